// hammlet — command line of the B200-native build.
//
// Same flags, defaults, token-group semantics, sampling-scheme interpreter, output files and error
// behaviour as the reference's src/main.cpp (flags :33-63, scheme :384-455, errors :467-474).  The data
// are parsed on the host, handed once to the device (hml_load_f32) and every sweep of the scheme runs
// there; see StateSequence.hpp.  Additional flags (none collides with the reference's):
//   -device N            CUDA device index (default 0)
//   -devices a b c ...   split the sequence into contiguous segments over the listed devices (also `a,b,c`): one
//                        process per device is forked after the input has been parsed, every process runs the same
//                        chain on the same parameters, the scan carries travel between the GPUs,
//                        process 0 writes the files — the same files a single device writes.
//   -F|-input-format X   auto (default: gzip'd text is recognised by its magic number, anything else is text), text, gz,
//                        f32 (raw little-endian float32: 4 bytes per value instead of ~9 of text)
//   -replay              draw the per-block uniforms from the shared mt19937 exactly as the reference
//                        does (draw-for-draw comparable runs; slower: one host round trip per sweep)
//   -timing              print sweeps/s per run token to stderr
#include <signal.h>
#include <sys/prctl.h>
#include <sys/wait.h>
#include <unistd.h>

#include <chrono>
#include <ctime>

#include "StateSequence.hpp"

// ---- one process per device (-devices): children of process 0, wired by pipes that carry the communicator id
static std::vector<pid_t> g_children;
static void killChildren() {
  for (pid_t p : g_children) kill(p, SIGTERM);
}
static void onChild(int) {  // a rank that dies would leave the others waiting in a collective
  int status = 0;
  pid_t p;
  while ((p = waitpid(-1, &status, WNOHANG)) > 0) {
    const bool ok = WIFEXITED(status) && WEXITSTATUS(status) == 0;
    bool mine = false;
    for (pid_t& c : g_children)
      if (c == p) {
        c = ok ? 0 : -1;
        mine = true;
      }
    if (mine && !ok) {
      const char msg[] = "\n[ERROR] A device process failed!\nTerminating HaMMLET. The rest is silence.\n";
      if (write(2, msg, sizeof(msg) - 1) < 0) {}
      for (pid_t c : g_children)
        if (c > 0) kill(c, SIGTERM);
      _exit(1);
    }
  }
}

using std::cerr;
using std::cout;
using std::endl;
using std::flush;
using std::string;
using std::vector;

static const char* kHelp =
    "hammlet (B200 build) - Bayesian HMM segmentation with dynamic wavelet compression\n"
    "  -f|-input-file FILE        whitespace-separated values (default: standard input)\n"
    "  -o|-output-pattern P S     output files are P<type>S (default: hammlet- .csv, or from -f)\n"
    "  -O|-output-data ...        marginals|M sequences|S parameters|P blocks|B compression|C segments|G\n"
    "  -w|-overwrite              allow overwriting existing files\n"
    "  -s|-states K               number of states (default 3); `C p [d]` = p parameters, d dimensions\n"
    "  -e|-emissions normal s2 p  automatic-prior tuning (default: normal 0.2 0.9)\n"
    "  -a|-auto-priors            derive emission priors from the data (required)\n"
    "  -t|-transitions a [b]      Dirichlet hyper-parameters: off-diagonal a, diagonal b (default 0.5 0.5)\n"
    "  -S|-no-self-transitions    ignore within-block self transitions\n"
    "  -I|-initial-dist a         Dirichlet hyper-parameter of the initial distribution (default 0.5)\n"
    "  -R|-random-seed N          seed (default: time)\n"
    "  -i|-iterations ...         scheme, e.g. `M 500 0 S P F 200 0 F 300 3` (M|F iterations thinning; P S D)\n"
    "  -m|-weight-multiplier x    multiply breakpoint weights (default 1)\n"
    "  -v -g -h                   verbose, print parsed arguments, this help\n"
    "  -device N  -replay  -timing   see the header of hammlet_main.cpp\n"
    "  -devices a b ...           split the sequence over several GPUs (one process each), same output files\n"
    "  -F|-input-format X         auto|text|gz|f32: gzip'd text (e.g. samToCounts' *-count.csv.gz) is recognised by\n"
    "                             itself; f32 = raw little-endian float32\n";

// everything but the wait for the other device processes: when this returns, the sequence (and with it the NCCL
// communicator, whose teardown is collective) has been released
static int hammletMain(int argc, const char* argv[]) {
  {
    Parser args(argc, argv);
    args.registerFlags({"-v", "-verbose"});
    args.registerFlags({"-g", "-arguments"});
    args.registerFlags({"-h", "-help", "--help"});
    args.registerFlags({"-f", "-input-file"});
    args.registerFlags({"-o", "-output-pattern"}, "hammlet- .csv");
    args.registerFlags({"-O", "-output-data"}, "marginals");
    args.registerFlags({"-w", "-overwrite"});
    args.registerFlags({"-s", "-states"}, "3");
    args.registerFlags({"-e", "-emissions"}, "normal 0.2 0.9");
    args.registerFlags({"-a", "-auto-priors"});
    args.registerFlags({"-t", "-transitions"}, "0.5 0.5");
    args.registerFlags({"-S", "-no-self-transitions"});
    args.registerFlags({"-I", "-initial-dist"}, "0.5");
    args.registerFlags({"-R", "-random-seed"}, std::to_string(time(0)));
    args.registerFlags({"-i", "-iterations"}, "M 500 0 S P F 200 0 F 300 3");
    args.registerFlags({"-m", "-weight-multiplier"}, "1");
    args.registerFlags({"-device"}, "0");
    args.registerFlags({"-devices"});
    args.registerFlags({"-F", "-input-format"}, "auto");
    args.registerFlags({"-replay"});
    args.registerFlags({"-timing"});
    args.parseArgs();

    if (args.isSet("-g")) args.print();
    bool verbose = args.isSet("-v");
    const bool overwrite = args.isSet("-w");
    if (args.isSet("-h")) {
      cout << endl << kHelp << endl;
      return 0;
    }

    // ---- output pattern: without -o, `-f name.ext` gives `name-` + `.ext`
    string outputPrefix, outputSuffix;
    if (!args.isSet("-o") && args.isSet("-f")) {
      const string filename = args.parse<string>("-f");
      const size_t dot = filename.find_last_of(".");
      outputPrefix = filename.substr(0, dot) + "-";
      outputSuffix = filename.substr(dot);
    } else {
      outputPrefix = args.parse<string>("-o", 0);
      outputSuffix = args.parse<string>("-o", 1);
    }

    const size_t rng_seed = args.parse<size_t>("-R", 0);
    rng_t RNG(rng_seed);

    // ---- states and mapping
    size_t nrParams, nrDataDim = 1;
    MappingType mappingType = combinations;
    if (args.nrTokens("-s") == 1) {
      nrParams = args.parse<size_t>("-s", 0);
    } else {
      mappingType = args.parse<MappingType>("-s", 0);
      nrParams = args.parse<size_t>("-s", 1);
      if (args.nrTokens("-s") >= 3) nrDataDim = args.parse<size_t>("-s", 2);
    }
    Mapping mapping(nrDataDim, nrParams, mappingType);
    const size_t nrStates = mapping.nrStates();

    // ---- transitions: first token off-diagonal, second (optional) diagonal
    const real_t trans = args.parse<real_t>("-t", 0);
    const real_t selfTrans = args.nrTokens("-t") > 1 ? args.parse<real_t>("-t", 1) : trans;
    Transitions<DirichletVector> A(nrStates, RNG);
    TransitionHyperParam<DirichletParamVector> tau_A(nrStates, trans, selfTrans);
    const bool useSelfTrans = !args.isSet("-S");

    const real_t initialAlpha = args.parse<real_t>("-I", 0);
    Initial<Dirichlet> pi(nrStates, RNG);
    InitialHyperParam<DirichletParam> tau_pi(nrStates, initialAlpha);

    const real_t weightMultiplier = args.parse<real_t>("-m");

    vector<vector<real_t>> thetaParams;
    if (args.isSet("-a")) {
      const vector<real_t> thp = args.parseVector<real_t>("-e", 1, 3);
      for (size_t i = 0; i < nrParams; ++i) thetaParams.push_back(thp);
    } else {
      throw std::runtime_error("Manual theta priors not implemented, use -a!");
    }

    if (verbose) {
      cout << "Data dimensions: " << nrDataDim << endl;
      cout << "Emission distributions: " << nrParams << endl;
      cout << "States: " << nrStates << endl;
      cout << "Sampling scheme: " << hammlet::concat(args.tokens("-i"), " ") << endl;
      cout << "Random seed: " << rng_seed << endl;
    }

    Parser outputArgs = args.subparser("-output-data");
    outputArgs.registerFlags({"M", "marginals"});
    outputArgs.registerFlags({"S", "sequences"});
    outputArgs.registerFlags({"P", "parameters"});
    outputArgs.registerFlags({"B", "blocks"});
    outputArgs.registerFlags({"C", "compression"});
    outputArgs.registerFlags({"D", "mapping"});
    outputArgs.registerFlags({"G", "segments"});
    outputArgs.parseArgs();

    // ---- devices: one (-device) or several (-devices: the sequence is split into contiguous segments)
    vector<int> devices{args.parse<int>("-device")};
    if (args.isSet("-devices")) {
      devices.clear();
      for (const string& tok : args.tokens("-devices")) {
        std::stringstream ss(tok);
        string item;
        while (std::getline(ss, item, ',')) {
          size_t used = 0;
          int d = -1;
          try {
            d = std::stoi(item, &used);
          } catch (std::exception&) {
            used = 0;
          }
          if (item.empty() || used != item.size() || d < 0) throw std::runtime_error("Cannot parse device list \"" + tok + "\"!");
          devices.push_back(d);
        }
      }
      if (devices.empty()) throw std::runtime_error("Flag -devices needs at least one device index!");
    }
    const int world = (int)devices.size();

    // ---- load: parse on the host (before any process is forked and before CUDA is touched), transform on the device
    const fastparse::Format inputFormat = fastparse::formatFromName(args.parse<string>("-F"));
    vector<float> values;
    if (args.isSet("-f")) {
      const vector<string> files = args.parseVector<string>("-f");
      if (files.size() > 1) throw std::runtime_error("Coefficient array must be empty!");  // as wavelet.hpp:111-113
      if (verbose) cout << "Reading " + files[0] << endl << flush;
      values = readValuesFile(files[0], inputFormat);
    } else {
      if (verbose) cout << "Reading from standard input" << endl << flush;
      values = readValues(std::cin, 0, inputFormat);
    }
    if (nrDataDim <= 0) throw std::runtime_error("Number of dimensions must be positive!");

    int rank = 0;
    uint8_t commId[HML_UNIQUE_ID_BYTES] = {0};
    if (world > 1) {
      cout << flush;
      vector<int> toChild(world, -1);
      int fromParent = -1;
      signal(SIGCHLD, onChild);
      for (int r = 1; r < world; ++r) {
        int fds[2];
        if (pipe(fds) != 0) throw std::runtime_error("Cannot create a pipe for a device process!");
        const pid_t pid = fork();
        if (pid < 0) throw std::runtime_error("Cannot fork a device process!");
        if (pid == 0) {  // child: rank r
          prctl(PR_SET_PDEATHSIG, SIGTERM);
          signal(SIGCHLD, SIG_DFL);
          g_children.clear();
          for (int q = 1; q < r; ++q) close(toChild[q]);
          close(fds[1]);
          fromParent = fds[0];
          rank = r;
          break;
        }
        close(fds[0]);
        toChild[r] = fds[1];
        g_children.push_back(pid);
      }
      if (rank == 0) {
        if (hml_comm_unique_id(commId) != HML_OK) throw std::runtime_error(hml_last_error(nullptr));
        for (int r = 1; r < world; ++r) {
          if (write(toChild[r], commId, sizeof(commId)) != (ssize_t)sizeof(commId))
            throw std::runtime_error("Cannot reach a device process!");
          close(toChild[r]);
        }
      } else {
        size_t got = 0;
        while (got < sizeof(commId)) {
          const ssize_t n = read(fromParent, commId + got, sizeof(commId) - got);
          if (n <= 0) throw std::runtime_error("Lost the connection to device process 0!");
          got += (size_t)n;
        }
        close(fromParent);
      }
    }
    const bool lead = rank == 0;  // the process that talks and writes
    if (!lead) verbose = false;

    DeviceSequence sequence(devices[rank]);
    if (world > 1)
      sequence.loadSegment(values, (float)weightMultiplier, rank, world, commId, nrDataDim);
    else
      sequence.load(values, (float)weightMultiplier, nrDataDim);
    vector<float>().swap(values);
    if (verbose) cout << "Output will be written to " + outputPrefix + "*" + outputSuffix << endl << flush;
    const size_t T = sequence.size();
    if (verbose) cout << "Number of data points: " + std::to_string(T) << endl << flush;
    const double stdEstimate = sequence.noiseStdev();
    if (verbose) cout << "Calculating Haar breakpoint weights" << endl << flush;

    Records records(T, outputPrefix, outputSuffix, nrStates);
    records.setMute(!lead);
    records.setRecordStateSequence(outputArgs.isSet("sequences"), overwrite);
    records.setRecordTheta(outputArgs.isSet("parameters"), overwrite);
    records.setRecordBlocks(outputArgs.isSet("blocks"), overwrite);
    records.setRecordCompression(outputArgs.isSet("compression"), overwrite);
    records.setRecordMarginals(outputArgs.isSet("marginals"), overwrite);
    records.setRecordSegments(outputArgs.isSet("segments"), overwrite);
    // with -O M alone (the default) the marginals accumulate on the device; any per-iteration host output
    // (sequences, blocks, segments) keeps them in the host-side StateMarginals
    DeviceMarginals deviceMarginals(sequence, nrStates);
    records.setMarginalsSink(&deviceMarginals);

    typedef Statistics<IntegralArray, Normal> S;
    typedef Blocks<BreakpointArray> B;
    S ia(sequence, nrDataDim);
    B waveletBlocks(sequence);
    Emissions<S, B> y(ia, waveletBlocks);

    thetaParams[0] = autoPrior(thetaParams[0][0], thetaParams[0][1], y, stdEstimate);
    for (auto& param : thetaParams) param = thetaParams[0];
    ThetaHyperParam<NormalInverseGammaParam> tau_theta(thetaParams);
    Theta<NormalInverseGamma> theta(tau_theta, nrDataDim, mappingType, RNG);

    // ---- sampling scheme
    size_t nrTokens = 0;
    for (const string& c : args.tokens("-i"))
      if (c != "P" && c != "S" && c != "D") nrTokens++;
    if (nrTokens % 3 != 0) throw std::runtime_error("Parameters for -i, excluding \"P\", \"S\" and \"D\", must be multiples of 3!");
    nrTokens = args.nrTokens("-i");

    const bool replay = args.isSet("-replay");
    const bool timing = args.isSet("-timing");
    bool samplePrior = true, dynamic = true;
    if (verbose) cout << "Setting block structure to dynamic" << endl << flush;
    for (size_t i = 0; i < nrTokens;) {
      const string method = args.parse<string>("-i", i);
      if (samplePrior) {
        if (verbose) cout << "Sampling prior" << endl << flush;
        theta.sample(tau_theta);
        pi.sample(tau_pi);
        A.sample(tau_A);
        samplePrior = false;
      }
      if (method == "P") {
        samplePrior = true;
        i++;
        continue;
      } else if (method == "S") {
        if (verbose) cout << "Setting block structure to static" << endl << flush;
        y.createBlocks(theta);
        dynamic = false;
        i++;
        continue;
      } else if (method == "D") {
        if (verbose) cout << "Setting block structure to dynamic" << endl << flush;
        dynamic = true;
        i++;
        continue;
      }
      if (i + 2 >= nrTokens) throw std::runtime_error("Incomplete command line for -i!");
      const size_t iterations = args.parse<size_t>("-i", i + 1);
      const size_t thinning = args.parse<size_t>("-i", i + 2);
      i += 3;
      const auto t0 = std::chrono::steady_clock::now();
      if (method == "F") {
        if (verbose) cout << "Sampling Forward-Backward" << endl << flush;
        StateSequence<ForwardBackward> q(RNG);
        q.setReplay(replay);
        sampleHMM(y, q, theta, tau_theta, A, tau_A, pi, tau_pi, mapping, iterations, thinning, records, dynamic, useSelfTrans);
      } else if (method == "M") {
        if (verbose) cout << "Sampling mixture" << endl << flush;
        StateSequence<Mixture> q(RNG);
        q.setReplay(replay);
        sampleHMM(y, q, theta, tau_theta, A, tau_A, pi, tau_pi, mapping, iterations, thinning, records, dynamic, useSelfTrans);
      } else {
        throw std::runtime_error("Unknown sampling type " + method + "!");
      }
      if (timing && lead) {
        const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        cerr << "[timing] " << method << " " << iterations << " sweeps in " << s << " s = " << (s > 0 ? iterations / s : 0)
             << " sweeps/s" << endl;
      }
    }
    records.close();  // writes the marginals (the reference does this in ~Records)
    if (verbose) cout << "Exit HaMMLET" << endl << flush;
    return 0;
  }
}

int main(int argc, const char* argv[]) {
  try {
    const int rc = hammletMain(argc, argv);
    // the other device processes have nothing left to do: collect them (only now — the communicator is torn down by all
    // ranks together, so process 0 must have left hammletMain before it waits for anybody)
    signal(SIGCHLD, SIG_DFL);
    for (pid_t p : g_children) {
      int status = 0;
      if (p < 0 || (p > 0 && waitpid(p, &status, 0) == p && !(WIFEXITED(status) && WEXITSTATUS(status) == 0)))
        throw std::runtime_error("A device process failed!");
    }
    g_children.clear();
    return rc;
  } catch (std::exception& e) {
    signal(SIGCHLD, SIG_DFL);
    killChildren();
    cout << flush;
    cerr << endl << flush << "[ERROR] " << e.what() << endl;
    cerr << "Terminating HaMMLET. The rest is silence." << endl << flush;
    return 1;
  }
}
