// hammlet_b200 host side — tiny driver that replays recorded iterations through Records, used by
// tests/test_host_records.py to check the record files byte for byte against the reference's.
// Input (stdin): T K nsweeps, then per sweep: B, B block sizes, B states.  Output: files PREFIX*SUFFIX.
// With a third argument "runs" the blocks are first merged into equal-state runs and handed over with
// Records::recordRun, the way the device delivers recorded iterations (hml_get_segments); block sizes are then not
// written (a run does not know them).
#include "Records.hpp"

int main(int argc, const char* argv[]) {
  try {
    if (argc != 3 && argc != 4) throw std::runtime_error("usage: records_tool PREFIX SUFFIX [runs] < iterations");
    const bool runs = argc == 4 && std::string(argv[3]) == "runs";
    size_t T, K, n;
    std::cin >> T >> K >> n;
    Records rec(T, argv[1], argv[2], K);
    rec.setRecordStateSequence(true, true);
    rec.setRecordBlocks(!runs, true);
    rec.setRecordCompression(true, true);
    rec.setRecordMarginals(true, true);
    rec.setRecordSegments(true, true);
    for (size_t it = 0; it < n; ++it) {
      size_t B;
      std::cin >> B;
      std::vector<size_t> sizes(B), states(B);
      for (auto& v : sizes) std::cin >> v;
      for (auto& v : states) std::cin >> v;
      if (!runs) {
        for (size_t b = 0; b < B; ++b) rec.record(states[b], sizes[b]);
      } else {
        size_t b = 0;
        bool first = true;
        while (b < B) {
          size_t e = b, N = 0;
          while (e < B && states[e] == states[b]) N += sizes[e++];
          rec.recordRun(states[b], N, first ? B : 0);  // the block count of the iteration rides on the first run
          first = false;
          b = e;
        }
      }
    }
    rec.close();
    return 0;
  } catch (std::exception& e) {
    std::cerr << "[ERROR] " << e.what() << std::endl;
    return 1;
  }
}
