// TEST INFRASTRUCTURE — not product code.  Nothing under hammlet_b200/ may include, link or
// execute this file or anything it builds.
//
// ref_probe: a harness that #includes the UNMODIFIED reference headers where they lie
// (-I/root/reference/src; nothing is copied into this repository) and drives the reference's own
// classes on inputs we choose, dumping every intermediate the parity tests need.  It exists
// because the reference ships no tests or golden vectors: these dumps pin
// oracle/hammlet_oracle_impl.h (our restatement) and, through it, the CUDA path.
//
// Built by oracle/Makefile into oracle/_ref/ (git-ignored) in four flavours:
//   ref_probe     real_t=float   reference Trellis           (bit-for-bit what `hammlet` computes)
//   ref_probe64   real_t=double  reference Trellis           (the fp64 oracle of SURVEY.md §8c)
//   ref_probe_t / ref_probe64_t  same, but the 79-line Trellis container (src/Trellis.hpp) is
//                 replaced by an instrumented double with identical arithmetic that snapshots
//                 the forward rows before the backward pass overwrites and clears them
//                 (StateSequence/ForwardBackward.hpp:140-162 mutates, :162 clears).
//                 oracle/make_golden.py asserts, for every fixture it writes, that both flavours sample identical
//                 states and leave identical posteriors; tests/test_oracle_vs_golden.py::test_probe_flavours_agree
//                 repeats that check wherever oracle/_ref is built.
//
// The include order below follows src/main.cpp:2-15; other orders do not compile (circular
// headers).  Private members are reached with -fno-access-control, so no reference line changes.
//
// Usage: ref_probe <job-file>     (job-file format: see oracle/refprobe.py, which writes it)

#include <random>
#include <vector>
#include <string>
#include <sstream>
#include <fstream>
#include <iostream>
#include <iomanip>
#include <chrono>
#include <map>
#include <cstdio>
#include <cstdint>
#include <limits>

#ifndef PROBE_REAL
#define PROBE_REAL float
#endif
typedef PROBE_REAL real_t;          // identical re-declaration of src/includes.hpp:10 (or its fp64 variant)
typedef std::mt19937 rng_t;         // identical re-declaration of src/Distribution.hpp:15

struct ProbeLog {
	std::vector<double> rows;       // forward rows (B+1) x K as they stand when backward sampling starts
	size_t K = 0;
	bool armed = false;
	std::vector<double> uniforms;   // 53-bit uniforms in consumption order
} g_probe;

#ifdef PROBE_TRELLIS
// Instrumented stand-in for src/Trellis.hpp:8-76 (same public interface, same arithmetic, same
// std::discrete_distribution draw).  Only additions: snapshot + uniform log.
#define TRELLIS_HPP
class Trellis {
		std::vector<real_t> mVec;
		size_t mNrStates;
		rng_t& mRNG;
	public:
		Trellis( const Trellis& ) = delete;
		Trellis( rng_t& RNG ) : mNrStates( 2 ), mRNG( RNG ) {}
		Trellis( size_t nrStates, rng_t& RNG ) : mNrStates( nrStates ), mRNG( RNG ) {}
		real_t& operator()( size_t t, size_t d ) { return mVec[t * mNrStates + d]; }
		real_t& back( size_t d ) { return mVec[mVec.size() - mNrStates + d]; }
		void setNrStates( size_t K ) { mNrStates = K; }
		size_t size() const { return mVec.size() / mNrStates; }
		void push_back( const std::vector<real_t>& v ) { mVec.insert( mVec.end(), v.begin(), v.end() ); }
		size_t sample( size_t t ) const {
			if ( g_probe.armed ) {   // first draw of a backward pass: rows are still the forward rows
				g_probe.rows.assign( mVec.begin(), mVec.end() );
				g_probe.K = mNrStates;
				g_probe.armed = false;
			}
			rng_t peek = mRNG;
			g_probe.uniforms.push_back( std::generate_canonical<double, 53>( peek ) );
			std::discrete_distribution<size_t> dist( mVec.begin() + ( t * mNrStates ), mVec.begin() + ( ( t + 1 ) * mNrStates ) );
			size_t r = dist( mRNG );
			dist.reset();
			return r;
		}
		void reserve( size_t N ) { mVec.reserve( N * mNrStates ); }
		void clear() { mVec.clear(); g_probe.armed = true; }
};
#endif

#include "Tags.hpp"
#include "HMM.hpp"
#include "Parser.hpp"
#include "Emissions.hpp"
#include "Blocks.hpp"
#include "AutoPriors.hpp"
#include "Records.hpp"
#include "wavelet.hpp"
#include "StateSequence.hpp"
#include "Statistics.hpp"
#include "includes.hpp"
#include "utils.hpp"

// ---------------------------------------------------------------- job file + dumps

static std::map<std::string, std::vector<std::string>> g_job;

static void readJob( const char* path ) {
	std::ifstream f( path );
	if ( !f ) { throw std::runtime_error( std::string( "cannot read job file " ) + path ); }
	std::string line;
	while ( std::getline( f, line ) ) {
		std::istringstream ls( line );
		std::string key, tok;
		if ( !( ls >> key ) ) { continue; }
		std::vector<std::string> v;
		while ( ls >> tok ) { v.push_back( tok ); }
		g_job[key] = v;
	}
}
static bool has( const std::string& k ) { return g_job.count( k ) > 0; }
static const std::vector<std::string>& toks( const std::string& k ) {
	if ( !has( k ) ) { throw std::runtime_error( "job lacks key " + k ); }
	return g_job[k];
}
static std::string jstr( const std::string& k ) { return toks( k ).at( 0 ); }
static double jnum( const std::string& k, size_t i = 0 ) { return std::stod( toks( k ).at( i ) ); }
static size_t jint( const std::string& k, size_t i = 0 ) { return ( size_t ) std::stoull( toks( k ).at( i ) ); }

template<typename V>
static void dumpF64( const std::string& name, const V& v ) {
	std::vector<double> d( v.begin(), v.end() );
	std::ofstream o( jstr( "out" ) + "/" + name + ".f64", std::ios::binary );
	o.write( ( const char* ) d.data(), d.size() * sizeof( double ) );
}
template<typename V>
static void dumpI64( const std::string& name, const V& v ) {
	std::vector<int64_t> d( v.begin(), v.end() );
	std::ofstream o( jstr( "out" ) + "/" + name + ".i64", std::ios::binary );
	o.write( ( const char* ) d.data(), d.size() * sizeof( int64_t ) );
}

// std::streambuf over a raw float32 file that yields the values as text, "%.9g\n" each (9 significant digits
// round-trip a float), 65536 values per refill
class TextOfFloats : public std::streambuf {
		std::ifstream mFile;
		size_t mCount;
		std::vector<float> mVals;
		std::vector<char> mText;
	public:
		explicit TextOfFloats( const std::string& path ) : mFile( path, std::ios::binary | std::ios::ate ), mCount( 0 ) {
			if ( !mFile ) { throw std::runtime_error( "cannot read " + path ); }
			mCount = ( size_t ) mFile.tellg() / sizeof( float );
			mFile.seekg( 0 );
			mVals.resize( 65536 );
			mText.resize( 65536 * 20 );
		}
		size_t count() const { return mCount; }
	protected:
		int_type underflow() override {
			mFile.read( ( char* ) mVals.data(), mVals.size() * sizeof( float ) );
			const size_t n = ( size_t ) mFile.gcount() / sizeof( float );
			if ( n == 0 ) { return traits_type::eof(); }
			size_t len = 0;
			for ( size_t i = 0; i < n; ++i ) { len += ( size_t ) snprintf( mText.data() + len, 20, "%.9g\n", ( double ) mVals[i] ); }
			setg( mText.data(), mText.data(), mText.data() + len );
			return traits_type::to_int_type( *gptr() );
		}
};

typedef Statistics<IntegralArray, Normal> S;
typedef Blocks<BreakpointArray> B;
typedef Emissions<S, B> Y;

// ---------------------------------------------------------------- main

int main( int argc, const char* argv[] ) {
	try {
		if ( argc != 2 ) { throw std::runtime_error( "usage: ref_probe <job-file>" ); }
		readJob( argv[1] );
		const std::string mode = jstr( "mode" );

		// ---- load: data arrives as raw float64; it is rendered with 17 significant digits so that
		// the reference's own text front end (wavelet.hpp:131 `input >> v`) recovers exactly the same
		// real number whether real_t is float or double (inputs are float-representable).
		// multivariate jobs ("dims D"): the data file holds T*D values, position-major (wavelet.hpp:131-136)
		const size_t nrDataDim = has( "dims" ) ? jint( "dims" ) : 1;
		std::vector<real_t> inputValues;
		std::vector<SufficientStatistics<Normal>> stats;
		// "quiet 1" (the bench arm at 1e8..1e9 observations): no dumps of per-observation arrays
		const bool quiet = has( "quiet" ) && jint( "quiet" ) != 0;
		if ( has( "data32" ) ) {
			// large inputs: raw float32 on disk, rendered as text in 64k-value pieces by a stream buffer, so the
			// reference's text front end still does the parsing but nothing of size T besides its own arrays exists
			TextOfFloats tb( jstr( "data32" ) );
			std::istream text( &tb );
			MaxletTransform( text, inputValues, stats, nrDataDim, tb.count() + 1 );
		} else {
			std::vector<double> raw;
			{
				std::ifstream f( jstr( "data" ), std::ios::binary | std::ios::ate );
				if ( !f ) { throw std::runtime_error( "cannot read data" ); }
				size_t bytes = f.tellg();
				f.seekg( 0 );
				raw.resize( bytes / sizeof( double ) );
				f.read( ( char* ) raw.data(), bytes );
			}
			std::stringstream text;
			{
				char buf[64];
				for ( double v : raw ) { snprintf( buf, sizeof buf, "%.17g\n", v ); text << buf; }
			}
			MaxletTransform( text, inputValues, stats, nrDataDim, raw.size() + 1 );      // main.cpp:277
		}
		const size_t T = inputValues.size();
		if ( !quiet ) { dumpF64( "coeffs", inputValues ); }

		// noise estimate, main.cpp:303-311 (glue inside main(); restated, it is not callable)
		double stdEstimate = 0;
		size_t nrDetailCoeffs = 0;
		for ( size_t i = 1; i < inputValues.size(); i += 2 ) { stdEstimate += inputValues[i]; nrDetailCoeffs++; }
		stdEstimate /= nrDetailCoeffs;
		stdEstimate /= 0.797884560802865355879892119868763736951717262329869315331;
		dumpF64( "sigma_hat", std::vector<double> {stdEstimate} );

		HaarBreakpointWeights( inputValues );                                            // main.cpp:318
		const real_t weightMultiplier = has( "wmult" ) ? ( real_t ) jnum( "wmult" ) : ( real_t ) 1;
		for ( auto& w : inputValues ) { w *= weightMultiplier; }                        // main.cpp:332-334
		if ( !quiet ) { dumpF64( "weights", inputValues ); }
		if ( mode == "weights" ) { return 0; }

		S ia( stats, nrDataDim );                                                        // main.cpp:340
		B waveletBlocks( inputValues );                                                  // main.cpp:341
		Y y( ia, waveletBlocks );                                                        // main.cpp:343

		auto dumpBlocks = [&]( const std::string & prefix ) {
			std::vector<int64_t> st, en;
			std::vector<double> sm, sq;          // block-major, dimension-minor
			y.initForward();
			while ( y.next() ) {
				st.push_back( y.start() );
				en.push_back( y.end() );
				for ( size_t dim = 0; dim < nrDataDim; ++dim ) {
					sm.push_back( y.suffStat( dim ).sum() );
					sq.push_back( y.suffStat( dim ).sumSq() );
				}
			}
			dumpI64( prefix + "starts", st );
			dumpI64( prefix + "ends", en );
			dumpF64( prefix + "sum", sm );
			dumpF64( prefix + "sumsq", sq );
		};

		if ( mode == "blocks" ) {       // one dump per threshold listed
			const auto& th = toks( "thr" );
			for ( size_t i = 0; i < th.size(); ++i ) {
				y.createBlocks( ( real_t ) std::stod( th[i] ) );
				dumpBlocks( "t" + std::to_string( i ) + "_" );
			}
			return 0;
		}

		if ( mode == "autoprior" ) {    // AutoPriors.hpp:86-110 through the reference's own function
			std::vector<real_t> ap = autoPrior( ( real_t ) jnum( "s2" ), ( real_t ) jnum( "p" ), y, stdEstimate );
			dumpF64( "autoprior", ap );
			dumpBlocks( "ap_" );
			return 0;
		}

		const size_t P = jint( "K" );            // emission parameters ("-s C P D"); univariate: P == number of states
		rng_t RNG( jint( "seed" ) );
		Mapping mapping( nrDataDim, P, combinations );
		const size_t K = mapping.nrStates();     // main.cpp:137

		if ( mode == "sweep" ) {
			// ---- model objects exactly as main.cpp:154-166,354-362 builds them
			Transitions<DirichletVector> A( K, RNG );
			TransitionHyperParam<DirichletParamVector> tau_A( K, ( real_t ) jnum( "tau_A", 0 ), ( real_t ) jnum( "tau_A", 1 ) );
			Initial<Dirichlet> pi( K, RNG );
			InitialHyperParam<DirichletParam> tau_pi( K, ( real_t ) jnum( "tau_pi" ) );
			std::vector<std::vector<real_t>> thetaParams( P, std::vector<real_t> {
				( real_t ) jnum( "tau_theta", 0 ), ( real_t ) jnum( "tau_theta", 1 ), ( real_t ) jnum( "tau_theta", 2 ), ( real_t ) jnum( "tau_theta", 3 )} );
			ThetaHyperParam<NormalInverseGammaParam> tau_theta( thetaParams );
			Theta<NormalInverseGamma> theta( tau_theta, nrDataDim, combinations, RNG );
			// ---- overwrite the sampled values with the job's fixed parameters
			for ( size_t s = 0; s < P; ++s ) {
				theta.mParams[s].setValue( ( real_t ) jnum( "theta", 2 * s ), ( real_t ) jnum( "theta", 2 * s + 1 ) );
			}
			for ( size_t s = 0; s < K; ++s ) {
				pi.mValue.mProbs[s] = ( real_t ) jnum( "pi", s );
				for ( size_t j = 0; j < K; ++j ) { A( s, j ) = ( real_t ) jnum( "A", s * K + j ); }
			}
			const bool useSelf = jint( "self" ) != 0;
			const std::string method = jstr( "method" );
			const size_t nsweeps = has( "nsweeps" ) ? jint( "nsweeps" ) : 1;
			const bool dynamic = has( "dynamic" ) && jint( "dynamic" ) != 0;
			Records records( T, jstr( "out" ) + "/rec-", ".csv", K );
			records.setRecordStateSequence( true, true );
			records.setRecordTheta( true, true );
			records.setRecordBlocks( true, true );
			records.setRecordCompression( true, true );
			records.setRecordMarginals( true, true );
			records.setRecordSegments( true, true );

			if ( dynamic ) { y.createBlocks( theta ); } else { y.createBlocks( ( real_t ) jnum( "thr" ) ); }
			RNG.seed( jint( "seed" ) );
			std::vector<int64_t> allStates;
			std::vector<double> allUniforms, drawn;
			for ( size_t it = 0; it < nsweeps; ++it ) {
				if ( dynamic && it > 0 ) { y.createBlocks( theta ); }            // HMM.hpp:100-102
				rng_t clone = RNG;
				g_probe.armed = true;
				g_probe.uniforms.clear();
				size_t nb = 0;
				if ( method == "F" ) {
					StateSequence<ForwardBackward> q( RNG );
					q.sample( y, theta, tau_theta, A, tau_A, pi, tau_pi, mapping, records, true, useSelf );
					nb = q.size();
					if ( it == 0 ) { dumpI64( "states", q.states() ); }
					allStates.insert( allStates.end(), q.states().begin(), q.states().end() );
				} else {
					StateSequence<Mixture> q( RNG );
					q.sample( y, theta, tau_theta, A, tau_A, pi, tau_pi, mapping, records, true, useSelf );
					nb = y.nrBlocks();
				}
				std::vector<double> u( nb );
				for ( size_t b = 0; b < nb; ++b ) { u[b] = std::generate_canonical<double, 53>( clone ); }
				if ( it == 0 ) {
					dumpF64( "uniforms", u );
#ifdef PROBE_TRELLIS
					if ( method == "F" ) {
						dumpF64( "rows", g_probe.rows );
						dumpF64( "uniforms_seen", g_probe.uniforms );
					}
#endif
					dumpBlocks( "" );
					std::vector<double> pt, pa, pp;
					for ( size_t s = 0; s < P; ++s ) {
						const auto& po = tau_theta.posterior( s );
						pt.push_back( po.alpha() ); pt.push_back( po.beta() ); pt.push_back( po.mu0() ); pt.push_back( po.nu() );
					}
					for ( size_t s = 0; s < K; ++s ) {
						pp.push_back( tau_pi.posterior()[s] );
						for ( size_t j = 0; j < K; ++j ) { pa.push_back( tau_A.posterior()[s][j] ); }
					}
					dumpF64( "post_theta", pt );
					dumpF64( "post_A", pa );
					dumpF64( "post_pi", pp );
				}
				allUniforms.insert( allUniforms.end(), u.begin(), u.end() );
				// parameter draws in the order of HMM.hpp:112-116
				theta.sample( tau_theta );
				pi.sample( tau_pi );
				A.sample( tau_A );
				records.record( theta );
				for ( size_t s = 0; s < P; ++s ) { drawn.push_back( theta.value()[s].mean() ); drawn.push_back( theta.value()[s].var() ); }
				for ( size_t s = 0; s < K; ++s ) { drawn.push_back( pi.valueVector()[s] ); }
				for ( size_t s = 0; s < K; ++s ) for ( size_t j = 0; j < K; ++j ) { drawn.push_back( A( s, j ) ); }
			}
			dumpF64( "drawn", drawn );               // per sweep: P*(mean,var), K pi, K*K A
			dumpF64( "all_uniforms", allUniforms );
			dumpI64( "all_states", allStates );
			return 0;
		}

		if ( mode == "bench" ) {
			// Reference arm of bench.py: the reference's own setup (auto priors) and its own sampleHMM
			// loop (HMM.hpp:60-125), FBG, dynamic blocks, no recording; sweeps timed with std::chrono.
			Transitions<DirichletVector> A( K, RNG );
			TransitionHyperParam<DirichletParamVector> tau_A( K, ( real_t ) 0.5, ( real_t ) 0.5 );
			Initial<Dirichlet> pi( K, RNG );
			InitialHyperParam<DirichletParam> tau_pi( K, ( real_t ) 0.5 );
			std::vector<real_t> ap = autoPrior( ( real_t ) 0.2, ( real_t ) 0.9, y, stdEstimate );
			std::vector<std::vector<real_t>> thetaParams( P, ap );
			ThetaHyperParam<NormalInverseGammaParam> tau_theta( thetaParams );
			Theta<NormalInverseGamma> theta( tau_theta, nrDataDim, combinations, RNG );
			theta.sample( tau_theta );
			pi.sample( tau_pi );
			A.sample( tau_A );
			Records records( T, jstr( "out" ) + "/bench-", ".csv", K );
			records.setRecordMarginals( false, true );
			const std::string method = has( "method" ) ? jstr( "method" ) : "F";
			auto run = [&]( size_t n ) {
				if ( method == "F" ) {
					StateSequence<ForwardBackward> q( RNG );
					sampleHMM( y, q, theta, tau_theta, A, tau_A, pi, tau_pi, mapping, n, 0, records, true, true );
				} else {
					StateSequence<Mixture> q( RNG );
					sampleHMM( y, q, theta, tau_theta, A, tau_A, pi, tau_pi, mapping, n, 0, records, true, true );
				}
			};
			run( jint( "burn" ) );
			const size_t reps = jint( "reps" ), per = jint( "timed" );
			std::vector<double> secs;
			for ( size_t r = 0; r < reps; ++r ) {
				auto t0 = std::chrono::steady_clock::now();
				run( per );
				auto t1 = std::chrono::steady_clock::now();
				secs.push_back( std::chrono::duration<double>( t1 - t0 ).count() );
			}
			// block count of the last sweep's structure
			size_t nb = 0;
			y.createBlocks( theta );
			y.initForward();
			while ( y.next() ) { nb++; }
			dumpF64( "bench_secs", secs );
			dumpI64( "bench_blocks", std::vector<int64_t> {( int64_t ) nb} );
			return 0;
		}
		throw std::runtime_error( "unknown mode " + mode );
	} catch ( std::exception& e ) {
		std::cerr << "[ref_probe ERROR] " << e.what() << std::endl;
		return 1;
	}
}
