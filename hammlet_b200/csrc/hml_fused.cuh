// hammlet_b200 — the whole Gibbs sweep in ONE persistent kernel, for block structures of up to kFusedMaxTiles tiles
// (65 536 blocks) and K <= 8, with the conjugate updates and the parameter draws on the device.
//
// Reference path: one iteration of sampleHMM, HMM.hpp:99-121 — createBlocks(theta), StateSequence<ForwardBackward>::
// sample (StateSequence/ForwardBackward.hpp:16-213), theta.sample, pi.sample, A.sample (Theta.hpp:203-211,
// Initial.hpp:35-40, Transitions.hpp:75-79 with Conjugate.hpp:120-205 and Distribution.hpp:76-139).
//
// Why: the multi-kernel sweep (hml_sweep_impl.cuh) is 13 dependent launches and one host round trip per sweep.  At
// 1e9 observations that overhead hides behind 240 us of work; at 1e6 (1.3 k blocks) or 1e7 observations (15 k blocks)
// every kernel sits on its 7-17 us floor and the sweep costs 65-85 us of which the GPU works a fraction.  Here a
// cooperative grid of one CTA per quarter tile (256 blocks; at most one CTA per SM) keeps everything of those blocks —
// emission terms, chunk operators, forward rows, backward maps, states — in that CTA; the four CTAs of a tile combine
// their chunk operators and chunk maps at a tile-local barrier, the phases of a sweep are separated by six grid-wide
// barriers instead of kernel boundaries, and the kernel loops over sweeps: the model of sweep i+1 is drawn by CTA 0
// from the statistics of sweep i (Philox streams) while the other CTAs wait at the barrier, so nothing returns to the
// host until the requested sweeps are done or the kernel stands down (the threshold left the candidate list, the block
// arrays are too small, a forward sum vanished: the host then handles that one sweep through the multi-kernel path).
//
// Per-block arrays, layouts and numerics are those of the multi-kernel path (same Layout::perm, same helpers), so
// every getter of the C ABI (states, runs, marginals, blocks) works on the result, and under uniform replay the
// sampled states are the reference's.
#pragma once
#include <cooperative_groups.h>

#include "hml_sweep_impl.cuh"

namespace hml {

namespace cg = cooperative_groups;

// ---- Philox-driven draws for the parameter phase: every draw owns a stream, successive uniforms advance its counter
struct DrawStream {
  unsigned long long seed, sweep;
  uint32_t stream;
  unsigned long long ctr;
  // four uniforms in (0, 1] from one Philox call (24 bits each: the draws are real_t = float in the reference,
  // std::gamma_distribution<float> / std::normal_distribution<float>)
  __device__ __forceinline__ void next4(float (&u)[4]) {
    uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)sweep, (uint32_t)(sweep >> 32) ^ (stream << 24)};
    ++ctr;
    Philox::run(c, (uint32_t)seed, (uint32_t)(seed >> 32));
#pragma unroll
    for (int k = 0; k < 4; ++k) u[k] = (float)((c[k] >> 8) + 1u) * (1.0f / 16777216.0f);
  }
  // The transcendental functions are float (a tenth of the instructions of their double versions: the parameter phase is
  // one warp on the critical path of every sweep); the value d * v is formed in double.
  __device__ __forceinline__ float normal_from(float u1, float u2) {  // Box-Muller
    return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
  }
  __device__ __forceinline__ float normal() {
    float u[4];
    next4(u);
    return normal_from(u[0], u[1]);
  }
  // Gamma(alpha, 1): Marsaglia & Tsang 2000 (the method libstdc++'s gamma_distribution uses) with its squeeze test,
  // boosted for alpha < 1; one Philox call per trial
  __device__ double gamma(double alpha) {
    const double a = alpha < 1.0 ? alpha + 1.0 : alpha;
    const double d = a - 1.0 / 3.0;
    const float c = rsqrtf((float)(9.0 * d));
    double g = 0.0;
    float u[4] = {1.f, 1.f, 1.f, 1.f};
    for (int tries = 0; tries < 1000; ++tries) {
      next4(u);
      const float x = normal_from(u[0], u[1]);
      const float v1 = 1.0f + c * x;
      if (v1 <= 0.0f) continue;
      const double v = (double)v1 * (double)v1 * (double)v1;
      const float x2 = x * x;
      // log u < x^2 / 2 + d (1 - v + log v): for large d the bracket is a small difference, taken in double
      if (u[2] < 1.0f - 0.0331f * x2 * x2 ||
          (double)logf(u[2]) < 0.5 * (double)x2 + d * (1.0 - v + (double)log1pf((float)(v - 1.0)))) {
        g = d * v;
        break;
      }
    }
    if (alpha < 1.0) g *= (double)powf(u[3], (float)(1.0 / alpha));
    return g;
  }
};

// theta, pi, A of the next sweep from the statistics of this one, by the first 64 + K*K threads of one CTA.
// Conjugate.hpp:120-205 in real_t = float like the reference; posterior := prior after every draw (Theta.hpp:209).
// s_g: K*K + 2*K floats of shared memory.
template <int KP>
__device__ void chain_sample_params(ChainDev* ch, const unsigned long long* out_u64, const double* out_f64, float* s_g) {
  const int K = ch->K, tid = threadIdx.x;
  DrawStream rs{ch->seed, ch->sweep, 0u, 0ull};
  float* g_pi = s_g;               // K
  float* g_A = s_g + KP;           // K*K
  float* s_var = s_g + KP + KP * KP;  // K
  if (tid < K) {                   // ---- theta_k: NIG posterior, then var = 1 / Gamma(alpha, 1 / beta), mean ~ N(mu0, sqrt(var / nu))
    float alpha = ch->prior_theta[tid][0], beta = ch->prior_theta[tid][1], mu0 = ch->prior_theta[tid][2], nu = ch->prior_theta[tid][3];
    const unsigned long long n = __ldcg(out_u64 + tid);
    if (n > 0) {
      const float sum = (float)__ldcg(out_f64 + tid), sumSq = (float)__ldcg(out_f64 + KP + tid);
      const double N = (double)n;
      const float xbar = (float)(sum / N);
      float ssN = (float)((sum * sum) / N);
      if (ssN > sumSq) ssN = sumSq;
      const float a2 = (float)(alpha + N / 2.0);
      const float b2 = (float)(beta + ((sumSq + (N * nu / (N + nu)) * ((xbar - mu0) * (xbar - mu0))) - ssN) / 2.0);
      const float m2 = (float)((nu * mu0 + sum) / (nu + N));
      const float n2 = (float)(nu + N);
      alpha = a2; beta = b2; mu0 = m2; nu = n2;
    }
    rs.stream = 2u + (uint32_t)tid;
    const float gam = (float)(rs.gamma((double)alpha) / (double)beta);
    const float var = 1.0f / gam;
    const float mean = mu0 + sqrtf(var / nu) * rs.normal();
    if (!(var > 0.0f) || !isfinite(var) || !isfinite(mean)) ch->phase_abort[6] = kChainNumeric;  // Observation.hpp:177-179
    ch->mean[tid] = (double)mean;
    ch->var[tid] = (double)var;
    // what the kernels use of theta_k, once per sweep instead of once per CTA (make_model)
    ch->inv2var[tid] = 1.0 / (2.0 * (double)var);
    ch->lognorm[tid] = log(sqrt((double)var)) + (double)mean * (double)mean / (2.0 * (double)var);
    s_var[tid] = var;
  }
  if (tid >= 32 && tid < 32 + K) {  // ---- pi ~ Dirichlet(alpha_I + occupancy) (FB.hpp:211: the quirk of A11)
    const int s = tid - 32;
    rs.stream = 2u + (uint32_t)KP + (uint32_t)s;
    g_pi[s] = (float)rs.gamma((double)(ch->prior_pi + (float)__ldcg(out_u64 + s)));
  }
  if (tid >= 64 && tid < 64 + K * K) {  // ---- rows of A ~ Dirichlet(prior row + transition counts)
    const int i = (tid - 64) / K, j = (tid - 64) % K;
    rs.stream = 2u + 2u * (uint32_t)KP + (uint32_t)(i * KP + j);
    const float prior = i == j ? ch->prior_self : ch->prior_trans;
    g_A[i * KP + j] = (float)rs.gamma((double)(prior + (float)__ldcg(out_u64 + KP + i * KP + j)));
  }
  __syncthreads();
  if (tid < K) {
    float s = 0.f;
    for (int j = 0; j < K; ++j) s += g_A[tid * KP + j];
    for (int j = 0; j < K; ++j) ch->A[tid * K + j] = (double)(g_A[tid * KP + j] / s);
    ch->loga[tid] = ch->use_self ? log((double)(g_A[tid * KP + tid] / s)) : 0.0;  // FB.hpp:47-50
  }
  if (tid == 32) {
    float s = 0.f;
    for (int j = 0; j < K; ++j) s += g_pi[j];
    for (int j = 0; j < K; ++j) ch->pi[j] = (double)(g_pi[j] / s);
    float mv = s_var[0];
    for (int j = 1; j < K; ++j) mv = fminf(mv, s_var[j]);
    ch->thr = sqrtf(2.0f * logf((float)ch->T) * mv);  // BreakpointArray.hpp:195-199, Theta.hpp:226-234
    ch->sweep += 1ull;
  }
  __syncthreads();
}

// stand-alone: the parameter phase after a sweep of the multi-kernel path
template <int KP>
__global__ void __launch_bounds__(kFusedThreads) k_chain_params(ChainDev* ch, const unsigned long long* out_u64, const double* out_f64) {
  __shared__ float s_g[KP * KP + 2 * KP];
  chain_sample_params<KP>(ch, out_u64, out_f64, s_g);
}

// What other CTAs of the same launch wrote is read from L2 (ld.global.cg): the L1 of this SM may still hold the line as
// it was a sweep ago.
template <int KP>
__device__ __forceinline__ void load_op_cg(OpVals<KP>& o, const double* M, const int* X) {
#pragma unroll
  for (int k = 0; k < KP * KP; ++k) o.m[k] = __ldcg(M + k);
#pragma unroll
  for (int k = 0; k < KP; ++k) o.x[k] = __ldcg(X + k);
}
template <int KP>
__device__ __forceinline__ Map<KP> load_map_cg(const uint8_t* p) {
  Map<KP> r;
  const unsigned long long* q = reinterpret_cast<const unsigned long long*>(p);
#pragma unroll
  for (int i = 0; i < Map<KP>::W; ++i) r.w[i] = __ldcg(q + i);
  return r;
}

// thread i < KP fills row i of the model from the chain (the derived values come with it, chain_derive)
template <int KP>
__device__ __forceinline__ void model_from_chain(const ChainDev* ch, ModelDev<KP>& m, int i) {
  const int K = ch->K;
  if (i == 0) {
    m.K = K;
    m.use_self = ch->use_self;
  }
  const bool on = i < K;
  m.mean[i] = on ? __ldcg(&ch->mean[i]) : 0.0;
  m.inv2var[i] = on ? __ldcg(&ch->inv2var[i]) : 0.0;
  m.lognorm[i] = on ? __ldcg(&ch->lognorm[i]) : 0.0;
  m.loga[i] = on ? __ldcg(&ch->loga[i]) : 0.0;
  m.pi[i] = on ? __ldcg(&ch->pi[i]) : 0.0;
  for (int j = 0; j < KP; ++j) m.A[i][j] = (on && j < K) ? __ldcg(&ch->A[i * K + j]) : 0.0;
}

// as block_sums, with the block list read from L2 (the scatter phase of other CTAs wrote it)
__device__ __forceinline__ void block_sums_cg(const SweepBuffers& buf, uint64_t b, uint32_t& n, double& sx, double& sq) {
  const uint32_t s = __ldcg(buf.starts + b), e = __ldcg(buf.starts + b + 1);
  const double2 ps = __ldcg(buf.spq + b), pe = __ldcg(buf.spq + b + 1);
  range_sums_from(buf, ps, pe, s, e, sx, sq);
  n = e - s;
}

// ---- barriers of the persistent kernel.  All CTAs of a cooperative launch are resident, so spinning is safe.  The
// counters only grow (the host zeroes them before a launch): the k-th barrier of a group of n CTAs waits for k * n.
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void spin_barrier(unsigned* ctr, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1u);
    while (ld_acquire_u32(ctr) < target) {
    }
    __threadfence();
  }
  __syncthreads();
}

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// ---------------------------------------------------------------------------------------------------------------
// One CTA per QUARTER of a tile (256 consecutive blocks = 8 chunks of 32 = 32 sub-chunks of 8); the four CTAs of a tile
// meet at a tile-local barrier where the tile's operators / maps are combined.  Grid = 4 x tiles (a multiple of 4, at most one
// CTA per SM); with more quarters than CTAs a CTA takes several, the four of a tile always in the same round.
template <int KP>
__global__ void __launch_bounds__(kFusedThreads, 1) k_sweep_fused(SweepBuffers buf, FusedArgs args) {
  constexpr int L = Layout::L, C = Layout::C, TB = Layout::TB, MB = 8 * Map<KP>::W;
  constexpr int QC = 8;                 // chunks per quarter
  constexpr int NS = 32, SL = 8;        // sub-chunks per quarter, blocks per sub-chunk
  constexpr int PE = 33 * KP, PB = 36;  // row pitches of the emission staging (see k_block_emit)
  constexpr int OPW = KP * KP + KP;     // doubles of one operator in the exchanges (mantissas, then exponents)
  static_assert(Map<KP>::W == 1, "K <= 8: maps are one word");
  static_assert(NS * KP <= kFusedThreads, "thread (sub-chunk, row)");
  extern __shared__ __align__(16) double s_fdyn[];  // two operator buffers of the quarter's 32 sub-chunks (scan ping-pong)
  double* const s_opA = s_fdyn;
  double* const s_opB = s_fdyn + NS * KP * KP;
  double* const s_top = s_fdyn + 2 * NS * KP * KP;  // the tile operators in front of this CTA's tile (tree product)
  __shared__ int s_tex[kFusedMaxTiles * KP];
  __shared__ int s_exA[NS * KP], s_exB[NS * KP];
  double* s_pfx = s_opA;                  // where the prefixes of the sub-chunks inside the tile ended up
  int* s_pfxx = s_exA;
  __shared__ unsigned long long s_subx[NS];  // per sub-chunk: map of the later sub-chunks of the quarter
  __shared__ ModelDev<KP> m;
  __shared__ double s_tab[64];
  __shared__ double s_base[KP * KP];      // product of the quarters in front of this one (identity for quarter 0)
  __shared__ int s_bex[KP];
  __shared__ double s_e[QC * PE];
  __shared__ double s_sx[QC * PB], s_sq[QC * PB];
  __shared__ uint32_t s_n[QC * PB];
  __shared__ double s_red[kFusedThreads / 32][2 * KP];
  __shared__ unsigned long long s_cnt[KP * KP + KP];
  __shared__ float s_g[KP * KP + 2 * KP];
  __shared__ uint32_t s_warp[8];
  __shared__ uint32_t s_misc[4];
  // hand-offs between the phases of a quarter: forward rows (for the maps), maps (for the chunk maps and the states),
  // states (for the statistics).  What another phase of the same CTA wrote to global memory comes back from L2 at
  // ~0.35 us per dependent access; a CTA that owns one quarter per phase (`resident`) keeps all of it in shared memory.
  // (the rows take the place of the emission terms they were computed from: same chunk, same step, same pitch)
  double* const s_alpha = s_e;
  __shared__ unsigned long long s_maps[256];
  __shared__ uint8_t s_states[256 + 8];
  ChainDev* ch = args.chain;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t G = gridDim.x, cta = blockIdx.x;
  exp_table_load(s_tab);
  const uint64_t capacity = buf.capacity;
  unsigned long long* nb_out = const_cast<unsigned long long*>(buf.nblocks);
  uint32_t* starts = const_cast<uint32_t*>(buf.starts);
  double2* spq = const_cast<double2*>(buf.spq);
  unsigned* gctr = args.barriers;  // [0]: grid barrier, [1 + tile]: barrier of the tile's four CTAs
  unsigned ggen = 0;
  // Uses of the tile barrier so far, per round of the quarter loops: round r of this CTA is always the same tile, and
  // its four CTAs take part in exactly the sweeps in which the tile exists — a count per (CTA, round) is the tile's.
  __shared__ unsigned s_tuse[4 * kFusedMaxTiles / 4];
  for (int i = threadIdx.x; i < 4 * kFusedMaxTiles / 4; i += kFusedThreads) s_tuse[i] = 0;

  // phase time stamps of the sweep in progress (globaltimer, CTA 0): where a sweep of the persistent kernel spends its time
#define HML_STAMP(k)                                                    \
  do {                                                                  \
    if (args.phase_ns && cta == 0 && tid == 0) args.phase_ns[k] = global_ns(); \
  } while (0)
  for (int it = 0; it < args.nsweeps; ++it) {
    HML_STAMP(0);
    // ---------------- model of this sweep (written by CTA 0 before the barrier that ended the previous one)
    if (tid < KP) model_from_chain<KP>(ch, m, tid);
    __syncthreads();
    const float thr = __ldcg(&ch->thr);
    const int K = m.K;
    const unsigned long long sweep_key = args.philox_sweep_from_chain ? __ldcg(&ch->sweep) : args.sweep;
    if (cta == 0) {
      // the first phase that writes the result block of this sweep
      for (int i = tid; i < KP + KP * KP + 2; i += kFusedThreads) buf.out_u64[i] = 0;
      for (int i = tid; i < 2 * KP + 1; i += kFusedThreads) buf.out_f64[i] = 0.0;
      if (tid == 0 && !(thr >= ch->cand_floor && thr > 0.f && isfinite(thr))) ch->phase_abort[1] = kChainThreshold;
    }

    // ---------------- boundaries among the candidates: count
    const uint32_t per = ((args.nc + G - 1) / G + 255u) / 256u * 256u;  // slice of this CTA, whole chunks of 256
    const uint32_t lo = min(args.nc, cta * per), hi = min(args.nc, lo + per);
    {
      uint32_t c = 0;
      for (uint32_t i = lo + tid; i < hi; i += kFusedThreads) c += !(args.cand_w[i] < thr) ? 1u : 0u;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
      if (lane == 0) s_warp[warp] = c;
      __syncthreads();
      if (tid == 0) {
        uint32_t t = 0;
        for (int i = 0; i < 8; ++i) t += s_warp[i];
        args.cta_count[cta] = t;
      }
    }
    HML_STAMP(1);
    spin_barrier(gctr, ++ggen * G);  // (1) counts of all slices
    HML_STAMP(2);
    if (const unsigned code = __ldcg(&ch->phase_abort[1])) {
      if (cta == 0 && tid == 0) ch->abort_code = code;
      break;
    }
    // ---------------- ... and scatter: block starts and their integral pairs, in order
    if (warp == 0) {
      uint32_t before = 0, total = 0;
      for (uint32_t c0 = 0; c0 < G; c0 += 32) {
        const uint32_t v = (c0 + lane < G) ? __ldcg(args.cta_count + c0 + lane) : 0u;
        uint32_t b = (c0 + lane < cta) ? v : 0u, t = v;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          b += __shfl_xor_sync(0xffffffffu, b, o);
          t += __shfl_xor_sync(0xffffffffu, t, o);
        }
        before += b;
        total += t;
      }
      if (lane == 0) {
        s_misc[0] = before;
        s_misc[1] = total;
      }
    }
    __syncthreads();
    const uint64_t B = s_misc[1];
    if (cta == 0 && tid == 0) *nb_out = B;
    if (B > capacity || B > (uint64_t)kFusedMaxTiles * TB) {
      if (cta == 0 && tid == 0) {
        ch->nblocks_seen = B;
        ch->phase_abort[2] = kChainCapacity;
      }
    } else {
      uint32_t run = s_misc[0];
      for (uint32_t base = lo; base < hi; base += kFusedThreads) {
        const uint32_t i = base + tid;
        const bool f = i < hi && !(args.cand_w[i] < thr);
        const unsigned bal = __ballot_sync(0xffffffffu, f);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        uint32_t bw = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
          const uint32_t v = s_warp[w];
          if (w < warp) bw += v;
          tot += v;
        }
        if (f) {
          const uint32_t o = run + bw + __popc(bal & ((1u << lane) - 1u));
          starts[o] = args.cand_pos[i];
          spq[o] = args.cand_pq[i];
        }
        run += tot;
        __syncthreads();
      }
      if (cta == 0 && tid == 0) {
        starts[B] = args.T_local;  // sentinel: block b = [starts[b], starts[b + 1])
        spq[B] = buf.pq[args.T_local];
      }
    }
    HML_STAMP(3);
    spin_barrier(gctr, ++ggen * G);  // (2) the block list
    HML_STAMP(4);
    if (const unsigned code = __ldcg(&ch->phase_abort[2])) {
      if (cta == 0 && tid == 0) ch->abort_code = code;
      break;
    }
    const uint32_t ntiles = (uint32_t)((B + TB - 1) / TB);
    const uint32_t nq = 4 * ntiles;  // quarters, the empty ones of the last tile included (their CTAs keep the barriers whole)

    // ================ phase I, per quarter: block statistics, emission terms, sub-chunk operators and their prefixes
    // A quarter is 32 SUB-CHUNKS of 8 consecutive blocks (four per chunk).  The phases that walk blocks one after the
    // other — operator recursion, forward rows, map composition, state look-up — are a single warp per SM-quarter paced
    // by instruction latency (~6.5 cycles per dependent instruction): 8 steps per lane instead of 32 is four times less
    // of that, paid for with a 32-element operator scan inside the quarter.
    const bool resident = nq <= G;  // one quarter per CTA: what a phase leaves in shared memory is there for the next
    for (uint32_t Q = cta; Q < nq; Q += G) {
      const uint32_t tile = Q >> 2;
      const int qi = (int)(Q & 3u), c0 = qi * QC;
      {  // ---- 256 consecutive blocks, one per thread (as k_block_emit)
        const uint64_t b = (uint64_t)Q * 256 + tid;
        const int cl = tid >> 5, t = tid & 31;
        const int oc = tid & 7, ot = tid >> 3;
        const bool valid = b < B;
        uint32_t n = 0;
        double sx = 0.0, sq = 0.0;
        if (valid) block_sums_cg(buf, b, n, sx, sq);
        s_n[cl * PB + t] = n;
        s_sx[cl * PB + t] = sx;
        s_sq[cl * PB + t] = sq;
        double E[KP];
        const double mx = emission_terms<KP, false>(m, (double)n, sx, sq, E);
#pragma unroll
        for (int s = 0; s < KP; ++s) s_e[cl * PE + t * KP + s] = (valid && s < K) ? exp_nonpos(E[s] - mx, s_tab) : 0.0;
        __syncthreads();
        const uint64_t ob = (uint64_t)tile * TB + (uint64_t)(c0 + oc) * L + ot;
        const uint64_t op = Layout::at(tile, c0 + oc, ot);
        if (ob < B) {
          buf.bN[op] = s_n[oc * PB + ot];
          buf.bS[op] = make_double2(s_sx[oc * PB + ot], s_sq[oc * PB + ot]);
        }
        if (!resident) {  // a CTA with several quarters reads the emission terms back in phase II
#pragma unroll
          for (int k = 0; k < KP; ++k) {
            const int q = k * 256 + tid;
            const int qt = q / (8 * KP), within = q % (8 * KP);
            buf.e[(Layout::at(tile, c0, qt)) * KP + within] = s_e[(within / KP) * PE + qt * KP + within % KP];
          }
        }
      }
      // ---- sub-chunk operators: thread (sub-chunk, row) runs the row recursion over its 8 blocks, emission terms
      // straight from the staging area in shared memory
      const int sc = tid / KP, i = tid % KP;
      const uint64_t qfirst = (uint64_t)Q * 256;
      if (sc < NS) {
        const uint64_t first = qfirst + (uint64_t)sc * SL;
        int steps = 0;
        if (first < B) steps = (B - first) < (uint64_t)SL ? (int)(B - first) : SL;
        double r[KP];
        int rex = 0;
#pragma unroll
        for (int j = 0; j < KP; ++j) r[j] = (j == i) ? 1.0 : 0.0;
        const double* ev0 = s_e + (sc >> 2) * PE + (sc & 3) * SL * KP;
#pragma unroll
        for (int h4 = 0; h4 < SL; h4 += 4) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (h4 + q < steps) {
              const double* ev = ev0 + (h4 + q) * KP;
              double y[KP];
#pragma unroll
              for (int j = 0; j < KP; ++j) y[j] = 0.0;
#pragma unroll
              for (int k = 0; k < KP; ++k) {
#pragma unroll
                for (int j = 0; j < KP; ++j) y[j] = fma(r[k], m.A[k][j], y[j]);
              }
#pragma unroll
              for (int j = 0; j < KP; ++j) r[j] = y[j] * ev[j];
            }
          }
          if (rex != kDeadExp && h4 < steps) renorm_pow2<KP>(r, rex);
        }
#pragma unroll
        for (int j = 0; j < KP; ++j) s_opA[(sc * KP + i) * KP + j] = r[j];
        s_exA[sc * KP + i] = rex;
      }
      __syncthreads();
      // ---- inclusive scan over the 32 sub-chunk operators (Hillis-Steele, ping-pong between two buffers): after level d
      // entry sc holds op[max(0, sc - 2d + 1)] x ... x op[sc]
      double* src = s_opA;
      double* dst = s_opB;
      int* srcx = s_exA;
      int* dstx = s_exB;
#pragma unroll 1
      for (int d = 1; d < NS; d <<= 1) {
        if (sc < NS) {
          double r[KP];
          int rex;
          if (sc >= d) {
#pragma unroll
            for (int j = 0; j < KP; ++j) r[j] = src[((sc - d) * KP + i) * KP + j];
            rex = srcx[(sc - d) * KP + i];
            row_times_op<KP, false>(r, rex, src + sc * KP * KP, srcx + sc * KP);
          } else {
#pragma unroll
            for (int j = 0; j < KP; ++j) r[j] = src[(sc * KP + i) * KP + j];
            rex = srcx[sc * KP + i];
          }
#pragma unroll
          for (int j = 0; j < KP; ++j) dst[(sc * KP + i) * KP + j] = r[j];
          dstx[sc * KP + i] = rex;
        }
        __syncthreads();
        double* t1 = src; src = dst; dst = t1;
        int* t2 = srcx; srcx = dstx; dstx = t2;
      }
      // src: inclusive products.  The quarter's total goes to the other CTAs of the tile.
      if (sc == NS - 1) {
        double* out = args.qtot + (size_t)Q * OPW;
#pragma unroll
        for (int j = 0; j < KP; ++j) out[i * KP + j] = src[((NS - 1) * KP + i) * KP + j];
        out[KP * KP + i] = (double)srcx[(NS - 1) * KP + i];
      }
      {
        const uint32_t round = (Q - cta) / G;
        const unsigned target = 4u * (s_tuse[round] + 1u);
        spin_barrier(gctr + 1 + tile, target);  // the four quarters of the tile
        if (tid == 0) s_tuse[round] += 1u;
      }
      // ---- product of the quarters in front of this one (row i by thread i); the last quarter also forms the tile operator
      if (tid < KP) {
        double r[KP];
        int rex = 0;
#pragma unroll
        for (int j = 0; j < KP; ++j) r[j] = (j == tid) ? 1.0 : 0.0;
        for (int k = 0; k <= qi; ++k) {
          if (k == qi && qi != 3) break;
          double M[KP * KP];
          int X[KP];
          const double* qsrc = args.qtot + (size_t)(tile * 4 + k) * OPW;
#pragma unroll
          for (int w = 0; w < KP * KP; ++w) M[w] = __ldcg(qsrc + w);
#pragma unroll
          for (int w = 0; w < KP; ++w) X[w] = (int)__ldcg(qsrc + KP * KP + w);
          if (k == qi) {  // qi == 3: tile operator = base x own total
            double r2[KP];
            int rex2 = rex;
#pragma unroll
            for (int j = 0; j < KP; ++j) r2[j] = r[j];
            row_times_op<KP, false>(r2, rex2, M, X);
#pragma unroll
            for (int j = 0; j < KP; ++j) buf.tile_ops[((uint64_t)tile * KP + tid) * KP + j] = r2[j];
            buf.tile_exp[(uint64_t)tile * KP + tid] = rex2;
          } else {
            row_times_op<KP, false>(r, rex, M, X);
          }
        }
#pragma unroll
        for (int j = 0; j < KP; ++j) s_base[tid * KP + j] = r[j];
        s_bex[tid] = rex;
      }
      __syncthreads();
      // ---- prefix of sub-chunk sc inside the tile = base x (inclusive product up to sc - 1); into the free buffer
      if (sc < NS) {
        double r[KP];
        int rex = s_bex[i];
#pragma unroll
        for (int j = 0; j < KP; ++j) r[j] = s_base[i * KP + j];
        if (sc > 0) row_times_op<KP, false>(r, rex, src + (sc - 1) * KP * KP, srcx + (sc - 1) * KP);
#pragma unroll
        for (int j = 0; j < KP; ++j) dst[(sc * KP + i) * KP + j] = r[j];
        dstx[sc * KP + i] = rex;
        if (!resident) {
          double* g = args.subops + ((size_t)Q * NS + sc) * OPW;
#pragma unroll
          for (int j = 0; j < KP; ++j) g[i * KP + j] = r[j];
          g[KP * KP + i] = (double)rex;
        }
      }
      __syncthreads();
      s_pfx = dst;  // (uniform across the CTA: the scan makes the same number of swaps everywhere)
      s_pfxx = dstx;
    }
    HML_STAMP(5);
    spin_barrier(gctr, ++ggen * G);  // (3) tile operators
    HML_STAMP(6);

    // ================ phase II, per quarter: forward rows, backward maps, sub-chunk and quarter maps, tile map
    unsigned fallbacks = 0;
    for (uint32_t Q = cta; Q < nq; Q += G) {
      const uint32_t tile = Q >> 2;
      const int qi = (int)(Q & 3u), c0 = qi * QC;
      const uint64_t qfirst = (uint64_t)Q * 256;
      // The vector entering the tile = normalise(pi x T_0 x ... x T_{tile-1}).  Every CTA forms it itself: a few tiles by
      // walking them (0.5 us each: one L2 round trip and one vector-operator product after the other), more by a tree
      // over the operators in shared memory, the whole CTA working (log2(tile) levels of operator products, ~0.45 us
      // each) — the walk made the CTAs of the last tiles arrive at the next barrier 9 us late at 18 tiles.
      const bool tree = tile > 4;
      if (tree) {
        for (uint32_t w = tid; w < tile * KP * KP; w += kFusedThreads) s_top[w] = __ldcg(buf.tile_ops + w);
        for (uint32_t w = tid; w < tile * KP; w += kFusedThreads) s_tex[w] = __ldcg(buf.tile_exp + w);
        __syncthreads();
        const int row = tid % KP;
        for (uint32_t d = 1; d < tile; d <<= 1) {
          for (uint32_t pr = tid / KP; pr < (uint32_t)(kFusedThreads / KP); pr += kFusedThreads / KP) {
            for (uint32_t left = 2 * d * pr; left + d < tile; left += 2 * d * (kFusedThreads / KP)) {
              // row `row` of T_left <- (row of T_left) x T_{left + d}: in place, nobody else touches this row
              double r[KP];
              int rex = s_tex[left * KP + row];
#pragma unroll
              for (int j = 0; j < KP; ++j) r[j] = s_top[(left * KP + row) * KP + j];
              row_times_op<KP, false>(r, rex, s_top + (left + d) * KP * KP, s_tex + (left + d) * KP);
#pragma unroll
              for (int j = 0; j < KP; ++j) s_top[(left * KP + row) * KP + j] = r[j];
              s_tex[left * KP + row] = rex;
            }
          }
          __syncthreads();
        }
      }
      if (warp == 0) {
        double a[KP];
#pragma unroll
        for (int j = 0; j < KP; ++j) a[j] = m.pi[j];
        if (tree) {
          OpVals<KP> o;
          load_op<KP>(o, s_top, s_tex);
          if (!vec_apply_op<KP>(a, o)) fallbacks++;
        } else {
          for (uint32_t t = 0; t < tile; ++t) {
            OpVals<KP> o;
            load_op_cg<KP>(o, buf.tile_ops + (uint64_t)t * KP * KP, buf.tile_exp + (uint64_t)t * KP);
            if (!vec_apply_op<KP>(a, o)) fallbacks++;
          }
        }
        // ---- rows: lane = sub-chunk; vector entering it = normalise(vector entering the tile x its prefix), then the
        // recursion over its 8 blocks with a power-of-two rescaling after every fourth (the backward pass normalises its
        // weights itself; four steps cannot move the largest entry by more than min(A)^4 unless the product dies)
        const int sc = lane;
        const uint64_t first = qfirst + (uint64_t)sc * SL;
        int steps = 0;
        if (first < B) steps = (B - first) < (uint64_t)SL ? (int)(B - first) : SL;
        if ((qi > 0 || sc > 0) && steps > 0) {
          OpVals<KP> o;
          if (resident) {
#pragma unroll
            for (int w = 0; w < KP * KP; ++w) o.m[w] = s_pfx[sc * KP * KP + w];
#pragma unroll
            for (int w = 0; w < KP; ++w) o.x[w] = s_pfxx[sc * KP + w];
          } else {
            const double* g = args.subops + ((size_t)Q * NS + sc) * OPW;
#pragma unroll
            for (int w = 0; w < KP * KP; ++w) o.m[w] = g[w];
#pragma unroll
            for (int w = 0; w < KP; ++w) o.x[w] = (int)g[KP * KP + w];
          }
          if (!vec_apply_op<KP>(a, o)) fallbacks++;
        }
        const int cq = sc >> 2, t0 = (sc & 3) * SL;
        double* const row = s_alpha + cq * PE + t0 * KP;  // also where the emission terms of these blocks are (resident)
        const double* ep = buf.e + Layout::at(tile, c0 + cq, t0) * KP;
#pragma unroll
        for (int h4 = 0; h4 < SL; h4 += 4) {
          double ev[4][KP];
          if (resident) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
#pragma unroll
              for (int j = 0; j < KP; ++j) ev[q][j] = row[(h4 + q) * KP + j];
            }
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
#pragma unroll
              for (int j = 0; j < KP; ++j) ev[q][j] = (h4 + q < steps) ? ep[(uint64_t)(h4 + q) * C * KP + j] : 0.0;
            }
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (h4 + q < steps) {
              double f[KP];
#pragma unroll
              for (int j = 0; j < KP; ++j) f[j] = 0.0;
#pragma unroll
              for (int k = 0; k < KP; ++k) {
#pragma unroll
                for (int j = 0; j < KP; ++j) f[j] = fma(a[k], m.A[k][j], f[j]);
              }
#pragma unroll
              for (int j = 0; j < KP; ++j) {
                a[j] = f[j] * ev[q][j];
                row[(h4 + q) * KP + j] = a[j];
              }
            }
          }
          if (h4 < steps) {
            double mxv = 0.0;
#pragma unroll
            for (int j = 0; j < KP; ++j) mxv = fmax(mxv, a[j]);
            if (mxv > 1e-250) {
              const double scl = pow2i(-exponent_of(mxv));
#pragma unroll
              for (int j = 0; j < KP; ++j) a[j] *= scl;
            } else {  // a vanished (or nearly vanished) forward sum, FB.hpp:106-111: the host re-runs the sweep
              fallbacks++;
            }
          }
        }
      }
      __syncthreads();
      HML_STAMP(13);
      {  // ---- backward map of this thread's block (as k_bwd_maps)
        const uint64_t b = qfirst + tid;
        const uint64_t p = Layout::perm(b);
        Map<KP> fm = Map<KP>::identity();
        if (b < B) {
          const bool last = b + 1 == B;
          const double Nm1 = (double)(resident ? s_n[(tid >> 5) * PB + (tid & 31)] : buf.bN[p]) - 1.0;
          double ap[KP];
#pragma unroll
          for (int j = 0; j < KP; ++j) {
            const double al = s_alpha[(tid >> 5) * PE + (tid & 31) * KP + j];
            ap[j] = (last || !m.use_self || j >= K) ? al : al * exp_nonpos(Nm1 * m.loga[j], s_tab);
          }
          const double u = buf.replay_u ? buf.replay_u[B - 1 - b] : Philox::uniform(args.seed, sweep_key, 0u, b);
          fm = Map<KP>::zero();
          if (last) {
            const uint32_t q = discrete_draw<KP>(ap, K, u);
#pragma unroll
            for (int j = 0; j < KP; ++j) fm.set(j, q);
          } else {
#pragma unroll
            for (int j = 0; j < KP; ++j) {
              if (j < K) {
                double cs[KP];
                cs[0] = ap[0] * m.A[0][j];
#pragma unroll
                for (int k = 1; k < KP; ++k) cs[k] = fma(ap[k], m.A[k][j], cs[k - 1]);
                bool tie;
                uint32_t q = discrete_draw_fast<KP>(cs, K, u, tie);
                if (tie) {
                  Weights<KP> w;
#pragma unroll
                  for (int k = 0; k < KP; ++k) w.v[k] = __dmul_rn(ap[k], m.A[k][j]);
                  q = discrete_draw_exact<KP>(w, K, u);
                }
                fm.set(j, q);
              }
            }
          }
        }
        s_maps[tid] = fm.w[0];
        if (!resident) fm.store(buf.maps + p * MB);  // phase III of a CTA with several quarters reads them back
      }
      __syncthreads();
      HML_STAMP(14);
      if (warp == 0) {
        // maps of the quarter's 32 sub-chunks, their suffixes inside the quarter, the quarter map
        Map<KP> Gm = Map<KP>::identity();
#pragma unroll
        for (int t = 0; t < SL; ++t) {
          Map<KP> f;
          f.w[0] = s_maps[lane * SL + t];
          Gm = Gm.after(f);
        }
        Map<KP> inc = Gm;  // inclusive suffix over lanes lane..31
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          Map<KP> other;
          other.w[0] = __shfl_down_sync(0xffffffffu, inc.w[0], o);
          if (lane + o < 32) inc = inc.after(other);
        }
        Map<KP> excl;
        excl.w[0] = __shfl_down_sync(0xffffffffu, inc.w[0], 1);
        if (lane == 31) excl = Map<KP>::identity();
        s_subx[lane] = excl.w[0];  // the later sub-chunks of the quarter
        if (!resident) args.submaps[(size_t)Q * NS + lane] = excl.w[0];
        if (lane == 0) args.qmap[Q] = inc.w[0];
      }
      HML_STAMP(15);
      {
        const uint32_t round = (Q - cta) / G;
        const unsigned target = 4u * (s_tuse[round] + 1u);
        spin_barrier(gctr + 1 + tile, target);
        if (tid == 0) s_tuse[round] += 1u;
      }
      if (qi == 0 && tid == 0) {  // tile map = composition of the four quarter maps
        unsigned long long w4[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) w4[k] = __ldcg(args.qmap + Q + k);
        Map<KP> tm;
        tm.w[0] = w4[0];
#pragma unroll
        for (int k = 1; k < 4; ++k) {
          Map<KP> o;
          o.w[0] = w4[k];
          tm = tm.after(o);
        }
        tm.store(buf.tile_maps + (uint64_t)tile * MB);
      }
    }
    if (fallbacks) ch->phase_abort[4] = kChainFallback;
    HML_STAMP(7);
    spin_barrier(gctr, ++ggen * G);  // (4) tile maps
    HML_STAMP(8);
    if (const unsigned code = __ldcg(&ch->phase_abort[4])) {
      if (cta == 0 && tid == 0) ch->abort_code = code;
      break;
    }

    // ================ phase III, per quarter: the state following it, the states of its blocks, their statistics
    double ax[KP], aq[KP];
    unsigned long long an[KP], ad[KP];
#pragma unroll
    for (int s = 0; s < KP; ++s) {
      ax[s] = aq[s] = 0.0;
      an[s] = ad[s] = 0;
    }
    for (int i = tid; i < KP * KP + KP; i += kFusedThreads) s_cnt[i] = 0;
    __syncthreads();
    for (uint32_t Q = cta; Q < nq; Q += G) {
      const uint32_t tile = Q >> 2;
      const int qi = (int)(Q & 3u), c0 = qi * QC;
      const uint64_t qfirst = (uint64_t)Q * 256;
      if (warp == 0) {
        // maps of the later tiles (lane t holds tiles 32 k + t) and of the later quarters of this tile: all loads are in
        // flight together, then the look-ups run from the last tile down
        unsigned long long tw[kFusedMaxTiles / 32];
#pragma unroll
        for (int k = 0; k < kFusedMaxTiles / 32; ++k) {
          const uint32_t t = 32u * k + lane;
          tw[k] = (t > tile && t < ntiles) ? __ldcg(reinterpret_cast<const unsigned long long*>(buf.tile_maps + (uint64_t)t * MB)) : 0ull;
        }
        const unsigned long long qw = (lane > (unsigned)qi && lane < 4) ? __ldcg(args.qmap + tile * 4 + lane) : 0ull;
        const unsigned long long sx0 = resident ? s_subx[lane] : args.submaps[(size_t)Q * NS + lane];
        uint32_t q = 0;  // the last block of the sequence carries a constant map: the start value is irrelevant
        for (uint32_t t = ntiles - 1; t > tile; --t) {
          unsigned long long w = 0;
#pragma unroll
          for (int k = 0; k < kFusedMaxTiles / 32; ++k)
            if ((t >> 5) == (uint32_t)k) w = __shfl_sync(0xffffffffu, tw[k], t & 31);
          q = (uint32_t)(w >> (8 * q)) & 0xffu;
        }
        if (qi == 0 && lane == 0) buf.tile_qin[tile] = (uint8_t)q;
        for (int k = 3; k > qi; --k) {
          const unsigned long long w = __shfl_sync(0xffffffffu, qw, k);
          q = (uint32_t)(w >> (8 * q)) & 0xffu;
        }
        if (lane == 0) s_misc[2] = q;  // state of the block that follows the quarter
        // ---- states of the quarter's sub-chunks (as k_bwd_replay): lane = sub-chunk, 8 look-ups
        const uint64_t first = qfirst + (uint64_t)lane * SL;
        if (first < B) {
          const int steps = (B - first) < (uint64_t)SL ? (int)(B - first) : SL;
          uint32_t qs = (uint32_t)(sx0 >> (8 * q)) & 0xffu;  // state of the block that follows the sub-chunk
          const int cq = lane >> 2, t0 = (lane & 3) * SL;
          unsigned long long f8[SL];
#pragma unroll
          for (int t = 0; t < SL; ++t)
            f8[t] = resident ? s_maps[lane * SL + t]
                             : *reinterpret_cast<const unsigned long long*>(buf.maps + Layout::at(tile, c0 + cq, t0 + t) * MB);
#pragma unroll
          for (int t = SL - 1; t >= 0; --t) {
            if (t < steps) {
              qs = (uint32_t)(f8[t] >> (8 * qs)) & 0xffu;
              s_states[lane * SL + t] = (uint8_t)qs;
              buf.states[Layout::at(tile, c0 + cq, t0 + t)] = (uint8_t)qs;
            }
          }
        }
      }
      __syncthreads();
      const uint32_t qafter = s_misc[2];
      {  // statistics of this thread's block (as k_reduce_partial; the successor of the quarter's last block is qafter)
        const uint64_t b = qfirst + tid;
        if (b < B) {
          const uint64_t p = Layout::perm(b);
          const uint32_t st = s_states[tid];
          const bool has_next = b + 1 < B;
          uint32_t next = st;
          if (has_next) next = (tid == 255) ? qafter : (uint32_t)s_states[tid + 1];
          uint32_t n;
          double2 v;
          if (resident) {
            const int si = (tid >> 5) * PB + (tid & 31);
            n = s_n[si];
            v = make_double2(s_sx[si], s_sq[si]);
          } else {
            n = buf.bN[p];
            v = buf.bS[p];
          }
          unsigned long long dg = (unsigned long long)(n - 1) + ((has_next && next == st) ? 1ull : 0ull);
          if (b == 0) {  // the phantom transition 0 -> q_0 (FB.hpp:177,182-184)
            if (st == 0u)
              dg += 1ull;
            else
              atomicAdd(&s_cnt[st], 1ull);
          }
#pragma unroll
          for (int s = 0; s < KP; ++s) {
            const bool hit = st == (uint32_t)s;
            ax[s] += hit ? v.x : 0.0;
            aq[s] += hit ? v.y : 0.0;
            an[s] += hit ? (unsigned long long)n : 0ull;
            ad[s] += hit ? dg : 0ull;
          }
          if (has_next && next != st) atomicAdd(&s_cnt[st * KP + next], 1ull);
        }
      }
      __syncthreads();
    }
#pragma unroll
    for (int s = 0; s < KP; ++s) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        ax[s] += shfl_xor_double(ax[s], o);
        aq[s] += shfl_xor_double(aq[s], o);
        an[s] += __shfl_xor_sync(0xffffffffu, an[s], o);
        ad[s] += __shfl_xor_sync(0xffffffffu, ad[s], o);
      }
    }
    if (lane == 0) {
#pragma unroll
      for (int s = 0; s < KP; ++s) {
        s_red[warp][s] = ax[s];
        s_red[warp][KP + s] = aq[s];
        if (an[s]) atomicAdd(&s_cnt[KP * KP + s], an[s]);
        if (ad[s]) atomicAdd(&s_cnt[s * KP + s], ad[s]);
      }
    }
    __syncthreads();
    if (tid < 2 * KP) {
      double t = 0.0;
      for (int w = 0; w < kFusedThreads / 32; ++w) t += s_red[w][tid];
      buf.partials[(size_t)cta * 2 * KP + tid] = t;
    }
    for (int i = tid; i < KP * KP; i += kFusedThreads)
      if (s_cnt[i]) atomicAdd(&buf.out_u64[KP + i], s_cnt[i]);
    for (int i = tid; i < KP; i += kFusedThreads)
      if (s_cnt[KP * KP + i]) atomicAdd(&buf.out_u64[i], s_cnt[KP * KP + i]);
    HML_STAMP(9);
    spin_barrier(gctr, ++ggen * G);  // (5) partial statistics
    HML_STAMP(10);
    if (cta == 0) {
      // final sums in a fixed order (deterministic), then the parameters of the next sweep
      for (int v = warp; v < 2 * KP; v += kFusedThreads / 32) {  // lanes stride over the CTAs, then a fixed tree
        double t = 0.0;
        for (uint32_t c2 = lane; c2 < G; c2 += 32) t += __ldcg(buf.partials + (size_t)c2 * 2 * KP + v);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += shfl_xor_double(t, o);
        if (lane == 0) buf.out_f64[v] = t;
      }
      __syncthreads();
      if (tid == 0) ch->sweeps_done += 1u;
      if (args.sample_params) {
        __threadfence();
        chain_sample_params<KP>(ch, buf.out_u64, buf.out_f64, s_g);
      }
      __threadfence();
    }
    HML_STAMP(11);
    spin_barrier(gctr, ++ggen * G);  // (6) the model of the next sweep
    HML_STAMP(12);
    if (const unsigned code = __ldcg(&ch->phase_abort[6])) {
      if (cta == 0 && tid == 0) ch->abort_code = code;
      break;
    }
  }
#undef HML_STAMP
}

template <int KP>
constexpr size_t fused_dyn_smem() { return (size_t)(2 * 32 + kFusedMaxTiles) * KP * KP * sizeof(double); }

// one cooperative launch of `a.nsweeps` sweeps; returns the cudaError_t of the launch
template <int KP>
int fused_impl(const SweepBuffers& b, const FusedArgs& a, int grid, cudaStream_t s) {
  SweepBuffers bb = b;
  FusedArgs aa = a;
  void* params[] = {(void*)&bb, (void*)&aa};
  cudaFuncSetAttribute(k_sweep_fused<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fused_dyn_smem<KP>());
  return (int)cudaLaunchCooperativeKernel((const void*)k_sweep_fused<KP>, dim3(grid), dim3(kFusedThreads), params,
                                          fused_dyn_smem<KP>(), s);
}
// CTAs of k_sweep_fused<KP> that can be resident at once on the current device (0: the kernel does not fit)
template <int KP>
int fused_max_grid_impl(int sms) {
  int per_sm = 0;
  cudaFuncSetAttribute(k_sweep_fused<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fused_dyn_smem<KP>());
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sweep_fused<KP>, kFusedThreads, fused_dyn_smem<KP>()) != cudaSuccess)
    return 0;
  return per_sm > 0 ? sms : 0;  // one CTA per SM is all the kernel asks for
}
template <int KP>
int chain_params_impl(ChainDev* ch, const unsigned long long* out_u64, const double* out_f64, cudaStream_t s) {
  k_chain_params<KP><<<1, kFusedThreads, 0, s>>>(ch, out_u64, out_f64);
  return (int)cudaGetLastError();
}

}  // namespace hml
