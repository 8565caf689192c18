// hammlet_b200 — templated block-level kernels of one Gibbs sweep (included by hml_sweep.cu, which carries
// the overview, and by hml_sweep_inst.cu, which instantiates them per padded state count).
#pragma once
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/hammlet_b200.h"
#include "hml_common.cuh"
#include "hml_kernels.h"
#include "hml_p2p.cuh"
#include "hml_seg_head.cuh"

namespace hml {

// ------------------------------------------------------------------------------------------------
// layout

struct Layout {
  static constexpr int L = kChunkLen;  // blocks per chunk
  static constexpr int C = 32;         // chunks per tile
  static constexpr int TB = L * C;     // blocks per tile
  static_assert(L * C == kTileBlocks, "layout");
  // natural block index -> storage index: [tile][step t][chunk c]
  __host__ __device__ static inline uint64_t perm(uint64_t b) {
    const uint64_t tile = b / TB, r = b % TB;
    return tile * TB + (r % L) * C + (r / L);
  }
  __host__ __device__ static inline uint64_t inv(uint64_t p) {
    const uint64_t tile = p / TB, r = p % TB;
    return tile * TB + (r % C) * L + (r / C);
  }
  __host__ __device__ static inline uint64_t at(uint64_t tile, int c, int t) { return tile * TB + (uint64_t)t * C + c; }
};

template <int KP>
struct ModelDev {
  double A[KP][KP];
  double mean[KP], inv2var[KP], lognorm[KP], loga[KP], pi[KP];
  int K, use_self;
};

template <int KP>
static ModelDev<KP> make_model(const ModelHost& m) {
  ModelDev<KP> d;
  memset(&d, 0, sizeof(d));
  d.K = m.K;
  d.use_self = m.use_self;
  for (int i = 0; i < m.K; ++i) {
    d.mean[i] = m.mean[i];
    d.inv2var[i] = 1.0 / (2.0 * m.var[i]);
    // EFD.hpp:35-38 with the cached stdev of Observation.hpp:175-185
    d.lognorm[i] = log(sqrt(m.var[i])) + m.mean[i] * m.mean[i] / (2 * m.var[i]);
    d.loga[i] = m.use_self ? log(m.A[i * m.K + i]) : 0.0;  // FB.hpp:47-50
    d.pi[i] = m.pi[i];
    for (int j = 0; j < m.K; ++j) d.A[i][j] = m.A[i * m.K + j];
  }
  return d;
}

// Multivariate data (D > 1): state s uses parameter mapping[s][d] in dimension d (Mapping.hpp:89-117), so the
// emission term is a sum over the dimensions (EFD.hpp:83-93) and the log-normaliser a sum over the state's
// parameters (Theta.hpp:154-164).
template <int KP>
struct EmitMD {
  double mean[kMaxDims][KP], inv2var[kMaxDims][KP];
  double lognorm[KP], loga[KP];
  int K, D;
};

template <int KP>
static EmitMD<KP> make_emit_md(const ModelHost& m) {
  EmitMD<KP> d;
  memset(&d, 0, sizeof(d));
  d.K = m.K;
  d.D = m.D;
  for (int s = 0; s < m.K; ++s) {
    double ln = 0.0;
    for (int dim = 0; dim < m.D; ++dim) {
      const double mu = m.mean_sd[dim][s], var = m.var_sd[dim][s];
      d.mean[dim][s] = mu;
      d.inv2var[dim][s] = 1.0 / (2.0 * var);
      ln += log(sqrt(var)) + mu * mu / (2 * var);
    }
    d.lognorm[s] = ln;
    d.loga[s] = m.use_self ? log(m.A[s * m.K + s]) : 0.0;
  }
  return d;
}

// ------------------------------------------------------------------------------------------------
// small-vector helpers (all loops fully unrolled: register-resident vectors)

constexpr int kDeadExp = -(1 << 30);

template <int KP>
__device__ __forceinline__ void renorm_pow2(double (&v)[KP], int& ex) {
  double m = v[0];
#pragma unroll
  for (int j = 1; j < KP; ++j) m = fmax(m, v[j]);
  if (m > 0.0) {
    int e = exponent_of(m);
    if (e < -1000) e = -1000;
    const double s = pow2i(-e);
#pragma unroll
    for (int j = 0; j < KP; ++j) v[j] *= s;
    ex += e;
  } else {
    ex = kDeadExp;
  }
}

// r <- r * Op, Op given as KP x KP row-major mantissas M with per-row binary exponents X.
template <int KP, bool kViaL2>
__device__ __forceinline__ void row_times_op(double (&r)[KP], int& rex, const double* M, const int* X) {
  int xk[KP];
  int xm = kDeadExp;
#pragma unroll
  for (int k = 0; k < KP; ++k) {
    xk[k] = kViaL2 ? __ldcg(X + k) : X[k];
    if (r[k] > 0.0 && xk[k] > xm) xm = xk[k];
  }
  if (rex == kDeadExp || xm == kDeadExp) {
#pragma unroll
    for (int j = 0; j < KP; ++j) r[j] = 0.0;
    rex = kDeadExp;
    return;
  }
  double y[KP];
#pragma unroll
  for (int j = 0; j < KP; ++j) y[j] = 0.0;
#pragma unroll
  for (int k = 0; k < KP; ++k) {
    const double a = r[k] * pow2i(xk[k] - xm);
#pragma unroll
    for (int j = 0; j < KP; ++j) {
      const double mkj = kViaL2 ? __ldcg(M + k * KP + j) : M[k * KP + j];
      y[j] = fma(a, mkj, y[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < KP; ++j) r[j] = y[j];
  rex += xm;
  renorm_pow2<KP>(r, rex);
}

// Warp-cooperative a <- normalise(a * Op): lane j holds a_j (lanes >= KP hold 0) and column j of the
// operator (LaneOp), loaded one step ahead of its use so the dependent chain never waits on memory.
template <int KP>
struct LaneOp {
  double col[KP];  // col[k] = M[k][lane]
  int x;           // lane k holds the binary exponent of row k
};

template <int KP>
__device__ __forceinline__ void load_lane_op(LaneOp<KP>& o, const double* M, const int* X, int lane) {
#pragma unroll
  for (int k = 0; k < KP; ++k) o.col[k] = lane < KP ? M[k * KP + lane] : 0.0;
  o.x = lane < KP ? X[lane] : kDeadExp;
}

// Returns false if the product vanished (the caller flags a uniform-fallback suspect).
template <int KP>
__device__ __forceinline__ bool warp_apply_op(double& a, const LaneOp<KP>& o) {
  int xm = kDeadExp;
  int xk[KP];
  double ak[KP];
#pragma unroll
  for (int k = 0; k < KP; ++k) {
    ak[k] = shfl_double(a, k);
    xk[k] = __shfl_sync(0xffffffffu, o.x, k);
    if (ak[k] > 0.0 && xk[k] > xm) xm = xk[k];
  }
  double y = 0.0;
  if (xm != kDeadExp) {
#pragma unroll
    for (int k = 0; k < KP; ++k) y = fma(ak[k] * pow2i(xk[k] - xm), o.col[k], y);
  }
  double s = y;
#pragma unroll
  for (int o2 = 16; o2 > 0; o2 >>= 1) s += shfl_xor_double(s, o2);
  if (!(s > 0.0)) {
    a = 0.0;
    return false;
  }
  a = y / s;
  return true;
}

// Thread-local a <- normalise(a * Op) for small K (operator held in registers).  Returns false if the
// product vanished.
template <int KP, typename OpT>
__device__ __forceinline__ bool vec_apply_op(double (&a)[KP], const OpT& o) {
  int xm = kDeadExp;
#pragma unroll
  for (int k = 0; k < KP; ++k)
    if (a[k] > 0.0 && o.x[k] > xm) xm = o.x[k];
  double y[KP];
#pragma unroll
  for (int j = 0; j < KP; ++j) y[j] = 0.0;
  if (xm != kDeadExp) {
#pragma unroll
    for (int k = 0; k < KP; ++k) {
      const double ak = a[k] * pow2i(o.x[k] - xm);
#pragma unroll
      for (int j = 0; j < KP; ++j) y[j] = fma(ak, o.m[k * KP + j], y[j]);
    }
  }
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < KP; ++j) s += y[j];
  if (!(s > 0.0)) {
#pragma unroll
    for (int j = 0; j < KP; ++j) a[j] = 0.0;
    return false;
  }
  const double inv = 1.0 / s;
#pragma unroll
  for (int j = 0; j < KP; ++j) a[j] = y[j] * inv;
  return true;
}

// Thread-local operator copy (small K only) for prefetching ahead of a dependent row recursion.
template <int KP>
struct OpVals {
  double m[KP * KP];
  int x[KP];
};
template <int KP>
__device__ __forceinline__ void load_op(OpVals<KP>& o, const double* M, const int* X) {
#pragma unroll
  for (int k = 0; k < KP * KP; ++k) o.m[k] = M[k];
#pragma unroll
  for (int k = 0; k < KP; ++k) o.x[k] = X[k];
}

// ---- byte-packed state maps (KPB = 8, 16 or 32 entries of one byte)
template <int KP>
struct Map {
  static constexpr int W = (KP <= 8 ? 1 : (KP <= 16 ? 2 : 4));
  uint64_t w[W];
  __device__ __forceinline__ static Map identity() {
    Map m;
#pragma unroll
    for (int i = 0; i < W; ++i) m.w[i] = 0x0706050403020100ull + 0x0808080808080808ull * i;
    return m;
  }
  __device__ __forceinline__ uint32_t get(uint32_t j) const {
    if (W == 1) return (uint32_t)(w[0] >> (8 * j)) & 0xffu;
    uint64_t x = w[0];
#pragma unroll
    for (int i = 1; i < W; ++i)
      if ((j >> 3) == (uint32_t)i) x = w[i];
    return (uint32_t)(x >> (8 * (j & 7))) & 0xffu;
  }
  // static index only
  __device__ __forceinline__ void set(int j, uint32_t v) { w[j >> 3] |= (uint64_t)v << (8 * (j & 7)); }
  __device__ __forceinline__ static Map zero() {
    Map m;
#pragma unroll
    for (int i = 0; i < W; ++i) m.w[i] = 0;
    return m;
  }
  // this o f : j -> this[f[j]]
  __device__ __forceinline__ Map after(const Map& f) const {
    Map r = zero();
#pragma unroll
    for (int j = 0; j < KP; ++j) r.set(j, get(f.get(j)));
    return r;
  }
  __device__ __forceinline__ void store(uint8_t* p) const {
    uint64_t* q = reinterpret_cast<uint64_t*>(p);
#pragma unroll
    for (int i = 0; i < W; ++i) q[i] = w[i];
  }
  __device__ __forceinline__ static Map load(const uint8_t* p) {
    Map m;
    const uint64_t* q = reinterpret_cast<const uint64_t*>(p);
#pragma unroll
    for (int i = 0; i < W; ++i) m.w[i] = q[i];
    return m;
  }
  // through L2: the map was written by another CTA of the running kernel
  __device__ __forceinline__ static Map load_cg(const uint8_t* p) {
    Map m;
    const unsigned long long* q = reinterpret_cast<const unsigned long long*>(p);
#pragma unroll
    for (int i = 0; i < W; ++i) m.w[i] = __ldcg(q + i);
    return m;
  }
};

// std::discrete_distribution rule (libstdc++ bits/random.tcc, used by Trellis.hpp:61-66 and
// Mixture.hpp:111-112): normalise (every weight divided by the sum), partial sums, last := 1, first k with
// cp[k] >= u.  A row without positive mass gives NaN partial sums in the reference, for which lower_bound
// returns index 0.
//
// discrete_draw_exact follows it operation for operation (round-to-nearest intrinsics, so nothing is contracted).
// discrete_draw_fast needs no division at all: with c[k] the running sums of the weights, cp[k] >= u is decided as
// c[k] >= u * c[K-1].  The two can disagree only if some cp[k] lies within K * 2^-51 of u; the fast version reports
// |c[k] - u s| < 2^-40 s as a tie and the caller then takes the exact path — about once in 1e11 draws — so the result
// is the reference's for every u.  The comparison is made on the sign and the exponent field of d = c[k] - u s in the
// integer pipe (a rounded difference has the sign of the exact one).
template <int KP>
struct Weights {
  double v[KP];
};
// out of line: it runs about once in 1e11 draws and its divisions must not cost the callers registers
template <int KP>
__device__ __noinline__ uint32_t discrete_draw_exact(const Weights<KP> wp, int K, double u) {
  const double (&p)[KP] = wp.v;
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < KP; ++k) s = (k < K) ? __dadd_rn(s, p[k]) : s;
  if (!(s > 0.0)) return 0u;
  double acc = 0.0;
  uint32_t res = (uint32_t)(K - 1);
  bool found = false;
#pragma unroll
  for (int k = 0; k < KP - 1; ++k) {
    if (k < K - 1) {
      acc = __dadd_rn(acc, __ddiv_rn(p[k], s));
      if (!found && acc >= u) {
        res = (uint32_t)k;
        found = true;
      }
    }
  }
  return res;
}

// c[k] = p[0] + ... + p[k] (c[KP-1] = the sum: weights of padded states are zero)
template <int KP>
__device__ __forceinline__ uint32_t discrete_draw_fast(const double (&c)[KP], int K, double u, bool& tie) {
  const double s = c[KP - 1];
  tie = false;
  if (!(s > 0.0)) return 0u;
  const double us = u * s;
  const int lim = (__double2hiint(s) & 0x7ff00000) - (40 << 20);
  uint32_t res = (uint32_t)(K - 1);
  bool found = false;
#pragma unroll
  for (int k = 0; k < KP - 1; ++k) {
    if (k < K - 1) {
      const int hi = __double2hiint(c[k] - us);
      tie |= (hi & 0x7ff00000) < lim;
      if (!found && hi >= 0) {
        res = (uint32_t)k;
        found = true;
      }
    }
  }
  return res;
}

// weights given directly (the last block of the backward pass, the mixture sampler)
template <int KP>
__device__ __forceinline__ uint32_t discrete_draw(const double (&p)[KP], int K, double u) {
  double c[KP];
  c[0] = p[0];
#pragma unroll
  for (int k = 1; k < KP; ++k) c[k] = c[k - 1] + ((k < K) ? p[k] : 0.0);
  bool tie;
  uint32_t r = discrete_draw_fast<KP>(c, K, u, tie);
  if (tie) {
    Weights<KP> w;
#pragma unroll
    for (int k = 0; k < KP; ++k) w.v[k] = p[k];
    r = discrete_draw_exact<KP>(w, K, u);
  }
  return r;
}

// Number of blocks the kernels may touch.  If boundary detection found more blocks than the per-block
// arrays hold, the block list is incomplete: every kernel then sees an empty structure and the host,
// which reads the same counter after the sweep, grows the arrays and repeats the sweep.
__device__ __forceinline__ uint64_t device_nblocks(const unsigned long long* nb, uint64_t capacity) {
  const uint64_t b = *nb;
  return b <= capacity ? b : 0;
}

// ---- segment mode: what a rank derives from the all-gathered heads {blocks, head length, head sums}
__device__ __forceinline__ uint64_t seg_first_block(const SegInfo& g) {  // global index of this rank's block 0
  uint64_t o = 0;
  for (int r = 0; r < g.rank; ++r) o += (uint64_t)g.heads[kHeadWords * r];
  return o;
}
__device__ __forceinline__ uint64_t seg_global_blocks(const SegInfo& g, uint64_t local) {
  if (g.world <= 1) return local;
  uint64_t o = 0;
  for (int r = 0; r < g.world; ++r) o += (uint64_t)g.heads[kHeadWords * r];
  return o;
}
__device__ __forceinline__ bool seg_later_blocks(const SegInfo& g) {  // does any later rank own a block?
  for (int r = g.rank + 1; r < g.world; ++r)
    if (g.heads[kHeadWords * r] > 0.0) return true;
  return false;
}

// (sum x, sum x^2) over the local observations [s, e) from the integral arrays
// (Statistics/IntegralArray.hpp:104-124): cell-local running sums + double-double cell offsets
__device__ __forceinline__ void range_sums_from(const SweepBuffers& buf, const double2 ps, const double2 pe, uint32_t s,
                                                uint32_t e, double& sx, double& sq) {
  sx = pe.x - ps.x;
  sq = pe.y - ps.y;
  const uint32_t cs = s >> kCellLog2, ce = e >> kCellLog2;
  if (cs != ce) {
    const double4 a = buf.cell_pref[cs], z = buf.cell_pref[ce];
    sx += (z.x - a.x) + (z.y - a.y);
    sq += (z.z - a.z) + (z.w - a.w);
  }
}
__device__ __forceinline__ void range_sums(const SweepBuffers& buf, uint32_t s, uint32_t e, double& sx, double& sq) {
  const double2 ps = buf.pq[s], pe = buf.pq[e];
  sx = pe.x - ps.x;
  sq = pe.y - ps.y;
  const uint32_t cs = s >> kCellLog2, ce = e >> kCellLog2;
  if (cs != ce) {
    const double4 a = buf.cell_pref[cs], z = buf.cell_pref[ce];
    sx += (z.x - a.x) + (z.y - a.y);
    sq += (z.z - a.z) + (z.w - a.w);
  }
}

// (range_sums_dim and the head of a rank's segment: hml_seg_head.cuh)

// k_seg_head: the head partial as a kernel of its own (the candidate scatter does it in its last CTA when it can)
static __global__ void __launch_bounds__(256) k_seg_head(SweepBuffers buf, uint32_t seg_len, unsigned long long seq) {
  pdl_enter();
  seg_head_cta(buf, seg_len, seq);
}

// ------------------------------------------------------------------------------------------------
// k_block_emit: block statistics + emission terms

// block b's (N, sum x, sum x^2) from the integral arrays; in segment mode the rank's last block continues on the
// following ranks up to their first boundary
__device__ __forceinline__ void block_sums(const SweepBuffers& buf, uint64_t b, uint64_t B, uint32_t& n, double& sx, double& sq) {
  const uint32_t s = buf.starts[b], e = buf.starts[b + 1];
  if (buf.spq)  // the pairs of this block's start and end sit next to each other (candidate list)
    range_sums_from(buf, buf.spq[b], buf.spq[b + 1], s, e, sx, sq);
  else
    range_sums(buf, s, e, sx, sq);
  n = e - s;
  if (buf.seg.world > 1 && b + 1 == B) {
    for (int r = buf.seg.rank + 1; r < buf.seg.world; ++r) {
      const double* hd = buf.seg.heads + kHeadWords * r;
      n += (uint32_t)hd[1];
      sx += hd[2];
      sq += hd[3];
      if (hd[0] > 0.0) break;
    }
  }
}

// the K emission log-weights of a block, EFD.hpp:23-32 then FB.hpp:74-81 (Mixture.hpp:98 has no self-transition term)
template <int KP, bool kMix>
__device__ __forceinline__ double emission_terms(const ModelDev<KP>& m, double N, double sx, double sq, double (&E)[KP]) {
  double mx = -INFINITY;
#pragma unroll
  for (int s = 0; s < KP; ++s) {
    double v = (2.0 * m.mean[s] * sx - sq) * m.inv2var[s] - N * m.lognorm[s];
    if (!kMix) v += (N - 1.0) * m.loga[s];
    E[s] = v;
    if (s < m.K) mx = fmax(mx, v);
  }
  return mx;
}

// The self-transition rescale A_ss^(N-1) of FB.hpp:115-119 is not stored: k_bwd_maps, its only reader, derives it
// from the block size (5 table-based exps there against 8 K bytes per block written here and read there).
template <int KP, bool kGather, bool kEmit, bool kMix>
__global__ void __launch_bounds__(256) k_block_emit(SweepBuffers buf, ModelDev<KP> m, int want_maxe) {
  pdl_enter();
  // K <= 8, dynamic blocks (the default path): the CTA takes 256 CONSECUTIVE blocks — eight chunks of a tile — so that
  // the block starts and the integral pairs are read as contiguous pieces (a thread per storage slot reads 32 different
  // sectors per warp instruction: 11 of 32 bytes used, the L1 pipe 70 % busy, the kernel bound by it), and hands its
  // results to the chunk-interleaved layout through shared memory: the eight chunks' values of a step sit next to each
  // other in storage, so every step is one contiguous piece of 8 * KP doubles.  Row pitches (33 * KP doubles, 36
  // scalars) keep both the natural-order writes and the storage-order reads at the minimum of two wavefronts per
  // 64-bit access.
  constexpr bool kNatural = kEmit && kGather && KP <= 8;
  constexpr bool kStage = kEmit && KP <= 8 && !kNatural;
  constexpr int PE = 33 * KP, PB = 36;
  __shared__ double s_e[kNatural ? 8 * PE : (kStage ? 256 * KP : 1)];
  __shared__ double s_sx[kNatural ? 8 * PB : 1], s_sq[kNatural ? 8 * PB : 1];
  __shared__ uint32_t s_n[kNatural ? 8 * PB : 1];
  __shared__ double s_tab[kEmit ? 64 : 1];  // 2^(j/64) for exp_nonpos
  if (kEmit) {
    exp_table_load(s_tab);
    __syncthreads();
  }
  if (kEmit && blockIdx.x == 0) {
    // first kernel of a sweep: zero the result block that the later kernels accumulate into
    for (int i = threadIdx.x; i < KP + KP * KP + 2; i += blockDim.x) buf.out_u64[i] = 0;
    for (int i = threadIdx.x; i < 2 * KP + 1; i += blockDim.x) buf.out_f64[i] = 0.0;
  }
  const uint64_t B = device_nblocks(buf.nblocks, buf.capacity);
  const uint64_t slots = (B + Layout::TB - 1) / Layout::TB * Layout::TB;
  if constexpr (kNatural) {
    const int cl = threadIdx.x >> 5, t = threadIdx.x & 31;   // chunk within the group, step
    const int oc = threadIdx.x & 7, ot = threadIdx.x >> 3;   // the slot this thread writes out: chunk, step
    // where the words this thread copies out sit in the staging area and, relative to the group's first slot, in storage
    // (the grid is a few CTAs per SM and every CTA takes several groups: worked out once)
    int src_w[KP], dst_w[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) {
      const int q = k * 256 + threadIdx.x;  // word of the group's emission terms in storage order
      const int qt = q / (8 * KP), within = q % (8 * KP);
      src_w[k] = (within / KP) * PE + qt * KP + within % KP;
      dst_w[k] = qt * Layout::C * KP + within;
    }
    for (uint64_t g = blockIdx.x; g < slots / 256; g += gridDim.x) {
      const uint64_t b = g * 256 + threadIdx.x;
      const uint64_t tile = g >> 2;
      const int c0 = (int)(g & 3) * 8;
      const bool valid = b < B;
      uint32_t n = 0;
      double sx = 0.0, sq = 0.0;
      if (valid) block_sums(buf, b, B, n, sx, sq);
      s_n[cl * PB + t] = n;
      s_sx[cl * PB + t] = sx;
      s_sq[cl * PB + t] = sq;
      const double N = (double)n;
      double E[KP];
      const double mx = emission_terms<KP, kMix>(m, N, sx, sq, E);
#pragma unroll
      for (int s = 0; s < KP; ++s) s_e[cl * PE + t * KP + s] = (valid && s < m.K) ? exp_nonpos(E[s] - mx, s_tab) : 0.0;
      if (want_maxe && valid) buf.maxE[Layout::at(tile, c0 + cl, t)] = mx;
      __syncthreads();
      const uint64_t ob = tile * Layout::TB + (uint64_t)(c0 + oc) * Layout::L + ot;  // natural index of the slot written
      const uint64_t op = Layout::at(tile, c0 + oc, ot);
      if (ob < B) {
        buf.bN[op] = s_n[oc * PB + ot];
        buf.bS[op] = make_double2(s_sx[oc * PB + ot], s_sq[oc * PB + ot]);
      }
      double* const eg = buf.e + (tile * Layout::TB + c0) * KP;
#pragma unroll
      for (int k = 0; k < KP; ++k) eg[dst_w[k]] = s_e[src_w[k]];
      __syncthreads();
    }
  }
  // slots is a multiple of 1024 and the stride a multiple of 256: all threads of a CTA make the same trips
  for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; !kNatural && p < slots; p += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t b = Layout::inv(p);
    const bool valid = b < B;
    if (!kStage && !valid) continue;
    uint32_t n = 0;
    double sx = 0.0, sq = 0.0;
    if (valid) {
      if (kGather) {
        block_sums(buf, b, B, n, sx, sq);
        buf.bN[p] = n;
        buf.bS[p] = make_double2(sx, sq);
      } else {
        n = buf.bN[p];
        const double2 v = buf.bS[p];
        sx = v.x;
        sq = v.y;
      }
    }
    if (kEmit) {
      const double N = (double)n;
      double E[KP];
      const double mx = emission_terms<KP, kMix>(m, N, sx, sq, E);
      if (kStage) {
        // the KP terms of a block sit KP * 8 bytes apart from the next block's: staged in shared memory and written as
        // one contiguous piece (lane-strided stores cost five times the sectors at K = 5)
#pragma unroll
        for (int s = 0; s < KP; ++s) s_e[threadIdx.x * KP + s] = (valid && s < m.K) ? exp_nonpos(E[s] - mx, s_tab) : 0.0;
        __syncthreads();
        const uint64_t base = (p - threadIdx.x) * KP;  // first word of the CTA's 256 slots
#pragma unroll
        for (int k = 0; k < KP; ++k) buf.e[base + k * 256 + threadIdx.x] = s_e[k * 256 + threadIdx.x];
        __syncthreads();
      } else {
#pragma unroll
        for (int s = 0; s < KP; ++s) buf.e[p * KP + s] = (s < m.K) ? exp_nonpos(E[s] - mx, s_tab) : 0.0;
      }
      if (want_maxe && valid) buf.maxE[p] = mx;
    }
  }
}

// k_block_emit_md: the same for multivariate data — block sums of every dimension (plane d of bS at d * capacity)
// and emission terms summed over the dimensions.  Single handle only (no segment heads).
template <int KP, bool kGather, bool kEmit, bool kMix>
__global__ void __launch_bounds__(256) k_block_emit_md(SweepBuffers buf, EmitMD<KP> m, int want_maxe) {
  pdl_enter();
  __shared__ double s_tab[kEmit ? 64 : 1];
  if (kEmit) {
    exp_table_load(s_tab);
    __syncthreads();
  }
  if (kEmit && blockIdx.x == 0) {
    for (int i = threadIdx.x; i < KP + KP * KP + 2; i += blockDim.x) buf.out_u64[i] = 0;
    for (int i = threadIdx.x; i < 2 * KP + 1; i += blockDim.x) buf.out_f64[i] = 0.0;
  }
  const uint64_t B = device_nblocks(buf.nblocks, buf.capacity);
  const uint64_t slots = (B + Layout::TB - 1) / Layout::TB * Layout::TB;
  const int D = buf.D;
  for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < slots; p += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t b = Layout::inv(p);
    if (b >= B) continue;
    uint32_t n;
    double sx[kMaxDims], sq[kMaxDims];
    if (kGather) {
      const uint32_t s = buf.starts[b], e = buf.starts[b + 1];
      n = e - s;
#pragma unroll
      for (int d = 0; d < kMaxDims; ++d)
        if (d < D) range_sums_dim(buf, d, s, e, sx[d], sq[d]);
      if (buf.seg.world > 1 && b + 1 == B) {
        // the rank's last block continues on the following ranks up to their first boundary
        for (int r = buf.seg.rank + 1; r < buf.seg.world; ++r) {
          const double* hd = buf.seg.heads + kHeadWords * r;
          n += (uint32_t)hd[1];
#pragma unroll
          for (int d = 0; d < kMaxDims; ++d) {
            if (d < D) {
              sx[d] += hd[2 + 2 * d];
              sq[d] += hd[3 + 2 * d];
            }
          }
          if (hd[0] > 0.0) break;
        }
      }
      buf.bN[p] = n;
#pragma unroll
      for (int d = 0; d < kMaxDims; ++d)
        if (d < D) buf.bS[(size_t)d * buf.capacity + p] = make_double2(sx[d], sq[d]);
    } else {
      n = buf.bN[p];
#pragma unroll
      for (int d = 0; d < kMaxDims; ++d) {
        if (d < D) {
          const double2 v = buf.bS[(size_t)d * buf.capacity + p];
          sx[d] = v.x;
          sq[d] = v.y;
        }
      }
    }
    if (kEmit) {
      const double N = (double)n;
      double E[KP];
      double mx = -INFINITY;
#pragma unroll
      for (int s = 0; s < KP; ++s) {
        double ip = 0.0;  // EFD.hpp:83-93: sum over the dimensions of the per-dimension inner products
#pragma unroll
        for (int d = 0; d < kMaxDims; ++d)
          if (d < D) ip += (2.0 * m.mean[d][s] * sx[d] - sq[d]) * m.inv2var[d][s];
        double v = ip - N * m.lognorm[s];
        if (!kMix) v += (N - 1.0) * m.loga[s];
        E[s] = v;
        if (s < m.K) mx = fmax(mx, v);
      }
#pragma unroll
      for (int s = 0; s < KP; ++s) buf.e[p * KP + s] = (s < m.K) ? exp_nonpos(E[s] - mx, s_tab) : 0.0;
      if (want_maxe) buf.maxE[p] = mx;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// k_fwd_chunks: CTA per tile; thread (chunk c, row i) runs the row recursion over the chunk.

template <int KP>
struct FwdCfg {
  static constexpr int CG = (KP <= 8) ? 32 : (KP <= 16 ? 16 : 8);  // chunks handled per pass
  static constexpr int THREADS = ((CG * KP + 31) / 32) * 32;
  static constexpr bool SMEM_OPS = KP <= 8;
};

template <int KP>
__global__ void __launch_bounds__(FwdCfg<KP>::THREADS) k_fwd_chunks(SweepBuffers buf, ModelDev<KP> m) {
  constexpr int L = Layout::L, C = Layout::C, CG = FwdCfg<KP>::CG;
  __shared__ double s_ops[FwdCfg<KP>::SMEM_OPS ? C * KP * KP : 1];
  __shared__ int s_exp[FwdCfg<KP>::SMEM_OPS ? C * KP : 1];
  const uint64_t B = device_nblocks(buf.nblocks, buf.capacity);
  const uint64_t ntiles = (B + Layout::TB - 1) / Layout::TB;
  const int cl = threadIdx.x / KP, i = threadIdx.x % KP;
  for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    for (int cg = 0; cg < C; cg += CG) {
      const int c = cg + cl;
      if (cl < CG) {
        const uint64_t first = tile * Layout::TB + (uint64_t)c * L;
        int steps = 0;
        if (first < B) steps = (B - first) < (uint64_t)L ? (int)(B - first) : L;
        double r[KP];
        int rex = 0;
#pragma unroll
        for (int j = 0; j < KP; ++j) r[j] = (j == i) ? 1.0 : 0.0;
        const double* ep = buf.e + Layout::at(tile, c, 0) * KP;
        double en[KP];  // emission terms of the next step, loaded one step ahead of their use
#pragma unroll
        for (int j = 0; j < KP; ++j) en[j] = steps > 0 ? ep[j] : 0.0;
#pragma unroll 1
        for (int t = 0; t < steps; ++t) {
          double ev[KP];
#pragma unroll
          for (int j = 0; j < KP; ++j) ev[j] = en[j];
          if (t + 1 < steps) {
#pragma unroll
            for (int j = 0; j < KP; ++j) en[j] = ep[(uint64_t)(t + 1) * C * KP + j];
          }
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
          if (rex != kDeadExp) renorm_pow2<KP>(r, rex);
        }
        const uint64_t ch = tile * C + c;
#pragma unroll
        for (int j = 0; j < KP; ++j) buf.chunk_ops[(ch * KP + i) * KP + j] = r[j];
        buf.chunk_exp[ch * KP + i] = rex;
        if (FwdCfg<KP>::SMEM_OPS) {
#pragma unroll
          for (int j = 0; j < KP; ++j) s_ops[(c * KP + i) * KP + j] = r[j];
          s_exp[c * KP + i] = rex;
        }
      }
    }
    __threadfence_block();
    __syncthreads();
    // tile operator = ordered product of the 32 chunk operators
    if (FwdCfg<KP>::SMEM_OPS) {
      // pairwise tree in shared memory: thread (pair, row) replaces row `row` of the left operator by
      // its product with the right operator; 5 levels
      for (int stride = 1; stride < C; stride <<= 1) {
        const int pairs = C / (2 * stride);
        const int pr = threadIdx.x / KP, row = threadIdx.x % KP;
        if (pr < pairs) {
          const int a = pr * 2 * stride, b = a + stride;
          double r[KP];
#pragma unroll
          for (int j = 0; j < KP; ++j) r[j] = s_ops[(a * KP + row) * KP + j];
          int rex = s_exp[a * KP + row];
          row_times_op<KP, false>(r, rex, s_ops + b * KP * KP, s_exp + b * KP);
#pragma unroll
          for (int j = 0; j < KP; ++j) s_ops[(a * KP + row) * KP + j] = r[j];
          s_exp[a * KP + row] = rex;
        }
        __syncthreads();
      }
      if (threadIdx.x < KP) {
#pragma unroll
        for (int j = 0; j < KP; ++j) buf.tile_ops[(tile * KP + threadIdx.x) * KP + j] = s_ops[threadIdx.x * KP + j];
        buf.tile_exp[tile * KP + threadIdx.x] = s_exp[threadIdx.x];
      }
    } else if (threadIdx.x < KP) {
      double r[KP];
      int rex = 0;
#pragma unroll
      for (int j = 0; j < KP; ++j) r[j] = (j == (int)threadIdx.x) ? 1.0 : 0.0;
#pragma unroll 1
      for (int c = 0; c < C; ++c)
        row_times_op<KP, true>(r, rex, buf.chunk_ops + (tile * C + c) * KP * KP, buf.chunk_exp + (tile * C + c) * KP);
#pragma unroll
      for (int j = 0; j < KP; ++j) buf.tile_ops[(tile * KP + threadIdx.x) * KP + j] = r[j];
      buf.tile_exp[tile * KP + threadIdx.x] = rex;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// k_fwd_chunks_wide (K > 8): the arithmetic of k_fwd_chunks (same row recursion, same rescaling, so the chunk
// operators are bit-identical), laid out for the fp64 pipe instead of for registers:
//   * thread (chunk, row) for ALL chunks of the tile at once where the register file allows (K <= 20: 32 K threads,
//     12-20 warps per SM instead of 5), the row vector and its product in registers, A as constant-bank operands, the
//     emission terms read at their use (the K threads of a chunk read the same values: one L1 line, broadcast);
//   * the tile operator is a pairwise tree over the 32 chunk operators (5 levels of row-times-operator products
//     through L2, every thread busy) instead of one warp walking the 32 operators one after the other.
// Intermediate tree nodes live in a per-CTA scratch area in global memory (30 operators; L2-resident); the right
// operands of a level are staged in shared memory by the whole CTA in one coalesced pass (reading them element by
// element through L2 inside the products cost 11 us per level: 40 dependent round trips at 100 registers).
template <int KP>
struct WideCfg {
  static constexpr int CG = (KP <= 20) ? 32 : 8;  // chunks per pass
  static constexpr int THREADS = ((CG * KP + 31) / 32) * 32;
  static constexpr int kTreeNodes = 31;            // 16 + 8 + 4 + 2 + 1
  static constexpr size_t kScratchDoubles = (size_t)kTreeNodes * KP * KP;
  static constexpr size_t kScratchInts = (size_t)kTreeNodes * KP;
  // dynamic shared memory: the (at most 16) right-hand operators of one tree level and their row exponents
  static constexpr size_t kSmem = (size_t)16 * KP * KP * sizeof(double) + (size_t)16 * KP * sizeof(int);
};

// r <- r * Op with the operator staged in shared memory, without keeping the KP exponents in registers
template <int KP>
__device__ __forceinline__ void row_times_op_lean(double (&r)[KP], int& rex, const double* M, const int* X) {
  int xm = kDeadExp;
#pragma unroll
  for (int k = 0; k < KP; ++k) {
    const int x = X[k];
    if (r[k] > 0.0 && x > xm) xm = x;
  }
  if (rex == kDeadExp || xm == kDeadExp) {
#pragma unroll
    for (int j = 0; j < KP; ++j) r[j] = 0.0;
    rex = kDeadExp;
    return;
  }
  double y[KP];
#pragma unroll
  for (int j = 0; j < KP; ++j) y[j] = 0.0;
#pragma unroll
  for (int k = 0; k < KP; ++k) {
    const double a = r[k] * pow2i(X[k] - xm);
#pragma unroll
    for (int j = 0; j < KP; ++j) y[j] = fma(a, M[k * KP + j], y[j]);
  }
#pragma unroll
  for (int j = 0; j < KP; ++j) r[j] = y[j];
  rex += xm;
  renorm_pow2<KP>(r, rex);
}

template <int KP>
__global__ void __launch_bounds__(WideCfg<KP>::THREADS) k_fwd_chunks_wide(SweepBuffers buf, ModelDev<KP> m,
                                                                         double* __restrict__ scratch_ops,
                                                                         int* __restrict__ scratch_exp) {
  pdl_enter();
  constexpr int L = Layout::L, C = Layout::C, CG = WideCfg<KP>::CG;
  static_assert(KP % 2 == 0, "emission terms are read as double2");
  const uint64_t B = device_nblocks(buf.nblocks, buf.capacity);
  const uint64_t ntiles = (B + Layout::TB - 1) / Layout::TB;
  const int cl = threadIdx.x / KP, i = threadIdx.x % KP;
  double* const sops = scratch_ops + (size_t)blockIdx.x * WideCfg<KP>::kScratchDoubles;
  int* const sexp = scratch_exp + (size_t)blockIdx.x * WideCfg<KP>::kScratchInts;
  extern __shared__ __align__(16) unsigned char s_dyn_wide[];
  double* const s_rop = reinterpret_cast<double*>(s_dyn_wide);       // [pair][KP*KP]
  int* const s_rex = reinterpret_cast<int*>(s_rop + 16 * KP * KP);   // [pair][KP]
  for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    // ---- chunk operators: K independent row recursions per chunk (FB.hpp:64-125 as operator products)
#pragma unroll 1
    for (int cg = 0; cg < C; cg += CG) {
      const int c = cg + cl;
      if (cl < CG) {
        const uint64_t first = tile * Layout::TB + (uint64_t)c * L;
        int steps = 0;
        if (first < B) steps = (B - first) < (uint64_t)L ? (int)(B - first) : L;
        double r[KP];
        int rex = 0;
#pragma unroll
        for (int j = 0; j < KP; ++j) r[j] = (j == i) ? 1.0 : 0.0;
        const double2* ep = reinterpret_cast<const double2*>(buf.e + Layout::at(tile, c, 0) * KP);
#pragma unroll 1
        for (int t = 0; t < steps; ++t) {
          double y[KP];
#pragma unroll
          for (int j = 0; j < KP; ++j) y[j] = 0.0;
#pragma unroll
          for (int k = 0; k < KP; ++k) {
#pragma unroll
            for (int j = 0; j < KP; ++j) y[j] = fma(r[k], m.A[k][j], y[j]);
          }
          const double2* et = ep + (size_t)t * C * (KP / 2);
#pragma unroll
          for (int j = 0; j < KP; j += 2) {
            const double2 ev = et[j / 2];
            r[j] = y[j] * ev.x;
            r[j + 1] = y[j + 1] * ev.y;
          }
          if (rex != kDeadExp) renorm_pow2<KP>(r, rex);
        }
        const uint64_t ch = tile * C + c;
#pragma unroll
        for (int j = 0; j < KP; ++j) buf.chunk_ops[(ch * KP + i) * KP + j] = r[j];
        buf.chunk_exp[ch * KP + i] = rex;
      }
    }
    __threadfence_block();
    __syncthreads();
    // ---- tile operator: pairwise tree.  Level 0 reads the chunk operators, level l > 0 the nodes of level l - 1;
    // node p of a level is (left operand 2p) x (right operand 2p + 1); the root goes to tile_ops.
    int in_base = 0;   // first node of the previous level inside the scratch area
    int out_base = 0;  // first node of this level
#pragma unroll 1
    for (int pairs = C / 2, level = 0; pairs >= 1; pairs >>= 1, ++level) {
      // operands of this level: nodes 2p (left) and 2p + 1 (right) of the previous one
      const double* const in_ops = level == 0 ? buf.chunk_ops + tile * C * KP * KP : sops + (size_t)in_base * KP * KP;
      const int* const in_exp = level == 0 ? buf.chunk_exp + tile * C * KP : sexp + (size_t)in_base * KP;
      for (int idx = threadIdx.x; idx < pairs * KP * KP; idx += blockDim.x) {
        const int pr = idx / (KP * KP), off = idx % (KP * KP);
        s_rop[idx] = __ldcg(in_ops + (size_t)(2 * pr + 1) * KP * KP + off);
      }
      for (int idx = threadIdx.x; idx < pairs * KP; idx += blockDim.x) {
        const int pr = idx / KP, off = idx % KP;
        s_rex[idx] = __ldcg(in_exp + (2 * pr + 1) * KP + off);
      }
      __syncthreads();
#pragma unroll 1
      for (int task = threadIdx.x; task < pairs * KP; task += blockDim.x) {
        const int pr = task / KP, row = task % KP;
        const double* lop = in_ops + (size_t)(2 * pr) * KP * KP;
        double r[KP];
#pragma unroll
        for (int j = 0; j < KP; ++j) r[j] = __ldcg(lop + row * KP + j);
        int rx = __ldcg(in_exp + (2 * pr) * KP + row);
        row_times_op_lean<KP>(r, rx, s_rop + pr * KP * KP, s_rex + pr * KP);
        double* dst = pairs == 1 ? buf.tile_ops + (tile * KP + row) * KP : sops + ((size_t)(out_base + pr) * KP + row) * KP;
        int* dex = pairs == 1 ? buf.tile_exp + tile * KP + row : sexp + (size_t)(out_base + pr) * KP + row;
#pragma unroll
        for (int j = 0; j < KP; ++j) dst[j] = r[j];
        *dex = rx;
      }
      __threadfence_block();
      __syncthreads();
      in_base = out_base;
      out_base += pairs;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// k_fwd_chunks_prefix (K <= 8): as k_fwd_chunks, but what it leaves in chunk_ops is, for every chunk, the
// ordered product of the EARLIER chunk operators of its tile (exclusive prefix; identity for chunk 0), so that
// the replay kernel gets the vector entering a chunk from one vector-operator product instead of a serial walk.
//   rows    thread (chunk, row): row recursion over the 32 blocks of the chunk; emission terms loaded four
//           steps at a time, exact power-of-two rescaling every fourth step
//   scan    work-efficient exclusive scan over the 32 chunk operators in shared memory (up-sweep: tile
//           operator; down-sweep: prefixes); products are formed in registers between two barriers because a
//           node's operand is overwritten in place
template <int KP>
__global__ void __launch_bounds__(FwdCfg<KP>::THREADS, (KP <= 5 ? 4 : (KP <= 6 ? 3 : 2))) k_fwd_chunks_prefix(SweepBuffers buf, ModelDev<KP> m) {
  pdl_enter();
  constexpr int L = Layout::L, C = Layout::C;
  static_assert(FwdCfg<KP>::CG == C && L % 4 == 0, "small-K configuration");
  __shared__ double s_ops[C * KP * KP];
  __shared__ int s_exp[C * KP];
  const uint64_t B = device_nblocks(buf.nblocks, buf.capacity);
  const uint64_t ntiles = (B + Layout::TB - 1) / Layout::TB;
  const int c = threadIdx.x / KP, i = threadIdx.x % KP;
  for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    if (c < C) {
      const uint64_t first = tile * Layout::TB + (uint64_t)c * L;
      int steps = 0;
      if (first < B) steps = (B - first) < (uint64_t)L ? (int)(B - first) : L;
      double r[KP];
      int rex = 0;
#pragma unroll
      for (int j = 0; j < KP; ++j) r[j] = (j == i) ? 1.0 : 0.0;
      const double* ep = buf.e + Layout::at(tile, c, 0) * KP;
#pragma unroll 1
      for (int t0 = 0; t0 < steps; t0 += 4) {
        double ev[4][KP];  // emission terms of four steps: all loads in flight before the first use
#pragma unroll
        for (int q = 0; q < 4; ++q) {
#pragma unroll
          for (int j = 0; j < KP; ++j) ev[q][j] = (t0 + q < steps) ? ep[(uint64_t)(t0 + q) * C * KP + j] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (t0 + q < steps) {
            double y[KP];
#pragma unroll
            for (int j = 0; j < KP; ++j) y[j] = 0.0;
#pragma unroll
            for (int k = 0; k < KP; ++k) {
#pragma unroll
              for (int j = 0; j < KP; ++j) y[j] = fma(r[k], m.A[k][j], y[j]);
            }
#pragma unroll
            for (int j = 0; j < KP; ++j) r[j] = y[j] * ev[q][j];
          }
        }
        // four steps shrink the largest entry by at most min(A)^4 unless the product dies; a row that underflows
        // here is reported like a dead one and the host falls back to the exact sequential recursion
        if (rex != kDeadExp) renorm_pow2<KP>(r, rex);
      }
#pragma unroll
      for (int j = 0; j < KP; ++j) s_ops[(c * KP + i) * KP + j] = r[j];
      s_exp[c * KP + i] = rex;
    }
    __syncthreads();
    // ---- up-sweep: node at index n = (k+1)*2*stride - 1 becomes (its left neighbour's product) x (itself)
    const int pr = threadIdx.x / KP, row = threadIdx.x % KP;
    for (int stride = 1; stride < C; stride <<= 1) {
      const int n = (pr + 1) * 2 * stride - 1;
      const bool on = n < C;
      double r[KP];
      int rex = 0;
      if (on) {
#pragma unroll
        for (int j = 0; j < KP; ++j) r[j] = s_ops[((n - stride) * KP + row) * KP + j];
        rex = s_exp[(n - stride) * KP + row];
        row_times_op<KP, false>(r, rex, s_ops + n * KP * KP, s_exp + n * KP);
      }
      __syncthreads();
      if (on) {
#pragma unroll
        for (int j = 0; j < KP; ++j) s_ops[(n * KP + row) * KP + j] = r[j];
        s_exp[n * KP + row] = rex;
      }
      __syncthreads();
    }
    if (threadIdx.x < KP) {
#pragma unroll
      for (int j = 0; j < KP; ++j) {
        buf.tile_ops[(tile * KP + threadIdx.x) * KP + j] = s_ops[((C - 1) * KP + threadIdx.x) * KP + j];
        s_ops[((C - 1) * KP + threadIdx.x) * KP + j] = (j == (int)threadIdx.x) ? 1.0 : 0.0;  // root prefix: identity
      }
      buf.tile_exp[tile * KP + threadIdx.x] = s_exp[(C - 1) * KP + threadIdx.x];
      s_exp[(C - 1) * KP + threadIdx.x] = 0;
    }
    __syncthreads();
    // ---- down-sweep: left child takes the node's prefix, right child takes prefix x (left child's product)
    for (int stride = C / 2; stride >= 1; stride >>= 1) {
      const int n = (pr + 1) * 2 * stride - 1;
      const bool on = n < C;
      double pfx[KP], r[KP];
      int pex = 0, rex = 0;
      if (on) {
#pragma unroll
        for (int j = 0; j < KP; ++j) pfx[j] = r[j] = s_ops[(n * KP + row) * KP + j];
        pex = rex = s_exp[n * KP + row];
        row_times_op<KP, false>(r, rex, s_ops + (n - stride) * KP * KP, s_exp + (n - stride) * KP);
      }
      __syncthreads();
      if (on) {
#pragma unroll
        for (int j = 0; j < KP; ++j) {
          s_ops[((n - stride) * KP + row) * KP + j] = pfx[j];
          s_ops[(n * KP + row) * KP + j] = r[j];
        }
        s_exp[(n - stride) * KP + row] = pex;
        s_exp[n * KP + row] = rex;
      }
      __syncthreads();
    }
    // ---- prefixes to global memory (coalesced)
    for (int k = threadIdx.x; k < C * KP * KP; k += blockDim.x) buf.chunk_ops[tile * C * KP * KP + k] = s_ops[k];
    for (int k = threadIdx.x; k < C * KP; k += blockDim.x) buf.chunk_exp[tile * C * KP + k] = s_exp[k];
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// Speculative forward pass (rank convergence).  The filter forgets where it started: two runs of the recursion
// alpha_t = normalise(e_t o (alpha_{t-1} A)) from different vectors become parallel after a few informative blocks.
// So instead of products of K x K chunk operators (2 K^3 flop per block) every chunk runs the VECTOR recursion
// (2 K^2) from a guess — uniform, pushed through the `warm` blocks in front of the chunk — and a second
// pass repairs the first rows of every chunk from the (by then known) last row of its predecessor until the repaired
// row is parallel to the stored one; the rest of the chunk is then right as it stands.  A chunk whose rows have not
// met by its last block reports a failure and the host runs the sweep again through the operator scan, which needs
// no such assumption.  (ForwardBackward.hpp:64-125 is one sequential loop; SURVEY.md §7 "rank-1 collapse".)
constexpr double kSpecTol = 1e-13;  // relative agreement of every component of two rows that count as parallel

// index of the result slot that counts chunks whose repair did not converge (the pad word after the fallbacks)
template <int KP>
__device__ __forceinline__ int spec_fail_slot() {
  return KP + KP * KP + 1;
}

// the guess for the vector entering the piece that starts at block `first`: uniform pushed through the `warm` blocks
// in front of it (rescaled by exact powers of two; a vanished vector restarts uniform).  A piece with fewer than `warm`
// blocks in front of it starts from pi at block 0, i.e. exactly; the first piece of a later rank of a split sequence
// has no blocks in front of it on this device and starts uniform (k_fwd_fixup_head repairs it).
template <int KP>
__device__ __forceinline__ void spec_entry(const SweepBuffers& buf, const ModelDev<KP>& m, uint64_t first, int warm,
                                           double (&a)[KP]) {
  const int K = m.K;
  uint64_t b0 = first;  // first warm-up block (`first`: block index of the piece's first block)
  if (first <= (uint64_t)warm) {
    b0 = 0;
    if (buf.seg.world > 1 && buf.seg.rank > 0) {
#pragma unroll
      for (int j = 0; j < KP; ++j) a[j] = (j < K) ? 1.0 : 0.0;
    } else {
#pragma unroll
      for (int j = 0; j < KP; ++j) a[j] = m.pi[j];
    }
  } else {
    b0 = first - (uint64_t)warm;
#pragma unroll
    for (int j = 0; j < KP; ++j) a[j] = (j < K) ? 1.0 : 0.0;
  }
  if (b0 == first) return;
  double en[KP];  // emission terms one step ahead of their use
  {
    const double* ep = buf.e + Layout::perm(b0) * KP;
#pragma unroll
    for (int j = 0; j < KP; ++j) en[j] = ep[j];
  }
#pragma unroll 1
  for (uint64_t b = b0; b < first; ++b) {
    double ev[KP];
#pragma unroll
    for (int j = 0; j < KP; ++j) ev[j] = en[j];
    if (b + 1 < first) {
      const double* ep = buf.e + Layout::perm(b + 1) * KP;
#pragma unroll
      for (int j = 0; j < KP; ++j) en[j] = ep[j];
    }
    double f[KP];
#pragma unroll
    for (int j = 0; j < KP; ++j) f[j] = 0.0;
#pragma unroll
    for (int k = 0; k < KP; ++k) {
#pragma unroll
      for (int j = 0; j < KP; ++j) f[j] = fma(a[k], m.A[k][j], f[j]);
    }
    double mxv = 0.0;
#pragma unroll
    for (int j = 0; j < KP; ++j) {
      f[j] *= ev[j];
      mxv = fmax(mxv, f[j]);
    }
    if (mxv > 0.0) {
      int e2 = exponent_of(mxv);
      if (e2 < -1000) e2 = -1000;
      const double sc = pow2i(-e2);
#pragma unroll
      for (int j = 0; j < KP; ++j) a[j] = f[j] * sc;
    } else {
#pragma unroll
      for (int j = 0; j < KP; ++j) a[j] = (j < K) ? 1.0 : 0.0;
    }
  }
}

// are x and y parallel?  |x_j / sum(x) - y_j / sum(y)| <= tol * x_j / sum(x) for every j, without dividing
template <int KP>
__device__ __forceinline__ bool rows_parallel(const double (&x)[KP], const double (&y)[KP]) {
  double sx = 0.0, sy = 0.0;
#pragma unroll
  for (int j = 0; j < KP; ++j) {
    sx += x[j];
    sy += y[j];
  }
  bool ok = sx > 0.0 && sy > 0.0;
#pragma unroll
  for (int j = 0; j < KP; ++j) {
    const double l = x[j] * sy, r = y[j] * sx;
    ok = ok && (fabs(l - r) <= kSpecTol * l + 1e-300);
  }
  return ok;
}

// ------------------------------------------------------------------------------------------------
// k_fwd_replay_prefix (K <= 8): one warp per tile, lane c owns chunk c.  The vector entering the chunk is
// normalise(vector entering the tile x prefix operator of the chunk); then the reference's own recursion
// alpha_t = e_t o (alpha_{t-1} A) runs over the chunk, with e arriving in slabs of kSlab steps by
// double-buffered bulk copies.
//   kExact   every step divides by the forward sum like the reference (FB.hpp:101-105); required for the
//            log-likelihood and for kept rows
//   !kExact  every step rescales by an exact power of two instead: the rows written are c_t * alpha_t with
//            c_t = 2^k, which is all the backward pass needs (it normalises its weights itself)
template <int KP>
struct ReplayCfg {
  static constexpr int kSlab = 4;                                                    // steps per stage
  static constexpr size_t kStage = (size_t)kSlab * Layout::C * KP * sizeof(double);  // e of a slab
  static constexpr size_t kSmem = 2 * kStage;
};

//   kSpec    speculative pass: the vector entering the chunk is a guess (spec_entry), k_fwd_fixup repairs the rows;
//            the log-likelihood terms go to `lognorm` per block and are summed after the repair
template <int KP, bool kExact, bool kLoglik, bool kSpec = false>
//            and the warp takes one PIECE of `sub` steps (8, 16 or 32) of every chunk of a tile at a time — the chunk
//            length is no longer tied to anything, and shorter pieces mean more warps for a latency-bound recursion
__global__ void __launch_bounds__(32) k_fwd_replay_prefix(SweepBuffers buf, ModelDev<KP> m, double* __restrict__ lognorm, int warm,
                                                          int sub) {
  pdl_enter();
  using Cfg = ReplayCfg<KP>;
  constexpr int L = Layout::L, C = Layout::C, S = Cfg::kSlab, NS = L / S;
  static_assert(!kLoglik || kExact, "the log-likelihood needs the forward sums");
  extern __shared__ __align__(16) unsigned char s_dyn[];
  __shared__ __align__(8) uint64_t s_bar[2];
  unsigned char* stage0 = s_dyn;
  unsigned char* stage1 = s_dyn + Cfg::kStage;
  const uint64_t B = device_nblocks(buf.nblocks, buf.capacity);
  const uint64_t ntiles = (B + Layout::TB - 1) / Layout::TB;
  const int lane = threadIdx.x;
  const int K = m.K;
  if (lane == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    mbar_init_fence();
  }
  __syncwarp();
  uint32_t ph0 = 0, ph1 = 0;
  double ll = 0.0;
  unsigned fallbacks = 0;
  auto issue_slab = [&](uint64_t tile, int slab, unsigned char* dst, uint64_t* bar) {
    const uint64_t off = (tile * Layout::TB + (uint64_t)slab * S * C) * KP;
    mbar_expect_tx(bar, (uint32_t)Cfg::kStage);
    bulk_g2s(dst, buf.e + off, (uint32_t)Cfg::kStage, bar);
  };
  const int nsub = kSpec ? L / sub : 1;               // pieces per chunk
  const int slab0_of = kSpec ? sub / S : NS;          // slabs per piece
  for (uint64_t unit = blockIdx.x; unit < ntiles * nsub; unit += gridDim.x) {
    const uint64_t tile = unit / nsub;
    const int piece = (int)(unit % nsub);
    const int s_lo = piece * slab0_of, s_hi = s_lo + slab0_of;  // this unit's slabs
    const int t_lo = s_lo * S;
    if (lane == 0) {
      fence_proxy_async();
      issue_slab(tile, s_lo, stage0, &s_bar[0]);
      issue_slab(tile, s_lo + 1, stage1, &s_bar[1]);
    }
    const int c = lane;
    const uint64_t first = tile * Layout::TB + (uint64_t)c * L;
    int steps = 0;  // steps of the chunk that exist (the piece covers [t_lo, t_lo + slab0_of * S) of them)
    if (first < B) steps = (B - first) < (uint64_t)L ? (int)(B - first) : L;
    // ---- vector entering chunk c
    double a[KP];
    if constexpr (kSpec) {
#pragma unroll
      for (int j = 0; j < KP; ++j) a[j] = 0.0;
      if (steps > t_lo) spec_entry<KP>(buf, m, first + (uint64_t)t_lo, warm, a);
    } else {
#pragma unroll
      for (int j = 0; j < KP; ++j) a[j] = buf.tile_ain[tile * KP + j];
      if (c > 0 && steps > 0) {
        OpVals<KP> o;
        load_op<KP>(o, buf.chunk_ops + (tile * C + c) * KP * KP, buf.chunk_exp + (tile * C + c) * KP);
        if (!vec_apply_op<KP>(a, o)) fallbacks++;
      }
    }
#pragma unroll 1
    for (int slab = s_lo; slab < s_hi; ++slab) {
      const int bi = (slab - s_lo) & 1;
      if (bi == 0) {
        mbar_wait(&s_bar[0], ph0);
        ph0 ^= 1u;
      } else {
        mbar_wait(&s_bar[1], ph1);
        ph1 ^= 1u;
      }
      const double* se = reinterpret_cast<const double*>(bi ? stage1 : stage0);
#pragma unroll
      for (int tt = 0; tt < S; ++tt) {
        const int t = slab * S + tt;
        if (t < steps) {
          const uint64_t p = Layout::at(tile, c, t);
          const double* ev = se + (tt * C + c) * KP;
          double f[KP];
#pragma unroll
          for (int j = 0; j < KP; ++j) f[j] = 0.0;
#pragma unroll
          for (int k = 0; k < KP; ++k) {
#pragma unroll
            for (int j = 0; j < KP; ++j) f[j] = fma(a[k], m.A[k][j], f[j]);
          }
          if (kExact) {
            double fs = 0.0;
#pragma unroll
            for (int j = 0; j < KP; ++j) {
              f[j] *= ev[j];
              fs += f[j];
            }
            if (fs != 0.0) {  // FB.hpp:101-105
              const double inv = 1.0 / fs;
#pragma unroll
              for (int j = 0; j < KP; ++j) a[j] = f[j] * inv;
              if (kLoglik) {
                if (kSpec)
                  lognorm[p] = buf.maxE[p] + log(fs);
                else
                  ll += buf.maxE[p] + log(fs);
              }
            } else {          // FB.hpp:106-111: uniform fallback (the host then re-runs the sweep sequentially)
              fallbacks++;
#pragma unroll
              for (int j = 0; j < KP; ++j) a[j] = (j < K) ? 1.0 / (double)K : 0.0;
            }
          } else {
            double mxv = 0.0;
#pragma unroll
            for (int j = 0; j < KP; ++j) {
              f[j] *= ev[j];
              mxv = fmax(mxv, f[j]);
            }
            if (mxv > 0.0) {
              int e2 = exponent_of(mxv);
              if (e2 < -1000) e2 = -1000;
              const double sc = pow2i(-e2);
#pragma unroll
              for (int j = 0; j < KP; ++j) a[j] = f[j] * sc;
            } else {
              fallbacks++;
#pragma unroll
              for (int j = 0; j < KP; ++j) a[j] = (j < K) ? 1.0 / (double)K : 0.0;
            }
          }
#pragma unroll
          for (int j = 0; j < KP; ++j) buf.alpha[p * KP + j] = a[j];
        }
      }
      __syncwarp();
      if (slab + 2 < s_hi && lane == 0) {
        fence_proxy_async();
        issue_slab(tile, slab + 2, bi ? stage1 : stage0, bi ? &s_bar[1] : &s_bar[0]);
      }
    }
    __syncwarp();
  }
  if (kLoglik && !kSpec) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ll += shfl_xor_double(ll, o);
    if (lane == 0) buf.partials[blockIdx.x] = ll;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) fallbacks += __shfl_xor_sync(0xffffffffu, fallbacks, o);
  // a vanished forward sum under a guessed start proves nothing about the true recursion: the sweep is run again
  if (lane == 0 && fallbacks) atomicAdd(&buf.out_u64[kSpec ? spec_fail_slot<KP>() : KP + KP * KP], (unsigned long long)fallbacks);
}

// k_fwd_fixup_head (segment mode): what k_fwd_fixup cannot do for chunk 0 of a later rank.  Every rank publishes the
// last row it holds (+ its block count; all-gather of KP + 1 words, inside this kernel when the peer mailboxes are up),
// then restarts its first chunk from the last row of the nearest earlier rank that owns blocks.  The published row is
// right if the rank's own chunk 0 meets its guess before its last block (nothing behind chunk 0 then depended on the
// guess); a rank with a single piece cannot promise that and reports a failure.
//   phase 1: publish only (the caller runs the all-gather), 2: repair only, 3: publish + embedded exchange + repair
template <int KP, bool kExact>
__device__ __forceinline__ void fwd_fixup_head_cta(const SweepBuffers& buf, const ModelDev<KP>& m, int phase, int stride,
                                                   unsigned long long seq, int sub) {
  const int L = sub;  // the rank's first piece
  const uint64_t B = device_nblocks(buf.nblocks, buf.capacity);
  if (phase & 1) {
    if (threadIdx.x == 0) {
      const double* last = buf.alpha + (B ? Layout::perm(B - 1) : 0) * KP;
#pragma unroll
      for (int j = 0; j < KP; ++j) buf.seg.send_op[j] = B ? __ldcg(last + j) : 0.0;
      buf.seg.send_op[KP] = (double)B;
    }
    if (phase == 1) return;
    __threadfence();
    __syncthreads();
    p2p_exchange_cta(buf.seg.p2p, kSlotOps, seq, reinterpret_cast<const uint64_t*>(buf.seg.send_op), (uint32_t)(KP + 1),
                     reinterpret_cast<uint64_t*>(const_cast<double*>(buf.seg.ops)));
  }
  if (threadIdx.x != 0 || buf.seg.rank == 0 || B == 0) return;
  int src = buf.seg.rank - 1;
  while (src > 0 && !(__ldcg(buf.seg.ops + (size_t)src * stride + KP) > 0.0)) --src;  // rank 0 always owns block 0
  if (B <= (uint64_t)L) {  // one chunk: the row this rank published came straight from the guess
    atomicAdd(&buf.out_u64[spec_fail_slot<KP>()], 1ull);
    return;
  }
  double a[KP];
#pragma unroll
  for (int j = 0; j < KP; ++j) a[j] = __ldcg(buf.seg.ops + (size_t)src * stride + j);
#pragma unroll 1
  for (int t = 0; t < L; ++t) {
    const uint64_t p = Layout::at(0, 0, t);
    double f[KP], as[KP];
#pragma unroll
    for (int j = 0; j < KP; ++j) {
      f[j] = 0.0;
      as[j] = __ldcg(buf.alpha + p * KP + j);
    }
#pragma unroll
    for (int k = 0; k < KP; ++k) {
#pragma unroll
      for (int j = 0; j < KP; ++j) f[j] = fma(a[k], m.A[k][j], f[j]);
    }
    double fs = 0.0, mxv = 0.0;
#pragma unroll
    for (int j = 0; j < KP; ++j) {
      f[j] *= buf.e[p * KP + j];
      fs += f[j];
      mxv = fmax(mxv, f[j]);
    }
    bool met = false;
    if (fs > 0.0) {
      if (kExact) {
        const double inv = 1.0 / fs;
#pragma unroll
        for (int j = 0; j < KP; ++j) a[j] = f[j] * inv;
      } else {
        int e2 = exponent_of(mxv);
        if (e2 < -1000) e2 = -1000;
        const double sc = pow2i(-e2);
#pragma unroll
        for (int j = 0; j < KP; ++j) a[j] = f[j] * sc;
      }
      met = rows_parallel<KP>(a, as);
    }
    if (!(fs > 0.0) || (!met && t + 1 == L)) {
      atomicAdd(&buf.out_u64[spec_fail_slot<KP>()], 1ull);
      return;
    }
    if (met) return;
#pragma unroll
    for (int j = 0; j < KP; ++j) buf.alpha[p * KP + j] = a[j];
  }
}

template <int KP, bool kExact>
__global__ void __launch_bounds__(256) k_fwd_fixup_head(SweepBuffers buf, ModelDev<KP> m, int phase, int stride,
                                                        unsigned long long seq, int sub) {
  pdl_enter();
  fwd_fixup_head_cta<KP, kExact>(buf, m, phase, stride, seq, sub);
}

// ------------------------------------------------------------------------------------------------
// k_fwd_fixup: second pass of the speculative forward filter, thread per chunk.  Chunk g > 0 restarts from the stored
// last row of chunk g - 1 and rewrites its own rows until the new row is parallel to the stored one.  The last row of
// a chunk is never rewritten (the next chunk reads it in this same pass): a chunk that has not converged before it
// counts as a failure — except the very last chunk of the sequence, which nobody continues from.
//   kExact   rows are normalised by their sum (kept rows, log-likelihood), otherwise by a power of two
//   kLoglik  sums the per-block terms after the repair (partials[blockIdx.x])
//   kHead    split sequence over peer mailboxes: the CTA that arrives last publishes the rank's last row, exchanges and
//            repairs the rank's first chunk (fwd_fixup_head_cta) in the same launch
template <int KP, bool kExact, bool kLoglik, bool kHead = false>
__global__ void __launch_bounds__(128) k_fwd_fixup(SweepBuffers buf, ModelDev<KP> m, double* __restrict__ lognorm,
                                                   unsigned long long seq, int sub) {
  pdl_enter();
  constexpr int L = Layout::L, C = Layout::C;
  static_assert(!kLoglik || kExact, "the log-likelihood needs the forward sums");
  const uint64_t B = device_nblocks(buf.nblocks, buf.capacity);
  // pieces of `sub` blocks, the unit of the speculative pass; thread index = ((tile, piece), chunk): the 32 lanes of a
  // warp take the same piece of the 32 chunks of a tile, whose rows sit next to each other
  const int nsub = L / sub;
  const uint64_t ntiles = (B + Layout::TB - 1) / Layout::TB;
  const uint64_t nthreads = ntiles * (uint64_t)nsub * C;  // a multiple of 32
  const uint64_t rounded = (nthreads + blockDim.x - 1) / blockDim.x * blockDim.x;
  double ll = 0.0;
  unsigned fails = 0;
  for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < rounded; g += (uint64_t)gridDim.x * blockDim.x) {
    if (g >= nthreads) continue;
    const int c = (int)(g % C);
    const uint64_t unit = g / C, tile = unit / nsub;
    const int t_lo = (int)(unit % nsub) * sub;
    const uint64_t first = tile * Layout::TB + (uint64_t)c * L + t_lo;
    if (first >= B) continue;
    const int steps = (B - first) < (uint64_t)sub ? (int)(B - first) : sub;
    const bool last_chunk = first + steps == B;
    const bool have_entry = first > 0;  // (segment mode: piece 0 of a later rank is repaired by k_fwd_fixup_head)
    if (have_entry) {
      double a[KP];
      const double* ap = buf.alpha + Layout::perm(first - 1) * KP;
#pragma unroll
      for (int j = 0; j < KP; ++j) a[j] = ap[j];
#pragma unroll 1
      for (int t = 0; t < steps; ++t) {
        const uint64_t p = Layout::at(tile, c, t_lo + t);
        double ev[KP], as[KP];
#pragma unroll
        for (int j = 0; j < KP; ++j) {
          ev[j] = buf.e[p * KP + j];
          as[j] = buf.alpha[p * KP + j];
        }
        double f[KP];
#pragma unroll
        for (int j = 0; j < KP; ++j) f[j] = 0.0;
#pragma unroll
        for (int k = 0; k < KP; ++k) {
#pragma unroll
          for (int j = 0; j < KP; ++j) f[j] = fma(a[k], m.A[k][j], f[j]);
        }
        double fs = 0.0, mxv = 0.0;
#pragma unroll
        for (int j = 0; j < KP; ++j) {
          f[j] *= ev[j];
          fs += f[j];
          mxv = fmax(mxv, f[j]);
        }
        if (!(fs > 0.0)) {  // FB.hpp:106-111 would reset the filter here: not this pass's business
          fails++;
          break;
        }
        if (kExact) {
          const double inv = 1.0 / fs;
#pragma unroll
          for (int j = 0; j < KP; ++j) a[j] = f[j] * inv;
        } else {
          int e2 = exponent_of(mxv);
          if (e2 < -1000) e2 = -1000;
          const double sc = pow2i(-e2);
#pragma unroll
          for (int j = 0; j < KP; ++j) a[j] = f[j] * sc;
        }
        const bool met = rows_parallel<KP>(a, as);
        if (!met && t + 1 == steps && !last_chunk) {
          fails++;
          break;
        }
        if (kLoglik) lognorm[p] = buf.maxE[p] + log(fs);  // also at the meeting step: its stored term came from a wrong row
        if (met) break;
#pragma unroll
        for (int j = 0; j < KP; ++j) buf.alpha[p * KP + j] = a[j];
      }
    }
    if (kLoglik) {
      for (int t = 0; t < steps; ++t) ll += lognorm[Layout::at(tile, c, t_lo + t)];
    }
  }
  if (kLoglik) {
    __shared__ double s_ll[4];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ll += shfl_xor_double(ll, o);
    if ((threadIdx.x & 31) == 0) s_ll[threadIdx.x >> 5] = ll;
    __syncthreads();
    if (threadIdx.x == 0) buf.partials[blockIdx.x] = s_ll[0] + s_ll[1] + s_ll[2] + s_ll[3];
  }
  if (fails) atomicAdd(&buf.out_u64[spec_fail_slot<KP>()], (unsigned long long)fails);
  if constexpr (kHead) {
    if (last_cta_arrives(buf.tickets + kTicketFixup)) fwd_fixup_head_cta<KP, kExact>(buf, m, 3, KP + 1, seq, sub);
  }
}

// ------------------------------------------------------------------------------------------------
// k_fwd_tilescan: one CTA of 1024 threads.
//   step 1  thread (group, row): operator of each group of S consecutive tiles (row recursion, operators
//           prefetched one tile ahead for K <= 8)
//   step 2  warp 0 walks the group operators (shared memory for K <= 8): forward vector entering each group
//   step 3  warp per group (two groups per warp if G > 32): forward vector entering each tile

template <int KP>
struct ScanCfg {
  static constexpr int GMAX = 32;
  static constexpr bool SMEM = false;
};

// Segment mode splits the tile scan in two phases around the all-gather of the segment operators:
//   kPhase 1  group operators (kept in global memory) and this rank's segment operator -> seg.send_op
//   kPhase 2  forward vector entering the segment = pi * Op_0 * ... * Op_{rank-1} (normalised), then steps 2-3
//   kPhase 0  everything in one launch (single handle)
template <int KP>
__device__ __forceinline__ void load_gathered_op(OpVals<KP>& o, const double* g) {  // through L2: another CTA may have
#pragma unroll                                                                     // written it in this very kernel
  for (int k = 0; k < KP * KP; ++k) o.m[k] = __ldcg(g + k);
#pragma unroll
  for (int k = 0; k < KP; ++k) o.x[k] = (int)__ldcg(g + KP * KP + k);
}

// Small K (<= 8): 256 threads, every vector and operator in registers, operators prefetched one step
// ahead of the dependent chain.  G groups of S consecutive tiles, G ~ sqrt(#tiles):
//   step 1  thread (group, row): group operator = product of its S tile operators   -> shared memory
//   step 2  thread 0: forward vector entering each group (G sequential operator applications)
//   step 3  thread per group: forward vector entering each of its tiles
template <int KP, int kPhase>
__global__ void __launch_bounds__(256) k_fwd_tilescan_small(SweepBuffers buf, ModelDev<KP> m, int skip_upto,
                                                            unsigned long long seq) {
  pdl_enter();
  constexpr int GMAX = 256 / KP < 48 ? 256 / KP : 48;
  __shared__ double s_gain[GMAX][KP];
  __shared__ double s_gop[GMAX * KP * KP];
  __shared__ int s_gexp[GMAX * KP];
  const uint64_t B = device_nblocks(buf.nblocks, buf.capacity);
  const int nt = (int)((B + Layout::TB - 1) / Layout::TB);
  int G = 0, S = 1;
  if (nt > 0) {
    G = (int)ceil(sqrt((double)nt));
    if (G > GMAX) G = GMAX;
    S = (nt + G - 1) / G;
    G = (nt + S - 1) / S;
  }
  if (nt <= skip_upto) return;  // the cluster kernel took this sweep
  if (kPhase == 0 && nt == 0) return;
  if (kPhase != 2) {
    const int g = threadIdx.x / KP, i = threadIdx.x % KP;
    if (g < G) {
      double r[KP];
      int rex = 0;
#pragma unroll
      for (int j = 0; j < KP; ++j) r[j] = (j == i) ? 1.0 : 0.0;
      const int t0 = g * S, t1 = min(nt, (g + 1) * S);
      OpVals<KP> cur, nxt;
      load_op<KP>(cur, buf.tile_ops + (uint64_t)t0 * KP * KP, buf.tile_exp + (uint64_t)t0 * KP);
#pragma unroll 1
      for (int t = t0; t < t1; ++t) {
        const int tn = (t + 1 < t1) ? t + 1 : t;
        load_op<KP>(nxt, buf.tile_ops + (uint64_t)tn * KP * KP, buf.tile_exp + (uint64_t)tn * KP);
        row_times_op<KP, false>(r, rex, cur.m, cur.x);
        cur = nxt;
      }
#pragma unroll
      for (int j = 0; j < KP; ++j) s_gop[(g * KP + i) * KP + j] = r[j];
      s_gexp[g * KP + i] = rex;
      if (kPhase == 1) {
#pragma unroll
        for (int j = 0; j < KP; ++j) buf.group_ops[(g * KP + i) * KP + j] = r[j];
        buf.group_exp[g * KP + i] = rex;
      }
    }
    __syncthreads();
    if (kPhase == 1 || kPhase == 3) {
      // segment operator: row i of the ordered product of the group operators (identity without blocks)
      if (threadIdx.x < KP) {
        const int i = threadIdx.x;
        double r[KP];
        int rex = 0;
#pragma unroll
        for (int j = 0; j < KP; ++j) r[j] = (j == i) ? 1.0 : 0.0;
#pragma unroll 1
        for (int g = 0; g < G; ++g) row_times_op<KP, false>(r, rex, s_gop + g * KP * KP, s_gexp + g * KP);
#pragma unroll
        for (int j = 0; j < KP; ++j) buf.seg.send_op[i * KP + j] = r[j];
        buf.seg.send_op[KP * KP + i] = (double)rex;
      }
      if (kPhase == 1) return;
      __threadfence();
      __syncthreads();
      p2p_exchange_cta(buf.seg.p2p, kSlotOps, seq, reinterpret_cast<const uint64_t*>(buf.seg.send_op), KP * KP + KP,
                       reinterpret_cast<uint64_t*>(const_cast<double*>(buf.seg.ops)));
      if (nt == 0) return;
    }
  } else {
    if (nt == 0) return;
    for (int k = threadIdx.x; k < G * KP * KP; k += blockDim.x) s_gop[k] = buf.group_ops[k];
    for (int k = threadIdx.x; k < G * KP; k += blockDim.x) s_gexp[k] = buf.group_exp[k];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double a[KP];
#pragma unroll
    for (int j = 0; j < KP; ++j) a[j] = m.pi[j];  // row 0 of the trellis is pi itself (FB.hpp:57)
    if (kPhase >= 2) {
#pragma unroll 1
      for (int r = 0; r < buf.seg.rank; ++r) {
        OpVals<KP> o;
        load_gathered_op<KP>(o, buf.seg.ops + (size_t)r * (KP * KP + KP));
        if (buf.seg.heads[kHeadWords * r] > 0.0 && !vec_apply_op<KP>(a, o)) atomicAdd(&buf.out_u64[KP + KP * KP], 1ull);
      }
    }
#pragma unroll 1
    for (int g = 0; g < G; ++g) {
#pragma unroll
      for (int j = 0; j < KP; ++j) s_gain[g][j] = a[j];
      if (g + 1 < G) {
        OpVals<KP> o;
        load_op<KP>(o, s_gop + g * KP * KP, s_gexp + g * KP);
        if (!vec_apply_op<KP>(a, o)) atomicAdd(&buf.out_u64[KP + KP * KP], 1ull);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < G) {
    const int g = threadIdx.x;
    double a[KP];
#pragma unroll
    for (int j = 0; j < KP; ++j) a[j] = s_gain[g][j];
    const int t0 = g * S, t1 = min(nt, (g + 1) * S);
    OpVals<KP> cur, nxt;
    load_op<KP>(cur, buf.tile_ops + (uint64_t)t0 * KP * KP, buf.tile_exp + (uint64_t)t0 * KP);
#pragma unroll 1
    for (int t = t0; t < t1; ++t) {
#pragma unroll
      for (int j = 0; j < KP; ++j) buf.tile_ain[(uint64_t)t * KP + j] = a[j];
      const int tn = (t + 1 < t1) ? t + 1 : t;
      load_op<KP>(nxt, buf.tile_ops + (uint64_t)tn * KP * KP, buf.tile_exp + (uint64_t)tn * KP);
      if (t + 1 < t1 && !vec_apply_op<KP>(a, cur)) atomicAdd(&buf.out_u64[KP + KP * KP], 1ull);
      cur = nxt;
    }
  }
}

// Small K, cluster version: 8 CTAs of one thread-block cluster split the tile operators among them and keep
// their share in shared memory (one coalesced load with every access in flight), so that none of the serial walks
// below ever waits on global memory:
//   step 1  thread (subgroup, row): operator of each subgroup of S consecutive tiles, S ~ sqrt(own tiles)
//   step 2  pairwise tree over the subgroup operators -> the CTA's operator, published in shared memory
//   step 3  after a cluster barrier, thread 0 of CTA r applies the operators of CTAs 0..r-1 (distributed shared
//           memory) to the start vector and walks its own subgroups
//   step 4  thread per subgroup: forward vector entering each of its tiles
// Handles up to kClusterCtas * CAP tiles; larger sweeps are left to k_fwd_tilescan_small, launched right after.
constexpr int kClusterCtas = 8;
template <int KP>
struct ClusterScanCfg {
  static constexpr int CAP = KP <= 5 ? 512 : (KP <= 6 ? 384 : 224);  // tiles per CTA
  static constexpr int SGMAX = 32;
  static constexpr size_t kSmem = (size_t)CAP * (KP * KP * sizeof(double) + KP * sizeof(int));
  static constexpr int kMaxTiles = kClusterCtas * CAP;
};

template <int KP, int kPhase>
__global__ void __cluster_dims__(kClusterCtas, 1, 1) __launch_bounds__(256)
    k_fwd_tilescan_cluster(SweepBuffers buf, ModelDev<KP> m, unsigned long long seq) {
  pdl_enter();
  namespace cg = cooperative_groups;
  using Cfg = ClusterScanCfg<KP>;
  constexpr int SGMAX = Cfg::SGMAX;
  extern __shared__ __align__(16) unsigned char s_dyn[];
  double* s_op = reinterpret_cast<double*>(s_dyn);
  int* s_ex = reinterpret_cast<int*>(s_op + (size_t)Cfg::CAP * KP * KP);
  __shared__ double s_gop[SGMAX * KP * KP];   // subgroup operators (kept for step 3)
  __shared__ int s_gex[SGMAX * KP];
  __shared__ double s_tree[SGMAX * KP * KP];  // working copy consumed by the tree
  __shared__ int s_treex[SGMAX * KP];
  __shared__ double s_cop[KP * KP];           // this CTA's operator, read by the other CTAs of the cluster
  __shared__ int s_cex[KP];
  __shared__ double s_gain[SGMAX][KP];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int tid = threadIdx.x;
  const uint64_t B = device_nblocks(buf.nblocks, buf.capacity);
  const int nt = (int)((B + Layout::TB - 1) / Layout::TB);
  if (nt > Cfg::kMaxTiles) return;  // uniform over the cluster: nobody reaches a cluster barrier
  const int per = (nt + kClusterCtas - 1) / kClusterCtas;
  const int t0 = min(nt, rank * per), t1 = min(nt, t0 + per), n_own = t1 - t0;
  // ---- step 0
  {
    const double* src = buf.tile_ops + (uint64_t)t0 * KP * KP;
    for (int i = tid; i < n_own * KP * KP; i += 256) s_op[i] = src[i];
    const int* srx = buf.tile_exp + (uint64_t)t0 * KP;
    for (int i = tid; i < n_own * KP; i += 256) s_ex[i] = srx[i];
  }
  int S = 1, SG = 0;
  if (n_own > 0) {
    S = (int)ceil(sqrt((double)n_own));
    if (S * SGMAX < n_own) S = (n_own + SGMAX - 1) / SGMAX;
    SG = (n_own + S - 1) / S;
  }
  __syncthreads();
  // ---- step 1
  {
    const int g = tid / KP, i = tid % KP;
    if (g < SG) {
      double r[KP];
      int rex = 0;
#pragma unroll
      for (int j = 0; j < KP; ++j) r[j] = (j == i) ? 1.0 : 0.0;
      const int a = g * S, b = min(n_own, a + S);
#pragma unroll 1
      for (int t = a; t < b; ++t) row_times_op<KP, false>(r, rex, s_op + t * KP * KP, s_ex + t * KP);
#pragma unroll
      for (int j = 0; j < KP; ++j) s_gop[(g * KP + i) * KP + j] = s_tree[(g * KP + i) * KP + j] = r[j];
      s_gex[g * KP + i] = s_treex[g * KP + i] = rex;
    }
  }
  __syncthreads();
  // ---- step 2
  for (int stride = 1; stride < SG; stride <<= 1) {
    const int pr = tid / KP, row = tid % KP;
    const int a = pr * 2 * stride, b = a + stride;
    if (b < SG) {
      double r[KP];
#pragma unroll
      for (int j = 0; j < KP; ++j) r[j] = s_tree[(a * KP + row) * KP + j];
      int rex = s_treex[a * KP + row];
      row_times_op<KP, false>(r, rex, s_tree + b * KP * KP, s_treex + b * KP);
#pragma unroll
      for (int j = 0; j < KP; ++j) s_tree[(a * KP + row) * KP + j] = r[j];
      s_treex[a * KP + row] = rex;
    }
    __syncthreads();
  }
  if (tid < KP) {
#pragma unroll
    for (int j = 0; j < KP; ++j) s_cop[tid * KP + j] = SG ? s_tree[tid * KP + j] : (j == tid ? 1.0 : 0.0);
    s_cex[tid] = SG ? s_treex[tid] : 0;
  }
  cluster.sync();
  if (kPhase == 1 || kPhase == 3) {
    // segment operator: row i of the ordered product of the CTA operators (identity without blocks)
    if (rank == 0 && tid < KP) {
      const int i = tid;
      double r[KP];
      int rex = 0;
#pragma unroll
      for (int j = 0; j < KP; ++j) r[j] = (j == i) ? 1.0 : 0.0;
#pragma unroll 1
      for (int p = 0; p < kClusterCtas; ++p) {
        if (min(nt, p * per) >= nt) break;  // CTA p and the later ones own no tile
        row_times_op<KP, false>(r, rex, cluster.map_shared_rank(s_cop, p), cluster.map_shared_rank(s_cex, p));
      }
#pragma unroll
      for (int j = 0; j < KP; ++j) buf.seg.send_op[i * KP + j] = r[j];
      buf.seg.send_op[KP * KP + i] = (double)rex;
    }
    if (kPhase == 3 && rank == 0) {
      // the collective runs right here: CTA 0 trades segment operators with the other GPUs while the rest of
      // the cluster waits at the barrier below
      __threadfence();
      __syncthreads();
      p2p_exchange_cta(buf.seg.p2p, kSlotOps, seq, reinterpret_cast<const uint64_t*>(buf.seg.send_op), KP * KP + KP,
                       reinterpret_cast<uint64_t*>(const_cast<double*>(buf.seg.ops)));
    }
    cluster.sync();  // shared memory of every CTA stays alive until CTA 0 has read it
    if (kPhase == 1) return;
  }
  // ---- step 3
  if (tid == 0 && n_own > 0) {
    double a[KP];
#pragma unroll
    for (int j = 0; j < KP; ++j) a[j] = m.pi[j];  // row 0 of the trellis is pi itself (FB.hpp:57)
    if (kPhase >= 2) {
#pragma unroll 1
      for (int r = 0; r < buf.seg.rank; ++r) {
        OpVals<KP> o;
        load_gathered_op<KP>(o, buf.seg.ops + (size_t)r * (KP * KP + KP));
        if (buf.seg.heads[kHeadWords * r] > 0.0 && !vec_apply_op<KP>(a, o)) atomicAdd(&buf.out_u64[KP + KP * KP], 1ull);
      }
    }
#pragma unroll 1
    for (int p = 0; p < rank; ++p) {
      OpVals<KP> o;
      load_op<KP>(o, cluster.map_shared_rank(s_cop, p), cluster.map_shared_rank(s_cex, p));
      if (!vec_apply_op<KP>(a, o)) atomicAdd(&buf.out_u64[KP + KP * KP], 1ull);
    }
#pragma unroll 1
    for (int g = 0; g < SG; ++g) {
#pragma unroll
      for (int j = 0; j < KP; ++j) s_gain[g][j] = a[j];
      if (g + 1 < SG) {
        OpVals<KP> o;
        load_op<KP>(o, s_gop + g * KP * KP, s_gex + g * KP);
        if (!vec_apply_op<KP>(a, o)) atomicAdd(&buf.out_u64[KP + KP * KP], 1ull);
      }
    }
  }
  __syncthreads();
  // ---- step 4
  if (tid < SG) {
    double a[KP];
#pragma unroll
    for (int j = 0; j < KP; ++j) a[j] = s_gain[tid][j];
    const int lo = tid * S, hi = min(n_own, lo + S);
#pragma unroll 1
    for (int t = lo; t < hi; ++t) {
#pragma unroll
      for (int j = 0; j < KP; ++j) buf.tile_ain[(uint64_t)(t0 + t) * KP + j] = a[j];
      if (t + 1 < hi) {
        OpVals<KP> o;
        load_op<KP>(o, s_op + t * KP * KP, s_ex + t * KP);
        if (!vec_apply_op<KP>(a, o)) atomicAdd(&buf.out_u64[KP + KP * KP], 1ull);
      }
    }
  }
  cluster.sync();  // no CTA leaves while another may still read its operator
}

template <int KP, int kPhase>
__global__ void __launch_bounds__(1024) k_fwd_tilescan(SweepBuffers buf, ModelDev<KP> m) {
  constexpr int GMAX = ScanCfg<KP>::GMAX;
  __shared__ double s_gain[GMAX][KP];
  const uint64_t B = device_nblocks(buf.nblocks, buf.capacity);
  const int nt = (int)((B + Layout::TB - 1) / Layout::TB);
  if (kPhase != 1 && nt == 0) return;
  int G = 0, S = 1;
  if (nt > 0) {
    G = (int)ceil(sqrt((double)nt));
    if (G > GMAX) G = GMAX;
    if (G > 1024 / KP) G = 1024 / KP;
    S = (nt + G - 1) / G;
    G = (nt + S - 1) / S;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* gop = buf.group_ops;
  int* gexp = buf.group_exp;
  // ---- step 1
  if (kPhase != 2) {
    const int g = threadIdx.x / KP, i = threadIdx.x % KP;
    if (g < G) {
      double r[KP];
      int rex = 0;
#pragma unroll
      for (int j = 0; j < KP; ++j) r[j] = (j == i) ? 1.0 : 0.0;
      const int t0 = g * S, t1 = min(nt, (g + 1) * S);
#pragma unroll 1
      for (int t = t0; t < t1; ++t)
        row_times_op<KP, false>(r, rex, buf.tile_ops + (uint64_t)t * KP * KP, buf.tile_exp + (uint64_t)t * KP);
#pragma unroll
      for (int j = 0; j < KP; ++j) gop[(g * KP + i) * KP + j] = r[j];
      gexp[g * KP + i] = rex;
    }
    __threadfence_block();
    __syncthreads();
    if (kPhase == 1) {
      if (threadIdx.x < KP) {
        const int i = threadIdx.x;
        double r[KP];
        int rex = 0;
#pragma unroll
        for (int j = 0; j < KP; ++j) r[j] = (j == i) ? 1.0 : 0.0;
#pragma unroll 1
        for (int g = 0; g < G; ++g) row_times_op<KP, true>(r, rex, gop + g * KP * KP, gexp + g * KP);
#pragma unroll
        for (int j = 0; j < KP; ++j) buf.seg.send_op[i * KP + j] = r[j];
        buf.seg.send_op[KP * KP + i] = (double)rex;
      }
      return;
    }
  }
  // ---- step 2
  if (warp == 0) {
    double a = lane < KP ? m.pi[lane] : 0.0;  // row 0 of the trellis is pi itself (FB.hpp:57)
    if (kPhase == 2) {
#pragma unroll 1
      for (int r = 0; r < buf.seg.rank; ++r) {
        const double* go = buf.seg.ops + (size_t)r * (KP * KP + KP);
        LaneOp<KP> o;
#pragma unroll
        for (int k = 0; k < KP; ++k) o.col[k] = lane < KP ? go[k * KP + lane] : 0.0;
        o.x = lane < KP ? (int)go[KP * KP + lane] : kDeadExp;
        if (buf.seg.heads[kHeadWords * r] > 0.0 && !warp_apply_op<KP>(a, o) && lane == 0)
          atomicAdd(&buf.out_u64[KP + KP * KP], 1ull);
      }
    }
    LaneOp<KP> cur, nxt;
    load_lane_op<KP>(cur, gop, gexp, lane);
#pragma unroll 1
    for (int g = 0; g < G; ++g) {
      if (lane < KP) s_gain[g][lane] = a;
      const int gn = (g + 1 < G) ? g + 1 : g;
      load_lane_op<KP>(nxt, gop + gn * KP * KP, gexp + gn * KP, lane);
      if (g + 1 < G && !warp_apply_op<KP>(a, cur) && lane == 0) atomicAdd(&buf.out_u64[KP + KP * KP], 1ull);
      cur = nxt;
    }
  }
  __syncthreads();
  // ---- step 3
  for (int g = warp; g < G; g += 32) {
    double a = lane < KP ? s_gain[g][lane] : 0.0;
    const int t0 = g * S, t1 = min(nt, (g + 1) * S);
    LaneOp<KP> cur, nxt;
    load_lane_op<KP>(cur, buf.tile_ops + (uint64_t)t0 * KP * KP, buf.tile_exp + (uint64_t)t0 * KP, lane);
#pragma unroll 1
    for (int t = t0; t < t1; ++t) {
      if (lane < KP) buf.tile_ain[(uint64_t)t * KP + lane] = a;
      const int tn = (t + 1 < t1) ? t + 1 : t;
      load_lane_op<KP>(nxt, buf.tile_ops + (uint64_t)tn * KP * KP, buf.tile_exp + (uint64_t)tn * KP, lane);
      if (t + 1 < t1 && !warp_apply_op<KP>(a, cur) && lane == 0) atomicAdd(&buf.out_u64[KP + KP * KP], 1ull);
      cur = nxt;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// k_fwd_replay: one warp per tile; lane c owns chunk c.  Only the inherently sequential part of the
// filter runs here: alpha_t = normalise(e_t o (alpha_{t-1} A)), written per block (interleaved order).

//   kSpec: speculative pass (see spec_entry / k_fwd_fixup): no operators, the entering vectors are guesses
template <int KP, bool kLoglik, bool kSpec = false>
__global__ void __launch_bounds__(32) k_fwd_replay(SweepBuffers buf, ModelDev<KP> m, double* __restrict__ lognorm, int warm) {
  pdl_enter();
  constexpr int L = Layout::L, C = Layout::C;
  __shared__ double s_ain[C][KP + 1];
  const uint64_t B = device_nblocks(buf.nblocks, buf.capacity);
  const uint64_t ntiles = (B + Layout::TB - 1) / Layout::TB;
  const int lane = threadIdx.x;
  const int K = m.K;
  double ll = 0.0;
  unsigned fallbacks = 0;
  for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    // ---- vector entering each chunk (warp-cooperative walk over the 32 chunk operators)
    if constexpr (!kSpec) {
      double a = lane < KP ? buf.tile_ain[tile * KP + lane] : 0.0;
      LaneOp<KP> cur, nxt;
      load_lane_op<KP>(cur, buf.chunk_ops + tile * C * KP * KP, buf.chunk_exp + tile * C * KP, lane);
#pragma unroll 1
      for (int c = 0; c < C; ++c) {
        if (lane < KP) s_ain[c][lane] = a;
        const uint64_t ch = tile * C + c;
        const uint64_t chn = (c + 1 < C) ? ch + 1 : ch;
        load_lane_op<KP>(nxt, buf.chunk_ops + chn * KP * KP, buf.chunk_exp + chn * KP, lane);
        if (c + 1 < C && (ch + 1) * L < B) {
          if (!warp_apply_op<KP>(a, cur)) fallbacks += (lane == 0);
        }
        cur = nxt;
      }
    }
    __syncwarp();
    // ---- replay of chunk `lane`
    const int c = lane;
    const uint64_t first = tile * Layout::TB + (uint64_t)c * L;
    int steps = 0;
    if (first < B) steps = (B - first) < (uint64_t)L ? (int)(B - first) : L;
    double a[KP];
    if constexpr (kSpec) {
#pragma unroll
      for (int j = 0; j < KP; ++j) a[j] = 0.0;
      if (steps > 0) spec_entry<KP>(buf, m, first, warm, a);
    } else {
#pragma unroll
      for (int j = 0; j < KP; ++j) a[j] = s_ain[c][j];
    }
    const double* ep = buf.e + Layout::at(tile, c, 0) * KP;
    double en[KP];
#pragma unroll
    for (int j = 0; j < KP; ++j) en[j] = steps > 0 ? ep[j] : 0.0;
    double mxn = (kLoglik && steps > 0) ? buf.maxE[Layout::at(tile, c, 0)] : 0.0;
#pragma unroll 1
    for (int t = 0; t < steps; ++t) {
      const uint64_t p = Layout::at(tile, c, t);
      double ev[KP];
#pragma unroll
      for (int j = 0; j < KP; ++j) ev[j] = en[j];
      const double mx = mxn;
      if (t + 1 < steps) {
#pragma unroll
        for (int j = 0; j < KP; ++j) en[j] = ep[(uint64_t)(t + 1) * C * KP + j];
        if (kLoglik) mxn = buf.maxE[p + C];
      }
      double f[KP];
#pragma unroll
      for (int j = 0; j < KP; ++j) f[j] = 0.0;
#pragma unroll
      for (int k = 0; k < KP; ++k) {
#pragma unroll
        for (int j = 0; j < KP; ++j) f[j] = fma(a[k], m.A[k][j], f[j]);
      }
      double fs = 0.0;
#pragma unroll
      for (int j = 0; j < KP; ++j) {
        f[j] *= ev[j];
        fs += f[j];
      }
      if (fs != 0.0) {  // FB.hpp:101-105
        const double inv = 1.0 / fs;
#pragma unroll
        for (int j = 0; j < KP; ++j) a[j] = f[j] * inv;
        if (kLoglik) {
          if (kSpec)
            lognorm[p] = mx + log(fs);
          else
            ll += mx + log(fs);
        }
      } else {          // FB.hpp:106-111: uniform fallback (the host then re-runs the sweep sequentially)
        fallbacks++;
#pragma unroll
        for (int j = 0; j < KP; ++j) a[j] = (j < K) ? 1.0 / (double)K : 0.0;
      }
#pragma unroll
      for (int j = 0; j < KP; ++j) buf.alpha[p * KP + j] = a[j];
    }
    __syncwarp();
  }
  if (kLoglik && !kSpec) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ll += shfl_xor_double(ll, o);
    if (lane == 0) buf.partials[blockIdx.x] = ll;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) fallbacks += __shfl_xor_sync(0xffffffffu, fallbacks, o);
  if (lane == 0 && fallbacks) atomicAdd(&buf.out_u64[kSpec ? spec_fail_slot<KP>() : KP + KP * KP], (unsigned long long)fallbacks);
}

// Sequential forward recursion (one thread): the exact reference recursion, used only when the
// parallel pass reported a uniform fallback, whose effect on later blocks the operator products
// cannot express.
// Segment mode: ranks run it one after the other; `ain` is the final vector of the previous rank and
// the rank's own final vector goes to seg.send_op[0..KP) for the next one.
template <int KP, bool kLoglik>
__global__ void __launch_bounds__(32) k_fwd_sequential(SweepBuffers buf, ModelDev<KP> m, const double* ain) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const uint64_t B = device_nblocks(buf.nblocks, buf.capacity);
  const int K = m.K;
  double a[KP];
#pragma unroll
  for (int j = 0; j < KP; ++j) a[j] = ain ? ain[j] : m.pi[j];
  double ll = 0.0;
  unsigned long long fallbacks = 0;
  for (uint64_t b = 0; b < B; ++b) {
    const uint64_t p = Layout::perm(b);
    double f[KP];
#pragma unroll
    for (int j = 0; j < KP; ++j) f[j] = 0.0;
#pragma unroll
    for (int k = 0; k < KP; ++k) {
#pragma unroll
      for (int j = 0; j < KP; ++j) f[j] = fma(a[k], m.A[k][j], f[j]);
    }
    double fs = 0.0;
#pragma unroll
    for (int j = 0; j < KP; ++j) {
      f[j] *= buf.e[p * KP + j];
      fs += f[j];
    }
    if (fs != 0.0) {
      const double inv = 1.0 / fs;
#pragma unroll
      for (int j = 0; j < KP; ++j) a[j] = f[j] * inv;
      if (kLoglik) ll += buf.maxE[p] + log(fs);
    } else {
      fallbacks++;
#pragma unroll
      for (int j = 0; j < KP; ++j) a[j] = (j < K) ? 1.0 / (double)K : 0.0;
    }
#pragma unroll
    for (int j = 0; j < KP; ++j) buf.alpha[p * KP + j] = a[j];
  }
  if (kLoglik) buf.partials[0] = ll;
  buf.out_u64[KP + KP * KP] = fallbacks;
  if (buf.seg.world > 1) {
#pragma unroll
    for (int j = 0; j < KP; ++j) buf.seg.send_op[j] = a[j];
  }
}

// k_bwd_maps: thread per block.  Given alpha_t, everything the backward pass will do at block t is a
// function of the state q_{t+1} it arrives with: the map j -> discrete_distribution(alpha'_t(.) A(., j))(u_t)
// with alpha'_t = alpha_t * A_ss^(N_t - 1) (FB.hpp:115-119,145-146); the last block draws from alpha_B
// itself (FB.hpp:138).  u_t is the block's counter-based Philox uniform, or the replayed one.
// the map of the block in storage slot p (natural index b): j -> state drawn at this block if its successor is in state j
template <int KP, bool kRows>
__device__ __forceinline__ Map<KP> block_map(const SweepBuffers& buf, const ModelDev<KP>& m, uint64_t p, uint64_t b, uint64_t B,
                                            uint64_t seed, uint64_t sweep, const double* s_tab) {
  const int K = m.K;
  Map<KP> fm = Map<KP>::identity();  // slots past the last block compose as the identity
  if (b < B) {
    const bool last = (b + 1 == B) && !(buf.seg.world > 1 && seg_later_blocks(buf.seg));
    const uint64_t gb = buf.seg.world > 1 ? seg_first_block(buf.seg) + b : b;  // global block index
    // alpha'_t = alpha_t * A_ss^(N_t - 1): the rescale FB.hpp:115-119 applies to a row once its successor exists
    const double Nm1 = (double)buf.bN[p] - 1.0;
    double ap[KP];
#pragma unroll
    for (int j = 0; j < KP; ++j) {
      const double al = buf.alpha[p * KP + j];
      ap[j] = (last || !m.use_self || j >= K) ? al : al * exp_nonpos(Nm1 * m.loga[j], s_tab);
    }
    if (kRows) {
      if (b == 0)
        for (int j = 0; j < K; ++j) buf.rows[j] = m.pi[j];
      for (int j = 0; j < K; ++j) buf.rows[(b + 1) * K + j] = ap[j];
    }
    const double u = buf.replay_u ? buf.replay_u[seg_global_blocks(buf.seg, B) - 1 - gb]
                                  : Philox::uniform(seed, sweep, 0u, gb);
    fm = Map<KP>::zero();
    if (last) {
      const uint32_t q = discrete_draw<KP>(ap, K, u);
#pragma unroll
      for (int j = 0; j < KP; ++j) fm.set(j, q);
    } else {
#pragma unroll
      for (int j = 0; j < KP; ++j) {
        if (j < K) {
          // running sums of alpha'_t(k) A(k, j) (rows and columns of padded states are zero)
          double c[KP];
          c[0] = ap[0] * m.A[0][j];
#pragma unroll
          for (int k = 1; k < KP; ++k) c[k] = fma(ap[k], m.A[k][j], c[k - 1]);
          bool tie;
          uint32_t q = discrete_draw_fast<KP>(c, K, u, tie);
          if (tie) {  // the reference's own arithmetic: rounded products (FB.hpp:145-146), then discrete_distribution
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
  return fm;
}

template <int KP, bool kRows>
__global__ void __launch_bounds__(256, KP <= 5 ? 5 : 1) k_bwd_maps(SweepBuffers buf, ModelDev<KP> m, uint64_t seed, uint64_t sweep) {
  pdl_enter();
  __shared__ double s_tab[64];  // 2^(j/64) for exp_nonpos
  exp_table_load(s_tab);
  __syncthreads();
  const uint64_t B = device_nblocks(buf.nblocks, buf.capacity);
  const uint64_t slots = (B + Layout::TB - 1) / Layout::TB * Layout::TB;
  for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < slots; p += (uint64_t)gridDim.x * blockDim.x) {
    const Map<KP> fm = block_map<KP, kRows>(buf, m, p, Layout::inv(p), B, seed, sweep, s_tab);
    fm.store(buf.maps + p * (8 * Map<KP>::W));
  }
}

// ------------------------------------------------------------------------------------------------
// k_bwd_scan: one CTA; suffix composition of the tile maps.
// tile_qin[t] = state of the first block after tile t (irrelevant for the last tile, whose last block
// carries a constant map).

// kSegMap: only the composed map of the whole segment is wanted (-> seg.send_map, identity without
// blocks); otherwise the state following the segment comes from the gathered maps of the later ranks
// (the last block of the sequence carries a constant map, so the start value is irrelevant).
// kMode 0: resolve (gathered maps of the later ranks, if any); 1: segment map only; 2: segment map, map
// exchange through the peer mailboxes and resolution in one kernel
// NT threads of one CTA; the tile maps are read through L2 (the chunk-map kernel's last CTA calls this in the same launch)
template <int KP, int kMode, int NT>
__device__ __forceinline__ void bwd_scan_cta(const SweepBuffers& buf, unsigned long long seq) {
  constexpr bool kSegMap = kMode == 1;
  constexpr int MB = 8 * Map<KP>::W, NW = NT / 32;
  __shared__ uint64_t s_w[NW][Map<KP>::W];
  const uint64_t B = device_nblocks(buf.nblocks, buf.capacity);
  const int64_t nt = (int64_t)((B + Layout::TB - 1) / Layout::TB);
  if (nt == 0) {
    if (kMode != 0 && threadIdx.x == 0) {
      const Map<KP> id = Map<KP>::identity();
#pragma unroll
      for (int i = 0; i < 4; ++i) buf.seg.send_map[i] = i < Map<KP>::W ? id.w[i] : 0ull;
    }
    if (kMode == 2) {  // a rank without blocks still takes part in the collective
      __threadfence();
      __syncthreads();
      p2p_exchange_cta(buf.seg.p2p, kSlotMaps, seq, reinterpret_cast<const uint64_t*>(buf.seg.send_map), 4,
                       const_cast<uint64_t*>(buf.seg.maps));
    }
    return;
  }
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t per = (nt + NT - 1) / NT;
  // thread `tid` owns tiles [lo, hi); threads are ordered by position, so thread 0 holds the earliest
  const int64_t lo = min(nt, (int64_t)tid * per), hi = min(nt, lo + per);
  Map<KP> own = Map<KP>::identity();  // own = f_lo o f_{lo+1} o ... o f_{hi-1}
  for (int64_t t = lo; t < hi; ++t) own = own.after(Map<KP>::load_cg(buf.tile_maps + t * MB));
  // inclusive suffix within the warp by shuffles
  Map<KP> inc = own;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    Map<KP> other;
#pragma unroll
    for (int i = 0; i < Map<KP>::W; ++i) other.w[i] = __shfl_down_sync(0xffffffffu, inc.w[i], o);
    if (lane + o < 32) inc = inc.after(other);
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < Map<KP>::W; ++i) s_w[warp][i] = inc.w[i];
  }
  __syncthreads();
  if (kMode != 0) {
    if (tid == 0) {
      Map<KP> tot = Map<KP>::identity();
      for (int wv = 0; wv < NW; ++wv) {
        Map<KP> o;
#pragma unroll
        for (int i = 0; i < Map<KP>::W; ++i) o.w[i] = s_w[wv][i];
        tot = tot.after(o);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) buf.seg.send_map[i] = i < Map<KP>::W ? tot.w[i] : 0ull;
    }
    if (kSegMap) return;
    __threadfence();
    __syncthreads();
    p2p_exchange_cta(buf.seg.p2p, kSlotMaps, seq, reinterpret_cast<const uint64_t*>(buf.seg.send_map), 4,
                     const_cast<uint64_t*>(buf.seg.maps));
  }
  uint32_t q_end = 0;
  for (int r = buf.seg.world - 1; r > buf.seg.rank; --r) {
    Map<KP> o;
#pragma unroll
    for (int i = 0; i < Map<KP>::W; ++i) o.w[i] = __ldcg(buf.seg.maps + 4 * r + i);
    q_end = o.get(q_end);
  }
  Map<KP> after_warp = Map<KP>::identity();  // map of everything after this warp
  for (int wv = warp + 1; wv < NW; ++wv) {
    Map<KP> o;
#pragma unroll
    for (int i = 0; i < Map<KP>::W; ++i) o.w[i] = s_w[wv][i];
    after_warp = after_warp.after(o);
  }
  Map<KP> nxt;
#pragma unroll
  for (int i = 0; i < Map<KP>::W; ++i) nxt.w[i] = __shfl_down_sync(0xffffffffu, inc.w[i], 1);
  const Map<KP> suf = (lane == 31) ? after_warp : nxt.after(after_warp);
  // on the rank holding the sequence's last block the suffix map is constant in its argument
  uint32_t q = suf.get(q_end);
  for (int64_t t = hi - 1; t >= lo; --t) {
    buf.tile_qin[t] = (uint8_t)q;
    q = Map<KP>::load_cg(buf.tile_maps + t * MB).get(q);
  }
}

template <int KP, int kMode>
__global__ void __launch_bounds__(1024) k_bwd_scan(SweepBuffers buf, unsigned long long seq) {
  pdl_enter();
  bwd_scan_cta<KP, kMode, 1024>(buf, seq);
}

// k_bwd_chunkmaps: warp per tile, lane per chunk: G_c = f_first o ... o f_last of the chunk, then a
// suffix scan over the warp: X_c = G_{c+1} o ... o G_31 (per chunk) and the tile map G_0 o ... o G_31.
// kScan -1: chunk and tile maps only; 0 / 2: the CTA that arrives last also runs the scan over the tile maps
// (bwd_scan_cta<kScan>: 0 resolve, 2 with the map exchange of a split sequence), saving the follow-up launch
template <int KP, int kScan>
__global__ void __launch_bounds__(128) k_bwd_chunkmaps(SweepBuffers buf, unsigned long long seq) {
  pdl_enter();
  constexpr int L = Layout::L, C = Layout::C, MB = 8 * Map<KP>::W;
  static_assert(L == 32, "quarter chunks of 8 blocks");
  const uint64_t B = device_nblocks(buf.nblocks, buf.capacity);
  const uint64_t ntiles = (B + Layout::TB - 1) / Layout::TB;
  const int lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t tile = warp; tile < ntiles; tile += nwarps) {
    // from the chunk's last block down: G = f_t o (f_{t+1} o ... o f_31); the maps of the last 24, 16 and 8 blocks are
    // kept for the replay, whose threads take a quarter chunk each
    Map<KP> G = Map<KP>::identity();
#pragma unroll
    for (int q4 = 3; q4 >= 0; --q4) {
#pragma unroll
      for (int t = 8 * q4 + 7; t >= 8 * q4; --t) G = Map<KP>::load(buf.maps + Layout::at(tile, lane, t) * MB).after(G);
      if (q4 > 0) G.store(buf.chunk_submaps + ((tile * C + lane) * 3 + (q4 - 1)) * MB);
    }
    Map<KP> inc = G;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      Map<KP> other;
#pragma unroll
      for (int i = 0; i < Map<KP>::W; ++i) other.w[i] = __shfl_down_sync(0xffffffffu, inc.w[i], o);
      if (lane + o < 32) inc = inc.after(other);
    }
    Map<KP> excl;
#pragma unroll
    for (int i = 0; i < Map<KP>::W; ++i) excl.w[i] = __shfl_down_sync(0xffffffffu, inc.w[i], 1);
    if (lane == 31) excl = Map<KP>::identity();
    excl.store(buf.chunk_maps + (tile * C + lane) * MB);
    if (lane == 0) inc.store(buf.tile_maps + tile * MB);
  }
  if constexpr (kScan >= 0) {
    if (last_cta_arrives(buf.tickets + kTicketChunkMaps)) bwd_scan_cta<KP, kScan, 128>(buf, seq);
  }
}

// k_bwd_replay: thread per chunk, q_t = f_t[q_{t+1}]  (FB.hpp:140-160)
template <int KP>
__global__ void __launch_bounds__(128) k_bwd_replay(SweepBuffers buf) {
  pdl_enter();
  constexpr int L = Layout::L, C = Layout::C, MB = 8 * Map<KP>::W;
  const uint64_t B = device_nblocks(buf.nblocks, buf.capacity);
  const uint64_t nch = (B + L - 1) / L;
  for (uint64_t ch = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; ch < nch; ch += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t tile = ch / C;
    const int c = (int)(ch % C);
    const uint64_t first = ch * L;
    const int steps = (B - first) < (uint64_t)L ? (int)(B - first) : L;
    // state following this chunk = (maps of the later chunks of the tile)(state following the tile)
    uint32_t q = Map<KP>::load(buf.chunk_maps + ch * MB).get(buf.tile_qin[tile]);
    // the maps do not depend on the state being resolved: eight loads in flight, then the dependent look-ups
    for (int t0 = steps - 1; t0 >= 0; t0 -= 8) {
      Map<KP> mp[8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (t0 - i >= 0) mp[i] = Map<KP>::load(buf.maps + Layout::at(tile, c, t0 - i) * MB);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (t0 - i >= 0) {
          q = mp[i].get(q);
          buf.states[Layout::at(tile, c, t0 - i)] = (uint8_t)q;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// k_mix_sample: StateSequence/Mixture.hpp:90-112, thread per storage slot

template <int KP>
__global__ void __launch_bounds__(256) k_mix_sample(SweepBuffers buf, ModelDev<KP> m, uint64_t seed, uint64_t sweep) {
  pdl_enter();
  const uint64_t B = device_nblocks(buf.nblocks, buf.capacity);
  const uint64_t first = buf.seg.world > 1 ? seg_first_block(buf.seg) : 0;
  const uint64_t slots = (B + Layout::TB - 1) / Layout::TB * Layout::TB;
  for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < slots; p += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t b = Layout::inv(p);
    if (b >= B) continue;
    double w[KP];
#pragma unroll
    for (int s = 0; s < KP; ++s) w[s] = buf.e[p * KP + s];
    const uint64_t gb = buf.seg.world > 1 ? first + b : b;  // global block index
    const double u = buf.replay_u ? buf.replay_u[gb] : Philox::uniform(seed, sweep, 1u, gb);
    buf.states[p] = (uint8_t)discrete_draw<KP>(w, m.K, u);
  }
}

// ------------------------------------------------------------------------------------------------
// reductions (FB.hpp:170-200).  Fixed grid, fixed per-thread order, fixed trees: results do not
// depend on scheduling.

constexpr int kReduceThreads = 256;

// kFinal 0: partial sums only (k_reduce_final / k_reduce_final_exchange follow); 1: the CTA that arrives last also
// forms the final sums (same fixed order and tree as k_reduce_final_exchange); 2: ... and runs the statistics exchange
// Final sums by one CTA of NT threads (the one that arrived last): out_f64[v] = sum over the CTAs' partial rows, fixed
// assignment and fixed tree => deterministic.  Thread t adds up rows t, t + NT, ... (all 2 KP loads of a row in flight),
// then warp shuffles and one pass over the warps.
template <int KP, int NT>
__device__ __forceinline__ void reduce_final_cta(const SweepBuffers& buf, int nparts) {
  __shared__ double s_fin[NT / 32][2 * KP];
  double acc[2 * KP];
#pragma unroll
  for (int v = 0; v < 2 * KP; ++v) acc[v] = 0.0;
  for (int i = threadIdx.x; i < nparts; i += NT) {
    double row[2 * KP];
#pragma unroll
    for (int v = 0; v < 2 * KP; ++v) row[v] = __ldcg(buf.partials + (size_t)i * 2 * KP + v);
#pragma unroll
    for (int v = 0; v < 2 * KP; ++v) acc[v] += row[v];
  }
#pragma unroll
  for (int v = 0; v < 2 * KP; ++v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[v] += shfl_xor_double(acc[v], o);
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int v = 0; v < 2 * KP; ++v) s_fin[threadIdx.x >> 5][v] = acc[v];
  }
  __syncthreads();
  if (threadIdx.x < 2 * KP) {
    double t = 0.0;
    for (int wv = 0; wv < NT / 32; ++wv) t += s_fin[wv][threadIdx.x];
    buf.out_f64[threadIdx.x] = t;
  }
}

template <int KP, int kFinal>
__global__ void __launch_bounds__(kReduceThreads) k_reduce_partial(SweepBuffers buf, int K, uint32_t stats_words,
                                                                   unsigned long long seq) {
  pdl_enter();
  __shared__ unsigned long long s_trans[KP * KP];
  __shared__ unsigned long long s_n[KP];
  __shared__ double s_sum[kReduceThreads / 32][2 * KP];
  const uint64_t B = device_nblocks(buf.nblocks, buf.capacity);
  const uint64_t slots = (B + Layout::TB - 1) / Layout::TB * Layout::TB;
  for (int i = threadIdx.x; i < KP * KP; i += blockDim.x) s_trans[i] = 0;
  for (int i = threadIdx.x; i < KP; i += blockDim.x) s_n[i] = 0;
  __syncthreads();
  double ax[KP], aq[KP];
  unsigned long long an[KP], ad[KP];  // observations per state; diagonal transition counts
#pragma unroll
  for (int s = 0; s < KP; ++s) {
    ax[s] = aq[s] = 0.0;
    an[s] = ad[s] = 0;
  }
  // Transitions are counted at their source block: (q_b -> q_{b+1}) for every block that has a successor
  // anywhere in the sequence, plus the phantom 0 -> q_0 of the first block (FB.hpp:177,182-184).  In
  // segment mode the successor of a rank's last block is the state following the segment.
  const bool seg = buf.seg.world > 1;
  const bool has_after = seg && seg_later_blocks(buf.seg);
  const bool first_rank = !seg || buf.seg.rank == 0;
  for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < slots; p += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t b = Layout::inv(p);
    if (b >= B) continue;
    const uint32_t st = buf.states[p];
    const bool has_next = (b + 1 < B) || has_after;
    uint32_t next = st;
    if (b + 1 < B)
      next = buf.states[Layout::perm(b + 1)];
    else if (has_after)
      next = buf.tile_qin[(B - 1) / Layout::TB];
    const uint32_t n = buf.bN[p];
    const double2 v = buf.bS[p];
    unsigned long long dg = (unsigned long long)(n - 1) + ((has_next && next == st) ? 1ull : 0ull);
    if (b == 0 && first_rank) {
      if (st == 0u)
        dg += 1ull;
      else
        atomicAdd(&s_trans[st], 1ull);  // row 0, column st
    }
#pragma unroll
    for (int s = 0; s < KP; ++s) {
      const bool hit = st == (uint32_t)s;
      ax[s] += hit ? v.x : 0.0;
      aq[s] += hit ? v.y : 0.0;
      an[s] += hit ? (unsigned long long)n : 0ull;
      ad[s] += hit ? dg : 0ull;
    }
    if (has_next && next != st) atomicAdd(&s_trans[st * KP + next], 1ull);  // state changes are rare: low contention
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
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int s = 0; s < KP; ++s) {
      s_sum[threadIdx.x >> 5][s] = ax[s];
      s_sum[threadIdx.x >> 5][KP + s] = aq[s];
      if (an[s]) atomicAdd(&s_n[s], an[s]);
      if (ad[s]) atomicAdd(&s_trans[s * KP + s], ad[s]);
    }
  }
  __syncthreads();
  if (threadIdx.x < 2 * KP) {
    double t = 0.0;
    for (int wv = 0; wv < kReduceThreads / 32; ++wv) t += s_sum[wv][threadIdx.x];
    buf.partials[(size_t)blockIdx.x * 2 * KP + threadIdx.x] = t;
  }
  for (int i = threadIdx.x; i < KP * KP; i += blockDim.x)
    if (s_trans[i]) atomicAdd(&buf.out_u64[KP + i], s_trans[i]);
  for (int i = threadIdx.x; i < KP; i += blockDim.x)
    if (s_n[i]) atomicAdd(&buf.out_u64[i], s_n[i]);
  if constexpr (kFinal != 0) {
    if (!last_cta_arrives(buf.tickets + kTicketReduce)) return;
    reduce_final_cta<KP, kReduceThreads>(buf, (int)gridDim.x);
    if constexpr (kFinal == 2) {
      __threadfence();
      __syncthreads();
      p2p_exchange_cta(buf.seg.p2p, kSlotStats, seq, reinterpret_cast<const uint64_t*>(buf.seg.stats_send), stats_words,
                       reinterpret_cast<uint64_t*>(buf.seg.stats_recv));
    }
  }
}

// k_bwd_replay_reduce (K <= 8): k_bwd_replay and the statistics pass in one kernel, thread per quarter chunk (the chunk-map
// kernel keeps the maps of a chunk's later quarters: a latency-bound walk wants many short ones).  Walking its 8
// blocks backwards the thread knows every block's state and its successor's, so transitions are counted on the way; the
// sums go to the thread's own per-state slots in shared memory, indexed by the state (4 read-modify-writes per block;
// K-way selects into registers cost 4 K, and "flush when the state changes" does not help a warp whose 32 lanes change
// state at different blocks).  The CTA that arrives last forms the final sums (and, kFinal == 2, runs the statistics
// exchange of a split sequence).
template <int KP, int kFinal>
__global__ void __launch_bounds__(128) k_bwd_replay_reduce(SweepBuffers buf, uint32_t stats_words, unsigned long long seq) {
  pdl_enter();
  constexpr int L = Layout::L, C = Layout::C, MB = 8 * Map<KP>::W, NT = 128;
  __shared__ unsigned long long s_trans[KP * KP];
  __shared__ unsigned long long s_n[KP];
  static_assert(NT == 128, "the CTA sums add four slots per lane");
  // [value][state][thread]: a warp's 32 lanes hit 32 different banks whatever their states are
  __shared__ double s_ax[KP][NT], s_aq[KP][NT];
  __shared__ unsigned long long s_an[KP][NT], s_ad[KP][NT];
  const int tid = threadIdx.x;
  for (int i = tid; i < KP * KP; i += NT) s_trans[i] = 0;
  for (int i = tid; i < KP; i += NT) s_n[i] = 0;
#pragma unroll
  for (int s = 0; s < KP; ++s) {
    s_ax[s][tid] = s_aq[s][tid] = 0.0;
    s_an[s][tid] = s_ad[s][tid] = 0ull;
  }
  __syncthreads();
  const uint64_t B = device_nblocks(buf.nblocks, buf.capacity);
  const uint64_t nch = (B + L - 1) / L;
  const bool seg = buf.seg.world > 1;
  const bool has_after = seg && seg_later_blocks(buf.seg);
  const bool first_rank = !seg || buf.seg.rank == 0;
  // thread index = ((tile, quarter), chunk): the 32 lanes of a warp take the same quarter of the 32 chunks of a tile
  const uint64_t ntiles = (B + Layout::TB - 1) / Layout::TB;
  const uint64_t nthreads = ntiles * 4 * C;
  for (uint64_t g = (uint64_t)blockIdx.x * NT + tid; g < nthreads; g += (uint64_t)gridDim.x * NT) {
    const int c = (int)(g % C);
    const uint64_t unit = g / C, tile = unit / 4;
    const int q4 = (int)(unit % 4), t_lo = 8 * q4;
    const uint64_t ch = tile * C + c;
    const uint64_t first = ch * L + t_lo;  // first block of the quarter
    if (first >= B) continue;
    const int steps = (B - first) < 8ull ? (int)(B - first) : 8;
    // state following the chunk = (maps of the later chunks of the tile)(state following the tile); the state following
    // the quarter = (map of the chunk's later quarters)(state following the chunk)
    uint32_t next = Map<KP>::load(buf.chunk_maps + ch * MB).get(buf.tile_qin[tile]);
    if (q4 < 3) next = Map<KP>::load(buf.chunk_submaps + (ch * 3 + q4) * MB).get(next);
    bool has_next = (first + steps < B) || has_after;  // does the quarter's last block have a successor anywhere?
    {
      const int t0 = steps - 1;
      Map<KP> mp[8];
      uint32_t nn[8];
      double2 vv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (t0 - i >= 0) {
          const uint64_t p = Layout::at(tile, c, t_lo + t0 - i);
          mp[i] = Map<KP>::load(buf.maps + p * MB);
          nn[i] = buf.bN[p];
          vv[i] = buf.bS[p];
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (t0 - i >= 0) {
          const uint32_t q = mp[i].get(next);
          buf.states[Layout::at(tile, c, t_lo + t0 - i)] = (uint8_t)q;
          // FB.hpp:177,182-184: N - 1 self transitions inside the block, one transition to the successor
          unsigned long long dg = (unsigned long long)(nn[i] - 1u);
          if (has_next) {
            if (next == q)
              dg += 1ull;
            else
              atomicAdd(&s_trans[q * KP + next], 1ull);  // state changes are rare: low contention
          }
          if (first + (uint64_t)(t0 - i) == 0 && first_rank) {  // the phantom 0 -> q_0 in front of the sequence
            if (q == 0u)
              dg += 1ull;
            else
              atomicAdd(&s_trans[q], 1ull);  // row 0, column q
          }
          s_ax[q][tid] += vv[i].x;
          s_aq[q][tid] += vv[i].y;
          s_an[q][tid] += nn[i];
          s_ad[q][tid] += dg;
          next = q;
          has_next = true;
        }
      }
    }
  }
  // sums over the CTA's threads: warp w takes the arrays w, w + 4, ... of the 4 K (value, state) arrays — each lane adds
  // the four slots tid, tid + 32, ... and five shuffle steps finish it (fixed order: deterministic); a thread reducing
  // its own 4 K slots by shuffles would spend more instructions here than in its eight blocks
  __syncthreads();
  {
    const int lane = tid & 31, warp = tid >> 5;
    for (int a = warp; a < 4 * KP; a += NT / 32) {
      const int kind = a / KP, st = a % KP;  // 0: sum x, 1: sum x^2, 2: observations, 3: diagonal transitions
      if (kind < 2) {
        const double* src = kind == 0 ? s_ax[st] : s_aq[st];
        double t = (src[lane] + src[lane + 32]) + (src[lane + 64] + src[lane + 96]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += shfl_xor_double(t, o);
        if (lane == 0) buf.partials[(size_t)blockIdx.x * 2 * KP + kind * KP + st] = t;
      } else {
        const unsigned long long* src = kind == 2 ? s_an[st] : s_ad[st];
        unsigned long long t = (src[lane] + src[lane + 32]) + (src[lane + 64] + src[lane + 96]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0 && t) {
          if (kind == 2)
            atomicAdd(&s_n[st], t);
          else
            atomicAdd(&s_trans[st * KP + st], t);
        }
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < KP * KP; i += NT)
    if (s_trans[i]) atomicAdd(&buf.out_u64[KP + i], s_trans[i]);
  for (int i = tid; i < KP; i += NT)
    if (s_n[i]) atomicAdd(&buf.out_u64[i], s_n[i]);
  if (!last_cta_arrives(buf.tickets + kTicketReduce)) return;
  reduce_final_cta<KP, NT>(buf, (int)gridDim.x);
  __threadfence();
  __syncthreads();
  if constexpr (kFinal == 2)
    p2p_exchange_cta(buf.seg.p2p, kSlotStats, seq, reinterpret_cast<const uint64_t*>(buf.seg.stats_send), stats_words,
                     reinterpret_cast<uint64_t*>(buf.seg.stats_recv));
  // the result block (of every rank after the exchange) straight into the host's pinned mirror: the kernel's end is then
  // all the host waits for, no copy behind it
  if (buf.result_host != nullptr) {
    const unsigned long long* src = kFinal == 2 ? buf.seg.stats_recv : buf.nblocks;
    const uint32_t n = kFinal == 2 ? stats_words * (uint32_t)buf.seg.world : buf.result_words;
    for (uint32_t i = tid; i < n; i += NT) buf.result_host[i] = __ldcg(src + i);
  }
}

// one CTA per output value; fixed assignment and fixed tree => deterministic
template <int KP>
__global__ void __launch_bounds__(128) k_reduce_final(SweepBuffers buf, int nparts) {
  pdl_enter();
  __shared__ double sh[4];
  const int v = blockIdx.x;  // 0 .. 2*KP-1
  double t = 0.0;
  for (int i = threadIdx.x; i < nparts; i += 128) t += buf.partials[(size_t)i * 2 * KP + v];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += shfl_xor_double(t, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) buf.out_f64[v] = (sh[0] + sh[1]) + (sh[2] + sh[3]);
}

// Multivariate data: per-state (sum x, sum x^2) of one further dimension (plane `bS`), same fixed trees as
// k_reduce_partial; counts and transitions do not depend on the dimension and are not recomputed.
template <int KP>
__global__ void __launch_bounds__(kReduceThreads) k_reduce_dim(SweepBuffers buf, const double2* __restrict__ bS,
                                                               double* __restrict__ partials) {
  pdl_enter();
  __shared__ double s_sum[kReduceThreads / 32][2 * KP];
  const uint64_t B = device_nblocks(buf.nblocks, buf.capacity);
  const uint64_t slots = (B + Layout::TB - 1) / Layout::TB * Layout::TB;
  double ax[KP], aq[KP];
#pragma unroll
  for (int s = 0; s < KP; ++s) ax[s] = aq[s] = 0.0;
  for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < slots; p += (uint64_t)gridDim.x * blockDim.x) {
    if (Layout::inv(p) >= B) continue;
    const uint32_t st = buf.states[p];
    const double2 v = bS[p];
#pragma unroll
    for (int s = 0; s < KP; ++s) {
      const bool hit = st == (uint32_t)s;
      ax[s] += hit ? v.x : 0.0;
      aq[s] += hit ? v.y : 0.0;
    }
  }
#pragma unroll
  for (int s = 0; s < KP; ++s) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ax[s] += shfl_xor_double(ax[s], o);
      aq[s] += shfl_xor_double(aq[s], o);
    }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int s = 0; s < KP; ++s) {
      s_sum[threadIdx.x >> 5][s] = ax[s];
      s_sum[threadIdx.x >> 5][KP + s] = aq[s];
    }
  }
  __syncthreads();
  if (threadIdx.x < 2 * KP) {
    double t = 0.0;
    for (int wv = 0; wv < kReduceThreads / 32; ++wv) t += s_sum[wv][threadIdx.x];
    partials[(size_t)blockIdx.x * 2 * KP + threadIdx.x] = t;
  }
}

// Segment mode with peer mailboxes: the final sums and the statistics exchange in one single-CTA kernel.  Warp
// w sums output values w, w + 8, ... over the partials (fixed order, fixed tree => deterministic).
template <int KP>
__global__ void __launch_bounds__(256) k_reduce_final_exchange(SweepBuffers buf, int nparts, uint32_t stats_words,
                                                               unsigned long long seq) {
  pdl_enter();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int v = warp; v < 2 * KP; v += 8) {
    double t = 0.0;
    for (int i = lane; i < nparts; i += 32) t += buf.partials[(size_t)i * 2 * KP + v];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += shfl_xor_double(t, o);
    if (lane == 0) buf.out_f64[v] = t;
  }
  __threadfence();
  __syncthreads();
  p2p_exchange_cta(buf.seg.p2p, kSlotStats, seq, reinterpret_cast<const uint64_t*>(buf.seg.stats_send), stats_words,
                   reinterpret_cast<uint64_t*>(buf.seg.stats_recv));
}

template <int KP>
__global__ void k_sum_partials(const double* partials, int n, double* out) {
  pdl_enter();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < n; ++i) t += partials[i];
    *out = t;
  }
}

template <int KP>
__global__ void k_clear_out(SweepBuffers buf) {
  const int n = KP + KP * KP + 2;
  for (int i = threadIdx.x; i < n; i += blockDim.x) buf.out_u64[i] = 0;
  for (int i = threadIdx.x; i < 2 * KP + 1; i += blockDim.x) buf.out_f64[i] = 0.0;
}

// ------------------------------------------------------------------------------------------------
// Tile scan for K > 8 on a single handle, spread over the device (the single-CTA k_fwd_tilescan spends 15 ms on
// 5 000 tiles at K = 20: a thousand threads at 64 registers walking 3 kB operators).  Same three steps, one launch
// each; the number of tiles is only known on the device, so the grid is sized from the host's hint and every CTA
// derives the same split of the tiles into G groups of S:
//   k_tilescan_groups  CTA g, thread = operator row: group operator = ordered product of the group's tile operators
//   k_tilescan_top     one warp: forward vector entering every group (G vector-operator products)
//   k_tilescan_apply   CTA g, one warp: forward vector entering every tile of the group
constexpr int kScanGroupsMax = 256;  // group_ops / group_exp / group_ain hold this many entries

__device__ __forceinline__ void scan_split(int nt, int grid, int& G, int& S) {
  const int gmax = grid < kScanGroupsMax ? grid : kScanGroupsMax;
  S = nt > 0 ? (nt + gmax - 1) / gmax : 1;
  G = nt > 0 ? (nt + S - 1) / S : 0;
}

template <int KP>
__global__ void __launch_bounds__(((KP + 31) / 32) * 32) k_tilescan_groups(SweepBuffers buf) {
  pdl_enter();
  const uint64_t B = device_nblocks(buf.nblocks, buf.capacity);
  const int nt = (int)((B + Layout::TB - 1) / Layout::TB);
  int G, S;
  scan_split(nt, (int)gridDim.x, G, S);
  const int g = blockIdx.x, i = threadIdx.x;
  if (g >= G || i >= KP) return;
  double r[KP];
  int rex = 0;
#pragma unroll
  for (int j = 0; j < KP; ++j) r[j] = (j == i) ? 1.0 : 0.0;
  const int t0 = g * S, t1 = min(nt, (g + 1) * S);
#pragma unroll 1
  for (int t = t0; t < t1; ++t)
    row_times_op<KP, false>(r, rex, buf.tile_ops + (uint64_t)t * KP * KP, buf.tile_exp + (uint64_t)t * KP);
#pragma unroll
  for (int j = 0; j < KP; ++j) buf.group_ops[((size_t)g * KP + i) * KP + j] = r[j];
  buf.group_exp[g * KP + i] = rex;
}

template <int KP>
__global__ void __launch_bounds__(32) k_tilescan_top(SweepBuffers buf, ModelDev<KP> m, int grid) {
  pdl_enter();
  const uint64_t B = device_nblocks(buf.nblocks, buf.capacity);
  const int nt = (int)((B + Layout::TB - 1) / Layout::TB);
  int G, S;
  scan_split(nt, grid, G, S);
  if (G == 0) return;
  const int lane = threadIdx.x;
  double a = lane < KP ? m.pi[lane] : 0.0;  // row 0 of the trellis is pi itself (FB.hpp:57)
  LaneOp<KP> cur, nxt;
  load_lane_op<KP>(cur, buf.group_ops, buf.group_exp, lane);
#pragma unroll 1
  for (int g = 0; g < G; ++g) {
    if (lane < KP) buf.group_ain[g * KP + lane] = a;
    const int gn = (g + 1 < G) ? g + 1 : g;
    load_lane_op<KP>(nxt, buf.group_ops + (size_t)gn * KP * KP, buf.group_exp + gn * KP, lane);
    if (g + 1 < G && !warp_apply_op<KP>(a, cur) && lane == 0) atomicAdd(&buf.out_u64[KP + KP * KP], 1ull);
    cur = nxt;
  }
}

template <int KP>
__global__ void __launch_bounds__(32) k_tilescan_apply(SweepBuffers buf) {
  pdl_enter();
  const uint64_t B = device_nblocks(buf.nblocks, buf.capacity);
  const int nt = (int)((B + Layout::TB - 1) / Layout::TB);
  int G, S;
  scan_split(nt, (int)gridDim.x, G, S);
  const int g = blockIdx.x, lane = threadIdx.x;
  if (g >= G) return;
  double a = lane < KP ? buf.group_ain[g * KP + lane] : 0.0;
  const int t0 = g * S, t1 = min(nt, (g + 1) * S);
  LaneOp<KP> cur, nxt;
  load_lane_op<KP>(cur, buf.tile_ops + (uint64_t)t0 * KP * KP, buf.tile_exp + (uint64_t)t0 * KP, lane);
#pragma unroll 1
  for (int t = t0; t < t1; ++t) {
    if (lane < KP) buf.tile_ain[(uint64_t)t * KP + lane] = a;
    const int tn = (t + 1 < t1) ? t + 1 : t;
    load_lane_op<KP>(nxt, buf.tile_ops + (uint64_t)tn * KP * KP, buf.tile_exp + (uint64_t)tn * KP, lane);
    if (t + 1 < t1 && !warp_apply_op<KP>(a, cur) && lane == 0) atomicAdd(&buf.out_u64[KP + KP * KP], 1ull);
    cur = nxt;
  }
}

// ------------------------------------------------------------------------------------------------
// host-side orchestration (one instantiation per padded state count)

inline int grid_for(uint64_t items, int threads, int sms, int per_sm) {
  uint64_t g = (items + threads - 1) / threads;
  const uint64_t cap = (uint64_t)sms * per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// Mixture sampler in segment mode: the "map" a rank publishes is the constant map onto the state of its
// first block (identity without blocks), so the same resolution over later ranks yields the state of the
// next block of the sequence.
template <int KP>
__global__ void k_mix_segmap(SweepBuffers buf) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const uint64_t B = device_nblocks(buf.nblocks, buf.capacity);
  Map<KP> mp = Map<KP>::identity();
  if (B > 0) {
    const uint32_t q = buf.states[Layout::perm(0)];
    mp = Map<KP>::zero();
#pragma unroll
    for (int j = 0; j < KP; ++j) mp.set(j, q);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) buf.seg.send_map[i] = i < Map<KP>::W ? mp.w[i] : 0ull;
}
template <int KP>
__global__ void k_mix_qend(SweepBuffers buf) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const uint64_t B = device_nblocks(buf.nblocks, buf.capacity);
  if (B == 0) return;
  uint32_t q_end = 0;
  for (int r = buf.seg.world - 1; r > buf.seg.rank; --r) {
    Map<KP> o;
#pragma unroll
    for (int i = 0; i < Map<KP>::W; ++i) o.w[i] = buf.seg.maps[4 * r + i];
    q_end = o.get(q_end);
  }
  buf.tile_qin[(B - 1) / Layout::TB] = (uint8_t)q_end;
}

// phase 0: single handle; 1 / 2: before / after the all-gather of the segment operators.  Returns the number of
// launches.  K <= 8: the cluster kernel handles sweeps of up to kMaxTiles tiles; if the block arrays could hold
// more than that, the single-CTA kernel is launched too and takes over exactly when the cluster kernel stood down
// (the block count is only known on the device).
template <int KP, int kPhase>
int launch_fwd_tilescan_phase(const SweepBuffers& b, const ModelDev<KP>& m, uint64_t ntiles_hint, unsigned long long seq,
                              cudaStream_t s) {
  if constexpr (KP <= 8) {
    using Cfg = ClusterScanCfg<KP>;
    // per device, so set on every launch (a host-side table lookup)
    cudaFuncSetAttribute(k_fwd_tilescan_cluster<KP, kPhase>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmem);
    launch_k(k_fwd_tilescan_cluster<KP, kPhase>, kClusterCtas, 256, Cfg::kSmem, s, b, m, seq);
    if (ntiles_hint <= (uint64_t)Cfg::kMaxTiles) return 1;
    launch_k(k_fwd_tilescan_small<KP, kPhase>, 1, 256, 0, s, b, m, (int)Cfg::kMaxTiles, seq);
    return 2;
  } else {
    static_assert(KP <= 8 || kPhase != 3, "the embedded exchange exists for K <= 8");
    if constexpr (kPhase == 0) {
      // groups of tiles: about one per SM, never more than the tiles there can be
      int grid = ntiles_hint < (uint64_t)kScanGroupsMax ? (int)ntiles_hint : kScanGroupsMax;
      if (grid > 148) grid = 148;
      if (grid < 1) grid = 1;
      launch_k(k_tilescan_groups<KP>, grid, ((KP + 31) / 32) * 32, 0, s, b);
      launch_k(k_tilescan_top<KP>, 1, 32, 0, s, b, m, grid);
      launch_k(k_tilescan_apply<KP>, grid, 32, 0, s, b);
      return 3;
    } else {
      k_fwd_tilescan<KP, kPhase><<<1, 1024, 0, s>>>(b, m);
      return 1;
    }
  }
}
// phase 3 (K <= 8, peer mailboxes): phases 1 and 2 with the operator exchange between them, in one kernel
template <int KP>
int launch_fwd_tilescan(const SweepBuffers& b, const ModelDev<KP>& m, int phase, uint64_t ntiles_hint,
                        unsigned long long seq, cudaStream_t s) {
  if (phase == 0) return launch_fwd_tilescan_phase<KP, 0>(b, m, ntiles_hint, 0, s);
  if (phase == 1) return launch_fwd_tilescan_phase<KP, 1>(b, m, ntiles_hint, 0, s);
  if (phase == 2) return launch_fwd_tilescan_phase<KP, 2>(b, m, ntiles_hint, 0, s);
  if constexpr (KP <= 8) return launch_fwd_tilescan_phase<KP, 3>(b, m, ntiles_hint, seq, s);
  return 0;
}

// maps -> chunk/tile maps -> suffix scan over tiles -> states; returns the number of launches
template <int KP>
int launch_backward(const SweepBuffers& b, const ModelDev<KP>& m, const SweepLaunch& l, bool rows, uint64_t nb,
                    cudaStream_t s, stage_cb_t cb, void* user, bool have_maps = false, bool* reduce_too = nullptr) {
  const uint64_t ntiles = (nb + Layout::TB - 1) / Layout::TB;
  int launches = 2;
  bool scanned = false;
  if (!have_maps) {  // k_replay_maps already wrote the block, chunk and tile maps
    launches += 2;
    if (cb) cb(user, "bwd_maps");
    const int gm = grid_for(ntiles * Layout::TB, 256, l.sms, 32);
    if (rows)
      launch_k(k_bwd_maps<KP, true>, gm, 256, 0, s, b, m, l.seed, l.sweep);
    else
      launch_k(k_bwd_maps<KP, false>, gm, 256, 0, s, b, m, l.seed, l.sweep);
    if (cb) cb(user, "bwd_chunkmaps");
    const int gc = grid_for(ntiles * 32, 128, l.sms, 16);
    if (b.tickets != nullptr && !(b.seg.world > 1 && b.seg.p2p == nullptr)) {
      // the CTA that arrives last scans the tile maps (and runs the map exchange of a split sequence) in the same launch
      if (b.seg.world > 1)
        launch_k(k_bwd_chunkmaps<KP, 2>, gc, 128, 0, s, b, l.next_seq(l.exchange_user, kExchangeMaps));
      else
        launch_k(k_bwd_chunkmaps<KP, 0>, gc, 128, 0, s, b, 0ull);
      scanned = true;
      --launches;
    } else {
      launch_k(k_bwd_chunkmaps<KP, -1>, gc, 128, 0, s, b, 0ull);
    }
  }
  if (!scanned) {
    if (cb) cb(user, "bwd_scan");
    if (b.seg.world > 1 && b.seg.p2p != nullptr) {
      launch_k(k_bwd_scan<KP, 2>, 1, 1024, 0, s, b, l.next_seq(l.exchange_user, kExchangeMaps));
    } else {
      if (b.seg.world > 1) {
        launch_k(k_bwd_scan<KP, 1>, 1, 1024, 0, s, b, 0ull);
        ++launches;
        if (cb) cb(user, "exchange_maps");
        if (l.exchange(l.exchange_user, kExchangeMaps) != 0) return -1;
        if (cb) cb(user, "bwd_scan2");
      }
      launch_k(k_bwd_scan<KP, 0>, 1, 1024, 0, s, b, 0ull);
    }
  }
  if (cb) cb(user, "bwd_replay");
  const int gp = grid_for((nb + Layout::L - 1) / Layout::L, 128, l.sms, 16);
  if constexpr (KP <= 8) {
    if (reduce_too != nullptr && b.tickets != nullptr) {  // states and statistics in one kernel, thread per quarter chunk
      const int gq = grid_for(ntiles * Layout::C * 4, 128, l.sms, 8);
      // the host's mirror of the result is written by the kernel when nothing adds to the result afterwards: one data
      // dimension; on a split sequence only if the kernel runs the statistics exchange itself
      const bool in_kernel_exchange = b.seg.world > 1 && b.seg.p2p != nullptr && l.stats_words;
      SweepBuffers bb = b;
      const bool to_host = b.result_host != nullptr && b.D == 1 && (b.seg.world <= 1 || in_kernel_exchange);
      if (!to_host) bb.result_host = nullptr;
      if (in_kernel_exchange)
        launch_k(k_bwd_replay_reduce<KP, 2>, gq, 128, 0, s, bb, l.stats_words, l.next_seq(l.exchange_user, kExchangeStats));
      else
        launch_k(k_bwd_replay_reduce<KP, 1>, gq, 128, 0, s, bb, (uint32_t)0, 0ull);
      if (to_host && l.result_on_host) *l.result_on_host = true;
      *reduce_too = true;
      return launches;
    }
  }
  launch_k(k_bwd_replay<KP>, gp, 128, 0, s, b);
  return launches;
}

template <int KP>
int launch_reduce(const SweepBuffers& b, int K, uint64_t nb, const SweepLaunch& l, cudaStream_t s, bool first_done = false) {
  const uint64_t ntiles = (nb + Layout::TB - 1) / Layout::TB;
  const int g = grid_for(ntiles * Layout::TB, kReduceThreads, l.sms, 4);
  int launches = first_done ? 0 : 1;
  if (first_done) {
    // k_bwd_replay_reduce left the counts and the sums of the first dimension
  } else if (b.D == 1 && b.tickets != nullptr) {  // the last CTA finishes: final sums (+ statistics exchange) in the same launch
    if (b.seg.world > 1 && b.seg.p2p != nullptr && l.stats_words)
      launch_k(k_reduce_partial<KP, 2>, g, kReduceThreads, 0, s, b, K, l.stats_words, l.next_seq(l.exchange_user, kExchangeStats));
    else
      launch_k(k_reduce_partial<KP, 1>, g, kReduceThreads, 0, s, b, K, (uint32_t)0, 0ull);
  } else {
    launch_k(k_reduce_partial<KP, 0>, g, kReduceThreads, 0, s, b, K, (uint32_t)0, 0ull);
    if (b.seg.world > 1 && b.seg.p2p != nullptr && l.stats_words)
      launch_k(k_reduce_final_exchange<KP>, 1, 256, 0, s, b, g, l.stats_words, l.next_seq(l.exchange_user, kExchangeStats));
    else
      launch_k(k_reduce_final<KP>, 2 * KP, 128, 0, s, b, g);
    ++launches;
  }
  for (int d = 1; d < b.D; ++d) {  // multivariate data: the remaining dimensions, written behind the log-likelihood
    SweepBuffers bd = b;
    bd.out_f64 = b.out_f64 + 2 * KP + 1 + (size_t)(d - 1) * 2 * KP;
    launch_k(k_reduce_dim<KP>, g, kReduceThreads, 0, s, b, (const double2*)(b.bS + (size_t)d * b.capacity), b.partials);
    launch_k(k_reduce_final<KP>, 2 * KP, 128, 0, s, bd, g);
    launches += 2;
  }
  return launches;
}

template <int KP>
int sweep_impl(const ModelHost& mh, const SweepBuffers& b, const SweepLaunch& l, cudaStream_t s, stage_cb_t cb,
               void* user) {
  const ModelDev<KP> m = make_model<KP>(mh);
  const bool loglik = (l.flags & HML_SWEEP_LOGLIK) != 0;
  const bool rows = (l.flags & HML_SWEEP_KEEP_ROWS) != 0 && b.rows != nullptr;
  const bool seg = b.seg.world > 1;
  const uint64_t nb = l.nblocks_hint;
  const uint64_t ntiles = (nb + Layout::TB - 1) / Layout::TB;
  int launches = 0;
  auto stage = [&](const char* name) {
    if (cb) cb(user, name);
  };
  constexpr bool kPrefix = KP <= 8;  // k_fwd_chunks_prefix + k_fwd_replay_prefix
  bool reduced = false;              // the backward replay kernel also did the statistics pass
  stage("block_emit");
  if (b.D > 1) {
    const EmitMD<KP> md = make_emit_md<KP>(mh);
    const int g = grid_for(ntiles * Layout::TB, 256, l.sms, 32);
    if (l.mixture) {
      if (l.gather)
        launch_k(k_block_emit_md<KP, true, true, true>, g, 256, 0, s, b, md, (int)0);
      else
        launch_k(k_block_emit_md<KP, false, true, true>, g, 256, 0, s, b, md, (int)0);
    } else {
      if (l.gather)
        launch_k(k_block_emit_md<KP, true, true, false>, g, 256, 0, s, b, md, (int)loglik);
      else
        launch_k(k_block_emit_md<KP, false, true, false>, g, 256, 0, s, b, md, (int)loglik);
    }
    ++launches;
  } else {
    // K <= 8 with fresh block sums: CTAs take groups of 256 consecutive blocks in a loop, 4 resident CTAs per SM (60 registers)
    const int g = (l.gather && KP <= 8) ? grid_for(ntiles * Layout::TB, 256, l.sms, 4) : grid_for(ntiles * Layout::TB, 256, l.sms, 32);
    if (l.mixture) {
      if (l.gather)
        launch_k(k_block_emit<KP, true, true, true>, g, 256, 0, s, b, m, (int)0);
      else
        launch_k(k_block_emit<KP, false, true, true>, g, 256, 0, s, b, m, (int)0);
    } else {
      if (l.gather)
        launch_k(k_block_emit<KP, true, true, false>, g, 256, 0, s, b, m, (int)loglik);
      else
        launch_k(k_block_emit<KP, false, true, false>, g, 256, 0, s, b, m, (int)loglik);
    }
    ++launches;
  }
  if (l.mixture) {
    stage("mix_sample");
    launch_k(k_mix_sample<KP>, grid_for(ntiles * Layout::TB, 256, l.sms, 32), 256, 0, s, b, m, l.seed, l.sweep);
    ++launches;
    if (seg) {
      k_mix_segmap<KP><<<1, 32, 0, s>>>(b);
      stage("exchange_maps");
      if (l.exchange(l.exchange_user, kExchangeMaps) != 0) return -1;
      k_mix_qend<KP><<<1, 32, 0, s>>>(b);
      launches += 2;
    }
  } else if (l.speculate && !(seg && loglik)) {
    // speculative filter: vector recursions from guessed starts, then the repair pass (two kernels, no operators)
    double* const lognorm = reinterpret_cast<double*>(b.maps);  // per-block scratch until k_bwd_maps writes the maps
    stage("fwd_spec");
    const int gr = grid_for(ntiles, 1, l.sms, 32);
    const int sub = kPrefix ? l.spec_sub : Layout::L;                    // piece length of the speculative pass
    const int gs = grid_for(ntiles * (Layout::L / sub), 1, l.sms, 32);  // one warp per (tile, piece)
    if constexpr (kPrefix) {
      using RCfg = ReplayCfg<KP>;
      if (loglik)
        launch_k(k_fwd_replay_prefix<KP, true, true, true>, gs, 32, RCfg::kSmem, s, b, m, lognorm, (int)l.spec_warm, sub);
      else if (rows)
        launch_k(k_fwd_replay_prefix<KP, true, false, true>, gs, 32, RCfg::kSmem, s, b, m, lognorm, (int)l.spec_warm, sub);
      else
        launch_k(k_fwd_replay_prefix<KP, false, false, true>, gs, 32, RCfg::kSmem, s, b, m, lognorm, (int)l.spec_warm, sub);
    } else {
      if (loglik)
        launch_k(k_fwd_replay<KP, true, true>, gr, 32, 0, s, b, m, lognorm, (int)l.spec_warm);
      else
        launch_k(k_fwd_replay<KP, false, true>, gr, 32, 0, s, b, m, lognorm, (int)l.spec_warm);
    }
    ++launches;
    stage("fwd_fixup");
    const int gf = grid_for(ntiles * Layout::C * (Layout::L / sub), 128, l.sms, 8);
    bool head_done = false;
    if (loglik) {
      launch_k(k_fwd_fixup<KP, true, true>, gf, 128, 0, s, b, m, lognorm, 0ull, sub);
      launch_k(k_sum_partials<KP>, 1, 32, 0, s, b.partials, gf, b.out_f64 + 2 * KP);
      ++launches;
    } else if (seg && b.seg.p2p != nullptr && b.tickets != nullptr) {
      // split sequence: the repair of the rank's first chunk (one all-gather of K + 1 words) rides in the last CTA
      launch_k(k_fwd_fixup<KP, !kPrefix, false, true>, gf, 128, 0, s, b, m, lognorm, l.next_seq(l.exchange_user, kExchangeOps), sub);
      head_done = true;
    } else if (rows || !kPrefix) {
      launch_k(k_fwd_fixup<KP, true, false>, gf, 128, 0, s, b, m, lognorm, 0ull, sub);
    } else {
      launch_k(k_fwd_fixup<KP, false, false>, gf, 128, 0, s, b, m, lognorm, 0ull, sub);
    }
    ++launches;
    if (seg && !head_done) {  // chunk 0 of the later ranks, from the last row of the rank before
      stage("fwd_fixup_head");
      constexpr bool kEx = !kPrefix;
      if (b.seg.p2p != nullptr) {
        launch_k(k_fwd_fixup_head<KP, kEx>, 1, 256, 0, s, b, m, (int)3, (int)(KP + 1), l.next_seq(l.exchange_user, kExchangeOps), sub);
        ++launches;
      } else {
        launch_k(k_fwd_fixup_head<KP, kEx>, 1, 256, 0, s, b, m, (int)1, (int)(KP * KP + KP), 0ull, sub);
        stage("exchange_ops");
        if (l.exchange(l.exchange_user, kExchangeOps) != 0) return -1;
        stage("fwd_fixup_head2");
        launch_k(k_fwd_fixup_head<KP, kEx>, 1, 256, 0, s, b, m, (int)2, (int)(KP * KP + KP), 0ull, sub);
        launches += 2;
      }
    }
    const int nbw = launch_backward<KP>(b, m, l, rows, nb, s, cb, user, false, &reduced);
    if (nbw < 0) return -1;
    launches += nbw;
  } else {
    stage("fwd_chunks");
    if constexpr (kPrefix) {
      launch_k(k_fwd_chunks_prefix<KP>, grid_for(ntiles, 1, l.sms, 16), FwdCfg<KP>::THREADS, 0, s, b, m);
    } else if (b.wide_ops != nullptr) {
      // at most kWideCtasPerSm CTAs per SM: that many scratch areas exist (alloc_blocks)
      cudaFuncSetAttribute(k_fwd_chunks_wide<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WideCfg<KP>::kSmem);
      launch_k(k_fwd_chunks_wide<KP>, grid_for(ntiles, 1, l.sms, kWideCtasPerSm), WideCfg<KP>::THREADS, WideCfg<KP>::kSmem,
               s, b, m, b.wide_ops, b.wide_exp);
    } else {
      k_fwd_chunks<KP><<<grid_for(ntiles, 1, l.sms, 16), FwdCfg<KP>::THREADS, 0, s>>>(b, m);
    }
    ++launches;
    stage("fwd_tilescan");
    const bool embed = seg && b.seg.p2p != nullptr && KP <= 8;  // collectives inside the producing kernels
    if (embed) {
      launches += launch_fwd_tilescan<KP>(b, m, 3, ntiles, l.next_seq(l.exchange_user, kExchangeOps), s);
    } else if (seg) {
      launches += launch_fwd_tilescan<KP>(b, m, 1, ntiles, 0, s);
      stage("exchange_ops");
      if (l.exchange(l.exchange_user, kExchangeOps) != 0) return -1;
      stage("fwd_tilescan2");
      launches += launch_fwd_tilescan<KP>(b, m, 2, ntiles, 0, s);
    } else {
      launches += launch_fwd_tilescan<KP>(b, m, 0, ntiles, 0, s);
    }
    stage("fwd_replay");
    const int gr = grid_for(ntiles, 1, l.sms, 32);
    if constexpr (kPrefix) {
      using RCfg = ReplayCfg<KP>;
      if (loglik) {
        launch_k(k_fwd_replay_prefix<KP, true, true>, gr, 32, RCfg::kSmem, s, b, m, (double*)nullptr, (int)0, (int)Layout::L);
        launch_k(k_sum_partials<KP>, 1, 32, 0, s, b.partials, gr, b.out_f64 + 2 * KP);
        ++launches;
      } else if (rows) {
        launch_k(k_fwd_replay_prefix<KP, true, false>, gr, 32, RCfg::kSmem, s, b, m, (double*)nullptr, (int)0, (int)Layout::L);
      } else {
        launch_k(k_fwd_replay_prefix<KP, false, false>, gr, 32, RCfg::kSmem, s, b, m, (double*)nullptr, (int)0, (int)Layout::L);
      }
    } else {
      if (loglik) {
        launch_k(k_fwd_replay<KP, true>, gr, 32, 0, s, b, m, (double*)nullptr, (int)0);
        launch_k(k_sum_partials<KP>, 1, 32, 0, s, b.partials, gr, b.out_f64 + 2 * KP);
        ++launches;
      } else {
        launch_k(k_fwd_replay<KP, false>, gr, 32, 0, s, b, m, (double*)nullptr, (int)0);
      }
    }
    ++launches;
    const int nbw = launch_backward<KP>(b, m, l, rows, nb, s, cb, user, false, &reduced);
    if (nbw < 0) return -1;
    launches += nbw;
  }
  if (!reduced) stage("reduce");
  launches += launch_reduce<KP>(b, mh.K, nb, l, s, reduced);
  stage("end");
  return launches;
}

// Exact sequential forward pass.  In segment mode the ranks take turns: rank r starts from the final
// vector of rank r-1, which travels through the operator slots of the all-gather.
template <int KP>
int sequential_impl(const ModelHost& mh, const SweepBuffers& b, const SweepLaunch& l, cudaStream_t s) {
  const ModelDev<KP> m = make_model<KP>(mh);
  const bool loglik = (l.flags & HML_SWEEP_LOGLIK) != 0;
  const bool rows = (l.flags & HML_SWEEP_KEEP_ROWS) != 0 && b.rows != nullptr;
  const uint64_t nb = l.nblocks_hint;
  int launches = 1;
  k_clear_out<KP><<<1, 256, 0, s>>>(b);
  const int world = b.seg.world > 1 ? b.seg.world : 1;
  for (int turn = 0; turn < world; ++turn) {
    if (turn == (world > 1 ? b.seg.rank : 0)) {
      const double* ain = turn == 0 ? nullptr : b.seg.ops + (size_t)(turn - 1) * (KP * KP + KP);
      if (loglik) {
        k_fwd_sequential<KP, true><<<1, 32, 0, s>>>(b, m, ain);
        launch_k(k_sum_partials<KP>, 1, 32, 0, s, b.partials, 1, b.out_f64 + 2 * KP);
        launches += 2;
      } else {
        k_fwd_sequential<KP, false><<<1, 32, 0, s>>>(b, m, ain);
        launches += 1;
      }
    }
    if (world > 1 && l.exchange(l.exchange_user, kExchangeOps) != 0) return -1;
  }
  const int nbw = launch_backward<KP>(b, m, l, rows, nb, s, nullptr, nullptr);
  if (nbw < 0) return -1;
  launches += nbw;
  launches += launch_reduce<KP>(b, mh.K, nb, l, s);
  return launches;
}

}  // namespace hml
