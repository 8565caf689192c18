// hammlet_b200 — block-level kernels of one Gibbs sweep (sm_100a).
//
// Reference path: StateSequence<ForwardBackward>::sample, StateSequence/ForwardBackward.hpp:16-213
// (forward filter :64-125, backward sampling :130-162 with Trellis.hpp:61-66, statistics pass
// :170-212) and StateSequence<Mixture>::sample, StateSequence/Mixture.hpp:31-144.  The reference
// walks the blocks sequentially three times; here every stage is parallel over blocks:
//
//   k_block_emit   per block: (N, sum x, sum x^2) gathered from the integral arrays
//                  (Statistics/IntegralArray.hpp:104-124), emission terms e_s = exp(E_s - max E)
//                  (EFD.hpp:23-38, FB.hpp:74-84) and the self-transition rescale A_ss^(N-1) (FB.hpp:115-119)
//   forward filter, speculative (default): k_fwd_replay_prefix / k_fwd_replay with kSpec run the reference's vector
//                  recursion per piece of 8-32 blocks from a guessed start, k_fwd_fixup repairs the heads of the pieces
//                  from the true row of the block before and reports pieces that did not meet their guess (the host
//                  then repeats the sweep through the operator scan); on a split sequence its last CTA repairs the
//                  rank's first piece from the previous rank's last row
//   forward filter, operator scan (fallback): k_fwd_chunks* reduce each chunk of 32 blocks to one K x K operator
//                  M = prod A diag(e_t) (K independent row recursions, exact power-of-two rescaling) and each tile
//                  to one operator, k_fwd_tilescan* scan the tile operators, k_fwd_replay* run the vector recursion
//                  from the exact incoming vector
//   k_bwd_maps     per block: the backward map j -> i = discrete_distribution(alpha'_t(.) A(., j))(u_t)
//   k_bwd_chunkmaps  composed maps per chunk / quarter chunk / tile; its last CTA scans the tile maps (bwd_scan_cta)
//   k_bwd_replay_reduce (K <= 8)  per quarter chunk: q_t = f_t[q_{t+1}] and the statistics pass in the same walk;
//                  k_bwd_replay + k_reduce_partial otherwise
//   k_mix_sample   mixture sampler: independent categorical draw per block
//   k_reduce_*     per-state (N, sum x, sum x^2), KxK transition counts incl. the phantom 0 -> q0
//                  transition, occupancy (FB.hpp:177-200); deterministic summation order
//
// Per-block arrays live in a chunk-interleaved order (Layout::perm) so that the threads of a warp,
// which own 32 consecutive chunks, read and write consecutive addresses at every step.
#include <math.h>

#include "../../include/hammlet_b200.h"
#include "hml_common.cuh"
#include "hml_kernels.h"

#include "hml_sweep_impl.cuh"

namespace hml {

int padded_states(int K) {
  static const int opts[] = {2, 3, 4, 5, 6, 8, 12, 16, 20, 32};
  for (int o : opts)
    if (K <= o) return K < 2 ? 0 : o;
  return 0;
}
int map_bytes(int KP) { return KP <= 8 ? 8 : (KP <= 16 ? 16 : 32); }
int chunks_per_tile(int) { return Layout::C; }

__global__ void k_unpermute(const uint8_t* states, const double2* bS, uint64_t B, int16_t* dst_states, double* dst_sum,
                            double* dst_sumsq) {
  for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t p = Layout::perm(b);
    if (dst_states) dst_states[b] = (int16_t)states[p];
    if (dst_sum) {
      const double2 v = bS[p];
      dst_sum[b] = v.x;
      dst_sumsq[b] = v.y;
    }
  }
}


// ---- run-length view of the sampled state sequence (what Records::record forms block by block, Records.hpp:166-188):
// a run starts at block 0 and wherever the state differs from the previous block's.  Three small kernels — heads per
// tile, exclusive scan of the tile counts, ordered write of (start position, state) — so that a recorded sweep moves
// one entry per run to the host instead of one per block.
// Segment mode (one sequence split over ranks): a run may continue across a rank border, so whether a rank's first
// block starts a run depends on the state of the previous rank's last block (`prev`, all-gathered per recorded sweep,
// RunCtx::last_states).  Every rank > 0 keeps a run start at its local position 0 — the device-side marginals need
// one (P[0] = 0) —: the rank's block 0 if that block begins exactly at the rank's first observation, else a VIRTUAL
// run in the previous rank's last state that covers the observations in front of the first boundary (they belong to a
// block owned by an earlier rank).  Whether position 0 is also a run start of the whole sequence (`border_real`) is
// remembered so that the merged outputs can drop the border where it never was one.
struct RunCtx {
  int rank, world;
  const double* heads;          // world x 4 (block count of every rank first), from the sweep's head exchange
  const uint64_t* last_states;  // world words: state of every rank's last block (~0: the rank has no block)
  uint32_t* border;             // [0] |= 1 if position 0 of this rank started a run of the whole sequence in a recorded
                                // iteration (k_seg_write with `accumulate`), [1] = the same for the last iteration only
};
__device__ __forceinline__ uint32_t run_prev_state(const RunCtx& c) {
  for (int r = c.rank - 1; r >= 0; --r)
    if (c.heads[kHeadWords * r] > 0.0) return (uint32_t)c.last_states[r];
  return 0u;  // unreachable: rank 0 always owns the block that starts at position 0
}
// does a virtual run precede the rank's blocks?
__device__ __forceinline__ bool run_virtual(const RunCtx& c, const uint32_t* starts, uint64_t B) {
  return c.world > 1 && c.rank > 0 && (B == 0 || starts[0] != 0u);
}

__device__ __forceinline__ bool seg_is_head(const uint8_t* states, uint64_t b, uint64_t B, bool virt, uint32_t prev) {
  if (b >= B) return false;
  if (b == 0) return virt ? states[Layout::perm(0)] != prev : true;
  return states[Layout::perm(b)] != states[Layout::perm(b - 1)];
}

// tiles: max(1, ceil(B / 1024)) in segment mode (a rank without blocks still has its virtual run)
__global__ void __launch_bounds__(1024) k_seg_count(const uint8_t* __restrict__ states, const uint32_t* __restrict__ starts,
                                                    uint64_t B, uint64_t ntiles, RunCtx ctx, uint32_t* tile_counts) {
  const bool virt = run_virtual(ctx, starts, B);
  const uint32_t prev = virt ? run_prev_state(ctx) : 0u;
  for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int n = __syncthreads_count(seg_is_head(states, tile * 1024 + threadIdx.x, B, virt, prev) ? 1 : 0);
    if (threadIdx.x == 0) tile_counts[tile] = (uint32_t)n + ((tile == 0 && virt) ? 1u : 0u);
  }
}

// the state of this rank's last block for the neighbour (segment mode, before k_seg_count)
__global__ void k_seg_last_state(const uint8_t* __restrict__ states, const unsigned long long* __restrict__ nblocks,
                                 uint64_t capacity, unsigned long long* send) {
  const uint64_t raw = *nblocks;
  const uint64_t B = raw <= capacity ? raw : 0;
  send[0] = B ? (unsigned long long)states[Layout::perm(B - 1)] : ~0ull;
}

// in-place exclusive scan of tile_counts[0..ntiles) by one CTA; tile_counts[ntiles] receives the total.
// Eight consecutive elements per thread and trip (8192 per trip): the marginal merge scans ~1.6e5 flags with it.
__device__ __forceinline__ void seg_scan_body(uint32_t* tile_counts, uint32_t ntiles) {
  constexpr int V = 8;
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_running;
  if (threadIdx.x == 0) s_running = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t base = 0; base < ntiles; base += 1024 * V) {
    const uint32_t i0 = base + threadIdx.x * V;
    uint32_t v[V];
    uint32_t mine = 0;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      v[k] = (i0 + k < ntiles) ? tile_counts[i0 + k] : 0u;
      mine += v[k];
    }
    uint32_t x = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += y;
      }
      s_warp[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const uint32_t before = s_running + (warp ? s_warp[warp - 1] : 0u);
    uint32_t run = before + x - mine;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      if (i0 + k < ntiles) tile_counts[i0 + k] = run;
      run += v[k];
    }
    __syncthreads();
    if (threadIdx.x == 1023) s_running = before + x;
    __syncthreads();
  }
  if (threadIdx.x == 0) tile_counts[ntiles] = s_running;
}
__global__ void __launch_bounds__(1024) k_seg_scan(uint32_t* tile_counts, uint32_t ntiles) { seg_scan_body(tile_counts, ntiles); }
// the same with the length read from device memory (the host does not know the number of runs of a recorded sweep)
__global__ void __launch_bounds__(1024) k_seg_scan_dev(uint32_t* tile_counts, const uint32_t* __restrict__ n_ptr) {
  seg_scan_body(tile_counts, *n_ptr);
}

__global__ void __launch_bounds__(1024) k_seg_write(const uint8_t* __restrict__ states, const uint32_t* __restrict__ starts,
                                                    uint64_t B, uint64_t ntiles, RunCtx ctx, int accumulate,
                                                    const uint32_t* __restrict__ tile_offsets,
                                                    uint32_t* __restrict__ seg_start, int16_t* __restrict__ seg_state) {
  __shared__ uint32_t s_warp[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool virt = run_virtual(ctx, starts, B);
  const uint32_t prev = (ctx.world > 1 && ctx.rank > 0) ? run_prev_state(ctx) : 0u;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (virt) {
      seg_start[0] = 0u;
      seg_state[0] = (int16_t)prev;
    }
    if (ctx.border) {
      const uint32_t real = (ctx.world > 1 && ctx.rank > 0 && !virt && states[Layout::perm(0)] != prev) ? 1u : 0u;
      ctx.border[1] = real;
      if (accumulate && real) ctx.border[0] = 1u;
    }
  }
  for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const uint64_t b = tile * 1024 + threadIdx.x;
    const bool head = seg_is_head(states, b, B, virt, prev);
    const unsigned bal = __ballot_sync(0xffffffffu, head);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    uint32_t before = 0;
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    if (head) {
      const uint32_t idx = tile_offsets[tile] + ((tile == 0 && virt) ? 1u : 0u) + before + __popc(bal & ((1u << lane) - 1u));
      seg_start[idx] = starts[b];
      seg_state[idx] = (int16_t)states[Layout::perm(b)];
    }
    __syncthreads();
  }
}

// ---- state marginals on the device (StateMarginals.hpp:51-137): the common refinement of all recorded
// segmentations as a sorted array of segment starts P[n] with one count per state and segment.  Adding an iteration
// with run starts R[m] (sorted, R[0] = 0) and run states:
//   k_mg_rank   every old start finds its run (upper bound in R), every run start its place among the old starts
//               (lower bound in P) and whether it is a new boundary
//   k_seg_scan  exclusive scan of the "new" flags
//   k_mg_write  old start i goes to i + #new run starts below it, a new run start j to (#old starts below it) + #new
//               run starts before it; counts = counts of the old segment that contains the position + 1 for the state
//               of the run that contains it
__device__ __forceinline__ uint32_t mg_lower_bound(const uint32_t* a, uint32_t n, uint32_t v) {  // first a[k] >= v
  uint32_t lo = 0, hi = n;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}
__device__ __forceinline__ uint32_t mg_upper_bound(const uint32_t* a, uint32_t n, uint32_t v) {  // first a[k] > v
  uint32_t lo = 0, hi = n;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (a[mid] <= v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256) k_mg_rank(const uint32_t* __restrict__ P, const uint32_t* __restrict__ n_ptr,
                                                 const uint32_t* __restrict__ R, const uint32_t* __restrict__ m_ptr,
                                                 uint32_t* __restrict__ run_of_old, uint32_t* __restrict__ olds_below,
                                                 uint32_t* __restrict__ is_new) {
  const uint32_t n = *n_ptr, m = *m_ptr;  // segments so far, runs of this iteration (device-side counts)
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n + m; k += gridDim.x * blockDim.x) {
    if (k < n) {
      run_of_old[k] = mg_upper_bound(R, m, P[k]) - 1;
    } else {
      const uint32_t j = k - n, v = R[j];
      const uint32_t lb = mg_lower_bound(P, n, v);
      olds_below[j] = lb;
      is_new[j] = (lb == n || P[lb] != v) ? 1u : 0u;
    }
  }
}

__global__ void __launch_bounds__(256) k_mg_write(const uint32_t* __restrict__ P, const uint32_t* __restrict__ n_ptr,
                                                  const uint16_t* __restrict__ cnt, const uint32_t* __restrict__ R,
                                                  const int16_t* __restrict__ rstate, const uint32_t* __restrict__ m_ptr,
                                                  const uint32_t* __restrict__ run_of_old, const uint32_t* __restrict__ olds_below,
                                                  const uint32_t* __restrict__ new_before, int K, uint32_t* __restrict__ P2,
                                                  uint16_t* __restrict__ cnt2, uint32_t* __restrict__ n_out) {
  const uint32_t n = *n_ptr, m = *m_ptr;
  if (blockIdx.x == 0 && threadIdx.x == 0) *n_out = n + new_before[m];  // segments after this iteration
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n + m; k += gridDim.x * blockDim.x) {
    uint32_t out, parent, pos;
    int state;
    if (k < n) {
      const uint32_t run = run_of_old[k];
      out = k + new_before[run + 1];
      parent = k;
      pos = P[k];
      state = rstate[run];
    } else {
      const uint32_t j = k - n;
      if (new_before[j + 1] == new_before[j]) continue;  // an old boundary: written by its old start
      out = olds_below[j] + new_before[j];
      parent = olds_below[j] - 1;                        // P[0] = 0 <= every position
      pos = R[j];
      state = rstate[j];
    }
    P2[out] = pos;
    for (int s = 0; s < K; ++s) cnt2[(size_t)out * K + s] = (uint16_t)(cnt[(size_t)parent * K + s] + (s == state ? 1 : 0));
  }
}

void launch_marginals_merge(const uint32_t* P, const uint32_t* n_ptr, uint64_t n_upper, const uint16_t* cnt, const uint32_t* R,
                            const int16_t* rstate, const uint32_t* m_ptr, uint64_t m_upper, uint32_t* run_of_old,
                            uint32_t* olds_below, uint32_t* new_flags, int K, uint32_t* P2, uint16_t* cnt2, uint32_t* n_out,
                            int sms, cudaStream_t s) {
  uint64_t g64 = (n_upper + m_upper + 255) / 256;  // the counts live on the device; the grid is sized from upper bounds
  int g = (int)(g64 > (uint64_t)sms * 8 ? (uint64_t)sms * 8 : g64);
  if (g < 1) g = 1;
  k_mg_rank<<<g, 256, 0, s>>>(P, n_ptr, R, m_ptr, run_of_old, olds_below, new_flags);
  k_seg_scan_dev<<<1, 1024, 0, s>>>(new_flags, m_ptr);  // exclusive, new_flags[m] = number of new boundaries
  k_mg_write<<<g, 256, 0, s>>>(P, n_ptr, cnt, R, rstate, m_ptr, run_of_old, olds_below, new_flags, K, P2, cnt2, n_out);
}

static RunCtx make_run_ctx(const SweepBuffers& b, const unsigned long long* last_states, uint32_t* border) {
  RunCtx c;
  c.rank = b.seg.world > 1 ? b.seg.rank : 0;
  c.world = b.seg.world > 1 ? b.seg.world : 1;
  c.heads = b.seg.heads;
  c.last_states = reinterpret_cast<const uint64_t*>(last_states);
  c.border = border;
  return c;
}
static uint64_t run_tiles(const SweepBuffers& b, uint64_t nblocks) {
  const uint64_t ntiles = (nblocks + 1023) / 1024;
  return (b.seg.world > 1 && ntiles == 0) ? 1 : ntiles;
}
void launch_segments_last_state(const SweepBuffers& b, unsigned long long* send, cudaStream_t s) {
  k_seg_last_state<<<1, 1, 0, s>>>(b.states, b.nblocks, b.capacity, send);
}
void launch_segments_count(const SweepBuffers& b, uint64_t nblocks, const unsigned long long* last_states,
                           uint32_t* tile_counts, int sms, cudaStream_t s) {
  const uint64_t ntiles = run_tiles(b, nblocks);
  const int g = (int)(ntiles < (uint64_t)sms * 2 ? ntiles : (uint64_t)sms * 2);
  k_seg_count<<<g, 1024, 0, s>>>(b.states, b.starts, nblocks, ntiles, make_run_ctx(b, last_states, nullptr), tile_counts);
  k_seg_scan<<<1, 1024, 0, s>>>(tile_counts, (uint32_t)ntiles);
}
void launch_segments_write(const SweepBuffers& b, uint64_t nblocks, const unsigned long long* last_states, uint32_t* border,
                           int accumulate, const uint32_t* tile_offsets, uint32_t* seg_start, int16_t* seg_state, int sms,
                           cudaStream_t s) {
  const uint64_t ntiles = run_tiles(b, nblocks);
  const int g = (int)(ntiles < (uint64_t)sms * 2 ? ntiles : (uint64_t)sms * 2);
  k_seg_write<<<g, 1024, 0, s>>>(b.states, b.starts, nblocks, ntiles, make_run_ctx(b, last_states, border), accumulate,
                                 tile_offsets, seg_start, seg_state);
}

size_t reduce_partials_doubles(int KP, int grid) { return (size_t)grid * 2 * KP + 64; }
size_t wide_scratch_doubles(int KP) { return KP > 8 ? (size_t)31 * KP * KP : 0; }
size_t wide_scratch_ints(int KP) { return KP > 8 ? (size_t)31 * KP : 0; }

#define HML_EXTERN(KP)                                                                                         \
  extern template int sweep_impl<KP>(const ModelHost&, const SweepBuffers&, const SweepLaunch&, cudaStream_t, \
                                     stage_cb_t, void*);                                                      \
  extern template int sequential_impl<KP>(const ModelHost&, const SweepBuffers&, const SweepLaunch&, cudaStream_t);
HML_EXTERN(2) HML_EXTERN(3) HML_EXTERN(4) HML_EXTERN(5) HML_EXTERN(6) HML_EXTERN(8) HML_EXTERN(12) HML_EXTERN(16)
HML_EXTERN(20) HML_EXTERN(32)

#define HML_DISPATCH_KP(KP_, CALL) \
  switch (KP_) {                   \
    case 2: CALL(2); break;        \
    case 3: CALL(3); break;        \
    case 4: CALL(4); break;        \
    case 5: CALL(5); break;        \
    case 6: CALL(6); break;        \
    case 8: CALL(8); break;        \
    case 12: CALL(12); break;      \
    case 16: CALL(16); break;      \
    case 20: CALL(20); break;      \
    case 32: CALL(32); break;      \
    default: break;                \
  }

int launch_sweep(const ModelHost& m, const SweepBuffers& b, const SweepLaunch& l, cudaStream_t s, stage_cb_t cb,
                 void* user) {
  const int KP = padded_states(m.K);
  int n = -2;  // unsupported number of states (-1: a carry exchange failed)
#define CALL(X) n = sweep_impl<X>(m, b, l, s, cb, user)
  HML_DISPATCH_KP(KP, CALL)
#undef CALL
  return n;
}

template <int KP> int fused_launch_kp(const SweepBuffers& b, const FusedArgs& a, int grid, cudaStream_t s);
template <int KP> int fused_max_grid_kp(int sms);
template <int KP> int chain_params_kp(ChainDev* ch, const unsigned long long* o64, const double* of, cudaStream_t s);
#define HML_EXTERN_FUSED(KP)                                                                                 \
  extern template int fused_launch_kp<KP>(const SweepBuffers&, const FusedArgs&, int, cudaStream_t);           \
  extern template int fused_max_grid_kp<KP>(int);                                                             \
  extern template int chain_params_kp<KP>(ChainDev*, const unsigned long long*, const double*, cudaStream_t);
HML_EXTERN_FUSED(2) HML_EXTERN_FUSED(3) HML_EXTERN_FUSED(4) HML_EXTERN_FUSED(5) HML_EXTERN_FUSED(6) HML_EXTERN_FUSED(8)
HML_EXTERN_FUSED(12) HML_EXTERN_FUSED(16) HML_EXTERN_FUSED(20) HML_EXTERN_FUSED(32)

int launch_sweep_fused(int KP, const SweepBuffers& b, const FusedArgs& a, int grid, cudaStream_t s) {
  int n = -2;
#define CALL(X) n = fused_launch_kp<X>(b, a, grid, s)
  HML_DISPATCH_KP(KP, CALL)
#undef CALL
  return n;
}
int fused_max_grid(int KP, int sms) {
  int n = 0;
#define CALL(X) n = fused_max_grid_kp<X>(sms)
  HML_DISPATCH_KP(KP, CALL)
#undef CALL
  return n;
}
int launch_chain_params(int KP, ChainDev* ch, const unsigned long long* out_u64, const double* out_f64, cudaStream_t s) {
  int n = -2;
#define CALL(X) n = chain_params_kp<X>(ch, out_u64, out_f64, s)
  HML_DISPATCH_KP(KP, CALL)
#undef CALL
  return n;
}

int launch_sweep_sequential(const ModelHost& m, const SweepBuffers& b, const SweepLaunch& l, cudaStream_t s) {
  const int KP = padded_states(m.K);
  int n = -1;
#define CALL(X) n = sequential_impl<X>(m, b, l, s)
  HML_DISPATCH_KP(KP, CALL)
#undef CALL
  return n;
}

void launch_unpermute(const SweepBuffers& b, int, uint64_t nblocks, int16_t* dst_states, double* dst_sum,
                      double* dst_sumsq, cudaStream_t s) {
  if (nblocks == 0) return;
  const int g = (int)((nblocks + 255) / 256 > 4096 ? 4096 : (nblocks + 255) / 256);
  k_unpermute<<<g, 256, 0, s>>>(b.states, b.bS, nblocks, dst_states, dst_sum, dst_sumsq);
}

void launch_seg_head(const SweepBuffers& b, uint64_t seg_len, unsigned long long seq, cudaStream_t s) {
  launch_k(k_seg_head, 1, 256, 0, s, b, (uint32_t)seg_len, seq);
}

void launch_block_stats(const SweepBuffers& b, int, uint64_t nblocks_hint, int sms, cudaStream_t s) {
  const uint64_t ntiles = (nblocks_hint + Layout::TB - 1) / Layout::TB;
  if (b.D > 1) {
    EmitMD<2> dummy;
    memset(&dummy, 0, sizeof(dummy));
    k_block_emit_md<2, true, false, false><<<grid_for(ntiles * Layout::TB, 256, sms, 8), 256, 0, s>>>(b, dummy, 0);
    return;
  }
  ModelDev<2> dummy;
  memset(&dummy, 0, sizeof(dummy));
  k_block_emit<2, true, false, false><<<grid_for(ntiles * Layout::TB, 256, sms, 8), 256, 0, s>>>(b, dummy, 0);
}

}  // namespace hml
