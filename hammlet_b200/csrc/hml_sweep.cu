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
//   k_fwd_chunks   forward filter as a scan of KxK operators M_t = A diag(e_t): each chunk of 32
//                  blocks is reduced to one operator (K independent row recursions, exact
//                  power-of-two rescaling), 32 chunk operators to one tile operator
//   k_fwd_tilescan single CTA: prefix over tile operators -> normalised forward vector entering each tile
//   k_fwd_replay   per chunk: the reference's vector recursion from the exact incoming vector; emits,
//                  per block, the backward map j -> i = discrete_distribution(alpha'_t(.) A(., j))(u_t)
//                  and the composed map of the chunk (backward sampling as map composition)
//   k_bwd_scan     single CTA: suffix composition of chunk maps -> state following each chunk
//   k_bwd_replay   per chunk: q_t = f_t[q_{t+1}]
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
  k_seg_head<<<1, 256, 0, s>>>(b, (uint32_t)seg_len, seq);
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
