// hammlet_b200 — the head of a rank's segment (split sequences), shared by the sweep kernels and the candidate scatter.
#pragma once
#include "hml_common.cuh"
#include "hml_kernels.h"
#include "hml_p2p.cuh"

namespace hml {

// (sum x, sum x^2) of data dimension d over the local observations [s, e) from the integral arrays
// (Statistics/IntegralArray.hpp:104-124, 176-182): cell-local running sums + double-double cell offsets
__device__ __forceinline__ void range_sums_dim(const SweepBuffers& buf, int d, uint32_t s, uint32_t e, double& sx,
                                               double& sq) {
  const double2* pq = buf.pq + (size_t)d * buf.pq_stride;
  const double2 ps = pq[s], pe = pq[e];
  sx = pe.x - ps.x;
  sq = pe.y - ps.y;
  const uint32_t cs = s >> kCellLog2, ce = e >> kCellLog2;
  if (cs != ce) {
    const double4* cp = buf.cell_pref + (size_t)d * buf.cell_stride;
    const double4 a = cp[cs], z = cp[ce];
    sx += (z.x - a.x) + (z.y - a.y);
    sq += (z.z - a.z) + (z.w - a.w);
  }
}

// The observations in front of this rank's first boundary belong to a block that starts on an earlier rank; their
// partial statistics travel with the rank's block count (seg.send_head).  Called by all threads of one CTA once the
// block list is complete (the first boundary is read through L2: another CTA of the same launch may have written it).
// seq != 0: the head exchange runs here too (peer mailboxes), else the caller exchanges afterwards.
__device__ __forceinline__ void seg_head_cta(const SweepBuffers& buf, uint32_t seg_len, unsigned long long seq) {
  if (threadIdx.x == 0) {
    const uint64_t raw = *buf.nblocks;
    const uint64_t B = raw <= buf.capacity ? raw : 0;  // (an overflowing list reads as empty: the host repeats the sweep)
    if (buf.seg.overflow) *buf.seg.overflow = raw > buf.capacity ? 1ull : 0ull;
    const uint32_t e = B ? __ldcg(buf.starts) : seg_len;
    buf.seg.send_head[0] = (double)B;
    buf.seg.send_head[1] = (double)e;
    for (int d = 0; d < kMaxDims; ++d) {  // one pair per data dimension
      double sx = 0.0, sq = 0.0;
      if (e > 0 && d < buf.D) range_sums_dim(buf, d, 0u, e, sx, sq);
      buf.seg.send_head[2 + 2 * d] = sx;
      buf.seg.send_head[3 + 2 * d] = sq;
    }
  }
  if (seq && buf.seg.p2p) {
    __threadfence();
    __syncthreads();
    p2p_exchange_cta(buf.seg.p2p, kSlotHeads, seq, reinterpret_cast<const uint64_t*>(buf.seg.send_head), kHeadWords,
                     reinterpret_cast<uint64_t*>(const_cast<double*>(buf.seg.heads)));
  }
}

}  // namespace hml
