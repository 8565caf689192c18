// hammlet_b200 — device side of the peer-memory carry exchange (see hml_p2p.cu for the protocol).
// p2p_exchange_cta is called by ALL threads of one CTA, from inside the kernel that produced the payload,
// so that a compute step and the collective that follows it are one launch.
#pragma once
#include "hml_common.cuh"
#include "hml_kernels.h"

namespace hml {

__device__ __forceinline__ void st_relaxed_sys_u64(uint64_t* p, uint64_t v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_relaxed_sys_u64(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// All-gather of `words` 8-byte words per rank: send -> recv[rank-major].  The payload must be visible to the
// calling CTA (written by it before a __syncthreads, or by an earlier kernel).  On return every thread of the
// CTA may read recv.
//
// Wire format: every 8-byte store carries 4 bytes of payload and the 32-bit sequence number of the exchange, and
// an aligned 8-byte store arrives as a whole, so data and "it is there" travel in ONE NVLink write: no fence, no
// separate flag, the latency of an exchange is a single one-way write plus the poll of local memory.  The
// receiver polls each 8-byte cell of its own mailbox until the number matches.  A peer that does not show up
// within kP2PTimeoutNs raises *d->timeout_flag (mapped host memory); the call then returns with recv incomplete
// and the host reports the failure after the sweep.
__device__ __forceinline__ void p2p_exchange_cta(const P2PDev* __restrict__ d, int slot, uint64_t seq,
                                                 const uint64_t* send, uint32_t words, uint64_t* recv) {
  __shared__ int s_failed;
  const int rank = d->rank, world = d->world;
  const size_t entry = p2p_entry_offset((int)(seq & 1u), slot, rank, world);
  const uint64_t tag = (uint64_t)(uint32_t)seq << 32;
  if (threadIdx.x == 0) s_failed = 0;
  __syncthreads();
  // 1. my payload, 4 bytes per cell, into every mailbox (my own included: the gather below treats all ranks alike)
  const uint32_t cells = 2 * words;
  for (uint32_t i = threadIdx.x; i < cells * (uint32_t)world; i += blockDim.x) {
    const uint32_t r = i / cells, j = i - r * cells;
    const uint64_t v = __ldcg(send + (j >> 1));
    const uint64_t half = (j & 1u) ? (v >> 32) : (v & 0xffffffffull);
    st_relaxed_sys_u64(reinterpret_cast<uint64_t*>(d->peers.box[r] + entry) + j, tag | half);
  }
  // 2. gather: poll the cells of every rank's entry in my own mailbox
  const uint64_t t0 = global_timer_ns();
  for (uint32_t i = threadIdx.x; i < words * (uint32_t)world; i += blockDim.x) {
    const uint32_t r = i / words, k = i - r * words;
    const uint64_t* cell = reinterpret_cast<const uint64_t*>(
                               d->peers.box[rank] + p2p_entry_offset((int)(seq & 1u), slot, (int)r, world)) + 2 * k;
    uint64_t lo, hi;
    uint32_t spins = 0;
    bool ok = true;
    while (((lo = ld_relaxed_sys_u64(cell)) >> 32) != (tag >> 32) || ((hi = ld_relaxed_sys_u64(cell + 1)) >> 32) != (tag >> 32)) {
      if ((++spins & 1023u) == 0 && global_timer_ns() - t0 > kP2PTimeoutNs) {
        ok = false;
        break;
      }
    }
    if (ok)
      recv[i] = (lo & 0xffffffffull) | (hi << 32);
    else
      s_failed = 1;
  }
  __threadfence();
  __syncthreads();
  if (s_failed && threadIdx.x == 0) *d->timeout_flag = 1u;
}

}  // namespace hml
