// hammlet_b200 — device side of the peer-memory carry exchange (see hml_p2p.cu for the protocol).
// p2p_exchange_cta is called by ALL threads of one CTA, from inside the kernel that produced the payload,
// so that a compute step and the collective that follows it are one launch.
#pragma once
#include "hml_common.cuh"
#include "hml_kernels.h"

namespace hml {

__device__ __forceinline__ void st_release_sys_u64(uint64_t* p, uint64_t v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_acquire_sys_u64(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// All-gather of `words` 8-byte words per rank: send -> recv[rank-major].  The payload must be visible to the
// calling CTA (written by it before a __syncthreads, or by an earlier kernel).  On return every thread of the
// CTA may read recv.  A peer that does not show up within kP2PTimeoutNs raises *d->timeout_flag (mapped host
// memory); the call then returns with recv incomplete and the host reports the failure after the sweep.
__device__ __forceinline__ void p2p_exchange_cta(const P2PDev* __restrict__ d, int slot, uint64_t seq,
                                                 const uint64_t* send, uint32_t words, uint64_t* recv) {
  __shared__ int s_failed;
  const int rank = d->rank, world = d->world;
  const size_t entry = p2p_entry_offset((int)(seq & 1u), slot, rank, world);
  if (threadIdx.x == 0) s_failed = 0;
  // 1. my payload into every mailbox (my own included: the copy-out below treats all ranks alike)
  const uint32_t total = words * (uint32_t)world;
  for (uint32_t i = threadIdx.x; i < total; i += blockDim.x) {
    const uint32_t r = i / words, k = i - r * words;
    uint64_t* dst = reinterpret_cast<uint64_t*>(d->peers.box[r] + entry + kP2PHeader);
    dst[k] = __ldcg(send + k);
  }
  __threadfence_system();
  __syncthreads();
  // 2. publish, 3. wait for everybody's number in my own mailbox
  if ((int)threadIdx.x < world) {
    st_release_sys_u64(reinterpret_cast<uint64_t*>(d->peers.box[threadIdx.x] + entry), seq);
    const uint64_t* flag = reinterpret_cast<const uint64_t*>(
        d->peers.box[rank] + p2p_entry_offset((int)(seq & 1u), slot, (int)threadIdx.x, world));
    const uint64_t t0 = global_timer_ns();
    uint32_t spins = 0;
    while (ld_acquire_sys_u64(flag) < seq) {
      if ((++spins & 1023u) == 0 && global_timer_ns() - t0 > kP2PTimeoutNs) {
        s_failed = 1;
        break;
      }
    }
  }
  __syncthreads();
  if (s_failed && threadIdx.x == 0) *d->timeout_flag = 1u;
  // 4. gathered payloads, rank-major
  for (uint32_t i = threadIdx.x; i < total; i += blockDim.x) {
    const uint32_t r = i / words, k = i - r * words;
    const uint64_t* src = reinterpret_cast<const uint64_t*>(
        d->peers.box[rank] + p2p_entry_offset((int)(seq & 1u), slot, (int)r, world) + kP2PHeader);
    recv[i] = __ldcg(src + k);
  }
  __threadfence();
  __syncthreads();
}

}  // namespace hml
