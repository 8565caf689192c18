// hammlet_b200 — carry exchange of the segment-split mode over NVLink peer memory (sm_100a).
//
// One sequence split over the GPUs of a box needs four tiny all-to-all exchanges per sweep: the
// partial block in front of each rank's first boundary, one KxK segment operator, one K->K segment map and
// the per-rank statistics; 32 bytes to 9 kB each.  An NCCL all-gather of that size costs about 20 us of
// launch and protocol latency, four of them a third of the sweep at 8 GPUs.  Here every rank owns a mailbox
// in its own HBM that its peers map through CUDA IPC; an exchange is ONE single-CTA kernel per rank that
//   1. stores the rank's payload straight into every peer's mailbox over NVLink (plain stores, 8-byte words),
//   2. fences system-wide and releases a sequence number next to each copy,
//   3. spins (acquire loads on its OWN mailbox, i.e. local L2) until all peers' numbers arrived,
//   4. copies the gathered payloads where the following kernels read them.
// No rank ever waits on a remote load.  Mailbox entries are double-buffered by sequence parity, so a rank
// that is one exchange ahead never overwrites an entry a slower rank is still reading; it cannot be two
// ahead because each exchange needs every rank's contribution.  A spin that exceeds its time budget (a
// peer died) raises a flag in mapped host memory and returns, so the host reports an error instead of
// hanging the device.
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

__global__ void __launch_bounds__(256)
    k_p2p_exchange(P2PPeers peers, int rank, int world, int slot, uint64_t seq, const uint64_t* __restrict__ send,
                   uint32_t words, uint64_t* __restrict__ recv, unsigned int* __restrict__ timeout_flag) {
  __shared__ int s_failed;
  const size_t entry = p2p_entry_offset((int)(seq & 1u), slot, rank, world);
  if (threadIdx.x == 0) s_failed = 0;
  // 1. my payload into every mailbox (my own included: the copy-out below treats all ranks alike)
  const uint32_t total = words * (uint32_t)world;
  for (uint32_t i = threadIdx.x; i < total; i += blockDim.x) {
    const uint32_t r = i / words, k = i - r * words;
    uint64_t* dst = reinterpret_cast<uint64_t*>(peers.box[r] + entry + kP2PHeader);
    dst[k] = send[k];
  }
  __threadfence_system();
  __syncthreads();
  // 2. publish
  if ((int)threadIdx.x < world)
    st_release_sys_u64(reinterpret_cast<uint64_t*>(peers.box[threadIdx.x] + entry), seq);
  // 3. wait for everybody's number in my own mailbox
  if ((int)threadIdx.x < world) {
    const uint64_t* flag = reinterpret_cast<const uint64_t*>(
        peers.box[rank] + p2p_entry_offset((int)(seq & 1u), slot, (int)threadIdx.x, world));
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
  if (s_failed) {
    if (threadIdx.x == 0) *timeout_flag = 1u;  // mapped host memory
    return;
  }
  // 4. gathered payloads, rank-major
  for (uint32_t i = threadIdx.x; i < total; i += blockDim.x) {
    const uint32_t r = i / words, k = i - r * words;
    const uint64_t* src = reinterpret_cast<const uint64_t*>(
        peers.box[rank] + p2p_entry_offset((int)(seq & 1u), slot, (int)r, world) + kP2PHeader);
    recv[i] = __ldcg(src + k);
  }
}

void launch_p2p_exchange(const P2PPeers& peers, int rank, int world, int slot, uint64_t seq, const void* send,
                         size_t bytes, void* recv, unsigned int* timeout_flag_dev, cudaStream_t s) {
  k_p2p_exchange<<<1, 256, 0, s>>>(peers, rank, world, slot, seq, reinterpret_cast<const uint64_t*>(send),
                                   (uint32_t)(bytes / 8), reinterpret_cast<uint64_t*>(recv), timeout_flag_dev);
}

}  // namespace hml
