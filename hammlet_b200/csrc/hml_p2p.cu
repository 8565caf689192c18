// hammlet_b200 — carry exchange of the segment-split mode over NVLink peer memory (sm_100a).
//
// One sequence split over the GPUs of a box needs four tiny all-to-all exchanges per sweep: the
// partial block in front of each rank's first boundary, one KxK segment operator, one K->K segment map and
// the per-rank statistics; 32 bytes to 9 kB each.  An NCCL all-gather of that size costs about 20 us of
// launch and protocol latency, four of them a third of the sweep at 8 GPUs.  Here every rank owns a mailbox
// in its own HBM that its peers map through CUDA IPC; an exchange is one CTA per rank — inside the kernel that produced the carry (hml_p2p.cuh) — that
//   1. stores the rank's payload straight into every peer's mailbox over NVLink, as 8-byte cells holding 4 bytes
//      of payload and the 32-bit sequence number of the exchange (an aligned 8-byte store arrives whole, so no
//      fence and no separate flag are needed: one one-way write per cell),
//   2. polls the cells of its OWN mailbox (local L2) until all peers' numbers match and assembles the payloads
//      where the following code reads them.
// No rank ever waits on a remote load.  Mailbox entries are double-buffered by sequence parity, so a rank
// that is one exchange ahead never overwrites an entry a slower rank is still reading; it cannot be two
// ahead because each exchange needs every rank's contribution.  A poll that exceeds its time budget (a
// peer died) raises a flag in mapped host memory and returns, so the host reports an error instead of
// hanging the device.
#include "hml_p2p.cuh"

namespace hml {

// stand-alone exchange (carries produced by kernels that do not embed the collective)
__global__ void __launch_bounds__(256)
    k_p2p_exchange(const P2PDev* __restrict__ d, int slot, uint64_t seq, const uint64_t* __restrict__ send, uint32_t words,
                   uint64_t* __restrict__ recv) {
  p2p_exchange_cta(d, slot, seq, send, words, recv);
}

void launch_p2p_exchange(const P2PDev* d, int slot, uint64_t seq, const void* send, size_t bytes, void* recv,
                         cudaStream_t s) {
  k_p2p_exchange<<<1, 256, 0, s>>>(d, slot, seq, reinterpret_cast<const uint64_t*>(send), (uint32_t)(bytes / 8),
                                   reinterpret_cast<uint64_t*>(recv));
}

}  // namespace hml
