// hammlet_b200 — k_detect_compact: threshold-driven block-boundary detection + ordered compaction.
//
// Replaces the per-sweep walk of Blocks<BreakpointArray>::next (Blocks/BreakpointArray.hpp:216-235):
// a block starts at t iff t == 0 or !(w[t] < thr).  The reference finds those positions by chasing
// uint16 skip pointers; here the fp32 weight array is streamed once per sweep (4 B/observation, the
// HBM-roofline term of SURVEY.md §8d) and the boundary positions are written in increasing order.
//
// Shape: persistent CTAs of 256 threads pull 4096-observation tiles from an atomic ticket counter.
// Each thread issues four 16-byte streaming loads (tile k-th quarter, float4 index = k*256 + tid, so a
// warp reads 512 contiguous bytes per instruction), and the loads of the NEXT tile are issued before
// the current tile's scan and look-back so HBM stays busy through the synchronisation points.
// Flags -> ranks by warp ballot + popc; the 32 (quarter, warp) counts of a tile are scanned by
// shuffles; tiles are chained by a single-word decoupled look-back (flag | epoch | count), so the
// kernel makes one pass over the weights and needs no zeroing of the descriptors between sweeps.
#include "hml_common.cuh"
#include "hml_kernels.h"

namespace hml {

namespace {
// descriptor word: [63:36] epoch  [35:34] status  [33:0] value
constexpr uint64_t kStatusAggregate = 1, kStatusPrefix = 2;
__device__ __forceinline__ uint64_t pack_desc(uint32_t epoch, uint64_t status, uint64_t value) {
  return ((uint64_t)epoch << 36) | (status << 34) | value;
}
}  // namespace

__global__ void __launch_bounds__(256, 6)
    k_detect_compact(const float4* __restrict__ w4, uint64_t T, float thr, int force_first, uint32_t num_tiles,
                     uint64_t* __restrict__ desc, uint32_t epoch, unsigned long long* __restrict__ ticket,
                     unsigned long long ticket_base, uint32_t* __restrict__ starts, uint64_t capacity,
                     unsigned long long* __restrict__ nblocks_out) {
  __shared__ uint32_t s_wcount[32];
  __shared__ uint64_t s_excl;
  __shared__ uint32_t s_tile[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned lt = lanemask_lt();

  if (tid == 0) s_tile[0] = (uint32_t)(atomicAdd(ticket, 1ull) - ticket_base);
  __syncthreads();
  uint32_t tile = s_tile[0];
  int buf = 0;
  float4 v[4];
  if (tile < num_tiles) {
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = ld_stream_f4(w4 + (uint64_t)tile * (kTile / 4) + k * 256 + tid);
  }
  while (tile < num_tiles) {
    // ---- flags of this tile
    const uint64_t tbase = (uint64_t)tile * kTile;
    uint32_t nib[4], pre[4], wtot[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint64_t p = tbase + 4ull * (k * 256 + tid);
      uint32_t f = 0;
      f |= (p + 0 < T && !(v[k].x < thr)) ? 1u : 0u;
      f |= (p + 1 < T && !(v[k].y < thr)) ? 2u : 0u;
      f |= (p + 2 < T && !(v[k].z < thr)) ? 4u : 0u;
      f |= (p + 3 < T && !(v[k].w < thr)) ? 8u : 0u;
      if (p == 0 && force_first) f |= 1u;
      nib[k] = f;
      const uint32_t any = __ballot_sync(0xffffffffu, f != 0);
      uint32_t before = 0, total = 0;
      if (any) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t m = __ballot_sync(0xffffffffu, (f >> c) & 1u);
          before += __popc(m & lt);
          total += __popc(m);
        }
      }
      pre[k] = before;
      wtot[k] = total;
    }
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < 4; ++k) s_wcount[k * 8 + warp] = wtot[k];
    }
    // ---- claim the next tile and put its loads in flight before any waiting
    if (tid == 0) s_tile[buf ^ 1] = (uint32_t)(atomicAdd(ticket, 1ull) - ticket_base);
    __syncthreads();
    const uint32_t next_tile = s_tile[buf ^ 1];
    buf ^= 1;
    if (next_tile < num_tiles) {
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = ld_stream_f4(w4 + (uint64_t)next_tile * (kTile / 4) + k * 256 + tid);
    }
    // ---- scan the 32 (quarter, warp) counts; every warp does it redundantly (no extra barrier)
    uint32_t incl = s_wcount[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += u;
    }
    const uint32_t tile_total = __shfl_sync(0xffffffffu, incl, 31);
    const uint32_t excl = incl - s_wcount[lane];
    uint32_t wbase[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) wbase[k] = __shfl_sync(0xffffffffu, excl, k * 8 + warp);

    // ---- decoupled look-back (warp 0)
    if (warp == 0) {
      if (lane == 0)
        st_volatile_u64(desc + tile, pack_desc(epoch, tile == 0 ? kStatusPrefix : kStatusAggregate, tile_total));
      uint64_t running = 0;
      int64_t look = (int64_t)tile - 1;
      while (look >= 0) {
        const int64_t idx = look - lane;
        uint64_t d = 0;
        uint64_t status = kStatusPrefix;  // lanes before the sequence start behave like a zero prefix
        uint64_t value = 0;
        if (idx >= 0) {
          do {
            d = ld_volatile_u64(desc + idx);
          } while ((uint32_t)(d >> 36) != epoch || ((d >> 34) & 3) == 0);
          status = (d >> 34) & 3;
          value = d & ((1ull << 34) - 1);
        }
        const uint32_t has_prefix = __ballot_sync(0xffffffffu, status == kStatusPrefix);
        const int first = __ffs(has_prefix) - 1;  // nearest predecessor holding an inclusive prefix
        uint64_t contrib = (first < 0 || lane <= first) ? value : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
        running += contrib;
        if (first >= 0) break;
        look -= 32;
      }
      if (lane == 0) {
        st_volatile_u64(desc + tile, pack_desc(epoch, kStatusPrefix, running + tile_total));
        s_excl = running;
        if (tile == num_tiles - 1) {
          const uint64_t nb = running + tile_total;
          *nblocks_out = nb;
          if (nb <= capacity) starts[nb] = (uint32_t)T;  // sentinel (starts holds capacity + 1 entries)
        }
      }
    }
    __syncthreads();
    const uint64_t gbase = s_excl;
    // ---- ordered scatter of the boundary positions
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      uint32_t f = nib[k];
      if (f) {
        uint64_t o = gbase + wbase[k] + pre[k];
        const uint32_t p = (uint32_t)(tbase + 4ull * (k * 256 + tid));
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if ((f >> c) & 1u) {
            if (o < capacity) starts[o] = p + c;
            ++o;
          }
        }
      }
    }
    tile = next_tile;
  }
}

int detect_grid_size(int sms) {
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_detect_compact, 256, 0);
  if (per_sm < 1) per_sm = 1;
  return sms * per_sm;
}

void launch_detect_compact(const float* w, uint64_t T, float thr, int force_first, uint64_t* desc, uint32_t epoch,
                           unsigned long long* ticket, unsigned long long ticket_base, uint32_t* starts,
                           uint64_t capacity, unsigned long long* nblocks_out, int grid, cudaStream_t s) {
  const uint32_t num_tiles = (uint32_t)((T + kTile - 1) / kTile);
  k_detect_compact<<<grid, 256, 0, s>>>(reinterpret_cast<const float4*>(w), T, thr, force_first, num_tiles, desc, epoch,
                                        ticket, ticket_base, starts, capacity, nblocks_out);
}

}  // namespace hml
