// hammlet_b200 — threshold-driven block-boundary detection + ordered compaction (sm_100a).
//
// Replaces the per-sweep walk of Blocks<BreakpointArray>::next (Blocks/BreakpointArray.hpp:216-235):
// a block starts at t iff t == 0 or !(w[t] < thr).  The reference finds those positions by chasing
// uint16 skip pointers; here the fp32 weight array is streamed once per sweep (4 B/observation, the
// HBM-roofline term of SURVEY.md §8d) and the boundary positions are written in increasing order.
//
// Three kernels, none of which waits on another CTA (a single-pass decoupled look-back was measured
// first and is limited by its hop latency: 32 tiles per L2 round trip = 0.5 TB/s on this part):
//
//   k_detect_flags   THE streaming kernel.  One warp per 4096-observation tile; each lane keeps four
//                    16-byte streaming loads in flight (a warp reads 2 KB contiguous per step).  Flags
//                    become warp ballots: per group of 128 observations four 32-bit masks (mask c holds
//                    observations 4*lane + c).  The 32 groups' masks are collected lane-wise and
//                    written as one coalesced 512-byte store; the tile's boundary count is a popc sum.
//                    Output: T/8 bytes of bit masks, one count per tile, one per CTA (8 tiles).
//   k_scan_counts    single CTA: exclusive prefix over the per-CTA counts; total -> block count.
//   k_scatter_starts one warp per tile: reads the tile's 512 B of masks (mostly L2 hits), ranks by
//                    popc + warp scan, writes the positions in increasing order.
#include "hml_common.cuh"
#include "hml_kernels.h"

namespace hml {

constexpr int kTilesPerCta = 8;  // 8 warps, one tile each

__global__ void __launch_bounds__(256)
    k_detect_flags(const float4* __restrict__ w4, uint64_t T, float thr, int force_first, uint32_t num_tiles,
                   uint4* __restrict__ masks, uint32_t* __restrict__ tile_count, uint32_t* __restrict__ cta_count) {
  __shared__ uint32_t s_cnt[kTilesPerCta];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t tile = blockIdx.x * kTilesPerCta + warp;
  uint32_t cnt = 0;
  if (tile < num_tiles) {
    const uint64_t tbase = (uint64_t)tile * kTile;
    const float4* src = w4 + (uint64_t)tile * (kTile / 4) + lane;
    uint4 mine = make_uint4(0u, 0u, 0u, 0u);
    const bool full = tbase + kTile <= T && !(tile == 0 && force_first);
#pragma unroll 2
    for (int it = 0; it < 8; ++it) {
      float4 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = ld_stream_f4(src + (it * 4 + k) * 32);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int g = it * 4 + k;  // group of 128 observations inside the tile
        uint32_t f = 0;
        if (full) {
          f = (!(v[k].x < thr) ? 1u : 0u) | (!(v[k].y < thr) ? 2u : 0u) | (!(v[k].z < thr) ? 4u : 0u) |
              (!(v[k].w < thr) ? 8u : 0u);
        } else {
          const uint64_t p = tbase + (uint64_t)g * 128 + 4u * lane;
          f = ((p + 0 < T && !(v[k].x < thr)) ? 1u : 0u) | ((p + 1 < T && !(v[k].y < thr)) ? 2u : 0u) |
              ((p + 2 < T && !(v[k].z < thr)) ? 4u : 0u) | ((p + 3 < T && !(v[k].w < thr)) ? 8u : 0u);
          if (p == 0 && force_first) f |= 1u;
        }
        if (__any_sync(0xffffffffu, f != 0)) {
          const uint32_t m0 = __ballot_sync(0xffffffffu, f & 1u), m1 = __ballot_sync(0xffffffffu, f & 2u);
          const uint32_t m2 = __ballot_sync(0xffffffffu, f & 4u), m3 = __ballot_sync(0xffffffffu, f & 8u);
          cnt += __popc(m0) + __popc(m1) + __popc(m2) + __popc(m3);
          if (lane == g) mine = make_uint4(m0, m1, m2, m3);
        }
      }
    }
    masks[(uint64_t)tile * 32 + lane] = mine;
    if (lane == 0) tile_count[tile] = cnt;
  }
  if (lane == 0) s_cnt[warp] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
#pragma unroll
    for (int i = 0; i < kTilesPerCta; ++i) t += s_cnt[i];
    cta_count[blockIdx.x] = t;
  }
}

// exclusive prefix over n per-CTA counts (n <= ~2^20): 1024 threads, 4 values per thread per round
__global__ void __launch_bounds__(1024)
    k_scan_counts(const uint32_t* __restrict__ cta_count, uint32_t n, uint32_t* __restrict__ cta_off,
                  unsigned long long* __restrict__ nblocks_out, uint32_t* __restrict__ starts, uint64_t capacity,
                  uint64_t T) {
  __shared__ uint64_t s_warp[32];
  __shared__ uint64_t s_carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < n; base += 4096) {
    const uint32_t i0 = base + 4u * tid;
    uint32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = (i0 + k < n) ? cta_count[i0 + k] : 0u;
    const uint64_t mine = (uint64_t)v[0] + v[1] + v[2] + v[3];
    uint64_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint64_t u = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += u;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint64_t wpre = 0;
    for (int k = 0; k < warp; ++k) wpre += s_warp[k];
    uint64_t run = s_carry + wpre + incl - mine;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (i0 + k < n) cta_off[i0 + k] = (uint32_t)run;  // < 2^32 because T < 2^32
      run += v[k];
    }
    __syncthreads();
    if (tid == 1023) s_carry = run;
    __syncthreads();
  }
  if (tid == 0) {
    const uint64_t nb = s_carry;
    *nblocks_out = nb;
    if (nb <= capacity) starts[nb] = (uint32_t)T;  // sentinel: block b = [starts[b], starts[b+1])
  }
}

__global__ void __launch_bounds__(256)
    k_scatter_starts(const uint4* __restrict__ masks, const uint32_t* __restrict__ tile_count,
                     const uint32_t* __restrict__ cta_off, uint32_t num_tiles, uint32_t* __restrict__ starts,
                     uint64_t capacity) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t tile0 = blockIdx.x * kTilesPerCta;
  const uint32_t tile = tile0 + warp;
  if (tile >= num_tiles) return;
  // offset of this tile = CTA offset + counts of the CTA's earlier tiles
  uint32_t c = (lane < warp) ? tile_count[tile0 + lane] : 0u;
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  c = __shfl_sync(0xffffffffu, c, 0);
  if (tile_count[tile] == 0) return;
  const uint4 m = masks[(uint64_t)tile * 32 + lane];
  const uint32_t mine = __popc(m.x) + __popc(m.y) + __popc(m.z) + __popc(m.w);
  uint32_t incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += u;
  }
  uint64_t o = (uint64_t)cta_off[blockIdx.x] + c + (incl - mine);
  const uint32_t pbase = tile * (uint32_t)kTile + (uint32_t)lane * 128u;
  uint32_t any = m.x | m.y | m.z | m.w;
  while (any) {
    const int l = __ffs(any) - 1;  // sub-position 4*l .. 4*l+3 of this group
    any &= any - 1;
    const uint32_t p = pbase + 4u * l;
    if ((m.x >> l) & 1u) { if (o < capacity) starts[o] = p; ++o; }
    if ((m.y >> l) & 1u) { if (o < capacity) starts[o] = p + 1; ++o; }
    if ((m.z >> l) & 1u) { if (o < capacity) starts[o] = p + 2; ++o; }
    if ((m.w >> l) & 1u) { if (o < capacity) starts[o] = p + 3; ++o; }
  }
}

size_t detect_scratch_bytes(uint64_t T) {
  const uint64_t tiles = (T + kTile - 1) / kTile;
  const uint64_t ctas = (tiles + kTilesPerCta - 1) / kTilesPerCta;
  return tiles * 32 * sizeof(uint4) + (tiles + 2 * ctas + 16) * sizeof(uint32_t);
}

int launch_detect(const float* w, uint64_t T, float thr, int force_first, void* scratch, uint32_t* starts,
                  uint64_t capacity, unsigned long long* nblocks_out, cudaStream_t s, stage_cb_t cb, void* user) {
  const uint32_t tiles = (uint32_t)((T + kTile - 1) / kTile);
  const uint32_t ctas = (tiles + kTilesPerCta - 1) / kTilesPerCta;
  uint4* masks = reinterpret_cast<uint4*>(scratch);
  uint32_t* tile_count = reinterpret_cast<uint32_t*>(masks + (uint64_t)tiles * 32);
  uint32_t* cta_count = tile_count + tiles;
  uint32_t* cta_off = cta_count + ctas;
  if (cb) cb(user, "detect_flags");
  k_detect_flags<<<ctas, 256, 0, s>>>(reinterpret_cast<const float4*>(w), T, thr, force_first, tiles, masks, tile_count,
                                      cta_count);
  if (cb) cb(user, "detect_scan");
  k_scan_counts<<<1, 1024, 0, s>>>(cta_count, ctas, cta_off, nblocks_out, starts, capacity, T);
  if (cb) cb(user, "detect_scatter");
  k_scatter_starts<<<ctas, 256, 0, s>>>(masks, tile_count, cta_off, tiles, starts, capacity);
  return 3;
}

}  // namespace hml
