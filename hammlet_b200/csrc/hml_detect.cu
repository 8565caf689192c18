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
//
// Pyramid mode (default).  The reference's walk is sub-linear in T: its uint16 skip pointers jump over
// runs of small weights (BreakpointArray.hpp:150-182,216-235).  The device analogue is a one-level max
// pyramid built at load: smax[g] = max of the 32 weights of sub-block g (NaN counts as +inf).  A sub-block
// can hold a boundary only if !(smax[g] < thr), so
//   k_detect_pyramid  reads the pyramid (T/8 bytes) and then only the hot sub-blocks (128 B each, eight
//                     loads in flight per lane); one ballot per hot sub-block is its 32-bit mask
//   k_scatter_pyramid ranks the masks of the hot sub-blocks as above
// The boundary set is identical to the streaming kernels' (tests compare both with the oracle); the traffic
// drops from 4 T bytes to T/8 + 128 * (hot sub-blocks).
#include "hml_common.cuh"
#include "hml_kernels.h"

namespace hml {

constexpr int kTilesPerCta = 8;  // 8 warps, one tile each
constexpr int kSub = 32;         // observations per pyramid entry
constexpr int kSubsPerTile = kTile / kSub;  // 128: four per lane

// Per tile: its boundary count and its offset inside the CTA, packed as count | offset << 16 (count <= 4096,
// offset <= 7 * 4096); per CTA: the total.  Called by all threads after s_cnt was filled and synchronised.
__device__ __forceinline__ void finish_cta_counts(const uint32_t* s_cnt, uint32_t cnt, uint32_t tile, uint32_t num_tiles,
                                                  uint32_t* __restrict__ tile_count, uint32_t* __restrict__ cta_count) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    uint32_t pre = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < kTilesPerCta; ++i) {
      if (i < warp) pre += s_cnt[i];
      tot += s_cnt[i];
    }
    if (tile < num_tiles) tile_count[tile] = cnt | (pre << 16);
    if (warp == 0) cta_count[blockIdx.x] = tot;
  }
}

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
  }
  if (lane == 0) s_cnt[warp] = cnt;
  __syncthreads();
  finish_cta_counts(s_cnt, cnt, tile, num_tiles, tile_count, cta_count);
}

// smax[g] = max over the sub-block's weights that exist (t < T); NaN -> +inf (a NaN weight is always a boundary)
__global__ void __launch_bounds__(256) k_build_pyramid(const float* __restrict__ w, uint64_t T, uint64_t num_subs,
                                                       float* __restrict__ smax) {
  const float inf = __int_as_float(0x7f800000);
  const int lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t g0 = warp * 32; g0 < num_subs; g0 += nwarps * 32) {
    float mine = -inf;
#pragma unroll 4
    for (int k = 0; k < 32; ++k) {  // sub-block g0 + k: one coalesced 128-byte row per step
      const uint64_t p = (g0 + k) * kSub + lane;
      float v = -inf;
      if (p < T) {
        v = w[p];
        if (v != v) v = inf;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
      if (lane == k) mine = v;
    }
    if (g0 + lane < num_subs) smax[g0 + lane] = mine;
  }
}

__global__ void __launch_bounds__(256)
    k_detect_pyramid(const float* __restrict__ w, const float4* __restrict__ smax4, uint64_t T, float thr, int force_first,
                     uint32_t num_tiles, uint4* __restrict__ masks, uint4* __restrict__ tile_hot,
                     uint32_t* __restrict__ tile_count, uint32_t* __restrict__ cta_count,
                     unsigned long long* __restrict__ hot_counter) {
  __shared__ uint32_t s_cnt[kTilesPerCta];
  __shared__ uint32_t s_hot[kTilesPerCta];
  __shared__ uint8_t s_list[kTilesPerCta][kSubsPerTile + 4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t tile = blockIdx.x * kTilesPerCta + warp;
  uint32_t cnt = 0, nhot = 0;
  if (tile < num_tiles) {
    const uint64_t tbase = (uint64_t)tile * kTile;
    const float* __restrict__ wt = w + tbase + lane;
    const float4 m = smax4[(uint64_t)tile * 32 + lane];  // sub-blocks 4*lane .. 4*lane+3
    uint32_t hot = (!(m.x < thr) ? 1u : 0u) | (!(m.y < thr) ? 2u : 0u) | (!(m.z < thr) ? 4u : 0u) | (!(m.w < thr) ? 8u : 0u);
    const bool first = tile == 0 && force_first;
    if (first && lane == 0) hot |= 1u;
    const uint32_t hb0 = __ballot_sync(0xffffffffu, hot & 1u), hb1 = __ballot_sync(0xffffffffu, hot & 2u);
    const uint32_t hb2 = __ballot_sync(0xffffffffu, hot & 4u), hb3 = __ballot_sync(0xffffffffu, hot & 8u);
    const uint32_t n0 = __popc(hb0), n1 = __popc(hb1), n2 = __popc(hb2);
    nhot = n0 + n1 + n2 + __popc(hb3);
    uint4 mine = make_uint4(0u, 0u, 0u, 0u);
    if (nhot) {
      // list of the hot sub-blocks (any order: every mask lands in its owner's register)
      const uint32_t lt = lanemask_lt();
      if (hot & 1u) s_list[warp][__popc(hb0 & lt)] = (uint8_t)(4 * lane);
      if (hot & 2u) s_list[warp][n0 + __popc(hb1 & lt)] = (uint8_t)(4 * lane + 1);
      if (hot & 4u) s_list[warp][n0 + n1 + __popc(hb2 & lt)] = (uint8_t)(4 * lane + 2);
      if (hot & 8u) s_list[warp][n0 + n1 + n2 + __popc(hb3 & lt)] = (uint8_t)(4 * lane + 3);
      __syncwarp();
      const bool full = tbase + kTile <= T && !first;
      const uint32_t rem = full ? (uint32_t)kTile : (uint32_t)(T > tbase ? T - tbase : 0);  // valid observations
      for (uint32_t i = 0; i < nhot; i += 4) {
        // four sub-blocks per round: the loads are issued before the first ballot consumes one
        const uint32_t quad = *reinterpret_cast<const uint32_t*>(&s_list[warp][i]);
        const uint32_t left = nhot - i;
        float v[4];
        uint32_t sub[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          sub[k] = (quad >> (8 * k)) & 0xffu;
          v[k] = (k == 0 || (uint32_t)k < left) ? wt[sub[k] * kSub] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (k == 0 || (uint32_t)k < left) {  // warp-uniform
            bool f = !(v[k] < thr);
            if (!full) {
              const uint32_t p = sub[k] * kSub + lane;
              f = (f && p < rem) || (first && p == 0);
            }
            const uint32_t mk = __ballot_sync(0xffffffffu, f);
            cnt += __popc(mk);
            if ((uint32_t)lane == (sub[k] >> 2)) {
              const uint32_t c = sub[k] & 3u;
              mine.x = c == 0 ? mk : mine.x;
              mine.y = c == 1 ? mk : mine.y;
              mine.z = c == 2 ? mk : mine.z;
              mine.w = c == 3 ? mk : mine.w;
            }
          }
        }
      }
      if (hot) masks[(uint64_t)tile * 32 + lane] = mine;
    }
    if (lane == 0) tile_hot[tile] = make_uint4(hb0, hb1, hb2, hb3);
  }
  if (lane == 0) {
    s_cnt[warp] = cnt;
    s_hot[warp] = nhot;
  }
  __syncthreads();
  finish_cta_counts(s_cnt, cnt, tile, num_tiles, tile_count, cta_count);
  if (threadIdx.x == 0) {
    uint32_t hsum = 0;
#pragma unroll
    for (int i = 0; i < kTilesPerCta; ++i) hsum += s_hot[i];
    if (hsum) atomicAdd(hot_counter, (unsigned long long)hsum);
  }
}

__global__ void __launch_bounds__(256)
    k_scatter_pyramid(const uint4* __restrict__ masks, const uint4* __restrict__ tile_hot,
                      const uint32_t* __restrict__ tile_count, const uint32_t* __restrict__ cta_off, uint32_t num_tiles,
                      uint32_t* __restrict__ starts, uint64_t capacity) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t tile0 = blockIdx.x * kTilesPerCta;
  const uint32_t tile = tile0 + warp;
  if (tile >= num_tiles) return;
  const uint32_t tc = tile_count[tile];
  if ((tc & 0xffffu) == 0) return;
  const uint32_t c = tc >> 16;  // offset of this tile inside the CTA
  const uint4 hb = tile_hot[tile];
  const bool any_hot = ((hb.x | hb.y | hb.z | hb.w) >> lane) & 1u;
  uint4 m = make_uint4(0u, 0u, 0u, 0u);
  if (any_hot) m = masks[(uint64_t)tile * 32 + lane];
  // components of sub-blocks that were not hot hold zeros (the producer starts from zero)
  const uint32_t mine = __popc(m.x) + __popc(m.y) + __popc(m.z) + __popc(m.w);
  uint32_t incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += u;
  }
  uint64_t o = (uint64_t)cta_off[blockIdx.x] + c + (incl - mine);
  const uint32_t pbase = tile * (uint32_t)kTile + (uint32_t)lane * (4u * kSub);
  const uint32_t word[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    uint32_t bits = word[k];
    while (bits) {
      const int l = __ffs(bits) - 1;
      bits &= bits - 1;
      if (o < capacity) starts[o] = pbase + (uint32_t)k * kSub + (uint32_t)l;
      ++o;
    }
  }
}

// exclusive prefix over n per-CTA counts (n <= ~2^20): 1024 threads, 4 values per thread per round
__global__ void __launch_bounds__(1024)
    k_scan_counts(const uint32_t* __restrict__ cta_count, uint32_t n, uint32_t* __restrict__ cta_off,
                  unsigned long long* __restrict__ nblocks_out, uint32_t* __restrict__ starts, uint64_t capacity,
                  uint64_t T, unsigned long long* __restrict__ hot_counter) {
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
    hot_counter[1] = hot_counter[0];  // hot sub-blocks of this pass (pyramid mode), for the traffic accounting
    hot_counter[0] = 0;
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
  const uint32_t tc = tile_count[tile];
  if ((tc & 0xffffu) == 0) return;
  const uint32_t c = tc >> 16;  // offset of this tile inside the CTA
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
  return tiles * 32 * sizeof(uint4) + tiles * sizeof(uint4) + (tiles + 2 * ctas + 16) * sizeof(uint32_t) + 64;
}

size_t pyramid_floats(uint64_t T) { return (T + kTile - 1) / kTile * kSubsPerTile; }

void launch_build_pyramid(const float* w, uint64_t T, float* smax, int sms, cudaStream_t s) {
  const uint64_t subs = pyramid_floats(T);
  uint64_t blocks = (subs + 255) / 256;
  if (blocks > (uint64_t)sms * 16) blocks = (uint64_t)sms * 16;
  k_build_pyramid<<<(unsigned)blocks, 256, 0, s>>>(w, T, subs, smax);
}

int launch_detect(const float* w, const float* smax, uint64_t T, float thr, int force_first, void* scratch,
                  uint32_t* starts, uint64_t capacity, unsigned long long* nblocks_out, cudaStream_t s, stage_cb_t cb,
                  void* user) {
  const uint32_t tiles = (uint32_t)((T + kTile - 1) / kTile);
  const uint32_t ctas = (tiles + kTilesPerCta - 1) / kTilesPerCta;
  uint4* masks = reinterpret_cast<uint4*>(scratch);
  uint4* tile_hot = masks + (uint64_t)tiles * 32;
  unsigned long long* hot_counter = reinterpret_cast<unsigned long long*>(tile_hot + tiles);  // [0] running, [1] last pass
  uint32_t* tile_count = reinterpret_cast<uint32_t*>(hot_counter + 2);
  uint32_t* cta_count = tile_count + tiles;
  uint32_t* cta_off = cta_count + ctas;
  const bool pyramid = smax != nullptr && thr == thr;  // a NaN threshold makes every position a boundary: stream
  if (cb) cb(user, pyramid ? "detect_pyramid" : "detect_flags");
  if (pyramid)
    k_detect_pyramid<<<ctas, 256, 0, s>>>(w, reinterpret_cast<const float4*>(smax), T, thr, force_first, tiles, masks,
                                          tile_hot, tile_count, cta_count, hot_counter);
  else
    k_detect_flags<<<ctas, 256, 0, s>>>(reinterpret_cast<const float4*>(w), T, thr, force_first, tiles, masks, tile_count,
                                        cta_count);
  if (cb) cb(user, "detect_scan");
  k_scan_counts<<<1, 1024, 0, s>>>(cta_count, ctas, cta_off, nblocks_out, starts, capacity, T, hot_counter);
  if (cb) cb(user, "detect_scatter");
  if (pyramid)
    k_scatter_pyramid<<<ctas, 256, 0, s>>>(masks, tile_hot, tile_count, cta_off, tiles, starts, capacity);
  else
    k_scatter_starts<<<ctas, 256, 0, s>>>(masks, tile_count, cta_off, tiles, starts, capacity);
  return 3;
}

// device address of the hot sub-block count of the last pass
const unsigned long long* detect_hot_count_ptr(const void* scratch, uint64_t T) {
  const uint64_t tiles = (T + kTile - 1) / kTile;
  const uint4* masks = reinterpret_cast<const uint4*>(scratch);
  return reinterpret_cast<const unsigned long long*>(masks + tiles * 32 + tiles) + 1;
}

}  // namespace hml
