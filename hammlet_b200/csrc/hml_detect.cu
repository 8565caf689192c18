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
// pyramid built at load: entry g = max of the 32 weights of sub-block g, rounded up to bf16 (NaN counts as
// +inf).  A sub-block can hold a boundary only if !(entry < thr), so
//   k_detect_hot   reads the pyramid (T/16 bytes) and then only the hot sub-blocks (128 B each); per hot
//                  sub-block one 32-bit boundary mask; the last CTA to finish scans the per-span counts
//   k_scatter_hot  writes the positions of the set bits in increasing order
// The boundary set is identical to the streaming kernels' (tests compare both with the oracle); the traffic
// drops from 4 T bytes to T/16 + 128 * (hot sub-blocks).
#include "hml_common.cuh"
#include "hml_kernels.h"
#include "hml_seg_head.cuh"

namespace hml {

constexpr int kTilesPerCta = 8;  // 8 warps, one tile each
constexpr int kSub = 32;         // observations per pyramid entry
constexpr int kSpanSubs = 4096;               // pyramid entries per CTA of the pyramid kernels
constexpr int kSpanObs = kSpanSubs * kSub;    // 131072 observations

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

// ---- pyramid mode -------------------------------------------------------------------------------------------
// Entry g of the pyramid is the maximum of the 32 weights of sub-block g, rounded UP to bf16 (2 bytes per 32
// observations).  Rounding up keeps the test conservative: a sub-block whose true maximum reaches the threshold
// is always flagged; the few extra sub-blocks flagged by the rounding are read and turn out empty.  The decision
// itself is always taken on the exact fp32 weights, so the boundary set is bit-identical to the stream kernels'.
__device__ __forceinline__ uint16_t bf16_round_up(float v) {
  if (v != v) return 0x7f80u;  // a NaN weight is always a boundary: +inf
  const uint32_t b = __float_as_uint(v);
  if (b & 0x80000000u) return (uint16_t)(b >> 16);  // negative: truncation rounds towards +inf
  return (uint16_t)((b + 0xffffu) >> 16);            // smallest bf16 >= v (overflows to +inf)
}

// entries past the last real sub-block (up to a whole number of spans) hold -inf: never hot
__global__ void __launch_bounds__(256) k_build_pyramid(const float* __restrict__ w, uint64_t T, uint64_t num_entries,
                                                       uint16_t* __restrict__ smax) {
  const float inf = __int_as_float(0x7f800000);
  const int lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t g0 = warp * 32; g0 < num_entries; g0 += nwarps * 32) {
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
    if (g0 + lane < num_entries) smax[g0 + lane] = bf16_round_up(mine);
  }
}

__device__ __forceinline__ uint4 ld_stream_u4(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// hot bits of the eight bf16 entries of one 16-byte pyramid word
__device__ __forceinline__ uint32_t hot_bits8(const uint4 v, float thr) {
  const uint32_t wd[4] = {v.x, v.y, v.z, v.w};
  uint32_t h = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float lo = __uint_as_float(wd[i] << 16), hi = __uint_as_float(wd[i] & 0xffff0000u);
    h |= (!(lo < thr) ? 1u : 0u) << (2 * i);
    h |= (!(hi < thr) ? 1u : 0u) << (2 * i + 1);
  }
  return h;
}

// k_detect_hot: one CTA per span of 4096 pyramid entries (131072 observations).
//   phase 1  every thread reads two 16-byte pyramid words (both loads in flight); the hot sub-blocks are compacted,
//            in position order, into a list in shared memory (ballot-free: packed warp scans of the popcounts)
//   phase 2  the hot sub-blocks are read 16 per warp and round: a quarter warp reads one sub-block as float4 per
//            lane, four loads in flight per lane; an OR butterfly over the quarter assembles the 32-bit boundary mask
//   phase 3  exclusive scan of the mask popcounts; (sub-block, offset, mask) triples go to global memory,
//            8 bytes per hot sub-block, and the span's boundary count to span_info
//   last CTA (atomic ticket): exclusive scan over the span counts -> span_off, total -> *nblocks_out and the
//            sentinel starts[total] = T.  No CTA ever waits for another one.
__global__ void __launch_bounds__(256)
    k_detect_hot(const float* __restrict__ w, const uint4* __restrict__ smax8, uint64_t T, float thr, int force_first,
                 uint32_t nspans, uint2* __restrict__ hot, uint32_t* __restrict__ span_info, uint32_t* __restrict__ span_off,
                 unsigned int* __restrict__ ticket, unsigned long long* __restrict__ nblocks_out,
                 uint32_t* __restrict__ starts, uint64_t capacity, unsigned long long* __restrict__ hot_counter) {
  pdl_enter();
  __shared__ uint16_t s_list[kSpanSubs];
  __shared__ uint32_t s_mask[kSpanSubs];
  __shared__ uint32_t s_warp[2][8];
  __shared__ uint32_t s_red[8];
  __shared__ uint64_t s_red64[8];
  __shared__ bool s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t span = blockIdx.x;
  const bool first = span == 0 && force_first;
  const uint64_t nsubs = (T + kSub - 1) / kSub;
  const bool tail_span = ((uint64_t)span + 1) * kSpanObs > T;  // only the last span can reach past the sequence

  // ---- phase 1
  uint4 pv[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) pv[k] = ld_stream_u4(smax8 + (uint64_t)span * (kSpanSubs / 8) + k * 256 + tid);
  uint32_t h[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    h[k] = hot_bits8(pv[k], thr);
    // entries past the last sub-block never count (a NaN or -inf threshold flags even their -inf padding)
    if (tail_span) {
      const uint64_t e0 = (uint64_t)span * kSpanSubs + (uint64_t)(k * 256 + tid) * 8;
      if (e0 + 8 > nsubs) h[k] = e0 < nsubs ? (h[k] & ((1u << (uint32_t)(nsubs - e0)) - 1u)) : 0u;
    }
  }
  if (first && tid == 0) h[0] |= 1u;
  const uint32_t c0 = __popc(h[0]), c1 = __popc(h[1]);
  uint32_t incl = c0 | (c1 << 16);  // both counts scanned at once: a warp total is at most 256
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += u;
  }
  if (lane == 31) {
    s_warp[0][warp] = incl & 0xffffu;
    s_warp[1][warp] = incl >> 16;
  }
  __syncthreads();
  uint32_t base0 = 0, base1 = 0, nhot = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint32_t a = s_warp[0][i], b = s_warp[1][i];
    if (i < warp) {
      base0 += a;
      base1 += b;
    }
    nhot += a + b;
    base1 += a;  // every k = 0 entry precedes the k = 1 entries
  }
  {
    uint32_t p0 = base0 + (incl & 0xffffu) - c0, p1 = base1 + (incl >> 16) - c1;
    uint32_t b0 = h[0], b1 = h[1];
    while (b0) {
      const int l = __ffs(b0) - 1;
      b0 &= b0 - 1;
      s_list[p0++] = (uint16_t)(tid * 8 + l);
    }
    while (b1) {
      const int l = __ffs(b1) - 1;
      b1 &= b1 - 1;
      s_list[p1++] = (uint16_t)((256 + tid) * 8 + l);
    }
  }
  __syncthreads();

  // ---- phase 2
  {
    const int q = lane >> 3, l8 = lane & 7;
    const float4* wq = reinterpret_cast<const float4*>(w + (uint64_t)span * kSpanObs) + l8;
    for (uint32_t i0 = warp * 16; i0 < nhot; i0 += 8 * 16) {
      uint32_t sub[4];
      float4 v[4];
      bool ok[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t idx = i0 + 4 * k + q;
        ok[k] = idx < nhot;
        sub[k] = ok[k] ? s_list[idx] : 0u;
        if (ok[k]) v[k] = ld_stream_f4(wq + sub[k] * (kSub / 4));
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint32_t f = 0;
        if (ok[k])
          f = (!(v[k].x < thr) ? 1u : 0u) | (!(v[k].y < thr) ? 2u : 0u) | (!(v[k].z < thr) ? 4u : 0u) |
              (!(v[k].w < thr) ? 8u : 0u);
        // OR over the quarter warp: three butterfly steps stay inside groups of eight lanes (a redux.sync with a
        // partial member mask compiles to a slow generic path)
        uint32_t m = f << (4 * l8);
        m |= __shfl_xor_sync(0xffffffffu, m, 1);
        m |= __shfl_xor_sync(0xffffffffu, m, 2);
        m |= __shfl_xor_sync(0xffffffffu, m, 4);
        if (ok[k] && l8 == 0) {
          if (tail_span) {  // positions past the end of the sequence never count
            const uint64_t pos0 = ((uint64_t)span * kSpanSubs + sub[k]) * kSub;
            if (pos0 + kSub > T) m = pos0 < T ? (m & ((1u << (uint32_t)(T - pos0)) - 1u)) : 0u;
          }
          if (first && sub[k] == 0) m |= 1u;
          s_mask[i0 + 4 * k + q] = m;
        }
      }
    }
  }
  __syncthreads();

  // ---- phase 3: thread t owns entries [t * per, t * per + per)
  const uint32_t per = (nhot + 255) >> 8;
  const uint32_t lo = min(nhot, (uint32_t)tid * per), hi = min(nhot, lo + per);
  uint32_t mine = 0;
  for (uint32_t i = lo; i < hi; ++i) mine += __popc(s_mask[i]);
  uint32_t inc2 = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t u = __shfl_up_sync(0xffffffffu, inc2, o);
    if (lane >= o) inc2 += u;
  }
  if (lane == 31) s_red[warp] = inc2;
  __syncthreads();
  uint32_t wpre = 0, total = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (i < warp) wpre += s_red[i];
    total += s_red[i];
  }
  {
    uint32_t off = wpre + inc2 - mine;
    uint2* dst = hot + (uint64_t)span * kSpanSubs;
    for (uint32_t i = lo; i < hi; ++i) {
      const uint32_t m = s_mask[i];
      dst[i] = make_uint2((uint32_t)s_list[i] | (off << 12), m);
      off += __popc(m);
    }
  }
  if (tid == 0) {
    span_info[span] = total | (nhot << 18);
    __threadfence();
    s_last = atomicAdd(ticket, 1u) == nspans - 1;
  }
  __syncthreads();
  if (!s_last) return;

  // ---- last CTA: exclusive scan over the span counts
  __threadfence();
  const uint32_t per2 = (nspans + 255) >> 8;
  const uint32_t lo2 = min(nspans, (uint32_t)tid * per2), hi2 = min(nspans, lo2 + per2);
  uint64_t sum = 0, hsum = 0;
  for (uint32_t i = lo2; i < hi2; ++i) {
    const uint32_t v = __ldcg(span_info + i);
    sum += v & 0x3ffffu;
    hsum += v >> 18;
  }
  uint64_t inc3 = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint64_t u = __shfl_up_sync(0xffffffffu, inc3, o);
    if (lane >= o) inc3 += u;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) hsum += __shfl_xor_sync(0xffffffffu, hsum, o);
  __syncthreads();  // s_red is reused below
  if (lane == 31) s_red64[warp] = inc3;
  if (lane == 0) s_red[warp] = (uint32_t)hsum;  // a sequence holds fewer than 2^32 / 32 sub-blocks
  __syncthreads();
  uint64_t wpre3 = 0, tot3 = 0, hot_total = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (i < warp) wpre3 += s_red64[i];
    tot3 += s_red64[i];
    hot_total += s_red[i];
  }
  uint64_t run = wpre3 + inc3 - sum;
  for (uint32_t i = lo2; i < hi2; ++i) {
    span_off[i] = (uint32_t)run;  // < 2^32 because T < 2^32
    run += __ldcg(span_info + i) & 0x3ffffu;
  }
  if (tid == 0) {
    *nblocks_out = tot3;
    if (tot3 <= capacity) starts[tot3] = (uint32_t)T;  // sentinel: block b = [starts[b], starts[b+1])
    hot_counter[1] = hot_total;                        // hot sub-blocks of this pass, for the traffic accounting
    *ticket = 0u;
  }
}

// k_scatter_hot: thread per hot sub-block: its boundary positions go to starts[span_off + offset ...]
__global__ void __launch_bounds__(128)
    k_scatter_hot(const uint2* __restrict__ hot, const uint32_t* __restrict__ span_info, const uint32_t* __restrict__ span_off,
                  uint32_t* __restrict__ starts, uint64_t capacity) {
  pdl_enter();
  const uint32_t span = blockIdx.x;
  const uint32_t info = span_info[span];
  if ((info & 0x3ffffu) == 0) return;
  const uint32_t nhot = info >> 18;
  const uint64_t base = span_off[span];
  const uint2* src = hot + (uint64_t)span * kSpanSubs;
  for (uint32_t i = threadIdx.x; i < nhot; i += blockDim.x) {
    const uint2 e = src[i];
    uint64_t o = base + (e.x >> 12);
    const uint32_t pos = (span * (uint32_t)kSpanSubs + (e.x & 0xfffu)) * (uint32_t)kSub;
    uint32_t bits = e.y;
    while (bits) {
      const int l = __ffs(bits) - 1;
      bits &= bits - 1;
      if (o < capacity) starts[o] = pos + (uint32_t)l;
      ++o;
    }
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

// ------------------------------------------------------------------------------------------------
// Candidate mode (default).  Between two sweeps of a chain the threshold moves by a hair (it follows the smallest
// sampled variance), so almost all of the weight array can never hold a boundary.  The candidate list keeps, in
// position order, every position whose weight is not below a FLOOR (0.75 x the threshold it was built for) together
// with that weight; as long as thr >= floor the boundary set {t : !(w[t] < thr)} is a subset of the list and one
// coalesced pass over the list's weights (8 bytes per candidate, ~1.3 candidates per block) finds it: same set, same
// order, bit for bit, 6 us instead of 88 at T = 1e9.  The list is (re)built by one pyramid pass at the floor whenever
// the threshold drops below the floor or the list has become much longer than needed (hml_api.cu: run_detect).
//   k_cand_gather   cand_w[i] = w[starts[i]], cand_pos[i] = starts[i]          (build)
//   k_cand_count    float4 loads, flags, per-CTA counts; the last CTA scans them, writes the block count + sentinel
//   k_cand_scatter  flags again, block scan, starts[offset + rank] = cand_pos[i]
constexpr int kCandPerThread = 8;                      // two float4
constexpr int kCandPerCta = 256 * kCandPerThread;      // 2048

__global__ void __launch_bounds__(256) k_cand_gather(const float* __restrict__ w, const double2* __restrict__ pq,
                                                     const uint32_t* __restrict__ starts, uint32_t n,
                                                     float* __restrict__ cand_w, uint32_t* __restrict__ cand_pos,
                                                     double2* __restrict__ cand_pq) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t p = starts[i];
    cand_pos[i] = p;
    cand_w[i] = w[p];
    // the integral pair of the candidate travels with it (univariate data): the random 16-byte gathers into the
    // T-sized integral arrays — a 64-byte DRAM burst each — are paid once per list, not twice per block and sweep
    if (cand_pq) cand_pq[i] = pq[p];
  }
}

// flags of the eight candidates of a thread: bits 0-3 = candidates base + 4 tid .. + 3, bits 4-7 = the same + 1024
__device__ __forceinline__ uint32_t cand_flags(const float4* __restrict__ cw4, uint32_t nc, float thr) {
  const uint32_t base = blockIdx.x * (uint32_t)kCandPerCta;
  uint32_t f = 0;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const uint32_t i0 = base + half * 1024u + 4u * threadIdx.x;
    if (i0 < nc) {  // the arrays are padded to a multiple of 4
      const float4 v = cw4[i0 >> 2];
      uint32_t m = (!(v.x < thr) ? 1u : 0u) | (!(v.y < thr) ? 2u : 0u) | (!(v.z < thr) ? 4u : 0u) | (!(v.w < thr) ? 8u : 0u);
      const uint32_t left = nc - i0;
      if (left < 4) m &= (1u << left) - 1u;
      f |= m << (4 * half);
    }
  }
  return f;
}

__global__ void __launch_bounds__(256)
    k_cand_count(const float4* __restrict__ cw4, uint32_t nc, float thr, uint32_t nctas, uint32_t* __restrict__ cta_count,
                 uint32_t* __restrict__ cta_off, unsigned int* __restrict__ ticket, unsigned long long* __restrict__ nblocks_out,
                 uint32_t* __restrict__ starts, uint64_t capacity, uint64_t T, const double2* __restrict__ pq,
                 double2* __restrict__ spq) {
  pdl_enter();
  __shared__ uint32_t s_warp[8];
  __shared__ bool s_last;
  const uint32_t f = cand_flags(cw4, nc, thr);
  uint32_t c = __popc(f);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t tot = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += s_warp[i];
    cta_count[blockIdx.x] = tot;
    __threadfence();
    s_last = atomicAdd(ticket, 1u) == nctas - 1;
  }
  __syncthreads();
  if (!s_last) return;
  // the last CTA to arrive: exclusive scan of the per-CTA counts, block count, sentinel
  __threadfence();
  __shared__ uint32_t s_part[256];
  const uint32_t per = (nctas + 255) / 256;
  const uint32_t lo = threadIdx.x * per, hi = min(nctas, lo + per);
  uint32_t sum = 0;
  for (uint32_t i = lo; i < hi; ++i) sum += __ldcg(cta_count + i);
  s_part[threadIdx.x] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    for (int i = 0; i < 256; ++i) {
      const uint32_t v = s_part[i];
      s_part[i] = run;
      run += v;
    }
    *nblocks_out = run;
    if ((uint64_t)run <= capacity) {
      starts[run] = (uint32_t)T;  // sentinel: block b = [starts[b], starts[b + 1])
      if (spq) spq[run] = pq[T];
    }
    *ticket = 0;
  }
  __syncthreads();
  uint32_t run = s_part[threadIdx.x];
  for (uint32_t i = lo; i < hi; ++i) {
    const uint32_t v = __ldcg(cta_count + i);
    cta_off[i] = run;
    run += v;
  }
}

// kHead (split sequence, peer mailboxes): the CTA that arrives last also forms the head partial of the rank's segment
// and runs the head exchange (seg_head_cta) — the block list is complete at that point — instead of a kernel of its own
template <bool kHead>
__global__ void __launch_bounds__(256)
    k_cand_scatter(const float4* __restrict__ cw4, const uint32_t* __restrict__ cpos, uint32_t nc, float thr,
                   const uint32_t* __restrict__ cta_off, uint32_t* __restrict__ starts, uint64_t capacity,
                   const double2* __restrict__ cand_pq, double2* __restrict__ spq, SweepBuffers head, uint32_t seg_len,
                   unsigned long long seq) {
  pdl_enter();
  __shared__ uint32_t s_warp[8];
  __shared__ uint16_t s_rank[kCandPerCta];  // rank of every candidate of the CTA among its boundaries, 0xffff: none
  const uint32_t f = cand_flags(cw4, nc, thr);
  // counts of the two halves packed (each <= 4 per thread, <= 1024 per CTA): one scan ranks both
  const uint32_t mine = __popc(f & 0xfu) | (__popc(f >> 4) << 16);
  uint32_t incl = mine;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += u;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  uint32_t before = 0, total = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint32_t v = s_warp[i];
    if (i < warp) before += v;
    total += v;
  }
  const uint32_t excl = before + incl - mine;
  uint32_t r0 = excl & 0xffffu;                       // first half: ranks among first-half flags
  uint32_t r1 = (total & 0xffffu) + (excl >> 16);     // second half comes after the whole first half
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const uint32_t m = (f >> (4 * half)) & 0xfu;
    uint32_t& r = half ? r1 : r0;
    uint16_t v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      v[k] = (m >> k) & 1u ? (uint16_t)r : (uint16_t)0xffffu;
      r += (m >> k) & 1u;
    }
    *reinterpret_cast<uint2*>(&s_rank[half * 1024 + 4 * threadIdx.x]) =
        make_uint2((uint32_t)v[0] | ((uint32_t)v[1] << 16), (uint32_t)v[2] | ((uint32_t)v[3] << 16));
  }
  __syncthreads();
  // Second phase, thread per candidate with consecutive threads on consecutive candidates: positions and integral
  // pairs are read as contiguous pieces and the survivors land next to each other in block order — the emission kernel
  // reads the pairs of block b and b + 1 from adjacent slots.  (Four consecutive candidates per thread, loaded where
  // they are stored, cost a 64-byte stride between lanes and one DRAM round trip after the other: 22 us against 11.)
  const uint32_t base = blockIdx.x * (uint32_t)kCandPerCta;
  const uint64_t off = cta_off[blockIdx.x];
#pragma unroll
  for (int k = 0; k < kCandPerThread; ++k) {
    const uint32_t li = k * 256u + threadIdx.x;
    const uint32_t r = s_rank[li];
    if (r != 0xffffu && off + r < capacity) {
      starts[off + r] = cpos[base + li];
      if (spq) spq[off + r] = cand_pq[base + li];
    }
  }
  if constexpr (kHead) {
    if (last_cta_arrives(head.tickets + kTicketScatter)) seg_head_cta(head, seg_len, seq);
  }
}

void launch_cand_gather(const float* w, const double2* pq, const uint32_t* starts, uint32_t n, float* cand_w,
                        uint32_t* cand_pos, double2* cand_pq, int sms, cudaStream_t s) {
  if (n == 0) return;
  int g = (int)((n + 255) / 256);
  if (g > sms * 16) g = sms * 16;
  k_cand_gather<<<g, 256, 0, s>>>(w, pq, starts, n, cand_w, cand_pos, cand_pq);
}

// boundaries among the candidates; cta_scratch holds 2 * ctas + 1 words (counts, offsets, ticket = last word, zeroed
// once by the caller and left at zero by the kernel).  Returns the number of launches.
int launch_detect_candidates(const float* cand_w, const uint32_t* cand_pos, const double2* cand_pq, uint32_t nc, float thr,
                             uint32_t* cta_scratch, uint32_t scratch_ctas, uint32_t* starts, double2* spq, const double2* pq,
                             uint64_t capacity, uint64_t T, unsigned long long* nblocks_out, cudaStream_t s, stage_cb_t cb,
                             void* user, const SweepBuffers* head, unsigned long long head_seq) {
  if (!cand_pq) spq = nullptr;
  const uint32_t ctas = (nc + kCandPerCta - 1) / kCandPerCta;
  uint32_t* cta_count = cta_scratch;
  uint32_t* cta_off = cta_scratch + scratch_ctas;
  unsigned int* ticket = cta_scratch + 2 * scratch_ctas;
  if (cb) cb(user, "detect_cand");
  launch_k(k_cand_count, ctas, 256, 0, s, reinterpret_cast<const float4*>(cand_w), nc, thr, ctas, cta_count, cta_off, ticket,
           nblocks_out, starts, capacity, T, pq, spq);
  if (cb) cb(user, "detect_scatter");
  if (head != nullptr && head->tickets != nullptr) {
    launch_k(k_cand_scatter<true>, ctas, 256, 0, s, reinterpret_cast<const float4*>(cand_w), cand_pos, nc, thr,
             (const uint32_t*)cta_off, starts, capacity, cand_pq, spq, *head, (uint32_t)T, head_seq);
  } else {
    SweepBuffers none;
    memset(&none, 0, sizeof(none));
    launch_k(k_cand_scatter<false>, ctas, 256, 0, s, reinterpret_cast<const float4*>(cand_w), cand_pos, nc, thr,
             (const uint32_t*)cta_off, starts, capacity, cand_pq, spq, none, (uint32_t)0, 0ull);
  }
  return 2;
}
uint32_t cand_ctas(uint32_t nc) { return (nc + kCandPerCta - 1) / kCandPerCta; }

// scratch layout: [stream mode] masks (512 B per tile), per-tile counts, per-CTA counts and offsets;
// [pyramid mode] hot triples (8 B per pyramid entry), span_info, span_off, ticket; [both] hot counters
struct DetectScratch {
  uint4* masks;
  uint32_t *tile_count, *cta_count, *cta_off;
  uint2* hot;
  uint32_t *span_info, *span_off;
  unsigned int* ticket;
  unsigned long long* hot_counter;  // [0] unused, [1] hot sub-blocks of the last pyramid pass
  size_t bytes;
};

static DetectScratch carve_scratch(void* scratch, uint64_t T) {
  const uint64_t tiles = (T + kTile - 1) / kTile;
  const uint64_t ctas = (tiles + kTilesPerCta - 1) / kTilesPerCta;
  const uint64_t spans = (T + kSpanObs - 1) / kSpanObs;
  DetectScratch d;
  char* p = reinterpret_cast<char*>(scratch);
  auto take = [&](size_t n) {
    char* r = p;
    p += (n + 255) / 256 * 256;
    return r;
  };
  d.masks = reinterpret_cast<uint4*>(take(tiles * 32 * sizeof(uint4)));
  d.hot = reinterpret_cast<uint2*>(take(spans * kSpanSubs * sizeof(uint2)));
  d.tile_count = reinterpret_cast<uint32_t*>(take(tiles * 4));
  d.cta_count = reinterpret_cast<uint32_t*>(take(ctas * 4));
  d.cta_off = reinterpret_cast<uint32_t*>(take(ctas * 4));
  d.span_info = reinterpret_cast<uint32_t*>(take(spans * 4));
  d.span_off = reinterpret_cast<uint32_t*>(take((spans + 1) * 4));
  d.hot_counter = reinterpret_cast<unsigned long long*>(take(16));
  d.ticket = reinterpret_cast<unsigned int*>(take(4));
  d.bytes = (size_t)(p - reinterpret_cast<char*>(scratch));
  return d;
}

size_t detect_scratch_bytes(uint64_t T) { return carve_scratch(nullptr, T).bytes; }

size_t pyramid_entries(uint64_t T) { return (T + kSpanObs - 1) / kSpanObs * kSpanSubs; }

void launch_build_pyramid(const float* w, uint64_t T, uint16_t* smax, int sms, cudaStream_t s) {
  const uint64_t n = pyramid_entries(T);
  uint64_t blocks = (n + 255) / 256;
  if (blocks > (uint64_t)sms * 16) blocks = (uint64_t)sms * 16;
  k_build_pyramid<<<(unsigned)blocks, 256, 0, s>>>(w, T, n, smax);
}

int launch_detect(const float* w, const uint16_t* smax, uint64_t T, float thr, int force_first, void* scratch,
                  uint32_t* starts, uint64_t capacity, unsigned long long* nblocks_out, cudaStream_t s, stage_cb_t cb,
                  void* user) {
  const DetectScratch d = carve_scratch(scratch, T);
  if (smax != nullptr) {
    const uint32_t spans = (uint32_t)((T + kSpanObs - 1) / kSpanObs);
    if (cb) cb(user, "detect_hot");
    launch_k(k_detect_hot, spans, 256, 0, s, w, reinterpret_cast<const uint4*>(smax), T, thr, force_first, spans, d.hot,
             d.span_info, d.span_off, d.ticket, nblocks_out, starts, capacity, d.hot_counter);
    if (cb) cb(user, "detect_scatter");
    launch_k(k_scatter_hot, spans, 128, 0, s, d.hot, d.span_info, d.span_off, starts, capacity);
    return 2;
  }
  const uint32_t tiles = (uint32_t)((T + kTile - 1) / kTile);
  const uint32_t ctas = (tiles + kTilesPerCta - 1) / kTilesPerCta;
  if (cb) cb(user, "detect_flags");
  k_detect_flags<<<ctas, 256, 0, s>>>(reinterpret_cast<const float4*>(w), T, thr, force_first, tiles, d.masks, d.tile_count,
                                      d.cta_count);
  if (cb) cb(user, "detect_scan");
  k_scan_counts<<<1, 1024, 0, s>>>(d.cta_count, ctas, d.cta_off, nblocks_out, starts, capacity, T);
  if (cb) cb(user, "detect_scatter");
  k_scatter_starts<<<ctas, 256, 0, s>>>(d.masks, d.tile_count, d.cta_off, tiles, starts, capacity);
  return 3;
}

// device address of the hot sub-block count of the last pyramid pass
const unsigned long long* detect_hot_count_ptr(const void* scratch, uint64_t T) {
  return carve_scratch(const_cast<void*>(scratch), T).hot_counter + 1;
}

}  // namespace hml
