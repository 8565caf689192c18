// hammlet_b200 — one-time load kernels (sm_100a).
//
//   k_maxlet_level   MaxletTransform, wavelet.hpp:97-188: absolute Haar detail coefficients with the
//                    reference's fixed pairwise fp32 summation tree and its running-product level
//                    normaliser; 12 levels per pass, applied recursively to the per-tile sums.
//   k_bp_weights     HaarBreakpointWeights, wavelet.hpp:68-93, in closed form (see below) + the
//                    weight multiplier of main.cpp:332-334.
//   k_integral_*     Statistics<IntegralArray> ctor, Statistics/IntegralArray.hpp:136-191: integral
//                    arrays of (x, x^2); here fp64, forward, cell-local (4096) + double-double cell
//                    offsets, so block sums carry ~1e-12 absolute error instead of the 1e-3 of the
//                    reference's fp32 arrays.
//   k_sum_odd        sigma-hat numerator, main.cpp:303-311.
//
// Everything that decides a bit of an fp32 weight uses __fadd_rn/__fsub_rn/__fmul_rn so that no FMA
// contraction can change the rounding the reference's SSE2 code performs.
#include "hml_common.cuh"
#include "hml_kernels.h"

namespace hml {

__constant__ float c_level_norm[64];  // c_level_norm[l] = normaliser of level l (1-based), running fp32 product

void upload_level_norms(const float* host64, cudaStream_t s) {
  cudaMemcpyToSymbolAsync(c_level_norm, host64, 64 * sizeof(float), 0, cudaMemcpyHostToDevice, s);
}

// One pass = 12 Haar levels over tiles of 4096 inputs.
//   in[0..n_valid)      complete inputs of this pass (pass 0: the data; pass p: complete tile sums of pass p-1)
//   n_pos               number of positions of this pass that exist in the sequence (ceil(T / stride))
//   coeffs[j * stride]  receives the coefficient of the wavelet whose mid discontinuity is input j,
//                       for every j < n_pos that is not a multiple of 4096 (those belong to later passes);
//                       incomplete wavelets get +inf (wavelet.hpp:140, never overwritten).
//   tile_sums[tile]     sum of the tile in the reference's tree order, for complete tiles.
__global__ void __launch_bounds__(256) k_maxlet_level(const float* __restrict__ in, uint64_t n_valid, uint64_t n_pos,
                                                      uint64_t stride, int level0, float* __restrict__ coeffs,
                                                      float* __restrict__ tile_sums) {
  __shared__ float bufA[kTile];
  __shared__ float bufB[kTile / 2];
  __shared__ float cout[kTile];
  const uint64_t base = (uint64_t)blockIdx.x * kTile;
  const float inf = __int_as_float(0x7f800000);
  for (int i = threadIdx.x; i < kTile; i += 256) {
    uint64_t g = base + i;
    bufA[i] = g < n_valid ? in[g] : 0.f;
    cout[i] = inf;
  }
  __syncthreads();
  float* cur = bufA;
  float* nxt = bufB;
#pragma unroll 1
  for (int l = 1; l <= kTileLog2; ++l) {
    const int nodes = kTile >> l;        // nodes at this level
    const uint64_t span = 1ull << l;     // inputs per node
    const float norm = c_level_norm[level0 + l];
    for (int p = threadIdx.x; p < nodes; p += 256) {
      const uint64_t right_end = base + (uint64_t)(p + 1) * span;
      if (right_end <= n_valid) {  // complete wavelet
        const float a = cur[2 * p], b = cur[2 * p + 1];
        cout[p * span + span / 2] = fmaxf(0.f, __fmul_rn(norm, fabsf(__fsub_rn(a, b))));
        nxt[p] = __fadd_rn(a, b);
      } else {
        nxt[p] = 0.f;
      }
    }
    __syncthreads();
    float* t = cur;
    cur = nxt;
    nxt = (t == bufA) ? bufA : bufB;  // ping-pong: level l sums live in the buffer level l-2 used
  }
  // cur[0] holds the tile sum (level 12) if the tile is complete
  if (threadIdx.x == 0 && base + kTile <= n_valid) tile_sums[blockIdx.x] = cur[0];
  for (int i = threadIdx.x; i < kTile; i += 256) {
    const uint64_t j = base + i;
    if (i != 0 && j < n_pos) coeffs[j * stride] = cout[i];
  }
}

// Multivariate input arrives position-major (wavelet.hpp:131-136); every dimension is transformed as its own plane.
__global__ void __launch_bounds__(256) k_deinterleave(const float* __restrict__ x, uint64_t T, int nr_dims, int dim,
                                                      float* __restrict__ plane) {
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < T; t += (uint64_t)gridDim.x * blockDim.x)
    plane[t] = x[t * nr_dims + dim];
}

// Maxlet coefficient = maximum over the dimensions of the normalised |detail| (wavelet.hpp:155-160; max is exact,
// incomplete wavelets are +inf in every dimension).
__global__ void __launch_bounds__(256) k_max_combine(float* __restrict__ dst, const float* __restrict__ src, uint64_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    dst[i] = fmaxf(dst[i], src[i]);
}

// HaarBreakpointWeights in closed form.  The reference walks the levels top-down and pushes each
// wavelet's |coefficient| to its left end L, its mid point and its right end R by max(), turning
// wavelets with R >= T into +inf (also at L).  A wavelet with mid point j has half width
// h = 2^ctz(j), so position t receives
//     c'[t],  c'[t - 2^k] (t is that wavelet's R)  and  c'[t + 2^k] (t is its L)   for all k < ctz(t),
// with c'[j] = (j + 2^ctz(j) < T) ? c[j] : +inf; position 0 is +inf.  max is exact, so the result is
// bit-identical to the sequential procedure.
__global__ void __launch_bounds__(256) k_bp_weights(const float* __restrict__ c, uint64_t T, float mult,
                                                    float* __restrict__ w) {
  const float inf = __int_as_float(0x7f800000);
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < T; t += (uint64_t)gridDim.x * blockDim.x) {
    float v;
    if (t == 0) {
      v = inf;
    } else {
      const int z = __ffsll((long long)t) - 1;  // ctz
      const uint64_t h = 1ull << z;
      v = (t + h < T) ? c[t] : inf;
      for (int k = 0; k < z; ++k) {
        const uint64_t d = 1ull << k;
        // wavelet with mid point t - d: half width d, right end t < T always
        const uint64_t jl = t - d;
        const float cl = (jl + d < T) ? c[jl] : inf;
        v = fmaxf(v, cl);
        const uint64_t jr = t + d;  // wavelet with mid point t + d: left end t
        if (jr < T) {
          const float cr = (jr + d < T) ? c[jr] : inf;
          v = fmaxf(v, cr);
        }
      }
    }
    w[t] = __fmul_rn(v, mult);
  }
}

// Segment mode.  Rank r holds the coefficients of its own observations; the closed form above reaches
// outside the segment only (a) at multiples of 4096, whose coefficients belong to Haar levels above 12
// and are derived by every rank from the all-gathered tile sums (ctop[m] = c[4096 m]), and (b) for the
// first position of the segment, at seg_start - 2^k, k < 12, which the previous rank sends (halo[k]).
__global__ void __launch_bounds__(256)
    k_bp_weights_segment(const float* __restrict__ c_local, const float* __restrict__ ctop,
                         const float* __restrict__ halo, uint64_t seg_start, uint64_t len, uint64_t T, float mult,
                         float* __restrict__ w) {
  const float inf = __int_as_float(0x7f800000);
  auto coef = [&](uint64_t j) -> float {  // c'[j]: +inf if the wavelet's right end is not inside the sequence
    const uint64_t h = 1ull << (__ffsll((long long)j) - 1);
    if (!(j + h < T)) return inf;
    if ((j & (uint64_t)(kTile - 1)) == 0) return ctop[j >> kTileLog2];
    if (j >= seg_start) return c_local[j - seg_start];
    return halo[__ffsll((long long)(seg_start - j)) - 1];
  };
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t t = seg_start + i;
    float v;
    if (t == 0) {
      v = inf;
    } else {
      const int z = __ffsll((long long)t) - 1;
      v = coef(t);
      for (int k = 0; k < z; ++k) {
        const uint64_t d = 1ull << k;
        v = fmaxf(v, coef(t - d));
        if (t + d < T) v = fmaxf(v, coef(t + d));
      }
    }
    w[i] = __fmul_rn(v, mult);
  }
}

__global__ void k_pack_edge(const float* __restrict__ c_local, uint64_t len, float* __restrict__ edge16) {
  const int k = threadIdx.x;
  if (k < 16) edge16[k] = (k < kTileLog2 && len >= (1ull << k)) ? c_local[len - (1ull << k)] : 0.f;
}

// sum of coeffs[1], coeffs[3], ... in fp64; per-CTA partials, summed on the host.
__global__ void __launch_bounds__(256) k_sum_odd(const float* __restrict__ c, uint64_t T, double* __restrict__ partial) {
  double s = 0.0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; 2 * i + 1 < T;
       i += (uint64_t)gridDim.x * blockDim.x)
    s += (double)c[2 * i + 1];
  __shared__ double sh[8];
  for (int o = 16; o > 0; o >>= 1) s += shfl_xor_double(s, o);
  if (lane_id() == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int i = 0; i < 8; ++i) t += sh[i];
    partial[blockIdx.x] = t;
  }
}

// Integral arrays.  One CTA per cell of 4096 observations; warp w owns 512 consecutive observations.
// pq[i] = (sum_{cell start <= j < i} x_j, same for x_j^2) for i in [0, T]; cell_tot[c] = totals.
__global__ void __launch_bounds__(256) k_integral_cells(const float* __restrict__ x, uint64_t T, double2* __restrict__ pq,
                                                        double2* __restrict__ cell_tot) {
  __shared__ double2 wtot[8];
  const uint64_t base = (uint64_t)blockIdx.x * kCell;
  const int warp = threadIdx.x >> 5, lane = lane_id();
  const uint64_t wbase = base + (uint64_t)warp * 512;
  // pass 1: warp totals
  double sx = 0, sq = 0;
#pragma unroll 4
  for (int it = 0; it < 16; ++it) {
    const uint64_t i = wbase + it * 32 + lane;
    const double d = i < T ? (double)x[i] : 0.0;
    sx += d;
    sq += d * d;  // exact: product of two fp32 values fits 48 bits
  }
  for (int o = 16; o > 0; o >>= 1) {
    sx += shfl_xor_double(sx, o);
    sq += shfl_xor_double(sq, o);
  }
  if (lane == 0) wtot[warp] = make_double2(sx, sq);
  __syncthreads();
  double cx = 0, cq = 0;
  for (int k = 0; k < warp; ++k) {
    cx += wtot[k].x;
    cq += wtot[k].y;
  }
  if (threadIdx.x == 0) {
    double tx = 0, tq = 0;
    for (int k = 0; k < 8; ++k) {
      tx += wtot[k].x;
      tq += wtot[k].y;
    }
    cell_tot[blockIdx.x] = make_double2(tx, tq);
  }
  // pass 2: exclusive running sums, 32 observations per step
#pragma unroll 1
  for (int it = 0; it < 16; ++it) {
    const uint64_t i = wbase + it * 32 + lane;
    const double d = i < T ? (double)x[i] : 0.0;
    double ix = d, iq = d * d;
    for (int o = 1; o < 32; o <<= 1) {
      const double ux = shfl_up_double(ix, o), uq = shfl_up_double(iq, o);
      if (lane >= o) {
        ix += ux;
        iq += uq;
      }
    }
    double ex = shfl_up_double(ix, 1), eq = shfl_up_double(iq, 1);
    if (lane == 0) ex = eq = 0.0;
    if (i <= T) pq[i] = make_double2(cx + ex, cq + eq);
    cx += shfl_double(ix, 31);
    cq += shfl_double(iq, 31);
  }
}

// ---------------------------------------------------------------- host-side launchers

void launch_maxlet_level(const float* in, uint64_t n_valid, uint64_t n_pos, uint64_t stride, int level0, float* coeffs,
                         float* tile_sums, cudaStream_t s) {
  const uint64_t tiles = (n_pos + kTile - 1) / kTile;
  if (tiles == 0) return;
  k_maxlet_level<<<(unsigned)tiles, 256, 0, s>>>(in, n_valid, n_pos, stride, level0, coeffs, tile_sums);
}
void launch_bp_weights(const float* c, uint64_t T, float mult, float* w, int sms, cudaStream_t s) {
  uint64_t blocks = (T + 255) / 256;
  if (blocks > (uint64_t)sms * 32) blocks = (uint64_t)sms * 32;
  k_bp_weights<<<(unsigned)blocks, 256, 0, s>>>(c, T, mult, w);
}
void launch_pack_edge(const float* coeffs_local, uint64_t len, float* edge16, cudaStream_t s) {
  k_pack_edge<<<1, 32, 0, s>>>(coeffs_local, len, edge16);
}
void launch_bp_weights_segment(const float* c_local, const float* ctop, const float* halo, uint64_t seg_start,
                               uint64_t len, uint64_t T, float mult, float* w, int sms, cudaStream_t s) {
  uint64_t blocks = (len + 255) / 256;
  if (blocks > (uint64_t)sms * 32) blocks = (uint64_t)sms * 32;
  if (blocks == 0) return;
  k_bp_weights_segment<<<(unsigned)blocks, 256, 0, s>>>(c_local, ctop, halo, seg_start, len, T, mult, w);
}
void launch_sum_odd(const float* c, uint64_t T, double* partial, int nblocks, cudaStream_t s) {
  k_sum_odd<<<nblocks, 256, 0, s>>>(c, T, partial);
}
void launch_deinterleave(const float* x, uint64_t T, int nr_dims, int dim, float* plane, int sms, cudaStream_t s) {
  uint64_t blocks = (T + 255) / 256;
  if (blocks > (uint64_t)sms * 32) blocks = (uint64_t)sms * 32;
  if (blocks == 0) return;
  k_deinterleave<<<(unsigned)blocks, 256, 0, s>>>(x, T, nr_dims, dim, plane);
}
void launch_max_combine(float* dst, const float* src, uint64_t n, int sms, cudaStream_t s) {
  uint64_t blocks = (n + 255) / 256;
  if (blocks > (uint64_t)sms * 32) blocks = (uint64_t)sms * 32;
  if (blocks == 0) return;
  k_max_combine<<<(unsigned)blocks, 256, 0, s>>>(dst, src, n);
}
void launch_integral_cells(const float* x, uint64_t T, double2* pq, double2* cell_tot, cudaStream_t s) {
  const uint64_t cells = T / kCell + 1;  // covers index T
  k_integral_cells<<<(unsigned)cells, 256, 0, s>>>(x, T, pq, cell_tot);
}

}  // namespace hml
