// hammlet_b200 — the whole Gibbs sweep in ONE persistent kernel, for block structures of up to kFusedMaxTiles tiles
// (65 536 blocks) and K <= 8, with the conjugate updates and the parameter draws on the device.
//
// Reference path: one iteration of sampleHMM, HMM.hpp:99-121 — createBlocks(theta), StateSequence<ForwardBackward>::
// sample (StateSequence/ForwardBackward.hpp:16-213), theta.sample, pi.sample, A.sample (Theta.hpp:203-211,
// Initial.hpp:35-40, Transitions.hpp:75-79 with Conjugate.hpp:120-205 and Distribution.hpp:76-139).
//
// Why: the multi-kernel sweep (hml_sweep_impl.cuh) is 13 dependent launches and one host round trip per sweep.  At
// 1e9 observations that overhead hides behind 240 us of work; at 1e6 (1.3 k blocks) or 1e7 observations (15 k blocks)
// every kernel sits on its 7-17 us floor and the sweep costs 65-85 us of which the GPU works a fraction.  Here a
// cooperative grid of one CTA per tile (at most one per SM) keeps everything of a tile — emission terms, chunk
// operators, forward rows, backward maps, states — in that CTA, the phases of a sweep are separated by six grid-wide
// barriers instead of kernel boundaries, and the kernel loops over sweeps: the model of sweep i+1 is drawn by CTA 0
// from the statistics of sweep i (Philox streams) while the other CTAs wait at the barrier, so nothing returns to the
// host until the requested sweeps are done or the kernel stands down (the threshold left the candidate list, the block
// arrays are too small, a forward sum vanished: the host then handles that one sweep through the multi-kernel path).
//
// Per-block arrays, layouts and numerics are those of the multi-kernel path (same Layout::perm, same helpers), so
// every getter of the C ABI (states, runs, marginals, blocks) works on the result, and under uniform replay the
// sampled states are the reference's.
#pragma once
#include <cooperative_groups.h>

#include "hml_sweep_impl.cuh"

namespace hml {

namespace cg = cooperative_groups;

// ---- Philox-driven draws for the parameter phase: every draw owns a stream, successive uniforms advance its counter
struct DrawStream {
  unsigned long long seed, sweep;
  uint32_t stream;
  unsigned long long ctr;
  __device__ __forceinline__ double uniform() { return Philox::uniform(seed, sweep, stream, ctr++); }
  __device__ __forceinline__ double normal() {  // Box-Muller, one value per two uniforms
    const double u1 = 1.0 - uniform(), u2 = uniform();
    return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
  }
  // Gamma(alpha, 1): Marsaglia & Tsang 2000 (the method libstdc++'s gamma_distribution uses), boosted for alpha < 1
  __device__ double gamma(double alpha) {
    const double a = alpha < 1.0 ? alpha + 1.0 : alpha;
    const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    double g = 0.0;
    for (int tries = 0; tries < 1000; ++tries) {
      const double x = normal();
      double v = 1.0 + c * x;
      if (v <= 0.0) continue;
      v = v * v * v;
      const double u = 1.0 - uniform();
      if (log(u) < 0.5 * x * x + d - d * v + d * log(v)) {
        g = d * v;
        break;
      }
    }
    if (alpha < 1.0) g *= pow(1.0 - uniform(), 1.0 / alpha);
    return g;
  }
};

// theta, pi, A of the next sweep from the statistics of this one, by the first 64 + K*K threads of one CTA.
// Conjugate.hpp:120-205 in real_t = float like the reference; posterior := prior after every draw (Theta.hpp:209).
// s_g: K*K + 2*K floats of shared memory.
template <int KP>
__device__ void chain_sample_params(ChainDev* ch, const unsigned long long* out_u64, const double* out_f64, float* s_g) {
  const int K = ch->K, tid = threadIdx.x;
  DrawStream rs{ch->seed, ch->sweep, 0u, 0ull};
  float* g_pi = s_g;               // K
  float* g_A = s_g + KP;           // K*K
  float* s_var = s_g + KP + KP * KP;  // K
  if (tid < K) {                   // ---- theta_k: NIG posterior, then var = 1 / Gamma(alpha, 1 / beta), mean ~ N(mu0, sqrt(var / nu))
    float alpha = ch->prior_theta[tid][0], beta = ch->prior_theta[tid][1], mu0 = ch->prior_theta[tid][2], nu = ch->prior_theta[tid][3];
    const unsigned long long n = __ldcg(out_u64 + tid);
    if (n > 0) {
      const float sum = (float)__ldcg(out_f64 + tid), sumSq = (float)__ldcg(out_f64 + KP + tid);
      const double N = (double)n;
      const float xbar = (float)(sum / N);
      float ssN = (float)((sum * sum) / N);
      if (ssN > sumSq) ssN = sumSq;
      const float a2 = (float)(alpha + N / 2.0);
      const float b2 = (float)(beta + ((sumSq + (N * nu / (N + nu)) * ((xbar - mu0) * (xbar - mu0))) - ssN) / 2.0);
      const float m2 = (float)((nu * mu0 + sum) / (nu + N));
      const float n2 = (float)(nu + N);
      alpha = a2; beta = b2; mu0 = m2; nu = n2;
    }
    rs.stream = 2u + (uint32_t)tid;
    const float gam = (float)(rs.gamma((double)alpha) / (double)beta);
    const float var = 1.0f / gam;
    const float mean = (float)((double)mu0 + sqrt((double)(var / nu)) * rs.normal());
    if (!(var > 0.0f) || !isfinite(var) || !isfinite(mean)) ch->phase_abort[6] = kChainNumeric;  // Observation.hpp:177-179
    ch->mean[tid] = (double)mean;
    ch->var[tid] = (double)var;
    s_var[tid] = var;
  }
  if (tid >= 32 && tid < 32 + K) {  // ---- pi ~ Dirichlet(alpha_I + occupancy) (FB.hpp:211: the quirk of A11)
    const int s = tid - 32;
    rs.stream = 2u + (uint32_t)KP + (uint32_t)s;
    g_pi[s] = (float)rs.gamma((double)(ch->prior_pi + (float)__ldcg(out_u64 + s)));
  }
  if (tid >= 64 && tid < 64 + K * K) {  // ---- rows of A ~ Dirichlet(prior row + transition counts)
    const int i = (tid - 64) / K, j = (tid - 64) % K;
    rs.stream = 2u + 2u * (uint32_t)KP + (uint32_t)(i * KP + j);
    const float prior = i == j ? ch->prior_self : ch->prior_trans;
    g_A[i * KP + j] = (float)rs.gamma((double)(prior + (float)__ldcg(out_u64 + KP + i * KP + j)));
  }
  __syncthreads();
  if (tid < K) {
    float s = 0.f;
    for (int j = 0; j < K; ++j) s += g_A[tid * KP + j];
    for (int j = 0; j < K; ++j) ch->A[tid * K + j] = (double)(g_A[tid * KP + j] / s);
  }
  if (tid == 32) {
    float s = 0.f;
    for (int j = 0; j < K; ++j) s += g_pi[j];
    for (int j = 0; j < K; ++j) ch->pi[j] = (double)(g_pi[j] / s);
    float mv = s_var[0];
    for (int j = 1; j < K; ++j) mv = fminf(mv, s_var[j]);
    ch->thr = sqrtf(2.0f * logf((float)ch->T) * mv);  // BreakpointArray.hpp:195-199, Theta.hpp:226-234
    ch->sweep += 1ull;
  }
  __syncthreads();
}

// stand-alone: the parameter phase after a sweep of the multi-kernel path
template <int KP>
__global__ void __launch_bounds__(kFusedThreads) k_chain_params(ChainDev* ch, const unsigned long long* out_u64, const double* out_f64) {
  __shared__ float s_g[KP * KP + 2 * KP];
  chain_sample_params<KP>(ch, out_u64, out_f64, s_g);
}

// What other CTAs of the same launch wrote is read from L2 (ld.global.cg): the L1 of this SM may still hold the line as
// it was a sweep ago.
template <int KP>
__device__ __forceinline__ void load_op_cg(OpVals<KP>& o, const double* M, const int* X) {
#pragma unroll
  for (int k = 0; k < KP * KP; ++k) o.m[k] = __ldcg(M + k);
#pragma unroll
  for (int k = 0; k < KP; ++k) o.x[k] = __ldcg(X + k);
}
template <int KP>
__device__ __forceinline__ Map<KP> load_map_cg(const uint8_t* p) {
  Map<KP> r;
  const unsigned long long* q = reinterpret_cast<const unsigned long long*>(p);
#pragma unroll
  for (int i = 0; i < Map<KP>::W; ++i) r.w[i] = __ldcg(q + i);
  return r;
}

// thread i < KP fills row i of the model (mirrors make_model)
template <int KP>
__device__ __forceinline__ void model_from_chain(const ChainDev* ch, ModelDev<KP>& m, int i) {
  const int K = ch->K;
  if (i == 0) {
    m.K = K;
    m.use_self = ch->use_self;
  }
  const bool on = i < K;
  const double mu = on ? __ldcg(&ch->mean[i]) : 0.0, var = on ? __ldcg(&ch->var[i]) : 1.0;
  m.mean[i] = mu;
  m.inv2var[i] = on ? 1.0 / (2.0 * var) : 0.0;
  m.lognorm[i] = on ? log(sqrt(var)) + mu * mu / (2 * var) : 0.0;
  m.loga[i] = (on && ch->use_self) ? log(__ldcg(&ch->A[i * K + i])) : 0.0;
  m.pi[i] = on ? __ldcg(&ch->pi[i]) : 0.0;
  for (int j = 0; j < KP; ++j) m.A[i][j] = (on && j < K) ? __ldcg(&ch->A[i * K + j]) : 0.0;
}

// as block_sums, with the block list read from L2 (the scatter phase of other CTAs wrote it)
__device__ __forceinline__ void block_sums_cg(const SweepBuffers& buf, uint64_t b, uint32_t& n, double& sx, double& sq) {
  const uint32_t s = __ldcg(buf.starts + b), e = __ldcg(buf.starts + b + 1);
  const double2 ps = __ldcg(buf.spq + b), pe = __ldcg(buf.spq + b + 1);
  range_sums_from(buf, ps, pe, s, e, sx, sq);
  n = e - s;
}

// ---------------------------------------------------------------------------------------------------------------
template <int KP>
__global__ void __launch_bounds__(kFusedThreads, 1) k_sweep_fused(SweepBuffers buf, FusedArgs args) {
  constexpr int L = Layout::L, C = Layout::C, TB = Layout::TB, MB = 8 * Map<KP>::W;
  constexpr int PE = 33 * KP, PB = 36;
  cg::grid_group grid = cg::this_grid();
  __shared__ ModelDev<KP> m;
  __shared__ double s_tab[64];
  __shared__ double s_ops[C * KP * KP];
  __shared__ int s_exp[C * KP];
  __shared__ double s_e[8 * PE];
  __shared__ double s_sx[8 * PB], s_sq[8 * PB];
  __shared__ uint32_t s_n[8 * PB];
  __shared__ double s_red[kFusedThreads / 32][2 * KP];
  __shared__ unsigned long long s_cnt[KP * KP + KP];
  __shared__ float s_g[KP * KP + 2 * KP];
  __shared__ uint32_t s_warp[8];
  __shared__ uint32_t s_misc[4];
  __shared__ double s_ain[KP];
  ChainDev* ch = args.chain;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t G = gridDim.x, cta = blockIdx.x;
  exp_table_load(s_tab);
  const uint64_t capacity = buf.capacity;
  unsigned long long* nb_out = const_cast<unsigned long long*>(buf.nblocks);
  uint32_t* starts = const_cast<uint32_t*>(buf.starts);
  double2* spq = const_cast<double2*>(buf.spq);

  for (int it = 0; it < args.nsweeps; ++it) {
    // ---------------- model of this sweep (written by CTA 0 before the barrier that ended the previous one)
    if (tid < KP) model_from_chain<KP>(ch, m, tid);
    __syncthreads();
    const float thr = __ldcg(&ch->thr);
    const int K = m.K;
    const unsigned long long sweep_key = args.philox_sweep_from_chain ? __ldcg(&ch->sweep) : args.sweep;
    if (cta == 0) {
      // the first phase that writes the result block of this sweep
      for (int i = tid; i < KP + KP * KP + 1; i += kFusedThreads) buf.out_u64[i] = 0;
      for (int i = tid; i < 2 * KP + 1; i += kFusedThreads) buf.out_f64[i] = 0.0;
      if (tid == 0 && !(thr >= ch->cand_floor && thr > 0.f && isfinite(thr))) ch->phase_abort[1] = kChainThreshold;
    }

    // ---------------- boundaries among the candidates: count
    const uint32_t per = ((args.nc + G - 1) / G + 255u) / 256u * 256u;  // slice of this CTA, whole chunks of 256
    const uint32_t lo = min(args.nc, cta * per), hi = min(args.nc, lo + per);
    {
      uint32_t c = 0;
      for (uint32_t i = lo + tid; i < hi; i += kFusedThreads) c += !(args.cand_w[i] < thr) ? 1u : 0u;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
      if (lane == 0) s_warp[warp] = c;
      __syncthreads();
      if (tid == 0) {
        uint32_t t = 0;
        for (int i = 0; i < 8; ++i) t += s_warp[i];
        args.cta_count[cta] = t;
      }
    }
    grid.sync();  // (1) counts of all slices
    if (const unsigned code = __ldcg(&ch->phase_abort[1])) {
      if (cta == 0 && tid == 0) ch->abort_code = code;
      break;
    }
    // ---------------- ... and scatter: block starts and their integral pairs, in order
    if (warp == 0) {
      uint32_t before = 0, total = 0;
      for (uint32_t c0 = 0; c0 < G; c0 += 32) {
        const uint32_t v = (c0 + lane < G) ? __ldcg(args.cta_count + c0 + lane) : 0u;
        uint32_t b = (c0 + lane < cta) ? v : 0u, t = v;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          b += __shfl_xor_sync(0xffffffffu, b, o);
          t += __shfl_xor_sync(0xffffffffu, t, o);
        }
        before += b;
        total += t;
      }
      if (lane == 0) {
        s_misc[0] = before;
        s_misc[1] = total;
      }
    }
    __syncthreads();
    const uint64_t B = s_misc[1];
    if (cta == 0 && tid == 0) *nb_out = B;
    if (B > capacity || B > (uint64_t)kFusedMaxTiles * TB) {
      if (cta == 0 && tid == 0) {
        ch->nblocks_seen = B;
        ch->phase_abort[2] = kChainCapacity;
      }
    } else {
      uint32_t run = s_misc[0];
      for (uint32_t base = lo; base < hi; base += kFusedThreads) {
        const uint32_t i = base + tid;
        const bool f = i < hi && !(args.cand_w[i] < thr);
        const unsigned bal = __ballot_sync(0xffffffffu, f);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        uint32_t bw = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
          const uint32_t v = s_warp[w];
          if (w < warp) bw += v;
          tot += v;
        }
        if (f) {
          const uint32_t o = run + bw + __popc(bal & ((1u << lane) - 1u));
          starts[o] = args.cand_pos[i];
          spq[o] = args.cand_pq[i];
        }
        run += tot;
        __syncthreads();
      }
      if (cta == 0 && tid == 0) {
        starts[B] = args.T_local;  // sentinel: block b = [starts[b], starts[b + 1])
        spq[B] = buf.pq[args.T_local];
      }
    }
    grid.sync();  // (2) the block list
    if (const unsigned code = __ldcg(&ch->phase_abort[2])) {
      if (cta == 0 && tid == 0) ch->abort_code = code;
      break;
    }
    const uint32_t ntiles = (uint32_t)((B + TB - 1) / TB);

    // ---------------- per owned tile: block statistics, emission terms, chunk operators, their scan inside the tile
    for (uint32_t tile = cta; tile < ntiles; tile += G) {
      for (int g4 = 0; g4 < 4; ++g4) {
        const uint64_t b = (uint64_t)tile * TB + g4 * 256 + tid;
        const int cl = tid >> 5, t = tid & 31, c0 = g4 * 8;
        const int oc = tid & 7, ot = tid >> 3;
        const bool valid = b < B;
        uint32_t n = 0;
        double sx = 0.0, sq = 0.0;
        if (valid) block_sums_cg(buf, b, n, sx, sq);
        s_n[cl * PB + t] = n;
        s_sx[cl * PB + t] = sx;
        s_sq[cl * PB + t] = sq;
        double E[KP];
        const double mx = emission_terms<KP, false>(m, (double)n, sx, sq, E);
#pragma unroll
        for (int s = 0; s < KP; ++s) s_e[cl * PE + t * KP + s] = (valid && s < K) ? exp_nonpos(E[s] - mx, s_tab) : 0.0;
        __syncthreads();
        const uint64_t ob = (uint64_t)tile * TB + (uint64_t)(c0 + oc) * L + ot;
        const uint64_t op = Layout::at(tile, c0 + oc, ot);
        if (ob < B) {
          buf.bN[op] = s_n[oc * PB + ot];
          buf.bS[op] = make_double2(s_sx[oc * PB + ot], s_sq[oc * PB + ot]);
        }
#pragma unroll
        for (int k = 0; k < KP; ++k) {
          const int q = k * 256 + tid;
          const int qt = q / (8 * KP), within = q % (8 * KP);
          buf.e[(Layout::at(tile, c0, qt)) * KP + within] = s_e[(within / KP) * PE + qt * KP + within % KP];
        }
        __syncthreads();
      }
      // chunk operators: thread (chunk, row), row recursion over the 32 blocks of the chunk (as k_fwd_chunks_prefix)
      {
        const int c = tid / KP, i = tid % KP;
        if (c < C) {
          const uint64_t first = (uint64_t)tile * TB + (uint64_t)c * L;
          int steps = 0;
          if (first < B) steps = (B - first) < (uint64_t)L ? (int)(B - first) : L;
          double r[KP];
          int rex = 0;
#pragma unroll
          for (int j = 0; j < KP; ++j) r[j] = (j == i) ? 1.0 : 0.0;
          const double* ep = buf.e + Layout::at(tile, c, 0) * KP;
#pragma unroll 1
          for (int t0 = 0; t0 < steps; t0 += 4) {
            double ev[4][KP];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
#pragma unroll
              for (int j = 0; j < KP; ++j) ev[q][j] = (t0 + q < steps) ? __ldcg(ep + (uint64_t)(t0 + q) * C * KP + j) : 0.0;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (t0 + q < steps) {
                double y[KP];
#pragma unroll
                for (int j = 0; j < KP; ++j) y[j] = 0.0;
#pragma unroll
                for (int k = 0; k < KP; ++k) {
#pragma unroll
                  for (int j = 0; j < KP; ++j) y[j] = fma(r[k], m.A[k][j], y[j]);
                }
#pragma unroll
                for (int j = 0; j < KP; ++j) r[j] = y[j] * ev[q][j];
              }
            }
            if (rex != kDeadExp) renorm_pow2<KP>(r, rex);
          }
#pragma unroll
          for (int j = 0; j < KP; ++j) s_ops[(c * KP + i) * KP + j] = r[j];
          s_exp[c * KP + i] = rex;
        }
        __syncthreads();
        const int pr = tid / KP, row = tid % KP;
        for (int stride = 1; stride < C; stride <<= 1) {  // up-sweep
          const int n2 = (pr + 1) * 2 * stride - 1;
          const bool on = n2 < C;
          double r[KP];
          int rex = 0;
          if (on) {
#pragma unroll
            for (int j = 0; j < KP; ++j) r[j] = s_ops[((n2 - stride) * KP + row) * KP + j];
            rex = s_exp[(n2 - stride) * KP + row];
            row_times_op<KP, false>(r, rex, s_ops + n2 * KP * KP, s_exp + n2 * KP);
          }
          __syncthreads();
          if (on) {
#pragma unroll
            for (int j = 0; j < KP; ++j) s_ops[(n2 * KP + row) * KP + j] = r[j];
            s_exp[n2 * KP + row] = rex;
          }
          __syncthreads();
        }
        if (tid < KP) {
#pragma unroll
          for (int j = 0; j < KP; ++j) {
            buf.tile_ops[((uint64_t)tile * KP + tid) * KP + j] = s_ops[((C - 1) * KP + tid) * KP + j];
            s_ops[((C - 1) * KP + tid) * KP + j] = (j == tid) ? 1.0 : 0.0;
          }
          buf.tile_exp[(uint64_t)tile * KP + tid] = s_exp[(C - 1) * KP + tid];
          s_exp[(C - 1) * KP + tid] = 0;
        }
        __syncthreads();
        for (int stride = C / 2; stride >= 1; stride >>= 1) {  // down-sweep
          const int n2 = (pr + 1) * 2 * stride - 1;
          const bool on = n2 < C;
          double pfx[KP], r[KP];
          int pex = 0, rex = 0;
          if (on) {
#pragma unroll
            for (int j = 0; j < KP; ++j) pfx[j] = r[j] = s_ops[(n2 * KP + row) * KP + j];
            pex = rex = s_exp[n2 * KP + row];
            row_times_op<KP, false>(r, rex, s_ops + (n2 - stride) * KP * KP, s_exp + (n2 - stride) * KP);
          }
          __syncthreads();
          if (on) {
#pragma unroll
            for (int j = 0; j < KP; ++j) {
              s_ops[((n2 - stride) * KP + row) * KP + j] = pfx[j];
              s_ops[(n2 * KP + row) * KP + j] = r[j];
            }
            s_exp[(n2 - stride) * KP + row] = pex;
            s_exp[n2 * KP + row] = rex;
          }
          __syncthreads();
        }
        for (int k = tid; k < C * KP * KP; k += kFusedThreads) buf.chunk_ops[(uint64_t)tile * C * KP * KP + k] = s_ops[k];
        for (int k = tid; k < C * KP; k += kFusedThreads) buf.chunk_exp[(uint64_t)tile * C * KP + k] = s_exp[k];
        __syncthreads();
      }
    }
    grid.sync();  // (3) tile operators

    // ---------------- forward: vector entering each owned tile, rows of the tile, backward maps, chunk and tile maps
    unsigned fallbacks = 0;
    for (uint32_t tile = cta; tile < ntiles; tile += G) {
      if (warp == 0) {
        // every CTA walks the tile operators in front of its tile itself (at most 63 vector-operator products)
        double a[KP];
#pragma unroll
        for (int j = 0; j < KP; ++j) a[j] = m.pi[j];
        OpVals<KP> cur, nxt;
        if (tile > 0) load_op_cg<KP>(cur, buf.tile_ops, buf.tile_exp);
        for (uint32_t t = 0; t < tile; ++t) {
          if (t + 1 < tile) load_op_cg<KP>(nxt, buf.tile_ops + (uint64_t)(t + 1) * KP * KP, buf.tile_exp + (uint64_t)(t + 1) * KP);
          if (!vec_apply_op<KP>(a, cur)) fallbacks++;
          cur = nxt;
        }
        // ---- rows: lane c owns chunk c (as k_fwd_replay_prefix, power-of-two rescaling instead of the division)
        const int c = lane;
        const uint64_t first = (uint64_t)tile * TB + (uint64_t)c * L;
        int steps = 0;
        if (first < B) steps = (B - first) < (uint64_t)L ? (int)(B - first) : L;
        if (c > 0 && steps > 0) {
          OpVals<KP> o;
          load_op<KP>(o, buf.chunk_ops + ((uint64_t)tile * C + c) * KP * KP, buf.chunk_exp + ((uint64_t)tile * C + c) * KP);
          if (!vec_apply_op<KP>(a, o)) fallbacks++;
        }
        const double* ep = buf.e + Layout::at(tile, c, 0) * KP;
#pragma unroll 1
        for (int t0 = 0; t0 < steps; t0 += 4) {
          double ev[4][KP];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
#pragma unroll
            for (int j = 0; j < KP; ++j) ev[q][j] = (t0 + q < steps) ? __ldcg(ep + (uint64_t)(t0 + q) * C * KP + j) : 0.0;
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (t0 + q < steps) {
              double f[KP];
#pragma unroll
              for (int j = 0; j < KP; ++j) f[j] = 0.0;
#pragma unroll
              for (int k = 0; k < KP; ++k) {
#pragma unroll
                for (int j = 0; j < KP; ++j) f[j] = fma(a[k], m.A[k][j], f[j]);
              }
              double mxv = 0.0;
#pragma unroll
              for (int j = 0; j < KP; ++j) {
                f[j] *= ev[q][j];
                mxv = fmax(mxv, f[j]);
              }
              if (mxv > 0.0) {
                int e2 = exponent_of(mxv);
                if (e2 < -1000) e2 = -1000;
                const double sc = pow2i(-e2);
#pragma unroll
                for (int j = 0; j < KP; ++j) a[j] = f[j] * sc;
              } else {  // FB.hpp:106-111: the uniform fallback is not an operator product; the host re-runs the sweep
                fallbacks++;
#pragma unroll
                for (int j = 0; j < KP; ++j) a[j] = (j < K) ? 1.0 / (double)K : 0.0;
              }
              const uint64_t p = Layout::at(tile, c, t0 + q);
#pragma unroll
              for (int j = 0; j < KP; ++j) buf.alpha[p * KP + j] = a[j];
            }
          }
        }
      }
      __syncthreads();
      // ---- backward maps of the tile's blocks (as k_bwd_maps)
      for (int g4 = 0; g4 < 4; ++g4) {
        const uint64_t p = (uint64_t)tile * TB + g4 * 256 + tid;
        const uint64_t b = Layout::inv(p);
        Map<KP> fm = Map<KP>::identity();
        if (b < B) {
          const bool last = b + 1 == B;
          const double Nm1 = (double)buf.bN[p] - 1.0;
          double ap[KP];
#pragma unroll
          for (int j = 0; j < KP; ++j) {
            const double al = buf.alpha[p * KP + j];
            ap[j] = (last || !m.use_self || j >= K) ? al : al * exp_nonpos(Nm1 * m.loga[j], s_tab);
          }
          const double u = buf.replay_u ? buf.replay_u[B - 1 - b] : Philox::uniform(args.seed, sweep_key, 0u, b);
          fm = Map<KP>::zero();
          if (last) {
            const uint32_t q = discrete_draw<KP>(ap, K, u);
#pragma unroll
            for (int j = 0; j < KP; ++j) fm.set(j, q);
          } else {
#pragma unroll
            for (int j = 0; j < KP; ++j) {
              if (j < K) {
                double cs[KP];
                cs[0] = ap[0] * m.A[0][j];
#pragma unroll
                for (int k = 1; k < KP; ++k) cs[k] = fma(ap[k], m.A[k][j], cs[k - 1]);
                bool tie;
                uint32_t q = discrete_draw_fast<KP>(cs, K, u, tie);
                if (tie) {
                  Weights<KP> w;
#pragma unroll
                  for (int k = 0; k < KP; ++k) w.v[k] = __dmul_rn(ap[k], m.A[k][j]);
                  q = discrete_draw_exact<KP>(w, K, u);
                }
                fm.set(j, q);
              }
            }
          }
        }
        fm.store(buf.maps + p * MB);
      }
      __syncthreads();
      if (warp == 0) {  // chunk maps and the tile map (as k_bwd_chunkmaps)
        Map<KP> Gm = Map<KP>::identity();
#pragma unroll 4
        for (int t = 0; t < L; ++t) Gm = Gm.after(Map<KP>::load(buf.maps + Layout::at(tile, lane, t) * MB));
        Map<KP> inc = Gm;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          Map<KP> other;
#pragma unroll
          for (int i = 0; i < Map<KP>::W; ++i) other.w[i] = __shfl_down_sync(0xffffffffu, inc.w[i], o);
          if (lane + o < 32) inc = inc.after(other);
        }
        Map<KP> excl;
#pragma unroll
        for (int i = 0; i < Map<KP>::W; ++i) excl.w[i] = __shfl_down_sync(0xffffffffu, inc.w[i], 1);
        if (lane == 31) excl = Map<KP>::identity();
        excl.store(buf.chunk_maps + ((uint64_t)tile * C + lane) * MB);
        if (lane == 0) inc.store(buf.tile_maps + (uint64_t)tile * MB);
      }
      __syncthreads();
    }
    if (fallbacks) ch->phase_abort[4] = kChainFallback;
    grid.sync();  // (4) tile maps
    if (const unsigned code = __ldcg(&ch->phase_abort[4])) {
      if (cta == 0 && tid == 0) ch->abort_code = code;
      break;
    }

    // ---------------- backward: state following each owned tile, states of the tile, statistics
    double ax[KP], aq[KP];
    unsigned long long an[KP], ad[KP];
#pragma unroll
    for (int s = 0; s < KP; ++s) {
      ax[s] = aq[s] = 0.0;
      an[s] = ad[s] = 0;
    }
    for (int i = tid; i < KP * KP + KP; i += kFusedThreads) s_cnt[i] = 0;
    __syncthreads();
    for (uint32_t tile = cta; tile < ntiles; tile += G) {
      if (tid == 0) {
        uint32_t q = 0;  // the last block of the sequence carries a constant map: the start value is irrelevant
        for (uint32_t t = ntiles - 1; t > tile; --t) q = load_map_cg<KP>(buf.tile_maps + (uint64_t)t * MB).get(q);
        s_misc[2] = q;
        buf.tile_qin[tile] = (uint8_t)q;
      }
      __syncthreads();
      const uint32_t qin = s_misc[2];
      if (tid < C) {  // states of the tile's chunks (as k_bwd_replay)
        const int c = tid;
        const uint64_t first = (uint64_t)tile * TB + (uint64_t)c * L;
        if (first < B) {
          const int steps = (B - first) < (uint64_t)L ? (int)(B - first) : L;
          uint32_t q = Map<KP>::load(buf.chunk_maps + ((uint64_t)tile * C + c) * MB).get(qin);
          for (int t0 = steps - 1; t0 >= 0; t0 -= 8) {
            Map<KP> mp[8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (t0 - i >= 0) mp[i] = Map<KP>::load(buf.maps + Layout::at(tile, c, t0 - i) * MB);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (t0 - i >= 0) {
                q = mp[i].get(q);
                buf.states[Layout::at(tile, c, t0 - i)] = (uint8_t)q;
              }
            }
          }
        }
      }
      __syncthreads();
      // statistics of the tile's blocks (as k_reduce_partial; the successor of the tile's last block is qin)
      for (int g4 = 0; g4 < 4; ++g4) {
        const uint64_t p = (uint64_t)tile * TB + g4 * 256 + tid;
        const uint64_t b = Layout::inv(p);
        if (b >= B) continue;
        const uint32_t st = buf.states[p];
        const bool has_next = b + 1 < B;
        uint32_t next = st;
        if (has_next) next = ((b + 1) % TB == 0) ? qin : (uint32_t)buf.states[Layout::perm(b + 1)];
        const uint32_t n = buf.bN[p];
        const double2 v = buf.bS[p];
        unsigned long long dg = (unsigned long long)(n - 1) + ((has_next && next == st) ? 1ull : 0ull);
        if (b == 0) {  // the phantom transition 0 -> q_0 (FB.hpp:177,182-184)
          if (st == 0u)
            dg += 1ull;
          else
            atomicAdd(&s_cnt[st], 1ull);
        }
#pragma unroll
        for (int s = 0; s < KP; ++s) {
          const bool hit = st == (uint32_t)s;
          ax[s] += hit ? v.x : 0.0;
          aq[s] += hit ? v.y : 0.0;
          an[s] += hit ? (unsigned long long)n : 0ull;
          ad[s] += hit ? dg : 0ull;
        }
        if (has_next && next != st) atomicAdd(&s_cnt[st * KP + next], 1ull);
      }
    }
#pragma unroll
    for (int s = 0; s < KP; ++s) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        ax[s] += shfl_xor_double(ax[s], o);
        aq[s] += shfl_xor_double(aq[s], o);
        an[s] += __shfl_xor_sync(0xffffffffu, an[s], o);
        ad[s] += __shfl_xor_sync(0xffffffffu, ad[s], o);
      }
    }
    if (lane == 0) {
#pragma unroll
      for (int s = 0; s < KP; ++s) {
        s_red[warp][s] = ax[s];
        s_red[warp][KP + s] = aq[s];
        if (an[s]) atomicAdd(&s_cnt[KP * KP + s], an[s]);
        if (ad[s]) atomicAdd(&s_cnt[s * KP + s], ad[s]);
      }
    }
    __syncthreads();
    if (tid < 2 * KP) {
      double t = 0.0;
      for (int w = 0; w < kFusedThreads / 32; ++w) t += s_red[w][tid];
      buf.partials[(size_t)cta * 2 * KP + tid] = t;
    }
    for (int i = tid; i < KP * KP; i += kFusedThreads)
      if (s_cnt[i]) atomicAdd(&buf.out_u64[KP + i], s_cnt[i]);
    for (int i = tid; i < KP; i += kFusedThreads)
      if (s_cnt[KP * KP + i]) atomicAdd(&buf.out_u64[i], s_cnt[KP * KP + i]);
    grid.sync();  // (5) partial statistics
    if (cta == 0) {
      // final sums in a fixed order (deterministic), then the parameters of the next sweep
      if (tid < 2 * KP) {
        double t = 0.0;
        for (uint32_t c2 = 0; c2 < G; ++c2) t += __ldcg(buf.partials + (size_t)c2 * 2 * KP + tid);
        buf.out_f64[tid] = t;
      }
      __syncthreads();
      if (tid == 0) ch->sweeps_done += 1u;
      if (args.sample_params) {
        __threadfence();
        chain_sample_params<KP>(ch, buf.out_u64, buf.out_f64, s_g);
      }
      __threadfence();
    }
    grid.sync();  // (6) the model of the next sweep
    if (const unsigned code = __ldcg(&ch->phase_abort[6])) {
      if (cta == 0 && tid == 0) ch->abort_code = code;
      break;
    }
  }
}

// one cooperative launch of `a.nsweeps` sweeps; returns the cudaError_t of the launch
template <int KP>
int fused_impl(const SweepBuffers& b, const FusedArgs& a, int grid, cudaStream_t s) {
  SweepBuffers bb = b;
  FusedArgs aa = a;
  void* params[] = {(void*)&bb, (void*)&aa};
  return (int)cudaLaunchCooperativeKernel((const void*)k_sweep_fused<KP>, dim3(grid), dim3(kFusedThreads), params, 0, s);
}
// CTAs of k_sweep_fused<KP> that can be resident at once on the current device (0: the kernel does not fit)
template <int KP>
int fused_max_grid_impl(int sms) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sweep_fused<KP>, kFusedThreads, 0) != cudaSuccess) return 0;
  return per_sm > 0 ? sms : 0;  // one CTA per SM is all the kernel asks for
}
template <int KP>
int chain_params_impl(ChainDev* ch, const unsigned long long* out_u64, const double* out_f64, cudaStream_t s) {
  k_chain_params<KP><<<1, kFusedThreads, 0, s>>>(ch, out_u64, out_f64);
  return (int)cudaGetLastError();
}

}  // namespace hml
