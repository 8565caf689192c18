/* TEST INFRASTRUCTURE — CPU restatement of the reference's hot path ("the oracle").
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load the library built
 * from this file.  The product (hammlet_b200/) never includes, links or calls it.
 *
 * This header is included twice by hammlet_oracle.c, once with REAL=float (what the shipped
 * `hammlet` computes, src/includes.hpp:10) and once with REAL=double (the real_t=double build used
 * as the fp64 oracle).  Every function cites the reference lines it follows
 * (paths relative to /root/reference/src).  Mixed float/double promotions are reproduced
 * deliberately; they decide the last bits of the float build.
 *
 * Parity pin: tests/test_oracle_vs_golden.py checks these functions against dumps of the
 * reference's own classes (oracle/ref_probe.cpp -> tests/golden/*.npz), bit-for-bit for both
 * REAL types.  The reference itself holds no tests or golden vectors.
 */

#define HO_CAT2(a, b) a##_##b
#define HO_CAT(a, b) HO_CAT2(a, b)
#define FN(name) HO_CAT(name, SFX)

#ifndef HO_COMMON_ONCE
#define HO_COMMON_ONCE
#define HO_CELL 65535 /* Statistics/IntegralArray.hpp:24 */

/* uintmath.hpp:128 */
static size_t ho_ceil_pow2(size_t n) {
  size_t p = 1;
  while (p < n) p <<= 1;
  return p;
}

/* std::discrete_distribution as libstdc++ 13 implements it (bits/random.tcc): normalise in double,
 * partial sums, last := 1.0, lower_bound(u).  Used by Trellis.hpp:61-66 and Mixture.hpp:111-112.
 * An all-zero row gives NaN partial sums; lower_bound then lands on index 0. */
static int ho_discrete(const double* w, int K, double u) {
  double cp[HO_MAXK];
  double sum = 0.0;
  int k;
  if (K < 2) return 0;
  for (k = 0; k < K; ++k) sum += w[k];
  {
    double acc = 0.0;
    for (k = 0; k < K; ++k) {
      acc += w[k] / sum;
      cp[k] = acc;
    }
  }
  cp[K - 1] = 1.0;
  { /* std::lower_bound */
    int first = 0, len = K;
    while (len > 0) {
      int half = len >> 1;
      int mid = first + half;
      if (cp[mid] < u) {
        first = mid + 1;
        len = len - half - 1;
      } else
        len = half;
    }
    return first;
  }
}
#endif /* HO_COMMON_ONCE */

/* ------------------------------------------------------------------ load-time transforms */

/* wavelet.hpp:97-188 (MaxletTransform): streaming in-order Haar over D interleaved dimensions (x holds T*D
 * values, position-major).  coeffs[j] is the maximum over the dimensions (:155-160, starting from 0) of the
 * absolute detail coefficient of the wavelet whose mid discontinuity sits at j, normalised by a RUNNING product
 * of sqrt2half (includes.hpp:126-128); incomplete wavelets and position 0 stay inf. */
void FN(ho_maxlet_md)(const REAL* x, size_t T, size_t D, REAL* coeffs) {
  const REAL sqrt2 = (REAL)sqrt(2.0);
  const REAL sqrt2half = (REAL)(sqrt2 / 2.0);
  REAL* stack = (REAL*)malloc(sizeof(REAL) * 80 * D);
  size_t sp = 0, i, d;
  for (i = 0; i < T; ++i) {
    size_t j = i, m = 1;
    REAL normalizer = sqrt2half;
    for (d = 0; d < D; ++d) stack[sp++] = x[i * D + d];
    coeffs[i] = (REAL)INFINITY;
    while ((j & m) > 0) {
      REAL maxCoeff = 0;
      size_t L = sp - 2 * D, R = L + D;
      for (d = 0; d < D; ++d) {
        REAL df = stack[L] - stack[R];
        REAL c = normalizer * (REAL)RFABS(df);
        if (maxCoeff < c) maxCoeff = c; /* std::max(maxCoeff, c): NaN-safe order as in :156 */
        stack[L] += stack[R];
        L++;
        R++;
      }
      coeffs[j] = maxCoeff;
      sp -= D;
      j = j - m;
      m *= 2;
      normalizer *= sqrt2half;
    }
  }
  if (T > 0) coeffs[0] = (REAL)INFINITY;
  free(stack);
}
void FN(ho_maxlet)(const REAL* x, size_t T, REAL* coeffs) { FN(ho_maxlet_md)(x, T, 1, coeffs); }

/* main.cpp:303-311: mean of the finest-level coefficients (odd indices) over sqrt(2/pi) */
double FN(ho_sigma_hat)(const REAL* coeffs, size_t T) {
  double s = 0;
  size_t n = 0, i;
  for (i = 1; i < T; i += 2) {
    s += coeffs[i];
    n++;
  }
  s /= n;
  s /= 0.797884560802865355879892119868763736951717262329869315331;
  return s;
}

/* wavelet.hpp:68-93 (HaarBreakpointWeights), in place; then main.cpp:332-334 (weight multiplier) */
void FN(ho_breakpoint_weights)(REAL* w, size_t T, REAL mult) {
  size_t interval, index;
  for (interval = ho_ceil_pow2(T) / 2; interval >= 1; interval /= 2) {
    const size_t shift = 2 * interval;
    for (index = interval; index < T; index += shift) {
      const size_t L = index - interval, R = index + interval;
      if (R < T) {
        if (w[R] < w[index]) w[R] = w[index];
      } else {
        w[L] = (REAL)INFINITY;
        w[index] = (REAL)INFINITY;
      }
      if (w[L] < w[index]) w[L] = w[index];
    }
  }
  for (index = 0; index < T; ++index) w[index] *= mult;
}

/* Blocks/BreakpointArray.hpp:203-235: block starts for a threshold.  The uint16 skip pointers
 * (:130-184) only accelerate the search for the next t with !(w[t] < thr); the set is the same. */
size_t FN(ho_boundaries)(const REAL* w, size_t T, REAL thr, uint64_t* starts) {
  size_t t, n = 0;
  for (t = 0; t < T; ++t)
    if (t == 0 || !(w[t] < thr)) {
      if (starts) starts[n] = t;
      n++;
    }
  return n;
}

/* Blocks/BreakpointArray.hpp:195-199 + Theta.hpp:226-234: threshold from the smallest variance */
REAL FN(ho_threshold)(size_t T, const REAL* var, int nparams) {
  REAL mv = (REAL)INFINITY;
  int k;
  for (k = 0; k < nparams; ++k)
    if (var[k] < mv) mv = var[k];
  return (REAL)RSQRT(2 * RLOG((REAL)T) * mv);
}

/* Statistics/IntegralArray.hpp:136-191 + utils.hpp:15-76: (x, x*x) per value
 * (SufficientStatistics.hpp:61-65), one zero entry appended, then inside each cell of 65535
 * entries a REVERSE Kahan running sum.  isum/isq hold T+1 entries. */
void FN(ho_integral_build)(const REAL* x, size_t T, REAL* isum, REAL* isq) {
  size_t i, start;
  const size_t n = T + 1;
  for (i = 0; i < T; ++i) {
    isum[i] = x[i];
    isq[i] = x[i] * x[i];
  }
  isum[T] = 0;
  isq[T] = 0;
  for (start = 0; start < n; start += HO_CELL) {
    size_t right = start + HO_CELL;
    if (right > n) right = n;
    right--;
    if (start < right) {
      REAL s1 = isum[right], c1 = 0, s2 = isq[right], c2 = 0;
      size_t k = right;
      while (k > start) {
        REAL y, tmp;
        k--;
        y = isum[k] - c1;
        tmp = s1 + y;
        c1 = (tmp - s1) - y;
        s1 = tmp;
        isum[k] = s1;
        y = isq[k] - c2;
        tmp = s2 + y;
        c2 = (tmp - s2) - y;
        s2 = tmp;
        isq[k] = s2;
      }
    }
  }
}

/* Statistics/IntegralArray.hpp:104-124 (addBlockStats) with KahanAggregator.hpp:26-45: separate
 * compensated positive and negative accumulators, result = pos - neg. */
void FN(ho_block_stats)(const REAL* isum, const REAL* isq, size_t start, size_t end, REAL* sum, REAL* sumsq) {
  REAL ps[2] = {0, 0}, pe[2] = {0, 0}, ns[2] = {0, 0}, ne[2] = {0, 0};
  const REAL* arr[2];
  int c;
  size_t p;
  arr[0] = isum;
  arr[1] = isq;
  for (c = 0; c < 2; ++c) {
    REAL y, tmp;
    y = arr[c][start] - pe[c];
    tmp = ps[c] + y;
    pe[c] = (tmp - ps[c]) - y;
    ps[c] = tmp;
    for (p = ((start + HO_CELL) / HO_CELL) * HO_CELL; p < end; p += HO_CELL) {
      y = arr[c][p] - pe[c];
      tmp = ps[c] + y;
      pe[c] = (tmp - ps[c]) - y;
      ps[c] = tmp;
    }
    if (end % HO_CELL != 0) {
      y = arr[c][end] - ne[c];
      tmp = ns[c] + y;
      ne[c] = (tmp - ns[c]) - y;
      ns[c] = tmp;
    }
  }
  *sum = ps[0] - ns[0];
  *sumsq = ps[1] - ns[1];
}

/* ------------------------------------------------------------------ emission log-weights */

/* EFD.hpp:35-38 with the stdev cached by Observation.hpp:175-185 */
static REAL FN(ho_log_normalizer)(REAL mean, REAL var) {
  const REAL stdev = (REAL)RSQRT(var);
  return (REAL)RLOG(stdev) + mean * mean / (2 * var);
}

/* EFD.hpp:23-32: the 2.0 literals promote the whole expression to double */
static REAL FN(ho_inner_product)(REAL mean, REAL var, REAL sum, REAL sumsq) {
  return (REAL)((2.0 * mean * sum - sumsq) / (2.0 * var));
}

/* ------------------------------------------------------------------ one FBG sweep over blocks */

/* StateSequence/ForwardBackward.hpp:16-213.  Inputs: B blocks with size/sum/sumsq, K states with
 * (mean,var), A row-major K*K, pi, the uniforms in CONSUMPTION order (last block first).
 * Outputs (any may be NULL): rows[(B+1)*K] = trellis as the backward pass finds it (forward rows
 * incl. the self-transition rescale quirk, :115-119), states[B], per-state stat sums (Kahan over
 * blocks, :189-191), stat_n[K] term counts, trans[K*K] (:182-184 incl. phantom 0->q0),
 * counts[K] (:185), loglik = sum_t (maxE_t + log forwardSum_t) (never materialised by the
 * reference; accumulated here in double for the parity tests).  Returns the number of
 * "uniform fallback" events (:106-111), or -1 for a negative backward variable (:147-149). */
int FN(ho_fb_sweep_md)(size_t B, const uint64_t* bsize, const REAL* bsum, const REAL* bsq, int D, int P,
                       const int* mapping, int K, const REAL* mean, const REAL* var, const REAL* A, const REAL* pi,
                       int use_self, const double* uniforms, REAL* rows_out, int16_t* states, REAL* stat_sum,
                       REAL* stat_sq, uint64_t* stat_n, uint64_t* trans, uint64_t* counts, double* loglik) {
  REAL logA[HO_MAXK], logNorm[HO_MAXK], forward[HO_MAXK];
  REAL* rows = rows_out ? rows_out : (REAL*)malloc(sizeof(REAL) * (B + 1) * (size_t)K);
  int s, i, j, fallbacks = 0;
  size_t t;
  int d;
  REAL prevN = 1;
  double ll = 0;
  for (s = 0; s < K; ++s) {
    logA[s] = use_self ? (REAL)RLOG(A[s * K + s]) : 0;
    logNorm[s] = 0; /* Theta.hpp:154-164: sum over the state's parameters, accumulated in REAL */
    for (d = 0; d < D; ++d) logNorm[s] += FN(ho_log_normalizer)(mean[mapping[s * D + d]], var[mapping[s * D + d]]);
    rows[s] = pi[s]; /* :57 */
  }
  for (t = 1; t <= B; ++t) { /* :65-125 */
    REAL maxE = -RMAX;
    const REAL N = (REAL)bsize[t - 1];
    REAL fsum = 0;
    REAL* prev = rows + (t - 1) * (size_t)K;
    for (s = 0; s < K; ++s) {
      REAL ip = 0, E; /* EFD.hpp:83-93: per-dimension inner products accumulated in REAL */
      for (d = 0; d < D; ++d) {
        const int p = mapping[s * D + d];
        ip += FN(ho_inner_product)(mean[p], var[p], bsum[(t - 1) * D + d], bsq[(t - 1) * D + d]);
      }
      E = ip - N * logNorm[s];
      if (use_self) E += (N - 1) * logA[s];
      forward[s] = E;
      if (maxE < E) maxE = E;
    }
    for (s = 0; s < K; ++s) forward[s] = (REAL)REXP(forward[s] - maxE);
    for (j = 0; j < K; ++j) {
      REAL tt = 0;
      for (i = 0; i < K; ++i) tt += prev[i] * A[i * K + j];
      forward[j] *= tt;
      fsum += forward[j];
    }
    if (fsum != 0) {
      for (j = 0; j < K; ++j) forward[j] /= fsum;
      ll += (double)maxE + log((double)fsum);
    } else {
      fallbacks++;
      for (j = 0; j < K; ++j) forward[j] = (REAL)(1.0 / ((REAL)K));
    }
    if (use_self)
      for (s = 0; s < K; ++s) prev[s] *= (REAL)REXP((prevN - 1) * logA[s]); /* :115-119 quirk */
    for (s = 0; s < K; ++s) rows[t * (size_t)K + s] = forward[s];
    prevN = N;
  }
  if (loglik) *loglik = ll;

  if (states && B > 0) { /* :133-162 */
    double w[HO_MAXK];
    REAL* work = (REAL*)malloc(sizeof(REAL) * (size_t)K);
    size_t u = 0, tt;
    for (s = 0; s < K; ++s) w[s] = rows[B * (size_t)K + s];
    j = ho_discrete(w, K, uniforms[u++]);
    states[B - 1] = (int16_t)j;
    for (tt = B - 1; tt > 0; --tt) {
      for (i = 0; i < K; ++i) {
        work[i] = rows[tt * (size_t)K + i] * A[i * K + j];
        if (work[i] < 0) {
          free(work);
          if (!rows_out) free(rows);
          return -1;
        }
        w[i] = work[i];
      }
      j = ho_discrete(w, K, uniforms[u++]);
      states[tt - 1] = (int16_t)j;
    }
    free(work);
  }

  if (states && stat_sum) { /* :170-212 */
    REAL e1[HO_MAXK], e2[HO_MAXK];
    size_t prevState = 0;
    for (s = 0; s < P; ++s) {
      stat_sum[s] = stat_sq[s] = 0;
      e1[s] = e2[s] = 0;
      stat_n[s] = 0;
    }
    for (s = 0; s < K; ++s) counts[s] = 0;
    for (s = 0; s < K * K; ++s) trans[s] = 0;
    for (t = 0; t < B; ++t) {
      const int st = states[t];
      REAL y, tmp;
      trans[st * K + st] += bsize[t] - 1;
      trans[prevState * K + st] += 1;
      counts[st] += bsize[t];
      for (d = 0; d < D; ++d) { /* :189-191: stats[mapping[state][d]].add(y.suffStat(d), N) */
        const int p = mapping[st * D + d];
        y = bsum[t * D + d] - e1[p];
        tmp = stat_sum[p] + y;
        e1[p] = (tmp - stat_sum[p]) - y;
        stat_sum[p] = tmp;
        y = bsq[t * D + d] - e2[p];
        tmp = stat_sq[p] + y;
        e2[p] = (tmp - stat_sq[p]) - y;
        stat_sq[p] = tmp;
        stat_n[p] += bsize[t];
      }
      prevState = (size_t)st;
    }
  }
  if (!rows_out) free(rows);
  return fallbacks;
}

/* univariate data: one parameter per state, identity mapping */
int FN(ho_fb_sweep)(size_t B, const uint64_t* bsize, const REAL* bsum, const REAL* bsq, int K, const REAL* mean,
                    const REAL* var, const REAL* A, const REAL* pi, int use_self, const double* uniforms,
                    REAL* rows_out, int16_t* states, REAL* stat_sum, REAL* stat_sq, uint64_t* stat_n,
                    uint64_t* trans, uint64_t* counts, double* loglik) {
  int ident[HO_MAXK], s;
  for (s = 0; s < K; ++s) ident[s] = s;
  return FN(ho_fb_sweep_md)(B, bsize, bsum, bsq, 1, K, ident, K, mean, var, A, pi, use_self, uniforms, rows_out,
                            states, stat_sum, stat_sq, stat_n, trans, counts, loglik);
}

/* StateSequence/Mixture.hpp:31-144: independent categorical draw per block; uniforms are consumed
 * in block order; no transition term, no self-transition term, pi ignored. */
int FN(ho_mix_sweep_md)(size_t B, const uint64_t* bsize, const REAL* bsum, const REAL* bsq, int D, int P,
                        const int* mapping, int K, const REAL* mean, const REAL* var, const double* uniforms,
                        int16_t* states, REAL* stat_sum, REAL* stat_sq, uint64_t* stat_n, uint64_t* trans,
                        uint64_t* counts) {
  REAL logNorm[HO_MAXK], wts[HO_MAXK], e1[HO_MAXK], e2[HO_MAXK];
  double w[HO_MAXK];
  int s, d;
  size_t t, prevState = 0;
  for (s = 0; s < K; ++s) {
    logNorm[s] = 0;
    for (d = 0; d < D; ++d) logNorm[s] += FN(ho_log_normalizer)(mean[mapping[s * D + d]], var[mapping[s * D + d]]);
    counts[s] = 0;
  }
  for (s = 0; s < P; ++s) {
    stat_sum[s] = stat_sq[s] = 0;
    e1[s] = e2[s] = 0;
    stat_n[s] = 0;
  }
  for (s = 0; s < K * K; ++s) trans[s] = 0;
  for (t = 0; t < B; ++t) {
    REAL maxE = -RMAX, y, tmp;
    const size_t N = bsize[t];
    int st;
    for (s = 0; s < K; ++s) {
      /* Mixture.hpp:98: `N * logNormalizers[s]` with N a size_t -> converted to REAL */
      REAL ip = 0, E;
      for (d = 0; d < D; ++d) {
        const int p = mapping[s * D + d];
        ip += FN(ho_inner_product)(mean[p], var[p], bsum[t * D + d], bsq[t * D + d]);
      }
      E = ip - (REAL)N * logNorm[s];
      wts[s] = E;
      if (maxE < E) maxE = E;
    }
    for (s = 0; s < K; ++s) {
      wts[s] = (REAL)REXP(wts[s] - maxE);
      w[s] = wts[s];
    }
    st = ho_discrete(w, K, uniforms[t]);
    states[t] = (int16_t)st;
    counts[st] += N;
    trans[st * K + st] += N - 1;
    trans[prevState * K + st] += 1;
    for (d = 0; d < D; ++d) {
      const int p = mapping[st * D + d];
      y = bsum[t * D + d] - e1[p];
      tmp = stat_sum[p] + y;
      e1[p] = (tmp - stat_sum[p]) - y;
      stat_sum[p] = tmp;
      y = bsq[t * D + d] - e2[p];
      tmp = stat_sq[p] + y;
      e2[p] = (tmp - stat_sq[p]) - y;
      stat_sq[p] = tmp;
      stat_n[p] += N;
    }
    prevState = (size_t)st;
  }
  return 0;
}

int FN(ho_mix_sweep)(size_t B, const uint64_t* bsize, const REAL* bsum, const REAL* bsq, int K, const REAL* mean,
                     const REAL* var, const double* uniforms, int16_t* states, REAL* stat_sum, REAL* stat_sq,
                     uint64_t* stat_n, uint64_t* trans, uint64_t* counts) {
  int ident[HO_MAXK], s;
  for (s = 0; s < K; ++s) ident[s] = s;
  return FN(ho_mix_sweep_md)(B, bsize, bsum, bsq, 1, K, ident, K, mean, var, uniforms, states, stat_sum, stat_sq,
                             stat_n, trans, counts);
}

/* ------------------------------------------------------------------ conjugate updates */

/* Conjugate.hpp:120-168; hp = {alpha, beta, mu0, nu} updated in place.  Returns 0, or 1 if the
 * observation count is zero (reference prints a warning and leaves hp unchanged), or -1 where the
 * reference throws (negative sum of squares; zero count with positive sum of squares). */
int FN(ho_nig_update)(REAL* hp, REAL sum, REAL sumSq, uint64_t counts) {
  if (counts == 0) return sumSq > 0 ? -1 : 1;
  if (sumSq < 0) return -1;
  {
    const double N = (double)counts;
    const REAL xbar = (REAL)(sum / N);
    const REAL alpha = hp[0], beta = hp[1], mu0 = hp[2], nu = hp[3];
    REAL ssN = (REAL)((sum * sum) / N);
    if (ssN > sumSq) ssN = sumSq;
    hp[0] = (REAL)(alpha + N / 2.0);
    hp[1] = (REAL)(beta + ((sumSq + (N * nu / (N + nu)) * ((xbar - mu0) * (xbar - mu0))) - ssN) / 2.0);
    hp[2] = (REAL)((nu * mu0 + sum) / (nu + N));
    hp[3] = (REAL)(nu + N);
  }
  return 0;
}

/* Conjugate.hpp:177-205: alpha += count (size_t converted to REAL, added in REAL) */
void FN(ho_dirichlet_update)(REAL* alphas, const uint64_t* counts, size_t n) {
  size_t i;
  for (i = 0; i < n; ++i) alphas[i] += (REAL)counts[i];
}

/* AutoPriors.hpp:18-110 given the block list at threshold (REAL)(sqrt(2 log T) * sigma_hat):
 * block means and their squares are accumulated in REAL (SufficientStatistics.hpp:88-91), mean and
 * variance formed in double but RETURNED as REAL (EFD.hpp:41-60), closed form in mixed precision. */
int FN(ho_auto_prior_md)(size_t B, const uint64_t* bsize, const REAL* bsum, size_t D, REAL s2, REAL p, REAL* out4) {
  REAL mSum = 0, mSumSq = 0;
  size_t t, d;
  for (t = 0; t < B; ++t)
    for (d = 0; d < D; ++d) { /* :99-104: block-major, dimension-minor; N = nrBlocks * nrDim */
      const REAL m = bsum[t * D + d] / (REAL)bsize[t];
      mSum += m;
      mSumSq += m * m;
    }
  {
    const double n = (double)(B * D);
    const REAL meanR = (REAL)(mSum / n);
    const double blocksMean = meanR;
    const double avg = meanR;
    const REAL varR = (REAL)(mSumSq / n - (avg * avg));
    const double blocksVariance = varR;
    const REAL dataMean = (REAL)blocksMean, dataVar = (REAL)blocksVariance;
    const REAL M1 = (REAL)0.3361, M2 = (REAL)-0.0042, M3 = (REAL)-0.0201;
    const REAL b = -(REAL)RLOG(p);
    const REAL sb = (REAL)RSQRT(b);
    const REAL alpha = (REAL)2.0;
    const REAL beta = (REAL)(s2 * ((2.0 * sb) / (M1 * sb + sqrt(2.0) * (M2 * b * (REAL)REXP(M3 * sb) + 1)) + b));
    const REAL nu = beta / dataVar;
    if (p < 0 || p > 1 || s2 <= 0 || dataVar <= 0) return -1;
    if (!(beta > 0) || !(nu > 0) || !isfinite(beta) || !isfinite(nu) || !isfinite(dataMean)) return -1;
    out4[0] = alpha;
    out4[1] = beta;
    out4[2] = dataMean;
    out4[3] = nu;
  }
  return 0;
}

int FN(ho_auto_prior)(size_t B, const uint64_t* bsize, const REAL* bsum, REAL s2, REAL p, REAL* out4) {
  return FN(ho_auto_prior_md)(B, bsize, bsum, 1, s2, p, out4);
}

#undef FN
