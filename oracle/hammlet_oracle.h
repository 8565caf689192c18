/* TEST INFRASTRUCTURE — prototypes of the oracle (CPU restatement of the reference's hot path).
 * See hammlet_oracle_impl.h for the reference lines each function follows.  X = f32 | f64. */
#ifndef HAMMLET_ORACLE_H
#define HAMMLET_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#define HO_DECL(R, X)                                                                                        \
  void ho_maxlet_##X(const R* x, size_t T, R* coeffs);                                                       \
  void ho_maxlet_md_##X(const R* x, size_t T, size_t D, R* coeffs);                                          \
  int ho_fb_sweep_md_##X(size_t B, const uint64_t* bsize, const R* bsum, const R* bsq, int D, int P,         \
                         const int* mapping, int K, const R* mean, const R* var, const R* A, const R* pi,   \
                         int use_self, const double* uniforms, R* rows_out, int16_t* states, R* stat_sum,   \
                         R* stat_sq, uint64_t* stat_n, uint64_t* trans, uint64_t* counts, double* loglik);  \
  int ho_mix_sweep_md_##X(size_t B, const uint64_t* bsize, const R* bsum, const R* bsq, int D, int P,        \
                          const int* mapping, int K, const R* mean, const R* var, const double* uniforms,   \
                          int16_t* states, R* stat_sum, R* stat_sq, uint64_t* stat_n, uint64_t* trans,      \
                          uint64_t* counts);                                                                 \
  int ho_auto_prior_md_##X(size_t B, const uint64_t* bsize, const R* bsum, size_t D, R s2, R p, R* out4);    \
  double ho_sigma_hat_##X(const R* coeffs, size_t T);                                                        \
  void ho_breakpoint_weights_##X(R* w, size_t T, R mult);                                                    \
  size_t ho_boundaries_##X(const R* w, size_t T, R thr, uint64_t* starts);                                   \
  R ho_threshold_##X(size_t T, const R* var, int nparams);                                                   \
  void ho_integral_build_##X(const R* x, size_t T, R* isum, R* isq);                                         \
  void ho_block_stats_##X(const R* isum, const R* isq, size_t start, size_t end, R* sum, R* sumsq);          \
  int ho_fb_sweep_##X(size_t B, const uint64_t* bsize, const R* bsum, const R* bsq, int K, const R* mean,    \
                      const R* var, const R* A, const R* pi, int use_self, const double* uniforms,           \
                      R* rows_out, int16_t* states, R* stat_sum, R* stat_sq, uint64_t* stat_n,               \
                      uint64_t* trans, uint64_t* counts, double* loglik);                                    \
  int ho_mix_sweep_##X(size_t B, const uint64_t* bsize, const R* bsum, const R* bsq, int K, const R* mean,   \
                       const R* var, const double* uniforms, int16_t* states, R* stat_sum, R* stat_sq,       \
                       uint64_t* stat_n, uint64_t* trans, uint64_t* counts);                                 \
  int ho_nig_update_##X(R* hp, R sum, R sumSq, uint64_t counts);                                             \
  void ho_dirichlet_update_##X(R* alphas, const uint64_t* counts, size_t n);                                 \
  int ho_auto_prior_##X(size_t B, const uint64_t* bsize, const R* bsum, R s2, R p, R* out4);
HO_DECL(float, f32)
HO_DECL(double, f64)
int ho_max_states(void);
#endif
