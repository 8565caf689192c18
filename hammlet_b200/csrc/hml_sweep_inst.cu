// hammlet_b200 — explicit instantiations of the sweep kernels for a group of padded state counts.
// Compiled several times with different -DHML_INST_LIST=... so the groups build in parallel.
#include "hml_fused.cuh"

namespace hml {
// the fused kernel exists for K <= 8 only; the other groups get stubs that report "unsupported"
template <int KP>
struct FusedInst {
  static int launch(const SweepBuffers& b, const FusedArgs& a, int grid, cudaStream_t s) {
    if constexpr (KP <= kChainMaxStates) return fused_impl<KP>(b, a, grid, s);
    return -2;
  }
  static int max_grid(int sms) {
    if constexpr (KP <= kChainMaxStates) return fused_max_grid_impl<KP>(sms);
    return 0;
  }
  static int params(ChainDev* ch, const unsigned long long* o64, const double* of, cudaStream_t s) {
    if constexpr (KP <= kChainMaxStates) return chain_params_impl<KP>(ch, o64, of, s);
    return -2;
  }
};
template <int KP> int fused_launch_kp(const SweepBuffers& b, const FusedArgs& a, int grid, cudaStream_t s) { return FusedInst<KP>::launch(b, a, grid, s); }
template <int KP> int fused_max_grid_kp(int sms) { return FusedInst<KP>::max_grid(sms); }
template <int KP> int chain_params_kp(ChainDev* ch, const unsigned long long* o64, const double* of, cudaStream_t s) { return FusedInst<KP>::params(ch, o64, of, s); }
#define HML_INST(KP)                                                                                            \
  template int sweep_impl<KP>(const ModelHost&, const SweepBuffers&, const SweepLaunch&, cudaStream_t, stage_cb_t, \
                              void*);                                                                           \
  template int sequential_impl<KP>(const ModelHost&, const SweepBuffers&, const SweepLaunch&, cudaStream_t);  \
  template int fused_launch_kp<KP>(const SweepBuffers&, const FusedArgs&, int, cudaStream_t);                      \
  template int fused_max_grid_kp<KP>(int);                                                                        \
  template int chain_params_kp<KP>(ChainDev*, const unsigned long long*, const double*, cudaStream_t);
HML_INST_LIST
}  // namespace hml
