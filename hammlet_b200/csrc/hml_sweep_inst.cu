// hammlet_b200 — explicit instantiations of the sweep kernels for a group of padded state counts.
// Compiled several times with different -DHML_INST_LIST=... so the groups build in parallel.
#include "hml_sweep_impl.cuh"

namespace hml {
#define HML_INST(KP)                                                                                            \
  template int sweep_impl<KP>(const ModelHost&, const SweepBuffers&, const SweepLaunch&, cudaStream_t, stage_cb_t, \
                              void*);                                                                           \
  template int sequential_impl<KP>(const ModelHost&, const SweepBuffers&, const SweepLaunch&, cudaStream_t);
HML_INST_LIST
}  // namespace hml
