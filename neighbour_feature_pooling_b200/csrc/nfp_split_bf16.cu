// bf16 instantiations of the cluster-split fused NFP kernels (see nfp_split_impl.cuh).
#include "nfp_split_impl.cuh"

namespace nfp {
namespace split {
PlanInfo plan_bf16(const KParams& P, int mode) {
  const Plan p = plan_dtype<__nv_bfloat16>(P, mode);
  return PlanInfo{p.ok, p.S, p.Cs, p.NSUB, p.lanech, p.ctas_per_sm, p.smem};
}
int launch_bf16(const KParams& P, int mode, const SplitArgs& a, cudaStream_t stream) {
  return launch_dtype<__nv_bfloat16>(P, mode, a, stream);
}
}  // namespace split
}  // namespace nfp
