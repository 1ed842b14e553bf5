// bf16-I/O (fp32 accumulate) instantiations of the streaming fused NFP kernels (see nfp_stream_impl.cuh).
#include "nfp_stream_impl.cuh"

namespace nfp {
namespace stream {
bool plan_ok_bf16(const KParams& P, int mode) { return plan_ok_dtype<__nv_bfloat16>(P, mode); }
int launch_bf16(const KParams& P, int mode, const StreamArgs& a, cudaStream_t stream) {
  return launch_dtype<__nv_bfloat16>(P, mode, a, stream);
}
}  // namespace stream
}  // namespace nfp
