// fp32 instantiations of the streaming fused NFP kernels (see nfp_stream_impl.cuh).
#include "nfp_stream_impl.cuh"

namespace nfp {
namespace stream {
bool plan_ok_f32(const KParams& P, int mode) { return plan_ok_dtype<float>(P, mode); }
int launch_f32(const KParams& P, int mode, const StreamArgs& a, cudaStream_t stream) {
  return launch_dtype<float>(P, mode, a, stream);
}
}  // namespace stream
}  // namespace nfp
