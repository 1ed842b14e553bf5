// Dispatcher of the streaming fused NFP kernels: geometry check, dtype selection, entry points used
// by the C ABI (nfp_capi.cu).  The kernels themselves live in nfp_stream_impl.cuh.
#include "nfp_stream.h"

namespace nfp {
namespace stream {
unsigned long long* g_debug_stamps = nullptr;
}
namespace {

bool geometry_ok(const KParams& P, int measure) {
  return measure == NFPB200_COSINE && P.stride == 1 && P.dil == 1 && P.pad == P.R &&
         P.mode != NFPB200_PAD_CIRCULAR;
}

int op_mode(int op) {
  switch (op) {
    case NFPB200_OP_FORWARD: return stream::MODE_FWD;
    case NFPB200_OP_BACKWARD: return stream::MODE_BWD;
    case NFPB200_OP_POOL_FORWARD: return stream::MODE_POOL_FWD;
    default: return stream::MODE_POOL_BWD;
  }
}

int run(const KParams& P, int dtype, int mode, stream::StreamArgs a, cudaStream_t s) {
  a.B = P.B; a.C = P.C;
  a.pad_mode = P.mode; a.similarity = P.similarity; a.eps = P.eps;
  a.dbg = stream::g_debug_stamps;
  return dtype == NFPB200_BF16 ? stream::launch_bf16(P, mode, a, s) : stream::launch_f32(P, mode, a, s);
}

}  // namespace

bool stream_supported(const KParams& P, int dtype, int measure, int op) {
  if (!geometry_ok(P, measure)) return false;
  return dtype == NFPB200_BF16 ? stream::plan_ok_bf16(P, op_mode(op)) : stream::plan_ok_f32(P, op_mode(op));
}

const char* stream_name(const KParams& P, int dtype, int measure, int op) {
  (void)dtype; (void)measure; (void)op;
  const char* nm = "fused/stream";
#define X(H_, W_, R_, TW_) \
  if (P.H == H_ && P.W == W_ && P.R == R_) nm = "fused/stream_" #H_ "x" #W_ "_r" #R_;
  NFP_STREAM_SHAPES(X)
#undef X
  return nm;
}

int stream_forward(const KParams& P, int dtype, const void* x, void* y, const LaunchCtx& ctx) {
  stream::StreamArgs a{};
  a.x = x; a.y = y;
  return run(P, dtype, stream::MODE_FWD, a, ctx.stream);
}
int stream_backward(const KParams& P, int dtype, const void* x, const void* gy, void* gx, const LaunchCtx& ctx) {
  stream::StreamArgs a{};
  a.x = x; a.gy = gy; a.gx = gx;
  a.x_early = P.x_stable;
  return run(P, dtype, stream::MODE_BWD, a, ctx.stream);
}
int stream_pool_forward(const KParams& P, int dtype, const void* x, float* gap_x, float* gap_nfp,
                        const LaunchCtx& ctx) {
  stream::StreamArgs a{};
  a.x = x; a.gap_x = gap_x; a.gap_nfp = gap_nfp;
  return run(P, dtype, stream::MODE_POOL_FWD, a, ctx.stream);
}
int stream_pool_backward(const KParams& P, int dtype, const void* x, const float* g_gap_x, const float* g_gap_nfp,
                         void* gx, const LaunchCtx& ctx) {
  stream::StreamArgs a{};
  a.x = x; a.g_gap_x = g_gap_x; a.g_gap_nfp = g_gap_nfp; a.gx = gx;
  a.x_early = P.x_stable;
  a.ggx_tma = (P.C % 4 == 0) && (reinterpret_cast<uintptr_t>(g_gap_x) % 16 == 0);
  return run(P, dtype, stream::MODE_POOL_BWD, a, ctx.stream);
}

}  // namespace nfp
