// Internal interface between the dtype-specific translation units of the cluster-split fused kernels
// (nfp_split_f32.cu, nfp_split_bf16.cu) and the dispatcher (nfp_stream.cu).
#pragma once

#include "nfp_common.cuh"

namespace nfp {
namespace split {

struct SplitArgs {
  const void* x;
  const void* gy;
  void* y;
  void* gx;
  const float* g_gap_x;
  const float* g_gap_nfp;
  float* gap_x;
  float* gap_nfp;
  int B, C;
  int S;         // CTAs per cluster = channel slices per image
  int Cs;        // channels per unit (C / S)
  int NSUB;      // TMA sub-chunks per unit
  int sub_ch;    // channels per sub-chunk (a whole number of work items)
  int pad_mode, similarity;
  int lanech;    // backward: lane-per-channel pass B (Cs a multiple of 64)
  int y_f32;     // forward: y is fp32 regardless of T
  int x_early;   // backward: x is stable across the preceding launch -> load it before griddepcontrol.wait
  float eps;
  unsigned long long* dbg;  // optional: 8 globaltimer stamps per CTA, see nfpb200_debug_phase_timing
};
struct PlanInfo {
  bool ok;
  int S, Cs, NSUB, lanech, ctas_per_sm;
  size_t smem;
};
PlanInfo plan_f32(const KParams& P, int mode);
PlanInfo plan_bf16(const KParams& P, int mode);
int launch_f32(const KParams& P, int mode, const SplitArgs& a, cudaStream_t stream);
int launch_bf16(const KParams& P, int mode, const SplitArgs& a, cudaStream_t stream);

}  // namespace split
}  // namespace nfp
