// Microbenchmark: do warp shuffles share the shared-memory data path with LDS on sm_100a?
// Three kernels per configuration: N x LDS, N x SHFL, N x (LDS + SHFL) interleaved; if T(both) ~ max -> separate
// paths, if ~ sum -> shared.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o shfl_vs_lds shfl_vs_lds.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int N = 4096;
template <int MODE>
__global__ void __launch_bounds__(512) k(float* out) {
  __shared__ float sm[512 * 4];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = (float)i;
  __syncthreads();
  float a0 = threadIdx.x, a1 = 1.f, a2 = 2.f, a3 = 3.f;
  int idx = threadIdx.x;
#pragma unroll 1
  for (int it = 0; it < N / 4; ++it) {
    if (MODE & 1) {
      a0 += sm[idx]; a1 += sm[idx + 512]; a2 += sm[idx + 1024]; a3 += sm[idx + 1536];
      idx = (idx + 32) & 511;  // addresses change every iteration: the loads cannot be hoisted
    }
    if (MODE & 2) {
      a0 += __shfl_down_sync(0xffffffffu, a1, 7); a1 += __shfl_down_sync(0xffffffffu, a2, 7);
      a2 += __shfl_down_sync(0xffffffffu, a3, 7); a3 += __shfl_down_sync(0xffffffffu, a0, 7);
    }
    if (MODE == 4) {  // FMA-only reference
      a0 = fmaf(a0, 1.0001f, a1); a1 = fmaf(a1, 1.0001f, a2); a2 = fmaf(a2, 1.0001f, a3); a3 = fmaf(a3, 1.0001f, a0);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
}
template <int MODE>
float run(float* d, int blocks) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks, 512>>>(d); cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int i = 0; i < 10; ++i) k<MODE><<<blocks, 512>>>(d);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / 10;
}
int main() {
  float* d; cudaMalloc(&d, 148 * 2 * 512 * 4 * 4);
  for (int blocks : {148, 296}) {
    float t1 = run<1>(d, blocks), t2 = run<2>(d, blocks), t3 = run<3>(d, blocks), t4 = run<4>(d, blocks);
    // per SM: warps = blocks/148*16; instr per warp N; cycles/instr/SM
    double w = blocks / 148.0 * 16;
    printf("blocks %d (%.0f warps/SM): LDS %.3f ms (%.2f clk/warp-instr/SM @1.9GHz)  SHFL %.3f ms (%.2f)  both %.3f ms  fma %.3f ms\n",
           blocks, w, t1, t1 * 1e-3 * 1.9e9 / (N * w), t2, t2 * 1e-3 * 1.9e9 / (N * w), t3, t4);
  }
  return 0;
}
