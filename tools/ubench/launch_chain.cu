// Microbenchmark: the per-launch floor of a CUDA-graph chain of kernels shaped like the fused NFP launches (256 CTAs x 288
// threads, 113 KB of dynamic shared memory -> two CTAs per SM fill the SM's shared memory, so a dependent grid's CTAs can
// only become resident as the previous grid's CTAs exit), with and without programmatic dependent launch, and with one
// HBM-cold 12.5 KB TMA-sized read per CTA after the dependency wait (first-byte latency of a cold stream).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o launch_chain launch_chain.cu && ./launch_chain
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>  // 0: empty, 1: one cold 12.5 KB read per CTA (coalesced 16-byte loads), 2: read 100 KB per CTA
__global__ void __launch_bounds__(288, 2) k(const uint4* __restrict__ src, float* sink, size_t stride16) {
  extern __shared__ unsigned char sm[];
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (MODE >= 1) {
    const int n16 = MODE == 1 ? 784 : 6272;
    const uint4* p = src + (size_t)blockIdx.x * stride16;
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < n16; i += 288) {
      const uint4 v = p[i];
      acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
    }
    if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x12345678u) sink[threadIdx.x] = 1.f;
  }
  if (threadIdx.x == 9999) sm[0] = 1;
}
template <int MODE>
float run(bool pdl, const uint4* bufs, size_t buf16, int nbuf, float* sink) {
  cudaStream_t st; cudaStreamCreate(&st);
  auto kern = k<MODE>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
  cudaGraph_t g; cudaGraphExec_t ge;
  const int N = 200;
  cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal);
  for (int i = 0; i < N; ++i) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(256); cfg.blockDim = dim3(288); cfg.dynamicSmemBytes = 113 * 1024; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    const uint4* src = bufs + (size_t)(i % nbuf) * buf16;
    cudaLaunchKernelEx(&cfg, kern, src, sink, (size_t)6272);
  }
  cudaStreamEndCapture(st, &g);
  cudaGraphInstantiate(&ge, g, 0);
  cudaGraphLaunch(ge, st); cudaStreamSynchronize(st);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, st);
  for (int r = 0; r < 5; ++r) cudaGraphLaunch(ge, st);
  cudaEventRecord(e1, st);
  cudaStreamSynchronize(st);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms * 1e3f / (5 * N);
}
int main() {
  const size_t buf16 = (size_t)256 * 6272;  // 25.7 MB per buffer (256 x 100 KB)
  const int nbuf = 12;                      // 308 MB > 2 x L2: every read is HBM-cold
  uint4* bufs; float* sink;
  cudaMalloc(&bufs, buf16 * 16 * nbuf); cudaMalloc(&sink, 4096);
  cudaMemset(bufs, 1, buf16 * 16 * nbuf);
  for (int pdl = 0; pdl < 2; ++pdl) {
    printf("%s: empty kernel %.2f us/launch | + one cold 12.5 KB read per CTA %.2f us | + 100 KB per CTA (26 MB per launch), "
           "16-byte loads %.2f us   (TMA forms of the 26 MB read: cold_read.cu)\n",
           pdl ? "programmatic dependent launch" : "plain stream order          ", run<0>(pdl, bufs, buf16, nbuf, sink),
           run<1>(pdl, bufs, buf16, nbuf, sink), run<2>(pdl, bufs, buf16, nbuf, sink));
  }
  return 0;
}
