// Microbenchmark: how fast can ONE launch read 26 MB (256 x 100 KB, the forward's input at B = 256) that is cold in HBM,
// as a function of how the bytes are cut over CTAs and how they are requested?  Every launch reads a different buffer of a
// 308 MB rotation (> 2 x L2); launches are chained in a CUDA graph with programmatic dependent launch, like the NFP kernels.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cold_read cold_read.cu && ./cold_read
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr size_t kTotal = (size_t)256 * 100352;  // bytes per launch
// LDG form: `threads` threads per CTA, each CTA reads `per_cta` bytes with UNROLL independent 16-byte loads in flight per thread
template <int UNROLL>
__global__ void k_ldg(const uint4* __restrict__ src, float* sink, int per_cta16) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const uint4* p = src + (size_t)blockIdx.x * per_cta16;
  uint4 acc = make_uint4(0, 0, 0, 0);
  int i = threadIdx.x;
  for (; i + (UNROLL - 1) * (int)blockDim.x < per_cta16; i += UNROLL * blockDim.x) {
    uint4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) v[u] = p[i + u * blockDim.x];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) { acc.x ^= v[u].x; acc.y ^= v[u].y; acc.z ^= v[u].z; acc.w ^= v[u].w; }
  }
  for (; i < per_cta16; i += blockDim.x) { const uint4 v = p[i]; acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w; }
  if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x12345678u) sink[threadIdx.x] = 1.f;
}
// TMA form: one thread issues `ncopy` cp.async.bulk copies of `copy_bytes` each, all in flight, onto one mbarrier
__global__ void k_tma(const unsigned char* __restrict__ src, float* sink, int ncopy, int copy_bytes, float* out, int out_words) {
  extern __shared__ __align__(128) unsigned char sm[];
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(sm), dst = (uint32_t)__cvta_generic_to_shared(sm + 128);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(ncopy * copy_bytes) : "memory");
    const unsigned char* p = src + (size_t)blockIdx.x * ncopy * copy_bytes;
    for (int c = 0; c < ncopy; ++c)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(dst + c * copy_bytes), "l"(p + (size_t)c * copy_bytes), "r"(copy_bytes), "r"(bar_a) : "memory");
  }
  __syncthreads();
  asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(bar_a) : "memory");
  if (sm[128 + threadIdx.x] == 77 && sm[300 + threadIdx.x] == 78) sink[threadIdx.x] = 2.f;
  if (out)  // a small result per CTA (the forward writes 1.5 KB of similarities per image)
    for (int i = threadIdx.x; i < out_words; i += blockDim.x) out[(size_t)blockIdx.x * out_words + i] = (float)sm[128 + i];
}
// copy form (the backward's traffic without its arithmetic): every chunk is stored to `dst` as soon as it has landed
__global__ void k_copy(const unsigned char* __restrict__ src, unsigned char* __restrict__ dst_g, float* sink, int ncopy, int copy_bytes) {
  extern __shared__ __align__(128) unsigned char sm[];
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm);
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(bars), dst = (uint32_t)__cvta_generic_to_shared(sm + 128);
  if (threadIdx.x == 0) {
    for (int c = 0; c < ncopy; ++c) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8 * c));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const unsigned char* p = src + (size_t)blockIdx.x * ncopy * copy_bytes;
    for (int c = 0; c < ncopy; ++c) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8 * c), "r"(copy_bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(dst + c * copy_bytes), "l"(p + (size_t)c * copy_bytes), "r"(copy_bytes), "r"(bar0 + 8 * c) : "memory");
    }
    unsigned char* q = dst_g + (size_t)blockIdx.x * ncopy * copy_bytes;
    for (int c = 0; c < ncopy; ++c) {
      asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(bar0 + 8 * c) : "memory");
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(q + (size_t)c * copy_bytes), "r"(dst + c * copy_bytes), "r"(copy_bytes) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
  __syncthreads();
  if (sm[128 + threadIdx.x] == 77 && sm[300 + threadIdx.x] == 78) sink[threadIdx.x] = 2.f;
}
template <class F>
float timed(F launch) {
  cudaStream_t st; cudaStreamCreate(&st);
  cudaGraph_t g; cudaGraphExec_t ge;
  const int N = 120;
  cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal);
  for (int i = 0; i < N; ++i) launch(st, i);
  cudaStreamEndCapture(st, &g);
  cudaGraphInstantiate(&ge, g, 0);
  cudaGraphLaunch(ge, st); cudaStreamSynchronize(st);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, st);
  for (int r = 0; r < 5; ++r) cudaGraphLaunch(ge, st);
  cudaEventRecord(e1, st); cudaStreamSynchronize(st);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) printf("  (CUDA error: %s)\n", cudaGetErrorString(err));
  cudaGraphExecDestroy(ge); cudaGraphDestroy(g); cudaStreamDestroy(st);
  return ms * 1e3f / (5 * N);
}
void cfg_launch(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* at, int grid, int block, size_t smem, cudaStream_t st) {
  cfg = cudaLaunchConfig_t{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
}
int main() {
  const int nbuf = 12;
  unsigned char* bufs; float* sink;
  cudaMalloc(&bufs, kTotal * nbuf); cudaMalloc(&sink, 8192);
  cudaMemset(bufs, 1, kTotal * nbuf);
  cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  printf("one launch = %.1f MB cold from HBM; floor at the measured 6557 GB/s: %.2f us\n", kTotal / 1e6, kTotal / 6557e3);
  struct L { int grid, block, unroll; } ldg[] = {{256, 288, 1}, {256, 288, 4}, {256, 512, 4}, {256, 1024, 4}, {296, 512, 4},
                                                 {592, 512, 4}, {1184, 256, 4}, {2368, 256, 4}, {148, 1024, 8}};
  for (auto c : ldg) {
    const int per16 = (int)(kTotal / 16 / c.grid);
    auto f = [&](cudaStream_t st, int i) {
      cudaLaunchConfig_t cfg; cudaLaunchAttribute at[1];
      cfg_launch(cfg, at, c.grid, c.block, 0, st);
      const uint4* src = reinterpret_cast<const uint4*>(bufs + (size_t)(i % nbuf) * kTotal);
      if (c.unroll == 1) cudaLaunchKernelEx(&cfg, k_ldg<1>, src, sink, per16);
      else if (c.unroll == 4) cudaLaunchKernelEx(&cfg, k_ldg<4>, src, sink, per16);
      else cudaLaunchKernelEx(&cfg, k_ldg<8>, src, sink, per16);
    };
    const float us = timed(f);
    printf("LDG.128  grid %4d x %4d threads, %d loads in flight per thread, %6.1f KB per CTA: %5.2f us  (%4.0f GB/s)\n", c.grid, c.block,
           c.unroll, per16 * 16 / 1024.0, us, kTotal / us / 1e3);
  }
  struct T { int grid, ncopy, bytes, block, pad; } tma[] = {{256, 8, 12544, 128, 0}, {256, 1, 100352, 128, 0}, {256, 32, 3136, 128, 0},
      {512, 4, 12544, 128, 0}, {512, 1, 50176, 128, 0}, {1024, 2, 12544, 128, 0}, {2048, 1, 12544, 128, 0}, {296, 8, 10848, 128, 0},
      {148, 16, 10848, 128, 0}, {256, 8, 12544, 288, 0}, {256, 8, 12544, 288, 12000}, {256, 8, 12544, 288, 14700}};
  for (auto c : tma) {
    const size_t smem = 128 + (size_t)c.ncopy * c.bytes + 512 + c.pad;
    auto f = [&](cudaStream_t st, int i) {
      cudaLaunchConfig_t cfg; cudaLaunchAttribute at[1];
      cfg_launch(cfg, at, c.grid, c.block, smem, st);
      const unsigned char* src = bufs + (size_t)(i % nbuf) * kTotal;
      cudaLaunchKernelEx(&cfg, k_tma, src, sink, c.ncopy, c.bytes, (float*)nullptr, 0);
    };
    const float us = timed(f);
    const double tot = (double)c.grid * c.ncopy * c.bytes;
    printf("TMA bulk grid %4d x %3d threads, %2d copies of %6d B per CTA (%5.1f KB of shared memory), all in flight: %5.2f us  (%4.0f GB/s)\n",
           c.grid, c.block, c.ncopy, c.bytes, smem / 1024.0, us, tot / us / 1e3);
  }
  {  // the read launch + a 1.5 KB result per CTA written at the end
    float* outy; cudaMalloc(&outy, (size_t)256 * 392 * 4 * nbuf);
    const size_t smem = 128 + (size_t)8 * 12544 + 512;
    auto f = [&](cudaStream_t st, int i) {
      cudaLaunchConfig_t cfg; cudaLaunchAttribute at[1];
      cfg_launch(cfg, at, 256, 288, smem, st);
      cudaLaunchKernelEx(&cfg, k_tma, (const unsigned char*)(bufs + (size_t)(i % nbuf) * kTotal), sink, 8, 12544, outy + (size_t)(i % nbuf) * 256 * 392, 392);
    };
    printf("TMA bulk grid  256 x 288 threads, 8 x 12544 B per CTA + 1568 B written per CTA at the end: %5.2f us\n", timed(f));
  }
  // the step's traffic without its arithmetic: a read launch (forward: 25.7 MB in) followed by a copy launch (backward: 25.7 MB in,
  // 25.7 MB out), chained; the copy's dirty lines leave L2 during the launches that follow, as in the real step
  {
    unsigned char* outs;
    cudaMalloc(&outs, kTotal * nbuf);
    cudaFuncSetAttribute(k_copy, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    const size_t smem = 128 + (size_t)8 * 12544 + 512;
    auto fcopy = [&](cudaStream_t st, int i) {
      cudaLaunchConfig_t cfg; cudaLaunchAttribute at[1];
      cfg_launch(cfg, at, 256, 128, smem, st);
      cudaLaunchKernelEx(&cfg, k_copy, (const unsigned char*)(bufs + (size_t)(i % nbuf) * kTotal), outs + (size_t)(i % nbuf) * kTotal, sink, 8, 12544);
    };
    const float us_copy = timed(fcopy);
    auto fstep = [&](cudaStream_t st, int i) {
      cudaLaunchConfig_t cfg; cudaLaunchAttribute at[1];
      cfg_launch(cfg, at, 256, 128, smem, st);
      if (i & 1) cudaLaunchKernelEx(&cfg, k_copy, (const unsigned char*)(bufs + (size_t)((i + 5) % nbuf) * kTotal), outs + (size_t)(i % nbuf) * kTotal, sink, 8, 12544);
      else cudaLaunchKernelEx(&cfg, k_tma, (const unsigned char*)(bufs + (size_t)(i % nbuf) * kTotal), sink, 8, 12544, (float*)nullptr, 0);
    };
    const float us_step = 2 * timed(fstep);
    printf("TMA copy grid  256, 8 x 12544 B in and out per CTA (25.7 MB read + 25.7 MB written per launch): %5.2f us  (%4.0f GB/s)\n", us_copy,
           2 * kTotal / us_copy / 1e3);
    printf("read launch + copy launch, chained (the step's 77 MB without its arithmetic): %5.2f us per pair  (%4.0f GB/s)\n", us_step,
           3 * kTotal / us_step / 1e3);
  }
  return 0;
}
