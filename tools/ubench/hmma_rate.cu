// Microbenchmark: issue interval and dependent latency of the legacy mma.sync.m16n8k16 bf16 (SASS HMMA.16816.F32.BF16)
// on sm_100a, and of ldmatrix.x4.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hmma_rate hmma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <int NACC, bool LDSM>
__global__ void k(float* out, int iters, long long* cyc) {
  __shared__ __align__(128) uint32_t sm[32 * 32 * 4];
  for (int i = threadIdx.x; i < 32 * 32 * 4; i += blockDim.x) sm[i] = 0x3f803f80u;
  __syncthreads();
  float acc[NACC][4] = {};
  uint32_t a[4] = {0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u}, b0 = 0x3f803f80u, b1 = 0x3f803f80u;
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 16 + (threadIdx.x >> 5) * 512;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (LDSM) {
      asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(addr));
    }
#pragma unroll
    for (int j = 0; j < NACC; ++j) mma(acc[j], a, b0, b1);
  }
  long long t1 = clock64();
  float s = 0;
  for (int j = 0; j < NACC; ++j) s += acc[j][0] + acc[j][1] + acc[j][2] + acc[j][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int NACC, bool LDSM>
void run(int warps, float* out, long long* dc) {
  const int iters = 2000;
  k<NACC, LDSM><<<148, warps * 32>>>(out, iters, dc);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
  const double per_sm_hmma = (double)iters * NACC * warps;
  printf("warps/SM %2d  acc chains/warp %d  ldsm %d: %8lld cycles  -> %.1f cycles per HMMA per SM sub-partition (4 SMSPs), %.0f dense TFLOP/s at 1.965 GHz\n",
         warps, NACC, (int)LDSM, c, c / (per_sm_hmma / 4), per_sm_hmma * 4096 / c * 148 * 1.965e9 / 1e12);
}
int main() {
  float* out; cudaMalloc(&out, 148 * 1024 * 4);
  long long* dc; cudaMalloc(&dc, 8);
  for (int w : {1, 4, 8, 16}) { run<1, false>(w, out, dc); run<3, false>(w, out, dc); run<8, false>(w, out, dc); run<3, true>(w, out, dc); run<8, true>(w, out, dc); }
  return 0;
}
