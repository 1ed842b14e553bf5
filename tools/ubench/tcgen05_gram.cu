// Microbenchmark + numerics check: the two channel contractions of the channels-last NFP kernels (csrc/nfp_token.cu) on
// the 5th-generation tensor cores -- tcgen05.mma kind::f16 with TMEM accumulators -- for one 512x7x7 bf16 image per step:
//
//   Gram   G (64x64 fp32)  = X X^T          X = the image as a P x C matrix (49 pixel rows padded to 64, C = 512 channels)
//                                           A = B = X, both K-major (channels contiguous); 8 k-blocks x 4 MMAs (M64 N64 K16)
//   apply  D (64x512 fp32) = (Mhi + Mlo) X  M = the banded P x P coefficient matrix (bf16 hi + lo parts, K-major A),
//                                           B = the SAME shared-memory image read MN-major (channel blocks of 64 = LBO,
//                                           8-pixel groups = SBO); two halves of N = 256, 4 k-steps, hi + lo: 16 MMAs
//
// One CTA (4 warps) per SM, persistent over the images; X sits in shared memory in the canonical 128-byte-swizzled layout
// (tile = 64 pixel rows x 128 bytes per 64-channel block, 16-byte chunk j of row r stored at j ^ (r & 7)): the same bytes
// serve as K-major operand of the Gram and as MN-major B operand of the apply.  Thread 0 issues the MMAs and commits them
// to an mbarrier; the four warps read the accumulators back with tcgen05.ld (M = 64: row m lives in TMEM lane
// 32*(m/16) + m%16, so lanes 0..15 of every warp hold rows).  Per phase, thread 0 records clock64() deltas.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tcgen05_gram tcgen05_gram.cu && ./tcgen05_gram
//
// Prints max |error| of both products against a double-precision CPU evaluation on the first images, and the cycles per
// image of each phase (median over CTAs).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <algorithm>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

constexpr int P = 49, PP = 64, C = 512, KB = C / 64;  // 8 channel blocks of 64
constexpr int TILE = PP * 128;                        // bytes per (64 rows x 64 bf16) swizzled tile
constexpr int SM_X = 0, SM_MHI = KB * TILE, SM_MLO = SM_MHI + TILE, SM_BAR = SM_MLO + TILE;
constexpr int kSmemRequest = 160 * 1024;              // > half an SM: one CTA per SM (it allocates all 512 TMEM columns)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address, leading / stride byte offsets (>> 4),
// version 1 (sm_100), 128-byte swizzle
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): fp32 accumulate, bf16 x bf16, operand majors, N >> 3, M >> 4
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\tbra WAIT;\n\tDONE:\n\t}"
      ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// byte offset of element (row r, element e of the 64 in a row) inside a swizzled tile
__device__ __forceinline__ int tile_off(int r, int e) { return r * 128 + ((((e >> 3) ^ (r & 7)) << 4) | ((e & 7) << 1)); }

// x: (B, P, C) bf16 channels-last.  mhi / mlo: (64, 64) bf16 row-major.  gram: (nver, 64, 64) fp32, dapp: (nver, 64, C) fp32.
// cyc: (gridDim, 6): summed clock64 deltas of Gram MMAs, Gram read-back, apply MMAs, apply read-back; then the
// issue-bound cycles per image-sized Gram / apply (MMAs of 16 images back to back, no read-back in between).
__global__ void __launch_bounds__(128, 1)
k_tcgen05(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ mhi, const __nv_bfloat16* __restrict__ mlo,
          float* __restrict__ gram, float* __restrict__ dapp, int B, int nver, long long* __restrict__ cyc, float* sink) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + SM_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM_BAR + 16);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {  // one warp allocates all 512 columns and hands the permit back
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // zero everything once (rows 49..63 of every X tile stay zero), then the coefficient tiles
  for (int i = tid; i < SM_BAR / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  for (int i = tid; i < 64 * 64; i += 128) {
    const int r = i >> 6, e = i & 63;
    *reinterpret_cast<__nv_bfloat16*>(smem + SM_MHI + tile_off(r, e)) = mhi[i];
    *reinterpret_cast<__nv_bfloat16*>(smem + SM_MLO + tile_off(r, e)) = mlo[i];
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t xs = smem_u32(smem + SM_X);
  constexpr uint32_t IDESC_GRAM = umma_idesc(64, 64, 0, 0), IDESC_APPLY = umma_idesc(64, 256, 0, 1);
  uint32_t parity = 0;
  long long c_gram = 0, c_gram_rd = 0, c_app = 0, c_app_rd = 0;
  float keep = 0.f;

  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    // ---- image -> swizzled tiles (16-byte chunks; plain loads: the load path is not what is measured here)
    const uint4* src = reinterpret_cast<const uint4*>(x + (size_t)b * P * C);
    for (int i = tid; i < P * (C / 8); i += 128) {
      const int r = i / (C / 8), ch = i - r * (C / 8), kb = ch >> 3, j = ch & 7;
      *reinterpret_cast<uint4*>(smem + SM_X + kb * TILE + r * 128 + ((j ^ (r & 7)) << 4)) = src[i];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
    __syncthreads();
    long long t0 = clock64();
    // ---- Gram: 8 k-blocks x 4 k-steps of 16 channels (32 bytes inside the 128-byte swizzle atom)
    if (tid == 0) {
      fence_after();
      for (int kb = 0; kb < KB; ++kb)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t d = umma_desc(xs + kb * TILE + ks * 32, 16, 1024);
          umma(tmem, d, d, IDESC_GRAM, (kb | ks) ? 1u : 0u);
        }
      umma_commit(bar);
    }
    mbar_wait(bar, parity);
    parity ^= 1u;
    fence_after();
    long long t1 = clock64();
    {  // read-back: warp w owns TMEM lanes 32w .. 32w+31; lanes 0..15 of it are rows 16w .. 16w+15
      uint32_t v[32];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        tmem_ld32(tmem + ((uint32_t)(32 * warp) << 16) + h * 32, v);
        tmem_ld_wait();
        if (lane < 16) {
          if (b < nver) {
            float* g = gram + ((size_t)b * 64 + 16 * warp + lane) * 64 + h * 32;
#pragma unroll
            for (int i = 0; i < 32; ++i) g[i] = __uint_as_float(v[i]);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) keep += __uint_as_float(v[i]);
          }
        }
      }
    }
    fence_before();
    __syncthreads();
    long long t2 = clock64();
    // ---- apply: D (64 x 256 per half) = (Mhi + Mlo) X; K = 64 pixels in 4 steps of 16 (2 KB apart in the tiles)
    for (int half = 0; half < 2; ++half) {
      long long t3 = clock64();
      if (tid == 0) {
        fence_after();
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t bd = umma_desc(xs + half * 4 * TILE + ks * 16 * 128, TILE, 1024);
          umma(tmem + 64, umma_desc(smem_u32(smem + SM_MHI) + ks * 32, 16, 1024), bd, IDESC_APPLY, ks ? 1u : 0u);
          umma(tmem + 64, umma_desc(smem_u32(smem + SM_MLO) + ks * 32, 16, 1024), bd, IDESC_APPLY, 1u);
        }
        umma_commit(bar);
      }
      mbar_wait(bar, parity);
      parity ^= 1u;
      fence_after();
      long long t4 = clock64();
      uint32_t v[32];
#pragma unroll 1
      for (int h = 0; h < 8; ++h) {
        tmem_ld32(tmem + ((uint32_t)(32 * warp) << 16) + 64 + h * 32, v);
        tmem_ld_wait();
        if (lane < 16) {
          if (b < nver) {
            float* g = dapp + ((size_t)b * 64 + 16 * warp + lane) * C + half * 256 + h * 32;
#pragma unroll
            for (int i = 0; i < 32; ++i) g[i] = __uint_as_float(v[i]);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) keep += __uint_as_float(v[i]);
          }
        }
      }
      fence_before();
      __syncthreads();
      long long t5 = clock64();
      c_app += t4 - t3;
      c_app_rd += t5 - t4;
    }
    c_gram += t1 - t0;
    c_gram_rd += t2 - t1;
  }
  // ---- issue-rate bound: the same MMAs back to back (no read-back in between), several accumulators in flight
  long long thr[2] = {0, 0};
  {
    constexpr int REP = 16;
    __syncthreads();
    long long t0 = clock64();
    if (tid == 0) {
      fence_after();
      for (int rep = 0; rep < REP; ++rep)
        for (int kb = 0; kb < KB; ++kb)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t d = umma_desc(xs + kb * TILE + ks * 32, 16, 1024);
            umma(tmem + (rep & 3) * 64, d, d, IDESC_GRAM, (kb | ks) ? 1u : 0u);
          }
      umma_commit(bar);
    }
    mbar_wait(bar, parity);
    parity ^= 1u;
    fence_after();
    long long t1 = clock64();
    if (tid == 0) {
      for (int rep = 0; rep < REP; ++rep)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t bd = umma_desc(xs + (rep & 1) * 4 * TILE + ks * 16 * 128, TILE, 1024);
          umma(tmem + (rep & 1) * 256, umma_desc(smem_u32(smem + SM_MHI) + ks * 32, 16, 1024), bd, IDESC_APPLY, ks ? 1u : 0u);
          umma(tmem + (rep & 1) * 256, umma_desc(smem_u32(smem + SM_MLO) + ks * 32, 16, 1024), bd, IDESC_APPLY, 1u);
        }
      umma_commit(bar);
    }
    mbar_wait(bar, parity);
    parity ^= 1u;
    fence_after();
    long long t2 = clock64();
    thr[0] = (t1 - t0) / REP;          // cycles per image-sized Gram (32 MMAs)
    thr[1] = (t2 - t1) * 2 / REP;      // cycles per image-sized apply (2 halves = 16 MMAs)
  }
  if (tid == 0) {
    cyc[blockIdx.x * 6 + 4] = thr[0];
    cyc[blockIdx.x * 6 + 5] = thr[1];
    cyc[blockIdx.x * 6 + 0] = c_gram;
    cyc[blockIdx.x * 6 + 1] = c_gram_rd;
    cyc[blockIdx.x * 6 + 2] = c_app;
    cyc[blockIdx.x * 6 + 3] = c_app_rd;
  }
  if (keep == 123.456f) sink[tid] = keep;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

#define CK(e) do { cudaError_t err_ = (e); if (err_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(err_), __LINE__); return 1; } } while (0)

int main() {
  const int B = 148 * 8, nver = 3;
  std::vector<__nv_bfloat16> hx((size_t)B * P * C), hmhi(64 * 64), hmlo(64 * 64);
  std::vector<float> hm(64 * 64, 0.f);
  srand(1);
  for (auto& v : hx) v = __float2bfloat16((rand() / (float)RAND_MAX) * 2.f - 1.f);
  for (int p = 0; p < P; ++p)   // a banded matrix like the 3x3 stencil's
    for (int q = 0; q < P; ++q) {
      const int dy = q / 7 - p / 7, dx = q % 7 - p % 7;
      if (abs(dy) <= 1 && abs(dx) <= 1) hm[p * 64 + q] = (rand() / (float)RAND_MAX) - 0.5f;
    }
  for (int i = 0; i < 64 * 64; ++i) {
    hmhi[i] = __float2bfloat16(hm[i]);
    hmlo[i] = __float2bfloat16(hm[i] - __bfloat162float(hmhi[i]));
  }
  __nv_bfloat16 *dx, *dmhi, *dmlo;
  float *dg, *dd, *dsink;
  long long* dc;
  CK(cudaMalloc(&dx, hx.size() * 2)); CK(cudaMalloc(&dmhi, 64 * 64 * 2)); CK(cudaMalloc(&dmlo, 64 * 64 * 2));
  CK(cudaMalloc(&dg, (size_t)nver * 64 * 64 * 4)); CK(cudaMalloc(&dd, (size_t)nver * 64 * C * 4));
  CK(cudaMalloc(&dsink, 128 * 4)); CK(cudaMalloc(&dc, 148 * 6 * 8));
  CK(cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dmhi, hmhi.data(), 64 * 64 * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dmlo, hmlo.data(), 64 * 64 * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dg, 0, (size_t)nver * 64 * 64 * 4)); CK(cudaMemset(dd, 0, (size_t)nver * 64 * C * 4));
  CK(cudaFuncSetAttribute(k_tcgen05, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemRequest));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    k_tcgen05<<<148, 128, kSmemRequest>>>(dx, dmhi, dmlo, dg, dd, B, nver, dc, dsink);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
  }
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  std::vector<float> hg((size_t)nver * 64 * 64), hd((size_t)nver * 64 * C);
  std::vector<long long> hc(148 * 6);
  CK(cudaMemcpy(hg.data(), dg, hg.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hd.data(), dd, hd.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hc.data(), dc, hc.size() * 8, cudaMemcpyDeviceToHost));
  double eg = 0, sg = 0, ed = 0, sd = 0;
  for (int b = 0; b < nver; ++b) {
    const __nv_bfloat16* xb = hx.data() + (size_t)b * P * C;
    for (int p = 0; p < P; ++p)
      for (int q = 0; q < P; ++q) {
        double s = 0;
        for (int c = 0; c < C; ++c) s += (double)__bfloat162float(xb[p * C + c]) * (double)__bfloat162float(xb[q * C + c]);
        eg = std::max(eg, fabs(s - hg[((size_t)b * 64 + p) * 64 + q]));
        sg = std::max(sg, fabs(s));
      }
    for (int p = 0; p < P; ++p)
      for (int c = 0; c < C; ++c) {
        double s = 0;
        for (int q = 0; q < P; ++q)
          s += ((double)__bfloat162float(hmhi[p * 64 + q]) + (double)__bfloat162float(hmlo[p * 64 + q])) *
               (double)__bfloat162float(xb[q * C + c]);
        ed = std::max(ed, fabs(s - hd[((size_t)b * 64 + p) * C + c]));
        sd = std::max(sd, fabs(s));
      }
  }
  printf("numerics vs double on %d images: Gram max|err| %.3e (max |G| %.1f), apply max|err| %.3e (max |D| %.2f)\n", nver, eg, sg, ed, sd);
  const int per_cta = B / 148;
  const char* names[6] = {"Gram: 32 x tcgen05.mma M64 N64 K16 (+ commit, barrier wait)", "Gram read-back: 2 x tcgen05.ld 32x32b.x32 per warp",
                          "apply: 16 x tcgen05.mma M64 N256 K16 (two halves, hi + lo)", "apply read-back: 16 x tcgen05.ld 32x32b.x32 per warp",
                          "issue-bound Gram (16 images' MMAs back to back, one commit)", "issue-bound apply (same)"};
  for (int k = 0; k < 6; ++k) {
    std::vector<double> v;
    for (int c = 0; c < 148; ++c) v.push_back((double)hc[c * 6 + k] / (k < 4 ? per_cta : 1));
    std::sort(v.begin(), v.end());
    printf("%-62s: %7.0f cycles per image (median over CTAs; min %.0f, max %.0f) = %.2f us at 1.965 GHz\n", names[k], v[74], v[0], v[147],
           v[74] / 1965.0);
  }
  printf("whole kernel: %.1f us for %d images on 148 CTAs (incl. the plain-load staging of X and the verification stores)\n", ms * 1e3, B);
  return (eg < 1e-2 && ed < 1e-3) ? 0 : 2;
}
