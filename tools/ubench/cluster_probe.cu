// Microbenchmark: how do thread-block clusters of 1 / 2 / 4 / 8 CTAs pack onto the 148 SMs of a B200 when several
// CTAs fit per SM, and what does a cluster barrier + a DSMEM table exchange cost?  Each CTA records its SM id, start and
// end time; the body loads `bytes` from global with one TMA bulk copy, does a cluster barrier pair and spins `spin_ns`.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cluster_probe cluster_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <map>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t gtimer() { uint64_t t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ uint32_t smid() { uint32_t s; asm volatile("mov.u32 %0, %%smid;" : "=r"(s)); return s; }
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }

struct Rec { uint32_t sm, rank; uint64_t t0, t1, ts, t2; };

__global__ void probe(Rec* rec, const float* src, int bytes, int spin_ns, int csize, int nsync) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  const uint64_t t0 = gtimer();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0 && bytes) {
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(smem)),
                 "l"(src + (size_t)blockIdx.x * (bytes / 4)), "r"(bytes), "r"(b)
                 : "memory");
  }
  if (bytes) {
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                   : "=r"(ok)
                   : "r"(b)
                   : "memory");
    }
  }
  const uint64_t t1 = gtimer();
  if (csize > 1) {
    for (int i = 0; i < nsync; ++i) {
      // exchange: write one float per thread into the next CTA's shared memory
      uint32_t remote;
      const uint32_t local = (uint32_t)__cvta_generic_to_shared(smem) + threadIdx.x * 4;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"((cluster_rank() + 1) % csize));
      asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(1.0f) : "memory");
      cluster_arrive();
      cluster_wait();
    }
  } else {
    for (int i = 0; i < nsync; ++i) __syncthreads();
  }
  const uint64_t t2 = gtimer();
  while (gtimer() - t2 < (uint64_t)spin_ns) {}
  if (threadIdx.x == 0) rec[blockIdx.x] = Rec{smid(), csize > 1 ? cluster_rank() : 0u, t0, t1, t2, gtimer()};
}

int main(int argc, char** argv) {
  const int grid = 1024;
  Rec* d;
  cudaMalloc(&d, grid * sizeof(Rec));
  float* src;
  cudaMalloc(&src, (size_t)grid * 64 * 1024);
  cudaMemset(src, 0, (size_t)grid * 64 * 1024);
  std::vector<Rec> h(grid);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(probe, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  printf("csize threads smemKB bytes nsync spin | SMs used, max CTAs co-resident/SM, CTAs/SM min..max, mates on same SM | first-wave CTAs | load med us | sync med us | total us | event us\n");
  for (int csize : {1, 2, 4, 8}) {
    for (int cfg = 0; cfg < 4; ++cfg) {
      const int threads = cfg == 3 ? 288 : 160;
      const int smem_kb = cfg == 0 ? 200 : cfg == 1 ? 100 : cfg == 2 ? 44 : 100;
      const int bytes = 25088, nsync = 2, spin = 2000;
      cudaLaunchConfig_t c{};
      c.gridDim = dim3(grid);
      c.blockDim = dim3(threads);
      c.dynamicSmemBytes = smem_kb * 1024;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = csize;
      at[0].val.clusterDim.y = 1;
      at[0].val.clusterDim.z = 1;
      c.attrs = at;
      c.numAttrs = 1;
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      float ms = 0;
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        cudaError_t e = cudaLaunchKernelEx(&c, probe, d, (const float*)src, bytes, spin, csize, nsync);
        cudaEventRecord(e1);
        if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); break; }
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("run failed: %s\n", cudaGetErrorString(e)); return 1; }
        cudaEventElapsedTime(&ms, e0, e1);
      }
      cudaMemcpy(h.data(), d, grid * sizeof(Rec), cudaMemcpyDeviceToHost);
      uint64_t tmin = ~0ull, tmax = 0;
      std::map<int, int> per_sm;
      for (auto& r : h) { tmin = std::min(tmin, r.t0); tmax = std::max(tmax, r.t2); per_sm[r.sm]++; }
      // co-residency: max overlap per SM
      int maxco = 0;
      for (auto& kv : per_sm) {
        std::vector<std::pair<uint64_t, int>> ev;
        for (auto& r : h) if ((int)r.sm == kv.first) { ev.push_back({r.t0, 1}); ev.push_back({r.t2, -1}); }
        std::sort(ev.begin(), ev.end());
        int cur = 0;
        for (auto& e : ev) { cur += e.second; maxco = std::max(maxco, cur); }
      }
      int mn = 1 << 30, mx = 0;
      for (auto& kv : per_sm) { mn = std::min(mn, kv.second); mx = std::max(mx, kv.second); }
      int same = 0;
      for (int i = 0; i + csize <= grid; i += csize)
        for (int j = 1; j < csize; ++j) if (h[i + j].sm == h[i].sm) { ++same; break; }
      int first = 0;
      for (auto& r : h) if (r.t0 - tmin < 1000) ++first;
      std::vector<double> ld, sy;
      for (auto& r : h) { ld.push_back((r.t1 - r.t0) * 1e-3); sy.push_back((r.ts - r.t1) * 1e-3); }
      std::sort(ld.begin(), ld.end());
      std::sort(sy.begin(), sy.end());
      printf("%d %d %d %d %d %d | %zu, %d, %d..%d, %d | %d | %.2f | %.2f | %.2f | %.2f\n", csize, threads, smem_kb, bytes, nsync, spin,
             per_sm.size(), maxco, mn, mx, same, first, ld[ld.size() / 2], sy[sy.size() / 2], (tmax - tmin) * 1e-3, ms * 1e3);
    }
  }
  return 0;
}
