// Streaming fused NFP kernels for sm_100a: cosine measure, stride 1, dilation 1, padding = R
// (the configuration every live model path of the reference uses: models/NFP_Pooling.py:10-16,
// models/texture_pooling.py:232,302).
//
// One CTA owns one image at a time (persistent loop over b = blockIdx.x, += gridDim.x).  Inside
// the CTA one PRODUCER warp streams the image's x as chunks of CC channels (CC x H x W contiguous
// elements of the NCHW tensor) into a shared-memory ring with TMA bulk copies (cp.async.bulk +
// mbarrier complete_tx); NW CONSUMER warps work on the chunks as they land:
//
//   pass A   per-pixel |x_p|^2 and the dot products with the (k*k-1)/2 "forward" window
//            neighbours (dot(p,q) == dot(q,p): half the window suffices), each lane owning a
//            TW-pixel row strip of one channel with the strip's accumulators in registers;
//            reduced over lanes (shuffles) and warps (shared memory) once per image.
//   forward  y = dot / (max(|p|,eps) max(|q|,eps)) for the K taps -> the only HBM write;
//            pooled mode reduces y and x over the plane instead (nfp_pooling head).
//   backward the table + gy become a per-pixel k x k stencil of coefficients Wd[p][o] (closed form
//            of ATen's cosine_similarity backward); then
//   pass B   gx[c][p] = sum_o Wd[p][o] * x[c][p+o], chunk by chunk.  If the whole image fits in
//            the ring ("resident") the chunks of pass A are still there; otherwise the producer
//            streams them a second time -- they were read microseconds ago by the same SM, so
//            the second read is served by the 126 MB L2, not by HBM.  Each warp stages its planes
//            in shared memory and writes them with 16-byte fully coalesced stores.
//
// The (B, C*(k*k-1), H, W) neighbour tensor of the reference (nfp.py:153-154) never exists, x is
// read from HBM once per kernel, and no cluster / grid synchronisation is needed: images are
// independent, all B CTAs are resident at once (2 per SM), so loads of all images are in flight
// together and HBM stays saturated while individual CTAs sit in their reduction phases.
#include <stdlib.h>

#include "nfp_common.cuh"

namespace nfp {
namespace {

enum { MODE_FWD = 0, MODE_BWD = 1, MODE_POOL_FWD = 2, MODE_POOL_BWD = 3 };

struct StreamArgs {
  const void* x;
  const void* gy;
  void* y;
  void* gx;
  const float* g_gap_x;
  const float* g_gap_nfp;
  float* gap_x;
  float* gap_nfp;
  int B, C;
  int CC;        // channels per chunk (multiple of CPW, divides C)
  int NCH;       // chunks per image
  int resident;  // backward: the image fits in the ring, pass B re-walks the slots of pass A
  int mode, pad_mode, similarity;
  float eps;
};

template <int H_, int W_, int R_, int TW_>
struct Cfg {
  static constexpr int H = H_, W = W_, R = R_, TW = TW_;
  static constexpr int k = 2 * R + 1, KK = k * k, K = KK - 1, CTR = R * k + R;
  static constexpr int P = H * W;
  static constexpr int NSX = W / TW;        // strips per row
  static constexpr int NS = H * NSX;        // strips per channel plane
  static constexpr int ND = K / 2;          // forward directions
  static constexpr int NV = ND + 1;         // table entries per pixel: |x|^2 + ND dots
  static constexpr int CPW = 32 / NS;       // channels per warp iteration
  static constexpr int LANES = CPW * NS;    // active lanes
  static constexpr int XW = (NSX == 1) ? TW : TW + 2 * R;  // loaded columns per row (halo only if strips abut)
  static constexpr int XOFF = (NSX == 1) ? 0 : R;          // column index of strip pixel 0 inside a loaded row
  static_assert(W % TW == 0, "strip width must divide W");
  static_assert(NS <= 32 && CPW >= 1 && (CPW & (CPW - 1)) == 0, "channels per warp must be a power of two");
};

// ---- PTX helpers: mbarrier, TMA bulk copy, named barriers ---------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
template <int NTHREADS>
__device__ __forceinline__ void consumer_sync() {  // named barrier 1: the consumer warps only
  asm volatile("bar.sync 1, %0;" ::"n"(NTHREADS) : "memory");
}

template <typename T> __device__ __forceinline__ float lds(const T* p);
template <> __device__ __forceinline__ float lds<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float lds<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __uint_as_float(((uint32_t) * reinterpret_cast<const unsigned short*>(p)) << 16);
}

// dot(x_p, x_q) from the symmetric table; (dy, dx) = q - p, both within the window
template <class C>
__device__ __forceinline__ float table_dot(const float* tab, int p, int q, int dy, int dx) {
  int o = (dy + C::R) * C::k + dx + C::R;
  if (o == C::CTR) return tab[p * C::NV];
  return o > C::CTR ? tab[p * C::NV + (o - C::CTR)] : tab[q * C::NV + (C::CTR - o)];
}

constexpr int kMaxStages = 8;

template <typename T, class C, int NW>
struct Smem {
  // byte offsets from the (1024-aligned) start of dynamic shared memory
  int ring, gyraw, tfull, uni, bars, total;
  int wtab, gyS, wd, inv, rn, selfw, stg, ytab;  // inside the union region
  int slot_bytes;
  __host__ __device__ Smem(int CC, int nst) {
    int o = 0;
    auto take = [&](int n) { int r = o; o += (n + 127) & ~127; return r; };
    slot_bytes = (CC * C::P * (int)sizeof(T) + 127) & ~127;
    ring = take(nst * slot_bytes);
    gyraw = take(2 * ((C::K * C::P * (int)sizeof(T) + 15) & ~15));
    tfull = take(C::P * C::NV * 4);
    bars = take((2 * kMaxStages + 4) * 8);
    uni = o;
    wtab = take(NW * C::P * C::NV * 4);
    const int u1 = o;
    o = uni;
    gyS = take(C::K * C::P * 4);
    wd = take(C::P * C::KK * 4);
    inv = take(C::P * 4);
    rn = take(C::P * 4);
    selfw = take(C::P * 4);
    stg = take(NW * C::LANES * C::TW * 4);
    const int u2 = o;
    o = uni;
    ytab = take(C::K * C::P * 4);
    const int u3 = o;
    total = u1 > u2 ? (u1 > u3 ? u1 : u3) : (u2 > u3 ? u2 : u3);
  }
};

template <typename T, class C, int NW, int MINB>
__global__ void __launch_bounds__((NW + 1) * 32, MINB) stream_kernel(StreamArgs a, int nst) {
  constexpr int H = C::H, W = C::W, R = C::R, TW = C::TW, k = C::k, KK = C::KK, K = C::K, P = C::P;
  constexpr int NV = C::NV, NS = C::NS, NSX = C::NSX, CPW = C::CPW, LANES = C::LANES, XW = C::XW, XOFF = C::XOFF;
  constexpr int NT = NW * 32;  // consumer threads
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const Smem<T, C, NW> L(a.CC, nst);
  unsigned char* ring = smem_raw + L.ring;
  float* tfull = reinterpret_cast<float*>(smem_raw + L.tfull);
  float* wtab = reinterpret_cast<float*>(smem_raw + L.wtab);
  float* gyS = reinterpret_cast<float*>(smem_raw + L.gyS);
  float* Wd = reinterpret_cast<float*>(smem_raw + L.wd);
  float* inv = reinterpret_cast<float*>(smem_raw + L.inv);
  float* rn = reinterpret_cast<float*>(smem_raw + L.rn);
  float* selfw = reinterpret_cast<float*>(smem_raw + L.selfw);
  float* stg = reinterpret_cast<float*>(smem_raw + L.stg);
  float* ytab = reinterpret_cast<float*>(smem_raw + L.ytab);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + L.bars);
  uint64_t* empty = full + kMaxStages;
  uint64_t* gyfull = empty + kMaxStages;
  uint64_t* gyempty = gyfull + 2;
  const int gy_bytes = K * P * (int)sizeof(T);
  const int gy_stride = (gy_bytes + 15) & ~15;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool bwd = (a.mode == MODE_BWD || a.mode == MODE_POOL_BWD);
  const bool pooled = (a.mode == MODE_POOL_FWD || a.mode == MODE_POOL_BWD);
  const bool gy_tma = (a.mode == MODE_BWD);
  const bool resident = bwd && a.resident;
  const int NCH = a.NCH, CC = a.CC;
  const uint32_t chunk_bytes = (uint32_t)(CC * P * (int)sizeof(T));

  if (tid == 0) {
    for (int s = 0; s < nst; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], NW);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&gyfull[s], 1);
      mbar_init(&gyempty[s], NW);
    }
    fence_mbar_init();
  }
  __syncthreads();

  // ================================ producer warp ==================================================
  if (warp == NW) {
    if (lane == 0) {
      uint32_t seq = 0;
      int img = 0;
      const int npass = (bwd && !resident) ? 2 : 1;
      for (int b = blockIdx.x; b < a.B; b += gridDim.x, ++img) {
        if (gy_tma) {
          const int par = img & 1;
          mbar_wait(&gyempty[par], ((img >> 1) & 1) ^ 1);
          mbar_expect_tx(&gyfull[par], (uint32_t)gy_bytes);
          bulk_g2s(smem_raw + L.gyraw + par * gy_stride, reinterpret_cast<const T*>(a.gy) + (size_t)b * K * P,
                   (uint32_t)gy_bytes, &gyfull[par]);
        }
        const T* xb = reinterpret_cast<const T*>(a.x) + (size_t)b * a.C * P;
        for (int pass = 0; pass < npass; ++pass)
          for (int ch = 0; ch < NCH; ++ch, ++seq) {
            const int slot = seq % nst;
            const uint32_t ph = (seq / nst) & 1;
            mbar_wait(&empty[slot], ph ^ 1);
            mbar_expect_tx(&full[slot], chunk_bytes);
            bulk_g2s(ring + (size_t)slot * L.slot_bytes, xb + (size_t)ch * CC * P, chunk_bytes, &full[slot]);
          }
      }
    }
    return;
  }

  // ================================ consumer warps =================================================
  const float sgn = a.similarity ? 1.f : -1.f;
  const bool lane_on = lane < LANES;
  const int chslot = lane_on ? lane / NS : 0;
  const int pos = lane_on ? lane % NS : 0;
  const int r = pos / NSX, c0 = (pos % NSX) * TW;
  const int strip_off = r * W + c0;
  int coff[XW];  // column offsets (relative to c0) of the loaded row window, clamped into the map
#pragma unroll
  for (int jj = 0; jj < XW; ++jj) {
    int c = c0 + jj - XOFF;
    c = c < 0 ? 0 : (c > W - 1 ? W - 1 : c);
    coff[jj] = c - c0;
  }
  const int ngroups = CC / CPW;
  auto row_of = [&](int p) { return p / W; };
  auto col_of = [&](int p) { return p - (p / W) * W; };

  uint32_t seq = 0;
  int img = 0;
  for (int b = blockIdx.x; b < a.B; b += gridDim.x, ++img) {
    const size_t img_off = (size_t)b * a.C * P;
    const uint32_t seq0 = seq;

    // ---- pass A: per-pixel |x|^2 and forward-direction dots, streamed over the chunks -----------
    {
      float acc[TW][NV];
#pragma unroll
      for (int j = 0; j < TW; ++j)
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[j][v] = 0.f;
      int roff[R + 1];
#pragma unroll
      for (int dy = 0; dy <= R; ++dy) roff[dy] = ((r + dy > H - 1 ? H - 1 : r + dy) - r) * W;

      for (int ch = 0; ch < NCH; ++ch, ++seq) {
        const int slot = seq % nst;
        mbar_wait(&full[slot], (seq / nst) & 1);
        const T* sl = reinterpret_cast<const T*>(ring + (size_t)slot * L.slot_bytes);
        if (a.mode == MODE_POOL_FWD) {
          // GAP(x) of this chunk's channels (NFP_Pooling.py:27): one lane per channel plane; the
          // job rotates over the warps chunk by chunk
          const int w0 = (ch * 2) % NW;
          for (int c = ((warp - w0 + NW) % NW) * 32 + lane; c < CC; c += NT) {
            const T* pl = sl + c * P;
            float s = 0.f;
#pragma unroll 7
            for (int e = 0; e < P; ++e) s += lds<T>(pl + e);
            a.gap_x[(size_t)b * a.C + ch * CC + c] = s / (float)P;
          }
        }
#pragma unroll 2
        for (int g = warp; g < ngroups; g += NW) {
          if (lane_on) {
            const T* base = sl + (g * CPW + chslot) * P + strip_off;
            float xr[R + 1][XW];
#pragma unroll
            for (int dy = 0; dy <= R; ++dy)
#pragma unroll
              for (int jj = 0; jj < XW; ++jj) xr[dy][jj] = lds<T>(base + roff[dy] + coff[jj]);
#pragma unroll
            for (int j = 0; j < TW; ++j) {
              const float c = xr[0][j + XOFF];
              acc[j][0] = fmaf(c, c, acc[j][0]);
#pragma unroll
              for (int dx = 1; dx <= R; ++dx) {
                if (j + dx + XOFF < XW) acc[j][dx] = fmaf(c, xr[0][j + dx + XOFF], acc[j][dx]);
              }
#pragma unroll
              for (int dy = 1; dy <= R; ++dy)
#pragma unroll
                for (int dx = -R; dx <= R; ++dx) {
                  if (j + dx + XOFF >= 0 && j + dx + XOFF < XW)
                    acc[j][dy * k + dx] = fmaf(c, xr[dy][j + dx + XOFF], acc[j][dy * k + dx]);
                }
            }
          }
        }
        if (!resident) {
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[slot]);
        }
      }
      // reduce over the channel slots inside the warp, then publish one table per warp
#pragma unroll
      for (int j = 0; j < TW; ++j)
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          float s = acc[j][v];
#pragma unroll
          for (int d = CPW / 2; d >= 1; d >>= 1) s += __shfl_down_sync(0xffffffffu, s, d * NS);
          if (lane < NS) wtab[warp * (P * NV) + (pos * TW + j) * NV + v] = s;
        }
    }
    consumer_sync<NT>();
    for (int i = tid; i < P * NV; i += NT) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) s += wtab[w * (P * NV) + i];
      tfull[i] = s;
    }
    consumer_sync<NT>();

    // ---- forward value -----------------------------------------------------------------------------
    if (!bwd) {
      for (int idx = tid; idx < K * P; idx += NT) {
        const int n = idx / P, p = idx - n * P;
        int ta, tb;
        tap_rc(n, k, K, ta, tb);
        const int pr = row_of(p), pc = col_of(p);
        const int qr = map_index(pr + ta - R, H, a.pad_mode), qc = map_index(pc + tb - R, W, a.pad_mode);
        float yv = 0.f;
        if (qr >= 0 && qc >= 0) {
          const int q = qr * W + qc;
          const float d = table_dot<C>(tfull, p, q, qr - pr, qc - pc);
          const float Np = fmaxf(sqrtf(tfull[p * NV]), a.eps), Nq = fmaxf(sqrtf(tfull[q * NV]), a.eps);
          yv = d / (Np * Nq);
        }
        if (!a.similarity) yv = 1.f - yv;
        if (pooled) {
          ytab[idx] = yv;
        } else {
          reinterpret_cast<T*>(a.y)[(size_t)b * K * P + idx] = from_f32<T>(yv);
        }
      }
      if (pooled) {
        // GAP over the plane of every tap (NFP_Pooling.py:31)
        consumer_sync<NT>();
        for (int n = warp; n < K; n += NW) {
          float s = 0.f;
          for (int p = lane; p < P; p += 32) s += ytab[n * P + p];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
          if (lane == 0) a.gap_nfp[(size_t)b * K + n] = s / (float)P;
        }
      }
      consumer_sync<NT>();  // tfull / ytab / wtab are rewritten by the next image
      continue;
    }

    // ---- backward: stencil coefficients ------------------------------------------------------------
    for (int p = tid; p < P; p += NT) {
      const float nsq = tfull[p * NV];
      const float nrm = sqrtf(nsq), N = fmaxf(nrm, a.eps);
      inv[p] = 1.f / N;
      rn[p] = nrm > 0.f ? 1.f / (N * nrm) : 0.f;
      selfw[p] = 0.f;
    }
    if (pooled) {
      for (int idx = tid; idx < K * P; idx += NT)
        gyS[idx] = sgn * a.g_gap_nfp[(size_t)b * K + idx / P] * (1.f / (float)P);
    } else {
      const int par = img & 1;
      mbar_wait(&gyfull[par], (img >> 1) & 1);
      const T* g = reinterpret_cast<const T*>(smem_raw + L.gyraw + par * gy_stride);
      for (int idx = tid; idx < K * P; idx += NT) gyS[idx] = sgn * lds<T>(g + idx);
      __syncwarp();
      if (lane == 0) mbar_arrive(&gyempty[par]);
    }
    consumer_sync<NT>();
    // direct (in-map) pairs: Wd[p][o] = (G[n_o][p] + G[~n_o][q]) / (N_p N_q), q = p + off(o)
    for (int idx = tid; idx < P * KK; idx += NT) {
      const int p = idx / KK, o = idx - p * KK;
      float wv = 0.f;
      if (o != C::CTR) {
        const int dy = o / k - R, dx = o % k - R;
        const int qr = row_of(p) + dy, qc = col_of(p) + dx;
        if (qr >= 0 && qr < H && qc >= 0 && qc < W) {
          const int q = qr * W + qc;
          const int n = o < C::CTR ? o : o - 1;
          wv = (gyS[n * P + p] + gyS[(K - 1 - n) * P + q]) * inv[p] * inv[q];
        }
      }
      Wd[idx] = wv;
    }
    consumer_sync<NT>();
    // padded taps: the neighbour is a reflected / replicated in-map pixel v
    if (a.pad_mode != NFPB200_PAD_ZEROS) {
      for (int idx = tid; idx < K * P; idx += NT) {
        const int n = idx / P, p = idx - n * P;
        int ta, tb;
        tap_rc(n, k, K, ta, tb);
        const int pr = row_of(p), pc = col_of(p);
        const int rr = pr + ta - R, cc = pc + tb - R;
        if (rr >= 0 && rr < H && cc >= 0 && cc < W) continue;
        const int vr = map_index(rr, H, a.pad_mode), vc = map_index(cc, W, a.pad_mode);
        const int v = vr * W + vc;
        const float wv = gyS[idx] * inv[p] * inv[v];
        if (v == p) {
          atomicAdd(&selfw[p], 2.f * wv);
        } else {
          atomicAdd(&Wd[p * KK + (vr - pr + R) * k + (vc - pc + R)], wv);
          atomicAdd(&Wd[v * KK + (pr - vr + R) * k + (pc - vc + R)], wv);
        }
      }
      consumer_sync<NT>();
    }
    // centre tap: -(1/(N_p |x_p|)) * sum_o Wd[p][o] dot(p, q_o)   (+ the self pairs)
    for (int p = tid; p < P; p += NT) {
      float s = 0.f;
      const int pr = row_of(p), pc = col_of(p);
#pragma unroll
      for (int o = 0; o < KK; ++o) {
        if (o == C::CTR) continue;
        const int dy = o / k - R, dx = o % k - R;
        const int qr = pr + dy, qc = pc + dx;
        if (qr >= 0 && qr < H && qc >= 0 && qc < W)
          s = fmaf(Wd[p * KK + o], table_dot<C>(tfull, p, qr * W + qc, dy, dx), s);
      }
      const float sw = selfw[p];
      Wd[p * KK + C::CTR] = sw - rn[p] * (s + sw * tfull[p * NV]);
    }
    consumer_sync<NT>();

    // ---- pass B: gx = stencil(x), chunk by chunk ------------------------------------------------------
    {
      float wr[TW][KK];
#pragma unroll
      for (int j = 0; j < TW; ++j)
#pragma unroll
        for (int o = 0; o < KK; ++o) wr[j][o] = Wd[(pos * TW + j) * KK + o];
      int roff[k];
#pragma unroll
      for (int dy = -R; dy <= R; ++dy) {
        int rr = r + dy;
        rr = rr < 0 ? 0 : (rr > H - 1 ? H - 1 : rr);
        roff[dy + R] = (rr - r) * W;
      }
      float* mystg = stg + warp * (LANES * TW);
      const float invP = 1.f / (float)P;
      T* gxg = reinterpret_cast<T*>(a.gx) + img_off;
      if (resident) seq = seq0;  // re-walk the slots pass A left in place
      for (int ch = 0; ch < NCH; ++ch, ++seq) {
        const int slot = seq % nst;
        mbar_wait(&full[slot], (seq / nst) & 1);
        const T* sl = reinterpret_cast<const T*>(ring + (size_t)slot * L.slot_bytes);
        for (int g = warp; g < ngroups; g += NW) {
          const int ch_w = g * CPW;  // first channel of this warp's group inside the chunk
          if (lane_on) {
            const T* base = sl + (ch_w + chslot) * P + strip_off;
            float out[TW];
            const float g0 = a.mode == MODE_POOL_BWD
                                 ? a.g_gap_x[(size_t)b * a.C + ch * CC + ch_w + chslot] * invP
                                 : 0.f;
#pragma unroll
            for (int j = 0; j < TW; ++j) out[j] = g0;
#pragma unroll
            for (int dy = -R; dy <= R; ++dy) {
              float xr[XW];
#pragma unroll
              for (int jj = 0; jj < XW; ++jj) xr[jj] = lds<T>(base + roff[dy + R] + coff[jj]);
#pragma unroll
              for (int j = 0; j < TW; ++j)
#pragma unroll
                for (int dx = -R; dx <= R; ++dx) {
                  if (j + dx + XOFF >= 0 && j + dx + XOFF < XW)
                    out[j] = fmaf(wr[j][(dy + R) * k + dx + R], xr[j + dx + XOFF], out[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < TW; ++j) mystg[lane * TW + j] = out[j];
          }
          __syncwarp();
          // the warp's LANES*TW outputs are contiguous in gx: channels [ch_w, ch_w + CPW) of the chunk
          constexpr int NOUT = LANES * TW;
          static_assert(NOUT % 4 == 0, "");
          T* dst = gxg + ((size_t)ch * CC + ch_w) * P;
          if constexpr (sizeof(T) == 4) {
            for (int i = lane; i < NOUT / 4; i += 32)
              reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(mystg)[i];
          } else {
            for (int i = lane; i < NOUT / 4; i += 32) {
              const float4 v = reinterpret_cast<const float4*>(mystg)[i];
              __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
              uint2 pk;
              pk.x = *reinterpret_cast<uint32_t*>(&lo);
              pk.y = *reinterpret_cast<uint32_t*>(&hi);
              reinterpret_cast<uint2*>(dst)[i] = pk;
            }
          }
          __syncwarp();
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[slot]);
      }
    }
    consumer_sync<NT>();  // the union region (Wd / stg) is rewritten as wtab by the next image
  }
}

// ---- host side --------------------------------------------------------------------------------------

constexpr int kNW = 8;
constexpr int kSmemPerSM = 227 * 1024;

struct Plan {
  bool ok;
  int CC, NCH, nst, resident, ctas_per_sm;
  size_t smem;
};

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

template <typename T, class C>
Plan plan_for(const KParams& P, int op) {
  Plan pl{false, 0, 0, 0, 0, 0, 0};
  const int esz = (int)sizeof(T);
  const bool bwd = (op == NFPB200_OP_BACKWARD || op == NFPB200_OP_POOL_BACKWARD);
  // chunk: the largest CC <= target that divides C, is a multiple of CPW, keeps 16-byte TMA granularity
  // and 16-byte (fp32) / 8-byte (bf16) vector stores of whole warp groups
  static const int target_bytes = env_int("NFPB200_CHUNK_BYTES", 16 * 1024);
  static const int want_ctas = env_int("NFPB200_CTAS_PER_SM", 2);
  static const int force_stream = env_int("NFPB200_NO_RESIDENT", 0);
  if (((size_t)C::CPW * C::P * esz) % (esz == 4 ? 16 : 8)) return pl;
  if (((size_t)C::K * C::P * esz) % 16) return pl;
  int best = 0;
  for (int cc = C::CPW; cc <= P.C; cc += C::CPW) {
    if (P.C % cc) continue;
    if (((size_t)cc * C::P * esz) % 16) continue;
    if ((size_t)cc * C::P * esz > (size_t)target_bytes && best) break;
    best = cc;
    if ((size_t)cc * C::P * esz >= (size_t)target_bytes) break;
  }
  if (!best) return pl;
  pl.CC = best;
  pl.NCH = P.C / best;
  for (int ctas = want_ctas; ctas >= 1 && !pl.ok; --ctas) {
    const int budget = kSmemPerSM / ctas - 1024;
    for (int nst = kMaxStages; nst >= 2; --nst) {  // as many stages as fit: deeper prefetch across images
      Smem<T, C, kNW> L(pl.CC, nst);
      if (L.total + 1024 > budget) continue;
      pl.nst = nst;
      pl.smem = (size_t)L.total + 1024;  // slack for the 1024-byte alignment of the ring
      pl.ctas_per_sm = ctas;
      pl.ok = true;
      break;
    }
  }
  if (!pl.ok) return pl;
  pl.resident = (bwd && !force_stream && pl.NCH <= pl.nst) ? 1 : 0;
  return pl;
}

template <typename T, class C>
int launch_t(const KParams& P, StreamArgs a, int op, cudaStream_t stream) {
  Plan pl = plan_for<T, C>(P, op);
  if (!pl.ok) return NFPB200_EUNSUPPORTED;
  a.CC = pl.CC;
  a.NCH = pl.NCH;
  a.resident = pl.resident;
  auto kern = stream_kernel<T, C, kNW, 2>;
  static const cudaError_t attr_rc =
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemPerSM);
  if (attr_rc != cudaSuccess) return (int)attr_rc;
  static const int num_sms = [] {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
  }();
  const int slots = num_sms * pl.ctas_per_sm;
  const int grid = P.B < slots ? P.B : slots;
  stream_kernel<T, C, kNW, 2><<<grid, (kNW + 1) * 32, pl.smem, stream>>>(a, pl.nst);
  return (int)cudaGetLastError();
}

// the (H, W, R) shapes with a streaming instantiation
#define NFP_STREAM_SHAPES(X) \
  X(7, 7, 1, 7)              \
  X(14, 14, 1, 7)            \
  X(2, 2, 1, 2)              \
  X(4, 4, 1, 4)

template <typename F>
bool for_shape(const KParams& P, F&& f) {
#define X(H_, W_, R_, TW_) \
  if (P.H == H_ && P.W == W_ && P.R == R_) { f(Cfg<H_, W_, R_, TW_>{}); return true; }
  NFP_STREAM_SHAPES(X)
#undef X
  return false;
}

bool geometry_ok(const KParams& P, int measure) {
  return measure == NFPB200_COSINE && P.stride == 1 && P.dil == 1 && P.pad == P.R &&
         P.mode != NFPB200_PAD_CIRCULAR;
}

int run(const KParams& P, int dtype, int op, StreamArgs a, cudaStream_t stream) {
  a.B = P.B; a.C = P.C;
  a.pad_mode = P.mode; a.similarity = P.similarity; a.eps = P.eps;
  int rc = NFPB200_EUNSUPPORTED;
  for_shape(P, [&](auto cfg) {
    using C = decltype(cfg);
    rc = dtype == NFPB200_BF16 ? launch_t<__nv_bfloat16, C>(P, a, op, stream) : launch_t<float, C>(P, a, op, stream);
  });
  return rc;
}

}  // namespace

bool stream_supported(const KParams& P, int dtype, int measure, int op) {
  if (!geometry_ok(P, measure)) return false;
  bool ok = false;
  for_shape(P, [&](auto cfg) {
    using C = decltype(cfg);
    ok = dtype == NFPB200_BF16 ? plan_for<__nv_bfloat16, C>(P, op).ok : plan_for<float, C>(P, op).ok;
  });
  return ok;
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
  StreamArgs a{};
  a.x = x; a.y = y; a.mode = MODE_FWD;
  return run(P, dtype, NFPB200_OP_FORWARD, a, ctx.stream);
}
int stream_backward(const KParams& P, int dtype, const void* x, const void* gy, void* gx, const LaunchCtx& ctx) {
  StreamArgs a{};
  a.x = x; a.gy = gy; a.gx = gx; a.mode = MODE_BWD;
  return run(P, dtype, NFPB200_OP_BACKWARD, a, ctx.stream);
}
int stream_pool_forward(const KParams& P, int dtype, const void* x, float* gap_x, float* gap_nfp,
                        const LaunchCtx& ctx) {
  StreamArgs a{};
  a.x = x; a.gap_x = gap_x; a.gap_nfp = gap_nfp; a.mode = MODE_POOL_FWD;
  return run(P, dtype, NFPB200_OP_POOL_FORWARD, a, ctx.stream);
}
int stream_pool_backward(const KParams& P, int dtype, const void* x, const float* g_gap_x, const float* g_gap_nfp,
                         void* gx, const LaunchCtx& ctx) {
  StreamArgs a{};
  a.x = x; a.g_gap_x = g_gap_x; a.g_gap_nfp = g_gap_nfp; a.gx = gx; a.mode = MODE_POOL_BWD;
  return run(P, dtype, NFPB200_OP_POOL_BACKWARD, a, ctx.stream);
}

}  // namespace nfp
