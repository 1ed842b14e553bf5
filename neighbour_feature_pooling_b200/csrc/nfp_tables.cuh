// Compile-time geometry (Cfg) and stencil index tables (Tables) shared by the fused NFP kernels
// (nfp_stream_impl.cuh: one CTA per image; nfp_split_impl.cuh: one cluster per image).
// The index arithmetic of the stencil (reflect / replicate / zero padding, tap order of nfp.py:64-67) is evaluated at
// COMPILE time into __device__ const tables.
#pragma once

#include <stdint.h>

#include "nfp_common.cuh"

namespace nfp {
namespace stream {

__host__ __device__ constexpr int align_up(int n, int a) { return (n + a - 1) / a * a; }

template <int H_, int W_, int R_, int TW_>
struct Cfg {
  static constexpr int H = H_, W = W_, R = R_, TW = TW_;
  static constexpr int k = 2 * R + 1, KK = k * k, K = KK - 1, CTR = R * k + R;
  static constexpr int P = H * W;
  static constexpr int NSX = W / TW;        // strips per row
  static constexpr int NS = H * NSX;        // strips per channel plane
  static constexpr int ND = K / 2;          // forward directions
  static constexpr int NV = ND + 1;         // table entries per pixel: |x|^2 + ND dots
  static constexpr int PNV = P * NV;
  static constexpr int CPW = 32 / NS;       // channels per group (one warp pass)
  static constexpr int LANES = CPW * NS;    // active lanes
  static constexpr int XW = (NSX == 1) ? TW : TW + 2 * R;  // loaded columns per row (halo only if strips abut)
  static constexpr int XOFF = (NSX == 1) ? 0 : R;          // column index of strip pixel 0 inside a loaded row
  static constexpr int HALO = R * W + R;    // elements a strip may read before / after its channel plane
  static constexpr bool PACK = (R == 1 && NSX == 1);  // pass A on packed fp32 pairs where the 96-register budget allows it
  static constexpr int MINB = (R == 1) ? 2 : 1;  // CTAs per SM the register budget is sized for
  // coefficient table Wd: one row of RS floats per map row (W pixels x KK offsets); padded to whole float4s where a
  // lane strip is a full row, so the lane-per-channel pass B can fetch a row's coefficients with broadcast LDS.128
  // (an ODD number of float4s per row: rows start 16-byte aligned and in different banks, so the strip form's
  // per-lane coefficient loads stay conflict-free; other shapes keep the dense p*KK + o layout)
  static constexpr bool LANECH = (NSX == 1 && P >= 49);  // shapes with a lane-per-channel pass B (measured: no gain on 2x2)
  static constexpr int RS4 = align_up(W * KK, 4) / 4;
  static constexpr int RS = LANECH ? 4 * (RS4 % 2 ? RS4 : RS4 + 1) : W * KK;
  static constexpr int TASK = 64;                        // its work item: 64 channels = 2 per lane
  __host__ __device__ static constexpr int widx(int p, int o) { return (p / W) * RS + (p % W) * KK + o; }
  static_assert(W % TW == 0, "strip width must divide W");
  static_assert(HALO * 4 <= 128, "the zeroed lead pad in front of the ring must cover the halo");
  static_assert(NS <= 32 && CPW >= 1 && (CPW & (CPW - 1)) == 0, "channels per group must be a power of two");
};

// ---- compile-time stencil tables ----------------------------------------------------------------

// taps that point outside the map (they fold back onto a window pixel under reflect / replicate padding)
template <class C>
constexpr int count_outside_taps() {
  int n = 0;
  for (int p = 0; p < C::P; ++p)
    for (int o = 0; o < C::KK; ++o) {
      if (o == C::CTR) continue;
      const int qr = p / C::W + o / C::k - C::R, qc = p % C::W + o % C::k - C::R;
      if (qr < 0 || qr >= C::H || qc < 0 || qc >= C::W) ++n;
    }
  return n;
}

template <class C>
struct Tables {
  // every array padded to a multiple of 16 bytes: the kernels fetch [fv, fd] (forward) or [q, fsrc, fdst, fptr]
  // (backward) with one TMA bulk copy
  static constexpr int NF = align_up(C::K * C::P, 8);
  static constexpr int NQ = align_up(C::P * C::KK, 8);
  static constexpr int NOUT = count_outside_taps<C>();
  static constexpr int NFS = align_up(NOUT + 1, 8);
  static constexpr int NFP = align_up(NOUT + 2, 8);
  static constexpr int FWD_BYTES = 2 * NF * 2, BWD_BYTES = (NQ + 2 * NFS + NFP) * 2, BWD_OFFSET = FWD_BYTES;
  alignas(16) int16_t fv[NF];   // forward: pixel the (padded) tap n of pixel p lands on, -1 = implicit zero
  alignas(16) int16_t fd[NF];   // forward: index into the table of dot(p, fv)
  alignas(16) int16_t q[NQ];    // window pixel p + off(o) when inside the map, else -1
  // backward, folded taps in CSR form: window entry fdst[i] = p*KK + o of a border pixel p additionally receives
  // the upstream gradient elements fsrc[fptr[i] .. fptr[i+1]) (flat n*P + p) of p's taps that point outside the
  // map and are folded onto p + off(o) by the padding (o == CTR: onto p itself); fptr[NFP-1] = number of entries
  alignas(16) int16_t fsrc[NFS];
  alignas(16) int16_t fdst[NFS];
  alignas(16) int16_t fptr[NFP];
};

constexpr int cmap_index(int i, int n, int mode) {
  if (i >= 0 && i < n) return i;
  if (mode == NFPB200_PAD_REFLECT) return i < 0 ? -i : 2 * (n - 1) - i;
  if (mode == NFPB200_PAD_REPLICATE) return i < 0 ? 0 : n - 1;
  return -1;
}

template <class C>
constexpr Tables<C> make_tables(int mode) {
  Tables<C> t{};
  for (int p = 0; p < C::P; ++p)
    for (int o = 0; o < C::KK; ++o) {
      const int qr = p / C::W + o / C::k - C::R, qc = p % C::W + o % C::k - C::R;
      t.q[p * C::KK + o] = (qr >= 0 && qr < C::H && qc >= 0 && qc < C::W) ? (int16_t)(qr * C::W + qc) : (int16_t)-1;
    }
  for (int n = 0; n < C::K; ++n) {
    const int tt = n < (C::K >> 1) ? n : n + 1;  // row-major window with the centre removed (nfp.py:64-67)
    const int ta = tt / C::k, tb = tt % C::k;
    for (int p = 0; p < C::P; ++p) {
      const int pr = p / C::W, pc = p % C::W;
      const int vr = cmap_index(pr + ta - C::R, C::H, mode), vc = cmap_index(pc + tb - C::R, C::W, mode);
      if (vr < 0 || vc < 0) {
        t.fv[n * C::P + p] = -1;
        t.fd[n * C::P + p] = 0;
        continue;
      }
      const int v = vr * C::W + vc;
      const int o = (vr - pr + C::R) * C::k + (vc - pc + C::R);
      t.fv[n * C::P + p] = (int16_t)v;
      t.fd[n * C::P + p] = (int16_t)(o == C::CTR ? p * C::NV
                                                 : (o > C::CTR ? p * C::NV + (o - C::CTR) : v * C::NV + (C::CTR - o)));
    }
  }
  // folded taps, grouped by the window entry they land on (border pixels only)
  int nfd = 0, nfs = 0;
  for (int p = 0; p < C::P; ++p) {
    const int pr = p / C::W, pc = p % C::W;
    if (pr >= C::R && pr < C::H - C::R && pc >= C::R && pc < C::W - C::R) continue;  // no tap leaves the map
    int land[C::K] = {};  // window entry the outside tap n folds onto, -1 = none
    for (int n = 0; n < C::K; ++n) {
      const int tt = n < (C::K >> 1) ? n : n + 1;
      const int rr = pr + tt / C::k - C::R, cc = pc + tt % C::k - C::R;
      land[n] = -1;
      if (rr >= 0 && rr < C::H && cc >= 0 && cc < C::W) continue;  // a direct tap
      const int vr = cmap_index(rr, C::H, mode), vc = cmap_index(cc, C::W, mode);
      if (vr < 0 || vc < 0) continue;  // zero padding
      land[n] = (vr - pr + C::R) * C::k + (vc - pc + C::R);
    }
    for (int o = 0; o < C::KK; ++o) {
      int cnt = 0;
      for (int n = 0; n < C::K; ++n)
        if (land[n] == o) {
          if (cnt == 0) {
            t.fdst[nfd] = (int16_t)(p * C::KK + o);
            t.fptr[nfd] = (int16_t)nfs;
          }
          t.fsrc[nfs++] = (int16_t)(n * C::P + p);
          ++cnt;
        }
      if (cnt) ++nfd;
    }
  }
  t.fptr[nfd] = (int16_t)nfs;
  t.fptr[Tables<C>::NFP - 1] = (int16_t)nfd;
  return t;
}

template <class C, int PADMODE>
__device__ const Tables<C> g_tables = make_tables<C>(PADMODE);

template <class C>
const Tables<C>* tables_for(int pad_mode) {
  const Tables<C>* p = nullptr;
  cudaError_t e;
  switch (pad_mode) {
    case NFPB200_PAD_REFLECT: e = cudaGetSymbolAddress((void**)&p, g_tables<C, NFPB200_PAD_REFLECT>); break;
    case NFPB200_PAD_REPLICATE: e = cudaGetSymbolAddress((void**)&p, g_tables<C, NFPB200_PAD_REPLICATE>); break;
    default: e = cudaGetSymbolAddress((void**)&p, g_tables<C, NFPB200_PAD_ZEROS>); break;
  }
  return e == cudaSuccess ? p : nullptr;
}

}  // namespace stream
}  // namespace nfp
