// Warp-level tensor-core convolutions for the 16/32-channel levels of UNet-B (networks/unet.py: n = 16 base
// width), bf16 in / fp32 accumulate.
//
// With Cin or Cout in {16, 32} a tcgen05 tile (64-channel K chunks, N >= 64) would be 75-94 % padding and these
// layers are HBM-bound anyway (arithmetic intensity 72-144 FLOP/B); the generic CUDA-core kernel
// (simt_conv.cu) ran them at 4-6 TFLOP/s and made UNet-B 4.6x SLOWER than the 9x larger UNet-A.  Here a warp owns
// 16 consecutive pixels of an image row:
//   k_conv_mid_mma  : forward / dgrad.  A fragments are gathered tap by tap with predicated 4-byte loads (two
//                     channels; the 3x3 neighbourhood of a 16-pixel run is L1-resident), the packed weights
//                     [Cout][K] sit in shared memory in B-fragment order, mma.sync.m16n8k16 accumulates the
//                     16 x Cout tile, which is transposed through a warp-private shared tile and stored as 16-byte
//                     vectors; BatchNorm partial sums accumulate per thread (one partial row per block).
//   k_wgrad_mid_mma : weight gradient.  A block stages a 64-pixel run of dY and the 3 x 66-pixel neighbourhood of X
//                     in shared memory; warp t of nine owns tap t: A = dY^T and B = X shifted by the tap, both
//                     through ldmatrix.trans, fp32 accumulators in registers for the whole kernel; one partial row
//                     per block in the layout k_wgrad_reduce expects.
#include <stdlib.h>

#include "common.cuh"

namespace ustrun {

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], const void* smem_row) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem_row);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x2_trans(uint32_t (&r)[2], const void* smem_row) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem_row);
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(a));
}

// ------------------------------------------------------------------------------------------
// forward / dgrad
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* smem_row) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem_row);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}

// A warp owns 16 consecutive pixels of an image row.  Per tile: the KS x (16 + 2R)-pixel input neighbourhood is staged
// in a warp-private shared slab with 16-byte loads (prefetched into registers one tile ahead, zero padded), every
// 16-channel K step is ONE ldmatrix.x4 (row-shifted by the tap) instead of four predicated global loads, and the
// weights are kept in shared memory in B-fragment order so that one LDS.128 per lane feeds two n-tiles.
template <int CIN, int NT, int KS>          // NT = Cout / 8, KS = 1 | 3
struct MidCfg {
  static constexpr int R = KS >> 1, TAPS = KS * KS, K = TAPS * CIN, KSTEPS = K / 16, COUT = NT * 8, CPT = CIN / 16;
  static constexpr int XW = 16 + 2 * R, XP = CIN + 8, SP = COUT + 8;
  static constexpr int slab_elems = KS * XW * XP, stg_elems = 16 * SP;
  static constexpr int warp_elems = ((slab_elems > stg_elems ? slab_elems : stg_elems) + 7) / 8 * 8;   // bf16 per warp, 16-byte multiple
  static constexpr int wfrag_u32 = KSTEPS * (NT / 2) * 32 * 4;                                          // [s][j2][lane][4]
  static constexpr int NCH = (KS * XW * (CIN / 8) + 31) / 32;                                           // 16-byte slab chunks per lane
  static constexpr size_t smem = (size_t)wfrag_u32 * 4 + (size_t)8 * warp_elems * 2 + (size_t)8 * 2 * COUT * 4;
};

// HEAD = true: the logits head of UNet-B (out1 / seg1: conv3x3 2n -> n_classes + bias, unet.py:182,312).  The output tile is
// padded to 16 columns (cout_valid <= 16 real classes, zero weights beyond) and leaves as fp32 NCHW class planes -- eight
// consecutive pixels per store instruction -- instead of a bf16 NHWC tile; no BatchNorm statistics.
template <int CIN, int NT, int KS, bool HEAD = false>
__global__ void __launch_bounds__(256)
k_conv_mid_mma(const __nv_bfloat16* __restrict__ x, int ldx, const __nv_bfloat16* __restrict__ wp, const float* __restrict__ bias,
               __nv_bfloat16* __restrict__ y, int ldy, int B, int H, int W, float* __restrict__ partials, float* __restrict__ y_nchw = nullptr,
               int cout_valid = NT * 8) {
  using Cfg = MidCfg<CIN, NT, KS>;
  constexpr int R = Cfg::R, TAPS = Cfg::TAPS, K = Cfg::K, KSTEPS = Cfg::KSTEPS, COUT = Cfg::COUT, CPT = Cfg::CPT;
  constexpr int XW = Cfg::XW, XP = Cfg::XP, SP = Cfg::SP, NCH = Cfg::NCH;
  extern __shared__ __align__(16) uint8_t mid_smem[];
  uint32_t* wfrag = reinterpret_cast<uint32_t*>(mid_smem);                                   // B fragments, [s][j2][lane][4]
  __nv_bfloat16* warp_all = reinterpret_cast<__nv_bfloat16*>(wfrag + Cfg::wfrag_u32);        // [8 warps][warp_elems]
  float* red = reinterpret_cast<float*>(warp_all + 8 * Cfg::warp_elems);                     // [8][2*COUT]
  // weights -> fragment order: entry (s, j, lane) = { w[co][16s + 2tig, +1], w[co][16s + 8 + 2tig, +1] }, co = 8j + g
  for (int i = threadIdx.x; i < KSTEPS * NT * 32; i += 256) {
    const int ln = i & 31, j = (i >> 5) % NT, s_ = i / (32 * NT);
    const uint32_t* wr = reinterpret_cast<const uint32_t*>(wp + (size_t)(8 * j + (ln >> 2)) * K + 16 * s_ + 2 * (ln & 3));
    uint32_t* d = wfrag + ((s_ * (NT / 2) + (j >> 1)) * 32 + ln) * 4 + 2 * (j & 1);
    const bool real = !HEAD || (8 * j + (ln >> 2)) < cout_valid;
    d[0] = real ? wr[0] : 0u;
    d[1] = real ? wr[4] : 0u;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int g = lane >> 2, tig = lane & 3;
  __nv_bfloat16* slab = warp_all + wrp * Cfg::warp_elems;       // [KS][XW][XP]; reused as the [16][SP] output staging tile
  float bs[NT][2], ssum[NT][2], ssq[NT][2];
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    bs[j][0] = (bias && 8 * j + 2 * tig < cout_valid) ? bias[8 * j + 2 * tig] : 0.f;
    bs[j][1] = (bias && 8 * j + 2 * tig + 1 < cout_valid) ? bias[8 * j + 2 * tig + 1] : 0.f;
    ssum[j][0] = ssum[j][1] = ssq[j][0] = ssq[j][1] = 0.f;
  }
  const int tiles_w = W >> 4;
  const long long ntiles = (long long)B * H * tiles_w, tstep = (long long)gridDim.x * 8;
  uint4 pre[NCH];
  auto load_slab = [&](long long tile) {       // global -> registers, zero padding; chunk c = (row r, pixel p, 8-channel group q)
    int tw, h_, b_;
    pix_decomp(tile, tiles_w, H, b_, h_, tw);
    const int w0 = tw << 4;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = lane + 32 * i;
      pre[i] = make_uint4(0, 0, 0, 0);
      if (c < KS * XW * (CIN / 8)) {
        const int q = c % (CIN / 8), p = (c / (CIN / 8)) % XW, r = c / ((CIN / 8) * XW);
        const int hh = h_ + r - R, ww = w0 + p - R;
        if (hh >= 0 && hh < H && ww >= 0 && ww < W) pre[i] = *reinterpret_cast<const uint4*>(x + ((long long)(b_ * H + hh) * W + ww) * ldx + q * 8);
      }
    }
  };
  long long tl = (long long)blockIdx.x * 8 + wrp;
  if (tl < ntiles) load_slab(tl);
  const int mat = lane >> 3, rr = lane & 7;
  for (; tl < ntiles; tl += tstep) {
    int tw, h_, b_;
    pix_decomp(tl, tiles_w, H, b_, h_, tw);
    const int w0 = tw << 4;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = lane + 32 * i;
      if (c < KS * XW * (CIN / 8)) {
        const int q = c % (CIN / 8), p = (c / (CIN / 8)) % XW, r = c / ((CIN / 8) * XW);
        *reinterpret_cast<uint4*>(slab + (r * XW + p) * XP + q * 8) = pre[i];
      }
    }
    __syncwarp();
    if (tl + tstep < ntiles) load_slab(tl + tstep);
    float acc[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
    for (int t = 0; t < TAPS; ++t) {
      const int dyi = KS == 3 ? t / 3 : 0, dxi = KS == 3 ? t % 3 : 0;
      // A fragment rows = pixels (slab column p + dxi), matrices: (pixels 0-7 | 8-15) x (channels +0 | +8)
      const __nv_bfloat16* arow = slab + (dyi * XW + dxi + rr + 8 * (mat & 1)) * XP + 8 * (mat >> 1);
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        uint32_t af[4];
        ldsm_x4(af, arow + 16 * c);
        const int s_ = t * CPT + c;
#pragma unroll
        for (int j2 = 0; j2 < NT / 2; ++j2) {
          const uint4 bq = *reinterpret_cast<const uint4*>(wfrag + ((s_ * (NT / 2) + j2) * 32 + lane) * 4);
          mma16816(acc[2 * j2], af, bq.x, bq.y);
          mma16816(acc[2 * j2 + 1], af, bq.z, bq.w);
        }
      }
    }
    __syncwarp();                                // every lane is done with the slab: reuse it as the output staging tile
    if constexpr (HEAD) {
      const long long plane = (long long)H * W;
      float* o = y_nchw + (long long)b_ * cout_valid * plane + (long long)h_ * W + w0 + g;
#pragma unroll
      for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int co = 8 * j + 2 * tig + e;
          if (co < cout_valid) {
            o[co * plane] = acc[j][e] + bs[j][e];
            o[co * plane + 8] = acc[j][2 + e] + bs[j][e];
          }
        }
      continue;
    }
    __nv_bfloat16* st = slab;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      ssum[j][0] += acc[j][0] + acc[j][2]; ssum[j][1] += acc[j][1] + acc[j][3];
      ssq[j][0] += acc[j][0] * acc[j][0] + acc[j][2] * acc[j][2]; ssq[j][1] += acc[j][1] * acc[j][1] + acc[j][3] * acc[j][3];
      *reinterpret_cast<__nv_bfloat162*>(st + g * SP + 8 * j + 2 * tig) = __floats2bfloat162_rn(acc[j][0] + bs[j][0], acc[j][1] + bs[j][1]);
      *reinterpret_cast<__nv_bfloat162*>(st + (g + 8) * SP + 8 * j + 2 * tig) = __floats2bfloat162_rn(acc[j][2] + bs[j][0], acc[j][3] + bs[j][1]);
    }
    __syncwarp();
    __nv_bfloat16* yrow = y + ((long long)(b_ * H + h_) * W + w0) * ldy;
#pragma unroll
    for (int c = lane; c < 16 * NT; c += 32) {
      const int r = c / NT, cc = c - r * NT;
      *reinterpret_cast<uint4*>(yrow + (long long)r * ldy + cc * 8) = *reinterpret_cast<const uint4*>(st + r * SP + cc * 8);
    }
    __syncwarp();
  }
  if (partials) {
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float a = ssum[j][e], q = ssq[j][e];
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
        if (g == 0) { red[wrp * 2 * COUT + 8 * j + 2 * tig + e] = a; red[wrp * 2 * COUT + COUT + 8 * j + 2 * tig + e] = q; }
      }
    __syncthreads();
    for (int o = threadIdx.x; o < 2 * COUT; o += 256) {
      float s_ = 0.f;
#pragma unroll
      for (int w_ = 0; w_ < 8; ++w_) s_ += red[w_ * 2 * COUT + o];
      partials[(size_t)blockIdx.x * 2 * COUT + o] = s_;
    }
  }
}

// ------------------------------------------------------------------------------------------
// weight gradient: D_t[co][ci] = sum_p dY[p][co] * X[p + tap t][ci]
// ------------------------------------------------------------------------------------------
template <int CIN, int COUT, int KS>
__global__ void __launch_bounds__(KS == 3 ? 288 : 128)
k_wgrad_mid_mma(const __nv_bfloat16* __restrict__ dy, int lddy, const __nv_bfloat16* __restrict__ x, int ldx, int B, int H, int W, float* __restrict__ ws) {
  constexpr int TAPS = KS * KS, NWARP = KS == 3 ? 9 : 4, NTHR = NWARP * 32, R = KS >> 1;
  constexpr int MT = COUT / 16, NT = CIN / 8, YP = COUT + 8, XP = CIN + 8, XW = 64 + 2 * R, XR = KS;
  extern __shared__ __align__(16) uint8_t mid_smem[];
  __nv_bfloat16* ys = reinterpret_cast<__nv_bfloat16*>(mid_smem);              // [64 pixels][YP]
  __nv_bfloat16* xs = ys + 64 * YP;                                             // [XR rows][XW pixels][XP]
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  // KS == 3: warp = tap.  KS == 1: the four warps split the 64-pixel run (one 16-pixel K step each).
  const int tap = KS == 3 ? wrp : 0;
  const int tdy = KS == 3 ? tap / 3 : 0, tdx = KS == 3 ? tap % 3 : 0;
  float acc[MT][NT][4];
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int n = 0; n < NT; ++n) acc[m][n][0] = acc[m][n][1] = acc[m][n][2] = acc[m][n][3] = 0.f;
  const int segs_w = (W + 63) >> 6;
  const long long nsegs = (long long)B * H * segs_w;
  // 16-byte chunks of a run: first the dY run [64][COUT/8], then the X neighbourhood [XR][XW][CIN/8]; every thread
  // prefetches its chunks of the NEXT run into registers while the current one is consumed from shared memory
  constexpr int NY = 64 * (COUT / 8), NX = XR * XW * (CIN / 8), NCH = (NY + NX + NTHR - 1) / NTHR;
  uint4 pre[NCH];
  auto load_run = [&](long long seg) {
    int sw, h_, b_;
    pix_decomp(seg, segs_w, H, b_, h_, sw);
    const int w0 = sw << 6;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      const int i = threadIdx.x + k * NTHR;
      pre[k] = make_uint4(0, 0, 0, 0);
      if (i < NY) {
        const int p = i / (COUT / 8), c = i - p * (COUT / 8);
        if (w0 + p < W) pre[k] = *reinterpret_cast<const uint4*>(dy + ((long long)(b_ * H + h_) * W + w0 + p) * lddy + c * 8);
      } else if (i < NY + NX) {
        const int j = i - NY;
        const int c = j % (CIN / 8), p = (j / (CIN / 8)) % XW, r = j / ((CIN / 8) * XW);
        const int hh = h_ + r - R, ww = w0 + p - R;
        if (hh >= 0 && hh < H && ww >= 0 && ww < W) pre[k] = *reinterpret_cast<const uint4*>(x + ((long long)(b_ * H + hh) * W + ww) * ldx + c * 8);
      }
    }
  };
  if ((long long)blockIdx.x < nsegs) load_run(blockIdx.x);
  for (long long sg = blockIdx.x; sg < nsegs; sg += gridDim.x) {
    __syncthreads();                                            // the previous run has been consumed
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      const int i = threadIdx.x + k * NTHR;
      if (i < NY) {
        const int p = i / (COUT / 8), c = i - p * (COUT / 8);
        *reinterpret_cast<uint4*>(ys + p * YP + c * 8) = pre[k];
      } else if (i < NY + NX) {
        const int j = i - NY;
        const int c = j % (CIN / 8), p = (j / (CIN / 8)) % XW, r = j / ((CIN / 8) * XW);
        *reinterpret_cast<uint4*>(xs + (r * XW + p) * XP + c * 8) = pre[k];
      }
    }
    __syncthreads();
    if (sg + gridDim.x < nsegs) load_run(sg + gridDim.x);
    const int mat = lane >> 3, rr = lane & 7;
#pragma unroll
    for (int ks_ = 0; ks_ < 4; ++ks_) {                         // 16-pixel K steps of the run
      if (KS == 1 && ks_ != wrp) continue;
      const int p0 = 16 * ks_;
      uint32_t af[MT][4];
#pragma unroll
      for (int m = 0; m < MT; ++m)       // A = dY^T: matrices (pixels 0-7 | 8-15) x (channels 16m.. | 16m+8..)
        ldsm_x4_trans(af[m], ys + (p0 + rr + 8 * (mat >> 1)) * YP + 16 * m + 8 * (mat & 1));
      const __nv_bfloat16* xb = xs + (tdy * XW + p0 + tdx) * XP;   // X run shifted by the tap
#pragma unroll
      for (int n2 = 0; n2 < NT / 2; ++n2) {   // B fragments of two n tiles per ldmatrix.x4: (pixels 0-7 | 8-15) x (ci 16n2.. | 16n2+8..)
        uint32_t bf[4];
        ldsm_x4_trans(bf, xb + (rr + 8 * (mat & 1)) * XP + 16 * n2 + 8 * (mat >> 1));
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          mma16816(acc[m][2 * n2], af[m], bf[0], bf[1]);
          mma16816(acc[m][2 * n2 + 1], af[m], bf[2], bf[3]);
        }
      }
    }
  }
  // one partial row per block: ws[block][co][tap*CIN + ci]; the KS == 1 warps add up through shared memory
  const int g = lane >> 2, tig = lane & 3;
  float* row = ws + (size_t)blockIdx.x * ((size_t)COUT * TAPS * CIN);
  if (KS == 3) {
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        float* d0 = row + (size_t)(16 * m + g) * (TAPS * CIN) + tap * CIN + 8 * n + 2 * tig;
        float* d1 = row + (size_t)(16 * m + g + 8) * (TAPS * CIN) + tap * CIN + 8 * n + 2 * tig;
        *reinterpret_cast<float2*>(d0) = make_float2(acc[m][n][0], acc[m][n][1]);
        *reinterpret_cast<float2*>(d1) = make_float2(acc[m][n][2], acc[m][n][3]);
      }
  } else {
    __syncthreads();
    float* redw = reinterpret_cast<float*>(mid_smem);           // [COUT][CIN], reused
    for (int i = threadIdx.x; i < COUT * CIN; i += NTHR) redw[i] = 0.f;
    __syncthreads();
    for (int w_ = 0; w_ < NWARP; ++w_) {
      if (wrp == w_) {
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
          for (int n = 0; n < NT; ++n) {
            redw[(16 * m + g) * CIN + 8 * n + 2 * tig] += acc[m][n][0];
            redw[(16 * m + g) * CIN + 8 * n + 2 * tig + 1] += acc[m][n][1];
            redw[(16 * m + g + 8) * CIN + 8 * n + 2 * tig] += acc[m][n][2];
            redw[(16 * m + g + 8) * CIN + 8 * n + 2 * tig + 1] += acc[m][n][3];
          }
      }
      __syncthreads();
    }
    for (int i = threadIdx.x; i < COUT * CIN; i += NTHR) row[i] = redw[i];
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static int mid_mode() {      // USTRUN_MID_MMA=0 keeps the CUDA-core kernels (A/B comparisons)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("USTRUN_MID_MMA");
    v = e ? atoi(e) : 1;
  }
  return v;
}
static inline bool mid_c(int c) { return c == 16 || c == 32 || c == 64; }
bool mid_conv_ok(int Cin, int Cout, int ks, int W, int ldx, int ldy) {
  return mid_mode() > 0 && mid_c(Cin) && mid_c(Cout) && !(Cin == 64 && Cout == 64) && (ks == 1 || ks == 3) && W % 16 == 0 && ldx % 8 == 0 && ldy % 8 == 0;
}
constexpr int kMidWgradBlocks = 148 * 2;
bool mid_wgrad_ok(int Cin, int Cout, int ks, int ldx, int lddy) {
  return mid_mode() > 0 && mid_c(Cin) && mid_c(Cout) && !(Cin == 64 && Cout == 64) && (ks == 1 || ks == 3) && ldx % 8 == 0 && lddy % 8 == 0;
}
long long mid_wgrad_ws_bytes(int Cin, int Cout, int ks) { return (long long)kMidWgradBlocks * Cout * ks * ks * Cin * (long long)sizeof(float); }

template <int CIN, int NT, int KS>
static int mid_conv_launch_t(const void* x, int ldx, const void* w, const float* bias, void* y, int ldy, int B, int H, int W, float* partials,
                             int* nparts_host, cudaStream_t st) {
  const size_t smem = MidCfg<CIN, NT, KS>::smem;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(k_conv_mid_mma<CIN, NT, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(k_conv_mid_mma): %s", cudaGetErrorString(e)); return (int)e; }
    attr = true;
  }
  const long long ntiles = (long long)B * H * (W / 16);
  long long g = (ntiles + 7) / 8;
  if (g > 148 * 4) g = 148 * 4;
  if (nparts_host) *nparts_host = (int)g;
  k_conv_mid_mma<CIN, NT, KS><<<(int)g, 256, smem, st>>>((const __nv_bfloat16*)x, ldx, (const __nv_bfloat16*)w, bias, (__nv_bfloat16*)y, ldy, B, H, W, partials);
  return check_launch("conv_mid_mma");
}
template <int CIN, int KS>
static int mid_conv_launch_c(int Cout, const void* x, int ldx, const void* w, const float* bias, void* y, int ldy, int B, int H, int W, float* partials,
                             int* nparts_host, cudaStream_t st) {
  if (Cout == 16) return mid_conv_launch_t<CIN, 2, KS>(x, ldx, w, bias, y, ldy, B, H, W, partials, nparts_host, st);
  if (Cout == 32) return mid_conv_launch_t<CIN, 4, KS>(x, ldx, w, bias, y, ldy, B, H, W, partials, nparts_host, st);
  return mid_conv_launch_t<CIN, 8, KS>(x, ldx, w, bias, y, ldy, B, H, W, partials, nparts_host, st);
}
int mid_conv_launch(const void* x, int ldx, const void* w, const float* bias, void* y, int ldy, int B, int H, int W, int Cin, int Cout, int ks,
                    float* partials, int* nparts_host, cudaStream_t st) {
#define MC(CI) (ks == 3 ? mid_conv_launch_c<CI, 3>(Cout, x, ldx, w, bias, y, ldy, B, H, W, partials, nparts_host, st) \
                        : mid_conv_launch_c<CI, 1>(Cout, x, ldx, w, bias, y, ldy, B, H, W, partials, nparts_host, st))
  if (Cin == 16) return MC(16);
  if (Cin == 32) return MC(32);
  return MC(64);
#undef MC
}

// logits head of UNet-B: conv3x3 (16 | 32 channels -> <= 16 classes) + bias -> fp32 NCHW
bool mid_head_ok(int Cin, int Cout, int ks, int W, int ldx) {
  return mid_mode() > 0 && (Cin == 16 || Cin == 32) && Cout >= 1 && Cout <= 16 && ks == 3 && W % 16 == 0 && ldx % 8 == 0;
}
template <int CIN>
static int mid_head_launch_t(const void* x, int ldx, const void* w, const float* bias, float* y_nchw, int B, int H, int W, int Cout, cudaStream_t st) {
  const size_t smem = MidCfg<CIN, 2, 3>::smem;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(k_conv_mid_mma<CIN, 2, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(k_conv_mid_mma head): %s", cudaGetErrorString(e)); return (int)e; }
    attr = true;
  }
  const long long ntiles = (long long)B * H * (W / 16);
  long long g = (ntiles + 7) / 8;
  if (g > 148 * 4) g = 148 * 4;
  k_conv_mid_mma<CIN, 2, 3, true><<<(int)g, 256, smem, st>>>((const __nv_bfloat16*)x, ldx, (const __nv_bfloat16*)w, bias, nullptr, 0, B, H, W, nullptr, y_nchw, Cout);
  return check_launch("conv_mid_mma(head)");
}
int mid_head_launch(const void* x, int ldx, const void* w, const float* bias, float* y_nchw, int B, int H, int W, int Cin, int Cout, cudaStream_t st) {
  return Cin == 16 ? mid_head_launch_t<16>(x, ldx, w, bias, y_nchw, B, H, W, Cout, st) : mid_head_launch_t<32>(x, ldx, w, bias, y_nchw, B, H, W, Cout, st);
}

int launch_wgrad_reduce(const float* ws, int splits, int Mo, int Nin, int taps, float* dw, int accumulate, cudaStream_t st, int swapped);

template <int CIN, int COUT, int KS>
static int mid_wgrad_launch_t(const void* dy, int lddy, const void* x, int ldx, int B, int H, int W, float* dw, int accumulate, void* ws, cudaStream_t st) {
  constexpr int R = KS >> 1;
  const size_t smem_run = (size_t)64 * (COUT + 8) * 2 + (size_t)KS * (64 + 2 * R) * (CIN + 8) * 2;
  const size_t smem = smem_run > (size_t)COUT * CIN * 4 ? smem_run : (size_t)COUT * CIN * 4;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(k_wgrad_mid_mma<CIN, COUT, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(k_wgrad_mid_mma): %s", cudaGetErrorString(e)); return (int)e; }
    attr = true;
  }
  const long long nsegs = (long long)B * H * ((W + 63) / 64);
  int grid = (int)(nsegs < kMidWgradBlocks ? nsegs : kMidWgradBlocks);
  k_wgrad_mid_mma<CIN, COUT, KS><<<grid, KS == 3 ? 288 : 128, smem, st>>>((const __nv_bfloat16*)dy, lddy, (const __nv_bfloat16*)x, ldx, B, H, W, (float*)ws);
  int rc = check_launch("wgrad_mid_mma");
  if (rc) return rc;
  return launch_wgrad_reduce((const float*)ws, grid, COUT, CIN, KS * KS, dw, accumulate, st, 0);
}
int mid_wgrad_launch(const void* dy, int lddy, const void* x, int ldx, float* dw, int accumulate, int B, int H, int W, int Cin, int Cout, int ks,
                     void* workspace, long long ws_bytes, cudaStream_t st) {
  if (!workspace || ws_bytes < mid_wgrad_ws_bytes(Cin, Cout, ks)) { set_error("mid wgrad: workspace too small"); return USTRUN_ERR_ARG; }
#define MW(CI, CO) (ks == 3 ? mid_wgrad_launch_t<CI, CO, 3>(dy, lddy, x, ldx, B, H, W, dw, accumulate, workspace, st) \
                            : mid_wgrad_launch_t<CI, CO, 1>(dy, lddy, x, ldx, B, H, W, dw, accumulate, workspace, st))
  if (Cin == 16 && Cout == 16) return MW(16, 16);
  if (Cin == 16 && Cout == 32) return MW(16, 32);
  if (Cin == 16 && Cout == 64) return MW(16, 64);
  if (Cin == 32 && Cout == 16) return MW(32, 16);
  if (Cin == 32 && Cout == 32) return MW(32, 32);
  if (Cin == 32 && Cout == 64) return MW(32, 64);
  if (Cin == 64 && Cout == 16) return MW(64, 16);
  return MW(64, 32);
#undef MW
}

}  // namespace ustrun
