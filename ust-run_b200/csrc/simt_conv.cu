// CUDA-core implicit-GEMM convolutions (fp32 accumulate) for
//   (a) the fp32 validation mode (every layer), and
//   (b) the layers a tensor-core tile cannot serve in bf16 mode (first conv with Cin=1..3, the
//       n_classes-wide logits head, UNet-B's 16/32-channel levels).
// One generic forward-type kernel (conv3x3 / conv1x1 / dgrad via flipped packing / the four 1x1
// GEMMs of ConvTranspose2d(k2,s2)) and one generic weight-gradient kernel with split-K.
#include "common.cuh"

namespace ustrun {

constexpr int TM = 64, TN = 64, TK = 16;

enum { GATHER_CONV = 0, GATHER_UP = 1 };

// address of input element for GEMM row pixel (b,h,w), tap t, channel c; returns false if padded
__device__ __forceinline__ bool in_offset(const ConvGeom& g, int b, int h, int w, int t, long long& pix) {
  if (g.gather == GATHER_CONV) {
    int r = g.ks >> 1;
    int hh = h + (g.ks == 3 ? t / 3 : 0) - r, ww = w + (g.ks == 3 ? t % 3 : 0) - r;
    if (hh < 0 || hh >= g.H || ww < 0 || ww >= g.W) return false;
    pix = ((long long)b * g.H + hh) * g.W + ww;
    return true;
  }
  pix = ((long long)b * 2 * g.H + 2 * h + (t >> 1)) * 2 * g.W + 2 * w + (t & 1);
  return true;
}

template <typename T>
__global__ void __launch_bounds__(256)
k_conv_simt(const T* __restrict__ x, int ldx, const T* __restrict__ wp, const float* __restrict__ bias, T* __restrict__ y, int ldy,
            float* __restrict__ y_nchw, ConvGeom g, float* __restrict__ partials) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  __shared__ float red[16][TN];
  const int taps = g.gather == GATHER_UP ? 4 : g.ks * g.ks;
  const int K = taps * g.Cin;
  const long long M = (long long)g.B * g.H * g.W;
  const int mtiles = (int)((M + TM - 1) / TM);
  const int n0 = blockIdx.y * TN;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float csum[4] = {0, 0, 0, 0}, csq[4] = {0, 0, 0, 0};

  for (int mt = blockIdx.x; mt < mtiles; mt += gridDim.x) {
    const long long m0 = (long long)mt * TM;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < K; k0 += TK) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int e = tid + i * 256;
        int kk = e & 15, mm = e >> 4;
        int k = k0 + kk;
        long long m = m0 + mm;
        float v = 0.f;
        if (k < K && m < M) {
          int t = (int)((unsigned)k / (unsigned)g.Cin), c = k - t * g.Cin;
          int w_, h_, b_;
          pix_decomp(m, g.W, g.H, b_, h_, w_);
          long long pix;
          if (in_offset(g, b_, h_, w_, t, pix)) v = to_f(x[pix * ldx + c]);
        }
        As[kk][mm] = v;
        int nn = e >> 4;                       // reuse mapping: 64 n x 16 k
        int n = n0 + nn;
        Bs[kk][nn] = (k < K && n < g.Cout) ? to_f(wp[(size_t)n * K + k]) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < TK; ++kk) {
        float a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
    // epilogue
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      long long m = m0 + ty * 4 + i;
      if (m >= M) continue;
      int w_, h_, b_;
      pix_decomp(m, g.W, g.H, b_, h_, w_);
      long long opix = m;
      int OH = g.H, OW = g.W;
      if (g.scatter_ij >= 0) {
        OH = 2 * g.H; OW = 2 * g.W;
        opix = ((long long)b_ * OH + 2 * h_ + (g.scatter_ij >> 1)) * OW + 2 * w_ + (g.scatter_ij & 1);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int n = n0 + tx * 4 + j;
        if (n >= g.Cout) continue;
        float v = acc[i][j];
        csum[j] += v;
        csq[j] += v * v;
        if (bias) v += bias[n];
        if (y_nchw) y_nchw[(((long long)b_ * g.Cout + n) * OH + (opix / OW) % OH) * OW + opix % OW] = v;
        else y[opix * ldy + n] = from_f<T>(v);
      }
    }
  }
  if (partials) {
    for (int which = 0; which < 2; ++which) {
#pragma unroll
      for (int j = 0; j < 4; ++j) red[ty][tx * 4 + j] = which ? csq[j] : csum[j];
      __syncthreads();
      if (tid < TN) {
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < 16; ++r) s += red[r][tid];
        int n = n0 + tid;
        if (n < g.Cout) partials[(size_t)blockIdx.x * 2 * g.Cout + which * g.Cout + n] = s;
      }
      __syncthreads();
    }
  }
}

// ws[split][Mo][taps*Nin] = sum_{p in split} A[p][m] * Bg[p (+tap)][c];  n = t*Nin + c
//   conv  : A = dy (Mo=Cout), Bg = x gathered with conv taps (Nin=Cin)
//   convT : A = x  (Mo=Cin),  Bg = dy gathered on the up-sampled grid (Nin=Cout)
template <typename T>
__global__ void __launch_bounds__(256)
k_wgrad_simt(const T* __restrict__ a, int lda, const T* __restrict__ bsrc, int ldb, float* __restrict__ ws, ConvGeom g, int Mo,
             int Nin, long long pix_per_split) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const int taps = g.gather == GATHER_UP ? 4 : g.ks * g.ks;
  const int N = taps * Nin;
  const long long P = (long long)g.B * g.H * g.W;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const long long p_begin = (long long)blockIdx.z * pix_per_split;
  long long p_end = p_begin + pix_per_split;
  if (p_end > P) p_end = P;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (long long p0 = p_begin; p0 < p_end; p0 += TK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int e = tid + i * 256;
      int mm = e & 63, kk = e >> 6;             // 64 channels contiguous per pixel
      long long p = p0 + kk;
      int m = m0 + mm;
      As[kk][mm] = (p < p_end && m < Mo) ? to_f(a[p * lda + m]) : 0.f;
      int n = n0 + mm;
      float v = 0.f;
      if (p < p_end && n < N) {
        int t = n / Nin, c = n - t * Nin;
        int w_, h_, b_;
        pix_decomp(p, g.W, g.H, b_, h_, w_);
        long long pix;
        if (in_offset(g, b_, h_, w_, t, pix)) v = to_f(bsrc[pix * ldb + c]);
      }
      Bs[kk][mm] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= Mo) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n < N) ws[((size_t)blockIdx.z * Mo + m) * N + n] = acc[i][j];
    }
  }
}

// dw[(m*Nin + c)*taps + t] (+)= sum_s ws[s][m][t*Nin + c].  One WARP per (m, 32 consecutive c): lane = c, so the
// taps*splits loads of a lane are independent and coalesced along c; the 32 x taps results are transposed through a
// per-warp shared-memory tile (no block barrier) and leave as ONE contiguous run of 32*taps floats of the OIHW
// tensor (read-modify-write when accumulating), 128 bytes per store instruction.
template <int TAPS>
__global__ void __launch_bounds__(256)
k_wgrad_reduce(const float* __restrict__ ws, int splits, int Mo, int Nin, float* __restrict__ dw, int accumulate, int swapped) {
  __shared__ float tile[8][32 * TAPS + 1];
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int cchunks = (Nin + 31) >> 5;
  const long long n = (long long)Mo * Nin * TAPS, groups = (long long)Mo * cchunks;
  for (long long g = (long long)blockIdx.x * 8 + wrp; g < groups; g += (long long)gridDim.x * 8) {
    const int m = (int)(g / cchunks), c0 = (int)(g - (long long)m * cchunks) * 32;
    const int nc = min(32, Nin - c0);
    float acc[TAPS];
#pragma unroll
    for (int t = 0; t < TAPS; ++t) acc[t] = 0.f;
    if (lane < nc) {
      const float* src = ws + ((size_t)m * TAPS) * Nin + c0 + lane;
#pragma unroll 4
      for (int k = 0; k < splits; ++k) {               // up to ~49 splits for the 64-channel layers: keep 4 x TAPS loads in flight
#pragma unroll
        for (int t = 0; t < TAPS; ++t) acc[t] += src[(size_t)k * n + (size_t)t * Nin];
      }
    }
    if (swapped) {                                    // ws rows are input channels: dw[(c*Mo + m)*taps + t] (small layers only)
      if (lane < nc) {
        float* dst = dw + ((size_t)(c0 + lane) * Mo + m) * TAPS;
#pragma unroll
        for (int t = 0; t < TAPS; ++t) dst[t] = (accumulate ? dst[t] : 0.f) + acc[t];
      }
      continue;
    }
#pragma unroll
    for (int t = 0; t < TAPS; ++t) tile[wrp][lane * TAPS + t] = acc[t];
    __syncwarp();
    float* dst = dw + ((size_t)m * Nin + c0) * TAPS;
    for (int i = lane; i < nc * TAPS; i += 32) dst[i] = (accumulate ? dst[i] : 0.f) + tile[wrp][i];
    __syncwarp();
  }
}

// Small layers, many splits (64x64 weights, 49 K splits): the kernel above has only Mo*Nin/32 warps -- 16 CTAs, each
// lane walking 49 x 9 dependent-latency loads (measured 37 us for 7 MB).  Here one CTA owns a (m, 32 c) group and its
// 8 warps share the splits; partial sums meet in shared memory.
template <int TAPS>
__global__ void __launch_bounds__(256)
k_wgrad_reduce_wide(const float* __restrict__ ws, int splits, int Mo, int Nin, float* __restrict__ dw, int accumulate, int swapped) {
  __shared__ float part[8][32 * TAPS + 1];
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int cchunks = (Nin + 31) >> 5;
  const long long n = (long long)Mo * Nin * TAPS;
  const int m = blockIdx.x / cchunks, c0 = (blockIdx.x - m * cchunks) * 32;
  const int nc = min(32, Nin - c0);
  float acc[TAPS];
#pragma unroll
  for (int t = 0; t < TAPS; ++t) acc[t] = 0.f;
  if (lane < nc) {
    const float* src = ws + ((size_t)m * TAPS) * Nin + c0 + lane;
#pragma unroll 2
    for (int k = wrp; k < splits; k += 8) {
#pragma unroll
      for (int t = 0; t < TAPS; ++t) acc[t] += src[(size_t)k * n + (size_t)t * Nin];
    }
  }
#pragma unroll
  for (int t = 0; t < TAPS; ++t) part[wrp][lane * TAPS + t] = acc[t];
  __syncthreads();
  for (int i = threadIdx.x; i < nc * TAPS; i += 256) {
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += part[w][i];
    float* dst;
    if (swapped) {                                    // ws rows are input channels: dw[(c*Mo + m)*taps + t]
      const int cl = i / TAPS, t = i - cl * TAPS;
      dst = dw + ((size_t)(c0 + cl) * Mo + m) * TAPS + t;
    } else {
      dst = dw + ((size_t)m * Nin + c0) * TAPS + i;
    }
    *dst = (accumulate ? *dst : 0.f) + sum;
  }
}

int launch_wgrad_reduce(const float* ws, int splits, int Mo, int Nin, int taps, float* dw, int accumulate, cudaStream_t st, int swapped) {
  const long long groups = (long long)Mo * ((Nin + 31) / 32);
  int rb = (int)((groups + 7) / 8);
  if (rb < 148 && splits >= 8) {                     // fewer CTAs than SMs: spread the splits over the warps of a CTA instead
    switch (taps) {
      case 9: k_wgrad_reduce_wide<9><<<(int)groups, 256, 0, st>>>(ws, splits, Mo, Nin, dw, accumulate, swapped); break;
      case 4: k_wgrad_reduce_wide<4><<<(int)groups, 256, 0, st>>>(ws, splits, Mo, Nin, dw, accumulate, swapped); break;
      case 1: k_wgrad_reduce_wide<1><<<(int)groups, 256, 0, st>>>(ws, splits, Mo, Nin, dw, accumulate, swapped); break;
      default: set_error("wgrad_reduce: taps must be 1, 4 or 9 (got %d)", taps); return USTRUN_ERR_ARG;
    }
    return check_launch("wgrad_reduce_wide");
  }
  if (rb > 148 * 8) rb = 148 * 8;
  switch (taps) {
    case 9: k_wgrad_reduce<9><<<rb, 256, 0, st>>>(ws, splits, Mo, Nin, dw, accumulate, swapped); break;
    case 4: k_wgrad_reduce<4><<<rb, 256, 0, st>>>(ws, splits, Mo, Nin, dw, accumulate, swapped); break;
    case 1: k_wgrad_reduce<1><<<rb, 256, 0, st>>>(ws, splits, Mo, Nin, dw, accumulate, swapped); break;
    default: set_error("wgrad_reduce: taps must be 1, 4 or 9 (got %d)", taps); return USTRUN_ERR_ARG;
  }
  return check_launch("wgrad_reduce");
}

static int simt_splits(long long P, int Mo, int N) {
  long long tiles = (long long)ceil_div(Mo, TM) * ceil_div(N, TN);
  long long s = (592 + tiles - 1) / tiles;
  long long maxs = (P + 255) / 256;
  if (s > maxs) s = maxs;
  if (s < 1) s = 1;
  if (s > 256) s = 256;
  return (int)s;
}

long long simt_wgrad_ws_bytes(long long P, int Mo, int Nin, int taps) {
  return (long long)simt_splits(P, Mo, taps * Nin) * Mo * taps * Nin * (long long)sizeof(float);
}

template <typename T>
int simt_conv_launch(const void* x, int ldx, const void* w, const float* bias, void* y, int ldy, float* y_nchw, ConvGeom g,
                     float* partials, int* nparts_host, cudaStream_t st) {
  long long M = (long long)g.B * g.H * g.W;
  int mtiles = (int)((M + TM - 1) / TM);
  int gx = mtiles < 592 ? mtiles : 592;
  if (nparts_host) *nparts_host = gx;
  dim3 grid(gx, ceil_div(g.Cout, TN));
  k_conv_simt<T><<<grid, 256, 0, st>>>((const T*)x, ldx, (const T*)w, bias, (T*)y, ldy, y_nchw, g, partials);
  return check_launch("conv_simt");
}

template <typename T>
int simt_wgrad_launch(const void* a, int lda, const void* b, int ldb, float* dw, int accumulate, ConvGeom g, int Mo, int Nin,
                      void* workspace, long long ws_bytes, cudaStream_t st) {
  int taps = g.gather == GATHER_UP ? 4 : g.ks * g.ks;
  long long P = (long long)g.B * g.H * g.W;
  int N = taps * Nin;
  int splits = simt_splits(P, Mo, N);
  long long need = (long long)splits * Mo * N * (long long)sizeof(float);
  if (!workspace || ws_bytes < need) {
    set_error("wgrad: workspace too small (%lld < %lld)", ws_bytes, need);
    return USTRUN_ERR_ARG;
  }
  long long pps = (P + splits - 1) / splits;
  pps = (pps + TK - 1) / TK * TK;
  dim3 grid(ceil_div(N, TN), ceil_div(Mo, TM), splits);
  k_wgrad_simt<T><<<grid, 256, 0, st>>>((const T*)a, lda, (const T*)b, ldb, (float*)workspace, g, Mo, Nin, pps);
  int rc = check_launch("wgrad_simt");
  if (rc) return rc;
  return launch_wgrad_reduce((const float*)workspace, splits, Mo, Nin, taps, dw, accumulate, st, 0);
}

template int simt_conv_launch<float>(const void*, int, const void*, const float*, void*, int, float*, ConvGeom, float*, int*, cudaStream_t);
template int simt_conv_launch<__nv_bfloat16>(const void*, int, const void*, const float*, void*, int, float*, ConvGeom, float*, int*, cudaStream_t);
template int simt_wgrad_launch<float>(const void*, int, const void*, int, float*, int, ConvGeom, int, int, void*, long long, cudaStream_t);
template int simt_wgrad_launch<__nv_bfloat16>(const void*, int, const void*, int, float*, int, ConvGeom, int, int, void*, long long, cudaStream_t);

}  // namespace ustrun
