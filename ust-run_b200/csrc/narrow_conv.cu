// HBM-bound convolutions with one very narrow side (SURVEY App. B marks them "HBM"):
//   * narrow-K  : first conv of the network (Cin = 1..4 image channels, K = 9*Cin) and the dgrad
//                 of the logits head (K = taps * n_classes)           -> k_conv_narrow_in
//   * narrow-N  : the logits head itself (Cout = n_classes <= 8)       -> k_conv_narrow_out
//   * their weight gradients (one operand 8..256 channels wide, the other <= 8) -> k_wgrad_narrow
// A tensor-core tile would be >90 % padding here; these kernels stream the wide tensor once with
// 128-bit accesses and keep the small operand / weights in shared memory or registers.
#include <stdlib.h>

#include "common.cuh"

namespace ustrun {

constexpr int NARROW_MAX_W = 6144;   // floats of weights kept in shared memory

// y[p][co] = sum_{tap,ci} x[p+tap][ci] * w[co][tap][ci];  K = taps*Cin small, Cout % 8 == 0
template <typename T, int CIN>     // CIN = compile-time input channels for the 3x3 fast path (0: runtime Cin)
__global__ void __launch_bounds__(256)
k_conv_narrow_in(const T* __restrict__ x, int ldx, const T* __restrict__ wp, const float* __restrict__ bias, T* __restrict__ y, int ldy,
                 int B, int H, int W, int Cin_rt, int Cout, int ks, float* __restrict__ partials) {
  const int Cin = CIN > 0 ? CIN : Cin_rt;
  __shared__ __align__(16) float w_s[NARROW_MAX_W];     // [k][co]
  __shared__ float red[256][16];
  const int taps = ks * ks, K = taps * Cin;
  for (int i = threadIdx.x; i < K * Cout; i += 256) {
    int k = i / Cout, co = i - k * Cout;
    w_s[i] = to_f(wp[(size_t)co * K + k]);
  }
  __syncthreads();
  const int CG = Cout >> 3, lanes = 256 / CG;
  const int cg = threadIdx.x % CG, lane = threadIdx.x / CG;
  const long long M = (long long)B * H * W;
  const int r = ks >> 1;
  float bs[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) bs[j] = bias ? bias[cg * 8 + j] : 0.f;
  float csum[8], csq[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { csum[j] = 0.f; csq[j] = 0.f; }
  for (long long p = (long long)blockIdx.x * lanes + lane; p < M; p += (long long)gridDim.x * lanes) {
    int w_, h_, b_;
    pix_decomp(p, W, H, b_, h_, w_);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    if (CIN > 0) {
      // compile-time Cin, 3x3: issue all 9*Cin neighbour loads before any FMA (independent loads)
      constexpr int CI = CIN > 0 ? CIN : 1;
      float xv[9][CI];
      const bool interior = h_ > 0 && h_ < H - 1 && w_ > 0 && w_ < W - 1;
      const T* xc = x + p * ldx;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int dh = t / 3 - 1, dw = t % 3 - 1;
        const bool ok = interior || (h_ + dh >= 0 && h_ + dh < H && w_ + dw >= 0 && w_ + dw < W);
        const T* xp = xc + (ok ? ((long long)dh * W + dw) * ldx : 0);
#pragma unroll
        for (int ci = 0; ci < CI; ++ci) xv[t][ci] = ok ? to_f(xp[ci]) : 0.f;
      }
#pragma unroll
      for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int ci = 0; ci < CI; ++ci)
          {
            const float4* wv = reinterpret_cast<const float4*>(w_s + (t * Cin + ci) * Cout + cg * 8);
            const float4 a = wv[0], c = wv[1];
            const float v = xv[t][ci];
            acc[0] = fmaf(v, a.x, acc[0]); acc[1] = fmaf(v, a.y, acc[1]); acc[2] = fmaf(v, a.z, acc[2]); acc[3] = fmaf(v, a.w, acc[3]);
            acc[4] = fmaf(v, c.x, acc[4]); acc[5] = fmaf(v, c.y, acc[5]); acc[6] = fmaf(v, c.z, acc[6]); acc[7] = fmaf(v, c.w, acc[7]);
          }
    } else {
      for (int t = 0; t < taps; ++t) {
        const int hh = h_ + (ks == 3 ? t / 3 : 0) - r, ww = w_ + (ks == 3 ? t % 3 : 0) - r;
        if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
        const T* xp = x + (((long long)b_ * H + hh) * W + ww) * ldx;
        float xv[8];
#pragma unroll
        for (int ci = 0; ci < 8; ++ci) xv[ci] = ci < Cin ? to_f(xp[ci < Cin ? ci : 0]) : 0.f;
#pragma unroll
        for (int ci = 0; ci < 8; ++ci)
          if (ci < Cin) {
            const float4* wv = reinterpret_cast<const float4*>(w_s + (t * Cin + ci) * Cout + cg * 8);
            const float4 a = wv[0], c = wv[1];
            const float v = xv[ci];
            acc[0] = fmaf(v, a.x, acc[0]); acc[1] = fmaf(v, a.y, acc[1]); acc[2] = fmaf(v, a.z, acc[2]); acc[3] = fmaf(v, a.w, acc[3]);
            acc[4] = fmaf(v, c.x, acc[4]); acc[5] = fmaf(v, c.y, acc[5]); acc[6] = fmaf(v, c.z, acc[6]); acc[7] = fmaf(v, c.w, acc[7]);
          }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { csum[j] += acc[j]; csq[j] += acc[j] * acc[j]; acc[j] += bs[j]; }
    Vec8<T>::store(y + p * ldy + cg * 8, acc);
  }
  if (partials) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { red[threadIdx.x][j] = csum[j]; red[threadIdx.x][8 + j] = csq[j]; }
    __syncthreads();
    for (int o = threadIdx.x; o < 2 * Cout; o += 256) {
      const int which = o / Cout, c = o - which * Cout;
      float s = 0.f;
      for (int l = 0; l < lanes; ++l) s += red[l * CG + (c >> 3)][which * 8 + (c & 7)];
      partials[(size_t)blockIdx.x * 2 * Cout + o] = s;
    }
  }
}

// Pixel-quad variant of the narrow-K convolution (compile-time Cin <= 4, ks in {1,3}, W % 4 == 0): a thread
// owns 8 output channels of FOUR consecutive pixels of one image row, so every weight vector read from
// shared memory (2 x LDS.128) feeds 32 FMAs and the 3 x 6 input window is loaded once for the quad.
// FMA-bound at ~20 us for the 1 -> 64 first conv of cfg2; the 16-byte stores of the 8 threads of a pixel
// form one full 128-byte line.
template <typename T, int CIN, int KS>
__global__ void __launch_bounds__(256, KS == 1 ? 2 : 1)
k_conv_narrow_quad(const T* __restrict__ x, int ldx, const T* __restrict__ wp, const float* __restrict__ bias, T* __restrict__ y, int ldy,
                   int B, int H, int W, int Cout, float* __restrict__ partials) {
  constexpr int TAPS = KS * KS, K = TAPS * CIN, R = KS >> 1, WIN = 4 + 2 * R;
  extern __shared__ __align__(16) float nq_smem[];     // w_s[K][Cout], then red[256][16] when partials
  float* w_s = nq_smem;
  for (int i = threadIdx.x; i < K * Cout; i += 256) {
    const int k = i / Cout, co = i - k * Cout;
    w_s[i] = to_f(wp[(size_t)co * K + k]);
  }
  __syncthreads();
  const int CG = Cout >> 3, pgs = 256 / CG;
  const int cg = threadIdx.x % CG, pg = threadIdx.x / CG;
  const int Wq = W >> 2;
  const long long nquads = (long long)B * H * Wq;
  float bs[8], csum[8], csq[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { bs[j] = bias ? bias[cg * 8 + j] : 0.f; csum[j] = 0.f; csq[j] = 0.f; }
  for (long long q = (long long)blockIdx.x * pgs + pg; q < nquads; q += (long long)gridDim.x * pgs) {
    int wq, h_, b_;
    pix_decomp(q, Wq, H, b_, h_, wq);
    const int w0 = wq * 4;
    float acc[4][8];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[u][j] = 0.f;
    // one window row (WIN x CIN values) live at a time; the next row's loads are issued before this row's FMAs
    float xw[2][WIN][CIN];
    auto load_row = [&](int dy, float (&dst)[WIN][CIN]) {
      const int hh = h_ + dy - R;
      const bool hok = hh >= 0 && hh < H;
      const T* rowp = x + ((long long)(b_ * H + (hok ? hh : h_)) * W) * ldx;
#pragma unroll
      for (int c = 0; c < WIN; ++c) {
        const int ww = w0 + c - R;
        const bool ok = hok && ww >= 0 && ww < W;
        const T* xp = rowp + (long long)(ok ? ww : w0) * ldx;
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) dst[c][ci] = ok ? to_f(xp[ci]) : 0.f;
      }
    };
    load_row(0, xw[0]);
#pragma unroll
    for (int dy = 0; dy < KS; ++dy) {
      if (dy + 1 < KS) load_row(dy + 1, xw[(dy + 1) & 1]);
#pragma unroll
      for (int dx = 0; dx < KS; ++dx)
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) {
          const float4* wv = reinterpret_cast<const float4*>(w_s + ((dy * KS + dx) * CIN + ci) * Cout + cg * 8);
          const float4 a = wv[0], c = wv[1];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float v = xw[dy & 1][u + dx][ci];
            acc[u][0] = fmaf(v, a.x, acc[u][0]); acc[u][1] = fmaf(v, a.y, acc[u][1]); acc[u][2] = fmaf(v, a.z, acc[u][2]); acc[u][3] = fmaf(v, a.w, acc[u][3]);
            acc[u][4] = fmaf(v, c.x, acc[u][4]); acc[u][5] = fmaf(v, c.y, acc[u][5]); acc[u][6] = fmaf(v, c.z, acc[u][6]); acc[u][7] = fmaf(v, c.w, acc[u][7]);
          }
        }
    }
    T* yp = y + ((long long)(b_ * H + h_) * W + w0) * ldy + cg * 8;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { csum[j] += acc[u][j]; csq[j] = fmaf(acc[u][j], acc[u][j], csq[j]); acc[u][j] += bs[j]; }
      Vec8<T>::store(yp + (long long)u * ldy, acc[u]);
    }
  }
  if (partials) {
    float* red = w_s + K * Cout;                       // [256][16]
#pragma unroll
    for (int j = 0; j < 8; ++j) { red[threadIdx.x * 16 + j] = csum[j]; red[threadIdx.x * 16 + 8 + j] = csq[j]; }
    __syncthreads();
    for (int o = threadIdx.x; o < 2 * Cout; o += 256) {
      const int which = o / Cout, c = o - which * Cout;
      float s = 0.f;
      for (int l = 0; l < pgs; ++l) s += red[(l * CG + (c >> 3)) * 16 + which * 8 + (c & 7)];
      partials[(size_t)blockIdx.x * 2 * Cout + o] = s;
    }
  }
}

// ---- warp-level tensor-core variant of the first convolution (bf16, 3x3, Cin <= 4) ------------------------------
// K = 9*Cin <= 36 is far too short for a tcgen05 tile (the A operand would have to be materialised as an im2col
// buffer first), and the layer is HBM-bound (write 128 B per pixel) -- the CUDA-core kernel above needs 576 FMAs per
// pixel and ends up issue-bound at ~4x the memory time.  Here a warp owns 16 consecutive pixels of an image row and
// runs mma.sync.m16n8k16 (bf16 in, fp32 accumulate): the A fragment is gathered straight from the 3x3 neighbourhood
// with predicated 2-byte loads (the input is tiny and L1-resident), the weights live in registers as B fragments,
// the 16 x Cout tile is transposed through a warp-private shared tile and stored as full 16-byte vectors, and the
// BatchNorm partial sums accumulate per thread and are reduced once per block.
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(__nv_bfloat16 lo, __nv_bfloat16 hi) {
  return (uint32_t)__bfloat16_as_ushort(lo) | ((uint32_t)__bfloat16_as_ushort(hi) << 16);
}

template <int CIN, int NT>          // NT = Cout / 8
__global__ void __launch_bounds__(256)
k_conv_first_mma(const __nv_bfloat16* __restrict__ x, int ldx, const __nv_bfloat16* __restrict__ wp, const float* __restrict__ bias,
                 __nv_bfloat16* __restrict__ y, int ldy, int B, int H, int W, float* __restrict__ partials) {
  constexpr int K = 9 * CIN, KS = (K + 15) / 16, COUT = NT * 8, PITCH = COUT + 8;       // staging row pitch in bf16 (+16 B: bank spread)
  __shared__ __align__(16) __nv_bfloat16 stg[8][16 * PITCH];
  __shared__ float red[8][2 * COUT];
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int g = lane >> 2, tig = lane & 3;
  const __nv_bfloat16 zero = __float2bfloat16(0.f);
  // B fragments: B[k][n] = w[co = 8j + g][k]; b0 = {k = 16s + 2tig, +1}, b1 = {k = 16s + 2tig + 8, +9}
  uint32_t bf[NT][KS][2];
#pragma unroll
  for (int j = 0; j < NT; ++j)
#pragma unroll
    for (int s_ = 0; s_ < KS; ++s_)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int k0 = 16 * s_ + 2 * tig + 8 * hh;
        const __nv_bfloat16* wr = wp + (size_t)(8 * j + g) * K;
        bf[j][s_][hh] = pack_bf16(k0 < K ? wr[k0] : zero, k0 + 1 < K ? wr[k0 + 1] : zero);
      }
  float bs[NT][2];
#pragma unroll
  for (int j = 0; j < NT; ++j) { bs[j][0] = bias ? bias[8 * j + 2 * tig] : 0.f; bs[j][1] = bias ? bias[8 * j + 2 * tig + 1] : 0.f; }
  float ssum[NT][2], ssq[NT][2];
#pragma unroll
  for (int j = 0; j < NT; ++j) { ssum[j][0] = ssum[j][1] = 0.f; ssq[j][0] = ssq[j][1] = 0.f; }

  const int tiles_w = W >> 4;
  const long long ntiles = (long long)B * H * tiles_w;
  for (long long tl = (long long)blockIdx.x * 8 + wrp; tl < ntiles; tl += (long long)gridDim.x * 8) {
    int tw, h_, b_;
    pix_decomp(tl, tiles_w, H, b_, h_, tw);
    const int w0 = tw << 4;
    // A fragments: A[r][k] = x[b, h + dy - 1, w0 + r + dx - 1, ci], k = (dy*3 + dx)*CIN + ci
    uint32_t af[KS][4];
#pragma unroll
    for (int s_ = 0; s_ < KS; ++s_)
#pragma unroll
      for (int q = 0; q < 4; ++q) {              // q: (row half, k half) = a0:(g, lo) a1:(g+8, lo) a2:(g, hi) a3:(g+8, hi)
        const int r = g + 8 * (q & 1);
        __nv_bfloat16 v[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int k = 16 * s_ + 2 * tig + 8 * (q >> 1) + e;
          const int t = k / CIN, ci = k - t * CIN, dy = t / 3, dx = t - dy * 3;
          const int hh = h_ + dy - 1, ww = w0 + r + dx - 1;
          const bool ok = k < K && hh >= 0 && hh < H && ww >= 0 && ww < W;
          v[e] = ok ? x[((long long)(b_ * H + hh) * W + ww) * ldx + ci] : zero;
        }
        af[s_][q] = pack_bf16(v[0], v[1]);
      }
    float acc[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
      for (int s_ = 0; s_ < KS; ++s_) mma_bf16_16816(acc[j], af[s_], bf[j][s_][0], bf[j][s_][1]);
    }
    // statistics of the fp32 accumulators (before bias / rounding), then stage the bf16 tile
    __nv_bfloat16* st = stg[wrp];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      ssum[j][0] += acc[j][0] + acc[j][2]; ssum[j][1] += acc[j][1] + acc[j][3];
      ssq[j][0] += acc[j][0] * acc[j][0] + acc[j][2] * acc[j][2]; ssq[j][1] += acc[j][1] * acc[j][1] + acc[j][3] * acc[j][3];
      const __nv_bfloat162 lo = __floats2bfloat162_rn(acc[j][0] + bs[j][0], acc[j][1] + bs[j][1]);
      const __nv_bfloat162 hi = __floats2bfloat162_rn(acc[j][2] + bs[j][0], acc[j][3] + bs[j][1]);
      *reinterpret_cast<__nv_bfloat162*>(st + g * PITCH + 8 * j + 2 * tig) = lo;
      *reinterpret_cast<__nv_bfloat162*>(st + (g + 8) * PITCH + 8 * j + 2 * tig) = hi;
    }
    __syncwarp();
    __nv_bfloat16* yrow = y + ((long long)(b_ * H + h_) * W + w0) * ldy;
#pragma unroll
    for (int c = lane; c < 16 * NT; c += 32) {            // 16-byte chunks: NT per pixel row
      const int r = c / NT, cc = c - r * NT;
      *reinterpret_cast<uint4*>(yrow + (long long)r * ldy + cc * 8) = *reinterpret_cast<const uint4*>(st + r * PITCH + cc * 8);
    }
    __syncwarp();
  }
  if (partials) {
    // columns 8j + 2tig (+1) are shared by the 8 lanes with the same tig: xor-reduce over g, then over the 8 warps
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float a = ssum[j][e], q = ssq[j][e];
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
        if (g == 0) { red[wrp][8 * j + 2 * tig + e] = a; red[wrp][COUT + 8 * j + 2 * tig + e] = q; }
      }
    __syncthreads();
    for (int o = threadIdx.x; o < 2 * COUT; o += 256) {
      float s_ = 0.f;
#pragma unroll
      for (int w_ = 0; w_ < 8; ++w_) s_ += red[w_][o];
      partials[(size_t)blockIdx.x * 2 * COUT + o] = s_;
    }
  }
}

// Cout <= 8 outputs per pixel, Cin % 8 == 0: 8 threads per pixel each own one 16-byte channel chunk
template <typename T>
__global__ void __launch_bounds__(256)
k_conv_narrow_out(const T* __restrict__ x, int ldx, const T* __restrict__ wp, const float* __restrict__ bias, T* __restrict__ y, int ldy,
                  float* __restrict__ y_nchw, int B, int H, int W, int Cin, int Cout, int ks) {
  __shared__ float w_s[NARROW_MAX_W];                   // [tap][ci][co]
  const int taps = ks * ks, K = taps * Cin;
  for (int i = threadIdx.x; i < K * Cout; i += 256) {
    int k = i / Cout, co = i - k * Cout;
    w_s[i] = to_f(wp[(size_t)co * K + k]);
  }
  __syncthreads();
  const int sub = threadIdx.x & 7, slot = threadIdx.x >> 3;
  const long long M = (long long)B * H * W;
  const int r = ks >> 1, chunks = Cin >> 3;
  constexpr int UN = 4;                              // pixels per thread per iteration (memory-level parallelism)
  for (long long p0 = (long long)blockIdx.x * 32 * UN; p0 < M; p0 += (long long)gridDim.x * 32 * UN) {
    float acc[UN][8];
#pragma unroll
    for (int u = 0; u < UN; ++u)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[u][j] = 0.f;
    if (ks == 1 && chunks <= 8) {
      float xv[UN][8];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const long long p = p0 + u * 32 + slot;
        if (p < M && sub < chunks) Vec8<T>::load(x + p * ldx + sub * 8, xv[u]);
        else {
#pragma unroll
          for (int j = 0; j < 8; ++j) xv[u][j] = 0.f;
        }
      }
      const float* wv = w_s + (sub < chunks ? sub : 0) * 8 * Cout;
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int co = 0; co < 8; ++co)
          if (co < Cout) {
            const float wgt = wv[j * Cout + co];
#pragma unroll
            for (int u = 0; u < UN; ++u) acc[u][co] = fmaf(xv[u][j], wgt, acc[u][co]);
          }
    } else {
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const long long p = p0 + u * 32 + slot;
        if (p >= M) continue;
        int w_, h_, b_;
    pix_decomp(p, W, H, b_, h_, w_);
        for (int t = 0; t < taps; ++t) {
          const int hh = h_ + (ks == 3 ? t / 3 : 0) - r, ww = w_ + (ks == 3 ? t % 3 : 0) - r;
          if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
          const T* xp = x + (((long long)b_ * H + hh) * W + ww) * ldx;
          for (int c8 = sub; c8 < chunks; c8 += 8) {
            float xv[8];
            Vec8<T>::load(xp + c8 * 8, xv);
            const float* wv = w_s + (t * Cin + c8 * 8) * Cout;
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
              for (int co = 0; co < 8; ++co)
                if (co < Cout) acc[u][co] = fmaf(xv[j], wv[j * Cout + co], acc[u][co]);
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
#pragma unroll
      for (int co = 0; co < 8; ++co)
        if (co < Cout) {
          acc[u][co] += __shfl_xor_sync(0xffffffffu, acc[u][co], 1);
          acc[u][co] += __shfl_xor_sync(0xffffffffu, acc[u][co], 2);
          acc[u][co] += __shfl_xor_sync(0xffffffffu, acc[u][co], 4);
        }
      const long long p = p0 + u * 32 + slot;
      if (p < M && sub < Cout) {
        int w_, h_, b_;
    pix_decomp(p, W, H, b_, h_, w_);
        float v = 0.f;
#pragma unroll
        for (int co = 0; co < 8; ++co) v = (co == sub) ? acc[u][co] : v;
        if (bias) v += bias[sub];
        if (y_nchw) y_nchw[(((long long)b_ * Cout + sub) * H + h_) * W + w_] = v;
        else y[p * ldy + sub] = from_f<T>(v);
      }
    }
  }
}

// 1x1 logits head (Cin <= 64, compile-time COUT <= 4, H*W % 4 == 0), fp32 NCHW output: 8 threads share a
// pixel (one 16-byte channel chunk each, weights in registers), a thread handles 4 CONSECUTIVE pixels so the
// class planes are written with 16-byte stores; 3 xor-shuffles reduce the 8 chunk partials.
template <typename T, int COUT>
__global__ void __launch_bounds__(256)
k_head1x1_nchw(const T* __restrict__ x, int ldx, const T* __restrict__ wp, const float* __restrict__ bias, float* __restrict__ y_nchw,
               long long M, int HW, int Cin) {
  const int sub = threadIdx.x & 7, slot = threadIdx.x >> 3;
  const int chunks = Cin >> 3;
  float wreg[8][COUT];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int co = 0; co < COUT; ++co) wreg[j][co] = sub < chunks ? to_f(wp[(size_t)co * Cin + sub * 8 + j]) : 0.f;
  float bs[COUT];
#pragma unroll
  for (int co = 0; co < COUT; ++co) bs[co] = bias ? bias[co] : 0.f;
  for (long long p0 = (long long)blockIdx.x * 128 + slot * 4; p0 < M; p0 += (long long)gridDim.x * 128) {
    float xv[4][8];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (sub < chunks) Vec8<T>::load(x + (p0 + u) * ldx + sub * 8, xv[u]);
      else {
#pragma unroll
        for (int j = 0; j < 8; ++j) xv[u][j] = 0.f;
      }
    }
    float acc[4][COUT];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int co = 0; co < COUT; ++co) {
        float a = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) a = fmaf(xv[u][j], wreg[j][co], a);
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        a += __shfl_xor_sync(0xffffffffu, a, 2);
        a += __shfl_xor_sync(0xffffffffu, a, 4);
        acc[u][co] = a;
      }
    if (sub < COUT) {
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      float b = 0.f;
#pragma unroll
      for (int co = 0; co < COUT; ++co)
        if (co == sub) { o = make_float4(acc[0][co], acc[1][co], acc[2][co], acc[3][co]); b = bs[co]; }
      o.x += b; o.y += b; o.z += b; o.w += b;
      const long long img = p0 / HW, pin = p0 - img * HW;
      *reinterpret_cast<float4*>(y_nchw + (img * COUT + sub) * HW + pin) = o;
    }
  }
}

// acc[wc][tap][nc] = sum_q wide[q][wc] * narrow[q + sgn*tap][nc];  one row of partial sums per block.
// A thread owns 8 wide channels x TG taps x CN narrow channels (TG*CN <= 12): the 16-byte wide vector
// is loaded once per pixel and reused for all its taps.
template <typename T, int CN, int TG>
__global__ void __launch_bounds__(256, 2)
k_wgrad_narrow(const T* __restrict__ wide, int ldw, int Cw, const T* __restrict__ nar, int ldn, int ks, int sgn, int B, int H, int W,
               float* __restrict__ ws) {
  __shared__ float red[8192];                      // [per][8*TG*CN] block reduction over the pixel lanes
  constexpr int A = TG * CN;
  const int taps = ks * ks, WG = Cw >> 3, ngroups = taps / TG;
  const int per = WG * ngroups, lanes = 256 / per;
  const int lane = threadIdx.x / per;
  const bool active = lane < lanes;
  const int idx = threadIdx.x - lane * per;
  const int wg = idx % WG, tg = idx / WG;
  const long long M = (long long)B * H * W;
  float acc[8][A];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int n = 0; n < A; ++n) acc[j][n] = 0.f;
  if (active) {
    const long long step = (long long)gridDim.x * lanes;
#pragma unroll 2
    for (long long q = (long long)blockIdx.x * lanes + lane; q < M; q += step) {
      int w_, h_, b_;
      pix_decomp(q, W, H, b_, h_, w_);
      float wv[8], nv[A];
      Vec8<T>::load(wide + q * ldw + wg * 8, wv);
      const T* nc0 = nar + q * ldn;
#pragma unroll
      for (int tt = 0; tt < TG; ++tt) {
        const int t = tg * TG + tt;
        const int dh = sgn * ((ks == 3 ? t / 3 : 0) - (ks >> 1)), dw = sgn * ((ks == 3 ? t % 3 : 0) - (ks >> 1));
        const bool ok = h_ + dh >= 0 && h_ + dh < H && w_ + dw >= 0 && w_ + dw < W;
        const T* np = nc0 + (ok ? ((long long)dh * W + dw) * ldn : 0);
#pragma unroll
        for (int n = 0; n < CN; ++n) nv[tt * CN + n] = ok ? to_f(np[n]) : 0.f;
      }
#pragma unroll
      for (int n = 0; n < A; ++n)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j][n] = fmaf(wv[j], nv[n], acc[j][n]);
    }
  }
  // sequential rounds: lane l adds into shared memory, lane 0 ends up with the block total
  constexpr int width = 8 * A;
  for (int l = 0; l < lanes; ++l) {
    if (active && lane == l) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int n = 0; n < A; ++n) {
          float* p = &red[idx * width + j * A + n];
          *p = (l == 0 ? 0.f : *p) + acc[j][n];
        }
    }
    __syncthreads();
  }
  if (active && lane == 0) {
    float* row = ws + (size_t)blockIdx.x * ((size_t)Cw * taps * CN);
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int tt = 0; tt < TG; ++tt)
#pragma unroll
        for (int n = 0; n < CN; ++n)
          row[((size_t)(wg * 8 + j) * taps + (tg * TG + tt)) * CN + n] = red[idx * width + j * A + tt * CN + n];
  }
}

// ---- warp-level tensor-core variant of the narrow weight gradients (bf16) -----------------------------------------
// D[wc][n] = sum_q wide[q][wc] * narrow[q + sgn*tap(n)][nc(n)],  n = tap*CN + nc  (first conv: wide = dY, narrow = x;
// logits head: wide = x, narrow = dY).  A warp walks 16-pixel row segments: the 16 x Cw slab of the wide tensor goes
// through a warp-private shared tile (four 16-byte loads per lane) and comes back as m16k16 A fragments via
// ldmatrix.trans (rows = channels, K = pixels); the B fragments are gathered from the narrow tensor with predicated
// 2-byte loads; mma.sync.m16n8k16 accumulates in fp32 registers for the whole kernel.  The CUDA-core version spent
// 72 FMAs per 16 bytes of the wide tensor and ran at ~0.7 TB/s; this one is bound by streaming the wide tensor.
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_row) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem_row);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}

template <int CN, int TAPS, int MT>          // MT = Cw / 16 (m-tiles), N = TAPS*CN columns
__global__ void __launch_bounds__(256)
k_wgrad_narrow_mma(const __nv_bfloat16* __restrict__ wide, int ldw, const __nv_bfloat16* __restrict__ nar, int ldn, int sgn, int B, int H, int W,
                   float* __restrict__ ws) {
  constexpr int N = TAPS * CN, NKT = (N + 7) / 8, CW = MT * 16, PITCH = CW + 8, KSZ = TAPS == 9 ? 3 : 1;
  __shared__ __align__(16) __nv_bfloat16 stg[8][16 * PITCH];
  __shared__ float red[MT * 16 * NKT * 8];
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int g = lane >> 2, tig = lane & 3;
  const __nv_bfloat16 zero = __float2bfloat16(0.f);
  float acc[MT][NKT][4];
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int n = 0; n < NKT; ++n) acc[m][n][0] = acc[m][n][1] = acc[m][n][2] = acc[m][n][3] = 0.f;
  __nv_bfloat16* st = stg[wrp];
  const int tiles_w = W >> 4;
  const long long ntiles = (long long)B * H * tiles_w;
  const long long tstep = (long long)gridDim.x * 8;
  constexpr int CH = (2 * CW + 31) / 32;                    // 16-byte chunks of the 16 x CW slab per lane
  uint4 pre[CH];
  auto load_slab = [&](long long tile) {                     // global -> registers (in flight while the previous slab is consumed)
    int tw, h_, b_;
    pix_decomp(tile, tiles_w, H, b_, h_, tw);
    const __nv_bfloat16* wrow = wide + ((long long)(b_ * H + h_) * W + (tw << 4)) * ldw;
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int c = lane + 32 * i;
      if (c < 2 * CW) {
        const int r = c / (CW / 8), cc = c - r * (CW / 8);
        pre[i] = *reinterpret_cast<const uint4*>(wrow + (long long)r * ldw + cc * 8);
      }
    }
  };
  long long tl = (long long)blockIdx.x * 8 + wrp;
  if (tl < ntiles) load_slab(tl);
  for (; tl < ntiles; tl += tstep) {
    int tw, h_, b_;
    pix_decomp(tl, tiles_w, H, b_, h_, tw);
    const int w0 = tw << 4;
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int c = lane + 32 * i;
      if (c < 2 * CW) {
        const int r = c / (CW / 8), cc = c - r * (CW / 8);
        *reinterpret_cast<uint4*>(st + r * PITCH + cc * 8) = pre[i];
      }
    }
    if (tl + tstep < ntiles) load_slab(tl + tstep);
    // B fragments: B[p][n] = nar[b, h + sgn*(dy-1), w0 + p + sgn*(dx-1), nc]; b0: p = 2tig, 2tig+1; b1: p = 2tig+8, +9; n = 8*nt + g
    uint32_t bfr[NKT][2];
#pragma unroll
    for (int nt = 0; nt < NKT; ++nt) {
      const int n = 8 * nt + g;
      const int t = n / CN, nc = n - t * CN;
      const int dy = KSZ == 3 ? t / 3 - 1 : 0, dx = KSZ == 3 ? t % 3 - 1 : 0;
      const int hh = h_ + sgn * dy;
      const bool rowok = n < N && hh >= 0 && hh < H;
      const __nv_bfloat16* nrow = nar + ((long long)(b_ * H + (rowok ? hh : h_)) * W) * ldn + nc;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        __nv_bfloat16 v[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int ww = w0 + 2 * tig + 8 * hf + e + sgn * dx;
          v[e] = (rowok && ww >= 0 && ww < W) ? nrow[(long long)ww * ldn] : zero;
        }
        bfr[nt][hf] = pack_bf16(v[0], v[1]);
      }
    }
    __syncwarp();
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      // A fragment (rows = channels 16m.., K = 16 pixels) from the [pixel][channel] tile: four transposed 8x8 matrices
      uint32_t af[4];
      const int mat = lane >> 3, rr = lane & 7;
      ldmatrix_x4_trans(af, st + (rr + 8 * (mat >> 1)) * PITCH + 16 * m + 8 * (mat & 1));
#pragma unroll
      for (int nt = 0; nt < NKT; ++nt) mma_bf16_16816(acc[m][nt], af, bfr[nt][0], bfr[nt][1]);
    }
    __syncwarp();
  }
  // block total: the 8 warps add their fragments through shared memory (fixed order: deterministic)
  for (int o = threadIdx.x; o < MT * 16 * NKT * 8; o += 256) red[o] = 0.f;
  __syncthreads();
  for (int w_ = 0; w_ < 8; ++w_) {
    if (wrp == w_) {
#pragma unroll
      for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int nt = 0; nt < NKT; ++nt) {
          float* base = red + ((16 * m) * NKT * 8) + 8 * nt;
          base[(g)*NKT * 8 + 2 * tig] += acc[m][nt][0];
          base[(g)*NKT * 8 + 2 * tig + 1] += acc[m][nt][1];
          base[(g + 8) * NKT * 8 + 2 * tig] += acc[m][nt][2];
          base[(g + 8) * NKT * 8 + 2 * tig + 1] += acc[m][nt][3];
        }
    }
    __syncthreads();
  }
  float* row = ws + (size_t)blockIdx.x * ((size_t)CW * N);
  for (int o = threadIdx.x; o < CW * N; o += 256) {
    const int wc = o / N, n = o - wc * N;
    row[o] = red[wc * NKT * 8 + n];
  }
}

// mode 0: dw[(wc*Cn + nc)*taps + t]  (wide = output channels);  mode 1: dw[(nc*Cw + wc)*taps + t]; one warp per element
__global__ void k_wgrad_narrow_reduce(const float* __restrict__ ws, int rows, int Cw, int taps, int Cn, int mode, float* __restrict__ dw,
                                      int accumulate) {
  const int n = Cw * taps * Cn;
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= n) return;
  double s = 0.0;
  for (int r = lane; r < rows; r += 32) s += (double)ws[(size_t)r * n + i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane != 0) return;
  const int nc = i % Cn, t = (i / Cn) % taps, wc = i / (Cn * taps);
  const size_t o = mode == 0 ? ((size_t)wc * Cn + nc) * taps + t : ((size_t)nc * Cw + wc) * taps + t;
  dw[o] = (accumulate ? dw[o] : 0.f) + (float)s;
}

// ---- host-side eligibility + launches ---------------------------------------------------------
static inline bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

bool narrow_in_ok(int Cin, int Cout, int ks) {
  return Cin * ks * ks <= 72 && Cin <= 8 && Cout % 8 == 0 && pow2(Cout / 8) && Cout / 8 <= 256 && Cin * ks * ks * Cout <= NARROW_MAX_W;
}
bool narrow_out_ok(int Cin, int Cout, int ks) { return Cout <= 8 && Cin % 8 == 0 && Cin * ks * ks * Cout <= NARROW_MAX_W; }
static int narrow_wgrad_grid(long long M, int lanes) {
  long long g = (M + lanes - 1) / lanes;          // exactly one wave: 2 resident blocks per SM (__launch_bounds__(256, 2))
  const long long cap = 148 * 2;
  if (g > cap) g = cap;
  return (int)(g < 1 ? 1 : g);
}
static int narrow_tg(int Cn, int ks) { return ks == 1 ? 1 : (Cn == 1 ? 9 : 3); }
bool narrow_wgrad_ok(int Cw, int Cn, int ks) {
  if (!(Cn >= 1 && Cn <= 4 && Cw % 8 == 0 && (ks == 1 || ks == 3))) return false;
  const int per = (Cw / 8) * (ks * ks / narrow_tg(Cn, ks));
  return per <= 128 && per * 8 * narrow_tg(Cn, ks) * Cn <= 8192;
}
long long narrow_wgrad_ws_bytes(long long M, int Cw, int Cn, int ks) {
  const int per = (Cw / 8) * (ks * ks / narrow_tg(Cn, ks)), lanes = 256 / per;
  return (long long)narrow_wgrad_grid(M, lanes) * Cw * ks * ks * Cn * (long long)sizeof(float);
}

template <typename T, int CIN, int KS>
static int narrow_quad_launch(const void* x, int ldx, const void* w, const float* bias, void* y, int ldy, int B, int H, int W, int Cout,
                              float* partials, int* nparts_host, cudaStream_t st) {
  const int pgs = 256 / (Cout / 8);
  const long long nquads = (long long)B * H * (W / 4);
  long long g = (nquads + pgs - 1) / pgs;
  if (g > 148 * 8) g = 148 * 8;                          // <= USTRUN_MAX_PARTS partial-stat rows
  if (nparts_host) *nparts_host = (int)g;
  const size_t smem = ((size_t)KS * KS * CIN * Cout + (partials ? 256 * 16 : 0)) * sizeof(float);
  k_conv_narrow_quad<T, CIN, KS><<<(int)g, 256, smem, st>>>((const T*)x, ldx, (const T*)w, bias, (T*)y, ldy, B, H, W, Cout, partials);
  return check_launch("conv_narrow_quad");
}

template <int CIN>
static int first_mma_launch(const void* x, int ldx, const void* w, const float* bias, void* y, int ldy, int B, int H, int W, int Cout,
                            float* partials, int* nparts_host, cudaStream_t st) {
  const long long ntiles = (long long)B * H * (W / 16);
  long long g = (ntiles + 7) / 8;
  if (g > 148 * 4) g = 148 * 4;
  if (nparts_host) *nparts_host = (int)g;
#define FM(NTV) k_conv_first_mma<CIN, NTV><<<(int)g, 256, 0, st>>>((const __nv_bfloat16*)x, ldx, (const __nv_bfloat16*)w, bias, (__nv_bfloat16*)y, ldy, B, H, W, partials)
  if (Cout == 64) FM(8);
  else if (Cout == 32) FM(4);
  else FM(2);
#undef FM
  return check_launch("conv_first_mma");
}
static int first_mma_mode() {      // USTRUN_FIRST_MMA=0 keeps the CUDA-core quad kernel (A/B comparisons)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("USTRUN_FIRST_MMA");
    v = e ? atoi(e) : 1;
  }
  return v;
}

template <typename T>
int narrow_in_launch(const void* x, int ldx, const void* w, const float* bias, void* y, int ldy, int B, int H, int W, int Cin, int Cout, int ks,
                     float* partials, int* nparts_host, cudaStream_t st) {
  if (sizeof(T) == 2 && ks == 3 && W % 16 == 0 && Cin >= 1 && Cin <= 4 && (Cout == 64 || Cout == 32 || Cout == 16) && ldy % 8 == 0 && first_mma_mode() > 0) {
    switch (Cin) {
      case 1: return first_mma_launch<1>(x, ldx, w, bias, y, ldy, B, H, W, Cout, partials, nparts_host, st);
      case 2: return first_mma_launch<2>(x, ldx, w, bias, y, ldy, B, H, W, Cout, partials, nparts_host, st);
      case 3: return first_mma_launch<3>(x, ldx, w, bias, y, ldy, B, H, W, Cout, partials, nparts_host, st);
      default: return first_mma_launch<4>(x, ldx, w, bias, y, ldy, B, H, W, Cout, partials, nparts_host, st);
    }
  }
  if (W % 4 == 0 && Cin >= 1 && Cin <= 4 && Cout <= 256 && ((size_t)ks * ks * Cin * Cout + 256 * 16) * sizeof(float) <= 48 * 1024) {
#define NQ(CI, KSV) return narrow_quad_launch<T, CI, KSV>(x, ldx, w, bias, y, ldy, B, H, W, Cout, partials, nparts_host, st)
    if (ks == 3) { switch (Cin) { case 1: NQ(1, 3); case 2: NQ(2, 3); case 3: NQ(3, 3); default: NQ(4, 3); } }
    else { switch (Cin) { case 1: NQ(1, 1); case 2: NQ(2, 1); case 3: NQ(3, 1); default: NQ(4, 1); } }
#undef NQ
  }
  const int lanes = 256 / (Cout / 8);
  long long M = (long long)B * H * W;
  long long g = (M + lanes - 1) / lanes;
  if (g > USTRUN_MAX_PARTS) g = USTRUN_MAX_PARTS;       // partial-stat rows = blocks
  if (nparts_host) *nparts_host = (int)g;
  k_conv_narrow_in<T, 0><<<(int)g, 256, 0, st>>>((const T*)x, ldx, (const T*)w, bias, (T*)y, ldy, B, H, W, Cin, Cout, ks, partials);
  return check_launch("conv_narrow_in");
}
template <typename T>
int narrow_out_launch(const void* x, int ldx, const void* w, const float* bias, void* y, int ldy, float* y_nchw, int B, int H, int W, int Cin,
                      int Cout, int ks, cudaStream_t st) {
  long long M = (long long)B * H * W;
  if (ks == 1 && y_nchw && Cout <= 4 && Cin <= 64 && (H * W) % 4 == 0) {
    long long g = (M + 127) / 128;
    if (g > 148 * 8) g = 148 * 8;
#define HL(CO) k_head1x1_nchw<T, CO><<<(int)g, 256, 0, st>>>((const T*)x, ldx, (const T*)w, bias, y_nchw, M, H * W, Cin)
    switch (Cout) { case 1: HL(1); break; case 2: HL(2); break; case 3: HL(3); break; default: HL(4); break; }
#undef HL
    return check_launch("head1x1_nchw");
  }
  long long g = (M + 127) / 128;
  if (g > 148 * 16) g = 148 * 16;
  k_conv_narrow_out<T><<<(int)g, 256, 0, st>>>((const T*)x, ldx, (const T*)w, bias, (T*)y, ldy, y_nchw, B, H, W, Cin, Cout, ks);
  return check_launch("conv_narrow_out");
}
static int narrow_mma_mode() {      // USTRUN_NARROW_MMA=0 keeps the CUDA-core narrow weight-gradient kernel
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("USTRUN_NARROW_MMA");
    v = e ? atoi(e) : 1;
  }
  return v;
}
template <int CN, int TAPS>
static int narrow_wgrad_mma_launch(const void* wide, int ldw, int Cw, const void* nar, int ldn, int sgn, int B, int H, int W, float* ws, int grid,
                                   cudaStream_t st) {
#define NM(MTV) k_wgrad_narrow_mma<CN, TAPS, MTV><<<grid, 256, 0, st>>>((const __nv_bfloat16*)wide, ldw, (const __nv_bfloat16*)nar, ldn, sgn, B, H, W, ws)
  if (Cw == 64) NM(4);
  else if (Cw == 32) NM(2);
  else NM(1);
#undef NM
  return check_launch("wgrad_narrow_mma");
}

template <typename T>
int narrow_wgrad_launch(const void* wide, int ldw, int Cw, const void* nar, int ldn, int Cn, int ks, int sgn, int mode, int B, int H, int W,
                        float* dw, int accumulate, void* workspace, long long ws_bytes, cudaStream_t st) {
  long long M = (long long)B * H * W;
  const int TGv = narrow_tg(Cn, ks);
  const int per = (Cw / 8) * (ks * ks / TGv), lanes = 256 / per;
  int grid = narrow_wgrad_grid(M, lanes);
  long long need = narrow_wgrad_ws_bytes(M, Cw, Cn, ks);
  if (!workspace || ws_bytes < need) { set_error("narrow wgrad: workspace too small (%lld < %lld)", ws_bytes, need); return USTRUN_ERR_ARG; }
  const bool mma_ok = sizeof(T) == 2 && narrow_mma_mode() > 0 && W % 16 == 0 && ldw % 8 == 0 && (Cw == 64 || Cw == 32 || Cw == 16) && Cn >= 1 && Cn <= 4 &&
                      (ks == 1 || ks == 3);
  if (mma_ok) {
    long long tiles = M / 16;
    long long g = (tiles + 7) / 8;
    if (g > grid) g = grid;                         // the workspace holds `grid` partial rows
    grid = (int)(g < 1 ? 1 : g);
    int rc;
#define NMM(CNV) (ks == 3 ? narrow_wgrad_mma_launch<CNV, 9>(wide, ldw, Cw, nar, ldn, sgn, B, H, W, (float*)workspace, grid, st) \
                          : narrow_wgrad_mma_launch<CNV, 1>(wide, ldw, Cw, nar, ldn, sgn, B, H, W, (float*)workspace, grid, st))
    switch (Cn) { case 1: rc = NMM(1); break; case 2: rc = NMM(2); break; case 3: rc = NMM(3); break; default: rc = NMM(4); break; }
#undef NMM
    if (rc) return rc;
  } else {
#define NW_LAUNCH(CNV, TGV) k_wgrad_narrow<T, CNV, TGV><<<grid, 256, 0, st>>>((const T*)wide, ldw, Cw, (const T*)nar, ldn, ks, sgn, B, H, W, (float*)workspace)
    if (ks == 1) {
      switch (Cn) { case 1: NW_LAUNCH(1, 1); break; case 2: NW_LAUNCH(2, 1); break; case 3: NW_LAUNCH(3, 1); break; default: NW_LAUNCH(4, 1); break; }
    } else if (Cn == 1) NW_LAUNCH(1, 9);
    else if (Cn == 2) NW_LAUNCH(2, 3);
    else if (Cn == 3) NW_LAUNCH(3, 3);
    else NW_LAUNCH(4, 3);
#undef NW_LAUNCH
    int rc = check_launch("wgrad_narrow");
    if (rc) return rc;
  }
  const int n = Cw * ks * ks * Cn;
  k_wgrad_narrow_reduce<<<(n + 7) / 8, 256, 0, st>>>((const float*)workspace, grid, Cw, ks * ks, Cn, mode, dw, accumulate);
  return check_launch("wgrad_narrow_reduce");
}

#define INST(T)                                                                                                                              \
  template int narrow_in_launch<T>(const void*, int, const void*, const float*, void*, int, int, int, int, int, int, int, float*, int*,     \
                                   cudaStream_t);                                                                                            \
  template int narrow_out_launch<T>(const void*, int, const void*, const float*, void*, int, float*, int, int, int, int, int, int,           \
                                    cudaStream_t);                                                                                           \
  template int narrow_wgrad_launch<T>(const void*, int, int, const void*, int, int, int, int, int, int, int, int, float*, int, void*,        \
                                      long long, cudaStream_t);
INST(float)
INST(__nv_bfloat16)

}  // namespace ustrun
