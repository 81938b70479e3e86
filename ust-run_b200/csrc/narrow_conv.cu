// HBM-bound convolutions with one very narrow side (SURVEY App. B marks them "HBM"):
//   * narrow-K  : first conv of the network (Cin = 1..4 image channels, K = 9*Cin) and the dgrad
//                 of the logits head (K = taps * n_classes)           -> k_conv_narrow_in
//   * narrow-N  : the logits head itself (Cout = n_classes <= 8)       -> k_conv_narrow_out
//   * their weight gradients (one operand 8..256 channels wide, the other <= 8) -> k_wgrad_narrow
// A tensor-core tile would be >90 % padding here; these kernels stream the wide tensor once with
// 128-bit accesses and keep the small operand / weights in shared memory or registers.
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

template <typename T>
int narrow_in_launch(const void* x, int ldx, const void* w, const float* bias, void* y, int ldy, int B, int H, int W, int Cin, int Cout, int ks,
                     float* partials, int* nparts_host, cudaStream_t st) {
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
template <typename T>
int narrow_wgrad_launch(const void* wide, int ldw, int Cw, const void* nar, int ldn, int Cn, int ks, int sgn, int mode, int B, int H, int W,
                        float* dw, int accumulate, void* workspace, long long ws_bytes, cudaStream_t st) {
  long long M = (long long)B * H * W;
  const int TGv = narrow_tg(Cn, ks);
  const int per = (Cw / 8) * (ks * ks / TGv), lanes = 256 / per;
  const int grid = narrow_wgrad_grid(M, lanes);
  long long need = narrow_wgrad_ws_bytes(M, Cw, Cn, ks);
  if (!workspace || ws_bytes < need) { set_error("narrow wgrad: workspace too small (%lld < %lld)", ws_bytes, need); return USTRUN_ERR_ARG; }
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
