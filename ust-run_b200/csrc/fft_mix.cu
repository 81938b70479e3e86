// Frequency-domain style mix on the device (reference train.py:158-207 called from
// train.py:628-636): for every sample the amplitude spectrum of the SOURCE image is blended with the amplitude
// of the TARGET image inside the centred low-frequency window |fh|,|fw| <= b, b = floor(min(H,W)*L), the source
// phase is kept, and the inverse transform is clipped to [0,255] and mapped back to [-1,1].
//
// The reference does two full fft2 + one ifft2 per sample on the HOST (D2H copy, numpy, H2D copy).  Only the
// (2b+1)^2 window bins change (b = 3 for the default --LB 0.01 at 384x384), and the transform is linear, so
//     out = src + Re ifft2( delta ),   delta = ratio * (|F_trg| - |F_src|) * F_src/|F_src|  on the window, else 0
// needs the forward DFT of both images at the window bins only (two thin separable passes) and a thin inverse.
// Everything runs in float64 like numpy's pocketfft (the work is ~10 MFLOP per image), so the result agrees with
// the reference to float32 rounding.  No host synchronisation, no cuFFT.
#include "common.cuh"

namespace ustrun {

constexpr int FFT_MAXB = 12;                 // window half-width supported (2b+1 <= 25)
constexpr int FFT_MAXNB = 2 * FFT_MAXB + 1;

// twiddles: tw[k] = (cos(2 pi k / n), sin(2 pi k / n)), k in [0, n)
__global__ void k_fft_twiddles(double2* __restrict__ tw, int n) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  double s, c;
  sincospi(2.0 * (double)k / (double)n, &s, &c);
  tw[k] = make_double2(c, s);
}

// pass 1 (rows): T[img][h][j] = sum_w v[h][w] * exp(-2 pi i fw w / W), fw = j - b, v = (x + 1) * 127.5 (fp32, as the
// reference forms it before handing the array to numpy).  One block per image row, both images.
__global__ void __launch_bounds__(128)
k_fft_rows(const float* __restrict__ src, const float* __restrict__ trg, const double2* __restrict__ twW, int H, int W, int b,
           double2* __restrict__ Ts, double2* __restrict__ Tt) {
  extern __shared__ double fsm[];            // vs[W], vt[W]
  double* vs = fsm;
  double* vt = fsm + W;
  const long long row = blockIdx.x;          // img * H + h, img = n * C + c
  const float* ps = src + row * W;
  const float* pt = trg + row * W;
  for (int w = threadIdx.x; w < W; w += blockDim.x) {
    vs[w] = (double)((ps[w] + 1.f) * 127.5f);
    vt[w] = (double)((pt[w] + 1.f) * 127.5f);
  }
  __syncthreads();
  const int nb = 2 * b + 1;
  // 128 threads: (j, part) = nb columns x (128 / nb) partial sums over w, combined through shared memory
  const int parts = blockDim.x / nb;
  const int j = threadIdx.x % nb, part = threadIdx.x / nb;
  double sr = 0.0, si = 0.0, tr = 0.0, ti = 0.0;
  if (part < parts) {
    const int fw = j - b;
    for (int w = part; w < W; w += parts) {
      int k = (int)(((long long)fw * w) % W);
      if (k < 0) k += W;
      const double2 t = twW[k];              // exp(-i a) = cos a - i sin a
      sr += vs[w] * t.x; si -= vs[w] * t.y;
      tr += vt[w] * t.x; ti -= vt[w] * t.y;
    }
  }
  __syncthreads();
  double* red = fsm;                         // [4][128]
  red[threadIdx.x] = sr; red[128 + threadIdx.x] = si; red[256 + threadIdx.x] = tr; red[384 + threadIdx.x] = ti;
  __syncthreads();
  if (threadIdx.x < nb) {
    double a = 0.0, c = 0.0, d = 0.0, e = 0.0;
    for (int p = 0; p < parts; ++p) {
      a += red[p * nb + threadIdx.x]; c += red[128 + p * nb + threadIdx.x];
      d += red[256 + p * nb + threadIdx.x]; e += red[384 + p * nb + threadIdx.x];
    }
    Ts[row * nb + threadIdx.x] = make_double2(a, c);
    Tt[row * nb + threadIdx.x] = make_double2(d, e);
  }
}

// pass 2 (columns) + amplitude blend: F[i][j] = sum_h exp(-2 pi i fh h / H) T[h][j];
// delta = ratio * (|F_trg| - |F_src|) * F_src / |F_src|   (np.angle(0) = 0 => unit phase 1 when |F_src| == 0)
__global__ void k_fft_cols_delta(const double2* __restrict__ Ts, const double2* __restrict__ Tt, const double2* __restrict__ twH,
                                 const double* __restrict__ ratio, int C, int H, int b, double2* __restrict__ delta) {
  const int nb = 2 * b + 1;
  const int img = blockIdx.x;
  const int i = threadIdx.x / nb, j = threadIdx.x % nb;
  if (i >= nb) return;
  const int fh = i - b;
  double sr = 0.0, si = 0.0, tr = 0.0, ti = 0.0;
  for (int h = 0; h < H; ++h) {
    int k = (int)(((long long)fh * h) % H);
    if (k < 0) k += H;
    const double2 t = twH[k];
    const double2 a = Ts[((long long)img * H + h) * nb + j], c = Tt[((long long)img * H + h) * nb + j];
    // (a.x + i a.y) * (t.x - i t.y)
    sr += a.x * t.x + a.y * t.y; si += a.y * t.x - a.x * t.y;
    tr += c.x * t.x + c.y * t.y; ti += c.y * t.x - c.x * t.y;
  }
  const double as = hypot(sr, si), at = hypot(tr, ti);
  const double r = ratio[img / C];
  const double g = r * (at - as);
  const double ur = as > 0.0 ? sr / as : 1.0, ui = as > 0.0 ? si / as : 0.0;
  delta[((long long)img * nb + i) * nb + j] = make_double2(g * ur, g * ui);
}

// inverse pass 1 (columns): U[h][j] = sum_i delta[i][j] * exp(+2 pi i fh h / H)
__global__ void k_ifft_cols(const double2* __restrict__ delta, const double2* __restrict__ twH, long long imgs, int H, int b, double2* __restrict__ U) {
  const int nb = 2 * b + 1;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;      // (img * H + h) * nb + j
  const long long rows = idx / nb;
  const int j = (int)(idx - rows * nb);
  const long long img = rows / H;
  const int h = (int)(rows - img * H);
  if (img >= imgs) return;
  double ur = 0.0, ui = 0.0;
  for (int i = 0; i < nb; ++i) {
    int k = (int)(((long long)(i - b) * h) % H);
    if (k < 0) k += H;
    const double2 t = twH[k], d = delta[(img * nb + i) * nb + j];
    ur += d.x * t.x - d.y * t.y; ui += d.x * t.y + d.y * t.x;
  }
  U[idx] = make_double2(ur, ui);
}

// inverse pass 2 (rows) + output: out = clip(src255 + Re sum_j U[h][j] exp(+2 pi i fw w / W) / (H W), 0, 255) / 127.5 - 1
__global__ void __launch_bounds__(128)
k_ifft_rows_out(const float* __restrict__ src, const double2* __restrict__ U, const double2* __restrict__ twW, int H, int W, int b,
                float* __restrict__ out) {
  __shared__ double2 u[FFT_MAXNB];
  const int nb = 2 * b + 1;
  const long long row = blockIdx.x;
  if (threadIdx.x < nb) u[threadIdx.x] = U[row * nb + threadIdx.x];
  __syncthreads();
  const double inv = 1.0 / ((double)H * (double)W);
  for (int w = threadIdx.x; w < W; w += blockDim.x) {
    double acc = 0.0;
    for (int j = 0; j < nb; ++j) {
      int k = (int)(((long long)(j - b) * w) % W);
      if (k < 0) k += W;
      const double2 t = twW[k];
      acc += u[j].x * t.x - u[j].y * t.y;
    }
    double v = (double)((src[row * W + w] + 1.f) * 127.5f) + acc * inv;
    v = v < 0.0 ? 0.0 : (v > 255.0 ? 255.0 : v);
    out[row * W + w] = __fdiv_rn((float)v, 127.5f) - 1.f;       // .astype(float32), then tensor / 127.5 - 1 in fp32
  }
}

}  // namespace ustrun

using namespace ustrun;

extern "C" {

long long ustrun_fft_amp_mix_workspace_bytes(int N, int C, int H, int W, double L) {
  const int b = (int)floor((double)(H < W ? H : W) * L);      // numpy: np.floor(np.amin((h, w)) * L) in float64
  const long long nb = 2LL * b + 1, imgs = (long long)N * C;
  // twW[W], twH[H], Ts, Tt, U: [imgs][H][nb], delta: [imgs][nb][nb]   (double2 each)
  return (long long)sizeof(double2) * (W + H + 3 * imgs * H * nb + imgs * nb * nb);
}

int ustrun_fft_amp_mix(const float* src, const float* trg, const double* ratio, double L, float* out, int N, int C, int H, int W,
                       void* workspace, long long ws_bytes, void* stream) {
  USTRUN_REQUIRE(src && trg && ratio && out && N > 0 && C > 0 && H > 1 && W > 1 && workspace, "fft_amp_mix: bad args");
  const int b = (int)floor((double)(H < W ? H : W) * L);      // exactly numpy's float64 product (train.py:170)
  USTRUN_REQUIRE(b >= 0 && b <= FFT_MAXB && 2 * b + 1 <= H && 2 * b + 1 <= W, "fft_amp_mix: window half-width %d unsupported (max %d)", b, FFT_MAXB);
  USTRUN_REQUIRE(ws_bytes >= ustrun_fft_amp_mix_workspace_bytes(N, C, H, W, L), "fft_amp_mix: workspace too small");
  const int nb = 2 * b + 1;
  const long long imgs = (long long)N * C;
  USTRUN_REQUIRE(imgs * H < (1LL << 31), "fft_amp_mix: too many rows");
  cudaStream_t st = (cudaStream_t)stream;
  double2* twW = (double2*)workspace;
  double2* twH = twW + W;
  double2* Ts = twH + H;
  double2* Tt = Ts + imgs * H * nb;
  double2* U = Tt + imgs * H * nb;
  double2* delta = U + imgs * H * nb;
  k_fft_twiddles<<<ceil_div(W, 128), 128, 0, st>>>(twW, W);
  k_fft_twiddles<<<ceil_div(H, 128), 128, 0, st>>>(twH, H);
  size_t smem = sizeof(double) * (size_t)(2 * W > 512 ? 2 * W : 512);
  k_fft_rows<<<(unsigned)(imgs * H), 128, smem, st>>>(src, trg, twW, H, W, b, Ts, Tt);
  k_fft_cols_delta<<<(unsigned)imgs, ((nb * nb + 31) / 32) * 32, 0, st>>>(Ts, Tt, twH, ratio, C, H, b, delta);
  k_ifft_cols<<<(unsigned)ceil_div((long long)H * nb * imgs, 128), 128, 0, st>>>(delta, twH, imgs, H, b, U);
  k_ifft_rows_out<<<(unsigned)(imgs * H), 128, 0, st>>>(src, U, twW, H, W, b, out);
  return check_launch("fft_amp_mix");
}

}  // extern "C"
