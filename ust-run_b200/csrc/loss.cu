// Pseudo-label pass (K12), CE+Dice loss forward/backward (K13/K14), image CutMix blend.
// All HBM-bound: logits are fp32 NCHW, so for a fixed class consecutive threads read consecutive
// pixels (float4 = 128-bit per thread); label/mask planes are uint8 (uchar4 per thread).
// Arithmetic mirrors ATen so that masks/argmax are bit-exact given identical logits:
// softmax = expf(x - max) / sum (sequential class order, IEEE division, no fast-math), argmax over
// the PROBABILITIES with first-index tie-break, strict '>' against float(threshold).
#include <float.h>

#include "common.cuh"

namespace ustrun {

constexpr int MAXC = 8;

template <int C>
__device__ __forceinline__ void softmax_argmax(const float (&x)[MAXC], float& conf, int& arg, float (&p)[MAXC]) {
  float m = x[0];
#pragma unroll
  for (int c = 1; c < C; ++c) m = fmaxf(m, x[c]);
  float e[MAXC];
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) { e[c] = expf(x[c] - m); s += e[c]; }
  conf = -1.f; arg = 0;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    p[c] = e[c] / s;
    if (p[c] > conf) { conf = p[c]; arg = c; }
  }
}

struct PLArgs {
  const float *t1, *t2, *t3, *s0;
  const uint8_t *box, *cut_label, *cut_mask;
  const int* choice;
  float thr, thr_lo;
  int Bu, C, HW;
  uint8_t *pl, *mask, *pl_w, *mask_w, *pl_ul, *mask_ul, *pl_lu, *mask_lu, *stu_pl;
};

template <int C, int V>   // V pixels per thread (4 when HW % 4 == 0, else 1)
__global__ void __launch_bounds__(256) k_pseudo_label_softmax(PLArgs a) {
  const long long nvec = (long long)a.Bu * a.HW / V;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i * V;
    const int b = (int)((unsigned int)pix / (unsigned int)a.HW);
    const int hw = (int)(pix - (long long)b * a.HW);
    const size_t lbase = (size_t)b * C * a.HW + hw;
    float x1[V][MAXC], x2[V][MAXC], x3[V][MAXC], xs[V][MAXC];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      if (V == 4) {
        float4 v1 = *reinterpret_cast<const float4*>(a.t1 + lbase + (size_t)c * a.HW);
        float4 v2 = *reinterpret_cast<const float4*>(a.t2 + lbase + (size_t)c * a.HW);
        float4 v3 = *reinterpret_cast<const float4*>(a.t3 + lbase + (size_t)c * a.HW);
        x1[0][c] = v1.x; x1[1 % V][c] = v1.y; x1[2 % V][c] = v1.z; x1[3 % V][c] = v1.w;
        x2[0][c] = v2.x; x2[1 % V][c] = v2.y; x2[2 % V][c] = v2.z; x2[3 % V][c] = v2.w;
        x3[0][c] = v3.x; x3[1 % V][c] = v3.y; x3[2 % V][c] = v3.z; x3[3 % V][c] = v3.w;
        if (a.s0) {
          float4 v4 = *reinterpret_cast<const float4*>(a.s0 + lbase + (size_t)c * a.HW);
          xs[0][c] = v4.x; xs[1 % V][c] = v4.y; xs[2 % V][c] = v4.z; xs[3 % V][c] = v4.w;
        }
      } else {
        x1[0][c] = a.t1[lbase + (size_t)c * a.HW];
        x2[0][c] = a.t2[lbase + (size_t)c * a.HW];
        x3[0][c] = a.t3[lbase + (size_t)c * a.HW];
        if (a.s0) xs[0][c] = a.s0[lbase + (size_t)c * a.HW];
      }
    }
    const int ch = a.choice[b];
    uint8_t bx[V], cl[V], cm[V];
    if (V == 4) {
      uchar4 v = *reinterpret_cast<const uchar4*>(a.box + pix);
      bx[0] = v.x; bx[1 % V] = v.y; bx[2 % V] = v.z; bx[3 % V] = v.w;
      v = *reinterpret_cast<const uchar4*>(a.cut_label + (size_t)ch * a.HW + hw);
      cl[0] = v.x; cl[1 % V] = v.y; cl[2 % V] = v.z; cl[3 % V] = v.w;
      v = *reinterpret_cast<const uchar4*>(a.cut_mask + (size_t)ch * a.HW + hw);
      cm[0] = v.x; cm[1 % V] = v.y; cm[2 % V] = v.z; cm[3 % V] = v.w;
    } else {
      bx[0] = a.box[pix]; cl[0] = a.cut_label[(size_t)ch * a.HW + hw]; cm[0] = a.cut_mask[(size_t)ch * a.HW + hw];
    }
    uint8_t o[9][V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      float conf, p[MAXC];
      int pl, pl2, pl3, spl = 0;
      softmax_argmax<C>(x1[v], conf, pl, p);
      const uint8_t m = conf > a.thr;
      softmax_argmax<C>(x2[v], conf, pl2, p);
      const uint8_t m2 = conf > a.thr;
      softmax_argmax<C>(x3[v], conf, pl3, p);
      const uint8_t m3 = conf > a.thr;
      if (a.s0) softmax_argmax<C>(xs[v], conf, spl, p);
      const bool in = bx[v] != 0;
      const int plw = in ? pl3 : pl2;                       // train.py:679
      uint8_t mw = in ? m3 : m2;                            // train.py:677
      if (!(plw == pl && m)) mw = 0;                        // train.py:684-685
      o[0][v] = (uint8_t)pl; o[1][v] = m; o[2][v] = (uint8_t)plw; o[3][v] = mw;
      o[4][v] = in ? cl[v] : (uint8_t)pl;                   // pseudo_label_ul  :690
      o[5][v] = in ? cm[v] : m;                             // mask_ul          :691
      o[6][v] = in ? (uint8_t)pl : cl[v];                   // pseudo_label_lu  :693
      o[7][v] = in ? m : cm[v];                             // mask_lu          :697
      o[8][v] = (uint8_t)spl;
    }
    uint8_t* outs[9] = {a.pl, a.mask, a.pl_w, a.mask_w, a.pl_ul, a.mask_ul, a.pl_lu, a.mask_lu, a.stu_pl};
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      if (!outs[k]) continue;
      if (V == 4) *reinterpret_cast<uchar4*>(outs[k] + pix) = make_uchar4(o[k][0], o[k][1 % V], o[k][2 % V], o[k][3 % V]);
      else outs[k][pix] = o[k][0];
    }
  }
}

__device__ __forceinline__ float sigmoidf_aten(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(256) k_pseudo_label_sigmoid(PLArgs a) {
  const long long n = (long long)a.Bu * a.C * a.HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int hw = (int)(i % a.HW);
    const int b = (int)(i / ((long long)a.C * a.HW));
    const size_t cidx = (size_t)(i - (long long)b * a.C * a.HW);      // (c, hw) offset inside one sample
    const float p1 = sigmoidf_aten(a.t1[i]), p2 = sigmoidf_aten(a.t2[i]), p3 = sigmoidf_aten(a.t3[i]);
    const uint8_t pl = p1 >= 0.5f, m = (uint8_t)((p1 >= a.thr) + (p1 <= a.thr_lo));
    const uint8_t pl2 = p2 >= 0.5f, m2 = (uint8_t)((p2 >= a.thr) + (p2 <= a.thr_lo));
    const uint8_t pl3 = p3 >= 0.5f, m3 = (uint8_t)((p3 >= a.thr) + (p3 <= a.thr_lo));
    const bool in = a.box[(size_t)b * a.HW + hw] != 0;
    const int ch = a.choice[b];
    const uint8_t cl = a.cut_label[(size_t)ch * a.C * a.HW + cidx], cm = a.cut_mask[(size_t)ch * a.C * a.HW + cidx];
    const uint8_t plw = in ? pl3 : pl2;
    uint8_t mw = in ? m3 : m2;
    if (!(plw == pl && m)) mw = 0;
    if (a.pl) a.pl[i] = pl;
    if (a.mask) a.mask[i] = m;
    if (a.pl_w) a.pl_w[i] = plw;
    if (a.mask_w) a.mask_w[i] = mw;
    if (a.pl_ul) a.pl_ul[i] = in ? cl : pl;
    if (a.mask_ul) a.mask_ul[i] = in ? cm : m;
    if (a.pl_lu) a.pl_lu[i] = in ? pl : cl;
    if (a.mask_lu) a.mask_lu[i] = in ? m : cm;
    if (a.stu_pl && a.s0) a.stu_pl[i] = sigmoidf_aten(a.s0[i]) >= 0.5f;
  }
}

// dst[b][hw][c] = box ? bsrc[idx[b]][c][hw] : a[b][c][hw]
template <typename T>
__global__ void k_mix_to_nhwc(const float* __restrict__ a, const float* __restrict__ bsrc, const int* __restrict__ b_index,
                              const uint8_t* __restrict__ box, T* __restrict__ dst, int ld, int B, int C, int HW) {
  const long long n = (long long)B * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)((unsigned int)i / (unsigned int)HW), hw = (int)((unsigned int)i - (unsigned int)b * (unsigned int)HW);
    const bool in = box && box[i] != 0;
    const int sb = b_index ? b_index[b] : b;
    for (int c = 0; c < C; ++c) {
      float v = in ? bsrc[((size_t)sb * C + c) * HW + hw] : a[((size_t)b * C + c) * HW + hw];
      dst[i * ld + c] = from_f<T>(v);
    }
  }
}

// Normalize_tf (dataloaders/custom_transforms.py:650-684): float32(u8) / 127.5 - 1.0, two separately rounded operations
__device__ __forceinline__ float normalize_tf(uint8_t v) { return __fsub_rn(__fdiv_rn((float)v, 127.5f), 1.0f); }

// Same composition with either source given as the uint8 H x W x C image the reference's loaders hold BEFORE Normalize_tf +
// ToTensor (NHWC, normalised on the fly) or as a float32 NCHW tensor: dst[b][hw][c] = box ? B[idx[b]] : A[b]
template <typename T>
__global__ void k_mix_any_to_nhwc(const float* __restrict__ a_f, const uint8_t* __restrict__ a_u, const float* __restrict__ b_f, const uint8_t* __restrict__ b_u,
                                  const int* __restrict__ b_index, const uint8_t* __restrict__ box, T* __restrict__ dst, int ld, int B, int C, int HW) {
  const long long n = (long long)B * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)((unsigned int)i / (unsigned int)HW), hw = (int)((unsigned int)i - (unsigned int)b * (unsigned int)HW);
    const bool in = box && box[i] != 0;
    const int sb = b_index ? b_index[b] : b;
    for (int c = 0; c < C; ++c) {
      float v;
      if (in) v = b_u ? normalize_tf(b_u[((size_t)sb * HW + hw) * C + c]) : b_f[((size_t)sb * C + c) * HW + hw];
      else v = a_u ? normalize_tf(a_u[((size_t)b * HW + hw) * C + c]) : a_f[((size_t)b * C + c) * HW + hw];
      dst[i * ld + c] = from_f<T>(v);
    }
  }
}

// Normalize_tf + ToTensor on the device: uint8 [B,H,W,C] -> float32 [B,C,H,W] (bit-exact with the reference's numpy ops)
__global__ void k_normalize_u8_to_nchw(const uint8_t* __restrict__ src, float* __restrict__ dst, int B, int C, int HW) {
  const long long n = (long long)B * C * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int hw = (int)(i % HW);
    const long long bc = i / HW;
    const int c = (int)(bc % C), b = (int)(bc / C);
    dst[i] = normalize_tf(src[((size_t)b * HW + hw) * C + c]);
  }
}

// ------------------------------------------------------------------------------------------
// CE + Dice, softmax branch
// ------------------------------------------------------------------------------------------
template <int NACC>
__device__ __forceinline__ void block_reduce_store(float (&acc)[NACC], float* __restrict__ row) {
  __shared__ float sm[8][NACC];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NACC; ++k) {
    float v = warp_sum(acc[k]);
    if (lane == 0) sm[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < NACC) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += sm[w][threadIdx.x];
    row[threadIdx.x] = s;
  }
}

template <int C>
__global__ void __launch_bounds__(256) k_ce_dice_pass1(const float* __restrict__ logits, const uint8_t* __restrict__ target,
                                                      const uint8_t* __restrict__ mask, int B, int HW, float* __restrict__ partials) {
  constexpr int NACC = 3 * C + 1;
  float acc[NACC];
#pragma unroll
  for (int k = 0; k < NACC; ++k) acc[k] = 0.f;
  const long long n = (long long)B * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)((unsigned int)i / (unsigned int)HW), hw = (int)((unsigned int)i - (unsigned int)b * (unsigned int)HW);
    float x[MAXC];
#pragma unroll
    for (int c = 0; c < C; ++c) x[c] = logits[((size_t)b * C + c) * HW + hw];
    float mx = x[0];
#pragma unroll
    for (int c = 1; c < C; ++c) mx = fmaxf(mx, x[c]);
    float e[MAXC], s = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) { e[c] = expf(x[c] - mx); s += e[c]; }
    const float lse = logf(s);
    const int t = target[i];
    const float m = mask ? (float)mask[i] : 1.f;
    float xt = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float p = e[c] / s;
      const float tc = (t == c) ? 1.f : 0.f;
      const float mp = (c == 0 || !mask) ? 1.f : (mask[i] == 1 ? 1.f : 0.f);   // losses.py:207-213
      if (t == c) xt = x[c];
      acc[c] += p * tc * mp;
      acc[C + c] += p * p * mp;
      acc[2 * C + c] += tc * mp;
    }
    acc[3 * C] += m * (lse - (xt - mx));
  }
  block_reduce_store<NACC>(acc, partials + (size_t)blockIdx.x * NACC);
}

// one block: sum partial rows in double, emit loss + pass-2 coefficients
__global__ void k_ce_dice_finalize(const float* __restrict__ partials, int nparts, int C, double npix, float ce_w, float dice_w,
                                   const float* __restrict__ class_weight, float* __restrict__ coef, float* __restrict__ loss_out) {
  __shared__ double tot[3 * MAXC + 1];
  const int nacc = 3 * C + 1;
  // one warp per accumulator, lanes stride over the partial rows (fixed order: deterministic)
  for (int a = threadIdx.x >> 5; a < nacc; a += blockDim.x >> 5) {
    double s = 0.0;
    for (int r = threadIdx.x & 31; r < nparts; r += 32) s += (double)partials[(size_t)r * nacc + a];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) tot[a] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const double eps = 1e-10;
    double dice = 0.0;
    for (int c = 0; c < C; ++c) {
      const double I = tot[c], Z = tot[C + c], Y = tot[2 * C + c];
      const double D = Z + Y + eps;
      const double w = (class_weight ? (double)class_weight[c] : 1.0) / C;
      dice += w * (1.0 - (2.0 * I + eps) / D);
      coef[c] = (float)(dice_w * w * (-2.0 / D));
      coef[C + c] = (float)(dice_w * w * 2.0 * (2.0 * I + eps) / (D * D));
    }
    const double ce = tot[3 * C] / npix;
    coef[2 * C] = (float)(ce_w / npix);
    loss_out[0] = (float)(ce_w * ce + dice_w * dice);
    loss_out[1] = (float)ce;
    loss_out[2] = (float)dice;
  }
}

template <int C>
__global__ void __launch_bounds__(256) k_ce_dice_pass2(const float* __restrict__ logits, const uint8_t* __restrict__ target,
                                                      const uint8_t* __restrict__ mask, int B, int HW, const float* __restrict__ coef,
                                                      const float* __restrict__ upstream, float gscale, float* __restrict__ dlogits,
                                                      int accumulate) {
  const float up = gscale * (upstream ? *upstream : 1.f);
  float A[MAXC], Bc[MAXC];
#pragma unroll
  for (int c = 0; c < C; ++c) { A[c] = coef[c]; Bc[c] = coef[C + c]; }
  const float ce_scale = coef[2 * C];
  const long long n = (long long)B * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)((unsigned int)i / (unsigned int)HW), hw = (int)((unsigned int)i - (unsigned int)b * (unsigned int)HW);
    float x[MAXC];
#pragma unroll
    for (int c = 0; c < C; ++c) x[c] = logits[((size_t)b * C + c) * HW + hw];
    float mx = x[0];
#pragma unroll
    for (int c = 1; c < C; ++c) mx = fmaxf(mx, x[c]);
    float p[MAXC], s = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) { p[c] = expf(x[c] - mx); s += p[c]; }
    const int t = target[i];
    const float m = mask ? (float)mask[i] : 1.f;
    const float mp1 = (!mask || mask[i] == 1) ? 1.f : 0.f;
    float g[MAXC], dot = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      p[c] = p[c] / s;
      const float tc = (t == c) ? 1.f : 0.f;
      const float mp = (c == 0) ? 1.f : mp1;
      g[c] = mp * (A[c] * tc + Bc[c] * p[c]);
      dot += g[c] * p[c];
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float tc = (t == c) ? 1.f : 0.f;
      const float d = up * (ce_scale * m * (p[c] - tc) + p[c] * (g[c] - dot));
      const size_t o = ((size_t)b * C + c) * HW + hw;
      dlogits[o] = accumulate ? dlogits[o] + d : d;
    }
  }
}

// ------------------------------------------------------------------------------------------
// BCE + global Dice, sigmoid/multi (fundus) branch
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_bce_dice_pass1(const float* __restrict__ logits, const uint8_t* __restrict__ target,
                                                       const uint8_t* __restrict__ mask, long long n, float* __restrict__ partials) {
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float x = logits[i];
    const float t = (float)target[i];
    const float m = mask ? (float)mask[i] : 1.f;
    const float s = sigmoidf_aten(x);
    // ATen binary_cross_entropy_with_logits: (1-t)*x + max(-x,0) + log(exp(-max) + exp(-x-max))
    const float mxv = fmaxf(-x, 0.f);
    const float bce = (1.f - t) * x + mxv + logf(expf(-mxv) + expf(-x - mxv));
    acc[0] += s * t * m;
    acc[1] += s * s * m;
    acc[2] += t * t * m;
    acc[3] += bce * m;
  }
  block_reduce_store<4>(acc, partials + (size_t)blockIdx.x * 4);
}
__global__ void k_bce_dice_finalize(const float* __restrict__ partials, int nparts, double n, float ce_w, float dice_w,
                                    float* __restrict__ coef, float* __restrict__ loss_out) {
  __shared__ double tot[4];
  for (int a = threadIdx.x >> 5; a < 4; a += blockDim.x >> 5) {
    double s = 0.0;
    for (int r = threadIdx.x & 31; r < nparts; r += 32) s += (double)partials[(size_t)r * 4 + a];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) tot[a] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const double eps = 1e-10, I = tot[0], Z = tot[1], Y = tot[2], D = Z + Y + eps;
    const double dice = 1.0 - (2.0 * I + eps) / D, ce = tot[3] / n;
    coef[0] = (float)(dice_w * (-2.0 / D));
    coef[1] = (float)(dice_w * 2.0 * (2.0 * I + eps) / (D * D));
    coef[2] = (float)(ce_w / n);
    loss_out[0] = (float)(ce_w * ce + dice_w * dice);
    loss_out[1] = (float)ce;
    loss_out[2] = (float)dice;
  }
}
__global__ void __launch_bounds__(256) k_bce_dice_pass2(const float* __restrict__ logits, const uint8_t* __restrict__ target,
                                                       const uint8_t* __restrict__ mask, long long n, const float* __restrict__ coef,
                                                       const float* __restrict__ upstream, float gscale, float* __restrict__ dlogits,
                                                       int accumulate) {
  const float up = gscale * (upstream ? *upstream : 1.f);
  const float A = coef[0], Bc = coef[1], ce_scale = coef[2];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float x = logits[i];
    const float t = (float)target[i];
    const float m = mask ? (float)mask[i] : 1.f;
    const float s = sigmoidf_aten(x);
    const float d = up * m * (ce_scale * (s - t) + (A * t + Bc * s) * s * (1.f - s));
    dlogits[i] = accumulate ? dlogits[i] + d : d;
  }
}

__global__ void k_reduce_rows(const float* __restrict__ rows, int nrows, int ncols, float* __restrict__ out) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= ncols) return;
  double s = 0.0;
  for (int r = lane; r < nrows; r += 32) s += (double)rows[(size_t)r * ncols + c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[c] = (float)s;
}

static inline int loss_grid(long long n) {
  long long b = (n + 255) / 256;
  if (b > USTRUN_MAX_PARTS) b = USTRUN_MAX_PARTS;
  if (b > 148 * 8) b = 148 * 8;
  return (int)(b < 1 ? 1 : b);
}

// ------------------------------------------------------------------------------------------
// hardness of the unlabelled samples (train.py:705-718, utils/metrics.py:114-231)
// ------------------------------------------------------------------------------------------
// Per sample and part: S = #student, G = #teacher, I = #both (booleans), exactly what dice_coefficient_numpy counts
// after its .cpu() copies.  mode 0: one part, label != 0 (dice_coeff: prostate / BUSI); mode 1: planes [B][2][HW], one
// part per channel (dice_coeff_2label: fundus); mode 2: parts = classes 1..3 of one label plane (dice_coeff_3label: M&Ms).
__global__ void __launch_bounds__(256)
k_hardness_counts(const uint8_t* __restrict__ stu, const uint8_t* __restrict__ tea, int HW, int mode, int parts, unsigned int* __restrict__ counts) {
  const int b = blockIdx.x, part = blockIdx.y;
  const long long plane = mode == 1 ? ((long long)b * parts + part) * HW : (long long)b * HW;
  const uint8_t* ps = stu + plane;
  const uint8_t* pt = tea + plane;
  const int cls = part + 1;
  unsigned int S = 0, G = 0, I = 0;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) {
    const bool s_ = mode == 2 ? ps[i] == cls : ps[i] != 0;
    const bool g_ = mode == 2 ? pt[i] == cls : pt[i] != 0;
    S += s_; G += g_; I += (s_ && g_);
  }
  __shared__ unsigned int sm[3][8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    S += __shfl_xor_sync(0xffffffffu, S, o); G += __shfl_xor_sync(0xffffffffu, G, o); I += __shfl_xor_sync(0xffffffffu, I, o);
  }
  if ((threadIdx.x & 31) == 0) { sm[0][threadIdx.x >> 5] = S; sm[1][threadIdx.x >> 5] = G; sm[2][threadIdx.x >> 5] = I; }
  __syncthreads();
  if (threadIdx.x < 3) {
    unsigned int t = 0;
    for (int w = 0; w < 8; ++w) t += sm[threadIdx.x][w];
    counts[((size_t)b * parts + part) * 3 + threadIdx.x] = t;
  }
}
// dice = (2I + 1) / (1.001 + S + G), 0 when both are empty (metrics.py:139-143); hardness = 1 - mean over parts
// (train.py:706-710), all ones in the first epoch (:711-713); lq_idx = first maximum (:714-718).  float64 like Python.
__global__ void k_hardness_finalize(const unsigned int* __restrict__ counts, int B, int parts, int first_epoch, double* __restrict__ hardness,
                                    double* __restrict__ dice_out, int* __restrict__ lq_idx) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int best = 0;
  double maxv = -1.0;
  for (int b = 0; b < B; ++b) {
    double tmp = 0.0;
    for (int p = 0; p < parts; ++p) {
      const unsigned int* c = counts + ((size_t)b * parts + p) * 3;
      const double S = (double)c[0], G = (double)c[1], I = (double)c[2];
      const double d = (c[0] == 0 && c[1] == 0) ? 0.0 : (2.0 * I + 1.0) / (1.001 + S + G);
      if (dice_out) dice_out[(size_t)p * B + b] = d;
      tmp = p == 0 ? d : tmp + d;
    }
    double h = 1.0 - tmp / (double)parts;
    if (first_epoch) h = 1.0;
    hardness[b] = h;
    if (h > maxv) { maxv = h; best = b; }
  }
  *lq_idx = best;
}

// ------------------------------------------------------------------------------------------
// evaluation helpers: label encodings, predictions, per-part segmentation metrics
// ------------------------------------------------------------------------------------------
// train.py:590-608 / 281-288 and train_mnms.py:549-556.  mode 0: y == 0 (prostate); 1: y == 255 (BUSI);
// 2: two planes {y == 0, y <= 128} (fundus: cup, disc); 3: y is [B,H,W,3], label = 1/2/3 where channel 0/1/2 == 255,
// later channels override (M&Ms).  y: the float32 label image the data loader yields.
__global__ void __launch_bounds__(256)
k_encode_labels(const float* __restrict__ y, int mode, int B, int HW, uint8_t* __restrict__ out) {
  const long long n = (long long)B * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (mode == 0) out[i] = y[i] == 0.f;
    else if (mode == 1) out[i] = y[i] == 255.f;
    else if (mode == 2) {
      const long long b = i / HW, p = i - b * HW;
      out[(b * 2) * HW + p] = y[i] == 0.f;
      out[(b * 2 + 1) * HW + p] = y[i] <= 128.f;
    } else {
      uint8_t l = y[i * 3] == 255.f ? 1 : 0;
      if (y[i * 3 + 1] == 255.f) l = 2;
      if (y[i * 3 + 2] == 255.f) l = 3;
      out[i] = l;
    }
  }
}
// softmax branch: torch.max(torch.softmax(output, 1), 1)[1] (first maximum of the PROBABILITIES)
template <int C>
__global__ void __launch_bounds__(256) k_predict_softmax(const float* __restrict__ logits, int B, int HW, uint8_t* __restrict__ pred) {
  const long long n = (long long)B * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / HW, p = i - b * HW;
    float x[MAXC], pr[MAXC], conf;
    int arg;
#pragma unroll
    for (int c = 0; c < C; ++c) x[c] = logits[(b * C + c) * HW + p];
    softmax_argmax<C>(x, conf, arg, pr);
    pred[i] = (uint8_t)arg;
  }
}
// sigmoid branch: torch.sigmoid(output).ge(0.5)
__global__ void __launch_bounds__(256) k_predict_sigmoid(const float* __restrict__ logits, long long n, uint8_t* __restrict__ pred) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    pred[i] = sigmoidf_aten(logits[i]) >= 0.5f;
}
// batch means per part of: Dice as utils/metrics.py:114-146, and medpy.metric.binary dc = 2I/(S+G) (0 when S+G == 0)
// and jc = I/(S+G-I) (medpy raises ZeroDivisionError for an empty union; 0 here) -- train.py:304-320
__global__ void k_seg_metrics_finalize(const unsigned int* __restrict__ counts, int B, int parts, double* __restrict__ out) {
  const int p = threadIdx.x;
  if (p >= parts) return;
  double dice = 0.0, dc = 0.0, jc = 0.0;
  for (int b = 0; b < B; ++b) {
    const unsigned int* c = counts + ((size_t)b * parts + p) * 3;
    const double S = (double)c[0], G = (double)c[1], I = (double)c[2];
    dice += (c[0] == 0 && c[1] == 0) ? 0.0 : (2.0 * I + 1.0) / (1.001 + S + G);
    dc += (S + G) > 0.0 ? 2.0 * I / (S + G) : 0.0;
    jc += (S + G - I) > 0.0 ? I / (S + G - I) : 0.0;
  }
  out[p] = dice / (double)B;               // sum(all_dice) / len(all_dice): sequential sum, then one division
  out[parts + p] = dc / (double)B;
  out[2 * parts + p] = jc / (double)B;
}

}  // namespace ustrun

using namespace ustrun;

#define DISPATCH_C(C, ...)                                                     \
  switch (C) {                                                                 \
    case 2: { constexpr int kC = 2; __VA_ARGS__; } break;                      \
    case 3: { constexpr int kC = 3; __VA_ARGS__; } break;                      \
    case 4: { constexpr int kC = 4; __VA_ARGS__; } break;                      \
    case 5: { constexpr int kC = 5; __VA_ARGS__; } break;                      \
    case 6: { constexpr int kC = 6; __VA_ARGS__; } break;                      \
    case 7: { constexpr int kC = 7; __VA_ARGS__; } break;                      \
    case 8: { constexpr int kC = 8; __VA_ARGS__; } break;                      \
    default: set_error("n_classes %d unsupported (2..8)", (int)(C)); return USTRUN_ERR_ARG; \
  }

extern "C" {

int ustrun_pseudo_label_softmax(const float* t1, const float* t2, const float* t3, const float* s0, const uint8_t* box,
                                const uint8_t* cut_label, const uint8_t* cut_mask, const int* choice, float threshold, int Bu, int C, int H,
                                int W, uint8_t* pl, uint8_t* mask, uint8_t* pl_w, uint8_t* mask_w, uint8_t* pl_ul, uint8_t* mask_ul,
                                uint8_t* pl_lu, uint8_t* mask_lu, uint8_t* stu_pl, void* stream) {
  USTRUN_REQUIRE(t1 && t2 && t3 && box && cut_label && cut_mask && choice && Bu > 0 && H > 0 && W > 0, "pseudo_label_softmax: null/empty arg");
  PLArgs a{t1, t2, t3, s0, box, cut_label, cut_mask, choice, threshold, 0.f, Bu, C, H * W, pl, mask, pl_w, mask_w, pl_ul, mask_ul, pl_lu, mask_lu, stu_pl};
  const long long n = (long long)Bu * H * W;
  cudaStream_t st = (cudaStream_t)stream;
  if ((H * W) % 4 == 0) {
    int grid = (int)((n / 4 + 255) / 256);
    if (grid > 148 * 8) grid = 148 * 8;
    DISPATCH_C(C, (k_pseudo_label_softmax<kC, 4><<<grid, 256, 0, st>>>(a)));
  } else {
    int grid = (int)((n + 255) / 256);
    if (grid > 148 * 8) grid = 148 * 8;
    DISPATCH_C(C, (k_pseudo_label_softmax<kC, 1><<<grid, 256, 0, st>>>(a)));
  }
  return check_launch("pseudo_label_softmax");
}

int ustrun_pseudo_label_sigmoid(const float* t1, const float* t2, const float* t3, const float* s0, const uint8_t* box,
                                const uint8_t* cut_label, const uint8_t* cut_mask, const int* choice, float thr_hi, float thr_lo, int Bu, int C,
                                int H, int W, uint8_t* pl, uint8_t* mask, uint8_t* pl_w, uint8_t* mask_w, uint8_t* pl_ul, uint8_t* mask_ul,
                                uint8_t* pl_lu, uint8_t* mask_lu, uint8_t* stu_pl, void* stream) {
  USTRUN_REQUIRE(t1 && t2 && t3 && box && cut_label && cut_mask && choice && Bu > 0 && C > 0 && H > 0 && W > 0, "pseudo_label_sigmoid: null/empty arg");
  PLArgs a{t1, t2, t3, s0, box, cut_label, cut_mask, choice, thr_hi, thr_lo, Bu, C, H * W, pl, mask, pl_w, mask_w, pl_ul, mask_ul, pl_lu, mask_lu, stu_pl};
  const long long n = (long long)Bu * C * H * W;
  int grid = (int)((n + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  k_pseudo_label_sigmoid<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  return check_launch("pseudo_label_sigmoid");
}

int ustrun_mix_to_nhwc(const float* a, const float* b, const int* b_index, const uint8_t* box, void* dst, int ld_dst, int dtype, int B, int C,
                       int H, int W, void* stream) {
  USTRUN_REQUIRE(a && dst && B > 0 && C > 0 && ld_dst >= C && (!box || b), "mix_to_nhwc: bad args");
  const long long n = (long long)B * H * W;
  int grid = (int)((n + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  if (dtype == USTRUN_F32) k_mix_to_nhwc<float><<<grid, 256, 0, (cudaStream_t)stream>>>(a, b, b_index, box, (float*)dst, ld_dst, B, C, H * W);
  else if (dtype == USTRUN_BF16) k_mix_to_nhwc<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(a, b, b_index, box, (__nv_bfloat16*)dst, ld_dst, B, C, H * W);
  else { set_error("bad dtype"); return USTRUN_ERR_ARG; }
  return check_launch("mix_to_nhwc");
}

int ustrun_mix_any_to_nhwc(const float* a_f32, const uint8_t* a_u8, const float* b_f32, const uint8_t* b_u8, const int* b_index, const uint8_t* box,
                           void* dst, int ld_dst, int dtype, int B, int C, int H, int W, void* stream) {
  USTRUN_REQUIRE((a_f32 != nullptr) != (a_u8 != nullptr), "mix_any_to_nhwc: exactly one of a_f32 / a_u8");
  USTRUN_REQUIRE(dst && B > 0 && C > 0 && ld_dst >= C && !(b_f32 && b_u8) && (!box || b_f32 || b_u8), "mix_any_to_nhwc: bad args");
  const long long n = (long long)B * H * W;
  int grid = (int)((n + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  if (dtype == USTRUN_F32) k_mix_any_to_nhwc<float><<<grid, 256, 0, (cudaStream_t)stream>>>(a_f32, a_u8, b_f32, b_u8, b_index, box, (float*)dst, ld_dst, B, C, H * W);
  else if (dtype == USTRUN_BF16)
    k_mix_any_to_nhwc<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(a_f32, a_u8, b_f32, b_u8, b_index, box, (__nv_bfloat16*)dst, ld_dst, B, C, H * W);
  else { set_error("bad dtype"); return USTRUN_ERR_ARG; }
  return check_launch("mix_any_to_nhwc");
}

int ustrun_normalize_u8_to_nchw(const uint8_t* src, float* dst, int B, int C, int H, int W, void* stream) {
  USTRUN_REQUIRE(src && dst && B > 0 && C > 0 && H > 0 && W > 0, "normalize_u8_to_nchw: bad args");
  const long long n = (long long)B * C * H * W;
  int grid = (int)((n + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  k_normalize_u8_to_nchw<<<grid, 256, 0, (cudaStream_t)stream>>>(src, dst, B, C, H * W);
  return check_launch("normalize_u8_to_nchw");
}

int ustrun_reduce_rows(const float* rows, int nrows, int ncols, float* out, void* stream) {
  USTRUN_REQUIRE(rows && out && nrows > 0 && ncols > 0, "reduce_rows: bad args");
  k_reduce_rows<<<ceil_div(ncols, 8), 256, 0, (cudaStream_t)stream>>>(rows, nrows, ncols, out);
  return check_launch("reduce_rows");
}

int ustrun_ce_dice_softmax_partials(const float* logits, const uint8_t* target, const uint8_t* mask, int B, int C, int H, int W,
                                    float* workspace, int* nparts_host, void* stream) {
  USTRUN_REQUIRE(logits && target && workspace && nparts_host && B > 0 && H > 0 && W > 0, "ce_dice_softmax_partials: null/empty arg");
  const long long n = (long long)B * H * W;
  const int grid = loss_grid(n);
  *nparts_host = grid;
  DISPATCH_C(C, (k_ce_dice_pass1<kC><<<grid, 256, 0, (cudaStream_t)stream>>>(logits, target, mask, B, H * W, workspace)));
  return check_launch("ce_dice_softmax_partials");
}
int ustrun_ce_dice_softmax_finalize(const float* workspace, int nparts, int C, double npix_total, float ce_w, float dice_w,
                                    const float* class_weight, float* coef, float* loss_out, void* stream) {
  USTRUN_REQUIRE(workspace && nparts > 0 && C >= 2 && C <= 8 && npix_total > 0 && coef && loss_out, "ce_dice_softmax_finalize: bad args");
  k_ce_dice_finalize<<<1, 256, 0, (cudaStream_t)stream>>>(workspace, nparts, C, npix_total, ce_w, dice_w, class_weight, coef, loss_out);
  return check_launch("ce_dice_softmax_finalize");
}
int ustrun_ce_dice_softmax_fwd(const float* logits, const uint8_t* target, const uint8_t* mask, int B, int C, int H, int W, float ce_w,
                               float dice_w, const float* class_weight, float* workspace, float* coef, float* loss_out, void* stream) {
  USTRUN_REQUIRE(coef && loss_out, "ce_dice_softmax_fwd: null arg");
  int nparts = 0;
  int rc = ustrun_ce_dice_softmax_partials(logits, target, mask, B, C, H, W, workspace, &nparts, stream);
  if (rc) return rc;
  return ustrun_ce_dice_softmax_finalize(workspace, nparts, C, (double)B * H * W, ce_w, dice_w, class_weight, coef, loss_out, stream);
}
int ustrun_ce_dice_softmax_bwd(const float* logits, const uint8_t* target, const uint8_t* mask, int B, int C, int H, int W, const float* coef,
                               const float* upstream, float gscale, float* dlogits, int accumulate, void* stream) {
  USTRUN_REQUIRE(logits && target && coef && dlogits && B > 0 && H > 0 && W > 0, "ce_dice_softmax_bwd: null/empty arg");
  const long long n = (long long)B * H * W;
  int grid = (int)((n + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  DISPATCH_C(C, (k_ce_dice_pass2<kC><<<grid, 256, 0, (cudaStream_t)stream>>>(logits, target, mask, B, H * W, coef, upstream, gscale, dlogits, accumulate)));
  return check_launch("ce_dice_softmax_bwd");
}
int ustrun_bce_dice_sigmoid_partials(const float* logits, const uint8_t* target, const uint8_t* mask, int B, int C, int H, int W,
                                     float* workspace, int* nparts_host, void* stream) {
  USTRUN_REQUIRE(logits && target && workspace && nparts_host && B > 0 && C > 0 && H > 0 && W > 0, "bce_dice_sigmoid_partials: null/empty arg");
  const long long n = (long long)B * C * H * W;
  const int grid = loss_grid(n);
  *nparts_host = grid;
  k_bce_dice_pass1<<<grid, 256, 0, (cudaStream_t)stream>>>(logits, target, mask, n, workspace);
  return check_launch("bce_dice_sigmoid_partials");
}
int ustrun_bce_dice_sigmoid_finalize(const float* workspace, int nparts, double nelem_total, float ce_w, float dice_w, float* coef,
                                     float* loss_out, void* stream) {
  USTRUN_REQUIRE(workspace && nparts > 0 && nelem_total > 0 && coef && loss_out, "bce_dice_sigmoid_finalize: bad args");
  k_bce_dice_finalize<<<1, 128, 0, (cudaStream_t)stream>>>(workspace, nparts, nelem_total, ce_w, dice_w, coef, loss_out);
  return check_launch("bce_dice_sigmoid_finalize");
}
int ustrun_bce_dice_sigmoid_fwd(const float* logits, const uint8_t* target, const uint8_t* mask, int B, int C, int H, int W, float ce_w,
                                float dice_w, float* workspace, float* coef, float* loss_out, void* stream) {
  USTRUN_REQUIRE(coef && loss_out, "bce_dice_sigmoid_fwd: null arg");
  int nparts = 0;
  int rc = ustrun_bce_dice_sigmoid_partials(logits, target, mask, B, C, H, W, workspace, &nparts, stream);
  if (rc) return rc;
  return ustrun_bce_dice_sigmoid_finalize(workspace, nparts, (double)B * C * H * W, ce_w, dice_w, coef, loss_out, stream);
}
int ustrun_bce_dice_sigmoid_bwd(const float* logits, const uint8_t* target, const uint8_t* mask, int B, int C, int H, int W, const float* coef,
                                const float* upstream, float gscale, float* dlogits, int accumulate, void* stream) {
  USTRUN_REQUIRE(logits && target && coef && dlogits && B > 0 && C > 0 && H > 0 && W > 0, "bce_dice_sigmoid_bwd: null/empty arg");
  const long long n = (long long)B * C * H * W;
  int grid = (int)((n + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  k_bce_dice_pass2<<<grid, 256, 0, (cudaStream_t)stream>>>(logits, target, mask, n, coef, upstream, gscale, dlogits, accumulate);
  return check_launch("bce_dice_sigmoid_bwd");
}

int ustrun_hardness(const uint8_t* stu_pl, const uint8_t* tea_pl, int B, int H, int W, int mode, int first_epoch, unsigned int* workspace,
                    double* hardness, double* dice, int* lq_idx, void* stream) {
  USTRUN_REQUIRE(stu_pl && tea_pl && workspace && hardness && lq_idx && B > 0 && H > 0 && W > 0 && mode >= 0 && mode <= 2, "hardness: bad args");
  const int parts = mode == 0 ? 1 : (mode == 1 ? 2 : 3);
  k_hardness_counts<<<dim3(B, parts), 256, 0, (cudaStream_t)stream>>>(stu_pl, tea_pl, H * W, mode, parts, workspace);
  k_hardness_finalize<<<1, 32, 0, (cudaStream_t)stream>>>(workspace, B, parts, first_epoch, hardness, dice, lq_idx);
  return check_launch("hardness");
}

int ustrun_encode_labels(const float* y, int mode, int B, int H, int W, unsigned char* out, void* stream) {
  USTRUN_REQUIRE(y && out && B > 0 && H > 0 && W > 0 && mode >= 0 && mode <= 3, "encode_labels: bad args");
  const long long n = (long long)B * H * W;
  k_encode_labels<<<loss_grid(n), 256, 0, (cudaStream_t)stream>>>(y, mode, B, H * W, out);
  return check_launch("encode_labels");
}
int ustrun_predict(const float* logits, int sigmoid, int B, int C, int H, int W, unsigned char* pred, void* stream) {
  USTRUN_REQUIRE(logits && pred && B > 0 && C > 0 && H > 0 && W > 0, "predict: bad args");
  if (sigmoid) {
    const long long n = (long long)B * C * H * W;
    k_predict_sigmoid<<<loss_grid(n), 256, 0, (cudaStream_t)stream>>>(logits, n, pred);
  } else {
    USTRUN_REQUIRE(C >= 2 && C <= 8, "predict: softmax branch supports 2..8 classes");
    const long long n = (long long)B * H * W;
    DISPATCH_C(C, (k_predict_softmax<kC><<<loss_grid(n), 256, 0, (cudaStream_t)stream>>>(logits, B, H * W, pred)));
  }
  return check_launch("predict");
}
int ustrun_seg_metrics(const uint8_t* pred, const uint8_t* target, int B, int H, int W, int mode, unsigned int* workspace, double* out, void* stream) {
  USTRUN_REQUIRE(pred && target && workspace && out && B > 0 && H > 0 && W > 0 && mode >= 0 && mode <= 2, "seg_metrics: bad args");
  const int parts = mode == 0 ? 1 : (mode == 1 ? 2 : 3);
  k_hardness_counts<<<dim3(B, parts), 256, 0, (cudaStream_t)stream>>>(pred, target, H * W, mode, parts, workspace);
  k_seg_metrics_finalize<<<1, 32, 0, (cudaStream_t)stream>>>(workspace, B, parts, out);
  return check_launch("seg_metrics");
}

}  // extern "C"
