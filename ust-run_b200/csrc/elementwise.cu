// Memory-bound kernels of the UNet forward/backward: layout converters, weight packing, BatchNorm
// finalize/apply/backward, max-pool backward, bilinear x2 up-sampling, per-channel sums.
// All are HBM-bound: 128-bit vector accesses along the NHWC channel axis, warp/block reductions,
// two-stage (partials -> finalize) deterministic sums, no atomics.
#include <stdarg.h>
#include <stdio.h>

#include <stdlib.h>

#include <stdlib.h>

#include "common.cuh"

namespace ustrun {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}
const char* last_error() { return g_err; }

#define DISPATCH_DTYPE(dtype, ...)                                   \
  if ((dtype) == USTRUN_F32) {                                       \
    using T = float;                                                 \
    __VA_ARGS__;                                                     \
  } else if ((dtype) == USTRUN_BF16) {                               \
    using T = __nv_bfloat16;                                         \
    __VA_ARGS__;                                                     \
  } else {                                                           \
    set_error("bad dtype %d", (int)(dtype));                         \
    return USTRUN_ERR_ARG;                                           \
  }

// The tcgen05 conv CTAs run with the maximum shared-memory carve-out (228 KB).  A kernel launched with another carve-out
// cannot be resident on an SM until that SM has drained, so the HBM-bound BatchNorm / pooling kernels -- whose blocks are
// meant to run NEXT TO a conv CTA of another lane (multi-lane step) -- ask for the same configuration (they stream and do not
// need the L1).  Measured on B200 (tools/coreside_probe.py, profiles/r02_coreside_probe.txt): OFF by default -- with the hint
// the streaming kernels themselves slow down (bn_act 52 -> 59 us, bn_bwd_apply 77 -> 101 us for 151 MB) and the conv + BN
// pair still takes ~0.85 x the sum of the two, i.e. the blocks do not run side by side.  USTRUN_BN_CARVEOUT=1 enables it.
static int carveout_mode() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("USTRUN_BN_CARVEOUT");
    v = e ? atoi(e) : 0;
  }
  return v;
}
static int prefer_max_shared_impl(const void* kernel) {
  static const void* seen[64];
  static int nseen = 0;
  for (int i = 0; i < nseen; ++i)
    if (seen[i] == kernel) return 0;
  if (nseen < 64) seen[nseen++] = kernel;
  if (carveout_mode() > 0) cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  return 0;
}
#define PREFER_MAX_SHARED(kernel) prefer_max_shared_impl((const void*)(kernel))

static inline int grid_for(long long items, int threads, int cap = 148 * 16) {
  long long b = (items + threads - 1) / threads;
  if (b < 1) b = 1;
  return (int)(b > cap ? cap : b);
}

// ------------------------------------------------------------------------------------------
// layout
// ------------------------------------------------------------------------------------------
// src [B][C][HW] fp32 -> dst [B][HW][ld] T, 32x32 smem transpose tiles
template <typename T>
__global__ void k_nchw_to_nhwc(const float* __restrict__ src, T* __restrict__ dst, int C, int HW, int ld) {
  __shared__ float tile[32][33];
  int b = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i, p = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && p < HW) ? src[((size_t)b * C + c) * HW + p] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int p = p0 + i, c = c0 + threadIdx.x;
    if (p < HW && c < C) dst[((size_t)b * HW + p) * ld + c] = from_f<T>(tile[threadIdx.x][i]);
  }
}
// C <= 8 (logit gradients, 1..4-channel images): one thread per pixel, plane reads coalesced across the warp,
// the C outputs of a pixel written back to back (the 32x32 tile kernel above would use 2 of its 32 channel lanes)
template <typename T, int C>
__global__ void __launch_bounds__(256)
k_nchw_to_nhwc_small(const float* __restrict__ src, T* __restrict__ dst, long long HW, int ld, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / HW, p = i - b * HW;
    float v[C];
#pragma unroll
    for (int c = 0; c < C; ++c) v[c] = src[(b * C + c) * HW + p];
    T* d = dst + i * ld;
#pragma unroll
    for (int c = 0; c < C; ++c) d[c] = from_f<T>(v[c]);
  }
}

template <typename T>
__global__ void k_nhwc_to_nchw(const T* __restrict__ src, float* __restrict__ dst, int C, int HW, int ld) {
  __shared__ float tile[32][33];
  int b = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int p = p0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && p < HW) ? to_f(src[((size_t)b * HW + p) * ld + c]) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i, p = p0 + threadIdx.x;
    if (p < HW && c < C) dst[((size_t)b * C + c) * HW + p] = tile[threadIdx.x][i];
  }
}

// OIHW fp32 -> wf[co][tap][ci] and wd[ci][taps-1-tap][co] (flipped, transposed: the dgrad operand).
// One block = a 32(co) x 32(ci) x taps tile staged in shared memory so that the fp32 reads (runs of
// 32*taps floats) and both packed writes (32 consecutive ci / co) are coalesced.
template <typename T>
__global__ void __launch_bounds__(256)
k_pack_conv(const float* __restrict__ w, T* __restrict__ wf, T* __restrict__ wd, int Cout, int Cin, int taps) {
  extern __shared__ float tile[];                      // [32 co][32 ci][taps] (+1 pad per co row)
  const int co0 = blockIdx.x * 32, ci0 = blockIdx.y * 32;
  const int nco = min(32, Cout - co0), nci = min(32, Cin - ci0);
  const int row = 32 * taps + 1;
  for (int i = threadIdx.x; i < nco * nci * taps; i += 256) {
    const int co = i / (nci * taps), r = i - co * (nci * taps);          // r = ci*taps + t, contiguous in w
    tile[co * row + r] = w[((size_t)(co0 + co) * Cin + ci0) * taps + r];
  }
  __syncthreads();
  if (wf)
    for (int i = threadIdx.x; i < nco * taps * nci; i += 256) {
      const int ci = i % nci, t = (i / nci) % taps, co = i / (nci * taps);
      wf[((size_t)(co0 + co) * taps + t) * Cin + ci0 + ci] = from_f<T>(tile[co * row + ci * taps + t]);
    }
  if (wd)
    for (int i = threadIdx.x; i < nci * taps * nco; i += 256) {
      const int co = i % nco, t = (i / nco) % taps, ci = i / (nco * taps);
      wd[((size_t)(ci0 + ci) * taps + (taps - 1 - t)) * Cout + co0 + co] = from_f<T>(tile[co * row + ci * taps + t]);
    }
}
// Multi-tensor packing: ONE launch re-packs every conv / transposed-conv weight of a model after the optimiser step (46
// launches per step for UNet-A + teacher, ~200 for UNet-B, each a few microseconds of work behind a launch).  Block i works
// on tile blk_tile[i] of table entry blk_entry[i]: conv entries use the 32 x 32 x taps shared-memory tile of k_pack_conv,
// transposed-conv entries a 4096-element run of the flat weight.
template <typename T>
__global__ void __launch_bounds__(256)
k_pack_multi(const ustrun_pack_t* __restrict__ table, const int* __restrict__ blk_entry, const int* __restrict__ blk_tile) {
  extern __shared__ float tile[];
  const ustrun_pack_t e = table[blk_entry[blockIdx.x]];
  const int tl = blk_tile[blockIdx.x];
  T* wf = (T*)e.wf;
  T* wd = (T*)e.wd;
  if (e.transposed) {
    const long long n = (long long)e.Cin * e.Cout * 4, i0 = (long long)tl * 4096;
    for (long long i = i0 + threadIdx.x; i < n && i < i0 + 4096; i += 256) {
      const int ij = (int)(i % 4), co = (int)((i / 4) % e.Cout), ci = (int)(i / (4LL * e.Cout));
      const float v = e.w[i];
      if (wf) wf[((size_t)ij * e.Cout + co) * e.Cin + ci] = from_f<T>(v);
      if (wd) wd[((size_t)ci * 4 + ij) * e.Cout + co] = from_f<T>(v);
    }
    return;
  }
  const int Cout = e.Cout, Cin = e.Cin, taps = e.taps;
  const int ci_tiles = (Cin + 31) >> 5;
  const int co0 = (tl / ci_tiles) * 32, ci0 = (tl % ci_tiles) * 32;
  const int nco = min(32, Cout - co0), nci = min(32, Cin - ci0);
  const int row = 32 * taps + 1;
  for (int i = threadIdx.x; i < nco * nci * taps; i += 256) {
    const int co = i / (nci * taps), r = i - co * (nci * taps);
    tile[co * row + r] = e.w[((size_t)(co0 + co) * Cin + ci0) * taps + r];
  }
  __syncthreads();
  if (wf)
    for (int i = threadIdx.x; i < nco * taps * nci; i += 256) {
      const int ci = i % nci, t = (i / nci) % taps, co = i / (nci * taps);
      wf[((size_t)(co0 + co) * taps + t) * Cin + ci0 + ci] = from_f<T>(tile[co * row + ci * taps + t]);
    }
  if (wd)
    for (int i = threadIdx.x; i < nci * taps * nco; i += 256) {
      const int co = i % nco, t = (i / nco) % taps, ci = i / (nco * taps);
      wd[((size_t)(ci0 + ci) * taps + (taps - 1 - t)) * Cout + co0 + co] = from_f<T>(tile[co * row + ci * taps + t]);
    }
}
// ConvTranspose2d weight [Cin][Cout][2][2]
template <typename T>
__global__ void k_pack_convT(const float* __restrict__ w, T* __restrict__ wf, T* __restrict__ wd, int Cin, int Cout) {
  long long n = (long long)Cin * Cout * 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int ij = (int)(i % 4);
    int co = (int)((i / 4) % Cout);
    int ci = (int)(i / (4LL * Cout));
    float v = w[i];
    if (wf) wf[((size_t)ij * Cout + co) * Cin + ci] = from_f<T>(v);
    if (wd) wd[((size_t)ci * 4 + ij) * Cout + co] = from_f<T>(v);
  }
}

// ------------------------------------------------------------------------------------------
// BatchNorm
// ------------------------------------------------------------------------------------------
__global__ void k_bn_reduce_partials(const float* __restrict__ partials, int nparts, int C, float* __restrict__ sums) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * C) return;
  double s = 0.0;
  for (int r = 0; r < nparts; ++r) s += (double)partials[(size_t)r * 2 * C + i];
  sums[i] = (float)s;
}

// Deferred running-statistics update (multi-lane step): every train-mode forward of a step wrote its batch statistics
// (mean incl. the conv bias, unbiased variance) into slot s of stats[slot][2][C]; this kernel applies the slots named by
// `mask` in increasing slot order = the order of the reference's forwards (train.py:643-647,668,699-702,740), whatever
// order the forwards actually executed in.  One launch for all BatchNorm layers of a model.
__global__ void k_bn_running_update(const ustrun_bn_update_t* __restrict__ table) {
  const ustrun_bn_update_t t = table[blockIdx.x];
  for (int c = blockIdx.y * blockDim.x + threadIdx.x; c < t.C; c += gridDim.y * blockDim.x) {
    float rm = t.running_mean[c], rv = t.running_var[c];
    for (int s = 0; s < 32; ++s)
      if (t.mask >> s & 1u) {
        rm = bn_running(rm, t.stats[(size_t)s * 2 * t.C + c], t.momentum);
        rv = bn_running(rv, t.stats[(size_t)s * 2 * t.C + t.C + c], t.momentum);
      }
    t.running_mean[c] = rm;
    t.running_var[c] = rv;
  }
  if (blockIdx.y == 0 && threadIdx.x == 0 && t.nbt) *t.nbt += __popc(t.mask);
}

__global__ void k_bn_finalize(const float* __restrict__ partials, int nparts, int C, double count,
                              const float* __restrict__ gamma, const float* __restrict__ beta,
                              const float* __restrict__ conv_bias, float* running_mean, float* running_var,
                              long long* nbt, float momentum, float eps, int training, float* scale, float* shift,
                              float* mean_out, float* rstd_out, float* stat_out) {
  // one warp per channel: lanes stride over the partial rows, double accumulation, shuffle reduce
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (blockIdx.x == 0 && threadIdx.x == 0 && training && nbt) *nbt += 1;
  if (c >= C) return;
  float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  float cb = conv_bias ? conv_bias[c] : 0.f;
  if (training) {
    double s = 0.0, q = 0.0;
    for (int r = lane; r < nparts; r += 32) {
      s += (double)partials[(size_t)r * 2 * C + c];
      q += (double)partials[(size_t)r * 2 * C + C + c];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if (lane != 0) return;
    double m = s / count;
    double var = q / count - m * m;
    if (var < 0.0) var = 0.0;
    float rstd = (float)(1.0 / sqrt(var + (double)eps));
    float mf = (float)m;
    scale[c] = g * rstd;
    shift[c] = b - mf * g * rstd;                    // conv bias cancels in train mode
    mean_out[c] = mf;
    rstd_out[c] = rstd;
    if (running_mean || stat_out) {
      const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
      const float bm = __fadd_rn(mf, cb), bv = (float)unbiased;
      if (stat_out) {                // deferred: k_bn_running_update applies the step's forwards in the reference's order
        stat_out[c] = bm;
        stat_out[C + c] = bv;
      } else {
        running_mean[c] = bn_running(running_mean[c], bm, momentum);
        running_var[c] = bn_running(running_var[c], bv, momentum);
      }
    }
  } else {
    if (lane != 0) return;
    float rstd = 1.0f / sqrtf(running_var[c] + eps);
    scale[c] = g * rstd;
    shift[c] = b + (cb - running_mean[c]) * g * rstd;
    mean_out[c] = running_mean[c] - cb;
    rstd_out[c] = rstd;
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
k_bn_act(const T* __restrict__ x, int ldx, const float* __restrict__ scale, const float* __restrict__ shift,
         int act, T* __restrict__ y, int ldy, long long npix, int C) {
  // (gridDim.x * 256) % (C/8) == 0 (host): one channel group per thread, scale/shift in registers
  const int CG = C >> 3;
  const long long n = npix * CG, stride = (long long)gridDim.x * blockDim.x;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int cg = (int)(i % CG);
  float sc[8], sh[8];
  Vec8<float>::load(scale + cg * 8, sc);
  Vec8<float>::load(shift + cg * 8, sh);
  const long long pstep = stride / CG;             // stride % CG == 0: no division inside the loop
  (void)n;
#pragma unroll 4
  for (long long p = i / CG; p < npix; p += pstep) {
    float v[8];
    Vec8<T>::load(x + p * ldx + cg * 8, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = act_fwd(fmaf(v[k], sc[k], sh[k]), act);
    Vec8<T>::store(y + p * ldy + cg * 8, v);
  }
}

// BN-apply + activation + 2x2 max pool: one thread = one 2x2 quad x 8 channels
template <typename T>
__global__ void k_bn_act_pool(const T* __restrict__ x, int ldx, const float* __restrict__ scale,
                              const float* __restrict__ shift, int act, T* __restrict__ y, int ldy, T* __restrict__ pooled,
                              int ldp, int B, int H, int W, int C) {
  int CG = C >> 3, Hp = H >> 1, Wp = W >> 1;
  long long n = (long long)B * Hp * Wp * CG;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const unsigned int iu = (unsigned int)i;
    const unsigned int qu = iu / (unsigned int)CG;
    int cg = (int)(iu - qu * (unsigned int)CG);
    long long q = qu;
    int wp, hp, b;
    pix_decomp(q, Wp, Hp, b, hp, wp);
    float sc[8], sh[8], m[8];
    Vec8<float>::load(scale + cg * 8, sc);
    Vec8<float>::load(shift + cg * 8, sh);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      long long p = ((long long)b * H + (hp * 2 + (j >> 1))) * W + (wp * 2 + (j & 1));
      float v[8];
      Vec8<T>::load(x + p * ldx + cg * 8, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        v[k] = act_fwd(fmaf(v[k], sc[k], sh[k]), act);
        // pool over the values as stored (rounded to T), like pooling the stored tensor
        float r = to_f(from_f<T>(v[k]));
        m[k] = (j == 0) ? r : fmaxf(m[k], r);
      }
      Vec8<T>::store(y + p * ldy + cg * 8, v);
    }
    Vec8<T>::store(pooled + q * ldp + cg * 8, m);
  }
}

// partial sums of (g', g'*xhat) per channel, g' = g * act'(bn(x)).  256 threads; a thread owns channel group
// tid % CG.  The loop accumulates sum g' and sum g'*x only (2 FMAs per element); xhat = (x - mean) * rstd is
// applied to the block totals: sum g'*xhat = rstd * (sum g'*x - mean * sum g').
template <typename T>
__global__ void __launch_bounds__(256, 3)
k_bn_bwd_reduce(const T* __restrict__ g, int ldg, const T* __restrict__ x, int ldx,
                const float* __restrict__ mean, const float* __restrict__ rstd,
                const float* __restrict__ scale, const float* __restrict__ shift, int act, long long npix,
                int C, float* __restrict__ partials) {
  extern __shared__ float sm[];   // [8][256]
  const int CG = C >> 3;
  const int tid = threadIdx.x;
  float acc[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) acc[k] = 0.f;
  {
    const int cg = tid % CG;
    const int lanes = 256 / CG;                       // pixel lanes per block
    float sc[8], sh[8];
    Vec8<float>::load(scale + cg * 8, sc);
    Vec8<float>::load(shift + cg * 8, sh);
    auto accum = [&](const Raw8<T>& gr, const Raw8<T>& xr) {
      float gv[8], xv[8];
      gr.get(gv);
      xr.get(xv);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float gp = gv[k] * act_grad(fmaf(xv[k], sc[k], sh[k]), act);
        acc[k] += gp;
        acc[8 + k] = fmaf(gp, xv[k], acc[8 + k]);
      }
    };
    constexpr int UB = sizeof(T) == 2 ? 4 : 1;        // pixels per iteration: 8 independent 16-byte loads in flight per thread (bf16)
    const long long step = (long long)gridDim.x * lanes;
    long long p = (long long)blockIdx.x * lanes + tid / CG;
    for (; p + (UB - 1) * step < npix; p += UB * step) {
      Raw8<T> gr[UB], xr[UB];
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        gr[u].load(g + (p + u * step) * ldg + cg * 8);
        xr[u].load(x + (p + u * step) * ldx + cg * 8);
      }
#pragma unroll
      for (int u = 0; u < UB; ++u) accum(gr[u], xr[u]);
    }
    for (; p < npix; p += step) {
      Raw8<T> gr, xr;
      gr.load(g + p * ldg + cg * 8);
      xr.load(x + p * ldx + cg * 8);
      accum(gr, xr);
    }
  }
  // block totals in two phases through one 8 KB buffer (sum g', then sum g'*x): with <= 9 KB of shared memory two of
  // these blocks fit next to a resident weight-gradient CTA (200 KB) that runs concurrently on the side stream
  const int lanes = 256 / CG;
  float s1r[8], s2r[8];                  // C <= 2048 (C/8 <= 256 channel groups): a thread finalises at most 8 channels
#pragma unroll
  for (int ph = 0; ph < 2; ++ph) {
#pragma unroll
    for (int k = 0; k < 8; ++k) sm[k * 256 + tid] = acc[ph * 8 + k];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = tid + i * 256;
      float s = 0.f;
      if (c < C) {
        const int cg = c >> 3, k = c & 7;
        for (int l = 0; l < lanes; ++l) s += sm[k * 256 + l * CG + cg];
      }
      if (ph == 0) s1r[i] = s; else s2r[i] = s;
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = tid + i * 256;
    if (c < C) {
      partials[(size_t)blockIdx.x * 2 * C + c] = s1r[i];
      partials[(size_t)blockIdx.x * 2 * C + C + c] = rstd[c] * (s2r[i] - mean[c] * s1r[i]);
    }
  }
}

__global__ void k_bn_bwd_finalize(const float* __restrict__ partials, int nparts, int C, double count,
                                  const float* __restrict__ gamma, const float* __restrict__ rstd, float* dgamma,
                                  float* dbeta, int accumulate, float param_grad_scale, float* coef) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= C) return;
  // up to 888 partial rows: four independent accumulator pairs keep eight loads in flight per lane (a single dependent
  // chain made this tiny kernel 14 us long, 72 times per step)
  double s1 = 0.0, s2 = 0.0, t1 = 0.0, t2 = 0.0, u1 = 0.0, u2 = 0.0, v1 = 0.0, v2 = 0.0;
  int r = lane;
  for (; r + 96 < nparts; r += 128) {
    const float a0 = partials[(size_t)r * 2 * C + c], b0 = partials[(size_t)r * 2 * C + C + c];
    const float a1 = partials[(size_t)(r + 32) * 2 * C + c], b1 = partials[(size_t)(r + 32) * 2 * C + C + c];
    const float a2 = partials[(size_t)(r + 64) * 2 * C + c], b2 = partials[(size_t)(r + 64) * 2 * C + C + c];
    const float a3 = partials[(size_t)(r + 96) * 2 * C + c], b3 = partials[(size_t)(r + 96) * 2 * C + C + c];
    s1 += (double)a0; s2 += (double)b0; t1 += (double)a1; t2 += (double)b1;
    u1 += (double)a2; u2 += (double)b2; v1 += (double)a3; v2 += (double)b3;
  }
  for (; r < nparts; r += 32) {
    s1 += (double)partials[(size_t)r * 2 * C + c];
    s2 += (double)partials[(size_t)r * 2 * C + C + c];
  }
  s1 += t1 + (u1 + v1);
  s2 += t2 + (u2 + v2);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if (lane != 0) return;
  if (dgamma) dgamma[c] = (accumulate ? dgamma[c] : 0.f) + param_grad_scale * (float)s2;
  if (dbeta) dbeta[c] = (accumulate ? dbeta[c] : 0.f) + param_grad_scale * (float)s1;
  if (coef) {               // nullable: the multi-lane step accumulates the parameter gradients in a second call on the side stream
    float g = gamma ? gamma[c] : 1.f;
    coef[c] = g * rstd[c];
    coef[C + c] = (float)(s1 / count);
    coef[2 * C + c] = (float)(s2 / count);
  }
}

template <typename T>
__global__ void __launch_bounds__(256, 3)
k_bn_bwd_apply(const T* __restrict__ g, int ldg, const T* __restrict__ x, int ldx, const float* __restrict__ mean,
               const float* __restrict__ rstd, const float* __restrict__ scale, const float* __restrict__ shift,
               const float* __restrict__ coef, int act, T* __restrict__ dx, int lddx, long long npix, int C) {
  // host guarantees (gridDim.x * 256) % (C/8) == 0: a thread keeps one channel group, so the 7 per-channel
  // vectors live in registers and the loop streams only g and x (2 x 16 B in, 16 B out)
  const int CG = C >> 3;
  const long long n = npix * CG, stride = (long long)gridDim.x * blockDim.x;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int cg = (int)(i % CG);
  // dx = a*(g' - b - xhat*d), xhat = (x - mean)*rstd  ==  a*g' + cx*x + c0 with cx = -a*d*rstd, c0 = -cx*mean - a*b:
  // five per-channel vectors stay live in the loop (a, cx, c0 and scale/shift for the activation mask)
  float sc[8], sh[8], a[8], cx[8], c0[8];
  {
    float mu[8], rs[8], b[8], d[8];
    Vec8<float>::load(mean + cg * 8, mu);
    Vec8<float>::load(rstd + cg * 8, rs);
    Vec8<float>::load(coef + cg * 8, a);
    Vec8<float>::load(coef + C + cg * 8, b);
    Vec8<float>::load(coef + 2 * C + cg * 8, d);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      cx[k] = -a[k] * d[k] * rs[k];
      c0[k] = -cx[k] * mu[k] - a[k] * b[k];
    }
  }
  Vec8<float>::load(scale + cg * 8, sc);
  Vec8<float>::load(shift + cg * 8, sh);
  const long long pstep = stride / CG;             // stride % CG == 0: no division inside the loop
  (void)n;
  auto apply = [&](const Raw8<T>& gr, const Raw8<T>& xr, long long p) {
    float gv[8], xv[8], o[8];
    gr.get(gv);
    xr.get(xv);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float gp = gv[k] * act_grad(fmaf(xv[k], sc[k], sh[k]), act);
      o[k] = fmaf(a[k], gp, fmaf(cx[k], xv[k], c0[k]));
    }
    Vec8<T>::store(dx + p * lddx + cg * 8, o);
  };
  constexpr int UB = sizeof(T) == 2 ? 3 : 1;          // pixels per iteration: 6 independent 16-byte loads in flight per thread (bf16)
  long long p = i / CG;
  for (; p + (UB - 1) * pstep < npix; p += UB * pstep) {
    Raw8<T> gr[UB], xr[UB];
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      gr[u].load(g + (p + u * pstep) * ldg + cg * 8);
      xr[u].load(x + (p + u * pstep) * ldx + cg * 8);
    }
#pragma unroll
    for (int u = 0; u < UB; ++u) apply(gr[u], xr[u], p + u * pstep);
  }
  for (; p < npix; p += pstep) {
    Raw8<T> gr, xr;
    gr.load(g + p * ldg + cg * 8);
    xr.load(x + p * ldx + cg * 8);
    apply(gr, xr, p);
  }
}

template <typename T>
__global__ void k_maxpool_bwd(const T* __restrict__ y, int ldy, const T* __restrict__ dpool, int ldp,
                              const T* __restrict__ gskip, int ldgs, T* __restrict__ gout, int ldgo, int B, int H, int W, int C) {
  int CG = C >> 3, Hp = H >> 1, Wp = W >> 1;
  long long n = (long long)B * Hp * Wp * CG;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const unsigned int iu = (unsigned int)i;
    const unsigned int qu = iu / (unsigned int)CG;
    int cg = (int)(iu - qu * (unsigned int)CG);
    long long q = qu;
    int wp, hp, b;
    pix_decomp(q, Wp, Hp, b, hp, wp);
    float v[4][8], dp[8];
    long long p[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      p[j] = ((long long)b * H + (hp * 2 + (j >> 1))) * W + (wp * 2 + (j & 1));
      Vec8<T>::load(y + p[j] * ldy + cg * 8, v[j]);
    }
    Vec8<T>::load(dpool + q * ldp + cg * 8, dp);
    int arg[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      int a = 0;
      float m = v[0][k];
#pragma unroll
      for (int j = 1; j < 4; ++j)
        if (v[j][k] > m) { m = v[j][k]; a = j; }      // strict > keeps the first max (ATen max_pool2d)
      arg[k] = a;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float o[8];
      if (gskip) Vec8<T>::load(gskip + p[j] * ldgs + cg * 8, o);
      else {
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = 0.f;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] += (arg[k] == j) ? dp[k] : 0.f;
      Vec8<T>::store(gout + p[j] * ldgo + cg * 8, o);
    }
  }
}

// ------------------------------------------------------------------------------------------
// bilinear x2 (ATen upsample_bilinear2d index rule: area_pixel_compute_source_index)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void src_index(int o, int in_size, int out_size, int align, int& i0, int& i1, float& l0, float& l1) {
  float r;
  if (align) {
    float sc = out_size > 1 ? (float)(in_size - 1) / (float)(out_size - 1) : 0.f;
    r = sc * o;
  } else {
    r = 0.5f * (o + 0.5f) - 0.5f;
    if (r < 0.f) r = 0.f;
  }
  i0 = (int)r;
  i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
  l1 = r - (float)i0;
  l0 = 1.f - l1;
}

template <typename T>
__global__ void k_upsample2x_fwd(const T* __restrict__ x, int ldx, T* __restrict__ y, int ldy, int B, int H, int W, int C, int align) {
  int CG = C >> 3, Ho = 2 * H, Wo = 2 * W;
  long long n = (long long)B * Ho * Wo * CG;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const unsigned int iu = (unsigned int)i;
    const unsigned int qu = iu / (unsigned int)CG;
    int cg = (int)(iu - qu * (unsigned int)CG);
    long long q = qu;
    int ow, oh, b;
    pix_decomp(q, Wo, Ho, b, oh, ow);
    int h0, h1, w0, w1;
    float lh0, lh1, lw0, lw1;
    src_index(oh, H, Ho, align, h0, h1, lh0, lh1);
    src_index(ow, W, Wo, align, w0, w1, lw0, lw1);
    const T* base = x + (size_t)b * H * W * ldx + cg * 8;
    float a[8], bb[8], c[8], d[8], o[8];
    Vec8<T>::load(base + ((size_t)h0 * W + w0) * ldx, a);
    Vec8<T>::load(base + ((size_t)h0 * W + w1) * ldx, bb);
    Vec8<T>::load(base + ((size_t)h1 * W + w0) * ldx, c);
    Vec8<T>::load(base + ((size_t)h1 * W + w1) * ldx, d);
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = lh0 * (lw0 * a[k] + lw1 * bb[k]) + lh1 * (lw0 * c[k] + lw1 * d[k]);
    Vec8<T>::store(y + q * ldy + cg * 8, o);
  }
}

// gather form of the backward: every input pixel collects from the <=5x5 outputs that can touch it
template <typename T>
__global__ void k_upsample2x_bwd(const T* __restrict__ dy, int lddy, T* __restrict__ dx, int lddx, int B, int H, int W, int C, int align) {
  int CG = C >> 3, Ho = 2 * H, Wo = 2 * W;
  long long n = (long long)B * H * W * CG;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const unsigned int iu = (unsigned int)i;
    const unsigned int qu = iu / (unsigned int)CG;
    int cg = (int)(iu - qu * (unsigned int)CG);
    long long q = qu;
    int w, h, b;
    pix_decomp(q, W, H, b, h, w);
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    int oh_lo, oh_hi, ow_lo, ow_hi;
    if (align) { oh_lo = 0; oh_hi = Ho - 1; ow_lo = 0; ow_hi = Wo - 1;
      // align_corners: outputs touching row h lie within [2h-2, 2h+3] as well (scale < 1)
      oh_lo = max(0, 2 * h - 2); oh_hi = min(Ho - 1, 2 * h + 3); ow_lo = max(0, 2 * w - 2); ow_hi = min(Wo - 1, 2 * w + 3);
    } else { oh_lo = max(0, 2 * h - 2); oh_hi = min(Ho - 1, 2 * h + 2); ow_lo = max(0, 2 * w - 2); ow_hi = min(Wo - 1, 2 * w + 2); }
    for (int oh = oh_lo; oh <= oh_hi; ++oh) {
      int h0, h1; float lh0, lh1;
      src_index(oh, H, Ho, align, h0, h1, lh0, lh1);
      float wh = (h0 == h ? lh0 : 0.f) + (h1 == h ? lh1 : 0.f);
      if (wh == 0.f) continue;
      for (int ow = ow_lo; ow <= ow_hi; ++ow) {
        int w0, w1; float lw0, lw1;
        src_index(ow, W, Wo, align, w0, w1, lw0, lw1);
        float ww = (w0 == w ? lw0 : 0.f) + (w1 == w ? lw1 : 0.f);
        if (ww == 0.f) continue;
        float g[8];
        Vec8<T>::load(dy + (((size_t)b * Ho + oh) * Wo + ow) * lddy + cg * 8, g);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += wh * ww * g[k];
      }
    }
    Vec8<T>::store(dx + q * lddx + cg * 8, acc);
  }
}

// ------------------------------------------------------------------------------------------
// per-channel sums (bias gradients)
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_channel_sum_partial(const T* __restrict__ x, int ldx, long long npix, int C, float* __restrict__ part) {
  __shared__ float sm[8][33];
  int c = blockIdx.y * 32 + threadIdx.x;
  float s = 0.f;
  if (c < C)
    for (long long p = (long long)blockIdx.x * 8 + threadIdx.y; p < npix; p += (long long)gridDim.x * 8) s += to_f(x[p * ldx + c]);
  sm[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += sm[j][threadIdx.x];
    part[(size_t)blockIdx.x * C + c] = t;
  }
}
// 128-bit variant: a thread owns channel group tid % (C/8); requires C % 8 == 0, ld % 8 == 0, 256 % (C/8) == 0
template <typename T>
__global__ void __launch_bounds__(256)
k_channel_sum_partial_v8(const T* __restrict__ x, int ldx, long long npix, int C, float* __restrict__ part) {
  __shared__ float sm[256][9];
  const int CG = C >> 3, tid = threadIdx.x;
  const int cg = tid % CG, lanes = 256 / CG;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll 4
  for (long long p = (long long)blockIdx.x * lanes + tid / CG; p < npix; p += (long long)gridDim.x * lanes) {
    float v[8];
    Vec8<T>::load(x + p * ldx + cg * 8, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] += v[k];
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) sm[tid][k] = acc[k];
  __syncthreads();
  for (int o = tid; o < C; o += 256) {
    const int g = o >> 3, k = o & 7;
    float s = 0.f;
    for (int l = 0; l < lanes; ++l) s += sm[l * CG + g][k];
    part[(size_t)blockIdx.x * C + o] = s;
  }
}
// one warp per channel: lanes stride over the partial rows
__global__ void k_channel_sum_final(const float* __restrict__ part, int nparts, int C, float* out, int accumulate) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= C) return;
  double s = 0.0;
  for (int r = lane; r < nparts; r += 32) s += (double)part[(size_t)r * C + c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[c] = (accumulate ? out[c] : 0.f) + (float)s;
}

}  // namespace ustrun

using namespace ustrun;

extern "C" {

int ustrun_abi_version(void) { return USTRUN_ABI_VERSION; }
const char* ustrun_last_error_string(void) { return ustrun::last_error(); }
int ustrun_device_supported(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { set_error("cudaGetDevice: %s", cudaGetErrorString(e)); return -(int)e; }
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) { set_error("cudaDeviceGetAttribute: %s", cudaGetErrorString(e)); return -(int)e; }
  return major == 10 ? 1 : 0;
}

int ustrun_nchw_to_nhwc(const float* src, void* dst, int dtype, int B, int C, int H, int W, int ld_dst, void* stream) {
  USTRUN_REQUIRE(src && dst && B > 0 && C > 0 && H > 0 && W > 0 && ld_dst >= C, "nchw_to_nhwc: bad args");
  if (C <= 4) {
    const long long HW = (long long)H * W, total = (long long)B * HW;
    const int g = grid_for(total, 256);
#define SMALL(CV) DISPATCH_DTYPE(dtype, (k_nchw_to_nhwc_small<T, CV><<<g, 256, 0, (cudaStream_t)stream>>>(src, (T*)dst, HW, ld_dst, total)))
    switch (C) { case 1: SMALL(1); break; case 2: SMALL(2); break; case 3: SMALL(3); break; default: SMALL(4); break; }
#undef SMALL
    return check_launch("nchw_to_nhwc_small");
  }
  dim3 grid(ceil_div((long long)H * W, 32), ceil_div(C, 32), B), block(32, 8);
  DISPATCH_DTYPE(dtype, (k_nchw_to_nhwc<T><<<grid, block, 0, (cudaStream_t)stream>>>(src, (T*)dst, C, H * W, ld_dst)));
  return check_launch("nchw_to_nhwc");
}
int ustrun_nhwc_to_nchw(const void* src, int dtype, int ld_src, float* dst, int B, int C, int H, int W, void* stream) {
  USTRUN_REQUIRE(src && dst && B > 0 && C > 0 && H > 0 && W > 0 && ld_src >= C, "nhwc_to_nchw: bad args");
  dim3 grid(ceil_div((long long)H * W, 32), ceil_div(C, 32), B), block(32, 8);
  DISPATCH_DTYPE(dtype, (k_nhwc_to_nchw<T><<<grid, block, 0, (cudaStream_t)stream>>>((const T*)src, dst, C, H * W, ld_src)));
  return check_launch("nhwc_to_nchw");
}
int ustrun_pack_conv_weight(const float* w, void* wf, void* wd, int dtype, int Cout, int Cin, int ksize, void* stream) {
  USTRUN_REQUIRE(w && (wf || wd) && Cout > 0 && Cin > 0 && (ksize == 1 || ksize == 3), "pack_conv_weight: bad args");
  const int taps = ksize * ksize;
  dim3 grid(ceil_div(Cout, 32), ceil_div(Cin, 32));
  const size_t smem = (size_t)32 * (32 * taps + 1) * sizeof(float);
  DISPATCH_DTYPE(dtype, (k_pack_conv<T><<<grid, 256, smem, (cudaStream_t)stream>>>(w, (T*)wf, (T*)wd, Cout, Cin, taps)));
  return check_launch("pack_conv_weight");
}
int ustrun_pack_weights_multi(const ustrun_pack_t* table, const int* blk_entry, const int* blk_tile, int nblocks, int dtype, void* stream) {
  USTRUN_REQUIRE(table && blk_entry && blk_tile && nblocks > 0, "pack_weights_multi: bad args");
  const size_t smem = (size_t)32 * (32 * 9 + 1) * sizeof(float);
  DISPATCH_DTYPE(dtype, (k_pack_multi<T><<<nblocks, 256, smem, (cudaStream_t)stream>>>(table, blk_entry, blk_tile)));
  return check_launch("pack_weights_multi");
}
int ustrun_pack_convT_weight(const float* w, void* wf, void* wd, int dtype, int Cin, int Cout, void* stream) {
  USTRUN_REQUIRE(w && (wf || wd) && Cout > 0 && Cin > 0, "pack_convT_weight: bad args");
  long long n = (long long)Cout * Cin * 4;
  DISPATCH_DTYPE(dtype, (k_pack_convT<T><<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(w, (T*)wf, (T*)wd, Cin, Cout)));
  return check_launch("pack_convT_weight");
}

int ustrun_bn_reduce_partials(const float* partials, int nparts, int C, float* sums, void* stream) {
  USTRUN_REQUIRE(partials && sums && nparts > 0 && C > 0, "bn_reduce_partials: bad args");
  k_bn_reduce_partials<<<ceil_div(2 * C, 128), 128, 0, (cudaStream_t)stream>>>(partials, nparts, C, sums);
  return check_launch("bn_reduce_partials");
}
int ustrun_bn_finalize(const float* partials, int nparts, int C, double count, const float* gamma, const float* beta,
                       const float* conv_bias, float* running_mean, float* running_var, long long* nbt, float momentum,
                       float eps, int training, float* scale, float* shift, float* mean, float* rstd, float* stat_out, void* stream) {
  USTRUN_REQUIRE(C > 0 && scale && shift && mean && rstd, "bn_finalize: bad args");
  USTRUN_REQUIRE(!stat_out || (training && !running_mean && !running_var && !nbt), "bn_finalize: stat_out replaces the in-place running update");
  USTRUN_REQUIRE(!training || (partials && nparts > 0 && count > 0), "bn_finalize: training needs partials");
  USTRUN_REQUIRE(training || (running_mean && running_var), "bn_finalize: eval needs running stats");
  PREFER_MAX_SHARED(k_bn_finalize);
  k_bn_finalize<<<ceil_div(C, 8), 256, 0, (cudaStream_t)stream>>>(partials, nparts, C, count, gamma, beta, conv_bias, running_mean,
                                                                   running_var, nbt, momentum, eps, training, scale, shift, mean, rstd, stat_out);
  return check_launch("bn_finalize");
}
int ustrun_bn_running_update(const ustrun_bn_update_t* table, int n, int Cmax, void* stream) {
  USTRUN_REQUIRE(table && n > 0 && Cmax > 0, "bn_running_update: bad args");
  dim3 grid(n, ceil_div(Cmax, 256));
  k_bn_running_update<<<grid, 256, 0, (cudaStream_t)stream>>>(table);
  return check_launch("bn_running_update");
}
int ustrun_bn_act_fwd(const void* x, int ldx, const float* scale, const float* shift, int act, void* y, int ldy, void* pooled,
                      int ldp, int dtype, int B, int H, int W, int C, void* stream) {
  USTRUN_REQUIRE(x && y && scale && shift && C % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0, "bn_act_fwd: need C, ld %% 8 == 0");
  long long npix = (long long)B * H * W;
  if (pooled) {
    USTRUN_REQUIRE(H % 2 == 0 && W % 2 == 0 && ldp % 8 == 0, "bn_act_fwd: pooling needs even H, W");
    long long n = npix / 4 * (C / 8);
    DISPATCH_DTYPE(dtype, (PREFER_MAX_SHARED(k_bn_act_pool<T>), k_bn_act_pool<T><<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>((const T*)x, ldx, scale, shift, act, (T*)y,
                                                                                               ldy, (T*)pooled, ldp, B, H, W, C)));
  } else {
    long long n = npix * (C / 8);
    int grid = grid_for(n, 256);
    while (((long long)grid * 256) % (C / 8)) ++grid;          // a thread must keep its channel group
    DISPATCH_DTYPE(dtype, (PREFER_MAX_SHARED(k_bn_act<T>), k_bn_act<T><<<grid, 256, 0, (cudaStream_t)stream>>>((const T*)x, ldx, scale, shift, act, (T*)y, ldy, npix, C)));
  }
  return check_launch("bn_act_fwd");
}
int ustrun_bn_bwd_reduce(const void* g, int ldg, const void* x, int ldx, const float* mean, const float* rstd, const float* scale,
                         const float* shift, int act, int dtype, long long npix, int C, float* partials, int* nparts_host, void* stream) {
  USTRUN_REQUIRE(g && x && mean && rstd && scale && shift && partials && nparts_host, "bn_bwd_reduce: null arg");
  USTRUN_REQUIRE(C % 8 == 0 && ldg % 8 == 0 && ldx % 8 == 0 && (C / 8) <= 256 && 256 % (C / 8) == 0, "bn_bwd_reduce: C=%d unsupported", C);
  int lanes = 256 / (C / 8);
  int grid = (int)((npix + 4LL * lanes - 1) / (4LL * lanes));       // >= 4 pixels per thread
  if (grid < 1) grid = 1;
  // ONE wave of 3 resident blocks per SM: the grid-stride loop takes any number of pixels, and every block is one more
  // partial row that k_bn_bwd_finalize has to walk (888 rows made that 2-CTA kernel 15 us long, 72 times per step)
  static int waves = 0;
  if (!waves) { const char* e = getenv("USTRUN_BN_BWD_WAVES"); waves = e ? atoi(e) : 1; if (waves < 1) waves = 1; }
  if (grid > 148 * 3 * waves) grid = 148 * 3 * waves;
  *nparts_host = grid;
  DISPATCH_DTYPE(dtype, (PREFER_MAX_SHARED(k_bn_bwd_reduce<T>), k_bn_bwd_reduce<T><<<grid, 256, 256 * 8 * sizeof(float), (cudaStream_t)stream>>>(
                            (const T*)g, ldg, (const T*)x, ldx, mean, rstd, scale, shift, act, npix, C, partials)));
  return check_launch("bn_bwd_reduce");
}
int ustrun_bn_bwd_finalize(const float* partials, int nparts, int C, double count, const float* gamma, const float* rstd, float* dgamma,
                           float* dbeta, int accumulate, float param_grad_scale, float* coef, void* stream) {
  USTRUN_REQUIRE(partials && nparts > 0 && C > 0 && rstd && count > 0 && (coef || dgamma || dbeta), "bn_bwd_finalize: bad args");
  PREFER_MAX_SHARED(k_bn_bwd_finalize);
  k_bn_bwd_finalize<<<ceil_div(C, 8), 256, 0, (cudaStream_t)stream>>>(partials, nparts, C, count, gamma, rstd, dgamma, dbeta, accumulate, param_grad_scale, coef);
  return check_launch("bn_bwd_finalize");
}
int ustrun_bn_bwd_apply(const void* g, int ldg, const void* x, int ldx, const float* mean, const float* rstd, const float* scale,
                        const float* shift, const float* coef, int act, void* dx, int lddx, int dtype, long long npix, int C, void* stream) {
  USTRUN_REQUIRE(g && x && dx && coef && C % 8 == 0 && ldg % 8 == 0 && ldx % 8 == 0 && lddx % 8 == 0, "bn_bwd_apply: bad args");
  long long n = npix * (C / 8);
  int grid = grid_for(n, 256);
  while (((long long)grid * 256) % (C / 8)) ++grid;            // a thread must keep its channel group
  DISPATCH_DTYPE(dtype, (PREFER_MAX_SHARED(k_bn_bwd_apply<T>), k_bn_bwd_apply<T><<<grid, 256, 0, (cudaStream_t)stream>>>((const T*)g, ldg, (const T*)x, ldx, mean, rstd, scale,
                                                                                  shift, coef, act, (T*)dx, lddx, npix, C)));
  return check_launch("bn_bwd_apply");
}
int ustrun_maxpool_bwd(const void* y, int ldy, const void* dpool, int ldp, const void* gskip, int ldgs, void* gout, int ldgo, int dtype,
                       int B, int H, int W, int C, void* stream) {
  USTRUN_REQUIRE(y && dpool && gout && C % 8 == 0 && H % 2 == 0 && W % 2 == 0 && ldy % 8 == 0 && ldp % 8 == 0 && ldgo % 8 == 0 &&
                     (!gskip || ldgs % 8 == 0), "maxpool_bwd: bad args");
  long long n = (long long)B * (H / 2) * (W / 2) * (C / 8);
  DISPATCH_DTYPE(dtype, (PREFER_MAX_SHARED(k_maxpool_bwd<T>), k_maxpool_bwd<T><<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>((const T*)y, ldy, (const T*)dpool, ldp, (const T*)gskip,
                                                                                             ldgs, (T*)gout, ldgo, B, H, W, C)));
  return check_launch("maxpool_bwd");
}
int ustrun_upsample2x_fwd(const void* x, int ldx, void* y, int ldy, int dtype, int B, int H, int W, int C, int align, void* stream) {
  USTRUN_REQUIRE(x && y && C % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0, "upsample2x_fwd: bad args");
  long long n = (long long)B * 4 * H * W * (C / 8);
  DISPATCH_DTYPE(dtype, (k_upsample2x_fwd<T><<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>((const T*)x, ldx, (T*)y, ldy, B, H, W, C, align)));
  return check_launch("upsample2x_fwd");
}
int ustrun_upsample2x_bwd(const void* dy, int lddy, void* dx, int lddx, int dtype, int B, int H, int W, int C, int align, void* stream) {
  USTRUN_REQUIRE(dy && dx && C % 8 == 0 && lddy % 8 == 0 && lddx % 8 == 0, "upsample2x_bwd: bad args");
  long long n = (long long)B * H * W * (C / 8);
  DISPATCH_DTYPE(dtype, (k_upsample2x_bwd<T><<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>((const T*)dy, lddy, (T*)dx, lddx, B, H, W, C, align)));
  return check_launch("upsample2x_bwd");
}
int ustrun_channel_sum(const void* x, int ldx, int dtype, long long npix, int C, float* out, int accumulate, float* workspace, void* stream) {
  USTRUN_REQUIRE(x && out && workspace && npix > 0 && C > 0, "channel_sum: bad args");
  int parts;
  if (C % 8 == 0 && ldx % 8 == 0 && (C / 8) <= 256 && 256 % (C / 8) == 0) {
    const int lanes = 256 / (C / 8);
    parts = (int)((npix + 4LL * lanes - 1) / (4LL * lanes));          // >= 4 pixels per thread
    if (parts > 148 * 6) parts = 148 * 6;
    if (parts < 1) parts = 1;
    DISPATCH_DTYPE(dtype, (k_channel_sum_partial_v8<T><<<parts, 256, 0, (cudaStream_t)stream>>>((const T*)x, ldx, npix, C, workspace)));
  } else {
    parts = (int)((npix + 63) / 64);
    if (parts > USTRUN_MAX_PARTS) parts = USTRUN_MAX_PARTS;
    dim3 grid(parts, ceil_div(C, 32)), block(32, 8);
    DISPATCH_DTYPE(dtype, (k_channel_sum_partial<T><<<grid, block, 0, (cudaStream_t)stream>>>((const T*)x, ldx, npix, C, workspace)));
  }
  k_channel_sum_final<<<ceil_div(C, 8), 256, 0, (cudaStream_t)stream>>>(workspace, parts, C, out, accumulate);
  return check_launch("channel_sum");
}

}  // extern "C"
