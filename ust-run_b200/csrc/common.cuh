// Shared device/host helpers for libustrun_sm100.so (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ustrun.h"

namespace ustrun {

// ---- error plumbing: every extern "C" entry returns 0 or a negative / cudaError code ---------
void set_error(const char* fmt, ...);
int check_launch(const char* what);   // cudaGetLastError -> code (0 ok)

#define USTRUN_REQUIRE(cond, ...)                 \
  do {                                            \
    if (!(cond)) {                                \
      ::ustrun::set_error(__VA_ARGS__);           \
      return USTRUN_ERR_ARG;                      \
    }                                             \
  } while (0)

// geometry of a CUDA-core implicit-GEMM convolution (simt_conv.cu)
struct ConvGeom {
  int B, H, W;        // GEMM-row pixel grid (for gather==1: the LOW-res grid)
  int Cin, Cout, ks;  // ks = 1 or 3 (gather==0); gather==1 has 4 "taps"
  int gather;         // 0: conv taps with zero padding, 1: taps = (i,j) of the 2x up-sampled grid
  int scatter_ij;     // >=0: write output pixel (2h+i, 2w+j) of a [B,2H,2W] grid (convT forward)
};

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- scalar conversions -------------------------------------------------------------------
__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// ---- 8-element vector access (16 B of bf16 / 2 x 16 B of fp32) ------------------------------
template <typename T> struct Vec8;
template <> struct Vec8<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 r = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x;
      v[2 * i + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 r;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = r;
  }
};
template <> struct Vec8<float> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
    float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
};

// raw 8-element loads (kept unconverted in registers so that several independent loads can be issued back to back
// before the first conversion: the streaming kernels batch 2-4 pixels per loop iteration by hand -- with a bounds check
// per unrolled iteration the compiler serialises load -> use -> load and only one pixel's loads are in flight)
template <typename T> struct Raw8;
template <> struct Raw8<__nv_bfloat16> {
  uint4 r;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { r = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void get(float (&v)[8]) const {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x;
      v[2 * i + 1] = f.y;
    }
  }
};
template <> struct Raw8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) { a = *reinterpret_cast<const float4*>(p); b = *reinterpret_cast<const float4*>(p + 4); }
  __device__ __forceinline__ void get(float (&v)[8]) const { v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w; }
};

// linear pixel index -> (b, h, w) with 32-bit arithmetic (64-bit div/mod costs ~100 instructions each;
// every entry point checks B*H*W < 2^31)
__device__ __forceinline__ void pix_decomp(long long p, int W, int H, int& b, int& h, int& w) {
  const unsigned int q = (unsigned int)p;
  const unsigned int t = q / (unsigned int)W;
  w = (int)(q - t * (unsigned int)W);
  const unsigned int bb = t / (unsigned int)H;
  h = (int)(t - bb * (unsigned int)H);
  b = (int)bb;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// running <- (1 - momentum) * running + momentum * batch, with the rounding points pinned (no FMA contraction) so that the
// in-place update (k_bn_finalize, k_bn_finalize_peer) and the deferred one (k_bn_running_update) are bit-identical
__device__ __forceinline__ float bn_running(float running, float batch, float momentum) {
  return __fadd_rn(__fmul_rn(1.f - momentum, running), __fmul_rn(momentum, batch));
}

// activation codes shared with the host side
__device__ __forceinline__ float act_fwd(float y, int act) {
  if (act == USTRUN_ACT_RELU) return y > 0.f ? y : 0.f;
  if (act == USTRUN_ACT_LEAKY) return y > 0.f ? y : 0.01f * y;
  return y;
}
__device__ __forceinline__ float act_grad(float y, int act) {   // derivative given pre-activation y
  if (act == USTRUN_ACT_RELU) return y > 0.f ? 1.f : 0.f;
  if (act == USTRUN_ACT_LEAKY) return y > 0.f ? 1.f : 0.01f;
  return 1.f;
}

}  // namespace ustrun
