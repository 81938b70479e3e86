// extern "C" convolution entry points: argument checks + dispatch to the CUDA-core (simt_conv.cu)
// or tcgen05 (tc_conv.cu) implementation.  `impl` is the caller's explicit choice -- there is no
// silent fallback: an unsupported (impl, shape, dtype) combination is an argument error.
#include "common.cuh"

namespace ustrun {
template <typename T>
int simt_conv_launch(const void* x, int ldx, const void* w, const float* bias, void* y, int ldy, float* y_nchw, ConvGeom g, float* partials,
                     int* nparts_host, cudaStream_t st);
template <typename T>
int simt_wgrad_launch(const void* a, int lda, const void* b, int ldb, float* dw, int accumulate, ConvGeom g, int Mo, int Nin, void* workspace,
                      long long ws_bytes, cudaStream_t st);
long long simt_wgrad_ws_bytes(long long P, int Mo, int Nin, int taps);
bool narrow_in_ok(int Cin, int Cout, int ks);
bool narrow_out_ok(int Cin, int Cout, int ks);
bool narrow_wgrad_ok(int Cw, int Cn, int ks);
long long narrow_wgrad_ws_bytes(long long M, int Cw, int Cn, int ks);
template <typename T>
int narrow_in_launch(const void* x, int ldx, const void* w, const float* bias, void* y, int ldy, int B, int H, int W, int Cin, int Cout, int ks,
                     float* partials, int* nparts_host, cudaStream_t st);
template <typename T>
int narrow_out_launch(const void* x, int ldx, const void* w, const float* bias, void* y, int ldy, float* y_nchw, int B, int H, int W, int Cin,
                      int Cout, int ks, cudaStream_t st);
template <typename T>
int narrow_wgrad_launch(const void* wide, int ldw, int Cw, const void* nar, int ldn, int Cn, int ks, int sgn, int mode, int B, int H, int W,
                        float* dw, int accumulate, void* workspace, long long ws_bytes, cudaStream_t st);

bool mid_conv_ok(int Cin, int Cout, int ks, int W, int ldx, int ldy);
bool mid_wgrad_ok(int Cin, int Cout, int ks, int ldx, int lddy);
long long mid_wgrad_ws_bytes(int Cin, int Cout, int ks);
int mid_conv_launch(const void* x, int ldx, const void* w, const float* bias, void* y, int ldy, int B, int H, int W, int Cin, int Cout, int ks,
                    float* partials, int* nparts_host, cudaStream_t st);
bool mid_head_ok(int Cin, int Cout, int ks, int W, int ldx);
int mid_head_launch(const void* x, int ldx, const void* w, const float* bias, float* y_nchw, int B, int H, int W, int Cin, int Cout, cudaStream_t st);
int mid_wgrad_launch(const void* dy, int lddy, const void* x, int ldx, float* dw, int accumulate, int B, int H, int W, int Cin, int Cout, int ks,
                     void* workspace, long long ws_bytes, cudaStream_t st);

int tc_conv_fwd(const void* x, int ldx, const void* w, const float* bias, void* y, int ldy, int B, int H, int W, int Cin, int Cout, int ksize,
                float* partials, int* nparts_host, cudaStream_t st);
int tc_convT_fwd(const void* x, int ldx, const void* wf, const float* bias, void* y, int ldy, int B, int H, int W, int Cin, int Cout, cudaStream_t st);
int tc_convT_dgrad(const void* dy, int lddy, const void* wd, void* dx, int lddx, int B, int H, int W, int Cin, int Cout, cudaStream_t st);
int tc_conv_wgrad(const void* dy, int lddy, const void* x, int ldx, float* dw, int accumulate, int B, int H, int W, int Cin, int Cout, int ksize,
                  void* workspace, long long ws_bytes, cudaStream_t st);
long long tc_conv_wgrad_ws(int B, int H, int W, int Cin, int Cout, int ksize);
int tc_convT_wgrad(const void* dy, int lddy, const void* x, int ldx, float* dw, int accumulate, int B, int H, int W, int Cin, int Cout,
                   void* workspace, long long ws_bytes, cudaStream_t st);
long long tc_convT_wgrad_ws(int B, int H, int W, int Cin, int Cout);
int tc_plan_query(int what, int B, int H, int W, int Cin, int Cout, int ksize, int* out);
}  // namespace ustrun

using namespace ustrun;

#define CHECK_COMMON(name)                                                                                            \
  USTRUN_REQUIRE(B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, name ": empty shape");                               \
  USTRUN_REQUIRE((long long)B * H * W < (1LL << 31), name ": B*H*W must be < 2^31 (32-bit pixel indexing)");            \
  USTRUN_REQUIRE(dtype == USTRUN_F32 || dtype == USTRUN_BF16, name ": bad dtype");                                    \
  USTRUN_REQUIRE(impl == USTRUN_IMPL_SIMT || (impl == USTRUN_IMPL_TCGEN05 && dtype == USTRUN_BF16), name ": tcgen05 path is bf16 only")

extern "C" {

int ustrun_conv_fwd(int impl, const void* x, int ldx, const void* w_packed, const float* bias, void* y, int ldy, int dtype, int B, int H,
                    int W, int Cin, int Cout, int ksize, int out_nchw_f32, float* partials, int* nparts_host, void* stream) {
  CHECK_COMMON("conv_fwd");
  USTRUN_REQUIRE(x && w_packed && y && (ksize == 1 || ksize == 3) && ldx >= Cin, "conv_fwd: bad args");
  USTRUN_REQUIRE(!partials || nparts_host, "conv_fwd: partials needs nparts_host");
  cudaStream_t st = (cudaStream_t)stream;
  if (impl == USTRUN_IMPL_TCGEN05) {
    USTRUN_REQUIRE(!out_nchw_f32, "conv_fwd: tcgen05 path writes NHWC bf16 only");
    return tc_conv_fwd(x, ldx, w_packed, bias, y, ldy, B, H, W, Cin, Cout, ksize, partials, nparts_host, st);
  }
  float* ynchw = out_nchw_f32 ? (float*)y : nullptr;
  if (!out_nchw_f32 && narrow_in_ok(Cin, Cout, ksize) && ldy % 8 == 0) {      // first conv / head dgrad: HBM-bound streaming kernel
    if (dtype == USTRUN_F32) return narrow_in_launch<float>(x, ldx, w_packed, bias, y, ldy, B, H, W, Cin, Cout, ksize, partials, nparts_host, st);
    return narrow_in_launch<__nv_bfloat16>(x, ldx, w_packed, bias, y, ldy, B, H, W, Cin, Cout, ksize, partials, nparts_host, st);
  }
  if (dtype == USTRUN_BF16 && out_nchw_f32 && !partials && mid_head_ok(Cin, Cout, ksize, W, ldx))     // UNet-B logits head (3x3): warp-level MMA
    return mid_head_launch(x, ldx, w_packed, bias, ynchw, B, H, W, Cin, Cout, st);
  if (!partials && narrow_out_ok(Cin, Cout, ksize) && ldx % 8 == 0) {          // logits head
    if (dtype == USTRUN_F32) return narrow_out_launch<float>(x, ldx, w_packed, bias, y, ldy, ynchw, B, H, W, Cin, Cout, ksize, st);
    return narrow_out_launch<__nv_bfloat16>(x, ldx, w_packed, bias, y, ldy, ynchw, B, H, W, Cin, Cout, ksize, st);
  }
  if (dtype == USTRUN_BF16 && !out_nchw_f32 && mid_conv_ok(Cin, Cout, ksize, W, ldx, ldy))       // UNet-B 16/32-channel levels: warp-level MMA
    return mid_conv_launch(x, ldx, w_packed, bias, y, ldy, B, H, W, Cin, Cout, ksize, partials, nparts_host, st);
  ConvGeom g{B, H, W, Cin, Cout, ksize, 0, -1};
  if (dtype == USTRUN_F32) return simt_conv_launch<float>(x, ldx, w_packed, bias, y, ldy, ynchw, g, partials, nparts_host, st);
  return simt_conv_launch<__nv_bfloat16>(x, ldx, w_packed, bias, y, ldy, ynchw, g, partials, nparts_host, st);
}

int ustrun_tc_plan_query(int what, int B, int H, int W, int Cin, int Cout, int ksize, int* out8) {
  USTRUN_REQUIRE(out8 && B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && Cin % 64 == 0 && Cout % 64 == 0 && (ksize == 1 || ksize == 3),
                 "tc_plan_query: bad args");
  return tc_plan_query(what, B, H, W, Cin, Cout, ksize, out8);
}

long long ustrun_conv_wgrad_workspace_bytes(int impl, int B, int H, int W, int Cin, int Cout, int ksize) {
  // ksize 2 = ConvTranspose2d(k2,s2): Mo = Cin, N = 4*Cout
  if (ksize == 2) return impl == USTRUN_IMPL_TCGEN05 ? tc_convT_wgrad_ws(B, H, W, Cin, Cout) : simt_wgrad_ws_bytes((long long)B * H * W, Cin, Cout, 4);
  if (impl == USTRUN_IMPL_TCGEN05) return tc_conv_wgrad_ws(B, H, W, Cin, Cout, ksize);
  long long generic = simt_wgrad_ws_bytes((long long)B * H * W, Cout, Cin, ksize * ksize), narrow = 0;
  if (Cin <= 8 && narrow_wgrad_ok(Cout, Cin, ksize)) narrow = narrow_wgrad_ws_bytes((long long)B * H * W, Cout, Cin, ksize);
  else if (Cout <= 8 && narrow_wgrad_ok(Cin, Cout, ksize)) narrow = narrow_wgrad_ws_bytes((long long)B * H * W, Cin, Cout, ksize);
  if (mid_wgrad_ok(Cin, Cout, ksize, 8, 8)) {
    const long long mid = mid_wgrad_ws_bytes(Cin, Cout, ksize);
    if (mid > narrow) narrow = mid;
  }
  return generic > narrow ? generic : narrow;
}

int ustrun_conv_wgrad(int impl, const void* dy, int lddy, const void* x, int ldx, float* dw, int accumulate, int dtype, int B, int H, int W,
                      int Cin, int Cout, int ksize, void* workspace, long long workspace_bytes, void* stream) {
  CHECK_COMMON("conv_wgrad");
  USTRUN_REQUIRE(dy && x && dw && (ksize == 1 || ksize == 3), "conv_wgrad: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  if (impl == USTRUN_IMPL_TCGEN05) return tc_conv_wgrad(dy, lddy, x, ldx, dw, accumulate, B, H, W, Cin, Cout, ksize, workspace, workspace_bytes, st);
  if (Cin <= 8 && narrow_wgrad_ok(Cout, Cin, ksize) && lddy % 8 == 0) {        // first conv: wide = dy, narrow = x shifted by +tap
    if (dtype == USTRUN_F32) return narrow_wgrad_launch<float>(dy, lddy, Cout, x, ldx, Cin, ksize, +1, 0, B, H, W, dw, accumulate, workspace, workspace_bytes, st);
    return narrow_wgrad_launch<__nv_bfloat16>(dy, lddy, Cout, x, ldx, Cin, ksize, +1, 0, B, H, W, dw, accumulate, workspace, workspace_bytes, st);
  }
  if (Cout <= 8 && narrow_wgrad_ok(Cin, Cout, ksize) && ldx % 8 == 0) {         // logits head: wide = x, narrow = dy shifted by -tap
    if (dtype == USTRUN_F32) return narrow_wgrad_launch<float>(x, ldx, Cin, dy, lddy, Cout, ksize, -1, 1, B, H, W, dw, accumulate, workspace, workspace_bytes, st);
    return narrow_wgrad_launch<__nv_bfloat16>(x, ldx, Cin, dy, lddy, Cout, ksize, -1, 1, B, H, W, dw, accumulate, workspace, workspace_bytes, st);
  }
  if (dtype == USTRUN_BF16 && mid_wgrad_ok(Cin, Cout, ksize, ldx, lddy))
    return mid_wgrad_launch(dy, lddy, x, ldx, dw, accumulate, B, H, W, Cin, Cout, ksize, workspace, workspace_bytes, st);
  ConvGeom g{B, H, W, Cin, Cout, ksize, 0, -1};
  if (dtype == USTRUN_F32) return simt_wgrad_launch<float>(dy, lddy, x, ldx, dw, accumulate, g, Cout, Cin, workspace, workspace_bytes, st);
  return simt_wgrad_launch<__nv_bfloat16>(dy, lddy, x, ldx, dw, accumulate, g, Cout, Cin, workspace, workspace_bytes, st);
}

int ustrun_convT2x2_fwd(int impl, const void* x, int ldx, const void* wf, const float* bias, void* y, int ldy, int dtype, int B, int H, int W,
                        int Cin, int Cout, void* stream) {
  CHECK_COMMON("convT2x2_fwd");
  USTRUN_REQUIRE(x && wf && y, "convT2x2_fwd: null arg");
  cudaStream_t st = (cudaStream_t)stream;
  if (impl == USTRUN_IMPL_TCGEN05) return tc_convT_fwd(x, ldx, wf, bias, y, ldy, B, H, W, Cin, Cout, st);
  const size_t esz = dtype == USTRUN_F32 ? 4 : 2;
  for (int ij = 0; ij < 4; ++ij) {
    ConvGeom g{B, H, W, Cin, Cout, 1, 0, ij};
    const void* wij = (const char*)wf + (size_t)ij * Cout * Cin * esz;
    int rc = dtype == USTRUN_F32 ? simt_conv_launch<float>(x, ldx, wij, bias, y, ldy, nullptr, g, nullptr, nullptr, st)
                                 : simt_conv_launch<__nv_bfloat16>(x, ldx, wij, bias, y, ldy, nullptr, g, nullptr, nullptr, st);
    if (rc) return rc;
  }
  return 0;
}

int ustrun_convT2x2_dgrad(int impl, const void* dy, int lddy, const void* wd, void* dx, int lddx, int dtype, int B, int H, int W, int Cin,
                          int Cout, void* stream) {
  CHECK_COMMON("convT2x2_dgrad");
  USTRUN_REQUIRE(dy && wd && dx, "convT2x2_dgrad: null arg");
  cudaStream_t st = (cudaStream_t)stream;
  if (impl == USTRUN_IMPL_TCGEN05) return tc_convT_dgrad(dy, lddy, wd, dx, lddx, B, H, W, Cin, Cout, st);
  ConvGeom g{B, H, W, /*gathered channels*/ Cout, /*outputs*/ Cin, 1, 1, -1};
  if (dtype == USTRUN_F32) return simt_conv_launch<float>(dy, lddy, wd, nullptr, dx, lddx, nullptr, g, nullptr, nullptr, st);
  return simt_conv_launch<__nv_bfloat16>(dy, lddy, wd, nullptr, dx, lddx, nullptr, g, nullptr, nullptr, st);
}

int ustrun_convT2x2_wgrad(int impl, const void* dy, int lddy, const void* x, int ldx, float* dw, int accumulate, int dtype, int B, int H, int W,
                          int Cin, int Cout, void* workspace, long long workspace_bytes, void* stream) {
  CHECK_COMMON("convT2x2_wgrad");
  USTRUN_REQUIRE(dy && x && dw, "convT2x2_wgrad: null arg");
  cudaStream_t st = (cudaStream_t)stream;
  if (impl == USTRUN_IMPL_TCGEN05) return tc_convT_wgrad(dy, lddy, x, ldx, dw, accumulate, B, H, W, Cin, Cout, workspace, workspace_bytes, st);
  ConvGeom g{B, H, W, Cout, Cin, 1, 1, -1};
  if (dtype == USTRUN_F32) return simt_wgrad_launch<float>(x, ldx, dy, lddy, dw, accumulate, g, Cin, Cout, workspace, workspace_bytes, st);
  return simt_wgrad_launch<__nv_bfloat16>(x, ldx, dy, lddy, dw, accumulate, g, Cin, Cout, workspace, workspace_bytes, st);
}

}  // extern "C"
