// Fused multi-tensor SGD(momentum, weight decay) + EMA teacher update (K15+K16): ONE launch over
// every parameter tensor instead of torch's per-tensor foreach passes and the reference's
// 2 launches x #tensors EMA loop (train.py:92-93).  28 B/param of HBM traffic, 128-bit accesses.
#include "common.cuh"

namespace ustrun {

__global__ void __launch_bounds__(256) k_sgd_ema(const ustrun_param_t* __restrict__ table, const int* __restrict__ blk_tensor,
                                                const long long* __restrict__ blk_offset, float lr, float mu, float wd, float alpha,
                                                float gscale, int do_sgd, int do_ema, const float* __restrict__ hyper) {
  if (hyper) {            // per-step scalars in device memory: a captured CUDA graph replays with new values
    lr = hyper[0];
    alpha = hyper[1];
    gscale = hyper[2];
  }
  const ustrun_param_t t = table[blk_tensor[blockIdx.x]];
  const long long off = blk_offset[blockIdx.x];
  long long cnt = t.n - off;
  if (cnt > USTRUN_OPT_CHUNK) cnt = USTRUN_OPT_CHUNK;
  float* p = t.p + off;
  float* g = t.g ? t.g + off : nullptr;
  float* buf = t.buf ? t.buf + off : nullptr;
  float* ema = t.ema ? t.ema + off : nullptr;
  const bool sgd = do_sgd && g != nullptr && buf != nullptr;
  const bool do_e = do_ema && ema != nullptr;
  const float oma = 1.f - alpha;
  const bool vec = (((uintptr_t)p | (uintptr_t)g | (uintptr_t)buf | (uintptr_t)ema) & 15) == 0 && (cnt & 3) == 0;
  if (vec) {
    for (long long i = threadIdx.x * 4; i < cnt; i += 256 * 4) {
      float4 pv = *reinterpret_cast<float4*>(p + i);
      if (sgd) {
        float4 gv = *reinterpret_cast<float4*>(g + i);
        float4 bv = t.first ? make_float4(0, 0, 0, 0) : *reinterpret_cast<float4*>(buf + i);
        gv.x = fmaf(wd, pv.x, gv.x * gscale); gv.y = fmaf(wd, pv.y, gv.y * gscale);
        gv.z = fmaf(wd, pv.z, gv.z * gscale); gv.w = fmaf(wd, pv.w, gv.w * gscale);
        if (t.first) bv = gv;
        else { bv.x = fmaf(mu, bv.x, gv.x); bv.y = fmaf(mu, bv.y, gv.y); bv.z = fmaf(mu, bv.z, gv.z); bv.w = fmaf(mu, bv.w, gv.w); }
        pv.x = fmaf(-lr, bv.x, pv.x); pv.y = fmaf(-lr, bv.y, pv.y); pv.z = fmaf(-lr, bv.z, pv.z); pv.w = fmaf(-lr, bv.w, pv.w);
        *reinterpret_cast<float4*>(buf + i) = bv;
        *reinterpret_cast<float4*>(p + i) = pv;
      }
      if (do_e) {
        float4 ev = *reinterpret_cast<float4*>(ema + i);
        ev.x = fmaf(oma, pv.x, alpha * ev.x); ev.y = fmaf(oma, pv.y, alpha * ev.y);
        ev.z = fmaf(oma, pv.z, alpha * ev.z); ev.w = fmaf(oma, pv.w, alpha * ev.w);
        *reinterpret_cast<float4*>(ema + i) = ev;
      }
    }
  } else {
    for (long long i = threadIdx.x; i < cnt; i += 256) {
      float pv = p[i];
      if (sgd) {
        float gv = fmaf(wd, pv, g[i] * gscale);
        float bv = t.first ? gv : fmaf(mu, buf[i], gv);
        pv = fmaf(-lr, bv, pv);
        buf[i] = bv;
        p[i] = pv;
      }
      if (do_e) ema[i] = fmaf(oma, pv, alpha * ema[i]);
    }
  }
}

}  // namespace ustrun

using namespace ustrun;

extern "C" int ustrun_sgd_ema_multi(const ustrun_param_t* table, const int* blk_tensor, const long long* blk_offset, int nblocks, float lr,
                                    float momentum, float weight_decay, float alpha, float grad_scale, int do_sgd, int do_ema, void* stream) {
  USTRUN_REQUIRE(table && blk_tensor && blk_offset && nblocks > 0, "sgd_ema_multi: bad args");
  k_sgd_ema<<<nblocks, 256, 0, (cudaStream_t)stream>>>(table, blk_tensor, blk_offset, lr, momentum, weight_decay, alpha, grad_scale, do_sgd, do_ema, nullptr);
  return check_launch("sgd_ema_multi");
}

extern "C" int ustrun_sgd_ema_multi_dev(const ustrun_param_t* table, const int* blk_tensor, const long long* blk_offset, int nblocks,
                                        const float* hyper, float momentum, float weight_decay, int do_sgd, int do_ema, void* stream) {
  USTRUN_REQUIRE(table && blk_tensor && blk_offset && nblocks > 0 && hyper, "sgd_ema_multi_dev: bad args");
  k_sgd_ema<<<nblocks, 256, 0, (cudaStream_t)stream>>>(table, blk_tensor, blk_offset, 0.f, momentum, weight_decay, 0.f, 1.f, do_sgd, do_ema, hyper);
  return check_launch("sgd_ema_multi_dev");
}
