// Micro-benchmark: tcgen05.mma rate (bf16, M=128) as a function of N, operand major-ness and operand placement.
// One CTA per SM, one elected thread issues (elect.sync => no waterfall loops), descriptors precomputed.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int N, int amn, int bmn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)amn << 15) | ((uint32_t)bmn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\t@px mov.s32 %0, 1;\n\t}" : "+r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mma(uint32_t tm, uint64_t a, uint64_t b, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tm), "l"(a), "l"(b), "r"(idesc), "r"(1));
}
// mode 0: 12 MMAs per iteration = conv ROW pattern (3 row-shifted A views x 4 K slices, 3 different B slots)
// mode 1: same descriptors every time
__global__ void __launch_bounds__(128, 1) k(int N, int amn, int bmn, int reps, int aoff, int boff, int mode, long long* out, const uint8_t* gsrc = nullptr, int copy_kb = 0, int coff = 0) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint64_t cbar[4];
  __shared__ volatile int done;
  __shared__ uint32_t holder;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&cbar[i])));
    done = 0;
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&holder)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tm = holder;
  if (threadIdx.x < 32) {
    const uint32_t idesc = make_idesc(N, amn, bmn);
    const uint32_t sa = smem_u32(smem) + aoff, sb = smem_u32(smem) + boff;
    const uint64_t ad = amn ? make_sdesc(sa, 8192, 1024) : make_sdesc(sa, 16, 1024);
    const uint64_t bd = bmn ? make_sdesc(sb, 9216, 1024) : make_sdesc(sb, 16, 1024);
    const uint64_t kstep = (amn ? 128 : 2), kstepb = (bmn ? 128 : 2);
    const uint64_t bslot = (uint64_t)(N * 128) >> 4;
    long long t0 = clock64();
    if (elect_one()) {
      for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int dx = 0; dx < 3; ++dx)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            mma(tm, ad + kk * kstep + (mode == 0 ? dx * 8 : 0), bd + kk * kstepb + (mode == 0 ? dx * bslot : 0), idesc);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
    }
    __syncwarp();
    asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&bar)));
    long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) *out = t1 - t0;
    done = 1;
  } else if (threadIdx.x >= 32 && threadIdx.x < 64 && copy_kb > 0) {
    // concurrent TMA-style traffic: 4 bulk copies of copy_kb KB in flight into shared memory at coff
    if (elect_one()) {
      long long n = 0;
      uint32_t ph[4] = {0, 0, 0, 0};
      const uint32_t bytes = copy_kb * 1024;
      const uint8_t* src = gsrc + (size_t)blockIdx.x * 4 * bytes;
      for (int i = 0; i < 4; ++i) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&cbar[i])), "r"(bytes));
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem) + coff + i * bytes), "l"(src + (size_t)i * bytes), "r"(bytes), "r"(smem_u32(&cbar[i])));
      }
      while (!done) {
        for (int i = 0; i < 4; ++i) {
          asm volatile("{\n\t.reg .pred p;\n\tW2:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D2;\n\tbra W2;\n\tD2:\n\t}" ::"r"(smem_u32(&cbar[i])), "r"(ph[i]));
          ph[i] ^= 1;
          ++n;
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&cbar[i])), "r"(bytes));
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem) + coff + i * bytes), "l"(src + (size_t)i * bytes), "r"(bytes), "r"(smem_u32(&cbar[i])));
        }
      }
      for (int i = 0; i < 4; ++i)
        asm volatile("{\n\t.reg .pred p;\n\tW3:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D3;\n\tbra W3;\n\tD3:\n\t}" ::"r"(smem_u32(&cbar[i])), "r"(ph[i]));
      if (blockIdx.x == 0) out[1] = n * bytes;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm));
}
int main() {
  long long* d;
  cudaMalloc(&d, 16);
  uint8_t* gsrc;
  cudaMalloc(&gsrc, 148 * 4 * 32 * 1024);
  cudaMemset(gsrc, 0, 148 * 4 * 32 * 1024);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  const int reps = 1000;
  auto run = [&](int N, int amn, int bmn, int aoff, int boff, int mode) {
    k<<<148, 128, 220 * 1024>>>(N, amn, bmn, reps, aoff, boff, mode, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    printf("N=%3d A=%s B=%s aoff=%6d boff=%6d mode=%d: %6.1f cycles/MMA (nominal %d) %s\n", N, amn ? "MN" : "K ", bmn ? "MN" : "K ", aoff, boff, mode,
           (double)c / (reps * 12), N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
  };
  for (int N : {64, 128, 192, 256}) {
    run(N, 0, 0, 0, 32768, 1);
    run(N, 1, 1, 0, 32768, 1);
  }
  run(64, 0, 0, 0, 17408, 0);                // non-resident layout: B right behind the A box
  run(64, 0, 0, 73728, 0, 0);                // resident layout: weights at 0, A ring behind them
  run(64, 0, 0, 73728 + 17408, 0, 0);
  run(64, 0, 0, 73728 + 3 * 17408, 24576, 0);
  run(128, 0, 0, 0, 17408, 0);
  run(128, 0, 0, 147456, 0, 0);
  run(128, 0, 0, 147456 + 17408, 49152, 0);
  auto runc = [&](int N, int copy_kb, int coff) {
    cudaMemset(d, 0, 16);
    k<<<148, 128, 220 * 1024>>>(N, 0, 0, reps, 0, 32768, 0, d, gsrc, copy_kb, coff);
    cudaError_t e = cudaDeviceSynchronize();
    long long c[2] = {0, 0};
    cudaMemcpy(c, d, 16, cudaMemcpyDeviceToHost);
    printf("N=%3d with concurrent bulk copies of %2d KB (x4 in flight, L2-resident source): %6.1f cycles/MMA, copy %.1f B/clk/SM %s\n", N, copy_kb, (double)c[0] / (reps * 12),
           (double)c[1] / (double)c[0], e == cudaSuccess ? "" : cudaGetErrorString(e));
  };
  for (int N : {64, 128, 256})
    for (int kb : {4, 16, 32}) runc(N, kb, 65536);
  return 0;
}
