// Micro-benchmark: tcgen05.mma issue rate for K-major vs MN-major shared-memory operands (bf16, M=128).
// One CTA per SM, operands are whatever is in shared memory (zeros); measures cycles per MMA.
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
__global__ void __launch_bounds__(128, 1) k(int N, int amn, int bmn, int reps, int shift, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t holder;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
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
  uint32_t tm = holder;
  if (threadIdx.x == 0) {
    uint32_t idesc = make_idesc(N, amn, bmn);
    uint32_t sa = smem_u32(smem), sb = sa + 32768;
    // K-major: 128 rows x 128 B, SBO 1024, K slice +32 B.  MN-major: atoms of 64 ch x 64 px (8 KB), LBO 8192, SBO 1024, K slice +2048 B
    uint64_t ad = amn ? make_sdesc(sa, 8192, 1024) : make_sdesc(sa, 16, 1024);
    uint64_t bd = bmn ? make_sdesc(sb + shift * 128, 9216, 1024) : make_sdesc(sb + shift * 128, 16, 1024);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint64_t a2 = ad + (uint64_t)(amn ? kk * 128 : kk * 2), b2 = bd + (uint64_t)(bmn ? kk * 128 : kk * 2);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tm), "l"(a2), "l"(b2), "r"(idesc), "r"(1));
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
    asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&bar)));
    long long t1 = clock64();
    if (blockIdx.x == 0) *out = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm));
}
int main() {
  long long* d;
  cudaMalloc(&d, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int reps = 2000;
  for (int N : {64, 128, 256})
    for (int mode = 0; mode < 4; ++mode)
      for (int shift : {0, 1}) {
        int amn = mode & 1, bmn = mode >> 1;
        for (int grid : {1, 148}) {
          k<<<grid, 128, 100 * 1024>>>(N, amn, bmn, reps, shift, d);
          cudaError_t e = cudaDeviceSynchronize();
          long long c = 0;
          cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
          printf("N=%3d A=%s B=%s shift=%d grid=%3d: %.1f cycles/MMA (nominal %d)%s\n", N, amn ? "MN" : "K ", bmn ? "MN" : "K ", shift, grid, (double)c / (reps * 4), N / 2,
                 e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
      }
  return 0;
}
