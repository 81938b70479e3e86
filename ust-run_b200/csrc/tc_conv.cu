// tcgen05 / TMEM / TMA implicit-GEMM convolution kernels for sm_100a (bf16 in, fp32 accumulate).
//
//   k_tc_conv  : forward-type GEMM  D[pixel][cout] = sum_{tap,ci} X[pixel+tap][ci] * W[cout][tap][ci]
//                (conv3x3 fwd, dgrad via flipped packing, conv1x1, the 4 GEMMs of ConvTranspose2d,
//                 ConvTranspose2d dgrad).  A (activations) arrives as 4-D TMA boxes over the NHWC
//                tensor -- one box of TH x TW pixels x 64 channels per (tap, channel chunk), halo and
//                padding produced by TMA out-of-bounds zero fill -- B (weights, K-major) as 2-D
//                boxes.  Both land in 128B-swizzled smem and feed tcgen05.mma (M=128, N=BN, K=16)
//                accumulating in TMEM.  Persistent CTAs, 4..8-stage mbarrier ring, double-buffered
//                accumulators; the epilogue warps read TMEM with tcgen05.ld, fuse bias and the
//                per-channel BatchNorm statistics (sum, sum of squares of the fp32 accumulators,
//                butterfly warp-shuffle reduction over the 32 pixel rows a warp owns), pack bf16
//                into a swizzled staging tile and TMA-store it.
//   k_tc_wgrad : weight gradient  D[co][ci] = sum_pixel dY[pixel][co] * X[pixel+tap][ci]; both
//                operands are MN-major views of the same NHWC TMA boxes (K = pixels), split-K over
//                CTAs with fp32 partials reduced by k_wgrad_reduce.
//
// Warp roles (256 threads): w0 TMA producer, w1 MMA issuer, w2 TMEM allocator, w4-7 epilogue.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace ustrun {

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
                   smem_u32(dst)),
               "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
               "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"((uint64_t)map), "r"(smem_u32(src)),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
      "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
        "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// elect.sync: exactly one lane of a converged warp returns true.  The compiler knows the guarded region runs
// on a single thread, so tcgen05/TMA operands move to uniform registers directly; guarding with
// `lane == 0` instead makes it wrap EVERY such instruction in an ELECT/BRA.U.ANY waterfall loop (~100
// cycles per MMA issue -- measured: N=64/128 tiles were issue-bound).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .b32 rx;\n\t"
      ".reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "@px mov.s32 %0, 1;\n\t"
      "}"
      : "+r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// ------------------------------------------------------------------------------------------
// descriptors (cute/arch/mma_sm100_desc.hpp bit layout)
// ------------------------------------------------------------------------------------------
// instruction descriptor: D=f32, A=B=bf16, M=128
__host__ __device__ constexpr uint32_t make_idesc(int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(128 >> 4) << 24);
}
// shared-memory matrix descriptor, SWIZZLE_128B, version 1 (Blackwell)
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t base_offset = 0) {
  uint64_t d = (uint64_t)(base_offset & 7) << 49;   // start address not 1024B-aligned: (addr >> 7) & 7
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// Column sums over the 32 lanes of a warp by a halving butterfly: level o (16, 8, .., 1) leaves o values per lane.
// butterfly_levels<HI, LO> runs levels HI, HI/2, .., LO on v[0 .. 2*HI); after level 1 lane j holds the total of
// column j.  All levels are linear, so the first levels can run per tile and the rest once per CTA on
// accumulated partials (see the epilogue: 128 accumulator registers for every BN).
template <int HI, int LO>
__device__ __forceinline__ void butterfly_levels(float* v, int lane) {
#pragma unroll
  for (int o = HI; o >= LO; o >>= 1) {
    const bool hi = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      float send = hi ? v[i] : v[i + o];
      float keep = hi ? v[i + o] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
}

// ------------------------------------------------------------------------------------------
// forward-type kernel
// ------------------------------------------------------------------------------------------
enum { TAP_CONV3 = 0, TAP_NONE = 1, TAP_PERMAP = 2 };

struct TcConvParams {
  int B, H, W;           // GEMM-row pixel grid
  int Cin, Cout;
  int ntaps, tap_mode;
  int TW, TH, tiles_w, tiles_h;
  int m_tiles, n_tiles, ctas_per_n;
  int halves, units;     // halves == 2: an M tile is two independent TW x TH (<= 64 pixel) boxes, rows 0-63 / 64-127;
                         // units = number of boxes (= m_tiles when halves == 1)
  float* partials;       // [ctas_per_n][2][Cout] or null
  const float* bias;     // [Cout] or null ([out_split] when out_split > 0)
  int out_split;         // > 0: the N dimension is `Cout / out_split` output tensors of out_split channels each (fused
                         // ConvTranspose2d: column block t goes to tensor map mapO<t>, bias index = column % out_split)
  int stages, nstaging;  // depth of the TMA ring / number of 16 KB store staging buffers
  uint32_t wres_bytes;   // RESB only: bytes of the resident weight block (ntaps * Cin/64 slots of BN x 128 B)
};

struct BoxCoord { int w0, h0, b; };
__device__ __forceinline__ BoxCoord box_coord(const TcConvParams& p, int unit) {
  BoxCoord c;
  const int r = unit / p.tiles_w;
  c.w0 = (unit - r * p.tiles_w) * p.TW;
  c.b = r / p.tiles_h;
  c.h0 = (r - c.b * p.tiles_h) * p.TH;
  return c;
}

constexpr int kEpiBar0 = 1, kEpiBar1 = 2;
constexpr uint32_t kStageA = 128 * 128;        // 128 pixel rows x 128 B

// ROW mode (3x3 conv, tile = 128 consecutive pixels of ONE image row): a stage holds one (dy, channel chunk)
// box of 136 pixels starting at w0-1 plus the weights of the three dx taps; the three dx taps are three
// UMMA descriptors into the SAME box, start address advanced by dx*128 B (one pixel row; the 128B swizzle
// is address-based, so the descriptor base-offset field stays 0 -- measured, tools/diag_row.py).  Activation traffic from L2 drops 3x (the operand that bounds the
// Cout=64/128 layers at the top of the UNet).
constexpr uint32_t kRowBoxPixels = 136;
template <int BN, bool ROW> struct FwdCfg {
  static constexpr uint32_t stageA = ROW ? kRowBoxPixels * 128 : kStageA;
  static constexpr uint32_t stageB = (ROW ? 3 : 1) * BN * 128;
  static constexpr uint32_t stage = stageA + stageB;
  static constexpr int stages = ROW ? (BN == 64 ? 4 : 3) : ((BN == 256) ? 4 : (BN == 128 ? 6 : 8));
  static constexpr int nstaging = (ROW && BN == 128) ? 1 : 2;
  static constexpr uint32_t smem = stages * stage + nstaging * 16384 + 1024 /*barriers*/ + 1024 /*align slack*/;
};

// RESB ("resident B"): when the whole weight block of the CTA's N tile fits in shared memory (<= ~147 KB: the
// 64/128-channel layers at the top of the UNet) it is loaded ONCE per CTA and the ring carries activations
// only.  Those layers were bound by re-loading the weights for every 128-pixel tile (L2->SM traffic and
// shared-memory fill bandwidth), not by the tensor pipe.
template <int BN, bool ROW, bool RESB>
__global__ void __launch_bounds__(384, 1)
k_tc_conv(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1, const __grid_constant__ CUtensorMap mapA2,
          const __grid_constant__ CUtensorMap mapA3, const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapO,
          const __grid_constant__ CUtensorMap mapO1, const __grid_constant__ CUtensorMap mapO2, const __grid_constant__ CUtensorMap mapO3,
          const TcConvParams p) {
  using Cfg = FwdCfg<BN, ROW>;
  const int S = p.stages;          // ring depth and store-staging buffers: chosen on the host (launch_fwd)
  const int nstg = p.nstaging;
  const uint32_t stage_stride = RESB ? Cfg::stageA : Cfg::stage;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem_al = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* wres = smem_al;                                       // RESB: [ntaps * kchunks][BN rows x 128 B]
  uint8_t* smem = smem_al + (RESB ? p.wres_bytes : 0u);          // the ring
  uint8_t* staging = smem + S * stage_stride;
  uint64_t* full_bar = (uint64_t*)(staging + nstg * 16384);
  uint64_t* empty_bar = full_bar + 8;
  uint64_t* tfull_bar = empty_bar + 8;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* wres_bar = tempty_bar + 2;
  uint32_t* tmem_holder = (uint32_t*)(wres_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tile = blockIdx.x % p.n_tiles;
  const int cta_j = blockIdx.x / p.n_tiles;
  const int n0 = n_tile * BN;
  const int kchunks = p.Cin >> 6;
  const int num_kb = (ROW ? 3 : p.ntaps) * kchunks;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapA0);
    prefetch_tmap(&mapW);
    prefetch_tmap(&mapO);
    for (int s = 0; s < S; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 4); }
    mbar_init(wres_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_holder, 2 * BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  // Register budget per role (setmaxnreg): the kernel is compiled for 168 registers per thread (__launch_bounds__(384, 1);
  // it is launched with 256 threads = 43 K registers per CTA instead of the whole 64 K file), the producer / MMA-issue /
  // TMEM-allocator warpgroup gives back all but 56, the four epilogue warps -- 128 BatchNorm-statistics accumulators plus a
  // 32-column TMEM slice per thread -- grow to 248.  The 22 K registers the CTA no longer holds are what lets a block of
  // the HBM-bound BatchNorm / pooling kernels of ANOTHER lane (multi-lane step) run on the same SM next to this CTA.
  // ptxas allocates per region only for code DOMINATED by the setmaxnreg instruction: the role code is nested under it.
  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == 0) {
    if (elect_one()) {
      if constexpr (RESB) {                    // the CTA's whole weight block, once
        mbar_expect_tx(wres_bar, p.wres_bytes);
        const int nslots = p.ntaps * kchunks;
        for (int i = 0; i < nslots; ++i) tma_load_2d(wres + (size_t)i * BN * 128, &mapW, wres_bar, i * 64, n0);
      }
      int stage = 0; uint32_t phase = 0;
      for (int mt = cta_j; mt < p.m_tiles; mt += p.ctas_per_n) {
        const BoxCoord c0 = box_coord(p, mt * p.halves);
        const int w0 = c0.w0, h0 = c0.h0, b = c0.b;
        if constexpr (ROW) {
          for (int dy = 0; dy < 3; ++dy)
            for (int kc = 0; kc < kchunks; ++kc) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              uint8_t* sa = smem + stage * stage_stride;
              mbar_expect_tx(&full_bar[stage], stage_stride);
              tma_load_4d(sa, &mapA1, &full_bar[stage], kc * 64, w0 - 1, h0 + dy - 1, b);      // mapA1: box of 136 pixels
              if constexpr (!RESB) {
#pragma unroll
                for (int dx = 0; dx < 3; ++dx)
                  tma_load_2d(sa + Cfg::stageA + dx * BN * 128, &mapW, &full_bar[stage], ((dy * 3 + dx) * kchunks + kc) * 64, n0);
              }
              if (++stage == S) { stage = 0; phase ^= 1; }
            }
        } else {
          // second 64-pixel box of the tile (rows 64..127 of the A stage); the last tile may have only one
          const bool two = p.halves == 2 && mt * 2 + 1 < p.units;
          const BoxCoord c1 = box_coord(p, two ? mt * 2 + 1 : 0);
          const uint32_t tx_bytes = (uint32_t)(p.TH * p.TW * 128) * (two ? 2u : 1u) + (RESB ? 0u : Cfg::stageB);
          for (int tap = 0; tap < p.ntaps; ++tap) {
            int dh = 0, dw = 0;
            const CUtensorMap* mA = &mapA0;
            if (p.tap_mode == TAP_CONV3) { dh = tap / 3 - 1; dw = tap % 3 - 1; }
            else if (p.tap_mode == TAP_PERMAP) mA = tap == 0 ? &mapA0 : (tap == 1 ? &mapA1 : (tap == 2 ? &mapA2 : &mapA3));
            for (int kc = 0; kc < kchunks; ++kc) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              uint8_t* sa = smem + stage * stage_stride;
              mbar_expect_tx(&full_bar[stage], tx_bytes);
              tma_load_4d(sa, mA, &full_bar[stage], kc * 64, w0 + dw, h0 + dh, b);
              if (two) tma_load_4d(sa + kStageA / 2, mA, &full_bar[stage], kc * 64, c1.w0 + dw, c1.h0 + dh, c1.b);
              if constexpr (!RESB) tma_load_2d(sa + Cfg::stageA, &mapW, &full_bar[stage], (tap * kchunks + kc) * 64, n0);
              if (++stage == S) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // One elected lane walks the pipeline and issues every MMA.  The per-stage work of this thread is the
    // critical path of the narrow tiles (a stage is only 4..12 MMAs of 48..128 cycles), so the 64-bit
    // shared-memory descriptors are never rebuilt: `a_cur` / `b_cur` advance by the stage stride (the 14-bit
    // address field cannot carry: addresses < 256 KB) and every MMA adds a compile-time constant.
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc(BN, 0, 0);
      const uint64_t a_first = make_sdesc(smem_u32(smem), 16, 1024);
      const uint64_t stride_d = (uint64_t)(stage_stride >> 4);
      const uint64_t w_first = make_sdesc(smem_u32(wres), 16, 1024);       // RESB: slot 0 of the resident weights
      constexpr uint64_t kSlotD = (uint64_t)(BN * 128) >> 4;               // one weight slot (BN rows x 128 B)
      constexpr uint64_t kBoffD = (uint64_t)Cfg::stageA >> 4;              // non-RESB: weights sit behind the A box of the stage
      uint64_t a_cur = a_first;
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      if constexpr (RESB) { mbar_wait(wres_bar, 0); tc_fence_after(); }
      for (int mt = cta_j; mt < p.m_tiles; mt += p.ctas_per_n, ++it) {
        const int buf = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[buf], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * BN;
        uint64_t w_cur = w_first;                                          // RESB: walks the slots in producer order
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if constexpr (ROW) {
            // stage = (dy, channel chunk); slots of its three dx taps: ((dy*3 + dx) * kchunks + kc)
            const uint64_t b0 = RESB ? w_cur : a_cur + kBoffD;
            const uint64_t bdx = RESB ? (uint64_t)kchunks * kSlotD : kSlotD;
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {          // three taps out of one box: start += dx pixel rows (128 B = 8 descriptor units)
#pragma unroll
              for (int k = 0; k < 4; ++k)
                tc_mma_bf16(d_tmem, a_cur + (uint64_t)(dx * 8 + k * 2), b0 + (uint64_t)dx * bdx + (uint64_t)(k * 2), idesc, (kb | dx | k) != 0);
            }
            if constexpr (RESB) {                     // next stage: kc+1, or (dy+1, kc=0) = +2*kchunks slots further
              if ((kb + 1) % kchunks == 0) w_cur += (uint64_t)(2 * kchunks + 1) * kSlotD;
              else w_cur += kSlotD;
            }
          } else {
            const uint64_t b0 = RESB ? w_cur : a_cur + kBoffD;
#pragma unroll
            for (int k = 0; k < 4; ++k)     // 64 channels = 4 x UMMA_K(16); +32 B inside the swizzle row
              tc_mma_bf16(d_tmem, a_cur + (uint64_t)(k * 2), b0 + (uint64_t)(k * 2), idesc, (kb | k) != 0);
            if constexpr (RESB) w_cur += kSlotD;
          }
          tc_commit(&empty_bar[stage]);
          if (kb == num_kb - 1) tc_commit(&tfull_bar[buf]);
          a_cur += stride_d;
          if (++stage == S) { stage = 0; phase ^= 1; a_cur = a_first; }
        }
      }
    }
    __syncwarp();
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 248;");
    const int q = warp - 4;                     // TMEM lane quadrant
    const int row = q * 32 + lane;              // pixel row inside the tile
    const int epi_tid = threadIdx.x - 128;
    // BatchNorm statistics: LV butterfly levels per tile, the remaining levels once per CTA.  Per lane
    // 2 * (BN/32) * (32 >> LV) = 128 running sums for BN = 64 / 128 / 256 (LV = 0 / 1 / 2).
    constexpr int LV = BN == 64 ? 0 : (BN == 128 ? 1 : 2);
    constexpr int NP = 32 >> LV;
    float asum[BN / 32][NP], asq[BN / 32][NP];
#pragma unroll
    for (int i = 0; i < BN / 32; ++i)
#pragma unroll
      for (int j = 0; j < NP; ++j) { asum[i][j] = 0.f; asq[i][j] = 0.f; }
    int it = 0;
    uint32_t chunk_ctr = 0;
    for (int mt = cta_j; mt < p.m_tiles; mt += p.ctas_per_n, ++it) {
      const int buf = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      // box this row belongs to (a warp's 32 rows always share one) and its pixel inside the box
      const int hsel = p.halves == 2 ? (row >> 6) : 0;
      const int within = p.halves == 2 ? (row & 63) : row;
      const int unit = mt * p.halves + hsel;
      const bool unit_ok = unit < p.units;
      const BoxCoord cr = box_coord(p, unit_ok ? unit : 0);
      const int th = within / p.TW, tw = within - th * p.TW;
      const bool valid = unit_ok && (within < p.TH * p.TW) && (cr.h0 + th < p.H) && (cr.w0 + tw < p.W);
      const bool tile_partial = !__all_sync(0xffffffffu, valid);         // warp-uniform: only warps with masked rows pay for it
      const BoxCoord s0 = box_coord(p, mt * p.halves);                     // store coordinates of the tile's box(es)
      const bool two = p.halves == 2 && mt * 2 + 1 < p.units;
      const BoxCoord s1 = box_coord(p, two ? mt * 2 + 1 : 0);
      mbar_wait(&tfull_bar[buf], acc_phase);
      tc_fence_after();
#pragma unroll
      for (int c64 = 0; c64 < BN / 64; ++c64, ++chunk_ctr) {
        uint8_t* stg = staging + (chunk_ctr % nstg) * 16384;
        if (q == 0 && elect_one()) {                                           // the store that last used this buffer has drained
          if (nstg == 2) tma_store_wait_read<1>();
          else tma_store_wait_read<0>();
        }
        named_bar_sync(kEpiBar0, 128);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int c32 = c64 * 2 + half;
          float v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + c32 * 32), v);
          uint32_t packed[16];
          const int bcol = p.out_split > 0 ? (n0 + c32 * 32) % p.out_split : n0 + c32 * 32;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float a = v[2 * i], c = v[2 * i + 1];
            if (p.bias) { a += p.bias[bcol + 2 * i]; c += p.bias[bcol + 2 * i + 1]; }
            __nv_bfloat162 h = __floats2bfloat162_rn(a, c);
            packed[i] = *reinterpret_cast<uint32_t*>(&h);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {                 // 4 x 16 B chunks of this half row, 128B swizzle
            const int chunk = half * 4 + j;
            uint4 val = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
            *reinterpret_cast<uint4*>(stg + row * 128 + ((chunk ^ (row & 7)) << 4)) = val;
          }
          if (p.partials) {
            if (tile_partial) {                       // warp-uniform: only edge tiles pay for the row mask
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = valid ? v[i] : 0.f;
            }
            float sq[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) sq[i] = v[i] * v[i];
            if constexpr (LV > 0) {
              butterfly_levels<16, (16 >> (LV - 1))>(v, lane);
              butterfly_levels<16, (16 >> (LV - 1))>(sq, lane);
            }
#pragma unroll
            for (int j = 0; j < NP; ++j) { asum[c32][j] += v[j]; asq[c32][j] += sq[j]; }
          }
        }
        fence_proxy_async();
        named_bar_sync(kEpiBar1, 128);
        if (q == 0) {                 // bulk-group state is per thread: elect.sync picks the same lane for the same mask every time
          if (elect_one()) {
            int ocol = n0 + c64 * 64;
            const CUtensorMap* mo = &mapO;
            if (p.out_split > 0) {                 // fused transposed conv: 64-channel block -> (tap, channel) of the 2x up-sampled grid
              const int t = ocol / p.out_split;
              ocol -= t * p.out_split;
              mo = t == 0 ? &mapO : (t == 1 ? &mapO1 : (t == 2 ? &mapO2 : &mapO3));
            }
            tma_store_4d(mo, stg, ocol, s0.w0, s0.h0, s0.b);
            if (two) tma_store_4d(mo, stg + kStageA / 2, ocol, s1.w0, s1.h0, s1.b);
            tma_store_commit();
          }
          __syncwarp();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[buf]);
    }
    if (q == 0 && elect_one()) tma_store_wait_all();
    if (p.partials) {
      named_bar_sync(kEpiBar0, 128);
      float* red = reinterpret_cast<float*>(staging);   // [4 warps][2][BN]
#pragma unroll
      for (int i = 0; i < BN / 32; ++i) {
        butterfly_levels<(16 >> LV), 1>(asum[i], lane);        // remaining levels: lane j ends with the total of column j
        butterfly_levels<(16 >> LV), 1>(asq[i], lane);
        red[(q * 2 + 0) * BN + i * 32 + lane] = asum[i][0];
        red[(q * 2 + 1) * BN + i * 32 + lane] = asq[i][0];
      }
      named_bar_sync(kEpiBar1, 128);
      for (int o = epi_tid; o < 2 * BN; o += 128) {
        const int which = o / BN, c = o - which * BN;
        float s = red[(0 * 2 + which) * BN + c] + red[(1 * 2 + which) * BN + c] + red[(2 * 2 + which) * BN + c] + red[(3 * 2 + which) * BN + c];
        p.partials[(size_t)cta_j * 2 * p.Cout + (size_t)which * p.Cout + n0 + c] = s;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 2 * BN);
}

// ------------------------------------------------------------------------------------------
// weight-gradient kernel
// ------------------------------------------------------------------------------------------
struct TcWgradParams {
  int B, H, W;             // pixel grid of the K loop
  int Mo, Nin;             // GEMM M (<= channels of the A-side tensor) and N channels per tap
  int ntaps, tap_mode;     // TAP_CONV3 shifts the N-side operand; TAP_PERMAP picks mapB[tap]
  int TW, TH, tiles_w, tiles_h, pix_tiles;
  int m_tiles, n_tiles, splits, tiles_per_split;
  float* ws;               // [splits][Mo][ntaps*Nin]
};

template <int BN> struct WgCfg {
  static constexpr uint32_t stageA = 2 * 64 * 128;            // two 64-channel atoms x 64 pixels
  static constexpr uint32_t stageB = (BN / 64) * 64 * 128;
  static constexpr uint32_t stage = stageA + stageB;
  static constexpr int stages = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr uint32_t smem = stages * stage + 1024 + 1024;
};

template <int BN>
__global__ void __launch_bounds__(256, 1)
k_tc_wgrad(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB0, const __grid_constant__ CUtensorMap mapB1,
           const __grid_constant__ CUtensorMap mapB2, const __grid_constant__ CUtensorMap mapB3, const TcWgradParams p) {
  using Cfg = WgCfg<BN>;
  constexpr int S = Cfg::stages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + S * Cfg::stage);
  uint64_t* empty_bar = full_bar + S;
  uint64_t* tfull_bar = empty_bar + S;
  uint32_t* tmem_holder = (uint32_t*)(tfull_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // work item: tap fastest so the taps of one (tile, split) run concurrently and share L2 lines
  int wi = blockIdx.x;
  const int tap = wi % p.ntaps; wi /= p.ntaps;
  const int n_tile = wi % p.n_tiles; wi /= p.n_tiles;
  const int m_tile = wi % p.m_tiles; wi /= p.m_tiles;
  const int split = wi;
  const int pt_begin = split * p.tiles_per_split;
  const int pt_end = min(pt_begin + p.tiles_per_split, p.pix_tiles);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapA);
    prefetch_tmap(&mapB0);
    for (int s = 0; s < S; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_holder, BN < 32 ? 32 : BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    if (elect_one()) {
      int dh = 0, dw = 0;
      const CUtensorMap* mB = &mapB0;
      if (p.tap_mode == TAP_CONV3) { dh = tap / 3 - 1; dw = tap % 3 - 1; }
      else if (p.tap_mode == TAP_PERMAP) mB = tap == 0 ? &mapB0 : (tap == 1 ? &mapB1 : (tap == 2 ? &mapB2 : &mapB3));
      int stage = 0; uint32_t phase = 0;
      for (int pt = pt_begin; pt < pt_end; ++pt) {
        const int txi = pt % p.tiles_w, tyi = (pt / p.tiles_w) % p.tiles_h, b = pt / (p.tiles_w * p.tiles_h);
        const int w0 = txi * p.TW, h0 = tyi * p.TH;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * Cfg::stage;
        mbar_expect_tx(&full_bar[stage], Cfg::stage);
        tma_load_4d(sa, &mapA, &full_bar[stage], m_tile * 128, w0, h0, b);
        tma_load_4d(sa + 8192, &mapA, &full_bar[stage], m_tile * 128 + 64, w0, h0, b);
#pragma unroll
        for (int j = 0; j < BN / 64; ++j)
          tma_load_4d(sa + Cfg::stageA + j * 8192, mB, &full_bar[stage], n_tile * BN + j * 64, w0 + dw, h0 + dh, b);
        if (++stage == S) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc(BN, 1, 1);
      // MN-major, 128B swizzle: LBO = stride between 64-channel atoms (8 KB), SBO = 8 pixel rows (1 KB)
      const uint64_t a_first = make_sdesc(smem_u32(smem), 8192, 1024);
      constexpr uint64_t kStrideD = (uint64_t)Cfg::stage >> 4, kBoffD = (uint64_t)Cfg::stageA >> 4;
      uint64_t a_cur = a_first;
      int stage = 0; uint32_t phase = 0;
      for (int pt = pt_begin; pt < pt_end; ++pt) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)       // 64 pixels = 4 x UMMA_K(16): advance 16 rows = 2 KB
          tc_mma_bf16(tmem_base, a_cur + (uint64_t)(k * 128), a_cur + kBoffD + (uint64_t)(k * 128), idesc, (pt > pt_begin) || (k > 0));
        tc_commit(&empty_bar[stage]);
        a_cur += kStrideD;
        if (++stage == S) { stage = 0; phase ^= 1; a_cur = a_first; }
      }
      tc_commit(tfull_bar);
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int q = warp - 4;
    const int m = m_tile * 128 + q * 32 + lane;
    const int N = p.ntaps * p.Nin;
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
#pragma unroll
    for (int c32 = 0; c32 < BN / 32; ++c32) {
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c32 * 32), v);
      const int c = n_tile * BN + c32 * 32;
      if (m < p.Mo && c < p.Nin && pt_end > pt_begin) {
        float* dst = p.ws + ((size_t)split * p.Mo + m) * N + (size_t)tap * p.Nin + c;
#pragma unroll
        for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(dst + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, BN < 32 ? 32 : BN);
}

// ------------------------------------------------------------------------------------------
// weight-gradient kernel, ROW mode (3x3):  D_dx[m][n] = sum_pixels TA[p][m] * TB[p + (dy-1, dx-1)][n]
// ------------------------------------------------------------------------------------------
// One CTA = (dy, 128 M-side channels, BN N-side channels, K split).  The K loop walks row segments of TW
// pixels; per segment ONE box of the unshifted M-side tensor and ONE box of TW+8 pixels of the N-side
// tensor (starting at w0-1 of row h+dy-1; TMA zero-fills the padding) are loaded, and the three dx taps are
// three UMMA descriptors into that same box, start address advanced by dx pixel rows (128 B) -- the 128B
// swizzle is a function of the shared-memory address, so a row-shifted MN-major view stays consistent with
// what TMA wrote.  Three accumulators (one per dx) live in TMEM for the whole CTA.  Compared with
// k_tc_wgrad (one tap per CTA) both operands cross L2->SM three times less often.
struct TcWgradRowParams {
  int B, H, W;
  int Mo, Nn;                 // valid channels on the M side / N side
  int TW, segs_w, nsegs;      // row segments
  int R, hgroups;             // image rows per K segment (R * TW <= 64: narrow feature maps put several rows in a stage)
  int nstages;                // ring depth (<= 8) and stage stride: sized on the host from TW and R
  uint32_t stage_bytes;
  int m_tiles, n_tiles, splits, segs_per_split;
  int flip;                   // 1: operands swapped (M side = x, N side = dy shifted by -(tap)): tap index 8 - t
  float* ws;                  // [splits][Mo][9 * Nn]
};
template <int BN> struct WgRowCfg {
  static constexpr uint32_t tmem_cols = BN == 128 ? 512 : 256;
  // a stage = two M-side atoms of R x TW pixels + BN/64 N-side atoms of R x (TW+8) pixels (128 B per pixel)
  static uint32_t stage_bytes(int TW, int R) { return (uint32_t)(2 * R * TW + (BN / 64) * R * (TW + 8)) * 128u; }
  static int stages(int TW, int R) {
    int n = (int)((232448u - 2048u) / stage_bytes(TW, R));
    return n > 8 ? 8 : n;
  }
  static uint32_t smem(int TW, int R) { return (uint32_t)stages(TW, R) * stage_bytes(TW, R) + 1024 + 1024; }
};

// MMA stream of k_tc_wgrad_row for KS K-slices (16 pixels each) per image row.  The issuing thread is the critical path
// when a stage holds only 6..12 MMAs of 48..64 cycles: with run-time trip counts nvcc emitted ~12 uniform-datapath
// instructions (R2UR, 64-bit adds, predicate logic) per MMA and the tensor pipe idled 40 % of the time (ncu:
// sm__pipe_tensor_cycles_active 61 %, the MMA warp never waiting on a full barrier).  Everything here is
// straight-line: constant descriptor offsets, accumulate flag folded except for the first slice of a row.
template <int BN, int KS>
__device__ __forceinline__ void wg_row_mma_loop(const TcWgradRowParams& p, uint32_t tmem_base, uint64_t* full_bar, uint64_t* empty_bar,
                                                uint64_t a_first, uint64_t b_first, int sg_begin, int sg_end) {
  constexpr uint32_t idesc = make_idesc(BN, 1, 1);
  constexpr uint32_t idesc3 = make_idesc(192, 1, 1);
  const uint64_t stride_d = (uint64_t)p.stage_bytes >> 4;
  const uint64_t rowA = (uint64_t)(p.TW * 8), rowB = (uint64_t)((p.TW + 8) * 8);      // one image row, in 16-byte units
  const int S = p.nstages, R = p.R;
  uint64_t a_cur = a_first, b_cur = b_first;
  int stage = 0; uint32_t phase = 0;
  uint32_t acc_first = 0;                               // 0 only for the very first slice of each accumulator
  for (int sg = sg_begin; sg < sg_end; ++sg) {
    mbar_wait(&full_bar[stage], phase);
    tc_fence_after();
    uint64_t ar = a_cur, br = b_cur;
    for (int r = 0; r < R; ++r) {                      // each image row of the stage is its own K run with its own halo
      if constexpr (BN == 64) {
#pragma unroll
        for (int k = 0; k < KS; ++k)                   // 16 pixel rows = 2 KB per K slice
          tc_mma_bf16(tmem_base, ar + (uint64_t)(k * 128), br + (uint64_t)(k * 128), idesc3, k == 0 ? acc_first : 1u);
      } else {
#pragma unroll
        for (int dx = 0; dx < 3; ++dx)
#pragma unroll
          for (int k = 0; k < KS; ++k)                 // dx shifts the N-side view by one pixel row (128 B)
            tc_mma_bf16(tmem_base + dx * BN, ar + (uint64_t)(k * 128), br + (uint64_t)(dx * 8 + k * 128), idesc, k == 0 ? acc_first : 1u);
      }
      acc_first = 1u;
      ar += rowA; br += rowB;
    }
    tc_commit(&empty_bar[stage]);
    a_cur += stride_d; b_cur += stride_d;
    if (++stage == S) { stage = 0; phase ^= 1; a_cur = a_first; b_cur = b_first; }
  }
}

template <int BN>
__global__ void __launch_bounds__(256, 1)
k_tc_wgrad_row(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const TcWgradRowParams p) {
  using Cfg = WgRowCfg<BN>;
  const int S = p.nstages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + (size_t)S * p.stage_bytes);
  uint64_t* empty_bar = full_bar + 8;
  uint64_t* tfull_bar = empty_bar + 8;
  uint32_t* tmem_holder = (uint32_t*)(tfull_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int wi = blockIdx.x;
  const int dy = wi % 3; wi /= 3;
  const int n_tile = wi % p.n_tiles; wi /= p.n_tiles;
  const int m_tile = wi % p.m_tiles; wi /= p.m_tiles;
  const int split = wi;
  const int sg_begin = split * p.segs_per_split;
  const int sg_end = min(sg_begin + p.segs_per_split, p.nsegs);
  const uint32_t atomA = (uint32_t)(p.R * p.TW) * 128, atomB = (uint32_t)(p.R * (p.TW + 8)) * 128;
  const uint32_t stageA = 2 * atomA;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapA);
    prefetch_tmap(&mapB);
    for (int s = 0; s < S; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_holder, Cfg::tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      const uint32_t tx = 2 * atomA + (BN / 64) * atomB;
      // segment -> (image, row group, row segment), walked incrementally: no integer division on the per-stage path
      int ws_ = sg_begin % p.segs_w, hg = (sg_begin / p.segs_w) % p.hgroups, b = sg_begin / (p.segs_w * p.hgroups);
      for (int sg = sg_begin; sg < sg_end; ++sg) {
        const int w0 = ws_ * p.TW, h = hg * p.R;          // boxes of R rows: rows past H are zero-filled on both sides
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + (size_t)stage * p.stage_bytes;
        mbar_expect_tx(&full_bar[stage], tx);
        tma_load_4d(sa, &mapA, &full_bar[stage], m_tile * 128, w0, h, b);
        tma_load_4d(sa + atomA, &mapA, &full_bar[stage], m_tile * 128 + 64, w0, h, b);
#pragma unroll
        for (int j = 0; j < BN / 64; ++j)
          tma_load_4d(sa + stageA + j * atomB, &mapB, &full_bar[stage], n_tile * BN + j * 64, w0 - 1, h + dy - 1, b);
        if (++stage == S) { stage = 0; phase ^= 1; }
        if (++ws_ == p.segs_w) { ws_ = 0; if (++hg == p.hgroups) { hg = 0; ++b; } }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {                  // one thread owns the whole MMA stream; descriptors advance incrementally
      const uint64_t a_first = make_sdesc(smem_u32(smem), atomA, 1024);
      // BN == 64: ONE N=192 MMA covers the three dx taps: the N-side "atoms" are the same 64-channel box at a stride
      // (LBO) of one pixel row, i.e. atom j IS the view shifted by j pixels.  The M-side operand is then read from
      // shared memory once instead of three times (an N=64 MMA is bound by that read: 48 cycles instead of 32).
      const uint64_t b_first = make_sdesc(smem_u32(smem) + stageA, BN == 64 ? 128u : atomB, 1024);
      switch (p.TW >> 4) {              // K slices per image row: compile-time, so the MMAs of a row are straight-line code
        case 1: wg_row_mma_loop<BN, 1>(p, tmem_base, full_bar, empty_bar, a_first, b_first, sg_begin, sg_end); break;
        case 2: wg_row_mma_loop<BN, 2>(p, tmem_base, full_bar, empty_bar, a_first, b_first, sg_begin, sg_end); break;
        case 3: wg_row_mma_loop<BN, 3>(p, tmem_base, full_bar, empty_bar, a_first, b_first, sg_begin, sg_end); break;
        default: wg_row_mma_loop<BN, 4>(p, tmem_base, full_bar, empty_bar, a_first, b_first, sg_begin, sg_end); break;
      }
      tc_commit(tfull_bar);
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int q = warp - 4;
    const int m = m_tile * 128 + q * 32 + lane;
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      const int t = p.flip ? 8 - (dy * 3 + dx) : dy * 3 + dx;
#pragma unroll
      for (int c32 = 0; c32 < BN / 32; ++c32) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(dx * BN + c32 * 32), v);
        const int c = n_tile * BN + c32 * 32;
        if (m < p.Mo && c < p.Nn && sg_end > sg_begin) {
          float* dst = p.ws + (((size_t)split * p.Mo + m) * 9 + t) * p.Nn + c;
#pragma unroll
          for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(dst + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::tmem_cols);
}

// ------------------------------------------------------------------------------------------
// host side: tensor maps + launches
// ------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)ptr;
  }
  return fn;
}

// 4-D NHWC bf16 view: dims (C, W, H, B) with element strides (1, sw, sh, sb); box (64, TW, TH, 1)
static int make_act_map(CUtensorMap* m, const void* base, int C, int W, int H, int B, long long sw, long long sh, long long sb, int TW, int TH) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return USTRUN_ERR_ARG; }
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t gstr[3] = {(cuuint64_t)sw * 2, (cuuint64_t)sh * 2, (cuuint64_t)sb * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)TW, (cuuint32_t)TH, 1};
  cuuint32_t est[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(act C=%d W=%d H=%d B=%d box %dx%d) failed: %d", C, W, H, B, TW, TH, (int)r); return USTRUN_ERR_ARG; }
  return 0;
}
static int make_w_map(CUtensorMap* m, const void* base, long long K, int rows, int box_rows) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return USTRUN_ERR_ARG; }
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t est[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(weights K=%lld rows=%d) failed: %d", K, rows, (int)r); return USTRUN_ERR_ARG; }
  return 0;
}

// pick the TH x TW pixel box (TH*TW <= target) covering an H x W image with the least padding
static void pick_tile(int H, int W, int target, bool exact_pow2, int& TW, int& TH) {
  double best = -1.0;
  TW = 1; TH = 1;
  for (int tw = 1; tw <= W && tw <= target && tw <= 256; ++tw) {
    if (exact_pow2 && (tw & (tw - 1))) continue;
    int th = target / tw;
    if (th > 256) th = 256;
    if (th < 1) continue;
    if (exact_pow2 && tw * th != target) continue;
    if (!exact_pow2 && th > H) th = H;
    long long cover = (long long)((W + tw - 1) / tw) * ((H + th - 1) / th);
    double eff = (double)H * W / ((double)cover * target);
    if (eff > best + 1e-9 || (eff > best - 1e-9 && tw > TW)) { best = eff; TW = tw; TH = th; }
  }
}

static int row_mode() {      // USTRUN_TC_ROW=0 disables the row-reuse variant (A/B comparisons)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("USTRUN_TC_ROW");
    v = e ? atoi(e) : 1;
  }
  return v;
}
static int g_num_sms = 0;
static int num_sms() {
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

struct ActView {            // NHWC bf16 view in elements
  const void* base;
  int C, W, H, B;
  long long sw, sh, sb;
};

static int resb_mode() {      // USTRUN_TC_RESB=0 disables the resident-weights variant (A/B comparisons)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("USTRUN_TC_RESB");
    v = e ? atoi(e) : 1;
  }
  return v;
}
constexpr uint32_t kMaxDynSmem = 232448;       // 227 KB opt-in limit per CTA on sm_100

static int ctas_per_n_tile(int n_tiles, int m_tiles) {
  int per = num_sms() / n_tiles;
  if (per < 1) per = 1;
  if (per > m_tiles) per = m_tiles;
  if (per > USTRUN_MAX_PARTS) per = USTRUN_MAX_PARTS;
  return per;
}

// Tiling of the GEMM-row pixel grid and the N tile width.  Default: one TW x TH box of <= 128 pixels per M tile and
// the widest N tile that divides Cout.  Two things cost whole rounds of the persistent CTAs on small feature maps
// (the 24 x 24 bottleneck of a 384 x 384 input: 24x5 boxes, the fifth box of an image 80 % full, 40 tiles on the
// 37 CTAs of an N tile = two rounds for 1.08 rounds of work): (a) badly filled boxes -- an M tile can instead be
// built from two independent 64-pixel boxes (8x8 covers 24x24 exactly: 36 full tiles, one round); (b) too few tiles
// per N tile -- a narrower N tile doubles the CTAs that have work.  Cost model: rounds x relative time of one tile
// (N=128 tiles run at ~1000 vs ~1350 TFLOP/s for N=256, profiles/r01_kernel_microbench_cfg2.txt).
static int tile_plan_mode() {      // USTRUN_TC_TILEPLAN=0: always one box per tile, widest N tile (A/B comparisons)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("USTRUN_TC_TILEPLAN");
    v = e ? atoi(e) : 1;
  }
  return v;
}
static int plan_fwd_tiles(TcConvParams& p, int B, int H, int W, int Cout) {
  const int bn_default = Cout % 256 == 0 ? 256 : (Cout % 128 == 0 ? 128 : 64);
  int tw1, th1, tw2, th2;
  pick_tile(H, W, 128, false, tw1, th1);
  pick_tile(H, W, 64, false, tw2, th2);
  const int units1 = B * ((W + tw1 - 1) / tw1) * ((H + th1 - 1) / th1);
  const int units2 = B * ((W + tw2 - 1) / tw2) * ((H + th2 - 1) / th2);
  int best_bn = bn_default, best_halves = 1;
  double best = 1e30;
  for (int bn = bn_default; bn >= (bn_default == 256 ? 128 : bn_default); bn >>= 1) {
    const double tile_cost = bn == 256 ? 1.0 : (bn == 128 ? 0.66 : 0.45);
    for (int halves = 1; halves <= 2; ++halves) {
      const int m_tiles = halves == 1 ? units1 : (units2 + 1) / 2;
      const int per = ctas_per_n_tile(Cout / bn, m_tiles);
      const double cost = (double)((m_tiles + per - 1) / per) * tile_cost;
      if (cost < best * 0.97) { best = cost; best_bn = bn; best_halves = halves; }      // earlier candidates win near-ties
    }
    if (tile_plan_mode() <= 0) break;
  }
  if (tile_plan_mode() <= 0) { best_bn = bn_default; best_halves = 1; }
  p.halves = best_halves;
  p.TW = best_halves == 1 ? tw1 : tw2;
  p.TH = best_halves == 1 ? th1 : th2;
  p.tiles_w = (W + p.TW - 1) / p.TW; p.tiles_h = (H + p.TH - 1) / p.TH;
  p.units = B * p.tiles_w * p.tiles_h;
  p.m_tiles = (p.units + best_halves - 1) / best_halves;
  return best_bn;
}

static const ActView* g_extra_outs = nullptr;      // fused transposed conv: output tensors of taps 1..3 (set around the launch by tc_convT_fwd)

template <int BN, bool ROW, bool RESB>
static int launch_fwd_impl(const ActView* a, int nmaps, const void* w, long long Ktot, const ActView& out, TcConvParams p, uint32_t smem_bytes,
                           cudaStream_t st) {
  CUtensorMap mA[4], mW, mO, mOx[3];
  for (int i = 0; i < 4; ++i) {
    const ActView& v = a[i < nmaps ? i : 0];
    int rc = make_act_map(&mA[i], v.base, v.C, v.W, v.H, v.B, v.sw, v.sh, v.sb, (ROW && i == 1) ? (int)kRowBoxPixels : p.TW, p.TH);
    if (rc) return rc;
  }
  int rc = make_w_map(&mW, w, Ktot, p.Cout, BN);
  if (rc) return rc;
  rc = make_act_map(&mO, out.base, out.C, out.W, out.H, out.B, out.sw, out.sh, out.sb, p.TW, p.TH);
  if (rc) return rc;
  for (int i = 0; i < 3; ++i) {
    const ActView& o = (p.out_split > 0 && g_extra_outs) ? g_extra_outs[i] : out;
    rc = (p.out_split > 0 && g_extra_outs) ? make_act_map(&mOx[i], o.base, o.C, o.W, o.H, o.B, o.sw, o.sh, o.sb, p.TW, p.TH) : 0;
    if (rc) return rc;
    if (!(p.out_split > 0 && g_extra_outs)) mOx[i] = mO;
  }
  static bool attr_set = false;
  if (!attr_set) {
    const uint32_t want = RESB ? kMaxDynSmem : smem_bytes;
    cudaError_t e = cudaFuncSetAttribute(k_tc_conv<BN, ROW, RESB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)want);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(k_tc_conv<%d,%d,%d>, %u): %s", BN, (int)ROW, (int)RESB, want, cudaGetErrorString(e)); return (int)e; }
    attr_set = true;
  }
  k_tc_conv<BN, ROW, RESB><<<p.n_tiles * p.ctas_per_n, 256, smem_bytes, st>>>(mA[0], mA[1], mA[2], mA[3], mW, mO, mOx[0], mOx[1], mOx[2], p);
  return check_launch("k_tc_conv");
}

// USTRUN_TC_COSHARE=1: leave >= 16 KB of the SM's shared memory (and, always, via setmaxnreg 22 K registers) unused by a conv
// CTA, so that a block of an HBM-bound BatchNorm / pooling kernel running on another stream could be resident next to it.
// Measured (tools/coreside_probe.py): no co-residency gain on B200 -- conv + BN launched on two streams take ~0.85 x the sum
// of the two either way (tail overlap only) -- and the shallower ring costs the conv ~2 %, so it is OFF by default.
static int coshare_mode() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("USTRUN_TC_COSHARE");
    v = e ? atoi(e) : 0;
  }
  return v;
}
constexpr uint32_t kCoshareFree = 17 * 1024;

template <int BN, bool ROW = false>
static int launch_fwd(const ActView* a, int nmaps, const void* w, long long Ktot, const ActView& out, TcConvParams p, cudaStream_t st) {
  using Cfg = FwdCfg<BN, ROW>;
  p.n_tiles = p.Cout / BN;
  const int per = ctas_per_n_tile(p.n_tiles, p.m_tiles);
  p.ctas_per_n = per;
  if (p.halves < 1) { p.halves = 1; p.units = p.m_tiles; }
  if constexpr (BN <= 128) {
    // resident weights: the N tile's whole [BN][Ktot] block stays in shared memory, the ring carries activations only
    const long long wres = Ktot * BN * 2;
    const int tiles_per_cta = (p.m_tiles + per - 1) / per;
    // measured on B200 (profiles/r01_kernel_microbench_cfg2.txt): pays off for the row-mode N=128 tiles (3 x 16 KB of
    // weights per stage otherwise); for N=64 tiles the weight slots are small and the plain ring is as fast
    if ((resb_mode() > 1 || (resb_mode() == 1 && ROW && BN == 128)) && wres <= 148 * 1024 && tiles_per_cta >= 3) {
      const long long avail = (long long)kMaxDynSmem - 2048 - wres - (coshare_mode() > 0 ? kCoshareFree : 0);
      int nstg = 2;
      long long stages = (avail - nstg * 16384) / Cfg::stageA;
      if (stages < 4) { nstg = 1; stages = (avail - nstg * 16384) / Cfg::stageA; }
      if (stages > 8) stages = 8;
      { static int cap = -1; if (cap < 0) { const char* e = getenv("USTRUN_TC_RESB_STAGES"); cap = e ? atoi(e) : 8; } if (stages > cap) stages = cap; }
      if (stages >= 3) {
        p.stages = (int)stages; p.nstaging = nstg; p.wres_bytes = (uint32_t)wres;
        const uint32_t smem_bytes = (uint32_t)(wres + stages * Cfg::stageA + nstg * 16384 + 2048);
        return launch_fwd_impl<BN, ROW, true>(a, nmaps, w, Ktot, out, p, smem_bytes, st);
      }
    }
  }
  p.stages = Cfg::stages; p.nstaging = Cfg::nstaging;
  uint32_t smem_bytes = Cfg::smem;
  if (coshare_mode() > 0 && smem_bytes + kCoshareFree > kMaxDynSmem) {
    // BN = 256 keeps its 4 ring stages (48 KB each) and gives up one store-staging buffer; narrower tiles give up one stage
    if (BN == 256 && p.nstaging > 1) { p.nstaging -= 1; smem_bytes -= 16384; }
    else { p.stages -= 1; smem_bytes -= Cfg::stage; }
  }
  return launch_fwd_impl<BN, ROW, false>(a, nmaps, w, Ktot, out, p, smem_bytes, st);
}

// x: [B,H,W,Cin] (ldx), y: [B,H,W,Cout] (ldy); generic entry used by conv3x3/1x1 fwd+dgrad
int tc_conv_fwd(const void* x, int ldx, const void* w, const float* bias, void* y, int ldy, int B, int H, int W, int Cin, int Cout, int ksize,
                float* partials, int* nparts_host, cudaStream_t st) {
  if (Cin % 64 || Cout % 64 || ldx % 8 || ldy % 8) { set_error("tcgen05 conv needs Cin, Cout %% 64 == 0 (got %d, %d)", Cin, Cout); return USTRUN_ERR_ARG; }
  TcConvParams p{};
  p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  p.ntaps = ksize * ksize; p.tap_mode = ksize == 3 ? TAP_CONV3 : TAP_NONE;
  int BN = Cout % 256 == 0 ? 256 : (Cout % 128 == 0 ? 128 : 64);
  // row mode: 3x3, N tile <= 128 (the L2-bound layers at the top of the UNet), rows that split into whole
  // 128-pixel segments (measured on B200: 64->64 @384 222 -> 155 us; with a 75 %-filled last segment it loses)
  const bool row = row_mode() > 0 && ksize == 3 && BN <= 128 && W % 128 == 0;
  if (row) {
    p.TW = 128; p.TH = 1;
    p.tiles_w = (W + p.TW - 1) / p.TW; p.tiles_h = (H + p.TH - 1) / p.TH;
    p.m_tiles = B * p.tiles_w * p.tiles_h;
    p.halves = 1; p.units = p.m_tiles;
  } else {
    BN = plan_fwd_tiles(p, B, H, W, Cout);
  }
  p.partials = partials; p.bias = bias;
  ActView a{x, Cin, W, H, B, ldx, (long long)W * ldx, (long long)H * W * ldx};
  ActView o{y, Cout, W, H, B, ldy, (long long)W * ldy, (long long)H * W * ldy};
  if (nparts_host) *nparts_host = ctas_per_n_tile(Cout / BN, p.m_tiles);
  long long Ktot = (long long)p.ntaps * Cin;
  if (row) {
    ActView aa[2] = {a, a};
    if (BN == 128) return launch_fwd<128, true>(aa, 2, w, Ktot, o, p, st);
    return launch_fwd<64, true>(aa, 2, w, Ktot, o, p, st);
  }
  if (BN == 256) return launch_fwd<256>(&a, 1, w, Ktot, o, p, st);
  if (BN == 128) return launch_fwd<128>(&a, 1, w, Ktot, o, p, st);
  return launch_fwd<64>(&a, 1, w, Ktot, o, p, st);
}

// ConvTranspose2d(k2,s2) forward as ONE GEMM: the packed weights wf[4][Cout][Cin] are a K-major [4*Cout][Cin] matrix, so the
// four taps are column blocks of N = 4*Cout; the input tile is read once (not four times) and every 64-channel block of the
// epilogue is TMA-stored into its tap's (2h+i, 2w+j) sub-grid of the output (four output tensor maps).  USTRUN_TC_CONVT_FUSED=0
// restores the four separate launches (A/B comparisons).
static int convt_fused_mode() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("USTRUN_TC_CONVT_FUSED");
    v = e ? atoi(e) : 1;
  }
  return v;
}
int tc_convT_fwd(const void* x, int ldx, const void* wf, const float* bias, void* y, int ldy, int B, int H, int W, int Cin, int Cout,
                 cudaStream_t st) {
  if (Cin % 64 || Cout % 64 || ldx % 8 || ldy % 8) { set_error("tcgen05 convT needs Cin, Cout %% 64 == 0"); return USTRUN_ERR_ARG; }
  ActView a{x, Cin, W, H, B, ldx, (long long)W * ldx, (long long)H * W * ldx};
  ActView o[4];
  for (int ij = 0; ij < 4; ++ij) {
    const char* ybase = (const char*)y + ((long long)(ij >> 1) * 2 * W + (ij & 1)) * ldy * 2;
    o[ij] = ActView{ybase, Cout, W, H, B, 2LL * ldy, 4LL * W * ldy, 4LL * H * W * ldy};
  }
  if (convt_fused_mode() > 0) {
    TcConvParams p{};
    p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = 4 * Cout; p.ntaps = 1; p.tap_mode = TAP_NONE; p.out_split = Cout;
    const int BN = plan_fwd_tiles(p, B, H, W, 4 * Cout);
    p.partials = nullptr; p.bias = bias;
    g_extra_outs = &o[1];
    int rc = BN == 256 ? launch_fwd<256>(&a, 1, wf, Cin, o[0], p, st) : (BN == 128 ? launch_fwd<128>(&a, 1, wf, Cin, o[0], p, st) : launch_fwd<64>(&a, 1, wf, Cin, o[0], p, st));
    g_extra_outs = nullptr;
    return rc;
  }
  for (int ij = 0; ij < 4; ++ij) {
    TcConvParams p{};
    p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.ntaps = 1; p.tap_mode = TAP_NONE;
    const int BN = plan_fwd_tiles(p, B, H, W, Cout);
    p.partials = nullptr; p.bias = bias;
    const char* wij = (const char*)wf + (size_t)ij * Cout * Cin * 2;
    int rc = BN == 256 ? launch_fwd<256>(&a, 1, wij, Cin, o[ij], p, st) : (BN == 128 ? launch_fwd<128>(&a, 1, wij, Cin, o[ij], p, st) : launch_fwd<64>(&a, 1, wij, Cin, o[ij], p, st));
    if (rc) return rc;
  }
  return 0;
}

// ConvTranspose2d dgrad: dx[b,h,w,ci] = sum_{ij,co} dy[b,2h+i,2w+j,co] * wd[ci][ij][co]
int tc_convT_dgrad(const void* dy, int lddy, const void* wd, void* dx, int lddx, int B, int H, int W, int Cin, int Cout, cudaStream_t st) {
  if (Cin % 64 || Cout % 64 || lddy % 8 || lddx % 8) { set_error("tcgen05 convT dgrad needs Cin, Cout %% 64 == 0"); return USTRUN_ERR_ARG; }
  TcConvParams p{};
  p.B = B; p.H = H; p.W = W; p.Cin = Cout; p.Cout = Cin; p.ntaps = 4; p.tap_mode = TAP_PERMAP;
  const int BN = plan_fwd_tiles(p, B, H, W, Cin);
  ActView a[4];
  for (int ij = 0; ij < 4; ++ij) {
    const char* base = (const char*)dy + ((long long)(ij >> 1) * 2 * W + (ij & 1)) * lddy * 2;
    a[ij] = ActView{base, Cout, W, H, B, 2LL * lddy, 4LL * W * lddy, 4LL * H * W * lddy};
  }
  ActView o{dx, Cin, W, H, B, lddx, (long long)W * lddx, (long long)H * W * lddx};
  long long Ktot = 4LL * Cout;
  if (BN == 256) return launch_fwd<256>(a, 4, wd, Ktot, o, p, st);
  if (BN == 128) return launch_fwd<128>(a, 4, wd, Ktot, o, p, st);
  return launch_fwd<64>(a, 4, wd, Ktot, o, p, st);
}

int launch_wgrad_reduce(const float* ws, int splits, int Mo, int Nin, int taps, float* dw, int accumulate, cudaStream_t st, int swapped);

struct WgPlan { int TW, TH, tiles_w, tiles_h, pix_tiles, m_tiles, n_tiles, BN, splits, tps; };
static WgPlan plan_wgrad(int B, int H, int W, int Mo, int Nin, int ntaps) {
  WgPlan g;
  pick_tile(H, W, 64, true, g.TW, g.TH);
  g.tiles_w = (W + g.TW - 1) / g.TW; g.tiles_h = (H + g.TH - 1) / g.TH;
  g.pix_tiles = B * g.tiles_w * g.tiles_h;
  g.BN = Nin % 256 == 0 ? 256 : (Nin % 128 == 0 ? 128 : 64);
  g.m_tiles = (Mo + 127) / 128; g.n_tiles = Nin / g.BN;
  long long items = (long long)ntaps * g.m_tiles * g.n_tiles;
  long long want = (2LL * num_sms() + items - 1) / items;       // ~2 waves of CTAs
  if (want < 1) want = 1;
  long long maxs = (g.pix_tiles + 7) / 8;                       // >= 8 pixel tiles (512 pixels) per split
  if (maxs < 1) maxs = 1;
  if (want > maxs) want = maxs;
  g.tps = (int)((g.pix_tiles + want - 1) / want);
  g.splits = (g.pix_tiles + g.tps - 1) / g.tps;
  return g;
}
long long tc_wgrad_ws_bytes(int B, int H, int W, int Mo, int Nin, int ntaps) {
  WgPlan g = plan_wgrad(B, H, W, Mo, Nin, ntaps);
  return (long long)g.splits * Mo * ntaps * Nin * (long long)sizeof(float);
}

template <int BN>
static int launch_wgrad(const ActView& a, const ActView* bviews, int nb, TcWgradParams p, cudaStream_t st) {
  using Cfg = WgCfg<BN>;
  CUtensorMap mA, mB[4];
  int rc = make_act_map(&mA, a.base, a.C, a.W, a.H, a.B, a.sw, a.sh, a.sb, p.TW, p.TH);
  if (rc) return rc;
  for (int i = 0; i < 4; ++i) {
    const ActView& v = bviews[i < nb ? i : 0];
    rc = make_act_map(&mB[i], v.base, v.C, v.W, v.H, v.B, v.sw, v.sh, v.sb, p.TW, p.TH);
    if (rc) return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_tc_wgrad<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smem);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(k_tc_wgrad<%d>): %s", BN, cudaGetErrorString(e)); return (int)e; }
    attr_set = true;
  }
  int grid = p.ntaps * p.m_tiles * p.n_tiles * p.splits;
  k_tc_wgrad<BN><<<grid, 256, Cfg::smem, st>>>(mA, mB[0], mB[1], mB[2], mB[3], p);
  return check_launch("k_tc_wgrad");
}

// generic: A-side tensor (M channels, unshifted), B-side tensor(s) (N channels per tap)
static int tc_wgrad_generic(const ActView& a, const ActView* bviews, int nb, int tap_mode, int ntaps, int B, int H, int W, int Mo, int Nin,
                            float* dw, int accumulate, void* workspace, long long ws_bytes, cudaStream_t st) {
  WgPlan g = plan_wgrad(B, H, W, Mo, Nin, ntaps);
  long long need = (long long)g.splits * Mo * ntaps * Nin * (long long)sizeof(float);
  if (!workspace || ws_bytes < need) { set_error("tc wgrad: workspace too small (%lld < %lld)", ws_bytes, need); return USTRUN_ERR_ARG; }
  TcWgradParams p{};
  p.B = B; p.H = H; p.W = W; p.Mo = Mo; p.Nin = Nin; p.ntaps = ntaps; p.tap_mode = tap_mode;
  p.TW = g.TW; p.TH = g.TH; p.tiles_w = g.tiles_w; p.tiles_h = g.tiles_h; p.pix_tiles = g.pix_tiles;
  p.m_tiles = g.m_tiles; p.n_tiles = g.n_tiles; p.splits = g.splits; p.tiles_per_split = g.tps;
  p.ws = (float*)workspace;
  int rc = g.BN == 256 ? launch_wgrad<256>(a, bviews, nb, p, st) : (g.BN == 128 ? launch_wgrad<128>(a, bviews, nb, p, st) : launch_wgrad<64>(a, bviews, nb, p, st));
  if (rc) return rc;
  return launch_wgrad_reduce((const float*)workspace, g.splits, Mo, Nin, ntaps, dw, accumulate, st, 0);
}

// ---- row-mode plan (3x3) --------------------------------------------------------------------
static int wgrad_row_mode() {      // USTRUN_TC_WGRAD_ROW=0 falls back to the one-tap-per-CTA kernel (A/B comparisons)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("USTRUN_TC_WGRAD_ROW");
    v = e ? atoi(e) : 1;
  }
  return v;
}
struct WgRowPlan { bool ok; int swap, Mo, Nn, BN, TW, R, hgroups, segs_w, nsegs, m_tiles, n_tiles, splits, sps; };
static WgRowPlan plan_wgrad_row(int B, int H, int W, int Cin, int Cout, int ksize) {
  WgRowPlan g{};
  g.ok = false;
  if (ksize != 3 || wgrad_row_mode() <= 0) return g;
  int TW = 0;
  for (int tw : {64, 48, 32})
    if (W % tw == 0) { TW = tw; break; }
  if (!TW && W < 64 && W >= 16) TW = (W + 15) / 16 * 16;           // one zero-padded segment per row
  if (!TW) return g;
  g.TW = TW;
  g.segs_w = (W + TW - 1) / TW;
  // narrow feature maps (W <= 32): a 16/32-pixel segment is only 1-2 K slices per tap, and the pipeline is then bound by
  // TMA/barrier latency, not the tensor pipe -- put R image rows (R * TW <= 64 pixels) into one stage
  g.R = (tile_plan_mode() > 0 && TW <= 32) ? 64 / TW : 1;
  if (g.R > H) g.R = H;
  g.hgroups = (H + g.R - 1) / g.R;
  g.nsegs = B * g.hgroups * g.segs_w;
  // M side = the unshifted tensor.  Default: dy (Cout rows).  With Cout == 64 < Cin the operands swap so that
  // no half of the 128-row MMA is idle: M side = x (Cin rows), N side = dy shifted by -(tap).
  g.swap = (Cout == 64 && Cin >= 128) ? 1 : 0;
  g.Mo = g.swap ? Cin : Cout;
  g.Nn = g.swap ? Cout : Cin;
  g.BN = g.Nn % 128 == 0 ? 128 : 64;
  g.m_tiles = (g.Mo + 127) / 128;
  g.n_tiles = g.Nn / g.BN;
  const long long items = 3LL * g.m_tiles * g.n_tiles;
  long long want = num_sms() / items;                                // one wave of CTAs (1 CTA / SM)
  if (want < 1) want = 1;
  long long maxs = (g.nsegs + 15) / 16;                              // >= 16 segments per split
  if (maxs < 1) maxs = 1;
  if (want > maxs) want = maxs;
  if (tile_plan_mode() > 0 && (items * want > num_sms() || items * want * 5 < (long long)num_sms() * 4)) {
    // 96 or 192 work items on 148 SMs (the 512/1024-channel layers) leave a third of the SMs idle or cost a second,
    // mostly empty wave: more K splits even the waves out.  Cost = waves x (K time per split) + the extra passes of
    // the split-K reduction over the fp32 partials (~5 TB/s).
    const double t_item = (double)g.nsegs * g.R * 3.0 * (TW / 16) * (g.BN == 128 ? 64.0 : 48.0) * 1.3 / 1900.0;     // us
    const double t_red = (double)g.Mo * 9.0 * g.Nn * 4.0 * 2.0 / 5.0e6;                                       // us per split
    double best = 1e30;
    long long best_s = want;
    for (long long sct = want; sct <= maxs && sct <= want * 4; ++sct) {
      const long long waves = (items * sct + num_sms() - 1) / num_sms();
      const double cost = (double)waves * t_item / (double)sct + t_red * (double)sct;
      if (cost < best * 0.97) { best = cost; best_s = sct; }
    }
    want = best_s;
  }
  g.sps = (int)((g.nsegs + want - 1) / want);
  g.splits = (g.nsegs + g.sps - 1) / g.sps;
  g.ok = true;
  return g;
}

template <int BN>
static int launch_wgrad_row(const ActView& a, const ActView& b, const WgRowPlan& g, TcWgradRowParams p, cudaStream_t st) {
  using Cfg = WgRowCfg<BN>;
  CUtensorMap mA, mB;
  int rc = make_act_map(&mA, a.base, a.C, a.W, a.H, a.B, a.sw, a.sh, a.sb, g.TW, g.R);
  if (rc) return rc;
  rc = make_act_map(&mB, b.base, b.C, b.W, b.H, b.B, b.sw, b.sh, b.sb, g.TW + 8, g.R);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_tc_wgrad_row<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(k_tc_wgrad_row<%d>): %s", BN, cudaGetErrorString(e)); return (int)e; }
    attr_set = true;
  }
  p.nstages = Cfg::stages(g.TW, g.R);
  p.stage_bytes = Cfg::stage_bytes(g.TW, g.R);
  const int grid = 3 * g.m_tiles * g.n_tiles * g.splits;
  k_tc_wgrad_row<BN><<<grid, 256, Cfg::smem(g.TW, g.R), st>>>(mA, mB, p);
  return check_launch("k_tc_wgrad_row");
}

int tc_conv_wgrad(const void* dy, int lddy, const void* x, int ldx, float* dw, int accumulate, int B, int H, int W, int Cin, int Cout, int ksize,
                  void* workspace, long long ws_bytes, cudaStream_t st) {
  if (Cin % 64 || Cout % 64 || lddy % 8 || ldx % 8) { set_error("tcgen05 wgrad needs Cin, Cout %% 64 == 0"); return USTRUN_ERR_ARG; }
  ActView a{dy, Cout, W, H, B, lddy, (long long)W * lddy, (long long)H * W * lddy};
  ActView b{x, Cin, W, H, B, ldx, (long long)W * ldx, (long long)H * W * ldx};
  WgRowPlan g = plan_wgrad_row(B, H, W, Cin, Cout, ksize);
  if (g.ok) {
    long long need = (long long)g.splits * Cout * 9 * Cin * (long long)sizeof(float);
    if (!workspace || ws_bytes < need) { set_error("tc wgrad (row): workspace too small (%lld < %lld)", ws_bytes, need); return USTRUN_ERR_ARG; }
    TcWgradRowParams p{};
    p.B = B; p.H = H; p.W = W; p.Mo = g.Mo; p.Nn = g.Nn; p.TW = g.TW; p.segs_w = g.segs_w; p.nsegs = g.nsegs;
    p.R = g.R; p.hgroups = g.hgroups;
    p.m_tiles = g.m_tiles; p.n_tiles = g.n_tiles; p.splits = g.splits; p.segs_per_split = g.sps; p.flip = g.swap;
    p.ws = (float*)workspace;
    const ActView& ma = g.swap ? b : a;
    const ActView& nb = g.swap ? a : b;
    int rc = g.BN == 128 ? launch_wgrad_row<128>(ma, nb, g, p, st) : launch_wgrad_row<64>(ma, nb, g, p, st);
    if (rc) return rc;
    return launch_wgrad_reduce((const float*)workspace, g.splits, g.Mo, g.Nn, 9, dw, accumulate, st, g.swap);
  }
  return tc_wgrad_generic(a, &b, 1, ksize == 3 ? TAP_CONV3 : TAP_NONE, ksize * ksize, B, H, W, Cout, Cin, dw, accumulate, workspace, ws_bytes, st);
}
long long tc_conv_wgrad_ws(int B, int H, int W, int Cin, int Cout, int ksize) {
  WgRowPlan g = plan_wgrad_row(B, H, W, Cin, Cout, ksize);
  if (g.ok) return (long long)g.splits * Cout * 9 * Cin * (long long)sizeof(float);
  return tc_wgrad_ws_bytes(B, H, W, Cout, Cin, ksize * ksize);
}

// Host-side launch plans, exposed for tests and tools (no launch, no device access beyond the SM count):
//   what == 0: forward / dgrad tiling of a [B,H,W] pixel grid with Cout output channels
//              out = {BN, row_mode, boxes_per_tile, TW, TH, m_tiles, ctas_per_n_tile, n_tiles}
//   what == 1: row-mode weight gradient  out = {ok, swapped, BN, TW, rows_per_stage, segments, splits, work_items}
int tc_plan_query(int what, int B, int H, int W, int Cin, int Cout, int ksize, int* out) {
  if (what == 0) {
    TcConvParams p{};
    int BN = Cout % 256 == 0 ? 256 : (Cout % 128 == 0 ? 128 : 64);
    const bool row = row_mode() > 0 && ksize == 3 && BN <= 128 && W % 128 == 0;
    if (row) {
      p.TW = 128; p.TH = 1; p.halves = 1;
      p.m_tiles = B * ((W + 127) / 128) * H;
    } else {
      BN = plan_fwd_tiles(p, B, H, W, Cout);
    }
    out[0] = BN; out[1] = row ? 1 : 0; out[2] = p.halves; out[3] = p.TW; out[4] = p.TH; out[5] = p.m_tiles;
    out[6] = ctas_per_n_tile(Cout / BN, p.m_tiles); out[7] = Cout / BN;
    return 0;
  }
  if (what == 1) {
    WgRowPlan g = plan_wgrad_row(B, H, W, Cin, Cout, ksize);
    out[0] = g.ok ? 1 : 0; out[1] = g.swap; out[2] = g.BN; out[3] = g.TW; out[4] = g.R; out[5] = g.nsegs; out[6] = g.splits;
    out[7] = 3 * g.m_tiles * g.n_tiles;
    return 0;
  }
  set_error("tc_plan_query: what must be 0 or 1");
  return USTRUN_ERR_ARG;
}

// dW[ci][co][ij] = sum_p x[p][ci] * dy[b,2h+i,2w+j][co]
int tc_convT_wgrad(const void* dy, int lddy, const void* x, int ldx, float* dw, int accumulate, int B, int H, int W, int Cin, int Cout,
                   void* workspace, long long ws_bytes, cudaStream_t st) {
  if (Cin % 64 || Cout % 64 || lddy % 8 || ldx % 8) { set_error("tcgen05 convT wgrad needs Cin, Cout %% 64 == 0"); return USTRUN_ERR_ARG; }
  ActView a{x, Cin, W, H, B, ldx, (long long)W * ldx, (long long)H * W * ldx};
  ActView bv[4];
  for (int ij = 0; ij < 4; ++ij) {
    const char* base = (const char*)dy + ((long long)(ij >> 1) * 2 * W + (ij & 1)) * lddy * 2;
    bv[ij] = ActView{base, Cout, W, H, B, 2LL * lddy, 4LL * W * lddy, 4LL * H * W * lddy};
  }
  return tc_wgrad_generic(a, bv, 4, TAP_PERMAP, 4, B, H, W, Cin, Cout, dw, accumulate, workspace, ws_bytes, st);
}
long long tc_convT_wgrad_ws(int B, int H, int W, int Cin, int Cout) { return tc_wgrad_ws_bytes(B, H, W, Cin, Cout, 4); }

}  // namespace ustrun
