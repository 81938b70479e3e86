// BatchNorm finalize fused with its cross-rank reduction over NVLink/NVSwitch PEER MEMORY.
//
// Data-parallel BN needs, per layer and pass, the sum over ranks of a tiny [2*C] vector
// (forward: sum x, sum x^2; backward: sum g', sum g' xhat).  As separate launches that is a partial-row
// reduction kernel + an NCCL all-reduce (latency ~20-50 us, 230+ times per step) + the finalize
// kernel.  Here ONE kernel does all three: every warp reduces its channel's partial rows and PUSHES the
// two doubles into every rank's symmetric buffer as self-validating 8-byte packets {32 data bits, 32-bit
// sequence number} (the "LL" protocol of collective libraries: an aligned 8-byte store is single-copy atomic,
// so a packet whose flag equals this call's sequence number carries this call's data -- no fences, no flag
// array, no counters).  Each warp then polls the packets the other ranks pushed into ITS OWN buffer (local
// memory, L2 hits) and adds them in rank order: identical summation order on every rank, so all ranks compute
// bit-identical statistics.  Cost per barrier = one NVLink store latency, instead of (round 1) a system fence,
// a grid-wide counter, a flag store per peer and `world` dependent peer LOADS (~2 us each).
//
// Buffers come from torch.distributed._symmetric_memory (one allocation per rank, peer-mapped).
// Layout of each rank's buffer: uint2 packet[2 slots][PEER_MAXW source ranks][PEER_CMAX channels][4 parts]
// (parts: low/high word of the first double, low/high word of the second).  Two slots (sequence parity) are
// enough: a rank can run at most one barrier ahead of any peer, because finishing barrier k needs every
// peer's packets of barrier k, which a peer pushes only after it has finished reading barrier k-1.
// Sequence numbers: `seq` (launch argument) + `*seq_base` (device word the host advances once per step), so the launch
// arguments of a step never change and the whole data-parallel step can be replayed as a CUDA graph.
#include "common.cuh"

namespace ustrun {

constexpr int PEER_CMAX = 1024;
constexpr int PEER_MAXW = 8;
constexpr long long PEER_TIMEOUT_CYCLES = 20000000000LL;    // ~10 s: never hang the GPU on a lost peer (the host polls `error`)

struct PeerCtx {
  unsigned char* base[PEER_MAXW];   // peer-mapped base address of every rank's buffer (own included)
  int rank, world;
  unsigned int seq;                 // same non-zero value on every rank for this call (with seq_base: its 1-based index within the step)
  const unsigned int* seq_base;     // device word (or null): barriers completed in earlier steps, modulo 0x7FFFFFFE; lets the launch
                                    // arguments stay the same from step to step, so a captured CUDA graph of the step can be replayed
  int* error;                       // local: set to 1 on timeout
};

__device__ __forceinline__ uint2* peer_packet(const PeerCtx& p, int dst_rank, int slot, int src_rank, int c) {
  return reinterpret_cast<uint2*>(p.base[dst_rank]) + (((size_t)slot * PEER_MAXW + src_rank) * PEER_CMAX + c) * 4;
}
__device__ __forceinline__ void st_packet(uint2* dst, unsigned int data, unsigned int flag) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(dst), "r"(data), "r"(flag) : "memory");
}
__device__ __forceinline__ uint2 ld_packet(const uint2* src) {
  uint2 v;
  asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(src) : "memory");
  return v;
}

// Called by ALL 32 lanes of the warp that owns channel c, every lane holding the warp totals (a, b).  Lane l serves
// rank l / 4, part l % 4: world <= 8 ranks x 4 packets = one packet per lane each way.  Returns the sums over all ranks
// (in every lane); on a timeout sets *error and leaves the local sums.
__device__ __forceinline__ void peer_allreduce2(const PeerCtx& p, int c, int lane, double& a, double& b) {
  // effective sequence number 1 .. 2^31-2: consecutive over consecutive barriers (the modulus is even: slot parity alternates)
  const unsigned int seq = p.seq_base ? (p.seq - 1u + *p.seq_base) % 0x7FFFFFFEu + 1u : p.seq;
  const int slot = (int)(seq & 1u);
  const int r = lane >> 2, part = lane & 3;
  const bool active = r < p.world;
  if (active) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(part < 2 ? a : b);
    st_packet(peer_packet(p, r, slot, p.rank, c) + part, (part & 1) ? (unsigned int)(bits >> 32) : (unsigned int)bits, seq);
  }
  unsigned int word = 0u;
  bool ok = true;
  if (active) {
    const uint2* src = peer_packet(p, p.rank, slot, r, c) + part;
    const long long t0 = clock64();
    for (;;) {
      const uint2 v = ld_packet(src);
      if (v.y == seq) { word = v.x; break; }
      if (clock64() - t0 > PEER_TIMEOUT_CYCLES) { ok = false; break; }
    }
  }
  if (!__all_sync(0xffffffffu, ok)) {
    if (lane == 0) *p.error = 1;
    return;
  }
  double sa = 0.0, sb = 0.0;
  for (int rr = 0; rr < p.world; ++rr) {
    const unsigned int a_lo = __shfl_sync(0xffffffffu, word, rr * 4 + 0), a_hi = __shfl_sync(0xffffffffu, word, rr * 4 + 1);
    const unsigned int b_lo = __shfl_sync(0xffffffffu, word, rr * 4 + 2), b_hi = __shfl_sync(0xffffffffu, word, rr * 4 + 3);
    sa += __longlong_as_double((long long)(((unsigned long long)a_hi << 32) | a_lo));
    sb += __longlong_as_double((long long)(((unsigned long long)b_hi << 32) | b_lo));
  }
  a = sa;
  b = sb;
}

__global__ void k_bn_finalize_peer(const float* __restrict__ partials, int nparts, int C, double count_global,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ conv_bias,
                                   float* running_mean, float* running_var, long long* nbt, float momentum, float eps, float* scale,
                                   float* shift, float* mean_out, float* rstd_out, float* stat_out, PeerCtx p) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (blockIdx.x == 0 && threadIdx.x == 0 && nbt) *nbt += 1;
  if (c >= C) return;
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
  peer_allreduce2(p, c, lane, s, q);
  if (lane != 0) return;
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f, cb = conv_bias ? conv_bias[c] : 0.f;
  const double m = s / count_global;
  double var = q / count_global - m * m;
  if (var < 0.0) var = 0.0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps)), mf = (float)m;
  scale[c] = g * rstd;
  shift[c] = b - mf * g * rstd;
  mean_out[c] = mf;
  rstd_out[c] = rstd;
  if (running_mean || stat_out) {
    const double unbiased = count_global > 1.0 ? var * count_global / (count_global - 1.0) : var;
    const float bm = __fadd_rn(mf, cb), bv = (float)unbiased;
    if (stat_out) {              // deferred (multi-lane step): ustrun_bn_running_update applies the forwards in the reference's order
      stat_out[c] = bm;
      stat_out[C + c] = bv;
    } else {
      running_mean[c] = bn_running(running_mean[c], bm, momentum);
      running_var[c] = bn_running(running_var[c], bv, momentum);
    }
  }
}

__global__ void k_bn_bwd_finalize_peer(const float* __restrict__ partials, int nparts, int C, double count_global,
                                       const float* __restrict__ gamma, const float* __restrict__ rstd, float* dgamma, float* dbeta,
                                       int accumulate, float* coef, PeerCtx p) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= C) return;
  double s1 = 0.0, s2 = 0.0, t1 = 0.0, t2 = 0.0, u1 = 0.0, u2 = 0.0, v1 = 0.0, v2 = 0.0;   // independent chains: 8 loads in flight
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
  // parameter gradients stay LOCAL sums: the gradient all-reduce adds them across ranks exactly once
  if (lane == 0) {
    if (dgamma) dgamma[c] = (accumulate ? dgamma[c] : 0.f) + (float)s2;
    if (dbeta) dbeta[c] = (accumulate ? dbeta[c] : 0.f) + (float)s1;
  }
  peer_allreduce2(p, c, lane, s1, s2);
  if (lane != 0) return;
  const float g = gamma ? gamma[c] : 1.f;
  coef[c] = g * rstd[c];
  coef[C + c] = (float)(s1 / count_global);
  coef[2 * C + c] = (float)(s2 / count_global);
}

static int fill_ctx(PeerCtx& p, const void* const* peer_bases, int rank, int world, unsigned int seq, unsigned int* counter, int* error) {
  if (!peer_bases || world < 1 || world > PEER_MAXW || rank < 0 || rank >= world || !error || seq == 0u) {
    set_error("peer BN: bad peer arguments (world %d, rank %d)", world, rank);
    return USTRUN_ERR_ARG;
  }
  for (int r = 0; r < PEER_MAXW; ++r) p.base[r] = (unsigned char*)(r < world ? peer_bases[r] : nullptr);
  p.rank = rank; p.world = world; p.seq = seq; p.seq_base = counter; p.error = error;
  return 0;
}

}  // namespace ustrun

using namespace ustrun;

extern "C" {

long long ustrun_peer_buffer_bytes(void) { return 2LL * PEER_MAXW * PEER_CMAX * 4 * (long long)sizeof(uint2); }

int ustrun_bn_finalize_peer(const float* partials, int nparts, int C, double count_global, const float* gamma, const float* beta,
                            const float* conv_bias, float* running_mean, float* running_var, long long* nbt, float momentum, float eps,
                            float* scale, float* shift, float* mean, float* rstd, float* stat_out, const void* const* peer_bases, int rank, int world,
                            unsigned int seq, unsigned int* counter, int* error, void* stream) {
  USTRUN_REQUIRE(partials && nparts > 0 && C > 0 && C <= PEER_CMAX && count_global > 0 && scale && shift && mean && rstd, "bn_finalize_peer: bad args");
  USTRUN_REQUIRE(!stat_out || (!running_mean && !running_var && !nbt), "bn_finalize_peer: stat_out replaces the in-place running update");
  PeerCtx p;
  int rc = fill_ctx(p, peer_bases, rank, world, seq, counter, error);
  if (rc) return rc;
  k_bn_finalize_peer<<<ceil_div(C, 8), 256, 0, (cudaStream_t)stream>>>(partials, nparts, C, count_global, gamma, beta, conv_bias, running_mean, running_var,
                                                                      nbt, momentum, eps, scale, shift, mean, rstd, stat_out, p);
  return check_launch("bn_finalize_peer");
}

int ustrun_bn_bwd_finalize_peer(const float* partials, int nparts, int C, double count_global, const float* gamma, const float* rstd, float* dgamma,
                                float* dbeta, int accumulate, float* coef, const void* const* peer_bases, int rank, int world, unsigned int seq,
                                unsigned int* counter, int* error, void* stream) {
  USTRUN_REQUIRE(partials && nparts > 0 && C > 0 && C <= PEER_CMAX && count_global > 0 && rstd && coef, "bn_bwd_finalize_peer: bad args");
  PeerCtx p;
  int rc = fill_ctx(p, peer_bases, rank, world, seq, counter, error);
  if (rc) return rc;
  k_bn_bwd_finalize_peer<<<ceil_div(C, 8), 256, 0, (cudaStream_t)stream>>>(partials, nparts, C, count_global, gamma, rstd, dgamma, dbeta, accumulate, coef, p);
  return check_launch("bn_bwd_finalize_peer");
}

}  // extern "C"
