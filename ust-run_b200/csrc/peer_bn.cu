// BatchNorm finalize fused with its cross-rank reduction over NVLink/NVSwitch PEER MEMORY.
//
// Data-parallel BN needs, per layer and pass, the sum over ranks of a tiny [2*C] vector
// (forward: sum x, sum x^2; backward: sum g', sum g' xhat).  As separate launches that is a partial-row
// reduction kernel + an NCCL all-reduce (latency ~20-50 us, 230+ times per step) + the finalize
// kernel.  Here ONE kernel does all three: every warp reduces its channel's partial rows, publishes
// the two doubles in this rank's symmetric buffer, the last warp of the grid pushes a sequence flag into
// every peer's flag array (st over NVLink), all warps wait for the peers' flags in LOCAL memory, read
// the peers' doubles with peer loads and finish the BatchNorm maths -- identical summation order on
// every rank, so all ranks compute bit-identical statistics.
//
// Buffers come from torch.distributed._symmetric_memory (one allocation per rank, peer-mapped).
// Layout of each rank's buffer: uint32 flags[32] | double data[2 slots][2][PEER_CMAX].
#include "common.cuh"

namespace ustrun {

constexpr int PEER_CMAX = 1024;
constexpr int PEER_MAXW = 8;
constexpr long long PEER_TIMEOUT_CYCLES = 6000000000LL;     // ~3 s: never hang the GPU on a lost peer

struct PeerCtx {
  unsigned char* base[PEER_MAXW];   // peer-mapped base address of every rank's buffer (own included)
  int rank, world;
  unsigned int seq;                 // same value on every rank for this call
  unsigned int* counter;            // local: warps that published
  int* error;                       // local: set to 1 on timeout
};

__device__ __forceinline__ volatile unsigned int* peer_flags(const PeerCtx& p, int r) { return reinterpret_cast<volatile unsigned int*>(p.base[r]); }
__device__ __forceinline__ double* peer_data(const PeerCtx& p, int r, int slot) {
  return reinterpret_cast<double*>(p.base[r] + 128) + (size_t)slot * 2 * PEER_CMAX;
}

// lane 0 of every warp calls this with its channel's local sums; returns the sums over all ranks
__device__ __forceinline__ void peer_allreduce2(const PeerCtx& p, int c, int nwarps, double& a, double& b) {
  const int slot = p.seq & 1;
  double* mine = peer_data(p, p.rank, slot);
  mine[c] = a;
  mine[PEER_CMAX + c] = b;
  __threadfence_system();
  const unsigned int prev = atomicAdd(p.counter, 1u);
  if (prev == (unsigned int)nwarps - 1u) {                 // last publisher of this rank: signal every peer
    *p.counter = 0u;
    __threadfence_system();
    for (int r = 0; r < p.world; ++r) peer_flags(p, r)[p.rank] = p.seq;
    __threadfence_system();
  }
  const long long t0 = clock64();
  volatile unsigned int* myflags = peer_flags(p, p.rank);
  for (int r = 0; r < p.world; ++r) {
    // flags only grow; wrap-around safe comparison
    while ((int)(myflags[r] - p.seq) < 0) {
      if (clock64() - t0 > PEER_TIMEOUT_CYCLES) { *p.error = 1; return; }
    }
  }
  __threadfence_system();
  double sa = 0.0, sb = 0.0;
  for (int r = 0; r < p.world; ++r) {
    const volatile double* d = peer_data(p, r, slot);
    sa += d[c];
    sb += d[PEER_CMAX + c];
  }
  a = sa;
  b = sb;
}

__global__ void k_bn_finalize_peer(const float* __restrict__ partials, int nparts, int C, double count_global,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ conv_bias,
                                   float* running_mean, float* running_var, long long* nbt, float momentum, float eps, float* scale,
                                   float* shift, float* mean_out, float* rstd_out, PeerCtx p) {
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
  if (lane != 0) return;
  peer_allreduce2(p, c, C, s, q);
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f, cb = conv_bias ? conv_bias[c] : 0.f;
  const double m = s / count_global;
  double var = q / count_global - m * m;
  if (var < 0.0) var = 0.0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps)), mf = (float)m;
  scale[c] = g * rstd;
  shift[c] = b - mf * g * rstd;
  mean_out[c] = mf;
  rstd_out[c] = rstd;
  if (running_mean) {
    const double unbiased = count_global > 1.0 ? var * count_global / (count_global - 1.0) : var;
    running_mean[c] = bn_running(running_mean[c], __fadd_rn(mf, cb), momentum);
    running_var[c] = bn_running(running_var[c], (float)unbiased, momentum);
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
  if (lane != 0) return;
  // parameter gradients stay LOCAL sums: the gradient all-reduce adds them across ranks exactly once
  if (dgamma) dgamma[c] = (accumulate ? dgamma[c] : 0.f) + (float)s2;
  if (dbeta) dbeta[c] = (accumulate ? dbeta[c] : 0.f) + (float)s1;
  peer_allreduce2(p, c, C, s1, s2);
  const float g = gamma ? gamma[c] : 1.f;
  coef[c] = g * rstd[c];
  coef[C + c] = (float)(s1 / count_global);
  coef[2 * C + c] = (float)(s2 / count_global);
}

static int fill_ctx(PeerCtx& p, const void* const* peer_bases, int rank, int world, unsigned int seq, unsigned int* counter, int* error) {
  if (!peer_bases || world < 1 || world > PEER_MAXW || rank < 0 || rank >= world || !counter || !error) {
    set_error("peer BN: bad peer arguments (world %d, rank %d)", world, rank);
    return USTRUN_ERR_ARG;
  }
  for (int r = 0; r < PEER_MAXW; ++r) p.base[r] = (unsigned char*)(r < world ? peer_bases[r] : nullptr);
  p.rank = rank; p.world = world; p.seq = seq; p.counter = counter; p.error = error;
  return 0;
}

}  // namespace ustrun

using namespace ustrun;

extern "C" {

long long ustrun_peer_buffer_bytes(void) { return 128 + 2LL * 2 * PEER_CMAX * (long long)sizeof(double); }

int ustrun_bn_finalize_peer(const float* partials, int nparts, int C, double count_global, const float* gamma, const float* beta,
                            const float* conv_bias, float* running_mean, float* running_var, long long* nbt, float momentum, float eps,
                            float* scale, float* shift, float* mean, float* rstd, const void* const* peer_bases, int rank, int world,
                            unsigned int seq, unsigned int* counter, int* error, void* stream) {
  USTRUN_REQUIRE(partials && nparts > 0 && C > 0 && C <= PEER_CMAX && count_global > 0 && scale && shift && mean && rstd, "bn_finalize_peer: bad args");
  PeerCtx p;
  int rc = fill_ctx(p, peer_bases, rank, world, seq, counter, error);
  if (rc) return rc;
  k_bn_finalize_peer<<<ceil_div(C, 8), 256, 0, (cudaStream_t)stream>>>(partials, nparts, C, count_global, gamma, beta, conv_bias, running_mean, running_var,
                                                                      nbt, momentum, eps, scale, shift, mean, rstd, p);
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
