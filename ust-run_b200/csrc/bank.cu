// Confidence-bank bookkeeping on the device (second half): the reference does this with numpy on the host
// after copying the hardness, the label maps and the masks back every step (train.py:745-781), and draws the CutMix partners
// from the bank with host-side lengths (train.py:612-625).  Here the bank length and the adaptive threshold live in device
// memory, so a training loop built on SSLTrainer.step never synchronises with the host.
//
//   k_bank_plan    one thread: simple_ulb_idx = hardness < choice_th (float64, as numpy compares), the FIFO plan
//                  "selected samples of this batch first, then the first newlen old entries", the new length and the new
//                  threshold min(choice_th, max hardness in the bank) / min(increase * choice_th, 0.1)
//   k_bank_gather  moves images (fp32) and label / mask planes (uint8) into the OTHER copy of the bank as the plan says
//   k_bank_choice  CutMix partner indices from host-supplied uniform draws and the device-side bank length
//   k_lq_select    keeps the hardest ("low quality") sample of the batch, chosen by the device-side lq_idx
//   k_cover_box    obtain_all_cover_box (train.py:242-251): bounding box of the union of up to four uint8 planes
#include "common.cuh"

namespace ustrun {

__global__ void k_bank_plan(const double* __restrict__ hardness, int Bu, const double* __restrict__ old_hard, double* __restrict__ new_hard,
                            int* __restrict__ n_state, double* __restrict__ th_state, int max_len, double increase, int* __restrict__ plan) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int n_old = *n_state;
  double th = *th_state;
  int cur = 0;
  for (int b = 0; b < Bu; ++b)
    if (hardness[b] < th) {                       // train.py:754
      if (cur < max_len) { plan[cur] = b; new_hard[cur] = hardness[b]; }
      ++cur;
    }
  if (cur > max_len) cur = max_len;               // cannot happen for Bu <= max_len (checked on the host)
  int n_new;
  if (n_old == 0) {                               // train.py:756-764
    n_new = cur;
  } else if (cur > 0) {                           // train.py:766-777
    const int newlen = (n_old + cur > max_len) ? (max_len - cur) : n_old;
    for (int j = 0; j < newlen; ++j) { plan[cur + j] = -1 - j; new_hard[cur + j] = old_hard[j]; }
    n_new = cur + newlen;
  } else {                                        // nothing selected: the bank stays, the threshold relaxes (train.py:778-779)
    for (int j = 0; j < n_old; ++j) { plan[j] = -1 - j; new_hard[j] = old_hard[j]; }
    n_new = n_old;
    const double t = increase * th;
    th = t < 0.1 ? t : 0.1;
  }
  if (cur > 0 && n_new > 0) {
    double mx = new_hard[0];
    for (int j = 1; j < n_new; ++j) mx = new_hard[j] > mx ? new_hard[j] : mx;
    th = th < mx ? th : mx;                       // min(choice_th, cor_hardness.max())
  }
  *n_state = n_new;
  *th_state = th;
  plan[max_len] = n_new;
}

// grid (max_len, chunks): slot j of the new bank <- batch sample plan[j] >= 0, or old slot -1 - plan[j]
__global__ void __launch_bounds__(256)
k_bank_gather(const int* __restrict__ plan, int max_len, const float* __restrict__ b_img, const uint8_t* __restrict__ b_pl, const uint8_t* __restrict__ b_mask,
              const float* __restrict__ o_img, const uint8_t* __restrict__ o_pl, const uint8_t* __restrict__ o_mask, float* __restrict__ n_img,
              uint8_t* __restrict__ n_pl, uint8_t* __restrict__ n_mask, long long img_elems, long long lab_elems) {
  const int j = blockIdx.x;
  if (j >= plan[max_len]) return;
  const int src = plan[j];
  const float* si = src >= 0 ? b_img + (long long)src * img_elems : o_img + (long long)(-1 - src) * img_elems;
  const uint8_t* sp = src >= 0 ? b_pl + (long long)src * lab_elems : o_pl + (long long)(-1 - src) * lab_elems;
  const uint8_t* sm = src >= 0 ? b_mask + (long long)src * lab_elems : o_mask + (long long)(-1 - src) * lab_elems;
  const long long stride = (long long)gridDim.y * blockDim.x, t0 = (long long)blockIdx.y * blockDim.x + threadIdx.x;
  for (long long i = t0; i < img_elems; i += stride) n_img[(long long)j * img_elems + i] = si[i];
  for (long long i = t0; i < lab_elems; i += stride) {
    n_pl[(long long)j * lab_elems + i] = sp[i];
    n_mask[(long long)j * lab_elems + i] = sm[i];
  }
}

__global__ void k_bank_choice(const int* __restrict__ n_state, int Bl, int Bu, const int* __restrict__ r_lb, const double* __restrict__ r_u,
                              const int* __restrict__ perm, int* __restrict__ choice) {
  const int i = threadIdx.x;
  if (i >= Bu) return;
  const int n = *n_state;
  if (n == 0) { choice[i] = r_lb[i]; return; }                    // train.py:615
  int k = Bu / 2;                                                 // int(len(ulb_x_s) * 0.5)
  if (k > n) k = n;
  const int src = perm[i];                                        // np.random.permutation(concatenate((in_lb, in_simple)))
  choice[i] = src < Bu - k ? r_lb[src] : Bl + (int)floor(r_u[src - (Bu - k)] * (double)n);
}

__global__ void __launch_bounds__(256)
k_lq_select(const int* __restrict__ lq_idx, const float* __restrict__ img, const uint8_t* __restrict__ pl, const uint8_t* __restrict__ mask,
            float* __restrict__ o_img, uint8_t* __restrict__ o_pl, uint8_t* __restrict__ o_mask, long long img_elems, long long lab_elems) {
  const int s = *lq_idx;
  const long long stride = (long long)gridDim.x * blockDim.x, t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (long long i = t0; i < img_elems; i += stride) o_img[i] = img[(long long)s * img_elems + i];
  for (long long i = t0; i < lab_elems; i += stride) {
    o_pl[i] = pl[(long long)s * lab_elems + i];
    o_mask[i] = mask[(long long)s * lab_elems + i];
  }
}

// one block: rows y1..y2 = first / last row holding a non-zero pixel, columns x1..x2 = smallest / largest such column
__global__ void __launch_bounds__(256)
k_cover_box(const uint8_t* __restrict__ p0, const uint8_t* __restrict__ p1, const uint8_t* __restrict__ p2, const uint8_t* __restrict__ p3, int H, int W,
            const uint8_t* __restrict__ fallback, uint8_t* __restrict__ box) {
  __shared__ int s_y1, s_y2, s_x1, s_x2;
  if (threadIdx.x == 0) { s_y1 = H; s_y2 = -1; s_x1 = W; s_x2 = -1; }
  __syncthreads();
  int y1 = H, y2 = -1, x1 = W, x2 = -1;
  for (int i = threadIdx.x; i < H * W; i += blockDim.x) {
    const bool nz = (p0 && p0[i]) || (p1 && p1[i]) || (p2 && p2[i]) || (p3 && p3[i]);
    if (nz) {
      const int y = i / W, x = i - y * W;
      y1 = min(y1, y); y2 = max(y2, y); x1 = min(x1, x); x2 = max(x2, x);
    }
  }
  atomicMin(&s_y1, y1); atomicMax(&s_y2, y2); atomicMin(&s_x1, x1); atomicMax(&s_x2, x2);
  __syncthreads();
  const bool empty = s_y2 < 0;
  for (int i = threadIdx.x; i < H * W; i += blockDim.x) {
    const int y = i / W, x = i - y * W;
    box[i] = empty ? (fallback ? fallback[i] : (uint8_t)0) : (uint8_t)(y >= s_y1 && y <= s_y2 && x >= s_x1 && x <= s_x2);
  }
}

}  // namespace ustrun

using namespace ustrun;

extern "C" {

int ustrun_bank_update(const double* hardness, int Bu, const float* batch_img, const unsigned char* batch_pl, const unsigned char* batch_mask,
                       const float* old_img, const unsigned char* old_pl, const unsigned char* old_mask, const double* old_hard, float* new_img,
                       unsigned char* new_pl, unsigned char* new_mask, double* new_hard, int* n_state, double* th_state, int max_len, double increase,
                       long long img_elems, long long lab_elems, int* plan_ws, void* stream) {
  USTRUN_REQUIRE(hardness && batch_img && batch_pl && batch_mask && old_img && old_pl && old_mask && old_hard && new_img && new_pl && new_mask && new_hard &&
                     n_state && th_state && plan_ws,
                 "bank_update: null argument");
  USTRUN_REQUIRE(Bu > 0 && max_len >= Bu && img_elems > 0 && lab_elems > 0, "bank_update: need 0 < Bu <= max_len (got Bu=%d, max_len=%d)", Bu, max_len);
  cudaStream_t st = (cudaStream_t)stream;
  k_bank_plan<<<1, 32, 0, st>>>(hardness, Bu, old_hard, new_hard, n_state, th_state, max_len, increase, plan_ws);
  int rc = check_launch("bank_plan");
  if (rc) return rc;
  int chunks = (int)((img_elems + 256 * 8 - 1) / (256 * 8));
  if (chunks > 64) chunks = 64;
  if (chunks < 1) chunks = 1;
  k_bank_gather<<<dim3(max_len, chunks), 256, 0, st>>>(plan_ws, max_len, batch_img, batch_pl, batch_mask, old_img, old_pl, old_mask, new_img, new_pl, new_mask,
                                                      img_elems, lab_elems);
  return check_launch("bank_gather");
}

int ustrun_bank_choice(const int* n_state, int Bl, int Bu, const int* r_lb, const double* r_u, const int* perm, int* choice, void* stream) {
  USTRUN_REQUIRE(n_state && r_lb && r_u && perm && choice && Bl > 0 && Bu > 0 && Bu <= 1024, "bank_choice: bad args");
  k_bank_choice<<<1, ((Bu + 31) / 32) * 32, 0, (cudaStream_t)stream>>>(n_state, Bl, Bu, r_lb, r_u, perm, choice);
  return check_launch("bank_choice");
}

int ustrun_lq_select(const int* lq_idx, const float* img, const unsigned char* pl, const unsigned char* mask, float* out_img, unsigned char* out_pl,
                     unsigned char* out_mask, long long img_elems, long long lab_elems, void* stream) {
  USTRUN_REQUIRE(lq_idx && img && pl && mask && out_img && out_pl && out_mask && img_elems > 0 && lab_elems > 0, "lq_select: bad args");
  int grid = (int)((img_elems + 256 * 4 - 1) / (256 * 4));
  if (grid > 148) grid = 148;
  k_lq_select<<<grid, 256, 0, (cudaStream_t)stream>>>(lq_idx, img, pl, mask, out_img, out_pl, out_mask, img_elems, lab_elems);
  return check_launch("lq_select");
}

int ustrun_cover_box(const unsigned char* p0, const unsigned char* p1, const unsigned char* p2, const unsigned char* p3, int H, int W,
                     const unsigned char* fallback, unsigned char* box, void* stream) {
  USTRUN_REQUIRE((p0 || p1 || p2 || p3) && box && H > 0 && W > 0, "cover_box: bad args");
  k_cover_box<<<1, 256, 0, (cudaStream_t)stream>>>(p0, p1, p2, p3, H, W, fallback, box);
  return check_launch("cover_box");
}

}  // extern "C"
