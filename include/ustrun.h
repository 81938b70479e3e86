/* libustrun_sm100.so -- C ABI of the B200-native UST-RUN SSL train-step kernels.
 *
 * The reference (MQinghe/UST-RUN) is pure Python/PyTorch and has no FFI of its own; the boundary a
 * maintainer binds is therefore "one entry point per library kernel the reference's hot path
 * dispatches" (SURVEY.md section 2.2 K1-K16, section 8b).  Each declaration cites the reference
 * call site it replaces (paths relative to the reference repo).
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless marked "host".
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream).
 *   - returns 0 on success, USTRUN_ERR_ARG (-1) on an argument/shape error, otherwise the
 *     cudaError_t of the failed launch; `ustrun_last_error_string()` has the message.
 *   - never allocates, frees or synchronises; workspaces are caller-provided.
 *   - activations are NHWC with an explicit pixel pitch `ld*` in elements (so a tensor may be a
 *     channel slice of a wider concat buffer); `dtype` selects fp32 (validation mode) or bf16.
 *   - there is NO CPU path: every entry needs an sm_100 device.
 */
#ifndef USTRUN_H_
#define USTRUN_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define USTRUN_ABI_VERSION 1
#define USTRUN_ERR_ARG (-1)

enum { USTRUN_F32 = 0, USTRUN_BF16 = 1 };
enum { USTRUN_ACT_NONE = 0, USTRUN_ACT_RELU = 1, USTRUN_ACT_LEAKY = 2 };
enum { USTRUN_IMPL_SIMT = 0, USTRUN_IMPL_TCGEN05 = 1 };
/* rows a per-channel partial-sum workspace must provide: float[USTRUN_MAX_PARTS][2][C] */
#define USTRUN_MAX_PARTS 1280

int ustrun_abi_version(void);
const char* ustrun_last_error_string(void);
/* 1 if the current CUDA device is compute capability 10.x, 0 otherwise, <0 on CUDA error */
int ustrun_device_supported(void);

/* ---- layout / packing ------------------------------------------------------------------- */
/* inputs arrive NCHW fp32 (dataloaders/custom_transforms.py:739-747) */
int ustrun_nchw_to_nhwc(const float* src, void* dst, int dtype, int B, int C, int H, int W, int ld_dst, void* stream);
int ustrun_nhwc_to_nchw(const void* src, int dtype, int ld_src, float* dst, int B, int C, int H, int W, void* stream);
/* nn.Conv2d weight OIHW fp32 -> wf[Cout][k*k][Cin] (forward) and wd[Cin][k*k flipped][Cout] (dgrad) */
int ustrun_pack_conv_weight(const float* w, void* wf, void* wd, int dtype, int Cout, int Cin, int ksize, void* stream);
/* nn.ConvTranspose2d(k2,s2) weight [Cin][Cout][2][2] -> wf[4][Cout][Cin], wd[Cin][4][Cout] */
int ustrun_pack_convT_weight(const float* w, void* wf, void* wd, int dtype, int Cin, int Cout, void* stream);

/* Every conv / transposed-conv weight of a model in ONE launch (after the optimiser step all packed copies are stale).
 * Entry: OIHW (transposed = 0: wf[Cout][taps][Cin], wd[Cin][taps flipped][Cout]) or ConvTranspose2d [Cin][Cout][2][2]
 * (transposed = 1: wf[4][Cout][Cin], wd[Cin][4][Cout]); wf / wd nullable.  Block i handles tile blk_tile[i] of entry
 * blk_entry[i]: conv tiles are 32 x 32 (co, ci) numbered co_tile * ceil(Cin / 32) + ci_tile, transposed-conv tiles are runs of
 * 4096 elements. */
typedef struct {
  const float* w;
  void* wf;
  void* wd;
  int Cout, Cin, taps, transposed;
} ustrun_pack_t;
int ustrun_pack_weights_multi(const ustrun_pack_t* table, const int* blk_entry, const int* blk_tile, int nblocks, int dtype, void* stream);

/* ---- convolutions (K1,K2,K3,K8,K10) ------------------------------------------------------- */
/* nn.Conv2d 3x3 s1 p1 / 1x1 forward: unet_parts.py:16,19,74; unet.py:37-43,81,85,88,182.
 * Also used for dgrad with the `wd` packing (autograd of the same call sites).
 * y = conv(x) [+ bias]; if `partials` != NULL also writes per-channel (sum, sum of squares) of the
 * fp32 results (before bias, before rounding) as float[*nparts][2][Cout] for train-mode BatchNorm.
 * out_nchw_f32 != 0: y is a float NCHW tensor (the logits head) instead of NHWC `dtype`. */
int ustrun_conv_fwd(int impl, const void* x, int ldx, const void* w_packed, const float* bias, void* y, int ldy,
                    int dtype, int B, int H, int W, int Cin, int Cout, int ksize, int out_nchw_f32,
                    float* partials, int* nparts_host, void* stream);
/* dW[Cout][Cin][k][k] (fp32, OIHW like the nn.Parameter) (+)= sum_p dy[p][co] * x[p+tap][ci].
 * accumulate != 0 adds to dw (gradient accumulation over the four loss branches, train.py:838). */
int ustrun_conv_wgrad(int impl, const void* dy, int lddy, const void* x, int ldx, float* dw, int accumulate,
                      int dtype, int B, int H, int W, int Cin, int Cout, int ksize, void* workspace,
                      long long workspace_bytes, void* stream);
long long ustrun_conv_wgrad_workspace_bytes(int impl, int B, int H, int W, int Cin, int Cout, int ksize);
/* Host-side launch plan of the tcgen05 kernels for a shape (no launch; for tests and tools).
 * what == 0: forward/dgrad tiling -> out8 = {N tile, row mode, 64-pixel boxes per M tile (1|2), TW, TH, M tiles,
 *            CTAs per N tile (= BatchNorm partial rows), N tiles};
 * what == 1: row-mode weight gradient -> out8 = {usable, operands swapped, N tile, TW, image rows per stage, K segments,
 *            K splits, work items per split}. */
int ustrun_tc_plan_query(int what, int B, int H, int W, int Cin, int Cout, int ksize, int* out8);
/* nn.ConvTranspose2d(k2,s2)+bias: unet_parts.py:53,57.  x is [B,H,W,Cin]; y is [B,2H,2W,Cout]. */
int ustrun_convT2x2_fwd(int impl, const void* x, int ldx, const void* wf, const float* bias, void* y, int ldy,
                        int dtype, int B, int H, int W, int Cin, int Cout, void* stream);
int ustrun_convT2x2_dgrad(int impl, const void* dy, int lddy, const void* wd, void* dx, int lddx, int dtype,
                          int B, int H, int W, int Cin, int Cout, void* stream);
int ustrun_convT2x2_wgrad(int impl, const void* dy, int lddy, const void* x, int ldx, float* dw, int accumulate,
                          int dtype, int B, int H, int W, int Cin, int Cout, void* workspace,
                          long long workspace_bytes, void* stream);
/* out[c] (+)= sum_p x[p][c]  (bias gradients of convT / head convs) */
int ustrun_channel_sum(const void* x, int ldx, int dtype, long long npix, int C, float* out, int accumulate,
                       float* workspace, void* stream);

/* ---- BatchNorm / DSBN (K4,K5) : unet_parts.py:17,20; unet.py:19; dsbn.py:11,24-27 ---------- */
/* sums[2][C] = sum over parts (for cross-rank reduction before finalize) */
int ustrun_bn_reduce_partials(const float* partials, int nparts, int C, float* sums, void* stream);
/* training: batch statistics from partials (nparts rows), running-stat update (unbiased var,
 * momentum), num_batches_tracked += 1; eval: running statistics.  conv_bias (nullable) is the
 * bias of the preceding conv (UNet-B), which the conv kernels do not add.  Outputs per channel:
 * scale, shift (y = x*scale + shift), mean, rstd (saved for backward). */
int ustrun_bn_finalize(const float* partials, int nparts, int C, double count, const float* gamma, const float* beta,
                       const float* conv_bias, float* running_mean, float* running_var, long long* num_batches_tracked,
                       float momentum, float eps, int training, float* scale, float* shift, float* mean, float* rstd,
                       float* stat_out, void* stream);
/* Deferred running statistics (multi-lane step: independent forwards of one step run concurrently, the running statistics
 * must still be updated in the reference's forward order, train.py:643-647,668,699-702,740).  ustrun_bn_finalize with
 * stat_out != NULL (and running_* / nbt NULL) writes the batch statistics {mean + conv bias, unbiased variance} to
 * stat_out[2][C] instead of updating in place; ustrun_bn_running_update then applies, for n BatchNorm layers in ONE launch,
 * the slots named by `mask` in increasing slot order with the same arithmetic as the in-place update (bit-identical), and
 * adds popcount(mask) to num_batches_tracked.  stats: float[slots][2][C] of that layer. */
typedef struct {
  float* running_mean;
  float* running_var;
  long long* nbt;        /* num_batches_tracked or NULL */
  const float* stats;    /* [slot][2][C] */
  int C;
  unsigned int mask;     /* bit s: slot s holds a forward of this step */
  float momentum;
  int pad;
} ustrun_bn_update_t;
int ustrun_bn_running_update(const ustrun_bn_update_t* table, int n, int Cmax, void* stream);
/* y = act(x*scale+shift); optional 2x2 max-pooled copy (nn.MaxPool2d(2): unet_parts.py:34, unet.py:45) */
int ustrun_bn_act_fwd(const void* x, int ldx, const float* scale, const float* shift, int act, void* y, int ldy,
                      void* pooled, int ldp, int dtype, int B, int H, int W, int C, void* stream);
/* partials[*nparts][2][C] of (sum g', sum g'*xhat), g' = g * act'(x*scale+shift) */
int ustrun_bn_bwd_reduce(const void* g, int ldg, const void* x, int ldx, const float* mean, const float* rstd,
                         const float* scale, const float* shift, int act, int dtype, long long npix, int C,
                         float* partials, int* nparts_host, void* stream);
/* dgamma/dbeta (+)= param_grad_scale * (sum g' xhat, sum g'); coef[3][C] = (gamma*rstd, sum g'/count, sum g' xhat/count).
 * dgamma / dbeta / coef are each nullable (the multi-lane step computes coef on the lane's stream and accumulates the
 * parameter gradients with a second call on the weight-gradient stream, keeping the accumulation order of the branches).
 * With cross-rank statistics the sums are global: pass param_grad_scale = 1/world so that the later
 * gradient all-reduce (a sum over ranks) yields the global dgamma/dbeta exactly once. */
int ustrun_bn_bwd_finalize(const float* partials, int nparts, int C, double count, const float* gamma, const float* rstd,
                           float* dgamma, float* dbeta, int accumulate, float param_grad_scale, float* coef, void* stream);
int ustrun_bn_bwd_apply(const void* g, int ldg, const void* x, int ldx, const float* mean, const float* rstd,
                        const float* scale, const float* shift, const float* coef, int act, void* dx, int lddx,
                        int dtype, long long npix, int C, void* stream);
/* gout = gskip (nullable) + unpool(dpool) routed to the first max of each 2x2 window of y */
int ustrun_maxpool_bwd(const void* y, int ldy, const void* dpool, int ldp, const void* gskip, int ldgs, void* gout,
                       int ldgo, int dtype, int B, int H, int W, int C, void* stream);
/* nn.Upsample(x2, bilinear): unet.py:84,127 (align_corners=False), unet_parts.py:50 (True) */
int ustrun_upsample2x_fwd(const void* x, int ldx, void* y, int ldy, int dtype, int B, int H, int W, int C,
                          int align_corners, void* stream);
int ustrun_upsample2x_bwd(const void* dy, int lddy, void* dx, int lddx, int dtype, int B, int H, int W, int C,
                          int align_corners, void* stream);

/* ---- BatchNorm finalize fused with the cross-rank reduction over NVLink peer memory ----
 * peer_bases: HOST array of `world` device pointers, entry r = rank r's symmetric buffer of
 * ustrun_peer_buffer_bytes() bytes (zero-initialised once, e.g. torch.distributed._symmetric_memory),
 * mapped into this process.  The barrier's sequence number is (seq - 1 + *seq_base) % 0x7FFFFFFE + 1 and must be the
 * same on every rank and consecutive from call to call: `seq` (launch argument, >= 1) is the call's index within the
 * step, `seq_base` a local device word (uint32, nullable = 0) that the host advances once per step by the number of
 * calls of the step before -- the launch arguments then repeat from step to step and a captured CUDA graph of the
 * data-parallel step can be replayed.  error (int32) is a local device scalar, set to 1 when a peer did not answer
 * within ~10 s.  stat_out (nullable, excludes running_* / num_batches_tracked): [2*C] = {batch mean + conv bias,
 * unbiased variance} for a deferred ustrun_bn_running_update.  One kernel replaces {partial-row reduction, NCCL
 * all-reduce of 2*C floats, finalize}. */
long long ustrun_peer_buffer_bytes(void);
int ustrun_bn_finalize_peer(const float* partials, int nparts, int C, double count_global, const float* gamma, const float* beta,
                            const float* conv_bias, float* running_mean, float* running_var, long long* num_batches_tracked,
                            float momentum, float eps, float* scale, float* shift, float* mean, float* rstd, float* stat_out,
                            const void* const* peer_bases, int rank, int world, unsigned int seq, unsigned int* seq_base,
                            int* error, void* stream);
/* dgamma/dbeta receive the LOCAL sums (the gradient all-reduce adds the ranks), coef the global means */
int ustrun_bn_bwd_finalize_peer(const float* partials, int nparts, int C, double count_global, const float* gamma,
                                const float* rstd, float* dgamma, float* dbeta, int accumulate, float* coef,
                                const void* const* peer_bases, int rank, int world, unsigned int seq, unsigned int* seq_base,
                                int* error, void* stream);

/* ---- pseudo labels (K12): train.py:649-697, train_mnms.py:595-623 -------------------------- */
/* softmax branch.  t1,t2,t3 (teacher) and s0 (student, nullable) are fp32 NCHW logits [Bu,C,H,W];
 * box, cut_label, cut_mask are uint8 ([Bu,H,W], [Nc,H,W], [Nc,H,W]); choice int32[Bu].
 * Outputs are uint8 [Bu,H,W]. stats (nullable) float[2]: sum(mask), sum(mask_w) for logging. */
int ustrun_pseudo_label_softmax(const float* t1, const float* t2, const float* t3, const float* s0,
                                const uint8_t* box, const uint8_t* cut_label, const uint8_t* cut_mask,
                                const int* choice, float threshold, int Bu, int C, int H, int W,
                                uint8_t* pl, uint8_t* mask, uint8_t* pl_w, uint8_t* mask_w, uint8_t* pl_ul,
                                uint8_t* mask_ul, uint8_t* pl_lu, uint8_t* mask_lu, uint8_t* stu_pl, void* stream);
/* sigmoid (fundus) branch: everything is per element [Bu,C,H,W]; box stays [Bu,H,W].
 * thr_hi = float(threshold), thr_lo = float(1-threshold) as torch casts them (train.py:651). */
int ustrun_pseudo_label_sigmoid(const float* t1, const float* t2, const float* t3, const float* s0,
                                const uint8_t* box, const uint8_t* cut_label, const uint8_t* cut_mask,
                                const int* choice, float thr_hi, float thr_lo, int Bu, int C, int H, int W,
                                uint8_t* pl, uint8_t* mask, uint8_t* pl_w, uint8_t* mask_w, uint8_t* pl_ul,
                                uint8_t* mask_ul, uint8_t* pl_lu, uint8_t* mask_lu, uint8_t* stu_pl, void* stream);
/* a*(1-box)+b*box on images (train.py:644,646,689,692); fp32 NCHW in, NHWC `dtype` out */
int ustrun_mix_to_nhwc(const float* a, const float* b, const int* b_index, const uint8_t* box, void* dst, int ld_dst,
                       int dtype, int B, int C, int H, int W, void* stream);

/* Input pipeline on the device: the reference's loaders end with Normalize_tf (custom_transforms.py:650-684:
 * float32(u8) / 127.5 - 1.0) and ToTensor (:728-753: H x W x C -> C x H x W) on the host and copy float32 batches to the GPU
 * (train.py:582-589).  ustrun_normalize_u8_to_nchw does both on the device, bit-exactly (uint8 [B,H,W,C] -> float32 [B,C,H,W]);
 * ustrun_mix_any_to_nhwc is ustrun_mix_to_nhwc with either source given as such a uint8 image batch (a_u8 / b_u8, normalised on
 * the fly) or as a float32 NCHW tensor (a_f32 / b_f32): 4x fewer bytes over PCIe, no separate normalisation pass. */
int ustrun_normalize_u8_to_nchw(const uint8_t* src, float* dst, int B, int C, int H, int W, void* stream);
int ustrun_mix_any_to_nhwc(const float* a_f32, const uint8_t* a_u8, const float* b_f32, const uint8_t* b_u8, const int* b_index, const uint8_t* box,
                           void* dst, int ld_dst, int dtype, int B, int C, int H, int W, void* stream);

/* ---- CE + Dice (K13,K14): train.py:816-838, utils/losses.py:194-268 ------------------------- */
/* Pass 1 reduces (3C+1) scalars; the finalize stage (same call) writes
 *   loss_out[0] = ce_w * mean_all(CE*mask) + dice_w * DiceLossWithMask   (either weight may be 0)
 *   loss_out[1] = CE part, loss_out[2] = Dice part, and coef[] for pass 2.
 * target uint8 [B,H,W]; mask uint8 [B,H,W] or NULL; class_weight float[C] or NULL (losses.py:251).
 * workspace: float[USTRUN_MAX_PARTS*(3C+1)]; coef: float[4*C+4]. */
int ustrun_ce_dice_softmax_fwd(const float* logits, const uint8_t* target, const uint8_t* mask, int B, int C, int H,
                               int W, float ce_w, float dice_w, const float* class_weight, float* workspace,
                               float* coef, float* loss_out, void* stream);
/* The two halves of the forward, for data parallelism: `partials` writes *nparts_host rows of (3C+1)
 * partial sums; the caller may sum rows (ustrun_reduce_rows) and all-reduce them across ranks, then
 * `finalize` with the GLOBAL pixel count gives the global-batch loss and coefficients. */
int ustrun_ce_dice_softmax_partials(const float* logits, const uint8_t* target, const uint8_t* mask, int B, int C, int H,
                                    int W, float* workspace, int* nparts_host, void* stream);
int ustrun_ce_dice_softmax_finalize(const float* workspace, int nparts, int C, double npix_total, float ce_w, float dice_w,
                                    const float* class_weight, float* coef, float* loss_out, void* stream);
/* out[c] = sum_r rows[r][c] */
int ustrun_reduce_rows(const float* rows, int nrows, int ncols, float* out, void* stream);
/* dlogits (+)= gscale * (*upstream, nullable => 1) * dLoss/dlogits */
int ustrun_ce_dice_softmax_bwd(const float* logits, const uint8_t* target, const uint8_t* mask, int B, int C, int H,
                               int W, const float* coef, const float* upstream, float gscale, float* dlogits,
                               int accumulate, void* stream);
/* sigmoid/multi (fundus) branch: BCEWithLogits*mask mean + one global Dice (losses.py:244-249);
 * target, mask uint8 [B,C,H,W] */
int ustrun_bce_dice_sigmoid_fwd(const float* logits, const uint8_t* target, const uint8_t* mask, int B, int C, int H,
                                int W, float ce_w, float dice_w, float* workspace, float* coef, float* loss_out,
                                void* stream);
int ustrun_bce_dice_sigmoid_partials(const float* logits, const uint8_t* target, const uint8_t* mask, int B, int C, int H,
                                     int W, float* workspace, int* nparts_host, void* stream);
int ustrun_bce_dice_sigmoid_finalize(const float* workspace, int nparts, double nelem_total, float ce_w, float dice_w,
                                     float* coef, float* loss_out, void* stream);
int ustrun_bce_dice_sigmoid_bwd(const float* logits, const uint8_t* target, const uint8_t* mask, int B, int C, int H,
                                int W, const float* coef, const float* upstream, float gscale, float* dlogits,
                                int accumulate, void* stream);

/* ---- SGD + EMA (K15,K16): train.py:512,840-851,87-93 ---------------------------------------- */
typedef struct {
  float* p;      /* student parameter (fp32) */
  float* g;      /* gradient or NULL (parameter skipped, like torch SGD with grad=None) */
  float* buf;    /* momentum buffer */
  float* ema;    /* teacher parameter or NULL */
  long long n;   /* elements */
  int first;     /* 1: momentum buffer not initialised yet (buf = g) */
  int pad;
} ustrun_param_t;
/* One launch over all tensors. blk_tensor[i], blk_offset[i]: tensor index and element offset of
 * block i's chunk (chunk = USTRUN_OPT_CHUNK elements).  g <- g + wd*p; buf <- first? g : mu*buf+g;
 * p <- p - lr*buf; ema <- alpha*ema + (1-alpha)*p.  do_sgd/do_ema select the halves. */
#define USTRUN_OPT_CHUNK 4096
int ustrun_sgd_ema_multi(const ustrun_param_t* table, const int* blk_tensor, const long long* blk_offset, int nblocks,
                         float lr, float momentum, float weight_decay, float alpha, float grad_scale, int do_sgd,
                         int do_ema, void* stream);
/* Same update with the per-step scalars read from DEVICE memory: hyper = float[3] {lr, alpha, grad_scale}
 * (train.py:854-856 poly lr, :91 EMA alpha).  A captured CUDA graph of the step replays with new values
 * after a 12-byte copy; nothing in the launch arguments changes from step to step. */
int ustrun_sgd_ema_multi_dev(const ustrun_param_t* table, const int* blk_tensor, const long long* blk_offset, int nblocks,
                             const float* hyper, float momentum, float weight_decay, int do_sgd, int do_ema, void* stream);

/* ---- frequency-domain style mix -------------------------------------------
 * Replaces the per-sample host loop of train.py:628-636 (extract_amp_spectrum train.py:158-165,
 * low_freq_mutate_np :167-187, source_to_target_freq :189-207): out[n] = clip(Re ifft2(amp'(n) * exp(i*phase_src)), 0, 255)
 * / 127.5 - 1 with amp' = amp_src*(1-ratio[n]) + amp_trg*ratio[n] inside the centred window of half-width
 * b = floor(min(H,W)*L) and amp_src outside.  src (the CutMix partner, mix_img), trg (ulb_x_w), out: fp32 NCHW in
 * [-1,1]; ratio: one float64 per sample (the host's random.uniform(0, iter/max_iter), train.py:180).  float64
 * arithmetic on the device; 2b+1 <= 25.  workspace: ustrun_fft_amp_mix_workspace_bytes(). */
long long ustrun_fft_amp_mix_workspace_bytes(int N, int C, int H, int W, double L);
int ustrun_fft_amp_mix(const float* src, const float* trg, const double* ratio, double L, float* out, int N, int C, int H, int W,
                       void* workspace, long long ws_bytes, void* stream);

/* ---- hardness of the unlabelled samples -----------------------------------
 * Replaces train.py:705-718 (stu_pseudo_label.cpu() / pseudo_label.cpu() -> utils/metrics.py dice_coeff :149-174,
 * dice_coeff_2label :176-201, dice_coeff_3label :203-231 -> dice_coefficient_numpy :114-146): per sample
 * dice_p = (2I+1)/(1.001+S+G) (0 if S = G = 0) over `parts`, hardness[b] = 1 - mean_p dice_p (1 in the first epoch),
 * lq_idx = first argmax.  stu_pl / tea_pl: uint8 label planes from ustrun_pseudo_label_* ([B,H,W]; mode 1: [B,2,H,W]).
 * mode 0: label != 0 (prostate, BUSI); 1: two sigmoid channels (fundus); 2: classes 1..3 (M&Ms).
 * workspace: unsigned int[B*3*3]; hardness: double[B]; dice: double[parts][B] or NULL; lq_idx: int[1]. */
int ustrun_hardness(const unsigned char* stu_pl, const unsigned char* tea_pl, int B, int H, int W, int mode, int first_epoch,
                    unsigned int* workspace, double* hardness, double* dice, int* lq_idx, void* stream);

/* ---- confidence bank / low-quality sample bookkeeping (second half) --------------
 * Replaces the host-side numpy / torch.cat code of train.py:754-781 (bank FIFO + adaptive threshold), :612-626 (CutMix partner
 * pool and choice), :741-743 (hardest sample) and obtain_all_cover_box :242-251.  The bank length (int) and choice_th (double)
 * are DEVICE scalars; nothing is copied to the host.
 * ustrun_bank_update: simple_ulb_idx = hardness < choice_th; new bank = selected samples of this batch (batch order) followed by
 *   the first `newlen` old entries (newlen = max_len - cur if n + cur > max_len else n); choice_th = min(choice_th, max hardness
 *   in the new bank) if something was selected, else min(increase * choice_th, 0.1) (unchanged while the bank is empty).
 *   old_* / new_* are the two copies of the bank storage (max_len slots of img_elems floats / lab_elems bytes); new_hard /
 *   old_hard double[max_len]; plan_ws int[max_len + 1].  Requires Bu <= max_len.
 * ustrun_bank_choice: choice[i] (train.py:615,621-625) from host draws r_lb[i] in [0, Bl), r_u[i] in [0, 1), perm (a permutation
 *   of 0..Bu-1) and the device-side bank length: first Bu - k partners from the labelled batch, k = min(Bu / 2, n) from the bank
 *   (index Bl + floor(r_u * n)), permuted.
 * ustrun_lq_select: copies sample *lq_idx (device int, from ustrun_hardness) of the batch into the lq buffers.
 * ustrun_cover_box: box[H][W] = 1 inside the bounding box of the union of the non-zero pixels of up to four uint8 planes
 *   (nullable); if the union is empty the box is `fallback` (a host-drawn CutMix rectangle, train.py:245) or all zero. */
int ustrun_bank_update(const double* hardness, int Bu, const float* batch_img, const unsigned char* batch_pl, const unsigned char* batch_mask,
                       const float* old_img, const unsigned char* old_pl, const unsigned char* old_mask, const double* old_hard, float* new_img,
                       unsigned char* new_pl, unsigned char* new_mask, double* new_hard, int* n_state, double* th_state, int max_len, double increase,
                       long long img_elems, long long lab_elems, int* plan_ws, void* stream);
int ustrun_bank_choice(const int* n_state, int Bl, int Bu, const int* r_lb, const double* r_u, const int* perm, int* choice, void* stream);
int ustrun_lq_select(const int* lq_idx, const float* img, const unsigned char* pl, const unsigned char* mask, float* out_img, unsigned char* out_pl,
                     unsigned char* out_mask, long long img_elems, long long lab_elems, void* stream);
int ustrun_cover_box(const unsigned char* p0, const unsigned char* p1, const unsigned char* p2, const unsigned char* p3, int H, int W,
                     const unsigned char* fallback, unsigned char* box, void* stream);

/* ---- evaluation helpers --------------------------------------------------
 * ustrun_encode_labels: the label encodings of train.py:590-608 / :281-288 and train_mnms.py:549-556 on the float32
 *   label image the data loader yields.  mode 0: y == 0 (prostate); 1: y == 255 (BUSI); 2: two planes {y == 0, y <= 128}
 *   (fundus) -> out [B,2,H,W]; 3: y [B,H,W,3], label 1/2/3 where channel 0/1/2 == 255 (M&Ms).  out: uint8.
 * ustrun_predict: train.py:295-302: softmax branch argmax of the softmax PROBABILITIES (first maximum) -> [B,H,W];
 *   sigmoid branch sigmoid(x) >= 0.5 -> [B,C,H,W].  logits fp32 NCHW, pred uint8.
 * ustrun_seg_metrics: per label part, batch means of the reference's Dice (utils/metrics.py:114-146 via dice_coeff*,
 *   train.py:303) and of medpy.metric.binary dc / jc (train.py:307-311; medpy is not vendored: dc = 2I/(S+G), 0 if empty;
 *   jc = I/|union|, 0 where medpy raises on an empty union).  mode as in ustrun_hardness.  workspace: unsigned int[B*9];
 *   out: double[3][parts] = {dice, dc, jc}.  hd95 / asd stay on the host (medpy). */
int ustrun_encode_labels(const float* y, int mode, int B, int H, int W, unsigned char* out, void* stream);
int ustrun_predict(const float* logits, int sigmoid, int B, int C, int H, int W, unsigned char* pred, void* stream);
int ustrun_seg_metrics(const unsigned char* pred, const unsigned char* target, int B, int H, int W, int mode, unsigned int* workspace,
                       double* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* USTRUN_H_ */
