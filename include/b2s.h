/* libb2s — C ABI of the B200 (sm_100a) UNet segmentation hot path.
 *
 * The reference (WuJiaqiii/Thyroid-nodule-image-segmentation-UNet-DDTI) has no FFI of its own: its hot path is
 * the set of torch.nn call sites in models/model.py, models/vnet.py, models/loss.py and utils/trainer.py. Each
 * entry point below names the reference call site (file:line) whose arithmetic it replaces. The Python side
 * (package models/model.py, models/loss.py) binds these with ctypes; see INTEGRATION.md.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer on the current CUDA device unless stated otherwise; caller owns all
 *     buffers (outputs, saved tensors, workspaces); the library never allocates and never synchronises.
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued on it and the call returns immediately.
 *   - activations are NHWC bf16; `*_cstride` is the pixel stride in ELEMENTS of the buffer the pointer lives
 *     in (>= the channel count when the pointer addresses a channel slice of a wider concat buffer).
 *   - return value: 0 = ok, negative = error (B2S_ERR_*); b2s_last_error() returns a thread-local message.
 */
#ifndef B2S_H_
#define B2S_H_

#ifdef __cplusplus
extern "C" {
#endif

#define B2S_OK 0
#define B2S_ERR_ARG (-1)
#define B2S_ERR_CUDA (-2)

#define B2S_FLAG_RELU 1   /* apply max(x,0) in the conv epilogue            (models/model.py:37,40) */
#define B2S_FLAG_STATS 2  /* emit per-channel sum / sum-of-squares partials (BatchNorm2d, model.py:38,41) */
#define B2S_FLAG_BNRED 4  /* with STATS on an input-gradient launch: second partial = sum(dy * r) (BatchNorm backward) */

const char* b2s_last_error(void);
long long b2s_launch_count(void); /* kernels launched by this library in this process (bench.py gpu_launches) */
int b2s_version(void);

/* ---- tensor-core convolutions (tcgen05 + TMA + TMEM) ------------------------------------------------------ */

/* nn.Conv2d(Cin,Cout,3,padding=1) / nn.Conv2d(Cin,Cout,1) forward (models/model.py:36,39; models/vnet.py:43,59)
 * and, with rotated weights from b2s_pack_conv_weight, the 3x3 input gradient (autograd of the same call sites).
 *   x [N,H,W,Cin] bf16, w_packed [ksize*ksize][Cout][Cin] bf16, bias [Cout] fp32 or NULL,
 *   y [N,H,W,Cout] bf16, stats_partial [b2s_conv_stats_rows(N,H,W,Cout,tile_n)][2][Cout] fp32 when B2S_FLAG_STATS
 *   (two rows per CTA row-group of the persistent grid; at most 2 * SM count rows).
 *   tile_n: 0 = auto, else 64/128/256 (must divide Cout); the bits above bit 9 force a kernel variant (tests):
 *   +1024 tile-pair kernel, +2048 row-halo kernel (W % 128 == 0, even H, ksize 3), +4096 single-tile kernel.
 *   Cin, Cout multiples of 64. */
int b2s_conv_fwd(const void* x, int x_cstride, const void* w_packed, const float* bias, void* y, int y_cstride,
                 float* stats_partial, int N, int H, int W, int Cin, int Cout, int ksize, int flags, int tile_n,
                 void* stream);
int b2s_conv_stats_rows(int N, int H, int W, int Cout, int tile_n);
/* inference variant: y = act(conv(x) + bias) * post_scale + post_shift, i.e. the eval-mode BatchNorm2d
 * (models/model.py:38,41 with running statistics, b2s_bn_eval_affine) applied in the conv epilogue. */
int b2s_conv_fwd_affine(const void* x, int x_cstride, const void* w_packed, const float* bias, const float* post_scale,
                        const float* post_shift, void* y, int y_cstride, int N, int H, int W, int Cin, int Cout,
                        int ksize, int flags, int tile_n, void* stream);

/* nn.ConvTranspose2d(Cin,Cout,2,stride=2) forward (models/model.py:19,49): x [N,Hi,Wi,Cin] -> y [N,2Hi,2Wi,Cout];
 * w_packed [(a*2+b)*Cout+co][Cin] bf16 (b2s_pack_convt_weight). */
int b2s_convt2x2_fwd(const void* x, int x_cstride, const void* w_packed, const float* bias, void* y, int y_cstride,
                     int N, int Hi, int Wi, int Cin, int Cout, int tile_n, void* stream);
/* its input gradient: dy [N,2Hi,2Wi,Cout] -> dx [N,Hi,Wi,Cin]; w_packed [(a*2+b)*Cin+ci][Cout] bf16. */
int b2s_convt2x2_dgrad(const void* dy, int dy_cstride, const void* w_packed, void* dx, int dx_cstride, int N, int Hi,
                       int Wi, int Cin, int Cout, int tile_n, void* stream);

/* Weight gradients (autograd of the conv call sites): split-K fp32 partials ws[split][tap*Cin+ci][co], reduced and
 * re-laid-out by b2s_wgrad_reduce. taps: 9 for the 3x3 conv (pass ksize_or_taps = 3), 4 for the transposed conv. */
long long b2s_conv_wgrad_workspace(int N, int H, int W, int Cin, int Cout, int ksize_or_taps, int tile_n, int splits,
                                   int* splits_out);
int b2s_conv3x3_wgrad(const void* x, int x_cstride, const void* dz, int dz_cstride, float* ws, int N, int H, int W,
                      int Cin, int Cout, int tile_n, int splits, void* stream);
int b2s_convt2x2_wgrad(const void* x, int x_cstride, const void* dy, int dy_cstride, float* ws, int N, int Hi, int Wi,
                       int Cin, int Cout, int tile_n, int splits, void* stream);
/* layout 0: dw[co][ci][tap] (Conv2d OIHW); layout 1: dw[ci][co][tap] (ConvTranspose2d IOHW). taps: 9, 4 or 1. */
int b2s_wgrad_reduce(const float* ws, int splits, int taps, int Cin, int Cout, float* dw, int layout, void* stream);

/* ---- weight packing (fp32 parameters -> bf16 GEMM operands) ---------------------------------------------------- */
/* w [Cout][Cin][k][k] fp32 -> w_fwd [k*k][Cout][Cin] bf16 and (optional) w_dgrad [k*k][Cin][Cout] bf16 with the
 * taps reversed, so that the input gradient is a forward conv of dz with w_dgrad. */
int b2s_pack_conv_weight(const float* w, void* w_fwd, void* w_dgrad, int Cout, int Cin, int ksize, void* stream);
/* w [Cin][Cout][2][2] fp32 -> w_fwd [(a*2+b)*Cout+co][Cin], w_dgrad [(a*2+b)*Cin+ci][Cout] (either may be NULL). */
int b2s_pack_convt_weight(const float* w, void* w_fwd, void* w_dgrad, int Cin, int Cout, void* stream);

/* every conv / transposed-conv weight of a network in ONE launch (arrays in HOST memory, n <= 40): kind[i] 0 = Conv2d
 * weight (d0 = Cout, d1 = Cin, taps = k*k in {1, 9}), 1 = ConvTranspose2d weight (d0 = Cin, d1 = Cout, taps = 4); outputs
 * laid out as by b2s_pack_conv_weight / b2s_pack_convt_weight; wf[i] or wd[i] may be NULL. */
int b2s_pack_weights_all(int n, const void* const* w, void* const* wf, void* const* wd, const int* d0, const int* d1,
                         const int* taps, const int* kind, void* stream);

/* ---- bandwidth kernels ------------------------------------------------------------------------------------------ */
/* first layer nn.Conv2d(1,Cout,3,padding=1)+ReLU (models/model.py:10): x [N,H,W] fp32 -> r [N,H,W,Cout] bf16.
 * stats_partial [b2s_c1_rows(N,H,W)][2][Cout]. Cout multiple of 8, <= 128. */
int b2s_conv3x3_c1_fwd(const float* x, const float* w, const float* bias, void* r, float* stats_partial, int N, int H,
                       int W, int Cout, int flags, void* stream);
int b2s_c1_rows(int N, int H, int W);
/* inference variant: y = act(conv(x) + bias) * post_scale + post_shift (eval-mode BatchNorm in the same pass). */
int b2s_conv3x3_c1_fwd_affine(const float* x, const float* w, const float* bias, const float* post_scale,
                              const float* post_shift, void* y, int N, int H, int W, int Cout, int flags, void* stream);
/* its weight gradient: partial [b2s_c1_rows][Cout*9] fp32 (reduce with b2s_reduce_rows). */
int b2s_conv3x3_c1_wgrad(const float* x, const void* dz, float* partial, int N, int H, int W, int Cout, void* stream);

/* out[k] = sum_r in[r][k]; scratch >= 128*K floats (used when rows > 1024; the same rule applies to every
 * `scratch` argument below with K = the row width of the partial buffer). Deterministic. */
int b2s_reduce_rows(const float* in, int rows, int K, float* scratch, float* out, void* stream);

/* nn.BatchNorm2d training statistics (models/model.py:38,41): reduces conv-epilogue partials [rows][2][C] over
 * count = N*H*W elements; writes mean, invstd, scale = gamma*invstd, shift = beta - mean*scale; updates
 * running_mean/var with `momentum` (unbiased variance) and ++num_batches_tracked when the pointers are non-NULL. */
int b2s_bn_finalize(const float* partial, int rows, int C, double count, const float* gamma, const float* beta,
                    float* running_mean, float* running_var, long long* num_batches_tracked, float momentum,
                    float eps, float* scale, float* shift, float* mean, float* invstd, float* scratch, void* stream);
/* eval mode: scale/shift from the running statistics. */
int b2s_bn_eval_affine(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                       float eps, float* scale, float* shift, int C, void* stream);
/* y = r*scale + shift (bf16 out, may be a concat slice); if pooled != NULL also F.max_pool2d(y, 2)
 * (models/model.py:17,56-58) into the dense tensor pooled [N,H/2,W/2,C]. */
int b2s_bn_apply(const void* r, int r_cstride, const float* scale, const float* shift, void* y, int y_cstride,
                 void* pooled, int N, int H, int W, int C, void* stream);
/* F.max_pool2d(x, 2) on its own (inference path): x [N,H,W,C] -> pooled [N,H/2,W/2,C] dense. */
int b2s_maxpool2x2(const void* x, int x_cstride, void* pooled, int N, int H, int W, int C, void* stream);
int b2s_ew_rows(void); /* partial rows written by the element-wise reduction kernels below */
/* BatchNorm+ReLU backward, pass 1: partial [b2s_ew_rows][2][C] = sum dy, sum dy*xhat, where
 * dy = dy_in (+ max-pool routed dpool when dpool != NULL; first maximum in row-major window order). */
int b2s_bn_bwd_reduce(const void* dy, int dy_cstride, const void* dpool, const void* r, int r_cstride,
                      const float* scale, const float* shift, const float* mean, const float* invstd, float* partial,
                      int N, int H, int W, int C, void* stream);
/* pass 1 finalize: dgamma, dbeta and coef [3][C] = {gamma*invstd, mean(dy), mean(dy*xhat)}. */
int b2s_bn_bwd_finalize(const float* partial, int rows, int C, double count, const float* gamma, const float* invstd,
                        float* dgamma, float* dbeta, float* coef, float* scratch, void* stream);
/* pass 2: dz = (r>0) * coef0 * (dy - coef1 - xhat*coef2) (bf16) and conv-bias gradient partials [b2s_ew_rows][C]. */
int b2s_bn_bwd_apply(const void* dy, int dy_cstride, const void* dpool, const void* r, int r_cstride,
                     const float* scale, const float* shift, const float* mean, const float* invstd, const float* coef,
                     void* dz, int dz_cstride, float* dbias_partial, int N, int H, int W, int C, void* stream);

/* head: BatchNorm apply folded into nn.Conv2d(C,O,1) (models/model.py:30): logits [N,O,H,W] fp32 from r [N,H,W,C];
 * mask (optional, uint8 [N,O,H,W]) = sigmoid(logit) > 0.5 evaluated in fp32 (utils/trainer.py:101,152,217). */
int b2s_head_fwd(const void* r, int r_cstride, const float* scale, const float* shift, const float* w, const float* b,
                 float* logits, unsigned char* mask, int N, long long HW, int C, int O, void* stream);
/* head backward: dy [N,H,W,C] bf16 (gradient w.r.t. the BN output), partial [b2s_ew_rows][O*C+O] = dW, db. */
int b2s_head_bwd(const float* dlogits, const void* r, int r_cstride, const float* scale, const float* shift,
                 const float* w, void* dy, int dy_cstride, float* partial, int N, long long HW, int C, int O,
                 void* stream);

/* Dice + BCE (+ FocalTversky) loss (models/loss.py:13-24,34-46; nn.BCEWithLogitsLoss utils/trainer.py:37):
 * sums [B][4] = {sum p*t, sum p, sum t, sum bce}; out [8] = {total, bce, dice, focal_tversky, TP, sum p, sum t, -};
 * total = w_bce*bce + w_dice*dice + w_ft*ft; out+4 is the batch-global triple b2s_seg_loss_bwd takes as ft_tot. partial: [B][b2s_loss_chunks(per_sample)][4] scratch. */
int b2s_loss_chunks(long long per_sample);
int b2s_seg_loss_fwd(const float* logits, const float* targets, int B, long long per_sample, float* partial,
                     float* sums, float* out, float dice_smooth, float w_bce, float w_dice, float w_ft, float ft_alpha,
                     float ft_beta, float ft_gamma, float ft_smooth, void* stream);
/* dlogits = grad_out * d total / d logits; grad_out: device scalar or NULL (= 1). ft_tot [3] = batch-global
 * {TP, sum p, sum t} (all-reduced by the caller under data parallelism) or NULL to derive from sums. */
int b2s_seg_loss_bwd(const float* logits, const float* targets, const float* sums, const float* ft_tot, int B,
                     long long per_sample, long long bce_count, int dice_batch, const float* grad_out, float* dlogits,
                     float dice_smooth, float w_bce, float w_dice, float w_ft, float ft_alpha, float ft_beta,
                     float ft_gamma, float ft_smooth, void* stream);

/* Segmentation metric counters without a per-step device->host copy (utils/trainer.py:101-107,236-250 with
 * utils/utils.py:225-251): pred = sigmoid(logit) > 0.5 (fp32). ADDS to counters (int64 [7], device) =
 * {TP, FP, FN, TN against target.astype(int), intersection, union against target.astype(bool), elements}.
 * partial: uint32 [b2s_metrics_blocks(n)][6] scratch. */
int b2s_metrics_blocks(long long n);
int b2s_seg_metrics(const float* logits, const float* targets, long long n, unsigned int* partial, long long* counters,
                    void* stream);

/* torch.optim.AdamW step (utils/trainer.py:41,92) over a flat fp32 parameter/gradient bucket;
 * grad_scale multiplies g first (1/world_size after a sum all-reduce). step is 1-based. */
int b2s_adamw_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                   float eps, float weight_decay, int step, float grad_scale, void* stream);

/* same update with hyper = {lr, beta1, beta2, eps, weight_decay, 1-beta1^step, sqrt(1-beta2^step), grad_scale} read
 * from DEVICE memory: lets a captured CUDA graph of the whole step be replayed while lr / step change. */
int b2s_adamw_step_dev(float* p, const float* g, float* m, float* v, long long n, const float* hyper, void* stream);

/* torch.cat along channels for API-visible concat (models/model.py:64-70): strided channel-slice copy. */
int b2s_copy_channels(const void* src, int src_cstride, void* dst, int dst_cstride, long long npix, int C,
                      void* stream);

/* ---- V-Net variant (models/vnet.py) ------------------------------------------------------------------------- */

/* nn.Conv2d(C,2C,3,stride=2,padding=1) forward (models/vnet.py:97): x [N,H,W,Cin] -> y [N,H/2,W/2,Cout] (+bias);
 * the A tiles are dense TMA boxes over one pixel-parity class of x. Its input and weight gradients are b2s_conv_fwd (rotated
 * weights) and b2s_conv3x3_wgrad applied to the zero-inserted output gradient from b2s_upsample_zero2x. */
int b2s_conv3x3_s2_fwd(const void* x, int x_cstride, const void* w_packed, const float* bias, void* y, int y_cstride,
                       int N, int H, int W, int Cin, int Cout, int tile_n, void* stream);
/* its input gradient without zero insertion: four launches, one per output parity class (1/2/2/4 taps each), written
 * to the four sub-lattices of dx [N,H,W,Cin] through a 5-D TMA store; w_dgrad_packed [9][Cin][Cout] from
 * b2s_pack_conv_weight. Returns 1 (no error, nothing launched) for images too small for that store: use the
 * zero-insertion path then. */
int b2s_conv3x3_s2_dgrad(const void* dz, int dz_cstride, const void* w_dgrad_packed, void* dx, int dx_cstride, int N,
                         int H, int W, int Cin, int Cout, int tile_n, void* stream);
/* its weight gradient: x read as parity-class boxes (no zero insertion); workspace / splits from
 * b2s_conv_wgrad_workspace(N, H/2, W/2, Cin, Cout, 3, tile_n | 4096, splits, &s). */
int b2s_conv3x3_s2_wgrad(const void* x, int x_cstride, const void* dz, int dz_cstride, float* ws, int N, int H, int W,
                         int Cin, int Cout, int tile_n, int splits, void* stream);
/* dst [N,2Hs,2Ws,C]: dst[n,2i,2j,:] = src[n,i,j,:], zero elsewhere. */
int b2s_upsample_zero2x(const void* src, int src_cstride, void* dst, int dst_cstride, int N, int Hs, int Ws, int C,
                        void* stream);
/* nn.Conv2d(Cin,Cout,1) weight gradient (residual projections, models/vnet.py:46,58): ws[split][ci][co]; workspace
 * from b2s_conv_wgrad_workspace(..., ksize_or_taps = 1, ...), reduced by b2s_wgrad_reduce(taps = 1, layout 0). */
int b2s_conv1x1_wgrad(const void* x, int x_cstride, const void* dz, int dz_cstride, float* ws, int N, int H, int W,
                      int Cin, int Cout, int tile_n, int splits, void* stream);

/* BatchNorm -> ReLU -> Dropout (+ residual) of ConvBlock (models/vnet.py:51-59):
 * out = dropout_p(relu(z*scale + shift)) + res; scale/shift/res may be NULL, relu is a flag. The dropout mask is a
 * counter-based hash of (seed, NHWC element index): P(keep) = 1-p, kept values scaled by 1/(1-p); p = 0 disables it.
 * step_counter (optional DEVICE int64, e.g. the layer's BatchNorm num_batches_tracked) is mixed into the seed at run time,
 * so a captured CUDA graph draws a new mask at every replay; forward and backward of one step must see the same value. */
int b2s_bn_act_apply(const void* z, int z_cstride, const float* scale, const float* shift, const void* res,
                     int res_cstride, void* out, int out_cstride, long long npix, int C, int relu, float dropout_p,
                     unsigned seed, const long long* step_counter, void* stream);
/* its backward in two passes (same finalize as the UNet path, b2s_bn_bwd_finalize): pass 1 partial [b2s_ew_rows][2][C]
 * = sum dy, sum dy*xhat with dy = da * mask/(1-p) * (z*scale+shift > 0); pass 2 dz = c0 (dy - c1 - xhat c2) and the
 * conv-bias gradient partials [b2s_ew_rows][C]. */
int b2s_bn_act_bwd_reduce(const void* da, int da_cstride, const void* z, int z_cstride, const float* scale,
                          const float* shift, const float* mean, const float* invstd, float* partial, long long npix,
                          int C, int relu, float dropout_p, unsigned seed, const long long* step_counter, void* stream);
int b2s_bn_act_bwd_apply(const void* da, int da_cstride, const void* z, int z_cstride, const float* scale,
                         const float* shift, const float* mean, const float* invstd, const float* coef, void* dz,
                         int dz_cstride, float* dbias_partial, long long npix, int C, int relu, float dropout_p,
                         unsigned seed, const long long* step_counter, void* stream);
/* models/mod.py variants (Conv -> BN -> ReLU blocks, ResidualBlock: relu(BN(conv) + skip), models/mod.py:43-51,71-84):
 * b2s_bn_act_apply's `relu` argument is a mode: 0 none, 1 ReLU before the residual add (V-Net ConvBlock), 2 ReLU after it
 * (ResidualBlock). b2s_relu_bwd: dx = dy * (y > 0) from the ReLU OUTPUT y. b2s_maxpool2x2_bwd: nn.MaxPool2d(2,2) backward,
 * gradient to the first maximum of each window (x = the pooling input), dx dense. */
int b2s_relu_bwd(const void* dy, int dy_cstride, const void* y, int y_cstride, void* dx, int dx_cstride, long long npix,
                 int C, void* stream);
int b2s_maxpool2x2_bwd(const void* x, int x_cstride, const void* dpool, int dpool_cstride, void* dx, int dx_cstride, int N,
                       int H, int W, int C, void* stream);
/* partial [b2s_ew_rows][C] = per-channel sums over pixels (bias gradient of a conv that is not followed by BN). */
int b2s_channel_sums(const void* x, int x_cstride, float* partial, long long npix, int C, void* stream);

/* SEBlock (models/vnet.py:5-26). b2s_se_pool: partial [N][b2s_se_chunks(HW)][C] = sums over pixels of x (y NULL) or
 * of x*y; b2s_se_fc_fwd: mean = pooled/HW, hidden = relu(W1 mean + b1), gate = sigmoid(W2 hidden + b2), all fp32
 * [N][C] / [N][C/r]; b2s_se_scale: y = x*gate[n][c] (+ add[n][c]*add_scale when add != NULL: the backward's
 * dx = dy*gate + dmean/HW); b2s_se_fc_bwd: from partial = sum_hw dy*x: ds, dh, dmean per sample and the batch-summed
 * dW1 [C/r][C], db1, dW2 [C][C/r], db2. */
int b2s_se_chunks(long long HW);
int b2s_se_pool(const void* x, int x_cstride, const void* y, int y_cstride, float* partial, int N, long long HW, int C,
                void* stream);
int b2s_se_fc_fwd(const float* partial, int chunks, long long HW, const float* w1, const float* b1, const float* w2,
                  const float* b2, float* mean, float* hidden, float* gate, int N, int C, int Cr, void* stream);
int b2s_se_scale(const void* x, int x_cstride, const float* gate, const float* add, float add_scale, void* y,
                 int y_cstride, int N, long long HW, int C, void* stream);
int b2s_se_fc_bwd(const float* partial, int chunks, const float* gate, const float* hidden, const float* mean,
                  const float* w1, const float* w2, float* ds, float* dh, float* dmean, float* dw1, float* db1,
                  float* dw2, float* db2, int N, int C, int Cr, void* stream);

/* BatchNorm-backward reduction fused into the launch that PRODUCES dy (autograd of models/model.py:38,41 behind a conv /
 * transposed conv): input-gradient launches whose output dy is the gradient of a train-mode BatchNorm output also emit
 * partial [rows][2][C] = {sum dy, sum dy * r} per channel (r = that BatchNorm's saved input, same pixels / channels as
 * dy; rows = b2s_conv_stats_rows / b2s_convt2x2_dgrad_rows), so the separate b2s_bn_bwd_reduce pass over dy and r is not
 * needed; b2s_bn_bwd_finalize_raw turns them into dgamma, dbeta and the coefficients b2s_bn_bwd_apply expects. Both
 * *_bnred entry points return 1 (nothing launched) when the shape takes the one-tile fallback kernel. */
int b2s_conv_dgrad_bnred(const void* dz, int dz_cstride, const void* w_dgrad_packed, void* dy, int dy_cstride,
                         const void* r, int r_cstride, float* partial, int N, int H, int W, int Cin, int Cout, int ksize,
                         int tile_n, void* stream);
int b2s_convt2x2_dgrad_bnred(const void* dy, int dy_cstride, const void* w_packed, void* dx, int dx_cstride, const void* r,
                             int r_cstride, float* partial, int N, int Hi, int Wi, int Cin, int Cout, int tile_n,
                             void* stream);
int b2s_convt2x2_dgrad_rows(int N, int Hi, int Wi, int Cin, int tile_n);
int b2s_bn_bwd_finalize_raw(const float* partial, int rows, int C, double count, const float* gamma, const float* mean,
                            const float* invstd, float* dgamma, float* dbeta, float* coef, float* scratch, void* stream);

/* AttentionGate (models/mod.py:211-234): psi = sigmoid(BatchNorm2d(1)(conv1x1(relu(g1 + x1)))), out = x * psi.
 * The F_int -> 1 conv is b2s_head_fwd over channel slices of at most 256 channels (up to four fp32 maps, summed on
 * the fly by the kernels below). One-channel BatchNorm + sigmoid over n = N*H*W values:
 *   b2s_psi_stats       partial [b2s_psi_rows(n)][2] = {sum v, sum v^2}; finalise with b2s_bn_finalize(C = 1)
 *   b2s_psi_fwd         psi = sigmoid(v * scale[0] + shift[0])
 *   b2s_psi_bwd_reduce  partial [b2s_psi_rows(n)][2] = {sum g, sum g*xhat}, g = dpsi*psi*(1-psi); finalise with
 *                       b2s_bn_bwd_finalize(C = 1) -> dgamma, dbeta, coef[3]
 *   b2s_psi_bwd_apply   dv = coef0 * (g - coef1 - xhat*coef2)   (the gradient of every input map)
 * b2s_pixel_scale_fwd: out[p,c] = x[p,c]*psi[p]; b2s_pixel_scale_bwd: dx = dy*psi, dpsi[p] = sum_c dy[p,c]*x[p,c]. */
int b2s_psi_rows(long long n);
int b2s_psi_stats(const float* const* maps, int n_maps, long long n, float* partial, void* stream);
int b2s_psi_fwd(const float* const* maps, int n_maps, const float* scale, const float* shift, float* psi, long long n,
                void* stream);
int b2s_psi_bwd_reduce(const float* const* maps, int n_maps, const float* psi, const float* dpsi, const float* mean,
                       const float* invstd, long long n, float* partial, void* stream);
int b2s_psi_bwd_apply(const float* const* maps, int n_maps, const float* psi, const float* dpsi, const float* mean,
                      const float* invstd, const float* coef, float* dv, long long n, void* stream);
int b2s_pixel_scale_fwd(const void* x, int x_cstride, const float* psi, void* out, int out_cstride, long long npix, int C,
                        void* stream);
int b2s_pixel_scale_bwd(const void* x, int x_cstride, const float* psi, const void* dy, int dy_cstride, void* dx,
                        int dx_cstride, float* dpsi, long long npix, int C, void* stream);
/* F.interpolate(x, size=(Ho,Wo), mode='bilinear', align_corners=False) on NHWC bf16 (models/mod.py:61-62,126-127,
 * 289-290: taken when H or W is not a multiple of 2^depth) and its input gradient (deterministic gather). */
int b2s_bilinear_fwd(const void* x, int x_cstride, void* y, int y_cstride, int N, int Hi, int Wi, int Ho, int Wo, int C,
                     void* stream);
int b2s_bilinear_bwd(const void* dy, int dy_cstride, void* dx, int dx_cstride, int N, int Hi, int Wi, int Ho, int Wo,
                     int C, void* stream);
/* Multi-channel input images (UNet(in_channels > 1), models/model.py:6,10): x [N,C,H*W] fp32 -> y [N,H*W,Cpad] bf16 with
 * channels >= C zero; the first conv then takes the tensor-core path with its weight zero-padded to Cpad inputs. */
int b2s_image_to_nhwc(const float* x, void* y, int N, int C, long long HW, int Cpad, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B2S_H_ */
