/* ptivae.h -- C ABI of libptivae.so: the B200 (sm_100a) kernels behind the PTI-LDM-VAE hot path.
 *
 * The reference (Sukikui/PTI-LDM-VAE) is pure Python: its hot path is
 *   pti_ldm_vae.models.VAEModel            /root/reference/src/pti_ldm_vae/models/autoencoder.py:6-171
 *     -> monai.networks.nets.AutoencoderKL /root/reference/src/pti_ldm_vae/models/autoencoder.py:67-79
 * i.e. there is no FFI in the reference; what this library replaces are the ATen/cuDNN/cuBLAS calls
 * MONAI 1.5.1 issues for that path.  Each entry point below names the reference operator it stands in
 * for.  The Python binding a maintainer adds is the ctypes stub in INTEGRATION.md (mirrored by
 * pti-ldm-vae_b200/_lib.py).
 *
 * Conventions
 *   - all pointers are DEVICE pointers on the current CUDA device unless stated otherwise;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *   - nothing allocates, nothing synchronises; the caller owns every buffer;
 *   - return value: 0 = launched; <0 = argument/shape error (no launch happened):
 *        -1 bad argument, -2 unsupported shape, -3 driver entry point / tensor-map encode failure;
 *     >0 = the cudaError_t reported by the launch;
 *   - activations: NHWC [N][H][W][C].  GEMM operands are 16-bit ("h16"): IEEE fp16 when the call's
 *     f16 flag is non-zero (default of the Python host: 8x finer rounding than bf16 at the same
 *     tensor-core rate; values are clamped to +-65504), bfloat16 otherwise.  The residual stream between
 *     blocks is fp32.  Where a tensor may be any of the three, `fmt` is 0 = bf16, 1 = fp16, 2 = fp32.
 *   - latents / images at the network boundary: fp32 NCHW.
 */
#ifndef PTIVAE_H_
#define PTIVAE_H_

#ifdef __cplusplus
extern "C" {
#endif

/* Library / build identification: returns e.g. 100 for sm_100a builds, and the ABI revision. */
int ptivae_abi_version(void);

/* nn.Conv2d / nn.Linear on tensor cores (tcgen05 implicit GEMM).
 *   replaces: monai Convolution(conv_only) 3x3 s1 p1, AEKLDownsample (F.pad(0,1,0,1)+3x3 s2),
 *             UpSample(nearest x2)+postconv 3x3, nin_shortcut 1x1, SABlock.to_q/to_k/to_v/out_proj
 *             (monai 1.5.1 networks/nets/autoencoderkl.py, blocks/selfattention.py; called from
 *             autoencoder.py:114/:139/:151).
 *   in        h16 [N][H][W][Cin]          Cin in {32, 64k}
 *   w_packed  h16 [T][Cout][Cin]          from ptivae_pack_conv_weight (T = 9, 1, or 16 for mode 2)
 *   bias      fp32 [Cout]
 *   residual  same shape as out, h16 (res_f32 = 0) or fp32 (res_f32 = 1), added in the epilogue; may be NULL
 *   out       [N][Hout][Wout][Cout], h16 (out_f32 = 0) or fp32 (out_f32 = 1: the residual stream)
 *   out16     optional (NULL ok): when out is fp32, an additional h16 copy of it (the operand a following
 *             nin_shortcut 1x1 conv reads)
 *   gn_part   fp32 [N][P][gn_groups][2]: per-tile (sum, sum of squares) of the stored output, P =
 *             ptivae_conv_parts(H, W, mode); plain stores, fixed order (deterministic); ignored when
 *             gn_groups == 0.  Feed to ptivae_gn_finalize(..., P, ...).
 *   mode      0: 3x3 stride 1 pad 1 (Hout=H)      1: pad right/bottom + 3x3 stride 2 (Hout=H/2, H,W even)
 *             2: nearest x2 upsample + 3x3 pad 1 (Hout=2H)   3: 1x1
 *             data gradients (autograd's conv backward-input under train_vae.py:444), `in` = dY with the extent
 *             (H, W) of dY, Cin = channels of dY, w_packed = the TRANSPOSED pack [T][Cin_fwd][Cout_fwd]
 *             (ptivae_pack_conv_weight mode | 4):
 *             4: of mode 0 (Hout=H)   5: of mode 1 (Hout=2H)   6: of mode 2 (Hout=H/2, w_packed = transposed 16-slab pack)
 *             (the data gradient of mode 3 is mode 3 with the transposed pack) */
int ptivae_conv_umma(const void* in, const void* w_packed, const float* bias, const void* residual, void* out,
                     void* out16, float* gn_part, int gn_groups, int N, int H, int W, int Cin, int Cout, int mode, int out_f32,
                     int res_f32, int f16, void* stream);
/* number of statistics partials per image that ptivae_conv_umma writes for this shape */
int ptivae_conv_parts(int H, int W, int mode);

/* Fused ResBlock convolution: out = conv3x3_s1_p1( act(x*scale + shift) ) + bias (+ residual), tcgen05,
 * halo-resident: the normalised tensor never exists in HBM and every input byte is read once per tile.
 *   replaces: AEKLResBlock's  conv1(F.silu(norm1(x)))  and  conv2(F.silu(norm2(h))) + shortcut  pairs
 *             (monai 1.5.1 networks/nets/autoencoderkl.py AEKLResBlock.forward; via autoencoder.py:114).
 *   x           NHWC [N][H][W][Cin], storage in_fmt (fp32 stream or the h16 operand format)
 *   scale_shift fp32 [N][Cin][2] from ptivae_gn_finalize, or NULL (identity prologue); silu != 0 -> SiLU.
 *               Zero padding is applied AFTER the normalisation, as nn.Conv2d(padding=1) does.
 *   w_packed    h16 [9][Cout][Cin]; Cin, Cout in {32, 64, 128} (and 256 on the h16 stream, see ptivae_conv3x3_fused_query)
 *   residual/out/gn_part: as ptivae_conv_umma, with P = ptivae_conv3x3_fused_parts(H, W) (16x16 tiles).
 *   impl: 0 = auto; 1 = register-staged kernel (widths <= 128, both operand formats); 2 = TMA-staged kernel
 *         (Cin,Cout <= 64); 3 = chunk-pipelined TMA kernel (all widths); 4 = row-band kernel (Cout = 32, h16 in / out,
 *         none or h16 residual); 5 = two-SM kernel (tcgen05 cta_group::2; Cin, Cout in {128, 256}, h16 in / out, none or
 *         h16 residual).  2-5 need fp16 operands and return -2 for a combination they do not instantiate.  On the fp32
 *         stream all implementations produce the same results up to accumulation order; the h16-stream modes of 3, 4
 *         and 5 evaluate the prologue in packed half2 (h + h*tanh(h), three fp16 roundings): within twice the per-layer
 *         tolerance of the fp32 form (tests/test_gpu_kernels.py), end to end within the stated gates. */
int ptivae_conv3x3_fused(const void* x, int in_fmt, const float* scale_shift, int silu, const void* w_packed,
                         const float* bias, const void* residual, int res_f32, void* out, int out_f32,
                         float* gn_part, int gn_groups, int N, int H, int W, int Cin, int Cout, int f16,
                         int impl, void* stream);
/* Chained launches: the kernels of the forward chain can be launched with programmatic stream serialization, so their
 * set-up (barrier init, TMEM allocation, loads of the static weights) overlaps the drain of the previous kernel on the
 * stream; they then wait (griddepcontrol.wait) before reading or writing anything else.  mask bit 0: the tensor-core
 * kernels (default on), bit 1: the small statistics / direct-conv / latent kernels (default off: measured slower).
 * PTIVAE_CHAIN=<mask> in the environment sets the initial value; returns the previous mask.  Results are bit-identical
 * for every mask. */
int ptivae_set_chained_launch(int mask);
int ptivae_conv3x3_fused_parts(int H, int W);
/* 0 when ptivae_conv3x3_fused has a kernel for (in_fmt, residual kind 0 none | 1 fp32 | 2 h16, out_f32, Cin, Cout, f16), -2
 * otherwise (nothing is launched).  Widths 32/64/128: always; 256 (config B): h16 in, h16 or no residual, h16 out. */
int ptivae_conv3x3_fused_query(int in_fmt, int res_kind, int out_f32, int Cin, int Cout, int f16);
/* conv2 of an AEKLResBlock whose nin_shortcut is a 1x1 conv, with that shortcut fused into the same accumulators:
 *   out = conv3x3_s1_p1( act(h*scale + shift) ) + W_sc * x + bias      (fp32 stream out, no residual tensor)
 *   replaces: `return self.nin_shortcut(x) + h` of AEKLResBlock.forward together with its conv2.
 *   h         h16 [N][H][W][Cin];  sc_x: h16 [N][H][W][sc_cin] = the block input in the operand format
 *   sc_w_packed h16 [1][Cout][sc_cin] (ptivae_pack_conv_weight of the 1x1);  bias = conv2.bias + nin_shortcut.bias
 *   instantiated for (Cin, Cout, sc_cin) in {(32,32,64), (64,64,32)}, fp16 operands; -2 otherwise (callers then
 *   run the shortcut as its own ptivae_conv_umma and pass it as the residual of ptivae_conv3x3_fused) */
int ptivae_conv3x3_fused_sc(const void* h, const float* scale_shift, int silu, const void* w_packed, const float* bias,
                            const void* sc_x, const void* sc_w_packed, int sc_cin, void* out, int out_f32,
                            float* gn_part, int gn_groups, int N, int H, int W, int Cin, int Cout, int f16, void* stream);
/* Nearest x2 upsample + 3x3 conv (pad 1) in one halo-resident kernel: out = conv3x3(upsample2x(x)) + bias.
 *   replaces: monai UpSample(mode="nontrainable", interp_mode="nearest") + its post-conv inside the decoder
 *             (monai 1.5.1 networks/nets/autoencoderkl.py Decoder; via autoencoder.py:151), for the widths where
 *             it beats ptivae_conv_umma mode 2: C in {64, 128}, fp16 operands (anything else returns -2).
 *   x         h16 [N][H][W][C] raw operand (no normalisation: the producer stored the operand format)
 *   w_packed  h16 [16][C][C] from ptivae_pack_conv_weight(mode = 2) (4 phases x 2x2 pre-summed taps)
 *   out       fp32 [N][2H][2W][C] (the residual stream);  out16: optional h16 copy (NULL ok)
 *   gn_part   fp32 [N][P][gn_groups][2] statistics of the stored fp32 values, P = ptivae_up2x_conv3x3_parts(H, W)
 *             (one partial per 16x16 low-res tile); ignored when gn_groups == 0 */
int ptivae_up2x_conv3x3(const void* x, const void* w_packed, const float* bias, float* out, void* out16,
                        float* gn_part, int gn_groups, int N, int H, int W, int C, int f16, void* stream);
int ptivae_up2x_conv3x3_parts(int H, int W);
/* debug only: device buffer (64*32 uint64) that receives CTA 0's per-role clock64 timeline of subsequent
 * ptivae_conv3x3_fused launches; NULL switches tracing off. */
int ptivae_debug_set_trace(void* buf);

/* fp32 master weights [Cout][Cin][k][k] -> h16 UMMA operand [T][Cout][Cin].
 *   mode 0: T = k*k (k in {1,3}); mode 2: T = 16, the 4-phase x (2x2)-tap decomposition of
 *   nearest-x2-upsample + 3x3 (weights of taps that hit the same low-res pixel are pre-summed).
 *   mode 4 / 6: the same slabs transposed, [T][Cin][Cout] (operand of the data-gradient convolutions). */
int ptivae_pack_conv_weight(const float* w, void* out, int Cout, int Cin, int k, int mode, int f16, void* stream);

/* Many weight packs in ONE launch (after an optimizer step every conv / linear weight needs its forward pack and its
 * transposed backward pack again).  descs: DEVICE array of n records of ptivae_pack_desc_bytes() bytes:
 *   { const float* src [Cout][Cin][ksq]; uint16* dst; int64 start (prefix sum of T*Cout*Cin); int32 Cout, Cin, ksq, T;
 *     int64 st_t; int32 st_r, f16, transpose, up2x; }   element (t, r, c) -> dst[t*st_t + r*st_r + c],
 *   (r, c) = (co, ci), or (ci, co) when transpose; up2x selects the 16-slab pre-summed decomposition. */
int ptivae_pack_many(const void* descs, int n, long long total, void* stream);
int ptivae_pack_desc_bytes(void);

/* nn.GroupNorm statistics, deterministic two-stage form.
 *   gn_stats: partial[n][p][g] = (sum, sumsq) over pixel chunk p of x [N][HW][C] (storage in_fmt); P = ptivae_gn_stats_parts(N, HW, C) chunks per image.
 *   gn_finalize: partial [N][P][G][2] -> scale_shift fp32 [N][C][2], scale = gamma*rstd,
 *             shift = beta - mean*scale (biased variance, eps inside the sqrt:
 *             nn.GroupNorm(eps=norm_eps, affine=True)); partials are summed in index order.  mean_rstd (NULL ok):
 *             fp32 [N][G][2] = (mean, rstd) per group, saved for ptivae_gn_bwd.
 *   gn_apply: y(h16) = act(x*scale + shift); act = SiLU when silu != 0 (AEKLResBlock:
 *             F.silu(norm(x))), identity otherwise (SpatialAttentionBlock.norm).  raw16 (may be
 *             NULL) additionally receives x rounded to h16 (the nin_shortcut operand). */
int ptivae_gn_stats(const void* x, float* partial, int N, int HW, int C, int G, int in_fmt, void* stream);
int ptivae_gn_stats_parts(int N, int HW, int C);
int ptivae_gn_finalize(const float* partial, const float* gamma, const float* beta, float* scale_shift,
                       float* mean_rstd, int N, int HW, int C, int G, int P, float eps, void* stream);
int ptivae_gn_apply(const void* x, const float* scale_shift, void* y, void* raw16, int N, int HW, int C, int silu,
                    int in_fmt, int out_f16, void* stream);
/* fp16 RANGE CHECK.  Everything stored in the fp16 operand format is clamped to +-65504 (cvt.rn.satfinite); the reference
 * computes in fp32, so a clamped activation would be a silent parity error.  Every tensor that is stored in 16 bits carries
 * statistics partials (sum, sum of squares per tile and group); a clamped element puts >= 65504^2 into its tile's sum of
 * squares, so "no partial reaches 65504^2" PROVES that nothing was clamped (the converse flags magnitudes within a factor
 * sqrt(tile elements) of the limit -- a warning worth having).  gn_finalize_checked = gn_finalize that also sets
 * range_flag[0] = 1 (device int, never cleared by the library) when a partial trips; range_check does the same for the
 * partials of a tensor that no GroupNorm consumes (pairs = N*P*G).  AutoencoderKL.check_range(x) drives both. */
int ptivae_gn_finalize_checked(const float* partial, const float* gamma, const float* beta, float* scale_shift,
                               float* mean_rstd, int N, int HW, int C, int G, int P, float eps, int* range_flag, void* stream);
int ptivae_range_check(const float* partial, long long pairs, int* range_flag, void* stream);

/* Thin-end 3x3 s1 p1 convolutions on CUDA cores.
 *   small_cin : x fp32 NCHW [N][Cin<=16][H][W], w fp32 [Cout][Cin][3][3] -> out NHWC (storage out_fmt)
 *               (encoder.blocks.0, decoder.blocks.0).  gn_groups > 0: also writes the GroupNorm statistics partials
 *               of the stored values, gn_part fp32 [N][P][gn_groups][2], P = ptivae_conv3x3_small_cin_parts(H, W,
 *               Cout); fused only for Cin == 1 (returns -2 otherwise: use ptivae_gn_stats)
 *   small_cout: x NHWC (storage in_fmt), optional fused GroupNorm affine scale_shift [N][Cin][2] (NO activation:
 *               encoder.blocks.15/decoder.blocks.15 are bare GroupNorms), zero padding applied after
 *               the norm -> out fp32 NCHW [N][Cout<=16][H][W]  (encoder.blocks.16, decoder.blocks.16) */
int ptivae_conv3x3_small_cin(const float* x, const float* w, const float* bias, void* out, float* gn_part,
                             int gn_groups, int N, int H, int W, int Cin, int Cout, int out_fmt, void* stream);
int ptivae_conv3x3_small_cin_parts(int H, int W, int Cout);
int ptivae_conv3x3_small_cout(const void* x, const float* w, const float* bias, const float* scale_shift, float* out,
                              int N, int H, int W, int Cin, int Cout, int in_fmt, void* stream);
/* 1x1 conv, fp32 NCHW in/out, Cin,Cout <= 16 (quant_conv_mu, quant_conv_log_sigma, post_quant_conv).
 *   act 0: none;  act 1: exp(clamp(v,-30,20)/2)  == AutoencoderKL.encode's z_sigma. */
int ptivae_conv1x1_small(const float* x, const float* w, const float* bias, float* out, int N, int HW, int Cin,
                         int Cout, int act, void* stream);

/* Single-head self-attention core: out = softmax(Q K^T * D^-0.5) V; q,k,v h16 [B][L][ld] views (row stride ld >= D
 * elements, ld % 8 == 0: ld = 3*D for a fused q|k|v projection, ld = D for separate tensors), out h16 [B][L][D],
 * D in {64,128,256} (monai SABlock with num_heads = 1, use_flash_attention=False semantics).
 * lse (NULL ok): fp32 [B][L], log2-domain log-sum-exp of the scaled score rows, saved for the backward pass. */
int ptivae_attention_fwd(const void* q, const void* k, const void* v, void* out, float* lse, int B, int L, int D, int ld,
                         int f16, void* stream);

/* AutoencoderKL.sampling: z = mu + sigma*eps.  eps_in != NULL: use the injected noise; else draw
 * eps from Philox4x32-10 (key = seed, counter = (element/4, offset)) + Box-Muller.  eps_out (may be
 * NULL) receives the noise that was used.  rng_dev (may be NULL): device array {seed, offset} that
 * overrides the by-value pair, so a captured CUDA graph draws fresh noise on every replay
 * (ptivae_rng_advance bumps the offset).  n = element count. */
int ptivae_latent_sample(const float* mu, const float* sigma, const float* eps_in, float* z, float* eps_out,
                         const unsigned long long* rng_dev, long long n, unsigned long long seed,
                         unsigned long long offset, void* stream);
int ptivae_rng_advance(unsigned long long* rng_dev, void* stream);

/* compute_kl_loss (/root/reference/src/pti_ldm_vae/models/losses.py:4-30):
 *   out[0] = mean_b( -0.5 * sum_{chw}(1 + t - mu^2 - exp(t)) ), t = `t` if input_is_logvar else log(t^2+1e-8).
 *   workspace: N floats. */
int ptivae_kl_loss(const float* mu, const float* t, float* workspace, float* out, int N, int per_img,
                   int input_is_logvar, void* stream);
/* nn.L1Loss()/nn.MSELoss() (train_vae.py:289-296): out[0] = mean|a-b|, out[1] = mean (a-b)^2.
 *   n % 4 == 0, 16-byte aligned inputs; workspace: 2*1184 floats. */
int ptivae_l1l2(const float* a, const float* b, float* workspace, float* out, long long n, void* stream);

/* latent_vectors = z_mu.mean((2,3)) (vae_scripts/train_vae.py:387-389): x fp32 [BC][HW] -> out [BC]. */
int ptivae_spatial_mean(const float* x, float* out, int BC, int HW, void* stream);

/* compute_ar_vae_loss (/root/reference/src/pti_ldm_vae/models/losses.py:69-166) on the device.
 *   zbar [B][C] fp32 latent vectors; attrs [L][B] attribute values; channel [L] latent channel per attribute;
 *   delta [L]; pairs: NULL = all ordered pairs i != j ("all" mode), else int32 [P][2] explicit (i, j) pairs
 *   ("subset" mode: the host samples them exactly as the reference does).
 *   loss_per_attr [L] = mean over pairs with a_j != a_i of (tanh(delta*(z_j-z_i)) - sign(a_j-a_i))^2 (0 if none),
 *   pair_count [L] = number of such pairs, total[0] = sum of loss_per_attr. */
int ptivae_ar_vae_loss(const float* zbar, const float* attrs, const int* channel, const float* delta, const int* pairs,
                       int P, int B, int C, int L, float* loss_per_attr, int* pair_count, float* total, void* stream);

/* Gradient of compute_ar_vae_loss w.r.t. the latent vectors (the loss is a training regulariser: train_vae.py:407-415 adds
 * ar_gamma * total to loss_g before loss_g.backward()).  pair_count: as written by ptivae_ar_vae_loss for the SAME inputs;
 * g_total: device scalar dL/d(total) (NULL = 0);  g_attr: device [L] dL/d(loss_per_attr) (NULL = 0);
 * dzbar: fp32 [B][C], overwritten (channels no attribute maps to get 0).  One thread per sample, fixed order. */
int ptivae_ar_vae_loss_bwd(const float* zbar, const float* attrs, const int* channel, const float* delta, const int* pairs,
                           int P, int B, int C, int L, const int* pair_count, const float* g_total, const float* g_attr,
                           float* dzbar, void* stream);
/* backward of ptivae_spatial_mean: dx [BC][HW] = dmean[BC] / HW */
int ptivae_spatial_mean_bwd(const float* dmean, float* dx, int BC, int HW, void* stream);

/* y [B][O] = act(x [B][I] * W[O][I]^T + b): nn.Linear (+activation) of LatentRegressor
 * (/root/reference/src/pti_ldm_vae/models/regression_head.py:30-78).  act: 0 none, 1 relu, 2 gelu, 3 leaky_relu, 4 elu. */
int ptivae_linear_act(const float* x, const float* w, const float* bias, float* y, int B, int I, int O, int act,
                      void* stream);

/* ---- callers either side of the hot path --------------------------------------------------
 * Per-sample evaluation metrics of evaluate_vae.py in one pass:
 *   replaces: compute_psnr / compute_ssim (src/pti_ldm_vae/utils/eval_metrics.py:6-63) and the per-sample MSE/MAE
 *             of vae_scripts/evaluate_vae.py:87-98 (torch.clamp(.,0,1) of both images first when do_clamp != 0)
 *   pred, target fp32 NCHW [B][C][H][W];  window: `win` (odd, <= 15) fp32 taps of the separable SSIM window
 *   (the reference's 11-tap Gaussian, sigma 1.5, normalised); zero padding as conv2d(padding = win/2)
 *   out fp32 [B][4] = (mse, mae, psnr, ssim);  workspace: ptivae_eval_metrics_workspace(B, C, H, W) bytes */
int ptivae_eval_metrics(const float* pred, const float* target, const float* window, int win, float* out,
                        void* workspace, int B, int C, int H, int W, int do_clamp, float lo, float hi,
                        float data_range, float k1, float k2, void* stream);
int ptivae_eval_metrics_workspace(int B, int C, int H, int W);
/* LocalNormalizeByMask (src/pti_ldm_vae/data/transforms.py:8-32) for a batch on the device: per image, z-score
 * with the mean / population std of the NON-ZERO pixels (std <= 1e-5 -> 1), zero pixels stay exactly 0.
 *   x, out fp32 [B][per_img];  stats: optional fp32 [B][2] (mean, std used);  workspace:
 *   ptivae_local_normalize_workspace(B) bytes (fp64 partial sums, fixed order) */
int ptivae_local_normalize(const float* x, float* out, float* stats, void* workspace, int B, int per_img, void* stream);
int ptivae_local_normalize_workspace(int B);
/* Resize(patch_size) of the reference's preprocessing (data/dataloaders.py:263-272, :323-327: MONAI Resize, default mode
 * "area" = torch.nn.functional.interpolate(mode="area") = adaptive average pooling) for a batch of equally sized raw images:
 *   in  [B][H][W], in_fmt 0 = uint8, 1 = uint16, 2 = float32 (as decoded from the TIFF);  out fp32 [B][Ho][Wo]
 *   out[oy][ox] = mean of in[floor(oy*H/Ho) : ceil((oy+1)*H/Ho)][floor(ox*W/Wo) : ceil((ox+1)*W/Wo)] */
int ptivae_resize_area(const void* in, int in_fmt, float* out, int B, int H, int W, int Ho, int Wo, void* stream);

/* ---- backward pass (`loss_g.backward()`, vae_scripts/train_vae.py:444) -----------------
 * Gradient tensors are NHWC; every 16-bit operand of a backward GEMM is bf16 (the gradients' range; both operands of
 * one MMA must share a format).
 *
 * Weight gradient of a convolution on tcgen05 (replaces cuDNN wgrad): dw[co][ci][ky][kx] = sum dY[p][co]*X[p+tap][ci].
 *   dy   h16 NHWC gradient of the conv output, Ca channels;  x  h16 NHWC conv input (as the forward GEMM read it), Cb channels
 *   mode 0: 3x3 s1 p1, H,W = extent of x (= of dy)     1: F.pad(0,1,0,1)+3x3 s2, H,W = extent of x (dy is H/2 x W/2)
 *        2: nearest x2 upsample + 3x3, H,W = extent of x (dy is 2H x 2W)     3: 1x1
 *        4: same result as mode 0 through the generic one-box-per-tap kernel (mode 0 shares one activation box
 *           between the three kx taps of a kernel row; the tests compare the two)
 *   dw   fp32 [Ca][Cb][3][3] (modes 0-2) or [Ca][Cb] (mode 3): the master-weight layout, overwritten
 *   workspace: ptivae_wgrad_workspace(...) bytes of split-K partial sums (plain stores, summed in index order)
 *   f16: 16-bit format of BOTH operands (one tcgen05 MMA cannot mix fp16 and bf16) */
int ptivae_wgrad(const void* dy, const void* x, float* workspace, float* dw, int N, int H, int W, int Ca, int Cb, int mode,
                 int f16, void* stream);
long long ptivae_wgrad_workspace(int N, int H, int W, int Ca, int Cb, int mode);
/* debug only: 6 ints of host-mapped memory that receive (role, tile, stage, block x, y, z) of the first barrier wait that
 * timed out in subsequent ptivae_wgrad launches (the kernel then exits instead of trapping); NULL switches it off */
int ptivae_debug_set_wgrad_status(void* status);

/* Batched GEMM on tcgen05 with operands consumed as they lie in memory (attention backward):
 *   out[b][m][n] = epi( sum_k A[b](m,k) * B[b](n,k) ),  16-bit operands of ONE format, fp32 accumulate, 16-bit out.
 *   a_mn == 0: A(m,k) at a[b*bs_a + m*lda + k] (K-major);  a_mn != 0: at a[b*bs_a + k*lda + m] (MN-major); same for b.
 *   epi 0: acc*alpha;  1: exp2(acc*alpha - rowv[b][m]);  2: aux[b][m][n]*(acc - rowv[b][m])*alpha
 *   leading dimensions / batch strides in elements (multiples of 8); stores are 8 columns wide, so with N % 8 != 0 the
 *   last group spills into the row padding of out (ldo >= round_up(N, 8)). */
int ptivae_bgemm(const void* a, const void* b, void* out, int B, int M, int N, int K, long long lda, long long bs_a, int a_mn,
                 int a_f16, long long ldb, long long bs_b, int b_mn, int b_f16, long long ldo, long long bs_out, int out_f16,
                 int epi, float alpha, const float* rowv, const void* aux, long long ld_aux, long long bs_aux, int aux_f16,
                 void* stream);
/* out[row] = sum_d a[row][d]*b[row][d] (16-bit rows with strides lda/ldb, fp32 out): rowsum(dO o O) of the softmax backward */
int ptivae_rowdot(const void* a, const void* b, float* out, long long rows, int D, long long lda, long long ldb, int a_f16,
                  int b_f16, void* stream);

/* Backward of y = act(GroupNorm(x)) (act = SiLU when silu != 0), given da = dL/dy:
 *   dx = scale_c*du - e_g - x*f_g (+ residual), du = da*act'(x*scale+shift);  dgamma/dbeta fp32 [C] (overwritten)
 *   x (x_fmt), da (da_fmt), residual (res_fmt, NULL ok): NHWC [N][HW][C];  dx32 (fp32) and/or dx16 (bf16): at least one
 *   scale_shift [N][C][2], mean_rstd [N][G][2] from ptivae_gn_finalize;  coef: fp32 [N][C][2] scratch
 *   act_out (NULL ok): bf16 [N][HW][C] = act(x*scale+shift), the re-materialised forward operand (what the weight
 *             gradient of the conv that consumed y reads), written by the reduction pass that has x in registers anyway
 *   colsum_out (NULL ok): fp32 [C] = per-channel sum of dx = the bias gradient of the conv that produced x
 *   workspace: ptivae_gn_bwd_workspace(N, HW, C) floats.  Deterministic, batch-invariant chunking. */
int ptivae_gn_bwd(const void* x, int x_fmt, const void* da, int da_fmt, const float* scale_shift, const float* mean_rstd,
                  const float* gamma, const void* residual, int res_fmt, float* dx32, void* dx16, float* dgamma,
                  float* dbeta, void* act_out, float* colsum_out, float* coef, float* workspace, int N, int HW, int C, int G,
                  int silu, void* stream);
long long ptivae_gn_bwd_workspace(int N, int HW, int C);
/* bias gradient: out[c] = sum over rows of x[row][c] (x NHWC storage fmt); workspace ptivae_colsum_blocks(rows)*C floats */
int ptivae_colsum(const void* x, float* out, float* workspace, long long rows, int C, int fmt, void* stream);
int ptivae_colsum_blocks(long long rows);

/* Weight (and thin-side bias) gradient of the thin-end 3x3 convs (ptivae_conv3x3_small_cin / _small_cout):
 *   wide_is_input != 0: thin = dOut fp32 NCHW [N][Ct][H][W], wide = the conv's NHWC input (optional fused GroupNorm
 *       affine scale_shift, as the forward applied it);  dw [Ct][C][3][3], db [Ct] (NULL ok)
 *   wide_is_input == 0: thin = the conv's fp32 NCHW input, wide = dOut NHWC;  dw [C][Ct][3][3], db must be NULL */
int ptivae_thin_wgrad(const float* thin, const void* wide, const float* scale_shift, float* dw, float* db, float* workspace,
                      int N, int H, int W, int C, int Ct, int wide_fmt, int wide_is_input, void* stream);
long long ptivae_thin_wgrad_workspace(int N, int H, int W, int C, int Ct);
/* Gradient through the latent head (AutoencoderKL.encode tail + sampling + post_quant_conv), fp32 NCHW [N][L][HW]:
 *   in: dzq = dL/d(post_quant_conv out), dmu_ext / dsig_ext = gradients arriving at the returned z_mu / z_sigma (NULL ok),
 *       eps, h (encoder stack output), mu, sigma, the three [L][L] weights and quant_conv_log_sigma's bias
 *   out: dh = dL/dh, dmu, dlv (gradients at the two quant conv outputs), z = mu + sigma*eps (for dW of post_quant) */
int ptivae_latent_bwd(const float* dzq, const float* dmu_ext, const float* dsig_ext, const float* eps, const float* h,
                      const float* mu, const float* sigma, const float* wp, const float* wm, const float* ws, const float* bs,
                      float* dh, float* dmu, float* dlv, float* z, int N, int HW, int L, void* stream);
/* dw[i][j] = sum_{n,p} a[n][i][p]*b[n][j][p], db[i] = sum_{n,p} a[n][i][p] (NULL ok): gradients of a 1x1 latent conv */
int ptivae_outer_reduce(const float* a, const float* b, float* dw, float* db, int N, int I, int J, int HW, void* stream);

/* Gradients of the loss terms; gout points to the upstream gradient scalar(s) in DEVICE memory.
 *   l1l2_bwd: d = gout[0]*sign(a-b)/n + gout[1]*2(a-b)/n   (outputs of ptivae_l1l2);  kl_bwd: see ptivae_kl_loss */
int ptivae_l1l2_bwd(const float* a, const float* b, const float* gout, float* d, long long n, void* stream);
int ptivae_kl_bwd(const float* mu, const float* t, const float* gout, float* dmu, float* dt, int N, int per_img,
                  int input_is_logvar, void* stream);
/* dst (h16: fp16 when dst_f16 != 0, else bf16) = src (storage src_fmt); n % 8 == 0.  One tcgen05 MMA needs both
 * operands in the same 16-bit format (measured: mixing raises an illegal-instruction fault), the backward GEMMs run in
 * bf16, so the saved fp16 forward operands they read are converted once. */
int ptivae_cast16(const void* src, void* dst, long long n, int src_fmt, int dst_f16, void* stream);
/* torch.optim.Adam (defaults) over one flat fp32 buffer; step_dev: device float = 1-based step of this update,
 * incremented afterwards when advance != 0; grad_scale multiplies g (1/world_size after a sum all-reduce). */
int ptivae_adam(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                float grad_scale, float* step_dev, int advance, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PTIVAE_H_ */
