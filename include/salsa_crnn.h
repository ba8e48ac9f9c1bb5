/*
 * salsa_crnn -- C ABI of the SELD CRNN forward operators in libsalsa_b200.so.
 *
 * The reference has no FFI layer; its CRNN is torch.nn modules (models/encoders.py,
 * models/decoders.py, models/model_utils.py) executed by cuDNN / cuBLAS.  Each entry point cites the
 * module call it replaces.  Conventions are those of salsa_b200.h: device pointers owned by the
 * caller, asynchronous on `stream`, 0 / negative SALSA_E* return, message in salsa_last_error().
 *
 * Layouts
 *   activations  bf16 NHWC  [B][H][W][C], C a multiple of 64 (H = time frames, W = frequency)
 *   conv weights bf16       [kh*kw][Cout][Cin]  (BatchNorm scale already folded in)
 *   GEMM weights bf16       [N][K]              (nn.Linear.weight layout)
 *   bias         fp32       [Cout] / [N]        (folded BatchNorm shift or nn.Linear.bias)
 *
 * Precision: `planes` = 1 is plain bf16 operands with fp32 accumulation.  `planes` = 2 / 3 ("bf16x2" / "bf16x3") store every
 * activation and weight as the sum of two / three bf16 planes laid side by side along the channel axis ([..][planes][C],
 * hi | mid | lo) and accumulate the plane products that matter (3 for two planes: hi hi, hi lo, lo hi; 6 for three) on the
 * same tensor-core path: float32-grade results (logits 6-8e-5 / 3-4e-5 of the output scale from the float32 reference) at
 * three / six times the MMA work, used for parity with the reference.
 */
#ifndef SALSA_CRNN_H
#define SALSA_CRNN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- model level: one call per batch ------------------------------------------------------------------------------
 * What a non-Python host needs from models.seld_models.SeldModel (models/seld_models.py:39-49) and the checkpoint
 * loading of experiments/inference.py:102-116. */

/* One tensor of the reference's state dict (Lightning checkpoint ['state_dict'], keys `encoder.*` / `decoder.*`):
 * HOST pointer to contiguous float32 values in the reference's own layout (nn.Conv2d weight (Cout, Cin, k, k),
 * BatchNorm weight / bias / running_mean / running_var, nn.GRU weight_ih_l* / weight_hh_l* / bias_*, nn.Linear weight / bias). */
typedef struct crnn_tensor {
    const char *name;
    const float *data;
    int64_t numel;
} crnn_tensor_t;

/* Builds the inference model on the current device: folds every eval-mode BatchNorm into its convolution, packs the
 * weights for the tensor-core kernels (planes = 1: bf16; 2 / 3: two / three bf16 planes per value, float32-grade), stacks
 * the two GRU directions and fuses the four heads.  Entries the model does not use (num_batches_tracked, ...) are
 * ignored; a missing or mis-sized entry is SALSA_EINVAL.  PannResNet22(n_input_channels=7) + SeldDecoder(512, n_classes,
 * 'reg_xyz', 'bigru', 'avg', 256).  *model_out is released with crnn_free_model(). */
int crnn_load_weights(const crnn_tensor_t *tensors, int32_t n_tensors, int32_t planes, int32_t n_classes, void **model_out);
int crnn_free_model(void *model);

/* Bytes of device scratch crnn_forward needs for a (B, 7, T, F) batch (T = the frames actually used). */
size_t crnn_workspace_bytes(const void *model, int32_t B, int32_t T, int32_t F);

/* SeldModel.forward (models/seld_models.py:39-49): feat float32 [B][7][T_in][F] (the h5 'feature' layout; the first
 * T_use <= T_in frames of every clip are used: dataset/database.py:205-207 trims 4801 -> 4800) ->
 * logits float32 [B][T'][n_classes] ('event_frame_logit'), doa float32 [B][T'][3 n_classes] ('doa_frame_output', x | y | z),
 * T' = T_use / 16 (four 2x2 poolings, floor).  mean / std float32 [n_scaled][F] (or NULL with n_scaled = 0): the data
 * layer's (x - mean) / std on the first n_scaled channels (dataset/database.py:196-202), fused into the first kernel.
 * All pointers are device pointers owned by the caller; nothing is allocated.  The ~30 kernels of the forward are captured
 * into a CUDA graph the second time the same set of buffers is seen and replayed from then on (asynchronous on `stream`). */
int crnn_forward(void *model, const float *feat, int32_t B, int32_t T_in, int32_t T_use, int32_t F, const float *mean,
                 const float *std, int32_t n_scaled, float *logits, float *doa, void *workspace, size_t workspace_bytes,
                 void *stream);

/* "graph" (0 / 1): replay a captured CUDA graph (default) or launch the kernels one by one. */
int crnn_model_option(const char *name, int32_t value);

/* ---- operator level ------------------------------------------------------------------------------------------------ */

/* nn.Conv2d(k=3, pad=1 | k=1, stride 1, bias=False) + eval-mode BatchNorm2d (+ residual add) (+ ReLU):
 * ConvBlock.forward (models/model_utils.py:213-216), _ResnetBasicBlock.forward (:352-365), the
 * downsample branch (:474-481).  tcgen05 implicit GEMM, fp32 accumulation.
 * out (bf16) and/or out_f32 receive relu?(conv(x, w) + bias + residual).  pool = 1 (planes = 1, bf16 output only)
 * applies the following F.avg_pool2d(2x2) (:220, :349) in the epilogue: out is then [B][H/2][W/2][Cout]. */
int crnn_conv2d(const void *x, const void *w, const float *bias, const void *residual, void *out, float *out_f32,
                int32_t B, int32_t H, int32_t W, int32_t Cin, int32_t Cout, int32_t ksize, int32_t relu, int32_t planes,
                int32_t pool, void *stream);

/* Weight gradient of nn.Conv2d(k=3, pad=1, stride 1, bias=False) (the backward pass of models/model_utils.py:192-200,
 * :301-304 in the training step, models/seld_models.py:68-76) or, with ksize = 1, of the 1x1 downsample convolutions
 * (:307-309): x bf16 NHWC [B][H][W][Cin], gy = dLoss/dOutput bf16 NHWC [B][H][W][Cout] -> dw fp32 [ksize^2][Cout][Cin]
 * (tap-major like the packed weights; overwritten).  tcgen05 GEMMs over the pixel axis on MN-major operands, two taps
 * stacked per M = 128 MMA, split over the grid, accumulated with fp32 vector atomics.  (The input gradient is crnn_conv2d
 * itself on gy with the taps flipped and Cin / Cout exchanged.) */
int crnn_conv_wgrad(const void *x, const void *gy, float *dw, int32_t B, int32_t H, int32_t W, int32_t Cin, int32_t Cout,
                    int32_t ksize, void *stream);

/* nn.BatchNorm2d in TRAIN mode (batch statistics; models/model_utils.py:202-203, :216, :356 in the training step), fused with
 * the residual add, the ReLU and the element-wise dropout around it, on NHWC bf16 activations viewed as [n_pix][C]
 * (C = 64, 128, 256 or 512):
 *   z = dropout?(relu?(gamma (y - mean) / sqrt(var + eps) + beta (+ residual))),  mean / var = biased statistics of y
 *   stat fp32 [C][2] receives (mean, 1 / std) for the backward pass; sums float64 [C][2] is scratch; running_mean /
 *   running_var fp32 [C] (both or neither) are updated with `momentum` (unbiased variance), like nn.BatchNorm2d.
 *   Dropout (nn.Dropout(p) behind relu(bn1(.)), :356): drop_seed is a DEVICE scalar (NULL or drop_p = 0: none), drop_salt
 *   tells the layers sharing it apart; kept elements are scaled by 1 / (1 - drop_p).  The keep decisions are a function of
 *   (seed, salt, element index), so no mask is stored: the backward call recomputes them from the same three values, and a
 *   CUDA graph of the step sees the seed the device scalar holds at replay time. */
int crnn_bn_train_forward(const void *y, const float *gamma, const float *beta, const void *residual, void *z, float *stat,
                          double *sums, float *running_mean, float *running_var, int64_t n_pix, int32_t C, float eps,
                          float momentum, int32_t relu, const uint64_t *drop_seed, uint32_t drop_salt, float drop_p,
                          void *stream);

/* Its backward pass: dz = dLoss/dz (bf16), z / y / stat from the forward ->
 *   dy bf16 = gamma / std (g - mean(g) - xhat mean(g xhat)),  g = dz * mask (* keep / (1 - drop_p)),  xhat = (y - mean) / std
 *   d_residual bf16 (optional) = g;  dgamma fp32 [C] = sum g xhat;  dbeta fp32 [C] = sum g;  sums float64 [C][2] scratch.
 *   relu = 0: no ReLU in the forward (mask = 1);  1: mask = z > 0, read from z;  2: the forward had NO residual, so the mask
 *   is recomputed from y, gamma, beta and stat with the forward's own expression (bit-identical to z > 0; z may be NULL and
 *   one tensor less is read).  beta is only used by relu = 2.  drop_*: the forward call's values. */
int crnn_bn_train_backward(const void *dz, const void *z, const void *y, const float *stat, const float *gamma,
                           const float *beta, void *dy, void *d_residual, double *sums, float *dgamma, float *dbeta,
                           int64_t n_pix, int32_t C, int32_t relu, const uint64_t *drop_seed, uint32_t drop_salt,
                           float drop_p, void *stream);

/* BatchNorm (+ residual) + ReLU + the F.avg_pool2d(2) behind it as one pass each way, for the four places where the network
 * pools (models/model_utils.py:220, :349, :476) and the BatchNorm output has no other consumer: y bf16 NHWC [B][H][W][C] ->
 * pooled bf16 [B][H/2][W/2][C]; the full-resolution activation is never written.  Results are bit-identical to
 * crnn_bn_train_forward(relu = 1) followed by crnn_avgpool2.  stat / sums / running_*: as in crnn_bn_train_forward. */
int crnn_bn_train_forward_pool(const void *y, const float *gamma, const float *beta, const void *residual, void *pooled,
                               float *stat, double *sums, float *running_mean, float *running_var, int32_t B, int32_t H,
                               int32_t W, int32_t C, float eps, float momentum, void *stream);

/* Its backward pass: dpool bf16 [B][H/2][W/2][C] = dLoss/dpooled -> dy bf16 [B][H][W][C], d_residual (optional; needs
 * `residual`, the forward's residual input, from which together with y the ReLU mask is recomputed), dgamma, dbeta.
 * Bit-identical to crnn_avgpool2_backward followed by crnn_bn_train_backward. */
int crnn_bn_train_backward_pool(const void *dpool, const void *y, const void *residual, const float *stat,
                                const float *gamma, const float *beta, void *dy, void *d_residual, double *sums,
                                float *dgamma, float *dbeta, int32_t B, int32_t H, int32_t W, int32_t C, void *stream);

/* The first convolution of the encoder (conv_block1.conv1 + bn1 + ReLU, models/model_utils.py:213-215) on an input
 * padded to 16 channels: x bf16 NHWC [B][H][W][planes*16], w bf16 [9][64][planes*16] -> out bf16 [B][H][W][planes*64]. */
int crnn_conv_first(const void *x, const void *w, const float *bias, void *out, int32_t B, int32_t H, int32_t W,
                    int32_t relu, int32_t planes, void *stream);

/* Tuning knobs of the convolution kernels (not part of the reference surface): "resident_b" (0 / 1: weights resident
 * in shared memory for 64 -> 64 convolutions), "tma_store" (0 / 1: bf16 outputs leave through a swizzled staging
 * tile and TMA stores instead of per-thread stores), "gru_mma" (0 / 1: tensor-core recurrent step for planes = 1);
 * value -1 restores the built-in choice. */
int crnn_set_option(const char *name, int32_t value);

/* nn.Linear / GRU input projection (models/decoders.py:44-46, :75-92): out[M][N] = relu?(a[M][K] w[N][K]^T + bias).
 * `a` must be allocated with its row count rounded up to a multiple of 8. */
int crnn_gemm(const void *a, const void *w, const float *bias, void *out, float *out_f32, int32_t M, int32_t N,
              int32_t K, int32_t relu, int32_t planes, void *stream);

/* (B, C, T, F) fp32 NCHW feature batch (the tensor SeldModel.forward receives, models/seld_models.py:39-43)
 * -> bf16 NHWC [B][T_use][F][Cpad], channels C..Cpad-1 zero; frames >= T_use are dropped
 * (database.py:205-207 trims 4801 -> 4800).  With n_scaled > 0 the first n_scaled channels are normalised on the
 * way, (x - mean[c][f]) / std[c][f] with mean / std float32 [n_scaled][F] (the scaler h5, database.py:196-202). */
int crnn_pack_input(const float *x, void *y, int32_t B, int32_t C, int32_t T, int32_t F, int32_t T_use, int32_t Cpad,
                    int32_t planes, const float *mean, const float *std, int32_t n_scaled, void *stream);

/* F.avg_pool2d(x, kernel_size=(2, 2)) (models/model_utils.py:220, :349; nn.AvgPool2d :476), floor mode. */
int crnn_avgpool2(const void *x, void *y, int32_t B, int32_t H, int32_t W, int32_t C, int32_t planes, void *stream);

/* Backward of that pooling (training step): dy bf16 [B][H/2][W/2][C] -> dx bf16 [B][H][W][C] = dy / 4 under every 2x2 window,
 * 0 in an odd last row / column. */
int crnn_avgpool2_backward(const void *dy, void *dx, int32_t B, int32_t H, int32_t W, int32_t C, void *stream);

/* torch.mean(x, dim=3) + transpose (models/decoders.py:111, :123): [B*H][W][C] -> [B*H][C]. */
int crnn_freq_mean(const void *x, void *y, int32_t BH, int32_t W, int32_t C, int32_t planes, void *stream);

/* Recurrent part of one bidirectional nn.GRU layer, hidden size 256 (models/decoders.py:44-46, :126).
 *   xproj fp32 [B*T][1536] = x W_ih^T + b_ih, forward direction in columns 0..767 (r|z|n), backward in 768..1535
 *   w_hh  fp32 [2][768][256], b_hh fp32 [2][768]   (weight_hh_l*, weight_hh_l*_reverse)
 *   y     bf16 [B*T][512]     forward | backward hidden states */
int crnn_gru_layer(const float *xproj, const float *w_hh, const float *b_hh, void *y, int32_t B, int32_t T, int32_t planes,
                   void *stream);

/* The same recurrence for the TRAINING step (float32 throughout), keeping what back-propagation through time needs:
 *   y    fp32 [B*T][512]       forward | backward hidden states
 *   save fp32 [B*T][2][4][256] per direction r, z, n and hn = W_hn h + b_hn of every step */
int crnn_gru_layer_train(const float *xproj, const float *w_hh, const float *b_hh, float *y, float *save, int32_t B,
                         int32_t T, void *stream);

/* Back-propagation through time of that layer (the backward pass of nn.GRU in models/seld_models.py:68-76): dy fp32
 * [B*T][512] = dLoss/dy, y / save from crnn_gru_layer_train ->
 *   dgi fp32 [B*T][1536] = dLoss/d(x W_ih^T + b_ih)  and  dgh fp32 [B*T][1536] = dLoss/d(h_prev W_hh^T + b_hh),
 * same column layout as xproj.  The parameter and input gradients are GEMMs / column sums over them, left to the caller:
 * dW_ih = dgi^T x, db_ih = sum dgi, dx = dgi W_ih, dW_hh = dgh^T h_prev, db_hh = sum dgh. */
int crnn_gru_layer_backward(const float *dy, const float *y, const float *save, const float *w_hh, float *dgi, float *dgh,
                            int32_t B, int32_t T, void *stream);

/* Output stage of SeldDecoder.forward (models/decoders.py:137-147): z fp32 [rows][64] holds the fused
 * event|x|y|z second-layer outputs in columns 0..4*n_classes-1; logits = z[:, :n], doa = tanh(z[:, n:4n]). */
int crnn_head_finish(const float *z, float *logits, float *doa, int32_t rows, int32_t n_classes, void *stream);

/* Output decoding of BaseModel.write_classwise_output_to_file (models/interfaces.py:224-246), reg_xyz format, on
 * label-rate outputs: logits fp32 [rows][n_classes], doa fp32 [rows][3*n_classes] (x | y | z) ->
 * active[rows][n_classes] = sigmoid(logit) >= threshold, azi / ele int16 whole degrees (np.around, 180 -> -180). */
int crnn_decode_events(const float *logits, const float *doa, int32_t rows, int32_t n_classes, float threshold,
                       uint8_t *active, int16_t *azi, int16_t *ele, void *stream);

/* interpolate_tensor (models/model_utils.py:57-75): out[b][i][:] = in[b][idx[i]][:], idx computed by the
 * host exactly as the reference does (floor(arange(n_out) / ratio) in float32). */
int crnn_gather_time(const float *in, const int32_t *idx, float *out, int32_t B, int32_t n_in, int32_t n_out,
                     int32_t width, void *stream);

/* Training-time augmentations that commute with the SALSA feature layout, on a device batch:
 * TfmapRandomSwapChannelFoa.apply (utilities/transforms.py:394-437), TfmapRandomSwapChannelMic.apply (:470-523) and
 * RandomShiftUpDownNp.apply (:298-320, mode='reflect', all channels), composed as dataset/datamodule.py:45-83 composes
 * them (joint swap, then the shift).  The random draws stay with the host (the reference's NumPy draws):
 *   ops int32 [B][4] = {format (0 foa / 1 mic), swap flags (bit i = m[i], 0 = transform skipped),
 *                       shift_len (0 = skipped), direction (0 up / 1 down)}       -- a DEVICE array
 *   x, out    fp32 [B][7][T][F]  (out != x)
 *   y_doa, y_out fp32 [B][Ty][3*n_classes] (x | y | z per class), both NULL to leave the labels alone.
 * Permutations, sign flips and single float32 subtractions in the reference's order: bit-identical results. */
int crnn_augment(const float *x, float *out, const float *y_doa, float *y_out, const int32_t *ops, int32_t B,
                 int32_t T, int32_t F, int32_t Ty, int32_t n_classes, void *stream);

/* CompositeCutout.apply (utilities/transforms.py:257-283; composed behind the frequency shift for MIC SALSA features,
 * dataset/datamodule.py:76-82) on a device batch, IN PLACE: whichever of RandomCutoutNp (:58-125), SpecAugmentNp (:128-196)
 * or RandomCutoutHoleNp (:199-254) the host drew is a list of at most 8 rectangles per sample:
 *   rects   int32 [B][8][4] = {first frame, end frame, first frequency, end frequency} (ends exclusive), applied in order
 *   n_rects int32 [B]       (0 = transform skipped for this sample)
 *   u       float64 [B][8]  the uniform [0, 1) draw behind each rectangle's fill value
 *   minmax  fp32 [B][2]     scratch: np.min / np.max of each sample before any cut, computed here
 * A rectangle sets channels 0 .. C - n_zero_channels - 1 to float32(min + (max - min) * u) and the last n_zero_channels
 * channels to 0 (is_filled_last_channels = True): the reference's arithmetic, bit for bit. */
int crnn_cutout(float *x, const int32_t *rects, const int32_t *n_rects, const double *u, float *minmax, int32_t B, int32_t C,
                int32_t T, int32_t F, int32_t n_zero_channels, void *stream);

/* BaseModel.compute_loss for output_format='reg_xyz' (models/interfaces.py:273-355; SURVEY.md 8 f1, loss row):
 * sed_loss = mean BCE-with-logits over [rows][n_classes]; doa_loss = sum over x, y, z of sum(|pred - gt| * event_gt) /
 * sum(event_gt); loss = w_sed * sed_loss + w_doa * doa_loss (seld.yml:53-55: 0.3 / 0.7).
 *   logit, event_gt fp32 [rows][n_classes]; doa, doa_gt fp32 [rows][3*n_classes] (x | y | z), rows = batch * label frames,
 *   time axes already aligned by the caller (compute_masked_reg_loss :337-341)
 *   sums   float64 [5] device scratch (zeroed here); loss fp32 [3] device = {loss, sed_loss, doa_loss}
 *   g_logit / g_doa: optional gradients of `loss` with respect to logit / doa (same shapes), NULL to skip. */
int crnn_seld_loss(const float *logit, const float *doa, const float *event_gt, const float *doa_gt, int64_t rows,
                   int32_t n_classes, float w_sed, float w_doa, double *sums, float *loss, float *g_logit,
                   float *g_doa, void *stream);

/* One torch.optim.Adam step (models/interfaces.py:85-95: no weight decay, no amsgrad) on a flat fp32 buffer of n values,
 * with the lr / beta1 of this batch as LearningRateScheduler sets them (utilities/learning_utils.py:39-52); `step`
 * counts from 1; the hyper-parameters are doubles like torch's Python scalars (1 - beta is formed before rounding).  param, exp_avg and exp_avg_sq are updated in place (optimiser row). */
int crnn_adam_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, double lr, double beta1,
                   double beta2, double eps, int32_t step, void *stream);

/* The same step with its per-batch scalars in DEVICE memory, for a training step replayed as a CUDA graph (the reference
 * changes lr / beta1 every batch, utilities/learning_utils.py:44-52): crnn_adam_hyper fills hyper_host[6] = {1 - beta1, beta2,
 * 1 - beta2, eps, lr / (1 - beta1^step), sqrt(1 - beta2^step)} exactly as crnn_adam_step forms them; the caller copies the 24
 * bytes to the device (stream-ordered) before crnn_adam_step_hyper, which reads them from hyper_dev. */
int crnn_adam_hyper(double lr, double beta1, double beta2, double eps, int32_t step, float *hyper_host);
int crnn_adam_step_hyper(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n,
                         const float *hyper_dev, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SALSA_CRNN_H */
