/* xhved.h -- C ABI of the B200-native ViL-mLSTM + S-MVAE hot path of XLSTM-HVED.
 *
 * The reference (Quanato607/XLSTM-HVED) is pure Python/PyTorch and has no FFI
 * layer; this library is what sits underneath the reference's unchanged Python
 * module API (see INTEGRATION.md for the binding a maintainer adds).  Every
 * entry point takes raw DEVICE pointers, explicit sizes and a cudaStream_t
 * (passed as void*), never allocates, keeps no global state, is re-entrant per
 * stream, and returns 0 on success, a positive cudaError_t, or a negative
 * XHVED_ERR_* validation code.  Kernels are sm_100a only.
 *
 * Reference symbols each group replaces (paths relative to the reference root):
 *   xhved_poe_* / xhved_reparam_*   buildingblocks.py:846-886 (ProductOfExperts, ProductOfExperts2),
 *                                   loss.py:42-83 (duplicates), RA_HVED.py:741-753 (reparametrize, clip),
 *                                   loss.py:29-40,85-115 (KL_divergence / compute_KLD)
 *   xhved_mlstm_*                   UxLSTM/nnunetv2/nets/vision_lstm.py:48-130 (parallel_stabilized_simple)
 *   xhved_vil_pre_*                 vision_lstm.py:224-268 (LayerNorm), 415-435 (ViLLayer.forward up to q,k,v),
 *                                   178-221 (CausalConv1d), 133-175 (LinearHeadwiseExpand), 302-318 (gates)
 *   xhved_vil_post_*                vision_lstm.py:271-287 (MultiHeadLayerNorm), 437-453 (skip, z gate, proj_down,
 *                                   un-flip), vision_lstm_util.py:171-175 (residual), UxLSTMEnc_3d.py:54-63
 *
 * Tile-native layout ("tiles"): a [R rows][C cols] bf16 tile stores element (r,c) at byte
 * ((c/8)*R + r)*16 + (c%8)*2.  q/k/v/h/dh tiles have R = 128 (one chunk of tokens) and C = dhp
 * (head dim padded to a multiple of 16); tile index = (b*NH + head)*nc + chunk, nc = ceil(S/128).
 * "Padded gates" are fp32 (BH, nc*128) with i = -1e30, f = +1e30 in the padding rows.
 */
#ifndef XHVED_H_
#define XHVED_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XHVED_ERR_BAD_SHAPE (-1)
#define XHVED_ERR_UNSUPPORTED_DH (-2)
#define XHVED_ERR_BAD_ARG (-3)
#define XHVED_ERR_UNSUPPORTED_DIM (-4)

int xhved_version(void);

/* ---------------------------------------------------------------- S-MVAE product of experts (K4)
 * mu, logvar: (5, n) fp32, expert e at base + e*expert_stride (elements); expert 0 is the prior
 * (RA_HVED.py:576-580).  For every requested subset s (bit m of subset_masks[s] set = modality m
 * present; the prior is always included, buildingblocks.py:854-855):
 *   T_e = 1/(exp(logvar_e) + eps);  out_mu = sum mu_e T_e / sum T_e;  out_logvar = log(1 / sum T_e)
 * drop (optional, device, (n/per_sample, 4) uint8): per-sample missing flags of ProductOfExperts2
 * (buildingblocks.py:875-886): expert m+1 of sample b is removed where drop[b][m] != 0.
 * noise/out_z (optional, (n_subsets, n)): z = out_mu + noise * exp(0.5 out_logvar) (RA_HVED.py:741-747).
 * kld_out (optional, device float[n_subsets], must be zeroed by the caller): accumulates the sum over elements of
 *   -1 + lv_0 - lv + (exp(lv) + (mu - mu_0)^2) / (exp(lv_0) + 1e-8)   with (mu_0, lv_0) = slab 0, the prior the
 *   reference hands to KL_divergence (loss.py:95-97, 113, 29-40); the caller scales by 0.5/n.  With
 *   XHVED_POE_STANDARD_PRIOR slab 0 is taken to be (0, 0) without being read.
 * subset_masks is a HOST array of n_subsets (<= 15) entries.
 * flags: XHVED_POE_STANDARD_PRIOR = the caller guarantees that expert 0 is the standard normal the model always passes
 *   (mu = 0, logvar = 0, RA_HVED.py:576-580): its slabs are then never read (T_0 = 1/(1+eps)); mu / logvar still point
 *   at the position slab 0 would have (expert e at base + e*expert_stride). */
#define XHVED_POE_STANDARD_PRIOR 1
int xhved_poe_fwd(const float* mu, const float* logvar, int64_t n, int64_t expert_stride, const uint32_t* subset_masks,
                  int n_subsets, const uint8_t* drop, int64_t per_sample, float eps, float* out_mu, float* out_logvar,
                  const float* noise, float* out_z, float* kld_out, int flags, void* stream);

/* Backward of xhved_poe_fwd (optionally through z and the KL term): g_mu, g_logvar, g_z (each optional,
 * (n_subsets, n)); kld_scale[s] (host, optional) = dLoss/d(kld_out[s]).  Writes d_mu, d_logvar (5, n)
 * at stride expert_stride (summed over subsets; the prior slot receives its gradient too -- unless
 * XHVED_POE_STANDARD_PRIOR is set: then slab 0 is neither read nor written). */
int xhved_poe_bwd(const float* mu, const float* logvar, int64_t n, int64_t expert_stride, const uint32_t* subset_masks,
                  int n_subsets, const uint8_t* drop, int64_t per_sample, float eps, const float* g_mu, const float* g_logvar,
                  const float* noise, const float* g_z, const float* kld_scale, float* d_mu, float* d_logvar, int flags,
                  void* stream);

/* Several latent levels in ONE launch (the four latent resolutions of a volume, RA_HVED.py:567-605: the small levels are
 * latency-bound on their own).  Every field has the meaning of the xhved_poe_fwd / xhved_poe_bwd argument of the same name;
 * subsets, eps and flags are shared by the levels.  kld_scale is a HOST array of n_subsets floats per level (or NULL). */
#define XHVED_POE_MAX_LEVELS 4
typedef struct {
  const float* mu;
  const float* logvar;
  int64_t n, expert_stride;
  const uint8_t* drop;
  int64_t per_sample;
  float* out_mu;
  float* out_logvar;
  const float* noise;
  float* out_z;
  float* kld_out;
} xhved_poe_level;
typedef struct {
  const float* mu;
  const float* logvar;
  int64_t n, expert_stride;
  const uint8_t* drop;
  int64_t per_sample;
  const float* g_mu;
  const float* g_logvar;
  const float* noise;
  const float* g_z;
  const float* kld_scale;
  float* d_mu;
  float* d_logvar;
} xhved_poe_level_grad;
int xhved_poe_fwd_levels(const xhved_poe_level* levels, int n_levels, const uint32_t* subset_masks, int n_subsets, float eps,
                         int flags, void* stream);
int xhved_poe_bwd_levels(const xhved_poe_level_grad* levels, int n_levels, const uint32_t* subset_masks, int n_subsets, float eps,
                         int flags, void* stream);

/* The same two launches with the model's `clip` fused in (RA_HVED.py:580 applies clip = clamp(logvar, -50, 50),
 * RA_HVED.py:749-753, to every modality's logvar while the experts are concatenated): with XHVED_POE_CLIP in flags the
 * logvar slabs of experts 1..4 hold the RAW (unclipped) values, the kernel clamps them to [clip_lo, clip_hi] as it loads
 * them, and the backward returns zero for d_logvar where the raw value lies outside the interval (torch.clamp's
 * gradient).  The prior slab is never clipped (the reference does not clip it either). */
#define XHVED_POE_CLIP 2
typedef struct {
  float eps;
  int flags;              /* XHVED_POE_STANDARD_PRIOR | XHVED_POE_CLIP */
  float clip_lo, clip_hi;
} xhved_poe_opts;
int xhved_poe_fwd_levels_opts(const xhved_poe_level* levels, int n_levels, const uint32_t* subset_masks, int n_subsets,
                              const xhved_poe_opts* opts, void* stream);
int xhved_poe_bwd_levels_opts(const xhved_poe_level_grad* levels, int n_levels, const uint32_t* subset_masks, int n_subsets,
                              const xhved_poe_opts* opts, void* stream);

/* clip on its own (RA_HVED.py:749-753): y = clamp(x, lo, hi) (NaN passes through, like torch.clamp); backward
 * dx = g where lo <= x <= hi, else 0. */
int xhved_clip_fwd(const float* x, int64_t n, float lo, float hi, float* y, void* stream);
int xhved_clip_bwd(const float* x, const float* g, int64_t n, float lo, float hi, float* dx, void* stream);

/* ZeroLayerF (buildingblocks.py:308-323; call sites RA_HVED.py:559, U_Hemis.py:42, buildingblocks.py:880-881):
 * y[b, :] = mask[b] ? 0 : x[b, :] for a (rows, per_row) fp32 tensor and a per-row uint8 mask.  The backward of the
 * reference is the same map applied to the incoming gradient, so one entry point serves both directions. */
int xhved_zero_rows(const float* x, const uint8_t* mask, int64_t rows, int64_t per_row, float* y, void* stream);

/* Per-channel Dice coefficient of the training loss (compute_per_channel_dice, loss.py:257-285, called from DiceLoss.forward,
 * loss.py:201-209): p, t fp32 (N, C, spatial) contiguous.  xhved_dice_sums accumulates into sums[3 C] (zeroed by the caller)
 * {sum p t, sum p^2, sum t^2} per channel in one pass; dice_c = 2 sums[3c] / max(sums[3c+1] + sums[3c+2], eps) is left to the
 * caller (C numbers).  xhved_dice_bwd: dp = g_dice[c] * d dice_c / d p (zero where the clamp is active). */
int xhved_dice_sums(const float* p, const float* t, int N, int C, int64_t spatial, float* sums, void* stream);
int xhved_dice_bwd(const float* p, const float* t, const float* sums, const float* g_dice, int N, int C, int64_t spatial, float eps,
                   float* dp, void* stream);

/* ---------------------------------------------------------------- InstanceNorm3d / BatchNorm3d + LeakyReLU (K6)
 * The normalisation of the convolution path fused with the LeakyReLU that follows it: SingleConv order 'ilc'
 * (buildingblocks.py:400-462: nn.InstanceNorm3d -> nn.LeakyReLU(0.01) -> Conv3d), BasicConv (buildingblocks.py:11-31), the
 * BatchNorm3d layers of modules/DuSFE.py:17-36,108-110,187.  x, y, dy, dx: (N, C, spatial) contiguous, element type `dtype`
 * (0 fp32, 1 fp16, 2 bf16: the types autocast hands over, train.py:207); statistics, gamma, beta in fp32.
 *   y = lrelu_slope((x - mean_g) * rstd_g * gamma_c + beta_c),  rstd_g = 1 / sqrt(biased var_g + eps);  slope = 1: no activation
 *   mode 0 (instance): group g = (n, c), mean / rstd are OUTPUTS of N*C entries;
 *   mode 1 (batch, training): group g = c over all samples, OUTPUTS of C entries (the caller updates running statistics);
 *   mode 2 (frozen, eval-mode BatchNorm): mean / rstd are INPUTS of C entries.
 * gamma / beta may be NULL (InstanceNorm3d's default affine=False).  partials: device scratch of
 * xhved_norm_act_workspace(...) bytes (not needed in mode 2 forward).
 * Backward recomputes the activation mask from x and the saved mean / rstd:
 *   g = dy * (pre > 0 ? 1 : slope);  dx = gamma_c rstd_g (g - mean_g(g) - xhat mean_g(g xhat))   (mode 2: dx = gamma_c rstd_c g)
 *   dgamma_c += sum g xhat, dbeta_c += sum g   (optional; fp32[C], zeroed by the caller). */
typedef struct xhved_norm_shape {
  int N, C;
  int64_t spatial;
  int mode;
  int dtype;
  float eps;
  float slope;
} xhved_norm_shape;
int64_t xhved_norm_act_workspace(int N, int C, int64_t spatial, int dtype);
int xhved_norm_act_fwd(const void* x, const float* gamma, const float* beta, const xhved_norm_shape* shape, float* mean, float* rstd,
                       void* partials, void* y, void* stream);
int xhved_norm_act_bwd(const void* x, const void* dy, const float* gamma, const float* beta, const float* mean, const float* rstd,
                       const xhved_norm_shape* shape, void* partials, void* dx, float* dgamma, float* dbeta, void* stream);

/* ---------------------------------------------------------------- spatial gate of AttenModule2 (K7)
 * buildingblocks.py:259-301: gate = sigmoid(conv1x1x1(depthwise_conv7x7x7(x)))  (enc_spatial -> enc_spatial2 at 283-285,
 * seg_spatial -> seg_spatial2 at 294-296).  The caller composes the two linear layers into ONE dense G -> 1 convolution,
 * w[g][kd][kh][kw] = sum_j w2[e g + j] W1[e g + j][kd][kh][kw] (e = channel expansion), bias = sum_o w2[o] b1[o] + b2:
 *   gate[n][v] = sigmoid(bias + sum_g sum_tap x[n][g][v + tap - 3] w[g][tap])      zero padding 3, stride 1
 * x: (N, G, D, H, W) fp32 contiguous, w: (G, 343) fp32, bias: device float[1] or NULL, gate / dgate: (N, 1, D, H, W).
 * Backward: dpre = dgate gate (1 - gate); dx (optional) = the transposed convolution of dpre; dw (optional, (G, 343)) and
 * dbias (optional, [1]) need `partials`, a device scratch of xhved_gate7_workspace(...) bytes (per-tile partial sums, reduced in a
 * second launch: deterministic). */
int64_t xhved_gate7_workspace(int N, int G, int D, int H, int W);
int xhved_gate7_fwd(const float* x, const float* w, const float* bias, int N, int G, int D, int H, int W, float* gate, void* stream);
int xhved_gate7_bwd(const float* x, const float* w, const float* gate, const float* dgate, int N, int G, int D, int H, int W,
                    void* partials, float* dx, float* dw, float* dbias, void* stream);

/* ---------------------------------------------------------------- depthwise 3x3x3 convolution (K8)
 * nn.Conv3d(C, C, 3, padding=1, groups=C) -- the conv of BasicConv(C, C, 3, padding=1, groups=C), RA_HVED.py:406,
 * buildingblocks.py:11-31.  x, y, dy, dx: (N, C, D, H, W) fp32 contiguous; w: (C, 27); bias: (C) or NULL.
 *   y[n][c][v] = bias[c] + sum_tap x[n][c][v + tap - 1] w[c][tap]           stride 1, zero padding 1
 * Backward: dx (optional) = correlation of dy with the flipped kernel; dw (optional, (C, 27)) and dbias (optional, (C)) need
 * `partials`, a device scratch of xhved_dwconv3_workspace(...) bytes (per-tile partial sums reduced in a second launch). */
int64_t xhved_dwconv3_workspace(int N, int C, int D, int H, int W);
int xhved_dwconv3_fwd(const float* x, const float* w, const float* bias, int N, int C, int D, int H, int W, float* y, void* stream);
int xhved_dwconv3_bwd(const float* x, const float* w, const float* dy, int N, int C, int D, int H, int W, void* partials, float* dx,
                      float* dw, float* dbias, void* stream);

/* ---------------------------------------------------------------- 1x1x1 convolution (K9)
 * nn.Conv3d(Cin, Cout, 1): the squeeze / fuse layers of DuSEAttention (modules/DuSFE.py:100-104), the 1x1x1 layers of the VU blocks and
 * final_conv (RA_HVED.py:483, 567-605); Cin, Cout <= 32 (XHVED_ERR_UNSUPPORTED_DIM above).  x: (N, Cin, vol), y / dy: (N, Cout, vol)
 * contiguous, element type `dtype` (0 fp32, 1 fp16, 2 bf16); w: (Cout, Cin) fp32, bias: (Cout) fp32 or NULL; fp32 accumulation.
 *   y[n][o][v] = bias[o] + sum_i w[o][i] x[n][i][v]
 * Backward: dx (optional, type `dtype`), dw (optional, (Cout, Cin) fp32), dbias (optional, (Cout) fp32); `partials`: device scratch of
 * xhved_pwconv_workspace(...) bytes, needed for dw / dbias when Cin, Cout <= 8. */
int64_t xhved_pwconv_workspace(int N, int Cin, int Cout, int64_t vol);
int xhved_pwconv_fwd(const void* x, const float* w, const float* bias, int N, int Cin, int Cout, int64_t vol, int dtype, void* y,
                     void* stream);
int xhved_pwconv_bwd(const void* x, const float* w, const void* dy, int N, int Cin, int Cout, int64_t vol, int dtype, void* partials,
                     void* dx, float* dw, float* dbias, void* stream);

/* ---------------------------------------------------------------- dense 3x3x3 convolution, few channels (K10)
 * nn.Conv3d(Cin, Cout, 3, stride 1, padding 1), one group, Cin, Cout <= 64 (XHVED_ERR_UNSUPPORTED_DIM above): the SingleConv /
 * DoubleConv layers of the encoders and decoders (buildingblocks.py:444-507).  x: (N, Cin, D, H, W), y / dy: (N, Cout, D, H, W)
 * contiguous, element type `dtype` (0 fp32, 1 fp16, 2 bf16); w: (Cout, Cin, 27) fp32; bias: (Cout) fp32 or NULL; fp32 accumulation.
 *   y[n][o][v] = bias[o] + sum_i sum_tap x[n][i][v + tap - 1] w[o][i][tap]
 * Backward: dx (optional, type `dtype`), dw (optional, (Cout, Cin, 27) fp32), dbias (optional); dw / dbias need `partials`, a device
 * scratch of xhved_conv3_workspace(...) bytes. */
int64_t xhved_conv3_workspace(int N, int Cin, int Cout, int D, int H, int W);
int xhved_conv3_fwd(const void* x, const float* w, const float* bias, int N, int Cin, int Cout, int D, int H, int W, int dtype, void* y,
                    void* stream);
int xhved_conv3_bwd(const void* x, const float* w, const void* dy, int N, int Cin, int Cout, int D, int H, int W, int dtype, void* partials,
                    void* dx, float* dw, float* dbias, void* stream);

/* reparametrize (RA_HVED.py:741-747): z = mu + noise * exp(0.5 logvar); and its backward. */
int xhved_reparam_fwd(const float* mu, const float* logvar, const float* noise, int64_t n, float* z, void* stream);
int xhved_reparam_bwd(const float* logvar, const float* noise, const float* g_z, int64_t n, float* d_mu, float* d_logvar, void* stream);

/* ---------------------------------------------------------------- mLSTM cell (K1)
 * Chunkwise forward over BH = B*NH sequences of nc chunks.  dh = true head dim (scale 1/sqrt(dh)),
 * dhp in {16,32,64,128}.  Outputs: h_tiles (bf16 tiles), m and den (fp32, (BH, nc*128)): the row
 * stabiliser (vision_lstm.py:111) and sum_s C_ts (the argument of :123).
 * Scratch / saved-for-backward buffers, all caller allocated:
 *   ws_dstate  fp32  BH*nc*dhp*(dhp+16)      ws_g, ws_amax  fp32  BH*nc
 *   states     bf16  2*BH*nc*dhp*(dhp+16)  (state ENTERING each chunk as a hi/lo pair of tile-native [dhp][dhp+16] tiles)
 *   m_prev     fp32  BH*nc               (its log-scale) */
int xhved_mlstm_fwd(const void* q_tiles, const void* k_tiles, const void* v_tiles, const float* ig_padded, const float* fg_padded,
                    int BH, int nc, int dh, int dhp, float eps, void* h_tiles, float* m, float* den, float* ws_dstate,
                    float* ws_g, float* ws_amax, void* states, float* m_prev, void* stream);

/* Backward.  dh_tiles: bf16 tiles of dL/dh.  states/m_prev/m/den/h_tiles as produced by the forward.
 * Outputs: dq, dk, dv as bf16 tiles in the layout of q, k, v (their consumers -- xhved_vil_pre_bwd, the weight-gradient
 * MMAs -- take bf16 operands; xhved_mlstm_unpack turns them into fp32 (BH,S,dh)); dig, dfg fp32 (BH, nc*128).
 * Scratch: ws_dstate/ws_g/ws_amax as in the forward, rstates bf16 2*BH*nc*dhp*(dhp+16), mu_next fp32 BH*nc,
 * ws_dc fp32 (BH, nc*128).  The (~1e-6 relative) gradient through the row-max stabiliser is dropped. */
int xhved_mlstm_bwd(const void* q_tiles, const void* k_tiles, const void* v_tiles, const float* ig_padded, const float* fg_padded,
                    const void* h_tiles, const void* dh_tiles, const float* m, const float* den, const void* states,
                    const float* m_prev, int BH, int nc, int dh, int dhp, float eps, void* dq, void* dk, void* dv, float* dig,
                    float* dfg, float* ws_dstate, float* ws_g, float* ws_amax, void* rstates, float* mu_next, float* ws_dc,
                    void* stream);

/* Layout conversion for the stand-alone cell entry point (parallel_stabilized_simple drop-in). */
int xhved_mlstm_pack(const float* src /* (BH,S,dh) */, int BH, int S, int dh, int dhp, void* tiles, void* stream);
int xhved_mlstm_pack_gates(const float* ig, const float* fg /* (BH,S) */, int BH, int S, float* ig_padded, float* fg_padded, void* stream);
int xhved_mlstm_unpack(const void* tiles, int BH, int S, int dh, int dhp, float* dst /* (BH,S,dh) */, void* stream);

/* Sizes of the caller-allocated buffers (host-only queries, no device needed).  The library never allocates: outputs, saved
 * state and scratch are passed in, and these two functions are the single source of truth for how large they must be.
 *   tile_bytes    each of q, k, v, h, dh, dq, dk, dv tiles     (BH * nc * 128 * dhp bf16)
 *   row_bytes     each of ig, fg, m, den, dig, dfg, ws_dc       (BH * nc * 128 fp32)
 *   dstate_bytes  ws_dstate                                     (BH * nc * dhp * (dhp+16) fp32)
 *   chunk_bytes   each of ws_g, ws_amax, m_prev, mu_next        (BH * nc fp32)
 *   states_bytes  each of states, rstates                       (BH * nc bf16 hi/lo pairs of dhp * (dhp+16)) */
typedef struct {
  int nc, dhp;
  int64_t tile_bytes, row_bytes, dstate_bytes, chunk_bytes, states_bytes;
} xhved_mlstm_workspace;
int xhved_mlstm_workspace_query(int BH, int S, int dh, xhved_mlstm_workspace* out);
/* ViL block of width C on B sequences of S tokens: its cell runs with BH = 4*B, dh = C/2;
 *   token_tile_bytes    each of act, z, xm, d_act, dz, ws_dconv, ws_dxmv    (B * nc bf16 "token tiles" [128][2C]: every
 *                       kernel-to-kernel tensor over the E = 2C inner channels, tile-native layout with R = 128 tokens of one
 *                       chunk in traversal order, tile index b*nc + chunk)
 *   grad_replica_stride floats per replica of the flat parameter-gradient buffer (sum of the 14 parameter sizes, padded) */
typedef struct {
  xhved_mlstm_workspace cell;
  int64_t token_tile_bytes;
  int64_t grad_replica_stride;
} xhved_vil_workspace;
int xhved_vil_workspace_query(int B, int S, int C, xhved_vil_workspace* out);

/* Optional per-kernel timing with CUDA events on the launching stream (off by default; used by bench.py to
 * attribute time to kernels).  xhved_profile_read synchronises the device, returns accumulated milliseconds
 * and launch counts per kernel id since the last read, and resets. */
int xhved_profile_enable(int on);
int xhved_profile_kernel_count(void);
const char* xhved_profile_kernel_name(int id);
int xhved_profile_read(float* ms, int* launches, int n);

/* Diagnostic: D[128][N] = A * B^T through tcgen05 with tile-native operands (see mlstm_fwd.cu). */
int xhved_umma_selftest(const void* a_tile, const void* b_tile, int N, int K, int a_mn, int b_mn, float* d, void* stream);
/* Diagnostic: clock64 cycles one thread needs to ISSUE `reps` tcgen05.mma (M = 128, bf16, K = 16 each, spread round-robin over
 * n_acc independent accumulators; A in shared or tensor memory; K- or MN-major operands) and until they have COMPLETED.
 * out2: device int64[2] = {issue cycles, issue + completion cycles}. */
int xhved_umma_issue_bench(int N, int reps, int n_acc, int a_in_tmem, int mn_major, long long* out2, void* stream);

/* ---------------------------------------------------------------- ViL block around the cell (K2, K3)
 * Parameter block: pointers to the reference's parameters (fp32, contiguous), named by state_dict key
 * relative to the ViLBlock. */
typedef struct xhved_vil_params {
  const float* norm_weight;      /* norm.weight                         (C)        */
  const float* proj_up_weight;   /* layer.proj_up.weight                (2E, C)    */
  const float* conv_weight;      /* layer.conv1d.conv.weight            (E, 1, 4)  */
  const float* conv_bias;        /* layer.conv1d.conv.bias              (E)        */
  const float* q_weight;         /* layer.q_proj.weight                 (E/QB, QB, QB) */
  const float* k_weight;         /* layer.k_proj.weight                            */
  const float* v_weight;         /* layer.v_proj.weight                            */
  const float* igate_weight;     /* layer.mlstm_cell.igate.weight       (NH, 3E)   */
  const float* igate_bias;       /* layer.mlstm_cell.igate.bias         (NH)       */
  const float* fgate_weight;     /* layer.mlstm_cell.fgate.weight       (NH, 3E)   */
  const float* fgate_bias;       /* layer.mlstm_cell.fgate.bias         (NH)       */
  const float* outnorm_weight;   /* layer.mlstm_cell.outnorm.weight     (E)        */
  const float* learnable_skip;   /* layer.learnable_skip                (E)        */
  const float* proj_down_weight; /* layer.proj_down.weight              (C, E)     */
} xhved_vil_params;

/* Gradients of the same parameters (fp32, accumulated with atomics: zero them first). */
typedef struct xhved_vil_grads {
  float* norm_weight;
  float* proj_up_weight;
  float* conv_weight;
  float* conv_bias;
  float* q_weight;
  float* k_weight;
  float* v_weight;
  float* igate_weight;
  float* igate_bias;
  float* fgate_weight;
  float* fgate_bias;
  float* outnorm_weight;
  float* learnable_skip;
  float* proj_down_weight;
} xhved_vil_grads;

/* Token geometry: x[b, n, c] = x_base[b*stride_b + n*stride_n + c*stride_c] (elements).  The NCDHW
 * bottleneck feature (UxLSTMEnc_3d.py:59) is stride_n = 1, stride_c = S; a (B,S,C) token tensor is
 * stride_n = C, stride_c = 1.  reverse != 0 = SequenceTraversal.ROWWISE_FROM_BOT_RIGHT. */
typedef struct xhved_vil_shape {
  int B, S, C;          /* tokens: batch, sequence, model dim; E = 2C inner dim                */
  int NH, QB;           /* cell heads (= qkv_block_size, vision_lstm.py:402-405) and block size */
  int reverse;
  int64_t x_stride_b, x_stride_n, x_stride_c;
  int64_t y_stride_b, y_stride_n, y_stride_c;
  /* Backward only: parameter gradients are accumulated with global atomics.  To spread that traffic over more L2
   * slices the caller may provide grad_replicas (R >= 1) zero-filled copies of the whole gradient block, replica r
   * starting grad_replica_stride elements after replica r-1 (the xhved_vil_grads pointers address replica 0); CTA i adds
   * into replica i % R, and xhved_reduce_replicas sums them.  R <= 1 disables it. */
  int grad_replicas;
  int64_t grad_replica_stride;
} xhved_vil_shape;

/* K2: LayerNorm -> proj_up -> causal conv -> SiLU -> q,k,v (tiles) + gates (padded) + act, z, xm.
 * act (conv activation), z (gate branch), xm (pre-conv x_mlstm, kept for the backward): bf16 token tiles (SURVEY 8d counts
 * the block's intermediates as bf16: 12 C + 16 E + 8 NH bytes per token for K2 + K3 forward). */
int xhved_vil_pre_fwd(const float* x, const xhved_vil_params* p, const xhved_vil_shape* sh, void* q_tiles, void* k_tiles,
                      void* v_tiles, float* ig_padded, float* fg_padded, void* act, void* z, void* xm, void* stream);
/* K3: outnorm(h) + skip*act, * silu(z), proj_down, + x residual -> y (same geometry family as x). */
int xhved_vil_post_fwd(const float* x, const void* h_tiles, const void* act, const void* z, const xhved_vil_params* p,
                       const xhved_vil_shape* sh, float* y, void* stream);
/* K3 backward: from dy computes dh (bf16 tiles), d_act_skip, dz (bf16 token tiles [128][E], token_tile_bytes each),
 * dx_residual is dy itself; accumulates outnorm / skip / proj_down gradients. */
int xhved_vil_post_bwd(const float* dy, const void* h_tiles, const void* act, const void* z, const xhved_vil_params* p,
                       const xhved_vil_shape* sh, void* dh_tiles, void* d_act, void* dz, const xhved_vil_grads* g, void* stream);
/* K2 backward: from the forward's saved xm, the cell's dq, dk, dv (bf16 tiles), dig, dfg (padded), d_act (skip path) and
 * dz (bf16 token tiles) computes dx = dy + d(branch)/dx (dy and dx both use the y_* strides of sh) and accumulates parameter
 * gradients.  q_tiles / k_tiles / v_tiles are accepted for ABI stability and not read: the gate-weight gradient is taken
 * through the block-diagonal projections ([dig|dfg]^T q = ([dig|dfg]^T act) Wq^T).  Scratch: ws_dconv, ws_dxmv, bf16 token
 * tiles (token_tile_bytes each).  Rows of dq / dk / dv / dig / dfg / d_act behind the end of a sequence (S not a multiple
 * of 128) must be zero, as xhved_mlstm_bwd and xhved_vil_post_bwd write them: the kernel does not mask them again.
 * Every kernel-to-kernel tensor of the backward is bf16: its consumers round to bf16 MMA
 * operands anyway, and the fp32 versions were 2.9 KB of HBM traffic per token and block (dim 32). */
int xhved_vil_pre_bwd(const float* x, const float* dy, const void* xm, const void* q_tiles, const void* k_tiles, const void* v_tiles,
                      const void* dq, const void* dk, const void* dv, const float* dig, const float* dfg, const void* d_act,
                      const void* dz, const xhved_vil_params* p, const xhved_vil_shape* sh, float* dx, const xhved_vil_grads* g,
                      void* ws_dconv, void* ws_dxmv, void* stream);

/* One whole ViL block per call (ViLBlock.forward, vision_lstm.py:494-502, and its backward): K2 -> cell -> K3 enqueued on
 * `stream` into caller-allocated blobs.  xhved_vil_block_workspace reports the blob sizes: `saved` is written by the forward
 * and read by the backward of the same call (it must stay untouched in between: one blob per live forward), `scratch` is
 * the backward's own; n_param_grads = the number of floats of `param_grads`, which receives the 14 parameter gradients
 * back to back in xhved_vil_params order.  sh->grad_replicas / grad_replica_stride as for xhved_vil_pre_bwd (the replicas
 * live inside `scratch`, zero-filled by the call).  Nothing is allocated and no state is kept between calls. */
int xhved_vil_block_workspace(int B, int S, int C, int grad_replicas, int64_t* saved_bytes, int64_t* scratch_bytes,
                              int64_t* n_param_grads);
int xhved_vil_block_fwd(const float* x, const xhved_vil_params* p, const xhved_vil_shape* sh, float eps, void* saved, float* y,
                        void* stream);
int xhved_vil_block_bwd(const float* x, const float* dy, const xhved_vil_params* p, const xhved_vil_shape* sh, float eps, void* saved,
                        void* scratch, float* dx, float* param_grads, void* stream);

/* ---------------------------------------------------------------- ViL blocks wider than the fused K2 / K3 kernels
 * dim 128 / 256 (E = 256 / 512, head dim 64 / 128).  The block's three Linear layers (LayerNorm +
 * proj_up, proj_down) stay plain GEMMs on the caller's side; these four kernels fuse everything between them and the cell.
 * Row-major tensors are in NATURAL token order (B, S, .), the direction flip is an index map (reverse != 0).
 *   up     (B, S, 2E) fp32   proj_up output [x_mlstm | z] (vision_lstm.py:427-428);  d_up: its gradient, same layout
 *   act    (B, S, E)  fp32   SiLU(conv(x_mlstm)), kept for the skip (437) and the backward
 *   hg     (B, S, E)  fp32   (outnorm(h) + skip * act) * SiLU(z): the row proj_down multiplies (437-443)
 *   q / k / v / h / dh / dq / dk / dv tiles, padded gates: as for xhved_mlstm_fwd / _bwd (dhp = E / 4)
 *   dg_rm  (B, S, 8)  fp32   [dig | dfg] in token order: the caller takes [dig|dfg]^T act and [dig|dfg]^T x_mlstm as two plain
 *                            GEMMs and applies the 4x4 projections to get the gate-weight gradient (as vil_pre.cu does)
 * Parameter gradients (g_*) are accumulated with atomics: zero them first. */
int xhved_vil_wide_pre_fwd(const float* up, const float* conv_w, const float* conv_b, const float* qw, const float* kw, const float* vw,
                           const float* igw, const float* igb, const float* fgw, const float* fgb, int B, int S, int E, int reverse,
                           void* q_tiles, void* k_tiles, void* v_tiles, float* ig_padded, float* fg_padded, float* act, void* stream);
int xhved_vil_wide_post_fwd(const void* h_tiles, const float* act, const float* up, const float* outnorm_w, const float* skip, int B,
                            int S, int E, int reverse, float* hg, void* stream);
int xhved_vil_wide_post_bwd(const float* dhg, const void* h_tiles, const float* act, const float* up, const float* outnorm_w,
                            const float* skip, int B, int S, int E, int reverse, void* dh_tiles, float* d_act, float* d_up,
                            float* g_outnorm_w, float* g_skip, void* stream);
int xhved_vil_wide_pre_bwd(const float* up, const float* conv_w, const float* conv_b, const float* qw, const float* kw, const float* vw,
                           const float* igw, const float* fgw, int B, int S, int E, int reverse, const void* dq_tiles,
                           const void* dk_tiles, const void* dv_tiles, const float* dig, const float* dfg, const float* d_act,
                           float* d_up, float* dg_rm, float* g_conv_w, float* g_conv_b, float* g_qw, float* g_kw, float* g_vw,
                           void* stream);

/* fp32 (rows, cols) row-major -> bf16 (rows, 3 cols): [hi | lo | hi] (b_side = 0) or [hi | hi | lo] (b_side != 0), hi = the value
 * rounded to bf16, lo = the rounded residual.  One bf16 GEMM of an a-side and a b-side operand over the tripled contraction
 * dimension is the 3-product hi*hi + lo*hi + hi*lo with fp32 accumulation (~16 mantissa bits): how the plain Linear layers of
 * the wide blocks run on the tensor cores through the library GEMM.  cols % 8 == 0. */
int xhved_split_hilo_cat(const float* x, int64_t rows, int cols, int b_side, void* out, void* stream);

/* dst[i] = sum_r src[r*stride + i], i < n  (reduction of the gradient replicas above). */
int xhved_reduce_replicas(const float* src, int replicas, int64_t stride, int64_t n, float* dst, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* XHVED_H_ */
