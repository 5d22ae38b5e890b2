/* tdvc_b200.h -- C ABI of libtdvc_b200.so: the sm_100a kernels behind td-vc-gan's
 * Generator / ConditionalInstanceNorm / Discriminator / GAN-loss hot path.
 *
 * The reference (vicpc00/td-vc-gan) has NO FFI of its own: it is pure PyTorch and every
 * arithmetic op below is an ATen call made from its nn.Modules.  Each entry point therefore
 * cites the reference call site (file:line, relative to the reference root) whose ATen op it
 * replaces; the Python host (td-vc-gan_b200/tdvc/ops.py) binds them with ctypes and wraps them
 * in torch.autograd.Functions, see INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers to DEVICE memory, sizes as int/int64, `stream` is a cudaStream_t passed
 *     as void*; no torch types.  The caller allocates every output and workspace.
 *   - activations are NCW, contiguous, fp32 (layout of the reference's tensors).
 *   - every function returns 0 on success, a negative tdvc_status otherwise;
 *     tdvc_last_error() gives a thread-local message.  Nothing synchronises the device.
 *   - no hidden global state except lazily cached function attributes / driver entry points.
 */
#ifndef TDVC_B200_H
#define TDVC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum tdvc_status { TDVC_OK = 0, TDVC_ERR_ARG = -1, TDVC_ERR_CUDA = -2, TDVC_ERR_UNSUPPORTED = -3 };
enum tdvc_pad_mode { TDVC_PAD_ZEROS = 0, TDVC_PAD_REFLECT = 1 };
enum tdvc_act { TDVC_ACT_NONE = 0, TDVC_ACT_LRELU = 1, TDVC_ACT_TANH = 2 };

/* Geometry of one Conv1d / ConvTranspose1d call.  For Conv1d: x[B,Cin,Tin] -> y[B,Cout,Tout],
 * weight [Cout, Cin/groups, K].  For ConvTranspose1d: x[B,Cin,Tin] -> y[B,Cout,Tout], weight
 * [Cin, Cout, K] (groups must be 1). */
typedef struct tdvc_conv_geom {
  int32_t B, Cin, Tin, Cout, Tout, K;
  int32_t stride, pad, dilation, groups;
  int32_t pad_mode;   /* tdvc_pad_mode; reflect only for Conv1d */
  float in_slope;     /* LeakyReLU slope applied to x as it is read (1.0f = identity): the
                         nn.LeakyReLU that precedes every conv in model/generator.py */
  int32_t out_act;    /* tdvc_act applied after bias (+residual) */
  float out_slope;    /* slope for TDVC_ACT_LRELU */
} tdvc_conv_geom;

const char* tdvc_last_error(void);
int tdvc_version(void);
/* kernels launched by this library in this process so far (bench.py reports the per-step delta) */
int64_t tdvc_launch_count(void);
/* 2 * multiply-accumulates handed to one kernel family by this process so far (operand channel padding included):
 * 0 conv_tc_fwd_k (tcgen05, one tile per CTA), 1 conv_tc_ws_k (weight-stationary), 2 conv_tc_wt_k (stacked weights),
 * 3 conv_tc_wgrad_k, 4 the fp32 CUDA-core conv kernels, 5 the fused MRF chain kernels.  bench.py divides the per-step
 * deltas by each family's measured time. */
double tdvc_flop_count(int family);
/* 1 if the current device is sm_100 (tcgen05/TMEM/TMA paths usable), 0 otherwise, <0 on error */
int tdvc_device_is_sm100(void);

/* ---- weight norm: torch.nn.utils.weight_norm (util/__init__.py:16-20, model/generator.py:14,
 *      model/discriminator.py:11): w[r,:] = g[r] * v[r,:] / ||v[r,:]||_2, rows = dim 0. */
int tdvc_weight_norm_fwd(const float* v, const float* g, float* w, float* inv_norm /*[rows]*/,
                         int rows, int cols, void* stream);
int tdvc_weight_norm_bwd(const float* dw, const float* v, const float* g, const float* inv_norm,
                         float* dv, float* dg, int rows, int cols, void* stream);
/* the forward for MANY weights in one launch (every weight of G and D once per step instead of ~530 launches):
 * table = device int64[n_weights][4] {v pointer, g pointer, offset of this w inside flat_w (floats), cols};
 * row_start = device int32[n_weights + 1], first global row of each weight; inv norms go to flat_inv[global row]. */
int tdvc_weight_norm_fwd_multi(const void* table, const void* row_start, int n_weights, int total_rows,
                               float* flat_w, float* flat_inv, void* stream);
/* the backward for MANY weights in one launch: table = device int64[n_weights][6] {v pointer, g pointer, offset of dw_j in
 * flat_dw, offset of dv_j in flat_dv, first row of dg_j in flat_dg, cols}; 1/||v|| is recomputed from v. */
int tdvc_weight_norm_bwd_multi(const void* table, const void* row_start, int n_weights, int total_rows,
                               const float* flat_dw, float* flat_dv, float* flat_dg, void* stream);

/* ---- Conv1d (model/generator.py:75-92,146-156,214-249,299-347; model/discriminator.py:17-38;
 *      depthwise Kaiser FIR F.conv1d at model/generator.py:165-168, model/discriminator.py:100-102).
 *      fwd: y = act(conv(lrelu_in(pad(x)), w) + bias + residual).  bias/residual may be NULL.      */
int tdvc_conv1d_fwd(const tdvc_conv_geom* g, const float* x, const float* w, const float* bias,
                    const float* residual, float* y, void* stream);
/* dx = d/dx of the above given dy = dL/d(pre-activation output).  `x` is needed only when
 * in_slope != 1 (sign mask); `ws` is a workspace of tdvc_conv1d_bwd_data_ws(g) floats (reflect
 * padding or strided polyphase staging), may be NULL when that returns 0. */
int64_t tdvc_conv1d_bwd_data_ws(const tdvc_conv_geom* g);
int tdvc_conv1d_bwd_data(const tdvc_conv_geom* g, const float* dy, const float* w, const float* x,
                         float* dx, float* ws, void* stream);
/* dbias[c] = sum_{b,t} dy[b,c,t] (OVERWRITTEN) */
int tdvc_bias_grad(const float* dy, float* dbias, int B, int C, int T, void* stream);
/* second half of bwd_data on its own (used after a tensor-core dgrad): folds the `halo` reflected samples of
 * stage[rows, T+2*halo] back and applies the input-LeakyReLU mask from x:  dx[rows, T]. x may be NULL when
 * in_slope == 1. */
int tdvc_pad_act_bwd(const float* stage, const float* x, float* dx, int64_t rows, int T, int halo, int reflect,
                     float in_slope, void* stream);
/* dw (same layout as w, OVERWRITTEN) and dbias (may be NULL) */
int tdvc_conv1d_bwd_weight(const tdvc_conv_geom* g, const float* dy, const float* x, float* dw,
                           float* dbias, void* stream);

/* ---- ConvTranspose1d (model/generator.py:311-315).  No fused input activation (in_slope must be 1). */
int tdvc_conv_transpose1d_fwd(const tdvc_conv_geom* g, const float* x, const float* w,
                              const float* bias, float* y, void* stream);
int tdvc_conv_transpose1d_bwd_data(const tdvc_conv_geom* g, const float* dy, const float* w,
                                   float* dx, void* stream);
int tdvc_conv_transpose1d_bwd_weight(const tdvc_conv_geom* g, const float* dy, const float* x,
                                     float* dw, float* dbias, void* stream);

/* ---- elementwise pieces of FiLMResnetBlock / MRFBlock / Discriminator
 *      (model/generator.py:96-111,186-194; model/discriminator.py:20,31,38) */
int tdvc_leaky_relu_fwd(const float* x, float* y, int64_t n, float slope, void* stream);
/* F0 track -> sine + noise excitation, the decoder's c_var conditioning (util/__init__.py:22-50 f0_to_excitation):
 * f0[rows, F] Hz (0 = unvoiced; the last frame is dropped), out[rows, (F-1)*step].  Per-sample angular frequency = the
 * reference's nearest / linear F.interpolate mix, phase = its cumulative sum (formed in fp64), voiced samples
 * 0.1 sin(phase + *phase0) + 0.003 noise_v, unvoiced samples noise_u * 0.003 * (0.1 / 0.009).  The random numbers are inputs
 * drawn by the caller in the reference's order: noise_v[rows, N]; noise_u full [rows, N] (u_offset = NULL) or compact, the
 * k-th unvoiced sample in row-major order taking noise_u[k], with u_offset[row] = unvoiced samples in the rows before
 * (tdvc_f0_unvoiced_count gives counts[rows]). */
int tdvc_f0_unvoiced_count(const float* f0, int* counts, int rows, int F, int step, float sampling_rate, int linear,
                           void* stream);
int tdvc_f0_excitation(const float* f0, const float* noise_v, const float* noise_u, const int64_t* u_offset,
                       const float* phase0 /* device scalar */, float* out, int rows, int F, int step, float sampling_rate,
                       int linear, void* stream);
/* YIN pitch estimation, util/yin.py:24-140 (`estimate`, hard search): signal[B, T] -> f0[B, n_frames] Hz (0 = no period below the
 * threshold).  Windows of frame_length = 2 * tau_max samples centred on multiples of frame_stride (zero padding),
 * n_frames = (max(T, frame_length) - 1) / frame_stride + 1; lags tau_min + 1 .. tau_max - 1 are searched. */
int tdvc_yin_estimate(const float* signal, float* f0, int B, int T, int n_frames, int frame_length, int frame_stride,
                      int tau_min, int tau_max, float threshold, float sample_rate, void* stream);
/* WaveNet gate of the SSL content encoder (model/ssl_encoder.py:7-14 fused_add_tanh_sigmoid_multiply):
 * y[B,H,T] = tanh(a[:, :H] + g[:, :H]) * sigmoid(a[:, H:] + g[:, H:]) for a, g [B,2H,T] (g optional);
 * backward: da[B,2H,T] from dy and the recomputed activations (the gradient of g is the same tensor). */
int tdvc_gate_fwd(const float* a, const float* g, float* y, int B, int H, int T, void* stream);
int tdvc_gate_bwd(const float* dy, const float* a, const float* g, float* da, int B, int H, int T, void* stream);
/* dz = dy * act'(.) evaluated from the activation OUTPUT y (lrelu: sign(y); tanh: 1-y^2) */
int tdvc_act_bwd_from_output(const float* dy, const float* y, float* dz, int64_t n, int act,
                             float slope, void* stream);
/* FiLM: y = h*(1+gamma)+beta with gb = [B, 2C, T] holding gamma (first C) then beta */
int tdvc_film_fwd(const float* h, const float* gb, float* y, int B, int C, int T, void* stream);
int tdvc_film_bwd(const float* dy, const float* h, const float* gb, float* dh, float* dgb,
                  int B, int C, int T, void* stream);
/* y = alpha * (a + b + c); b, c may be NULL */
int tdvc_add3_scale(const float* a, const float* b, const float* c, float* y, int64_t n,
                    float alpha, void* stream);
/* F.normalize(x, dim=1) on [B,C,T] (model/generator.py:271), eps 1e-12 */
int tdvc_l2norm_fwd(const float* x, float* y, float* inv /*[B,T]*/, int B, int C, int T, void* stream);
int tdvc_l2norm_bwd(const float* dy, const float* y, const float* inv, float* dx, int B, int C, int T,
                    void* stream);
/* c_first != 0: cat([c.unsqueeze(2).repeat(1,1,T), e], dim=1) (model/generator.py:387-399);
 * c_first == 0: cat([e, c...], dim=1) (model/generator.py:260-261,380-381).  out [B, Cc+Ce, T] */
int tdvc_cond_concat_fwd(const float* c /*[B,Cc]*/, const float* e /*[B,Ce,T]*/, float* out,
                         int B, int Cc, int Ce, int T, int c_first, void* stream);
/* dc[b,cc] = sum_t dout[b,cc,t] ; de = the e rows of dout (de may be NULL). */
int tdvc_cond_concat_bwd(const float* dout, float* dc, float* de, int B, int Cc, int Ce, int T,
                         int c_first, void* stream);

/* ---- instance norm / conditional instance norm (model/conditional_instance_norm.py:4-19;
 *      nn.InstanceNorm1d(affine=False): biased variance over T, eps).  gb may be NULL (plain
 *      instance norm) or [B,2C,Tg] with Tg in {1, T} (Linear- or Conv-produced gamma|beta).
 *      y = (1+gamma)*xhat + beta, optionally followed by LeakyReLU(out_slope) (1.0 = none).     */
int tdvc_instnorm_stats(const float* x, float* mean, float* rstd, int BC, int T, float eps, void* stream);
int tdvc_cin_apply_fwd(const float* x, const float* mean, const float* rstd, const float* gb, int Tg,
                       float* y, int B, int C, int T, float out_slope, void* stream);
/* dy -> dx, dgb (dgb may be NULL when gb is NULL).  y_act is the forward output (only read when
 * out_slope != 1). */
int tdvc_cin_apply_bwd(const float* dy, const float* x, const float* mean, const float* rstd,
                       const float* gb, int Tg, const float* y_act, float* dx, float* dgb,
                       int B, int C, int T, float out_slope, void* stream);

/* ---- AvgPool1d(4, 2, 1, count_include_pad=False) (model/discriminator.py:63,72) */
int tdvc_avgpool4s2_fwd(const float* x, float* y, int BC, int Tin, int Tout, void* stream);
int tdvc_avgpool4s2_bwd(const float* dy, float* dx, int BC, int Tin, int Tout, void* stream);

/* ---- frame view of a signal for the strided (de)convolutions whose kernel is a whole number of strides -- the
 *      encoder's downsampling Conv1d(k=2r, stride=r) and the decoder's ConvTranspose1d(k=2r, stride=r)
 *      (model/generator.py:214-249, 299-347) run as stride-1 convolutions over frames of s samples:
 *      out[b, p*C + c, q] = x[b, c, s*q + p - pad] (0 outside [0,T)), q < Tq;  depth_to_space is the inverse map
 *      y[b, c, u] = in[b, p*C + c, q], s*q + p = u + pad (0 when q >= Tq), u < Tout.  Each is the other's gradient. */
int tdvc_space_to_depth(const float* x, float* out, int B, int C, int T, int s, int pad, int Tq, void* stream);
int tdvc_depth_to_space(const float* in, float* y, int B, int C, int Tq, int s, int pad, int Tout, void* stream);

/* ---- x.gather(1, label) (model/discriminator.py:49-51): y[b,0,t] = x[b,label[b],t] */
int tdvc_select_channel_fwd(const float* x, const int64_t* label, float* y, int B, int C, int T, void* stream);
int tdvc_select_channel_bwd(const float* dy, const int64_t* label, float* dx /*zero-filled here*/,
                            int B, int C, int T, void* stream);

/* ---- the discriminator's output layer with the label gather folded in (model/discriminator.py:36 `self.output`, a
 *      weight-normed Conv1d(width, num_classes, 3, padding=1, bias=False), followed by :49-51 `x.gather(1, label)`):
 *      only the selected speaker's row is computed,
 *        y[b,0,t] = sum_{c,k} w[label[b], c, k] * x[b, c, t + k - pad]      (stride 1, zero padding, K in {1,3,5,7}, pad = (K-1)/2);
 *      a label outside [0, NC) yields zeros.  bwd: dx[B,C,T] (may be NULL) and dw[NC,C,K] (may be NULL; ACCUMULATED into,
 *      the caller zero-fills it: samples that share a label add into the same row). */
int tdvc_conv1d_select_fwd(const float* x, const float* w, const int64_t* label, float* y, int B, int C, int T, int NC,
                           int K, int pad, void* stream);
int tdvc_conv1d_select_bwd(const float* dy, const float* x, const float* w, const int64_t* label, float* dx, float* dw,
                           int B, int C, int T, int NC, int K, int pad, void* stream);

/* ---- losses (train.py:273-281,327-331 LSGAN F.mse_loss; util/losses.py:55-68 F.l1_loss).
 *      out_sum += scale * sum(...) (caller zeroes out_sum; scale carries 1/N and any lambda). */
int tdvc_sq_err_const_sum(const float* a, float target, float scale, float* out_sum, int64_t n, void* stream);
/* da = gscale[0] * 2*scale*(a-target) */
int tdvc_sq_err_const_bwd(const float* a, float target, float scale, const float* gscale, float* da,
                          int64_t n, void* stream);
int tdvc_abs_diff_sum(const float* a, const float* b, float scale, float* out_sum, int64_t n, void* stream);
/* da = gscale[0] * scale * sign(a-b) */
int tdvc_abs_diff_bwd(const float* a, const float* b, float scale, const float* gscale, float* da,
                      int64_t n, void* stream);
/* the same for up to TDVC_L1_MAX_JOBS tensor pairs in one launch each (all 30 feature maps of util/losses.py:55-68):
 * out_sum += sum_j scale_j * sum |a_j - b_j|;  da_j = scale_j * gscale[0] * sign(a_j - b_j).  Pointers 16-byte aligned. */
#define TDVC_L1_MAX_JOBS 32
typedef struct tdvc_l1_job { const float* a; const float* b; float* da; int64_t n; float scale; } tdvc_l1_job;
int tdvc_abs_diff_sum_multi(const tdvc_l1_job* jobs, int n_jobs, float* out_sum, void* stream);
int tdvc_abs_diff_bwd_multi(const tdvc_l1_job* jobs, int n_jobs, const float* gscale, void* stream);

/* ---- contrastive (InfoNCE) loss over content-embedding frames, util/losses.py:70-116: gather of in-utterance
 *      negatives + cosine similarities + cross-entropy in one kernel.  One direction per call: anchors A, positives
 *      P (both [B,C,T]), raw[B,T,N] = the reference's torch.randint(0, T-1) draws.  loss_sum[0] += scale * sum of
 *      per-frame CE; dA / dP (optional, [B,C,T]) += scale * gradient.  Caller zeroes the accumulators. */
int tdvc_contrastive_dir(const float* A, const float* P, const int64_t* raw, float* loss_sum, float* dA, float* dP,
                         int B, int C, int T, int N, float scale, void* stream);

/* ---- fused multi-tensor AdamW (torch.optim.AdamW at train.py:188-189): one launch for a whole
 *      parameter list.  ptr tables are DEVICE arrays of n_tensors pointers / sizes.             */
int tdvc_adamw_multi(float* const* params, const float* const* grads, float* const* exp_avg,
                     float* const* exp_avg_sq, const int64_t* sizes, int n_tensors, int64_t max_size,
                     float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                     float grad_scale, float* step_dev /* optional device step counter: incremented, then used
                     for the bias corrections instead of `step` (CUDA-graph replay) */, void* stream);
/*      the same update with the work cut into equal blocks by the caller: blocks[n_blocks][2] = (tensor index, first element),
 *      a device array of int32 pairs, one CTA per block of 4096 elements (the form FusedAdamW uses: a 5 M-element layer next
 *      to 700 small tensors is spread over 1280 CTAs instead of 64). */
int tdvc_adamw_blocks(float* const* params, const float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                      const int64_t* sizes, const int32_t* blocks, int n_blocks, float lr, float beta1, float beta2, float eps,
                      float weight_decay, int step, float grad_scale, float* step_dev, void* stream);

/* ---- bf16 tensor-core path (tcgen05 / TMEM / TMA), dense stride-1 Conv1d as implicit GEMM.
 *      pack: x[B,C,T] fp32 NCW -> xp[B, Tp, Cp] bf16 channels-last with LeakyReLU(in_slope) and
 *      `halo` samples of reflect/zero padding on each side (Tp = T + 2*halo, Cp = C rounded up to
 *      8, zero filled).  */
int tdvc_pack_cl_bf16(const float* x, void* xp, int B, int C, int T, int Cp /* row pitch of xp */, int halo,
                      int pad_mode, float in_slope, float* chan_sum /* optional [C]: sum over (b,t) of x, OVERWRITTEN:
                      the bias gradient falls out of packing dL/dy */, int c_off /* first channel written */,
                      int Cw /* channels written (>= C, zero filled; <= 0: Cp - c_off) */,
                      int ones_ch /* >= C: that channel is set to 1 on the T valid rows (bias gradient through
                      the wgrad GEMM); -1: none */,
                      const float* film_gb /* optional [B,2C,T]: x*(1+gamma)+beta is applied before the LeakyReLU
                      (FiLM + activation + pack in one pass, generator.py:104-108) */, void* stream);
/*      the same pass on dL/dy of a layer that ends in LeakyReLU(slope) (model/discriminator.py:15-37: every layer of D):
 *      dyp = bf16 channels-last of dy * (y > 0 ? 1 : slope) -- the activation's backward applied while packing, chan_sum =
 *      the bias gradient.  Replaces torch's leaky_relu_backward + the pack. */
/*      the decoder's conditioning cat([c.unsqueeze(2).repeat(1, 1, T), e], dim=1) (model/generator.py:387-399; c[B,Cc] the
 *      speaker code, e[B,Ce,T] the excitation level) as the packed operand of the cond_var convs, without the fp32 tensor:
 *      cp[B, T, Cg] bf16, channel Cc+Ce = 1 (bias-gradient channel), zeros up to Cg. */
int tdvc_cond_pack_cl(const float* c, const float* e, void* cp, int B, int Cc, int Ce, int T, int Cg, void* stream);
int tdvc_pack_cl_bf16_masked(const float* dy, const float* y /* the layer's output, same shape */, float slope, void* dyp,
                             int B, int C, int T, int Cp, int halo /* zero rows on both sides: dyp[B, T + 2*halo, Cp] */,
                             float* chan_sum /* optional [C], OVERWRITTEN */, void* stream);
/* w[Cout,Cin,K] fp32 -> wp[K, Coutp, Cinp] bf16 (zero padded); transpose_flip!=0 produces the
 * dgrad operand wp[K, Cinp, Coutp] with taps reversed. */
int tdvc_pack_weight_bf16(const float* w, void* wp, int Cout, int Cin, int K, int Coutp, int Cinp,
                          int transpose_flip, int R_total, int r_off, int Q_total, int q_off, void* stream);
/* (R_total, r_off, Q_total, q_off): write the packed block at row r_off / column q_off of a larger
 * wp[K][R_total][Q_total] holding several convs' weights (grouped launches); <= 0 totals mean "the block is all". */
/* a list of small packs in one launch (job table by value): kind 0 = the weight block of tdvc_pack_weight_bf16 (src fp32
 * [Cout][Cin][K] -> rows [r_off, r_off+Rp) x columns [q_off, q_off+Qp) of dst bf16 [K][R_total][Q_total]; flip = transposed
 * with reversed taps), kind 1 = Rp floats of a bias vector at dst: the first Cout from src (may be NULL), zeros after. */
#define TDVC_PACK_MAX_JOBS 40
typedef struct tdvc_pack_job {
  const void* src;
  void* dst;
  int32_t kind, Cout, Cin, K, Rp, Qp, flip, R_total, r_off, Q_total, q_off, pad_;
} tdvc_pack_job;
int tdvc_pack_jobs(const tdvc_pack_job* jobs, int n_jobs, void* stream);
/* MANY weights in one launch: jobs = device int64[n_jobs][8] {offset of w inside flat_w (floats), offset of wp inside
 * flat_wp (bf16 elements), Cout, Cin, K, Rp, Qp, transpose_flip}; job output layout [K][Rp][Qp] as above. */
int tdvc_pack_weight_bf16_multi(const void* jobs, int n_jobs, int blocks_per_job, const float* flat_w, void* flat_wp,
                                void* stream);
/* y[B,Cout,Tout] fp32 NCW = epilogue( sum_k sum_ci xp[b, t + k*dilation + off, ci] * wp[k, co, ci] )
 * epilogue: + bias[co] ; FiLM  y*(1+gb[b,co,t]) + gb[b,Cout+co,t] when gb != NULL ; + residual ; act. */
int tdvc_conv1d_tc_fwd(const void* xp, const void* wp, const float* bias, const float* gb,
                       const float* residual, float* y, int B, int Cinp, int Tp, int Cout, int Coutp,
                       int Tout, int K, int dilation, int t_off, int out_act, float out_slope,
                       void* stream);

/* Grouped / packed-operand form of the same kernel.  All `groups` share the time geometry; group g reads input
 * channels [a_ch_off + g*a_ch_stride, +Cinp_g) of xp[B,Tp,Cp_total], uses weight rows [g*Coutp_g, +Coutp_g) of
 * wp[K][groups*Coutp_g][Cinp_g] and bias[g*bias_stride + n].  Output: fp32 NCW y[groups][B][Cout_g][Tout], or
 * (out_packed) bf16 channels-last yp[B][tp_out][cp_out] at row t+out_halo, channel out_ch_off + g*out_ch_stride + n
 * -- the operand layout of the next conv, so no pack pass.  maskp (optional): a packed activated tensor
 * (row t+mask_halo, channel mask_ch_off + g*mask_ch_stride + n); where it is <= 0 the result is multiplied by
 * mask_slope: the LeakyReLU derivative fused into a data-gradient conv (generator.py:89 cond_var[1]). */
typedef struct tdvc_tc_conv {
  const void* xp; const void* wp; const float* bias; const float* gb; const float* residual;
  float* y; void* yp; const void* maskp;
  int32_t B, Tp, Tout, K, dilation, t_off;
  int32_t Cp_total, groups, a_ch_off, a_ch_stride, Cinp_g, Cout_g, Coutp_g, bias_stride;
  int32_t out_act; float out_slope;
  int32_t out_packed, tp_out, cp_out, out_halo, out_ch_off, out_ch_stride;
  int32_t tm, cm, mask_halo, mask_ch_off, mask_ch_stride; float mask_slope;
  /* fp32 output layout: element (group g, batch b, channel n, step t) at g*y_grp_stride + b*y_b_stride + n*Tout + t.
   * Both 0 = group-major [groups][B][Cout_g][Tout]; y_grp_stride = Cout_g*Tout, y_b_stride = groups*Cout_g*Tout is the
   * ordinary NCW tensor [B][groups*Cout_g][Tout] of a grouped nn.Conv1d (model/discriminator.py:26-30). */
  int64_t y_grp_stride, y_b_stride;
  /* Chain epilogues: an MRF stage (model/generator.py:69-111,175-194) kept in bf16 channels-last between its
   * convolutions; group = kernel-size branch, channel counts multiples of 16, residual stream fp32 [branch][B][C][T].
   *   chain_mode 3 (conv.1):     h0 = conv + bias -> yp2 (optional);  leaky_relu(h0 * (1 + gamma) + beta, out_slope) -> yp
   *                              (gamma|beta = gb + g*gb_grp_stride, fp32 [B][2*Cout_g][Tout]; no FiLM when gb is NULL)
   *   chain_mode 4 (posconv.1):  x' = conv + bias + residual[g*res_grp_stride ...] -> y (y_*_stride layout);
   *                              leaky_relu(x', pk_slope) -> yp rows out_halo + t plus the reflect halo rows (yp optional)
   *   chain_mode 5 (posconv^T):  d = conv * lrelu'(maskp);  dgamma = d * auxp, dbeta = d -> dgbp[b][t][dgb_ch_off +
   *                              g*dgb_ch_stride + {n, Cout_g + n}];  d * (1 + gamma) -> yp   (gb NULL: d -> yp)
   *   chain_mode 6 (conv.1^T):   rows = positions of the padded input (t_valid + 2*halo); d = conv * lrelu'(maskp);
   *                              interior rows: d + residual -> y and yp (optional); halo rows -> halo_buf[g][B][Cout_g][2*halo]
   *                              (folded onto the rows they mirror by tdvc_chain_fold)
   * kg[g] > 0 (groups <= 4): group g uses the centred kg[g] of the K taps. */
  int32_t chain_mode;
  int32_t kg[4];
  int64_t gb_grp_stride, res_grp_stride;
  void* yp2; const void* auxp; void* dgbp;
  int32_t dgb_cp, dgb_ch_off, dgb_ch_stride;
  float* halo_buf;
  int32_t halo, t_valid;
  float pk_slope;
  /* unframe_s > 0 (fp32 output, no chain mode): the launch is the data gradient of a Conv1d(k, stride = s, padding = pad)
   * run over frames (tdvc_frame_pack_bf16): output channel f of row q is sample u = s*q + f % s - unframe_pad of channel f / s,
   * written to y[b, f / s, u] of a [B, unframe_C, unframe_T] tensor (rows outside [0, unframe_T) are dropped) -- the
   * inverse frame view in the epilogue instead of tdvc_frame_unpack.  s must divide 16. */
  int32_t unframe_s, unframe_pad, unframe_T, unframe_C;
  /* flat_tp > 0 (fp32 output, groups = 1, no chain / mask / residual): short sequences run batch-flattened -- the packed input
   * [B_real, flat_tp, Cp] with flat_halo zero rows on both sides of every sample is handed over as ONE sequence (B = 1,
   * Tp = Tout = B_real * flat_tp), so a 128-row tile holds several samples instead of one sample's 9 .. 35 time steps (the
   * discriminator's 1024-channel tail, model/discriminator.py:31-38, and the T/320 convs of the generator); output row r is
   * time step r % flat_tp - flat_halo of sample r / flat_tp and is stored to y[B_real, Cout_g, flat_T] when inside [0, flat_T). */
  int32_t flat_tp, flat_halo, flat_T;
} tdvc_tc_conv;
int tdvc_conv1d_tc_fwd_ex(const tdvc_tc_conv* c, void* stream);
/* after a chain_mode 6 launch: y[g][b][c][mirror(r)] += halo_buf[g][b][c][r] for the 2*halo reflect rows r of every
 * (group, batch, channel) and, when yp is given, the packed bf16 copy of the touched samples is rewritten. */
int tdvc_chain_fold(const float* halo_buf, float* y, void* yp, int groups, int B, int C, int T, int halo,
                    int64_t y_grp_stride, int64_t y_b_stride, int cp_out, int out_ch_off, int out_ch_stride, void* stream);

/* Stacked-block form for n convs that read the SAME input -- the FiLM cond_var[0] convs of the 9 blocks of an MRF stage
 * (generator.py:85-92,141-194).  The GEMM is turned round: the stacked weights are the M=128 tcgen05 operand (resident in
 * shared memory), 256 time steps are N.  wp[K][R][Cinp] bf16 holds the blocks' rows densely, block j = rows
 * [j*rows_per_block, +rows_per_block), R >= n_blocks*rows_per_block (rows beyond are never read); bias[n_blocks *
 * rows_per_block] or NULL.  Input: channels [a_ch_off, +Cinp) of xp[B,Tp,Cp_total], row t + k*dilation + t_off
 * (out-of-range rows read as zero).  Output, packed bf16 channels-last only (the next conv's operand):
 * yp[b, t+out_halo, out_ch_off + j*out_ch_stride + r] = act(conv + bias) for row r of block j; the pad columns
 * [rows_per_block, out_ch_stride) of every block are written as zeros.  out_act: TDVC_ACT_NONE / TDVC_ACT_LRELU.
 * Needs Cinp >= 64 and K * 128 * Cinp * 2 bytes of weights + two activation stages within 227 KB of shared memory. */
int tdvc_conv1d_tc_fwd_stacked(const void* xp, const void* wp, const float* bias, void* yp, int B, int Cp_total,
                               int a_ch_off, int Cinp, int Tp, int Tout, int K, int dilation, int t_off, int R,
                               int n_blocks, int rows_per_block, int out_act, float out_slope, int tp_out,
                               int cp_out, int out_halo, int out_ch_off, int out_ch_stride, void* stream);

/* weight gradient of the same conv on tcgen05: dw[Cout,Cin,K] (OVERWRITTEN, fp32) from the packed bf16 operands
 * dyp[B,Tout,Cdp] and xp[B,Tp,Cp]; xp row read for output step t and tap k is t + k*dilation + t_off.
 * ws: workspace of tdvc_conv1d_tc_wgrad_ws(Cout, Cin, K) floats (split-K partial sums, [K][Coutp][Cinp] + a bias row). */
int64_t tdvc_conv1d_tc_wgrad_ws(int Cout, int Cin, int K);
int tdvc_conv1d_tc_wgrad(const void* dyp, const void* xp, float* dw, float* ws, int B, int Cdp, int Tout, int Cp,
                         int Tp, int Cout, int Cin, int K, int dilation, int t_off,
                         int x_ch_off /* first channel of the conv's input inside xp */,
                         int dy_ch_off /* first channel of the conv's output inside dyp */,
                         float* db /* optional: bias gradient [Cout] = sum over (b,t) of dL/dy, from the same GEMM */,
                         int ws_is_zero /* != 0: ws holds zeros on entry and is left zeroed on return (a persistent
                                           workspace: no memset per call); 0: ws is scratch, cleared here */,
                         void* stream);

/* Weight gradients of up to 65535 independent conv groups in one tcgen05 launch (conv_tc2.cu):
 *   dW_g[co, ci, tap] = sum_{b,t} dyp[b, t, dy_ch_off + g*dy_ch_stride + co] * xp[b, t + tap*dilation + t_off_g, x_ch_off + g*x_ch_stride + ci]
 * dyp[B,Tout,Cdp], xp[B,Tp,Cp] bf16 channels-last.  per_group != 0 (ngroups <= 4): group g has kg[g] <= K taps, row offset
 * t_off[g] and its own output tensors dw[g] ([Cout][Cin][kg[g]], OVERWRITTEN) / db[g] -- the three kernel sizes of an MRF
 * depth (model/generator.py:175-194).  per_group == 0: K taps and t_off[0] for every group, outputs dw[0] + g*dw_grp_stride,
 * db[0] + g*db_grp_stride.  frame_s > 0 (per_group == 0): the groups are bundles of `sub` conv groups of a
 * Conv1d(k = kreal, stride = frame_s, groups) seen as a stride-1 conv over frames (xp from tdvc_frame_pack_bf16: Cin = sub *
 * cin_conv_g * frame_s frame channels per bundle, Cout = sub * cout_conv_g); dw[0] is then that conv's own gradient
 * [ngroups*Cout][cin_conv_g][kreal] and db[0] its [ngroups*Cout] bias gradient (model/discriminator.py:26-30).
 * want_bias: also produce the bias gradients (sum over (b,t) of dyp), through one more accumulator fed by a tile of ones.
 * ws: tdvc_conv1d_tc_wgrad2_ws(c) floats; ws_is_zero != 0: it holds zeros on entry and is left zeroed (persistent). */
typedef struct tdvc_tc_wgrad2 {
  const void* dyp; const void* xp; float* ws;
  float* dw[4]; float* db[4];
  int64_t dw_grp_stride, db_grp_stride;
  int32_t B, Cdp, Tout, Cp, Tp, Cout, Cin, K, dilation;
  int32_t ngroups, per_group, x_ch_off, x_ch_stride, dy_ch_off, dy_ch_stride;
  int32_t kg[4], t_off[4];
  int32_t want_bias, ws_is_zero;
  int32_t haloed;      /* -1 default; 1: one haloed x tile per time unit, taps = row-shifted views; 0: one copy per tap */
  int32_t tapsm;       /* -1 default (on when Cin is 16, 32 or 64); 0: channels only on the M lanes; 1: 128 / Cin taps per MMA */
  int32_t swap;        /* -1 default (operand roles swapped, M = co / N = ci, when 128 < Cin <= 256 and all taps fit TMEM:
                          the FiLM conditioning convs); 0: never */
  int32_t frame_s, kreal, cin_conv_g, sub;
} tdvc_tc_wgrad2;
int64_t tdvc_conv1d_tc_wgrad2_ws(const tdvc_tc_wgrad2* c);
int tdvc_conv1d_tc_wgrad2(const tdvc_tc_wgrad2* c, void* stream);

/* Channel-major frame view of a signal for the grouped strided convs (model/discriminator.py:26-30, k = 41, stride 4):
 * xf[b, q, c*s + p] = x[b, c, s*q + p - pad] (bf16 channels-last, 0 outside the signal), q < Tq; and its inverse for the
 * data gradient: dx[b, c, u] = dxf[b, c*s + p, q] with s*q + p = u + pad, dxf fp32 [B, C*s, Tq]. */
int tdvc_frame_pack_bf16(const float* x, void* xf, int B, int C, int T, int s, int pad, int Tq, void* stream);
int tdvc_frame_unpack(const float* dxf, float* dx, int B, int C, int T, int s, int pad, int Tq, void* stream);

/* Operand packs in one launch each.  tdvc_chain_pack: the grouped bf16 operands of one MRF depth from the G <= 4 branches'
 * fp32 weights w[g][C][C][k[g]] (model/generator.py:175-194): wp[Kmax][G*C][C] (forward, taps centred inside Kmax),
 * wtp[Kmax][G*C][C] (data gradient: transposed, taps reversed) and bias_out[G*C] (bias[g] may be NULL).
 * tdvc_frame_weights_pack: the bundled frame-view operands of a grouped strided conv w[Cout][cin_g][K], stride s, m =
 * ceil(K/s) frame taps, bundles of `sub` conv groups (cin_b = sub*cin_g*s frame channels in, cout_b outputs):
 * wp[m][Cout][cin_b] and wtp[m][(Cout/cout_b)*cin_b][cout_b] (model/discriminator.py:26-30). */
int tdvc_chain_pack(const float* const* w, const float* const* bias, const int* k, int G, int C, int Kmax, void* wp,
                    void* wtp, float* bias_out, void* stream);
int tdvc_frame_weights_pack(const float* w, void* wp, void* wtp, int Cout, int cin_g, int K, int s, int m, int sub,
                            int cin_b, int cout_b, void* stream);

/* ---- log-mel spectrogram loss without cuFFT (util/losses.py:28-53; mel.cu).  The STFT of the 18 frames is a GEMM
 *      against a (window x cos | -sin) basis run on the convolution kernels above; these are the pieces around it.
 * frames: F[k][b*NF + n] = win[k] * x[b][reflect(n*hop + k - pad)] (fp32, [n_fft][B*NF]) and its adjoint dx[B][T];
 *         split != 0: three row blocks [hi | lo | hi], hi = bf16(F), lo = bf16(F - hi), for a bf16 GEMM against the basis
 *         blocks [B_hi | B_hi | B_lo] with fp32-class accuracy (the adjoint sums the gradients of blocks 0 and 2);
 * power:  P[f][col] = S[f][col]^2 + S[im_off + f][col]^2 over the stacked (re | im) rows of the GEMM output S[rows][cols];
 * log_clamp: y = log(max(x, floor)), gradient dy / x where x > floor else 0 (torch.clamp + torch.log). */
int tdvc_stft_frames_fwd(const float* x, const float* win, float* F, int B, int T, int n_fft, int hop, int pad, int NF,
                         int split, void* stream);
int tdvc_stft_frames_bwd(const float* dF, const float* win, float* dx, int B, int T, int n_fft, int hop, int pad, int NF,
                         int split, void* stream);
int tdvc_power_fwd(const float* S, float* P, int nfreq, int im_off, int64_t cols, int split /* P = [hi | lo | hi] */,
                   void* stream);
int tdvc_power_bwd(const float* S, const float* dP, float* dS, int nfreq, int im_off, int rows_total, int64_t cols,
                   int split, void* stream);
int tdvc_log_clamp_fwd(const float* x, float* y, int64_t n, float floor_, void* stream);
int tdvc_log_clamp_bwd(const float* x, const float* dy, float* dx, int64_t n, float floor_, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TDVC_B200_H */
