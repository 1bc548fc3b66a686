/* a2c_b200.h -- C ABI of liba2c_b200.so: the B200 (sm_100a) implementation of
 * the A2C rollout-post-processing + update hot path of grantsrb/PyTorch-A2C.
 *
 * The reference is pure Python and has no FFI for this path; each entry point
 * below names the reference interface (file:line under the reference root) whose
 * arithmetic it replaces.  INTEGRATION.md shows the ctypes binding a reference
 * maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless
 *     the parameter name ends in `_host`;
 *   - `stream` is a cudaStream_t passed as void*; kernels are enqueued on it and
 *     the call returns without synchronising (the *_host entry points are the
 *     exception: they synchronise `stream` before returning);
 *   - no allocation, no global state besides cached TMA descriptors and an
 *     internally owned staging buffer for the *_host calls;
 *   - returns 0 on success, A2C_ERR_ARG (-1) for a rejected argument, else the
 *     cudaError_t of the failing runtime call; a2c_last_error() gives the text;
 *   - rollout buffers are flat and rollout-major: index n = r*T + t
 *     (a2c/training.py:87-101, a2c/runner.py:196).
 */
#ifndef A2C_B200_H
#define A2C_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define A2C_ERR_ARG (-1)
#define A2C_ABI_VERSION 3

int a2c_abi_version(void);
const char* a2c_last_error(void);
/* Number of kernel launches issued through this library since load (all threads). */
long long a2c_launch_count(void);

/* ------------------------------------------------------------------------
 * Stage 1 -- observation prep + frame stack
 * replaces a2c/preprocessing.py:11-17 (pong_prep), a2c/utils.py:26-43
 * (next_state) and the state store of a2c/runner.py:199.
 *
 * Plane ring (HBM layout): planes u8 [R, P, 80*80], pflag u8 [R, P].  Plane
 * index j holds the prepped frame j - (fs-1); state t of slot r is the window
 * planes[r, t .. t+fs-1] (oldest first).  pflag[r, j] != 0 marks a frame that
 * started a fresh stack (env reset): older planes of any window containing j
 * read as zero, which reproduces next_state's fs-1 zero planes.
 * ---------------------------------------------------------------------- */

/* frames u8 [R, F, 210,160,3], reset u8 [R, F] -> planes[r, j0+f], pflag[r, j0+f]
 * for f < F.  `variant`: 0 = TMA-staged, 2-deep rings, two CTAs per SM (default); 1 = direct
 * 128-bit loads; 2 = TMA-staged, 4-deep ring, one CTA per SM. */
int a2c_obs_prep_stack(const uint8_t* frames, const uint8_t* reset, uint8_t* planes,
                       uint8_t* pflag, int R, int F, int P, int j0, int variant, void* stream);

/* Same, frames/reset in pinned (page-locked, mapped) HOST memory.  Zero-copy: the stage-1
 * TMA boxes are fetched straight from host memory, so only the 80 selected rows of each frame
 * (38 % of its bytes) cross PCIe.  Synchronises `stream` before returning (the caller may
 * reuse the host buffers immediately): the per-step acting path. */
int a2c_obs_prep_stack_host(const uint8_t* frames_host, const uint8_t* reset_host,
                            uint8_t* planes, uint8_t* pflag, int R, int F, int P, int j0,
                            void* stream);

/* Pipelined host ingest for whole rollouts (the update path): ONE asynchronous copy per
 * rollout on a copy stream, overlapped with compute, then stage 1 from the device staging
 * buffer.  `layout` selects what crosses PCIe and how the staging buffer is laid out:
 *   0: the 80 selected rows (35, 37, ..., 193) of each frame, u8 [n, 80, 480]: 38,400 B per
 *      frame, one cudaMemcpy3DAsync with 480-byte rows;
 *   1: the cropped band (rows 35..193) of each frame, u8 [n, 159, 480]: 76,320 contiguous
 *      bytes per frame, one cudaMemcpy2DAsync;
 *   2: whole frames, u8 [n, 210, 160, 3]: one contiguous cudaMemcpyAsync.
 * a2c_obs_ingest_rows_async does not synchronise: the pinned host buffer must stay untouched
 * until the caller's event on `stream` completes.  a2c_obs_prep_stack_rows is
 * a2c_obs_prep_stack reading the staging layout. */
long long a2c_obs_ingest_bytes_per_frame(int layout);
int a2c_obs_ingest_rows_async(const uint8_t* frames_host, uint8_t* rows, long long n_frames,
                              int layout, void* stream);
int a2c_obs_prep_stack_rows(const uint8_t* rows, const uint8_t* reset, uint8_t* planes,
                            uint8_t* pflag, int R, int F, int P, int j0, int layout, void* stream);

/* The reference's other prep_fxn callables as batched device kernels (variants of the stage-1
 * kernel for non-Pong environments), bit-identical to the numpy arithmetic:
 *   kind 1  normalize_prep (a2c/preprocessing.py:4-6):  frames u8 [n,H,W,C] -> f64 [n,H,W,C]
 *   kind 2  snake_prep     (a2c/preprocessing.py:25-32): frames u8 [n,H,W,C>=3] -> f32 [n,H,W]
 *   kind 3  breakout_prep  (a2c/preprocessing.py:19-23): frames u8 [n,H>=195,W,C] -> u8 [n,80,(W-16+1)/2]
 * (null_prep, :8-9, adds an axis and needs no kernel). */
int a2c_prep_other(const uint8_t* frames, void* out, int kind, long long n_frames, int H, int W,
                   int C, void* stream);

/* Moves the last `fs` planes (+flags) of every slot to the head of its ring:
 * the bookmark state of a2c/runner.py:246. */
int a2c_ring_carry(uint8_t* planes, uint8_t* pflag, int R, int P, int fs, void* stream);

/* Strided description of N = R*T stacked states; used by every first layer. */
typedef struct {
  const void* base;        /* u8 or f32 elements */
  int dtype;               /* 0 = u8, 1 = f32 */
  int fs;                  /* channels (n_frame_stack) */
  int T;                   /* steps per slot: n = r*T + t */
  int reserved;
  long long stride_r;      /* elements between slots */
  long long stride_t;      /* elements between consecutive steps */
  long long stride_c;      /* elements between channels */
  const uint8_t* pflag;    /* NULL, or reset flags of the ring */
  long long pflag_stride_r;
} a2c_state_view;

/* Writes states (N, fs, 80, 80) contiguous, as f32 (out_dtype 1; the reference's
 * `states` tensor, a2c/training.py:94) or u8 (out_dtype 0). */
int a2c_states_materialize(const a2c_state_view* view, void* out, int out_dtype, int R,
                           void* stream);

/* ------------------------------------------------------------------------
 * Acting -- replaces a2c/runner.py:94-97 + a2c/utils.py:45-60.
 * x f32 [B, A] (logits, or probabilities when is_prob != 0), u f32 [B] ->
 * actions i64 [B] (-1 when the fp32 running sum never reaches u).
 * ---------------------------------------------------------------------- */
int a2c_sample_action(const float* x, const float* u, long long* actions, int B, int A,
                      int is_prob, void* stream);

/* ------------------------------------------------------------------------
 * Stage 2 -- rollout bookkeeping + reverse segmented scan
 * replaces a2c/runner.py:222-245 and a2c/utils.py:63-79.
 * ---------------------------------------------------------------------- */

/* rewards/dones f32 [R*T] are updated IN PLACE to their post-rollout form
 * (bootstrap folded, dones[end]=1); values f32 [R, T+1]; outputs deltas, advs
 * (factor gamma*lambda) and returns (factor gamma), f32 [R*T]; stats f64 [4] gets
 * sum(adv), sum(adv^2), sum(ret), sum(ret^2) ADDED to it (zero it first).
 * Any output pointer except deltas may be NULL. */
int a2c_delta_scan(float* rewards, float* dones, const float* values, float* deltas,
                   float* advs, float* returns, double* stats, int R, int T, float gamma,
                   float gamma_lambda, void* stream);

/* utils.discount(array, dones, factor) on a flat length-N buffer, bit-exact;
 * stats f64 [2] (may be NULL) gets sum, sum of squares added. */
int a2c_discount(const float* array, const float* dones, float* out, double* stats,
                 long long N, float factor, void* stream);

/* Both scans of Updater.update_model in one launch (a2c/updater.py:70-71: advs =
 * discount(deltas, dones, gamma*lambda); :86-88: returns = discount(rewards, dones,
 * gamma)), bit-exact.  rewards / returns may both be NULL (use_nstep_rets).
 * stats f64 [4] (may be NULL) gets sum(adv), sum(adv^2), sum(ret), sum(ret^2) added. */
int a2c_discount2(const float* deltas, const float* rewards, const float* dones, float* advs,
                  float* returns, double* stats, long long N, float gamma_lambda, float gamma,
                  void* stream);

/* ------------------------------------------------------------------------
 * Stage 3 -- fused A2C loss forward + backward (discrete actions)
 * replaces a2c/updater.py:97-106,119-127 and the autograd of those lines.
 *
 * advs / returns are the raw scans.  norm_advs != 0: advs are normalised with
 * mean / unbiased std derived from adv_stats = {sum, sumsq} over n_global
 * elements (allreduce them first when sharded).  ret_norm (may be NULL): {mean,
 * std} applied to returns (norm_returns).  use_nstep != 0: returns = advs + vals.
 * Every mean divides by n_global so per-rank gradients sum to the global one.
 * out f32 [4]: Loss, Pi_Loss, ValLoss, Entropy (this rank's share), zeroed by
 * the call.  dlogits f32 [N, A], dvals f32 [N].
 * ---------------------------------------------------------------------- */
typedef struct {
  float pi_coef, val_coef, entr_coef;
  int norm_advs, use_nstep;
  long long n_global;
} a2c_loss_cfg;

int a2c_loss_fwd_bwd(const float* logits, const float* vals, const long long* actions,
                     const float* advs, const float* returns, const double* adv_stats,
                     const float* ret_norm, const a2c_loss_cfg* cfg, float* dlogits,
                     float* dvals, float* out, long long N, int A, void* stream);

/* ------------------------------------------------------------------------
 * Stage 4 -- policy forward / backward building blocks (fp32, NCHW)
 * replace the torch.nn calls of a2c/models.py (Conv2d :98,312,685; Linear
 * :40,46-47,247-263,380-394,626-636; LayerNorm :392; GRU.forward :465-476) and
 * their autograd.
 * ---------------------------------------------------------------------- */

/* C[M,N] (+)= A[M,K] * B[K,N] with element strides: A(m,k) at A[m*sam + k*sak],
 * B(k,n) at B[k*sbk + n*sbn], C row-major ldc.  bias (may be NULL) is added per
 * column n; relu != 0 applies max(0, .); beta is 0 or 1.  `workspace` (may be
 * NULL) of `workspace_bytes` enables split-K for skinny outputs.
 * precision: 0 = fp32 FFMA; 1 = 3xTF32 on tcgen05 when the shape qualifies
 * (falls back to 0 otherwise); 2 = single-pass TF32 on tcgen05 (tolerance 2e-3). */
int a2c_gemm(const float* A, long long sam, long long sak, const float* B, long long sbk,
             long long sbn, float* C, long long ldc, const float* bias, int M, int N, int K,
             int relu, int beta, int precision, void* workspace, long long workspace_bytes,
             void* stream);

/* a2c_gemm whose result is gated by a ReLU mask laid out like C: C = Cmask > 0 ? value : 0.
 * Used for the input gradient of a Linear layer that reads a ReLU output (ConvModel /
 * GRUModel `resize_emb` on the conv features, a2c/models.py:247,626): the mask of the last conv
 * layer is applied where its output gradient is produced. */
int a2c_gemm_masked(const float* A, long long sam, long long sak, const float* B, long long sbk,
                    long long sbn, float* C, long long ldc, const float* bias, const float* Cmask,
                    int M, int N, int K, int relu, int beta, int precision, void* workspace,
                    long long workspace_bytes, void* stream);

/* First-layer linear on a state view: Y[n, :] = relu?(flatten(state_n) * W^T + b).
 * W f32 [Nout, fs*6400].  `workspace` (may be NULL) enables split-K when the output has
 * fewer tiles than the GPU has SMs (C1: 60 x 256). */
int a2c_linear_state_fwd(const a2c_state_view* x, const float* W, const float* bias, float* Y,
                         int R, int Nout, int relu, void* workspace, long long workspace_bytes,
                         void* stream);
/* dW[Nout, fs*6400] = dY^T * flatten(states). */
int a2c_linear_state_wgrad(const a2c_state_view* x, const float* dY, float* dW, int R, int Nout,
                           void* workspace, long long workspace_bytes, void* stream);

/* out[n] = sum_m X[m, n] (bias gradient); X row-major [M, N]. */
int a2c_colsum(const float* X, float* out, long long M, int N, void* stream);
/* dX = dY * (Y > 0), elementwise, n elements (in place allowed). */
int a2c_relu_bwd(const float* dY, const float* Y, float* dX, long long n, void* stream);

typedef struct {
  int Cin, Cout, H, W, OH, OW, K, stride, pad;
} a2c_conv_shape;

/* Y = relu?(conv2d(X, Wt) + bias).  X f32 [N, Cin, H, W], or a state view when
 * xview != NULL (then X is ignored and N = R*view.T).  Wt f32 [Cout, Cin, K, K].
 * precision (all three conv entry points): 0 = fp32 FFMA direct convolution;
 * 1 = implicit GEMM on tcgen05 with the 3xTF32 split (fp32 parity); 2 = single-pass
 * TF32.  The tensor-core backward kernels need dY already ReLU-masked (Ymask NULL,
 * see a2c_relu_bwd); with a non-NULL Ymask they run the FFMA kernels. */
int a2c_conv2d_fwd(const float* X, const a2c_state_view* xview, const float* Wt,
                   const float* bias, float* Y, int N, const a2c_conv_shape* s, int relu,
                   int precision, void* stream);
/* dX = conv2d_transpose(dY * (Ymask > 0)).  Ymask (may be NULL) is the forward
 * output whose ReLU gates dY. */
int a2c_conv2d_bwd_data(const float* dY, const float* Ymask, const float* Wt, float* dX, int N,
                        const a2c_conv_shape* s, int precision, void* stream);
/* "Mask at the source" form used by a stack of Conv+ReLU layers (a2c/models.py:312,685): X is
 * this layer's input, i.e. the previous layer's ReLU output.  dX = conv2d_transpose(dY * (Ymask >
 * 0)) * (X > 0) is the previous layer's MASKED output gradient and dbias_prev f32 [Cin] (may be
 * NULL) its bias gradient (per-channel sums of dX), both produced in the epilogue of the
 * input-gradient kernel: the previous layer's backward kernels then need no mask and no
 * separate bias-sum pass over dY.  Ymask may be NULL (dY already masked).
 * workspace: >= 8*Cin bytes when dbias_prev. */
int a2c_conv2d_bwd_data_relu(const float* dY, const float* Ymask, const float* Wt, const float* X,
                             float* dX, float* dbias_prev, int N, const a2c_conv_shape* s,
                             int precision, void* workspace, long long workspace_bytes,
                             void* stream);
/* 1-bit ReLU masks (a2c/models.py:312,685: every Conv2d of the stacks is followed by nn.ReLU).
 * relu_bits u8 [N, Cout/8, OH*OW]: bit (c & 7) of byte [n, c >> 3, pos] = Y[n, c, pos] > 0; Cout % 8 == 0.
 * a2c_conv2d_fwd_relu_bits = a2c_conv2d_fwd with relu = 1 that also writes the bits (in the epilogue
 * of the tcgen05 shift-GEMM; any other path runs the plain forward and one pass over Y).
 * a2c_conv2d_bwd_data_relu_bits = a2c_conv2d_bwd_data_relu whose epilogue reads Xbits (the bits of the
 * layer that produced X) instead of the fp32 X: one byte per 8 channels and position instead of eight
 * strided floats.  X is still passed: paths without the fused epilogue mask with it. */
int a2c_conv2d_fwd_relu_bits(const float* X, const a2c_state_view* xview, const float* Wt,
                             const float* bias, float* Y, unsigned char* relu_bits, int N,
                             const a2c_conv_shape* s, int precision, void* stream);
int a2c_conv2d_bwd_data_relu_bits(const float* dY, const float* Ymask, const float* Wt, const float* X,
                                  const unsigned char* Xbits, float* dX, float* dbias_prev, int N,
                                  const a2c_conv_shape* s, int precision, void* workspace,
                                  long long workspace_bytes, void* stream);
/* out f32 [C] = per-channel sums of X f32 [N, C, HW] where mask > 0 (mask may be NULL): the bias
 * gradient of the last conv layer from its masked output gradient.  workspace >= 8*C bytes. */
int a2c_channel_sum(const float* X, const float* mask, float* out, int N, int C, int HW,
                    void* workspace, long long workspace_bytes, void* stream);
/* Kernel family used by the last a2c_conv2d_* call of the calling host thread: 0 = tcgen05
 * shift-GEMM, 1 = tcgen05 implicit GEMM over im2col tiles, 2 = fp32 FFMA direct convolution
 * (-1: none yet).  Lets a host warn when a precision >= 1 request lands on the FFMA kernels. */
int a2c_conv2d_last_path(void);
/* Which backward entries apply the ReLU mask (Ymask > 0) themselves on the path they take for
 * this shape / precision: bit 0 = a2c_conv2d_bwd_data, bit 1 = a2c_conv2d_bwd_weight; bit 2 =
 * a2c_conv2d_bwd_weight also yields dbias at no extra cost on that path (otherwise a non-NULL
 * dbias costs a separate pass over dY).  A caller
 * holding an unmasked dY passes Ymask to those and pre-masks dY (a2c_relu_bwd) for the others;
 * passing Ymask to an entry whose bit is clear is still correct but runs the fp32 FFMA kernel. */
int a2c_conv2d_fused_mask(const a2c_conv_shape* s, int precision, int input_is_view);
/* dW, dbias from X (or xview) and dY * (Ymask > 0).  workspace >=
 * a2c_conv2d_bwd_weight_workspace(s) bytes. */
long long a2c_conv2d_bwd_weight_workspace(const a2c_conv_shape* s);
int a2c_conv2d_bwd_weight(const float* X, const a2c_state_view* xview, const float* dY,
                          const float* Ymask, float* dW, float* dbias, int N,
                          const a2c_conv_shape* s, void* workspace, long long workspace_bytes,
                          int precision, void* stream);

/* Explicit patch matrices for convolutions without a shift-GEMM form (A3CModel's 8x8 s4 and
 * 4x4 s2 layers, a2c/models.py:22-38): col f32 [N*OH*OW, Cin*K*K] is written once and the layer's
 * forward / weight gradient / input gradient run as a2c_gemm on it.  Exactly one source:
 *   xview   planar states (uint8 ring or f32), column order k = (ci*K + ky)*K + kx  (the
 *           reference's W.view(Cout, -1) as is); needs K % 4 == 0;
 *   x_nhwc  channels-last f32 [N, H, W, Cin] (the GEMM output of the previous such layer), column
 *           order k = (ky*K + kx)*Cin + ci; needs Cin % 4 == 0.
 * a2c_col2im_nhwc: dX [N, H, W, Cin] = gather of dcol (channels-fastest order), gated by
 * xmask > 0 (channels-last like dX; may be NULL). */
int a2c_im2col(const a2c_state_view* xview, const float* x_nhwc, float* col, int N,
               const a2c_conv_shape* s, void* stream);
int a2c_col2im_nhwc(const float* dcol, const float* xmask, float* dx, int N,
                    const a2c_conv_shape* s, void* stream);

/* LayerNorm over the last dim (eps 1e-5), rows x H.  Saves mean / rstd f32 [rows]. */
int a2c_layernorm_fwd(const float* X, const float* gamma, const float* beta, float* Y,
                      float* mean, float* rstd, long long rows, int Hdim, void* stream);
/* dX, and dgamma/dbeta partials reduced into dgamma, dbeta f32 [H] (overwritten). */
int a2c_layernorm_bwd(const float* dY, const float* X, const float* gamma, const float* mean,
                      const float* rstd, float* dX, float* dgamma, float* dbeta, long long rows,
                      int Hdim, void* stream);

/* GRU cell of a2c/models.py:465-476, gate math only (the six products come from
 * a2c_gemm):  gx f32 [B, 3H] = x*[Wx0|Wx1|Wx2] + b,  gh f32 [B, 2H] = h*[Wh0|Wh1].
 * step 1: z, r = sigmoid(gx[:, :2H] + gh);  rh = r*h.
 * step 2 (after gc = rh*Wh2):  cand = tanh(gx[:, 2H:] + gc);  hnew = z*h + (1-z)*cand,
 *         hnext = hnew * keep  (keep f32 [B] = 1-done; NULL = no masking, a2c/updater.py:164). */
int a2c_gru_gates1(const float* gx, const float* gh, const float* h, float* z, float* r,
                   float* rh, long long B, int Hdim, void* stream);
int a2c_gru_gates2(const float* gx, const float* gc, const float* h, const float* z,
                   const float* keep, float* cand, float* hnew, float* hnext, long long B,
                   int Hdim, void* stream);
/* Backward of step 2.  g = dhnew + dhnext * keep (either gradient may be NULL;
 * keep NULL = 1).  Produces dgc (grad of the candidate pre-activation = dgx[:, 2H:]),
 * dz_out (grad wrt z, sigmoid derivative NOT applied) and the direct part of dh. */
int a2c_gru_gates2_bwd(const float* dhnew, const float* dhnext, const float* keep,
                       const float* h, const float* z, const float* cand, float* dgc,
                       float* dz_out, float* dh, long long B, int Hdim, void* stream);
/* Backward of step 1: given drh (= dgc * Wh2^T) and dz_out, produces the z/r
 * pre-activation grads dgzr f32 [B, 2H] (= dgx[:, :2H] = dgh) and adds r*drh to dh. */
int a2c_gru_gates1_bwd(const float* drh, const float* dz_out, const float* h, const float* z,
                       const float* r, float* dgzr, float* dh, long long B, int Hdim,
                       void* stream);

/* ------------------------------------------------------------------------
 * Clip + optimizer over one flat fp32 parameter buffer
 * replaces a2c/updater.py:129-132 (clip_grad_norm_, optim.step, zero_grad).
 * ---------------------------------------------------------------------- */
/* sumsq f64 [1] += sum(g^2) (zero it first). */
int a2c_grad_sumsq(const float* grad, double* sumsq, long long n, void* stream);
/* clip coef = min(1, max_norm / (sqrt(sumsq) + 1e-6)), read on device.
 * kind 0: RMSprop (alpha, eps; state1 = square_avg).  kind 1: Adam (beta1, beta2,
 * eps, step; state1 = exp_avg, state2 = exp_avg_sq).  zero_grad != 0 clears grad.
 * norm_out f32 [1] (may be NULL) receives the pre-clip norm. */
typedef struct {
  int kind, zero_grad;
  float lr, max_norm, alpha, eps, beta1, beta2;
  long long step;
} a2c_optim_cfg;
int a2c_clip_optim_step(float* param, float* grad, float* state1, float* state2,
                        const double* sumsq, const a2c_optim_cfg* cfg, float* norm_out,
                        long long n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* A2C_B200_H */
