/* hulc_b200.h — C ABI of libhulc_b200.so: the sm_100a kernels behind the HULC training hot path.
 *
 * The reference (lukashermann/hulc) is pure Python: its "FFI" for this path is torch's operator set called from the
 * nn.Module.forward / loss methods cited next to each entry point below (paths relative to the reference root).  A
 * maintainer binds these symbols with ctypes (see INTEGRATION.md and hulc_b200/_lib.py) and calls them from
 * torch.autograd.Function bodies in place of those torch ops.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to fp32 unless the parameter type says otherwise; row-major, leading dimensions
 *     in elements; no torch types cross this boundary;
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on that stream and never synchronises;
 *   - return value: 0 on success, otherwise the cudaError_t of the failing launch / argument check;
 *   - dropout arguments: `drop_p` = 0 disables; with `drop_keep` != NULL it is an injected uint8 keep-mask laid out like
 *     the tensor it applies to, otherwise keep decisions are Philox4x32-10(seed, site, element index) and the backward
 *     entry points regenerate them from the same (seed, site);
 *   - `workspace`: caller-owned scratch (16-byte aligned, zero-initialised once, then reusable by calls on the SAME stream).
 */
#ifndef HULC_B200_H
#define HULC_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- dense layers ---------------------------------------------------------------------------------------------------
 * hulc_gemm: C[M,N] = epi(alpha * op(A)[M,K] * op(B)[K,N]); transA=1: A stored KxM; transB=1: B stored NxK (torch Linear
 * weight).  epi(v) = dropout(gate(act(v + bias[n] + addend[(add_mod ? m % add_mod : m), n] + beta*C[m,n]))).
 * Replaces torch.nn.Linear forward / backward everywhere on the path (e.g. plan_encoders/plan_proposal_net.py:26-47,
 * encoders/goal_encoders.py:20-36, decoders/logistic_decoder_rnn.py:278-283) and the per-step matmuls of torch.nn.RNN
 * (decoders/utils/rnn.py:5-14). */
int hulc_gemm(const float* A, const float* B, float* C, int M, int N, int K, int lda, int ldb, int ldc, int transA, int transB,
              float alpha, float beta, const float* bias, const float* addend, int ldadd, int add_mod, int act,
              const float* gate, int ldg, float drop_p, unsigned long long drop_seed, unsigned drop_site,
              const unsigned char* drop_keep, float* workspace, size_t workspace_bytes, void* stream);

/* hulc_gemm_tc: the same contract on the tensor cores (tcgen05.mma kind::tf32, fp32 accumulation in TMEM).  passes = 1:
 * operands consumed as tf32; passes = 3: 3xTF32 split products (fp32-level accuracy; the residuals x - tf32(x) are formed
 * in shared memory).  Skinny products are split along K (partials and tickets in the workspace, fixed-order reduction by
 * the last CTA of each tile).  Operands must be 16-byte aligned with lda, ldb and their contiguous extents multiples of 4,
 * else cudaErrorInvalidValue. */
int hulc_gemm_tc(const float* A, const float* B, float* C, int M, int N, int K, int lda, int ldb, int ldc, int transA, int transB,
                 float alpha, float beta, const float* bias, const float* addend, int ldadd, int add_mod, int act,
                 const float* gate, int ldg, float drop_p, unsigned long long drop_seed, unsigned drop_site,
                 const unsigned char* drop_keep, int passes, float* workspace, size_t workspace_bytes, void* stream);

/* Device-resident RNG offset: when set (non-NULL device pointer), every Philox-drawn quantity (dropout keep decisions, latent
 * plan samples) uses seed + *ptr.  Advancing *ptr on the device between steps gives a captured CUDA graph of the whole
 * training step fresh randomness on every replay. */
int hulc_set_rng_offset_ptr(const unsigned long long* device_ptr);

/* *out (HOST pointer) = number of kernel launches this library has issued since it was loaded (bench.py's gpu_launches). */
int hulc_launch_count(unsigned long long* out);

/* out[c] = beta*out[c] + sum_r X[r*ldx + c]  (bias gradients).  Long reductions are split over CTAs through the workspace
 * and combined in a fixed order (bit-reproducible). */
int hulc_colsum(const float* X, int rows, int cols, int ldx, float* out, float beta, float* workspace, size_t workspace_bytes,
                void* stream);
/* njobs column sums in one launch: table = njobs x 8 int64 on the device — {X, out, rows, cols, ldx, float bits of beta, first block of the
 * job, 0}; job j owns blocks [first_j, first_j + ceil(cols_j / 32)); out_j[c] = beta_j * out_j[c] + sum_r X_j[r][c].  The bias gradients
 * (sum over rows of dY for every nn.Linear / nn.Conv2d of the backward pass) go through this once per step. */
int hulc_colsum_multi(const void* table, int njobs, int total_blocks, void* stream);

/* activation / gate selectors for hulc_gemm's `act` argument */
#define HULC_ACT_NONE 0
#define HULC_ACT_RELU 1
#define HULC_ACT_TANH 2
#define HULC_GATE_TANH 4 /* OR-ed in: `gate` holds tanh outputs, C *= 1 - gate^2 (default: C = gate > 0 ? C : 0) */

/* ---- GRU gate math ---------------------------------------------------------------------------------------------------
 * torch.nn.GRU built at decoders/utils/rnn.py:27-36 (gate order r|z|n).  gi = x W_ih^T + b_ih, gh = h_prev W_hh^T + b_hh
 * (both [B,3H], produced by hulc_gemm); fwd writes h and saved[B,4H] = r|z|n|gh_n; bwd takes dh = dh_above + dh_rec and
 * writes dgi, dgh and dh_carry = dh*z (the caller adds dgh W_hh). */
int hulc_gru_gates_fwd(const float* gi, int ldgi, const float* gh, int ldgh, const float* hprev, int ldhp, float* h, int ldh,
                       float* saved, int B, int H, void* stream);
int hulc_gru_gates_bwd(const float* dh_above, int lda, const float* dh_rec, int ldr, const float* saved, const float* hprev,
                       int ldhp, float* dgi, int ldgi, float* dgh, int ldgh, float* dh_carry, int ldc, int B, int H, void* stream);

/* ---- a whole Elman recurrence in one persistent launch (tensor cores, W_hh resident in shared memory) ----------------------
 * torch.nn.RNN as built at decoders/utils/rnn.py:5-14 (ReLU) and plan_encoders/plan_recognition_net.py:27-34 (tanh,
 * bidirectional), one layer and direction per call, all S dependent steps:
 *   transW = 0 (forward):  out_s = act(add_s + prev_s W^T)            W [H,H] = weight_hh (out x in)
 *   transW = 1 (BPTT):     out_s = (add_s + prev_s W) * act'(gate_s)   gate_s = h_t; act & HULC_GATE_TANH selects 1 - g^2
 * for s = 0..S-1 with X_s = X0 + s * X_step (element strides, may be negative: reverse direction / backward in time);
 * every operand is a [B, H] view with its own leading dimension.  out_s of one step is prev_{s+1} of the next (the caller
 * lays the slots out so).  H must be 2048 and B <= 64, else cudaErrorInvalidValue (callers fall back to per-step
 * hulc_gemm_tc).  Operands are rounded to tf32 (nearest), accumulation and epilogue are fp32. */
int hulc_rnn_tc_seq(const float* W, int ldw, int transW, const float* prev0, long long prev_step, int ldp, float* out0, long long out_step,
                    int ldo, const float* add0, long long add_step, int ldadd, const float* gate0, long long gate_step, int ldg, int act,
                    int B, int H, int S, float* workspace, size_t workspace_bytes, void* stream);

/* bf16 variant (reference under 16-bit autocast, conf/trainer/play_trainer.yaml:3): W16 is the bf16 copy of weight_hh [H, ldw]; the
 * hidden state travels between the steps as bf16 through x16 (workspace, (S + 1) * B * H bf16; slot 0 is filled from the fp32 initial
 * state prev0 [B, ldp]); the fp32 result of step s goes to out0 + s * out_step (may be NULL).  add / gate / act / transW as above. */
int hulc_rnn_seq_bf16(const void* W16, int ldw, int transW, const float* prev0, int ldp, void* x16, float* out0, long long out_step, int ldo,
                      const float* add0, long long add_step, int ldadd, const float* gate0, long long gate_step, int ldg, int act, int B, int H,
                      int S, void* stream);

/* Diagnostics: number of 4-CTA clusters of the second-generation recurrence kernels (csrc/rnn_push_tc.cu) the current device can hold
 * at once (they need 32 co-resident; fewer: hulc_rnn_seq_bf16 returns cudaErrorLaunchOutOfResources, hulc_rnn_tc_seq runs its first kernel).
 * out[0]: bf16 kernel, out[1]: tf32 kernel. */
int hulc_rnn_push_max_clusters(int* out);

/* ---- convolutions of the perceptual encoders -------------------------------------------------------------------------
 * perceptual_encoders/vision_network.py:36-47 and vision_network_gripper.py:11-17: nn.Conv2d (valid, NCHW) + ReLU for
 * the three layer shapes (3->32 k8 s4, 32->64 k4 s2, 64->64 k3 s1).  x [N,CIN,H,W], w [COUT,CIN,KS,KS], y [N,COUT,HO,WO].
 * dgrad gates dx by (gate > 0) when gate != NULL (the ReLU output that fed this conv); wgrad: dw = beta*dw + grad. */
int hulc_conv2d_fwd(const float* x, const float* w, const float* b, float* y, int N, int CIN, int H, int W, int COUT, int KS, int S,
                    int relu, float* workspace, size_t workspace_bytes, void* stream);
int hulc_conv2d_dgrad(const float* dy, const float* w, const float* gate, float* dx, int N, int CIN, int H, int W, int COUT, int KS,
                      int S, float* workspace, size_t workspace_bytes, void* stream);
int hulc_conv2d_wgrad(const float* x, const float* dy, float* dw, float beta, int N, int CIN, int H, int W, int COUT, int KS, int S,
                      float* workspace, size_t workspace_bytes, void* stream);
/* The same three layers on the tensor cores (tcgen05.mma kind::tf32) with channels-last activations: x NHWC [N,H,W,CIN] (or,
 * for the 3-channel first layer, the reference's NCHW frames: fwd reads them as they are, wgrad with x_nchw = 1);
 * y / dy NHWC [N,HO,WO,COUT]; gate / dx NHWC; w / dw keep the reference layout [COUT,CIN,KS,KS].  The first layer reads the frames in
 * 16-byte pieces: W must be a multiple of 4 and x 16-byte aligned (cudaErrorInvalidValue otherwise; the reference's cameras are 200 and 84
 * pixels wide — other widths take the exact-fp32 hulc_conv2d_* kernels). */
/* relu_bits (optional, fwd): the sign mask of y, bit c % 32 of word [pixel][c / 32] = (y > 0), written next to y; gate_bits (optional,
 * dgrad): the same mask of the gating activation — the data gradient then reads 4 bytes instead of 128 per (pixel, 32 channels); `gate`
 * must still be given (kernels that do not take the mask use it). */
int hulc_conv2d_tc_fwd(const float* x, const float* w, const float* b, float* y, int N, int CIN, int H, int W, int COUT, int KS, int S,
                       int relu, unsigned* relu_bits, float* workspace, size_t workspace_bytes, void* stream);
int hulc_conv2d_tc_dgrad(const float* dy, const float* w, const float* gate, const unsigned* gate_bits, float* dx, int N, int CIN, int H, int W,
                         int COUT, int KS, int S, float* workspace, size_t workspace_bytes, void* stream);
/* db (optional): the bias gradient sum_pixels dy[.., co] is ADDED to db[co] — by a row of ones appended to the im2col matrix of the same
 * tensor-core pass where the kernel has a spare GEMM row (first layer), otherwise by a column-sum pass over dy. */
int hulc_conv2d_tc_wgrad(const float* x, const float* dy, float* dw, float beta, float* db, int N, int CIN, int H, int W, int COUT, int KS, int S,
                         int x_nchw, float* workspace, size_t workspace_bytes, void* stream);
/* out[c] += sum_{n,p} x[n,c,p] (conv bias gradient; accumulates like the other parameter-gradient outputs) */
int hulc_nchw_channel_sum(const float* x, float* out, int N, int C, int P, void* stream);

/* ---- SpatialSoftmax (vision_network.py:74-108) ------------------------------------------------------------------------
 * rows = N*C feature maps of H*W; out[row] = (E[x_map], E[y_map]) interleaved, i.e. the (N, 2C) tensor of the reference.
 * bwd with relu_gate=1 also applies the mask of the ReLU that produced x (x > 0). */
int hulc_spatial_softmax_fwd(const float* x, float* out, int rows, int H, int W, float inv_temp, void* stream);
int hulc_spatial_softmax_bwd(const float* x, const float* dout, float* dx, int rows, int H, int W, float inv_temp, int relu_gate,
                             void* stream);

/* the same on channels-last maps x [N, H*W, C] (the layout of the tensor-core convolutions); out [N, 2C] as above */
int hulc_spatial_softmax_nhwc_fwd(const float* x, float* out, int N, int C, int H, int W, float inv_temp, void* stream);
int hulc_spatial_softmax_nhwc_bwd(const float* x, const float* dout, float* dx, int N, int C, int H, int W, float inv_temp, int relu_gate,
                                  void* stream);

/* ---- LayerNorm (+ residual + dropout) ----------------------------------------------------------------------------------
 * nn.LayerNorm at vision_network.py:53, goal_encoders.py:29, and the post-norm residual blocks of
 * nn.TransformerEncoderLayer (plan_recognition_net.py:83-85): z = res + dropout(x) (z = x when res == NULL),
 * y = (z - mean) * rstd * w + b; stats[row] = (mean, rstd).  bwd: dz (to the residual), dx = dz*dropout, dw/db ACCUMULATE. */
int hulc_layernorm_fwd(const float* x, int ldx, const float* res, int ldres, const float* w, const float* b, float* y, int ldy,
                       float* z, int ldz, float* stats, int rows, int D, float eps, float drop_p, unsigned long long drop_seed,
                       unsigned drop_site, const unsigned char* drop_keep, void* stream);
int hulc_layernorm_bwd(const float* dy, int lddy, const float* z, int ldz, const float* stats, const float* w, float* dz, int lddz,
                       float* dx, int lddx, float* dw, float* db, int rows, int D, float drop_p, unsigned long long drop_seed,
                       unsigned drop_site, const unsigned char* drop_keep, void* stream);

/* ---- posterior transformer pieces (plan_encoders/plan_recognition_net.py:94-117) ----------------------------------------
 * add_posemb: y[b,s,:] = dropout(x[b,s,:] + pos[s,:]) (:101-111).  attention: packed qkv [B*S, 3*H*dh] batch-first,
 * softmax(q k^T / sqrt(dh)) with dropout on the probabilities, out [B*S, H*dh]; probs [B,H,S,S] saved for bwd. */
int hulc_add_posemb_fwd(const float* x, const float* pos, float* y, int B, int S, int D, float drop_p, unsigned long long drop_seed,
                        unsigned drop_site, const unsigned char* drop_keep, void* stream);
int hulc_dropout_apply(const float* x, float* y, long long n, float drop_p, unsigned long long drop_seed, unsigned drop_site,
                       const unsigned char* drop_keep, void* stream);
int hulc_attention_fwd(const float* qkv, float* out, float* probs, int B, int S, int H, int dh, float drop_p,
                       unsigned long long drop_seed, unsigned drop_site, const unsigned char* drop_keep, void* stream);
int hulc_attention_bwd(const float* qkv, const float* probs, const float* dout, float* dqkv, int B, int S, int H, int dh,
                       float drop_p, unsigned long long drop_seed, unsigned drop_site, const unsigned char* drop_keep, void* stream);

/* ---- data movement helpers -------------------------------------------------------------------------------------------------
 * strided_copy: dst[i0,i1,i2] (+)= alpha*src[i0,i1,i2] with element strides (torch slicing / permute / expand at
 * logistic_decoder_rnn.py:267-272, concat_encoders.py:103-107); reduce_mid: out[b,d] = scale*sum_s x[b,s,d] (the mean over
 * time at plan_recognition_net.py:114); sum: out[0] = scale*sum x; scale: x *= alpha. */
int hulc_strided_copy(float* dst, const float* src, int n0, int n1, int n2, long long d0, long long d1, long long d2, long long s0,
                      long long s1, long long s2, float alpha, int accumulate, void* stream);
int hulc_reduce_mid(const float* x, float* out, int B, int S, int D, float scale, void* stream);
int hulc_sum(const float* x, int n, float* out, float scale, void* stream);
int hulc_scale(float* x, long long n, float alpha, void* stream);
/* x *= *alpha_ptr with the factor in device memory (the gradient autograd hands to `loss.backward()`, hulc/training.py ->
 * Lightning's backward of the tensor returned by Hulc.training_step, hulc.py:537); a factor of exactly 1 leaves x untouched. */
int hulc_scale_dev(float* x, long long n, const float* alpha_ptr, void* stream);

/* ---- uint8 camera frames -> normalised fp32 (the input pipeline on the device) ---------------------------
 * dst[i] = ((src[i] / 255) - mean) / std in fp32, the deterministic part of the reference's image transforms (ScaleImageTensor +
 * Normalize(0.5, 0.5), conf/datamodule/transforms/rand_shift.yaml:3-22; hulc/utils/transforms.py:8-29) — lets the host hand over the
 * uint8 frames it read from disk (4x fewer PCIe bytes).  The random-shift augmentation stays with the data pipeline. */
int hulc_frames_u8_to_f32(const unsigned char* src, float* dst, long long n, float mean, float stdv, void* stream);
/* The training-time image pipeline of the reference on the device (conf/datamodule/transforms/rand_shift.yaml:2-22): RandomShiftsAug
 * (hulc/utils/transforms.py:8-29: replicate-pad by `pad`, shift by (sx, sy) in [0, 2 pad] whole pixels per frame) fused with the scale +
 * normalise of hulc_frames_u8_to_f32.  src uint8 [N][C][H][W] -> dst fp32 [N][C][H][W]; shifts [N][2] (sx, sy) int32 on the device, or NULL to
 * draw them from the Philox stream (seed, site) — one pair per frame, like torch.randint in the reference. */
int hulc_frames_u8_shift_to_f32(const unsigned char* src, float* dst, int N, int C, int H, int W, int pad, const int* shifts, unsigned long long seed,
                                unsigned site, float mean, float stdv, void* stream);

/* ---- world_to_tcp_frame (decoders/utils/gripper_control.py:16-36) ----------------------------------------------------------
 * actions [n,7], robot_obs [n,obs_dim] (euler XYZ at 3:6) -> out [n,7].  *nan_flag is OR-ed with 1 if any output is NaN
 * (the reference asserts on the host, :35; the flag is checked without stalling the stream). */
int hulc_world_to_tcp(const float* actions, const float* robot_obs, int obs_dim, float* out, int n_tokens, int* nan_flag, void* stream);

/* ---- validation path: tcp_to_world_frame (gripper_control.py:39-63), sampling from the logistic mixture
 * (logistic_decoder_rnn.py:234-258) and the MAE / gripper-accuracy reductions of lmp_val (hulc/models/hulc.py:346-385) -------
 * hulc_tcp_to_world: sampled actions [n,7] in the TCP frame + robot_obs -> world frame [n,7]; *nan_flag as in hulc_world_to_tcp.
 * hulc_logistic_sample: heads rows as in hulc_logistic_loss (row = b*S+t or t*B+b); for the sequences [b0,b0+Bm) writes
 *   out[b][t][0..n_dims) = mean_k + exp(max(log_scale_k, log_scale_min)) * (log u - log(1-u)), k = argmax_k(logit_k - log(-log u_k)),
 *   and, with has_gripper, out[b][t][n_dims] = gripper logit 1 > logit 0 ? grip_hi : grip_lo.  u_mix [Bm*S][n_dims][n_mix] and
 *   u_inv [Bm*S][n_dims] are U[0,1) draws (mapped to [1e-5, 1-1e-5] like the reference); NULL -> Philox(seed, site / site+1).
 * hulc_val_metrics: pred / actions [B,S,n_dims+1] -> mae [B,n_dims] (mean over S of |pred - actions|), hits [B] (steps whose
 *   sign(pred gripper) equals the ground-truth gripper command). */
int hulc_tcp_to_world(const float* actions, const float* robot_obs, int obs_dim, float* out, int n_tokens, int* nan_flag, void* stream);
int hulc_logistic_sample(const float* heads, int ldh, const float* u_mix, const float* u_inv, float* out, int B, int S, int b0, int Bm,
                         int time_major, int n_dims, int n_mix, float log_scale_min, int has_gripper, float grip_lo, float grip_hi,
                         unsigned long long seed, unsigned site, void* stream);
int hulc_val_metrics(const float* pred, const float* actions, float* mae, float* hits, int B, int S, int n_dims, void* stream);

/* ---- discretised logistic mixture NLL + gripper CE (decoders/logistic_decoder_rnn.py:136-155,184-231) ---------------------
 * heads rows: [logit_probs n_dims*n_mix | means | raw log_scales | gripper 2]; B sequences of S tokens, row = b*S+t or
 * t*B+b (time_major); the loss covers sequences [b0,b0+Bm).  losses[0] = NLL, losses[1] = CE (means over Bm*S tokens);
 * dheads = grad_scale * d(losses[0] + gripper_alpha*losses[1])/d heads, written in the same pass. */
int hulc_logistic_loss(const float* heads, int ldh, const float* actions, int act_dim, float* dheads, float* losses, int B, int S,
                       int b0, int Bm, int time_major, int n_dims, int n_mix, int num_classes, float log_scale_min, float act_min,
                       float act_max, int has_gripper, float gripper_alpha, float grad_scale, float* workspace,
                       size_t workspace_bytes, void* stream);

/* ---- latent plan: sample + KL (hulc/utils/distributions.py:23-60, hulc/models/hulc.py:289-291,539-561) ---------------------
 * discrete: rows = batch*category_size rows of class_size (=32) logits; the sample index is idx_in[row] if given, else the
 * inverse CDF of u[row], else of Philox(seed, site, row); plan = one-hot (straight-through: the bwd adds the softmax
 * Jacobian); kl_rows[row] = KL(post || prior) of that category.  bwd: d_pr = ST(dplan) + dkl*coef_rhs*dKL/dpost,
 * d_pp = dkl*coef_lhs*dKL/dprior (dkl == NULL -> 1).  continuous: state = [mean | raw_std], plan = mean + std*eps. */
int hulc_plan_discrete_fwd(const float* pr_logit, const float* pp_logit, const float* u, const int* idx_in, float* plan, int* idx_out,
                           float* kl_rows, int rows, int class_size, unsigned long long seed, unsigned site, void* stream);
int hulc_plan_discrete_bwd(const float* pr_logit, const float* pp_logit, const float* dplan, const float* dkl, float coef_lhs,
                           float coef_rhs, float* d_pr, float* d_pp, int rows, int class_size, void* stream);
int hulc_plan_cont_fwd(const float* pr_state, const float* pp_state, const float* eps, float* plan, float* kl_elem, int batch,
                       int plan_features, unsigned long long seed, unsigned site, void* stream);
int hulc_plan_cont_bwd(const float* pr_state, const float* pp_state, const float* eps, const float* dplan, const float* dkl,
                       float coef_lhs, float coef_rhs, float* d_pr, float* d_pp, int batch, int plan_features,
                       unsigned long long seed, unsigned site, void* stream);

/* ---- CLIP-style auxiliary loss (hulc/models/hulc.py:650-695) on the ProjVisLang outputs --------------------------------------
 * im, tx [n,D]; rows with mask[i]==0 are left out (all masked -> loss 0, zero grads, :669-676).  Writes the loss and
 * grad_scale * its gradients w.r.t. im, tx and logit_scale. */
int hulc_clip_loss(const float* im, const float* tx, const float* logit_scale, const unsigned char* mask, float* loss, float* d_im,
                   float* d_tx, float* d_logit_scale, int n, int D, float grad_scale, void* stream);

/* ---- auxiliary losses of the ablation configs ------------------------------------------------------------------------------------
 * hulc_cosine_loss: Hulc.bc_z_auxiliary_loss (hulc/models/hulc.py:567-604) on the BCZLangDecoder output: loss[0] = mean_b (1 - cos(pred_b, target_b)),
 *   dpred = grad_scale * d loss / d pred.  pred, target [B, D] with leading dimensions; B <= 256.
 * hulc_bce_logits_loss: the binary cross entropy with logits of Hulc.mia_auxiliary_loss (hulc.py:606-648) over n_pos matching scores (label 1)
 *   followed by n_neg rolled-pair scores (label 0): loss[0] = mean, dlogits = grad_scale * d loss / d logits. */
int hulc_cosine_loss(const float* pred, int ldp, const float* target, int ldt, float* dpred, int ldd, float* loss, int B, int D, float grad_scale,
                     void* stream);
int hulc_bce_logits_loss(const float* logits, float* dlogits, float* loss, int n_pos, int n_neg, float grad_scale, void* stream);

/* ---- bf16 path (BASELINE config 3: the reference trains under 16-bit autocast, conf/trainer/play_trainer.yaml:3) ------------------
 * hulc_cast_bf16(_rows): y = bf16(x), round to nearest even — what torch.autocast does to the inputs of every nn.Linear / nn.Conv2d.
 * hulc_gemm_bf16: the nn.Linear products (forward x W^T, data gradient dY W, weight gradient dY^T x — the same call sites as hulc_gemm)
 * with bf16 operands A, B on the tensor cores (tcgen05.mma kind::f16, fp32 accumulation, operands staged by TMA) and hulc_gemm's fused
 * epilogue in fp32.  The result goes to C (fp32; beta*C is added first when beta != 0) and / or Cb (bf16, the operand of the next
 * product); the gate is read from `gate` (fp32) or `gate_bf16`.  Requirements: 16-byte aligned A / B, lda % 8 == ldb % 8 == 0. */
int hulc_cast_bf16(const float* x, void* y, long long n, void* stream);
int hulc_cast_bf16_rows(const float* x, int ldx, void* y, int ldy, int rows, int cols, void* stream);
int hulc_gemm_bf16(const void* A, const void* B, float* C, void* Cb, int M, int N, int K, int lda, int ldb, int ldc, int ldcb, int transA, int transB,
                   float alpha, float beta, const float* bias, const float* addend, int ldadd, int add_mod, int act, const float* gate, const void* gate_bf16,
                   int ldg, float drop_p, unsigned long long drop_seed, unsigned drop_site, const unsigned char* drop_keep, void* stream);

/* bf16 activations between the conv layers (vision_network.py:36-47 under autocast): the same three layers as hulc_conv2d_tc_*, with
 * y / dy / dx (and x of layers 2, 3) channels-last bf16; layer 1 still reads the reference's fp32 NCHW frames.  Weights, bias and both
 * gradients dw / db stay fp32 (reference layout); relu_bits / gate_bits as above (the data gradient gates on the sign bits only).
 * wgrad: dw = beta * dw + dL/dw; db (optional) += column sums of dy, out of the same tensor-core pass. */
int hulc_conv2d_bf16_fwd(const void* x, const float* w, const float* b, void* y, int N, int CIN, int H, int W, int COUT, int KS, int S, int relu,
                         unsigned* relu_bits, float* workspace, size_t workspace_bytes, void* stream);
int hulc_conv2d_bf16_dgrad(const void* dy, const float* w, const unsigned* gate_bits, void* dx, int N, int CIN, int H, int W, int COUT, int KS, int S,
                           float* workspace, size_t workspace_bytes, void* stream);
int hulc_conv2d_bf16_wgrad(const void* x, const void* dy, float* dw, float beta, float* db, int N, int CIN, int H, int W, int COUT, int KS, int S,
                           float* workspace, size_t workspace_bytes, void* stream);

/* SpatialSoftmax (vision_network.py:100-108) on a bf16 channels-last map; dx (bf16) is gated by x > 0 when relu_gate (the conv's ReLU). */
int hulc_spatial_softmax_nhwc_bf16_fwd(const void* x, float* out, int N, int C, int H, int W, float inv_temp, void* stream);
int hulc_spatial_softmax_nhwc_bf16_bwd(const void* x, const float* dout, void* dx, int N, int C, int H, int W, float inv_temp, int relu_gate, void* stream);

/* ---- optimizer: torch.optim.Adam(lr, betas, eps), no weight decay (hulc/models/hulc.py:239-252) over a flat buffer ---------
 * g is multiplied by grad_scale first (1/world after the all-reduce); `step` is the 1-based step count, read from the device
 * integer *step_ptr instead when step_ptr != NULL (so the launch can be replayed from a CUDA graph). */
int hulc_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps, int step,
                   const int* step_ptr, float grad_scale, void* stream);
/* same update; additionally writes the new parameters as bf16 into p_bf16 (n elements) — the tensor-core operand copy of the bf16
 * path, refreshed in the same pass (fp32 master weights stay the source of truth, as under torch.autocast). */
int hulc_adam_step_bf16(float* p, const float* g, float* m, float* v, void* p_bf16, long long n, float lr, float beta1, float beta2, float eps, int step,
                        const int* step_ptr, float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HULC_B200_H */
