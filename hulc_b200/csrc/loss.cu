// loss.cu — the scalar-producing ends of the step, each as ONE fused forward+backward kernel:
//   world_to_tcp_frame, discretised-logistic-mixture NLL + gripper cross-entropy, latent-plan sampling + balanced KL
//   (discrete 32x32 categorical and continuous Gaussian), CLIP-style contrastive loss.
// The reference runs each of these as dozens of elementwise aten kernels plus host syncs (SURVEY.md §2.3 K13,K17-K19).
#include "common.cuh"

namespace {

__device__ __forceinline__ float softplusf(float x) { return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x))); }
__device__ __forceinline__ float sigmoidf(float x) { return 1.f / (1.f + expf(-x)); }

// ---------------------------------------------------------------------------------------------------------------------
// world_to_tcp_frame (decoders/utils/gripper_control.py:16-36)
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mat3_mul(const float* A, const float* B, float* C) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}
// R = Rx(a) Ry(b) Rz(c)  (pytorch3d_transforms.py:162-218, convention "XYZ")
__device__ __forceinline__ void euler_xyz(float a, float b, float c, float* R) {
  float sa, ca, sb, cb, sc, cc;
  sincosf(a, &sa, &ca); sincosf(b, &sb, &cb); sincosf(c, &sc, &cc);
  float Rx[9] = {1, 0, 0, 0, ca, -sa, 0, sa, ca};
  float Ry[9] = {cb, 0, sb, 0, 1, 0, -sb, 0, cb};
  float Rz[9] = {cc, -sc, 0, sc, cc, 0, 0, 0, 1};
  float T[9];
  mat3_mul(Rx, Ry, T);
  mat3_mul(T, Rz, R);
}
// general 3x3 inverse by cofactors (the reference calls torch.inverse)
__device__ __forceinline__ void mat3_inv(const float* m, float* o) {
  float c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
  float det = m[0] * c00 + m[1] * c01 + m[2] * c02;
  float id = 1.f / det;
  o[0] = c00 * id; o[1] = (m[2] * m[7] - m[1] * m[8]) * id; o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  o[3] = c01 * id; o[4] = (m[0] * m[8] - m[2] * m[6]) * id; o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  o[6] = c02 * id; o[7] = (m[1] * m[6] - m[0] * m[7]) * id; o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}

__global__ void world_to_tcp_kernel(const float* __restrict__ act, const float* __restrict__ robot_obs, int obs_dim, float* __restrict__ out, int n,
                                    int* __restrict__ nan_flag) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* a = act + (size_t)i * 7;
  const float* o = robot_obs + (size_t)i * obs_dim;
  const float PI = 3.14159265358979323846f;
  float R[9], Rinv[9], Rn[9], Rninv[9], M[9];
  euler_xyz(o[3], o[4], o[5], R);
  mat3_inv(R, Rinv);
  float r[7];
#pragma unroll
  for (int k = 0; k < 3; ++k) r[k] = Rinv[k * 3] * a[0] + Rinv[k * 3 + 1] * a[1] + Rinv[k * 3 + 2] * a[2];
  euler_xyz(o[3] + a[3] * 0.01f, o[4] + a[4] * 0.01f, o[5] + a[5] * 0.01f, Rn);
  mat3_inv(Rn, Rninv);
  mat3_mul(Rninv, R, M);
  // matrix_to_euler_angles(M, "XYZ") (pytorch3d_transforms.py:264-303)
  float e[3] = {atan2f(-M[5], M[8]), asinf(M[2]), atan2f(-M[1], M[0])};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float v = e[k];
    if (v < -PI) v += 2.f * PI;
    if (v > PI) v -= 2.f * PI;
    r[3 + k] = v * 100.f;
  }
  r[6] = a[6];
  bool bad = false;
#pragma unroll
  for (int k = 0; k < 7; ++k) { out[(size_t)i * 7 + k] = r[k]; bad |= (r[k] != r[k]); }
  if (bad && nan_flag) atomicOr(nan_flag, 1);
}

// tcp_to_world_frame (decoders/utils/gripper_control.py:39-63): the inverse map, applied to SAMPLED actions on the validation path.
// R = R(euler); pos_w = R pos_tcp; R_new = R * inv(R(0.01 * orn_tcp)); orn_w = 100 * wrap_pi(euler(R_new) - euler); gripper passes through.
__global__ void tcp_to_world_kernel(const float* __restrict__ act, const float* __restrict__ robot_obs, int obs_dim, float* __restrict__ out, int n,
                                    int* __restrict__ nan_flag) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* a = act + (size_t)i * 7;
  const float* o = robot_obs + (size_t)i * obs_dim;
  const float PI = 3.14159265358979323846f;
  float R[9], Rrel[9], Rrelinv[9], M[9];
  euler_xyz(o[3], o[4], o[5], R);
  float r[7];
#pragma unroll
  for (int k = 0; k < 3; ++k) r[k] = R[k * 3] * a[0] + R[k * 3 + 1] * a[1] + R[k * 3 + 2] * a[2];
  euler_xyz(a[3] * 0.01f, a[4] * 0.01f, a[5] * 0.01f, Rrel);
  mat3_inv(Rrel, Rrelinv);
  mat3_mul(R, Rrelinv, M);
  float e[3] = {atan2f(-M[5], M[8]), asinf(M[2]), atan2f(-M[1], M[0])};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float v = e[k] - o[3 + k];
    if (v < -PI) v += 2.f * PI;
    if (v > PI) v -= 2.f * PI;
    r[3 + k] = v * 100.f;
  }
  r[6] = a[6];
  bool bad = false;
#pragma unroll
  for (int k = 0; k < 7; ++k) { out[(size_t)i * 7 + k] = r[k]; bad |= (r[k] != r[k]); }
  if (bad && nan_flag) atomicOr(nan_flag, 1);
}

// ---------------------------------------------------------------------------------------------------------------------
// Sampling from the logistic mixture (logistic_decoder_rnn.py:234-258, validation / inference): per (token, action dim) a Gumbel-max
// choice of the mixture component, then inverse-CDF sampling of that logistic; the gripper command is gripper_bounds[argmax].
// Uniforms: injected (u_mix [token][dim][mix], u_inv [token][dim], tokens in (b, t) order) or Philox(seed, site / site + 1).
// One thread per (token, dim), dim == n_dims being the gripper.  out is [B][S][n_dims + has_gripper], batch-major.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void logistic_sample_kernel(const float* __restrict__ heads, int ldh, const float* __restrict__ u_mix, const float* __restrict__ u_inv,
                                       float* __restrict__ out, int B, int S, int b0, int Bm, int time_major, int n_dims, int n_mix, float log_scale_min,
                                       int has_gripper, float grip_lo, float grip_hi, unsigned long long seed, const unsigned long long* seed_ptr,
                                       unsigned site) {
  const int per_tok = n_dims + (has_gripper ? 1 : 0);
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= Bm * S * per_tok) return;
  const int tok = gid / per_tok, d = gid - tok * per_tok;  // tok = bl * S + t over the Bm sequences of this call
  const int bl = tok / S, t = tok - bl * S, b = b0 + bl;
  const float* h = heads + (size_t)(time_major ? t * B + b : b * S + t) * ldh;
  float* o = out + ((size_t)b * S + t) * per_tok;
  const int NM = n_dims * n_mix;
  if (d == n_dims) {  // gripper: argmax of the two logits (first index on ties, like torch.argmax)
    o[d] = h[3 * NM + 1] > h[3 * NM] ? grip_hi : grip_lo;
    return;
  }
  const float r1 = 1e-5f, r2 = 1.0f - 1e-5f;
  const unsigned long long sd = rng_seed(seed, seed_ptr);
  int best = 0;
  float best_v = -FLT_MAX;
  for (int k = 0; k < n_mix; ++k) {
    const size_t e = ((size_t)tok * n_dims + d) * n_mix + k;
    const float u = u_mix ? u_mix[e] : philox_uniform(sd, site, e);
    const float v = h[d * n_mix + k] - logf(-logf((r1 - r2) * u + r2));
    if (v > best_v) { best_v = v; best = k; }
  }
  const float mean = h[NM + d * n_mix + best];
  const float scale = expf(fmaxf(h[2 * NM + d * n_mix + best], log_scale_min));
  const size_t e2 = (size_t)tok * n_dims + d;
  const float u = (r1 - r2) * (u_inv ? u_inv[e2] : philox_uniform(sd, site + 1, e2)) + r2;
  o[d] = mean + scale * (logf(u) - logf(1.0f - u));
}

// Validation metrics (hulc.py:346-385): mae[b][d] = mean_t |pred - actions| for the first n_dims action dims; hits[b] = number of steps whose
// sampled gripper command (sign of pred[..., -1]) equals the ground-truth one.  One block per sequence.
__global__ void val_metrics_kernel(const float* __restrict__ pred, const float* __restrict__ actions, float* __restrict__ mae, float* __restrict__ hits, int S,
                                   int n_dims) {
  const int b = blockIdx.x, d = threadIdx.x;
  const int A = n_dims + 1;
  if (d < n_dims) {
    float s = 0.f;
    for (int t = 0; t < S; ++t) s += fabsf(pred[((size_t)b * S + t) * A + d] - actions[((size_t)b * S + t) * A + d]);
    mae[(size_t)b * n_dims + d] = s / (float)S;
  } else if (d == n_dims) {
    float c = 0.f;
    for (int t = 0; t < S; ++t) {
      const float g = pred[((size_t)b * S + t) * A + n_dims] > 0.f ? 1.f : -1.f;
      c += (actions[((size_t)b * S + t) * A + n_dims] == g) ? 1.f : 0.f;
    }
    hits[b] = c;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Discretised logistic mixture NLL + gripper CE (decoders/logistic_decoder_rnn.py:136-155,184-231), forward + gradient
// w.r.t. the head pre-activations in one pass.  One thread per (token, action dim).  heads holds B sequences (rows
// b*S+t, or t*B+b when time_major); the loss is taken over sequences [b0, b0+Bm) only and dheads is written for those.
//   heads row layout: [logit_probs n_dims*n_mix | means n_dims*n_mix | raw log_scales n_dims*n_mix | gripper 2]
//   losses[0] = mean_tokens sum_dims NLL, losses[1] = mean_tokens CE;   dheads = grad_scale * d(losses[0] + alpha*losses[1]) / d heads
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kMaxMix = 16;

__global__ void __launch_bounds__(256) logistic_loss_kernel(const float* __restrict__ heads, int ldh, const float* __restrict__ actions, int act_dim,
                                                            float* __restrict__ dheads, float* __restrict__ partial, unsigned* __restrict__ counter,
                                                            float* __restrict__ losses, int B, int S, int b0, int Bm, int time_major, int n_dims,
                                                            int n_mix, int num_classes, float log_scale_min, float act_min, float act_max,
                                                            int has_gripper, float gripper_alpha, float grad_scale) {
  __shared__ float red[32];
  __shared__ unsigned s_ticket;
  const int T = Bm * S;  // tokens of the sequences [b0, b0+Bm) — one modality of the batched step
  const int per_tok = n_dims + (has_gripper ? 1 : 0);
  int gid = blockIdx.x * blockDim.x + threadIdx.x;
  float nll = 0.f, ce = 0.f;
  if (gid < T * per_tok) {
    int tok = gid / per_tok, d = gid % per_tok;
    int b = b0 + (time_major ? tok % Bm : tok / S), t = time_major ? tok / Bm : tok % S;
    int row = time_major ? t * B + b : b * S + t;
    const float* a_tok = actions + ((size_t)b * S + t) * act_dim;
    const float* h = heads + (size_t)row * ldh;
    float* dh = dheads + (size_t)row * ldh;
    const float inv_T = grad_scale / (float)T;
    const int NM = n_dims * n_mix;
    if (d < n_dims) {
      float a = a_tok[d];
      const float* lp = h + d * n_mix;
      const float* mu = h + NM + d * n_mix;
      const float* lsr = h + 2 * NM + d * n_mix;
      float half_bin = (act_max - act_min) * 0.5f / (float)(num_classes - 1);
      float log_half_classes = logf((float)(num_classes - 1) * 0.5f);
      float tot[kMaxMix], dmu[kMaxMix], dls[kMaxMix];
      float mlp = -FLT_MAX;
      for (int k = 0; k < n_mix; ++k) mlp = fmaxf(mlp, lp[k]);
      float selp = 0.f;
      for (int k = 0; k < n_mix; ++k) selp += expf(lp[k] - mlp);
      float lse_p = mlp + logf(selp);
      float mt = -FLT_MAX;
      for (int k = 0; k < n_mix; ++k) {
        float ls = fmaxf(lsr[k], log_scale_min);
        float c = a - mu[k];
        float inv = expf(-ls);
        float u = inv * (c + half_bin), v = inv * (c - half_bin), m = inv * c;
        float su = sigmoidf(u), sv = sigmoidf(v);
        float delta = su - sv;
        float f, fu = 0.f, fv = 0.f, fm = 0.f, fls = 0.f;  // f and its partials w.r.t. u, v, m and (direct) ls
        if (a < act_min + 1e-3f) { f = u - softplusf(u); fu = 1.f - su; }
        else if (a > act_max - 1e-3f) { f = -softplusf(v); fv = -sv; }
        else if (delta > 1e-5f) { f = logf(fmaxf(delta, 1e-12f)); fu = su * (1.f - su) / delta; fv = -sv * (1.f - sv) / delta; }
        else { f = m - ls - 2.f * softplusf(m) - log_half_classes; fm = 1.f - 2.f * sigmoidf(m); fls = -1.f; }
        dmu[k] = -inv * (fu + fv + fm);
        dls[k] = -(u * fu + v * fv + m * fm) + fls;
        tot[k] = f + lp[k] - lse_p;
        mt = fmaxf(mt, tot[k]);
      }
      float se = 0.f;
      for (int k = 0; k < n_mix; ++k) se += expf(tot[k] - mt);
      float L = mt + logf(se);
      nll = -L;
      for (int k = 0; k < n_mix; ++k) {
        float w = expf(tot[k] - L);
        float pi = expf(lp[k] - lse_p);
        dh[d * n_mix + k] = (pi - w) * inv_T;
        dh[NM + d * n_mix + k] = -w * dmu[k] * inv_T;
        dh[2 * NM + d * n_mix + k] = (lsr[k] >= log_scale_min) ? -w * dls[k] * inv_T : 0.f;
      }
    } else {
      float g0 = h[3 * NM], g1 = h[3 * NM + 1];
      int label = (a_tok[act_dim - 1] == -1.f) ? 0 : 1;
      float mx = fmaxf(g0, g1);
      float lse = mx + logf(expf(g0 - mx) + expf(g1 - mx));
      ce = lse - (label ? g1 : g0);
      float p0 = expf(g0 - lse), p1 = expf(g1 - lse);
      dh[3 * NM] = gripper_alpha * (p0 - (label == 0 ? 1.f : 0.f)) * inv_T;
      dh[3 * NM + 1] = gripper_alpha * (p1 - (label == 1 ? 1.f : 0.f)) * inv_T;
    }
  }
  // deterministic two-level reduction: per-block partials, last block sums them in order
  float bn = block_sum(nll, red);
  float bc = block_sum(ce, red);
  if (threadIdx.x == 0) {
    partial[2 * blockIdx.x] = bn; partial[2 * blockIdx.x + 1] = bc;
    __threadfence();
    s_ticket = atomicAdd(counter, 1u);
  }
  __syncthreads();
  if (s_ticket == gridDim.x - 1 && threadIdx.x == 0) {
    __threadfence();
    float sn = 0.f, sc = 0.f;
    for (unsigned i = 0; i < gridDim.x; ++i) { sn += __ldcg(&partial[2 * i]); sc += __ldcg(&partial[2 * i + 1]); }
    losses[0] = sn / (float)T; losses[1] = sc / (float)T;
    *counter = 0u;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Discrete latent plan (hulc/utils/distributions.py:23-41, hulc/models/hulc.py:289-291,539-561).
// One warp per (sequence, category); lane = class (class_size must be 32).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) plan_discrete_fwd_kernel(const float* __restrict__ pr_logit, const float* __restrict__ pp_logit,
                                                                const float* __restrict__ u_in, const int* __restrict__ idx_in,
                                                                float* __restrict__ plan, int* __restrict__ idx_out, float* __restrict__ kl_rows,
                                                                int rows, unsigned long long seed, const unsigned long long* seed_ptr, unsigned site) {
  int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  float lq = pr_logit[(size_t)row * 32 + lane];
  float lp = pp_logit ? pp_logit[(size_t)row * 32 + lane] : 0.f;
  float mq = warp_max(lq), mp = warp_max(lp);
  float eq = expf(lq - mq), ep = expf(lp - mp);
  float sq = warp_sum(eq), sp = warp_sum(ep);
  float q = eq / sq;
  float logq = lq - mq - logf(sq), logp = lp - mp - logf(sp);
  if (kl_rows) {
    float kl = warp_sum(q * (logq - logp));
    if (lane == 0) kl_rows[row] = kl;
  }
  int idx;
  if (idx_in) idx = idx_in[row];
  else {
    float u = u_in ? u_in[row] : philox_uniform(rng_seed(seed, seed_ptr), site, (unsigned long long)row);
    // sequential inclusive prefix sum (same association as torch.cumsum) -> #{j : c_j <= u}
    float c = 0.f;
    int cnt = 0;
    for (int j = 0; j < 32; ++j) {
      c += __shfl_sync(0xffffffffu, q, j);
      cnt += (c <= u) ? 1 : 0;
    }
    idx = min(cnt, 31);
  }
  if (plan) plan[(size_t)row * 32 + lane] = (lane == idx) ? 1.f : 0.f;
  if (idx_out && lane == 0) idx_out[row] = idx;
}

// d pr = q (g - sum q g)  [straight-through]  +  c_rhs * q ((log q - log p) - KL)     d pp = c_lhs * (p - q)
__global__ void __launch_bounds__(256) plan_discrete_bwd_kernel(const float* __restrict__ pr_logit, const float* __restrict__ pp_logit,
                                                                const float* __restrict__ dplan, const float* __restrict__ dkl, float coef_lhs,
                                                                float coef_rhs, float* __restrict__ d_pr, float* __restrict__ d_pp, int rows) {
  int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  float lq = pr_logit[(size_t)row * 32 + lane];
  float mq = warp_max(lq);
  float eq = expf(lq - mq);
  float sq = warp_sum(eq);
  float q = eq / sq;
  float g = dplan ? dplan[(size_t)row * 32 + lane] : 0.f;
  float gq = warp_sum(q * g);
  float dq = q * (g - gq);
  if (pp_logit) {
    float up = dkl ? *dkl : 1.f;
    float lp = pp_logit[(size_t)row * 32 + lane];
    float mp = warp_max(lp);
    float ep = expf(lp - mp);
    float sp = warp_sum(ep);
    float p = ep / sp;
    float logq = lq - mq - logf(sq), logp = lp - mp - logf(sp);
    float t = logq - logp;
    float kl = warp_sum(q * t);
    dq += up * coef_rhs * q * (t - kl);
    d_pp[(size_t)row * 32 + lane] = up * coef_lhs * (p - q);
  }
  d_pr[(size_t)row * 32 + lane] = dq;
}

// ---------------------------------------------------------------------------------------------------------------------
// Continuous latent plan (distributions.py:28-29,55-59): state = [mean | raw_std], std = softplus(raw) + 1e-4,
// plan = mean + std * eps, KL(N_q || N_p) summed over dims.  One thread per (sequence, dim).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void plan_cont_fwd_kernel(const float* __restrict__ pr, const float* __restrict__ pp, const float* __restrict__ eps_in,
                                     float* __restrict__ plan, float* __restrict__ kl_elem, int Bn, int P, unsigned long long seed0,
                                     const unsigned long long* seed_ptr, unsigned site) {
  const unsigned long long seed = rng_seed(seed0, seed_ptr);
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Bn * P) return;
  int b = i / P, d = i % P;
  float mq = pr[(size_t)b * 2 * P + d], sq = softplusf(pr[(size_t)b * 2 * P + P + d]) + 1e-4f;
  float e;
  if (eps_in) e = eps_in[i];
  else {  // Box-Muller on two philox uniforms
    float u1 = fmaxf(philox_uniform(seed, site, 2ull * i), 1e-7f), u2 = philox_uniform(seed, site, 2ull * i + 1);
    e = sqrtf(-2.f * logf(u1)) * cosf(6.283185307179586f * u2);
  }
  plan[i] = mq + sq * e;
  if (kl_elem) {
    float mp = pp[(size_t)b * 2 * P + d], sp = softplusf(pp[(size_t)b * 2 * P + P + d]) + 1e-4f;
    float vr = (sq / sp) * (sq / sp), t1 = (mq - mp) / sp;
    kl_elem[i] = 0.5f * (vr + t1 * t1 - 1.f - logf(vr));
  }
}
__global__ void plan_cont_bwd_kernel(const float* __restrict__ pr, const float* __restrict__ pp, const float* __restrict__ eps_in,
                                     const float* __restrict__ dplan, const float* __restrict__ dkl, float coef_lhs, float coef_rhs,
                                     float* __restrict__ d_pr, float* __restrict__ d_pp, int Bn, int P, unsigned long long seed0,
                                     const unsigned long long* seed_ptr, unsigned site) {
  const unsigned long long seed = rng_seed(seed0, seed_ptr);
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Bn * P) return;
  int b = i / P, d = i % P;
  float rq = pr[(size_t)b * 2 * P + P + d], rp = pp[(size_t)b * 2 * P + P + d];
  float mq = pr[(size_t)b * 2 * P + d], sq = softplusf(rq) + 1e-4f;
  float mp = pp[(size_t)b * 2 * P + d], sp = softplusf(rp) + 1e-4f;
  float e;
  if (eps_in) e = eps_in[i];
  else {
    float u1 = fmaxf(philox_uniform(seed, site, 2ull * i), 1e-7f), u2 = philox_uniform(seed, site, 2ull * i + 1);
    e = sqrtf(-2.f * logf(u1)) * cosf(6.283185307179586f * u2);
  }
  float g = dplan ? dplan[i] : 0.f;
  float up = dkl ? *dkl : 1.f;
  // KL = 0.5 (sq^2/sp^2 + (mq-mp)^2/sp^2 - 1 - 2 log sq + 2 log sp)
  float dmq = (mq - mp) / (sp * sp), dsq = sq / (sp * sp) - 1.f / sq;
  float dmp = -dmq, dsp = -(sq * sq) / (sp * sp * sp) - (mq - mp) * (mq - mp) / (sp * sp * sp) + 1.f / sp;
  float gmq = g + up * coef_rhs * dmq, gsq = g * e + up * coef_rhs * dsq;
  d_pr[(size_t)b * 2 * P + d] = gmq;
  d_pr[(size_t)b * 2 * P + P + d] = gsq * sigmoidf(rq);
  d_pp[(size_t)b * 2 * P + d] = up * coef_lhs * dmp;
  d_pp[(size_t)b * 2 * P + P + d] = up * coef_lhs * dsp * sigmoidf(rp);
}

// out[0] = scale * sum_i x[i], fixed order (single block)
__global__ void __launch_bounds__(256) sum_kernel(const float* __restrict__ x, int n, float* __restrict__ out, float scale) {
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += x[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) out[0] = s * scale;
}

// ---------------------------------------------------------------------------------------------------------------------
// CLIP-style contrastive loss (hulc/models/hulc.py:650-695) on already projected features, forward + gradients, one CTA.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) clip_loss_kernel(const float* __restrict__ im, const float* __restrict__ tx, const float* __restrict__ logit_scale,
                                                        const unsigned char* __restrict__ mask, float* __restrict__ loss, float* __restrict__ d_im,
                                                        float* __restrict__ d_tx, float* __restrict__ d_logit_scale, int n, int D, float grad_scale) {
  HULC_DYN_SMEM(float, sm);
  __shared__ int s_nv;
  __shared__ float red[32];
  int* sel = reinterpret_cast<int*>(sm);       // [n] compacted indices of selected rows
  float* a = sm + n;                           // [n][D] normalised image features
  float* t = a + (size_t)n * D;                // [n][D] normalised text features
  float* na = t + (size_t)n * D;               // [n] norms
  float* nt = na + n;                          // [n]
  float* L = nt + n;                           // [n][n] logits, then gradient G
  float* rlse = L + (size_t)n * n;             // [n] row logsumexp
  float* clse = rlse + n;                      // [n] column logsumexp
  float* U = clse + n;                         // [2n][D] un-normalised feature gradients
  float* pr = U + 2 * (size_t)n * D;           // [2n] their projections on the features
  const int tid = threadIdx.x;
  int* flag = reinterpret_cast<int*>(pr);      // mask bytes fetched in parallel, compacted in row order by one thread
  for (int i = tid; i < n; i += blockDim.x) flag[i] = (!mask || mask[i]) ? 1 : 0;
  for (int i = tid; i < n * D; i += blockDim.x) { d_im[i] = 0.f; d_tx[i] = 0.f; }
  __syncthreads();
  if (tid == 0) {
    int c = 0;
    for (int i = 0; i < n; ++i)
      if (flag[i]) sel[c++] = i;
    s_nv = c;
  }
  __syncthreads();
  const int nv = s_nv;
  if (nv == 0) {  // reference: dummy pass scaled by 0 -> zero loss, zero gradients
    if (tid == 0) { loss[0] = 0.f; d_logit_scale[0] = 0.f; }
    return;
  }
  const float s = expf(logit_scale[0]);
  // selected rows -> shared memory (coalesced), then the norms and the normalisation work on shared memory only
  for (int e = tid; e < nv * D; e += blockDim.x) {
    const int r = e / D, k = e - r * D;
    a[e] = im[(size_t)sel[r] * D + k];
    t[e] = tx[(size_t)sel[r] * D + k];
  }
  __syncthreads();
  for (int r = tid; r < 2 * nv; r += blockDim.x) {
    const bool is_t = r >= nv;
    const int i = is_t ? r - nv : r;
    const float* p = (is_t ? t : a) + i * D;
    float q = 0.f;
    for (int k = 0; k < D; ++k) q += p[k] * p[k];
    (is_t ? nt : na)[i] = sqrtf(q);
  }
  __syncthreads();
  for (int e = tid; e < nv * D; e += blockDim.x) {
    const int r = e / D;
    a[e] = a[e] / na[r];
    t[e] = t[e] / nt[r];
  }
  __syncthreads();
  for (int e = tid; e < nv * nv; e += blockDim.x) {
    int i = e / nv, j = e % nv;
    float dot = 0.f;
    for (int k = 0; k < D; ++k) dot += a[i * D + k] * t[j * D + k];
    L[i * nv + j] = s * dot;
  }
  __syncthreads();
  for (int r = tid; r < 2 * nv; r += blockDim.x) {
    bool col = r >= nv;
    int i = col ? r - nv : r;
    float mx = -FLT_MAX;
    for (int j = 0; j < nv; ++j) mx = fmaxf(mx, col ? L[j * nv + i] : L[i * nv + j]);
    float se = 0.f;
    for (int j = 0; j < nv; ++j) se += expf((col ? L[j * nv + i] : L[i * nv + j]) - mx);
    (col ? clse : rlse)[i] = mx + logf(se);
  }
  __syncthreads();
  float part = 0.f;
  for (int i = tid; i < nv; i += blockDim.x) part += (rlse[i] - L[i * nv + i]) + (clse[i] - L[i * nv + i]);
  part = block_sum(part, red);
  if (tid == 0) loss[0] = 0.5f * part / (float)nv;
  __syncthreads();
  // G = dloss/dL ; d s accumulates sum G * (a.t)
  float ds = 0.f;
  const float c = grad_scale * 0.5f / (float)nv;
  for (int e = tid; e < nv * nv; e += blockDim.x) {
    int i = e / nv, j = e % nv;
    float l = L[e];
    float g = c * (expf(l - rlse[i]) + expf(l - clse[j]) - (i == j ? 2.f : 0.f));
    ds += g * (l / s);
    L[e] = g;
  }
  ds = block_sum(ds, red);
  if (tid == 0) d_logit_scale[0] = ds * s;
  __syncthreads();
  // d a_i = s * sum_j G_ij t_j ; d t_j = s * sum_i G_ij a_i ; then through the L2 normalisation.  One (row, feature) pair per
  // thread and pass; the un-normalised gradients are parked in shared memory.
  for (int e = tid; e < 2 * nv * D; e += blockDim.x) {
    const int r = e / D, k = e - r * D;
    const bool is_t = r >= nv;
    const int i = is_t ? r - nv : r;
    const float* other = is_t ? a : t;
    float gk = 0.f;
    for (int j = 0; j < nv; ++j) gk += (is_t ? L[j * nv + i] : L[i * nv + j]) * other[j * D + k];
    U[e] = gk * s;
  }
  __syncthreads();
  for (int r = tid; r < 2 * nv; r += blockDim.x) {
    const float* self = r >= nv ? t + (r - nv) * D : a + r * D;
    float proj = 0.f;
    for (int k = 0; k < D; ++k) proj += U[r * D + k] * self[k];
    pr[r] = proj;
  }
  __syncthreads();
  for (int e = tid; e < 2 * nv * D; e += blockDim.x) {
    const int r = e / D, k = e - r * D;
    const bool is_t = r >= nv;
    const int i = is_t ? r - nv : r;
    const float self = (is_t ? t : a)[i * D + k];
    (is_t ? d_tx : d_im)[(size_t)sel[i] * D + k] = (U[e] - self * pr[r]) / (is_t ? nt[i] : na[i]);
  }
}

}  // namespace

HULC_API int hulc_world_to_tcp(const float* actions, const float* robot_obs, int obs_dim, float* out, int n_tokens, int* nan_flag, void* stream) {
  if (n_tokens <= 0) return 0;
  HULC_LAUNCH(world_to_tcp_kernel, dim3(hulc_cdiv(n_tokens, 128)), dim3(128), 0, (cudaStream_t)stream, actions, robot_obs, obs_dim, out, n_tokens,
              nan_flag);
  HULC_RETURN_LAST();
}

HULC_API int hulc_tcp_to_world(const float* actions, const float* robot_obs, int obs_dim, float* out, int n_tokens, int* nan_flag, void* stream) {
  if (n_tokens <= 0) return 0;
  HULC_LAUNCH(tcp_to_world_kernel, dim3(hulc_cdiv(n_tokens, 128)), dim3(128), 0, (cudaStream_t)stream, actions, robot_obs, obs_dim, out, n_tokens,
              nan_flag);
  HULC_RETURN_LAST();
}

HULC_API int hulc_logistic_sample(const float* heads, int ldh, const float* u_mix, const float* u_inv, float* out, int B, int S, int b0, int Bm,
                                  int time_major, int n_dims, int n_mix, float log_scale_min, int has_gripper, float grip_lo, float grip_hi,
                                  unsigned long long seed, unsigned site, void* stream) {
  if (Bm * S <= 0) return 0;
  if (b0 < 0 || b0 + Bm > B || n_mix <= 0) return (int)cudaErrorInvalidValue;
  const int total = Bm * S * (n_dims + (has_gripper ? 1 : 0));
  HULC_LAUNCH(logistic_sample_kernel, dim3(hulc_cdiv(total, 128)), dim3(128), 0, (cudaStream_t)stream, heads, ldh, u_mix, u_inv, out, B, S, b0, Bm,
              time_major, n_dims, n_mix, log_scale_min, has_gripper, grip_lo, grip_hi, seed, g_hulc_rng_offset_ptr, site);
  HULC_RETURN_LAST();
}

HULC_API int hulc_val_metrics(const float* pred, const float* actions, float* mae, float* hits, int B, int S, int n_dims, void* stream) {
  if (B <= 0 || S <= 0) return 0;
  if (n_dims + 1 > 32) return (int)cudaErrorInvalidValue;
  HULC_LAUNCH(val_metrics_kernel, dim3(B), dim3(32), 0, (cudaStream_t)stream, pred, actions, mae, hits, S, n_dims);
  HULC_RETURN_LAST();
}

HULC_API int hulc_logistic_loss(const float* heads, int ldh, const float* actions, int act_dim, float* dheads, float* losses, int B, int S,
                                int b0, int Bm, int time_major, int n_dims, int n_mix, int num_classes, float log_scale_min, float act_min,
                                float act_max, int has_gripper, float gripper_alpha, float grad_scale, float* workspace, size_t workspace_bytes,
                                void* stream) {
  if (Bm * S <= 0) return 0;
  if (b0 < 0 || b0 + Bm > B) return (int)cudaErrorInvalidValue;
  if (n_mix > kMaxMix) return (int)cudaErrorInvalidValue;
  int total = Bm * S * (n_dims + (has_gripper ? 1 : 0));
  int blocks = hulc_cdiv(total, 256);
  if (workspace_bytes < 4096 + sizeof(float) * 2 * (size_t)blocks) return (int)cudaErrorInvalidValue;
  HULC_LAUNCH(logistic_loss_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, heads, ldh, actions, act_dim, dheads, workspace + 1024,
              reinterpret_cast<unsigned*>(workspace) + 1023, losses, B, S, b0, Bm, time_major, n_dims, n_mix, num_classes, log_scale_min, act_min,
              act_max, has_gripper, gripper_alpha, grad_scale);
  HULC_RETURN_LAST();
}

HULC_API int hulc_plan_discrete_fwd(const float* pr_logit, const float* pp_logit, const float* u, const int* idx_in, float* plan, int* idx_out,
                                    float* kl_rows, int rows, int class_size, unsigned long long seed, unsigned site, void* stream) {
  if (rows <= 0) return 0;
  if (class_size != 32) return (int)cudaErrorInvalidValue;
  HULC_LAUNCH(plan_discrete_fwd_kernel, dim3(hulc_cdiv(rows, 8)), dim3(256), 0, (cudaStream_t)stream, pr_logit, pp_logit, u, idx_in, plan, idx_out,
              kl_rows, rows, seed, g_hulc_rng_offset_ptr, site);
  HULC_RETURN_LAST();
}

HULC_API int hulc_plan_discrete_bwd(const float* pr_logit, const float* pp_logit, const float* dplan, const float* dkl, float coef_lhs,
                                    float coef_rhs, float* d_pr, float* d_pp, int rows, int class_size, void* stream) {
  if (rows <= 0) return 0;
  if (class_size != 32) return (int)cudaErrorInvalidValue;
  HULC_LAUNCH(plan_discrete_bwd_kernel, dim3(hulc_cdiv(rows, 8)), dim3(256), 0, (cudaStream_t)stream, pr_logit, pp_logit, dplan, dkl, coef_lhs,
              coef_rhs, d_pr, d_pp, rows);
  HULC_RETURN_LAST();
}

HULC_API int hulc_plan_cont_fwd(const float* pr_state, const float* pp_state, const float* eps, float* plan, float* kl_elem, int batch,
                                int plan_features, unsigned long long seed, unsigned site, void* stream) {
  int n = batch * plan_features;
  if (n <= 0) return 0;
  HULC_LAUNCH(plan_cont_fwd_kernel, dim3(hulc_cdiv(n, 256)), dim3(256), 0, (cudaStream_t)stream, pr_state, pp_state, eps, plan, kl_elem, batch,
              plan_features, seed, g_hulc_rng_offset_ptr, site);
  HULC_RETURN_LAST();
}

HULC_API int hulc_plan_cont_bwd(const float* pr_state, const float* pp_state, const float* eps, const float* dplan, const float* dkl,
                                float coef_lhs, float coef_rhs, float* d_pr, float* d_pp, int batch, int plan_features, unsigned long long seed,
                                unsigned site, void* stream) {
  int n = batch * plan_features;
  if (n <= 0) return 0;
  HULC_LAUNCH(plan_cont_bwd_kernel, dim3(hulc_cdiv(n, 256)), dim3(256), 0, (cudaStream_t)stream, pr_state, pp_state, eps, dplan, dkl, coef_lhs,
              coef_rhs, d_pr, d_pp, batch, plan_features, seed, g_hulc_rng_offset_ptr, site);
  HULC_RETURN_LAST();
}

HULC_API int hulc_sum(const float* x, int n, float* out, float scale, void* stream) {
  HULC_LAUNCH(sum_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, x, n, out, scale);
  HULC_RETURN_LAST();
}

// ---------------------------------------------------------------------------------------------------------------------
// Auxiliary losses of the ablation configs (hulc/models/hulc.py:567-648): mean cosine distance between the regressed and the
// true language embedding (BC-Z), binary cross entropy with logits over matching / rolled pairs (MIA).  One small CTA each;
// rows are reduced by warps in a fixed order, so the results are deterministic.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cosine_loss_kernel(const float* __restrict__ pred, int ldp, const float* __restrict__ tgt, int ldt, float* __restrict__ dpred,
                                                         int ldd, float* __restrict__ loss, int B, int D, float grad_scale) {
  __shared__ float row_loss[256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int b = warp; b < B; b += nw) {
    const float* p = pred + (size_t)b * ldp;
    const float* g = tgt + (size_t)b * ldt;
    float pg = 0.f, pp = 0.f, gg = 0.f;
    for (int j = lane; j < D; j += 32) { pg += p[j] * g[j]; pp += p[j] * p[j]; gg += g[j] * g[j]; }
    pg = warp_sum(pg); pp = warp_sum(pp); gg = warp_sum(gg);
    const float np_ = sqrtf(pp), ng = sqrtf(gg), cosv = pg / (np_ * ng);
    // d(1 - cos)/dp = -(g / (|p||g|) - cos * p / |p|^2), averaged over the B rows
    const float k = grad_scale / (float)B;
    for (int j = lane; j < D; j += 32) dpred[(size_t)b * ldd + j] = -k * (g[j] / (np_ * ng) - cosv * p[j] / pp);
    if (lane == 0) row_loss[b] = 1.f - cosv;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float acc = 0.f;
    for (int b = 0; b < B; ++b) acc += row_loss[b];
    loss[0] = acc / (float)B;
  }
}

__global__ void __launch_bounds__(256) bce_logits_kernel(const float* __restrict__ x, float* __restrict__ dx, float* __restrict__ loss, int n_pos, int n_neg,
                                                        float grad_scale) {
  __shared__ float part[256];
  const int n = n_pos + n_neg;
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = x[i], y = i < n_pos ? 1.f : 0.f;
    acc += fmaxf(v, 0.f) - v * y + log1pf(expf(-fabsf(v)));  // the numerically stable form torch uses
    dx[i] = grad_scale * (1.f / (1.f + expf(-v)) - y) / (float)n;
  }
  part[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)blockDim.x; ++i) t += part[i];
    loss[0] = t / (float)n;
  }
}

// See include/hulc_b200.h.
HULC_API int hulc_cosine_loss(const float* pred, int ldp, const float* target, int ldt, float* dpred, int ldd, float* loss, int B, int D, float grad_scale,
                              void* stream) {
  if (B <= 0 || D <= 0) return 0;
  if (B > 256) return (int)cudaErrorInvalidValue;
  HULC_LAUNCH(cosine_loss_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, pred, ldp, target, ldt, dpred, ldd, loss, B, D, grad_scale);
  HULC_RETURN_LAST();
}
HULC_API int hulc_bce_logits_loss(const float* logits, float* dlogits, float* loss, int n_pos, int n_neg, float grad_scale, void* stream) {
  if (n_pos + n_neg <= 0) return 0;
  HULC_LAUNCH(bce_logits_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, logits, dlogits, loss, n_pos, n_neg, grad_scale);
  HULC_RETURN_LAST();
}

HULC_API int hulc_clip_loss(const float* im, const float* tx, const float* logit_scale, const unsigned char* mask, float* loss, float* d_im,
                            float* d_tx, float* d_logit_scale, int n, int D, float grad_scale, void* stream) {
  if (n <= 0) return 0;
  size_t smem = sizeof(float) * ((size_t)n + 2 * (size_t)n * D + 2 * n + (size_t)n * n + 2 * n + 2 * (size_t)n * D + 2 * n);
  if (smem > 200 * 1024) return (int)cudaErrorInvalidValue;
  auto kfn = clip_loss_kernel;
  if (smem > 48 * 1024) HULC_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  HULC_LAUNCH(kfn, dim3(1), dim3(256), smem, (cudaStream_t)stream, im, tx, logit_scale, mask, loss, d_im, d_tx, d_logit_scale, n, D, grad_scale);
  HULC_RETURN_LAST();
}
