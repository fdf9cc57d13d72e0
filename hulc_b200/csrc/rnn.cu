// rnn.cu — gate math of torch.nn.GRU (decoders/utils/rnn.py:27-36, gate order r|z|n in the packed weights) as one
// fused elementwise kernel per time step, forward and backward.  The matrix products around it (x W_ih^T for all
// steps at once, h W_hh^T per step) are hulc_gemm calls; the Elman RNN (rnn.py:5-14) needs no kernel of its own —
// its step is a hulc_gemm with the input projection as addend and ReLU/tanh in the epilogue.
#include "common.cuh"

namespace {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// gi = W_i x + b_i (row stride ldgi), gh = W_h h_prev + b_h (ldgh), both [B, 3H] as r|z|n.
// h = (1 - z) n + z h_prev;  saved[B,4H] = r | z | n | gh_n for the backward pass.
__global__ void gru_gates_fwd_kernel(const float* __restrict__ gi, int ldgi, const float* __restrict__ gh, int ldgh, const float* __restrict__ hprev,
                                     int ldhp, float* __restrict__ h, int ldh, float* __restrict__ saved, int B, int H) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * H) return;
  int b = i / H, j = i % H;
  const float* a = gi + (size_t)b * ldgi;
  const float* c = gh + (size_t)b * ldgh;
  float r = sigmoidf_(a[j] + c[j]);
  float z = sigmoidf_(a[H + j] + c[H + j]);
  float ghn = c[2 * H + j];
  float n = tanhf(a[2 * H + j] + r * ghn);
  float hp = hprev ? hprev[(size_t)b * ldhp + j] : 0.f;
  h[(size_t)b * ldh + j] = (1.f - z) * n + z * hp;
  float* s = saved + (size_t)b * 4 * H;
  s[j] = r; s[H + j] = z; s[2 * H + j] = n; s[3 * H + j] = ghn;
}

// dh = dh_above + dh_rec.  Writes dgi, dgh ([B,3H]) and dh_carry = dh * z (the direct path to h_prev; the caller adds
// dgh W_hh with a GEMM).
__global__ void gru_gates_bwd_kernel(const float* __restrict__ dh_above, int lda, const float* __restrict__ dh_rec, int ldr,
                                     const float* __restrict__ saved, const float* __restrict__ hprev, int ldhp, float* __restrict__ dgi, int ldgi,
                                     float* __restrict__ dgh, int ldgh, float* __restrict__ dh_carry, int ldc, int B, int H) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * H) return;
  int b = i / H, j = i % H;
  float dh = (dh_above ? dh_above[(size_t)b * lda + j] : 0.f) + (dh_rec ? dh_rec[(size_t)b * ldr + j] : 0.f);
  const float* s = saved + (size_t)b * 4 * H;
  float r = s[j], z = s[H + j], n = s[2 * H + j], ghn = s[3 * H + j];
  float hp = hprev ? hprev[(size_t)b * ldhp + j] : 0.f;
  float dn_pre = dh * (1.f - z) * (1.f - n * n);
  float dz_pre = dh * (hp - n) * z * (1.f - z);
  float dr_pre = dn_pre * ghn * r * (1.f - r);
  float* di = dgi + (size_t)b * ldgi;
  float* dg = dgh + (size_t)b * ldgh;
  di[j] = dr_pre; di[H + j] = dz_pre; di[2 * H + j] = dn_pre;
  dg[j] = dr_pre; dg[H + j] = dz_pre; dg[2 * H + j] = dn_pre * r;
  dh_carry[(size_t)b * ldc + j] = dh * z;
}

}  // namespace

HULC_API int hulc_gru_gates_fwd(const float* gi, int ldgi, const float* gh, int ldgh, const float* hprev, int ldhp, float* h, int ldh, float* saved,
                                int B, int H, void* stream) {
  if (B * H <= 0) return 0;
  HULC_LAUNCH(gru_gates_fwd_kernel, dim3(hulc_cdiv((long long)B * H, 256)), dim3(256), 0, (cudaStream_t)stream, gi, ldgi, gh, ldgh, hprev, ldhp, h, ldh,
              saved, B, H);
  HULC_RETURN_LAST();
}

HULC_API int hulc_gru_gates_bwd(const float* dh_above, int lda, const float* dh_rec, int ldr, const float* saved, const float* hprev, int ldhp,
                                float* dgi, int ldgi, float* dgh, int ldgh, float* dh_carry, int ldc, int B, int H, void* stream) {
  if (B * H <= 0) return 0;
  HULC_LAUNCH(gru_gates_bwd_kernel, dim3(hulc_cdiv((long long)B * H, 256)), dim3(256), 0, (cudaStream_t)stream, dh_above, lda, dh_rec, ldr, saved, hprev,
              ldhp, dgi, ldgi, dgh, ldgh, dh_carry, ldc, B, H);
  HULC_RETURN_LAST();
}
