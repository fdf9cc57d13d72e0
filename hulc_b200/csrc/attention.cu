// attention.cu — multi-head self-attention for the posterior transformer (plan_recognition_net.py:83-90 builds
// nn.TransformerEncoderLayer(d=128, nhead=8)): S <= 64 keys, head_dim <= 32, so one warp owns a whole (sequence, head)
// problem: K/V live in shared memory, lane i owns query row i, the S scores of a row stay in registers and the softmax
// needs no cross-lane traffic.  Tokens are batch-first rows (b*S + s) of the packed qkv matrix [T, 3*D].
#include "common.cuh"

namespace {

constexpr int kMaxS = 64;
constexpr int kMaxDh = 32;

// probs [B,H,S,S] keeps the pre-dropout softmax for the backward pass; dropout index = ((b*H+h)*S + i)*S + j
__global__ void __launch_bounds__(32) attention_fwd_kernel(const float* __restrict__ qkv, float* __restrict__ out, float* __restrict__ probs, int B,
                                                           int S, int H, int dh, float scale, DropSpec drop) {
  HULC_DYN_SMEM(float, sm);
  const int D = H * dh, ldq = 3 * D;
  const int b = blockIdx.x / H, h = blockIdx.x % H, lane = threadIdx.x;
  float* Ks = sm;                        // [S][dh+1]
  float* Vs = sm + (size_t)S * (dh + 1);  // [S][dh+1]
  for (int e = lane; e < S * dh; e += 32) {
    int j = e / dh, d = e % dh;
    const float* row = qkv + (size_t)(b * S + j) * ldq + h * dh + d;
    Ks[j * (dh + 1) + d] = row[D];
    Vs[j * (dh + 1) + d] = row[2 * D];
  }
  __syncwarp();
  for (int i = lane; i < S; i += 32) {
    float q[kMaxDh], sc[kMaxS], o[kMaxDh];
    const float* qrow = qkv + (size_t)(b * S + i) * ldq + h * dh;
#pragma unroll
    for (int d = 0; d < kMaxDh; ++d) { q[d] = (d < dh) ? qrow[d] * scale : 0.f; o[d] = 0.f; }
    float mx = -FLT_MAX;
#pragma unroll
    for (int j = 0; j < kMaxS; ++j) {
      if (j < S) {
        float s = 0.f;
#pragma unroll
        for (int d = 0; d < kMaxDh; ++d)
          if (d < dh) s = fmaf(q[d], Ks[j * (dh + 1) + d], s);
        sc[j] = s; mx = fmaxf(mx, s);
      }
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxS; ++j)
      if (j < S) { sc[j] = expf(sc[j] - mx); sum += sc[j]; }
    float inv = 1.f / sum;
    size_t pbase = ((size_t)(b * H + h) * S + i) * S;
#pragma unroll
    for (int j = 0; j < kMaxS; ++j) {
      if (j < S) {
        float p = sc[j] * inv;
        probs[pbase + j] = p;
        float pd = p * drop_factor(drop, pbase + j);
#pragma unroll
        for (int d = 0; d < kMaxDh; ++d)
          if (d < dh) o[d] = fmaf(pd, Vs[j * (dh + 1) + d], o[d]);
      }
    }
    float* orow = out + (size_t)(b * S + i) * D + h * dh;
#pragma unroll
    for (int d = 0; d < kMaxDh; ++d)
      if (d < dh) orow[d] = o[d];
  }
}

__global__ void __launch_bounds__(32) attention_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ probs, const float* __restrict__ dout,
                                                           float* __restrict__ dqkv, int B, int S, int H, int dh, float scale, DropSpec drop) {
  HULC_DYN_SMEM(float, sm);
  const int D = H * dh, ldq = 3 * D, P = dh + 1, SP = S + 1;
  const int b = blockIdx.x / H, h = blockIdx.x % H, lane = threadIdx.x;
  float* Qs = sm;                    // [S][P]
  float* Ks = Qs + (size_t)S * P;
  float* Vs = Ks + (size_t)S * P;
  float* dOs = Vs + (size_t)S * P;
  float* dSs = dOs + (size_t)S * P;  // [S][SP]  dS_ij
  float* Pds = dSs + (size_t)S * SP;  // [S][SP]  dropped probabilities
  for (int e = lane; e < S * dh; e += 32) {
    int j = e / dh, d = e % dh;
    const float* row = qkv + (size_t)(b * S + j) * ldq + h * dh + d;
    Qs[j * P + d] = row[0];
    Ks[j * P + d] = row[D];
    Vs[j * P + d] = row[2 * D];
    dOs[j * P + d] = dout[(size_t)(b * S + j) * D + h * dh + d];
  }
  __syncwarp();
  // phase 1: lane = query row i -> dS_i,:  and dQ_i
  for (int i = lane; i < S; i += 32) {
    size_t pbase = ((size_t)(b * H + h) * S + i) * S;
    float dp[kMaxS];
    float dot = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxS; ++j) {
      if (j < S) {
        float p = probs[pbase + j];
        float f = drop_factor(drop, pbase + j);
        float g = 0.f;
        for (int d = 0; d < dh; ++d) g = fmaf(dOs[i * P + d], Vs[j * P + d], g);
        g *= f;  // d loss / d p_ij
        Pds[i * SP + j] = p * f;
        dp[j] = g;
        dot += g * p;
      }
    }
    float dq[kMaxDh];
#pragma unroll
    for (int d = 0; d < kMaxDh; ++d) dq[d] = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxS; ++j) {
      if (j < S) {
        float ds = probs[pbase + j] * (dp[j] - dot);
        dSs[i * SP + j] = ds;
#pragma unroll
        for (int d = 0; d < kMaxDh; ++d)
          if (d < dh) dq[d] = fmaf(ds, Ks[j * P + d], dq[d]);
      }
    }
    float* dst = dqkv + (size_t)(b * S + i) * ldq + h * dh;
#pragma unroll
    for (int d = 0; d < kMaxDh; ++d)
      if (d < dh) dst[d] = dq[d] * scale;
  }
  __syncwarp();
  // phase 2: lane = key row j -> dK_j = scale * sum_i dS_ij q_i ,  dV_j = sum_i Pd_ij dO_i
  for (int j = lane; j < S; j += 32) {
    float dk[kMaxDh], dv[kMaxDh];
#pragma unroll
    for (int d = 0; d < kMaxDh; ++d) dk[d] = dv[d] = 0.f;
    for (int i = 0; i < S; ++i) {
      float ds = dSs[i * SP + j], pd = Pds[i * SP + j];
#pragma unroll
      for (int d = 0; d < kMaxDh; ++d)
        if (d < dh) { dk[d] = fmaf(ds, Qs[i * P + d], dk[d]); dv[d] = fmaf(pd, dOs[i * P + d], dv[d]); }
    }
    float* dst = dqkv + (size_t)(b * S + j) * ldq + h * dh;
#pragma unroll
    for (int d = 0; d < kMaxDh; ++d)
      if (d < dh) { dst[D + d] = dk[d] * scale; dst[2 * D + d] = dv[d]; }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Block-parallel variants for S in {4, 8, 16, 32, 64} (S divides 128): one CTA of 128 threads per (sequence, head), G = 128 / S
// lanes per query row (all in one warp).  The single-warp kernels above serialise 32 keys x dh per lane and run one warp per
// CTA; here a lane owns S / G keys (then dh / G output columns) and row statistics meet through warp shuffles.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kAttThreads = 128;

__device__ __forceinline__ float group_max(float v, int G) {
  for (int o = G >> 1; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float group_sum(float v, int G) {
  for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(kAttThreads) attention_fwd_block_kernel(const float* __restrict__ qkv, float* __restrict__ out, float* __restrict__ probs,
                                                                          int B, int S, int H, int dh, float scale, DropSpec drop) {
  HULC_DYN_SMEM(float, sm);
  const int D = H * dh, ldq = 3 * D, P = dh + 1, SP = S + 1, G = kAttThreads / S;
  const int b = blockIdx.x / H, h = blockIdx.x % H, t = threadIdx.x;
  float* Qs = sm;  // [S][P], pre-scaled
  float* Ks = Qs + (size_t)S * P;
  float* Vs = Ks + (size_t)S * P;
  float* Ps = Vs + (size_t)S * P;  // [S][SP] scores, then dropped probabilities
  for (int e = t; e < S * dh; e += kAttThreads) {
    const int j = e / dh, d = e - j * dh;
    const float* row = qkv + (size_t)(b * S + j) * ldq + h * dh + d;
    Qs[j * P + d] = row[0] * scale;
    Ks[j * P + d] = row[D];
    Vs[j * P + d] = row[2 * D];
  }
  __syncthreads();
  const int i = t / G, g = t - i * G;
  float mx = -FLT_MAX;
  for (int j = g; j < S; j += G) {
    float a = 0.f;
    for (int d = 0; d < dh; ++d) a = fmaf(Qs[i * P + d], Ks[j * P + d], a);
    Ps[i * SP + j] = a;
    mx = fmaxf(mx, a);
  }
  mx = group_max(mx, G);
  float sum = 0.f;
  for (int j = g; j < S; j += G) {
    const float e = expf(Ps[i * SP + j] - mx);
    Ps[i * SP + j] = e;
    sum += e;
  }
  sum = group_sum(sum, G);
  const float inv = 1.f / sum;
  const size_t pbase = ((size_t)(b * H + h) * S + i) * S;
  for (int j = g; j < S; j += G) {
    const float pr = Ps[i * SP + j] * inv;
    probs[pbase + j] = pr;
    Ps[i * SP + j] = pr * drop_factor(drop, pbase + j);
  }
  __syncwarp();  // the G lanes of a row share a warp
  float* orow = out + (size_t)(b * S + i) * D + h * dh;
  for (int d = g; d < dh; d += G) {
    float o = 0.f;
    for (int j = 0; j < S; ++j) o = fmaf(Ps[i * SP + j], Vs[j * P + d], o);
    orow[d] = o;
  }
}

__global__ void __launch_bounds__(kAttThreads) attention_bwd_block_kernel(const float* __restrict__ qkv, const float* __restrict__ probs,
                                                                          const float* __restrict__ dout, float* __restrict__ dqkv, int B, int S, int H, int dh,
                                                                          float scale, DropSpec drop) {
  HULC_DYN_SMEM(float, sm);
  const int D = H * dh, ldq = 3 * D, P = dh + 1, SP = S + 1, G = kAttThreads / S;
  const int b = blockIdx.x / H, h = blockIdx.x % H, t = threadIdx.x;
  float* Qs = sm;
  float* Ks = Qs + (size_t)S * P;
  float* Vs = Ks + (size_t)S * P;
  float* dOs = Vs + (size_t)S * P;
  float* Pp = dOs + (size_t)S * P;    // [S][SP] probabilities
  float* dSs = Pp + (size_t)S * SP;   // [S][SP] d loss / d p, then dS
  float* Pds = dSs + (size_t)S * SP;  // [S][SP] dropped probabilities
  for (int e = t; e < S * dh; e += kAttThreads) {
    const int j = e / dh, d = e - j * dh;
    const float* row = qkv + (size_t)(b * S + j) * ldq + h * dh + d;
    Qs[j * P + d] = row[0];
    Ks[j * P + d] = row[D];
    Vs[j * P + d] = row[2 * D];
    dOs[j * P + d] = dout[(size_t)(b * S + j) * D + h * dh + d];
  }
  __syncthreads();
  const int i = t / G, g = t - i * G;
  {  // phase 1: query row i
    const size_t pbase = ((size_t)(b * H + h) * S + i) * S;
    float dot = 0.f;
    for (int j = g; j < S; j += G) {
      const float pr = probs[pbase + j], f = drop_factor(drop, pbase + j);
      float a = 0.f;
      for (int d = 0; d < dh; ++d) a = fmaf(dOs[i * P + d], Vs[j * P + d], a);
      a *= f;  // d loss / d p_ij
      Pp[i * SP + j] = pr;
      Pds[i * SP + j] = pr * f;
      dSs[i * SP + j] = a;
      dot += a * pr;
    }
    dot = group_sum(dot, G);
    for (int j = g; j < S; j += G) dSs[i * SP + j] = Pp[i * SP + j] * (dSs[i * SP + j] - dot);
    __syncwarp();
    float* dst = dqkv + (size_t)(b * S + i) * ldq + h * dh;
    for (int d = g; d < dh; d += G) {
      float a = 0.f;
      for (int j = 0; j < S; ++j) a = fmaf(dSs[i * SP + j], Ks[j * P + d], a);
      dst[d] = a * scale;
    }
  }
  __syncthreads();
  {  // phase 2: key row j = i
    const int j = i;
    float* dst = dqkv + (size_t)(b * S + j) * ldq + h * dh;
    for (int d = g; d < dh; d += G) {
      float dk = 0.f, dv = 0.f;
      for (int r = 0; r < S; ++r) {
        dk = fmaf(dSs[r * SP + j], Qs[r * P + d], dk);
        dv = fmaf(Pds[r * SP + j], dOs[r * P + d], dv);
      }
      dst[D + d] = dk * scale;
      dst[2 * D + d] = dv;
    }
  }
}

inline bool block_variant_ok(int S) { return S >= 4 && S <= 64 && (kAttThreads % S) == 0; }

}  // namespace

HULC_API int hulc_attention_fwd(const float* qkv, float* out, float* probs, int B, int S, int H, int dh, float drop_p, unsigned long long drop_seed,
                                unsigned drop_site, const unsigned char* drop_keep, void* stream) {
  if (B <= 0) return 0;
  if (S > kMaxS || dh > kMaxDh || S <= 0) return (int)cudaErrorInvalidValue;
  if (block_variant_ok(S)) {
    const size_t sm = sizeof(float) * (3 * (size_t)S * (dh + 1) + (size_t)S * (S + 1));
    HULC_LAUNCH(attention_fwd_block_kernel, dim3(B * H), dim3(kAttThreads), sm, (cudaStream_t)stream, qkv, out, probs, B, S, H, dh, 1.0f / sqrtf((float)dh),
                make_drop(drop_p, drop_seed, drop_site, drop_keep));
    HULC_RETURN_LAST();
  }
  size_t smem = sizeof(float) * 2 * (size_t)S * (dh + 1);
  HULC_LAUNCH(attention_fwd_kernel, dim3(B * H), dim3(32), smem, (cudaStream_t)stream, qkv, out, probs, B, S, H, dh, 1.0f / sqrtf((float)dh),
              make_drop(drop_p, drop_seed, drop_site, drop_keep));
  HULC_RETURN_LAST();
}

HULC_API int hulc_attention_bwd(const float* qkv, const float* probs, const float* dout, float* dqkv, int B, int S, int H, int dh, float drop_p,
                                unsigned long long drop_seed, unsigned drop_site, const unsigned char* drop_keep, void* stream) {
  if (B <= 0) return 0;
  if (S > kMaxS || dh > kMaxDh || S <= 0) return (int)cudaErrorInvalidValue;
  if (block_variant_ok(S)) {
    const size_t sm = sizeof(float) * (4 * (size_t)S * (dh + 1) + 3 * (size_t)S * (S + 1));
    auto kb = attention_bwd_block_kernel;
    if (sm > 48 * 1024) HULC_TRY(cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    HULC_LAUNCH(kb, dim3(B * H), dim3(kAttThreads), sm, (cudaStream_t)stream, qkv, probs, dout, dqkv, B, S, H, dh, 1.0f / sqrtf((float)dh),
                make_drop(drop_p, drop_seed, drop_site, drop_keep));
    HULC_RETURN_LAST();
  }
  size_t smem = sizeof(float) * (4 * (size_t)S * (dh + 1) + 2 * (size_t)S * (S + 1));
  auto kfn = attention_bwd_kernel;
  if (smem > 48 * 1024) HULC_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  HULC_LAUNCH(kfn, dim3(B * H), dim3(32), smem, (cudaStream_t)stream, qkv, probs, dout, dqkv, B, S, H, dh, 1.0f / sqrtf((float)dh),
              make_drop(drop_p, drop_seed, drop_site, drop_keep));
  HULC_RETURN_LAST();
}
