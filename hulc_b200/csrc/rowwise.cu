// rowwise.cu — warp-per-row kernels: LayerNorm (+residual +dropout) fwd/bwd, spatial softmax fwd/bwd,
// position-embedding add, strided copies/reductions, fused Adam.  All HBM-bound; rows map to warps, lanes stride the
// row so global accesses are coalesced; reductions are warp shuffles.
#include "common.cuh"

namespace {

constexpr int kWarpsPerBlock = 8;

// ---------------------------------------------------------------------------------------------------------------------
// LayerNorm:  z = res + drop(x)   (z = x when res == null);   y = (z - mean) * rstd * w + b
// ---------------------------------------------------------------------------------------------------------------------
template <int MAXV>  // elements per lane held in registers: D <= 32*MAXV
__global__ void __launch_bounds__(kWarpsPerBlock * 32) layernorm_fwd_kernel(
    const float* __restrict__ x, int ldx, const float* __restrict__ res, int ldres, const float* __restrict__ w,
    const float* __restrict__ b, float* __restrict__ y, int ldy, float* __restrict__ z, int ldz, float* __restrict__ stats,
    int rows, int D, float eps, DropSpec drop) {
  int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  float v[MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int c = lane + 32 * i;
    float t = 0.f;
    if (c < D) {
      t = x[(size_t)row * ldx + c];
      if (res) t = res[(size_t)row * ldres + c] + t * drop_factor(drop, (unsigned long long)row * D + c);
    }
    v[i] = t; s += t;
  }
  float mean = warp_sum(s) / D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int c = lane + 32 * i;
    float d = (c < D) ? v[i] - mean : 0.f;
    q += d * d;
  }
  float rstd = rsqrtf(warp_sum(q) / D + eps);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int c = lane + 32 * i;
    if (c < D) {
      y[(size_t)row * ldy + c] = (v[i] - mean) * rstd * w[c] + b[c];
      if (z) z[(size_t)row * ldz + c] = v[i];
    }
  }
  if (lane == 0) { stats[2 * row] = mean; stats[2 * row + 1] = rstd; }
}

// dz = rstd * (g - mean(g) - xhat * mean(g*xhat)),  g = dy*w;   dx = dz * drop;   dw += sum dy*xhat;  db += sum dy
template <int MAXV>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) layernorm_bwd_kernel(
    const float* __restrict__ dy, int lddy, const float* __restrict__ z, int ldz, const float* __restrict__ stats,
    const float* __restrict__ w, float* __restrict__ dz, int lddz, float* __restrict__ dx, int lddx, float* __restrict__ dw,
    float* __restrict__ db, int rows, int D, int rows_per_block, DropSpec drop) {
  __shared__ float s_dw[kWarpsPerBlock][32 * MAXV];
  __shared__ float s_db[kWarpsPerBlock][32 * MAXV];
  int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float aw[MAXV], ab[MAXV];
#pragma unroll
  for (int i = 0; i < MAXV; ++i) aw[i] = ab[i] = 0.f;
  int r0 = blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  for (int row = r0 + wid; row < r1; row += kWarpsPerBlock) {
    float mean = stats[2 * row], rstd = stats[2 * row + 1];
    float g[MAXV], xh[MAXV];
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      int c = lane + 32 * i;
      g[i] = xh[i] = 0.f;
      if (c < D) {
        float d = dy[(size_t)row * lddy + c];
        xh[i] = (z[(size_t)row * ldz + c] - mean) * rstd;
        g[i] = d * w[c];
        aw[i] += d * xh[i]; ab[i] += d;
        sg += g[i]; sgx += g[i] * xh[i];
      }
    }
    sg = warp_sum(sg) / D; sgx = warp_sum(sgx) / D;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      int c = lane + 32 * i;
      if (c < D) {
        float t = rstd * (g[i] - sg - xh[i] * sgx);
        if (dz) dz[(size_t)row * lddz + c] = t;
        if (dx) dx[(size_t)row * lddx + c] = t * drop_factor(drop, (unsigned long long)row * D + c);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < MAXV; ++i) { s_dw[wid][lane + 32 * i] = aw[i]; s_db[wid][lane + 32 * i] = ab[i]; }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float tw = 0.f, tb = 0.f;
#pragma unroll
    for (int k = 0; k < kWarpsPerBlock; ++k) { tw += s_dw[k][c]; tb += s_db[k][c]; }
    atomicAdd(&dw[c], tw); atomicAdd(&db[c], tb);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Spatial softmax over H*W positions of one (frame, channel) row.  out[row] = (E[x_map], E[y_map]), x_map follows the
// row index of the feature map, y_map the column index, both linspace(-1, 1).
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float lin_coord(int i, int n) { return n > 1 ? -1.0f + 2.0f * (float)i / (float)(n - 1) : -1.0f; }

template <int MAXV>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) spatial_softmax_fwd_kernel(const float* __restrict__ x, float* __restrict__ out,
                                                                              int rows, int H, int W, float inv_temp) {
  int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int P = H * W;
  const float* xr = x + (size_t)row * P;
  float v[MAXV];
  float mx = -FLT_MAX;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int p = lane + 32 * i;
    v[i] = (p < P) ? xr[p] * inv_temp : -FLT_MAX;
    mx = fmaxf(mx, v[i]);
  }
  mx = warp_max(mx);
  float s = 0.f, sx = 0.f, sy = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int p = lane + 32 * i;
    if (p < P) {
      float e = expf(v[i] - mx);
      s += e; sx += e * lin_coord(p / W, H); sy += e * lin_coord(p % W, W);
    }
  }
  s = warp_sum(s); sx = warp_sum(sx); sy = warp_sum(sy);
  if (lane == 0) { out[2 * (size_t)row] = sx / s; out[2 * (size_t)row + 1] = sy / s; }
}

// dx_p = a_p * inv_temp * ((gx*xm_p + gy*ym_p) - (gx*ex + gy*ey)), optionally gated by x_p > 0 (the ReLU that produced x)
template <int MAXV>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) spatial_softmax_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dout,
                                                                              float* __restrict__ dx, int rows, int H, int W, float inv_temp,
                                                                              int relu_gate) {
  int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int P = H * W;
  const float* xr = x + (size_t)row * P;
  float v[MAXV];
  float mx = -FLT_MAX;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int p = lane + 32 * i;
    v[i] = (p < P) ? xr[p] * inv_temp : -FLT_MAX;
    mx = fmaxf(mx, v[i]);
  }
  mx = warp_max(mx);
  float gx = dout[2 * (size_t)row], gy = dout[2 * (size_t)row + 1];
  float s = 0.f, sc = 0.f;
  float e[MAXV];
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int p = lane + 32 * i;
    e[i] = 0.f;
    if (p < P) {
      e[i] = expf(v[i] - mx);
      s += e[i]; sc += e[i] * (gx * lin_coord(p / W, H) + gy * lin_coord(p % W, W));
    }
  }
  s = warp_sum(s); sc = warp_sum(sc);
  float inv_s = 1.f / s, mean_c = sc * inv_s;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int p = lane + 32 * i;
    if (p < P) {
      float c = gx * lin_coord(p / W, H) + gy * lin_coord(p % W, W);
      float g = e[i] * inv_s * inv_temp * (c - mean_c);
      if (relu_gate && !(v[i] > 0.f)) g = 0.f;
      dx[(size_t)row * P + p] = g;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Spatial softmax on channels-last feature maps x [N, P = H*W, C] (the layout the tensor-core convolutions produce).
// One CTA per frame, 4 position lanes x 64 channels: every global access is a contiguous 256-byte run of channels;
// online softmax statistics per (position lane, channel), merged through shared memory.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kSsPosLanes = 4;

struct SsStat { float m, s, sx, sy; };

__device__ __forceinline__ void ss_merge(SsStat& a, const SsStat& b) {
  float m = fmaxf(a.m, b.m);
  float fa = expf(a.m - m), fb = expf(b.m - m);
  a.s = a.s * fa + b.s * fb; a.sx = a.sx * fa + b.sx * fb; a.sy = a.sy * fa + b.sy * fb; a.m = m;
}

// stats of channel c of frame n over positions pl, pl + kSsPosLanes, ...
__device__ __forceinline__ SsStat ss_partial(const float* __restrict__ xf, int C, int H, int W, int c, int pl, float inv_temp, float gx, float gy, bool weighted) {
  SsStat st{-FLT_MAX, 0.f, 0.f, 0.f};
  const int P = H * W;
  for (int p = pl; p < P; p += kSsPosLanes) {
    float v = xf[(size_t)p * C + c] * inv_temp;
    float m = fmaxf(st.m, v);
    float f = expf(st.m - m), e = expf(v - m);
    float cx = lin_coord(p / W, H), cy = lin_coord(p % W, W);
    st.s = st.s * f + e;
    if (weighted) { st.sx = st.sx * f + e * (gx * cx + gy * cy); }
    else { st.sx = st.sx * f + e * cx; st.sy = st.sy * f + e * cy; }
    st.m = m;
  }
  return st;
}

__global__ void __launch_bounds__(256) spatial_softmax_nhwc_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, int C, int H, int W, float inv_temp) {
  __shared__ SsStat sh[kSsPosLanes][64];
  const int n = blockIdx.x, P = H * W;
  const float* xf = x + (size_t)n * P * C;
  for (int c0 = 0; c0 < C; c0 += 64) {
    const int c = c0 + (threadIdx.x & 63), pl = threadIdx.x >> 6;
    if (c < C) sh[pl][threadIdx.x & 63] = ss_partial(xf, C, H, W, c, pl, inv_temp, 0.f, 0.f, false);
    __syncthreads();
    if (pl == 0 && c < C) {
      SsStat a = sh[0][threadIdx.x];
#pragma unroll
      for (int k = 1; k < kSsPosLanes; ++k) ss_merge(a, sh[k][threadIdx.x]);
      out[(size_t)n * 2 * C + 2 * c] = a.sx / a.s;
      out[(size_t)n * 2 * C + 2 * c + 1] = a.sy / a.s;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) spatial_softmax_nhwc_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dout, float* __restrict__ dx, int C,
                                                                       int H, int W, float inv_temp, int relu_gate) {
  __shared__ SsStat sh[kSsPosLanes][64];
  const int n = blockIdx.x, P = H * W;
  const float* xf = x + (size_t)n * P * C;
  float* df = dx + (size_t)n * P * C;
  for (int c0 = 0; c0 < C; c0 += 64) {
    const int lc = threadIdx.x & 63, c = c0 + lc, pl = threadIdx.x >> 6;
    float gx = 0.f, gy = 0.f;
    if (c < C) {
      gx = dout[(size_t)n * 2 * C + 2 * c]; gy = dout[(size_t)n * 2 * C + 2 * c + 1];
      sh[pl][lc] = ss_partial(xf, C, H, W, c, pl, inv_temp, gx, gy, true);
    }
    __syncthreads();
    if (c < C) {
      SsStat a = sh[0][lc];
#pragma unroll
      for (int k = 1; k < kSsPosLanes; ++k) ss_merge(a, sh[k][lc]);
      const float inv_s = 1.f / a.s, mean_c = a.sx * inv_s;  // sx holds sum e*(gx*xm + gy*ym)
      for (int p = pl; p < P; p += kSsPosLanes) {
        float xv = xf[(size_t)p * C + c];
        float v = xv * inv_temp;
        float cc = gx * lin_coord(p / W, H) + gy * lin_coord(p % W, W);
        float g = expf(v - a.m) * inv_s * inv_temp * (cc - mean_c);
        if (relu_gate && !(xv > 0.f)) g = 0.f;
        df[(size_t)p * C + c] = g;
      }
    }
    __syncthreads();
  }
}

// Register-resident variant for maps of up to PL*MAXV positions (the 21x21 map of the static camera: 16 x 28): a CTA of
// PL x 32 threads per (frame, 32 channels), thread (pl, c) keeps positions pl, pl+PL, ... of channel c in registers, so every load of the
// frame is issued before the first use (the serial online-softmax chain above is latency bound) and the backward pass reads
// x once.  Statistics of the PL position lanes meet in shared memory and are added in lane order.
// bf16 (round to nearest even) of a finite or infinite fp32; NaN stays NaN
__device__ __forceinline__ unsigned short f32_to_bf16_bits(float f) {
  unsigned u = __float_as_uint(f);
  if ((u & 0x7FFFFFFFu) > 0x7F800000u) return (unsigned short)((u >> 16) | 0x40u);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (unsigned short)(u >> 16);
}

// BF: the map x (and its gradient dx) are bf16 (the bf16 path's channels-last activations); statistics and outputs stay fp32.
template <int PL, int MAXV, bool BWD, bool BF = false>
__global__ void __launch_bounds__(PL * 32, 2) spatial_softmax_nhwc_reg_kernel(const void* __restrict__ x_, const float* __restrict__ dout, float* __restrict__ out,
                                                                           void* __restrict__ dx_, int C, int H, int W, float inv_temp, int relu_gate) {
  const float* x = reinterpret_cast<const float*>(x_);
  const unsigned short* x16 = reinterpret_cast<const unsigned short*>(x_);
  static_assert(MAXV <= 32, "the ReLU signs of a thread's values are kept in one 32-bit mask");
  __shared__ float sh_m[PL][32], sh_s[PL][32], sh_a[PL][32], sh_b[PL][32];
  __shared__ float cxs[PL * MAXV], cys[PL * MAXV];  // coordinate maps of the flattened positions (no per-element divisions)
  const int n = blockIdx.x, P = H * W;
  const int lc = threadIdx.x & 31, pl = threadIdx.x >> 5;
  const float* xf = x + (size_t)n * P * C;
  const unsigned short* xh = x16 + (size_t)n * P * C;
  for (int p = threadIdx.x; p < P; p += PL * 32) { cxs[p] = lin_coord(p / W, H); cys[p] = lin_coord(p % W, W); }
  const int c = blockIdx.y * 32 + lc;  // one CTA per (frame, group of 32 channels): two CTAs share an SM
  const bool cv = c < C;
  float v[MAXV];
  float mx = -FLT_MAX;
  unsigned pos_mask = 0u;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int p = pl + PL * i;
    v[i] = (cv && p < P) ? (BF ? __uint_as_float((unsigned)xh[(size_t)p * C + c] << 16) : xf[(size_t)p * C + c]) * inv_temp : -FLT_MAX;
    mx = fmaxf(mx, v[i]);
    pos_mask |= (v[i] > 0.f ? 1u : 0u) << i;  // inv_temp > 0: sign(v) == sign(x)
  }
  sh_m[pl][lc] = mx;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < PL; ++k) mx = fmaxf(mx, sh_m[k][lc]);
  float gx = 0.f, gy = 0.f;
  if (BWD && cv) { gx = dout[(size_t)n * 2 * C + 2 * c]; gy = dout[(size_t)n * 2 * C + 2 * c + 1]; }
  float s = 0.f, a = 0.f, b = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int p = pl + PL * i;
    if (p < P) {
      const float e = expf(v[i] - mx);
      v[i] = e;  // the values are only needed as exponentials from here on
      s += e;
      if (BWD) { a += e * (gx * cxs[p] + gy * cys[p]); }
      else { a += e * cxs[p]; b += e * cys[p]; }
    }
  }
  sh_s[pl][lc] = s; sh_a[pl][lc] = a;
  if (!BWD) sh_b[pl][lc] = b;
  __syncthreads();
  s = 0.f; a = 0.f; b = 0.f;
#pragma unroll
  for (int k = 0; k < PL; ++k) { s += sh_s[k][lc]; a += sh_a[k][lc]; if (!BWD) b += sh_b[k][lc]; }
  if (!BWD) {
    if (pl == 0 && cv) {
      out[(size_t)n * 2 * C + 2 * c] = a / s;
      out[(size_t)n * 2 * C + 2 * c + 1] = b / s;
    }
  } else if (cv) {
    const float scale = inv_temp / s, mean_c = a / s;
    float* df = reinterpret_cast<float*>(dx_) + (size_t)n * P * C;
    unsigned short* dh = reinterpret_cast<unsigned short*>(dx_) + (size_t)n * P * C;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int p = pl + PL * i;
      if (p < P) {
        float g = v[i] * scale * (gx * cxs[p] + gy * cys[p] - mean_c);
        if (relu_gate && !((pos_mask >> i) & 1u)) g = 0.f;
        if (BF) dh[(size_t)p * C + c] = f32_to_bf16_bits(g);
        else df[(size_t)p * C + c] = g;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// y[b,s,:] = drop(x[b,s,:] + pos[s,:])   and the generic dropout re-application used by its backward
// ---------------------------------------------------------------------------------------------------------------------
__global__ void add_posemb_kernel(const float* __restrict__ x, const float* __restrict__ pos, float* __restrict__ y, long long n, int S, int D,
                                  DropSpec drop) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int d = (int)(i % D), s = (int)((i / D) % S);
  y[i] = (x[i] + pos[(size_t)s * D + d]) * drop_factor(drop, (unsigned long long)i);
}
__global__ void dropout_apply_kernel(const float* __restrict__ x, float* __restrict__ y, long long n, DropSpec drop) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = x[i] * drop_factor(drop, (unsigned long long)i);
}

// dst[i0,i1,i2] (+)= alpha * src[i0,i1,i2] with arbitrary element strides (slices, transposes, broadcasts via stride 0)
__global__ void strided_copy_kernel(float* __restrict__ dst, const float* __restrict__ src, int n0, int n1, int n2, long long d0, long long d1,
                                    long long d2, long long s0, long long s1, long long s2, float alpha, int accumulate) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long n = (long long)n0 * n1 * n2;
  if (i >= n) return;
  int i2 = (int)(i % n2), i1 = (int)((i / n2) % n1), i0 = (int)(i / ((long long)n1 * n2));
  float v = alpha * src[i0 * s0 + i1 * s1 + i2 * s2];
  float* p = dst + i0 * d0 + i1 * d1 + i2 * d2;
  *p = accumulate ? *p + v : v;
}

// out[b,d] = scale * sum_s x[b,s,d]   (x contiguous [B,S,D])
__global__ void reduce_mid_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int S, int D, float scale) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * D) return;
  int b = i / D, d = i % D;
  const float* p = x + (size_t)b * S * D + d;
  float s = 0.f;
  for (int t = 0; t < S; ++t) s += p[(size_t)t * D];
  out[i] = s * scale;
}

// out[c] += sum_{n,p} x[n,c,p]  (NCHW bias gradient, accumulated); one block per (channel, chunk of n)
__global__ void __launch_bounds__(256) nchw_channel_sum_kernel(const float* __restrict__ x, float* __restrict__ out, int N, int C, int P, int n_per_block) {
  __shared__ float red[32];
  int c = blockIdx.x;
  int n0 = blockIdx.y * n_per_block, n1 = min(N, n0 + n_per_block);
  float s = 0.f;
  for (int n = n0; n < n1; ++n) {
    const float* p = x + ((size_t)n * C + c) * P;
    for (int i = threadIdx.x; i < P; i += blockDim.x) s += p[i];
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(&out[c], s);
}

// torch.optim.Adam (no weight decay, no amsgrad): one launch over the flat parameter buffer
// BF16: also write the updated parameters as bf16 into pb (the operand copy of the bf16 path: one pass instead of Adam + a cast)
template <bool BF16>
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, unsigned short* __restrict__ pb,
                            long long n, float lr, float b1, float b2, float eps, int step, const int* __restrict__ step_ptr, float grad_scale) {
  if (step_ptr) step = *step_ptr;  // device-resident step count: the launch can live in a replayed CUDA graph
  const float bc1 = 1.f - powf(b1, (float)step);
  const float bc2_sqrt = sqrtf(1.f - powf(b2, (float)step));
  const float step_size = lr / bc1;
  auto upd = [&](float& pi, float gi, float& mi, float& vi) {
    gi *= grad_scale;
    mi = b1 * mi + (1.f - b1) * gi;
    vi = b2 * vi + (1.f - b2) * gi * gi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi -= step_size * (mi / denom);
  };
  // 16-byte accesses, two independent quads per thread and trip (the kernel streams 7 x 4 bytes per parameter: HBM bound)
  const long long n4 = (((reinterpret_cast<size_t>(p) | reinterpret_cast<size_t>(g) | reinterpret_cast<size_t>(m) | reinterpret_cast<size_t>(v)) & 15) == 0 &&
                        (!BF16 || (reinterpret_cast<size_t>(pb) & 7) == 0)) ? n / 4 : 0;
  auto stb = [&](long long q, const float4& x) {  // four bf16 = 8 bytes
    if (BF16) {
      uint2 o;
      o.x = (unsigned)f32_to_bf16_bits(x.x) | ((unsigned)f32_to_bf16_bits(x.y) << 16);
      o.y = (unsigned)f32_to_bf16_bits(x.z) | ((unsigned)f32_to_bf16_bits(x.w) << 16);
      reinterpret_cast<uint2*>(pb)[q] = o;
    }
  };
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + stride < n4; i += 2 * stride) {
    float4 pa = p4[i], ga = g4[i], ma = m4[i], va = v4[i];
    float4 pb2 = p4[i + stride], gb = g4[i + stride], mb = m4[i + stride], vb = v4[i + stride];
    upd(pa.x, ga.x, ma.x, va.x); upd(pa.y, ga.y, ma.y, va.y); upd(pa.z, ga.z, ma.z, va.z); upd(pa.w, ga.w, ma.w, va.w);
    upd(pb2.x, gb.x, mb.x, vb.x); upd(pb2.y, gb.y, mb.y, vb.y); upd(pb2.z, gb.z, mb.z, vb.z); upd(pb2.w, gb.w, mb.w, vb.w);
    p4[i] = pa; m4[i] = ma; v4[i] = va;
    p4[i + stride] = pb2; m4[i + stride] = mb; v4[i + stride] = vb;
    stb(i, pa); stb(i + stride, pb2);
  }
  for (; i < n4; i += stride) {
    float4 pa = p4[i], ga = g4[i], ma = m4[i], va = v4[i];
    upd(pa.x, ga.x, ma.x, va.x); upd(pa.y, ga.y, ma.y, va.y); upd(pa.z, ga.z, ma.z, va.z); upd(pa.w, ga.w, ma.w, va.w);
    p4[i] = pa; m4[i] = ma; v4[i] = va;
    stb(i, pa);
  }
  for (long long e = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
    float pi = p[e], mi = m[e], vi = v[e];
    upd(pi, g[e], mi, vi);
    p[e] = pi; m[e] = mi; v[e] = vi;
    if (BF16) pb[e] = f32_to_bf16_bits(pi);
  }
}

__global__ void scale_kernel(float* __restrict__ x, long long n, float alpha) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] *= alpha;
}

// x *= *alpha_ptr; every block returns at once when the factor is exactly 1 (the common case: autograd's root gradient)
__global__ void scale_dev_kernel(float* __restrict__ x, long long n, const float* __restrict__ alpha_ptr) {
  const float alpha = *alpha_ptr;
  if (alpha == 1.0f) return;
  long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    float4 v = *reinterpret_cast<float4*>(x + i);
    v.x *= alpha; v.y *= alpha; v.z *= alpha; v.w *= alpha;
    *reinterpret_cast<float4*>(x + i) = v;
  } else {
    for (; i < n; ++i) x[i] *= alpha;
  }
}

}  // namespace

HULC_API int hulc_layernorm_fwd(const float* x, int ldx, const float* res, int ldres, const float* w, const float* b, float* y, int ldy,
                                float* z, int ldz, float* stats, int rows, int D, float eps, float drop_p, unsigned long long drop_seed,
                                unsigned drop_site, const unsigned char* drop_keep, void* stream) {
  if (rows <= 0) return 0;
  if (D > 128 || D <= 0) return (int)cudaErrorInvalidValue;
  DropSpec d = make_drop(drop_p, drop_seed, drop_site, drop_keep);
  dim3 grid(hulc_cdiv(rows, kWarpsPerBlock)), block(kWarpsPerBlock * 32);
  HULC_LAUNCH(layernorm_fwd_kernel<4>, grid, block, 0, (cudaStream_t)stream, x, ldx, res, ldres, w, b, y, ldy, z, ldz, stats, rows, D, eps, d);
  HULC_RETURN_LAST();
}

HULC_API int hulc_layernorm_bwd(const float* dy, int lddy, const float* z, int ldz, const float* stats, const float* w, float* dz, int lddz,
                                float* dx, int lddx, float* dw, float* db, int rows, int D, float drop_p, unsigned long long drop_seed,
                                unsigned drop_site, const unsigned char* drop_keep, void* stream) {
  if (rows <= 0) return 0;
  if (D > 128 || D <= 0) return (int)cudaErrorInvalidValue;
  DropSpec d = make_drop(drop_p, drop_seed, drop_site, drop_keep);
  int nblk = min(2 * kNumSMs, hulc_cdiv(rows, kWarpsPerBlock));
  int rpb = hulc_cdiv(rows, nblk);
  nblk = hulc_cdiv(rows, rpb);
  HULC_LAUNCH(layernorm_bwd_kernel<4>, dim3(nblk), dim3(kWarpsPerBlock * 32), 0, (cudaStream_t)stream, dy, lddy, z, ldz, stats, w, dz, lddz, dx,
              lddx, dw, db, rows, D, rpb, d);
  HULC_RETURN_LAST();
}

HULC_API int hulc_spatial_softmax_fwd(const float* x, float* out, int rows, int H, int W, float inv_temp, void* stream) {
  if (rows <= 0) return 0;
  if (H * W > 32 * 16) return (int)cudaErrorInvalidValue;
  HULC_LAUNCH(spatial_softmax_fwd_kernel<16>, dim3(hulc_cdiv(rows, kWarpsPerBlock)), dim3(kWarpsPerBlock * 32), 0, (cudaStream_t)stream, x, out, rows,
              H, W, inv_temp);
  HULC_RETURN_LAST();
}

HULC_API int hulc_spatial_softmax_bwd(const float* x, const float* dout, float* dx, int rows, int H, int W, float inv_temp, int relu_gate,
                                      void* stream) {
  if (rows <= 0) return 0;
  if (H * W > 32 * 16) return (int)cudaErrorInvalidValue;
  HULC_LAUNCH(spatial_softmax_bwd_kernel<16>, dim3(hulc_cdiv(rows, kWarpsPerBlock)), dim3(kWarpsPerBlock * 32), 0, (cudaStream_t)stream, x, dout, dx,
              rows, H, W, inv_temp, relu_gate);
  HULC_RETURN_LAST();
}

// Vectorised variant for the channels-last maps of the static camera (C = 64, up to 448 positions): ONE CTA per frame; a thread owns a
// 16-byte group of channels (8 bf16 / 4 fp32) at 14 positions, so a warp's load instruction covers whole positions back to back (512
// contiguous bytes) and all 14 loads of a thread are in flight before the first use — the frame is read from HBM exactly once, forward and
// backward (the backward keeps the frame in registers — packed, for bf16 — between the statistics and the gradient pass).  The 32
// position lanes of a channel meet by two shuffles inside a warp and a fixed-order sum over the warps in shared memory (deterministic).
// The scalar kernel above issues one 2- or 4-byte load per value and ran at 0.14 / 0.24 ms (forward / backward) per 2048 frames where the
// HBM traffic allows 0.02 / 0.04; this one is bound by its instruction count (an exponential per element and pass), hence the fast exp.
template <bool BF>
struct SsVec {
  static constexpr int kVec = BF ? 8 : 4;          // channels per thread = one 16-byte load
  static constexpr int kC = 64;
  static constexpr int kGroups = kC / kVec;        // 8 / 16 threads per position
  static constexpr int kLanes = 32;                // position lanes: 256 / 512 threads (64 lanes for bf16 — 512 threads, 7 positions each — measured
                                                   // 0.124 / 0.224 ms against 0.067 / 0.147: one 16-warp CTA per SM hides less than two 8-warp ones)
  static constexpr int kThreads = kLanes * kGroups;
  static constexpr int kWarps = kThreads / 32;
  static constexpr int kPosPerWarp = 32 / kGroups; // 4 / 2
  static constexpr int kNPos = 448 / kLanes;       // positions per thread (7 / 14): up to 448 per frame
};

template <bool BF>
__device__ __forceinline__ void ss_unpack(const uint4& u, float* v) {
  if (BF) {
    v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xFFFF0000u);
    v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xFFFF0000u);
    v[4] = __uint_as_float(u.z << 16); v[5] = __uint_as_float(u.z & 0xFFFF0000u);
    v[6] = __uint_as_float(u.w << 16); v[7] = __uint_as_float(u.w & 0xFFFF0000u);
  } else {
    v[0] = __uint_as_float(u.x); v[1] = __uint_as_float(u.y); v[2] = __uint_as_float(u.z); v[3] = __uint_as_float(u.w);
  }
}

template <bool BF, bool BWD>
__global__ void __launch_bounds__(SsVec<BF>::kThreads, BF ? 2 : 1) spatial_softmax_vec_kernel(const void* __restrict__ x_, const float* __restrict__ dout,
                                                                                           float* __restrict__ out, void* __restrict__ dx_, int H, int W,
                                                                                           float inv_temp, int relu_gate) {
  using K = SsVec<BF>;
  constexpr int V = K::kVec, NP = K::kNPos, C = K::kC, LN = K::kLanes;
  __shared__ float sh[3][K::kWarps][C];
  __shared__ float cxs[LN * NP], cys[LN * NP];  // coordinate maps of the flattened positions (no per-element divisions)
  const int n = blockIdx.x, P = H * W;
  const int g = threadIdx.x % K::kGroups, pl = threadIdx.x / K::kGroups;  // channel group, position lane
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = g * V;
  const unsigned char* xb = reinterpret_cast<const unsigned char*>(x_) + (size_t)n * P * C * (BF ? 2 : 4) + (size_t)g * 16;
  for (int p = threadIdx.x; p < P; p += K::kThreads) { cxs[p] = lin_coord(p / W, H); cys[p] = lin_coord(p % W, W); }
  uint4 u[NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    const int p = pl + LN * i;
    u[i] = make_uint4(0u, 0u, 0u, 0u);
    if (p < P) u[i] = __ldg(reinterpret_cast<const uint4*>(xb + (size_t)p * C * (BF ? 2 : 4)));
  }
  // ---- per-channel maximum over the frame ----
  float mx[V];
#pragma unroll
  for (int j = 0; j < V; ++j) mx[j] = -FLT_MAX;
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    if (pl + LN * i < P) {
      float v[V];
      ss_unpack<BF>(u[i], v);
#pragma unroll
      for (int j = 0; j < V; ++j) mx[j] = fmaxf(mx[j], v[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < V; ++j) {
#pragma unroll
    for (int m = K::kGroups; m < 32; m <<= 1) mx[j] = fmaxf(mx[j], __shfl_xor_sync(0xffffffffu, mx[j], m));
  }
  if (lane < K::kGroups) {
#pragma unroll
    for (int j = 0; j < V; ++j) sh[0][warp][c0 + j] = mx[j];
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < V; ++j) {
    float m = sh[0][0][c0 + j];
    for (int k = 1; k < K::kWarps; ++k) m = fmaxf(m, sh[0][k][c0 + j]);
    mx[j] = m * inv_temp;  // inv_temp > 0: max(x * inv_temp) = max(x) * inv_temp
  }
  float gx[V], gy[V];
  if (BWD) {
#pragma unroll
    for (int j = 0; j < V; ++j) { gx[j] = dout[(size_t)n * 2 * C + 2 * (c0 + j)]; gy[j] = dout[(size_t)n * 2 * C + 2 * (c0 + j) + 1]; }
  }
  // ---- sums of the exponentials and of the coordinate-weighted exponentials ----
  float s[V], a[V], b[V];
#pragma unroll
  for (int j = 0; j < V; ++j) s[j] = a[j] = b[j] = 0.f;
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    const int p = pl + LN * i;
    if (p < P) {
      float v[V];
      ss_unpack<BF>(u[i], v);
      const float cx = cxs[p], cy = cys[p];
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const float e = __expf(v[j] * inv_temp - mx[j]);  // ex2.approx: 2^-22 relative, against ~10 instructions of expf per element
        s[j] += e;
        if (BWD) a[j] += e * (gx[j] * cx + gy[j] * cy);
        else { a[j] += e * cx; b[j] += e * cy; }
      }
    }
  }
  __syncthreads();  // (every thread has read the maxima)
#pragma unroll
  for (int j = 0; j < V; ++j) {
#pragma unroll
    for (int m = K::kGroups; m < 32; m <<= 1) {
      s[j] += __shfl_xor_sync(0xffffffffu, s[j], m);
      a[j] += __shfl_xor_sync(0xffffffffu, a[j], m);
      if (!BWD) b[j] += __shfl_xor_sync(0xffffffffu, b[j], m);
    }
  }
  if (lane < K::kGroups) {
#pragma unroll
    for (int j = 0; j < V; ++j) { sh[0][warp][c0 + j] = s[j]; sh[1][warp][c0 + j] = a[j]; if (!BWD) sh[2][warp][c0 + j] = b[j]; }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < V; ++j) {
    float ss = 0.f, aa = 0.f, bb = 0.f;
    for (int k = 0; k < K::kWarps; ++k) { ss += sh[0][k][c0 + j]; aa += sh[1][k][c0 + j]; if (!BWD) bb += sh[2][k][c0 + j]; }
    s[j] = ss; a[j] = aa; b[j] = bb;
  }
  if (!BWD) {
    if (pl == 0) {
#pragma unroll
      for (int j = 0; j < V; ++j) {
        out[(size_t)n * 2 * C + 2 * (c0 + j)] = a[j] / s[j];
        out[(size_t)n * 2 * C + 2 * (c0 + j) + 1] = b[j] / s[j];
      }
    }
    return;
  }
  // ---- gradient w.r.t. the map (gated by the ReLU that produced it) ----
  unsigned char* db = reinterpret_cast<unsigned char*>(dx_) + (size_t)n * P * C * (BF ? 2 : 4) + (size_t)g * 16;
  float scale[V], mean_c[V];
#pragma unroll
  for (int j = 0; j < V; ++j) { scale[j] = inv_temp / s[j]; mean_c[j] = a[j] / s[j]; }
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    const int p = pl + LN * i;
    if (p < P) {
      float v[V], gr[V];
      ss_unpack<BF>(u[i], v);
      const float cx = cxs[p], cy = cys[p];
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const float e = __expf(v[j] * inv_temp - mx[j]);  // ex2.approx: 2^-22 relative, against ~10 instructions of expf per element
        gr[j] = e * scale[j] * (gx[j] * cx + gy[j] * cy - mean_c[j]);
        if (relu_gate && !(v[j] > 0.f)) gr[j] = 0.f;
      }
      uint4 o;
      if (BF) {
        o.x = (unsigned)f32_to_bf16_bits(gr[0]) | ((unsigned)f32_to_bf16_bits(gr[1]) << 16);
        o.y = (unsigned)f32_to_bf16_bits(gr[2]) | ((unsigned)f32_to_bf16_bits(gr[3]) << 16);
        o.z = (unsigned)f32_to_bf16_bits(gr[4 % V]) | ((unsigned)f32_to_bf16_bits(gr[5 % V]) << 16);
        o.w = (unsigned)f32_to_bf16_bits(gr[6 % V]) | ((unsigned)f32_to_bf16_bits(gr[7 % V]) << 16);
      } else {
        o = make_uint4(__float_as_uint(gr[0]), __float_as_uint(gr[1]), __float_as_uint(gr[2]), __float_as_uint(gr[3]));
      }
      *reinterpret_cast<uint4*>(db + (size_t)p * C * (BF ? 2 : 4)) = o;
    }
  }
}

// the vectorised kernel takes 64-channel maps of up to 448 positions with 16-byte aligned rows
static bool ss_vec_ok(const void* x, const void* dx, int C, int H, int W, float inv_temp) {
  return C == 64 && H * W <= 448 && inv_temp > 0.f && ((reinterpret_cast<size_t>(x) | reinterpret_cast<size_t>(dx)) & 15) == 0;
}

HULC_API int hulc_spatial_softmax_nhwc_fwd(const float* x, float* out, int N, int C, int H, int W, float inv_temp, void* stream) {
  if (N <= 0) return 0;
  if (ss_vec_ok(x, nullptr, C, H, W, inv_temp)) {
    HULC_LAUNCH((spatial_softmax_vec_kernel<false, false>), dim3(N), dim3(SsVec<false>::kThreads), 0, (cudaStream_t)stream, (const void*)x, (const float*)nullptr, out,
                (void*)nullptr, H, W, inv_temp, 0);
    HULC_RETURN_LAST();
  }
  if (H * W <= 16 * 28 && inv_temp > 0.f) {
    HULC_LAUNCH((spatial_softmax_nhwc_reg_kernel<16, 28, false>), dim3(N, hulc_cdiv(C, 32)), dim3(512), 0, (cudaStream_t)stream, (const void*)x, (const float*)nullptr, out,
                (void*)nullptr, C, H, W, inv_temp, 0);
    HULC_RETURN_LAST();
  }
  HULC_LAUNCH(spatial_softmax_nhwc_fwd_kernel, dim3(N), dim3(256), 0, (cudaStream_t)stream, x, out, C, H, W, inv_temp);
  HULC_RETURN_LAST();
}

HULC_API int hulc_spatial_softmax_nhwc_bwd(const float* x, const float* dout, float* dx, int N, int C, int H, int W, float inv_temp, int relu_gate,
                                           void* stream) {
  if (N <= 0) return 0;
  if (ss_vec_ok(x, dx, C, H, W, inv_temp)) {
    HULC_LAUNCH((spatial_softmax_vec_kernel<false, true>), dim3(N), dim3(SsVec<false>::kThreads), 0, (cudaStream_t)stream, (const void*)x, dout, (float*)nullptr, (void*)dx,
                H, W, inv_temp, relu_gate);
    HULC_RETURN_LAST();
  }
  if (H * W <= 16 * 28 && inv_temp > 0.f) {
    HULC_LAUNCH((spatial_softmax_nhwc_reg_kernel<16, 28, true>), dim3(N, hulc_cdiv(C, 32)), dim3(512), 0, (cudaStream_t)stream, (const void*)x, dout, (float*)nullptr, (void*)dx, C, H, W,
                inv_temp, relu_gate);
    HULC_RETURN_LAST();
  }
  HULC_LAUNCH(spatial_softmax_nhwc_bwd_kernel, dim3(N), dim3(256), 0, (cudaStream_t)stream, x, dout, dx, C, H, W, inv_temp, relu_gate);
  HULC_RETURN_LAST();
}

// bf16 maps (x, dx channels-last bf16); maps of up to 448 positions (the 21x21 map of the static camera)
HULC_API int hulc_spatial_softmax_nhwc_bf16_fwd(const void* x, float* out, int N, int C, int H, int W, float inv_temp, void* stream) {
  if (N <= 0) return 0;
  if (H * W > 16 * 28 || !(inv_temp > 0.f)) return (int)cudaErrorInvalidValue;
  if (ss_vec_ok(x, nullptr, C, H, W, inv_temp)) {
    HULC_LAUNCH((spatial_softmax_vec_kernel<true, false>), dim3(N), dim3(SsVec<true>::kThreads), 0, (cudaStream_t)stream, x, (const float*)nullptr, out, (void*)nullptr, H, W,
                inv_temp, 0);
    HULC_RETURN_LAST();
  }
  HULC_LAUNCH((spatial_softmax_nhwc_reg_kernel<16, 28, false, true>), dim3(N, hulc_cdiv(C, 32)), dim3(512), 0, (cudaStream_t)stream, x, (const float*)nullptr, out,
              (void*)nullptr, C, H, W, inv_temp, 0);
  HULC_RETURN_LAST();
}
HULC_API int hulc_spatial_softmax_nhwc_bf16_bwd(const void* x, const float* dout, void* dx, int N, int C, int H, int W, float inv_temp, int relu_gate, void* stream) {
  if (N <= 0) return 0;
  if (H * W > 16 * 28 || !(inv_temp > 0.f)) return (int)cudaErrorInvalidValue;
  if (ss_vec_ok(x, dx, C, H, W, inv_temp)) {
    HULC_LAUNCH((spatial_softmax_vec_kernel<true, true>), dim3(N), dim3(SsVec<true>::kThreads), 0, (cudaStream_t)stream, x, dout, (float*)nullptr, dx, H, W, inv_temp,
                relu_gate);
    HULC_RETURN_LAST();
  }
  HULC_LAUNCH((spatial_softmax_nhwc_reg_kernel<16, 28, true, true>), dim3(N, hulc_cdiv(C, 32)), dim3(512), 0, (cudaStream_t)stream, x, dout, (float*)nullptr, dx, C, H, W,
              inv_temp, relu_gate);
  HULC_RETURN_LAST();
}

HULC_API int hulc_add_posemb_fwd(const float* x, const float* pos, float* y, int B, int S, int D, float drop_p, unsigned long long drop_seed,
                                 unsigned drop_site, const unsigned char* drop_keep, void* stream) {
  long long n = (long long)B * S * D;
  if (n <= 0) return 0;
  HULC_LAUNCH(add_posemb_kernel, dim3(hulc_cdiv(n, 256)), dim3(256), 0, (cudaStream_t)stream, x, pos, y, n, S, D,
              make_drop(drop_p, drop_seed, drop_site, drop_keep));
  HULC_RETURN_LAST();
}

HULC_API int hulc_dropout_apply(const float* x, float* y, long long n, float drop_p, unsigned long long drop_seed, unsigned drop_site,
                                const unsigned char* drop_keep, void* stream) {
  if (n <= 0) return 0;
  HULC_LAUNCH(dropout_apply_kernel, dim3(hulc_cdiv(n, 256)), dim3(256), 0, (cudaStream_t)stream, x, y, n,
              make_drop(drop_p, drop_seed, drop_site, drop_keep));
  HULC_RETURN_LAST();
}

HULC_API int hulc_strided_copy(float* dst, const float* src, int n0, int n1, int n2, long long d0, long long d1, long long d2, long long s0,
                               long long s1, long long s2, float alpha, int accumulate, void* stream) {
  long long n = (long long)n0 * n1 * n2;
  if (n <= 0) return 0;
  HULC_LAUNCH(strided_copy_kernel, dim3(hulc_cdiv(n, 256)), dim3(256), 0, (cudaStream_t)stream, dst, src, n0, n1, n2, d0, d1, d2, s0, s1, s2, alpha,
              accumulate);
  HULC_RETURN_LAST();
}

HULC_API int hulc_reduce_mid(const float* x, float* out, int B, int S, int D, float scale, void* stream) {
  if (B * D <= 0) return 0;
  HULC_LAUNCH(reduce_mid_kernel, dim3(hulc_cdiv((long long)B * D, 128)), dim3(128), 0, (cudaStream_t)stream, x, out, B, S, D, scale);
  HULC_RETURN_LAST();
}

HULC_API int hulc_nchw_channel_sum(const float* x, float* out, int N, int C, int P, void* stream) {
  if (N <= 0 || C <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  int gy = max(1, min(N, (4 * kNumSMs) / C));
  int npb = hulc_cdiv(N, gy);
  gy = hulc_cdiv(N, npb);
  HULC_LAUNCH(nchw_channel_sum_kernel, dim3(C, gy), dim3(256), 0, st, x, out, N, C, P, npb);
  HULC_RETURN_LAST();
}

HULC_API int hulc_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps, int step,
                            const int* step_ptr, float grad_scale, void* stream) {
  if (n <= 0) return 0;
  int blocks = (int)min((long long)kNumSMs * 8, (n + 255) / 256);
  HULC_LAUNCH(adam_kernel<false>, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, p, g, m, v, (unsigned short*)nullptr, n, lr, beta1, beta2, eps, step, step_ptr, grad_scale);
  HULC_RETURN_LAST();
}

HULC_API int hulc_adam_step_bf16(float* p, const float* g, float* m, float* v, void* p_bf16, long long n, float lr, float beta1, float beta2, float eps, int step,
                                 const int* step_ptr, float grad_scale, void* stream) {
  if (n <= 0) return 0;
  if (!p_bf16) return (int)cudaErrorInvalidValue;
  int blocks = (int)min((long long)kNumSMs * 8, (n + 255) / 256);
  HULC_LAUNCH(adam_kernel<true>, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, p, g, m, v, reinterpret_cast<unsigned short*>(p_bf16), n, lr, beta1, beta2, eps, step,
              step_ptr, grad_scale);
  HULC_RETURN_LAST();
}

namespace {
// uint8 camera frames -> the normalised fp32 the encoders take: ((x / 255) - mean) / std, evaluated in that order with IEEE
// fp32 operations so it matches the reference's CPU transforms bit for bit.  16 pixels per thread (one 16-byte load).
__global__ void frames_u8_kernel(const uint4* __restrict__ src, float4* __restrict__ dst, long long n16, float mean, float stdv) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n16) return;
  const uint4 v = src[i];
  const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float o[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) o[b] = (((float)((w[q] >> (8 * b)) & 0xFFu) / 255.0f) - mean) / stdv;
    dst[i * 4 + q] = make_float4(o[0], o[1], o[2], o[3]);
  }
}
__global__ void frames_u8_tail_kernel(const unsigned char* __restrict__ src, float* __restrict__ dst, long long n0, long long n, float mean, float stdv) {
  long long i = n0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (((float)src[i] / 255.0f) - mean) / stdv;
}
// RandomShiftsAug (hulc/utils/transforms.py:8-29) + scale + normalise in one pass: the reference pads the frame by `pad` replicated pixels and
// samples it on a grid displaced by an integer number of pixels (sx, sy) in [0, 2 pad] drawn per frame — i.e. out[y][x] = in[clamp(y + sy - pad)]
// [clamp(x + sx - pad)].  shifts: [N][2] = (sx, sy) injected, or NULL: drawn from Philox(seed, site, frame).  4 pixels of a row per thread.
__global__ void frames_u8_shift_kernel(const unsigned char* __restrict__ src, float* __restrict__ dst, int N, int C, int H, int W, int pad,
                                       const int* __restrict__ shifts, unsigned long long seed, const unsigned long long* seed_ptr, unsigned site, float mean,
                                       float stdv) {
  const int Wq = (W + 3) / 4;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * C * H * Wq) return;
  const int xq = (int)(i % Wq);
  const long long r = i / Wq;
  const int y = (int)(r % H);
  const long long nc = r / H;
  const int n = (int)(nc / C);
  int sx, sy;
  if (shifts) {
    sx = shifts[2 * n]; sy = shifts[2 * n + 1];
  } else {  // torch.randint(0, 2 pad + 1): uniform integers, two per frame
    const uint4 rnd = philox4x32(rng_seed(seed, seed_ptr), site, (unsigned long long)n);
    sx = (int)(rnd.x % (unsigned)(2 * pad + 1)); sy = (int)(rnd.y % (unsigned)(2 * pad + 1));
  }
  const int cy = min(max(y + sy - pad, 0), H - 1);
  const unsigned char* row = src + ((size_t)nc * H + cy) * W;
  float* out = dst + ((size_t)nc * H + y) * W + xq * 4;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int x = xq * 4 + e;
    if (x < W) out[e] = (((float)row[min(max(x + sx - pad, 0), W - 1)] / 255.0f) - mean) / stdv;
  }
}

}  // namespace

HULC_API int hulc_frames_u8_to_f32(const unsigned char* src, float* dst, long long n, float mean, float stdv, void* stream) {
  if (n <= 0) return 0;
  if (!src || !dst || stdv == 0.f) return (int)cudaErrorInvalidValue;
  cudaStream_t st = (cudaStream_t)stream;
  long long n16 = ((reinterpret_cast<size_t>(src) | reinterpret_cast<size_t>(dst)) & 15) ? 0 : n / 16;
  if (n16 > 0) HULC_LAUNCH(frames_u8_kernel, dim3(hulc_cdiv(n16, 256)), dim3(256), 0, st, reinterpret_cast<const uint4*>(src), reinterpret_cast<float4*>(dst), n16, mean, stdv);
  if (n16 * 16 < n) HULC_LAUNCH(frames_u8_tail_kernel, dim3(hulc_cdiv(n - n16 * 16, 256)), dim3(256), 0, st, src, dst, n16 * 16, n, mean, stdv);
  HULC_RETURN_LAST();
}

HULC_API int hulc_frames_u8_shift_to_f32(const unsigned char* src, float* dst, int N, int C, int H, int W, int pad, const int* shifts, unsigned long long seed,
                                         unsigned site, float mean, float stdv, void* stream) {
  if (N <= 0) return 0;
  if (!src || !dst || stdv == 0.f || pad < 0) return (int)cudaErrorInvalidValue;
  const long long n = (long long)N * C * H * ((W + 3) / 4);
  HULC_LAUNCH(frames_u8_shift_kernel, dim3(hulc_cdiv(n, 256)), dim3(256), 0, (cudaStream_t)stream, src, dst, N, C, H, W, pad, shifts, seed, g_hulc_rng_offset_ptr, site, mean,
              stdv);
  HULC_RETURN_LAST();
}

HULC_API int hulc_scale(float* x, long long n, float alpha, void* stream) {
  if (n <= 0) return 0;
  HULC_LAUNCH(scale_kernel, dim3(hulc_cdiv(n, 256)), dim3(256), 0, (cudaStream_t)stream, x, n, alpha);
  HULC_RETURN_LAST();
}

HULC_API int hulc_scale_dev(float* x, long long n, const float* alpha_ptr, void* stream) {
  if (n <= 0) return 0;
  if ((reinterpret_cast<size_t>(x) & 15) != 0 || !alpha_ptr) return (int)cudaErrorInvalidValue;
  HULC_LAUNCH(scale_dev_kernel, dim3(hulc_cdiv(hulc_cdiv(n, 4), 256)), dim3(256), 0, (cudaStream_t)stream, x, n, alpha_ptr);
  HULC_RETURN_LAST();
}
