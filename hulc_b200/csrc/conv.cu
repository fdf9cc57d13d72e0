// conv.cu — the three valid (no padding) NCHW convolutions of the perceptual encoders (vision_network.py:36-47,
// vision_network_gripper.py:11-17) as implicit GEMMs: forward (+bias +ReLU), data gradient (gather form, per stride
// phase, gated by the producer's ReLU) and weight gradient (split over output positions, fixed-order reduction).
// Exact-fp32 CUDA-core path; im2col is never materialised: every operand element is gathered straight from the
// activation tensor with a per-row base offset and a per-k offset.
#include "common.cuh"

namespace {

constexpr int BK = 16;
constexpr int NT = 256;

struct ConvDims {
  int N, CIN, H, W, COUT, HO, WO;
};

// ---------------------------------------------------------------------------------------------------------------------
// Gather-GEMM:  dst[row(m), c] = epi( sum_k src[base(m) + off(k)] * Bmat[k][c] ),  c < NOUT
// ---------------------------------------------------------------------------------------------------------------------
template <int CIN, int KS, int S>
struct FwdGeo {
  static constexpr int K = CIN * KS * KS;
  ConvDims d;
  __device__ __forceinline__ int rows() const { return d.N * d.HO * d.WO; }
  // row decode -> (source base offset, y, x, destination base offset); destination channel stride returned by cstride()
  __device__ __forceinline__ void row(int m, long long& sbase, int& y, int& x, long long& dbase) const {
    int P = d.HO * d.WO;
    int n = m / P, p = m % P;
    y = p / d.WO; x = p % d.WO;
    sbase = (long long)n * CIN * d.H * d.W + (long long)y * S * d.W + x * S;
    dbase = (long long)n * d.COUT * P + p;
  }
  __device__ __forceinline__ long long cstride() const { return (long long)d.HO * d.WO; }
  __device__ __forceinline__ bool tap(int k, int, int, int& off) const {
    int ci = k / (KS * KS), r = k % (KS * KS);
    off = ci * d.H * d.W + (r / KS) * d.W + (r % KS);
    return true;
  }
};

// Data gradient of a stride-S, kernel-KS conv, restricted to the input pixels of one stride phase (py, px):
// dX[n,ci,S*y2+py,S*x2+px] = sum_{co,jy,jx} dY[n,co,y2-jy,x2-jx] * W[co,ci,py+S*jy,px+S*jx]
template <int COUT, int KS, int S>
struct DgradGeo {
  static constexpr int R = KS / S;  // taps per dimension that hit one phase (KS % S == 0 for the three layers)
  static constexpr int K = COUT * R * R;
  ConvDims d;
  int py, px, HP, WP;
  __device__ __forceinline__ int rows() const { return d.N * HP * WP; }
  __device__ __forceinline__ void row(int m, long long& sbase, int& y, int& x, long long& dbase) const {
    int P = HP * WP;
    int n = m / P, p = m % P;
    y = p / WP; x = p % WP;
    sbase = (long long)n * COUT * d.HO * d.WO + (long long)y * d.WO + x;
    dbase = (long long)n * d.CIN * d.H * d.W + (long long)(S * y + py) * d.W + (S * x + px);
  }
  __device__ __forceinline__ long long cstride() const { return (long long)d.H * d.W; }
  __device__ __forceinline__ bool tap(int k, int y, int x, int& off) const {
    int co = k / (R * R), r = k % (R * R);
    int jy = r / R, jx = r % R;
    off = co * d.HO * d.WO - jy * d.WO - jx;
    int oy = y - jy, ox = x - jx;
    return oy >= 0 && oy < d.HO && ox >= 0 && ox < d.WO;
  }
};

template <class Geo, int NOUT, int BM>
__global__ void __launch_bounds__(NT) igemm_kernel(Geo geo, const float* __restrict__ src, const float* __restrict__ Bmat, const float* __restrict__ bias,
                                                  const float* __restrict__ gate, float* __restrict__ dst, int relu) {
  constexpr int K = Geo::K;
  static_assert(K % BK == 0, "K must be a multiple of BK");
  constexpr int NGRP = NOUT / 4, MGRP = NT / NGRP, TM = BM / MGRP;
  static_assert(TM == 8, "thread tile is 8 rows x 4 channels");
  constexpr int A_PER = BM * BK / NT;  // A elements gathered per thread per k-step
  constexpr int KSTEP = NT / BM;       // threads sharing one row split the BK columns
  constexpr int LDA = BM + 4, LDB = NOUT + 4;
  constexpr int SMEM_AB = 2 * BK * (LDA + LDB);
  constexpr int SMEM_C = NOUT * (BM + 1);
  constexpr int SMEM = SMEM_AB > SMEM_C ? SMEM_AB : SMEM_C;
  __shared__ __align__(16) float sm[SMEM];
  __shared__ long long s_dbase[BM];
  float* As = sm;                 // [2][BK][LDA]
  float* Bs = sm + 2 * BK * LDA;  // [2][BK][LDB]

  const int tid = threadIdx.x;
  const int M = geo.rows();
  const int m0 = blockIdx.x * BM;
  if (m0 >= M) return;
  const int tn = tid % NGRP, tm = tid / NGRP;

  // this thread gathers row (tid % BM), columns (tid / BM) + KSTEP*j
  const int lrow = tid % BM, lk0 = tid / BM;
  long long sbase = 0, dbase = 0;
  int ry = 0, rx = 0;
  const bool rvalid = (m0 + lrow) < M;
  if (rvalid) geo.row(m0 + lrow, sbase, ry, rx, dbase);
  if (tid < BM) s_dbase[tid] = rvalid ? dbase : -1;

  float ra[A_PER];
  float4 rb;
  constexpr int B_V = BK * NOUT / 4;  // float4 in a B tile (<= NT)
  auto load = [&](int k0) {
#pragma unroll
    for (int j = 0; j < A_PER; ++j) {
      int k = k0 + lk0 + KSTEP * j, off;
      bool ok = geo.tap(k, ry, rx, off) && rvalid;
      ra[j] = ok ? src[sbase + off] : 0.f;
    }
    if (tid < B_V) rb = *reinterpret_cast<const float4*>(Bmat + (size_t)(k0 + tid / (NOUT / 4)) * NOUT + (tid % (NOUT / 4)) * 4);
  };
  auto store = [&](int buf) {
#pragma unroll
    for (int j = 0; j < A_PER; ++j) As[(buf * BK + lk0 + KSTEP * j) * LDA + lrow] = ra[j];
    if (tid < B_V) *reinterpret_cast<float4*>(&Bs[(buf * BK + tid / (NOUT / 4)) * LDB + (tid % (NOUT / 4)) * 4]) = rb;
  };

  float acc[TM][4];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  load(0);
  store(0);
  __syncthreads();
  constexpr int NKI = K / BK;
  for (int it = 0; it < NKI; ++it) {
    const int buf = it & 1;
    if (it + 1 < NKI) load((it + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float* ap = &As[(buf * BK + k) * LDA];
      float4 a0 = *reinterpret_cast<const float4*>(ap + tm * 4);
      float4 a1 = *reinterpret_cast<const float4*>(ap + BM / 2 + tm * 4);
      float4 b = *reinterpret_cast<const float4*>(&Bs[(buf * BK + k) * LDB + tn * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    if (it + 1 < NKI) store(buf ^ 1);
    __syncthreads();
  }
  // stage the tile as Cs[c][m] so the global stores run along m (contiguous pixels of one channel plane)
  float* Cs = sm;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int ml = (i / 4) * (BM / 2) + tm * 4 + (i % 4);
#pragma unroll
    for (int j = 0; j < 4; ++j) Cs[(tn * 4 + j) * (BM + 1) + ml] = acc[i][j];
  }
  __syncthreads();
  const long long cs = geo.cstride();
  for (int e = tid; e < NOUT * BM; e += NT) {
    int c = e / BM, ml = e % BM;
    long long db = s_dbase[ml];
    if (db < 0) continue;
    float v = Cs[c * (BM + 1) + ml];
    if (bias) v += bias[c];
    if (relu) v = fmaxf(v, 0.f);
    long long o = db + c * cs;
    if (gate) v = gate[o] > 0.f ? v : 0.f;
    dst[o] = v;
  }
}

// w[COUT][CIN*KS*KS] -> wt[K][COUT]
__global__ void weight_fwd_prep_kernel(const float* __restrict__ w, float* __restrict__ wt, int COUT, int K) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= COUT * K) return;
  int k = i / COUT, co = i % COUT;
  wt[i] = w[(size_t)co * K + k];
}
// w[COUT][CIN][KS][KS] -> wp[phase = py*S+px][k' = (co,jy,jx)][ci] with ky = py + S*jy, kx = px + S*jx
__global__ void weight_dgrad_prep_kernel(const float* __restrict__ w, float* __restrict__ wp, int COUT, int CIN, int KS, int S) {
  int R = KS / S, Kp = COUT * R * R;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S * S * Kp * CIN) return;
  int ci = i % CIN, k = (i / CIN) % Kp, ph = i / (CIN * Kp);
  int py = ph / S, px = ph % S;
  int co = k / (R * R), r = k % (R * R), jy = r / R, jx = r % R;
  wp[i] = w[(((size_t)co * CIN + ci) * KS + (py + S * jy)) * KS + (px + S * jx)];
}

// ---------------------------------------------------------------------------------------------------------------------
// Weight gradient: dW[co][k] = sum_m dY[m][co] * X[m][k]; CTA = (chunk of rows, tile of k); partials -> fixed-order sum.
// ---------------------------------------------------------------------------------------------------------------------
template <int CIN, int KS, int S, int COUT>
__global__ void __launch_bounds__(NT) wgrad_kernel(ConvDims d, const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ partial,
                                                  int rows_per_chunk) {
  constexpr int K = CIN * KS * KS;
  constexpr int CG = COUT / 4, KG = NT / CG, KT = KG * 8;  // thread tile 4 (co) x 8 (k)
  constexpr int BR = 16;
  constexpr int LDD = COUT + 4, LDX = KT + 4;
  __shared__ __align__(16) float Ds[2][BR][LDD];
  __shared__ __align__(16) float Xs[2][BR][LDX];
  __shared__ long long s_xbase[2][BR];
  __shared__ long long s_ybase[2][BR];
  const int tid = threadIdx.x;
  const int cg = tid % CG, kg = tid / CG;
  const int P = d.HO * d.WO, M = d.N * P;
  const int kt0 = blockIdx.y * KT;
  const int mbeg = blockIdx.x * rows_per_chunk, mend = min(M, mbeg + rows_per_chunk);

  constexpr int X_PER = BR * KT / NT, D_PER = (BR * COUT + NT - 1) / NT;
  float rx[X_PER], rd[D_PER];
  auto rowinfo = [&](int buf, int mrow0) {
    if (tid < BR) {
      int m = mrow0 + tid;
      long long xb = -1, yb = -1;
      if (m < mend) {
        int n = m / P, p = m % P;
        int oy = p / d.WO, ox = p % d.WO;
        xb = (long long)n * CIN * d.H * d.W + (long long)oy * S * d.W + ox * S;
        yb = (long long)n * COUT * P + p;
      }
      s_xbase[buf][tid] = xb; s_ybase[buf][tid] = yb;
    }
  };
  auto load = [&](int buf) {
    const int r = tid % BR;
    long long xb = s_xbase[buf][r], yb = s_ybase[buf][r];
#pragma unroll
    for (int j = 0; j < X_PER; ++j) {
      int kk = tid / BR + (NT / BR) * j, k = kt0 + kk;
      float v = 0.f;
      if (xb >= 0 && k < K) {
        int ci = k / (KS * KS), q = k % (KS * KS);
        v = x[xb + (long long)ci * d.H * d.W + (q / KS) * d.W + (q % KS)];
      }
      rx[j] = v;
    }
#pragma unroll
    for (int j = 0; j < D_PER; ++j) {
      int co = tid / BR + (NT / BR) * j;
      rd[j] = (yb >= 0 && co < COUT) ? dy[yb + (long long)co * P] : 0.f;
    }
  };
  auto store = [&](int buf) {
    const int r = tid % BR;
#pragma unroll
    for (int j = 0; j < X_PER; ++j) Xs[buf][r][tid / BR + (NT / BR) * j] = rx[j];
#pragma unroll
    for (int j = 0; j < D_PER; ++j) {
      int co = tid / BR + (NT / BR) * j;
      if (co < COUT) Ds[buf][r][co] = rd[j];
    }
  };

  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int nit = (mend - mbeg + BR - 1) / BR;
  if (nit > 0) {
    rowinfo(0, mbeg);
    __syncthreads();
    load(0);
    store(0);
    if (nit > 1) rowinfo(1, mbeg + BR);
  }
  __syncthreads();
  for (int it = 0; it < nit; ++it) {
    const int buf = it & 1;
    if (it + 1 < nit) load(buf ^ 1);
#pragma unroll
    for (int r = 0; r < BR; ++r) {
      float4 dv = *reinterpret_cast<const float4*>(&Ds[buf][r][cg * 4]);
      float4 x0 = *reinterpret_cast<const float4*>(&Xs[buf][r][kg * 4]);
      float4 x1 = *reinterpret_cast<const float4*>(&Xs[buf][r][KT / 2 + kg * 4]);
      float a[4] = {dv.x, dv.y, dv.z, dv.w};
      float b[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();  // everyone is done with tile `buf` and with row table buf^1 being loaded from
    if (it + 1 < nit) {
      store(buf ^ 1);
      if (it + 2 < nit) rowinfo(buf, mbeg + (it + 2) * BR);
    }
    __syncthreads();
  }
  float* out = partial + (size_t)blockIdx.x * COUT * K;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int co = cg * 4 + i;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int k = kt0 + (j / 4) * (KT / 2) + kg * 4 + (j % 4);
      if (k < K) out[(size_t)co * K + k] = acc[i][j];
    }
  }
}

// out[e] = beta*out[e] + sum_c partial[c][e]
__global__ void reduce_partials_kernel(const float* __restrict__ partial, float* __restrict__ out, int n, int chunks, float beta) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  float s = 0.f;
  for (int c = 0; c < chunks; ++c) s += partial[(size_t)c * n + e];
  out[e] = (beta != 0.f ? beta * out[e] : 0.f) + s;
}

template <int CIN, int KS, int S, int COUT, int BM>
int launch_fwd(const ConvDims& d, const float* x, const float* wt, const float* bias, float* y, int relu, cudaStream_t st) {
  FwdGeo<CIN, KS, S> geo;
  geo.d = d;
  int M = d.N * d.HO * d.WO;
  auto kfn = igemm_kernel<FwdGeo<CIN, KS, S>, COUT, BM>;
  HULC_LAUNCH(kfn, dim3(hulc_cdiv(M, BM)), dim3(NT), 0, st, geo, x, wt, bias, (const float*)nullptr, y, relu);
  HULC_RETURN_LAST();
}

template <int CIN, int KS, int S, int COUT, int BM>
int launch_dgrad(const ConvDims& d, const float* dy, const float* wp, const float* gate, float* dx, cudaStream_t st) {
  using Geo = DgradGeo<COUT, KS, S>;
  auto kfn = igemm_kernel<Geo, CIN, BM>;
  for (int ph = 0; ph < S * S; ++ph) {
    Geo geo;
    geo.d = d; geo.py = ph / S; geo.px = ph % S;
    geo.HP = (d.H - geo.py + S - 1) / S; geo.WP = (d.W - geo.px + S - 1) / S;
    int M = d.N * geo.HP * geo.WP;
    if (M <= 0) continue;
    HULC_LAUNCH(kfn, dim3(hulc_cdiv(M, BM)), dim3(NT), 0, st, geo, dy, wp + (size_t)ph * Geo::K * CIN, (const float*)nullptr, gate, dx, 0);
  }
  HULC_RETURN_LAST();
}

template <int CIN, int KS, int S, int COUT>
int launch_wgrad(const ConvDims& d, const float* x, const float* dy, float* dw, float beta, float* ws, size_t ws_bytes, cudaStream_t st) {
  constexpr int K = CIN * KS * KS;
  constexpr int KT = (NT / (COUT / 4)) * 8;
  int ktiles = hulc_cdiv(K, KT);
  int M = d.N * d.HO * d.WO;
  int chunks = max(1, min(hulc_cdiv(2 * kNumSMs, ktiles), hulc_cdiv(M, 256)));
  while (chunks > 1 && (size_t)chunks * COUT * K * sizeof(float) > ws_bytes) --chunks;
  if ((size_t)chunks * COUT * K * sizeof(float) > ws_bytes) return (int)cudaErrorInvalidValue;
  int rpc = hulc_cdiv(hulc_cdiv(M, chunks), 16) * 16;
  chunks = hulc_cdiv(M, rpc);
  auto kfn = wgrad_kernel<CIN, KS, S, COUT>;
  HULC_LAUNCH(kfn, dim3(chunks, ktiles), dim3(NT), 0, st, d, x, dy, ws, rpc);
  HULC_LAUNCH(reduce_partials_kernel, dim3(hulc_cdiv(COUT * K, 256)), dim3(256), 0, st, (const float*)ws, dw, COUT * K, chunks, beta);
  HULC_RETURN_LAST();
}

int conv_kind(int CIN, int COUT, int KS, int S) {
  if (CIN == 3 && COUT == 32 && KS == 8 && S == 4) return 1;
  if (CIN == 32 && COUT == 64 && KS == 4 && S == 2) return 2;
  if (CIN == 64 && COUT == 64 && KS == 3 && S == 1) return 3;
  return 0;
}

// The first kCounterFloats floats of the shared workspace hold the zero-initialised tickets of the split-K GEMM and the
// loss reductions; scratch starts after them.
constexpr size_t kCounterFloats = 1024;

}  // namespace

// y = [relu](conv2d(x, w, stride) + b).  x [N,CIN,H,W], w [COUT,CIN,KS,KS], y [N,COUT,HO,WO]; workspace >= COUT*CIN*KS*KS floats.
HULC_API int hulc_conv2d_fwd(const float* x, const float* w, const float* b, float* y, int N, int CIN, int H, int W, int COUT, int KS, int S,
                             int relu, float* workspace, size_t workspace_bytes, void* stream) {
  if (N <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (workspace_bytes <= kCounterFloats * sizeof(float)) return (int)cudaErrorInvalidValue;
  workspace += kCounterFloats; workspace_bytes -= kCounterFloats * sizeof(float);
  ConvDims d{N, CIN, H, W, COUT, (H - KS) / S + 1, (W - KS) / S + 1};
  int K = CIN * KS * KS;
  if (workspace_bytes < sizeof(float) * (size_t)COUT * K) return (int)cudaErrorInvalidValue;
  HULC_LAUNCH(weight_fwd_prep_kernel, dim3(hulc_cdiv(COUT * K, 256)), dim3(256), 0, st, w, workspace, COUT, K);
  switch (conv_kind(CIN, COUT, KS, S)) {
    case 1: return launch_fwd<3, 8, 4, 32, 256>(d, x, workspace, b, y, relu, st);
    case 2: return launch_fwd<32, 4, 2, 64, 128>(d, x, workspace, b, y, relu, st);
    case 3: return launch_fwd<64, 3, 1, 64, 128>(d, x, workspace, b, y, relu, st);
  }
  return (int)cudaErrorInvalidValue;
}

// dx = conv2d_transpose(dy, w) gated by (gate > 0) when gate != NULL (gate has dx's shape: the ReLU output that fed the conv)
HULC_API int hulc_conv2d_dgrad(const float* dy, const float* w, const float* gate, float* dx, int N, int CIN, int H, int W, int COUT, int KS, int S,
                               float* workspace, size_t workspace_bytes, void* stream) {
  if (N <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (workspace_bytes <= kCounterFloats * sizeof(float)) return (int)cudaErrorInvalidValue;
  workspace += kCounterFloats; workspace_bytes -= kCounterFloats * sizeof(float);
  ConvDims d{N, CIN, H, W, COUT, (H - KS) / S + 1, (W - KS) / S + 1};
  size_t n = (size_t)COUT * CIN * KS * KS;
  if (workspace_bytes < sizeof(float) * n || KS % S != 0) return (int)cudaErrorInvalidValue;
  HULC_LAUNCH(weight_dgrad_prep_kernel, dim3(hulc_cdiv((long long)n, 256)), dim3(256), 0, st, w, workspace, COUT, CIN, KS, S);
  switch (conv_kind(CIN, COUT, KS, S)) {
    case 2: return launch_dgrad<32, 4, 2, 64, 256>(d, dy, workspace, gate, dx, st);
    case 3: return launch_dgrad<64, 3, 1, 64, 128>(d, dy, workspace, gate, dx, st);
  }
  return (int)cudaErrorInvalidValue;  // conv1 needs no data gradient: its input is the image
}

// dw = beta*dw + sum dy (x) x ;  workspace holds the per-chunk partials
HULC_API int hulc_conv2d_wgrad(const float* x, const float* dy, float* dw, float beta, int N, int CIN, int H, int W, int COUT, int KS, int S,
                               float* workspace, size_t workspace_bytes, void* stream) {
  if (N <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (workspace_bytes <= kCounterFloats * sizeof(float)) return (int)cudaErrorInvalidValue;
  workspace += kCounterFloats; workspace_bytes -= kCounterFloats * sizeof(float);
  ConvDims d{N, CIN, H, W, COUT, (H - KS) / S + 1, (W - KS) / S + 1};
  switch (conv_kind(CIN, COUT, KS, S)) {
    case 1: return launch_wgrad<3, 8, 4, 32>(d, x, dy, dw, beta, workspace, workspace_bytes, st);
    case 2: return launch_wgrad<32, 4, 2, 64>(d, x, dy, dw, beta, workspace, workspace_bytes, st);
    case 3: return launch_wgrad<64, 3, 1, 64>(d, x, dy, dw, beta, workspace, workspace_bytes, st);
  }
  return (int)cudaErrorInvalidValue;
}
