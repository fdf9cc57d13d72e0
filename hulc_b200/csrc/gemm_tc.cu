// gemm_tc.cu — dense GEMM on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM) with the same
// fused epilogue as the CUDA-core path in gemm.cu (bias / broadcast addend / beta*C / ReLU|tanh / gate / dropout).
//
//   passes = 1 : operands rounded to tf32 (10-bit mantissa) — used where the parity budget allows (weight gradients)
//   passes = 3 : 3xTF32 split (hi*hi + hi*lo + lo*hi), fp32-level accuracy at 3 MMAs per product — used on the forward
//                path whose outputs are compared with the fp32 reference at rtol 1e-3 / atol 1e-4.
// The main loop lives in tc_pipeline.cuh.
#include "common.cuh"
#include "tc_pipeline.cuh"

namespace {

struct TcEpilogue {
  float* C;
  int M, N, ldc;
  float alpha, beta;
  const float* bias;
  const float* addend;
  int ldadd, add_mod;
  int act;
  const float* gate;
  int ldg;
  DropSpec drop;
  int BN, tiles_n;
  int splits;      // > 1: split-K, the kernel stores raw partial sums and splitk_epilogue_kernel finishes
  float* partial;  // [splits][M][N]
  float* C_lo;     // optional: C - tf32(C), the residual operand of a following 3xTF32 product
  int ldc_lo;

  __device__ __forceinline__ float one(float acc, int m, int n) const {
    float v = alpha * acc;
    if (bias) v += bias[n];
    if (addend) v += addend[(size_t)(add_mod ? m % add_mod : m) * ldadd + n];
    if (beta != 0.f) v += beta * C[(size_t)m * ldc + n];
    if ((act & 3) == 1) v = fmaxf(v, 0.f);
    else if ((act & 3) == 2) v = tanhf(v);
    if (gate) {
      float g = gate[(size_t)m * ldg + n];
      v = (act & 4) ? v * (1.f - g * g) : (g > 0.f ? v : 0.f);
    }
    v *= drop_factor(drop, (unsigned long long)m * N + n);
    return v;
  }
  __device__ __forceinline__ void store(float* dst_lo, float* dst, float v) const {
    *dst = v;
    if (C_lo) *dst_lo = v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
  }
  __device__ __forceinline__ void operator()(int tile, int row, int col0, const float* v) const {
    const int t2 = tile / splits, ks = tile - t2 * splits;
    const int tm = t2 / tiles_n, tn = t2 % tiles_n;
    const int m = tm * tc::kBM + row;
    if (m >= M) return;
    const int n0 = tn * BN + col0;
    if (n0 >= N) return;
    if (splits > 1) {
      float* dst = partial + ((size_t)ks * M + m) * N + n0;
      if (n0 + 32 <= N && (N & 3) == 0) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      } else {
        for (int j = 0; j < 32; ++j)
          if (n0 + j < N) dst[j] = v[j];
      }
      return;
    }
    float* dst = C + (size_t)m * ldc + n0;
    if (C_lo) {
      float* dlo = C_lo + (size_t)m * ldc_lo + n0;
      for (int j = 0; j < 32; ++j)
        if (n0 + j < N) store(dlo + j, dst + j, one(v[j], m, n0 + j));
      return;
    }
    const bool vec = ((reinterpret_cast<size_t>(dst) & 15) == 0) && n0 + 32 <= N;
    if (vec) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 o = make_float4(one(v[j], m, n0 + j), one(v[j + 1], m, n0 + j + 1), one(v[j + 2], m, n0 + j + 2), one(v[j + 3], m, n0 + j + 3));
        *reinterpret_cast<float4*>(dst + j) = o;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (n0 + j < N) dst[j] = one(v[j], m, n0 + j);
    }
  }
};

template <int BN, bool SPLIT, class AL, class BL>
__global__ void __launch_bounds__(tc::PipeCfg<BN, SPLIT>::kThreads, 1) gemm_tc_kernel(AL al, BL bl, TcEpilogue ep, int num_tiles, int num_kb) {
  tc::run_pipeline<BN, SPLIT, tc::kBK>(al, bl, ep, num_tiles, num_kb);
}

// lo[r][c] = x[r][c] - tf32_trunc(x[r][c]) for an R x Ccols region with leading dimension ld -> compact [R][Ccols]
__global__ void split_lo_kernel(const float* __restrict__ x, int ld, float* __restrict__ lo, int R, int Ccols) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long n = (long long)R * (Ccols / 4);
  if (i >= n) return;
  int r = (int)(i / (Ccols / 4)), c = (int)(i % (Ccols / 4)) * 4;
  float4 v = *reinterpret_cast<const float4*>(x + (size_t)r * ld + c);
  auto res = [](float f) { return f - __uint_as_float(__float_as_uint(f) & 0xFFFFE000u); };
  *reinterpret_cast<float4*>(lo + (size_t)r * Ccols + c) = make_float4(res(v.x), res(v.y), res(v.z), res(v.w));
}

// split-K finish: C[m][n] = epi(sum_s partial[s][m][n]) in a fixed order (bit-reproducible)
__global__ void splitk_epilogue_kernel(TcEpilogue ep) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)ep.M * ep.N) return;
  const int m = (int)(i / ep.N), n = (int)(i - (long long)m * ep.N);
  float s = 0.f;
  for (int k = 0; k < ep.splits; ++k) s += ep.partial[((size_t)k * ep.M + m) * ep.N + n];
  ep.store(ep.C_lo ? ep.C_lo + (size_t)m * ep.ldc_lo + n : nullptr, ep.C + (size_t)m * ep.ldc + n, ep.one(s, m, n));
}

// how many k-splits a skinny product gets so that its tiles cover the 148 SMs
inline int choose_splits(int tiles, int num_kb) {
  if (tiles * 2 > kNumSMs || num_kb < 16) return 1;
  int s = min(kNumSMs / tiles, num_kb / 4);
  return max(1, s);
}

template <int BN, bool SPLIT, class AL, class BL>
int launch(AL al, BL bl, TcEpilogue ep, int M, int N, int K, cudaStream_t st) {
  using Cfg = tc::PipeCfg<BN, SPLIT>;
  auto kfn = gemm_tc_kernel<BN, SPLIT, AL, BL>;
  HULC_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
  int tiles_m = hulc_cdiv(M, tc::kBM), tiles_n = hulc_cdiv(N, BN);
  const int num_kb = hulc_cdiv(K, tc::kBK);
  int kbps = num_kb;
  if (ep.splits > 1) {
    kbps = hulc_cdiv(num_kb, ep.splits);
    ep.splits = hulc_cdiv(num_kb, kbps);
  }
  const int tiles = tiles_m * tiles_n * ep.splits;
  ep.BN = BN; ep.tiles_n = tiles_n;
  al.tiles_other = tiles_n; al.is_n = 0; al.splits = ep.splits; al.kb_per_split = kbps;
  bl.tiles_other = tiles_n; bl.is_n = 1; bl.splits = ep.splits; bl.kb_per_split = kbps;
  HULC_LAUNCH(kfn, dim3(min(kNumSMs, tiles)), dim3(Cfg::kThreads), Cfg::kSmemBytes, st, al, bl, ep, tiles, kbps);
  if (ep.splits > 1) HULC_LAUNCH(splitk_epilogue_kernel, dim3(hulc_cdiv((long long)M * N, 256)), dim3(256), 0, st, ep);
  HULC_RETURN_LAST();
}

template <int BN, bool SPLIT>
int dispatch_layout(const float* A, const float* B, const float* Alo, const float* Blo, int M, int N, int K, int lda, int ldb, int transA,
                    int transB, TcEpilogue ep, cudaStream_t st) {
  // op(A) is M x K: stored M x K (transA = 0, K contiguous -> K-major tile) or K x M (transA = 1 -> MN-major tile).
  // op(B)^T is N x K: B stored N x K (transB = 1, the torch Linear weight) is the K-contiguous case.
  // lo buffers are compact copies (leading dimension = number of stored columns)
  tc::KMajorLoader<tc::kBM> ak{A, Alo, M, K, lda, K, 0, 0, 1, 0, 0, 0};
  tc::MNMajorLoader<tc::kBM> am{A, Alo, M, K, lda, M, 0, 0, 1, 0, 0, 0};
  tc::KMajorLoader<BN> bk{B, Blo, N, K, ldb, K, 0, 0, 1, 0, 0, 0};
  tc::MNMajorLoader<BN> bm{B, Blo, N, K, ldb, N, 0, 0, 1, 0, 0, 0};
  if (!transA && transB) return launch<BN, SPLIT>(ak, bk, ep, M, N, K, st);
  if (!transA && !transB) return launch<BN, SPLIT>(ak, bm, ep, M, N, K, st);
  if (transA && transB) return launch<BN, SPLIT>(am, bk, ep, M, N, K, st);
  return launch<BN, SPLIT>(am, bm, ep, M, N, K, st);
}

}  // namespace

HULC_API int hulc_split_lo(const float* x, int ld, float* lo, int rows, int cols, void* stream) {
  if (rows <= 0 || cols <= 0) return 0;
  if ((cols & 3) || (ld & 3) || (reinterpret_cast<size_t>(x) & 15) || (reinterpret_cast<size_t>(lo) & 15)) return (int)cudaErrorInvalidValue;
  HULC_LAUNCH(split_lo_kernel, dim3(hulc_cdiv((long long)rows * (cols / 4), 256)), dim3(256), 0, (cudaStream_t)stream, x, ld, lo, rows, cols);
  HULC_RETURN_LAST();
}

// Same contract as hulc_gemm (include/hulc_b200.h) plus `passes` (1 = tf32, 3 = 3xTF32).  Operand requirements: 16-byte
// aligned A and B, lda % 4 == ldb % 4 == 0, and the contiguous extent of each operand a multiple of 4 (K for K-contiguous
// storage, M / N for the transposed storage); otherwise cudaErrorInvalidValue (callers use hulc_gemm for such shapes).
// passes = 3 reads residual operands x - tf32(x): A_lo / B_lo when given (compact copies made with hulc_split_lo, e.g. a
// weight reused by every step of a recurrence), else computed into the workspace.  C_lo (optional) receives the residual of
// the result.  Skinny products (few output tiles, long K) are split along K; partial sums live in the workspace.
HULC_API int hulc_gemm_tc(const float* A, const float* B, float* C, int M, int N, int K, int lda, int ldb, int ldc, int transA, int transB,
                          float alpha, float beta, const float* bias, const float* addend, int ldadd, int add_mod, int act, const float* gate,
                          int ldg, float drop_p, unsigned long long drop_seed, unsigned drop_site, const unsigned char* drop_keep, int passes,
                          const float* A_lo, const float* B_lo, float* C_lo, int ldc_lo, float* workspace, size_t workspace_bytes, void* stream) {
  if (M <= 0 || N <= 0) return 0;
  if (K <= 0 || !A || !B || !C || (passes != 1 && passes != 3)) return (int)cudaErrorInvalidValue;
  if ((reinterpret_cast<size_t>(A) & 15) || (reinterpret_cast<size_t>(B) & 15) || (lda & 3) || (ldb & 3)) return (int)cudaErrorInvalidValue;
  if (((transA ? M : K) & 3) || ((transB ? K : N) & 3)) return (int)cudaErrorInvalidValue;
  TcEpilogue ep;
  ep.C = C; ep.M = M; ep.N = N; ep.ldc = ldc; ep.alpha = alpha; ep.beta = beta; ep.bias = bias; ep.addend = addend; ep.ldadd = ldadd;
  ep.add_mod = add_mod; ep.act = act; ep.gate = gate; ep.ldg = ldg; ep.drop = make_drop(drop_p, drop_seed, drop_site, drop_keep);
  ep.BN = 0; ep.tiles_n = 0; ep.splits = 1; ep.partial = nullptr; ep.C_lo = C_lo; ep.ldc_lo = ldc_lo;
  cudaStream_t st = (cudaStream_t)stream;
  const bool wide = N > 64;
  float* ws = workspace ? workspace + 1024 : nullptr;
  size_t ws_left = workspace_bytes > 4096 ? (workspace_bytes - 4096) / sizeof(float) : 0;
  if (passes == 3) {
    const size_t na = (size_t)M * K, nb = (size_t)N * K;
    if (!A_lo) {
      if (ws_left < na) return (int)cudaErrorInvalidValue;
      const int ar = transA ? K : M, ac = transA ? M : K;
      HULC_LAUNCH(split_lo_kernel, dim3(hulc_cdiv((long long)ar * (ac / 4), 256)), dim3(256), 0, st, A, lda, ws, ar, ac);
      A_lo = ws; ws += na; ws_left -= na;
    }
    if (!B_lo) {
      if (ws_left < nb) return (int)cudaErrorInvalidValue;
      const int br = transB ? N : K, bc = transB ? K : N;
      HULC_LAUNCH(split_lo_kernel, dim3(hulc_cdiv((long long)br * (bc / 4), 256)), dim3(256), 0, st, B, ldb, ws, br, bc);
      B_lo = ws; ws += nb; ws_left -= nb;
    }
  }
  const int tiles = hulc_cdiv(M, tc::kBM) * hulc_cdiv(N, wide ? 128 : 64);
  int splits = choose_splits(tiles, hulc_cdiv(K, tc::kBK));
  while (splits > 1 && (size_t)splits * M * N > ws_left) --splits;
  if (splits > 1) { ep.splits = splits; ep.partial = ws; }
  if (passes == 3) {
    if (wide) return dispatch_layout<128, true>(A, B, A_lo, B_lo, M, N, K, lda, ldb, transA, transB, ep, st);
    return dispatch_layout<64, true>(A, B, A_lo, B_lo, M, N, K, lda, ldb, transA, transB, ep, st);
  }
  if (wide) return dispatch_layout<128, false>(A, B, nullptr, nullptr, M, N, K, lda, ldb, transA, transB, ep, st);
  return dispatch_layout<64, false>(A, B, nullptr, nullptr, M, N, K, lda, ldb, transA, transB, ep, st);
}
