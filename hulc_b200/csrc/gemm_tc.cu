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
  __device__ __forceinline__ void operator()(int tile, int row, int col0, const float* v) const {
    const int tm = tile / tiles_n, tn = tile % tiles_n;
    const int m = tm * tc::kBM + row;
    if (m >= M) return;
    const int n0 = tn * BN + col0;
    if (n0 >= N) return;
    float* dst = C + (size_t)m * ldc + n0;
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

template <int BN, bool SPLIT, class AL, class BL>
int launch(AL al, BL bl, TcEpilogue ep, int M, int N, int K, cudaStream_t st) {
  using Cfg = tc::PipeCfg<BN, SPLIT>;
  auto kfn = gemm_tc_kernel<BN, SPLIT, AL, BL>;
  HULC_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
  int tiles_m = hulc_cdiv(M, tc::kBM), tiles_n = hulc_cdiv(N, BN);
  int grid = min(kNumSMs, tiles_m * tiles_n);
  ep.BN = BN; ep.tiles_n = tiles_n;
  al.tiles_other = tiles_n; al.is_n = 0;
  bl.tiles_other = tiles_n; bl.is_n = 1;
  HULC_LAUNCH(kfn, dim3(grid), dim3(Cfg::kThreads), Cfg::kSmemBytes, st, al, bl, ep, tiles_m * tiles_n, hulc_cdiv(K, tc::kBK));
  HULC_RETURN_LAST();
}

template <int BN, bool SPLIT>
int dispatch_layout(const float* A, const float* B, const float* Alo, const float* Blo, int M, int N, int K, int lda, int ldb, int transA,
                    int transB, TcEpilogue ep, cudaStream_t st) {
  // op(A) is M x K: stored M x K (transA = 0, K contiguous -> K-major tile) or K x M (transA = 1 -> MN-major tile).
  // op(B)^T is N x K: B stored N x K (transB = 1, the torch Linear weight) is the K-contiguous case.
  // lo buffers are compact copies (leading dimension = number of stored columns)
  tc::KMajorLoader<tc::kBM> ak{A, Alo, M, K, lda, K, 0, 0, 0};
  tc::MNMajorLoader<tc::kBM> am{A, Alo, M, K, lda, M, 0, 0, 0};
  tc::KMajorLoader<BN> bk{B, Blo, N, K, ldb, K, 0, 0, 0};
  tc::MNMajorLoader<BN> bm{B, Blo, N, K, ldb, N, 0, 0, 0};
  if (!transA && transB) return launch<BN, SPLIT>(ak, bk, ep, M, N, K, st);
  if (!transA && !transB) return launch<BN, SPLIT>(ak, bm, ep, M, N, K, st);
  if (transA && transB) return launch<BN, SPLIT>(am, bk, ep, M, N, K, st);
  return launch<BN, SPLIT>(am, bm, ep, M, N, K, st);
}

}  // namespace

// Same contract as hulc_gemm (include/hulc_b200.h) plus `passes` (1 = tf32, 3 = 3xTF32).  Operand requirements: 16-byte
// aligned A and B, lda % 4 == ldb % 4 == 0, and the contiguous extent of each operand a multiple of 4 (K for K-contiguous
// storage, M / N for the transposed storage); otherwise cudaErrorInvalidValue (callers use hulc_gemm for such shapes).
// passes = 3 needs workspace for the residual copies: (M*K + N*K) floats after the first 4096 bytes.
HULC_API int hulc_gemm_tc(const float* A, const float* B, float* C, int M, int N, int K, int lda, int ldb, int ldc, int transA, int transB,
                          float alpha, float beta, const float* bias, const float* addend, int ldadd, int add_mod, int act, const float* gate,
                          int ldg, float drop_p, unsigned long long drop_seed, unsigned drop_site, const unsigned char* drop_keep, int passes,
                          float* workspace, size_t workspace_bytes, void* stream) {
  if (M <= 0 || N <= 0) return 0;
  if (K <= 0 || !A || !B || !C || (passes != 1 && passes != 3)) return (int)cudaErrorInvalidValue;
  if ((reinterpret_cast<size_t>(A) & 15) || (reinterpret_cast<size_t>(B) & 15) || (lda & 3) || (ldb & 3)) return (int)cudaErrorInvalidValue;
  if (((transA ? M : K) & 3) || ((transB ? K : N) & 3)) return (int)cudaErrorInvalidValue;
  TcEpilogue ep;
  ep.C = C; ep.M = M; ep.N = N; ep.ldc = ldc; ep.alpha = alpha; ep.beta = beta; ep.bias = bias; ep.addend = addend; ep.ldadd = ldadd;
  ep.add_mod = add_mod; ep.act = act; ep.gate = gate; ep.ldg = ldg; ep.drop = make_drop(drop_p, drop_seed, drop_site, drop_keep);
  ep.BN = 0; ep.tiles_n = 0;
  cudaStream_t st = (cudaStream_t)stream;
  const bool wide = N > 64;
  if (passes == 3) {
    const size_t na = (size_t)M * K, nb = (size_t)N * K;
    if (!workspace || workspace_bytes < 4096 + (na + nb) * sizeof(float)) return (int)cudaErrorInvalidValue;
    float* Alo = workspace + 1024;
    float* Blo = Alo + na;
    const int ar = transA ? K : M, ac = transA ? M : K, br = transB ? N : K, bc = transB ? K : N;
    HULC_LAUNCH(split_lo_kernel, dim3(hulc_cdiv((long long)ar * (ac / 4), 256)), dim3(256), 0, st, A, lda, Alo, ar, ac);
    HULC_LAUNCH(split_lo_kernel, dim3(hulc_cdiv((long long)br * (bc / 4), 256)), dim3(256), 0, st, B, ldb, Blo, br, bc);
    if (wide) return dispatch_layout<128, true>(A, B, Alo, Blo, M, N, K, lda, ldb, transA, transB, ep, st);
    return dispatch_layout<64, true>(A, B, Alo, Blo, M, N, K, lda, ldb, transA, transB, ep, st);
  }
  if (wide) return dispatch_layout<128, false>(A, B, nullptr, nullptr, M, N, K, lda, ldb, transA, transB, ep, st);
  return dispatch_layout<64, false>(A, B, nullptr, nullptr, M, N, K, lda, ldb, transA, transB, ep, st);
}
