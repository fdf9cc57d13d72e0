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
  int splits;          // > 1: split-K: every CTA stores raw partial sums, the last one to finish a tile reduces them
  float* partial;      // [splits][M][N]
  unsigned* counters;  // one ticket per output tile, zero on entry and on exit

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
  // split-K: the CTA that delivers the last partial of an output tile sums all of them in a fixed order (bit-reproducible)
  // and applies the epilogue.  Called by the 128 epilogue threads after their partial stores.
  __device__ __forceinline__ void finish(int tile, int warp, int lane, uint32_t* flag) const {
    if (splits <= 1) return;
    const int t2 = tile / splits;
    __threadfence();
    tc::epi_bar_sync();
    if (warp == 0 && lane == 0) {
      const unsigned t = atomicAdd(&counters[t2], 1u);
      *flag = (t == (unsigned)(splits - 1));
      if (*flag) counters[t2] = 0u;
    }
    tc::epi_bar_sync();
    if (*flag) {
      __threadfence();
      const int tm = t2 / tiles_n, tn = t2 % tiles_n;
      const bool vec = (N & 3) == 0 && (ldc & 3) == 0 && (reinterpret_cast<size_t>(C) & 15) == 0;
      constexpr int RG = 4;  // rows per round: RG x 4 independent 16-byte loads in flight per thread hide the L2 latency
      for (int rb = warp * RG; rb < tc::kBM; rb += tc::kEpiWarps * RG) {
        const int m0 = tm * tc::kBM + rb;
        if (m0 >= M) break;
        for (int c = lane * 4; c < BN; c += 128) {
          const int n = tn * BN + c;
          if (n >= N) break;
          if (vec) {
            float4 acc[RG];
#pragma unroll
            for (int i = 0; i < RG; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int k0 = 0; k0 < splits; k0 += 4) {
              float4 p[RG][4];
#pragma unroll
              for (int i = 0; i < RG; ++i)
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                  const bool ok = m0 + i < M && k0 + kk < splits;
                  p[i][kk] = ok ? __ldcg(reinterpret_cast<const float4*>(partial + ((size_t)(k0 + kk) * M + m0 + i) * N + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
              for (int i = 0; i < RG; ++i)
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) { acc[i].x += p[i][kk].x; acc[i].y += p[i][kk].y; acc[i].z += p[i][kk].z; acc[i].w += p[i][kk].w; }
            }
#pragma unroll
            for (int i = 0; i < RG; ++i) {
              const int m = m0 + i;
              if (m < M)
                *reinterpret_cast<float4*>(C + (size_t)m * ldc + n) =
                    make_float4(one(acc[i].x, m, n), one(acc[i].y, m, n + 1), one(acc[i].z, m, n + 2), one(acc[i].w, m, n + 3));
            }
          } else {
            for (int i = 0; i < RG && m0 + i < M; ++i)
              for (int e = 0; e < 4 && n + e < N; ++e) {
                float s = 0.f;
                for (int k = 0; k < splits; ++k) s += __ldcg(partial + ((size_t)k * M + m0 + i) * N + n + e);
                C[(size_t)(m0 + i) * ldc + n + e] = one(s, m0 + i, n + e);
              }
          }
        }
      }
    }
    tc::epi_bar_sync();
  }
};

template <int BN, bool SPLIT, class AL, class BL>
__global__ void __launch_bounds__(tc::PipeCfg<BN, SPLIT>::kThreads, 1) gemm_tc_kernel(AL al, BL bl, TcEpilogue ep, int num_tiles, int num_kb) {
  tc::run_pipeline<BN, SPLIT, tc::kBK>(al, bl, ep, num_tiles, num_kb);
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
  al.tiles_n = tiles_n; al.is_n = 0; al.splits = ep.splits; al.kb_per_split = kbps;
  bl.tiles_n = tiles_n; bl.is_n = 1; bl.splits = ep.splits; bl.kb_per_split = kbps;
  HULC_LAUNCH(kfn, dim3(min(kNumSMs, tiles)), dim3(Cfg::kThreads), Cfg::kSmemBytes, st, al, bl, ep, tiles, kbps);
  HULC_RETURN_LAST();
}

template <int BN, bool SPLIT>
int dispatch_layout(const float* A, const float* B, int M, int N, int K, int lda, int ldb, int transA, int transB, TcEpilogue ep, cudaStream_t st) {
  // op(A) is M x K: stored M x K (transA = 0, K contiguous -> K-major tile) or K x M (transA = 1 -> MN-major tile).
  // op(B)^T is N x K: B stored N x K (transB = 1, the torch Linear weight) is the K-contiguous case.
  tc::KMajorLoader<tc::kBM> ak{A, M, K, lda, 0, 0, 1, 0, 0, 0};
  tc::MNMajorLoader<tc::kBM> am{A, M, K, lda, 0, 0, 1, 0, 0, 0};
  tc::KMajorLoader<BN> bk{B, N, K, ldb, 0, 0, 1, 0, 0, 0};
  tc::MNMajorLoader<BN> bm{B, N, K, ldb, 0, 0, 1, 0, 0, 0};
  if (!transA && transB) return launch<BN, SPLIT>(ak, bk, ep, M, N, K, st);
  if (!transA && !transB) return launch<BN, SPLIT>(ak, bm, ep, M, N, K, st);
  if (transA && transB) return launch<BN, SPLIT>(am, bk, ep, M, N, K, st);
  return launch<BN, SPLIT>(am, bm, ep, M, N, K, st);
}

}  // namespace

// Same contract as hulc_gemm (include/hulc_b200.h) plus `passes` (1 = tf32, 3 = 3xTF32).  Operand requirements: 16-byte
// aligned A and B, lda % 4 == ldb % 4 == 0, and the contiguous extent of each operand a multiple of 4 (K for K-contiguous
// storage, M / N for the transposed storage); otherwise cudaErrorInvalidValue (callers use hulc_gemm for such shapes).
// Skinny products (few output tiles, long K — the recurrent steps, the prior / goal MLPs) are split along K; the partial
// sums and tickets live in the workspace (zero-initialised head, as for hulc_gemm).
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
  ep.BN = 0; ep.tiles_n = 0; ep.splits = 1; ep.partial = nullptr; ep.counters = nullptr;
  cudaStream_t st = (cudaStream_t)stream;
  const bool wide = N > 64;
  const size_t ws_floats = workspace && workspace_bytes > 4096 ? (workspace_bytes - 4096) / sizeof(float) : 0;
  const int tiles = hulc_cdiv(M, tc::kBM) * hulc_cdiv(N, wide ? 128 : 64);
  int splits = choose_splits(tiles, hulc_cdiv(K, tc::kBK));
  while (splits > 1 && (size_t)splits * M * N > ws_floats) --splits;
  if (splits > 1) {
    ep.splits = splits; ep.partial = workspace + 1024; ep.counters = reinterpret_cast<unsigned*>(workspace);
  }
  if (passes == 3) {
    if (wide) return dispatch_layout<128, true>(A, B, M, N, K, lda, ldb, transA, transB, ep, st);
    return dispatch_layout<64, true>(A, B, M, N, K, lda, ldb, transA, transB, ep, st);
  }
  if (wide) return dispatch_layout<128, false>(A, B, M, N, K, lda, ldb, transA, transB, ep, st);
  return dispatch_layout<64, false>(A, B, M, N, K, lda, ldb, transA, transB, ep, st);
}
