// gemm_tc.cu — dense GEMM on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM) with the same
// fused epilogue as the CUDA-core path in gemm.cu (bias / broadcast addend / beta*C / ReLU|tanh / gate / dropout).
//
//   passes = 1 : operands rounded to tf32 (10-bit mantissa) — used where the parity budget allows (weight gradients)
//   passes = 3 : 3xTF32 split (hi*hi + hi*lo + lo*hi), fp32-level accuracy at 3 MMAs per product — used on the forward
//                path whose outputs are compared with the fp32 reference at rtol 1e-3 / atol 1e-4.
// The main loop lives in tc_pipeline.cuh.
#include "common.cuh"
#include "tc_pipeline.cuh"

int hulc_apply_dropout_rows(float* C, int M, int N, int ldc, DropSpec drop, cudaStream_t st);  // gemm.cu

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
  int BN, tiles_n;
  int splits;  // > 1: split-K over a thread-block cluster of that size (tile index = output tile * splits + k-slice)
  unsigned vec;  // 16-byte-aligned rows: bit 0 = C, bit 1 = addend, bit 2 = gate (epilogue reads go as float4); bits 3-5: 32-byte-aligned (256-bit reads)

  // The epilogue stages are applied as short vector passes over W contiguous columns of one row (uniform branches hoisted
  // out of the element loops keeps the unrolled code small — it is instruction-fetch bound otherwise).  Dropout is applied
  // by a separate elementwise pass (hulc_apply_dropout_rows).
  template <int W>
  __device__ __forceinline__ void apply(float* o, int m, int n0) const {  // all W columns valid
#pragma unroll
    for (int j = 0; j < W; ++j) o[j] *= alpha;
    if (bias) {
#pragma unroll
      for (int j = 0; j < W; ++j) o[j] += bias[n0 + j];
    }
    if (addend) {
      const float* a = addend + (size_t)(add_mod ? m % add_mod : m) * ldadd + n0;
      if (W >= 8 && (vec & 16)) {
#pragma unroll
        for (int j = 0; j + 7 < W; j += 8) {
          float t[8];
          tc::ld_global_v8(a + j, t);
#pragma unroll
          for (int e = 0; e < 8; ++e) o[j + e] += t[e];
        }
      } else if (W >= 4 && (vec & 2)) {
#pragma unroll
        for (int j = 0; j + 3 < W; j += 4) {
          const float4 t = *reinterpret_cast<const float4*>(a + j);
          o[j] += t.x; o[j + 1] += t.y; o[j + 2] += t.z; o[j + 3] += t.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < W; ++j) o[j] += a[j];
      }
    }
    if (beta != 0.f) {
      const float* c = C + (size_t)m * ldc + n0;
      if (W >= 8 && (vec & 8)) {
#pragma unroll
        for (int j = 0; j + 7 < W; j += 8) {
          float t[8];
          tc::ld_global_v8(c + j, t);
#pragma unroll
          for (int e = 0; e < 8; ++e) o[j + e] += beta * t[e];
        }
      } else if (W >= 4 && (vec & 1)) {
#pragma unroll
        for (int j = 0; j + 3 < W; j += 4) {
          const float4 t = *reinterpret_cast<const float4*>(c + j);
          o[j] += beta * t.x; o[j + 1] += beta * t.y; o[j + 2] += beta * t.z; o[j + 3] += beta * t.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < W; ++j) o[j] += beta * c[j];
      }
    }
    if ((act & 3) == 1) {
#pragma unroll
      for (int j = 0; j < W; ++j) o[j] = fmaxf(o[j], 0.f);
    } else if ((act & 3) == 2) {
#pragma unroll
      for (int j = 0; j < W; ++j) o[j] = tanhf(o[j]);
    }
    if (gate) {
      const float* gp = gate + (size_t)m * ldg + n0;
      float g[W];
      if (W >= 8 && (vec & 32)) {
#pragma unroll
        for (int j = 0; j + 7 < W; j += 8) tc::ld_global_v8(gp + j, g + j);
      } else if (W >= 4 && (vec & 4)) {
#pragma unroll
        for (int j = 0; j + 3 < W; j += 4) {
          const float4 t = *reinterpret_cast<const float4*>(gp + j);
          g[j] = t.x; g[j + 1] = t.y; g[j + 2] = t.z; g[j + 3] = t.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < W; ++j) g[j] = gp[j];
      }
      if (act & 4) {
#pragma unroll
        for (int j = 0; j < W; ++j) o[j] *= 1.f - g[j] * g[j];
      } else {
#pragma unroll
        for (int j = 0; j < W; ++j) o[j] = g[j] > 0.f ? o[j] : 0.f;
      }
    }
  }
  __device__ __forceinline__ float one(float acc, int m, int n) const {  // single element (ragged edges)
    float o[1] = {acc};
    apply<1>(o, m, n);
    return o[0];
  }
  __device__ __forceinline__ void operator()(int tile, int row, int col0, const float* v) const {
    const int t2 = tile / splits;
    const int tm = t2 / tiles_n, tn = t2 % tiles_n;
    const int m = tm * tc::kBM + row;
    if (m >= M) return;
    const int n0 = tn * BN + col0;
    if (n0 >= N) return;
    float* dst = C + (size_t)m * ldc + n0;
    if (n0 + 32 <= N) {
      float o[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) o[j] = v[j];
      apply<32>(o, m, n0);
      if ((reinterpret_cast<size_t>(dst) & 31) == 0) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) tc::st_global_v8(dst + j, o[j], o[j + 1], o[j + 2], o[j + 3], o[j + 4], o[j + 5], o[j + 6], o[j + 7]);
      } else if ((reinterpret_cast<size_t>(dst) & 15) == 0) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) dst[j] = o[j];
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (n0 + j < N) dst[j] = one(v[j], m, n0 + j);
    }
  }
  // cluster split-K hooks (see tc_pipeline.cuh)
  __device__ __forceinline__ int rows_valid(int tile) const { return min(tc::kBM, M - ((tile / splits) / tiles_n) * tc::kBM); }
  __device__ __forceinline__ void store4(int tile, int row, int col, float4 v) const {
    const int t2 = tile / splits;
    const int m = (t2 / tiles_n) * tc::kBM + row, n = (t2 % tiles_n) * BN + col;
    if (n >= N) return;
    float* dst = C + (size_t)m * ldc + n;
    if (n + 3 < N && (reinterpret_cast<size_t>(dst) & 15) == 0) {
      float o[4] = {v.x, v.y, v.z, v.w};
      apply<4>(o, m, n);
      *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
      const float a[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (n + e < N) dst[e] = one(a[e], m, n + e);
    }
  }
};

// 32-wide k-blocks per pipeline stage (tc_pipeline.cuh)
template <bool SPLIT>
constexpr int kps() { return 1; }  // measured: two k-blocks per stage (half the stages) is slower for BN = 64 / 128

template <int BN, bool SPLIT, int CLUSTER, class AL, class BL>
__global__ void __launch_bounds__(tc::PipeCfg<BN, SPLIT, tc::kBK, kps<SPLIT>()>::kThreads, 1) gemm_tc_kernel(AL al, BL bl, TcEpilogue ep, int num_tiles, int num_kb) {
  tc::run_pipeline<BN, SPLIT, tc::kBK, CLUSTER, AL, BL, TcEpilogue, kps<SPLIT>()>(al, bl, ep, num_tiles, num_kb);
}

// Tile width and k-slices (= cluster size) of a product.  Clusters of 8 / 4 / 2 CTAs of this kernel (one CTA per SM, ~200 KB
// of shared memory) fit 15 / 33 / 74 at a time on the 148 SMs (cudaOccupancyMaxActiveClusters); a skinny product takes the
// combination that puts the most CTAs to work in ONE wave.
inline void choose_config(int M, int N, int K, int& bn, int& splits) {
  const int num_kb = hulc_cdiv(K, tc::kBK), tiles_m = hulc_cdiv(M, tc::kBM);
  static const int kClusters[3] = {8, 4, 2}, kCap[3] = {15, 33, 74};
  bn = N > 64 ? 128 : 64; splits = 1;
  if (tiles_m * hulc_cdiv(N, bn) * 2 > kNumSMs) return;  // enough tiles to fill the machine without splitting
  int best = tiles_m * hulc_cdiv(N, bn);
  for (int b = 128; b >= 64; b >>= 1) {
    if (b == 128 && N <= 64) continue;
    const int tiles = tiles_m * hulc_cdiv(N, b);
    for (int i = 0; i < 3; ++i) {
      const int c = kClusters[i];
      if (tiles <= kCap[i] && num_kb >= 4 * c && tiles * c > best) { best = tiles * c; bn = b; splits = c; }
    }
  }
}

template <int BN, bool SPLIT, int CLUSTER, class AL, class BL>
int launch_one(AL al, BL bl, TcEpilogue ep, int tiles, int kbps, cudaStream_t st) {
  using Cfg = tc::PipeCfg<BN, SPLIT, tc::kBK, kps<SPLIT>()>;
  auto kfn = gemm_tc_kernel<BN, SPLIT, CLUSTER, AL, BL>;
  HULC_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
  if (CLUSTER == 1) {
    HULC_LAUNCH(kfn, dim3(min(kNumSMs, tiles)), dim3(Cfg::kThreads), Cfg::kSmemBytes, st, al, bl, ep, tiles, kbps);
    HULC_RETURN_LAST();
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(tiles);  // one work item per CTA; consecutive CTAs = the k-slices of one output tile = one cluster
  cfg.blockDim = dim3(Cfg::kThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CLUSTER; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  ++g_hulc_launches;
  HULC_TRY(cudaLaunchKernelEx(&cfg, kfn, al, bl, ep, tiles, kbps));
  HULC_RETURN_LAST();
}

template <int BN, bool SPLIT, class AL, class BL>
int launch(AL al, BL bl, TcEpilogue ep, int M, int N, int K, cudaStream_t st) {
  int tiles_m = hulc_cdiv(M, tc::kBM), tiles_n = hulc_cdiv(N, BN);
  const int num_kb = hulc_cdiv(K, tc::kBK);
  // k-blocks per k-slice, a whole number of stage rounds (a slice must not read into its neighbour's range; past K the loaders zero-fill)
  const int kbps = hulc_cdiv(hulc_cdiv(num_kb, ep.splits), kps<SPLIT>()) * kps<SPLIT>();
  const int tiles = tiles_m * tiles_n * ep.splits;
  ep.BN = BN; ep.tiles_n = tiles_n;
  al.tiles_n = tiles_n; al.is_n = 0; al.splits = ep.splits; al.kb_per_split = kbps;
  bl.tiles_n = tiles_n; bl.is_n = 1; bl.splits = ep.splits; bl.kb_per_split = kbps;
  switch (ep.splits) {
    case 8: return launch_one<BN, SPLIT, 8>(al, bl, ep, tiles, kbps, st);
    case 4: return launch_one<BN, SPLIT, 4>(al, bl, ep, tiles, kbps, st);
    case 2: return launch_one<BN, SPLIT, 2>(al, bl, ep, tiles, kbps, st);
    default: return launch_one<BN, SPLIT, 1>(al, bl, ep, tiles, kbps, st);
  }
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

// development aid: how many clusters of the skinny-product kernel can be resident at once
HULC_API int hulc_debug_max_clusters(int cluster, int split, int* out) {
  using AL = tc::KMajorLoader<tc::kBM>;
  using BL = tc::KMajorLoader<128>;
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cfg.gridDim = dim3(128);
  int n = -1;
  cudaError_t e;
  if (split) {
    using Cfg = tc::PipeCfg<128, true, tc::kBK, 1>;
    cfg.blockDim = dim3(Cfg::kThreads); cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    auto k8 = gemm_tc_kernel<128, true, 8, AL, BL>; auto k4 = gemm_tc_kernel<128, true, 4, AL, BL>; auto k2 = gemm_tc_kernel<128, true, 2, AL, BL>;
    cudaFuncSetAttribute(k8, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    cudaFuncSetAttribute(k4, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    e = cluster == 8 ? cudaOccupancyMaxActiveClusters(&n, k8, &cfg) : cluster == 4 ? cudaOccupancyMaxActiveClusters(&n, k4, &cfg) : cudaOccupancyMaxActiveClusters(&n, k2, &cfg);
  } else {
    using Cfg = tc::PipeCfg<128, false, tc::kBK, 1>;
    cfg.blockDim = dim3(Cfg::kThreads); cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    auto k8 = gemm_tc_kernel<128, false, 8, AL, BL>; auto k4 = gemm_tc_kernel<128, false, 4, AL, BL>; auto k2 = gemm_tc_kernel<128, false, 2, AL, BL>;
    cudaFuncSetAttribute(k8, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    cudaFuncSetAttribute(k4, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    e = cluster == 8 ? cudaOccupancyMaxActiveClusters(&n, k8, &cfg) : cluster == 4 ? cudaOccupancyMaxActiveClusters(&n, k4, &cfg) : cudaOccupancyMaxActiveClusters(&n, k2, &cfg);
  }
  *out = n;
  return (int)e;
}

#ifdef HULC_TC_TRACE
HULC_API int hulc_tc_trace_read(unsigned long long* host_out) {
  cudaDeviceSynchronize();
  return (int)cudaMemcpyFromSymbol(host_out, tc::g_tc_trace, sizeof(unsigned long long) * 64);
}
#endif

// Same contract as hulc_gemm (include/hulc_b200.h) plus `passes` (1 = tf32, 3 = 3xTF32).  Operand requirements: 16-byte
// aligned A and B, lda % 4 == ldb % 4 == 0, and the contiguous extent of each operand a multiple of 4 (K for K-contiguous
// storage, M / N for the transposed storage); otherwise cudaErrorInvalidValue (callers use hulc_gemm for such shapes).
// Skinny products (few output tiles, long K — the recurrent steps, the prior / goal MLPs, the decoder heads) are split along
// K over a thread-block cluster and reduced through distributed shared memory; the workspace is not used.
// gemm_bf16_tc.cu: the TMA-fed kernel with fp32 operands consumed as tf32 (one pass)
int hulc_gemm_tf32_tma(const float* A, const float* B, float* C, int M, int N, int K, int lda, int ldb, int ldc, int transA, int transB, float alpha, float beta,
                       const float* bias, const float* addend, int ldadd, int add_mod, int act, const float* gate, int ldg, float drop_p,
                       unsigned long long drop_seed, unsigned drop_site, const unsigned char* drop_keep, int passes, cudaStream_t st);

HULC_API int hulc_gemm_tc(const float* A, const float* B, float* C, int M, int N, int K, int lda, int ldb, int ldc, int transA, int transB,
                          float alpha, float beta, const float* bias, const float* addend, int ldadd, int add_mod, int act, const float* gate,
                          int ldg, float drop_p, unsigned long long drop_seed, unsigned drop_site, const unsigned char* drop_keep, int passes,
                          float* workspace, size_t workspace_bytes, void* stream) {
  if (M <= 0 || N <= 0) return 0;
  if (K <= 0 || !A || !B || !C || (passes != 1 && passes != 3)) return (int)cudaErrorInvalidValue;
  if ((reinterpret_cast<size_t>(A) & 15) || (reinterpret_cast<size_t>(B) & 15) || (lda & 3) || (ldb & 3)) return (int)cudaErrorInvalidValue;
  if (((transA ? M : K) & 3) || ((transB ? K : N) & 3)) return (int)cudaErrorInvalidValue;
  {
    // operands by TMA — twice the L2 -> shared-memory delivery rate of the cp.async producers
    // below (64 vs 30 B/clk/SM measured), same arithmetic (the tensor core truncates the fp32 containers either way).  HULC_B200_GEMM_TF32_TMA=0
    // keeps the cp.async kernel.
    // HULC_B200_GEMM_TF32_TMA: 0 = cp.async kernels for everything, 1 = TMA for the one-pass products only, 3 (default) = also for 3xTF32
    static const int use_tma = [] { const char* e = getenv("HULC_B200_GEMM_TF32_TMA"); return e ? atoi(e) : 3; }();
    if (use_tma >= passes) {
      const int rc = hulc_gemm_tf32_tma(A, B, C, M, N, K, lda, ldb, ldc, transA, transB, alpha, beta, bias, addend, ldadd, add_mod, act, gate, ldg, drop_p, drop_seed,
                                        drop_site, drop_keep, passes, (cudaStream_t)stream);
      if (rc != (int)cudaErrorNotSupported) return rc;
    }
  }
  TcEpilogue ep;
  ep.C = C; ep.M = M; ep.N = N; ep.ldc = ldc; ep.alpha = alpha; ep.beta = beta; ep.bias = bias; ep.addend = addend; ep.ldadd = ldadd;
  ep.add_mod = add_mod; ep.act = act; ep.gate = gate; ep.ldg = ldg;
  ep.BN = 0; ep.tiles_n = 0;
  auto aligned = [](const float* p, int ld) { return p && (reinterpret_cast<size_t>(p) & 15) == 0 && (ld & 3) == 0; };
  auto aligned32 = [](const float* p, int ld) { return p && (reinterpret_cast<size_t>(p) & 31) == 0 && (ld & 7) == 0; };
  ep.vec = (aligned(C, ldc) ? 1u : 0u) | (aligned(addend, ldadd) ? 2u : 0u) | (aligned(gate, ldg) ? 4u : 0u) | (aligned32(C, ldc) ? 8u : 0u) |
           (aligned32(addend, ldadd) ? 16u : 0u) | (aligned32(gate, ldg) ? 32u : 0u);
  cudaStream_t st = (cudaStream_t)stream;
  int bn;
  choose_config(M, N, K, bn, ep.splits);
  const bool wide = bn == 128;
  if (passes == 3) {
    if (wide) HULC_TRY((dispatch_layout<128, true>(A, B, M, N, K, lda, ldb, transA, transB, ep, st)));
    else HULC_TRY((dispatch_layout<64, true>(A, B, M, N, K, lda, ldb, transA, transB, ep, st)));
  } else {
    if (wide) HULC_TRY((dispatch_layout<128, false>(A, B, M, N, K, lda, ldb, transA, transB, ep, st)));
    else HULC_TRY((dispatch_layout<64, false>(A, B, M, N, K, lda, ldb, transA, transB, ep, st)));
  }
  return hulc_apply_dropout_rows(C, M, N, ldc, make_drop(drop_p, drop_seed, drop_site, drop_keep), st);
}
