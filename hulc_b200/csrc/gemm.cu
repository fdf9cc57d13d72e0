// gemm.cu — fp32 GEMM with fused epilogues (bias / broadcast addend / beta*C / ReLU / ReLU-gate / dropout) and a
// deterministic split-K for the skinny shapes of the HULC step (M = 32..64 sequences against 2048-wide weights).
//
// This is the exact-fp32 path (CUDA-core FFMA); it is what config 2 (fp32, rtol 1e-3) is validated on.  The tcgen05
// path (gemm_tc.cu) reuses the same C-ABI.
//
//   C[M,N] = epi( alpha * op(A)[M,K] * op(B)[K,N] )
//   transA = 0: A is M x K row-major (lda), 1: A is stored K x M row-major (lda)        [wgrad: dY^T]
//   transB = 0: B is K x N row-major (ldb), 1: B is stored N x K row-major (ldb)        [torch Linear weight]
#include "common.cuh"

namespace {

struct GemmParams {
  const float* A; const float* B; float* C;
  int M, N, K, lda, ldb, ldc;
  float alpha, beta;
  const float* bias;            // [N] or null
  const float* addend;          // null or [*, N]: C += addend[(add_mod ? m % add_mod : m) * ldadd + n]
  int ldadd, add_mod;
  int act;                      // bits 0-1: 0 none, 1 relu, 2 tanh;  bit 2: the gate is a tanh output
  const float* gate; int ldg;   // null or [M,N]: C = gate > 0 ? C : 0   (bit 2 set: C *= 1 - gate^2)
  DropSpec drop;                // applied last, element index m*N+n
  // split-K
  int splits, k_per_split;
  float* partial;               // [splits][tiles][BM*BN]
  unsigned* counters;           // [tiles], zero on entry, zero on exit
};

// Epilogue stages as short vector passes over W contiguous columns of one row: the branches are uniform, hoisting them
// out of the element loop keeps the fully unrolled thread-tile epilogue small (it is instruction-fetch bound otherwise).
// Dropout is applied afterwards by drop_rows_kernel.
template <int W>
__device__ __forceinline__ void epilogue_vec(const GemmParams& p, float* o, int m, int n0) {
#pragma unroll
  for (int j = 0; j < W; ++j) o[j] *= p.alpha;
  if (p.bias) {
#pragma unroll
    for (int j = 0; j < W; ++j) o[j] += p.bias[n0 + j];
  }
  if (p.addend) {
    const float* a = p.addend + (size_t)(p.add_mod ? m % p.add_mod : m) * p.ldadd + n0;
#pragma unroll
    for (int j = 0; j < W; ++j) o[j] += a[j];
  }
  if (p.beta != 0.f) {
    const float* c = p.C + (size_t)m * p.ldc + n0;
#pragma unroll
    for (int j = 0; j < W; ++j) o[j] += p.beta * c[j];
  }
  if ((p.act & 3) == 1) {
#pragma unroll
    for (int j = 0; j < W; ++j) o[j] = fmaxf(o[j], 0.f);
  } else if ((p.act & 3) == 2) {
#pragma unroll
    for (int j = 0; j < W; ++j) o[j] = tanhf(o[j]);
  }
  if (p.gate) {
    const float* g = p.gate + (size_t)m * p.ldg + n0;
    if (p.act & 4) {
#pragma unroll
      for (int j = 0; j < W; ++j) o[j] *= 1.f - g[j] * g[j];
    } else {
#pragma unroll
      for (int j = 0; j < W; ++j) o[j] = g[j] > 0.f ? o[j] : 0.f;
    }
  }
}

template <int BM, int BN, int BK, int TM, int TN, bool TA, bool TB>
__global__ void __launch_bounds__((BM / TM) * (BN / TN)) sgemm_kernel(GemmParams p) {
  constexpr int NT = (BM / TM) * (BN / TN);
  constexpr int PAD = 4;
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];
  __shared__ unsigned s_ticket;

  const int tid = threadIdx.x;
  const int tn = tid % (BN / TN), tm = tid / (BN / TN);
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * p.k_per_split;
  const int kend = min(p.K, kbeg + p.k_per_split);

  // ---- global -> register staging -------------------------------------------------------------------------------
  // A tile: BM x BK.  !TA: vectors run along k (4 per row);  TA: vectors run along m.
  constexpr int A_V = BM * BK / 4 / NT;  // float4 per thread
  constexpr int B_V = BN * BK / 4 / NT;
  static_assert(A_V >= 1 && B_V >= 1, "tile too small for the thread count");
  float4 ra[A_V], rb[B_V];
  const bool a_vec = ((reinterpret_cast<size_t>(p.A) & 15) == 0) && (p.lda % 4 == 0);
  const bool b_vec = ((reinterpret_cast<size_t>(p.B) & 15) == 0) && (p.ldb % 4 == 0);

  auto load_a = [&](int k0) {
#pragma unroll
    for (int j = 0; j < A_V; ++j) {
      int v = tid + j * NT;
      float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!TA) {
        int row = v / (BK / 4), kq = (v % (BK / 4)) * 4;
        int m = m0 + row, k = k0 + kq;
        if (m < p.M) {
          const float* src = p.A + (size_t)m * p.lda + k;
          if (a_vec && k + 3 < kend) r = *reinterpret_cast<const float4*>(src);
          else {
            if (k < kend) r.x = src[0];
            if (k + 1 < kend) r.y = src[1];
            if (k + 2 < kend) r.z = src[2];
            if (k + 3 < kend) r.w = src[3];
          }
        }
      } else {
        int kk = v / (BM / 4), mq = (v % (BM / 4)) * 4;
        int m = m0 + mq, k = k0 + kk;
        if (k < kend) {
          const float* src = p.A + (size_t)k * p.lda + m;
          if (a_vec && m + 3 < p.M) r = *reinterpret_cast<const float4*>(src);
          else {
            if (m < p.M) r.x = src[0];
            if (m + 1 < p.M) r.y = src[1];
            if (m + 2 < p.M) r.z = src[2];
            if (m + 3 < p.M) r.w = src[3];
          }
        }
      }
      ra[j] = r;
    }
  };
  auto load_b = [&](int k0) {
#pragma unroll
    for (int j = 0; j < B_V; ++j) {
      int v = tid + j * NT;
      float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
      if (TB) {
        int row = v / (BK / 4), kq = (v % (BK / 4)) * 4;
        int n = n0 + row, k = k0 + kq;
        if (n < p.N) {
          const float* src = p.B + (size_t)n * p.ldb + k;
          if (b_vec && k + 3 < kend) r = *reinterpret_cast<const float4*>(src);
          else {
            if (k < kend) r.x = src[0];
            if (k + 1 < kend) r.y = src[1];
            if (k + 2 < kend) r.z = src[2];
            if (k + 3 < kend) r.w = src[3];
          }
        }
      } else {
        int kk = v / (BN / 4), nq = (v % (BN / 4)) * 4;
        int n = n0 + nq, k = k0 + kk;
        if (k < kend) {
          const float* src = p.B + (size_t)k * p.ldb + n;
          if (b_vec && n + 3 < p.N) r = *reinterpret_cast<const float4*>(src);
          else {
            if (n < p.N) r.x = src[0];
            if (n + 1 < p.N) r.y = src[1];
            if (n + 2 < p.N) r.z = src[2];
            if (n + 3 < p.N) r.w = src[3];
          }
        }
      }
      rb[j] = r;
    }
  };
  auto store_a = [&](int buf) {
#pragma unroll
    for (int j = 0; j < A_V; ++j) {
      int v = tid + j * NT;
      if (!TA) {
        int row = v / (BK / 4), kq = (v % (BK / 4)) * 4;
        As[buf][kq + 0][row] = ra[j].x; As[buf][kq + 1][row] = ra[j].y;
        As[buf][kq + 2][row] = ra[j].z; As[buf][kq + 3][row] = ra[j].w;
      } else {
        int kk = v / (BM / 4), mq = (v % (BM / 4)) * 4;
        *reinterpret_cast<float4*>(&As[buf][kk][mq]) = ra[j];
      }
    }
  };
  auto store_b = [&](int buf) {
#pragma unroll
    for (int j = 0; j < B_V; ++j) {
      int v = tid + j * NT;
      if (TB) {
        int row = v / (BK / 4), kq = (v % (BK / 4)) * 4;
        Bs[buf][kq + 0][row] = rb[j].x; Bs[buf][kq + 1][row] = rb[j].y;
        Bs[buf][kq + 2][row] = rb[j].z; Bs[buf][kq + 3][row] = rb[j].w;
      } else {
        int kk = v / (BN / 4), nq = (v % (BN / 4)) * 4;
        *reinterpret_cast<float4*>(&Bs[buf][kk][nq]) = rb[j];
      }
    }
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  // Each thread owns TM rows as TM/4 groups of 4 spaced BM/(TM/4) apart (conflict-free LDS.128), same for columns.
  constexpr int GM = TM / 4, GN = TN / 4;
  constexpr int SM_STRIDE = BM / GM, SN_STRIDE = BN / GN;

  const int nk = (kend - kbeg + BK - 1) / BK;
  if (nk > 0) {
    load_a(kbeg); load_b(kbeg);
    store_a(0); store_b(0);
  }
  __syncthreads();
  for (int it = 0; it < nk; ++it) {
    const int buf = it & 1;
    if (it + 1 < nk) { load_a(kbeg + (it + 1) * BK); load_b(kbeg + (it + 1) * BK); }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int g = 0; g < GM; ++g) {
        float4 t = *reinterpret_cast<const float4*>(&As[buf][k][g * SM_STRIDE + tm * 4]);
        a[g * 4 + 0] = t.x; a[g * 4 + 1] = t.y; a[g * 4 + 2] = t.z; a[g * 4 + 3] = t.w;
      }
#pragma unroll
      for (int g = 0; g < GN; ++g) {
        float4 t = *reinterpret_cast<const float4*>(&Bs[buf][k][g * SN_STRIDE + tn * 4]);
        b[g * 4 + 0] = t.x; b[g * 4 + 1] = t.y; b[g * 4 + 2] = t.z; b[g * 4 + 3] = t.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (it + 1 < nk) { store_a(buf ^ 1); store_b(buf ^ 1); }
    __syncthreads();
  }

  // ---- epilogue -------------------------------------------------------------------------------------------------
  if (p.splits > 1) {
    const int tile = blockIdx.y * gridDim.x + blockIdx.x;
    const int ntiles = gridDim.x * gridDim.y;
    float* mine = p.partial + ((size_t)blockIdx.z * ntiles + tile) * (BM * BN);
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int g = 0; g < GN; ++g) {
        int lm = (i / 4) * SM_STRIDE + tm * 4 + (i % 4), ln = g * SN_STRIDE + tn * 4;
        *reinterpret_cast<float4*>(&mine[lm * BN + ln]) = make_float4(acc[i][g * 4], acc[i][g * 4 + 1], acc[i][g * 4 + 2], acc[i][g * 4 + 3]);
      }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_ticket = atomicAdd(&p.counters[tile], 1u);
    __syncthreads();
    if (s_ticket != (unsigned)(p.splits - 1)) return;
    __threadfence();
    // last CTA of this tile: fixed-order reduction over the splits -> bit-reproducible
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    for (int z = 0; z < p.splits; ++z) {
      const float* src = p.partial + ((size_t)z * ntiles + tile) * (BM * BN);
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int g = 0; g < GN; ++g) {
          int lm = (i / 4) * SM_STRIDE + tm * 4 + (i % 4), ln = g * SN_STRIDE + tn * 4;
          float4 t = __ldcg(reinterpret_cast<const float4*>(&src[lm * BN + ln]));
          acc[i][g * 4] += t.x; acc[i][g * 4 + 1] += t.y; acc[i][g * 4 + 2] += t.z; acc[i][g * 4 + 3] += t.w;
        }
    }
    if (tid == 0) p.counters[tile] = 0u;
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int m = m0 + (i / 4) * SM_STRIDE + tm * 4 + (i % 4);
    if (m >= p.M) continue;
#pragma unroll
    for (int g = 0; g < GN; ++g) {
      const int n = n0 + g * SN_STRIDE + tn * 4;
      float* dst = p.C + (size_t)m * p.ldc + n;
      if (n + 3 < p.N) {
        float o[4] = {acc[i][g * 4], acc[i][g * 4 + 1], acc[i][g * 4 + 2], acc[i][g * 4 + 3]};
        epilogue_vec<4>(p, o, m, n);
        if ((reinterpret_cast<size_t>(dst) & 15) == 0) *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
        else { dst[0] = o[0]; dst[1] = o[1]; dst[2] = o[2]; dst[3] = o[3]; }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n + j < p.N) {
            float o[1] = {acc[i][g * 4 + j]};
            epilogue_vec<1>(p, o, m, n + j);
            dst[j] = o[0];
          }
      }
    }
  }
}

template <int BM, int BN, int TM, int TN>
int launch_cfg(GemmParams p, int transA, int transB, float* workspace, size_t workspace_bytes, cudaStream_t st) {
  constexpr int BK = 16;
  constexpr int NT = (BM / TM) * (BN / TN);
  int gm = hulc_cdiv(p.M, BM), gn = hulc_cdiv(p.N, BN);
  int tiles = gm * gn;
  // split-K when the tile grid cannot fill the 148 SMs and K is deep enough to be worth a second pass
  int splits = 1;
  if (workspace && tiles < kNumSMs && p.K >= 8 * BK) {
    splits = min(hulc_cdiv(2 * kNumSMs, tiles), p.K / (4 * BK));
    splits = max(1, min(splits, 32));
    size_t need = 4096 + (size_t)splits * tiles * BM * BN * sizeof(float);
    while (splits > 1 && (need > workspace_bytes || tiles > 1024)) {
      --splits;
      need = 4096 + (size_t)splits * tiles * BM * BN * sizeof(float);
    }
  }
  int kps = hulc_cdiv(hulc_cdiv(p.K, splits), BK) * BK;
  splits = hulc_cdiv(p.K, kps);
  p.splits = splits; p.k_per_split = kps;
  if (splits > 1) {
    p.counters = reinterpret_cast<unsigned*>(workspace);
    p.partial = workspace + 1024;
  }
  dim3 grid(gn, gm, splits), block(NT);
  void (*kfn)(GemmParams);
  if (!transA && !transB) kfn = sgemm_kernel<BM, BN, BK, TM, TN, false, false>;
  else if (!transA && transB) kfn = sgemm_kernel<BM, BN, BK, TM, TN, false, true>;
  else if (transA && !transB) kfn = sgemm_kernel<BM, BN, BK, TM, TN, true, false>;
  else kfn = sgemm_kernel<BM, BN, BK, TM, TN, true, true>;
  HULC_LAUNCH(kfn, grid, block, 0, st, p);
  HULC_RETURN_LAST();
}

// C[m][n] *= dropout factor of element m*N + n  (the `drop` stage of the GEMM epilogue contract)
__global__ void drop_rows_kernel(float* __restrict__ C, int M, int N, int ldc, DropSpec drop) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)M * N) return;
  const int m = (int)(i / N), n = (int)(i - (long long)m * N);
  C[(size_t)m * ldc + n] *= drop_factor(drop, (unsigned long long)i);
}

// same, four consecutive columns per thread: one Philox call yields the four keep decisions (philox_uniform(idx) reads word
// idx & 3 of counter idx >> 2), and the row is read / written as float4
__global__ void drop_rows_vec4_kernel(float* __restrict__ C, int M, int N4, int ldc, DropSpec drop) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (long long)M * N4) return;
  const int m = (int)(q / N4), n = (int)(q - (long long)m * N4) * 4;
  float4* p = reinterpret_cast<float4*>(C + (size_t)m * ldc + n);
  float4 v = *p;
  bool k0, k1, k2, k3;
  if (drop.keep) {
    const uchar4 k = *reinterpret_cast<const uchar4*>(drop.keep + (unsigned long long)q * 4);
    k0 = k.x != 0; k1 = k.y != 0; k2 = k.z != 0; k3 = k.w != 0;
  } else {
    const uint4 r = philox4x32(rng_seed(drop.seed, drop.seed_ptr), drop.site, (unsigned long long)q);
    const float s = 1.0f / 16777216.0f;
    k0 = (float)(r.x >> 8) * s >= drop.p; k1 = (float)(r.y >> 8) * s >= drop.p;
    k2 = (float)(r.z >> 8) * s >= drop.p; k3 = (float)(r.w >> 8) * s >= drop.p;
  }
  v.x = k0 ? v.x * drop.scale : 0.f; v.y = k1 ? v.y * drop.scale : 0.f;
  v.z = k2 ? v.z * drop.scale : 0.f; v.w = k3 ? v.w * drop.scale : 0.f;
  *p = v;
}

// out[c] = beta*out[c] + sum_r X[r*ldx + c]
// Column sums out[c] = beta*out[c] + sum_r X[r][c].  A block of 256 threads = CL column lanes (VEC columns each) x 256/CL row
// lanes; the rows are split over gridDim.y blocks whose partial sums meet in `partial` and are added in a fixed order by the
// last block to arrive (ticket) -> bit-reproducible, no atomics on the data.
template <int VEC>
__global__ void colsum_kernel(const float* __restrict__ X, int rows, int cols, int ldx, float* __restrict__ out, float beta, int rows_per_block,
                              int CL, float* __restrict__ partial, unsigned* __restrict__ counters) {
  __shared__ float red[256 * VEC];
  __shared__ unsigned s_ticket;
  const int tid = threadIdx.x;
  const int cl = tid % CL, rl = tid / CL, RL = 256 / CL;
  const int c = (blockIdx.x * CL + cl) * VEC;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  float s[4][VEC];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < VEC; ++v) s[u][v] = 0.f;
  if (c < cols) {
    int r = r0 + rl;
    for (; r + 3 * RL < r1; r += 4 * RL) {  // four independent loads in flight per thread
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float* src = X + (size_t)(r + u * RL) * ldx + c;
        if (VEC == 4) {
          const float4 t = *reinterpret_cast<const float4*>(src);
          s[u][0] += t.x; s[u][VEC > 1 ? 1 : 0] += t.y; s[u][VEC > 2 ? 2 : 0] += t.z; s[u][VEC > 3 ? 3 : 0] += t.w;
        } else {
          s[u][0] += src[0];
        }
      }
    }
    for (; r < r1; r += RL) {
      const float* src = X + (size_t)r * ldx + c;
#pragma unroll
      for (int v = 0; v < VEC; ++v) s[0][v] += src[v];
    }
  }
#pragma unroll
  for (int v = 0; v < VEC; ++v) red[tid * VEC + v] = (s[0][v] + s[1][v]) + (s[2][v] + s[3][v]);
  __syncthreads();
  float t[VEC];
  if (rl == 0 && c < cols) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) t[v] = 0.f;
    for (int i = 0; i < RL; ++i)
#pragma unroll
      for (int v = 0; v < VEC; ++v) t[v] += red[(i * CL + cl) * VEC + v];
    if (gridDim.y == 1) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) out[c + v] = (beta != 0.f ? beta * out[c + v] : 0.f) + t[v];
    } else {
#pragma unroll
      for (int v = 0; v < VEC; ++v) partial[(size_t)blockIdx.y * cols + c + v] = t[v];
    }
  }
  if (gridDim.y == 1) return;
  __threadfence();
  __syncthreads();
  if (tid == 0) s_ticket = atomicAdd(&counters[blockIdx.x], 1u);
  __syncthreads();
  if (s_ticket != gridDim.y - 1) return;
  __threadfence();
  // last block of this column group: every row lane adds its share of the partial rows, then the lanes meet in shared
  // memory in lane order (fixed summation order for a given grid)
#pragma unroll
  for (int v = 0; v < VEC; ++v) t[v] = 0.f;
  if (c < cols)
    for (int y = rl; y < (int)gridDim.y; y += RL)
#pragma unroll
      for (int v = 0; v < VEC; ++v) t[v] += __ldcg(&partial[(size_t)y * cols + c + v]);
#pragma unroll
  for (int v = 0; v < VEC; ++v) red[tid * VEC + v] = t[v];
  __syncthreads();
  if (rl == 0 && c < cols) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) t[v] = 0.f;
    for (int i = 0; i < RL; ++i)
#pragma unroll
      for (int v = 0; v < VEC; ++v) t[v] += red[(i * CL + cl) * VEC + v];
#pragma unroll
    for (int v = 0; v < VEC; ++v) out[c + v] = (beta != 0.f ? beta * out[c + v] : 0.f) + t[v];
  }
  if (tid == 0) counters[blockIdx.x] = 0u;
}
__global__ void scale_vec_kernel(float* x, int n, float s) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] *= s;
}

// Many column sums in ONE launch (the bias gradients of a whole backward pass: ~45 small reductions that are otherwise launch-bound).
// Job j (8 x int64 in `table`): X, out, rows, cols, ldx, beta (float bits), first block, unused.  A block owns 32 consecutive columns of one
// job for all of its rows: 8 column lanes x 4 columns, 32 row lanes whose partial sums meet in shared memory in lane order (fixed
// summation order, no atomics).
__global__ void __launch_bounds__(256) colsum_multi_kernel(const long long* __restrict__ table, int njobs) {
  __shared__ float red[32][33];
  __shared__ int s_job;
  if (threadIdx.x == 0) {
    int j = 0;
    while (j + 1 < njobs && (int)table[(j + 1) * 8 + 6] <= (int)blockIdx.x) ++j;
    s_job = j;
  }
  __syncthreads();
  const long long* t = table + (size_t)s_job * 8;
  const float* X = reinterpret_cast<const float*>(t[0]);
  float* out = reinterpret_cast<float*>(t[1]);
  const int rows = (int)t[2], cols = (int)t[3], ldx = (int)t[4];
  const float beta = __int_as_float((int)t[5]);
  const int c0 = ((int)blockIdx.x - (int)t[6]) * 32;
  const int cl = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const int c = c0 + cl * 4;
  float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
  if (c + 3 < cols && (ldx & 3) == 0 && (reinterpret_cast<size_t>(X) & 15) == 0) {
    int r = rl;
    for (; r + 32 < rows; r += 64) {  // two independent 16-byte loads in flight per trip
      const float4 a = *reinterpret_cast<const float4*>(X + (size_t)r * ldx + c);
      const float4 b = *reinterpret_cast<const float4*>(X + (size_t)(r + 32) * ldx + c);
      s0[0] += a.x; s0[1] += a.y; s0[2] += a.z; s0[3] += a.w;
      s1[0] += b.x; s1[1] += b.y; s1[2] += b.z; s1[3] += b.w;
    }
    if (r < rows) {
      const float4 a = *reinterpret_cast<const float4*>(X + (size_t)r * ldx + c);
      s0[0] += a.x; s0[1] += a.y; s0[2] += a.z; s0[3] += a.w;
    }
  } else {
    for (int r = rl; r < rows; r += 32)
#pragma unroll
      for (int v = 0; v < 4; ++v)
        if (c + v < cols) s0[v] += X[(size_t)r * ldx + c + v];
  }
#pragma unroll
  for (int v = 0; v < 4; ++v) red[rl][cl * 4 + v] = s0[v] + s1[v];
  __syncthreads();
  if (threadIdx.x < 32 && c0 + (int)threadIdx.x < cols) {
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) acc += red[i][threadIdx.x];
    float* o = out + c0 + threadIdx.x;
    *o = (beta != 0.f ? beta * *o : 0.f) + acc;
  }
}

}  // namespace

// shared with gemm_tc.cu
int hulc_apply_dropout_rows(float* C, int M, int N, int ldc, DropSpec drop, cudaStream_t st) {
  if (drop.p <= 0.f) return 0;
  if (N % 4 == 0 && ldc % 4 == 0 && (reinterpret_cast<size_t>(C) & 15) == 0 && (reinterpret_cast<size_t>(drop.keep) & 3) == 0) {
    HULC_LAUNCH(drop_rows_vec4_kernel, dim3(hulc_cdiv((long long)M * (N / 4), 256)), dim3(256), 0, st, C, M, N / 4, ldc, drop);
    HULC_RETURN_LAST();
  }
  HULC_LAUNCH(drop_rows_kernel, dim3(hulc_cdiv((long long)M * N, 256)), dim3(256), 0, st, C, M, N, ldc, drop);
  HULC_RETURN_LAST();
}

unsigned long long g_hulc_launches = 0;
const unsigned long long* g_hulc_rng_offset_ptr = nullptr;

HULC_API int hulc_set_rng_offset_ptr(const unsigned long long* device_ptr) {
  g_hulc_rng_offset_ptr = device_ptr;
  return 0;
}

HULC_API int hulc_launch_count(unsigned long long* out) {
  if (!out) return (int)cudaErrorInvalidValue;
  *out = g_hulc_launches;
  return 0;
}

// See include/hulc_b200.h for the contract.
HULC_API int hulc_gemm(const float* A, const float* B, float* C, int M, int N, int K, int lda, int ldb, int ldc, int transA,
                       int transB, float alpha, float beta, const float* bias, const float* addend, int ldadd, int add_mod, int act,
                       const float* gate, int ldg, float drop_p, unsigned long long drop_seed, unsigned drop_site,
                       const unsigned char* drop_keep, float* workspace, size_t workspace_bytes, void* stream) {
  if (M <= 0 || N <= 0) return 0;
  if (K < 0 || !A || !B || !C) return (int)cudaErrorInvalidValue;
  GemmParams p;
  p.A = A; p.B = B; p.C = C; p.M = M; p.N = N; p.K = K; p.lda = lda; p.ldb = ldb; p.ldc = ldc;
  p.alpha = alpha; p.beta = beta; p.bias = bias; p.addend = addend; p.ldadd = ldadd; p.add_mod = add_mod;
  p.act = act; p.gate = gate; p.ldg = ldg; p.drop = make_drop(drop_p, drop_seed, drop_site, drop_keep);
  p.splits = 1; p.k_per_split = K; p.partial = nullptr; p.counters = nullptr;
  cudaStream_t st = (cudaStream_t)stream;
  if (M >= 256 && N >= 96) HULC_TRY((launch_cfg<128, 128, 8, 8>(p, transA, transB, workspace, workspace_bytes, st)));
  else HULC_TRY((launch_cfg<64, 64, 4, 4>(p, transA, transB, workspace, workspace_bytes, st)));
  return hulc_apply_dropout_rows(C, M, N, ldc, p.drop, st);
}

HULC_API int hulc_colsum(const float* X, int rows, int cols, int ldx, float* out, float beta, float* workspace, size_t workspace_bytes,
                        void* stream) {
  if (cols <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (cols % 4 == 0) && (ldx % 4 == 0) && ((reinterpret_cast<size_t>(X) & 15) == 0);
  const int V = vec ? 4 : 1;
  int CL = 32;
  while (CL > 8 && (CL / 2) * V >= cols) CL /= 2;  // narrow matrices: more row lanes per block
  const int gx = hulc_cdiv(cols, CL * V);
  // split the rows over blocks until ~4 blocks per SM are in flight (each row lane keeps >= 8 rows); the partial sums go
  // through the workspace (tickets in its first 1024 words, shared with the split-K GEMM: zero on entry, zero on exit)
  int gy = 1;
  if (workspace && gx <= 1024) {
    gy = max(1, min(rows / (8 * (256 / CL)), (4 * kNumSMs) / gx));
    while (gy > 1 && (1024 + (size_t)gy * cols) * sizeof(float) > workspace_bytes) --gy;
  }
  const int rpb = hulc_cdiv(rows, gy);
  gy = hulc_cdiv(rows, rpb);
  unsigned* counters = reinterpret_cast<unsigned*>(workspace);
  float* partial = workspace ? workspace + 1024 : nullptr;
  if (vec) HULC_LAUNCH(colsum_kernel<4>, dim3(gx, gy), dim3(256), 0, st, X, rows, cols, ldx, out, beta, rpb, CL, partial, counters);
  else HULC_LAUNCH(colsum_kernel<1>, dim3(gx, gy), dim3(256), 0, st, X, rows, cols, ldx, out, beta, rpb, CL, partial, counters);
  HULC_RETURN_LAST();
}

HULC_API int hulc_colsum_multi(const void* table, int njobs, int total_blocks, void* stream) {
  if (njobs <= 0 || total_blocks <= 0) return 0;
  HULC_LAUNCH(colsum_multi_kernel, dim3(total_blocks), dim3(256), 0, (cudaStream_t)stream, reinterpret_cast<const long long*>(table), njobs);
  HULC_RETURN_LAST();
}
