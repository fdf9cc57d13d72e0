// gemm_bf16_tc.cu — the dense GEMM of the bf16 path (BASELINE config 3): bf16 operands staged by TMA (cp.async.bulk.tensor, 128-byte
// swizzle), tcgen05.mma kind::f16 with fp32 accumulators in TMEM, and the same fused epilogue contract as hulc_gemm / hulc_gemm_tc
// (alpha, bias, broadcast addend, beta*C, ReLU | tanh, ReLU / tanh' gate, dropout) computed in fp32 on the accumulator.  The result
// is written as fp32 (C) and / or as bf16 (Cb — the operand of the next product), so a chain of Linear layers never re-reads an
// fp32 activation just to narrow it.
//
//   D[128 x BN] (TMEM) (+)= A_tile[128 x 64] * B_tile[BN x 64]^T  per k-block of 64 bf16 (= one 128-byte swizzled row), 4 MMAs of K = 16
//
// One persistent CTA per SM walks the output tiles.  Roles (192 threads):
//   warps 0-3  epilogue (warp w owns TMEM lanes 32w..32w+31 = tile rows); two accumulator buffers in TMEM let the epilogue of tile i
//              overlap the main loop of tile i+1
//   warp  4    allocates TMEM; one lane issues every tcgen05.mma and the tcgen05.commit that recycles the stage
//   warp  5    one lane issues the TMA loads of both operands
// Operand layouts (the four combinations torch's Linear forward / dgrad / wgrad need):
//   K-major   (stored rows x K, K contiguous)   : tensor map (K, rows), box 64 x ROWS -> [rows][128 B], SWIZZLE_128B, SBO 1024
//   MN-major  (stored K x rows, rows contiguous): tensor map (rows, K), boxes 64 x 64 -> [64 k][128 B] per group of 64 rows, the
//              canonical MN-major SWIZZLE_128B layout of 16-bit types (8 k-rows per swizzle atom, SBO 1024, groups LBO = 8 KB apart)
// Out-of-range rows / k are the TMA's zero fill: no bounds logic in the main loop.
// Skinny products (few output tiles, long K) split K over a thread-block cluster; the partial tiles meet in distributed shared
// memory and are reduced in a fixed order (no workspace, bit-reproducible) — the scheme of tc_pipeline.cuh.
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_pipeline.cuh"
#include "tma.cuh"

namespace {

using namespace tc;

constexpr int kBK16 = 64;                           // bf16 elements per k-block
constexpr int kATile16 = kBM * kRowBytes;           // 16 KB
constexpr int kGroupB = 64 * kRowBytes;             // one MN-major group: 64 k-rows x 128 B
constexpr int kThreadsG = (kEpiWarps + 2) * 32;
// The same kernel serves fp32 operands consumed as tf32 (TF = true; hulc_gemm_tc's 1-pass products): a k-block row is still 128 bytes (32 floats),
// the stage layout and byte counts are identical, only the MN-major layout differs — 32-element groups under SWIZZLE_128B_BASE32B, which the TMA
// writes with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B (scripts/micro/tma_mn_probe.cu) — and the instruction kind.
template <bool TF>
struct Elem {
  static constexpr int kBK = TF ? 32 : 64;            // elements per k-block (128 bytes)
  static constexpr int kG = TF ? 32 : 64;             // rows of an MN-major group = k-rows per k-block
  static constexpr int kGroupBytes = kG * kRowBytes;  // 4 KB / 8 KB
  static constexpr uint32_t kMnStep = TF ? 1024u : 2048u;  // bytes between the k-slices of consecutive MMAs in an MN-major tile (8 / 16 k-rows)
};

// P3 (fp32 operands only): 3xTF32 — the TMA lands the raw fp32 tiles (the tensor core reads them as their tf32 truncation, "hi"); four splitter
// warps write the residual x - tf32(x) of every 16-byte chunk into a second ("lo") copy of the stage, and the issuer adds lo*hi + hi*lo (side
// accumulator) to hi*hi: fp32-level accuracy, the arithmetic of hulc_gemm_tc's 3-pass kernel with its operands delivered by TMA.
template <int BN, bool P3 = false>
struct GCfg {
  static constexpr int kBTile = BN * kRowBytes;
  static constexpr int kHi = kATile16 + kBTile;          // what the TMA delivers per stage
  static constexpr int kStage = (P3 ? 2 : 1) * kHi;
  static constexpr int kStagesRaw = (200 * 1024) / kStage;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kSmem = kStages * kStage + 1024 + 256;
  static constexpr int kAcc = (P3 ? 2 : 1) * BN;         // TMEM columns of one accumulator buffer (main + side)
  static constexpr int kTmemCols = 2 * kAcc < 32 ? 32 : 2 * kAcc;
  static constexpr int kSplitWarps = P3 ? 4 : 0;
  static constexpr int kThreads = kThreadsG + kSplitWarps * 32;
  static_assert(kStages >= 3 && kTmemCols <= 512, "pipeline depth / TMEM columns");
};

struct GBars {
  uint64_t full[8];
  uint64_t split[8];   // (3xTF32) the lo tiles of the stage are written
  uint64_t empty[8];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

// Instruction descriptor, kind::f16 with bf16 operands and fp32 accumulate (cute::UMMA::InstrDescriptor):
//   [4,6) D format = 1 (F32) | [7,10) A format = 1 (BF16) | [10,13) B format = 1 | [15] A major (0 = K, 1 = MN) | [16] B major |
//   [17,23) N >> 3 | [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// operand slice of one MMA (K = 16 bf16) inside a stage tile
template <bool MN, bool TF = false>
__device__ __forceinline__ uint64_t desc16(uint32_t tile) {
  if (MN && TF) return make_smem_desc(tile, (uint32_t)Elem<true>::kGroupBytes, 512u, 1u);  // SWIZZLE_128B_BASE32B: 4-row atoms, k-step + 8 k-rows = 1024 B
  if (MN) return make_smem_desc(tile, (uint32_t)kGroupB, 1024u, 2u);  // k-step: + 16 k-rows = 2048 B
  return make_smem_desc(tile, 16u, 1024u, 2u);                        // k-step: + 32 B inside the 128-byte row
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void st_global_v8_b32(void* p, const uint32_t* v) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]),
               "r"(v[7])
               : "memory");
}

struct BfEpilogue {
  float* C;             // fp32 result (may be null)
  __nv_bfloat16* Cb;    // bf16 result (may be null)
  int M, N, ldc, ldcb;
  float alpha, beta;    // beta applies to the fp32 C (read-modify-write); requires C
  const float* bias;
  const float* addend;
  int ldadd, add_mod;
  int act;              // bits 0-1: 1 ReLU, 2 tanh; bit 2: the gate is tanh' (1 - g^2) instead of ReLU' (g > 0)
  const float* gate;    // fp32 gate ...
  const __nv_bfloat16* gate_b;  // ... or its bf16 twin (one of the two at most)
  int ldg;
  DropSpec drop;
  int BN, tiles_n, splits;

  template <int W>
  __device__ __forceinline__ void apply(float* o, int m, int n0) const {  // W consecutive valid columns of row m (W = 32, 4 or 1)
#pragma unroll
    for (int j = 0; j < W; ++j) o[j] *= alpha;
    if (bias) {
      if (W >= 4 && ((reinterpret_cast<size_t>(bias + n0) & 15) == 0)) {
#pragma unroll
        for (int j = 0; j + 3 < W; j += 4) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(bias + n0 + j));
          o[j] += t.x; o[j + 1] += t.y; o[j + 2] += t.z; o[j + 3] += t.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < W; ++j) o[j] += __ldg(bias + n0 + j);
      }
    }
    if (addend) {
      const float* a = addend + (size_t)(add_mod ? m % add_mod : m) * ldadd + n0;
      if (W >= 4 && ((reinterpret_cast<size_t>(a) & 15) == 0)) {
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
      if (W >= 4 && ((reinterpret_cast<size_t>(c) & 15) == 0)) {
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
    if (gate || gate_b) {
      float g[W];
      if (gate) {
        const float* gp = gate + (size_t)m * ldg + n0;
        if (W >= 4 && ((reinterpret_cast<size_t>(gp) & 15) == 0)) {
#pragma unroll
          for (int j = 0; j + 3 < W; j += 4) {
            const float4 t = *reinterpret_cast<const float4*>(gp + j);
            g[j] = t.x; g[j + 1] = t.y; g[j + 2] = t.z; g[j + 3] = t.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < W; ++j) g[j] = gp[j];
        }
      } else {
        const __nv_bfloat16* gp = gate_b + (size_t)m * ldg + n0;
        if (W >= 8 && ((reinterpret_cast<size_t>(gp) & 15) == 0)) {
#pragma unroll
          for (int j = 0; j + 7 < W; j += 8) {
            const uint4 t = *reinterpret_cast<const uint4*>(gp + j);
            const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {  // bf16 -> fp32 is a 16-bit shift
              g[j + 2 * e] = __uint_as_float(w[e] << 16);
              g[j + 2 * e + 1] = __uint_as_float(w[e] & 0xFFFF0000u);
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < W; ++j) g[j] = __bfloat162float(gp[j]);
        }
      }
      if (act & 4) {
#pragma unroll
        for (int j = 0; j < W; ++j) o[j] *= 1.f - g[j] * g[j];
      } else {
#pragma unroll
        for (int j = 0; j < W; ++j) o[j] = g[j] > 0.f ? o[j] : 0.f;
      }
    }
    if (drop.p > 0.f) {  // element index m*N + n, as hulc_apply_dropout_rows (gemm.cu): four consecutive elements share one Philox call
      const unsigned long long e0 = (unsigned long long)m * (unsigned long long)N + (unsigned long long)n0;
      if (W >= 4 && !drop.keep && (e0 & 3ull) == 0) {
        const unsigned long long seed = rng_seed(drop.seed, drop.seed_ptr);
#pragma unroll
        for (int j = 0; j + 3 < W; j += 4) {
          const uint4 r = philox4x32(seed, drop.site, (e0 + j) >> 2);
          const float s = 1.0f / 16777216.0f;
          o[j] = (float)(r.x >> 8) * s >= drop.p ? o[j] * drop.scale : 0.f;
          o[j + 1] = (float)(r.y >> 8) * s >= drop.p ? o[j + 1] * drop.scale : 0.f;
          o[j + 2] = (float)(r.z >> 8) * s >= drop.p ? o[j + 2] * drop.scale : 0.f;
          o[j + 3] = (float)(r.w >> 8) * s >= drop.p ? o[j + 3] * drop.scale : 0.f;
        }
      } else {
#pragma unroll
        for (int j = 0; j < W; ++j) o[j] *= drop_factor(drop, e0 + j);
      }
    }
  }
  template <int W>
  __device__ __forceinline__ void store(const float* o, int m, int n0) const {
    if (C) {
      float* dst = C + (size_t)m * ldc + n0;
      if (W == 32 && (reinterpret_cast<size_t>(dst) & 31) == 0) {
#pragma unroll
        for (int j = 0; j < W; j += 8) st_global_v8(dst + j, o[j], o[j + 1], o[j + 2], o[j + 3], o[j + 4], o[j + 5], o[j + 6], o[j + 7]);
      } else if (W >= 4 && (reinterpret_cast<size_t>(dst) & 15) == 0) {
#pragma unroll
        for (int j = 0; j + 3 < W; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < W; ++j) dst[j] = o[j];
      }
    }
    if (Cb) {
      __nv_bfloat16* dst = Cb + (size_t)m * ldcb + n0;
      if (W == 32 && (reinterpret_cast<size_t>(dst) & 31) == 0) {
        uint32_t v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = pack_bf16(o[2 * j], o[2 * j + 1]);
        st_global_v8_b32(dst, v);
        st_global_v8_b32(dst + 16, v + 8);
      } else if (W >= 4 && (reinterpret_cast<size_t>(dst) & 7) == 0) {
#pragma unroll
        for (int j = 0; j + 3 < W; j += 4) *reinterpret_cast<uint2*>(dst + j) = make_uint2(pack_bf16(o[j], o[j + 1]), pack_bf16(o[j + 2], o[j + 3]));
      } else {
#pragma unroll
        for (int j = 0; j < W; ++j) dst[j] = __float2bfloat16_rn(o[j]);
      }
    }
  }
  // one 32-column chunk of tile row `row`
  __device__ __forceinline__ void operator()(int tile, int row, int col0, const float* v) const {
    const int t2 = tile / splits;
    const int tm = t2 / tiles_n, tn = t2 - tm * tiles_n;
    const int m = tm * kBM + row;
    if (m >= M) return;
    const int n0 = tn * BN + col0;
    if (n0 >= N) return;
    if (n0 + 32 <= N) {
      float o[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) o[j] = v[j];
      apply<32>(o, m, n0);
      store<32>(o, m, n0);
    } else {
      for (int j = 0; j < 32 && n0 + j < N; ++j) {
        float o[1] = {v[j]};
        apply<1>(o, m, n0 + j);
        store<1>(o, m, n0 + j);
      }
    }
  }
  // cluster split-K hooks
  __device__ __forceinline__ int rows_valid(int tile) const { return min(kBM, M - ((tile / splits) / tiles_n) * kBM); }
  __device__ __forceinline__ void store4(int tile, int row, int col, float4 s) const {
    const int t2 = tile / splits;
    const int m = (t2 / tiles_n) * kBM + row, n = (t2 % tiles_n) * BN + col;
    if (n >= N) return;
    if (n + 3 < N) {
      float o[4] = {s.x, s.y, s.z, s.w};
      apply<4>(o, m, n);
      store<4>(o, m, n);
    } else {
      const float a[4] = {s.x, s.y, s.z, s.w};
      for (int e = 0; e < 4 && n + e < N; ++e) {
        float o[1] = {a[e]};
        apply<1>(o, m, n + e);
        store<1>(o, m, n + e);
      }
    }
  }
};

template <int BN, int CLUSTER, bool A_MN, bool B_MN, bool TF = false, bool P3 = false>
__global__ void __launch_bounds__((GCfg<BN, P3>::kThreads), 1) gemm_bf16_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                                                                                BfEpilogue ep, int num_tiles, int kb_per_split) {
  static_assert(!P3 || TF, "the 3-pass split exists for fp32 operands");
  using Cfg = GCfg<BN, P3>;
  constexpr int kThr = Cfg::kThreads;
  constexpr int S = Cfg::kStages;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  GBars* bars = reinterpret_cast<GBars*>(smem + S * Cfg::kStage);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->split[s], 4);
      mbar_init(&bars->empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bars->tmem_full[a], 1);
      mbar_init(&bars->tmem_empty[a], kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == kEpiWarps) tmem_alloc(&bars->tmem_base, Cfg::kTmemCols);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;
  const int splits = ep.splits, tiles_n = ep.tiles_n;

  if (warp < kEpiWarps) {
    // ================================ epilogue ================================
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int a = it & 1;
      mbar_wait(&bars->tmem_full[a], (it >> 1) & 1);
      tc_fence_after_sync();
      if constexpr (CLUSTER > 1) asm volatile("bar.sync 1, %0;" ::"n"(kThr) : "memory");  // producers are done with the stage buffers (see tc_pipeline.cuh)
      const int row = warp * 32 + lane;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(a * Cfg::kAcc + c0), r);
        if constexpr (P3) {  // + the side accumulator (cross terms)
          uint32_t r2[32];
          tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(a * Cfg::kAcc + BN + c0), r2);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) + __uint_as_float(r2[i]));
        } else {
          tmem_ld_wait();
        }
        if constexpr (CLUSTER > 1) {
          float* prow = reinterpret_cast<float*>(smem) + (size_t)row * (BN + 4) + c0;
#pragma unroll
          for (int q = 0; q < 32; q += 4)
            *reinterpret_cast<float4*>(prow + q) = make_float4(__uint_as_float(r[q]), __uint_as_float(r[q + 1]), __uint_as_float(r[q + 2]), __uint_as_float(r[q + 3]));
        } else {
          if (c0 + 32 == BN) {  // the accumulator is in registers: hand the TMEM buffer back before the global stores
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->tmem_empty[a]);
          }
          ep(tile, row, c0, reinterpret_cast<const float*>(r));
        }
      }
      if constexpr (CLUSTER > 1) {
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->tmem_empty[a]);
      }
    }
  } else if (warp == kEpiWarps) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      constexpr uint32_t idesc = TF ? make_idesc_tf32(kBM, BN, A_MN, B_MN) : make_idesc_bf16(kBM, BN, A_MN, B_MN);
      constexpr uint32_t kAStep = (A_MN ? Elem<TF>::kMnStep : 32u) >> 4, kBStep = (B_MN ? Elem<TF>::kMnStep : 32u) >> 4;
      const uint32_t s0 = smem_u32(smem);
      const uint64_t a0 = desc16<A_MN, TF>(s0), b0 = desc16<B_MN, TF>(s0 + kATile16);
      int j = 0, it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int a = it & 1;
        mbar_wait(&bars->tmem_empty[a], ((it >> 1) & 1) ^ 1);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + (uint32_t)(a * Cfg::kAcc);
        for (int kb = 0; kb < kb_per_split; ++kb, ++j) {
          const int stage = j % S;
          mbar_wait(P3 ? &bars->split[stage] : &bars->full[stage], (j / S) & 1);
          tc_fence_after_sync();
          const uint32_t soff = (uint32_t)(stage * Cfg::kStage) >> 4;
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // 4 MMAs per 128-byte k-block: K = 16 bf16 or 8 tf32 each
            const uint32_t first = (uint32_t)((kb | k) != 0);
            const uint64_t da = a0 + (soff + k * kAStep), db = b0 + (soff + k * kBStep);
            if constexpr (P3) {
              constexpr uint32_t kLo = (uint32_t)Cfg::kHi >> 4;  // the lo copy of a tile sits kHi bytes behind it
              umma_tf32(d_tmem + BN, da + kLo, db, idesc, first);
              umma_tf32(d_tmem + BN, da, db + kLo, idesc, 1u);
              umma_tf32(d_tmem, da, db, idesc, first);
            } else if (TF) {
              umma_tf32(d_tmem, da, db, idesc, first);
            } else {
              umma_bf16(d_tmem, da, db, idesc, first);
            }
          }
          umma_commit(&bars->empty[stage]);
          if (kb == kb_per_split - 1) umma_commit(&bars->tmem_full[a]);
        }
      }
    }
  } else if (P3 && warp >= kEpiWarps + 2) {
    // ================================ splitters (3xTF32): lo = x - tf32(x), chunk by chunk, smem -> smem ================================
    const int stid = threadIdx.x - (kEpiWarps + 2) * 32;  // 0..127
    int j = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < kb_per_split; ++kb, ++j) {
        const int stage = j % S;
        mbar_wait(&bars->full[stage], (j / S) & 1);
        const uint32_t hi = smem_u32(smem) + stage * Cfg::kStage;
#pragma unroll 4
        for (int i = stid; i < Cfg::kHi / 16; i += 128) lo_chunk(hi + (uint32_t)i * 16u, hi + (uint32_t)Cfg::kHi + (uint32_t)i * 16u);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->split[stage]);
      }
    }
  } else {
    // ================================ TMA producer ================================
    if (lane == 0) {
      tma::prefetch_map(&mapA);
      tma::prefetch_map(&mapB);
      int j = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int t2 = tile / splits, ks = tile - t2 * splits;
        const int tm = t2 / tiles_n, tn = t2 - tm * tiles_n;
        const int m0 = tm * kBM, n0 = tn * BN, kb0 = ks * kb_per_split;
        for (int kb = 0; kb < kb_per_split; ++kb, ++j) {
          const int stage = j % S;
          mbar_wait(&bars->empty[stage], ((j / S) & 1) ^ 1);
          const uint32_t dstA = smem_u32(smem) + stage * Cfg::kStage, dstB = dstA + kATile16;
          using El = Elem<TF>;
          const int k0 = (kb0 + kb) * El::kBK;
          tma::expect_tx(&bars->full[stage], (uint32_t)Cfg::kHi);
          if (A_MN) {
#pragma unroll
            for (int g = 0; g < kBM / El::kG; ++g) tma::load_2d(dstA + g * El::kGroupBytes, &mapA, &bars->full[stage], m0 + g * El::kG, k0);
          } else {
            tma::load_2d(dstA, &mapA, &bars->full[stage], k0, m0);
          }
          if (B_MN) {
#pragma unroll
            for (int g = 0; g < BN / El::kG; ++g) tma::load_2d(dstB + g * El::kGroupBytes, &mapB, &bars->full[stage], n0 + g * El::kG, k0);
          } else {
            tma::load_2d(dstB, &mapB, &bars->full[stage], k0, n0);
          }
        }
      }
    }
  }

  if constexpr (CLUSTER > 1) {
    if (warp >= kEpiWarps && (int)blockIdx.x < num_tiles) asm volatile("bar.sync 1, %0;" ::"n"(kThr) : "memory");
    static_assert(kBM * (BN + 4) * 4 <= Cfg::kStages * Cfg::kStage, "partial tile must fit in the stage buffers");
    cluster_sync_all();  // all partial tiles of the cluster are in place
    if (warp < kEpiWarps && (int)blockIdx.x < num_tiles) {
      const int tile = blockIdx.x;
      const uint32_t rank = cluster_ctarank();
      const int rows = ep.rows_valid(tile);
      constexpr int CQ = BN / 4;
      const uint32_t base = smem_u32(smem);
      for (int idx = threadIdx.x; idx < kBM * CQ; idx += kEpiWarps * 32) {
        const int rr = idx / CQ, cq = idx - rr * CQ;
        const int row = rr * CLUSTER + (int)rank;
        if (row >= rows) break;
        const uint32_t off = base + (uint32_t)(row * (BN + 4) + cq * 4) * 4u;
        float4 s = ld_dsmem16(off, 0u);
#pragma unroll
        for (int q = 1; q < CLUSTER; ++q) {
          const float4 p = ld_dsmem16(off, (uint32_t)q);
          s.x += p.x; s.y += p.y; s.z += p.z; s.w += p.w;
        }
        ep.store4(tile, row, cq * 4, s);
      }
    }
    cluster_sync_all();  // nobody leaves (and releases its shared memory) while a peer may still read it
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == kEpiWarps) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// fp32 -> bf16 (round to nearest even), 8 elements per thread
__global__ void cast_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n8, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n8) {
    float v[8];
    ld_global_v8(x + i * 8, v);
    uint4 o = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    *reinterpret_cast<uint4*>(y + i * 8) = o;
  } else if (i == n8) {
    for (long long e = n8 * 8; e < n; ++e) y[e] = __float2bfloat16_rn(x[e]);
  }
}
// rows x cols with leading dimensions (views)
__global__ void cast_bf16_rows_kernel(const float* __restrict__ x, int ldx, __nv_bfloat16* __restrict__ y, int ldy, int rows, int cols) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)rows * cols) return;
  const int r = (int)(i / cols), c = (int)(i - (long long)r * cols);
  y[(size_t)r * ldy + c] = __float2bfloat16_rn(x[(size_t)r * ldx + c]);
}

// Tile width and k-slices of a product.  Clusters of 8 / 4 / 2 CTAs of this kernel (one CTA per SM) fit 15 / 33 / 74 at a time on the
// 148 SMs (measured for gemm_tc_kernel, same shared-memory footprint): a skinny product takes the combination that puts the most CTAs
// to work in one wave.
inline void choose_config(int M, int N, int K, int& bn, int& splits, int bk = kBK16) {
  const int num_kb = hulc_cdiv(K, bk), tiles_m = hulc_cdiv(M, kBM);
  static const int kClusters[3] = {8, 4, 2}, kCap[3] = {15, 33, 74};
  splits = 1;
  bn = N > 64 ? 128 : 64;
  // wide tiles where they still fill the machine: N = 256 instructions read 96 B/clk of shared memory instead of 128 (the limit)
  if (N >= 256 && tiles_m * hulc_cdiv(N, 256) >= 96) { bn = 256; return; }
  if (tiles_m * hulc_cdiv(N, bn) * 2 > kNumSMs) return;
  int best = tiles_m * hulc_cdiv(N, bn);
  for (int b = 128; b >= 64; b >>= 1) {
    if (b == 128 && N <= 64) continue;
    const int tiles = tiles_m * hulc_cdiv(N, b);
    for (int i = 0; i < 3; ++i) {
      const int c = kClusters[i];
      if (tiles <= kCap[i] && num_kb >= 2 * c && tiles * c > best) { best = tiles * c; bn = b; splits = c; }
    }
  }
}

// fp32 operand (consumed as tf32): 32 floats per k-block row; MN-major tiles are 32 x 32 boxes under the ATOM_32B flavour of the 128-byte swizzle
int operand_map_f32(CUtensorMap* m, const float* p, int rows, int K, int ld, bool mn_major, int box_rows) {
  if (mn_major) {  // stored K x rows
    const uint64_t dims[2] = {(uint64_t)rows, (uint64_t)K};
    const uint64_t strides[1] = {(uint64_t)ld * 4};
    const uint32_t box[2] = {32, 32};
    return tma::make_map(m, p, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  }
  const uint64_t dims[2] = {(uint64_t)K, (uint64_t)rows};
  const uint64_t strides[1] = {(uint64_t)ld * 4};
  const uint32_t box[2] = {32, (uint32_t)box_rows};
  return tma::make_map(m, p, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

int operand_map(CUtensorMap* m, const __nv_bfloat16* p, int rows, int K, int ld, bool mn_major, int box_rows) {
  if (mn_major) {  // stored K x rows
    const uint64_t dims[2] = {(uint64_t)rows, (uint64_t)K};
    const uint64_t strides[1] = {(uint64_t)ld * 2};
    const uint32_t box[2] = {64, 64};
    return tma::make_map(m, p, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, nullptr, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
  }
  const uint64_t dims[2] = {(uint64_t)K, (uint64_t)rows};
  const uint64_t strides[1] = {(uint64_t)ld * 2};
  const uint32_t box[2] = {64, (uint32_t)box_rows};
  return tma::make_map(m, p, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, nullptr, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
}

template <int BN, int CLUSTER, bool A_MN, bool B_MN, bool TF = false, bool P3 = false>
int launch_one(const CUtensorMap& ma, const CUtensorMap& mb, const BfEpilogue& ep, int tiles, int kbps, cudaStream_t st) {
  using Cfg = GCfg<BN, P3>;
  auto kfn = gemm_bf16_kernel<BN, CLUSTER, A_MN, B_MN, TF, P3>;
  HULC_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem));
  if (CLUSTER == 1) {
    HULC_LAUNCH(kfn, dim3(min(kNumSMs, tiles)), dim3(Cfg::kThreads), Cfg::kSmem, st, ma, mb, ep, tiles, kbps);
    HULC_RETURN_LAST();
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(tiles);  // one work item per CTA; consecutive CTAs = the k-slices of one output tile = one cluster
  cfg.blockDim = dim3(Cfg::kThreads);
  cfg.dynamicSmemBytes = Cfg::kSmem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CLUSTER; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  ++g_hulc_launches;
  HULC_TRY(cudaLaunchKernelEx(&cfg, kfn, ma, mb, ep, tiles, kbps));
  HULC_RETURN_LAST();
}

template <int BN, bool A_MN, bool B_MN, bool TF = false, typename T = __nv_bfloat16, bool P3 = false>
int launch(const T* A, const T* B, int M, int N, int K, int lda, int ldb, BfEpilogue ep, cudaStream_t st) {
  CUtensorMap ma, mb;
  if constexpr (TF) {
    if (operand_map_f32(&ma, A, M, K, lda, A_MN, kBM) != 0 || operand_map_f32(&mb, B, N, K, ldb, B_MN, BN) != 0) return (int)cudaErrorInvalidValue;
  } else {
    if (operand_map(&ma, A, M, K, lda, A_MN, kBM) != 0 || operand_map(&mb, B, N, K, ldb, B_MN, BN) != 0) return (int)cudaErrorInvalidValue;
  }
  const int tiles_m = hulc_cdiv(M, kBM), tiles_n = hulc_cdiv(N, BN);
  const int kbps = hulc_cdiv(hulc_cdiv(K, Elem<TF>::kBK), ep.splits);
  const int tiles = tiles_m * tiles_n * ep.splits;
  ep.BN = BN; ep.tiles_n = tiles_n;
  if constexpr (BN <= 128) {
    switch (ep.splits) {
      case 8: return launch_one<BN, 8, A_MN, B_MN, TF, P3>(ma, mb, ep, tiles, kbps, st);
      case 4: return launch_one<BN, 4, A_MN, B_MN, TF, P3>(ma, mb, ep, tiles, kbps, st);
      case 2: return launch_one<BN, 2, A_MN, B_MN, TF, P3>(ma, mb, ep, tiles, kbps, st);
      default: break;
    }
  }
  return launch_one<BN, 1, A_MN, B_MN, TF, P3>(ma, mb, ep, tiles, kbps, st);
}

template <int BN, bool TF = false, typename T = __nv_bfloat16, bool P3 = false>
int dispatch_layout(const T* A, const T* B, int M, int N, int K, int lda, int ldb, int transA, int transB, const BfEpilogue& ep, cudaStream_t st) {
  // op(A) is M x K: stored M x K (transA = 0: K-major) or K x M (transA = 1: MN-major).  op(B)^T is N x K: B stored N x K
  // (transB = 1, the torch Linear weight: K-major) or K x N (transB = 0: MN-major).
  if (!transA && transB) return launch<BN, false, false, TF, T, P3>(A, B, M, N, K, lda, ldb, ep, st);
  if (!transA && !transB) return launch<BN, false, true, TF, T, P3>(A, B, M, N, K, lda, ldb, ep, st);
  if (transA && transB) return launch<BN, true, false, TF, T, P3>(A, B, M, N, K, lda, ldb, ep, st);
  return launch<BN, true, true, TF, T, P3>(A, B, M, N, K, lda, ldb, ep, st);
}

}  // namespace

// fp32 operands consumed as tf32 (one pass), TMA-fed: the kernel behind hulc_gemm_tc's 1-pass products (gemm_tc.cu).  Same contract as hulc_gemm_tc
// (dropout applied in the epilogue with hulc_apply_dropout_rows' element indexing); cudaErrorNotSupported = operands the TMA cannot address.
int hulc_gemm_tf32_tma(const float* A, const float* B, float* C, int M, int N, int K, int lda, int ldb, int ldc, int transA, int transB, float alpha, float beta,
                       const float* bias, const float* addend, int ldadd, int add_mod, int act, const float* gate, int ldg, float drop_p,
                       unsigned long long drop_seed, unsigned drop_site, const unsigned char* drop_keep, int passes, cudaStream_t st) {
  if ((reinterpret_cast<size_t>(A) & 15) || (reinterpret_cast<size_t>(B) & 15) || (lda & 3) || (ldb & 3)) return (int)cudaErrorNotSupported;
  BfEpilogue ep{};
  ep.C = C; ep.Cb = nullptr; ep.M = M; ep.N = N; ep.ldc = ldc; ep.ldcb = 0; ep.alpha = alpha; ep.beta = beta; ep.bias = bias;
  ep.addend = addend; ep.ldadd = ldadd; ep.add_mod = add_mod; ep.act = act; ep.gate = gate; ep.gate_b = nullptr; ep.ldg = ldg;
  ep.drop = make_drop(drop_p, drop_seed, drop_site, drop_keep);
  int bn;
  choose_config(M, N, K, bn, ep.splits, Elem<true>::kBK);
  if (passes == 3) {  // (hi + lo stages and two accumulators per buffer: tiles of at most 128 columns)
    if (bn >= 128) return dispatch_layout<128, true, float, true>(A, B, M, N, K, lda, ldb, transA, transB, ep, st);
    return dispatch_layout<64, true, float, true>(A, B, M, N, K, lda, ldb, transA, transB, ep, st);
  }
  if (bn == 256) return dispatch_layout<256, true, float>(A, B, M, N, K, lda, ldb, transA, transB, ep, st);
  if (bn == 128) return dispatch_layout<128, true, float>(A, B, M, N, K, lda, ldb, transA, transB, ep, st);
  return dispatch_layout<64, true, float>(A, B, M, N, K, lda, ldb, transA, transB, ep, st);
}

// y = bf16(x) for n contiguous elements (x 32-byte aligned, y 16-byte aligned) — the cast that makes an fp32 tensor a GEMM operand.
HULC_API int hulc_cast_bf16(const float* x, void* y, long long n, void* stream) {
  if (n <= 0) return 0;
  if ((reinterpret_cast<size_t>(x) & 31) || (reinterpret_cast<size_t>(y) & 15)) return (int)cudaErrorInvalidValue;
  const long long n8 = n / 8;
  HULC_LAUNCH(cast_bf16_kernel, dim3(hulc_cdiv(n8 + 1, 256)), dim3(256), 0, (cudaStream_t)stream, x, reinterpret_cast<__nv_bfloat16*>(y), n8, n);
  HULC_RETURN_LAST();
}
// same for a rows x cols view with leading dimensions
HULC_API int hulc_cast_bf16_rows(const float* x, int ldx, void* y, int ldy, int rows, int cols, void* stream) {
  if (rows <= 0 || cols <= 0) return 0;
  HULC_LAUNCH(cast_bf16_rows_kernel, dim3(hulc_cdiv((long long)rows * cols, 256)), dim3(256), 0, (cudaStream_t)stream, x, ldx, reinterpret_cast<__nv_bfloat16*>(y), ldy, rows, cols);
  HULC_RETURN_LAST();
}

// C / Cb = epi(alpha * op(A) @ op(B)) with bf16 A, B (see include/hulc_b200.h).  Operand requirements (TMA): 16-byte aligned bases,
// lda % 8 == ldb % 8 == 0; otherwise cudaErrorInvalidValue.
HULC_API int hulc_gemm_bf16(const void* A, const void* B, float* C, void* Cb, int M, int N, int K, int lda, int ldb, int ldc, int ldcb, int transA, int transB,
                            float alpha, float beta, const float* bias, const float* addend, int ldadd, int add_mod, int act, const float* gate, const void* gate_bf16,
                            int ldg, float drop_p, unsigned long long drop_seed, unsigned drop_site, const unsigned char* drop_keep, void* stream) {
  if (M <= 0 || N <= 0) return 0;
  if (K <= 0 || !A || !B || (!C && !Cb) || (beta != 0.f && !C) || (gate && gate_bf16)) return (int)cudaErrorInvalidValue;
  if ((reinterpret_cast<size_t>(A) & 15) || (reinterpret_cast<size_t>(B) & 15) || (lda & 7) || (ldb & 7)) return (int)cudaErrorInvalidValue;
  BfEpilogue ep{};
  ep.C = C; ep.Cb = reinterpret_cast<__nv_bfloat16*>(Cb); ep.M = M; ep.N = N; ep.ldc = ldc; ep.ldcb = ldcb; ep.alpha = alpha; ep.beta = beta; ep.bias = bias;
  ep.addend = addend; ep.ldadd = ldadd; ep.add_mod = add_mod; ep.act = act; ep.gate = gate; ep.gate_b = reinterpret_cast<const __nv_bfloat16*>(gate_bf16); ep.ldg = ldg;
  ep.drop = make_drop(drop_p, drop_seed, drop_site, drop_keep);
  cudaStream_t st = (cudaStream_t)stream;
  int bn;
  choose_config(M, N, K, bn, ep.splits);
  const __nv_bfloat16* a = reinterpret_cast<const __nv_bfloat16*>(A);
  const __nv_bfloat16* b = reinterpret_cast<const __nv_bfloat16*>(B);
  if (bn == 256) return dispatch_layout<256>(a, b, M, N, K, lda, ldb, transA, transB, ep, st);
  if (bn == 128) return dispatch_layout<128>(a, b, M, N, K, lda, ldb, transA, transB, ep, st);
  return dispatch_layout<64>(a, b, M, N, K, lda, ldb, transA, transB, ep, st);
}
