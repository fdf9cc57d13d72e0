// tc_pipeline.cuh — the warp-specialised tcgen05 main loop shared by the dense GEMM (gemm_tc.cu) and the implicit-GEMM
// convolutions (conv_tc.cu).
//
//   D[128 x BN] (TMEM, fp32) (+)= A_tile[128 x BK] * B_tile[BN x BK]^T  over the k-blocks of a work item   (kind::tf32)
//
// One persistent CTA per SM walks the work items ("tiles").  Roles:
//   warps 0-3        epilogue: tcgen05.ld the finished accumulator (warp w owns TMEM lanes 32w..32w+31 = tile rows), apply
//                    the functor, store to global.  Two accumulator buffers in TMEM let the epilogue of item i overlap the
//                    main loop of item i+1.
//   warp  4          allocates TMEM; one lane issues every tcgen05.mma and the tcgen05.commit that recycles the stage.
//   warps 5-12        producers: every producer thread owns a fixed set of 16-byte chunks of the stage tiles (so the im2col
//                    row decode happens once per tile per thread) and fills them with cp.async straight into the swizzled
//                    UMMA layout, no register staging.  Each thread keeps kLag+1 cp.async groups in flight; when the group
//                    of k-block j-kLag has landed it fences (fence.proxy.async) and its warp arrives on that stage's
//                    mbarrier (8 arrivals complete a stage).
// Operands are gathered by cp.async rather than TMA because the A operands on this path are im2col views of NHWC / NCHW
// activations (per-row base addresses, zero-filled halo taps) and the dense GEMMs reuse the same machinery.
//
// fp32 operands are consumed as tf32 (the tensor core ignores the low 13 mantissa bits).  In the 3-pass mode (SPLIT) each
// producer thread, once its copies of a k-block have landed, reads its own chunks back from shared memory and writes the
// residual x - tf32(x) into a second ("lo") tile; the MMA warp then issues lo*hi + hi*lo (into a side accumulator) and
// hi*hi: 3xTF32, fp32-level accuracy without any extra global traffic.
#pragma once
#include "tc_common.cuh"

namespace tc {

constexpr int kEpiWarps = 4;
constexpr int kProdWarps = 8;
constexpr int kProdThreads = kProdWarps * 32;
constexpr int kBM = 128;

template <int BN, bool SPLIT, int BK = kBK, int KPS = 1>
struct PipeCfg {
  static constexpr int kATile = kBM * BK * 4;
  static constexpr int kBTile = BN * BK * 4;
  static constexpr int kSubBytes = (SPLIT ? 2 : 1) * (kATile + kBTile);  // one BK-wide k-block: [A hi][A lo][B hi][B lo]
  // KPS k-blocks share one stage (one full/empty barrier round trip, one wait and one commit of the MMA thread per KPS*4 MMAs):
  // the per-round cost of the issuing thread (~200 cycles: try_wait + commit) is what bounds small-N products
  static constexpr int kStageBytes = KPS * kSubBytes;
  static constexpr int kStagesRaw = (200 * 1024) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kLag = kStages - 1 < 3 ? kStages - 1 : 3;  // cp.async groups a producer thread leaves in flight
  static constexpr int kThreads = (kEpiWarps + 1 + kProdWarps) * 32;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*alignment slack*/ + 256 /*barriers*/;
  // two accumulator buffers; the 3-pass mode keeps the small cross terms (lo*hi + hi*lo) in their own accumulator so the
  // tensor core's truncating fp32 accumulation touches the main sum once per k-step instead of three times
  static constexpr int kAccCols = (SPLIT ? 2 : 1) * BN;
  static constexpr int kTmemCols = 2 * kAccCols < 32 ? 32 : 2 * kAccCols;
  static_assert(kStages >= 2, "at least two stages");
};

struct PipeBarriers {
  uint64_t full[8];
  uint64_t empty[8];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

// ---- thread-block cluster helpers (split-K over a cluster, reduced through distributed shared memory) ------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {  // every thread of every CTA of the cluster
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 16 bytes from the shared memory of CTA `rank` of this cluster, at the same offset as local address `addr`
__device__ __forceinline__ float4 ld_dsmem16(uint32_t addr, uint32_t rank) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(addr), "r"(rank));
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(ra) : "memory");
  return v;
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  // src-size 0 zero-fills the 16 bytes (out-of-range rows, k tails, halo taps)
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(valid ? 16 : 0) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// lo[off] = x - tf32(x) for one 16-byte chunk already resident at hi[off] (shared-memory addresses)
__device__ __forceinline__ void lo_chunk(uint32_t hi, uint32_t lo) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(hi));
  auto res = [](float f) { return f - __uint_as_float(__float_as_uint(f) & 0xFFFFE000u); };
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(lo), "f"(res(v.x)), "f"(res(v.y)), "f"(res(v.z)), "f"(res(v.w)) : "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Loader interface (called by all kProdThreads producer threads, ptid = 0..255):
//   static constexpr bool kMNMajor;                          // tile layout / descriptor flavour
//   void start_tile(int tile, int ptid);
//   void issue(int kb, uint32_t dst, int ptid);              // cp.async this thread's 16-byte chunks of the [rows x BK] tile
//   void split(uint32_t hi, uint32_t lo, int ptid);          // (3-pass) residuals of the same chunks, smem -> smem
// Epilogue interface (called by the 128 epilogue threads):  void operator()(int tile, int row, int col0, const float* v32)
//
// CLUSTER > 1 (split-K): the CLUSTER consecutive CTAs of a thread-block cluster hold the k-slices of ONE output tile (tile
// index = output tile * CLUSTER + k-slice = blockIdx.x, exactly one work item per CTA).  Each CTA parks its partial
// accumulator in its own shared memory; after a cluster barrier CTA r sums rows r, r+CLUSTER, ... of all the partials
// through distributed shared memory in a fixed order and hands them to the epilogue
//   int rows_valid(int tile) const;  void store4(int tile, int row, int col, float4 v) const;
// — no partials in global memory, no second launch, bit-reproducible.
template <int BN, bool SPLIT, int BK, int CLUSTER, class ALoader, class BLoader, class Epilogue, int KPS = 1>
__device__ __forceinline__ void run_pipeline(ALoader& al, BLoader& bl, const Epilogue& ep, int num_tiles, int num_kb) {
  using Cfg = PipeCfg<BN, SPLIT, BK, KPS>;
  const int num_rounds = (num_kb + KPS - 1) / KPS;  // a ragged last round reads k-blocks past num_kb: loaders zero-fill them
  constexpr int S = Cfg::kStages;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  PipeBarriers* bars = reinterpret_cast<PipeBarriers*>(smem + S * Cfg::kStageBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) trace(0);  // kernel entry

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&bars->full[s], kProdWarps);
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
  if (warp == 0) trace(1);  // barriers initialised, TMEM allocated

  if (warp < kEpiWarps) {
    // ================================ epilogue ================================
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int a = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&bars->tmem_full[a], aphase);
      tc_fence_after_sync();
      if (warp == 0) trace(4);  // accumulator ready
      if constexpr (CLUSTER > 1) {
        // the partial tile is parked in the stage buffers below: meet the producer and MMA warps (which are done: one work item per CTA,
        // and its accumulator is complete) at a CTA barrier first.  The mbarrier chain full -> tcgen05.commit -> tmem_full already orders
        // the producers' last stage writes before this point; the barrier makes that ordering explicit (and visible to racecheck).
        asm volatile("bar.sync 1, %0;" ::"n"(Cfg::kThreads) : "memory");
      }
      const int row = warp * 32 + lane;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(a * Cfg::kAccCols + c0), r);
        if (SPLIT) {
          uint32_t r2[32];
          tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(a * Cfg::kAccCols + BN + c0), r2);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) + __uint_as_float(r2[i]));
        } else {
          tmem_ld_wait();
        }
        if constexpr (CLUSTER > 1) {
          // park the partial tile in shared memory (the stage buffers are idle: every MMA of this CTA has completed);
          // rows are padded by 4 floats so the per-row 16-byte stores of a warp spread over all banks
          float* prow = reinterpret_cast<float*>(smem) + (size_t)row * (BN + 4) + c0;
#pragma unroll
          for (int q = 0; q < 32; q += 4)
            *reinterpret_cast<float4*>(prow + q) = make_float4(__uint_as_float(r[q]), __uint_as_float(r[q + 1]), __uint_as_float(r[q + 2]), __uint_as_float(r[q + 3]));
        } else {
          ep(tile, row, c0, reinterpret_cast<const float*>(r));
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->tmem_empty[a]);
    }
  } else if (warp == kEpiWarps) {
    // ================================ MMA issuer ================================
    // One thread runs the whole loop.  Descriptors are formed once (for stage 0, k-step 0) and advanced by adding the
    // byte offset >> 4 to the start-address field: every tcgen05.mma costs its thread a handful of integer instructions.
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(kBM, BN, ALoader::kMNMajor, BLoader::kMNMajor);
      constexpr uint32_t kAStep = (ALoader::kMNMajor ? 1024u : 32u) >> 4, kBStep = (BLoader::kMNMajor ? 1024u : 32u) >> 4;  // per k-step of 8
      const uint32_t s0 = smem_u32(smem);
      const uint64_t a_hi0 = make_desc<ALoader::kMNMajor, kBM, BK>(s0, 0);
      const uint64_t b_hi0 = make_desc<BLoader::kMNMajor, BN, BK>(s0 + (SPLIT ? 2 : 1) * Cfg::kATile, 0);
      int j = 0;  // stage-round sequence number of this CTA
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int a = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&bars->tmem_empty[a], aphase ^ 1);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + (uint32_t)(a * Cfg::kAccCols);
        const uint32_t d_small = d_tmem + BN;
        for (int r = 0; r < num_rounds; ++r, ++j) {
          const int stage = j % S;
          const uint32_t phase = (j / S) & 1;
          mbar_wait(&bars->full[stage], phase);
          tc_fence_after_sync();
          if (r == 0) trace(2);  // first stage landed
          if (r == num_rounds - 1) trace(3);  // last stage landed
          const uint32_t soff = (uint32_t)(stage * Cfg::kStageBytes) >> 4;
#pragma unroll
          for (int u = 0; u < KPS; ++u) {
#pragma unroll
            for (int k = 0; k < BK / 8; ++k) {
              const uint32_t first = (uint32_t)((r | u | k) != 0);
              const uint64_t da_hi = a_hi0 + (soff + ((uint32_t)(u * Cfg::kSubBytes) >> 4) + k * kAStep);
              const uint64_t db_hi = b_hi0 + (soff + ((uint32_t)(u * Cfg::kSubBytes) >> 4) + k * kBStep);
              if (SPLIT) {
                const uint64_t da_lo = da_hi + ((uint32_t)Cfg::kATile >> 4), db_lo = db_hi + ((uint32_t)Cfg::kBTile >> 4);
                umma_tf32(d_small, da_lo, db_hi, idesc, first);
                umma_tf32(d_small, da_hi, db_lo, idesc, 1);
                umma_tf32(d_tmem, da_hi, db_hi, idesc, first);
              } else {
                umma_tf32(d_tmem, da_hi, db_hi, idesc, first);
              }
            }
          }
          umma_commit(&bars->empty[stage]);                             // stage reusable once these MMAs have read it
          if (r == num_rounds - 1) umma_commit(&bars->tmem_full[a]);  // accumulator complete
        }
      }
    }
  } else {
    // ================================ producers ================================
    constexpr int L = Cfg::kLag;
    const int ptid = threadIdx.x - (kEpiWarps + 1) * 32;
    int j = 0;  // k-block sequence number of this CTA
    auto publish = [&](int jj) {  // the cp.async group of stage-round jj has landed: hand this warp's share to the MMA warp
      if constexpr (SPLIT) {
#pragma unroll
        for (int u = 0; u < KPS; ++u) {
          const uint32_t hi = smem_u32(smem + (jj % S) * Cfg::kStageBytes + u * Cfg::kSubBytes);
          al.split(hi, hi + Cfg::kATile, ptid);
          bl.split(hi + 2 * Cfg::kATile, hi + 2 * Cfg::kATile + Cfg::kBTile, ptid);
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->full[jj % S]);
    };
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      al.start_tile(tile, ptid);
      bl.start_tile(tile, ptid);
      for (int r = 0; r < num_rounds; ++r, ++j) {
        const int stage = j % S;
        mbar_wait(&bars->empty[stage], ((j / S) & 1) ^ 1);
#pragma unroll
        for (int u = 0; u < KPS; ++u) {
          const uint32_t a_hi = smem_u32(smem + stage * Cfg::kStageBytes + u * Cfg::kSubBytes);
          const uint32_t b_hi = a_hi + (SPLIT ? 2 : 1) * Cfg::kATile;
          al.issue(r * KPS + u, a_hi, ptid);
          bl.issue(r * KPS + u, b_hi, ptid);
        }
        cp_async_commit();
        if (j >= L) {
          cp_async_wait<L>();
          publish(j - L);
        }
      }
    }
    cp_async_wait<0>();
    for (int jj = j > L ? j - L : 0; jj < j; ++jj) publish(jj);
  }
  if constexpr (CLUSTER > 1) {
    if (warp >= kEpiWarps && (int)blockIdx.x < num_tiles) asm volatile("bar.sync 1, %0;" ::"n"(Cfg::kThreads) : "memory");  // see the epilogue
  }

  if (warp == 0) trace(5);  // epilogue warp 0 done with its role loop
  if constexpr (CLUSTER > 1) {
    static_assert(kBM * (BN + 4) * 4 <= Cfg::kStages * Cfg::kStageBytes, "partial tile must fit in the stage buffers");
    cluster_sync_all();  // all partial tiles of the cluster are in place
    if (warp == 0) trace(6);
    if (warp < kEpiWarps && (int)blockIdx.x < num_tiles) {
      const int tile = blockIdx.x;
      const uint32_t rank = cluster_ctarank();
      const int rows = ep.rows_valid(tile);
      constexpr int CQ = BN / 4;  // 16-byte column chunks per row
      const int e = threadIdx.x;  // 0..127
      const uint32_t base = smem_u32(smem);
      for (int idx = e; idx < kBM * CQ; idx += kEpiWarps * 32) {
        const int rr = idx / CQ, cq = idx - rr * CQ;
        const int row = rr * CLUSTER + (int)rank;  // this CTA's share of the tile's rows
        if (row >= rows) break;
        const uint32_t off = base + (uint32_t)(row * (BN + 4) + cq * 4) * 4u;
        float4 p[CLUSTER];
#pragma unroll
        for (int q = 0; q < CLUSTER; ++q) p[q] = ld_dsmem16(off, (uint32_t)q);
        float4 s = p[0];
#pragma unroll
        for (int q = 1; q < CLUSTER; ++q) { s.x += p[q].x; s.y += p[q].y; s.z += p[q].z; s.w += p[q].w; }
        ep.store4(tile, row, cq * 4, s);
      }
    }
    if (warp == 0) trace(7);
    cluster_sync_all();  // nobody leaves (and releases its shared memory) while a peer may still read it
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) trace(8);
  if (warp == kEpiWarps) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ---- dense operand loaders ----------------------------------------------------------------------------------------------
// Logical operand: ROWS x K (rows = M or N index).  Requirements (checked by the host wrappers): 16-byte aligned base,
// ld % 4 == 0, K % 4 == 0 (K-major) / rows % 4 == 0 (MN-major), so every 16-byte chunk is entirely valid or entirely zero.
// Work items: tile -> ((tm, tn), k-split).

// Producer threads own a FIXED set of 16-byte chunks of every stage tile (chunk q = ptid + 256 i), so everything that does not
// depend on the k-block — row pointers, bounds, swizzled destination offsets — is computed once per work item in start_tile and
// issue() is one pointer bump, one predicate and one cp.async per chunk.  (The producers are bound by the latency of their own
// instruction stream: the first version re-derived row / column / address per chunk per k-block, ~150 instructions per warp.)

// stored rows x K row-major (K contiguous): K-major tile [ROWS][BK], 128-byte rows, SWIZZLE_128B
template <int ROWS>
struct KMajorLoader {
  static constexpr bool kMNMajor = false;
  static constexpr int kChunks = ROWS * 8 / kProdThreads;  // per thread: rows (ptid >> 3) + 32 i, column chunk ptid & 7
  const float* base;
  int rows, K, ld, tiles_n, is_n, splits, kb_per_split;
  int row0, kb0;
  const float* ptr[kChunks];  // (row_i, first column of this thread's chunk in k-block 0 of the work item)
  int kcol;                   // that column
  uint32_t okmask;
  __device__ __forceinline__ void start_tile(int tile, int ptid) {
    const int t2 = tile / splits;
    kb0 = (tile - t2 * splits) * kb_per_split;
    row0 = (is_n ? t2 % tiles_n : t2 / tiles_n) * ROWS;
    kcol = kb0 * kBK + (ptid & 7) * 4;
    okmask = 0;
#pragma unroll
    for (int i = 0; i < kChunks; ++i) {
      const int row = row0 + (ptid >> 3) + 32 * i;
      const bool ok = row < rows;
      okmask |= (ok ? 1u : 0u) << i;
      ptr[i] = base + (size_t)(ok ? row : 0) * ld + kcol;
    }
  }
  __device__ __forceinline__ void issue(int kb, uint32_t dst, int ptid) const {
    const bool kok = kcol + kb * kBK < K;
    const int r = ptid >> 3, c = ptid & 7;
#pragma unroll
    for (int i = 0; i < kChunks; ++i) {
      const bool ok = kok && ((okmask >> i) & 1u);
      cp_async16(dst + swz(r + 32 * i, c), ok ? (const void*)(ptr[i] + kb * kBK) : (const void*)base, ok);
    }
  }
  __device__ __forceinline__ void split(uint32_t hi, uint32_t lo, int ptid) const {
#pragma unroll
    for (int q = ptid; q < ROWS * 8; q += kProdThreads) lo_chunk(hi + swz(q >> 3, q & 7), lo + swz(q >> 3, q & 7));
  }
};

// stored K x rows row-major (rows contiguous): MN-major tile [ROWS/32 groups][BK k-rows][32 elements], 128-byte rows,
// SWIZZLE_128B_BASE32B over each 4 k-rows; groups BK*128 B apart
template <int ROWS>
struct MNMajorLoader {
  static constexpr bool kMNMajor = true;
  static constexpr int RQ = ROWS / 4;                            // 16-byte chunks per k-row
  static constexpr int kChunks = kBK * RQ / kProdThreads;        // per thread: column chunk ptid % RQ, k-rows ptid / RQ + (256 / RQ) i
  static constexpr int kKStep = kProdThreads / RQ;
  const float* base;
  int rows, K, ld, tiles_n, is_n, splits, kb_per_split;
  int row0, kb0;
  const float* ptr;  // (k-row kb0 * 32 + ptid / RQ, column row0 + 4 (ptid % RQ))
  int krow;
  bool colok;
  __device__ __forceinline__ void start_tile(int tile, int ptid) {
    const int t2 = tile / splits;
    kb0 = (tile - t2 * splits) * kb_per_split;
    row0 = (is_n ? t2 % tiles_n : t2 / tiles_n) * ROWS;
    const int col = row0 + (ptid % RQ) * 4;
    colok = col < rows;
    krow = kb0 * kBK + ptid / RQ;
    ptr = base + (size_t)krow * ld + (colok ? col : 0);
  }
  static __device__ __forceinline__ uint32_t offset(int q) {
    const int kk = q / RQ, r = (q % RQ) * 4;
    return (uint32_t)(r >> 5) * (kBK * kRowBytes) + swz32(kk, (r & 31) >> 2);
  }
  __device__ __forceinline__ void issue(int kb, uint32_t dst, int ptid) const {
    const float* p = ptr + (size_t)kb * kBK * ld;
#pragma unroll
    for (int i = 0; i < kChunks; ++i) {
      const bool ok = colok && (krow + kb * kBK + kKStep * i < K);
      cp_async16(dst + offset(ptid + kProdThreads * i), ok ? (const void*)(p + (size_t)(kKStep * i) * ld) : (const void*)base, ok);
    }
  }
  __device__ __forceinline__ void split(uint32_t hi, uint32_t lo, int ptid) const {
#pragma unroll
    for (int q = ptid; q < kBK * (ROWS / 4); q += kProdThreads) lo_chunk(hi + offset(q), lo + offset(q));
  }
};

}  // namespace tc
