// rnn_tc.cu — a whole Elman recurrence (all S dependent steps of one layer / direction of torch.nn.RNN, forward or BPTT;
// decoders/utils/rnn.py:5-14, plan_encoders/plan_recognition_net.py:27-42) in ONE persistent launch.
//
//   forward :  out_s = act(add_s + prev_s * W^T)          prev_s = h_{t-1},  add_s = x_t W_ih^T + b          (transW = 0)
//   backward:  out_s = (add_s + prev_s * W) * act'(gate_s) prev_s = dpre_{t+1}, add_s = dL/dh_t, gate_s = h_t  (transW = 1)
//
// Per step the product is [B <= 64] x [2048] x [2048]: far too small to fill the tensor cores and strictly sequential, so
// the cost of a step is latency.  Launching it as a GEMM per step re-reads W_hh (16.8 MB) from L2 and pays a launch, a
// pipeline fill and a drain 128 times per training step.  Here instead
//   * the grid is 32 clusters x 4 CTAs = 128 CTAs (one per SM, all co-resident).  Cluster i owns the 64 output features
//     [64 i, 64 i + 64); its CTA j owns the K-slice [512 j, 512 j + 512).  Each CTA loads ITS 64 x 512 block of W_hh once,
//     rounds it to tf32 (round-to-nearest; the tensor core itself would truncate) and keeps it in shared memory (128 KB)
//     in the UMMA operand layout for all S steps;
//   * per step sixteen producer warps — one per 32-wide k-block — each wait for the ONE feature tile of the previous step their
//     k-block comes from and fetch their B x 32 piece of the previous hidden state in ONE round trip: the output slots of all S steps
//     are pre-filled with a sentinel bit pattern (0xFFFFFFFF, a NaN no arithmetic produces) and a producer simply re-reads its 16-byte
//     pieces (ld.relaxed.gpu, every load in flight at once) until none of their words is the sentinel — the data is its own flag, so
//     there is no release fence, no arrival counter and no second dependent load (flag, then data) on the critical path; it rounds the
//     values to tf32 and stores them into the swizzled A tile; one thread issues 64 tcgen05.mma (kind::tf32, M = 64 batch rows, N = 64, fp32 accumulator in TMEM) with two
//     barrier waits and three commits per step; the CTA parks its partial tile in shared memory and after ONE cluster barrier
//     each CTA sums the four partials of its quarter of the rows through distributed shared memory in a fixed order, applies
//     the epilogue (addend and gate prefetched while the MMAs run) and writes h_t with relaxed gpu-scope stores.
// Measured with arrival counters (release increment per CTA, acquire poll, then the loads): 10.3 us per step — flag propagation 2.6,
// L2 load 1.1, stage + MMA 4.4, reduce + store + release 2.2 — against 25 us (forward, 3xTF32) / 17 us (backward) for the same step as a
// split-K cluster GEMM launch; the data-as-flag hand-off removes the flag propagation and the release.
// A single tf32 pass with round-to-nearest on both operands keeps the action logits within 0.3 of the parity tolerance
// (rtol 1e-3 / atol 1e-4) over the 32-step chain; accumulation, addend and activation are fp32.
#include "common.cuh"
#include "tc_pipeline.cuh"

namespace {

using namespace tc;

constexpr int kFT = 64;                       // output features per cluster (MMA N)
constexpr int kKS = 512;                      // K-slice per CTA
constexpr int kCl = 4;                        // CTAs per cluster = K-slices
constexpr int kH = kKS * kCl;                 // hidden size this kernel is built for
constexpr int kKB = kKS / kBK;                // k-blocks per step = producer warps (one k-block each)
constexpr int kBmax = 64;                     // batch rows = MMA M
constexpr int kAStageBytes = kBmax * kRowBytes;
constexpr int kAStages = 8;                   // stage = k-block & 7: warps w and w + 8 alternate on stage w
constexpr int kWTile = kFT * kRowBytes;       // one k-block of the resident weight block
constexpr int kWBytes = kKB * kWTile;         // 128 KB
constexpr int kParkBytes = kBmax * kFT * 4;   // parked partial tile, 16-byte chunks XOR-swizzled by row (no padding)
constexpr int kSmemBytes = kWBytes + kAStages * kAStageBytes + 2 * kParkBytes + 256 + 1024;
constexpr int kRnnProdWarps = kKB;
constexpr int kThreads = (kEpiWarps + 1 + kRnnProdWarps) * 32;
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

struct RnnParams {
  const float* W; int ldw;
  const float* prev; long long prev_step; int ldp;
  float* out; long long out_step; int ldo;
  const float* add; long long add_step; int ldadd;
  const float* gate; long long gate_step; int ldg;
  int act, B, S;
};
constexpr unsigned kSentinel = 0xFFFFFFFFu;  // "not written yet" (see the header); hulc_rnn_tc_seq fills every output slot with it before the launch

struct Bars {
  uint64_t full[2];   // k-blocks 0..7 / 8..15 of the step staged (8 producer warps each)
  uint64_t empty[2];  // stages 0..3 / 4..7 read by the MMAs of the first half (tcgen05.commit)
  uint64_t tmem_full, tmem_empty;
  uint32_t tmem_base;
};

// Optional per-step timeline of CTA 0 (development aid, -DHULC_RNN_TRACE): clock64 stamps, slot meanings in scripts/dbg_rnn_trace.py
#ifdef HULC_RNN_TRACE
__device__ long long g_rnn_trace[64 * 64];
__device__ __forceinline__ void rtrace(int step, int slot) {
  if (blockIdx.x == 0 && step < 64) g_rnn_trace[step * 64 + slot] = clock64();
}
#else
__device__ __forceinline__ void rtrace(int, int) {}
#endif

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint4 ld_relaxed16(const float* p) {  // gpu-scope relaxed: served by L2, never by a stale L1 line
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed16(float* p, float a, float b, float c, float d) {
  asm volatile("st.relaxed.gpu.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float4 rna4(float4 v) { return make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w)); }
__device__ __forceinline__ void st_shared16(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// byte offset of 16-byte chunk q of an MN-major [kFT features x 32 k] tile: q = kk * (kFT/4) + feature/4
__device__ __forceinline__ uint32_t mn_offset(int q) {
  constexpr int RQ = kFT / 4;
  const int kk = q / RQ, r = (q % RQ) * 4;
  return (uint32_t)(r >> 5) * (kBK * kRowBytes) + swz32(kk, (r & 31) >> 2);
}
// byte offset of 16-byte chunk cq (0..15) of row `row` in a parked tile
__device__ __forceinline__ uint32_t park_off(int row, int cq) { return (uint32_t)(row * kFT * 4 + ((cq ^ (row & 15)) << 4)); }

template <bool TRANSW>
__global__ void __launch_bounds__(kThreads, 1) rnn_seq_kernel(RnnParams p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* w_smem = smem;
  unsigned char* a_smem = smem + kWBytes;
  unsigned char* park = a_smem + kAStages * kAStageBytes;
  Bars* bars = reinterpret_cast<Bars*>(park + 2 * kParkBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x / kCl;          // feature tile = cluster index
  const int slice = blockIdx.x - tile * kCl;  // K-slice = rank in the cluster
  const int f0 = tile * kFT, k0 = slice * kKS;

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars->full[s], kAStages);
      mbar_init(&bars->empty[s], 1);
    }
    mbar_init(&bars->tmem_full, 1);
    mbar_init(&bars->tmem_empty, kEpiWarps);
    fence_barrier_init();
  }
  if (warp == kEpiWarps) tmem_alloc(&bars->tmem_base, 64);
  // the resident weight block, rounded to tf32, in the UMMA layout of the B operand
  {
    const uint32_t wb = smem_u32(w_smem);
    for (int q = threadIdx.x; q < kKB * kFT * 8; q += kThreads) {
      const int kb = q / (kFT * 8), qq = q - kb * (kFT * 8);
      const float* src;
      uint32_t dst;
      if (!TRANSW) {  // B[n = feature][k]: rows of W are K-contiguous -> K-major tile
        const int n = qq >> 3, c = qq & 7;
        src = p.W + (size_t)(f0 + n) * p.ldw + k0 + kb * kBK + c * 4;
        dst = wb + kb * kWTile + swz(n, c);
      } else {        // B[n = input feature][k = output feature] = W[k][n]: n contiguous -> MN-major tile
        constexpr int RQ = kFT / 4;
        const int kk = qq / RQ, r = (qq % RQ) * 4;
        src = p.W + (size_t)(k0 + kb * kBK + kk) * p.ldw + f0 + r;
        dst = wb + kb * kWTile + mn_offset(qq);
      }
      st_shared16(dst, rna4(__ldg(reinterpret_cast<const float4*>(src))));
    }
  }
  fence_proxy_async();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_d = bars->tmem_base;
  cluster_sync_all();  // every CTA of the cluster has initialised its barriers

  if (warp < kEpiWarps) {
    // ================================ epilogue + cluster reduction ================================
    // M = 64 accumulators sit in TMEM lanes 32w .. 32w+15 of sub-partition w (rows 16w .. 16w+15)
    const int e = threadIdx.x;  // 0..127
    constexpr int kItems = (kBmax / kCl) * (kFT / 4) / (kEpiWarps * 32);  // (row, 16-byte chunk) items of the reduction per thread
    for (int s = 0; s < p.S; ++s) {
      unsigned char* pk = park + (size_t)(s & 1) * kParkBytes;
      // operands of the epilogue that do not depend on this step's product: fetch them while the MMAs run
      const float* add = p.add + s * p.add_step;
      const float* gate = p.gate ? p.gate + s * p.gate_step : nullptr;
      float4 ad[kItems], gt[kItems];
#pragma unroll
      for (int it = 0; it < kItems; ++it) {
        const int idx = e + it * (kEpiWarps * 32);
        const int rr = idx / (kFT / 4), cq = idx - rr * (kFT / 4);
        const int row = rr * kCl + slice, col = f0 + cq * 4;
        ad[it] = gt[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < p.B) {
          ad[it] = __ldg(reinterpret_cast<const float4*>(add + (size_t)row * p.ldadd + col));
          if (gate) gt[it] = __ldg(reinterpret_cast<const float4*>(gate + (size_t)row * p.ldg + col));
        }
      }
      mbar_wait(&bars->tmem_full, s & 1);
      tc_fence_after_sync();
      if (e == 0) rtrace(s, 6);
      {
        const int row = warp * 16 + (lane & 15);
#pragma unroll
        for (int c0 = 0; c0 < kFT; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
          tmem_ld_wait();
          if (lane < 16) {
#pragma unroll
            for (int q = 0; q < 32; q += 4)
              *reinterpret_cast<float4*>(pk + park_off(row, (c0 + q) >> 2)) =
                  make_float4(__uint_as_float(r[q]), __uint_as_float(r[q + 1]), __uint_as_float(r[q + 2]), __uint_as_float(r[q + 3]));
          }
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->tmem_empty);
      if (e == 0) rtrace(s, 7);
      cluster_arrive();
      cluster_wait();  // the four partial tiles of this step are parked
      if (e == 0) rtrace(s, 8);
      float* out = p.out + s * p.out_step;
      const uint32_t base = smem_u32(pk);
      float4 part[kItems][kCl];
#pragma unroll
      for (int it = 0; it < kItems; ++it) {
        const int idx = e + it * (kEpiWarps * 32);
        const int rr = idx / (kFT / 4), cq = idx - rr * (kFT / 4);
        const int row = rr * kCl + slice;  // this CTA's quarter of the rows
#pragma unroll
        for (int q = 0; q < kCl; ++q) part[it][q] = ld_dsmem16(base + park_off(row, cq), (uint32_t)q);
      }
#pragma unroll
      for (int it = 0; it < kItems; ++it) {
        const int idx = e + it * (kEpiWarps * 32);
        const int rr = idx / (kFT / 4), cq = idx - rr * (kFT / 4);
        const int row = rr * kCl + slice, col = f0 + cq * 4;
        if (row < p.B) {
          const float4* a = part[it];
          float o[4] = {((a[0].x + a[1].x) + a[2].x) + a[3].x + ad[it].x, ((a[0].y + a[1].y) + a[2].y) + a[3].y + ad[it].y,
                        ((a[0].z + a[1].z) + a[2].z) + a[3].z + ad[it].z, ((a[0].w + a[1].w) + a[2].w) + a[3].w + ad[it].w};
          if ((p.act & 3) == 1) {
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = fmaxf(o[j], 0.f);
          } else if ((p.act & 3) == 2) {
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = tanhf(o[j]);
          }
          if (gate) {
            const float g[4] = {gt[it].x, gt[it].y, gt[it].z, gt[it].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = (p.act & 4) ? o[j] * (1.f - g[j] * g[j]) : (g[j] > 0.f ? o[j] : 0.f);
          }
          // the values are their own "ready" flag for the next step's producers (32-bit words are written atomically)
          st_relaxed16(out + (size_t)row * p.ldo + col, o[0], o[1], o[2], o[3]);
        }
      }
      if (e == 0) rtrace(s, 9);
    }
  } else if (warp == kEpiWarps) {
    // ================================ MMA issuer ================================
    // The issuing thread is the bottleneck of a product this small (a tcgen05.mma costs it ~50-60 cycles, a barrier wait
    // ~100, a commit ~60): two waits and three commits per step, descriptors advanced by immediates.
    constexpr uint32_t idesc = make_idesc_tf32(kBmax, kFT, false, TRANSW);
    constexpr uint32_t kBStep = (TRANSW ? 1024u : 32u) >> 4;
    const uint64_t a0 = make_desc<false, kBmax, kBK>(smem_u32(a_smem), 0), b0 = make_desc<TRANSW, kFT, kBK>(smem_u32(w_smem), 0);
    for (int s = 0; s < p.S; ++s) {
      cluster_arrive();
      if (lane == 0) {
        mbar_wait(&bars->tmem_empty, (s & 1) ^ 1);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          mbar_wait(&bars->full[half], s & 1);
          tc_fence_after_sync();
          if (half == 0) rtrace(s, 4);
#pragma unroll
          for (int q = 0; q < kAStages; ++q) {
            const int kb = half * kAStages + q;
#pragma unroll
            for (int k = 0; k < kBK / 8; ++k)
              umma_tf32(tmem_d, a0 + (uint32_t)(q * (kAStageBytes >> 4) + k * 2), b0 + (uint32_t)(kb * (kWTile >> 4) + k * kBStep), idesc, (kb | k) != 0);
            if (half == 0 && (q & 3) == 3) umma_commit(&bars->empty[q >> 2]);  // stages 0-3 / 4-7 may be refilled
          }
        }
        umma_commit(&bars->tmem_full);
        rtrace(s, 5);
      }
      __syncwarp();
      cluster_wait();
    }
  } else {
    // ================================ producers: warp w stages k-block w of every step ================================
    const int kb = warp - (kEpiWarps + 1);  // 0..15
    const int stage = kb & (kAStages - 1);
    const int r0 = lane >> 3, c = lane & 7;  // rows r0 + 4 i (i = 0..15), 16-byte chunk c
    const uint32_t dst = smem_u32(a_smem) + stage * kAStageBytes;
    for (int s = 0; s < p.S; ++s) {
      cluster_arrive();
      if (kb == 0 && lane == 0) rtrace(s, 0);
      // this lane's 16 pieces of the previous step's output (step 0: the initial state): re-read until every word has been written
      const float* src = p.prev + s * p.prev_step + k0 + kb * kBK + c * 4;
      uint4 u[16];
      unsigned pending = 0u;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        u[i] = make_uint4(0u, 0u, 0u, 0u);
        if (r0 + 4 * i < p.B) pending |= 1u << i;
      }
#ifdef HULC_RNN_TRACE
      if (p.act & 64) pending = 0u;
#endif
      // (1) cheap wait: every lane re-reads only its FIRST piece (rows 0..3 of the block: one row from each of the four CTAs that write this
      //     feature tile) until it has been written — 512 bytes per warp and round, not the whole 8 KB block;
      // (2) then all 16 pieces, every load in flight at once; any word still holding the sentinel sends that piece round again (rare).
      if (pending & 1u) {
        for (unsigned spins = 0;; ++spins) {
          u[0] = ld_relaxed16(src + (size_t)r0 * p.ldp);
          const bool ok = u[0].x != kSentinel && u[0].y != kSentinel && u[0].z != kSentinel && u[0].w != kSentinel;
          if (__all_sync(__activemask(), ok)) break;
          if (spins > (1u << 22)) __trap();  // a protocol bug (or a genuine 0xFFFFFFFF NaN in the state) must surface as a launch failure, never as a hung GPU
        }
      }
      for (unsigned spins = 0; __any_sync(0xffffffffu, pending != 0u); ++spins) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if ((pending >> i) & 1u) u[i] = ld_relaxed16(src + (size_t)(r0 + 4 * i) * p.ldp);
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (u[i].x != kSentinel && u[i].y != kSentinel && u[i].z != kSentinel && u[i].w != kSentinel) pending &= ~(1u << i);
        if (spins > (1u << 22)) __trap();
      }
      if (kb == 0 && lane == 0) rtrace(s, 1);
      // stage w is shared by warps w (first half of the step) and w + 8 (second half): the second user waits until the MMAs of
      // the first half have read it; the first user needs no wait — the cluster barrier that ended the previous step is
      // only passed once every MMA of that step has completed
      if (kb >= kAStages) mbar_wait(&bars->empty[(stage >> 2) & 1], s & 1);
      if (kb == 0 && lane == 0) rtrace(s, 2);
      if (lane == 0) rtrace(s, 48 + kb);
#pragma unroll
      for (int i = 0; i < 16; ++i)
        st_shared16(dst + swz(r0 + 4 * i, c), rna4(make_float4(__uint_as_float(u[i].x), __uint_as_float(u[i].y), __uint_as_float(u[i].z), __uint_as_float(u[i].w))));
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->full[kb >> 3]);
      if (kb == 0 && lane == 0) rtrace(s, 3);
      if (lane == 0) rtrace(s, 32 + kb);
      if (kb == 15 && lane == 0) rtrace(s, 15);
      cluster_wait();
    }
  }

  cluster_sync_all();  // nobody leaves while a peer may still read its parked tile
  tc_fence_before_sync();
  __syncthreads();
  if (warp == kEpiWarps) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_d, 64);
  }
}

// every output slot (S steps x B rows x kH columns) := the sentinel
__global__ void fill_sentinel_kernel(float* out0, long long out_step, int ldo, int B, int S) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int q = (int)(i % (kH / 4));
  const long long rs = i / (kH / 4);
  const int row = (int)(rs % B), s = (int)(rs / B);
  if (s >= S) return;
  *reinterpret_cast<uint4*>(out0 + s * out_step + (long long)row * ldo + q * 4) = make_uint4(kSentinel, kSentinel, kSentinel, kSentinel);
}

template <bool TRANSW>
int launch_seq(const RnnParams& p, cudaStream_t st) {
  auto kfn = rnn_seq_kernel<TRANSW>;
  HULC_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((kH / kFT) * kCl);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  ++g_hulc_launches;
  HULC_TRY(cudaLaunchKernelEx(&cfg, kfn, p));
  HULC_RETURN_LAST();
}

}  // namespace

#ifdef HULC_RNN_TRACE
HULC_API int hulc_rnn_trace_read(long long* host_out) {
  cudaDeviceSynchronize();
  return (int)cudaMemcpyFromSymbol(host_out, g_rnn_trace, sizeof(long long) * 64 * 64);
}
#endif

// rnn_push_tc.cu: the 8-CTA-cluster generation of this kernel (same contract); cudaErrorLaunchOutOfResources = cannot run here
int hulc_rnn_push_tf32(const float* W, int ldw, int transW, const float* prev0, long long prev_step, int ldp, float* out0, long long out_step, int ldo,
                       const float* add0, long long add_step, int ldadd, const float* gate0, long long gate_step, int ldg, int act, int B, int S,
                       cudaStream_t st);

// See include/hulc_b200.h.
HULC_API int hulc_rnn_tc_seq(const float* W, int ldw, int transW, const float* prev0, long long prev_step, int ldp, float* out0, long long out_step,
                             int ldo, const float* add0, long long add_step, int ldadd, const float* gate0, long long gate_step, int ldg, int act,
                             int B, int H, int S, float* workspace, size_t workspace_bytes, void* stream) {
  if (S <= 0 || B <= 0) return 0;
  if (H != kH || B > kBmax || !W || !prev0 || !out0 || !add0 || !workspace || workspace_bytes < (1024 + kH / kFT) * sizeof(float))
    return (int)cudaErrorInvalidValue;
  // every vector access is 16 bytes wide
  if ((reinterpret_cast<size_t>(W) | reinterpret_cast<size_t>(prev0) | reinterpret_cast<size_t>(out0) | reinterpret_cast<size_t>(add0) |
       reinterpret_cast<size_t>(gate0)) & 15)
    return (int)cudaErrorInvalidValue;
  if ((ldw | ldp | ldo | ldadd | ldg) & 3 || (prev_step | out_step | add_step | gate_step) & 3) return (int)cudaErrorInvalidValue;
  cudaStream_t st = (cudaStream_t)stream;
  {
    static int gen = -1;  // HULC_B200_RNN_GEN=2 selects the push kernel of rnn_push_tc.cu: with fp32 state it moves 128 KB per CTA and step and measured 11.4 us per step against 9.7 here
    if (gen < 0) {
      const char* e = getenv("HULC_B200_RNN_GEN");
      gen = e ? atoi(e) : 1;
    }
    if (gen >= 2) {
      const int rc = hulc_rnn_push_tf32(W, ldw, transW, prev0, prev_step, ldp, out0, out_step, ldo, add0, add_step, ldadd, gate0, gate_step, ldg, act, B, S, st);
      if (rc != (int)cudaErrorLaunchOutOfResources) return rc;
      (void)cudaGetLastError();
    }
  }
  // the 128 CTAs spin on each other's counters: they must all be resident at once
  static int max_clusters = -1;
  if (max_clusters < 0) {
    auto kfn = rnn_seq_kernel<false>;
    HULC_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((kH / kFT) * kCl); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = kSmemBytes;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kCl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = 0;
    HULC_TRY(cudaOccupancyMaxActiveClusters(&n, kfn, &cfg));
    max_clusters = n;
  }
  if (max_clusters < kH / kFT) return (int)cudaErrorLaunchOutOfResources;
  // the hand-off between steps is by address: step s reads what step s - 1 wrote
  if (S > 1 && (prev0 + prev_step != out0 || prev_step != out_step || ldp != ldo)) return (int)cudaErrorInvalidValue;
  HULC_LAUNCH(fill_sentinel_kernel, dim3(hulc_cdiv((long long)S * B * (kH / 4), 256)), dim3(256), 0, st, out0, out_step, ldo, B, S);
  RnnParams p;
  p.W = W; p.ldw = ldw; p.prev = prev0; p.prev_step = prev_step; p.ldp = ldp; p.out = out0; p.out_step = out_step; p.ldo = ldo;
  p.add = add0; p.add_step = add_step; p.ldadd = ldadd; p.gate = gate0; p.gate_step = gate_step; p.ldg = ldg;
  p.act = act; p.B = B; p.S = S;
  return transW ? launch_seq<true>(p, st) : launch_seq<false>(p, st);
}
