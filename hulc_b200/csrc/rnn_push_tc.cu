// rnn_push_tc.cu — the whole Elman recurrence (all S dependent steps of one layer / direction of torch.nn.RNN, forward or BPTT;
// decoders/utils/rnn.py:5-14, plan_encoders/plan_recognition_net.py:27-42) in ONE persistent launch — second generation of
// rnn_tc.cu, built around what the first one's per-step timeline showed: a step is a dependency chain, and every hop on it counts.
//
//   forward :  out_s = act(add_s + prev_s * W^T)          prev_s = h_{t-1},  add_s = x_t W_ih^T + b          (transW = 0)
//   backward:  out_s = (add_s + prev_s * W) * act'(gate_s) prev_s = dpre_{t+1}, add_s = dL/dh_t, gate_s = h_t  (transW = 1)
//
// 32 clusters x 4 CTAs = 128 CTAs, one per SM, all co-resident (a B200 holds 33 four-CTA clusters of a 1-CTA-per-SM kernel but only 15
// eight-CTA ones — scripts/micro/cluster_occ.cu — so the split over K stays 4-way).  A CTA keeps ITS block of W_hh (or W_hh^T) in shared
// memory, in the UMMA K-major layout, for all S steps, as the A operand (M = features) of D[features x batch rows] = W_blk x prev^T:
//   bf16 (kind::f16): 128 features x K-slice 512 (128 KB) x 32 batch rows — two clusters share a feature tile and split the batch, which
//                     halves what a CTA pulls from L2 per step (32 KB) and gives the full-rate M = 128 instruction shape (32 MMAs);
//   tf32            :  64 features x K-slice 512 (128 KB) x 64 batch rows (64 MMAs, 128 KB of state per CTA and step).
// Per step:
//   * producers (16 warps): the state of the previous step is exchanged through global memory (L2) and every element is its own
//     "ready" flag: all slots start as a sentinel (a NaN pattern no arithmetic produces), a producer thread re-reads its 16-byte pieces
//     (ld.relaxed.gpu, all in flight at once) until none holds the sentinel, then stores them into the swizzled B tile; every 128-byte
//     k-block has its own mbarrier, so the MMAs start when the first k-block has landed;
//   * one or two threads issue the MMAs (even / odd k-blocks into separate TMEM accumulators) and commit;
//   * epilogue (4 warps, thread = feature): the partial tile goes from TMEM straight to its owners — CTA j of the cluster reduces a
//     quarter of the batch rows — by st.async into the owner's shared memory, each store completing bytes on the owner's mbarrier
//     (a one-way push: no cluster barrier, no remote-load round trip); the owner sums the 4 partials in a fixed order, applies the
//     epilogue (addend / gate prefetched while the MMAs run) and writes the new state with relaxed gpu-scope stores — in the operand
//     type for the next step's producers, plus (bf16) the fp32 result the rest of the training step reads.
// Skew between CTAs is at most one step (a CTA's step s + 2 transitively needs every CTA's step s), so two receive buffers and two
// mbarriers alternate; no credits are needed.
#include "common.cuh"
#include "tc_pipeline.cuh"
#include <cuda_bf16.h>

namespace {

using namespace tc;

constexpr int kH = 2048;                      // hidden size this kernel is built for
constexpr int kCl = 4;                        // CTAs per cluster = K-slices = owners of a quarter of the cluster's batch rows
constexpr int kKS = kH / kCl;                 // K-slice per CTA
constexpr int kProd = 16;                     // producer warps
constexpr unsigned kSent32 = 0xFFFFFFFFu;     // "not written yet": fp32 word / pair of bf16
constexpr int kRecvBytes = 16 * 1024;         // one receive buffer: [src 4][piece][feature][16 B] = 4 partial quarter-tiles of fp32

template <typename E, int ISSUERS>
struct Cfg {
  static constexpr bool kBf = sizeof(E) == 2;
  static constexpr int kFT = kBf ? 128 : 64;               // features per CTA (MMA M)
  static constexpr int kNB = kBf ? 32 : 64;                // batch rows per CTA (MMA N)
  static constexpr int kBSplit = 64 / kNB;                 // clusters sharing a feature tile (each takes kNB batch rows)
  static constexpr int kRO = kNB / kCl;                    // batch rows an owner reduces: 8 / 16
  static constexpr int kPP = kRO / 4;                      // 16-byte pieces per (source, feature) in an owner's receive buffer: 2 / 4
  static constexpr int kEPB = kRowBytes / (int)sizeof(E);  // elements per 128-byte k-block row: 64 / 32
  static constexpr int kNKB = kKS / kEPB;                  // k-blocks per step: 8 / 16
  static constexpr int kStages = kBf ? kNKB : kNKB / 2;    // B-operand stages (tf32: a ring of half a step — shared memory budget)
  static constexpr int kWTile = kFT * kRowBytes;           // one k-block of the resident weights: 16 / 8 KB
  static constexpr int kWBytes = kNKB * kWTile;            // 128 KB
  static constexpr int kBTile = kNB * kRowBytes;           // 4 / 8 KB
  static constexpr int kWarpsPerKB = kProd / kNKB;         // producer warps per k-block: 2 / 1
  static constexpr int kRowsPerWarp = kNB / kWarpsPerKB;   // 16 / 64
  static constexpr int kPieces = kRowsPerWarp / 4;         // 16-byte pieces per producer thread and step: 4 / 16
  static constexpr int kThreads = (kEpiWarps + ISSUERS + kProd) * 32;
  static constexpr int kTmemCols = ISSUERS * kNB < 32 ? 32 : ISSUERS * kNB;
  static constexpr int kGrid = (kH / kFT) * kBSplit * kCl;  // 128
  static constexpr int kSmem = kWBytes + kStages * kBTile + 2 * kRecvBytes + 512 + 1024;
  static_assert(kCl * kPP * kFT * 16 == kRecvBytes, "receive buffer");
  static_assert(kSmem <= 227 * 1024, "shared memory budget");
  static_assert(kNKB % ISSUERS == 0 && kStages % ISSUERS == 0, "an issuer's k-blocks are kb = me, me + ISSUERS, ...");
};

struct PushParams {
  const void* W; int ldw;                     // E
  const void* x; long long x_step; int ldx;   // exchange buffer, E: slot s is read by step s, slot s + 1 written by it
  float* out; long long out_step; int ldo;    // fp32 result (bf16 kernel; null: none).  tf32: the exchange buffer IS the result
  const float* add; long long add_step; int ldadd;
  const float* gate; long long gate_step; int ldg;
  int act, B, S, poll;
};

struct PBars {
  uint64_t full[16];   // k-block kb of the step staged (kWarpsPerKB arrivals)
  uint64_t empty[16];  // stage read by the MMAs (ring only)
  uint64_t recv[2];    // the 4 partial quarter-tiles of a step have landed (transaction bytes)
  uint64_t tmem_full, tmem_empty;
  uint32_t tmem_base;
};

#ifdef HULC_RNN_TRACE
__device__ long long g_push_trace[64 * 16];
__device__ __forceinline__ void ptrace(int step, int slot) {
  if (blockIdx.x == 0 && step < 64) g_push_trace[step * 16 + slot] = clock64();
}
#else
__device__ __forceinline__ void ptrace(int, int) {}
#endif

__device__ __forceinline__ uint4 ld_relaxed16(const void* p) {  // gpu-scope relaxed: served by L2, never by a stale L1 line
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_f32(float* p, float v) { asm volatile("st.relaxed.gpu.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }
__device__ __forceinline__ void st_relaxed_b16(__nv_bfloat16* p, __nv_bfloat16 v) {
  asm volatile("st.relaxed.gpu.global.b16 [%0], %1;" ::"l"(p), "h"(*reinterpret_cast<unsigned short*>(&v)) : "memory");
}
__device__ __forceinline__ void st_shared16u(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(addr), "r"(rank));
  return ra;
}
// 16 bytes into a peer's shared memory; the peer's mbarrier `rbar` (same CTA as `raddr`) is credited 16 transaction bytes on arrival
__device__ __forceinline__ void st_async16(uint32_t raddr, uint32_t rbar, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(raddr), "r"(a), "r"(b), "r"(c),
               "r"(d), "r"(rbar)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 32 lanes x 16 columns of fp32 from TMEM
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
        "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
template <typename E>
__device__ __forceinline__ bool piece_ready(const uint4& u) {
  if (sizeof(E) == 4) return u.x != kSent32 && u.y != kSent32 && u.z != kSent32 && u.w != kSent32;
  // bf16: no half-word may be 0xFFFF
  return (__vcmpeq2(u.x, kSent32) | __vcmpeq2(u.y, kSent32) | __vcmpeq2(u.z, kSent32) | __vcmpeq2(u.w, kSent32)) == 0u;
}
__device__ __forceinline__ uint4 round_piece_tf32(uint4 u) {
  return make_uint4(__float_as_uint(to_tf32(__uint_as_float(u.x))), __float_as_uint(to_tf32(__uint_as_float(u.y))),
                    __float_as_uint(to_tf32(__uint_as_float(u.z))), __float_as_uint(to_tf32(__uint_as_float(u.w))));
}
__device__ __forceinline__ void umma_any(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate, bool f16) {
  if (f16) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(adesc),
                 "l"(bdesc), "r"(idesc), "r"(accumulate)
                 : "memory");
  } else {
    umma_tf32(tmem_d, adesc, bdesc, idesc, accumulate);
  }
}
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

template <typename E, bool TRANSW, int ISSUERS>
__global__ void __launch_bounds__(Cfg<E, ISSUERS>::kThreads, 1) rnn_push_kernel(PushParams p) {
  using C = Cfg<E, ISSUERS>;
  constexpr bool kBf = C::kBf;
  constexpr int FT = C::kFT, NB = C::kNB, RO = C::kRO, PP = C::kPP;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* w_smem = smem;
  unsigned char* b_smem = smem + C::kWBytes;
  unsigned char* recv = b_smem + C::kStages * C::kBTile;
  PBars* bars = reinterpret_cast<PBars*>(recv + 2 * kRecvBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ci = blockIdx.x / kCl;                   // cluster index
  const int rank = (int)cluster_ctarank();           // K-slice and owner index
  const int tile = ci / C::kBSplit, bh = ci - tile * C::kBSplit;
  const int f0 = tile * FT, k0 = rank * kKS, rb = bh * NB;  // first feature, first k, first batch row of this CTA
  const E* Wg = reinterpret_cast<const E*>(p.W);
  const E* xg = reinterpret_cast<const E*>(p.x);

  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) {
      mbar_init(&bars->full[i], C::kWarpsPerKB);
      mbar_init(&bars->empty[i], 1);
    }
    mbar_init(&bars->recv[0], 1);
    mbar_init(&bars->recv[1], 1);
    mbar_init(&bars->tmem_full, ISSUERS);
    mbar_init(&bars->tmem_empty, kEpiWarps);
    fence_barrier_init();
    mbar_arrive_expect(&bars->recv[0], kRecvBytes);  // armed for steps 0 and 1
    mbar_arrive_expect(&bars->recv[1], kRecvBytes);
  }
  if (warp == kEpiWarps) tmem_alloc(&bars->tmem_base, C::kTmemCols);
  // the B stages start as zeros (batch rows >= B are never written)
  for (int q = threadIdx.x; q < C::kStages * C::kBTile / 16; q += C::kThreads) st_shared16u(smem_u32(b_smem) + q * 16, make_uint4(0u, 0u, 0u, 0u));
  // the resident weight block A[m = feature][k], K-major SWIZZLE_128B tiles of [FT rows][128 B], one per k-block
  {
    const uint32_t wb = smem_u32(w_smem);
    constexpr int EP16 = 16 / (int)sizeof(E);  // elements per 16-byte chunk
    if (!TRANSW) {  // A[m][k] = W[f0 + m][k0 + k]: rows are K-contiguous
      for (int q = threadIdx.x; q < C::kNKB * FT * 8; q += C::kThreads) {
        const int kb = q / (FT * 8), qq = q - kb * (FT * 8);
        const int m = qq >> 3, c = qq & 7;
        uint4 v = __ldg(reinterpret_cast<const uint4*>(Wg + (size_t)(f0 + m) * p.ldw + k0 + kb * C::kEPB + c * EP16));
        if (!kBf) v = round_piece_tf32(v);
        st_shared16u(wb + kb * C::kWTile + swz(m, c), v);
      }
    } else {        // A[m][k] = W[k0 + k][f0 + m]: read rows of W along m (coalesced), scatter the elements into the K-major rows
      for (int q = threadIdx.x; q < kKS * (FT / EP16); q += C::kThreads) {
        const int k = q / (FT / EP16), m0 = (q - k * (FT / EP16)) * EP16;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(Wg + (size_t)(k0 + k) * p.ldw + f0 + m0));
        const int kb = k / C::kEPB, kk = k - kb * C::kEPB;
        const uint32_t base = wb + kb * C::kWTile;
        if (kBf) {
          const unsigned short* h = reinterpret_cast<const unsigned short*>(&v);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int m = m0 + e;
            asm volatile("st.shared.b16 [%0], %1;" ::"r"(base + swz(m, kk >> 3) + (kk & 7) * 2), "h"(h[e]) : "memory");
          }
        } else {
          const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int m = m0 + e;
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(base + swz(m, kk >> 2) + (kk & 3) * 4), "f"(to_tf32(__uint_as_float(w4[e]))) : "memory");
          }
        }
      }
    }
  }
  fence_proxy_async();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_d = bars->tmem_base;
  cluster_sync_all();  // every CTA of the cluster has initialised and armed its barriers before anybody pushes

  if (warp < kEpiWarps) {
    // ================================ epilogue: push partials, reduce the own rows ================================
    // M = 128: thread t reads TMEM lane t = feature t.  M = 64: the accumulator rows 16 w .. 16 w + 15 sit in lanes 32 w .. 32 w + 15, so
    // lanes 0-15 of a warp push; the reduction of the owner's 16 rows is split over both half-warps (8 rows each).
    const int f = FT == 128 ? (int)threadIdx.x : 16 * warp + (lane & 15);
    const bool sender = FT == 128 || lane < 16;
    const int qb = FT == 128 ? 0 : 2 * (lane >> 4);        // first of this thread's two 16-byte pieces (4 rows each) in the reduction
    const int row0 = rb + rank * RO + qb * 4;              // the 8 batch rows this thread reduces
    const int col = f0 + f;
    const uint32_t recv_u = smem_u32(recv);
    const uint32_t my_slot = recv_u + (uint32_t)(rank * PP * FT + f) * 16u;  // [src = me][piece 0][feature f] of receive buffer 0 (same offset in every CTA)
    const uint32_t bar_u = smem_u32(&bars->recv[0]);
    for (int s = 0; s < p.S; ++s) {
      const int buf = s & 1;
      // operands of the epilogue that do not depend on this step's product: fetch them while the MMAs run
      float ad[8], gt[8];
      {
        const float* add = p.add + s * p.add_step;
        const float* gate = p.gate ? p.gate + s * p.gate_step : nullptr;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          ad[r] = gt[r] = 0.f;
          if (row0 + r < p.B) {
            ad[r] = __ldg(add + (size_t)(row0 + r) * p.ldadd + col);
            if (gate) gt[r] = __ldg(gate + (size_t)(row0 + r) * p.ldg + col);
          }
        }
      }
      mbar_wait(&bars->tmem_full, s & 1);
      tc_fence_after_sync();
      if (threadIdx.x == 0) ptrace(s, 4);
      // push: columns [RO j, RO j + RO) of this feature's row go to owner j, 16 columns per pass (register budget)
#pragma unroll
      for (int pass = 0; pass < NB / 16; ++pass) {
        uint32_t v[16];
        tmem_ld16(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)(pass * 16), v);
        if (ISSUERS == 2) {  // the second issuer's accumulator (odd k-blocks)
          uint32_t v2[16];
          tmem_ld16(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)(NB + pass * 16), v2);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(v2[i]));
        } else {
          tmem_ld_wait();
        }
        if (pass == NB / 16 - 1) {
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->tmem_empty);
        }
        if (sender) {
#pragma unroll
          for (int jj = 0; jj < 16 / RO; ++jj) {
            const uint32_t j = (uint32_t)(pass * (16 / RO) + jj);
            const uint32_t dst = mapa(my_slot, j) + (uint32_t)buf * kRecvBytes, bar = mapa(bar_u, j) + (uint32_t)buf * 8u;
#pragma unroll
            for (int q = 0; q < PP; ++q)
              st_async16(dst + (uint32_t)(q * FT) * 16u, bar, v[RO * jj + 4 * q], v[RO * jj + 4 * q + 1], v[RO * jj + 4 * q + 2], v[RO * jj + 4 * q + 3]);
          }
        }
      }
      if (threadIdx.x == 0) ptrace(s, 5);
      // reduce: the 4 partials of rows row0 .. row0 + 7, feature f, summed in a fixed order
      // (a plain try_wait: the transaction-count completion orders the st.async data before the phase flip, as for TMA traffic; an
      //  acquire at cluster scope makes ptxas emit CCTL.IVALL — an L1 flush per step that sent every later local / read-only load to L2)
      mbar_wait(&bars->recv[buf], (s >> 1) & 1);
      if (threadIdx.x == 0) {
        ptrace(s, 6);
        mbar_arrive_expect(&bars->recv[buf], kRecvBytes);  // arm the barrier for step s + 2 (nobody can be there before this step's result is out)
      }
      float o[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) o[r] = 0.f;
      const unsigned char* rbuf = recv + (size_t)buf * kRecvBytes + (size_t)(qb * FT + f) * 16;
#pragma unroll
      for (int src = 0; src < kCl; ++src) {
        const float4 a = *reinterpret_cast<const float4*>(rbuf + (size_t)(src * PP) * FT * 16);
        const float4 b = *reinterpret_cast<const float4*>(rbuf + (size_t)(src * PP + 1) * FT * 16);
        o[0] += a.x; o[1] += a.y; o[2] += a.z; o[3] += a.w;
        o[4] += b.x; o[5] += b.y; o[6] += b.z; o[7] += b.w;
      }
      E* xo = const_cast<E*>(xg) + (long long)(s + 1) * p.x_step + (long long)row0 * p.ldx + col;
      // (the activation / gate kind is uniform: select the loop once, not per element)
#pragma unroll
      for (int r = 0; r < 8; ++r) o[r] += ad[r];
      if ((p.act & 3) == 1) {
#pragma unroll
        for (int r = 0; r < 8; ++r) o[r] = fmaxf(o[r], 0.f);
      } else if ((p.act & 3) == 2) {
#pragma unroll
        for (int r = 0; r < 8; ++r) o[r] = tanhf(o[r]);
      }
      if (p.gate) {
        if (p.act & 4) {
#pragma unroll
          for (int r = 0; r < 8; ++r) o[r] *= 1.f - gt[r] * gt[r];
        } else {
#pragma unroll
          for (int r = 0; r < 8; ++r) o[r] = gt[r] > 0.f ? o[r] : 0.f;
        }
      }
      // the value is its own "ready" flag for the next step's producers: these stores go out first
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        if (row0 + r < p.B) {
          if (kBf) st_relaxed_b16(reinterpret_cast<__nv_bfloat16*>(xo) + (long long)r * p.ldx, __float2bfloat16_rn(o[r]));
          else st_relaxed_f32(reinterpret_cast<float*>(xo) + (long long)r * p.ldx, o[r]);
        }
      }
      if (p.out) {
        float* out = p.out + s * p.out_step + (long long)row0 * p.ldo + col;
#pragma unroll
        for (int r = 0; r < 8; ++r)
          if (row0 + r < p.B) out[(long long)r * p.ldo] = o[r];
      }
      if (threadIdx.x == 0) ptrace(s, 7);
    }
  } else if (warp < kEpiWarps + ISSUERS) {
    // ================================ MMA issuers: k-blocks me, me + ISSUERS, ... into accumulator `me` ================================
    const int me = warp - kEpiWarps;
    constexpr uint32_t idesc = kBf ? idesc_bf16(FT, NB) : make_idesc_tf32(FT, NB, false, false);
    const uint64_t a0 = make_smem_desc(smem_u32(w_smem), 16u, 1024u, 2u), b0 = make_smem_desc(smem_u32(b_smem), 16u, 1024u, 2u);
    const uint32_t acc = tmem_d + (uint32_t)(me * NB);
    if (lane == 0) {
      for (int s = 0; s < p.S; ++s) {
        mbar_wait(&bars->tmem_empty, (s & 1) ^ 1);
        tc_fence_after_sync();
#pragma unroll
        for (int kb = me; kb < C::kNKB; kb += ISSUERS) {
          mbar_wait(&bars->full[kb], s & 1);
          tc_fence_after_sync();
          if (kb == 0) ptrace(s, 2);
          const int st = kb % C::kStages;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_any(acc, a0 + (uint32_t)(kb * (C::kWTile >> 4) + k * 2), b0 + (uint32_t)(st * (C::kBTile >> 4) + k * 2), idesc, (uint32_t)(kb >= ISSUERS || k != 0), kBf);
          if (C::kStages < C::kNKB && kb < C::kNKB - C::kStages) umma_commit(&bars->empty[st]);  // the stage may be refilled with k-block kb + kStages
        }
        umma_commit(&bars->tmem_full);
        if (me == 0) ptrace(s, 3);
      }
    }
    __syncwarp();
  } else {
    // ================================ producers ================================
    const int pw = warp - (kEpiWarps + ISSUERS);   // 0..15
    const int kb = pw % C::kNKB, rg = pw / C::kNKB;
    const int st = kb % C::kStages;
    const int r0 = rg * C::kRowsPerWarp + (lane >> 3), c = lane & 7;  // tile rows r0 + 4 i, 16-byte chunk c of the k-block
    const uint32_t dst = smem_u32(b_smem) + st * C::kBTile;
    constexpr int EP16 = 16 / (int)sizeof(E);
    for (int s = 0; s < p.S; ++s) {
      // the stages still hold the operands of the previous step until all its MMAs have completed (a CTA may lag its suppliers by one step)
      if (s > 0) mbar_wait(&bars->tmem_full, (s - 1) & 1);
      const E* src = xg + (long long)s * p.x_step + (size_t)rb * p.ldx + k0 + kb * C::kEPB + c * EP16;
      uint4 u[C::kPieces];
      unsigned pending = 0u;
#pragma unroll
      for (int i = 0; i < C::kPieces; ++i) {
        u[i] = make_uint4(0u, 0u, 0u, 0u);
        if (rb + r0 + 4 * i < p.B) pending |= 1u << i;
      }
      if (pw == 0 && lane == 0) ptrace(s, 0);
      if (p.poll == 1 && (pending & 1u)) {
        // cheap wait first: every lane re-reads only its first piece until it has been written
        for (unsigned spins = 0;; ++spins) {
          u[0] = ld_relaxed16(src + (size_t)r0 * p.ldx);
          if (__all_sync(__activemask(), piece_ready<E>(u[0]))) break;
          if (spins > (1u << 22)) __trap();  // a protocol bug must surface as a launch failure, never as a hung GPU
        }
      }
      for (unsigned spins = 0; __any_sync(0xffffffffu, pending != 0u); ++spins) {
#pragma unroll
        for (int i = 0; i < C::kPieces; ++i)
          if ((pending >> i) & 1u) u[i] = ld_relaxed16(src + (size_t)(r0 + 4 * i) * p.ldx);
#pragma unroll
        for (int i = 0; i < C::kPieces; ++i)
          if (piece_ready<E>(u[i])) pending &= ~(1u << i);
        if (spins > (1u << 22)) __trap();
      }
      if (pw == 0 && lane == 0) ptrace(s, 1);
      // ring (tf32): the stage is shared with k-block kb - kStages of the same step
      if (C::kStages < C::kNKB && kb >= C::kStages) mbar_wait(&bars->empty[st], s & 1);
#pragma unroll
      for (int i = 0; i < C::kPieces; ++i)
        if (rb + r0 + 4 * i < p.B) st_shared16u(dst + swz(r0 + 4 * i, c), kBf ? u[i] : round_piece_tf32(u[i]));
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->full[kb]);
    }
  }

  cluster_sync_all();  // nobody leaves while a peer's pushes may still be in flight towards it
  tc_fence_before_sync();
  __syncthreads();
  if (warp == kEpiWarps) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_d, C::kTmemCols);
  }
}

// exchange buffer: slot 0 := the initial state (converted), slots 1..S := the sentinel
template <typename E>
__global__ void push_init_kernel(const float* prev0, int ldp, E* x, long long x_step, int ldx, int B, int S, int fill_only) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int q = (int)(i % (kH / 4));
  const long long rs = i / (kH / 4);
  const int row = (int)(rs % B), s = (int)(rs / B);
  if (s > S) return;
  E* dst = x + s * x_step + (long long)row * ldx + q * 4;
  if (s == 0) {
    if (fill_only) return;
    const float4 v = *reinterpret_cast<const float4*>(prev0 + (long long)row * ldp + q * 4);
    if (sizeof(E) == 2) {
      __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
      *reinterpret_cast<uint2*>(dst) = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
    } else {
      *reinterpret_cast<float4*>(dst) = v;
    }
  } else if (sizeof(E) == 2) {
    *reinterpret_cast<uint2*>(dst) = make_uint2(kSent32, kSent32);
  } else {
    *reinterpret_cast<uint4*>(dst) = make_uint4(kSent32, kSent32, kSent32, kSent32);
  }
}

template <typename E, bool TRANSW, int ISSUERS>
cudaError_t push_config(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, cudaStream_t st) {
  using C = Cfg<E, ISSUERS>;
  auto kfn = rnn_push_kernel<E, TRANSW, ISSUERS>;
  cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem);
  if (e != cudaSuccess) return e;
  cfg = cudaLaunchConfig_t{};
  cfg.gridDim = dim3(C::kGrid);
  cfg.blockDim = dim3(C::kThreads);
  cfg.dynamicSmemBytes = C::kSmem;
  cfg.stream = st;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaSuccess;
}

template <typename E, bool TRANSW, int ISSUERS>
int launch_push(const PushParams& p, cudaStream_t st) {
  using C = Cfg<E, ISSUERS>;
  static int max_clusters = -1;
  auto kfn = rnn_push_kernel<E, TRANSW, ISSUERS>;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  auto cfg_fn = push_config<E, TRANSW, ISSUERS>;
  HULC_TRY(cfg_fn(cfg, attr, st));
  if (max_clusters < 0) {  // the 128 CTAs wait on each other's results: they must all be resident at once
    int n = 0;
    HULC_TRY(cudaOccupancyMaxActiveClusters(&n, kfn, &cfg));
    max_clusters = n;
  }
  if (max_clusters < C::kGrid / kCl) return (int)cudaErrorLaunchOutOfResources;
  ++g_hulc_launches;
  HULC_TRY(cudaLaunchKernelEx(&cfg, kfn, p));
  HULC_RETURN_LAST();
}

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
int poll_mode() {
  static int mode = env_int("HULC_B200_RNN_POLL", 1);  // 1: wait on the first piece, then fetch all (less L2 polling traffic; measured 3-4 % faster than polling every piece)
  return mode;
}
int issuers() {
  static int n = env_int("HULC_B200_RNN_ISSUERS", 2);
  return n == 1 ? 1 : 2;
}

template <typename E>
int launch_any(const PushParams& p, int transW, cudaStream_t st) {
  if (issuers() == 2) return transW ? launch_push<E, true, 2>(p, st) : launch_push<E, false, 2>(p, st);
  return transW ? launch_push<E, true, 1>(p, st) : launch_push<E, false, 1>(p, st);
}

}  // namespace

#ifdef HULC_RNN_TRACE
HULC_API int hulc_rnn_push_trace_read(long long* host_out) {
  cudaDeviceSynchronize();
  return (int)cudaMemcpyFromSymbol(host_out, g_push_trace, sizeof(long long) * 64 * 16);
}
#endif

// tf32 variant behind hulc_rnn_tc_seq (rnn_tc.cu): same contract; cudaErrorLaunchOutOfResources = this launch is not for this kernel
// (the 32 clusters cannot be co-resident, or a single step whose input slot is not adjacent to its output slot).
int hulc_rnn_push_tf32(const float* W, int ldw, int transW, const float* prev0, long long prev_step, int ldp, float* out0, long long out_step, int ldo,
                       const float* add0, long long add_step, int ldadd, const float* gate0, long long gate_step, int ldg, int act, int B, int S,
                       cudaStream_t st) {
  if (prev0 + prev_step != out0 || ldp != ldo || (S > 1 && prev_step != out_step)) return (int)cudaErrorLaunchOutOfResources;
  HULC_LAUNCH(push_init_kernel<float>, dim3(hulc_cdiv((long long)(S + 1) * B * (kH / 4), 256)), dim3(256), 0, st, prev0, ldp, const_cast<float*>(prev0), prev_step, ldp, B,
              S, 1);
  PushParams p;
  p.W = W; p.ldw = ldw; p.x = prev0; p.x_step = prev_step; p.ldx = ldp; p.out = nullptr; p.out_step = 0; p.ldo = 0;
  p.add = add0; p.add_step = add_step; p.ldadd = ldadd; p.gate = gate0; p.gate_step = gate_step; p.ldg = ldg;
  p.act = act; p.B = B; p.S = S; p.poll = poll_mode();
  return launch_any<float>(p, transW, st);
}

// Diagnostics: how many 4-CTA clusters of the recurrence kernels this device can hold at once (32 are needed); out[0] bf16, out[1] tf32.
HULC_API int hulc_rnn_push_max_clusters(int* out) {
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  auto cfg16 = push_config<__nv_bfloat16, false, 2>;
  auto cfg32 = push_config<float, false, 2>;
  auto k16 = rnn_push_kernel<__nv_bfloat16, false, 2>;
  auto k32 = rnn_push_kernel<float, false, 2>;
  HULC_TRY(cfg16(cfg, attr, nullptr));
  HULC_TRY(cudaOccupancyMaxActiveClusters(&out[0], k16, &cfg));
  HULC_TRY(cfg32(cfg, attr, nullptr));
  HULC_TRY(cudaOccupancyMaxActiveClusters(&out[1], k32, &cfg));
  return 0;
}

// See include/hulc_b200.h.
HULC_API int hulc_rnn_seq_bf16(const void* W16, int ldw, int transW, const float* prev0, int ldp, void* x16, float* out0, long long out_step, int ldo,
                               const float* add0, long long add_step, int ldadd, const float* gate0, long long gate_step, int ldg, int act, int B, int H,
                               int S, void* stream) {
  if (S <= 0 || B <= 0) return 0;
  if (H != kH || B > 64 || !W16 || !prev0 || !x16 || !add0) return (int)cudaErrorInvalidValue;
  if ((reinterpret_cast<size_t>(W16) | reinterpret_cast<size_t>(prev0) | reinterpret_cast<size_t>(x16)) & 15) return (int)cudaErrorInvalidValue;
  if ((ldw & 7) || (ldp & 3)) return (int)cudaErrorInvalidValue;
  cudaStream_t st = (cudaStream_t)stream;
  __nv_bfloat16* x = reinterpret_cast<__nv_bfloat16*>(x16);
  const long long x_step = (long long)B * kH;
  HULC_LAUNCH(push_init_kernel<__nv_bfloat16>, dim3(hulc_cdiv((long long)(S + 1) * B * (kH / 4), 256)), dim3(256), 0, st, prev0, ldp, x, x_step, kH, B, S, 0);
  PushParams p;
  p.W = W16; p.ldw = ldw; p.x = x; p.x_step = x_step; p.ldx = kH; p.out = out0; p.out_step = out_step; p.ldo = ldo;
  p.add = add0; p.add_step = add_step; p.ldadd = ldadd; p.gate = gate0; p.gate_step = gate_step; p.ldg = ldg;
  p.act = act; p.B = B; p.S = S; p.poll = poll_mode();
  return launch_any<__nv_bfloat16>(p, transW, st);
}
