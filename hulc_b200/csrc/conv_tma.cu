// conv_tma.cu — the channels-last convolutions of the perceptual encoders (layers 2 and 3: 32 -> 64, k4, s2 and 64 -> 64,
// k3, s1; vision_network.py:39-43, vision_network_gripper.py:14-16), forward and data gradient, as implicit GEMMs whose A
// operand is delivered by TMA.
//
// With NHWC activations the im2col block of one kernel tap and 32 channels is a BOX of the activation tensor: output
// pixels (n, y0.., 0..OW-1) read input pixels (n, S*y + ky, S*x + kx), i.e. a 4-d box (32 ch, OW, RT, NB) walked with element
// strides (1, S, S, 1) from the corner (c0, kx, S*y0 + ky, n0).  One cp.async.bulk.tensor lands it in shared memory as
// [pixel][128 B] rows with the 128-byte swizzle — the K-major UMMA operand tile — and the halo of the data gradient is the
// TMA's out-of-bounds zero fill.  One thread issues the copies, the prepared weights [BN x K] stay resident in shared memory
// for the whole kernel, four warps run the epilogue.  tcgen05.mma (kind::tf32) is issued by FOUR threads, each taking every
// fourth k-block into its own TMEM accumulator (summed by the epilogue): a k-block of 4 MMAs costs its issuing thread ~600
// cycles (barrier wait, descriptor setup, commit) but the tensor pipe only ~256, so one issuer leaves the pipe idle half the
// time (scripts/micro/mma_rate.cu: 640 cycles per k-block with one issuer, 330 with two).  Compared with the
// cp.async gather of conv_tc.cu there is no per-thread address arithmetic at all (that kernel is bound by the latency of
// the ~150 instructions per warp per k-block its producers execute).
//
//   forward      : source = x,  output grid = (HO, WO), taps (+ky, +kx), B = w as [COUT][(ky, kx, ci)], epilogue bias + ReLU
//   data gradient: source = dY, output grid = the input pixels of one stride phase (py, px), taps (-jy, -jx),
//                  B = w as [CIN][(jy, jx, co)] for that phase, epilogue = ReLU mask of the activation that fed the layer
#include <cstdlib>

#include "common.cuh"
#include "tc_pipeline.cuh"
#include "tma.cuh"

namespace {

using namespace tc;

constexpr int kATileB = kBM * kRowBytes;  // 16 KB: one k-block of the A operand (128 pixels x 32 channels)
constexpr int kNI = 4;  // MMA-issuing threads (one warp each): every 4th k-block each, into an accumulator of its own
constexpr int kThr = (kEpiWarps + kNI + 1) * 32;
constexpr int kSmemBudget = 224 * 1024;
// HULC_B200_CONV_BAND=0 keeps the box-per-tap kernel for the stride-1 layer (A/B comparison)
const bool g_use_band = [] { const char* e = getenv("HULC_B200_CONV_BAND"); return !(e && e[0] == '0'); }();

struct TcParams {
  int N, OH, OW;       // output grid of this launch
  int NB, RT, TPF;     // frames / rows per tile, tiles per frame (NB == 1) — GEMM rows per tile = NB * RT * OW <= 128
  int S;               // element stride of the source walk (the layer's stride forward, 1 for the data gradient)
  int taps_x, cblocks, sign;  // k-block kb -> tap = kb / cblocks (ty = tap / taps_x, tx = tap % taps_x), channel block kb % cblocks
  const float* wprep;  // [BN][num_kb * 32] K-major
  float* out;          // NHWC, [N][out_H][out_W][BN]
  int out_H, out_W, o_mul, oy_add, ox_add;  // output pixel of grid point (y, x): (o_mul * y + oy_add, o_mul * x + ox_add)
  const float* bias;
  const float* gate;   // same geometry as out (data gradient), or null
  int relu;
};

struct TcBars {
  uint64_t full[8];
  uint64_t empty[8];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ void st_shared16(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <int BN, int NKB, int STAGES>
__global__ void __launch_bounds__(kThr, 1) conv_tma_kernel(const __grid_constant__ CUtensorMap smap, TcParams p, int num_tiles) {
  constexpr int kWTileB = BN * kRowBytes;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* w_smem = smem;
  unsigned char* a_smem = smem + NKB * kWTileB;
  TcBars* bars = reinterpret_cast<TcBars*>(a_smem + STAGES * kATileB);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bars->tmem_full[a], kNI);
      mbar_init(&bars->tmem_empty[a], kEpiWarps);
    }
    fence_barrier_init();
  }
  static_assert(2 * kNI * BN <= 512, "TMEM columns");
  if (warp == kEpiWarps) tmem_alloc(&bars->tmem_base, 2 * kNI * BN);
  // rows of the A stages beyond the pixels of a tile are never written by the TMA box: keep them finite
  for (int i = threadIdx.x; i < STAGES * kATileB / 16; i += kThr) reinterpret_cast<float4*>(a_smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  // resident weights: NKB K-major swizzled tiles [BN rows x 128 B]
  for (int q = threadIdx.x; q < NKB * BN * 8; q += kThr) {
    const int kb = q / (BN * 8), qq = q - kb * (BN * 8);
    const int n = qq >> 3, c = qq & 7;
    st_shared16(smem_u32(w_smem) + kb * kWTileB + swz(n, c), __ldg(reinterpret_cast<const float4*>(p.wprep + (size_t)n * (NKB * kBK) + kb * kBK + c * 4)));
  }
  fence_proxy_async();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;
  const int frame_px = p.RT * p.OW;

  if (warp < kEpiWarps) {
    // ================================ epilogue ================================
    const int r = warp * 32 + lane;
    const int nl = r / frame_px, rem = r - nl * frame_px;
    const int yl = rem / p.OW, xx = rem - yl * p.OW;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int a = it & 1;
      const int n0 = p.NB > 1 ? tile * p.NB : tile / p.TPF;
      const int y0 = p.NB > 1 ? 0 : (tile - n0 * p.TPF) * p.RT;
      const bool valid = nl < p.NB && n0 + nl < p.N && y0 + yl < p.OH;
      const size_t off = (((size_t)(n0 + nl) * p.out_H + (p.o_mul * (y0 + yl) + p.oy_add)) * p.out_W + (p.o_mul * xx + p.ox_add)) * BN;
      mbar_wait(&bars->tmem_full[a], (it >> 1) & 1);
      tc_fence_after_sync();
#pragma unroll
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(a * kNI * BN + c0), v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 1; i < kNI; ++i) {  // partial sums of the other issuers
          uint32_t u[32];
          tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)((a * kNI + i) * BN + c0), u);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 32; ++q) v[q] = __float_as_uint(__uint_as_float(v[q]) + __uint_as_float(u[q]));
        }
        if (c0 + 32 == BN) {  // the accumulator is in registers: hand it back before the stores
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->tmem_empty[a]);
        }
        if (valid) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            if (p.bias) {
              const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + j));
              o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
            }
            if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            if (p.gate) {
              const float4 g = __ldg(reinterpret_cast<const float4*>(p.gate + off + c0 + j));
              o.x = g.x > 0.f ? o.x : 0.f; o.y = g.y > 0.f ? o.y : 0.f; o.z = g.z > 0.f ? o.z : 0.f; o.w = g.w > 0.f ? o.w : 0.f;
            }
            *reinterpret_cast<float4*>(p.out + off + c0 + j) = o;
          }
        }
      }
    }
  } else if (warp < kEpiWarps + kNI) {
    // ================================ MMA issuers ================================
    if (lane == 0) {
      const int me = warp - kEpiWarps;
      constexpr uint32_t idesc = make_idesc_tf32(kBM, BN, false, false);
      const uint64_t a0 = make_desc<false, kBM, kBK>(smem_u32(a_smem), 0), b0 = make_desc<false, BN, kBK>(smem_u32(w_smem), 0);
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int a = it & 1;
        mbar_wait(&bars->tmem_empty[a], ((it >> 1) & 1) ^ 1);
        tc_fence_after_sync();
        const uint32_t d = tmem_base + (uint32_t)((a * kNI + me) * BN);
        // k-block j of the CTA's stage sequence belongs to issuer j % kNI; STAGES is a multiple of kNI, so every stage has ONE
        // consumer that sees each of its uses in order (a consumer that skipped uses could mistake an older phase of the
        // stage's barrier for the one it waits for)
        const int kb0 = ((me - it * NKB) % kNI + kNI) % kNI;
        for (int kb = kb0; kb < NKB; kb += kNI) {
          const int j = it * NKB + kb;
          const int stage = j % STAGES;
          mbar_wait(&bars->full[stage], (j / STAGES) & 1);
          tc_fence_after_sync();
          const uint64_t da = a0 + (uint32_t)(stage * (kATileB >> 4)), db = b0 + (uint32_t)(kb * (kWTileB >> 4));
#pragma unroll
          for (int k = 0; k < kBK / 8; ++k) umma_tf32(d, da + (uint32_t)(k * 2), db + (uint32_t)(k * 2), idesc, (uint32_t)(kb != kb0 || k != 0));
          umma_commit(&bars->empty[stage]);
        }
        umma_commit(&bars->tmem_full[a]);
      }
    }
  } else {
    // ================================ TMA producer ================================
    if (lane == 0) {
      tma::prefetch_map(&smap);
      const uint32_t box_bytes = (uint32_t)(p.NB * frame_px * kRowBytes);
      int j = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n0 = p.NB > 1 ? tile * p.NB : tile / p.TPF;
        const int y0 = p.NB > 1 ? 0 : (tile - n0 * p.TPF) * p.RT;
        int tap = 0, cb = 0;
        for (int kb = 0; kb < NKB; ++kb, ++j) {
          const int stage = j % STAGES;
          mbar_wait(&bars->empty[stage], ((j / STAGES) & 1) ^ 1);
          const int ty = tap / p.taps_x, tx = tap - ty * p.taps_x;
          tma::expect_tx(&bars->full[stage], box_bytes);
          tma::load_4d(smem_u32(a_smem) + stage * kATileB, &smap, &bars->full[stage], cb * kBK, p.sign * tx, p.S * y0 + p.sign * ty, n0);
          if (++cb == p.cblocks) { cb = 0; ++tap; }
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == kEpiWarps) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 2 * kNI * BN);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Shifted views of staged bands.
// For a stride-1 convolution the A tile of tap (ty, tx) is the A tile of tap (0, 0) moved by ty * PW + tx pixels along the
// flattened band, so a tile loads a band of PH source rows x PW columns x 32 channels ONCE ([pixel][128 B] rows, 128-byte
// swizzle keyed on the absolute shared-memory address) and every tap reads it through a descriptor whose start address is
// shifted by whole rows (verified on the B200, scripts/micro/desc_shift_probe.cu: arbitrary row shifts work with the
// descriptor's base-offset field left at zero).  GEMM rows are positions of the band's PW-wide grid (the extra columns per
// row are computed and dropped).  This removes the per-tap re-fetch (4x / 9x the activation through L2 -> SM) that bounds
// conv_tma_kernel.
//   3x3 stride 1 forward : one band per 32-channel block, origin (0, y0), 9 taps, shift = ty PW + tx
//   4x4 stride 2 forward : one band per input phase (py, px) — the TMA walks the activation with element stride 2 from
//                          (px, 2 y0 + py), which de-interleaves the phase plane for free — 2 x 2 taps (ky, kx) = (2 ty + py, 2 tx + px)
//   data gradients       : the band starts R - 1 rows / columns before the output pixel (out-of-bounds zero fill = the halo),
//                          shift = (R-1-ty) PW + (R-1-tx); for stride 2 one launch per output phase (R = 2), stride 1 R = 3
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kBandRows = 184;             // 127 + largest shift (2 * 25 + 2) + 1, rounded up to a multiple of 8
constexpr int kBandB = kBandRows * kRowBytes;
constexpr int kBandSlots = 3;
constexpr int kMaxBands = 4, kMaxTaps = 9;
// (An epilogue that transposes each warp's 32 x 32 chunk through shared memory so that global accesses are whole 128-byte rows
// was measured SLOWER than the row-owner stores used here — conv2 forward 0.22 -> 0.24 ms, the fused stride-2 gradient
// 0.56 -> 1.97 ms — and was dropped.)

struct BandParams {
  int N, OH, OWv;        // output grid rows; valid output columns (< PW)
  int PW, PH, RT, TPF;   // band pitch / rows, output rows per tile, tiles per frame
  int nbands, es, flip;  // bands per tile; element stride of the source walk; flip = 1 for a data gradient
  int bc[kMaxBands], bx[kMaxBands], by[kMaxBands];  // TMA corner of band b: (bc, bx, es * y0 + by, n)
  int wkb[kMaxBands * kMaxTaps];                    // weight k-block of (band, local tap)
  const void* wprep;     // [BN][NKB k-blocks of 128 bytes] K-major: 32 fp32 or 64 bf16 per k-block
  void* out;             // NHWC [N][out_H][out_W][BN] (fp32, or bf16 in the BF16 kernels); grid point (y, x) -> pixel (o_mul y + oy_add, o_mul x + ox_add)
  int out_H, out_W, o_mul, oy_add, ox_add;
  int scatter;           // 1: the BN = 128 columns are 4 stride phases x 32 channels; chunk ph goes to pixel (2 y + ph / 2, 2 x + ph % 2)
  const float* bias;
  const float* gate;
  const unsigned* gate_bits;  // ReLU sign mask of the gating activation, one word per (pixel, 32 channels); preferred over `gate`
  unsigned* relu_bits;        // forward: sign mask of the output, written next to it (lets the data gradient skip re-reading the activation)
  int relu;
};

struct BandBars {
  uint64_t full[kBandSlots];
  uint64_t empty[kBandSlots];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

// bf16 operand helpers (BF16 = true: the band and the weights are bf16, 64 channels per 128-byte row, tcgen05.mma kind::f16 with K = 16
// per instruction — the same 32 bytes of every row as a K = 8 tf32 instruction, so every descriptor offset below is shared)
__host__ __device__ constexpr uint32_t make_idesc_bf16_kk(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16_kk(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ uint32_t pack2_bf16(float a, float b) {  // round to nearest even
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ void st_global_v8_u32(void* p, const uint32_t* v) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]),
               "r"(v[7])
               : "memory");
}

// R = local taps per axis of one band (3 or 2); NI = MMA-issuing threads (local taps me, me + NI, ... of every band)
template <int BN, int NKB, int R, int NI, bool BF16 = false>
__global__ void __launch_bounds__((kEpiWarps + NI + 1) * 32, 1) conv_band_kernel(const __grid_constant__ CUtensorMap smap, BandParams p, int num_tiles) {
  constexpr int kWTileB = BN * kRowBytes, kTaps = R * R, kThreadsB = (kEpiWarps + NI + 1) * 32;
  static_assert(kTaps % NI == 0 && 2 * NI * BN <= 512, "issuer split / TMEM columns");
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* w_smem = smem;
  unsigned char* band_smem = smem + NKB * kWTileB;
  BandBars* bars = reinterpret_cast<BandBars*>(band_smem + kBandSlots * kBandB);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kBandSlots; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], NI);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bars->tmem_full[a], NI);
      mbar_init(&bars->tmem_empty[a], kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == kEpiWarps) tmem_alloc(&bars->tmem_base, 512);
  for (int i = threadIdx.x; i < kBandSlots * kBandB / 16; i += kThreadsB) reinterpret_cast<float4*>(band_smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int q = threadIdx.x; q < NKB * BN * 8; q += kThreadsB) {  // 16-byte chunks: row n of the prepared weights is NKB * 128 bytes in either type
    const int kb = q / (BN * 8), qq = q - kb * (BN * 8);
    const int n = qq >> 3, c = qq & 7;
    st_shared16(smem_u32(w_smem) + kb * kWTileB + swz(n, c),
                __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const unsigned char*>(p.wprep) + ((size_t)n * NKB + kb) * kRowBytes + c * 16)));
  }
  fence_proxy_async();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp < kEpiWarps) {
    // ================================ epilogue ================================
    const int r = warp * 32 + lane;
    const int yl = r / p.PW, xx = r - yl * p.PW;
    float bias[BN];
#pragma unroll
    for (int j = 0; j < BN; ++j) bias[j] = p.bias ? __ldg(p.bias + j) : 0.f;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int a = it & 1;
      const int n = tile / p.TPF, y0 = (tile - n * p.TPF) * p.RT;
      const bool valid0 = yl < p.RT && xx < p.OWv && y0 + yl < p.OH;
      const size_t off0 = (((size_t)n * p.out_H + (p.o_mul * (y0 + yl) + p.oy_add)) * p.out_W + (p.o_mul * xx + p.ox_add)) * BN;
      if (!BF16 && p.gate && !p.gate_bits && valid0) {
        // the ReLU masks this thread will need come from DRAM: pull their lines into L2 while the tile's MMAs are still running
        // (four epilogue warps cannot hide one DRAM latency per 32-column chunk)
#pragma unroll
        for (int c0 = 0; c0 < BN; c0 += 32) {
          size_t off = off0 + c0;
          if (p.scatter) {
            const int oy = min(2 * (y0 + yl) + (c0 >> 6), p.out_H - 1), ox = min(2 * xx + ((c0 >> 5) & 1), p.out_W - 1);
            off = (((size_t)n * p.out_H + oy) * p.out_W + ox) * 32;
          }
          asm volatile("prefetch.global.L2 [%0];" ::"l"(p.gate + off));
        }
      }
      // the sign-mask words of all the tile's chunks: loaded here, before the accumulator wait — inside the chunk loop each was a dependent
      // DRAM round trip (the masks were written a whole forward pass ago), four per tile for the stride-2 data gradient: ~3 us of latency
      // per tile, which was the kernel's time
      unsigned gmw[BN / 32];
#pragma unroll
      for (int c0 = 0; c0 < BN; c0 += 32) {
        gmw[c0 >> 5] = 0xFFFFFFFFu;
        if (p.gate_bits && valid0) {
          size_t off = off0 + c0;
          bool valid = true;
          if (p.scatter) {
            const int oy = 2 * (y0 + yl) + (c0 >> 6), ox = 2 * xx + ((c0 >> 5) & 1);
            valid = oy < p.out_H && ox < p.out_W;
            off = (((size_t)n * p.out_H + oy) * p.out_W + ox) * 32;
          }
          if (valid) gmw[c0 >> 5] = __ldg(p.gate_bits + (off >> 5));
        }
      }
      mbar_wait(&bars->tmem_full[a], (it >> 1) & 1);
      tc_fence_after_sync();
#pragma unroll
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32], u[NI > 1 ? NI - 1 : 1][32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(a * NI * BN + c0), v);
#pragma unroll
        for (int i = 1; i < NI; ++i) tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)((a * NI + i) * BN + c0), u[i - 1]);
        tmem_ld_wait();  // one wait for the partial sums of all issuers
#pragma unroll
        for (int i = 1; i < NI; ++i)
#pragma unroll
          for (int q = 0; q < 32; ++q) v[q] = __float_as_uint(__uint_as_float(v[q]) + __uint_as_float(u[i - 1][q]));
        if (c0 + 32 == BN) {
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->tmem_empty[a]);
        }
        bool valid = valid0;
        size_t off = off0 + c0;
        if (p.scatter) {  // chunk c0 / 32 = stride phase (py, px): 32 channels of pixel (2 y + py, 2 x + px)
          const int oy = 2 * (y0 + yl) + (c0 >> 6), ox = 2 * xx + ((c0 >> 5) & 1);
          valid = valid0 && oy < p.out_H && ox < p.out_W;
          off = (((size_t)n * p.out_H + oy) * p.out_W + ox) * 32;
        }
        if (valid) {
          // word of the sign masks for this (pixel, 32-channel chunk): `off` is the element offset of the chunk's first channel
          const size_t word = off >> 5;
          const unsigned gm = gmw[c0 >> 5];
          unsigned om = 0u;
          float o[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            o[j] = __uint_as_float(v[j]) + bias[c0 + j];
            if (p.relu) o[j] = fmaxf(o[j], 0.f);
          }
          if (p.gate_bits) {
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = (gm >> j) & 1u ? o[j] : 0.f;
          } else if (!BF16 && p.gate) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 g = __ldg(reinterpret_cast<const float4*>(p.gate + off + j));
              o[j] = g.x > 0.f ? o[j] : 0.f; o[j + 1] = g.y > 0.f ? o[j + 1] : 0.f; o[j + 2] = g.z > 0.f ? o[j + 2] : 0.f; o[j + 3] = g.w > 0.f ? o[j + 3] : 0.f;
            }
          }
          if (p.relu_bits) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (o[j] > 0.f) om |= 1u << j;
          }
          if (BF16) {  // 32 channels = 64 bytes of the channels-last bf16 tensor
            uint32_t w16[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) w16[j] = pack2_bf16(o[2 * j], o[2 * j + 1]);
            unsigned char* dst = reinterpret_cast<unsigned char*>(p.out) + off * 2;
            st_global_v8_u32(dst, w16);
            st_global_v8_u32(dst + 32, w16 + 8);
          } else {
            float* dst = reinterpret_cast<float*>(p.out) + off;
            if ((reinterpret_cast<size_t>(dst) & 31) == 0) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) st_global_v8(dst + j, o[j], o[j + 1], o[j + 2], o[j + 3], o[j + 4], o[j + 5], o[j + 6], o[j + 7]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
            }
          }
          if (p.relu_bits) p.relu_bits[word] = om;
        }
      }
    }
  } else if (warp < kEpiWarps + NI) {
    // ================================ MMA issuers ================================
    if (lane == 0) {
      const int me = warp - kEpiWarps;
      constexpr uint32_t idesc = BF16 ? make_idesc_bf16_kk(kBM, BN) : make_idesc_tf32(kBM, BN, false, false);
      const uint64_t a0 = make_desc<false, kBM, kBK>(smem_u32(band_smem), 0), b0 = make_desc<false, BN, kBK>(smem_u32(w_smem), 0);
      int it = 0, u = 0;  // u: band sequence number of this CTA
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int a = it & 1;
        mbar_wait(&bars->tmem_empty[a], ((it >> 1) & 1) ^ 1);
        tc_fence_after_sync();
        const uint32_t d = tmem_base + (uint32_t)((a * NI + me) * BN);
        for (int b = 0; b < p.nbands; ++b, ++u) {
          const int slot = u % kBandSlots;
          mbar_wait(&bars->full[slot], (u / kBandSlots) & 1);
          tc_fence_after_sync();
#pragma unroll
          for (int tap = me; tap < kTaps; tap += NI) {
            const int ty = tap / R, tx = tap - ty * R;
            const int shift = p.flip ? (R - 1 - ty) * p.PW + (R - 1 - tx) : ty * p.PW + tx;  // rows of 128 B
            const uint64_t da = a0 + (uint32_t)((slot * kBandB + shift * kRowBytes) >> 4), db = b0 + (uint32_t)((p.wkb[b * kMaxTaps + tap] * kWTileB) >> 4);
#pragma unroll
            for (int k = 0; k < kBK / 8; ++k) {  // four 32-byte k-steps of the 128-byte rows (K = 8 tf32 or 16 bf16 each)
              if (BF16) umma_bf16_kk(d, da + (uint32_t)(k * 2), db + (uint32_t)(k * 2), idesc, (uint32_t)(b != 0 || tap != me || k != 0));
              else umma_tf32(d, da + (uint32_t)(k * 2), db + (uint32_t)(k * 2), idesc, (uint32_t)(b != 0 || tap != me || k != 0));
            }
          }
          umma_commit(&bars->empty[slot]);
        }
        umma_commit(&bars->tmem_full[a]);
      }
    }
  } else {
    // ================================ TMA producer ================================
    if (lane == 0) {
      tma::prefetch_map(&smap);
      const uint32_t box_bytes = (uint32_t)(p.PW * p.PH * kRowBytes);
      int u = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n = tile / p.TPF, y0 = (tile - n * p.TPF) * p.RT;
        for (int b = 0; b < p.nbands; ++b, ++u) {
          const int slot = u % kBandSlots;
          mbar_wait(&bars->empty[slot], ((u / kBandSlots) & 1) ^ 1);
          tma::expect_tx(&bars->full[slot], box_bytes);
          tma::load_4d(smem_u32(band_smem) + slot * kBandB, &smap, &bars->full[slot], p.bc[b], p.bx[b], p.es * y0 + p.by[b], n);
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == kEpiWarps) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int BN, int NKB, int R, int NI>
int launch_band(const float* src, int N, int SH, int SW, int SC, BandParams& p, cudaStream_t st) {
  if (p.PW > kBM || p.PW * p.es > 256 || p.PH * p.es > 256) return (int)cudaErrorNotSupported;
  if (p.PW * p.PH > kBandRows || 127 + (R - 1) * p.PW + R - 1 >= kBandRows) return (int)cudaErrorNotSupported;
  CUtensorMap m;
  const uint64_t dims[4] = {(uint64_t)SC, (uint64_t)SW, (uint64_t)SH, (uint64_t)N};
  const uint64_t strides[3] = {(uint64_t)SC * 4, (uint64_t)SW * SC * 4, (uint64_t)SH * SW * SC * 4};
  const uint32_t box[4] = {32, (uint32_t)(p.PW * p.es), (uint32_t)(p.PH * p.es), 1};
  const uint32_t es[4] = {1, (uint32_t)p.es, (uint32_t)p.es, 1};
  if (tma::make_map(&m, src, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, es) != 0) return (int)cudaErrorNotSupported;
  const long long tiles = (long long)N * p.TPF;
  if (tiles >= (1ll << 31)) return (int)cudaErrorNotSupported;
  constexpr int smem = NKB * BN * kRowBytes + kBandSlots * kBandB + 256 + 1024;
  static_assert(smem <= 227 * 1024, "shared memory budget");
  auto kfn = conv_band_kernel<BN, NKB, R, NI>;
  HULC_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  HULC_LAUNCH(kfn, dim3((unsigned)min((long long)kNumSMs, tiles)), dim3((kEpiWarps + NI + 1) * 32), smem, st, m, p, (int)tiles);
  HULC_RETURN_LAST();
}

// forward of the two channels-last layers on the band scheme (cudaErrorNotSupported: geometry does not fit)
int band_fwd(const float* x, const float* wprep, const float* b, float* y, unsigned* relu_bits, int N, int CIN, int H, int W, int KS, int S, int OH, int OW, int relu,
             cudaStream_t st) {
  BandParams p{};
  p.N = N; p.OH = OH; p.OWv = OW; p.flip = 0; p.es = S;
  p.wprep = wprep; p.out = y; p.out_H = OH; p.out_W = OW; p.o_mul = 1; p.oy_add = 0; p.ox_add = 0; p.bias = b; p.gate = nullptr; p.gate_bits = nullptr; p.relu_bits = relu_bits; p.relu = relu;
  if (KS == 3 && S == 1 && CIN == 64) {  // bands = channel blocks
    p.PW = W; p.RT = min(OH, kBM / p.PW); p.PH = p.RT + 2; p.TPF = hulc_cdiv(OH, p.RT); p.nbands = 2;
    for (int cb = 0; cb < 2; ++cb) {
      p.bc[cb] = cb * 32; p.bx[cb] = 0; p.by[cb] = 0;
      for (int tap = 0; tap < 9; ++tap) p.wkb[cb * kMaxTaps + tap] = tap * 2 + cb;
    }
    return launch_band<64, 18, 3, 3>(x, N, H, W, 64, p, st);
  }
  if (KS == 4 && S == 2 && CIN == 32) {  // bands = input phases (py, px); local tap (ty, tx) is kernel tap (2 ty + py, 2 tx + px)
    p.PW = (W + 1) / 2; p.RT = min(OH, kBM / p.PW); p.PH = p.RT + 1; p.TPF = hulc_cdiv(OH, p.RT); p.nbands = 4;
    for (int ph = 0; ph < 4; ++ph) {
      const int py = ph >> 1, px = ph & 1;
      p.bc[ph] = 0; p.bx[ph] = px; p.by[ph] = py;
      for (int tap = 0; tap < 4; ++tap) p.wkb[ph * kMaxTaps + tap] = (2 * (tap >> 1) + py) * 4 + 2 * (tap & 1) + px;
    }
    return launch_band<64, 16, 2, 2>(x, N, H, W, 32, p, st);
  }
  return (int)cudaErrorNotSupported;
}

// ALL four stride phases of the 4x4 stride-2 data gradient in one launch: the phases read the same shifted views of dY and differ
// only in their weights, so they are stacked along N (4 phases x 32 input channels = 128 columns: one tcgen05.mma does the work of
// four N = 32 ones) and the epilogue scatters column chunk ph to pixel (2 y + ph / 2, 2 x + ph % 2).  wall = [4][CIN][(jy, jx, co)].
int band_dgrad_s2_all(const float* dy, const float* wall, const float* gate, const unsigned* gate_bits, float* dx, int N, int H, int W, int HO, int WO,
                      cudaStream_t st) {
  constexpr int R = 2;
  BandParams p{};
  p.N = N; p.OH = (H + 1) / 2; p.OWv = (W + 1) / 2; p.flip = 1; p.es = 1; p.scatter = 1;
  p.wprep = wall; p.out = dx; p.out_H = H; p.out_W = W; p.o_mul = 1; p.oy_add = 0; p.ox_add = 0; p.bias = nullptr; p.gate = gate; p.gate_bits = gate_bits; p.relu_bits = nullptr; p.relu = 0;
  p.PW = p.OWv + R - 1; p.RT = min(p.OH, kBM / p.PW); p.PH = p.RT + R - 1; p.TPF = hulc_cdiv(p.OH, p.RT); p.nbands = 2;
  for (int cb = 0; cb < 2; ++cb) {
    p.bc[cb] = cb * 32; p.bx[cb] = -(R - 1); p.by[cb] = -(R - 1);
    for (int tap = 0; tap < R * R; ++tap) p.wkb[cb * kMaxTaps + tap] = tap * 2 + cb;
  }
  return launch_band<128, 8, 2, 2>(dy, N, HO, WO, 64, p, st);
}

// one stride phase of a data gradient on the band scheme: output grid (OH, OW) = the phase's input pixels
int band_dgrad(const float* dy, const float* wphase, const float* gate, const unsigned* gate_bits, float* dx, int N, int CIN, int H, int W, int HO, int WO, int R,
               int S, int py, int px, int OH, int OW, cudaStream_t st) {
  BandParams p{};
  p.N = N; p.OH = OH; p.OWv = OW; p.flip = 1; p.es = 1;
  p.wprep = wphase; p.out = dx; p.out_H = H; p.out_W = W; p.o_mul = S; p.oy_add = py; p.ox_add = px; p.bias = nullptr; p.gate = gate; p.gate_bits = gate_bits; p.relu_bits = nullptr; p.relu = 0;
  p.PW = OW + R - 1; p.RT = min(OH, kBM / p.PW); p.PH = p.RT + R - 1; p.TPF = hulc_cdiv(OH, p.RT); p.nbands = 2;
  for (int cb = 0; cb < 2; ++cb) {
    p.bc[cb] = cb * 32; p.bx[cb] = -(R - 1); p.by[cb] = -(R - 1);
    for (int tap = 0; tap < R * R; ++tap) p.wkb[cb * kMaxTaps + tap] = tap * 2 + cb;
  }
  if (CIN == 64 && R == 3) return launch_band<64, 18, 3, 3>(dy, N, HO, WO, 64, p, st);
  if (CIN == 32 && R == 2) return launch_band<32, 8, 2, 2>(dy, N, HO, WO, 64, p, st);
  return (int)cudaErrorNotSupported;
}

// ---- bf16 bands: 64 channels (or, for the 32-channel stride-2 layer, a PAIR of horizontally adjacent pixels) per 128-byte row -------
// pos_w positions of 64 bf16 per source row; row / frame strides in bytes; es_y = element stride of the walk in y (the stride-2 forward
// takes every second row: one band per row phase).
template <int BN, int NKB, int R, int NI>
int launch_band_bf16(const void* src, int N, int SH, int pos_w, size_t row_bytes, size_t frame_bytes, int es_y, BandParams& p, cudaStream_t st) {
  if (p.PW > kBM || p.PW > 256 || p.PH * es_y > 256) return (int)cudaErrorNotSupported;
  if (p.PW * p.PH > kBandRows || 127 + (R - 1) * p.PW + R - 1 >= kBandRows) return (int)cudaErrorNotSupported;
  if ((row_bytes & 15) || (frame_bytes & 15) || (reinterpret_cast<size_t>(src) & 15)) return (int)cudaErrorNotSupported;
  CUtensorMap m;
  const uint64_t dims[4] = {64, (uint64_t)pos_w, (uint64_t)SH, (uint64_t)N};
  const uint64_t strides[3] = {128, (uint64_t)row_bytes, (uint64_t)frame_bytes};
  const uint32_t box[4] = {64, (uint32_t)p.PW, (uint32_t)(p.PH * es_y), 1};
  const uint32_t es[4] = {1, 1, (uint32_t)es_y, 1};
  if (tma::make_map(&m, src, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, es, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16) != 0) return (int)cudaErrorNotSupported;
  const long long tiles = (long long)N * p.TPF;
  if (tiles >= (1ll << 31)) return (int)cudaErrorNotSupported;
  p.es = es_y;
  constexpr int smem = NKB * BN * kRowBytes + kBandSlots * kBandB + 256 + 1024;
  static_assert(smem <= 227 * 1024, "shared memory budget");
  auto kfn = conv_band_kernel<BN, NKB, R, NI, true>;
  HULC_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  HULC_LAUNCH(kfn, dim3((unsigned)min((long long)kNumSMs, tiles)), dim3((kEpiWarps + NI + 1) * 32), smem, st, m, p, (int)tiles);
  HULC_RETURN_LAST();
}

template <int BN, int NKB>
int launch(const CUtensorMap& smap, const TcParams& p, int num_tiles, cudaStream_t st) {
  constexpr int kW = NKB * BN * kRowBytes;
  constexpr int kStagesRaw = (kSmemBudget - kW - 1280) / kATileB;
  constexpr int STAGES = kStagesRaw >= 8 ? 8 : (kStagesRaw / kNI) * kNI;  // a multiple of the issuer count (see the MMA issuers)
  static_assert(STAGES >= kNI && NKB >= kNI, "weights too large to stay resident");
  constexpr int smem = kW + STAGES * kATileB + 256 + 1024;
  auto kfn = conv_tma_kernel<BN, NKB, STAGES>;
  HULC_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  HULC_LAUNCH(kfn, dim3(min(kNumSMs, num_tiles)), dim3(kThr), smem, st, smap, p, num_tiles);
  HULC_RETURN_LAST();
}

// tiling of an (OH, OW) output grid: whole rows; several frames per tile when a frame is small
bool tiling(int N, int OH, int OW, TcParams& p, int& num_tiles) {
  if (OW > kBM || OW <= 0 || OH <= 0) return false;
  p.RT = min(OH, kBM / OW);
  if (p.RT == OH) { p.NB = max(1, min(kBM / (OH * OW), 256)); p.TPF = 1; num_tiles = hulc_cdiv(N, p.NB); }
  else { p.NB = 1; p.TPF = hulc_cdiv(OH, p.RT); num_tiles = N * p.TPF; }
  return true;
}

// source tensor map: NHWC [N][H][W][C], box = 32 channels x OW x RT x NB grid points walked with element stride S
int source_map(CUtensorMap* m, const float* src, int N, int H, int W, int C, const TcParams& p) {
  const uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)N};
  const uint64_t strides[3] = {(uint64_t)C * 4, (uint64_t)W * C * 4, (uint64_t)H * W * C * 4};
  const uint32_t box[4] = {32, (uint32_t)(p.OW * p.S), (uint32_t)(p.RT * p.S), (uint32_t)p.NB};
  const uint32_t es[4] = {1, (uint32_t)p.S, (uint32_t)p.S, 1};
  if (box[1] > 256 || box[2] > 256) return (int)cudaErrorNotSupported;
  return tma::make_map(m, src, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, es);
}

}  // namespace

// y (NHWC) = relu?(conv(x NHWC, w) + b); wprep = w as [COUT][(ky, kx, ci)] (prep_fwd_weights_kernel).  cudaErrorNotSupported
// when the geometry does not fit: the caller falls back to the gather kernel.
int hulc_conv_tma_fwd(const float* x, const float* wprep, const float* b, float* y, unsigned* relu_bits, int N, int CIN, int H, int W, int COUT, int KS, int S,
                      int relu, cudaStream_t st) {
  if ((reinterpret_cast<size_t>(x) | reinterpret_cast<size_t>(wprep) | reinterpret_cast<size_t>(y)) & 15) return (int)cudaErrorNotSupported;
  TcParams p{};
  p.N = N; p.OH = (H - KS) / S + 1; p.OW = (W - KS) / S + 1; p.S = S;
  int num_tiles;
  if (!tiling(N, p.OH, p.OW, p, num_tiles)) return (int)cudaErrorNotSupported;
  p.taps_x = KS; p.cblocks = CIN / 32; p.sign = 1; p.wprep = wprep; p.out = y; p.out_H = p.OH; p.out_W = p.OW; p.o_mul = 1; p.oy_add = 0; p.ox_add = 0;
  p.bias = b; p.gate = nullptr; p.relu = relu;
  CUtensorMap m;
  const int rc = source_map(&m, x, N, H, W, CIN, p);
  if (rc != 0) return (int)cudaErrorNotSupported;
  if (g_use_band && COUT == 64) {
    const int rb = band_fwd(x, wprep, b, y, relu_bits, N, CIN, H, W, KS, S, p.OH, p.OW, relu, st);
    if (rb != (int)cudaErrorNotSupported) return rb;
  }
  if (relu_bits) return (int)cudaErrorNotSupported;  // only the band kernels (and the gather fallback) write the sign mask
  if (CIN == 32 && COUT == 64 && KS == 4) return launch<64, 16>(m, p, num_tiles, st);
  if (CIN == 64 && COUT == 64 && KS == 3) return launch<64, 18>(m, p, num_tiles, st);
  return (int)cudaErrorNotSupported;
}

// dx (NHWC [N][H][W][32]) = data gradient of the 32 -> 64, 4x4, stride-2 layer, all stride phases in one launch; wall = the four
// prepared phase weight matrices, contiguous.  cudaErrorNotSupported -> per-phase launches.
int hulc_conv_tma_dgrad_s2_all(const float* dy, const float* wall, const float* gate, const unsigned* gate_bits, float* dx, int N, int CIN, int H, int W, int COUT,
                               int HO, int WO, cudaStream_t st) {
  if (!g_use_band || CIN != 32 || COUT != 64) return (int)cudaErrorNotSupported;
  if ((reinterpret_cast<size_t>(dy) | reinterpret_cast<size_t>(wall) | reinterpret_cast<size_t>(dx) | reinterpret_cast<size_t>(gate)) & 15)
    return (int)cudaErrorNotSupported;
  return band_dgrad_s2_all(dy, wall, gate, gate_bits, dx, N, H, W, HO, WO, st);
}

// One stride phase (py, px) of dx (NHWC [N][H][W][CIN]) = conv_transpose(dy NHWC [N][HO][WO][COUT], w), masked by gate > 0.
// wphase = w as [CIN][(jy, jx, co)] for this phase (prep_dgrad_weights_kernel), R = KS / S taps per axis.
int hulc_conv_tma_dgrad_phase(const float* dy, const float* wphase, const float* gate, const unsigned* gate_bits, float* dx, int N, int CIN, int H, int W, int COUT,
                              int HO, int WO, int R, int S, int py, int px, cudaStream_t st) {
  if ((reinterpret_cast<size_t>(dy) | reinterpret_cast<size_t>(wphase) | reinterpret_cast<size_t>(dx) | reinterpret_cast<size_t>(gate)) & 15)
    return (int)cudaErrorNotSupported;
  TcParams p{};
  p.N = N; p.OH = (H - py + S - 1) / S; p.OW = (W - px + S - 1) / S; p.S = 1;
  int num_tiles;
  if (p.OH <= 0 || p.OW <= 0) return 0;
  if (!tiling(N, p.OH, p.OW, p, num_tiles)) return (int)cudaErrorNotSupported;
  p.taps_x = R; p.cblocks = COUT / 32; p.sign = -1; p.wprep = wphase; p.out = dx; p.out_H = H; p.out_W = W; p.o_mul = S; p.oy_add = py; p.ox_add = px;
  p.bias = nullptr; p.gate = gate; p.relu = 0;
  CUtensorMap m;
  const int rc = source_map(&m, dy, N, HO, WO, COUT, p);
  if (rc != 0) return (int)cudaErrorNotSupported;
  if (g_use_band && COUT == 64) {
    const int rb = band_dgrad(dy, wphase, gate, gate_bits, dx, N, CIN, H, W, HO, WO, R, S, py, px, p.OH, p.OW, st);
    if (rb != (int)cudaErrorNotSupported) return rb;
  }
  if (CIN == 32 && COUT == 64 && R == 2) return launch<32, 8>(m, p, num_tiles, st);
  if (CIN == 64 && COUT == 64 && R == 3) return launch<64, 18>(m, p, num_tiles, st);
  return (int)cudaErrorNotSupported;
}

// ---- bf16 activations (BASELINE config 3): x / y / dy / dx are channels-last bf16, prepared weights are bf16 in the same K order as the
// fp32 kernels use (forward [COUT][(ky, kx, ci)], data gradient [phase][CIN][(jy, jx, co)]), bias fp32, ReLU sign masks as above.
int hulc_conv_band_bf16_fwd(const void* x, const void* wprep, const float* b, void* y, unsigned* relu_bits, int N, int CIN, int H, int W, int COUT, int KS, int S,
                            int relu, cudaStream_t st) {
  const int OH = (H - KS) / S + 1, OW = (W - KS) / S + 1;
  if (COUT != 64 || OH <= 0 || OW <= 0) return (int)cudaErrorNotSupported;
  BandParams p{};
  p.N = N; p.OH = OH; p.OWv = OW; p.flip = 0;
  p.wprep = wprep; p.out = y; p.out_H = OH; p.out_W = OW; p.o_mul = 1; p.oy_add = 0; p.ox_add = 0; p.bias = b; p.gate = nullptr; p.gate_bits = nullptr; p.relu_bits = relu_bits; p.relu = relu;
  if (KS == 3 && S == 1 && CIN == 64) {  // one band of 64 channels per tile; k-block = tap
    p.PW = W; p.RT = min(OH, kBM / p.PW); p.PH = p.RT + 2; p.TPF = hulc_cdiv(OH, p.RT); p.nbands = 1;
    p.bc[0] = 0; p.bx[0] = 0; p.by[0] = 0;
    for (int tap = 0; tap < 9; ++tap) p.wkb[tap] = tap;
    return launch_band_bf16<64, 9, 3, 3>(x, N, H, W, (size_t)W * 128, (size_t)H * W * 128, 1, p, st);
  }
  if (KS == 4 && S == 2 && CIN == 32) {
    // a row of the band = the pixel PAIR (2 q, 2 q + 1) x 32 channels = kernel columns kx = 2 tx, 2 tx + 1 of local tap tx, so a k-block is
    // (ky, tx) x (px, ci) — 64 contiguous elements of the (ky, kx, ci)-ordered weights; one band per row phase py, local tap (ty, tx) is
    // kernel row ky = 2 ty + py.  The pair view needs no de-interleave in x (a row of W pixels is W / 2 whole pairs; an odd last pixel is
    // never read by a stride-2 window).
    p.PW = W / 2; p.RT = min(OH, kBM / p.PW); p.PH = p.RT + 1; p.TPF = hulc_cdiv(OH, p.RT); p.nbands = 2;
    if (p.PW < OW + 1) return (int)cudaErrorNotSupported;
    for (int py = 0; py < 2; ++py) {
      p.bc[py] = 0; p.bx[py] = 0; p.by[py] = py;
      for (int tap = 0; tap < 4; ++tap) p.wkb[py * kMaxTaps + tap] = (2 * (tap >> 1) + py) * 2 + (tap & 1);
    }
    return launch_band_bf16<64, 8, 2, 2>(x, N, H, W / 2, (size_t)W * 64, (size_t)H * W * 64, 2, p, st);
  }
  return (int)cudaErrorNotSupported;
}

// dx (bf16 NHWC [N][H][W][CIN]) = conv_transpose(dy bf16 [N][HO][WO][64], w), masked by the ReLU sign bits of the activation that fed the layer.
// KS = 3, S = 1: wprep = [CIN = 64][(jy, jx, co)]; KS = 4, S = 2: wprep = the four phase matrices [4][CIN = 32][(jy, jx, co)], all phases in one launch.
int hulc_conv_band_bf16_dgrad(const void* dy, const void* wprep, const unsigned* gate_bits, void* dx, int N, int CIN, int H, int W, int COUT, int HO, int WO, int KS,
                              int S, cudaStream_t st) {
  if (COUT != 64) return (int)cudaErrorNotSupported;
  BandParams p{};
  p.N = N; p.flip = 1;
  p.wprep = wprep; p.out = dx; p.out_H = H; p.out_W = W; p.o_mul = 1; p.oy_add = 0; p.ox_add = 0; p.bias = nullptr; p.gate = nullptr; p.gate_bits = gate_bits; p.relu_bits = nullptr; p.relu = 0;
  p.nbands = 1; p.bc[0] = 0;
  if (KS == 3 && S == 1 && CIN == 64) {
    constexpr int R = 3;
    p.OH = H; p.OWv = W; p.PW = W + R - 1; p.RT = min(p.OH, kBM / p.PW); p.PH = p.RT + R - 1; p.TPF = hulc_cdiv(p.OH, p.RT);
    p.bx[0] = -(R - 1); p.by[0] = -(R - 1);
    for (int tap = 0; tap < 9; ++tap) p.wkb[tap] = tap;
    return launch_band_bf16<64, 9, 3, 3>(dy, N, HO, WO, (size_t)WO * 128, (size_t)HO * WO * 128, 1, p, st);
  }
  if (KS == 4 && S == 2 && CIN == 32) {  // the four stride phases stacked along N = 128, scattered by the epilogue
    constexpr int R = 2;
    p.OH = (H + 1) / 2; p.OWv = (W + 1) / 2; p.scatter = 1;
    p.PW = p.OWv + R - 1; p.RT = min(p.OH, kBM / p.PW); p.PH = p.RT + R - 1; p.TPF = hulc_cdiv(p.OH, p.RT);
    p.bx[0] = -(R - 1); p.by[0] = -(R - 1);
    for (int tap = 0; tap < 4; ++tap) p.wkb[tap] = tap;
    return launch_band_bf16<128, 4, 2, 2>(dy, N, HO, WO, (size_t)WO * 128, (size_t)HO * WO * 128, 1, p, st);
  }
  return (int)cudaErrorNotSupported;
}
