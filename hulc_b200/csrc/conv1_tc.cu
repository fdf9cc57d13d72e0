// conv1_tc.cu — the first convolution of the perceptual encoders (3 -> 32 channels, 8x8, stride 4, NCHW fp32 frames;
// vision_network.py:37, vision_network_gripper.py:12), forward, with the input staged ONCE per tile in shared memory.
//
// The layer is bound by data movement (37 FLOP per algorithmic byte).  Every input pixel takes part in 2 x 2 output windows,
// so an im2col gather from global memory (conv_tc.cu, FwdNchw3Loader) pulls each frame through L2 four times at ~26 B/clk/SM —
// the measured L2 -> SM delivery rate of cp.async and TMA gathers alike — and spends ~150 instructions per producer warp per
// k-block on addresses.  Here a tile is RT whole output rows of one frame (RT * WO <= 128 pixels: 2 rows of 49 for the
// 200x200 camera, 6 rows of 20 for the 84x84 one) and
//   * one thread fetches the band of 4 RT + 4 input rows x 3 channels the tile needs with three cp.async.bulk copies
//     (each channel's rows are one contiguous run) into a 3-deep ring: every input byte crosses L2 -> SM about once;
//   * eight builder warps, each owning one stage of the A ring, copy 16-byte pieces of the band into the swizzled K-major UMMA
//     tile of one k-block (4 (ci, ky) pairs x 8 kx; 32 ld.shared / st.shared per lane, offsets precomputed once per kernel);
//     k-block j of the CTA's sequence goes to warp / stage j % 8, so the proxy fences and barrier round trips of eight
//     k-blocks overlap and a stage always has one producer and one consumer;
//   * the weights [32 x 192] stay resident in shared memory; two threads issue tcgen05.mma (kind::tf32, M = 128, N = 32),
//     even and odd k-blocks into separate TMEM accumulators (a single issuing thread is bound by its own barrier-wait /
//     commit overhead, scripts/micro/mma_rate.cu);
//   * four warps run the epilogue: sum the two accumulators, bias + ReLU, channels-last stores (128 B per pixel).
#include "common.cuh"
#include "tc_pipeline.cuh"
#include "tma.cuh"

namespace {

using namespace tc;

constexpr int kCout = 32, kKtot = 192, kNumKb = kKtot / kBK;  // 6 k-blocks = 6 builder warps = 6 stages
constexpr int kAStage = kBM * kRowBytes;                       // 16 KB
constexpr int kWTileB = kCout * kRowBytes;                     // one k-block of the weights: 4 KB
constexpr int kBandBytes = 30 * 1024;
constexpr int kBands = 2;
constexpr int kIssuers = 2;
constexpr int kStages = 8;    // A-tile ring; k-block j of the CTA's sequence lives in stage j % 8 and is built by builder warp j % 8
constexpr int kBuilders = kStages;
constexpr int kSmem = kNumKb * kWTileB + kStages * kAStage + kBands * kBandBytes + 256 + 1024;
// warps: 0-3 epilogue | 4-5 MMA issuers | 6 band loader | 7-14 builders
constexpr int kLoaderWarp = kEpiWarps + kIssuers, kBuilder0 = kLoaderWarp + 1;
constexpr int kThr = (kBuilder0 + kBuilders) * 32;
static_assert(kStages % kIssuers == 0 && kNumKb % kIssuers == 0, "every stage has one issuer; an issuer's k-blocks of a tile are kb = me, me + 2, ...");

struct C1Params {
  const float* x;   // [N, 3, H, W]
  const float* w;   // [32, 3, 8, 8] = [32][192], K-major in (ci, ky, kx) order
  const float* b;
  float* y;         // [N, HO, WO, 32]
  unsigned* bits;   // optional: sign mask of y, one word per pixel
  int N, H, W, HO, WO;
  int RT, TPF, BR;  // output rows per tile, tiles per frame, band rows = 4 RT + 4
  int relu;
};

struct C1Bars {
  uint64_t full[kStages];
  uint64_t empty[kStages];
  uint64_t band_full[kBands];
  uint64_t band_empty[kBands];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ float4 ld_shared16(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void st_shared16(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__global__ void __launch_bounds__(kThr, 1) conv1_band_fwd_kernel(C1Params p, int num_tiles) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* w_smem = smem;
  unsigned char* a_smem = smem + kNumKb * kWTileB;
  unsigned char* band_smem = a_smem + kStages * kAStage;
  C1Bars* bars = reinterpret_cast<C1Bars*>(band_smem + kBands * kBandBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    for (int s = 0; s < kBands; ++s) {
      mbar_init(&bars->band_full[s], 1);
      mbar_init(&bars->band_empty[s], kNumKb);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bars->tmem_full[a], kIssuers);
      mbar_init(&bars->tmem_empty[a], kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == kEpiWarps) tmem_alloc(&bars->tmem_base, 2 * kIssuers * kCout);
  // rows of the A stages that no pixel maps to are never written: keep them finite
  for (int i = threadIdx.x; i < kStages * kAStage / 16; i += kThr) reinterpret_cast<float4*>(a_smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  // resident weights: 6 K-major swizzled tiles [32 rows x 128 B]
  for (int q = threadIdx.x; q < kNumKb * kCout * 8; q += kThr) {
    const int kb = q / (kCout * 8), qq = q - kb * (kCout * 8);
    const int n = qq >> 3, c = qq & 7;
    st_shared16(smem_u32(w_smem) + kb * kWTileB + swz(n, c), __ldg(reinterpret_cast<const float4*>(p.w + (size_t)n * kKtot + kb * kBK + c * 4)));
  }
  fence_proxy_async();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;
  const int tile_px = p.RT * p.WO;

  if (warp < kEpiWarps) {
    // ================================ epilogue ================================
    const int r = warp * 32 + lane;
    const int yl = r / p.WO, xx = r - yl * p.WO;
    float bias[kCout];
#pragma unroll
    for (int j = 0; j < kCout; ++j) bias[j] = p.b ? __ldg(p.b + j) : 0.f;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int a = it & 1;
      mbar_wait(&bars->tmem_full[a], (it >> 1) & 1);
      tc_fence_after_sync();
      uint32_t v[32], u[32];
      tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(a * kIssuers * kCout), v);
      tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(a * kIssuers * kCout + kCout), u);
      tmem_ld_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->tmem_empty[a]);  // the values are in registers: the accumulators may be reused
      const int n = tile / p.TPF, y0 = (tile - n * p.TPF) * p.RT;
      if (r < tile_px && y0 + yl < p.HO) {
        const size_t pix = (size_t)(n * p.HO + y0 + yl) * p.WO + xx;
        float* dst = p.y + pix * kCout;
        unsigned om = 0u;
#pragma unroll
        for (int j = 0; j < kCout; j += 4) {
          float4 o = make_float4(__uint_as_float(v[j]) + __uint_as_float(u[j]) + bias[j], __uint_as_float(v[j + 1]) + __uint_as_float(u[j + 1]) + bias[j + 1],
                                 __uint_as_float(v[j + 2]) + __uint_as_float(u[j + 2]) + bias[j + 2], __uint_as_float(v[j + 3]) + __uint_as_float(u[j + 3]) + bias[j + 3]);
          if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
          if (o.x > 0.f) om |= 1u << j;
            if (o.y > 0.f) om |= 1u << (j + 1);
            if (o.z > 0.f) om |= 1u << (j + 2);
            if (o.w > 0.f) om |= 1u << (j + 3);
          *reinterpret_cast<float4*>(dst + j) = o;
        }
        if (p.bits) p.bits[pix] = om;
      }
    }
  } else if (warp < kLoaderWarp) {
    // ================================ MMA issuers: k-blocks me, me + 2, me + 4 ================================
    if (lane == 0) {
      const int me = warp - kEpiWarps;
      constexpr uint32_t idesc = make_idesc_tf32(kBM, kCout, false, false);
      const uint64_t a0 = make_desc<false, kBM, kBK>(smem_u32(a_smem), 0), b0 = make_desc<false, kCout, kBK>(smem_u32(w_smem), 0);
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int a = it & 1;
        mbar_wait(&bars->tmem_empty[a], ((it >> 1) & 1) ^ 1);
        tc_fence_after_sync();
        const uint32_t d = tmem_base + (uint32_t)((a * kIssuers + me) * kCout);
#pragma unroll
        for (int kb = me; kb < kNumKb; kb += kIssuers) {
          const int j = it * kNumKb + kb, stage = j % kStages;
          mbar_wait(&bars->full[stage], (j / kStages) & 1);
          tc_fence_after_sync();
#pragma unroll
          for (int k = 0; k < kBK / 8; ++k)
            umma_tf32(d, a0 + (uint32_t)(stage * (kAStage >> 4) + k * 2), b0 + (uint32_t)(kb * (kWTileB >> 4) + k * 2), idesc, (uint32_t)(kb != me || k != 0));
          umma_commit(&bars->empty[stage]);
        }
        umma_commit(&bars->tmem_full[a]);
      }
    }
  } else if (warp == kLoaderWarp) {
    // ================================ band loader ================================
    if (lane == 0) {
      const uint32_t rowbytes = (uint32_t)p.W * 4;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int slot = it % kBands;
        mbar_wait(&bars->band_empty[slot], ((it / kBands) & 1) ^ 1);
        const int n = tile / p.TPF, y0 = (tile - n * p.TPF) * p.RT;
        const int rows = min(p.BR, p.H - 4 * y0);  // rows past the frame are not needed by any valid pixel
        const uint32_t bytes = (uint32_t)rows * rowbytes;
        expect_tx(&bars->band_full[slot], 3 * bytes);
        const uint32_t dst = smem_u32(band_smem) + slot * kBandBytes;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci)
          bulk_load(dst + (uint32_t)(ci * p.BR) * rowbytes, p.x + ((size_t)(n * 3 + ci) * p.H + 4 * y0) * p.W, bytes, &bars->band_full[slot]);
      }
    }
  } else {
    // ================================ builders: warp w builds k-blocks w, w + 8, ... of the CTA's sequence into stage w ====
    const int bw = warp - kBuilder0;
    const int c = lane & 7, pair = c >> 1, half = c & 1;  // 16-byte chunk c of a tile row = (ci, ky) pair `pair` of the k-block, kx half
    const uint32_t rowbytes = (uint32_t)p.W * 4;
    // rows (pixels) (lane >> 3) + 4 i, i = 0..31: source offset inside one channel plane of the band, relative to kernel row ky0
    uint32_t soff[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int r = (lane >> 3) + 4 * i;
      const int yl = r / p.WO, xx = r - yl * p.WO;
      soff[i] = (uint32_t)(yl * 4 + pair) * rowbytes + (uint32_t)((xx * 4 + half * 4) * 4);
    }
    const uint32_t dst = smem_u32(a_smem) + bw * kAStage;
    const int my_tiles = ((int)blockIdx.x < num_tiles) ? (num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    for (int j = bw; j < my_tiles * kNumKb; j += kBuilders) {
      const int it = j / kNumKb, kb = j - it * kNumKb;
      const int tile = blockIdx.x + it * gridDim.x;
      const int slot = it % kBands;
      const int n = tile / p.TPF, y0 = (tile - n * p.TPF) * p.RT;
      const int px_valid = min(p.RT, p.HO - y0) * p.WO;  // pixels of this tile that exist (ragged last tile of a frame)
      const uint32_t band = smem_u32(band_smem) + slot * kBandBytes + (uint32_t)((kb >> 1) * p.BR + (kb & 1) * 4) * rowbytes;  // channel kb/2, ky0 = 4 (kb%2)
      mbar_wait(&bars->band_full[slot], (it / kBands) & 1);
      mbar_wait(&bars->empty[bw], ((j / kBuilders) & 1) ^ 1);  // the MMAs of this stage's previous k-block have read it
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int r = (lane >> 3) + 4 * i;
        if (r < px_valid) st_shared16(dst + swz(r, c), ld_shared16(band + soff[i]));
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&bars->full[bw]);
        mbar_arrive(&bars->band_empty[slot]);
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == kEpiWarps) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 2 * kIssuers * kCout);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Weight gradient of the same layer: dW[k = (ci, ky, kx)][co] = sum over pixels of Xcol[pixel][k] * dY[pixel][co].
// GEMM-K = pixels, so both operands are MN-major tiles ([32 elements of M or N][pixel k-rows][128 B], 32-byte-atom swizzle):
//   A = Xcol^T: six groups of 32 k's (group g = (ci, ky) pairs 4g .. 4g+3 x 8 kx), built from the SAME staged input band as the
//       forward kernel — a pixel's 32 k's of a group are four 32-byte pieces of four band rows;
//   B = dY^T  : one group of 32 channels, copied from the contiguous channels-last rows of dY.
// A tile is ONE output row of one frame (WO <= 56 pixels, padded to a multiple of 8 k-rows with zero rows); seven builder warps
// (six A groups + dY) fill a 3-deep ring of tile slots; two threads issue tcgen05.mma — k 0..127 into accumulator 0 and
// k 128..191 into accumulator 1 — and the accumulators stay in TMEM across ALL tiles of the CTA: each CTA writes one partial
// [192 x 32] at the end, which the existing reduction sums over CTAs in a fixed order.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kWgKR = 56;                       // k-rows per tile, multiple of 8
constexpr int kWgGroupB = kWgKR * kRowBytes;    // one 32-element group of a tile: 7 KB
// The contraction runs over PIXEL GROUPS, not pixels: with kx = 4 a + b the window element x[ci][4 oy + ky][4 ox + kx] is
// Xp[ci][4 oy + ky][ox + a][b], Xp[..][j][b] = x[..][4 j + b] — the input row cut into W/4 = WO + 1 groups of 4 floats.  So
//   dW[co][ci][ky][4 a + b] = sum_j Xp[ci][4 oy + ky][j][b] * dY_a[j][co],   dY_0[j] = dY[oy][j],  dY_1[j] = dY[oy][j - 1]
// and the two halves of the kernel window (a = 0, 1) share ONE staged operand of 96 columns (ci, ky, b) instead of an im2col matrix of
// 192 in which every input value appears twice; the shift moves into the small operand (dY staged twice, one k-row apart) and both
// halves ride in one instruction of N = 64.  Per tile the builders write 35 KB instead of 49 KB and the tensor core reads 42 KB of
// operands instead of 70 KB — the first kernel spent its time on exactly that shared-memory traffic (≈190 KB per tile at 128 B/clk).
// slot = 6 groups: 0-2 Xp^T of channel ci (column = 4 ky + b), 3 = a constant group whose element 0 is 1 (accumulator row 96 = sum over
// the pixels of dY = the BIAS gradient, for free), 4 = dY_0^T, 5 = dY_1^T
constexpr int kWgOnes = 3, kWgDy = 4;
constexpr int kWgSlotB = 6 * kWgGroupB;         // 42 KB
constexpr int kWgSlots = 2;                     // tiles being built / multiplied; slot s belongs to issuer s
constexpr int kWgXB = 19 * 1024;                // 8 input rows x 3 channels x W floats (W <= 202)
constexpr int kWgDyStageB = kWgKR * kRowBytes;  // the tile's dY row, fetched by the same bulk-copy thread (fp32: 128 B per pixel, bf16: 64)
constexpr int kWgBandB = kWgXB + kWgDyStageB;   // 26 KB
constexpr int kWgBands = 4;                     // bands in flight: the loader runs up to four tiles ahead of the builders, so no builder
                                                // ever waits on a global-memory round trip
constexpr int kWgSmem = kWgSlots * kWgSlotB + kWgBands * kWgBandB + 256 + 1024;
static_assert(kWgSmem <= 227 * 1024, "shared memory budget");
// warps: 0-3 epilogue (end of kernel only) | 4-5 MMA issuers | 6 band loader | 7-12 Xp builders (two per channel) | 13-16 dY builders (two per half)
constexpr int kWgXBuilders = 6, kWgDyBuilders = 4, kWgBuilders = kWgXBuilders + kWgDyBuilders;
constexpr int kWgLoader = kEpiWarps + 2, kWgBuilder0 = kWgLoader + 1;
constexpr int kWgThr = (kWgBuilder0 + kWgBuilders) * 32;

struct C1WgParams {
  const float* x;    // [N, 3, H, W]
  const float* dy;   // [N, HO, WO, 32] fp32, or bf16 when dy_bf16 = 1 (converted while the tile is built: the products stay tf32 x tf32)
  int dy_bf16;
  float* partial;    // [gridDim.x][192][32]
  float* bias_partial;  // [gridDim.x][32] or null
  int N, H, W, HO, WO, KR;  // KR = WO + 1 pixel groups rounded up to 8
};

struct C1WgBars {
  uint64_t full[kWgSlots];
  uint64_t empty[kWgSlots];
  uint64_t band_full[kWgBands];
  uint64_t band_empty[kWgBands];
  uint64_t done;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(kWgThr, 1) conv1_band_wgrad_kernel(C1WgParams p, int num_tiles) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* slot_smem = smem;
  unsigned char* band_smem = smem + kWgSlots * kWgSlotB;
  C1WgBars* bars = reinterpret_cast<C1WgBars*>(band_smem + kWgBands * kWgBandB);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kWgSlots; ++s) {
      mbar_init(&bars->full[s], kWgBuilders);
      mbar_init(&bars->empty[s], 1);
    }
    for (int s = 0; s < kWgBands; ++s) {
      mbar_init(&bars->band_full[s], 1);
      mbar_init(&bars->band_empty[s], kWgBuilders);
    }
    mbar_init(&bars->done, 2);
    fence_barrier_init();
  }
  if (warp == kEpiWarps) tmem_alloc(&bars->tmem_base, 128);
  // k-rows the builders never write must contribute zero: pixel groups past WO, row WO of dY_0, row 0 of dY_1
  for (int i = threadIdx.x; i < kWgSlots * kWgSlotB / 16; i += kWgThr) reinterpret_cast<float4*>(slot_smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  for (int i = threadIdx.x; i < kWgSlots * (p.WO + 1); i += kWgThr) {  // the constant ones group of every slot
    const int s_ = i / (p.WO + 1), j = i - s_ * (p.WO + 1);
    *reinterpret_cast<float*>(slot_smem + s_ * kWgSlotB + kWgOnes * kWgGroupB + swz32(j, 0)) = 1.0f;
  }
  fence_proxy_async();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;
  const int my_tiles = ((int)blockIdx.x < num_tiles) ? (num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp < kEpiWarps) {
    // ================================ epilogue: once, after the last tile ================================
    mbar_wait(&bars->done, 0);
    tc_fence_after_sync();
    const int m = warp * 32 + lane;  // accumulator row: (ci, ky, b) for m < 96, the bias row at 96
    float* dst = p.partial + (size_t)blockIdx.x * (kKtot * kCout);
#pragma unroll
    for (int a = 0; a < 2; ++a) {  // columns [32 a, 32 a + 32): the kx half a of the window
      uint32_t v[32], v1[32];
      tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(a * kCout), v);
      tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(64 + a * kCout), v1);
      tmem_ld_wait();
      float o[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) o[j] = (my_tiles > 0 ? __uint_as_float(v[j]) : 0.f) + (my_tiles > 1 ? __uint_as_float(v1[j]) : 0.f);  // issuer 1 ran only if there were two tiles
      if (m == 96 && a == 0 && p.bias_partial) {  // the ones row: sum over this CTA's pixels of dY
#pragma unroll
        for (int j = 0; j < kCout; ++j) p.bias_partial[(size_t)blockIdx.x * kCout + j] = o[j];
      }
      if (m < 96) {
        const int k = (m >> 5) * 64 + ((m >> 2) & 7) * 8 + 4 * a + (m & 3);  // (ci, ky, kx = 4 a + b)
#pragma unroll
        for (int j = 0; j < kCout; j += 4) *reinterpret_cast<float4*>(dst + (size_t)k * kCout + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
      }
    }
  } else if (warp < kWgLoader) {
    // ================================ MMA issuers: issuer me takes the tiles of slot me (it = me, me + 2, ...) into accumulator me ================================
    if (lane == 0) {
      const int me = warp - kEpiWarps;
      constexpr uint32_t idesc = make_idesc_tf32(kBM, 2 * kCout, true, true);
      const uint32_t sb = smem_u32(slot_smem) + me * kWgSlotB;
      const uint32_t d = tmem_base + (uint32_t)(me * 2 * kCout);
      const int ksteps = p.KR >> 3;
      for (int it = me; it < my_tiles; it += 2) {
        mbar_wait(&bars->full[me], (it >> 1) & 1);
        tc_fence_after_sync();
        const uint32_t a = sb, b = sb + kWgDy * kWgGroupB;
        for (int k = 0; k < ksteps; ++k) {
          // MN-major operands: 8 k-rows per MMA = two 512-byte swizzle atoms (SBO), 32-element groups kWgGroupB apart (LBO)
          const uint64_t da = make_smem_desc(a + k * 1024, kWgGroupB, 512u, 1u), db = make_smem_desc(b + k * 1024, kWgGroupB, 512u, 1u);
          umma_tf32(d, da, db, idesc, (uint32_t)(it >= 2 || k != 0));
        }
        umma_commit(&bars->empty[me]);
      }
      umma_commit(&bars->done);
    }
  } else if (warp == kWgLoader) {
    // ================================ band loader ================================
    if (lane == 0) {
      const uint32_t rowbytes = (uint32_t)p.W * 4;
      for (int it = 0; it < my_tiles; ++it) {
        const int tile = blockIdx.x + it * gridDim.x;
        const int slot = it % kWgBands;
        mbar_wait(&bars->band_empty[slot], ((it / kWgBands) & 1) ^ 1);
        const int n = tile / p.HO, y = tile - n * p.HO;
        const uint32_t dy_bytes = (uint32_t)p.WO * kCout * (p.dy_bf16 ? 2u : 4u);
        expect_tx(&bars->band_full[slot], 3 * 8 * rowbytes + dy_bytes);
        const uint32_t dst = smem_u32(band_smem) + slot * kWgBandB;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) bulk_load(dst + (uint32_t)(ci * 8) * rowbytes, p.x + ((size_t)(n * 3 + ci) * p.H + 4 * y) * p.W, 8 * rowbytes, &bars->band_full[slot]);
        bulk_load(dst + kWgXB, reinterpret_cast<const unsigned char*>(p.dy) + ((size_t)(n * p.HO + y) * p.WO) * kCout * (p.dy_bf16 ? 2 : 4), dy_bytes, &bars->band_full[slot]);
      }
    }
  } else {
    // ================================ builders ================================
    // warps 0-5: channel ci = bw / 2, every other 32-chunk batch of group ci; warps 6-9: window half a = (bw - 6) / 2, every other batch of dY_a^T
    const int bw = warp - kWgBuilder0;
    const bool is_x = bw < kWgXBuilders;
    const int grp = is_x ? (bw >> 1) : kWgDy + ((bw - kWgXBuilders) >> 1), half = bw & 1;
    const int shift = is_x ? 0 : ((bw - kWgXBuilders) >> 1);  // dY_1 sits one k-row further down
    const uint32_t rowbytes = (uint32_t)p.W * 4;
    const int nchunks = (is_x ? p.WO + 1 : p.WO) * 8;  // 16-byte chunks of one group: k-row = q / 8, chunk c = q % 8
    constexpr int kIters = kWgKR * 8 / 64;             // 7 chunks per lane at most; all loads are issued before the first store
    for (int it = 0; it < my_tiles; ++it) {
      const int slot = it % kWgSlots, bslot = it % kWgBands;
      const uint32_t dst = smem_u32(slot_smem) + slot * kWgSlotB + grp * kWgGroupB;
      mbar_wait(&bars->empty[slot], ((it / kWgSlots) & 1) ^ 1);
      float4 v[kIters];
      mbar_wait(&bars->band_full[bslot], (it / kWgBands) & 1);
      const uint32_t band = smem_u32(band_smem) + bslot * kWgBandB;
      if (is_x) {
#pragma unroll
        for (int i = 0; i < kIters; ++i) {
          const int q = lane + 32 * (2 * i + half), j = q >> 3, c = q & 7;  // column 4 c + b of group ci: input row ky = c, floats 4 j .. 4 j + 3
          if (q < nchunks) v[i] = ld_shared16(band + (uint32_t)(grp * 8 + c) * rowbytes + (uint32_t)j * 16u);
        }
      } else if (p.dy_bf16) {  // the staged dY row: chunk q = four bf16 = 8 bytes -> four fp32 (a 16-bit shift)
#pragma unroll
        for (int i = 0; i < kIters; ++i) {
          const int q = lane + 32 * (2 * i + half);
          if (q < nchunks) {
            uint32_t lo, hi;
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(band + kWgXB + (uint32_t)q * 8u));
            v[i] = make_float4(__uint_as_float(lo << 16), __uint_as_float(lo & 0xFFFF0000u), __uint_as_float(hi << 16), __uint_as_float(hi & 0xFFFF0000u));
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < kIters; ++i) {
          const int q = lane + 32 * (2 * i + half);
          if (q < nchunks) v[i] = ld_shared16(band + kWgXB + (uint32_t)q * 16u);  // rows of 32 floats are contiguous: chunk q is at q * 16 B
        }
      }
#pragma unroll
      for (int i = 0; i < kIters; ++i) {
        const int q = lane + 32 * (2 * i + half);
        if (q < nchunks) st_shared16(dst + swz32((q >> 3) + shift, q & 7), v[i]);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&bars->full[slot]);
        mbar_arrive(&bars->band_empty[bslot]);
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == kEpiWarps) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 128);
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// Forward without an im2col copy: the tensor core reads the raw input rows.
//
// For a fixed (ci, ky) the im2col row of output pixel (oy, ox) is the 8 consecutive floats x[ci][4 oy + ky][4 ox .. 4 ox + 7]:
// rows of consecutive ox start 16 bytes apart and overlap by half.  That IS a K-major, un-swizzled UMMA operand — core matrix
// = 8 rows x 16 bytes at a 16-byte pitch (128 contiguous bytes), the second K core matrix 16 bytes further (LBO = 16), the
// next 8 rows 128 bytes further (SBO = 128) — so an MMA of K = 8 per (ci, ky) needs no staging pass at all.  To keep the
// 16-byte pitch across output rows the input rows of a work unit are laid out in shared memory by row phase: sub-band
// (p = y mod 4, ci) holds rows y = 4 j + p for j = oy0 .. oy0 + R at a pitch of W floats = W/4 pixels, and with
// m' = (W/4) * (oy - oy0) + ox the operand row of (oy, ox) for tap row ky sits at  sub-band(ky mod 4, ci) + (ky / 4) * 4W + 16 m'.
// Four TMA boxes per unit (one per phase: every 4th row through the tensor map's element stride) write exactly that layout;
// ox = WO .. W/4 - 1 are dummy pixels (1 in 50).
//
// A unit is R output rows of one frame (R * W/4 <= 256 = two M = 128 accumulators: 5 rows of the 200x200 camera, 12 of the 84x84
// one).  Warps: 0-3 epilogue | 4-5 MMA issuers (one per accumulator half, 24 MMAs each per unit) | 6 TMA producer.
// (The weight gradient cannot do the same: its operands are MN-major, and an un-swizzled MN-major tf32 descriptor reads zeros —
// scripts/micro/mn_noswz_probe.cu — so conv1_band_wgrad_kernel keeps its staging warps.)
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kVwSlots = 3;
constexpr int kVwWBytes = 24 * 1024;   // weights: 24 (ci, ky) blocks of [2 k-halves][4 groups of 8 channels][8][4 floats]
constexpr int kVwThr = (kEpiWarps + 3) * 32;

struct V1Params {
  const float* w;
  const float* b;
  float* y;           // channels-last output: fp32, or bf16 (y_bf16 = 1: 32 channels = 64 bytes per pixel)
  int y_bf16;
  unsigned* bits;
  int N, H, W, HO, WO;
  int Wq, R, UPF;        // W / 4, output rows per unit, units per frame
  int sub_bytes;         // (R + 1) * W * 4: one (phase, ci) sub-band
  int ph_stride;         // three sub-bands rounded up to 128 bytes (TMA destinations are 128-byte aligned)
  int slot_bytes;        // 12 sub-bands + slack for the reads of the dummy rows past the last one
  int relu;
};

struct V1Bars {
  uint64_t band_full[kVwSlots];
  uint64_t band_empty[kVwSlots];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(kVwThr, 1) conv1_view_fwd_kernel(const __grid_constant__ CUtensorMap xmap, V1Params p, int num_units) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* w_smem = smem;
  unsigned char* band_smem = smem + kVwWBytes;
  V1Bars* bars = reinterpret_cast<V1Bars*>(band_smem + (size_t)kVwSlots * p.slot_bytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kVwSlots; ++s) {
      mbar_init(&bars->band_full[s], 1);
      mbar_init(&bars->band_empty[s], 2);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bars->tmem_full[a], 2);
      mbar_init(&bars->tmem_empty[a], kEpiWarps);
    }
    fence_barrier_init();
    tma::prefetch_map(&xmap);
  }
  if (warp == kEpiWarps) tmem_alloc(&bars->tmem_base, 128);
  // the slack behind each slot is read (for dummy rows only) but never written by the TMA: keep it finite
  for (int i = threadIdx.x; i < kVwSlots * p.slot_bytes / 16; i += kVwThr) reinterpret_cast<float4*>(band_smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  // resident weights, un-swizzled K-major core matrices: element (n, k = cky * 8 + kx) -> [cky][kx / 4][n / 8][n % 8][kx % 4]
  for (int q = threadIdx.x; q < kCout * 48; q += kVwThr) {
    const int n = q / 48, kq = q - n * 48;  // kq: 16-byte piece of the 192-float row
    const int cky = kq >> 1, half = kq & 1;
    const float4 v = __ldg(reinterpret_cast<const float4*>(p.w + (size_t)n * kKtot + kq * 4));
    *reinterpret_cast<float4*>(w_smem + cky * 1024 + half * 512 + (n >> 3) * 128 + (n & 7) * 16) = v;
  }
  fence_proxy_async();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp < kEpiWarps) {
    // ================================ epilogue: thread = one pixel of each accumulator half ================================
    const int r = warp * 32 + lane;
    float bias[kCout];
#pragma unroll
    for (int j = 0; j < kCout; ++j) bias[j] = p.b ? __ldg(p.b + j) : 0.f;
    int it = 0;
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++it) {
      const int a = it & 1;
      const int n = unit / p.UPF, oy0 = (unit - n * p.UPF) * p.R;
      mbar_wait(&bars->tmem_full[a], (it >> 1) & 1);
      tc_fence_after_sync();
      uint32_t v[2][32];
      tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(a * 64), v[0]);
      tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(a * 64 + 32), v[1]);
      tmem_ld_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->tmem_empty[a]);  // the values are in registers: the accumulators may be reused
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int m = h * 128 + r;
        const int dy = m / p.Wq, ox = m - dy * p.Wq;
        if (dy >= p.R || ox >= p.WO || oy0 + dy >= p.HO) continue;
        const size_t pix = (size_t)(n * p.HO + oy0 + dy) * p.WO + ox;
        float* dst = p.y + pix * kCout;
        unsigned om = 0u;
        float o[kCout];
#pragma unroll
        for (int j = 0; j < kCout; ++j) {
          o[j] = __uint_as_float(v[h][j]) + bias[j];
          if (p.relu) o[j] = fmaxf(o[j], 0.f);
          if (o[j] > 0.f) om |= 1u << j;
        }
        if (p.y_bf16) {
          uint32_t w16[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w16[j]) : "f"(o[2 * j + 1]), "f"(o[2 * j]));
          unsigned char* d16 = reinterpret_cast<unsigned char*>(p.y) + pix * (kCout * 2);
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2)
            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(d16 + 32 * h2), "r"(w16[8 * h2]), "r"(w16[8 * h2 + 1]), "r"(w16[8 * h2 + 2]),
                         "r"(w16[8 * h2 + 3]), "r"(w16[8 * h2 + 4]), "r"(w16[8 * h2 + 5]), "r"(w16[8 * h2 + 6]), "r"(w16[8 * h2 + 7])
                         : "memory");
        } else if ((reinterpret_cast<size_t>(dst) & 31) == 0) {
#pragma unroll
          for (int j = 0; j < kCout; j += 8) st_global_v8(dst + j, o[j], o[j + 1], o[j + 2], o[j + 3], o[j + 4], o[j + 5], o[j + 6], o[j + 7]);
        } else {
#pragma unroll
          for (int j = 0; j < kCout; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
        }
        if (p.bits) p.bits[pix] = om;
      }
    }
  } else if (warp < kEpiWarps + 2) {
    // ================================ MMA issuers: accumulator half h = rows m' = 128 h .. 128 h + 127 ================================
    if (lane == 0) {
      const int h = warp - kEpiWarps;
      constexpr uint32_t idesc = make_idesc_tf32(kBM, kCout, false, false);
      const uint64_t a0 = make_smem_desc(smem_u32(band_smem) + (uint32_t)h * 2048u, 16u, 128u, 0u);
      const uint64_t b0 = make_smem_desc(smem_u32(w_smem), 512u, 128u, 0u);
      uint32_t aoff[24];  // (ci, ky) -> offset of its operand in the slot, in 16-byte units
#pragma unroll
      for (int cky = 0; cky < 24; ++cky) {
        const int ci = cky >> 3, ky = cky & 7;
        aoff[cky] = (uint32_t)((ky & 3) * p.ph_stride + ci * p.sub_bytes + (ky >> 2) * p.W * 4) >> 4;
      }
      const uint32_t slot16 = (uint32_t)p.slot_bytes >> 4;
      int it = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++it) {
        const int a = it & 1, slot = it % kVwSlots;
        mbar_wait(&bars->tmem_empty[a], ((it >> 1) & 1) ^ 1);
        mbar_wait(&bars->band_full[slot], (it / kVwSlots) & 1);
        tc_fence_after_sync();
        const uint32_t d = tmem_base + (uint32_t)(a * 64 + h * 32);
        const uint64_t as = a0 + (uint64_t)(slot * slot16);
#pragma unroll
        for (int cky = 0; cky < 24; ++cky) umma_tf32(d, as + aoff[cky], b0 + (uint32_t)(cky * 64), idesc, (uint32_t)(cky != 0));
        umma_commit(&bars->band_empty[slot]);
        umma_commit(&bars->tmem_full[a]);
      }
    }
  } else {
    // ================================ TMA producer: one box per row phase and unit ================================
    if (lane == 0) {
      const uint32_t box_bytes = 12u * (uint32_t)p.sub_bytes, ph_bytes = (uint32_t)p.ph_stride;
      int it = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++it) {
        const int slot = it % kVwSlots;
        mbar_wait(&bars->band_empty[slot], ((it / kVwSlots) & 1) ^ 1);
        const int n = unit / p.UPF, oy0 = (unit - n * p.UPF) * p.R;
        tma::expect_tx(&bars->band_full[slot], box_bytes);
        const uint32_t dst = smem_u32(band_smem) + (uint32_t)slot * (uint32_t)p.slot_bytes;
#pragma unroll
        for (int ph = 0; ph < 4; ++ph) tma::load_4d(dst + ph * ph_bytes, &xmap, &bars->band_full[slot], 0, 4 * oy0 + ph, 0, n);
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == kEpiWarps) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 128);
  }
}

bool g_conv1_view = [] {
  const char* e = getenv("HULC_B200_CONV1_VIEW");
  return !(e && e[0] == '0');
}();

// cudaErrorNotSupported when the geometry does not fit (the caller then uses the band kernel below).
int conv1_view_fwd(const float* x, const float* w, const float* b, float* y, unsigned* relu_bits, int N, int H, int W, int relu, cudaStream_t st, int y_bf16 = 0) {
  const int HO = (H - 8) / 4 + 1, WO = (W - 8) / 4 + 1;
  if (!g_conv1_view || HO <= 0 || WO <= 0 || (W & 3) || (H & 3) || W > 256 || (reinterpret_cast<size_t>(x) & 15) || (reinterpret_cast<size_t>(w) & 15) ||
      (reinterpret_cast<size_t>(y) & 15))
    return (int)cudaErrorNotSupported;
  V1Params p;
  p.w = w; p.b = b; p.y = y; p.y_bf16 = y_bf16; p.bits = relu_bits; p.N = N; p.H = H; p.W = W; p.HO = HO; p.WO = WO; p.relu = relu;
  p.Wq = W / 4;
  p.R = min(HO, 256 / p.Wq);
  if (p.R < 1) return (int)cudaErrorNotSupported;
  p.UPF = hulc_cdiv(HO, p.R);
  p.sub_bytes = (p.R + 1) * W * 4;
  p.ph_stride = (3 * p.sub_bytes + 127) / 128 * 128;
  // reads reach the last sub-band + one more row + 16 * 255 + 32 bytes into the slot
  p.slot_bytes = (max(4 * p.ph_stride, 3 * p.ph_stride + 2 * p.sub_bytes + W * 4 + 16 * 256 + 32) + 1023) / 1024 * 1024;
  const size_t smem = 1024 + kVwWBytes + (size_t)kVwSlots * p.slot_bytes + sizeof(V1Bars);
  if (smem > 227 * 1024 || 4 * (p.R + 1) > 256) return (int)cudaErrorNotSupported;
  const long long units = (long long)N * p.UPF;
  if (units >= (1ll << 31)) return (int)cudaErrorNotSupported;
  // x[n][ci][y][xx]; a box takes every 4th row (element stride 4) starting at row 4 oy0 + phase: [ci][j][xx] in shared memory.
  // Rows past the frame read as zero.
  CUtensorMap m;
  const uint64_t dims[4] = {(uint64_t)W, (uint64_t)H, 3, (uint64_t)N};
  const uint64_t strides[3] = {(uint64_t)W * 4, (uint64_t)H * W * 4, (uint64_t)3 * H * W * 4};
  const uint32_t box[4] = {(uint32_t)W, (uint32_t)(4 * (p.R + 1)), 3, 1};
  const uint32_t es[4] = {1, 4, 1, 1};
  if (tma::make_map(&m, x, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE, es) != 0) return (int)cudaErrorNotSupported;
  HULC_TRY(cudaFuncSetAttribute(conv1_view_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  HULC_LAUNCH(conv1_view_fwd_kernel, dim3((unsigned)min((long long)kNumSMs, units)), dim3(kVwThr), smem, st, m, p, (int)units);
  HULC_RETURN_LAST();
}

}  // namespace

// Returns cudaErrorNotSupported when the geometry does not fit the band scheme (the caller then uses the gather kernel).
int hulc_conv1_band_fwd(const float* x, const float* w, const float* b, float* y, unsigned* relu_bits, int N, int H, int W, int relu, cudaStream_t st) {
  const int HO = (H - 8) / 4 + 1, WO = (W - 8) / 4 + 1;
  if (HO <= 0 || WO <= 0 || WO > kBM || (W & 3) || (reinterpret_cast<size_t>(x) & 15) || (reinterpret_cast<size_t>(w) & 15) || (reinterpret_cast<size_t>(y) & 15))
    return (int)cudaErrorNotSupported;
  const int rc_view = conv1_view_fwd(x, w, b, y, relu_bits, N, H, W, relu, st);
  if (rc_view != (int)cudaErrorNotSupported) return rc_view;
  C1Params p;
  p.x = x; p.w = w; p.b = b; p.y = y; p.bits = relu_bits; p.N = N; p.H = H; p.W = W; p.HO = HO; p.WO = WO; p.relu = relu;
  p.RT = min(HO, kBM / WO);
  p.TPF = hulc_cdiv(HO, p.RT);
  p.BR = 4 * p.RT + 4;
  if ((size_t)3 * p.BR * W * 4 > (size_t)kBandBytes) return (int)cudaErrorNotSupported;
  const long long tiles = (long long)N * p.TPF;
  if (tiles >= (1ll << 31) || (long long)N * 3 * H * W >= (1ll << 31)) return (int)cudaErrorNotSupported;
  HULC_TRY(cudaFuncSetAttribute(conv1_band_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
  HULC_LAUNCH(conv1_band_fwd_kernel, dim3((unsigned)min((long long)kNumSMs, tiles)), dim3(kThr), kSmem, st, p, (int)tiles);
  HULC_RETURN_LAST();
}

// Per-CTA partial weight gradients [ctas][192][32] of the first layer (k in the reference's (ci, ky, kx) order), followed — when
// want_bias — by per-CTA partial bias gradients [ctas][32]; *ctas_out = number of partials to reduce.  cudaErrorNotSupported when the geometry does not fit (the caller then uses the gather kernel).
int hulc_conv1_band_wgrad_partials(const float* x, const float* dy, float* partial, size_t partial_bytes, int want_bias, int N, int H, int W, int* ctas_out,
                                   cudaStream_t st, int dy_bf16) {
  const int HO = (H - 8) / 4 + 1, WO = (W - 8) / 4 + 1;
  if (HO <= 0 || WO <= 0 || WO + 1 > kWgKR || (W & 3) || (size_t)3 * 8 * W * 4 > (size_t)kWgXB) return (int)cudaErrorNotSupported;
  if ((reinterpret_cast<size_t>(x) | reinterpret_cast<size_t>(dy) | reinterpret_cast<size_t>(partial)) & 15) return (int)cudaErrorNotSupported;
  const long long tiles = (long long)N * HO;
  if (tiles >= (1ll << 31) || (long long)N * 3 * H * W >= (1ll << 31)) return (int)cudaErrorNotSupported;
  const int ctas = (int)min((long long)kNumSMs, tiles);
  if ((size_t)ctas * (kKtot + 1) * kCout * sizeof(float) > partial_bytes) return (int)cudaErrorNotSupported;
  C1WgParams p;
  p.x = x; p.dy = dy; p.dy_bf16 = dy_bf16; p.partial = partial; p.bias_partial = want_bias ? partial + (size_t)ctas * kKtot * kCout : nullptr;  // bias partials follow the weight partials
  p.N = N; p.H = H; p.W = W; p.HO = HO; p.WO = WO; p.KR = (WO + 1 + 7) & ~7;  // WO + 1 = W / 4 pixel groups per input row
  HULC_TRY(cudaFuncSetAttribute(conv1_band_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmem));
  HULC_LAUNCH(conv1_band_wgrad_kernel, dim3(ctas), dim3(kWgThr), kWgSmem, st, p, (int)tiles);
  *ctas_out = ctas;
  HULC_RETURN_LAST();
}

// first layer with a bf16 channels-last output (bf16 path): the view kernel only (H, W multiples of 4)
int hulc_conv1_view_fwd_bf16(const float* x, const float* w, const float* b, void* y, unsigned* relu_bits, int N, int H, int W, int relu, cudaStream_t st) {
  return conv1_view_fwd(x, w, b, reinterpret_cast<float*>(y), relu_bits, N, H, W, relu, st, 1);
}
