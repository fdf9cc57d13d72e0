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

}  // namespace

// Returns cudaErrorNotSupported when the geometry does not fit the band scheme (the caller then uses the gather kernel).
int hulc_conv1_band_fwd(const float* x, const float* w, const float* b, float* y, unsigned* relu_bits, int N, int H, int W, int relu, cudaStream_t st) {
  const int HO = (H - 8) / 4 + 1, WO = (W - 8) / 4 + 1;
  if (HO <= 0 || WO <= 0 || WO > kBM || (W & 3) || (reinterpret_cast<size_t>(x) & 15) || (reinterpret_cast<size_t>(w) & 15) || (reinterpret_cast<size_t>(y) & 15))
    return (int)cudaErrorNotSupported;
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
