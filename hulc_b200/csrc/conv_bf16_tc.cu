// conv_bf16_tc.cu — the perceptual encoders' convolutions with bf16 activations (BASELINE config 3): C-ABI entry points and the
// weight-gradient kernel of the two channels-last layers.
//
//   forward / data gradient : conv_band_kernel<..., BF16> (conv_tma.cu): the staged band is [pixel][64 bf16 = 128 B], every kernel tap a
//                             row-shifted view of it, tcgen05.mma kind::f16; the first layer reads the fp32 NCHW frames as before
//                             (conv1_view_fwd_kernel) and only its epilogue narrows to bf16.
//   weight gradient (here)  : dW[tap][ci][co] = sum over pixels of x[pixel + tap][ci] * dY[pixel][co].  GEMM-K = pixels, so both operands
//                             are MN-major — and a [pixel][64 channels = 128 B] band under the 128-byte swizzle IS the canonical MN-major
//                             SWIZZLE_128B layout of a 16-bit type (8 k-rows per swizzle atom).  So the SAME band the forward kernel
//                             stages is the A operand of every tap: a tap is a view whose start address is shifted by whole rows, and two
//                             taps are stacked into one M = 128 instruction by setting the descriptor's leading-dimension byte offset
//                             (the distance between the two 64-channel groups) to the distance between the two views.  B = the tile of dY
//                             on the band's PW-wide grid, fetched by a TMA box whose out-of-range columns / rows are zero-filled — the
//                             dummy grid columns and ragged last tiles contribute nothing.  A constant "ones" group stacked next to the
//                             last tap makes one accumulator row the column sum of dY = the BIAS gradient, in the same pass.
//                             Accumulators stay in TMEM across all tiles of a CTA; each CTA writes one partial, reduced in a fixed order.
// DRAM traffic: the activation and dY are read exactly once (the gather kernel of conv_tc.cu reads the activation once per 128-row tile
// of the (tap, ci) dimension: 1.9x).
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_pipeline.cuh"
#include "tma.cuh"

// conv_tma.cu / conv1_tc.cu
int hulc_conv_band_bf16_fwd(const void* x, const void* wprep, const float* b, void* y, unsigned* relu_bits, int N, int CIN, int H, int W, int COUT, int KS, int S,
                            int relu, cudaStream_t st);
int hulc_conv_band_bf16_dgrad(const void* dy, const void* wprep, const unsigned* gate_bits, void* dx, int N, int CIN, int H, int W, int COUT, int HO, int WO, int KS,
                              int S, cudaStream_t st);
int hulc_conv1_view_fwd_bf16(const float* x, const float* w, const float* b, void* y, unsigned* relu_bits, int N, int H, int W, int relu, cudaStream_t st);
int hulc_conv1_band_wgrad_partials(const float* x, const float* dy, float* partial, size_t partial_bytes, int want_bias, int N, int H, int W, int* ctas_out,
                                   cudaStream_t st, int dy_bf16);
int hulc_conv1_wgrad_reduce(const float* partial, float* dw, float* db, int ctas, float beta, cudaStream_t st);  // conv_tc.cu

namespace {

using namespace tc;

constexpr size_t kCounterFloats = 1024;  // head of the shared workspace reserved for tickets (conv_tc.cu)
constexpr int kBandRowsW = 184, kBandBW = kBandRowsW * kRowBytes;  // as conv_tma.cu
constexpr int kDyRows = 128, kDyB = kDyRows * kRowBytes;
constexpr int kOnesRows = 136, kOnesB = kOnesRows * kRowBytes;  // 128 k-rows + the 8-row alignment slack of a view
constexpr int kWgSlots = 3;
constexpr int kMaxJobs = 6;
constexpr int kNIw = 2;  // MMA-issuing threads
constexpr int kThrW = (kEpiWarps + kNIw + 1) * 32;

struct WgJob {
  int band;      // which band of the slot the first group views
  int off_rows;  // row shift of the first group's view (tap shift)
  int lbo_rows;  // second group = the view lbo_rows further down the same band; -1: the constant ones group
};

struct WgParams {
  int N, OH, PW, PH, RT, TPF;
  int nbands, es_y;
  int by[2];
  int njobs;
  WgJob jobs[kMaxJobs];
  float* partial;  // [ctas][njobs * 128][64]
};

struct WgBars {
  uint64_t full[kWgSlots];
  uint64_t empty[kWgSlots];
  uint64_t done;
  uint32_t tmem_base;
};

__host__ __device__ constexpr uint32_t idesc_bf16_mn(int M, int N) {  // kind::f16, bf16 x bf16 -> fp32, both operands MN-major
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int NBANDS>
__global__ void __launch_bounds__(kThrW, 1) conv_wgrad_bf16_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap dymap, WgParams p,
                                                                    int num_tiles) {
  constexpr int kSlotB = NBANDS * kBandBW + kDyB;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* slots = smem;
  unsigned char* ones = smem + kWgSlots * kSlotB;
  WgBars* bars = reinterpret_cast<WgBars*>(ones + kOnesB);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kWgSlots; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], kNIw);
    }
    mbar_init(&bars->done, kNIw);
    fence_barrier_init();
  }
  if (warp == kEpiWarps) tmem_alloc(&bars->tmem_base, 512);
  // rows of a slot the TMA boxes never write (past the band / past the dY tile) are read by the 16-row k-steps: keep them zero
  for (int i = threadIdx.x; i < (kWgSlots * kSlotB + kOnesB) / 16; i += kThrW) reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  // the ones group: logical element 0 of every k-row is 1.0 (bf16 0x3F80).  Under the 128-byte swizzle the 16-byte chunk c of the row
  // at address a sits at chunk c ^ ((a >> 7) & 7): chunk 0 of row r (the region is 1024-byte aligned) is stored at chunk r & 7.
  for (int r = threadIdx.x; r < kOnesRows; r += kThrW) *reinterpret_cast<unsigned short*>(ones + r * kRowBytes + ((r & 7) << 4)) = 0x3F80;
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
    const int r = warp * 32 + lane;
    float* dst = p.partial + (size_t)blockIdx.x * ((size_t)p.njobs * 128 * 64);
    for (int j = 0; j < p.njobs; ++j) {
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(j * 64 + c0), v);
        tmem_ld_wait();
        float* d = dst + ((size_t)j * 128 + r) * 64 + c0;
#pragma unroll
        for (int q = 0; q < 32; q += 4)
          *reinterpret_cast<float4*>(d + q) = my_tiles ? make_float4(__uint_as_float(v[q]), __uint_as_float(v[q + 1]), __uint_as_float(v[q + 2]), __uint_as_float(v[q + 3]))
                                                       : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  } else if (warp < kEpiWarps + kNIw) {
    // ================================ MMA issuers: jobs me, me + 2, ... ================================
    if (lane == 0) {
      const int me = warp - kEpiWarps;
      constexpr uint32_t idesc = idesc_bf16_mn(kBM, 64);
      const uint32_t sb = smem_u32(slots), ones_a = smem_u32(ones);
      for (int it = 0; it < my_tiles; ++it) {
        const int tile = blockIdx.x + it * gridDim.x;
        const int slot = it % kWgSlots;
        const int y0 = (tile % p.TPF) * p.RT;
        const int ksteps = (min(p.RT, p.OH - y0) * p.PW + 15) >> 4;  // 16 pixels per instruction; rows past the tile's pixels are zero in dY
        mbar_wait(&bars->full[slot], (it / kWgSlots) & 1);
        tc_fence_after_sync();
        const uint32_t slot_a = sb + slot * kSlotB;
        const uint32_t dy_a = slot_a + NBANDS * kBandBW;
        for (int j = me; j < p.njobs; j += kNIw) {
          const WgJob jb = p.jobs[j];
          const uint32_t a = slot_a + jb.band * kBandBW + (uint32_t)jb.off_rows * kRowBytes;
          // group 1 of the M = 128 operand: another tap's view of the same band, or the ones group (at a fixed address: the distance to it
          // does not depend on the k-step because both views advance by the same 16 rows ... the ones rows are all alike)
          const uint32_t lbo = jb.lbo_rows >= 0 ? (uint32_t)jb.lbo_rows * kRowBytes : ones_a - a;
          const uint32_t d = tmem_base + (uint32_t)(j * 64);
          for (int k = 0; k < ksteps; ++k) {
            const uint64_t da = make_smem_desc(a + k * 2048, jb.lbo_rows >= 0 ? lbo : lbo - k * 2048, 1024u, 2u);
            const uint64_t db = make_smem_desc(dy_a + k * 2048, 0u, 1024u, 2u);
            umma_f16(d, da, db, idesc, (uint32_t)(it != 0 || k != 0));
          }
        }
        umma_commit(&bars->empty[slot]);
      }
      umma_commit(&bars->done);
    }
  } else {
    // ================================ TMA producer ================================
    if (lane == 0) {
      tma::prefetch_map(&xmap);
      tma::prefetch_map(&dymap);
      const uint32_t bytes = (uint32_t)(NBANDS * p.PW * p.PH + p.PW * p.RT) * kRowBytes;
      for (int it = 0; it < my_tiles; ++it) {
        const int tile = blockIdx.x + it * gridDim.x;
        const int slot = it % kWgSlots;
        const int n = tile / p.TPF, y0 = (tile - n * p.TPF) * p.RT;
        mbar_wait(&bars->empty[slot], ((it / kWgSlots) & 1) ^ 1);
        tma::expect_tx(&bars->full[slot], bytes);
        const uint32_t dst = smem_u32(slots) + slot * kSlotB;
#pragma unroll
        for (int b = 0; b < NBANDS; ++b) tma::load_4d(dst + b * kBandBW, &xmap, &bars->full[slot], 0, 0, p.es_y * y0 + p.by[b], n);
        tma::load_4d(dst + NBANDS * kBandBW, &dymap, &bars->full[slot], 0, 0, y0, n);
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

// dw[co][ci][ky][kx] = beta * dw + sum over CTAs of their partial rows; db[co] += the ones row.  8 lanes per output element.
//   mode 3 (64 -> 64, 3x3): job j stacks taps 2j, 2j + 1 (job 4: tap 8 and the ones group): row = j * 128 + h * 64 + ci
//   mode 2 (32 -> 64, 4x4, stride 2): job j = (py, ty) stacks local taps tx = 0, 1 of band py; a band row is the pixel pair (px, ci):
//           row = j * 128 + tx * 64 + px * 32 + ci with ky = 2 ty + py, kx = 2 tx + px; job 4 = two ones groups
__global__ void wgrad_bf16_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, float* __restrict__ db, int ctas, int njobs, int mode, float beta) {
  const int KS = mode == 3 ? 3 : 4, CIN = mode == 3 ? 64 : 32;
  const int ktot = KS * KS * CIN;
  const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, l = threadIdx.x & 7;
  const int total = (ktot + 1) * 64;  // + the bias row
  const bool live = e < total;
  const int k = live ? e >> 6 : 0, co = e & 63;
  int row;
  int ky = 0, kx = 0, ci = 0;
  if (k == ktot) {
    row = 4 * 128 + 64;  // job 4, second group = the ones group
  } else if (mode == 3) {
    const int tap = k / 64;
    ci = k - tap * 64; ky = tap / 3; kx = tap - ky * 3;
    row = (tap >> 1) * 128 + (tap & 1) * 64 + ci;
  } else {
    const int tap = k / 32;
    ci = k - tap * 32; ky = tap >> 2; kx = tap & 3;
    const int py = ky & 1, ty = ky >> 1, px = kx & 1, tx = kx >> 1;
    row = (py * 2 + ty) * 128 + tx * 64 + px * 32 + ci;
  }
  float s = 0.f;
  if (live)
    for (int i = l; i < ctas; i += 8) s += partial[((size_t)i * njobs * 128 + row) * 64 + co];
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  if (!live || l != 0) return;
  if (k == ktot) {
    if (db) db[co] += s;
    return;
  }
  float* d = dw + (((size_t)co * CIN + ci) * KS + ky) * KS + kx;
  *d = (beta != 0.f ? beta * *d : 0.f) + s;
}

// forward weights: wf[co][(ky,kx,ci)] = w[co][ci][ky][kx]; data-gradient weights: wd[ph][ci][(jy,jx,co)] = w[co][ci][py+S*jy][px+S*jx] (as conv_tc.cu), bf16
__global__ void prep_fwd_bf16_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wf, int COUT, int CIN, int KS) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int K = CIN * KS * KS;
  if (i >= COUT * K) return;
  const int co = i / K, k = i - co * K;
  const int tap = k / CIN, ci = k - tap * CIN;
  wf[i] = __float2bfloat16_rn(w[(size_t)co * K + ci * KS * KS + tap]);
}
__global__ void prep_dgrad_bf16_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wd, int COUT, int CIN, int KS, int S) {
  const int R = KS / S, Kp = R * R * COUT;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S * S * CIN * Kp) return;
  const int k = i % Kp, ci = (i / Kp) % CIN, ph = i / (Kp * CIN);
  const int py = ph / S, px = ph - py * S;
  const int tap = k / COUT, co = k - tap * COUT;
  const int jy = tap / R, jx = tap - jy * R;
  wd[i] = __float2bfloat16_rn(w[(((size_t)co * CIN + ci) * KS + (py + S * jy)) * KS + (px + S * jx)]);
}

int band_map(CUtensorMap* m, const void* src, int N, int SH, int pos_w, size_t row_bytes, size_t frame_bytes, int box_w, int box_h, int es_y) {
  const uint64_t dims[4] = {64, (uint64_t)pos_w, (uint64_t)SH, (uint64_t)N};
  const uint64_t strides[3] = {128, (uint64_t)row_bytes, (uint64_t)frame_bytes};
  const uint32_t box[4] = {64, (uint32_t)box_w, (uint32_t)(box_h * es_y), 1};
  const uint32_t es[4] = {1, 1, (uint32_t)es_y, 1};
  return tma::make_map(m, src, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, es, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
}

template <int NBANDS>
int launch_wgrad(const CUtensorMap& xm, const CUtensorMap& dm, const WgParams& p, int tiles, int ctas, cudaStream_t st) {
  constexpr int smem = kWgSlots * (NBANDS * kBandBW + kDyB) + kOnesB + 256 + 1024;
  static_assert(smem <= 227 * 1024, "shared memory budget");
  auto kfn = conv_wgrad_bf16_kernel<NBANDS>;
  HULC_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  HULC_LAUNCH(kfn, dim3(ctas), dim3(kThrW), smem, st, xm, dm, p, tiles);
  HULC_RETURN_LAST();
}

}  // namespace

// ---- C ABI --------------------------------------------------------------------------------------------------------------------------
// Layer 1 (CIN = 3): x = the reference's fp32 NCHW frames; layers 2 / 3: x = bf16 NHWC.  y = bf16 NHWC, bias fp32, relu_bits as
// hulc_conv2d_tc_fwd.  cudaErrorNotSupported: the geometry does not fit the band scheme (the engine then stays on the fp32-activation path).
HULC_API int hulc_conv2d_bf16_fwd(const void* x, const float* w, const float* b, void* y, int N, int CIN, int H, int W, int COUT, int KS, int S, int relu,
                                  unsigned* relu_bits, float* workspace, size_t workspace_bytes, void* stream) {
  if (N <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (CIN == 3 && COUT == 32 && KS == 8 && S == 4) return hulc_conv1_view_fwd_bf16(reinterpret_cast<const float*>(x), w, b, y, relu_bits, N, H, W, relu, st);
  const int K = CIN * KS * KS;
  if (workspace_bytes < kCounterFloats * sizeof(float) + (size_t)COUT * K * 2) return (int)cudaErrorInvalidValue;
  __nv_bfloat16* wp = reinterpret_cast<__nv_bfloat16*>(workspace + kCounterFloats);
  HULC_LAUNCH(prep_fwd_bf16_kernel, dim3(hulc_cdiv(COUT * K, 256)), dim3(256), 0, st, w, wp, COUT, CIN, KS);
  return hulc_conv_band_bf16_fwd(x, wp, b, y, relu_bits, N, CIN, H, W, COUT, KS, S, relu, st);
}

HULC_API int hulc_conv2d_bf16_dgrad(const void* dy, const float* w, const unsigned* gate_bits, void* dx, int N, int CIN, int H, int W, int COUT, int KS, int S,
                                    float* workspace, size_t workspace_bytes, void* stream) {
  if (N <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (KS % S != 0 || !gate_bits) return (int)cudaErrorInvalidValue;
  const int R = KS / S, Kp = R * R * COUT, n = S * S * CIN * Kp;
  if (workspace_bytes < kCounterFloats * sizeof(float) + (size_t)n * 2) return (int)cudaErrorInvalidValue;
  __nv_bfloat16* wp = reinterpret_cast<__nv_bfloat16*>(workspace + kCounterFloats);
  HULC_LAUNCH(prep_dgrad_bf16_kernel, dim3(hulc_cdiv(n, 256)), dim3(256), 0, st, w, wp, COUT, CIN, KS, S);
  return hulc_conv_band_bf16_dgrad(dy, wp, gate_bits, dx, N, CIN, H, W, COUT, (H - KS) / S + 1, (W - KS) / S + 1, KS, S, st);
}

// dw (fp32, reference layout [COUT][CIN][KS][KS]) = beta * dw + dL/dw; db (fp32, optional) += sum over pixels of dy.
// Layer 1: x = fp32 NCHW frames, dy = bf16 NHWC [N][HO][WO][32]; layers 2 / 3: x, dy bf16 NHWC.
HULC_API int hulc_conv2d_bf16_wgrad(const void* x, const void* dy, float* dw, float beta, float* db, int N, int CIN, int H, int W, int COUT, int KS, int S,
                                    float* workspace, size_t workspace_bytes, void* stream) {
  if (N <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (workspace_bytes <= kCounterFloats * sizeof(float)) return (int)cudaErrorInvalidValue;
  float* ws = workspace + kCounterFloats;
  const size_t wsb = workspace_bytes - kCounterFloats * sizeof(float);
  const int OH = (H - KS) / S + 1, OW = (W - KS) / S + 1;
  if (CIN == 3 && COUT == 32 && KS == 8 && S == 4) {
    int ctas = 0;
    HULC_TRY(hulc_conv1_band_wgrad_partials(reinterpret_cast<const float*>(x), reinterpret_cast<const float*>(dy), ws, wsb, db != nullptr, N, H, W, &ctas, st, 1));
    return hulc_conv1_wgrad_reduce(ws, dw, db, ctas, beta, st);  // the fp32 path's reduction (conv_tc.cu): same partial layout
  }
  if (COUT != 64 || OH <= 0 || OW <= 0) return (int)cudaErrorNotSupported;
  WgParams p{};
  p.N = N; p.OH = OH;
  CUtensorMap xm, dm;
  int mode, nb;
  if (KS == 3 && S == 1 && CIN == 64) {
    mode = 3; nb = 1;
    p.PW = W; p.RT = min(OH, kBM / p.PW); p.PH = p.RT + 2; p.TPF = hulc_cdiv(OH, p.RT); p.nbands = 1; p.es_y = 1; p.by[0] = 0;
    p.njobs = 5;
    for (int j = 0; j < 5; ++j) {
      const int t0 = 2 * j, t1 = 2 * j + 1;
      const int s0 = (t0 / 3) * p.PW + t0 % 3, s1 = (t1 / 3) * p.PW + t1 % 3;
      p.jobs[j] = WgJob{0, s0, j < 4 ? s1 - s0 : -1};
    }
    if (band_map(&xm, x, N, H, W, (size_t)W * 128, (size_t)H * W * 128, p.PW, p.PH, 1) != 0) return (int)cudaErrorNotSupported;
  } else if (KS == 4 && S == 2 && CIN == 32) {
    mode = 2; nb = 2;
    p.PW = W / 2; p.RT = min(OH, kBM / p.PW); p.PH = p.RT + 1; p.TPF = hulc_cdiv(OH, p.RT); p.nbands = 2; p.es_y = 2; p.by[0] = 0; p.by[1] = 1;
    if (p.PW < OW + 1) return (int)cudaErrorNotSupported;
    p.njobs = db ? 5 : 4;
    for (int py = 0; py < 2; ++py)
      for (int ty = 0; ty < 2; ++ty) p.jobs[py * 2 + ty] = WgJob{py, ty * p.PW, 1};  // taps (ty, 0) and (ty, 1): one row apart
    p.jobs[4] = WgJob{0, 0, -1};  // bias only: rows 64.. = the ones group (rows 0..63 repeat a tap and are ignored)
    if (band_map(&xm, x, N, H, W / 2, (size_t)W * 64, (size_t)H * W * 64, p.PW, p.PH, 2) != 0) return (int)cudaErrorNotSupported;
  } else {
    return (int)cudaErrorNotSupported;
  }
  if (p.PW > kBM || p.PW * p.PH > kBandRowsW || p.RT * p.PW > kDyRows || p.PH * p.es_y > 256) return (int)cudaErrorNotSupported;
  // a view starts (taps) up to 2 PW + 2 rows into the band and runs up to 128 rows: it must stay inside the band's 184-row slot
  if (127 + 2 * p.PW + 2 >= kBandRowsW) return (int)cudaErrorNotSupported;
  if (band_map(&dm, dy, N, OH, OW, (size_t)OW * 128, (size_t)OH * OW * 128, p.PW, p.RT, 1) != 0) return (int)cudaErrorNotSupported;
  const long long tiles = (long long)N * p.TPF;
  if (tiles >= (1ll << 31)) return (int)cudaErrorNotSupported;
  const int ctas = (int)min((long long)kNumSMs, tiles);
  const size_t need = (size_t)ctas * p.njobs * 128 * 64 * sizeof(float);
  if (need > wsb) return (int)cudaErrorInvalidValue;
  p.partial = ws;
  HULC_TRY(nb == 1 ? launch_wgrad<1>(xm, dm, p, (int)tiles, ctas, st) : launch_wgrad<2>(xm, dm, p, (int)tiles, ctas, st));
  const int ktot = KS * KS * CIN;
  HULC_LAUNCH(wgrad_bf16_reduce_kernel, dim3(hulc_cdiv((ktot + 1) * 64 * 8, 256)), dim3(256), 0, st, (const float*)ws, dw, (mode == 3 || p.njobs == 5) ? db : nullptr, ctas,
              p.njobs, mode, beta);
  HULC_RETURN_LAST();
}
