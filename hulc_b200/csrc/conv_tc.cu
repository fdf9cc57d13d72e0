// conv_tc.cu — the three convolutions of the perceptual encoders (vision_network.py:36-47, vision_network_gripper.py:11-17)
// as implicit GEMMs on the tensor cores (tcgen05.mma kind::tf32, fp32 accumulation in TMEM; main loop in tc_pipeline.cuh).
//
// Activations are kept channels-last (NHWC) between the layers: the GEMM view of a convolution is
//   out[pixel][cout] = sum_{tap, ci} in[pixel + tap][ci] * w[cout][tap][ci]
// so with NHWC every 16-byte cp.async of the A operand is 4 consecutive channels of one tap, a k-block of 32 is one tap
// (CIN = 32) or half a tap (CIN = 64), and the accumulator tile [128 pixels x COUT] is written back as contiguous rows.
// The first layer reads the reference's NCHW fp32 frames directly (CIN = 3: the 8 kx of a (ci, ky) pair are contiguous).
//   forward : A = im2col(x) [pixels x K] (K-major tiles), B = w reordered to [COUT][tap][ci]; epilogue bias + ReLU
//   dgrad   : per stride phase, A = taps of dY [phase pixels x (tap, cout)], B = w reordered to [CIN][tap][cout]; the
//             epilogue applies the ReLU mask of the activation that fed this conv
//   wgrad   : dW[k][cout] = sum_pixels im2col(x)[pixel][k] * dY[pixel][cout]: both operands are read with the GEMM-K
//             dimension (pixels) strided -> MN-major tiles; the pixel range is split over CTAs, partial sums are reduced
//             in a fixed order and written back in the reference's [COUT][CIN][KS][KS] layout.
// Operands are consumed as tf32 (10-bit mantissa), as cuDNN does by default for the reference's convolutions on
// Ampere-or-newer GPUs; the effect on the parity metrics is quantified in DESIGN.md.
#include <cstdlib>

#include "common.cuh"
#include "tc_pipeline.cuh"

// conv_tma.cu: the same layers with the A operand delivered by TMA (cudaErrorNotSupported -> use the gather kernels below)
int hulc_conv_tma_fwd(const float* x, const float* wprep, const float* b, float* y, unsigned* relu_bits, int N, int CIN, int H, int W, int COUT, int KS, int S,
                      int relu, cudaStream_t st);
int hulc_conv_tma_dgrad_s2_all(const float* dy, const float* wall, const float* gate, const unsigned* gate_bits, float* dx, int N, int CIN, int H, int W, int COUT,
                               int HO, int WO, cudaStream_t st);
int hulc_conv_tma_dgrad_phase(const float* dy, const float* wphase, const float* gate, const unsigned* gate_bits, float* dx, int N, int CIN, int H, int W, int COUT,
                              int HO, int WO, int R, int S, int py, int px, cudaStream_t st);

int hulc_conv1_band_wgrad_partials(const float* x, const float* dy, float* partial, size_t partial_bytes, int want_bias, int N, int H, int W, int* ctas_out,
                                   cudaStream_t st, int dy_bf16 = 0);
extern "C" int hulc_colsum(const float* X, int rows, int cols, int ldx, float* out, float beta, float* workspace, size_t workspace_bytes, void* stream);
int hulc_conv1_band_fwd(const float* x, const float* w, const float* b, float* y, unsigned* relu_bits, int N, int H, int W, int relu, cudaStream_t st);  // conv1_tc.cu

namespace {

using tc::cp_async16;
using tc::kProdThreads;
using tc::swz;
using tc::swz32;

constexpr uint32_t kInvalid = 0xFFFFFFFFu;
// HULC_B200_CONV_TMA=0 keeps the cp.async gather kernels (A/B comparison, fallback)
const bool g_use_tma = [] { const char* e = getenv("HULC_B200_CONV_TMA"); return !(e && e[0] == '0'); }();
constexpr size_t kCounterFloats = 1024;  // head of the shared workspace reserved for the split-K / loss tickets

struct Geom {
  int N, H, W, CIN, HO, WO, COUT;
};

// ---------------------------------------------------------------------------------------------------------------------
// forward A operand, NHWC input
// ---------------------------------------------------------------------------------------------------------------------
template <int CIN, int KS, int S>
struct FwdNhwcLoader {
  static constexpr bool kMNMajor = false;
  const float* x;
  Geom g;
  int M;
  uint32_t rowoff[4];
  __device__ __forceinline__ void start_tile(int tile, int ptid) {
    const int P = g.HO * g.WO;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = tile * tc::kBM + (ptid >> 3) + 32 * i;
      uint32_t off = kInvalid;
      if (m < M) {
        const int n = m / P, p = m - n * P;
        const int y = p / g.WO, xx = p - y * g.WO;
        off = (uint32_t)(((n * g.H + y * S) * g.W + xx * S) * CIN);
      }
      rowoff[i] = off;
    }
  }
  __device__ __forceinline__ void issue(int kb, uint32_t dst, int ptid) const {
    constexpr int KB_PER_TAP = CIN / 32;
    const int c = ptid & 7;
    const int tap = kb / KB_PER_TAP, ci0 = (kb % KB_PER_TAP) * 32;
    const int ky = tap / KS, kx = tap - ky * KS;
    const uint32_t off = (uint32_t)((ky * g.W + kx) * CIN + ci0 + c * 4);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = (ptid >> 3) + 32 * i;
      const bool ok = rowoff[i] != kInvalid;
      cp_async16(dst + swz(r, c), ok ? (const void*)(x + rowoff[i] + off) : (const void*)x, ok);
    }
  }
};

// forward A operand of the first layer: NCHW input with 3 channels, K = (ci, ky, kx), kx contiguous
template <int KS, int S>
struct FwdNchw3Loader {
  static constexpr bool kMNMajor = false;
  const float* x;
  Geom g;
  int M;
  uint32_t rowoff[4];
  __device__ __forceinline__ void start_tile(int tile, int ptid) {
    const int P = g.HO * g.WO;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = tile * tc::kBM + (ptid >> 3) + 32 * i;
      uint32_t off = kInvalid;
      if (m < M) {
        const int n = m / P, p = m - n * P;
        const int y = p / g.WO, xx = p - y * g.WO;
        off = (uint32_t)((n * 3 * g.H + y * S) * g.W + xx * S);
      }
      rowoff[i] = off;
    }
  }
  __device__ __forceinline__ void issue(int kb, uint32_t dst, int ptid) const {
    static_assert(KS == 8, "k-block = 4 (ci, ky) pairs of 8 kx");
    const int c = ptid & 7;
    const int pair = kb * 4 + (c >> 1);
    const int ci = pair >> 3, ky = pair & 7;
    const uint32_t off = (uint32_t)((ci * g.H + ky) * g.W + (c & 1) * 4);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = (ptid >> 3) + 32 * i;
      const bool ok = rowoff[i] != kInvalid;
      cp_async16(dst + swz(r, c), ok ? (const void*)(x + rowoff[i] + off) : (const void*)x, ok);
    }
  }
};

// B operand for forward / dgrad: a dense [ROWS][K] K-major matrix (prepared weights), the same rows for every tile
template <int ROWS>
struct WeightLoader {
  static constexpr bool kMNMajor = false;
  const float* w;
  int K;
  __device__ __forceinline__ void start_tile(int, int) {}
  __device__ __forceinline__ void issue(int kb, uint32_t dst, int ptid) const {
#pragma unroll
    for (int q = ptid; q < ROWS * 8; q += kProdThreads) {
      const int r = q >> 3, c = q & 7;
      cp_async16(dst + swz(r, c), w + (size_t)r * K + kb * tc::kBK + c * 4, true);
    }
  }
};

struct FwdEpilogue {
  float* y;
  const float* bias;
  int M, COUT, relu;
  unsigned* bits;  // optional sign mask of the output: word [pixel][channel / 32]
  __device__ __forceinline__ void operator()(int tile, int row, int col0, const float* v) const {
    const int m = tile * tc::kBM + row;
    if (m >= M) return;
    float* dst = y + (size_t)m * COUT + col0;
    unsigned om = 0u;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      float4 o;
      o.x = v[j] + bias[col0 + j]; o.y = v[j + 1] + bias[col0 + j + 1]; o.z = v[j + 2] + bias[col0 + j + 2]; o.w = v[j + 3] + bias[col0 + j + 3];
      if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
      if (o.x > 0.f) om |= 1u << j;
            if (o.y > 0.f) om |= 1u << (j + 1);
            if (o.z > 0.f) om |= 1u << (j + 2);
            if (o.w > 0.f) om |= 1u << (j + 3);
      *reinterpret_cast<float4*>(dst + j) = o;
    }
    if (bits) bits[((size_t)m * COUT + col0) >> 5] = om;
  }
};

// ---------------------------------------------------------------------------------------------------------------------
// data gradient, one stride phase (py, px): rows = input pixels (S*y2+py, S*x2+px); K = (jy, jx, cout) with
// dX[.., ci] = sum dY[n, y2-jy, x2-jx, co] * W[co][ci][py+S*jy][px+S*jx]
// ---------------------------------------------------------------------------------------------------------------------
template <int COUT, int KS, int S>
struct DgradLoader {
  static constexpr bool kMNMajor = false;
  static constexpr int R = KS / S;
  const float* dy;
  Geom g;
  int M, HP, WP;
  int rowoff[4];
  int rowyx[4];
  __device__ __forceinline__ void start_tile(int tile, int ptid) {
    const int P = HP * WP;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = tile * tc::kBM + (ptid >> 3) + 32 * i;
      int off = 0, yx = -1;
      if (m < M) {
        const int n = m / P, p = m - n * P;
        const int y2 = p / WP, x2 = p - y2 * WP;
        off = ((n * g.HO + y2) * g.WO + x2) * COUT;
        yx = (y2 << 16) | x2;
      }
      rowoff[i] = off; rowyx[i] = yx;
    }
  }
  __device__ __forceinline__ void issue(int kb, uint32_t dst, int ptid) const {
    constexpr int KB_PER_TAP = COUT / 32;
    const int c = ptid & 7;
    const int tap = kb / KB_PER_TAP, co0 = (kb % KB_PER_TAP) * 32;
    const int jy = tap / R, jx = tap - jy * R;
    const int off = -(jy * g.WO + jx) * COUT + co0 + c * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = (ptid >> 3) + 32 * i;
      const int oy = (rowyx[i] >> 16) - jy, ox = (rowyx[i] & 0xFFFF) - jx;
      const bool ok = rowyx[i] >= 0 && oy >= 0 && oy < g.HO && ox >= 0 && ox < g.WO;
      cp_async16(dst + swz(r, c), ok ? (const void*)(dy + rowoff[i] + off) : (const void*)dy, ok);
    }
  }
};

template <int S>
struct DgradEpilogue {
  float* dx;
  const float* gate;
  Geom g;
  int M, HP, WP, py, px;
  __device__ __forceinline__ void operator()(int tile, int row, int col0, const float* v) const {
    const int m = tile * tc::kBM + row;
    if (m >= M) return;
    const int P = HP * WP;
    const int n = m / P, p = m - n * P;
    const int y2 = p / WP, x2 = p - y2 * WP;
    const size_t off = ((size_t)(n * g.H + S * y2 + py) * g.W + S * x2 + px) * g.CIN + col0;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      if (gate) {
        const float4 a = *reinterpret_cast<const float4*>(gate + off + j);
        o.x = a.x > 0.f ? o.x : 0.f; o.y = a.y > 0.f ? o.y : 0.f; o.z = a.z > 0.f ? o.z : 0.f; o.w = a.w > 0.f ? o.w : 0.f;
      }
      *reinterpret_cast<float4*>(dx + off + j) = o;
    }
  }
};

// ---------------------------------------------------------------------------------------------------------------------
// weight gradient: D[k][co] = sum_pixels X_col[pixel][k] * dY[pixel][co]; tile = (k-tile tm, pixel split s)
// ---------------------------------------------------------------------------------------------------------------------
struct WgradTiling {
  int splits, pix_per_split, M;
  __device__ __forceinline__ void decode(int tile, int& tm, int& pix0, int& pix_end) const {
    tm = tile / splits;
    const int s = tile - tm * splits;
    pix0 = s * pix_per_split;
    pix_end = min(M, pix0 + pix_per_split);
  }
};

// A operand (MN-major, 128 k's x 32 pixels).  NCHW3 = first layer (k = ci*64 + ky*8 + kx), otherwise NHWC (k = tap*CIN + ci).
// Thread (p = ptid >> 3, c4 = ptid & 7) owns pixel p of every k-block and chunk c4 of each of the 4 k-groups: the group offsets are
// fixed per work item, the pixel advances by 32 per k-block (tracked incrementally: no divisions in issue()).
template <int CIN, int KS, int S, bool NCHW3>
struct WgradXLoader {
  static constexpr bool kMNMajor = true;
  static constexpr int KTOT = CIN * KS * KS;
  const float* x;
  Geom g;
  WgradTiling t;
  int tm, pix0, pix_end;
  int pix, n, y, xx, next_kb;  // this thread's pixel of k-block next_kb
  uint32_t goff[4];
  uint32_t gokmask;
  __device__ __forceinline__ void start_tile(int tile, int ptid) {
    t.decode(tile, tm, pix0, pix_end);
    const int c4 = ptid & 7;
    pix = pix0 + (ptid >> 3);
    const int P = g.HO * g.WO;
    n = pix / P;
    const int q = pix - n * P;
    y = q / g.WO; xx = q - y * g.WO;
    next_kb = 0;
    gokmask = 0;
#pragma unroll
    for (int grp = 0; grp < 4; ++grp) {
      if (NCHW3) {
        const int k = tm * tc::kBM + grp * 32 + (c4 >> 1) * 8;  // (ci, ky) pair start
        const int ci = k / (KS * KS), ky = (k / KS) % KS;
        goff[grp] = (uint32_t)((ci * g.H + ky) * g.W + (c4 & 1) * 4);
        gokmask |= (k < KTOT ? 1u : 0u) << grp;
      } else {
        const int k = tm * tc::kBM + grp * 32;
        const int tap = k / CIN, ci0 = k - tap * CIN;
        const int ky = tap / KS, kx = tap - ky * KS;
        goff[grp] = (uint32_t)((ky * g.W + kx) * CIN + ci0 + c4 * 4);
        gokmask |= (k < KTOT ? 1u : 0u) << grp;
      }
    }
  }
  __device__ __forceinline__ void issue(int kb, uint32_t dst, int ptid) {
    while (next_kb < kb) {  // k-blocks are visited in order: advance this thread's pixel by 32 per k-block
      pix += tc::kBK; xx += tc::kBK;
      while (xx >= g.WO) { xx -= g.WO; if (++y == g.HO) { y = 0; ++n; } }
      ++next_kb;
    }
    const int c4 = ptid & 7, p = ptid >> 3;
    const bool pv = pix < pix_end;
    const uint32_t base = NCHW3 ? (uint32_t)((n * 3 * g.H + y * S) * g.W + xx * S) : (uint32_t)(((n * g.H + y * S) * g.W + xx * S) * CIN);
#pragma unroll
    for (int grp = 0; grp < 4; ++grp) {
      const bool ok = pv && ((gokmask >> grp) & 1u);
      cp_async16(dst + (uint32_t)grp * (tc::kBK * tc::kRowBytes) + swz32(p, c4), ok ? (const void*)(x + base + goff[grp]) : (const void*)x, ok);
    }
  }
};

// B operand (MN-major, COUT x 32 pixels) from dY stored [pixels][COUT]: thread (p, c4) owns chunk c4 of pixel p in each 32-channel group
template <int COUT>
struct WgradDyLoader {
  static constexpr bool kMNMajor = true;
  const float* dy;
  WgradTiling t;
  int tm, pix0, pix_end;
  const float* ptr;
  int pix;
  __device__ __forceinline__ void start_tile(int tile, int ptid) {
    t.decode(tile, tm, pix0, pix_end);
    pix = pix0 + (ptid >> 3);
    ptr = dy + (size_t)pix * COUT + (ptid & 7) * 4;
  }
  __device__ __forceinline__ void issue(int kb, uint32_t dst, int ptid) const {
    const bool ok = pix + kb * tc::kBK < pix_end;
    const float* p = ptr + (size_t)kb * tc::kBK * COUT;
#pragma unroll
    for (int grp = 0; grp < COUT / 32; ++grp)
      cp_async16(dst + (uint32_t)grp * (tc::kBK * tc::kRowBytes) + swz32(ptid >> 3, ptid & 7), ok ? (const void*)(p + grp * 32) : (const void*)dy, ok);
  }
};

struct WgradEpilogue {
  float* partial;  // [splits][KTOT][COUT]
  int splits, KTOT, COUT;
  __device__ __forceinline__ void operator()(int tile, int row, int col0, const float* v) const {
    const int tm = tile / splits, s = tile - tm * splits;
    const int k = tm * tc::kBM + row;
    if (k >= KTOT) return;
    float* dst = partial + ((size_t)s * KTOT + k) * COUT + col0;
#pragma unroll
    for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
  }
};

// dw[co][ci][ky][kx] = beta*dw + sum_s partial[s][k][co];  k = (ky, kx, ci) (NHWC order) or the natural order (nchw3)
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, int splits, int KTOT, int COUT, int CIN, int KS,
                                    int nchw3, float beta) {
  // 8 lanes per output element: lane j adds partials j, j + 8, ... (independent loads in flight), then a fixed-order shuffle tree
  const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, j = threadIdx.x & 7;
  const bool live = e < KTOT * COUT;
  const int k = live ? e / COUT : 0, co = live ? e - k * COUT : 0;
  float s = 0.f;
  if (live)
    for (int i = j; i < splits; i += 8) s += partial[((size_t)i * KTOT + k) * COUT + co];
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  if (!live || j != 0) return;
  int ko = k;
  if (!nchw3) {
    const int tap = k / CIN, ci = k - tap * CIN;
    ko = ci * KS * KS + tap;
  }
  float* d = dw + (size_t)co * KTOT + ko;
  *d = (beta != 0.f ? beta * *d : 0.f) + s;
}

// forward weights: wf[co][(ky,kx,ci)] = w[co][ci][ky][kx]
__global__ void prep_fwd_weights_kernel(const float* __restrict__ w, float* __restrict__ wf, int COUT, int CIN, int KS) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int K = CIN * KS * KS;
  if (i >= COUT * K) return;
  const int co = i / K, k = i - co * K;
  const int tap = k / CIN, ci = k - tap * CIN;
  wf[i] = w[(size_t)co * K + ci * KS * KS + tap];
}
// dgrad weights: wd[ph][ci][(jy,jx,co)] = w[co][ci][py+S*jy][px+S*jx]
__global__ void prep_dgrad_weights_kernel(const float* __restrict__ w, float* __restrict__ wd, int COUT, int CIN, int KS, int S) {
  const int R = KS / S, Kp = R * R * COUT;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S * S * CIN * Kp) return;
  const int k = i % Kp, ci = (i / Kp) % CIN, ph = i / (Kp * CIN);
  const int py = ph / S, px = ph - py * S;
  const int tap = k / COUT, co = k - tap * COUT;
  const int jy = tap / R, jx = tap - jy * R;
  wd[i] = w[(((size_t)co * CIN + ci) * KS + (py + S * jy)) * KS + (px + S * jx)];
}

// 32-wide k-blocks per pipeline stage (tc_pipeline.cuh).  Measured on the B200: the narrow tiles (BN = 32: the first layer and
// the data gradient of the second) gain from two k-blocks per barrier round; with BN = 64 the halved stage count costs more
// than the saved barrier traffic.  Every K of these layers is a multiple of 64.
template <int BN>
constexpr int kps() { return BN == 32 ? 2 : 1; }

template <int BN, class AL, class BL, class EP>
__global__ void __launch_bounds__(tc::PipeCfg<BN, false, tc::kBK, kps<BN>()>::kThreads, 1) conv_tc_kernel(AL al, BL bl, EP ep, int num_tiles, int num_kb) {
  tc::run_pipeline<BN, false, tc::kBK, 1, AL, BL, EP, kps<BN>()>(al, bl, ep, num_tiles, num_kb);
}

template <int BN, class AL, class BL, class EP>
int launch(AL al, BL bl, EP ep, int num_tiles, int num_kb, cudaStream_t st) {
  using Cfg = tc::PipeCfg<BN, false, tc::kBK, kps<BN>()>;
  if (num_kb % kps<BN>()) return (int)cudaErrorInvalidValue;
  auto kfn = conv_tc_kernel<BN, AL, BL, EP>;
  HULC_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
  HULC_LAUNCH(kfn, dim3(min(kNumSMs, num_tiles)), dim3(Cfg::kThreads), Cfg::kSmemBytes, st, al, bl, ep, num_tiles, num_kb);
  HULC_RETURN_LAST();
}

int conv_kind(int CIN, int COUT, int KS, int S) {
  if (CIN == 3 && COUT == 32 && KS == 8 && S == 4) return 1;
  if (CIN == 32 && COUT == 64 && KS == 4 && S == 2) return 2;
  if (CIN == 64 && COUT == 64 && KS == 3 && S == 1) return 3;
  return 0;
}

template <int CIN, int KS, int S, int COUT>
int fwd_nhwc(const Geom& g, const float* x, const float* w, const float* b, float* y, unsigned* bits, int relu, float* ws, cudaStream_t st) {
  const int K = CIN * KS * KS, M = g.N * g.HO * g.WO;
  HULC_LAUNCH(prep_fwd_weights_kernel, dim3(hulc_cdiv(COUT * K, 256)), dim3(256), 0, st, w, ws, COUT, CIN, KS);
  if (g_use_tma) {
    const int rc = hulc_conv_tma_fwd(x, ws, b, y, bits, g.N, CIN, g.H, g.W, COUT, KS, S, relu, st);
    if (rc != (int)cudaErrorNotSupported) return rc;
  }
  FwdNhwcLoader<CIN, KS, S> al{x, g, M, {0, 0, 0, 0}};
  WeightLoader<COUT> bl{ws, K};
  FwdEpilogue ep{y, b, M, COUT, relu, bits};
  return launch<COUT>(al, bl, ep, hulc_cdiv(M, tc::kBM), K / tc::kBK, st);
}

template <int CIN, int KS, int S, int COUT>
int dgrad_nhwc(const Geom& g, const float* dy, const float* w, const float* gate, const unsigned* gate_bits, float* dx, float* ws, cudaStream_t st) {
  constexpr int R = KS / S, Kp = R * R * COUT;
  HULC_LAUNCH(prep_dgrad_weights_kernel, dim3(hulc_cdiv(S * S * CIN * Kp, 256)), dim3(256), 0, st, w, ws, COUT, CIN, KS, S);
  if (g_use_tma && S == 2 && KS == 4) {  // all four stride phases in one launch
    const int rc = hulc_conv_tma_dgrad_s2_all(dy, ws, gate, gate_bits, dx, g.N, CIN, g.H, g.W, COUT, g.HO, g.WO, st);
    if (rc != (int)cudaErrorNotSupported) return rc;
  }
  for (int ph = 0; ph < S * S; ++ph) {
    const int py = ph / S, px = ph % S;
    const int HP = (g.H - py + S - 1) / S, WP = (g.W - px + S - 1) / S;
    const int M = g.N * HP * WP;
    if (M <= 0) continue;
    if (g_use_tma) {
      const int rc = hulc_conv_tma_dgrad_phase(dy, ws + (size_t)ph * CIN * Kp, gate, gate_bits, dx, g.N, CIN, g.H, g.W, COUT, g.HO, g.WO, R, S, py, px, st);
      if (rc == 0) continue;
      if (rc != (int)cudaErrorNotSupported) return rc;
    }
    DgradLoader<COUT, KS, S> al{dy, g, M, HP, WP, {0, 0, 0, 0}, {0, 0, 0, 0}};
    WeightLoader<CIN> bl{ws + (size_t)ph * CIN * Kp, Kp};
    DgradEpilogue<S> ep{dx, gate, g, M, HP, WP, py, px};
    HULC_TRY(launch<CIN>(al, bl, ep, hulc_cdiv(M, tc::kBM), Kp / tc::kBK, st));
  }
  return 0;
}

template <int CIN, int KS, int S, int COUT, bool NCHW3>
int wgrad(const Geom& g, const float* x, const float* dy, float* dw, float beta, float* ws, size_t ws_bytes, cudaStream_t st) {
  constexpr int KTOT = CIN * KS * KS;
  const int M = g.N * g.HO * g.WO;
  const int mt = hulc_cdiv(KTOT, tc::kBM);
  int splits = max(1, min((2 * kNumSMs) / mt, hulc_cdiv(M, 8 * tc::kBK)));
  while (splits > 1 && (size_t)splits * KTOT * COUT * sizeof(float) > ws_bytes) --splits;
  if ((size_t)splits * KTOT * COUT * sizeof(float) > ws_bytes) return (int)cudaErrorInvalidValue;
  const int pps = hulc_cdiv(hulc_cdiv(M, splits), 2 * tc::kBK) * (2 * tc::kBK);
  splits = hulc_cdiv(M, pps);
  WgradTiling t{splits, pps, M};
  WgradXLoader<CIN, KS, S, NCHW3> al{x, g, t, 0, 0, 0};
  WgradDyLoader<COUT> bl{dy, t, 0, 0, 0, nullptr, 0};
  WgradEpilogue ep{ws, splits, KTOT, COUT};
  HULC_TRY(launch<COUT>(al, bl, ep, mt * splits, pps / tc::kBK, st));
  HULC_LAUNCH(wgrad_reduce_kernel, dim3(hulc_cdiv(KTOT * COUT * 8, 256)), dim3(256), 0, st, (const float*)ws, dw, splits, KTOT, COUT, CIN, KS, NCHW3 ? 1 : 0, beta);
  HULC_RETURN_LAST();
}

}  // namespace

// Fixed-order reduction of the first layer's per-CTA partials [ctas][192][32] (+ [ctas][32] bias partials behind them when db != NULL)
int hulc_conv1_wgrad_reduce(const float* partial, float* dw, float* db, int ctas, float beta, cudaStream_t st) {
  HULC_LAUNCH(wgrad_reduce_kernel, dim3(hulc_cdiv(192 * 32 * 8, 256)), dim3(256), 0, st, partial, dw, ctas, 192, 32, 3, 8, 1, beta);
  if (db) return hulc_colsum(partial + (size_t)ctas * 192 * 32, ctas, 32, 32, db, 1.0f, nullptr, 0, (void*)st);  // the ones row of the same GEMM
  HULC_RETURN_LAST();
}

// Channels-last convolutions on the tensor cores.  x is NHWC [N,H,W,CIN] (or, for the 3-channel first layer, the
// reference's NCHW [N,3,H,W]); y / dy are NHWC [N,HO,WO,COUT]; w, dw keep the reference layout [COUT,CIN,KS,KS].
HULC_API int hulc_conv2d_tc_fwd(const float* x, const float* w, const float* b, float* y, int N, int CIN, int H, int W, int COUT, int KS, int S,
                                int relu, unsigned* relu_bits, float* workspace, size_t workspace_bytes, void* stream) {
  if (N <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (workspace_bytes < (kCounterFloats + (size_t)COUT * CIN * KS * KS) * sizeof(float)) return (int)cudaErrorInvalidValue;
  float* ws = workspace + kCounterFloats;
  Geom g{N, H, W, CIN, (H - KS) / S + 1, (W - KS) / S + 1, COUT};
  if ((long long)N * H * W * CIN >= (1ll << 32) / 4) return (int)cudaErrorInvalidValue;  // 32-bit element offsets
  switch (conv_kind(CIN, COUT, KS, S)) {
    case 1: {
      // every kernel of the first layer reads the NCHW rows in 16-byte pieces
      if ((W & 3) || (reinterpret_cast<size_t>(x) & 15)) return (int)cudaErrorInvalidValue;
      if (g_use_tma) {  // view / band-staged kernel (conv1_tc.cu) when the geometry fits; otherwise the im2col gather below
        const int rc = hulc_conv1_band_fwd(x, w, b, y, relu_bits, N, H, W, relu, st);
        if (rc != (int)cudaErrorNotSupported) return rc;
      }
      const int M = g.N * g.HO * g.WO;
      FwdNchw3Loader<8, 4> al{x, g, M, {0, 0, 0, 0}};
      WeightLoader<32> bl{w, 192};  // the reference layout [co][ci][ky][kx] already is K-major in (ci, ky, kx) order
      FwdEpilogue ep{y, b, M, 32, relu, relu_bits};
      return launch<32>(al, bl, ep, hulc_cdiv(M, tc::kBM), 192 / tc::kBK, st);
    }
    case 2: return fwd_nhwc<32, 4, 2, 64>(g, x, w, b, y, relu_bits, relu, ws, st);
    case 3: return fwd_nhwc<64, 3, 1, 64>(g, x, w, b, y, relu_bits, relu, ws, st);
  }
  return (int)cudaErrorInvalidValue;
}

// dx (NHWC) = conv_transpose(dy, w), masked by (gate > 0) when gate != NULL (gate: the NHWC activation that fed the conv)
HULC_API int hulc_conv2d_tc_dgrad(const float* dy, const float* w, const float* gate, const unsigned* gate_bits, float* dx, int N, int CIN, int H, int W, int COUT,
                                  int KS, int S, float* workspace, size_t workspace_bytes, void* stream) {
  if (N <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (workspace_bytes < (kCounterFloats + (size_t)COUT * CIN * KS * KS) * sizeof(float) || KS % S != 0) return (int)cudaErrorInvalidValue;
  float* ws = workspace + kCounterFloats;
  Geom g{N, H, W, CIN, (H - KS) / S + 1, (W - KS) / S + 1, COUT};
  if ((long long)N * g.HO * g.WO * COUT >= (1ll << 31)) return (int)cudaErrorInvalidValue;
  switch (conv_kind(CIN, COUT, KS, S)) {
    case 2: return dgrad_nhwc<32, 4, 2, 64>(g, dy, w, gate, gate_bits, dx, ws, st);
    case 3: return dgrad_nhwc<64, 3, 1, 64>(g, dy, w, gate, gate_bits, dx, ws, st);
  }
  return (int)cudaErrorInvalidValue;  // the first layer needs no data gradient: its input is the image
}

// dw = beta*dw + dL/dw.  x NHWC (x_nchw = 0) or the NCHW frames of the first layer (x_nchw = 1)
HULC_API int hulc_conv2d_tc_wgrad(const float* x, const float* dy, float* dw, float beta, float* db, int N, int CIN, int H, int W, int COUT, int KS, int S,
                                  int x_nchw, float* workspace, size_t workspace_bytes, void* stream) {
  if (N <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (workspace_bytes <= kCounterFloats * sizeof(float)) return (int)cudaErrorInvalidValue;
  float* ws = workspace + kCounterFloats;
  const size_t wsb = workspace_bytes - kCounterFloats * sizeof(float);
  Geom g{N, H, W, CIN, (H - KS) / S + 1, (W - KS) / S + 1, COUT};
  if ((long long)N * H * W * CIN >= (1ll << 32) / 4) return (int)cudaErrorInvalidValue;
  switch (conv_kind(CIN, COUT, KS, S)) {
    case 1: {
      if (!x_nchw || (W & 3) || (reinterpret_cast<size_t>(x) & 15)) return (int)cudaErrorInvalidValue;
      if (g_use_tma) {  // band-staged kernel (conv1_tc.cu): per-CTA partials, reduced in a fixed order below
        int ctas = 0;
        const int rc = hulc_conv1_band_wgrad_partials(x, dy, ws, wsb, db != nullptr, N, H, W, &ctas, st);
        if (rc == 0) return hulc_conv1_wgrad_reduce(ws, dw, db, ctas, beta, st);
        if (rc != (int)cudaErrorNotSupported) return rc;
      }
      HULC_TRY((wgrad<3, 8, 4, 32, true>(g, x, dy, dw, beta, ws, wsb, st)));
      break;
    }
    case 2:
      if (x_nchw) return (int)cudaErrorInvalidValue;
      HULC_TRY((wgrad<32, 4, 2, 64, false>(g, x, dy, dw, beta, ws, wsb, st)));
      break;
    case 3:
      if (x_nchw) return (int)cudaErrorInvalidValue;
      HULC_TRY((wgrad<64, 3, 1, 64, false>(g, x, dy, dw, beta, ws, wsb, st)));
      break;
    default: return (int)cudaErrorInvalidValue;
  }
  // bias gradient by a column-sum pass over dY where the weight-gradient kernel did not produce it
  if (db) return hulc_colsum(dy, N * g.HO * g.WO, COUT, COUT, db, 1.0f, workspace, workspace_bytes, stream);
  return 0;
}
