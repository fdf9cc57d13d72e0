// common.cuh — shared device helpers for the hulc_b200 kernels (sm_100a).
#pragma once

#if defined(HULC_HOST_EMULATION)
// tests/emu/cuda_emu.h is force-included by the emulator build (development aid, never shipped).
#else
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#define HULC_LAUNCH(kernel, grid, block, smem, stream, ...) \
  (++g_hulc_launches, kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__))
#define HULC_DYN_SMEM(T, name)                              \
  extern __shared__ __align__(16) unsigned char _dyn_smem[]; \
  T* name = reinterpret_cast<T*>(_dyn_smem)
#endif

#include <cfloat>
#include <cmath>

// number of kernel launches issued through this library since load (read by hulc_launch_count; bench.py reports it)
extern unsigned long long g_hulc_launches;

#define HULC_API extern "C" __attribute__((visibility("default")))

// Every C-ABI entry point returns 0 on success or a cudaError_t value.
#define HULC_RETURN_LAST() return (int)cudaGetLastError()
#define HULC_TRY(expr)                  \
  do {                                  \
    int _e = (int)(expr);               \
    if (_e != 0) return _e;             \
  } while (0)

static inline int hulc_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

constexpr int kNumSMs = 148;  // B200

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum; `red` is shared scratch of >= 32 floats.  All threads get the result.
__device__ __forceinline__ float block_sum(float v, float* red) {
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : 0.f;
  r = warp_sum(r);
  return r;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : -FLT_MAX;
  r = warp_max(r);
  return r;
}

// ---- counter-based RNG (Philox4x32-10) ------------------------------------------------------------------------------
// Dropout keep-decisions and latent-plan uniforms are pure functions of (seed, stream id, element index), so the
// backward pass regenerates them instead of storing masks.
__device__ __forceinline__ void philox_round(unsigned& c0, unsigned& c1, unsigned& c2, unsigned& c3, unsigned k0, unsigned k1) {
  const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  unsigned hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
  unsigned hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
  unsigned n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
  c0 = n0; c1 = n1; c2 = n2; c3 = n3;
}
__device__ __forceinline__ uint4 philox4x32(unsigned long long seed, unsigned stream, unsigned long long ctr) {
  unsigned c0 = (unsigned)ctr, c1 = (unsigned)(ctr >> 32), c2 = stream, c3 = 0x9E3779B9u;
  unsigned k0 = (unsigned)seed, k1 = (unsigned)(seed >> 32);
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    philox_round(c0, c1, c2, c3, k0, k1);
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
// uniform in [0,1) for element `idx` of random stream `stream`
__device__ __forceinline__ float philox_uniform(unsigned long long seed, unsigned stream, unsigned long long idx) {
  uint4 r = philox4x32(seed, stream, idx >> 2);
  unsigned w = (idx & 3) == 0 ? r.x : (idx & 3) == 1 ? r.y : (idx & 3) == 2 ? r.z : r.w;
  return (float)(w >> 8) * (1.0f / 16777216.0f);
}

// Dropout description passed by value to kernels.  p == 0 disables it.  When `keep` is non-null it is an injected
// keep-mask (uint8, 1 = keep) indexed like the tensor it applies to; otherwise the decision is philox(seed, site, idx).
struct DropSpec {
  float p;
  float scale;  // 1/(1-p)
  unsigned long long seed;
  const unsigned long long* seed_ptr;  // optional device-resident offset added to `seed` (see hulc_set_rng_offset_ptr)
  unsigned site;
  const unsigned char* keep;
};
// effective Philox seed: the by-value seed plus the device-resident offset.  Keeping the per-step part of the seed in device
// memory lets a captured CUDA graph of the whole training step draw fresh dropout masks / plan samples on every replay.
__device__ __forceinline__ unsigned long long rng_seed(unsigned long long seed, const unsigned long long* seed_ptr) {
  return seed + (seed_ptr ? *seed_ptr : 0ull);
}
__device__ __forceinline__ float drop_factor(const DropSpec& d, unsigned long long idx) {
  if (d.p <= 0.f) return 1.f;
  bool k = d.keep ? (d.keep[idx] != 0) : (philox_uniform(rng_seed(d.seed, d.seed_ptr), d.site, idx) >= d.p);
  return k ? d.scale : 0.f;
}
extern const unsigned long long* g_hulc_rng_offset_ptr;  // device pointer or null; set by hulc_set_rng_offset_ptr
static inline DropSpec make_drop(float p, unsigned long long seed, unsigned site, const unsigned char* keep) {
  DropSpec d;
  d.p = p; d.scale = p > 0.f ? 1.0f / (1.0f - p) : 1.0f; d.seed = seed; d.seed_ptr = g_hulc_rng_offset_ptr; d.site = site; d.keep = keep;
  return d;
}
