// cuda_emu.h — a minimal SIMT emulator so the CUDA kernel sources under hulc_b200/csrc compile with g++ and run on
// the host.  DEVELOPMENT / TEST INFRASTRUCTURE ONLY: this container has no GPU, and a `gpurun` round-trip costs
// minutes, so indexing / reduction / barrier-placement bugs in the SIMT kernels are shaken out here first
// (tests/test_emu_*.py).  The product never loads an emulated build: `hulc_b200/_lib.py` only opens the nvcc-built
// `libhulc_b200.so` and raises if it or a CUDA device is missing.  The tcgen05/TMA kernels are not emulated.
//
// Model: one OS worker thread per running block; the block's threads are ucontext fibers scheduled round-robin;
// `__syncthreads` / warp shuffles are generation barriers that yield.  `__shared__` becomes `static thread_local`.
#pragma once
#include <ucontext.h>

#include <algorithm>
#include <atomic>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#define HULC_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __restrict__ __restrict
#define __launch_bounds__(...)
#define __shared__ static thread_local
#define __constant__ static
#define __align__(n) __attribute__((aligned(n)))

struct uint3_ { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
struct int2 { int x, y; };
struct __attribute__((aligned(16))) int4 { int x, y, z, w; };
struct uint2 { unsigned x, y; };
struct __attribute__((aligned(4))) uchar4 { unsigned char x, y, z, w; };
struct __attribute__((aligned(16))) uint4 { unsigned x, y, z, w; };
static inline float2 make_float2(float x, float y) { return {x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
static inline int2 make_int2(int x, int y) { return {x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return {x, y, z, w}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return {x, y, z, w}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return {x, y}; }

typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorLaunchFailure = 719 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
enum cudaMemcpyKind { cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }

namespace emu {
constexpr int kMaxThreads = 1024;
constexpr size_t kStack = 96 * 1024;

struct Warp {
  int gen = 0, count = 0;
  uint64_t buf[2][32];
};
struct Block {
  ucontext_t main_ctx;
  ucontext_t ctx[kMaxThreads];
  char* stacks = nullptr;
  bool done[kMaxThreads];
  int nthreads = 0, exited = 0, cur = 0;
  int bar_gen = 0, bar_count = 0;
  int warp_exited[kMaxThreads / 32];
  Warp warps[kMaxThreads / 32];
  dim3 bdim, gdim, bidx;
  const std::function<void()>* body = nullptr;
  std::vector<char> dyn_smem;
};
extern thread_local Block* tb;            // the block this OS thread is running
extern thread_local uint3_ tl_threadIdx;  // refreshed on every fiber switch

static inline void yield() { Block* b = tb; swapcontext(&b->ctx[b->cur], &b->main_ctx); }

static inline void block_barrier() {
  Block* b = tb;
  int gen = b->bar_gen;
  if (++b->bar_count >= b->nthreads - b->exited) { b->bar_count = 0; b->bar_gen++; return; }
  while (b->bar_gen == gen) yield();
}
static inline int lane_id() { return tb->cur & 31; }
static inline Warp& my_warp() { return tb->warps[tb->cur >> 5]; }
static inline int warp_width() {
  Block* b = tb; int w = b->cur >> 5;
  return std::min(32, b->nthreads - w * 32) - b->warp_exited[w];
}
// arrive at the warp barrier; returns the generation index used for double buffering
static inline void warp_barrier() {
  Warp& w = my_warp();
  int gen = w.gen;
  if (++w.count >= warp_width()) { w.count = 0; w.gen++; return; }
  while (w.gen == gen) yield();
}
template <class T> static inline T shfl_idx(T v, int src) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  Warp& w = my_warp();
  int g = w.gen & 1;
  uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
  w.buf[g][lane_id()] = raw;
  warp_barrier();
  T r; memcpy(&r, &w.buf[g][src & 31], sizeof(T));
  return r;
}
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
static inline void* dyn_smem() { return tb->dyn_smem.data(); }
}  // namespace emu

#define threadIdx (emu::tl_threadIdx)
#define blockIdx (emu::tb->bidx)
#define blockDim (emu::tb->bdim)
#define gridDim (emu::tb->gdim)
static constexpr int warpSize = 32;

static inline void __syncthreads() { emu::block_barrier(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_barrier(); }
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __threadfence_block() {}
template <class T> static inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
  int lane = emu::lane_id();
  return emu::shfl_idx(v, (lane & ~(width - 1)) | (src & (width - 1)));
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m, int width = 32) { (void)width; return emu::shfl_idx(v, emu::lane_id() ^ m); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int width = 32) {
  int lane = emu::lane_id(); int src = lane + (int)d;
  if ((src & ~(width - 1)) != (lane & ~(width - 1))) src = lane;
  return emu::shfl_idx(v, src);
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int width = 32) {
  int lane = emu::lane_id(); int src = lane - (int)d;
  if (src < (lane & ~(width - 1))) src = lane;
  return emu::shfl_idx(v, src);
}
static inline unsigned __ballot_sync(unsigned, int pred) {
  unsigned bit = pred ? (1u << emu::lane_id()) : 0u, r = 0;
  for (int i = 0; i < 32; ++i) r |= emu::shfl_idx(bit, i);  // slow but simple
  return r;
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, !pred) == 0; }

static inline float atomicAdd(float* a, float v) {
  uint32_t* p = reinterpret_cast<uint32_t*>(a); uint32_t old = __atomic_load_n(p, __ATOMIC_RELAXED), nw; float f;
  do { memcpy(&f, &old, 4); f += v; memcpy(&nw, &f, 4); } while (!__atomic_compare_exchange_n(p, &old, nw, false, __ATOMIC_SEQ_CST, __ATOMIC_RELAXED));
  memcpy(&f, &old, 4); return f;
}
static inline int atomicAdd(int* a, int v) { return __atomic_fetch_add(a, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicAdd(unsigned* a, unsigned v) { return __atomic_fetch_add(a, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicAdd(unsigned long long* a, unsigned long long v) { return __atomic_fetch_add(a, v, __ATOMIC_SEQ_CST); }
static inline int atomicExch(int* a, int v) { return __atomic_exchange_n(a, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicExch(unsigned* a, unsigned v) { return __atomic_exchange_n(a, v, __ATOMIC_SEQ_CST); }
static inline int atomicMax(int* a, int v) { int old = *a; while (old < v && !__atomic_compare_exchange_n(a, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_RELAXED)) {} return old; }
static inline int atomicOr(int* a, int v) { return __atomic_fetch_or(a, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicInc(unsigned* a, unsigned lim) {
  unsigned old = __atomic_load_n(a, __ATOMIC_RELAXED), nw;
  do { nw = (old >= lim) ? 0 : old + 1; } while (!__atomic_compare_exchange_n(a, &old, nw, false, __ATOMIC_SEQ_CST, __ATOMIC_RELAXED));
  return old;
}

template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcg(const T* p) { return *p; }
template <class T> static inline T __ldcs(const T* p) { return *p; }
template <class T> static inline void __stcs(T* p, T v) { *p = v; }
#define __expf(x) expf(x)
#define __logf(x) logf(x)
#define __sinf(x) sinf(x)
#define __cosf(x) cosf(x)
#define __powf(x, y) powf(x, y)
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __saturatef(float x) { return x < 0 ? 0 : (x > 1 ? 1 : x); }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline float sinpif(float x) { return sinf(x * 3.14159265358979323846f); }
using std::max;
using std::min;

#define HULC_LAUNCH(kernel, grid, block, smem, stream, ...) \
  emu::launch((grid), (block), (smem), [=]() { kernel(__VA_ARGS__); })
#define HULC_DYN_SMEM(T, name) T* name = reinterpret_cast<T*>(emu::dyn_smem())
