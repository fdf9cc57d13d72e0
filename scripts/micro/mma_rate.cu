// Micro-benchmark: cycles per tcgen05.mma (kind::tf32 / kind::f16) issued back to back by one thread into one accumulator,
// operands in shared memory (K-major, SWIZZLE_128B).   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I hulc_b200/csrc
#include <cstdio>
#include "tc_common.cuh"
using namespace tc;

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc),
               "r"(idesc), "r"(acc)
               : "memory");
}
// kind::f16 with bf16 inputs: a/b format 1 (BF16), D fp32
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

template <int M, int N, bool BF16, int DISTINCT, int MODE = 0>
__global__ void __launch_bounds__(128, 1) k(long long* out, int iters) {
  extern __shared__ unsigned char raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint64_t bars2[8];
  __shared__ uint32_t tbase;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) ((float*)smem)[i] = 0.f;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int i = 0; i < 8; ++i) mbar_init(&bars2[i], 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(&tbase, 512);
  fence_proxy_async();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (threadIdx.x == 0) {
    const uint32_t a = smem_u32(smem), b = a + 64 * 1024;
    const uint32_t idesc = BF16 ? idesc_bf16(M, N) : make_idesc_tf32(M, N);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      // DISTINCT k-blocks of 16 KB (A) / up to 32 KB (B) are cycled through, 4 MMAs (32 B of K each) per k-block
      const uint32_t at = a + (it % DISTINCT) * 16384, bt = b + (it % DISTINCT) * 32768 % (96 * 1024);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint64_t da = make_desc<false, 128, 32>(at, kk), db = make_desc<false, 128, 32>(bt, kk);
        if (BF16) umma_f16(tbase, da, db, idesc, 1); else umma_tf32(tbase, da, db, idesc, 1);
      }
      if (MODE >= 1) umma_commit(&bars2[it & 7]);
      if (MODE >= 2) tc_fence_after_sync();
      if (MODE == 4) { while (!mbar_test_wait(&bars2[(it + 4) & 7], it >= 4 ? (((it - 4) >> 3) & 1) : 1)) {} }
      else if (MODE == 5) { while (!mbar_test_wait(&bars2[(it + 7) & 7], it >= 1 ? (((it - 1) >> 3) & 1) : 1)) {} }
      else if (MODE >= 3) mbar_wait(&bars2[(it + 4) & 7], it >= 4 ? (((it - 4) >> 3) & 1) : 1);  // completion of the k-block issued 4 iterations ago
    }
    long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    out[0] = t1 - t0; out[1] = t2 - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after_sync(); tmem_dealloc(tbase, 512); }
}

template <int M, int N, bool BF16, int DISTINCT, int MODE = 0>
void run(const char* name) {
  long long* d; cudaMalloc(&d, 16);
  auto fn = k<M, N, BF16, DISTINCT, MODE>;
  cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 256;
  for (int rep = 0; rep < 2; ++rep) fn<<<1, 128, 200 * 1024>>>(d, iters);
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  double per = (double)h[1] / (iters * 4);
  double macs = (double)M * N * (BF16 ? 16 : 8);
  printf("%-28s issue %.1f cyc/mma, complete %.1f cyc/mma -> %.0f MAC/clk/SM (%s)\n", name, (double)h[0] / (iters * 4), per, macs / per, cudaGetErrorString(e));
  cudaFree(d);
}

// The warp-specialised protocol with null producers: P producer warps wait empty[stage], fence, arrive full[stage]; one thread
// waits full[stage], issues 4 MMAs, commits empty[stage].  VAR: 0 = try_wait everywhere, 1 = MMA thread spins on test_wait,
// 2 = no __syncwarp/whole-warp participation (only lane 0 of the MMA warp runs the loop)
template <int N, int STAGES, int VAR>
__global__ void __launch_bounds__(32 * 10, 1) proto(long long* out, int iters) {
  extern __shared__ unsigned char raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full[8], empty[8];
  __shared__ uint32_t tbase;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((float*)smem)[i] = 0.f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) { mbar_init(&full[i], VAR == 4 ? 1 : 8); mbar_init(&empty[i], 1); } fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&tbase, 512);
  fence_proxy_async();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (warp == 1) {
    const uint32_t a = smem_u32(smem), b = a + 64 * 1024;
    const uint32_t idesc = make_idesc_tf32(128, N);
    long long t0 = clock64();
    long long tw = 0, tf = 0, tm = 0, tc_ = 0, ts = 0;
    if (VAR != 2 || lane == 0) {
      for (int it = 0; it < iters; ++it) {
        const int st = it % STAGES; const uint32_t ph = (it / STAGES) & 1;
        long long c0 = clock64();
        if (VAR == 1) { while (!mbar_test_wait(&full[st], ph)) {} } else mbar_wait(&full[st], ph);
        long long c1 = clock64();
        tc_fence_after_sync();
        long long c2 = clock64(), c3 = c2, c4 = c2;
        if (lane == 0) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_tf32(tbase, make_desc<false, 128, 32>(a + st * 16384, kk), make_desc<false, 128, 32>(b + (st & 1) * 32768, kk), idesc, 1);
          c3 = clock64();
          umma_commit(&empty[st]);
          c4 = clock64();
        }
        if (VAR != 2) __syncwarp();
        long long c5 = clock64();
        tw += c1 - c0; tf += c2 - c1; tm += c3 - c2; tc_ += c4 - c3; ts += c5 - c4;
      }
    }
    long long t1 = clock64();
    if (lane == 0) { out[0] = t1 - t0; out[1] = t1 - t0; out[2] = tw; out[3] = tf; out[4] = tm; out[5] = tc_; out[6] = ts; }
  } else if (warp >= 2) {
    for (int it = 0; it < iters; ++it) {
      const int st = it % STAGES; const uint32_t ph = (it / STAGES) & 1;
      if (VAR == 4 && st != warp - 2) continue;  // one producer warp per stage
      mbar_wait(&empty[st], ph ^ 1);
      if (VAR != 3) fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[st]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) { tc_fence_after_sync(); tmem_dealloc(tbase, 512); }
}
template <int N, int STAGES, int VAR>
void runp(const char* name) {
  long long* d; cudaMalloc(&d, 64);
  auto fn = proto<N, STAGES, VAR>;
  cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 512;
  for (int rep = 0; rep < 2; ++rep) fn<<<1, 320, 200 * 1024>>>(d, iters);
  long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%-40s %.0f cyc per k-block (4 MMAs; ideal %d): wait %.0f fence %.0f mma-issue %.0f commit %.0f syncwarp %.0f (%s)\n", name, (double)h[0] / iters,
         N <= 128 ? 256 : 512, (double)h[2] / iters, (double)h[3] / iters, (double)h[4] / iters, (double)h[5] / iters, (double)h[6] / iters, cudaGetErrorString(e));
  cudaFree(d);
}

// Two MMA-issuing threads (warps 1 and 2) take alternate k-blocks of the same stage ring into separate accumulators: do their
// per-k-block overheads (barrier wait, commit) overlap?  Producers are null (8 warps arriving on full[] as empty[] allows).
template <int N, int STAGES>
__global__ void __launch_bounds__(32 * 11, 1) dual(long long* out, int iters) {
  extern __shared__ unsigned char raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full[8], empty[8];
  __shared__ uint32_t tbase;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((float*)smem)[i] = 0.f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) { mbar_init(&full[i], 8); mbar_init(&empty[i], 1); } fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&tbase, 512);
  fence_proxy_async();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (warp == 1 || warp == 2) {
    const int me = warp - 1;
    const uint32_t a = smem_u32(smem), b = a + 64 * 1024;
    const uint32_t idesc = make_idesc_tf32(128, N);
    const uint64_t a0 = make_desc<false, 128, 32>(a, 0), b0 = make_desc<false, 128, 32>(b, 0);
    long long t0 = clock64();
    if (lane == 0) {
      for (int it = me; it < iters; it += 2) {
        const int st = it % STAGES; const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&full[st], ph);
        tc_fence_after_sync();
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_tf32(tbase + me * 256, a0 + (uint32_t)(st * 1024 + kk * 2), b0 + (uint32_t)((st & 1) * 2048 + kk * 2), idesc, 1);
        umma_commit(&empty[st]);
      }
    }
    long long t1 = clock64();
    if (lane == 0 && me == 0) { out[0] = t1 - t0; }
  } else if (warp >= 3) {
    for (int it = 0; it < iters; ++it) {
      const int st = it % STAGES; const uint32_t ph = (it / STAGES) & 1;
      mbar_wait(&empty[st], ph ^ 1);
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[st]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) { tc_fence_after_sync(); tmem_dealloc(tbase, 512); }
}
template <int N, int STAGES>
void rund(const char* name) {
  long long* d; cudaMalloc(&d, 64);
  auto fn = dual<N, STAGES>;
  cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 512;
  for (int rep = 0; rep < 2; ++rep) fn<<<1, 352, 200 * 1024>>>(d, iters);
  long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%-40s %.0f cyc per k-block (%s)\n", name, (double)h[0] / iters, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  rund<64, 8>("dual issue N64 8 stages");
  rund<64, 4>("dual issue N64 4 stages");
  rund<128, 4>("dual issue N128 4 stages");
  runp<64, 8, 0>("protocol N64 8 stages try_wait");
  runp<64, 4, 0>("protocol N64 4 stages try_wait");
  runp<64, 8, 1>("protocol N64 8 stages test_wait spin");
  runp<64, 8, 2>("protocol N64 8 stages lane0 only");
  runp<64, 8, 3>("protocol N64 8 stages, no proxy fence");
  runp<64, 8, 4>("protocol N64 8 stages, warp per stage");
  runp<256, 8, 4>("protocol N256 8 stages, warp per stage");
  runp<128, 4, 0>("protocol N128 4 stages try_wait");
  runp<256, 4, 0>("protocol N256 4 stages try_wait");
  run<128, 64, false, 4>("tf32 M128 N64");
  run<64, 64, false, 4>("tf32 M64 N64");
  run<128, 128, false, 3>("tf32 M128 N128");
  run<128, 256, false, 3>("tf32 M128 N256");
  run<128, 64, true, 4>("bf16 M128 N64");
  run<128, 128, true, 3>("bf16 M128 N128");
  run<128, 256, true, 3>("bf16 M128 N256");
  run<128, 64, false, 1>("tf32 M128 N64 same tile");
  run<128, 64, false, 4, 1>("tf32 M128 N64 +commit/kb");
  run<128, 64, false, 4, 2>("tf32 M128 N64 +commit+fence");
  run<128, 64, false, 4, 3>("tf32 M128 N64 +commit+fence+wait(-4)");
  run<128, 64, false, 4, 4>("tf32 M128 N64 +commit+test_wait(-4)");
  run<128, 64, false, 4, 5>("tf32 M128 N64 +commit+test_wait(-1)");
  run<128, 128, false, 3, 3>("tf32 M128 N128 +commit+fence+wait(-4)");
  return 0;
}
