// Probe: can a K-major SWIZZLE_128B UMMA operand start at a row that is not a multiple of 8 (start address not 1024-byte
// aligned)?  A [256 rows x 32 fp32] is laid out with the swizzle keyed on the absolute row; B = I (32 x 32); D[m][n] should
// equal A[m + shift][n].  Tried with the descriptor's base-offset field (bits 49-51) = 0 and = (start >> 7) & 7.
#include <cstdio>
#include <vector>
#include "tc_common.cuh"
using namespace tc;

__global__ void probe(float* out, int shift, int use_base_offset) {
  extern __shared__ unsigned char raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  unsigned char* a = smem;                // 256 rows x 128 B
  unsigned char* b = smem + 256 * 128;    // 32 rows x 128 B
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  for (int i = threadIdx.x; i < 256 * 32; i += blockDim.x) {
    const int r = i / 32, c = i % 32;
    *reinterpret_cast<float*>(a + swz(r, c >> 2) + (c & 3) * 4) = (float)(r * 100 + c);
  }
  for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) {
    const int r = i / 32, c = i % 32;
    *reinterpret_cast<float*>(b + swz(r, c >> 2) + (c & 3) * 4) = (r == c) ? 1.f : 0.f;
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(&tbase, 32);
  fence_proxy_async();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (threadIdx.x == 0) {
    const uint32_t start = smem_u32(a) + shift * 128;
    for (int k = 0; k < 4; ++k) {
      uint64_t da = make_desc<false, 128, 32>(start, k);
      if (use_base_offset) da |= (uint64_t)((start >> 7) & 7) << 49;
      const uint64_t db = make_desc<false, 32, 32>(smem_u32(b), k);
      umma_tf32(tbase, da, db, make_idesc_tf32(128, 32), k != 0);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
  }
  __syncthreads();
  tc_fence_after_sync();
  if (threadIdx.x < 128) {
    uint32_t v[32];
    tmem_ld32(tbase + ((uint32_t)((threadIdx.x >> 5) * 32) << 16), v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[threadIdx.x * 32 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after_sync(); tmem_dealloc(tbase, 32); }
}

int main() {
  float* d; cudaMalloc(&d, 128 * 32 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int shifts[] = {0, 8, 1, 3, 23, 25, 48};
  for (int ubo = 0; ubo < 2; ++ubo)
    for (int s : shifts) {
      probe<<<1, 128, 64 * 1024>>>(d, s, ubo);
      cudaError_t e = cudaDeviceSynchronize();
      std::vector<float> h(128 * 32);
      cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
      int bad = 0, first = -1;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 32; ++n)
          if (h[m * 32 + n] != (float)((m + s) * 100 + n)) { if (first < 0) first = m * 32 + n; ++bad; }
      printf("base_offset %s shift %2d: %s, mismatches %d", ubo ? "set " : "zero", s, cudaGetErrorString(e), bad);
      if (bad) printf("  (first at m=%d n=%d: got %.0f want %.0f; row0: %.0f %.0f %.0f %.0f | %.0f)", first / 32, first % 32, h[first], (float)((first / 32 + s) * 100 + first % 32), h[0], h[1], h[4], h[8], h[32]);
      printf("\n");
    }
  return 0;
}
