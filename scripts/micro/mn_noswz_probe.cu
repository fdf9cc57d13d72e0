// Probe: what does an UN-SWIZZLED MN-major tf32 UMMA operand descriptor read?  Shared memory word i holds the value i;
// B = I (8 x 8, K-major un-swizzled, the layout conv1_view_fwd_kernel uses), so D[m][n] = A[m][k = n] = the word index the
// tensor core fetched for operand element (m, k).  Tried for several (LBO, SBO) pairs.
#include <cstdio>
#include <vector>
#include "tc_common.cuh"
using namespace tc;

__global__ void probe(float* out, int lbo, int sbo, int a_mn) {
  extern __shared__ unsigned char raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  float* a = reinterpret_cast<float*>(smem);                 // 8192 words
  unsigned char* b = smem + 8192 * 4;
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) a[i] = (float)(i & 2047) ;
  for (int i = threadIdx.x; i < 64; i += blockDim.x) {
    const int n = i / 8, k = i % 8;
    *reinterpret_cast<float*>(b + (k >> 2) * 128 + n * 16 + (k & 3) * 4) = (n == k) ? 1.f : 0.f;
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(&tbase, 32);
  fence_proxy_async();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (threadIdx.x == 0) {
    const uint64_t da = make_smem_desc(smem_u32(a), (uint32_t)lbo, (uint32_t)sbo, 0u);
    const uint64_t db = make_smem_desc(smem_u32(b), 128u, 256u, 0u);
    umma_tf32(tbase, da, db, make_idesc_tf32(128, 8, a_mn != 0, false), 0);
    umma_commit(&bar);
    mbar_wait(&bar, 0);
  }
  __syncthreads();
  tc_fence_after_sync();
  if (threadIdx.x < 128) {
    uint32_t v[32];
    tmem_ld32(tbase + ((uint32_t)((threadIdx.x >> 5) * 32) << 16), v);
    tmem_ld_wait();
    for (int j = 0; j < 8; ++j) out[threadIdx.x * 8 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after_sync(); tmem_dealloc(tbase, 32); }
}

int main() {
  float* d; cudaMalloc(&d, 128 * 8 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int cfg[][3] = {{512, 64, 1}, {64, 512, 1}, {128, 16, 1}, {16, 128, 1}, {16, 128, 0}};
  for (auto& c : cfg) {
    probe<<<1, 128, 64 * 1024>>>(d, c[0], c[1], c[2]);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> h(128 * 8);
    cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
    printf("%s LBO %d SBO %d: %s   (word index fetched for (m, k))\n", c[2] ? "MN-major" : "K-major ", c[0], c[1], cudaGetErrorString(e));
    const int ms[] = {0, 1, 2, 3, 4, 5, 7, 8, 9, 12, 16, 32, 33, 64, 127};
    for (int m : ms) {
      printf("  m=%3d:", m);
      for (int k = 0; k < 8; ++k) printf(" %5.0f", h[m * 8 + k]);
      printf("\n");
    }
  }
  return 0;
}
