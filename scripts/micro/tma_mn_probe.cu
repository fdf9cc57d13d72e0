// Probe: does a TMA load with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B produce the MN-major tf32 UMMA operand layout
// (SWIZZLE_128B_BASE32B, layout type 1)?  A^T is stored [K = 32 rows][M = 128 contiguous]; the tile is 4 groups of 32 M-elements,
// each loaded by one TMA box (32 floats x 32 k-rows).  B = I (32 x 32, K-major).  D[m][n] should equal A[m][n] = At[n][m].
#include <cstdio>
#include <vector>
#include "tma.cuh"
using namespace tc;

__global__ void probe(const __grid_constant__ CUtensorMap m, float* out, int manual) {
  extern __shared__ unsigned char raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  unsigned char* a = smem;              // 4 groups x 32 k-rows x 128 B = 16 KB
  unsigned char* b = smem + 16384;      // 32 rows x 128 B
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t tbase;
  for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) {
    const int r = i / 32, c = i % 32;
    *reinterpret_cast<float*>(b + swz(r, c >> 2) + (c & 3) * 4) = (r == c) ? 1.f : 0.f;
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(&tbase, 32);
  fence_proxy_async();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (threadIdx.x == 0) {
    tma::expect_tx(&bar2, 16384);
    for (int g = 0; g < 4; ++g) tma::load_2d(smem_u32(a) + g * 4096, &m, &bar2, g * 32, 0);
    mbar_wait(&bar2, 0);
    for (int k = 0; k < 4; ++k) umma_tf32(tbase, make_desc<true, 128, 32>(smem_u32(a), k), make_desc<false, 32, 32>(smem_u32(b), k), make_idesc_tf32(128, 32, true, false), k != 0);
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
  const int K = 32, M = 128;
  std::vector<float> h(K * M);
  for (int k = 0; k < K; ++k) for (int mm = 0; mm < M; ++mm) h[k * M + mm] = (float)(mm * 4 + k % 4 + (k / 4) * 0.125f);  // exactly representable in tf32
  float *d, *o; cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, 128 * 32 * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
  const CUtensorMapSwizzle modes[3] = {CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B_FLIP_8B};
  const char* names[3] = {"128B_ATOM_32B", "128B", "128B_ATOM_32B_FLIP_8B"};
  for (int t = 0; t < 3; ++t) {
    CUtensorMap m;
    const uint64_t dims[2] = {(uint64_t)M, (uint64_t)K};
    const uint64_t strides[1] = {(uint64_t)M * 4};
    const uint32_t box[2] = {32, 32};
    int rc = tma::make_map(&m, d, 2, dims, strides, box, modes[t]);
    probe<<<1, 128, 48 * 1024>>>(m, o, 0);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> r(128 * 32);
    cudaMemcpy(r.data(), o, r.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int mm = 0; mm < M; ++mm) for (int n = 0; n < 32; ++n) if (r[mm * 32 + n] != h[n * M + mm]) ++bad;
    printf("%-24s make_map %d, kernel %s, mismatches %d  (D[1][0..3] = %.3f %.3f %.3f %.3f want %.3f %.3f %.3f %.3f)\n", names[t], rc, cudaGetErrorString(e), bad,
           r[32], r[33], r[34], r[35], h[0 * M + 1], h[1 * M + 1], h[2 * M + 1], h[3 * M + 1]);
  }
  return 0;
}
