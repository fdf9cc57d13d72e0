// Probe: how does a TMA box with a 32-byte inner dimension and SWIZZLE_128B land in shared memory?
#include <cstdio>
#include <vector>
#include "tma.cuh"
using namespace tc;

__global__ void probe(const __grid_constant__ CUtensorMap m, float* out, int box_floats, int c1, int c3) {
  extern __shared__ unsigned char raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) ((float*)smem)[i] = -1.f;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  fence_proxy_async();
  __syncthreads();
  if (threadIdx.x == 0) {
    tma::expect_tx(&bar, box_floats * 4);
    tma::load_5d(smem_u32(smem), &m, &bar, 0, c1, 0, c3, 1);
    mbar_wait(&bar, 0);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) out[i] = ((float*)smem)[i];
}

int main() {
  const int N = 2, H = 40, W = 40, HO = 9, WO = 9, RT = 2;
  std::vector<float> h(N * 3 * H * W);
  for (int pl = 0; pl < N * 3; ++pl) for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) h[(pl * H + y) * W + x] = pl * 10000 + y * 100 + x;
  float *d, *o; cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, 8192 * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  CUtensorMap m;
  const uint64_t dims[5] = {8, 8, WO, HO, (uint64_t)N * 3};
  const uint64_t strides[4] = {(uint64_t)W * 4, 16, (uint64_t)W * 16, (uint64_t)H * W * 4};
  const uint32_t box[5] = {8, 4, WO, RT, 1};
  for (int sw = 0; sw < 2; ++sw) {
    int rc = tma::make_map(&m, d, 5, dims, strides, box, sw ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE);
    printf("make_map rc=%d (swizzle %d)\n", rc, sw);
    probe<<<1, 128, 40 * 1024>>>(m, o, 8 * 4 * WO * RT, 4, 3);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    std::vector<float> r(8192);
    cudaMemcpy(r.data(), o, 8192 * 4, cudaMemcpyDeviceToHost);
    // expected for pixel (y_l, x): element [ky_l][kx] = plane 1, row 4*(3+y_l) + 4 + ky_l, col 4x + kx
    for (int row = 0; row < 20; ++row) {
      printf("row %2d:", row);
      for (int c = 0; c < 32; c += 4) printf(" %6.0f", r[row * 32 + c]);
      printf("\n");
    }
  }
  return 0;
}
