// tc_common.cuh — sm_100a tensor-core plumbing shared by the tcgen05 kernels (gemm_tc.cu, conv_tc.cu): mbarriers,
// proxy fences, TMEM allocation, UMMA shared-memory / instruction descriptors, tcgen05.mma / commit / ld wrappers and
// the K-major SWIZZLE_128B tile layout the producer warps write.
//
// Operand tiles live in shared memory as [rows][32 fp32] = 128-byte rows, K-major, 128-byte swizzle: within every
// 1024-byte group of 8 rows the 16-byte chunk c of row r sits at chunk position c ^ (r & 7).  One tcgen05.mma of
// kind::tf32 consumes K = 8 (32 bytes of every row), so a 32-wide k-block is 4 MMAs whose descriptors differ by +32 B.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace tc {

constexpr int kBK = 32;              // fp32 elements per k-block (= one 128-byte swizzled row)
constexpr int kRowBytes = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier -------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Spin with a watchdog: a protocol bug must surface as a launch failure (trap), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins) {
    if (spins > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// generic-proxy writes (st.shared) -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMEM -----------------------------------------------------------------------------------------------------------
// Called by ONE full warp; ncols is a power of two >= 32.  The base address lands in *dst (shared memory).
__device__ __forceinline__ void tmem_alloc(uint32_t* dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
// 32 lanes x 32 columns of fp32: thread (lane) l of the warp receives columns [col, col+32) of TMEM lane (lane_base + l).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 256-bit global store (sm_100: STG.E.ENL2.256): a thread that owns a row writes whole 32-byte sectors, half the store instructions of
// float4 and no half-written sectors when the lanes' rows are far apart.  p must be 32-byte aligned.
__device__ __forceinline__ void st_global_v8(float* p, float a0, float a1, float a2, float a3, float a4, float a5, float a6, float a7) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(a0), "f"(a1), "f"(a2), "f"(a3), "f"(a4), "f"(a5), "f"(a6), "f"(a7)
               : "memory");
}

// 256-bit global load (LDG.E.ENL2.256); p must be 32-byte aligned
__device__ __forceinline__ void ld_global_v8(const float* p, float* o) {
  asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(o[0]), "=f"(o[1]), "=f"(o[2]), "=f"(o[3]), "=f"(o[4]), "=f"(o[5]), "=f"(o[6]), "=f"(o[7])
               : "l"(p));
}

// ---- UMMA descriptors -----------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor), SWIZZLE_128B:
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 | [32,46) stride byte offset >> 4 | [46,48) version = 1 |
//   [61,64) layout type: 2 = SWIZZLE_128B (K-major tiles), 1 = SWIZZLE_128B_BASE32B (the only MN-major layout tf32 has).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}
// Descriptor of the operand slice one tcgen05.mma (K = 8 tf32) reads from a [ROWS x BK] stage tile.
//   K-major tile  [ROWS][BK fp32 = 128 B]: 8-row groups are 1024 B apart (SBO); the slice starts 32*kstep bytes into the row.
//   MN-major tile [ROWS/32][BK k-rows][32 fp32 = 128 B], swizzle atom = 4 k-rows (512 B, 32-byte chunk index ^ k-row & 3):
//                 the slice is the 8 k-rows starting at k-row 8*kstep = two atoms 512 B apart (SBO); the 32-element
//                 groups along M/N are BK*128 B apart (LBO).
template <bool MN_MAJOR, int ROWS, int BK>
__device__ __forceinline__ uint64_t make_desc(uint32_t tile, int kstep) {
  if (MN_MAJOR) return make_smem_desc(tile + (uint32_t)kstep * 1024u, (uint32_t)BK * kRowBytes, 512u, 1u);
  return make_smem_desc(tile + (uint32_t)kstep * 32u, 16u, 1024u, 2u);
}
// Instruction descriptor for kind::tf32, fp32 accumulate (cute::UMMA::InstrDescriptor):
//   [4,6) D format = 1 (F32) | [7,10) A format = 2 (TF32) | [10,13) B format = 2 | [15] A major (0 = K, 1 = MN) |
//   [16] B major | [17,23) N >> 3 | [24,29) M >> 4.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, bool a_mn = false, bool b_mn = false) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread on behalf of the CTA
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on `bar` once every tcgen05.mma issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- operand staging ------------------------------------------------------------------------------------------------
// round-to-nearest fp32 -> tf32 (10-bit mantissa); the tensor core would otherwise truncate
__device__ __forceinline__ float to_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}
// byte offset of 16-byte chunk c (0..7) of row r inside a swizzled tile
__device__ __forceinline__ uint32_t swz(int r, int c) { return (uint32_t)r * kRowBytes + (uint32_t)((c ^ (r & 7)) << 4); }

// MN-major tiles (SWIZZLE_128B_BASE32B): byte offset of 16-byte chunk c (0..7) of k-row kk inside a 32-element group
__device__ __forceinline__ uint32_t swz32(int kk, int c) {
  return (uint32_t)kk * kRowBytes + (uint32_t)((((c >> 1) ^ (kk & 3)) << 5) | ((c & 1) << 4));
}

__device__ __forceinline__ void st_chunk(unsigned char* tile, int r, int c, float4 v) {
  *reinterpret_cast<float4*>(tile + swz(r, c)) = v;
}
// 1 pass: hi only.  3 passes: hi = tf32(x), lo = tf32(x - hi): x*y ~= hi_x*hi_y + hi_x*lo_y + lo_x*hi_y to ~fp32 accuracy.
template <bool SPLIT>
__device__ __forceinline__ void st_chunk_split(unsigned char* hi, unsigned char* lo, int r, int c, float4 v) {
  float4 h = make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
  st_chunk(hi, r, c, h);
  if (SPLIT) {
    float4 l = make_float4(to_tf32(v.x - h.x), to_tf32(v.y - h.y), to_tf32(v.z - h.z), to_tf32(v.w - h.w));
    st_chunk(lo, r, c, l);
  }
}

// Optional device-side timeline (development aid, compiled in with -DHULC_TC_TRACE): globaltimer stamps of CTA 0.
#ifdef HULC_TC_TRACE
__device__ unsigned long long g_tc_trace[64];
__device__ __forceinline__ void trace(int slot) {
  if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_tc_trace[slot] = t;
  }
}
#else
__device__ __forceinline__ void trace(int) {}
#endif

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

}  // namespace tc
