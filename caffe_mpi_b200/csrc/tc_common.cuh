// tc_common.cuh -- Blackwell (sm_100a) building blocks for the tcgen05 kernels: mbarrier, proxy fences,
// TMEM allocation, UMMA shared-memory / instruction descriptors, tcgen05.mma / commit / ld wrappers,
// and the TF32 hi/lo split used by the fp32-equivalent (3xTF32) math mode.
//
// Shared-memory operand layout used by every kernel here: K-major, no swizzle ("interleave"),
// canonical form ((8,m),(4,2)):((4,SBO),(1,LBO)) in fp32 elements, i.e. 8-row x 16-byte core
// matrices of 128 contiguous bytes; consecutive 8-row groups SBO = 128 B apart, consecutive
// 16-byte K chunks LBO = rows*16 + 16 B apart.  The +16 pad makes the eight K chunks of one row
// land in eight different 16-byte bank groups, so both producer mappings (lanes along rows, and
// lanes along K) store conflict-free with 128-bit st.shared.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2c {
namespace tc {

constexpr int BK = 32;              // fp32 elements of K per pipeline stage (8 chunks of 16 B)
constexpr int KCHUNKS = BK / 4;

__host__ __device__ constexpr uint32_t tile_lbo(int rows) { return (uint32_t)rows * 16u + 16u; }
__host__ __device__ constexpr uint32_t tile_bytes(int rows) { return KCHUNKS * tile_lbo(rows); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Every wait is bounded: a protocol bug (a missed arrive, a wrong parity) must fail the launch with a trap
// ("unspecified launch failure" on the host) instead of hanging the GPU.  The bound is ~2^31 SM cycles (about a
// second), three orders of magnitude beyond the longest legitimate wait in any of these kernels.
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done != 0;
}
// Timeout handling (cold path).  Default: kill the kernel with `trap` (the host sees "illegal instruction").
// A translation unit that defines B2C_MBAR_DEBUG before including this header records the stuck waits instead:
// g_mbar_dbg[0] = number of timeouts, then up to 31 records of {block, thread, barrier smem address, parity}; with
// B2C_MBAR_TRAP=0 (b2c_debug_mbar_set_trap) the wait then gives up, so a deadlocked kernel drains and
// b2c_debug_mbar_timeouts() can say which role was stuck on which barrier.  (Off by default: the out-of-line call costs the
// register-tight gather kernels a few spilled registers.)
#ifdef B2C_MBAR_DEBUG
static __device__ unsigned int g_mbar_dbg[128];
static __device__ unsigned int g_mbar_trap = 1;
static __device__ __noinline__ void mbar_timeout_trap(uint32_t bar, uint32_t parity) {
  const unsigned int slot = atomicAdd(&g_mbar_dbg[0], 1u);
  if (slot < 31u) {
    unsigned int* r = &g_mbar_dbg[4 + 4 * slot];
    r[0] = blockIdx.x | (blockIdx.y << 16); r[1] = threadIdx.x; r[2] = bar; r[3] = parity;
  }
  __threadfence_system();
  if (g_mbar_trap) asm volatile("trap;");
}
#define TC_DEBUG_EXPORT(name)                                                                                   \
  int name(unsigned int* out, int cap_words, int set_trap) {                                                    \
    if (set_trap >= 0) { const unsigned int v = set_trap ? 1u : 0u; cudaMemcpyToSymbol(::b2c::tc::g_mbar_trap, &v, sizeof(v)); } \
    if (!out) return 0;                                                                                          \
    unsigned int h[128];                                                                                         \
    if (cudaMemcpyFromSymbol(h, ::b2c::tc::g_mbar_dbg, sizeof(h)) != cudaSuccess) return -1;                    \
    for (int i = 0; i < cap_words && i < 128; ++i) out[i] = h[i];                                                \
    unsigned int z[128] = {0};                                                                                   \
    cudaMemcpyToSymbol(::b2c::tc::g_mbar_dbg, z, sizeof(z));                                                     \
    return (int)h[0];                                                                                            \
  }
#else
__device__ __forceinline__ void mbar_timeout_trap(uint32_t, uint32_t) { asm volatile("trap;"); }
#define TC_DEBUG_EXPORT(name)                                                          \
  int name(unsigned int* out, int cap_words, int) {                                    \
    for (int i = 0; out && i < cap_words && i < 128; ++i) out[i] = 0;                  \
    return 0;                                                                          \
  }
#endif
constexpr long long MBAR_TIMEOUT_CYCLES = 1ll << 31;
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try(bar, parity)) {
    if ((++spins & 0x3ffu) == 0 && clock64() - t0 > MBAR_TIMEOUT_CYCLES) { mbar_timeout_trap(bar, parity); return; }
  }
}
// Waiters that are not on the critical path (epilogue warps waiting a whole tile for the accumulator,
// producers waiting for a free stage) must not spin hot: the warp scheduler prefers high warp ids, and a
// try_wait loop in 4-20 warps starves the single MMA-issuing thread of issue slots.  Back off with nanosleep.
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity, unsigned ns) {
  if (mbar_try(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  for (;;) {
    __nanosleep(ns);
    if (mbar_try(bar, parity)) break;
    if ((++spins & 0xffu) == 0 && clock64() - t0 > MBAR_TIMEOUT_CYCLES) { mbar_timeout_trap(bar, parity); return; }
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// generic-proxy st.shared -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// One lane of a CONVERGED warp.  tcgen05.mma / tcgen05.commit / TMA are uniform-datapath instructions: issued
// under `if (lane == 0)` the compiler wraps every one of them in an ELECT/branch loop over the active lanes
// (~80 cycles per MMA measured: the single issuing thread, not the tensor pipe, bounded the kernel).  With
// elect.sync on a converged warp they become straight-line uniform instructions.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t"
      "}" : "=r"(pred));
  return pred != 0;
}

// ---- tcgen05 ---------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// whole warp; writes the TMEM base address to *smem_dst
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// 64-bit shared-memory matrix descriptor, K-major, SWIZZLE_NONE, Blackwell version field = 1.
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// 32-bit instruction descriptor: D=F32, A=B=TF32, both K-major, M x N.
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T, 128 x N x 8 (tf32).  One thread issues.
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns (thread t gets lane base+t).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// registers -> TMEM: this warp's 32 lanes x 16 consecutive 32-bit columns (thread t writes lane base+t)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// registers -> TMEM, 8 columns
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem]^T: A operand read from tensor memory (rows = lanes, K along columns)
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// ---- TF32 split ------------------------------------------------------------------------------------
// fp32-equivalent mode: hi = a with the 13 low mantissa bits cleared (exactly a TF32 value), lo = a - hi
// (exact in fp32, |lo| < 2^-10 |a|; the tensor core reads its top 19 bits).  hi*hi' + hi*lo' + lo*hi'
// then carries ~2^-20 relative error per product -- fp32-class accuracy on the TF32 pipe for two ALU
// ops per element (cvt.rna.tf32 runs at 1/4 rate and made the producers conversion-bound).
// TF32 mode: round-to-nearest (ties away) on the integer pipe: (bits + 0x1000) & ~0x1fff.
__device__ __forceinline__ float to_tf32(float a) {
  return __uint_as_float((__float_as_uint(a) + 0x1000u) & 0xffffe000u);
}
__device__ __forceinline__ void split_tf32(float a, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(a) & 0xffffe000u);
  lo = a - hi;
}

__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// store a 4-element K chunk of one row into the hi (and lo) tile
template <bool SPLIT>
__device__ __forceinline__ void store_chunk(uint32_t hi_addr, uint32_t lo_addr, float v0, float v1, float v2, float v3) {
  if (SPLIT) {
    float h0, h1, h2, h3, l0, l1, l2, l3;
    split_tf32(v0, h0, l0); split_tf32(v1, h1, l1); split_tf32(v2, h2, l2); split_tf32(v3, h3, l3);
    sts128(hi_addr, h0, h1, h2, h3);
    sts128(lo_addr, l0, l1, l2, l3);
  } else {
    sts128(hi_addr, to_tf32(v0), to_tf32(v1), to_tf32(v2), to_tf32(v3));
  }
}

}  // namespace tc
}  // namespace b2c
