// conv_tc_wgrad_stg.cu -- tcgen05 weight gradient with TMA-STAGED operands and bf16x3 math (sm_100a), for the stride-1
// "same" convolutions with 1 or 9 taps (1x1 / pad 0 and 3x3 / pad 1: 40 of ResNet-50's 53 layers, all of VGG-16's).
//
// Replaces cudnnConvolutionBackwardFilter beta=1 (reference src/caffe/layers/cudnn_conv_layer.cu:95-99) and, semantically,
// the per-image weight_gpu_gemm loop (base_conv_layer.hpp:148-162):
//     dW[o][c][tap] += sum_{n,p} dY[n][o][p] * X[n][c][p + off(tap)]        (X read as zero outside the image)
// GEMM view: M = 128 output channels, N = CN input channels x T taps (columns ordered (c, tap) = dW's own memory order, so
// a tile row is one contiguous run of dW), K = pixels.  The gather kernel (conv_tc_wgrad.cu) builds the X operand with
// per-thread predicated 4-byte loads: ~2 900 cycles per 32-pixel K block against 1 536 of MMA (profiles/README.md), 50-110
// TFLOP/s on the 3x3 layers.  Here, per K block of 64 pixels of one image:
//   * TMA brings in dY as two [128 o][32 px] fp32 boxes (SWIZZLE_128B) and X as ONE [CN c][64 + 2*halo px] fp32 box; pixels
//     past the image end and channels past O / C read as zero, which is the padding both the K tail and the M / N tails need;
//   * 16 converter warps split dY into bf16 hi / lo and write it into TENSOR MEMORY as the A operand (tcgen05.st, lane =
//     output channel; the swizzled 16-byte chunks make the row-wise ld.shared.v4 conflict-free), and build the B operand in
//     shared memory: row (c, tap) of the K-major SWIZZLE_128B tile is the staged row of channel c shifted by the tap offset,
//     masked where the tap leaves the image (zero padding / row wrap; one bit per pixel and tap, computed once per K block),
//     split into bf16 hi / lo; lanes run along the pixels, so both the reads and the 4-byte writes are conflict-free;
//   * one converged warp issues tcgen05.mma.kind::f16: lo*hi + hi*lo + hi*hi per 16-pixel K step, N = 256 (+ 32 for the 288
//     columns of a 3x3 tile), accumulators in TMEM for the whole K range of the CTA;
//   * the reduction over pixels is split across CTAs in one wave; partial tiles meet through the fused, deterministic
//     split-K reduction of conv_tc_wgrad.cu (arrival counter + per-CTA slice, fixed split order).
// Three bf16 MMAs per MAC cost 1.5 TF32 MMAs (3xTF32: 3); dropped terms ~2^-17 per product.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>
#include <stdio.h>
#include "b2c_common.cuh"
#define B2C_MBAR_DEBUG 1     // this kernel records stuck mbarrier waits (tc_common.cuh)
#include "tc_common.cuh"

namespace b2c {
using namespace tc;

namespace wstg {

constexpr int NCW = 16;                       // converter warps (warps 0-3 also run the epilogue)
constexpr int W_TMA = NCW;
constexpr int W_MMA = NCW + 1;
constexpr int THREADS = (NCW + 2) * 32;       // 576
constexpr int KPX = 64;                       // pixels per K block
constexpr uint32_t DY_BYTES = 2u * 128u * 128u;   // two [128 rows][32 px] fp32 boxes

struct Params {
  int HW, H, W, C, O;
  int halo;                // pad*W + pad
  int bwx;                 // staged pixels per X channel row (multiple of 4)
  int bpi;                 // K blocks per image = ceil(HW / 64)
  long long nkb_total;     // N * bpi
  int kb_per_split, splits;
  int Kd;                  // C * T (row length of dW)
  float* out;              // splits == 1: dW (accumulated); else partials [splits][O*Kd] (overwritten)
  float* grad;             // dW (fused reduction target when splits > 1)
  unsigned int* counters;
};

template <int T, int CN>
struct Cfg {
  static constexpr int NROWS = T * CN;                                  // B rows = GEMM N
  static constexpr int N0 = NROWS > 256 ? 256 : NROWS;                  // first MMA
  static constexpr int N1 = NROWS - N0;                                 // second MMA (0 or 32)
  static constexpr uint32_t B_PLANE = (uint32_t)NROWS * 128u;           // [NROWS][64 bf16]
  static constexpr uint32_t B_STAGE = 2u * B_PLANE;                     // hi + lo
  static constexpr int BSTAGES = 2;
  // raw operands as TMA delivers them.  dY and X have their own rings: the converters are done with a dY stage after the
  // first fifth of a K block (it goes to tensor memory), so ONE dY stage reloads in the shadow of the B-tile build, while X
  // is read until the end of the block and wants two.  3x3: X pool = 2 x [32 c][184 px] (W <= 56; wider maps, up to the
  // 256-column TMA box, run with one X stage).
  static constexpr int NDY = T == 1 ? 2 : 1;
  static constexpr uint32_t XMAX = (uint32_t)CN * (T == 1 ? 64u : 256u) * 4u;   // raw X box upper bound (1x1: no halo; else bwx <= 256)
  static constexpr uint32_t XPOOL = T == 1 ? 2u * XMAX : 2u * (uint32_t)CN * 184u * 4u;
  static_assert(XPOOL >= XMAX, "one X stage always fits");
  static constexpr uint32_t B_OFF = 0;
  static constexpr uint32_t DY_OFF = BSTAGES * B_STAGE;
  static constexpr uint32_t X_OFF = DY_OFF + NDY * DY_BYTES;
  static constexpr uint32_t BAR_OFF = X_OFF + XPOOL;
  static constexpr uint32_t TOTAL = BAR_OFF + 256 + 1024;
  static constexpr uint32_t D_COLS = NROWS;                             // fp32 accumulator columns
  static constexpr uint32_t A_COL0 = ((NROWS + 31) / 32) * 32;
  static_assert(A_COL0 + BSTAGES * 64 <= 512, "TMEM budget");
  static_assert(NROWS % 16 == 0 && N0 % 16 == 0 && (N1 == 0 || N1 % 16 == 0), "UMMA N granularity");
  static_assert(TOTAL <= 227u * 1024u, "shared memory budget");
};

__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
template <int OFF>
__device__ __forceinline__ float lds32i(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(OFF));
  return v;
}
__device__ __forceinline__ float2 lds64(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
// fp32 pair (k even, k odd) -> packed bf16 hi word and packed bf16 lo word (element k in the low half)
__device__ __forceinline__ void split_pack_bf16(float e, float o, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(e, o);
  const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h);
  const float he = __uint_as_float(hb << 16), ho = __uint_as_float(hb & 0xffff0000u);
  const __nv_bfloat162 l = __floats2bfloat162_rn(e - he, o - ho);
  hi = hb;
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void tmem_st8u(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}

// the deterministic split-K reduction shared with the other weight-gradient kernels (conv_tc_wgrad.cu has the twin)
__device__ __forceinline__ void splitk_fused_reduce(const float* __restrict__ part, float* __restrict__ dw, unsigned int* counter,
                                                    int split, int splits, long long plane, long long tile_base, int ld, int nrows, int ncols) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    const long long t0 = clock64();
    unsigned int seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
      if (seen < (unsigned)splits) {
        __nanosleep(64);
        if (clock64() - t0 > MBAR_TIMEOUT_CYCLES) asm volatile("trap;");
      }
    } while (seen < (unsigned)splits);
  }
  __syncthreads();
  const int total = nrows * ncols;
  const int per = (total + splits - 1) / splits;
  const int e_end = min(total, (split + 1) * per);
  for (int e = split * per + (int)threadIdx.x; e < e_end; e += (int)blockDim.x) {
    const int r = e / ncols, c = e - r * ncols;
    const long long idx = tile_base + (long long)r * ld + c;
    // loads of 8 splits in flight, added in split order (the order, not the batching, fixes the rounding)
    float acc = 0.0f;
    int s2 = 0;
    for (; s2 + 8 <= splits; s2 += 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldcg(part + (long long)(s2 + u) * plane + idx);
#pragma unroll
      for (int u = 0; u < 8; ++u) acc += v[u];
    }
    for (; s2 < splits; ++s2) acc += __ldcg(part + (long long)s2 * plane + idx);
    dw[idx] += acc;
  }
}

// PLANE (7x7 maps: a 196-byte channel pitch is not a TMA stride): one K block = one whole image.  The blobs are viewed as
// {4*H*W, channels/4, N} (four channels per tensor row), so a box of 32 rows is 128 channels x H*W pixels, contiguous; dY is read
// with 4-byte loads (odd word pitch: conflict-free), pixels H*W..63 of the 64-wide K block are written as zeros on both operands.
template <int T, int CN, bool PLANE>
__global__ void __launch_bounds__(THREADS, 1)
wgrad_stg_kernel(const __grid_constant__ Params p, const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x) {
  using S = Cfg<T, CN>;
  constexpr int NROWS = S::NROWS, NDY = S::NDY, BST = S::BSTAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sptr = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t bar_dy_full = sbase + S::BAR_OFF;         // 2: dY boxes landed
  const uint32_t bar_dy_empty = bar_dy_full + 16;          // 2: converters have read the dY stage
  const uint32_t bar_x_full = bar_dy_empty + 16;           // 2: X box landed
  const uint32_t bar_x_empty = bar_x_full + 16;            // 2: converters have read the X stage
  const uint32_t bar_b_full = bar_x_empty + 16;            // BST: A in TMEM + B tile written
  const uint32_t bar_b_empty = bar_b_full + 16;            // BST: MMAs of the stage completed
  const uint32_t bar_done = bar_b_empty + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sptr + S::BAR_OFF + 128);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int split = blockIdx.x;
  const int c0 = blockIdx.y * CN;                          // input-channel tile
  const int m0 = blockIdx.z * 128;                         // output-channel tile
  const long long kb_begin = (long long)split * p.kb_per_split;
  long long kb_end = kb_begin + p.kb_per_split;
  if (kb_end > p.nkb_total) kb_end = p.nkb_total;
  const int nkb = (int)(kb_end - kb_begin);                // >= 1 by construction
  const uint32_t x_bytes = (uint32_t)CN * (uint32_t)p.bwx * 4u;      // plane mode: bwx = H*W (the channel pitch of the box)
  const uint32_t dy_bytes = PLANE ? 128u * (uint32_t)p.HW * 4u : DY_BYTES;
  const int nx = 2u * x_bytes <= S::XPOOL ? 2 : 1;         // X stages

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_dy_full + 8 * s, 1); mbar_init(bar_dy_empty + 8 * s, NCW);
      mbar_init(bar_x_full + 8 * s, 1); mbar_init(bar_x_empty + 8 * s, NCW);
    }
    for (int s = 0; s < BST; ++s) { mbar_init(bar_b_full + 8 * s, NCW); mbar_init(bar_b_empty + 8 * s, 1); }
    mbar_init(bar_done, 1);
    fence_barrier_init();
  }
  if (warp == W_MMA) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto b_hi = [&](int s) { return sbase + S::B_OFF + (uint32_t)s * S::B_STAGE; };
  auto b_lo = [&](int s) { return sbase + S::B_OFF + (uint32_t)s * S::B_STAGE + S::B_PLANE; };
  auto raw_dy = [&](int r) { return sbase + S::DY_OFF + (uint32_t)r * DY_BYTES; };
  auto raw_x = [&](int r) { return sbase + S::X_OFF + (uint32_t)r * x_bytes; };
  auto a_col = [&](int s) { return S::A_COL0 + (uint32_t)s * 64u; };
  // first pixel of the staged X row for K block j of an image: 16-byte aligned, never negative
  auto x_start = [&](int j) { return max((j * KPX - p.halo) & ~3, 0); };

  if (warp < NCW) {
    // ================= converters ===============================================================================
    const int rq = warp & 3, cg = warp >> 2;               // dY: TMEM lane quarter / 16-pixel column group
    const int row = rq * 32 + lane;
    const uint32_t a_lane = (uint32_t)(rq * 32) << 16;
    int n_img = (int)(kb_begin / p.bpi), j = (int)(kb_begin - (long long)n_img * p.bpi);
    const uint32_t lq = (uint32_t)lane >> 2, wr_off = (uint32_t)(lane & 3) * 4u;   // 16-byte chunk / byte inside it of this lane's pixel pair
    for (int i = 0; i < nkb; ++i) {
      const int rd = i % NDY, rx = i % nx, s = i % BST;
      mbar_wait(bar_dy_full + 8 * rd, (i / NDY) & 1);
      // ---- (a) dY -> bf16 hi / lo -> tensor memory: this thread's row (output channel), pixels [cg*16, cg*16 + 16)
      uint32_t hi[8], lo[8];
      if constexpr (PLANE) {
        const uint32_t rowb = raw_dy(rd) + (uint32_t)row * (uint32_t)(p.HW * 4) + (uint32_t)(cg * 64);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float v[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) v[e] = cg * 16 + q * 4 + e < p.HW ? lds32(rowb + (uint32_t)((q * 4 + e) * 4)) : 0.f;
          split_pack_bf16(v[0], v[1], hi[2 * q], lo[2 * q]);
          split_pack_bf16(v[2], v[3], hi[2 * q + 1], lo[2 * q + 1]);
        }
      } else {
        const uint32_t box = raw_dy(rd) + (uint32_t)(cg >> 1) * (128u * 128u) + (uint32_t)row * 128u;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t chunk = (uint32_t)((cg & 1) * 4 + q) ^ (uint32_t)(row & 7);
          const float4 v = lds128(box + chunk * 16u);
          split_pack_bf16(v.x, v.y, hi[2 * q], lo[2 * q]);
          split_pack_bf16(v.z, v.w, hi[2 * q + 1], lo[2 * q + 1]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_dy_empty + 8 * rd);      // the dY stage reloads while the B tile is built
      mbar_wait_backoff(bar_b_empty + 8 * s, ((i / BST) & 1) ^ 1, 32);
      tc_fence_after();
      {
        const uint32_t a0 = tmem_base + a_lane + a_col(s) + (uint32_t)(cg * 8);
        tmem_st8u(a0, hi);
        tmem_st8u(a0 + 32, lo);
      }
      // ---- (b) X -> B tile rows (c, tap): lanes run along the pixels, two adjacent pixels per lane
      const int col0 = j * KPX - x_start(j) + 2 * lane;       // staged column of this lane's first pixel at tap offset 0
      const uint32_t bh = b_hi(s);
      if constexpr (T == 9) {
        // per-tap validity of this lane's two pixels as an AND mask over the packed pair (zero padding, image borders, row
        // wrap, pixels past the image end), once per K block
        uint32_t mw[9];
        {
          uint32_t rb[2], cb[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int pp = j * KPX + 2 * lane + e;
            const int h = pp / p.W, w = pp - h * p.W;
            rb[e] = pp < p.HW ? ((h > 0 ? 1u : 0u) | 2u | (h + 1 < p.H ? 4u : 0u)) : 0u;
            cb[e] = (w > 0 ? 1u : 0u) | 2u | (w + 1 < p.W ? 4u : 0u);
          }
#pragma unroll
          for (int t = 0; t < 9; ++t)
            mw[t] = (((rb[0] >> (t / 3)) & (cb[0] >> (t % 3)) & 1u) ? 0x0000ffffu : 0u) |
                    (((rb[1] >> (t / 3)) & (cb[1] >> (t % 3)) & 1u) ? 0xffff0000u : 0u);
        }
        mbar_wait(bar_x_full + 8 * rx, (i / nx) & 1);
        constexpr int CPW = CN / NCW;                           // channels per warp (all 9 taps of each)
        const int w4 = p.W * 4;
#pragma unroll
        for (int c = 0; c < CPW; ++c) {
          const int cl = warp * CPW + c;
          const uint32_t srcc = raw_x(rx) + (uint32_t)((cl * p.bwx + col0) * 4);
          const uint32_t n0 = (uint32_t)cl * 9u;
#pragma unroll
          for (int ti = 0; ti < 3; ++ti) {
            // the three taps of a filter row read a 4-pixel window [-1, +2] around the pair
            const uint32_t srow = srcc + (uint32_t)((ti - 1) * w4);
            float v[4];
            v[0] = lds32i<-4>(srow); v[1] = lds32i<0>(srow); v[2] = lds32i<4>(srow); v[3] = lds32i<8>(srow);
#pragma unroll
            for (int tj = 0; tj < 3; ++tj) {
              const int t = ti * 3 + tj;
              uint32_t hw, lw;
              split_pack_bf16(v[tj], v[tj + 1], hw, lw);
              hw &= mw[t]; lw &= mw[t];                         // exact zeros whatever the staged bytes were
              const uint32_t n = n0 + (uint32_t)t;
              const uint32_t dst = bh + n * 128u + ((lq ^ (n & 7u)) << 4) + wr_off;
              sts32(dst, hw);
              sts32(dst + S::B_PLANE, lw);
            }
          }
        }
      } else {
        mbar_wait(bar_x_full + 8 * rx, (i / nx) & 1);
        constexpr int RPW = CN / NCW;                           // channels (= B rows) per warp
        // 1x1: no halo, the box starts at the K block's first pixel: the pair is 8-byte aligned
        const uint32_t srcw = raw_x(rx) + (uint32_t)((warp * RPW * KPX + 2 * lane) * 4);
#pragma unroll
        for (int k = 0; k < RPW; ++k) {
          float2 v;
          if constexpr (PLANE) {
            // channel pitch H*W*4 bytes (4-byte aligned only); pixels past the image are the next channel's: zeros instead
            const uint32_t a = raw_x(rx) + (uint32_t)(((warp * RPW + k) * p.HW + 2 * lane) * 4);
            v.x = 2 * lane < p.HW ? lds32(a) : 0.f;
            v.y = 2 * lane + 1 < p.HW ? lds32(a + 4u) : 0.f;
          } else {
            v = lds64(srcw + (uint32_t)(k * KPX * 4));
          }
          uint32_t hw, lw;
          split_pack_bf16(v.x, v.y, hw, lw);
          const uint32_t n = (uint32_t)(warp * RPW + k);
          const uint32_t dst = bh + n * 128u + ((lq ^ (n & 7u)) << 4) + wr_off;
          sts32(dst, hw);
          sts32(dst + S::B_PLANE, lw);
        }
      }
      // X stage fully read by this warp: hand it back to the TMA warp
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_x_empty + 8 * rx);
      fence_proxy_async();                                  // st.shared B tile -> visible to the MMA's async-proxy reads
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_b_full + 8 * s);
      if (++j == p.bpi) { j = 0; ++n_img; }
    }
    if (warp < 4) {
      // ================= epilogue: TMEM -> registers -> smem transpose -> coalesced row stores of the partial tile ===
      mbar_wait_backoff(bar_done, 0, 128);
      tc_fence_after();
      float* tpad = reinterpret_cast<float*>(sptr) + warp * (32 * 33);     // all MMAs done: the B stages are free
      float* obase = p.out + (p.splits > 1 ? (long long)split * p.O * p.Kd : 0LL) + (long long)(m0 + warp * 32) * p.Kd + (long long)c0 * T;
      const int rows_valid = p.O - (m0 + warp * 32);
      const int cols_valid = min(NROWS, (p.C - c0) * T);
      const bool accumulate = p.splits == 1;
#pragma unroll 1
      for (int q0 = 0; q0 < NROWS; q0 += 32) {
        if (q0 >= cols_valid) break;
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)q0, v);
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) tpad[lane * 33 + jj] = v[jj];
        __syncwarp();
        const bool colok = q0 + lane < cols_valid;
#pragma unroll 4
        for (int rr = 0; rr < 32; ++rr) {
          if (rr < rows_valid && colok) {
            float* dst = obase + (long long)rr * p.Kd + q0 + lane;
            const float t = tpad[rr * 33 + lane];
            *dst = accumulate ? *dst + t : t;
          }
        }
        __syncwarp();
      }
      tc_fence_before();
    }
  } else if (warp == W_TMA) {
    // ================= TMA producer ================================================================================
    int n_img = (int)(kb_begin / p.bpi), j = (int)(kb_begin - (long long)n_img * p.bpi);
    for (int i = 0; i < nkb; ++i) {
      const int rd = i % NDY, rx = i % nx;
      mbar_wait_backoff(bar_x_empty + 8 * rx, ((i / nx) & 1) ^ 1, 32);
      if (elect_one()) {
        arrive_expect_tx(bar_x_full + 8 * rx, x_bytes);
        if constexpr (PLANE) tma_load_3d(raw_x(rx), &map_x, bar_x_full + 8 * rx, 0, c0 / 4, n_img);
        else tma_load_3d(raw_x(rx), &map_x, bar_x_full + 8 * rx, x_start(j), c0, n_img);
      }
      __syncwarp();
      mbar_wait_backoff(bar_dy_empty + 8 * rd, ((i / NDY) & 1) ^ 1, 32);
      if (elect_one()) {
        arrive_expect_tx(bar_dy_full + 8 * rd, dy_bytes);
        if constexpr (PLANE) {
          tma_load_3d(raw_dy(rd), &map_dy, bar_dy_full + 8 * rd, 0, m0 / 4, n_img);
        } else {
          tma_load_3d(raw_dy(rd), &map_dy, bar_dy_full + 8 * rd, j * KPX, m0, n_img);
          tma_load_3d(raw_dy(rd) + 128u * 128u, &map_dy, bar_dy_full + 8 * rd, j * KPX + 32, m0, n_img);
        }
      }
      __syncwarp();
      if (++j == p.bpi) { j = 0; ++n_img; }
    }
  } else {
    // ================= MMA issuer ===================================================================================
    constexpr uint32_t IDESC0 = idesc_bf16(128, S::N0);
    constexpr uint32_t IDESC1 = idesc_bf16(128, S::N1 > 0 ? S::N1 : 16);
    int j = (int)(kb_begin % p.bpi);
    for (int i = 0; i < nkb; ++i) {
      const int s = i % BST;
      mbar_wait(bar_b_full + 8 * s, (i / BST) & 1);
      tc_fence_after();
      const int valid = min(KPX, p.HW - j * KPX);
      const int ks = (valid + 15) >> 4;                      // K steps with any in-image pixel (the rest is all zero)
      if (elect_one()) {
        for (int kk = 0; kk < ks; ++kk) {
          const uint32_t ah = tmem_base + a_col(s) + (uint32_t)(kk * 8);
          const uint64_t bh = desc_sw128(b_hi(s) + kk * 32), bl = desc_sw128(b_lo(s) + kk * 32);
          const uint32_t acc = (i | kk) != 0;
          umma_bf16_ts(tmem_base, ah + 32, bh, IDESC0, acc);      // lo * hi
          umma_bf16_ts(tmem_base, ah, bl, IDESC0, 1);              // hi * lo
          umma_bf16_ts(tmem_base, ah, bh, IDESC0, 1);              // hi * hi
          if (S::N1 > 0) {
            const uint64_t bh1 = desc_sw128(b_hi(s) + 256u * 128u + kk * 32), bl1 = desc_sw128(b_lo(s) + 256u * 128u + kk * 32);
            umma_bf16_ts(tmem_base + 256, ah + 32, bh1, IDESC1, acc);
            umma_bf16_ts(tmem_base + 256, ah, bl1, IDESC1, 1);
            umma_bf16_ts(tmem_base + 256, ah, bh1, IDESC1, 1);
          }
        }
        umma_commit(bar_b_empty + 8 * s);
        if (i == nkb - 1) umma_commit(bar_done);
      }
      __syncwarp();
      if (++j == p.bpi) j = 0;
    }
  }
  if (p.splits > 1) {
    const int tile_id = blockIdx.z * gridDim.y + blockIdx.y;
    splitk_fused_reduce(p.out, p.grad, p.counters + tile_id, split, p.splits, (long long)p.O * p.Kd, (long long)m0 * p.Kd + (long long)c0 * T, p.Kd,
                        min(128, p.O - m0), min(NROWS, (p.C - c0) * T));
  }
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace wstg

// ---- host side ------------------------------------------------------------------------------------------------------
typedef CUresult (*WsEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static WsEncodeTiledFn ws_encode_tiled() {
  static WsEncodeTiledFn fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<WsEncodeTiledFn>(f);
  }
  return fn;
}
// [N][rows][HW] fp32, box {bw, box_rows, 1}
static int ws_make_map(CUtensorMap* map, const float* base, int HW, int rows, int N, int bw, int box_rows, bool swizzle128) {
  WsEncodeTiledFn enc = ws_encode_tiled();
  if (!enc) return fail(B2C_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[3] = {(cuuint64_t)HW, (cuuint64_t)rows, (cuuint64_t)N};
  cuuint64_t strides[2] = {(cuuint64_t)HW * 4, (cuuint64_t)HW * 4 * (cuuint64_t)rows};
  cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(B2C_ERR_CUDA, "cuTensorMapEncodeTiled (staged wgrad) failed (%d)", (int)r);
  return B2C_OK;
}

static bool wstg_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("B2C_WGRAD_STAGED"); on = e ? atoi(e) : 1; }
  return on != 0;
}
struct WstgPlan { int T, cn, bwx, bpi, splits, kb_per_split; long long nkb; bool plane; };
static bool wstg_plane_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("B2C_WGRAD_STAGED_PLANE"); on = e ? atoi(e) : 1; }
  return on != 0;
}
static bool wstg_plan(const ConvShape& s, WstgPlan* pl) {
  if (!wstg_enabled()) return false;
  if (s.G != 1 || s.sh != 1 || s.sw != 1 || s.dh != 1 || s.dw != 1 || s.Ho != s.H || s.Wo != s.W) return false;
  const bool k1 = s.kh == 1 && s.kw == 1 && s.ph == 0 && s.pw == 0;
  const bool k3 = s.kh == 3 && s.kw == 3 && s.ph == 1 && s.pw == 1;
  if (!k1 && !k3) return false;
  const long long HW = (long long)s.H * s.W;
  if ((long long)s.N * HW > 0x7fffffffLL - 256) return false;
  // 7x7 maps: plane mode (one image per K block, blobs viewed four channels per tensor row)
  const bool plane = HW == 49 && s.C % 4 == 0 && s.O % 4 == 0 && wstg_plane_enabled();
  if (!plane && (HW % 4 != 0 || HW < 64)) return false;
  const int T = k1 ? 1 : 9;
  const int halo = s.ph * s.W + s.pw;
  const int need = wstg::KPX + 2 * halo + ((4 - halo % 4) % 4);
  if (!plane && need > 256) return false;
  if (!pl) return true;
  pl->T = T;
  pl->plane = plane;
  pl->cn = k1 ? (s.C > 64 ? 128 : 64) : 32;
  pl->bwx = plane ? (int)HW : (need + 3) & ~3;
  pl->bpi = (int)((HW + wstg::KPX - 1) / wstg::KPX);
  pl->nkb = (long long)s.N * pl->bpi;
  const long long mn = (long long)((s.O + 127) / 128) * ((s.C + pl->cn - 1) / pl->cn);
  long long splits = sm_count() / mn;
  const long long max_splits = pl->nkb / 4 > 0 ? pl->nkb / 4 : 1;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  const long long per = (pl->nkb + splits - 1) / splits;
  pl->splits = (int)((pl->nkb + per - 1) / per);
  pl->kb_per_split = (int)per;
  return true;
}
constexpr size_t WSTG_COUNTER_BYTES = 1024;
bool tc_wgrad_stg_supported(const ConvShape& s) { return wstg_plan(s, nullptr); }
size_t tc_wgrad_stg_workspace(const ConvShape& s) {
  WstgPlan pl;
  if (!wstg_plan(s, &pl)) return 0;
  return pl.splits > 1 ? WSTG_COUNTER_BYTES + sizeof(float) * (size_t)pl.splits * s.O * s.Kd : 0;
}

template <int T, int CN, bool PLANE>
static int wstg_launch(const wstg::Params& p, const CUtensorMap& mdy, const CUtensorMap& mx, int c_tiles, int o_tiles, cudaStream_t st) {
  using S = wstg::Cfg<T, CN>;
  B2C_CUDA_OK(cudaFuncSetAttribute(wstg::wgrad_stg_kernel<T, CN, PLANE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::TOTAL));
  dim3 grid(p.splits, c_tiles, o_tiles);
  wstg::wgrad_stg_kernel<T, CN, PLANE><<<grid, wstg::THREADS, S::TOTAL, st>>>(p, mdy, mx);
  B2C_POST_LAUNCH();
  return B2C_OK;
}

static int launch_wstg_planned(const ConvShape& s, const WstgPlan& pl, const float* x, const float* dy, float* dw, void* ws, size_t ws_bytes,
                               cudaStream_t st);
int launch_conv_tc_wgrad_stg(const ConvShape& s, const float* x, const float* dy, float* dw, void* ws, size_t ws_bytes, cudaStream_t st) {
  WstgPlan pl;
  if (!wstg_plan(s, &pl)) return fail(B2C_ERR_INVALID, "staged wgrad: shape not eligible");
  return launch_wstg_planned(s, pl, x, dy, dw, ws, ws_bytes, st);
}

// ---- caffe_gpu_gemm, NoTrans x Trans (math_functions.cu:11-26 as InnerProductLayer::Forward_gpu calls it): C[M][N] (+)= A[M][K] * B[N][K]^T.
// Both operands have the reduction axis contiguous, which is exactly the 1x1 weight-gradient problem with one "image" of K
// "pixels": A plays dY (rows = M), B plays X (rows = N), C is dW with row length N.  No workspace in the BLAS signature, so
// the reduction is not split (one CTA per 128 x 128 output tile walks all of K).
bool tc_gemm_supported(bool tA, bool tB, int M, int N, int K) {
  return wstg_enabled() && !tA && tB && M > 0 && N > 0 && K >= 64 && K % 4 == 0 && (long long)M * N * 4 < (1ll << 40);
}
// split count of the K loop when the caller brings a workspace (b2c_sgemm_ex): one wave of CTAs, >= 4 K blocks per split
static void gemm_tc_plan(int M, int N, int K, WstgPlan* pl) {
  pl->T = 1; pl->plane = false; pl->cn = N > 64 ? 128 : 64; pl->bwx = wstg::KPX; pl->bpi = (K + wstg::KPX - 1) / wstg::KPX; pl->nkb = pl->bpi;
  const long long mn = (long long)((M + 127) / 128) * ((N + pl->cn - 1) / pl->cn);
  long long splits = sm_count() / mn;
  const long long max_splits = pl->nkb / 4 > 0 ? pl->nkb / 4 : 1;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  const long long per = (pl->nkb + splits - 1) / splits;
  pl->splits = (int)((pl->nkb + per - 1) / per);
  pl->kb_per_split = (int)per;
}
size_t tc_gemm_workspace(bool tA, bool tB, int M, int N, int K) {
  if (!tc_gemm_supported(tA, tB, M, N, K)) return 0;
  WstgPlan pl;
  gemm_tc_plan(M, N, K, &pl);
  return pl.splits > 1 ? WSTG_COUNTER_BYTES + sizeof(float) * (size_t)pl.splits * M * N : 0;
}
int launch_sgemm_tc(bool tA, bool tB, int M, int N, int K, float alpha, const float* A, const float* B, float beta, float* Cm, int math,
                    void* ws, size_t ws_bytes, cudaStream_t st) {
  (void)math;
  if (!tc_gemm_supported(tA, tB, M, N, K) || alpha != 1.0f || (beta != 0.0f && beta != 1.0f) ||
      ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15u))
    return fail(B2C_ERR_INVALID, "tcgen05 GEMM: only C (+)= A * B^T with alpha = 1, beta in {0, 1}, K %% 4 == 0 and 16-byte aligned operands");
  if (beta == 0.0f) B2C_CUDA_OK(cudaMemsetAsync(Cm, 0, sizeof(float) * (size_t)M * N, st));
  ConvShape s{};
  s.N = 1; s.C = N; s.H = 1; s.W = K; s.O = M; s.G = 1; s.kh = s.kw = 1; s.sh = s.sw = 1; s.ph = s.pw = 0; s.dh = s.dw = 1; s.has_bias = 0;
  s.Ho = 1; s.Wo = K; s.Cg = N; s.Og = M; s.Kd = N; s.is_1x1 = true;
  WstgPlan pl;
  gemm_tc_plan(M, N, K, &pl);
  const size_t need = pl.splits > 1 ? WSTG_COUNTER_BYTES + sizeof(float) * (size_t)pl.splits * M * N : 0;
  if (!ws || ws_bytes < need || (reinterpret_cast<uintptr_t>(ws) & 15u)) { pl.splits = 1; pl.kb_per_split = pl.bpi; ws = nullptr; ws_bytes = 0; }   // no workspace: one CTA per tile walks all of K
  return launch_wstg_planned(s, pl, B, A, Cm, ws, ws_bytes, st);
}

static int launch_wstg_planned(const ConvShape& s, const WstgPlan& pl, const float* x, const float* dy, float* dw, void* ws, size_t ws_bytes,
                               cudaStream_t st) {
  const size_t need = pl.splits > 1 ? WSTG_COUNTER_BYTES + sizeof(float) * (size_t)pl.splits * s.O * s.Kd : 0;
  if (need && (!ws || ws_bytes < need)) return fail(B2C_ERR_WORKSPACE, "staged wgrad: workspace too small");
  wstg::Params p;
  p.HW = s.H * s.W; p.H = s.H; p.W = s.W; p.C = s.C; p.O = s.O;
  p.halo = s.ph * s.W + s.pw; p.bwx = pl.bwx; p.bpi = pl.bpi; p.nkb_total = pl.nkb;
  p.kb_per_split = pl.kb_per_split; p.splits = pl.splits; p.Kd = s.Kd;
  p.out = pl.splits > 1 ? reinterpret_cast<float*>(static_cast<char*>(ws) + WSTG_COUNTER_BYTES) : dw;
  p.grad = dw;
  p.counters = static_cast<unsigned int*>(ws);
  if (pl.splits > 1) B2C_CUDA_OK(cudaMemsetAsync(ws, 0, WSTG_COUNTER_BYTES, st));
  alignas(64) CUtensorMap mdy, mx;
  const int c_tiles = (s.C + pl.cn - 1) / pl.cn, o_tiles = (s.O + 127) / 128;
  if (pl.plane) {
    // {4*HW, channels/4, N}: four channels per tensor row (a 16-byte-multiple pitch); box rows = tile channels / 4
    if (int rc = ws_make_map(&mdy, dy, p.HW * 4, s.O / 4, s.N, p.HW * 4, 32, false)) return rc;
    if (int rc = ws_make_map(&mx, x, p.HW * 4, s.C / 4, s.N, p.HW * 4, pl.cn / 4, false)) return rc;
    if (pl.T == 9) return wstg_launch<9, 32, true>(p, mdy, mx, c_tiles, o_tiles, st);
    if (pl.cn == 128) return wstg_launch<1, 128, true>(p, mdy, mx, c_tiles, o_tiles, st);
    return wstg_launch<1, 64, true>(p, mdy, mx, c_tiles, o_tiles, st);
  }
  if (int rc = ws_make_map(&mdy, dy, p.HW, s.O, s.N, 32, 128, true)) return rc;
  if (int rc = ws_make_map(&mx, x, p.HW, s.C, s.N, pl.bwx, pl.cn, false)) return rc;
  if (pl.T == 9) return wstg_launch<9, 32, false>(p, mdy, mx, c_tiles, o_tiles, st);
  if (pl.cn == 128) return wstg_launch<1, 128, false>(p, mdy, mx, c_tiles, o_tiles, st);
  return wstg_launch<1, 64, false>(p, mdy, mx, c_tiles, o_tiles, st);
}

TC_DEBUG_EXPORT(debug_mbar_wstg)

}  // namespace b2c
