// conv_tc_stg.cu -- tcgen05 implicit-GEMM convolution with BULK-COPY-STAGED activations and bf16x3 math (sm_100a):
// forward and data-gradient of stride-1 "same" convolutions (1x1 / pad 0, 3x3 / pad 1, 5x5 / pad 2 ...), which are
// 43 of ResNet-50's 53 layers and every VGG-16 / GoogLeNet inception convolution with C % 32 == 0.
//
// Replaces cudnnConvolutionForward / cudnnConvolutionBackwardData as the reference's CuDNNConvolutionLayer calls them
// (src/caffe/layers/cudnn_conv_layer.cu:25-29,118-123); same contract as conv_tc.cu: NCHW fp32 in and out, whole
// batch per launch, y / dx overwritten, bias fused.
//
// Why a second kernel.  conv_tc.cu gathers the im2col rows with per-thread __ldg: ~1 400 cycles per 128x32 K block
// against 768 cycles of MMA (profiles/README.md), and it re-reads every input pixel kh*kw times through L1/L2.  Here
// no thread touches global memory for the activation operand:
//   * GEMM rows are 128 CONSECUTIVE flattened pixels q = n*H*W + p.  For 32 input channels the pixels a tile needs --
//     its own 128 plus a halo of pad*W + pad on either side -- are ONE TMA box per image the tile touches (at most two):
//     a 3-D tensor map {H*W, C, N} over the NCHW blob, box {BWT pixels, 32 channels, 1 image}, start column
//     p0 - halo; columns outside the image read as zero.  (A first version used one 1-D bulk copy per channel: the TMA
//     unit retires ~1 small bulk operation per 90 cycles, 3 000 cycles per 32-channel group -- 1x1 layers ran 2-3x
//     SLOWER than the gather kernel, profiles/r02_c2_*.)
//   * All kh*kw taps of those 32 channels are then served from that ONE staged copy: tap (i,j) of row r is the staged
//     value at column r + halo + (i-pad)*W + (j-pad); rows whose tap falls outside the image (zero padding, image
//     borders, row wrap) are masked with a per-row bit computed once per tile.  Global->smem traffic per tile is
//     (128 + 2*halo)/128 of the input instead of kh*kw times it.
//   * 16 converter warps read the staged fp32 values (lanes = consecutive pixels: conflict-free; the row pitch BWT is a
//     template parameter so the eight channel loads of a thread are one base register + immediates), split each into
//     bf16 hi + bf16 lo (hi = rn(a), lo = rn(a - hi)), pack pairs along K and write the A operand straight into TENSOR
//     MEMORY with tcgen05.st (lane = GEMM row), 64 K-elements per stage.
//   * The filter is pre-split into bf16 hi / lo in GEMM-K order (channel group, tap, channel) by a prepass and streamed
//     by TMA (SWIZZLE_128B, [N_TILE x 64] boxes); one converged warp issues tcgen05.mma.kind::f16 (bf16 x bf16 -> fp32):
//     lo*hi + hi*lo + hi*hi per K step -- three bf16 MMAs cost 1.5 TF32-MMA equivalents (3xTF32 costs 3), and the
//     operand bytes in shared / tensor memory halve.  Dropped terms (lo*lo and the bf16 rounding of lo) are ~2^-17 per
//     product: measured 5e-6..1.3e-5 blob-level against the 3xTF32 kernel (profiles/r02_c2_diag.log), 100x inside 1e-3.
//   * Epilogue: double-buffered TMEM accumulators, tcgen05.ld, (+ bias), 32-channel chunks transposed through shared
//     memory [channel][128 pixels] and written by the TMA unit when the tile lies inside one image: one tensor store per
//     chunk (box {128, 32, 1} of the output blob, clipped by the hardware at the image end and the channel count; the start
//     column is never negative -- a negative store coordinate is an illegal-instruction fault, measured).  Tiles that span
//     two images write the staged chunk with 16-byte st.global.
//   * Plane mode (BWT < 128): 7x7 maps, whole image planes staged -- see Smem below.
// dgrad of a stride-1 same convolution is the same kernel on dY with the transposed + flipped filter.
#include <cuda.h>
#include <cuda_bf16.h>
#include <limits.h>
#include <stdlib.h>
#include <stdio.h>
#include "b2c_common.cuh"
#define B2C_MBAR_DEBUG 1     // this kernel records stuck mbarrier waits (tc_common.cuh)
#include "tc_common.cuh"
#include "filter_prep.cuh"

namespace b2c {
using namespace tc;

namespace stg {

constexpr int NCW = 16;                       // converter warps
constexpr int W_BTMA = NCW;                   // filter TMA warp
constexpr int W_MMA = NCW + 1;                // MMA issuer (also owns the TMEM allocation)
constexpr int W_STG = NCW + 2;                // activation TMA warp
constexpr int W_EPI0 = NCW + 3;               // first of 4 epilogue warps
constexpr int THREADS = (NCW + 7) * 32;       // 736
constexpr int CB = 32;                        // channels per staged group
constexpr int BKE = 64;                       // bf16 K elements per pipeline stage (two (group, tap) half blocks)
constexpr int STAGES = 3;                     // operand pipeline depth (B in smem, A in TMEM)
constexpr uint32_t STG_POOL = 96u * 1024u;    // activation staging pool, carved into slots of one TMA box each
constexpr int MAX_SLOTS = 6;
constexpr int MAX_TAPS = 32;                  // per-row validity mask is one 32-bit word
constexpr int PLANE_IMGS = 4;                 // plane mode: images per staging slot (a 128-row tile of 7x7 maps touches <= 4)

struct Params {
  int HW, H, W;
  int kh, kw, ph, pw;
  int Mtot;                // Nimg * H * W (output pixel grid == input pixel grid)
  int Ntot;                // GEMM columns (output channels)
  int taps, groups;        // kh*kw, Cin / 32
  int nhb, nkb;            // half blocks = groups * taps, K blocks = ceil(nhb / 2)
  int halo;                // pad*W + pad
  const float* bias;       // [Ntot] or null
  float* out;              // [Nimg, Ntot, H, W]: the epilogue's st.global path (tiles that span two images)
  int m_tiles, n_tiles, total_tiles;
  int dbg;                 // B2C_STG_DBG bit 0: never use TMA stores (st.global epilogue for every tile)
  int accumulate;          // out += result instead of out = result (TMA reduce-add store; the fan-out accumulation of Net::Backward)
  long long* prof;
};

template <int N_TILE, int BWT>
struct Smem {
  static constexpr uint32_t B_BYTES = (uint32_t)N_TILE * 128u;              // [N_TILE rows][64 bf16], SW128
  static constexpr uint32_t STAGE = 2u * B_BYTES;                           // hi + lo
  static constexpr uint32_t STG_OFF = STAGES * STAGE;
  // BWT >= 128: one box = [32 channels][BWT pixels] fp32.  BWT < 128 ("plane mode", maps of BWT = H*W pixels whose
  // channel pitch TMA cannot address, 7x7 = 196 bytes): one slot = up to PLANE_IMGS whole images x [32 channels][H*W]
  static constexpr bool PLANE = BWT < 128;
  static constexpr uint32_t IMG_BYTES = (uint32_t)CB * (uint32_t)BWT * 4u;
  static constexpr uint32_t SLOT_BYTES = PLANE ? PLANE_IMGS * IMG_BYTES : IMG_BYTES;
  static constexpr uint32_t PITCH = (uint32_t)BWT * 4u;                     // bytes between channels of a staged box
  static_assert(!PLANE || IMG_BYTES % 128u == 0, "TMA destinations are 128-byte aligned");
  static_assert(!PLANE || 127 / BWT + 2 <= PLANE_IMGS, "a 128-row tile touches at most PLANE_IMGS images");
  static constexpr int NSLOT_ = (int)(STG_POOL / SLOT_BYTES);
  static constexpr int NSLOT = NSLOT_ > MAX_SLOTS ? MAX_SLOTS : NSLOT_;
  static constexpr uint32_t EPI_OFF = STG_OFF + STG_POOL;                   // 2 x [32 channels][128 pixels] fp32
  static constexpr uint32_t EPI_BYTES = 2u * 32u * 128u * 4u;
  static constexpr uint32_t BAR_OFF = EPI_OFF + EPI_BYTES;
  static constexpr uint32_t TAP_OFF = BAR_OFF + 256;                        // int tap offsets [MAX_TAPS]
  static constexpr uint32_t TOTAL = TAP_OFF + MAX_TAPS * 4 + 1024;          // + alignment slack
  static constexpr uint32_t TX_BYTES = STAGE;
  static constexpr uint32_t A_COL0 = 2u * N_TILE;                           // after the two accumulators
  static constexpr uint32_t A_COLS = 64u;                                   // 32 packed hi columns + 32 packed lo columns
  static_assert(A_COL0 + STAGES * A_COLS <= 512u, "TMEM budget");
  static_assert(NSLOT >= 2, "a tile that spans two images needs two boxes per channel group");
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
// D=F32, A=B=BF16, both K-major
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
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// out[box] += smem tile: the TMA unit's reduction store (fp32 add in L2), same clipping as the plain store
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* map, uint32_t smem_src, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
template <int OFF>
__device__ __forceinline__ float lds32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(OFF));
  return v;
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// fp32 pair (k even, k odd) -> packed bf16 hi word and packed bf16 lo word; element k sits in the LOW half
// (tensor-memory A operand of kind::f16: two consecutive K elements per 32-bit column, little endian)
__device__ __forceinline__ void split_pack_bf16(float e, float o, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(e, o);            // .x = e (low half), .y = o
  const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h);
  const float he = __uint_as_float(hb << 16), ho = __uint_as_float(hb & 0xffff0000u);
  const __nv_bfloat162 l = __floats2bfloat162_rn(e - he, o - ho);   // a - hi is exact in fp32
  hi = hb;
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

template <int N_TILE, int BWT>
__global__ void __launch_bounds__(THREADS, 1)
igemm_stg_kernel(const __grid_constant__ Params p, const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
                 const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_y) {
  using S = Smem<N_TILE, BWT>;
  constexpr int NSLOT = S::NSLOT;
  constexpr bool PLANE = S::PLANE;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sptr = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t bar_full = sbase + S::BAR_OFF;                 // STAGES: 16 converter warps + the filter TMA's expect_tx
  const uint32_t bar_empty = bar_full + 8 * STAGES;             // STAGES: tcgen05.commit
  const uint32_t bar_tfull = bar_empty + 8 * STAGES;            // 2
  const uint32_t bar_tempty = bar_tfull + 16;                   // 2
  const uint32_t bar_sfull = bar_tempty + 16;                   // MAX_SLOTS: staged box landed
  const uint32_t bar_sempty = bar_sfull + 8 * MAX_SLOTS;        // MAX_SLOTS: the converters are done with the box
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sptr + S::BAR_OFF + 8 * (2 * STAGES + 4 + 2 * MAX_SLOTS));
  int* tapoff = reinterpret_cast<int*>(sptr + S::TAP_OFF);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, NCW + 1); mbar_init(bar_empty + 8 * s, 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(bar_tfull + 8 * b, 1); mbar_init(bar_tempty + 8 * b, 4); }
    for (int b = 0; b < MAX_SLOTS; ++b) { mbar_init(bar_sfull + 8 * b, 1); mbar_init(bar_sempty + 8 * b, NCW); }
    fence_barrier_init();
  }
  if (warp == W_MMA) tmem_alloc(smem_u32(tmem_slot), 512);
  if (tid < p.taps) {
    const int i = tid / p.kw, j = tid - i * p.kw;
    tapoff[tid] = ((i - p.ph) * p.W + (j - p.pw)) * 4;          // byte offset inside a staged channel row
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto stage_b_hi = [&](int s) { return sbase + (uint32_t)s * S::STAGE; };
  auto stage_b_lo = [&](int s) { return sbase + (uint32_t)s * S::STAGE + S::B_BYTES; };
  auto stage_a_col = [&](int s) { return S::A_COL0 + (uint32_t)s * S::A_COLS; };
  auto slot_addr = [&](int slot) { return sbase + S::STG_OFF + (uint32_t)slot * S::SLOT_BYTES; };
  auto tile_coords = [&](int tile, int& m0, int& n0) {
    const int nt = tile % p.n_tiles;
    m0 = (tile / p.n_tiles) * 128; n0 = nt * N_TILE;
  };
  // images a tile touches: 1 or 2 (H*W >= 128); the second one starts at tile row `split`
  // plane mode: nseg = 1 (all the tile's images share one slot), nimg = how many of them there are (<= PLANE_IMGS)
  auto tile_images = [&](int m0, int& n_first, int& nseg, int& split, int& nimg) {
    const int q_last = min(m0 + 128, p.Mtot) - 1;
    n_first = m0 / p.HW;
    nimg = q_last / p.HW - n_first + 1;
    nseg = PLANE ? 1 : nimg;
    split = PLANE ? 128 : (n_first + 1) * p.HW - m0;          // >= 128 when nseg == 1
  };
  const bool prof = p.prof != nullptr && blockIdx.x == 0;

  if (warp < NCW) {
    // ================= converters: staged fp32 -> bf16 hi/lo -> tensor memory ===============================
    const int rq = warp & 3, sub = warp >> 2;                 // TMEM lane quarter, channel octet [sub*8, sub*8+8) of every half block
    const int row = rq * 32 + lane;
    const uint32_t a_lane = (uint32_t)(rq * 32) << 16;
    long long c_t0 = 0;
    if (prof) c_t0 = clock64();
    int kbg = 0;                                               // K blocks processed by this CTA (stage ring position)
    int u_base = 0;                                            // staged boxes consumed before this tile
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      int m0, n0, n_first, nseg, split, nimg;
      tile_coords(tile, m0, n0);
      tile_images(m0, n_first, nseg, split, nimg);
      // per-row tap validity (zero padding / image borders / row wrap), once per tile
      uint32_t mask = 0;
      const int seg_r = row >= split ? 1 : 0;                  // which of the tile's images this row belongs to
      // plane mode: the row's image relative to the tile's first one (rows past the batch: any staged address will do)
      const int img_r = PLANE ? (m0 + row < p.Mtot ? (m0 + row) / p.HW - n_first : 0) : seg_r;
      {
        const int q = m0 + row;
        if (q < p.Mtot) {
          const int pp = q - (n_first + img_r) * p.HW;
          const int h = pp / p.W, w = pp - h * p.W;
          if (p.taps == 1) mask = 1u;
          else {
            int t = 0;
            for (int i = 0; i < p.kh; ++i) {
              const bool hok = (unsigned)(h + i - p.ph) < (unsigned)p.H;
              for (int j = 0; j < p.kw; ++j, ++t)
                if (hok && (unsigned)(w + j - p.pw) < (unsigned)p.W) mask |= 1u << t;
            }
          }
        }
      }
      // byte offset of this row at tap offset 0 inside its image's box, plus the channel octet of this warp.  Box column 0
      // is pixel max(first pixel of the segment - halo, 0) of the image rounded DOWN to a multiple of 4: the innermost TMA
      // coordinate has to be 16-byte aligned (an odd start column is an illegal-instruction fault, measured) and never
      // negative; the TMA unit zero-fills past the image END; taps that reach before pixel 0 belong to masked rows, whatever
      // they read is discarded
      // plane mode: whole images are staged, [image][32 channels][H*W]
      const int pp_r = m0 + row - (n_first + img_r) * p.HW;                       // this row's pixel inside its image
      const int start_r = (PLANE || seg_r) ? 0 : max((m0 - n_first * p.HW - p.halo) & ~3, 0);
      const uint32_t row_off = (uint32_t)((pp_r - start_r) * 4) + (uint32_t)(sub * 8) * S::PITCH + (PLANE ? (uint32_t)img_r * S::IMG_BYTES : 0u);
      int g = 0, tap = 0;
      uint32_t src_base = 0;
      for (int kb = 0; kb < p.nkb; ++kb, ++kbg) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (2 * kb + h < p.nhb) {
            const int u0 = u_base + g * nseg;
            if (tap == 0) {                                    // first tap of a channel group: its box(es) must have landed
              mbar_wait(bar_sfull + 8 * (u0 % NSLOT), (u0 / NSLOT) & 1);
              if (nseg == 2) mbar_wait(bar_sfull + 8 * ((u0 + 1) % NSLOT), ((u0 + 1) / NSLOT) & 1);
              src_base = slot_addr((u0 + seg_r) % NSLOT) + row_off;
            }
            const bool ok = (mask >> tap) & 1u;
            const uint32_t src = src_base + (uint32_t)tapoff[tap];
            float v[8];
            constexpr int CP = (int)S::PITCH;
            v[0] = lds32<0>(src); v[1] = lds32<CP>(src); v[2] = lds32<2 * CP>(src); v[3] = lds32<3 * CP>(src);
            v[4] = lds32<4 * CP>(src); v[5] = lds32<5 * CP>(src); v[6] = lds32<6 * CP>(src); v[7] = lds32<7 * CP>(src);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              uint32_t hw, lw;
              split_pack_bf16(v[2 * e], v[2 * e + 1], hw, lw);
              hi[h * 4 + e] = ok ? hw : 0u;                    // masked taps contribute exact zeros (never NaN from stale smem)
              lo[h * 4 + e] = ok ? lw : 0u;
            }
            if (++tap == p.taps) {                             // group fully consumed by this warp: hand its box(es) back
              tap = 0;
              __syncwarp();
              if (lane == 0) {
                mbar_arrive(bar_sempty + 8 * (u0 % NSLOT));
                if (nseg == 2) mbar_arrive(bar_sempty + 8 * ((u0 + 1) % NSLOT));
              }
              ++g;
            }
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) { hi[h * 4 + e] = 0u; lo[h * 4 + e] = 0u; }   // K padding: finite zeros
          }
        }
        const int s = kbg % STAGES, it = kbg / STAGES;
        mbar_wait_backoff(bar_empty + 8 * s, (it & 1) ^ 1, 32);
        tc_fence_after();
        const uint32_t a0 = tmem_base + a_lane + stage_a_col(s) + (uint32_t)(sub * 4);
        tmem_st4(a0, hi[0], hi[1], hi[2], hi[3]);
        tmem_st4(a0 + 16, hi[4], hi[5], hi[6], hi[7]);
        tmem_st4(a0 + 32, lo[0], lo[1], lo[2], lo[3]);
        tmem_st4(a0 + 48, lo[4], lo[5], lo[6], lo[7]);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_full + 8 * s);
      }
      u_base += p.groups * nseg;
    }
    if (prof && tid == 0) { p.prof[0] = clock64() - c_t0; p.prof[3] = kbg; }
  } else if (warp == W_BTMA) {
    // ================= filter TMA producer ====================================================================
    int kbg = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      int m0, n0;
      tile_coords(tile, m0, n0);
      for (int kb = 0; kb < p.nkb; ++kb, ++kbg) {
        const int s = kbg % STAGES, it = kbg / STAGES;
        mbar_wait_backoff(bar_empty + 8 * s, (it & 1) ^ 1, 32);
        if (elect_one()) {
          arrive_expect_tx(bar_full + 8 * s, S::TX_BYTES);
          tma_load_3d(stage_b_hi(s), &map_hi, bar_full + 8 * s, kb * BKE, n0, 0);
          tma_load_3d(stage_b_lo(s), &map_lo, bar_full + 8 * s, kb * BKE, n0, 0);
        }
        __syncwarp();
      }
    }
  } else if (warp == W_STG) {
    // ================= activation TMA producer: one box per channel group and image ==============================
    int u = 0;
    long long s_wait = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      int m0, n0, n_first, nseg, split, nimg;
      tile_coords(tile, m0, n0);
      tile_images(m0, n_first, nseg, split, nimg);
      for (int g = 0; g < p.groups; ++g) {
        for (int seg = 0; seg < nseg; ++seg, ++u) {
          const int slot = u % NSLOT;
          if (prof) { const long long t0 = clock64(); mbar_wait_backoff(bar_sempty + 8 * slot, ((u / NSLOT) & 1) ^ 1, 32); s_wait += clock64() - t0; }
          else mbar_wait_backoff(bar_sempty + 8 * slot, ((u / NSLOT) & 1) ^ 1, 32);
          if (elect_one()) {
            if constexpr (PLANE) {
              // the blob viewed as {4*H*W, C/4, N} (four channels per row: a 16-byte multiple pitch); box {4*H*W, 8, 1} is
              // 32 channels of one image, contiguous
              arrive_expect_tx(bar_sfull + 8 * slot, (uint32_t)nimg * S::IMG_BYTES);
              for (int k = 0; k < nimg; ++k)
                tma_load_3d(slot_addr(slot) + (uint32_t)k * S::IMG_BYTES, &map_x, bar_sfull + 8 * slot, 0, g * (CB / 4), n_first + k);
            } else {
              arrive_expect_tx(bar_sfull + 8 * slot, S::SLOT_BYTES);
              const int col0 = seg ? 0 : max((m0 - n_first * p.HW - p.halo) & ~3, 0);   // 16-byte aligned; columns past the image end read as zero
              tma_load_3d(slot_addr(slot), &map_x, bar_sfull + 8 * slot, col0, g * CB, n_first + seg);
            }
          }
          __syncwarp();
        }
      }
    }
    if (prof && lane == 0) { p.prof[4] = s_wait; p.prof[5] = u; }
  } else if (warp == W_MMA) {
    // ================= MMA issuer ==============================================================================
    constexpr uint32_t IDESC = idesc_bf16(128, N_TILE);
    int kbg = 0, ti = 0;
    long long m_full = 0, m_t0 = 0;
    if (prof) m_t0 = clock64();
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++ti) {
      const int buf = ti & 1, use = ti >> 1;
      mbar_wait(bar_tempty + 8 * buf, (use & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(buf * N_TILE);
      for (int kb = 0; kb < p.nkb; ++kb, ++kbg) {
        const int s = kbg % STAGES, it = kbg / STAGES;
        if (prof) { const long long t0 = clock64(); mbar_wait(bar_full + 8 * s, it & 1); m_full += clock64() - t0; }
        else mbar_wait(bar_full + 8 * s, it & 1);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < BKE / 16; ++kk) {
            const uint32_t ah = tmem_base + stage_a_col(s) + (uint32_t)(kk * 8);
            const uint64_t bh = desc_sw128(stage_b_hi(s) + kk * 32);
            const uint64_t bl = desc_sw128(stage_b_lo(s) + kk * 32);
            umma_bf16_ts(d_tmem, ah + 32, bh, IDESC, (kb | kk) != 0);   // lo * hi
            umma_bf16_ts(d_tmem, ah, bl, IDESC, 1);                      // hi * lo
            umma_bf16_ts(d_tmem, ah, bh, IDESC, 1);                      // hi * hi
          }
          umma_commit(bar_empty + 8 * s);
          if (kb == p.nkb - 1) umma_commit(bar_tfull + 8 * buf);
        }
        __syncwarp();
      }
    }
    if (prof && lane == 0) { p.prof[8] = clock64() - m_t0; p.prof[9] = m_full; p.prof[10] = kbg; }
  } else {
    // ================= epilogue warps ============================================================================
    // tcgen05.ld -> (+ bias) -> 32-channel chunk transposed into shared memory [channel][128 pixels] -> one TMA tensor
    // store per image the tile touches.  (The warps' own st.global path cost ~3 100 cycles per 128x128 tile on the
    // short-K 1x1 layers, profiles/r01_prof_1x1_fwd_store_ablation.txt; 32 one-channel bulk stores per chunk cost ~2 900.)
    const int lg = warp & 3;                  // TMEM lane quarter this warp may read (warp id % 4)
    const int r = lg * 32 + lane;
    const bool issuer = warp == W_EPI0;
    float* epi = reinterpret_cast<float*>(sptr + S::EPI_OFF);
    int epi_chunk = 0, ti = 0;
    long long e_wait = 0, e_work = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++ti) {
      int m0, n0, n_first, nseg, split, nimg;
      tile_coords(tile, m0, n0);
      tile_images(m0, n_first, nseg, split, nimg);
      const int buf = ti & 1, use = ti >> 1;
      const float* brow = p.bias ? p.bias + n0 : nullptr;
      long long e0 = 0;
      if (prof) e0 = clock64();
      mbar_wait_backoff(bar_tfull + 8 * buf, use & 1, 128);
      long long e1 = 0;
      if (prof) { e1 = clock64(); e_wait += e1 - e0; }
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < N_TILE; c0 += 32, ++epi_chunk) {
        if (n0 + c0 >= p.Ntot) break;                     // uniform over the four epilogue warps
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(buf * N_TILE + c0), v);
        if (c0 + 32 >= N_TILE || n0 + c0 + 32 >= p.Ntot) {   // accumulator fully read: hand it back before storing
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
        }
        if (brow) {
          const int nb = min(32, p.Ntot - n0 - c0);
#pragma unroll
          for (int j = 0; j < 32; ++j) if (j < nb) v[j] += __ldg(brow + c0 + j);    // warp-uniform address: broadcast
        }
        if constexpr (PLANE) {
          // lanes = consecutive pixels of an image plane: every channel's store is one (unaligned) 128-byte run per warp,
          // straight from the registers
          const int q = m0 + r;
          if (q < p.Mtot) {
            const int nq = q / p.HW;
            float* ob = p.out + ((size_t)nq * p.Ntot + n0 + c0) * p.HW + (q - nq * p.HW);
            const int nb = min(32, p.Ntot - n0 - c0);
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < nb) { float* d = ob + (size_t)j * p.HW; *d = p.accumulate ? *d + v[j] : v[j]; }
          }
          continue;
        }
        // the store issued two chunks ago from this staging buffer has finished reading it (elect.sync on a converged
        // warp always elects the same lane, which owns the bulk-async groups)
        if (issuer && elect_one()) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
        float* E = epi + (epi_chunk & 1) * (32 * 128);
#pragma unroll
        for (int j = 0; j < 32; ++j) E[j * 128 + r] = v[j];           // lanes = consecutive pixels: conflict-free
        fence_proxy_async();                                           // generic-proxy writes -> visible to the TMA store
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (nseg == 1 && !(p.dbg & 1)) {
          // the tile lies inside one image: box column 0 = tile row 0 = pixel m0 - n*HW >= 0; columns past the image end
          // (last tile of an image, rows past the batch) and channels past Ntot are clipped by the TMA unit
          if (issuer && elect_one()) {
            if (p.accumulate) tma_reduce_add_3d(&map_y, smem_u32(E), m0 - n_first * p.HW, n0 + c0, n_first);
            else tma_store_3d(&map_y, smem_u32(E), m0 - n_first * p.HW, n0 + c0, n_first);
          }
        } else {
          // the tile spans two images (1 in 24 tiles at 56x56, 2 in 3 at 14x14): 16-byte st.global from the staged chunk
          const int mv = m0 + lane * 4;                    // this lane's 4 pixels (H*W % 4 == 0: never across an image end)
          if (mv < p.Mtot) {
            const int nv = mv / p.HW, pv = mv - nv * p.HW;
            float* vbase = p.out + ((size_t)nv * p.Ntot + n0 + c0) * p.HW + pv;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int ch = lg * 8 + q;
              if (n0 + c0 + ch < p.Ntot) {
                float4 t = *reinterpret_cast<const float4*>(E + ch * 128 + lane * 4);
                float4* dst = reinterpret_cast<float4*>(vbase + (size_t)ch * p.HW);
                if (p.accumulate) { const float4 o = *dst; t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w; }
                *dst = t;
              }
            }
          }
        }
        if (issuer && elect_one()) asm volatile("cp.async.bulk.commit_group;" ::: "memory");   // one (possibly empty) group per chunk
      }
      if (prof) e_work += clock64() - e1;
    }
    if (issuer && elect_one()) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (prof && r == 0) { p.prof[16] = e_wait; p.prof[17] = e_work; p.prof[18] = ti; }
  }
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace stg

// ---- host side ------------------------------------------------------------------------------------------------------
typedef CUresult (*StgEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static StgEncodeTiledFn stg_encode_tiled() {
  static StgEncodeTiledFn fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<StgEncodeTiledFn>(f);
  }
  return fn;
}
static int stg_make_filter_map(CUtensorMap* map, const void* base, int Kp, int rows, int n_tile) {
  StgEncodeTiledFn enc = stg_encode_tiled();
  if (!enc) return fail(B2C_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[3] = {(cuuint64_t)Kp, (cuuint64_t)rows, 1};
  cuuint64_t strides[2] = {(cuuint64_t)Kp * 2, (cuuint64_t)Kp * 2 * (cuuint64_t)rows};
  cuuint32_t box[3] = {64, (cuuint32_t)n_tile, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(B2C_ERR_CUDA, "cuTensorMapEncodeTiled (staged conv filter) failed (%d)", (int)r);
  return B2C_OK;
}

// [N][rows][HW] fp32, box {bw, 32, 1}, no swizzle, out-of-range elements read as zero / are not written
static int stg_make_act_map(CUtensorMap* map, const float* base, int HW, int rows, int N, int bw) {
  StgEncodeTiledFn enc = stg_encode_tiled();
  if (!enc) return fail(B2C_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[3] = {(cuuint64_t)HW, (cuuint64_t)rows, (cuuint64_t)N};
  cuuint64_t strides[2] = {(cuuint64_t)HW * 4, (cuuint64_t)HW * 4 * (cuuint64_t)rows};
  cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)stg::CB, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(B2C_ERR_CUDA, "cuTensorMapEncodeTiled (staged conv activations) failed (%d)", (int)r);
  return B2C_OK;
}

// plane mode: the [N][C][HW] blob viewed as {4*HW, C/4, N}; box {4*HW, 8, 1} = 32 channels of one image
static int stg_make_plane_map(CUtensorMap* map, const float* base, int HW, int C, int N) {
  StgEncodeTiledFn enc = stg_encode_tiled();
  if (!enc) return fail(B2C_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[3] = {(cuuint64_t)HW * 4, (cuuint64_t)C / 4, (cuuint64_t)N};
  cuuint64_t strides[2] = {(cuuint64_t)HW * 16, (cuuint64_t)HW * 4 * (cuuint64_t)C};
  cuuint32_t box[3] = {(cuuint32_t)HW * 4, (cuuint32_t)stg::CB / 4, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(B2C_ERR_CUDA, "cuTensorMapEncodeTiled (staged conv, plane mode) failed (%d)", (int)r);
  return B2C_OK;
}

static bool stg_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("B2C_CONV_STAGED"); on = e ? atoi(e) : 1; }
  return on != 0;
}

static bool stg_plane_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("B2C_CONV_STAGED_PLANE"); on = e ? atoi(e) : 1; }
  return on != 0;
}

// geometry shared by the eligibility test, the workspace size and the launch
struct StgGeom { int Cin, Cout, halo, bwt, taps, groups, nhb, nkb, Kp; };
static bool stg_geom(const ConvShape& s, int op, StgGeom* g) {
  if (!stg_enabled()) return false;
  if (op != B2C_OP_FORWARD && op != B2C_OP_BACKWARD_DATA) return false;
  if (s.G != 1 || s.sh != 1 || s.sw != 1 || s.dh != 1 || s.dw != 1) return false;
  if (s.Ho != s.H || s.Wo != s.W || 2 * s.ph != s.kh - 1 || 2 * s.pw != s.kw - 1) return false;     // "same" convolution
  const long long HW = (long long)s.H * s.W;
  if ((long long)s.N * HW > 0x7fffffffLL - 256) return false;
  const int Cin = op == B2C_OP_FORWARD ? s.C : s.O, Cout = op == B2C_OP_FORWARD ? s.O : s.C;
  if (Cin % stg::CB != 0 || Cout < 32) return false;
  const int taps = s.kh * s.kw;
  if (taps > stg::MAX_TAPS) return false;
  const int halo = s.ph * s.W + s.pw;
  int bwt;
  if (HW == 49) {
    // 7x7 maps (ResNet-50 res5, GoogLeNet inception 5): a 196-byte channel pitch is not a TMA stride; whole image planes are
    // staged instead ("plane mode", template parameter BWT = H*W)
    if (!stg_plane_enabled()) return false;
    bwt = 49;
  } else {
    if (HW % 4 != 0 || HW < 128) return false;               // 16-byte TMA pitches; <= 2 images per tile
    const int need = 128 + 2 * halo + ((4 - halo % 4) % 4);   // staged pixels per channel (+ what rounding the box start down to 16 bytes costs)
    bwt = need <= 128 ? 128 : need <= 160 ? 160 : need <= 192 ? 192 : need <= 256 ? 256 : 0;   // TMA boxes are <= 256 wide
    if (!bwt) return false;
  }
  if (g) {
    g->Cin = Cin; g->Cout = Cout; g->halo = halo; g->bwt = bwt; g->taps = taps;
    g->groups = Cin / stg::CB; g->nhb = g->groups * taps; g->nkb = (g->nhb + 1) / 2; g->Kp = g->nkb * stg::BKE;
  }
  return true;
}
bool tc_stg_supported(const ConvShape& s, int op) { return stg_geom(s, op, nullptr); }
size_t tc_stg_workspace(const ConvShape& s, int op) {
  StgGeom g;
  if (!stg_geom(s, op, &g)) return 0;
  return 2 * sizeof(__nv_bfloat16) * (size_t)g.Cout * g.Kp + 256;
}

template <int N_TILE, int BWT>
static int stg_launch_inst(const stg::Params& p, const CUtensorMap& mh, const CUtensorMap& ml, const CUtensorMap& mx, const CUtensorMap& my,
                           cudaStream_t st) {
  using S = stg::Smem<N_TILE, BWT>;
  B2C_CUDA_OK(cudaFuncSetAttribute(stg::igemm_stg_kernel<N_TILE, BWT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::TOTAL));
  const int grid = p.total_tiles < sm_count() ? p.total_tiles : sm_count();
  stg::igemm_stg_kernel<N_TILE, BWT><<<grid, stg::THREADS, S::TOTAL, st>>>(p, mh, ml, mx, my);
  B2C_POST_LAUNCH();
  return B2C_OK;
}
template <int N_TILE>
static int stg_launch_n(int bwt, const stg::Params& p, const CUtensorMap& mh, const CUtensorMap& ml, const CUtensorMap& mx,
                        const CUtensorMap& my, cudaStream_t st) {
  switch (bwt) {
    case 49: return stg_launch_inst<N_TILE, 49>(p, mh, ml, mx, my, st);
    case 128: return stg_launch_inst<N_TILE, 128>(p, mh, ml, mx, my, st);
    case 160: return stg_launch_inst<N_TILE, 160>(p, mh, ml, mx, my, st);
    case 192: return stg_launch_inst<N_TILE, 192>(p, mh, ml, mx, my, st);
    default: return stg_launch_inst<N_TILE, 256>(p, mh, ml, mx, my, st);
  }
}

// The filter operand of `op` in this kernel's GEMM layout, written into `dst` (tc_stg_workspace(s, op) bytes).
bool tc_stg_prep_entry(const ConvShape& s, int op, const float* w, void* dst, PrepEntry* q) {
  StgGeom g;
  if (!stg_geom(s, op, &g)) return false;
  __nv_bfloat16* wbase = reinterpret_cast<__nv_bfloat16*>((reinterpret_cast<uintptr_t>(dst) + 255) & ~(uintptr_t)255);
  q->w = w; q->kind = 1; q->G = 1; q->Og = s.O; q->Cg = s.C; q->taps = g.taps; q->rows = g.Cout; q->K = g.Cin * g.taps; q->Kp = g.Kp;
  q->mode = op == B2C_OP_FORWARD ? 0 : 1; q->tap_major = 0; q->flip = q->mode; q->kch = g.Cin;
  q->total = (long long)g.Cout * g.Kp;
  q->hi = wbase; q->lo = wbase + q->total;
  return true;
}

// a = x (forward) or dy (dgrad); b = w; out = y or dx; `prepared`: filter already in GEMM layout, or null = prepass here
int launch_conv_tc_stg(const ConvShape& s, int op, const float* a, const float* w, const float* bias, float* out, void* ws,
                       size_t ws_bytes, const void* prepared, bool accumulate, cudaStream_t st) {
  StgGeom g;
  if (!stg_geom(s, op, &g)) return fail(B2C_ERR_INVALID, "staged tcgen05 conv: shape not eligible");
  if (!prepared && (!ws || ws_bytes < tc_stg_workspace(s, op))) return fail(B2C_ERR_WORKSPACE, "staged tcgen05 conv: workspace too small");
  if ((reinterpret_cast<uintptr_t>(a) & 15u) || (reinterpret_cast<uintptr_t>(out) & 15u))
    return fail(B2C_ERR_INVALID, "staged tcgen05 conv: activations must be 16-byte aligned");
  PrepEntry q;
  tc_stg_prep_entry(s, op, w, prepared ? const_cast<void*>(prepared) : ws, &q);
  if (!prepared)
    if (int rc = launch_filter_prep(&q, 1, st)) return rc;

  stg::Params p;
  p.H = s.H; p.W = s.W; p.HW = s.H * s.W;
  p.kh = s.kh; p.kw = s.kw; p.ph = s.ph; p.pw = s.pw;
  p.Mtot = s.N * p.HW; p.Ntot = g.Cout;
  p.taps = g.taps; p.groups = g.groups; p.nhb = g.nhb; p.nkb = g.nkb;
  p.halo = g.halo;
  p.bias = op == B2C_OP_FORWARD ? bias : nullptr;
  p.out = out;
  p.accumulate = accumulate ? 1 : 0;
  { static int dbg = -1; if (dbg < 0) { const char* e = getenv("B2C_STG_DBG"); dbg = e ? atoi(e) : 0; } p.dbg = dbg; }
  const int n_tile = g.Cout > 64 ? 128 : g.Cout > 32 ? 64 : 32;
  p.m_tiles = (p.Mtot + 127) / 128; p.n_tiles = (g.Cout + n_tile - 1) / n_tile; p.total_tiles = p.m_tiles * p.n_tiles;
  static long long* prof_buf = nullptr;
  static int prof_on = -1;
  if (prof_on < 0) { const char* e = getenv("B2C_PROF"); prof_on = e ? atoi(e) : 0; if (prof_on) cudaMalloc(&prof_buf, 32 * sizeof(long long)); }
  p.prof = prof_on ? prof_buf : nullptr;
  if (prof_on) cudaMemsetAsync(prof_buf, 0, 32 * sizeof(long long), st);
  alignas(64) CUtensorMap mh, ml, mx, my;
  if (int rc = stg_make_filter_map(&mh, q.hi, g.Kp, g.Cout, n_tile)) return rc;
  if (int rc = stg_make_filter_map(&ml, q.lo, g.Kp, g.Cout, n_tile)) return rc;
  if (g.bwt < 128) {
    if (int rc = stg_make_plane_map(&mx, a, p.HW, g.Cin, s.N)) return rc;
    my = mx;                                                  // plane mode stores with st.global
  } else {
    if (int rc = stg_make_act_map(&mx, a, p.HW, g.Cin, s.N, g.bwt)) return rc;
    if (int rc = stg_make_act_map(&my, out, p.HW, g.Cout, s.N, 128)) return rc;
  }
  int rc;
  switch (n_tile) {
    case 128: rc = stg_launch_n<128>(g.bwt, p, mh, ml, mx, my, st); break;
    case 64: rc = stg_launch_n<64>(g.bwt, p, mh, ml, mx, my, st); break;
    default: rc = stg_launch_n<32>(g.bwt, p, mh, ml, mx, my, st); break;
  }
  if (rc == B2C_OK && prof_on) {
    long long h[32];
    cudaMemcpy(h, prof_buf, sizeof(h), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[stg-prof] N_TILE=%d BWT=%d nkb=%d groups=%d taps=%d tiles=%d | conv(t0): total=%lld kblocks=%lld | act-tma: wait_empty=%lld boxes=%lld | mma: total=%lld wait_full=%lld | epi: wait=%lld work=%lld tiles=%lld\n",
            n_tile, g.bwt, g.nkb, g.groups, g.taps, p.total_tiles, h[0], h[3], h[4], h[5], h[8], h[9], h[16], h[17], h[18]);
  }
  return rc;
}

TC_DEBUG_EXPORT(debug_mbar_stg)

}  // namespace b2c
