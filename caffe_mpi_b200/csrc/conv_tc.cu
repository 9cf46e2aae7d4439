// conv_tc.cu -- tcgen05 (5th-gen tensor core) implicit-GEMM convolution, forward and data-gradient (sm_100a).
//
// Replaces cudnnConvolutionForward / cudnnConvolutionBackwardData as called by the reference's
// CuDNNConvolutionLayer (src/caffe/layers/cudnn_conv_layer.cu:25-29,118-123) and, semantically, the
// per-image forward_gpu_gemm / backward_gpu_gemm loops (base_conv_layer.hpp:105-146): whole batch per
// launch, NCHW fp32 in and out, y / dx overwritten, bias fused, no col buffer in HBM.
//
// GEMM view (batch folded into the pixel dimension):
//   forward : D[q=(n,ho,wo)][o] = sum_k A[q][k] * B[o][k],  A = im2col gather of X, B = filter
//   dgrad   : the same kernel run on dY with the transposed (+ flipped) filter for stride-1 layers, or
//             on the top grid with a strided scatter epilogue for 1x1 / stride>1 / pad 0 layers.
//
// Kernel anatomy -- persistent, one CTA per SM, 128 x N_TILE output tiles, 704 threads:
//   warps 0-15 A producers.  Four threads per GEMM row (pixel), each owning two of the eight 16-byte K
//              chunks of a K block; the im2col gather reads NCHW global memory directly (coalesced along
//              W) from a per-kernel smem table of chunk offsets (the (c,i,j) decode is the same for every
//              tile), keeps two K blocks of loads in flight in registers across tile boundaries, splits
//              to TF32 hi (+ lo) on the ALU pipe and writes its row of the A tile straight into tensor
//              memory (tcgen05.st, lane = GEMM row).  K is ordered (channel/4, tap, channel%4) when C/g % 4 == 0.
//   warp  16   TMA producer for the filter operand: a prepass kernel writes the filter in GEMM-K order,
//              already split into TF32 hi / lo and zero padded to a multiple of 32; cp.async.bulk.tensor
//              (SWIZZLE_128B) drops [N_TILE x 32] boxes into smem, completion on the stage mbarrier.
//   warp  17   allocates TMEM and issues tcgen05.mma.kind::tf32 (M=128, N=N_TILE, K=8) with the A operand in
//              TENSOR MEMORY (written there by the producers with tcgen05.st) and B from smem: the 3xTF32 mode
//              is otherwise shared-memory-bandwidth bound (each K step reads A and B three times); tcgen05.commit releases smem stages and publishes finished accumulators.
//   warps 18-21 epilogue: tcgen05.ld TMEM -> registers -> bias -> coalesced NCHW stores, overlapped with
//              the next tile's main loop through the double-buffered accumulator.
// The activation operand is not TMA-staged: with NCHW the GEMM-K axis (c,i,j) is not unit-stride and
// 7x7 maps have 196-byte channel pitches (TMA needs 16-byte multiples); see DESIGN.md.
// fp32 math mode = 3 TF32 MMAs per K step (lo*hi, hi*lo, hi*hi); TF32 mode = 1.
#include <cuda.h>
#include <limits.h>
#include <stdlib.h>
#include <stdio.h>
#include "b2c_common.cuh"
#include "tc_common.cuh"
#include "filter_prep.cuh"

namespace b2c {
using namespace tc;

constexpr int NPW = 16;                     // A-producer warps
constexpr int FW_THREADS = (NPW + 6) * 32;  // + TMA warp, MMA warp, 4 epilogue warps = 704
constexpr int TAB_ENTRIES = 1152;   // gather table: one entry per K chunk (KMODE 1) or per K element (KMODE 0)

struct FwdParams {
  // A: activations [Nimg, Cin_tot, H, W]; group g reads channels [g*Cg, (g+1)*Cg)
  const float* x;
  int Cin_tot, H, W, Cg;
  int kh, kw, sh, sw, ph, pw, dh, dw;
  int Ho, Wo;        // pixel grid of the GEMM rows: m = (n*Ho + ho)*Wo + wo
  int Mtot;          // Nimg*Ho*Wo
  int Ntot;          // GEMM columns per group
  int K, Kp;         // true / padded (multiple of 32) reduction length
  // out[(n*Cout_tot + g*Ntot + col)*out_plane + ho*out_hs + wo*out_ws]
  float* out;
  int Cout_tot, out_hs, out_ws;
  long long out_plane;
  const float* bias;  // [G*Ntot] or null
  int m_tiles, n_tiles, G, total_tiles;
  long long* prof;    // optional cycle counters written by CTA 0 (B2C_PROF=1), else null
  int dbg;            // B2C_DBG experiments: 1 = skip the global gather (producer-side ceiling), 2 = skip the MMAs
};

// Resources of one CTA.  Shared memory holds only the filter (B) stages; the activation (A) operand
// lives in tensor memory next to the accumulators:
//   TMEM columns [0, ACC_BUFS*N_TILE)            fp32 accumulators (double buffered unless N_TILE = 256)
//                [A_COL0 + s*A_COLS, +A_COLS)     A stage s: 32 columns TF32 hi (+ 32 columns lo)
template <int N_TILE, bool SPLIT>
struct FwdSmem {
  static constexpr uint32_t B_BYTES = N_TILE * 128u;                       // [N_TILE rows][32 fp32], SW128
  static constexpr uint32_t NP = SPLIT ? 2u : 1u;
  static constexpr uint32_t STAGE = NP * B_BYTES;                          // multiple of 1024
  static constexpr int ACC_BUFS = N_TILE == 256 ? 1 : 2;
  static constexpr uint32_t A_COL0 = ACC_BUFS * N_TILE;
  static constexpr uint32_t A_COLS = SPLIT ? 64u : 32u;
  static constexpr int TMEM_STAGES = (int)((512u - A_COL0) / A_COLS);
  static constexpr int SMEM_STAGES = (int)((164u * 1024u) / STAGE);         // leaves room for the epilogue staging below
  static constexpr int STAGES_ = TMEM_STAGES < SMEM_STAGES ? TMEM_STAGES : SMEM_STAGES;
  static constexpr int STAGES = STAGES_ > 6 ? 6 : STAGES_;
  static constexpr uint32_t BAR_OFF = STAGES * STAGE;
  static constexpr uint32_t TAB_OFF = BAR_OFF + 256;
  static constexpr uint32_t EPI_OFF = TAB_OFF + TAB_ENTRIES * 8;           // 2 x [32 channels][128 pixels] fp32 (epilogue transpose)
  static constexpr uint32_t EPI_BYTES = 2u * 32u * 128u * 4u;
  static constexpr uint32_t TOTAL = EPI_OFF + EPI_BYTES + 1024;             // + alignment slack
  static constexpr uint32_t TX_BYTES = NP * B_BYTES;
};

__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;                    // LBO (ignored for swizzled K-major)
  d |= (uint64_t)(1024u >> 4) << 32;         // SBO: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                    // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void tma_load_3d(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

// KMODE 0: natural K order (c,i,j), per-element decode (any C/g).
// KMODE 1 (C/g % 4 == 0): K order (c/4, i, j, c%4): a 16-byte chunk = 4 channels of one tap (one bounds check),
// and the kh*kw taps of a channel block are consecutive chunks, so their overlapping pixel windows hit in L1
// instead of re-reading L2 kh*kw times.
template <int N_TILE, bool SPLIT, int KMODE>
__global__ void __launch_bounds__(FW_THREADS, 1)
igemm_fwd_kernel(const __grid_constant__ FwdParams p, const __grid_constant__ CUtensorMap map_hi,
                 const __grid_constant__ CUtensorMap map_lo) {
  using S = FwdSmem<N_TILE, SPLIT>;
  constexpr int STAGES = S::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;   // SW128 tiles need 1024-byte alignment
  uint8_t* sptr = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t bar_full = sbase + S::BAR_OFF;
  const uint32_t bar_empty = bar_full + 8 * STAGES;
  const uint32_t bar_tfull = bar_empty + 8 * STAGES;     // 2 x 8 B
  const uint32_t bar_tempty = bar_tfull + 16;            // 2 x 8 B
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sptr + S::BAR_OFF + 8 * (2 * STAGES + 4));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nkb = p.Kp / BK;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, NPW + 1);   // one elected arrive per producer warp + the TMA thread's expect_tx
      mbar_init(bar_empty + 8 * s, 1);        // one tcgen05.commit
    }
    for (int b = 0; b < 2; ++b) { mbar_init(bar_tfull + 8 * b, 1); mbar_init(bar_tempty + 8 * b, 4); }
    fence_barrier_init();
  }
  if (warp == NPW + 1) tmem_alloc(smem_u32(tmem_slot), 512);
  // gather table (identical for every tile): element offset inside one image's group slab + tap offsets
  {
    int2* tab = reinterpret_cast<int2*>(sptr + S::TAB_OFF);
    const int HWi = p.H * p.W;
    const int entries = KMODE == 1 ? p.Kp / 4 : p.Kp;
    for (int e = tid; e < entries; e += FW_THREADS) {
      int2 t = make_int2(INT_MIN, 0);
      const int k = KMODE == 1 ? e * 4 : e;
      if (k < p.K) {
        int c, i, j;
        if (KMODE == 1) { const int taps = p.kh * p.kw; const int cb = k / (4 * taps); const int tap = (k - cb * 4 * taps) >> 2;
                          c = cb * 4; i = tap / p.kw; j = tap - i * p.kw; }
        else { c = k / (p.kh * p.kw); const int tap = k - c * p.kh * p.kw; i = tap / p.kw; j = tap - i * p.kw; }
        t.x = c * HWi + i * p.dh * p.W + j * p.dw;
        t.y = (i * p.dh) | ((j * p.dw) << 16);
      }
      tab[e] = t;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // smem stage layout: [B_hi | B_lo], 1024-aligned; A stages are TMEM columns
  auto stage_b_hi = [&](int s) { return sbase + s * S::STAGE; };
  auto stage_b_lo = [&](int s) { return sbase + s * S::STAGE + S::B_BYTES; };
  auto stage_a_col = [&](int s) { return S::A_COL0 + (uint32_t)s * S::A_COLS; };

  auto tile_coords = [&](int tile, int& m0, int& n0, int& g) {
    const int nt = tile % p.n_tiles;
    const int r = tile / p.n_tiles;
    const int mt = r % p.m_tiles;
    g = r / p.m_tiles;
    m0 = mt * 128; n0 = nt * N_TILE;
  };

  if (warp < NPW) {
    // ================= A producers =====================================================================
    // 16 warps: four threads per GEMM row, each owning two of the eight 16-byte K chunks of every K block.
    // (A warp is one sequential instruction stream: with 4-8 producer warps the ~170 dependent instructions
    // per K block and warp, not bandwidth, set the K-block period -- measured ~1450 cycles regardless of
    // MMA count, tile width or staging path.  More, lighter warps cut that chain to ~60 instructions.)
    const int2* tab = reinterpret_cast<const int2*>(sptr + S::TAB_OFF);
    const long long HW = (long long)p.H * p.W;
    const int P = p.Ho * p.Wo;
    const int row = tid & 127, cp = tid >> 7;              // chunk pair 0..3 -> chunks 2cp, 2cp+1 (8 K columns)
    const uint32_t a_lane = (uint32_t)((warp & 3) * 32) << 16;   // TMEM lane quarter of this warp; row == (warp&3)*32 + lane
    const int my_tiles = (p.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int total_blocks = my_tiles * nkb;
    const float* xrow = p.x;
    int ih0 = 0, iw0 = 0;
    bool mvalid = false;
    int lt = 0, lkb = 0;       // load cursor: local tile index, K block inside it

    long long pw_load = 0;
    auto load_block_ = [&](float (&v)[8]) {
      if (lkb == 0) {
        int m0, n0, g;
        tile_coords((int)blockIdx.x + lt * (int)gridDim.x, m0, n0, g);
        const int m = m0 + row;
        mvalid = m < p.Mtot;
        const int mm = mvalid ? m : 0;
        const int n = mm / P, pix = mm - n * P;
        const int ho = pix / p.Wo, wo = pix - ho * p.Wo;
        ih0 = ho * p.sh - p.ph; iw0 = wo * p.sw - p.pw;
        xrow = p.x + ((long long)n * p.Cin_tot + (long long)g * p.Cg) * HW + (long long)ih0 * p.W + iw0;
      }
      if (p.dbg & 1) {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = 1.0f;
      } else if (KMODE == 1) {
        const int2* te = tab + lkb * KCHUNKS + cp * 2;
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          const int2 tt = te[ch];
          const bool ok = mvalid && tt.x != INT_MIN && (unsigned)(ih0 + (tt.y & 0xffff)) < (unsigned)p.H &&
                          (unsigned)(iw0 + (tt.y >> 16)) < (unsigned)p.W;
          const float* src = xrow + tt.x;
#pragma unroll
          for (int e = 0; e < 4; ++e) v[ch * 4 + e] = ok ? __ldg(src + e * HW) : 0.0f;
        }
      } else {
        const int2* te = tab + lkb * BK + cp * 8;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int2 tt = te[e];
          const bool ok = mvalid && tt.x != INT_MIN && (unsigned)(ih0 + (tt.y & 0xffff)) < (unsigned)p.H &&
                          (unsigned)(iw0 + (tt.y >> 16)) < (unsigned)p.W;
          v[e] = ok ? __ldg(xrow + tt.x) : 0.0f;
        }
      }
      if (++lkb == nkb) { lkb = 0; ++lt; }
    };
    auto load_block = [&](float (&v)[8]) { const long long a0 = clock64(); load_block_(v); pw_load += clock64() - a0; };
    int kbg = 0;   // pipeline position of the store cursor
    long long pw_empty = 0, pw_store = 0;
    const long long pt0 = clock64();
    auto store_block = [&](const float (&v)[8]) {
      const int s = kbg % STAGES, it = kbg / STAGES;
      const long long t0 = clock64();
      mbar_wait_backoff(bar_empty + 8 * s, (it & 1) ^ 1, 40);
      const long long t1 = clock64();
      tc_fence_after();
      const uint32_t a_hi = tmem_base + a_lane + stage_a_col(s) + (uint32_t)(cp * 8);
      if (SPLIT) {
        float hi[8], lo[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) split_tf32(v[e], hi[e], lo[e]);
        tmem_st8(a_hi, hi);
        tmem_st8(a_hi + 32, lo);
      } else {
        float hi[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) hi[e] = to_tf32(v[e]);
        tmem_st8(a_hi, hi);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_full + 8 * s);
      ++kbg;
      pw_empty += t1 - t0; pw_store += clock64() - t1;
    };

    // two K blocks of loads in flight ahead of the block being stored, across tile boundaries
    float v0[8], v1[8], v2[8];
    int issued = 0;
    if (issued < total_blocks) { load_block(v0); ++issued; }
    if (issued < total_blocks) { load_block(v1); ++issued; }
    for (int j = 0; j < total_blocks; j += 3) {
      if (issued < total_blocks) { load_block(v2); ++issued; }
      store_block(v0);
      if (j + 1 >= total_blocks) break;
      if (issued < total_blocks) { load_block(v0); ++issued; }
      store_block(v1);
      if (j + 2 >= total_blocks) break;
      if (issued < total_blocks) { load_block(v1); ++issued; }
      store_block(v2);
    }
    if (p.prof && blockIdx.x == 0 && tid == 0) {
      p.prof[0] = clock64() - pt0; p.prof[1] = pw_empty; p.prof[2] = pw_store; p.prof[3] = pw_load; p.prof[4] = total_blocks;
    }
  } else if (warp == NPW) {
    // ================= TMA producer (filter operand) ===================================================
    // (whole warp runs the loop converged; one elected lane issues)
    {
      int kbg = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int m0, n0, g;
        tile_coords(tile, m0, n0, g);
        for (int kb = 0; kb < nkb; ++kb, ++kbg) {
          const int s = kbg % STAGES, it = kbg / STAGES;
          mbar_wait_backoff(bar_empty + 8 * s, (it & 1) ^ 1, 40);
          if (elect_one()) {
            if (p.dbg & 8) {
              mbar_arrive(bar_full + 8 * s);
            } else {
              mbar_arrive_expect_tx(bar_full + 8 * s, S::TX_BYTES);
              tma_load_3d(stage_b_hi(s), &map_hi, bar_full + 8 * s, kb * BK, n0, g);
              if (SPLIT) tma_load_3d(stage_b_lo(s), &map_lo, bar_full + 8 * s, kb * BK, n0, g);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == NPW + 1) {
    // ================= MMA issuer ========================================================================
    // whole warp converged; the elected lane issues tcgen05.mma / tcgen05.commit
    {
      constexpr uint32_t IDESC = idesc_tf32(128, N_TILE);
      int kbg = 0, ti = 0;
      long long mw_full = 0, mw_issue = 0, mw_tempty = 0;
      const long long mt0 = clock64();
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++ti) {
        const int buf = ti % S::ACC_BUFS, use = ti / S::ACC_BUFS;
        const long long q0 = clock64();
        mbar_wait(bar_tempty + 8 * buf, (use & 1) ^ 1);         // epilogue has drained this accumulator
        mw_tempty += clock64() - q0;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * N_TILE);
        uint32_t acc = 0;
        for (int kb = 0; kb < nkb; ++kb, ++kbg) {
          const int s = kbg % STAGES, it = kbg / STAGES;
          const long long f0 = clock64();
          mbar_wait(bar_full + 8 * s, it & 1);
          const long long f1 = clock64();
          mw_full += f1 - f0;
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < BK / 8; ++kk) {
              if (p.dbg & 2) break;
              const uint32_t ah = tmem_base + stage_a_col(s) + (uint32_t)(kk * 8);
              const uint64_t bh = smem_desc_sw128(stage_b_hi(s) + kk * 32);
              if (SPLIT) {
                const uint64_t bl = smem_desc_sw128(stage_b_lo(s) + kk * 32);
                umma_tf32_ts(d_tmem, ah + 32, bh, IDESC, (kb | kk) != 0);   // lo * hi
                umma_tf32_ts(d_tmem, ah, bl, IDESC, 1);                      // hi * lo
                umma_tf32_ts(d_tmem, ah, bh, IDESC, 1);                      // hi * hi
              } else {
                umma_tf32_ts(d_tmem, ah, bh, IDESC, (kb | kk) != 0);
              }
            }
            umma_commit(bar_empty + 8 * s);
            if (kb == nkb - 1) umma_commit(bar_tfull + 8 * buf);
          }
          __syncwarp();
          mw_issue += clock64() - f1;
        }
        (void)acc;
      }
      if (p.prof && blockIdx.x == 0 && lane == 0) { p.prof[8] = clock64() - mt0; p.prof[9] = mw_full; p.prof[10] = mw_issue; p.prof[11] = mw_tempty; }
    }
  } else {
    // ================= epilogue warps NPW+2 .. NPW+5 ============================================================
    const int lg = warp & 3;                 // TMEM lane group this warp may access
    const int r = lg * 32 + lane;            // GEMM row inside the tile
    const int P = p.Ho * p.Wo;
    int ti = 0;
    long long ep_wait = 0, ep_work = 0;
    float* epi = reinterpret_cast<float*>(sptr + S::EPI_OFF);
    int epi_chunk = 0;
    const bool vec_epi = p.out_ws == 1 && p.out_hs == p.Wo && (P & 3) == 0 && (p.out_plane & 3) == 0 &&
                         (reinterpret_cast<uintptr_t>(p.out) & 15u) == 0 && !(p.dbg & 16);
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++ti) {
      int m0, n0, g;
      tile_coords(tile, m0, n0, g);
      const int buf = ti % S::ACC_BUFS, use = ti / S::ACC_BUFS;
      const int m = m0 + r;
      const bool mvalid = m < p.Mtot;
      const int mm = mvalid ? m : 0;
      const int n = mm / P, pix = mm - n * P;
      const int ho = pix / p.Wo, wo = pix - ho * p.Wo;
      float* orow = p.out + ((long long)n * p.Cout_tot + (long long)g * p.Ntot + n0) * p.out_plane +
                    (long long)ho * p.out_hs + (long long)wo * p.out_ws;
      const float* brow = p.bias ? p.bias + (long long)g * p.Ntot + n0 : nullptr;
      const long long e0 = clock64();
      mbar_wait_backoff(bar_tfull + 8 * buf, use & 1, 200);
      const long long e1 = clock64();
      ep_wait += e1 - e0;
      tc_fence_after();
      if (vec_epi) {
        // Contiguous output planes: 4-byte stores (128 B per warp instruction, one per channel) left this kernel
        // bound by the number of store requests in flight on the short-K 1x1 layers (8 000 cycles of stores per
        // tile against 1 500 of MMA).  Transpose each 32-channel chunk through shared memory so that one warp
        // instruction writes 128 consecutive pixels of a channel: 512 B per request, a quarter of the requests.
        const int mv = m0 + lane * 4;                      // this lane's 4 pixels in the store phase
        const bool mv_ok = mv < p.Mtot;                    // P % 4 == 0: the group is valid or not as a whole
        const int nv = (mv_ok ? mv : 0) / P, pv = (mv_ok ? mv : 0) - nv * P;
        float* vbase = p.out + ((long long)nv * p.Cout_tot + (long long)g * p.Ntot + n0) * p.out_plane + pv;
#pragma unroll 1
        for (int c0 = 0; c0 < N_TILE; c0 += 32, ++epi_chunk) {
          if (n0 + c0 >= p.Ntot) break;   // uniform over the four epilogue warps
          float v[32];
          tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(buf * N_TILE + c0), v);
          if (c0 + 32 >= N_TILE || n0 + c0 + 32 >= p.Ntot) {   // accumulator fully read: hand it back before storing
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
          }
          float* E = epi + (epi_chunk & 1) * (32 * 128);
#pragma unroll
          for (int j = 0; j < 32; ++j) E[j * 128 + r] = v[j];           // lanes = consecutive pixels: conflict-free
          asm volatile("bar.sync 1, 128;" ::: "memory");                // the four epilogue warps
          if (!(p.dbg & 4)) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int ch = lg * 8 + q;
              if (mv_ok && n0 + c0 + ch < p.Ntot) {
                float4 t = *reinterpret_cast<const float4*>(E + ch * 128 + lane * 4);
                if (brow) { const float b = __ldg(brow + c0 + ch); t.x += b; t.y += b; t.z += b; t.w += b; }
                *reinterpret_cast<float4*>(vbase + (long long)(c0 + ch) * p.out_plane) = t;
              }
            }
          }
        }
        ep_work += clock64() - e1;
        continue;
      }
#pragma unroll 1
      for (int c0 = 0; c0 < N_TILE; c0 += 32) {
        if (n0 + c0 >= p.Ntot) break;   // warp-uniform
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(buf * N_TILE + c0), v);
        if (mvalid && !(p.dbg & 4)) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (n0 + c0 + j < p.Ntot) {
              float t = v[j];
              if (brow) t += __ldg(brow + c0 + j);
              orow[(long long)(c0 + j) * p.out_plane] = t;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
      ep_work += clock64() - e1;
    }
    if (p.prof && blockIdx.x == 0 && r == 0) { p.prof[16] = ep_wait; p.prof[17] = ep_work; p.prof[18] = ti; }
  }
  __syncthreads();
  if (warp == NPW + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---- host side ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  }
  return fn;
}

static int make_filter_map(CUtensorMap* map, const float* base, int Kp, int rows, int G, int n_tile) {
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return fail(B2C_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[3] = {(cuuint64_t)Kp, (cuuint64_t)rows, (cuuint64_t)G};
  cuuint64_t strides[2] = {(cuuint64_t)Kp * 4, (cuuint64_t)Kp * 4 * (cuuint64_t)rows};
  cuuint32_t box[3] = {32, (cuuint32_t)n_tile, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(B2C_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return B2C_OK;
}

template <int N_TILE, bool SPLIT, int KMODE>
static int launch_fwd_inst(const FwdParams& p, const CUtensorMap& mh, const CUtensorMap& ml, cudaStream_t st) {
  using S = FwdSmem<N_TILE, SPLIT>;
  // per-device attribute: set on every launch (a process may drive several GPUs)
  B2C_CUDA_OK(cudaFuncSetAttribute(igemm_fwd_kernel<N_TILE, SPLIT, KMODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)S::TOTAL));
  const int grid = p.total_tiles < sm_count() ? p.total_tiles : sm_count();
  igemm_fwd_kernel<N_TILE, SPLIT, KMODE><<<grid, FW_THREADS, S::TOTAL, st>>>(p, mh, ml);
  B2C_POST_LAUNCH();
  return B2C_OK;
}

template <int N_TILE>
static int launch_fwd_n(const FwdParams& p, const CUtensorMap& mh, const CUtensorMap& ml, int math, int kmode,
                        cudaStream_t st) {
  if (math == B2C_MATH_FP32)
    return kmode ? launch_fwd_inst<N_TILE, true, 1>(p, mh, ml, st) : launch_fwd_inst<N_TILE, true, 0>(p, mh, ml, st);
  return kmode ? launch_fwd_inst<N_TILE, false, 1>(p, mh, ml, st) : launch_fwd_inst<N_TILE, false, 0>(p, mh, ml, st);
}

// N tile choice, from the per-layer sweeps in profiles/: the producers stage 32 KB of TF32 hi/lo per K block
// into tensor memory regardless of N, so narrow tiles are staging-bound (N = 64: ~2x slower per FLOP than
// N = 128) and N = 256 loses the second accumulator buffer (exposed epilogue).  128 whenever the layer has
// more than 64 output columns, else the smallest tile that covers them.
static int pick_n_tile(int Mtot, int Ntot, int G, int Kp, int math) {
  (void)Mtot; (void)G; (void)Kp; (void)math;
  if (Ntot > 64) return 128;
  if (Ntot > 32) return 64;
  return 32;
}

bool tc_wgrad_supported(const ConvShape& s);
size_t tc_wgrad_workspace(const ConvShape& s);
int launch_conv_tc_wgrad(const ConvShape& s, int math, const float* x, const float* dy, float* dw, void* ws,
                         size_t ws_bytes, cudaStream_t st);

static bool dgrad_as_fwd(const ConvShape& s) { return s.sh == 1 && s.sw == 1; }
static bool dgrad_scatter(const ConvShape& s) { return s.kh == 1 && s.kw == 1 && s.ph == 0 && s.pw == 0 && (s.sh > 1 || s.sw > 1); }

static int padded_k(int K) { return (K + BK - 1) / BK * BK; }

bool tc_conv_supported(const ConvShape& s, int op) {
  const long long Mtot = (long long)s.N * s.Ho * s.Wo;
  if (Mtot > 0x7fffffffLL || (long long)s.N * s.H * s.W > 0x7fffffffLL) return false;
  if ((long long)s.C * s.H * s.W >= 0x7fffffffLL || (long long)s.O * s.Ho * s.Wo >= 0x7fffffffLL) return false;
  if ((s.kh - 1) * s.dh > 0x7fff || (s.kw - 1) * s.dw > 0x7fff) return false;
  auto table_fits = [](int K, int chan) { const int Kp = (K + 31) / 32 * 32; return (chan % 4 == 0 ? Kp / 4 : Kp) <= TAB_ENTRIES; };
  if (op == B2C_OP_FORWARD) return table_fits(s.Kd, s.Cg);
  if (op == B2C_OP_BACKWARD_DATA) return (dgrad_as_fwd(s) || dgrad_scatter(s)) && table_fits(s.Og * s.kh * s.kw, s.Og);
  return tc_wgrad_supported(s);
}

size_t tc_conv_workspace(const ConvShape& s, int op, int math) {
  const size_t np = math == B2C_MATH_FP32 ? 2 : 1;
  if (op == B2C_OP_FORWARD) return sizeof(float) * np * (size_t)s.O * padded_k(s.Kd) + 256;
  if (op == B2C_OP_BACKWARD_DATA) return sizeof(float) * np * (size_t)s.C * padded_k(s.Og * s.kh * s.kw) + 256;
  return tc_wgrad_workspace(s);
}

// The filter operand of `op` in this kernel's GEMM layout, written into `dst` (tc_conv_workspace(s, op, math) bytes).
void tc_conv_prep_entry(const ConvShape& s, int op, int math, const float* w, void* dst, PrepEntry* q) {
  float* wbase = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(dst) + 255) & ~(uintptr_t)255);
  const int chan = op == B2C_OP_FORWARD ? s.Cg : s.Og;
  q->w = w; q->kind = 0; q->G = s.G; q->Og = s.Og; q->Cg = s.Cg; q->taps = s.kh * s.kw;
  q->mode = op == B2C_OP_FORWARD ? 0 : 1;
  q->rows = op == B2C_OP_FORWARD ? s.Og : s.Cg;
  q->K = chan * s.kh * s.kw; q->Kp = padded_k(q->K);
  q->tap_major = (chan % 4 == 0) ? 1 : 0;
  q->flip = (op == B2C_OP_BACKWARD_DATA && dgrad_as_fwd(s)) ? 1 : 0;
  q->kch = chan;
  q->total = (long long)s.G * q->rows * q->Kp;
  q->hi = wbase;
  q->lo = math == B2C_MATH_FP32 ? static_cast<void*>(wbase + q->total) : nullptr;
}

// `prepared`: the filter already in GEMM layout (b2c_conv_prepare_filter) or null = run the prepass here into `ws`
int launch_conv_tc(const ConvShape& s, int op, int math, const float* a, const float* b, const float* bias, float* out,
                   void* ws, size_t ws_bytes, const void* prepared, cudaStream_t st) {
  if (op == B2C_OP_BACKWARD_FILTER) return launch_conv_tc_wgrad(s, math, a, b, out, ws, ws_bytes, st);
  if (!prepared && (ws_bytes < tc_conv_workspace(s, op, math) || !ws)) return fail(B2C_ERR_WORKSPACE, "tcgen05 conv: workspace too small");
  FwdParams p;
  PrepEntry q;
  tc_conv_prep_entry(s, op, math, b, prepared ? const_cast<void*>(prepared) : ws, &q);
  if (op == B2C_OP_FORWARD) {
    p.x = a; p.Cin_tot = s.C; p.H = s.H; p.W = s.W; p.Cg = s.Cg;
    p.kh = s.kh; p.kw = s.kw; p.sh = s.sh; p.sw = s.sw; p.ph = s.ph; p.pw = s.pw; p.dh = s.dh; p.dw = s.dw;
    p.Ho = s.Ho; p.Wo = s.Wo; p.Mtot = s.N * s.Ho * s.Wo;
    p.Ntot = s.Og; p.K = s.Kd;
    p.out = out; p.Cout_tot = s.O; p.out_plane = (long long)s.Ho * s.Wo; p.out_hs = s.Wo; p.out_ws = 1;
    p.bias = bias;
  } else {
    // a = dy [N,O,Ho,Wo], b = w, out = dx [N,C,H,W]
    p.x = a; p.Cin_tot = s.O; p.Cg = s.Og;
    p.Ntot = s.Cg; p.K = s.Og * s.kh * s.kw;
    p.out = out; p.Cout_tot = s.C; p.out_plane = (long long)s.H * s.W; p.bias = nullptr;
    if (dgrad_as_fwd(s)) {
      // stride-1 conv of dY with the flipped filter, pad' = (k-1)*d - p, rows = bottom pixels
      p.H = s.Ho; p.W = s.Wo;
      p.kh = s.kh; p.kw = s.kw; p.sh = 1; p.sw = 1; p.dh = s.dh; p.dw = s.dw;
      p.ph = (s.kh - 1) * s.dh - s.ph; p.pw = (s.kw - 1) * s.dw - s.pw;
      p.Ho = s.H; p.Wo = s.W; p.Mtot = s.N * s.H * s.W;
      p.out_hs = s.W; p.out_ws = 1;
    } else {
      // 1x1, stride > 1, pad 0: rows = top pixels, scatter to bottom[ho*sh][wo*sw]; the rest of dx is 0
      B2C_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)s.N * s.C * s.H * s.W, st));
      p.H = s.Ho; p.W = s.Wo;
      p.kh = 1; p.kw = 1; p.sh = 1; p.sw = 1; p.dh = 1; p.dw = 1; p.ph = 0; p.pw = 0;
      p.Ho = s.Ho; p.Wo = s.Wo; p.Mtot = s.N * s.Ho * s.Wo;
      p.out_hs = s.sh * s.W; p.out_ws = s.sw;
    }
  }
  p.Kp = padded_k(p.K);
  static long long* prof_buf = nullptr;
  static int prof_on = -1;
  if (prof_on < 0) { const char* e = getenv("B2C_PROF"); prof_on = e ? atoi(e) : 0; if (prof_on) cudaMalloc(&prof_buf, 32 * sizeof(long long)); }
  p.prof = prof_on ? prof_buf : nullptr;
  { static int dbg = -1; if (dbg < 0) { const char* e = getenv("B2C_DBG"); dbg = e ? atoi(e) : 0; } p.dbg = dbg; }
  const int kmode = (p.Cg % 4 == 0) ? 1 : 0;
  if (!prepared)
    if (int rc = launch_filter_prep(&q, 1, st)) return rc;

  int n_tile = pick_n_tile(p.Mtot, p.Ntot, s.G, p.Kp, math);
  { static int force = -1; if (force < 0) { const char* e = getenv("B2C_NTILE"); force = e ? atoi(e) : 0; }
    if (force && force / 2 < p.Ntot) n_tile = force; }
  p.m_tiles = (p.Mtot + 127) / 128; p.n_tiles = (p.Ntot + n_tile - 1) / n_tile; p.G = s.G;
  p.total_tiles = p.m_tiles * p.n_tiles * p.G;
  alignas(64) CUtensorMap mh, ml;
  if (int rc = make_filter_map(&mh, static_cast<const float*>(q.hi), p.Kp, q.rows, s.G, n_tile)) return rc;
  if (int rc = make_filter_map(&ml, static_cast<const float*>(q.lo ? q.lo : q.hi), p.Kp, q.rows, s.G, n_tile)) return rc;
  if (prof_on) {
    cudaMemsetAsync(prof_buf, 0, 32 * sizeof(long long), st);
    int rc = n_tile == 128 ? launch_fwd_n<128>(p, mh, ml, math, kmode, st) : n_tile == 64 ? launch_fwd_n<64>(p, mh, ml, math, kmode, st) : launch_fwd_n<32>(p, mh, ml, math, kmode, st);
    long long h[32];
    cudaMemcpy(h, prof_buf, sizeof(h), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[prof] N_TILE=%d nkb=%d tiles=%d | producer(t0): total=%lld wait_empty=%lld store=%lld load=%lld blocks=%lld | mma: total=%lld wait_full=%lld issue=%lld wait_tempty=%lld | epi: wait=%lld work=%lld tiles=%lld\n",
            n_tile, p.Kp / 32, p.total_tiles, h[0], h[1], h[2], h[3], h[4], h[8], h[9], h[10], h[11], h[16], h[17], h[18]);
    return rc;
  }
  switch (n_tile) {
    case 256: return launch_fwd_n<256>(p, mh, ml, math, kmode, st);
    case 128: return launch_fwd_n<128>(p, mh, ml, math, kmode, st);
    case 64: return launch_fwd_n<64>(p, mh, ml, math, kmode, st);
    default: return launch_fwd_n<32>(p, mh, ml, math, kmode, st);
  }
}

TC_DEBUG_EXPORT(debug_mbar_fwd)

}  // namespace b2c
