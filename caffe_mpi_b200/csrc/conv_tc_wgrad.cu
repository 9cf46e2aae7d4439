// conv_tc_wgrad.cu -- tcgen05 implicit-GEMM weight-gradient kernel (sm_100a).
//
// Replaces cudnnConvolutionBackwardFilter beta=1 (reference src/caffe/layers/cudnn_conv_layer.cu:95-99)
// and, semantically, the per-image weight_gpu_gemm loop (base_conv_layer.hpp:148-162):
//     dW[o][k'=(c,i,j)] += sum_{q=(n,ho,wo)} dY[n,o,ho,wo] * X[n,c,ho*s-p+i*d,wo*s-p+j*d]
// GEMM view: M = output channels of one group (tile 128), N = K_dim = (C/g)*kh*kw (tile N_TILE), the
// reduction runs over q = N*Ho*Wo, which is the CONTIGUOUS axis of both operands in NCHW memory, so both
// smem tiles are filled "lanes along K": dY with 128-bit loads, X through the im2col gather.  The
// reduction is split across CTAs (grid.x); partial tiles go to a workspace and the tile's CTAs then add them
// to dW in split order (splitk_fused_reduce; the reference accumulates image by image, also a fixed order).
#include <cuda.h>
#include <stdlib.h>
#include "b2c_common.cuh"
#include "tc_common.cuh"

namespace b2c {
using namespace tc;

constexpr int WG_YW = 8;                        // dY producer warps (warps 0-3 of them run the epilogue afterwards)
constexpr int WG_GW = 16;                       // X-gather producer warps
constexpr int WG_THREADS = (WG_YW + WG_GW + 1) * 32;   // + the MMA warp

struct WgradParams {
  const float* dy;   // [N, O, Ho, Wo]
  const float* x;    // [N, C, H, W]
  int N, C, H, W, O, Cg, Og, Kd;
  int kh, kw, sh, sw, ph, pw, dh, dw, Ho, Wo;
  long long Q;       // N*Ho*Wo
  int kb_per_split;  // k-blocks (of 32 q) per split
  int splits;
  float* out;        // splits == 1: dW (accumulated);  else partials [splits][O*Kd] (overwritten)
  float* grad;       // splits > 1: the gradient blob the fused reduction accumulates into
  unsigned int* counters;   // splits > 1: one zeroed arrival counter per output tile
  long long* prof;   // optional cycle counters from CTA (0,0,0) (B2C_PROF=1)
};

template <int N_TILE, bool SPLIT>
struct WgradSmem {
  static constexpr uint32_t A_BYTES = tile_bytes(128);
  static constexpr uint32_t B_BYTES = tile_bytes(N_TILE);
  static constexpr uint32_t STAGE = (SPLIT ? 2u : 1u) * (A_BYTES + B_BYTES);
  static constexpr uint32_t TABLE = N_TILE * 8u;
  static constexpr int STAGES = (int)((216u * 1024u) / STAGE) > 6 ? 6 : (int)((216u * 1024u) / STAGE);
  static constexpr uint32_t BAR_OFF = STAGES * STAGE;
  static constexpr uint32_t TAB_OFF = BAR_OFF + 256;
  static constexpr uint32_t TOTAL = TAB_OFF + TABLE;
};

// ---- split-K reduction fused into the tile's own CTAs ---------------------------------------------------------------------
// Round 1 ran a second kernel per layer (wgrad_reduce_kernel, 53 launches / 0.58 ms per ResNet-50 step).  Here the `splits` CTAs
// of an output tile meet at a counter in the workspace once their partial tiles are in global memory -- the whole grid is one
// wave by construction (wgrad_plan: never more CTAs than SMs), so every CTA of the tile is resident -- and then each CTA adds
// up its 1/splits slice of the tile over the partials IN SPLIT ORDER and accumulates it into dW: the summation order of every
// element is fixed, so the result is bit-reproducible and equal to what the separate kernel produced.
__device__ __forceinline__ void splitk_fused_reduce(const float* __restrict__ part, float* __restrict__ dw, unsigned int* counter,
                                                    int split, int splits, long long plane, long long tile_base, int ld, int nrows, int ncols) {
  __syncthreads();                                   // this CTA's partial tile is complete (epilogue warps)
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    const long long t0 = clock64();
    unsigned int seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
      if (seen < (unsigned)splits) {
        __nanosleep(64);
        if (clock64() - t0 > tc::MBAR_TIMEOUT_CYCLES) asm volatile("trap;");
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

template <int N_TILE, bool SPLIT, bool X1X1>
__global__ void __launch_bounds__(WG_THREADS, 1)
igemm_wgrad_kernel(const __grid_constant__ WgradParams p) {
  using S = WgradSmem<N_TILE, SPLIT>;
  constexpr int STAGES = S::STAGES;
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_full = sbase + S::BAR_OFF;
  const uint32_t bar_empty = bar_full + 8 * STAGES;
  const uint32_t bar_tmem = bar_empty + 8 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::BAR_OFF + 8 * (2 * STAGES + 1));
  int2* rowtab = reinterpret_cast<int2*>(smem + S::TAB_OFF);   // per B row: {channel offset, hoff | woff<<16}

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int split = blockIdx.x;
  const int n0 = blockIdx.y * N_TILE;                 // k' tile
  const int mtiles = (p.Og + 127) / 128;
  const int g = blockIdx.z / mtiles, m0 = (blockIdx.z % mtiles) * 128;
  const int P = p.Ho * p.Wo;
  const long long HW = (long long)p.H * p.W;
  const long long nkb_total = (p.Q + BK - 1) / BK;
  const long long kb_begin = (long long)split * p.kb_per_split;
  long long kb_end = kb_begin + p.kb_per_split;
  if (kb_end > nkb_total) kb_end = nkb_total;
  const int nkb = (int)(kb_end - kb_begin);           // >= 1 by construction

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, WG_YW + WG_GW); mbar_init(bar_empty + 8 * s, 1); }   // one elected arrive per producer warp
    mbar_init(bar_tmem, 1);
    fence_barrier_init();
  }
  if (warp == WG_YW + WG_GW) tmem_alloc(smem_u32(tmem_slot), N_TILE);
  // row table for the X gather: k' = (c,i,j)
  for (int r = tid; r < N_TILE; r += WG_THREADS) {
    const int kp = n0 + r;
    int2 e = make_int2(-1, 0);
    if (kp < p.Kd) {
      const int j = kp % p.kw, i = (kp / p.kw) % p.kh, c = kp / (p.kw * p.kh);
      e.x = (int)(((long long)g * p.Cg + c) * HW);      // channel offset inside one image (C*H*W < 2^31)
      e.y = (i * p.dh) | ((j * p.dw) << 16);
    }
    rowtab[r] = e;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  constexpr uint32_t LBO_A = tile_lbo(128), LBO_B = tile_lbo(N_TILE);
  auto stage_a_hi = [&](int s) { return sbase + s * S::STAGE; };
  auto stage_a_lo = [&](int s) { return sbase + s * S::STAGE + S::A_BYTES; };
  auto stage_b_hi = [&](int s) { return sbase + s * S::STAGE + (SPLIT ? 2u : 1u) * S::A_BYTES; };
  auto stage_b_lo = [&](int s) { return stage_b_hi(s) + S::B_BYTES; };

  if (warp < WG_YW) {
    // ================= A producer: dY rows (output channels), lanes along q ==========================
    const int kc = tid & 7, rb = tid >> 3;              // rb in [0, 32): rows r*32 + rb, r < 4
    const bool vec_ok = (P % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.dy) & 15u) == 0);
    // loads of k-block kb+1 are issued before the stores of k-block kb (register double buffer)
    // q = n*P + p of this thread's first element; advanced by BK per K block with 32-bit arithmetic only
    // (the 64-bit q / P per block made the load phase ~2000 cycles)
    int cur_n, cur_p;
    {
      const long long q0 = kb_begin * BK + kc * 4;
      cur_n = (int)(q0 / P); cur_p = (int)(q0 - (long long)cur_n * P);
    }
    auto load_block = [&](int /*kb*/, float (&v)[16]) {
      int nn[4], pp[4];
      bool qv[4];
      {
        int n_ = cur_n, p_ = cur_p;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          qv[e] = n_ < p.N;
          nn[e] = n_; pp[e] = p_;
          if (++p_ == P) { p_ = 0; ++n_; }
        }
        cur_p += BK;
        while (cur_p >= P) { cur_p -= P; ++cur_n; }
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int o = m0 + r * 32 + rb;
        float* vv = v + r * 4;
        vv[0] = vv[1] = vv[2] = vv[3] = 0.f;
        if (o < p.Og) {
          const long long ch = (long long)g * p.Og + o;
          if (vec_ok && qv[3] && nn[3] == nn[0]) {
            const float4 t4 = __ldg(reinterpret_cast<const float4*>(p.dy + ((long long)nn[0] * p.O + ch) * P + pp[0]));
            vv[0] = t4.x; vv[1] = t4.y; vv[2] = t4.z; vv[3] = t4.w;
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (qv[e]) vv[e] = __ldg(p.dy + ((long long)nn[e] * p.O + ch) * P + pp[e]);
          }
        }
      }
    };
    long long pa_wait = 0, pa_store = 0;
    const long long pa_t0 = clock64();
    auto store_block = [&](int kb, const float (&v)[16]) {
      const int s = kb % STAGES, it = kb / STAGES;
      const long long w0 = clock64();
      mbar_wait_backoff(bar_empty + 8 * s, (it & 1) ^ 1, 40);
      const long long w1 = clock64();
      pa_wait += w1 - w0;
      const uint32_t a_hi = stage_a_hi(s) + kc * LBO_A, a_lo = stage_a_lo(s) + kc * LBO_A;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int row = r * 32 + rb;
        store_chunk<SPLIT>(a_hi + row * 16, a_lo + row * 16, v[r * 4], v[r * 4 + 1], v[r * 4 + 2], v[r * 4 + 3]);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_full + 8 * s);
      pa_store += clock64() - w1;
    };
    {
      float va[16], vb[16];
      load_block(0, va);
      int kb = 0;
      for (; kb + 2 <= nkb; kb += 2) {
        load_block(kb + 1, vb);
        store_block(kb, va);
        if (kb + 2 < nkb) load_block(kb + 2, va);
        store_block(kb + 1, vb);
      }
      if (kb < nkb) store_block(kb, va);
    }
    const long long pa_t1 = clock64();
    if (warp < 4) {
    // ================= epilogue: TMEM -> registers -> smem transpose -> coalesced row stores ===========
    mbar_wait_backoff(bar_tmem, 0, 100);
    const long long pa_t2 = clock64();
    tc_fence_after();
    // all MMAs have completed, so the pipeline stages are free: each warp uses a private 32 x 33 fp32 pad
    float* tpad = reinterpret_cast<float*>(smem) + warp * (32 * 33);
    float* obase = p.out + (p.splits > 1 ? (long long)split * p.O * p.Kd : 0LL) + ((long long)g * p.Og + m0 + warp * 32) * p.Kd + n0;
    const int rows_valid = p.Og - (m0 + warp * 32);       // rows of this warp's 32 that exist
    const bool accumulate = p.splits == 1;
#pragma unroll 1
    for (int c0 = 0; c0 < N_TILE; c0 += 32) {
      if (n0 + c0 >= p.Kd) break;
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) tpad[lane * 33 + j] = v[j];     // row = lane, conflict-free (stride 33)
      __syncwarp();
      const bool colok = n0 + c0 + lane < p.Kd;
#pragma unroll 4
      for (int r = 0; r < 32; ++r) {
        if (r < rows_valid && colok) {
          float* dst = obase + (long long)r * p.Kd + c0 + lane;     // 32 lanes -> 128 contiguous bytes
          const float t = tpad[r * 33 + lane];
          *dst = accumulate ? *dst + t : t;
        }
      }
      __syncwarp();
    }
    tc_fence_before();
    if (p.prof && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && tid == 0) {
      p.prof[0] = pa_t1 - pa_t0; p.prof[1] = pa_wait; p.prof[2] = pa_store; p.prof[3] = pa_t2 - pa_t1; p.prof[4] = clock64() - pa_t2; p.prof[5] = nkb;
    }
    }
  } else if (warp < WG_YW + WG_GW) {
    // ================= B producer: im2col gather of X, rows k'=(c,i,j), lanes along q ==================
    const int t = tid - WG_YW * 32;
    const int kc = t & 7, rb = t >> 3;                 // rb in [0, 64)
    const bool vec_ok = X1X1 && (P % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.x) & 15u) == 0);
    constexpr int RB = N_TILE >= 64 ? N_TILE / 64 : 1;
    const bool row_active = rb < N_TILE;               // N_TILE = 32: upper half of the threads only arrive
    int cur_n, cur_ho, cur_wo;
    {
      const long long q0 = kb_begin * BK + kc * 4;
      cur_n = (int)(q0 / P);
      const int p0 = (int)(q0 - (long long)cur_n * P);
      cur_ho = p0 / p.Wo; cur_wo = p0 - cur_ho * p.Wo;
    }
    const long long img = (long long)p.C * HW;
    auto load_block = [&](int /*kb*/, float (&v)[RB * 4]) {
      long long base[4];
      int ihb[4], iwb[4];
      bool qv[4];
      {
        int n_ = cur_n, ho = cur_ho, wo = cur_wo;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          qv[e] = n_ < p.N;
          base[e] = n_ * img;
          ihb[e] = ho * p.sh - p.ph; iwb[e] = wo * p.sw - p.pw;
          if (++wo == p.Wo) { wo = 0; if (++ho == p.Ho) { ho = 0; ++n_; } }
        }
        // advance the cursor by BK pixels (32-bit divisions only)
        int w2 = cur_wo + BK;
        const int dh = w2 / p.Wo;
        cur_wo = w2 - dh * p.Wo;
        int h2 = cur_ho + dh;
        const int dn = h2 / p.Ho;
        cur_ho = h2 - dn * p.Ho;
        cur_n += dn;
      }
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        float* vv = v + r * 4;
        vv[0] = vv[1] = vv[2] = vv[3] = 0.f;
        if (!row_active) continue;
        const int2 rt = rowtab[r * 64 + rb];
        if (rt.x >= 0) {
          if (X1X1) {
            // k=1, s=1, p=0: X rows are contiguous in q exactly like dY
            if (vec_ok && qv[3] && base[3] == base[0]) {
              const float4 t4 = __ldg(reinterpret_cast<const float4*>(p.x + base[0] + rt.x + (long long)ihb[0] * p.W + iwb[0]));
              vv[0] = t4.x; vv[1] = t4.y; vv[2] = t4.z; vv[3] = t4.w;
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (qv[e]) vv[e] = __ldg(p.x + base[e] + rt.x + (long long)ihb[e] * p.W + iwb[e]);
            }
          } else {
            const int hoff = rt.y & 0xffff, woff = rt.y >> 16;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int ih = ihb[e] + hoff, iw = iwb[e] + woff;
              if (qv[e] && (unsigned)ih < (unsigned)p.H && (unsigned)iw < (unsigned)p.W)
                v[r * 4 + e] = __ldg(p.x + base[e] + rt.x + ih * p.W + iw);
            }
          }
        }
      }
    };
    long long pb_wait = 0, pb_store = 0;
    const long long pb_t0 = clock64();
    auto store_block = [&](int kb, const float (&v)[RB * 4]) {
      const int s = kb % STAGES, it = kb / STAGES;
      const long long w0 = clock64();
      mbar_wait_backoff(bar_empty + 8 * s, (it & 1) ^ 1, 40);
      const long long w1 = clock64();
      pb_wait += w1 - w0;
      const uint32_t b_hi = stage_b_hi(s) + kc * LBO_B, b_lo = stage_b_lo(s) + kc * LBO_B;
      if (row_active) {
#pragma unroll
        for (int r = 0; r < RB; ++r) {
          const int row = r * 64 + rb;
          store_chunk<SPLIT>(b_hi + row * 16, b_lo + row * 16, v[r * 4], v[r * 4 + 1], v[r * 4 + 2], v[r * 4 + 3]);
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_full + 8 * s);
      pb_store += clock64() - w1;
    };
    {
      float va[RB * 4], vb[RB * 4];
      load_block(0, va);
      int kb = 0;
      for (; kb + 2 <= nkb; kb += 2) {
        load_block(kb + 1, vb);
        store_block(kb, va);
        if (kb + 2 < nkb) load_block(kb + 2, va);
        store_block(kb + 1, vb);
      }
      if (kb < nkb) store_block(kb, va);
    }
    if (p.prof && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && tid == WG_YW * 32) {  // first X-gather thread
      p.prof[8] = clock64() - pb_t0; p.prof[9] = pb_wait; p.prof[10] = pb_store;
    }
  } else {
    // MMA issuer: whole warp converged, the elected lane issues (see elect_one() in tc_common.cuh)
    constexpr uint32_t IDESC = idesc_tf32(128, N_TILE);
    long long mw = 0, mi = 0;
    const long long m_t0 = clock64();
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES, it = kb / STAGES;
      const long long f0 = clock64();
      mbar_wait(bar_full + 8 * s, it & 1);
      const long long f1 = clock64();
      mw += f1 - f0;
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < BK / 8; ++kk) {
          const uint64_t ah = smem_desc(stage_a_hi(s) + 2 * kk * LBO_A, LBO_A, 128);
          const uint64_t bh = smem_desc(stage_b_hi(s) + 2 * kk * LBO_B, LBO_B, 128);
          if (SPLIT) {
            const uint64_t al = smem_desc(stage_a_lo(s) + 2 * kk * LBO_A, LBO_A, 128);
            const uint64_t bl = smem_desc(stage_b_lo(s) + 2 * kk * LBO_B, LBO_B, 128);
            umma_tf32(tmem_base, al, bh, IDESC, (kb | kk) != 0);
            umma_tf32(tmem_base, ah, bl, IDESC, 1);
            umma_tf32(tmem_base, ah, bh, IDESC, 1);
          } else {
            umma_tf32(tmem_base, ah, bh, IDESC, (kb | kk) != 0);
          }
        }
        umma_commit(bar_empty + 8 * s);
        if (kb == nkb - 1) umma_commit(bar_tmem);
      }
      __syncwarp();
      mi += clock64() - f1;
    }
    if (p.prof && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0) { p.prof[16] = clock64() - m_t0; p.prof[17] = mw; p.prof[18] = mi; }
  }
  if (p.splits > 1) {
    const int tile_id = blockIdx.z * gridDim.y + blockIdx.y;
    splitk_fused_reduce(p.out, p.grad, p.counters + tile_id, split, p.splits, (long long)p.O * p.Kd,
                        ((long long)g * p.Og + m0) * p.Kd + n0, p.Kd, min(128, p.Og - m0), min(N_TILE, p.Kd - n0));
  }
  __syncthreads();
  if (warp == WG_YW + WG_GW) {
    tc_fence_after();
    tmem_dealloc(tmem_base, N_TILE);
  }
}

// =====================================================================================================================
// 1x1 / stride 1 / pad 0 layers (24 of ResNet-50's 53): per image both operands are plain row-major matrices with the
// reduction axis q = (ho,wo) contiguous -- dY[n] is [O x P], X[n] is [C x P] -- so neither needs a gather:
//   * one elected thread drops [128 x 32] (dY) and [N_TILE x 32] (X) fp32 boxes into shared memory with TMA
//     (3-D tensor maps {P, rows, N}, SWIZZLE_128B = the canonical K-major UMMA layout; rows past O / C and columns
//     past P are zero-filled by the TMA unit, which is exactly the padding the GEMM needs);
//   * the RAW tile is the TF32 "hi" operand as it stands: kind::tf32 reads the top 19 bits of each fp32 word, i.e.
//     a & 0xffffe000 -- the same truncation split_tf32() applies in software;
//   * 8 converter warps produce the "lo" tiles, lo = a - (a & 0xffffe000), position by position (the copy is
//     layout-agnostic: same swizzled offset in a second buffer), 12 x (LDS.128 + 8 ALU + STS.128) per thread and K block;
//   * the MMA warp issues lo*hi + hi*lo + hi*hi from shared-memory descriptors as in the gather kernel.
// K blocks are enumerated per image (ceil(P/32) chunks, the last one zero padded) and split across CTAs in one wave.
// fp32-equivalent math only: single-pass TF32 mode needs round-to-nearest operands and keeps the gather kernel.
constexpr int WT_CW = 8;                               // converter warps (warps 0-3 also run the epilogue)
constexpr int WT_THREADS = (WT_CW + 2) * 32;           // + TMA warp + MMA warp

struct WgradTmaParams {
  int O, C;              // rows of dY / X per image (G = 1)
  int cpi;               // K chunks of 32 per image = ceil(P / 32)
  long long nkb_total;   // N * cpi
  int kb_per_split, splits;
  float* out;            // splits == 1: dW (accumulated);  else partials [splits][O*C] (overwritten)
  float* grad;           // splits > 1: the gradient blob the fused reduction accumulates into
  unsigned int* counters;   // splits > 1: one zeroed arrival counter per output tile
};

template <int N_TILE>
struct WgradTmaSmem {
  static constexpr uint32_t A_BYTES = 128u * 128u;                     // [128 rows][32 fp32], SW128
  static constexpr uint32_t B_BYTES = (uint32_t)N_TILE * 128u;
  static constexpr uint32_t RAW_BYTES = A_BYTES + B_BYTES;             // one TMA transaction pair
  static constexpr uint32_t STAGE = 2u * RAW_BYTES;                    // raw (= hi) tiles, then lo tiles at +RAW_BYTES
  static constexpr int STAGES_ = (int)((208u * 1024u) / STAGE);
  static constexpr int STAGES = STAGES_ > 6 ? 6 : STAGES_;
  static constexpr uint32_t BAR_OFF = STAGES * STAGE;
  static constexpr uint32_t TOTAL = BAR_OFF + 256 + 1024;              // + 1024-byte alignment slack
};

__device__ __forceinline__ uint64_t wg_desc_sw128(uint32_t addr) {    // K-major, SWIZZLE_128B, 8-row groups 1024 B apart
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void wg_tma_load_3d(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void wg_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ float4 wg_lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

template <int N_TILE>
__global__ void __launch_bounds__(WT_THREADS, 1)
wgrad1x1_tma_kernel(const __grid_constant__ WgradTmaParams p, const __grid_constant__ CUtensorMap map_dy,
                    const __grid_constant__ CUtensorMap map_x) {
  using S = WgradTmaSmem<N_TILE>;
  constexpr int STAGES = S::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;        // SW128 tiles need 1024-byte alignment
  uint8_t* sptr = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t bar_raw = sbase + S::BAR_OFF;                         // TMA bytes have landed
  const uint32_t bar_full = bar_raw + 8 * STAGES;                      // lo tiles written
  const uint32_t bar_empty = bar_full + 8 * STAGES;                    // MMAs of the stage have completed
  const uint32_t bar_tmem = bar_empty + 8 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sptr + S::BAR_OFF + 8 * (3 * STAGES + 1));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int split = blockIdx.x;
  const int n0 = blockIdx.y * N_TILE;                                   // input-channel tile (GEMM N)
  const int m0 = blockIdx.z * 128;                                      // output-channel tile (GEMM M)
  const long long kb_begin = (long long)split * p.kb_per_split;
  long long kb_end = kb_begin + p.kb_per_split;
  if (kb_end > p.nkb_total) kb_end = p.nkb_total;
  const int nkb = (int)(kb_end - kb_begin);                             // >= 1 by construction

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_raw + 8 * s, 1);
      mbar_init(bar_full + 8 * s, WT_CW);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_tmem, 1);
    fence_barrier_init();
  }
  if (warp == WT_CW + 1) tmem_alloc(smem_u32(tmem_slot), N_TILE);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto stage_a = [&](int s) { return sbase + (uint32_t)s * S::STAGE; };            // raw dY tile (hi operand)
  auto stage_b = [&](int s) { return sbase + (uint32_t)s * S::STAGE + S::A_BYTES; };  // raw X tile

  if (warp < WT_CW) {
    // ================= converters: lo = a - trunc_tf32(a), same offset in the lo half of the stage =============
    constexpr int UNITS = (int)(S::RAW_BYTES / 16u);                   // 16-byte units per stage, multiple of 256
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES, it = kb / STAGES;
      mbar_wait_backoff(bar_raw + 8 * s, it & 1, 20);
      const uint32_t src = stage_a(s);
#pragma unroll 4
      for (int u = tid; u < UNITS; u += WT_CW * 32) {
        const float4 v = wg_lds128(src + (uint32_t)u * 16u);
        float h, l0, l1, l2, l3;
        split_tf32(v.x, h, l0); split_tf32(v.y, h, l1); split_tf32(v.z, h, l2); split_tf32(v.w, h, l3);
        sts128(src + S::RAW_BYTES + (uint32_t)u * 16u, l0, l1, l2, l3);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_full + 8 * s);
    }
    if (warp < 4) {
      // ================= epilogue: TMEM -> registers -> smem transpose -> coalesced row stores ==================
      mbar_wait_backoff(bar_tmem, 0, 100);
      tc_fence_after();
      float* tpad = reinterpret_cast<float*>(sptr) + warp * (32 * 33);   // all MMAs done: the stages are free
      float* obase = p.out + (p.splits > 1 ? (long long)split * p.O * p.C : 0LL) + (long long)(m0 + warp * 32) * p.C + n0;
      const int rows_valid = p.O - (m0 + warp * 32);
      const bool accumulate = p.splits == 1;
#pragma unroll 1
      for (int c0 = 0; c0 < N_TILE; c0 += 32) {
        if (n0 + c0 >= p.C) break;
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) tpad[lane * 33 + j] = v[j];
        __syncwarp();
        const bool colok = n0 + c0 + lane < p.C;
#pragma unroll 4
        for (int r = 0; r < 32; ++r) {
          if (r < rows_valid && colok) {
            float* dst = obase + (long long)r * p.C + c0 + lane;
            const float t = tpad[r * 33 + lane];
            *dst = accumulate ? *dst + t : t;
          }
        }
        __syncwarp();
      }
      tc_fence_before();
    }
  } else if (warp == WT_CW) {
    // ================= TMA producer (whole warp converged, one elected lane issues) ============================
    int n = (int)(kb_begin / p.cpi), j = (int)(kb_begin - (long long)n * p.cpi);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES, it = kb / STAGES;
      mbar_wait_backoff(bar_empty + 8 * s, (it & 1) ^ 1, 20);
      if (elect_one()) {
        wg_arrive_expect_tx(bar_raw + 8 * s, S::RAW_BYTES);
        wg_tma_load_3d(stage_a(s), &map_dy, bar_raw + 8 * s, j * 32, m0, n);
        wg_tma_load_3d(stage_b(s), &map_x, bar_raw + 8 * s, j * 32, n0, n);
      }
      __syncwarp();
      if (++j == p.cpi) { j = 0; ++n; }
    }
  } else {
    // ================= MMA issuer (whole warp converged, elect.sync) ============================================
    constexpr uint32_t IDESC = idesc_tf32(128, N_TILE);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES, it = kb / STAGES;
      mbar_wait(bar_full + 8 * s, it & 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < BK / 8; ++kk) {
          const uint64_t ah = wg_desc_sw128(stage_a(s) + kk * 32);
          const uint64_t bh = wg_desc_sw128(stage_b(s) + kk * 32);
          const uint64_t al = wg_desc_sw128(stage_a(s) + S::RAW_BYTES + kk * 32);
          const uint64_t bl = wg_desc_sw128(stage_b(s) + S::RAW_BYTES + kk * 32);
          umma_tf32(tmem_base, al, bh, IDESC, (kb | kk) != 0);
          umma_tf32(tmem_base, ah, bl, IDESC, 1);
          umma_tf32(tmem_base, ah, bh, IDESC, 1);
        }
        umma_commit(bar_empty + 8 * s);
        if (kb == nkb - 1) umma_commit(bar_tmem);
      }
      __syncwarp();
    }
  }
  if (p.splits > 1) {
    const int tile_id = blockIdx.z * gridDim.y + blockIdx.y;
    splitk_fused_reduce(p.out, p.grad, p.counters + tile_id, split, p.splits, (long long)p.O * p.C, (long long)m0 * p.C + n0, p.C,
                        min(128, p.O - m0), min(N_TILE, p.C - n0));
  }
  __syncthreads();
  if (warp == WT_CW + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, N_TILE);
  }
}

struct WgradPlan { int n_tile, splits, kb_per_split; };

static WgradPlan wgrad_plan(const ConvShape& s) {
  WgradPlan pl;
  pl.n_tile = s.Kd > 128 ? 256 : s.Kd > 64 ? 128 : s.Kd > 32 ? 64 : 32;
  const long long Q = (long long)s.N * s.Ho * s.Wo;
  const long long nkb = (Q + BK - 1) / BK;
  const long long mn = (long long)((s.Og + 127) / 128) * ((s.Kd + pl.n_tile - 1) / pl.n_tile) * s.G;
  // one wave: every CTA pays a fixed prologue + epilogue (partial tile write), so fewer, longer CTAs win;
  // never exceed the SM count (a partial second wave costs a full CTA time)
  long long splits = sm_count() / mn;
  const long long max_splits = nkb / 8 > 0 ? nkb / 8 : 1;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  long long per = (nkb + splits - 1) / splits;
  splits = (nkb + per - 1) / per;
  pl.splits = (int)splits;
  pl.kb_per_split = (int)per;
  return pl;
}

bool tc_wgrad_supported(const ConvShape& s) {
  if ((long long)s.C * s.H * s.W >= 0x7fffffffLL) return false;
  if ((s.kh - 1) * s.dh > 0xffff || (s.kw - 1) * s.dw > 0x7fff) return false;
  const long long nkb = ((long long)s.N * s.Ho * s.Wo + BK - 1) / BK;
  return nkb < 0x7fffffffLL;
}

static WgradPlan wgrad_tma_plan(const ConvShape& s, int* cpi_out);
// the TMA-staged bf16x3 kernel (conv_tc_wgrad_stg.cu): the dense 1x1 problem a compacted strided layer turns into goes there when it can
bool tc_wgrad_stg_supported(const ConvShape&);
size_t tc_wgrad_stg_workspace(const ConvShape&);
int launch_conv_tc_wgrad_stg(const ConvShape&, const float* x, const float* dy, float* dw, void* ws, size_t ws_bytes, cudaStream_t);
static bool wgrad_tma_shape_ok(const ConvShape& s);
static bool wgrad_compact_shape_ok(const ConvShape& s);
static ConvShape wgrad_compact_dense_shape(const ConvShape& s);
constexpr size_t WG_COUNTER_BYTES = 1024;      // arrival counters of the fused split-K reduction: one u32 per output tile (<= #SMs)
size_t tc_wgrad_workspace(const ConvShape& s) {
  const WgradPlan pl = wgrad_plan(s);
  size_t need = pl.splits > 1 ? WG_COUNTER_BYTES + sizeof(float) * (size_t)pl.splits * s.O * s.Kd : 0;
  if (wgrad_tma_shape_ok(s)) {                       // the TMA path of 1x1 layers plans its own split count
    const WgradPlan pt = wgrad_tma_plan(s, nullptr);
    const size_t nt = pt.splits > 1 ? WG_COUNTER_BYTES + sizeof(float) * (size_t)pt.splits * s.O * s.C : 0;
    if (nt > need) need = nt;
  }
  if (wgrad_compact_shape_ok(s)) {                   // strided 1x1: compacted input + the TMA path's partials
    const ConvShape d = wgrad_compact_dense_shape(s);
    const WgradPlan pt = wgrad_tma_plan(d, nullptr);
    size_t part = pt.splits > 1 ? WG_COUNTER_BYTES + sizeof(float) * (size_t)pt.splits * s.O * s.C : 0;
    if (tc_wgrad_stg_supported(d)) part = tc_wgrad_stg_workspace(d);
    const size_t nt = part + 256 + sizeof(float) * (size_t)s.N * s.C * s.Ho * s.Wo;
    if (nt > need) need = nt;
  }
  return need;
}

// ---- TMA path for 1x1 layers: eligibility, plan, launch -----------------------------------------------------------
static bool wgrad_tma_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("B2C_WGRAD_TMA"); on = e ? atoi(e) : 1; }
  return on != 0;
}
static bool wgrad_tma_shape_ok(const ConvShape& s) {
  const long long P = (long long)s.Ho * s.Wo;
  return wgrad_tma_enabled() && s.is_1x1 && s.G == 1 && P % 4 == 0 && P >= 32 && s.H == s.Ho && s.W == s.Wo;
}
static WgradPlan wgrad_tma_plan(const ConvShape& s, int* cpi_out) {
  WgradPlan pl;
  pl.n_tile = s.C > 128 ? 256 : s.C > 64 ? 128 : s.C > 32 ? 64 : 32;
  const int cpi = (s.Ho * s.Wo + BK - 1) / BK;
  const long long nkb = (long long)s.N * cpi;
  const long long mn = (long long)((s.O + 127) / 128) * ((s.C + pl.n_tile - 1) / pl.n_tile);
  long long splits = sm_count() / mn;
  const long long max_splits = nkb / 8 > 0 ? nkb / 8 : 1;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  const long long per = (nkb + splits - 1) / splits;
  pl.splits = (int)((nkb + per - 1) / per);
  pl.kb_per_split = (int)per;
  if (cpi_out) *cpi_out = cpi;
  return pl;
}
typedef CUresult (*WgEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static WgEncodeTiledFn wg_encode_tiled() {
  static WgEncodeTiledFn fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<WgEncodeTiledFn>(f);
  }
  return fn;
}
// [N][rows][P] fp32, box {32, box_rows, 1}, SWIZZLE_128B, out-of-bounds elements read as zero
static int wg_make_map(CUtensorMap* map, const float* base, long long P, int rows, int N, int box_rows) {
  WgEncodeTiledFn enc = wg_encode_tiled();
  if (!enc) return fail(B2C_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[3] = {(cuuint64_t)P, (cuuint64_t)rows, (cuuint64_t)N};
  cuuint64_t strides[2] = {(cuuint64_t)P * 4, (cuuint64_t)P * 4 * (cuuint64_t)rows};
  cuuint32_t box[3] = {32, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(B2C_ERR_CUDA, "cuTensorMapEncodeTiled (wgrad) failed (%d)", (int)r);
  return B2C_OK;
}
template <int N_TILE>
static int launch_wgrad_tma_inst(const WgradTmaParams& p, const CUtensorMap& mdy, const CUtensorMap& mx, cudaStream_t st) {
  using S = WgradTmaSmem<N_TILE>;
  // per-device attribute: set on every launch (a process may drive several GPUs)
  B2C_CUDA_OK(cudaFuncSetAttribute(wgrad1x1_tma_kernel<N_TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::TOTAL));
  dim3 grid(p.splits, (p.C + N_TILE - 1) / N_TILE, (p.O + 127) / 128);
  wgrad1x1_tma_kernel<N_TILE><<<grid, WT_THREADS, S::TOTAL, st>>>(p, mdy, mx);
  B2C_POST_LAUNCH();
  return B2C_OK;
}
static int launch_conv_tc_wgrad_tma(const ConvShape& s, const float* x, const float* dy, float* dw, void* ws, size_t ws_bytes,
                                    cudaStream_t st) {
  int cpi = 0;
  const WgradPlan pl = wgrad_tma_plan(s, &cpi);
  const long long P = (long long)s.Ho * s.Wo;
  WgradTmaParams p;
  p.O = s.O; p.C = s.C; p.cpi = cpi; p.nkb_total = (long long)s.N * cpi;
  p.kb_per_split = pl.kb_per_split; p.splits = pl.splits;
  const size_t need = pl.splits > 1 ? WG_COUNTER_BYTES + sizeof(float) * (size_t)pl.splits * s.O * s.C : 0;
  if (need && (!ws || ws_bytes < need)) return fail(B2C_ERR_WORKSPACE, "wgrad (tma): workspace too small");
  p.out = pl.splits > 1 ? reinterpret_cast<float*>(static_cast<char*>(ws) + WG_COUNTER_BYTES) : dw;
  p.grad = dw;
  p.counters = static_cast<unsigned int*>(ws);
  if (pl.splits > 1) B2C_CUDA_OK(cudaMemsetAsync(ws, 0, WG_COUNTER_BYTES, st));
  alignas(64) CUtensorMap mdy, mx;
  int rc = wg_make_map(&mdy, dy, P, s.O, s.N, 128);
  if (rc) return rc;
  rc = wg_make_map(&mx, x, P, s.C, s.N, pl.n_tile);
  if (rc) return rc;
  switch (pl.n_tile) {
    case 256: rc = launch_wgrad_tma_inst<256>(p, mdy, mx, st); break;
    case 128: rc = launch_wgrad_tma_inst<128>(p, mdy, mx, st); break;
    case 64: rc = launch_wgrad_tma_inst<64>(p, mdy, mx, st); break;
    default: rc = launch_wgrad_tma_inst<32>(p, mdy, mx, st); break;
  }
  if (rc) return rc;
  return B2C_OK;
}

// ---- strided 1x1 layers (pad 0, stride > 1): subsample, then the TMA path ------------------------------------------
// On by default for layers with more than 128 output channels (B2C_WGRAD_COMPACT=0 switches it off, =2 forces it for every
// eligible layer).  Measured on ResNet-50's strided projections at N = 64 (profiles/r02_c1_sweep_compact.txt): C256->O512@56
// 183 -> 136 us, C512->O1024@28 196 -> 124 us, C512->O256@28 89 -> 72 us, but C256->O128@56 89 -> 97 us (the extra pass over X
// costs more than the gather it replaces when the GEMM is that small).  dW[o][c] = sum_q dY[o][q] * X[c][ho*sh][wo*sw]: the gather kernel reads X with 8-byte-strided
// 4-byte loads (36 TFLOP/s on ResNet-50's six such layers, 0.92 ms per step against 0.17 ms of MMA).  Copying the
// sampled pixels into a dense [N][C][Ho*Wo] buffer first is one HBM-bound pass over half of X's rows, after which the
// layer is an ordinary 1x1 / stride 1 weight gradient for the TMA-fed kernel.
static int wgrad_compact_mode() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("B2C_WGRAD_COMPACT"); on = e ? atoi(e) : 1; }
  return on;
}
static bool wgrad_compact_shape_ok(const ConvShape& s) {
  const long long P = (long long)s.Ho * s.Wo;
  const int mode = wgrad_compact_mode();
  if (!(mode != 0 && wgrad_tma_enabled() && s.kh == 1 && s.kw == 1 && s.ph == 0 && s.pw == 0 && (s.sh > 1 || s.sw > 1) && s.G == 1))
    return false;
  // with the staged kernel behind it the extra pass over X pays for every layer it takes; with the 3xTF32 TMA kernel only for
  // the wide ones (measurements above).  7x7 outputs (P = 49) exist only on the staged kernel (plane mode).
  ConvShape d = s;
  d.H = s.Ho; d.W = s.Wo; d.sh = d.sw = 1; d.is_1x1 = true;
  const bool staged = tc_wgrad_stg_supported(d);
  if (!(P % 4 == 0 && P >= 32)) return staged && P == 49;
  return mode == 2 || s.O > 128 || staged;
}
static ConvShape wgrad_compact_dense_shape(const ConvShape& s) {
  ConvShape d = s;
  d.H = s.Ho; d.W = s.Wo; d.sh = d.sw = 1; d.is_1x1 = true;
  return d;
}
__global__ void __launch_bounds__(256)
subsample_kernel(const float* __restrict__ x, float* __restrict__ y, long long planes, int H, int W, int Ho, int Wo, int sh, int sw) {
  const long long total = planes * Ho * Wo;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int wo = (int)(i % Wo);
    const long long t = i / Wo;
    const int ho = (int)(t % Ho);
    const long long pl = t / Ho;
    y[i] = __ldg(x + (pl * H + (long long)ho * sh) * W + (long long)wo * sw);
  }
}

template <int N_TILE, bool SPLIT, bool X1X1>
static int launch_wgrad_inst(const WgradParams& p, int G, cudaStream_t st) {
  using S = WgradSmem<N_TILE, SPLIT>;
  // per-device attribute: set on every launch (a process may drive several GPUs)
  B2C_CUDA_OK(cudaFuncSetAttribute(igemm_wgrad_kernel<N_TILE, SPLIT, X1X1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)S::TOTAL));
  dim3 grid(p.splits, (p.Kd + N_TILE - 1) / N_TILE, G * ((p.Og + 127) / 128));
  igemm_wgrad_kernel<N_TILE, SPLIT, X1X1><<<grid, WG_THREADS, S::TOTAL, st>>>(p);
  B2C_POST_LAUNCH();
  return B2C_OK;
}

template <int N_TILE>
static int launch_wgrad_n(const WgradParams& p, int G, int math, bool x1, cudaStream_t st) {
  if (math == B2C_MATH_FP32)
    return x1 ? launch_wgrad_inst<N_TILE, true, true>(p, G, st) : launch_wgrad_inst<N_TILE, true, false>(p, G, st);
  return x1 ? launch_wgrad_inst<N_TILE, false, true>(p, G, st) : launch_wgrad_inst<N_TILE, false, false>(p, G, st);
}

int launch_conv_tc_wgrad(const ConvShape& s, int math, const float* x, const float* dy, float* dw, void* ws,
                         size_t ws_bytes, cudaStream_t st) {
  if (math == B2C_MATH_FP32 && wgrad_tma_shape_ok(s) && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy)) & 15u) == 0)
    return launch_conv_tc_wgrad_tma(s, x, dy, dw, ws, ws_bytes, st);
  if (math == B2C_MATH_FP32 && wgrad_compact_shape_ok(s) && (reinterpret_cast<uintptr_t>(dy) & 15u) == 0 && ws) {
    const ConvShape d = wgrad_compact_dense_shape(s);
    const WgradPlan pt = wgrad_tma_plan(d, nullptr);
    const bool staged = tc_wgrad_stg_supported(d);
    const size_t part = staged ? tc_wgrad_stg_workspace(d) : pt.splits > 1 ? WG_COUNTER_BYTES + sizeof(float) * (size_t)pt.splits * s.O * s.C : 0;
    const size_t xoff = (part + 255) & ~(size_t)255;
    const size_t xbytes = sizeof(float) * (size_t)s.N * s.C * s.Ho * s.Wo;
    if (ws_bytes < xoff + xbytes) return fail(B2C_ERR_WORKSPACE, "wgrad (compact): workspace too small");
    float* xc = reinterpret_cast<float*>(static_cast<char*>(ws) + xoff);       // cudaMalloc'ed workspace: 256-byte aligned base
    if ((reinterpret_cast<uintptr_t>(xc) & 15u) == 0) {
      const long long planes = (long long)s.N * s.C;
      subsample_kernel<<<grid_for((size_t)(planes * s.Ho * s.Wo), 256), 256, 0, st>>>(x, xc, planes, s.H, s.W, s.Ho, s.Wo, s.sh, s.sw);
      B2C_POST_LAUNCH();
      if (staged) return launch_conv_tc_wgrad_stg(d, xc, dy, dw, ws, part, st);
      return launch_conv_tc_wgrad_tma(d, xc, dy, dw, ws, part, st);
    }
  }
  const WgradPlan pl = wgrad_plan(s);
  WgradParams p;
  p.dy = dy; p.x = x;
  p.N = s.N; p.C = s.C; p.H = s.H; p.W = s.W; p.O = s.O; p.Cg = s.Cg; p.Og = s.Og; p.Kd = s.Kd;
  p.kh = s.kh; p.kw = s.kw; p.sh = s.sh; p.sw = s.sw; p.ph = s.ph; p.pw = s.pw; p.dh = s.dh; p.dw = s.dw;
  p.Ho = s.Ho; p.Wo = s.Wo;
  p.Q = (long long)s.N * s.Ho * s.Wo;
  p.kb_per_split = pl.kb_per_split; p.splits = pl.splits;
  static long long* prof_buf = nullptr;
  static int prof_on = -1;
  if (prof_on < 0) { const char* e = getenv("B2C_PROF"); prof_on = e ? atoi(e) : 0; if (prof_on) cudaMalloc(&prof_buf, 32 * sizeof(long long)); }
  p.prof = prof_on ? prof_buf : nullptr;
  if (prof_on) cudaMemsetAsync(prof_buf, 0, 32 * sizeof(long long), st);
  const size_t need = pl.splits > 1 ? WG_COUNTER_BYTES + sizeof(float) * (size_t)pl.splits * s.O * s.Kd : 0;
  if (need && (!ws || ws_bytes < need)) return fail(B2C_ERR_WORKSPACE, "wgrad: workspace too small");
  p.out = pl.splits > 1 ? reinterpret_cast<float*>(static_cast<char*>(ws) + WG_COUNTER_BYTES) : dw;
  p.grad = dw;
  p.counters = static_cast<unsigned int*>(ws);
  if (pl.splits > 1) B2C_CUDA_OK(cudaMemsetAsync(ws, 0, WG_COUNTER_BYTES, st));
  int rc;
  switch (pl.n_tile) {
    case 256: rc = launch_wgrad_n<256>(p, s.G, math, s.is_1x1, st); break;
    case 128: rc = launch_wgrad_n<128>(p, s.G, math, s.is_1x1, st); break;
    case 64: rc = launch_wgrad_n<64>(p, s.G, math, s.is_1x1, st); break;
    default: rc = launch_wgrad_n<32>(p, s.G, math, s.is_1x1, st); break;
  }
  if (rc) return rc;
  if (prof_on) {
    long long h[32];
    cudaMemcpy(h, prof_buf, sizeof(h), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[wprof] N_TILE=%d splits=%d nkb=%lld | dY-prod: loop=%lld wait_empty=%lld store=%lld wait_tmem=%lld epilogue=%lld | X-prod: loop=%lld wait_empty=%lld store=%lld | mma: total=%lld wait_full=%lld issue=%lld\n",
            pl.n_tile, pl.splits, h[5], h[0], h[1], h[2], h[3], h[4], h[8], h[9], h[10], h[16], h[17], h[18]);
  }
  return B2C_OK;
}

TC_DEBUG_EXPORT(debug_mbar_wgrad)

}  // namespace b2c
