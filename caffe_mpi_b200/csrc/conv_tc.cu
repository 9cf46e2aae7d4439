// conv_tc.cu -- tcgen05 (5th-gen tensor core) implicit-GEMM convolution kernels for sm_100a.
//
// Replaces cudnnConvolutionForward / BackwardData / BackwardFilter as called by the reference's
// CuDNNConvolutionLayer (src/caffe/layers/cudnn_conv_layer.cu:25-29,118-123,95-99): whole batch per
// launch, NCHW fp32 in and out, y / dx overwritten, dw accumulated, no col buffer in HBM.
//
// GEMM view (SURVEY.md 8a rows a4/a7/a8 with the batch folded into the pixel dimension):
//   forward : D[q=(n,ho,wo)][o]      = sum_{k=(c,i,j)} X[n,c,ho*s-p+i*d,wo*s-p+j*d] * W[o][k]
//   dgrad   : the same kernel run on dY with the transposed / flipped filter (stride 1), or on the
//             output grid with a strided scatter epilogue (1x1, stride > 1)
//   wgrad   : D[o][k'=(c,i,j)]      += sum_{q=(n,ho,wo)} dY[n,o,q] * X[n,c,...]      (split over q)
//
// Kernel anatomy (one 128 x N_TILE output tile per CTA, 288 threads):
//   warps 0-3  A producers: gather 128 rows x 32 K of the activation operand straight from NCHW global
//              memory (coalesced along W), convert to TF32 (hi [+ lo]) in registers, 128-bit st.shared
//              into the canonical K-major UMMA layout; afterwards they are the epilogue warps
//              (tcgen05.ld TMEM -> registers -> coalesced NCHW stores, bias fused).
//   warps 4-7  B producers: weight / K-contiguous operand, 128-bit global loads, same conversion.
//   warp  8    allocates TMEM, one elected lane issues tcgen05.mma.kind::tf32 (M=128, N=N_TILE, K=8)
//              against smem descriptors, tcgen05.commit releases pipeline stages through mbarriers.
// The activation operand cannot be staged by TMA from NCHW: the implicit-GEMM K index (c,i,j) is not
// a unit-stride axis of the tensor and the 7x7 maps have 196-byte channel strides (TMA needs 16-byte
// multiples), see DESIGN.md.  fp32 mode issues 3 TF32 MMAs per K step (lo*hi, hi*lo, hi*hi).
#include "b2c_common.cuh"
#include "tc_common.cuh"

namespace b2c {
using namespace tc;

constexpr int TC_THREADS = 288;

struct FwdLikeParams {
  // A: activations [Nimg, Cin_tot, H, W]; group g reads channels [g*Cg, (g+1)*Cg)
  const float* x;
  int Cin_tot, H, W, Cg;
  int kh, kw, sh, sw, ph, pw, dh, dw;
  int Ho, Wo;        // pixel grid of the GEMM rows: m = (n*Ho + ho)*Wo + wo
  int Mtot;          // Nimg*Ho*Wo
  // B: [G][Ntot][K] row-major, K contiguous (K = Cg*kh*kw)
  const float* w;
  int Ntot, K;
  // out[(n*Cout_tot + g*Ntot + col)*out_plane + ho*out_hs + wo*out_ws]
  float* out;
  int Cout_tot, out_hs, out_ws;
  long long out_plane;
  const float* bias;  // [G*Ntot] or null
};

template <int N_TILE, bool SPLIT>
struct FwdLikeSmem {
  static constexpr uint32_t A_BYTES = tile_bytes(128);
  static constexpr uint32_t B_BYTES = tile_bytes(N_TILE);
  static constexpr uint32_t STAGE = (SPLIT ? 2u : 1u) * (A_BYTES + B_BYTES);
  static constexpr int STAGES = (int)((220u * 1024u) / STAGE) > 6 ? 6 : (int)((220u * 1024u) / STAGE);
  static constexpr uint32_t BAR_OFF = STAGES * STAGE;
  static constexpr uint32_t TOTAL = BAR_OFF + 256;
};

template <int N_TILE, bool SPLIT, bool K1X1>
__global__ void __launch_bounds__(TC_THREADS, 1)
igemm_fwdlike_kernel(const __grid_constant__ FwdLikeParams p) {
  using S = FwdLikeSmem<N_TILE, SPLIT>;
  constexpr int STAGES = S::STAGES;
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_full = sbase + S::BAR_OFF;            // STAGES x 8 B
  const uint32_t bar_empty = bar_full + 8 * STAGES;        // STAGES x 8 B
  const uint32_t bar_tmem = bar_empty + 8 * STAGES;        // 8 B
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::BAR_OFF + 8 * (2 * STAGES + 1));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * 128, n0 = blockIdx.y * N_TILE, g = blockIdx.z;
  const int nkb = (p.K + BK - 1) / BK;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 256);   // every producer thread arrives
      mbar_init(bar_empty + 8 * s, 1);    // one tcgen05.commit
    }
    mbar_init(bar_tmem, 1);
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc(smem_u32(tmem_slot), N_TILE);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  constexpr uint32_t LBO_A = tile_lbo(128), LBO_B = tile_lbo(N_TILE);
  auto stage_a_hi = [&](int s) { return sbase + s * S::STAGE; };
  auto stage_a_lo = [&](int s) { return sbase + s * S::STAGE + S::A_BYTES; };
  auto stage_b_hi = [&](int s) { return sbase + s * S::STAGE + (SPLIT ? 2u : 1u) * S::A_BYTES; };
  auto stage_b_lo = [&](int s) { return stage_b_hi(s) + S::B_BYTES; };

  if (warp < 4) {
    // ================= A producer: one GEMM row (output pixel) per thread =========================
    const int m = m0 + tid;
    const bool mvalid = m < p.Mtot;
    const int P = p.Ho * p.Wo;
    const int mm = mvalid ? m : 0;
    const int n = mm / P, pix = mm - n * P;
    const int ho = pix / p.Wo, wo = pix - ho * p.Wo;
    const int ih0 = ho * p.sh - p.ph, iw0 = wo * p.sw - p.pw;
    const long long HW = (long long)p.H * p.W;
    const float* xrow = p.x + ((long long)n * p.Cin_tot + (long long)g * p.Cg) * HW + (long long)ih0 * p.W + iw0;
    const bool row_inb = mvalid && (unsigned)ih0 < (unsigned)p.H && (unsigned)iw0 < (unsigned)p.W;  // used when K1X1
    // incremental decode of k -> (c,i,j): koff = c*HW + i*dh*W + j*dw, hoff = i*dh, woff = j*dw
    int ki = 0, kj = 0, hoff = 0, woff = 0, k = 0;
    long long koff = 0;
    const uint32_t row_off = (uint32_t)tid * 16u;
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES, it = kb / STAGES;
      mbar_wait(bar_empty + 8 * s, (it & 1) ^ 1);
      const uint32_t a_hi = stage_a_hi(s) + row_off, a_lo = stage_a_lo(s) + row_off;
#pragma unroll
      for (int kc = 0; kc < KCHUNKS; ++kc) {
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          bool ok;
          if (K1X1) {
            ok = row_inb && k < p.K;
          } else {
            ok = mvalid && k < p.K && (unsigned)(ih0 + hoff) < (unsigned)p.H && (unsigned)(iw0 + woff) < (unsigned)p.W;
          }
          v[e] = ok ? __ldg(xrow + koff) : 0.0f;
          ++k;
          if (K1X1) {
            koff += HW;
          } else {
            ++kj; woff += p.dw; koff += p.dw;
            if (kj == p.kw) {
              kj = 0; koff -= woff; woff = 0;
              ++ki; hoff += p.dh; koff += (long long)p.dh * p.W;
              if (ki == p.kh) { ki = 0; koff -= (long long)hoff * p.W; hoff = 0; koff += HW; }
            }
          }
        }
        store_chunk<SPLIT>(a_hi + kc * LBO_A, a_lo + kc * LBO_A, v[0], v[1], v[2], v[3]);
      }
      fence_proxy_async();
      mbar_arrive(bar_full + 8 * s);
    }
    // ================= epilogue: TMEM -> registers -> NCHW global ==================================
    mbar_wait(bar_tmem, 0);
    tc_fence_after();
    float* orow = p.out + ((long long)n * p.Cout_tot + (long long)g * p.Ntot + n0) * p.out_plane +
                  (long long)ho * p.out_hs + (long long)wo * p.out_ws;
    const float* brow = p.bias ? p.bias + (long long)g * p.Ntot + n0 : nullptr;
#pragma unroll 1
    for (int c0 = 0; c0 < N_TILE; c0 += 32) {
      if (n0 + c0 >= p.Ntot) break;   // warp-uniform
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
      if (mvalid) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (n0 + c0 + j < p.Ntot) {
            float r = v[j];
            if (brow) r += __ldg(brow + c0 + j);
            orow[(long long)(c0 + j) * p.out_plane] = r;
          }
        }
      }
    }
    tc_fence_before();
  } else if (warp < 8) {
    // ================= B producer: lanes along K, 4 rows x 8 chunks per warp pass ==================
    const int t = tid - 128;
    const float* wg = p.w + (long long)g * p.Ntot * p.K;
    const bool vec_ok = (p.K % 4 == 0) && ((reinterpret_cast<uintptr_t>(wg) & 15u) == 0);
    const int kc = t & 7;
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES, it = kb / STAGES;
      mbar_wait(bar_empty + 8 * s, (it & 1) ^ 1);
      const uint32_t b_hi = stage_b_hi(s) + kc * LBO_B, b_lo = stage_b_lo(s) + kc * LBO_B;
      const int k = kb * BK + kc * 4;
#pragma unroll
      for (int r = 0; r < N_TILE / 16; ++r) {
        const int row = r * 16 + (t >> 3);
        const int col = n0 + row;
        float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
        if (col < p.Ntot) {
          const float* src = wg + (long long)col * p.K + k;
          if (vec_ok && k + 3 < p.K) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(src));
            v0 = q.x; v1 = q.y; v2 = q.z; v3 = q.w;
          } else {
            if (k < p.K) v0 = __ldg(src);
            if (k + 1 < p.K) v1 = __ldg(src + 1);
            if (k + 2 < p.K) v2 = __ldg(src + 2);
            if (k + 3 < p.K) v3 = __ldg(src + 3);
          }
        }
        store_chunk<SPLIT>(b_hi + row * 16, b_lo + row * 16, v0, v1, v2, v3);
      }
      fence_proxy_async();
      mbar_arrive(bar_full + 8 * s);
    }
  } else {
    // ================= MMA issuer ====================================================================
    if (lane == 0) {
      constexpr uint32_t IDESC = idesc_tf32(128, N_TILE);
      uint32_t acc = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES, it = kb / STAGES;
        mbar_wait(bar_full + 8 * s, it & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < BK / 8; ++kk) {
          const uint64_t ah = smem_desc(stage_a_hi(s) + 2 * kk * LBO_A, LBO_A, 128);
          const uint64_t bh = smem_desc(stage_b_hi(s) + 2 * kk * LBO_B, LBO_B, 128);
          if (SPLIT) {
            const uint64_t al = smem_desc(stage_a_lo(s) + 2 * kk * LBO_A, LBO_A, 128);
            const uint64_t bl = smem_desc(stage_b_lo(s) + 2 * kk * LBO_B, LBO_B, 128);
            umma_tf32(tmem_base, al, bh, IDESC, acc); acc = 1;
            umma_tf32(tmem_base, ah, bl, IDESC, 1);
          }
          umma_tf32(tmem_base, ah, bh, IDESC, acc); acc = 1;
        }
        umma_commit(bar_empty + 8 * s);     // stage reusable once these MMAs have read it
      }
      umma_commit(bar_tmem);                // accumulator complete
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, N_TILE);
  }
}

// ---- weight transform for dgrad-as-forward --------------------------------------------------------
// wt[g][c][o][i'][j'] = w[g*Og+o][c][kh-1-i'][kw-1-j']   (flip == 0 keeps (i,j))
__global__ void __launch_bounds__(256)
weight_transpose_flip_kernel(const float* __restrict__ w, float* __restrict__ wt, int G, int Og, int Cg, int kh, int kw) {
  const long long total = (long long)G * Og * Cg * kh * kw;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx % kw), i = (int)((idx / kw) % kh);
    const int o = (int)((idx / ((long long)kw * kh)) % Og), c = (int)((idx / ((long long)kw * kh * Og)) % Cg);
    const int g = (int)(idx / ((long long)kw * kh * Og * Cg));
    wt[idx] = __ldg(w + ((((long long)g * Og + o) * Cg + c) * kh + (kh - 1 - i)) * kw + (kw - 1 - j));
  }
}

template <int N_TILE, bool SPLIT, bool K1X1>
static int launch_fwdlike_inst(const FwdLikeParams& p, int G, cudaStream_t st) {
  using S = FwdLikeSmem<N_TILE, SPLIT>;
  static bool attr_set = false;
  if (!attr_set) {
    B2C_CUDA_OK(cudaFuncSetAttribute(igemm_fwdlike_kernel<N_TILE, SPLIT, K1X1>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::TOTAL));
    attr_set = true;
  }
  dim3 grid((p.Mtot + 127) / 128, (p.Ntot + N_TILE - 1) / N_TILE, G);
  igemm_fwdlike_kernel<N_TILE, SPLIT, K1X1><<<grid, TC_THREADS, S::TOTAL, st>>>(p);
  B2C_POST_LAUNCH();
  return B2C_OK;
}

template <int N_TILE>
static int launch_fwdlike_n(const FwdLikeParams& p, int G, int math, bool k1, cudaStream_t st) {
  if (math == B2C_MATH_FP32) {
    return k1 ? launch_fwdlike_inst<N_TILE, true, true>(p, G, st) : launch_fwdlike_inst<N_TILE, true, false>(p, G, st);
  }
  return k1 ? launch_fwdlike_inst<N_TILE, false, true>(p, G, st) : launch_fwdlike_inst<N_TILE, false, false>(p, G, st);
}

static int launch_fwdlike(const FwdLikeParams& p, int G, int math, cudaStream_t st) {
  const bool k1 = p.kh == 1 && p.kw == 1;
  if (p.Ntot > 128) return launch_fwdlike_n<256>(p, G, math, k1, st);
  if (p.Ntot > 64) return launch_fwdlike_n<128>(p, G, math, k1, st);
  if (p.Ntot > 32) return launch_fwdlike_n<64>(p, G, math, k1, st);
  return launch_fwdlike_n<32>(p, G, math, k1, st);
}

// ---- public entry points of this translation unit ---------------------------------------------------
bool tc_wgrad_supported(const ConvShape& s);
size_t tc_wgrad_workspace(const ConvShape& s);
int launch_conv_tc_wgrad(const ConvShape& s, int math, const float* x, const float* dy, float* dw, void* ws,
                         size_t ws_bytes, cudaStream_t st);

static bool dgrad_as_fwd(const ConvShape& s) { return s.sh == 1 && s.sw == 1; }
static bool dgrad_scatter(const ConvShape& s) { return s.kh == 1 && s.kw == 1 && s.ph == 0 && s.pw == 0 && (s.sh > 1 || s.sw > 1); }

bool tc_conv_supported(const ConvShape& s, int op) {
  const long long Mtot = (long long)s.N * s.Ho * s.Wo;
  if (Mtot > 0x7fffffffLL || (long long)s.N * s.H * s.W > 0x7fffffffLL) return false;
  if (op == B2C_OP_FORWARD) return true;
  if (op == B2C_OP_BACKWARD_DATA) return dgrad_as_fwd(s) || dgrad_scatter(s);
  return tc_wgrad_supported(s);
}

size_t tc_conv_workspace(const ConvShape& s, int op, int math) {
  (void)math;
  if (op == B2C_OP_BACKWARD_DATA) return sizeof(float) * (size_t)s.O * s.Cg * s.kh * s.kw;  // transposed filter
  if (op == B2C_OP_BACKWARD_FILTER) return tc_wgrad_workspace(s);                          // split-K partials
  return 0;
}

int launch_conv_tc(const ConvShape& s, int op, int math, const float* a, const float* b, const float* bias, float* out,
                   void* ws, size_t ws_bytes, cudaStream_t st) {
  FwdLikeParams p;
  if (op == B2C_OP_FORWARD) {
    p.x = a; p.Cin_tot = s.C; p.H = s.H; p.W = s.W; p.Cg = s.Cg;
    p.kh = s.kh; p.kw = s.kw; p.sh = s.sh; p.sw = s.sw; p.ph = s.ph; p.pw = s.pw; p.dh = s.dh; p.dw = s.dw;
    p.Ho = s.Ho; p.Wo = s.Wo; p.Mtot = s.N * s.Ho * s.Wo;
    p.w = b; p.Ntot = s.Og; p.K = s.Kd;
    p.out = out; p.Cout_tot = s.O; p.out_plane = (long long)s.Ho * s.Wo; p.out_hs = s.Wo; p.out_ws = 1;
    p.bias = bias;
    return launch_fwdlike(p, s.G, math, st);
  }
  if (op == B2C_OP_BACKWARD_DATA) {
    // a = dy [N,O,Ho,Wo], b = w [O,Cg,kh,kw], out = dx [N,C,H,W]
    const size_t need = sizeof(float) * (size_t)s.O * s.Cg * s.kh * s.kw;
    if (!ws || ws_bytes < need) return fail(B2C_ERR_WORKSPACE, "dgrad: workspace too small");
    float* wt = static_cast<float*>(ws);
    const long long total = (long long)s.O * s.Cg * s.kh * s.kw;
    weight_transpose_flip_kernel<<<grid_for((size_t)total, 256), 256, 0, st>>>(b, wt, s.G, s.Og, s.Cg, s.kh, s.kw);
    B2C_POST_LAUNCH();
    p.x = a; p.Cin_tot = s.O; p.Cg = s.Og;
    p.w = wt; p.Ntot = s.Cg; p.K = s.Og * s.kh * s.kw;
    p.out = out; p.Cout_tot = s.C; p.out_plane = (long long)s.H * s.W; p.bias = nullptr;
    if (dgrad_as_fwd(s)) {
      // stride-1 conv of dY with the flipped filter, pad' = (k-1)*d - p, output grid = bottom grid
      p.H = s.Ho; p.W = s.Wo;
      p.kh = s.kh; p.kw = s.kw; p.sh = 1; p.sw = 1; p.dh = s.dh; p.dw = s.dw;
      p.ph = (s.kh - 1) * s.dh - s.ph; p.pw = (s.kw - 1) * s.dw - s.pw;
      p.Ho = s.H; p.Wo = s.W; p.Mtot = s.N * s.H * s.W;
      p.out_hs = s.W; p.out_ws = 1;
      return launch_fwdlike(p, s.G, math, st);
    }
    // 1x1, stride > 1, pad 0: GEMM over the top grid, scatter to bottom[ho*sh][wo*sw]; everything else is 0
    B2C_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)s.N * s.C * s.H * s.W, st));
    p.H = s.Ho; p.W = s.Wo;
    p.kh = 1; p.kw = 1; p.sh = 1; p.sw = 1; p.dh = 1; p.dw = 1; p.ph = 0; p.pw = 0;
    p.Ho = s.Ho; p.Wo = s.Wo; p.Mtot = s.N * s.Ho * s.Wo;
    p.out_hs = s.sh * s.W; p.out_ws = s.sw;
    return launch_fwdlike(p, s.G, math, st);
  }
  if (op == B2C_OP_BACKWARD_FILTER) return launch_conv_tc_wgrad(s, math, a, b, out, ws, ws_bytes, st);
  return fail(B2C_ERR_INVALID, "tcgen05 path: unsupported op %d", op);
}

bool tc_gemm_supported(bool, bool, int, int, int) { return false; }
int launch_sgemm_tc(bool, bool, int, int, int, float, const float*, const float*, float, float*, int, cudaStream_t) {
  return fail(B2C_ERR_INVALID, "tcgen05 GEMM not built");
}

}  // namespace b2c
