// layers.cu -- the non-convolution layers on ResNet-50's training path, fp32 NCHW.
// All are HBM-bound elementwise / reduction kernels; each entry point cites the reference CPU code whose
// semantics it keeps (the parity oracle for them is oracle/layers_oracle.py).
//
//   ReLU            src/caffe/layers/relu_layer.cpp:10-41           y = max(x,0) + slope*min(x,0); dx = dy*(x>0 ? 1 : slope)
//   BatchNorm       src/caffe/layers/batch_norm_layer.cpp:140-300   batch statistics (biased variance, eps added before the
//                   (NVCaffe, scale_bias)                            inverse sqrt AND before the running average), x_norm kept
//                                                                    for backward; dgamma/dbeta OVERWRITTEN (compute_sum_*)
//   Pooling MAX/AVE src/caffe/layers/pooling_layer.cpp:129-318      ceil-mode extents, first-max index mask, AVE divides by the
//                                                                    padded window size
//   Eltwise SUM     src/caffe/layers/eltwise_layer.cpp
//   SoftmaxWithLoss src/caffe/layers/softmax_loss_layer.cpp:96-160  loss = -sum log p[label] / N (VALID normalisation, no
//                                                                    ignore_label); dx = (p - onehot) * loss_weight / N
#include <float.h>
#include <cooperative_groups.h>
#include "b2c_common.cuh"
#include "bn_common.cuh"

namespace cg = cooperative_groups;

namespace b2c {

// ---- ReLU ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) relu_fwd_kernel(size_t n, const float* __restrict__ x, float* __restrict__ y, float slope) {
  const size_t n4 = n / 4, tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, step = (size_t)gridDim.x * blockDim.x;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  float4* y4 = reinterpret_cast<float4*>(y);
  for (size_t i = tid; i < n4; i += step) {
    float4 v = x4[i];
    v.x = v.x > 0 ? v.x : v.x * slope; v.y = v.y > 0 ? v.y : v.y * slope; v.z = v.z > 0 ? v.z : v.z * slope; v.w = v.w > 0 ? v.w : v.w * slope;
    y4[i] = v;
  }
  for (size_t i = n4 * 4 + tid; i < n; i += step) { const float v = x[i]; y[i] = v > 0 ? v : v * slope; }
}
__global__ void __launch_bounds__(256) relu_bwd_kernel(size_t n, const float* __restrict__ dy, const float* __restrict__ x,
                                                       float* __restrict__ dx, float slope) {
  const size_t n4 = n / 4, tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, step = (size_t)gridDim.x * blockDim.x;
  const float4* d4 = reinterpret_cast<const float4*>(dy);
  const float4* x4 = reinterpret_cast<const float4*>(x);
  float4* o4 = reinterpret_cast<float4*>(dx);
  for (size_t i = tid; i < n4; i += step) {
    const float4 d = d4[i], v = x4[i];
    float4 o;
    o.x = d.x * (v.x > 0 ? 1.f : slope); o.y = d.y * (v.y > 0 ? 1.f : slope); o.z = d.z * (v.z > 0 ? 1.f : slope); o.w = d.w * (v.w > 0 ? 1.f : slope);
    o4[i] = o;
  }
  for (size_t i = n4 * 4 + tid; i < n; i += step) dx[i] = dy[i] * (x[i] > 0 ? 1.f : slope);
}

// ---- BatchNorm --------------------------------------------------------------------------------------------------
// Per-channel reductions run as one thread-block CLUSTER of 1..8 CTAs per channel (sized by the host from the channel's
// extent): each CTA streams its share of the channel's N*S values (float4 when the planes are 16-byte aligned), the eight partial sums meet in rank 0 through
// distributed shared memory in rank order (deterministic), no scratch buffer, no atomics, one launch.
// Statistics use sums of (x - k) and (x - k)^2 with k = the channel's first value, combined in double: one pass over
// HBM instead of the reference's two, without the cancellation of a raw E[x^2] - E[x]^2.
// (geometry, cursors and block sums: bn_common.cuh, shared with the fused kernels of layers_fused.cu)

template <bool VEC>
__global__ void __launch_bounds__(BN_THREADS)
bn_stats_kernel(int N, int C, int S, const float* __restrict__ x, float eps, float maf, int first,
                float* __restrict__ mean, float* __restrict__ invstd, float* __restrict__ run_mean, float* __restrict__ run_var) {
  __shared__ float2 part;
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank();
  const int c = blockIdx.y;
  const float k = x[(size_t)c * S];
  float a, b;
  bn_channel_partial<VEC, 0>(N, C, S, c, x, nullptr, k, rank, cluster.num_blocks(), a, b);
  block_sum2(a, b);
  if (threadIdx.x == 0) part = make_float2(a, b);
  cluster.sync();
  if (rank == 0 && threadIdx.x == 0) {
    double s1 = 0.0, s2 = 0.0;
    for (unsigned r = 0; r < cluster.num_blocks(); ++r) { const float2 v = *cluster.map_shared_rank(&part, r); s1 += v.x; s2 += v.y; }
    const double cnt = (double)N * S, m1 = s1 / cnt;
    const float m = (float)((double)k + m1);
    const float var_eps = (float)fmax(s2 / cnt - m1 * m1, 0.0) + eps;   // batch_norm_layer.cpp:183-186 (eps folded in before the average)
    mean[c] = m;
    invstd[c] = 1.0f / sqrtf(var_eps);
    if (first) { run_mean[c] = m; run_var[c] = var_eps; }                         // iter_ <= 1: copy (:199-204)
    else { run_mean[c] = (1.f - maf) * m + maf * run_mean[c]; run_var[c] = (1.f - maf) * var_eps + maf * run_var[c]; }
  }
  cluster.sync();                                   // remote shared memory stays valid until rank 0 has read it
}
// per channel: dgamma = sum dy*xn, dbeta = sum dy (both OVERWRITTEN, batch_norm_layer.cpp:247-252)
template <bool VEC>
__global__ void __launch_bounds__(BN_THREADS)
bn_bwd_reduce_kernel(int N, int C, int S, const float* __restrict__ dy, const float* __restrict__ xnorm,
                     float* __restrict__ sum_dy_xn, float* __restrict__ sum_dy) {
  __shared__ float2 part;
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank();
  const int c = blockIdx.y;
  float a, b;
  bn_channel_partial<VEC, 1>(N, C, S, c, dy, xnorm, 0.f, rank, cluster.num_blocks(), a, b);
  block_sum2(a, b);
  if (threadIdx.x == 0) part = make_float2(a, b);
  cluster.sync();
  if (rank == 0 && threadIdx.x == 0) {
    double s1 = 0.0, s2 = 0.0;
    for (unsigned r = 0; r < cluster.num_blocks(); ++r) { const float2 v = *cluster.map_shared_rank(&part, r); s1 += v.x; s2 += v.y; }
    sum_dy_xn[c] = (float)s1; sum_dy[c] = (float)s2;
  }
  cluster.sync();
}

// Elementwise passes: each block owns BN_EW_PER_THREAD * 256 consecutive units (float4 or float) of the NCHW tensor and
// walks (plane, offset) incrementally, so there is one integer division per thread, not per element.
constexpr int BN_EW_PER_THREAD = 8;
template <bool VEC> struct ChanCursor {
  unsigned p, c;
  __device__ __forceinline__ void init(size_t i, unsigned units, unsigned C) { const size_t plane = i / units; p = (unsigned)(i - plane * units); c = (unsigned)(plane % C); }
  __device__ __forceinline__ void advance(unsigned step, unsigned units, unsigned C) { p += step; while (p >= units) { p -= units; if (++c == C) c = 0; } }
};
template <bool VEC>
__global__ void __launch_bounds__(256)
bn_norm_kernel(size_t total_units, int C, int S, const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ invstd,
               const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ xnorm, float* __restrict__ y) {
  const unsigned units = VEC ? S / 4 : S;
  size_t i = (size_t)blockIdx.x * (256 * BN_EW_PER_THREAD) + threadIdx.x;
  if (i >= total_units) return;
  ChanCursor<VEC> cur;
  cur.init(i, units, C);
#pragma unroll 2
  for (int k = 0; k < BN_EW_PER_THREAD && i < total_units; ++k, i += 256) {
    const float m = mean[cur.c], is = invstd[cur.c], g = gamma ? gamma[cur.c] : 1.f, bt = gamma ? beta[cur.c] : 0.f;
    if (VEC) {
      const float4 v = reinterpret_cast<const float4*>(x)[i];
      float4 n4, o4;
      const bool affine = gamma != nullptr;
      n4.x = bn_xn(v.x, m, is); n4.y = bn_xn(v.y, m, is); n4.z = bn_xn(v.z, m, is); n4.w = bn_xn(v.w, m, is);
      o4.x = bn_y(n4.x, g, bt, affine); o4.y = bn_y(n4.y, g, bt, affine); o4.z = bn_y(n4.z, g, bt, affine); o4.w = bn_y(n4.w, g, bt, affine);
      reinterpret_cast<float4*>(xnorm)[i] = n4;
      reinterpret_cast<float4*>(y)[i] = o4;
    } else {
      const float xn = bn_xn(x[i], m, is);
      xnorm[i] = xn;
      y[i] = bn_y(xn, g, bt, gamma != nullptr);
    }
    cur.advance(256, units, C);
  }
}
// dx = gamma * invstd * (dy - mean(dy) - xn * mean(dy*xn))     (means over N*S; with gamma folded: :254-281)
template <bool VEC>
__global__ void __launch_bounds__(256)
bn_bwd_dx_kernel(size_t total_units, int C, int S, float inv_cnt, const float* __restrict__ dy, const float* __restrict__ xnorm,
                 const float* __restrict__ gamma, const float* __restrict__ invstd, const float* __restrict__ sum_dy_xn,
                 const float* __restrict__ sum_dy, float* __restrict__ dx) {
  const unsigned units = VEC ? S / 4 : S;
  size_t i = (size_t)blockIdx.x * (256 * BN_EW_PER_THREAD) + threadIdx.x;
  if (i >= total_units) return;
  ChanCursor<VEC> cur;
  cur.init(i, units, C);
#pragma unroll 2
  for (int k = 0; k < BN_EW_PER_THREAD && i < total_units; ++k, i += 256) {
    const float gi = __fmul_rn(gamma ? gamma[cur.c] : 1.f, invstd[cur.c]);
    const float mdy = bn_mean_term(sum_dy[cur.c], inv_cnt), mdx = bn_mean_term(sum_dy_xn[cur.c], inv_cnt);
    if (VEC) {
      const float4 d = reinterpret_cast<const float4*>(dy)[i], n4 = reinterpret_cast<const float4*>(xnorm)[i];
      float4 o;
      o.x = bn_dx(d.x, n4.x, gi, mdy, mdx); o.y = bn_dx(d.y, n4.y, gi, mdy, mdx); o.z = bn_dx(d.z, n4.z, gi, mdy, mdx); o.w = bn_dx(d.w, n4.w, gi, mdy, mdx);
      reinterpret_cast<float4*>(dx)[i] = o;
    } else {
      dx[i] = bn_dx(dy[i], xnorm[i], gi, mdy, mdx);
    }
    cur.advance(256, units, C);
  }
}

// ---- LRN across channels ------------------------------------------------------------------------------------------
// s^-beta: beta = 0.75 (every BASELINE net) is two reciprocal square roots instead of powf's ~30 instructions per element
__device__ __forceinline__ float lrn_pow_neg(float s, float beta) {
  if (beta == 0.75f) { const float r = rsqrtf(s); return r * sqrtf(r); }
  return powf(s, -beta);
}
// one thread per (image, pixel): walks the channels with a running window sum, loads coalesced along the pixel axis
__global__ void __launch_bounds__(256)
lrn_fwd_kernel(int N, int C, int S, int size, float alpha_over_size, float beta, float k, const float* __restrict__ x,
               float* __restrict__ scale, float* __restrict__ y) {
  const long long total = (long long)N * S;
  const int pre = (size - 1) / 2;                              // lrn_layer.cpp: pre_pad_
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / S, sp = i - n * S;
    const float* xp = x + n * C * S + sp;
    float* sc = scale + n * C * S + sp;
    float* yp = y + n * C * S + sp;
    float acc = 0.f;                                           // window of channel c: [c - pre, c - pre + size)
    for (int c = 0; c < size - 1 - pre && c < C; ++c) { const float v = xp[(long long)c * S]; acc += v * v; }
    for (int c = 0; c < C; ++c) {
      const int head = c - pre + size - 1, tail = c - pre - 1;
      if (head < C && head >= 0) { const float v = xp[(long long)head * S]; acc += v * v; }
      if (tail >= 0 && tail < C) { const float v = xp[(long long)tail * S]; acc -= v * v; }
      const float s_ = k + alpha_over_size * acc;
      sc[(long long)c * S] = s_;
      yp[(long long)c * S] = xp[(long long)c * S] * lrn_pow_neg(s_, beta);
    }
  }
}
__global__ void __launch_bounds__(256)
lrn_bwd_kernel(int N, int C, int S, int size, float cache_ratio, float beta, const float* __restrict__ x, const float* __restrict__ y,
               const float* __restrict__ scale, const float* __restrict__ dy, float* __restrict__ dx) {
  const long long total = (long long)N * S;
  const int ipp = size - (size + 1) / 2;                       // lrn_layer.cpp: inverse_pre_pad
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / S, sp = i - n * S, base = n * C * S + sp;
    auto ratio = [&](int c) { const long long o = base + (long long)c * S; return dy[o] * y[o] / scale[o]; };
    float acc = 0.f;                                           // window of channel c: [c - ipp, c - ipp + size)
    for (int c = 0; c < size - 1 - ipp && c < C; ++c) acc += ratio(c);
    for (int c = 0; c < C; ++c) {
      const int head = c - ipp + size - 1, tail = c - ipp - 1;
      if (head < C && head >= 0) acc += ratio(head);
      if (tail >= 0 && tail < C) acc -= ratio(tail);
      const long long o = base + (long long)c * S;
      dx[o] = dy[o] * lrn_pow_neg(scale[o], beta) - cache_ratio * x[o] * acc;
    }
  }
}

// ---- Dropout mask / elementwise product ----------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__global__ void __launch_bounds__(256)
dropout_mask_kernel(size_t n, unsigned threshold24, float scale, unsigned long long seed, unsigned long long offset, float* __restrict__ mask) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const unsigned u = (unsigned)(splitmix64(seed + offset + i) >> 40);      // 24 uniform bits
    mask[i] = u >= threshold24 ? scale : 0.f;
  }
}
__global__ void __launch_bounds__(256) mul_kernel(size_t n, const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y) {
  const size_t n4 = n / 4, tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, step = (size_t)gridDim.x * blockDim.x;
  for (size_t i = tid; i < n4; i += step) {
    const float4 u = reinterpret_cast<const float4*>(a)[i], v = reinterpret_cast<const float4*>(b)[i];
    reinterpret_cast<float4*>(y)[i] = make_float4(u.x * v.x, u.y * v.y, u.z * v.z, u.w * v.w);
  }
  for (size_t i = n4 * 4 + tid; i < n; i += step) y[i] = a[i] * b[i];
}

// ---- Pooling ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pool_max_fwd_kernel(size_t total, int H, int W, int Ho, int Wo, int kh, int kw, int sh, int sw, int ph, int pw,
                    const float* __restrict__ x, float* __restrict__ y, int* __restrict__ mask) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int wo = (int)(i % Wo), ho = (int)((i / Wo) % Ho);
    const size_t nc = i / ((size_t)Wo * Ho);
    int hs = ho * sh - ph, ws = wo * sw - pw;
    const int he = min(hs + kh, H), we = min(ws + kw, W);
    hs = max(hs, 0); ws = max(ws, 0);
    const float* src = x + nc * H * W;
    float best = -FLT_MAX;
    int bi = -1;
    for (int h = hs; h < he; ++h)
      for (int w = ws; w < we; ++w) {
        const float v = src[h * W + w];
        if (v > best) { best = v; bi = h * W + w; }       // strict '>' keeps the FIRST maximum, as the reference
      }
    y[i] = best;
    mask[i] = bi;
  }
}
// The common window shapes with Wo % Q == 0: one thread per Q consecutive outputs of a pooled row.  The (Q-1)*SW + KW input columns
// those windows cover are loaded once per window row (27 loads instead of 36 for 3x3 / 2 with Q = 4), results leave as one 16- or
// 8-byte store each for y and the mask.  Scan order and the strict '>' are those of the scalar kernel: same first maximum.
template <int KH, int KW, int SH, int SW, int Q>
__global__ void __launch_bounds__(256)
pool_max_fwd_q_kernel(long long groups, int H, int W, int Ho, int Wo, int ph, int pw, const float* __restrict__ x,
                      float* __restrict__ y, int* __restrict__ mask) {
  constexpr int NCOL = (Q - 1) * SW + KW;
  const int Wq = Wo / Q;
  for (long long gidx = (long long)blockIdx.x * blockDim.x + threadIdx.x; gidx < groups; gidx += (long long)gridDim.x * blockDim.x) {
    const long long orow = gidx / Wq;                             // pooled row index over (plane, ho)
    const int wo0 = (int)(gidx - orow * Wq) * Q;
    const long long nc = orow / Ho;
    const int ho = (int)(orow - nc * Ho);
    const int hs = ho * SH - ph, ws0 = wo0 * SW - pw;
    const float* src = x + nc * H * W;
    float best[Q];
    int bi[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) { best[q] = -FLT_MAX; bi[q] = -1; }
#pragma unroll
    for (int i = 0; i < KH; ++i) {
      const int h = hs + i;
      const bool hok = (unsigned)h < (unsigned)H;
      float v[NCOL];
#pragma unroll
      for (int j = 0; j < NCOL; ++j) {
        const int w = ws0 + j;
        v[j] = (hok && (unsigned)w < (unsigned)W) ? __ldg(src + (long long)h * W + w) : -FLT_MAX;
      }
#pragma unroll
      for (int q = 0; q < Q; ++q)
#pragma unroll
        for (int j = 0; j < KW; ++j) {
          const int w = ws0 + q * SW + j;
          if (hok && (unsigned)w < (unsigned)W && v[q * SW + j] > best[q]) { best[q] = v[q * SW + j]; bi[q] = h * W + w; }
        }
    }
    const long long o = orow * Wo + wo0;
    if (Q == 4) {
      *reinterpret_cast<float4*>(y + o) = make_float4(best[0], best[1], best[2], best[Q - 1]);
      *reinterpret_cast<int4*>(mask + o) = make_int4(bi[0], bi[1], bi[2], bi[Q - 1]);
    } else {
      *reinterpret_cast<float2*>(y + o) = make_float2(best[0], best[Q - 1]);
      *reinterpret_cast<int2*>(mask + o) = make_int2(bi[0], bi[Q - 1]);
    }
  }
}
// One block per input row (plane nc, row h): the row's window range in h is computed once, threads walk w.  (The first version
// decoded a flat 64-bit element index with three runtime divisions per element and ran at 380 GB/s -- 0.67 ms per ResNet-50 step for
// its one 3x3 / stride 2 pool, profiles/r02_c6_fullnet_launches.csv.)  Same ascending (a, b) summation order as before.
__global__ void __launch_bounds__(128)
pool_max_bwd_kernel(int rows, int H, int W, int Ho, int Wo, int kh, int kw, int sh, int sw, int ph, int pw,
                    const float* __restrict__ dy, const int* __restrict__ mask, float* __restrict__ dx) {
  const int rpb = W >= 128 ? 1 : 128 / W;                    // input rows per block (narrow maps: several rows share the 128 threads)
  const int tr = rpb == 1 ? 0 : (int)threadIdx.x / W;        // this thread's row inside the block's group, first column
  const int tw = rpb == 1 ? (int)threadIdx.x : (int)threadIdx.x - tr * W;
  if (tr >= rpb) return;
  for (int row = blockIdx.x * rpb + tr; row < rows; row += gridDim.x * rpb) {
    const int nc = row / H, h = row - nc * H;
    const int phs = (h + ph < kh) ? 0 : (h + ph - kh) / sh + 1, phe = min((h + ph) / sh + 1, Ho);
    const float* d = dy + (size_t)nc * Ho * Wo;
    const int* m = mask + (size_t)nc * Ho * Wo;
    float* out = dx + (size_t)row * W;
    for (int w = tw; w < W; w += 128) {
      const int pws = (w + pw < kw) ? 0 : (w + pw - kw) / sw + 1, pwe = min((w + pw) / sw + 1, Wo);
      const int me = h * W + w;
      float g = 0.f;
      for (int a = phs; a < phe; ++a)
        for (int b = pws; b < pwe; ++b)
          if (m[a * Wo + b] == me) g += d[a * Wo + b];
      out[w] = g;
    }
  }
}
// The common window shapes (3x3 / stride 2, 2x2 / stride 2, 3x3 / stride 1) with W % 4 == 0: one thread per FOUR consecutive input
// columns.  The windows those four columns can sit in are MA rows x NB columns of the pooled map; their (mask, dy) pairs are loaded
// ONCE, unconditionally (12 loads for 3x3 / 2), and each of the four elements picks its own from registers -- the row kernel
// above has one element and two dependent round trips per thread (0.35 TB/s: 0.6 ms per ResNet-50 step, 5.9 ms per GoogLeNet
// step); loading per element instead of per window ran into the load/store unit (32 loads per 16-byte store, 0.94 TB/s).
// Same ascending (a, b) summation order: same bits.
template <int KH, int KW, int SH, int SW>
__global__ void __launch_bounds__(256)
pool_max_bwd_q4_kernel(long long quads, int H, int W, int Ho, int Wo, int ph, int pw, const float* __restrict__ dy,
                       const int* __restrict__ mask, float* __restrict__ dx) {
  constexpr int MA = (KH + SH - 1) / SH;                        // pooled rows an input row can sit in
  constexpr int NB = (3 + KW - 1) / SW + 1;                    // pooled columns four consecutive input columns can sit in
  const int Wq = W / 4;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += (long long)gridDim.x * blockDim.x) {
    const long long row = q / Wq;
    const int w0 = (int)(q - row * Wq) * 4;
    const long long nc = row / H;
    const int h = (int)(row - nc * H);
    const int phs = (h + ph < KH) ? 0 : (h + ph - KH) / SH + 1, phe = min((h + ph) / SH + 1, Ho);
    const int b0 = (w0 + pw < KW) ? 0 : (w0 + pw - KW) / SW + 1;                    // first pooled column of input column w0
    const float* d = dy + nc * Ho * Wo;
    const int* m = mask + nc * Ho * Wo;
    int mm[MA][NB];
    float dd[MA][NB];
#pragma unroll
    for (int a = 0; a < MA; ++a)
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const bool ok = phs + a < phe && b0 + b < Wo;
        const int idx = ok ? (phs + a) * Wo + b0 + b : 0;
        mm[a][b] = ok ? __ldg(m + idx) : -1;                    // -1 never equals an element index
        dd[a][b] = __ldg(d + idx);
      }
    float g[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int w = w0 + e;
      const int pws = (w + pw < KW) ? 0 : (w + pw - KW) / SW + 1, pwe = min((w + pw) / SW + 1, Wo);
      const int me = h * W + w;
      float acc = 0.f;
#pragma unroll
      for (int a = 0; a < MA; ++a)
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          const int bb = b0 + b;
          acc += (bb >= pws && bb < pwe && mm[a][b] == me) ? dd[a][b] : 0.f;
        }
      g[e] = acc;
    }
    *reinterpret_cast<float4*>(dx + row * W + w0) = make_float4(g[0], g[1], g[2], g[3]);
  }
}
__global__ void __launch_bounds__(256)
pool_ave_fwd_kernel(size_t total, int H, int W, int Ho, int Wo, int kh, int kw, int sh, int sw, int ph, int pw,
                    const float* __restrict__ x, float* __restrict__ y) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int wo = (int)(i % Wo), ho = (int)((i / Wo) % Ho);
    const size_t nc = i / ((size_t)Wo * Ho);
    int hs = ho * sh - ph, ws = wo * sw - pw;
    int he = min(hs + kh, H + ph), we = min(ws + kw, W + pw);
    const int pool_size = (he - hs) * (we - ws);          // padded window size (pooling_layer.cpp:200-206)
    hs = max(hs, 0); ws = max(ws, 0); he = min(he, H); we = min(we, W);
    const float* src = x + nc * H * W;
    float s = 0.f;
    for (int h = hs; h < he; ++h)
      for (int w = ws; w < we; ++w) s += src[h * W + w];
    y[i] = s / pool_size;
  }
}
__global__ void __launch_bounds__(256)
pool_ave_bwd_kernel(size_t total, int H, int W, int Ho, int Wo, int kh, int kw, int sh, int sw, int ph, int pw,
                    const float* __restrict__ dy, float* __restrict__ dx) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int w = (int)(i % W) + pw, h = (int)((i / W) % H) + ph;
    const size_t nc = i / ((size_t)W * H);
    const int phs = (h < kh) ? 0 : (h - kh) / sh + 1, phe = min(h / sh + 1, Ho);
    const int pws = (w < kw) ? 0 : (w - kw) / sw + 1, pwe = min(w / sw + 1, Wo);
    const float* d = dy + nc * Ho * Wo;
    float g = 0.f;
    for (int a = phs; a < phe; ++a)
      for (int b = pws; b < pwe; ++b) {
        const int hs = a * sh - ph, ws = b * sw - pw;
        const int he = min(hs + kh, H + ph), we = min(ws + kw, W + pw);
        g += d[a * Wo + b] / ((he - hs) * (we - ws));
      }
    dx[i] = g;
  }
}

// ---- elementwise ----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) add_kernel(size_t n, const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y) {
  const size_t n4 = n / 4, tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, step = (size_t)gridDim.x * blockDim.x;
  for (size_t i = tid; i < n4; i += step) {
    const float4 u = reinterpret_cast<const float4*>(a)[i], v = reinterpret_cast<const float4*>(b)[i];
    reinterpret_cast<float4*>(y)[i] = make_float4(u.x + v.x, u.y + v.y, u.z + v.z, u.w + v.w);
  }
  for (size_t i = n4 * 4 + tid; i < n; i += step) y[i] = a[i] + b[i];
}

// ---- softmax with loss ------------------------------------------------------------------------------------------
// one block per sample; prob[N,C]; labels are float class ids.  The loss itself is summed by softmax_loss_sum_kernel in a fixed
// order (an atomicAdd per sample made the last bit of the reported loss depend on block scheduling)
__global__ void __launch_bounds__(256)
softmax_loss_fwd_kernel(int C, const float* __restrict__ logits, float* __restrict__ prob) {
  const int n = blockIdx.x;
  const float* z = logits + (size_t)n * C;
  float mx = -FLT_MAX, dummy = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) mx = fmaxf(mx, z[c]);
#pragma unroll
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  __shared__ float smx[32];
  if ((threadIdx.x & 31) == 0) smx[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = smx[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) mx = fmaxf(mx, smx[w]);
  float s = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) { const float e = expf(z[c] - mx); prob[(size_t)n * C + c] = e; s += e; }
  block_sum2(s, dummy);
  const float inv = 1.f / s;
  for (int c = threadIdx.x; c < C; c += blockDim.x) prob[(size_t)n * C + c] *= inv;
}
// loss = scale * sum_n -log(max(prob[n][label[n]], FLT_MIN))   (softmax_loss_layer.cpp:100-116); one block, fixed summation order
__global__ void __launch_bounds__(256)
softmax_loss_sum_kernel(int N, int C, const float* __restrict__ prob, const float* __restrict__ labels, float scale, float* __restrict__ loss) {
  float s = 0.f, dummy = 0.f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    // the reference DCHECKs 0 <= label < C (:108-109); an out-of-range label must not read out of bounds
    const int lab = min(max((int)labels[n], 0), C - 1);
    s += -logf(fmaxf(prob[(size_t)n * C + lab], FLT_MIN));
  }
  block_sum2(s, dummy);
  if (threadIdx.x == 0) *loss = s * scale;
}
__global__ void __launch_bounds__(256)
softmax_loss_bwd_kernel(size_t total, int C, const float* __restrict__ prob, const float* __restrict__ labels, float scale,
                        float* __restrict__ dx) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t n = i / C;
    const int c = (int)(i - n * C);
    dx[i] = (prob[i] - (c == (int)labels[n] ? 1.f : 0.f)) * scale;
  }
}

}  // namespace b2c

using namespace b2c;
#define NEED(cond, msg) if (!(cond)) return fail(B2C_ERR_INVALID, msg)

extern "C" int b2c_relu_forward(size_t n, const float* x, float* y, float negative_slope, void* stream) {
  NEED(x && y, "b2c_relu_forward: null");
  if (!n) return B2C_OK;
  relu_fwd_kernel<<<grid_for((n + 3) / 4, 256), 256, 0, as_stream(stream)>>>(n, x, y, negative_slope);
  B2C_POST_LAUNCH();
  return B2C_OK;
}
extern "C" int b2c_relu_backward(size_t n, const float* dy, const float* x, float* dx, float negative_slope, void* stream) {
  NEED(dy && x && dx, "b2c_relu_backward: null");
  if (!n) return B2C_OK;
  relu_bwd_kernel<<<grid_for((n + 3) / 4, 256), 256, 0, as_stream(stream)>>>(n, dy, x, dx, negative_slope);
  B2C_POST_LAUNCH();
  return B2C_OK;
}
template <typename... Args>
static void launch_clustered(void (*kernel)(Args...), unsigned cs, int C, void* stream, unsigned threads, Args... args) {
  bn_launch_clustered(kernel, cs, C, 0, stream, threads, args...);
}
static bool vec_ok(int S, std::initializer_list<const void*> ptrs) {
  if (S % 4) return false;
  for (const void* p : ptrs) if (reinterpret_cast<uintptr_t>(p) & 15) return false;
  return true;
}
namespace b2c {
int launch_bn_stats(int N, int C, int S, const float* x, float eps, float maf, int first, float* mean, float* invstd, float* run_mean,
                    float* run_var, bool vec, void* stream) {
  const unsigned cs = bn_cluster_size(N, C, S);
  launch_clustered(vec ? bn_stats_kernel<true> : bn_stats_kernel<false>, cs, C, stream, bn_threads(N, S, true), N, C, S, x, eps, maf, first, mean, invstd, run_mean, run_var);
  B2C_POST_LAUNCH();
  return B2C_OK;
}
}  // namespace b2c
extern "C" int b2c_bn_forward_train(int N, int C, int S, const float* x, const float* gamma, const float* beta, float eps,
                                    float moving_average_fraction, int first_iteration, float* running_mean, float* running_var,
                                    float* save_mean, float* save_invstd, float* xnorm, float* y, void* stream) {
  NEED(x && y && xnorm && save_mean && save_invstd && running_mean && running_var && N > 0 && C > 0 && S > 0, "b2c_bn_forward_train: bad argument");
  NEED((gamma == nullptr) == (beta == nullptr), "b2c_bn_forward_train: gamma and beta go together");
  NEED((size_t)N * S < (1ull << 31) && C <= 65535, "b2c_bn_forward_train: channel extent out of range");
  const bool vec = vec_ok(S, {x, y, xnorm});
  const unsigned cs = bn_cluster_size(N, C, S);
  launch_clustered(vec ? bn_stats_kernel<true> : bn_stats_kernel<false>, cs, C, stream, bn_threads(N, S, true), N, C, S, x, eps, moving_average_fraction, first_iteration,
                   save_mean, save_invstd, running_mean, running_var);
  B2C_POST_LAUNCH();
  const size_t units = (size_t)N * C * (vec ? S / 4 : S);
  const unsigned blocks = (unsigned)((units + 256 * BN_EW_PER_THREAD - 1) / (256 * BN_EW_PER_THREAD));
  if (vec) bn_norm_kernel<true><<<blocks, 256, 0, as_stream(stream)>>>(units, C, S, x, save_mean, save_invstd, gamma, beta, xnorm, y);
  else bn_norm_kernel<false><<<blocks, 256, 0, as_stream(stream)>>>(units, C, S, x, save_mean, save_invstd, gamma, beta, xnorm, y);
  B2C_POST_LAUNCH();
  return B2C_OK;
}
extern "C" int b2c_bn_backward(int N, int C, int S, const float* dy, const float* xnorm, const float* gamma, const float* save_invstd,
                               float* dgamma, float* dbeta, float* dx, void* stream) {
  NEED(dy && xnorm && save_invstd && dgamma && dbeta && dx, "b2c_bn_backward: null (dgamma/dbeta double as the reduction scratch)");
  NEED(N > 0 && C > 0 && S > 0 && (size_t)N * S < (1ull << 31) && C <= 65535, "b2c_bn_backward: channel extent out of range");
  const bool vec = vec_ok(S, {dy, xnorm, dx});
  launch_clustered(vec ? bn_bwd_reduce_kernel<true> : bn_bwd_reduce_kernel<false>, bn_cluster_size(N, C, S), C, stream, bn_threads(N, S, false), N, C, S, dy, xnorm, dgamma, dbeta);
  B2C_POST_LAUNCH();
  const size_t units = (size_t)N * C * (vec ? S / 4 : S);
  const unsigned blocks = (unsigned)((units + 256 * BN_EW_PER_THREAD - 1) / (256 * BN_EW_PER_THREAD));
  const float inv_cnt = 1.0f / ((float)N * S);
  if (vec) bn_bwd_dx_kernel<true><<<blocks, 256, 0, as_stream(stream)>>>(units, C, S, inv_cnt, dy, xnorm, gamma, save_invstd, dgamma, dbeta, dx);
  else bn_bwd_dx_kernel<false><<<blocks, 256, 0, as_stream(stream)>>>(units, C, S, inv_cnt, dy, xnorm, gamma, save_invstd, dgamma, dbeta, dx);
  B2C_POST_LAUNCH();
  return B2C_OK;
}
extern "C" int b2c_lrn_forward(int N, int C, int S, int local_size, float alpha, float beta, float k, const float* x, float* scale,
                               float* y, void* stream) {
  NEED(x && scale && y && N > 0 && C > 0 && S > 0 && local_size > 0 && (local_size & 1), "b2c_lrn_forward: bad argument (LRN only supports odd values for local_size)");
  lrn_fwd_kernel<<<grid_for((size_t)N * S, 256), 256, 0, as_stream(stream)>>>(N, C, S, local_size, alpha / local_size, beta, k, x, scale, y);
  B2C_POST_LAUNCH();
  return B2C_OK;
}
extern "C" int b2c_lrn_backward(int N, int C, int S, int local_size, float alpha, float beta, const float* x, const float* y,
                                const float* scale, const float* dy, float* dx, void* stream) {
  NEED(x && y && scale && dy && dx && N > 0 && C > 0 && S > 0 && local_size > 0, "b2c_lrn_backward: bad argument");
  lrn_bwd_kernel<<<grid_for((size_t)N * S, 256), 256, 0, as_stream(stream)>>>(N, C, S, local_size, 2.f * alpha * beta / local_size, beta, x, y, scale, dy, dx);
  B2C_POST_LAUNCH();
  return B2C_OK;
}
extern "C" int b2c_dropout_mask(size_t n, float ratio, unsigned long long seed, unsigned long long offset, float* mask, void* stream) {
  NEED(mask && ratio >= 0.f && ratio < 1.f, "b2c_dropout_mask: dropout_ratio must be in [0, 1)");
  if (!n) return B2C_OK;
  const unsigned thr = (unsigned)((double)ratio * 16777216.0);
  dropout_mask_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(n, thr, 1.0f / (1.0f - ratio), seed, offset, mask);
  B2C_POST_LAUNCH();
  return B2C_OK;
}
extern "C" int b2c_mul(size_t n, const float* a, const float* b, float* y, void* stream) {
  NEED(a && b && y, "b2c_mul: null");
  if (!n) return B2C_OK;
  NEED(((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(y)) & 15) == 0, "b2c_mul: pointers must be 16-byte aligned");
  mul_kernel<<<grid_for((n + 3) / 4, 256), 256, 0, as_stream(stream)>>>(n, a, b, y);
  B2C_POST_LAUNCH();
  return B2C_OK;
}
static int pool_out(int in, int k, int s, int p) {
  int o = (int)ceilf((float)(in + 2 * p - k) / s) + 1;
  if (p > 0 && (o - 1) * s >= in + p) --o;
  return o;
}
extern "C" int b2c_pool_forward(int method, int NC, int H, int W, int kh, int kw, int sh, int sw, int ph, int pw, const float* x,
                                float* y, int* mask, void* stream) {
  NEED(x && y && NC > 0 && kh > 0 && kw > 0 && sh > 0 && sw > 0, "b2c_pool_forward: bad argument");
  const int Ho = pool_out(H, kh, sh, ph), Wo = pool_out(W, kw, sw, pw);
  const size_t total = (size_t)NC * Ho * Wo;
  if (method == 0) {
    NEED(mask, "b2c_pool_forward: MAX needs a mask buffer");
    const bool common = kh == kw && sh == sw && ((kh == 3 && sh == 2) || (kh == 2 && sh == 2) || (kh == 3 && sh == 1));
    const int q = Wo % 4 == 0 ? 4 : Wo % 2 == 0 ? 2 : 0;
    const uintptr_t al = reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(mask);
    if (common && q && (al & (q == 4 ? 15 : 7)) == 0) {
      const long long groups = (long long)NC * Ho * (Wo / q);
      const unsigned grid = grid_for((size_t)groups, 256);
      cudaStream_t s_ = as_stream(stream);
#define B2C_POOLF(KH, S, Q) pool_max_fwd_q_kernel<KH, KH, S, S, Q><<<grid, 256, 0, s_>>>(groups, H, W, Ho, Wo, ph, pw, x, y, mask)
      if (kh == 3 && sh == 2) { if (q == 4) B2C_POOLF(3, 2, 4); else B2C_POOLF(3, 2, 2); }
      else if (kh == 2) { if (q == 4) B2C_POOLF(2, 2, 4); else B2C_POOLF(2, 2, 2); }
      else { if (q == 4) B2C_POOLF(3, 1, 4); else B2C_POOLF(3, 1, 2); }
#undef B2C_POOLF
    } else {
      pool_max_fwd_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(total, H, W, Ho, Wo, kh, kw, sh, sw, ph, pw, x, y, mask);
    }
  } else if (method == 1) {
    pool_ave_fwd_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(total, H, W, Ho, Wo, kh, kw, sh, sw, ph, pw, x, y);
  } else {
    return fail(B2C_ERR_INVALID, "b2c_pool_forward: STOCHASTIC pooling is outside this path");
  }
  B2C_POST_LAUNCH();
  return B2C_OK;
}
extern "C" int b2c_pool_backward(int method, int NC, int H, int W, int kh, int kw, int sh, int sw, int ph, int pw, const float* dy,
                                 const int* mask, float* dx, void* stream) {
  NEED(dy && dx && NC > 0, "b2c_pool_backward: bad argument");
  const int Ho = pool_out(H, kh, sh, ph), Wo = pool_out(W, kw, sw, pw);
  const size_t total = (size_t)NC * H * W;
  if (method == 0) {
    NEED(mask, "b2c_pool_backward: MAX needs the forward mask");
    {
      const long long rows = (long long)NC * H;
      NEED(rows < 0x7fffffffLL, "b2c_pool_backward: too many rows");
      if (W % 4 == 0 && (reinterpret_cast<uintptr_t>(dx) & 15) == 0 && kh == kw && sh == sw &&
          ((kh == 3 && sh == 2) || (kh == 2 && sh == 2) || (kh == 3 && sh == 1))) {
        const long long quads = rows * (W / 4);
        const unsigned grid = grid_for((size_t)quads, 256);
        cudaStream_t s_ = as_stream(stream);
        if (kh == 3 && sh == 2) pool_max_bwd_q4_kernel<3, 3, 2, 2><<<grid, 256, 0, s_>>>(quads, H, W, Ho, Wo, ph, pw, dy, mask, dx);
        else if (kh == 2) pool_max_bwd_q4_kernel<2, 2, 2, 2><<<grid, 256, 0, s_>>>(quads, H, W, Ho, Wo, ph, pw, dy, mask, dx);
        else pool_max_bwd_q4_kernel<3, 3, 1, 1><<<grid, 256, 0, s_>>>(quads, H, W, Ho, Wo, ph, pw, dy, mask, dx);
        B2C_POST_LAUNCH();
        return B2C_OK;
      }
      const long long groups = (rows + (W >= 128 ? 1 : 128 / W) - 1) / (W >= 128 ? 1 : 128 / W);
      const int grid = (int)(groups < (long long)sm_count() * 64 ? groups : (long long)sm_count() * 64);
      pool_max_bwd_kernel<<<grid, 128, 0, as_stream(stream)>>>((int)rows, H, W, Ho, Wo, kh, kw, sh, sw, ph, pw, dy, mask, dx);
    }
  } else if (method == 1) {
    pool_ave_bwd_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(total, H, W, Ho, Wo, kh, kw, sh, sw, ph, pw, dy, dx);
  } else {
    return fail(B2C_ERR_INVALID, "b2c_pool_backward: STOCHASTIC pooling is outside this path");
  }
  B2C_POST_LAUNCH();
  return B2C_OK;
}
// dst[c][r] = src[r][c]: 32 x 32 tiles through shared memory (padded: conflict-free both ways), coalesced reads and writes
__global__ void __launch_bounds__(256)
transpose_kernel(int rows, int cols, const float* __restrict__ src, float* __restrict__ dst) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;             // 32 x 8 threads
  const long long c0 = (long long)blockIdx.x * 32, r0 = (long long)blockIdx.y * 32;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const long long r = r0 + ty + 8 * j, c = c0 + tx;
    if (r < rows && c < cols) tile[ty + 8 * j][tx] = src[r * cols + c];
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const long long c = c0 + ty + 8 * j, r = r0 + tx;
    if (r < rows && c < cols) dst[c * rows + r] = tile[tx][ty + 8 * j];
  }
}
extern "C" int b2c_transpose(int rows, int cols, const float* src, float* dst, void* stream) {
  NEED(src && dst && rows > 0 && cols > 0 && src != dst, "b2c_transpose: bad argument");
  dim3 grid((cols + 31) / 32, (rows + 31) / 32);
  NEED(grid.y <= 65535, "b2c_transpose: too many rows");
  transpose_kernel<<<grid, 256, 0, as_stream(stream)>>>(rows, cols, src, dst);
  B2C_POST_LAUNCH();
  return B2C_OK;
}
extern "C" int b2c_add(size_t n, const float* a, const float* b, float* y, void* stream) {
  NEED(a && b && y, "b2c_add: null");
  if (!n) return B2C_OK;
  add_kernel<<<grid_for((n + 3) / 4, 256), 256, 0, as_stream(stream)>>>(n, a, b, y);
  B2C_POST_LAUNCH();
  return B2C_OK;
}
extern "C" int b2c_softmax_loss_forward(int N, int C, const float* logits, const float* labels, float* prob, float* loss, void* stream) {
  NEED(logits && labels && prob && loss && N > 0 && C > 0, "b2c_softmax_loss_forward: bad argument");
  softmax_loss_fwd_kernel<<<N, 256, 0, as_stream(stream)>>>(C, logits, prob);
  B2C_POST_LAUNCH();
  softmax_loss_sum_kernel<<<1, 256, 0, as_stream(stream)>>>(N, C, prob, labels, 1.0f / (float)N, loss);   // VALID normalisation, no ignore_label
  B2C_POST_LAUNCH();
  return B2C_OK;
}
extern "C" int b2c_softmax_loss_backward(int N, int C, const float* prob, const float* labels, float loss_weight, float* dx, void* stream) {
  NEED(prob && labels && dx && N > 0 && C > 0, "b2c_softmax_loss_backward: bad argument");
  const size_t total = (size_t)N * C;
  softmax_loss_bwd_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(total, C, prob, labels, loss_weight / (float)N, dx);
  B2C_POST_LAUNCH();
  return B2C_OK;
}
// y[n][o] += b[o] and db[o] += sum_n dy[n][o]: InnerProduct bias (inner_product_layer.cpp), via the conv bias kernels with P = 1
namespace b2c {
int launch_bias_add(int, int, int, const float*, float*, cudaStream_t);
int launch_bias_grad(int, int, int, const float*, float*, cudaStream_t);
}
extern "C" int b2c_bias_forward(int N, int O, int P, const float* bias, float* y, void* stream) {
  NEED(bias && y, "b2c_bias_forward: null");
  return launch_bias_add(N, O, P, bias, y, as_stream(stream));
}
extern "C" int b2c_bias_backward(int N, int O, int P, const float* dy, float* db, void* stream) {
  NEED(dy && db, "b2c_bias_backward: null");
  return launch_bias_grad(N, O, P, dy, db, as_stream(stream));
}
