/*
 * b2o_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference's CPU algorithm for the data-parallel
 * conv + gradient-allreduce + SGD hot path.  It is the parity checker for the
 * CUDA kernels under caffe_mpi_b200/csrc.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  Nothing in
 * the product path calls into this file.
 *
 * Parity status: PINNED.  The functions below are checked in tests/ against
 *   - the reference's own known-answer vectors (test_util_blas.cpp:19-175 integer
 *     GEMM/GEMV, test_convolution_layer.cpp:511-604 Sobel, test_im2col_layer.cpp:63-77),
 *   - oracle/_ref (the reference's src/caffe/util/im2col.cpp compiled verbatim),
 *   - fixtures under tests/golden/ produced from oracle/_ref.
 *
 * Citations are relative to /root/reference.  All tensors fp32, dense row-major,
 * activations NCHW, weights [O, C/g, kh, kw].
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int N, C, H, W;   /* bottom shape                                  */
  int O, G;         /* num_output, group                             */
  int kh, kw, sh, sw, ph, pw, dh, dw;
  int has_bias;
} b2o_conv_params;

/* conv_layer.cpp:7-22: truncating integer division for the output extent. */
static int out_extent(int in, int k, int s, int p, int d) {
  const int ext = d * (k - 1) + 1;
  return (in + 2 * p - ext) / s + 1;
}
int b2o_out_h(const b2o_conv_params* p) { return out_extent(p->H, p->kh, p->sh, p->ph, p->dh); }
int b2o_out_w(const b2o_conv_params* p) { return out_extent(p->W, p->kw, p->sw, p->pw, p->dw); }

/* ------------------------------------------------------------------------- *
 * im2col / col2im, 2-D.  Follows src/caffe/util/im2col.cpp:18-55 (im2col_cpu)
 * and :176-211 (col2im_cpu).  Written from the index formula
 *   col[(c*kh+i)*kw+j][ho][wo] = im[c][ho*sh-ph+i*dh][wo*sw-pw+j*dw]  (0 outside)
 * col2im zero-fills then accumulates in (c,i,j,ho,wo) order, which fixes the
 * fp32 summation order per input pixel: ascending (i,j).
 * ------------------------------------------------------------------------- */
void b2o_im2col(const float* im, int C, int H, int W, int kh, int kw,
                int ph, int pw, int sh, int sw, int dh, int dw, float* col) {
  const int Ho = out_extent(H, kh, sh, ph, dh), Wo = out_extent(W, kw, sw, pw, dw);
  size_t q = 0;
  for (int c = 0; c < C; ++c)
    for (int i = 0; i < kh; ++i)
      for (int j = 0; j < kw; ++j)
        for (int ho = 0; ho < Ho; ++ho) {
          const int h = ho * sh - ph + i * dh;
          for (int wo = 0; wo < Wo; ++wo, ++q) {
            const int w = wo * sw - pw + j * dw;
            const int inside = (h >= 0) & (h < H) & (w >= 0) & (w < W);
            col[q] = inside ? im[((size_t)c * H + h) * W + w] : 0.0f;
          }
        }
}

void b2o_col2im(const float* col, int C, int H, int W, int kh, int kw,
                int ph, int pw, int sh, int sw, int dh, int dw, float* im) {
  const int Ho = out_extent(H, kh, sh, ph, dh), Wo = out_extent(W, kw, sw, pw, dw);
  memset(im, 0, sizeof(float) * (size_t)C * H * W);
  size_t q = 0;
  for (int c = 0; c < C; ++c)
    for (int i = 0; i < kh; ++i)
      for (int j = 0; j < kw; ++j)
        for (int ho = 0; ho < Ho; ++ho) {
          const int h = ho * sh - ph + i * dh;
          for (int wo = 0; wo < Wo; ++wo, ++q) {
            const int w = wo * sw - pw + j * dw;
            if (h >= 0 && h < H && w >= 0 && w < W)
              im[((size_t)c * H + h) * W + w] += col[q];
          }
        }
}

/* ------------------------------------------------------------------------- *
 * N-D im2col / col2im (1..10 spatial axes), im2col.cpp:76-174,231-258.
 * shapes: im_shape = [C, d0, d1, ...], col_shape = [C*prod(k), o0, o1, ...].
 * Restated with a flat decode of the column index instead of the reference's
 * odometer; the visiting order (and therefore the col2im summation order per
 * image element) is identical: ascending flat column index.
 * ------------------------------------------------------------------------- */
static void nd_core(const float* in, int to_col, int nax, const int* im_shape,
                    const int* col_shape, const int* k, const int* pad,
                    const int* stride, const int* dil, float* out) {
  size_t im_size = im_shape[0], out_sp = 1;
  int ksize = 1;
  for (int a = 0; a < nax; ++a) { im_size *= im_shape[1 + a]; out_sp *= col_shape[1 + a]; ksize *= k[a]; }
  if (!to_col) memset(out, 0, sizeof(float) * im_size);
  const int ccol = col_shape[0];
  int koff[10], opos[10];
  for (int cc = 0; cc < ccol; ++cc) {
    int r = cc;
    for (int a = nax - 1; a >= 0; --a) { koff[a] = r % k[a]; r /= k[a]; }
    const int c_im = cc / ksize;
    for (size_t s = 0; s < out_sp; ++s) {
      size_t t = s;
      for (int a = nax - 1; a >= 0; --a) { opos[a] = (int)(t % col_shape[1 + a]); t /= col_shape[1 + a]; }
      size_t idx_im = c_im;
      int padded = 0;
      for (int a = 0; a < nax; ++a) {
        const int d_im = opos[a] * stride[a] - pad[a] + koff[a] * dil[a];
        padded |= (d_im < 0) || (d_im >= im_shape[1 + a]);
        idx_im = idx_im * im_shape[1 + a] + (size_t)d_im;
      }
      const size_t idx_col = (size_t)cc * out_sp + s;
      if (to_col) out[idx_col] = padded ? 0.0f : in[idx_im];
      else if (!padded) out[idx_im] += in[idx_col];
    }
  }
}
void b2o_im2col_nd(const float* im, int nax, const int* im_shape, const int* col_shape,
                   const int* k, const int* pad, const int* stride, const int* dil, float* col) {
  nd_core(im, 1, nax, im_shape, col_shape, k, pad, stride, dil, col);
}
void b2o_col2im_nd(const float* col, int nax, const int* im_shape, const int* col_shape,
                   const int* k, const int* pad, const int* stride, const int* dil, float* im) {
  nd_core(col, 0, nax, im_shape, col_shape, k, pad, stride, dil, im);
}

/* ------------------------------------------------------------------------- *
 * GEMM / GEMV with the reference's argument conventions,
 * src/caffe/util/math_functions.cpp:14-23 (caffe_cpu_gemm -> cblas_sgemm RowMajor,
 * lda = transA ? M : K, ldb = transB ? K : N, ldc = N) and :57-62 (gemv).
 * The arithmetic itself lives in OpenBLAS (not under /root/reference, unpinned:
 * Makefile.config:49 "BLAS := open"); restated as the textbook sum over k in
 * ascending order.  acc64 != 0 accumulates in double (a tighter yardstick that
 * does not depend on any BLAS's summation order).
 * ------------------------------------------------------------------------- */
void b2o_gemm(int transA, int transB, int M, int N, int K, float alpha,
              const float* A, const float* B, float beta, float* C, int acc64) {
  const int lda = transA ? M : K, ldb = transB ? K : N;
  if (!acc64 && !transB) {
    /* i-k-j order: same per-element ascending-k order, but vectorisable */
    float* row = (float*)malloc(sizeof(float) * (size_t)N);
    for (int i = 0; i < M; ++i) {
      for (int j = 0; j < N; ++j) row[j] = 0.0f;
      for (int k = 0; k < K; ++k) {
        const float a = transA ? A[(size_t)k * lda + i] : A[(size_t)i * lda + k];
        const float* b = B + (size_t)k * ldb;
        for (int j = 0; j < N; ++j) row[j] += a * b[j];
      }
      float* c = C + (size_t)i * N;
      if (beta == 0.0f) for (int j = 0; j < N; ++j) c[j] = alpha * row[j];
      else for (int j = 0; j < N; ++j) c[j] = alpha * row[j] + beta * c[j];
    }
    free(row);
    return;
  }
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < N; ++j) {
      double sd = 0.0; float sf = 0.0f;
      for (int k = 0; k < K; ++k) {
        const float a = transA ? A[(size_t)k * lda + i] : A[(size_t)i * lda + k];
        const float b = transB ? B[(size_t)j * ldb + k] : B[(size_t)k * ldb + j];
        if (acc64) sd += (double)a * (double)b; else sf += a * b;
      }
      float* c = C + (size_t)i * N + j;
      if (acc64) *c = (float)((double)alpha * sd + (beta == 0.0f ? 0.0 : (double)beta * (double)*c));
      else *c = (beta == 0.0f) ? alpha * sf : alpha * sf + beta * *c;
    }
}

/* y = alpha*op(A)*x + beta*y, A is M x N row-major (math_functions.cpp:57-62). */
void b2o_gemv(int transA, int M, int N, float alpha, const float* A,
              const float* x, float beta, float* y, int acc64) {
  const int leny = transA ? N : M, lenx = transA ? M : N;
  for (int i = 0; i < leny; ++i) {
    double sd = 0.0; float sf = 0.0f;
    for (int k = 0; k < lenx; ++k) {
      const float a = transA ? A[(size_t)k * N + i] : A[(size_t)i * N + k];
      if (acc64) sd += (double)a * (double)x[k]; else sf += a * x[k];
    }
    const float s = acc64 ? (float)((double)alpha * sd) : alpha * sf;
    y[i] = (beta == 0.0f) ? s : s + beta * y[i];
  }
}

/* ------------------------------------------------------------------------- *
 * ConvolutionLayer::Forward_cpu (src/caffe/layers/conv_layer.cpp:24-40) with
 * forward_cpu_gemm / forward_cpu_bias (include/caffe/layers/base_conv_layer.hpp:36-60).
 * Per image: col = im2col(x_n) unless is_1x1_ (base_conv_layer.cpp:99-103),
 * per group y_{n,g} = W_g * col_g (beta=0), then y_n += bias * ones (rank-1 gemm).
 * ------------------------------------------------------------------------- */
static int is_1x1(const b2o_conv_params* p) {
  return p->kh == 1 && p->kw == 1 && p->sh == 1 && p->sw == 1 && p->ph == 0 && p->pw == 0;
}

void b2o_conv_forward(const b2o_conv_params* p, const float* x, const float* w,
                      const float* bias, float* y, int acc64) {
  const int Ho = b2o_out_h(p), Wo = b2o_out_w(p), P = Ho * Wo;
  const int Cg = p->C / p->G, Og = p->O / p->G, Kd = Cg * p->kh * p->kw;
  const size_t bottom_dim = (size_t)p->C * p->H * p->W, top_dim = (size_t)p->O * P;
  const size_t col_off = (size_t)Kd * P, out_off = (size_t)Og * P, w_off = (size_t)Og * Kd;
  float* col = is_1x1(p) ? NULL : (float*)malloc(sizeof(float) * (size_t)Kd * p->G * P);
  float* ones = (float*)malloc(sizeof(float) * (size_t)P);
  for (int i = 0; i < P; ++i) ones[i] = 1.0f;
  for (int n = 0; n < p->N; ++n) {
    const float* xn = x + n * bottom_dim;
    float* yn = y + n * top_dim;
    const float* cb = xn;
    if (col) { b2o_im2col(xn, p->C, p->H, p->W, p->kh, p->kw, p->ph, p->pw, p->sh, p->sw, p->dh, p->dw, col); cb = col; }
    for (int g = 0; g < p->G; ++g)
      b2o_gemm(0, 0, Og, P, Kd, 1.0f, w + w_off * g, cb + col_off * g, 0.0f, yn + out_off * g, acc64);
    if (p->has_bias && bias) b2o_gemm(0, 0, p->O, P, 1, 1.0f, bias, ones, 1.0f, yn, acc64);
  }
  free(ones);
  free(col);
}

/* ------------------------------------------------------------------------- *
 * ConvolutionLayer::Backward_cpu (conv_layer.cpp:42-73) with
 * backward_cpu_bias (base_conv_layer.hpp:96-100, gemv beta=1),
 * weight_cpu_gemm (:80-94, dW_g += dY_{n,g} * col_g^T, beta=1) and
 * backward_cpu_gemm (:62-78, col_g = W_g^T * dY_{n,g}, beta=0, then col2im).
 * dw/db are ACCUMULATED into; dx is OVERWRITTEN.  Null dw/db/dx skips that leg
 * (param_propagate_down_ / propagate_down).
 * ------------------------------------------------------------------------- */
void b2o_conv_backward(const b2o_conv_params* p, const float* x, const float* w,
                       const float* dy, float* dw, float* db, float* dx, int acc64) {
  const int Ho = b2o_out_h(p), Wo = b2o_out_w(p), P = Ho * Wo;
  const int Cg = p->C / p->G, Og = p->O / p->G, Kd = Cg * p->kh * p->kw;
  const size_t bottom_dim = (size_t)p->C * p->H * p->W, top_dim = (size_t)p->O * P;
  const size_t col_off = (size_t)Kd * P, out_off = (size_t)Og * P, w_off = (size_t)Og * Kd;
  const int one = is_1x1(p);
  float* col = one ? NULL : (float*)malloc(sizeof(float) * (size_t)Kd * p->G * P);
  float* ones = (float*)malloc(sizeof(float) * (size_t)P);
  for (int i = 0; i < P; ++i) ones[i] = 1.0f;
  if (p->has_bias && db)
    for (int n = 0; n < p->N; ++n)
      b2o_gemv(0, p->O, P, 1.0f, dy + n * top_dim, ones, 1.0f, db, acc64);
  if (dw || dx) {
    for (int n = 0; n < p->N; ++n) {
      const float* xn = x + n * bottom_dim;
      const float* dyn = dy + n * top_dim;
      if (dw) {
        const float* cb = xn;
        if (!one) { b2o_im2col(xn, p->C, p->H, p->W, p->kh, p->kw, p->ph, p->pw, p->sh, p->sw, p->dh, p->dw, col); cb = col; }
        for (int g = 0; g < p->G; ++g)
          b2o_gemm(0, 1, Og, Kd, P, 1.0f, dyn + out_off * g, cb + col_off * g, 1.0f, dw + w_off * g, acc64);
      }
      if (dx) {
        float* dxn = dx + n * bottom_dim;
        float* cb = one ? dxn : col;
        for (int g = 0; g < p->G; ++g)
          b2o_gemm(1, 0, Kd, P, Og, 1.0f, w + w_off * g, dyn + out_off * g, 0.0f, cb + col_off * g, acc64);
        if (!one) b2o_col2im(col, p->C, p->H, p->W, p->kh, p->kw, p->ph, p->pw, p->sh, p->sw, p->dh, p->dw, dxn);
      }
    }
  }
  free(ones);
  free(col);
}

/* ------------------------------------------------------------------------- *
 * Independent direct-definition convolution, the role caffe_conv plays in
 * src/caffe/test/test_convolution_layer.cpp:24-139 (2-D case): a second,
 * structurally different statement of the same function, used to cross-check
 * the im2col+GEMM restatement above.
 * ------------------------------------------------------------------------- */
void b2o_conv_direct(const b2o_conv_params* p, const float* x, const float* w,
                     const float* bias, float* y) {
  const int Ho = b2o_out_h(p), Wo = b2o_out_w(p);
  const int Cg = p->C / p->G, Og = p->O / p->G;
  for (int n = 0; n < p->N; ++n)
    for (int o = 0; o < p->O; ++o) {
      const int g = o / Og;
      for (int ho = 0; ho < Ho; ++ho)
        for (int wo = 0; wo < Wo; ++wo) {
          double acc = (p->has_bias && bias) ? (double)bias[o] : 0.0;
          for (int c = 0; c < Cg; ++c)
            for (int i = 0; i < p->kh; ++i)
              for (int j = 0; j < p->kw; ++j) {
                const int h = ho * p->sh - p->ph + i * p->dh, ww = wo * p->sw - p->pw + j * p->dw;
                if (h < 0 || h >= p->H || ww < 0 || ww >= p->W) continue;
                acc += (double)x[(((size_t)n * p->C + g * Cg + c) * p->H + h) * p->W + ww] *
                       (double)w[(((size_t)o * Cg + c) * p->kh + i) * p->kw + j];
              }
          y[(((size_t)n * p->O + o) * Ho + ho) * Wo + wo] = (float)acc;
        }
    }
}

/* ------------------------------------------------------------------------- *
 * SGD with momentum, CPU branch of SGDSolver::ApplyUpdate
 * (src/caffe/solvers/sgd_solver.cpp:143-149): Normalize (:152-158, x 1/iter_size),
 * Regularize (:161-192, L2: g += decay*w, L1: g += decay*sign(w)),
 * ComputeUpdateValue (:203-219: h = lr*g + m*h; g = h; Blob::Update w -= g
 * (src/caffe/blob.cpp:129-154); clear_grads -> g = 0).
 * grad_scale restates Net::ReduceAndUpdate's 1/global_grad_scale and the
 * 1/solver_count of Net::ReduceBucket (src/caffe/net.cpp:815-817,899-912): g is
 * multiplied by it first.  Pass 1 for the single-GPU fp32 case.
 * ------------------------------------------------------------------------- */
void b2o_sgd_update(size_t n, float* g, float* w, float* h, float momentum,
                    float local_rate, float local_decay, int l2, float grad_scale,
                    int iter_size, int clear_grads) {
  const float accum_norm = 1.0f / (float)iter_size;
  for (size_t i = 0; i < n; ++i) {
    float gi = g[i];
    if (grad_scale != 1.0f) gi *= grad_scale;
    if (iter_size != 1) gi *= accum_norm;
    if (local_decay != 0.0f) {
      const float reg = l2 ? w[i] : (float)((0.0f < w[i]) - (w[i] < 0.0f));
      gi += local_decay * reg;
    }
    const float hi = local_rate * gi + momentum * h[i];
    h[i] = hi;
    w[i] -= hi;
    g[i] = clear_grads ? 0.0f : hi;
  }
}

/* SGDSolver::GetLearningRate (sgd_solver.cpp:24-65).  policy: 0 fixed, 1 step,
 * 2 exp, 3 inv, 4 multistep (current_step passed in), 5 poly, 6 sigmoid.      */
float b2o_learning_rate(int policy, int iter, float base_lr, float gamma, float power,
                        int stepsize, int max_iter, float min_lr, int current_step,
                        int rampup_interval, float rampup_lr) {
  if (iter < rampup_interval) {
    const float alpha = (float)iter / (float)rampup_interval;
    return rampup_lr + (base_lr - rampup_lr) * alpha;
  }
  switch (policy) {
    case 0: return base_lr;
    case 1: return base_lr * powf(gamma, (float)(iter / stepsize));
    case 2: return base_lr * powf(gamma, (float)iter);
    case 3: return base_lr * powf(1.0f + gamma * (float)iter, -power);
    case 4: return base_lr * powf(gamma, (float)current_step);
    case 5: return min_lr + (base_lr - min_lr) * powf(1.0f - ((float)iter / (float)max_iter), power);
    case 6: return base_lr / (1.0f + (float)exp(-(double)gamma * (double)(iter - stepsize)));
    default: return -1.0f;
  }
}

/* ------------------------------------------------------------------------- *
 * Multi-rank gradient exchange: in-place sum over ranks of a diff bucket
 * (P2PSync::allreduce_bucket, src/caffe/parallel.cpp:245-253, ncclSum) followed
 * by x 1/solver_count (Net::ReduceBucket, src/caffe/net.cpp:899-912).  One
 * process, R replicas, summed on the host in rank order.
 * ------------------------------------------------------------------------- */
void b2o_allreduce_avg(int R, size_t count, float** bufs, int solver_count) {
  const float scale = 1.0f / (float)solver_count;
  for (size_t i = 0; i < count; ++i) {
    float s = 0.0f;
    for (int r = 0; r < R; ++r) s += bufs[r][i];
    s *= scale;
    for (int r = 0; r < R; ++r) bufs[r][i] = s;
  }
}

#ifdef __cplusplus
}
#endif
