// ref_driver.cpp -- C entry points over the reference's own im2col.cpp (compiled
// verbatim from /root/reference into oracle/_ref/libb2o_ref.so) plus the per-image
// / per-group ConvolutionLayer CPU loop driven through OpenBLAS's cblas_sgemm /
// cblas_sgemv (the reference's "BLAS := open", Makefile.config:49), resolved at
// run time with dlopen so the CPU baseline uses the same BLAS family as the
// reference build would.  TEST / BASELINE INFRASTRUCTURE ONLY.
#include <dlfcn.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "caffe/util/im2col.hpp"

extern "C" {

void ref_im2col_cpu(const float* im, int C, int H, int W, int kh, int kw, int ph, int pw,
                    int sh, int sw, int dh, int dw, float* col) {
  caffe::im2col_cpu<float>(im, C, H, W, kh, kw, ph, pw, sh, sw, dh, dw, col);
}
void ref_col2im_cpu(const float* col, int C, int H, int W, int kh, int kw, int ph, int pw,
                    int sh, int sw, int dh, int dw, float* im) {
  caffe::col2im_cpu<float>(col, C, H, W, kh, kw, ph, pw, sh, sw, dh, dw, im);
}
void ref_im2col_nd_cpu(const float* im, int nax, const int* im_shape, const int* col_shape,
                       const int* k, const int* pad, const int* stride, const int* dil, float* col) {
  caffe::im2col_nd_cpu<float>(im, nax, im_shape, col_shape, k, pad, stride, dil, col);
}
void ref_col2im_nd_cpu(const float* col, int nax, const int* im_shape, const int* col_shape,
                       const int* k, const int* pad, const int* stride, const int* dil, float* im) {
  caffe::col2im_nd_cpu<float>(col, nax, im_shape, col_shape, k, pad, stride, dil, im);
}

// ---- OpenBLAS via dlopen (CBLAS enums: RowMajor=101, NoTrans=111, Trans=112) ----
typedef void (*sgemm_fn)(int, int, int, int, int, int, float, const float*, int,
                         const float*, int, float, float*, int);
typedef void (*sgemv_fn)(int, int, int, int, float, const float*, int, const float*, int,
                         float, float*, int);
typedef void (*setthr_fn)(int);
static sgemm_fn p_sgemm = nullptr;
static sgemv_fn p_sgemv = nullptr;
static setthr_fn p_setthr = nullptr;

int ref_blas_open(const char* path, int threads) {
  void* h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!h) { std::fprintf(stderr, "ref_blas_open: %s\n", dlerror()); return -1; }
  p_sgemm = (sgemm_fn)dlsym(h, "cblas_sgemm");
  p_sgemv = (sgemv_fn)dlsym(h, "cblas_sgemv");
  p_setthr = (setthr_fn)dlsym(h, "openblas_set_num_threads");
  if (!p_sgemm || !p_sgemv) return -2;
  if (p_setthr && threads > 0) p_setthr(threads);
  return 0;
}

static void gemm(int tA, int tB, int M, int N, int K, float alpha, const float* A,
                 const float* B, float beta, float* C) {
  const int lda = tA ? M : K, ldb = tB ? K : N;
  p_sgemm(101, tA ? 112 : 111, tB ? 112 : 111, M, N, K, alpha, A, lda, B, ldb, beta, C, N);
}

struct ref_conv_params { int N, C, H, W, O, G, kh, kw, sh, sw, ph, pw, dh, dw, has_bias; };

static int oext(int in, int k, int s, int p, int d) { return (in + 2 * p - (d * (k - 1) + 1)) / s + 1; }

// Forward + (optionally) backward of one conv layer the way the reference CPU
// layer does it: per image, per group, im2col + sgemm (+ rank-1 bias gemm / gemv).
int ref_conv_fwd_bwd(const ref_conv_params* p, const float* x, const float* w, const float* bias,
                     float* y, const float* dy, float* dw, float* db, float* dx) {
  if (!p_sgemm) return -1;
  const int Ho = oext(p->H, p->kh, p->sh, p->ph, p->dh), Wo = oext(p->W, p->kw, p->sw, p->pw, p->dw);
  const int P = Ho * Wo, Cg = p->C / p->G, Og = p->O / p->G, Kd = Cg * p->kh * p->kw;
  const size_t bdim = (size_t)p->C * p->H * p->W, tdim = (size_t)p->O * P;
  const size_t coff = (size_t)Kd * P, ooff = (size_t)Og * P, woff = (size_t)Og * Kd;
  const bool one = p->kh == 1 && p->kw == 1 && p->sh == 1 && p->sw == 1 && p->ph == 0 && p->pw == 0;
  std::vector<float> col(one ? 0 : (size_t)Kd * p->G * P), ones(P, 1.0f);
  for (int n = 0; n < p->N; ++n) {
    const float* xn = x + n * bdim;
    if (y) {
      const float* cb = xn;
      if (!one) { ref_im2col_cpu(xn, p->C, p->H, p->W, p->kh, p->kw, p->ph, p->pw, p->sh, p->sw, p->dh, p->dw, col.data()); cb = col.data(); }
      for (int g = 0; g < p->G; ++g) gemm(0, 0, Og, P, Kd, 1.f, w + woff * g, cb + coff * g, 0.f, y + n * tdim + ooff * g);
      if (p->has_bias && bias) gemm(0, 0, p->O, P, 1, 1.f, bias, ones.data(), 1.f, y + n * tdim);
    }
  }
  if (dy) {
    if (p->has_bias && db)
      for (int n = 0; n < p->N; ++n)
        p_sgemv(101, 111, p->O, P, 1.f, dy + n * tdim, P, ones.data(), 1, 1.f, db, 1);
    for (int n = 0; n < p->N; ++n) {
      const float* xn = x + n * bdim;
      const float* dyn = dy + n * tdim;
      if (dw) {
        const float* cb = xn;
        if (!one) { ref_im2col_cpu(xn, p->C, p->H, p->W, p->kh, p->kw, p->ph, p->pw, p->sh, p->sw, p->dh, p->dw, col.data()); cb = col.data(); }
        for (int g = 0; g < p->G; ++g) gemm(0, 1, Og, Kd, P, 1.f, dyn + ooff * g, cb + coff * g, 1.f, dw + woff * g);
      }
      if (dx) {
        float* cb = one ? dx + n * bdim : col.data();
        for (int g = 0; g < p->G; ++g) gemm(1, 0, Kd, P, Og, 1.f, w + woff * g, dyn + ooff * g, 0.f, cb + coff * g);
        if (!one) ref_col2im_cpu(col.data(), p->C, p->H, p->W, p->kh, p->kw, p->ph, p->pw, p->sh, p->sw, p->dh, p->dw, dx + n * bdim);
      }
    }
  }
  return 0;
}

}  // extern "C"
