// fake_kernels_conv.cpp -- host stand-ins for the libb2c entry points a ResNet-style net needs beyond fake_kernels.cpp's:
// the convolution descriptor API and its four passes, BatchNorm in its plain / fused / residual-tail forms, ReLU, pooling,
// Eltwise.  Written from the contracts in include/b2c.h (what is overwritten, what is accumulated, what the fused forms stand
// for), with double accumulation, and enqueued on tests/sim/fake_cuda.cpp's streams like the real launches.  With them
// TrainNet -- the fusion pass, the fan-out shadow diffs, the deferred adds, iter_size accumulation, the update on its side
// stream -- runs whole training steps of small nets on the CPU under the stream-order model and is held against
// tests/netoracle.py (tests/test_trainer_sim.py).  TEST INFRASTRUCTURE ONLY; the real kernels are checked on hardware.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <functional>
#include <vector>

#include "../../include/b2c.h"

void fakecuda_launch(cudaStream_t st, std::function<void()> fn);
static cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }

struct b2c_conv_desc {             // the library's descriptor is opaque to callers; this translation unit supplies its own
  b2c_conv_params p;
  int Ho, Wo;
};

namespace {
inline size_t X(const b2c_conv_params& p, int n, int c, int h, int w) { return (((size_t)n * p.C + c) * p.H + h) * p.W + w; }
inline size_t Y(const b2c_conv_desc& d, int n, int o, int h, int w) { return (((size_t)n * d.p.O + o) * d.Ho + h) * d.Wo + w; }
inline size_t Wt(const b2c_conv_params& p, int o, int c, int i, int j) { return (((size_t)o * (p.C / p.G) + c) * p.kh + i) * p.kw + j; }

// per-channel statistics of x [N, C, S]: mean, 1 / sqrt(var + eps)  (batch_norm_layer.cpp:140-198; eps goes in before the root)
void bn_stats(int N, int C, int Sp, const float* x, float eps, std::vector<float>* mean, std::vector<float>* invstd, std::vector<float>* var_eps) {
  mean->resize(C); invstd->resize(C); var_eps->resize(C);
  const double cnt = (double)N * Sp;
  for (int c = 0; c < C; ++c) {
    double s = 0;
    for (int n = 0; n < N; ++n) for (int i = 0; i < Sp; ++i) s += x[((size_t)n * C + c) * Sp + i];
    const float m = (float)(s / cnt);
    double q = 0;
    for (int n = 0; n < N; ++n) for (int i = 0; i < Sp; ++i) { const double dxm = x[((size_t)n * C + c) * Sp + i] - m; q += dxm * dxm; }
    const float ve = (float)(q / cnt) + eps;
    (*mean)[c] = m; (*var_eps)[c] = ve; (*invstd)[c] = (float)(1.0 / std::sqrt((double)ve));
  }
}
void bn_forward_common(int N, int C, int Sp, const float* x, const float* gamma, const float* beta, float eps, float maf, int first,
                       float* run_mean, float* run_var, float* save_mean, float* save_invstd, float* xnorm, float* y, const float* residual,
                       int relu) {
  std::vector<float> mean, invstd, ve;
  bn_stats(N, C, Sp, x, eps, &mean, &invstd, &ve);
  for (int c = 0; c < C; ++c) {
    save_mean[c] = mean[c]; save_invstd[c] = invstd[c];
    if (run_mean) run_mean[c] = first ? mean[c] : (1.f - maf) * mean[c] + maf * run_mean[c];
    if (run_var) run_var[c] = first ? ve[c] : (1.f - maf) * ve[c] + maf * run_var[c];
  }
  for (int n = 0; n < N; ++n)
    for (int c = 0; c < C; ++c)
      for (int i = 0; i < Sp; ++i) {
        const size_t at = ((size_t)n * C + c) * Sp + i;
        const float xn = (x[at] - mean[c]) * invstd[c];
        if (xnorm) xnorm[at] = xn;
        float v = gamma ? xn * gamma[c] + beta[c] : xn;
        if (residual) v += residual[at];
        y[at] = relu && v < 0.f ? 0.f : v;
      }
}
// BatchNormLayer::Backward_cpu (batch_norm_layer.cpp:234-283): dgamma / dbeta OVERWRITTEN
void bn_backward_common(int N, int C, int Sp, const float* dy, const float* xnorm, const float* gamma, const float* invstd, float* dgamma,
                        float* dbeta, float* dx) {
  const double cnt = (double)N * Sp;
  for (int c = 0; c < C; ++c) {
    double dg = 0, db = 0;
    for (int n = 0; n < N; ++n)
      for (int i = 0; i < Sp; ++i) { const size_t at = ((size_t)n * C + c) * Sp + i; dg += (double)dy[at] * xnorm[at]; db += dy[at]; }
    dgamma[c] = (float)dg; dbeta[c] = (float)db;
    const double g = gamma ? gamma[c] : 1.0;
    for (int n = 0; n < N; ++n)
      for (int i = 0; i < Sp; ++i) {
        const size_t at = ((size_t)n * C + c) * Sp + i;
        dx[at] = (float)(g * invstd[c] * (dy[at] - db / cnt - xnorm[at] * (dg / cnt)));
      }
  }
}
int pooled(int in, int k, int s, int p) {                 // PoolingLayer::Reshape: ceil mode + last-window clip
  int o = (int)std::ceil((float)(in + 2 * p - k) / s) + 1;
  if (p > 0 && (o - 1) * s >= in + p) --o;
  return o;
}
}  // namespace

extern "C" {

// ---- convolution ----------------------------------------------------------------------------------------------------------------------------
int b2c_conv_desc_create(const b2c_conv_params* p, int, b2c_conv_desc** out) {
  if (!p || !out || p->N <= 0 || p->C <= 0 || p->O <= 0 || p->kh <= 0 || p->kw <= 0 || p->sh <= 0 || p->sw <= 0 || p->G <= 0 || p->C % p->G || p->O % p->G) return B2C_ERR_INVALID;
  b2c_conv_desc* d = new b2c_conv_desc;
  d->p = *p;
  d->Ho = (p->H + 2 * p->ph - (p->dh * (p->kh - 1) + 1)) / p->sh + 1;
  d->Wo = (p->W + 2 * p->pw - (p->dw * (p->kw - 1) + 1)) / p->sw + 1;
  *out = d;
  return B2C_OK;
}
int b2c_conv_desc_destroy(b2c_conv_desc* d) { delete d; return B2C_OK; }
int b2c_conv_desc_set_math(b2c_conv_desc*, int) { return B2C_OK; }
int b2c_conv_desc_set_algo(b2c_conv_desc*, int) { return B2C_OK; }
int b2c_conv_out_shape(const b2c_conv_desc* d, int* Ho, int* Wo) { if (Ho) *Ho = d->Ho; if (Wo) *Wo = d->Wo; return B2C_OK; }
size_t b2c_conv_workspace_bytes(const b2c_conv_desc*, int) { return 0; }
int b2c_conv_algo_used(const b2c_conv_desc*, int) { return B2C_ALGO_SIMT; }
size_t b2c_conv_filter_cache_bytes(const b2c_conv_desc*) { return 0; }
int b2c_conv_desc_bind_filter_cache(b2c_conv_desc*, const void*) { return B2C_OK; }
int b2c_conv_prepare_filters(int, const b2c_conv_desc* const*, const float* const*, void* const*, void*) { return B2C_OK; }
int b2c_conv_backward_data_accumulate_supported(const b2c_conv_desc*) { return 1; }

int b2c_conv_forward(const b2c_conv_desc* d, const float* x, const float* w, const float* bias, float* y, void*, size_t, void* stream) {
  const b2c_conv_desc D = *d;
  fakecuda_launch(S(stream), [=] {
    const b2c_conv_params& p = D.p;
    const int Cg = p.C / p.G, Og = p.O / p.G;
    for (int n = 0; n < p.N; ++n)
      for (int o = 0; o < p.O; ++o)
        for (int ho = 0; ho < D.Ho; ++ho)
          for (int wo = 0; wo < D.Wo; ++wo) {
            double acc = bias && p.has_bias ? bias[o] : 0.0;
            const int g = o / Og;
            for (int c = 0; c < Cg; ++c)
              for (int i = 0; i < p.kh; ++i)
                for (int j = 0; j < p.kw; ++j) {
                  const int h = ho * p.sh - p.ph + i * p.dh, ww = wo * p.sw - p.pw + j * p.dw;
                  if (h < 0 || h >= p.H || ww < 0 || ww >= p.W) continue;
                  acc += (double)x[X(p, n, g * Cg + c, h, ww)] * w[Wt(p, o, c, i, j)];
                }
            y[Y(D, n, o, ho, wo)] = (float)acc;
          }
  });
  return B2C_OK;
}
static int conv_dgrad(const b2c_conv_desc* d, const float* dy, const float* w, float* dx, bool accumulate, void* stream) {
  const b2c_conv_desc D = *d;
  fakecuda_launch(S(stream), [=] {
    const b2c_conv_params& p = D.p;
    const int Cg = p.C / p.G, Og = p.O / p.G;
    std::vector<double> acc((size_t)p.N * p.C * p.H * p.W, 0.0);
    for (int n = 0; n < p.N; ++n)
      for (int o = 0; o < p.O; ++o)
        for (int ho = 0; ho < D.Ho; ++ho)
          for (int wo = 0; wo < D.Wo; ++wo) {
            const double g = dy[Y(D, n, o, ho, wo)];
            const int grp = o / Og;
            for (int c = 0; c < Cg; ++c)
              for (int i = 0; i < p.kh; ++i)
                for (int j = 0; j < p.kw; ++j) {
                  const int h = ho * p.sh - p.ph + i * p.dh, ww = wo * p.sw - p.pw + j * p.dw;
                  if (h < 0 || h >= p.H || ww < 0 || ww >= p.W) continue;
                  acc[X(p, n, grp * Cg + c, h, ww)] += g * w[Wt(p, o, c, i, j)];
                }
          }
    for (size_t i = 0; i < acc.size(); ++i) dx[i] = (float)((accumulate ? (double)dx[i] : 0.0) + acc[i]);
  });
  return B2C_OK;
}
int b2c_conv_backward_data(const b2c_conv_desc* d, const float* dy, const float* w, float* dx, void*, size_t, void* stream) {
  return conv_dgrad(d, dy, w, dx, false, stream);
}
int b2c_conv_backward_data_accumulate(const b2c_conv_desc* d, const float* dy, const float* w, float* dx, void*, size_t, void* stream) {
  return conv_dgrad(d, dy, w, dx, true, stream);
}
int b2c_conv_backward_filter(const b2c_conv_desc* d, const float* x, const float* dy, float* dw, void*, size_t, void* stream) {
  const b2c_conv_desc D = *d;
  fakecuda_launch(S(stream), [=] {
    const b2c_conv_params& p = D.p;
    const int Cg = p.C / p.G, Og = p.O / p.G;
    for (int o = 0; o < p.O; ++o)
      for (int c = 0; c < Cg; ++c)
        for (int i = 0; i < p.kh; ++i)
          for (int j = 0; j < p.kw; ++j) {
            double acc = 0;
            const int g = o / Og;
            for (int n = 0; n < p.N; ++n)
              for (int ho = 0; ho < D.Ho; ++ho)
                for (int wo = 0; wo < D.Wo; ++wo) {
                  const int h = ho * p.sh - p.ph + i * p.dh, ww = wo * p.sw - p.pw + j * p.dw;
                  if (h < 0 || h >= p.H || ww < 0 || ww >= p.W) continue;
                  acc += (double)x[X(p, n, g * Cg + c, h, ww)] * dy[Y(D, n, o, ho, wo)];
                }
            float& out = dw[Wt(p, o, c, i, j)];
            out = (float)(out + acc);                       // dW += (beta = 1)
          }
  });
  return B2C_OK;
}
int b2c_conv_backward_bias(const b2c_conv_desc* d, const float* dy, float* db, void* stream) {
  const b2c_conv_desc D = *d;
  fakecuda_launch(S(stream), [=] {
    for (int o = 0; o < D.p.O; ++o) {
      double acc = 0;
      for (int n = 0; n < D.p.N; ++n) for (int h = 0; h < D.Ho; ++h) for (int w = 0; w < D.Wo; ++w) acc += dy[Y(D, n, o, h, w)];
      db[o] = (float)(db[o] + acc);
    }
  });
  return B2C_OK;
}

// ---- ReLU / Eltwise -----------------------------------------------------------------------------------------------------------------------------
int b2c_relu_forward(size_t n, const float* x, float* y, float slope, void* stream) {
  fakecuda_launch(S(stream), [=] { for (size_t i = 0; i < n; ++i) y[i] = x[i] > 0.f ? x[i] : slope * x[i]; });
  return B2C_OK;
}
int b2c_relu_backward(size_t n, const float* dy, const float* x, float* dx, float slope, void* stream) {
  fakecuda_launch(S(stream), [=] { for (size_t i = 0; i < n; ++i) dx[i] = dy[i] * (x[i] > 0.f ? 1.f : slope); });
  return B2C_OK;
}
int b2c_add(size_t n, const float* a, const float* b, float* y, void* stream) {
  fakecuda_launch(S(stream), [=] { for (size_t i = 0; i < n; ++i) y[i] = a[i] + b[i]; });
  return B2C_OK;
}
int b2c_add_relu(size_t n, const float* a, const float* b, float* y, void* stream) {
  fakecuda_launch(S(stream), [=] { for (size_t i = 0; i < n; ++i) { const float v = a[i] + b[i]; y[i] = v > 0.f ? v : 0.f; } });
  return B2C_OK;
}
int b2c_relu_backward2(size_t n, const float* dy, const float* y, float* dx_a, float* dx_b, void* stream) {
  fakecuda_launch(S(stream), [=] {
    for (size_t i = 0; i < n; ++i) { const float g = y[i] > 0.f ? dy[i] : 0.f; if (dx_a) dx_a[i] = g; if (dx_b) dx_b[i] = g; }
  });
  return B2C_OK;
}

// ---- BatchNorm ---------------------------------------------------------------------------------------------------------------------------------------
int b2c_bn_forward_train(int N, int C, int Sp, const float* x, const float* gamma, const float* beta, float eps, float maf, int first,
                         float* run_mean, float* run_var, float* save_mean, float* save_invstd, float* xnorm, float* y, void* stream) {
  fakecuda_launch(S(stream), [=] { bn_forward_common(N, C, Sp, x, gamma, beta, eps, maf, first, run_mean, run_var, save_mean, save_invstd, xnorm, y, nullptr, 0); });
  return B2C_OK;
}
int b2c_bn_forward_train_fused(int N, int C, int Sp, const float* x, const float* gamma, const float* beta, float eps, float maf, int first,
                               float* run_mean, float* run_var, float* save_mean, float* save_invstd, float* y, int relu, void* stream) {
  fakecuda_launch(S(stream), [=] { bn_forward_common(N, C, Sp, x, gamma, beta, eps, maf, first, run_mean, run_var, save_mean, save_invstd, nullptr, y, nullptr, relu); });
  return B2C_OK;
}
int b2c_bn_forward_train_fused_res(int N, int C, int Sp, const float* x, const float* gamma, const float* beta, float eps, float maf, int first,
                                   float* run_mean, float* run_var, float* save_mean, float* save_invstd, const float* residual, float* y,
                                   int relu, void* stream) {
  fakecuda_launch(S(stream), [=] { bn_forward_common(N, C, Sp, x, gamma, beta, eps, maf, first, run_mean, run_var, save_mean, save_invstd, nullptr, y, residual, relu); });
  return B2C_OK;
}
int b2c_bn_backward(int N, int C, int Sp, const float* dy, const float* xnorm, const float* gamma, const float* save_invstd, float* dgamma,
                    float* dbeta, float* dx, void* stream) {
  fakecuda_launch(S(stream), [=] { bn_backward_common(N, C, Sp, dy, xnorm, gamma, save_invstd, dgamma, dbeta, dx); });
  return B2C_OK;
}
// x_norm recomputed from the layer input and the saved statistics; with relu, dy is first masked by (BatchNorm output > 0)
int b2c_bn_backward_fused(int N, int C, int Sp, const float* dy, const float* x, const float* save_mean, const float* save_invstd,
                          const float* gamma, const float* beta, float* dgamma, float* dbeta, float* dx, int relu, void* stream) {
  fakecuda_launch(S(stream), [=] {
    const size_t total = (size_t)N * C * Sp;
    std::vector<float> xn(total), d(total);
    for (int n = 0; n < N; ++n)
      for (int c = 0; c < C; ++c)
        for (int i = 0; i < Sp; ++i) {
          const size_t at = ((size_t)n * C + c) * Sp + i;
          xn[at] = (x[at] - save_mean[c]) * save_invstd[c];
          const float yv = gamma ? xn[at] * gamma[c] + beta[c] : xn[at];
          d[at] = relu && !(yv > 0.f) ? 0.f : dy[at];
        }
    bn_backward_common(N, C, Sp, d.data(), xn.data(), gamma, save_invstd, dgamma, dbeta, dx);
  });
  return B2C_OK;
}
// residual tail: the sum's top diff (in up to two parts) masked by the sum's post-ReLU data goes to the other bottom and through BatchNorm
int b2c_bn_backward_fused_res(int N, int C, int Sp, const float* d_sum, const float* d_sum2, const float* y_sum, const float* x,
                              const float* save_mean, const float* save_invstd, const float* gamma, const float* beta, float* dgamma,
                              float* dbeta, float* dx, float* d_residual, void* stream) {
  (void)beta;
  fakecuda_launch(S(stream), [=] {
    const size_t total = (size_t)N * C * Sp;
    std::vector<float> xn(total), d(total);
    for (int n = 0; n < N; ++n)
      for (int c = 0; c < C; ++c)
        for (int i = 0; i < Sp; ++i) {
          const size_t at = ((size_t)n * C + c) * Sp + i;
          xn[at] = (x[at] - save_mean[c]) * save_invstd[c];
          const float g = d_sum[at] + (d_sum2 ? d_sum2[at] : 0.f);
          d[at] = y_sum[at] > 0.f ? g : 0.f;
        }
    if (d_residual) std::memcpy(d_residual, d.data(), sizeof(float) * total);
    bn_backward_common(N, C, Sp, d.data(), xn.data(), gamma, save_invstd, dgamma, dbeta, dx);
  });
  return B2C_OK;
}

// ---- Pooling (pooling_layer.cpp:129-318): MAX keeps the first maximum's plane index in `mask`; AVE divides by the padded window ---------------------
int b2c_pool_forward(int method, int NC, int H, int W, int kh, int kw, int sh, int sw, int ph, int pw, const float* x, float* y, int* mask,
                     void* stream) {
  fakecuda_launch(S(stream), [=] {
    const int Ho = pooled(H, kh, sh, ph), Wo = pooled(W, kw, sw, pw);
    for (int nc = 0; nc < NC; ++nc)
      for (int ho = 0; ho < Ho; ++ho)
        for (int wo = 0; wo < Wo; ++wo) {
          int hs = ho * sh - ph, ws = wo * sw - pw;
          const size_t out = ((size_t)nc * Ho + ho) * Wo + wo;
          if (method == 0) {
            const int he = std::min(hs + kh, H), we = std::min(ws + kw, W);
            hs = std::max(hs, 0); ws = std::max(ws, 0);
            float best = -3.402823466e38f;
            int arg = -1;
            for (int h = hs; h < he; ++h)
              for (int w = ws; w < we; ++w) { const float v = x[((size_t)nc * H + h) * W + w]; if (v > best) { best = v; arg = h * W + w; } }
            y[out] = best;
            if (mask) mask[out] = arg;
          } else {
            int he = std::min(hs + kh, H + ph), we = std::min(ws + kw, W + pw);
            const int size = (he - hs) * (we - ws);
            hs = std::max(hs, 0); ws = std::max(ws, 0); he = std::min(he, H); we = std::min(we, W);
            double acc = 0;
            for (int h = hs; h < he; ++h) for (int w = ws; w < we; ++w) acc += x[((size_t)nc * H + h) * W + w];
            y[out] = (float)(acc / size);
          }
        }
  });
  return B2C_OK;
}
int b2c_pool_backward(int method, int NC, int H, int W, int kh, int kw, int sh, int sw, int ph, int pw, const float* dy, const int* mask,
                      float* dx, void* stream) {
  fakecuda_launch(S(stream), [=] {
    const int Ho = pooled(H, kh, sh, ph), Wo = pooled(W, kw, sw, pw);
    std::vector<double> acc((size_t)NC * H * W, 0.0);
    for (int nc = 0; nc < NC; ++nc)
      for (int ho = 0; ho < Ho; ++ho)
        for (int wo = 0; wo < Wo; ++wo) {
          const size_t out = ((size_t)nc * Ho + ho) * Wo + wo;
          if (method == 0) {
            acc[(size_t)nc * H * W + mask[out]] += dy[out];
          } else {
            int hs = ho * sh - ph, ws = wo * sw - pw;
            int he = std::min(hs + kh, H + ph), we = std::min(ws + kw, W + pw);
            const int size = (he - hs) * (we - ws);
            hs = std::max(hs, 0); ws = std::max(ws, 0); he = std::min(he, H); we = std::min(we, W);
            for (int h = hs; h < he; ++h) for (int w = ws; w < we; ++w) acc[((size_t)nc * H + h) * W + w] += (double)dy[out] / size;
          }
        }
    for (size_t i = 0; i < acc.size(); ++i) dx[i] = (float)acc[i];
  });
  return B2C_OK;
}

}  // extern "C"

// ---- the layers AlexNet / GoogLeNet / VGG-16 add: LRN across channels, Dropout, Accuracy ------------------------------------------------------------
extern "C" {

int b2c_lrn_forward(int N, int C, int Sp, int size, float alpha, float beta, float k, const float* x, float* scale, float* y, void* stream) {
  fakecuda_launch(S(stream), [=] {
    const int pre = (size - 1) / 2;
    for (int n = 0; n < N; ++n)
      for (int c = 0; c < C; ++c)
        for (int i = 0; i < Sp; ++i) {
          double acc = 0;
          for (int cc = std::max(c - pre, 0); cc < std::min(c - pre + size, C); ++cc) { const double v = x[((size_t)n * C + cc) * Sp + i]; acc += v * v; }
          const size_t at = ((size_t)n * C + c) * Sp + i;
          const double sc = k + acc * alpha / size;
          scale[at] = (float)sc;
          y[at] = (float)(x[at] * std::pow(sc, -(double)beta));
        }
  });
  return B2C_OK;
}
int b2c_lrn_backward(int N, int C, int Sp, int size, float alpha, float beta, const float* x, const float* y, const float* scale, const float* dy,
                     float* dx, void* stream) {
  fakecuda_launch(S(stream), [=] {
    const int ipp = size - (size + 1) / 2;                    // inverse_pre_pad, lrn_layer.cpp CrossChannelBackward_cpu
    for (int n = 0; n < N; ++n)
      for (int c = 0; c < C; ++c)
        for (int i = 0; i < Sp; ++i) {
          const size_t at = ((size_t)n * C + c) * Sp + i;
          double acc = 0;
          for (int cc = std::max(c - ipp, 0); cc < std::min(c - ipp + size, C); ++cc) {
            const size_t o = ((size_t)n * C + cc) * Sp + i;
            acc += (double)dy[o] * y[o] / scale[o];
          }
          dx[at] = (float)(dy[at] * std::pow((double)scale[at], -(double)beta) - 2.0 * alpha * beta / size * x[at] * acc);
        }
  });
  return B2C_OK;
}
int b2c_dropout_mask(size_t n, float ratio, unsigned long long seed, unsigned long long offset, float* mask, void* stream) {
  fakecuda_launch(S(stream), [=] {
    const unsigned thr = (unsigned)(int)((double)ratio * 16777216.0);
    const float keep = 1.f / (1.f - ratio);
    for (size_t i = 0; i < n; ++i) {
      unsigned long long z = seed + offset + i + 0x9E3779B97F4A7C15ull;          // splitmix64
      z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
      z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
      z ^= z >> 31;
      mask[i] = (unsigned)(z >> 40) >= thr ? keep : 0.f;
    }
  });
  return B2C_OK;
}
int b2c_mul(size_t n, const float* a, const float* b, float* y, void* stream) {
  fakecuda_launch(S(stream), [=] { for (size_t i = 0; i < n; ++i) y[i] = a[i] * b[i]; });
  return B2C_OK;
}
// the label must be among the top_k of (score, class) pairs ordered by std::greater (accuracy_layer.cpp:44-100)
int b2c_accuracy(int N, int C, int top_k, const float* scores, const float* labels, float* accuracy, void*, void* stream) {
  fakecuda_launch(S(stream), [=] {
    int hits = 0;
    for (int n = 0; n < N; ++n) {
      const int lab = (int)labels[n];
      const float s = scores[(size_t)n * C + lab];
      int better = 0;
      for (int c = 0; c < C; ++c) { const float v = scores[(size_t)n * C + c]; if (v > s || (v == s && c > lab)) ++better; }
      hits += better < top_k;
    }
    *accuracy = (float)hits / N;
  });
  return B2C_OK;
}

}  // extern "C"

// ---- N-D im2col / col2im (im2col_nd_cpu / col2im_nd_cpu, src/caffe/util/im2col.cpp:76-174): the host layer's N-D convolution path calls
// these directly (b2caffe.cpp), per image, around b2c_sgemm.  Shape arrays are HOST arrays: im_shape [C, d0..], col_shape [C*prod(k), o0..].
namespace {
void nd_core(bool to_col, const float* src, int num_axes, std::vector<int> ims, std::vector<int> cols, std::vector<int> k, std::vector<int> pad,
             std::vector<int> stride, std::vector<int> dil, float* dst) {
  size_t im_size = ims[0], kernel_size = 1, out_spatial = 1;
  for (int i = 0; i < num_axes; ++i) { im_size *= (size_t)ims[1 + i]; kernel_size *= (size_t)k[i]; out_spatial *= (size_t)cols[1 + i]; }
  const int channels_col = cols[0];
  if (!to_col) for (size_t i = 0; i < im_size; ++i) dst[i] = 0.f;
  std::vector<int> d_off(num_axes), d_iter(num_axes);
  for (int c_col = 0; c_col < channels_col; ++c_col) {
    int offset = c_col;                                           // which kernel tap this column row is
    for (int d = num_axes - 1; d >= 0; --d) { if (d < num_axes - 1) offset /= k[d + 1]; d_off[d] = offset % k[d]; }
    const int c_im = c_col / (int)kernel_size;
    std::fill(d_iter.begin(), d_iter.end(), 0);
    for (size_t o = 0; o < out_spatial; ++o) {
      size_t idx_im = (size_t)c_im;
      bool padded = false;
      for (int d = 0; d < num_axes; ++d) {
        const int pos = d_iter[d] * stride[d] - pad[d] + d_off[d] * dil[d];
        padded |= pos < 0 || pos >= ims[1 + d];
        idx_im = idx_im * (size_t)ims[1 + d] + (size_t)(pos < 0 ? 0 : pos);
      }
      const size_t idx_col = (size_t)c_col * out_spatial + o;
      if (to_col) dst[idx_col] = padded ? 0.f : src[idx_im];
      else if (!padded) dst[idx_im] += src[idx_col];
      for (int d = num_axes - 1; d >= 0; --d) { if (++d_iter[d] < cols[1 + d]) break; d_iter[d] = 0; }
    }
  }
}
}  // namespace

extern "C" {
int b2c_im2col_nd(const float* im, int num_axes, const int* im_shape, const int* col_shape, const int* kernel, const int* pad, const int* stride,
                  const int* dilation, float* col, void* stream) {
  std::vector<int> a(im_shape, im_shape + num_axes + 1), b(col_shape, col_shape + num_axes + 1), k(kernel, kernel + num_axes), p(pad, pad + num_axes),
      s(stride, stride + num_axes), d(dilation, dilation + num_axes);
  fakecuda_launch(S(stream), [=] { nd_core(true, im, num_axes, a, b, k, p, s, d, col); });
  return B2C_OK;
}
int b2c_col2im_nd(const float* col, int num_axes, const int* im_shape, const int* col_shape, const int* kernel, const int* pad, const int* stride,
                  const int* dilation, float* im, void* stream) {
  std::vector<int> a(im_shape, im_shape + num_axes + 1), b(col_shape, col_shape + num_axes + 1), k(kernel, kernel + num_axes), p(pad, pad + num_axes),
      s(stride, stride + num_axes), d(dilation, dilation + num_axes);
  fakecuda_launch(S(stream), [=] { nd_core(false, col, num_axes, a, b, k, p, s, d, im); });
  return B2C_OK;
}
}  // extern "C"
