// fake_kernels.cpp -- host stand-ins for the handful of libb2c kernels a fully-connected net needs (InnerProduct, SoftmaxWithLoss,
// the SGD update), enqueued on tests/sim/fake_cuda.cpp's streams like the real launches are.  Linked into libtrainsim.so so that
// TrainNet::Step -- data layer, forward, backward, the update on its side stream, the event hand-overs between them -- can run END
// TO END on the stream-order model for small nets (tests/test_trainer_sim.py).  TEST INFRASTRUCTURE ONLY: double-accumulating
// loops written from the entry points' contracts in include/b2c.h; the real kernels are checked against the oracle on hardware.
// Every other b2c_* compute entry point resolves to the real libb2c.so, which refuses to launch without a device.
#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <functional>
#include <vector>

#include "../../include/b2c.h"

void fakecuda_launch(cudaStream_t st, std::function<void()> fn);
static cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }

extern "C" {

int b2c_sgemm_tc_supported(int, int, int, int, int) { return 0; }          // keep InnerProduct on the plain b2c_sgemm calls
size_t b2c_sgemm_workspace_bytes(int, int, int, int, int) { return 0; }

// row-major C[M x N] = alpha * op(A) * op(B) + beta * C   (caffe_gpu_gemm's semantics, math_functions.cu:11-26)
int b2c_sgemm(int tA, int tB, int M, int N, int K, float alpha, const float* A, const float* B, float beta, float* Cm, void* stream) {
  fakecuda_launch(S(stream), [=] {
    for (int i = 0; i < M; ++i)
      for (int j = 0; j < N; ++j) {
        double acc = 0;
        for (int k = 0; k < K; ++k) acc += (double)(tA ? A[(size_t)k * M + i] : A[(size_t)i * K + k]) * (double)(tB ? B[(size_t)j * K + k] : B[(size_t)k * N + j]);
        float& c = Cm[(size_t)i * N + j];
        c = (float)(alpha * acc + (beta == 0.f ? 0.0 : (double)beta * c));
      }
  });
  return B2C_OK;
}
int b2c_sgemm_ex(int tA, int tB, int M, int N, int K, float alpha, const float* A, const float* B, float beta, float* Cm, void*, size_t,
                 void* stream) {
  return b2c_sgemm(tA, tB, M, N, K, alpha, A, B, beta, Cm, stream);
}
int b2c_transpose(int rows, int cols, const float* src, float* dst, void* stream) {
  fakecuda_launch(S(stream), [=] { for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c) dst[(size_t)c * rows + r] = src[(size_t)r * cols + c]; });
  return B2C_OK;
}
int b2c_bias_forward(int N, int O, int P, const float* bias, float* y, void* stream) {
  fakecuda_launch(S(stream), [=] { for (int n = 0; n < N; ++n) for (int o = 0; o < O; ++o) for (int p = 0; p < P; ++p) y[((size_t)n * O + o) * P + p] += bias[o]; });
  return B2C_OK;
}
int b2c_bias_backward(int N, int O, int P, const float* dy, float* db, void* stream) {
  fakecuda_launch(S(stream), [=] {
    for (int o = 0; o < O; ++o) {
      double acc = 0;
      for (int n = 0; n < N; ++n) for (int p = 0; p < P; ++p) acc += dy[((size_t)n * O + o) * P + p];
      db[o] = (float)(db[o] + acc);
    }
  });
  return B2C_OK;
}
int b2c_softmax_loss_forward(int N, int C, const float* logits, const float* labels, float* prob, float* loss, void* stream) {
  fakecuda_launch(S(stream), [=] {
    double total = 0;
    for (int n = 0; n < N; ++n) {
      const float* z = logits + (size_t)n * C;
      double mx = z[0], sum = 0;
      for (int c = 1; c < C; ++c) mx = z[c] > mx ? z[c] : mx;
      for (int c = 0; c < C; ++c) sum += std::exp((double)z[c] - mx);
      for (int c = 0; c < C; ++c) prob[(size_t)n * C + c] = (float)(std::exp((double)z[c] - mx) / sum);
      const int lab = (int)labels[n];
      const double p = lab >= 0 && lab < C ? prob[(size_t)n * C + lab] : 1.0;
      total -= std::log(p > 1.17549435e-38 ? p : 1.17549435e-38);
    }
    *loss = (float)(total / N);
  });
  return B2C_OK;
}
int b2c_softmax_loss_backward(int N, int C, const float* prob, const float* labels, float loss_weight, float* dx, void* stream) {
  fakecuda_launch(S(stream), [=] {
    for (int n = 0; n < N; ++n)
      for (int c = 0; c < C; ++c) dx[(size_t)n * C + c] = (prob[(size_t)n * C + c] - ((int)labels[n] == c ? 1.f : 0.f)) * (loss_weight / N);
  });
  return B2C_OK;
}
// h <- m*h + rate*(g*grad_scale + decay*w);  w <- w - h;  g <- 0      (sgd_solver.cu:9-20, L2)
int b2c_sgd_update_arena(int nseg, const size_t* offset, const size_t* count, const float* local_rate, const float* local_decay, float* g,
                         float* w, float* h, float momentum, int l2, float grad_scale, int clear_grads, void* stream) {
  std::vector<size_t> off(offset, offset + nseg), cnt(count, count + nseg);
  std::vector<float> rate(local_rate, local_rate + nseg), decay(local_decay, local_decay + nseg);
  fakecuda_launch(S(stream), [=] {
    for (int s = 0; s < nseg; ++s)
      for (size_t i = off[s]; i < off[s] + cnt[s]; ++i) {
        const float reg = l2 ? w[i] : (w[i] > 0.f ? 1.f : w[i] < 0.f ? -1.f : 0.f);
        const float gi = g[i] * grad_scale + decay[s] * reg;
        h[i] = momentum * h[i] + rate[s] * gi;
        w[i] -= h[i];
        if (clear_grads) g[i] = 0.f;
      }
  });
  return B2C_OK;
}

}  // extern "C"

// y = alpha * op(A) * x + beta * y, A row-major M x N (caffe_gpu_gemv, math_functions.cu:73-82)
extern "C" int b2c_sgemv(int tA, int M, int N, float alpha, const float* A, const float* x, float beta, float* y, void* stream) {
  fakecuda_launch(S(stream), [=] {
    const int rows = tA ? N : M, cols = tA ? M : N;
    for (int i = 0; i < rows; ++i) {
      double acc = 0;
      for (int j = 0; j < cols; ++j) acc += (double)(tA ? A[(size_t)j * N + i] : A[(size_t)i * N + j]) * x[j];
      y[i] = (float)(alpha * acc + (beta == 0.f ? 0.0 : (double)beta * y[i]));
    }
  });
  return B2C_OK;
}
