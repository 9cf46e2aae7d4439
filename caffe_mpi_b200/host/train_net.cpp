// train_net.cpp -- see train_net.hpp.
#include "train_net.hpp"
#include "proto_wire.hpp"

#include <algorithm>
#include <cstring>
#include <cstdlib>

namespace caffe {

static cudaStream_t S() { return Caffe::thread_stream(); }

// ================================================================================================ ReLU
void ReLULayer::Forward_gpu(const vector<Blob*>& b, const vector<Blob*>& t) {
  if (fused_away_) return;
  B2C_CHECK(b2c_relu_forward(b[0]->count(), b[0]->gpu_data(), t[0]->mutable_gpu_data(), slope_, S()));
}
void ReLULayer::Backward_gpu(const vector<Blob*>& t, const vector<bool>& pd, const vector<Blob*>& b) {
  if (!pd[0] || fused_away_) return;
  // in place: bottom data == top data (post-activation), same (x > 0) mask for slope 0 (relu_layer.cpp:27-41)
  B2C_CHECK(b2c_relu_backward(b[0]->count(), t[0]->gpu_diff(), b[0]->gpu_data(), b[0]->mutable_gpu_diff(), slope_, S()));
}

// ================================================================================================ BatchNorm
void BatchNormLayer::LayerSetUp(const vector<Blob*>& b, const vector<Blob*>&) {
  const int C = b[0]->num_axes() > 1 ? b[0]->shape(1) : 1;
  blobs_.resize(scale_bias_ ? 5 : 3);
  blobs_[0].reset(new Blob(vector<int>{C}));      // running mean
  blobs_[1].reset(new Blob(vector<int>{C}));      // running variance (+eps)
  blobs_[2].reset(new Blob(vector<int>{1}));      // variance correction (kept for blob-count compatibility)
  if (scale_bias_) {
    blobs_[3].reset(new Blob(vector<int>{C}));
    blobs_[4].reset(new Blob(vector<int>{C}));
    Fill(scale_filler_, blobs_[3].get());         // defaults: scale = 1, bias = 0 (batch_norm_layer.cpp:52-63)
    Fill(bias_filler_, blobs_[4].get());
  }
  param_propagate_down_.assign(blobs_.size(), false);
  if (scale_bias_) param_propagate_down_[3] = param_propagate_down_[4] = true;
}
void BatchNormLayer::Reshape(const vector<Blob*>& b, const vector<Blob*>& t) {
  if (t[0] != b[0]) t[0]->ReshapeLike(*b[0]);
  const int C = b[0]->shape(1);
  if (!(recompute_ && t[0] != b[0])) xnorm_.ReshapeLike(*b[0]);
  save_mean_.Reshape({C}); save_invstd_.Reshape({C}); scratch_.Reshape({2 * C});
}
void BatchNormLayer::Forward_gpu(const vector<Blob*>& b, const vector<Blob*>& t) {
  const int N = b[0]->shape(0), C = b[0]->shape(1), Sp = (int)(b[0]->count() / ((size_t)N * C));
  if (res_sum_) {
    // residual tail: sum_top = max(0, BatchNorm(x) + other); this layer's own top blob is not materialised
    B2C_CHECK(b2c_bn_forward_train_fused_res(N, C, Sp, b[0]->gpu_data(), scale_bias_ ? blobs_[3]->gpu_data() : nullptr,
                                             scale_bias_ ? blobs_[4]->gpu_data() : nullptr, eps_, maf_, iter_ <= 1 ? 1 : 0,
                                             blobs_[0]->mutable_gpu_data(), blobs_[1]->mutable_gpu_data(), save_mean_.mutable_gpu_data(),
                                             save_invstd_.mutable_gpu_data(), res_other_->gpu_data(), res_sum_->mutable_gpu_data(), 1, S()));
    ++iter_;
    return;
  }
  if (recompute_ && t[0] != b[0]) {
    // fused form: no x_norm blob (backward recomputes it from the bottom), optional ReLU on the way out
    B2C_CHECK(b2c_bn_forward_train_fused(N, C, Sp, b[0]->gpu_data(), scale_bias_ ? blobs_[3]->gpu_data() : nullptr,
                                         scale_bias_ ? blobs_[4]->gpu_data() : nullptr, eps_, maf_, iter_ <= 1 ? 1 : 0,
                                         blobs_[0]->mutable_gpu_data(), blobs_[1]->mutable_gpu_data(), save_mean_.mutable_gpu_data(),
                                         save_invstd_.mutable_gpu_data(), t[0]->mutable_gpu_data(), fuse_relu_ ? 1 : 0, S()));
    ++iter_;
    return;
  }
  B2_CHECK(!fuse_relu_, "BatchNorm: ReLU fusion needs the recompute form (top != bottom)");
  B2C_CHECK(b2c_bn_forward_train(N, C, Sp, b[0]->gpu_data(), scale_bias_ ? blobs_[3]->gpu_data() : nullptr,
                                 scale_bias_ ? blobs_[4]->gpu_data() : nullptr, eps_, maf_, iter_ <= 1 ? 1 : 0,
                                 blobs_[0]->mutable_gpu_data(), blobs_[1]->mutable_gpu_data(), save_mean_.mutable_gpu_data(),
                                 save_invstd_.mutable_gpu_data(), xnorm_.mutable_gpu_data(), t[0]->mutable_gpu_data(), S()));
  ++iter_;
}
void BatchNormLayer::Backward_gpu(const vector<Blob*>& t, const vector<bool>&, const vector<Blob*>& b) {
  const int N = b[0]->shape(0), C = b[0]->shape(1), Sp = (int)(b[0]->count() / ((size_t)N * C));
  float* dg = scale_bias_ ? blobs_[3]->mutable_gpu_diff() : scratch_.mutable_gpu_data();
  float* db = scale_bias_ ? blobs_[4]->mutable_gpu_diff() : scratch_.mutable_gpu_data() + C;
  if (res_sum_) {
    B2C_CHECK(b2c_bn_backward_fused_res(N, C, Sp, res_sum_->gpu_diff(), res_part2_ ? res_part2_->gpu_diff() : nullptr, res_sum_->gpu_data(),
                                        b[0]->gpu_data(), save_mean_.gpu_data(),
                                        save_invstd_.gpu_data(), scale_bias_ ? blobs_[3]->gpu_data() : nullptr,
                                        scale_bias_ ? blobs_[4]->gpu_data() : nullptr, dg, db, b[0]->mutable_gpu_diff(),
                                        res_prop_ ? res_other_->mutable_gpu_diff() : nullptr, S()));
    return;
  }
  if (recompute_ && t[0] != b[0]) {
    B2C_CHECK(b2c_bn_backward_fused(N, C, Sp, t[0]->gpu_diff(), b[0]->gpu_data(), save_mean_.gpu_data(), save_invstd_.gpu_data(),
                                    scale_bias_ ? blobs_[3]->gpu_data() : nullptr, scale_bias_ ? blobs_[4]->gpu_data() : nullptr, dg, db,
                                    b[0]->mutable_gpu_diff(), fuse_relu_ ? 1 : 0, S()));
    return;
  }
  B2C_CHECK(b2c_bn_backward(N, C, Sp, t[0]->gpu_diff(), xnorm_.gpu_data(), scale_bias_ ? blobs_[3]->gpu_data() : nullptr,
                            save_invstd_.gpu_data(), dg, db, b[0]->mutable_gpu_diff(), S()));
}

// ================================================================================================ Pooling
PoolingLayer::~PoolingLayer() { if (mask_) cudaFree(mask_); }
static int pooled(int in, int k, int s, int p) {
  int o = (int)std::ceil((float)(in + 2 * p - k) / s) + 1;
  if (p > 0 && (o - 1) * s >= in + p) --o;
  return o;
}
void PoolingLayer::Reshape(const vector<Blob*>& b, const vector<Blob*>& t) {
  eff_ = q_;
  if (q_.global_pooling) { eff_.kernel_h = b[0]->shape(2); eff_.kernel_w = b[0]->shape(3); eff_.stride_h = eff_.stride_w = 1; eff_.pad_h = eff_.pad_w = 0; }
  t[0]->Reshape({b[0]->shape(0), b[0]->shape(1), pooled(b[0]->shape(2), eff_.kernel_h, eff_.stride_h, eff_.pad_h),
                 pooled(b[0]->shape(3), eff_.kernel_w, eff_.stride_w, eff_.pad_w)});
  if (eff_.pool == 0 && t[0]->count() > mask_cap_) {
    if (mask_) CUDA_CHECK(cudaFree(mask_));
    CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&mask_), sizeof(int) * t[0]->count()));
    mask_cap_ = t[0]->count();
  }
}
void PoolingLayer::Forward_gpu(const vector<Blob*>& b, const vector<Blob*>& t) {
  B2C_CHECK(b2c_pool_forward(eff_.pool, b[0]->shape(0) * b[0]->shape(1), b[0]->shape(2), b[0]->shape(3), eff_.kernel_h, eff_.kernel_w,
                             eff_.stride_h, eff_.stride_w, eff_.pad_h, eff_.pad_w, b[0]->gpu_data(), t[0]->mutable_gpu_data(), mask_, S()));
}
void PoolingLayer::Backward_gpu(const vector<Blob*>& t, const vector<bool>& pd, const vector<Blob*>& b) {
  if (!pd[0]) return;
  B2C_CHECK(b2c_pool_backward(eff_.pool, b[0]->shape(0) * b[0]->shape(1), b[0]->shape(2), b[0]->shape(3), eff_.kernel_h, eff_.kernel_w,
                              eff_.stride_h, eff_.stride_w, eff_.pad_h, eff_.pad_w, t[0]->gpu_diff(), mask_, b[0]->mutable_gpu_diff(), S()));
}

// ================================================================================================ Eltwise SUM
void EltwiseLayer::Forward_gpu(const vector<Blob*>& b, const vector<Blob*>& t) {
  B2_CHECK(b.size() >= 2, "Eltwise needs two bottoms");
  if (fused_away_) return;
  if (fuse_relu_ && b.size() == 2) {        // y = max(0, a + b): the in-place ReLU that follows is folded in
    B2C_CHECK(b2c_add_relu(t[0]->count(), b[0]->gpu_data(), b[1]->gpu_data(), t[0]->mutable_gpu_data(), S()));
    return;
  }
  B2C_CHECK(b2c_add(t[0]->count(), b[0]->gpu_data(), b[1]->gpu_data(), t[0]->mutable_gpu_data(), S()));
  for (size_t i = 2; i < b.size(); ++i) B2C_CHECK(b2c_add(t[0]->count(), t[0]->gpu_data(), b[i]->gpu_data(), t[0]->mutable_gpu_data(), S()));
}
void EltwiseLayer::Backward_gpu(const vector<Blob*>& t, const vector<bool>& pd, const vector<Blob*>& b) {
  if (fused_away_) return;
  if (fuse_relu_ && b.size() == 2) {        // dx_a = dx_b = dy * (y > 0): ReLU backward + the two copies of SUM's backward
    if (pd[0] || pd[1])
      B2C_CHECK(b2c_relu_backward2(t[0]->count(), t[0]->gpu_diff(), t[0]->gpu_data(), pd[0] ? b[0]->mutable_gpu_diff() : nullptr,
                                   pd[1] ? b[1]->mutable_gpu_diff() : nullptr, S()));
    return;
  }
  for (size_t i = 0; i < b.size(); ++i)
    if (pd[i]) CUDA_CHECK(cudaMemcpyAsync(b[i]->mutable_gpu_diff(), t[0]->gpu_diff(), sizeof(float) * t[0]->count(), cudaMemcpyDeviceToDevice, S()));
}

// ================================================================================================ InnerProduct
void InnerProductLayer::LayerSetUp(const vector<Blob*>& b, const vector<Blob*>&) {
  K_ = (int)b[0]->count(1);
  blobs_.resize(bias_ ? 2 : 1);
  blobs_[0].reset(new Blob(vector<int>{num_output_, K_}));
  Fill(wf_, blobs_[0].get());
  if (bias_) { blobs_[1].reset(new Blob(vector<int>{num_output_})); Fill(bf_, blobs_[1].get()); }
  param_propagate_down_.assign(blobs_.size(), true);
}
void InnerProductLayer::Reshape(const vector<Blob*>& b, const vector<Blob*>& t) {
  M_ = b[0]->shape(0);
  B2_CHECK((int)b[0]->count(1) == K_, "Input size incompatible with inner product parameters.");
  t[0]->Reshape({M_, num_output_});
  const size_t wsb = b2c_sgemm_workspace_bytes(0, 1, M_, num_output_, K_);
  if (wsb) gemm_ws_.Reshape({(int)((wsb + 3) / 4)});
  // backward: dW = dy^T[N x M] * (x^T[K x M])^T (reduction over the batch), dx = dy[M x N] * (W^T[K x N])^T (reduction over the outputs)
  const char* e = getenv("B2C_IP_BWD_TC");
  const bool on = !e || atoi(e) != 0;
  // dW: the reduction runs over the batch -- only a few K blocks per 128 x 128 output tile unless the batch is large; measured on
  // AlexNet (batch 256, 4096 x 9216 outputs) the tile set-up and the read-modify-write epilogue cost more than the FFMA kernel
  // (InnerProduct backward 2.95 -> 3.11 ms, profiles/r02_c12_bench_alexnet.json), so the batch has to be at least 512
  bwd_w_tc_ = on && M_ >= 512 && b2c_sgemm_tc_supported(0, 1, num_output_, K_, M_) != 0;
  bwd_x_tc_ = on && b2c_sgemm_tc_supported(0, 1, M_, K_, num_output_) != 0;
  size_t bws = 0;
  if (bwd_w_tc_) { dyt_.Reshape({num_output_, M_}); xt_.Reshape({K_, M_}); bws = std::max(bws, b2c_sgemm_workspace_bytes(0, 1, num_output_, K_, M_)); }
  if (bwd_x_tc_) { wt_.Reshape({K_, num_output_}); bws = std::max(bws, b2c_sgemm_workspace_bytes(0, 1, M_, K_, num_output_)); }
  if (bws) bwd_ws_.Reshape({(int)((bws + 3) / 4)});
}
void InnerProductLayer::Forward_gpu(const vector<Blob*>& b, const vector<Blob*>& t) {
  // y[M x N] = x[M x K] * W[N x K]^T (+ bias)   (inner_product_layer.cpp: caffe_gpu_gemm(NoTrans, Trans, M, N, K))
  const float* w = blobs_[0]->gpu_data();
  if ((reinterpret_cast<uintptr_t>(w) & 15u) && K_ >= 64 && K_ % 4 == 0) {     // arena slot on an 8-byte boundary: aligned copy for the TMA-fed GEMM
    if (w_aligned_.count() != blobs_[0]->count()) w_aligned_.Reshape(blobs_[0]->shape());
    CUDA_CHECK(cudaMemcpyAsync(w_aligned_.mutable_gpu_data(), w, sizeof(float) * blobs_[0]->count(), cudaMemcpyDeviceToDevice, S()));
    w = w_aligned_.gpu_data();
  }
  B2C_CHECK(b2c_sgemm_ex(0, 1, M_, num_output_, K_, 1.f, b[0]->gpu_data(), w, 0.f, t[0]->mutable_gpu_data(),
                         gemm_ws_.count() ? gemm_ws_.mutable_gpu_data() : nullptr, sizeof(float) * gemm_ws_.count(), S()));
  if (bias_) B2C_CHECK(b2c_bias_forward(M_, num_output_, 1, blobs_[1]->gpu_data(), t[0]->mutable_gpu_data(), S()));
}
void InnerProductLayer::Backward_gpu(const vector<Blob*>& t, const vector<bool>& pd, const vector<Blob*>& b) {
  // dW[N x K] += dy^T[N x M] * x[M x K];  db += sum_m dy;  dx[M x K] = dy[M x N] * W[N x K]
  float* ws = bwd_ws_.count() ? bwd_ws_.mutable_gpu_data() : nullptr;
  const size_t wsb = sizeof(float) * bwd_ws_.count();
  float* dw = blobs_[0]->mutable_gpu_diff();
  if (bwd_w_tc_) {                                  // (the GEMM's output has no alignment requirement: 4-byte stores)
    B2C_CHECK(b2c_transpose(M_, num_output_, t[0]->gpu_diff(), dyt_.mutable_gpu_data(), S()));
    B2C_CHECK(b2c_transpose(M_, K_, b[0]->gpu_data(), xt_.mutable_gpu_data(), S()));
    B2C_CHECK(b2c_sgemm_ex(0, 1, num_output_, K_, M_, 1.f, dyt_.gpu_data(), xt_.gpu_data(), 1.f, dw, ws, wsb, S()));
  } else {
    B2C_CHECK(b2c_sgemm(1, 0, num_output_, K_, M_, 1.f, t[0]->gpu_diff(), b[0]->gpu_data(), 1.f, dw, S()));
  }
  if (bias_) B2C_CHECK(b2c_bias_backward(M_, num_output_, 1, t[0]->gpu_diff(), blobs_[1]->mutable_gpu_diff(), S()));
  if (pd[0]) {
    if (bwd_x_tc_) {
      B2C_CHECK(b2c_transpose(num_output_, K_, blobs_[0]->gpu_data(), wt_.mutable_gpu_data(), S()));
      B2C_CHECK(b2c_sgemm_ex(0, 1, M_, K_, num_output_, 1.f, t[0]->gpu_diff(), wt_.gpu_data(), 0.f, b[0]->mutable_gpu_diff(), ws, wsb, S()));
    } else {
      B2C_CHECK(b2c_sgemm(0, 0, M_, K_, num_output_, 1.f, t[0]->gpu_diff(), blobs_[0]->gpu_data(), 0.f, b[0]->mutable_gpu_diff(), S()));
    }
  }
}

// ================================================================================================ LRN / Dropout / Concat
void LRNLayer::Forward_gpu(const vector<Blob*>& b, const vector<Blob*>& t) {
  const int N = b[0]->shape(0), C = b[0]->shape(1), Sp = (int)(b[0]->count() / ((size_t)N * C));
  B2C_CHECK(b2c_lrn_forward(N, C, Sp, size_, alpha_, beta_, k_, b[0]->gpu_data(), scale_.mutable_gpu_data(), t[0]->mutable_gpu_data(), S()));
}
void LRNLayer::Backward_gpu(const vector<Blob*>& t, const vector<bool>& pd, const vector<Blob*>& b) {
  if (!pd[0]) return;
  const int N = b[0]->shape(0), C = b[0]->shape(1), Sp = (int)(b[0]->count() / ((size_t)N * C));
  B2C_CHECK(b2c_lrn_backward(N, C, Sp, size_, alpha_, beta_, b[0]->gpu_data(), t[0]->gpu_data(), scale_.gpu_data(), t[0]->gpu_diff(),
                             b[0]->mutable_gpu_diff(), S()));
}
void DropoutLayer::Forward_gpu(const vector<Blob*>& b, const vector<Blob*>& t) {
  B2C_CHECK(b2c_dropout_mask(b[0]->count(), ratio_, seed_, offset_, mask_.mutable_gpu_data(), S()));
  offset_ += b[0]->count();
  B2C_CHECK(b2c_mul(b[0]->count(), b[0]->gpu_data(), mask_.gpu_data(), t[0]->mutable_gpu_data(), S()));
}
void DropoutLayer::Backward_gpu(const vector<Blob*>& t, const vector<bool>& pd, const vector<Blob*>& b) {
  if (!pd[0]) return;
  B2C_CHECK(b2c_mul(b[0]->count(), t[0]->gpu_diff(), mask_.gpu_data(), b[0]->mutable_gpu_diff(), S()));
}
void ConcatLayer::Reshape(const vector<Blob*>& b, const vector<Blob*>& t) {
  vector<int> s = b[0]->shape();
  B2_CHECK(s.size() >= 2, "Concat: need at least 2 axes");
  for (size_t i = 1; i < b.size(); ++i) {
    B2_CHECK(b[i]->shape().size() == s.size() && b[i]->shape(0) == s[0] && b[i]->count(2) == b[0]->count(2),
             "All inputs must have the same shape, except at concat_axis.");
    s[1] += b[i]->shape(1);
  }
  t[0]->Reshape(s);
}
void ConcatLayer::Forward_gpu(const vector<Blob*>& b, const vector<Blob*>& t) {
  const size_t inner = t[0]->count(2), N = t[0]->shape(0), Ct = t[0]->shape(1);
  size_t coff = 0;
  for (Blob* src : b) {
    const size_t w = sizeof(float) * src->shape(1) * inner;
    CUDA_CHECK(cudaMemcpy2DAsync(t[0]->mutable_gpu_data() + coff * inner, sizeof(float) * Ct * inner, src->gpu_data(), w, w, N,
                                 cudaMemcpyDeviceToDevice, S()));
    coff += src->shape(1);
  }
}
void ConcatLayer::Backward_gpu(const vector<Blob*>& t, const vector<bool>& pd, const vector<Blob*>& b) {
  const size_t inner = t[0]->count(2), N = t[0]->shape(0), Ct = t[0]->shape(1);
  size_t coff = 0;
  for (size_t i = 0; i < b.size(); ++i) {
    const size_t w = sizeof(float) * b[i]->shape(1) * inner;
    if (pd[i])
      CUDA_CHECK(cudaMemcpy2DAsync(b[i]->mutable_gpu_diff(), w, t[0]->gpu_diff() + coff * inner, sizeof(float) * Ct * inner, w, N,
                                   cudaMemcpyDeviceToDevice, S()));
    coff += b[i]->shape(1);
  }
}

// ================================================================================================ SoftmaxWithLoss
void SoftmaxWithLossLayer::Reshape(const vector<Blob*>& b, const vector<Blob*>& t) {
  prob_.ReshapeLike(*b[0]);
  t[0]->Reshape({1});
}
void SoftmaxWithLossLayer::Forward_gpu(const vector<Blob*>& b, const vector<Blob*>& t) {
  B2C_CHECK(b2c_softmax_loss_forward(b[0]->shape(0), (int)b[0]->count(1), b[0]->gpu_data(), b[1]->gpu_data(), prob_.mutable_gpu_data(),
                                     t[0]->mutable_gpu_data(), S()));
}
void SoftmaxWithLossLayer::Backward_gpu(const vector<Blob*>&, const vector<bool>& pd, const vector<Blob*>& b) {
  if (!pd[0]) return;
  B2C_CHECK(b2c_softmax_loss_backward(b[0]->shape(0), (int)b[0]->count(1), prob_.gpu_data(), b[1]->gpu_data(), loss_weight_,
                                      b[0]->mutable_gpu_diff(), S()));
}

// ================================================================================================ Accuracy
void AccuracyLayer::Forward_gpu(const vector<Blob*>& b, const vector<Blob*>& t) {
  B2C_CHECK(b2c_accuracy(b[0]->shape(0), (int)b[0]->count(1), top_k_, b[0]->gpu_data(), b[1]->gpu_data(), t[0]->mutable_gpu_data(),
                         hits_.mutable_gpu_data(), S()));
}

// ================================================================================================ synthetic data
SyntheticDataLayer::~SyntheticDataLayer() {
  if (host_) cudaFreeHost(host_);
  if (copy_stream_) cudaStreamSynchronize(copy_stream_);
  if (host_u8_) cudaFreeHost(host_u8_);
  for (Slot& sl : slot_) {
    if (sl.host_off) cudaFreeHost(sl.host_off);
    if (sl.dev_u8) cudaFree(sl.dev_u8);
    if (sl.dev_off) cudaFree(sl.dev_off);
    if (sl.copied) cudaEventDestroy(sl.copied);
    if (sl.consumed) cudaEventDestroy(sl.consumed);
  }
  if (copy_stream_) cudaStreamDestroy(copy_stream_);
  if (dev_mean_) cudaFree(dev_mean_);
}
static inline uint64_t splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
void SyntheticDataLayer::LayerSetUp(const vector<Blob*>&, const vector<Blob*>& t) {
  std::mt19937 rng((uint32_t)seed_);
  for (size_t i = 0; i < t.size(); ++i) t[i]->Reshape(shapes_[i]);
  n0_ = t[0]->count();
  if (t.size() > 1) {
    std::uniform_int_distribution<int> u(0, classes_ - 1);
    float* l = t[1]->mutable_cpu_data();
    for (size_t i = 0; i < t[1]->count(); ++i) l[i] = (float)u(rng);
  }
  datum_mode_ = tf_.on && tf_.crop > 0 && shapes_[0].size() == 4 && shapes_[0][2] == tf_.crop && shapes_[0][3] == tf_.crop;
  if (datum_mode_) {
    const int N = shapes_[0][0], C = shapes_[0][1];
    hd_ = wd_ = tf_.crop + 32;
    u8_bytes_ = (size_t)N * C * hd_ * wd_;
    CUDA_CHECK(cudaMallocHost(reinterpret_cast<void**>(&host_u8_), u8_bytes_));
    CUDA_CHECK(cudaStreamCreateWithFlags(&copy_stream_, cudaStreamNonBlocking));
    for (Slot& sl : slot_) {
      CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&sl.dev_u8), u8_bytes_));
      CUDA_CHECK(cudaMallocHost(reinterpret_cast<void**>(&sl.host_off), sizeof(int) * 3 * N));
      CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&sl.dev_off), sizeof(int) * 3 * N));
      CUDA_CHECK(cudaEventCreateWithFlags(&sl.copied, cudaEventDisableTiming));
      CUDA_CHECK(cudaEventCreateWithFlags(&sl.consumed, cudaEventDisableTiming));
    }
    for (size_t i = 0; i < u8_bytes_; i += 8) {          // uniform bytes, 8 per draw
      uint64_t r = splitmix64(seed_ * 0x100000001B3ull + i);
      for (size_t k = 0; k < 8 && i + k < u8_bytes_; ++k, r >>= 8) host_u8_[i + k] = (unsigned char)(r & 0xff);
    }
    if (!tf_.mean_value.empty()) {
      B2_CHECK(tf_.mean_value.size() == 1 || (int)tf_.mean_value.size() == C, "Specify either 1 mean_value or as many as channels");
      vector<float> mv(C, tf_.mean_value[0]);
      if ((int)tf_.mean_value.size() == C) mv = tf_.mean_value;
      CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&dev_mean_), sizeof(float) * C));
      CUDA_CHECK(cudaMemcpy(dev_mean_, mv.data(), sizeof(float) * C, cudaMemcpyHostToDevice));
    }
    for (Blob* b : t) b->gpu_data();
    LoadBatch(t[0], Caffe::thread_stream());              // the resident batch of the non-e2e steps
    CUDA_CHECK(cudaStreamSynchronize(Caffe::thread_stream()));
    return;
  }
  CUDA_CHECK(cudaMallocHost(reinterpret_cast<void**>(&host_), sizeof(float) * n0_));
  std::normal_distribution<float> g(0.f, 1.f);
  for (size_t i = 0; i < n0_; ++i) host_[i] = g(rng);
  memcpy(t[0]->mutable_cpu_data(), host_, sizeof(float) * n0_);
  for (Blob* b : t) b->gpu_data();   // upload once; the blobs stay resident
}
void SyntheticDataLayer::LoadBatch(Blob* top, cudaStream_t st) {
  if (!datum_mode_) {
    CUDA_CHECK(cudaMemcpyAsync(top->mutable_gpu_data(), host_, sizeof(float) * n0_, cudaMemcpyHostToDevice, st));
    return;
  }
  const int N = shapes_[0][0], C = shapes_[0][1];
  Slot& sl = slot_[cur_];
  if (!sl.in_flight) IssueCopy(cur_);                    // first batch, or nothing was prefetched
  CUDA_CHECK(cudaStreamWaitEvent(st, sl.copied, 0));
  B2C_CHECK(b2c_transform_u8(sl.dev_u8, N, C, hd_, wd_, tf_.crop, tf_.crop, sl.dev_off, sl.dev_off + N,
                             reinterpret_cast<const unsigned char*>(sl.dev_off + 2 * N), dev_mean_, nullptr, tf_.scale, top->mutable_gpu_data(), st));
  CUDA_CHECK(cudaEventRecord(sl.consumed, st));
  sl.in_flight = false; sl.used = true;
  cur_ ^= 1;
  IssueCopy(cur_);                                       // the next batch travels while this one is being computed on
}
void SyntheticDataLayer::IssueCopy(int s) {
  const int N = shapes_[0][0];
  Slot& sl = slot_[s];
  if (sl.used) {
    CUDA_CHECK(cudaEventSynchronize(sl.copied));         // the previous copy out of this slot's pinned offsets finished long ago
    CUDA_CHECK(cudaStreamWaitEvent(copy_stream_, sl.consumed, 0));   // ... and the transform that read the slot has to be done before it is overwritten
  }
  // DataTransformer's draws (data_transformer.cpp:129-135,187,224-225): mirror = rand0 % 2, offsets = rand % (extent - crop + 1)
  unsigned char* mir = reinterpret_cast<unsigned char*>(sl.host_off + 2 * N);
  for (int i = 0; i < N; ++i) {
    const uint64_t base = (draws_ * (uint64_t)N + (uint64_t)i) * 3;
    const unsigned r0 = (unsigned)(splitmix64(seed_ + base) >> 33) + 1, r1 = (unsigned)(splitmix64(seed_ + base + 1) >> 33) + 1,
                   r2 = (unsigned)(splitmix64(seed_ + base + 2) >> 33) + 1;
    mir[i] = (tf_.mirror && (r0 % 2)) ? 1 : 0;
    sl.host_off[i] = (int)(r1 % (unsigned)(hd_ - tf_.crop + 1));
    sl.host_off[N + i] = (int)(r2 % (unsigned)(wd_ - tf_.crop + 1));
  }
  ++draws_;
  CUDA_CHECK(cudaMemcpyAsync(sl.dev_u8, host_u8_, u8_bytes_, cudaMemcpyHostToDevice, copy_stream_));
  CUDA_CHECK(cudaMemcpyAsync(sl.dev_off, sl.host_off, sizeof(int) * 3 * N, cudaMemcpyHostToDevice, copy_stream_));
  CUDA_CHECK(cudaEventRecord(sl.copied, copy_stream_));
  sl.in_flight = true;
}

// ================================================================================================ TrainNet
TrainNet::TrainNet(const Net& net, const SolverParameter& sp, int num_classes, uint64_t seed, int math) {
  name_ = net.name();
  Caffe::set_random_seed(seed);
  std::map<string, bool> need;
  vector<ParamSpec> specs;
  for (size_t li = 0; li < net.layers().size(); ++li) {
    const NetLayer& L = net.layers()[li];
    const string& type = L.param.type;
    Node node;
    for (auto& bn : L.param.bottom) {
      auto it = blobs_.find(bn);
      B2_CHECK(it != blobs_.end(), "Unknown bottom blob '" + bn + "'");
      node.bottom.push_back(it->second.get());
    }
    for (auto& tn : L.param.top) {
      auto& sp_ = blobs_[tn];
      if (!sp_) sp_.reset(new Blob());
      node.top.push_back(sp_.get());
    }
    shared_ptr<LayerBase> layer;
    if (type == "Data" && L.use_database) {
      // data_param.source opened: the layer reads it (parser threads -> pinned batch -> device transform), data_layer.hpp
      B2_CHECK(db_data_ == nullptr && data_ == nullptr, "TrainNet: one data source per net");
      auto* d = new DataLayer(L, seed);
      layer.reset(d);
      db_data_ = d;
      db_data_node_ = (int)layers_.size();
    } else if (type == "Data" || type == "Input" || type == "DummyData" || type == "ImageData") {
      vector<vector<int>> shapes;
      for (size_t t = 0; t < L.param.top.size(); ++t) shapes.push_back(net.top_shape((int)li, (int)t));
      SyntheticDataLayer::Transform tf;
      tf.on = L.has_transform; tf.mirror = L.mirror; tf.crop = L.crop_size; tf.scale = L.transform_scale; tf.mean_value = L.mean_value;
      auto* d = new SyntheticDataLayer(L.param, shapes, num_classes, seed, tf);
      layer.reset(d);
      data_ = d;
    } else if (type == "Convolution") {
      LayerParameter lp = L.param;
      lp.convolution_param.math = math;
      layer = LayerRegistry::CreateLayer(lp);
    } else if (type == "ReLU") layer.reset(new ReLULayer(L.param, L.relu_slope));
    else if (type == "BatchNorm") {
      auto* bn = new BatchNormLayer(L.param, L.bn_scale_bias, L.bn_eps, L.bn_maf);
      if (L.bn_has_scale_filler) bn->set_scale_filler(L.bn_scale_filler);      // batch_norm_layer.cpp:52-63
      if (L.bn_has_bias_filler) bn->set_bias_filler(L.bn_bias_filler);
      layer.reset(bn);
    }
    else if (type == "Pooling") layer.reset(new PoolingLayer(L.param, L.pooling));
    else if (type == "Eltwise") {
      B2_CHECK(L.eltwise_op == 1, "TrainNet: Eltwise PROD / MAX are not built (the BASELINE nets use SUM)");
      for (float c : L.eltwise_coeff) B2_CHECK(c == 1.f, "TrainNet: Eltwise SUM is built for unit coefficients only");
      layer.reset(new EltwiseLayer(L.param));
    }
    else if (type == "InnerProduct") layer.reset(new InnerProductLayer(L.param, L.ip_num_output, L.ip_bias, L.ip_weight_filler, L.ip_bias_filler));
    else if (type == "SoftmaxWithLoss") {
      auto* sl = new SoftmaxWithLossLayer(L.param);
      if (!L.loss_weight.empty()) sl->set_loss_weight(L.loss_weight[0]);      // GoogLeNet's auxiliary classifiers: 0.3
      layer.reset(sl);
    }
    else if (type == "Accuracy") layer.reset(new AccuracyLayer(L.param, L.accuracy_top_k));     // forward only, no gradient
    else if (type == "LRN") {
      B2_CHECK(L.lrn_region == 0, "TrainNet: LRN WITHIN_CHANNEL is not built (the BASELINE nets use ACROSS_CHANNELS)");
      layer.reset(new LRNLayer(L.param, L.lrn_size, L.lrn_alpha, L.lrn_beta, L.lrn_k));
    }
    else if (type == "Dropout") layer.reset(new DropoutLayer(L.param, L.dropout_ratio, seed * 0x100000001B3ull + li));
    else if (type == "Concat") {
      B2_CHECK(L.concat_axis == 1, "TrainNet: Concat is built for the channel axis only");
      layer.reset(new ConcatLayer(L.param));
    }
    else B2_CHECK(false, "TrainNet: layer type '" + type + "' is not built yet (SURVEY 8f rank 2 covers the ResNet-50 set)");
    layer->SetUp(node.bottom, node.top);
    // learnable blobs in layer order: Net::AppendParam (net.cpp:573-606) appends EVERY blob of every layer, so param ids,
    // reduce-bucket boundaries and the SolverState history list (5 blobs per BatchNorm layer) match the reference.
    // Blobs the layer never differentiates (BatchNorm mean / variance / correction, batch_norm_layer.cpp: lr_mult forced
    // to 0) are marked `statistic`; trainable_ids_ lists the rest for the trainer API.
    node.first_param = (int)learnable_.size();
    for (size_t bi = 0; bi < layer->blobs().size(); ++bi) {
      ParamSpec ps = bi < L.param.param.size() ? L.param.param[bi] : ParamSpec();
      if (!layer->param_propagate_down((int)bi)) { ps.lr_mult = 0.f; ps.decay_mult = 0.f; ps.statistic = true; }
      else trainable_ids_.push_back((int)learnable_.size());
      learnable_.push_back(layer->blobs()[bi]);
      specs.push_back(ps);
      ++node.num_params;
    }
    // need-backward analysis (net.cpp:160-283)
    bool nb = false;
    for (size_t i = 0; i < node.bottom.size(); ++i) {
      const bool bneed = need.count(L.param.bottom[i]) ? need[L.param.bottom[i]] : false;
      node.propagate_down.push_back(bneed && !(type == "SoftmaxWithLoss" && i == 1));
      nb |= bneed;
    }
    for (int k = 0; k < node.num_params; ++k) nb |= specs[node.first_param + k].lr_mult != 0.f;
    if (layer.get() == data_ || type == "Accuracy") nb = false;      // Accuracy: forward-only metric (accuracy_layer.hpp: "cannot backpropagate")
    node.need_backward = nb;
    for (auto& tn : L.param.top) need[tn] = nb;
    if (type == "SoftmaxWithLoss") loss_blob_ = node.top[0];
    node.bottom_diff_tmp.assign(node.bottom.size(), nullptr);
    layers_.push_back(layer);
    layer_names_.push_back(L.param.name);
    layer_types_.push_back(type);
    nodes_.push_back(node);
  }
  // Fusion pass (B2C_FUSE=0 disables): BatchNorm -> in-place ReLU and Eltwise(SUM, 2 bottoms) -> in-place ReLU run as one
  // kernel each way, and BatchNorm stops materialising x_norm (csrc/layers_fused.cu).  The ReLU layer stays in the graph as a
  // no-op so that layer indices, names and blobs are those of the prototxt.  Results are bit-identical to the unfused graph.
  {
    const char* e = getenv("B2C_FUSE");
    const bool fuse = !e || atoi(e) != 0;
    // conv layers adding their bottom gradient straight into a fan-out blob's diff (TMA reduce-add store): measured SLOWER than
    // the shadow blob + add pass on ResNet-50 (dgrad 3.04 -> 3.82 ms/step against 0.50 ms of adds saved; the L2 reduction
    // units retire the 16 large accumulating stores at about half the plain-store rate), so it is opt-in: B2C_FUSE_FANOUT=1
    const char* ef = getenv("B2C_FUSE_FANOUT");
    fuse_fanout_ = fuse && ef && atoi(ef) != 0;
    for (size_t i = 0; fuse && i < layers_.size(); ++i) {
      auto* bn = dynamic_cast<BatchNormLayer*>(layers_[i].get());
      auto* el = dynamic_cast<EltwiseLayer*>(layers_[i].get());
      if (!bn && !el) continue;
      const bool out_of_place = nodes_[i].top[0] != nodes_[i].bottom[0];
      ReLULayer* relu = nullptr;
      if (i + 1 < layers_.size()) {
        relu = dynamic_cast<ReLULayer*>(layers_[i + 1].get());
        if (relu && !(nodes_[i + 1].bottom[0] == nodes_[i].top[0] && nodes_[i + 1].top[0] == nodes_[i].top[0] && relu->negative_slope() == 0.f))
          relu = nullptr;
      }
      if (bn && out_of_place) {
        bn->set_fusion(true, relu != nullptr);
        bn->Reshape(nodes_[i].bottom, nodes_[i].top);
        if (relu) relu->set_fused_away(true);
      } else if (el && relu && nodes_[i].bottom.size() == 2) {
        el->set_fuse_relu(true);
        relu->set_fused_away(true);
      }
    }
  }
  // diff accumulation where a blob fans out (insert_splits.cpp / SplitLayer::Backward): the consumer that runs FIRST in
  // the backward pass writes the blob's diff; every other consumer writes a shadow diff that is then added in
  std::map<Blob*, int> writers;
  for (int li = (int)nodes_.size() - 1; li >= 0; --li) {
    Node& nd = nodes_[li];
    if (!nd.need_backward) continue;
    for (size_t i = 0; i < nd.bottom.size(); ++i) {
      if (!nd.propagate_down[i]) continue;
      const bool in_place = i < nd.top.size() && nd.top[i] == nd.bottom[i];
      if (in_place) continue;
      if (writers[nd.bottom[i]]++ > 0) {
        if (fuse_fanout_ && layers_[li]->SupportsBottomDiffAccumulate((int)i)) {
          nd.accumulate_bottom.resize(nd.bottom.size(), false);       // the layer adds into the blob's diff itself
          nd.accumulate_bottom[i] = true;
        } else {
          tmp_diffs_.emplace_back(new Blob(nd.bottom[i]->shape()));
          nd.bottom_diff_tmp[i] = tmp_diffs_.back().get();
        }
      }
    }
  }
  // Residual tails (B2C_FUSE_RES=0 disables): BatchNorm -> Eltwise(SUM, 2 bottoms) -> in-place ReLU.  The BatchNorm that runs LAST
  // among the producers of the sum's bottoms takes over the sum and the ReLU in both directions (csrc/layers_fused.cu, *_res): its
  // top is not materialised, four passes over a block-output-sized tensor go away per block and step.  Conditions: the
  // BatchNorm's top has no other consumer, nothing between it and the sum rewrites the other operand, the sum is the first
  // backward writer of both bottoms (no shadow diffs).
  {
    const char* e = getenv("B2C_FUSE");
    const char* er = getenv("B2C_FUSE_RES");
    const bool fuse_res = (!e || atoi(e) != 0) && (!er || atoi(er) != 0);
    for (size_t i = 0; fuse_res && i < layers_.size(); ++i) {
      auto* el = dynamic_cast<EltwiseLayer*>(layers_[i].get());
      Node& en = nodes_[i];
      if (!el || !el->fuse_relu() || en.bottom.size() != 2 || !en.need_backward) continue;
      if (!en.accumulate_bottom.empty() || en.bottom_diff_tmp[0] || en.bottom_diff_tmp[1]) continue;
      int prod[2] = {-1, -1};
      for (int k = 0; k < 2; ++k)
        for (int j = (int)i - 1; j >= 0 && prod[k] < 0; --j)
          for (Blob* tb : nodes_[j].top) if (tb == en.bottom[k]) prod[k] = j;
      if (prod[0] < 0 || prod[1] < 0 || prod[0] == prod[1]) continue;
      const int k = prod[0] > prod[1] ? 0 : 1, j = prod[k];
      auto* bn = dynamic_cast<BatchNormLayer*>(layers_[j].get());
      if (!bn || !bn->recompute() || bn->fuse_relu() || nodes_[j].top[0] == nodes_[j].bottom[0] || !en.propagate_down[k]) continue;
      int consumers = 0;
      for (size_t l = 0; l < nodes_.size(); ++l)
        for (Blob* bb : nodes_[l].bottom) if (bb == en.bottom[k]) ++consumers;
      bool other_rewritten = false;
      for (size_t l = (size_t)j + 1; l < i; ++l) {       // ... nor reads it (its backward would add into a diff this layer has yet to write)
        for (Blob* tb : nodes_[l].top) if (tb == en.bottom[1 - k]) other_rewritten = true;
        for (Blob* bb : nodes_[l].bottom) if (bb == en.bottom[1 - k]) other_rewritten = true;
      }
      if (consumers != 1 || other_rewritten) continue;
      bn->set_residual(en.bottom[1 - k], en.top[0], en.propagate_down[1 - k]);
      el->set_fused_away(true);
    }
    // a fused tail's sum blob with a second backward writer: hand the shadow diff to the tail instead of adding it in a pass of its own
    const char* ed = getenv("B2C_FUSE_SPLIT");
    const bool fuse_split = fuse_res && (!ed || atoi(ed) != 0);
    std::map<Blob*, BatchNormLayer*> tails;
    for (size_t j = 0; j < layers_.size(); ++j)
      if (auto* bn = dynamic_cast<BatchNormLayer*>(layers_[j].get()))
        if (bn->residual_sum()) tails[bn->residual_sum()] = bn;
    std::map<Blob*, int> shadows;
    for (Node& nd : nodes_)
      for (size_t i = 0; i < nd.bottom.size(); ++i) if (nd.bottom_diff_tmp[i]) ++shadows[nd.bottom[i]];
    for (Node& nd : nodes_) {
      nd.deferred_add.assign(nd.bottom.size(), false);
      for (size_t i = 0; fuse_split && i < nd.bottom.size(); ++i) {
        if (!nd.bottom_diff_tmp[i] || shadows[nd.bottom[i]] != 1) continue;       // exactly one shadow per blob
        auto it = tails.find(nd.bottom[i]);
        if (it == tails.end()) continue;
        it->second->set_sum_diff_part2(nd.bottom_diff_tmp[i]);
        nd.deferred_add[i] = true;
      }
    }
  }
  solver_.reset(new SGDSolver(sp));
  solver_->SetParams(learnable_, specs);
  for (auto& l : layers_)
    if (auto* c = dynamic_cast<ConvolutionLayer*>(l.get()))
      if (c->EnableFilterCache()) cached_convs_.push_back(c);
  sched_.reset(new ReduceScheduler(solver_.get(), nullptr));
  CUDA_CHECK(cudaStreamSynchronize(S()));
}
TrainNet::~TrainNet() {}

void TrainNet::AttachSync(P2PSync* sync) {
  sync_ = sync;
  if (db_data_) db_data_->set_solver(sync->nranks(), sync->rank());     // each solver reads its own stripe of the database
  sync_->on_start(solver_->arena());
  filters_dirty_ = true;
  sched_.reset(new ReduceScheduler(solver_.get(), sync_));
}
Blob* TrainNet::blob(const string& name) {
  auto it = blobs_.find(name);
  return it == blobs_.end() ? nullptr : it->second.get();
}
size_t TrainNet::activation_floats() const {
  size_t c = 0;
  for (auto& kv : blobs_) c += kv.second->count();
  return c;
}
void TrainNet::PrepareFilters() {
  // one multi-tensor launch per 24 filter layouts instead of one prepass per forward and per dgrad call
  vector<const b2c_conv_desc*> descs;
  vector<const float*> ws;
  vector<void*> caches;
  for (ConvolutionLayer* c : cached_convs_) {
    if (!c->desc() || !c->filter_cache()) continue;
    descs.push_back(c->desc()); ws.push_back(c->blobs()[0]->gpu_data()); caches.push_back(c->filter_cache());
  }
  if (!descs.empty()) B2C_CHECK(b2c_conv_prepare_filters((int)descs.size(), descs.data(), ws.data(), caches.data(), S()));
  filters_dirty_ = false;
}
void TrainNet::Forward(bool copy_input) {
  if (filters_dirty_) PrepareFilters();
  if (db_data_) {                                                        // database source: a new batch per e2e step; the first one always
    if (copy_input || !db_data_->loaded()) db_data_->LoadBatch(nodes_[db_data_node_].top, S());
  } else if (copy_input && data_) data_->LoadBatch(nodes_[0].top[0], S());     // H2D of the batch (+ the device transform of uint8 datums)
  EventProfiler* prof = Caffe::profiler();
  for (size_t i = 0; i < layers_.size(); ++i) {
    size_t h = 0;
    if (prof) { prof->set_layer((int)i); h = prof->begin(EventProfiler::FWD, S()); }
    layers_[i]->Forward(nodes_[i].bottom, nodes_[i].top);
    if (prof) prof->end(h, S());
  }
}
void TrainNet::Backward(bool update) {
  for (int li = (int)layers_.size() - 1; li >= 0; --li) {
    Node& nd = nodes_[li];
    EventProfiler* prof = Caffe::profiler();
    size_t ph = 0;
    if (prof && nd.need_backward) { prof->set_layer(li); ph = prof->begin(EventProfiler::BWD, S()); }
    if (nd.need_backward) {
      // redirect fan-out bottoms to their shadow blobs (data shared, private diff), run, then accumulate
      vector<Blob*> bvec = nd.bottom;
      for (size_t i = 0; i < bvec.size(); ++i)
        if (nd.bottom_diff_tmp[i]) { nd.bottom_diff_tmp[i]->ShareData(*nd.bottom[i]); bvec[i] = nd.bottom_diff_tmp[i]; }
      if (!nd.accumulate_bottom.empty()) layers_[li]->set_accumulate_bottom_diff(nd.accumulate_bottom);
      layers_[li]->Backward(nd.top, nd.propagate_down, bvec);
      for (size_t i = 0; i < bvec.size(); ++i)
        if (nd.bottom_diff_tmp[i] && !(i < nd.deferred_add.size() && nd.deferred_add[i]))
          B2C_CHECK(b2c_add(nd.bottom[i]->count(), nd.bottom[i]->gpu_diff(), nd.bottom_diff_tmp[i]->gpu_diff(), nd.bottom[i]->mutable_gpu_diff(), S()));
      if (prof) prof->end(ph, S());
    }
    if (update && nd.num_params) sched_->on_param_ready(nd.first_param, S());     // net.cpp:738-746
  }
}
float TrainNet::ForwardBackward() {
  Forward(false);
  Backward(false);              // diffs stay in place for inspection; Step() is the path that reduces + updates
  return last_loss();
}
void TrainNet::ClearParamDiffs() {
  // the reference's solver relies on ApplyUpdate clearing the diffs (solver.cpp:237-239) and so does Step(); a caller that
  // ran ForwardBackward() for inspection clears them before stepping, as with Net::ClearParamDiffs
  CUDA_CHECK(cudaMemsetAsync(solver_->arena().diff(), 0, sizeof(float) * solver_->arena().total(), S()));
}
void TrainNet::Step(bool copy_input_from_host) {
  // Solver::Step (solver.cpp:277-288): iter_size micro-batches accumulate into the parameter diffs (weight / bias
  // gradients are accumulated by every layer, as in the reference); only the last one releases parameters to the
  // reduction queue (ForwardBackward(apply_update = i + 1 == iter_size)), and the update folds 1/iter_size in.
  const int iter_size = std::max(1, solver_->param().iter_size);
  for (int i = 0; i < iter_size; ++i) {
    Forward(copy_input_from_host);
    Backward(i + 1 == iter_size);
  }
  sched_->end_of_iteration(S());
  filters_dirty_ = true;          // the update changed every weight: the next Forward re-prepares (after the compute stream's wait)
}
float TrainNet::TimedSteps(int n, bool copy_input, bool read_loss) {
  cudaEvent_t a, b;
  CUDA_CHECK(cudaEventCreate(&a)); CUDA_CHECK(cudaEventCreate(&b));
  CUDA_CHECK(cudaStreamSynchronize(S()));
  CUDA_CHECK(cudaEventRecord(a, S()));
  for (int i = 0; i < n; ++i) {
    Step(copy_input);
    if (read_loss) {                               // device -> host read of the step's result, on the same stream
      CUDA_CHECK(cudaMemcpyAsync(&host_loss_, loss_blob_->gpu_data(), sizeof(float), cudaMemcpyDeviceToHost, S()));
      CUDA_CHECK(cudaStreamSynchronize(S()));
    }
  }
  CUDA_CHECK(cudaEventRecord(b, S()));
  CUDA_CHECK(cudaEventSynchronize(b));
  float ms = 0.f;
  CUDA_CHECK(cudaEventElapsedTime(&ms, a, b));
  cudaEventDestroy(a); cudaEventDestroy(b);
  return ms;
}
void TrainNet::ProfileSteps(int n, vector<int>* layer, vector<int>* op, vector<float>* ms) {
  EventProfiler prof;
  CUDA_CHECK(cudaDeviceSynchronize());
  Caffe::set_profiler(&prof);
  try { for (int i = 0; i < n; ++i) Step(false); } catch (...) { Caffe::set_profiler(nullptr); throw; }
  Caffe::set_profiler(nullptr);
  std::map<std::pair<int, int>, float> acc;
  prof.collect(&acc);
  for (auto& kv : acc) { layer->push_back(kv.first.first); op->push_back(kv.first.second); ms->push_back(kv.second / (float)n); }
}
// ---- snapshot / restore -------------------------------------------------------------------------------------------
static BlobData to_blob_data(const vector<int>& shape, const float* host, size_t count) {
  BlobData b;
  b.shape = shape;
  b.data.assign(host, host + count);
  return b;
}
string TrainNet::Snapshot(const string& prefix) {
  CUDA_CHECK(cudaDeviceSynchronize());
  NetWeights nw;
  nw.name = name_;
  for (size_t i = 0; i < layers_.size(); ++i) {
    LayerWeights lw;
    lw.name = layer_names_[i]; lw.type = layer_types_[i];
    // read the DEVICE copy explicitly: the fused SGD kernel updates the arena behind the Blob's back, so a blob whose
    // head was left SYNCED by an earlier cpu_data() would hand out stale host values (second snapshot of a process)
    for (auto& b : layers_[i]->blobs()) {
      vector<float> h(b->count());
      CUDA_CHECK(cudaMemcpy(h.data(), b->gpu_data(), sizeof(float) * h.size(), cudaMemcpyDeviceToHost));
      lw.blobs.push_back(to_blob_data(b->shape(), h.data(), h.size()));
    }
    nw.layers.push_back(std::move(lw));
  }
  const string stem = prefix + "_iter_" + std::to_string(solver_->iter());
  WriteBinaryFile(stem + ".caffemodel", SerializeNetWeights(nw));
  SolverStateData st;
  st.iter = solver_->iter(); st.current_step = solver_->current_step(); st.learned_net = stem + ".caffemodel";
  ParamArena& ar = solver_->arena();
  for (size_t i = 0; i < learnable_.size(); ++i) {
    vector<float> h(learnable_[i]->count());
    CUDA_CHECK(cudaMemcpy(h.data(), ar.history() + ar.offset((int)i), sizeof(float) * h.size(), cudaMemcpyDeviceToHost));
    st.history.push_back(to_blob_data(learnable_[i]->shape(), h.data(), h.size()));
  }
  WriteBinaryFile(stem + ".solverstate", SerializeSolverState(st));
  return stem + ".solverstate";
}
int TrainNet::CopyTrainedLayersFrom(const string& path) {
  const NetWeights nw = ParseNetWeights(ReadBinaryFile(path));
  CUDA_CHECK(cudaDeviceSynchronize());
  int copied = 0;
  for (const LayerWeights& lw : nw.layers) {
    size_t li = 0;
    while (li < layer_names_.size() && layer_names_[li] != lw.name) ++li;
    if (li == layer_names_.size() || lw.blobs.empty()) continue;          // "Ignoring source layer" (net.cpp)
    auto& target = layers_[li]->blobs();
    B2_CHECK(target.size() == lw.blobs.size(), "Incompatible number of blobs for layer " + lw.name);
    for (size_t j = 0; j < target.size(); ++j) {
      B2_CHECK(target[j]->count() == lw.blobs[j].data.size(), "Cannot copy param " + std::to_string(j) + " weights from layer '" + lw.name +
                                                                  "'; shape mismatch.");
      CUDA_CHECK(cudaMemcpy(target[j]->mutable_gpu_data(), lw.blobs[j].data.data(), sizeof(float) * target[j]->count(), cudaMemcpyHostToDevice));
    }
    ++copied;
  }
  filters_dirty_ = true;
  return copied;
}
void TrainNet::Restore(const string& state_path) {
  const SolverStateData st = ParseSolverState(ReadBinaryFile(state_path));
  if (!st.learned_net.empty()) CopyTrainedLayersFrom(st.learned_net);
  B2_CHECK(st.history.size() == learnable_.size(), "Incorrect length of history blobs.");      // sgd_solver.cpp:334
  ParamArena& ar = solver_->arena();
  for (size_t i = 0; i < learnable_.size(); ++i) {
    B2_CHECK(st.history[i].data.size() == learnable_[i]->count(), "history blob " + std::to_string(i) + ": size mismatch");
    CUDA_CHECK(cudaMemcpy(ar.history() + ar.offset((int)i), st.history[i].data.data(), sizeof(float) * learnable_[i]->count(), cudaMemcpyHostToDevice));
  }
  solver_->set_iter(st.iter);
  solver_->set_current_step(st.current_step);
  for (auto& l : layers_)
    if (auto* bn = dynamic_cast<BatchNormLayer*>(l.get())) bn->set_iter(st.iter + 1);   // running statistics are warm
}
float TrainNet::last_loss() { return loss_blob_ ? loss_blob_->cpu_data()[0] : 0.f; }

}  // namespace caffe
