// train_net.hpp -- the non-convolution layers of the BASELINE nets as caffe::LayerBase subclasses, and TrainNet: a
// caffe::Net executor (forward / backward over the prototxt graph with automatic diff accumulation where a blob fans
// out, i.e. what insert_splits.cpp + SplitLayer do in the reference) wired to SGDSolver / ReduceScheduler / P2PSync.
//
// Reference map: src/caffe/layers/{relu,batch_norm,pooling,eltwise,inner_product,softmax_loss,split}_layer.cpp,
// Net::ForwardFromTo / BackwardFromToAu (src/caffe/net.cpp:669-751), Solver::Step (src/caffe/solver.cpp:187-353).
#pragma once
#include "b2caffe.hpp"
#include "prototxt.hpp"
#include "data_layer.hpp"

namespace caffe {

class ReLULayer : public LayerBase {
 public:
  ReLULayer(const LayerParameter& p, float negative_slope) : LayerBase(p), slope_(negative_slope) {}
  const char* type() const override { return "ReLU"; }
  void Reshape(const vector<Blob*>& b, const vector<Blob*>& t) override { if (t[0] != b[0]) t[0]->ReshapeLike(*b[0]); }
 protected:
  void Forward_gpu(const vector<Blob*>& b, const vector<Blob*>& t) override;
  void Backward_gpu(const vector<Blob*>& t, const vector<bool>& pd, const vector<Blob*>& b) override;
  float slope_;     // relu_param.negative_slope (relu_layer.cpp:13-16)
  bool fused_away_ = false;
 public:
  float negative_slope() const { return slope_; }
  void set_fused_away(bool v) { fused_away_ = v; }   // the producing BatchNorm / Eltwise layer applies (and back-propagates) this ReLU
};

class BatchNormLayer : public LayerBase {      // NVCaffe BatchNorm with scale_bias (batch_norm_layer.cpp)
 public:
  BatchNormLayer(const LayerParameter& p, bool scale_bias, float eps, float maf) : LayerBase(p), scale_bias_(scale_bias), eps_(eps), maf_(maf) {
    scale_filler_.type = "constant"; scale_filler_.value = 1.f;
  }
  void set_scale_filler(const FillerParameter& f) { scale_filler_ = f; }
  void set_bias_filler(const FillerParameter& f) { bias_filler_ = f; }
  const char* type() const override { return "BatchNorm"; }
  void LayerSetUp(const vector<Blob*>& b, const vector<Blob*>& t) override;
  void Reshape(const vector<Blob*>& b, const vector<Blob*>& t) override;
 protected:
  void Forward_gpu(const vector<Blob*>& b, const vector<Blob*>& t) override;
  void Backward_gpu(const vector<Blob*>& t, const vector<bool>& pd, const vector<Blob*>& b) override;
  bool scale_bias_;
  float eps_, maf_;
  FillerParameter scale_filler_, bias_filler_;
  bool fuse_relu_ = false;   // an in-place ReLU on the top blob is folded into this layer's kernels (TrainNet's fusion pass)
  bool recompute_ = false;   // do not store x_norm: backward recomputes it from the bottom blob (needs top != bottom)
  int iter_ = 0;
 public:
  void set_iter(int i) { iter_ = i; }
  void set_fusion(bool recompute_xnorm, bool fuse_relu) { recompute_ = recompute_xnorm; fuse_relu_ = fuse_relu; }
  bool recompute() const { return recompute_; }
  bool fuse_relu() const { return fuse_relu_; }
  // Residual tail: this layer also does the Eltwise SUM (+ in-place ReLU) that consumes its top.  Forward writes
  // max(0, BatchNorm(x) + other) into sum_top; backward takes sum_top's diff / data and also writes other's diff.
  void set_residual(Blob* other, Blob* sum_top, bool propagate_other) { res_other_ = other; res_sum_ = sum_top; res_prop_ = propagate_other; }
  Blob* residual_sum() const { return res_sum_; }
  // the sum's top has a second consumer whose bottom gradient lands in a shadow blob: this layer's backward reads both parts and
  // adds them on the fly instead of TrainNet running b2c_add over them first
  void set_sum_diff_part2(Blob* shadow) { res_part2_ = shadow; }
 protected:
  Blob* res_other_ = nullptr;
  Blob* res_sum_ = nullptr;
  Blob* res_part2_ = nullptr;
  bool res_prop_ = false;
  Blob xnorm_, save_mean_, save_invstd_, scratch_;
};

class PoolingLayer : public LayerBase {
 public:
  PoolingLayer(const LayerParameter& p, const PoolingParameter& q) : LayerBase(p), q_(q) {}
  const char* type() const override { return "Pooling"; }
  void Reshape(const vector<Blob*>& b, const vector<Blob*>& t) override;
 protected:
  void Forward_gpu(const vector<Blob*>& b, const vector<Blob*>& t) override;
  void Backward_gpu(const vector<Blob*>& t, const vector<bool>& pd, const vector<Blob*>& b) override;
  PoolingParameter q_, eff_;
  int* mask_ = nullptr;
  size_t mask_cap_ = 0;
 public:
  ~PoolingLayer() override;
};

class EltwiseLayer : public LayerBase {        // SUM with unit coefficients (the only use in the BASELINE nets)
 public:
  using LayerBase::LayerBase;
  void set_fuse_relu(bool v) { fuse_relu_ = v; }
  bool fuse_relu() const { return fuse_relu_; }
  void set_fused_away(bool v) { fused_away_ = v; }     // the BatchNorm layer that produces one of the bottoms does the sum (TrainNet's fusion pass)
 protected:
  bool fuse_relu_ = false;
  bool fused_away_ = false;
 public:
  const char* type() const override { return "Eltwise"; }
  void Reshape(const vector<Blob*>& b, const vector<Blob*>& t) override { t[0]->ReshapeLike(*b[0]); }
 protected:
  void Forward_gpu(const vector<Blob*>& b, const vector<Blob*>& t) override;
  void Backward_gpu(const vector<Blob*>& t, const vector<bool>& pd, const vector<Blob*>& b) override;
};

class InnerProductLayer : public LayerBase {
 public:
  InnerProductLayer(const LayerParameter& p, int num_output, bool bias, const FillerParameter& wf, const FillerParameter& bf)
      : LayerBase(p), num_output_(num_output), bias_(bias), wf_(wf), bf_(bf) {}
  const char* type() const override { return "InnerProduct"; }
  void LayerSetUp(const vector<Blob*>& b, const vector<Blob*>& t) override;
  void Reshape(const vector<Blob*>& b, const vector<Blob*>& t) override;
 protected:
  void Forward_gpu(const vector<Blob*>& b, const vector<Blob*>& t) override;
  void Backward_gpu(const vector<Blob*>& t, const vector<bool>& pd, const vector<Blob*>& b) override;
  int num_output_, K_ = 0, M_ = 0;
  bool bias_;
  FillerParameter wf_, bf_;
  // forward runs on the tensor-core NoTrans x Trans GEMM (b2c_sgemm_ex): split-K scratch, and a 16-byte aligned copy of the
  // weights when their slot in the parameter arena is only 8-byte aligned (slots are padded to even counts, net.cpp:1356-1371;
  // TMA wants 16).  The copy is 4 bytes per weight once per forward -- against a 4x slower FFMA kernel.
  Blob gemm_ws_, w_aligned_;
  // backward on the same GEMM: dW[N x K] += dy^T x and dx[M x K] = dy W are NoTrans x Trans products of TRANSPOSED operand copies
  // (dy^T [N x M], x^T [K x M]; W^T [K x N]) -- the reduction axis must be the contiguous one.  The copies are one pass over
  // each operand; the FFMA kernel they replace runs at ~20 TFLOP/s (2.9 of AlexNet's 17 ms step, 1.4 of VGG-16's 19).
  Blob dyt_, xt_, wt_, bwd_ws_;
  bool bwd_w_tc_ = false, bwd_x_tc_ = false;
};

// ---- layers the AlexNet / GoogLeNet / VGG-16 nets add (parity: tests/test_layers_extra_gpu.py, test_inception_style_net_matches_oracle) ----
class LRNLayer : public LayerBase {            // ACROSS_CHANNELS only (the BASELINE nets' use)
 public:
  LRNLayer(const LayerParameter& p, int size, float alpha, float beta, float k) : LayerBase(p), size_(size), alpha_(alpha), beta_(beta), k_(k) {}
  const char* type() const override { return "LRN"; }
  void Reshape(const vector<Blob*>& b, const vector<Blob*>& t) override { t[0]->ReshapeLike(*b[0]); scale_.ReshapeLike(*b[0]); }
 protected:
  void Forward_gpu(const vector<Blob*>& b, const vector<Blob*>& t) override;
  void Backward_gpu(const vector<Blob*>& t, const vector<bool>& pd, const vector<Blob*>& b) override;
  int size_;
  float alpha_, beta_, k_;
  Blob scale_;
};

class DropoutLayer : public LayerBase {        // TRAIN phase: y = x * mask, mask in {0, 1/(1-ratio)}
 public:
  DropoutLayer(const LayerParameter& p, float ratio, uint64_t seed) : LayerBase(p), ratio_(ratio), seed_(seed) {}
  const char* type() const override { return "Dropout"; }
  void Reshape(const vector<Blob*>& b, const vector<Blob*>& t) override { if (t[0] != b[0]) t[0]->ReshapeLike(*b[0]); mask_.ReshapeLike(*b[0]); }
 protected:
  void Forward_gpu(const vector<Blob*>& b, const vector<Blob*>& t) override;
  void Backward_gpu(const vector<Blob*>& t, const vector<bool>& pd, const vector<Blob*>& b) override;
  float ratio_;
  uint64_t seed_, offset_ = 0;
  Blob mask_;
};

class ConcatLayer : public LayerBase {         // along the channel axis (axis 1), strided 2-D copies
 public:
  using LayerBase::LayerBase;
  const char* type() const override { return "Concat"; }
  void Reshape(const vector<Blob*>& b, const vector<Blob*>& t) override;
 protected:
  void Forward_gpu(const vector<Blob*>& b, const vector<Blob*>& t) override;
  void Backward_gpu(const vector<Blob*>& t, const vector<bool>& pd, const vector<Blob*>& b) override;
};

class SoftmaxWithLossLayer : public LayerBase {
 public:
  using LayerBase::LayerBase;
  void set_loss_weight(float w) { loss_weight_ = w; }
  float loss_weight() const { return loss_weight_; }
  const char* type() const override { return "SoftmaxWithLoss"; }
  void Reshape(const vector<Blob*>& b, const vector<Blob*>& t) override;
 protected:
  void Forward_gpu(const vector<Blob*>& b, const vector<Blob*>& t) override;
  void Backward_gpu(const vector<Blob*>& t, const vector<bool>& pd, const vector<Blob*>& b) override;
  Blob prob_;
  float loss_weight_ = 1.f;
};

class AccuracyLayer : public LayerBase {       // forward only (accuracy_layer.cpp); top[0] is a scalar
 public:
  AccuracyLayer(const LayerParameter& p, int top_k) : LayerBase(p), top_k_(top_k) {}
  const char* type() const override { return "Accuracy"; }
  void Reshape(const vector<Blob*>& b, const vector<Blob*>& t) override {
    B2_CHECK(top_k_ <= (int)(b[0]->count() / b[1]->count()), "top_k must be less than or equal to the number of classes.");
    t[0]->Reshape({1}); hits_.Reshape({1});
  }
 protected:
  void Forward_gpu(const vector<Blob*>& b, const vector<Blob*>& t) override;
  void Backward_gpu(const vector<Blob*>&, const vector<bool>&, const vector<Blob*>&) override {}   // "Accuracy cannot backpropagate"
  int top_k_;
  Blob hits_;
};

// Synthetic in-memory source standing in for DataLayer's LMDB reader.  Two forms:
//   * float: N(0,1) images from mt19937(seed), uniform labels (Input / DummyData layers, nets without transform_param);
//   * datum: a pinned batch of uint8 datums (crop_size + 32 on a side, like the 256x256 ImageNet LMDBs for a 224 crop) that
//     goes through the reference's own path every e2e step: host -> device copy of the BYTES, then DataTransformer::Transform
//     on the device (b2c_transform_u8: random crop window and mirror per image drawn with Caffe's rule rand % (extent), mean
//     values, scale) into the float data blob.
class SyntheticDataLayer : public LayerBase {
 public:
  struct Transform { bool on = false, mirror = false; int crop = 0; float scale = 1.f; vector<float> mean_value; };
  SyntheticDataLayer(const LayerParameter& p, const vector<vector<int>>& shapes, int num_classes, uint64_t seed, const Transform& tf)
      : LayerBase(p), shapes_(shapes), classes_(num_classes), seed_(seed), tf_(tf) {}
  const char* type() const override { return "Data"; }
  void LayerSetUp(const vector<Blob*>& b, const vector<Blob*>& t) override;
  void Reshape(const vector<Blob*>&, const vector<Blob*>&) override {}
  // one batch from pinned host memory into top[0] on `stream` (the e2e step): float copy, or uint8 copy + device transform
  void LoadBatch(Blob* top, cudaStream_t stream);
  size_t h2d_bytes() const { return datum_mode_ ? u8_bytes_ + 9 * (size_t)shapes_[0][0] : sizeof(float) * n0_; }
  bool datum_mode() const { return datum_mode_; }
  ~SyntheticDataLayer() override;
 protected:
  void Forward_gpu(const vector<Blob*>&, const vector<Blob*>&) override {}
  void Backward_gpu(const vector<Blob*>&, const vector<bool>&, const vector<Blob*>&) override {}
  vector<vector<int>> shapes_;
  int classes_;
  uint64_t seed_;
  Transform tf_;
  float* host_ = nullptr;
  size_t n0_ = 0;
  // datum mode
  bool datum_mode_ = false;
  int hd_ = 0, wd_ = 0;
  size_t u8_bytes_ = 0;
  unsigned char* host_u8_ = nullptr;      // pinned [N][C][hd][wd]
  float* dev_mean_ = nullptr;
  uint64_t draws_ = 0;                    // batches drawn so far (crop / mirror stream position)
  // Prefetch (BasePrefetchingDataLayer's thread, as a copy stream): while step i computes, batch i+1's datums and its crop / mirror
  // draws travel host -> device into the other of two slots; LoadBatch makes the compute stream wait for its slot, transforms out
  // of it and starts the next copy.  Every step still moves one batch over PCIe; the copy no longer sits in front of the forward.
  struct Slot {
    unsigned char* dev_u8 = nullptr;
    int* host_off = nullptr;              // pinned: h_off[N], w_off[N], then mirror[N] as bytes
    int* dev_off = nullptr;
    cudaEvent_t copied = nullptr, consumed = nullptr;
    bool in_flight = false, used = false;
  } slot_[2];
  int cur_ = 0;
  cudaStream_t copy_stream_ = nullptr;
  void IssueCopy(int s);
};

class TrainNet {
 public:
  TrainNet(const Net& net, const SolverParameter& sp, int num_classes, uint64_t seed, int math);
  ~TrainNet();
  void AttachSync(P2PSync* sync);                  // weights broadcast from rank 0, gradients exchanged per bucket
  float ForwardBackward();                         // Net::ForwardBackward (net.cpp:711-716); returns the loss (host sync)
  void ClearParamDiffs();                          // Net::ClearParamDiffs (net.cpp:1256-1270): zero every learnable blob's diff
  void Step(bool copy_input_from_host);            // one Solver::Step iteration, fully asynchronous on the thread stream
  // n Steps bracketed by CUDA events on the thread stream; returns milliseconds (end-of-iteration makes the compute
  // stream wait for the last update, so the closing event covers reduce + update too)
  float TimedSteps(int n, bool copy_input, bool read_loss);
  // n Steps with CUDA events around every layer call (and around the conv weight- / data-gradient calls inside Backward);
  // per (layer, EventProfiler::Op) mean milliseconds per step.  Events serialise nothing, but the numbers are per-call device
  // times on the compute stream, not a decomposition of the overlapped step.
  void ProfileSteps(int n, vector<int>* layer, vector<int>* op, vector<float>* ms);
  const string& layer_name(int i) const { return layer_names_[i]; }
  const string& layer_type(int i) const { return layer_types_[i]; }
  float last_loss();                               // device -> host read of the loss blob
  SGDSolver& solver() { return *solver_; }
  Blob* blob(const string& name);
  int num_layers() const { return (int)layers_.size(); }
  LayerBase* layer(int i) { return layers_[i].get(); }
  const vector<shared_ptr<Blob>>& learnable_params() const { return learnable_; }   // Net::learnable_params(): every layer blob
  const vector<int>& trainable_ids() const { return trainable_ids_; }               // ids the layers differentiate
  size_t activation_floats() const;
  size_t input_bytes() const { return db_data_ ? db_data_->h2d_bytes() : data_ ? data_->h2d_bytes() : 0; }   // host -> device bytes of one e2e step's batch
  DataLayer* database_layer() { return db_data_; }   // non-null when the net's Data layer reads its LMDB (data_layer.hpp)
  // Solver::Snapshot (solver.cpp:447-520): <prefix>_iter_<N>.caffemodel (every layer's blobs, NVCaffe raw BlobProto) and
  // <prefix>_iter_<N>.solverstate (iter, learned_net, one history blob per learnable parameter, current_step).
  // Returns the .solverstate path.
  string Snapshot(const string& prefix);
  void Restore(const string& solverstate_path);          // SGDSolver::RestoreSolverStateFromBinaryProto + weights
  // Net::CopyTrainedLayersFrom (net.cpp): layers matched by name, blob shapes must agree; returns the number of layers copied
  int CopyTrainedLayersFrom(const string& caffemodel_path);
  const string& name() const { return name_; }
 private:
  struct Node {
    vector<Blob*> bottom, top;
    vector<bool> propagate_down;
    vector<Blob*> bottom_diff_tmp;                 // per bottom: null = write the blob's own diff, else accumulate through this
    vector<bool> accumulate_bottom;                // per bottom: the layer itself adds into the blob's diff (no shadow blob)
    vector<bool> deferred_add;                     // per bottom: the shadow diff is added by the blob's producer (fused BatchNorm residual tail)
    bool need_backward = false;
    int first_param = -1, num_params = 0;
  };
  void Forward(bool copy_input);
  void Backward(bool update);
  void PrepareFilters();                           // b2c_conv_prepare_filters over every cached conv layer (weights changed)
 public:
  void MarkParamsDirty() { filters_dirty_ = true; } // call after writing parameter data behind the net's back
 private:
  bool filters_dirty_ = true;
  bool fuse_fanout_ = true;                        // B2C_FUSE: conv layers add their bottom gradient into a fan-out blob's diff directly
  vector<ConvolutionLayer*> cached_convs_;
  float host_loss_ = 0.f;
  string name_;
  vector<string> layer_names_, layer_types_;
  std::map<string, shared_ptr<Blob>> blobs_;
  vector<shared_ptr<LayerBase>> layers_;
  vector<Node> nodes_;
  vector<shared_ptr<Blob>> tmp_diffs_;
  vector<shared_ptr<Blob>> learnable_;
  vector<int> trainable_ids_;
  std::unique_ptr<SGDSolver> solver_;
  std::unique_ptr<ReduceScheduler> sched_;
  P2PSync* sync_ = nullptr;
  Blob* loss_blob_ = nullptr;
  SyntheticDataLayer* data_ = nullptr;
  DataLayer* db_data_ = nullptr;                   // the LMDB-backed source, when data_param.source opened
  int db_data_node_ = 0;
};

}  // namespace caffe
