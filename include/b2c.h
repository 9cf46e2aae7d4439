/*
 * b2c.h -- C ABI of the B200-native Caffe-MPI data-parallel hot path
 *          (ConvolutionLayer forward/backward, im2col/col2im, SGEMM/SGEMV, fused
 *          SGD-momentum update, NCCL gradient exchange).
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch or
 * C++ types.  Every entry point names the reference interface it replaces (paths
 * relative to the Caffe-MPI tree).  All tensor pointers are DEVICE pointers, fp32,
 * dense row-major; activations NCHW, weights [O, C/g, kh, kw], bias [O].  `stream`
 * is the caller's cudaStream_t passed as void* (NULL = legacy default stream).
 * Calls are asynchronous on `stream`; nothing host-synchronises (the reference syncs
 * after every BLAS call, math_functions.cu:25 -- callers that need that behaviour
 * synchronise the stream themselves).
 *
 * Error model: the reference aborts through glog CHECK / CUDA_CHECK
 * (include/caffe/util/nccl.hpp:10-15).  Here every function returns 0 on success or
 * a negative b2c_status and leaves a thread-local message readable with
 * b2c_last_error(); the C++ layer shim turns non-zero into a fatal check.
 * There is NO CPU fallback: on a machine without a CUDA device every compute entry
 * point returns B2C_ERR_CUDA.
 */
#ifndef B2C_H_
#define B2C_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  B2C_OK = 0,
  B2C_ERR_INVALID = -1,     /* bad argument / unsupported shape            */
  B2C_ERR_CUDA = -2,        /* CUDA runtime / launch error                 */
  B2C_ERR_NCCL = -3,        /* NCCL error                                  */
  B2C_ERR_WORKSPACE = -4    /* workspace too small                         */
} b2c_status;

/* ConvolutionParameter.engine, src/caffe/proto/caffe.proto:745-750 (same values).
 * CAFFE  -> explicit per-image im2col + GEMM (base_conv_layer.hpp:105-168).
 * CUDNN / DEFAULT -> implicit-GEMM kernels (replaces cudnn_conv_layer.cu:13-132). */
typedef enum { B2C_ENGINE_DEFAULT = 0, B2C_ENGINE_CAFFE = 1, B2C_ENGINE_CUDNN = 2 } b2c_engine;

/* Math mode of the tensor-core kernels (NetParameter default_forward_math /
 * default_backward_math, caffe.proto:124-127, is the reference's analogous knob).
 * FP32        : split-precision tcgen05 MMA, fp32-equivalent results (default): bf16 hi/lo split, three
 *               kind::f16 MMAs per K step, on the layers the bulk-copy-staged kernel takes (stride-1 "same"
 *               convolutions, forward and dgrad); 3xTF32 split everywhere else.  Blob error ~1e-6..1e-5.
 * TF32        : single-pass TF32 (round-to-nearest operands), ~1e-3 relative.
 * FP32_3XTF32 : the 3xTF32 split on every layer (the round-1 default; ~1e-7..4e-5).          */
typedef enum { B2C_MATH_FP32 = 0, B2C_MATH_TF32 = 1, B2C_MATH_FP32_3XTF32 = 2 } b2c_math;

/* Kernel family used by the implicit-GEMM engine (diagnostics / tests). */
typedef enum { B2C_ALGO_AUTO = 0, B2C_ALGO_SIMT = 1, B2C_ALGO_TCGEN05 = 2 } b2c_algo;

typedef enum { B2C_OP_FORWARD = 0, B2C_OP_BACKWARD_DATA = 1, B2C_OP_BACKWARD_FILTER = 2 } b2c_conv_op;

/* Shape of one ConvolutionLayer call: ConvolutionParameter (caffe.proto:718-786) plus
 * the bottom shape BaseConvolutionLayer::Reshape reads (base_conv_layer.cpp:172-239). */
typedef struct {
  int N, C, H, W;     /* bottom blob                                              */
  int O, G;           /* num_output, group                                        */
  int kh, kw;         /* kernel_size / kernel_h,kernel_w                          */
  int sh, sw;         /* stride                                                   */
  int ph, pw;         /* pad                                                      */
  int dh, dw;         /* dilation                                                 */
  int has_bias;       /* bias_term                                                */
} b2c_conv_params;

typedef struct b2c_conv_desc b2c_conv_desc;
typedef struct b2c_comm b2c_comm;

/* ---- library ------------------------------------------------------------------- */
const char* b2c_last_error(void);
const char* b2c_version(void);
/* number of kernels THIS library has launched in this process (all threads).      */
uint64_t b2c_launch_count(void);
/* process-wide defaults used by descriptors created afterwards                    */
int b2c_set_default_math(int math /* b2c_math */);
int b2c_set_default_algo(int algo /* b2c_algo */);

/* ---- convolution descriptor -------------------------------------------------------
 * Replaces BaseConvolutionLayer::LayerSetUp/Reshape bookkeeping (base_conv_layer.cpp:12-239)
 * and CuDNNConvolutionLayer's descriptor set-up (cudnn_conv_layer.cpp:78-210).        */
int b2c_conv_desc_create(const b2c_conv_params* p, int engine, b2c_conv_desc** out);
int b2c_conv_desc_destroy(b2c_conv_desc* d);
int b2c_conv_desc_set_math(b2c_conv_desc* d, int math);
int b2c_conv_desc_set_algo(b2c_conv_desc* d, int algo);
/* output extent, conv_layer.cpp:7-22 (truncating division)                          */
int b2c_conv_out_shape(const b2c_conv_desc* d, int* Ho, int* Wo);
/* scratch bytes the given op needs (col buffer for CAFFE engine: base_conv_layer.cpp:225-233;
 * split-K partials / transformed weights for the implicit engine).                   */
size_t b2c_conv_workspace_bytes(const b2c_conv_desc* d, int op);
/* which kernel family the op will run (b2c_algo) -- lets tests assert "tcgen05 ran". */
int b2c_conv_algo_used(const b2c_conv_desc* d, int op);

/* Prepared filters.  The tcgen05 forward / dgrad kernels read the filter through TMA in a GEMM-ordered, hi/lo-split copy
 * that b2c_conv_forward / b2c_conv_backward_data otherwise rebuild on every call.  Weights change once per iteration
 * (SGDSolver::ApplyUpdate, sgd_solver.cpp:143-149), so a caller that owns the iteration can build the copies ONCE:
 *   bytes = b2c_conv_filter_cache_bytes(d);            cudaMalloc(&cache, bytes)  (256-byte aligned; 0 = nothing to cache)
 *   b2c_conv_desc_bind_filter_cache(d, cache);          forward / backward_data now trust the cache
 *   b2c_conv_prepare_filters(n, descs, ws, caches, s)   after every weight update: ONE multi-tensor launch per 24 layouts
 * Binding NULL restores the self-contained behaviour.  The cache is the caller's promise that it was prepared from the `w`
 * passed to forward / backward_data; nothing checks it.  (cuDNN's analogue: cudnnTransformFilter + a persistent
 * filter descriptor; the reference re-reads the blob each call, cudnn_conv_layer.cu:25-29.) */
size_t b2c_conv_filter_cache_bytes(const b2c_conv_desc* d);
int b2c_conv_desc_bind_filter_cache(b2c_conv_desc* d, const void* cache);
int b2c_conv_prepare_filter(const b2c_conv_desc* d, const float* w, void* cache, size_t cache_bytes, void* stream);
int b2c_conv_prepare_filters(int n, const b2c_conv_desc* const* descs, const float* const* ws, void* const* caches,
                             void* stream);


/* Y = conv(X, W) (+ bias).  y OVERWRITTEN.
 * Replaces ConvolutionLayer::Forward_gpu (conv_layer.cu:7-23) = forward_gpu_gemm +
 * forward_gpu_bias (base_conv_layer.hpp:105-128) and CuDNNConvolutionLayer::Forward_gpu
 * (cudnn_conv_layer.cu:13-58: cudnnConvolutionForward beta=0 + cudnnAddTensor).      */
int b2c_conv_forward(const b2c_conv_desc* d, const float* x, const float* w, const float* bias,
                     float* y, void* ws, size_t ws_bytes, void* stream);
/* dX = conv_bwd_data(dY, W).  dx OVERWRITTEN.
 * Replaces backward_gpu_gemm + col2im (base_conv_layer.hpp:130-146) /
 * cudnnConvolutionBackwardData beta=0 (cudnn_conv_layer.cu:118-123).                 */
int b2c_conv_backward_data(const b2c_conv_desc* d, const float* dy, const float* w, float* dx,
                           void* ws, size_t ws_bytes, void* stream);
/* dW += conv_bwd_filter(X, dY).  dw ACCUMULATED (beta = 1).
 * Replaces weight_gpu_gemm (base_conv_layer.hpp:148-162) /
 * cudnnConvolutionBackwardFilter beta=1 (cudnn_conv_layer.cu:95-99).                 */
/* dx += the bottom gradient (instead of dx = ...): where a blob feeds several layers the reference sums their bottom diffs
 * in SplitLayer::Backward (split_layer.cpp); here the second and later writers add straight into the blob's diff through
 * the kernel's TMA reduce-add store.  Available for the layers the staged kernel takes; ask first. */
int b2c_conv_backward_data_accumulate_supported(const b2c_conv_desc* d);
int b2c_conv_backward_data_accumulate(const b2c_conv_desc* d, const float* dy, const float* w, float* dx,
                                      void* ws, size_t ws_bytes, void* stream);
int b2c_conv_backward_filter(const b2c_conv_desc* d, const float* x, const float* dy, float* dw,
                             void* ws, size_t ws_bytes, void* stream);
/* db[o] += sum_{n,h,w} dY.  db ACCUMULATED (beta = 1).
 * Replaces backward_gpu_bias (base_conv_layer.hpp:164-168, gemv per image) /
 * cudnnConvolutionBackwardBias beta=1 (cudnn_conv_layer.cu:73-75).                   */
int b2c_conv_backward_bias(const b2c_conv_desc* d, const float* dy, float* db, void* stream);

/* ---- im2col / col2im ----------------------------------------------------------------
 * One image [C,H,W] <-> [C*kh*kw, Ho, Wo].  Replace im2col_gpu / col2im_gpu
 * (src/caffe/util/im2col.cu:42-62, 298-317).  col2im OVERWRITES im and adds the
 * contributions of each pixel in ascending (kh,kw) order (bit-identical to the
 * reference's col2im_cpu, im2col.cpp:176-211).                                         */
int b2c_im2col(const float* im, int C, int H, int W, int kh, int kw, int ph, int pw,
               int sh, int sw, int dh, int dw, float* col, void* stream);
int b2c_col2im(const float* col, int C, int H, int W, int kh, int kw, int ph, int pw,
               int sh, int sw, int dh, int dw, float* im, void* stream);
/* N-D (1..10 spatial axes): im2col_nd_gpu / col2im_nd_gpu (im2col.cu:165-253, 443-530).
 * Shape arrays are HOST int arrays: im_shape [C, d0..], col_shape [C*prod(k), o0..].   */
int b2c_im2col_nd(const float* im, int num_axes, const int* im_shape, const int* col_shape,
                  const int* kernel, const int* pad, const int* stride, const int* dilation,
                  float* col, void* stream);
int b2c_col2im_nd(const float* col, int num_axes, const int* im_shape, const int* col_shape,
                  const int* kernel, const int* pad, const int* stride, const int* dilation,
                  float* im, void* stream);

/* ---- BLAS-shaped entry points -----------------------------------------------------------
 * Row-major C[MxN] = alpha*op(A)*op(B) + beta*C with the reference's leading
 * dimensions (lda = transA ? M : K, ldb = transB ? K : N, ldc = N).
 * Replaces caffe_gpu_gemm<float> -> cublasSgemm (src/caffe/util/math_functions.cu:11-26). */
int b2c_sgemm(int transA, int transB, int M, int N, int K, float alpha, const float* A,
              const float* B, float beta, float* C, void* stream);
/* The same product with caller-provided scratch (b2c_sgemm_workspace_bytes(...) bytes, 16-byte aligned; may be null / short: then
 * as b2c_sgemm).  With it the NoTrans x Trans tensor-core path splits the K loop over one wave of CTAs and reduces the partial
 * tiles in a fixed order (deterministic) -- InnerProduct forward at small batch is otherwise one CTA per 128 x 128 output tile. */
/* 1 when b2c_sgemm / b2c_sgemm_ex would run this product (alpha = 1, beta in {0, 1}, 16-byte aligned operands) on the tensor cores:
 * NoTrans x Trans, K >= 64, K % 4 == 0, default math = fp32-equivalent.  Callers that can choose their operand layout
 * (InnerProductLayer::Backward_gpu with transposed copies) ask before preparing them. */
int b2c_sgemm_tc_supported(int transA, int transB, int M, int N, int K);
size_t b2c_sgemm_workspace_bytes(int transA, int transB, int M, int N, int K);
/* dst[c][r] = src[r][c] (rows x cols -> cols x rows), src != dst. */
int b2c_transpose(int rows, int cols, const float* src, float* dst, void* stream);
int b2c_sgemm_ex(int transA, int transB, int M, int N, int K, float alpha, const float* A, const float* B, float beta, float* C,
                 void* workspace, size_t workspace_bytes, void* stream);
/* y = alpha*op(A)*x + beta*y, A row-major MxN.
 * Replaces caffe_gpu_gemv<float> -> cublasSgemv (math_functions.cu:73-82).               */
int b2c_sgemv(int transA, int M, int N, float alpha, const float* A, const float* x,
              float beta, float* y, void* stream);

/* ---- the non-convolution layers on ResNet-50's training path ----------------------------
 * fp32 NCHW, asynchronous on `stream`.  Each keeps the semantics of the reference layer's CPU code.                */
/* ReLULayer::Forward/Backward (src/caffe/layers/relu_layer.cpp:10-41); may run in place (y == x).                  */
int b2c_relu_forward(size_t n, const float* x, float* y, float negative_slope, void* stream);
int b2c_relu_backward(size_t n, const float* dy, const float* x, float* dx, float negative_slope, void* stream);
/* BatchNormLayer (NVCaffe, src/caffe/layers/batch_norm_layer.cpp:140-300), TRAIN phase.  x,y,xnorm: [N,C,S].
 * gamma/beta null <=> scale_bias false.  running_var stores (variance + eps) like the reference's blobs_[1].
 * backward OVERWRITES dgamma/dbeta (also the reduction scratch when gamma is null) and dx.                          */
int b2c_bn_forward_train(int N, int C, int S, const float* x, const float* gamma, const float* beta, float eps,
                         float moving_average_fraction, int first_iteration, float* running_mean, float* running_var,
                         float* save_mean, float* save_invstd, float* xnorm, float* y, void* stream);
int b2c_bn_backward(int N, int C, int S, const float* dy, const float* xnorm, const float* gamma, const float* save_invstd,
                    float* dgamma, float* dbeta, float* dx, void* stream);
/* PoolingLayer (src/caffe/layers/pooling_layer.cpp:129-318): method 0 = MAX (mask = argmax index inside the H*W
 * plane, first maximum), 1 = AVE.  NC = N*C planes; output extent is the reference's ceil mode.  dx overwritten.    */
/* DataTransformer::Transform on a batch of uint8 datums (data_transformer.cpp:178-312): crop window and mirror flag per image
 * (device arrays of N), per-channel mean_values[C] or a per-pixel mean_image[C*Hd*Wd] in datum coordinates (or neither), scale.
 * out[n][c][h][w] = (datum[n][c][h_off[n]+h][w_off[n] + (mirror[n] ? crop_w-1-w : w)] - mean) * scale. */
int b2c_transform_u8(const unsigned char* src, int N, int C, int Hd, int Wd, int crop_h, int crop_w, const int* h_off,
                     const int* w_off, const unsigned char* mirror, const float* mean_values, const float* mean_image,
                     float scale, float* dst, void* stream);

/* Fused forms of the BatchNorm -> ReLU and Eltwise(SUM) -> ReLU chains (csrc/layers_fused.cu): bit-identical to the unfused
 * layer sequence, one HBM pass less each way; x_norm is recomputed in backward from the layer input and the saved statistics. */
int b2c_bn_forward_train_fused(int N, int C, int S, const float* x, const float* gamma, const float* beta, float eps,
                               float moving_average_fraction, int first_iteration, float* running_mean, float* running_var,
                               float* save_mean, float* save_invstd, float* y, int relu, void* stream);
int b2c_bn_backward_fused(int N, int C, int S, const float* dy, const float* x, const float* save_mean, const float* save_invstd,
                          const float* gamma, const float* beta, float* dgamma, float* dbeta, float* dx, int relu, void* stream);
/* The residual tail of a ResNet block -- BatchNorm -> Eltwise(SUM, 2 bottoms) -> in-place ReLU (batch_norm_layer.cpp,
 * eltwise_layer.cpp:47-60,100-140, relu_layer.cpp) -- as one launch each way.  Forward: y = [max(0, .)] (BatchNorm(x) + residual).
 * Backward: d_sum / y_sum are the diff and the post-ReLU data of the sum's top; dx = BatchNorm::Backward of d_sum * (y_sum > 0), and
 * that masked gradient is also written to d_residual (the sum's other bottom) unless it is null.  d_sum2 (may be null) is a second
 * part of the sum's top diff -- the shadow diff of a blob with two consumers -- added to d_sum on the fly (SplitLayer::Backward's
 * accumulation without its pass over memory).  Same bits as the layers run one by one. */
int b2c_bn_forward_train_fused_res(int N, int C, int S, const float* x, const float* gamma, const float* beta, float eps,
                                   float moving_average_fraction, int first_iteration, float* running_mean, float* running_var,
                                   float* save_mean, float* save_invstd, const float* residual, float* y, int relu, void* stream);
int b2c_bn_backward_fused_res(int N, int C, int S, const float* d_sum, const float* d_sum2, const float* y_sum, const float* x,
                              const float* save_mean, const float* save_invstd, const float* gamma, const float* beta,
                              float* dgamma, float* dbeta, float* dx, float* d_residual, void* stream);
/* AccuracyLayer::Forward (accuracy_layer.cpp:44-100), labels as float class ids, ties ranked like the reference's
 * std::greater<pair<score, index>>; `scratch`: 4 bytes of device memory. */
int b2c_accuracy(int N, int C, int top_k, const float* scores, const float* labels, float* accuracy, void* scratch, void* stream);
int b2c_add_relu(size_t n, const float* a, const float* b, float* y, void* stream);
int b2c_relu_backward2(size_t n, const float* dy, const float* y, float* dx_a, float* dx_b, void* stream);
int b2c_pool_forward(int method, int NC, int H, int W, int kh, int kw, int sh, int sw, int ph, int pw, const float* x,
                     float* y, int* mask, void* stream);
int b2c_pool_backward(int method, int NC, int H, int W, int kh, int kw, int sh, int sw, int ph, int pw, const float* dy,
                      const int* mask, float* dx, void* stream);
/* y = a + b: EltwiseLayer SUM forward (eltwise_layer.cpp) and the diff accumulation of SplitLayer::Backward.        */
int b2c_add(size_t n, const float* a, const float* b, float* y, void* stream);
/* SoftmaxWithLossLayer (src/caffe/layers/softmax_loss_layer.cpp:96-160), VALID normalisation, labels as float ids:
 * forward writes prob [N,C] and *loss (device scalar); backward writes dx = (prob - onehot) * loss_weight / N.       */
int b2c_softmax_loss_forward(int N, int C, const float* logits, const float* labels, float* prob, float* loss, void* stream);
int b2c_softmax_loss_backward(int N, int C, const float* prob, const float* labels, float loss_weight, float* dx, void* stream);
/* y[n][o][p] += bias[o];  db[o] += sum_{n,p} dy[n][o][p]  (InnerProduct bias with P = 1; conv bias with P = Ho*Wo). */
int b2c_bias_forward(int N, int O, int P, const float* bias, float* y, void* stream);
int b2c_bias_backward(int N, int O, int P, const float* dy, float* db, void* stream);
/* ---- layers the AlexNet / GoogLeNet / VGG-16 BASELINE nets add (second half);
 * GPU parity: tests/test_layers_extra_gpu.py.
 * LRNLayer ACROSS_CHANNELS (src/caffe/layers/lrn_layer.cpp CrossChannelForward_cpu / CrossChannelBackward_cpu):
 * scale = k + alpha/size * sum_{window} x^2, y = x * scale^-beta; x,y,scale,dy,dx: [N,C,S].                          */
int b2c_lrn_forward(int N, int C, int S, int local_size, float alpha, float beta, float k, const float* x, float* scale,
                    float* y, void* stream);
int b2c_lrn_backward(int N, int C, int S, int local_size, float alpha, float beta, const float* x, const float* y,
                     const float* scale, const float* dy, float* dx, void* stream);
/* DropoutLayer TRAIN (src/caffe/layers/dropout_layer.cpp): mask[i] = keep ? 1/(1-ratio) : 0 with
 * keep = (splitmix64(seed + offset + i) >> 40) >= ratio * 2^24  (counter based: reproducible, no state);
 * forward y = x * mask and backward dx = dy * mask are b2c_mul.                                                        */
int b2c_dropout_mask(size_t n, float ratio, unsigned long long seed, unsigned long long offset, float* mask, void* stream);
int b2c_mul(size_t n, const float* a, const float* b, float* y, void* stream);

/* ---- fused SGD-momentum update -----------------------------------------------------------
 * h = momentum*h + local_rate*(grad_scale*g + local_decay*reg(w)); w -= h;
 * g = clear_grads ? 0 : h.   reg(w) = w (l2 != 0) or sign(w).
 * Replaces SGDRegUpdateAllAndClear + its launcher (src/caffe/solvers/sgd_solver.cu:9-72),
 * with Net::ReduceBucket's 1/solver_count scal (net.cpp:910), the 1/global_grad_scale
 * scal (net.cpp:815-817) and SGDSolver::Normalize's 1/iter_size (sgd_solver.cpp:152-158)
 * folded into grad_scale.                                                                 */
int b2c_sgd_update(size_t n, float* g, float* w, float* h, float momentum, float local_rate,
                   float local_decay, int l2, float grad_scale, int clear_grads, void* stream);
/* Multi-tensor form over a contiguous arena (Net::InitializeLearnableDiffSpace,
 * net.cpp:1350-1373): nseg segments, segment s covers elements [offset[s], offset[s]+count[s])
 * of g, w and h (same offsets in all three arenas) with its own local_rate / local_decay
 * (lr_mult / decay_mult, sgd_solver.cpp:208-210,254-259).  The four arrays are HOST arrays.
 * One launch replaces one SGDRegUpdateAllAndClear launch per learnable blob.               */
int b2c_sgd_update_arena(int nseg, const size_t* offset, const size_t* count,
                         const float* local_rate, const float* local_decay,
                         float* g, float* w, float* h, float momentum, int l2,
                         float grad_scale, int clear_grads, void* stream);

/* ---- gradient exchange ---------------------------------------------------------------------
 * Replaces P2PManager/P2PSync's NCCL use (src/caffe/parallel.cpp:36-87,145-253) and the MPI
 * bootstrap (src/caffe/clusters.cpp:8-16; parallel.cpp:42-45,163-172).  One communicator per
 * process (one process per GPU).  The 128-byte id is produced on rank 0 and carried to the
 * other ranks by the caller (MPI_Bcast in the reference; any byte transport here).           */
/* Debugging aid for the mbarrier-synchronised kernels: every wait is bounded (~1 s); a timeout records
 * {block, thread, barrier, parity} and kills the kernel with a trap.  b2c_debug_mbar_set_trap(0) makes timeouts non-fatal so
 * that a deadlocked kernel drains; b2c_debug_mbar_timeouts fills four 128-word blocks (gather fwd/dgrad, staged fwd/dgrad,
 * gather weight-gradient, staged weight-gradient kernels), each {count, -, -, -, records...}, clears them and returns the
 * total count (cap_words >= 512). */
int b2c_debug_mbar_timeouts(unsigned int* out, int cap_words);
int b2c_debug_mbar_set_trap(int on);

#define B2C_UNIQUE_ID_BYTES 128
int b2c_comm_get_unique_id(void* id_out /* B2C_UNIQUE_ID_BYTES */);
int b2c_comm_init(int nranks, int rank, const void* id, b2c_comm** out);
int b2c_comm_destroy(b2c_comm* c);
int b2c_comm_nranks(const b2c_comm* c);
/* ncclBcast of parameter data from `root` (P2PSync::on_start, parallel.cpp:208-227).        */
int b2c_comm_bcast(b2c_comm* c, float* buf, size_t count, int root, void* stream);
/* in-place ncclSum allreduce of a diff bucket (P2PSync::allreduce_bucket, parallel.cpp:245-253) */
int b2c_comm_allreduce_sum(b2c_comm* c, float* buf, size_t count, void* stream);
/* NVLS-capable buffers: ncclMemAlloc / ncclMemFree and ncclCommRegister (handles are released by b2c_comm_destroy).  An
 * in-place allreduce on a registered ncclMemAlloc buffer can be reduced inside the NVSwitch without NCCL's staging copies.
 * Environment: B2C_NCCL_MAX_CTAS / B2C_NCCL_MIN_CTAS set ncclConfig_t.maxCTAs / minCTAs of b2c_comm_init. */
int b2c_comm_mem_alloc(void** ptr, size_t bytes);
int b2c_comm_mem_free(void* ptr);
int b2c_comm_register(b2c_comm* c, void* buf, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif  /* B2C_H_ */
