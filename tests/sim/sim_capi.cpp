// sim_capi.cpp -- drives caffe::DataLayer (the product class, compiled from caffe_mpi_b200/host/data_layer.cpp unchanged) on top of
// tests/sim/fake_cuda.cpp's stream-order model, so that its slot protocol -- who waits for whom between the parser threads, the
// copy stream and the compute stream -- can be tested without a GPU.  TEST INFRASTRUCTURE ONLY.
//
// The one device kernel the layer launches, b2c_transform_u8, is replaced here by a host closure enqueued on the same stream with
// the same arithmetic (out = (datum[h_off + h][w_off + (mirror ? W-1-w : w)] - mean) * scale, include/b2c.h); the real kernel is
// checked bit-exactly on hardware by tests/test_layers_extra_gpu.py::test_transform_u8_is_bit_exact.
#include <cuda_runtime.h>

#include <cstring>
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "../../caffe_mpi_b200/host/data_layer.hpp"

void fakecuda_launch(cudaStream_t st, std::function<void()> fn);

// what every executed transform actually READ (sum of the source bytes of its batch), in execution order: lets a test see batches
// it never read back, e.g. one whose device slot was overwritten by a later copy before the transform ran
static std::vector<unsigned long long> g_transform_log;

extern "C" int b2c_transform_u8(const unsigned char* src, int N, int C, int Hd, int Wd, int crop_h, int crop_w, const int* h_off,
                                const int* w_off, const unsigned char* mirror, const float* mean_values, const float* mean_image,
                                float scale, float* dst, void* stream) {
  fakecuda_launch(static_cast<cudaStream_t>(stream), [=] {
    unsigned long long sum = 0;
    for (size_t i = 0; i < (size_t)N * C * Hd * Wd; ++i) sum += src[i];
    g_transform_log.push_back(sum);
    for (int n = 0; n < N; ++n)
      for (int c = 0; c < C; ++c)
        for (int h = 0; h < crop_h; ++h)
          for (int w = 0; w < crop_w; ++w) {
            const int ws = w_off[n] + (mirror[n] ? crop_w - 1 - w : w);
            const size_t di = ((size_t)c * Hd + h_off[n] + h) * Wd + ws;
            const float x = (float)src[(size_t)n * C * Hd * Wd + di];
            const float m = mean_image ? mean_image[di] : mean_values ? mean_values[c] : 0.f;
            dst[(((size_t)n * C + c) * crop_h + h) * crop_w + w] = (x - m) * scale;
          }
  });
  return 0;
}

using namespace caffe;

static thread_local std::string g_serr;
struct SimHandle {
  std::unique_ptr<Net> net;
  std::unique_ptr<DataLayer> layer;
  Blob data, label;
  vector<Blob*> tops;
};

extern "C" {

const char* sim_last_error() { return g_serr.c_str(); }

void* sim_create(const char* net_text, unsigned long long seed, int solver_count, int solver_rank) {
  try {
    std::unique_ptr<SimHandle> h(new SimHandle);
    h->net.reset(new Net(ParseTextProto(net_text), TRAIN));
    const NetLayer& L = h->net->layers().at(0);
    B2_CHECK(L.param.type == "Data" && L.use_database, "sim_create: layer 0 must be a Data layer whose source opens");
    h->layer.reset(new DataLayer(L, seed));
    h->tops = {&h->data, &h->label};
    h->layer->SetUp({}, h->tops);
    h->layer->set_solver(solver_count, solver_rank);
    return h.release();
  } catch (const std::exception& e) { g_serr = e.what(); return nullptr; }
}
void sim_destroy(void* hv) { delete static_cast<SimHandle*>(hv); }
int sim_shape(void* hv, int* nchw) {
  auto* h = static_cast<SimHandle*>(hv);
  for (int i = 0; i < 4; ++i) nchw[i] = h->data.shape(i);
  return 0;
}
int sim_load_batch(void* hv) {
  auto* h = static_cast<SimHandle*>(hv);
  try { h->layer->LoadBatch(h->tops, nullptr); return 0; } catch (const std::exception& e) { g_serr = e.what(); return -1; }
}
// Blob::cpu_data(): an asynchronous device -> host copy on the compute stream followed by a stream synchronise
int sim_read(void* hv, float* data, float* label) {
  auto* h = static_cast<SimHandle*>(hv);
  try {
    std::memcpy(data, h->data.cpu_data(), sizeof(float) * h->data.count());
    std::memcpy(label, h->label.cpu_data(), sizeof(float) * h->label.count());
    return 0;
  } catch (const std::exception& e) { g_serr = e.what(); return -1; }
}
int sim_transform_log(unsigned long long* out, int cap, int clear) {
  const int n = (int)g_transform_log.size();
  for (int i = 0; i < n && i < cap; ++i) out[i] = g_transform_log[i];
  if (clear) g_transform_log.clear();
  return n;
}
// Several "ranks" in one process: each gets its own compute stream (Caffe::thread_stream() is what every layer launches on); call
// before creating / stepping that rank's trainer.  rank < 0 goes back to the legacy stream.
static std::vector<cudaStream_t>& rank_streams() { static std::vector<cudaStream_t> v; return v; }
int sim_use_rank_stream(int rank) {
  std::vector<cudaStream_t>& streams = rank_streams();
  if (rank < 0) { Caffe::set_thread_stream(nullptr); return 0; }
  while ((int)streams.size() <= rank) { cudaStream_t s = nullptr; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking); streams.push_back(s); }
  Caffe::set_thread_stream(streams[rank]);
  return 0;
}
extern "C" void fakecuda_set_eager(cudaStream_t st, int eager);
void sim_set_rank_stream_eager(int rank, int eager) {
  if (rank >= 0 && rank < (int)rank_streams().size()) fakecuda_set_eager(rank_streams()[rank], eager);
}
// P2PSync's constructor sets the process-wide solver count (one process is one rank in the product); a test that afterwards runs a
// single solver in the same process puts it back.
void sim_set_solver_count(int n) { Caffe::set_solver_count(n); Caffe::set_root_solver(true); }
long long sim_batches(void* hv) { return (long long)static_cast<SimHandle*>(hv)->layer->batches_loaded(); }

}  // extern "C"
