// fake_comm.cpp -- b2c_comm_* (the NCCL communicator of csrc/comm.cu) for N "ranks" living in ONE process on the stream-order model:
// every rank is a TrainNet + P2PSync + ReduceScheduler with its own compute stream, its allreduce a COLLECTIVE op of
// tests/sim/fake_cuda.cpp that completes only when every rank's comm stream has reached it.  This is what lets the reference's
// multi-device check -- N solvers on batch/N each == one solver on the batch (test_gradient_based_solver.cpp:471-509) -- run through
// the product's own C++ exchange code on a machine without GPUs (tests/test_trainer_sim.py).  TEST INFRASTRUCTURE ONLY.
#include <cuda_runtime.h>

#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/b2c.h"

void fakecuda_launch(cudaStream_t st, std::function<void()> fn);
std::shared_ptr<void> fakecuda_collective(cudaStream_t st, std::shared_ptr<void> handle, int nranks, int rank, std::function<void()> run);

namespace {
struct Call { std::shared_ptr<void> handle; std::vector<float*> buf; size_t count = 0; };
struct Group {
  int nranks = 0;
  std::map<unsigned long long, std::shared_ptr<Call>> calls;     // allreduce number -> its per-rank buffers
  const float* root_buf = nullptr;                               // broadcast source, once the root has run its part
};
std::map<std::string, std::shared_ptr<Group>> g_groups;
unsigned long long g_next_id = 1;
}  // namespace

struct b2c_comm {
  std::shared_ptr<Group> g;
  int nranks = 1, rank = 0;
  unsigned long long seq = 0;
};

extern "C" {

int b2c_comm_get_unique_id(void* id_out) {
  std::memset(id_out, 0, B2C_UNIQUE_ID_BYTES);
  const unsigned long long id = g_next_id++;
  std::memcpy(id_out, &id, sizeof(id));
  return B2C_OK;
}
int b2c_comm_init(int nranks, int rank, const void* id, b2c_comm** out) {
  if (nranks < 1 || rank < 0 || rank >= nranks || !id || !out) return B2C_ERR_INVALID;
  const std::string key(static_cast<const char*>(id), B2C_UNIQUE_ID_BYTES);
  std::shared_ptr<Group>& g = g_groups[key];
  if (!g) { g = std::make_shared<Group>(); g->nranks = nranks; }
  if (g->nranks != nranks) return B2C_ERR_INVALID;
  b2c_comm* c = new b2c_comm;
  c->g = g; c->nranks = nranks; c->rank = rank;
  *out = c;
  return B2C_OK;
}
int b2c_comm_destroy(b2c_comm* c) { delete c; return B2C_OK; }
int b2c_comm_nranks(const b2c_comm* c) { return c->nranks; }
// Rank 0's weights to everyone (P2PSync::on_start).  The callers synchronise right after it, one rank at a time, so this is not a
// rendezvous here: the root publishes its buffer, the others copy from it (the root attaches first, as rank 0 does in the launchers).
int b2c_comm_bcast(b2c_comm* c, float* buf, size_t count, int root, void* stream) {
  std::shared_ptr<Group> g = c->g;
  const bool is_root = c->rank == root;
  fakecuda_launch(static_cast<cudaStream_t>(stream), [=] {
    if (is_root) g->root_buf = buf;
    else if (g->root_buf) std::memcpy(buf, g->root_buf, sizeof(float) * count);
  });
  return B2C_OK;
}
// In-place sum over the ranks, in rank order (so every rank ends with the same bits, like NCCL's deterministic ring / tree for a fixed
// topology), delivered to all of them when the last rank's comm stream arrives.
int b2c_comm_allreduce_sum(b2c_comm* c, float* buf, size_t count, void* stream) {
  std::shared_ptr<Call>& call = c->g->calls[c->seq++];
  if (!call) { call = std::make_shared<Call>(); call->buf.assign(c->nranks, nullptr); call->count = count; }
  if (call->count != count) return B2C_ERR_INVALID;
  call->buf[c->rank] = buf;
  std::shared_ptr<Call> k = call;
  call->handle = fakecuda_collective(static_cast<cudaStream_t>(stream), call->handle, c->nranks, c->rank, [k] {
    std::vector<float> sum(k->buf[0], k->buf[0] + k->count);
    for (size_t r = 1; r < k->buf.size(); ++r) for (size_t i = 0; i < k->count; ++i) sum[i] += k->buf[r][i];
    for (float* b : k->buf) std::memcpy(b, sum.data(), sizeof(float) * k->count);
  });
  return B2C_OK;
}
int b2c_comm_mem_alloc(void** ptr, size_t bytes) { return cudaMalloc(ptr, bytes) == cudaSuccess ? B2C_OK : B2C_ERR_CUDA; }
int b2c_comm_mem_free(void* ptr) { cudaFree(ptr); return B2C_OK; }
int b2c_comm_register(b2c_comm*, void*, size_t) { return B2C_OK; }

}  // extern "C"
