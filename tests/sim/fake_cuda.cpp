// fake_cuda.cpp -- a CUDA-runtime stand-in that models STREAM ORDER, for testing host code that orchestrates streams and events
// on a machine without a GPU (tests/test_data_layer_sim.py).  TEST INFRASTRUCTURE ONLY: it is linked into tests/sim/libdatasim.so
// and nowhere else; the product libraries link the real runtime.
//
// Model.  "Device" memory is host memory.  Every stream is a FIFO of pending operations; nothing runs when it is enqueued.  An
// operation runs only when (a) its stream is EAGER and everything before it has run and its event waits are satisfied, or (b) a
// host-side synchronisation (cudaStreamSynchronize, cudaEventSynchronize, cudaDeviceSynchronize, a blocking cudaMemcpy, cudaFree)
// forces it -- recursively forcing whatever it waits for.  A LAZY stream therefore runs as LATE as the program's dependencies
// allow and an EAGER one as EARLY as they allow; running a host program under both extremes (per stream) exposes missing
// dependencies in either direction: a consumer that forgot to wait for its producer reads stale data under lazy producers, a
// producer that forgot to wait for the previous consumer overwrites live data under eager producers.  cudaMemcpyAsync reads its
// source WHEN IT RUNS, which is the real semantics for pinned host memory (and the reason a pinned buffer must not be refilled
// before its copy has completed).  Single host thread only (the caller's); other threads may touch plain memory but not this API.
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <vector>

namespace {

struct Stream;
struct Event {
  unsigned long long gen = 0;                                   // generation of the latest cudaEventRecord
  unsigned long long done = 0;                                  // latest generation that has executed
  std::map<unsigned long long, std::pair<Stream*, size_t>> at;  // generation -> (stream, absolute op index of its record op)
};
// A collective over several streams (tests/sim/fake_comm.cpp: the allreduce of N in-process "ranks"): one COLLECTIVE op per rank,
// on that rank's stream.  It can run only when every rank's op has reached the FRONT of its stream -- everything each rank enqueued
// before it has executed -- and then runs once, for all ranks; the other ranks' ops complete as no-ops.
struct Collective {
  std::vector<char> enqueued;
  std::vector<std::pair<Stream*, size_t>> where;                // per rank: stream and absolute index of its op
  std::function<void()> run;
  bool done = false;
};
struct Op {
  enum Kind { EXEC, RECORD, WAIT, COLLECTIVE } kind = EXEC;
  std::function<void()> fn;
  Event* ev = nullptr;
  unsigned long long gen = 0;
  std::shared_ptr<Collective> coll;
};
struct Stream {
  std::deque<Op> q;
  size_t base = 0;          // absolute index of q.front()
  bool eager = false;
};

Stream g_legacy;                            // the legacy default stream (cudaStream_t 0)
std::vector<Stream*> g_streams{&g_legacy};
unsigned long long g_executed = 0;

Stream* S(cudaStream_t s) { return s ? reinterpret_cast<Stream*>(s) : &g_legacy; }
Event* E(cudaEvent_t e) { return reinterpret_cast<Event*>(e); }

void force(Stream* s, size_t upto_abs);     // run everything on s with absolute index < upto_abs

bool collective_ready(const Collective& c);
bool runnable(const Op& op) {
  if (op.kind == Op::WAIT) return op.ev->done >= op.gen;
  if (op.kind == Op::COLLECTIVE) return op.coll->done || collective_ready(*op.coll);
  return true;
}

void run_front(Stream* s) {
  Op op = std::move(s->q.front());
  s->q.pop_front();
  ++s->base;
  ++g_executed;
  if (op.kind == Op::EXEC) op.fn();
  else if (op.kind == Op::RECORD) { if (op.ev->done < op.gen) op.ev->done = op.gen; }
  else if (op.kind == Op::COLLECTIVE && !op.coll->done) { op.coll->done = true; op.coll->run(); }
}
bool collective_ready(const Collective& c) {
  for (size_t r = 0; r < c.where.size(); ++r)
    if (!c.enqueued[r] || c.where[r].first->base != c.where[r].second) return false;
  return true;
}

void progress() {                            // eager streams run as far as their dependencies allow
  for (bool moved = true; moved;) {
    moved = false;
    for (Stream* s : g_streams)
      while (s->eager && !s->q.empty() && runnable(s->q.front())) { run_front(s); moved = true; }
  }
}

void force_event(Event* e, unsigned long long gen) {
  if (e->done >= gen) return;
  auto it = e->at.find(gen);
  if (it == e->at.end()) throw std::runtime_error("fake_cuda: wait on an event generation that was never recorded");
  force(it->second.first, it->second.second + 1);
}

void force(Stream* s, size_t upto_abs) {
  while (s->base < upto_abs && !s->q.empty()) {
    Op& op = s->q.front();
    if (op.kind == Op::WAIT && op.ev->done < op.gen) { force_event(op.ev, op.gen); continue; }   // re-read the queue: it may have moved
    if (op.kind == Op::COLLECTIVE && !op.coll->done && !collective_ready(*op.coll)) {
      // bring every other rank up to its own op of this collective; a rank that has not even enqueued it means the host program
      // synchronised on one rank before launching the others -- a deadlock on real hardware too
      std::shared_ptr<Collective> c = op.coll;
      for (size_t r = 0; r < c->where.size(); ++r) {
        if (!c->enqueued[r]) throw std::runtime_error("fake_cuda: collective waits for a rank that has not enqueued it (host-side deadlock)");
        if (c->where[r].first != s) force(c->where[r].first, c->where[r].second);
      }
      continue;
    }
    run_front(s);
    progress();
  }
}

void push(Stream* s, Op op) {
  s->q.push_back(std::move(op));
  progress();
}
void push_exec(cudaStream_t st, std::function<void()> fn) {
  Op op; op.kind = Op::EXEC; op.fn = std::move(fn);
  push(S(st), std::move(op));
}
void force_all() {
  for (Stream* s : g_streams) force(s, s->base + s->q.size());
}

}  // namespace

// ---- simulator controls (tests/sim/sim_capi.cpp) ---------------------------------------------------------------------------------
extern "C" void fakecuda_set_eager(cudaStream_t st, int eager) { S(st)->eager = eager != 0; progress(); }
extern "C" void fakecuda_set_all_eager(int eager) { for (Stream* s : g_streams) s->eager = eager != 0; progress(); }
extern "C" unsigned long long fakecuda_executed() { return g_executed; }
extern "C" unsigned long long fakecuda_pending() { unsigned long long n = 0; for (Stream* s : g_streams) n += s->q.size(); return n; }
// enqueue an arbitrary "kernel" (used by the fake b2c_transform_u8)
void fakecuda_launch(cudaStream_t st, std::function<void()> fn) { push_exec(st, std::move(fn)); }
// one rank's part of a collective over `nranks` streams; `run` (taken from the first caller) executes it once for all ranks
std::shared_ptr<void> fakecuda_collective(cudaStream_t st, std::shared_ptr<void> handle, int nranks, int rank, std::function<void()> run) {
  std::shared_ptr<Collective> c = handle ? std::static_pointer_cast<Collective>(handle) : std::make_shared<Collective>();
  if (c->where.empty()) { c->enqueued.assign(nranks, 0); c->where.assign(nranks, {nullptr, 0}); c->run = std::move(run); }
  Stream* s = S(st);
  c->enqueued[rank] = 1;
  c->where[rank] = {s, s->base + s->q.size()};
  Op op; op.kind = Op::COLLECTIVE; op.coll = c;
  push(s, std::move(op));
  return c;
}

// ---- the runtime entry points the host layer uses -----------------------------------------------------------------------------------
extern "C" {

const char* cudaGetErrorString(cudaError_t) { return "fake_cuda error"; }
static cudaError_t aligned_zeroed(void** p, size_t n) {        // cudaMalloc / cudaMallocHost return at least 256-byte aligned memory
  if (posix_memalign(p, 256, n ? n : 1) != 0) { *p = nullptr; return cudaErrorMemoryAllocation; }
  std::memset(*p, 0, n ? n : 1);
  return cudaSuccess;
}
cudaError_t cudaMalloc(void** p, size_t n) { return aligned_zeroed(p, n); }
cudaError_t cudaMallocHost(void** p, size_t n) { return aligned_zeroed(p, n); }
cudaError_t cudaFree(void* p) { force_all(); std::free(p); return cudaSuccess; }            // cudaFree synchronises the device
cudaError_t cudaFreeHost(void* p) { force_all(); std::free(p); return cudaSuccess; }

cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) {
  Stream* st = new Stream();
  g_streams.push_back(st);
  *s = reinterpret_cast<cudaStream_t>(st);
  return cudaSuccess;
}
cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned f, int) { return cudaStreamCreateWithFlags(s, f); }
cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { *lo = 0; *hi = -1; return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) {
  Stream* st = S(s);
  force(st, st->base + st->q.size());
  for (size_t i = 0; i < g_streams.size(); ++i) if (g_streams[i] == st) { g_streams.erase(g_streams.begin() + i); break; }
  delete st;
  return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t s) { Stream* st = S(s); force(st, st->base + st->q.size()); return cudaSuccess; }
cudaError_t cudaDeviceSynchronize() { force_all(); return cudaSuccess; }

cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = reinterpret_cast<cudaEvent_t>(new Event()); return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { return cudaEventCreateWithFlags(e, 0); }
cudaError_t cudaEventDestroy(cudaEvent_t e) { force_all(); delete E(e); return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s) {
  Event* ev = E(e);
  Stream* st = S(s);
  ++ev->gen;
  ev->at[ev->gen] = {st, st->base + st->q.size()};
  Op op; op.kind = Op::RECORD; op.ev = ev; op.gen = ev->gen;
  push(st, std::move(op));
  return cudaSuccess;
}
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned) {
  Event* ev = E(e);
  if (ev->gen == 0) return cudaSuccess;              // never recorded: the wait is a no-op, as in CUDA
  Op op; op.kind = Op::WAIT; op.ev = ev; op.gen = ev->gen;   // the wait captures the record that precedes it in program order
  push(S(s), std::move(op));
  return cudaSuccess;
}
cudaError_t cudaEventSynchronize(cudaEvent_t e) { Event* ev = E(e); if (ev->gen) force_event(ev, ev->gen); return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }

cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t n, cudaMemcpyKind, cudaStream_t s) {
  push_exec(s, [=] { std::memmove(dst, src, n); });  // the source is read when the copy RUNS
  return cudaSuccess;
}
cudaError_t cudaMemcpy2DAsync(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, cudaMemcpyKind,
                              cudaStream_t s) {
  push_exec(s, [=] {
    for (size_t r = 0; r < height; ++r) std::memmove(static_cast<char*>(dst) + r * dpitch, static_cast<const char*>(src) + r * spitch, width);
  });
  return cudaSuccess;
}
cudaError_t cudaMemsetAsync(void* dst, int v, size_t n, cudaStream_t s) {
  push_exec(s, [=] { std::memset(dst, v, n); });
  return cudaSuccess;
}
cudaError_t cudaMemcpy(void* dst, const void* src, size_t n, cudaMemcpyKind k) {              // blocking, legacy stream
  cudaMemcpyAsync(dst, src, n, k, nullptr);
  return cudaStreamSynchronize(nullptr);
}
cudaError_t cudaMemset(void* dst, int v, size_t n) { cudaMemsetAsync(dst, v, n, nullptr); return cudaStreamSynchronize(nullptr); }

}  // extern "C"
