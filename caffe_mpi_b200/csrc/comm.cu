// comm.cu -- gradient exchange over NCCL (NVLink 5 / NVSwitch), one communicator per process.
//
// Replaces the NCCL use of P2PManager / P2PSync (reference src/caffe/parallel.cpp:36-87,
// 145-253: ncclGetUniqueId + MPI_Bcast, ncclCommInitRank between MPI_Barriers, ncclBcast of the
// initial weights, in-place ncclSum ncclAllReduce of each diff bucket followed by a host
// cudaStreamSynchronize) and the MPI bootstrap of src/caffe/clusters.cpp:8-16.  Differences by
// design: one process per GPU (no solver threads, no boost barriers, no GPUMemory write-lock),
// every call is asynchronous on the caller's stream -- ordering against the backward pass and the
// SGD kernel is done with CUDA events by the caller (caffe_mpi_b200/host), never by host syncs.
#include <nccl.h>
#include "b2c_common.cuh"

struct b2c_comm {
  ncclComm_t comm;
  int nranks, rank;
};

#define B2C_NCCL_OK(expr)                                                                          \
  do {                                                                                             \
    ncclResult_t r__ = (expr);                                                                     \
    if (r__ != ncclSuccess)                                                                        \
      return ::b2c::fail(B2C_ERR_NCCL, "%s:%d %s: %s", __FILE__, __LINE__, #expr, ncclGetErrorString(r__)); \
  } while (0)

static_assert(sizeof(ncclUniqueId) == B2C_UNIQUE_ID_BYTES, "ncclUniqueId size");

extern "C" int b2c_comm_get_unique_id(void* id_out) {
  if (!id_out) return b2c::fail(B2C_ERR_INVALID, "b2c_comm_get_unique_id: null");
  ncclUniqueId id;
  B2C_NCCL_OK(ncclGetUniqueId(&id));
  memcpy(id_out, &id, sizeof(id));
  return B2C_OK;
}

extern "C" int b2c_comm_init(int nranks, int rank, const void* id, b2c_comm** out) {
  if (!id || !out || nranks < 1 || rank < 0 || rank >= nranks)
    return b2c::fail(B2C_ERR_INVALID, "b2c_comm_init: bad argument");
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof(uid));
  b2c_comm* c = new b2c_comm{nullptr, nranks, rank};
  ncclResult_t r = ncclCommInitRank(&c->comm, nranks, uid, rank);
  if (r != ncclSuccess) {
    delete c;
    return b2c::fail(B2C_ERR_NCCL, "ncclCommInitRank: %s", ncclGetErrorString(r));
  }
  *out = c;
  return B2C_OK;
}

extern "C" int b2c_comm_destroy(b2c_comm* c) {
  if (!c) return B2C_OK;
  ncclCommDestroy(c->comm);
  delete c;
  return B2C_OK;
}

extern "C" int b2c_comm_nranks(const b2c_comm* c) { return c ? c->nranks : 0; }

extern "C" int b2c_comm_bcast(b2c_comm* c, float* buf, size_t count, int root, void* stream) {
  if (!c || !buf) return b2c::fail(B2C_ERR_INVALID, "b2c_comm_bcast: null");
  B2C_NCCL_OK(ncclBroadcast(buf, buf, count, ncclFloat, root, c->comm, b2c::as_stream(stream)));
  return B2C_OK;
}

extern "C" int b2c_comm_allreduce_sum(b2c_comm* c, float* buf, size_t count, void* stream) {
  if (!c || !buf) return b2c::fail(B2C_ERR_INVALID, "b2c_comm_allreduce_sum: null");
  B2C_NCCL_OK(ncclAllReduce(buf, buf, count, ncclFloat, ncclSum, c->comm, b2c::as_stream(stream)));
  return B2C_OK;
}
