// b2h_data_capi.cpp -- flat C entry points over the input pipeline's host half (lmdb_reader / proto_wire Datum / DataReader /
// TransformDraws), used by tests/test_data_cpu.py through ctypes.  None of these touch the device.  Same conventions as
// b2h_capi.cpp: 0 / a handle on success, -1 / NULL with the message in b2h_data_last_error().
#include <cstdio>
#include <cstring>
#include <memory>

#include "../../include/b2h_data.h"
#include "b2caffe.hpp"
#include "data_reader.hpp"
#include "jpeg_decode.hpp"
#include "lmdb_reader.hpp"
#include "proto_wire.hpp"

using namespace caffe;

static thread_local std::string g_derr;
#define B2D_TRY(body)                          \
  try { body; return 0; }                      \
  catch (const std::exception& e) { g_derr = e.what(); return -1; }

struct LmdbHandle {
  db::LMDB env;
  std::unique_ptr<db::LMDBCursor> cur;
  std::unique_ptr<db::LMDBTransaction> txn;
};
struct ReaderHandle {
  std::unique_ptr<DataReader> reader;
  // a small pool of pageable batches circulating through the reader (the data layer circulates pinned ones)
  struct Owned { std::vector<uint8_t> data; std::vector<float> label; std::vector<uint32_t> ids; BatchBuf buf; };
  std::vector<std::unique_ptr<Owned>> pool;
};

extern "C" {

const char* b2h_data_last_error() { return g_derr.c_str(); }

// ---- db::LMDB / LMDBCursor ------------------------------------------------------------------------------------------------
int b2h_lmdb_exists(const char* source) { return db::LMDB::Exists(source) ? 1 : 0; }
void* b2h_lmdb_open(const char* source) {
  try {
    std::unique_ptr<LmdbHandle> h(new LmdbHandle);
    h->env.Open(source, db::READ);
    h->cur.reset(h->env.NewCursor());
    return h.release();
  } catch (const std::exception& e) { g_derr = e.what(); return nullptr; }
}
// mode: 0 READ, 1 WRITE, 2 NEW (db::Mode)
void* b2h_lmdb_open_mode(const char* source, int mode) {
  try {
    std::unique_ptr<LmdbHandle> h(new LmdbHandle);
    h->env.Open(source, mode == 2 ? db::NEW : mode == 1 ? db::WRITE : db::READ);
    h->cur.reset(h->env.NewCursor());
    return h.release();
  } catch (const std::exception& e) { g_derr = e.what(); return nullptr; }
}
// Transaction::Put on the handle's pending transaction (db_lmdb.cpp:52-55)
int b2h_lmdb_put(void* hv, const void* key, size_t key_size, const void* value, size_t value_size) {
  auto* h = static_cast<LmdbHandle*>(hv);
  B2D_TRY({
    if (!h->txn) h->txn.reset(h->env.NewTransaction());
    h->txn->Put(std::string(static_cast<const char*>(key), key_size), std::string(static_cast<const char*>(value), value_size));
  });
}
// Transaction::Commit (db_lmdb.cpp:57-96); the handle's cursor is re-created on the committed state, at the first record
int b2h_lmdb_commit(void* hv) {
  auto* h = static_cast<LmdbHandle*>(hv);
  B2D_TRY({
    if (!h->txn) h->txn.reset(h->env.NewTransaction());
    h->cur.reset();
    h->txn->Commit();
    h->txn.reset();
    h->cur.reset(h->env.NewCursor());
  });
}
void b2h_lmdb_close(void* hv) { delete static_cast<LmdbHandle*>(hv); }
int b2h_lmdb_stat(void* hv, long long* entries, unsigned* page_size, unsigned* depth, unsigned long long* txnid) {
  auto* h = static_cast<LmdbHandle*>(hv);
  B2D_TRY({ *entries = (long long)h->env.entries(); *page_size = h->env.page_size(); *depth = h->env.depth(); *txnid = h->env.txnid(); });
}
int b2h_lmdb_seek_to_first(void* hv) {
  auto* h = static_cast<LmdbHandle*>(hv);
  try { h->cur->SeekToFirst(); return h->cur->valid() ? 1 : 0; } catch (const std::exception& e) { g_derr = e.what(); return -1; }
}
int b2h_lmdb_next(void* hv) {
  auto* h = static_cast<LmdbHandle*>(hv);
  try { h->cur->Next(); return h->cur->valid() ? 1 : 0; } catch (const std::exception& e) { g_derr = e.what(); return -1; }
}
int b2h_lmdb_valid(void* hv) { return static_cast<LmdbHandle*>(hv)->cur->valid() ? 1 : 0; }
// pointers into the mapping, valid until b2h_lmdb_close
int b2h_lmdb_current(void* hv, const void** key, size_t* key_size, const void** value, size_t* value_size) {
  auto* h = static_cast<LmdbHandle*>(hv);
  B2D_TRY({
    B2_CHECK(h->cur->valid(), "cursor is not positioned on a record");
    *key = h->cur->key_data(); *key_size = h->cur->key_size(); *value = h->cur->data(); *value_size = h->cur->size();
  });
}

// ---- Datum ------------------------------------------------------------------------------------------------------------------
// out[0..5] = channels, height, width, label, encoded, record_id; data / data_size view the input; float_data is copied when
// float_cap > 0.  Returns 1 if the bytes parse, 0 if not (Datum::ParseFromArray's bool).
int b2h_datum_parse(const void* bytes, size_t n, long long* out, const void** data, size_t* data_size, float* float_data, int float_cap,
                    int* n_float) {
  Datum d;
  if (!ParseDatum(bytes, n, &d)) return 0;
  out[0] = d.channels; out[1] = d.height; out[2] = d.width; out[3] = d.label; out[4] = d.encoded ? 1 : 0; out[5] = d.record_id;
  *data = d.data; *data_size = d.data_size;
  *n_float = (int)d.float_data.size();
  for (int i = 0; i < *n_float && i < float_cap; ++i) float_data[i] = d.float_data[i];
  return 1;
}
// serialises into `out` (cap bytes); returns the length, or -1 if it does not fit
long long b2h_datum_serialize(int channels, int height, int width, const void* data, size_t data_size, int label, int encoded,
                              const float* float_data, int n_float, void* out, size_t cap) {
  std::vector<float> fd;
  if (n_float > 0) fd.assign(float_data, float_data + n_float);
  const std::string s = SerializeDatum(channels, height, width, data, data_size, label, encoded != 0, n_float > 0 ? &fd : nullptr);
  if (s.size() > cap) { g_derr = "b2h_datum_serialize: buffer too small"; return -1; }
  memcpy(out, s.data(), s.size());
  return (long long)s.size();
}
// mean_file: a bare BlobProto.  shape gets up to 8 axes; data may be null (query the count first)
int b2h_blobproto_load(const char* path, int* ndim, int* shape, long long* count, float* data) {
  B2D_TRY({
    const BlobData b = ParseBlobProto(ReadBinaryFile(path));
    *ndim = (int)b.shape.size();
    for (size_t k = 0; k < b.shape.size() && k < 8; ++k) shape[k] = b.shape[k];
    *count = (long long)b.data.size();
    if (data) memcpy(data, b.data.data(), sizeof(float) * b.data.size());
  });
}
int b2h_blobproto_save(const char* path, int ndim, const int* shape, const float* data, int raw) {
  B2D_TRY({
    BlobData b;
    size_t cnt = 1;
    for (int i = 0; i < ndim; ++i) { b.shape.push_back(shape[i]); cnt *= (size_t)shape[i]; }
    b.data.assign(data, data + cnt);
    WriteBinaryFile(path, SerializeBlobProto(b, raw != 0));
  });
}

// ---- encoded datums ------------------------------------------------------------------------------------------------------------------
// DecodeDatumToCVMatNative / DecodeDatumToCVMat(force_color) + CVMatToDatum's [channel][row][column] layout, channels B, G, R
// (src/caffe/util/io.cpp:167-230).  chw[0..2] = channels, height, width; `out` may be null (query the shape); returns -1 on
// unsupported or damaged files.
int b2h_jpeg_decode(const void* bytes, size_t n, int force_color, int* chw, unsigned char* out, size_t cap) {
  B2D_TRY({
    DecodedImage img;
    DecodeImage(bytes, n, force_color != 0, &img);             // JPEG or PNG by signature (the name is historical)
    chw[0] = img.channels; chw[1] = img.height; chw[2] = img.width;
    if (out) {
      B2_CHECK(img.chw.size() <= cap, "b2h_jpeg_decode: buffer too small");
      memcpy(out, img.chw.data(), img.chw.size());
    }
  });
}

// ---- DataReader ----------------------------------------------------------------------------------------------------------------
void* b2h_data_reader_create(const char* source, int batch_size, int solver_count, int solver_rank, int node_count, int node_rank,
                             int parser_threads, int depth, int force_encoded_color) {
  try {
    DataReaderParam p;
    p.source = source; p.batch_size = batch_size;
    p.solver_count = (size_t)solver_count; p.solver_rank = (size_t)solver_rank;
    p.node_count = (size_t)node_count; p.node_rank = (size_t)node_rank;
    p.parser_threads = (size_t)parser_threads;
    p.force_encoded_color = force_encoded_color != 0;
    std::unique_ptr<ReaderHandle> h(new ReaderHandle);
    h->reader.reset(new DataReader(p));
    const size_t bytes = h->reader->datum_bytes() * (size_t)batch_size;
    const int nbuf = std::max(1, depth) * std::max(1, parser_threads);
    for (int i = 0; i < nbuf; ++i) {
      std::unique_ptr<ReaderHandle::Owned> o(new ReaderHandle::Owned);
      o->data.resize(bytes); o->label.resize(batch_size); o->ids.resize(batch_size);
      o->buf.data = o->data.data(); o->buf.label = o->label.data(); o->buf.record_id = o->ids.data();
      h->pool.push_back(std::move(o));
    }
    for (auto& o : h->pool) h->reader->free_push(&o->buf);
    return h.release();
  } catch (const std::exception& e) { g_derr = e.what(); return nullptr; }
}
void b2h_data_reader_destroy(void* hv) {
  auto* h = static_cast<ReaderHandle*>(hv);
  if (h) h->reader.reset();          // joins the parser threads before their buffers go away
  delete h;
}
int b2h_data_reader_info(void* hv, int* chw, long long* entries, long long* full_cycle) {
  auto* h = static_cast<ReaderHandle*>(hv);
  B2D_TRY({
    chw[0] = h->reader->channels(); chw[1] = h->reader->height(); chw[2] = h->reader->width();
    *entries = (long long)h->reader->entries(); *full_cycle = (long long)h->reader->full_cycle();
  });
}
long long b2h_data_reader_first_record(void* hv, long long batch) {
  return (long long)static_cast<ReaderHandle*>(hv)->reader->first_record_of_batch((size_t)batch);
}
// the next batch in batch order, copied out; the buffer goes straight back to the reader
int b2h_data_reader_next(void* hv, unsigned char* data, float* label, unsigned* record_id, long long* batch_id) {
  auto* h = static_cast<ReaderHandle*>(hv);
  B2D_TRY({
    BatchBuf* b = h->reader->full_pop();
    const size_t B = h->pool[0]->label.size();
    memcpy(data, b->data, h->reader->datum_bytes() * B);
    memcpy(label, b->label, sizeof(float) * B);
    if (record_id) memcpy(record_id, b->record_id, sizeof(uint32_t) * B);
    if (batch_id) *batch_id = (long long)b->batch_id;
    h->reader->free_push(b);
  });
}

// ---- DataTransformer's draws -----------------------------------------------------------------------------------------------------
int b2h_transform_draws(unsigned long long seed, int mirror, int crop, int train, int n, int datum_h, int datum_w, int* h_off, int* w_off,
                        unsigned char* do_mirror) {
  B2D_TRY({
    TransformDraws dr(seed, mirror != 0, crop, train != 0);
    for (int i = 0; i < n; ++i) dr.Draw(datum_h, datum_w, h_off + i, w_off + i, do_mirror + i);
  });
}

}  // extern "C"
