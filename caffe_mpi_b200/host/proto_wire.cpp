// proto_wire.cpp -- see proto_wire.hpp.
#include "proto_wire.hpp"

#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>

#include "b2caffe.hpp"

namespace caffe {
namespace {

// ---- writer -------------------------------------------------------------------------------------------------------
void put_varint(std::string& o, uint64_t v) {
  while (v >= 0x80) { o.push_back((char)((v & 0x7f) | 0x80)); v >>= 7; }
  o.push_back((char)v);
}
void put_tag(std::string& o, int field, int wt) { put_varint(o, ((uint64_t)field << 3) | (uint64_t)wt); }
void put_bytes(std::string& o, int field, const std::string& b) { put_tag(o, field, 2); put_varint(o, b.size()); o += b; }
void put_int(std::string& o, int field, int64_t v) { put_tag(o, field, 0); put_varint(o, (uint64_t)v); }

std::string blob_bytes(const BlobData& b, bool raw) {
  std::string shape;                               // BlobShape { repeated int64 dim = 1 [packed] }
  { std::string dims; for (int d : b.shape) put_varint(dims, (uint64_t)(int64_t)d); put_bytes(shape, 1, dims); }
  std::string o;
  const std::string payload(reinterpret_cast<const char*>(b.data.data()), b.data.size() * sizeof(float));
  if (!raw) put_bytes(o, 5, payload);              // repeated float data = 5 [packed]: little-endian IEEE words
  put_bytes(o, 7, shape);
  if (raw) { put_int(o, 10, 1 /* FLOAT */); put_bytes(o, 12, payload); }
  return o;
}

// ---- reader -------------------------------------------------------------------------------------------------------
struct Cursor {
  const uint8_t* p; const uint8_t* end;
  bool done() const { return p >= end; }
  uint64_t varint() {
    uint64_t v = 0; int shift = 0;
    for (;;) {
      B2_CHECK(p < end && shift < 64, "protobuf: truncated varint");
      const uint8_t c = *p++;
      v |= (uint64_t)(c & 0x7f) << shift;
      if (!(c & 0x80)) return v;
      shift += 7;
    }
  }
  Cursor sub() { const uint64_t n = varint(); B2_CHECK(n <= (uint64_t)(end - p), "protobuf: length past the end"); Cursor c{p, p + n}; p += n; return c; }
  std::string str() { Cursor c = sub(); return std::string(reinterpret_cast<const char*>(c.p), c.end - c.p); }
  void skip(int wt) {
    if (wt == 0) varint();
    else if (wt == 1) { B2_CHECK(end - p >= 8, "protobuf: truncated fixed64"); p += 8; }
    else if (wt == 2) sub();
    else if (wt == 5) { B2_CHECK(end - p >= 4, "protobuf: truncated fixed32"); p += 4; }
    else B2_CHECK(false, "protobuf: unsupported wire type " + std::to_string(wt));
  }
};

float half_to_float(uint16_t h) {
  const uint32_t s = (h >> 15) & 1, e = (h >> 10) & 0x1f, m = h & 0x3ff;
  float v;
  if (e == 0) v = std::ldexp((float)m, -24);
  else if (e == 31) v = m ? NAN : INFINITY;
  else v = std::ldexp((float)(m | 0x400), (int)e - 25);
  return s ? -v : v;
}

BlobData parse_blob(Cursor c) {
  BlobData b;
  int legacy[4] = {0, 0, 0, 0};
  bool has_legacy = false;
  int raw_type = -1;
  std::string raw;
  std::vector<double> dd;
  while (!c.done()) {
    const uint64_t key = c.varint();
    const int f = (int)(key >> 3), wt = (int)(key & 7);
    if (f >= 1 && f <= 4 && wt == 0) { legacy[f - 1] = (int)c.varint(); has_legacy = true; }
    else if (f == 7 && wt == 2) {
      Cursor s = c.sub();
      while (!s.done()) {
        const uint64_t k2 = s.varint();
        if ((k2 >> 3) == 1 && (k2 & 7) == 2) { Cursor d = s.sub(); while (!d.done()) b.shape.push_back((int)(int64_t)d.varint()); }
        else if ((k2 >> 3) == 1 && (k2 & 7) == 0) b.shape.push_back((int)(int64_t)s.varint());
        else s.skip((int)(k2 & 7));
      }
    } else if (f == 5 && wt == 2) {
      Cursor d = c.sub();
      const size_t n = (size_t)(d.end - d.p) / 4, at = b.data.size();
      b.data.resize(at + n);
      if (n) memcpy(b.data.data() + at, d.p, n * 4);
    } else if (f == 5 && wt == 5) {
      B2_CHECK(c.end - c.p >= 4, "protobuf: truncated float"); float v; memcpy(&v, c.p, 4); c.p += 4; b.data.push_back(v);
    } else if (f == 8 && wt == 2) {
      Cursor d = c.sub();
      const size_t n = (size_t)(d.end - d.p) / 8, at = dd.size();
      dd.resize(at + n);
      if (n) memcpy(dd.data() + at, d.p, n * 8);
    } else if (f == 8 && wt == 1) {
      B2_CHECK(c.end - c.p >= 8, "protobuf: truncated double"); double v; memcpy(&v, c.p, 8); c.p += 8; dd.push_back(v);
    } else if (f == 10 && wt == 0) raw_type = (int)c.varint();
    else if (f == 12 && wt == 2) raw = c.str();
    else c.skip(wt);
  }
  // Blob::FromProto precedence: double_data, then data, then raw_data (blob.cpp:375-414)
  if (!dd.empty()) { b.data.assign(dd.begin(), dd.end()); }
  else if (b.data.empty() && !raw.empty()) {
    B2_CHECK(raw_type >= 0, "Missing raw data type");
    if (raw_type == 1) { b.data.resize(raw.size() / 4); if (!b.data.empty()) memcpy(b.data.data(), raw.data(), b.data.size() * 4); }
    else if (raw_type == 0) { b.data.resize(raw.size() / 8); for (size_t i = 0; i < b.data.size(); ++i) { double v; memcpy(&v, raw.data() + 8 * i, 8); b.data[i] = (float)v; } }
    else if (raw_type == 2) { b.data.resize(raw.size() / 2); for (size_t i = 0; i < b.data.size(); ++i) { uint16_t h; memcpy(&h, raw.data() + 2 * i, 2); b.data[i] = half_to_float(h); } }
    else B2_CHECK(false, "Unsupported raw type " + std::to_string(raw_type));
  }
  if (b.shape.empty() && has_legacy) b.shape.assign(legacy, legacy + 4);     // (num, channels, height, width), blob.cpp:355-363
  size_t cnt = 1;
  for (int d : b.shape) cnt *= (size_t)d;
  B2_CHECK(b.data.empty() || cnt == b.data.size(), "BlobProto: shape and data size disagree");
  return b;
}

LayerWeights parse_layer(Cursor c, bool v1) {
  LayerWeights L;
  const int f_name = v1 ? 4 : 1, f_blobs = v1 ? 6 : 7, f_bottom = v1 ? 2 : 3, f_top = v1 ? 3 : 4;
  while (!c.done()) {
    const uint64_t key = c.varint();
    const int f = (int)(key >> 3), wt = (int)(key & 7);
    if (f == f_name && wt == 2) L.name = c.str();
    else if (!v1 && f == 2 && wt == 2) L.type = c.str();
    else if (v1 && f == 5 && wt == 0) L.type = "V1:" + std::to_string(c.varint());
    else if (f == f_bottom && wt == 2) L.bottom.push_back(c.str());
    else if (f == f_top && wt == 2) L.top.push_back(c.str());
    else if (f == f_blobs && wt == 2) L.blobs.push_back(parse_blob(c.sub()));
    else c.skip(wt);
  }
  return L;
}

}  // namespace

std::string SerializeNetWeights(const NetWeights& net, bool raw) {
  std::string o;
  put_bytes(o, 1, net.name);
  for (const LayerWeights& L : net.layers) {
    std::string l;
    put_bytes(l, 1, L.name);
    put_bytes(l, 2, L.type);
    for (auto& b : L.bottom) put_bytes(l, 3, b);
    for (auto& t : L.top) put_bytes(l, 4, t);
    for (auto& b : L.blobs) put_bytes(l, 7, blob_bytes(b, raw));
    put_bytes(o, 100, l);
  }
  return o;
}

NetWeights ParseNetWeights(const std::string& bytes) {
  NetWeights net;
  Cursor c{reinterpret_cast<const uint8_t*>(bytes.data()), reinterpret_cast<const uint8_t*>(bytes.data()) + bytes.size()};
  while (!c.done()) {
    const uint64_t key = c.varint();
    const int f = (int)(key >> 3), wt = (int)(key & 7);
    if (f == 1 && wt == 2) net.name = c.str();
    else if (f == 100 && wt == 2) net.layers.push_back(parse_layer(c.sub(), false));
    else if (f == 2 && wt == 2) net.layers.push_back(parse_layer(c.sub(), true));      // V1LayerParameter (old model zoo files)
    else c.skip(wt);
  }
  return net;
}

std::string SerializeSolverState(const SolverStateData& st, bool raw) {
  std::string o;
  put_int(o, 1, st.iter);
  put_bytes(o, 2, st.learned_net);
  for (auto& b : st.history) put_bytes(o, 3, blob_bytes(b, raw));
  put_int(o, 4, st.current_step);
  return o;
}

SolverStateData ParseSolverState(const std::string& bytes) {
  SolverStateData st;
  Cursor c{reinterpret_cast<const uint8_t*>(bytes.data()), reinterpret_cast<const uint8_t*>(bytes.data()) + bytes.size()};
  while (!c.done()) {
    const uint64_t key = c.varint();
    const int f = (int)(key >> 3), wt = (int)(key & 7);
    if (f == 1 && wt == 0) st.iter = (int)c.varint();
    else if (f == 2 && wt == 2) st.learned_net = c.str();
    else if (f == 3 && wt == 2) st.history.push_back(parse_blob(c.sub()));
    else if (f == 4 && wt == 0) st.current_step = (int)c.varint();
    else c.skip(wt);
  }
  return st;
}

bool ParseDatum(const void* bytes, size_t n, Datum* d) {
  *d = Datum();
  try {
    Cursor c{static_cast<const uint8_t*>(bytes), static_cast<const uint8_t*>(bytes) + n};
    while (!c.done()) {
      const uint64_t key = c.varint();
      const int f = (int)(key >> 3), wt = (int)(key & 7);
      if (f == 1 && wt == 0) d->channels = (int)(int64_t)c.varint();
      else if (f == 2 && wt == 0) d->height = (int)(int64_t)c.varint();
      else if (f == 3 && wt == 0) d->width = (int)(int64_t)c.varint();
      else if (f == 4 && wt == 2) { Cursor b = c.sub(); d->data = b.p; d->data_size = (size_t)(b.end - b.p); }
      else if (f == 5 && wt == 0) d->label = (int)(int64_t)c.varint();
      else if (f == 6 && wt == 2) {
        Cursor b = c.sub();
        if ((b.end - b.p) % 4) return false;
        const size_t k = (size_t)(b.end - b.p) / 4, at = d->float_data.size();
        d->float_data.resize(at + k);
        if (k) memcpy(d->float_data.data() + at, b.p, k * 4);
      } else if (f == 6 && wt == 5) {
        if (c.end - c.p < 4) return false;
        float v; memcpy(&v, c.p, 4); c.p += 4; d->float_data.push_back(v);
      }
      else if (f == 7 && wt == 0) d->encoded = c.varint() != 0;
      else if (f == 8 && wt == 0) d->record_id = (uint32_t)c.varint();
      else if (f == 0) return false;
      else c.skip(wt);
    }
  } catch (const FatalError&) { return false; }
  return true;
}

std::string SerializeDatum(int channels, int height, int width, const void* data, size_t data_size, int label, bool encoded,
                           const std::vector<float>* float_data) {
  std::string o;
  put_int(o, 1, channels); put_int(o, 2, height); put_int(o, 3, width);
  if (data_size) put_bytes(o, 4, std::string(static_cast<const char*>(data), data_size));
  put_int(o, 5, label);
  if (float_data) for (float v : *float_data) { put_tag(o, 6, 5); o.append(reinterpret_cast<const char*>(&v), 4); }   // unpacked, as proto2 writes it
  if (encoded) put_int(o, 7, 1);
  return o;
}

BlobData ParseBlobProto(const std::string& bytes) {
  return parse_blob(Cursor{reinterpret_cast<const uint8_t*>(bytes.data()), reinterpret_cast<const uint8_t*>(bytes.data()) + bytes.size()});
}
std::string SerializeBlobProto(const BlobData& b, bool raw) { return blob_bytes(b, raw); }

void WriteBinaryFile(const std::string& path, const std::string& bytes) {
  std::ofstream f(path, std::ios::binary | std::ios::trunc);
  B2_CHECK((bool)f, "Cannot open " + path + " for writing");
  f.write(bytes.data(), (std::streamsize)bytes.size());
  B2_CHECK((bool)f, "Short write to " + path);
}
std::string ReadBinaryFile(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  B2_CHECK((bool)f, "File not found: " + path);
  return std::string((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

}  // namespace caffe
