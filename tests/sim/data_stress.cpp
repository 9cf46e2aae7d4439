// data_stress.cpp -- sanitizer run of the input pipeline's host code (tests/test_data_sanitize.py builds it with
// -fsanitize=address,undefined and runs it).  TEST INFRASTRUCTURE ONLY.
//   1. db::LMDB writer against a std::map model: random commits (appends, overwrites, out-of-order keys), every state read back;
//   2. the reader on DAMAGED files: random byte flips and truncations of a valid database must end in caffe::FatalError or in a
//      clean walk -- never in a crash or an out-of-bounds access;
//   3. ParseDatum on random bytes and on truncated valid datums;
//   4. DataReader: parser threads started, drained and destroyed at random points;
//   5. the two parsers of files that come from outside -- prototxt text (models, solvers) and the .caffemodel / .solverstate wire
//      format (model-zoo weights) -- on mutated inputs: FatalError or a parse, never a crash;
//   6. the JPEG and PNG decoders on mutated copies of the files <scratch dir>/seed*.jpg|png (written by the test): flips in headers,
//      tables and entropy-coded / compressed data, truncations.
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <random>

#include "../../caffe_mpi_b200/host/b2caffe.hpp"
#include "../../caffe_mpi_b200/host/data_reader.hpp"
#include "../../caffe_mpi_b200/host/jpeg_decode.hpp"
#include "../../caffe_mpi_b200/host/prototxt.hpp"

using namespace caffe;

static std::string slurp(const std::string& f) {
  std::ifstream in(f, std::ios::binary);
  return std::string((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
}
static void spit(const std::string& f, const std::string& b) {
  std::ofstream out(f, std::ios::binary | std::ios::trunc);
  out.write(b.data(), (std::streamsize)b.size());
}
#define REQUIRE(c) do { if (!(c)) { fprintf(stderr, "data_stress: %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: data_stress <scratch dir> [seed]\n"); return 2; }
  const std::string dir = argv[1];
  mkdir(dir.c_str(), 0755);
  std::mt19937 rng(argc > 2 ? (unsigned)atoi(argv[2]) : 7u);
  auto rnd = [&](unsigned n) { return (unsigned)(rng() % n); };
  auto blob = [&](size_t n) { std::string s(n, '\0'); for (auto& c : s) c = (char)rnd(256); return s; };
  const bool threads_only = argc > 3 && std::string(argv[3]) == "threads";      // the ThreadSanitizer build runs section 4 alone
  std::map<std::string, std::string> model;
  int refused = 0, walked = 0;
  if (!threads_only) {

  // ---- 1. writer vs model
  const std::string db = dir + "/w";
  {
    db::LMDB env;
    env.Open(db, db::NEW);
    const size_t sizes[] = {0, 1, 17, 600, 2010, 2025, 2040, 4096, 9000, 70000};
    unsigned next_id = 0;
    for (int commit = 0; commit < 40; ++commit) {
      std::unique_ptr<db::LMDBTransaction> txn(env.NewTransaction());
      const int n = (int)rnd(30);
      for (int i = 0; i < n; ++i) {
        char key[64];
        const bool ascending = rnd(10) < 8;
        snprintf(key, sizeof(key), "%08u_k", ascending ? next_id++ : rnd(next_id + 1));
        const std::string v = blob(sizes[rnd(10)]);
        txn->Put(key, v);
        model[key] = v;
      }
      txn->Commit();
      REQUIRE(env.entries() == model.size());
      std::unique_ptr<db::LMDBCursor> cur(env.NewCursor());
      auto it = model.begin();
      for (; cur->valid(); cur->Next(), ++it) {
        REQUIRE(it != model.end());
        REQUIRE(cur->key() == it->first && cur->value() == it->second);
      }
      REQUIRE(it == model.end());
    }
  }
  // ---- 2. damaged files
  const std::string good = slurp(db + "/data.mdb");
  REQUIRE(good.size() > 8192);
  mkdir((dir + "/bad").c_str(), 0755);
  for (int trial = 0; trial < 400; ++trial) {
    std::string bad = good;
    if (trial % 4 == 0) bad.resize(4096 * (2 + rnd((unsigned)(good.size() / 4096 - 2))) + rnd(4096));
    const int flips = 1 + (int)rnd(6);
    for (int f = 0; f < flips; ++f) {
      // mostly page headers, node offset tables and the meta pages: where the structure lives
      size_t at = (size_t)rnd((unsigned)(bad.size() / 4096)) * 4096 + (rnd(3) ? rnd(64) : rnd(4096));
      if (rnd(4) == 0) at = rnd(2) * 4096 + 16 + rnd(140);
      if (at < bad.size()) bad[at] = (char)rnd(256);
    }
    spit(dir + "/bad/data.mdb", bad);
    try {
      db::LMDB env;
      env.Open(dir + "/bad", db::READ);
      std::unique_ptr<db::LMDBCursor> cur(env.NewCursor());
      size_t n = 0, bytes = 0;
      for (; cur->valid() && n < 100000; cur->Next(), ++n) {
        const std::string k = cur->key(), v = cur->value();      // touches every byte the cursor claims
        bytes += k.size() + v.size();
        Datum d;
        ParseDatum(cur->data(), cur->size(), &d);
      }
      ++walked;
    } catch (const FatalError&) { ++refused; }
  }
  REQUIRE(refused > 0 && walked > 0);
  // ---- 3. Datum parser
  const uint8_t px[24] = {1, 2, 3};
  const std::string datum = SerializeDatum(2, 3, 4, px, sizeof(px), 5);
  for (size_t cut = 0; cut <= datum.size(); ++cut) { Datum d; ParseDatum(datum.data(), cut, &d); }
  for (int trial = 0; trial < 2000; ++trial) { const std::string junk = blob(rnd(64)); Datum d; ParseDatum(junk.data(), junk.size(), &d); }
  }  // !threads_only
  // ---- 4. DataReader life cycle
  {
    const std::string ddb = dir + "/d";
    db::LMDB env;
    env.Open(ddb, db::NEW);
    std::unique_ptr<db::LMDBTransaction> txn(env.NewTransaction());
    for (int i = 0; i < 23; ++i) {
      char key[32]; snprintf(key, sizeof(key), "%08d_x", i);
      const std::string img = blob(3 * 6 * 5);
      txn->Put(key, SerializeDatum(3, 6, 5, img.data(), img.size(), i));
    }
    txn->Commit();
    env.Close();
    for (int trial = 0; trial < 30; ++trial) {
      DataReaderParam p;
      p.source = ddb; p.batch_size = 1 + (int)rnd(5); p.parser_threads = 1 + rnd(3);
      p.solver_count = 1 + rnd(3); p.solver_rank = rnd((unsigned)p.solver_count);
      // the buffers are declared BEFORE the reader: they must outlive it, its threads write into them until it is destroyed
      // (the data layer and the C binding release theirs after reader.reset() for the same reason)
      const int nbuf = 1 + (int)rnd(4);
      std::vector<std::vector<uint8_t>> data(nbuf, std::vector<uint8_t>((size_t)3 * 6 * 5 * p.batch_size));
      std::vector<std::vector<float>> lab(nbuf, std::vector<float>(p.batch_size));
      std::vector<BatchBuf> bufs(nbuf);
      DataReader rd(p);
      REQUIRE(rd.datum_bytes() == 3 * 6 * 5);
      for (int i = 0; i < nbuf; ++i) { bufs[i].data = data[i].data(); bufs[i].label = lab[i].data(); rd.free_push(&bufs[i]); }
      const int pops = (int)rnd(12);
      for (int i = 0; i < pops; ++i) {
        BatchBuf* b = rd.full_pop();
        REQUIRE(b->batch_id == (size_t)i);
        unsigned sum = 0;
        for (size_t k = 0; k < rd.datum_bytes() * p.batch_size; ++k) sum += b->data[k];      // read what the parser thread wrote
        REQUIRE(sum > 0 && b->label[0] >= 0.f);
        rd.free_push(b);
      }
    }                                                      // ~DataReader with batches in flight
  }
  // ---- 5. text and wire parsers on mutated input
  if (!threads_only) {
    const std::string net =
        "name: \"fuzz\"\n"
        "layer { name: \"in\" type: \"Input\" top: \"data\" top: \"label\" input_param { shape { dim: 4 dim: 3 dim: 12 dim: 12 } shape { dim: 4 } } }\n"
        "layer { name: \"c1\" type: \"Convolution\" bottom: \"data\" top: \"c1\" param { lr_mult: 1 decay_mult: 1 } param { lr_mult: 2 decay_mult: 0 }\n"
        "  convolution_param { num_output: 8 kernel_size: 3 pad: 1 stride: 2 group: 1 weight_filler { type: \"msra\" } bias_filler { type: \"constant\" value: 0.1 } } }\n"
        "layer { name: \"bn\" type: \"BatchNorm\" bottom: \"c1\" top: \"bn\" batch_norm_param { scale_bias: true eps: 1e-4 } }\n"
        "layer { name: \"r\" type: \"ReLU\" bottom: \"bn\" top: \"bn\" }  # in place\n"
        "layer { name: \"p\" type: \"Pooling\" bottom: \"bn\" top: \"p\" pooling_param { pool: MAX kernel_size: 3 stride: 2 } }\n"
        "layer { name: \"ip\" type: \"InnerProduct\" bottom: \"p\" top: \"ip\" inner_product_param { num_output: 10 } }\n"
        "layer { name: \"acc\" type: \"Accuracy\" bottom: \"ip\" bottom: \"label\" top: \"acc\" include { phase: TEST } }\n"
        "layer { name: \"loss\" type: \"SoftmaxWithLoss\" bottom: \"ip\" bottom: \"label\" top: \"loss\" loss_weight: 1 }\n";
    { Net ok(ParseTextProto(net), TRAIN); REQUIRE(ok.layers().size() == 7 && ok.learnable_params().size() == 6); }   // TEST-only Accuracy filtered out; BatchNorm lists its scale and bias here
    const char alphabet[] = "{}:\"#\n 0123456789-.eE_abclmnoprstuyINPUTX";
    int parsed = 0, rejected = 0;
    for (int trial = 0; trial < 1500; ++trial) {
      std::string t = net;
      const int edits = 1 + (int)rnd(4);
      for (int e = 0; e < edits; ++e) {
        const size_t at = rnd((unsigned)t.size());
        switch (rnd(4)) {
          case 0: t[at] = alphabet[rnd(sizeof(alphabet) - 1)]; break;
          case 1: t.erase(at, 1 + rnd(12)); break;
          case 2: t.insert(at, 1, alphabet[rnd(sizeof(alphabet) - 1)]); break;
          default: t.insert(at, t.substr(rnd((unsigned)t.size()), rnd(40))); break;
        }
      }
      try { Net n(ParseTextProto(t), rnd(2) ? TRAIN : TEST, (int)rnd(3)); ReadSolverParameter(ParseTextProto(t)); ++parsed; }
      catch (const FatalError&) { ++rejected; }
      catch (const std::exception&) { ++rejected; }          // std::stoi and friends on damaged numbers
    }
    REQUIRE(parsed > 0 && rejected > 0);
    NetWeights w;
    w.name = "fuzz";
    for (int l = 0; l < 3; ++l) {
      LayerWeights lw; lw.name = "layer" + std::to_string(l); lw.type = "Convolution"; lw.bottom = {"a"}; lw.top = {"b"};
      BlobData b; b.shape = {2, 3, 1 + l, 2}; b.data.resize((size_t)2 * 3 * (1 + l) * 2, 0.5f * l);
      lw.blobs.push_back(b);
      w.layers.push_back(lw);
    }
    SolverStateData st; st.iter = 7; st.learned_net = "x.caffemodel"; st.history.push_back(w.layers[0].blobs[0]);
    for (int raw = 0; raw < 2; ++raw) {
      const std::string model = SerializeNetWeights(w, raw != 0), state = SerializeSolverState(st, raw != 0);
      REQUIRE(ParseNetWeights(model).layers.size() == 3 && ParseSolverState(state).iter == 7);
      for (int trial = 0; trial < 1500; ++trial) {
        std::string m = trial % 2 ? model : state;
        if (rnd(5) == 0) m.resize(rnd((unsigned)m.size() + 1));
        const int flips = 1 + (int)rnd(5);
        for (int f = 0; f < flips && !m.empty(); ++f) m[rnd((unsigned)m.size())] = (char)rnd(256);
        try { if (trial % 2) ParseNetWeights(m); else ParseSolverState(m); ParseBlobProto(m); } catch (const FatalError&) {}
      }
    }
  }
  // ---- 6. JPEG decoder on damaged files
  if (!threads_only) {
    int decoded = 0, rejected = 0, files = 0;
    for (int f = 0; f < 8; ++f) {
      std::string jpg = slurp(dir + "/seed" + std::to_string(f) + ".jpg");
      if (jpg.empty()) jpg = slurp(dir + "/seed" + std::to_string(f) + ".png");
      if (jpg.empty()) continue;
      ++files;
      { DecodedImage img; DecodeImage(jpg.data(), jpg.size(), false, &img); REQUIRE(img.channels >= 1 && img.height > 0 && img.width > 0); }
      for (int trial = 0; trial < 1200; ++trial) {
        std::string m = jpg;
        if (rnd(6) == 0) m.resize(rnd((unsigned)m.size() + 1));
        const int flips = 1 + (int)rnd(4);
        for (int k = 0; k < flips && !m.empty(); ++k) {
          const size_t at = rnd(3) ? rnd((unsigned)std::min<size_t>(m.size(), 700)) : rnd((unsigned)m.size());   // headers and tables first
          m[at] = (char)rnd(256);
        }
        try { DecodedImage img; DecodeImage(m.data(), m.size(), rnd(2) != 0, &img); ++decoded; } catch (const FatalError&) { ++rejected; }
      }
    }
    if (files) REQUIRE(decoded > 0 && rejected > 0);
  }
  printf("data_stress ok: %zu records, damaged files: %d refused, %d walked\n", model.size(), refused, walked);
  return 0;
}
