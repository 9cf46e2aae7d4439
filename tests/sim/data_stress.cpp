// data_stress.cpp -- sanitizer run of the input pipeline's host code (tests/test_data_sanitize.py builds it with
// -fsanitize=address,undefined and runs it).  TEST INFRASTRUCTURE ONLY.
//   1. db::LMDB writer against a std::map model: random commits (appends, overwrites, out-of-order keys), every state read back;
//   2. the reader on DAMAGED files: random byte flips and truncations of a valid database must end in caffe::FatalError or in a
//      clean walk -- never in a crash or an out-of-bounds access;
//   3. ParseDatum on random bytes and on truncated valid datums;
//   4. DataReader: parser threads started, drained and destroyed at random points.
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <random>

#include "../../caffe_mpi_b200/host/b2caffe.hpp"
#include "../../caffe_mpi_b200/host/data_reader.hpp"

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
  std::mt19937 rng(argc > 2 ? (unsigned)atoi(argv[2]) : 7u);
  auto rnd = [&](unsigned n) { return (unsigned)(rng() % n); };
  auto blob = [&](size_t n) { std::string s(n, '\0'); for (auto& c : s) c = (char)rnd(256); return s; };

  // ---- 1. writer vs model
  const std::string db = dir + "/w";
  std::map<std::string, std::string> model;
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
  int refused = 0, walked = 0;
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
      DataReader rd(p);
      const int nbuf = 1 + (int)rnd(4);
      std::vector<std::vector<uint8_t>> data(nbuf, std::vector<uint8_t>(rd.datum_bytes() * p.batch_size));
      std::vector<std::vector<float>> lab(nbuf, std::vector<float>(p.batch_size));
      std::vector<BatchBuf> bufs(nbuf);
      for (int i = 0; i < nbuf; ++i) { bufs[i].data = data[i].data(); bufs[i].label = lab[i].data(); rd.free_push(&bufs[i]); }
      const int pops = (int)rnd(12);
      for (int i = 0; i < pops; ++i) { BatchBuf* b = rd.full_pop(); REQUIRE(b->batch_id == (size_t)i); rd.free_push(b); }
    }                                                      // ~DataReader with batches in flight
  }
  printf("data_stress ok: %zu records, damaged files: %d refused, %d walked\n", model.size(), refused, walked);
  return 0;
}
