// lmdb_reader.cpp -- see lmdb_reader.hpp.
#include "lmdb_reader.hpp"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <cstring>
#include <map>

#include "b2caffe.hpp"

namespace caffe { namespace db {
namespace {

constexpr unsigned kPageHdr = 16;
constexpr unsigned kNodeHdr = 8;
constexpr uint16_t P_BRANCH = 0x01, P_LEAF = 0x02, P_OVERFLOW = 0x04, P_META = 0x08, P_LEAF2 = 0x20;
constexpr uint16_t F_BIGDATA = 0x01, F_SUBDATA = 0x02, F_DUPDATA = 0x04;
constexpr uint32_t kMagic = 0xBEEFC0DE;
constexpr uint64_t kInvalid = ~0ull;

inline uint16_t rd16(const uint8_t* p) { uint16_t v; memcpy(&v, p, 2); return v; }
inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline uint64_t rd64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }

// NUMKEYS(page) with the sanity a corrupted file needs: `lower` must lie inside the page and after the header
unsigned num_keys(const uint8_t* pg, unsigned psize) {
  const unsigned lower = rd16(pg + 12);
  B2_CHECK(lower >= kPageHdr && lower <= psize && rd16(pg + 14) <= psize, "lmdb: MDB_CORRUPTED: page bounds outside the page");
  return (lower - kPageHdr) >> 1;
}

std::string data_file(const std::string& source) {
  struct stat st;
  if (stat(source.c_str(), &st) == 0 && S_ISDIR(st.st_mode)) return source + "/data.mdb";
  return source;
}

struct Meta { bool ok = false; uint32_t psize = 0; uint16_t depth = 0; uint64_t entries = 0, root = kInvalid, last_pg = 0, txnid = 0; uint16_t flags = 0; };
Meta read_meta(const uint8_t* pg) {
  Meta m;
  if (!(rd16(pg + 10) & P_META)) return m;
  const uint8_t* mm = pg + kPageHdr;
  if (rd32(mm) != kMagic || rd32(mm + 4) != 1) return m;
  const uint8_t* free_db = mm + 24;                 // after magic, version, address, mapsize
  const uint8_t* main_db = free_db + 48;
  m.psize = rd32(free_db);                           // mm_psize = mm_dbs[FREE_DBI].md_pad
  m.flags = rd16(main_db + 4);
  m.depth = rd16(main_db + 6);
  m.entries = rd64(main_db + 32);
  m.root = rd64(main_db + 40);
  m.last_pg = rd64(main_db + 48);
  m.txnid = rd64(main_db + 56);
  m.ok = true;
  return m;
}

}  // namespace

// ------------------------------------------------------------------------------------------------------------ writer
// Bulk builder behind Mode NEW / WRITE (see the header): packs leaf pages left to right, sends large values to overflow runs,
// and on every commit writes the partial leaf, a fresh set of branch pages and the transaction's meta page.
struct LMDB::Writer {
  static constexpr unsigned kPsize = 4096;
  static constexpr unsigned kNodeMax = (((kPsize - kPageHdr) / 2) & ~1u) - 2;      // me_nodemax for 4 KiB pages: 2038
  static constexpr size_t kMapSize = 1099511627776ull;                              // LMDB_MAP_SIZE, db_lmdb.hpp:16-20

  struct PageBuilder {                       // node offsets grow up from byte 16, nodes grow down from the end (MDB_page)
    std::vector<uint8_t> buf;
    std::vector<uint16_t> ptrs;
    unsigned upper = kPsize;
    PageBuilder() : buf(kPsize, 0) {}
    bool room(size_t node_bytes) const { return upper - (kPageHdr + 2 * ptrs.size()) >= ((node_bytes + 1) & ~(size_t)1) + 2; }
    uint8_t* add(size_t node_bytes) {
      upper -= (unsigned)((node_bytes + 1) & ~(size_t)1);
      ptrs.push_back((uint16_t)upper);
      return buf.data() + upper;
    }
    const uint8_t* finish(uint64_t pgno, uint16_t flags) {
      const uint16_t pad = 0, lower = (uint16_t)(kPageHdr + 2 * ptrs.size()), up = (uint16_t)upper;
      memcpy(buf.data(), &pgno, 8); memcpy(buf.data() + 8, &pad, 2); memcpy(buf.data() + 10, &flags, 2);
      memcpy(buf.data() + 12, &lower, 2); memcpy(buf.data() + 14, &up, 2);
      memcpy(buf.data() + kPageHdr, ptrs.data(), 2 * ptrs.size());
      return buf.data();
    }
  };
  struct Ref { std::string first_key; uint64_t pgno; };

  int fd = -1;
  bool indexed = false;                      // `leaves` / `cur` describe the file (this object built it)
  uint64_t next_pg = 2, txnid = 0, n_overflow = 0, entries = 0;
  std::vector<Ref> leaves;                   // finished leaf pages, in key order
  PageBuilder cur;                           // the rightmost, still filling leaf
  uint64_t cur_pgno = 0;
  std::string cur_first, last_key;
  bool cur_open = false, have_last = false;

  ~Writer() { if (fd >= 0) ::close(fd); }

  void put_page(uint64_t pgno, const void* bytes, size_t n) {
    const uint8_t* p = static_cast<const uint8_t*>(bytes);
    off_t off = (off_t)(pgno * kPsize);
    while (n) {
      const ssize_t k = ::pwrite(fd, p, n, off);
      B2_CHECK(k > 0, std::string("lmdb: write failed: ") + std::strerror(errno));
      p += k; n -= (size_t)k; off += k;
    }
  }
  void put_meta(unsigned which, uint64_t tx, uint16_t depth, uint64_t branch, uint64_t leaf, uint64_t overflow, uint64_t n, uint64_t root,
                uint64_t last_pg) {
    std::vector<uint8_t> pg(kPsize, 0);
    const uint64_t pgno = which; const uint16_t flags = P_META;
    memcpy(pg.data(), &pgno, 8); memcpy(pg.data() + 10, &flags, 2);
    uint8_t* mm = pg.data() + kPageHdr;
    const uint32_t magic = kMagic, version = 1, psize = kPsize;
    const uint64_t address = 0, mapsize = kMapSize, invalid = kInvalid;
    memcpy(mm, &magic, 4); memcpy(mm + 4, &version, 4); memcpy(mm + 8, &address, 8); memcpy(mm + 16, &mapsize, 8);
    uint8_t* free_db = mm + 24;
    memcpy(free_db, &psize, 4); memcpy(free_db + 40, &invalid, 8);                      // empty free list; md_pad = page size
    uint8_t* main_db = free_db + 48;
    memcpy(main_db + 6, &depth, 2); memcpy(main_db + 8, &branch, 8); memcpy(main_db + 16, &leaf, 8); memcpy(main_db + 24, &overflow, 8);
    memcpy(main_db + 32, &n, 8); memcpy(main_db + 40, &root, 8);
    memcpy(main_db + 48, &last_pg, 8); memcpy(main_db + 56, &tx, 8);
    put_page(which, pg.data(), kPsize);
  }
  void open_fd(const std::string& file, bool create) {
    if (fd >= 0) ::close(fd);
    fd = ::open(file.c_str(), create ? (O_RDWR | O_CREAT | O_TRUNC) : O_RDWR, 0664);
    B2_CHECK(fd >= 0, "Failed to open lmdb " + file + " for writing: " + std::strerror(errno));
  }
  void reset_tree() {
    next_pg = 2; n_overflow = 0; entries = 0; leaves.clear(); cur = PageBuilder(); cur_open = false; have_last = false; last_key.clear();
  }
  void CreateEmpty(const std::string& file) {           // what mdb_env_open leaves behind for a new environment
    open_fd(file, true);
    reset_tree();
    txnid = 0;
    put_meta(0, 0, 0, 0, 0, 0, 0, kInvalid, 1);
    put_meta(1, 0, 0, 0, 0, 0, 0, kInvalid, 1);
    indexed = true;
  }
  void Append(const std::string& key, const std::string& val) {
    B2_CHECK(!key.empty() && key.size() <= 511, "lmdb: MDB_BAD_VALSIZE: key size must be 1..511");
    B2_CHECK(val.size() <= 0xffffffffull, "lmdb: MDB_BAD_VALSIZE: value too large");
    const bool big = kNodeHdr + key.size() + val.size() > kNodeMax;
    const size_t node_bytes = kNodeHdr + key.size() + (big ? 8 : val.size());
    if (!cur_open || !cur.room(node_bytes)) {
      if (cur_open) { put_page(cur_pgno, cur.finish(cur_pgno, P_LEAF), kPsize); leaves.push_back(Ref{cur_first, cur_pgno}); }
      cur = PageBuilder();
      cur_pgno = next_pg++;
      cur_first = key;
      cur_open = true;
    }
    uint64_t opg = 0;
    if (big) {                                           // OVPAGES(size, psize) = (PAGEHDRSZ - 1 + size) / psize + 1
      const uint32_t npg = (uint32_t)((kPageHdr - 1 + val.size()) / kPsize + 1);
      opg = next_pg;
      next_pg += npg;
      std::vector<uint8_t> run((size_t)npg * kPsize, 0);
      const uint16_t flags = P_OVERFLOW;
      memcpy(run.data(), &opg, 8); memcpy(run.data() + 10, &flags, 2); memcpy(run.data() + 12, &npg, 4);
      memcpy(run.data() + kPageHdr, val.data(), val.size());
      put_page(opg, run.data(), run.size());
      n_overflow += npg;
    }
    uint8_t* nd = cur.add(node_bytes);
    const uint16_t lo = (uint16_t)(val.size() & 0xffff), hi = (uint16_t)(val.size() >> 16), fl = big ? F_BIGDATA : 0, ks = (uint16_t)key.size();
    memcpy(nd, &lo, 2); memcpy(nd + 2, &hi, 2); memcpy(nd + 4, &fl, 2); memcpy(nd + 6, &ks, 2);
    memcpy(nd + kNodeHdr, key.data(), key.size());
    if (big) memcpy(nd + kNodeHdr + key.size(), &opg, 8);
    else if (!val.empty()) memcpy(nd + kNodeHdr + key.size(), val.data(), val.size());
    last_key = key; have_last = true;
    ++entries;
  }
  void FinishCommit() {
    ++txnid;
    std::vector<Ref> level = leaves;
    if (cur_open) { put_page(cur_pgno, cur.finish(cur_pgno, P_LEAF), kPsize); level.push_back(Ref{cur_first, cur_pgno}); }
    const uint64_t n_leaf = level.size();
    uint64_t n_branch = 0, root = kInvalid;
    uint16_t depth = 0;
    if (!level.empty()) {
      depth = 1;
      while (level.size() > 1) {                         // one branch level per pass, pages written as they fill
        std::vector<Ref> parents;
        PageBuilder b;
        uint64_t bpg = 0;
        bool open = false;
        std::string bfirst;
        auto flush = [&] { put_page(bpg, b.finish(bpg, P_BRANCH), kPsize); parents.push_back(Ref{bfirst, bpg}); ++n_branch; };
        for (const Ref& c : level) {
          const bool first_of_page = !open || !b.room(kNodeHdr + c.first_key.size());
          if (first_of_page) {
            if (open) flush();
            b = PageBuilder(); bpg = next_pg++; bfirst = c.first_key; open = true;
          }
          const size_t ks = first_of_page ? 0 : c.first_key.size();      // node 0 of a branch page carries no key
          uint8_t* nd = b.add(kNodeHdr + ks);
          const uint16_t lo = (uint16_t)(c.pgno & 0xffff), hi = (uint16_t)((c.pgno >> 16) & 0xffff), top = (uint16_t)((c.pgno >> 32) & 0xffff), k16 = (uint16_t)ks;
          memcpy(nd, &lo, 2); memcpy(nd + 2, &hi, 2); memcpy(nd + 4, &top, 2); memcpy(nd + 6, &k16, 2);
          if (ks) memcpy(nd + kNodeHdr, c.first_key.data(), ks);
        }
        flush();
        level.swap(parents);
        ++depth;
      }
      root = level[0].pgno;
    }
    ::fdatasync(fd);                                       // data pages before the meta page that publishes them
    put_meta((unsigned)(txnid & 1), txnid, depth, n_branch, n_leaf, n_overflow, entries, root, next_pg - 1);
    ::fdatasync(fd);
  }
};

void LMDB::Commit(std::vector<std::string>& keys, std::vector<std::string>& values) {
  B2_CHECK(mode_ != READ && w_, "db::LMDB: Commit on an environment opened READ");
  // the transaction's puts in key order; a key put twice keeps its last value (mdb_put overwrites)
  std::vector<size_t> order(keys.size());
  for (size_t i = 0; i < order.size(); ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return keys[a] < keys[b]; });
  std::vector<size_t> uniq;
  for (size_t i = 0; i < order.size(); ++i) {
    if (i + 1 < order.size() && keys[order[i + 1]] == keys[order[i]]) continue;
    uniq.push_back(order[i]);
  }
  Writer& w = *w_;
  if (w.fd < 0) { w.open_fd(file_, false); w.txnid = txnid_; }
  const bool appendable = w.indexed && (uniq.empty() || !w.have_last || keys[uniq.front()] > w.last_key);
  if (appendable) {
    for (size_t i : uniq) w.Append(keys[i], values[i]);
    w.FinishCommit();
  } else {
    // merge with what the file holds and rebuild it next to the old one
    std::map<std::string, std::string> all;
    if (stale_ || !map_) Map(file_);
    {
      LMDBCursor cur(this);
      for (; cur.valid(); cur.Next()) all[cur.key()] = cur.value();
    }
    for (size_t i : uniq) all[keys[i]] = values[i];
    const uint64_t next_txn = txnid_ + 1;
    const std::string tmp = file_ + ".tmp";
    w.open_fd(tmp, true);
    w.reset_tree();
    w.put_meta(0, 0, 0, 0, 0, 0, 0, kInvalid, 1);
    w.put_meta(1, 0, 0, 0, 0, 0, 0, kInvalid, 1);
    w.txnid = next_txn - 1;
    for (auto& kv : all) w.Append(kv.first, kv.second);
    w.FinishCommit();
    w.indexed = true;
    Unmap();
    B2_CHECK(::rename(tmp.c_str(), file_.c_str()) == 0, "lmdb: rename " + tmp + " failed: " + std::strerror(errno));
  }
  Map(file_);
}

bool LMDB::Exists(const std::string& source) {
  struct stat st;
  const std::string f = data_file(source);
  return !source.empty() && stat(f.c_str(), &st) == 0 && S_ISREG(st.st_mode);
}

LMDB::LMDB() {}
LMDB::~LMDB() { Close(); }

void LMDB::Map(const std::string& f) {
  Unmap();
  const int fd = ::open(f.c_str(), O_RDONLY);
  B2_CHECK(fd >= 0, "Failed to open lmdb " + f + ": " + std::strerror(errno));
  struct stat st;
  if (fstat(fd, &st) != 0 || st.st_size < (off_t)(2 * 512)) { ::close(fd); Fatal(__FILE__, __LINE__, "lmdb " + f + ": file too small to hold the meta pages"); }
  void* p = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_SHARED, fd, 0);
  ::close(fd);
  B2_CHECK(p != MAP_FAILED, "lmdb " + f + ": mmap failed: " + std::strerror(errno));
  map_ = static_cast<const uint8_t*>(p);
  map_bytes_ = (size_t)st.st_size;
  // meta page 0 sits at offset 0 and tells the page size; meta page 1 at offset psize
  Meta m0 = read_meta(map_);
  if (!m0.ok || m0.psize < 512 || (m0.psize & (m0.psize - 1)) || (size_t)2 * m0.psize > map_bytes_) {
    Unmap();
    Fatal(__FILE__, __LINE__, "lmdb " + f + ": MDB_INVALID: File is not an LMDB file");
  }
  Meta m1 = read_meta(map_ + m0.psize);
  const Meta& m = (m1.ok && m1.txnid > m0.txnid) ? m1 : m0;
  if (m.flags & 0x04 /* MDB_DUPSORT */) { Unmap(); Fatal(__FILE__, __LINE__, "lmdb " + f + ": DUPSORT databases are not built"); }
  psize_ = m0.psize;
  depth_ = m.depth;
  entries_ = (size_t)m.entries;
  root_ = m.root;
  last_pg_ = m.last_pg;
  txnid_ = m.txnid;
  if (root_ != kInvalid && root_ >= map_bytes_ / psize_) { Unmap(); Fatal(__FILE__, __LINE__, "lmdb " + f + ": root page past the end of the file (truncated copy?)"); }
  // no madvise: like liblmdb, default read-ahead.  (MADV_SEQUENTIAL would also drop pages behind the cursor, which costs the
  // second epoch of a database that fits the page cache.)
  stale_ = false;
}

void LMDB::Unmap() {
  if (map_) munmap(const_cast<uint8_t*>(map_), map_bytes_);
  map_ = nullptr; map_bytes_ = 0; psize_ = 0; depth_ = 0; root_ = kInvalid; entries_ = 0; last_pg_ = 0; txnid_ = 0;
}

void LMDB::Open(const std::string& source, Mode mode) {
  Close();
  mode_ = mode;
  if (mode == READ) {
    file_ = data_file(source);
    Map(file_);
    return;
  }
  if (mode == NEW) B2_CHECK(mkdir(source.c_str(), 0744) == 0, "mkdir " + source + " failed");      // db_lmdb.cpp:12-14
  struct stat st;
  B2_CHECK(stat(source.c_str(), &st) == 0 && S_ISDIR(st.st_mode), "Failed to open lmdb " + source + ": not a directory");
  file_ = source + "/data.mdb";
  w_.reset(new Writer());
  if (stat(file_.c_str(), &st) == 0) Map(file_);          // WRITE on an existing environment
  else { w_->CreateEmpty(file_); Map(file_); }             // mdb_env_open creates data.mdb with two empty meta pages
}

void LMDB::Close() {
  Unmap();
  w_.reset();
  file_.clear();
  mode_ = READ;
  stale_ = false;
}

LMDBCursor* LMDB::NewCursor() {
  if (stale_) Map(file_);                                 // a transaction changed the file since it was mapped
  return new LMDBCursor(this);
}

LMDBTransaction* LMDB::NewTransaction() {
  B2_CHECK(mode_ != READ && w_, "db::LMDB: NewTransaction on an environment opened READ (MDB_RDONLY)");
  return new LMDBTransaction(this);
}

void LMDBTransaction::Commit() {
  env_->Commit(keys_, values_);
  keys_.clear();
  values_.clear();
}

const uint8_t* LMDB::page(uint64_t pgno) const {
  B2_CHECK(map_ != nullptr, "lmdb: environment is closed");
  B2_CHECK(psize_ && pgno < map_bytes_ / psize_, "lmdb: MDB_PAGE_NOTFOUND: page " + std::to_string(pgno) + " is outside the file");   // no pgno * psize overflow
  return map_ + pgno * (uint64_t)psize_;
}

// ---------------------------------------------------------------------------------------------------------- cursor
LMDBCursor::LMDBCursor(const LMDB* env) : env_(env) { SeekToFirst(); }

void LMDBCursor::descend_leftmost(uint64_t pgno) {
  for (;;) {
    const uint8_t* pg = env_->page(pgno);
    const uint16_t flags = rd16(pg + 10);
    const unsigned nkeys = num_keys(pg, env_->page_size());
    B2_CHECK(!(flags & P_LEAF2), "lmdb: LEAF2 (fixed-size key) pages are not built");
    B2_CHECK(nkeys > 0 && stack_.size() < 64, "lmdb: MDB_CORRUPTED: empty page or runaway depth");
    stack_.push_back(Level{pgno, 0});
    if (flags & P_LEAF) return;
    B2_CHECK(flags & P_BRANCH, "lmdb: MDB_CORRUPTED: page " + std::to_string(pgno) + " is neither branch nor leaf");
    const unsigned off = rd16(pg + kPageHdr);
    B2_CHECK(off + kNodeHdr <= env_->page_size(), "lmdb: MDB_CORRUPTED: node offset past the page");
    const uint8_t* nd = pg + off;
    pgno = (uint64_t)rd16(nd) | ((uint64_t)rd16(nd + 2) << 16) | ((uint64_t)rd16(nd + 4) << 32);
  }
}

void LMDBCursor::load() const {
  const Level& lv = stack_.back();
  const uint8_t* pg = env_->page(lv.pgno);
  const unsigned psize = env_->page_size();
  B2_CHECK(lv.idx >= 0 && (unsigned)lv.idx < num_keys(pg, psize), "lmdb: MDB_CORRUPTED: node index past the page's keys");
  const unsigned off = rd16(pg + kPageHdr + 2 * lv.idx);
  B2_CHECK(off >= kPageHdr && off + kNodeHdr <= psize, "lmdb: MDB_CORRUPTED: node offset past the page");
  const uint8_t* nd = pg + off;
  const uint32_t dsz = (uint32_t)rd16(nd) | ((uint32_t)rd16(nd + 2) << 16);
  const uint16_t nflags = rd16(nd + 4);
  ksize_ = rd16(nd + 6);
  key_ = nd + kNodeHdr;
  B2_CHECK(!(nflags & (F_SUBDATA | F_DUPDATA)), "lmdb: sub-databases / duplicate data are not built");
  B2_CHECK(off + kNodeHdr + ksize_ <= psize, "lmdb: MDB_CORRUPTED: key runs past the page");
  if (nflags & F_BIGDATA) {
    B2_CHECK(off + kNodeHdr + ksize_ + 8 <= psize, "lmdb: MDB_CORRUPTED: overflow reference runs past the page");
    const uint64_t opg = rd64(key_ + ksize_);
    const uint8_t* ov = env_->page(opg);
    B2_CHECK(rd16(ov + 10) & P_OVERFLOW, "lmdb: MDB_CORRUPTED: page " + std::to_string(opg) + " is not an overflow page");
    const uint32_t npages = rd32(ov + 12);
    B2_CHECK((uint64_t)kPageHdr + dsz <= (uint64_t)npages * psize, "lmdb: MDB_CORRUPTED: value larger than its overflow run");
    B2_CHECK(npages >= 1 && opg + (uint64_t)npages > opg, "lmdb: MDB_CORRUPTED: overflow run length");
    env_->page(opg + npages - 1);                     // the whole run lies inside the file
    data_ = ov + kPageHdr;
  } else {
    B2_CHECK(off + kNodeHdr + ksize_ + dsz <= psize, "lmdb: MDB_CORRUPTED: value runs past the page");
    data_ = key_ + ksize_;
  }
  dsize_ = dsz;
  loaded_ = true;
}

void LMDBCursor::SeekToFirst() {
  stack_.clear();
  valid_ = false;
  if (env_->root_ == kInvalid || env_->entries_ == 0) return;   // MDB_NOTFOUND on an empty database
  descend_leftmost(env_->root_);
  loaded_ = false;
  valid_ = true;
}

void LMDBCursor::Next() {
  if (!valid_) return;                                // mdb_cursor_get(MDB_NEXT) past the end keeps answering MDB_NOTFOUND
  // advance in the leaf; when it is exhausted climb until a parent has a next child, then descend leftmost
  while (!stack_.empty()) {
    Level& lv = stack_.back();
    const uint8_t* pg = env_->page(lv.pgno);
    const int nkeys = (int)num_keys(pg, env_->page_size());
    if (lv.idx + 1 < nkeys) {
      ++lv.idx;
      if (rd16(pg + 10) & P_LEAF) { loaded_ = false; return; }
      const unsigned off = rd16(pg + kPageHdr + 2 * lv.idx);
      B2_CHECK(off + kNodeHdr <= env_->page_size(), "lmdb: MDB_CORRUPTED: node offset past the page");
      const uint8_t* nd = pg + off;
      const uint64_t child = (uint64_t)rd16(nd) | ((uint64_t)rd16(nd + 2) << 16) | ((uint64_t)rd16(nd + 4) << 32);
      descend_leftmost(child);
      loaded_ = false;
      return;
    }
    stack_.pop_back();
  }
  valid_ = false;
}

}  // namespace db
}  // namespace caffe
