// lmdb_reader.cpp -- see lmdb_reader.hpp.
#include "lmdb_reader.hpp"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cerrno>
#include <cstring>

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

bool LMDB::Exists(const std::string& source) {
  struct stat st;
  const std::string f = data_file(source);
  return !source.empty() && stat(f.c_str(), &st) == 0 && S_ISREG(st.st_mode);
}

void LMDB::Open(const std::string& source, Mode mode) {
  B2_CHECK(mode == READ, "db::LMDB: only Mode READ is built (write with the reference's convert_imageset or caffe_mpi_b200.lmdb_io)");
  Close();
  const std::string f = data_file(source);
  const int fd = ::open(f.c_str(), O_RDONLY);
  B2_CHECK(fd >= 0, "Failed to open lmdb " + source + ": " + std::strerror(errno));
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
    Close();
    Fatal(__FILE__, __LINE__, "lmdb " + f + ": MDB_INVALID: File is not an LMDB file");
  }
  Meta m1 = read_meta(map_ + m0.psize);
  const Meta& m = (m1.ok && m1.txnid > m0.txnid) ? m1 : m0;
  if (m.flags & 0x04 /* MDB_DUPSORT */) { Close(); Fatal(__FILE__, __LINE__, "lmdb " + f + ": DUPSORT databases are not built"); }
  psize_ = m0.psize;
  depth_ = m.depth;
  entries_ = (size_t)m.entries;
  root_ = m.root;
  last_pg_ = m.last_pg;
  txnid_ = m.txnid;
  if (root_ != kInvalid && (root_ + 1) * (uint64_t)psize_ > map_bytes_) { Close(); Fatal(__FILE__, __LINE__, "lmdb " + f + ": root page past the end of the file (truncated copy?)"); }
  madvise(const_cast<uint8_t*>(map_), map_bytes_, MADV_SEQUENTIAL);   // Caffe reads in key order, front to back
}

void LMDB::Close() {
  if (map_) munmap(const_cast<uint8_t*>(map_), map_bytes_);
  map_ = nullptr; map_bytes_ = 0; psize_ = 0; depth_ = 0; root_ = kInvalid; entries_ = 0; last_pg_ = 0; txnid_ = 0;
}

const uint8_t* LMDB::page(uint64_t pgno) const {
  B2_CHECK(map_ != nullptr, "lmdb: environment is closed");
  B2_CHECK(pgno != kInvalid && (pgno + 1) * (uint64_t)psize_ <= map_bytes_, "lmdb: MDB_PAGE_NOTFOUND: page " + std::to_string(pgno) + " is outside the file");
  return map_ + pgno * (uint64_t)psize_;
}

// ---------------------------------------------------------------------------------------------------------- cursor
LMDBCursor::LMDBCursor(const LMDB* env) : env_(env) { SeekToFirst(); }

void LMDBCursor::descend_leftmost(uint64_t pgno) {
  for (;;) {
    const uint8_t* pg = env_->page(pgno);
    const uint16_t flags = rd16(pg + 10);
    const unsigned nkeys = (rd16(pg + 12) - kPageHdr) >> 1;
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

void LMDBCursor::load() {
  const Level& lv = stack_.back();
  const uint8_t* pg = env_->page(lv.pgno);
  const unsigned psize = env_->page_size();
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
    env_->page(opg + npages - 1);                     // the whole run lies inside the file
    data_ = ov + kPageHdr;
  } else {
    B2_CHECK(off + kNodeHdr + ksize_ + dsz <= psize, "lmdb: MDB_CORRUPTED: value runs past the page");
    data_ = key_ + ksize_;
  }
  dsize_ = dsz;
}

void LMDBCursor::SeekToFirst() {
  stack_.clear();
  valid_ = false;
  if (env_->root_ == kInvalid || env_->entries_ == 0) return;   // MDB_NOTFOUND on an empty database
  descend_leftmost(env_->root_);
  load();
  valid_ = true;
}

void LMDBCursor::Next() {
  if (!valid_) return;                                // mdb_cursor_get(MDB_NEXT) past the end keeps answering MDB_NOTFOUND
  // advance in the leaf; when it is exhausted climb until a parent has a next child, then descend leftmost
  while (!stack_.empty()) {
    Level& lv = stack_.back();
    const uint8_t* pg = env_->page(lv.pgno);
    const int nkeys = (int)((rd16(pg + 12) - kPageHdr) >> 1);
    if (lv.idx + 1 < nkeys) {
      ++lv.idx;
      if (rd16(pg + 10) & P_LEAF) { load(); return; }
      const unsigned off = rd16(pg + kPageHdr + 2 * lv.idx);
      B2_CHECK(off + kNodeHdr <= env_->page_size(), "lmdb: MDB_CORRUPTED: node offset past the page");
      const uint8_t* nd = pg + off;
      const uint64_t child = (uint64_t)rd16(nd) | ((uint64_t)rd16(nd + 2) << 16) | ((uint64_t)rd16(nd + 4) << 32);
      descend_leftmost(child);
      load();
      return;
    }
    stack_.pop_back();
  }
  valid_ = false;
}

}  // namespace db
}  // namespace caffe
