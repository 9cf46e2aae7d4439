// lmdb_reader.hpp -- LMDB environment, cursor and transaction (storage half of the input pipeline).
//
// Mirrors caffe::db::LMDB / LMDBCursor / LMDBTransaction (reference include/caffe/util/db.hpp:12-52, db_lmdb.hpp:27-106,
// src/caffe/util/db_lmdb.cpp): Open(source, READ | WRITE | NEW), NewCursor(), cursor SeekToFirst / Next / valid / key / value /
// data / size, NewTransaction(), Put, Commit.  There is no liblmdb in the toolchain, so the on-disk format is read and written
// directly:
//
//   data.mdb = array of pages of `psize` bytes (psize = meta.mm_dbs[FREE_DBI].md_pad, 4096 by default).
//   page header, 16 bytes:  pgno u64 | pad u16 | flags u16 | lower u16, upper u16  (overflow pages: pages u32 instead)
//       flags: P_BRANCH 0x01, P_LEAF 0x02, P_OVERFLOW 0x04, P_META 0x08, P_LEAF2 0x20, P_SUBP 0x40
//       node offsets u16[] start at byte 16; number of keys = (lower - 16) / 2
//   pages 0 and 1: meta pages; header, then magic u32 0xBEEFC0DE | version u32 (1) | address u64 | mapsize u64 |
//       MDB_db[2] (free list, main) | last_pg u64 | txnid u64.  The meta with the larger txnid is the current one.
//       MDB_db, 48 bytes: pad u32 | flags u16 | depth u16 | branch_pages u64 | leaf_pages u64 | overflow_pages u64 |
//                         entries u64 | root u64 (~0 = empty tree)
//   node, 8-byte header:  lo u16 | hi u16 | flags u16 | ksize u16 | key bytes | data
//       branch node: child page = lo | hi << 16 | flags << 32;  leaf node: data size = lo | hi << 16
//       leaf flag F_BIGDATA 0x01: the data is a u64 page number of an overflow run holding the value after its 16-byte header
//   (LMDB 0.9.x, mdb.c: MDB_page, MDB_node, MDB_db, MDB_meta; little-endian 64-bit build, which is what Caffe's LMDBs are.)
//
// The whole file is mmap'ed PROT_READ / MAP_SHARED, like mdb_env_open(MDB_RDONLY | MDB_NOLOCK) does; values are returned as
// pointers into the mapping (zero copy), so a parser thread moves a datum's bytes once: page cache -> pinned batch buffer.
//
// Writing (Mode NEW / WRITE; what convert_imageset and the reference's tests do through db::Transaction) is a bulk builder, not
// LMDB's copy-on-write engine: a Commit whose keys all sort after the database's last key -- convert_imageset's "%08d_name" keys,
// committed every 1 000 records -- APPENDS: values go to overflow runs and packed leaf pages as they come, only (first key, page)
// per leaf stays in memory, and each Commit writes fresh branch pages plus the meta page of its transaction (meta page txnid & 1;
// the previous transaction's tree stays intact, its branch pages become unreferenced: ~0.3 % of an ImageNet database at 1 000
// records per commit).  Any other Commit (a key that sorts earlier, an overwrite, WRITE on a database this object did not build)
// merges everything in memory and rewrites the file (data.mdb.tmp, then rename): fine for the small databases tests and tools
// make, not for bulk loads in random key order.  One writer, no concurrent readers, no free-list reuse.
// Not built: named sub-databases, DUPSORT, mdb_del, the lock file (Caffe opens readers with MDB_NOLOCK).
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

namespace caffe { namespace db {

enum Mode { READ, WRITE, NEW };

class LMDB;

class LMDBCursor {
 public:
  explicit LMDBCursor(const LMDB* env);        // positioned on the first record, like the reference's (db_lmdb.hpp:29-33)
  void SeekToFirst();
  void Next();
  bool valid() const { return valid_; }
  std::string key() const { ensure(); return std::string(reinterpret_cast<const char*>(key_), ksize_); }
  std::string value() const { ensure(); return std::string(reinterpret_cast<const char*>(data_), dsize_); }
  const void* data() const { ensure(); return data_; }   // into the mapping; stable while the LMDB object lives
  size_t size() const { ensure(); return dsize_; }
  const void* key_data() const { ensure(); return key_; }
  size_t key_size() const { ensure(); return ksize_; }
 private:
  struct Level { uint64_t pgno; int idx; };
  void descend_leftmost(uint64_t pgno);        // push pages down to the leftmost leaf under pgno
  void load() const;                           // key_ / data_ from the top of the stack
  // The node is decoded on first access, not on Next(): a DataReader that strides over the other solvers' records
  // (CursorManager::next steps full_cycle - batch_size records between batches) then touches only leaf pages, never the
  // overflow pages holding the datums it skips.
  void ensure() const { if (valid_ && !loaded_) load(); }
  mutable bool loaded_ = false;
  const LMDB* env_;
  std::vector<Level> stack_;                   // root .. leaf
  mutable const uint8_t* key_ = nullptr;
  mutable const uint8_t* data_ = nullptr;
  mutable size_t ksize_ = 0, dsize_ = 0;
  bool valid_ = false;
};

class LMDBTransaction {
 public:
  explicit LMDBTransaction(LMDB* env) : env_(env) {}
  void Put(const std::string& key, const std::string& value) { keys_.push_back(key); values_.push_back(value); }   // db_lmdb.cpp:52-55
  void Commit();                                                                                                     // db_lmdb.cpp:57-96
 private:
  LMDB* env_;
  std::vector<std::string> keys_, values_;
};

class LMDB {
 public:
  LMDB();
  ~LMDB();
  LMDB(const LMDB&) = delete;
  LMDB& operator=(const LMDB&) = delete;
  // `source` is the environment directory (holding data.mdb) or, for READ, the data file itself.  NEW creates the directory
  // (mkdir must succeed, db_lmdb.cpp:12-14); WRITE opens an existing environment or creates an empty one in an existing directory.
  // A missing / truncated file or a foreign format is fatal (caffe::FatalError), like MDB_CHECK in the reference.
  void Open(const std::string& source, Mode mode = READ);
  void Close();
  LMDBCursor* NewCursor();
  LMDBTransaction* NewTransaction();
  size_t entries() const { return entries_; }          // MDB_stat.ms_entries of the main database
  unsigned page_size() const { return psize_; }
  unsigned depth() const { return depth_; }
  uint64_t txnid() const { return txnid_; }
  static bool Exists(const std::string& source);       // data.mdb present (directory form) or `source` is a regular file

 private:
  friend class LMDBCursor;
  friend class LMDBTransaction;
  struct Writer;                                       // bulk builder state (lmdb_reader.cpp)
  void Map(const std::string& file);                   // mmap + meta selection
  void Unmap();
  void Commit(std::vector<std::string>& keys, std::vector<std::string>& values);
  const uint8_t* page(uint64_t pgno) const;            // bounds-checked
  std::string file_;
  Mode mode_ = READ;
  std::unique_ptr<Writer> w_;
  bool stale_ = false;                                 // the file changed since it was mapped
  const uint8_t* map_ = nullptr;
  size_t map_bytes_ = 0;
  unsigned psize_ = 0, depth_ = 0;
  uint64_t root_ = ~0ull, last_pg_ = 0, txnid_ = 0;
  size_t entries_ = 0;
};

}  // namespace db
}  // namespace caffe
