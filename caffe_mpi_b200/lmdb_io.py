"""LMDB data files and Caffe Datum records in pure Python: a bulk writer, and a reader that restates the page walk.

Why this exists: the input pipeline's storage half (`host/lmdb_reader.cpp`, `host/data_reader.cpp`) reads the databases Caffe's
`convert_imageset` / `create_imagenet.sh` produce (reference `src/caffe/util/db_lmdb.cpp`, `tools/convert_imageset.cpp`), and
there is neither liblmdb nor the python `lmdb` module in this image.  `write_lmdb` produces a `data.mdb` in LMDB 0.9's on-disk
format from (key, value) pairs so that tests and `tools/make_lmdb.py` have databases to read; `read_lmdb` is an independent
(recursive) restatement of the B+tree walk that the tests hold the C++ cursor against.

Format (LMDB 0.9.x `mdb.c`: MDB_page, MDB_node, MDB_db, MDB_meta; little-endian, 64-bit):
  page header 16 B:  pgno u64 | pad u16 | flags u16 | lower u16 | upper u16     (overflow pages: pages u32 in place of lower/upper)
  node header  8 B:  lo u16 | hi u16 | flags u16 | ksize u16 | key | data        (nodes are 2-byte aligned)
  leaf node  : data size = lo | hi << 16; F_BIGDATA: data is the u64 number of the first page of an overflow run
  branch node: child page = lo | hi << 16 | flags << 32; the key of node 0 is empty
  a value goes to an overflow run when 8 + ksize + dsize > nodemax = (((psize - 16) // 2) & ~1) - 2   (2038 for 4 KiB pages)
  meta pages 0 and 1: header | magic 0xBEEFC0DE | version 1 | address u64 | mapsize u64 | MDB_db free | MDB_db main | last_pg | txnid
  MDB_db 48 B: pad u32 (the free DB's pad holds the page size) | flags u16 | depth u16 | branch u64 | leaf u64 | overflow u64 |
               entries u64 | root u64 (2**64 - 1 when empty)
A transaction with id T commits to meta page T & 1; the reader uses the meta with the larger txnid.
"""
import os
import struct

P_BRANCH, P_LEAF, P_OVERFLOW, P_META = 0x01, 0x02, 0x04, 0x08
F_BIGDATA = 0x01
MAGIC, VERSION = 0xBEEFC0DE, 1
PAGEHDR, NODEHDR = 16, 8
INVALID = (1 << 64) - 1
MAXKEY = 511


def _even(n):
    return (n + 1) & ~1


class _Page:
    """A branch or leaf page being filled: node offsets grow up from byte 16, nodes grow down from the end."""

    def __init__(self, pgno, flags, psize):
        self.pgno, self.flags, self.psize = pgno, flags, psize
        self.buf = bytearray(psize)
        self.ptrs = []
        self.upper = psize
        self.first_key = None

    def room(self, node_bytes):
        return self.upper - (PAGEHDR + 2 * len(self.ptrs)) >= _even(node_bytes) + 2

    def used(self):
        return PAGEHDR + 2 * len(self.ptrs) + self.psize - self.upper

    def add(self, node, key):
        if self.first_key is None:
            self.first_key = key
        self.upper -= _even(len(node))
        self.buf[self.upper:self.upper + len(node)] = node
        self.ptrs.append(self.upper)

    def bytes(self):
        lower = PAGEHDR + 2 * len(self.ptrs)
        struct.pack_into("<QHHHH", self.buf, 0, self.pgno, 0, self.flags, lower, self.upper)
        struct.pack_into("<%dH" % len(self.ptrs), self.buf, PAGEHDR, *self.ptrs)
        return bytes(self.buf)


def _db(pad, flags, depth, branch, leaf, overflow, entries, root):
    return struct.pack("<IHHQQQQQ", pad, flags, depth, branch, leaf, overflow, entries, root)


def _meta(pgno, psize, mapsize, main, last_pg, txnid):
    page = bytearray(psize)
    struct.pack_into("<QHHHH", page, 0, pgno, 0, P_META, 0, 0)
    body = struct.pack("<IIQQ", MAGIC, VERSION, 0, mapsize) + _db(psize, 0, 0, 0, 0, 0, 0, INVALID) + main + struct.pack("<QQ", last_pg, txnid)
    page[PAGEHDR:PAGEHDR + len(body)] = body
    return bytes(page)


def write_lmdb(path, items, psize=4096, txnid=1, mapsize=1 << 40, subdir=True, leaf_fill=1.0, branch_fanout=0):
    """Write `items` (iterable of (key: bytes, value: bytes)) as an LMDB environment at `path` (a directory holding data.mdb,
    or the file itself with subdir=False).  Keys are stored in memcmp order, as mdb_put would.  `txnid` is the id of the (single)
    transaction that wrote the data: its meta goes to page txnid & 1, the other meta page holds the empty state before it.
    `leaf_fill` < 1 leaves leaf pages partly empty and `branch_fanout` > 0 caps the children per branch page (both force deeper
    trees out of few records).  Returns the number of pages."""
    items = sorted((bytes(k), bytes(v)) for k, v in items)
    for i in range(1, len(items)):
        if items[i][0] == items[i - 1][0]:
            raise ValueError("duplicate key %r" % items[i][0])
    nodemax = (((psize - PAGEHDR) // 2) & ~1) - 2
    pages = {}
    next_pg = [2]

    def alloc(n=1):
        p = next_pg[0]
        next_pg[0] += n
        return p

    n_overflow = 0
    leaves = []
    cur = None
    budget = int(psize * leaf_fill)
    for key, val in items:
        if not 0 < len(key) <= MAXKEY:
            raise ValueError("key length must be 1..%d" % MAXKEY)
        if NODEHDR + len(key) + len(val) > nodemax:
            npg = (PAGEHDR - 1 + len(val)) // psize + 1
            opg = alloc(npg)
            run = bytearray(npg * psize)
            struct.pack_into("<QHHI", run, 0, opg, 0, P_OVERFLOW, npg)
            run[PAGEHDR:PAGEHDR + len(val)] = val
            for j in range(npg):
                pages[opg + j] = bytes(run[j * psize:(j + 1) * psize])
            n_overflow += npg
            node = struct.pack("<HHHH", len(val) & 0xFFFF, len(val) >> 16, F_BIGDATA, len(key)) + key + struct.pack("<Q", opg)
        else:
            node = struct.pack("<HHHH", len(val) & 0xFFFF, len(val) >> 16, 0, len(key)) + key + val
        if cur is None or not cur.room(len(node)) or (cur.ptrs and cur.used() + _even(len(node)) + 2 > budget):
            cur = _Page(alloc(), P_LEAF, psize)
            leaves.append(cur)
        cur.add(node, key)
    n_leaf, n_branch, depth = len(leaves), 0, 0
    root = INVALID
    level = leaves
    if level:
        depth = 1
        while len(level) > 1:
            parents = []
            cur = None
            for child in level:
                key = child.first_key
                if cur is None or not cur.room(NODEHDR + len(key)) or (branch_fanout and len(cur.ptrs) >= branch_fanout):
                    cur = _Page(alloc(), P_BRANCH, psize)
                    parents.append(cur)
                    key = b""                      # node 0 of a branch page carries no key
                node = struct.pack("<HHHH", child.pgno & 0xFFFF, (child.pgno >> 16) & 0xFFFF, (child.pgno >> 32) & 0xFFFF, len(key)) + key
                cur.add(node, child.first_key)
            n_branch += len(parents)
            for pg in level:
                pages[pg.pgno] = pg.bytes()
            level = parents
            depth += 1
        pages[level[0].pgno] = level[0].bytes()
        root = level[0].pgno
    last_pg = next_pg[0] - 1
    main = _db(0, 0, depth, n_branch, n_leaf, n_overflow, len(items), root)
    empty = _db(0, 0, 0, 0, 0, 0, 0, INVALID)
    live = txnid & 1
    metas = [None, None]
    metas[live] = _meta(live, psize, mapsize, main, last_pg, txnid)
    metas[1 - live] = _meta(1 - live, psize, mapsize, empty, 1, max(txnid - 1, 0))
    fname = os.path.join(path, "data.mdb") if subdir else path
    if subdir:
        os.makedirs(path, exist_ok=True)
    with open(fname, "wb") as f:
        f.write(metas[0])
        f.write(metas[1])
        for p in range(2, next_pg[0]):
            f.write(pages[p])
    return next_pg[0]


def read_lmdb(path):
    """All (key, value) pairs of the main database in key order -- a recursive restatement of the walk, independent of
    host/lmdb_reader.cpp's stack cursor."""
    fname = os.path.join(path, "data.mdb") if os.path.isdir(path) else path
    with open(fname, "rb") as f:
        buf = f.read()

    def meta(off):
        flags = struct.unpack_from("<H", buf, off + 10)[0]
        magic, version = struct.unpack_from("<II", buf, off + PAGEHDR)
        if not flags & P_META or magic != MAGIC or version != VERSION:
            return None
        psize = struct.unpack_from("<I", buf, off + PAGEHDR + 24)[0]
        depth, = struct.unpack_from("<H", buf, off + PAGEHDR + 72 + 6)
        entries, root, last_pg, txnid = struct.unpack_from("<QQQQ", buf, off + PAGEHDR + 72 + 32)
        return dict(psize=psize, depth=depth, entries=entries, root=root, last_pg=last_pg, txnid=txnid)

    m0 = meta(0)
    if m0 is None:
        raise ValueError("not an LMDB file")
    m1 = meta(m0["psize"])
    m = m1 if m1 is not None and m1["txnid"] > m0["txnid"] else m0
    psize = m0["psize"]
    out = []

    def walk(pgno):
        off = pgno * psize
        flags, lower = struct.unpack_from("<HH", buf, off + 10)
        n = (lower - PAGEHDR) // 2
        for i in range(n):
            noff = off + struct.unpack_from("<H", buf, off + PAGEHDR + 2 * i)[0]
            lo, hi, nflags, ksize = struct.unpack_from("<HHHH", buf, noff)
            if flags & P_BRANCH:
                walk(lo | hi << 16 | nflags << 32)
            else:
                key = buf[noff + NODEHDR:noff + NODEHDR + ksize]
                dsz = lo | hi << 16
                if nflags & F_BIGDATA:
                    opg, = struct.unpack_from("<Q", buf, noff + NODEHDR + ksize)
                    val = buf[opg * psize + PAGEHDR:opg * psize + PAGEHDR + dsz]
                else:
                    val = buf[noff + NODEHDR + ksize:noff + NODEHDR + ksize + dsz]
                out.append((key, val))

    if m["root"] != INVALID:
        walk(m["root"])
    assert len(out) == m["entries"], (len(out), m["entries"])
    return out


# ---- Datum (caffe.proto:43-56) ------------------------------------------------------------------------------------------------
def _varint(v):
    v &= (1 << 64) - 1
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def datum_bytes(array_u8, label, encoded=False):
    """Serialized Datum of a [C][H][W] uint8 array, field order as protobuf writes it (channels, height, width, data, label)."""
    c, h, w = array_u8.shape
    data = array_u8.tobytes()
    out = b"\x08" + _varint(c) + b"\x10" + _varint(h) + b"\x18" + _varint(w) + b"\x22" + _varint(len(data)) + data + b"\x28" + _varint(label)
    if encoded:
        out += b"\x38\x01"
    return out


def caffe_key(index, name=""):
    """convert_imageset's key: "%08d_" + file name (tools/convert_imageset.cpp)."""
    return ("%08d_%s" % (index, name)).encode()


def write_datum_lmdb(path, images_u8, labels, **kw):
    """images_u8: [n][C][H][W] uint8; one Datum per image under convert_imageset's keys."""
    return write_lmdb(path, ((caffe_key(i, "img%d.jpg" % i), datum_bytes(images_u8[i], int(labels[i]))) for i in range(len(labels))), **kw)
