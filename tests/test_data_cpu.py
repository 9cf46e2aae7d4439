"""Input pipeline, host half: db::LMDB cursor, Datum wire format, DataReader's record partition and the
DataTransformer's draws -- all on the CPU.

Checkers, none of them product code:
  * caffe_mpi_b200/lmdb_io.py  -- pure-Python LMDB writer + an independent recursive reader of the same format;
  * google.protobuf            -- a Datum / BlobProto schema built from caffe.proto:15-56 field numbers;
  * CursorManagerOracle below  -- a line-by-line restatement of the reference's CursorManager (src/caffe/data_reader.cpp:206-310)
                                  and of the queue order DataLayer consumes (data_reader.cpp:93-98, data_layer.cpp:116-125,232-256);
  * numpy RandomState(seed)    -- MT19937 seeded like boost::mt19937(seed), the reference's caffe::rng_t.
No real LMDB file exists in this image to pin the format against (no liblmdb, no python lmdb, no *.mdb on disk): the format is
restated from LMDB 0.9's structure definitions in lmdb_io.py's header; "format parity unpinned" is recorded in DESIGN.md."""
import os
import struct

import numpy as np
import pytest

from caffe_mpi_b200 import data_api, lmdb_io


def _items(n, sizes, seed=0):
    rng = np.random.default_rng(seed)
    return [(lmdb_io.caffe_key(i, "x%d" % i), rng.integers(0, 256, int(rng.choice(sizes)), dtype=np.uint8).tobytes()) for i in range(n)]


# ------------------------------------------------------------------------------------------------------------ LMDB format
@pytest.mark.parametrize("n,sizes,kw", [
    (0, [10], {}),                                              # empty main database: SeekToFirst answers "not valid"
    (1, [5], {}),                                               # root is a leaf
    (1, [300000], {}),                                          # single record on an overflow run
    (40, [0, 1, 7, 100], {}),                                   # inline values incl. empty ones
    (200, [2014, 2015, 2016, 2017, 2018, 2030], {}),            # around nodemax (2038 - 8 - keylen): inline / overflow boundary
    (500, [3 * 8 * 8, 3 * 32 * 32 + 12], {}),                   # depth 2
    (3000, [20, 60], {"leaf_fill": 0.2, "branch_fanout": 5}),   # deep tree (depth >= 4)
    (64, [3 * 256 * 256 + 14], {"txnid": 2}),                   # ImageNet-sized datums, live meta on page 0
    (300, [10, 5000], {"psize": 8192, "txnid": 7}),             # other page size, live meta on page 1
    (120, [33, 70000], {"psize": 16384, "subdir": False}),      # MDB_NOSUBDIR form: the path is the data file
])
def test_cursor_walks_what_was_written(tmp_path, n, sizes, kw):
    items = _items(n, sizes, seed=n)
    path = str(tmp_path / "db")
    lmdb_io.write_lmdb(path, items, **kw)
    assert lmdb_io.read_lmdb(path) == sorted(items)             # the two Python halves agree with each other
    env = data_api.LMDB(path)
    st = env.stat()
    assert st["entries"] == n and st["page_size"] == kw.get("psize", 4096) and st["txnid"] == kw.get("txnid", 1)
    got = env.items()
    assert got == sorted(items)
    assert not env.valid()
    assert not env.next()                                       # MDB_NEXT past the end stays "not found"
    if n:
        assert env.seek_to_first() and env.current() == sorted(items)[0]      # DataReader's wrap: SeekToFirst after the end
    if kw.get("branch_fanout"):
        assert st["depth"] >= 4
    env.close()


def test_keys_come_back_in_memcmp_order(tmp_path):
    keys = [b"b", b"a", b"ab", b"a\x00", b"\xff", b"B", b"aa", b"a" * 511]
    lmdb_io.write_lmdb(str(tmp_path / "db"), [(k, k[::-1]) for k in keys])
    got = data_api.LMDB(str(tmp_path / "db")).items()
    assert [k for k, _ in got] == sorted(keys) and all(v == k[::-1] for k, v in got)


def test_open_failures_are_fatal(tmp_path):
    with pytest.raises(data_api.DataError, match="Failed to open lmdb"):
        data_api.LMDB(str(tmp_path / "nowhere"))
    bad = tmp_path / "bad"
    bad.mkdir()
    (bad / "data.mdb").write_bytes(b"\x00" * 8192)
    with pytest.raises(data_api.DataError, match="not an LMDB file"):
        data_api.LMDB(str(bad))
    (bad / "data.mdb").write_bytes(b"\x00" * 100)
    with pytest.raises(data_api.DataError, match="too small"):
        data_api.LMDB(str(bad))
    # a copy cut short: the root page lies beyond the end of the file
    items = _items(50, [5000], seed=3)
    good = str(tmp_path / "good")
    lmdb_io.write_lmdb(good, items)
    blob = open(os.path.join(good, "data.mdb"), "rb").read()
    cut = tmp_path / "cut"
    cut.mkdir()
    (cut / "data.mdb").write_bytes(blob[:len(blob) // 2])
    with pytest.raises(data_api.DataError, match="past the end of the file|outside the file"):
        data_api.LMDB(str(cut)).items()
    assert data_api.lmdb_exists(good) and not data_api.lmdb_exists(str(tmp_path / "nowhere")) and not data_api.lmdb_exists("")


def test_stale_meta_page_is_ignored(tmp_path):
    """Two commits: the meta page with the smaller txnid describes an older tree; the reader must follow the newer one."""
    path = str(tmp_path / "db")
    items = _items(30, [50], seed=9)
    lmdb_io.write_lmdb(path, items, txnid=4)                    # live meta on page 0, page 1 = txnid 3 (empty tree)
    f = os.path.join(path, "data.mdb")
    blob = bytearray(open(f, "rb").read())
    # make the stale meta (page 1) claim a bogus root with a larger entry count but keep its smaller txnid
    struct.pack_into("<QQ", blob, 4096 + 16 + 72 + 32, 999, 5)
    open(f, "wb").write(blob)
    env = data_api.LMDB(path)
    assert env.stat()["entries"] == 30 and env.items() == sorted(items)


# ------------------------------------------------------------------------------------------------------------ Datum
def _datum_schema():
    pytest.importorskip("google.protobuf")
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    F = descriptor_pb2.FieldDescriptorProto
    fd = descriptor_pb2.FileDescriptorProto(name="b2_datum.proto", package="b2d", syntax="proto2")
    m = fd.message_type.add(name="Datum")                                           # caffe.proto:43-56
    for name, num, typ, lab in (("channels", 1, F.TYPE_INT32, F.LABEL_OPTIONAL), ("height", 2, F.TYPE_INT32, F.LABEL_OPTIONAL),
                                ("width", 3, F.TYPE_INT32, F.LABEL_OPTIONAL), ("data", 4, F.TYPE_BYTES, F.LABEL_OPTIONAL),
                                ("label", 5, F.TYPE_INT32, F.LABEL_OPTIONAL), ("float_data", 6, F.TYPE_FLOAT, F.LABEL_REPEATED),
                                ("encoded", 7, F.TYPE_BOOL, F.LABEL_OPTIONAL), ("record_id", 8, F.TYPE_UINT32, F.LABEL_OPTIONAL)):
        m.field.add(name=name, number=num, type=typ, label=lab)
    p = fd.message_type.add(name="PackedDatum")                                     # same message, float_data packed
    for name, num, typ, lab in (("channels", 1, F.TYPE_INT32, F.LABEL_OPTIONAL), ("float_data", 6, F.TYPE_FLOAT, F.LABEL_REPEATED)):
        f = p.field.add(name=name, number=num, type=typ, label=lab)
        if num == 6:
            f.options.packed = True
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return (message_factory.GetMessageClass(pool.FindMessageTypeByName("b2d.Datum")),
            message_factory.GetMessageClass(pool.FindMessageTypeByName("b2d.PackedDatum")))


def test_datum_parse_matches_protobuf():
    Datum, Packed = _datum_schema()
    rng = np.random.default_rng(5)
    for c, h, w, label, enc in ((3, 8, 9, 7, False), (1, 28, 28, 0, False), (3, 4, 4, -3, False), (3, 256, 256, 999, False), (0, 0, 0, 12, True)):
        d = Datum(channels=c, height=h, width=w, label=label, data=rng.integers(0, 256, max(1, c * h * w), dtype=np.uint8).tobytes())
        if enc:
            d.encoded = True
        d.record_id = 4000000000
        got = data_api.datum_parse(d.SerializeToString())
        assert got == dict(channels=c, height=h, width=w, label=label, encoded=enc, record_id=4000000000, data=d.data, float_data=[])
    d = Datum(channels=2, height=1, width=3, label=1, float_data=[0.5, -1.25, 3.0, 1e-3, 7.0, -0.0])
    got = data_api.datum_parse(d.SerializeToString())
    assert got["data"] == b"" and np.array_equal(np.float32(got["float_data"]), np.float32(d.float_data))
    got = data_api.datum_parse(Packed(channels=2, float_data=[1.5, 2.5]).SerializeToString())
    assert got["float_data"] == [1.5, 2.5] and got["channels"] == 2
    # unknown fields are skipped, truncated input is refused (ParseFromArray's false)
    assert data_api.datum_parse(b"\x08\x03" + b"\xfa\x01\x02hi" + b"\x28\x05")["label"] == 5
    assert data_api.datum_parse(d.SerializeToString()[:-3]) is None
    assert data_api.datum_parse(b"\x22\x7fabc") is None


def test_datum_written_here_parses_with_protobuf():
    Datum, _ = _datum_schema()
    img = np.arange(3 * 5 * 7, dtype=np.uint8).reshape(3, 5, 7)
    for blob in (lmdb_io.datum_bytes(img, 42), data_api.datum_serialize(3, 5, 7, img.tobytes(), 42)):
        d = Datum()
        d.ParseFromString(blob)
        assert (d.channels, d.height, d.width, d.label, d.data, d.encoded) == (3, 5, 7, 42, img.tobytes(), False)
    assert lmdb_io.datum_bytes(img, 42) == Datum(channels=3, height=5, width=7, data=img.tobytes(), label=42).SerializeToString()
    blob = data_api.datum_serialize(1, 1, 2, b"", -1, encoded=True, float_data=[1.0, 2.0])
    d = Datum()
    d.ParseFromString(blob)
    assert d.label == -1 and d.encoded and list(d.float_data) == [1.0, 2.0]


def test_mean_file_blobproto_round_trip(tmp_path):
    mean = np.random.default_rng(2).uniform(90, 130, (1, 3, 6, 6)).astype(np.float32)
    for raw in (False, True):                                   # BVLC `data` form (compute_image_mean writes it) and NVCaffe raw form
        p = str(tmp_path / ("mean%d.binaryproto" % raw))
        data_api.blobproto_save(p, mean, raw=raw)
        assert np.array_equal(data_api.blobproto_load(p), mean)
    # legacy 4-D header (num / channels / height / width) as old compute_image_mean builds wrote it
    legacy = b"\x08\x01\x10\x03\x18\x06\x20\x06" + b"\x2a" + lmdb_io._varint(mean.size * 4) + mean.tobytes()
    p = str(tmp_path / "legacy.binaryproto")
    open(p, "wb").write(legacy)
    assert np.array_equal(data_api.blobproto_load(p), mean)


# ------------------------------------------------------------------------------------------------------------ DataReader
class CursorManagerOracle:
    """src/caffe/data_reader.cpp:206-310, statement by statement; the cursor is an index into the key-ordered record list."""

    def __init__(self, n_entries, solver_count, solver_rank, parser_threads, parser_thread_id, batch_size, node_count, node_rank):
        self.n = n_entries
        self.solver_count, self.solver_rank, self.batch_size = solver_count, solver_rank, batch_size
        self.parser_threads, self.parser_thread_id = parser_threads, parser_thread_id
        self.node_count, self.node_rank = node_count, node_rank
        self.rank_cycle = parser_threads * batch_size                                   # :221
        self.full_cycle = self.rank_cycle * solver_count * node_count                   # :222
        self.rec_id = self.rec_end = 0
        self.pos = 0

    def _cursor_next(self):
        self.pos += 1
        if self.pos >= self.n:                                                          # !cursor_->valid() -> SeekToFirst
            self.pos = 0

    def rewind(self):                                                                   # :288-305
        rank_cycle_per_solver = self.parser_threads * self.batch_size
        rank_cycle_per_node = rank_cycle_per_solver * self.solver_count
        rank_cycle_begin = rank_cycle_per_solver * self.solver_rank + rank_cycle_per_node * self.node_rank
        self.rec_id = rank_cycle_begin + self.parser_thread_id * self.batch_size
        self.rec_end = self.rec_id + self.batch_size
        self.pos = 0
        for _ in range(self.rec_id):
            self._cursor_next()

    def next(self):                                                                     # :233-259
        datum_pos, record_id = self.pos, self.rec_id
        old_id = self.rec_id
        self.rec_id += 1
        if self.rec_id == self.rec_end:
            self.rec_id += self.full_cycle - self.batch_size
            self.rec_end += self.full_cycle
        for _ in range(old_id, self.rec_id):
            self._cursor_next()
        return datum_pos, record_id


def oracle_batches(n_entries, n_batches, B, S, s, P, N=1, node=0):
    """[(record positions, record ids)] of the first n_batches batches solver s consumes: a single transformer thread walks the
    parser queues round-robin (data_layer.cpp:116-125), parser thread t feeding queue (ranked_rec * P + t) % P = t."""
    cms = []
    for t in range(P):
        cm = CursorManagerOracle(n_entries, S, s, P, t, B, N, node)
        cm.rewind()
        cms.append(cm)
    out = []
    for n in range(n_batches):
        t = n % P
        recs = [cms[t].next() for _ in range(B)]
        ids = [r for _, r in recs]
        assert all(r // cms[t].full_cycle * P + t == n for r in ids)                    # batch_on_solver, data_reader.cpp:96-97
        items = [None] * B
        for pos, r in recs:
            items[r % B] = (pos, r)                                                     # item_id, data_layer.cpp:256
        out.append(items)
    return out


def _datum_db(tmp_path, n, c=3, h=6, w=5, name="db"):
    rng = np.random.default_rng(n)
    imgs = rng.integers(0, 256, (n, c, h, w), dtype=np.uint8)
    labels = rng.integers(0, 1000, n)
    path = str(tmp_path / name)
    lmdb_io.write_datum_lmdb(path, imgs, labels)
    return path, imgs, labels


@pytest.mark.parametrize("n_entries,B,S,P,N", [
    (64, 4, 1, 1, 1),          # one solver, one parser: the database front to back
    (37, 4, 1, 1, 1),          # wraps in the middle of a batch
    (100, 8, 2, 1, 1),         # two solvers
    (100, 5, 4, 2, 1),         # four solvers x two parser threads
    (53, 3, 2, 3, 2),          # two nodes x two solvers x three parser threads, entries prime to everything
    (10, 4, 8, 1, 1),          # full cycle (32) larger than the database
])
def test_reader_partition_matches_cursor_manager(tmp_path, n_entries, B, S, P, N):
    path, imgs, labels = _datum_db(tmp_path, n_entries)
    n_batches = 7
    seen_first_cycle = []
    for node in range(N):
        for s in range(S):
            rd = data_api.DataReader(path, B, solver_count=S, solver_rank=s, node_count=N, node_rank=node, parser_threads=P)
            assert rd.shape == imgs.shape[1:] and rd.entries == n_entries and rd.full_cycle == P * B * S * N
            want = oracle_batches(n_entries, n_batches, B, S, s, P, N, node)
            for n in range(n_batches):
                data, label, ids, bid = rd.next()
                pos = [p for p, _ in want[n]]
                assert bid == n and rd.first_record(n) == want[n][0][1]
                assert ids.tolist() == [r for _, r in want[n]]
                assert np.array_equal(data, imgs[pos]) and np.array_equal(label, labels[pos].astype(np.float32))
                if n < P:
                    seen_first_cycle += ids.tolist()
            rd.close()
    # one full cycle: every record id 0 .. full_cycle-1 is read by exactly one (node, solver, thread)
    assert sorted(seen_first_cycle) == list(range(P * B * S * N))


def test_reader_keeps_up_with_a_slow_consumer_and_shuts_down(tmp_path):
    path, imgs, labels = _datum_db(tmp_path, 30, c=1, h=4, w=4)
    rd = data_api.DataReader(path, 6, parser_threads=2, depth=3)
    got = [rd.next()[2].tolist() for _ in range(11)]             # more batches than buffers in flight: buffers recycle
    want = oracle_batches(30, 11, 6, 1, 0, 2)
    assert got == [[r for _, r in b] for b in want]
    rd.close()                                                   # joins parser threads blocked on their free queues
    rd2 = data_api.DataReader(path, 6)
    del rd2                                                      # destruction without a single pop


def test_reader_refuses_what_it_cannot_feed(tmp_path):
    img = np.zeros((3, 4, 4), np.uint8)
    enc = str(tmp_path / "enc")
    lmdb_io.write_lmdb(enc, [(lmdb_io.caffe_key(0), lmdb_io.datum_bytes(img, 1, encoded=True))])
    with pytest.raises(data_api.DataError, match="neither a JPEG nor a PNG"):
        data_api.DataReader(enc, 2)                               # `encoded` set, but the bytes are no image file
    mixed = str(tmp_path / "mixed")
    lmdb_io.write_lmdb(mixed, [(lmdb_io.caffe_key(0), lmdb_io.datum_bytes(img, 1)), (lmdb_io.caffe_key(1), lmdb_io.datum_bytes(np.zeros((3, 5, 4), np.uint8), 1))])
    rd = data_api.DataReader(mixed, 2)
    with pytest.raises(data_api.DataError, match="height can't vary"):
        rd.next()
    rd.close()
    empty = str(tmp_path / "empty")
    lmdb_io.write_lmdb(empty, [])
    with pytest.raises(data_api.DataError, match="is empty"):
        data_api.DataReader(empty, 2)
    junk = str(tmp_path / "junk")
    lmdb_io.write_lmdb(junk, [(b"k", b"\x22\x7fabc")])
    with pytest.raises(data_api.DataError, match="failed to parse Datum"):
        data_api.DataReader(junk, 1)
    with pytest.raises(data_api.DataError, match="solver_rank"):
        data_api.DataReader(enc, 2, solver_count=2, solver_rank=2)


# ------------------------------------------------------------------------------------------------------------ draws
def _mt19937(seed, n):
    """n raw 32-bit outputs of MT19937 seeded by init_genrand(seed) -- boost::mt19937(seed), std::mt19937(seed)."""
    return np.random.RandomState(seed).randint(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint64)


def test_mt19937_helper_is_the_textbook_stream():
    # first outputs of mt19937 with the default seed 5489 (the C++ standard's check value is the 10000th: 4123659995)
    assert _mt19937(5489, 1)[0] == 3499211612
    assert _mt19937(5489, 10000)[-1] == 4123659995


@pytest.mark.parametrize("mirror,crop,train", [(True, 24, True), (False, 24, True), (True, 0, True), (True, 24, False), (False, 0, False)])
def test_transform_draws_follow_fill3randoms(mirror, crop, train):
    seed, n, H, W = 1701, 50, 32, 29
    h, w, m = data_api.transform_draws(seed, mirror, crop, train, n, H, W)
    per = (1 if mirror else 0) + (2 if (train and crop) else 0)
    raw = _mt19937(seed, max(1, per * n)).tolist()
    k = 0
    for i in range(n):
        r0 = r1 = r2 = 0
        if mirror:                                               # data_transformer.cpp:130-132
            r0 = (raw[k] + 1) & 0xFFFFFFFF
            k += 1
        if train and crop:                                       # :133-136
            r1, r2 = (raw[k] + 1) & 0xFFFFFFFF, (raw[k + 1] + 1) & 0xFFFFFFFF
            k += 2
        assert m[i] == (1 if mirror and r0 % 2 else 0)           # :187
        if crop and train:
            assert (h[i], w[i]) == (r1 % (H - crop + 1), r2 % (W - crop + 1))        # :224-225
        elif crop:
            assert (h[i], w[i]) == ((H - crop) // 2, (W - crop) // 2)                # :227-228
        else:
            assert (h[i], w[i]) == (0, 0)
    with pytest.raises(data_api.DataError, match="crop_size larger"):
        data_api.transform_draws(1, True, 40, True, 1, 32, 48)


# ------------------------------------------------------------------------------------------------------------ Net shape from the database
def test_net_sizes_the_data_top_from_the_first_datum(tmp_path, monkeypatch):
    """DataLayerSetUp reads one datum to shape top[0] (data_layer.cpp:176-183); the prototxt graph builder does the same when the
    source opens, and falls back to the synthetic defaults when it does not."""
    from caffe_mpi_b200 import host_api
    path, imgs, labels = _datum_db(tmp_path, 12, c=1, h=28, w=30, name="mnist_like")
    proto = ('name: "t" layer { name: "d" type: "Data" top: "data" top: "label" data_param { source: "%s" backend: LMDB batch_size: 4 } %s }\n'
             'layer { name: "ip" type: "InnerProduct" bottom: "data" top: "ip" inner_product_param { num_output: 10 } }\n'
             'layer { name: "loss" type: "SoftmaxWithLoss" bottom: "ip" bottom: "label" top: "loss" }\n')
    monkeypatch.delenv("B2C_DATA", raising=False)
    net = host_api.Net(proto % (path, 'transform_param { scale: 0.00390625 }'), is_text=True)
    assert net.uses_database(0) and net.layers()[0][2] == (4, 1, 28, 30)
    net = host_api.Net(proto % (path, 'transform_param { crop_size: 24 mirror: true }'), is_text=True)
    assert net.layers()[0][2] == (4, 1, 24, 24)
    with pytest.raises(host_api.HostError, match="crop_size larger"):
        host_api.Net(proto % (path, 'transform_param { crop_size: 29 }'), is_text=True)
    # no database at the source: synthetic stand-in with the caller's defaults
    net = host_api.Net(proto % (str(tmp_path / "absent"), 'transform_param { crop_size: 24 }'), is_text=True, default_channels=3)
    assert not net.uses_database(0) and net.layers()[0][2] == (4, 3, 24, 24)
    monkeypatch.setenv("B2C_DATA", "synthetic")
    assert not host_api.Net(proto % (path, ""), is_text=True).uses_database(0)
    monkeypatch.setenv("B2C_DATA", "db")
    with pytest.raises(host_api.HostError, match="Failed to open lmdb"):
        host_api.Net(proto % (str(tmp_path / "absent"), ""), is_text=True)
    monkeypatch.delenv("B2C_DATA")
    (tmp_path / "a_leveldb").mkdir()                             # there, but no data.mdb inside: not silently replaced by synthetic data
    with pytest.raises(host_api.HostError, match="holds no data.mdb"):
        host_api.Net(proto % (str(tmp_path / "a_leveldb"), ""), is_text=True)
    with pytest.raises(host_api.HostError, match="LEVELDB"):
        host_api.Net(proto.replace("backend: LMDB ", "") % (path, ""), is_text=True)


def test_library_exports_every_symbol_b2h_data_h_declares():
    import re
    from caffe_mpi_b200 import host_api
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    txt = re.sub(r"/\*.*?\*/", "", open(os.path.join(root, "include", "b2h_data.h")).read(), flags=re.S)
    syms = sorted(set(re.findall(r"\b(b2h_[a-z0-9_]+)\s*\(", txt)))
    assert len(syms) == 23
    L = host_api.lib()
    assert not [s for s in syms if not hasattr(L, s)]


# ------------------------------------------------------------------------------------------------------------ db::LMDB NEW / WRITE
def test_writer_appends_like_convert_imageset(tmp_path):
    """tools/convert_imageset.cpp: Open(NEW), Put under "%08d_name" keys, Commit every 1 000 records (here: every 37) and once at
    the end.  Every commit is a readable database; the C++ cursor and the independent Python reader agree with what was put."""
    path = str(tmp_path / "new_db")
    items = _items(400, [0, 5, 300, 2014, 2017, 2030, 9000, 70000], seed=21)      # inline, around nodemax, overflow runs
    env = data_api.LMDB(path, "NEW")
    assert env.stat()["entries"] == 0 and not env.valid()
    done = 0
    for i, (k, v) in enumerate(items):
        env.put(k, v)
        if (i + 1) % 37 == 0 or i + 1 == len(items):
            env.commit()
            done = i + 1
            st = env.stat()
            assert st["entries"] == done and st["txnid"] == (i // 37) + 1
            if (i + 1) % 111 == 0:
                assert lmdb_io.read_lmdb(path) == items[:done]                     # the Python reader on a mid-way state
                assert data_api.LMDB(path).items() == items[:done]
    assert env.items() == items and lmdb_io.read_lmdb(path) == items
    assert env.stat()["depth"] >= 2
    env.close()
    with pytest.raises(data_api.DataError, match="mkdir .* failed"):
        data_api.LMDB(path, "NEW")                                                 # db_lmdb.cpp:12-14: the directory must not exist
    rd = data_api.LMDB(path)
    with pytest.raises(data_api.DataError, match="READ"):
        rd.put(b"k", b"v")


def test_writer_merges_unordered_and_overwriting_commits(tmp_path):
    """test_db.cpp TestWrite: Open(WRITE) on an existing database and Put keys it already holds; plus keys that sort before the last
    one, empty commits, a database written by someone else (the Python writer) and reopening."""
    path = str(tmp_path / "db")
    base = {lmdb_io.caffe_key(i): bytes([i]) * (i * 40) for i in range(30)}
    lmdb_io.write_lmdb(path, base.items(), txnid=5)
    env = data_api.LMDB(path, "WRITE")
    assert env.stat()["entries"] == 30
    env.put(lmdb_io.caffe_key(3), b"replaced")                  # overwrite
    env.put(b"0000000", b"sorts first")                         # before every existing key
    env.put(b"zz", b"sorts last")
    env.put(b"zz", b"put twice: the last value stays")
    env.commit()
    want = dict(base)
    want[lmdb_io.caffe_key(3)] = b"replaced"
    want[b"0000000"] = b"sorts first"
    want[b"zz"] = b"put twice: the last value stays"
    assert env.stat()["entries"] == 32 and env.stat()["txnid"] == 6
    assert env.items() == sorted(want.items()) == lmdb_io.read_lmdb(path)
    env.commit()                                                 # an empty transaction is still a transaction
    assert env.stat()["txnid"] == 7 and env.items() == sorted(want.items())
    env.put(b"zzz", b"x" * 5000)                                 # after the rebuild, ascending keys append again
    env.commit()
    want[b"zzz"] = b"x" * 5000
    assert env.items() == sorted(want.items()) == lmdb_io.read_lmdb(path)
    env.close()
    env = data_api.LMDB(path, "WRITE")                           # reopen: this object does not know the leaf index -> merge path
    env.put(b"zzzz", b"tail")
    env.commit()
    want[b"zzzz"] = b"tail"
    assert env.items() == sorted(want.items()) == lmdb_io.read_lmdb(path)
    assert not os.path.exists(os.path.join(path, "data.mdb.tmp"))
    with pytest.raises(data_api.DataError, match="BAD_VALSIZE"):
        env.put(b"", b"empty key")
        env.commit()
    # WRITE in an existing, empty directory creates the environment
    empty = tmp_path / "fresh"
    empty.mkdir()
    e2 = data_api.LMDB(str(empty), "WRITE")
    e2.put(b"a", b"1")
    e2.commit()
    assert data_api.LMDB(str(empty)).items() == [(b"a", b"1")]


def test_reader_trains_from_a_database_the_cpp_writer_made(tmp_path):
    """Datums serialised and written by the C++ side, read back by DataReader: the whole storage path without Python's writer."""
    rng = np.random.default_rng(3)
    imgs = rng.integers(0, 256, (12, 3, 5, 4), dtype=np.uint8)
    labels = rng.integers(0, 9, 12)
    path = str(tmp_path / "cpp_db")
    env = data_api.LMDB(path, "NEW")
    for i in range(12):
        env.put(lmdb_io.caffe_key(i, "img%d" % i), data_api.datum_serialize(3, 5, 4, imgs[i].tobytes(), int(labels[i])))
        if i % 5 == 4:
            env.commit()
    env.commit()
    env.close()
    rd = data_api.DataReader(path, 4)
    for k in range(3):
        data, label, ids, _ = rd.next()
        assert np.array_equal(data, imgs[4 * k:4 * k + 4]) and np.array_equal(label, labels[4 * k:4 * k + 4].astype(np.float32))
    rd.close()
