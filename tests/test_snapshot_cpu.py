"""The .caffemodel / .solverstate binary wire format of caffe_mpi_b200/host/proto_wire.cpp against google.protobuf built
from the same message schema (caffe.proto:15-35 BlobShape/BlobProto, LayerParameter 1/2/3/4/7, V1LayerParameter 2/3/4/5/6,
NetParameter 1/2/100, SolverState :303-308): files written by the host layer parse with real protobuf, and files written by
real protobuf in every encoding Blob::FromProto accepts (data, double_data, raw FLOAT / DOUBLE / FLOAT16, legacy 4-D dims,
V1 layers) load through the host layer."""
import ctypes as C

import numpy as np
import pytest

from caffe_mpi_b200 import host_api

pb = pytest.importorskip("google.protobuf")
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory  # noqa: E402

F = descriptor_pb2.FieldDescriptorProto


def _schema():
    fd = descriptor_pb2.FileDescriptorProto(name="b2_caffe_subset.proto", package="b2t", syntax="proto2")

    def msg(name, fields):
        m = fd.message_type.add(name=name)
        for (fname, num, ftype, label, extra) in fields:
            f = m.field.add(name=fname, number=num, type=ftype, label=label)
            if extra == "packed":
                f.options.packed = True
            elif extra:
                f.type_name = ".b2t." + extra
        return m

    OPT, REP = F.LABEL_OPTIONAL, F.LABEL_REPEATED
    msg("BlobShape", [("dim", 1, F.TYPE_INT64, REP, "packed")])
    msg("BlobProto", [("num", 1, F.TYPE_INT32, OPT, None), ("channels", 2, F.TYPE_INT32, OPT, None), ("height", 3, F.TYPE_INT32, OPT, None),
                      ("width", 4, F.TYPE_INT32, OPT, None), ("data", 5, F.TYPE_FLOAT, REP, "packed"), ("diff", 6, F.TYPE_FLOAT, REP, "packed"),
                      ("shape", 7, F.TYPE_MESSAGE, OPT, "BlobShape"), ("double_data", 8, F.TYPE_DOUBLE, REP, "packed"),
                      ("raw_data_type", 10, F.TYPE_INT32, OPT, None), ("raw_data", 12, F.TYPE_BYTES, OPT, None)])
    msg("LayerParameter", [("name", 1, F.TYPE_STRING, OPT, None), ("type", 2, F.TYPE_STRING, OPT, None), ("bottom", 3, F.TYPE_STRING, REP, None),
                           ("top", 4, F.TYPE_STRING, REP, None), ("phase", 10, F.TYPE_INT32, OPT, None), ("blobs", 7, F.TYPE_MESSAGE, REP, "BlobProto")])
    msg("V1LayerParameter", [("bottom", 2, F.TYPE_STRING, REP, None), ("top", 3, F.TYPE_STRING, REP, None), ("name", 4, F.TYPE_STRING, OPT, None),
                             ("type", 5, F.TYPE_INT32, OPT, None), ("blobs", 6, F.TYPE_MESSAGE, REP, "BlobProto")])
    msg("NetParameter", [("name", 1, F.TYPE_STRING, OPT, None), ("layers", 2, F.TYPE_MESSAGE, REP, "V1LayerParameter"),
                         ("force_backward", 5, F.TYPE_BOOL, OPT, None), ("layer", 100, F.TYPE_MESSAGE, REP, "LayerParameter")])
    msg("SolverState", [("iter", 1, F.TYPE_INT32, OPT, None), ("learned_net", 2, F.TYPE_STRING, OPT, None),
                        ("history", 3, F.TYPE_MESSAGE, REP, "BlobProto"), ("current_step", 4, F.TYPE_INT32, OPT, None)])
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    get = lambda n: message_factory.GetMessageClass(pool.FindMessageTypeByName("b2t." + n))
    return {n: get(n) for n in ("BlobProto", "LayerParameter", "NetParameter", "SolverState")}


@pytest.fixture(scope="module")
def schema():
    return _schema()


@pytest.fixture(scope="module")
def L():
    lib = host_api.lib()
    lib.b2h_wire_new.restype = C.c_void_p
    lib.b2h_wire_load.restype = C.c_void_p
    lib.b2h_wire_load.argtypes = [C.c_char_p, C.c_int]
    lib.b2h_wire_destroy.argtypes = [C.c_void_p]
    lib.b2h_wire_add_layer.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
    lib.b2h_wire_add_blob.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_float)]
    lib.b2h_wire_save_model.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int]
    lib.b2h_wire_save_state.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.c_int]
    lib.b2h_wire_num_layers.argtypes = [C.c_void_p]
    lib.b2h_wire_num_history.argtypes = [C.c_void_p]
    lib.b2h_wire_layer.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_int)]
    lib.b2h_wire_blob.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_longlong), C.c_void_p]
    lib.b2h_wire_state.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_char_p, C.c_int, C.c_char_p, C.c_int]
    return lib


def add_blob(L, h, layer, arr):
    arr = np.ascontiguousarray(arr, np.float32)
    shp = (C.c_int * arr.ndim)(*arr.shape)
    assert L.b2h_wire_add_blob(h, layer, arr.ndim, shp, arr.ctypes.data_as(C.POINTER(C.c_float))) == 0


def get_blob(L, h, layer, j):
    nd, shp, cnt = C.c_int(), (C.c_int * 8)(), C.c_longlong()
    assert L.b2h_wire_blob(h, layer, j, C.byref(nd), shp, C.byref(cnt), None) == 0
    out = np.empty(cnt.value, np.float32)
    assert L.b2h_wire_blob(h, layer, j, C.byref(nd), shp, C.byref(cnt), out.ctypes.data_as(C.c_void_p)) == 0
    return tuple(shp[k] for k in range(nd.value)), out


def layer_info(L, h, i):
    nm, ty, nb = C.create_string_buffer(256), C.create_string_buffer(256), C.c_int()
    assert L.b2h_wire_layer(h, i, nm, ty, 256, C.byref(nb)) == 0
    return nm.value.decode(), ty.value.decode(), nb.value


@pytest.mark.parametrize("raw", [1, 0])
def test_written_caffemodel_parses_with_protobuf(L, schema, rng, tmp_path, raw):
    w = rng.standard_normal((64, 3, 7, 7)).astype(np.float32)
    g, b = rng.standard_normal(64).astype(np.float32), rng.standard_normal(64).astype(np.float32)
    fc = rng.standard_normal((10, 300)).astype(np.float32)          # > 127 elements per dim: multi-byte varints
    h = L.b2h_wire_new()
    assert L.b2h_wire_add_layer(h, b"data", b"Data") == 0
    assert L.b2h_wire_add_layer(h, b"conv1", b"Convolution") == 1
    add_blob(L, h, 1, w)
    assert L.b2h_wire_add_layer(h, b"conv1/bn", b"BatchNorm") == 2
    add_blob(L, h, 2, g); add_blob(L, h, 2, b)
    assert L.b2h_wire_add_layer(h, b"fc", b"InnerProduct") == 3
    add_blob(L, h, 3, fc)
    path = tmp_path / "m.caffemodel"
    assert L.b2h_wire_save_model(h, str(path).encode(), b"Resnet50", raw) == 0
    L.b2h_wire_destroy(h)
    net = schema["NetParameter"]()
    net.ParseFromString(path.read_bytes())
    assert net.name == "Resnet50" and [l.name for l in net.layer] == ["data", "conv1", "conv1/bn", "fc"]
    assert [l.type for l in net.layer] == ["Data", "Convolution", "BatchNorm", "InnerProduct"]
    assert len(net.layer[0].blobs) == 0 and len(net.layer[2].blobs) == 2
    for blob, ref in ((net.layer[1].blobs[0], w), (net.layer[2].blobs[0], g), (net.layer[2].blobs[1], b), (net.layer[3].blobs[0], fc)):
        assert tuple(blob.shape.dim) == ref.shape
        if raw:                                  # Blob::ToProto of the reference: raw_data_type FLOAT (= 1) + raw bytes
            assert blob.raw_data_type == 1 and len(blob.data) == 0
            got = np.frombuffer(blob.raw_data, np.float32)
        else:
            got = np.array(blob.data, np.float32)
        assert np.array_equal(got, ref.reshape(-1))


def test_written_solverstate_parses_with_protobuf(L, schema, rng, tmp_path):
    hs = [rng.standard_normal(s).astype(np.float32) for s in ((64, 3, 7, 7), (64,), (1000, 2048))]
    h = L.b2h_wire_new()
    for a in hs:
        add_blob(L, h, -1, a)
    path = tmp_path / "s.solverstate"
    assert L.b2h_wire_save_state(h, str(path).encode(), 450000, 3, b"snap/resnet_iter_450000.caffemodel", 1) == 0
    L.b2h_wire_destroy(h)
    st = schema["SolverState"]()
    st.ParseFromString(path.read_bytes())
    assert (st.iter, st.current_step, st.learned_net) == (450000, 3, "snap/resnet_iter_450000.caffemodel")
    assert len(st.history) == 3
    for blob, ref in zip(st.history, hs):
        assert tuple(blob.shape.dim) == ref.shape and np.array_equal(np.frombuffer(blob.raw_data, np.float32), ref.reshape(-1))


def test_protobuf_written_files_load(L, schema, rng, tmp_path):
    """Every encoding Blob::FromProto accepts (blob.cpp:352-445), plus V1 layers and unknown fields to skip."""
    a = rng.standard_normal((4, 3, 2, 2)).astype(np.float32)
    net = schema["NetParameter"](name="mixed", force_backward=True)
    l0 = net.layer.add(name="packed_data", type="Convolution", bottom=["data"], top=["c"], phase=1)
    b0 = l0.blobs.add(); b0.shape.dim.extend(a.shape); b0.data.extend(a.reshape(-1).tolist())
    l1 = net.layer.add(name="double_data", type="InnerProduct")
    b1 = l1.blobs.add(); b1.shape.dim.extend([4, 12]); b1.double_data.extend(a.reshape(-1).astype(np.float64).tolist())
    l2 = net.layer.add(name="raw_float", type="Scale")
    b2 = l2.blobs.add(); b2.shape.dim.extend([48]); b2.raw_data_type = 1; b2.raw_data = a.tobytes()
    l3 = net.layer.add(name="raw_double", type="Scale")
    b3 = l3.blobs.add(); b3.shape.dim.extend([48]); b3.raw_data_type = 0; b3.raw_data = a.astype(np.float64).tobytes()
    l4 = net.layer.add(name="raw_half", type="Scale")
    b4 = l4.blobs.add(); b4.shape.dim.extend([48]); b4.raw_data_type = 2; b4.raw_data = a.astype(np.float16).tobytes()
    l5 = net.layer.add(name="legacy_dims", type="Convolution")
    b5 = l5.blobs.add(); b5.num, b5.channels, b5.height, b5.width = 4, 3, 2, 2; b5.data.extend(a.reshape(-1).tolist())
    v1 = net.layers.add(name="v1_conv", type=4, bottom=["x"], top=["y"])
    bv = v1.blobs.add(); bv.num, bv.channels, bv.height, bv.width = 1, 1, 4, 12; bv.data.extend(a.reshape(-1).tolist())
    path = tmp_path / "mixed.caffemodel"
    path.write_bytes(net.SerializeToString())
    h = L.b2h_wire_load(str(path).encode(), 0)
    assert h, host_api.lib().b2h_last_error()
    n = L.b2h_wire_num_layers(h)
    got = {layer_info(L, h, i)[0]: (i, layer_info(L, h, i)) for i in range(n)}
    assert set(got) == {"packed_data", "double_data", "raw_float", "raw_double", "raw_half", "legacy_dims", "v1_conv"}
    want_shape = {"packed_data": (4, 3, 2, 2), "double_data": (4, 12), "raw_float": (48,), "raw_double": (48,), "raw_half": (48,),
                  "legacy_dims": (4, 3, 2, 2), "v1_conv": (1, 1, 4, 12)}
    for name, (i, (_, ty, nb)) in got.items():
        assert nb == 1
        shape, data = get_blob(L, h, i, 0)
        assert shape == want_shape[name], name
        ref = a.astype(np.float16).astype(np.float32) if name == "raw_half" else a
        assert np.array_equal(data, ref.reshape(-1)), name
    assert got["packed_data"][1][1] == "Convolution"
    L.b2h_wire_destroy(h)
    st = schema["SolverState"](iter=7, learned_net="x.caffemodel", current_step=2)
    hb = st.history.add(); hb.shape.dim.extend([48]); hb.data.extend(a.reshape(-1).tolist())
    sp = tmp_path / "s.solverstate"
    sp.write_bytes(st.SerializeToString())
    h = L.b2h_wire_load(str(sp).encode(), 1)
    it, cs, ln, nn = C.c_int(), C.c_int(), C.create_string_buffer(256), C.create_string_buffer(256)
    assert L.b2h_wire_state(h, C.byref(it), C.byref(cs), ln, 256, nn, 256) == 0
    assert (it.value, cs.value, ln.value) == (7, 2, b"x.caffemodel") and L.b2h_wire_num_history(h) == 1
    assert np.array_equal(get_blob(L, h, -1, 0)[1], a.reshape(-1))
    L.b2h_wire_destroy(h)


def test_malformed_files_are_rejected(L, tmp_path):
    p = tmp_path / "bad.caffemodel"
    p.write_bytes(b"\xa2\x06\xff\xff\xff\x7f" + b"\x00" * 8)       # layer field whose length runs past the end
    assert not L.b2h_wire_load(str(p).encode(), 0)
    assert b"protobuf" in host_api.lib().b2h_last_error()
    assert not L.b2h_wire_load(str(tmp_path / "missing.caffemodel").encode(), 0)
    assert b"File not found" in host_api.lib().b2h_last_error()
