"""Encoded datums (convert_imageset --encoded): host/jpeg_decode.cpp and host/png_decode.cpp against this image's cv2.imdecode -- OpenCV on libjpeg-turbo,
i.e. what the reference's DecodeDatumToCVMat[Native] (src/caffe/util/io.cpp:167-190) calls.  Bit-exact, over sampling modes
4:4:4 / 4:2:2 / 4:2:0, qualities 10..100, restart intervals, optimised Huffman tables, grayscale files, force_color, odd sizes
down to 1x1, progressive files, and files written by a second encoder (PIL); then a database of encoded datums through DataReader."""
import io

import numpy as np
import pytest

from caffe_mpi_b200 import data_api, lmdb_io

cv2 = pytest.importorskip("cv2")


def _img(rng, h, w, c=3):
    base = rng.integers(0, 256, (h // 8 + 2, w // 8 + 2, c), dtype=np.uint8)
    img = cv2.resize(base, (w, h), interpolation=cv2.INTER_CUBIC).reshape(h, w, c)
    return np.clip(img.astype(int) + rng.integers(-25, 25, (h, w, c)), 0, 255).astype(np.uint8)


def _same(enc, flags=cv2.IMREAD_UNCHANGED, force=False):
    ref = cv2.imdecode(np.frombuffer(enc, np.uint8), flags)
    got = data_api.jpeg_decode(enc, force_color=force)
    if ref.ndim == 2:
        ref = ref[:, :, None]
    return got.shape == (ref.shape[2], ref.shape[0], ref.shape[1]) and np.array_equal(got.transpose(1, 2, 0), ref)


SAMPLING = {"444": "IMWRITE_JPEG_SAMPLING_FACTOR_444", "422": "IMWRITE_JPEG_SAMPLING_FACTOR_422", "420": "IMWRITE_JPEG_SAMPLING_FACTOR_420"}


@pytest.mark.parametrize("sampling", sorted(SAMPLING))
def test_decoder_is_bit_identical_to_imdecode(sampling):
    if not hasattr(cv2, SAMPLING[sampling]):
        pytest.skip("this OpenCV cannot choose the chroma sampling")
    rng = np.random.default_rng(int(sampling))
    sizes = [(1, 1), (1, 17), (17, 1), (2, 2), (3, 5), (8, 8), (16, 16), (15, 17), (33, 31), (64, 48), (100, 3), (5, 200), (256, 256)]
    sizes += [(int(rng.integers(1, 90)), int(rng.integers(1, 90))) for _ in range(40)]
    for h, w in sizes:
        params = [cv2.IMWRITE_JPEG_QUALITY, int(rng.choice([10, 40, 75, 90, 100])), cv2.IMWRITE_JPEG_SAMPLING_FACTOR, getattr(cv2, SAMPLING[sampling])]
        if rng.random() < 0.35:
            params += [cv2.IMWRITE_JPEG_RST_INTERVAL, int(rng.integers(1, 6))]
        if rng.random() < 0.35:
            params += [cv2.IMWRITE_JPEG_OPTIMIZE, 1]
        ok, enc = cv2.imencode(".jpg", _img(rng, h, w), params)
        assert ok and _same(enc.tobytes()), (h, w, params)


def test_grayscale_files_and_force_color():
    rng = np.random.default_rng(9)
    for h, w in [(1, 1), (7, 9), (40, 24), (65, 130)]:
        ok, enc = cv2.imencode(".jpg", _img(rng, h, w, 1), [cv2.IMWRITE_JPEG_QUALITY, 80])
        assert _same(enc.tobytes())                                            # IMREAD_UNCHANGED: one channel (DecodeDatumToCVMatNative)
        assert _same(enc.tobytes(), cv2.IMREAD_COLOR, force=True)              # force_color: three equal channels


def test_files_from_another_encoder():
    PIL = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(4)
    for k in range(30):
        h, w = int(rng.integers(8, 120)), int(rng.integers(8, 120))
        for ss in (0, 1, 2):
            b = io.BytesIO()
            PIL.fromarray(_img(rng, h, w)[:, :, ::-1]).save(b, "JPEG", quality=int(rng.integers(20, 98)), subsampling=ss, optimize=bool(k % 2))
            assert _same(b.getvalue()), (h, w, ss)
    b = io.BytesIO()
    PIL.fromarray(_img(rng, 32, 32, 4), "CMYK").save(b, "JPEG")
    with pytest.raises(data_api.DataError, match="CMYK"):
        data_api.jpeg_decode(b.getvalue())
    with pytest.raises(data_api.DataError, match="not a JPEG|premature|no image"):
        data_api.jpeg_decode(b"\xff\xd8\xff\xd9")
    ok, enc = cv2.imencode(".jpg", _img(rng, 48, 48))
    cut = enc.tobytes()[:len(enc) // 2]                                        # libjpeg pads a truncated scan with zeros and warns; so does this
    assert data_api.jpeg_decode(cut).shape == (3, 48, 48)


def test_progressive_files_are_bit_identical_too():
    """SOF2: spectral selection + successive approximation (DC / AC first and refinement scans, end-of-band runs), as written by
    libjpeg's default progression script (PIL, cv2 IMWRITE_JPEG_PROGRESSIVE): ImageNet's original files, stored as they are by
    `convert_imageset --encoded` without a resize, include them."""
    PIL = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(21)
    for k in range(40):
        h, w = int(rng.integers(1, 100)), int(rng.integers(1, 100))
        for ss in (0, 1, 2):
            b = io.BytesIO()
            PIL.fromarray(_img(rng, h, w)).save(b, "JPEG", quality=int(rng.integers(15, 98)), subsampling=ss, progressive=True, optimize=bool(k % 2))
            assert _same(b.getvalue()), ("PIL progressive", h, w, ss)
        b = io.BytesIO()
        PIL.fromarray(_img(rng, h, w, 1)[:, :, 0]).save(b, "JPEG", progressive=True)
        assert _same(b.getvalue()), ("progressive gray", h, w)
        params = [cv2.IMWRITE_JPEG_PROGRESSIVE, 1, cv2.IMWRITE_JPEG_QUALITY, int(rng.integers(10, 100))]
        if k % 3 == 0:
            params += [cv2.IMWRITE_JPEG_RST_INTERVAL, int(rng.integers(1, 5))]
        ok, enc = cv2.imencode(".jpg", _img(rng, h, w), params)
        assert ok and _same(enc.tobytes()), ("cv2 progressive", h, w, params)


def test_png_files_decode_to_what_imdecode_returns():
    """--encode_type png: lossless, so the only question is OpenCV's layouts -- gray -> 1 channel, RGB -> B,G,R, RGBA and gray+alpha
    -> B,G,R,A, palette -> B,G,R (IMREAD_UNCHANGED); always B,G,R with force_color (IMREAD_COLOR) -- over every zlib compression
    level (stored, fixed and dynamic Huffman blocks), all five scanline filters, sub-byte palettes, sizes down to 1x1."""
    PIL = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(31)
    for t in range(40):
        h, w = int(rng.integers(1, 70)), int(rng.integers(1, 70))
        for c in (1, 3, 4):
            ok, enc = cv2.imencode(".png", _img(rng, h, w, c), [cv2.IMWRITE_PNG_COMPRESSION, int(rng.integers(0, 10))])
            assert ok and _same(enc.tobytes()) and _same(enc.tobytes(), cv2.IMREAD_COLOR, force=True), ("cv2 png", h, w, c)
        pil = PIL.fromarray(_img(rng, h, w))
        b = io.BytesIO()
        pil.convert("P", palette=PIL.ADAPTIVE, colors=int(rng.choice([2, 4, 16, 200]))).save(b, "PNG", optimize=bool(t % 2))
        assert _same(b.getvalue()) and _same(b.getvalue(), cv2.IMREAD_COLOR, force=True), ("palette", h, w)
        b = io.BytesIO()
        PIL.fromarray(_img(rng, h, w, 2), "LA").save(b, "PNG")
        assert _same(b.getvalue()) and _same(b.getvalue(), cv2.IMREAD_COLOR, force=True), ("gray+alpha", h, w)
        b = io.BytesIO()
        pil.save(b, "PNG", compress_level=0)
        assert _same(b.getvalue()), ("stored blocks", h, w)
    ok, enc = cv2.imencode(".png", _img(rng, 20, 20))
    enc = bytearray(enc.tobytes())
    enc[60] ^= 0x10                                                            # one flipped bit in IDAT: the chunk CRC catches it
    with pytest.raises(data_api.DataError, match="CRC|PNG"):
        data_api.jpeg_decode(bytes(enc))
    with pytest.raises(data_api.DataError, match="neither a JPEG nor a PNG"):
        data_api.jpeg_decode(b"GIF89a" + bytes(40))


def test_database_of_encoded_datums_feeds_the_reader(tmp_path):
    """convert_imageset --encoded: Datum{data = the .jpg file's bytes, encoded = true, label} with no shape fields
    (ReadFileToDatum, io.cpp:117-134); the parser threads decode, the batch holds what cv::imdecode + CVMatToDatum would."""
    rng = np.random.default_rng(12)
    files, labels = [], rng.integers(0, 10, 9)
    for i in range(9):
        ok, enc = cv2.imencode(".jpg", _img(rng, 20, 24), [cv2.IMWRITE_JPEG_QUALITY, 85])
        files.append(enc.tobytes())
    path = str(tmp_path / "enc_db")
    env = data_api.LMDB(path, "NEW")
    for i, f in enumerate(files):
        env.put(lmdb_io.caffe_key(i, "img%d.jpg" % i), data_api.datum_serialize(0, 0, 0, f, int(labels[i]), encoded=True))
    env.commit()
    env.close()
    want = np.stack([cv2.imdecode(np.frombuffer(f, np.uint8), cv2.IMREAD_UNCHANGED).transpose(2, 0, 1) for f in files])
    rd = data_api.DataReader(path, 3, parser_threads=2)
    assert rd.shape == (3, 20, 24)
    for k in range(4):                                                          # the fourth batch wraps to the first records
        data, label, ids, _ = rd.next()
        pos = [(3 * k + j) % 9 for j in range(3)]
        assert np.array_equal(data, want[pos]) and np.array_equal(label, labels[pos].astype(np.float32))
    rd.close()
    # the graph builder sizes the Data layer's top from the decoded first datum
    from caffe_mpi_b200 import host_api
    net = host_api.Net('layer { name: "d" type: "Data" top: "data" top: "label" data_param { source: "%s" backend: LMDB batch_size: 3 } '
                       'transform_param { crop_size: 16 } }' % path, is_text=True)
    assert net.uses_database(0) and net.layers()[0][2] == (3, 3, 16, 16)
