#!/usr/bin/env python
"""compute_image_mean (the reference's tools/compute_image_mean.cpp): the per-pixel mean of every datum of an LMDB, written as the
BlobProto `transform_param { mean_file: ... }` reads, plus the per-channel means `mean_value:` wants.

  python tools/compute_image_mean.py INPUT_DB [OUTPUT_FILE]

Same arithmetic as the reference: a float32 running sum per pixel in database order, divided by the count at the end
(compute_image_mean.cpp:75-99) -- so the file is the one the reference's tool writes for the same database, legacy
num / channels / height / width header included.  Encoded datums are decoded first (DecodeDatumNative)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def image_mean(db_path):
    """(mean float32 [C][H][W], count)"""
    import numpy as np
    from caffe_mpi_b200 import data_api
    env = data_api.LMDB(db_path)
    total, count, shape = None, 0, None
    ok = env.seek_to_first()
    while ok:
        _, value = env.current()
        d = data_api.datum_parse(value)
        if d is None:
            sys.exit("compute_image_mean: record %d does not parse as a Datum" % count)
        if d["encoded"]:
            img = data_api.jpeg_decode(d["data"]).astype(np.float32)
        elif d["data"]:
            img = np.frombuffer(d["data"], np.uint8).astype(np.float32).reshape(d["channels"], d["height"], d["width"])
        else:
            img = np.asarray(d["float_data"], np.float32).reshape(d["channels"], d["height"], d["width"])
        if total is None:
            total, shape = np.zeros(img.shape, np.float32), img.shape
        if img.shape != shape:
            sys.exit("compute_image_mean: Incorrect data field size %s (record %d), the first datum has %s" % (img.shape, count, shape))
        total += img                                           # float32 += float32, element by element, like sum_blob.set_data(i, ... + x)
        count += 1
        if count % 10000 == 0:
            print("Processed %d files." % count, flush=True)
        ok = env.next()
    env.close()
    if not count:
        sys.exit("compute_image_mean: the database is empty")
    return total / np.float32(count), count


def legacy_blobproto(mean):
    """BlobProto{num = 1, channels, height, width, data} (caffe.proto:22-35), the layout WriteProtoToBinaryFile(sum_blob) produces."""
    from caffe_mpi_b200.lmdb_io import _varint
    c, h, w = mean.shape
    data = mean.astype("<f4").tobytes()
    return b"\x08\x01" + b"\x10" + _varint(c) + b"\x18" + _varint(h) + b"\x20" + _varint(w) + b"\x2a" + _varint(len(data)) + data


def main():
    if len(sys.argv) not in (2, 3):
        sys.exit(__doc__)
    mean, count = image_mean(sys.argv[1])
    if count % 10000:
        print("Processed %d files." % count)
    if len(sys.argv) == 3:
        print("Write to %s" % sys.argv[2])
        with open(sys.argv[2], "wb") as f:
            f.write(legacy_blobproto(mean))
    print("Number of channels: %d" % mean.shape[0])
    for c in range(mean.shape[0]):
        print("mean_value channel [%d]: %g" % (c, float(mean[c].astype("float32").sum(dtype="float32") / mean[c].size)))


if __name__ == "__main__":
    main()
