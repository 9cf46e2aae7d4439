#!/usr/bin/env python
"""Write an LMDB of Caffe Datums without liblmdb (the C++ db::LMDB writer of host/lmdb_reader.cpp) -- the stand-in for the reference's
`convert_imageset` (tools/convert_imageset.cpp) on a machine with neither LMDB nor OpenCV.

  python tools/make_lmdb.py OUT_DIR --random N --shape 3x256x256 [--classes 1000] [--seed 0]     synthetic uint8 images
  python tools/make_lmdb.py OUT_DIR --npz FILE.npz [--images-key x --labels-key y]                  arrays: [n][C][H][W] uint8, [n] int
  python tools/make_lmdb.py OUT_DIR --random N --shape 3x256x256 --mean MEAN.binaryproto            also write the per-pixel mean
                                                                                                   (compute_image_mean's output)
Keys are convert_imageset's ("%08d_<name>"), values are Datum{channels, height, width, data, label}; train with
  data_param { source: "OUT_DIR" backend: LMDB batch_size: ... }   and   python tools/caffe.py train --solver=...
"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    from caffe_mpi_b200 import lmdb_io
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("out")
    ap.add_argument("--random", type=int, default=0)
    ap.add_argument("--shape", default="3x256x256")
    ap.add_argument("--classes", type=int, default=1000)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--npz", default="")
    ap.add_argument("--images-key", default="images")
    ap.add_argument("--labels-key", default="labels")
    ap.add_argument("--mean", default="", help="write the mean image as a BlobProto (1 x C x H x W) to this path")
    a = ap.parse_args()
    if a.npz:
        z = np.load(a.npz)
        imgs, labels = np.ascontiguousarray(z[a.images_key], np.uint8), np.asarray(z[a.labels_key]).astype(np.int64)
        if imgs.ndim != 4 or len(imgs) != len(labels):
            sys.exit("make_lmdb: images must be [n][C][H][W] and labels [n]")
    elif a.random > 0:
        c, h, w = (int(v) for v in a.shape.lower().split("x"))
        rng = np.random.default_rng(a.seed)
        imgs = rng.integers(0, 256, (a.random, c, h, w), dtype=np.uint8)
        labels = rng.integers(0, a.classes, a.random)
    else:
        sys.exit("make_lmdb: give --random N or --npz FILE")
    # the C++ writer (host/lmdb_reader.cpp, db::LMDB Mode NEW + Transaction), filled the way convert_imageset fills a database:
    # ascending "%08d_name" keys, one Commit per 1 000 records -- each commit appends, nothing but the leaf index stays in memory
    from caffe_mpi_b200 import data_api
    env = data_api.LMDB(a.out, "NEW")
    for i in range(len(labels)):
        env.put(lmdb_io.caffe_key(i, "img%d.jpg" % i), lmdb_io.datum_bytes(imgs[i], int(labels[i])))
        if (i + 1) % 1000 == 0:
            env.commit()
    env.commit()
    st = env.stat()
    env.close()
    size = os.path.getsize(os.path.join(a.out, "data.mdb"))
    print("wrote %d datums of %s to %s (%.1f MB, tree depth %d, %d transactions)" % (st["entries"], "x".join(map(str, imgs.shape[1:])), a.out,
                                                                                    size / 1e6, st["depth"], st["txnid"]))
    if a.mean:
        data_api.blobproto_save(a.mean, imgs.mean(axis=0, dtype=np.float64).astype(np.float32)[None])
        print("wrote the mean image to %s" % a.mean)


if __name__ == "__main__":
    main()
