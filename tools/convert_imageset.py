#!/usr/bin/env python
"""convert_imageset (the reference's tools/convert_imageset.cpp): a list of image files -> an LMDB of Datums, through the C++
db::LMDB writer of host/lmdb_reader.cpp.  Image reading / resizing / re-encoding is OpenCV's, as in the reference.

  python tools/convert_imageset.py [FLAGS] ROOTFOLDER/ LISTFILE DB_NAME
      LISTFILE: lines of  `subfolder/file.JPEG 7`
      --gray  --shuffle  --resize_width W --resize_height H  --check_size  --encoded  --encode_type {png,jpg}

Same records as the reference's tool: key "%08d_<file name>" by line number (after the optional shuffle), Datum{channels, height,
width, data = B,G,R planes, label} or -- with --encoded -- Datum{data = the image file's bytes, encoded = true, label}; one commit per
1 000 images.  The shuffle uses Python's RNG (the reference's uses Caffe's; neither is reproducible across implementations)."""
import argparse, os, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def match_ext(fn, en):                                         # io.cpp:104-115
    p = fn.rfind(".")
    ext = (fn[p:] if p >= 0 else fn).lower()
    en = en.lower()
    return ext == en or (en == "jpg" and ext == "jpeg")


def image_to_datum(path, label, height, width, is_color, encoding):
    """ReadImageToDatum (io.cpp:117-153); None when the file cannot be read.  `encoding` is "" (raw datum), the --encode_type as
    given ("jpg"), or the file's own extension with its dot (".jpg") when it was guessed -- only the latter can match in matchExt,
    so only then is the original file stored untouched, exactly as in the reference."""
    import cv2
    from caffe_mpi_b200 import data_api
    img = cv2.imread(path, cv2.IMREAD_COLOR if is_color else cv2.IMREAD_GRAYSCALE)
    if img is None:
        print("Could not open or find file " + path, file=sys.stderr)
        return None
    if height > 0 and width > 0:
        img = cv2.resize(img, (width, height))
    if encoding:
        channels = 1 if img.ndim == 2 else img.shape[2]
        if (channels == 3) == is_color and not height and not width and match_ext(path, encoding):
            data = open(path, "rb").read()                       # ReadFileToDatum: the file as it is
        else:
            ok, buf = cv2.imencode("." + encoding.lstrip("."), img)
            if not ok:
                return None
            data = buf.tobytes()
        return data_api.datum_serialize(0, 0, 0, data, label, encoded=True)
    if img.ndim == 2:
        img = img[:, :, None]
    chw = img.transpose(2, 0, 1).copy()                          # CVMatToDatum: [channel][row][column], OpenCV's channel order
    return data_api.datum_serialize(chw.shape[0], chw.shape[1], chw.shape[2], chw.tobytes(), label)


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("root")
    ap.add_argument("listfile")
    ap.add_argument("db")
    ap.add_argument("--gray", action="store_true")
    ap.add_argument("--shuffle", action="store_true")
    ap.add_argument("--backend", default="lmdb")
    ap.add_argument("--resize_width", type=int, default=0)
    ap.add_argument("--resize_height", type=int, default=0)
    ap.add_argument("--check_size", action="store_true")
    ap.add_argument("--encoded", action="store_true")
    ap.add_argument("--encode_type", default="")
    a = ap.parse_args()
    if a.backend != "lmdb":
        sys.exit("convert_imageset: only --backend lmdb is built")
    from caffe_mpi_b200 import data_api
    lines = []
    for ln in open(a.listfile):
        parts = ln.split()
        if len(parts) >= 2:
            lines.append((parts[0], int(parts[1])))
    if a.shuffle:
        print("Shuffling data")
        random.shuffle(lines)
    print("A total of %d images." % len(lines))
    if a.encode_type and not a.encoded:
        print("encode_type specified, assuming encoded=true.")
    env = data_api.LMDB(a.db, "NEW")
    count, data_size = 0, None
    for line_id, (name, label) in enumerate(lines):
        enc = a.encode_type
        if a.encoded and not enc:                               # guess the encoding from the file name
            p = name.rfind(".")
            if p < 0:
                print("Failed to guess the encoding of '%s'" % name, file=sys.stderr)
            enc = name[p:].lower() if p >= 0 else ""
        datum = image_to_datum(os.path.join(a.root, name) if not a.root.endswith("/") else a.root + name, label,
                               max(0, a.resize_height), max(0, a.resize_width), not a.gray, enc)
        if datum is None:
            continue
        if a.check_size and not enc:
            d = data_api.datum_parse(datum)
            if data_size is None:
                data_size = len(d["data"])
            elif len(d["data"]) != data_size:
                sys.exit("Check failed: data.size() == data_size Incorrect data field size %d" % len(d["data"]))
        env.put(("%08d_%s" % (line_id, name)).encode(), datum)
        count += 1
        if count % 1000 == 0:
            env.commit()
            print("Processed %d files." % count, flush=True)
    if count % 1000 != 0:
        env.commit()
        print("Processed %d files." % count)
    env.close()


if __name__ == "__main__":
    main()
