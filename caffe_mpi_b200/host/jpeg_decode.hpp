// jpeg_decode.hpp -- JPEG decoder (baseline, extended sequential and progressive Huffman) for ENCODED datums (Datum.encoded = true: `data` holds the image file's bytes).
//
// Reference: DecodeDatumToCVMatNative / DecodeDatumToCVMat (src/caffe/util/io.cpp:167-190) hand the bytes to cv::imdecode, i.e. to
// OpenCV's libjpeg(-turbo), and DataTransformer / CVMatToDatum (io.cpp:205-230) read the result as [channel][row][column] with
// OpenCV's channel order (B, G, R).  There is no OpenCV in this toolchain's C++ side and no libjpeg headers, so the decoder is
// written out here, following libjpeg's DEFAULT decompression path step for step so that the pixels agree bit for bit with what the
// reference's cv::imdecode returns (tests/test_jpeg_cpu.py holds it against this image's cv2, a libjpeg-turbo build):
//   * Huffman entropy decoding, baseline / extended sequential DCT (SOF0, SOF1) and PROGRESSIVE DCT (SOF2: spectral selection and
//     successive approximation, DC / AC first and refinement scans, end-of-band runs; T.81 Annex G, libjpeg jdphuff.c), 8-bit
//     samples, restart intervals, interleaved and per-component scans (ITU T.81 Annex F);
//   * dequantisation + the "islow" integer inverse DCT (libjpeg jidctint.c: 13-bit constants, two passes, the default method);
//   * "fancy" (triangle-filter) chroma upsampling for 2x1 and 2x2 subsampling with libjpeg's edge rules (jdsample.c: h2v1 / h2v2
//     fancy when the component is wider than two samples, pixel replication otherwise), full-size chroma as is;
//   * YCbCr -> RGB with libjpeg's 16-bit fixed-point tables (jdcolor.c), written in B, G, R order; one-component files stay one
//     channel (IMREAD_UNCHANGED) or are replicated to three (force_color, io.cpp:183).
// Not built (fatal with a message): arithmetic-coded, lossless and hierarchical files (SOF3, SOF5+), 12-bit samples, CMYK / four-component
// files, sampling ratios other than 1x1, 2x1, 2x2 against the luma.  PNG-encoded datums are not decoded either.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace caffe {

struct DecodedImage {
  int channels = 0, height = 0, width = 0;
  std::vector<uint8_t> chw;            // [channel][row][column], channels in OpenCV's order (B, G, R)
};

// true for a file that starts with the JPEG SOI marker
bool LooksLikeJpeg(const void* bytes, size_t n);
// Decodes into `out`; throws caffe::FatalError on unsupported or damaged input (cv::imdecode returns an empty Mat there and the
// reference goes on with "Could not decode datum"; a training run on holes is not something to continue).
void DecodeJpeg(const void* bytes, size_t n, bool force_color, DecodedImage* out);

// PNG (png_decode.cpp): 8-bit gray / RGB / RGBA / gray+alpha and palette images, non-interlaced; OpenCV's channel layouts
// (gray -> 1, RGB -> B,G,R, with alpha -> B,G,R,A; force_color -> always B,G,R).
bool LooksLikePng(const void* bytes, size_t n);
void DecodePng(const void* bytes, size_t n, bool force_color, DecodedImage* out);
// whichever of the two the bytes are; anything else is fatal
void DecodeImage(const void* bytes, size_t n, bool force_color, DecodedImage* out);

}  // namespace caffe
