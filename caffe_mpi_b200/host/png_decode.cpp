// png_decode.cpp -- PNG files in ENCODED datums (convert_imageset --encoded --encode_type png; cv::imdecode in the reference,
// src/caffe/util/io.cpp:167-190).  PNG is lossless, so "the pixels cv::imdecode returns" is simply the image: this decoder inflates
// the IDAT stream (RFC 1950 / 1951: stored, fixed and dynamic Huffman blocks), undoes the five scanline filters (PNG spec 9.2) and
// lays the samples out the way OpenCV does -- gray -> 1 channel, RGB -> B, G, R, RGBA -> B, G, R, A, palette -> B, G, R
// (IMREAD_UNCHANGED), or always three channels with force_color (IMREAD_COLOR).  Chunk CRCs and the zlib Adler-32 are verified.
// Built: 8-bit gray / RGB / RGBA / gray+alpha and 1-, 2-, 4-, 8-bit palette images, non-interlaced -- what cv::imencode writes and
// what photographs come as.  Not built (fatal with a message): Adam7 interlacing, 16-bit samples, sub-byte gray, palette transparency.
#include <cstring>

#include "b2caffe.hpp"
#include "jpeg_decode.hpp"

namespace caffe {
namespace {

[[noreturn]] void bad(const std::string& what) { Fatal(__FILE__, __LINE__, "Could not decode datum: PNG: " + what); }

uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

uint32_t crc32(const uint8_t* p, size_t n) {
  static uint32_t table[256];
  static const bool init = [] {
    for (uint32_t i = 0; i < 256; ++i) { uint32_t c = i; for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1; table[i] = c; }
    return true;
  }();
  (void)init;
  uint32_t c = 0xFFFFFFFFu;
  for (size_t i = 0; i < n; ++i) c = table[(c ^ p[i]) & 0xFF] ^ (c >> 8);
  return c ^ 0xFFFFFFFFu;
}

// ---- inflate (RFC 1951), canonical-Huffman decoding by code-length counts ---------------------------------------------------------------
struct Inflater {
  const uint8_t* p;
  const uint8_t* end;
  uint32_t acc = 0;
  int n = 0;
  std::vector<uint8_t>* out;
  size_t limit;

  int bits(int k) {
    while (n < k) { if (p >= end) bad("truncated compressed data"); acc |= (uint32_t)*p++ << n; n += 8; }
    const int v = (int)(acc & ((1u << k) - 1));
    acc >>= k; n -= k;
    return v;
  }
  struct Table { uint16_t count[16]; uint16_t symbol[288]; };
  static bool build(Table* t, const uint8_t* lengths, int nsym) {
    memset(t->count, 0, sizeof(t->count));
    for (int i = 0; i < nsym; ++i) ++t->count[lengths[i]];
    int left = 1;
    for (int len = 1; len <= 15; ++len) { left <<= 1; left -= t->count[len]; if (left < 0) return false; }     // over-subscribed
    uint16_t offs[16];
    offs[1] = 0;
    for (int len = 1; len < 15; ++len) offs[len + 1] = (uint16_t)(offs[len] + t->count[len]);
    for (int i = 0; i < nsym; ++i) if (lengths[i]) t->symbol[offs[lengths[i]]++] = (uint16_t)i;
    return true;
  }
  int decode(const Table& t) {
    int code = 0, first = 0, index = 0;
    for (int len = 1; len <= 15; ++len) {
      code |= bits(1);
      const int count = t.count[len];
      if (code - count < first) return t.symbol[index + (code - first)];
      index += count; first += count; first <<= 1; code <<= 1;
    }
    bad("corrupt Huffman code in the compressed data");
  }
  void codes(const Table& lit, const Table& dist) {
    static const uint16_t lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    static const uint16_t lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    static const uint16_t dbase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    static const uint16_t dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    for (;;) {
      int sym = decode(lit);
      if (sym < 256) { if (out->size() >= limit) bad("more image data than the header announces"); out->push_back((uint8_t)sym); continue; }
      if (sym == 256) return;
      sym -= 257;
      if (sym >= 29) bad("corrupt length code");
      const int len = lbase[sym] + bits(lext[sym]);
      const int ds = decode(dist);
      if (ds >= 30) bad("corrupt distance code");
      const size_t d = (size_t)dbase[ds] + (size_t)bits(dext[ds]);
      if (d > out->size()) bad("distance reaches before the start of the data");
      if (out->size() + (size_t)len > limit) bad("more image data than the header announces");
      for (int i = 0; i < len; ++i) out->push_back((*out)[out->size() - d]);
    }
  }
  void run() {
    for (;;) {
      const int last = bits(1), type = bits(2);
      if (type == 0) {
        acc = 0; n = 0;                                              // to the byte boundary
        if (end - p < 4) bad("truncated stored block");
        const unsigned len = p[0] | (p[1] << 8), nlen = p[2] | (p[3] << 8);
        p += 4;
        if ((len ^ 0xFFFF) != nlen || (size_t)(end - p) < len) bad("corrupt stored block");
        if (out->size() + len > limit) bad("more image data than the header announces");
        out->insert(out->end(), p, p + len);
        p += len;
      } else if (type == 1) {
        static Table lit, dist;
        static const bool init = [] {
          uint8_t l[288];
          for (int i = 0; i < 144; ++i) l[i] = 8;
          for (int i = 144; i < 256; ++i) l[i] = 9;
          for (int i = 256; i < 280; ++i) l[i] = 7;
          for (int i = 280; i < 288; ++i) l[i] = 8;
          build(&lit, l, 288);
          for (int i = 0; i < 30; ++i) l[i] = 5;
          build(&dist, l, 30);
          return true;
        }();
        (void)init;
        codes(lit, dist);
      } else if (type == 2) {
        const int nlen = bits(5) + 257, ndist = bits(5) + 1, ncode = bits(4) + 4;
        if (nlen > 286 || ndist > 30) bad("corrupt dynamic block header");
        static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
        uint8_t lengths[320];
        memset(lengths, 0, sizeof(lengths));
        for (int i = 0; i < ncode; ++i) lengths[order[i]] = (uint8_t)bits(3);
        Table lencode;
        if (!build(&lencode, lengths, 19)) bad("corrupt code-length code");
        int idx = 0;
        while (idx < nlen + ndist) {
          const int sym = decode(lencode);
          if (sym < 16) { lengths[idx++] = (uint8_t)sym; continue; }
          int prev = 0, rep;
          if (sym == 16) { if (idx == 0) bad("repeat with no previous length"); prev = lengths[idx - 1]; rep = 3 + bits(2); }
          else if (sym == 17) rep = 3 + bits(3);
          else rep = 11 + bits(7);
          if (idx + rep > nlen + ndist) bad("too many code lengths");
          while (rep--) lengths[idx++] = (uint8_t)prev;
        }
        if (lengths[256] == 0) bad("no end-of-block code");
        Table lit, dist;
        if (!build(&lit, lengths, nlen) || !build(&dist, lengths + nlen, ndist)) bad("over-subscribed Huffman code");
        codes(lit, dist);
      } else {
        bad("reserved block type");
      }
      if (last) return;
    }
  }
};

inline int paeth(int a, int b, int c) {
  const int p = a + b - c, pa = p > a ? p - a : a - p, pb = p > b ? p - b : b - p, pc = p > c ? p - c : c - p;
  return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

}  // namespace

bool LooksLikePng(const void* bytes, size_t n) {
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  return n >= 8 && memcmp(bytes, sig, 8) == 0;
}

void DecodePng(const void* bytes, size_t n, bool force_color, DecodedImage* out) {
  const uint8_t* p = static_cast<const uint8_t*>(bytes);
  const uint8_t* end = p + n;
  if (!LooksLikePng(bytes, n)) bad("not a PNG file");
  p += 8;
  uint32_t W = 0, H = 0;
  int depth = 0, ctype = -1, interlace = 0;
  bool have_ihdr = false, have_plte = false, iend = false, have_trns = false;
  uint8_t palette[256][3];
  memset(palette, 0, sizeof(palette));
  std::vector<uint8_t> idat;
  while (!iend) {
    if (end - p < 12) bad("truncated file");
    const uint32_t len = be32(p);
    if ((size_t)(end - p) - 12 < len) bad("chunk runs past the end of the file");
    const uint8_t* type = p + 4;
    const uint8_t* data = p + 8;
    if (crc32(type, 4 + (size_t)len) != be32(data + len)) bad("CRC error");
    if (memcmp(type, "IHDR", 4) == 0) {
      if (len != 13 || have_ihdr) bad("bad IHDR");
      W = be32(data); H = be32(data + 4); depth = data[8]; ctype = data[9]; interlace = data[12];
      if (W == 0 || H == 0 || (uint64_t)W * H > ((uint64_t)1 << 28)) bad("empty or oversized image");
      if (data[10] != 0 || data[11] != 0) bad("unknown compression / filter method");
      have_ihdr = true;
    } else if (memcmp(type, "PLTE", 4) == 0) {
      if (len % 3 || len > 768) bad("bad PLTE");
      for (uint32_t i = 0; i < len / 3; ++i) { palette[i][0] = data[3 * i]; palette[i][1] = data[3 * i + 1]; palette[i][2] = data[3 * i + 2]; }
      have_plte = true;
    } else if (memcmp(type, "tRNS", 4) == 0) {
      have_trns = true;
    } else if (memcmp(type, "IDAT", 4) == 0) {
      if (!have_ihdr) bad("IDAT before IHDR");
      idat.insert(idat.end(), data, data + len);
    } else if (memcmp(type, "IEND", 4) == 0) {
      iend = true;
    } else if (!(type[0] & 0x20)) {
      bad("unknown critical chunk");
    }
    p = data + len + 4;
  }
  if (!have_ihdr || idat.size() < 6) bad("no image data");
  if (interlace) bad("Adam7-interlaced files are not built");
  int samples;
  switch (ctype) {
    case 0: samples = 1; if (depth != 8) bad("only 8-bit gray images are built"); break;
    case 2: samples = 3; if (depth != 8) bad("only 8-bit RGB images are built"); break;
    case 3: samples = 1; if (depth != 1 && depth != 2 && depth != 4 && depth != 8) bad("bad palette bit depth"); if (!have_plte) bad("palette image without PLTE"); break;
    case 4: samples = 2; if (depth != 8) bad("only 8-bit gray+alpha images are built"); break;
    case 6: samples = 4; if (depth != 8) bad("only 8-bit RGBA images are built"); break;
    default: bad("bad colour type");
  }
  if (have_trns && ctype != 6 && ctype != 4) bad("tRNS transparency is not built");
  const size_t bpp = (size_t)std::max(1, samples * depth / 8);
  const size_t rowbytes = ((size_t)W * samples * depth + 7) / 8;
  // zlib wrapper (RFC 1950)
  if ((idat[0] & 0x0F) != 8 || ((idat[0] << 8) | idat[1]) % 31 || (idat[1] & 0x20)) bad("bad zlib header");
  std::vector<uint8_t> raw;
  raw.reserve((rowbytes + 1) * H);
  Inflater inf{idat.data() + 2, idat.data() + idat.size()};
  inf.out = &raw;
  inf.limit = (rowbytes + 1) * (size_t)H;
  inf.run();
  if (raw.size() != (rowbytes + 1) * (size_t)H) bad("less image data than the header announces");
  {
    uint32_t a = 1, b = 0;
    for (uint8_t v : raw) { a = (a + v) % 65521; b = (b + a) % 65521; }
    if (inf.end - inf.p < 4 || be32(inf.p) != ((b << 16) | a)) bad("Adler-32 mismatch");
  }
  // scanline filters (PNG 9.2), in place
  std::vector<uint8_t> zero(rowbytes, 0);
  for (uint32_t y = 0; y < H; ++y) {
    uint8_t* row = raw.data() + (size_t)y * (rowbytes + 1);
    const int ft = row[0];
    uint8_t* cur = row + 1;
    const uint8_t* up = y ? raw.data() + (size_t)(y - 1) * (rowbytes + 1) + 1 : zero.data();   // the (already unfiltered) row above
    switch (ft) {
      case 0: break;
      case 1: for (size_t i = bpp; i < rowbytes; ++i) cur[i] = (uint8_t)(cur[i] + cur[i - bpp]); break;
      case 2: for (size_t i = 0; i < rowbytes; ++i) cur[i] = (uint8_t)(cur[i] + up[i]); break;
      case 3: for (size_t i = 0; i < rowbytes; ++i) cur[i] = (uint8_t)(cur[i] + (((i >= bpp ? cur[i - bpp] : 0) + up[i]) >> 1)); break;
      case 4: for (size_t i = 0; i < rowbytes; ++i) cur[i] = (uint8_t)(cur[i] + paeth(i >= bpp ? cur[i - bpp] : 0, up[i], i >= bpp ? up[i - bpp] : 0)); break;
      default: bad("unknown filter type");
    }
  }
  // OpenCV's layouts
  const bool gray_src = ctype == 0 || ctype == 4;
  const bool alpha_src = ctype == 4 || ctype == 6;
  int C;
  if (force_color) C = 3;                                           // IMREAD_COLOR: always B, G, R
  else if (ctype == 0) C = 1;
  else if (ctype == 4) C = 4;                                       // OpenCV expands gray + alpha to B, G, R, A
  else if (ctype == 6) C = 4;
  else C = 3;
  out->channels = C; out->height = (int)H; out->width = (int)W;
  out->chw.assign((size_t)C * H * W, 0);
  const size_t plane = (size_t)H * W;
  for (uint32_t y = 0; y < H; ++y) {
    const uint8_t* row = raw.data() + (size_t)y * (rowbytes + 1) + 1;
    for (uint32_t x = 0; x < W; ++x) {
      uint8_t r, g, b, a = 255;
      if (ctype == 3) {
        const int idx = depth == 8 ? row[x] : (row[(size_t)x * depth / 8] >> (8 - depth - (x * depth) % 8)) & ((1 << depth) - 1);
        r = palette[idx][0]; g = palette[idx][1]; b = palette[idx][2];
      } else if (gray_src) {
        r = g = b = row[(size_t)x * samples];
        if (alpha_src) a = row[(size_t)x * samples + 1];
      } else {
        r = row[(size_t)x * samples]; g = row[(size_t)x * samples + 1]; b = row[(size_t)x * samples + 2];
        if (alpha_src) a = row[(size_t)x * samples + 3];
      }
      const size_t at = (size_t)y * W + x;
      if (C == 1) { out->chw[at] = r; continue; }
      out->chw[at] = b; out->chw[plane + at] = g; out->chw[2 * plane + at] = r;
      if (C == 4) out->chw[3 * plane + at] = a;
    }
  }
}

void DecodeImage(const void* bytes, size_t n, bool force_color, DecodedImage* out) {
  if (LooksLikeJpeg(bytes, n)) return DecodeJpeg(bytes, n, force_color, out);
  if (LooksLikePng(bytes, n)) return DecodePng(bytes, n, force_color, out);
  Fatal(__FILE__, __LINE__, "Could not decode datum: encoded datum is neither a JPEG nor a PNG file (other encodings are not built)");
}

}  // namespace caffe
