// jpeg_decode.cpp -- see jpeg_decode.hpp.
#include "jpeg_decode.hpp"

#include <cstring>

#include "b2caffe.hpp"

namespace caffe {
namespace {

const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

[[noreturn]] void bad(const std::string& what) { Fatal(__FILE__, __LINE__, "Could not decode datum: JPEG: " + what); }

// ---- Huffman tables (T.81 Annex C, decoding procedure of F.2.2.3) ---------------------------------------------------------------
struct Huff {
  bool present = false;
  uint8_t bits[17] = {0};
  uint8_t vals[256] = {0};
  int maxcode[18];
  int valoff[17];
  uint16_t look[512];                       // 9-bit prefix -> (length << 8) | symbol, 0 = longer code
  void build() {
    int code = 0, k = 0;
    memset(look, 0, sizeof(look));
    for (int l = 1; l <= 16; ++l) {
      valoff[l] = k - code;
      if (code + bits[l] > (1 << l)) bad("bad Huffman table");     // more codes of this length than the prefix space has left
      if (bits[l]) {
        if (l <= 9)
          for (int i = 0; i < bits[l]; ++i) {
            const int c = code + i, sym = vals[k + i];
            for (int fill = 0; fill < (1 << (9 - l)); ++fill) look[(c << (9 - l)) | fill] = (uint16_t)((l << 8) | sym);
          }
        k += bits[l];
        code += bits[l];
        maxcode[l] = code - 1;
      } else {
        maxcode[l] = -1;
      }
      code <<= 1;
    }
    maxcode[17] = 0x7fffffff;
    present = true;
  }
};

// ---- entropy-coded segment reader: byte stuffing (FF 00), stops feeding at a marker and pads with zero bits like libjpeg ------------
struct BitReader {
  const uint8_t* p;
  const uint8_t* end;
  uint32_t acc = 0;
  int n = 0;
  void fill() {
    while (n <= 24) {
      uint32_t b = 0;
      if (p < end) {
        if (*p == 0xFF) {
          if (p + 1 < end && p[1] == 0x00) { b = 0xFF; p += 2; }
          else b = 0;                        // a marker: leave it for the scan loop, feed zeros
        } else {
          b = *p++;
        }
      }
      acc |= b << (24 - n);
      n += 8;
    }
  }
  int get(int k) {                           // k in 0..16
    if (k == 0) return 0;
    if (n < k) fill();
    const int v = (int)(acc >> (32 - k));
    acc <<= k;
    n -= k;
    return v;
  }
  int decode(const Huff& h) {
    if (n < 16) fill();
    const uint16_t e = h.look[acc >> 23];
    if (e) {
      const int l = e >> 8;
      acc <<= l;
      n -= l;
      return e & 0xFF;
    }
    for (int l = 10; l <= 16; ++l) {
      const int code = (int)(acc >> (32 - l));
      if (h.maxcode[l] >= 0 && code <= h.maxcode[l]) {
        acc <<= l;
        n -= l;
        const int idx = code + h.valoff[l];
        if (idx < 0 || idx > 255) bad("corrupt Huffman code");
        return h.vals[idx];
      }
    }
    bad("corrupt Huffman code");
  }
  void reset() { acc = 0; n = 0; }
};

inline int extend(int v, int t) { return v < (1 << (t - 1)) ? v - (1 << t) + 1 : v; }       // T.81 F.2.2.1 EXTEND

// ---- jidctint.c "islow": dequantised coefficients -> 8x8 samples (level-shifted, clamped) ------------------------------------------------
constexpr int CONST_BITS = 13, PASS1_BITS = 2;
constexpr int64_t FIX_0_298631336 = 2446, FIX_0_390180644 = 3196, FIX_0_541196100 = 4433, FIX_0_765366865 = 6270, FIX_0_899976223 = 7373,
                  FIX_1_175875602 = 9633, FIX_1_501321110 = 12299, FIX_1_847759065 = 15137, FIX_1_961570560 = 16069, FIX_2_053119869 = 16819,
                  FIX_2_562915447 = 20995, FIX_3_072711026 = 25172;
inline int64_t descale(int64_t x, int n) { return (x + (1 << (n - 1))) >> n; }
inline uint8_t clamp255(int64_t v) { return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); }

// 64-bit temporaries: a VALID file never leaves the 32-bit range libjpeg computes in (so the results are libjpeg's bit for bit); a
// damaged one, whose coefficients can be anything, must not be undefined behaviour.
void idct_islow(const int16_t* coef, const uint16_t* quant, uint8_t* out, int stride) {
  int64_t ws[64];
  for (int c = 0; c < 8; ++c) {                                   // pass 1: columns
    const int64_t in0 = (int64_t)coef[c] * quant[c], in1 = (int64_t)coef[8 + c] * quant[8 + c], in2 = (int64_t)coef[16 + c] * quant[16 + c], in3 = (int64_t)coef[24 + c] * quant[24 + c],
                  in4 = (int64_t)coef[32 + c] * quant[32 + c], in5 = (int64_t)coef[40 + c] * quant[40 + c], in6 = (int64_t)coef[48 + c] * quant[48 + c],
                  in7 = (int64_t)coef[56 + c] * quant[56 + c];
    int64_t z2 = in2, z3 = in6;
    int64_t z1 = (z2 + z3) * FIX_0_541196100;
    int64_t tmp2 = z1 + z3 * (-FIX_1_847759065);
    int64_t tmp3 = z1 + z2 * FIX_0_765366865;
    z2 = in0; z3 = in4;
    int64_t tmp0 = (z2 + z3) * (1 << CONST_BITS);
    int64_t tmp1 = (z2 - z3) * (1 << CONST_BITS);
    const int64_t tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    tmp0 = in7; tmp1 = in5; tmp2 = in3; tmp3 = in1;
    z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
    int64_t z4 = tmp1 + tmp3;
    const int64_t z5 = (z3 + z4) * FIX_1_175875602;
    tmp0 *= FIX_0_298631336; tmp1 *= FIX_2_053119869; tmp2 *= FIX_3_072711026; tmp3 *= FIX_1_501321110;
    z1 *= -FIX_0_899976223; z2 *= -FIX_2_562915447; z3 *= -FIX_1_961570560; z4 *= -FIX_0_390180644;
    z3 += z5; z4 += z5;
    tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
    ws[c] = descale(tmp10 + tmp3, CONST_BITS - PASS1_BITS);       ws[56 + c] = descale(tmp10 - tmp3, CONST_BITS - PASS1_BITS);
    ws[8 + c] = descale(tmp11 + tmp2, CONST_BITS - PASS1_BITS);   ws[48 + c] = descale(tmp11 - tmp2, CONST_BITS - PASS1_BITS);
    ws[16 + c] = descale(tmp12 + tmp1, CONST_BITS - PASS1_BITS);  ws[40 + c] = descale(tmp12 - tmp1, CONST_BITS - PASS1_BITS);
    ws[24 + c] = descale(tmp13 + tmp0, CONST_BITS - PASS1_BITS);  ws[32 + c] = descale(tmp13 - tmp0, CONST_BITS - PASS1_BITS);
  }
  for (int r = 0; r < 8; ++r) {                                   // pass 2: rows
    const int64_t* w = ws + 8 * r;
    int64_t z2 = w[2], z3 = w[6];
    int64_t z1 = (z2 + z3) * FIX_0_541196100;
    int64_t tmp2 = z1 + z3 * (-FIX_1_847759065);
    int64_t tmp3 = z1 + z2 * FIX_0_765366865;
    int64_t tmp0 = (w[0] + w[4]) * (1 << CONST_BITS);
    int64_t tmp1 = (w[0] - w[4]) * (1 << CONST_BITS);
    const int64_t tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    tmp0 = w[7]; tmp1 = w[5]; tmp2 = w[3]; tmp3 = w[1];
    z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
    int64_t z4 = tmp1 + tmp3;
    const int64_t z5 = (z3 + z4) * FIX_1_175875602;
    tmp0 *= FIX_0_298631336; tmp1 *= FIX_2_053119869; tmp2 *= FIX_3_072711026; tmp3 *= FIX_1_501321110;
    z1 *= -FIX_0_899976223; z2 *= -FIX_2_562915447; z3 *= -FIX_1_961570560; z4 *= -FIX_0_390180644;
    z3 += z5; z4 += z5;
    tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
    constexpr int SH = CONST_BITS + PASS1_BITS + 3;
    uint8_t* o = out + (size_t)r * stride;
    o[0] = clamp255(descale(tmp10 + tmp3, SH) + 128);  o[7] = clamp255(descale(tmp10 - tmp3, SH) + 128);
    o[1] = clamp255(descale(tmp11 + tmp2, SH) + 128);  o[6] = clamp255(descale(tmp11 - tmp2, SH) + 128);
    o[2] = clamp255(descale(tmp12 + tmp1, SH) + 128);  o[5] = clamp255(descale(tmp12 - tmp1, SH) + 128);
    o[3] = clamp255(descale(tmp13 + tmp0, SH) + 128);  o[4] = clamp255(descale(tmp13 - tmp0, SH) + 128);
  }
}

struct Component {
  int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0;
  int w = 0, hh = 0;                  // true down-sampled size: ceil(W * h / maxh), ceil(H * v / maxv)
  int bw = 0, bh = 0;                 // blocks allocated (MCU-padded)
  int pred = 0;
  std::vector<uint8_t> plane;         // (bw * 8) x (bh * 8)
  std::vector<int16_t> coef;          // progressive files: every block's 64 coefficients (natural order), built up scan by scan
};

struct Decoder {
  const uint8_t* p;
  const uint8_t* end;
  int W = 0, H = 0, ncomp = 0, maxh = 1, maxv = 1, mcux = 0, mcuy = 0, restart_interval = 0;
  bool have_sof = false, adobe = false, progressive = false;
  int eobrun = 0;
  int adobe_transform = -1;
  uint16_t quant[4][64];
  bool have_quant[4] = {false, false, false, false};
  Huff dc[4], ac[4];
  Component comp[4];

  int u8() { if (p >= end) bad("premature end of file"); return *p++; }
  int u16() { const int a = u8(); return (a << 8) | u8(); }

  void read_dqt(int len) {
    const uint8_t* seg_end = p + len;
    if (seg_end > end) bad("premature end of file");
    while (p < seg_end) {
      const int pq = u8(), prec = pq >> 4, id = pq & 15;
      if (id > 3 || prec > 1) bad("bad quantisation table");
      for (int i = 0; i < 64; ++i) quant[id][kZigzag[i]] = (uint16_t)(prec ? u16() : u8());
      have_quant[id] = true;
    }
  }
  void read_dht(int len) {
    const uint8_t* seg_end = p + len;
    if (seg_end > end) bad("premature end of file");
    while (p < seg_end) {
      const int tc = u8(), cls = tc >> 4, id = tc & 15;
      if (cls > 1 || id > 3) bad("bad Huffman table index");
      Huff& h = cls ? ac[id] : dc[id];
      int total = 0;
      h.bits[0] = 0;
      for (int l = 1; l <= 16; ++l) { h.bits[l] = (uint8_t)u8(); total += h.bits[l]; }
      if (total > 256) bad("bad Huffman table");
      memset(h.vals, 0, sizeof(h.vals));
      for (int i = 0; i < total; ++i) h.vals[i] = (uint8_t)u8();
      h.build();
    }
  }
  void read_sof(int len, bool prog) {
    if (have_sof) bad("more than one frame header");
    progressive = prog;
    if (u8() != 8) bad("only 8-bit samples are built");
    H = u16(); W = u16(); ncomp = u8();
    if (H <= 0 || W <= 0) bad("empty image");
    if ((size_t)W * H > ((size_t)1 << 28)) bad("image too large");
    if (ncomp != 1 && ncomp != 3) bad("only 1- and 3-component files are built (CMYK / YCCK are not)");
    if (len != 6 + 3 * ncomp) bad("bad frame header length");
    for (int i = 0; i < ncomp; ++i) {
      Component& c = comp[i];
      c.id = u8();
      const int hv = u8();
      c.h = hv >> 4; c.v = hv & 15; c.tq = u8();
      if (c.h < 1 || c.h > 4 || c.v < 1 || c.v > 4 || c.tq > 3) bad("bad component parameters");
      maxh = std::max(maxh, c.h); maxv = std::max(maxv, c.v);
    }
    mcux = (W + 8 * maxh - 1) / (8 * maxh);
    mcuy = (H + 8 * maxv - 1) / (8 * maxv);
    for (int i = 0; i < ncomp; ++i) {
      Component& c = comp[i];
      c.w = (W * c.h + maxh - 1) / maxh;
      c.hh = (H * c.v + maxv - 1) / maxv;
      c.bw = mcux * c.h; c.bh = mcuy * c.v;
      c.plane.assign((size_t)c.bw * 8 * c.bh * 8, 0);
      if (progressive) c.coef.assign((size_t)c.bw * c.bh * 64, 0);
    }
    have_sof = true;
  }

  void decode_block(BitReader& br, Component& c, int bx, int by) {
    int16_t coef[64];
    memset(coef, 0, sizeof(coef));
    const Huff& hd = dc[c.td];
    const Huff& ha = ac[c.ta];
    const int t = br.decode(hd);
    if (t > 15) bad("corrupt DC coefficient");      // 8-bit baseline: at most 11
    const int diff = t ? extend(br.get(t), t) : 0;
    c.pred += diff;
    if (c.pred < -32768 || c.pred > 32767) bad("corrupt DC coefficient");        // valid 8-bit data stays within +-2047
    coef[0] = (int16_t)c.pred;
    for (int k = 1; k < 64;) {
      const int rs = br.decode(ha), r = rs >> 4, s = rs & 15;
      if (s == 0) {
        if (r != 15) break;                         // EOB
        k += 16;                                    // ZRL
        continue;
      }
      k += r;
      if (k > 63) bad("corrupt AC coefficient run");
      coef[kZigzag[k]] = (int16_t)extend(br.get(s), s);
      ++k;
    }
    idct_islow(coef, quant[c.tq], c.plane.data() + ((size_t)by * 8 * c.bw + bx) * 8, c.bw * 8);
  }

  void restart(BitReader& br, int* next_rst) {
    br.reset();
    // the marker the bit reader stopped in front of (possibly after fill bytes)
    while (br.p < br.end && *br.p != 0xFF) ++br.p;
    while (br.p + 1 < br.end && br.p[1] == 0xFF) ++br.p;
    if (br.p + 1 >= br.end || br.p[1] != 0xD0 + *next_rst) bad("missing restart marker");
    br.p += 2;
    *next_rst = (*next_rst + 1) & 7;
    for (int i = 0; i < ncomp; ++i) comp[i].pred = 0;
    eobrun = 0;
  }

  // ---- progressive scans (T.81 Annex G; libjpeg jdphuff.c) ----------------------------------------------------------------------------
  void prog_dc(BitReader& br, Component& c, int16_t* blk, int ah, int al) {
    if (ah == 0) {                                    // first pass: the DC difference, scaled by the point transform
      const int t = br.decode(dc[c.td]);
      if (t > 15) bad("corrupt DC coefficient");
      c.pred += t ? extend(br.get(t), t) : 0;
      if (c.pred < -32768 || c.pred > 32767) bad("corrupt DC coefficient");
      blk[0] = (int16_t)(c.pred * (1 << al));
    } else if (br.get(1)) {                           // refinement: one more bit of every DC value
      blk[0] = (int16_t)(blk[0] | (1 << al));
    }
  }
  void prog_ac_first(BitReader& br, const Huff& h, int16_t* blk, int ss, int se, int al) {
    if (eobrun > 0) { --eobrun; return; }
    for (int k = ss; k <= se; ++k) {
      const int rs = br.decode(h), r = rs >> 4, s = rs & 15;
      if (s) {
        k += r;
        if (k > 63) bad("corrupt AC coefficient run");
        blk[kZigzag[k]] = (int16_t)(extend(br.get(s), s) * (1 << al));
      } else if (r == 15) {
        k += 15;                                      // ZRL: sixteen zeros
      } else {                                        // EOBr: this band ends here in the next 2^r + extra - 1 blocks too
        eobrun = 1 << r;
        if (r) eobrun += br.get(r);
        --eobrun;
        break;
      }
    }
  }
  void prog_ac_refine(BitReader& br, const Huff& h, int16_t* blk, int ss, int se, int al) {
    const int p1 = 1 << al, m1 = -(1 << al);
    int k = ss;
    if (eobrun == 0) {
      for (; k <= se; ++k) {
        const int rs = br.decode(h);
        int r = rs >> 4;
        const int s = rs & 15;
        int value = 0;
        if (s) {
          if (s != 1) bad("corrupt AC refinement");
          value = br.get(1) ? p1 : m1;                // a coefficient that becomes non-zero in this pass
        } else if (r != 15) {
          eobrun = 1 << r;
          if (r) eobrun += br.get(r);
          break;                                      // the rest of the block only gets correction bits (below)
        }
        // skip r still-zero coefficients, handing a correction bit to every already non-zero one on the way
        while (k <= se) {
          int16_t& cf = blk[kZigzag[k]];
          if (cf != 0) {
            if (br.get(1) && (cf & p1) == 0) cf = (int16_t)(cf + (cf >= 0 ? p1 : m1));
          } else {
            if (--r < 0) break;
          }
          ++k;
        }
        if (s) {
          if (k > 63) bad("corrupt AC refinement");
          blk[kZigzag[k]] = (int16_t)value;
        }
      }
    }
    if (eobrun > 0) {
      for (; k <= se; ++k) {
        int16_t& cf = blk[kZigzag[k]];
        if (cf != 0 && br.get(1) && (cf & p1) == 0) cf = (int16_t)(cf + (cf >= 0 ? p1 : m1));
      }
      --eobrun;
    }
  }
  void read_sos_progressive(Component** sc, int ns, int ss, int se, int ah, int al) {
    if (ss > se || se > 63 || ah > 13 || al > 13 || (ss == 0 && se != 0) || (ss > 0 && ns != 1)) bad("bad progressive scan parameters");
    for (int i = 0; i < ncomp; ++i) comp[i].pred = 0;
    eobrun = 0;
    BitReader br{p, end};
    int next_rst = 0;
    long done = 0;
    auto one = [&](Component& c, int bx, int by) {
      int16_t* blk = c.coef.data() + ((size_t)by * c.bw + bx) * 64;
      if (ss == 0) prog_dc(br, c, blk, ah, al);
      else if (ah == 0) prog_ac_first(br, ac[c.ta], blk, ss, se, al);
      else prog_ac_refine(br, ac[c.ta], blk, ss, se, al);
    };
    if (ns == 1) {
      Component& c = *sc[0];
      const int nbx = (c.w + 7) / 8, nby = (c.hh + 7) / 8;
      for (int by = 0; by < nby; ++by)
        for (int bx = 0; bx < nbx; ++bx) {
          if (restart_interval && done && done % restart_interval == 0) restart(br, &next_rst);
          one(c, bx, by);
          ++done;
        }
    } else {
      for (int my = 0; my < mcuy; ++my)
        for (int mx = 0; mx < mcux; ++mx) {
          if (restart_interval && done && done % restart_interval == 0) restart(br, &next_rst);
          for (int i = 0; i < ns; ++i)
            for (int by = 0; by < sc[i]->v; ++by)
              for (int bx = 0; bx < sc[i]->h; ++bx) one(*sc[i], mx * sc[i]->h + bx, my * sc[i]->v + by);
          ++done;
        }
    }
    p = br.p;
    while (p < end && *p != 0xFF) ++p;
  }
  void finish_progressive() {                          // all scans are in: dequantise + inverse DCT of every block
    for (int i = 0; i < ncomp; ++i) {
      Component& c = comp[i];
      if (!have_quant[c.tq]) bad("component uses a quantisation table that was not defined");
      for (int by = 0; by < c.bh; ++by)
        for (int bx = 0; bx < c.bw; ++bx)
          idct_islow(c.coef.data() + ((size_t)by * c.bw + bx) * 64, quant[c.tq], c.plane.data() + ((size_t)by * 8 * c.bw + bx) * 8, c.bw * 8);
    }
  }

  void read_sos(int len) {
    if (!have_sof) bad("scan before the frame header");
    const int ns = u8();
    if (ns < 1 || ns > ncomp || len != 4 + 2 * ns) bad("bad scan header");
    Component* sc[4];
    for (int i = 0; i < ns; ++i) {
      const int id = u8(), tt = u8();
      sc[i] = nullptr;
      for (int k = 0; k < ncomp; ++k) if (comp[k].id == id) sc[i] = &comp[k];
      if (!sc[i]) bad("scan names an unknown component");
      sc[i]->td = tt >> 4; sc[i]->ta = tt & 15;
      if (sc[i]->td > 3 || sc[i]->ta > 3) bad("scan uses a Huffman table that was not defined");
    }
    const int ss = u8(), se = u8(), ahal = u8();
    for (int i = 0; i < ns; ++i) {
      const bool need_dc = !progressive || (ss == 0 && (ahal >> 4) == 0), need_ac = !progressive || ss > 0;
      if ((need_dc && !dc[sc[i]->td].present) || (need_ac && !ac[sc[i]->ta].present)) bad("scan uses a Huffman table that was not defined");
      if (!progressive && !have_quant[sc[i]->tq]) bad("component uses a quantisation table that was not defined");
    }
    if (progressive) { read_sos_progressive(sc, ns, ss, se, ahal >> 4, ahal & 15); return; }
    if (ss != 0 || se != 63 || ahal != 0) bad("progressive scan parameters in a sequential file");
    for (int i = 0; i < ncomp; ++i) comp[i].pred = 0;
    BitReader br{p, end};
    int next_rst = 0;
    long done = 0;
    if (ns == 1) {                                   // non-interleaved: the MCU is one block, the scan covers the true component size
      Component& c = *sc[0];
      const int nbx = (c.w + 7) / 8, nby = (c.hh + 7) / 8;
      for (int by = 0; by < nby; ++by)
        for (int bx = 0; bx < nbx; ++bx) {
          if (restart_interval && done && done % restart_interval == 0) restart(br, &next_rst);
          decode_block(br, c, bx, by);
          ++done;
        }
    } else {
      for (int my = 0; my < mcuy; ++my)
        for (int mx = 0; mx < mcux; ++mx) {
          if (restart_interval && done && done % restart_interval == 0) restart(br, &next_rst);
          for (int i = 0; i < ns; ++i)
            for (int by = 0; by < sc[i]->v; ++by)
              for (int bx = 0; bx < sc[i]->h; ++bx) decode_block(br, *sc[i], mx * sc[i]->h + bx, my * sc[i]->v + by);
          ++done;
        }
    }
    // resume marker parsing at the marker that ended the entropy-coded segment
    p = br.p;
    while (p < end && *p != 0xFF) ++p;
  }

  void parse() {
    if (end - p < 4 || p[0] != 0xFF || p[1] != 0xD8) bad("not a JPEG file (no SOI marker)");
    p += 2;
    bool eoi = false;
    int scans = 0;
    while (!eoi) {
      if (p >= end) { if (scans) break; bad("premature end of file"); }       // libjpeg: "premature end", inserts EOI
      if (*p != 0xFF) { ++p; continue; }
      while (p < end && *p == 0xFF) ++p;
      if (p >= end) break;
      const int m = *p++;
      if (m == 0x00 || (m >= 0xD0 && m <= 0xD7) || m == 0x01) continue;       // stuffed byte / stray RSTn / TEM
      if (m == 0xD9) { eoi = true; break; }
      const int len = u16();
      if (len < 2 || p + (len - 2) > end) bad("bad marker length");
      const uint8_t* next = p + (len - 2);
      switch (m) {
        case 0xDB: read_dqt(len - 2); break;
        case 0xC4: read_dht(len - 2); break;
        case 0xC0: case 0xC1: read_sof(len - 2, false); break;
        case 0xC2: read_sof(len - 2, true); break;
        case 0xC3: case 0xC5: case 0xC6: case 0xC7: case 0xC9: case 0xCA: case 0xCB: case 0xCD: case 0xCE: case 0xCF:
          bad("lossless / hierarchical / arithmetic-coded JPEG files are not built");
        case 0xDD: restart_interval = u16(); break;
        case 0xEE:                                                             // APP14 "Adobe": colour transform flag
          if (len - 2 >= 12 && memcmp(p, "Adobe", 5) == 0) { adobe = true; adobe_transform = p[11]; }
          break;
        case 0xDA: read_sos(len - 2); ++scans; next = p; break;
        default: break;                                                        // APPn, COM, DNL, ...: skipped
      }
      p = next;
    }
    if (!have_sof || !scans) bad("no image data");
    if (progressive) finish_progressive();
  }
};

// jdsample.c: one full-resolution row of a component from its down-sampled plane
void upsample_row(const Component& c, int maxh, int maxv, int y, int W, uint8_t* out, std::vector<int>& colsum) {
  const int stride = c.bw * 8;
  const uint8_t* base = c.plane.data();
  const int hr = maxh / c.h, vr = maxv / c.v;
  const int r0 = vr == 2 ? y / 2 : y;
  const uint8_t* in0 = base + (size_t)r0 * stride;
  if (hr == 1 && vr == 1) { memcpy(out, in0, (size_t)W); return; }
  const int w = c.w;
  const bool fancy = w > 2;                          // "do_fancy && downsampled_width > 2"
  if (hr == 2 && vr == 1) {
    if (!fancy) { for (int x = 0; x < W; ++x) out[x] = in0[x / 2]; return; }
    std::vector<uint8_t> tmp((size_t)2 * w);
    tmp[0] = in0[0];
    tmp[1] = (uint8_t)((in0[0] * 3 + in0[1] + 2) >> 2);
    for (int i = 1; i < w - 1; ++i) {
      const int v3 = in0[i] * 3;
      tmp[2 * i] = (uint8_t)((v3 + in0[i - 1] + 1) >> 2);
      tmp[2 * i + 1] = (uint8_t)((v3 + in0[i + 1] + 2) >> 2);
    }
    tmp[2 * (w - 1)] = (uint8_t)((in0[w - 1] * 3 + in0[w - 2] + 1) >> 2);
    tmp[2 * (w - 1) + 1] = in0[w - 1];
    memcpy(out, tmp.data(), (size_t)W);
    return;
  }
  if (hr == 2 && vr == 2) {
    if (!fancy) { for (int x = 0; x < W; ++x) out[x] = in0[x / 2]; return; }
    int r1 = (y & 1) ? r0 + 1 : r0 - 1;               // the nearer neighbour row; the image edge repeats its own row
    if (r1 < 0) r1 = 0;
    if (r1 > c.hh - 1) r1 = c.hh - 1;
    const uint8_t* in1 = base + (size_t)r1 * stride;
    colsum.resize((size_t)w);
    for (int i = 0; i < w; ++i) colsum[i] = in0[i] * 3 + in1[i];
    std::vector<uint8_t> tmp((size_t)2 * w);
    tmp[0] = (uint8_t)((colsum[0] * 4 + 8) >> 4);
    tmp[1] = (uint8_t)((colsum[0] * 3 + colsum[1] + 7) >> 4);
    for (int i = 1; i < w - 1; ++i) {
      tmp[2 * i] = (uint8_t)((colsum[i] * 3 + colsum[i - 1] + 8) >> 4);
      tmp[2 * i + 1] = (uint8_t)((colsum[i] * 3 + colsum[i + 1] + 7) >> 4);
    }
    tmp[2 * (w - 1)] = (uint8_t)((colsum[w - 1] * 3 + colsum[w - 2] + 8) >> 4);
    tmp[2 * (w - 1) + 1] = (uint8_t)((colsum[w - 1] * 4 + 7) >> 4);
    memcpy(out, tmp.data(), (size_t)W);
    return;
  }
  bad("chroma sampling other than 1x1, 2x1 and 2x2 is not built");
}

}  // namespace

bool LooksLikeJpeg(const void* bytes, size_t n) {
  const uint8_t* b = static_cast<const uint8_t*>(bytes);
  return n >= 3 && b[0] == 0xFF && b[1] == 0xD8 && b[2] == 0xFF;
}

void DecodeJpeg(const void* bytes, size_t n, bool force_color, DecodedImage* out) {
  Decoder d;
  d.p = static_cast<const uint8_t*>(bytes);
  d.end = d.p + n;
  d.parse();
  const int W = d.W, H = d.H;
  for (int i = 0; i < d.ncomp; ++i) {
    const Component& c = d.comp[i];
    if (d.maxh % c.h || d.maxv % c.v) bad("fractional sampling ratios are not built");
    const int hr = d.maxh / c.h, vr = d.maxv / c.v;
    if (!((hr == 1 && vr == 1) || (hr == 2 && vr == 1) || (hr == 2 && vr == 2))) bad("chroma sampling other than 1x1, 2x1 and 2x2 is not built");
  }
  out->height = H; out->width = W;
  if (d.ncomp == 1) {
    out->channels = force_color ? 3 : 1;
    out->chw.resize((size_t)out->channels * H * W);
    const Component& c = d.comp[0];
    for (int y = 0; y < H; ++y) memcpy(out->chw.data() + (size_t)y * W, c.plane.data() + (size_t)y * c.bw * 8, (size_t)W);
    for (int k = 1; k < out->channels; ++k) memcpy(out->chw.data() + (size_t)k * H * W, out->chw.data(), (size_t)H * W);
    return;
  }
  out->channels = 3;
  out->chw.resize((size_t)3 * H * W);
  uint8_t* B = out->chw.data();
  uint8_t* G = B + (size_t)H * W;
  uint8_t* R = G + (size_t)H * W;
  const bool rgb = d.adobe && d.adobe_transform == 0;          // Adobe APP14 transform 0: the components are R, G, B already
  // jdcolor.c build_ycc_rgb_table: SCALEBITS = 16, FIX(x) = (int)(x * 65536 + 0.5)
  static int cr_r[256], cb_b[256], cr_g[256], cb_g[256];
  static const bool tables = [] {
    for (int i = 0; i < 256; ++i) {
      const int x = i - 128;
      cr_r[i] = (91881 * x + 32768) >> 16;
      cb_b[i] = (116130 * x + 32768) >> 16;
      cr_g[i] = -46802 * x;
      cb_g[i] = -22554 * x + 32768;
    }
    return true;
  }();
  (void)tables;
  std::vector<uint8_t> row[3] = {std::vector<uint8_t>((size_t)W), std::vector<uint8_t>((size_t)W), std::vector<uint8_t>((size_t)W)};
  std::vector<int> colsum;
  for (int y = 0; y < H; ++y) {
    for (int i = 0; i < 3; ++i) upsample_row(d.comp[i], d.maxh, d.maxv, y, W, row[i].data(), colsum);
    uint8_t* b = B + (size_t)y * W;
    uint8_t* g = G + (size_t)y * W;
    uint8_t* r = R + (size_t)y * W;
    if (rgb) { memcpy(r, row[0].data(), (size_t)W); memcpy(g, row[1].data(), (size_t)W); memcpy(b, row[2].data(), (size_t)W); continue; }
    for (int x = 0; x < W; ++x) {
      const int Y = row[0][x], cb = row[1][x], cr = row[2][x];
      r[x] = clamp255(Y + cr_r[cr]);
      g[x] = clamp255(Y + ((cb_g[cb] + cr_g[cr]) >> 16));
      b[x] = clamp255(Y + cb_b[cb]);
    }
  }
}

}  // namespace caffe
