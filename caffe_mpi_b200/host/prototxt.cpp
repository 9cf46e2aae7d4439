// prototxt.cpp -- text-format protobuf reader + Net graph builder (see prototxt.hpp for the reference map).
#include "prototxt.hpp"
#include "data_reader.hpp"

#include <cctype>
#include <cmath>
#include <algorithm>
#include <cstdlib>
#include <fstream>
#include <set>
#include <sstream>

namespace caffe {

// ================================================================================================ tokenizer
namespace {
struct Tok { enum Kind { IDENT, STRING, NUMBER, COLON, LBRACE, RBRACE, END } kind; std::string text; int line; };

class Lexer {
 public:
  explicit Lexer(const std::string& s) : s_(s) {}
  Tok next() {
    skip();
    if (i_ >= s_.size()) return {Tok::END, "", line_};
    const char c = s_[i_];
    if (c == ':') { ++i_; return {Tok::COLON, ":", line_}; }
    if (c == '{' || c == '<') { ++i_; return {Tok::LBRACE, "{", line_}; }
    if (c == '}' || c == '>') { ++i_; return {Tok::RBRACE, "}", line_}; }
    if (c == '"' || c == '\'') {
      const char q = c;
      std::string out;
      ++i_;
      while (i_ < s_.size() && s_[i_] != q) {
        if (s_[i_] == '\\' && i_ + 1 < s_.size()) {
          const char e = s_[++i_];
          out += e == 'n' ? '\n' : e == 't' ? '\t' : e;
        } else {
          if (s_[i_] == '\n') ++line_;
          out += s_[i_];
        }
        ++i_;
      }
      if (i_ >= s_.size()) Fatal(__FILE__, __LINE__, "prototxt line " + std::to_string(line_) + ": unterminated string");
      ++i_;
      return {Tok::STRING, out, line_};
    }
    if (std::isalpha((unsigned char)c) || c == '_') {
      size_t b = i_;
      while (i_ < s_.size() && (std::isalnum((unsigned char)s_[i_]) || s_[i_] == '_' || s_[i_] == '.')) ++i_;
      return {Tok::IDENT, s_.substr(b, i_ - b), line_};
    }
    if (std::isdigit((unsigned char)c) || c == '-' || c == '+' || c == '.') {
      size_t b = i_;
      ++i_;
      while (i_ < s_.size() && (std::isalnum((unsigned char)s_[i_]) || s_[i_] == '.' || s_[i_] == '-' || s_[i_] == '+')) {
        // allow exponents like 1e-4; stop a sign that is not part of an exponent
        if ((s_[i_] == '-' || s_[i_] == '+') && !(s_[i_ - 1] == 'e' || s_[i_ - 1] == 'E')) break;
        ++i_;
      }
      return {Tok::NUMBER, s_.substr(b, i_ - b), line_};
    }
    Fatal(__FILE__, __LINE__, "prototxt line " + std::to_string(line_) + ": unexpected character '" + std::string(1, c) + "'");
  }
 private:
  void skip() {
    for (;;) {
      while (i_ < s_.size() && (std::isspace((unsigned char)s_[i_]) || s_[i_] == ',' || s_[i_] == ';')) { if (s_[i_] == '\n') ++line_; ++i_; }
      if (i_ < s_.size() && s_[i_] == '#') { while (i_ < s_.size() && s_[i_] != '\n') ++i_; continue; }
      break;
    }
  }
  const std::string& s_;
  size_t i_ = 0;
  int line_ = 1;
};

void parse_message(Lexer& lx, PMessage& m, bool top) {
  for (;;) {
    Tok t = lx.next();
    if (t.kind == Tok::END) { if (!top) Fatal(__FILE__, __LINE__, "prototxt: missing '}'"); return; }
    if (t.kind == Tok::RBRACE) { if (top) Fatal(__FILE__, __LINE__, "prototxt line " + std::to_string(t.line) + ": unmatched '}'"); return; }
    if (t.kind != Tok::IDENT) Fatal(__FILE__, __LINE__, "prototxt line " + std::to_string(t.line) + ": expected a field name, got '" + t.text + "'");
    Tok v = lx.next();
    if (v.kind == Tok::COLON) v = lx.next();
    PField f;
    if (v.kind == Tok::LBRACE) {
      f.msg.reset(new PMessage);
      parse_message(lx, *f.msg, false);
    } else if (v.kind == Tok::STRING || v.kind == Tok::NUMBER || v.kind == Tok::IDENT) {
      f.scalar = v.text;
    } else {
      Fatal(__FILE__, __LINE__, "prototxt line " + std::to_string(v.line) + ": expected a value for '" + t.text + "'");
    }
    m.fields.emplace_back(t.text, std::move(f));
  }
}
}  // namespace

PMessage ParseTextProto(const std::string& text) {
  Lexer lx(text);
  PMessage m;
  parse_message(lx, m, true);
  return m;
}
PMessage ParseTextProtoFile(const std::string& path) {
  std::ifstream in(path);
  if (!in) Fatal(__FILE__, __LINE__, "File not found: " + path);   // ReadProtoFromTextFile CHECK
  std::stringstream ss;
  ss << in.rdbuf();
  return ParseTextProto(ss.str());
}

bool PMessage::has(const std::string& k) const { for (auto& f : fields) if (f.first == k) return true; return false; }
std::vector<const PField*> PMessage::all(const std::string& k) const {
  std::vector<const PField*> v;
  for (auto& f : fields) if (f.first == k) v.push_back(&f.second);
  return v;
}
const PMessage* PMessage::sub(const std::string& k) const {
  for (auto& f : fields) if (f.first == k && f.second.is_msg()) return f.second.msg.get();
  return nullptr;
}
std::string PMessage::str(const std::string& k, const std::string& d) const {
  for (auto& f : fields) if (f.first == k && !f.second.is_msg()) return f.second.scalar;
  return d;
}
double PMessage::num(const std::string& k, double d) const { const std::string s = str(k); return s.empty() ? d : std::atof(s.c_str()); }
long long PMessage::integer(const std::string& k, long long d) const { const std::string s = str(k); return s.empty() ? d : std::atoll(s.c_str()); }
bool PMessage::boolean(const std::string& k, bool d) const { const std::string s = str(k); return s.empty() ? d : (s == "true" || s == "1" || s == "True"); }
std::vector<long long> PMessage::ints(const std::string& k) const {
  std::vector<long long> v;
  for (auto* f : all(k)) if (!f->is_msg()) v.push_back(std::atoll(f->scalar.c_str()));
  return v;
}

// ================================================================================================ Net
namespace {
bool rule_met(const PMessage& rule, Phase phase) {     // Net::StateMeetsRule with level 0 and no stages
  if (rule.has("phase")) {
    const std::string p = rule.str("phase");
    const Phase rp = (p == "TEST" || p == "1") ? TEST : TRAIN;
    if (rp != phase) return false;
  }
  if (rule.has("min_level") && rule.integer("min_level") > 0) return false;
  if (rule.has("max_level") && rule.integer("max_level") < 0) return false;
  if (rule.has("stage")) return false;                  // the net state carries no stages
  return true;
}
bool layer_included(const PMessage& lp, Phase phase) {  // Net::FilterNet
  const auto inc = lp.all("include"), exc = lp.all("exclude");
  if (!inc.empty() && !exc.empty()) Fatal(__FILE__, __LINE__, "Specify either include rules or exclude rules; not both.");
  if (inc.empty()) {
    for (auto* e : exc) if (e->is_msg() && rule_met(*e->msg, phase)) return false;
    return true;
  }
  for (auto* i : inc) if (i->is_msg() && rule_met(*i->msg, phase)) return true;
  return false;
}
int engine_of(const std::string& s) { return s == "CAFFE" || s == "1" ? B2C_ENGINE_CAFFE : s == "CUDNN" || s == "2" ? B2C_ENGINE_CUDNN : B2C_ENGINE_DEFAULT; }
FillerParameter filler_of(const PMessage* m) {
  FillerParameter f;
  if (!m) return f;
  f.type = m->str("type", "constant");
  f.value = (float)m->num("value", 0); f.min = (float)m->num("min", 0); f.max = (float)m->num("max", 1);
  f.mean = (float)m->num("mean", 0); f.std = (float)m->num("std", 1);
  const std::string vn = m->str("variance_norm", "FAN_IN");
  f.variance_norm = vn == "FAN_OUT" ? 1 : vn == "AVERAGE" ? 2 : 0;
  return f;
}
std::vector<int> to_int(const std::vector<long long>& v) { return std::vector<int>(v.begin(), v.end()); }
size_t prod(const std::vector<int>& s, size_t from = 0) { size_t c = 1; for (size_t i = from; i < s.size(); ++i) c *= (size_t)s[i]; return c; }
int pooled_extent(int in, int k, int s, int p) {        // PoolingLayer::Reshape: ceil mode + last-window clip
  int o = (int)std::ceil((float)(in + 2 * p - k) / s) + 1;
  if (p > 0 && (o - 1) * s >= in + p) --o;
  return o;
}
}  // namespace

Net Net::FromFile(const std::string& path, Phase phase, int batch_override) {
  return Net(ParseTextProtoFile(path), phase, batch_override);
}
size_t Net::learnable_count() const { size_t c = 0; for (auto& p : params_) c += p.count; return c; }

Net::Net(const PMessage& np, Phase phase, int batch_override, int default_channels, int default_size) {
  name_ = np.str("name");
  reduce_buckets_ = (int)np.integer("reduce_buckets", 6);          // caffe.proto:140
  global_grad_scale_ = (float)np.num("global_grad_scale", 1.0);    // caffe.proto:130
  if (!np.all("layers").empty()) Fatal(__FILE__, __LINE__, "V1 'layers' prototxt: run upgrade_net_proto_text first (not on this path)");
  // net-level inputs (deploy-style prototxts): input: "data" input_shape { dim: ... } / input_dim: ...
  {
    const auto names = np.all("input");
    const auto shp = np.all("input_shape");
    const auto dims = np.ints("input_dim");
    for (size_t i = 0; i < names.size(); ++i) {
      std::vector<int> s;
      if (i < shp.size() && shp[i]->is_msg()) s = to_int(shp[i]->msg->ints("dim"));
      else if (dims.size() >= 4 * (i + 1)) s.assign(dims.begin() + 4 * i, dims.begin() + 4 * i + 4);
      if (batch_override > 0 && !s.empty()) s[0] = batch_override;
      shapes_[names[i]->scalar] = s;
    }
  }
  std::map<std::string, bool> need_bw;
  for (auto* lf : np.all("layer")) {
    if (!lf->is_msg()) continue;
    const PMessage& lp = *lf->msg;
    if (!layer_included(lp, phase)) continue;
    NetLayer L;
    L.param.name = lp.str("name");
    L.param.type = lp.str("type");
    for (const PField* f : lp.all("loss_weight")) L.loss_weight.push_back((float)std::atof(f->scalar.c_str()));
    for (auto* b : lp.all("bottom")) L.param.bottom.push_back(b->scalar);
    for (auto* t : lp.all("top")) L.param.top.push_back(t->scalar);
    for (auto* ps : lp.all("param")) {
      ParamSpec s;
      if (ps->is_msg()) { s.lr_mult = (float)ps->msg->num("lr_mult", 1.0); s.decay_mult = (float)ps->msg->num("decay_mult", 1.0); }
      L.param.param.push_back(s);
    }
    const std::string& type = L.param.type;
    const int id = (int)layers_.size();
    auto bottom_shape = [&](int i) -> const std::vector<int>& {
      B2_CHECK(i < (int)L.param.bottom.size(), "layer " + L.param.name + " has too few bottoms");
      auto it = shapes_.find(L.param.bottom[i]);
      if (it == shapes_.end()) Fatal(__FILE__, __LINE__, "Unknown bottom blob '" + L.param.bottom[i] + "' (layer '" + L.param.name + "')");
      return it->second;
    };
    std::vector<std::vector<int>> tops;
    std::vector<std::pair<std::vector<int>, int>> blobs;   // learnable blob shapes (+ index into ParamSpecs)
    bool is_data = false;

    if (type == "Data" || type == "ImageData" || type == "HDF5Data") {
      is_data = true;
      const PMessage* dp = lp.sub("data_param");
      if (!dp) dp = lp.sub("image_data_param");
      const PMessage* tp = lp.sub("transform_param");
      L.batch_size = batch_override > 0 ? batch_override : (int)(dp ? dp->integer("batch_size", 1) : 1);
      L.crop_size = (int)(tp ? tp->integer("crop_size", 0) : 0);
      if (tp) {
        L.has_transform = true;
        L.mirror = tp->boolean("mirror", false);
        L.transform_scale = (float)tp->num("scale", 1.0);
        for (auto* f : tp->all("mean_value")) if (!f->is_msg()) L.mean_value.push_back((float)std::atof(f->scalar.c_str()));
      }
      int dc = default_channels, dh = default_size, dw = default_size;
      if (type == "Data" && dp) {
        L.data_source = dp->str("source");
        const std::string be = dp->str("backend", "LEVELDB");
        L.data_backend = (be == "LMDB" || be == "1") ? 1 : 0;
        L.parser_threads = (int)dp->integer("parser_threads", 0);
        L.force_encoded_color = dp->boolean("force_encoded_color", false);
        L.data_cache = dp->boolean("cache", false);
        L.data_shuffle = dp->boolean("shuffle", false);
        if (tp) { L.mean_file = tp->str("mean_file"); L.transform_random_seed = tp->integer("random_seed", -1); }
        // DataLayerSetUp reads one datum to size the top blob (data_layer.cpp:176-183); so does this, when the database is there
        L.use_database = UseDatabase(L.data_source, L.data_backend);
        if (L.use_database) PeekDatumShape(L.data_source, &dc, &dh, &dw, L.force_encoded_color);
      }
      if (L.crop_size > 0) {
        if (L.use_database) B2_CHECK(dh >= L.crop_size && dw >= L.crop_size, "crop_size larger than the datums of " + L.data_source);
        dh = dw = L.crop_size;
      }
      tops.push_back({L.batch_size, dc, dh, dw});
      tops.push_back({L.batch_size});
    } else if (type == "Input" || type == "DummyData") {
      is_data = true;
      const PMessage* ip = lp.sub(type == "Input" ? "input_param" : "dummy_data_param");
      B2_CHECK(ip != nullptr, type + " layer needs its parameter message");
      const auto shp = ip->all("shape");
      for (size_t i = 0; i < L.param.top.size(); ++i) {
        B2_CHECK(!shp.empty(), "Input layer needs shape");
        const PField* sf = shp[std::min(i, shp.size() - 1)];
        B2_CHECK(sf->is_msg(), type + " layer: `shape` must be a message { dim: ... }");
        std::vector<int> s = to_int(sf->msg->ints("dim"));
        if (batch_override > 0 && !s.empty()) s[0] = batch_override;
        tops.push_back(s);
      }
    } else if (type == "Convolution") {
      const PMessage* cp = lp.sub("convolution_param");
      B2_CHECK(cp != nullptr, "Convolution layer without convolution_param");
      ConvolutionParameter& c = L.param.convolution_param;
      c.num_output = (int)cp->integer("num_output");
      c.bias_term = cp->boolean("bias_term", true);
      c.pad = to_int(cp->ints("pad")); c.kernel_size = to_int(cp->ints("kernel_size"));
      c.stride = to_int(cp->ints("stride")); c.dilation = to_int(cp->ints("dilation"));
      if (cp->has("pad_h")) c.pad_h = (int)cp->integer("pad_h"); if (cp->has("pad_w")) c.pad_w = (int)cp->integer("pad_w");
      if (cp->has("kernel_h")) c.kernel_h = (int)cp->integer("kernel_h"); if (cp->has("kernel_w")) c.kernel_w = (int)cp->integer("kernel_w");
      if (cp->has("stride_h")) c.stride_h = (int)cp->integer("stride_h"); if (cp->has("stride_w")) c.stride_w = (int)cp->integer("stride_w");
      c.group = (int)cp->integer("group", 1);
      c.weight_filler = filler_of(cp->sub("weight_filler"));
      c.bias_filler = filler_of(cp->sub("bias_filler"));
      c.engine = engine_of(cp->str("engine", "DEFAULT"));
      c.axis = (int)cp->integer("axis", 1);
      c.force_nd_im2col = cp->boolean("force_nd_im2col", false);
      const std::vector<int>& bs = bottom_shape(0);
      B2_CHECK(bs.size() == 4, "2-D convolution expects a 4-D bottom");
      auto ax = [&](const std::vector<int>& rep, int hv, int wv, int dflt, int which) {
        if (hv >= 0 || wv >= 0) return which == 0 ? hv : wv;
        if (rep.empty()) return dflt;
        return rep[rep.size() == 1 ? 0 : which];
      };
      b2c_conv_params p{};
      p.N = bs[0]; p.C = bs[1]; p.H = bs[2]; p.W = bs[3]; p.O = c.num_output; p.G = c.group;
      p.kh = ax(c.kernel_size, c.kernel_h, c.kernel_w, 0, 0); p.kw = ax(c.kernel_size, c.kernel_h, c.kernel_w, 0, 1);
      p.sh = ax(c.stride, c.stride_h, c.stride_w, 1, 0); p.sw = ax(c.stride, c.stride_h, c.stride_w, 1, 1);
      p.ph = ax(c.pad, c.pad_h, c.pad_w, 0, 0); p.pw = ax(c.pad, c.pad_h, c.pad_w, 0, 1);
      p.dh = ax(c.dilation, -1, -1, 1, 0); p.dw = ax(c.dilation, -1, -1, 1, 1);
      p.has_bias = c.bias_term;
      B2_CHECK(p.kh > 0 && p.kw > 0, "Filter dimensions must be nonzero.");
      B2_CHECK(p.sh > 0 && p.sw > 0, "Stride dimensions must be nonzero.");                 // base_conv_layer.cpp:66-69
      B2_CHECK(p.dh > 0 && p.dw > 0 && p.ph >= 0 && p.pw >= 0, "dilation must be positive and pad non-negative");
      B2_CHECK(p.G > 0 && p.O > 0, "num_output and group must be positive");
      B2_CHECK(p.C % p.G == 0 && p.O % p.G == 0, "channels / num_output must be multiples of group");
      const int Ho = (p.H + 2 * p.ph - (p.dh * (p.kh - 1) + 1)) / p.sh + 1;     // conv_layer.cpp:7-22
      const int Wo = (p.W + 2 * p.pw - (p.dw * (p.kw - 1) + 1)) / p.sw + 1;
      tops.push_back({p.N, p.O, Ho, Wo});
      blobs.push_back({{p.O, p.C / p.G, p.kh, p.kw}, 0});
      if (c.bias_term) blobs.push_back({{p.O}, 1});
      const bool pd = need_bw.count(L.param.bottom[0]) ? need_bw[L.param.bottom[0]] : false;   // net.cpp:183-191,278-283
      convs_.push_back(ConvEntry{L.param.name, id, p, pd});
    } else if (type == "Pooling") {
      const PMessage* pp = lp.sub("pooling_param");
      const std::vector<int>& bs = bottom_shape(0);
      B2_CHECK(bs.size() == 4, "Pooling expects a 4-D bottom (num, channels, height, width)");   // pooling_layer.cpp:84-86
      PoolingParameter& q = L.pooling;
      if (pp) {
        const std::string pool = pp->str("pool", "MAX");
        q.pool = pool == "AVE" || pool == "1" ? 1 : pool == "STOCHASTIC" || pool == "2" ? 2 : 0;
        q.global_pooling = pp->boolean("global_pooling", false);
        const int k = (int)pp->integer("kernel_size", 0), s = (int)pp->integer("stride", 1), p = (int)pp->integer("pad", 0);
        q.kernel_h = pp->has("kernel_h") ? (int)pp->integer("kernel_h") : k; q.kernel_w = pp->has("kernel_w") ? (int)pp->integer("kernel_w") : k;
        q.stride_h = pp->has("stride_h") ? (int)pp->integer("stride_h") : s; q.stride_w = pp->has("stride_w") ? (int)pp->integer("stride_w") : s;
        q.pad_h = pp->has("pad_h") ? (int)pp->integer("pad_h") : p; q.pad_w = pp->has("pad_w") ? (int)pp->integer("pad_w") : p;
      }
      if (q.global_pooling) { q.kernel_h = bs[2]; q.kernel_w = bs[3]; q.stride_h = q.stride_w = 1; q.pad_h = q.pad_w = 0; }
      B2_CHECK(q.kernel_h > 0 && q.kernel_w > 0, "Filter dimensions cannot be zero.");
      B2_CHECK(q.stride_h > 0 && q.stride_w > 0 && q.pad_h >= 0 && q.pad_w >= 0, "Pooling stride must be positive and pad non-negative");
      tops.push_back({bs[0], bs[1], pooled_extent(bs[2], q.kernel_h, q.stride_h, q.pad_h), pooled_extent(bs[3], q.kernel_w, q.stride_w, q.pad_w)});
    } else if (type == "InnerProduct") {
      const PMessage* ip = lp.sub("inner_product_param");
      B2_CHECK(ip != nullptr, "InnerProduct layer without inner_product_param");
      L.ip_num_output = (int)ip->integer("num_output");
      L.ip_bias = ip->boolean("bias_term", true);
      L.ip_weight_filler = filler_of(ip->sub("weight_filler"));
      L.ip_bias_filler = filler_of(ip->sub("bias_filler"));
      const std::vector<int>& bs = bottom_shape(0);
      B2_CHECK(bs.size() >= 2 && L.ip_num_output > 0, "InnerProduct needs a bottom with at least 2 axes and a positive num_output");
      const int K = (int)prod(bs, 1);
      tops.push_back({bs[0], L.ip_num_output});
      blobs.push_back({{L.ip_num_output, K}, 0});
      if (L.ip_bias) blobs.push_back({{L.ip_num_output}, 1});
    } else if (type == "BatchNorm") {
      const PMessage* bp = lp.sub("batch_norm_param");
      L.bn_scale_bias = bp ? (bp->boolean("scale_bias", false) || bp->has("scale_filler") || bp->has("bias_filler")) : false;
      if (bp) { L.bn_eps = std::max((float)bp->num("eps", 1e-5), 1e-5f); L.bn_maf = (float)bp->num("moving_average_fraction", 0.999); }
      if (bp && bp->has("scale_filler")) { L.bn_scale_filler = filler_of(bp->sub("scale_filler")); L.bn_has_scale_filler = true; }
      if (bp && bp->has("bias_filler")) { L.bn_bias_filler = filler_of(bp->sub("bias_filler")); L.bn_has_bias_filler = true; }
      const std::vector<int>& bs = bottom_shape(0);
      B2_CHECK(bs.size() >= 2, "BatchNorm needs a bottom with a channel axis");
      tops.push_back(bs);
      // blobs_[0..2] = running mean / variance / correction (statistics, not exchanged with the cuDNN engine:
      // include/caffe/layers/cudnn_batch_norm_layer.hpp:30-32); [3], [4] = scale, bias when scale_bias
      if (L.bn_scale_bias) { blobs.push_back({{bs[1]}, 3}); blobs.push_back({{bs[1]}, 4}); }
    } else if (type == "Scale") {
      const PMessage* sp = lp.sub("scale_param");
      const std::vector<int>& bs = bottom_shape(0);
      B2_CHECK(bs.size() >= 2, "Scale needs a bottom with a channel axis");
      tops.push_back(bs);
      if (L.param.bottom.size() == 1) {
        blobs.push_back({{bs[1]}, 0});
        if (sp && sp->boolean("bias_term", false)) blobs.push_back({{bs[1]}, 1});
      }
    } else if (type == "Concat") {
      const PMessage* cp = lp.sub("concat_param");
      L.concat_axis = cp ? (int)cp->integer("axis", 1) : 1;
      std::vector<int> s = bottom_shape(0);
      B2_CHECK(L.concat_axis >= 0 && L.concat_axis < (int)s.size(), "Concat: axis outside the bottom's axes");
      for (size_t i = 1; i < L.param.bottom.size(); ++i) {
        const std::vector<int>& b = bottom_shape((int)i);
        B2_CHECK(b.size() == s.size(), "Concat: all inputs must have the same #axes.");
        s[L.concat_axis] += b[L.concat_axis];
      }
      tops.push_back(s);
    } else if (type == "SoftmaxWithLoss" || type == "EuclideanLoss" || type == "SigmoidCrossEntropyLoss" || type == "Accuracy") {
      if (const PMessage* ap = lp.sub("accuracy_param")) L.accuracy_top_k = (int)ap->integer("top_k", 1);
      for (size_t i = 0; i < L.param.top.size(); ++i) tops.push_back({});
    } else if (type == "ReLU" && (L.relu_slope = (float)(lp.sub("relu_param") ? lp.sub("relu_param")->num("negative_slope", 0.0) : 0.0), false)) {
    } else if ((type == "LRN" || type == "Dropout") && ([&] {
                 if (const PMessage* q = lp.sub("lrn_param")) {
                   L.lrn_size = (int)q->integer("local_size", 5); L.lrn_alpha = (float)q->num("alpha", 1.0); L.lrn_beta = (float)q->num("beta", 0.75);
                   L.lrn_k = (float)q->num("k", 1.0);
                   const std::string r = q->str("norm_region", "ACROSS_CHANNELS");
                   L.lrn_region = (r == "WITHIN_CHANNEL" || r == "1") ? 1 : 0;
                 }
                 if (const PMessage* q = lp.sub("dropout_param")) L.dropout_ratio = (float)q->num("dropout_ratio", 0.5);
               }(), false)) {
    } else if (type == "Eltwise" && ([&] {
                 if (const PMessage* q = lp.sub("eltwise_param")) {
                   const std::string op = q->str("operation", "SUM");
                   L.eltwise_op = (op == "PROD" || op == "0") ? 0 : (op == "MAX" || op == "2") ? 2 : 1;
                   for (auto* f : q->all("coeff")) if (!f->is_msg()) L.eltwise_coeff.push_back((float)std::atof(f->scalar.c_str()));
                 }
               }(), false)) {
    } else if (type == "ReLU" || type == "Dropout" || type == "LRN" || type == "Eltwise" || type == "Softmax" || type == "Sigmoid" ||
               type == "TanH" || type == "Power" || type == "Bias" || type == "ELU" || type == "PReLU" || type == "Split") {
      for (size_t i = 0; i < std::max<size_t>(1, L.param.top.size()); ++i) tops.push_back(bottom_shape(0));
    } else {
      Fatal(__FILE__, __LINE__, "Unknown layer type: " + type + " (layer '" + L.param.name + "')");
    }

    // need-backward analysis (Net::Init, net.cpp:160-283): data tops never need it; a layer needs backward if a
    // bottom does or it owns a parameter with a non-zero lr_mult
    bool layer_bw = false;
    for (auto& b : L.param.bottom) layer_bw |= need_bw.count(b) ? need_bw[b] : false;
    for (auto& bl : blobs) {
      const ParamSpec spec = bl.second < (int)L.param.param.size() ? L.param.param[bl.second] : ParamSpec();
      if (spec.lr_mult != 0.f) layer_bw = true;
    }
    if (is_data) layer_bw = false;
    for (size_t i = 0; i < L.param.top.size(); ++i) {
      B2_CHECK(i < tops.size(), "layer '" + L.param.name + "' declares more tops than its type produces");
      shapes_[L.param.top[i]] = tops[i];
      need_bw[L.param.top[i]] = layer_bw;
    }
    int bid = 0;
    for (auto& bl : blobs) {
      const ParamSpec spec = bl.second < (int)L.param.param.size() ? L.param.param[bl.second] : ParamSpec();
      params_.push_back(LearnableParam{L.param.name, id, bid++, prod(bl.first), bl.first, spec.lr_mult, spec.decay_mult});
    }
    layer_top_shapes_.push_back(tops);
    layers_.push_back(std::move(L));
  }
}

// ================================================================================================ solver
SolverParameter ReadSolverParameter(const PMessage& m, std::string* net_path) {
  SolverParameter p;
  p.base_lr = (float)m.num("base_lr", p.base_lr);
  p.lr_policy = m.str("lr_policy", "fixed");
  p.gamma = (float)m.num("gamma", p.gamma);
  p.power = (float)m.num("power", p.power);
  p.momentum = (float)m.num("momentum", 0.0);
  p.momentum_policy = m.str("momentum_policy", "fixed");
  p.max_momentum = (float)m.num("max_momentum", 0.99);
  p.momentum_power = (float)m.num("momentum_power", 1.0);
  p.weight_decay = (float)m.num("weight_decay", 0.0);
  p.regularization_type = m.str("regularization_type", "L2");
  p.stepsize = (int)m.integer("stepsize", 1);
  for (long long v : m.ints("stepvalue")) p.stepvalue.push_back((int)v);
  p.max_iter = (int)m.integer("max_iter", 1);
  p.iter_size = (int)m.integer("iter_size", 1);
  p.clip_gradients = (float)m.num("clip_gradients", -1.0);
  p.rampup_interval = (int)m.integer("rampup_interval", 0);
  p.rampup_lr = (float)m.num("rampup_lr", 0.0);
  p.min_lr = (float)m.num("min_lr", 0.0);
  p.snapshot_diff = m.boolean("snapshot_diff", false);
  const std::string type = m.str("type", "SGD");
  if (type != "SGD") Fatal(__FILE__, __LINE__, "solver type '" + type + "' is outside this path (SGD only; REGISTER_SOLVER_CLASS(SGD), sgd_solver.cpp:357)");
  if (net_path) *net_path = m.str("net", m.str("train_net"));
  return p;
}

}  // namespace caffe
