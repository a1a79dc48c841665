// Native host-side data loader: multi-threaded parser for the reference's
// "name:type" tab-separated node / edge tables
// (docs/en/gl/graph/data_loader.md:117-153; behaviour of
// graphlearn/src/core/io/{edge_loader,node_loader,parser,slice_reader}.cc).
//
// Design: the file is read once into memory, cut into `threads` slices at line
// boundaries (the reference slices per server x thread,
// graphlearn/src/core/io/slice_reader.h:60-86), every slice is parsed by its
// own thread straight into columnar vectors, and the columns are concatenated
// into torch tensors ready for a pinned-host -> HBM copy.  Columnar output
// (ids, weights, labels, timestamps, int/float attribute matrices, string
// blob + offsets) replaces the reference's per-record AttributeValue objects.
#include <torch/extension.h>
#include <algorithm>
#include <cerrno>
#include <charconv>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <thread>
#include <vector>

namespace glb {

namespace {

struct Schema {
  bool is_edge;
  bool weighted, labeled, timestamped;
  std::vector<int> attr_types;      // 0 = int, 1 = float, 2 = string
  std::vector<int64_t> buckets;     // per attribute: >0 -> hash into buckets (becomes an int attr)
  char attr_delim;
  char field_delim;
  int n_int = 0, n_float = 0, n_str = 0;
};

struct Columns {
  std::vector<int64_t> a, b;        // id | (src, dst)
  std::vector<float> w;
  std::vector<int64_t> label, ts;
  std::vector<int64_t> iattr;
  std::vector<float> fattr;
  std::vector<char> sblob;
  std::vector<int64_t> soff;        // end offset of every string attr, relative to this slice
  std::string error;
};

inline uint64_t fnv1a(const char* s, size_t n) {
  uint64_t h = 1469598103934665603ull;
  for (size_t i = 0; i < n; ++i) { h ^= (unsigned char)s[i]; h *= 1099511628211ull; }
  return h;
}

inline bool parse_i64(const char* s, const char* e, int64_t* out) {
  if (s == e) return false;
  bool neg = false;
  if (*s == '-') { neg = true; ++s; } else if (*s == '+') { ++s; }
  if (s == e) return false;
  uint64_t v = 0;
  for (; s < e; ++s) {
    if (*s < '0' || *s > '9') {
      if (*s == '\r' && s + 1 == e) break;
      return false;
    }
    v = v * 10 + (uint64_t)(*s - '0');
  }
  *out = neg ? -(int64_t)v : (int64_t)v;
  return true;
}

inline bool parse_f32_slow(const char* s, const char* e, float* out) {
  char buf[64];
  size_t n = (size_t)(e - s);
  if (n == 0 || n >= sizeof(buf)) return false;
  std::memcpy(buf, s, n);
  buf[n] = 0;
  char* endp = nullptr;
  errno = 0;
  float v = std::strtof(buf, &endp);
  if (endp == buf) return false;
  while (*endp == '\r' || *endp == ' ') ++endp;
  if (*endp != 0) return false;
  *out = v;
  return true;
}

// std::from_chars (Eisel-Lemire, correctly rounded, no locale, no copy) is ~4x faster than strtof on attribute
// strings; anything it does not take verbatim (leading '+', blanks, hex floats, ...) goes to the strtof path.
inline bool parse_f32(const char* s, const char* e, float* out) {
  while (e > s && (e[-1] == '\r' || e[-1] == ' ')) --e;
  if (s == e) return false;
  const char* b = (*s == '+') ? s + 1 : s;
  float v;
  auto r = std::from_chars(b, e, v);
  if (r.ec == std::errc() && r.ptr == e) { *out = v; return true; }
  return parse_f32_slow(s, e, out);
}

// parse lines in [begin, end)
void parse_slice(const char* begin, const char* end, const Schema& sc, Columns* out) {
  const char* p = begin;
  int64_t line_no = 0;
  while (p < end) {
    const char* eol = (const char*)memchr(p, '\n', (size_t)(end - p));
    if (!eol) eol = end;
    const char* le = eol;
    if (le > p && le[-1] == '\r') --le;
    if (le > p) {
      // split fields
      const char* f = p;
      int field = 0;
      auto next = [&](const char** fs, const char** fe) -> bool {
        if (f > le) return false;
        const char* t = (const char*)memchr(f, sc.field_delim, (size_t)(le - f));
        if (!t) t = le;
        *fs = f; *fe = t;
        f = t + 1;
        ++field;
        return true;
      };
      const char *fs, *fe;
      bool ok = true;
      int64_t v;
      float fv;
      if (!next(&fs, &fe) || !parse_i64(fs, fe, &v)) ok = false; else out->a.push_back(v);
      if (ok && sc.is_edge) { if (!next(&fs, &fe) || !parse_i64(fs, fe, &v)) ok = false; else out->b.push_back(v); }
      if (ok && sc.weighted) { if (!next(&fs, &fe) || !parse_f32(fs, fe, &fv)) ok = false; else out->w.push_back(fv); }
      if (ok && sc.labeled) { if (!next(&fs, &fe) || !parse_i64(fs, fe, &v)) ok = false; else out->label.push_back(v); }
      if (ok && sc.timestamped) { if (!next(&fs, &fe) || !parse_i64(fs, fe, &v)) ok = false; else out->ts.push_back(v); }
      if (ok && !sc.attr_types.empty()) {
        if (!next(&fs, &fe)) ok = false;
        else {
          // the attribute column runs to the end of the line (it may itself contain the field delimiter
          // only if it differs from attr_delim; keep reference behaviour: one column)
          const char* as = fs;
          const char* ae = fe;
          size_t na = sc.attr_types.size();
          for (size_t i = 0; i < na && ok; ++i) {
            const char* t = (i + 1 == na) ? ae : (const char*)memchr(as, sc.attr_delim, (size_t)(ae - as));
            if (!t) { ok = false; break; }
            int ty = sc.attr_types[i];
            int64_t bucket = sc.buckets.empty() ? 0 : sc.buckets[i];
            if (ty == 0) {
              if (!parse_i64(as, t, &v)) { ok = false; break; }
              out->iattr.push_back(v);   // int buckets are applied by the feature encoder, not the loader
            } else if (ty == 1) {
              if (!parse_f32(as, t, &fv)) { ok = false; break; }
              out->fattr.push_back(fv);
            } else {
              if (bucket > 0) out->iattr.push_back((int64_t)(fnv1a(as, (size_t)(t - as)) % (uint64_t)bucket));
              else { out->sblob.insert(out->sblob.end(), as, t); out->soff.push_back((int64_t)out->sblob.size()); }
            }
            as = t + 1;
          }
        }
      }
      if (!ok) {
        if (out->error.empty())
          out->error = "malformed record (slice line " + std::to_string(line_no) + "): " +
                       std::string(p, std::min<size_t>((size_t)(le - p), 120));
        return;
      }
    }
    ++line_no;
    p = eol + 1;
  }
}

template <typename T>
at::Tensor concat(const std::vector<Columns>& cols, std::vector<T> Columns::*field, at::ScalarType st) {
  int64_t n = 0;
  for (auto& c : cols) n += (int64_t)(c.*field).size();
  auto t = at::empty({n}, at::TensorOptions().dtype(st));
  T* dst = reinterpret_cast<T*>(t.data_ptr());
  for (auto& c : cols) {
    auto& v = c.*field;
    if (!v.empty()) std::memcpy(dst, v.data(), v.size() * sizeof(T));
    dst += v.size();
  }
  return t;
}

}  // namespace

// returns [a, b, weights, labels, timestamps, int_attrs[n,i], float_attrs[n,f], str_blob(uint8), str_offsets(int64)]
std::vector<at::Tensor> load_table(const std::string& path, bool is_edge, bool weighted, bool labeled,
                                   bool timestamped, std::vector<int64_t> attr_types,
                                   std::vector<int64_t> buckets, const std::string& attr_delim,
                                   const std::string& field_delim, int64_t threads, int64_t part_index,
                                   int64_t part_count) {
  Schema sc;
  sc.is_edge = is_edge; sc.weighted = weighted; sc.labeled = labeled; sc.timestamped = timestamped;
  for (auto t : attr_types) sc.attr_types.push_back((int)t);
  sc.buckets = buckets;
  TORCH_CHECK(buckets.empty() || buckets.size() == attr_types.size(), "bucket list must match attr_types");
  sc.attr_delim = attr_delim.empty() ? ':' : attr_delim[0];
  sc.field_delim = field_delim.empty() ? '\t' : field_delim[0];
  for (size_t i = 0; i < sc.attr_types.size(); ++i) {
    int64_t bk = sc.buckets.empty() ? 0 : sc.buckets[i];
    if (sc.attr_types[i] == 0) sc.n_int++;
    else if (sc.attr_types[i] == 1) sc.n_float++;
    else if (bk > 0) sc.n_int++;
    else sc.n_str++;
  }

  // Read only this part's byte range of the file (SliceReader semantics,
  // graphlearn/src/core/io/slice_reader.h:60-86: `part_count` contiguous record ranges).  A record
  // belongs to the part whose raw byte range contains its first byte.
  FILE* fp = std::fopen(path.c_str(), "rb");
  TORCH_CHECK(fp != nullptr, "cannot open data source: ", path);
  std::fseek(fp, 0, SEEK_END);
  const long sz = std::ftell(fp);
  TORCH_CHECK(part_count >= 1 && part_index >= 0 && part_index < part_count, "bad part index/count");
  // header detection: "name:type<TAB>name:type..." (first field not numeric)
  long data_begin = 0;
  {
    std::vector<char> head((size_t)std::min<long>(sz, 1 << 16));
    std::fseek(fp, 0, SEEK_SET);
    size_t hn = head.empty() ? 0 : std::fread(head.data(), 1, head.size(), fp);
    const char* hb = head.data();
    const char* he = hb + hn;
    if (hb < he) {
      const char* eol = (const char*)memchr(hb, '\n', (size_t)(he - hb));
      if (!eol) eol = he;
      const char* t = (const char*)memchr(hb, sc.field_delim, (size_t)(eol - hb));
      if (!t) t = eol;
      int64_t dummy;
      const char* fe = t;
      if (fe > hb && fe[-1] == '\r') --fe;
      if (!parse_i64(hb, fe, &dummy)) data_begin = (eol < he) ? (long)(eol - hb) + 1 : (long)hn;
    }
  }
  auto line_start_at_or_after = [&](long c) -> long {   // c > data_begin
    long pos = c - 1;
    char tmp[1 << 14];
    std::fseek(fp, pos, SEEK_SET);
    while (pos < sz) {
      size_t n = std::fread(tmp, 1, (size_t)std::min<long>((long)sizeof(tmp), sz - pos), fp);
      if (n == 0) break;
      const char* nl = (const char*)memchr(tmp, '\n', n);
      if (nl) return pos + (long)(nl - tmp) + 1;
      pos += (long)n;
    }
    return sz;
  };
  const long span = sz - data_begin;
  long lo = data_begin, hi = sz;
  if (part_count > 1) {
    const long c0 = data_begin + (long)((__int128)span * part_index / part_count);
    const long c1 = data_begin + (long)((__int128)span * (part_index + 1) / part_count);
    lo = (part_index == 0 || c0 <= data_begin) ? data_begin : line_start_at_or_after(c0);
    hi = (part_index == part_count - 1) ? sz : (c1 <= data_begin ? data_begin : line_start_at_or_after(c1));
    if (hi < lo) hi = lo;
  }
  std::vector<char> buf((size_t)(hi - lo));
  std::fseek(fp, lo, SEEK_SET);
  size_t rd = buf.empty() ? 0 : std::fread(buf.data(), 1, buf.size(), fp);
  std::fclose(fp);
  TORCH_CHECK(rd == buf.size(), "short read on ", path);
  const char* begin = buf.data();
  const char* end = begin + buf.size();

  int nt = (int)std::max<int64_t>(1, std::min<int64_t>(threads, 64));
  if ((end - begin) < (1 << 16)) nt = 1;
  std::vector<const char*> cuts(nt + 1);
  cuts[0] = begin; cuts[nt] = end;
  for (int i = 1; i < nt; ++i) {
    const char* c = begin + (size_t)((end - begin) / nt) * i;
    if (c < cuts[i - 1]) c = cuts[i - 1];
    const char* nl = (const char*)memchr(c, '\n', (size_t)(end - c));
    cuts[i] = nl ? nl + 1 : end;
  }
  std::vector<Columns> cols(nt);
  std::vector<std::thread> pool;
  for (int i = 0; i < nt; ++i)
    pool.emplace_back([&, i] { parse_slice(cuts[i], cuts[i + 1], sc, &cols[i]); });
  for (auto& th : pool) th.join();
  for (auto& c : cols) TORCH_CHECK(c.error.empty(), path, ": ", c.error);

  auto a = concat<int64_t>(cols, &Columns::a, at::kLong);
  auto b = concat<int64_t>(cols, &Columns::b, at::kLong);
  auto w = concat<float>(cols, &Columns::w, at::kFloat);
  auto lb = concat<int64_t>(cols, &Columns::label, at::kLong);
  auto ts = concat<int64_t>(cols, &Columns::ts, at::kLong);
  auto ia = concat<int64_t>(cols, &Columns::iattr, at::kLong);
  auto fa = concat<float>(cols, &Columns::fattr, at::kFloat);
  int64_t n = a.numel();
  if (sc.n_int > 0) ia = ia.view({n, sc.n_int});
  if (sc.n_float > 0) fa = fa.view({n, sc.n_float});
  // strings: rebase per-slice offsets
  int64_t nblob = 0, noff = 0;
  for (auto& c : cols) { nblob += (int64_t)c.sblob.size(); noff += (int64_t)c.soff.size(); }
  auto blob = at::empty({nblob}, at::TensorOptions().dtype(at::kByte));
  auto off = at::zeros({noff + 1}, at::TensorOptions().dtype(at::kLong));
  {
    char* bd = reinterpret_cast<char*>(blob.data_ptr());
    int64_t* od = off.data_ptr<int64_t>() + 1;
    int64_t base = 0;
    for (auto& c : cols) {
      if (!c.sblob.empty()) std::memcpy(bd + base, c.sblob.data(), c.sblob.size());
      for (auto o : c.soff) *od++ = base + o;
      base += (int64_t)c.sblob.size();
    }
  }
  return {a, b, w, lb, ts, ia, fa, blob, off};
}

// Embedding / result dump in the reference's "id:int64\temb:string" dialect
// (graphlearn/examples/tf/trainer.py:214-279) so results can be re-ingested as a node table.
void save_embeddings(const std::string& path, const at::Tensor& ids, const at::Tensor& emb, bool header) {
  auto idc = ids.to(at::kCPU).to(at::kLong).contiguous();
  auto ec = emb.to(at::kCPU).to(at::kFloat).contiguous();
  TORCH_CHECK(ec.dim() == 2 && ec.size(0) == idc.numel(), "emb must be [n, d]");
  FILE* fp = std::fopen(path.c_str(), "wb");
  TORCH_CHECK(fp != nullptr, "cannot open for write: ", path);
  if (header) std::fputs("id:int64\temb:string\n", fp);
  const int64_t* ip = idc.data_ptr<int64_t>();
  const float* ep = ec.data_ptr<float>();
  int64_t n = idc.numel(), d = ec.size(1);
  std::string line;
  char tmp[64];
  for (int64_t i = 0; i < n; ++i) {
    line.clear();
    std::snprintf(tmp, sizeof(tmp), "%lld\t", (long long)ip[i]);
    line += tmp;
    for (int64_t j = 0; j < d; ++j) {
      std::snprintf(tmp, sizeof(tmp), j + 1 == d ? "%.6g" : "%.6g,", ep[i * d + j]);
      line += tmp;
    }
    line += '\n';
    std::fwrite(line.data(), 1, line.size(), fp);
  }
  std::fclose(fp);
}

// ---------------------------------------------------------------------------------------------------------------
// Streaming-record parser of the dynamic graph service's file loader (dgs/file_loader.py; reference: the C++ data-loader
// SDK + apps/file_loader, dynamic_graph_service/dataloader/apps/file_loader/loader.cc).  A data line is
//     <type name><delim><field 1><delim>...                e.g.  "u2i,3,17,1001,0.5"   /   "item,17,1000,0.1:0.2:0.3"
// with the field order given per type by a pattern file.  One call parses up to `max_records` records starting at byte
// `offset` and returns them as COLUMNS per type, ready for `apply_updates`:
//     vertex type: ids int64 [n], ts int64 [n], feat fp32 [n, d]          edge type: src, dst, ts int64 [n], w fp32 [n]
// (4 tensor slots per type, in `names` order; unused slots are empty) + meta int64 [next_offset, records, skipped_lines, eof].
// Lines of unknown types or with a wrong field count are skipped like the Python loader does; malformed numbers are errors.
// ---------------------------------------------------------------------------------------------------------------
std::vector<at::Tensor> parse_records(const std::string& path, int64_t offset, int64_t max_records, const std::string& delim,
                                      const std::string& list_delim, const std::vector<std::string>& names,
                                      const std::vector<int64_t>& kinds,          // 0 vertex, 1 edge
                                      const std::vector<int64_t>& n_fields,       // fields after the type name
                                      const std::vector<int64_t>& ts_idx,         // field index of the timestamp or -1
                                      const std::vector<int64_t>& w_idx,          // field index of the weight or -1 (edges)
                                      const std::vector<int64_t>& feat_off,       // [types + 1] ranges into the two lists below
                                      const std::vector<int64_t>& feat_field,     // field indices forming the feature row, in order
                                      const std::vector<int64_t>& feat_is_list,   // 1: list_delim separated values
                                      int64_t window_bytes) {
  const size_t T = names.size();
  TORCH_CHECK(delim.size() == 1 && list_delim.size() == 1, "single-character delimiters only");
  TORCH_CHECK(kinds.size() == T && n_fields.size() == T && ts_idx.size() == T && w_idx.size() == T && feat_off.size() == T + 1 &&
              feat_field.size() == feat_is_list.size() && (int64_t)feat_field.size() == feat_off[T], "inconsistent record specs");
  TORCH_CHECK(max_records > 0 && offset >= 0 && window_bytes >= 4096);
  const char dl = delim[0], ldl = list_delim[0];
  FILE* fp = std::fopen(path.c_str(), "rb");
  TORCH_CHECK(fp != nullptr, "cannot open ", path);
  std::fseek(fp, 0, SEEK_END);
  const long sz = std::ftell(fp);
  std::vector<char> buf;
  long got = 0;
  bool window_hits_eof = false;
  // read a window; grow it until it holds at least one complete line (or the rest of the file)
  for (int64_t want = window_bytes;; want *= 4) {
    const long n = (long)std::min<int64_t>(want, std::max<long>(0, sz - (long)offset));
    buf.resize((size_t)n);
    std::fseek(fp, (long)offset, SEEK_SET);
    got = n > 0 ? (long)std::fread(buf.data(), 1, (size_t)n, fp) : 0;
    window_hits_eof = (long)offset + got >= sz;
    if (window_hits_eof || got == 0 || memchr(buf.data(), '\n', (size_t)got) != nullptr) break;
  }
  std::fclose(fp);
  const char* base = buf.data();
  const char* end = base + got;
  if (!window_hits_eof) {                        // cut at the last complete line
    const char* q = end;
    while (q > base && q[-1] != '\n') --q;
    end = q;
  }
  struct Col { std::vector<int64_t> a, b, ts; std::vector<float> w, feat; int64_t d = -1; };
  std::vector<Col> cols(T);
  int64_t records = 0, skipped = 0;
  std::vector<std::pair<const char*, const char*>> f;
  const char* p = base;
  while (p < end && records < max_records) {
    const char* eol = (const char*)memchr(p, '\n', (size_t)(end - p));
    const char* next = eol ? eol + 1 : end;
    const char* le = eol ? eol : end;
    if (le > p && le[-1] == '\r') --le;
    if (le == p) { p = next; continue; }
    // split
    f.clear();
    const char* fs = p;
    for (const char* c = p; c <= le; ++c) {
      if (c == le || *c == dl) { f.emplace_back(fs, c); fs = c + 1; }
    }
    size_t t = T;
    for (size_t i = 0; i < T; ++i)
      if (names[i].size() == (size_t)(f[0].second - f[0].first) && std::memcmp(names[i].data(), f[0].first, names[i].size()) == 0) { t = i; break; }
    if (t == T || (int64_t)f.size() - 1 != n_fields[t] || n_fields[t] < (kinds[t] == 1 ? 2 : 1)) { ++skipped; p = next; continue; }
    Col& c = cols[t];
    auto fld = [&](int64_t i) { return f[(size_t)i + 1]; };
    auto fail = [&](const char* what) {
      TORCH_CHECK(false, path, ": bad ", what, " in record at byte ", (long long)(offset + (p - base)), ": ", std::string(p, (size_t)std::min<long>(le - p, 120)));
    };
    int64_t ts = 0;
    if (ts_idx[t] >= 0 && !parse_i64(fld(ts_idx[t]).first, fld(ts_idx[t]).second, &ts)) fail("timestamp");
    int64_t k0 = 0, k1 = 0;
    if (!parse_i64(fld(0).first, fld(0).second, &k0)) fail("id");
    if (kinds[t] == 1) {
      if (n_fields[t] < 2 || !parse_i64(fld(1).first, fld(1).second, &k1)) fail("destination id");
      float w = 1.f;
      if (w_idx[t] >= 0 && !parse_f32(fld(w_idx[t]).first, fld(w_idx[t]).second, &w)) fail("weight");
      c.a.push_back(k0); c.b.push_back(k1); c.ts.push_back(ts); c.w.push_back(w);
    } else {
      const size_t before = c.feat.size();
      for (int64_t j = feat_off[t]; j < feat_off[t + 1]; ++j) {
        const auto fv = fld(feat_field[(size_t)j]);
        if (feat_is_list[(size_t)j]) {
          const char* vs = fv.first;
          for (const char* q = fv.first; q <= fv.second; ++q) {
            if (q == fv.second || *q == ldl) {
              if (q > vs) { float v; if (!parse_f32(vs, q, &v)) fail("list attribute"); c.feat.push_back(v); }
              vs = q + 1;
            }
          }
        } else if (fv.second > fv.first) {
          float v;
          if (!parse_f32(fv.first, fv.second, &v)) fail("attribute");
          c.feat.push_back(v);
        }
      }
      const int64_t d = (int64_t)(c.feat.size() - before);
      if (c.d < 0) c.d = d;
      if (d != c.d) fail("feature width (ragged rows)");
      c.a.push_back(k0); c.ts.push_back(ts);
    }
    ++records;
    p = next;
  }
  const int64_t consumed = (int64_t)(p - base);
  auto i64 = [](const std::vector<int64_t>& v) { return at::from_blob((void*)v.data(), {(int64_t)v.size()}, at::kLong).clone(); };
  auto f32 = [](const std::vector<float>& v) { return at::from_blob((void*)v.data(), {(int64_t)v.size()}, at::kFloat).clone(); };
  std::vector<at::Tensor> out;
  out.reserve(T * 4 + 1);
  for (size_t t = 0; t < T; ++t) {
    const Col& c = cols[t];
    if (kinds[t] == 1) { out.push_back(i64(c.a)); out.push_back(i64(c.b)); out.push_back(i64(c.ts)); out.push_back(f32(c.w)); }
    else {
      out.push_back(i64(c.a)); out.push_back(i64(c.ts));
      out.push_back(f32(c.feat).reshape({(int64_t)c.a.size(), c.d < 0 ? 0 : c.d}));
      out.push_back(at::empty({0}, at::kFloat));
    }
  }
  const bool eof = (long)(offset + consumed) >= sz;
  auto meta = at::empty({4}, at::kLong);
  meta[0] = offset + consumed; meta[1] = records; meta[2] = skipped; meta[3] = eof ? 1 : 0;
  out.push_back(meta);
  return out;
}

}  // namespace glb
