#include "model_dir.h"

#include <cstdio>
#include <cstring>
#include <fstream>
#include <iterator>
#include <sstream>

namespace pf {

namespace {
#pragma pack(push, 1)
struct CfgRec { char key[32]; double value; };
struct TensorRec { char name[96]; uint32_t ndim; uint64_t dims[4]; uint64_t offset; uint64_t nbytes; };
#pragma pack(pop)
static_assert(sizeof(CfgRec) == 40, "cfg record");
static_assert(sizeof(TensorRec) == 148, "tensor record");
}  // namespace

bool read_weight_file(const std::string& path, WeightFile* out, std::string* err) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) { *err = "cannot open " + path; return false; }
  auto fail = [&](const std::string& m) { *err = m + " (" + path + ")"; fclose(f); return false; };
  char magic[8];
  if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "B2PFWTS1", 8) != 0) return fail("bad magic");
  uint32_t n_cfg = 0, n_t = 0;
  if (fread(&n_cfg, 4, 1, f) != 1 || n_cfg > 4096) return fail("bad cfg count");
  for (uint32_t i = 0; i < n_cfg; ++i) {
    CfgRec r;
    if (fread(&r, sizeof(r), 1, f) != 1) return fail("truncated cfg");
    r.key[31] = 0;
    out->cfg[r.key] = r.value;
  }
  if (fread(&n_t, 4, 1, f) != 1 || n_t > (1u << 20)) return fail("bad tensor count");
  uint64_t file_size = 0;
  {
    const long pos = ftell(f);
    if (pos < 0 || fseek(f, 0, SEEK_END) != 0) return fail("seek");
    const long end = ftell(f);
    if (end < 0 || fseek(f, pos, SEEK_SET) != 0) return fail("seek");
    file_size = (uint64_t)end;
  }
  std::vector<TensorRec> recs(n_t);
  if (n_t && fread(recs.data(), sizeof(TensorRec), n_t, f) != n_t) return fail("truncated tensor table");
  for (auto& r : recs) {
    r.name[95] = 0;
    if (r.ndim > 4) return fail("bad ndim");
    HostTensor t;
    uint64_t n = 1;
    for (uint32_t d = 0; d < r.ndim; ++d) {
      // every dimension and the running product are bounded by the file size before anything is multiplied further or allocated
      if (r.dims[d] > file_size || (r.dims[d] != 0 && n > file_size / r.dims[d])) return fail(std::string("bad shape for ") + r.name);
      t.shape.push_back((int64_t)r.dims[d]);
      n *= r.dims[d];
    }
    if (n * 4 != r.nbytes || r.offset > file_size || r.nbytes > file_size - r.offset) return fail(std::string("size mismatch for ") + r.name);
    t.data.resize(n);
    if (fseek(f, (long)r.offset, SEEK_SET) != 0) return fail("seek");
    if (n && fread(t.data.data(), 4, (size_t)n, f) != (size_t)n) return fail(std::string("truncated data for ") + r.name);
    out->tensors[r.name] = std::move(t);
  }
  fclose(f);
  return true;
}

bool read_am_mvn(const std::string& path, std::vector<float>* means, std::vector<float>* vars, std::string* err) {
  std::ifstream in(path);
  if (!in.is_open()) { *err = "cannot open " + path; return false; }
  std::string line;
  auto split = [](const std::string& s) {
    std::istringstream iss(s);
    return std::vector<std::string>{std::istream_iterator<std::string>{iss}, std::istream_iterator<std::string>{}};
  };
  while (std::getline(in, line)) {
    auto it = split(line);
    if (it.empty()) continue;
    const bool shift = it[0] == "<AddShift>", rescale = it[0] == "<Rescale>";
    if (!shift && !rescale) continue;
    if (!std::getline(in, line)) break;
    auto v = split(line);
    if (!v.empty() && v[0] == "<LearnRateCoef>") {
      // tokens [3 .. size-2] hold the values: "<LearnRateCoef> 0 [ v0 ... vN ]"
      for (size_t j = 3; j + 1 < v.size(); ++j) (shift ? means : vars)->push_back(std::stof(v[j]));
    }
  }
  if (means->empty() || means->size() != vars->size()) { *err = "am.mvn: no <AddShift>/<Rescale> rows in " + path; return false; }
  return true;
}

static void append_utf8(std::string* s, uint32_t cp) {
  if (cp < 0x80) s->push_back((char)cp);
  else if (cp < 0x800) { s->push_back((char)(0xC0 | (cp >> 6))); s->push_back((char)(0x80 | (cp & 0x3F))); }
  else if (cp < 0x10000) {
    s->push_back((char)(0xE0 | (cp >> 12))); s->push_back((char)(0x80 | ((cp >> 6) & 0x3F))); s->push_back((char)(0x80 | (cp & 0x3F)));
  } else {
    s->push_back((char)(0xF0 | (cp >> 18))); s->push_back((char)(0x80 | ((cp >> 12) & 0x3F)));
    s->push_back((char)(0x80 | ((cp >> 6) & 0x3F))); s->push_back((char)(0x80 | (cp & 0x3F)));
  }
}

bool read_tokens_json(const std::string& path, std::vector<std::string>* tokens, std::string* err) {
  std::ifstream in(path, std::ios::binary);
  if (!in.is_open()) { *err = "cannot open " + path; return false; }
  std::string s((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
  size_t i = 0;
  auto ws = [&]() { while (i < s.size() && (s[i] == ' ' || s[i] == '\n' || s[i] == '\r' || s[i] == '\t')) ++i; };
  ws();
  if (i >= s.size() || s[i] != '[') { *err = "tokens.json: expected a JSON array"; return false; }
  ++i;
  for (;;) {
    ws();
    if (i >= s.size()) { *err = "tokens.json: unterminated array"; return false; }
    if (s[i] == ']') break;
    if (s[i] == ',') { ++i; continue; }
    if (s[i] != '"') { *err = "tokens.json: expected a string"; return false; }
    ++i;
    std::string tok;
    while (i < s.size() && s[i] != '"') {
      if (s[i] == '\\' && i + 1 < s.size()) {
        const char c = s[i + 1];
        i += 2;
        switch (c) {
          case 'n': tok.push_back('\n'); break;
          case 't': tok.push_back('\t'); break;
          case 'r': tok.push_back('\r'); break;
          case 'b': tok.push_back('\b'); break;
          case 'f': tok.push_back('\f'); break;
          case 'u': {
            if (i + 4 > s.size()) { *err = "tokens.json: bad \\u escape"; return false; }
            uint32_t cp = (uint32_t)std::stoul(s.substr(i, 4), nullptr, 16);
            i += 4;
            if (cp >= 0xD800 && cp < 0xDC00 && i + 6 <= s.size() && s[i] == '\\' && s[i + 1] == 'u') {
              const uint32_t lo = (uint32_t)std::stoul(s.substr(i + 2, 4), nullptr, 16);
              cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
              i += 6;
            }
            append_utf8(&tok, cp);
            break;
          }
          default: tok.push_back(c);
        }
      } else {
        tok.push_back(s[i++]);
      }
    }
    ++i;  // closing quote
    tokens->push_back(tok);
  }
  if (tokens->empty()) { *err = "tokens.json: empty"; return false; }
  return true;
}

bool read_config_yaml(const std::string& path, int* fs, std::string* lang, std::string* err) {
  *fs = 16000;
  *lang = "zh-cn";
  std::ifstream in(path);
  if (!in.is_open()) { *err = "cannot open " + path; return false; }
  std::string line;
  bool in_frontend = false;
  auto trim = [](std::string v) {
    const size_t c = v.find('#');
    if (c != std::string::npos) v = v.substr(0, c);
    size_t a = v.find_first_not_of(" \t\"'"), b = v.find_last_not_of(" \t\r\"'");
    return a == std::string::npos ? std::string() : v.substr(a, b - a + 1);
  };
  while (std::getline(in, line)) {
    if (line.empty() || line[0] == '#') continue;
    const bool top = line[0] != ' ' && line[0] != '\t';
    const size_t colon = line.find(':');
    if (colon == std::string::npos) continue;
    const std::string key = trim(line.substr(0, colon)), val = trim(line.substr(colon + 1));
    if (top) {
      in_frontend = key == "frontend_conf";
      if (key == "lang" && !val.empty()) *lang = val;
    } else if (in_frontend && key == "fs" && !val.empty()) {
      *fs = std::stoi(val);
    }
  }
  return true;
}

}  // namespace pf
