#include "text.h"

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <numeric>

namespace pf {
namespace host {

namespace {

const char kSubwordMark[] = "@@";
const char kBpeSpace[] = "\xE2\x96\x81";  // U+2581

bool IsControl(const std::string& w) { return w == "<s>" || w == "</s>" || w == "<unk>"; }
bool HasSubwordMark(const std::string& w) { return w.find(kSubwordMark) != std::string::npos; }
std::string DropMark(const std::string& w) { return w.size() >= 2 ? w.substr(0, w.size() - 2) : std::string(); }

std::string FixPronoun(const std::string& w) {  // Vocab::WordFormat
  if (w == "i") return "I";
  if (w == "i'm") return "I'm";
  if (w == "i've") return "I've";
  if (w == "i'll") return "I'll";
  return w;
}

std::string Join(const std::vector<std::string>& parts) {
  std::string out;
  for (const auto& p : parts) out += p;
  return out;
}

std::string F2S(float v) {  // std::to_string(float)
  char buf[64];
  snprintf(buf, sizeof(buf), "%f", (double)v);
  return buf;
}

}  // namespace

bool IsCjk(const std::string& s) {
  if (s.size() != 3) return false;
  const unsigned char a = s[0], b = s[1], c = s[2];
  if ((a & 0xf0) != 0xe0 || (b & 0xc0) != 0x80 || (c & 0xc0) != 0x80) return false;
  const int cp = ((a & 0x0f) << 12) | ((b & 0x3f) << 6) | (c & 0x3f);
  return cp >= 0x4E00 && cp <= 0x9FFF;
}

int Detokenizer::IdOf(const std::string& tok) const {
  for (size_t i = 0; i < tokens_.size(); ++i)
    if (tokens_[i] == tok) return (int)i;
  return -1;
}

std::vector<std::string> Detokenizer::ToPieces(const std::vector<int>& ids) const {
  std::vector<std::string> out;
  out.reserve(ids.size());
  for (int id : ids) out.push_back(tokens_[id]);
  return out;
}

std::string Detokenizer::ToText(const std::vector<int>& ids, const std::string& language) {
  bool ended = ended_on_english_word_;
  std::string text = ToText(ids, language, ended_on_english_word_, &ended);
  ended_on_english_word_ = ended;
  return text;
}

std::string Detokenizer::ToText(const std::vector<int>& ids, const std::string& language, bool after_english_word, bool* ended_out) const {
  std::vector<std::string> out;
  const bool lead_space = after_english_word;  // decided by the previous call
  bool ended_on_english_word_ = after_english_word;   // local copy of the member of the same name: left unchanged by an empty call
  struct Publish { bool* dst; const bool* src; ~Publish() { if (dst) *dst = *src; } } publish{ended_out, &ended_on_english_word_};
  const size_t n = ids.size();

  if (language == "en-bpe") {
    // sentencepiece pieces: a piece carrying U+2581 starts a new word
    std::string pending;
    auto flush = [&]() {
      if (pending.empty()) return;
      std::string w = FixPronoun(pending);
      out.push_back(out.empty() ? w : " " + w);
    };
    for (size_t i = 0; i < n; ++i) {
      const std::string& piece = tokens_[ids[i]];
      if (IsControl(piece)) continue;
      if (piece.find(kBpeSpace) != std::string::npos) {
        flush();
        pending = piece.size() >= 3 ? piece.substr(3) : std::string();
      } else {
        pending += piece;
      }
    }
    flush();
    return Join(out);
  }

  bool prev_english = false;   // previous emitted word was non-CJK
  size_t prev_len = 0;         // its byte length
  bool joining = false;        // inside an "@@" chain
  std::string chain;
  for (size_t i = 0; i < n; ++i) {
    std::string word = tokens_[ids[i]];
    if (IsControl(word)) continue;
    const bool marked = HasSubwordMark(word);
    if (marked) {
      const bool last = (i + 1 == n);
      const bool next_cjk = !last && IsCjk(tokens_[ids[i + 1]]);
      if (!next_cjk && !last) {        // chain continues
        chain += DropMark(word);
        joining = true;
        continue;
      }
      // the chain is cut short: by a CJK token (keep a separating space) or by the end of the sequence
      word = DropMark(word) + (next_cjk ? " " : "");
      if (joining) { word = chain + word; chain.clear(); joining = false; }
      if (!next_cjk) ended_on_english_word_ = false;
    } else if (joining) {
      word = chain + word;
      chain.clear();
      joining = false;
    }

    if (IsCjk(word)) {
      out.push_back(word);
      prev_english = false;
    } else {
      if (!prev_english) {
        if (lead_space) out.push_back(" ");
        out.push_back(word);
      } else {
        // single letters after a single letter are glued ("a b" -> "ab"); everything else is spaced
        if (prev_len > 1 || word.size() > 1) out.push_back(" ");
        out.push_back(word);
      }
      prev_len = word.size();
      prev_english = true;
    }
    ended_on_english_word_ = (i + 1 == n) && !IsCjk(word) && !marked;
  }
  return Join(out);
}

std::vector<Span> TimestampFromPeaks(std::vector<float>* us_alphas, const std::vector<float>& us_cif_peak,
                                     std::vector<std::string>* pieces, std::string* debug_str, float begin_time_ms,
                                     float total_offset) {
  std::vector<Span> result;
  if (pieces->empty()) return result;
  const float kEdge = 5.0f;        // START_END_THRESHOLD, upsampled frames
  const float kMaxDur = 30.0f;     // MAX_TOKEN_DURATION
  const float kRate = 10.0 * 6 / 1000 / 3;  // seconds per 3x-upsampled LFR frame
  const double kFire = 1.0 - 1e-4;
  std::vector<float> peak = us_cif_peak;
  const int frames = (int)peak.size();
  if (pieces->back() == "</s>") pieces->pop_back();
  if (pieces->empty()) return result;

  auto collect = [&](std::vector<float>* where) {
    where->clear();
    for (int i = 0; i < frames; ++i)
      if (peak[i] > kFire) where->push_back(i + total_offset);
  };
  std::vector<float> fire;
  collect(&fire);
  const int want = (int)pieces->size() + 1;
  if ((int)fire.size() != want) {
    // peaks and tokens disagree: renormalise the upsampled alphas so they integrate to #tokens + 1
    float total = std::accumulate(us_alphas->begin(), us_alphas->end(), 0.0f);
    const float scale = total / want;
    if (scale == 0) return result;
    peak.clear();
    float run = 0.0;
    for (float& a : *us_alphas) {
      a = a / scale;
      run += a;
      peak.push_back(run);
      if (run >= kFire) run -= kFire;
    }
    for (int k = (int)peak.size() - 1; run >= kFire && k >= 0; --k) {
      if (peak[k] < kFire) { peak[k] = run; run -= kFire; }
    }
    collect(&fire);
  }
  const int n_fire = (int)fire.size();
  if (n_fire == 0) return result;

  std::vector<std::string> labels;
  std::vector<Span> spans;
  if (fire[0] > kEdge) { labels.push_back("<sil>"); spans.push_back(Span(0.0f, fire[0] * kRate)); }
  for (int i = 0; i + 1 < n_fire; ++i) {
    labels.push_back((*pieces)[i]);
    const bool is_last = (i == n_fire - 2);
    if (is_last || kMaxDur < 0 || fire[i + 1] - fire[i] < kMaxDur) {
      spans.push_back(Span(fire[i] * kRate, fire[i + 1] * kRate));
    } else {  // over-long token: keep 30 frames, the rest becomes silence
      const float cut = fire[i] + kMaxDur;
      spans.push_back(Span(fire[i] * kRate, cut * kRate));
      spans.push_back(Span(cut * kRate, fire[i + 1] * kRate));
      labels.push_back("<sil>");
    }
  }
  if (spans.empty()) return result;
  if (frames - fire.back() > kEdge) {
    const float mid = (frames + fire.back()) / 2.0;
    spans.back().second = mid * kRate;
    spans.push_back(Span(mid * kRate, frames * kRate));
    labels.push_back("<sil>");
  } else {
    spans.back().second = frames * kRate;
  }
  if (begin_time_ms) {
    for (auto& s : spans) { s.first += begin_time_ms / 1000.0; s.second += begin_time_ms / 1000.0; }
  }
  for (size_t i = 0; i < labels.size(); ++i) {
    if (debug_str) *debug_str += labels[i] + " " + F2S(spans[i].first) + " " + F2S(spans[i].second) + ";";
    if (labels[i] != "<sil>") result.push_back(spans[i]);
  }
  return result;
}

std::string MergeWithStamps(const std::vector<std::string>& pieces, const std::vector<Span>& spans) {
  std::vector<std::string> out;
  std::vector<Span> merged;
  bool prev_english = false;
  bool joining = false;
  std::string chain;
  float open_begin = -1;  // begin time of a word whose "@@" chain is still open
  const size_t n = pieces.size();
  // the reference indexes timestamp_list[i] unchecked; stay defined when the lists disagree in length
  auto span_at = [&](size_t i) { return i < spans.size() ? spans[i] : (spans.empty() ? Span(0.f, 0.f) : spans.back()); };
  for (size_t i = 0; i < n; ++i) {
    std::string word = pieces[i];
    if (IsControl(word)) continue;
    if (HasSubwordMark(word)) {
      const bool cut = (i + 1 == n) || IsCjk(pieces[i + 1]);
      if (!cut) {
        chain += DropMark(word);
        if (!joining) open_begin = span_at(i).first;
        joining = true;
        continue;
      }
      word = DropMark(word) + " ";
      if (joining) { word = chain + word; chain.clear(); joining = false; }
    } else if (joining) {
      word = chain + word;
      chain.clear();
      joining = false;
    }
    if (IsCjk(word)) {
      out.push_back(word);
      merged.push_back(span_at(i));
      prev_english = false;
    } else {
      if (prev_english) out.push_back(" ");
      out.push_back(word);
      const float b = (open_begin == -1) ? span_at(i).first : open_begin;
      merged.push_back(Span(b, span_at(i).second));
      open_begin = -1;
      prev_english = true;
    }
  }
  std::string stamps;
  for (size_t i = 0; i < merged.size(); ++i) {
    stamps += F2S(merged[i].first) + ", " + F2S(merged[i].second);
    if (i + 1 != merged.size()) stamps += ",";
  }
  return Join(out) + " | " + stamps;
}

void StitchSegments(const std::vector<std::string>& msgs, const std::vector<float>& start_s, const std::string& lang,
                    std::string* text, std::string* stamp) {
  text->clear();
  stamp->clear();
  std::string acc = "[";
  for (size_t k = 0; k < msgs.size(); ++k) {
    const std::string& msg = msgs[k];
    if (msg.empty()) continue;
    const size_t bar = msg.find(" | ");
    const std::string head = bar == std::string::npos ? msg : msg.substr(0, bar);
    if (lang == "en-bpe" && !text->empty()) *text += " ";
    *text += head;
    if (bar == std::string::npos) continue;
    // "b, e,b, e" -> numbers in order
    std::vector<std::string> fields;
    std::string cur;
    for (char c : msg.substr(bar + 3)) {
      if (c == ',') { fields.push_back(cur); cur.clear(); } else cur.push_back(c);
    }
    fields.push_back(cur);
    if (fields.size() < 2) continue;
    for (size_t i = 0; i + 1 < fields.size(); i += 2) {
      const float b = std::stof(fields[i]) + start_s[k];
      const float e = std::stof(fields[i + 1]) + start_s[k];
      acc += "[" + std::to_string((int)(1000 * b)) + "," + std::to_string((int)(1000 * e)) + "],";
    }
  }
  if (acc != "[") {
    acc.erase(acc.size() - 1);
    *stamp = acc + "]";
  }
}


namespace {
// "[[b,e],[b,e]]" -> pairs; anything that is not a pair empties the result (ParseTimestamps, util.cpp:268-297)
std::vector<std::pair<int, int>> ParseStampList(const std::string& str) {
  std::vector<std::pair<int, int>> out;
  size_t pos = str.empty() ? 0 : 1;   // the opening '['
  while (pos < str.size()) {
    size_t close = str.find(']', pos);
    const bool had_delim = close != std::string::npos;
    if (!had_delim) close = str.size();
    std::string seg = str.substr(pos, close - pos);
    if (!seg.empty()) seg.erase(0, 1);   // its own '['
    std::vector<int> nums;
    size_t a = 0;
    while (a <= seg.size() && !seg.empty()) {
      size_t b = seg.find(',', a);
      if (b == std::string::npos) b = seg.size();
      nums.push_back(atoi(seg.substr(a, b - a).c_str()));
      a = b + 1;
      if (b == seg.size()) break;
    }
    if (nums.size() != 2) return std::vector<std::pair<int, int>>();
    out.emplace_back(nums[0], nums[1]);
    pos = had_delim ? close + 2 : str.size();   // the ']' and the ',' (or the closing ']') after it
  }
  return out;
}

// every BYTE of s occurs among the bytes of "，。？、,?" (TimestampIsPunctuation(const std::string&), util.cpp:257-266)
bool AllPunctuationBytes(const std::string& s) {
  static const std::string set = "\xEF\xBC\x8C\xE3\x80\x82\xEF\xBC\x9F\xE3\x80\x81,?";
  for (char c : s)
    if (set.find(c) == std::string::npos) return false;
  return true;
}

bool PunctuationCodePoint(uint32_t u) {   // TimestampIsPunctuation(U16CHAR_T&), util.cpp:307-318
  if (u == 0x26 || u == 0x27 || u == 0x2D) return false;
  return (u >= 0x21 && u <= 0x2F) || (u >= 0x3A && u <= 0x40) || (u >= 0x5B && u <= 0x60) || (u >= 0x7B && u <= 0x7E) ||
         (u >= 0x2000 && u <= 0x206F) || (u >= 0x3000 && u <= 0x303F);
}

// CJK characters, digits and punctuation marks stand alone, other characters accumulate into words, spaces separate
// (TimestampSplitChiEngCharacters, util.cpp:320-366)
std::vector<std::string> SplitForStamps(const std::string& s) {
  std::vector<std::string> out;
  std::string word;
  size_t i = 0;
  while (i < s.size()) {
    const unsigned char c = (unsigned char)s[i];
    size_t n = 1;
    uint32_t u = c;
    if ((c & 0xF0) == 0xE0 && i + 2 < s.size()) { n = 3; u = ((c & 0x0F) << 12) | (((unsigned char)s[i + 1] & 0x3F) << 6) | ((unsigned char)s[i + 2] & 0x3F); }
    else if ((c & 0xE0) == 0xC0 && i + 1 < s.size()) { n = 2; u = ((c & 0x1F) << 6) | ((unsigned char)s[i + 1] & 0x3F); }
    const std::string ch = s.substr(i, n);
    i += n;
    const bool cjk = (u >= 0x4e00 && u <= 0x9fff) || (u >= 0x3400 && u <= 0x4dff);
    if (cjk || (u >= '0' && u <= '9') || PunctuationCodePoint(u)) {
      if (!word.empty()) { out.push_back(word); word.clear(); }
      out.push_back(ch);
    } else if (u == 0x20) {
      if (!word.empty()) { out.push_back(word); word.clear(); }
    } else {
      word += ch;
    }
  }
  if (!word.empty()) out.push_back(word);
  return out;
}

std::string StampListJson(const std::vector<std::pair<int, int>>& v) {
  if (v.empty()) return "[]";
  std::string s = "[";
  for (size_t i = 0; i < v.size(); ++i) {
    s += "[" + std::to_string(v[i].first) + "," + std::to_string(v[i].second) + "]";
    if (i + 1 < v.size()) s += ",";
  }
  return s + "]";
}
}  // namespace

std::string SentenceStamps(const std::string& text, const std::string& stamp) {
  const std::vector<std::string> chars = SplitForStamps(text);
  const std::vector<std::pair<int, int>> ts = ParseStampList(stamp);
  size_t it = 0;
  int start = -1, end = -1;
  std::string seg_text, out;
  std::vector<std::pair<int, int>> seg;
  auto emit = [&](const std::string& punc) {
    if (!seg.empty()) { start = seg.front().first; end = seg.back().second; }
    out += "{\"text_seg\":\"" + seg_text + "\",\"punc\":\"" + punc + "\",\"start\":" + std::to_string(start) + ",\"end\":" + std::to_string(end) +
           ",\"ts_list\":" + StampListJson(seg) + "}";
  };
  for (size_t k = 0; k < chars.size(); ++k) {
    if (AllPunctuationBytes(chars[k])) {
      emit(chars[k]);
      if (k + 1 != chars.size()) out += ",";
      seg_text.clear();
      seg.clear();
      start = 0;
      end = 0;
    } else if (it < ts.size()) {
      seg_text += seg_text.empty() ? chars[k] : " " + chars[k];
      seg.push_back(ts[it++]);
    }
  }
  if (!seg.empty()) emit("");
  return "[" + out + "]";
}

}  // namespace host
}  // namespace pf
