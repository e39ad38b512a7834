// Host-side text and timestamp post-processing of the Paraformer::Forward path, mirroring the reference's
// behaviour byte for byte:
//   Detokenizer::ToText        Vocab::Vector2StringV2   onnxruntime/src/vocab.cpp:164-305
//   Detokenizer::ToPieces      Vocab::Vector2String     onnxruntime/src/vocab.cpp:98-104
//   TimestampFromPeaks         TimestampOnnx            onnxruntime/src/util.cpp:838-963
//   MergeWithStamps            PostProcess              onnxruntime/src/util.cpp:720-836
//   StitchSegments             FunOfflineInferBuffer    onnxruntime/src/funasrruntime.cpp:291-316
//   SentenceStamps             TimestampSentence        onnxruntime/src/util.cpp:569-637
#pragma once
#include <string>
#include <utility>
#include <vector>

namespace pf {
namespace host {

bool IsCjk(const std::string& s);  // one 3-byte UTF-8 code point in U+4E00..U+9FFF (vocab.cpp:131-141)

class Detokenizer {
 public:
  explicit Detokenizer(std::vector<std::string> tokens) : tokens_(std::move(tokens)) {}
  const std::vector<std::string>& tokens() const { return tokens_; }
  int IdOf(const std::string& tok) const;
  std::vector<std::string> ToPieces(const std::vector<int>& ids) const;
  // Stateful like the reference: whether the previous call ended on a complete English word decides if the
  // next call starts with a space (Vocab::last_is_complete_english_, vocab.cpp:177,259-261,283-288).
  std::string ToText(const std::vector<int>& ids, const std::string& language);
  // The same text for an explicit incoming state; *ended_out is the state the call leaves behind.  Lets a caller build the two
  // possible texts of a segment ahead of time (in parallel) and pick one when the previous segment's end state is known.
  std::string ToText(const std::vector<int>& ids, const std::string& language, bool after_english_word, bool* ended_out) const;
  bool ended_on_english_word() const { return ended_on_english_word_; }
  void set_ended_on_english_word(bool v) { ended_on_english_word_ = v; }
  void ResetState() { ended_on_english_word_ = false; }

 private:
  std::vector<std::string> tokens_;
  bool ended_on_english_word_ = false;
};

typedef std::pair<float, float> Span;  // seconds

// us_alphas is rescaled in place in the peak-count-mismatch branch, pieces loses a trailing "</s>",
// exactly as the reference mutates its arguments.  Returns the spans of the non-<sil> tokens.
std::vector<Span> TimestampFromPeaks(std::vector<float>* us_alphas, const std::vector<float>& us_cif_peak,
                                     std::vector<std::string>* pieces, std::string* debug_str, float begin_time_ms = 0.0f,
                                     float total_offset = -1.5f);

// "text | b0, e0,b1, e1" (std::to_string floats), merging "@@" sub-words and their spans.
std::string MergeWithStamps(const std::vector<std::string>& pieces, const std::vector<Span>& spans);

// Joins per-segment Forward() strings (original time order) into the final text and "[[b,e],...]" in ms.
void StitchSegments(const std::vector<std::string>& msgs, const std::vector<float>& start_s, const std::string& lang,
                    std::string* text, std::string* stamp);

// Punctuated text + "[[b,e],...]" (ms) -> the stamp_sents JSON array: one {"text_seg","punc","start","end","ts_list"} object per
// punctuation-terminated sentence, plus a last one with "punc":"" for text after the final punctuation mark (BMP text).
std::string SentenceStamps(const std::string& text, const std::string& stamp);

}  // namespace host
}  // namespace pf
