#include "vad_segmenter.h"

#include <algorithm>
#include <cmath>

namespace pf {
namespace host {

namespace {

constexpr int kFrameMs = 10;
constexpr int kWindowFrames = 20;        // window_size_ms 200
constexpr int kToSpeechFrames = 15;      // sil_to_speech_time_thres 150
constexpr int kToSilFrames = 15;         // speech_to_sil_time_thres 150
constexpr int kSpeechToSilMs = 150;
constexpr int kLookbackStartMs = 200;    // lookback_time_start_point
constexpr int kLookaheadEndMs = 100;     // lookahead_time_end_point
constexpr int kStartLatency = kWindowFrames + kLookbackStartMs / kFrameMs;   // LatencyFrmNumAtStartPoint

enum class Change { SilToSpeech, SpeechToSil, SilToSil, SpeechToSpeech };
enum class Phase { Searching, InSpeech, Ended };

// sliding majority vote over the last 20 frame decisions (WindowDetector::DetectOneFrame, e2e-vad.h:234-260)
struct Window {
  int ring[kWindowFrames];
  int sum = 0, pos = 0;
  bool speech = false;
  Window() { Reset(); }
  void Reset() { std::fill(ring, ring + kWindowFrames, 0); sum = 0; pos = 0; speech = false; }
  Change Push(bool frame_is_speech) {
    sum += (frame_is_speech ? 1 : 0) - ring[pos];
    ring[pos] = frame_is_speech ? 1 : 0;
    pos = (pos + 1) % kWindowFrames;
    if (!speech && sum >= kToSpeechFrames) { speech = true; return Change::SilToSpeech; }
    if (speech && sum <= kToSilFrames) { speech = false; return Change::SpeechToSil; }
    return speech ? Change::SpeechToSpeech : Change::SilToSil;
  }
};

struct Machine {
  const VadOptions& opt;
  Window win;
  Phase phase = Phase::Searching;
  int quiet_run = 0;            // consecutive SilToSil frames
  int last_speech = 0;          // latest frame confirmed as speech
  int last_quiet = -1;          // latest frame confirmed as silence
  int seg_start = -1;           // confirmed start frame of the open segment
  int ends_seen = 0;            // segments closed so far (never reset)
  int buf_front = 0;            // first frame still held in the (virtual) audio buffer
  struct Seg { int start_ms, end_ms; bool closed; };
  std::vector<Seg> out;

  explicit Machine(const VadOptions& o) : opt(o) {}

  // PopDataToOutputBuf: frame `f` joins the open segment, or opens a new one
  void Emit(int f, bool opens, bool closes = false) {
    buf_front = std::max(buf_front, f);
    if (out.empty() || opens) out.push_back(Seg{f * kFrameMs, f * kFrameMs, false});
    buf_front += 1;
    out.back().end_ms = (f + 1) * kFrameMs;
    if (closes) out.back().closed = true;
  }
  void Speech(int f) { last_speech = f; Emit(f, false); }
  void Quiet(int f) {
    last_quiet = f;
    if (phase == Phase::Searching) buf_front = std::max(buf_front, f);
  }
  void Open(int f, bool fake) {
    if (seg_start == -1) seg_start = f;
    if (!fake && phase == Phase::Searching) Emit(seg_start, true);
  }
  void Close(int f, bool fake) {
    for (int u = last_speech + 1; u < f; ++u) Speech(u);
    if (!fake) Emit(f, false, true);
    ++ends_seen;
  }
  void ResetDetection() {
    quiet_run = 0; last_speech = 0; last_quiet = -1; seg_start = -1;
    phase = Phase::Searching;
    win.Reset();
  }
  bool TooLong(int f) const { return f - seg_start + 1 > opt.max_single_segment_ms / kFrameMs; }

  // one frame inside an open segment: close on the length cap, extend, or close at the very last frame
  void Continue(int f, bool final_frame) {
    if (TooLong(f)) { Close(f, false); phase = Phase::Ended; }
    else if (!final_frame) Speech(f);
    else { Close(f, false); phase = Phase::Ended; }
  }

  void Step(int f, bool frame_is_speech, bool final_frame) {
    const Change ch = win.Push(frame_is_speech);
    if (ch == Change::SilToSpeech) {
      quiet_run = 0;
      if (phase == Phase::Searching) {
        const int s = std::max(buf_front, f - kStartLatency);
        Open(s, false);
        phase = Phase::InSpeech;
        for (int u = s + 1; u <= f; ++u) Speech(u);
      } else if (phase == Phase::InSpeech) {
        for (int u = last_speech + 1; u < f; ++u) Speech(u);
        Continue(f, final_frame);
      }
    } else if (ch == Change::SpeechToSil || ch == Change::SpeechToSpeech) {
      quiet_run = 0;
      if (phase == Phase::InSpeech) Continue(f, final_frame);
    } else {  // SilToSil
      ++quiet_run;
      if (phase == Phase::Searching) {
        if (final_frame && ends_seen == 0) {
          // a recording without any speech: the reference closes a zero-length fake segment so that callers see an end
          for (int u = last_quiet + 1; u < f; ++u) Quiet(u);
          Open(0, true);
          Close(0, true);
          phase = Phase::Ended;
        } else if (f >= kStartLatency) {
          Quiet(f - kStartLatency);
        }
      } else if (phase == Phase::InSpeech) {
        const int end_sil_ms = opt.max_end_silence_ms - kSpeechToSilMs;
        if (quiet_run * kFrameMs >= end_sil_ms) {
          int back = end_sil_ms / kFrameMs - kLookaheadEndMs / kFrameMs - 1;
          back = std::max(0, back);
          Close(f - back, false);
          phase = Phase::Ended;
        } else if (TooLong(f)) {
          Close(f, false);
          phase = Phase::Ended;
        } else if (!final_frame) {
          if (quiet_run <= kLookaheadEndMs / kFrameMs) Speech(f);
        } else {
          Close(f, false);
          phase = Phase::Ended;
        }
      }
    }
    if (phase == Phase::Ended) ResetDetection();
  }
};

// E2EVadModel::GetFrameState with the decibel / SNR gates off: speech iff exp(log(1 - p)) >= exp(log(p)) + thres, evaluated
// in float exactly as written there (the round trip through log / exp is kept: it decides frames that sit on the threshold)
bool FrameIsSpeech(float p_sil, float thres) {
  const float noise_prob = std::log(p_sil) * 1.0f;
  const float speech = 1.0f - p_sil;
  const float speech_prob = std::log(speech);
  return std::exp(speech_prob) >= std::exp(noise_prob) + thres;
}

}  // namespace

std::vector<std::pair<int, int>> SegmentVad(const float* sil_prob, int n_frames, const VadOptions& opt) {
  Machine m(opt);
  for (int f = 0; f < n_frames; ++f) m.Step(f, FrameIsSpeech(sil_prob[f], opt.speech_noise_thres), f == n_frames - 1);
  std::vector<std::pair<int, int>> segs;
  // A segment whose start is confirmed on the very last frame never gets an end point; Audio::CutSplit pairs start and end
  // points and therefore drops it (audio.cpp:1198-1226).
  for (const auto& s : m.out)
    if (s.closed) segs.emplace_back(s.start_ms, s.end_ms);
  return segs;
}

// ---- the same machine fed incrementally (the 2-pass stream: FsmnVadOnline + Audio::Split, audio.cpp:1242-1370) ----
struct StreamingVad::Impl {
  VadOptions opt;
  Machine m;
  int next_frame = 0;
  size_t reported = 0;     // segments of m.out already handed to the caller
  explicit Impl(const VadOptions& o) : opt(o), m(opt) {}
};

StreamingVad::StreamingVad(const VadOptions& opt) : impl_(new Impl(opt)) {}
StreamingVad::~StreamingVad() { delete impl_; }

void StreamingVad::Reset() {
  const VadOptions o = impl_->opt;
  delete impl_;
  impl_ = new Impl(o);
}

void StreamingVad::SetOptions(int max_end_silence_ms, int max_single_segment_ms) {
  impl_->opt.max_end_silence_ms = max_end_silence_ms;       // Machine holds a reference to impl_->opt
  impl_->opt.max_single_segment_ms = max_single_segment_ms;
}

int StreamingVad::frames() const { return impl_->next_frame; }

int StreamingVad::open_start_ms() const {
  const Machine& m = impl_->m;
  if (!m.out.empty() && !m.out.back().closed && impl_->reported < m.out.size()) return m.out.back().start_ms;
  return -1;
}

std::vector<std::pair<int, int>> StreamingVad::Push(const float* sil_prob, int n_frames, bool last_is_final) {
  Machine& m = impl_->m;
  for (int i = 0; i < n_frames; ++i) {
    const int f = impl_->next_frame++;
    m.Step(f, FrameIsSpeech(sil_prob[i], impl_->opt.speech_noise_thres), last_is_final && i == n_frames - 1);
  }
  std::vector<std::pair<int, int>> segs;
  while (impl_->reported < m.out.size() && m.out[impl_->reported].closed) {
    segs.emplace_back(m.out[impl_->reported].start_ms, m.out[impl_->reported].end_ms);
    ++impl_->reported;
  }
  if (last_is_final) {
    // an unclosed tail (a start confirmed on the very last frame) is dropped, as in SegmentVad
    impl_->reported = m.out.size();
  }
  return segs;
}

}  // namespace host
}  // namespace pf
