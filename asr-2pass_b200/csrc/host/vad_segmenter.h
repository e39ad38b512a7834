// VadSegmenter — per-frame silence probabilities -> speech segments, the job funasr::E2EVadModel does in the reference
// (onnxruntime/src/e2e-vad.h, driven by FsmnVadOnline::Infer / Audio::CutSplit, audio.cpp:1172-1226).  Written from the
// reference's behaviour for whole recordings (one pass over all frames, the last one flagged final) with the reference's
// default VADXOptions (multiple-utterance mode, 200 ms window with 150 / 150 ms thresholds, 200 ms look-back at the start
// point, 100 ms look-ahead at the end point, decibel / SNR gates off: decibel_thres = snr_thres = -100).  Pinned against the
// reference's own compiled E2EVadModel, in its chunked online form and in one offline call (tests/test_vad.py,
// tests/golden/vad_segments_golden.npz).
#pragma once
#include <utility>
#include <vector>

namespace pf {
namespace host {

struct VadOptions {
  int max_end_silence_ms = 800;       // vad_tail_sil / max_end_silence_time (FunOfflineInferBuffer's vad_tail_sil)
  int max_single_segment_ms = 15000;  // vad_max_len
  float speech_noise_thres = 0.8f;    // model_conf.speech_noise_thres
};

// sil_prob[t] = probability of pdf 0 (silence) of 10 ms frame t.  Returns [start_ms, end_ms) pairs in time order.
std::vector<std::pair<int, int>> SegmentVad(const float* sil_prob, int n_frames, const VadOptions& opt);

// The same state machine fed chunk by chunk -- what the 2-pass stream does (FsmnVadOnline::Forward per chunk, Audio::Split,
// audio.cpp:1242-1370): Push() takes the silence probabilities of the NEXT frames and returns the segments that closed on
// them; the frame flagged final closes an open segment like the last frame of SegmentVad.  Fed any chunking of the same
// scores it returns, in total, exactly SegmentVad's segments (tests/test_vad.py).
class StreamingVad {
 public:
  explicit StreamingVad(const VadOptions& opt);
  ~StreamingVad();
  StreamingVad(const StreamingVad&) = delete;
  StreamingVad& operator=(const StreamingVad&) = delete;
  std::vector<std::pair<int, int>> Push(const float* sil_prob, int n_frames, bool last_is_final);
  void SetOptions(int max_end_silence_ms, int max_single_segment_ms);   // VadModel::SetConfig, per call like the reference
  void Reset();                 // a new stream (Audio::ResetIndex after input_finished)
  int frames() const;           // frames consumed so far
  int open_start_ms() const;    // start of the segment currently open, or -1

 private:
  struct Impl;
  Impl* impl_;
};

}  // namespace host
}  // namespace pf
