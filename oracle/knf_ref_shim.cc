// ORACLE (test infrastructure only).  Thin extern "C" wrapper around the REFERENCE's own
// kaldi-native-fbank, compiled in place from /root/reference by oracle/Makefile into
// oracle/_ref/libknf_ref.so.  It performs exactly the calls Paraformer::FbankKaldi makes
// (onnxruntime/src/paraformer.cpp:309-323) with the options of Paraformer::InitAsr (:24-31).
// No reference source is copied into this repository.
#include <cstdint>
#include <vector>
#include "kaldi-native-fbank/csrc/online-feature.h"
#include "kaldi-native-fbank/csrc/rfft.h"

extern "C" int knf_ref_fbank(const float *pcm, int n, float *out, int max_frames) {
  knf::FbankOptions opts;
  opts.frame_opts.dither = 0;
  opts.mel_opts.num_bins = 80;
  opts.frame_opts.samp_freq = 16000;
  opts.frame_opts.window_type = "hamming";
  opts.frame_opts.frame_shift_ms = 10;
  opts.frame_opts.frame_length_ms = 25;
  opts.energy_floor = 0;
  opts.mel_opts.debug_mel = false;
  knf::OnlineFbank fbank(opts);
  std::vector<float> buf(n);
  for (int32_t i = 0; i != n; ++i) buf[i] = pcm[i] * 32768;
  fbank.AcceptWaveform(16000, buf.data(), buf.size());
  int32_t frames = fbank.NumFramesReady();
  for (int32_t i = 0; i < frames && i < max_frames; ++i) {
    const float *frame = fbank.GetFrame(i);
    for (int b = 0; b < 80; ++b) out[(int64_t)i * 80 + b] = frame[b];
  }
  return frames;
}

extern "C" void knf_ref_rfft(float *in_out, int n) {
  knf::Rfft fft(n);
  fft.Compute(in_out);
}
