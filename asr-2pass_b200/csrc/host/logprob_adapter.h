// Pruned posteriors -> the rows the reference's log-prob consumers read.
//
// WfstDecoder::Search (onnxruntime/src/wfst-decoder.cpp:27-57) and CtcPrefixDecoder::CtcSearch (ctc-prefix-decoder.cpp:157) take the
// graph's log_softmax output as dense rows [len, vocab] (the reference hands them `floatData` of the logits tensor,
// paraformer.cpp:566,575).  The B200 engine does not ship [L, 8404] floats per segment to the host; with the engine option
// "logprob_topk" = k it returns, per token, the k largest log-softmax values with their ids (b200pf_result.topk_*).  This adapter
// rebuilds a dense, properly normalised row from them: the k listed classes keep their exact log-probabilities, the probability
// mass that is left, 1 - sum_k p, is spread evenly over the other vocab - k classes.  A beam search with beam <= k never looks
// below the listed classes (CtcPrefixDecoder's first_beam_size is 10), so its result is the one the full rows give.
#pragma once
#include <stdint.h>

#include <vector>

namespace pf {
namespace host {

// dense receives rows * vocab floats.  Ids outside [0, vocab) are ignored.
void ExpandPrunedPosteriors(const float* topk_logprob, const int32_t* topk_ids, int rows, int k, int vocab, std::vector<float>* dense);
// log-probability of `id` in row `row` of the pruned posterior (what kaldi's DecodableInterface::LogLikelihood(frame, id) asks for)
float PrunedLogLikelihood(const float* topk_logprob, const int32_t* topk_ids, int row, int k, int vocab, int id);

}  // namespace host
}  // namespace pf
