#include "logprob_adapter.h"

#include <cmath>

namespace pf {
namespace host {

namespace {
// log of the probability left for each class that is not listed: log((1 - sum_k p) / (vocab - k)), floored so that a row whose
// listed classes already hold all the mass (to float precision) stays finite
float RestLogProb(const float* lp, int k, int vocab) {
  double mass = 0.0;
  for (int j = 0; j < k; ++j) mass += std::exp((double)lp[j]);
  const int rest = vocab - k;
  if (rest <= 0) return -INFINITY;
  const double left = 1.0 - mass;
  const double floor_p = 1e-30;
  return (float)std::log((left > floor_p ? left : floor_p) / rest);
}
}  // namespace

void ExpandPrunedPosteriors(const float* topk_logprob, const int32_t* topk_ids, int rows, int k, int vocab, std::vector<float>* dense) {
  dense->assign((size_t)rows * vocab, 0.f);
  for (int r = 0; r < rows; ++r) {
    const float* lp = topk_logprob + (size_t)r * k;
    const int32_t* id = topk_ids + (size_t)r * k;
    float* out = dense->data() + (size_t)r * vocab;
    const float rest = RestLogProb(lp, k, vocab);
    for (int c = 0; c < vocab; ++c) out[c] = rest;
    for (int j = 0; j < k; ++j)
      if (id[j] >= 0 && id[j] < vocab) out[id[j]] = lp[j];
  }
}

float PrunedLogLikelihood(const float* topk_logprob, const int32_t* topk_ids, int row, int k, int vocab, int id) {
  const float* lp = topk_logprob + (size_t)row * k;
  const int32_t* ids = topk_ids + (size_t)row * k;
  for (int j = 0; j < k; ++j)
    if (ids[j] == id) return lp[j];
  return RestLogProb(lp, k, vocab);
}

}  // namespace host
}  // namespace pf
