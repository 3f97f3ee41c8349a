// vsr_beam.cu -- the constraint mask of the beam search, on the device.
//
// Replaces the per-step, per-beam host loop of Model.fitfunc2 (reference
// src/visymre/architectures/model.py:385-411): `generated[i].cpu().tolist()` -- a device
// synchronisation per beam per decode step -- followed by the Python stack walk
// _analyze_prefix_tree_context (model.py:522-560) and the assembly of the -inf mask.  One
// thread per beam walks its prefix (<= length_eq tokens) with the same stack discipline; token
// sets are 64-bit masks (the vocabulary has 47-60 words).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vsr.h"

namespace {

constexpr int kMaxDepth = 128;  // prefix sequences are capped at length_eq (<= 100) tokens

__global__ void beam_mask_kernel(const int64_t* __restrict__ generated, int64_t ld, int beam, int cur_len,
                                 const float* __restrict__ beam_scores, vsr_beam_rules r, int n_words,
                                 float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= beam) return;
  float* row = out + (int64_t)i * n_words;
  for (int w = 0; w < n_words; ++w) row[w] = 0.0f;
  if (beam_scores[i] < -1e8f) return;  // dead beam: the reference skips it (model.py:387)

  // frames of the walk: operator (or -1), children still missing, constraints for the children
  int8_t op[kMaxDepth];
  int8_t missing[kMaxDepth];
  uint64_t cons[kMaxDepth];
  int depth = 1;
  op[0] = -1;
  missing[0] = 1;
  cons[0] = 0;
  const int64_t* seq = generated + (int64_t)i * ld;
  int start = (cur_len > 0 && seq[0] == r.start_id) ? 1 : 0;
  for (int t = start; t < cur_len && depth > 0; ++t) {
    const int tok = (int)seq[t];
    const uint64_t bit = (tok >= 0 && tok < 64) ? (1ull << tok) : 0ull;
    missing[depth - 1] -= 1;
    uint64_t inherited = cons[depth - 1];
    if (r.c_id >= 0 && op[depth - 1] == r.pow_id && missing[depth - 1] == 0)
      inherited |= 1ull << r.c_id;  // the exponent slot of a pow
    uint64_t for_children = inherited;
    if (bit & r.transcendental) for_children |= r.transcendental;
    if (r.pow_id >= 0 && tok == r.pow_id) for_children |= 1ull << r.pow_id;
    if ((bit & r.arity2) && depth < kMaxDepth) {
      op[depth] = (int8_t)tok;
      missing[depth] = 2;
      cons[depth] = for_children;
      ++depth;
    } else if ((bit & r.arity1) && depth < kMaxDepth) {
      op[depth] = (int8_t)tok;
      missing[depth] = 1;
      cons[depth] = for_children;
      ++depth;
    }
    while (depth > 0 && missing[depth - 1] == 0) --depth;
  }
  int valency = 0;
  for (int d = 0; d < depth; ++d) valency += missing[d];
  uint64_t forbidden = depth > 0 ? cons[depth - 1] : 0ull;
  if (r.c_id >= 0 && depth > 0 && op[depth - 1] == r.pow_id && missing[depth - 1] == 1)
    forbidden |= 1ull << r.c_id;
  // model.py:398-406
  if (valency >= r.length_eq - cur_len) forbidden |= r.all_ops;
  if (valency > 0) {
    if (r.finish_id >= 0) forbidden |= 1ull << r.finish_id;
    if (r.pad_id >= 0) forbidden |= 1ull << r.pad_id;
  }
  forbidden |= r.masked_vars;
  const float ninf = __int_as_float(0xff800000);
  for (int w = 0; w < n_words && w < 64; ++w)
    if ((forbidden >> w) & 1ull) row[w] = ninf;
}

// One decode step of the same walk with the frames kept in global memory (see vsr.h).
__global__ void beam_mask_step_kernel(int8_t* __restrict__ op_all, int8_t* __restrict__ missing_all,
                                      uint64_t* __restrict__ cons_all, int32_t* __restrict__ depth_all,
                                      int32_t* __restrict__ pos_all, int max_depth, const int64_t* __restrict__ tokens,
                                      int beam, const float* __restrict__ beam_scores, vsr_beam_rules r, int n_words,
                                      float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= beam) return;
  int8_t* op = op_all + (int64_t)i * max_depth;
  int8_t* missing = missing_all + (int64_t)i * max_depth;
  uint64_t* cons = cons_all + (int64_t)i * max_depth;
  int depth = depth_all[i];
  const int pos = pos_all[i];
  const int tok = (int)tokens[i];
  // the walk skips a leading start token and stops for good once the tree is complete
  if (!(pos == 0 && tok == r.start_id) && depth > 0) {
    const uint64_t bit = (tok >= 0 && tok < 64) ? (1ull << tok) : 0ull;
    missing[depth - 1] -= 1;
    uint64_t inherited = cons[depth - 1];
    if (r.c_id >= 0 && op[depth - 1] == r.pow_id && missing[depth - 1] == 0) inherited |= 1ull << r.c_id;
    uint64_t for_children = inherited;
    if (bit & r.transcendental) for_children |= r.transcendental;
    if (r.pow_id >= 0 && tok == r.pow_id) for_children |= 1ull << r.pow_id;
    if ((bit & r.arity2) && depth < max_depth) {
      op[depth] = (int8_t)tok;
      missing[depth] = 2;
      cons[depth] = for_children;
      ++depth;
    } else if ((bit & r.arity1) && depth < max_depth) {
      op[depth] = (int8_t)tok;
      missing[depth] = 1;
      cons[depth] = for_children;
      ++depth;
    }
    while (depth > 0 && missing[depth - 1] == 0) --depth;
    depth_all[i] = depth;
  }
  const int cur_len = pos + 1;
  pos_all[i] = cur_len;
  float* row = out + (int64_t)i * n_words;
  for (int w = 0; w < n_words; ++w) row[w] = 0.0f;
  if (beam_scores[i] < -1e8f) return;
  int valency = 0;
  for (int d = 0; d < depth; ++d) valency += missing[d];
  uint64_t forbidden = depth > 0 ? cons[depth - 1] : 0ull;
  if (r.c_id >= 0 && depth > 0 && op[depth - 1] == r.pow_id && missing[depth - 1] == 1) forbidden |= 1ull << r.c_id;
  if (valency >= r.length_eq - cur_len) forbidden |= r.all_ops;
  if (valency > 0) {
    if (r.finish_id >= 0) forbidden |= 1ull << r.finish_id;
    if (r.pad_id >= 0) forbidden |= 1ull << r.pad_id;
  }
  forbidden |= r.masked_vars;
  const float ninf = __int_as_float(0xff800000);
  for (int w = 0; w < n_words && w < 64; ++w)
    if ((forbidden >> w) & 1ull) row[w] = ninf;
}

}  // namespace

extern "C" int vsr_beam_mask_step(int8_t* op_dev, int8_t* missing_dev, uint64_t* cons_dev, int32_t* depth_dev,
                                  int32_t* pos_dev, int32_t max_depth, const int64_t* tokens_dev, int32_t beam,
                                  const float* beam_scores_dev, const vsr_beam_rules* rules, int32_t n_words,
                                  float* out_mask_dev, void* stream) {
  if (!op_dev || !missing_dev || !cons_dev || !depth_dev || !pos_dev || !tokens_dev || !beam_scores_dev || !rules ||
      !out_mask_dev || beam <= 0 || max_depth < 2 || max_depth > 128 || n_words <= 0 || n_words > 64)
    return VSR_EINVAL;
  const int threads = 64;
  beam_mask_step_kernel<<<(beam + threads - 1) / threads, threads, 0, (cudaStream_t)stream>>>(
      op_dev, missing_dev, cons_dev, depth_dev, pos_dev, max_depth, tokens_dev, beam, beam_scores_dev, *rules, n_words,
      out_mask_dev);
  return cudaGetLastError() == cudaSuccess ? VSR_OK : VSR_ECUDA;
}

extern "C" int vsr_beam_mask(const int64_t* generated_dev, int64_t ld, int32_t beam, int32_t cur_len,
                             const float* beam_scores_dev, const vsr_beam_rules* rules, int32_t n_words,
                             float* out_mask_dev, void* stream) {
  if (!generated_dev || !beam_scores_dev || !rules || !out_mask_dev || beam <= 0 || cur_len < 0 ||
      n_words <= 0 || n_words > 64 || ld < cur_len)
    return VSR_EINVAL;
  const int threads = 64;
  beam_mask_kernel<<<(beam + threads - 1) / threads, threads, 0, (cudaStream_t)stream>>>(
      generated_dev, ld, beam, cur_len, beam_scores_dev, *rules, n_words, out_mask_dev);
  return cudaGetLastError() == cudaSuccess ? VSR_OK : VSR_ECUDA;
}
