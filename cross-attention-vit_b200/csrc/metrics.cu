// cavit-sm100 — per-step classification metrics without host synchronisation.
//
// The reference logs seven metrics after EVERY training / validation step (`log_stats`,
// /root/reference/model_cross.py:243-255 → `compute_metrics`, /root/reference/utils.py:18-62, and
// `torchmetrics.functional.auroc`): six torchmetrics objects are built, each result is read back with `.item()` (a host
// synchronisation per metric, per step), and Lightning averages the per-batch values over the epoch, weighted by batch size
// (`on_epoch=True`). Here one single-block launch per step takes the logits and labels the step already has on the device and
// adds   B * (accuracy, precision, recall, specificity, F1, NPV, AUROC, loss),  B  and  1   to a 16-double accumulator
// (10 values + 2 words of scratch + 4 reserved);
// the host reads it once per epoch (cavit/metrics.py).
//
// Definitions (torchmetrics binary metrics on `argmax(logits, 1)`; 0 / 0 = 0 everywhere, `_safe_divide`; fp32 divisions):
//   accuracy (tp+tn)/B, precision tp/(tp+fp), recall tp/(tp+fn), specificity tn/(tn+fp), F1 2tp/(2tp+fn+fp),
//   NPV tn/(tn+fn); AUROC of softmax(logits)[:, 1] with ties counted half (the trapezoid over distinct thresholds equals
//   the Mann-Whitney statistic), 0 when the batch holds one class only. The pair count is exact integer arithmetic.
#include "common.cuh"
#include "internal.h"

namespace cavit {

constexpr int METRICS_MAX_B = 8192;
constexpr int METRICS_ROWS_PER_BLOCK = 64;  // positives-candidates (rows i of the pair count) per block

// grid = ceil(B / 64) blocks (one for the reference's batch sizes). Every block recomputes the B probabilities (cheap) and
// counts the pairs of its own slice of rows; the pair counts meet in accum[10] (as an integer) and the block that draws the
// last ticket (accum[11]) finishes the metrics and clears both words for the next launch.
__global__ void __launch_bounds__(256)
batch_metrics_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, const float* __restrict__ loss,
                     double* __restrict__ accum, int B) {
  __shared__ float prob[METRICS_MAX_B];
  __shared__ uint8_t pos[METRICS_MAX_B];
  __shared__ unsigned long long cnt[5];  // tn, fp, fn, tp, 2 * U of this block
  __shared__ bool last;
  if (threadIdx.x < 5) cnt[threadIdx.x] = 0ull;
  __syncthreads();
  unsigned c[4] = {0u, 0u, 0u, 0u};
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    const float z0 = logits[2 * i], z1 = logits[2 * i + 1];
    const int pred = z1 > z0;  // torch.argmax keeps the first maximum
    const int y = labels[i] != 0;
    const float m = fmaxf(z0, z1);
    const float e0 = expf(z0 - m), e1 = expf(z1 - m);
    prob[i] = e1 / (e0 + e1);
    pos[i] = static_cast<uint8_t>(y);
    ++c[2 * y + pred];
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    unsigned v = c[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&cnt[k], static_cast<unsigned long long>(v));
  }
  __syncthreads();
  // pairs (i, j), i a positive of this block's slice, j any negative: 2 for prob[i] > prob[j], 1 for a tie
  const int i_lo = blockIdx.x * METRICS_ROWS_PER_BLOCK, i_hi = min(B, i_lo + METRICS_ROWS_PER_BLOCK);
  unsigned long long u2 = 0ull;
  for (int i = i_lo; i < i_hi; ++i) {
    if (!pos[i]) continue;  // block-uniform
    const float p = prob[i];
    unsigned w = 0u;
    for (int j = threadIdx.x; j < B; j += blockDim.x) {
      const float q = prob[j];
      w += pos[j] ? 0u : (p > q ? 2u : (p == q ? 1u : 0u));
    }
    u2 += w;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) u2 += __shfl_xor_sync(0xffffffffu, u2, o);
  if ((threadIdx.x & 31) == 0 && u2) atomicAdd(&cnt[4], u2);
  __syncthreads();
  unsigned long long* pairs = reinterpret_cast<unsigned long long*>(accum + 10);
  unsigned long long* ticket = reinterpret_cast<unsigned long long*>(accum + 11);
  if (threadIdx.x == 0) {
    if (cnt[4]) atomicAdd(pairs, cnt[4]);
    __threadfence();
    last = atomicAdd(ticket, 1ull) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last || threadIdx.x != 0) return;
  __threadfence();
  const unsigned long long u_all = atomicExch(pairs, 0ull);
  *ticket = 0ull;
  const float tn = static_cast<float>(cnt[0]), fp = static_cast<float>(cnt[1]), fn = static_cast<float>(cnt[2]),
              tp = static_cast<float>(cnt[3]);
  auto safe = [](float a, float b) { return b != 0.f ? a / b : 0.f; };
  const float npos = tp + fn, nneg = tn + fp;
  float m[8];
  m[0] = safe(tp + tn, tp + tn + fp + fn);
  m[1] = safe(tp, tp + fp);
  m[2] = safe(tp, tp + fn);
  m[3] = safe(tn, tn + fp);
  m[4] = safe(2.f * tp, 2.f * tp + fn + fp);
  m[5] = safe(tn, tn + fn);
  m[6] = (npos > 0.f && nneg > 0.f)
             ? static_cast<float>(static_cast<double>(u_all) /
                                  (2.0 * static_cast<double>(cnt[2] + cnt[3]) * static_cast<double>(cnt[0] + cnt[1])))
             : 0.f;
  m[7] = loss ? *loss : 0.f;
  const double w = static_cast<double>(B);
#pragma unroll
  for (int k = 0; k < 8; ++k) accum[k] += w * static_cast<double>(m[k]);
  accum[8] += w;
  accum[9] += 1.0;
}

}  // namespace cavit

using namespace cavit;

extern "C" int cavit_batch_metrics(const float* logits, const int64_t* labels, const float* loss, double* accum, int32_t B,
                                   int32_t classes, void* stream) {
  if (!logits || !labels || !accum) return fail(CAVIT_E_BADARG, "cavit_batch_metrics: null pointer");
  if (classes != 2) return fail(CAVIT_E_UNSUPPORTED_SHAPE, "cavit_batch_metrics: binary metrics need 2 classes, got %d", classes);
  if (B < 1 || B > METRICS_MAX_B)
    return fail(CAVIT_E_UNSUPPORTED_SHAPE, "cavit_batch_metrics: batch %d outside 1..%d", B, METRICS_MAX_B);
  batch_metrics_kernel<<<(B + METRICS_ROWS_PER_BLOCK - 1) / METRICS_ROWS_PER_BLOCK, 256, 0, as_stream(stream)>>>(
      logits, labels, loss, accum, B);
  count_launch();
  return check_launch("cavit_batch_metrics");
}
