// On-device versions of the two host-side steps that sit in front of every pre-training batch:
//   * span masking        -- compute_mask_indices, WavLM/WavLM.py:35-159 (static span length, overlapping spans allowed: the
//                            released recipes; called from apply_mask :271-287)
//   * utterance mixing    -- src/fairseq/data/audio/utterance_mixing_dataset.py:373-438 (utterance branch)
// The reference runs both on the host with numpy's RNG (a `.item()` device sync per row in the sampler, python loops over 16 kHz
// samples in the mixer).  numpy's streams cannot be reproduced on the device, so these kernels use the library's counter-based hash
// (dropout.cuh) and are held to the reference STATISTICALLY (tests: span count law, uniform starts, equal masked count per row,
// SNR law of the mixture) while every deterministic rule of the reference is kept exactly:
//   count_b = max(min_masks, floor(mask_prob * sz_b / L + u_b));   starts = count_b DISTINCT uniform draws from [0, sz_b - L);
//   masked frames = union of the spans; every row is trimmed to the batch minimum by dropping uniformly chosen masked frames.
#include <algorithm>

#include "../../include/unispeech_b200.h"
#include "common.h"
#include "dropout.cuh"
#include "ptx.cuh"

namespace b200 {

namespace {

constexpr int kMaskThreads = 256;
constexpr int kMaxFrames = 4096;   // frames per utterance the sampler supports (T <= 4096: 82 s of audio)

__device__ __forceinline__ float u01(uint32_t k0, uint32_t k1, uint32_t ctr) {
  return (static_cast<float>(drop_bits(k0, k1, ctr) >> 8) + 0.5f) * (1.0f / 16777216.0f);   // 24-bit uniform in (0, 1)
}

// One block per utterance.  keys: a random 32-bit key per candidate start; the `count` candidates with the smallest keys are a
// uniform sample without replacement.  Selecting them = ranking every key against the others (n <= 4096: O(n^2 / threads)).
__global__ void __launch_bounds__(kMaskThreads) span_mask_kernel(const int* __restrict__ valid_len, int T, float mask_prob,
                                                                 int L, int min_masks, uint32_t k0, uint32_t k1,
                                                                 uint8_t* __restrict__ mask, int* __restrict__ counts) {
  pdl_grid_sync();
  __shared__ uint32_t keys[kMaxFrames];
  __shared__ uint32_t bitmap[kMaxFrames / 32];
  __shared__ int n_masked;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int sz = valid_len ? min(valid_len[b], T) : T;
  const uint32_t row = static_cast<uint32_t>(b) * 3u * static_cast<uint32_t>(kMaxFrames);  // disjoint counter ranges per row
  // (without a padding mask the reference draws ONE span count for the whole batch, WavLM.py:74-82; with one, a count per row)
  int count = static_cast<int>(mask_prob * static_cast<float>(sz) / static_cast<float>(L) + u01(k0, k1, valid_len ? row : 0u));
  count = max(min_masks, count);
  int span = L;
  if (sz - span <= count) span = max(1, sz - count - 1);        // WavLM.py:117-118: `min_len = sz - num_mask - 1`
  const int n_cand = max(0, sz - span);
  count = min(count, n_cand);
  for (int i = tid; i < kMaxFrames / 32; i += kMaskThreads) bitmap[i] = 0u;
  if (tid == 0) n_masked = 0;
  for (int i = tid; i < n_cand; i += kMaskThreads) keys[i] = drop_bits(k0, k1, row + 1u + static_cast<uint32_t>(i));
  __syncthreads();
  for (int i = tid; i < n_cand; i += kMaskThreads) {
    const uint32_t ki = keys[i];
    int rank = 0;
    for (int j = 0; j < n_cand; ++j) {
      const uint32_t kj = keys[j];
      rank += (kj < ki || (kj == ki && j < i)) ? 1 : 0;
    }
    if (rank < count) {
      for (int o = 0; o < span && i + o < sz; ++o) atomicOr(&bitmap[(i + o) >> 5], 1u << ((i + o) & 31));
    }
  }
  __syncthreads();
  int local = 0;
  for (int i = tid; i < kMaxFrames / 32; i += kMaskThreads) local += __popc(bitmap[i]);
  local = static_cast<int>(warp_sum(static_cast<float>(local)) + 0.5f);
  if ((tid & 31) == 0 && local) atomicAdd(&n_masked, local);
  __syncthreads();
  if (tid == 0) counts[b] = n_masked;
  for (int t = tid; t < T; t += kMaskThreads) mask[static_cast<long long>(b) * T + t] = (bitmap[t >> 5] >> (t & 31)) & 1u;
}

// Every row keeps exactly min_b counts[b] masked frames: the (counts[b] - keep) masked frames with the smallest random keys are
// cleared (np.random.choice(mask_idc, min_len, replace=False), WavLM.py:147-150).
__global__ void __launch_bounds__(kMaskThreads) span_mask_trim_kernel(int B, int T, uint32_t k0, uint32_t k1,
                                                                      uint8_t* __restrict__ mask, const int* __restrict__ counts) {
  pdl_grid_sync();
  __shared__ uint32_t keys[kMaxFrames];
  __shared__ int pos[kMaxFrames];
  __shared__ int n_s;
  const int b = blockIdx.x, tid = threadIdx.x;
  int keep = counts[0];
  for (int i = 1; i < B; ++i) keep = min(keep, counts[i]);
  const int drop = counts[b] - keep;
  if (drop <= 0) return;
  if (tid == 0) n_s = 0;
  __syncthreads();
  const uint32_t row = static_cast<uint32_t>(b) * 3u * static_cast<uint32_t>(kMaxFrames) + 2u * static_cast<uint32_t>(kMaxFrames);
  for (int t = tid; t < T; t += kMaskThreads) {
    if (mask[static_cast<long long>(b) * T + t]) {
      const int k = atomicAdd(&n_s, 1);
      pos[k] = t;
      keys[k] = drop_bits(k0, k1, row + static_cast<uint32_t>(t));   // the key belongs to the frame, not to its slot
    }
  }
  __syncthreads();
  const int n = n_s;
  for (int i = tid; i < n; i += kMaskThreads) {
    const uint32_t ki = keys[i];
    const int pi = pos[i];
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const uint32_t kj = keys[j];
      rank += (kj < ki || (kj == ki && pos[j] < pi)) ? 1 : 0;
    }
    if (rank < drop) mask[static_cast<long long>(b) * T + pi] = 0;
  }
}

// ---------------------------------------------------------------------------------------------------- utterance mixing
// power[b] += sum x[b, :]^2   (fp32 waveform rows)
__global__ void __launch_bounds__(256) row_power_kernel(const float* __restrict__ x, long long x_bs, int L, double* __restrict__ power) {
  pdl_grid_sync();
  const int b = blockIdx.y;
  const float* row = x + b * x_bs;
  float acc = 0.f;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < L; i += gridDim.x * 256) acc = fmaf(row[i], row[i], acc);
  double d = static_cast<double>(warp_sum(acc));
  __shared__ double red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = d;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(power + b, t);
  }
}

// plan[b] = {src utterance c (or -1: not mixed), chunk length, source start, destination start, snr_db as float bits}; the
// mixture reads the ORIGINAL batch `src` and writes `dst` (= a copy of it):
//   dst[b, s_start + t] += src[c, c_start + t] * sqrt(P_b / (P_c 10^(snr/10)))      (utterance_mixing_dataset.py:415-432)
struct MixPlan { int c, len, c_start, s_start; float snr_db; };
__global__ void __launch_bounds__(256) mix_apply_kernel(const float* __restrict__ src, long long bs, int L,
                                                        const MixPlan* __restrict__ plan, const double* __restrict__ power,
                                                        float* __restrict__ dst) {
  pdl_grid_sync();
  const int b = blockIdx.y;
  const MixPlan pl = plan[b];
  if (pl.c < 0 || pl.len <= 0) return;
  const double pc = power[pl.c] / L, pb = power[b] / L;
  if (pc == 0.0) return;                                   // `if mix_pow == 0: scale = 0`
  const float scale = static_cast<float>(sqrt(pb / (pc * pow(10.0, static_cast<double>(pl.snr_db) / 10.0))));
  const float* s = src + pl.c * bs + pl.c_start;
  float* d = dst + b * bs + pl.s_start;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < pl.len; i += gridDim.x * 256) d[i] = fmaf(s[i], scale, d[i]);
}

// per-utterance normalisation of the data path (F.layer_norm(wav, wav.shape), utterance_mixing_dataset.py:571-573) for rows
// flagged in `which` (NULL = all): two passes, statistics in fp64 accumulators [B, 2] zeroed by the caller
__global__ void __launch_bounds__(256) row_stats_kernel(const float* __restrict__ x, long long bs, int L, const int* __restrict__ valid,
                                                        double* __restrict__ stats) {
  pdl_grid_sync();
  const int b = blockIdx.y;
  const int n = valid ? min(valid[b], L) : L;
  const float* row = x + b * bs;
  float s1 = 0.f, s2 = 0.f;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    const float v = row[i];
    s1 += v;
    s2 = fmaf(v, v, s2);
  }
  double d1 = static_cast<double>(warp_sum(s1)), d2 = static_cast<double>(warp_sum(s2));
  __shared__ double red[16];
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = d1; red[8 + (threadIdx.x >> 5)] = d2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t1 = 0.0, t2 = 0.0;
    for (int w = 0; w < 8; ++w) { t1 += red[w]; t2 += red[8 + w]; }
    atomicAdd(stats + 2 * b, t1);
    atomicAdd(stats + 2 * b + 1, t2);
  }
}
__global__ void __launch_bounds__(256) row_normalize_kernel(float* __restrict__ x, long long bs, int L, const int* __restrict__ valid,
                                                            const double* __restrict__ stats, const MixPlan* __restrict__ plan) {
  pdl_grid_sync();
  const int b = blockIdx.y;
  if (plan != nullptr && plan[b].c < 0) return;            // the reference re-normalises only the utterances it mixed (:433-435)
  const int n = valid ? min(valid[b], L) : L;
  if (n <= 0) return;
  const double mean = stats[2 * b] / n;
  const double var = fmax(stats[2 * b + 1] / n - mean * mean, 0.0);
  const float m = static_cast<float>(mean), r = static_cast<float>(1.0 / sqrt(var + 1e-5));
  float* row = x + b * bs;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) row[i] = (row[i] - m) * r;
}

}  // namespace

}  // namespace b200

using namespace b200;

extern "C" {

int b200s_span_mask(const int* valid_len, int B, int T, float mask_prob, int mask_length, int min_masks, uint32_t key0,
                    uint32_t key1, uint8_t* mask, int* counts, b200s_stream stream) {
  B200_CHECK_ARG(mask && counts, "span_mask: null pointer");
  B200_CHECK_ARG(B > 0 && T > 0 && T <= kMaxFrames, "span_mask: T=%d out of range (1..%d)", T, kMaxFrames);
  B200_CHECK_ARG(mask_length >= 1 && mask_prob >= 0.f, "span_mask: bad mask_length / mask_prob");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  B200_CHECK_CUDA(launch_pdl(span_mask_kernel, dim3(B), dim3(kMaskThreads), 0, st, valid_len, T, mask_prob, mask_length, min_masks,
                             key0, key1, mask, counts));
  B200_CHECK_LAUNCH();
  B200_CHECK_CUDA(launch_pdl(span_mask_trim_kernel, dim3(B), dim3(kMaskThreads), 0, st, B, T, key0, key1, mask,
                             static_cast<const int*>(counts)));
  B200_CHECK_LAUNCH();
  return 0;
}

int b200s_row_power(const float* x, long long x_bs, int B, int L, double* power, b200s_stream stream) {
  B200_CHECK_ARG(x && power && B > 0 && L > 0, "row_power: bad arguments");
  dim3 grid(std::min(ceil_div(L, 256 * 8), 64), B);
  B200_CHECK_CUDA(launch_pdl(row_power_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), x, x_bs, L, power));
  B200_CHECK_LAUNCH();
  return 0;
}

int b200s_mix_apply(const float* src, long long bs, int B, int L, const void* plan, const double* power, float* dst,
                    b200s_stream stream) {
  static_assert(sizeof(MixPlan) == 20, "plan record layout is part of the ABI");
  B200_CHECK_ARG(src && plan && power && dst && B > 0 && L > 0, "mix_apply: bad arguments");
  dim3 grid(std::min(ceil_div(L, 256 * 8), 64), B);
  B200_CHECK_CUDA(launch_pdl(mix_apply_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), src, bs, L,
                             static_cast<const MixPlan*>(plan), power, dst));
  B200_CHECK_LAUNCH();
  return 0;
}

int b200s_row_normalize(float* x, long long bs, int B, int L, const int* valid_len, double* stats, const void* plan,
                        b200s_stream stream) {
  B200_CHECK_ARG(x && stats && B > 0 && L > 0, "row_normalize: bad arguments");
  dim3 grid(std::min(ceil_div(L, 256 * 8), 64), B);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  B200_CHECK_CUDA(launch_pdl(row_stats_kernel, dim3(grid), dim3(256), 0, st, static_cast<const float*>(x), bs, L, valid_len, stats));
  B200_CHECK_LAUNCH();
  B200_CHECK_CUDA(launch_pdl(row_normalize_kernel, dim3(grid), dim3(256), 0, st, x, bs, L, valid_len,
                             static_cast<const double*>(stats), static_cast<const MixPlan*>(plan)));
  B200_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
