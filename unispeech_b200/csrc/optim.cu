// Optimizer step on the flat gradient buffer: global gradient norm, clipping, gradient scaling and the fairseq Adam update
// in two launches (sum of squares; update), with no host synchronisation in between.
//
// Replaces, for the fp32 masters of this model:
//   * utils.clip_grad_norm_            src/fairseq/utils.py:338-381        (per-tensor norms -> stack -> norm -> clamp -> mul_)
//   * FP16Optimizer._unscale_grads / multiply_grads / clip_grad_norm      src/fairseq/optim/fp16_optimizer.py:176-214
//   * Adam.step                        src/fairseq/optim/adam.py:150-228   (5 elementwise passes per tensor, ~500 tensors)
// Both kernels are HBM-bound: the update reads g, m, v, p and writes m, v, p [and zeroes g] = 28-32 bytes per parameter.
#include <algorithm>
#include <cmath>

#include "../../include/unispeech_b200.h"
#include "common.h"
#include "ptx.cuh"

namespace b200 {

namespace {

constexpr int kAdamThreads = 256;
constexpr int kAdamChunk = 2048;  // elements per block: two float4 per thread

struct AdamTensor {   // 32 bytes, part of the ABI (b200s_adam_step)
  float* param;       // fp32 master
  long long goff;     // element offset of this tensor's gradient (and Adam moments) in the flat buffers, multiple of 4
  long long numel;
  long long chunk0;   // first global chunk index of this tensor (prefix sum of ceil(numel / 2048))
};

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long long n, double* __restrict__ out) {
  pdl_grid_sync();
  float acc = 0.f;
  const long long nv = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nv;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 v = g4[i];
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const float v = g[(nv << 2) + threadIdx.x];
    acc += v * v;
  }
  double d = static_cast<double>(warp_sum(acc));
  __shared__ double red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = d;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += red[w];
    atomicAdd(out, s);
  }
}

struct AdamHyper {
  float grad_scale, max_norm, lr, beta1, beta2, eps, weight_decay, step_size;
  int zero_grad;
};

__device__ __forceinline__ void adam_one(float& p, float& g, float& m, float& v, float gs, const AdamHyper& h) {
  const float gg = g * gs;
  m = m * h.beta1 + (1.f - h.beta1) * gg;                 // exp_avg.mul_(beta1).add_(grad, alpha=1-beta1)
  v = v * h.beta2 + (1.f - h.beta2) * gg * gg;            // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1-beta2)
  const float denom = sqrtf(v) + h.eps;                   // exp_avg_sq.sqrt().add_(eps)
  float pp = p;
  if (h.weight_decay != 0.f) pp += pp * (-h.weight_decay * h.lr);   // p.add_(p, alpha=-weight_decay*lr)
  p = pp - h.step_size * (m / denom);                     // p.addcdiv_(exp_avg, denom, value=-step_size)
  if (h.zero_grad) g = 0.f;
}

// Squared norm of the gradients of the tensors in the descriptor table ONLY (fairseq's clip_grad_norm_ sees the parameters the
// optimizer owns whose .grad exists, src/fairseq/utils.py:338-345): regions of the flat buffer that belong to excluded / frozen
// parameters are not counted.  Same block -> tensor mapping as the update kernel.
__global__ void __launch_bounds__(kAdamThreads) sumsq_table_kernel(const AdamTensor* __restrict__ table, int n_tensors,
                                                                   const float* __restrict__ g, double* __restrict__ out) {
  pdl_grid_sync();
  int lo = 0, hi = n_tensors - 1;
  const long long c = blockIdx.x;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (table[mid].chunk0 <= c) lo = mid; else hi = mid - 1;
  }
  const AdamTensor t = table[lo];
  const long long base = (c - t.chunk0) * kAdamChunk;
  float acc = 0.f;
#pragma unroll
  for (int part = 0; part < kAdamChunk / (kAdamThreads * 4); ++part) {
    const long long i = base + (part * kAdamThreads + threadIdx.x) * 4;
    if (i >= t.numel) break;
    const float* gp = g + t.goff + i;
    if (i + 4 <= t.numel) {
      const float4 G = *reinterpret_cast<const float4*>(gp);   // every view of the flat buffer starts 16-byte aligned
      acc += G.x * G.x + G.y * G.y + G.z * G.z + G.w * G.w;
    } else {
      for (int e = 0; e < 4 && i + e < t.numel; ++e) acc = fmaf(gp[e], gp[e], acc);
    }
  }
  double d = static_cast<double>(warp_sum(acc));
  __shared__ double red[kAdamThreads / 32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = d;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < kAdamThreads / 32; ++w) s += red[w];
    atomicAdd(out, s);
  }
}

__global__ void __launch_bounds__(kAdamThreads) adam_step_kernel(const AdamTensor* __restrict__ table, int n_tensors,
                                                                 float* __restrict__ g, float* __restrict__ m,
                                                                 float* __restrict__ v, const double* __restrict__ sumsq,
                                                                 const AdamHyper h) {
  pdl_grid_sync();
  int lo = 0, hi = n_tensors - 1;
  const long long c = blockIdx.x;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (table[mid].chunk0 <= c) lo = mid; else hi = mid - 1;
  }
  const AdamTensor t = table[lo];
  // gradient scale: the caller's multiply factor times the clip coefficient of the global norm
  //   clip_coef = (max_norm / (total_norm + 1e-6)).clamp_(max=1)        src/fairseq/utils.py:378-381, fp16_optimizer.py:207-209
  float gs = h.grad_scale;
  if (h.max_norm > 0.f && sumsq != nullptr) {
    const float total = fabsf(h.grad_scale) * static_cast<float>(sqrt(*sumsq));
    gs *= fminf(1.f, h.max_norm / (total + 1e-6f));
  }
  const long long base = (c - t.chunk0) * kAdamChunk;
#pragma unroll
  for (int part = 0; part < kAdamChunk / (kAdamThreads * 4); ++part) {
    const long long i = base + (part * kAdamThreads + threadIdx.x) * 4;
    if (i >= t.numel) break;
    float* pp = t.param + i;
    float* gp = g + t.goff + i;
    float* mp = m + t.goff + i;
    float* vp = v + t.goff + i;
    if (i + 4 <= t.numel && (reinterpret_cast<uintptr_t>(pp) & 15) == 0) {
      float4 P = *reinterpret_cast<float4*>(pp), G = *reinterpret_cast<float4*>(gp);
      float4 M = *reinterpret_cast<float4*>(mp), V = *reinterpret_cast<float4*>(vp);
      adam_one(P.x, G.x, M.x, V.x, gs, h);
      adam_one(P.y, G.y, M.y, V.y, gs, h);
      adam_one(P.z, G.z, M.z, V.z, gs, h);
      adam_one(P.w, G.w, M.w, V.w, gs, h);
      *reinterpret_cast<float4*>(pp) = P;
      *reinterpret_cast<float4*>(mp) = M;
      *reinterpret_cast<float4*>(vp) = V;
      if (h.zero_grad) *reinterpret_cast<float4*>(gp) = G;
    } else {
      for (int e = 0; e < 4 && i + e < t.numel; ++e) {
        float P = pp[e], G = gp[e], M = mp[e], V = vp[e];
        adam_one(P, G, M, V, gs, h);
        pp[e] = P;
        mp[e] = M;
        vp[e] = V;
        if (h.zero_grad) gp[e] = G;
      }
    }
  }
}

}  // namespace

}  // namespace b200

using namespace b200;

extern "C" {

int b200s_sumsq_f32(const float* g, long long n, double* out, b200s_stream stream) {
  B200_CHECK_ARG(g && out && n >= 0, "sumsq_f32: bad arguments");
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(g) & 15) == 0, "sumsq_f32: buffer must be 16-byte aligned");
  if (n == 0) return 0;
  const int grid = static_cast<int>(std::min<long long>(ceil_div_ll(n / 4 + 1, 256), 8LL * sm_count()));
  B200_CHECK_CUDA(launch_pdl(sumsq_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), g, n, out));
  B200_CHECK_LAUNCH();
  return 0;
}

int b200s_sumsq_table(const void* table, int n_tensors, long long total_chunks, const float* g, double* out,
                      b200s_stream stream) {
  B200_CHECK_ARG(table && g && out && n_tensors > 0 && total_chunks > 0, "sumsq_table: bad arguments");
  B200_CHECK_ARG(total_chunks < (1LL << 31), "sumsq_table: too many chunks");
  B200_CHECK_CUDA(launch_pdl(sumsq_table_kernel, dim3(static_cast<unsigned>(total_chunks)), dim3(kAdamThreads), 0,
                             static_cast<cudaStream_t>(stream), static_cast<const AdamTensor*>(table), n_tensors, g, out));
  B200_CHECK_LAUNCH();
  return 0;
}

int b200s_adam_step(const void* table, int n_tensors, long long total_chunks, float* g, float* m, float* v,
                    const double* sumsq, float grad_scale, float max_norm, float lr, float beta1, float beta2, float eps,
                    float weight_decay, int step, int zero_grad, b200s_stream stream) {
  static_assert(sizeof(AdamTensor) == 32, "descriptor layout is part of the ABI");
  B200_CHECK_ARG(table && g && m && v && n_tensors > 0 && total_chunks > 0, "adam_step: bad arguments");
  B200_CHECK_ARG(step >= 1, "adam_step: step=%d must be >= 1 (state[\"step\"] after the increment)", step);
  B200_CHECK_ARG(max_norm <= 0.f || sumsq != nullptr, "adam_step: clipping needs the squared gradient norm (b200s_sumsq_f32)");
  B200_CHECK_ARG(total_chunks < (1LL << 31), "adam_step: too many chunks");
  AdamHyper h;
  h.grad_scale = grad_scale; h.max_norm = max_norm; h.lr = lr; h.beta1 = beta1; h.beta2 = beta2; h.eps = eps;
  h.weight_decay = weight_decay; h.zero_grad = zero_grad;
  // bias_correction{1,2} = 1 - beta^step;  step_size = lr * sqrt(bias_correction2) / bias_correction1     adam.py:213-215
  const double bc1 = 1.0 - pow(static_cast<double>(beta1), step), bc2 = 1.0 - pow(static_cast<double>(beta2), step);
  h.step_size = static_cast<float>(static_cast<double>(lr) * sqrt(bc2) / bc1);
  B200_CHECK_CUDA(launch_pdl(adam_step_kernel, dim3(static_cast<unsigned>(total_chunks)), dim3(kAdamThreads), 0,
                             static_cast<cudaStream_t>(stream), static_cast<const AdamTensor*>(table), n_tensors, g, m, v, sumsq,
                             h));
  B200_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
