// GroupNorm(C, C) path of conv layer 0 ("default" extractor, WavLM-Base; WavLM/WavLM.py:420-426, Fp32GroupNorm
// WavLM/modules.py:45-57) without ever re-scanning the 512-channel activation for statistics.
//
// Because the layer is a Cin = 1 convolution, everything GroupNorm needs is a function of the waveform's k x k
// autocorrelation per utterance:   conv[c,t] = sum_j w[c,j] x[s t + j]
//     sum_t conv[c,t]      = sum_j w[c,j] X1[j]                X1[j]    = sum_t x[s t + j]
//     sum_t conv[c,t]^2    = sum_jj' w[c,j] w[c,j'] A[j,j']    A[j,j']  = sum_t x[s t + j] x[s t + j']
// so the forward statistics cost one tiny pass over the waveform (fp64 accumulation) instead of a full conv pass.
// The backward uses the same trick to be SINGLE pass over the incoming gradient: with dz = da * gelu'(z), dxhat = dz*gamma,
//     dW[c,j] = sum_b rstd_bc * ( P_b[c,j] - m1_bc X1_b[j] - m2_bc Q_b[c,j] )
//     P_b[c,j] = sum_t dxhat[c,t] x[s t + j]   (the only data-dependent reduction, accumulated by the pass kernel)
//     m1 = gamma*dbeta_b/T,  m2 = gamma*dgamma_b/T,   Q_b[c,j] = rstd_bc * ( sum_j' w[c,j'] A_b[j',j] - mean_bc X1_b[j] )
// Pass-kernel mapping: thread <-> 2 channels (coalesced bf16x2 loads of da[t][c]), loop over frames, 2*(k+2) accumulators.
#include <algorithm>

#include "../../include/unispeech_b200.h"
#include "common.h"
#include "ptx.cuh"

namespace b200 {

constexpr int kGnTaps = 10;                 // fast path: kernel size <= 10 (every WavLM / UniSpeech config uses 10)
constexpr int kAc = kGnTaps + kGnTaps * kGnTaps;  // X1[10] + A[10][10] per utterance (fp64), stored with stride 128
constexpr int kAcStride = 128;
constexpr int kFrTile = 64;  // frames staged per shared-memory tile (apply / backward pass kernels)

// ---------------------------------------------------------------------------------------------- waveform autocorrelation
__global__ void __launch_bounds__(256) conv0_autocorr_kernel(const float* __restrict__ wav, long long L, int T, int k, int s,
                                                             double* __restrict__ acorr) {
  pdl_grid_sync();
  const int b = blockIdx.y;
  const float* x = wav + static_cast<long long>(b) * L;
  float x1[kGnTaps], a[kGnTaps * (kGnTaps + 1) / 2];
#pragma unroll
  for (int j = 0; j < kGnTaps; ++j) x1[j] = 0.f;
#pragma unroll
  for (int j = 0; j < kGnTaps * (kGnTaps + 1) / 2; ++j) a[j] = 0.f;
  // each thread accumulates a bounded number of frames in fp32 (then fp64 across threads / blocks)
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x) {
    float w[kGnTaps];
#pragma unroll
    for (int j = 0; j < kGnTaps; ++j) {
      const long long p = static_cast<long long>(t) * s + j;
      w[j] = (j < k && p < L) ? x[p] : 0.f;
    }
    int idx = 0;
#pragma unroll
    for (int j = 0; j < kGnTaps; ++j) {
      x1[j] += w[j];
#pragma unroll
      for (int jj = j; jj < kGnTaps; ++jj) a[idx++] += w[j] * w[jj];
    }
  }
  __shared__ double red[8][kGnTaps + kGnTaps * (kGnTaps + 1) / 2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < kGnTaps; ++j) {
    const float v = warp_sum(x1[j]);
    if (lane == 0) red[warp][j] = v;
  }
#pragma unroll
  for (int j = 0; j < kGnTaps * (kGnTaps + 1) / 2; ++j) {
    const float v = warp_sum(a[j]);
    if (lane == 0) red[warp][kGnTaps + j] = v;
  }
  __syncthreads();
  const int n = kGnTaps + kGnTaps * (kGnTaps + 1) / 2;
  if (threadIdx.x < n) {
    double sum = 0.0;
    for (int w8 = 0; w8 < 8; ++w8) sum += red[w8][threadIdx.x];
    double* dst = acorr + static_cast<long long>(b) * kAcStride;
    if (threadIdx.x < kGnTaps) {
      atomicAdd(dst + threadIdx.x, sum);
    } else {
      // unpack the upper-triangular index into (j, jj) and mirror
      int r = threadIdx.x - kGnTaps, j = 0;
      while (r >= kGnTaps - j) { r -= kGnTaps - j; ++j; }
      const int jj = j + r;
      atomicAdd(dst + kGnTaps + j * kGnTaps + jj, sum);
      if (jj != j) atomicAdd(dst + kGnTaps + jj * kGnTaps + j, sum);
    }
  }
}

// stats[b][c] = {sum_t conv, sum_t conv^2}
__global__ void conv0_gn_stats_finalize_kernel(const float* __restrict__ w, const double* __restrict__ acorr, int B, int C,
                                               int k, double* __restrict__ stats) {
  pdl_grid_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int b = i / C, c = i % C;
  const double* ac = acorr + static_cast<long long>(b) * kAcStride;
  double s1 = 0.0, s2 = 0.0;
  for (int j = 0; j < k; ++j) {
    const double wj = w[c * k + j];
    s1 += wj * ac[j];
    for (int jj = 0; jj < k; ++jj) s2 += wj * static_cast<double>(w[c * k + jj]) * ac[kGnTaps + j * kGnTaps + jj];
  }
  stats[static_cast<long long>(i) * 2] = s1;
  stats[static_cast<long long>(i) * 2 + 1] = s2;
}

// ---------------------------------------------------------------------------------------------- forward apply pass
// out[b,t,c] = gelu(gamma_c * (conv[b,t,c] - mean_bc) * rstd_bc + beta_c), written bf16 channels-last.  Thread <-> 2 adjacent
// channels with the ten taps in registers, GroupNorm folded INTO the taps (w' = gamma*rstd*w, b' = beta - gamma*rstd*mean), so
// an output element costs 10 FMA + GELU; the waveform window of each frame is broadcast from shared memory (three 16-byte
// loads per frame) and every warp store covers one full 128-byte line of a frame row.
template <int C>
__global__ void __launch_bounds__(C / 2) conv0_gn_fwd_apply_kernel(const float* __restrict__ wav, long long L, int T, int k,
                                                                   int s, const float* __restrict__ w,
                                                                   const float* __restrict__ gamma,
                                                                   const float* __restrict__ beta,
                                                                   const double* __restrict__ stats, int t_chunk,
                                                                   __nv_bfloat16* __restrict__ out, long long out_bs) {
  pdl_grid_sync();
  __shared__ __align__(16) float xs[kFrTile][12];
  const int b = blockIdx.y;
  const int t_begin = blockIdx.x * t_chunk;
  const int t_end = min(T, t_begin + t_chunk);
  const int c0 = threadIdx.x * 2;
  const float* x = wav + static_cast<long long>(b) * L;
  float wr[2][kGnTaps], bias[2];
#pragma unroll
  for (int ch = 0; ch < 2; ++ch) {
    const double m = stats[(static_cast<long long>(b) * C + c0 + ch) * 2] / T;
    const double var = stats[(static_cast<long long>(b) * C + c0 + ch) * 2 + 1] / T - m * m;
    const float rstd = static_cast<float>(1.0 / sqrt((var > 0 ? var : 0) + 1e-5));
    const float a = gamma[c0 + ch] * rstd;
#pragma unroll
    for (int j = 0; j < kGnTaps; ++j) wr[ch][j] = (j < k) ? a * w[(c0 + ch) * k + j] : 0.f;
    bias[ch] = beta[c0 + ch] - a * static_cast<float>(m);
  }
  __nv_bfloat16* out_b = out + b * out_bs + c0;
  for (int t0 = t_begin; t0 < t_end; t0 += kFrTile) {
    const int nf = min(kFrTile, t_end - t0);
    __syncthreads();
    for (int i = threadIdx.x; i < kFrTile * 12; i += blockDim.x) {
      const int f = i / 12, j = i % 12;
      const long long p = static_cast<long long>(t0 + f) * s + j;
      xs[f][j] = (f < nf && j < k && p < L) ? x[p] : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int f = 0; f < nf; ++f) {
      const float4 w0 = *reinterpret_cast<const float4*>(&xs[f][0]);
      const float4 w1 = *reinterpret_cast<const float4*>(&xs[f][4]);
      const float4 w2 = *reinterpret_cast<const float4*>(&xs[f][8]);
      const float win[12] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w};
      float z[2];
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        float acc = bias[ch];
#pragma unroll
        for (int j = 0; j < kGnTaps; ++j) acc = fmaf(wr[ch][j], win[j], acc);
        z[ch] = gelu_f(acc);
      }
      *reinterpret_cast<uint32_t*>(out_b + static_cast<long long>(t0 + f) * C) = pack_bf16x2(z[0], z[1]);
    }
  }
}

// ---------------------------------------------------------------------------------------------- backward pass
template <int C>
__global__ void __launch_bounds__(C / 2) conv0_gn_bwd_pass_kernel(const float* __restrict__ wav, long long L, int T, int k,
                                                                  int s, const float* __restrict__ w,
                                                                  const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta,
                                                                  const double* __restrict__ stats,
                                                                  const __nv_bfloat16* __restrict__ da, long long da_bs,
                                                                  int t_chunk, float* __restrict__ bstats) {
  pdl_grid_sync();
  __shared__ __align__(16) float xs[kFrTile][12];
  const int b = blockIdx.y;
  const int t_begin = blockIdx.x * t_chunk;
  const int t_end = min(T, t_begin + t_chunk);
  const int c0 = threadIdx.x * 2;
  const float* x = wav + static_cast<long long>(b) * L;
  float wr[2][kGnTaps], P[2][kGnTaps], g[2], be[2], mean[2], rstd[2], dbe[2], dga[2];
#pragma unroll
  for (int ch = 0; ch < 2; ++ch) {
#pragma unroll
    for (int j = 0; j < kGnTaps; ++j) {
      wr[ch][j] = (j < k) ? w[(c0 + ch) * k + j] : 0.f;
      P[ch][j] = 0.f;
    }
    g[ch] = gamma[c0 + ch];
    be[ch] = beta[c0 + ch];
    const double m = stats[(static_cast<long long>(b) * C + c0 + ch) * 2] / T;
    const double var = stats[(static_cast<long long>(b) * C + c0 + ch) * 2 + 1] / T - m * m;
    mean[ch] = static_cast<float>(m);
    rstd[ch] = static_cast<float>(1.0 / sqrt((var > 0 ? var : 0) + 1e-5));
    dbe[ch] = dga[ch] = 0.f;
  }
  const __nv_bfloat16* da_b = da + b * da_bs + c0;
  for (int t0 = t_begin; t0 < t_end; t0 += kFrTile) {
    const int nf = min(kFrTile, t_end - t0);
    __syncthreads();
    for (int i = threadIdx.x; i < kFrTile * 12; i += blockDim.x) {
      const int f = i / 12, j = i % 12;
      const long long p = static_cast<long long>(t0 + f) * s + j;
      xs[f][j] = (f < nf && j < k && p < L) ? x[p] : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int f = 0; f < nf; ++f) {
      const float2 d = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(da_b + static_cast<long long>(t0 + f) * C));
      const float4 w0 = *reinterpret_cast<const float4*>(&xs[f][0]);
      const float4 w1 = *reinterpret_cast<const float4*>(&xs[f][4]);
      const float4 w2 = *reinterpret_cast<const float4*>(&xs[f][8]);
      const float win[12] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w};
      const float dv[2] = {d.x, d.y};
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        float conv = 0.f;
#pragma unroll
        for (int j = 0; j < kGnTaps; ++j) conv = fmaf(wr[ch][j], win[j], conv);
        const float xh = (conv - mean[ch]) * rstd[ch];
        const float dz = dv[ch] * gelu_grad_f(fmaf(g[ch], xh, be[ch]));
        dbe[ch] += dz;
        dga[ch] = fmaf(dz, xh, dga[ch]);
        const float dxh = dz * g[ch];
#pragma unroll
        for (int j = 0; j < kGnTaps; ++j) P[ch][j] = fmaf(dxh, win[j], P[ch][j]);
      }
    }
  }
#pragma unroll
  for (int ch = 0; ch < 2; ++ch) {
    float* dst = bstats + (static_cast<long long>(b) * C + c0 + ch) * 12;
#pragma unroll
    for (int j = 0; j < kGnTaps; ++j) atomicAdd(dst + j, P[ch][j]);
    atomicAdd(dst + 10, dbe[ch]);
    atomicAdd(dst + 11, dga[ch]);
  }
}

// dW[c,j], dgamma[c], dbeta[c] from the per-utterance partials (fp64 arithmetic, one thread per (c, j))
__global__ void conv0_gn_bwd_finalize_kernel(const float* __restrict__ w, const float* __restrict__ gamma,
                                             const double* __restrict__ stats, const double* __restrict__ acorr,
                                             const float* __restrict__ bstats, int B, int C, int k, int T,
                                             float* __restrict__ dw, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  pdl_grid_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C * k) return;
  const int c = i / k, j = i % k;
  double acc = 0.0, sg = 0.0, sb = 0.0;
  for (int b = 0; b < B; ++b) {
    const double* ac = acorr + static_cast<long long>(b) * kAcStride;
    const float* bs = bstats + (static_cast<long long>(b) * C + c) * 12;
    const double m = stats[(static_cast<long long>(b) * C + c) * 2] / T;
    const double var = stats[(static_cast<long long>(b) * C + c) * 2 + 1] / T - m * m;
    const double rstd = 1.0 / sqrt((var > 0 ? var : 0) + 1e-5);
    const double db = bs[10], dg = bs[11];
    const double m1 = gamma[c] * db / T, m2 = gamma[c] * dg / T;
    double cx = 0.0;
    for (int jj = 0; jj < k; ++jj) cx += static_cast<double>(w[c * k + jj]) * ac[kGnTaps + jj * kGnTaps + j];
    const double q = rstd * (cx - m * ac[j]);
    acc += rstd * (static_cast<double>(bs[j]) - m1 * ac[j] - m2 * q);
    sg += dg;
    sb += db;
  }
  atomicAdd(dw + i, static_cast<float>(acc));
  if (j == 0) {
    atomicAdd(dgamma + c, static_cast<float>(sg));
    atomicAdd(dbeta + c, static_cast<float>(sb));
  }
}

// ---------------------------------------------------------------------------------------------- launch helpers (used by conv0.cu)
int conv0_gn_stats_launch(const float* wav, long long L, int B, int T, int C, int k, int s, const float* w, double* stats,
                          cudaStream_t st) {
  B200_CHECK_ARG(k <= kGnTaps, "conv0 GroupNorm path supports kernel size <= %d (got %d)", kGnTaps, k);
  double* acorr = stats + static_cast<long long>(B) * C * 2;  // caller allocates B*C*2 + B*128 doubles
  B200_CHECK_CUDA(cudaMemsetAsync(acorr, 0, sizeof(double) * B * kAcStride, st));
  dim3 grid(std::max(1, std::min(ceil_div(T, 256 * 8), 64)), B);
  B200_CHECK_CUDA(launch_pdl(conv0_autocorr_kernel, dim3(grid), dim3(256), 0, st, wav, L, T, k, s, acorr));
  B200_CHECK_LAUNCH();
  B200_CHECK_CUDA(launch_pdl(conv0_gn_stats_finalize_kernel, dim3(ceil_div(B * C, 128)), dim3(128), 0, st, w, acorr, B, C, k, stats));
  B200_CHECK_LAUNCH();
  return 0;
}

int conv0_gn_fwd_apply_launch(const float* wav, long long L, int B, int T, int C, int k, int s, const float* w,
                              const float* gamma, const float* beta, const double* stats, void* out, long long out_bs,
                              cudaStream_t st) {
  B200_CHECK_ARG(k <= kGnTaps, "conv0 GroupNorm path supports kernel size <= %d (got %d)", kGnTaps, k);
  int chunks = std::max(1, (8 * sm_count()) / std::max(1, B));
  int t_chunk = ceil_div(ceil_div(T, chunks), kFrTile) * kFrTile;
  chunks = ceil_div(T, t_chunk);
  dim3 grid(chunks, B);
  __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out);
  if (C == 512) {
    B200_CHECK_CUDA(launch_pdl(conv0_gn_fwd_apply_kernel<512>, dim3(grid), dim3(256), 0, st, wav, L, T, k, s, w, gamma, beta, stats, t_chunk, o, out_bs));
  } else if (C == 64) {
    B200_CHECK_CUDA(launch_pdl(conv0_gn_fwd_apply_kernel<64>, dim3(grid), dim3(32), 0, st, wav, L, T, k, s, w, gamma, beta, stats, t_chunk, o, out_bs));
  } else {
    set_last_error("conv0: channel count %d not supported (64 / 512)", C);
    return -1;
  }
  B200_CHECK_LAUNCH();
  return 0;
}

int conv0_gn_bwd_launch(const float* wav, long long L, int B, int T, int C, int k, int s, const float* w, const float* gamma,
                        const float* beta, const double* stats, float* bstats, const void* da, long long da_bs, float* dw,
                        float* dgamma, float* dbeta, cudaStream_t st) {
  B200_CHECK_ARG(k <= kGnTaps, "conv0 GroupNorm path supports kernel size <= %d (got %d)", kGnTaps, k);
  B200_CHECK_CUDA(cudaMemsetAsync(bstats, 0, sizeof(float) * 12 * B * C, st));
  // enough blocks to fill the machine a few times over, chunks a multiple of the frame tile
  int chunks = std::max(1, (8 * sm_count()) / std::max(1, B));
  int t_chunk = ceil_div(ceil_div(T, chunks), kFrTile) * kFrTile;
  chunks = ceil_div(T, t_chunk);
  dim3 grid(chunks, B);
  const __nv_bfloat16* dap = static_cast<const __nv_bfloat16*>(da);
  const double* acorr = stats + static_cast<long long>(B) * C * 2;
  if (C == 512) {
    B200_CHECK_CUDA(launch_pdl(conv0_gn_bwd_pass_kernel<512>, dim3(grid), dim3(256), 0, st, wav, L, T, k, s, w, gamma, beta, stats, dap, da_bs, t_chunk, bstats));
  } else if (C == 64) {
    B200_CHECK_CUDA(launch_pdl(conv0_gn_bwd_pass_kernel<64>, dim3(grid), dim3(32), 0, st, wav, L, T, k, s, w, gamma, beta, stats, dap, da_bs, t_chunk, bstats));
  } else {
    set_last_error("conv0: channel count %d not supported (64 / 512)", C);
    return -1;
  }
  B200_CHECK_LAUNCH();
  B200_CHECK_CUDA(launch_pdl(conv0_gn_bwd_finalize_kernel, dim3(ceil_div(C * k, 128)), dim3(128), 0, st, w, gamma, stats, acorr, bstats, B, C, k, T, dw, dgamma,
                                                                    dbeta));
  B200_CHECK_LAUNCH();
  return 0;
}

}  // namespace b200
