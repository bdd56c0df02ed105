// Counter-based dropout masks shared by the row kernel (hidden / activation / input dropout) and the attention kernels
// (dropout on the softmax probabilities).  Replaces the Philox streams behind nn.Dropout / F.dropout in
// WavLM/WavLM.py:350,584,659-661,702-738 and the dropout_p of F.multi_head_attention_forward (WavLM/modules.py:551):
// the decision for one element is a pure function of (key0, key1, row, column), so the backward pass regenerates (or, for the
// attention probabilities, re-reads) exactly the mask of the forward pass and nothing but a 64-bit key is saved.
//
//   bits(ctr)          = fmix32((ctr ^ k1) * 0x9E3779B1 + k0)            one 32-bit word per PAIR of adjacent columns
//   keep(element 2c+e) = ((bits(c) >> 16e) & 0xffff) >= thr16,          thr16 = round(p * 65536)
//   y                  = keep ? x / (1 - p) : 0                          (same scaling as F.dropout)
//
// fmix32 is the MurmurHash3 finaliser (a bijection with full avalanche); p is represented to 2^-16.
// tests/test_dropout_gpu.py holds a numpy restatement of these formulas and checks the kernels bit for bit against it.
#pragma once
#include <stdint.h>

namespace b200 {

__host__ __device__ __forceinline__ uint32_t fmix32(uint32_t h) {
  h ^= h >> 16;
  h *= 0x85EBCA6Bu;
  h ^= h >> 13;
  h *= 0xC2B2AE35u;
  h ^= h >> 16;
  return h;
}

__host__ __device__ __forceinline__ uint32_t drop_bits(uint32_t k0, uint32_t k1, uint32_t ctr) {
  return fmix32((ctr ^ k1) * 0x9E3779B1u + k0);
}

// low / high half of a bits word against the threshold pre-shifted into the high half (thr_hi = thr16 << 16)
__host__ __device__ __forceinline__ bool drop_keep_lo(uint32_t bits, uint32_t thr_hi) { return (bits << 16) >= thr_hi; }
__host__ __device__ __forceinline__ bool drop_keep_hi(uint32_t bits, uint32_t thr_hi) { return bits >= thr_hi; }

// per-row keys of the attention-probability mask: row = (b*H + h)*T + i, counter = key column >> 1
__host__ __device__ __forceinline__ uint32_t drop_row_k0(uint32_t k0, uint32_t row) { return fmix32(k0 + row * 0x9E3779B1u); }
__host__ __device__ __forceinline__ uint32_t drop_row_k1(uint32_t k1, uint32_t row) { return fmix32(k1 ^ (row * 0x85EBCA6Bu)); }

static inline uint32_t drop_threshold16(float p) {
  double t = static_cast<double>(p) * 65536.0 + 0.5;
  if (t < 0.0) t = 0.0;
  if (t > 65535.0) t = 65535.0;
  return static_cast<uint32_t>(t);
}

}  // namespace b200
