// Flash-style attention forward with the WavLM gated relative-position bias, tcgen05 + TMEM + TMA (sm_100a).
//
// One CTA = 256 query rows of one (batch, head): two warpgroups (WG) of 128 threads, each owning one 128-row query tile,
// plus one TMA producer warp.  Thread r of a WG owns query row r (TMEM lane r): the row maximum / sum of the online softmax
// are thread-local, no shuffles.  The two WGs share every K/V tile (one TMA load feeds both) and ping-pong on the tensor
// core: while one WG runs its softmax on the CUDA cores, the other WG's S = Q K^T and O = P V MMAs run.
// Per WG and key tile n (128 keys):
//   S_n  = Q K_n^T           tcgen05.mma 128x128x64  -> TMEM
//   p    = exp2(S*scale*log2e + gate_i*log2e*tab[j-i] + keymask - m)   (two passes over TMEM: max, then exp)
//   P_n -> shared memory in the K-major SWIZZLE_128B operand layout (bf16)
//   O_n  = P_n V_n           tcgen05.mma 128x64x128, V_n read as an MN-major operand straight from the TMA tile
//   O_reg = O_reg*alpha + O_{n-1}   (registers; the rescale never touches TMEM and is deferred by one tile, so the PV MMA
//                                    of tile n overlaps the softmax of tile n+1)
// K/V stages are released by tcgen05.commit arrivals of BOTH WGs' issuing threads (mbarrier count 2).
#include "../../include/unispeech_b200.h"
#include "attn_common.cuh"
#include "common.h"

namespace b200 {

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 32 x 32 bit-matrix transpose across a warp: on entry bit j of lane l's word is element (l, j); on return bit l of lane j's
// word is that element.  Five butterfly stages (one shuffle + three logic ops each) replace 32 ballots.
__device__ __forceinline__ uint32_t warp_bit_transpose(uint32_t x, int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const uint32_t m = (s == 16) ? 0x0000FFFFu : (s == 8) ? 0x00FF00FFu : (s == 4) ? 0x0F0F0F0Fu : (s == 2) ? 0x33333333u : 0x55555555u;
    const uint32_t y = __shfl_xor_sync(0xffffffffu, x, s);
    x = (lane & s) ? ((x & ~m) | ((y >> s) & m)) : ((x & m) | ((y << s) & ~m));
  }
  return x;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

constexpr int kFwdQ = 0;                   // 2 x 16 KB (one Q tile per warpgroup)
constexpr int kFwdK = 32768;               // 2 stages x 16 KB
constexpr int kFwdV = 65536;               // 2 stages x 16 KB
constexpr int kFwdP = 98304;               // 2 x 32 KB (one P tile per warpgroup)
constexpr int kFwdTab = 163840;            // fp32 bias-table slice, key mask, tile flags
constexpr int kFwdThreads = 288;           // 2 warpgroups + 1 producer warp

template <bool HAS_BIAS, bool DROP>
__global__ void __launch_bounds__(kFwdThreads, 1) attn_fwd_kernel(const __grid_constant__ CUtensorMap tm,
                                                                 const __grid_constant__ AttnParams p) {
  pdl_grid_sync();
  const int tid = threadIdx.x, warp = tid >> 5;
  const int wg = warp >> 2;  // 0, 1 = softmax warpgroups; 2 = producer warp
  const int q0 = blockIdx.x * 2 * kAttnTile, h = blockIdx.y, b = blockIdx.z;
  const int T = p.T, D = p.D, N = p.n_tiles;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // 1024-aligned, still a __shared__ pointer (LDS/STS, not generic)
  uint8_t* sQ = smem + kFwdQ;
  uint8_t* sK = smem + kFwdK;
  uint8_t* sV = smem + kFwdV;
  uint8_t* sP = smem + kFwdP;
  float* tab_s = reinterpret_cast<float*>(smem + kFwdTab);  // [(N+2)*128]: index j - r + 255, r = row inside the 256-row block
  float* kbias = tab_s + (N + 2) * kAttnTile;                // [N*128]
  int* tile_flags = reinterpret_cast<int*>(kbias + N * kAttnTile);

  __shared__ uint64_t q_full, k_full[2], k_empty[2], v_full[2], v_empty[2], s_full[2], o_full[2];
  __shared__ uint32_t tmem_base_s;

  if (tid == 0) {
    tma_prefetch_desc(&tm);
    mbar_init(&q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&k_empty[i], 2);  // one tcgen05.commit arrival per warpgroup
      mbar_init(&v_empty[i], 2);
      mbar_init(&s_full[i], 1);   // index = warpgroup
      mbar_init(&o_full[i], 1);
    }
    fence_mbar_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (HAS_BIAS) {
    const int len = (N + 2) * kAttnTile;
    const int base = (T - 1) - (q0 + 2 * kAttnTile - 1);
    for (int i = tid; i < len; i += blockDim.x) {
      const int gi = i + base;
      tab_s[i] = (gi >= 0 && gi < 2 * T - 1) ? p.tab[static_cast<long long>(h) * (2 * T - 1) + gi] : 0.f;
    }
  }
  load_key_mask(kbias, tile_flags, p.key_pad, b, T, N);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (wg == 2) {
    // ------------------------------------------------------------------ TMA producer warp
    if ((tid & 31) == 0) {
      mbar_expect_tx(&q_full, 32768);
      tma_load_4d(sQ, &tm, &q_full, h * kHeadDim, q0, b, 0);
      tma_load_4d(sQ + 16384, &tm, &q_full, h * kHeadDim, q0 + kAttnTile, b, 0);
      for (int n = 0; n < N; ++n) {
        const int s = n & 1;
        const uint32_t ph = (n >> 1) & 1;
        mbar_wait(&k_empty[s], ph ^ 1);
        mbar_expect_tx(&k_full[s], 16384);
        tma_load_4d(sK + s * 16384, &tm, &k_full[s], D + h * kHeadDim, n * kAttnTile, b, 0);
        mbar_wait(&v_empty[s], ph ^ 1);
        mbar_expect_tx(&v_full[s], 16384);
        tma_load_4d(sV + s * 16384, &tm, &v_full[s], 2 * D + h * kHeadDim, n * kAttnTile, b, 0);
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax / MMA-issuing warpgroups
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t idesc_pv = make_idesc_bf16(128, 64, 0, 1);
    const int r = tid & 127;                 // row inside this warpgroup's tile == TMEM lane
    const int r256 = wg * kAttnTile + r;     // row inside the CTA's 256-row block
    const bool issuer = (r == 0);
    const bool row_valid = (q0 + r256) < T;
    const uint32_t tmem_s = tmem + wg * 256;        // S: 128 columns
    const uint32_t tmem_o = tmem + wg * 256 + 128;  // O tile: 64 columns
    uint8_t* sQw = sQ + wg * 16384;
    uint8_t* sPw = sP + wg * 32768;
    const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;

    auto issue_s = [&](int n) {  // S_n = Q K_n^T, then release the K stage once the MMAs retire
      const uint32_t a = smem_u32(sQw), bb = smem_u32(sK + (n & 1) * 16384);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16(tmem_s, make_smem_desc_sw128(a + k * 32, 16, 1024), make_smem_desc_sw128(bb + k * 32, 16, 1024), idesc_s,
                  k > 0 ? 1u : 0u);
      umma_commit(&s_full[wg]);
      umma_commit(&k_empty[n & 1]);
    };

    if (issuer) {
      mbar_wait(&q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_s(0);
    }
    __syncwarp();

    float gl = 0.f;
    if (HAS_BIAS) {
      const float g = (p.gate != nullptr && row_valid) ? p.gate[(static_cast<long long>(b) * p.H + h) * T + q0 + r256] : 1.0f;
      gl = g * kLog2e;
    }
    const float sc = p.scale * kLog2e;
    const float* tabrow = tab_s + (2 * kAttnTile - 1 - r256);
    // dropout on the probabilities: per-row hash keys, and where this warp's 32 rows keep their bits (one word per key column)
    uint32_t rk0 = 0, rk1 = 0;
    uint32_t* mask_row = nullptr;
    if (DROP) {
      const uint32_t rowid = static_cast<uint32_t>(b * p.H + h) * static_cast<uint32_t>(T) + static_cast<uint32_t>(q0 + r256);
      rk0 = drop_row_k0(p.drop_k0, rowid);
      rk1 = drop_row_k1(p.drop_k1, rowid);
      if (q0 + wg * kAttnTile < N * kAttnTile)  // (a 256-row CTA may reach past the last 128-row tile: nothing to record there)
        mask_row = p.drop_mask + (static_cast<long long>(b * p.H + h) * (4 * N) + ((q0 + r256) >> 5)) * (N * kAttnTile);
    }

    float m_run = -INFINITY, l_run = 0.f, alpha_prev = 0.f;
    float o[kHeadDim];
#pragma unroll
    for (int i = 0; i < kHeadDim; ++i) o[i] = 0.f;

    auto accumulate_o = [&]() {
      uint32_t t0[32], t1[32];
      tmem_ld_32x32b_x32(tmem_o + lane_addr, t0);
      tmem_ld_32x32b_x32(tmem_o + lane_addr + 32, t1);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        o[i] = o[i] * alpha_prev + __uint_as_float(t0[i]);
        o[32 + i] = o[32 + i] * alpha_prev + __uint_as_float(t1[i]);
      }
    };

    for (int n = 0; n < N; ++n) {
      const int k0 = n * kAttnTile;
      mbar_wait(&s_full[wg], n & 1);
      tc_fence_after();
      const bool msk = tile_flags[n] != 0;
      const uint32_t s_addr = tmem_s + lane_addr;
      // The softmax is invariant to the reference value subtracted in the exponent, so after the first tile the running
      // maximum does NOT have to be refreshed: probabilities are taken relative to the reference found so far (ONE pass over the
      // scores instead of max-then-exp), fp32 row sums / accumulators absorb factors up to 2^64.  A warp falls back to the
      // two-pass form while some row has no finite reference yet (first tile, or only masked keys so far) or if a score
      // exceeds the reference by more than 64 (log2 units), which re-bases that tile.
      bool fast = !__any_sync(0xffffffffu, m_run == -INFINITY);
      bool folded = false;
      float m_new, m_use, alpha_cur, lsum;
      while (true) {
        if (!fast) {
          // ---- pass 1: row maximum of this tile
          float mx = -INFINITY;
#pragma unroll 1
          for (int c0 = 0; c0 < kAttnTile; c0 += 32) {
            uint32_t su[32];
            tmem_ld_32x32b_x32(s_addr + c0, su);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float x = __uint_as_float(su[j]) * sc;
              if (HAS_BIAS) x = fmaf(gl, tabrow[k0 + c0 + j], x);
              if (msk) x += kbias[k0 + c0 + j];
              mx = fmaxf(mx, x);
            }
          }
          m_new = fmaxf(m_run, mx);
          m_use = (m_new == -INFINITY) ? 0.f : m_new;
          alpha_cur = fast_exp2(m_run - m_use);
        } else {
          m_new = m_run;
          m_use = m_run;
          alpha_cur = 1.0f;
        }
        // ---- the PV MMA of the previous tile has long finished: fold its result in (this also frees the P buffer and O tile)
        if (n >= 1 && !folded) {
          mbar_wait(&o_full[wg], (n - 1) & 1);
          tc_fence_after();
          accumulate_o();
          folded = true;
        }
        // ---- probabilities, row sum, bf16 P tile into shared memory (operand layout)
        lsum = 0.f;
        float over = -INFINITY;  // largest exponent argument seen (fast path guard)
        const float neg_ref = -m_use;
#pragma unroll 1
        for (int c0 = 0; c0 < kAttnTile; c0 += 32) {
          uint32_t su[32];
          tmem_ld_32x32b_x32(s_addr + c0, su);
          tmem_ld_wait();
          float pv[32];
          uint32_t rowbits = 0, hbits = 0;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float x = fmaf(__uint_as_float(su[j]), sc, neg_ref);
            if (HAS_BIAS) x = fmaf(gl, tabrow[k0 + c0 + j], x);
            if (msk) x += kbias[k0 + c0 + j];
            over = fmaxf(over, x);
            const float e = fast_exp2(x);
            lsum += e;  // the softmax normaliser is taken before dropout
            if (DROP) {
              if ((j & 1) == 0) hbits = drop_bits(rk0, rk1, static_cast<uint32_t>(k0 + c0 + j) >> 1);
              const bool keep = (j & 1) ? drop_keep_hi(hbits, p.drop_thr_hi) : drop_keep_lo(hbits, p.drop_thr_hi);
              if (keep) rowbits |= (1u << j);  // this row's decisions for the 32 key columns
              pv[j] = keep ? e : 0.f;
            } else {
              pv[j] = e;
            }
          }
          if (DROP) {
            // the backward walks key-major: store, per key column, one word whose bit l is the decision of query row l of this warp
            const uint32_t mword = warp_bit_transpose(rowbits, tid & 31);
            if (mask_row != nullptr) mask_row[k0 + c0 + (tid & 31)] = mword;
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 w;
            w.x = pack_bf16x2(pv[g * 8 + 0], pv[g * 8 + 1]);
            w.y = pack_bf16x2(pv[g * 8 + 2], pv[g * 8 + 3]);
            w.z = pack_bf16x2(pv[g * 8 + 4], pv[g * 8 + 5]);
            w.w = pack_bf16x2(pv[g * 8 + 6], pv[g * 8 + 7]);
            store_sw128_chunk(sPw, r, (c0 >> 3) + g, w);
          }
        }
        if (fast && __any_sync(0xffffffffu, over > 64.0f)) {
          fast = false;  // re-base this tile on its own maximum (the P tile is simply rewritten)
          continue;
        }
        break;
      }
      l_run = l_run * alpha_cur + lsum;
      m_run = m_new;
      alpha_prev = alpha_cur;

      fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc_fence_before();
      named_bar_sync(1 + wg, kAttnTile);  // this warpgroup only: P complete, S and O tile fully read
      if (issuer) {
        tc_fence_after();
        mbar_wait(&v_full[n & 1], (n >> 1) & 1);
        tc_fence_after();
        const uint32_t a = smem_u32(sPw), bb = smem_u32(sV + (n & 1) * 16384);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_bf16(tmem_o, make_smem_desc_sw128(a + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024),
                    make_smem_desc_sw128(bb + k * 2048, 8192, 1024), idesc_pv, k > 0 ? 1u : 0u);
        umma_commit(&o_full[wg]);
        umma_commit(&v_empty[n & 1]);
        if (n + 1 < N) {
          mbar_wait(&k_full[(n + 1) & 1], ((n + 1) >> 1) & 1);
          tc_fence_after();
          issue_s(n + 1);
        }
      }
      __syncwarp();
    }
    mbar_wait(&o_full[wg], (N - 1) & 1);
    tc_fence_after();
    accumulate_o();

    if (row_valid) {
      const float inv = l_run > 0.f ? (DROP ? p.drop_rp : 1.0f) / l_run : 0.f;
      __nv_bfloat16* dst = p.out + (static_cast<long long>(b) * T + q0 + r256) * D + h * kHeadDim;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        uint4 w;
        w.x = pack_bf16x2(o[g * 8 + 0] * inv, o[g * 8 + 1] * inv);
        w.y = pack_bf16x2(o[g * 8 + 2] * inv, o[g * 8 + 3] * inv);
        w.z = pack_bf16x2(o[g * 8 + 4] * inv, o[g * 8 + 5] * inv);
        w.w = pack_bf16x2(o[g * 8 + 6] * inv, o[g * 8 + 7] * inv);
        *reinterpret_cast<uint4*>(dst + g * 8) = w;
      }
      if (p.lse != nullptr)
        p.lse[(static_cast<long long>(b) * p.H + h) * T + q0 + r256] = (l_run > 0.f) ? (m_run + log2f(l_run)) : INFINITY;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc(tmem, 512);
  }
}

int make_qkv_tmap(CUtensorMap* out, const void* qkv, int T, int B, int D3, int box_rows);

}  // namespace b200

using namespace b200;

extern "C" {

// out[b,t,h*64+d] = softmax_j(scale q.k + gate*tab[j-i], key padding) v     (WavLM/modules.py:540-563 replaced)
// qkv: bf16 [B,T,3D] fused projection output; gate: fp32 [B,H,T] or NULL; tab: fp32 [H,2T-1] or NULL (no bias);
// key_pad: uint8 [B,T] or NULL; out: bf16 [B,T,D]; lse: fp32 [B,H,T] (log2-domain log-sum-exp, saved for backward).
int b200s_attn_fwd_dropout(const void* qkv, const float* gate, const float* tab, const uint8_t* key_pad, void* out, float* lse,
                           int B, int T, int H, float scale, float drop_p, uint32_t key0, uint32_t key1, uint32_t* drop_mask,
                           b200s_stream stream) {
  B200_CHECK_ARG(qkv && out, "attn_fwd: null pointer");
  B200_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f, "attn_fwd: dropout p=%f out of range [0,1)", static_cast<double>(drop_p));
  B200_CHECK_ARG(drop_p == 0.f || drop_mask != nullptr, "attn_fwd: dropout needs the mask buffer (b200s_attn_dropout_mask_words)");
  B200_CHECK_ARG(static_cast<long long>(B) * H * T < (1LL << 32), "attn_fwd: B*H*T exceeds the 32-bit dropout row counter");
  B200_CHECK_ARG(T >= 1 && T <= 4096, "attn_fwd: T=%d out of range (1..4096)", T);
  const int D = H * kHeadDim;
  CUtensorMap tm;
  if (make_qkv_tmap(&tm, qkv, T, B, 3 * D, kAttnTile)) return -3;
  AttnParams p;
  memset(&p, 0, sizeof(p));
  p.T = T; p.H = H; p.B = B; p.D = D;
  p.n_tiles = ceil_div(T, kAttnTile);
  p.scale = scale;
  p.gate = gate; p.tab = tab; p.key_pad = key_pad;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.lse = lse;
  const bool drop = drop_p > 0.f;
  p.drop_mask = drop_mask;
  p.drop_k0 = key0; p.drop_k1 = key1;
  p.drop_thr_hi = drop_threshold16(drop_p) << 16;
  p.drop_rp = 1.0f / (1.0f - drop_p);
  const int smem = kFwdTab + sizeof(float) * ((p.n_tiles + 2) * kAttnTile + p.n_tiles * kAttnTile) +
                   sizeof(int) * p.n_tiles + 1024;
  dim3 grid(ceil_div(T, 2 * kAttnTile), H, B);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  void (*kern)(const CUtensorMap, const AttnParams) =
      tab != nullptr ? (drop ? attn_fwd_kernel<true, true> : attn_fwd_kernel<true, false>)
                     : (drop ? attn_fwd_kernel<false, true> : attn_fwd_kernel<false, false>);
  B200_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  B200_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(kFwdThreads), smem, st, tm, p));
  B200_CHECK_LAUNCH();
  return 0;
}

int b200s_attn_fwd(const void* qkv, const float* gate, const float* tab, const uint8_t* key_pad, void* out, float* lse,
                   int B, int T, int H, float scale, b200s_stream stream) {
  return b200s_attn_fwd_dropout(qkv, gate, tab, key_pad, out, lse, B, T, H, scale, 0.f, 0u, 0u, nullptr, stream);
}

long long b200s_attn_dropout_mask_words(int B, int T, int H) { return attn_drop_mask_words(B, H, T); }

}  // extern "C"
