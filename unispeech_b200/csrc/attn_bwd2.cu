// Fused attention backward with the gated relative-position bias (autograd of WavLM/modules.py:521-563): ONE tensor-core
// kernel produces dQ, dK, dV, d gate and d tab, so the probabilities are recomputed once (the two-kernel version in
// attn_bwd.cu recomputes S, dP and the exponentials twice, and its tensor-core and CUDA-core phases never overlap).
//
// CTA = 128 keys of one (batch, head); it walks over the queries in HALF tiles of 64 (transposed orientation: thread = key
// row, so dK / dV accumulate in TMEM over the whole loop and need no cross-thread reduction):
//   S^T  = K Q_h^T          128x64x64   -> TMEM stage h&1, columns [0,64)
//   dP^T = V dO_h^T         128x64x64   -> TMEM stage h&1, columns [64,128)
//   P^T  = exp2(S^T*scale*log2e + gate_i*log2e*tab[j-i] + keymask_j - lse_i);   dS^T = P^T o (dP^T - Delta_i)
//   dV  += P^T  dO_h        (P^T, dS^T written once to shared memory as bf16 K-major operand tiles)
//   dK  += dS^T Q_h * scale
//   dQ_i = dS K * scale     once per full 128-query tile (dS^T read as an MN-major A operand), accumulator read back from
//                           TMEM and added to an fp32 [B,T,D] buffer with vector reductions (one writer CTA per key tile)
//   d gate_i = sum_j dS_ij tab[j-i]        column sums over the key rows: warp butterfly + shared-memory accumulators
//   d tab[d] = sum_i gate_i dS_{i,i+d}     diagonal sums of a staged bf16 tile (double buffered), per-CTA accumulators
// Warp roles: warps 0-7 = CUDA-core warps (thread t: key row t & 127, query columns 32*(t>>7).. of the half tile);
// warps 8-11 = producer warpgroup (lane 0 of warp 8 issues every TMA load and MMA; the group only exists so that
// setmaxnreg can hand its registers to the CUDA-core warpgroups).  All hand-offs are mbarriers, there is no __syncthreads in the loop; the
// S^T/dP^T accumulators are double buffered in TMEM so the MMAs of half tile n+1 run under the exponentials of n.
#include "../../include/unispeech_b200.h"
#include "attn_common.cuh"
#include "common.h"

namespace b200 {

namespace {

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void mbar_arrive_cta(uint64_t* bar) {
  asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// shared-memory map (bytes from the 1024-aligned base)
constexpr int kFK = 0;             // K tile          16 KB
constexpr int kFV = 16384;         // V tile          16 KB
constexpr int kFQ = 32768;         // Q tiles, 2 stages x 16 KB
constexpr int kFDO = 65536;        // dO tiles, 2 stages x 16 KB
constexpr int kFPT = 98304;        // P^T  : two [128 keys][64 queries] blocks (one per half tile), 32 KB
constexpr int kFDST = 131072;      // dS^T : same layout, 32 KB
constexpr int kFW = 163840;        // gate*dS^T staging for the diagonal sums: 2 x [128][66] bf16
constexpr int kWStride2 = 66;      // bf16 per staged row (33 words: conflict-free row writes and diagonal reads)
constexpr int kFWBytes = 128 * kWStride2 * 2;   // 16896
constexpr int kFVec = kFW + 2 * kFWBytes;        // 197632: colvec [2][128] float4
constexpr int kFTab = kFVec + 2 * 128 * 16;      // 201728: tab_s[(N+1)*128], dtab_acc[(N+1)*128], dgate_s[N*128]
constexpr int kNC = 4;                        // threads per key row: each handles kCW query columns of a 64-wide half tile
constexpr int kCW = 64 / kNC;                 // 16
constexpr int kCudaThreads = 128 * kNC;       // 16 CUDA-core warps: four per scheduler hide the TMEM / SFU / LDS latencies
constexpr int kFThreads = kCudaThreads + 128; // + 1 producer warpgroup (only its first lane works)
constexpr int kProdWarp = kCudaThreads / 32;

}  // namespace

template <bool HAS_BIAS, bool DROP>
__global__ void __launch_bounds__(kFThreads, 1) attn_bwd_fused_kernel(const __grid_constant__ CUtensorMap tm_qkv,
                                                                      const __grid_constant__ CUtensorMap tm_do,
                                                                      const __grid_constant__ AttnParams p,
                                                                      float* __restrict__ dq_acc) {
  pdl_grid_sync();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int k0 = blockIdx.x * kAttnTile, h = blockIdx.y, b = blockIdx.z;
  const int T = p.T, D = p.D, N = p.n_tiles;
  const int NH = 2 * N;  // half tiles

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // 1024-aligned, still a __shared__ pointer (LDS/STS, not generic)
  uint8_t* sK = smem + kFK;
  uint8_t* sV = smem + kFV;
  uint8_t* sQ = smem + kFQ;
  uint8_t* sDO = smem + kFDO;
  uint8_t* sPT = smem + kFPT;
  uint8_t* sDST = smem + kFDST;
  float4* colvec = reinterpret_cast<float4*>(smem + kFVec);  // [2][128] {lse2, delta, gate*log2e, gate}
  float* tab_s = reinterpret_cast<float*>(smem + kFTab);     // [(N+1)*128]
  float* dtab_acc = tab_s + (N + 1) * kAttnTile;             // [(N+1)*128]
  float* dgate_s = dtab_acc + (N + 1) * kAttnTile;           // [N*128]

  __shared__ uint64_t kv_full, qdo_full[2], qdo_free[2], st_full[2], ready[2], mma_done[2], dq_full, acc_done;
  __shared__ uint32_t tmem_base_s;

  if (tid == 0) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_do);
    mbar_init(&kv_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&qdo_full[i], 1);
      mbar_init(&qdo_free[i], 1);
      mbar_init(&st_full[i], 1);
      mbar_init(&ready[i], kCudaThreads);
      mbar_init(&mma_done[i], 1);
    }
    mbar_init(&dq_full, 1);
    mbar_init(&acc_done, 1);
    fence_mbar_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);

  // per-CTA tables: tab_s[l] = tab[h, l + base], base = k0 - (N*128-1) + (T-1); element (key row r, query i) -> l = r + N*128-1 - i
  const int tab_base = k0 - (N * kAttnTile - 1) + (T - 1);
  if (HAS_BIAS) {
    const int len = (N + 1) * kAttnTile;
    for (int l = tid; l < len; l += kFThreads) {
      const int gi = l + tab_base;
      tab_s[l] = (gi >= 0 && gi < 2 * T - 1) ? p.tab[static_cast<long long>(h) * (2 * T - 1) + gi] : 0.f;
      dtab_acc[l] = 0.f;
    }
    for (int l = tid; l < N * kAttnTile; l += kFThreads) dgate_s[l] = 0.f;
  }
  auto load_colvec = [&](int qi) {  // executed by threads 0..127: one query column each
    const int i = qi * kAttnTile + tid;
    float4 v;
    if (i < T) {
      const long long idx = (static_cast<long long>(b) * p.H + h) * T + i;
      const float g = HAS_BIAS ? ((p.gate != nullptr) ? p.gate[idx] : 1.0f) : 0.f;
      v.x = p.lse[idx];
      v.y = p.delta[idx];
      v.z = g * kLog2e;
      v.w = g;
    } else {
      v.x = INFINITY;  // p = exp2(-inf) = 0 for out-of-range queries
      v.y = 0.f;
      v.z = 0.f;
      v.w = 0.f;
    }
    colvec[(qi & 1) * kAttnTile + tid] = v;
  };
  if (tid < kAttnTile) load_colvec(0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  // TMEM columns: S^T/dP^T stage s at s*128 (S^T) and s*128+64 (dP^T); dV 256; dK 320; dQ 384
  constexpr uint32_t kColDV = 256, kColDK = 320, kColDQ = 384;

  if (warp >= kProdWarp) {
    // registers are granted per warpgroup (96 per thread at launch): the producer group keeps 40, the CUDA-core groups get 104
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == kProdWarp && lane == 0) {
      // ================================================================== TMA producer + MMA issuer
      constexpr uint32_t idesc_st = make_idesc_bf16(128, 64, 0, 0);   // K-major A (K/V), K-major B (Q/dO half tile)
      constexpr uint32_t idesc_acc = make_idesc_bf16(128, 64, 0, 1);  // K-major A (P^T/dS^T), MN-major B (dO/Q)
      constexpr uint32_t idesc_dq = make_idesc_bf16(128, 64, 1, 1);   // MN-major A (dS^T read as dS), MN-major B (K)
      auto load_qdo = [&](int qi) {
        const int s = qi & 1;
        mbar_expect_tx(&qdo_full[s], 32768);
        tma_load_4d(sQ + s * 16384, &tm_qkv, &qdo_full[s], h * kHeadDim, qi * kAttnTile, b, 0);
        tma_load_4d(sDO + s * 16384, &tm_do, &qdo_full[s], h * kHeadDim, qi * kAttnTile, b, 0);
      };
      auto issue_st = [&](int hh) {  // S^T and dP^T of half tile hh into TMEM stage hh & 1
        const int qi = hh >> 1, hf = hh & 1;
        const uint32_t ak = smem_u32(sK), av = smem_u32(sV);
        const uint32_t bq = smem_u32(sQ + (qi & 1) * 16384 + hf * 8192), bd = smem_u32(sDO + (qi & 1) * 16384 + hf * 8192);
        const uint32_t d0 = tmem + hf * 128;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(d0, make_smem_desc_sw128(ak + k * 32, 16, 1024), make_smem_desc_sw128(bq + k * 32, 16, 1024), idesc_st,
                    k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(d0 + 64, make_smem_desc_sw128(av + k * 32, 16, 1024), make_smem_desc_sw128(bd + k * 32, 16, 1024),
                    idesc_st, k > 0 ? 1u : 0u);
        umma_commit(&st_full[hf]);
      };

      mbar_expect_tx(&kv_full, 32768);
      tma_load_4d(sK, &tm_qkv, &kv_full, D + h * kHeadDim, k0, b, 0);
      tma_load_4d(sV, &tm_qkv, &kv_full, 2 * D + h * kHeadDim, k0, b, 0);
      load_qdo(0);
      if (N > 1) load_qdo(1);
      mbar_wait(&kv_full, 0);
      mbar_wait(&qdo_full[0], 0);
      tc_fence_after();
      issue_st(0);
      issue_st(1);

      for (int hh = 0; hh < NH; ++hh) {
        const int qi = hh >> 1, hf = hh & 1, st = qi & 1;
        mbar_wait(&ready[hf], qi & 1);  // P^T / dS^T of this half tile are in shared memory; TMEM stage hf has been read
        tc_fence_after();
        const uint32_t apt = smem_u32(sPT + hf * 16384), ads = smem_u32(sDST + hf * 16384);
        const uint32_t bdo = smem_u32(sDO + st * 16384 + hf * 8192), bq = smem_u32(sQ + st * 16384 + hf * 8192);
#pragma unroll
        for (int k = 0; k < 4; ++k)  // dV += P^T dO   (K = 64 queries)
          umma_bf16(tmem + kColDV, make_smem_desc_sw128(apt + k * 32, 16, 1024),
                    make_smem_desc_sw128(bdo + k * 2048, 8192, 1024), idesc_acc, (hh > 0 || k > 0) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k)  // dK += dS^T Q
          umma_bf16(tmem + kColDK, make_smem_desc_sw128(ads + k * 32, 16, 1024),
                    make_smem_desc_sw128(bq + k * 2048, 8192, 1024), idesc_acc, (hh > 0 || k > 0) ? 1u : 0u);
        if (hf == 1) {
          // dQ_i = dS K over the full 128-query tile: A = dS^T tile read MN-major (M = queries: two 64-wide atoms 16 KB
          // apart, K = key rows: 16 rows = 2048 B per step), B = K tile MN-major
          const uint32_t adq = smem_u32(sDST), bk = smem_u32(sK);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_bf16(tmem + kColDQ, make_smem_desc_sw128(adq + k * 2048, 16384, 1024),
                      make_smem_desc_sw128(bk + k * 2048, 8192, 1024), idesc_dq, k > 0 ? 1u : 0u);
          umma_commit(&dq_full);
          umma_commit(&qdo_free[st]);  // every MMA reading Q_i / dO_i has been issued before this commit
        }
        umma_commit(&mma_done[hf]);
        if (hh + 2 < NH) {
          const int q2 = (hh + 2) >> 1;
          if (hf == 0) {  // first half of a new query tile: its TMA load must have landed
            mbar_wait(&qdo_full[q2 & 1], (q2 >> 1) & 1);
            tc_fence_after();
          }
          issue_st(hh + 2);
        } else if (hh + 1 == NH) {
          umma_commit(&acc_done);
        }
        if (hf == 1 && qi + 2 < N) {  // refill this Q/dO stage once its readers have retired
          mbar_wait(&qdo_free[st], (qi >> 1) & 1);
          load_qdo(qi + 2);
        }
      }
    }
  } else {
    // ==================================================================== CUDA-core warps
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    const int r = tid & (kAttnTile - 1);  // key row inside the tile == TMEM lane
    const int ch = tid >> 7;              // which kCW query columns of the 64-wide half tile
    const int key = k0 + r;
    const bool key_valid = key < T;
    const bool key_masked = !key_valid || (p.key_pad != nullptr && p.key_pad[static_cast<long long>(b) * T + key] != 0);
    const float kb = key_masked ? -INFINITY : 0.f;
    const float sc = p.scale * kLog2e;
    const float* tabrow = tab_s + r + N * kAttnTile - 1;
    const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
    // dropout keep bits written by the forward kernel: word (query block i >> 5, key) holds the bits of 32 consecutive queries
    const uint32_t* mask_col =
        DROP ? p.drop_mask + static_cast<long long>(b * p.H + h) * (4 * N) * (N * kAttnTile) + k0 + r : nullptr;

    // dQ tile of query tile qi: TMEM -> fp32 reductions into dq_acc[b, q, h*64 + ch*32 ..]
    auto flush_dq = [&](int qi) {
      mbar_wait(&dq_full, qi & 1);
      tc_fence_after();
      uint32_t t0[kCW];
      tmem_ld_32x32b_x16(tmem + lane_addr + kColDQ + ch * kCW, t0);
      tmem_ld_wait();
      const int q = qi * kAttnTile + r;  // TMEM lane = query row of the dQ accumulator
      if (q < T) {
        float* dst = dq_acc + (static_cast<long long>(b) * T + q) * D + h * kHeadDim + ch * kCW;
#pragma unroll
        for (int g = 0; g < kCW / 4; ++g)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + g * 4), "f"(__uint_as_float(t0[g * 4 + 0])),
                       "f"(__uint_as_float(t0[g * 4 + 1])), "f"(__uint_as_float(t0[g * 4 + 2])),
                       "f"(__uint_as_float(t0[g * 4 + 3]))
                       : "memory");  // the staged dS^T already carries the softmax scale
      }
    };
    // diagonal sums of the staged gate*dS^T tile of half tile hh (thread: diagonal d = r, query columns 32*ch..)
    // Element (key row rr, local query c) of the staged tile belongs to diagonal rr - c; thread (d = r, ch) walks
    // rr = (d + c) & 127 over its kCW columns: the wrap point is a per-thread constant, so every load is base + immediate.
    const int diag_w0 = kAttnTile - r - ch * kCW;  // columns e >= diag_w0 (e = c - ch*kCW) are on the wrapped diagonal
    const uint32_t diag_base = static_cast<uint32_t>(r * kWStride2 + ch * kCW * (kWStride2 + 1)) * 2u;
    auto diag_sums = [&](int hh) {
      const uint32_t w_nowrap = smem_u32(smem + kFW + (hh & 1) * kFWBytes) + diag_base;
      const uint32_t w_wrap = w_nowrap - static_cast<uint32_t>(kAttnTile * kWStride2 * 2);
      float acc_all = 0.f, acc_pos = 0.f;
#pragma unroll
      for (int e = 0; e < kCW; ++e) {
        const bool wrapped = e >= diag_w0;
        uint32_t v16;
        asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v16) : "r"((wrapped ? w_wrap : w_nowrap) + e * (kWStride2 + 1) * 2));
        const float v = __uint_as_float(v16 << 16);
        acc_all += v;
        if (!wrapped) acc_pos += v;
      }
      const int l_pos = r + N * kAttnTile - 1 - (hh >> 1) * kAttnTile - (hh & 1) * 64;
      atomicAdd(&dtab_acc[l_pos], acc_pos);
      if (diag_w0 < kCW) atomicAdd(&dtab_acc[l_pos - kAttnTile], acc_all - acc_pos);
    };

    for (int hh = 0; hh < NH; ++hh) {
      const int qi = hh >> 1, hf = hh & 1;
      const int i0 = qi * kAttnTile + hf * 64 + ch * kCW;  // first global query of this thread's columns
      mbar_wait(&st_full[hf], qi & 1);
      tc_fence_after();
      uint32_t keep_bits = 0xffffffffu;
      if (DROP) keep_bits = mask_col[static_cast<long long>(i0 >> 5) * (N * kAttnTile)] >> (i0 & 31);  // bit e = query i0 + e
      uint32_t su[kCW], du[kCW];
      tmem_ld_32x32b_x16(tmem + lane_addr + hf * 128 + ch * kCW, su);
      tmem_ld_32x32b_x16(tmem + lane_addr + hf * 128 + 64 + ch * kCW, du);
      if (hf == 1 && qi >= 1) flush_dq(qi - 1);  // dQ of the previous query tile finished a whole phase ago
      tmem_ld_wait();
      const float4* cv = colvec + (qi & 1) * kAttnTile + hf * 64 + ch * kCW;
      uint32_t pw[kCW / 2], dw[kCW / 2], ww[kCW / 2];
      float dgc[kCW];
#pragma unroll
      for (int j = 0; j < kCW; j += 2) {
        float pr2[2], ds2[2], w2[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float4 c = cv[j + e];
          float x = fmaf(__uint_as_float(su[j + e]), sc, kb);
          float tb = 0.f;
          if (HAS_BIAS) {
            tb = tabrow[-(i0 + j + e)];
            x = fmaf(c.z, tb, x);
          }
          const float pr = ex2f(x - c.x);
          float dpv = __uint_as_float(du[j + e]);
          bool keep = true;
          if (DROP) {  // O = (P o M) V / (1-p):  dP = M o (dO V^T) / (1-p);  dV takes the dropped probabilities (scaled at the end)
            keep = ((keep_bits >> (j + e)) & 1u) != 0u;
            dpv = keep ? dpv * p.drop_rp : 0.f;
          }
          const float ds = pr * (dpv - c.y);
          pr2[e] = keep ? pr : 0.f;
          ds2[e] = ds * p.scale;
          w2[e] = c.w * ds;
          dgc[j + e] = ds * tb;
        }
        pw[j >> 1] = pack_bf16x2(pr2[0], pr2[1]);
        dw[j >> 1] = pack_bf16x2(ds2[0], ds2[1]);
        ww[j >> 1] = pack_bf16x2(w2[0], w2[1]);
      }
      if (HAS_BIAS) {
        // d gate: sum over the 32 key rows of this warp for each of its query columns, then one shared atomic per column
        const float csum = warp_colsum16(dgc, lane);
        if ((lane & 1) == 0) atomicAdd(&dgate_s[i0 + (lane >> 1)], csum);
      }
      if (hh >= 1) {
        // Every thread has staged half tile hh-1 (needed by its diagonal sums; the same wait orders the reuse of the double-
        // buffered staging tile and of the colvec buffers).  Waiting HERE, a whole phase after the arrivals, means the warps
        // never run in lock step: a warp may be up to one phase ahead of the slowest one.
        mbar_wait(&ready[(hh - 1) & 1], ((hh - 1) >> 1) & 1);
        if (HAS_BIAS) diag_sums(hh - 1);
        // MMAs that read the P^T / dS^T blocks (and, in order, everything issued before them) have retired
        mbar_wait(&mma_done[(hh - 1) & 1], ((hh - 1) >> 1) & 1);
      }
#pragma unroll
      for (int g = 0; g < kCW / 8; ++g) {
        store_sw128_chunk(sPT, r, hf * 8 + ch * (kCW / 8) + g, make_uint4(pw[g * 4], pw[g * 4 + 1], pw[g * 4 + 2], pw[g * 4 + 3]));
        store_sw128_chunk(sDST, r, hf * 8 + ch * (kCW / 8) + g, make_uint4(dw[g * 4], dw[g * 4 + 1], dw[g * 4 + 2], dw[g * 4 + 3]));
      }
      if (HAS_BIAS) {
        uint32_t* wrow = reinterpret_cast<uint32_t*>(smem + kFW + (hh & 1) * kFWBytes) + r * (kWStride2 / 2) + ch * (kCW / 2);
#pragma unroll
        for (int j = 0; j < kCW / 2; ++j) wrow[j] = ww[j];
      }
      if (hf == 0 && qi + 1 < N && tid < kAttnTile) load_colvec(qi + 1);  // other buffer: last read in query tile qi-1
      fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc_fence_before();
      mbar_arrive_cta(&ready[hf]);
    }
    // ---- tail: last staged tile's diagonals, last dQ tile, then the dK / dV accumulators
    mbar_wait(&ready[(NH - 1) & 1], ((NH - 1) >> 1) & 1);
    if (HAS_BIAS) diag_sums(NH - 1);
    flush_dq(N - 1);
    mbar_wait(&acc_done, 0);
    tc_fence_after();
    {
      // 128 key rows x (64 dV + 64 dK) columns over 512 threads: ch 0,1 -> dV columns 32*(ch&1).., ch 2,3 -> dK
      uint32_t t0[32];
      const uint32_t col = ((ch < kNC / 2) ? kColDV : kColDK) + (ch & (kNC / 2 - 1)) * (128 / kNC);
      tmem_ld_32x32b_x32(tmem + lane_addr + col, t0);
      tmem_ld_wait();
      if (DROP && ch < kNC / 2) {  // dV = (P o M)^T dO / (1-p)
#pragma unroll
        for (int i = 0; i < 32; ++i) t0[i] = __float_as_uint(__uint_as_float(t0[i]) * p.drop_rp);
      }
      if (key_valid) {
        __nv_bfloat16* dst = p.dqkv + (static_cast<long long>(b) * T + key) * (3 * D) + ((ch < kNC / 2) ? 2 * D : D) +
                             h * kHeadDim + (ch & (kNC / 2 - 1)) * (128 / kNC);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(t0[g * 8 + 0]), __uint_as_float(t0[g * 8 + 1]));
          w.y = pack_bf16x2(__uint_as_float(t0[g * 8 + 2]), __uint_as_float(t0[g * 8 + 3]));
          w.z = pack_bf16x2(__uint_as_float(t0[g * 8 + 4]), __uint_as_float(t0[g * 8 + 5]));
          w.w = pack_bf16x2(__uint_as_float(t0[g * 8 + 6]), __uint_as_float(t0[g * 8 + 7]));
          *reinterpret_cast<uint4*>(dst + g * 8) = w;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (HAS_BIAS) {
    // per-CTA accumulators -> global (the relative-position table is shared by all layers: atomics; d gate: one CTA per key tile)
    if (p.dtab != nullptr) {
      for (int l = tid; l < (N + 1) * kAttnTile; l += kFThreads) {
        const int gi = l + tab_base;
        const float v = dtab_acc[l];
        if (gi >= 0 && gi < 2 * T - 1 && v != 0.f) atomicAdd(p.dtab + static_cast<long long>(h) * (2 * T - 1) + gi, v);
      }
    }
    if (p.dgate != nullptr) {
      for (int i = tid; i < N * kAttnTile; i += kFThreads)
        if (i < T) atomicAdd(p.dgate + (static_cast<long long>(b) * p.H + h) * T + i, dgate_s[i]);
    }
  }
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc(tmem, 512);
  }
}

// Delta_i = sum_d dO_id O_id (fp32 [B,H,T]); also clears d gate, which the fused kernel accumulates with atomics.
__global__ void __launch_bounds__(256) attn_delta2_kernel(const __nv_bfloat16* __restrict__ o,
                                                          const __nv_bfloat16* __restrict__ dout, int B, int T, int H,
                                                          float* __restrict__ delta, float* __restrict__ dgate) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (row >= static_cast<long long>(B) * T) return;
  const int D = H * kHeadDim;
  const long long b = row / T, t = row % T;
  for (int h = 0; h < H; ++h) {
    const float2 a = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(o + row * D + h * kHeadDim + lane * 2));
    const float2 g = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(dout + row * D + h * kHeadDim + lane * 2));
    const float s = warp_sum(a.x * g.x + a.y * g.y);
    if (lane == 0) {
      delta[(b * H + h) * T + t] = s;
      if (dgate != nullptr) dgate[(b * H + h) * T + t] = 0.f;
    }
  }
}

// dq_acc (fp32 [B*T, D]) -> bf16 into the q columns of dqkv [B*T, 3D]; the accumulator is cleared for the next layer.
__global__ void __launch_bounds__(256) attn_dq_convert_kernel(float* __restrict__ dq_acc, __nv_bfloat16* __restrict__ dqkv,
                                                              long long rows, int D) {
  pdl_grid_sync();
  const int vec_per_row = D / 8;
  const long long n = rows * vec_per_row;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / vec_per_row;
    const int c = static_cast<int>(i % vec_per_row) * 8;
    float4* src = reinterpret_cast<float4*>(dq_acc + row * D + c);
    const float4 a = src[0], bq = src[1];
    uint4 w;
    w.x = pack_bf16x2(a.x, a.y); w.y = pack_bf16x2(a.z, a.w);
    w.z = pack_bf16x2(bq.x, bq.y); w.w = pack_bf16x2(bq.z, bq.w);
    *reinterpret_cast<uint4*>(dqkv + row * 3 * D + c) = w;
    src[0] = make_float4(0.f, 0.f, 0.f, 0.f);
    src[1] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

int make_qkv_tmap(CUtensorMap* out, const void* qkv, int T, int B, int D3, int box_rows);

}  // namespace b200

using namespace b200;

extern "C" {

// Fused backward of b200s_attn_fwd.  Same contract as b200s_attn_bwd plus dq_acc: fp32 [B,T,D] workspace that must be ZERO
// on entry and is zero again on return (the q gradient is reduced there across key tiles before it is rounded to bf16).
int b200s_attn_bwd_fused_dropout(const void* qkv, const void* out, const void* dout, const float* gate, const float* tab,
                                 const uint8_t* key_pad, const float* lse, float* delta, float* dq_acc, void* dqkv,
                                 float* dgate, float* dtab, int B, int T, int H, float scale, float drop_p,
                                 const uint32_t* drop_mask, b200s_stream stream) {
  B200_CHECK_ARG(qkv && out && dout && lse && delta && dqkv && dq_acc, "attn_bwd_fused: null pointer");
  B200_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f, "attn_bwd_fused: dropout p=%f out of range [0,1)", static_cast<double>(drop_p));
  B200_CHECK_ARG(drop_p == 0.f || drop_mask != nullptr, "attn_bwd_fused: dropout needs the mask written by b200s_attn_fwd_dropout");
  B200_CHECK_ARG(T >= 1 && T <= 2048, "attn_bwd_fused: T=%d out of range (1..2048)", T);
  B200_CHECK_ARG(!tab || (dgate && dtab), "attn_bwd_fused: bias given but dgate/dtab missing");
  const int D = H * kHeadDim;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long rows = static_cast<long long>(B) * T;
  B200_CHECK_CUDA(launch_pdl(attn_delta2_kernel, dim3(static_cast<unsigned>(ceil_div_ll(rows * 32, 256))), dim3(256), 0, st, 
      static_cast<const __nv_bfloat16*>(out), static_cast<const __nv_bfloat16*>(dout), B, T, H, delta,
      tab != nullptr ? dgate : nullptr));
  B200_CHECK_LAUNCH();

  CUtensorMap tm_qkv, tm_do;
  if (make_qkv_tmap(&tm_qkv, qkv, T, B, 3 * D, kAttnTile)) return -3;
  if (make_qkv_tmap(&tm_do, dout, T, B, D, kAttnTile)) return -3;
  AttnParams p;
  memset(&p, 0, sizeof(p));
  p.T = T; p.H = H; p.B = B; p.D = D;
  p.n_tiles = ceil_div(T, kAttnTile);
  p.scale = scale;
  p.gate = gate; p.tab = tab; p.key_pad = key_pad;
  p.lse = const_cast<float*>(lse);
  p.dout = static_cast<const __nv_bfloat16*>(dout);
  p.delta = delta;
  p.dqkv = static_cast<__nv_bfloat16*>(dqkv);
  p.dgate = dgate;
  p.dtab = dtab;
  const bool drop = drop_p > 0.f;
  p.drop_mask = const_cast<uint32_t*>(drop_mask);
  p.drop_rp = 1.0f / (1.0f - drop_p);
  const int N = p.n_tiles;
  const int smem = kFTab + sizeof(float) * ((N + 1) * kAttnTile * 2 + N * kAttnTile) + 1024;
  B200_CHECK_ARG(smem <= 232448 - 512, "attn_bwd_fused: T=%d needs %d bytes of shared memory", T, smem);
  dim3 grid(N, H, B);
  void (*kern)(const CUtensorMap, const CUtensorMap, const AttnParams, float*) =
      tab != nullptr ? (drop ? attn_bwd_fused_kernel<true, true> : attn_bwd_fused_kernel<true, false>)
                     : (drop ? attn_bwd_fused_kernel<false, true> : attn_bwd_fused_kernel<false, false>);
  B200_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  B200_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(kFThreads), smem, st, tm_qkv, tm_do, p, dq_acc));
  B200_CHECK_LAUNCH();
  const long long nvec = rows * (D / 8);
  const int blocks = static_cast<int>(std::min<long long>(ceil_div_ll(nvec, 256), 148 * 16));
  B200_CHECK_CUDA(launch_pdl(attn_dq_convert_kernel, dim3(blocks), dim3(256), 0, st, dq_acc, static_cast<__nv_bfloat16*>(dqkv), rows, D));
  B200_CHECK_LAUNCH();
  return 0;
}

int b200s_attn_bwd_fused(const void* qkv, const void* out, const void* dout, const float* gate, const float* tab,
                         const uint8_t* key_pad, const float* lse, float* delta, float* dq_acc, void* dqkv, float* dgate,
                         float* dtab, int B, int T, int H, float scale, b200s_stream stream) {
  return b200s_attn_bwd_fused_dropout(qkv, out, dout, gate, tab, key_pad, lse, delta, dq_acc, dqkv, dgate, dtab, B, T, H, scale,
                                      0.f, nullptr, stream);
}

}  // extern "C"
