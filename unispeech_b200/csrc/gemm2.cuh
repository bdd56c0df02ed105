// Persistent CTA-pair bf16 GEMM: tcgen05.mma cta_group::2, 256 x 256 output tile per pair, fp32 accumulators double-buffered
// in TMEM (2 x 256 columns) so the epilogue of one tile overlaps the main loop of the next.
//
// Why pairs: a 128x128 tile moves 64 flop per byte from L2 into shared memory and a 128x256 tile 85; at the measured
// tensor rate that is more than the L2 can deliver.  With cta_group::2 each CTA stages its 128 rows of A and HALF of the B
// tile (128 of the 256 N rows); the tensor cores of both SMs read both halves, so the pair moves 64 KB per 256x256x64
// MAC block (128 flop/B) and per-SM shared-memory traffic halves.
//
// One pair per two SMs (grid = #SM CTAs, cluster (2,1,1)); static round-robin tile scheduler, N fastest so that
// concurrently running pairs share the A rows in L2.  Warp roles per CTA:
//   warp 0  TMA producer (both CTAs; completion bytes of both land on the LEADER's full barrier)
//   warp 1  TMEM allocator; in the leader also the MMA issuer (commit multicasts to both CTAs' barriers)
//   warps 2-9  epilogue of this CTA's 128 accumulator rows (same fused tail as gemm.cuh), two warps per TMEM lane quadrant
// Operand majorness, TMA coordinate matrices, split-K and the epilogue flags are shared with gemm.cuh.
#pragma once
#include "gemm.cuh"

namespace b200 {

// Epilogue kinds of the pair kernel (compile-time: keeps each instantiation's code small -- the fused tail used to be one
// 58 KB body that thrashed the instruction cache).
enum : int {
  EK_LINEAR = 0,  // bf16 out = acc (+bias) (+in0) (+in1) (+in2)
  EK_GELU = 1,    // bf16 out2 = acc + bias (optional); out = gelu(acc + bias) (+inputs)
  EK_DGELU = 2,   // bf16 out = (acc (+bias)) * gelu'(in0) (+in1) (+in2), optional column sums
  EK_F32 = 3,     // fp32 out (+)= acc: vector reductions (split-K) or read-modify-write (single writer)
};

template <int KIND>
struct Gemm2Cfg {
  // bf16 kinds: 4 operand stages + two 4 KB staging buffers per epilogue warp (input prefetch + coalesced output);
  // fp32 kind (weight gradients): 6 stages + one 4 KB staging buffer per epilogue warp.
  static constexpr int kStages = (KIND == EK_F32) ? 6 : 4;
  static constexpr int kABytes = 128 * 128;  // 128 rows x 64 bf16 per CTA
  static constexpr int kBBytes = 128 * 128;  // this CTA's half (128 N rows) of the 256-wide B tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kEpiWarps = 8;
  static constexpr int kStageBufBytes = 4096;                       // 32 rows x 128 B, 16-byte chunks XOR-swizzled by row & 7
  static constexpr int kStageBufs = (KIND == EK_F32) ? 1 : 2;
  static constexpr int kStagingBytes = kEpiWarps * kStageBufs * kStageBufBytes;
  static constexpr int kBiasBytes = (KIND == EK_F32) ? 0 : kEpiWarps * 128 * 4;  // 128 fp32 bias values per epilogue warp
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + kBiasBytes + 1024;
  static constexpr int kThreads = 320;  // TMA warp + MMA warp + 8 epilogue warps
  static constexpr int kTileM = 256, kTileN = 256;
  static_assert(kSmemBytes + 1792 <= 232448, "pair GEMM exceeds the 227 KB shared-memory limit");
};

// ---------------------------------------------------------------- staged (coalesced) epilogue helpers
// A warp's staging buffer holds 32 rows x 128 bytes (64 bf16 or 32 fp32 columns).  Byte offset of 16-byte chunk c of row r:
// r*128 + ((c ^ (r & 7)) << 4) -- conflict-free both for "one row per lane" accesses (compute side) and for "8 lanes per row"
// accesses (global-memory side: a warp instruction then covers 4 full 128-byte lines instead of 32 partial sectors).
__device__ __forceinline__ uint32_t stg_off(int row, int chunk) {
  return static_cast<uint32_t>(row * 128 + ((chunk ^ (row & 7)) << 4));
}
__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// Where one epilogue warp works for one tile: 32 accumulator rows x 128 columns.
struct EpiTile {
  int mb;          // batch index of the tile
  long long row0;  // first row (within the batch) of this warp's 32 rows
  int rows_valid;  // 0..32
  int col0;        // first global output column of this warp's 128 columns
  int cols_valid;  // 0..128 (multiple of 8)
};

// Start the asynchronous, coalesced copy of one 32-row x 64-column bf16 block of an epilogue input into a staging buffer.
__device__ __noinline__ void stage_input_async(const EpiTensor t, const EpiTile et, int cc, uint32_t buf, int lane) {
  const int chunk = lane & 7;
  const int colc = cc * 64 + chunk * 8;
  if (colc >= et.cols_valid) return;
  const __nv_bfloat16* base = static_cast<const __nv_bfloat16*>(t.p) + et.mb * t.bs + et.row0 * t.ld + et.col0 + colc;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = i * 4 + (lane >> 3);
    if (row < et.rows_valid) cp_async16(buf + stg_off(row, chunk), base + row * t.ld);
  }
}

// Coalesced store of a staged 32 x 64 bf16 block (rows beyond rows_valid / columns beyond cols_valid are skipped).
__device__ __noinline__ void flush_bf16(const EpiTensor t, const EpiTile et, int cc, uint32_t buf, int lane) {
  const int chunk = lane & 7;
  const int colc = cc * 64 + chunk * 8;
  __nv_bfloat16* base = static_cast<__nv_bfloat16*>(t.p) + et.mb * t.bs + et.row0 * t.ld + et.col0 + colc;
  const bool col_ok = colc < et.cols_valid;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = i * 4 + (lane >> 3);
    const uint4 v = lds128(buf + stg_off(row, chunk));
    if (col_ok && row < et.rows_valid) *reinterpret_cast<uint4*>(base + row * t.ld) = v;
  }
}

// This lane's row of a staged bf16 block: acc[j] (op)= staged[lane][j], j = 0..63   (MUL_DGELU: multiply by gelu'(staged), or by
// the staged value itself when it already is the derivative)
template <bool MUL_DGELU>
__device__ __forceinline__ void consume_row(float* acc, uint32_t buf, int lane, bool aux_is_grad = false) {
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const uint4 w = lds128(buf + stg_off(lane, g));
    const uint32_t wu[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2(wu[j]);
      if (MUL_DGELU) {
        acc[g * 8 + 2 * j] *= aux_is_grad ? f.x : gelu_grad_f(f.x);
        acc[g * 8 + 2 * j + 1] *= aux_is_grad ? f.y : gelu_grad_f(f.y);
      } else {
        acc[g * 8 + 2 * j] += f.x;
        acc[g * 8 + 2 * j + 1] += f.y;
      }
    }
  }
}
__device__ __forceinline__ void stage_row_bf16(const float* acc, uint32_t buf, int lane) {
#pragma unroll
  for (int g = 0; g < 8; ++g) sts128(buf + stg_off(lane, g), pack_bf16x8(acc + g * 8));
}

// NPAIR = 1: cluster = one CTA pair.  NPAIR = 2: cluster = two CTA pairs working on M-adjacent tiles of the SAME N tile; every
// CTA loads one quarter of the B tile (64 of its 256 rows) and TMA-multicasts it to the CTA of the same rank in the other
// pair, so the B operand crosses the L2 -> SM fabric once per cluster instead of once per pair.  The kernel is bound by that
// fabric (about 6.9 KB/clk chip-wide, measured: the loads alone take longer than the MMAs), not by the tensor pipe.
template <bool A_MN, bool B_MN, int KIND, int NPAIR>
__global__ void __launch_bounds__(320, 1) gemm_bf16_pair_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                const __grid_constant__ CUtensorMap tmB,
                                                                const __grid_constant__ CUtensorMap tmO,
                                                                const __grid_constant__ CUtensorMap tmO2,
                                                                const __grid_constant__ GemmParams p) {
  pdl_launch_dependents();  // the next kernel's CTAs may start their prologue as soon as SMs free up
  using Cfg = Gemm2Cfg<KIND>;
  constexpr int kStages = Cfg::kStages;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();     // rank inside the cluster
  const uint32_t rank = crank & 1u;             // rank inside the CTA pair, 0 = leader (issues the MMAs)
  const int cp = static_cast<int>(crank >> 1);  // pair index inside the cluster
  const int pair = blockIdx.x / (2 * NPAIR);    // work unit (cluster) index
  const int n_pairs = gridDim.x / (2 * NPAIR);
  const int m_tiles_total = p.tiles_total / p.n_tiles;
  const int sup_tiles = ((m_tiles_total + NPAIR - 1) / NPAIR) * p.n_tiles;  // work items per K split
  // (stream-K: item = pair + seg * n_pairs is the seg-th tile segment of that pair's range, see decode)
  const int n_items = (NPAIR == 1 && KIND == EK_F32 && p.stream_k) ? n_pairs * p.splits : sup_tiles * p.splits;
  constexpr uint16_t kAllCtas = static_cast<uint16_t>((1u << (2 * NPAIR)) - 1u);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // 1024-aligned, still a __shared__ pointer (LDS/STS, not generic)
  __shared__ uint64_t full_bar[kStages];
  __shared__ uint64_t empty_bar[kStages];
  __shared__ uint64_t tmem_full_bar[2];
  __shared__ uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_smem;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (KIND != EK_F32) tma_prefetch_desc(&tmO);
#pragma unroll
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], NPAIR);  // one tcgen05.commit arrival per pair that reads (a multicast copy of) the stage
    }
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 16);  // 8 epilogue warps x 2 CTAs arrive on the leader's barrier
    }
    fence_mbar_init();
  }
  __syncwarp();
  if (warp == 1) tmem_alloc_2cta(&tmem_base_smem, 512);
  tc_fence_before();
  cluster_sync_all();  // barriers of BOTH CTAs initialised (and TMEM allocated) before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_wait();  // everything above overlapped the previous kernel's tail; global memory is touched only from here on

  // Ragged batches (p.m_valid: valid rows per batch of an activation GEMM; p.k_valid: the same for the reduction dimension of a
  // weight gradient).  Padding is a suffix of every batch, so the live units of batch b are its first live_b M tiles (or 64-row K
  // blocks); rag_pref[b] = live units of batches < b.  Work is then enumerated over LIVE units only -- a static round-robin over
  // the dense tile grid would leave most of the gain on the table (the dead tiles of a short utterance are consecutive items) --
  // and the dead M tiles, which only have to be zero-filled, are appended after the live ones.
  __shared__ uint16_t rag_pref[kMaxRagBatches + 1];
  const int* rag = p.m_valid != nullptr ? p.m_valid : p.k_valid;
  const int rag_unit = p.m_valid != nullptr ? p.m_tile_stride : 64;
  const int rag_per_batch = p.m_valid != nullptr ? p.m_tiles_per_batch : p.k_blocks_per_batch;
  const int rag_batches = rag == nullptr ? 0 : (p.m_valid != nullptr ? m_tiles_total : p.k_blocks) / rag_per_batch;
  if (rag != nullptr) {
    if (warp == 2) {
      int carry = 0;
      for (int b0 = 0; b0 < rag_batches; b0 += 32) {
        const int b = b0 + lane;
        int incl = 0;
        if (b < rag_batches) incl = min(rag_per_batch, max(0, (rag[b] + rag_unit - 1) / rag_unit));
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += t;
        }
        if (b < rag_batches) rag_pref[b + 1] = static_cast<uint16_t>(carry + incl);
        carry += __shfl_sync(0xffffffffu, incl, 31);
      }
      if (lane == 0) rag_pref[0] = 0;
    }
    __syncthreads();
  }
  const int rag_live = rag != nullptr ? rag_pref[rag_batches] : 0;
  // batch holding live unit r (0 <= r < rag_live) / dead unit d (0 <= d < total - rag_live)
  auto live_batch = [&](int r) {
    int lo = 0, hi = rag_batches;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (rag_pref[mid] <= r) lo = mid; else hi = mid;
    }
    return lo;
  };
  auto dead_batch = [&](int d) {
    int lo = 0, hi = rag_batches;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (mid * rag_per_batch - rag_pref[mid] <= d) lo = mid; else hi = mid;
    }
    return lo;
  };

  // item -> (split, m tile of THIS pair, n tile) and its K-block range; identical in every role of the pair's CTAs.
  // `active` is false for the second pair of a cluster when the M tile count is odd: it still loads and multicasts its part
  // of B and keeps the stage barriers flowing, but issues no MMA and has no epilogue.
  // `active` is also false for a DEAD tile of a ragged batch (p.m_valid: every row of the tile is padding): no loads, no MMAs, the
  // epilogue writes zeros (NPAIR = 1 only: the host never combines ragged activations with the multicast cluster).  For ragged
  // weight gradients (p.k_valid) [kb_begin, kb_end) is a range of LIVE K-block ranks, split evenly.
  auto decode = [&](int item, int& mb, int& m0, int& n_tile, int& kb_begin, int& kb_end, bool& active) {
    if (NPAIR == 1 && KIND == EK_F32 && p.stream_k) {
      // stream-K: this pair owns units [u0, u1) of the tile-major (tile, K block) space -- every pair gets the same number of K
      // blocks (64 tiles on 74 pairs would otherwise leave 10 pairs idle); the range is walked one tile segment per item.
      // K blocks are LIVE ranks for ragged batches.
      const int kbt = p.k_valid != nullptr ? rag_live : p.k_blocks;
      const int total = sup_tiles * kbt;
      const int per = (total + n_pairs - 1) / n_pairs;
      const int seg = item / n_pairs;
      const int u0 = min(total, (item - seg * n_pairs) * per), u1 = min(total, u0 + per);
      const int t0 = kbt > 0 ? u0 / kbt : 0;
      const int tile = min(t0 + seg, sup_tiles - 1);
      kb_begin = seg == 0 ? u0 - t0 * kbt : 0;
      kb_end = (t0 + seg < sup_tiles) ? min(kbt, u1 - (t0 + seg) * kbt) : 0;   // <= kb_begin: empty segment, skipped by every role
      if (kb_end < kb_begin) kb_end = kb_begin;
      n_tile = tile % p.n_tiles;
      const int mt = tile / p.n_tiles;
      active = true;
      mb = mt / p.m_tiles_per_batch;
      m0 = (mt % p.m_tiles_per_batch) * p.m_tile_stride;
      return;
    }
    const int split = item / sup_tiles;
    const int tile = item % sup_tiles;
    n_tile = tile % p.n_tiles;
    if (NPAIR == 1 && p.m_valid != nullptr) {
      const int tr = tile / p.n_tiles;
      int mt_in;
      if (tr < rag_live) {
        mb = live_batch(tr);
        mt_in = tr - rag_pref[mb];
        active = true;
      } else {
        const int d = tr - rag_live;
        mb = dead_batch(d);
        mt_in = (rag_pref[mb + 1] - rag_pref[mb]) + d - (mb * rag_per_batch - rag_pref[mb]);
        active = false;
      }
      m0 = mt_in * p.m_tile_stride;
    } else {
      const int mt = (tile / p.n_tiles) * NPAIR + cp;
      active = mt < m_tiles_total;
      mb = mt / p.m_tiles_per_batch;
      m0 = (mt % p.m_tiles_per_batch) * p.m_tile_stride;
    }
    if (p.k_valid != nullptr) {
      const int per = (rag_live + p.splits - 1) / p.splits;
      kb_begin = split * per;
      kb_end = min(kb_begin + per, rag_live);
    } else {
      kb_begin = split * p.k_blocks_per_split;
      kb_end = min(kb_begin + p.k_blocks_per_split, p.k_blocks);
    }
  };

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------ TMA producer (both CTAs)
      // TMA coordinates are affine in (m0, mb, n_tile, k0, kbatch, kb, sub): the per-tile part is evaluated once per tile, the
      // per-K-block part is three multiply-adds per coordinate (this single thread must stay well ahead of the tensor core).
      int v[kCoordVars];
      int it = 0;
      for (int item = pair; item < n_items; item += n_pairs) {
        int mb, m0, n_tile, kb_begin, kb_end;
        bool active;
        decode(item, mb, m0, n_tile, kb_begin, kb_end, active);
        if (kb_begin >= kb_end) continue;
        if (NPAIR == 1 && !active) continue;  // dead tile of a ragged batch: nothing to load
        v[0] = 1; v[1] = m0 + 128 * static_cast<int>(rank); v[2] = mb; v[3] = n_tile;
        v[4] = 0; v[5] = 0; v[6] = 0; v[7] = 0;
        int ta[4], tb[4];
        // B rows of this CTA: its half of the 256-row tile (NPAIR = 1) or its quarter (NPAIR = 2, multicast to the other pair)
        const int b_sub = 128 * static_cast<int>(rank) + (NPAIR == 2 ? 64 * cp : 0);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          ta[i] = coord_dot(p.ca[i], v);
          tb[i] = coord_dot(p.cb[i], v) + p.cb[i][7] * b_sub;
        }
        const uint32_t tx_bytes = (active ? 2 * Cfg::kABytes : 0) + 2 * Cfg::kBBytes;
        const uint16_t mc_mask = static_cast<uint16_t>((1u << rank) | (1u << (rank + 2)));
        // K position: kb = kbatch * k_blocks_per_batch + kin; `klive` = K blocks of the current batch that are walked
        int kbatch = 0, kin = kb_begin, klive = p.k_blocks_per_batch;
        if (p.k_valid != nullptr) {
          kbatch = live_batch(kb_begin);
          kin = kb_begin - rag_pref[kbatch];
          klive = rag_pref[kbatch + 1] - rag_pref[kbatch];
        } else if (p.k_blocks_per_batch > 0) {
          kbatch = kb_begin / p.k_blocks_per_batch;
          kin = kb_begin - kbatch * p.k_blocks_per_batch;
        }
        auto advance = [&](int r) {
          ++kin;
          if (p.k_blocks_per_batch > 0 && kin == klive) {
            kin = 0;
            ++kbatch;
            if (p.k_valid != nullptr && r + 1 < kb_end) {
              while ((klive = rag_pref[kbatch + 1] - rag_pref[kbatch]) == 0) ++kbatch;  // utterances without a live block
            }
          }
        };
        for (int r = kb_begin; r < kb_end; ++r) {
          const int kb = p.k_valid != nullptr ? kbatch * p.k_blocks_per_batch + kin : r;
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          ++it;
          int ka[4], kbv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            ka[i] = ta[i] + p.ca[i][4] * (kin * 64) + p.ca[i][5] * kbatch + p.ca[i][6] * kb;
            kbv[i] = tb[i] + p.cb[i][4] * (kin * 64) + p.cb[i][5] * kbatch + p.cb[i][6] * kb;
          }
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + s * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          if (NPAIR == 1 && (p.debug & 4)) {  // diagnostics: barriers flow, no loads
            if (rank == 0) mbar_arrive(&full_bar[s]);
            advance(r);
            continue;
          }
          if (rank == 0) mbar_expect_tx(&full_bar[s], tx_bytes);  // bytes landing in both CTAs of this pair
          if (active) {
            if constexpr (!A_MN) {
              tma_load_4d_2cta(sa, &tmA, &full_bar[s], ka[0], ka[1], ka[2], ka[3]);
            } else {
#pragma unroll
              for (int i = 0; i < 2; ++i)
                tma_load_4d_2cta(sa + i * 8192, &tmA, &full_bar[s], ka[0] + p.ca[0][7] * 64 * i, ka[1] + p.ca[1][7] * 64 * i,
                                 ka[2] + p.ca[2][7] * 64 * i, ka[3] + p.ca[3][7] * 64 * i);
            }
          }
          if constexpr (NPAIR == 1) {
            if constexpr (!B_MN) {
              tma_load_4d_2cta(sb, &tmB, &full_bar[s], kbv[0], kbv[1], kbv[2], kbv[3]);
            } else {
#pragma unroll
              for (int i = 0; i < 2; ++i)
                tma_load_4d_2cta(sb + i * 8192, &tmB, &full_bar[s], kbv[0] + p.cb[0][7] * 64 * i, kbv[1] + p.cb[1][7] * 64 * i,
                                 kbv[2] + p.cb[2][7] * 64 * i, kbv[3] + p.cb[3][7] * 64 * i);
            }
          } else {
            // one 64-row quarter of the B tile (8 KB: box 64 x 64 in both majornesses), delivered to this CTA and to the CTA
            // of the same pair rank in the other pair; each destination signals its own pair leader's full barrier
            tma_load_4d_2cta_mc(sb + cp * 8192, &tmB, &full_bar[s], kbv[0], kbv[1], kbv[2], kbv[3], mc_mask);
          }
          advance(r);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      // ------------------------------------------------------------ MMA issuer (leader CTA only)
      constexpr uint32_t idesc = make_idesc_bf16(256, 256, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int it = 0, local = 0;  // `local` counts the non-empty items this pair has processed (accumulator stage = local & 1)
      for (int item = pair; item < n_items; item += n_pairs) {
        int mb, m0, n_tile, kb_begin, kb_end;
        bool active;
        decode(item, mb, m0, n_tile, kb_begin, kb_end, active);
        if (kb_begin >= kb_end) continue;
        if (NPAIR == 1 && !active) continue;  // dead tile of a ragged batch: the producer loaded nothing
        if (!active) {  // nothing to compute: just hand the stages (which received the multicast B parts) back
          for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
            mbar_wait(&full_bar[it % kStages], (it / kStages) & 1);
            umma_commit_2cta(&empty_bar[it % kStages], kAllCtas);
          }
          continue;
        }
        const int a = local & 1;
        mbar_wait(&tmem_empty_bar[a], ((local >> 1) & 1) ^ 1);  // both CTAs' epilogues drained this accumulator
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + a * 256;
        bool first = true;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          ++it;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * Cfg::kStageBytes);
          const uint32_t sb = sa + Cfg::kABytes;
          if (!(p.debug & 2)) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t da = A_MN ? make_smem_desc_sw128(sa + k * 2048, 8192, 1024)
                                       : make_smem_desc_sw128(sa + k * 32, 16, 1024);
              const uint64_t db = B_MN ? make_smem_desc_sw128(sb + k * 2048, 8192, 1024)
                                       : make_smem_desc_sw128(sb + k * 32, 16, 1024);
              umma_bf16_2cta(tmem_acc, da, db, idesc, (first && k == 0) ? 0u : 1u);
            }
          }
          first = false;
          umma_commit_2cta(&empty_bar[s], kAllCtas);  // frees the stage in every CTA that writes into it
        }
        umma_commit_2cta(&tmem_full_bar[a], static_cast<uint16_t>(0x3u << (2 * cp)));
        ++local;
      }
    }
  } else {
    // -------------------------------------------------------------- epilogue (both CTAs, 128 rows each)
    // 8 epilogue warps: two per TMEM lane quadrant (a warp may only touch lanes 32*(warp%4)..+31); the pair of warps splits
    // the 256 accumulator columns in halves, each half is processed as two 64-column chunks.  All global traffic of the
    // epilogue goes through the warp's swizzled staging buffers so that it is line-coalesced (see stg_off).
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int ew = warp - 2;
    const uint32_t stg0 = smem_u32(smem + kStages * Cfg::kStageBytes + ew * Cfg::kStageBufs * Cfg::kStageBufBytes);
    float* bias_s = reinterpret_cast<float*>(smem + kStages * Cfg::kStageBytes + Cfg::kStagingBytes) + ew * 128;
    const int n_in = (KIND == EK_F32) ? 0 : p.n_in;

    // which 32 x 128 block of which tile this warp owns (independent of the ragged validity)
    auto tile_of = [&](int item) {
      int mb, m0, n_tile, kb_begin, kb_end;
      bool active;
      decode(item, mb, m0, n_tile, kb_begin, kb_end, active);
      EpiTile et;
      et.mb = mb;
      const int m_valid = min(p.m_tile_valid, p.m_rows - m0) - 128 * static_cast<int>(rank) - q * 32;
      et.rows_valid = max(0, min(32, m_valid));
      et.row0 = static_cast<long long>(m0) + 128 * static_cast<int>(rank) + q * 32;
      const int col_base = n_tile * p.n_out_stride;
      const int n_valid = min(p.n_tile_valid, p.n_total - col_base) - half * 128;
      et.cols_valid = max(0, min(128, n_valid));
      et.col0 = col_base + half * 128;
      return et;
    };
    // dead tile of a ragged batch: its output rows (and the saved pre-activation / derivative) become zeros -- padded frames must
    // stay FINITE (a NaN in a padded key row would survive the -inf mask as NaN * 0 in the probabilities x V product)
    auto zero_tile = [&](int item) {
      if constexpr (KIND != EK_F32) {
        const EpiTile zt = tile_of(item);
        const int chunk = lane & 7;
#pragma unroll 1
        for (int t2 = 0; t2 < 2; ++t2) {
          const EpiTensor& t = (t2 == 0) ? p.out : p.out2;
          if (t.p == nullptr) continue;
#pragma unroll 1
          for (int cc = 0; cc < 2; ++cc) {
            const int colc = cc * 64 + chunk * 8;
            if (colc >= zt.cols_valid) continue;
            __nv_bfloat16* base = static_cast<__nv_bfloat16*>(t.p) + zt.mb * t.bs + zt.row0 * t.ld + zt.col0 + colc;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int row = i * 4 + (lane >> 3);
              if (row < zt.rows_valid) *reinterpret_cast<uint4*>(base + row * t.ld) = make_uint4(0u, 0u, 0u, 0u);
            }
          }
        }
      }
    };
    // next non-empty work item of this pair at or after `item` (n_items if none); dead tiles met on the way are zero-filled
    auto next_item = [&](int item) {
      for (; item < n_items; item += n_pairs) {
        int mb, m0, n_tile, kb_begin, kb_end;
        bool active;
        decode(item, mb, m0, n_tile, kb_begin, kb_end, active);
        if (kb_begin < kb_end && active) break;
        if (NPAIR == 1 && p.m_valid != nullptr && kb_begin < kb_end && !active) zero_tile(item);
      }
      return item;
    };
    // Input staging protocol.  `early` (exactly one input, no pre-activation output): the input of sequence element
    // (tile, chunk) #seq lives in buffer seq & 1 and the output is staged through the same buffer once the input is consumed, so
    // the prefetch of #seq+1 into the other buffer is issued at the start of #seq and is fully hidden.  Otherwise: input 0 ->
    // buffer 0, input 1 -> buffer 1, output through buffer 0, pre-activation (EK_GELU, <= 1 input) through buffer 1, and the
    // next prefetch is issued after the output has been flushed.
    const bool early = (KIND != EK_GELU) && (n_in == 1);
    auto prefetch = [&](const EpiTile& t, int cc, int seq) {
      if (early) {
        stage_input_async(p.in[0], t, cc, stg0 + (seq & 1) * Cfg::kStageBufBytes, lane);
      } else {
        stage_input_async(p.in[0], t, cc, stg0, lane);
        if (n_in >= 2) stage_input_async(p.in[1], t, cc, stg0 + Cfg::kStageBufBytes, lane);
      }
      cp_async_commit();
    };

    int item = next_item(pair);
    int local = 0;
    EpiTile et = tile_of(item < n_items ? item : 0);
    if (n_in >= 1 && item < n_items) prefetch(et, 0, 0);
    while (item < n_items) {
      const int item_next = next_item(item + n_pairs);
      const EpiTile et_next = tile_of(item_next < n_items ? item_next : 0);
      const int a = local & 1;
      if (KIND != EK_F32 && p.bias != nullptr) {  // this warp's 128 bias values -> shared memory (one coalesced load)
        const int c = lane * 4;
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < et.cols_valid) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + et.col0 + c));
        __syncwarp();
        *reinterpret_cast<float4*>(bias_s + c) = b4;
        __syncwarp();
      }
      mbar_wait(&tmem_full_bar[a], (local >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        const int seq = 2 * local + cc;
        const bool chunk_live = (cc * 64 < et.cols_valid) && !(p.debug & 1);  // warp-uniform
        // ---- accumulators of this lane's row: 64 fp32 columns
        float acc[64];
        {
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + a * 256 + half * 128 + cc * 64;
          tmem_ld_32x32b_x32(taddr, reinterpret_cast<uint32_t*>(acc));
          tmem_ld_32x32b_x32(taddr + 32, reinterpret_cast<uint32_t*>(acc) + 32);
        }
        const bool have_next = (cc == 0) || (item_next < n_items);   // is there a sequence element #seq+1?
        const EpiTile et_pf = (cc == 0) ? et : et_next;
        if (early) {  // the next sequence element's input goes into the other buffer right away
          if (lane == 0) bulk_wait_read0();  // ... once the TMA store that still reads that buffer has drained it
          __syncwarp();
          if (have_next) prefetch(et_pf, cc ^ 1, seq + 1);
          else cp_async_commit();
        }
        tmem_ld_wait();
        if (cc == 1) {  // all TMEM reads of this tile are done: hand the accumulator back to the MMA warp early
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(&tmem_empty_bar[a], crank & ~1u);  // this pair's leader
        }
        if constexpr (KIND == EK_F32) {
          // ---- fp32 (+)= acc, two 32-column passes through the staging buffer
          if (chunk_live) {
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                const float* a4 = acc + h2 * 32 + g * 4;
                uint4 w;
                w.x = __float_as_uint(a4[0]); w.y = __float_as_uint(a4[1]);
                w.z = __float_as_uint(a4[2]); w.w = __float_as_uint(a4[3]);
                sts128(stg0 + stg_off(lane, g), w);
              }
              __syncwarp();
              const int chunk = lane & 7;
              const int colc = cc * 64 + h2 * 32 + chunk * 4;
              float* base = static_cast<float*>(p.out.p) + et.mb * p.out.bs + et.row0 * p.out.ld + et.col0 + colc;
              const bool col_ok = colc < et.cols_valid;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int row = i * 4 + (lane >> 3);
                const uint4 w = lds128(stg0 + stg_off(row, chunk));
                if (col_ok && row < et.rows_valid) {
                  float* o = base + row * p.out.ld;
                  // fire-and-forget vector reduction for the single-writer case too: a read-modify-write here is 32 DEPENDENT
                  // global round trips per warp and tile (the compiler cannot move a row's load above the previous row's
                  // store), which left the weight-gradient GEMMs at half the speed of the forward ones
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o), "f"(__uint_as_float(w.x)),
                               "f"(__uint_as_float(w.y)), "f"(__uint_as_float(w.z)), "f"(__uint_as_float(w.w))
                               : "memory");
                }
              }
              __syncwarp();
            }
          }
        } else {
          const uint32_t buf_a = stg0 + (early ? (seq & 1) * Cfg::kStageBufBytes : 0);
          const uint32_t buf_b = stg0 + Cfg::kStageBufBytes;
          if (chunk_live) {
            if (p.bias != nullptr) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float4 b4 = *reinterpret_cast<const float4*>(bias_s + cc * 64 + j * 4);
                acc[4 * j] += b4.x; acc[4 * j + 1] += b4.y; acc[4 * j + 2] += b4.z; acc[4 * j + 3] += b4.w;
              }
            }
            if constexpr (KIND == EK_GELU) {
              const bool store_grad = (p.flags & EPI_GELU_STORE_GRAD) != 0;
              if (p.out2.p != nullptr) {  // saved for the backward pass (buffer 1 is free: n_in <= 1): the pre-activation, or
                if (lane == 0) bulk_wait_read0();  // directly gelu'(pre-activation) so that the backward epilogue is one multiply
                __syncwarp();
                if (store_grad) {
#pragma unroll
                  for (int g = 0; g < 8; ++g) {
                    float g8[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[g * 8 + j] = gelu_with_grad_f(acc[g * 8 + j], g8[j]);
                    sts128(buf_b + stg_off(lane, g), pack_bf16x8(g8));
                  }
                } else {
                  stage_row_bf16(acc, buf_b, lane);
                }
                fence_proxy_async_smem();  // generic-proxy writes -> visible to the TMA engine
                __syncwarp();
                if (lane == 0) {
                  tma_store_4d(&tmO2, buf_b, et.col0 + cc * 64, static_cast<int>(et.row0), et.mb, 0);
                  bulk_commit();
                }
              }
              if (!(store_grad && p.out2.p != nullptr)) {
#pragma unroll
                for (int j = 0; j < 64; ++j) acc[j] = gelu_f(acc[j]);
              }
            }
          }
          // ---- inputs (prefetched with cp.async): wait, make them visible to the whole warp, consume row-wise
          if (n_in >= 1) {
            if (early) cp_async_wait<1>(); else cp_async_wait<0>();
            __syncwarp();
            if constexpr (KIND == EK_DGELU) {
              if (chunk_live) consume_row<true>(acc, buf_a, lane, (p.flags & EPI_AUX_IS_GRAD) != 0);
            }
#pragma unroll 1
            for (int k = (KIND == EK_DGELU) ? 1 : 0; k < n_in; ++k) {
              if (k >= 2) {  // rare third input: synchronous round through buffer 1
                __syncwarp();
                stage_input_async(p.in[2], et, cc, buf_b, lane);
                cp_async_commit();
                cp_async_wait<0>();
                __syncwarp();
              }
              if (chunk_live) consume_row<false>(acc, k == 0 ? buf_a : buf_b, lane);
            }
            __syncwarp();  // every lane is done reading the input buffers: they may be overwritten
          }
          if (chunk_live) {
            // the staged 32 x 64 tile is exactly one SWIZZLE_128B TMA box: one elected lane stores it (rows / columns beyond
            // the tensor are clipped by the tensor map), the warp moves on while the TMA engine drains the buffer
            if (lane == 0) bulk_wait_read0();
            __syncwarp();
            stage_row_bf16(acc, buf_a, lane);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_4d(&tmO, buf_a, et.col0 + cc * 64, static_cast<int>(et.row0), et.mb, 0);
              bulk_commit();
            }
            if (p.flags & EPI_COLSUM) {
              // column sums of the stored (bf16-rounded) values from the staged tile: lane l owns columns 2l, 2l+1 of the chunk
              float s0 = 0.f, s1 = 0.f;
              const int chunk = lane >> 2, within = (lane & 3) * 4;
#pragma unroll 4
              for (int row = 0; row < et.rows_valid; ++row) {
                uint32_t w;
                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w) : "r"(buf_a + stg_off(row, chunk) + within) : "memory");
                const float2 f = unpack_bf16x2(w);
                s0 += f.x;
                s1 += f.y;
              }
              const int c = cc * 64 + 2 * lane;
              if (c < et.cols_valid) {
                atomicAdd(p.colsum + et.col0 + c, s0);
                atomicAdd(p.colsum + et.col0 + c + 1, s1);
              }
            }
            __syncwarp();
          }
          if (n_in >= 1 && !early) {  // late mode: the buffers are free only once the output store has drained them
            if (lane == 0) bulk_wait_read0();
            __syncwarp();
            if (have_next) prefetch(et_pf, cc ^ 1, seq + 1);
            else cp_async_commit();
          }
        }
      }
      item = item_next;
      et = et_next;
      ++local;
    }
    cp_async_wait<0>();
    if (lane == 0) bulk_wait0();  // every output box has been written before the CTA (and its shared memory) goes away
  }

  tc_fence_before();
  __syncwarp();
  cluster_sync_all();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc_2cta(tmem_base, 512);
  }
}

}  // namespace b200
