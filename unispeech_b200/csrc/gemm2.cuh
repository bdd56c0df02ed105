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

struct Gemm2Cfg {
  static constexpr int kStages = 6;
  static constexpr int kABytes = 128 * 128;  // 128 rows x 64 bf16 per CTA
  static constexpr int kBBytes = 128 * 128;  // this CTA's half (128 N rows) of the 256-wide B tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024;
  static constexpr int kThreads = 320;  // TMA warp + MMA warp + 8 epilogue warps
  static constexpr int kTileM = 256, kTileN = 256;
};

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(320, 1) gemm_bf16_pair_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                const __grid_constant__ CUtensorMap tmB,
                                                                const __grid_constant__ GemmParams p) {
  using Cfg = Gemm2Cfg;
  constexpr int kStages = Cfg::kStages;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();      // 0 = leader
  const int pair = blockIdx.x >> 1;
  const int n_pairs = gridDim.x >> 1;
  const int n_items = p.tiles_total * p.splits;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full_bar[kStages];
  __shared__ uint64_t empty_bar[kStages];
  __shared__ uint64_t tmem_full_bar[2];
  __shared__ uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_smem;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
#pragma unroll
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
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

  // item -> (split, m tile, n tile) and its K-block range; identical in every role of both CTAs
  auto decode = [&](int item, int& mb, int& m0, int& n_tile, int& kb_begin, int& kb_end) {
    const int split = item / p.tiles_total;
    const int tile = item % p.tiles_total;
    n_tile = tile % p.n_tiles;
    const int mt = tile / p.n_tiles;
    mb = mt / p.m_tiles_per_batch;
    m0 = (mt % p.m_tiles_per_batch) * p.m_tile_stride;
    kb_begin = split * p.k_blocks_per_split;
    kb_end = min(kb_begin + p.k_blocks_per_split, p.k_blocks);
  };

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------ TMA producer (both CTAs)
      int v[kCoordVars];
      v[0] = 1;
      int it = 0;
      for (int item = pair; item < n_items; item += n_pairs) {
        int mb, m0, n_tile, kb_begin, kb_end;
        decode(item, mb, m0, n_tile, kb_begin, kb_end);
        v[1] = m0 + 128 * static_cast<int>(rank);
        v[2] = mb;
        v[3] = n_tile;
        for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          if (p.k_blocks_per_batch > 0) {
            v[5] = kb / p.k_blocks_per_batch;
            v[4] = (kb % p.k_blocks_per_batch) * 64;
          } else {
            v[5] = 0;
            v[4] = kb * 64;
          }
          v[6] = kb;
          uint8_t* sa = smem + s * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * Cfg::kStageBytes);  // bytes of both CTAs
          if constexpr (!A_MN) {
            v[7] = 0;
            tma_load_4d_2cta(sa, &tmA, &full_bar[s], coord_dot(p.ca[0], v), coord_dot(p.ca[1], v), coord_dot(p.ca[2], v),
                             coord_dot(p.ca[3], v));
          } else {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              v[7] = 64 * i;
              tma_load_4d_2cta(sa + i * 8192, &tmA, &full_bar[s], coord_dot(p.ca[0], v), coord_dot(p.ca[1], v),
                               coord_dot(p.ca[2], v), coord_dot(p.ca[3], v));
            }
          }
          if constexpr (!B_MN) {
            v[7] = 128 * static_cast<int>(rank);
            tma_load_4d_2cta(sb, &tmB, &full_bar[s], coord_dot(p.cb[0], v), coord_dot(p.cb[1], v), coord_dot(p.cb[2], v),
                             coord_dot(p.cb[3], v));
          } else {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              v[7] = 128 * static_cast<int>(rank) + 64 * i;
              tma_load_4d_2cta(sb + i * 8192, &tmB, &full_bar[s], coord_dot(p.cb[0], v), coord_dot(p.cb[1], v),
                               coord_dot(p.cb[2], v), coord_dot(p.cb[3], v));
            }
          }
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
        decode(item, mb, m0, n_tile, kb_begin, kb_end);
        if (kb_begin >= kb_end) continue;
        const int a = local & 1;
        mbar_wait(&tmem_empty_bar[a], ((local >> 1) & 1) ^ 1);  // both CTAs' epilogues drained this accumulator
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + a * 256;
        bool first = true;
        for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * Cfg::kStageBytes);
          const uint32_t sb = sa + Cfg::kABytes;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = A_MN ? make_smem_desc_sw128(sa + k * 2048, 8192, 1024)
                                     : make_smem_desc_sw128(sa + k * 32, 16, 1024);
            const uint64_t db = B_MN ? make_smem_desc_sw128(sb + k * 2048, 8192, 1024)
                                     : make_smem_desc_sw128(sb + k * 32, 16, 1024);
            umma_bf16_2cta(tmem_acc, da, db, idesc, (first && k == 0) ? 0u : 1u);
          }
          first = false;
          umma_commit_2cta(&empty_bar[s], 0x3);  // frees the stage in both CTAs
        }
        umma_commit_2cta(&tmem_full_bar[a], 0x3);
        ++local;
      }
    }
  } else {
    // -------------------------------------------------------------- epilogue (both CTAs, 128 rows each)
    // 8 epilogue warps: two per TMEM lane quadrant (a warp may only touch lanes 32*(warp%4)..+31); the pair splits the 256
    // accumulator columns in halves.  Two warps per scheduler roughly double the issue rate of this issue-bound phase.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    int local = 0;
    for (int item = pair; item < n_items; item += n_pairs) {
      int mb, m0, n_tile, kb_begin, kb_end;
      decode(item, mb, m0, n_tile, kb_begin, kb_end);
      if (kb_begin >= kb_end) continue;
      const int a = local & 1;
      const int m0c = m0 + 128 * static_cast<int>(rank);
      const int m_valid = min(128, min(p.m_tile_valid, p.m_rows - m0) - 128 * static_cast<int>(rank));
      const bool row_ok = r < m_valid;
      const int col_base = n_tile * p.n_out_stride;
      const int n_valid = min(p.n_tile_valid, p.n_total - col_base);
      const long long row = static_cast<long long>(m0c) + r;
      const EpiRow erow = make_epi_row(p, mb, row);
      mbar_wait(&tmem_full_bar[a], (local >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = half * 128; c0 < half * 128 + 128; c0 += 32) {
        if (c0 >= n_valid) break;
        epilogue_chunk32(p, erow, tmem_base + (static_cast<uint32_t>(q * 32) << 16) + a * 256 + c0, c0, n_valid, col_base,
                         row_ok, lane);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(&tmem_empty_bar[a], 0);  // tell the leader's MMA warp this accumulator is free
      ++local;
    }
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
