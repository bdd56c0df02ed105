// Generic bf16 tcgen05 GEMM for sm_100a:  D[M,N] (+)= A[M,K] * B[N,K]^T, fp32 accumulation in TMEM.
//
// One CTA computes one 128 x BLOCK_N output tile (optionally one K-split of it).  Warp roles:
//   warp 0 : TMA producer (one elected lane)   global -> 128B-swizzled shared memory ring
//   warp 1 : TMEM allocator + MMA issuer (one elected lane, tcgen05.mma cta_group::1 kind::f16)
//   warps 2-5 : epilogue (tcgen05.ld 32x32b, one accumulator row per thread) with a fused
//               bias / GELU / GELU' / residual / fp32-atomic (split-K) tail.
// Operands may be K-major (reduction dim contiguous) or MN-major (reduction dim strided; used by the
// weight-gradient GEMMs), selected per operand.  All the "view" tricks of the WavLM path (strided
// Conv1d as an overlapping-row view, grouped pos_conv taps, per-batch tiles) are expressed on the host
// as <=4-D TMA tensor maps plus a small integer matrix that maps tile indices to TMA coordinates.
#pragma once
#include "ptx.cuh"

namespace b200 {

struct EpiTensor {
  void* p;
  long long bs;  // batch stride (elements)
  long long ld;  // row stride (elements)
};

enum : int {
  EPI_GELU = 1,      // out = gelu(acc + bias); if out2.p: out2 = acc + bias (pre-activation)
  EPI_DGELU = 2,     // acc *= gelu'(aux)
  EPI_OUT_F32 = 4,   // out is fp32
  EPI_ATOMIC = 8,    // fp32 atomicAdd into out (split-K / accumulation)
  EPI_COLSUM = 16,   // atomically accumulate column sums of the final value into colsum[col] (bias grads)
  EPI_ACCUM = 32,    // fp32 out += value, non-atomic (single writer per element: split-K disabled)
  EPI_GELU_STORE_GRAD = 64,  // with EPI_GELU: out2 receives gelu'(acc + bias) instead of the pre-activation
  EPI_AUX_IS_GRAD = 128,     // with EPI_DGELU: aux already holds gelu'(.) (written by an EPI_GELU_STORE_GRAD forward): acc *= aux
};

// variables the coordinate matrices multiply: {1, m0, mb, n_tile, k0, kbatch, kb, sub}
constexpr int kCoordVars = 8;

// ragged batches: at most this many batches per launch take the live-unit schedule (the per-batch prefix lives in shared memory)
constexpr int kMaxRagBatches = 255;

struct GemmParams {
  int m_rows;            // valid rows per batch
  int m_tiles_per_batch; // ceil over m_tile_stride
  int m_tile_stride;     // rows between consecutive M tiles (128 normally)
  int m_tile_valid;      // max valid rows per tile (128 normally; Cg for grouped wgrad)
  int n_total;           // valid output columns
  int n_out_stride;      // output column stride per N tile
  int n_tile_valid;      // max valid columns per tile
  int k_blocks;          // total number of 64-wide K blocks
  int k_blocks_per_batch;  // >0: K iterates (batch, row-block); 0: plain
  int k_blocks_per_split;
  int ca[4][kCoordVars];
  int cb[4][kCoordVars];
  int flags;
  const float* bias;     // [n] fp32 or null
  float* colsum;         // [n] fp32 or null (EPI_COLSUM)
  EpiTensor out, out2, aux, res1, res2;
  // persistent CTA-pair kernel only (gemm2.cuh): work items = (split, m_tile, n_tile), n fastest
  int n_tiles, tiles_total, splits;
  // bf16 epilogue inputs in application order, compacted on the host: in[0] is the GELU' argument when EPI_DGELU is set,
  // the others are added (res1, res2)
  EpiTensor in[3];
  int n_in;
  int debug;  // diagnostics only (B200S_GEMM_DEBUG): 1 = no epilogue global traffic, 2 = no MMAs, 4 = no TMA loads
  // ragged batches (CTA-pair kernel only; null = every row counts).  m_valid[b]: rows of batch b that hold real frames -- an M
  // tile that starts at or beyond it is not computed, its output rows are written as zeros.  k_valid[b] (weight gradients, K
  // iterates over (batch, row block)): row blocks that start at or beyond it are not loaded / multiplied (their gradient rows
  // are zero by construction: nothing downstream of a padded frame reaches the loss).
  const int* m_valid;
  const int* k_valid;
  // weight gradients (fp32 kind, one pair per cluster): the (tile, K block) space is cut into one contiguous range per CTA pair
  // (stream-K) instead of whole (split, tile) items; `splits` then holds the largest number of tiles a range can touch
  int stream_k;
};

template <int BLOCK_N>
struct GemmCfg {
  static constexpr int kStages = (BLOCK_N <= 64) ? 4 : 3;
  static constexpr int kABytes = 128 * 128;          // 128 rows x 64 bf16
  static constexpr int kBBytes = BLOCK_N * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024;
  static constexpr int kThreads = 192;
};

__device__ __forceinline__ int coord_dot(const int* row, const int* v) {
  int s = 0;
#pragma unroll
  for (int i = 0; i < kCoordVars; ++i) s += row[i] * v[i];
  return s;
}

// Per-thread row pointers of the epilogue tensors (computed once per tile: the epilogue is issue-bound -- one warp per
// scheduler -- so every instruction removed from the per-element path is wall-clock time).
struct EpiRow {
  __nv_bfloat16* out_bf;
  float* out_f32;
  __nv_bfloat16* out2;
  const __nv_bfloat16* aux;
  const __nv_bfloat16* res1;
  const __nv_bfloat16* res2;
};
__device__ __forceinline__ EpiRow make_epi_row(const GemmParams& p, int mb, long long row) {
  EpiRow e;
  e.out_bf = static_cast<__nv_bfloat16*>(p.out.p) + mb * p.out.bs + row * p.out.ld;
  e.out_f32 = static_cast<float*>(p.out.p) + mb * p.out.bs + row * p.out.ld;
  e.out2 = p.out2.p ? static_cast<__nv_bfloat16*>(p.out2.p) + mb * p.out2.bs + row * p.out2.ld : nullptr;
  e.aux = p.aux.p ? static_cast<const __nv_bfloat16*>(p.aux.p) + mb * p.aux.bs + row * p.aux.ld : nullptr;
  e.res1 = p.res1.p ? static_cast<const __nv_bfloat16*>(p.res1.p) + mb * p.res1.bs + row * p.res1.ld : nullptr;
  e.res2 = p.res2.p ? static_cast<const __nv_bfloat16*>(p.res2.p) + mb * p.res2.bs + row * p.res2.ld : nullptr;
  return e;
}
__device__ __forceinline__ void add_bf16x8(float* a8, const __nv_bfloat16* src) {
  const uint4 w = *reinterpret_cast<const uint4*>(src);
  const uint32_t wu[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = unpack_bf16x2(wu[j]);
    a8[2 * j] += f.x;
    a8[2 * j + 1] += f.y;
  }
}
__device__ __forceinline__ uint4 pack_bf16x8(const float* a8) {
  uint4 w;
  w.x = pack_bf16x2(a8[0], a8[1]); w.y = pack_bf16x2(a8[2], a8[3]);
  w.z = pack_bf16x2(a8[4], a8[5]); w.w = pack_bf16x2(a8[6], a8[7]);
  return w;
}

// One 32-column chunk of the fused epilogue for this thread's accumulator row (taddr = TMEM address of the chunk).
__device__ __forceinline__ void epilogue_chunk32(const GemmParams& p, const EpiRow& e, uint32_t taddr, int c0, int n_valid,
                                                 int col_base, bool row_ok, int lane) {
  const int flags = p.flags;
  uint32_t acc_u[32];
  tmem_ld_32x32b_x32(taddr, acc_u);
  tmem_ld_wait();
  float* acc = reinterpret_cast<float*>(acc_u);
  const bool full = (c0 + 32 <= n_valid);  // warp-uniform
  if (p.bias != nullptr) {
    const float* bp = p.bias + col_base + c0;
    if (full) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(bp) + j);
        acc[4 * j] += b4.x; acc[4 * j + 1] += b4.y; acc[4 * j + 2] += b4.z; acc[4 * j + 3] += b4.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (c0 + j < n_valid) acc[j] += __ldg(bp + j);
    }
  }
  if (row_ok) {
    const int col0 = col_base + c0;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      if (!full && (c0 + g * 8 >= n_valid)) break;
      const int col = col0 + g * 8;
      float* a8 = acc + g * 8;
      if (flags & EPI_GELU) {
        if (flags & EPI_GELU_STORE_GRAD) {
          float g8[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) a8[j] = gelu_with_grad_f(a8[j], g8[j]);
          if (e.out2 != nullptr) *reinterpret_cast<uint4*>(e.out2 + col) = pack_bf16x8(g8);
        } else {
          if (e.out2 != nullptr) *reinterpret_cast<uint4*>(e.out2 + col) = pack_bf16x8(a8);
#pragma unroll
          for (int j = 0; j < 8; ++j) a8[j] = gelu_f(a8[j]);
        }
      }
      if (flags & EPI_DGELU) {
        const uint4 w = *reinterpret_cast<const uint4*>(e.aux + col);
        const uint32_t wu[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = unpack_bf16x2(wu[j]);
          a8[2 * j] *= (flags & EPI_AUX_IS_GRAD) ? f.x : gelu_grad_f(f.x);
          a8[2 * j + 1] *= (flags & EPI_AUX_IS_GRAD) ? f.y : gelu_grad_f(f.y);
        }
      }
      if (e.res1 != nullptr) add_bf16x8(a8, e.res1 + col);
      if (e.res2 != nullptr) add_bf16x8(a8, e.res2 + col);
      if (flags & EPI_OUT_F32) {
        float* o = e.out_f32 + col;
        if (flags & EPI_ATOMIC) {
          // 128-bit vector reductions (REDG.E.ADD.F32x4): the split-K epilogue is bound by LSU issue, not bytes
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o), "f"(a8[0]), "f"(a8[1]), "f"(a8[2]), "f"(a8[3])
                       : "memory");
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + 4), "f"(a8[4]), "f"(a8[5]), "f"(a8[6]),
                       "f"(a8[7])
                       : "memory");
        } else if (flags & EPI_ACCUM) {
          float4 o0 = *reinterpret_cast<float4*>(o), o1 = *reinterpret_cast<float4*>(o + 4);
          o0.x += a8[0]; o0.y += a8[1]; o0.z += a8[2]; o0.w += a8[3];
          o1.x += a8[4]; o1.y += a8[5]; o1.z += a8[6]; o1.w += a8[7];
          *reinterpret_cast<float4*>(o) = o0;
          *reinterpret_cast<float4*>(o + 4) = o1;
        } else {
          *reinterpret_cast<float4*>(o) = make_float4(a8[0], a8[1], a8[2], a8[3]);
          *reinterpret_cast<float4*>(o + 4) = make_float4(a8[4], a8[5], a8[6], a8[7]);
        }
      } else {
        *reinterpret_cast<uint4*>(e.out_bf + col) = pack_bf16x8(a8);
      }
    }
  }
  if (flags & EPI_COLSUM) {
    // column sums of the stored values over this warp's 32 rows: transpose-reduce, then one atomic per column
    float cv[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float t = row_ok ? acc[j] : 0.0f;
      if (!(flags & EPI_OUT_F32)) t = __bfloat162float(__float2bfloat16_rn(t));
      cv[j] = t;
    }
    const float csum = warp_colsum32(cv, lane);
    if ((c0 + lane) < n_valid) atomicAdd(p.colsum + col_base + c0 + lane, csum);
  }
}

template <int BLOCK_N, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(192) gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA,
                                                        const __grid_constant__ CUtensorMap tmB,
                                                        const __grid_constant__ GemmParams p) {
  pdl_launch_dependents();
  using Cfg = GemmCfg<BLOCK_N>;
  constexpr int kStages = Cfg::kStages;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tile = blockIdx.x;
  const int mb = blockIdx.y / p.m_tiles_per_batch;
  const int m0 = (blockIdx.y % p.m_tiles_per_batch) * p.m_tile_stride;
  const int kb_begin = blockIdx.z * p.k_blocks_per_split;
  const int kb_end = min(kb_begin + p.k_blocks_per_split, p.k_blocks);
  if (kb_begin >= kb_end) return;  // uniform per CTA

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // 1024-aligned, still a __shared__ pointer (LDS/STS, not generic)
  __shared__ uint64_t full_bar[kStages];
  __shared__ uint64_t empty_bar[kStages];
  __shared__ uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_smem;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
#pragma unroll
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_smem, BLOCK_N);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_wait();  // prologue overlapped the previous kernel's tail; global memory is touched only from here on

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------ TMA producer
      int v[kCoordVars];
      v[0] = 1; v[1] = m0; v[2] = mb; v[3] = n_tile;
      int it = 0;
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
        mbar_expect_tx(&full_bar[s], Cfg::kStageBytes);
        if constexpr (!A_MN) {
          v[7] = 0;
          tma_load_4d(sa, &tmA, &full_bar[s], coord_dot(p.ca[0], v), coord_dot(p.ca[1], v), coord_dot(p.ca[2], v),
                      coord_dot(p.ca[3], v));
        } else {
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            v[7] = 64 * i;
            tma_load_4d(sa + i * 8192, &tmA, &full_bar[s], coord_dot(p.ca[0], v), coord_dot(p.ca[1], v),
                        coord_dot(p.ca[2], v), coord_dot(p.ca[3], v));
          }
        }
        if constexpr (!B_MN) {
          v[7] = 0;
          tma_load_4d(sb, &tmB, &full_bar[s], coord_dot(p.cb[0], v), coord_dot(p.cb[1], v), coord_dot(p.cb[2], v),
                      coord_dot(p.cb[3], v));
        } else {
#pragma unroll
          for (int i = 0; i < BLOCK_N / 64; ++i) {
            v[7] = 64 * i;
            tma_load_4d(sb + i * 8192, &tmB, &full_bar[s], coord_dot(p.cb[0], v), coord_dot(p.cb[1], v),
                        coord_dot(p.cb[2], v), coord_dot(p.cb[3], v));
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------------------------------------ MMA issuer
      constexpr uint32_t idesc = make_idesc_bf16(128, BLOCK_N, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int it = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * Cfg::kStageBytes);
        const uint32_t sb = sa + Cfg::kABytes;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // K-major: advance 16 elements (32 B) inside the 128B swizzle row; SBO = 8 rows * 128 B.
          // MN-major: advance 16 K-rows (2048 B); LBO = stride between 64-wide MN atoms (8192 B), SBO = 1024 B.
          const uint64_t da = A_MN ? make_smem_desc_sw128(sa + k * 2048, 8192, 1024)
                                   : make_smem_desc_sw128(sa + k * 32, 16, 1024);
          const uint64_t db = B_MN ? make_smem_desc_sw128(sb + k * 2048, 8192, 1024)
                                   : make_smem_desc_sw128(sb + k * 32, 16, 1024);
          umma_bf16(tmem_base, da, db, idesc, (it > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);  // frees the smem slot once these MMAs retire
      }
      umma_commit(&tmem_full_bar);
    }
  } else {
    // -------------------------------------------------------------- epilogue
    const int q = warp & 3;             // TMEM lane quadrant this warp may access
    const int r = q * 32 + lane;        // row inside the tile
    const int m_valid = min(p.m_tile_valid, p.m_rows - m0);
    const bool row_ok = r < m_valid;
    const int col_base = n_tile * p.n_out_stride;
    const int n_valid = min(p.n_tile_valid, p.n_total - col_base);
    const long long row = static_cast<long long>(m0) + r;

    const EpiRow erow = make_epi_row(p, mb, row);
    mbar_wait(&tmem_full_bar, 0);
    tc_fence_after();

#pragma unroll 1
    for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
      if (c0 >= n_valid) break;  // warp-uniform
      epilogue_chunk32(p, erow, tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c0, c0, n_valid, col_base, row_ok, lane);
    }
  }

  // teardown: everyone done with TMEM before the allocating warp frees it
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, BLOCK_N);
  }
}

// ---------------------------------------------------------------------------------------------- pos_conv, windowed
// Grouped Conv1d(k = taps, padding = taps/2) as an implicit GEMM whose A operand is loaded ONCE per tile: the 128 frames of a
// tile need input rows m0 .. m0+127+taps-1 of the zero-padded activation, and tap j multiplies rows m0+j .. m0+j+127 -- the same
// shared-memory tile, shifted by j rows.  One 256-row TMA box brings the window in; every tap is four tcgen05.mma whose A
// descriptor starts j rows (j * 128 bytes) into the tile (the swizzle follows the absolute address, see the MMA loop).  Only the per-tap weight tile (8 KB) streams through the TMA ring, so the kernel
// reads ~1/3 of the bytes of the box-per-tap formulation (which was bound by the L2 -> SM fabric).
// grid (groups, m_tiles * batches); block 192 (TMA warp, MMA warp, 4 epilogue warps); fused tail = epilogue_chunk32.
struct PosconvCfg {
  static constexpr int kStages = 6;
  static constexpr int kABytes = 256 * 128;  // 256 rows x 64 bf16
  static constexpr int kBBytes = 64 * 128;   // 64 output channels x 64 input channels of one tap
  static constexpr int kSmemBytes = kABytes + kStages * kBBytes + 1024;
  static constexpr int kThreads = 192;
};

__global__ void __launch_bounds__(192) posconv_window_kernel(const __grid_constant__ CUtensorMap tmA,
                                                             const __grid_constant__ CUtensorMap tmB,
                                                             const __grid_constant__ GemmParams p, int taps, int cg) {
  pdl_launch_dependents();
  using Cfg = PosconvCfg;
  constexpr int kStages = Cfg::kStages;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int g = blockIdx.x;
  const int mb = blockIdx.y / p.m_tiles_per_batch;
  const int m0 = (blockIdx.y % p.m_tiles_per_batch) * 128;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + Cfg::kABytes;
  __shared__ uint64_t a_full, full_bar[kStages], empty_bar[kStages], tmem_full_bar;
  __shared__ uint32_t tmem_base_smem;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    mbar_init(&a_full, 1);
#pragma unroll
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_smem, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(&a_full, Cfg::kABytes);
      tma_load_4d(sA, &tmA, &a_full, g * cg, m0, mb, 0);  // rows m0 .. m0+255 of this group's channels (zero-filled past the end)
      for (int j = 0; j < taps; ++j) {
        const int s = j % kStages;
        mbar_wait(&empty_bar[s], ((j / kStages) & 1) ^ 1);
        mbar_expect_tx(&full_bar[s], Cfg::kBBytes);
        tma_load_4d(sB + s * Cfg::kBBytes, &tmB, &full_bar[s], j * 64, g * 64, 0, 0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
      mbar_wait(&a_full, 0);
      for (int j = 0; j < taps; ++j) {
        const int s = j % kStages;
        mbar_wait(&full_bar[s], (j / kStages) & 1);
        tc_fence_after();
        const uint32_t sa = smem_u32(sA) + j * 128;  // the window shifted by j rows
        const uint32_t sb = smem_u32(sB + s * Cfg::kBBytes);
        // base offset 0: measured on B200, the tensor core applies the 128-byte swizzle to the ABSOLUTE shared-memory address
        // bits (as TMA does when it writes the tile), so a row-shifted start needs no correction (B200S_GEMM_DEBUG=8 sets the
        // "(start >> 7) & 7" value the descriptor format documents for unaligned starts: it produces wrong results here)
        const uint32_t bo = (p.debug & 8) ? static_cast<uint32_t>(j & 7) : 0u;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_base, make_smem_desc_sw128_bo(sa + k * 32, 16, 1024, bo), make_smem_desc_sw128(sb + k * 32, 16, 1024),
                    idesc, (j > 0 || k > 0) ? 1u : 0u);
        umma_commit(&empty_bar[s]);
      }
      umma_commit(&tmem_full_bar);
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int m_valid = min(128, p.m_rows - m0);
    const bool row_ok = r < m_valid;
    const int col_base = g * cg;
    const EpiRow erow = make_epi_row(p, mb, static_cast<long long>(m0) + r);
    mbar_wait(&tmem_full_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c0 = 0; c0 < 64; c0 += 32) {
      if (c0 >= cg) break;
      epilogue_chunk32(p, erow, tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c0, c0, cg, col_base, row_ok, lane);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 64);
  }
}

}  // namespace b200
