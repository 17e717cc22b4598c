// Persistent, warp-specialised bf16 GEMM for sm_100a:  C[M,N] = A[M,K] * W[N,K]^T (+ epilogue)
//
//   TMA (cp.async.bulk.tensor, SWIZZLE_128B) -> 4-stage smem ring -> tcgen05.mma (kind::f16,
//   fp32 accumulators in TMEM, double-buffered 2 x 256 columns) -> tcgen05.ld epilogue.
//
// It replaces every nn.Linear / F.linear on RUArt's hot path:
//   BERT  query/key/value (Models/Bert/modeling.py:225-227, fused into one N=2304 GEMM),
//         BertSelfOutput.dense (:261), BertIntermediate.dense + gelu (:287-288),
//         BertOutput.dense (:300)
//   SDNet AttentionScore.linear (+ReLU, *diagonal) (Models/Layers.py:226-231),
//         the LSTM input projections W_ih x + b_ih + b_hh of every nn.LSTM (Layers.py:137,166).
//
// "Split" operands give fp32-grade accuracy on bf16 tensor cores: an fp32 matrix X is stored as
// up to three bf16 parts X = X0 + X1 + X2 laid side by side ([rows, nparts*Kp]); the K loop then
// walks a list of (part_a, part_b) terms (1 term: plain bf16; 3 terms: ~2^-16; 6 terms: ~2^-24).
#include <cuda.h>

#include <cstdlib>
#include "common.cuh"
#include "ruart_b200.h"

namespace {

using namespace ruart;

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int MAX_BN = 256;
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BM * BK * 2;      // 16 KB
constexpr int B_STAGE_BYTES = MAX_BN * BK * 2;  // 32 KB
constexpr int C_STAGE_BYTES = BM * 64 * 2;      // 16 KB: 128 rows x 64 bf16, TMA-store staging
constexpr int OFF_B = STAGES * A_STAGE_BYTES;
constexpr int OFF_C = OFF_B + STAGES * B_STAGE_BYTES;
constexpr int OFF_VEC = OFF_C + 2 * C_STAGE_BYTES;
constexpr int OFF_BAR = OFF_VEC + 2 * MAX_BN * 4;  // bias / scale slice, double-buffered per tile
constexpr int BAR_BYTES = 128;
constexpr int GEMM_SMEM_BYTES = OFF_BAR + BAR_BYTES;
constexpr int GEMM_THREADS = 384;  // TMA, MMA, TMEM-alloc, spare + 2 x 4 epilogue warps
constexpr int TMEM_COLS = 512;
static_assert(GEMM_SMEM_BYTES <= 232448, "shared memory budget");
// CTA-pair (cta_group::2) variant: pair tile 256 x 256; each CTA stages its 128 rows of A and HALF
// of the W tile, so a stage is 32 KB and the ring is 6 deep in the same shared memory
constexpr int STAGES2 = 6;
constexpr int B2_ROWS = MAX_BN / 2;
constexpr int B2_STAGE_BYTES = B2_ROWS * BK * 2;  // 16 KB
constexpr int OFF2_B = STAGES2 * A_STAGE_BYTES;
constexpr int OFF2_C = OFF2_B + STAGES2 * B2_STAGE_BYTES;
constexpr int OFF2_VEC = OFF2_C + 2 * C_STAGE_BYTES;
constexpr int OFF2_BAR = OFF2_VEC + 2 * MAX_BN * 4;
constexpr int GEMM2_SMEM_BYTES = OFF2_BAR + 256;
static_assert(GEMM2_SMEM_BYTES <= 232448, "shared memory budget (2-CTA)");

// epilogue kinds as compiled (the ABI's RUART_EPI_BIAS_GELU maps to one of the two GELUs)
enum : int { K_NONE = 0, K_BIAS, K_GELU_FAST, K_GELU_EXACT, K_RELU_SCALE, K_BIAS_RELU, K_GELU_TANHFIT, K_NUM };

struct GemmParams {
  int M, N, Kp;
  int block_n;
  int n_terms;
  uint32_t term_a, term_b;  // 4 bits per term: which part of A / W
  const float* vec;         // bias [N] or scale ([N] or [1]); nullptr for K_NONE
  int vec_stride;           // 1, or 0 for a broadcast scalar
  float* out_f32;
  long long ldo_f32;
  __nv_bfloat16* out_bf16;
  long long ldo_bf16;
  int out_parts;              // 1: plain bf16; 2/3: hi|mid|lo split parts
  long long out_part_stride;  // elements between parts inside one row
  const __nv_bfloat16* residual;  // TMA_OUT only: added (fp32) before rounding, or nullptr
  long long ld_res;
  // folded LayerNorm (CTA-pair kernel, FOLD != 0; see the comment above gemm_bf16_2cta_kernel)
  const float2* in_stats;   // [M][8] partial (sum, sum of squares) of the rows whose LayerNorm is pending:
                            // FOLD 1: the rows of A; FOLD 2: the rows of the residual
  const float* vec2;        // FOLD 1: colsum[n] = sum_k W'[n,k]; FOLD 2: gamma[n] of the residual's LayerNorm
  float2* out_stats;        // FOLD 2: [M][8] partial sums of the rows written here (slot 2 * n_blk + group)
  float ln_inv_dim;         // 1 / (width of the normalised rows)
  float ln_eps;
};

template <int EPI>
__device__ __forceinline__ float epi_fn(float acc, float v) {
  if constexpr (EPI == K_NONE) return acc;
  if constexpr (EPI == K_BIAS) return acc + v;
  if constexpr (EPI == K_GELU_FAST) return gelu_erf_fast(acc + v);
  if constexpr (EPI == K_GELU_EXACT) return gelu_erf(acc + v);
  if constexpr (EPI == K_GELU_TANHFIT) return gelu_erf_tanhfit(acc + v);
  if constexpr (EPI == K_RELU_SCALE) return fmaxf(acc, 0.0f) * v;
  if constexpr (EPI == K_BIAS_RELU) return fmaxf(acc + v, 0.0f);
  return acc;
}

// activation only (the folded-LayerNorm epilogue has already formed the pre-activation)
template <int EPI>
__device__ __forceinline__ float epi_act(float x) {
  if constexpr (EPI == K_GELU_FAST) return gelu_erf_fast(x);
  if constexpr (EPI == K_GELU_EXACT) return gelu_erf(x);
  if constexpr (EPI == K_GELU_TANHFIT) return gelu_erf_tanhfit(x);
  if constexpr (EPI == K_BIAS_RELU) return fmaxf(x, 0.0f);
  return x;
}

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
// epilogue group barrier (group 0 -> id 1, group 1 -> id 2), 128 threads each
__device__ __forceinline__ void grp_bar_sync(int grp) {
  asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory");
}
__device__ __forceinline__ void epi_all_bar_sync() { asm volatile("bar.sync 3, 256;" ::: "memory"); }
__device__ __forceinline__ void tma_store_2d(const void* tmap, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   tmap),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() {
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// EPI: epilogue kind.  TMA_OUT: plain bf16 output written with swizzled smem staging + TMA
// stores (the BERT path); otherwise generic direct stores (fp32 and/or split bf16 outputs).
template <int EPI, bool TMA_OUT, bool HAS_RES>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a,
                         const __grid_constant__ CUtensorMap tmap_b,
                         const __grid_constant__ CUtensorMap tmap_c,
                         const __grid_constant__ CUtensorMap tmap_r, const GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + OFF_B;
  uint8_t* smem_c = smem + OFF_C;
  float* s_vec = reinterpret_cast<float*>(smem + OFF_VEC);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full_bar = bars + 2 * STAGES;
  uint64_t* tmem_empty_bar = bars + 2 * STAGES + 2;
  uint64_t* res_bar = bars + 2 * STAGES + 4;  // one per epilogue group: residual tile landed
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 6);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (TMA_OUT) tma_prefetch_desc(&tmap_c);
    if (TMA_OUT && HAS_RES) tma_prefetch_desc(&tmap_r);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], TMA_OUT ? 256 : 128);
      mbar_init(&res_bar[s], 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc<TMEM_COLS>(tmem_ptr_smem);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int m_tiles = (p.M + BM - 1) / BM;
  const int n_tiles = (p.N + p.block_n - 1) / p.block_n;
  const int total_tiles = m_tiles * n_tiles;
  const int k_blocks = p.Kp / BK;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes = A_STAGE_BYTES + p.block_n * BK * 2;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m_blk = tile / n_tiles;
        const int n_blk = tile - m_blk * n_tiles;
        for (int t = 0; t < p.n_terms; ++t) {
          const int pa = (p.term_a >> (4 * t)) & 0xF;
          const int pb = (p.term_b >> (4 * t)) & 0xF;
          for (int kb = 0; kb < k_blocks; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
            tma_load_2d(&tmap_a, &full_bar[stage], smem_a + stage * A_STAGE_BYTES,
                        pa * p.Kp + kb * BK, m_blk * BM);
            tma_load_2d(&tmap_b, &full_bar[stage], smem_b + stage * B_STAGE_BYTES,
                        pb * p.Kp + kb * BK, n_blk * p.block_n);
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    {  // whole warp walks the loop, one elected lane issues (see gemm_bf16_2cta_kernel)
      const uint32_t idesc = make_idesc_bf16_f32(BM, p.block_n);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const int iters = p.n_terms * k_blocks;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * MAX_BN;
        for (int it = 0; it < iters; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t da = make_sw128_kmajor_desc(smem_u32(smem_a + stage * A_STAGE_BYTES));
          const uint64_t db = make_sw128_kmajor_desc(smem_u32(smem_b + stage * B_STAGE_BYTES));
          if (elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              // advance 16 bf16 = 32 B along K inside the 128 B swizzle row: +2 in (addr >> 4)
              umma_bf16_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, (it > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(&empty_bar[stage]);  // smem slot is free once these MMAs retire
          }
          __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (elect_one_sync()) umma_commit(&tmem_full_bar[acc]);  // accumulator complete -> epilogue
        __syncwarp();
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else if (warp >= 4 && (TMA_OUT || warp < 8)) {
    // ------------------------------------------------------------------ epilogue
    // TMA_OUT: two groups of 4 warps (warps 4-7 / 8-11, i.e. two warps per SM sub-partition so that
    // the MUFU/ALU latency of the GELU is hidden); group g converts the 64-column chunks
    // g, g+2, ... of the tile, each through its own staging buffer and TMA store.
    // Direct-store path: group 0 only.
    const int grp = (warp - 4) >> 2;
    const int ew = warp & 3;  // TMEM lane quadrant this warp may touch
    const int et = threadIdx.x - 128 - grp * 128;  // thread index inside the group
    const int r_local = ew * 32 + lane;
    const bool issuer = (et == 0);
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t res_phase = 0;
    uint32_t tile_ctr = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tile_ctr) {
      const int m_blk = tile / n_tiles;
      const int n_blk = tile - m_blk * n_tiles;
      const int tile_col0 = n_blk * p.block_n;
      float* vec = s_vec + (TMA_OUT ? (tile_ctr & 1u) * MAX_BN : 0);
      // stage this tile's bias / scale slice
      if (EPI != K_NONE) {
        const int stride = TMA_OUT ? 256 : 128;
        for (int i = et + grp * 128; i < p.block_n; i += stride) {
          const int col = tile_col0 + i;
          vec[i] = (col < p.N) ? __ldg(p.vec + static_cast<long long>(col) * p.vec_stride) : 0.0f;
        }
      }
      if (TMA_OUT) epi_all_bar_sync(); else epi_bar_sync();
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) +
                                static_cast<uint32_t>(acc * MAX_BN);
      if constexpr (TMA_OUT) {
        uint8_t* cbuf = smem_c + grp * C_STAGE_BYTES;
        constexpr bool has_res = HAS_RES;
        if (grp * 64 >= p.block_n) {  // this group has no chunk in such a narrow tile
          tc_fence_before();
          mbar_arrive(&tmem_empty_bar[acc]);
        }
        for (int c0 = grp * 64; c0 < p.block_n; c0 += 128) {
          // residual tile (BertSelfOutput / BertOutput: dense(x) + input_tensor, modeling.py:263,302):
          // TMA-loaded into the group's staging buffer while the accumulators are converted
          if (has_res && issuer) {
            bulk_wait_read<0>();  // the group's previous store has finished reading cbuf
            mbar_arrive_expect_tx(&res_bar[grp], C_STAGE_BYTES);
            tma_load_2d(&tmap_r, &res_bar[grp], cbuf, tile_col0 + c0, m_blk * BM);
          }
          uint32_t v0[32], v1[32];
          tmem_ld_32x32b_x32(tmem_acc + c0, v0);
          tmem_ld_32x32b_x32(tmem_acc + c0 + 32, v1);
          tmem_ld_wait();
          if (c0 + 128 >= p.block_n) {  // the group's last TMEM read of this tile
            tc_fence_before();
            mbar_arrive(&tmem_empty_bar[acc]);
          }
          uint8_t* row_ptr = cbuf + r_local * 128;
          if constexpr (!HAS_RES) {
            // convert and pack BEFORE waiting for the staging buffer: the math overlaps the wait
            uint4 pk4[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              uint32_t pk[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const int j = c * 8 + q * 2;  // column inside the 64-wide chunk
                const float x0 = __uint_as_float(j < 32 ? v0[j] : v1[j - 32]);
                const float x1 = __uint_as_float(j < 32 ? v0[j + 1] : v1[j - 31]);
                const float2 bv = *reinterpret_cast<const float2*>(vec + c0 + j);
                pk[q] = pack_bf16x2(epi_fn<EPI>(x0, bv.x), epi_fn<EPI>(x1, bv.y));
              }
              pk4[c] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
            if (issuer) bulk_wait_read<0>();  // the group's previous store has finished reading cbuf
            grp_bar_sync(grp);
#pragma unroll
            for (int c = 0; c < 8; ++c)
              *reinterpret_cast<uint4*>(row_ptr + ((c ^ (r_local & 7)) << 4)) = pk4[c];
          } else {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const float2 b0 = *reinterpret_cast<const float2*>(vec + c0 + j);
              const float2 b1 = *reinterpret_cast<const float2*>(vec + c0 + 32 + j);
              v0[j] = __float_as_uint(epi_fn<EPI>(__uint_as_float(v0[j]), b0.x));
              v0[j + 1] = __float_as_uint(epi_fn<EPI>(__uint_as_float(v0[j + 1]), b0.y));
              v1[j] = __float_as_uint(epi_fn<EPI>(__uint_as_float(v1[j]), b1.x));
              v1[j + 1] = __float_as_uint(epi_fn<EPI>(__uint_as_float(v1[j + 1]), b1.y));
            }
            mbar_wait(&res_bar[grp], res_phase);  // residual tile has landed in cbuf
            res_phase ^= 1u;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              uint4* slot = reinterpret_cast<uint4*>(row_ptr + ((c ^ (r_local & 7)) << 4));
              const uint4 r4 = *slot;
              const uint32_t rw[4] = {r4.x, r4.y, r4.z, r4.w};
              uint32_t pk[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const int j = c * 8 + q * 2;
                const float x0 = __uint_as_float(j < 32 ? v0[j] : v1[j - 32]);
                const float x1 = __uint_as_float(j < 32 ? v0[j + 1] : v1[j - 31]);
                pk[q] = pack_bf16x2(x0 + bf16_lo(rw[q]), x1 + bf16_hi(rw[q]));
              }
              *slot = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
          }
          fence_proxy_async();
          grp_bar_sync(grp);
          if (issuer) {
            tma_store_2d(&tmap_c, cbuf, tile_col0 + c0, m_blk * BM);
            bulk_commit();
          }
        }
      } else {
        const int row = m_blk * BM + r_local;
        const bool row_ok = row < p.M;
        for (int c0 = 0; c0 < p.block_n; c0 += 32) {
          uint32_t v[32];
          __syncwarp();
          tmem_ld_32x32b_x32(tmem_acc + c0, v);
          tmem_ld_wait();
          const int col0 = tile_col0 + c0;
          if (col0 >= p.N) break;  // warp-uniform: the rest of this tile is past the last column
          const bool full_chunk = (col0 + 32 <= p.N);
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = epi_fn<EPI>(__uint_as_float(v[j]), vec[c0 + j]);
          if (row_ok && p.out_f32 != nullptr) {
            float* dst = p.out_f32 + static_cast<long long>(row) * p.ldo_f32 + col0;
            if (full_chunk && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(dst + j) =
                    make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) dst[j] = f[j];
            }
          }
          if (row_ok && p.out_bf16 != nullptr) {
#pragma unroll 1
            for (int part = 0; part < p.out_parts; ++part) {
              __nv_bfloat16* dst = p.out_bf16 + static_cast<long long>(row) * p.ldo_bf16 +
                                   static_cast<long long>(part) * p.out_part_stride + col0;
              uint32_t pk[16];
#pragma unroll
              for (int j = 0; j < 32; j += 2) {
                const __nv_bfloat16 h0 = __float2bfloat16_rn(f[j]);
                const __nv_bfloat16 h1 = __float2bfloat16_rn(f[j + 1]);
                pk[j / 2] = static_cast<uint32_t>(__bfloat16_as_ushort(h0)) |
                            (static_cast<uint32_t>(__bfloat16_as_ushort(h1)) << 16);
                f[j] -= __bfloat162float(h0);  // residual feeds the next (finer) part
                f[j + 1] -= __bfloat162float(h1);
              }
              if (full_chunk && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                  *reinterpret_cast<uint4*>(dst + 2 * j) =
                      make_uint4(pk[j], pk[j + 1], pk[j + 2], pk[j + 3]);
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (col0 + j < p.N)
                    dst[j] = __ushort_as_bfloat16(
                        static_cast<unsigned short>((pk[j / 2] >> (16 * (j & 1))) & 0xFFFFu));
              }
            }
          }
        }
        tc_fence_before();
        mbar_arrive(&tmem_empty_bar[acc]);
        epi_bar_sync();  // s_vec may be overwritten for the next tile only after all reads
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
    if (TMA_OUT && issuer) bulk_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2) for the plain-bf16-output GEMMs of the BERT encoder.
// Two CTAs of a cluster (one TPC) own a 256 x 256 tile: CTA r stages rows [128 r, 128 r + 128) of the
// A tile and rows [128 r, 128 r + 128) of the W tile; the leader's single MMA thread issues
// M = 256, N = 256 instructions that read both CTAs' shared memory and write each CTA's 128 x 256
// half of the accumulator into that CTA's own TMEM.  Per CTA a stage shrinks from 48 KB to 32 KB, so
// the TMA ring is 6 deep instead of 4 (the 1-CTA kernel loses 6-9 % with 3 stages: it is bound by
// load latency) and the W traffic from L2 halves.  Barrier protocol:
//   full[s]   (leader's)  1 arrival: the leader's arrive.expect_tx(64 KB); both CTAs' TMA loads
//                         complete_tx on it
//   empty[s]  (each CTA)  tcgen05.commit multicast to both CTAs when the MMAs of stage s retire
//   tmem_full (each CTA)  commit multicast after the last k-block of a tile
//   tmem_empty (leader's) 4 arrivals: one elected thread per epilogue group of each CTA (the peer's
//                         arrive remotely)
// The epilogue is the TMA-store epilogue of the kernel above, per CTA on its own 128 rows.
//
// Folded LayerNorm (FOLD != 0, bf16 BERT).  BertLayerNorm is affine per row, y = (v - mu) r gamma + beta, so it
// never has to be materialised between two GEMMs:
//   FOLD 2 (BertSelfOutput / BertOutput dense, modeling.py:260-264,299-303): the epilogue adds the residual
//           and writes the PRE-LayerNorm rows v (bf16) plus, per row and per 128-column half-tile, the partial
//           sums (sum v, sum v^2) taken from the fp32 values; the residual itself is still "pending" its own
//           LayerNorm, which is applied on the fly from ITS partial sums: res = (v~ r - mu r) gamma + (beta + bias)
//   FOLD 1 (the GEMM that consumes LayerNorm(v): query/key/value, BertIntermediate): A = v~ as stored,
//           W' = W * gamma (folded on the host), and the epilogue finishes the normalisation,
//           out = r acc - r mu colsum(W') + (W beta + bias)
// 24 LayerNorm passes per forward (1.6 ms, 8.4 GB of DRAM traffic at cfg-3) disappear.  Shared memory is full
// (6 stages + 2 staging buffers), so the two per-column vectors of a FOLD tile are single-buffered and one more
// 256-thread barrier per tile protects them.
template <int EPI, bool HAS_RES, int FOLD>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_2cta_kernel(const __grid_constant__ CUtensorMap tmap_a,
                      const __grid_constant__ CUtensorMap tmap_b,
                      const __grid_constant__ CUtensorMap tmap_c,
                      const __grid_constant__ CUtensorMap tmap_r, const GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + OFF2_B;
  uint8_t* smem_c = smem + OFF2_C;
  float* s_vec = reinterpret_cast<float*>(smem + OFF2_VEC);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF2_BAR);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES2;
  uint64_t* tmem_full_bar = bars + 2 * STAGES2;
  uint64_t* tmem_empty_bar = bars + 2 * STAGES2 + 2;
  uint64_t* res_bar = bars + 2 * STAGES2 + 4;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES2 + 6);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_c);
    if (HAS_RES) tma_prefetch_desc(&tmap_r);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES2; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 4);  // one elected thread per epilogue group per CTA
      mbar_init(&res_bar[s], 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc_2sm<TMEM_COLS>(tmem_ptr_smem);
  tc_fence_before();
  cluster_sync_all();  // barriers of BOTH CTAs are initialised before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int m_pairs = (p.M + 2 * BM - 1) / (2 * BM);
  const int n_tiles = p.N / MAX_BN;
  const int total_tiles = m_pairs * n_tiles;
  const int k_blocks = p.Kp / BK;
  const int tile0 = blockIdx.x >> 1;
  const int tile_step = gridDim.x >> 1;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile0; tile < total_tiles; tile += tile_step) {
        const int m_pair = tile / n_tiles;
        const int n_blk = tile - m_pair * n_tiles;
        const int row_a = (2 * m_pair + static_cast<int>(rank)) * BM;
        const int row_b = n_blk * MAX_BN + static_cast<int>(rank) * B2_ROWS;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint32_t full_leader;
          asm volatile("mapa.shared::cluster.u32 %0, %1, %2;"
                       : "=r"(full_leader)
                       : "r"(smem_u32(&full_bar[stage])), "r"(0));
          // the leader alone arrives (expecting the bytes of BOTH CTAs); the peer's loads only
          // complete_tx on the leader's barrier.  (A remote arrive.release.cluster per stage from
          // the peer's producer serialised it behind its own TMA loads: 0.73 us per stage.)
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * (A_STAGE_BYTES + B2_STAGE_BYTES));
          tma_load_2d_2sm(&tmap_a, full_leader, smem_a + stage * A_STAGE_BYTES, kb * BK, row_a);
          tma_load_2d_2sm(&tmap_b, full_leader, smem_b + stage * B2_STAGE_BYTES, kb * BK, row_b);
          if (++stage == STAGES2) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader only)
    // The WHOLE warp walks the loop (uniform control flow, every operand provably warp-uniform) and one elected lane
    // issues.  With a single-lane loop (`lane == 0`) the compiler cannot keep the descriptors in uniform registers
    // and wraps every tcgen05.mma in an ELECT / 5 x R2UR.BROADCAST / BRA.U.ANY waterfall: ~95 instructions per
    // k-block on a scheduler shared with two epilogue warps — ncu showed the issuing warp busy 87 % of the time and
    // the tensor pipe idling in proportion to the epilogue's instruction count ("MMA issue loop").
    if (leader) {
      const uint32_t idesc = make_idesc_bf16_f32(2 * BM, MAX_BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = tile0; tile < total_tiles; tile += tile_step) {
        mbar_wait_warp(&tmem_empty_bar[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * MAX_BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait_warp(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t da = make_sw128_kmajor_desc(smem_u32(smem_a + stage * A_STAGE_BYTES));
          const uint64_t db = make_sw128_kmajor_desc(smem_u32(smem_b + stage * B2_STAGE_BYTES));
          if (elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16_ss_2sm(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit_2sm(&empty_bar[stage], 0x3);  // frees the slot in BOTH CTAs
          }
          __syncwarp();
          if (++stage == STAGES2) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (elect_one_sync()) umma_commit_2sm(&tmem_full_bar[acc], 0x3);  // both CTAs' epilogues
        __syncwarp();
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (both CTAs)
    static_assert(FOLD == 0 || (FOLD == 1 && !HAS_RES) || (FOLD == 2 && HAS_RES && EPI == K_BIAS), "fold forms");
    const int grp = (warp - 4) >> 2;
    const int ew = warp & 3;
    const int et = threadIdx.x - 128 - grp * 128;
    const int r_local = ew * 32 + lane;
    const bool issuer = (et == 0);
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t res_phase = 0;
    uint32_t tile_ctr = 0;
    // FOLD: the tile's two per-column vectors (one element each per thread) and the row's partial sums are fetched
    // ONE TILE AHEAD into registers: at the top of a tile they would cost two exposed L2 round trips (ncu: 18 % of
    // the epilogue warps' time on the first use of the sums, profiles/r02c_ncu_full_summary.txt)
    float pf_v = 0.f, pf_v2 = 0.f;
    float4 pf_a = make_float4(0.f, 0.f, 0.f, 0.f), pf_b = pf_a, pf_c = pf_a;
    auto prefetch = [&](int tile) {
      if constexpr (FOLD != 0) {
        if (tile < total_tiles) {
          const int m_pair = tile / n_tiles;
          const int n_blk = tile - m_pair * n_tiles;
          const int col = n_blk * MAX_BN + et + grp * 128;
          pf_v = __ldg(p.vec + col);
          pf_v2 = __ldg(p.vec2 + col);
          const int row = (2 * m_pair + static_cast<int>(rank)) * BM + r_local;
          if (row < p.M) {
            const float4* sp = reinterpret_cast<const float4*>(p.in_stats + static_cast<long long>(row) * 8);
            pf_a = __ldg(sp);
            pf_b = __ldg(sp + 1);
            pf_c = __ldg(sp + 2);
          } else {  // harmless unit-variance stand-in for rows past M
            pf_a = make_float4(0.f, 1.0f / p.ln_inv_dim, 0.f, 0.f);
            pf_b = pf_c = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      }
    };
    prefetch(tile0);
    for (int tile = tile0; tile < total_tiles; tile += tile_step, ++tile_ctr) {
      const int m_pair = tile / n_tiles;
      const int n_blk = tile - m_pair * n_tiles;
      const int m_blk = 2 * m_pair + static_cast<int>(rank);
      const int tile_col0 = n_blk * MAX_BN;
      float* vec = s_vec + (FOLD ? 0u : (tile_ctr & 1u) * MAX_BN);
      float* vec2 = s_vec + MAX_BN;  // FOLD only
      // pending LayerNorm of this thread's row: r = rstd, nmr = -mean * rstd
      float ln_r = 1.0f, ln_nmr = 0.0f;
      if constexpr (FOLD != 0) {
        epi_all_bar_sync();  // single-buffered vectors: the previous tile's readers are done
        vec[et + grp * 128] = pf_v;
        vec2[et + grp * 128] = pf_v2;
        const float s1 = ((pf_a.x + pf_a.z) + (pf_b.x + pf_b.z)) + (pf_c.x + pf_c.z);
        const float s2 = ((pf_a.y + pf_a.w) + (pf_b.y + pf_b.w)) + (pf_c.y + pf_c.w);
        const float mu = s1 * p.ln_inv_dim;
        const float var = fmaxf(fmaf(-mu, mu, s2 * p.ln_inv_dim), 0.0f);
        ln_r = rsqrtf(var + p.ln_eps);
        ln_nmr = -mu * ln_r;
        prefetch(tile + tile_step);
      } else if (EPI != K_NONE) {
        for (int i = et + grp * 128; i < MAX_BN; i += 256)
          vec[i] = __ldg(p.vec + static_cast<long long>(tile_col0 + i) * p.vec_stride);
      }
      float st1 = 0.f, st2 = 0.f;  // FOLD 2: this thread's partial sums over its 128 columns
      epi_all_bar_sync();
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) +
                                static_cast<uint32_t>(acc * MAX_BN);
      uint8_t* cbuf = smem_c + grp * C_STAGE_BYTES;
      for (int c0 = grp * 64; c0 < MAX_BN; c0 += 128) {
        if (HAS_RES && issuer) {
          bulk_wait_read<0>();
          mbar_arrive_expect_tx(&res_bar[grp], C_STAGE_BYTES);
          tma_load_2d(&tmap_r, &res_bar[grp], cbuf, tile_col0 + c0, m_blk * BM);
        }
        uint32_t v0[32], v1[32];
        tmem_ld_32x32b_x32(tmem_acc + c0, v0);
        tmem_ld_32x32b_x32(tmem_acc + c0 + 32, v1);
        tmem_ld_wait();
        if (c0 + 128 >= MAX_BN) {  // the group's last TMEM read of this tile
          // one (remote, for the peer) arrive per group instead of 128: every thread's TMEM loads are
          // ordered before the group barrier, the elected thread then releases the accumulator
          tc_fence_before();
          grp_bar_sync(grp);
          if (issuer) {
            if (leader) mbar_arrive(&tmem_empty_bar[acc]);
            else mbar_arrive_cluster_relaxed(&tmem_empty_bar[acc], 0);
          }
        }
        uint8_t* row_ptr = cbuf + r_local * 128;
        if constexpr (!HAS_RES) {
          uint4 pk4[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            uint32_t pk[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int j = c * 8 + q * 2;
              float x0 = __uint_as_float(j < 32 ? v0[j] : v1[j - 32]);
              float x1 = __uint_as_float(j < 32 ? v0[j + 1] : v1[j - 31]);
              const float2 bv = *reinterpret_cast<const float2*>(vec + c0 + j);
              if constexpr (FOLD == 1) {
                // LayerNorm(v) W^T + b  =  r acc + (-mu r) colsum + (W beta + b)
                const float2 cs = *reinterpret_cast<const float2*>(vec2 + c0 + j);
                x0 = fmaf(ln_r, x0, fmaf(ln_nmr, cs.x, bv.x));
                x1 = fmaf(ln_r, x1, fmaf(ln_nmr, cs.y, bv.y));
                pk[q] = pack_bf16x2(epi_act<EPI>(x0), epi_act<EPI>(x1));
              } else {
                pk[q] = pack_bf16x2(epi_fn<EPI>(x0, bv.x), epi_fn<EPI>(x1, bv.y));
              }
            }
            pk4[c] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
          if (issuer) bulk_wait_read<0>();
          grp_bar_sync(grp);
#pragma unroll
          for (int c = 0; c < 8; ++c)
            *reinterpret_cast<uint4*>(row_ptr + ((c ^ (r_local & 7)) << 4)) = pk4[c];
        } else {
          if constexpr (FOLD == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const float2 b0 = *reinterpret_cast<const float2*>(vec + c0 + j);
              const float2 b1 = *reinterpret_cast<const float2*>(vec + c0 + 32 + j);
              v0[j] = __float_as_uint(epi_fn<EPI>(__uint_as_float(v0[j]), b0.x));
              v0[j + 1] = __float_as_uint(epi_fn<EPI>(__uint_as_float(v0[j + 1]), b0.y));
              v1[j] = __float_as_uint(epi_fn<EPI>(__uint_as_float(v1[j]), b1.x));
              v1[j + 1] = __float_as_uint(epi_fn<EPI>(__uint_as_float(v1[j + 1]), b1.y));
            }
          }
          mbar_wait(&res_bar[grp], res_phase);
          res_phase ^= 1u;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            uint4* slot = reinterpret_cast<uint4*>(row_ptr + ((c ^ (r_local & 7)) << 4));
            const uint4 r4 = *slot;
            const uint32_t rw[4] = {r4.x, r4.y, r4.z, r4.w};
            uint32_t pk[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int j = c * 8 + q * 2;
              float x0 = __uint_as_float(j < 32 ? v0[j] : v1[j - 32]);
              float x1 = __uint_as_float(j < 32 ? v0[j + 1] : v1[j - 31]);
              if constexpr (FOLD == 2) {
                // residual = LayerNorm(stored rows) = (v~ r - mu r) gamma + beta; vec holds beta + dense bias
                const float2 gm = *reinterpret_cast<const float2*>(vec2 + c0 + j);
                const float2 bt = *reinterpret_cast<const float2*>(vec + c0 + j);
                x0 += fmaf(fmaf(bf16_lo(rw[q]), ln_r, ln_nmr), gm.x, bt.x);
                x1 += fmaf(fmaf(bf16_hi(rw[q]), ln_r, ln_nmr), gm.y, bt.y);
                st1 += x0 + x1;
                st2 = fmaf(x0, x0, fmaf(x1, x1, st2));
              } else {
                x0 += bf16_lo(rw[q]);
                x1 += bf16_hi(rw[q]);
              }
              pk[q] = pack_bf16x2(x0, x1);
            }
            *slot = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
        fence_proxy_async();
        grp_bar_sync(grp);
        if (issuer) {
          tma_store_2d(&tmap_c, cbuf, tile_col0 + c0, m_blk * BM);
          bulk_commit();
        }
      }
      if constexpr (FOLD == 2) {
        const int row = m_blk * BM + r_local;
        if (row < p.M) p.out_stats[static_cast<long long>(row) * 8 + n_blk * 2 + grp] = make_float2(st1, st2);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
    if (issuer) bulk_wait_all();
  }

  tc_fence_before();
  cluster_sync_all();  // the peer's MMAs / multicast commits no longer touch this CTA
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm<TMEM_COLS>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// Fused query/key/value projection + self-attention (BertSelfAttention.forward, modeling.py:224-250) for the
// bf16 encoder: the [T, 2304] Q|K|V matrix is never written to memory.
//
// The weight rows are permuted head-major on the host ([Wq_h / 8 ; Wk_h ; Wv_h], 192 rows per head, the softmax
// scale folded into the query rows), so one 192-column accumulator tile holds one head's Q, K and V of the tile's
// rows.  Row tiles come from ruart_seq_tiles: consecutive WHOLE sequences, at most 128 token rows, so every
// sequence meets all of its keys inside one CTA's 128 TMEM lanes (RUArt's sequences are 3-10 wordpieces for
// items, <= 50 for questions).  16 warps:
//   0 TMA producer, 1 MMA issuer (leader CTA), 2 TMEM allocation — the CTA-pair pipeline of the kernel above with
//     N = 192 and a 4-stage ring;
//   4-7 "converters": read the accumulator (tcgen05.ld, one TMEM lane quadrant each), finish the pending LayerNorm
//     of the A rows (FOLD 1 above), round to bf16 and store Q, K, V as three XOR-swizzled [128 x 64] tiles in
//     one of TWO shared-memory buffers; they release the TMEM buffer as soon as it is read;
//   8-15 "attention": 16 query rows per warp: S = Q K^T and O = P V on mma.sync.m16n8k16 (the register layout of
//     the stand-alone kernels in bert_kernels.cu), masked by the rows' sequence bounds, online softmax over the
//     16-key blocks of the rows' key range only; O is staged through the warp's own Q rows and written as
//     128-byte rows of ctx[:, 64 h : 64 h + 64].
// Converters and attention warps hand the tile buffers over with bar.arrive / bar.sync pairs, so the attention of
// tile i runs under the conversion of tile i + 1 and the MMAs of tiles i + 1, i + 2.  (First version: one buffer,
// all eight epilogue warps doing both phases in turn — 0.53 ms per layer at cfg-3 against 0.29 ms for the GEMM
// alone, tools/bench_qkv_attn.py; the MMAs were not the cost: without them the attention phase took as long.)
constexpr int QA_THREADS = 512;
constexpr int QA_BN = 192;
constexpr int QA_B_ROWS = QA_BN / 2;
constexpr int QA_B_STAGE = QA_B_ROWS * BK * 2;  // 12 KB
constexpr int QA_STAGES = 4;
constexpr int QA_TILE = 128 * 128;              // one [128 x 64] bf16 tile
constexpr int QA_BUF = 3 * QA_TILE;             // Q | K | V
constexpr int QA_OFF_B = QA_STAGES * A_STAGE_BYTES;
constexpr int QA_OFF_T = QA_OFF_B + QA_STAGES * QA_B_STAGE;
constexpr int QA_OFF_VEC = QA_OFF_T + 2 * QA_BUF;
constexpr int QA_OFF_BAR = QA_OFF_VEC + 2 * QA_BN * 4;
constexpr int QA_SMEM_BYTES = QA_OFF_BAR + 256;
static_assert(QA_SMEM_BYTES <= 232448 && (QA_OFF_T % 1024) == 0 && (QA_B_STAGE % 1024) == 0, "qkv+attention smem");
// named barriers: 1 = the four converter warps; 2, 3 = tile buffer 0 / 1 written; 4, 5 = tile buffer 0 / 1 read
__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

struct QaParams {
  int M, Kp, n_heads;
  const int32_t* meta;      // ruart_seq_tiles: [0] = tiles, [1 ..] = first row of each tile, then M
  const int2* bounds;       // [M] (first row, end row) of each token's sequence
  const float2* in_stats;   // [M][8] partial sums of the A rows (pending LayerNorm)
  const float* vec;         // [n_heads * 192] W0 beta + b, permuted like the weight rows
  const float* vec2;        // [n_heads * 192] colsum(W)
  float ln_inv_dim, ln_eps;
  __nv_bfloat16* out;       // ctx [M, n_heads * 64]
  long long ldo;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(QA_THREADS, 1)
qkv_attn_2cta_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                     const QaParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + QA_OFF_B;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + QA_OFF_BAR);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + QA_STAGES;
  uint64_t* tmem_full_bar = bars + 2 * QA_STAGES;
  uint64_t* tmem_empty_bar = bars + 2 * QA_STAGES + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * QA_STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < QA_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 2);  // one elected converter thread per CTA
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc_2sm<TMEM_COLS>(tmem_ptr_smem);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int n_mtiles = __ldg(p.meta);
  const int m_pairs = (n_mtiles + 1) >> 1;
  const int total_tiles = m_pairs * p.n_heads;
  const int k_blocks = p.Kp / BK;
  const int tile0 = blockIdx.x >> 1;
  const int tile_step = gridDim.x >> 1;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile0; tile < total_tiles; tile += tile_step) {
        const int m_pair = tile / p.n_heads;
        const int h = tile - m_pair * p.n_heads;
        const int mt = 2 * m_pair + static_cast<int>(rank);
        const int row_a = (mt < n_mtiles) ? __ldg(p.meta + 1 + mt) : p.M;  // past the end: zero-filled tile
        const int row_b = h * QA_BN + static_cast<int>(rank) * QA_B_ROWS;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint32_t full_leader;
          asm volatile("mapa.shared::cluster.u32 %0, %1, %2;"
                       : "=r"(full_leader)
                       : "r"(smem_u32(&full_bar[stage])), "r"(0));
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * (A_STAGE_BYTES + QA_B_STAGE));
          tma_load_2d_2sm(&tmap_a, full_leader, smem_a + stage * A_STAGE_BYTES, kb * BK, row_a);
          tma_load_2d_2sm(&tmap_b, full_leader, smem_b + stage * QA_B_STAGE, kb * BK, row_b);
          if (++stage == QA_STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader only)
    // single-lane loop here: the whole-warp / elected-lane form that helps the GEMM kernels above measured SLOWER in
    // this kernel (0.400 vs 0.377 ms per layer, tools/bench_qkv_attn.py) — its schedulers are shared with 12 busy warps
    if (leader && lane == 0) {
      {
      const uint32_t idesc = make_idesc_bf16_f32(2 * BM, QA_BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = tile0; tile < total_tiles; tile += tile_step) {
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * MAX_BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t da = make_sw128_kmajor_desc(smem_u32(smem_a + stage * A_STAGE_BYTES));
          const uint64_t db = make_sw128_kmajor_desc(smem_u32(smem_b + stage * QA_B_STAGE));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16_ss_2sm(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit_2sm(&empty_bar[stage], 0x3);
          if (++stage == QA_STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit_2sm(&tmem_full_bar[acc], 0x3);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ------------------------------------------------------------------ converters (both CTAs)
    const int ew = warp & 3;               // TMEM lane quadrant
    const int ct = threadIdx.x - 128;      // 0..127 = the tile row this thread converts
    float* vec = reinterpret_cast<float*>(smem + QA_OFF_VEC);
    float* vec2 = vec + QA_BN;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t tile_ctr = 0;
    // the tile's vectors (192 + 192 floats over 128 threads: 3 elements each) and the row's partial sums, fetched one
    // tile ahead into registers (two dependent L2 round trips: tile table -> partial sums)
    float pf_v[3], pf_v2[3];
    float4 pf_a, pf_b, pf_c;
    auto prefetch = [&](int tile) {
      if (tile >= total_tiles) return;
      const int m_pair = tile / p.n_heads;
      const int h = tile - m_pair * p.n_heads;
      const int mt = 2 * m_pair + static_cast<int>(rank);
      const int row = ((mt < n_mtiles) ? __ldg(p.meta + 1 + mt) : p.M) + ct;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int i = ct + 128 * k;          // i < 192 for k = 0, and for k = 1 when ct < 64
        pf_v[k] = (i < QA_BN) ? __ldg(p.vec + h * QA_BN + i) : 0.f;
        pf_v2[k] = (i < QA_BN) ? __ldg(p.vec2 + h * QA_BN + i) : 0.f;
      }
      if (row < p.M) {
        const float4* sp = reinterpret_cast<const float4*>(p.in_stats + static_cast<long long>(row) * 8);
        pf_a = __ldg(sp);
        pf_b = __ldg(sp + 1);
        pf_c = __ldg(sp + 2);
      } else {
        pf_a = make_float4(0.f, 1.0f / p.ln_inv_dim, 0.f, 0.f);
        pf_b = pf_c = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    prefetch(tile0);
    for (int tile = tile0; tile < total_tiles; tile += tile_step, ++tile_ctr) {
      const int buf = tile_ctr & 1u;
      // the previous tile's conversion is over (group barrier at its end): the vectors may be replaced
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int i = ct + 128 * k;
        if (i < QA_BN) {
          vec[i] = pf_v[k];
          vec2[i] = pf_v2[k];
        }
      }
      float ln_r, ln_nmr;
      {
        const float s1 = ((pf_a.x + pf_a.z) + (pf_b.x + pf_b.z)) + (pf_c.x + pf_c.z);
        const float s2 = ((pf_a.y + pf_a.w) + (pf_b.y + pf_b.w)) + (pf_c.y + pf_c.w);
        const float mu = s1 * p.ln_inv_dim;
        const float var = fmaxf(fmaf(-mu, mu, s2 * p.ln_inv_dim), 0.0f);
        ln_r = rsqrtf(var + p.ln_eps);
        ln_nmr = -mu * ln_r;
      }
      prefetch(tile + tile_step);
      named_bar_sync(1, 128);
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      if (tile_ctr >= 2) named_bar_sync(4 + buf, 384);  // the attention warps have finished with this buffer
      uint8_t* sT = smem + QA_OFF_T + buf * QA_BUF;
      const uint32_t tmem_acc = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + static_cast<uint32_t>(acc * MAX_BN);
#pragma unroll 1
      for (int cc = 0; cc < 3; ++cc) {       // Q, K, V: 64 columns each
        uint32_t v0[32], v1[32];
        tmem_ld_32x32b_x32(tmem_acc + cc * 64, v0);
        tmem_ld_32x32b_x32(tmem_acc + cc * 64 + 32, v1);
        tmem_ld_wait();
        uint8_t* trow = sT + cc * QA_TILE;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t pk[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int j = c * 8 + q * 2;
            const float x0 = __uint_as_float(j < 32 ? v0[j] : v1[j - 32]);
            const float x1 = __uint_as_float(j < 32 ? v0[j + 1] : v1[j - 31]);
            const float2 bv = *reinterpret_cast<const float2*>(vec + cc * 64 + j);
            const float2 cs = *reinterpret_cast<const float2*>(vec2 + cc * 64 + j);
            pk[q] = pack_bf16x2(fmaf(ln_r, x0, fmaf(ln_nmr, cs.x, bv.x)), fmaf(ln_r, x1, fmaf(ln_nmr, cs.y, bv.y)));
          }
          *reinterpret_cast<uint4*>(trow + tile_off(ct, c)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
      tc_fence_before();
      named_bar_arrive(2 + buf, 384);        // Q, K, V of this tile are in buffer `buf`
      named_bar_sync(1, 128);                // every converter's TMEM reads are done
      if (ct == 0) {
        if (leader) mbar_arrive(&tmem_empty_bar[acc]);
        else mbar_arrive_cluster_relaxed(&tmem_empty_bar[acc], 0);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  } else if (warp >= 8) {
    // ------------------------------------------------------------------ attention (both CTAs)
    const int q0 = (warp - 8) * 16;        // this warp's query rows [q0, q0 + 16) of the tile
    const int g = lane >> 2, t = lane & 3;
    const int a_row = (lane & 7) + ((lane >> 3) & 1) * 8;
    const int a_chk = lane >> 4;
    const int b_row = lane & 7;
    const int b_chk = lane >> 3;
    uint32_t tile_ctr = 0;
    // (row0, n_valid, this thread's two sequence bounds) of a tile; fetched one tile AHEAD so that the two dependent
    // L2 round trips (tile table -> bounds) are off the critical path
    struct TileInfo { int row0, n_valid; int2 bA, bB; };
    auto fetch = [&](int tile) {
      TileInfo ti;
      const int m_pair = tile / p.n_heads;
      const int mt = 2 * m_pair + static_cast<int>(rank);
      const bool live = tile < total_tiles && mt < n_mtiles;
      ti.row0 = live ? __ldg(p.meta + 1 + mt) : p.M;
      ti.n_valid = live ? __ldg(p.meta + 2 + mt) - ti.row0 : 0;
      ti.bA = make_int2(q0 + g, q0 + g + 1);
      ti.bB = make_int2(q0 + g + 8, q0 + g + 9);
      if (ti.row0 + q0 + g < p.M) {
        const int2 gb = __ldg(p.bounds + ti.row0 + q0 + g);
        ti.bA = make_int2(max(gb.x - ti.row0, 0), min(gb.y - ti.row0, 128));
      }
      if (ti.row0 + q0 + g + 8 < p.M) {
        const int2 gb = __ldg(p.bounds + ti.row0 + q0 + g + 8);
        ti.bB = make_int2(max(gb.x - ti.row0, 0), min(gb.y - ti.row0, 128));
      }
      return ti;
    };
    TileInfo nxt = fetch(tile0);
    for (int tile = tile0; tile < total_tiles; tile += tile_step, ++tile_ctr) {
      const int m_pair = tile / p.n_heads;
      const int h = tile - m_pair * p.n_heads;
      const TileInfo cur = nxt;
      const int row0 = cur.row0, n_valid = cur.n_valid;
      const int2 bA = cur.bA, bB = cur.bB;
      const int buf = tile_ctr & 1u;
      nxt = fetch(tile + tile_step);
      named_bar_sync(2 + buf, 384);          // the converters have written buffer `buf`
      if (q0 < n_valid) {
        uint8_t* sQ = smem + QA_OFF_T + buf * QA_BUF;
        const uint32_t aQ = smem_u32(sQ), aK = aQ + QA_TILE, aV = aK + QA_TILE;
        uint32_t qa[4][4];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          ldsm_x4(aQ + tile_off(q0 + a_row, 2 * ks + a_chk), qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3]);
        // rows are in sequence order: the warp's key range runs from its first row's sequence start (lane 0 holds
        // row q0) to its last row's sequence end (lanes 28-31 hold row q0 + 15)
        const int klo = __shfl_sync(0xffffffffu, bA.x, 0), khi = __shfl_sync(0xffffffffu, bB.y, 28);
        float mA = -INFINITY, mB = -INFINITY, lA = 0.f, lB = 0.f;
        float o[8][4];
#pragma unroll
        for (int d = 0; d < 8; ++d) o[d][0] = o[d][1] = o[d][2] = o[d][3] = 0.f;
        // 16-key blocks start at the first key (not at a multiple of 16): ceil(range / 16) blocks; rows past the tile
        // (all masked) are clamped to its last row
        for (int kb = klo; kb < khi; kb += 16) {
          float s[2][4];
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
          for (int kh = 0; kh < 2; ++kh) {
            uint32_t kf[2][4];
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
              ldsm_x4(aK + tile_off(min(kb + nt * 8 + b_row, 127), 4 * kh + b_chk), kf[nt][0], kf[nt][1], kf[nt][2], kf[nt][3]);
#pragma unroll
            for (int k2 = 0; k2 < 2; ++k2)
#pragma unroll
              for (int nt = 0; nt < 2; ++nt)
                mma_bf16_16816(s[nt], qa[2 * kh + k2][0], qa[2 * kh + k2][1], qa[2 * kh + k2][2], qa[2 * kh + k2][3],
                               kf[nt][2 * k2], kf[nt][2 * k2 + 1]);
          }
          float bmA = -INFINITY, bmB = -INFINITY;
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) {
            const int k0 = kb + nt * 8 + 2 * t;
            s[nt][0] = (k0 >= bA.x && k0 < bA.y) ? s[nt][0] : -INFINITY;
            s[nt][1] = (k0 + 1 >= bA.x && k0 + 1 < bA.y) ? s[nt][1] : -INFINITY;
            s[nt][2] = (k0 >= bB.x && k0 < bB.y) ? s[nt][2] : -INFINITY;
            s[nt][3] = (k0 + 1 >= bB.x && k0 + 1 < bB.y) ? s[nt][3] : -INFINITY;
            bmA = fmaxf(bmA, fmaxf(s[nt][0], s[nt][1]));
            bmB = fmaxf(bmB, fmaxf(s[nt][2], s[nt][3]));
          }
          bmA = fmaxf(bmA, __shfl_xor_sync(0xffffffffu, bmA, 1));
          bmA = fmaxf(bmA, __shfl_xor_sync(0xffffffffu, bmA, 2));
          bmB = fmaxf(bmB, __shfl_xor_sync(0xffffffffu, bmB, 1));
          bmB = fmaxf(bmB, __shfl_xor_sync(0xffffffffu, bmB, 2));
          const float nmA = fmaxf(mA, bmA), nmB = fmaxf(mB, bmB);
          // rows whose sequence has not started yet in this key block keep (m, l, o) = (-inf, 0, 0)
          const float fA = (nmA == -INFINITY) ? 1.0f : __expf(mA - nmA);
          const float fB = (nmB == -INFINITY) ? 1.0f : __expf(mB - nmB);
          const float zA = (nmA == -INFINITY) ? 0.0f : nmA, zB = (nmB == -INFINITY) ? 0.0f : nmB;
          mA = nmA;
          mB = nmB;
          float psA = 0.f, psB = 0.f;
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) {
            s[nt][0] = __expf(s[nt][0] - zA);
            s[nt][1] = __expf(s[nt][1] - zA);
            s[nt][2] = __expf(s[nt][2] - zB);
            s[nt][3] = __expf(s[nt][3] - zB);
            psA += s[nt][0] + s[nt][1];
            psB += s[nt][2] + s[nt][3];
          }
          lA = fmaf(lA, fA, psA);
          lB = fmaf(lB, fB, psB);
#pragma unroll
          for (int d = 0; d < 8; ++d) {
            o[d][0] *= fA; o[d][1] *= fA;
            o[d][2] *= fB; o[d][3] *= fB;
          }
          const uint32_t p0 = pack_bf16x2(s[0][0], s[0][1]), p1 = pack_bf16x2(s[0][2], s[0][3]);
          const uint32_t p2 = pack_bf16x2(s[1][0], s[1][1]), p3 = pack_bf16x2(s[1][2], s[1][3]);
#pragma unroll
          for (int dp = 0; dp < 4; ++dp) {
            uint32_t b0, b1, b2, b3;
            ldsm_x4_t(aV + tile_off(min(kb + a_row, 127), 2 * dp + a_chk), b0, b1, b2, b3);
            mma_bf16_16816(o[2 * dp], p0, p1, p2, p3, b0, b1);
            mma_bf16_16816(o[2 * dp + 1], p0, p1, p2, p3, b2, b3);
          }
        }
        lA += __shfl_xor_sync(0xffffffffu, lA, 1);
        lA += __shfl_xor_sync(0xffffffffu, lA, 2);
        lB += __shfl_xor_sync(0xffffffffu, lB, 1);
        lB += __shfl_xor_sync(0xffffffffu, lB, 2);
        const float iA = lA > 0.f ? __fdividef(1.0f, lA) : 0.f, iB = lB > 0.f ? __fdividef(1.0f, lB) : 0.f;
        __syncwarp();  // this warp's Q fragments are in registers: its 16 Q rows become the O staging rows
        const int rA = q0 + g, rB = rA + 8;
#pragma unroll
        for (int d = 0; d < 8; ++d) {
          *reinterpret_cast<uint32_t*>(sQ + tile_off(rA, d) + 4 * t) = pack_bf16x2(o[d][0] * iA, o[d][1] * iA);
          *reinterpret_cast<uint32_t*>(sQ + tile_off(rB, d) + 4 * t) = pack_bf16x2(o[d][2] * iB, o[d][3] * iB);
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = q0 + (lane >> 3) + 4 * i;
          if (r < n_valid) {
            const uint4 o4 = *reinterpret_cast<const uint4*>(sQ + tile_off(r, lane & 7));
            *(reinterpret_cast<uint4*>(p.out + static_cast<long long>(row0 + r) * p.ldo + h * 64) + (lane & 7)) = o4;
          }
        }
      }
      named_bar_arrive(4 + buf, 384);        // this warp no longer touches buffer `buf`
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm<TMEM_COLS>(tmem_base);
  }
}

// --------------------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = (PFN_encodeTiled)ptr;
  }
  return fn;
}

// 2-D bf16 row-major [rows, cols] with leading dimension ld (elements); box = [box_rows, 64].
int make_tmap_bf16(CUtensorMap* tm, const void* base, long long rows, long long cols, long long ld,
                   int box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (enc == nullptr) {
    ruart_set_error("cuTensorMapEncodeTiled is not available (no CUDA driver?)");
    return RUART_ERR_NO_DRIVER;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ruart_set_error("cuTensorMapEncodeTiled failed: %d (base=%p rows=%lld cols=%lld ld=%lld box=%d)",
                    (int)r, base, rows, cols, ld, box_rows);
    return RUART_ERR_CUDA;
  }
  return RUART_OK;
}

int pick_block_n(int N, int gran) {
  // N tile: multiple of `gran` (32 for the direct-store epilogue, 64 for TMA stores), <= 256;
  // prefer the largest tile that wastes the least of the last column block.
  if (N >= 256) {
    int best = 256, best_waste = 1 << 30;
    for (int bn = 256; bn >= 128; bn -= gran) {
      const int tiles = (N + bn - 1) / bn;
      const int waste = tiles * bn - N;
      if (waste < best_waste) {
        best_waste = waste;
        best = bn;
      }
    }
    return best;
  }
  return ((N + gran - 1) / gran) * gran;
}

typedef void (*GemmKernel)(const CUtensorMap, const CUtensorMap, const CUtensorMap,
                           const CUtensorMap, const GemmParams);

template <int EPI>
GemmKernel pick_kernel(bool tma_out) {
  return tma_out ? gemm_bf16_tcgen05_kernel<EPI, true, false>
                 : gemm_bf16_tcgen05_kernel<EPI, false, false>;
}

GemmKernel kernel_for(int kind, bool tma_out, bool has_res) {
  if (has_res) return gemm_bf16_tcgen05_kernel<K_BIAS, true, true>;  // the only fused-residual form
  switch (kind) {
    case K_NONE: return pick_kernel<K_NONE>(tma_out);
    case K_BIAS: return pick_kernel<K_BIAS>(tma_out);
    case K_GELU_FAST: return pick_kernel<K_GELU_FAST>(tma_out);
    case K_GELU_EXACT: return pick_kernel<K_GELU_EXACT>(tma_out);
    case K_GELU_TANHFIT: return pick_kernel<K_GELU_TANHFIT>(tma_out);
    case K_RELU_SCALE: return pick_kernel<K_RELU_SCALE>(tma_out);
    case K_BIAS_RELU: return pick_kernel<K_BIAS_RELU>(tma_out);
  }
  return nullptr;
}

}  // namespace

extern "C" int ruart_gemm_bf16(const void* A, long long lda, int a_parts, const void* W,
                               long long ldw, int w_parts, int M, int N, int Kp, int n_terms,
                               int epi, const float* bias, const float* scale, int scale_len,
                               float* out_f32, long long ldo_f32, void* out_bf16,
                               long long ldo_bf16, int out_parts, long long out_part_stride,
                               int fast_gelu, const void* residual_bf16, long long ld_res,
                               void* stream) {
  RUART_ARG_CHECK(M >= 0 && N > 0 && Kp > 0 && (Kp % BK) == 0);
  RUART_ARG_CHECK(a_parts >= 1 && a_parts <= 3 && w_parts >= 1 && w_parts <= 3);
  RUART_ARG_CHECK(n_terms == 1 || n_terms == 3 || n_terms == 6);
  RUART_ARG_CHECK((lda % 8) == 0 && (ldw % 8) == 0);
  RUART_ARG_CHECK((reinterpret_cast<uintptr_t>(A) & 15u) == 0 &&
                  (reinterpret_cast<uintptr_t>(W) & 15u) == 0);
  RUART_ARG_CHECK(out_f32 != nullptr || out_bf16 != nullptr);
  RUART_ARG_CHECK(out_parts >= 1 && out_parts <= 3);
  int kind = K_NONE;
  const float* vec = nullptr;
  int vec_stride = 1;
  switch (epi) {
    case RUART_EPI_NONE: kind = K_NONE; break;
    case RUART_EPI_BIAS: kind = K_BIAS; vec = bias; break;
    case RUART_EPI_BIAS_GELU:
      kind = fast_gelu == 2 ? K_GELU_TANHFIT : (fast_gelu ? K_GELU_FAST : K_GELU_EXACT);
      vec = bias;
      break;
    case RUART_EPI_BIAS_RELU: kind = K_BIAS_RELU; vec = bias; break;
    case RUART_EPI_RELU_SCALE:
      kind = K_RELU_SCALE;
      vec = scale;
      RUART_ARG_CHECK(scale_len == 1 || scale_len >= N);
      vec_stride = (scale_len > 1) ? 1 : 0;
      break;
    default: RUART_ARG_CHECK(!"unknown epilogue");
  }
  if (kind != K_NONE) RUART_ARG_CHECK(vec != nullptr);
  if (M == 0) return RUART_OK;

  // term tables: (a part, w part).  Order: smallest products first so they are not swamped.
  uint32_t ta = 0, tb = 0;
  if (n_terms == 3) {
    RUART_ARG_CHECK(a_parts >= 2 && w_parts >= 2);
    ta = 0x001;  // (1,0) (0,1) (0,0)
    tb = 0x010;
  } else if (n_terms == 6) {
    RUART_ARG_CHECK(a_parts >= 3 && w_parts >= 3);
    ta = 0x001021;  // (1,1) (2,0) (0,2) (1,0) (0,1) (0,0)
    tb = 0x010201;
  }

  // plain bf16 output with TMA-compatible alignment -> staged TMA-store epilogue
  const bool tma_out = out_f32 == nullptr && out_bf16 != nullptr && out_parts == 1 &&
                       (ldo_bf16 % 8) == 0 && (reinterpret_cast<uintptr_t>(out_bf16) & 15u) == 0;

  GemmParams p;
  p.M = M;
  p.N = N;
  p.Kp = Kp;
  p.block_n = pick_block_n(N, tma_out ? 64 : 32);
  p.n_terms = n_terms;
  p.term_a = ta;
  p.term_b = tb;
  p.vec = vec;
  p.vec_stride = vec_stride;
  p.out_f32 = out_f32;
  p.ldo_f32 = ldo_f32;
  p.out_bf16 = reinterpret_cast<__nv_bfloat16*>(out_bf16);
  p.ldo_bf16 = ldo_bf16;
  p.out_parts = out_parts;
  p.out_part_stride = out_part_stride;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(residual_bf16);
  p.ld_res = ld_res;
  p.in_stats = nullptr;
  p.vec2 = nullptr;
  p.out_stats = nullptr;
  p.ln_inv_dim = 0.f;
  p.ln_eps = 0.f;
  if (residual_bf16 != nullptr) {
    // fused residual: staged TMA-store epilogue only, whole 64-column chunks, 16-byte aligned rows
    RUART_ARG_CHECK(tma_out && kind == K_BIAS && (N % 64) == 0 && (ld_res % 8) == 0 &&
                    (reinterpret_cast<uintptr_t>(residual_bf16) & 15u) == 0);
  }

  CUtensorMap tma, tmb, tmc;
  int rc = make_tmap_bf16(&tma, A, M, (long long)a_parts * Kp, lda, BM);
  if (rc != RUART_OK) return rc;
  rc = make_tmap_bf16(&tmb, W, N, (long long)w_parts * Kp, ldw, p.block_n);
  if (rc != RUART_OK) return rc;
  if (tma_out) {
    rc = make_tmap_bf16(&tmc, out_bf16, M, N, ldo_bf16, BM);
    if (rc != RUART_OK) return rc;
  } else {
    tmc = tma;
  }
  CUtensorMap tmr = tmc;
  if (residual_bf16 != nullptr) {
    rc = make_tmap_bf16(&tmr, residual_bf16, M, N, ld_res, BM);
    if (rc != RUART_OK) return rc;
  }

  const bool has_res = residual_bf16 != nullptr;
  // CTA-pair kernel: plain bf16 output, whole 256-column tiles, one MMA term, a full wave of pairs
  static const bool one_cta = getenv("RUART_GEMM_1CTA") != nullptr;  // A/B aid
  GemmKernel kern2 = nullptr;
  static RuartDeviceOnce pair_launch_bad;  // this device rejected the CTA-pair launch
  if (!one_cta && !pair_launch_bad.done() && tma_out && n_terms == 1 && (N % MAX_BN) == 0 && M >= 2 * BM * 8) {
    if (has_res) kern2 = gemm_bf16_2cta_kernel<K_BIAS, true, 0>;
    else if (kind == K_BIAS) kern2 = gemm_bf16_2cta_kernel<K_BIAS, false, 0>;
    else if (kind == K_GELU_FAST) kern2 = gemm_bf16_2cta_kernel<K_GELU_FAST, false, 0>;
    else if (kind == K_GELU_TANHFIT) kern2 = gemm_bf16_2cta_kernel<K_GELU_TANHFIT, false, 0>;
  }
  if (kern2 != nullptr) {
    CUtensorMap tmb2;
    rc = make_tmap_bf16(&tmb2, W, N, (long long)w_parts * Kp, ldw, B2_ROWS);
    if (rc != RUART_OK) return rc;
    static RuartDeviceOnce attr2[4];
    const int slot2 = has_res ? 2 : (kind == K_BIAS ? 0 : (kind == K_GELU_FAST ? 1 : 3));
    if (!attr2[slot2].done()) {
      RUART_CUDA_CHECK(cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            GEMM2_SMEM_BYTES));
      attr2[slot2].set();
    }
    const int pair_tiles = ((M + 2 * BM - 1) / (2 * BM)) * (N / MAX_BN);
    int grid2 = ruart_num_sms() & ~1;
    if (grid2 > 2 * pair_tiles) grid2 = 2 * pair_tiles;
    kern2<<<grid2, GEMM_THREADS, GEMM2_SMEM_BYTES, (cudaStream_t)stream>>>(tma, tmb2, tmc, tmr, p);
    const cudaError_t e2 = cudaGetLastError();
    if (e2 == cudaSuccess) return RUART_OK;
    // a device / partition that cannot co-schedule CTA pairs rejects the cluster launch up front
    // (nothing ran): remember it and use the 1-CTA kernel from now on; anything else is an error
    if (e2 != cudaErrorInvalidConfiguration && e2 != cudaErrorLaunchOutOfResources &&
        e2 != cudaErrorNotSupported && e2 != cudaErrorInvalidValue) {
      ruart_set_error("%s:%d: launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e2));
      return RUART_ERR_CUDA;
    }
    pair_launch_bad.set();
  }
  GemmKernel kern = kernel_for(kind, tma_out, has_res);
  static RuartDeviceOnce attr_set[K_NUM][3];
  const int slot = has_res ? 2 : (tma_out ? 1 : 0);
  if (!attr_set[kind][slot].done()) {
    RUART_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          GEMM_SMEM_BYTES));
    attr_set[kind][slot].set();
  }
  const int m_tiles = (M + BM - 1) / BM;
  const int n_tiles = (N + p.block_n - 1) / p.block_n;
  const int total = m_tiles * n_tiles;
  const int grid = total < ruart_num_sms() ? total : ruart_num_sms();
  kern<<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, (cudaStream_t)stream>>>(tma, tmb, tmc, tmr, p);
  RUART_LAUNCH_CHECK();
  return RUART_OK;
}

// CTA-pair GEMM with a folded LayerNorm (see the comment above gemm_bf16_2cta_kernel; bf16 BERT encoder only).
//   fold 1: out = act( LayerNorm(A) W^T + b )   given A = the stored pre-LayerNorm rows, W = W0 * gamma (host),
//           vec = W0 beta + b, vec2 = colsum(W), in_stats = partial sums of A's rows (ln_dim = K)
//           act: RUART_EPI_BIAS (none) or RUART_EPI_BIAS_GELU (fitted-tanh erf-GELU)
//   fold 2: out = A W^T + LayerNorm(residual) + b, pre-LayerNorm rows + their partial sums (out_stats):
//           vec = beta + b, vec2 = gamma, in_stats = partial sums of the residual's rows (ln_dim = N)
// Needs the CTA-pair kernel (M >= 2048, N % 256 == 0, cluster launch available): returns RUART_ERR_ARG
// otherwise and the caller keeps the explicit LayerNorm pass.
extern "C" int ruart_gemm_bf16_fold(const void* A, long long lda, const void* W, long long ldw, int M, int N,
                                    int Kp, int fold, int epi, const float* vec, const float* vec2,
                                    const float* in_stats, float ln_eps, void* out_bf16, long long ldo,
                                    const void* residual_bf16, long long ld_res, float* out_stats,
                                    void* stream) {
  RUART_ARG_CHECK(M >= 2 * BM * 8 && N > 0 && (N % MAX_BN) == 0 && Kp > 0 && (Kp % BK) == 0);
  RUART_ARG_CHECK((lda % 8) == 0 && (ldw % 8) == 0 && (ldo % 8) == 0);
  RUART_ARG_CHECK((reinterpret_cast<uintptr_t>(A) & 15u) == 0 && (reinterpret_cast<uintptr_t>(W) & 15u) == 0 &&
                  (reinterpret_cast<uintptr_t>(out_bf16) & 15u) == 0);
  RUART_ARG_CHECK(vec != nullptr && vec2 != nullptr && in_stats != nullptr &&
                  (reinterpret_cast<uintptr_t>(in_stats) & 15u) == 0);
  RUART_ARG_CHECK(fold == 1 || fold == 2);
  RUART_ARG_CHECK((fold == 1 ? Kp : N) == 768);  // the epilogues sum 6 partial-sum slots = 3 column tiles x 2 groups
  GemmKernel kern = nullptr;
  int slot = 0;
  if (fold == 1) {
    RUART_ARG_CHECK(residual_bf16 == nullptr && (epi == RUART_EPI_BIAS || epi == RUART_EPI_BIAS_GELU));
    if (epi == RUART_EPI_BIAS) { kern = gemm_bf16_2cta_kernel<K_BIAS, false, 1>; slot = 0; }
    else { kern = gemm_bf16_2cta_kernel<K_GELU_TANHFIT, false, 1>; slot = 1; }
  } else {
    RUART_ARG_CHECK(residual_bf16 != nullptr && out_stats != nullptr && epi == RUART_EPI_BIAS && (ld_res % 8) == 0 &&
                    (reinterpret_cast<uintptr_t>(residual_bf16) & 15u) == 0 && N <= 4 * MAX_BN);
    kern = gemm_bf16_2cta_kernel<K_BIAS, true, 2>;
    slot = 2;
  }
  GemmParams p;
  p.M = M; p.N = N; p.Kp = Kp; p.block_n = MAX_BN; p.n_terms = 1; p.term_a = 0; p.term_b = 0;
  p.vec = vec; p.vec_stride = 1;
  p.out_f32 = nullptr; p.ldo_f32 = 0;
  p.out_bf16 = reinterpret_cast<__nv_bfloat16*>(out_bf16); p.ldo_bf16 = ldo; p.out_parts = 1; p.out_part_stride = 0;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(residual_bf16); p.ld_res = ld_res;
  p.in_stats = reinterpret_cast<const float2*>(in_stats);
  p.vec2 = vec2;
  p.out_stats = reinterpret_cast<float2*>(out_stats);
  p.ln_inv_dim = 1.0f / static_cast<float>(fold == 1 ? Kp : N);
  p.ln_eps = ln_eps;
  if (M == 0) return RUART_OK;
  CUtensorMap tma, tmb, tmc, tmr;
  int rc = make_tmap_bf16(&tma, A, M, Kp, lda, BM);
  if (rc != RUART_OK) return rc;
  rc = make_tmap_bf16(&tmb, W, N, Kp, ldw, B2_ROWS);
  if (rc != RUART_OK) return rc;
  rc = make_tmap_bf16(&tmc, out_bf16, M, N, ldo, BM);
  if (rc != RUART_OK) return rc;
  tmr = tmc;
  if (residual_bf16 != nullptr) {
    rc = make_tmap_bf16(&tmr, residual_bf16, M, N, ld_res, BM);
    if (rc != RUART_OK) return rc;
  }
  static RuartDeviceOnce attr[3];
  if (!attr[slot].done()) {
    RUART_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM2_SMEM_BYTES));
    attr[slot].set();
  }
  const int pair_tiles = ((M + 2 * BM - 1) / (2 * BM)) * (N / MAX_BN);
  int grid2 = ruart_num_sms() & ~1;
  if (grid2 > 2 * pair_tiles) grid2 = 2 * pair_tiles;
  kern<<<grid2, GEMM_THREADS, GEMM2_SMEM_BYTES, (cudaStream_t)stream>>>(tma, tmb, tmc, tmr, p);
  RUART_LAUNCH_CHECK();
  return RUART_OK;
}

// Fused query/key/value GEMM + self-attention of the folded bf16 encoder (qkv_attn_2cta_kernel above).
//   A [M, Kp] bf16: the stored pre-LayerNorm rows; in_stats their partial sums
//   W [n_heads * 192, Kp] bf16: (W0 * gamma) with the rows permuted head-major [q_h / 8 ; k_h ; v_h]
//   vec / vec2 [n_heads * 192]: W0 beta + b and colsum(W), permuted and scaled the same way
//   tile_meta / tok_bounds: from ruart_seq_tiles (every sequence <= 128 tokens)
//   out [M, n_heads * 64] bf16: the attention context (input of BertSelfOutput.dense)
extern "C" int ruart_qkv_attention_fold(const void* A, long long lda, const void* W, long long ldw, int M, int Kp,
                                        int n_heads, const float* vec, const float* vec2, const float* in_stats,
                                        float ln_eps, const int32_t* tile_meta, const int32_t* tok_bounds,
                                        void* out_bf16, long long ldo, void* stream) {
  RUART_ARG_CHECK(M >= 2 * BM * 8 && Kp == 768 && n_heads >= 1 && n_heads <= 64);
  RUART_ARG_CHECK((lda % 8) == 0 && (ldw % 8) == 0 && (ldo % 8) == 0);
  RUART_ARG_CHECK((reinterpret_cast<uintptr_t>(A) & 15u) == 0 && (reinterpret_cast<uintptr_t>(W) & 15u) == 0 &&
                  (reinterpret_cast<uintptr_t>(out_bf16) & 15u) == 0);
  RUART_ARG_CHECK(vec != nullptr && vec2 != nullptr && in_stats != nullptr && tile_meta != nullptr &&
                  tok_bounds != nullptr && (reinterpret_cast<uintptr_t>(in_stats) & 15u) == 0 &&
                  (reinterpret_cast<uintptr_t>(tok_bounds) & 7u) == 0);
  QaParams p;
  p.M = M; p.Kp = Kp; p.n_heads = n_heads;
  p.meta = tile_meta;
  p.bounds = reinterpret_cast<const int2*>(tok_bounds);
  p.in_stats = reinterpret_cast<const float2*>(in_stats);
  p.vec = vec; p.vec2 = vec2;
  p.ln_inv_dim = 1.0f / static_cast<float>(Kp);
  p.ln_eps = ln_eps;
  p.out = reinterpret_cast<__nv_bfloat16*>(out_bf16);
  p.ldo = ldo;
  CUtensorMap tma, tmb;
  int rc = make_tmap_bf16(&tma, A, M, Kp, lda, BM);
  if (rc != RUART_OK) return rc;
  rc = make_tmap_bf16(&tmb, W, (long long)n_heads * QA_BN, Kp, ldw, QA_B_ROWS);
  if (rc != RUART_OK) return rc;
  static RuartDeviceOnce attr;
  if (!attr.done()) {
    RUART_CUDA_CHECK(cudaFuncSetAttribute(qkv_attn_2cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          QA_SMEM_BYTES));
    attr.set();
  }
  const int grid = ruart_num_sms() & ~1;  // persistent; the tile count is read from tile_meta on the device
  qkv_attn_2cta_kernel<<<grid, QA_THREADS, QA_SMEM_BYTES, (cudaStream_t)stream>>>(tma, tmb, p);
  RUART_LAUNCH_CHECK();
  return RUART_OK;
}
