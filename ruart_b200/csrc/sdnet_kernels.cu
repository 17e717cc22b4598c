// fp32 kernels of the SDNet fusion stack (reference Models/Layers.py, Models/SDNet.py) for sm_100a.
// All of them are small, HBM/L2- or latency-bound; the dense projections that feed them go
// through the tcgen05 GEMM (gemm_tcgen05.cu) with split-bf16 operands.
#include <cstdlib>

#include "common.cuh"
#include "ruart_b200.h"

namespace {

using namespace ruart;

// ------------------------------------------------------------------------------------------
// Row gather: dst[dst_idx[k]] <- table[src_idx[k]]  (D floats per row, independent row pitches).
// Covers nn.Embedding lookups (SDNet.py:447-492), the pre-align pack/unpack loops
// (SDNet.py:504-520,540-550) and the item -> slot scatter (SDNet.py:300-318).
// idx arrays are int64 (the collate's dtype) or int32; a null dst_idx means k itself.
template <typename I>
__global__ void gather_rows_kernel(const float* __restrict__ src, long long src_pitch,
                                   const I* __restrict__ src_idx, float* __restrict__ dst,
                                   long long dst_pitch, const I* __restrict__ dst_idx,
                                   float* __restrict__ dst2, long long dst2_pitch, long long n,
                                   int D) {
  const int lanes = (D + 3) / 4;  // float4 lanes per row
  const long long total = n * lanes;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long k = i / lanes;
    const int c = static_cast<int>(i - k * lanes) * 4;
    const long long s = src_idx ? static_cast<long long>(src_idx[k]) : k;
    const long long d = dst_idx ? static_cast<long long>(dst_idx[k]) : k;
    if (s < 0 || d < 0) continue;
    const float* sp = src + s * src_pitch + c;
    float* dp = dst + d * dst_pitch + c;
    if (c + 4 <= D && ((reinterpret_cast<uintptr_t>(sp) | reinterpret_cast<uintptr_t>(dp)) & 15u) == 0) {
      const float4 v = *reinterpret_cast<const float4*>(sp);
      *reinterpret_cast<float4*>(dp) = v;
      if (dst2) {
        float* d2 = dst2 + d * dst2_pitch + c;
        d2[0] = v.x; d2[1] = v.y; d2[2] = v.z; d2[3] = v.w;
      }
    } else {
      for (int e = 0; e < 4 && c + e < D; ++e) {
        const float v = sp[e];
        dp[e] = v;
        if (dst2) dst2[d * dst2_pitch + c + e] = v;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Whole-tensor LayerNorm, F.layer_norm(x, x.size()) of Layers.py:167-168: ONE mean / variance over
// every element of the [rows, cols] block (pads included), eps 1e-5, no affine.
// Pass 1 writes per-CTA partial (sum, sum of squares) in double; pass 2 re-reduces the partials in
// a fixed order (deterministic) and normalises in place.
constexpr int LN_THREADS = 256;
constexpr int LN_MAX_PARTS = 1024;

__global__ void __launch_bounds__(LN_THREADS)
whole_ln_stats_kernel(const float* __restrict__ x, long long rows, int cols, long long pitch,
                      double* __restrict__ partials) {
  __shared__ double s_sum[LN_THREADS / 32], s_sq[LN_THREADS / 32];
  const long long total = rows * cols;
  double sum = 0.0, sq = 0.0;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / cols;
    const float v = x[r * pitch + (i - r * cols)];
    sum += v;
    sq += static_cast<double>(v) * v;
  }
  sum = warp_sum_d(sum);
  sq = warp_sum_d(sq);
  if ((threadIdx.x & 31) == 0) {
    s_sum[threadIdx.x >> 5] = sum;
    s_sq[threadIdx.x >> 5] = sq;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < LN_THREADS / 32; ++w) {
      a += s_sum[w];
      b += s_sq[w];
    }
    partials[2 * blockIdx.x] = a;
    partials[2 * blockIdx.x + 1] = b;
  }
}

__global__ void __launch_bounds__(LN_THREADS)
whole_ln_apply_kernel(float* __restrict__ x, long long rows, int cols, long long pitch,
                      const double* __restrict__ partials, int n_parts, float eps,
                      float* __restrict__ stats_out) {
  __shared__ float s_mean, s_inv;
  if (threadIdx.x < 32) {
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < n_parts; i += 32) {
      a += partials[2 * i];
      b += partials[2 * i + 1];
    }
    a = warp_sum_d(a);
    b = warp_sum_d(b);
    if (threadIdx.x == 0) {
      const double n = static_cast<double>(rows) * cols;
      const double mean = a / n;
      double var = b / n - mean * mean;
      if (var < 0.0) var = 0.0;
      s_mean = static_cast<float>(mean);
      s_inv = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
      if (stats_out != nullptr && blockIdx.x == 0) {   // (mean, rstd) for the backward pass
        stats_out[0] = s_mean;
        stats_out[1] = s_inv;
      }
    }
  }
  __syncthreads();
  const float mean = s_mean, inv = s_inv;
  const long long total = rows * cols;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / cols;
    float* p = x + r * pitch + (i - r * cols);
    *p = (*p - mean) * inv;
  }
}

// ------------------------------------------------------------------------------------------
// Fused tail of Attention.forward (Layers.py:272-288) after the two projections:
//   scores = p1 p2^T ; masked_fill(x2_mask == 0, -inf) ; softmax over keys ; out = alpha x3
// p1 [B, L1, Hd] = relu(x1 W^T) * D,  p2 [B, L2, Hd] = relu(x2 W^T) come from the GEMM.
// One CTA per (batch b, tile of QT query rows).  Phase 1: each warp owns keys k = w, w+8, ...;
// lanes stride the hidden dim (coalesced p2 reads), p1 tile lives in smem.  Phase 2: one warp per
// query row does the masked softmax.  Phase 3: threads stride the value dim.
// If add_to_out != 0 the result is added to what `out` already holds (x_od_ocr += pos_att,
// SDNet.py:399-401).
constexpr int ATT_QT = 16;
constexpr int ATT_THREADS = 256;

__global__ void __launch_bounds__(ATT_THREADS)
attention_tail_kernel(const float* __restrict__ p1, long long p1_pitch, const float* __restrict__ p2,
                      long long p2_pitch, int Hd, const uint8_t* __restrict__ mask,
                      const float* __restrict__ x3, long long x3_pitch, int D3,
                      float* __restrict__ out, long long out_pitch, int L1, int L2, int add_to_out) {
  extern __shared__ float smem[];
  float* s_p1 = smem;                    // [QT][Hd]
  float* s_sc = smem + ATT_QT * Hd;      // [QT][L2]
  const int b = blockIdx.y;
  const int q0 = blockIdx.x * ATT_QT;
  const int nq = min(ATT_QT, L1 - q0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* p1b = p1 + (static_cast<long long>(b) * L1 + q0) * p1_pitch;
  const float* p2b = p2 + static_cast<long long>(b) * L2 * p2_pitch;
  for (int i = threadIdx.x; i < ATT_QT * Hd; i += ATT_THREADS) {
    const int q = i / Hd, h = i - q * Hd;
    s_p1[i] = (q < nq) ? p1b[q * p1_pitch + h] : 0.f;
  }
  __syncthreads();
  // phase 1: scores
  for (int k = warp; k < L2; k += ATT_THREADS / 32) {
    float acc[ATT_QT];
#pragma unroll
    for (int q = 0; q < ATT_QT; ++q) acc[q] = 0.f;
    const float* kr = p2b + k * p2_pitch;
    for (int h = lane; h < Hd; h += 32) {
      const float kv = kr[h];
#pragma unroll
      for (int q = 0; q < ATT_QT; ++q) acc[q] = fmaf(s_p1[q * Hd + h], kv, acc[q]);
    }
    const bool keep = mask[static_cast<long long>(b) * L2 + k] != 0;
#pragma unroll
    for (int q = 0; q < ATT_QT; ++q) {
      const float v = warp_sum(acc[q]);
      if (lane == 0) s_sc[q * L2 + k] = keep ? v : -INFINITY;
    }
  }
  __syncthreads();
  // phase 2: softmax over keys (F.softmax(scores, dim=1), Layers.py:284)
  for (int q = warp; q < nq; q += ATT_THREADS / 32) {
    float* row = s_sc + q * L2;
    float m = -INFINITY;
    for (int k = lane; k < L2; k += 32) m = fmaxf(m, row[k]);
    m = warp_max(m);
    float s = 0.f;
    for (int k = lane; k < L2; k += 32) {
      const float e = expf(row[k] - m);
      row[k] = e;
      s += e;
    }
    s = warp_sum(s);
    const float inv = 1.0f / s;
    for (int k = lane; k < L2; k += 32) row[k] *= inv;
  }
  __syncthreads();
  // phase 3: out = alpha @ x3
  const float* x3b = x3 + static_cast<long long>(b) * L2 * x3_pitch;
  float* ob = out + (static_cast<long long>(b) * L1 + q0) * out_pitch;
  for (int d = threadIdx.x; d < D3; d += ATT_THREADS) {
    float acc[ATT_QT];
#pragma unroll
    for (int q = 0; q < ATT_QT; ++q) acc[q] = 0.f;
    for (int k = 0; k < L2; ++k) {
      const float v = x3b[k * x3_pitch + d];
#pragma unroll
      for (int q = 0; q < ATT_QT; ++q) acc[q] = fmaf(s_sc[q * L2 + k], v, acc[q]);
    }
#pragma unroll
    for (int q = 0; q < ATT_QT; ++q)
      if (q < nq) {
        float* o = ob + q * out_pitch + d;
        *o = add_to_out ? (*o + acc[q]) : acc[q];
      }
  }
}

// Register-tiled version for L2 <= 128 (every attention of the shipped conf): one CTA per
// (batch, 32 query rows), 256 threads = 8 query groups (4 rows each) x 32 lanes.
//   phase 1  S[32 x L2]: k-chunks of 32 of p1 / p2 staged in smem (row pitch 36 floats: 16-byte
//            aligned, conflict-free); thread tile 4 queries x 4 keys (keys lane, lane+32, ...)
//   phase 2  masked softmax per row (4 rows per warp), in smem
//   phase 3  O[32 x D3] = A x3: x3 staged in [32 keys x 128 dims] chunks; thread tile 4 queries x
//            4 consecutive dims (float4)
constexpr int AT_Q = 32, AT_KC = 32, AT_L2 = 128, AT_P = 36, AT_SP = 132, AT_DC = 128;

__global__ void __launch_bounds__(256)
attention_tail_tiled_kernel(const float* __restrict__ p1, long long p1_pitch,
                            const float* __restrict__ p2, long long p2_pitch, int Hd,
                            const uint8_t* __restrict__ mask, const float* __restrict__ x3,
                            long long x3_pitch, int D3, float* __restrict__ out,
                            long long out_pitch, int L1, int L2, int add_to_out) {
  __shared__ __align__(16) float s_a[AT_Q * AT_P];       // p1 chunk [32 q][32 k]
  __shared__ __align__(16) float s_b[AT_L2 * AT_P];      // p2 chunk [128 keys][32 k]; later x3 chunk
  __shared__ __align__(16) float s_s[AT_Q * AT_SP];      // scores / probabilities [32 q][L2]
  const int b = blockIdx.y;
  const int q0 = blockIdx.x * AT_Q;
  const int tid = threadIdx.x, ty = tid >> 5, tx = tid & 31;
  const float* p1b = p1 + (static_cast<long long>(b) * L1 + q0) * p1_pitch;
  const float* p2b = p2 + static_cast<long long>(b) * L2 * p2_pitch;
  const int nq = min(AT_Q, L1 - q0);
  const int nj = (L2 + 31) >> 5;  // 32-key groups actually present (warp-uniform)
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < Hd; k0 += AT_KC) {
    __syncthreads();
    // stage p1 chunk: 32 rows x 32 k (one float per thread x 4)
    for (int i = tid; i < AT_Q * AT_KC; i += 256) {
      const int r = i >> 5, k = i & 31;
      s_a[r * AT_P + k] = (r < nq && k0 + k < Hd) ? p1b[r * p1_pitch + k0 + k] : 0.f;
    }
    for (int i = tid; i < nj * 32 * AT_KC; i += 256) {
      const int r = i >> 5, k = i & 31;
      s_b[r * AT_P + k] = (r < L2 && k0 + k < Hd) ? p2b[r * p2_pitch + k0 + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < AT_KC; k += 4) {
      float4 av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = *reinterpret_cast<const float4*>(s_a + (4 * ty + i) * AT_P + k);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < nj) bv[j] = *reinterpret_cast<const float4*>(s_b + (tx + 32 * j) * AT_P + k);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j < nj) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            acc[i][j] = fmaf(av[i].x, bv[j].x, acc[i][j]);
            acc[i][j] = fmaf(av[i].y, bv[j].y, acc[i][j]);
            acc[i][j] = fmaf(av[i].z, bv[j].z, acc[i][j]);
            acc[i][j] = fmaf(av[i].w, bv[j].w, acc[i][j]);
          }
        }
      }
    }
  }
  // scores -> smem, masked
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int key = tx + 32 * j;
    if (key < L2) {
      const bool keep = mask[static_cast<long long>(b) * L2 + key] != 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) s_s[(4 * ty + i) * AT_SP + key] = keep ? acc[i][j] : -INFINITY;
    }
  }
  __syncthreads();
  // softmax over keys: warp ty owns rows 4ty .. 4ty+3
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float* row = s_s + (4 * ty + i) * AT_SP;
    float m = -INFINITY;
    for (int k = tx; k < L2; k += 32) m = fmaxf(m, row[k]);
    m = warp_max(m);
    float sum = 0.f;
    for (int k = tx; k < L2; k += 32) {
      const float e = expf(row[k] - m);
      row[k] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int k = tx; k < L2; k += 32) row[k] *= inv;
  }
  // out = A @ x3, 128 dims per pass, keys in chunks of 32 staged through s_b ([32][AT_SP])
  const float* x3b = x3 + static_cast<long long>(b) * L2 * x3_pitch;
  float* ob = out + (static_cast<long long>(b) * L1 + q0) * out_pitch;
  for (int d0 = 0; d0 < D3; d0 += AT_DC) {
    float o[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
    for (int kk = 0; kk < L2; kk += 32) {
      __syncthreads();
      for (int i = tid; i < 32 * AT_DC; i += 256) {
        const int r = i >> 7, d = i & 127;
        s_b[r * AT_SP + d] = (kk + r < L2 && d0 + d < D3) ? x3b[(kk + r) * x3_pitch + d0 + d] : 0.f;
      }
      __syncthreads();
      const int kmax = min(32, L2 - kk);
      for (int k = 0; k < kmax; ++k) {
        const float4 xv = *reinterpret_cast<const float4*>(s_b + k * AT_SP + 4 * tx);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float a = s_s[(4 * ty + i) * AT_SP + kk + k];
          o[i][0] = fmaf(a, xv.x, o[i][0]);
          o[i][1] = fmaf(a, xv.y, o[i][1]);
          o[i][2] = fmaf(a, xv.z, o[i][2]);
          o[i][3] = fmaf(a, xv.w, o[i][3]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int q = 4 * ty + i;
      if (q < nq) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int d = d0 + 4 * tx + j;
          if (d < D3) {
            float* op = ob + q * out_pitch + d;
            *op = add_to_out ? (*op + o[i][j]) : o[i][j];
          }
        }
      }
    }
  }
}

// Tensor-core form of the attention tail for the 2-part-split precision of the SDNet stack
// (bf16x2: every fp32 operand x is used as hi + lo bf16 parts and a product as the three terms
// hi*hi + lo*hi + hi*lo, ~2^-16 relative — the same rule the split GEMMs follow).  64 queries per
// CTA, one warp per 16 queries, mma.sync.m16n8k16 in the FlashAttention-2 register layout:
//   phase 1  S[16 x L2] = p1 p2^T over 64-wide k chunks; p1 / p2 chunks staged in shared memory as
//            hi | lo bf16 tiles (128-byte rows, XOR-swizzled 16-byte chunks)
//   phase 2  mask (-inf), softmax in registers (quad shuffles), P split into hi | lo A-fragments
//   phase 3  O[16 x D3] = P x3 over 64-wide dim chunks, x3 chunk staged as hi | lo tiles and read with
//            ldmatrix.trans
// L2 <= 128 (16 key tiles of accumulators per thread).  The fp32 CUDA-core kernel above stays for
// the 3-part (fp32-grade) mode and longer key lists.
constexpr int ATM_Q = 64, ATM_L2 = 128, ATM_THREADS = 128;

__device__ __forceinline__ void split_store4(uint8_t* hi_tile, uint8_t* lo_tile, int r, int col4,
                                             float4 v) {
  // four consecutive columns (col4 = first, multiple of 4) of row r -> 8 bytes in each tile
  const __nv_bfloat16 h0 = __float2bfloat16_rn(v.x), h1 = __float2bfloat16_rn(v.y);
  const __nv_bfloat16 h2 = __float2bfloat16_rn(v.z), h3 = __float2bfloat16_rn(v.w);
  const uint32_t hi01 = static_cast<uint32_t>(__bfloat16_as_ushort(h0)) |
                        (static_cast<uint32_t>(__bfloat16_as_ushort(h1)) << 16);
  const uint32_t hi23 = static_cast<uint32_t>(__bfloat16_as_ushort(h2)) |
                        (static_cast<uint32_t>(__bfloat16_as_ushort(h3)) << 16);
  const uint32_t lo01 = pack_bf16x2(v.x - __bfloat162float(h0), v.y - __bfloat162float(h1));
  const uint32_t lo23 = pack_bf16x2(v.z - __bfloat162float(h2), v.w - __bfloat162float(h3));
  const int off = tile_off(r, col4 >> 3) + ((col4 & 4) << 1);
  *reinterpret_cast<uint2*>(hi_tile + off) = make_uint2(hi01, hi23);
  *reinterpret_cast<uint2*>(lo_tile + off) = make_uint2(lo01, lo23);
}

// Stage a [rows x 64] chunk (columns c0 .. c0+63 of `src`, rows >= n_rows and columns >= n_cols zero)
// as hi | lo bf16 tiles.  `src` rows are `pitch` floats apart.
__device__ __forceinline__ void stage_split_chunk(const float* __restrict__ src, long long pitch,
                                                  int n_rows, int n_cols, int c0, int rows,
                                                  uint8_t* hi_tile, uint8_t* lo_tile, int tid) {
  for (int i = tid; i < rows * 16; i += ATM_THREADS) {
    const int r = i >> 4, col4 = (i & 15) << 2;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < n_rows) {
      const float* p = src + static_cast<long long>(r) * pitch + c0 + col4;
      const int left = n_cols - (c0 + col4);
      if (left >= 4 && ((reinterpret_cast<uintptr_t>(p) & 15u) == 0)) {
        v = *reinterpret_cast<const float4*>(p);
      } else {
        if (left > 0) v.x = p[0];
        if (left > 1) v.y = p[1];
        if (left > 2) v.z = p[2];
        if (left > 3) v.w = p[3];
      }
    }
    split_store4(hi_tile, lo_tile, r, col4, v);
  }
}

struct TailX3 {
  const float* p[4];
};

__global__ void __launch_bounds__(ATM_THREADS)
attention_tail_mma_kernel(const float* __restrict__ p1, long long p1_pitch,
                          const float* __restrict__ p2, long long p2_pitch, int Hd,
                          const uint8_t* __restrict__ mask, const TailX3 x3s,
                          long long x3_pitch, int D3, float* __restrict__ out, long long out_pitch,
                          int L1, int L2, int add_to_out) {
  // blockIdx.z = attention head of a DeepAttention group (Layers.py:493-524): head z reads columns
  // [z Hd, (z + 1) Hd) of the stacked projections, its own x3 and writes columns [z D3, (z + 1) D3) of out
  p1 += blockIdx.z * Hd;
  p2 += blockIdx.z * Hd;
  out += blockIdx.z * D3;
  const float* __restrict__ x3 = x3s.p[blockIdx.z];
  // [p1 hi | p1 lo] 2 x 8 KB, [p2 / x3 hi | lo] 2 x 16 KB
  __shared__ __align__(128) uint8_t s_q[2][ATM_Q * 128];
  __shared__ __align__(128) uint8_t s_k[2][ATM_L2 * 128];
  const int b = blockIdx.y;
  const int q0 = blockIdx.x * ATM_Q;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int a_row = (lane & 7) + ((lane >> 3) & 1) * 8;
  const int a_chk = lane >> 4;
  const int b_row = lane & 7;
  const int b_chk = lane >> 3;
  const uint32_t aQh = smem_u32(s_q[0]), aQl = smem_u32(s_q[1]);
  const uint32_t aKh = smem_u32(s_k[0]), aKl = smem_u32(s_k[1]);
  const float* p1b = p1 + (static_cast<long long>(b) * L1 + q0) * p1_pitch;
  const float* p2b = p2 + static_cast<long long>(b) * L2 * p2_pitch;
  const float* x3b = x3 + static_cast<long long>(b) * L2 * x3_pitch;
  const int nq = min(ATM_Q, L1 - q0);
  const int n_kt = (L2 + 7) >> 3;    // 8-key tiles in use
  const int n_kk = (L2 + 15) >> 4;   // 16-key steps in use
  const int key_rows = n_kk * 16;    // staged key rows (zero padded)

  // ---- phase 1: scores
  float s[16][4];
#pragma unroll
  for (int nt = 0; nt < 16; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
  for (int k0 = 0; k0 < Hd; k0 += 64) {
    __syncthreads();
    stage_split_chunk(p1b, p1_pitch, nq, Hd, k0, ATM_Q, s_q[0], s_q[1], tid);
    stage_split_chunk(p2b, p2_pitch, L2, Hd, k0, key_rows, s_k[0], s_k[1], tid);
    __syncthreads();
    uint32_t qh[4][4], ql[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const int off = tile_off(warp * 16 + a_row, 2 * ks + a_chk);
      ldsm_x4(aQh + off, qh[ks][0], qh[ks][1], qh[ks][2], qh[ks][3]);
      ldsm_x4(aQl + off, ql[ks][0], ql[ks][1], ql[ks][2], ql[ks][3]);
    }
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) {
      if (nt < n_kt) {
#pragma unroll
        for (int kh = 0; kh < 2; ++kh) {
          uint32_t kb[4], kl[4];
          const int off = tile_off(nt * 8 + b_row, 4 * kh + b_chk);
          ldsm_x4(aKh + off, kb[0], kb[1], kb[2], kb[3]);
          ldsm_x4(aKl + off, kl[0], kl[1], kl[2], kl[3]);
#pragma unroll
          for (int k2 = 0; k2 < 2; ++k2) {
            const int ks = 2 * kh + k2;
            mma_bf16_16816(s[nt], ql[ks][0], ql[ks][1], ql[ks][2], ql[ks][3], kb[2 * k2], kb[2 * k2 + 1]);
            mma_bf16_16816(s[nt], qh[ks][0], qh[ks][1], qh[ks][2], qh[ks][3], kl[2 * k2], kl[2 * k2 + 1]);
            mma_bf16_16816(s[nt], qh[ks][0], qh[ks][1], qh[ks][2], qh[ks][3], kb[2 * k2], kb[2 * k2 + 1]);
          }
        }
      }
    }
  }
  // ---- phase 2: masked softmax over keys; thread holds rows g (regs 0,1) and g+8 (regs 2,3)
  float mA = -INFINITY, mB = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < 16; ++nt) {
    const int k0 = nt * 8 + 2 * t;
    const bool v0 = k0 < L2 && mask[static_cast<long long>(b) * L2 + k0] != 0;
    const bool v1 = k0 + 1 < L2 && mask[static_cast<long long>(b) * L2 + k0 + 1] != 0;
    s[nt][0] = v0 ? s[nt][0] : -INFINITY;
    s[nt][1] = v1 ? s[nt][1] : -INFINITY;
    s[nt][2] = v0 ? s[nt][2] : -INFINITY;
    s[nt][3] = v1 ? s[nt][3] : -INFINITY;
    mA = fmaxf(mA, fmaxf(s[nt][0], s[nt][1]));
    mB = fmaxf(mB, fmaxf(s[nt][2], s[nt][3]));
  }
  mA = fmaxf(mA, __shfl_xor_sync(0xffffffffu, mA, 1));
  mA = fmaxf(mA, __shfl_xor_sync(0xffffffffu, mA, 2));
  mB = fmaxf(mB, __shfl_xor_sync(0xffffffffu, mB, 1));
  mB = fmaxf(mB, __shfl_xor_sync(0xffffffffu, mB, 2));
  float lA = 0.f, lB = 0.f;
#pragma unroll
  for (int nt = 0; nt < 16; ++nt) {
    s[nt][0] = expf(s[nt][0] - mA);
    s[nt][1] = expf(s[nt][1] - mA);
    s[nt][2] = expf(s[nt][2] - mB);
    s[nt][3] = expf(s[nt][3] - mB);
    lA += s[nt][0] + s[nt][1];
    lB += s[nt][2] + s[nt][3];
  }
  lA += __shfl_xor_sync(0xffffffffu, lA, 1);
  lA += __shfl_xor_sync(0xffffffffu, lA, 2);
  lB += __shfl_xor_sync(0xffffffffu, lB, 1);
  lB += __shfl_xor_sync(0xffffffffu, lB, 2);
  const float iA = 1.0f / lA, iB = 1.0f / lB;
  // probabilities as hi | lo A-fragments per 16-key step
  uint32_t ph[8][4], pl[8][4];
#pragma unroll
  for (int kk = 0; kk < 8; ++kk) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {  // key tiles 2kk, 2kk+1
      const float a0 = s[2 * kk + h][0] * iA, a1 = s[2 * kk + h][1] * iA;
      const float b0 = s[2 * kk + h][2] * iB, b1 = s[2 * kk + h][3] * iB;
      const __nv_bfloat162 ha = __floats2bfloat162_rn(a0, a1), hb = __floats2bfloat162_rn(b0, b1);
      ph[kk][2 * h] = *reinterpret_cast<const uint32_t*>(&ha);
      ph[kk][2 * h + 1] = *reinterpret_cast<const uint32_t*>(&hb);
      pl[kk][2 * h] = pack_bf16x2(a0 - __low2float(ha), a1 - __high2float(ha));
      pl[kk][2 * h + 1] = pack_bf16x2(b0 - __low2float(hb), b1 - __high2float(hb));
    }
  }
  // ---- phase 3: O = P x3, 64 output dims per pass
  float* ob = out + (static_cast<long long>(b) * L1 + q0) * out_pitch;
  const int rA = warp * 16 + g, rB = rA + 8;
  for (int d0 = 0; d0 < D3; d0 += 64) {
    __syncthreads();
    stage_split_chunk(x3b, x3_pitch, L2, D3, d0, key_rows, s_k[0], s_k[1], tid);
    __syncthreads();
    float o[8][4];
#pragma unroll
    for (int d = 0; d < 8; ++d) o[d][0] = o[d][1] = o[d][2] = o[d][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      if (kk < n_kk) {
#pragma unroll
        for (int dp = 0; dp < 4; ++dp) {
          uint32_t vh[4], vl[4];
          const int off = tile_off(kk * 16 + a_row, 2 * dp + a_chk);
          ldsm_x4_t(aKh + off, vh[0], vh[1], vh[2], vh[3]);
          ldsm_x4_t(aKl + off, vl[0], vl[1], vl[2], vl[3]);
          mma_bf16_16816(o[2 * dp], pl[kk][0], pl[kk][1], pl[kk][2], pl[kk][3], vh[0], vh[1]);
          mma_bf16_16816(o[2 * dp], ph[kk][0], ph[kk][1], ph[kk][2], ph[kk][3], vl[0], vl[1]);
          mma_bf16_16816(o[2 * dp], ph[kk][0], ph[kk][1], ph[kk][2], ph[kk][3], vh[0], vh[1]);
          mma_bf16_16816(o[2 * dp + 1], pl[kk][0], pl[kk][1], pl[kk][2], pl[kk][3], vh[2], vh[3]);
          mma_bf16_16816(o[2 * dp + 1], ph[kk][0], ph[kk][1], ph[kk][2], ph[kk][3], vl[2], vl[3]);
          mma_bf16_16816(o[2 * dp + 1], ph[kk][0], ph[kk][1], ph[kk][2], ph[kk][3], vh[2], vh[3]);
        }
      }
    }
#pragma unroll
    for (int d = 0; d < 8; ++d) {
      const int col = d0 + d * 8 + 2 * t;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int r = h ? rB : rA;
        if (r < nq) {
          float* dst = ob + static_cast<long long>(r) * out_pitch + col;
          float x0 = o[d][2 * h], x1 = o[d][2 * h + 1];
          if (col < D3) dst[0] = add_to_out ? dst[0] + x0 : x0;
          if (col + 1 < D3) dst[1] = add_to_out ? dst[1] + x1 : x1;
        }
      }
    }
  }
}


// ------------------------------------------------------------------------------------------
// LinearSelfAttn + weighted_avg (Layers.py:328-341,529-534; SDNet.py:414-415):
//   alpha = softmax(mask(x w + b)) over the sequence ; out[b] = sum_l alpha_l x[b, l]
// One CTA per batch row.
__global__ void __launch_bounds__(256)
self_attn_pool_kernel(const float* __restrict__ x, long long x_pitch, int L, int D,
                      const uint8_t* __restrict__ mask, const float* __restrict__ w,
                      const float* __restrict__ bias, float* __restrict__ out, long long out_pitch) {
  extern __shared__ float s_al[];  // [L]
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* xb = x + static_cast<long long>(b) * L * x_pitch;
  for (int l = warp; l < L; l += 8) {
    float acc = 0.f;
    for (int d = lane; d < D; d += 32) acc = fmaf(xb[l * x_pitch + d], w[d], acc);
    acc = warp_sum(acc);
    if (lane == 0) s_al[l] = (mask[static_cast<long long>(b) * L + l] != 0) ? acc + bias[0] : -INFINITY;
  }
  __syncthreads();
  if (warp == 0) {
    float m = -INFINITY;
    for (int l = lane; l < L; l += 32) m = fmaxf(m, s_al[l]);
    m = warp_max(m);
    float s = 0.f;
    for (int l = lane; l < L; l += 32) {
      const float e = expf(s_al[l] - m);
      s_al[l] = e;
      s += e;
    }
    s = warp_sum(s);
    const float inv = 1.0f / s;
    for (int l = lane; l < L; l += 32) s_al[l] *= inv;
  }
  __syncthreads();
  for (int d = threadIdx.x; d < D; d += 256) {
    float acc = 0.f;
    for (int l = 0; l < L; ++l) acc = fmaf(s_al[l], xb[l * x_pitch + d], acc);
    out[static_cast<long long>(b) * out_pitch + d] = acc;
  }
}

// ------------------------------------------------------------------------------------------
// GetFinalScores.forward (Layers.py:373-419) with useES + no_answer, after the three
// `linear(h0)` products (attn, attn2, noanswer_linear; Layers.py:422,457) were computed by the GEMM
// into wy [B, 3*X] = [attn | attn2 | noanswer].  One CTA per question:
//   logit[m] = x[m] . (m < es_len ? wy_attn2 : wy_attn), pads -> -inf   (BilinearSeqAttn, :446-468)
//   xWh[m]   = x[m] . wy_noans, pads -> -inf ; pooled = softmax(xWh) x ; noans = w . pooled + b
//   probs = softmax([logits..., noans])                                              (:413-419)
// A sticky NaN flag replaces the reference's `assert isnan == 0` host syncs (Layers.py:430,462).
__global__ void __launch_bounds__(256)
final_scores_kernel(const float* __restrict__ x, long long x_pitch, int M, int X,
                    const float* __restrict__ wy, const uint8_t* __restrict__ mask, int es_len,
                    const float* __restrict__ noans_w, const float* __restrict__ noans_b,
                    float* __restrict__ probs, float* __restrict__ logits, int* nan_flag) {
  extern __shared__ float sm[];
  float* s_logit = sm;          // [M + 1]
  float* s_xwh = sm + (M + 1);  // [M]
  float* s_red = s_xwh + M;     // [8]
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* xb = x + static_cast<long long>(b) * M * x_pitch;
  const float* wa = wy + static_cast<long long>(b) * 3 * X;
  const float* wa2 = wa + X;
  const float* wn = wa + 2 * X;
  for (int m = warp; m < M; m += 8) {
    const float* xr = xb + m * x_pitch;
    const float* wv = (m < es_len) ? wa2 : wa;
    float a = 0.f, c = 0.f;
    for (int d = lane; d < X; d += 32) {
      const float xv = xr[d];
      a = fmaf(xv, wv[d], a);
      c = fmaf(xv, wn[d], c);
    }
    a = warp_sum(a);
    c = warp_sum(c);
    if (lane == 0) {
      const bool keep = mask[static_cast<long long>(b) * M + m] != 0;
      s_logit[m] = keep ? a : -INFINITY;
      s_xwh[m] = keep ? c : -INFINITY;
    }
  }
  __syncthreads();
  // softmax(xWh) (all threads need it): warp 0 normalises in place
  if (warp == 0) {
    float mx = -INFINITY;
    for (int m = lane; m < M; m += 32) mx = fmaxf(mx, s_xwh[m]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int m = lane; m < M; m += 32) {
      const float e = expf(s_xwh[m] - mx);
      s_xwh[m] = e;
      s += e;
    }
    s = warp_sum(s);
    const float inv = 1.0f / s;
    for (int m = lane; m < M; m += 32) s_xwh[m] *= inv;
  }
  __syncthreads();
  // noans = noanswer_w . (sum_m p_m x[m]) + b  =  sum_m p_m (x[m] . noanswer_w): one more row pass with independent
  // loads (the column walk it replaces was a chain of M dependent loads per thread: 84 us for 256 questions)
  float part = 0.f;
  for (int m = warp; m < M; m += 8) {
    const float pm = s_xwh[m];
    if (pm != 0.f) {                       // masked slots have p = 0 exactly
      const float* xr = xb + m * x_pitch;
      float a = 0.f;
      for (int d = lane; d < X; d += 32) a = fmaf(xr[d], noans_w[d], a);
      part = fmaf(pm, a, part);
    }
  }
  part = warp_sum(part);
  if (lane == 0) s_red[warp] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += s_red[w];
    s_logit[M] = t + noans_b[0];
  }
  __syncthreads();
  if (warp == 0) {
    float mx = -INFINITY;
    for (int m = lane; m <= M; m += 32) mx = fmaxf(mx, s_logit[m]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int m = lane; m <= M; m += 32) s += expf(s_logit[m] - mx);
    s = warp_sum(s);
    const float inv = 1.0f / s;
    bool bad = false;
    for (int m = lane; m <= M; m += 32) {
      const float lg = s_logit[m];
      const float p = expf(lg - mx) * inv;
      probs[static_cast<long long>(b) * (M + 1) + m] = p;
      if (logits) logits[static_cast<long long>(b) * (M + 1) + m] = lg;
      bad |= (p != p);
    }
    if (bad && nan_flag) atomicExch(nan_flag, 1);
  }
}

// ------------------------------------------------------------------------------------------
// LSTM cell update for the step-synchronous multi2one path (uni-LSTM 1388 -> 300 over item words,
// SDNet.py:137,270-271): rows are the still-active items at this step.
//   g = gx[row_gx[r]] (+ gh[r])  ;  i,f,g,o ;  c = f*c + i*g ;  h = o*tanh(c)
// h is written as fp32, as the 3-part bf16 split operand of the next step's recurrent GEMM, and,
// when this is an item's last word (last_step[r] == step), into its slot row (SDNet.py:304,310).
__global__ void lstm_cell_kernel(const float* __restrict__ gx, const int32_t* __restrict__ row_gx,
                                 const float* __restrict__ gh, float* __restrict__ c,
                                 __nv_bfloat16* __restrict__ h_split, int parts, int Kp, int H, int n_rows,
                                 const int32_t* __restrict__ last_step, int step,
                                 const long long* __restrict__ slot_off, float* __restrict__ slots) {
  const long long total = static_cast<long long>(n_rows) * H;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / H);
    const int j = static_cast<int>(i - static_cast<long long>(r) * H);
    const float* g = gx + static_cast<long long>(row_gx ? row_gx[r] : r) * 4 * H;
    float gi = g[j], gf = g[H + j], gg = g[2 * H + j], go = g[3 * H + j];
    float cp = 0.f;
    if (gh != nullptr) {
      const float* q = gh + static_cast<long long>(r) * 4 * H;
      gi += q[j]; gf += q[H + j]; gg += q[2 * H + j]; go += q[3 * H + j];
      cp = c[i];
    }
    const float si = 1.0f / (1.0f + expf(-gi));
    const float sf = 1.0f / (1.0f + expf(-gf));
    const float so = 1.0f / (1.0f + expf(-go));
    const float cn = sf * cp + si * tanhf(gg);
    const float hn = so * tanhf(cn);
    c[i] = cn;
    float rem = hn;
    for (int p = 0; p < parts; ++p) {
      const __nv_bfloat16 hb = __float2bfloat16_rn(rem);
      h_split[static_cast<long long>(r) * parts * Kp + static_cast<long long>(p) * Kp + j] = hb;
      rem -= __bfloat162float(hb);
    }
    if (last_step[r] == step) slots[slot_off[r] + j] = hn;
  }
}

// ------------------------------------------------------------------------------------------
// Answer-index rule of SDNetTrainer.predict (SDNetTrainer.py:402-412) on the device: walking the
// slots by descending probability, the loop stops at the no-answer column (last), skips the
// `<OCR>` end slot (index num_cnt-1) and every index >= num_cnt, and accepts the first index
// < num_cnt.  That is: the highest-probability index in {M} U {i < num_cnt - 1}.  Ties (equal
// probabilities) resolve to the smallest index; torch.sort's order among ties is unspecified.
__global__ void select_answers_kernel(const float* __restrict__ probs, const int32_t* __restrict__ num_cnt,
                                      int B, int M1, int no_answer, int32_t* __restrict__ out) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const float* p = probs + static_cast<long long>(b) * M1;
  const int n_ok = num_cnt[b] - 1;  // indices 0 .. n_ok-1 are real OCR items
  float best = -INFINITY;
  int best_i = 0x7fffffff;
  for (int i = lane; i < M1; i += 32) {
    const bool ok = (i < n_ok) || (no_answer && i == M1 - 1);
    const float v = p[i];
    if (ok && (v > best || (v == best && i < best_i))) {
      best = v;
      best_i = i;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
    if (ov > best || (ov == best && oi < best_i)) {
      best = ov;
      best_i = oi;
    }
  }
  if (lane == 0) out[b] = (best_i == 0x7fffffff) ? (M1 - 1) : best_i;
}

inline unsigned cap_grid(long long work_items, int per_cta) {
  long long g = (work_items + per_cta - 1) / per_cta;
  const long long cap = static_cast<long long>(ruart_num_sms()) * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<unsigned>(g);
}

}  // namespace

extern "C" int ruart_gather_rows(const float* src, long long src_pitch, const void* src_idx,
                                 float* dst, long long dst_pitch, const void* dst_idx, float* dst2,
                                 long long dst2_pitch, long long n, int D, int idx_is_64,
                                 void* stream) {
  RUART_ARG_CHECK(src != nullptr && dst != nullptr && D > 0 && n >= 0);
  if (n == 0) return RUART_OK;
  const unsigned grid = cap_grid(n * ((D + 3) / 4), 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (idx_is_64)
    gather_rows_kernel<long long><<<grid, 256, 0, st>>>(src, src_pitch, (const long long*)src_idx,
                                                        dst, dst_pitch, (const long long*)dst_idx,
                                                        dst2, dst2_pitch, n, D);
  else
    gather_rows_kernel<int32_t><<<grid, 256, 0, st>>>(src, src_pitch, (const int32_t*)src_idx, dst,
                                                      dst_pitch, (const int32_t*)dst_idx, dst2,
                                                      dst2_pitch, n, D);
  RUART_LAUNCH_CHECK();
  return RUART_OK;
}

extern "C" int ruart_whole_layernorm_stats(float* x, long long rows, int cols, long long pitch, float eps,
                                           double* workspace, float* stats_out, void* stream);

extern "C" int ruart_whole_layernorm(float* x, long long rows, int cols, long long pitch, float eps,
                                     double* workspace, void* stream) {
  return ruart_whole_layernorm_stats(x, rows, cols, pitch, eps, workspace, nullptr, stream);
}

// Same, also writing (mean, rstd) to stats_out[2] (device) — what the backward pass needs.
extern "C" int ruart_whole_layernorm_stats(float* x, long long rows, int cols, long long pitch, float eps,
                                           double* workspace, float* stats_out, void* stream) {
  RUART_ARG_CHECK(x != nullptr && workspace != nullptr && rows > 0 && cols > 0);
  cudaStream_t st = (cudaStream_t)stream;
  long long g = (rows * cols + LN_THREADS * 8 - 1) / (LN_THREADS * 8);
  if (g > LN_MAX_PARTS) g = LN_MAX_PARTS;
  if (g < 1) g = 1;
  whole_ln_stats_kernel<<<static_cast<unsigned>(g), LN_THREADS, 0, st>>>(x, rows, cols, pitch,
                                                                          workspace);
  RUART_LAUNCH_CHECK();
  whole_ln_apply_kernel<<<cap_grid(rows * cols, LN_THREADS * 4), LN_THREADS, 0, st>>>(
      x, rows, cols, pitch, workspace, static_cast<int>(g), eps, stats_out);
  RUART_LAUNCH_CHECK();
  return RUART_OK;
}

extern "C" int ruart_attention_tail(const float* p1, long long p1_pitch, const float* p2,
                                    long long p2_pitch, int hidden, const uint8_t* mask,
                                    const float* x3, long long x3_pitch, int D3, float* out,
                                    long long out_pitch, int B, int L1, int L2, int add_to_out,
                                    int split_parts, void* stream) {
  RUART_ARG_CHECK(B > 0 && L1 > 0 && L2 > 0 && hidden > 0 && D3 > 0);
  static const bool no_mma = getenv("RUART_TAIL_NO_MMA") != nullptr;  // A/B aid
  if (split_parts == 2 && L2 <= ATM_L2 && !no_mma) {
    dim3 grid((L1 + ATM_Q - 1) / ATM_Q, B);
    TailX3 xs = {{x3, nullptr, nullptr, nullptr}};
    attention_tail_mma_kernel<<<grid, ATM_THREADS, 0, (cudaStream_t)stream>>>(
        p1, p1_pitch, p2, p2_pitch, hidden, mask, xs, x3_pitch, D3, out, out_pitch, L1, L2,
        add_to_out);
    RUART_LAUNCH_CHECK();
    return RUART_OK;
  }
  if (L2 <= AT_L2) {
    dim3 grid((L1 + AT_Q - 1) / AT_Q, B);
    attention_tail_tiled_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        p1, p1_pitch, p2, p2_pitch, hidden, mask, x3, x3_pitch, D3, out, out_pitch, L1, L2,
        add_to_out);
    RUART_LAUNCH_CHECK();
    return RUART_OK;
  }
  const size_t smem = (static_cast<size_t>(ATT_QT) * hidden + static_cast<size_t>(ATT_QT) * L2) *
                      sizeof(float);
  RUART_ARG_CHECK(smem <= 200 * 1024);
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    RUART_CUDA_CHECK(cudaFuncSetAttribute(attention_tail_kernel,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    smem_set = 200 * 1024;
  }
  dim3 grid((L1 + ATT_QT - 1) / ATT_QT, B);
  attention_tail_kernel<<<grid, ATT_THREADS, smem, (cudaStream_t)stream>>>(
      p1, p1_pitch, p2, p2_pitch, hidden, mask, x3, x3_pitch, D3, out, out_pitch, L1, L2,
      add_to_out);
  RUART_LAUNCH_CHECK();
  return RUART_OK;
}

// The n_heads (<= 4) attention heads of one DeepAttention call in ONE launch (grid.z = head): head z takes the
// columns [z hidden, (z + 1) hidden) of the stacked projections p1 / p2, x3s[z] (all [B, L2, D3], pitch x3_pitch)
// and writes out[:, :, z D3 : (z + 1) D3].  Tensor-core form only (2-part operands, L2 <= 128): RUART_ERR_ARG
// otherwise and the caller launches the heads one by one.
extern "C" int ruart_attention_tail_heads(const float* p1, long long p1_pitch, const float* p2,
                                          long long p2_pitch, int hidden, int n_heads, const uint8_t* mask,
                                          const float* const* x3s_host, long long x3_pitch, int D3, float* out,
                                          long long out_pitch, int B, int L1, int L2, void* stream) {
  RUART_ARG_CHECK(B > 0 && L1 > 0 && L2 > 0 && hidden > 0 && D3 > 0 && n_heads >= 1 && n_heads <= 4);
  RUART_ARG_CHECK(L2 <= ATM_L2 && x3s_host != nullptr);
  TailX3 xs = {{nullptr, nullptr, nullptr, nullptr}};
  for (int i = 0; i < n_heads; ++i) {
    RUART_ARG_CHECK(x3s_host[i] != nullptr);
    xs.p[i] = x3s_host[i];
  }
  dim3 grid((L1 + ATM_Q - 1) / ATM_Q, B, n_heads);
  attention_tail_mma_kernel<<<grid, ATM_THREADS, 0, (cudaStream_t)stream>>>(
      p1, p1_pitch, p2, p2_pitch, hidden, mask, xs, x3_pitch, D3, out, out_pitch, L1, L2, 0);
  RUART_LAUNCH_CHECK();
  return RUART_OK;
}

extern "C" int ruart_self_attn_pool(const float* x, long long x_pitch, int B, int L, int D,
                                    const uint8_t* mask, const float* w, const float* bias,
                                    float* out, long long out_pitch, void* stream) {
  RUART_ARG_CHECK(B > 0 && L > 0 && D > 0 && L * sizeof(float) <= 48 * 1024);
  self_attn_pool_kernel<<<B, 256, L * sizeof(float), (cudaStream_t)stream>>>(x, x_pitch, L, D, mask,
                                                                            w, bias, out, out_pitch);
  RUART_LAUNCH_CHECK();
  return RUART_OK;
}

extern "C" int ruart_final_scores(const float* x, long long x_pitch, int B, int M, int X,
                                  const float* wy, const uint8_t* mask, int es_len,
                                  const float* noans_w, const float* noans_b, float* probs,
                                  float* logits, int* nan_flag, void* stream) {
  RUART_ARG_CHECK(B > 0 && M > 0 && X > 0 && es_len >= 0 && es_len <= M);
  const size_t smem = (2 * static_cast<size_t>(M) + 1 + 8) * sizeof(float);
  RUART_ARG_CHECK(smem <= 48 * 1024);
  final_scores_kernel<<<B, 256, smem, (cudaStream_t)stream>>>(x, x_pitch, M, X, wy, mask, es_len,
                                                              noans_w, noans_b, probs, logits,
                                                              nan_flag);
  RUART_LAUNCH_CHECK();
  return RUART_OK;
}

extern "C" int ruart_lstm_cell(const float* gx, const int32_t* row_gx, const float* gh, float* c,
                               void* h_split, int parts, int Kp, int H, int n_rows,
                               const int32_t* last_step, int step, const long long* slot_off,
                               float* slots, void* stream) {
  RUART_ARG_CHECK(H > 0 && Kp >= H && (Kp % 64) == 0 && parts >= 1 && parts <= 3);
  if (n_rows == 0) return RUART_OK;
  lstm_cell_kernel<<<cap_grid(static_cast<long long>(n_rows) * H, 256), 256, 0,
                     (cudaStream_t)stream>>>(gx, row_gx, gh, c, (__nv_bfloat16*)h_split, parts,
                                             Kp, H, n_rows, last_step, step, slot_off, slots);
  RUART_LAUNCH_CHECK();
  return RUART_OK;
}

extern "C" int ruart_select_answers(const float* probs, const int32_t* num_cnt, int B, int M1,
                                    int label_no_answer, int32_t* out_idx, void* stream) {
  RUART_ARG_CHECK(B > 0 && M1 > 0 && probs != nullptr && num_cnt != nullptr && out_idx != nullptr);
  select_answers_kernel<<<(B + 7) / 8, 256, 0, (cudaStream_t)stream>>>(probs, num_cnt, B, M1,
                                                                      label_no_answer, out_idx);
  RUART_LAUNCH_CHECK();
  return RUART_OK;
}
