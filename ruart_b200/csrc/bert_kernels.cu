// Memory-bound BERT kernels for sm_100a over PACKED (pad-free) token rows.
//
// The reference runs BertModel on padded [N, L] id matrices (Models/Bert/modeling.py:585-614).
// Pad keys receive an additive -10000 (modeling.py:604) whose exp underflows to exactly 0 in
// fp32, and pad query rows are never read downstream (Models/Bert/Bert.py:153-165 only reads
// [st,ed) of real words), so we keep only the real tokens: row t of every activation matrix is
// one real wordpiece, sequences are delimited by cu_seqlens.  See DESIGN.md §"Pad skipping".
//
// Activations are either bf16 ("bf16 mode") or fp32 + a 3-part bf16 split for the next GEMM
// ("fp32 mode"); every kernel takes both kinds of pointers and uses the non-null ones.
//
// One warp owns one 768- (or 1024-) wide row: lane l holds columns c*256 + l*8 .. +7 of chunk c,
// i.e. 16-byte vector accesses for bf16 and 2 x 16 bytes for fp32, fully coalesced.
#include <cstdlib>
#include "common.cuh"
#include "ruart_b200.h"

namespace {

using namespace ruart;

constexpr int ROWS_PER_CTA = 8;  // 8 warps

template <int HC>
struct RowVec {
  float v[HC * 8];
};

template <int HC>
__device__ __forceinline__ void load_row_f32(const float* __restrict__ p, int lane, RowVec<HC>& r) {
#pragma unroll
  for (int c = 0; c < HC; ++c) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p + c * 256 + lane * 8));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p + c * 256 + lane * 8 + 4));
    r.v[c * 8 + 0] = a.x; r.v[c * 8 + 1] = a.y; r.v[c * 8 + 2] = a.z; r.v[c * 8 + 3] = a.w;
    r.v[c * 8 + 4] = b.x; r.v[c * 8 + 5] = b.y; r.v[c * 8 + 6] = b.z; r.v[c * 8 + 7] = b.w;
  }
}
template <int HC>
__device__ __forceinline__ void load_row_bf16(const __nv_bfloat16* __restrict__ p, int lane,
                                              RowVec<HC>& r) {
#pragma unroll
  for (int c = 0; c < HC; ++c) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p + c * 256 + lane * 8));
    r.v[c * 8 + 0] = bf16_lo(u.x); r.v[c * 8 + 1] = bf16_hi(u.x);
    r.v[c * 8 + 2] = bf16_lo(u.y); r.v[c * 8 + 3] = bf16_hi(u.y);
    r.v[c * 8 + 4] = bf16_lo(u.z); r.v[c * 8 + 5] = bf16_hi(u.z);
    r.v[c * 8 + 6] = bf16_lo(u.w); r.v[c * 8 + 7] = bf16_hi(u.w);
  }
}
// Activation row from whichever representation exists.
template <int HC>
__device__ __forceinline__ void load_act(const float* f32, const __nv_bfloat16* b16, long long row,
                                         int lane, RowVec<HC>& r) {
  constexpr int H = HC * 256;
  if (f32 != nullptr) load_row_f32<HC>(f32 + row * H, lane, r);
  else load_row_bf16<HC>(b16 + row * H, lane, r);
}
// Store a row: fp32 (if out_f32) and bf16 in `parts` split parts (part p at column p*H).
template <int HC>
__device__ __forceinline__ void store_act(float* out_f32, __nv_bfloat16* out_b16, int parts,
                                          long long row, int lane, RowVec<HC>& r) {
  constexpr int H = HC * 256;
  if (out_f32 != nullptr) {
    float* p = out_f32 + row * H;
#pragma unroll
    for (int c = 0; c < HC; ++c) {
      *reinterpret_cast<float4*>(p + c * 256 + lane * 8) =
          make_float4(r.v[c * 8 + 0], r.v[c * 8 + 1], r.v[c * 8 + 2], r.v[c * 8 + 3]);
      *reinterpret_cast<float4*>(p + c * 256 + lane * 8 + 4) =
          make_float4(r.v[c * 8 + 4], r.v[c * 8 + 5], r.v[c * 8 + 6], r.v[c * 8 + 7]);
    }
  }
  if (out_b16 != nullptr) {
    for (int part = 0; part < parts; ++part) {
      __nv_bfloat16* p = out_b16 + row * (static_cast<long long>(parts) * H) + part * H;
#pragma unroll
      for (int c = 0; c < HC; ++c) {
        uint32_t pk[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const __nv_bfloat16 h0 = __float2bfloat16_rn(r.v[c * 8 + 2 * q]);
          const __nv_bfloat16 h1 = __float2bfloat16_rn(r.v[c * 8 + 2 * q + 1]);
          pk[q] = static_cast<uint32_t>(__bfloat16_as_ushort(h0)) |
                  (static_cast<uint32_t>(__bfloat16_as_ushort(h1)) << 16);
          r.v[c * 8 + 2 * q] -= __bfloat162float(h0);
          r.v[c * 8 + 2 * q + 1] -= __bfloat162float(h1);
        }
        *reinterpret_cast<uint4*>(p + c * 256 + lane * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
    }
  }
}

// BertLayerNorm (modeling.py:155-168): u = mean, s = mean((x-u)^2), (x-u)/sqrt(s+eps)*gamma+beta
template <int HC>
__device__ __forceinline__ void layer_norm_row(RowVec<HC>& r, const float* __restrict__ gamma,
                                               const float* __restrict__ beta, float eps, int lane) {
  constexpr int H = HC * 256;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < HC * 8; ++i) s += r.v[i];
  const float u = warp_sum(s) * (1.0f / H);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < HC * 8; ++i) {
    const float d = r.v[i] - u;
    q = fmaf(d, d, q);
  }
  const float var = warp_sum(q) * (1.0f / H);
  const float inv = 1.0f / sqrtf(var + eps);
#pragma unroll
  for (int c = 0; c < HC; ++c) {
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c * 256 + lane * 8));
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + c * 256 + lane * 8 + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c * 256 + lane * 8));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + c * 256 + lane * 8 + 4));
    const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) r.v[c * 8 + i] = fmaf(g[i], (r.v[c * 8 + i] - u) * inv, b[i]);
  }
}

// ---------------------------------------------------------------------------------------
// BertEmbeddings (modeling.py:185-199): word[id] + position[pos] + token_type[0] -> LayerNorm
template <int HC>
__global__ void __launch_bounds__(ROWS_PER_CTA * 32)
bert_embed_ln_kernel(const int32_t* __restrict__ ids, const int32_t* __restrict__ pos,
                     const float* __restrict__ word_emb, const float* __restrict__ pos_emb,
                     const float* __restrict__ type_emb, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float eps, int T, float* out_f32,
                     __nv_bfloat16* out_b16, int parts) {
  constexpr int H = HC * 256;
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * ROWS_PER_CTA + (threadIdx.x >> 5);
  if (row >= T) return;
  const int id = ids[row];
  const int ps = pos[row];
  RowVec<HC> w, p, t;
  load_row_f32<HC>(word_emb + static_cast<long long>(id) * H, lane, w);
  load_row_f32<HC>(pos_emb + static_cast<long long>(ps) * H, lane, p);
  load_row_f32<HC>(type_emb, lane, t);
#pragma unroll
  for (int i = 0; i < HC * 8; ++i) w.v[i] = (w.v[i] + p.v[i]) + t.v[i];
  layer_norm_row<HC>(w, gamma, beta, eps, lane);
  store_act<HC>(out_f32, out_b16, parts, row, lane, w);
}

// Folded-LayerNorm form (bf16 encoder, see gemm_tcgen05.cu): the embedding sum is stored BEFORE its LayerNorm
// (bf16) together with the row's (sum, sum of squares) in slot 0 of its 8-slot partial-sum row; the first
// layer's query/key/value GEMM and the first residual add finish the normalisation.
template <int HC>
__global__ void __launch_bounds__(ROWS_PER_CTA * 32)
bert_embed_raw_kernel(const int32_t* __restrict__ ids, const int32_t* __restrict__ pos,
                      const float* __restrict__ word_emb, const float* __restrict__ pos_emb,
                      const float* __restrict__ type_emb, int T, __nv_bfloat16* __restrict__ out_b16,
                      float2* __restrict__ out_stats) {
  constexpr int H = HC * 256;
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * ROWS_PER_CTA + (threadIdx.x >> 5);
  if (row >= T) return;
  const int id = ids[row];
  const int ps = pos[row];
  RowVec<HC> w, p, t;
  load_row_f32<HC>(word_emb + static_cast<long long>(id) * H, lane, w);
  load_row_f32<HC>(pos_emb + static_cast<long long>(ps) * H, lane, p);
  load_row_f32<HC>(type_emb, lane, t);
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < HC * 8; ++i) {
    w.v[i] = (w.v[i] + p.v[i]) + t.v[i];
    s1 += w.v[i];
    s2 = fmaf(w.v[i], w.v[i], s2);
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if (lane < 8) out_stats[row * 8 + lane] = (lane == 0) ? make_float2(s1, s2) : make_float2(0.f, 0.f);
  __nv_bfloat16* o = out_b16 + row * H;
#pragma unroll
  for (int c = 0; c < HC; ++c) {
    uint4 u;
    u.x = pack_bf16x2(w.v[c * 8 + 0], w.v[c * 8 + 1]);
    u.y = pack_bf16x2(w.v[c * 8 + 2], w.v[c * 8 + 3]);
    u.z = pack_bf16x2(w.v[c * 8 + 4], w.v[c * 8 + 5]);
    u.w = pack_bf16x2(w.v[c * 8 + 6], w.v[c * 8 + 7]);
    *reinterpret_cast<uint4*>(o + c * 256 + lane * 8) = u;
  }
}

// BertSelfOutput / BertOutput tail (modeling.py:260-264, 299-303): LayerNorm(dense_out + input)
// (the dense bias is already added by the GEMM epilogue).
template <int HC>
__global__ void __launch_bounds__(ROWS_PER_CTA * 32)
add_ln_kernel(const float* x_f32, const __nv_bfloat16* x_b16, const float* res_f32,
              const __nv_bfloat16* res_b16, const float* __restrict__ gamma,
              const float* __restrict__ beta, float eps, int T, float* out_f32,
              __nv_bfloat16* out_b16, int parts) {
  // One warp per row, one CTA per 8 rows.  (Round 2 tried a persistent grid-stride form with the next row's
  // loads issued before the current row is reduced: 77 registers, 3 CTAs / SM, and SLOWER — 79.9 us vs 72.5 us
  // for 113 664 rows, gpurun_out/r02_ln.txt — because fewer resident warps hide the shuffle reductions worse.)
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * ROWS_PER_CTA + (threadIdx.x >> 5);
  if (row >= T) return;
  RowVec<HC> x;
  load_act<HC>(x_f32, x_b16, row, lane, x);
  if (res_f32 != nullptr || res_b16 != nullptr) {  // absent: the GEMM epilogue already added it
    RowVec<HC> r;
    load_act<HC>(res_f32, res_b16, row, lane, r);
#pragma unroll
    for (int i = 0; i < HC * 8; ++i) x.v[i] += r.v[i];
  }
  layer_norm_row<HC>(x, gamma, beta, eps, lane);
  store_act<HC>(out_f32, out_b16, parts, row, lane, x);
}

// bf16 -> bf16 specialisation of add_ln_kernel (residual already fused into the GEMM epilogue, one split part):
// the form every LayerNorm of the bf16 encoder takes.  Fewer live registers -> more resident warps.
template <int HC>
__global__ void __launch_bounds__(ROWS_PER_CTA * 32, 6)
ln_bf16_kernel(const __nv_bfloat16* __restrict__ x_b16, const float* __restrict__ gamma,
               const float* __restrict__ beta, float eps, int T, __nv_bfloat16* __restrict__ out_b16) {
  constexpr int H = HC * 256;
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * ROWS_PER_CTA + (threadIdx.x >> 5);
  if (row >= T) return;
  RowVec<HC> x;
  load_row_bf16<HC>(x_b16 + row * H, lane, x);
  layer_norm_row<HC>(x, gamma, beta, eps, lane);
  __nv_bfloat16* o = out_b16 + row * H;
#pragma unroll
  for (int c = 0; c < HC; ++c) {
    uint4 u;
    u.x = pack_bf16x2(x.v[c * 8 + 0], x.v[c * 8 + 1]);
    u.y = pack_bf16x2(x.v[c * 8 + 2], x.v[c * 8 + 3]);
    u.z = pack_bf16x2(x.v[c * 8 + 4], x.v[c * 8 + 5]);
    u.w = pack_bf16x2(x.v[c * 8 + 6], x.v[c * 8 + 7]);
    *reinterpret_cast<uint4*>(o + c * 256 + lane * 8) = u;
  }
}

// ---------------------------------------------------------------------------------------
// BertSelfAttention core (modeling.py:229-250) for packed sequences, head_dim 64.
// One warp per (sequence, head).  K and V of the sequence's head are staged in shared memory
// (bf16, 16-byte chunks XOR-swizzled by row) when the sequence has <= kMaxStage tokens; longer
// sequences stream K/V from global memory (L2) block by block.  Online softmax over blocks of 32
// keys; lane j scores key j, lane l accumulates output dims (2l, 2l+1).
constexpr int ATT_WARPS = 8;
constexpr int ATT_MAX_STAGE = 64;  // most tokens whose K,V the per-warp staging buffer may hold

template <typename T>
__device__ __forceinline__ float2 ld2(const T* p);
template <>
__device__ __forceinline__ float2 ld2<float>(const float* p) {
  return *reinterpret_cast<const float2*>(p);
}
template <>
__device__ __forceinline__ float2 ld2<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint32_t u = *reinterpret_cast<const uint32_t*>(p);
  return make_float2(bf16_lo(u), bf16_hi(u));
}

template <typename T>
__global__ void __launch_bounds__(ATT_WARPS * 32)
bert_attention_kernel(const T* __restrict__ qkv, const int32_t* __restrict__ cu_seqlens, int n_seq,
                      int n_heads, float scale, int stage_tokens, int skip_upto, float* out_f32,
                      __nv_bfloat16* out_b16, int parts) {
  // per warp: K and V [stage_tokens][64] fp32 + the scaled query row [64]
  extern __shared__ float att_smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  float* sK = att_smem + warp * (2 * stage_tokens * 64 + 64);
  float* sV = sK + stage_tokens * 64;
  float* sQ = sV + stage_tokens * 64;
  const long long task = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + warp;
  if (task >= static_cast<long long>(n_seq) * n_heads) return;
  const int seq = static_cast<int>(task / n_heads);
  const int head = static_cast<int>(task - static_cast<long long>(seq) * n_heads);
  const int t0 = cu_seqlens[seq];
  const int len = cu_seqlens[seq + 1] - t0;
  if (len <= skip_upto) return;  // handled by bert_attention_short_kernel
  const int H = n_heads * 64;
  const long long ld = 3LL * H;
  const T* qbase = qkv + static_cast<long long>(t0) * ld + head * 64;
  const T* kbase = qbase + H;
  const T* vbase = qbase + 2 * H;
  const bool staged = len <= stage_tokens;
  if (staged) {
    // row j, float index d stored at j*64 + (d ^ ((j & 7) << 2)) : conflict-free for "lane = row"
    for (int j = 0; j < len; ++j) {
      const float2 kk = ld2<T>(kbase + j * ld + 2 * lane);
      const float2 vv = ld2<T>(vbase + j * ld + 2 * lane);
      const int sw = (2 * lane) ^ ((j & 7) << 2);
      *reinterpret_cast<float2*>(sK + j * 64 + sw) = kk;
      *reinterpret_cast<float2*>(sV + j * 64 + 2 * lane) = vv;
    }
  }
  __syncwarp();
  for (int i = 0; i < len; ++i) {
    const float2 qq = ld2<T>(qbase + i * ld + 2 * lane);
    __syncwarp();
    *reinterpret_cast<float2*>(sQ + 2 * lane) = make_float2(qq.x * scale, qq.y * scale);
    __syncwarp();
    float m = -INFINITY, l = 0.f, o0 = 0.f, o1 = 0.f;
    for (int kb = 0; kb < len; kb += 32) {
      const int j = kb + lane;
      float s = -INFINITY;
      if (j < len) {
        float acc = 0.f;
        if (staged) {
          const float* kr = sK + j * 64;
          const int x = (j & 7) << 2;
#pragma unroll
          for (int d = 0; d < 64; d += 4) {
            const float4 kv = *reinterpret_cast<const float4*>(kr + (d ^ x));
            const float4 qv = *reinterpret_cast<const float4*>(sQ + d);
            acc = fmaf(qv.x, kv.x, acc);
            acc = fmaf(qv.y, kv.y, acc);
            acc = fmaf(qv.z, kv.z, acc);
            acc = fmaf(qv.w, kv.w, acc);
          }
        } else {
          const T* kr = kbase + j * ld;
#pragma unroll 8
          for (int d = 0; d < 64; d += 2) {
            const float2 kv = ld2<T>(kr + d);
            acc = fmaf(sQ[d], kv.x, acc);
            acc = fmaf(sQ[d + 1], kv.y, acc);
          }
        }
        s = acc;
      }
      const float m_new = fmaxf(m, warp_max(s));
      const float corr = __expf(m - m_new);  // m = -inf on the first block -> 0
      const float pj = (j < len) ? __expf(s - m_new) : 0.f;
      l = l * corr + warp_sum(pj);
      o0 *= corr;
      o1 *= corr;
      const int nk = min(32, len - kb);
      for (int jj = 0; jj < nk; ++jj) {
        const float pb = __shfl_sync(0xffffffffu, pj, jj);
        float2 vv;
        if (staged) vv = *reinterpret_cast<const float2*>(sV + (kb + jj) * 64 + 2 * lane);
        else vv = ld2<T>(vbase + (kb + jj) * ld + 2 * lane);
        o0 = fmaf(pb, vv.x, o0);
        o1 = fmaf(pb, vv.y, o1);
      }
      m = m_new;
    }
    const float inv = 1.0f / l;
    o0 *= inv;
    o1 *= inv;
    const long long orow = static_cast<long long>(t0 + i);
    const int ocol = head * 64 + 2 * lane;
    if (out_f32 != nullptr)
      *reinterpret_cast<float2*>(out_f32 + orow * H + ocol) = make_float2(o0, o1);
    if (out_b16 != nullptr) {
      float r0 = o0, r1 = o1;
      for (int part = 0; part < parts; ++part) {
        const __nv_bfloat16 h0 = __float2bfloat16_rn(r0);
        const __nv_bfloat16 h1 = __float2bfloat16_rn(r1);
        __nv_bfloat162 hh;
        hh.x = h0;
        hh.y = h1;
        *reinterpret_cast<__nv_bfloat162*>(out_b16 + orow * (static_cast<long long>(parts) * H) +
                                           part * H + ocol) = hh;
        r0 -= __bfloat162float(h0);
        r1 -= __bfloat162float(h1);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// Short-sequence variant (bf16 qkv, len <= SHORT_MAX = 16; RUArt's OCR / object-label items are
// 3..10 wordpieces).  One warp per sequence, looping over heads; per head the [len x 64] Q, K, V
// slices are staged in shared memory (16-byte chunks XOR-swizzled by row, rows >= len zeroed) and
// the two small products run on the warp-level tensor-core path (mma.sync m16n8k16, bf16 in, fp32
// accumulate) in the FlashAttention-2 register layout:
//     S[16 x 16] = Q K^T   (4 k-steps x 2 key tiles)   -> masked softmax in registers (quad shuffles)
//     O[16 x 64] = P V     (8 dim tiles, P re-used from the S accumulators as the A fragment)
// tcgen05 cannot help here (its smallest M is 64 and operands must sit in 1024-byte swizzle
// atoms); these tiles are 5x64.  O is staged back through the Q tile and leaves as 16-byte rows.
constexpr int SHORT_MAX = 16;
constexpr int SHORT_WARPS = 8;

// Item sequences (<= 8 wordpieces, the bulk of RUArt's rows) go two at a time: consecutive
// sequences A and B are also consecutive token rows, so A fills rows 0-7 and B rows 8-15 of ONE
// 16-row MMA tile; the cross blocks of S are masked to -inf, which makes P block-diagonal and
// O = P V exact for both.  A sequence of 9..16 tokens takes the tile alone.  One warp per
// (sequence pair, head).
__global__ void __launch_bounds__(SHORT_WARPS * 32, 4)
bert_attention_mma16_kernel(const __nv_bfloat16* __restrict__ qkv,
                            const int32_t* __restrict__ cu_seqlens, int n_seq, int n_heads,
                            float scale, __nv_bfloat16* __restrict__ out) {
  constexpr int TB = 16 * 128;
  __shared__ __align__(128) uint8_t sh16[SHORT_WARPS][3 * TB];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* sQ = sh16[warp];
  uint8_t* sK = sQ + TB;
  uint8_t* sV = sK + TB;
  const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(sV);
  const int H = n_heads * 64;
  const long long ld = 3LL * H;
  const int g = lane >> 2, t = lane & 3;
  const int a_row = (lane & 7) + ((lane >> 3) & 1) * 8;
  const int a_chk = lane >> 4;
  const int b_row = lane & 7;
  const int b_chk = lane >> 3;
  const int n_pairs = (n_seq + 1) >> 1;
  // one warp per sequence pair, looping over the heads: the sequence bookkeeping (three dependent
  // cu_seqlens loads) is paid once per pair instead of once per (pair, head)
  const int h_split = (n_heads % 2 == 0) ? 2 : 1;  // two warps share a pair's heads (finer tasks)
  const int h_per = n_heads / h_split;
  for (int task = blockIdx.x * SHORT_WARPS + warp; task < n_pairs * h_split; task += gridDim.x * SHORT_WARPS) {
    const int pair = task / h_split;
    const int h_lo = (task - pair * h_split) * h_per;
    const int sA = 2 * pair, sB = sA + 1;
    const int tA = cu_seqlens[sA];
    const int tB = cu_seqlens[sA + 1];
    const int lenA = tB - tA;
    const int lenB = (sB < n_seq) ? cu_seqlens[sB + 1] - tB : 0;
   for (int h = h_lo; h < h_lo + h_per; ++h) {
    // passes: paired (both <= 8), or each sequence of <= 16 tokens on its own
    const bool paired = lenA <= 8 && lenB <= 8;
    const int n_pass = paired ? 1 : 2;
    for (int pass = 0; pass < n_pass; ++pass) {
      int t0, l0, l1;  // rows 0.. : l0 tokens from t0 ; rows 8.. : l1 tokens from t0 + l0 (paired)
      if (paired) {
        t0 = tA; l0 = lenA; l1 = lenB;
      } else {
        t0 = pass == 0 ? tA : tB;
        l0 = pass == 0 ? lenA : lenB;
        l1 = 0;
        if (l0 <= 0 || l0 > 16) continue;  // longer sequences belong to the other kernels
      }
      if (l0 + l1 == 0) continue;
      // row r of the tile <-> token t0 + tok(r); rows without a token are zero
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int idx = lane + 32 * i;
        const int r = idx >> 3, c = idx & 7;
        int tok = -1;
        if (paired) {
          if (r < 8) { if (r < l0) tok = r; }
          else if (r - 8 < l1) tok = l0 + (r - 8);
        } else if (r < l0) {
          tok = r;
        }
        uint4 q4 = make_uint4(0, 0, 0, 0), k4 = q4, v4 = q4;
        if (tok >= 0) {
          const uint4* row = reinterpret_cast<const uint4*>(qkv + (static_cast<long long>(t0 + tok)) * ld + h * 64);
          q4 = __ldg(row + c);
          k4 = __ldg(row + (H >> 3) + c);
          v4 = __ldg(row + 2 * (H >> 3) + c);
        }
        const int off = tile_off(r, c);
        *reinterpret_cast<uint4*>(sQ + off) = q4;
        *reinterpret_cast<uint4*>(sK + off) = k4;
        *reinterpret_cast<uint4*>(sV + off) = v4;
      }
      __syncwarp();
      const bool two = paired ? (l1 > 0) : (l0 > 8);  // is the second key tile in use?
      float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
      uint32_t kb0[8], kb1[8];
      ldsm_x4(aK + tile_off(b_row, b_chk), kb0[0], kb0[1], kb0[2], kb0[3]);
      ldsm_x4(aK + tile_off(b_row, 4 + b_chk), kb0[4], kb0[5], kb0[6], kb0[7]);
      if (two) {
        ldsm_x4(aK + tile_off(8 + b_row, b_chk), kb1[0], kb1[1], kb1[2], kb1[3]);
        ldsm_x4(aK + tile_off(8 + b_row, 4 + b_chk), kb1[4], kb1[5], kb1[6], kb1[7]);
      }
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t a0, a1, a2, a3;
        ldsm_x4(aQ + tile_off(a_row, 2 * ks + a_chk), a0, a1, a2, a3);
        mma_bf16_16816(s0, a0, a1, a2, a3, kb0[2 * ks], kb0[2 * ks + 1]);
        if (two) mma_bf16_16816(s1, a0, a1, a2, a3, kb1[2 * ks], kb1[2 * ks + 1]);
      }
      // validity of (row, key): thread holds rows g and g+8, keys 2t, 2t+1 of tile 0 and of tile 1
      const int k0 = 2 * t, k1 = 2 * t + 1;
      bool vA0, vA1, vA2, vA3, vB0, vB1, vB2, vB3;  // A: row g ; B: row g+8 ; 0,1: tile 0 ; 2,3: tile 1
      if (paired) {
        vA0 = k0 < l0; vA1 = k1 < l0; vA2 = false; vA3 = false;
        vB0 = false; vB1 = false; vB2 = k0 < l1; vB3 = k1 < l1;
      } else {
        vA0 = vB0 = k0 < l0; vA1 = vB1 = k1 < l0;
        vA2 = vB2 = 8 + k0 < l0; vA3 = vB3 = 8 + k1 < l0;
      }
      const float xA0 = vA0 ? s0[0] * scale : -INFINITY, xA1 = vA1 ? s0[1] * scale : -INFINITY;
      const float xA2 = vA2 ? s1[0] * scale : -INFINITY, xA3 = vA3 ? s1[1] * scale : -INFINITY;
      const float xB0 = vB0 ? s0[2] * scale : -INFINITY, xB1 = vB1 ? s0[3] * scale : -INFINITY;
      const float xB2 = vB2 ? s1[2] * scale : -INFINITY, xB3 = vB3 ? s1[3] * scale : -INFINITY;
      float mA = fmaxf(fmaxf(xA0, xA1), fmaxf(xA2, xA3));
      float mB = fmaxf(fmaxf(xB0, xB1), fmaxf(xB2, xB3));
      mA = fmaxf(mA, __shfl_xor_sync(0xffffffffu, mA, 1));
      mA = fmaxf(mA, __shfl_xor_sync(0xffffffffu, mA, 2));
      mB = fmaxf(mB, __shfl_xor_sync(0xffffffffu, mB, 1));
      mB = fmaxf(mB, __shfl_xor_sync(0xffffffffu, mB, 2));
      // rows without any valid key (pad rows of the tile) would give exp(-inf + inf): clamp the max
      mA = (mA == -INFINITY) ? 0.f : mA;
      mB = (mB == -INFINITY) ? 0.f : mB;
      const float pA0 = __expf(xA0 - mA), pA1 = __expf(xA1 - mA), pA2 = __expf(xA2 - mA), pA3 = __expf(xA3 - mA);
      const float pB0 = __expf(xB0 - mB), pB1 = __expf(xB1 - mB), pB2 = __expf(xB2 - mB), pB3 = __expf(xB3 - mB);
      float lA = (pA0 + pA1) + (pA2 + pA3), lB = (pB0 + pB1) + (pB2 + pB3);
      lA += __shfl_xor_sync(0xffffffffu, lA, 1);
      lA += __shfl_xor_sync(0xffffffffu, lA, 2);
      lB += __shfl_xor_sync(0xffffffffu, lB, 1);
      lB += __shfl_xor_sync(0xffffffffu, lB, 2);
      const uint32_t p0 = pack_bf16x2(pA0, pA1), p1 = pack_bf16x2(pB0, pB1);
      const uint32_t p2 = pack_bf16x2(pA2, pA3), p3 = pack_bf16x2(pB2, pB3);
      const float iA = lA > 0.f ? 1.0f / lA : 0.f, iB = lB > 0.f ? 1.0f / lB : 0.f;
      __syncwarp();  // all lanes are done reading sQ: it becomes the O staging tile
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(aV + tile_off(a_row, 2 * dp + a_chk), b0, b1, b2, b3);
        float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f};
        mma_bf16_16816(o0, p0, p1, p2, p3, b0, b1);
        mma_bf16_16816(o1, p0, p1, p2, p3, b2, b3);
        *reinterpret_cast<uint32_t*>(sQ + tile_off(g, 2 * dp) + 4 * t) = pack_bf16x2(o0[0] * iA, o0[1] * iA);
        *reinterpret_cast<uint32_t*>(sQ + tile_off(g + 8, 2 * dp) + 4 * t) = pack_bf16x2(o0[2] * iB, o0[3] * iB);
        *reinterpret_cast<uint32_t*>(sQ + tile_off(g, 2 * dp + 1) + 4 * t) = pack_bf16x2(o1[0] * iA, o1[1] * iA);
        *reinterpret_cast<uint32_t*>(sQ + tile_off(g + 8, 2 * dp + 1) + 4 * t) = pack_bf16x2(o1[2] * iB, o1[3] * iB);
      }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int idx = lane + 32 * i;
        const int r = idx >> 3, c = idx & 7;
        int tok = -1;
        if (paired) {
          if (r < 8) { if (r < l0) tok = r; }
          else if (r - 8 < l1) tok = l0 + (r - 8);
        } else if (r < l0) {
          tok = r;
        }
        if (tok >= 0) {
          const uint4 o4 = *reinterpret_cast<const uint4*>(sQ + tile_off(r, c));
          *(reinterpret_cast<uint4*>(out + (static_cast<long long>(t0 + tok)) * H + h * 64) + c) = o4;
        }
      }
    }
   }
  }
}

// Same tiles and math, but the Q/K/V slices of head h+1 are fetched with cp.async (LDGSTS, 16 bytes
// per lane, zero-filled for rows without a token) into a second set of tiles while head h is being
// computed: the kernel is bound by the DRAM round trip of each (pair, head) stage, not by bytes or
// issue slots, so the prefetch hides one of the two phases.  2 x 6 KB of tiles per warp.
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__global__ void __launch_bounds__(SHORT_WARPS * 32, 2)
bert_attention_mma16_async_kernel(const __nv_bfloat16* __restrict__ qkv,
                                  const int32_t* __restrict__ cu_seqlens, int n_seq, int n_heads,
                                  float scale, int h_split, __nv_bfloat16* __restrict__ out) {
  constexpr int TB = 16 * 128;
  extern __shared__ __align__(128) uint8_t sh_dyn[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* sbase = sh_dyn + static_cast<size_t>(warp) * 6 * TB;  // [2 stages][Q | K | V]
  const uint32_t abase = smem_u32(sbase);
  const int H = n_heads * 64;
  const long long ld = 3LL * H;
  const int g = lane >> 2, t = lane & 3;
  const int a_row = (lane & 7) + ((lane >> 3) & 1) * 8;
  const int a_chk = lane >> 4;
  const int b_row = lane & 7;
  const int b_chk = lane >> 3;
  const int n_pairs = (n_seq + 1) >> 1;
  const int h_per = n_heads / h_split;
  const int cch = lane & 7;
  for (int task = blockIdx.x * SHORT_WARPS + warp; task < n_pairs * h_split; task += gridDim.x * SHORT_WARPS) {
    const int pair = task / h_split;
    const int h_lo = (task - pair * h_split) * h_per;
    const int sA = 2 * pair, sB = sA + 1;
    const int tA = cu_seqlens[sA];
    const int tB = cu_seqlens[sA + 1];
    const int lenA = tB - tA;
    const int lenB = (sB < n_seq) ? cu_seqlens[sB + 1] - tB : 0;
    const bool paired = lenA <= 8 && lenB <= 8;
    const int n_pass = paired ? 1 : 2;
    for (int pass = 0; pass < n_pass; ++pass) {
      int t0, l0, l1;  // rows 0.. : l0 tokens from t0 ; rows 8.. : l1 tokens from t0 + l0 (paired)
      if (paired) {
        t0 = tA; l0 = lenA; l1 = lenB;
      } else {
        t0 = pass == 0 ? tA : tB;
        l0 = pass == 0 ? lenA : lenB;
        l1 = 0;
        if (l0 <= 0 || l0 > 16) continue;
      }
      if (l0 + l1 == 0) continue;
      auto tok_of = [&](int r) {
        if (paired) {
          if (r < 8) return r < l0 ? r : -1;
          return (r - 8 < l1) ? l0 + (r - 8) : -1;
        }
        return r < l0 ? r : -1;
      };
      auto issue = [&](int h, int stage) {
        const uint32_t st_base = abase + stage * 3 * TB;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = (lane >> 3) + 4 * i;
          const int tok = tok_of(r);
          const __nv_bfloat16* row = qkv + static_cast<long long>(t0 + (tok >= 0 ? tok : 0)) * ld + h * 64 + cch * 8;
          const uint32_t off = st_base + tile_off(r, cch);
          cp_async16_zfill(off, row, tok >= 0);
          cp_async16_zfill(off + TB, row + H, tok >= 0);
          cp_async16_zfill(off + 2 * TB, row + 2 * H, tok >= 0);
        }
        cp_async_commit();
      };
      const bool two = paired ? (l1 > 0) : (l0 > 8);
      const int k0 = 2 * t, k1 = 2 * t + 1;
      bool vA0, vA1, vA2, vA3, vB0, vB1, vB2, vB3;
      if (paired) {
        vA0 = k0 < l0; vA1 = k1 < l0; vA2 = false; vA3 = false;
        vB0 = false; vB1 = false; vB2 = k0 < l1; vB3 = k1 < l1;
      } else {
        vA0 = vB0 = k0 < l0; vA1 = vB1 = k1 < l0;
        vA2 = vB2 = 8 + k0 < l0; vA3 = vB3 = 8 + k1 < l0;
      }
      __syncwarp();  // the previous task's tiles are no longer read
      issue(h_lo, 0);
      int stage = 0;
      for (int h = h_lo; h < h_lo + h_per; ++h, stage ^= 1) {
        if (h + 1 < h_lo + h_per) {
          issue(h + 1, stage ^ 1);
          cp_async_wait<1>();
        } else {
          cp_async_wait<0>();
        }
        __syncwarp();
        uint8_t* sQ = sbase + stage * 3 * TB;
        const uint32_t aQ = abase + stage * 3 * TB, aK = aQ + TB, aV = aK + TB;
        float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t kb0[8], kb1[8];
        ldsm_x4(aK + tile_off(b_row, b_chk), kb0[0], kb0[1], kb0[2], kb0[3]);
        ldsm_x4(aK + tile_off(b_row, 4 + b_chk), kb0[4], kb0[5], kb0[6], kb0[7]);
        if (two) {
          ldsm_x4(aK + tile_off(8 + b_row, b_chk), kb1[0], kb1[1], kb1[2], kb1[3]);
          ldsm_x4(aK + tile_off(8 + b_row, 4 + b_chk), kb1[4], kb1[5], kb1[6], kb1[7]);
        }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint32_t a0, a1, a2, a3;
          ldsm_x4(aQ + tile_off(a_row, 2 * ks + a_chk), a0, a1, a2, a3);
          mma_bf16_16816(s0, a0, a1, a2, a3, kb0[2 * ks], kb0[2 * ks + 1]);
          if (two) mma_bf16_16816(s1, a0, a1, a2, a3, kb1[2 * ks], kb1[2 * ks + 1]);
        }
        const float xA0 = vA0 ? s0[0] * scale : -INFINITY, xA1 = vA1 ? s0[1] * scale : -INFINITY;
        const float xA2 = vA2 ? s1[0] * scale : -INFINITY, xA3 = vA3 ? s1[1] * scale : -INFINITY;
        const float xB0 = vB0 ? s0[2] * scale : -INFINITY, xB1 = vB1 ? s0[3] * scale : -INFINITY;
        const float xB2 = vB2 ? s1[2] * scale : -INFINITY, xB3 = vB3 ? s1[3] * scale : -INFINITY;
        float mA = fmaxf(fmaxf(xA0, xA1), fmaxf(xA2, xA3));
        float mB = fmaxf(fmaxf(xB0, xB1), fmaxf(xB2, xB3));
        mA = fmaxf(mA, __shfl_xor_sync(0xffffffffu, mA, 1));
        mA = fmaxf(mA, __shfl_xor_sync(0xffffffffu, mA, 2));
        mB = fmaxf(mB, __shfl_xor_sync(0xffffffffu, mB, 1));
        mB = fmaxf(mB, __shfl_xor_sync(0xffffffffu, mB, 2));
        mA = (mA == -INFINITY) ? 0.f : mA;
        mB = (mB == -INFINITY) ? 0.f : mB;
        const float pA0 = __expf(xA0 - mA), pA1 = __expf(xA1 - mA), pA2 = __expf(xA2 - mA), pA3 = __expf(xA3 - mA);
        const float pB0 = __expf(xB0 - mB), pB1 = __expf(xB1 - mB), pB2 = __expf(xB2 - mB), pB3 = __expf(xB3 - mB);
        float lA = (pA0 + pA1) + (pA2 + pA3), lB = (pB0 + pB1) + (pB2 + pB3);
        lA += __shfl_xor_sync(0xffffffffu, lA, 1);
        lA += __shfl_xor_sync(0xffffffffu, lA, 2);
        lB += __shfl_xor_sync(0xffffffffu, lB, 1);
        lB += __shfl_xor_sync(0xffffffffu, lB, 2);
        const uint32_t p0 = pack_bf16x2(pA0, pA1), p1 = pack_bf16x2(pB0, pB1);
        const uint32_t p2 = pack_bf16x2(pA2, pA3), p3 = pack_bf16x2(pB2, pB3);
        const float iA = lA > 0.f ? 1.0f / lA : 0.f, iB = lB > 0.f ? 1.0f / lB : 0.f;
        __syncwarp();  // all lanes are done reading sQ: it becomes the O staging tile
#pragma unroll
        for (int dp = 0; dp < 4; ++dp) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4_t(aV + tile_off(a_row, 2 * dp + a_chk), b0, b1, b2, b3);
          float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f};
          mma_bf16_16816(o0, p0, p1, p2, p3, b0, b1);
          mma_bf16_16816(o1, p0, p1, p2, p3, b2, b3);
          *reinterpret_cast<uint32_t*>(sQ + tile_off(g, 2 * dp) + 4 * t) = pack_bf16x2(o0[0] * iA, o0[1] * iA);
          *reinterpret_cast<uint32_t*>(sQ + tile_off(g + 8, 2 * dp) + 4 * t) = pack_bf16x2(o0[2] * iB, o0[3] * iB);
          *reinterpret_cast<uint32_t*>(sQ + tile_off(g, 2 * dp + 1) + 4 * t) = pack_bf16x2(o1[0] * iA, o1[1] * iA);
          *reinterpret_cast<uint32_t*>(sQ + tile_off(g + 8, 2 * dp + 1) + 4 * t) = pack_bf16x2(o1[2] * iB, o1[3] * iB);
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = (lane >> 3) + 4 * i;
          const int tok = tok_of(r);
          if (tok >= 0) {
            const uint4 o4 = *reinterpret_cast<const uint4*>(sQ + tile_off(r, cch));
            *(reinterpret_cast<uint4*>(out + (static_cast<long long>(t0 + tok)) * H + h * 64) + cch) = o4;
          }
        }
        __syncwarp();  // the tiles of this stage are free for the prefetch issued next iteration
      }
    }
  }
}

// MAXT = 16 or 64: rows of the staged Q/K/V tiles.  One warp per (sequence, head) task with
// lo < len <= MAXT; other lengths are left to the other instantiation / the streaming kernel.
template <int MAXT>
__global__ void __launch_bounds__(SHORT_WARPS * 32)
bert_attention_mma_kernel(const __nv_bfloat16* __restrict__ qkv,
                          const int32_t* __restrict__ cu_seqlens, int n_seq, int n_heads,
                          float scale, int lo, __nv_bfloat16* __restrict__ out) {
  constexpr int TB = MAXT * 128;        // bytes of one [MAXT x 64] bf16 tile
  constexpr int NKT = MAXT / 8;         // 8-key tiles
  constexpr int NQT = MAXT / 16;        // 16-query tiles
  extern __shared__ __align__(128) uint8_t sh_dyn[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* sQ = sh_dyn + static_cast<size_t>(warp) * 3 * TB;
  uint8_t* sK = sQ + TB;
  uint8_t* sV = sK + TB;
  const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(sV);
  const int H = n_heads * 64;
  const long long ld = 3LL * H;
  const int g = lane >> 2, t = lane & 3;
  const int a_row = (lane & 7) + ((lane >> 3) & 1) * 8;
  const int a_chk = lane >> 4;
  const int b_row = lane & 7;
  const int b_chk = lane >> 3;
  const long long n_tasks = static_cast<long long>(n_seq) * n_heads;
  for (long long task = static_cast<long long>(blockIdx.x) * SHORT_WARPS + warp; task < n_tasks;
       task += static_cast<long long>(gridDim.x) * SHORT_WARPS) {
    const int seq = static_cast<int>(task / n_heads);
    const int h = static_cast<int>(task - static_cast<long long>(seq) * n_heads);
    const int t0 = cu_seqlens[seq];
    const int len = cu_seqlens[seq + 1] - t0;
    if (len <= lo || len > MAXT) continue;
    const int n_chunks = len * 8;
    const int n_qt = (len + 15) >> 4, n_kt = (len + 7) >> 3, n_kk = (len + 15) >> 4;
    const int rows_used = n_qt * 16;  // rows touched by the MMAs (pad rows zeroed)
    __syncwarp();
    // all chunks of the three tiles are requested at once with cp.async (zero-filled past the last
    // token): one DRAM round trip per task instead of one per 32-chunk loop iteration
    for (int idx = lane; idx < rows_used * 8; idx += 32) {
      const int r = idx >> 3, c = idx & 7;
      const bool valid = idx < n_chunks;
      const __nv_bfloat16* row = qkv + (static_cast<long long>(t0 + (valid ? r : 0))) * ld + h * 64 + c * 8;
      const uint32_t off = tile_off(r, c);
      cp_async16_zfill(aQ + off, row, valid);
      cp_async16_zfill(aK + off, row + H, valid);
      cp_async16_zfill(aV + off, row + 2 * H, valid);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncwarp();
    for (int qt = 0; qt < n_qt; ++qt) {
      uint32_t qa[4][4];
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        ldsm_x4(aQ + tile_off(qt * 16 + a_row, 2 * ks + a_chk), qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3]);
      // k-step outermost: the NKT accumulator chains are independent, so consecutive mma.sync
      // never wait on each other (the in-order issue would otherwise stall ~30 clk per MMA)
      float s[NKT][4];
#pragma unroll
      for (int nt = 0; nt < NKT; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
      for (int kh = 0; kh < 2; ++kh) {  // dims 0-31 / 32-63: one ldmatrix.x4 per key tile
        uint32_t kb[NKT][4];
#pragma unroll
        for (int nt = 0; nt < NKT; ++nt)
          if (nt < n_kt)
            ldsm_x4(aK + tile_off(nt * 8 + b_row, 4 * kh + b_chk), kb[nt][0], kb[nt][1], kb[nt][2], kb[nt][3]);
#pragma unroll
        for (int k2 = 0; k2 < 2; ++k2) {
          const int ks = 2 * kh + k2;
#pragma unroll
          for (int nt = 0; nt < NKT; ++nt)
            if (nt < n_kt)
              mma_bf16_16816(s[nt], qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3], kb[nt][2 * k2], kb[nt][2 * k2 + 1]);
        }
      }
      // masked softmax: rows g (regs 0,1) and g+8 (regs 2,3); keys nt*8 + 2t, +1
      float mA = -INFINITY, mB = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < NKT; ++nt) {
        const int k0 = nt * 8 + 2 * t;
        s[nt][0] = (k0 < len) ? s[nt][0] * scale : -INFINITY;
        s[nt][1] = (k0 + 1 < len) ? s[nt][1] * scale : -INFINITY;
        s[nt][2] = (k0 < len) ? s[nt][2] * scale : -INFINITY;
        s[nt][3] = (k0 + 1 < len) ? s[nt][3] * scale : -INFINITY;
        mA = fmaxf(mA, fmaxf(s[nt][0], s[nt][1]));
        mB = fmaxf(mB, fmaxf(s[nt][2], s[nt][3]));
      }
      mA = fmaxf(mA, __shfl_xor_sync(0xffffffffu, mA, 1));
      mA = fmaxf(mA, __shfl_xor_sync(0xffffffffu, mA, 2));
      mB = fmaxf(mB, __shfl_xor_sync(0xffffffffu, mB, 1));
      mB = fmaxf(mB, __shfl_xor_sync(0xffffffffu, mB, 2));
      float lA = 0.f, lB = 0.f;
#pragma unroll
      for (int nt = 0; nt < NKT; ++nt) {
        s[nt][0] = __expf(s[nt][0] - mA);
        s[nt][1] = __expf(s[nt][1] - mA);
        s[nt][2] = __expf(s[nt][2] - mB);
        s[nt][3] = __expf(s[nt][3] - mB);
        lA += s[nt][0] + s[nt][1];
        lB += s[nt][2] + s[nt][3];
      }
      lA += __shfl_xor_sync(0xffffffffu, lA, 1);
      lA += __shfl_xor_sync(0xffffffffu, lA, 2);
      lB += __shfl_xor_sync(0xffffffffu, lB, 1);
      lB += __shfl_xor_sync(0xffffffffu, lB, 2);
      const float iA = 1.0f / lA, iB = 1.0f / lB;
      // P fragments per 16-key step
      uint32_t pa[NQT][4];
#pragma unroll
      for (int kk = 0; kk < NQT; ++kk) {
        pa[kk][0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
        pa[kk][1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
        pa[kk][2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        pa[kk][3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
      }
      __syncwarp();  // Q fragments of this query tile are in registers: its rows become O staging
      float o[8][4];  // eight 8-wide dim tiles: independent accumulator chains, key step outermost
#pragma unroll
      for (int d = 0; d < 8; ++d) o[d][0] = o[d][1] = o[d][2] = o[d][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < NQT; ++kk) {
        if (kk < n_kk) {
#pragma unroll
          for (int dp = 0; dp < 4; ++dp) {
            uint32_t b0, b1, b2, b3;
            ldsm_x4_t(aV + tile_off(kk * 16 + a_row, 2 * dp + a_chk), b0, b1, b2, b3);
            mma_bf16_16816(o[2 * dp], pa[kk][0], pa[kk][1], pa[kk][2], pa[kk][3], b0, b1);
            mma_bf16_16816(o[2 * dp + 1], pa[kk][0], pa[kk][1], pa[kk][2], pa[kk][3], b2, b3);
          }
        }
      }
      const int rA = qt * 16 + g, rB = rA + 8;
#pragma unroll
      for (int d = 0; d < 8; ++d) {
        *reinterpret_cast<uint32_t*>(sQ + tile_off(rA, d) + 4 * t) = pack_bf16x2(o[d][0] * iA, o[d][1] * iA);
        *reinterpret_cast<uint32_t*>(sQ + tile_off(rB, d) + 4 * t) = pack_bf16x2(o[d][2] * iB, o[d][3] * iB);
      }
    }
    __syncwarp();
    for (int idx = lane; idx < n_chunks; idx += 32) {
      const int r = idx >> 3, c = idx & 7;
      const uint4 o4 = *reinterpret_cast<const uint4*>(sQ + tile_off(r, c));
      *(reinterpret_cast<uint4*>(out + (static_cast<long long>(t0 + r)) * H + h * 64) + c) = o4;
    }
  }
}

// ---------------------------------------------------------------------------------------
// Subword -> word averaging (Models/Bert/Bert.py:149-165) fused with the learned layer sum
// (Models/SDNet.py:573-583): for encoder layer `layer`,
//     dst[item, j] (+)= (mean_{t in [st,ed)} h[row_start[item] + t]) * softmax(alpha)[layer] * gamma
// One warp per word (item, j, st, ed).  Words whose x_mask[item, j] is 0 are skipped
// (Bert.py:155-156); ed == st+1 copies the row, ed > st+1 divides the sum by float(ed-st),
// ed <= st contributes zeros (Bert.py:160-165).  `first` overwrites instead of accumulating.
// alpha == nullptr: coefficient 1 (per-layer outputs of the plain Bert.forward API).
template <int HC>
__global__ void __launch_bounds__(ROWS_PER_CTA * 32)
subword_avg_accum_kernel(const float* h_f32, const __nv_bfloat16* h_b16,
                         const int32_t* __restrict__ words, int n_words,
                         const int32_t* __restrict__ row_start, const uint8_t* __restrict__ x_mask,
                         int W, float* __restrict__ dst, long long dst_stride,
                         const float* __restrict__ alpha, int n_layers,
                         const float* __restrict__ gamma_p, int layer, int first) {
  const int lane = threadIdx.x & 31;
  const long long w = static_cast<long long>(blockIdx.x) * ROWS_PER_CTA + (threadIdx.x >> 5);
  if (w >= n_words) return;
  const int item = words[w];
  const int j = words[n_words + w];
  const int st = words[2LL * n_words + w];
  const int ed = words[3LL * n_words + w];
  if (j >= W) return;
  if (x_mask != nullptr && x_mask[static_cast<long long>(item) * W + j] == 0) return;
  float a = 1.0f, g = 1.0f;
  if (alpha != nullptr) {
    // softmax(alpha)[layer] (F.softmax(alpha, dim=0), SDNet.py:574)
    float mx = -INFINITY;
    for (int i = 0; i < n_layers; ++i) mx = fmaxf(mx, alpha[i]);
    float den = 0.f;
    for (int i = 0; i < n_layers; ++i) den += expf(alpha[i] - mx);
    a = expf(alpha[layer] - mx) / den;
    g = gamma_p[0];
  }
  const int cnt = ed - st;
  const long long t0 = static_cast<long long>(row_start[item]) + st;
  RowVec<HC> acc;
#pragma unroll
  for (int i = 0; i < HC * 8; ++i) acc.v[i] = 0.f;
  for (int t = 0; t < cnt; ++t) {
    RowVec<HC> r;
    load_act<HC>(h_f32, h_b16, t0 + t, lane, r);
#pragma unroll
    for (int i = 0; i < HC * 8; ++i) acc.v[i] += r.v[i];
  }
  if (cnt > 1) {
    const float fc = static_cast<float>(cnt);
#pragma unroll
    for (int i = 0; i < HC * 8; ++i) acc.v[i] = acc.v[i] / fc;
  }
  float* d = dst + (static_cast<long long>(item) * W + j) * dst_stride;
#pragma unroll
  for (int c = 0; c < HC; ++c) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      float4* pp = reinterpret_cast<float4*>(d + c * 256 + lane * 8 + q * 4);
      float4 o = make_float4((acc.v[c * 8 + q * 4 + 0] * a) * g, (acc.v[c * 8 + q * 4 + 1] * a) * g,
                             (acc.v[c * 8 + q * 4 + 2] * a) * g, (acc.v[c * 8 + q * 4 + 3] * a) * g);
      if (!first) {
        const float4 old = *pp;
        o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
      }
      *pp = o;
    }
  }
}

// All-layers form of the kernel above: the encoder keeps its n_layers outputs ([n_layers][T][H],
// `layer_stride` elements apart) and ONE pass per word computes
//     dst[item, j] = sum_l ( mean_{t in [st,ed)} h_l[t] * softmax(alpha)[l] ) * gamma
// in layer order (res = t_0; res += t_1; ... as SDNet.py:575-580), writing dst once.
template <int HC>
__global__ void __launch_bounds__(ROWS_PER_CTA * 32)
subword_avg_layers_kernel(const float* h_f32, const __nv_bfloat16* h_b16, long long layer_stride,
                          const int32_t* __restrict__ words, int n_words,
                          const int32_t* __restrict__ row_start, const uint8_t* __restrict__ x_mask,
                          int W, float* __restrict__ dst, long long dst_stride,
                          const float* __restrict__ alpha, int n_layers,
                          const float* __restrict__ gamma_p) {
  const int lane = threadIdx.x & 31;
  const long long w = static_cast<long long>(blockIdx.x) * ROWS_PER_CTA + (threadIdx.x >> 5);
  if (w >= n_words) return;
  const int item = words[w];
  const int j = words[n_words + w];
  const int st = words[2LL * n_words + w];
  const int ed = words[3LL * n_words + w];
  if (j >= W) return;
  if (x_mask != nullptr && x_mask[static_cast<long long>(item) * W + j] == 0) {
    // masked word: the reference leaves zeros (Bert.py:155-156); written explicitly so that the
    // destination buffer needs no zero-fill
    float* dz = dst + (static_cast<long long>(item) * W + j) * dst_stride;
#pragma unroll
    for (int c = 0; c < HC; ++c) {
      *reinterpret_cast<float4*>(dz + c * 256 + lane * 8) = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(dz + c * 256 + lane * 8 + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    return;
  }
  float mx = -INFINITY;
  for (int i = 0; i < n_layers; ++i) mx = fmaxf(mx, alpha[i]);
  float den = 0.f;
  for (int i = 0; i < n_layers; ++i) den += expf(alpha[i] - mx);
  const float g = gamma_p[0];
  const int cnt = ed - st;
  const long long t0 = static_cast<long long>(row_start[item]) + st;
  const float fc = static_cast<float>(cnt);
  RowVec<HC> tot;
#pragma unroll
  for (int i = 0; i < HC * 8; ++i) tot.v[i] = 0.f;
  for (int l = 0; l < n_layers; ++l) {
    const float a = expf(alpha[l] - mx) / den;
    const float* hf = h_f32 ? h_f32 + l * layer_stride : nullptr;
    const __nv_bfloat16* hb = h_b16 ? h_b16 + l * layer_stride : nullptr;
    RowVec<HC> acc;
#pragma unroll
    for (int i = 0; i < HC * 8; ++i) acc.v[i] = 0.f;
    for (int t = 0; t < cnt; ++t) {
      RowVec<HC> r;
      load_act<HC>(hf, hb, t0 + t, lane, r);
#pragma unroll
      for (int i = 0; i < HC * 8; ++i) acc.v[i] += r.v[i];
    }
#pragma unroll
    for (int i = 0; i < HC * 8; ++i) {
      const float mean = (cnt > 1) ? acc.v[i] / fc : acc.v[i];
      const float term = (mean * a) * g;
      tot.v[i] = (l == 0) ? term : tot.v[i] + term;
    }
  }
  float* d = dst + (static_cast<long long>(item) * W + j) * dst_stride;
#pragma unroll
  for (int c = 0; c < HC; ++c) {
    *reinterpret_cast<float4*>(d + c * 256 + lane * 8) =
        make_float4(tot.v[c * 8 + 0], tot.v[c * 8 + 1], tot.v[c * 8 + 2], tot.v[c * 8 + 3]);
    *reinterpret_cast<float4*>(d + c * 256 + lane * 8 + 4) =
        make_float4(tot.v[c * 8 + 4], tot.v[c * 8 + 5], tot.v[c * 8 + 6], tot.v[c * 8 + 7]);
  }
}

// Pipelined form of subword_avg_layers_kernel for bf16 hidden states (the kernel that runs in bf16 mode).
// The synchronous kernel above keeps one layer's rows in flight per warp (1.5 - 4.5 KB) and sits at
// 0.33 of the HBM bandwidth (latency-, not bandwidth-bound).  Here every warp owns a
// 24 KB shared-memory ring of row slots and fetches the word's rows of SEVERAL layers ahead with cp.async
// (LDGSTS, 16 bytes per lane, the lane that copies a chunk is the lane that reads it back): a one-piece
// word has all 12 layers (18 KB) in flight at once, a two-piece word 8 layers (24 KB).  Registers no longer
// bound the bytes in flight.  Words with more than SW_MAX_CNT pieces take the synchronous loop.
constexpr int SW_WARPS = 16;
constexpr int SW_MAX_CNT = 4;
template <int HC>
struct SwRing {
  static constexpr int ROW_BYTES = HC * 512;               // one bf16 row
  // 12 KB per warp, 16 warps per CTA (192 KB): the first version (8 warps x 24 KB) had the same bytes in flight but
  // only 12.5 % of the warp slots, and every word is a serial issue -> wait -> reduce -> store chain (ncu:
  // profiles/r02_ncu_full2_summary.txt, DRAM 48 %); twice the warps overlap those chains
  static constexpr int SLOTS = (HC == 3) ? 8 : 6;
  static constexpr int WARP_BYTES = SLOTS * ROW_BYTES;
};

template <int HC, int CNT>
__device__ __forceinline__ void subword_word_async(const __nv_bfloat16* __restrict__ h, long long layer_stride,
                                                   long long t0, int n_layers, const float* __restrict__ alpha,
                                                   float mx, float den, float g, uint32_t ring, int lane,
                                                   RowVec<HC>& tot) {
  constexpr int H = HC * 256;
  constexpr int RB = SwRing<HC>::ROW_BYTES;
  constexpr int S = SwRing<HC>::SLOTS / CNT;               // layers in flight
  auto issue = [&](int l) {
    const __nv_bfloat16* src = h + static_cast<long long>(l) * layer_stride + t0 * H + lane * 8;
    const uint32_t dst = ring + static_cast<uint32_t>((l % S) * CNT) * RB + lane * 16;
#pragma unroll
    for (int t = 0; t < CNT; ++t)
#pragma unroll
      for (int c = 0; c < HC; ++c)
        cp_async16_zfill(dst + t * RB + c * 512, src + static_cast<long long>(t) * H + c * 256, true);
  };
#pragma unroll 1
  for (int l = 0; l < S; ++l) {
    if (l < n_layers) issue(l);
    cp_async_commit();
  }
  const float fc = static_cast<float>(CNT);
#pragma unroll 1
  for (int l = 0; l < n_layers; ++l) {
    cp_async_wait<S - 1>();                                // the group of layer l has landed
    const float a = expf(alpha[l] - mx) / den;
    RowVec<HC> acc;
#pragma unroll
    for (int i = 0; i < HC * 8; ++i) acc.v[i] = 0.f;
    const uint32_t base = ring + static_cast<uint32_t>((l % S) * CNT) * RB + lane * 16;
#pragma unroll
    for (int t = 0; t < CNT; ++t) {
#pragma unroll
      for (int c = 0; c < HC; ++c) {
        uint4 u;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w)
                     : "r"(base + t * RB + c * 512));
        acc.v[c * 8 + 0] += bf16_lo(u.x); acc.v[c * 8 + 1] += bf16_hi(u.x);
        acc.v[c * 8 + 2] += bf16_lo(u.y); acc.v[c * 8 + 3] += bf16_hi(u.y);
        acc.v[c * 8 + 4] += bf16_lo(u.z); acc.v[c * 8 + 5] += bf16_hi(u.z);
        acc.v[c * 8 + 6] += bf16_lo(u.w); acc.v[c * 8 + 7] += bf16_hi(u.w);
      }
    }
#pragma unroll
    for (int i = 0; i < HC * 8; ++i) {
      const float mean = (CNT > 1) ? acc.v[i] / fc : acc.v[i];
      const float term = (mean * a) * g;
      tot.v[i] = (l == 0) ? term : tot.v[i] + term;
    }
    // the slot just consumed is the slot of layer l + S (same lane wrote and read it: no warp sync needed)
    if (l + S < n_layers) issue(l + S);
    cp_async_commit();
  }
  cp_async_wait<0>();
}

template <int HC>
__global__ void __launch_bounds__(SW_WARPS * 32, 1)
subword_avg_layers_async_kernel(const __nv_bfloat16* __restrict__ h_b16, long long layer_stride,
                                const int32_t* __restrict__ words, int n_words,
                                const int32_t* __restrict__ row_start, const uint8_t* __restrict__ x_mask,
                                int W, float* __restrict__ dst, long long dst_stride,
                                const float* __restrict__ alpha, int n_layers,
                                const float* __restrict__ gamma_p) {
  extern __shared__ __align__(16) uint8_t sw_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t ring = smem_u32(sw_smem) + static_cast<uint32_t>(warp) * SwRing<HC>::WARP_BYTES;
  float mx = -INFINITY;
  for (int i = 0; i < n_layers; ++i) mx = fmaxf(mx, alpha[i]);
  float den = 0.f;
  for (int i = 0; i < n_layers; ++i) den += expf(alpha[i] - mx);
  const float g = gamma_p[0];
  for (long long w = static_cast<long long>(blockIdx.x) * SW_WARPS + warp; w < n_words;
       w += static_cast<long long>(gridDim.x) * SW_WARPS) {
    const int item = words[w];
    const int j = words[n_words + w];
    const int st = words[2LL * n_words + w];
    const int ed = words[3LL * n_words + w];
    if (j >= W) continue;
    float* d = dst + (static_cast<long long>(item) * W + j) * dst_stride;
    RowVec<HC> tot;
#pragma unroll
    for (int i = 0; i < HC * 8; ++i) tot.v[i] = 0.f;
    const int cnt = ed - st;
    // masked word / st >= ed: zeros (Bert.py:155-156,160-165), written explicitly (no zero-fill of dst)
    if (!(x_mask != nullptr && x_mask[static_cast<long long>(item) * W + j] == 0) && cnt > 0) {
      const long long t0 = static_cast<long long>(row_start[item]) + st;
      switch (cnt) {
        case 1: subword_word_async<HC, 1>(h_b16, layer_stride, t0, n_layers, alpha, mx, den, g, ring, lane, tot); break;
        case 2: subword_word_async<HC, 2>(h_b16, layer_stride, t0, n_layers, alpha, mx, den, g, ring, lane, tot); break;
        case 3: subword_word_async<HC, 3>(h_b16, layer_stride, t0, n_layers, alpha, mx, den, g, ring, lane, tot); break;
        case 4: subword_word_async<HC, 4>(h_b16, layer_stride, t0, n_layers, alpha, mx, den, g, ring, lane, tot); break;
        default: {
          const float fc = static_cast<float>(cnt);
          for (int l = 0; l < n_layers; ++l) {
            const float a = expf(alpha[l] - mx) / den;
            RowVec<HC> acc;
#pragma unroll
            for (int i = 0; i < HC * 8; ++i) acc.v[i] = 0.f;
            for (int t = 0; t < cnt; ++t) {
              RowVec<HC> r;
              load_row_bf16<HC>(h_b16 + l * layer_stride + (t0 + t) * (HC * 256), lane, r);
#pragma unroll
              for (int i = 0; i < HC * 8; ++i) acc.v[i] += r.v[i];
            }
#pragma unroll
            for (int i = 0; i < HC * 8; ++i) {
              const float term = ((acc.v[i] / fc) * a) * g;
              tot.v[i] = (l == 0) ? term : tot.v[i] + term;
            }
          }
        }
      }
    }
#pragma unroll
    for (int c = 0; c < HC; ++c) {
      *reinterpret_cast<float4*>(d + c * 256 + lane * 8) =
          make_float4(tot.v[c * 8 + 0], tot.v[c * 8 + 1], tot.v[c * 8 + 2], tot.v[c * 8 + 3]);
      *reinterpret_cast<float4*>(d + c * 256 + lane * 8 + 4) =
          make_float4(tot.v[c * 8 + 4], tot.v[c * 8 + 5], tot.v[c * 8 + 6], tot.v[c * 8 + 7]);
    }
  }
}

// ---------------------------------------------------------------------------------------
// Folded-LayerNorm form of the kernel above (bf16 encoder with fused LayerNorms, gemm_tcgen05.cu): the kept
// layer outputs are the rows BEFORE each layer's output LayerNorm plus their partial sums, so the normalisation
// happens here.  With a_l = softmax(alpha)_l * gamma and y = (v r - mu r) g_l + b_l (BertLayerNorm, modeling.py:155-168)
//     dst = sum_l a_l * mean_pieces(y)  =  sum_l G_l * mean_pieces(v r - mu r)  +  C,
//     G_l[c] = a_l g_l[c]   (table [n_layers][H], built by subword_coef_kernel),   C[c] = sum_l a_l b_l[c]
// i.e. ONE fma per element for the row's (r, -mu r) — the same count as the plain sum — and one fma per layer
// and column for G.  G lives in shared memory (36 KB for 12 x 768), every ring slot carries the row's 64 bytes
// of partial sums behind its 1536 bytes of data (fetched by the same cp.async group).
constexpr int SWF_WARPS = 16;
template <int HC>
struct SwFoldRing {
  static constexpr int DATA_BYTES = HC * 512;
  static constexpr int SLOT_BYTES = DATA_BYTES + 64;          // + 8 float2 partial sums of the row
  static constexpr int SLOTS = 7;
  static constexpr int WARP_BYTES = SLOTS * SLOT_BYTES;
};

__global__ void __launch_bounds__(256)
subword_coef_kernel(const float* __restrict__ alpha, const float* __restrict__ gamma_p, int n_layers,
                    const float* __restrict__ ln_g, const float* __restrict__ ln_b, int H,
                    float* __restrict__ G, float* __restrict__ C) {
  // ln_g / ln_b: [n_layers][H] output-LayerNorm weights of the encoder layers
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= H) return;
  float mx = -INFINITY;
  for (int i = 0; i < n_layers; ++i) mx = fmaxf(mx, alpha[i]);
  float den = 0.f;
  for (int i = 0; i < n_layers; ++i) den += expf(alpha[i] - mx);
  const float g = gamma_p[0];
  float cc = 0.f;
  for (int l = 0; l < n_layers; ++l) {
    const float a = (expf(alpha[l] - mx) / den) * g;
    G[static_cast<long long>(l) * H + c] = a * ln_g[static_cast<long long>(l) * H + c];
    cc = fmaf(a, ln_b[static_cast<long long>(l) * H + c], cc);
  }
  C[c] = cc;
}

__device__ __forceinline__ void row_norm_from_partials(const float4 a, const float4 b, const float4 c, float inv_dim,
                                                       float eps, float& r, float& nmr) {
  const float s1 = ((a.x + a.z) + (b.x + b.z)) + (c.x + c.z);
  const float s2 = ((a.y + a.w) + (b.y + b.w)) + (c.y + c.w);
  const float mu = s1 * inv_dim;
  const float var = fmaxf(fmaf(-mu, mu, s2 * inv_dim), 0.0f);
  r = rsqrtf(var + eps);
  nmr = -mu * r;
}

template <int HC, int CNT>
__device__ __forceinline__ void subword_word_fold(const __nv_bfloat16* __restrict__ h, long long layer_stride,
                                                  const float2* __restrict__ stats, long long stats_layer_stride,
                                                  long long t0, int n_layers, const float* __restrict__ Gs,
                                                  float eps, uint32_t ring, int lane, RowVec<HC>& tot) {
  constexpr int H = HC * 256;
  constexpr int SB = SwFoldRing<HC>::SLOT_BYTES;
  constexpr int DB = SwFoldRing<HC>::DATA_BYTES;
  constexpr int S = SwFoldRing<HC>::SLOTS / CNT;             // layers in flight
  auto issue = [&](int l) {
    const __nv_bfloat16* src = h + static_cast<long long>(l) * layer_stride + t0 * H + lane * 8;
    const float2* ssrc = stats + static_cast<long long>(l) * stats_layer_stride + t0 * 8 + lane * 2;
    const uint32_t dst = ring + static_cast<uint32_t>((l % S) * CNT) * SB;
#pragma unroll
    for (int t = 0; t < CNT; ++t) {
#pragma unroll
      for (int c = 0; c < HC; ++c)
        cp_async16_zfill(dst + t * SB + c * 512 + lane * 16, src + static_cast<long long>(t) * H + c * 256, true);
      // lanes 0-3 also fetch the row's 64 bytes of partial sums; every lane reads them back, hence the
      // __syncwarp() after the wait below
      if (lane < 4) cp_async16_zfill(dst + t * SB + DB + lane * 16, ssrc + static_cast<long long>(t) * 8, true);
    }
  };
#pragma unroll 1
  for (int l = 0; l < S; ++l) {
    if (l < n_layers) issue(l);
    cp_async_commit();
  }
  const float inv_cnt = 1.0f / static_cast<float>(CNT);
#pragma unroll 1
  for (int l = 0; l < n_layers; ++l) {
    cp_async_wait<S - 1>();                                  // the group of layer l has landed (this lane's copies)
    __syncwarp();                                            // ... and lanes 0-3's copies of the partial sums
    RowVec<HC> acc;
#pragma unroll
    for (int i = 0; i < HC * 8; ++i) acc.v[i] = 0.f;
    const uint32_t base = ring + static_cast<uint32_t>((l % S) * CNT) * SB;
#pragma unroll
    for (int t = 0; t < CNT; ++t) {
      float4 pa, pb, pc;
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(pa.x), "=f"(pa.y), "=f"(pa.z), "=f"(pa.w)
                   : "r"(base + t * SB + DB));
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(pb.x), "=f"(pb.y), "=f"(pb.z), "=f"(pb.w)
                   : "r"(base + t * SB + DB + 16));
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(pc.x), "=f"(pc.y), "=f"(pc.z), "=f"(pc.w)
                   : "r"(base + t * SB + DB + 32));
      float r, nmr;
      row_norm_from_partials(pa, pb, pc, 1.0f / H, eps, r, nmr);
#pragma unroll
      for (int c = 0; c < HC; ++c) {
        uint4 u;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w)
                     : "r"(base + t * SB + c * 512 + lane * 16));
        acc.v[c * 8 + 0] += fmaf(bf16_lo(u.x), r, nmr); acc.v[c * 8 + 1] += fmaf(bf16_hi(u.x), r, nmr);
        acc.v[c * 8 + 2] += fmaf(bf16_lo(u.y), r, nmr); acc.v[c * 8 + 3] += fmaf(bf16_hi(u.y), r, nmr);
        acc.v[c * 8 + 4] += fmaf(bf16_lo(u.z), r, nmr); acc.v[c * 8 + 5] += fmaf(bf16_hi(u.z), r, nmr);
        acc.v[c * 8 + 6] += fmaf(bf16_lo(u.w), r, nmr); acc.v[c * 8 + 7] += fmaf(bf16_hi(u.w), r, nmr);
      }
    }
    __syncwarp();                                            // all lanes have read the partial sums of this slot
    const float* gl = Gs + l * H;
#pragma unroll
    for (int c = 0; c < HC; ++c) {
      const float4 g0 = *reinterpret_cast<const float4*>(gl + c * 256 + lane * 8);
      const float4 g1 = *reinterpret_cast<const float4*>(gl + c * 256 + lane * 8 + 4);
      const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float mean = (CNT > 1) ? acc.v[c * 8 + i] * inv_cnt : acc.v[c * 8 + i];
        tot.v[c * 8 + i] = fmaf(mean, g[i], tot.v[c * 8 + i]);
      }
    }
    if (l + S < n_layers) issue(l + S);
    cp_async_commit();
  }
  cp_async_wait<0>();
}

template <int HC>
__global__ void __launch_bounds__(SWF_WARPS * 32, 1)
subword_avg_layers_fold_kernel(const __nv_bfloat16* __restrict__ h_b16, long long layer_stride,
                               const float2* __restrict__ stats, long long stats_layer_stride, float eps,
                               const int32_t* __restrict__ words, int n_words,
                               const int32_t* __restrict__ row_start, const uint8_t* __restrict__ x_mask,
                               int W, float* __restrict__ dst, long long dst_stride,
                               const float* __restrict__ G, const float* __restrict__ C, int n_layers) {
  constexpr int H = HC * 256;
  extern __shared__ __align__(16) uint8_t sw_smem[];
  float* Gs = reinterpret_cast<float*>(sw_smem);             // [n_layers][H]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < n_layers * H / 4; i += blockDim.x)
    reinterpret_cast<float4*>(Gs)[i] = __ldg(reinterpret_cast<const float4*>(G) + i);
  __syncthreads();
  const uint32_t ring = smem_u32(sw_smem) + static_cast<uint32_t>(n_layers * H * 4) +
                        static_cast<uint32_t>(warp) * SwFoldRing<HC>::WARP_BYTES;
  for (long long w = static_cast<long long>(blockIdx.x) * SWF_WARPS + warp; w < n_words;
       w += static_cast<long long>(gridDim.x) * SWF_WARPS) {
    const int item = words[w];
    const int j = words[n_words + w];
    const int st = words[2LL * n_words + w];
    const int ed = words[3LL * n_words + w];
    if (j >= W) continue;
    float* d = dst + (static_cast<long long>(item) * W + j) * dst_stride;
    RowVec<HC> tot;
#pragma unroll
    for (int i = 0; i < HC * 8; ++i) tot.v[i] = 0.f;
    const int cnt = ed - st;
    // masked word / st >= ed: zeros (Bert.py:155-156,160-165), written explicitly (no zero-fill of dst)
    if (!(x_mask != nullptr && x_mask[static_cast<long long>(item) * W + j] == 0) && cnt > 0) {
      const long long t0 = static_cast<long long>(row_start[item]) + st;
#pragma unroll
      for (int c = 0; c < HC; ++c) {                         // C = sum_l a_l beta_l
        const float4 c0 = __ldg(reinterpret_cast<const float4*>(C + c * 256 + lane * 8));
        const float4 c1 = __ldg(reinterpret_cast<const float4*>(C + c * 256 + lane * 8 + 4));
        tot.v[c * 8 + 0] = c0.x; tot.v[c * 8 + 1] = c0.y; tot.v[c * 8 + 2] = c0.z; tot.v[c * 8 + 3] = c0.w;
        tot.v[c * 8 + 4] = c1.x; tot.v[c * 8 + 5] = c1.y; tot.v[c * 8 + 6] = c1.z; tot.v[c * 8 + 7] = c1.w;
      }
      switch (cnt) {
        case 1: subword_word_fold<HC, 1>(h_b16, layer_stride, stats, stats_layer_stride, t0, n_layers, Gs, eps, ring, lane, tot); break;
        case 2: subword_word_fold<HC, 2>(h_b16, layer_stride, stats, stats_layer_stride, t0, n_layers, Gs, eps, ring, lane, tot); break;
        case 3: subword_word_fold<HC, 3>(h_b16, layer_stride, stats, stats_layer_stride, t0, n_layers, Gs, eps, ring, lane, tot); break;
        default: {
          const float inv_cnt = 1.0f / static_cast<float>(cnt);
          for (int l = 0; l < n_layers; ++l) {
            RowVec<HC> acc;
#pragma unroll
            for (int i = 0; i < HC * 8; ++i) acc.v[i] = 0.f;
            for (int t = 0; t < cnt; ++t) {
              const float4* sp = reinterpret_cast<const float4*>(stats + l * stats_layer_stride + (t0 + t) * 8);
              float r, nmr;
              row_norm_from_partials(__ldg(sp), __ldg(sp + 1), __ldg(sp + 2), 1.0f / H, eps, r, nmr);
              RowVec<HC> x;
              load_row_bf16<HC>(h_b16 + l * layer_stride + (t0 + t) * H, lane, x);
#pragma unroll
              for (int i = 0; i < HC * 8; ++i) acc.v[i] += fmaf(x.v[i], r, nmr);
            }
            const float* gl = Gs + l * H;
#pragma unroll
            for (int c = 0; c < HC; ++c)
#pragma unroll
              for (int i = 0; i < 8; ++i)
                tot.v[c * 8 + i] = fmaf(acc.v[c * 8 + i] * inv_cnt, gl[c * 256 + lane * 8 + i], tot.v[c * 8 + i]);
          }
        }
      }
    }
#pragma unroll
    for (int c = 0; c < HC; ++c) {
      *reinterpret_cast<float4*>(d + c * 256 + lane * 8) =
          make_float4(tot.v[c * 8 + 0], tot.v[c * 8 + 1], tot.v[c * 8 + 2], tot.v[c * 8 + 3]);
      *reinterpret_cast<float4*>(d + c * 256 + lane * 8 + 4) =
          make_float4(tot.v[c * 8 + 4], tot.v[c * 8 + 5], tot.v[c * 8 + 6], tot.v[c * 8 + 7]);
    }
  }
}

// Sequence bookkeeping of the packed layout, on the device (replaces a handful of torch integer ops):
//   seq_lengths_kernel : one warp per row -> number of real tokens of the row and of each of its
//                        512-token windows (written at the row's / windows' global slots)
//   seq_scan_kernel    : ONE CTA: exclusive prefix sums cu_rows[R+1], cu_seq[S+1] and the totals the
//                        host needs to size buffers: totals[0] = T, totals[1 + k] = longest window of
//                        segment k
constexpr int MAX_SEGMENTS = 8;
struct SegTable {
  int n_seg;
  int row0[MAX_SEGMENTS + 1];  // first global row of segment k (row0[n_seg] = R)
  int seq0[MAX_SEGMENTS + 1];  // first global window-sequence of segment k
};

__global__ void __launch_bounds__(256)
seq_lengths_kernel(const uint8_t* __restrict__ mask, int N, int L, int window, int n_win,
                   int32_t* __restrict__ row_len, int32_t* __restrict__ win_len) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= N) return;
  int total = 0;
  for (int w = 0; w < n_win; ++w) {
    int cnt = 0;
    const int c_end = min(L, (w + 1) * window);
    for (int c = w * window + lane; c < c_end; c += 32)
      cnt += mask[static_cast<long long>(row) * L + c] != 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) win_len[static_cast<long long>(row) * n_win + w] = cnt;
    total += cnt;
  }
  if (lane == 0) row_len[row] = total;
}

__device__ void block_exclusive_scan(const int32_t* __restrict__ in, int n, int32_t* __restrict__ out) {
  // out[0] = 0, out[i+1] = in[0] + ... + in[i]; one CTA of 1024 threads, chunked
  __shared__ int s_part[1024];
  const int t = threadIdx.x, nt = blockDim.x;
  const int per = (n + nt - 1) / nt;
  const int lo = min(n, t * per), hi = min(n, lo + per);
  int sum = 0;
  for (int i = lo; i < hi; ++i) sum += in[i];
  s_part[t] = sum;
  __syncthreads();
  for (int off = 1; off < nt; off <<= 1) {  // Hillis-Steele inclusive scan of the partials
    const int v = (t >= off) ? s_part[t - off] : 0;
    __syncthreads();
    s_part[t] += v;
    __syncthreads();
  }
  int run = (t == 0) ? 0 : s_part[t - 1];
  if (t == 0) out[0] = 0;
  for (int i = lo; i < hi; ++i) {
    run += in[i];
    out[i + 1] = run;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(1024)
seq_scan_kernel(const int32_t* __restrict__ row_len, int R, const int32_t* __restrict__ win_len, int S,
                const SegTable seg, int32_t* __restrict__ cu_rows, int32_t* __restrict__ cu_seq,
                int32_t* __restrict__ totals, int max_total) {
  block_exclusive_scan(row_len, R, cu_rows);
  block_exclusive_scan(win_len, S, cu_seq);
  __shared__ int s_max[MAX_SEGMENTS];
  if (threadIdx.x < MAX_SEGMENTS) s_max[threadIdx.x] = 0;
  __syncthreads();
  for (int k = 0; k < seg.n_seg; ++k) {
    int m = 0;
    for (int i = seg.seq0[k] + threadIdx.x; i < seg.seq0[k + 1]; i += blockDim.x) m = max(m, win_len[i]);
    atomicMax(&s_max[k], m);
  }
  __syncthreads();
  if (threadIdx.x == 0) totals[0] = cu_rows[R];  // the true count, also when clamped below
  if (threadIdx.x < seg.n_seg) totals[1 + threadIdx.x] = s_max[threadIdx.x];
  if (max_total >= 0) {
    // the caller sized its buffers from a host-side count: offsets never point past them, whatever
    // the masks say (a mismatch is reported by the caller from `totals` after the forward)
    __syncthreads();
    for (int i = threadIdx.x; i <= R; i += blockDim.x) cu_rows[i] = min(cu_rows[i], max_total);
    for (int i = threadIdx.x; i <= S; i += blockDim.x) cu_seq[i] = min(cu_seq[i], max_total);
  }
}

// Row tiles of the fused query/key/value + attention GEMM (gemm_tcgen05.cu): consecutive WHOLE sequences are
// grouped greedily into tiles of at most 128 token rows, so that every sequence has its Q, K and V rows inside
// one accumulator tile.  The packed layout itself is untouched: a tile simply starts at its first sequence's row
// and the MMA computes (and discards) up to 127 rows of the next tile.
//   meta[0] = number of tiles n, meta[1 .. n + 1] = first token row of each tile, meta[n + 1] = T.
// ONE CTA: cu_seq is staged through shared memory in chunks; warp 0 walks it 32 sequences at a time (ballot of
// "still fits in the current tile"), i.e. sequential only in the number of tiles.
constexpr int TILE_ROWS = 128;
constexpr int TILES_CHUNK = 32768;
__global__ void __launch_bounds__(1024)
seq_tiles_kernel(const int32_t* __restrict__ cu_seq, int S, int32_t* __restrict__ meta, int capacity) {
  extern __shared__ int32_t s_cu[];  // cu_seq[base .. base + n] of the current chunk
  __shared__ int s_state[2];         // row0 of the open tile, tiles closed so far
  if (threadIdx.x == 0) {
    s_state[0] = cu_seq[0];
    s_state[1] = 0;
    meta[1] = cu_seq[0];
  }
  for (int base = 0; base < S; base += TILES_CHUNK) {
    const int n = min(TILES_CHUNK, S - base);
    __syncthreads();
    for (int i = threadIdx.x; i <= n; i += blockDim.x) s_cu[i] = cu_seq[base + i];
    __syncthreads();
    if (threadIdx.x < 32) {
      const int lane = threadIdx.x;
      int row0 = s_state[0], closed = s_state[1];
      int s = 0;
      while (s < n) {
        const int i = s + lane;
        const bool fits = (i < n) && (s_cu[i + 1] - row0 <= TILE_ROWS);
        const unsigned bad = __ballot_sync(0xffffffffu, !fits);
        if (bad == 0u) { s += 32; continue; }
        const int k = __ffs(bad) - 1;
        if (s + k >= n) break;              // everything left in this chunk fits
        const int start = s_cu[s + k];      // sequence s + k opens a new tile
        if (k == 0 && start == row0) {      // a single sequence longer than a tile (the caller excludes it): skip it
          s += 1;
          continue;
        }
        closed += 1;
        row0 = start;
        if (lane == 0 && 1 + closed < capacity) meta[1 + closed] = row0;
        s += k;
      }
      if (lane == 0) {
        s_state[0] = row0;
        s_state[1] = closed;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int n_tiles = s_state[1] + 1;
    meta[0] = min(n_tiles, capacity - 2);
    if (1 + n_tiles < capacity) meta[1 + n_tiles] = cu_seq[S];
  }
}

// (first row, end row) of its sequence for every packed token
__global__ void __launch_bounds__(256)
seq_token_bounds_kernel(const int32_t* __restrict__ cu_seq, int S, int2* __restrict__ bounds) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const int a = cu_seq[s], b = cu_seq[s + 1];
  for (int t = a; t < b; ++t) bounds[t] = make_int2(a, b);
}

// Token packing: ids [N, L] + mask [N, L] -> the real tokens of every row, in row order, at
// out[row_start[r] ...]; position id = column index inside its 512-token window (the reference
// restarts positions per window, Bert.py:96-99,135-138 + modeling.py:186-187).  One warp per row.
__global__ void __launch_bounds__(256)
pack_tokens_kernel(const long long* __restrict__ ids, const uint8_t* __restrict__ mask, int N, int L,
                   const int32_t* __restrict__ row_start, int window, int32_t* __restrict__ out_ids,
                   int32_t* __restrict__ out_pos, int capacity) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= N) return;
  int base = row_start[row];
  for (int c0 = 0; c0 < L; c0 += 32) {
    const int c = c0 + lane;
    const bool keep = (c < L) && mask[static_cast<long long>(row) * L + c] != 0;
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      const int k = base + __popc(bal & ((1u << lane) - 1u));
      if (k < capacity) {  // a host-supplied token count smaller than the masks' never overruns the buffers
        out_ids[k] = static_cast<int32_t>(ids[static_cast<long long>(row) * L + c]);
        out_pos[k] = c % window;
      }
    }
    base += __popc(bal);
  }
}

// fp32 [rows, K] (row stride ld) -> bf16 split [rows, parts*Kp], zero padded to Kp per part.
// 8 columns per thread: four float2 loads (PAIRS: even ld and K, 8-byte aligned base) and one 16-byte store per part.
template <bool PAIRS>
__global__ void __launch_bounds__(256)
split_bf16_kernel(const float* __restrict__ src, long long ld, const int32_t* __restrict__ row_idx,
                  long long rows, int K, int Kp, int parts, __nv_bfloat16* __restrict__ dst) {
  const int groups = Kp / 8;
  const long long total = rows * groups;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / groups;
    const int c0 = static_cast<int>(i - r * groups) * 8;
    const long long sr = row_idx ? static_cast<long long>(row_idx[r]) : r;
    const float* srow = src + sr * ld;
    float x[8];
#pragma unroll
    for (int e = 0; e < 8; e += 2) {
      const int c = c0 + e;
      if (PAIRS) {
        float2 v = make_float2(0.f, 0.f);
        if (c < K) v = __ldg(reinterpret_cast<const float2*>(srow + c));
        x[e] = v.x;
        x[e + 1] = v.y;
      } else {
        x[e] = (c < K) ? __ldg(srow + c) : 0.f;
        x[e + 1] = (c + 1 < K) ? __ldg(srow + c + 1) : 0.f;
      }
    }
    __nv_bfloat16* drow = dst + r * (static_cast<long long>(parts) * Kp) + c0;
    for (int p = 0; p < parts; ++p) {
      uint4 u;
      uint32_t* w = reinterpret_cast<uint32_t*>(&u);
#pragma unroll
      for (int e = 0; e < 8; e += 2) {
        const __nv_bfloat16 h0 = __float2bfloat16_rn(x[e]);
        const __nv_bfloat16 h1 = __float2bfloat16_rn(x[e + 1]);
        w[e / 2] = static_cast<uint32_t>(__bfloat16_as_ushort(h0)) |
                   (static_cast<uint32_t>(__bfloat16_as_ushort(h1)) << 16);
        x[e] -= __bfloat162float(h0);
        x[e + 1] -= __bfloat162float(h1);
      }
      *reinterpret_cast<uint4*>(drow + static_cast<long long>(p) * Kp) = u;
    }
  }
}

// fp32 sources laid side by side (torch.cat(..., -1) of up to 8 tensors with equal row counts) -> the bf16 split
// GEMM operand of their concatenation [rows, parts*Kp], zero padded to Kp per part: the fp32 concatenation is
// never materialised (one pass instead of n copies + one split; DESIGN.md §8 "operand splits + concat copies").
struct ConcatSrc {
  const float* p[8];
  long long pitch[8];
  int start[9];   // first column of source k in the concatenation; start[n] = K
  int n;
};
// Each thread converts 8 consecutive columns of one row: four float2 loads (a pair never straddles two sources
// when every source starts at an even column — checked by the host; odd layouts take the scalar path) and ONE
// 16-byte store per split part.  (First version: one column pair per thread with scalar loads — 100 us for the
// 1800-wide self-attention input, slower than the copies + split it replaced; profiles/r02_launches_summary_v1.txt.)
template <bool PAIRS>
__global__ void __launch_bounds__(256)
split_concat_bf16_kernel(ConcatSrc cs, long long rows, int Kp, int parts, __nv_bfloat16* __restrict__ dst) {
  const int K = cs.start[cs.n];
  const int groups = Kp / 8;
  const long long total = rows * groups;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / groups;
    const int c0 = static_cast<int>(i - r * groups) * 8;
    float x[8];
    // fast path (PAIRS): the whole 8-column group lies inside ONE source -> one source search, four float2 loads
    // (the kernel was issue-bound on the per-pair search: 81 % issue slots busy, profiles/r02c_ncu_full2_summary.txt)
    bool done = false;
    if (PAIRS && c0 + 8 <= K) {
      int q = 0;
#pragma unroll
      for (int k = 1; k < 8; ++k) q += (k < cs.n && c0 >= cs.start[k]) ? 1 : 0;
      // select the source's fields with static indices (the struct stays in the constant bank)
      const float* base = cs.p[0];
      long long pitch = cs.pitch[0];
      int st = cs.start[0], en = cs.start[1];
#pragma unroll
      for (int k = 1; k < 8; ++k)
        if (q == k) {
          base = cs.p[k];
          pitch = cs.pitch[k];
          st = cs.start[k];
          en = cs.start[k + 1];
        }
      if (c0 + 8 <= en) {
        const float2* src = reinterpret_cast<const float2*>(base + r * pitch + (c0 - st));
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 v = __ldg(src + e);
          x[2 * e] = v.x;
          x[2 * e + 1] = v.y;
        }
        done = true;
      }
    }
    if (!done)
#pragma unroll
    for (int e = 0; e < 8; e += 2) {
      const int col = c0 + e;
      x[e] = 0.f;
      x[e + 1] = 0.f;
      if (col < K) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {     // static indices: the parameter struct stays in the constant bank
          if (q < cs.n && col >= cs.start[q] && col < cs.start[q + 1]) {
            const float* p = cs.p[q] + r * cs.pitch[q] + (col - cs.start[q]);
            if (PAIRS) {
              const float2 v = __ldg(reinterpret_cast<const float2*>(p));
              x[e] = v.x;
              x[e + 1] = v.y;
            } else {
              x[e] = __ldg(p);
            }
          }
          if (!PAIRS && q < cs.n && col + 1 >= cs.start[q] && col + 1 < cs.start[q + 1])
            x[e + 1] = __ldg(cs.p[q] + r * cs.pitch[q] + (col + 1 - cs.start[q]));
        }
      }
    }
    __nv_bfloat16* drow = dst + r * (static_cast<long long>(parts) * Kp) + c0;
    for (int p = 0; p < parts; ++p) {
      uint4 u;
      uint32_t* w = reinterpret_cast<uint32_t*>(&u);
#pragma unroll
      for (int e = 0; e < 8; e += 2) {
        const __nv_bfloat16 h0 = __float2bfloat16_rn(x[e]);
        const __nv_bfloat16 h1 = __float2bfloat16_rn(x[e + 1]);
        w[e / 2] = static_cast<uint32_t>(__bfloat16_as_ushort(h0)) |
                   (static_cast<uint32_t>(__bfloat16_as_ushort(h1)) << 16);
        x[e] -= __bfloat162float(h0);
        x[e + 1] -= __bfloat162float(h1);
      }
      *reinterpret_cast<uint4*>(drow + static_cast<long long>(p) * Kp) = u;
    }
  }
}

// Row-tiled form of the same conversion (the one that runs when float2 loads are legal and Kp <= 4096): a CTA walks
// tiles of RB = 256 SLOTS / (Kp / 8) rows, and a thread keeps the SAME (row within the tile, 8-column group) slots in
// every tile, so the index division and the source search happen once per thread instead of once per 8 columns (the
// flat form shows 69 % of the issue slots busy in ncu, profiles/r02d_movers_summary.txt).  All loads of a thread's
// slots are issued before the conversions.  Worth 5-10 % (3.5-4.8 TB/s of mixed read/write traffic; four slots per
// thread cost the occupancy and run at half the speed): profiles/r02_split_concat_microbench.txt.
template <int SCR_SLOTS>
__global__ void __launch_bounds__(256)
split_concat_rows_kernel(ConcatSrc cs, long long rows, int Kp, int parts, int RB, __nv_bfloat16* __restrict__ dst) {
  const int K = cs.start[cs.n];
  const int groups = Kp / 8;
  const int items = RB * groups;
  const long long dpitch = static_cast<long long>(parts) * Kp;
  const float* sp[SCR_SLOTS];      // source of the slot's 8 columns at row 0 of a tile (nullptr: zero columns / slow path)
  long long spitch[SCR_SLOTS];
  int rl[SCR_SLOTS], c0s[SCR_SLOTS];
  bool slow[SCR_SLOTS];
#pragma unroll
  for (int j = 0; j < SCR_SLOTS; ++j) {
    const int it = threadIdx.x + 256 * j;
    rl[j] = -1;
    sp[j] = nullptr;
    spitch[j] = 0;
    c0s[j] = 0;
    slow[j] = false;
    if (it < items) {
      rl[j] = it / groups;
      const int c0 = (it - rl[j] * groups) * 8;
      c0s[j] = c0;
      if (c0 + 8 <= K) {
        int q = 0;
#pragma unroll
        for (int k = 1; k < 8; ++k) q += (k < cs.n && c0 >= cs.start[k]) ? 1 : 0;
        const float* base = cs.p[0];
        long long pitch = cs.pitch[0];
        int st = cs.start[0], en = cs.start[1];
#pragma unroll
        for (int k = 1; k < 8; ++k)
          if (q == k) {
            base = cs.p[k];
            pitch = cs.pitch[k];
            st = cs.start[k];
            en = cs.start[k + 1];
          }
        if (c0 + 8 <= en) {
          sp[j] = base + (c0 - st);
          spitch[j] = pitch;
        } else {
          slow[j] = true;
        }
      } else if (c0 < K) {
        slow[j] = true;
      }
    }
  }
  for (long long row0 = static_cast<long long>(blockIdx.x) * RB; row0 < rows; row0 += static_cast<long long>(gridDim.x) * RB) {
    float x[SCR_SLOTS][8];
#pragma unroll
    for (int j = 0; j < SCR_SLOTS; ++j) {
      const long long r = row0 + rl[j];
#pragma unroll
      for (int e = 0; e < 8; ++e) x[j][e] = 0.f;
      if (rl[j] >= 0 && r < rows) {
        if (sp[j] != nullptr) {
          const float2* src = reinterpret_cast<const float2*>(sp[j] + r * spitch[j]);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 v = __ldg(src + e);
            x[j][2 * e] = v.x;
            x[j][2 * e + 1] = v.y;
          }
        } else if (slow[j]) {  // the group straddles two sources or the end of the concatenation
#pragma unroll
          for (int e = 0; e < 8; e += 2) {
            const int col = c0s[j] + e;
            if (col < K) {
#pragma unroll
              for (int q = 0; q < 8; ++q)
                if (q < cs.n && col >= cs.start[q] && col < cs.start[q + 1]) {
                  const float2 v = __ldg(reinterpret_cast<const float2*>(cs.p[q] + r * cs.pitch[q] + (col - cs.start[q])));
                  x[j][e] = v.x;
                  x[j][e + 1] = v.y;
                }
            }
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < SCR_SLOTS; ++j) {
      const long long r = row0 + rl[j];
      if (rl[j] >= 0 && r < rows) {
        __nv_bfloat16* drow = dst + r * dpitch + c0s[j];
        for (int p = 0; p < parts; ++p) {
          uint4 u;
          uint32_t* w = reinterpret_cast<uint32_t*>(&u);
#pragma unroll
          for (int e = 0; e < 8; e += 2) {
            const uint32_t h = pack_bf16x2(x[j][e], x[j][e + 1]);
            w[e / 2] = h;
            x[j][e] -= bf16_lo(h);
            x[j][e + 1] -= bf16_hi(h);
          }
          *reinterpret_cast<uint4*>(drow + static_cast<long long>(p) * Kp) = u;
        }
      }
    }
  }
}

inline unsigned row_grid(long long rows) {
  return static_cast<unsigned>((rows + ROWS_PER_CTA - 1) / ROWS_PER_CTA);
}


}  // namespace

extern "C" int ruart_bert_embed_ln(const int32_t* ids, const int32_t* pos, const float* word_emb,
                                   const float* pos_emb, const float* type_emb, const float* gamma,
                                   const float* beta, float eps, int T, int hidden, float* out_f32,
                                   void* out_bf16, int out_parts, void* stream) {
  RUART_ARG_CHECK(hidden == 768 || hidden == 1024);
  RUART_ARG_CHECK(out_f32 != nullptr || out_bf16 != nullptr);
  if (T == 0) return RUART_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (hidden == 768)
    bert_embed_ln_kernel<3><<<row_grid(T), ROWS_PER_CTA * 32, 0, st>>>(
        ids, pos, word_emb, pos_emb, type_emb, gamma, beta, eps, T, out_f32,
        (__nv_bfloat16*)out_bf16, out_parts);
  else
    bert_embed_ln_kernel<4><<<row_grid(T), ROWS_PER_CTA * 32, 0, st>>>(
        ids, pos, word_emb, pos_emb, type_emb, gamma, beta, eps, T, out_f32,
        (__nv_bfloat16*)out_bf16, out_parts);
  RUART_LAUNCH_CHECK();
  return RUART_OK;
}

extern "C" int ruart_bert_embed_raw(const int32_t* ids, const int32_t* pos, const float* word_emb,
                                    const float* pos_emb, const float* type_emb, int T, int hidden,
                                    void* out_bf16, float* out_stats, void* stream) {
  RUART_ARG_CHECK(hidden == 768 || hidden == 1024);
  RUART_ARG_CHECK(out_bf16 != nullptr && out_stats != nullptr);
  if (T == 0) return RUART_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (hidden == 768)
    bert_embed_raw_kernel<3><<<row_grid(T), ROWS_PER_CTA * 32, 0, st>>>(
        ids, pos, word_emb, pos_emb, type_emb, T, (__nv_bfloat16*)out_bf16, (float2*)out_stats);
  else
    bert_embed_raw_kernel<4><<<row_grid(T), ROWS_PER_CTA * 32, 0, st>>>(
        ids, pos, word_emb, pos_emb, type_emb, T, (__nv_bfloat16*)out_bf16, (float2*)out_stats);
  RUART_LAUNCH_CHECK();
  return RUART_OK;
}

extern "C" int ruart_subword_coef(const float* alpha, const float* gamma, int n_layers, const float* ln_gamma,
                                  const float* ln_beta, int hidden, float* G, float* C, void* stream) {
  RUART_ARG_CHECK(alpha != nullptr && gamma != nullptr && ln_gamma != nullptr && ln_beta != nullptr);
  RUART_ARG_CHECK(G != nullptr && C != nullptr && n_layers >= 1 && hidden > 0);
  subword_coef_kernel<<<(hidden + 255) / 256, 256, 0, (cudaStream_t)stream>>>(alpha, gamma, n_layers, ln_gamma,
                                                                            ln_beta, hidden, G, C);
  RUART_LAUNCH_CHECK();
  return RUART_OK;
}

extern "C" int ruart_subword_avg_layers_fold(const void* h_bf16, long long layer_stride, const float* stats,
                                             long long stats_layer_stride, float ln_eps, const int32_t* words,
                                             int n_words, const int32_t* row_start, const uint8_t* x_mask, int W,
                                             float* dst, long long dst_stride, const float* G, const float* C,
                                             int n_layers, int hidden, void* stream) {
  RUART_ARG_CHECK(hidden == 768 || hidden == 1024);
  RUART_ARG_CHECK(h_bf16 != nullptr && stats != nullptr && G != nullptr && C != nullptr && n_layers >= 1);
  RUART_ARG_CHECK((dst_stride % 4) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0);
  RUART_ARG_CHECK((reinterpret_cast<uintptr_t>(stats) & 15u) == 0 && (reinterpret_cast<uintptr_t>(G) & 15u) == 0);
  if (n_words == 0) return RUART_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t ring = hidden == 768 ? SwFoldRing<3>::WARP_BYTES : SwFoldRing<4>::WARP_BYTES;
  const size_t smem = static_cast<size_t>(n_layers) * hidden * 4 + SWF_WARPS * ring;
  RUART_ARG_CHECK(smem <= 232448);
  static RuartDeviceOnce attr_set;
  if (!attr_set.done()) {
    RUART_CUDA_CHECK(cudaFuncSetAttribute(subword_avg_layers_fold_kernel<3>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    RUART_CUDA_CHECK(cudaFuncSetAttribute(subword_avg_layers_fold_kernel<4>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_set.set();
  }
  int grid = (n_words + SWF_WARPS - 1) / SWF_WARPS;
  if (grid > ruart_num_sms()) grid = ruart_num_sms();
  if (hidden == 768)
    subword_avg_layers_fold_kernel<3><<<grid, SWF_WARPS * 32, smem, st>>>(
        (const __nv_bfloat16*)h_bf16, layer_stride, (const float2*)stats, stats_layer_stride, ln_eps, words, n_words,
        row_start, x_mask, W, dst, dst_stride, G, C, n_layers);
  else
    subword_avg_layers_fold_kernel<4><<<grid, SWF_WARPS * 32, smem, st>>>(
        (const __nv_bfloat16*)h_bf16, layer_stride, (const float2*)stats, stats_layer_stride, ln_eps, words, n_words,
        row_start, x_mask, W, dst, dst_stride, G, C, n_layers);
  RUART_LAUNCH_CHECK();
  return RUART_OK;
}

extern "C" int ruart_add_layernorm(const float* x_f32, const void* x_bf16, const float* res_f32,
                                   const void* res_bf16, const float* gamma, const float* beta,
                                   float eps, int T, int hidden, float* out_f32, void* out_bf16,
                                   int out_parts, void* stream) {
  RUART_ARG_CHECK(hidden == 768 || hidden == 1024);
  RUART_ARG_CHECK((x_f32 != nullptr) != (x_bf16 != nullptr));
  RUART_ARG_CHECK(!(res_f32 != nullptr && res_bf16 != nullptr));
  RUART_ARG_CHECK(out_f32 != nullptr || out_bf16 != nullptr);
  if (T == 0) return RUART_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (x_bf16 != nullptr && res_f32 == nullptr && res_bf16 == nullptr && out_f32 == nullptr && out_parts == 1) {
    if (hidden == 768)
      ln_bf16_kernel<3><<<row_grid(T), ROWS_PER_CTA * 32, 0, st>>>((const __nv_bfloat16*)x_bf16, gamma, beta, eps, T,
                                                                  (__nv_bfloat16*)out_bf16);
    else
      ln_bf16_kernel<4><<<row_grid(T), ROWS_PER_CTA * 32, 0, st>>>((const __nv_bfloat16*)x_bf16, gamma, beta, eps, T,
                                                                  (__nv_bfloat16*)out_bf16);
    RUART_LAUNCH_CHECK();
    return RUART_OK;
  }
  if (hidden == 768)
    add_ln_kernel<3><<<row_grid(T), ROWS_PER_CTA * 32, 0, st>>>(
        x_f32, (const __nv_bfloat16*)x_bf16, res_f32, (const __nv_bfloat16*)res_bf16, gamma, beta,
        eps, T, out_f32, (__nv_bfloat16*)out_bf16, out_parts);
  else
    add_ln_kernel<4><<<row_grid(T), ROWS_PER_CTA * 32, 0, st>>>(
        x_f32, (const __nv_bfloat16*)x_bf16, res_f32, (const __nv_bfloat16*)res_bf16, gamma, beta,
        eps, T, out_f32, (__nv_bfloat16*)out_bf16, out_parts);
  RUART_LAUNCH_CHECK();
  return RUART_OK;
}

extern "C" int ruart_bert_attention(const float* qkv_f32, const void* qkv_bf16,
                                    const int32_t* cu_seqlens, int n_seq, int n_heads, float scale,
                                    int max_len, float* out_f32, void* out_bf16, int out_parts,
                                    void* stream) {
  RUART_ARG_CHECK((qkv_f32 != nullptr) != (qkv_bf16 != nullptr));
  RUART_ARG_CHECK(out_f32 != nullptr || out_bf16 != nullptr);
  RUART_ARG_CHECK(n_heads > 0);
  if (n_seq == 0) return RUART_OK;
  cudaStream_t st = (cudaStream_t)stream;
  static RuartDeviceOnce attr_set;
  const size_t smem_max = 4 * (2 * ATT_MAX_STAGE * 64 + 64) * sizeof(float);
  if (!attr_set.done()) {
    RUART_CUDA_CHECK(cudaFuncSetAttribute(bert_attention_kernel<float>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)smem_max));
    RUART_CUDA_CHECK(cudaFuncSetAttribute(bert_attention_kernel<__nv_bfloat16>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)smem_max));
    RUART_CUDA_CHECK(cudaFuncSetAttribute(bert_attention_mma_kernel<64>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          SHORT_WARPS * 3 * 64 * 128));
    RUART_CUDA_CHECK(cudaFuncSetAttribute(bert_attention_mma16_async_kernel,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          SHORT_WARPS * 6 * 16 * 128));
    attr_set.set();
  }
  // bf16 in, plain bf16 out: sequences of <= 64 tokens run on the warp-level MMA kernels
  int skip_upto = 0;
  static const bool no_short = getenv("RUART_NO_SHORT_ATTN") != nullptr;  // debugging aid
  if (!no_short && qkv_bf16 != nullptr && out_f32 == nullptr && out_parts == 1) {
    const long long n_tasks = static_cast<long long>(n_seq) * n_heads;
    long long ctas = (n_tasks + SHORT_WARPS - 1) / SHORT_WARPS;
    const long long cap = static_cast<long long>(ruart_num_sms()) * 16;
    if (ctas > cap) ctas = cap;
    {
      const long long pair_tasks = static_cast<long long>((n_seq + 1) / 2) * ((n_heads % 2 == 0) ? 2 : 1);
      long long ctas16 = (pair_tasks + SHORT_WARPS - 1) / SHORT_WARPS;
      if (ctas16 > cap) ctas16 = cap;
      static const bool sync16 = getenv("RUART_ATTN16_SYNC") != nullptr;  // A/B aid: no cp.async prefetch
      if (sync16)
        bert_attention_mma16_kernel<<<static_cast<unsigned>(ctas16), SHORT_WARPS * 32, 0, st>>>(
            (const __nv_bfloat16*)qkv_bf16, cu_seqlens, n_seq, n_heads, scale, (__nv_bfloat16*)out_bf16);
      else {
        const int hs = (n_heads % 2 == 0) ? 2 : 1;  // two warps share a pair's heads (finer tasks)
        long long c16 = (static_cast<long long>((n_seq + 1) / 2) * hs + SHORT_WARPS - 1) / SHORT_WARPS;
        if (c16 > cap) c16 = cap;
        bert_attention_mma16_async_kernel<<<static_cast<unsigned>(c16), SHORT_WARPS * 32,
                                            SHORT_WARPS * 6 * 16 * 128, st>>>(
            (const __nv_bfloat16*)qkv_bf16, cu_seqlens, n_seq, n_heads, scale, hs, (__nv_bfloat16*)out_bf16);
      }
      RUART_LAUNCH_CHECK();
    }
    if (max_len <= 16) return RUART_OK;
    const size_t smem64 = static_cast<size_t>(SHORT_WARPS) * 3 * 64 * 128;
    bert_attention_mma_kernel<64><<<static_cast<unsigned>(ctas), SHORT_WARPS * 32, smem64, st>>>(
        (const __nv_bfloat16*)qkv_bf16, cu_seqlens, n_seq, n_heads, scale, 16,
        (__nv_bfloat16*)out_bf16);
    RUART_LAUNCH_CHECK();
    if (max_len <= 64) return RUART_OK;
    skip_upto = 64;
  }
  int stage_tokens = max_len < 1 ? 1 : (max_len > ATT_MAX_STAGE ? ATT_MAX_STAGE : max_len);
  stage_tokens = (stage_tokens + 7) / 8 * 8;
  if (stage_tokens > 24) stage_tokens = (stage_tokens > 48) ? 64 : 48;
  const int warps = stage_tokens <= 24 ? ATT_WARPS : 4;
  const size_t smem = warps * (2 * stage_tokens * 64 + 64) * sizeof(float);
  const long long tasks = static_cast<long long>(n_seq) * n_heads;
  const unsigned grid = static_cast<unsigned>((tasks + warps - 1) / warps);
  if (qkv_f32 != nullptr)
    bert_attention_kernel<float><<<grid, warps * 32, smem, st>>>(
        qkv_f32, cu_seqlens, n_seq, n_heads, scale, stage_tokens, skip_upto, out_f32,
        (__nv_bfloat16*)out_bf16, out_parts);
  else
    bert_attention_kernel<__nv_bfloat16><<<grid, warps * 32, smem, st>>>(
        (const __nv_bfloat16*)qkv_bf16, cu_seqlens, n_seq, n_heads, scale, stage_tokens, skip_upto,
        out_f32, (__nv_bfloat16*)out_bf16, out_parts);
  RUART_LAUNCH_CHECK();
  return RUART_OK;
}

extern "C" int ruart_subword_avg_accum(const float* h_f32, const void* h_bf16, const int32_t* words,
                                       int n_words, const int32_t* row_start,
                                       const uint8_t* x_mask, int W, float* dst,
                                       long long dst_stride, const float* alpha, int n_layers,
                                       const float* gamma, int layer, int first, int hidden,
                                       void* stream) {
  RUART_ARG_CHECK(hidden == 768 || hidden == 1024);
  RUART_ARG_CHECK((h_f32 != nullptr) != (h_bf16 != nullptr));
  RUART_ARG_CHECK(alpha == nullptr || (layer >= 0 && layer < n_layers && gamma != nullptr));
  RUART_ARG_CHECK((dst_stride % 4) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0);
  if (n_words == 0) return RUART_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (hidden == 768)
    subword_avg_accum_kernel<3><<<row_grid(n_words), ROWS_PER_CTA * 32, 0, st>>>(
        h_f32, (const __nv_bfloat16*)h_bf16, words, n_words, row_start, x_mask, W, dst, dst_stride,
        alpha, n_layers, gamma, layer, first);
  else
    subword_avg_accum_kernel<4><<<row_grid(n_words), ROWS_PER_CTA * 32, 0, st>>>(
        h_f32, (const __nv_bfloat16*)h_bf16, words, n_words, row_start, x_mask, W, dst, dst_stride,
        alpha, n_layers, gamma, layer, first);
  RUART_LAUNCH_CHECK();
  return RUART_OK;
}

extern "C" int ruart_subword_avg_layers(const float* h_f32, const void* h_bf16,
                                        long long layer_stride, const int32_t* words, int n_words,
                                        const int32_t* row_start, const uint8_t* x_mask, int W,
                                        float* dst, long long dst_stride, const float* alpha,
                                        int n_layers, const float* gamma, int hidden, void* stream) {
  RUART_ARG_CHECK(hidden == 768 || hidden == 1024);
  RUART_ARG_CHECK((h_f32 != nullptr) != (h_bf16 != nullptr));
  RUART_ARG_CHECK(alpha != nullptr && gamma != nullptr && n_layers >= 1);
  RUART_ARG_CHECK((dst_stride % 4) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0);
  if (n_words == 0) return RUART_OK;
  cudaStream_t st = (cudaStream_t)stream;
  // bf16 hidden states: the cp.async-pipelined kernel (RUART_SUBWORD_ASYNC=0: A/B aid, the synchronous one)
  static const bool use_async = []() {
    const char* e = getenv("RUART_SUBWORD_ASYNC");
    return e == nullptr || e[0] != '0';
  }();
  if (use_async && h_bf16 != nullptr) {
    static RuartDeviceOnce attr_set;
    if (!attr_set.done()) {
      RUART_CUDA_CHECK(cudaFuncSetAttribute(subword_avg_layers_async_kernel<3>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            SW_WARPS * SwRing<3>::WARP_BYTES));
      RUART_CUDA_CHECK(cudaFuncSetAttribute(subword_avg_layers_async_kernel<4>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            SW_WARPS * SwRing<4>::WARP_BYTES));
      attr_set.set();
    }
    int grid = (n_words + SW_WARPS - 1) / SW_WARPS;
    if (grid > ruart_num_sms()) grid = ruart_num_sms();
    if (hidden == 768)
      subword_avg_layers_async_kernel<3><<<grid, SW_WARPS * 32, SW_WARPS * SwRing<3>::WARP_BYTES, st>>>(
          (const __nv_bfloat16*)h_bf16, layer_stride, words, n_words, row_start, x_mask, W, dst, dst_stride,
          alpha, n_layers, gamma);
    else
      subword_avg_layers_async_kernel<4><<<grid, SW_WARPS * 32, SW_WARPS * SwRing<4>::WARP_BYTES, st>>>(
          (const __nv_bfloat16*)h_bf16, layer_stride, words, n_words, row_start, x_mask, W, dst, dst_stride,
          alpha, n_layers, gamma);
    RUART_LAUNCH_CHECK();
    return RUART_OK;
  }
  if (hidden == 768)
    subword_avg_layers_kernel<3><<<row_grid(n_words), ROWS_PER_CTA * 32, 0, st>>>(
        h_f32, (const __nv_bfloat16*)h_bf16, layer_stride, words, n_words, row_start, x_mask, W, dst,
        dst_stride, alpha, n_layers, gamma);
  else
    subword_avg_layers_kernel<4><<<row_grid(n_words), ROWS_PER_CTA * 32, 0, st>>>(
        h_f32, (const __nv_bfloat16*)h_bf16, layer_stride, words, n_words, row_start, x_mask, W, dst,
        dst_stride, alpha, n_layers, gamma);
  RUART_LAUNCH_CHECK();
  return RUART_OK;
}

extern "C" int ruart_seq_lengths(const uint8_t* mask, int N, int L, int window, int32_t* row_len,
                                 int32_t* win_len, void* stream) {
  RUART_ARG_CHECK(N >= 0 && L > 0 && window > 0);
  if (N == 0) return RUART_OK;
  const int n_win = (L + window - 1) / window;
  seq_lengths_kernel<<<(N + 7) / 8, 256, 0, (cudaStream_t)stream>>>(mask, N, L, window, n_win,
                                                                    row_len, win_len);
  RUART_LAUNCH_CHECK();
  return RUART_OK;
}

extern "C" int ruart_seq_scan(const int32_t* row_len, int R, const int32_t* win_len, int S, int n_seg,
                              const int32_t* seg_row0, const int32_t* seg_seq0, int32_t* cu_rows,
                              int32_t* cu_seq, int32_t* totals, int max_total, void* stream) {
  RUART_ARG_CHECK(n_seg >= 1 && n_seg <= MAX_SEGMENTS && R >= 0 && S >= 0);
  SegTable seg;
  seg.n_seg = n_seg;
  for (int k = 0; k <= n_seg; ++k) {  // host arrays of n_seg + 1 entries
    seg.row0[k] = seg_row0[k];
    seg.seq0[k] = seg_seq0[k];
  }
  seq_scan_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(row_len, R, win_len, S, seg, cu_rows, cu_seq,
                                                        totals, max_total);
  RUART_LAUNCH_CHECK();
  return RUART_OK;
}

extern "C" int ruart_seq_tiles(const int32_t* cu_seq, int S, int32_t* meta, int capacity, int32_t* tok_bounds,
                               void* stream) {
  RUART_ARG_CHECK(S >= 0 && meta != nullptr && capacity >= 4 && tok_bounds != nullptr);
  cudaStream_t st = (cudaStream_t)stream;
  static RuartDeviceOnce attr_set;
  const size_t smem = (TILES_CHUNK + 2) * sizeof(int32_t);
  if (!attr_set.done()) {
    RUART_CUDA_CHECK(cudaFuncSetAttribute(seq_tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set.set();
  }
  seq_tiles_kernel<<<1, 1024, smem, st>>>(cu_seq, S, meta, capacity);
  RUART_LAUNCH_CHECK();
  if (S > 0) {
    seq_token_bounds_kernel<<<(S + 255) / 256, 256, 0, st>>>(cu_seq, S, reinterpret_cast<int2*>(tok_bounds));
    RUART_LAUNCH_CHECK();
  }
  return RUART_OK;
}

extern "C" int ruart_pack_tokens(const long long* ids, const uint8_t* mask, int N, int L,
                                 const int32_t* row_start, int window, int32_t* out_ids,
                                 int32_t* out_pos, int capacity, void* stream) {
  RUART_ARG_CHECK(N >= 0 && L > 0 && window > 0 && capacity >= 0);
  if (N == 0) return RUART_OK;
  pack_tokens_kernel<<<(N + 7) / 8, 256, 0, (cudaStream_t)stream>>>(ids, mask, N, L, row_start,
                                                                    window, out_ids, out_pos, capacity);
  RUART_LAUNCH_CHECK();
  return RUART_OK;
}

extern "C" int ruart_split_concat_bf16(const float* const* srcs_host, const long long* pitches_host,
                                       const int* widths_host, int n_src, long long rows, int Kp, int parts,
                                       void* dst, void* stream) {
  RUART_ARG_CHECK(n_src >= 1 && n_src <= 8 && srcs_host != nullptr && pitches_host != nullptr &&
                  widths_host != nullptr);
  RUART_ARG_CHECK(Kp > 0 && (Kp % 64) == 0 && parts >= 1 && parts <= 3 && dst != nullptr);
  ConcatSrc cs;
  int col = 0;
  for (int k = 0; k < 8; ++k) {
    cs.p[k] = k < n_src ? srcs_host[k] : nullptr;
    cs.pitch[k] = k < n_src ? pitches_host[k] : 0;
    cs.start[k] = col;
    if (k < n_src) {
      RUART_ARG_CHECK(widths_host[k] > 0 && srcs_host[k] != nullptr);
      col += widths_host[k];
    }
  }
  cs.start[8] = col;
  for (int k = n_src; k <= 8; ++k) cs.start[k] = col;
  cs.n = n_src;
  RUART_ARG_CHECK(col <= Kp);
  RUART_ARG_CHECK((reinterpret_cast<uintptr_t>(dst) & 15u) == 0);
  if (rows == 0) return RUART_OK;
  // float2 loads need every source to start at an even column with an 8-byte aligned base and an even pitch
  bool pairs = true;
  for (int k = 0; k < n_src; ++k)
    pairs = pairs && (cs.start[k] % 2 == 0) && (widths_host[k] % 2 == 0) && (pitches_host[k] % 2 == 0) &&
            ((reinterpret_cast<uintptr_t>(srcs_host[k]) & 7u) == 0);
  const long long total = rows * (Kp / 8);
  long long blocks = (total + 255) / 256;
  const long long cap = static_cast<long long>(ruart_num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  static const char* sc_env = getenv("RUART_SPLIT_CONCAT_SLOTS");  // A/B aid: 0 = the first (flat-index) form
  const int groups = Kp / 8;
  int slots = groups <= 64 ? 1 : groups <= 512 ? 2 : 0;  // measured: profiles/r02_split_concat_microbench.txt
  if (sc_env) slots = atoi(sc_env);
  if (pairs && slots > 0 && groups <= 256 * slots) {
    const int RB = (256 * slots) / groups;
    long long tiles = (rows + RB - 1) / RB;
    const long long cap2 = static_cast<long long>(ruart_num_sms()) * 16;
    const unsigned grid = static_cast<unsigned>(tiles < cap2 ? tiles : cap2);
    if (slots == 1)
      split_concat_rows_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(cs, rows, Kp, parts, RB, (__nv_bfloat16*)dst);
    else if (slots == 2)
      split_concat_rows_kernel<2><<<grid, 256, 0, (cudaStream_t)stream>>>(cs, rows, Kp, parts, RB, (__nv_bfloat16*)dst);
    else
      split_concat_rows_kernel<4><<<grid, 256, 0, (cudaStream_t)stream>>>(cs, rows, Kp, parts, RB, (__nv_bfloat16*)dst);
  } else if (pairs)
    split_concat_bf16_kernel<true><<<static_cast<unsigned>(blocks), 256, 0, (cudaStream_t)stream>>>(
        cs, rows, Kp, parts, (__nv_bfloat16*)dst);
  else
    split_concat_bf16_kernel<false><<<static_cast<unsigned>(blocks), 256, 0, (cudaStream_t)stream>>>(
        cs, rows, Kp, parts, (__nv_bfloat16*)dst);
  RUART_LAUNCH_CHECK();
  return RUART_OK;
}

extern "C" int ruart_split_bf16(const float* src, long long ld, const int32_t* row_idx,
                                long long rows, int K, int Kp, int parts, void* dst, void* stream) {
  RUART_ARG_CHECK(K > 0 && Kp >= K && (Kp % 64) == 0 && parts >= 1 && parts <= 3);
  RUART_ARG_CHECK((reinterpret_cast<uintptr_t>(dst) & 15u) == 0);
  if (rows == 0) return RUART_OK;
  const long long total = rows * (Kp / 8);
  long long blocks = (total + 255) / 256;
  const long long cap = static_cast<long long>(ruart_num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  const bool pairs = (ld % 2 == 0) && (K % 2 == 0) && ((reinterpret_cast<uintptr_t>(src) & 7u) == 0);
  if (pairs)
    split_bf16_kernel<true><<<static_cast<unsigned>(blocks), 256, 0, (cudaStream_t)stream>>>(
        src, ld, row_idx, rows, K, Kp, parts, (__nv_bfloat16*)dst);
  else
    split_bf16_kernel<false><<<static_cast<unsigned>(blocks), 256, 0, (cudaStream_t)stream>>>(
        src, ld, row_idx, rows, K, Kp, parts, (__nv_bfloat16*)dst);
  RUART_LAUNCH_CHECK();
  return RUART_OK;
}
