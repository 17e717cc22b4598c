// Kernels of the DIFFERENTIABLE form of the SDNet stack (one
// SDNetTrainer.update, reference Models/SDNetTrainer.py:330-376 — forward with autograd, backward,
// clip, Adamax).  The inference path keeps its fused kernels; the training path is built from the
// primitives below, each with an exact backward, wired as torch.autograd.Function objects in
// ruart_b200/autograd_ops.py.  BERT is locked (LOCK_BERT, SDNet.py:91-94): it needs no backward
// except the 12 + 1 scalars of the learned layer mix (alphaBERT, gammaBERT).
//
//   bmm_f32                  C[b] (+)= alpha * op(A[b]) op(B[b])        per-image score / P.x3 products and
//                                                                       their gradients (CUDA cores, fp32)
//   masked_softmax (+ bwd)   Layers.py:237-244,283-288 (masked_fill -inf, softmax over keys)
//   eltwise                  a*b, relu mask, column scale, a+b          ReLU / diagonal of AttentionScore
//   colsum                   bias / diagonal gradients (deterministic two-stage column sums)
//   split_bf16_t             fp32 [rows, K] -> TRANSPOSED split-bf16 operand [K, parts*rows_p]: feeds the
//                            tcgen05 GEMM for wgrad (dW = dY^T X) and dgrad (dX = dY W)
//   whole_ln_backward        F.layer_norm(x, x.size()) backward (Layers.py:167-168)
//   embedding_grad           nn.Embedding weight gradient, one warp per vocabulary row, deterministic
//   subword_layers_backward  d/d(alphaBERT, gammaBERT) of the subword mean + layer mix
//                            (Bert.py:149-165, SDNet.py:573-583)
//   lstm_bptt                back-propagation through time of the persistent (Bi)LSTM (Layers.py:166)
//   lstm_cell_train/backward the step-synchronous multi2one LSTM (SDNet.py:270-271)
#include <cstdlib>

#include "common.cuh"
#include "ruart_b200.h"

namespace {

using namespace ruart;

inline unsigned grid_for(long long work_items, int per_cta) {
  long long g = (work_items + per_cta - 1) / per_cta;
  const long long cap = static_cast<long long>(ruart_num_sms()) * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<unsigned>(g);
}

// ------------------------------------------------------------------------------------------ bmm
constexpr int TB_M = 64, TB_N = 64, TB_K = 16;

__global__ void __launch_bounds__(256)
bmm_f32_kernel(const float* __restrict__ A, long long lda, long long sa, int transA,
               const float* __restrict__ Bm, long long ldb, long long sb, int transB,
               float* __restrict__ C, long long ldc, long long sc, int M, int N, int K, float alpha,
               int accumulate) {
  __shared__ float As[TB_K][TB_M + 4];
  __shared__ float Bs[TB_K][TB_N + 4];
  const int b = blockIdx.z;
  A += static_cast<long long>(b) * sa;
  Bm += static_cast<long long>(b) * sb;
  C += static_cast<long long>(b) * sc;
  const int m0 = blockIdx.y * TB_M, n0 = blockIdx.x * TB_N;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += TB_K) {
    for (int i = threadIdx.x; i < TB_M * TB_K; i += 256) {
      int m, k;
      if (transA) { m = i % TB_M; k = i / TB_M; } else { k = i % TB_K; m = i / TB_K; }
      const int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < M && gk < K)
        v = transA ? A[static_cast<long long>(gk) * lda + gm] : A[static_cast<long long>(gm) * lda + gk];
      As[k][m] = v;
    }
    for (int i = threadIdx.x; i < TB_N * TB_K; i += 256) {
      int n, k;
      if (transB) { k = i % TB_K; n = i / TB_K; } else { n = i % TB_N; k = i / TB_N; }
      const int gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < N && gk < K)
        v = transB ? Bm[static_cast<long long>(gn) * ldb + gk] : Bm[static_cast<long long>(gk) * ldb + gn];
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TB_K; ++kk) {
      float a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bb[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float* p = C + static_cast<long long>(gm) * ldc + gn;
      const float v = alpha * acc[i][j];
      *p = accumulate ? (*p + v) : v;
    }
  }
}

// ------------------------------------------------------------------------------------------ softmax
// One warp per row r of x [rows = B*L1, L2]; mask [B, L2] (0 = masked -> -inf, Layers.py:283-284),
// NULL = no mask.  A row without a live key gives NaN, like torch.softmax over all -inf.
__global__ void __launch_bounds__(256)
masked_softmax_kernel(const float* __restrict__ x, long long x_pitch, const uint8_t* __restrict__ mask,
                      int L1, int L2, long long rows, float* __restrict__ out, long long out_pitch) {
  const long long r = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float* xr = x + r * x_pitch;
  float* orow = out + r * out_pitch;
  const uint8_t* mr = mask ? mask + (r / L1) * L2 : nullptr;
  float mx = -INFINITY;
  for (int j = lane; j < L2; j += 32)
    if (!mr || mr[j]) mx = fmaxf(mx, xr[j]);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < L2; j += 32)
    if (!mr || mr[j]) sum += expf(xr[j] - mx);
  sum = warp_sum(sum);
  const bool dead = (mx == -INFINITY);   // every key masked (or all -inf): torch gives NaN
  for (int j = lane; j < L2; j += 32) {
    float v;
    if (dead) v = __int_as_float(0x7fc00000);
    else v = (!mr || mr[j]) ? expf(xr[j] - mx) / sum : 0.f;
    orow[j] = v;
  }
}

// dx = p * (dp - sum_j p_j dp_j)
__global__ void __launch_bounds__(256)
softmax_backward_kernel(const float* __restrict__ p, long long p_pitch, const float* __restrict__ dp,
                        long long dp_pitch, float* __restrict__ dx, long long dx_pitch, long long rows,
                        int cols) {
  const long long r = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float* pr = p + r * p_pitch;
  const float* dr = dp + r * dp_pitch;
  float dot = 0.f;
  for (int j = lane; j < cols; j += 32) dot = fmaf(pr[j], dr[j], dot);
  dot = warp_sum(dot);
  float* xr = dx + r * dx_pitch;
  for (int j = lane; j < cols; j += 32) xr[j] = pr[j] * (dr[j] - dot);
}

// ------------------------------------------------------------------------------------------ eltwise
// op 0: out = a * b          op 1: out = a * (b > 0)   (ReLU backward: a = dy, b = relu output)
// op 2: out = a + b          op 3: out = a * v[c]      (v has v_len entries; v_len == 1: scalar)
// op 4: out = max(a, 0)      op 5: out = a + v[c]
template <int OP>
__global__ void __launch_bounds__(256)
eltwise_kernel(const float* __restrict__ a, long long a_pitch, const float* __restrict__ b,
               long long b_pitch, const float* __restrict__ v, int v_len, float* __restrict__ out,
               long long out_pitch, long long rows, int cols) {
  const long long total = rows * cols;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / cols;
    const int c = static_cast<int>(i - r * cols);
    const float av = a[r * a_pitch + c];
    float o;
    if (OP == 0) o = av * b[r * b_pitch + c];
    else if (OP == 1) o = (b[r * b_pitch + c] > 0.f) ? av : 0.f;
    else if (OP == 2) o = av + b[r * b_pitch + c];
    else if (OP == 3) o = av * v[v_len > 1 ? c : 0];
    else if (OP == 4) o = fmaxf(av, 0.f);
    else o = av + v[v_len > 1 ? c : 0];
    out[r * out_pitch + c] = o;
  }
}

// out = mask ? x : value   (flat; the reference's scores.data.masked_fill_(mask == 0, -inf))
__global__ void mask_fill_kernel(const float* __restrict__ x, const uint8_t* __restrict__ mask, long long n,
                                 float value, float* __restrict__ out) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    out[i] = mask[i] ? x[i] : value;
}

// ------------------------------------------------------------------------------------------ colsum
// out[c] = sum_r x[r][c], deterministic: stage 1 sums row chunk blockIdx.y of column c in double,
// stage 2 adds the chunk partials in a fixed order.
constexpr int CS_CHUNKS = 64;
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const float* __restrict__ x, long long pitch, long long rows, int cols,
                      double* __restrict__ partials) {
  __shared__ double s[8][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int ry = threadIdx.x >> 5;
  const long long per = (rows + gridDim.y - 1) / gridDim.y;
  const long long r0 = per * blockIdx.y, r1 = (r0 + per < rows) ? r0 + per : rows;
  double acc = 0.0;
  if (c < cols)
    for (long long r = r0 + ry; r < r1; r += 8) acc += x[r * pitch + c];
  s[ry][threadIdx.x & 31] = acc;
  __syncthreads();
  if (ry == 0 && c < cols) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += s[i][threadIdx.x & 31];
    partials[static_cast<long long>(blockIdx.y) * cols + c] = t;
  }
}

__global__ void colsum_final_kernel(const double* __restrict__ partials, int chunks, int cols,
                                    float* __restrict__ out, int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  double t = 0.0;
  for (int i = 0; i < chunks; ++i) t += partials[static_cast<long long>(i) * cols + c];
  out[c] = accumulate ? out[c] + static_cast<float>(t) : static_cast<float>(t);
}

// ------------------------------------------------------------------------------------------ split^T
// dst[k][p * rows_p + r] = part p of src[r][k]   (bf16 hi / mid / lo parts: x = x0 + x1 + x2), zero for
// r >= rows.  32 x 32 tiles through shared memory; dst row pitch = parts * rows_p.
__global__ void __launch_bounds__(256)
split_bf16_t_kernel(const float* __restrict__ src, long long ld, long long rows, int K, long long rows_p,
                    int parts, __nv_bfloat16* __restrict__ dst) {
  __shared__ float tile[32][33];
  const long long r0 = static_cast<long long>(blockIdx.x) * 32;
  const int k0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 8 rows of threads
  for (int i = ty; i < 32; i += 8) {
    const long long r = r0 + i;
    const int k = k0 + tx;
    tile[i][tx] = (r < rows && k < K) ? src[r * ld + k] : 0.f;
  }
  __syncthreads();
  const long long dpitch = static_cast<long long>(parts) * rows_p;
  for (int i = ty; i < 32; i += 8) {
    const int k = k0 + i;
    const long long r = r0 + tx;
    if (k >= K || r >= rows_p) continue;
    float rem = tile[tx][i];
    for (int p = 0; p < parts; ++p) {
      const __nv_bfloat16 hb = __float2bfloat16_rn(rem);
      dst[static_cast<long long>(k) * dpitch + static_cast<long long>(p) * rows_p + r] = hb;
      rem -= __bfloat162float(hb);
    }
  }
}

// ------------------------------------------------------------------------------------------ whole LN
// y = (x - mean) * rstd over ALL elements (Layers.py:167-168); dx = rstd * (dy - mean(dy) - y * mean(dy*y)).
__global__ void __launch_bounds__(256)
whole_ln_bwd_stats_kernel(const float* __restrict__ y, long long y_pitch, const float* __restrict__ dy,
                          long long dy_pitch, long long rows, int cols, double* __restrict__ partials) {
  __shared__ double s_a[8], s_b[8];
  const long long total = rows * cols;
  double a = 0.0, b = 0.0;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / cols;
    const int c = static_cast<int>(i - r * cols);
    const float g = dy[r * dy_pitch + c];
    a += g;
    b += static_cast<double>(g) * y[r * y_pitch + c];
  }
  a = warp_sum_d(a);
  b = warp_sum_d(b);
  if ((threadIdx.x & 31) == 0) {
    s_a[threadIdx.x >> 5] = a;
    s_b[threadIdx.x >> 5] = b;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0.0, tb = 0.0;
    for (int w = 0; w < 8; ++w) {
      ta += s_a[w];
      tb += s_b[w];
    }
    partials[2 * blockIdx.x] = ta;
    partials[2 * blockIdx.x + 1] = tb;
  }
}

__global__ void __launch_bounds__(256)
whole_ln_bwd_apply_kernel(const float* __restrict__ y, long long y_pitch, const float* __restrict__ dy,
                          long long dy_pitch, long long rows, int cols, const double* __restrict__ partials,
                          int n_parts, const float* __restrict__ stats, float* __restrict__ dx,
                          long long dx_pitch) {
  __shared__ float s_m1, s_m2;
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
      s_m1 = static_cast<float>(a / n);
      s_m2 = static_cast<float>(b / n);
    }
  }
  __syncthreads();
  const float m1 = s_m1, m2 = s_m2, rstd = stats[1];
  const long long total = rows * cols;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / cols;
    const int c = static_cast<int>(i - r * cols);
    dx[r * dx_pitch + c] = rstd * (dy[r * dy_pitch + c] - m1 - y[r * y_pitch + c] * m2);
  }
}

// ------------------------------------------------------------------------------------------ embedding
// dW[v] (+)= sum over k with ids[k] == v of dy[k], in ascending k (deterministic, no atomics).
// Pass 1 flags the positions whose gradient row is not all zero (the pad-word slots of an item never reach
// the output: their rows are exactly zero, and there are ~10x more of them than real words — without the
// flags the warp of vocabulary row 0 would walk all of them serially).  Pass 2: one warp per vocabulary
// row scans the (id, flag) list 32 entries at a time and adds the matching rows in order.
__global__ void __launch_bounds__(256)
row_nonzero_kernel(const float* __restrict__ dy, long long dy_pitch, long long n, int D,
                   uint8_t* __restrict__ flag) {
  const int lane = threadIdx.x & 31;
  const long long warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long k = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; k < n; k += warps) {
    const float* row = dy + k * dy_pitch;
    bool nz = false;
    for (int c = lane; c < D; c += 32) nz |= (row[c] != 0.f);
    nz = __any_sync(0xffffffffu, nz);
    if (lane == 0) flag[k] = nz ? 1 : 0;
  }
}

template <typename I>
__global__ void __launch_bounds__(256)
embedding_grad_kernel(const I* __restrict__ ids, const uint8_t* __restrict__ flag, long long n,
                      const float* __restrict__ dy, long long dy_pitch, int D, int V, int n_seg,
                      float* __restrict__ partial, float* __restrict__ dW, long long dw_pitch, int accumulate) {
  // task = (segment of the id list, vocabulary row); n_seg > 1 (small tables: pos / ent embeddings) writes
  // partial[seg][v][:] and embedding_grad_reduce_kernel adds the segments in order
  const int lane = threadIdx.x & 31;
  const long long warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const long long tasks = static_cast<long long>(V) * n_seg;
  const long long seg_len = ((n + n_seg - 1) / n_seg + 31) / 32 * 32;
  for (long long task = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; task < tasks;
       task += warps) {
    const int seg = static_cast<int>(task / V);
    const int v = static_cast<int>(task - static_cast<long long>(seg) * V);
    const long long k0 = seg * seg_len, k1 = (k0 + seg_len < n) ? k0 + seg_len : n;
    for (int c0 = 0; c0 < D; c0 += 32 * 8) {
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.f;
      for (long long base = k0; base < k1; base += 32) {
        const long long k = base + lane;
        const bool hit = (k < k1) && (static_cast<long long>(ids[k]) == v) && (flag[k] != 0);
        unsigned m = __ballot_sync(0xffffffffu, hit);
        while (m) {
          const int t = __ffs(m) - 1;
          m &= m - 1;
          const float* row = dy + (base + t) * dy_pitch;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int c = c0 + lane + 32 * i;
            if (c < D) acc[i] += row[c];
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = c0 + lane + 32 * i;
        if (c < D) {
          if (n_seg > 1) {
            partial[(static_cast<long long>(seg) * V + v) * D + c] = acc[i];
          } else {
            float* p = dW + static_cast<long long>(v) * dw_pitch + c;
            *p = accumulate ? *p + acc[i] : acc[i];
          }
        }
      }
    }
  }
}

__global__ void embedding_grad_reduce_kernel(const float* __restrict__ partial, int n_seg, int V, int D,
                                             float* __restrict__ dW, long long dw_pitch, int accumulate) {
  const long long total = static_cast<long long>(V) * D;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float t = 0.f;
    for (int s = 0; s < n_seg; ++s) t += partial[static_cast<long long>(s) * total + i];
    const long long v = i / D;
    float* p = dW + v * dw_pitch + (i - v * D);
    *p = accumulate ? *p + t : t;
  }
}

// Sorted form (the default): work proportional to the number of live positions instead of V x n.
//   1. flag + count    live positions per vocabulary row (integer atomics: the COUNTS are deterministic)
//   2. scan            start[v] = exclusive prefix of the counts, start[V] = number of live positions
//   3. fill            the live positions of row v land in pos[start[v] .. start[v+1]) in arbitrary order
//   4. rank            every entry counts the smaller entries of its own bucket -> sorted[] holds each bucket in
//                      ascending k (one thread per entry: a heavy bucket is ranked by as many threads as it has
//                      entries, all reading the same addresses)
//   5. slab sums       one warp per EG_SLAB consecutive entries of sorted[], whatever rows they belong to: rows
//                      lying inside one slab are written straight to dW, a row that crosses slab borders leaves one
//                      partial sum per slab (at most the first and the last run of a slab can be such)
//   6. combine         rows that cross slab borders: partial sums added in slab order; rows without a live
//                      position: zeros
// The summation order is a function of the ids alone (ascending k inside a slab, slabs in ascending order), so the
// result is deterministic, and equal bit for bit to the one-warp-per-row walk for every row inside one slab.
constexpr int EG_SLAB = 64;
constexpr int EG_ACC = 10;   // columns per lane and pass: D <= 320 (the 300-wide word tables) in one pass

template <typename I>
__global__ void __launch_bounds__(256)
eg_flag_count_kernel(const float* __restrict__ dy, long long dy_pitch, long long n, int D,
                     const I* __restrict__ ids, int V, uint8_t* __restrict__ flag, int* __restrict__ cnt) {
  const int lane = threadIdx.x & 31;
  const long long warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long k = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; k < n; k += warps) {
    const float* row = dy + k * dy_pitch;
    bool nz = false;
    for (int c = lane; c < D; c += 32) nz |= (row[c] != 0.f);
    nz = __any_sync(0xffffffffu, nz);
    if (lane == 0) {
      const long long v = static_cast<long long>(ids[k]);
      const bool live = nz && v >= 0 && v < V;      // ids outside the table have no row to add to
      flag[k] = live ? 1 : 0;
      if (live) atomicAdd(cnt + v, 1);
    }
  }
}

__global__ void __launch_bounds__(1024) eg_scan_kernel(int* __restrict__ cnt, int* __restrict__ start, int V) {
  __shared__ int s_part[1024];
  const int tid = threadIdx.x;
  const int per = (V + 1023) / 1024;
  const int b = tid * per, e = (b + per < V) ? b + per : V;
  int sum = 0;
  for (int i = b; i < e; ++i) sum += cnt[i];
  s_part[tid] = sum;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {
    const int add = (tid >= off) ? s_part[tid - off] : 0;
    __syncthreads();
    s_part[tid] += add;
    __syncthreads();
  }
  int run = s_part[tid] - sum;
  for (int i = b; i < e; ++i) {
    const int c = cnt[i];
    start[i] = run;
    cnt[i] = 0;          // becomes the fill cursor
    run += c;
  }
  if (tid == 1023) start[V] = s_part[1023];
}

template <typename I>
__global__ void __launch_bounds__(256)
eg_fill_kernel(const I* __restrict__ ids, const uint8_t* __restrict__ flag, long long n,
               const int* __restrict__ start, int* __restrict__ cursor, int* __restrict__ pos) {
  for (long long k = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; k < n;
       k += static_cast<long long>(gridDim.x) * blockDim.x) {
    if (flag[k]) {
      const int v = static_cast<int>(ids[k]);
      pos[start[v] + atomicAdd(cursor + v, 1)] = static_cast<int>(k);
    }
  }
}

template <typename I>
__global__ void __launch_bounds__(256)
eg_rank_kernel(const I* __restrict__ ids, const int* __restrict__ start, int V, const int* __restrict__ pos,
               int* __restrict__ sorted) {
  const int total = start[V];
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int k = pos[i];
    const int v = static_cast<int>(ids[k]);
    const int s = start[v], e = start[v + 1];
    int rank = 0;
    for (int j = s; j < e; ++j) rank += (pos[j] < k) ? 1 : 0;
    sorted[s + rank] = k;
  }
}

template <typename I>
__global__ void __launch_bounds__(256)
eg_slab_kernel(const I* __restrict__ ids, const int* __restrict__ start, int V, const int* __restrict__ sorted,
               const float* __restrict__ dy, long long dy_pitch, int D, float* __restrict__ partial,
               float* __restrict__ dW, long long dw_pitch, int accumulate) {
  __shared__ int s_k[8][EG_SLAB];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int total = start[V];
  const int n_slabs = (total + EG_SLAB - 1) / EG_SLAB;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < n_slabs; t += warps) {
    const int b0 = t * EG_SLAB, b1 = (b0 + EG_SLAB < total) ? b0 + EG_SLAB : total;
    for (int j = lane; j < b1 - b0; j += 32) s_k[w][j] = sorted[b0 + j];
    __syncwarp();
    for (int c0 = 0; c0 < D; c0 += 32 * EG_ACC) {
      int j = 0;
      while (j < b1 - b0) {
        const int v = static_cast<int>(ids[s_k[w][j]]);
        const int s = start[v], e = start[v + 1];
        const int j2 = ((e < b1) ? e : b1) - b0;            // the run of row v inside this slab is [j, j2)
        float acc[EG_ACC];
#pragma unroll
        for (int i = 0; i < EG_ACC; ++i) acc[i] = 0.f;
        int jj = j;
        for (; jj + 4 <= j2; jj += 4) {                       // four gradient rows in flight, added in order
          float r[4][EG_ACC];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float* row = dy + static_cast<long long>(s_k[w][jj + u]) * dy_pitch;
#pragma unroll
            for (int i = 0; i < EG_ACC; ++i) {
              const int c = c0 + lane + 32 * i;
              r[u][i] = (c < D) ? row[c] : 0.f;
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < EG_ACC; ++i) acc[i] += r[u][i];
        }
        for (; jj < j2; ++jj) {
          const float* row = dy + static_cast<long long>(s_k[w][jj]) * dy_pitch;
#pragma unroll
          for (int i = 0; i < EG_ACC; ++i) {
            const int c = c0 + lane + 32 * i;
            if (c < D) acc[i] += row[c];
          }
        }
        const bool complete = (s >= b0) && (e <= b1);
        float* out = complete ? dW + static_cast<long long>(v) * dw_pitch
                              : partial + static_cast<long long>(j == 0 ? 2 * t : 2 * t + 1) * D;
#pragma unroll
        for (int i = 0; i < EG_ACC; ++i) {
          const int c = c0 + lane + 32 * i;
          if (c < D) out[c] = (complete && accumulate) ? out[c] + acc[i] : acc[i];
        }
        j = j2;
      }
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(256)
eg_combine_kernel(const int* __restrict__ start, int V, const float* __restrict__ partial, int D,
                  float* __restrict__ dW, long long dw_pitch, int accumulate) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; v < V; v += warps) {
    const int s = start[v], e = start[v + 1];
    float* out = dW + static_cast<long long>(v) * dw_pitch;
    if (e == s) {
      if (!accumulate)
        for (int c = lane; c < D; c += 32) out[c] = 0.f;
      continue;
    }
    const int ta = s / EG_SLAB, tb = (e - 1) / EG_SLAB;
    if (ta == tb) continue;                                   // inside one slab: written by eg_slab_kernel
    for (int c = lane; c < D; c += 32) {
      float sum = 0.f;
      for (int t = ta; t <= tb; ++t)
        sum += partial[static_cast<long long>(s <= t * EG_SLAB ? 2 * t : 2 * t + 1) * D + c];
      out[c] = accumulate ? out[c] + sum : sum;
    }
  }
}

inline long long eg_round16(long long x) { return (x + 15) / 16 * 16; }

// ------------------------------------------------------------------------------------------ subword
// s[l] = sum over words of < dy[word], mean_{t in [st,ed)} h_l[row_start[item] + t] >   (Bert.py:149-165):
// the only quantity the gradients of alphaBERT / gammaBERT need (SDNet.py:573-583).  One warp per word.
constexpr int SW_MAX_LAYERS = 24;
template <typename T>
__global__ void __launch_bounds__(256)
subword_layers_bwd_kernel(const T* __restrict__ h, long long layer_stride, const int32_t* __restrict__ words,
                          int n_words, const int32_t* __restrict__ row_start, const uint8_t* __restrict__ x_mask,
                          int W, const float* __restrict__ dy, long long dy_stride, int n_layers, int hidden,
                          double* __restrict__ partials) {
  __shared__ double s_acc[8][SW_MAX_LAYERS];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  if (lane < SW_MAX_LAYERS) s_acc[wid][lane] = 0.0;
  __syncwarp();
  for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n_words; w += warps) {
    const int item = words[w];
    const int j = words[1LL * n_words + w];
    const int st = words[2LL * n_words + w];
    const int ed = words[3LL * n_words + w];
    if (j >= W || x_mask[static_cast<long long>(item) * W + j] == 0 || ed <= st) continue;
    const int cnt = ed - st;
    const float inv = (cnt > 1) ? 1.0f / static_cast<float>(cnt) : 1.0f;
    const float* g = dy + (static_cast<long long>(item) * W + j) * dy_stride;
    const long long t0 = static_cast<long long>(row_start[item]) + st;
#pragma unroll 1
    for (int l = 0; l < n_layers; ++l) {
      const T* hl = h + static_cast<long long>(l) * layer_stride + t0 * hidden;
      float dot = 0.f;
      for (int c = lane; c < hidden; c += 32) {
        float sum = 0.f;
        for (int t = 0; t < cnt; ++t) sum += static_cast<float>(hl[static_cast<long long>(t) * hidden + c]);
        dot = fmaf(sum * inv, g[c], dot);
      }
      dot = warp_sum(dot);
      if (lane == 0) s_acc[wid][l] += dot;   // per-warp accumulator, words in a fixed order
    }
  }
  __syncthreads();
  if (threadIdx.x < n_layers) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s_acc[w][threadIdx.x];
    partials[static_cast<long long>(blockIdx.x) * n_layers + threadIdx.x] = t;
  }
}

// bf16 hidden states, hidden a multiple of 256 (768 for BERT-base): every lane owns hidden/256 groups of 8
// consecutive columns, so a token row is read with 16-byte loads (hidden/256 per lane, all independent), the
// gradient row is held in registers for the twelve layers, and two layers are in flight at a time.
constexpr int SWV_MAX = 4;   // hidden <= 1024
__global__ void __launch_bounds__(256)
subword_layers_bwd_bf16v_kernel(const __nv_bfloat16* __restrict__ h, long long layer_stride,
                                const int32_t* __restrict__ words, int n_words,
                                const int32_t* __restrict__ row_start, const uint8_t* __restrict__ x_mask, int W,
                                const float* __restrict__ dy, long long dy_stride, int n_layers, int hidden,
                                double* __restrict__ partials) {
  __shared__ double s_acc[8][SW_MAX_LAYERS];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int nv = hidden >> 8;
  if (lane < SW_MAX_LAYERS) s_acc[wid][lane] = 0.0;
  __syncwarp();
  for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n_words; w += warps) {
    const int item = words[w];
    const int j = words[1LL * n_words + w];
    const int st = words[2LL * n_words + w];
    const int ed = words[3LL * n_words + w];
    if (j >= W || x_mask[static_cast<long long>(item) * W + j] == 0 || ed <= st) continue;
    const int cnt = ed - st;
    const float inv = (cnt > 1) ? 1.0f / static_cast<float>(cnt) : 1.0f;
    const float* g = dy + (static_cast<long long>(item) * W + j) * dy_stride;
    float gr[SWV_MAX][8];
#pragma unroll
    for (int i = 0; i < SWV_MAX; ++i) {
      if (i < nv) {
        const float4 a = *reinterpret_cast<const float4*>(g + (lane + 32 * i) * 8);
        const float4 b = *reinterpret_cast<const float4*>(g + (lane + 32 * i) * 8 + 4);
        gr[i][0] = a.x * inv; gr[i][1] = a.y * inv; gr[i][2] = a.z * inv; gr[i][3] = a.w * inv;
        gr[i][4] = b.x * inv; gr[i][5] = b.y * inv; gr[i][6] = b.z * inv; gr[i][7] = b.w * inv;
      }
    }
    const long long t0 = static_cast<long long>(row_start[item]) + st;
#pragma unroll 2
    for (int l = 0; l < n_layers; ++l) {
      const __nv_bfloat16* hl = h + static_cast<long long>(l) * layer_stride + t0 * hidden;
      float dot = 0.f;
#pragma unroll
      for (int i = 0; i < SWV_MAX; ++i) {
        if (i < nv) {
          float sum[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) sum[u] = 0.f;
          for (int t = 0; t < cnt; ++t) {
            const uint4 v = *reinterpret_cast<const uint4*>(hl + static_cast<long long>(t) * hidden + (lane + 32 * i) * 8);
            const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float2 f = __bfloat1622float2(p[u]);
              sum[2 * u] += f.x;
              sum[2 * u + 1] += f.y;
            }
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) dot = fmaf(sum[u], gr[i][u], dot);
        }
      }
      dot = warp_sum(dot);
      if (lane == 0) s_acc[wid][l] += dot;   // per-warp accumulator, words in a fixed order
    }
  }
  __syncthreads();
  if (threadIdx.x < n_layers) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s_acc[w][threadIdx.x];
    partials[static_cast<long long>(blockIdx.x) * n_layers + threadIdx.x] = t;
  }
}

// one thread: s[l] = sum of partials; a = softmax(alpha); out = gamma * sum_l a_l m_l
//   d gamma = sum_l a_l s_l ; d a_l = gamma s_l ; d alpha_l = a_l (d a_l - sum_k a_k d a_k)
__global__ void layer_mix_bwd_kernel(const double* __restrict__ partials, int n_blocks, int n_layers,
                                     const float* __restrict__ alpha, const float* __restrict__ gamma,
                                     float* __restrict__ dalpha, float* __restrict__ dgamma, int accumulate) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double s[SW_MAX_LAYERS], a[SW_MAX_LAYERS];
  float mx = -INFINITY;
  for (int l = 0; l < n_layers; ++l) mx = fmaxf(mx, alpha[l]);
  double den = 0.0;
  for (int l = 0; l < n_layers; ++l) {
    a[l] = exp(static_cast<double>(alpha[l] - mx));
    den += a[l];
    double t = 0.0;
    for (int b = 0; b < n_blocks; ++b) t += partials[static_cast<long long>(b) * n_layers + l];
    s[l] = t;
  }
  const double gm = gamma[0];
  double dg = 0.0, dotda = 0.0;
  for (int l = 0; l < n_layers; ++l) {
    a[l] /= den;
    dg += a[l] * s[l];
    dotda += a[l] * gm * s[l];
  }
  for (int l = 0; l < n_layers; ++l) {
    const float v = static_cast<float>(a[l] * (gm * s[l] - dotda));
    dalpha[l] = accumulate ? dalpha[l] + v : v;
  }
  dgamma[0] = accumulate ? dgamma[0] + static_cast<float>(dg) : static_cast<float>(dg);
}

// ------------------------------------------------------------------------------------------ LSTM BPTT
// Backward of the persistent (Bi)LSTM recurrence (lstm.cu) for H <= 128.  One CTA owns BP_BT sequences
// of one direction and walks the steps in reverse.  Per step:
//   phase 1 (thread = (sequence b, unit j)): dh = dout[t] + dh_rec ; gate gradients from the saved
//           activations (i, f, g, o, c) ; writes d(pre-activation) to dxg[t] and to shared memory
//   phase 2 (thread = (unit k, gate q)): dh_rec[b][k] = sum_r d_pre[b][r] W_hh[r][k] — W_hh is read through
//           L1/L2 (250 KB, coalesced along k), partial sums of the four gates meet by warp shuffles.
// dW_ih, dW_hh, db and dx follow from dxg by GEMMs / column sums outside.
constexpr int BP_THREADS = 512;
constexpr int BP_BT = 4;
constexpr int BP_HP = 128;
constexpr int BP_ST = 132;      // shared-memory stride of one gate's gradients: the four gates a warp reads at once
                                // (lane = unit*4 + gate) start 33 quad-banks apart -> conflict-free LDS.128
constexpr int BP_UNROLL = 32;   // weights in flight per thread in the recurrent-gradient product

__global__ void __launch_bounds__(BP_THREADS, 1)
lstm_bptt_kernel(const float* __restrict__ gates, long long gates_pitch, const float* __restrict__ w_hh,
                 const float* __restrict__ dout, long long dout_pitch, float* __restrict__ dxg,
                 long long dxg_pitch, int B, int L, int H) {
  __shared__ __align__(16) float s_dg[BP_BT][4 * BP_ST];
  __shared__ float s_dh[BP_BT][BP_HP];
  const int t = threadIdx.x;
  const int dir = blockIdx.y;
  const int b0 = blockIdx.x * BP_BT;
  const float* W = w_hh + static_cast<long long>(dir) * 4 * H * H;
  for (int i = t; i < BP_BT * 4 * BP_ST; i += BP_THREADS) (&s_dg[0][0])[i] = 0.f;
  for (int i = t; i < BP_BT * BP_HP; i += BP_THREADS) (&s_dh[0][0])[i] = 0.f;
  __syncthreads();
  // phase-1 role
  const int pb = t / H, pj = t - pb * H;
  const int bb = b0 + pb;
  const bool p1 = (pb < BP_BT) && (bb < B);
  float dc_carry = 0.f;
  // phase-2 role
  const int k = t >> 2, q = t & 3;
  const bool p2 = k < H;
  for (int s = L - 1; s >= 0; --s) {
    const int tt = dir == 0 ? s : (L - 1 - s);
    if (p1) {
      const float* row = gates + (static_cast<long long>(bb) * L + tt) * gates_pitch + dir * 5 * H;
      const float gi = row[pj], gf = row[H + pj], gg = row[2 * H + pj], go = row[3 * H + pj];
      const float c = row[4 * H + pj];
      float c_prev = 0.f;
      if (s > 0) {
        const int tp = dir == 0 ? tt - 1 : tt + 1;
        c_prev = gates[(static_cast<long long>(bb) * L + tp) * gates_pitch + dir * 5 * H + 4 * H + pj];
      }
      const float dh = dout[(static_cast<long long>(bb) * L + tt) * dout_pitch + dir * H + pj] + s_dh[pb][pj];
      const float tc = tanhf(c);
      const float d_o = dh * tc;
      const float dc = dc_carry + dh * go * (1.0f - tc * tc);
      const float dai = dc * gg * gi * (1.0f - gi);
      const float daf = dc * c_prev * gf * (1.0f - gf);
      const float dag = dc * gi * (1.0f - gg * gg);
      const float dao = d_o * go * (1.0f - go);
      dc_carry = dc * gf;
      float* xr = dxg + (static_cast<long long>(bb) * L + tt) * dxg_pitch + dir * 4 * H;
      xr[pj] = dai;
      xr[H + pj] = daf;
      xr[2 * H + pj] = dag;
      xr[3 * H + pj] = dao;
      s_dg[pb][pj] = dai;
      s_dg[pb][BP_ST + pj] = daf;
      s_dg[pb][2 * BP_ST + pj] = dag;
      s_dg[pb][3 * BP_ST + pj] = dao;
    }
    __syncthreads();
    if (s > 0) {   // the recurrent gradient is not needed before the first step
      float acc[BP_BT];
#pragma unroll
      for (int b = 0; b < BP_BT; ++b) acc[b] = 0.f;
      if (p2) {
        // W_hh (250 KB) does not fit shared memory: its column k is streamed from L1/L2 with BP_UNROLL
        // independent loads in flight per thread (5 in flight made a step 20 us; the loads, not the FMAs, bound it)
        const float* wq = W + static_cast<long long>(q) * H * H + k;
        for (int j0 = 0; j0 < H; j0 += BP_UNROLL) {
          float w[BP_UNROLL];
#pragma unroll
          for (int u = 0; u < BP_UNROLL; ++u)
            w[u] = (j0 + u < H) ? __ldg(wq + static_cast<long long>(j0 + u) * H) : 0.f;
#pragma unroll
          for (int u = 0; u < BP_UNROLL; u += 4) {
#pragma unroll
            for (int b = 0; b < BP_BT; ++b) {
              // (the one-float-per-LDS form hit a 4-way bank conflict between the gates: 20 us per step)
              const float4 g4 = *reinterpret_cast<const float4*>(&s_dg[b][q * BP_ST + j0 + u]);
              acc[b] = fmaf(g4.x, w[u], acc[b]);
              acc[b] = fmaf(g4.y, w[u + 1], acc[b]);
              acc[b] = fmaf(g4.z, w[u + 2], acc[b]);
              acc[b] = fmaf(g4.w, w[u + 3], acc[b]);
            }
          }
        }
      }
#pragma unroll
      for (int b = 0; b < BP_BT; ++b) {
        acc[b] += __shfl_xor_sync(0xffffffffu, acc[b], 1);
        acc[b] += __shfl_xor_sync(0xffffffffu, acc[b], 2);
      }
      if (p2 && q == 0) {
#pragma unroll
        for (int b = 0; b < BP_BT; ++b) s_dh[b][k] = acc[b];
      }
    }
    __syncthreads();
  }
}

// Tensor-core form of the same backward recurrence — the kernel that runs.  Mirror image of
// lstm_recurrence_mma_kernel (lstm.cu): per step  dh_rec^T [H x 8] = W_hh^T [H x 4H] · d_pre^T [4H x 8]  on
// mma.sync.m16n8k16 with W_hh^T as the A operand.  8 warps, 8 sequences per CTA (= N).  Warp w owns units
// 16w .. 16w+15 (one 16-row tile); the K dimension is the four gates, each padded to 128 rows (32 k-steps).
// The hi halves of the thread's A fragments live in registers for the whole sequence (128 registers), the lo
// halves of BPM_KLO k-steps too, the others in shared memory in fragment order.  In the accumulator layout the
// thread holds dh_rec of units (16w+g, 16w+8+g) for sequences (2c, 2c+1): exactly the four cells whose gate
// gradients it computes next — no shuffle, no shared-memory round trip for dh.  Four accumulator chains
// (k-step mod 4) keep dependent HMMAs apart.  d_pre is published as bf16 hi | lo rows [sequence][gate*128+unit]
// (double buffered, ldmatrix.x4 as the B operand); the three product terms hi·hi + hi·lo + lo·hi give ~2^-16.
// The saved activations (i, f, g, o, c) and dout of a thread's cells are staged two steps ahead by the
// thread itself with 4-byte cp.async into a private slot of a 3-deep ring (c of the earlier time step — the
// c_prev of the forget-gate gradient — is then already there), so no global latency sits inside a step.
// Measured: profiles/r02_lstm_microbench.txt.
constexpr int BPM_THREADS = 256;
constexpr int BPM_BT = 8;
constexpr int BPM_KS = 32;            // k-steps: 4 gates x 128 padded units / 16
constexpr int BPM_KLO = 2;            // k-steps whose lo fragments stay in registers
constexpr int BPM_DS = 520;           // bf16 pitch of a d_pre row (1040 B: ldmatrix rows on distinct banks)
constexpr int BPM_SLOTS = 24;         // staged floats per thread and step: 4 cells x (i, f, g, o, c, dout)
constexpr size_t BPM_SMEM = static_cast<size_t>(BPM_KS - BPM_KLO) * BPM_THREADS * 16 +
                            2 * 2 * BPM_BT * BPM_DS * 2 + 3 * BPM_SLOTS * BPM_THREADS * 4;

__device__ __forceinline__ void cp_async4_zfill(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}

__global__ void __launch_bounds__(BPM_THREADS, 1)
lstm_bptt_mma_kernel(const float* __restrict__ gates, long long gates_pitch, const float* __restrict__ w_hh,
                     const float* __restrict__ dout, long long dout_pitch, float* __restrict__ dxg,
                     long long dxg_pitch, int B, int L, int H) {
  extern __shared__ __align__(128) unsigned char bpm_smem[];
  uint4* s_wlo = reinterpret_cast<uint4*>(bpm_smem);                                         // [KS-KLO][256]
  __nv_bfloat16* s_dg = reinterpret_cast<__nv_bfloat16*>(bpm_smem + (BPM_KS - BPM_KLO) * BPM_THREADS * 16);  // [2][hi|lo][8][DS]
  float* s_st = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(s_dg) + 2 * 2 * BPM_BT * BPM_DS * 2);  // [3][SLOTS][256]
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, c = lane & 3;
  const int dir = blockIdx.y;
  const int b0 = blockIdx.x * BPM_BT;
  const float* W = w_hh + static_cast<long long>(dir) * 4 * H * H;

  // A = W_hh^T: A[unit u][kk = q*128 + j] = W_hh[q*H + j][u]
  uint32_t whi[BPM_KS][4];
  uint32_t wlo[BPM_KLO][4];
  {
    const int u0 = 16 * w + g, u1 = u0 + 8;
#pragma unroll
    for (int ks = 0; ks < BPM_KS; ++ks) {
      uint32_t lo4[4];
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int kk = 16 * ks + 8 * hf + 2 * c;
        const int q = kk >> 7, j = kk & 127;
        const float* r = W + (static_cast<long long>(q) * H + j) * H;
        const float v00 = (j < H && u0 < H) ? __ldg(r + u0) : 0.f;
        const float v01 = (j + 1 < H && u0 < H) ? __ldg(r + H + u0) : 0.f;
        const float v10 = (j < H && u1 < H) ? __ldg(r + u1) : 0.f;
        const float v11 = (j + 1 < H && u1 < H) ? __ldg(r + H + u1) : 0.f;
        const uint32_t h0 = pack_bf16x2(v00, v01), h1 = pack_bf16x2(v10, v11);
        whi[ks][2 * hf] = h0;
        whi[ks][2 * hf + 1] = h1;
        lo4[2 * hf] = pack_bf16x2(v00 - bf16_lo(h0), v01 - bf16_hi(h0));
        lo4[2 * hf + 1] = pack_bf16x2(v10 - bf16_lo(h1), v11 - bf16_hi(h1));
      }
      if (ks < BPM_KLO) {
#pragma unroll
        for (int i = 0; i < 4; ++i) wlo[ks < BPM_KLO ? ks : 0][i] = lo4[i];
      } else {
        s_wlo[(ks - BPM_KLO) * BPM_THREADS + tid] = make_uint4(lo4[0], lo4[1], lo4[2], lo4[3]);
      }
    }
  }
  {
    uint32_t* z = reinterpret_cast<uint32_t*>(s_dg);
    for (int i = tid; i < 2 * 2 * BPM_BT * BPM_DS / 2; i += BPM_THREADS) z[i] = 0u;
  }
  __syncthreads();

  // the thread's four cells: (unit 16w + 8ug + g, sequence 2c + e); staged slot v = (ug*2 + e)*6 + {i,f,g,o,c,dout}
  long long rowL[2];   // (sequence of column e) * L, 0 when the sequence is outside the batch
  int uu[2];           // unit of group ug, 0 when beyond H
  bool ok[2][2];
#pragma unroll
  for (int e = 0; e < 2; ++e) rowL[e] = static_cast<long long>(b0 + 2 * c + e < B ? b0 + 2 * c + e : 0) * L;
#pragma unroll
  for (int ug = 0; ug < 2; ++ug) {
    const int unit = 16 * w + 8 * ug + g;
    uu[ug] = unit < H ? unit : 0;
#pragma unroll
    for (int e = 0; e < 2; ++e) ok[ug][e] = unit < H && b0 + 2 * c + e < B;
  }
  const uint32_t st_base = smem_u32(s_st) + tid * 4;
  auto stage = [&](int it) {   // iteration it handles step s = L-1-it, i.e. time index tt
    if (it < L) {
      const int s = L - 1 - it;
      const long long tt = dir == 0 ? s : (L - 1 - s);
      const uint32_t dst = st_base + (it % 3) * (BPM_SLOTS * BPM_THREADS * 4);
#pragma unroll
      for (int ug = 0; ug < 2; ++ug)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float* gr = gates + (rowL[e] + tt) * gates_pitch + dir * 5 * H + uu[ug];
          const int v0 = (ug * 2 + e) * 6;
#pragma unroll
          for (int q = 0; q < 5; ++q)
            cp_async4_zfill(dst + (v0 + q) * (BPM_THREADS * 4), gr + q * H, ok[ug][e]);
          cp_async4_zfill(dst + (v0 + 5) * (BPM_THREADS * 4), dout + (rowL[e] + tt) * dout_pitch + dir * H + uu[ug],
                          ok[ug][e]);
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  stage(0);
  stage(1);

  float dh_rec[2][2], dc_carry[2][2];
#pragma unroll
  for (int ug = 0; ug < 2; ++ug) dh_rec[ug][0] = dh_rec[ug][1] = dc_carry[ug][0] = dc_carry[ug][1] = 0.f;
  const uint32_t dg_base = smem_u32(s_dg);
  const uint32_t lane_off = static_cast<uint32_t>(((lane & 7) * BPM_DS + 8 * (lane >> 3)) * 2);
  constexpr uint32_t DBUF = 2 * BPM_BT * BPM_DS * 2;   // bytes per buffer (hi rows then lo rows)
  constexpr uint32_t DLO = BPM_BT * BPM_DS * 2;
  constexpr float L2E = 1.4426950408889634f;
  for (int it = 0; it < L; ++it) {
    const int s = L - 1 - it;
    const long long tt = dir == 0 ? s : (L - 1 - s);
    stage(it + 2);
    asm volatile("cp.async.wait_group 1;" ::: "memory");   // the groups of iterations it and it+1 have landed
    const float* cur = s_st + (it % 3) * (BPM_SLOTS * BPM_THREADS) + tid;
    const float* nxt = s_st + ((it + 1) % 3) * (BPM_SLOTS * BPM_THREADS) + tid;
    __nv_bfloat16* dgw = s_dg + (it & 1) * (2 * BPM_BT * BPM_DS);
#pragma unroll
    for (int ug = 0; ug < 2; ++ug)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int v0 = (ug * 2 + e) * 6;
        const float gi = cur[(v0 + 0) * BPM_THREADS], gf = cur[(v0 + 1) * BPM_THREADS];
        const float gg = cur[(v0 + 2) * BPM_THREADS], go = cur[(v0 + 3) * BPM_THREADS];
        const float cc = cur[(v0 + 4) * BPM_THREADS];
        const float c_prev = (s > 0) ? nxt[(v0 + 4) * BPM_THREADS] : 0.f;
        const float dh = cur[(v0 + 5) * BPM_THREADS] + dh_rec[ug][e];
        const float E = ex2_approx(fminf(-2.f * L2E * cc, 80.f));
        const float tc = (1.f - E) * rcp_approx(1.f + E);
        const float d_o = dh * tc;
        const float dc = fmaf(dh * go, 1.0f - tc * tc, dc_carry[ug][e]);
        const float dai = dc * gg * gi * (1.0f - gi);
        const float daf = dc * c_prev * gf * (1.0f - gf);
        const float dag = dc * gi * (1.0f - gg * gg);
        const float dao = d_o * go * (1.0f - go);
        dc_carry[ug][e] = dc * gf;
        if (ok[ug][e]) {
          float* xr = dxg + (rowL[e] + tt) * dxg_pitch + dir * 4 * H + uu[ug];
          xr[0] = dai;
          xr[H] = daf;
          xr[2 * H] = dag;
          xr[3 * H] = dao;
        }
        // cells outside the batch / beyond H staged zeros: their gradients are zero
        const int unit = 16 * w + 8 * ug + g, n = 2 * c + e;
        __nv_bfloat16* dr = dgw + n * BPM_DS + unit;
        const float da[4] = {dai, daf, dag, dao};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const __nv_bfloat16 hh = __float2bfloat16_rn(da[q]);
          dr[q * 128] = hh;
          dr[BPM_BT * BPM_DS + q * 128] = __float2bfloat16_rn(da[q] - __bfloat162float(hh));
        }
      }
    __syncthreads();
    if (s > 0) {   // the recurrent gradient is not needed before the first step
      float acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
      const uint32_t db = dg_base + (it & 1) * DBUF + lane_off;
#pragma unroll
      for (int kp = 0; kp < BPM_KS / 2; ++kp) {
        if (((32 * kp) & 127) < H) {   // both k-steps of a pair lie in one gate; all-padding pairs are skipped
          uint32_t bh[4], bl[4];
          ldsm_x4(db + kp * 64, bh[0], bh[1], bh[2], bh[3]);
          ldsm_x4(db + DLO + kp * 64, bl[0], bl[1], bl[2], bl[3]);
          // chains: (kp even, kk) -> 0, 1 ; (kp odd, kk) -> 2, 3, each takes its three terms two issues apart
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            const int ks = 2 * kp + kk;
            mma_bf16_16816(acc[2 * (kp & 1) + kk], whi[ks][0], whi[ks][1], whi[ks][2], whi[ks][3], bh[2 * kk], bh[2 * kk + 1]);
          }
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            const int ks = 2 * kp + kk;
            mma_bf16_16816(acc[2 * (kp & 1) + kk], whi[ks][0], whi[ks][1], whi[ks][2], whi[ks][3], bl[2 * kk], bl[2 * kk + 1]);
          }
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            const int ks = 2 * kp + kk;
            if (ks < BPM_KLO) {
              const int kr = ks < BPM_KLO ? ks : 0;
              mma_bf16_16816(acc[2 * (kp & 1) + kk], wlo[kr][0], wlo[kr][1], wlo[kr][2], wlo[kr][3], bh[2 * kk], bh[2 * kk + 1]);
            } else {
              const uint4 l4 = s_wlo[(ks - BPM_KLO) * BPM_THREADS + tid];
              mma_bf16_16816(acc[2 * (kp & 1) + kk], l4.x, l4.y, l4.z, l4.w, bh[2 * kk], bh[2 * kk + 1]);
            }
          }
        }
      }
      dh_rec[0][0] = (acc[0][0] + acc[1][0]) + (acc[2][0] + acc[3][0]);
      dh_rec[0][1] = (acc[0][1] + acc[1][1]) + (acc[2][1] + acc[3][1]);
      dh_rec[1][0] = (acc[0][2] + acc[1][2]) + (acc[2][2] + acc[3][2]);
      dh_rec[1][1] = (acc[0][3] + acc[1][3]) + (acc[2][3] + acc[3][3]);
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// ------------------------------------------------------------------------------------------ multi2one
// Training form of lstm_cell_kernel (sdnet_kernels.cu): same update with accurate expf / tanhf, also
// saving (i, f, g, o, c, h) of the step in save[r][6H].
__global__ void lstm_cell_train_kernel(const float* __restrict__ gx, const float* __restrict__ gh,
                                       float* __restrict__ c, __nv_bfloat16* __restrict__ h_split, int parts,
                                       int Kp, int H, int n_rows, const int32_t* __restrict__ last_step,
                                       int step, const long long* __restrict__ slot_off,
                                       float* __restrict__ slots, float* __restrict__ save) {
  const long long total = static_cast<long long>(n_rows) * H;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / H);
    const int j = static_cast<int>(i - static_cast<long long>(r) * H);
    const float* g = gx + static_cast<long long>(r) * 4 * H;
    float gi = g[j], gf = g[H + j], gg = g[2 * H + j], go = g[3 * H + j];
    float cp = 0.f;
    if (gh != nullptr) {
      const float* qh = gh + static_cast<long long>(r) * 4 * H;
      gi += qh[j]; gf += qh[H + j]; gg += qh[2 * H + j]; go += qh[3 * H + j];
      cp = c[i];
    }
    const float si = 1.0f / (1.0f + expf(-gi));
    const float sf = 1.0f / (1.0f + expf(-gf));
    const float so = 1.0f / (1.0f + expf(-go));
    const float tg = tanhf(gg);
    const float cn = sf * cp + si * tg;
    const float hn = so * tanhf(cn);
    c[i] = cn;
    float* sv = save + static_cast<long long>(r) * 6 * H;
    sv[j] = si; sv[H + j] = sf; sv[2 * H + j] = tg; sv[3 * H + j] = so; sv[4 * H + j] = cn; sv[5 * H + j] = hn;
    float rem = hn;
    for (int p = 0; p < parts; ++p) {
      const __nv_bfloat16 hb = __float2bfloat16_rn(rem);
      h_split[static_cast<long long>(r) * parts * Kp + static_cast<long long>(p) * Kp + j] = hb;
      rem -= __bfloat162float(hb);
    }
    if (last_step[r] == step) slots[slot_off[r] + j] = hn;
  }
}

// Backward of one step: rows r < n_rows are the items still active at `step`.
//   dh = (last_step[r] == step ? dslots[slot_off[r] + j] : 0) + (dh_rec ? dh_rec[r][j] : 0)
//   dc = dc_carry[r][j] (0 for rows that end at this step) + dh * o * (1 - tanh(c)^2)
// writes d(pre-activations) [n_rows, 4H] and the new dc_carry = dc * f.
__global__ void lstm_cell_bwd_kernel(const float* __restrict__ save, const float* __restrict__ save_prev,
                                     const float* __restrict__ dslots, const long long* __restrict__ slot_off,
                                     const int32_t* __restrict__ last_step, int step,
                                     const float* __restrict__ dh_rec, int n_rec, float* __restrict__ dc_carry,
                                     float* __restrict__ dgates, int H, int n_rows) {
  const long long total = static_cast<long long>(n_rows) * H;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / H);
    const int j = static_cast<int>(i - static_cast<long long>(r) * H);
    const float* sv = save + static_cast<long long>(r) * 6 * H;
    const float gi = sv[j], gf = sv[H + j], gg = sv[2 * H + j], go = sv[3 * H + j], c = sv[4 * H + j];
    const float c_prev = save_prev ? save_prev[static_cast<long long>(r) * 6 * H + 4 * H + j] : 0.f;
    const bool ends = last_step[r] == step;
    float dh = ends ? dslots[slot_off[r] + j] : 0.f;
    if (dh_rec != nullptr && r < n_rec) dh += dh_rec[static_cast<long long>(r) * H + j];
    const float tc = tanhf(c);
    const float dc = (ends ? 0.f : dc_carry[i]) + dh * go * (1.0f - tc * tc);
    float* dg = dgates + static_cast<long long>(r) * 4 * H;
    dg[j] = dc * gg * gi * (1.0f - gi);
    dg[H + j] = dc * c_prev * gf * (1.0f - gf);
    dg[2 * H + j] = dc * gi * (1.0f - gg * gg);
    dg[3 * H + j] = dh * tc * go * (1.0f - go);
    dc_carry[i] = dc * gf;
  }
}

}  // namespace

// =============================================================================================== C ABI
extern "C" int ruart_bmm_f32(const float* A, long long lda, long long stride_a, int trans_a, const float* B,
                             long long ldb, long long stride_b, int trans_b, float* C, long long ldc,
                             long long stride_c, int batch, int M, int N, int K, float alpha, int accumulate,
                             void* stream) {
  RUART_ARG_CHECK(A != nullptr && B != nullptr && C != nullptr && M >= 0 && N >= 0 && K >= 0 && batch >= 0);
  RUART_ARG_CHECK(batch <= 65535);
  if (batch == 0 || M == 0 || N == 0) return RUART_OK;
  dim3 grid((N + TB_N - 1) / TB_N, (M + TB_M - 1) / TB_M, batch);
  RUART_ARG_CHECK(grid.y <= 65535);
  bmm_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(A, lda, stride_a, trans_a, B, ldb, stride_b, trans_b,
                                                         C, ldc, stride_c, M, N, K, alpha, accumulate);
  RUART_LAUNCH_CHECK();
  return RUART_OK;
}

extern "C" int ruart_masked_softmax(const float* x, long long x_pitch, const uint8_t* mask, int B, int L1,
                                    int L2, float* out, long long out_pitch, void* stream) {
  RUART_ARG_CHECK(x != nullptr && out != nullptr && B >= 0 && L1 > 0 && L2 > 0);
  const long long rows = static_cast<long long>(B) * L1;
  if (rows == 0) return RUART_OK;
  masked_softmax_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
      x, x_pitch, mask, L1, L2, rows, out, out_pitch);
  RUART_LAUNCH_CHECK();
  return RUART_OK;
}

extern "C" int ruart_softmax_backward(const float* p, long long p_pitch, const float* dp, long long dp_pitch,
                                      float* dx, long long dx_pitch, long long rows, int cols, void* stream) {
  RUART_ARG_CHECK(p != nullptr && dp != nullptr && dx != nullptr && rows >= 0 && cols > 0);
  if (rows == 0) return RUART_OK;
  softmax_backward_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
      p, p_pitch, dp, dp_pitch, dx, dx_pitch, rows, cols);
  RUART_LAUNCH_CHECK();
  return RUART_OK;
}

extern "C" int ruart_eltwise(int op, const float* a, long long a_pitch, const float* b, long long b_pitch,
                             const float* v, int v_len, float* out, long long out_pitch, long long rows,
                             int cols, void* stream) {
  RUART_ARG_CHECK(a != nullptr && out != nullptr && rows >= 0 && cols > 0 && op >= 0 && op <= 5);
  if (op == 3 || op == 5) RUART_ARG_CHECK(v != nullptr && v_len >= 1 && (v_len == 1 || v_len >= cols));
  if (op <= 2) RUART_ARG_CHECK(b != nullptr);
  if (rows == 0) return RUART_OK;
  const unsigned grid = grid_for(rows * cols, 256 * 4);
  cudaStream_t st = (cudaStream_t)stream;
  switch (op) {
    case 0: eltwise_kernel<0><<<grid, 256, 0, st>>>(a, a_pitch, b, b_pitch, v, v_len, out, out_pitch, rows, cols); break;
    case 1: eltwise_kernel<1><<<grid, 256, 0, st>>>(a, a_pitch, b, b_pitch, v, v_len, out, out_pitch, rows, cols); break;
    case 2: eltwise_kernel<2><<<grid, 256, 0, st>>>(a, a_pitch, b, b_pitch, v, v_len, out, out_pitch, rows, cols); break;
    case 3: eltwise_kernel<3><<<grid, 256, 0, st>>>(a, a_pitch, b, b_pitch, v, v_len, out, out_pitch, rows, cols); break;
    case 4: eltwise_kernel<4><<<grid, 256, 0, st>>>(a, a_pitch, b, b_pitch, v, v_len, out, out_pitch, rows, cols); break;
    default: eltwise_kernel<5><<<grid, 256, 0, st>>>(a, a_pitch, b, b_pitch, v, v_len, out, out_pitch, rows, cols); break;
  }
  RUART_LAUNCH_CHECK();
  return RUART_OK;
}

extern "C" int ruart_mask_fill(const float* x, const uint8_t* mask, long long n, float value, float* out,
                               void* stream) {
  RUART_ARG_CHECK(x != nullptr && mask != nullptr && out != nullptr && n >= 0);
  if (n == 0) return RUART_OK;
  mask_fill_kernel<<<grid_for(n, 256 * 4), 256, 0, (cudaStream_t)stream>>>(x, mask, n, value, out);
  RUART_LAUNCH_CHECK();
  return RUART_OK;
}

extern "C" int ruart_colsum(const float* x, long long pitch, long long rows, int cols, double* workspace,
                            float* out, int accumulate, void* stream) {
  RUART_ARG_CHECK(x != nullptr && out != nullptr && workspace != nullptr && rows >= 0 && cols > 0);
  cudaStream_t st = (cudaStream_t)stream;
  int chunks = static_cast<int>((rows + 255) / 256);
  if (chunks > CS_CHUNKS) chunks = CS_CHUNKS;
  if (chunks < 1) chunks = 1;
  dim3 grid((cols + 31) / 32, chunks);
  colsum_partial_kernel<<<grid, 256, 0, st>>>(x, pitch, rows, cols, workspace);
  RUART_LAUNCH_CHECK();
  colsum_final_kernel<<<(cols + 255) / 256, 256, 0, st>>>(workspace, chunks, cols, out, accumulate);
  RUART_LAUNCH_CHECK();
  return RUART_OK;
}

extern "C" int ruart_split_bf16_t(const float* src, long long ld, long long rows, int K, long long rows_p,
                                  int parts, void* dst, void* stream) {
  RUART_ARG_CHECK(src != nullptr && dst != nullptr && rows >= 0 && K > 0 && rows_p >= rows &&
                  (rows_p % 64) == 0 && parts >= 1 && parts <= 3);
  if (rows_p == 0) return RUART_OK;
  dim3 grid(static_cast<unsigned>(rows_p / 32), (K + 31) / 32);
  RUART_ARG_CHECK(grid.y <= 65535);
  split_bf16_t_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, ld, rows, K, rows_p, parts,
                                                              (__nv_bfloat16*)dst);
  RUART_LAUNCH_CHECK();
  return RUART_OK;
}

extern "C" int ruart_whole_layernorm_backward(const float* y, long long y_pitch, const float* dy,
                                              long long dy_pitch, long long rows, int cols,
                                              const float* stats, double* workspace, float* dx,
                                              long long dx_pitch, void* stream) {
  RUART_ARG_CHECK(y != nullptr && dy != nullptr && dx != nullptr && stats != nullptr && workspace != nullptr);
  RUART_ARG_CHECK(rows > 0 && cols > 0);
  cudaStream_t st = (cudaStream_t)stream;
  long long g = (rows * cols + 256 * 8 - 1) / (256 * 8);
  if (g > 1024) g = 1024;
  if (g < 1) g = 1;
  whole_ln_bwd_stats_kernel<<<static_cast<unsigned>(g), 256, 0, st>>>(y, y_pitch, dy, dy_pitch, rows, cols,
                                                                     workspace);
  RUART_LAUNCH_CHECK();
  whole_ln_bwd_apply_kernel<<<grid_for(rows * cols, 256 * 4), 256, 0, st>>>(
      y, y_pitch, dy, dy_pitch, rows, cols, workspace, static_cast<int>(g), stats, dx, dx_pitch);
  RUART_LAUNCH_CHECK();
  return RUART_OK;
}

extern "C" long long ruart_embedding_grad_workspace_bytes(long long n, int V, int D) {
  if (n < 0 || V <= 0 || D <= 0) return -1;
  const long long slabs = (n + EG_SLAB - 1) / EG_SLAB;
  return eg_round16(n) + eg_round16((2LL * V + 1 + 2 * n) * 4) + 2 * slabs * D * 4;
}

template <typename I>
static int embedding_grad_sorted(const I* ids, long long n, const float* dy, long long dy_pitch, int D, int V,
                                 uint8_t* workspace, float* dW, long long dw_pitch, int accumulate,
                                 cudaStream_t st) {
  uint8_t* flag = workspace;
  int* cnt = reinterpret_cast<int*>(workspace + eg_round16(n));
  int* start = cnt + V;
  int* pos = start + V + 1;
  int* sorted = pos + n;
  float* partial = reinterpret_cast<float*>(workspace + eg_round16(n) + eg_round16((2LL * V + 1 + 2 * n) * 4));
  RUART_CUDA_CHECK(cudaMemsetAsync(cnt, 0, static_cast<size_t>(V) * 4, st));
  eg_flag_count_kernel<I><<<grid_for(n * 32, 256), 256, 0, st>>>(dy, dy_pitch, n, D, ids, V, flag, cnt);
  RUART_LAUNCH_CHECK();
  eg_scan_kernel<<<1, 1024, 0, st>>>(cnt, start, V);
  RUART_LAUNCH_CHECK();
  eg_fill_kernel<I><<<grid_for(n, 256), 256, 0, st>>>(ids, flag, n, start, cnt, pos);
  RUART_LAUNCH_CHECK();
  eg_rank_kernel<I><<<grid_for(n, 256), 256, 0, st>>>(ids, start, V, pos, sorted);
  RUART_LAUNCH_CHECK();
  const long long slabs = (n + EG_SLAB - 1) / EG_SLAB;
  eg_slab_kernel<I><<<grid_for(slabs * 32, 256), 256, 0, st>>>(ids, start, V, sorted, dy, dy_pitch, D, partial, dW,
                                                               dw_pitch, accumulate);
  RUART_LAUNCH_CHECK();
  eg_combine_kernel<<<grid_for(static_cast<long long>(V) * 32, 256), 256, 0, st>>>(start, V, partial, D, dW,
                                                                                   dw_pitch, accumulate);
  RUART_LAUNCH_CHECK();
  return RUART_OK;
}

extern "C" int ruart_embedding_grad(const void* ids, int idx_is_64, long long n, const float* dy,
                                    long long dy_pitch, int D, int V, uint8_t* workspace,
                                    long long workspace_bytes, float* dW, long long dw_pitch, int accumulate,
                                    void* stream) {
  RUART_ARG_CHECK(ids != nullptr && dy != nullptr && dW != nullptr && n >= 0 && D > 0 && V > 0);
  const long long flag_bytes = (n + 15) / 16 * 16;
  RUART_ARG_CHECK(workspace != nullptr && workspace_bytes >= flag_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  // sorted form whenever the caller's workspace holds it (ruart_embedding_grad_workspace_bytes); the
  // one-warp-per-row scan below stays for smaller workspaces and as the A/B aid RUART_EMBGRAD_SCAN
  static const bool scan_form = getenv("RUART_EMBGRAD_SCAN") != nullptr;
  if (!scan_form && n > 0 && n < (1LL << 31) - EG_SLAB && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0 &&
      workspace_bytes >= ruart_embedding_grad_workspace_bytes(n, V, D)) {
    return idx_is_64 ? embedding_grad_sorted<long long>((const long long*)ids, n, dy, dy_pitch, D, V, workspace, dW,
                                                        dw_pitch, accumulate, st)
                     : embedding_grad_sorted<int32_t>((const int32_t*)ids, n, dy, dy_pitch, D, V, workspace, dW,
                                                      dw_pitch, accumulate, st);
  }
  if (n > 0) {
    row_nonzero_kernel<<<grid_for(n * 32, 256), 256, 0, st>>>(dy, dy_pitch, n, D, workspace);
    RUART_LAUNCH_CHECK();
  }
  // small tables: split the id list into segments so that ~4096 warps have work (50 warps walking 128 k ids
  // serially cost 4 ms per call); as many segments as the caller's workspace holds partial sums for
  int n_seg = 1;
  if (V < 2048) {
    n_seg = (4096 + V - 1) / V;
    if (n_seg > 64) n_seg = 64;
    const long long room = (workspace_bytes - flag_bytes) / (static_cast<long long>(V) * D * 4);
    if (n_seg > room) n_seg = static_cast<int>(room);
    if (n_seg < 1) n_seg = 1;
  }
  float* partial = reinterpret_cast<float*>(workspace + flag_bytes);
  const unsigned grid = grid_for(static_cast<long long>(V) * n_seg * 32, 256);
  if (idx_is_64)
    embedding_grad_kernel<long long><<<grid, 256, 0, st>>>((const long long*)ids, workspace, n, dy, dy_pitch, D,
                                                           V, n_seg, partial, dW, dw_pitch, accumulate);
  else
    embedding_grad_kernel<int32_t><<<grid, 256, 0, st>>>((const int32_t*)ids, workspace, n, dy, dy_pitch, D, V,
                                                         n_seg, partial, dW, dw_pitch, accumulate);
  RUART_LAUNCH_CHECK();
  if (n_seg > 1) {
    embedding_grad_reduce_kernel<<<grid_for(static_cast<long long>(V) * D, 256), 256, 0, st>>>(
        partial, n_seg, V, D, dW, dw_pitch, accumulate);
    RUART_LAUNCH_CHECK();
  }
  return RUART_OK;
}

extern "C" int ruart_subword_layers_backward(const float* h_f32, const void* h_bf16, long long layer_stride,
                                             const int32_t* words, int n_words, const int32_t* row_start,
                                             const uint8_t* x_mask, int W, const float* dy,
                                             long long dy_stride, const float* alpha, int n_layers,
                                             const float* gamma, int hidden, double* workspace,
                                             float* dalpha, float* dgamma, int accumulate, void* stream) {
  RUART_ARG_CHECK((h_f32 != nullptr) != (h_bf16 != nullptr));
  RUART_ARG_CHECK(n_layers >= 1 && n_layers <= SW_MAX_LAYERS && workspace != nullptr);
  RUART_ARG_CHECK(alpha != nullptr && gamma != nullptr && dalpha != nullptr && dgamma != nullptr);
  cudaStream_t st = (cudaStream_t)stream;
  static const bool scalar_form = getenv("RUART_SUBWORD_BWD_SCALAR") != nullptr;   // A/B aid
  int blocks = (n_words + 7) / 8;
  if (blocks > 256) blocks = 256;
  if (blocks < 1) blocks = 1;
  if (h_f32 != nullptr)
    subword_layers_bwd_kernel<float><<<blocks, 256, 0, st>>>(h_f32, layer_stride, words, n_words, row_start,
                                                             x_mask, W, dy, dy_stride, n_layers, hidden, workspace);
  else if (hidden % 256 == 0 && hidden <= 256 * SWV_MAX && dy_stride % 4 == 0 &&
           (reinterpret_cast<uintptr_t>(dy) & 15) == 0 && (reinterpret_cast<uintptr_t>(h_bf16) & 15) == 0 &&
           layer_stride % 8 == 0 && !scalar_form)
    subword_layers_bwd_bf16v_kernel<<<blocks, 256, 0, st>>>((const __nv_bfloat16*)h_bf16, layer_stride, words, n_words,
                                                            row_start, x_mask, W, dy, dy_stride, n_layers, hidden,
                                                            workspace);
  else
    subword_layers_bwd_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(
        (const __nv_bfloat16*)h_bf16, layer_stride, words, n_words, row_start, x_mask, W, dy, dy_stride,
        n_layers, hidden, workspace);
  RUART_LAUNCH_CHECK();
  layer_mix_bwd_kernel<<<1, 32, 0, st>>>(workspace, blocks, n_layers, alpha, gamma, dalpha, dgamma, accumulate);
  RUART_LAUNCH_CHECK();
  return RUART_OK;
}

extern "C" int ruart_lstm_recurrence_backward(const float* gates, long long gates_pitch, const float* w_hh,
                                              const float* dout, long long dout_pitch, float* dxg,
                                              long long dxg_pitch, int B, int L, int H, int ndir,
                                              void* stream) {
  RUART_ARG_CHECK(gates != nullptr && w_hh != nullptr && dout != nullptr && dxg != nullptr);
  RUART_ARG_CHECK(B > 0 && L > 0 && H > 0 && H <= BP_HP && BP_BT * H <= BP_THREADS && (ndir == 1 || ndir == 2));
  static const bool fma_form = getenv("RUART_LSTM_BPTT_FMA") != nullptr;   // A/B aid: the fp32 FMA kernel
  if (!fma_form) {
    static RuartDeviceOnce attr_set;
    if (!attr_set.done()) {
      RUART_CUDA_CHECK(cudaFuncSetAttribute(lstm_bptt_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)BPM_SMEM));
      attr_set.set();
    }
    dim3 grid((B + BPM_BT - 1) / BPM_BT, ndir);
    lstm_bptt_mma_kernel<<<grid, BPM_THREADS, BPM_SMEM, (cudaStream_t)stream>>>(gates, gates_pitch, w_hh, dout,
                                                                               dout_pitch, dxg, dxg_pitch, B, L, H);
    RUART_LAUNCH_CHECK();
    return RUART_OK;
  }
  dim3 grid((B + BP_BT - 1) / BP_BT, ndir);
  lstm_bptt_kernel<<<grid, BP_THREADS, 0, (cudaStream_t)stream>>>(gates, gates_pitch, w_hh, dout, dout_pitch,
                                                                 dxg, dxg_pitch, B, L, H);
  RUART_LAUNCH_CHECK();
  return RUART_OK;
}

extern "C" int ruart_lstm_cell_train(const float* gx, const float* gh, float* c, void* h_split, int parts,
                                     int Kp, int H, int n_rows, const int32_t* last_step, int step,
                                     const long long* slot_off, float* slots, float* save, void* stream) {
  RUART_ARG_CHECK(H > 0 && Kp >= H && (Kp % 64) == 0 && parts >= 1 && parts <= 3 && save != nullptr);
  if (n_rows == 0) return RUART_OK;
  lstm_cell_train_kernel<<<grid_for(static_cast<long long>(n_rows) * H, 256), 256, 0, (cudaStream_t)stream>>>(
      gx, gh, c, (__nv_bfloat16*)h_split, parts, Kp, H, n_rows, last_step, step, slot_off, slots, save);
  RUART_LAUNCH_CHECK();
  return RUART_OK;
}

extern "C" int ruart_lstm_cell_backward(const float* save, const float* save_prev, const float* dslots,
                                        const long long* slot_off, const int32_t* last_step, int step,
                                        const float* dh_rec, int n_rec, float* dc_carry, float* dgates, int H,
                                        int n_rows, void* stream) {
  RUART_ARG_CHECK(save != nullptr && dslots != nullptr && dc_carry != nullptr && dgates != nullptr && H > 0);
  if (n_rows == 0) return RUART_OK;
  lstm_cell_bwd_kernel<<<grid_for(static_cast<long long>(n_rows) * H, 256), 256, 0, (cudaStream_t)stream>>>(
      save, save_prev, dslots, slot_off, last_step, step, dh_rec, n_rec, dc_carry, dgates, H, n_rows);
  RUART_LAUNCH_CHECK();
  return RUART_OK;
}
