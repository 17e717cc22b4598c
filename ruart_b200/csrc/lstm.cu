// Persistent (Bi)LSTM recurrence for sm_100a — the sequential half of torch.nn.LSTM as used by
// StackedBRNN.forward (reference Models/Layers.py:137,166): hidden size <= 128 (125 in the shipped
// conf), batch_first, both directions in one launch, pads included (the reference never packs).
//
// The input projection  xg = x W_ih^T + b_ih + b_hh  for every step comes from the tcgen05 GEMM.
// Here one CTA owns BT sequences of one direction for ALL steps, so no inter-CTA synchronisation
// exists: thread t owns gate row (unit j = t/4, gate g = t%4 in torch order i,f,g,o) of W_hh.
// W_hh (4H x H fp32 = 250 KB for H = 125) does not fit shared memory, so each thread keeps the
// first KR = 64 weights of its row in REGISTERS and the remaining H-64 columns live in shared
// memory, transposed ([k][row], conflict-free).  h_{t-1} of the BT sequences sits in shared memory
// and is read as broadcast float4.  The four gates of a unit are in adjacent lanes and meet through
// warp shuffles; the g == 0 lane keeps c in registers and publishes h.
//
// lstm_recurrence2_kernel (the kernel that runs) register-blocks two gate rows per
// thread: 256 threads, thread = (unit j, half); half 0 owns gates (i, g), half 1 owns (f, o).  Every
// broadcast h load now feeds four FFMA2 instead of two, and 88 weights of each row sit in
// registers, so the shared-memory loads per FMA halve: 3.75 -> 2.61 us per time step at B = 256,
// H = 125 (profiles/r01_lstm_microbench.txt).
#include <cstdint>
#include <cstdlib>

#include "common.cuh"
#include "ruart_b200.h"

namespace {

using namespace ruart;

constexpr int LSTM_THREADS = 512;
constexpr int KR = 64;     // weights per row kept in registers (as 32 float2)
constexpr int HP = 128;    // padded hidden size (h rows in smem, zero padded)
constexpr int ROWP = 512;  // padded gate-row count (threads)
constexpr int NQ = (HP - KR) / 4;  // float4 weight quads per row kept in smem

// Packed fp32x2 FMA (Blackwell FFMA2): d.x = a.x*b.x + c.x, d.y = a.y*b.y + c.y.  Measured on B200
// (profiles/r01_lstm_microbench.txt): FFMA2 issues every 3.7 clk per sub-partition (69 FMA/clk/SM)
// and scalar FFMA every 1.0 clk (128 FMA/clk/SM), yet this kernel is faster with FFMA2 (261 vs
// 290 us for 100 steps): with two warps per scheduler it is bound by issue slots and by the
// LDS.128 broadcasts of h, not by the FMA pipe, and FFMA2 halves the issued instructions.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(d)
      : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)),
        "l"(*reinterpret_cast<unsigned long long*>(&c)));
  return *reinterpret_cast<float2*>(&d);
}

// 1 / (1 + e^-x) = rcp(1 + ex2(-x log2 e)); saturates cleanly (ex2 -> inf -> rcp -> 0).
__device__ __forceinline__ float logistic(float x) {
  return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x));
}

template <int BT>
__global__ void __launch_bounds__(LSTM_THREADS, 1)
lstm_recurrence_kernel(const float* __restrict__ xg, long long xg_pitch,  // [B*L, ndir*4H]
                       const float* __restrict__ w_hh,                    // [ndir][4H][H]
                       float* __restrict__ out, long long out_pitch,      // [B*L, >= ndir*H]
                       int B, int L, int H) {
  extern __shared__ __align__(16) float smem[];
  float4* s_w = reinterpret_cast<float4*>(smem);   // [NQ][ROWP] quads: k = KR + 4q .. +3 of row t
  float* s_h = smem + NQ * ROWP * 4;               // [2][BT][HP]
  const int t = threadIdx.x;
  const int dir = blockIdx.y;
  const int b0 = blockIdx.x * BT;
  const int rows = 4 * H;
  const bool active = t < rows;
  const int j = t >> 2, g = t & 3;
  const int wrow = g * H + j;  // row of W_hh / column of xg for this thread
  const float* W = w_hh + static_cast<long long>(dir) * rows * H;

  float2 wreg[KR / 2];
#pragma unroll
  for (int k = 0; k < KR; k += 2) {
    wreg[k / 2].x = (active && k < H) ? W[static_cast<long long>(wrow) * H + k] : 0.f;
    wreg[k / 2].y = (active && k + 1 < H) ? W[static_cast<long long>(wrow) * H + k + 1] : 0.f;
  }
  for (int q = 0; q < NQ; ++q) {
    float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const int k = KR + 4 * q;
    if (active) {
      const float* wr = W + static_cast<long long>(wrow) * H;
      if (k < H) w4.x = wr[k];
      if (k + 1 < H) w4.y = wr[k + 1];
      if (k + 2 < H) w4.z = wr[k + 2];
      if (k + 3 < H) w4.w = wr[k + 3];
    }
    s_w[q * ROWP + t] = w4;
  }
  for (int i = t; i < 2 * BT * HP; i += LSTM_THREADS) s_h[i] = 0.f;
  __syncthreads();
  const int nq = (H > KR) ? (H - KR + 3) / 4 : 0;

  float c[BT];
  float nxt[BT];
#pragma unroll
  for (int b = 0; b < BT; ++b) c[b] = 0.f;
  const long long xcol = static_cast<long long>(dir) * rows + wrow;
  auto step_time = [&](int s) { return dir == 0 ? s : (L - 1 - s); };
#pragma unroll
  for (int b = 0; b < BT; ++b) {
    const int bb = b0 + b;
    nxt[b] = (active && bb < B)
                 ? __ldg(xg + (static_cast<long long>(bb) * L + step_time(0)) * xg_pitch + xcol)
                 : 0.f;
  }
  int cur = 0;
  for (int s = 0; s < L; ++s) {
    float2 acc[BT];
#pragma unroll
    for (int b = 0; b < BT; ++b) acc[b] = make_float2(nxt[b], 0.f);
    if (s + 1 < L) {
#pragma unroll
      for (int b = 0; b < BT; ++b) {
        const int bb = b0 + b;
        nxt[b] = (active && bb < B)
                     ? __ldg(xg + (static_cast<long long>(bb) * L + step_time(s + 1)) * xg_pitch + xcol)
                     : 0.f;
      }
    }
    const float* hc = s_h + cur * BT * HP;
#pragma unroll
    for (int k4 = 0; k4 < KR / 4; ++k4) {
#pragma unroll
      for (int b = 0; b < BT; ++b) {
        const float4 hv = *reinterpret_cast<const float4*>(hc + b * HP + 4 * k4);
        acc[b] = ffma2(wreg[2 * k4], make_float2(hv.x, hv.y), acc[b]);
        acc[b] = ffma2(wreg[2 * k4 + 1], make_float2(hv.z, hv.w), acc[b]);
      }
    }
    for (int q = 0; q < nq; ++q) {
      const float4 wv = s_w[q * ROWP + t];
#pragma unroll
      for (int b = 0; b < BT; ++b) {
        const float4 hv = *reinterpret_cast<const float4*>(hc + b * HP + KR + 4 * q);
        acc[b] = ffma2(make_float2(wv.x, wv.y), make_float2(hv.x, hv.y), acc[b]);
        acc[b] = ffma2(make_float2(wv.z, wv.w), make_float2(hv.z, hv.w), acc[b]);
      }
    }
    float* hn = s_h + (cur ^ 1) * BT * HP;
    const int tt = step_time(s);
#pragma unroll
    for (int b = 0; b < BT; ++b) {
      // every lane activates its own gate with one logistic evaluation on the MUFU ex2/rcp
      // approximations (<= 2 ulp each): i, f, o -> s(x) ; g -> tanh(x) = 2 s(2x) - 1
      const float pre = acc[b].x + acc[b].y;
      const float sg = logistic(g == 2 ? 2.0f * pre : pre);
      const float av = (g == 2) ? fmaf(2.0f, sg, -1.0f) : sg;
      const float a_f = __shfl_down_sync(0xffffffffu, av, 1);
      const float a_g = __shfl_down_sync(0xffffffffu, av, 2);
      const float a_o = __shfl_down_sync(0xffffffffu, av, 3);
      if (g == 0 && active) {
        const float cn = fmaf(a_f, c[b], av * a_g);
        const float hv = a_o * fmaf(2.0f, logistic(2.0f * cn), -1.0f);
        c[b] = cn;
        hn[b * HP + j] = hv;
        const int bb = b0 + b;
        if (bb < B) out[(static_cast<long long>(bb) * L + tt) * out_pitch + dir * H + j] = hv;
      }
    }
    __syncthreads();
    cur ^= 1;
  }
}


// ----------------------------------------------------------------------------------------------
// Two gate rows per thread (see the header).  KR2 weights of each row in registers, the remaining
// HP-KR2 columns as float4 quads in shared memory: s_w[(q*2 + r) * 256 + t].
//
// SAVE (training): the activated gates and the cell state of every step are also
// written to `gates` [B*L, ndir*5*H] (row pitch gates_pitch; columns dir*5H + {i,f,g,o,c}*H + j) —
// what back-propagation through time (lstm_bptt_kernel, backward_kernels.cu) reads.
constexpr int LSTM2_THREADS = 256;
template <int BT, int KR2, bool SAVE = false>
__global__ void __launch_bounds__(LSTM2_THREADS, 1)
lstm_recurrence2_kernel(const float* __restrict__ xg, long long xg_pitch,  // [B*L, ndir*4H]
                        const float* __restrict__ w_hh,                    // [ndir][4H][H]
                        float* __restrict__ out, long long out_pitch,      // [B*L, >= ndir*H]
                        int B, int L, int H, float* __restrict__ gates = nullptr,
                        long long gates_pitch = 0) {
  constexpr int NQ2 = (HP - KR2) / 4;
  extern __shared__ __align__(16) float smem[];
  float4* s_w = reinterpret_cast<float4*>(smem);        // [NQ2][2][256] quads
  float* s_h = smem + NQ2 * 2 * LSTM2_THREADS * 4;      // [2][BT][HP]
  const int t = threadIdx.x;
  const int dir = blockIdx.y;
  const int b0 = blockIdx.x * BT;
  const int rows = 4 * H;
  const int j = t >> 1, half = t & 1;
  const bool active = j < H;
  // torch gate order i, f, g, o: half 0 -> (i, g), half 1 -> (f, o)
  const int wrow0 = half * H + j;          // i or f
  const int wrow1 = (2 + half) * H + j;    // g or o
  const float* W = w_hh + static_cast<long long>(dir) * rows * H;
  const float* wr0 = W + static_cast<long long>(wrow0) * H;
  const float* wr1 = W + static_cast<long long>(wrow1) * H;

  float2 w0[KR2 / 2], w1[KR2 / 2];
#pragma unroll
  for (int k = 0; k < KR2; k += 2) {
    w0[k / 2].x = (active && k < H) ? wr0[k] : 0.f;
    w0[k / 2].y = (active && k + 1 < H) ? wr0[k + 1] : 0.f;
    w1[k / 2].x = (active && k < H) ? wr1[k] : 0.f;
    w1[k / 2].y = (active && k + 1 < H) ? wr1[k + 1] : 0.f;
  }
  for (int q = 0; q < NQ2; ++q) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), c4 = a;
    const int k = KR2 + 4 * q;
    if (active) {
      if (k < H) { a.x = wr0[k]; c4.x = wr1[k]; }
      if (k + 1 < H) { a.y = wr0[k + 1]; c4.y = wr1[k + 1]; }
      if (k + 2 < H) { a.z = wr0[k + 2]; c4.z = wr1[k + 2]; }
      if (k + 3 < H) { a.w = wr0[k + 3]; c4.w = wr1[k + 3]; }
    }
    s_w[(q * 2 + 0) * LSTM2_THREADS + t] = a;
    s_w[(q * 2 + 1) * LSTM2_THREADS + t] = c4;
  }
  for (int i = t; i < 2 * BT * HP; i += LSTM2_THREADS) s_h[i] = 0.f;
  __syncthreads();
  const int nq = (H > KR2) ? (H - KR2 + 3) / 4 : 0;

  float c[BT], nx0[BT], nx1[BT];
#pragma unroll
  for (int b = 0; b < BT; ++b) c[b] = 0.f;
  const long long xc0 = static_cast<long long>(dir) * rows + wrow0;
  const long long xc1 = static_cast<long long>(dir) * rows + wrow1;
  auto step_time = [&](int s) { return dir == 0 ? s : (L - 1 - s); };
  auto fetch = [&](int s) {
#pragma unroll
    for (int b = 0; b < BT; ++b) {
      const int bb = b0 + b;
      const bool ok = active && bb < B;
      const float* row = xg + (static_cast<long long>(bb) * L + step_time(s)) * xg_pitch;
      nx0[b] = ok ? __ldg(row + xc0) : 0.f;
      nx1[b] = ok ? __ldg(row + xc1) : 0.f;
    }
  };
  fetch(0);
  int cur = 0;
  for (int s = 0; s < L; ++s) {
    float2 a0[BT], a1[BT];
#pragma unroll
    for (int b = 0; b < BT; ++b) {
      a0[b] = make_float2(nx0[b], 0.f);
      a1[b] = make_float2(nx1[b], 0.f);
    }
    if (s + 1 < L) fetch(s + 1);
    const float* hc = s_h + cur * BT * HP;
#pragma unroll
    for (int k4 = 0; k4 < KR2 / 4; ++k4) {
#pragma unroll
      for (int b = 0; b < BT; ++b) {
        const float4 hv = *reinterpret_cast<const float4*>(hc + b * HP + 4 * k4);
        const float2 hlo = make_float2(hv.x, hv.y), hhi = make_float2(hv.z, hv.w);
        a0[b] = ffma2(w0[2 * k4], hlo, a0[b]);
        a1[b] = ffma2(w1[2 * k4], hlo, a1[b]);
        a0[b] = ffma2(w0[2 * k4 + 1], hhi, a0[b]);
        a1[b] = ffma2(w1[2 * k4 + 1], hhi, a1[b]);
      }
    }
    for (int q = 0; q < nq; ++q) {
      const float4 wa = s_w[(q * 2 + 0) * LSTM2_THREADS + t];
      const float4 wb = s_w[(q * 2 + 1) * LSTM2_THREADS + t];
#pragma unroll
      for (int b = 0; b < BT; ++b) {
        const float4 hv = *reinterpret_cast<const float4*>(hc + b * HP + KR2 + 4 * q);
        const float2 hlo = make_float2(hv.x, hv.y), hhi = make_float2(hv.z, hv.w);
        a0[b] = ffma2(make_float2(wa.x, wa.y), hlo, a0[b]);
        a1[b] = ffma2(make_float2(wb.x, wb.y), hlo, a1[b]);
        a0[b] = ffma2(make_float2(wa.z, wa.w), hhi, a0[b]);
        a1[b] = ffma2(make_float2(wb.z, wb.w), hhi, a1[b]);
      }
    }
    float* hn = s_h + (cur ^ 1) * BT * HP;
    const int tt = step_time(s);
#pragma unroll
    for (int b = 0; b < BT; ++b) {
      // half 0: s(i) * tanh(g) ; half 1: s(f), s(o).  tanh(x) = 2 s(2x) - 1 on the MUFU ex2/rcp path
      const float p0 = a0[b].x + a0[b].y;
      const float p1 = a1[b].x + a1[b].y;
      const float g0 = logistic(p0);
      const float s1 = logistic(half ? p1 : 2.0f * p1);
      const float g1 = half ? s1 : fmaf(2.0f, s1, -1.0f);
      const float ig = __shfl_xor_sync(0xffffffffu, g0 * g1, 1);  // half 1 receives i*g
      if (SAVE && active && b0 + b < B) {
        float* gr = gates + (static_cast<long long>(b0 + b) * L + tt) * gates_pitch + dir * 5 * H + j;
        gr[half * H] = g0;          // i (half 0) or f (half 1)
        gr[(2 + half) * H] = g1;    // g (half 0) or o (half 1)
      }
      if (half && active) {
        const float cn = fmaf(g0, c[b], ig);
        const float hv = g1 * fmaf(2.0f, logistic(2.0f * cn), -1.0f);
        c[b] = cn;
        hn[b * HP + j] = hv;
        const int bb = b0 + b;
        if (bb < B) out[(static_cast<long long>(bb) * L + tt) * out_pitch + dir * H + j] = hv;
        if (SAVE && bb < B)
          gates[(static_cast<long long>(bb) * L + tt) * gates_pitch + dir * 5 * H + 4 * H + j] = cn;
      }
    }
    __syncthreads();
    cur ^= 1;
  }
}

// ----------------------------------------------------------------------------------------------
// Tensor-core form — the kernel that runs for inference.  One CTA owns 8 sequences of one direction
// for all steps.  The gate pre-activations of a step are  W_hh [4H x H] · h^T [H x 8]  on mma.sync
// m16n8k16 with W_hh as the A operand: gate rows are the M dimension, the 8 sequences are N, so no
// column of the tile is padding.  fp32 fidelity comes from the bf16 hi|lo split the SDNet GEMMs use:
// W·h = Wh·hh + Wh·hl + Wl·hh (the dropped Wl·hl term is 2^-16 relative).
//
// 8 warps; warp w owns units 16w .. 16w+15 as four 16-row tiles: tile (ug, 0) = rows {i of units
// 16w+8ug+0..7 ; f of the same units}, tile (ug, 1) = {g ; o}.  In the accumulator layout lane
// (g = lane/4, c = lane%4) then holds i, f, g, o of unit 16w+8ug+g for sequences 2c and 2c+1: the cell
// update needs no shuffle.  The hi halves of the thread's A fragments stay in REGISTERS for the whole
// sequence (128 registers), the lo halves of the first KLO_REG k-steps too; the other lo fragments
// sit in shared memory in fragment order (one conflict-free LDS.128 each).  h_{t-1} lives in shared
// memory as bf16 hi and lo rows [sequence][unit] (pitch 272 B: ldmatrix rows on distinct banks) and is
// the B operand through ldmatrix.x4.  The HMMAs are issued term by term over the four tiles, so two
// HMMAs on the same accumulator sit four issues apart.  xg rows (x W_ih^T + b, from the tcgen05 GEMM)
// arrive through a 4-stage ring of 1-D TMA bulk copies, three steps ahead, and are added while the tensor
// pipe drains.
// Activations: E_x = 2^(-x log2 e) on MUFU.EX2 — the factor -log2 e (-2 log2 e for the g gate: tanh) is
// folded into the W_hh fragments when they are loaded and into the xg add — and
// s(i)·tanh(g) = (1-E_g) / ((1+E_i)(1+E_g))  shares one MUFU.RCP between two gates (likewise s(o)·tanh(c)):
// 8 MUFU per cell instead of 10.  The four cells of a thread are computed together (independent MUFU
// chains), the stores that follow are branch-free.
//
// Measured (profiles/r02_lstm_microbench.txt): 1.5 us per step against 2.6 for the FMA kernel, on half
// the CTAs (64 for B = 256, both directions), so two of the three branch LSTMs run side by side.  A step
// is ~1500 clk of HMMA (192 per scheduler at 8 clk, the legacy tensor path) + ~800 clk of cell update +
// the barrier; publishing the two halves of h behind separate mbarriers so that one warp of a scheduler
// could run its cell update under the other's HMMAs was tried: the phase offset decays to lockstep within
// a few steps and nothing is gained.
constexpr int LM_THREADS = 256;
constexpr int LM_BT = 8;      // sequences per CTA = N of the MMA
constexpr int LM_HS = 136;    // bf16 pitch of an h row
constexpr int LM_XS = 4;      // xg ring stages
constexpr int LM_XST = LM_BT * 4 * HP;  // floats per ring stage

__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int KLO_REG>
constexpr size_t lstm_mma_smem() {
  return static_cast<size_t>(8 - KLO_REG) * 4 * LM_THREADS * 16 + 2 * 2 * LM_BT * LM_HS * 2 +
         static_cast<size_t>(LM_XS) * LM_XST * 4 + LM_XS * 8;
}

template <int KLO_REG>
__global__ void __launch_bounds__(LM_THREADS, 1)
lstm_recurrence_mma_kernel(const float* __restrict__ xg, long long xg_pitch,  // [B*L, ndir*4H]
                           const float* __restrict__ w_hh,                    // [ndir][4H][H]
                           float* __restrict__ out, long long out_pitch,      // [B*L, >= ndir*H]
                           int B, int L, int H) {
  constexpr int KS = 8;
  constexpr int KLO_SM = KS - KLO_REG;
  constexpr float L2E = 1.4426950408889634f;
  extern __shared__ __align__(128) unsigned char lm_smem[];
  uint4* s_wlo = reinterpret_cast<uint4*>(lm_smem);                                    // [KLO_SM][4][256]
  __nv_bfloat16* s_h = reinterpret_cast<__nv_bfloat16*>(lm_smem + KLO_SM * 4 * LM_THREADS * 16);  // [2][hi|lo][8][LM_HS]
  float* s_x = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(s_h) + 2 * 2 * LM_BT * LM_HS * 2);
  uint64_t* xbar = reinterpret_cast<uint64_t*>(s_x + LM_XS * LM_XST);
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, c = lane & 3;
  const int dir = blockIdx.y;
  const int b0 = blockIdx.x * LM_BT;
  const int nb = min(LM_BT, B - b0);
  const int rs = 4 * H;  // floats per sequence row of a ring stage
  const float* W = w_hh + static_cast<long long>(dir) * 4 * H * H;

  uint32_t whi[4][KS][4];
  uint32_t wlo[4][KLO_REG > 0 ? KLO_REG : 1][4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int unit = 16 * w + 8 * (t >> 1) + g;
    const bool uok = unit < H;
    const float* r0 = W + (static_cast<long long>(2 * (t & 1)) * H + unit) * H;   // gate i or g
    const float* r1 = r0 + static_cast<long long>(H) * H;                          // gate f or o
    const float sc0 = (t & 1) ? -2.f * L2E : -L2E, sc1 = -L2E;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      uint32_t lo4[4];
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int k = 16 * ks + 8 * hf + 2 * c;
        const float v00 = (uok && k < H) ? sc0 * __ldg(r0 + k) : 0.f;
        const float v01 = (uok && k + 1 < H) ? sc0 * __ldg(r0 + k + 1) : 0.f;
        const float v10 = (uok && k < H) ? sc1 * __ldg(r1 + k) : 0.f;
        const float v11 = (uok && k + 1 < H) ? sc1 * __ldg(r1 + k + 1) : 0.f;
        const uint32_t h0 = pack_bf16x2(v00, v01), h1 = pack_bf16x2(v10, v11);
        whi[t][ks][2 * hf] = h0;
        whi[t][ks][2 * hf + 1] = h1;
        lo4[2 * hf] = pack_bf16x2(v00 - bf16_lo(h0), v01 - bf16_hi(h0));
        lo4[2 * hf + 1] = pack_bf16x2(v10 - bf16_lo(h1), v11 - bf16_hi(h1));
      }
      if (ks < KLO_REG) {
#pragma unroll
        for (int i = 0; i < 4; ++i) wlo[t][ks < KLO_REG ? ks : 0][i] = lo4[i];
      } else {
        s_wlo[((ks - KLO_REG) * 4 + t) * LM_THREADS + tid] = make_uint4(lo4[0], lo4[1], lo4[2], lo4[3]);
      }
    }
  }
  {
    uint32_t* z = reinterpret_cast<uint32_t*>(s_h);
    for (int i = tid; i < 2 * 2 * LM_BT * LM_HS / 2; i += LM_THREADS) z[i] = 0u;
    if (nb < LM_BT)  // rows of absent sequences are never loaded: keep them finite
      for (int i = tid; i < LM_XS * LM_XST; i += LM_THREADS) s_x[i] = 0.f;
  }
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < LM_XS; ++i) mbar_init(&xbar[i], LM_THREADS / 32);
    fence_mbar_init();
  }
  fence_proxy_async();
  __syncthreads();

  auto step_time = [&](int s) { return dir == 0 ? s : (L - 1 - s); };
  // lane 0 of warp n: the xg row of sequence n at step s into ring stage s % LM_XS (every warp arrives, so
  // no warp carries all eight copies on the critical path of a step); steps are issued in order, the source
  // pointer walks one row per step
  const long long xstep = (dir == 0 ? 1 : -1) * xg_pitch;
  const float* xsrc = xg + (static_cast<long long>(b0 + (w < nb ? w : 0)) * L + step_time(0)) * xg_pitch +
                      static_cast<long long>(dir) * rs;
  auto issue = [&](int s) {
    const int stg = s % LM_XS;
    if (w < nb) {
      const uint32_t bytes = static_cast<uint32_t>(rs) * 4u;
      mbar_arrive_expect_tx(&xbar[stg], bytes);
      bulk_load_1d(s_x + stg * LM_XST + w * rs, xsrc, bytes, &xbar[stg]);
      xsrc += xstep;
    } else {
      mbar_arrive(&xbar[stg]);
    }
  };
  constexpr int AHEAD = LM_XS - 1;  // steps the copies run ahead
  if (lane == 0)
    for (int s = 0; s < AHEAD && s < L; ++s) issue(s);

  float cst[2][2];
#pragma unroll
  for (int i = 0; i < 2; ++i) cst[i][0] = cst[i][1] = 0.f;
  // ldmatrix lane address inside an h half-buffer: matrix m = lane / 8 covers k = 8m .. 8m+7 of the pair of k-steps
  const uint32_t h_base = smem_u32(s_h);
  const uint32_t lane_off = static_cast<uint32_t>(((lane & 7) * LM_HS + 8 * (lane >> 3)) * 2);
  constexpr uint32_t HBUF = 2 * LM_BT * LM_HS * 2;  // bytes per buffer (hi rows then lo rows)
  constexpr uint32_t HLO = LM_BT * LM_HS * 2;
  // units >= H read the xg value of unit H-1 (finite; their W rows are zero and nothing of theirs is stored)
  const int ux0 = min(16 * w + g, H - 1), ux1 = min(16 * w + 8 + g, H - 1);
  float* orow = out + (static_cast<long long>(b0 + 2 * c) * L + step_time(0)) * out_pitch + dir * H + 16 * w + g;
  const long long ostep = (dir == 0 ? 1 : -1) * out_pitch;
  const long long oseq = static_cast<long long>(L) * out_pitch;
  int cur = 0;
  for (int s = 0; s < L; ++s) {
    float acc[4][4];
#pragma unroll
    for (int t = 0; t < 4; ++t) acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f;
    if (s > 0) {  // h_{-1} = 0
      const uint32_t hb = h_base + cur * HBUF + lane_off;
#pragma unroll
      for (int kp = 0; kp < KS / 2; ++kp) {
        uint32_t bh[4], bl[4];
        ldsm_x4(hb + kp * 64, bh[0], bh[1], bh[2], bh[3]);
        ldsm_x4(hb + HLO + kp * 64, bl[0], bl[1], bl[2], bl[3]);
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          const int ks = 2 * kp + kk;
#pragma unroll
          for (int t = 0; t < 4; ++t)
            mma_bf16_16816(acc[t], whi[t][ks][0], whi[t][ks][1], whi[t][ks][2], whi[t][ks][3], bh[2 * kk], bh[2 * kk + 1]);
#pragma unroll
          for (int t = 0; t < 4; ++t)
            mma_bf16_16816(acc[t], whi[t][ks][0], whi[t][ks][1], whi[t][ks][2], whi[t][ks][3], bl[2 * kk], bl[2 * kk + 1]);
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            if (ks < KLO_REG) {
              const int kr = ks < KLO_REG ? ks : 0;
              mma_bf16_16816(acc[t], wlo[t][kr][0], wlo[t][kr][1], wlo[t][kr][2], wlo[t][kr][3], bh[2 * kk], bh[2 * kk + 1]);
            } else {
              const uint4 l4 = s_wlo[((ks - KLO_REG) * 4 + t) * LM_THREADS + tid];
              mma_bf16_16816(acc[t], l4.x, l4.y, l4.z, l4.w, bh[2 * kk], bh[2 * kk + 1]);
            }
          }
        }
      }
    }
    // while the tensor pipe drains: the next copy goes out (its stage was last read in the cell update of step
    // s-1, which every warp has left) and this step's xg rows are fetched from the ring
    if (lane == 0 && s + AHEAD < L) issue(s + AHEAD);
    const int stg = s % LM_XS;
    mbar_wait(&xbar[stg], (s / LM_XS) & 1);
    const float* xs = s_x + stg * LM_XST + (2 * c) * rs;
    __nv_bfloat16* hn = s_h + (cur ^ 1) * (2 * LM_BT * LM_HS);
    float hv[2][2];
#pragma unroll
    for (int ug = 0; ug < 2; ++ug) {
      const float* x0 = xs + (ug ? ux1 : ux0);
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float* xe = x0 + e * rs;
        // pre-activations times -log2 e (-2 log2 e for g); exponents capped at 2^60 so that the shared
        // denominators stay finite
        const float pi = fmaf(-L2E, xe[0], acc[2 * ug][e]);
        const float pf = fmaf(-L2E, xe[H], acc[2 * ug][2 + e]);
        const float pg = fmaf(-2.f * L2E, xe[2 * H], acc[2 * ug + 1][e]);
        const float po = fmaf(-L2E, xe[3 * H], acc[2 * ug + 1][2 + e]);
        const float Ei = ex2_approx(fminf(pi, 60.f));
        const float Eg = ex2_approx(fminf(pg, 60.f));
        const float Ef = ex2_approx(pf);
        const float Eo = ex2_approx(fminf(po, 60.f));
        const float ig = (1.f - Eg) * rcp_approx((1.f + Ei) * (1.f + Eg));
        const float fg = rcp_approx(1.f + Ef);
        const float cn = fmaf(fg, cst[ug][e], ig);
        const float Ec = ex2_approx(fminf(-2.f * L2E * cn, 60.f));
        hv[ug][e] = (1.f - Ec) * rcp_approx((1.f + Eo) * (1.f + Ec));
        cst[ug][e] = cn;
      }
    }
    // units >= H produce finite values that meet zero columns of W_hh in the next step: the shared-memory
    // stores need no predicate
#pragma unroll
    for (int ug = 0; ug < 2; ++ug) {
      const int unit = 16 * w + 8 * ug + g;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int n = 2 * c + e;
        const float v = hv[ug][e];
        const __nv_bfloat16 hh = __float2bfloat16_rn(v);
        hn[n * LM_HS + unit] = hh;
        hn[LM_BT * LM_HS + n * LM_HS + unit] = __float2bfloat16_rn(v - __bfloat162float(hh));
        if (unit < H && n < nb) orow[e * oseq + 8 * ug] = v;
      }
    }
    orow += ostep;
    __syncthreads();
    cur ^= 1;
  }
}

template <int KLO_REG>
int launch_mma(const float* xg, long long xg_pitch, const float* w_hh, float* out, long long out_pitch,
               int B, int L, int H, int ndir, cudaStream_t st) {
  constexpr size_t smem = lstm_mma_smem<KLO_REG>();
  static RuartDeviceOnce attr_set;
  if (!attr_set.done()) {
    RUART_CUDA_CHECK(cudaFuncSetAttribute(lstm_recurrence_mma_kernel<KLO_REG>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set.set();
  }
  dim3 grid((B + LM_BT - 1) / LM_BT, ndir);
  lstm_recurrence_mma_kernel<KLO_REG><<<grid, LM_THREADS, smem, st>>>(xg, xg_pitch, w_hh, out, out_pitch,
                                                                       B, L, H);
  RUART_LAUNCH_CHECK();
  return RUART_OK;
}

template <int BT, int KR2, bool SAVE = false>
int launch2(const float* xg, long long xg_pitch, const float* w_hh, float* out, long long out_pitch,
            int B, int L, int H, int ndir, cudaStream_t st, float* gates = nullptr,
            long long gates_pitch = 0) {
  constexpr int NQ2 = (HP - KR2) / 4;
  const size_t smem = (static_cast<size_t>(NQ2) * 2 * LSTM2_THREADS * 4 + 2 * BT * HP) * sizeof(float);
  static RuartDeviceOnce attr_set;
  if (!attr_set.done()) {
    RUART_CUDA_CHECK(cudaFuncSetAttribute(lstm_recurrence2_kernel<BT, KR2, SAVE>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set.set();
  }
  dim3 grid((B + BT - 1) / BT, ndir);
  lstm_recurrence2_kernel<BT, KR2, SAVE><<<grid, LSTM2_THREADS, smem, st>>>(
      xg, xg_pitch, w_hh, out, out_pitch, B, L, H, gates, gates_pitch);
  RUART_LAUNCH_CHECK();
  return RUART_OK;
}

template <int BT>
int launch(const float* xg, long long xg_pitch, const float* w_hh, float* out, long long out_pitch,
           int B, int L, int H, int ndir, cudaStream_t st) {
  const size_t smem = (static_cast<size_t>(NQ) * ROWP * 4 + 2 * BT * HP) * sizeof(float);
  static RuartDeviceOnce attr_set;
  if (!attr_set.done()) {
    RUART_CUDA_CHECK(cudaFuncSetAttribute(lstm_recurrence_kernel<BT>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set.set();
  }
  dim3 grid((B + BT - 1) / BT, ndir);
  lstm_recurrence_kernel<BT><<<grid, LSTM_THREADS, smem, st>>>(xg, xg_pitch, w_hh, out, out_pitch,
                                                               B, L, H);
  RUART_LAUNCH_CHECK();
  return RUART_OK;
}

}  // namespace

static int lstm_recurrence_dispatch(const float* xg, long long xg_pitch, const float* w_hh,
                                    float* out, long long out_pitch, int B, int L, int H,
                                    int ndir, void* stream, bool tensor_cores) {
  RUART_ARG_CHECK(B > 0 && L > 0 && H > 0 && H <= 128 && (ndir == 1 || ndir == 2));
  cudaStream_t st = (cudaStream_t)stream;
  // fewest sequences per CTA that still fits one wave of CTAs on the device
  const int sms = ruart_num_sms();
  // tensor-core form unless switched off (A/B aid) or the xg rows cannot be bulk-copied (16-byte alignment)
  static const char* fma_env = getenv("RUART_LSTM_FMA");
  static const char* klo_env = getenv("RUART_LSTM_KLO");
  if (tensor_cores && !fma_env && (xg_pitch % 4) == 0 && (reinterpret_cast<uintptr_t>(xg) & 15) == 0) {
    const int klo = klo_env ? atoi(klo_env) : 3;  // lo fragments of that many k-steps in registers (A/B aid)
    if (klo == 0) return launch_mma<0>(xg, xg_pitch, w_hh, out, out_pitch, B, L, H, ndir, st);
    if (klo == 2) return launch_mma<2>(xg, xg_pitch, w_hh, out, out_pitch, B, L, H, ndir, st);
    return launch_mma<3>(xg, xg_pitch, w_hh, out, out_pitch, B, L, H, ndir, st);
  }
  static const bool one_row = getenv("RUART_LSTM_ONE_ROW") != nullptr;  // A/B aid: the 512-thread kernel
  if (one_row) {
    if (((B + 1) / 2) * ndir <= sms) return launch<2>(xg, xg_pitch, w_hh, out, out_pitch, B, L, H, ndir, st);
    if (((B + 3) / 4) * ndir <= sms) return launch<4>(xg, xg_pitch, w_hh, out, out_pitch, B, L, H, ndir, st);
  }
  if (((B + 1) / 2) * ndir <= sms) return launch2<2, 88>(xg, xg_pitch, w_hh, out, out_pitch, B, L, H, ndir, st);
  if (((B + 3) / 4) * ndir <= sms) return launch2<4, 88>(xg, xg_pitch, w_hh, out, out_pitch, B, L, H, ndir, st);
  // larger batches: several waves of the same kernel (measured 1.45x faster than one wave of the
  // one-row kernel with 8 sequences per CTA: 517 vs 751 us at B = 512, L = 100)
  return launch2<4, 88>(xg, xg_pitch, w_hh, out, out_pitch, B, L, H, ndir, st);
}

// The production form: tensor cores, W_hh and h as bf16 hi | lo splits (~2^-16 relative, the precision of the 2-part
// SDNet stack next to a bf16 BERT).
extern "C" int ruart_lstm_recurrence(const float* xg, long long xg_pitch, const float* w_hh,
                                     float* out, long long out_pitch, int B, int L, int H,
                                     int ndir, void* stream) {
  return lstm_recurrence_dispatch(xg, xg_pitch, w_hh, out, out_pitch, B, L, H, ndir, stream, true);
}

// fp32 FMA recurrence (the fp32 mode of the stack, 3-part operands everywhere else): same arguments.
extern "C" int ruart_lstm_recurrence_f32(const float* xg, long long xg_pitch, const float* w_hh,
                                         float* out, long long out_pitch, int B, int L, int H,
                                         int ndir, void* stream) {
  return lstm_recurrence_dispatch(xg, xg_pitch, w_hh, out, out_pitch, B, L, H, ndir, stream, false);
}

// Training form: same recurrence, additionally saving the activated gates and the
// cell state of every step (see lstm_recurrence2_kernel<.., SAVE>).
extern "C" int ruart_lstm_recurrence_train(const float* xg, long long xg_pitch, const float* w_hh,
                                           float* out, long long out_pitch, int B, int L, int H,
                                           int ndir, float* gates, long long gates_pitch,
                                           void* stream) {
  RUART_ARG_CHECK(B > 0 && L > 0 && H > 0 && H <= 128 && (ndir == 1 || ndir == 2));
  RUART_ARG_CHECK(gates != nullptr && gates_pitch >= static_cast<long long>(ndir) * 5 * H);
  cudaStream_t st = (cudaStream_t)stream;
  if (((B + 1) / 2) * ndir <= ruart_num_sms())
    return launch2<2, 88, true>(xg, xg_pitch, w_hh, out, out_pitch, B, L, H, ndir, st, gates, gates_pitch);
  return launch2<4, 88, true>(xg, xg_pitch, w_hh, out, out_pitch, B, L, H, ndir, st, gates, gates_pitch);
}
