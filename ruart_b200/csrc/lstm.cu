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
#include "common.cuh"
#include "ruart_b200.h"

namespace {

using namespace ruart;

constexpr int LSTM_THREADS = 512;
constexpr int KR = 64;     // weights per row kept in registers (as 32 float2)
constexpr int HP = 128;    // padded hidden size (h rows in smem, zero padded)
constexpr int ROWP = 512;  // padded gate-row count (threads)
constexpr int NQ = (HP - KR) / 4;  // float4 weight quads per row kept in smem

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  // packed fp32x2 FMA (Blackwell FFMA2): d.x = a.x*b.x + c.x, d.y = a.y*b.y + c.y
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

template <int BT>
int launch(const float* xg, long long xg_pitch, const float* w_hh, float* out, long long out_pitch,
           int B, int L, int H, int ndir, cudaStream_t st) {
  const size_t smem = (static_cast<size_t>(NQ) * ROWP * 4 + 2 * BT * HP) * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    RUART_CUDA_CHECK(cudaFuncSetAttribute(lstm_recurrence_kernel<BT>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  dim3 grid((B + BT - 1) / BT, ndir);
  lstm_recurrence_kernel<BT><<<grid, LSTM_THREADS, smem, st>>>(xg, xg_pitch, w_hh, out, out_pitch,
                                                               B, L, H);
  RUART_LAUNCH_CHECK();
  return RUART_OK;
}

}  // namespace

extern "C" int ruart_lstm_recurrence(const float* xg, long long xg_pitch, const float* w_hh,
                                     float* out, long long out_pitch, int B, int L, int H,
                                     int ndir, void* stream) {
  RUART_ARG_CHECK(B > 0 && L > 0 && H > 0 && H <= 128 && (ndir == 1 || ndir == 2));
  cudaStream_t st = (cudaStream_t)stream;
  // fewest sequences per CTA that still fits one wave of CTAs on the device
  const int sms = ruart_num_sms();
  if (((B + 1) / 2) * ndir <= sms) return launch<2>(xg, xg_pitch, w_hh, out, out_pitch, B, L, H, ndir, st);
  if (((B + 3) / 4) * ndir <= sms) return launch<4>(xg, xg_pitch, w_hh, out, out_pitch, B, L, H, ndir, st);
  return launch<8>(xg, xg_pitch, w_hh, out, out_pitch, B, L, H, ndir, st);
}
