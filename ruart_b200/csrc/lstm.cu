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
constexpr int KR = 64;        // weights per row kept in registers
constexpr int HP = 128;       // padded hidden size (h rows in smem)
constexpr int ROWP = 512;     // padded gate-row count (smem weight pitch)

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

template <int BT>
__global__ void __launch_bounds__(LSTM_THREADS, 1)
lstm_recurrence_kernel(const float* __restrict__ xg, long long xg_pitch,  // [B*L, ndir*4H]
                       const float* __restrict__ w_hh,                    // [ndir][4H][H]
                       float* __restrict__ out, long long out_pitch,      // [B*L, >= ndir*H]
                       int B, int L, int H) {
  extern __shared__ float smem[];
  float* s_w = smem;                               // [(H-KR)][ROWP]
  float* s_h = smem + (HP - KR) * ROWP;            // [2][BT][HP]
  const int t = threadIdx.x;
  const int dir = blockIdx.y;
  const int b0 = blockIdx.x * BT;
  const int rows = 4 * H;
  const bool active = t < rows;
  const int j = t >> 2, g = t & 3;
  const int wrow = g * H + j;  // row of W_hh / column of xg for this thread
  const float* W = w_hh + static_cast<long long>(dir) * rows * H;

  float wreg[KR];
#pragma unroll
  for (int k = 0; k < KR; ++k) wreg[k] = (active && k < H) ? W[static_cast<long long>(wrow) * H + k] : 0.f;
  for (int k = KR; k < H; ++k)
    if (active) s_w[(k - KR) * ROWP + t] = W[static_cast<long long>(wrow) * H + k];
  for (int i = t; i < 2 * BT * HP; i += LSTM_THREADS) s_h[i] = 0.f;
  __syncthreads();

  float c[BT];
  float nxt[BT];
#pragma unroll
  for (int b = 0; b < BT; ++b) c[b] = 0.f;
  const long long xcol = static_cast<long long>(dir) * rows + wrow;
  auto step_time = [&](int s) { return dir == 0 ? s : (L - 1 - s); };
  // prefetch step 0
#pragma unroll
  for (int b = 0; b < BT; ++b) {
    const int bb = b0 + b;
    nxt[b] = (active && bb < B)
                 ? __ldg(xg + (static_cast<long long>(bb) * L + step_time(0)) * xg_pitch + xcol)
                 : 0.f;
  }
  int cur = 0;
  for (int s = 0; s < L; ++s) {
    float acc[BT];
#pragma unroll
    for (int b = 0; b < BT; ++b) acc[b] = nxt[b];
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
    for (int k = 0; k < KR; k += 4) {
#pragma unroll
      for (int b = 0; b < BT; ++b) {
        const float4 hv = *reinterpret_cast<const float4*>(hc + b * HP + k);
        acc[b] = fmaf(wreg[k], hv.x, acc[b]);
        acc[b] = fmaf(wreg[k + 1], hv.y, acc[b]);
        acc[b] = fmaf(wreg[k + 2], hv.z, acc[b]);
        acc[b] = fmaf(wreg[k + 3], hv.w, acc[b]);
      }
    }
    for (int k = KR; k < H; ++k) {
      const float wv = s_w[(k - KR) * ROWP + t];
#pragma unroll
      for (int b = 0; b < BT; ++b) acc[b] = fmaf(wv, hc[b * HP + k], acc[b]);
    }
    float* hn = s_h + (cur ^ 1) * BT * HP;
    const int tt = step_time(s);
#pragma unroll
    for (int b = 0; b < BT; ++b) {
      const float vi = acc[b];
      const float vf = __shfl_down_sync(0xffffffffu, acc[b], 1);
      const float vg = __shfl_down_sync(0xffffffffu, acc[b], 2);
      const float vo = __shfl_down_sync(0xffffffffu, acc[b], 3);
      if (g == 0 && active) {
        const float cn = sigmoidf_(vf) * c[b] + sigmoidf_(vi) * tanhf(vg);
        const float hv = sigmoidf_(vo) * tanhf(cn);
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
  const size_t smem = (static_cast<size_t>(HP - KR) * ROWP + 2 * BT * HP) * sizeof(float);
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
