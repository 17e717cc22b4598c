// Optimizer side of the training step (SDNetTrainer.update,
// Models/SDNetTrainer.py:363-365): torch.nn.utils.clip_grad_norm_ + torch.optim.Adamax over ONE flat
// fp32 buffer holding all trainable parameters (12.25 M elements = 49 MB in the shipped conf; the same
// flat gradient buffer is what the data-parallel step all-reduces with NCCL).  Two HBM-bound passes:
//   grad_sqnorm   sum of squares of the flat gradient, double accumulation, deterministic two-stage
//                 reduction (per-CTA partials in fixed slots, then one CTA in fixed order)
//   adamax_step   reads the squared norm from device memory (no host sync), applies the clip
//                 coefficient min(1, max_norm / (norm + 1e-6)) on the fly and updates p, exp_avg,
//                 exp_inf in place: 4 reads + 3 writes of 4 bytes per element.
// The backward kernels that would produce the gradients are not built yet.
#include "common.cuh"
#include "ruart_b200.h"

namespace {

using namespace ruart;

constexpr int TR_THREADS = 256;
constexpr int TR_MAX_PARTIALS = 1024;

__global__ void __launch_bounds__(TR_THREADS)
grad_sqnorm_partial_kernel(const float* __restrict__ g, long long n, double* __restrict__ partials) {
  double acc = 0.0;
  const long long stride = static_cast<long long>(gridDim.x) * TR_THREADS * 4;
  for (long long i = (static_cast<long long>(blockIdx.x) * TR_THREADS + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 3 < n && ((reinterpret_cast<uintptr_t>(g + i) & 15u) == 0)) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(g + i));
      acc += static_cast<double>(v.x) * v.x + static_cast<double>(v.y) * v.y +
             static_cast<double>(v.z) * v.z + static_cast<double>(v.w) * v.w;
    } else {
      for (long long j = i; j < n && j < i + 4; ++j) acc += static_cast<double>(g[j]) * g[j];
    }
  }
  acc = warp_sum_d(acc);
  __shared__ double s_w[TR_THREADS / 32];
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < TR_THREADS / 32; ++w) t += s_w[w];
    partials[blockIdx.x] = t;
  }
}

__global__ void grad_sqnorm_final_kernel(const double* __restrict__ partials, int n_partials,
                                         double* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < n_partials; ++i) t += partials[i];
    *out = t;
  }
}

__global__ void __launch_bounds__(TR_THREADS)
adamax_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ exp_avg,
                   float* __restrict__ exp_inf, long long n, float clr, float beta1, float beta2,
                   float eps, const double* __restrict__ grad_sq, float max_norm) {
  float coef = 1.0f;
  if (grad_sq != nullptr && max_norm > 0.f) {
    const float norm = static_cast<float>(sqrt(*grad_sq));
    coef = fminf(max_norm / (norm + 1e-6f), 1.0f);
  }
  const long long stride = static_cast<long long>(gridDim.x) * TR_THREADS;
  for (long long i = static_cast<long long>(blockIdx.x) * TR_THREADS + threadIdx.x; i < n; i += stride) {
    const float gi = g[i] * coef;
    const float m = beta1 * exp_avg[i] + (1.0f - beta1) * gi;
    const float u = fmaxf(exp_inf[i] * beta2, fabsf(gi) + eps);
    exp_avg[i] = m;
    exp_inf[i] = u;
    p[i] = p[i] - clr * (m / u);
  }
}

}  // namespace

extern "C" int ruart_grad_sqnorm(const float* g, long long n, double* workspace, double* out_sq,
                                 void* stream) {
  RUART_ARG_CHECK(n >= 0 && workspace != nullptr && out_sq != nullptr && (g != nullptr || n == 0));
  cudaStream_t st = (cudaStream_t)stream;
  long long ctas = (n + TR_THREADS * 4 - 1) / (TR_THREADS * 4);
  const long long cap = static_cast<long long>(ruart_num_sms()) * 4;
  if (ctas > cap) ctas = cap;
  if (ctas > TR_MAX_PARTIALS) ctas = TR_MAX_PARTIALS;
  if (ctas < 1) ctas = 1;
  grad_sqnorm_partial_kernel<<<static_cast<unsigned>(ctas), TR_THREADS, 0, st>>>(g, n, workspace);
  RUART_LAUNCH_CHECK();
  grad_sqnorm_final_kernel<<<1, 32, 0, st>>>(workspace, static_cast<int>(ctas), out_sq);
  RUART_LAUNCH_CHECK();
  return RUART_OK;
}

extern "C" int ruart_adamax_step(float* p, const float* g, float* exp_avg, float* exp_inf, long long n,
                                 float lr, float beta1, float beta2, float eps, int step,
                                 const double* grad_sq, float max_norm, void* stream) {
  RUART_ARG_CHECK(n >= 0 && step >= 1 && beta1 >= 0.f && beta1 < 1.f);
  if (n == 0) return RUART_OK;
  RUART_ARG_CHECK(p != nullptr && g != nullptr && exp_avg != nullptr && exp_inf != nullptr);
  const float clr = lr / (1.0f - powf(beta1, static_cast<float>(step)));
  long long ctas = (n + TR_THREADS - 1) / TR_THREADS;
  const long long cap = static_cast<long long>(ruart_num_sms()) * 8;
  if (ctas > cap) ctas = cap;
  adamax_step_kernel<<<static_cast<unsigned>(ctas), TR_THREADS, 0, (cudaStream_t)stream>>>(
      p, g, exp_avg, exp_inf, n, clr, beta1, beta2, eps, grad_sq, max_norm);
  RUART_LAUNCH_CHECK();
  return RUART_OK;
}
