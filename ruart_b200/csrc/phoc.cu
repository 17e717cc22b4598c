// PHOC (pyramidal histogram of characters) featuriser, batch form, for sm_100a.
//
// Replaces Utils/cphoc.c:12-113 (`build_phoc`): 604 = 36 unigrams x (2+3+4+5) pyramid regions
// + 50 bigrams x 2 regions.  A feature is set iff the character's (bigram's) normalised extent
// overlaps the region by >= 0.5 *in IEEE float32 arithmetic, in the reference's operation order*
// (cphoc.c:34-35,56-61,89-98), so every division / subtraction below is an explicit
// round-to-nearest fp32 op (__fdiv_rn / __fsub_rn, no fast-math, nothing to contract into FMA).
//
// Layout / roofline: the kernel is pure HBM write traffic (2416 B out per ~10 B in).  One warp
// owns one string: lanes walk the characters and OR bits into a 19-word shared-memory bitmask,
// then the warp streams the row out as 151 coalesced float4 stores (st.global.cs).  Grid is a
// multiple of the SM count; warps grid-stride over strings.
#include "common.cuh"
#include "ruart_b200.h"

namespace {

constexpr int PHOC_WORDS = 19;  // ceil(604 / 32)
constexpr int WARPS_PER_CTA = 8;

// bigram table of cphoc.c:30, indexed [first*36 + second] -> 0..49 or -1
__constant__ int8_t c_bigram_lut[36 * 36];

const char* const kBigrams[50] = {
    "th", "he", "in", "er", "an", "re", "es", "on", "st", "nt", "en", "at", "ed",
    "nd", "to", "or", "ea", "ti", "ar", "te", "ng", "al", "it", "as", "is", "ha",
    "et", "se", "ou", "of", "le", "sa", "ve", "ro", "ra", "ri", "hi", "ne", "me",
    "de", "co", "ta", "ec", "si", "ll", "so", "na", "li", "la", "el"};

__device__ __forceinline__ int unigram_index(uint8_t c) {
  // order of cphoc.c:29: a..z then 0..9
  if (c >= 'a' && c <= 'z') return c - 'a';
  if (c >= '0' && c <= '9') return 26 + (c - '0');
  return -1;
}

// true iff (min(occ1, r1) - max(occ0, r0)) / (occ1 - occ0) >= 0.5f, all in fp32 RN.
__device__ __forceinline__ bool overlaps_half(float occ0, float occ1, int region, int level) {
  const float r0 = __fdiv_rn(static_cast<float>(region), static_cast<float>(level));
  const float r1 = __fdiv_rn(static_cast<float>(region + 1), static_cast<float>(level));
  const float o0 = (occ0 > r0) ? occ0 : r0;
  const float o1 = (occ1 < r1) ? occ1 : r1;
  const float ratio = __fdiv_rn(__fsub_rn(o1, o0), __fsub_rn(occ1, occ0));
  return ratio >= 0.5f;
}

// The region tests depend only on (string length n, character position idx), so for n <= LUT_N
// they are tabulated ONCE per process by lut_init_kernel — which evaluates exactly the same fp32
// expressions (overlaps_half) as the direct path, hence bit-identical — and the hot kernel replaces
// 28 IEEE divisions per character by one shared-memory lookup.
//   uni[n][idx] : bit (level_base_index + region) for the 14 (level, region) pairs, levels 2..5
//   bi[n][idx]  : bit region (0/1) for a bigram starting at idx
constexpr int LUT_N = 32;
__device__ uint16_t g_lut_uni[(LUT_N + 1) * LUT_N];
__device__ uint8_t g_lut_bi[(LUT_N + 1) * LUT_N];
__device__ int8_t g_bigram_lut[36 * 36];  // copy of c_bigram_lut for staging into shared memory

__global__ void lut_init_kernel() {
  const int n = blockIdx.x;  // 0 .. LUT_N
  const int idx = threadIdx.x;
  if (idx >= LUT_N) return;
  uint32_t m = 0, b = 0;
  if (n > 0 && idx < n) {
    const float fn = static_cast<float>(n);
    const float occ0 = __fdiv_rn(static_cast<float>(idx), fn);
    const float occ1 = __fdiv_rn(static_cast<float>(idx + 1), fn);
    int bit = 0;
    for (int level = 2; level < 6; ++level)
      for (int region = 0; region < level; ++region, ++bit)
        if (overlaps_half(occ0, occ1, region, level)) m |= 1u << bit;
    const float g1 = __fdiv_rn(static_cast<float>(idx + 2), fn);
    for (int region = 0; region < 2; ++region)
      if (overlaps_half(occ0, g1, region, 2)) b |= 1u << region;
  }
  g_lut_uni[n * LUT_N + idx] = static_cast<uint16_t>(m);
  g_lut_bi[n * LUT_N + idx] = static_cast<uint8_t>(b);
}

// feature offset of (level, region) pair number `bit` (levels 2..5 in order): region blocks of 36
__device__ __forceinline__ int pair_base(int bit) { return bit * 36; }

__device__ __forceinline__ void build_mask(const uint8_t* __restrict__ chars, int begin, int n,
                                           uint32_t* mask, int lane, int64_t str_idx, int32_t* err,
                                           const uint16_t* s_uni, const uint8_t* s_bi,
                                           const int8_t* s_big, uint32_t c_pref) {
  if (lane < PHOC_WORDS) mask[lane] = 0u;
  __syncwarp();
  bool bad = false;
  if (n <= LUT_N) {
    // tabulated path: one lane per character
    // c_pref: this lane's character, loaded one iteration ahead; the next one comes by shuffle
    const uint32_t c_next = __shfl_down_sync(0xffffffffu, c_pref, 1);
    if (lane < n) {
      const uint8_t c = static_cast<uint8_t>(c_pref);
      const int ci = unigram_index(c);
      if (ci < 0) {
        bad = true;
        const unsigned long long key = (static_cast<unsigned long long>(str_idx) << 24) |
                                       (static_cast<unsigned long long>(lane) << 8) | c;
        atomicMin(reinterpret_cast<unsigned long long*>(err), key);
      } else {
        uint32_t m = s_uni[n * LUT_N + lane];
        while (m) {
          const int bit = __ffs(m) - 1;
          m &= m - 1;
          const int f = pair_base(bit) + ci;
          atomicOr(&mask[f >> 5], 1u << (f & 31));
        }
        if (lane + 1 < n) {
          const int cj = unigram_index(static_cast<uint8_t>(c_next));
          if (cj >= 0) {
            const int bidx = s_big[ci * 36 + cj];
            if (bidx >= 0) {
              const uint32_t b = s_bi[n * LUT_N + lane];
              if (b & 1u) atomicOr(&mask[(504 + bidx) >> 5], 1u << ((504 + bidx) & 31));
              if (b & 2u) atomicOr(&mask[(554 + bidx) >> 5], 1u << ((554 + bidx) & 31));
            }
          }
        }
      }
    }
    const bool any_bad_fast = __any_sync(0xffffffffu, bad);
    __syncwarp();
    if (any_bad_fast && lane < PHOC_WORDS) mask[lane] = 0u;
    __syncwarp();
    return;
  }
  const float fn = static_cast<float>(n);
  for (int idx = lane; idx < n; idx += 32) {
    const uint8_t c = chars[begin + idx];
    const int ci = unigram_index(c);
    if (ci < 0) {
      bad = true;
      // first offending string wins, like the reference's early RuntimeError (cphoc.c:45-50)
      const unsigned long long key =
          (static_cast<unsigned long long>(str_idx) << 24) |
          (static_cast<unsigned long long>(idx < 65535 ? idx : 65535) << 8) | c;
      atomicMin(reinterpret_cast<unsigned long long*>(err), key);
      continue;
    }
    const float occ0 = __fdiv_rn(static_cast<float>(idx), fn);
    const float occ1 = __fdiv_rn(static_cast<float>(idx + 1), fn);
    int level_base = 0;  // (sum of l < level) * 36  (cphoc.c:64-66)
#pragma unroll
    for (int level = 2; level < 6; ++level) {
#pragma unroll
      for (int region = 0; region < level; ++region) {
        if (overlaps_half(occ0, occ1, region, level)) {
          const int f = level_base + region * 36 + ci;
          atomicOr(&mask[f >> 5], 1u << (f & 31));
        }
      }
      level_base += level * 36;
    }
    if (idx + 1 < n) {
      const int cj = unigram_index(chars[begin + idx + 1]);
      if (cj >= 0) {
        const int bi = c_bigram_lut[ci * 36 + cj];
        if (bi >= 0) {
          const float g0 = __fdiv_rn(static_cast<float>(idx), fn);
          const float g1 = __fdiv_rn(static_cast<float>(idx + 2), fn);
#pragma unroll
          for (int region = 0; region < 2; ++region) {
            if (overlaps_half(g0, g1, region, 2)) {
              const int f = 504 + region * 50 + bi;
              atomicOr(&mask[f >> 5], 1u << (f & 31));
            }
          }
        }
      }
    }
  }
  // a string with an unknown unigram yields no features at all in the reference (it raises)
  const bool any_bad = __any_sync(0xffffffffu, bad);
  __syncwarp();
  if (any_bad && lane < PHOC_WORDS) mask[lane] = 0u;
  __syncwarp();
}

template <bool kPacked>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
phoc_kernel(const uint8_t* __restrict__ chars, const int32_t* __restrict__ offsets, int64_t n,
            float* __restrict__ out, int64_t out_pitch, uint32_t* __restrict__ out_words, int32_t* err) {
  __shared__ uint32_t s_mask[WARPS_PER_CTA][PHOC_WORDS + 1];
  __shared__ uint16_t s_uni[(LUT_N + 1) * LUT_N];
  __shared__ uint8_t s_bi[(LUT_N + 1) * LUT_N];
  __shared__ int8_t s_big[36 * 36];
  for (int i = threadIdx.x; i < (LUT_N + 1) * LUT_N; i += blockDim.x) {
    s_uni[i] = g_lut_uni[i];
    s_bi[i] = g_lut_bi[i];
  }
  for (int i = threadIdx.x; i < 36 * 36; i += blockDim.x) s_big[i] = g_bigram_lut[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint32_t* mask = s_mask[warp];
  const int64_t warp_global = static_cast<int64_t>(blockIdx.x) * WARPS_PER_CTA + warp;
  const int64_t n_warps = static_cast<int64_t>(gridDim.x) * WARPS_PER_CTA;
  // two-deep software pipeline over this warp's strings: offsets are loaded two iterations ahead,
  // characters one ahead, so neither dependent global load stalls the in-order instruction stream
  auto load_off = [&](int64_t s, int& b, int& l) {
    b = 0;
    l = 0;
    if (s < n) {
      b = offsets[s];
      l = offsets[s + 1] - b;
    }
  };
  auto load_chr = [&](int b, int l) -> uint32_t {
    return (l <= LUT_N && lane < l) ? chars[b + lane] : 0u;
  };
  int b0, l0, b1, l1, b2, l2;
  load_off(warp_global, b0, l0);
  load_off(warp_global + n_warps, b1, l1);
  uint32_t c0 = load_chr(b0, l0);
  for (int64_t s = warp_global; s < n; s += n_warps) {
    const int begin = b0, len = l0;
    const uint32_t c_cur = c0;
    c0 = load_chr(b1, l1);                  // characters of the next string
    load_off(s + 2 * n_warps, b2, l2);      // offsets of the one after
    b0 = b1; l0 = l1; b1 = b2; l1 = l2;
    build_mask(chars, begin, len, mask, lane, s, err, s_uni, s_bi, s_big, c_cur);
    if (kPacked) {
      if (lane < PHOC_WORDS) out_words[s * PHOC_WORDS + lane] = mask[lane];
    } else {
      float4* row = reinterpret_cast<float4*>(out + s * out_pitch);
#pragma unroll
      for (int q = lane; q < RUART_PHOC_DIM / 4; q += 32) {
        const int f = q * 4;
        const uint32_t bits = (mask[f >> 5] >> (f & 31)) & 0xFu;  // 4 | 32: never straddles
        float4 v;
        v.x = (bits & 1u) ? 1.0f : 0.0f;
        v.y = (bits & 2u) ? 1.0f : 0.0f;
        v.z = (bits & 4u) ? 1.0f : 0.0f;
        v.w = (bits & 8u) ? 1.0f : 0.0f;
        __stcs(row + q, v);
      }
    }
    __syncwarp();
  }
}

int upload_lut() {
  static RuartDeviceOnce done;  // the LUT symbols live in each device's memory
  if (done.done()) return RUART_OK;
  int8_t lut[36 * 36];
  for (int i = 0; i < 36 * 36; ++i) lut[i] = -1;
  for (int k = 49; k >= 0; --k) {  // first match wins (cphoc.c:78-84): fill back to front
    const int a = kBigrams[k][0] - 'a';
    const int b = kBigrams[k][1] - 'a';
    lut[a * 36 + b] = static_cast<int8_t>(k);
  }
  RUART_CUDA_CHECK(cudaMemcpyToSymbol(c_bigram_lut, lut, sizeof(lut)));
  RUART_CUDA_CHECK(cudaMemcpyToSymbol(g_bigram_lut, lut, sizeof(lut)));
  lut_init_kernel<<<LUT_N + 1, 32>>>();
  RUART_CUDA_CHECK(cudaGetLastError());
  RUART_CUDA_CHECK(cudaDeviceSynchronize());
  done.set();
  return RUART_OK;
}

int launch(const uint8_t* chars, const int32_t* offsets, int64_t n, float* out, int64_t out_pitch,
           uint32_t* out_words, int32_t* err, cudaStream_t st) {
  RUART_ARG_CHECK(n >= 0 && offsets != nullptr && err != nullptr);
  RUART_ARG_CHECK((reinterpret_cast<uintptr_t>(err) & 7u) == 0);
  int rc = upload_lut();
  if (rc != RUART_OK) return rc;
  RUART_CUDA_CHECK(cudaMemsetAsync(err, 0xFF, 8, st));
  if (n == 0) return RUART_OK;
  // 8 CTAs of 8 warps per SM keeps 64 warps resident; cap by the work available.
  int64_t ctas = (n + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
  const int64_t max_ctas = static_cast<int64_t>(ruart_num_sms()) * 8;
  if (ctas > max_ctas) ctas = max_ctas;
  if (out != nullptr) {
    RUART_ARG_CHECK((reinterpret_cast<uintptr_t>(out) & 15u) == 0);
    RUART_ARG_CHECK(out_pitch >= RUART_PHOC_DIM && (out_pitch & 3) == 0);
    phoc_kernel<false><<<static_cast<unsigned>(ctas), WARPS_PER_CTA * 32, 0, st>>>(
        chars, offsets, n, out, out_pitch, nullptr, err);
  } else {
    RUART_ARG_CHECK(out_words != nullptr);
    phoc_kernel<true><<<static_cast<unsigned>(ctas), WARPS_PER_CTA * 32, 0, st>>>(
        chars, offsets, n, nullptr, 0, out_words, err);
  }
  RUART_LAUNCH_CHECK();
  return RUART_OK;
}

}  // namespace

extern "C" int ruart_phoc_batch(const uint8_t* chars, const int32_t* offsets, int64_t n, float* out,
                                int32_t* err, void* stream) {
  RUART_ARG_CHECK(out != nullptr || n == 0);
  return launch(chars, offsets, n, out, RUART_PHOC_DIM, nullptr, err, (cudaStream_t)stream);
}

extern "C" int ruart_phoc_batch_pitched(const uint8_t* chars, const int32_t* offsets, int64_t n,
                                        float* out, int64_t out_pitch, int32_t* err, void* stream) {
  RUART_ARG_CHECK(out != nullptr || n == 0);
  return launch(chars, offsets, n, out, out_pitch, nullptr, err, (cudaStream_t)stream);
}

extern "C" int ruart_phoc_batch_packed(const uint8_t* chars, const int32_t* offsets, int64_t n,
                                       uint32_t* out_words, int32_t* err, void* stream) {
  RUART_ARG_CHECK(out_words != nullptr || n == 0);
  return launch(chars, offsets, n, nullptr, 0, out_words, err, (cudaStream_t)stream);
}

extern "C" int ruart_phoc_batch_host(const char* chars_host, const int32_t* offsets_host, int64_t n,
                                     float* out_host, int64_t* bad_index, int32_t* bad_char) {
  RUART_ARG_CHECK(n >= 0 && offsets_host != nullptr && (out_host != nullptr || n == 0));
  if (bad_index) *bad_index = -1;
  if (bad_char) *bad_char = 0;
  if (n == 0) return RUART_OK;
  const int64_t total_chars = offsets_host[n];
  uint8_t* d_chars = nullptr;
  int32_t* d_off = nullptr;
  float* d_out = nullptr;
  int32_t* d_err = nullptr;
  int rc = RUART_OK;
  cudaError_t e;
#define PH_CHECK(x)                                                              \
  if ((e = (x)) != cudaSuccess) {                                                \
    ruart_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #x, cudaGetErrorString(e)); \
    rc = RUART_ERR_CUDA;                                                         \
    goto done;                                                                   \
  }
  PH_CHECK(cudaMalloc(&d_chars, total_chars > 0 ? total_chars : 1));
  PH_CHECK(cudaMalloc(&d_off, (n + 1) * sizeof(int32_t)));
  PH_CHECK(cudaMalloc(&d_out, n * RUART_PHOC_DIM * sizeof(float)));
  PH_CHECK(cudaMalloc(&d_err, 8));
  if (total_chars > 0) PH_CHECK(cudaMemcpy(d_chars, chars_host, total_chars, cudaMemcpyHostToDevice));
  PH_CHECK(cudaMemcpy(d_off, offsets_host, (n + 1) * sizeof(int32_t), cudaMemcpyHostToDevice));
  rc = launch(d_chars, d_off, n, d_out, RUART_PHOC_DIM, nullptr, d_err, 0);
  if (rc != RUART_OK) goto done;
  PH_CHECK(cudaMemcpy(out_host, d_out, n * RUART_PHOC_DIM * sizeof(float), cudaMemcpyDeviceToHost));
  {
    unsigned long long key = 0;
    PH_CHECK(cudaMemcpy(&key, d_err, 8, cudaMemcpyDeviceToHost));
    if (key != ~0ull) {
      if (bad_index) *bad_index = static_cast<int64_t>(key >> 24);
      if (bad_char) *bad_char = static_cast<int32_t>(key & 0xFF);
      ruart_set_error("Error: unigram %c is unknown", static_cast<char>(key & 0xFF));
      rc = RUART_ERR_PHOC_CHAR;
    }
  }
done:
#undef PH_CHECK
  cudaFree(d_chars);
  cudaFree(d_off);
  cudaFree(d_out);
  cudaFree(d_err);
  return rc;
}
