/* ruart_b200 — C ABI of the B200 (sm_100a) kernels behind RUArt's per-question inference path.
 *
 * The reference (xiaojino/RUArt) is pure Python/PyTorch plus one CPython extension
 * (Utils/cphoc.c).  It has no FFI for the model path, so each entry point below cites the
 * reference *Python/C interface it replaces* (file:line relative to the reference tree).
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary
 *   - every pointer is a DEVICE pointer unless the parameter name ends in `_host`
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream)
 *   - return value: 0 (RUART_OK) on success, otherwise an error code; the message is kept in a
 *     thread-local buffer readable with ruart_last_error()
 *   - no allocation inside unless stated; callers pass outputs / workspaces
 *   - all functions are asynchronous with respect to the host unless stated
 */
#ifndef RUART_B200_H_
#define RUART_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define RUART_API __attribute__((visibility("default")))
#else
#define RUART_API
#endif

#define RUART_PHOC_DIM 604

/* GEMM epilogues */
#define RUART_EPI_NONE 0       /* C = A W^T                                             */
#define RUART_EPI_BIAS 1       /* C = A W^T + b            nn.Linear                    */
#define RUART_EPI_BIAS_GELU 2  /* C = gelu_erf(A W^T + b)  modeling.py:52-57,286-289    */
#define RUART_EPI_RELU_SCALE 3 /* C = relu(A W^T) * d      Layers.py:226-231            */
#define RUART_EPI_BIAS_RELU 4  /* C = relu(A W^T + b)                                   */

/* ---------------------------------------------------------------- library management */
RUART_API const char* ruart_last_error(void);
RUART_API int ruart_version(void);
/* number of SMs of the current device (cached) */
RUART_API int ruart_num_sms(void);

/* ---------------------------------------------------------------- PHOC featuriser
 * Replaces Utils/cphoc.c:12-113 `build_phoc(str) -> float[604]` (bound in Python by
 * Utils/phoc.py:8-13).  Batch form: string i is chars[offsets[i] .. offsets[i+1]).
 * `err` is an 8-byte, 8-byte-aligned device buffer initialised by the callee to ~0; unknown
 * characters (outside [a-z0-9]) are reported like the reference's RuntimeError
 * (cphoc.c:45-50): that string's row is left all-zero and err receives the smallest key
 * (string_index << 24 | char_position << 8 | char) over all offending characters.           */
RUART_API int ruart_phoc_batch(const uint8_t* chars, const int32_t* offsets, int64_t n, float* out,
                     int32_t* err, void* stream);
/* Same, bit-packed: 19 uint32 words per string (bit b of word w = feature 32*w+b). */
RUART_API int ruart_phoc_batch_packed(const uint8_t* chars, const int32_t* offsets, int64_t n,
                            uint32_t* out_words, int32_t* err, void* stream);
/* Host-buffer convenience (synchronous): copies in, runs, copies out.  Returns
 * RUART_ERR_PHOC_CHAR and fills bad_index/bad_char when a string has an unknown unigram.  */
RUART_API int ruart_phoc_batch_host(const char* chars_host, const int32_t* offsets_host, int64_t n,
                          float* out_host, int64_t* bad_index, int32_t* bad_char);

/* ---------------------------------------------------------------- tcgen05 GEMM
 * C[M,N] = epi(A[M,K] W[N,K]^T).  A and W are bf16, K-major, possibly "split" into parts laid
 * side by side in each row ([rows, parts*Kp], Kp a multiple of 64, zero padded).
 * n_terms: 1 (plain bf16), 3 (two-part split, ~2^-16 rel) or 6 (three-part split, ~fp32).
 * Outputs: fp32 (out_f32, ldo_f32) and/or bf16 (out_bf16, ldo_bf16); the bf16 output may itself
 * be written as out_parts split parts, part p at column offset p*out_part_stride.
 * Replaces nn.Linear at modeling.py:225-227,261,287,300 and Layers.py:226-227,166 (W_ih).    */
RUART_API int ruart_gemm_bf16(const void* A, long long lda, int a_parts, const void* W, long long ldw,
                    int w_parts, int M, int N, int Kp, int n_terms, int epi, const float* bias,
                    const float* scale, int scale_len, float* out_f32, long long ldo_f32,
                    void* out_bf16, long long ldo_bf16, int out_parts, long long out_part_stride,
                    int fast_gelu, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RUART_B200_H_ */
