/* TEST INFRASTRUCTURE — CPU oracle for the PHOC featuriser.  Not part of the product path:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.
 *
 * Plain-C restatement of the algorithm of the reference's CPython extension
 * Utils/cphoc.c:12-113 (xiaojino/RUArt), written from its behaviour:
 *   - 604 floats = 36 unigrams x (2+3+4+5) regions  +  50 bigrams x 2 regions
 *   - character k of an n-character word occupies [k/n, (k+1)/n); region r of level L occupies
 *     [r/L, (r+1)/L); the feature is set when overlap / extent >= 0.5, evaluated in float32 in
 *     exactly this order: divisions first, then max/min, then subtract, subtract, divide
 *     (cphoc.c:34-35 occupancy; :56-61 unigram test; :89-98 bigram test)
 *   - feature index: (sum of levels below L)*36 + r*36 + unigram  (cphoc.c:64-67)
 *                    504 + r*50 + bigram                         (cphoc.c:73,100)
 *   - an unknown unigram aborts the word (the reference raises RuntimeError, cphoc.c:45-50)
 *
 * Pinned against: the reference's own build of cphoc.c (oracle/_ref/cphoc*.so, see
 * oracle/Makefile) on 200 000 random strings, and the known answers recorded in
 * tests/golden/phoc_known.json.
 */
#include <stdint.h>
#include <string.h>

#define PHOC_DIM 604

static const char k_unigrams[36] = {'a', 'b', 'c', 'd', 'e', 'f', 'g', 'h', 'i', 'j', 'k', 'l',
                                    'm', 'n', 'o', 'p', 'q', 'r', 's', 't', 'u', 'v', 'w', 'x',
                                    'y', 'z', '0', '1', '2', '3', '4', '5', '6', '7', '8', '9'};

static const char k_bigrams[50][3] = {
    "th", "he", "in", "er", "an", "re", "es", "on", "st", "nt", "en", "at", "ed",
    "nd", "to", "or", "ea", "ti", "ar", "te", "ng", "al", "it", "as", "is", "ha",
    "et", "se", "ou", "of", "le", "sa", "ve", "ro", "ra", "ri", "hi", "ne", "me",
    "de", "co", "ta", "ec", "si", "ll", "so", "na", "li", "la", "el"};

static int find_unigram(char c) {
  for (int k = 0; k < 36; ++k)
    if (k_unigrams[k] == c) return k;
  return -1;
}

static int find_bigram(const char* s) {
  for (int k = 0; k < 50; ++k)
    if (k_bigrams[k][0] == s[0] && k_bigrams[k][1] == s[1]) return k;
  return -1;
}

/* volatile stores keep every intermediate a genuine IEEE float32 (no x87 excess precision,
 * no contraction), mirroring what SSE scalar code does for the reference's expressions. */
static int half_overlap(float occ0, float occ1, int region, int level) {
  volatile float r0 = (float)region / (float)level;
  volatile float r1 = (float)(region + 1) / (float)level;
  volatile float lo = occ0 > r0 ? occ0 : r0;
  volatile float hi = occ1 < r1 ? occ1 : r1;
  volatile float num = hi - lo;
  volatile float den = occ1 - occ0;
  volatile float q = num / den;
  return q >= 0.5f;
}

/* One word of length n (not NUL-terminated) -> out[604].  Returns -1 on success, else the
 * position of the first character outside [a-z0-9] (out is then all zeros). */
int phoc_oracle_word(const char* w, int n, float* out) {
  memset(out, 0, PHOC_DIM * sizeof(float));
  for (int k = 0; k < n; ++k)
    if (find_unigram(w[k]) < 0) return k;
  for (int k = 0; k < n; ++k) {
    volatile float occ0 = (float)k / (float)n;
    volatile float occ1 = (float)(k + 1) / (float)n;
    const int u = find_unigram(w[k]);
    int base = 0;
    for (int level = 2; level <= 5; ++level) {
      for (int r = 0; r < level; ++r)
        if (half_overlap(occ0, occ1, r, level)) out[base + r * 36 + u] = 1.0f;
      base += level * 36;
    }
  }
  for (int k = 0; k + 1 < n; ++k) {
    const int b = find_bigram(w + k);
    if (b < 0) continue;
    volatile float occ0 = (float)k / (float)n;
    volatile float occ1 = (float)(k + 2) / (float)n;
    for (int r = 0; r < 2; ++r)
      if (half_overlap(occ0, occ1, r, 2)) out[504 + r * 50 + b] = 1.0f;
  }
  return -1;
}

/* Batch: word i = chars[offsets[i] .. offsets[i+1]).  Returns the index of the first word with an
 * unknown unigram (its row is zero), or -1. */
long long phoc_oracle_batch(const char* chars, const int32_t* offsets, long long n, float* out) {
  long long first_bad = -1;
  for (long long i = 0; i < n; ++i) {
    const int bad = phoc_oracle_word(chars + offsets[i], offsets[i + 1] - offsets[i],
                                     out + i * PHOC_DIM);
    if (bad >= 0 && first_bad < 0) first_bad = i;
  }
  return first_bad;
}
