// Library-level plumbing of the C ABI: error string, version, device properties.
#include <cstdarg>
#include <cstdio>
#include "common.cuh"
#include "ruart_b200.h"

static thread_local char g_err[1024] = "";

void ruart_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* ruart_last_error(void) { return g_err; }

extern "C" int ruart_version(void) { return 100; }

extern "C" int ruart_num_sms(void) {
  static std::atomic<int> sms[RUART_MAX_DEVICES];
  const int dev = ruart_current_device();
  int v = sms[dev].load(std::memory_order_relaxed);
  if (v == 0) {
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0)
      return 148;
    sms[dev].store(v, std::memory_order_relaxed);
  }
  return v;
}
