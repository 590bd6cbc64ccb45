// Library-wide state of libsan_b200.so: error string, device properties, launch counter.
#include <atomic>
#include <cstdarg>

#include "san_common.cuh"
#include "../../include/san_b200.h"

static thread_local char g_err[512] = "";
std::atomic<long long> g_san_launches{0};

void san_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int san_num_sms() {
  static int sms[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (sms[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    sms[dev] = v;
  }
  return sms[dev];
}

extern "C" {
const char* san_last_error(void) { return g_err; }
int san_version(void) { return 100; }
long long san_launch_count(void) { return g_san_launches.load(); }
}
