// Library-wide state of the neumesh_b200 C ABI: error string, launch counter, device queries.
#include <atomic>

#include "../../include/neumesh_b200.h"
#include "common.cuh"

namespace nmb {
static thread_local std::string g_error;
static std::atomic<int64_t> g_launches{0};
void set_error(const std::string& msg) { g_error = msg; }
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
    if (cached <= 0) cached = 148;
  }
  return cached;
}
}  // namespace nmb

extern "C" {
const char* nmb_last_error(void) { return nmb::g_error.c_str(); }
int nmb_version(void) { return NMB_VERSION; }
int64_t nmb_launch_count(void) { return nmb::g_launches.load(); }
}
