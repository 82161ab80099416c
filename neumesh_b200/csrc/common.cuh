// Shared helpers for the neumesh_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <mutex>
#include <string>

namespace nmb {

void set_error(const std::string& msg);
void count_launch(int n = 1);

#define NMB_CUDA_OK(expr)                                                                               \
  do {                                                                                                  \
    cudaError_t _e = (expr);                                                                            \
    if (_e != cudaSuccess) {                                                                            \
      ::nmb::set_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " (" + __FILE__ +  \
                       ":" + std::to_string(__LINE__) + ")");                                           \
      return 1;                                                                                         \
    }                                                                                                   \
  } while (0)

#define NMB_CHECK(cond, msg)                                                \
  do {                                                                      \
    if (!(cond)) {                                                          \
      ::nmb::set_error(std::string(msg) + " [" #cond "]");                  \
      return 2;                                                             \
    }                                                                       \
  } while (0)

#define NMB_LAUNCH_OK()                                      \
  do {                                                       \
    ::nmb::count_launch();                                   \
    NMB_CUDA_OK(cudaGetLastError());                         \
  } while (0)

// Keeps up to 1 GiB of freed stream-ordered scratch cached in the device's default memory pool (the default
// threshold of 0 hands every block back to the driver at the next synchronisation).  Once per device.
cudaError_t ensure_scratch_pool();

// Stream-ordered scratch (cudaMallocAsync) that is returned to the pool on every exit path of an API call.
struct StreamBuf {
  void* p = nullptr;
  cudaStream_t stream = nullptr;
  StreamBuf() = default;
  StreamBuf(const StreamBuf&) = delete;
  StreamBuf& operator=(const StreamBuf&) = delete;
  ~StreamBuf() {
    if (p) cudaFreeAsync(p, stream);
  }
  cudaError_t alloc(size_t bytes, cudaStream_t s) {
    stream = s;
    cudaError_t e = ensure_scratch_pool();
    if (e != cudaSuccess) return e;
    return cudaMallocAsync(&p, bytes ? bytes : 1, s);
  }
  template <typename T>
  T* as() const { return static_cast<T*>(p); }
};

// Run `f` (-> cudaError_t) until it has succeeded once per CUDA device of this process (kernel attributes are per
// device); thread-safe.
constexpr int NMB_MAX_DEVICES = 64;
struct DeviceOnce {
  std::mutex mu;
  bool done[NMB_MAX_DEVICES] = {};
  template <typename F>
  cudaError_t run(F&& f) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    bool& d = done[dev % NMB_MAX_DEVICES];
    if (d) return cudaSuccess;
    e = f();               // a failed attempt is reported AND retried by the next call (the flag is only set on success)
    if (e == cudaSuccess) d = true;
    return e;
  }
};

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t align_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

// Simple owning device buffer (build-time allocations; hot-path scratch comes from the caller's workspace).
template <typename T>
struct DevBuf {
  T* p = nullptr;
  int64_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  // (re)allocation; a buffer that already has `count` elements is kept as it is - a field that is re-packed every
  // training step must not pay cudaFree + cudaMalloc (both synchronise the device) for buffers of unchanged size
  cudaError_t alloc(int64_t count) {
    if (p && n == count) return cudaSuccess;
    release();
    n = count;
    if (count <= 0) return cudaSuccess;
    return cudaMalloc(reinterpret_cast<void**>(&p), sizeof(T) * static_cast<size_t>(count));
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
};

int sm_count();

// Optional per-kernel-class device timing (bench.py's roofline): CUDA events recorded on the launching stream
// around the launches of one class.  Disabled by default (no events, no overhead).
enum ProfTag { PROF_KNN = 0, PROF_BOUND = 1, PROF_GEO = 2, PROF_GEO_JVP = 3, PROF_COLOR = 4, PROF_SAMPLER = 5, PROF_KNN_LIST = 6, PROF_N = 7 };
bool prof_enabled();
void prof_begin(int tag, int64_t units, cudaStream_t stream);
void prof_end(int tag, cudaStream_t stream);
struct ProfScope {
  int tag;
  cudaStream_t stream;
  bool on;
  ProfScope(int t, int64_t units, cudaStream_t s) : tag(t), stream(s), on(prof_enabled()) {
    if (on) prof_begin(tag, units, stream);
  }
  ~ProfScope() {
    if (on) prof_end(tag, stream);
  }
};

}  // namespace nmb
