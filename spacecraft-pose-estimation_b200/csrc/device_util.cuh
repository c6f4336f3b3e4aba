// Per-device bookkeeping shared by the launchers: SM count and "configure this kernel once per
// device" (function attributes are per device; the library may be used from several threads and,
// in principle, on several devices of one process).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

namespace spe {

constexpr int kMaxDevices = 64;

inline cudaError_t current_device(int& dev, int& num_sms) {
  static std::atomic<int> sms[kMaxDevices];  // 0 = not queried yet
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= kMaxDevices) return cudaErrorInvalidDevice;
  num_sms = sms[dev].load(std::memory_order_relaxed);
  if (num_sms == 0) {
    e = cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    sms[dev].store(num_sms, std::memory_order_relaxed);
  }
  return cudaSuccess;
}

// run `configure()` the first time a kernel is launched on a device (idempotent if two threads race)
class PerDeviceOnce {
 public:
  template <typename F>
  cudaError_t run(int dev, F configure) {
    const uint64_t bit = 1ull << dev;
    if (done_.load(std::memory_order_acquire) & bit) return cudaSuccess;
    const cudaError_t e = configure();
    if (e == cudaSuccess) done_.fetch_or(bit, std::memory_order_release);
    return e;
  }

 private:
  std::atomic<uint64_t> done_{0};
};

}  // namespace spe
