// Host-side plumbing shared by every C-ABI entry point: thread-local error string, launch checks and
// TMA tensor-map encoding (driver entry point fetched at run time, so the .so has no link-time
// dependency on libcuda and loads on a CPU-only box for the symbol/export tests).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

namespace pi3b {

enum : int {
  PI3B_OK = 0,
  PI3B_ERR_ARG = -1,      // bad shape / alignment / null pointer
  PI3B_ERR_CUDA = -2,     // CUDA runtime or launch failure
  PI3B_ERR_DRIVER = -3,   // tensor-map encode / driver entry point failure
  PI3B_ERR_WORKSPACE = -4 // caller-provided workspace too small
};

void set_error(const char* fmt, ...);
int check_launch(const char* what);

// 2-D row-major tensor [rows, cols] with row pitch `pitch_bytes`; box = [box_rows, box_cols];
// 128-byte swizzle (box_cols * elem_bytes must be 128). OOB elements read as zero.
// dtype: 0 = bf16, 1 = fp32.
int make_tmap_2d(CUtensorMap* out, const void* base, int dtype, uint64_t rows, uint64_t cols, uint64_t pitch_bytes,
                 uint32_t box_rows, uint32_t box_cols);

// Per-device state. Function attributes (opt-in shared memory) and the SM count belong to a device, not to the process:
// a process that drives several GPUs, or several host threads racing through their first call, must each see them set.
constexpr int kMaxDevices = 64;
int current_device();          // cudaGetDevice, clamped to [0, kMaxDevices)
int num_sms_of(int dev);       // cached cudaDevAttrMultiProcessorCount
struct DeviceOnce {
  std::atomic<unsigned char> flag[kMaxDevices] = {};
  // true until done(dev) has been called: concurrent first callers may both run the (idempotent) setup
  bool first_call_on(int dev) const { return flag[dev].load(std::memory_order_acquire) == 0; }
  void done(int dev) { flag[dev].store(1, std::memory_order_release); }
};

}  // namespace pi3b

#define PI3B_REQUIRE(cond, ...)          \
  do {                                   \
    if (!(cond)) {                       \
      pi3b::set_error(__VA_ARGS__);      \
      return pi3b::PI3B_ERR_ARG;         \
    }                                    \
  } while (0)
