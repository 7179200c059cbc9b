// Internal declarations shared by the b200df translation units (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <algorithm>
#include <cstring>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/b200df.h"

namespace b200df
{
void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

#define B200DF_CUDA(expr)                                                                                          \
  do                                                                                                               \
  {                                                                                                                \
    cudaError_t e__ = (expr);                                                                                      \
    if (e__ != cudaSuccess)                                                                                        \
    {                                                                                                              \
      b200df::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(e__));             \
      return B200DF_ERR_CUDA;                                                                                      \
    }                                                                                                              \
  } while (0)

#define B200DF_REQUIRE(cond, msg)                                                                                  \
  do                                                                                                               \
  {                                                                                                                \
    if (!(cond))                                                                                                   \
    {                                                                                                              \
      b200df::set_error("%s (%s) at %s:%d", msg, #cond, __FILE__, __LINE__);                                       \
      return B200DF_ERR_INVALID;                                                                                   \
    }                                                                                                              \
  } while (0)

#define B200DF_LAUNCHED() (b200df::g_launches.fetch_add(1, std::memory_order_relaxed))

// RAII device switch
struct DeviceGuard
{
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev)
  {
    if (cudaGetDevice(&prev) != cudaSuccess)
      ok = false;
    else if (prev != dev && cudaSetDevice(dev) != cudaSuccess)
      ok = false;
  }
  ~DeviceGuard()
  {
    if (prev >= 0)
      cudaSetDevice(prev);
  }
};

int current_device();  // device chosen by b200df_set_device on this thread (default: CUDA current device)

// handles created inside the scope live on `dev`; the thread's own choice (b200df_set_device or none) comes back after
struct ThreadDeviceScope
{
  int prev_choice = -1, prev_cuda = -1;
  cudaError_t status = cudaSuccess;
  explicit ThreadDeviceScope(int dev);
  ~ThreadDeviceScope();
};

// host copy INTO a page-locked staging buffer that the GPU's copy engine reads next: non-temporal stores skip the
// read-for-ownership of the destination lines
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
__attribute__((target("avx2"))) inline void stream_copy_avx2(char* dst, const char* src, size_t bytes)
{
  const size_t head = std::min(bytes, (size_t)((32 - ((uintptr_t)dst & 31)) & 31));
  std::memcpy(dst, src, head);
  size_t k = head;
  for (; k + 32 <= bytes; k += 32)
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + k), _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + k)));
  _mm_sfence();
  std::memcpy(dst + k, src + k, bytes - k);
}
#endif
inline void stream_copy(void* dst, const void* src, size_t bytes)
{
#if defined(__x86_64__) && defined(__GNUC__)
  static const bool avx2 = __builtin_cpu_supports("avx2");
  if (avx2 && bytes >= 4096)
  {
    stream_copy_avx2(static_cast<char*>(dst), static_cast<const char*>(src), bytes);
    return;
  }
#endif
  std::memcpy(dst, src, bytes);
}

// Bounds-checked debug build (-DB200DF_BOUNDS_CHECK, `python tools/build_variant.py checked all -DB200DF_BOUNDS_CHECK`):
// every computed index into a voxel array or a device-side list is tested before use and a violation stops the kernel
// (printf + trap, which the host sees as a launch failure).  compute-sanitizer is closed on the pool this was developed
// on; tests/test_gpu_parity.py::test_bounds_checked_build_runs_every_kernel_family runs tools/sanitize_smoke.py on this
// build.  Compiled out otherwise.
#ifdef B200DF_BOUNDS_CHECK
#define B200DF_CHECK_INDEX(idx, limit)                                                                                 \
  do                                                                                                                   \
  {                                                                                                                    \
    if (!((long long)(idx) >= 0 && (long long)(idx) < (long long)(limit)))                                             \
    {                                                                                                                  \
      printf("b200df bounds check failed: %s:%d index %lld limit %lld\n", __FILE__, __LINE__, (long long)(idx),        \
             (long long)(limit));                                                                                      \
      __trap();                                                                                                        \
    }                                                                                                                  \
  } while (0)
#else
#define B200DF_CHECK_INDEX(idx, limit) ((void)0)
#endif

// Geometry of a voxel grid as the kernels see it (voxel_grid.h:367-399).
struct GridDev
{
  double om[3];  // origin_minus_ = origin - 0.5*resolution
  double oo;     // oo_resolution_ = 1/resolution
  int n[3];
  long stride1, stride2;  // ny*nz, nz
};

// What the validity / gradient kernels read of a field.
struct FieldDev
{
  GridDev g;
  const void* d2;       // uint8_t or uint16_t per voxel
  const void* neg;      // same type, nullptr when the field is unsigned
  const double* sqrt_table;  // sqrt(i)*resolution, i in [0, max_d2]
  int elem_size;
  int max_d2;
  double max_distance;  // getUninitializedDistance()
  int inv_twice_res;    // int(1/(2*res)) — Q3
};
}  // namespace b200df

struct b200df_field
{
  int device = 0;
  b200df_field_desc desc{};
  b200df::GridDev grid{};
  long n_voxels = 0;
  int max_d2 = 0;
  int radius = 0;  // smallest R with R*R >= max_d2
  int elem_size = 1;
  int inv_twice_res = 0;
  uint8_t* occ = nullptr;      // 1 = obstacle
  uint8_t* scratch8 = nullptr; // per-voxel marks for updatePointsInField
  void* d2 = nullptr;
  void* neg = nullptr;
  uint16_t* tmp_a = nullptr;
  uint16_t* tmp_b = nullptr;
  uint32_t* bits = nullptr;    // occupancy bit rows (1 bit per voxel) of the fast EDT path
  bool bits_valid = false;     // bits mirror occ (kept up to date by reset / add_points, else re-packed at rebuild)
  double* sqrt_table = nullptr;
  std::vector<double> sqrt_table_host;
  bool dirty = true;
  bool uploaded = false;       // d2 holds caller-supplied values (b200df_field_upload_d2), not the engine's exact EDT
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  static constexpr int kEdtStreams = 4;  // z slabs of one rebuild run concurrently (edt.cu)
  cudaStream_t edt_streams[kEdtStreams] = { nullptr, nullptr, nullptr, nullptr };
  cudaEvent_t ev_fork = nullptr, ev_join[kEdtStreams] = { nullptr, nullptr, nullptr, nullptr };
  std::mutex mu;
  float last_rebuild_ms = 0.f;
  // staging of host point lists (add / remove / update points): device buffer + page-locked buffer, grown on demand
  // and kept, so that a 1 M-point cloud costs neither a cudaMalloc / cudaFree pair nor the driver's pageable path
  double* pts_dev = nullptr;
  size_t pts_dev_bytes = 0;
  char* pts_pin = nullptr;
  size_t pts_pin_bytes = 0;
};

namespace b200df
{
// 32-bit words per bit row of the occupancy mirror: rows are padded to 16 bytes, the stride TMA needs (edt_tma.cu)
inline int bits_wpr(int nz)
{
  return ((nz + 127) / 128) * 4;
}
// builds the EDT when dirty; fills the kernel view.  Caller holds no lock.
int field_prepare(b200df_field* f, FieldDev* view, cudaStream_t wait_on);
// edt.cu: the two-launch exact EDT from bit rows (even nz, clamp radius <= 31); field.cu keeps the windowed fallback
bool edt_fast_applies(const b200df_field* f);
int edt_fast(b200df_field* f, bool invert, void* out);
// edt_tma.cu: the same two passes with TMA-moved tiles (nz a multiple of 16); -1 = tensor map not available
void keep_async_pool(int device);  // validity.cu: the stream-ordered pool keeps its memory between calls
bool edt_tma_applies(const b200df_field* f);
int edt_tma(b200df_field* f, void* out);
}
