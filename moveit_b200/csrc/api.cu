// Error reporting, device selection and bookkeeping shared by the b200df translation units.
#include "b200df_internal.cuh"

namespace b200df
{
static thread_local std::string t_error;
static thread_local int t_device = -1;
std::atomic<int64_t> g_launches{ 0 };

void set_error(const char* fmt, ...)
{
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  t_error = buf;
}

int current_device()
{
  if (t_device >= 0)
    return t_device;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess)
    dev = 0;
  return dev;
}
ThreadDeviceScope::ThreadDeviceScope(int dev)
{
  prev_choice = t_device;
  if (cudaGetDevice(&prev_cuda) != cudaSuccess)
    prev_cuda = -1;
  status = cudaSetDevice(dev);
  if (status == cudaSuccess)
    t_device = dev;
}

ThreadDeviceScope::~ThreadDeviceScope()
{
  t_device = prev_choice;
  if (prev_cuda >= 0)
    cudaSetDevice(prev_cuda);
}
}  // namespace b200df

extern "C" {

const char* b200df_last_error(void)
{
  return b200df::t_error.c_str();
}

int b200df_version(void)
{
  return B200DF_VERSION;
}

int b200df_device_count(int* count)
{
  B200DF_REQUIRE(count, "null argument");
  *count = 0;
  cudaError_t e = cudaGetDeviceCount(count);
  if (e != cudaSuccess)
  {
    *count = 0;
    b200df::set_error("cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
    return B200DF_ERR_CUDA;
  }
  return B200DF_OK;
}

int b200df_set_device(int device)
{
  int n = 0;
  B200DF_CUDA(cudaGetDeviceCount(&n));
  B200DF_REQUIRE(device >= 0 && device < n, "device index out of range");
  B200DF_CUDA(cudaSetDevice(device));
  b200df::t_device = device;
  return B200DF_OK;
}

int b200df_host_alloc(size_t bytes, void** out)
{
  B200DF_REQUIRE(out, "null argument");
  *out = nullptr;
  B200DF_CUDA(cudaHostAlloc(out, bytes > 0 ? bytes : 1, cudaHostAllocDefault));
  return B200DF_OK;
}

int b200df_host_free(void* p)
{
  if (p)
    B200DF_CUDA(cudaFreeHost(p));
  return B200DF_OK;
}

int64_t b200df_launch_count(void)
{
  return b200df::g_launches.load();
}

}  // extern "C"
