// common.cu -- error string, scratch buffers, device info for the C ABI (include/gcb200.h)
#include "gc_common.cuh"

namespace gc
{
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int DevBuf::reserve(size_t bytes)
{
  if (bytes <= cap)
    return GC_OK;
  size_t want = align_up(bytes + bytes / 2, 256);
  if (p)
    GC_CUDA(cudaFree(p));
  p = nullptr;
  cap = 0;
  GC_CUDA(cudaMalloc(&p, want));
  cap = want;
  return GC_OK;
}
void DevBuf::release()
{
  if (p)
    cudaFree(p);
  p = nullptr;
  cap = 0;
}
int PinBuf::reserve(size_t bytes)
{
  if (bytes <= cap)
    return GC_OK;
  size_t want = align_up(bytes + bytes / 2, 256);
  if (p)
    GC_CUDA(cudaFreeHost(p));
  p = nullptr;
  cap = 0;
  GC_CUDA(cudaMallocHost(&p, want));
  cap = want;
  return GC_OK;
}
void PinBuf::release()
{
  if (p)
    cudaFreeHost(p);
  p = nullptr;
  cap = 0;
}

int num_sms(int device)
{
  int v = 0;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || v <= 0)
    return 148;
  return v;
}
}  // namespace gc

extern "C" {

int gc_version(void) { return 100; }
const char* gc_last_error(void) { return gc::g_err; }

int gc_device_count(int* count)
{
  GC_REQUIRE(count != nullptr, "gc_device_count: count is NULL");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess)
  {
    *count = 0;
    gc::set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return GC_E_CUDA;
  }
  *count = n;
  return GC_OK;
}

int gc_device_info(int device, char* name, int len, int* sms, int* cc)
{
  cudaDeviceProp p;
  GC_CUDA(cudaGetDeviceProperties(&p, device));
  if (name && len > 0)
  {
    strncpy(name, p.name, (size_t)len - 1);
    name[len - 1] = 0;
  }
  if (sms)
    *sms = p.multiProcessorCount;
  if (cc)
    *cc = p.major * 10 + p.minor;
  return GC_OK;
}

}  // extern "C"
