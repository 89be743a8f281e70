// Error reporting and device selection for libvieo_b200.so.
#include <stdarg.h>

#include "common.cuh"

namespace vieo {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int use_device(int device) {
  static int s_count = -1;  // cached once a device has been seen
  int n = s_count;
  cudaError_t e = cudaSuccess;
  if (n <= 0) {
    e = cudaGetDeviceCount(&n);
    if (e == cudaSuccess && n > 0) s_count = n;
  }
  if (e != cudaSuccess || n == 0) {
    set_error("no CUDA device available (%s); libvieo_b200 has no CPU fallback",
              e == cudaSuccess ? "count = 0" : cudaGetErrorString(e));
    return VIEO_E_CUDA;
  }
  if (device < 0 || device >= n) {
    set_error("device %d out of range (0..%d)", device, n - 1);
    return VIEO_E_ARG;
  }
  VIEO_CK(cudaSetDevice(device));
  // the architecture check is cached per device: cudaGetDeviceProperties costs about a millisecond per call
  static int s_major[64], s_minor[64];
  static bool s_known[64];
  if (device >= 64 || !s_known[device]) {
    int major = 0, minor = 0;
    VIEO_CK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    VIEO_CK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
    if (device < 64) {
      s_major[device] = major;
      s_minor[device] = minor;
      s_known[device] = true;
    }
    if (major != 10) {
      set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, major, minor);
      return VIEO_E_CUDA;
    }
    return VIEO_OK;
  }
  if (s_major[device] != 10) {
    set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, s_major[device], s_minor[device]);
    return VIEO_E_CUDA;
  }
  return VIEO_OK;
}

void* CallScratch::get(int slot, size_t bytes) {
  if (bytes == 0) bytes = 1;
  if (cap[slot] >= bytes) return buf[slot];
  if (buf[slot]) cudaFree(buf[slot]);
  buf[slot] = nullptr;
  cap[slot] = 0;
  const size_t want = bytes + bytes / 2;
  if (cudaMalloc(&buf[slot], want) != cudaSuccess) {
    set_error("out of device memory (%zu bytes of call scratch)", want);
    return nullptr;
  }
  cap[slot] = want;
  return buf[slot];
}

void* CallScratch::get_pinned(size_t bytes) {
  if (bytes == 0) bytes = 1;
  if (hcap >= bytes) return hbuf;
  if (hbuf) cudaFreeHost(hbuf);
  hbuf = nullptr;
  hcap = 0;
  const size_t want = bytes + bytes / 2;
  if (cudaMallocHost(&hbuf, want) != cudaSuccess) {
    set_error("out of pinned host memory (%zu bytes of call staging)", want);
    return nullptr;
  }
  hcap = want;
  return hbuf;
}

CallScratch* call_scratch(int device) {
  static thread_local CallScratch pool[16];
  if (device < 0 || device >= 16) return nullptr;
  CallScratch& c = pool[device];
  if (c.device != device) {
    if (cudaStreamCreateWithFlags(&c.st, cudaStreamNonBlocking) != cudaSuccess) {
      set_error("cannot create a stream on device %d", device);
      return nullptr;
    }
    c.device = device;
  }
  return &c;
}

}  // namespace vieo

extern "C" {
const char* vieo_last_error(void) { return vieo::g_err; }
int vieo_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}
const char* vieo_version(void) { return "vieo_b200 0.1.0 (sm_100a)"; }
}
