// Error reporting and device selection for libvieo_b200.so.
#include <stdarg.h>

#include <cstdlib>

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

// Every handle owns a stream (one per camera extractor, per front-end chunk, per BA engine, per calling thread); with the
// default of 8 hardware work queues unrelated streams alias onto one queue and serialise (measured on B200: 8 LocalBA
// streams + front-end + tracking streams run 24 % faster with 32 queues).  The variable is read when the CUDA context is
// created, so it is set when the library is loaded, unless the user chose a value.
__attribute__((constructor)) static void vieo_more_hw_queues() { setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); }

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

CallScratch::~CallScratch() {
  if (device < 0) return;
  if (st) cudaStreamSynchronize(st);
  for (int k = 0; k < 12; ++k)
    if (buf[k]) cudaFree(buf[k]);
  if (hbuf) cudaFreeHost(hbuf);
  if (st) cudaStreamDestroy(st);
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

}  // namespace vieo

// ---------------------------------------------------------------------------------------------------------------
// SM partitions (CUDA green contexts).  The bundle-adjustment engines are chains of tiny, strictly ordered kernels
// (an LM trial is eight launches of a few CTAs); on a device that the ORB front-end keeps full they are starved: a
// stream priority only reorders CTAs that are waiting, it cannot free an SM, and the Cholesky CTA needs most of one
// (measured on B200, tools/green_ctx_probe.cu: a 300-kernel chain takes 1.2 ms alone, 31 ms beside a saturating kernel
// with priority streams, 1.3 ms with a 16-SM green context of its own).  A partition gives the BA streams `ba_sms` SMs
// and the front-end / tracking streams the remaining ones.  Driver entry points are resolved through the runtime
// (cudaGetDriverEntryPoint): no link-time dependency on libcuda.
#include <cuda.h>

#include <mutex>
#include <set>
struct vieo_sm_partition {
  int device = 0;
  CUgreenCtx ctx[2] = {nullptr, nullptr};  // [VIEO_SM_FRONTEND, VIEO_SM_BA]
  int sms[2] = {0, 0};
};
namespace {
struct DriverApi {
  CUresult (*DeviceGet)(CUdevice*, int) = nullptr;
  CUresult (*DeviceGetDevResource)(CUdevice, CUdevResource*, CUdevResourceType) = nullptr;
  CUresult (*DevSmResourceSplitByCount)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int,
                                        unsigned int) = nullptr;
  CUresult (*DevResourceGenerateDesc)(CUdevResourceDesc*, CUdevResource*, unsigned int) = nullptr;
  CUresult (*GreenCtxCreate)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int) = nullptr;
  CUresult (*GreenCtxDestroy)(CUgreenCtx) = nullptr;
  CUresult (*GreenCtxStreamCreate)(CUstream*, CUgreenCtx, unsigned int, int) = nullptr;
  bool ok = false;
};
const DriverApi& driver_api() {
  static DriverApi api = [] {
    DriverApi a;
    auto get = [](const char* name, void** fn) {
      cudaDriverEntryPointQueryResult q;
      return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess && *fn;
    };
    a.ok = get("cuDeviceGet", (void**)&a.DeviceGet) && get("cuDeviceGetDevResource", (void**)&a.DeviceGetDevResource) &&
           get("cuDevSmResourceSplitByCount", (void**)&a.DevSmResourceSplitByCount) &&
           get("cuDevResourceGenerateDesc", (void**)&a.DevResourceGenerateDesc) && get("cuGreenCtxCreate", (void**)&a.GreenCtxCreate) &&
           get("cuGreenCtxDestroy", (void**)&a.GreenCtxDestroy) && get("cuGreenCtxStreamCreate", (void**)&a.GreenCtxStreamCreate);
    return a;
  }();
  return api;
}
thread_local vieo_sm_partition* t_part = nullptr;
thread_local int t_which = 0;
// Partitions that exist: a thread's binding (t_part) is only a key into this set, so a thread that is still bound when another
// thread destroys the partition falls back to ordinary streams instead of dereferencing a freed object.
std::mutex g_part_mu;
std::set<vieo_sm_partition*> g_parts;
bool partition_alive(vieo_sm_partition* p) {
  std::lock_guard<std::mutex> lk(g_part_mu);
  return p && g_parts.count(p) != 0;
}
}  // namespace

namespace vieo {
cudaError_t make_stream(cudaStream_t* st, bool high_priority) {
  int lo = 0, hi = 0;
  cudaError_t e = cudaDeviceGetStreamPriorityRange(&lo, &hi);
  if (e != cudaSuccess) return e;
  if (t_part && !partition_alive(t_part)) t_part = nullptr;  // destroyed by another thread since this one was bound
  if (t_part && t_part->ctx[t_which]) {
    CUstream s = nullptr;
    if (driver_api().GreenCtxStreamCreate(&s, t_part->ctx[t_which], CU_STREAM_NON_BLOCKING, high_priority ? hi : 0) != CUDA_SUCCESS)
      return cudaErrorUnknown;
    *st = (cudaStream_t)s;
    return cudaSuccess;
  }
  return cudaStreamCreateWithPriority(st, cudaStreamNonBlocking, high_priority ? hi : 0);
}
}  // namespace vieo

extern "C" {
int vieo_sm_partition_create(int device, int ba_sms, vieo_sm_partition_t** out) {
  using namespace vieo;
  VIEO_ARG(out && ba_sms > 0, "bad argument");
  int rc = use_device(device);
  if (rc) return rc;
  VIEO_CK(cudaFree(0));
  const DriverApi& D = driver_api();
  if (!D.ok) {
    set_error("vieo_sm_partition_create: the driver does not export the green-context API");
    return VIEO_E_CUDA;
  }
  auto fail = [&](const char* what, CUresult r) {
    set_error("vieo_sm_partition_create: %s failed (CUresult %d)", what, (int)r);
    return VIEO_E_CUDA;
  };
  CUdevice dev;
  CUresult r = D.DeviceGet(&dev, device);
  if (r != CUDA_SUCCESS) return fail("cuDeviceGet", r);
  CUdevResource all, grp[1], rest;
  if ((r = D.DeviceGetDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM)) != CUDA_SUCCESS) return fail("cuDeviceGetDevResource", r);
  VIEO_ARG((unsigned)ba_sms < all.sm.smCount, "ba_sms must leave SMs for the front-end");
  unsigned n = 1;
  if ((r = D.DevSmResourceSplitByCount(grp, &n, &all, &rest, 0, (unsigned)ba_sms)) != CUDA_SUCCESS || n != 1)
    return fail("cuDevSmResourceSplitByCount", r);
  vieo_sm_partition* p = new vieo_sm_partition();
  p->device = device;
  CUdevResource* res[2] = {&rest, &grp[0]};
  for (int w = 0; w < 2; ++w) {
    CUdevResourceDesc desc;
    if ((r = D.DevResourceGenerateDesc(&desc, res[w], 1)) != CUDA_SUCCESS ||
        (r = D.GreenCtxCreate(&p->ctx[w], desc, dev, CU_GREEN_CTX_DEFAULT_STREAM)) != CUDA_SUCCESS) {
      vieo_sm_partition_destroy(p);
      return fail("cuGreenCtxCreate", r);
    }
    p->sms[w] = (int)res[w]->sm.smCount;
  }
  {
    std::lock_guard<std::mutex> lk(g_part_mu);
    g_parts.insert(p);
  }
  *out = p;
  return VIEO_OK;
}
void vieo_sm_partition_destroy(vieo_sm_partition_t* p) {
  if (!p) return;
  {
    std::lock_guard<std::mutex> lk(g_part_mu);
    g_parts.erase(p);  // other threads still bound to p see it gone at their next make_stream
  }
  if (t_part == p) t_part = nullptr;
  // Streams created inside the partition belong to handles / per-thread staging that the caller destroys first (documented in
  // the header); the device is drained so that no kernel of such a stream is in flight when the green contexts go away.
  cudaDeviceSynchronize();
  for (int w = 0; w < 2; ++w)
    if (p->ctx[w]) driver_api().GreenCtxDestroy(p->ctx[w]);
  delete p;
}
int vieo_sm_partition_sms(const vieo_sm_partition_t* p, int which) { return p && (which == 0 || which == 1) ? p->sms[which] : 0; }
int vieo_sm_partition_bind_thread(vieo_sm_partition_t* p, int which) {
  using namespace vieo;
  VIEO_ARG(which == VIEO_SM_FRONTEND || which == VIEO_SM_BA, "bad partition index");
  VIEO_ARG(p == nullptr || partition_alive(p), "the partition was destroyed");
  t_part = p;
  t_which = which;
  return VIEO_OK;
}
void* vieo_sm_partition_stream(vieo_sm_partition_t* p, int which, int high_priority) {
  if (!p || (which != 0 && which != 1)) return nullptr;
  vieo_sm_partition* keep = t_part;
  const int keep_w = t_which;
  t_part = p; t_which = which;
  cudaStream_t st = nullptr;
  if (cudaSetDevice(p->device) != cudaSuccess || vieo::make_stream(&st, high_priority != 0) != cudaSuccess) st = nullptr;
  t_part = keep; t_which = keep_w;
  return (void*)st;
}
}  // extern "C"

namespace vieo {
CallScratch* call_scratch(int device) {
  static thread_local CallScratch pool[16];
  if (device < 0 || device >= 16) return nullptr;
  CallScratch& c = pool[device];
  if (c.device != device) {
    if (make_stream(&c.st, false) != cudaSuccess) {
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

int vieo_set_host_sync(int device, int mode) {
  VIEO_ARG(mode >= 0 && mode <= 3, "mode must be 0 (auto), 1 (spin), 2 (yield) or 3 (blocking)");
  int rc = vieo::use_device(device);
  if (rc) return rc;
  const unsigned flags[4] = {cudaDeviceScheduleAuto, cudaDeviceScheduleSpin, cudaDeviceScheduleYield, cudaDeviceScheduleBlockingSync};
  VIEO_CK(cudaSetDeviceFlags(flags[mode]));
  return VIEO_OK;
}
}
