// Shared helpers for libvieo_b200.so (sm_100a only; no CPU fallback).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/vieo_b200.h"

namespace vieo {

void set_error(const char* fmt, ...);

#define VIEO_CK(call)                                                                          \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess) {                                                                   \
      vieo::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));    \
      return VIEO_E_CUDA;                                                                      \
    }                                                                                          \
  } while (0)

#define VIEO_ARG(cond, msg)                                   \
  do {                                                        \
    if (!(cond)) {                                            \
      vieo::set_error("%s:%d %s", __FILE__, __LINE__, msg);   \
      return VIEO_E_ARG;                                      \
    }                                                         \
  } while (0)

// selects `device` and checks it is a Blackwell (sm_100) part; there is no other code path.
int use_device(int device);

// Per-thread, per-device staging for the host-buffer entry points (vieo_imu_preint_batch, vieo_pose_opt_batch): a
// private non-blocking stream and device buffers that grow on demand and are reused by later calls, so a call costs
// copies + kernels, not cudaMalloc/cudaFree.  Slots are independent buffers; get() returns nullptr on failure.
struct CallScratch {
  int device = -1;
  cudaStream_t st = nullptr;
  void* buf[12] = {};
  size_t cap[12] = {};
  void* get(int slot, size_t bytes);
  // pinned host staging (grows on demand, reused): cudaMallocHost / cudaFreeHost per call would cost milliseconds and
  // cudaFreeHost synchronises the whole device
  void* hbuf = nullptr;
  size_t hcap = 0;
  void* get_pinned(size_t bytes);
  // a host thread that used a host-buffer entry point gives its stream and staging back when it exits (thread-local
  // destructors of the main thread run before the runtime's own teardown)
  ~CallScratch();
};
CallScratch* call_scratch(int device);

// Stream for a handle / call scratch created by the calling thread: inside the SM partition the thread is bound to
// (vieo_sm_partition_bind_thread), else an ordinary stream.  high_priority: the device's highest stream priority.
cudaError_t make_stream(cudaStream_t* st, bool high_priority);

// Opt a kernel into `bytes` of dynamic shared memory once per DEVICE (function attributes are per device / context; the
// C ABI serves several GPUs from one process).  `done` is the call site's own flag array; races only repeat the call.
struct SmemOptIn {
  bool done[64] = {};
};
template <class F>
inline cudaError_t smem_opt_in(F kernel, size_t bytes, SmemOptIn& flags) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 0 && dev < 64 && flags.done[dev]) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess && dev >= 0 && dev < 64) flags.done[dev] = true;
  return e;
}

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// Exclusive scan of a[0..n) in shared memory, in place, by warp 0 of the block; returns total via *total.
// Caller must __syncthreads() before (a written) and after (a read).
__device__ __forceinline__ void warp0_excl_scan(int* a, int n, int* total) {
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    int carry = 0;
    for (int base = 0; base < n; base += 32) {
      int i = base + lane;
      int v = i < n ? a[i] : 0;
      int inc = warp_incl_scan(v, lane);
      if (i < n) a[i] = carry + inc - v;
      carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) *total = carry;
  }
}

}  // namespace vieo

// internal cross-file hooks (comm.cu -> ba.cu)
struct vieo_comm;
namespace vieo {
int comm_allreduce_f64(vieo_comm* c, double* buf, size_t count, cudaStream_t stream);
void comm_info(const vieo_comm* c, int* rank, int* world);
}  // namespace vieo

// internal cross-file hooks (orb.cu -> frontend.cu)
struct vieo_orb;
namespace vieo {
int orb_enqueue_host(vieo_orb* h, int n_img, const uint8_t* imgs, size_t img_stride, int row_stride);
void orb_info(vieo_orb* h, int* device, int* max_batch);
void orb_dev_outputs(vieo_orb* h, VieoKeyPoint** kps, uint8_t** desc, int** nkp, int* cap, cudaStream_t* st);
}  // namespace vieo
