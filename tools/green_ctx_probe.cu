// Probe: do CUDA green contexts (SM partitions) isolate a chain of tiny latency-critical kernels (the LocalBA trial
// sequence) from a throughput kernel stream (the ORB front-end)?   nvcc -arch=sm_100a tools/green_ctx_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <vector>
#define CK(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char* s; cuGetErrorString(r_, &s); printf("%s -> %s\n", #x, s); return 1; } } while (0)
#define RK(x) do { cudaError_t r_ = (x); if (r_ != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(r_)); return 1; } } while (0)

__global__ void hog(float* out, int iters) {  // full-occupancy CTAs, ~tens of microseconds each
  float a = threadIdx.x;
  for (int i = 0; i < iters; ++i) a = a * 1.0001f + 0.5f;
  if (a == 123.f) out[0] = a;
}
__global__ void tiny(double* buf) {  // one CTA, 1024 threads, most of an SM's shared memory: like k_ba_chol
  extern __shared__ double sm[];
  sm[threadIdx.x] = buf[threadIdx.x];
  __syncthreads();
  buf[threadIdx.x] = sm[(threadIdx.x + 1) & 1023] + 1.0;
}

static double chain_ms(cudaStream_t s, double* buf, int n) {
  cudaStreamSynchronize(s);
  auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < n; ++i) tiny<<<1, 1024, 200 * 1024, s>>>(buf);
  cudaStreamSynchronize(s);
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

int main() {
  RK(cudaSetDevice(0));
  RK(cudaFree(0));
  RK(cudaFuncSetAttribute(tiny, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CUdevice dev; CK(cuDeviceGet(&dev, 0));
  CUdevResource all; CK(cuDeviceGetDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM));
  printf("device SMs: %u\n", all.sm.smCount);
  CUdevResource grp[1], rest; unsigned n = 1;
  CK(cuDevSmResourceSplitByCount(grp, &n, &all, &rest, 0, 16));
  printf("split: group %u SMs, remaining %u SMs (groups %u)\n", grp[0].sm.smCount, rest.sm.smCount, n);
  CUdevResourceDesc d0, d1;
  CK(cuDevResourceGenerateDesc(&d0, &grp[0], 1));
  CK(cuDevResourceGenerateDesc(&d1, &rest, 1));
  CUgreenCtx g0, g1;
  CK(cuGreenCtxCreate(&g0, d0, dev, CU_GREEN_CTX_DEFAULT_STREAM));
  CK(cuGreenCtxCreate(&g1, d1, dev, CU_GREEN_CTX_DEFAULT_STREAM));
  CUstream sB, sA;
  CK(cuGreenCtxStreamCreate(&sB, g0, CU_STREAM_NON_BLOCKING, 0));
  CK(cuGreenCtxStreamCreate(&sA, g1, CU_STREAM_NON_BLOCKING, 0));
  int lo, hi; RK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  cudaStream_t pA, pB;
  RK(cudaStreamCreateWithPriority(&pA, cudaStreamNonBlocking, lo));
  RK(cudaStreamCreateWithPriority(&pB, cudaStreamNonBlocking, hi));
  float* out; double* buf;
  RK(cudaMalloc(&out, 4)); RK(cudaMalloc(&buf, 8 * 1024)); RK(cudaMemset(buf, 0, 8 * 1024));
  const int N = 300;
  printf("chain of %d tiny kernels alone: plain %.2f ms, green %.2f ms\n", N, chain_ms(pB, buf, N), chain_ms((cudaStream_t)sB, buf, N));
  for (int mode = 0; mode < 2; ++mode) {
    cudaStream_t a = mode ? (cudaStream_t)sA : pA, b = mode ? (cudaStream_t)sB : pB;
    // keep ~40 ms of hog work queued: 20 launches of 148*16 CTAs x 1024 threads
    for (int i = 0; i < 20; ++i) hog<<<148 * 16, 1024, 0, a>>>(out, 20000);
    const double ms = chain_ms(b, buf, N);
    auto t0 = std::chrono::steady_clock::now();
    cudaStreamSynchronize(a);
    const double rest_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    printf("%s: chain under load %.2f ms (hog drained %.2f ms later)\n", mode ? "green contexts (16 | 132 SMs)" : "priority streams", ms, rest_ms);
  }
  // a CUDA graph captured on the green stream
  cudaGraph_t g; cudaGraphExec_t ge;
  RK(cudaStreamBeginCapture((cudaStream_t)sB, cudaStreamCaptureModeThreadLocal));
  for (int i = 0; i < 8; ++i) tiny<<<1, 1024, 200 * 1024, (cudaStream_t)sB>>>(buf);
  RK(cudaStreamEndCapture((cudaStream_t)sB, &g));
  RK(cudaGraphInstantiate(&ge, g, 0));
  for (int i = 0; i < 20; ++i) hog<<<148 * 16, 1024, 0, (cudaStream_t)sA>>>(out, 20000);
  auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < 40; ++i) RK(cudaGraphLaunch(ge, (cudaStream_t)sB));
  RK(cudaStreamSynchronize((cudaStream_t)sB));
  printf("green + graph: 40 x 8-node graph under load %.2f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
  RK(cudaDeviceSynchronize());
  printf("PROBE_OK\n");
  return 0;
}
