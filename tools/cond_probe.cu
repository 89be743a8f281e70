// Probe of the mechanism behind optimize()-as-one-graph (csrc/ba.cu): a CUDA graph whose only node is a WHILE
// conditional; the body is a captured stream of two kernels, the second one decides on the device whether the loop runs
// again (cudaGraphSetConditional).  Prints the iteration count reached with ONE cudaGraphLaunch and the time per
// iteration.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tools/cond_probe tools/cond_probe.cu (run on the box).
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e_ = (x);                                                            \
    if (e_ != cudaSuccess) {                                                         \
      std::printf("%s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
      return 1;                                                                      \
    }                                                                                \
  } while (0)

__global__ void k_work(double* x) {  // stands for the LM trial: something ordered before the control kernel
  x[threadIdx.x] = x[threadIdx.x] * 0.5 + 1.0;
}
__global__ void k_control(int* iter, int limit, cudaGraphConditionalHandle cond) {
  const int it = ++*iter;
  cudaGraphSetConditional(cond, it < limit ? 1 : 0);
}

int main() {
  const int limit = 37;
  double* x;
  int* iter;
  CK(cudaMalloc(&x, 32 * sizeof(double)));
  CK(cudaMalloc(&iter, sizeof(int)));
  CK(cudaMemset(x, 0, 32 * sizeof(double)));
  CK(cudaMemset(iter, 0, sizeof(int)));
  cudaStream_t st;
  CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  cudaGraph_t g, body, tmp;
  cudaGraphConditionalHandle hc;
  CK(cudaGraphCreate(&g, 0));
  CK(cudaGraphConditionalHandleCreate(&hc, g, 1, cudaGraphCondAssignDefault));  // default 1: the body runs at least once
  cudaGraphNodeParams np = {};
  np.type = cudaGraphNodeTypeConditional;
  np.conditional.handle = hc;
  np.conditional.type = cudaGraphCondTypeWhile;
  np.conditional.size = 1;
  cudaGraphNode_t node;
  CK(cudaGraphAddNode(&node, g, nullptr, 0, &np));
  body = np.conditional.phGraph_out[0];
  CK(cudaStreamBeginCaptureToGraph(st, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
  k_work<<<1, 32, 0, st>>>(x);
  k_control<<<1, 1, 0, st>>>(iter, limit, hc);
  CK(cudaStreamEndCapture(st, &tmp));
  cudaGraphExec_t exec;
  CK(cudaGraphInstantiate(&exec, g, 0));
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  CK(cudaGraphLaunch(exec, st));  // warm-up
  CK(cudaStreamSynchronize(st));
  CK(cudaMemsetAsync(iter, 0, sizeof(int), st));
  CK(cudaEventRecord(a, st));
  CK(cudaGraphLaunch(exec, st));
  CK(cudaEventRecord(b, st));
  CK(cudaStreamSynchronize(st));
  int n = 0;
  float ms = 0;
  CK(cudaMemcpy(&n, iter, sizeof(int), cudaMemcpyDeviceToHost));
  CK(cudaEventElapsedTime(&ms, a, b));
  std::printf("one graph launch ran %d iterations of the WHILE body (expected %d), %.2f us per iteration\n", n, limit,
              1e3f * ms / (n ? n : 1));
  return n == limit ? 0 : 2;
}
