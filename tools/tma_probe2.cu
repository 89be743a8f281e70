// Second probe: the CUDA programming guide's canonical 2-D TMA example, parameterised by element type and box.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap map, int x, int y, int bytes, uint8_t* out) {
  __shared__ alignas(128) uint8_t smem[16384];
  __shared__ alignas(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&bar)), "r"(1) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(s32(smem)), "l"((uint64_t)&map), "r"(x), "r"(y), "r"(s32(&bar)) : "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(bytes) : "memory");
  }
  asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(s32(&bar)) : "memory");
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = smem[i];
}
int main(int argc, char** argv) {
  const int esz = argc > 1 ? atoi(argv[1]) : 4, bw = argc > 2 ? atoi(argv[2]) : 32, bh = argc > 3 ? atoi(argv[3]) : 8;
  cudaSetDevice(0); cudaFree(0);
  const int W = 1024, H = 512;  // elements
  std::vector<uint8_t> h((size_t)W * H * esz);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)(i * 7 + i / 1000);
  uint8_t *d, *out; cudaMalloc(&d, h.size()); cudaMalloc(&out, 16384);
  cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H}, str[1] = {(cuuint64_t)W * esz};
  cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh}, es[2] = {1, 1};
  CUresult r = cuTensorMapEncodeTiled(&m, esz == 4 ? CU_TENSOR_MAP_DATA_TYPE_INT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, str, box, es,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  const int bytes = bw * bh * esz, x = argc > 4 ? atoi(argv[4]) : 16, y = 5;
  k<<<1, 128>>>(m, x, y, bytes, out);
  cudaError_t e = cudaDeviceSynchronize();
  int bad = -1;
  if (e == cudaSuccess) {
    std::vector<uint8_t> o(bytes);
    cudaMemcpy(o.data(), out, bytes, cudaMemcpyDeviceToHost);
    bad = 0;
    for (int rr = 0; rr < bh; ++rr)
      for (int q = 0; q < bw * esz; ++q) bad += o[rr * bw * esz + q] != h[((size_t)(y + rr) * W + x) * esz + q];
  }
  printf("elem %d B, box %dx%d: encode %d, run: %s, mismatches %d\n", esz, bw, bh, (int)r, cudaGetErrorString(e), bad);
  return 0;
}
