// Probe: which form of a u8 tile TMA load works from a plain <<<>>> launch on B200 (k_fast_cells staging).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
#include <cstdlib>
struct Maps { CUtensorMap m[8]; };
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n.reg .pred P1;\nelect.sync _|P1, 0xffffffff;\nselp.u32 %0, 1, 0, P1;\n}\n" : "=r"(pred));
  return pred != 0;
}
template <int RANK>
__global__ void k(const __grid_constant__ Maps maps, const Maps* gmaps, int variant, int level, int x, int y, int z, int bytes, uint8_t* out, int box_w, int box_h) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 8192);
  bool me = threadIdx.x == 0;
  if (variant & 1) me = (threadIdx.x < 32) && elect_one();
  if (me) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory");
    const CUtensorMap* mp = (variant & 2) ? &gmaps->m[level] : ((variant & 4) ? &maps.m[0] : &maps.m[level]);
    if (RANK == 3)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(s32(smem)), "l"((uint64_t)mp), "r"(s32(bar)), "r"(x), "r"(y), "r"(z) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(s32(smem)), "l"((uint64_t)mp), "r"(s32(bar)), "r"(x), "r"(y) : "memory");
  }
  __syncthreads();
  asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(s32(bar)) : "memory");
  for (int i = threadIdx.x; i < box_w * box_h; i += blockDim.x) out[i] = smem[i];
}
static bool enc(CUtensorMap* m, int rank, uint8_t* base, int pitch, int rows, int nimg, int bw, int bh) {
  cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)rows, (cuuint64_t)nimg};
  cuuint64_t str[2] = {(cuuint64_t)pitch, (cuuint64_t)pitch * rows};
  cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1}, es[3] = {1, 1, 1};
  CUresult r = cuTensorMapEncodeTiled(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, base, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) printf("  encode rank %d box %dx%d failed: %d\n", rank, bw, bh, (int)r);
  return r == CUDA_SUCCESS;
}
int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  printf("variant %d (1: elect.sync in a converged warp, 2: descriptor in global memory, 4: static index)\n", variant);
  cudaSetDevice(0); cudaFree(0);
  Maps* gmaps; cudaMalloc(&gmaps, sizeof(Maps));
  const int pitch = 768, rows = 480, nimg = 2;
  std::vector<uint8_t> h((size_t)pitch * rows * nimg);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)(i * 7 + i / pitch);
  uint8_t *d, *out; cudaMalloc(&d, h.size()); cudaMalloc(&out, 65536);
  cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
  struct Case { int rank, bw, bh; } cases[] = {{2, 64, 44}, {3, 64, 44}, {2, 64, 32}, {2, 128, 44}, {3, 64, 16}, {3, 128, 44}};
  for (auto c : cases) {
    Maps maps; memset(&maps, 0, sizeof(maps));
    bool ok = true;
    for (int l = 0; l < 8; ++l) ok &= enc(&maps.m[l], c.rank, d, pitch, rows, nimg, c.bw, c.bh);
    if (!ok) continue;
    cudaMemcpy(gmaps, &maps, sizeof(maps), cudaMemcpyHostToDevice);
    for (int level : {0, 3}) {
      const int x = 12, y = 13, z = 1;
      if (c.rank == 3) k<3><<<1, 128, 8192 + 64>>>(maps, gmaps, variant, level, x, y, z, c.bw * c.bh, out, c.bw, c.bh);
      else k<2><<<1, 128, 8192 + 64>>>(maps, gmaps, variant, level, x, y, z, c.bw * c.bh, out, c.bw, c.bh);
      cudaError_t e = cudaDeviceSynchronize();
      int bad = -1;
      if (e == cudaSuccess) {
        std::vector<uint8_t> o((size_t)c.bw * c.bh);
        cudaMemcpy(o.data(), out, o.size(), cudaMemcpyDeviceToHost);
        bad = 0;
        for (int r = 0; r < c.bh; ++r)
          for (int q = 0; q < c.bw; ++q) bad += o[r * c.bw + q] != h[(size_t)(c.rank == 3 ? z : 0) * pitch * rows + (size_t)(y + r) * pitch + x + q];
      }
      printf("rank %d box %dx%d level %d: %s, mismatches %d\n", c.rank, c.bw, c.bh, level, cudaGetErrorString(e), bad);
      if (e != cudaSuccess) { printf("  (context lost; stopping)\n"); return 0; }
    }
  }
  return 0;
}
