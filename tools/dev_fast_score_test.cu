// dev tool: candidate formulations of the FAST-9/16 arc score on the device vs a host loop (+ timing)
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define RING(p, pitch, F)                                                                                          \
  F(0, p[3 * pitch]) F(1, p[3 * pitch + 1]) F(2, p[2 * pitch + 2]) F(3, p[pitch + 3]) F(4, p[3]) F(5, p[-pitch + 3]) \
  F(6, p[-2 * pitch + 2]) F(7, p[-3 * pitch + 1]) F(8, p[-3 * pitch]) F(9, p[-3 * pitch - 1]) F(10, p[-2 * pitch - 2]) \
  F(11, p[-pitch - 3]) F(12, p[-3]) F(13, p[pitch - 3]) F(14, p[2 * pitch - 2]) F(15, p[3 * pitch - 1])

// V0: differences, int min/max trees (the first attempt: wrong on device)
__device__ __forceinline__ int score_v0(const uint8_t* p, int pitch) {
  const int v = p[0];
  int d[16];
#define F(i, e) d[i] = v - e;
  RING(p, pitch, F)
#undef F
  int lo2[16], hi2[16], lo4[16], hi4[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) { lo2[k] = min(d[k], d[(k + 1) & 15]); hi2[k] = max(d[k], d[(k + 1) & 15]); }
#pragma unroll
  for (int k = 0; k < 16; ++k) { lo4[k] = min(lo2[k], lo2[(k + 2) & 15]); hi4[k] = max(hi2[k], hi2[(k + 2) & 15]); }
  int best = 0;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int lo9 = min(min(lo4[k], lo4[(k + 4) & 15]), d[(k + 8) & 15]);
    const int hi9 = max(max(hi4[k], hi4[(k + 4) & 15]), d[(k + 8) & 15]);
    best = max(best, max(lo9, -hi9));
  }
  return min(best, 255);
}
// V1: raw pixel values, S = max(v - min_arcs(max9 p), max_arcs(min9 p) - v, 0)
__device__ __forceinline__ int score_v1(const uint8_t* p, int pitch) {
  const int v = p[0];
  int d[16];
#define F(i, e) d[i] = e;
  RING(p, pitch, F)
#undef F
  int lo2[16], hi2[16], lo4[16], hi4[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) { lo2[k] = min(d[k], d[(k + 1) & 15]); hi2[k] = max(d[k], d[(k + 1) & 15]); }
#pragma unroll
  for (int k = 0; k < 16; ++k) { lo4[k] = min(lo2[k], lo2[(k + 2) & 15]); hi4[k] = max(hi2[k], hi2[(k + 2) & 15]); }
  int a = 255, b = 0;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int lo9 = min(min(lo4[k], lo4[(k + 4) & 15]), d[(k + 8) & 15]);
    const int hi9 = max(max(hi4[k], hi4[(k + 4) & 15]), d[(k + 8) & 15]);
    a = min(a, hi9);
    b = max(b, lo9);
  }
  return max(max(v - a, b - v), 0);
}
// V2: two ring pixels per register (u16x2): lane lo = p[k], lane hi = p[k+8]; X[j+8] = halves of X[j] swapped
__device__ __forceinline__ int score_v2(const uint8_t* p, int pitch) {
  const int v = p[0];
  unsigned r[16];
#define F(i, e) r[i] = e;
  RING(p, pitch, F)
#undef F
  unsigned X[16];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    X[k] = r[k] | (r[k + 8] << 16);
    X[k + 8] = r[k + 8] | (r[k] << 16);
  }
  unsigned lo2[15], hi2[15], lo4[12], hi4[12];
#pragma unroll
  for (int j = 0; j < 15; ++j) { lo2[j] = __vminu2(X[j], X[j + 1]); hi2[j] = __vmaxu2(X[j], X[j + 1]); }
#pragma unroll
  for (int j = 0; j < 12; ++j) { lo4[j] = __vminu2(lo2[j], lo2[j + 2]); hi4[j] = __vmaxu2(hi2[j], hi2[j + 2]); }
  unsigned a = 0x00ff00ffu, b = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const unsigned lo9 = __vminu2(__vminu2(lo4[k], lo4[k + 4]), X[k + 8]);
    const unsigned hi9 = __vmaxu2(__vmaxu2(hi4[k], hi4[k + 4]), X[k + 8]);
    a = __vminu2(a, hi9);
    b = __vmaxu2(b, lo9);
  }
  const int A = min(a & 0xffff, a >> 16), B = max(b & 0xffff, b >> 16);
  return max(max(v - A, B - v), 0);
}

template <int V>
__global__ void k_score(const uint8_t* img, int w, int h, uint8_t* out) {
  int x = blockIdx.x * 128 + threadIdx.x, y = blockIdx.y;
  if (x < 3 || x >= w - 3 || y < 3 || y >= h - 3) return;
  const uint8_t* p = img + (size_t)y * w + x;
  int s = V == 0 ? score_v0(p, w) : V == 1 ? score_v1(p, w) : score_v2(p, w);
  out[(size_t)y * w + x] = (uint8_t)s;
}
static int host_score(const uint8_t* p, int pitch) {
  static const int ox[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
  static const int oy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
  int v = p[0], d[25], best = 0;
  for (int k = 0; k < 25; ++k) d[k] = v - p[oy[k % 16] * pitch + ox[k % 16]];
  for (int k = 0; k < 16; ++k) { int mn = d[k], mx = d[k]; for (int j = 1; j < 9; ++j) { mn = std::min(mn, d[k + j]); mx = std::max(mx, d[k + j]); } best = std::max(best, std::max(mn, -mx)); }
  return std::min(best, 255);
}
template <int V>
void run(const uint8_t* dimg, uint8_t* dout, const std::vector<uint8_t>& himg, int w, int h) {
  std::vector<uint8_t> hout((size_t)w * h);
  cudaMemset(dout, 0, (size_t)w * h);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_score<V><<<dim3((w + 127) / 128, h), 128>>>(dimg, w, h, dout);
  cudaEventRecord(e0);
  for (int i = 0; i < 10; ++i) k_score<V><<<dim3((w + 127) / 128, h), 128>>>(dimg, w, h, dout);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaMemcpy(hout.data(), dout, (size_t)w * h, cudaMemcpyDeviceToHost);
  long bad = 0;
  for (int y = 3; y < h - 3; ++y) for (int x = 3; x < w - 3; ++x) {
    int s = host_score(himg.data() + (size_t)y * w + x, w);
    if (s != hout[(size_t)y * w + x]) { if (bad++ < 3) printf("  V%d (%d,%d): host %d dev %d\n", V, x, y, s, hout[(size_t)y * w + x]); }
  }
  printf("V%d bad=%ld  %.3f ms/launch (%.1f Gpx/s) err=%s\n", V, bad, ms / 10, (double)w * h / (ms / 10) / 1e6, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  const int w = 4096, h = 2048;
  std::vector<uint8_t> himg((size_t)w * h);
  srand(3);
  for (auto& b : himg) b = rand() % 256;
  for (size_t i = 0; i < himg.size() / 2; ++i) himg[i] = 100 + rand() % 30;  // low-contrast half
  uint8_t *dimg, *dout;
  cudaMalloc(&dimg, himg.size()); cudaMalloc(&dout, himg.size());
  cudaMemcpy(dimg, himg.data(), himg.size(), cudaMemcpyHostToDevice);
  run<0>(dimg, dout, himg, w, h);
  run<1>(dimg, dout, himg, w, h);
  run<2>(dimg, dout, himg, w, h);
  return 0;
}
