// Stereo front-end: the Frame::Frame stereo constructor's hot work (src/Frame.cc:218-316) as one call over a
// batch of frames with HOST buffers: per-camera ORBextractor::operator() (:259-278) + the brute-force
// left->right knnMatch of ComputeStereoFishEyeMatches (:620-628).  The batch is cut into chunks that run on
// independent streams (own ORB scratch each), so the H2D copy of one chunk overlaps the kernels of another
// and the D2H of a third.  No CPU fallback.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "common.cuh"

struct vieo_frontend {
  int device;
  int max_frames;
  int chunk_frames;
  int cap;
  std::vector<vieo_orb_t*> orb;  // one per in-flight chunk
  std::vector<int32_t*> d_idx;   // [chunk_frames][cap][2]
  std::vector<int32_t*> d_dist;
  std::vector<float*> d_ur;  // rectified-stereo outputs per chunk: [chunk_frames][cap] uright | depth, then sad
  std::vector<int32_t*> d_sad;
  int last_frames = 0;
  int launches;
};

using namespace vieo;

extern "C" {

int vieo_frontend_create(const VieoOrbConfig* cfg, int max_frames, int device, vieo_frontend_t** out) {
  VIEO_ARG(cfg && out && max_frames >= 1, "bad argument");
  vieo_frontend* f = new vieo_frontend();
  f->device = device;
  f->max_frames = max_frames;
  int n_chunks = max_frames >= 8 ? 4 : 1;  // chunks on independent streams: copies of one overlap the kernels of another
  if (const char* e = getenv("VIEO_FE_CHUNKS")) n_chunks = std::max(1, std::min(atoi(e), max_frames));
  f->chunk_frames = (max_frames + n_chunks - 1) / n_chunks;
  VieoOrbConfig c = *cfg;
  c.max_batch = 2 * f->chunk_frames;
  for (int i = 0; i < n_chunks; ++i) {
    vieo_orb_t* o = nullptr;
    int rc = vieo_orb_create(&c, device, &o);
    if (rc) {
      vieo_frontend_destroy(f);
      return rc;
    }
    f->orb.push_back(o);
    f->cap = vieo_orb_max_keypoints(o);
    int32_t *a = nullptr, *b = nullptr;
    if (cudaMalloc(&a, sizeof(int32_t) * 2 * f->cap * f->chunk_frames) != cudaSuccess ||
        cudaMalloc(&b, sizeof(int32_t) * 2 * f->cap * f->chunk_frames) != cudaSuccess) {
      set_error("vieo_frontend_create: out of device memory");
      cudaFree(a);
      vieo_frontend_destroy(f);
      return VIEO_E_CUDA;
    }
    f->d_idx.push_back(a);
    f->d_dist.push_back(b);
  }
  *out = f;
  return VIEO_OK;
}

void vieo_frontend_destroy(vieo_frontend_t* f) {
  if (!f) return;
  cudaSetDevice(f->device);
  for (auto o : f->orb) vieo_orb_destroy(o);
  for (auto p : f->d_idx) cudaFree(p);
  for (auto p : f->d_dist) cudaFree(p);
  for (auto p : f->d_ur) cudaFree(p);
  for (auto p : f->d_sad) cudaFree(p);
  delete f;
}

int vieo_frontend_max_keypoints(const vieo_frontend_t* f) { return f ? f->cap : VIEO_E_ARG; }
int vieo_frontend_last_launches(const vieo_frontend_t* f) { return f ? f->launches : VIEO_E_ARG; }

int vieo_frontend_process(vieo_frontend_t* f, int n_frames, const uint8_t* imgs, int row_stride, VieoKeyPoint* kps,
                          uint8_t* desc, int32_t* n_kp, int32_t* match_idx, int32_t* match_dist) {
  VIEO_ARG(f && imgs && kps && desc && n_kp && match_idx && match_dist, "null argument");
  VIEO_ARG(n_frames >= 1 && n_frames <= f->max_frames, "n_frames exceeds max_frames");
  VIEO_CK(cudaSetDevice(f->device));
  const int cap = f->cap;
  f->launches = 0;
  f->last_frames = n_frames;
  int chunk = 0;
  // image geometry from the first ORB handle
  int32_t lw[16], lh[16];
  vieo_orb_get_tables(f->orb[0], nullptr, nullptr, nullptr, nullptr, nullptr, lw, lh);
  const size_t img_bytes = (size_t)row_stride * lh[0];
  for (int f0 = 0; f0 < n_frames; f0 += f->chunk_frames, ++chunk) {
    const int nf = std::min(f->chunk_frames, n_frames - f0);
    vieo_orb_t* o = f->orb[chunk];
    int rc = orb_enqueue_host(o, 2 * nf, imgs + (size_t)2 * f0 * img_bytes, img_bytes, row_stride);
    if (rc) return rc;
    VieoKeyPoint* dk;
    uint8_t* dd;
    int* dn;
    int ocap;
    cudaStream_t st;
    orb_dev_outputs(o, &dk, &dd, &dn, &ocap, &st);
    // pair p: query = left image 2p, train = right image 2p+1
    rc = vieo_hamming_knn2_batch_dev(dd, (size_t)2 * cap * 32, dn, cap, dd + (size_t)cap * 32, (size_t)2 * cap * 32,
                                     dn + 1, cap, 2, nf, f->d_idx[chunk], f->d_dist[chunk], st);
    if (rc) return rc;
    f->launches += vieo_orb_last_launches(o) + 1;
    VIEO_CK(cudaMemcpyAsync(kps + (size_t)2 * f0 * cap, dk, sizeof(VieoKeyPoint) * 2 * nf * cap, cudaMemcpyDeviceToHost, st));
    VIEO_CK(cudaMemcpyAsync(desc + (size_t)2 * f0 * cap * 32, dd, (size_t)32 * 2 * nf * cap, cudaMemcpyDeviceToHost, st));
    VIEO_CK(cudaMemcpyAsync(n_kp + 2 * f0, dn, sizeof(int) * 2 * nf, cudaMemcpyDeviceToHost, st));
    VIEO_CK(cudaMemcpyAsync(match_idx + (size_t)2 * f0 * cap, f->d_idx[chunk], sizeof(int32_t) * 2 * nf * cap,
                            cudaMemcpyDeviceToHost, st));
    VIEO_CK(cudaMemcpyAsync(match_dist + (size_t)2 * f0 * cap, f->d_dist[chunk], sizeof(int32_t) * 2 * nf * cap,
                            cudaMemcpyDeviceToHost, st));
  }
  for (int c = 0; c < chunk; ++c) {
    VieoKeyPoint* dk; uint8_t* dd; int* dn; int ocap; cudaStream_t st;
    orb_dev_outputs(f->orb[c], &dk, &dd, &dn, &ocap, &st);
    VIEO_CK(cudaStreamSynchronize(st));
  }
  return VIEO_OK;
}

int vieo_frontend_stereo_rectified(vieo_frontend_t* f, int n_frames, float bf, float min_z, float* uright, float* depth,
                                   int32_t* sad) {
  VIEO_ARG(f && uright && depth && sad, "null argument");
  VIEO_ARG(n_frames >= 1 && n_frames <= f->last_frames, "no processed frames to match");
  VIEO_CK(cudaSetDevice(f->device));
  const int cap = f->cap;
  if (f->d_ur.empty()) {
    for (size_t c = 0; c < f->orb.size(); ++c) {
      float* a = nullptr;
      int32_t* b = nullptr;
      if (cudaMalloc(&a, sizeof(float) * 2 * cap * f->chunk_frames) != cudaSuccess ||
          cudaMalloc(&b, sizeof(int32_t) * cap * f->chunk_frames) != cudaSuccess) {
        set_error("vieo_frontend_stereo_rectified: out of device memory");
        cudaFree(a);
        return VIEO_E_CUDA;
      }
      f->d_ur.push_back(a);
      f->d_sad.push_back(b);
    }
  }
  int chunk = 0;
  for (int f0 = 0; f0 < n_frames; f0 += f->chunk_frames, ++chunk) {
    const int nf = std::min(f->chunk_frames, n_frames - f0);
    VieoKeyPoint* dk;
    uint8_t* dd;
    int* dn;
    int ocap;
    cudaStream_t st;
    orb_dev_outputs(f->orb[chunk], &dk, &dd, &dn, &ocap, &st);
    float* ur = f->d_ur[chunk];
    float* dp = ur + (size_t)cap * f->chunk_frames;
    int rc = vieo_orb_stereo_match_dev(f->orb[chunk], nf, dk, dd, dn, cap, bf, min_z, ur, dp, f->d_sad[chunk], st);
    if (rc) return rc;
    f->launches += 1;
    VIEO_CK(cudaMemcpyAsync(uright + (size_t)f0 * cap, ur, sizeof(float) * nf * cap, cudaMemcpyDeviceToHost, st));
    VIEO_CK(cudaMemcpyAsync(depth + (size_t)f0 * cap, dp, sizeof(float) * nf * cap, cudaMemcpyDeviceToHost, st));
    VIEO_CK(cudaMemcpyAsync(sad + (size_t)f0 * cap, f->d_sad[chunk], sizeof(int32_t) * nf * cap, cudaMemcpyDeviceToHost, st));
  }
  for (int c = 0; c < chunk; ++c) {
    VieoKeyPoint* dk; uint8_t* dd; int* dn; int ocap; cudaStream_t st;
    orb_dev_outputs(f->orb[c], &dk, &dd, &dn, &ocap, &st);
    VIEO_CK(cudaStreamSynchronize(st));
  }
  return VIEO_OK;
}

}  // extern "C"
