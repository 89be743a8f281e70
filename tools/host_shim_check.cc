// Compile/link check of the C++ host shims against libvieo_b200.so (no GPU calls): g++ -std=c++17 ... -lvieo_b200
#include <cstdio>

#include "../vieo_slam_b200/host/vieo_shims.hpp"
int main() {
  using namespace VIEO_SLAM_B200;
  ORBextractor* e = nullptr;
  (void)e;
  ORBmatcher m(0.6f, true);
  IMUPreintegrator p;
  GlobalBA* g = nullptr;
  SmPartition* sp = nullptr;
  (void)g; (void)sp;
  std::vector<VieoSbpFrame> no_frames;
  std::vector<int32_t> a, b, c, d;
  VieoSbpQueries q{};
  if (m.SearchByProjection(VIEO_SBP_LAST_FRAME, no_frames, nullptr, nullptr, nullptr, q, nullptr, a, b, c, d) != 0) return 2;
  // the SURVEY 8(f) entry points: empty batches return without touching the device
  std::vector<VieoFrustumFrame> no_frustum;
  std::vector<VieoProjSearchFrame> no_kfs;
  std::vector<uint8_t> iv;
  std::vector<float> pr, vc, dp;
  std::vector<int32_t> lv, ni;
  if (m.SearchLocalPoints(no_frustum, no_frames, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                          nullptr, nullptr, nullptr, iv, pr, lv, vc, dp, ni, a, b, c, d) != 0) return 3;
  m.SearchByProjectionBase(no_kfs, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, a, b, c);
  m.ComputeDistinctiveDescriptors(nullptr, 0, {}, {0}, a, b);
  double bg[3] = {0, 0, 0};
  if (OptimizeInitialGyroBias({}, {}, bg) != 0) return 4;
  float tab[16];
  if (vieo_frustum_level_table(0.18232156f, 8, tab) != 0 || !(tab[1] > 1.0f && tab[1] < 1.0001f)) return 5;
  const double s2[4] = {1e-8, 4e-6, 1e-10, 9e-6};
  IMUPreintegrator::SetParam(s2, 1, 200.0);
  std::printf("%s %d %g %zu\n", vieo_version(), ORBmatcher::TH_HIGH, IMUPreintegrator::Noise().sigma_g, sizeof(p));
  return (IMUPreintegrator::Noise().freq_ref == 0 && m.mbCheckOrientation) ? 0 : 1;
}
