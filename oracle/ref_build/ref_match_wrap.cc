// ORACLE — TEST INFRASTRUCTURE ONLY.
// C interface over two self-contained member functions of the REFERENCE's ORBmatcher, compiled UNCHANGED:
// ORBmatcher::ComputeThreeMaxima (src/ORBmatcher.cc:1608-1641) and ORBmatcher::DescriptorDistance (:1645-1667).
// The rest of that translation unit needs Frame / KeyFrame / MapPoint / Eigen / DBoW2 and cannot be built here, so the
// recipe (Makefile) cuts exactly these two function definitions out of /root/reference/src/ORBmatcher.cc into
// oracle/_ref/gen/orbmatcher_fns.inc (git-ignored build output) and this file supplies only the class declaration
// they need.
#include <opencv2/core/core.hpp>
#include <stdint.h>
#include <vector>
using namespace std;
namespace VIEO_SLAM {
class ORBmatcher {  // declaration subset of include/ORBmatcher.h:18-113
 public:
  static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);
  void ComputeThreeMaxima(std::vector<int>* histo, const int L, int& ind1, int& ind2, int& ind3);
};
#include "orbmatcher_fns.inc"
}  // namespace VIEO_SLAM

extern "C" {
int ref_descriptor_distance(const uint8_t* a, const uint8_t* b) {
  cv::Mat ma(1, 32, CV_8UC1, (void*)a), mb(1, 32, CV_8UC1, (void*)b);
  return VIEO_SLAM::ORBmatcher::DescriptorDistance(ma, mb);
}
// histo_sizes[L]: number of entries per rotation bin
void ref_three_maxima(const int* histo_sizes, int L, int* ind) {
  std::vector<std::vector<int>> histo(L);
  for (int i = 0; i < L; i++) histo[i].assign(histo_sizes[i], 0);
  int i1 = -1, i2 = -1, i3 = -1;
  VIEO_SLAM::ORBmatcher m;
  m.ComputeThreeMaxima(histo.data(), L, i1, i2, i3);
  ind[0] = i1; ind[1] = i2; ind[2] = i3;
}
}
