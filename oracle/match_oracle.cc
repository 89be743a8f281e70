// ORACLE — TEST INFRASTRUCTURE ONLY (see orb_oracle.cc header).  CPU restatement of the Hamming paths.
#include <cstdint>
#include <cstring>
#include <climits>
#include "oracle.h"

extern "C" {

// ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:1645-1667): popcount of the XOR over 8 x u32.
int orc_descriptor_distance(const uint8_t* a, const uint8_t* b) {
  int d = 0;
  for (int i = 0; i < 8; ++i) {
    uint32_t x, y;
    memcpy(&x, a + 4 * i, 4);
    memcpy(&y, b + 4 * i, 4);
    uint32_t v = x ^ y;  // the reference's SWAR bit trick, same value as a popcount
    v = v - ((v >> 1) & 0x55555555u);
    v = (v & 0x33333333u) + ((v >> 2) & 0x33333333u);
    d += (int)((((v + (v >> 4)) & 0xF0F0F0Fu) * 0x1010101u) >> 24);
  }
  return d;
}

// cv::BFMatcher(NORM_HAMMING).knnMatch(q, t, k=2) as used by Frame::ComputeStereoFishEyeMatches
// (src/Frame.cc:620-628): two smallest distances ascending, ties -> lowest train index first.
// idx/dist are [nq][2]; missing neighbours (nt<2) are -1 / INT_MAX.  Pinned against cv2 (golden bf_*).
void orc_hamming_knn2(const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* idx, int32_t* dist) {
  for (int i = 0; i < nq; ++i) {
    int b0 = INT_MAX, b1 = INT_MAX, i0 = -1, i1 = -1;
    for (int j = 0; j < nt; ++j) {
      int d = orc_descriptor_distance(q + 32 * (size_t)i, t + 32 * (size_t)j);
      if (d < b0) {
        b1 = b0; i1 = i0; b0 = d; i0 = j;
      } else if (d < b1) {
        b1 = d; i1 = j;
      }
    }
    idx[2 * i] = i0; idx[2 * i + 1] = i1; dist[2 * i] = b0; dist[2 * i + 1] = b1;
  }
}

// Inner loop shared by the guided searches (src/ORBmatcher.cc:286-315, 1396-1424; src/Frame.cc:506-523):
// for query row r, scan its candidate list cand[row_ptr[r]..row_ptr[r+1]) in order, strict '<' keeps the
// first of equal distances; report best/second-best distance and the best candidate's train index.
// qsel[r] = descriptor row of the query in q (e.g. a MapPoint descriptor).
void orc_hamming_csr(const uint8_t* q, const uint8_t* t, const int32_t* row_ptr, const int32_t* cand, int nrows,
                     int32_t* best_dist, int32_t* best_idx, int32_t* second_dist, int32_t* second_idx) {
  for (int r = 0; r < nrows; ++r) {
    int b0 = 256, b1 = 256, i0 = -1, i1 = -1;
    for (int k = row_ptr[r]; k < row_ptr[r + 1]; ++k) {
      int d = orc_descriptor_distance(q + 32 * (size_t)r, t + 32 * (size_t)cand[k]);
      if (d < b0) {
        b1 = b0; i1 = i0; b0 = d; i0 = cand[k];
      } else if (d < b1) {
        b1 = d; i1 = cand[k];
      }
    }
    best_dist[r] = b0; best_idx[r] = i0; second_dist[r] = b1; second_idx[r] = i1;
  }
}

}  // extern "C"
