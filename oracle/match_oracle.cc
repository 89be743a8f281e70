// ORACLE — TEST INFRASTRUCTURE ONLY (see orb_oracle.cc header).  CPU restatement of the Hamming paths.
#include <cstdint>
#include <cstring>
#include <climits>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <utility>
#include <vector>
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

// Frame::ComputeStereoMatches (src/Frame.cc:451-611): rectified-stereo association.  Right keypoints are bucketed by
// row band (+-2 scale), a left keypoint takes the arg-min Hamming (< TH_HIGH, strict '<' keeps the first) over the
// candidates of its row with octave within +-1 and uR in [uL - maxD, uL - minD]; matches under (TH_HIGH+TH_LOW)/2 are
// refined by an 11x11 L1 block search over +-5 px in the keypoint's pyramid level + parabola fit; finally matches whose
// SAD exceeds 1.5 * 1.4 * median are dropped.  pyrL / pyrR: tightly packed level images (lw[l] x lh[l]).
// Outputs per left keypoint: uright, depth (-1 = no match), sad (block distance, -1 = none).  Returns #matches kept.
extern "C" int orc_stereo_matches(const OrcKeyPoint* kl, const uint8_t* dl, int nl, const OrcKeyPoint* kr, const uint8_t* dr,
                                  int nr, const uint8_t* const* pyrL, const uint8_t* const* pyrR, const int* lw, const int* lh,
                                  const float* scale, const float* inv_scale, float bf, float minZ, float* uright,
                                  float* depth, int32_t* sad) {
  const int thOrbDist = (100 + 50) / 2;
  const int nRows = lh[0];
  std::vector<std::vector<int>> rows(nRows);
  for (int iR = 0; iR < nr; ++iR) {
    const float kpY = kr[iR].y;
    const float r = 2.0f * scale[kr[iR].octave];
    const int maxr = (int)std::ceil(kpY + r), minr = (int)std::floor(kpY - r);
    for (int yi = minr; yi <= maxr; ++yi)
      if (yi >= 0 && yi < nRows) rows[yi].push_back(iR);  // (the reference indexes unchecked; keypoints stay 16 px inside)
  }
  const float minD = 0, maxD = bf / minZ;
  std::vector<std::pair<int, int>> vDistIdx;
  for (int iL = 0; iL < nl; ++iL) {
    uright[iL] = -1.0f;
    depth[iL] = -1.0f;
    sad[iL] = -1;
  }
  for (int iL = 0; iL < nl; ++iL) {
    const OrcKeyPoint& kpL = kl[iL];
    const int levelL = kpL.octave;
    const float vL = kpL.y, uL = kpL.x;
    const int row = (int)vL;
    if (row < 0 || row >= nRows) continue;
    const std::vector<int>& cand = rows[row];
    if (cand.empty()) continue;
    const float minU = uL - maxD, maxU = uL - minD;
    if (maxU < 0) continue;
    int bestDist = 100;
    int bestIdxR = 0;
    for (int iR : cand) {
      const OrcKeyPoint& kpR = kr[iR];
      if (kpR.octave < levelL - 1 || kpR.octave > levelL + 1) continue;
      const float uR = kpR.x;
      if (uR >= minU && uR <= maxU) {
        const int dist = orc_descriptor_distance(dl + 32 * (size_t)iL, dr + 32 * (size_t)iR);
        if (dist < bestDist) {
          bestDist = dist;
          bestIdxR = iR;
        }
      }
    }
    if (!(bestDist < thOrbDist)) continue;
    const float uR0 = kr[bestIdxR].x;
    const float scaleFactor = inv_scale[levelL];
    const float scaleduL = std::round(kpL.x * scaleFactor);
    const float scaledvL = std::round(kpL.y * scaleFactor);
    const float scaleduR0 = std::round(uR0 * scaleFactor);
    const int w = 5, L = 5;
    const int W = lw[levelL];
    const uint8_t* IL = pyrL[levelL];
    const uint8_t* IR = pyrR[levelL];
    const int cu = (int)scaleduL, cv = (int)scaledvL, cr = (int)scaleduR0;
    const float iniu = scaleduR0 + L - w, endu = scaleduR0 + L + w + 1;
    if (iniu < 0 || endu >= W) continue;
    int bestSad = INT_MAX, bestincR = 0;
    float vDists[2 * 5 + 1];
    const int cL = IL[(size_t)cv * W + cu];
    for (int incR = -L; incR <= L; ++incR) {
      const int cR = IR[(size_t)cv * W + cr + incR];
      float dist = 0;  // cv::norm(IL, IR, NORM_L1) over float patches of integers: exact
      for (int dy = -w; dy <= w; ++dy)
        for (int dx = -w; dx <= w; ++dx) {
          const int a = IL[(size_t)(cv + dy) * W + cu + dx] - cL;
          const int b = IR[(size_t)(cv + dy) * W + cr + incR + dx] - cR;
          dist += (float)std::abs(a - b);
        }
      if (dist < (float)bestSad) {
        bestSad = (int)dist;
        bestincR = incR;
      }
      vDists[L + incR] = dist;
    }
    if (bestincR == -L || bestincR == L) continue;
    const float dist1 = vDists[L + bestincR - 1], dist2 = vDists[L + bestincR], dist3 = vDists[L + bestincR + 1];
    const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2));
    if (deltaR < -1 || deltaR > 1) continue;
    float bestuR = scale[levelL] * ((float)scaleduR0 + (float)bestincR + deltaR);
    float disparity = uL - bestuR;
    if (disparity >= minD && disparity < maxD) {
      if (disparity <= 0) {
        disparity = 0.01f;
        bestuR = uL - 0.01f;
      }
      depth[iL] = bf / disparity;
      uright[iL] = bestuR;
      sad[iL] = bestSad;
      vDistIdx.emplace_back(bestSad, iL);
    }
  }
  if (vDistIdx.empty()) return 0;
  std::sort(vDistIdx.begin(), vDistIdx.end());
  const float median = (float)vDistIdx[vDistIdx.size() / 2].first;
  const float thDist = 1.5f * 1.4f * median;
  int kept = (int)vDistIdx.size();
  for (int i = (int)vDistIdx.size() - 1; i >= 0; --i) {
    if ((float)vDistIdx[i].first < thDist) break;
    uright[vDistIdx[i].second] = -1;
    depth[vDistIdx[i].second] = -1;
    --kept;
  }
  return kept;
}

// MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:314-378) for a batch of map points in CSR form: point p owns
// the observed descriptors rows[ptr[p] .. ptr[p+1]) of desc_pool (rows == nullptr: the rows themselves, in order) in the
// reference's vDescriptors order.  All-pairs Hamming distances, per row the sorted-distance entry at index
// int(0.5 * (N - 1)), the first row with the least median wins.  best[p] = index into the point's list (-1: no
// observations, the reference returns without touching mDescriptor), median[p] = that row's median.
extern "C" void orc_distinctive_descriptors(const uint8_t* desc_pool, const int32_t* rows, const int32_t* ptr, int n_points,
                                            int32_t* best, int32_t* median) {
  std::vector<int> d, v;
  for (int p = 0; p < n_points; ++p) {
    const int b = ptr[p], N = ptr[p + 1] - b;
    best[p] = -1;
    median[p] = -1;
    if (N <= 0) continue;
    d.assign((size_t)N * N, 0);
    for (int i = 0; i < N; ++i)
      for (int j = i + 1; j < N; ++j) {
        const int ri = rows ? rows[b + i] : b + i, rj = rows ? rows[b + j] : b + j;
        d[(size_t)i * N + j] = d[(size_t)j * N + i] =
            orc_descriptor_distance(desc_pool + 32 * (size_t)ri, desc_pool + 32 * (size_t)rj);
      }
    int BestMedian = INT32_MAX, BestIdx = 0;
    for (int i = 0; i < N; ++i) {
      v.assign(d.begin() + (size_t)i * N, d.begin() + (size_t)(i + 1) * N);
      std::sort(v.begin(), v.end());
      const int med = v[(size_t)(0.5 * (N - 1))];
      if (med < BestMedian) {
        BestMedian = med;
        BestIdx = i;
      }
    }
    best[p] = BestIdx;
    median[p] = BestMedian;
  }
}

// Frame::ComputeStereoFishEyeMatches, brute-force half (src/Frame.cc:613-663) for ONE frame of n_cams cameras whose
// keypoints are already lapping-ordered by ORBextractor::operator() (in-area ones at rows >= num_mono): for every camera
// pair i < j in the reference's loop order, knnMatch(desc_i[num_mono_i:], desc_j[num_mono_j:], k = 2) — the pair is
// skipped (:623) when either side has no in-area row — then `size >= 2 && (d0 < d1 * 0.7 || (d0 < 75 && d0 < d1 * 0.9))`
// with float distances against double constants (:659-663).  desc: n_cams blocks of `cap` rows.  Outputs [n_pairs][cap]:
// idx / dist [..][2] (-1 / INT_MAX = none), good (0 / 1).
extern "C" void orc_fisheye_matches(const uint8_t* desc, const int32_t* n_kp, const int32_t* n_mono, int n_cams, int cap,
                                    int32_t* idx, int32_t* dist, uint8_t* good) {
  const int thOrbDist = (100 + 50) / 2;  // (ORBmatcher::TH_HIGH + ORBmatcher::TH_LOW) / 2
  int pair = 0;
  for (int i = 0; i < n_cams - 1; ++i)
    for (int j = i + 1; j < n_cams; ++j, ++pair) {
      int32_t* pi = idx + (size_t)pair * cap * 2;
      int32_t* pd = dist + (size_t)pair * cap * 2;
      uint8_t* pg = good + (size_t)pair * cap;
      for (int r = 0; r < cap; ++r) {
        pi[2 * r] = pi[2 * r + 1] = -1;
        pd[2 * r] = pd[2 * r + 1] = INT_MAX;
        pg[r] = 0;
      }
      if (n_mono[i] >= n_kp[i] || n_mono[j] >= n_kp[j]) continue;
      const int nq = n_kp[i] - n_mono[i], nt = n_kp[j] - n_mono[j];
      orc_hamming_knn2(desc + ((size_t)i * cap + n_mono[i]) * 32, nq, desc + ((size_t)j * cap + n_mono[j]) * 32, nt, pi, pd);
      for (int r = 0; r < nq; ++r) {
        if (pi[2 * r + 1] < 0) continue;  // knnMatch returned fewer than 2 neighbours
        const float d0 = (float)pd[2 * r], d1 = (float)pd[2 * r + 1];
        pg[r] = (d0 < d1 * 0.7 || (d0 < thOrbDist && d0 < d1 * 0.9)) ? 1 : 0;
      }
    }
}
