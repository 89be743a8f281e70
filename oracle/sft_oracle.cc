// ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of ORBmatcher::SearchForTriangulation
// (/root/reference/src/ORBmatcher.cc:896-1150) for the single-pinhole case (usedistort_ == false, one camera per
// keyframe: vn_cams = {1, 1}, mapn2in_ empty, USE_STRATEGY_MIN_DIST as common/config.h:12 defines it), and of the
// FeatureVector walk + best-match loop of ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...) (:344-505).
// Each step cites the lines it follows.  Pinned by the reference itself: both functions are compiled UNCHANGED in oracle/_ref
// (ref_sft_wrap.cc: SearchForTriangulation with GeometricCamera::epipolarConstrain / FillMatchesFromPair; ref_sbp_wrap.cc:
// SearchByBoW) and tests/test_oracle_ref.py asserts the same match lists in the same order; tests/test_oracle_sft.py adds a
// deliberately naive python restatement and known answers.
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <utility>
#include <vector>

#include "oracle.h"

namespace {
const int TH_LOW = 50, TH_HIGH = 100, HISTO_LENGTH = 30;  // src/ORBmatcher.cc:20-22

// ORBmatcher::ComputeThreeMaxima (:1608-1641)
void three_maxima(const std::vector<int>* histo, int L, int& ind1, int& ind2, int& ind3) {
  int max1 = 0, max2 = 0, max3 = 0;
  for (int i = 0; i < L; i++) {
    const int s = (int)histo[i].size();
    if (s > max1) {
      max3 = max2; max2 = max1; max1 = s;
      ind3 = ind2; ind2 = ind1; ind1 = i;
    } else if (s > max2) {
      max3 = max2; max2 = s;
      ind3 = ind2; ind2 = i;
    } else if (s > max3) {
      max3 = s; ind3 = i;
    }
  }
  if (max2 < 0.1f * (float)max1) {
    ind2 = -1; ind3 = -1;
  } else if (max3 < 0.1f * (float)max1) {
    ind3 = -1;
  }
}

// GeometricCamera::epipolarConstrain, the #else (fundamental-matrix) branch (common/camera_models/camera_base.h:360-404)
// with bkp_distort == false.  F12 = K1^-T t12^ R12 K2^-1 is formed by the caller in double (Eigen inverse(): host side).
// Tdata = float (common/config.h:23), Tcalc = double (:24): a, b, c, num, den are rounded to float where the reference
// declares them `const Tdata`.
bool epipolar_ok(const double F12[9], float x1, float y1, float x2, float y2, float unc) {
  const double p1x = (double)x1, p1y = (double)y1, p2x = (double)x2, p2y = (double)y2;
  const float a = (float)(p1x * F12[0] + p1y * F12[3] + F12[6]);
  const float b = (float)(p1x * F12[1] + p1y * F12[4] + F12[7]);
  const float c = (float)(p1x * F12[2] + p1y * F12[5] + F12[8]);
  const float num = (float)((double)a * p2x + (double)b * p2y + (double)c);
  const float den = a * a + b * b;
  if (den == 0) return false;
  const float dsqr = num * num / den;
  return dsqr < 3.84f * unc;
}
}  // namespace

extern "C" {

// fv1 / fv2: DBoW2::FeatureVector flattened — node ids ascending (std::map order), ptr (n + 1), keypoint indices in push
// order.  has_mp: pKF->GetMapPoint(idx) != nullptr.  kps: mvKeysUn (== mvKeys for the angle: single pinhole).
// Outputs: pairs [cap][2] = (idx1, idx2) of vMatchedPairs in creation order (good matches only); returns nmatches.
int orc_search_for_triangulation(const OrcKeyPoint* kp1, const float* ur1, const uint8_t* desc1, const uint8_t* has_mp1,
                                 const int32_t* fv1_node, const int32_t* fv1_ptr, const int32_t* fv1_idx, int n_nodes1,
                                 const OrcKeyPoint* kp2, const float* ur2, const uint8_t* desc2, const uint8_t* has_mp2,
                                 const int32_t* fv2_node, const int32_t* fv2_ptr, const int32_t* fv2_idx, int n_nodes2,
                                 const double F12[9], float ex, float ey, const float* scale_factor2,
                                 const float* level_sigma2_2, int only_stereo, int check_orientation, int32_t* pairs, int cap) {
  std::vector<int> rotHist[HISTO_LENGTH];
  const float factor = 1.0f / HISTO_LENGTH;
  int nmatches = 0;
  // vidxs_matches / goodmatches / mapcamidx2idxs of FillMatchesFromPair (camera_base.h:408-560) for cameras {0} and {1}
  std::vector<std::pair<int, int>> vidxs;  // (idx1, idx2)
  std::vector<bool> good;
  std::map<std::pair<int, int>, int> cam2idx;  // (cam, idx) -> entry
  int a = 0, b2 = 0;
  while (a < n_nodes1 && b2 < n_nodes2) {  // :964-1135
    if (fv1_node[a] == fv2_node[b2]) {
      for (int i1 = fv1_ptr[a]; i1 < fv1_ptr[a + 1]; ++i1) {
        const int idx1 = fv1_idx[i1];
        if (has_mp1[idx1]) continue;  // :972-973
        const bool bStereo1 = ur1[idx1] >= 0;
        if (only_stereo && !bStereo1) continue;
        const OrcKeyPoint& k1 = kp1[idx1];
        int bestDist = TH_LOW, bestIdx2 = -1;
        for (int i2 = fv2_ptr[b2]; i2 < fv2_ptr[b2 + 1]; ++i2) {
          const int idx2 = fv2_idx[i2];
          if (has_mp2[idx2]) continue;  // :996-997
          // "avoid multi2one match": idx2 already in a match that has a camera-0 member (:1001-1002)
          auto it = cam2idx.find({1, idx2});
          if (it != cam2idx.end() && vidxs[it->second].first != -1) continue;
          const bool bStereo2 = ur2[idx2] >= 0;
          if (only_stereo && !bStereo2) continue;
          const int dist = orc_descriptor_distance(desc1 + 32 * (size_t)idx1, desc2 + 32 * (size_t)idx2);
          if (dist > bestDist) continue;  // :1026-1027
          const OrcKeyPoint& k2 = kp2[idx2];
          if (!bStereo1 && !bStereo2) {  // :1031-1035
            const float distex = ex - k2.x, distey = ey - k2.y;
            if (distex * distex + distey * distey < 100 * scale_factor2[k2.octave]) continue;
          }
          if (epipolar_ok(F12, k1.x, k1.y, k2.x, k2.y, level_sigma2_2[k2.octave])) {  // :1047-1051
            bestIdx2 = idx2;
            bestDist = dist;
          }
        }
        if (bestIdx2 >= 0) {  // :1056-1093: FillMatchesFromPair with {(0, idx1), (1, idx2)}
          // (0, idx1) is new (a keypoint lies in one node) and (1, idx2) was filtered above: both absent -> a new entry
          // (checkdepth = {1, 1}; psigmas == nullptr -> no triangulation test) and `true`
          const int id = (int)vidxs.size();
          cam2idx.emplace(std::make_pair(0, idx1), id);
          cam2idx.emplace(std::make_pair(1, bestIdx2), id);
          vidxs.emplace_back(idx1, bestIdx2);
          good.push_back(true);
          ++nmatches;
          if (check_orientation) {
            float rot = k1.angle - kp2[bestIdx2].angle;
            if (rot < 0.0) rot += 360.0f;
            int bin = (int)std::round(rot * factor);
            if (bin == HISTO_LENGTH) bin = 0;
            rotHist[bin].push_back(idx1);
          }
        }
      }
      ++a;
      ++b2;
    } else if (fv1_node[a] < fv2_node[b2]) {
      a = (int)(std::lower_bound(fv1_node, fv1_node + n_nodes1, fv2_node[b2]) - fv1_node);
    } else {
      b2 = (int)(std::lower_bound(fv2_node, fv2_node + n_nodes2, fv1_node[a]) - fv2_node);
    }
  }
  if (check_orientation) {  // :1137-1156
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; i++) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (int idx1 : rotHist[i]) {
        auto it = cam2idx.find({0, idx1});
        if (it == cam2idx.end()) continue;
        good[it->second] = false;
        nmatches--;
      }
    }
  }
  int n = 0;
  for (size_t i = 0; i < vidxs.size(); ++i) {  // :1158-1183 (count_num == 2 for every entry here)
    if (!good[i]) continue;
    if (n < cap) {
      pairs[2 * n] = vidxs[i].first;
      pairs[2 * n + 1] = vidxs[i].second;
    }
    ++n;
  }
  return nmatches;
}

// ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vector<MapPoint*>& vpMapPointMatches) (src/ORBmatcher.cc:344-505),
// single camera (F.mapn2in_ empty: img_id == 0).  mp_id [n_kf]: the keyframe keypoint's map point as a dense id, -1 for
// none / isBad().  A map point may sit on several keyframe keypoints only in multi-camera rigs; the (pMP, img_id) table of
// :363,414-429 is restated anyway.  Output: match_f [n_f] = keyframe keypoint index whose map point the frame keypoint
// received (vpMapPointMatches), -1 none; returns nmatches.
int orc_search_by_bow(const OrcKeyPoint* kp_kf, const uint8_t* desc_kf, const int32_t* mp_id, const int32_t* fv1_node,
                      const int32_t* fv1_ptr, const int32_t* fv1_idx, int n_nodes1, const OrcKeyPoint* kp_f,
                      const uint8_t* desc_f, int n_f, const int32_t* fv2_node, const int32_t* fv2_ptr, const int32_t* fv2_idx,
                      int n_nodes2, float nn_ratio, int check_orientation, int32_t* match_f) {
  for (int i = 0; i < n_f; ++i) match_f[i] = -1;
  int nmatches = 0;
  std::vector<int> rotHist[HISTO_LENGTH];
  std::vector<size_t> rothist2erase[HISTO_LENGTH];
  const float factor = 1.0f / HISTO_LENGTH;
  struct Rec { int dist; int idx_f; int bin; size_t pos; };
  std::map<int, Rec> mp2rec;  // mapmpcami2distkpidhist with img_id == 0
  int a = 0, b2 = 0;
  while (a < n_nodes1 && b2 < n_nodes2) {
    if (fv1_node[a] == fv2_node[b2]) {
      for (int iKF = fv1_ptr[a]; iKF < fv1_ptr[a + 1]; ++iKF) {
        const int realIdxKF = fv1_idx[iKF];
        const int mp = mp_id[realIdxKF];
        if (mp < 0) continue;  // !pMP || pMP->isBad()
        int bestDist1 = 256, bestDist2 = 256, bestIdxF = -1;
        for (int iF = fv2_ptr[b2]; iF < fv2_ptr[b2 + 1]; ++iF) {
          const int realIdxF = fv2_idx[iF];
          if (match_f[realIdxF] >= 0) continue;  // :385
          const int dist = orc_descriptor_distance(desc_kf + 32 * (size_t)realIdxKF, desc_f + 32 * (size_t)realIdxF);
          if (dist < bestDist1) {
            bestDist2 = bestDist1;
            bestDist1 = dist;
            bestIdxF = realIdxF;
          } else if (dist < bestDist2) {
            bestDist2 = dist;
          }
        }
        if (bestDist1 <= TH_LOW && (float)bestDist1 < nn_ratio * (float)bestDist2) {  // :408-411
          auto it = mp2rec.find(mp);
          if (it != mp2rec.end()) {
            if (it->second.dist <= bestDist1) continue;
            match_f[it->second.idx_f] = -1;
            --nmatches;
            if (check_orientation) rothist2erase[it->second.bin].push_back(it->second.pos);
          }
          match_f[bestIdxF] = realIdxKF;
          Rec r{bestDist1, bestIdxF, -1, 0};
          if (check_orientation) {
            float rot = kp_kf[realIdxKF].angle - kp_f[bestIdxF].angle;
            if (rot < 0.0) rot += 360.0f;
            int bin = (int)std::round(rot * factor);
            if (bin == HISTO_LENGTH) bin = 0;
            r.bin = bin;
            r.pos = rotHist[bin].size();
            rotHist[bin].push_back(bestIdxF);
          }
          mp2rec.emplace(mp, r);  // emplace does NOT overwrite an existing key (:446): the first record stays
          nmatches++;
        }
      }
      ++a;
      ++b2;
    } else if (fv1_node[a] < fv2_node[b2]) {
      a = (int)(std::lower_bound(fv1_node, fv1_node + n_nodes1, fv2_node[b2]) - fv1_node);
    } else {
      b2 = (int)(std::lower_bound(fv2_node, fv2_node + n_nodes2, fv1_node[a]) - fv2_node);
    }
  }
  if (check_orientation) {  // :463-489
    std::vector<int> rotHist2[HISTO_LENGTH];
    for (int i = 0; i < HISTO_LENGTH; ++i) {
      for (size_t j : rothist2erase[i]) rotHist[i][j] = -1;
      for (int v : rotHist[i])
        if (v != -1) rotHist2[i].push_back(v);
    }
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(rotHist2, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; i++) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (int v : rotHist2[i]) {
        match_f[v] = -1;
        nmatches--;
      }
    }
  }
  return nmatches;
}

}  // extern "C"
