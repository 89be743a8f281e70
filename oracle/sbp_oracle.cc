// ORACLE — TEST INFRASTRUCTURE ONLY (see orb_oracle.cc header).  CPU restatement of the tracking thread's guided
// searches: ORBmatcher::SearchByProjection(Frame&, const Frame& last, ...) (src/ORBmatcher.cc:1303-1467) and
// ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>, ...) (:230-335) on top of FrameBase::AssignFeaturesToGrid /
// GetFeaturesInArea / PosInGrid / IsInImage (src/FrameBase.cpp:95-174).  Single pinhole camera (rectified stereo /
// monocular, usedistort_ == false).  The reference's tests hold no vectors for this path; pinned by the reference's own
// functions compiled unchanged (oracle/_ref: both overloads and the grid functions, tests/test_oracle_ref.py) and by brute-force
// restatements in tests/test_oracle_sbp.py.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include "oracle.h"

namespace {
constexpr int kCols = 64, kRows = 48;  // FRAME_GRID_COLS / FRAME_GRID_ROWS (include/FrameBase.h:224-225)
constexpr int TH_HIGH = 100, HISTO_LENGTH = 30;

struct Grid {
  std::vector<std::vector<int>> cells;  // [ix * kRows + iy]
  // FrameBase::AssignFeaturesToGrid + PosInGrid (src/FrameBase.cpp:143-170)
  Grid(const OrcSbpFrame& f, const OrcKeyPoint* kps) : cells(kCols * kRows) {
    for (int i = 0; i < f.n_kp; ++i) {
      const int px = (int)std::round((kps[i].x - f.minx) * f.grid_winv);
      const int py = (int)std::round((kps[i].y - f.miny) * f.grid_hinv);
      if (px < 0 || px >= kCols || py < 0 || py >= kRows) continue;
      cells[px * kRows + py].push_back(i);
    }
  }
};

// FrameBase::GetFeaturesInArea (src/FrameBase.cpp:95-142)
void features_in_area(const OrcSbpFrame& f, const Grid& g, const OrcKeyPoint* kps, float x, float y, float r, int minlevel,
                      int maxlevel, std::vector<int>& out) {
  out.clear();
  const int min_cellx = std::max(0, (int)std::floor((x - f.minx - r) * f.grid_winv));
  if (min_cellx >= kCols) return;
  const int max_cellx = std::min(kCols - 1, (int)std::ceil((x - f.minx + r) * f.grid_winv));
  if (max_cellx < 0) return;
  const int min_celly = std::max(0, (int)std::floor((y - f.miny - r) * f.grid_hinv));
  if (min_celly >= kRows) return;
  const int max_celly = std::min(kRows - 1, (int)std::ceil((y - f.miny + r) * f.grid_hinv));
  if (max_celly < 0) return;
  const bool bchecklevel = (minlevel > 0) || (maxlevel >= 0);
  for (int ix = min_cellx; ix <= max_cellx; ++ix)
    for (int iy = min_celly; iy <= max_celly; ++iy)
      for (int j : g.cells[ix * kRows + iy]) {
        const OrcKeyPoint& kp = kps[j];
        if (bchecklevel) {
          if (kp.octave < minlevel) continue;
          if (maxlevel >= 0)
            if (kp.octave > maxlevel) continue;
        }
        const float distx = kp.x - x, disty = kp.y - y;
        if (std::fabs(distx) < r && std::fabs(disty) < r) out.push_back(j);
      }
}

// Eigen::Quaternion::_transformVector as used by Sophus::SO3 * point: v + w * (2 q x v) + q x (2 q x v)
void qrot(const double q[4], const double v[3], double o[3]) {
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  double uv[3] = {y * v[2] - z * v[1], z * v[0] - x * v[2], x * v[1] - y * v[0]};
  for (int i = 0; i < 3; ++i) uv[i] += uv[i];
  const double c[3] = {y * uv[2] - z * uv[1], z * uv[0] - x * uv[2], x * uv[1] - y * uv[0]};
  for (int i = 0; i < 3; ++i) o[i] = v[i] + w * uv[i] + c[i];
}

int descriptor_distance(const uint8_t* a, const uint8_t* b) { return orc_descriptor_distance(a, b); }

// ORBmatcher::ComputeThreeMaxima (src/ORBmatcher.cc:1608-1641)
void three_maxima(const int* cnt, int L, int& ind1, int& ind2, int& ind3) {
  int max1 = 0, max2 = 0, max3 = 0;
  for (int i = 0; i < L; i++) {
    const int s = cnt[i];
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
  if ((float)max2 < 0.1f * (float)max1) {
    ind2 = -1; ind3 = -1;
  } else if ((float)max3 < 0.1f * (float)max1) {
    ind3 = -1;
  }
}
}  // namespace

// SearchByProjection(CurrentFrame, LastFrame, th, bMono, th_far_pts).  Queries = the last frame's keypoints that hold a
// map point and are not outliers, in keypoint order (the host filters, :1325-1327).  q_flags bit 0: the map point has
// Observations() > 0, so the keypoint it takes is skipped by later queries (:1400-1401).  kp_blocked (nullable): keypoints
// that already hold such a map point on entry.  Outputs: kp_match[n_kp] = query that owns the keypoint at the end (-1
// none / erased by the rotation check), q_match / q_dist = keypoint chosen by each query (-1 / 256 when none) before the
// rotation check.  Returns nmatches.
extern "C" int orc_sbp_last_frame(const OrcSbpFrame* f, const OrcKeyPoint* kps, const float* uright, const uint8_t* desc,
                                  const double* q_Xw, const int32_t* q_octave, const float* q_angle, const uint8_t* q_desc,
                                  const uint8_t* q_flags, const uint8_t* kp_blocked, int32_t* kp_match, int32_t* q_match,
                                  int32_t* q_dist) {
  int nmatches = 0;
  Grid grid(*f, kps);
  std::vector<uint8_t> blocked(f->n_kp, 0);
  for (int i = 0; i < f->n_kp; ++i) {
    kp_match[i] = -1;
    if (kp_blocked) blocked[i] = kp_blocked[i];
  }
  std::vector<std::vector<int>> rotHist(HISTO_LENGTH);
  const float factor = 1.0f / HISTO_LENGTH;
  // Tlrcr = Tlrw * Tcrw^-1: translation = Rl * (-(Rc^-1 tc)) + tl
  double tz;
  {
    const double qci[4] = {f->qcw[0], -f->qcw[1], -f->qcw[2], -f->qcw[3]};
    double a[3], b3[3];
    qrot(qci, f->tcw, a);
    for (int i = 0; i < 3; ++i) a[i] = -a[i];
    qrot(f->qlw, a, b3);
    tz = b3[2] + f->tlw[2];
  }
  const bool bForward = tz > (double)f->b && !f->mono;
  const bool bBackward = -tz > (double)f->b && !f->mono;
  std::vector<int> cand;
  for (int i = 0; i < f->n_q; ++i) {
    q_match[i] = -1;
    q_dist[i] = 256;
    double x3Dr[3];
    qrot(f->qcw, q_Xw + 3 * (size_t)i, x3Dr);
    for (int k = 0; k < 3; ++k) x3Dr[k] += f->tcw[k];
    if (f->th_far > 0 && x3Dr[2] > (double)f->th_far) continue;
    const float xc = (float)x3Dr[0], yc = (float)x3Dr[1];
    const float invzc = (float)(1.0 / x3Dr[2]);
    if (invzc < 0) continue;
    // K (float) * (xc invzc, yc invzc, 1): Eigen's 3-term dot evaluates (a0 b0 + a1 b1) + a2 b2
    const float xn = xc * invzc, yn = yc * invzc;
    const float u = (f->fx * xn + 0.f * yn) + f->cx * 1.f;
    const float v = (0.f * xn + f->fy * yn) + f->cy * 1.f;
    if (!(u >= f->minx && u < f->maxx && v >= f->miny && v < f->maxy)) continue;
    const int nLastOctave = q_octave[i];
    const float radius = f->th * f->scale[nLastOctave];
    if (bForward) features_in_area(*f, grid, kps, u, v, radius, 0, nLastOctave, cand);
    else if (bBackward) features_in_area(*f, grid, kps, u, v, radius, nLastOctave, -1, cand);
    else features_in_area(*f, grid, kps, u, v, radius, nLastOctave - 1, nLastOctave + 1, cand);
    if (cand.empty()) continue;
    int bestDist = 256, bestIdx2 = -1;
    for (int i2 : cand) {
      if (blocked[i2]) continue;
      if (uright[i2] > 0) {
        const float ur = u - f->bf * invzc;
        const float er = std::fabs(ur - uright[i2]);
        if (er > radius) continue;
      }
      const int dist = descriptor_distance(q_desc + 32 * (size_t)i, desc + 32 * (size_t)i2);
      if (dist < bestDist) {
        bestDist = dist;
        bestIdx2 = i2;
      }
    }
    if (bestDist <= TH_HIGH) {
      kp_match[bestIdx2] = i;  // AddMapPoint
      if (q_flags[i] & 1) blocked[bestIdx2] = 1;
      q_match[i] = bestIdx2;
      q_dist[i] = bestDist;
      nmatches++;
      if (f->check_orientation) {
        float rot = q_angle[i] - kps[bestIdx2].angle;
        if (rot < 0.0) rot += 360.0f;
        int bin = (int)std::round(rot * factor);
        if (bin == HISTO_LENGTH) bin = 0;
        rotHist[bin].push_back(bestIdx2);
      }
    }
  }
  if (f->check_orientation) {
    int ind1 = -1, ind2 = -1, ind3 = -1, cnt[HISTO_LENGTH];
    for (int i = 0; i < HISTO_LENGTH; ++i) cnt[i] = (int)rotHist[i].size();
    three_maxima(cnt, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; i++)
      if (i != ind1 && i != ind2 && i != ind3)
        for (int k : rotHist[i]) {
          kp_match[k] = -1;  // EraseMapPointMatch
          nmatches--;
        }
  }
  return nmatches;
}

// SearchByProjection(F, vpMapPoints, th, th_far_pts).  Queries = the local map points that are in view (btrack_inview_,
// not bad; the host filters, :244-248) with the tracking info Frame::isInFrustum left: q_proj = (u, v, ur),
// q_level = predicted scale level, q_viewcos, q_depth = track_depth_.
extern "C" int orc_sbp_local_map(const OrcSbpFrame* f, const OrcKeyPoint* kps, const float* uright, const uint8_t* desc,
                                 const float* q_proj, const int32_t* q_level, const float* q_viewcos, const float* q_depth,
                                 const uint8_t* q_desc, const uint8_t* q_flags, const uint8_t* kp_blocked, int32_t* kp_match,
                                 int32_t* q_match, int32_t* q_dist) {
  int nmatches = 0;
  Grid grid(*f, kps);
  std::vector<uint8_t> blocked(f->n_kp, 0);
  for (int i = 0; i < f->n_kp; ++i) {
    kp_match[i] = -1;
    if (kp_blocked) blocked[i] = kp_blocked[i];
  }
  const bool bFactor = f->th != 1.0;
  std::vector<int> cand;
  for (int i = 0; i < f->n_q; ++i) {
    q_match[i] = -1;
    q_dist[i] = 256;
    if (q_level[i] < 0) continue;  // !btrack_inview_ (:244): the level isInFrustum leaves for a point out of view is -1
    if (f->th_far > 0 && q_depth[i] > f->th_far) continue;
    const int nPredictedLevel = q_level[i];
    float r = q_viewcos[i] > 0.998 ? 2.5f : 4.0f;  // RadiusByViewingCos (:337-342), float vs double literal compare
    if (bFactor) r *= f->th;
    const float rs = r * f->scale[nPredictedLevel];
    features_in_area(*f, grid, kps, q_proj[3 * i], q_proj[3 * i + 1], rs, nPredictedLevel - 1, nPredictedLevel, cand);
    if (cand.empty()) continue;
    int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
    for (int idx : cand) {
      if (blocked[idx]) continue;
      if (uright[idx] > 0) {
        const float er = std::fabs(q_proj[3 * i + 2] - uright[idx]);
        if (er > rs) continue;
      }
      const int dist = descriptor_distance(q_desc + 32 * (size_t)i, desc + 32 * (size_t)idx);
      if (dist < bestDist) {
        bestDist2 = bestDist;
        bestDist = dist;
        bestLevel2 = bestLevel;
        bestLevel = kps[idx].octave;
        bestIdx = idx;
      } else if (dist < bestDist2) {
        bestLevel2 = kps[idx].octave;
        bestDist2 = dist;
      }
    }
    if (bestDist <= TH_HIGH) {
      if (bestLevel == bestLevel2 && (float)bestDist > f->nn_ratio * (float)bestDist2) continue;
      kp_match[bestIdx] = i;
      if (q_flags[i] & 1) blocked[bestIdx] = 1;
      q_match[i] = bestIdx;
      q_dist[i] = bestDist;
      nmatches++;
    }
  }
  return nmatches;
}

// ---------------------------------------------------------------------------------------------------------------------
// Frame::isInFrustum (src/Frame.cc:335-416) + MapPoint::PredictScale (src/MapPoint.cc:491-509) for the single-camera
// case (mpCameras.size() == 1: GetTcr() is the identity and GetTrc().translation() is zero, so Pc = Pcr and twc = mOw
// exactly), usedistort_ == false.  Everything is float arithmetic.  Eigen 3.3's fixed-size products / dot / norm reduce a
// 3-term sum with redux_novec_unroller, i.e. a0 + (a1 + a2); this file is compiled without FP contraction.  "parity
// unpinned" for the last bit: the reference's own result depends on whether its compiler contracts to FMA.
//   in : Rcw / tcw / Ow = Tcw_ rotation, mtcw, mOw cast to float; per point wP, normal Pn, mfMaxDistance, mfMinDistance
//   out: inview (btrack_inview_), proj = (u, v, ur), level = vtrack_scalelevel_, viewcos, depth = track_depth_
// Returns the number of points in view (what SearchLocalPoints counts in nToMatch, src/Tracking.cc:2341-2344).
namespace {
inline float sum3(float a0, float a1, float a2) { return a0 + (a1 + a2); }
}  // namespace
extern "C" int orc_predict_scale(float max_distance, float current_dist, float log_scale_factor, int n_levels) {
  const float ratio = max_distance / current_dist;
  // log(float) resolves to the float overload; ceil(float); the int conversion of the reference is UB for inf / NaN,
  // defined here as "largest level" (ratio = +inf) resp. 0 (NaN)
  const float x = std::ceil(std::log(ratio) / log_scale_factor);
  if (std::isnan(x)) return 0;
  if (x < 0) return 0;
  if (x >= (float)n_levels) return n_levels - 1;
  return (int)x;
}
extern "C" int orc_is_in_frustum(const OrcFrustumFrame* f, int n, const float* wP, const float* Pn, const float* max_dist,
                                 const float* min_dist, uint8_t* inview, float* proj, int32_t* level, float* viewcos,
                                 float* depth) {
  int n_in = 0;
  for (int i = 0; i < n; ++i) {
    inview[i] = 0;
    level[i] = -1;
    proj[3 * i] = proj[3 * i + 1] = proj[3 * i + 2] = 0;
    viewcos[i] = 0;
    depth[i] = 0;
    const float* P = wP + 3 * i;
    const float maxDistance = 1.2f * max_dist[i], minDistance = 0.8f * min_dist[i];  // MapPoint.cc:481-489
    float Pc[3];
    for (int r = 0; r < 3; ++r)
      Pc[r] = sum3(f->Rcw[3 * r] * P[0], f->Rcw[3 * r + 1] * P[1], f->Rcw[3 * r + 2] * P[2]) + f->tcw[r];
    const float PcZ = Pc[2];
    if (PcZ < 0.0f) continue;
    const float invz = 1.0f / PcZ;
    const float xn = Pc[0] * invz, yn = Pc[1] * invz;
    // K * (xn, yn, 1): row 0 = fx*xn + (0*yn + cx*1), row 1 = 0*xn + (fy*yn + cy*1)
    const float u = sum3(f->fx * xn, 0.0f * yn, f->cx * 1.0f);
    const float v = sum3(0.0f * xn, f->fy * yn, f->cy * 1.0f);
    if (u < f->minx || u > f->maxx) continue;
    if (v < f->miny || v > f->maxy) continue;
    const float PO[3] = {P[0] - f->Ow[0], P[1] - f->Ow[1], P[2] - f->Ow[2]};
    const float dist3D = std::sqrt(sum3(PO[0] * PO[0], PO[1] * PO[1], PO[2] * PO[2]));
    if (dist3D < minDistance || dist3D > maxDistance) continue;
    const float vc = sum3(PO[0] * Pn[3 * i], PO[1] * Pn[3 * i + 1], PO[2] * Pn[3 * i + 2]) / dist3D;
    if (vc < f->cos_limit) continue;
    level[i] = orc_predict_scale(max_dist[i], dist3D, f->log_scale_factor, f->n_levels);
    proj[3 * i] = u;
    proj[3 * i + 1] = v;
    proj[3 * i + 2] = u - f->bf * invz;
    viewcos[i] = vc;
    depth[i] = dist3D;  // sum_depth / 1 camera
    inview[i] = 1;
    ++n_in;
  }
  return n_in;
}

// Frame::isInFrustum for a camera rig (src/Frame.cc:351-411 with mpCameras.size() > 1): the single-camera restatement above
// with the per-camera steps the reference adds.  Sophus::SE3f * Vector3f = unit_quaternion()._transformVector(p) +
// translation() (common/so3_extra.h:102-104; Eigen: uv = q.vec x p; uv += uv; p + w uv + q.vec x uv), all float.  With
// usedistort_ the pixel comes from the camera's Project in double, rounded to float (camera_pinhole.h:70-83,
// camera_kb8.h:68-106: atan2, the k4..k1 Horner chain as written, then the base class' projection of (mx, my, 1)).
namespace {
inline void cross3f(const float a[3], const float b[3], float o[3]) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
inline void project_double(const OrcFrustumCam& c, const float Pc[3], float& u, float& v) {
  const double x = (double)Pc[0], y = (double)Pc[1];
  if (c.model == 2) {
    const double x2 = x * x, y2 = y * y, r2 = x2 + y2, r = std::sqrt(r2);
    const float precision_r = 1e-5f;
    if (r > precision_r) {
      const double z = (double)Pc[2];
      const double theta = std::atan2(r, z), theta2 = theta * theta;
      double thetad = c.k[3] * theta2;
      thetad += c.k[2];
      thetad *= theta2;
      thetad += c.k[1];
      thetad *= theta2;
      thetad += c.k[0];
      thetad *= theta2;
      thetad += 1;
      thetad *= theta;
      const double mx = x * thetad / r, my = y * thetad / r;
      const double invz = 1. / 1.;
      u = (float)((double)c.fx * mx * invz + c.cx);
      v = (float)((double)c.fy * my * invz + c.cy);
      return;
    }
  }
  const double z = (double)Pc[2], invz = 1. / z;
  u = (float)((double)c.fx * x * invz + c.cx);
  v = (float)((double)c.fy * y * invz + c.cy);
}
}  // namespace
extern "C" int orc_is_in_frustum_rig(const OrcFrustumRigFrame* f, int n, const float* wP, const float* Pn, const float* max_dist,
                                     const float* min_dist, uint8_t* inview, uint8_t* cam_mask, float* proj, int32_t* level,
                                     float* viewcos, float* depth) {
  int n_in = 0;
  for (int i = 0; i < n; ++i) {
    inview[i] = 0;
    cam_mask[i] = 0;
    depth[i] = 0;
    for (int c = 0; c < 4; ++c) {
      level[4 * i + c] = -1;
      viewcos[4 * i + c] = 0;
      proj[12 * i + 3 * c] = proj[12 * i + 3 * c + 1] = proj[12 * i + 3 * c + 2] = 0;
    }
    const float* P = wP + 3 * i;
    const float maxDistance = 1.2f * max_dist[i], minDistance = 0.8f * min_dist[i];
    float Pcr[3];
    for (int r = 0; r < 3; ++r)
      Pcr[r] = sum3(f->Rcw[3 * r] * P[0], f->Rcw[3 * r + 1] * P[1], f->Rcw[3 * r + 2] * P[2]) + f->tcw[r];
    float sum_depth = 0;
    int count = 0;
    for (int ci = 0; ci < f->n_cams; ++ci) {
      const OrcFrustumCam& c = f->cam[ci];
      float uv[3], cr[3], Pc[3];
      cross3f(c.q_cr, Pcr, uv);
      for (int k = 0; k < 3; ++k) uv[k] += uv[k];
      cross3f(c.q_cr, uv, cr);
      for (int k = 0; k < 3; ++k) Pc[k] = ((Pcr[k] + c.q_cr[3] * uv[k]) + cr[k]) + c.t_cr[k];
      float twc[3];
      for (int k = 0; k < 3; ++k)  // Rcrw^T t: column k of Rcrw dotted with t
        twc[k] = f->Ow[k] + sum3(f->Rcw[k] * c.t_rc[0], f->Rcw[3 + k] * c.t_rc[1], f->Rcw[6 + k] * c.t_rc[2]);
      const float PcZ = Pc[2];
      if (PcZ < 0.0f) continue;
      const float invz = 1.0f / PcZ;
      float u, v;
      if (c.model == 0) {
        const float xn = Pc[0] * invz, yn = Pc[1] * invz;
        u = sum3(c.fx * xn, 0.0f * yn, c.cx * 1.0f);
        v = sum3(0.0f * xn, c.fy * yn, c.cy * 1.0f);
      } else {
        project_double(c, Pc, u, v);
      }
      if (u < c.minx || u > c.maxx) continue;
      if (v < c.miny || v > c.maxy) continue;
      const float PO[3] = {P[0] - twc[0], P[1] - twc[1], P[2] - twc[2]};
      const float dist3D = std::sqrt(sum3(PO[0] * PO[0], PO[1] * PO[1], PO[2] * PO[2]));
      if (dist3D < minDistance || dist3D > maxDistance) continue;
      const float vc = sum3(PO[0] * Pn[3 * i], PO[1] * Pn[3 * i + 1], PO[2] * Pn[3 * i + 2]) / dist3D;
      if (vc < f->cos_limit) continue;
      level[4 * i + ci] = orc_predict_scale(max_dist[i], dist3D, f->log_scale_factor, f->n_levels);
      proj[12 * i + 3 * ci] = u;
      proj[12 * i + 3 * ci + 1] = v;
      proj[12 * i + 3 * ci + 2] = u - f->bf * invz;
      viewcos[4 * i + ci] = vc;
      cam_mask[i] |= (uint8_t)(1u << ci);
      sum_depth += dist3D;
      ++count;
    }
    if (count) {
      depth[i] = sum_depth / (float)count;
      inview[i] = 1;
      ++n_in;
    }
  }
  return n_in;
}

// ---------------------------------------------------------------------------------------------------------------------
// ORBmatcher::SearchByProjectionBase (src/ORBmatcher.cc:26-227) — the search half shared by Fuse(KF, vpMapPoints, th)
// (:1152-1165, LocalMapping::SearchInNeighbors), Fuse(KF, Scw, ...) (:1167-1220, loop closing) and the Sim3 / relocalisation
// variants: per map point the projection into the keyframe, the image / scale-invariance / viewing-cone tests,
// PredictScale, GetFeaturesInArea(u, v, th_radius * scale[level]), level band [level - 1, level], the chi-square gate
// against the keypoint (stereo 7.8 / mono 5.99, only with pbf), and the strict-'<' Hamming arg-min.  Single camera,
// usedistort_ == false.  What the reference then does with (bestIdx, bestDist) — threshold, FuseMP / AddObservation /
// vpReplacePoint, the IsInKeyFrame skip — only touches map-point links, never keypoints or descriptors, so the search
// result of a point does not depend on the points before it; that policy stays on the host.  q_skip != 0: the point is
// null / bad / already in the keyframe (pvbAlreadyMatched1).  best_idx = -1, best_dist = 256: nothing found.
extern "C" void orc_sbp_base(const OrcProjSearchFrame* f, const OrcKeyPoint* kps, const float* uright, const uint8_t* desc,
                             const float* wP, const float* Pn, const float* max_dist, const float* min_dist,
                             const uint8_t* q_desc, const uint8_t* q_skip, int32_t* best_idx, int32_t* best_dist,
                             int32_t* level) {
  OrcSbpFrame g{};  // the grid helpers only read the grid fields
  g.n_kp = f->n_kp;
  g.minx = f->minx; g.maxx = f->maxx; g.miny = f->miny; g.maxy = f->maxy;
  g.grid_winv = f->grid_winv; g.grid_hinv = f->grid_hinv;
  Grid grid(g, kps);
  std::vector<int> cand;
  for (int i = 0; i < f->n_q; ++i) {
    best_idx[i] = -1;
    best_dist[i] = 256;
    level[i] = -1;
    if (q_skip && q_skip[i]) continue;
    const float* P = wP + 3 * i;
    float Pc[3];
    for (int r = 0; r < 3; ++r)
      Pc[r] = sum3(f->Rcw[3 * r] * P[0], f->Rcw[3 * r + 1] * P[1], f->Rcw[3 * r + 2] * P[2]) + f->tcw[r];
    if (Pc[2] <= 0.0) continue;
    const float invz = 1 / Pc[2];
    const float xn = Pc[0] * invz, yn = Pc[1] * invz;
    const float u = sum3(f->fx * xn, 0.0f * yn, f->cx * 1.0f);
    const float v = sum3(0.0f * xn, f->fy * yn, f->cy * 1.0f);
    if (!(u >= f->minx && u < f->maxx && v >= f->miny && v < f->maxy)) continue;  // IsInImage (FrameBase.cpp:171-174)
    const float PO[3] = {P[0] - f->Ow[0], P[1] - f->Ow[1], P[2] - f->Ow[2]};
    const float dist3D = std::sqrt(sum3(PO[0] * PO[0], PO[1] * PO[1], PO[2] * PO[2]));
    const float maxDistance = 1.2f * max_dist[i], minDistance = 0.8f * min_dist[i];
    if (dist3D < minDistance || dist3D > maxDistance) continue;
    if (f->check_viewing_angle) {
      const float dot = sum3(PO[0] * Pn[3 * i], PO[1] * Pn[3 * i + 1], PO[2] * Pn[3 * i + 2]);
      if ((double)dot < 0.5 * (double)dist3D) continue;
    }
    const int nPredictedLevel = orc_predict_scale(max_dist[i], dist3D, f->log_scale_factor, f->n_levels);
    level[i] = nPredictedLevel;
    const float radius = f->th_radius * f->scale[nPredictedLevel];
    features_in_area(g, grid, kps, u, v, radius, -1, -1, cand);
    if (cand.empty()) continue;
    int bestDist = INT32_MAX, bestIdx = -1;
    for (int idx : cand) {
      const OrcKeyPoint& kp = kps[idx];
      const int kpLevel = kp.octave;
      if (kpLevel < nPredictedLevel - 1 || kpLevel > nPredictedLevel) continue;
      if (f->use_bf) {
        const float kpr = uright[idx];
        if (kpr >= 0) {
          const float ex = u - kp.x, ey = v - kp.y;
          const float ur = u - f->bf * invz;
          const float er = ur - kpr;
          const float e2 = ex * ex + ey * ey + er * er;
          if ((double)(e2 * f->inv_level_sigma2[kpLevel]) > 7.8) continue;
        } else {
          const float ex = u - kp.x, ey = v - kp.y;
          const float e2 = ex * ex + ey * ey;
          if ((double)(e2 * f->inv_level_sigma2[kpLevel]) > 5.99) continue;
        }
      }
      const int dist = descriptor_distance(q_desc + 32 * (size_t)i, desc + 32 * (size_t)idx);
      if (dist < bestDist) {
        bestDist = dist;
        bestIdx = idx;
      }
    }
    if (bestIdx >= 0) {
      best_idx[i] = bestIdx;
      best_dist[i] = bestDist;
    }
  }
}


// ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const set<MapPoint*>& sAlreadyFound, th, ORBdist,
// th_far_pts) (src/ORBmatcher.cc:1471-1606) — Tracking::Relocalization's guided search after the PnP pose.  Single camera,
// usedistort_ == false.  Differences to the last-frame search restated above: no "invzc < 0" skip, the level comes from
// MapPoint::PredictScale of the camera-centre distance (with the 0.8 / 1.2 invariance gate), the band is always
// [level - 1, level + 1], no stereo ur gate, EVERY accepted match claims its keypoint (AddMapPoint makes
// CurrentFrame.GetMapPointMatches()[i2] non-null, :1553-1554), the acceptance threshold is the caller's ORBdist.
// q_level (nullable): the predicted level of each query (-1: failed a geometric test), for the tests.
extern "C" int orc_sbp_reloc(const OrcSbpFrame* f, int orb_dist, float log_scale_factor, const OrcKeyPoint* kps,
                             const uint8_t* desc, const double* q_Xw, const float* q_angle, const float* q_max_dist,
                             const float* q_min_dist, const uint8_t* q_desc, const uint8_t* kp_blocked, int32_t* kp_match,
                             int32_t* q_match, int32_t* q_dist, int32_t* q_level) {
  int nmatches = 0;
  Grid grid(*f, kps);
  std::vector<uint8_t> blocked(f->n_kp, 0);
  for (int i = 0; i < f->n_kp; ++i) {
    kp_match[i] = -1;
    if (kp_blocked) blocked[i] = kp_blocked[i];
  }
  std::vector<std::vector<int>> rotHist(HISTO_LENGTH);
  const float factor = 1.0f / HISTO_LENGTH;
  // Twcr = Tcrw.inverse(): translation = Rc^-1 * (tc * -1) (Sophus SE3::inverse)
  double twc[3];
  {
    const double qci[4] = {f->qcw[0], -f->qcw[1], -f->qcw[2], -f->qcw[3]};
    const double nt[3] = {f->tcw[0] * -1.0, f->tcw[1] * -1.0, f->tcw[2] * -1.0};
    qrot(qci, nt, twc);
  }
  std::vector<int> cand;
  for (int i = 0; i < f->n_q; ++i) {
    q_match[i] = -1;
    q_dist[i] = 256;
    if (q_level) q_level[i] = -1;
    const double* Xw = q_Xw + 3 * (size_t)i;
    double x3Dcr[3];
    qrot(f->qcw, Xw, x3Dcr);
    for (int k = 0; k < 3; ++k) x3Dcr[k] += f->tcw[k];
    if (f->th_far > 0 && x3Dcr[2] > (double)f->th_far) continue;
    const float invzc = (float)(1.0 / x3Dcr[2]);
    const float xc = (float)x3Dcr[0], yc = (float)x3Dcr[1];
    const float xn = xc * invzc, yn = yc * invzc;
    const float u = (f->fx * xn + 0.f * yn) + f->cx * 1.f;
    const float v = (0.f * xn + f->fy * yn) + f->cy * 1.f;
    if (!(u >= f->minx && u < f->maxx && v >= f->miny && v < f->maxy)) continue;
    const double PO[3] = {Xw[0] - twc[0], Xw[1] - twc[1], Xw[2] - twc[2]};
    const float dist3D = (float)std::sqrt(PO[0] * PO[0] + PO[1] * PO[1] + PO[2] * PO[2]);
    const float maxDistance = 1.2f * q_max_dist[i], minDistance = 0.8f * q_min_dist[i];
    if (dist3D < minDistance || dist3D > maxDistance) continue;
    const int nPredictedLevel = orc_predict_scale(q_max_dist[i], dist3D, log_scale_factor, f->n_levels);
    if (q_level) q_level[i] = nPredictedLevel;
    const float radius = f->th * f->scale[nPredictedLevel];
    features_in_area(*f, grid, kps, u, v, radius, nPredictedLevel - 1, nPredictedLevel + 1, cand);
    if (cand.empty()) continue;
    int bestDist = 256, bestIdx2 = -1;
    for (int i2 : cand) {
      if (blocked[i2]) continue;
      const int dist = descriptor_distance(q_desc + 32 * (size_t)i, desc + 32 * (size_t)i2);
      if (dist < bestDist) {
        bestDist = dist;
        bestIdx2 = i2;
      }
    }
    if (bestDist <= orb_dist) {
      kp_match[bestIdx2] = i;
      blocked[bestIdx2] = 1;
      q_match[i] = bestIdx2;
      q_dist[i] = bestDist;
      nmatches++;
      if (f->check_orientation) {
        float rot = q_angle[i] - kps[bestIdx2].angle;
        if (rot < 0.0) rot += 360.0f;
        int bin = (int)std::round(rot * factor);
        if (bin == HISTO_LENGTH) bin = 0;
        rotHist[bin].push_back(bestIdx2);
      }
    }
  }
  if (f->check_orientation) {
    int ind1 = -1, ind2 = -1, ind3 = -1, cnt[HISTO_LENGTH];
    for (int i = 0; i < HISTO_LENGTH; ++i) cnt[i] = (int)rotHist[i].size();
    three_maxima(cnt, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; i++)
      if (i != ind1 && i != ind2 && i != ind3)
        for (int k : rotHist[i]) {
          kp_match[k] = -1;
          nmatches--;
        }
  }
  return nmatches;
}


// Test hook: the grid alone — AssignFeaturesToGrid over one frame's keypoints, then GetFeaturesInArea / IsInImage per query
// (x, y, r, minlevel, maxlevel); the candidate lists in walk order.  Pinned against the reference's own four functions compiled
// unchanged (oracle/_ref, tests/test_oracle_ref.py).
extern "C" int orc_features_in_area(const OrcKeyPoint* kps, int n_kp, float minx, float maxx, float miny, float maxy, float winv,
                                    float hinv, const float* q_xyr, const int32_t* q_levels, int n_q, int32_t* out_ptr,
                                    int32_t* out_idx, int cap, uint8_t* in_image) {
  OrcSbpFrame f;
  memset(&f, 0, sizeof(f));
  f.n_kp = n_kp;
  f.minx = minx; f.maxx = maxx; f.miny = miny; f.maxy = maxy;
  f.grid_winv = winv; f.grid_hinv = hinv;
  const Grid grid(f, kps);
  std::vector<int> cand;
  int total = 0;
  out_ptr[0] = 0;
  for (int q = 0; q < n_q; ++q) {
    const float x = q_xyr[3 * q], y = q_xyr[3 * q + 1];
    features_in_area(f, grid, kps, x, y, q_xyr[3 * q + 2], q_levels[2 * q], q_levels[2 * q + 1], cand);
    for (int j : cand) {
      if (total < cap) out_idx[total] = j;
      ++total;
    }
    out_ptr[q + 1] = total;
    if (in_image) in_image[q] = (x >= f.minx && x < f.maxx && y >= f.miny && y < f.maxy) ? 1 : 0;
  }
  return total;
}
