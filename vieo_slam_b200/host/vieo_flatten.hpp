// Graph collection and write-back for the optimiser entry points, as code: what the reference's Optimizer methods do
// between "Set ... vertices / edges" and "Recover optimized ..." (src/Optimizer.cc:1611-1874 for the visual
// PoseOptimization, :133-520 + :704-768 for LocalBundleAdjustmentNavStatePRV), written against the REFERENCE's member names
// so that the bodies drop into src/Optimizer.cc unchanged.  Templates: the reference's Frame / KeyFrame / MapPoint types
// need OpenCV, Eigen and Sophus, which this repository cannot include; tools/host_shim_gpu_test.cc instantiates them with
// minimal stand-ins that carry the same member names, and the GPU test compares the results with the ctypes path.
//
// Customisation points the integrator supplies for its own types (two-line functions over NavState / Eigen vectors):
//   VieoNavState vieo_from_navstate(const NavStateT&);   void vieo_to_navstate(const VieoNavState&, NavStateT&);
//   void vieo_get_world_pos(const MapPointT&, double out[3]);   void vieo_set_world_pos(MapPointT&, const double in[3]);
#pragma once
#include <cmath>
#include <vector>

#include <map>
#include <set>

#include "../csrc/sim3.cuh"  // g2o::Sim3 product / inverse as plain host functions (the header is host-compilable)
#include "vieo_shims.hpp"

namespace VIEO_SLAM_B200 {

// ---- Optimizer::PoseOptimization(Frame* pFrame, Frame* pLastF) (visual, src/Optimizer.cc:1611-1874) -----------------------
// Collection (:1660-1760): every pFrame->mvpMapPoints[i] != nullptr gives one edge; mono when vuright_[i] < 0, else stereo;
// information = vinvlevelsigma2_[kpUn.octave]; pFrame->mvbOutlier[i] = false.  Write-back (:1850-1872): mvbOutlier from the
// last classification, pose from the optimised vertex.  Returns nInitialCorrespondences - nBad.
template <class FrameT>
int PoseOptimizationVisual(FrameT* pFrame, const VieoCamera& cam, int device = 0) {
  std::vector<double> Xw;
  std::vector<float> obs, w;
  std::vector<uint8_t> flags;
  std::vector<int> edge_kp;
  const int N = pFrame->N;
  for (int i = 0; i < N; i++) {
    auto* pMP = pFrame->mvpMapPoints[i];
    if (!pMP) continue;
    pFrame->mvbOutlier[i] = false;
    const auto& kpUn = pFrame->mvKeysUn[i];
    const float ur = pFrame->stereoinfo_.vuright_[i];
    double X[3];
    vieo_get_world_pos(*pMP, X);
    Xw.insert(Xw.end(), X, X + 3);
    obs.push_back(kpUn.pt.x); obs.push_back(kpUn.pt.y); obs.push_back(ur < 0 ? 0.f : ur);
    w.push_back(pFrame->scalepyrinfo_.vinvlevelsigma2_[kpUn.octave]);
    flags.push_back(ur < 0 ? 0 : VIEO_EDGE_STEREO);
    edge_kp.push_back(i);
  }
  if ((int)edge_kp.size() < 3) return 0;  // :1762-1763
  VieoPoseOptProblem pb{};
  pb.mode = 0;
  pb.cur = vieo_from_navstate(pFrame->GetNavState());
  VieoPoseOptResult res{};
  std::vector<uint8_t> outlier;
  const int inliers = Optimizer::PoseOptimization(pb, cam, Xw, obs, w, flags, res, outlier, device);
  for (size_t e = 0; e < edge_kp.size(); ++e) pFrame->mvbOutlier[edge_kp[e]] = outlier[e] != 0;
  auto ns = pFrame->GetNavState();
  vieo_to_navstate(res.cur, ns);
  pFrame->SetNavState(ns);  // also refreshes Tcw (UpdatePoseFromNS)
  return inliers;
}

// ---- Optimizer::LocalBundleAdjustmentNavStatePRV (src/Optimizer.cc:21-769): flattening of an already collected window --------
// lLocalKeyFrames (ascending id), lFixedCameras, lLocalMapPoints as the reference builds them (:52-131); observations()
// returns the (KeyFrame*, idx) pairs of a map point.  Produces the arrays of VieoBaProblem (keyframes local-first, edges
// sorted by point) and remembers which (KeyFrame*, MapPoint*) every edge was, for ErasePairObs in the write-back.
template <class KeyFrameT, class MapPointT>
struct LocalWindow {
  std::vector<KeyFrameT*> kfs;       // local first, then fixed
  std::vector<MapPointT*> mps;
  std::vector<VieoNavState> states;
  std::vector<uint8_t> state_flags;
  std::vector<double> points;
  std::vector<int32_t> edge_state, edge_point, imu_i, imu_j;
  std::vector<float> obs, inv_sigma2;
  std::vector<uint8_t> edge_flags;
  std::vector<VieoImuPreint> preint;
  std::vector<double> imu_dt_kf;

  VieoBaProblem problem(const double gw[3], double inv_sigma_bg2, double inv_sigma_ba2, bool bLarge, bool bRecInit) const {
    VieoBaProblem pb{};
    pb.n_states = (int)states.size(); pb.n_points = (int)mps.size(); pb.n_edges = (int)edge_state.size(); pb.n_imu = (int)imu_i.size();
    pb.states = states.data(); pb.state_flags = state_flags.data(); pb.points = points.data();
    pb.edge_state = edge_state.data(); pb.edge_point = edge_point.data(); pb.obs = obs.data(); pb.inv_sigma2 = inv_sigma2.data();
    pb.edge_flags = edge_flags.data(); pb.imu_i = imu_i.data(); pb.imu_j = imu_j.data(); pb.preint = preint.data();
    pb.imu_dt_kf = imu_dt_kf.data();
    pb.gw[0] = gw[0]; pb.gw[1] = gw[1]; pb.gw[2] = gw[2];
    pb.inv_sigma_bg2 = inv_sigma_bg2; pb.inv_sigma_ba2 = inv_sigma_ba2;
    pb.large = bLarge; pb.rec_init = bRecInit;
    return pb;
  }
};

template <class KeyFrameT, class MapPointT, class PreintOf>
LocalWindow<KeyFrameT, MapPointT> FlattenLocalWindow(const std::vector<KeyFrameT*>& lLocalKeyFrames,
                                                     const std::vector<KeyFrameT*>& lFixedCameras,
                                                     const std::vector<MapPointT*>& lLocalMapPoints, float th_dist_far,
                                                     PreintOf preint_of) {
  LocalWindow<KeyFrameT, MapPointT> W;
  auto index_of = [&](KeyFrameT* kf) {
    for (size_t k = 0; k < W.kfs.size(); ++k)
      if (W.kfs[k] == kf) return (int)k;
    return -1;
  };
  for (KeyFrameT* kf : lLocalKeyFrames) {  // PR + V + Bias vertices (:141-176)
    W.kfs.push_back(kf);
    W.states.push_back(vieo_from_navstate(kf->GetNavState()));
    W.state_flags.push_back(2);
  }
  for (KeyFrameT* kf : lFixedCameras) {  // fixed PR; the previous window's last keyframe also carries fixed V / Bias (:180-217)
    W.kfs.push_back(kf);
    W.states.push_back(vieo_from_navstate(kf->GetNavState()));
    W.state_flags.push_back(1);
  }
  for (KeyFrameT* kf1 : lLocalKeyFrames) {  // inertial + bias-walk edges to the previous keyframe (:219-330)
    KeyFrameT* kf0 = kf1->GetPrevKeyFrame();
    if (!kf0) continue;
    int i0 = index_of(kf0);
    if (i0 < 0) continue;
    if (W.state_flags[i0] == 1) W.state_flags[i0] = 1 | 2 | 4;  // fixed keyframe that takes part in an inertial edge
    W.imu_i.push_back(i0);
    W.imu_j.push_back(index_of(kf1));
    W.preint.push_back(preint_of(kf1));
    W.imu_dt_kf.push_back(kf1->ftimestamp_ - kf0->ftimestamp_);
  }
  const float chi_far = th_dist_far;
  for (MapPointT* mp : lLocalMapPoints) {  // point vertices and reprojection edges, point by point (:367-520)
    double X[3];
    vieo_get_world_pos(*mp, X);
    const int p = (int)W.mps.size();
    bool any = false;
    for (const auto& ob : mp->observations()) {
      KeyFrameT* kf = ob.first;
      const int s = index_of(kf);
      if (s < 0) continue;
      const size_t idx = ob.second;
      const auto& kpUn = kf->mvKeysUn[idx];
      const float ur = kf->stereoinfo_.vuright_[idx];
      W.edge_state.push_back(s);
      W.edge_point.push_back(p);
      W.obs.push_back(kpUn.pt.x); W.obs.push_back(kpUn.pt.y); W.obs.push_back(ur < 0 ? 0.f : ur);
      W.inv_sigma2.push_back(kf->scalepyrinfo_.vinvlevelsigma2_[kpUn.octave]);
      uint8_t fl = ur < 0 ? 0 : VIEO_EDGE_STEREO;
      if (!(kf->stereoinfo_.vdepth_[idx] > chi_far)) fl |= VIEO_EDGE_CLOSE;  // far-point guard (:393-396)
      W.edge_flags.push_back(fl);
      any = true;
    }
    if (!any) continue;
    W.mps.push_back(mp);
    W.points.insert(W.points.end(), X, X + 3);
  }
  return W;
}

// Write-back (:704-768): erase the outlier observations, SetNavState on the local keyframes, SetWorldPos +
// UpdateNormalAndDepth on the points.  Nothing is written when the result was rejected (res.accepted == 0, :663-666).
template <class KeyFrameT, class MapPointT>
void WriteBackLocalWindow(LocalWindow<KeyFrameT, MapPointT>& W, size_t n_local, const VieoBaResult& res,
                          const std::vector<VieoNavState>& states_out, const std::vector<double>& points_out,
                          const std::vector<uint8_t>& erase) {
  if (!res.accepted) return;
  for (size_t e = 0; e < erase.size(); ++e)
    if (erase[e]) ErasePairObs(W.kfs[W.edge_state[e]], W.mps[W.edge_point[e]]);
  for (size_t k = 0; k < n_local; ++k) {
    auto ns = W.kfs[k]->GetNavState();
    vieo_to_navstate(states_out[k], ns);
    W.kfs[k]->SetNavState(ns);
  }
  for (size_t p = 0; p < W.mps.size(); ++p) {
    vieo_set_world_pos(*W.mps[p], &points_out[3 * p]);
    W.mps[p]->UpdateNormalAndDepth();
  }
}

// ---- Optimizer::OptimizeSim3 (src/Optimizer.cc:2689-2920): collection and write-back ------------------------------------------
// Collection (:2753-2845): for every i with vpMatches1[i], pMP1 = pKF1->GetMapPointMatches()[i], pMP2 = vpMatches1[i], both
// good and pMP2 observed in pKF2 at i2: P3D1c = R1w * P3D1w + t1w and P3D2c = R2w * P3D2w + t2w in FLOAT (:2771-2783), the
// undistorted keypoints kpUn1 = pKF1->mvKeysUn[i], kpUn2 = pKF2->mvKeysUn[i2] and their level informations.  R1w / t1w /
// R2w / t2w: row-major float[9] / float[3] of pKF1->GetRotation() ... (:2711-2714).  q12 (w, x, y, z), t12, s12: g2oS12 in
// and out.  index_in_kf2(pMP2) returns i2 or -1.  Write-back (:2858-2868, 2894-2905): vpMatches1[i] = nullptr for dropped
// pairs; g2oS12 only when the call did not return 0 before the second stage.  Returns nIn.
template <class KeyFrameT, class MapPointT, class IndexInKF2>
int OptimizeSim3(KeyFrameT* pKF1, KeyFrameT* pKF2, std::vector<MapPointT*>& vpMatches1, const float R1w[9], const float t1w[3],
                 const float R2w[9], const float t2w[3], double q12[4], double t12[3], double& s12, float th2, bool bFixScale,
                 const VieoCamera& cam, IndexInKF2 index_in_kf2, int device = 0) {
  const int N = (int)vpMatches1.size();
  const auto vpMapPoints1 = pKF1->GetMapPointMatches();
  std::vector<double> Xc1, Xc2;
  std::vector<float> obs1, obs2, w1, w2;
  std::vector<int> vnIndexEdge;
  auto to_cam = [](const float R[9], const float t[3], const double Xw[3], std::vector<double>& out) {
    const float x = (float)Xw[0], y = (float)Xw[1], z = (float)Xw[2];
    for (int r = 0; r < 3; ++r) out.push_back((double)(R[3 * r] * x + R[3 * r + 1] * y + R[3 * r + 2] * z + t[r]));
  };
  for (int i = 0; i < N; i++) {
    if (!vpMatches1[i]) continue;
    MapPointT* pMP1 = vpMapPoints1[i];
    MapPointT* pMP2 = vpMatches1[i];
    if (!pMP1 || !pMP2) continue;
    const int i2 = index_in_kf2(pMP2);
    if (pMP1->isBad() || pMP2->isBad() || i2 < 0) continue;
    double X1[3], X2[3];
    vieo_get_world_pos(*pMP1, X1);
    vieo_get_world_pos(*pMP2, X2);
    to_cam(R1w, t1w, X1, Xc1);
    to_cam(R2w, t2w, X2, Xc2);
    const auto& kpUn1 = pKF1->mvKeysUn[i];
    const auto& kpUn2 = pKF2->mvKeysUn[i2];
    obs1.push_back(kpUn1.pt.x); obs1.push_back(kpUn1.pt.y);
    obs2.push_back(kpUn2.pt.x); obs2.push_back(kpUn2.pt.y);
    w1.push_back(pKF1->scalepyrinfo_.vinvlevelsigma2_[kpUn1.octave]);
    w2.push_back(pKF2->scalepyrinfo_.vinvlevelsigma2_[kpUn2.octave]);
    vnIndexEdge.push_back(i);
  }
  // the vertex: ns.mRwb = R12^-1, ns.mpwb = -(mRwb * t12) (:2717-2722)
  VieoSim3Problem pb{};
  const double w = q12[0], x = -q12[1], y = -q12[2], z = -q12[3];  // conjugate = inverse rotation
  pb.ns.q[0] = w; pb.ns.q[1] = x; pb.ns.q[2] = y; pb.ns.q[3] = z;
  const double R[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y), 2 * (x * y + w * z), 1 - 2 * (x * x + z * z),
                       2 * (y * z - w * x),     2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)};
  for (int r = 0; r < 3; ++r) pb.ns.p[r] = -(R[3 * r] * t12[0] + R[3 * r + 1] * t12[1] + R[3 * r + 2] * t12[2]);
  pb.scale = s12;
  pb.th2 = th2;
  pb.fix_scale = bFixScale ? 1 : 0;
  VieoSim3Result res{};
  std::vector<uint8_t> keep;
  const int nIn = Optimizer::OptimizeSim3(pb, cam, Xc1, Xc2, obs1, obs2, w1, w2, res, keep, device);
  for (size_t e = 0; e < vnIndexEdge.size(); ++e)
    if (!keep[e]) vpMatches1[vnIndexEdge[e]] = nullptr;
  if (res.n_corr - res.n_bad >= 10) {  // not the early "return 0" (:2878): "Recover optimized Sim3" (:2907-2917): R12 = mRwb^-1, t12 = -(R12 * mpwb)
    const double qw = res.ns.q[0], qx = -res.ns.q[1], qy = -res.ns.q[2], qz = -res.ns.q[3];
    q12[0] = qw; q12[1] = qx; q12[2] = qy; q12[3] = qz;
    const double Q[9] = {1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - qw * qz), 2 * (qx * qz + qw * qy), 2 * (qx * qy + qw * qz),
                         1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - qw * qx), 2 * (qx * qz - qw * qy), 2 * (qy * qz + qw * qx),
                         1 - 2 * (qx * qx + qy * qy)};
    for (int r = 0; r < 3; ++r) t12[r] = -(Q[3 * r] * res.ns.p[0] + Q[3 * r + 1] * res.ns.p[1] + Q[3 * r + 2] * res.ns.p[2]);
    s12 = res.scale;
  }
  return nIn;
}

// ---- Optimizer::OptimizeEssentialGraph (src/Optimizer.cc:2309-2688): collection and write-back -------------------------------
// Vertices are indexed by nid_ (0 .. GetMaxKFid()), as the reference's vScw / vCorrectedSwc are; a bad keyframe keeps its
// slot but gets no vertex (:2351) - here: no edge may name it, its slot is copied through.
// Customisation points: void vieo_get_Tcw(const KeyFrameT&, double Rcw[9], double tcw[3]) (GetRotation / GetTranslation),
// void vieo_set_Tcw(KeyFrameT&, const double Tcw[12]) (SetPose + UpdateNavStatePVRFromTcw, :2638-2641),
// bool vieo_odom_sigma(const KeyFrameT& kf, double& sigma_phi, double& sigma_p): the two norms of :2454-2463 / :2525-2538
// taken from kf's encoder / IMU pre-integrations (false: the keyframe has none), void vieo_get_world_pos_f(const MapPointT&,
// float[3]) / vieo_set_world_pos_f(MapPointT&, const float[3]).  Sim3From: the caller's g2o::Sim3 -> VieoSim3 is
// {r.coeffs(), t, s} member for member.
inline VieoSim3 vieo_sim3_from_Rt(const double R[9], const double t[3], double s) {  // Sim3(Matrix3d, Vector3d, double), sim3.h:59
  vieo::Sim3d o;
  vieo::s3_R2q(R, o.q);
  o.t[0] = t[0]; o.t[1] = t[1]; o.t[2] = t[2];
  o.s = s;
  VieoSim3 v;
  std::memcpy(&v, &o, sizeof(v));
  return v;
}
inline VieoSim3 vieo_sim3_mul(const VieoSim3& a, const VieoSim3& b) {
  vieo::Sim3d x, y;
  std::memcpy(&x, &a, sizeof(x));
  std::memcpy(&y, &b, sizeof(y));
  const vieo::Sim3d r = vieo::s3_mul(x, y);
  VieoSim3 v;
  std::memcpy(&v, &r, sizeof(v));
  return v;
}
inline VieoSim3 vieo_sim3_inv(const VieoSim3& a) {
  vieo::Sim3d x;
  std::memcpy(&x, &a, sizeof(x));
  const vieo::Sim3d r = vieo::s3_inv(x);
  VieoSim3 v;
  std::memcpy(&v, &r, sizeof(v));
  return v;
}

template <class KeyFrameT>
struct EssentialGraph {
  std::vector<KeyFrameT*> kf_of;  // vertex index (nid_) -> keyframe (nullptr: bad / absent)
  std::vector<VieoSim3> vScw;
  std::vector<uint8_t> fixed;
  std::vector<int32_t> edge_i, edge_j;
  std::vector<VieoSim3> Sji;
  std::vector<double> info;  // [E][49]; empty while every edge carries the identity
  bool any_odom_info = false;
  void add_edge(int i, int j, const VieoSim3& S, const double* om49) {
    edge_i.push_back(i);
    edge_j.push_back(j);
    Sji.push_back(S);
    for (int k = 0; k < 49; ++k) info.push_back(om49 ? om49[k] : (k % 8 == 0 ? 1.0 : 0.0));
    if (om49) any_odom_info = true;
  }
};

template <class MapT, class KeyFrameT, class Sim3Map, class ConnMap>
EssentialGraph<KeyFrameT> CollectEssentialGraph(MapT* pMap, KeyFrameT* pLoopKF, KeyFrameT* pCurKF, const Sim3Map& NonCorrectedSim3,
                                                const Sim3Map& CorrectedSim3, const ConnMap& LoopConnections, char tracking_ok_state) {
  EssentialGraph<KeyFrameT> G;
  const std::vector<KeyFrameT*> vpKFs = pMap->GetAllKeyFrames();
  const unsigned int nMaxKFid = pMap->GetMaxKFid();
  const int minFeat = 100;  // (:2345)
  G.kf_of.assign(nMaxKFid + 1, nullptr);
  G.fixed.assign(nMaxKFid + 1, 0);
  const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, z3[3] = {0, 0, 0};
  G.vScw.assign(nMaxKFid + 1, vieo_sim3_from_Rt(I3, z3, 1.0));
  for (KeyFrameT* pKF : vpKFs) {  // "Set KeyFrame vertices" (:2348-2385)
    if (pKF->isBad()) continue;
    const int nIDi = (int)pKF->nid_;
    auto it = CorrectedSim3.find(pKF);
    if (it != CorrectedSim3.end()) G.vScw[nIDi] = it->second;
    else {
      double Rcw[9], tcw[3];
      vieo_get_Tcw(*pKF, Rcw, tcw);
      G.vScw[nIDi] = vieo_sim3_from_Rt(Rcw, tcw, 1.0);
    }
    if (pKF == pLoopKF) G.fixed[nIDi] = 1;
    G.kf_of[nIDi] = pKF;
  }
  std::set<std::pair<long unsigned int, long unsigned int>> sInsertedEdges;
  for (const auto& mit : LoopConnections) {  // "Set Loop edges" (:2396-2430)
    KeyFrameT* pKF = mit.first;
    const long unsigned int nIDi = pKF->nid_;
    const VieoSim3 Swi = vieo_sim3_inv(G.vScw[nIDi]);
    for (KeyFrameT* pKFj : mit.second) {
      const long unsigned int nIDj = pKFj->nid_;
      if ((nIDi != pCurKF->nid_ || nIDj != pLoopKF->nid_) && pKF->GetWeight(pKFj) < minFeat) continue;
      if (!G.kf_of[nIDi] || !G.kf_of[nIDj]) continue;  // g2o refuses an edge to a missing vertex
      G.add_edge((int)nIDi, (int)nIDj, vieo_sim3_mul(G.vScw[nIDj], Swi), nullptr);
      sInsertedEdges.insert(std::make_pair(std::min(nIDi, nIDj), std::max(nIDi, nIDj)));
    }
  }
  auto pure_odom_pair = [&](KeyFrameT* pKF, KeyFrameT* pParentKF) {  // (:2449-2450, :2503-2504)
    return (pKF->getState() != tracking_ok_state && pKF->GetPrevKeyFrame() == pParentKF) ||
           (pParentKF->getState() != tracking_ok_state && pParentKF->GetPrevKeyFrame() == pKF);
  };
  float fOdomBase[2] = {1, 1};  // (:2444-2469): the smallest odometry sigmas over the pure-odometry spanning-tree edges
  for (KeyFrameT* pKF : vpKFs) {
    KeyFrameT* pParentKF = pKF->GetParent();
    if (pParentKF && pKF->GetWeight(pParentKF) < minFeat && pure_odom_pair(pKF, pParentKF)) {
      double sphi = 0, sp = 0;
      if (vieo_odom_sigma(pKF->getState() != tracking_ok_state ? *pKF : *pParentKF, sphi, sp)) {
        if (fOdomBase[0] > sphi) fOdomBase[0] = (float)sphi;
        if (fOdomBase[1] > sp) fOdomBase[1] = (float)sp;
      }
    }
  }
  auto nc = [&](KeyFrameT* kf) -> VieoSim3 {  // the pose before this loop correction
    auto it = NonCorrectedSim3.find(kf);
    return it != NonCorrectedSim3.end() ? it->second : G.vScw[kf->nid_];
  };
  for (KeyFrameT* pKF : vpKFs) {  // "Set normal edges" (:2472-2616)
    const int nIDi = (int)pKF->nid_;
    if (!G.kf_of[nIDi]) continue;
    const VieoSim3 Swi = vieo_sim3_inv(nc(pKF));
    KeyFrameT* pParentKF = pKF->GetParent();
    if (pParentKF && G.kf_of[pParentKF->nid_]) {  // spanning-tree edge (:2490-2549)
      const VieoSim3 Sji = vieo_sim3_mul(nc(pParentKF), Swi);
      double om[49];
      const double* use = nullptr;
      if (pKF->GetWeight(pParentKF) < minFeat && pure_odom_pair(pKF, pParentKF)) {
        double sphi = 0, sp = 0;
        if (vieo_odom_sigma(pKF->GetPrevKeyFrame() == pParentKF ? *pKF : *pParentKF, sphi, sp)) {
          for (int k = 0; k < 49; ++k) om[k] = k % 8 == 0 ? 1.0 : 0.0;
          const float EPS_MAX_INFO = 1e6f;
          float elemInfo = fOdomBase[0] / (float)sphi;
          if (!(!elemInfo || elemInfo > EPS_MAX_INFO)) om[0] = om[8] = om[16] = elemInfo;
          elemInfo = fOdomBase[1] / (float)sp;
          if (!(!elemInfo || elemInfo > EPS_MAX_INFO)) om[24] = om[32] = om[40] = elemInfo;
          use = om;
        }
      }
      G.add_edge(nIDi, (int)pParentKF->nid_, Sji, use);
    }
    const std::set<KeyFrameT*> sLoopEdges = pKF->GetLoopEdges();  // earlier loop edges (:2551-2577)
    for (KeyFrameT* pLKF : sLoopEdges)
      if (pLKF->nid_ < pKF->nid_ && G.kf_of[pLKF->nid_]) G.add_edge(nIDi, (int)pLKF->nid_, vieo_sim3_mul(nc(pLKF), Swi), nullptr);
    const std::vector<KeyFrameT*> vpConnectedKFs = pKF->GetCovisiblesByWeight(minFeat);  // covisibility edges (:2579-2615)
    for (KeyFrameT* pKFn : vpConnectedKFs) {
      if (pKFn && pKFn != pParentKF && !pKF->hasChild(pKFn) && !sLoopEdges.count(pKFn)) {
        if (!pKFn->isBad() && pKFn->nid_ < pKF->nid_) {
          if (sInsertedEdges.count(std::make_pair(std::min(pKF->nid_, pKFn->nid_), std::max(pKF->nid_, pKFn->nid_)))) continue;
          G.add_edge(nIDi, (int)pKFn->nid_, vieo_sim3_mul(nc(pKFn), Swi), nullptr);
        }
      }
    }
  }
  if (!G.any_odom_info) G.info.clear();
  return G;
}

// Optimise + write-back (:2618-2682): SetPose from [R | t / s], then every good map point through its reference keyframe's
// pose before / after; the caller holds pMap->mMutexMapUpdate around this call and calls pMap->InformNewBigChange() after.
template <class MapT, class KeyFrameT>
int OptimizeAndWriteBackEssentialGraph(MapT* pMap, EssentialGraph<KeyFrameT>& G, KeyFrameT* pCurKF, bool bFixScale,
                                       VieoPoseGraphStats& stats, int device = 0) {
  std::vector<VieoSim3> Scw_out;
  std::vector<double> Tcw;
  const int its = Optimizer::OptimizeEssentialGraph(G.vScw, G.fixed, bFixScale, G.edge_i, G.edge_j, G.Sji, G.info, Scw_out, Tcw, stats, device);
  for (size_t k = 0; k < G.kf_of.size(); ++k)
    if (G.kf_of[k]) vieo_set_Tcw(*G.kf_of[k], &Tcw[12 * k]);
  const auto vpMPs = pMap->GetAllMapPoints();
  std::vector<float> Pw, Pw_out;
  std::vector<int32_t> ref;
  std::vector<size_t> which;
  for (size_t i = 0; i < vpMPs.size(); ++i) {
    auto* pMP = vpMPs[i];
    if (pMP->isBad()) continue;
    const int nIDr = pMP->mnCorrectedByKF == pCurKF->nid_ ? (int)pMP->mnCorrectedReference : (int)pMP->GetReferenceKeyFrame()->nid_;
    float X[3];
    vieo_get_world_pos_f(*pMP, X);
    Pw.insert(Pw.end(), X, X + 3);
    ref.push_back(nIDr);
    which.push_back(i);
  }
  Pw_out.resize(Pw.size());
  vieo_check(vieo_essential_graph_correct_points((int)ref.size(), Pw.data(), ref.data(), (int)G.vScw.size(), G.vScw.data(), Scw_out.data(),
                                                 Pw_out.data(), device), "vieo_essential_graph_correct_points");
  for (size_t e = 0; e < which.size(); ++e) {
    vieo_set_world_pos_f(*vpMPs[which[e]], &Pw_out[3 * e]);
    vpMPs[which[e]]->UpdateNormalAndDepth();
  }
  G.vScw = Scw_out;
  return its;
}

}  // namespace VIEO_SLAM_B200
