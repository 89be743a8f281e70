// The C++ boundary on the GPU: the reference-facing classes of vieo_slam_b200/host/vieo_shims.hpp and the flatten /
// write-back templates of vieo_flatten.hpp driven with real data.  tests/test_gpu_shims.py writes the inputs as raw arrays
// into a directory, runs this binary, and byte-compares what it writes back with the results of the ctypes path.
//   host_shim_gpu_test <dir>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <cmath>
#include <map>
#include <set>
#include <string>

#include "../vieo_slam_b200/host/vieo_flatten.hpp"

using namespace VIEO_SLAM_B200;

template <class T>
std::vector<T> rd(const std::string& dir, const char* name) {
  std::ifstream f(dir + "/" + name, std::ios::binary | std::ios::ate);
  if (!f) { std::fprintf(stderr, "missing %s\n", name); std::exit(2); }
  const size_t n = (size_t)f.tellg();
  std::vector<T> v(n / sizeof(T));
  f.seekg(0);
  f.read((char*)v.data(), (std::streamsize)(v.size() * sizeof(T)));
  return v;
}
template <class T>
void wr(const std::string& dir, const char* name, const T* p, size_t n) {
  std::ofstream f(dir + "/" + name, std::ios::binary);
  f.write((const char*)p, (std::streamsize)(n * sizeof(T)));
}

// ---- minimal stand-ins with the reference's member names (see vieo_flatten.hpp) ---------------------------------------------
struct NavState { VieoNavState s; };
inline VieoNavState vieo_from_navstate(const NavState& n) { return n.s; }
inline void vieo_to_navstate(const VieoNavState& v, NavState& n) { n.s = v; }
struct KeyFrame;
struct MapPoint {
  double X[3];
  int updated = 0;
  std::vector<std::pair<KeyFrame*, size_t>> obs_;
  const std::vector<std::pair<KeyFrame*, size_t>>& observations() const { return obs_; }
  void UpdateNormalAndDepth() { ++updated; }
  bool bad = false;
  bool isBad() const { return bad; }
};
inline void vieo_get_world_pos(const MapPoint& m, double o[3]) { o[0] = m.X[0]; o[1] = m.X[1]; o[2] = m.X[2]; }
inline void vieo_set_world_pos(MapPoint& m, const double* i) { m.X[0] = i[0]; m.X[1] = i[1]; m.X[2] = i[2]; }
struct StereoInfo { std::vector<float> vuright_, vdepth_; };
struct ScaleInfo { std::vector<float> vinvlevelsigma2_; };
struct Frame {
  int N = 0;
  std::vector<MapPoint*> mvpMapPoints;
  std::vector<cv::KeyPoint> mvKeysUn;
  StereoInfo stereoinfo_;
  ScaleInfo scalepyrinfo_;
  std::vector<bool> mvbOutlier;
  NavState ns;
  NavState GetNavState() const { return ns; }
  void SetNavState(const NavState& n) { ns = n; }
};

struct KeyFrame {
  std::vector<cv::KeyPoint> mvKeysUn;
  StereoInfo stereoinfo_;
  ScaleInfo scalepyrinfo_;
  NavState ns;
  KeyFrame* prev = nullptr;
  double ftimestamp_ = 0;
  int erased = 0;
  NavState GetNavState() const { return ns; }
  void SetNavState(const NavState& n) { ns = n; }
  KeyFrame* GetPrevKeyFrame() const { return prev; }
  std::vector<MapPoint*> mvpMapPoints;
  std::vector<MapPoint*> GetMapPointMatches() const { return mvpMapPoints; }
};
inline void ErasePairObs(KeyFrame* kf, MapPoint*) { ++kf->erased; }

// ---- stand-ins for the essential-graph templates (member names of include/KeyFrame.h / MapPoint.h / Map.h) -----------------------
struct PgKeyFrame {
  unsigned long nid_ = 0;
  bool bad = false;
  char state = 2;  // Tracking::OK
  PgKeyFrame* parent = nullptr;
  PgKeyFrame* prev = nullptr;
  std::set<PgKeyFrame*> children, loop_edges;
  std::map<PgKeyFrame*, int> weights;
  double Rcw[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, tcw[3] = {0, 0, 0};
  double sig_phi = 0, sig_p = 0;  // odometry sigmas of this keyframe's pre-integration (0: none)
  int pose_sets = 0;
  bool isBad() const { return bad; }
  char getState() const { return state; }
  PgKeyFrame* GetParent() const { return parent; }
  PgKeyFrame* GetPrevKeyFrame() const { return prev; }
  int GetWeight(PgKeyFrame* o) const { auto it = weights.find(o); return it == weights.end() ? 0 : it->second; }
  std::set<PgKeyFrame*> GetLoopEdges() const { return loop_edges; }
  bool hasChild(PgKeyFrame* o) const { return children.count(o) != 0; }
  std::vector<PgKeyFrame*> GetCovisiblesByWeight(int w) const {
    std::vector<PgKeyFrame*> v;
    for (auto& kv : weights) if (kv.second >= w) v.push_back(kv.first);
    return v;
  }
};
inline void vieo_get_Tcw(const PgKeyFrame& k, double R[9], double t[3]) { std::memcpy(R, k.Rcw, sizeof(k.Rcw)); std::memcpy(t, k.tcw, sizeof(k.tcw)); }
inline void vieo_set_Tcw(PgKeyFrame& k, const double T[12]) {
  for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) k.Rcw[3 * r + c] = T[4 * r + c]; k.tcw[r] = T[4 * r + 3]; }
  ++k.pose_sets;
}
inline bool vieo_odom_sigma(const PgKeyFrame& k, double& a, double& b) { a = k.sig_phi; b = k.sig_p; return k.sig_phi > 0; }
struct PgMapPoint {
  float X[3];
  bool bad = false;
  unsigned long mnCorrectedByKF = 0, mnCorrectedReference = 0;
  PgKeyFrame* ref = nullptr;
  int updated = 0;
  bool isBad() const { return bad; }
  PgKeyFrame* GetReferenceKeyFrame() const { return ref; }
  void UpdateNormalAndDepth() { ++updated; }
};
inline void vieo_get_world_pos_f(const PgMapPoint& m, float o[3]) { o[0] = m.X[0]; o[1] = m.X[1]; o[2] = m.X[2]; }
inline void vieo_set_world_pos_f(PgMapPoint& m, const float* i) { m.X[0] = i[0]; m.X[1] = i[1]; m.X[2] = i[2]; }
struct PgMap {
  std::vector<PgKeyFrame*> kfs;
  std::vector<PgMapPoint*> mps;
  std::vector<PgKeyFrame*> GetAllKeyFrames() const { return kfs; }
  std::vector<PgMapPoint*> GetAllMapPoints() const { return mps; }
  unsigned int GetMaxKFid() const { unsigned int m = 0; for (auto* k : kfs) m = std::max<unsigned int>(m, (unsigned int)k->nid_); return m; }
};

// A hand-made map for the essential-graph templates: 24 keyframes on a circle, drifting heading, keyframe 23 closes the loop on
// keyframe 2.  Keyframe 9 is bad, keyframe 14 was tracked by odometry only (pure-odometry spanning-tree edge with reduced
// information), (17, 5) is an earlier loop edge, covisibility >= 100 links i -> i - 2.
struct PgWorld {
  std::vector<PgKeyFrame> kf;
  std::vector<PgMapPoint> mp;
  PgMap map;
  std::map<PgKeyFrame*, VieoSim3> NonCorrected, Corrected;
  std::map<PgKeyFrame*, std::set<PgKeyFrame*>> LoopConnections;
  PgWorld() : kf(24), mp(40) {
    for (int i = 0; i < 24; ++i) {
      PgKeyFrame& k = kf[i];
      k.nid_ = i;
      const double a = 0.26 * i + 0.004 * i * i * 0.1, c = std::cos(a), s_ = std::sin(a);  // Rwc = Rz(a) with drift; Rcw = Rz(-a)
      const double Rcw[9] = {c, s_, 0, -s_, c, 0, 0, 0, 1};
      const double pw[3] = {5 * std::cos(0.26 * i) + 0.01 * i, 5 * std::sin(0.26 * i), 0.02 * i};
      std::memcpy(k.Rcw, Rcw, sizeof(Rcw));
      for (int r = 0; r < 3; ++r) k.tcw[r] = -(Rcw[3 * r] * pw[0] + Rcw[3 * r + 1] * pw[1] + Rcw[3 * r + 2] * pw[2]);
      if (i > 0) { k.parent = &kf[i - 1]; k.prev = &kf[i - 1]; kf[i - 1].children.insert(&k); k.weights[&kf[i - 1]] = 150; kf[i - 1].weights[&k] = 150; }
      if (i > 1) { k.weights[&kf[i - 2]] = 120; kf[i - 2].weights[&k] = 120; }
      if (i > 2) { k.weights[&kf[i - 3]] = 60; kf[i - 3].weights[&k] = 60; }
      map.kfs.push_back(&k);
    }
    kf[9].bad = true;
    kf[10].parent = &kf[8]; kf[9].children.erase(&kf[10]); kf[8].children.insert(&kf[10]);  // the spanning tree skips the bad keyframe
    kf[14].state = 1; kf[14].weights[&kf[13]] = 20; kf[13].weights[&kf[14]] = 20; kf[14].sig_phi = 0.02; kf[14].sig_p = 0.05;
    kf[17].loop_edges.insert(&kf[5]); kf[5].loop_edges.insert(&kf[17]);
    // the loop: keyframes 21..23 get corrected Sim3s (a small rotation about z and a shift towards keyframe 2's neighbourhood)
    for (int i = 21; i < 24; ++i) {
      NonCorrected[&kf[i]] = vieo_sim3_from_Rt(kf[i].Rcw, kf[i].tcw, 1.0);
      const double d = -0.05, c = std::cos(d), s_ = std::sin(d);
      const double dR[9] = {c, -s_, 0, s_, c, 0, 0, 0, 1}, dt[3] = {0.08, -0.05, 0.01};
      Corrected[&kf[i]] = vieo_sim3_mul(NonCorrected[&kf[i]], vieo_sim3_from_Rt(dR, dt, 1.0));
      for (int j = 1; j <= 3; ++j) { LoopConnections[&kf[i]].insert(&kf[j]); kf[i].weights[&kf[j]] = (i == 23 && j == 2) ? 40 : (j == 3 ? 80 : 130); }
    }
    for (int m = 0; m < 40; ++m) {
      mp[m].X[0] = 0.3f * m - 4.f; mp[m].X[1] = 0.11f * m; mp[m].X[2] = 1.f + 0.05f * m;
      mp[m].ref = &kf[(m * 7) % 24 == 9 ? 10 : (m * 7) % 24];
      if (m % 5 == 0) { mp[m].mnCorrectedByKF = 23; mp[m].mnCorrectedReference = 22; }
      if (m == 13) mp[m].bad = true;
      map.mps.push_back(&mp[m]);
    }
  }
};

static int essential_graph_templates_check() {
  PgWorld W;
  auto G = CollectEssentialGraph(&W.map, &W.kf[2], &W.kf[23], W.NonCorrected, W.Corrected, W.LoopConnections, (char)2);
  if (G.vScw.size() != 24 || G.kf_of[9] != nullptr || !G.fixed[2] || G.fixed[3]) return 1;
  std::set<std::pair<int, int>> e;
  for (size_t k = 0; k < G.edge_i.size(); ++k) {
    if (G.edge_i[k] == 9 || G.edge_j[k] == 9) return 2;  // nothing touches the bad keyframe
    e.insert({G.edge_i[k], G.edge_j[k]});
  }
  // loop connections: weight >= 100 or the (cur, loop) pair itself (:2409)
  if (!e.count({23, 2}) || !e.count({23, 1}) || e.count({23, 3}) || !e.count({21, 2}) || e.count({22, 3})) return 3;
  // spanning tree incl. the re-parented keyframe, earlier loop edge, covisibility i -> i - 2 but not the weak i - 3 link,
  // and no covisibility duplicate of a parent / child pair
  if (!e.count({10, 8}) || !e.count({17, 5}) || !e.count({12, 10}) || e.count({12, 9}) || e.count({8, 5}) || !e.count({1, 0})) return 4;
  size_t n_10_8 = 0;
  for (size_t k = 0; k < G.edge_i.size(); ++k) n_10_8 += G.edge_i[k] == 10 && G.edge_j[k] == 8;
  if (n_10_8 != 1) return 5;  // parent edge only: GetCovisiblesByWeight's (10, 8) is skipped (pKFn != pParentKF)
  // the pure-odometry edge (14 -> 13) carries diag(a I3, b I3, 1) with a = b = 1 here (the only such edge defines fOdomBase)
  if (!G.any_odom_info || G.info.size() != 49 * G.edge_i.size()) return 6;
  for (size_t k = 0; k < G.edge_i.size(); ++k) {
    const double* om = &G.info[49 * k];
    const bool odom = G.edge_i[k] == 14 && G.edge_j[k] == 13;
    if (odom && !(om[0] == 1.0 && om[24] == 1.0 && om[48] == 1.0)) return 7;
    if (!odom) for (int q = 0; q < 49; ++q) if (om[q] != (q % 8 == 0 ? 1.0 : 0.0)) return 8;
  }
  // a measurement is Sjw * Swi of the non-corrected poses for normal edges: (1, 0) from the keyframes' own poses
  for (size_t k = 0; k < G.edge_i.size(); ++k)
    if (G.edge_i[k] == 1 && G.edge_j[k] == 0) {
      const VieoSim3 want = vieo_sim3_mul(vieo_sim3_from_Rt(W.kf[0].Rcw, W.kf[0].tcw, 1.0), vieo_sim3_inv(vieo_sim3_from_Rt(W.kf[1].Rcw, W.kf[1].tcw, 1.0)));
      if (std::memcmp(&want, &G.Sji[k], sizeof(want))) return 9;
    }
  return 0;
}

// FlattenLocalWindow / WriteBackLocalWindow on a tiny hand-made window: ordering and bookkeeping only (no device call)
static int window_templates_check() {
  KeyFrame k[3];
  MapPoint m[2];
  for (int i = 0; i < 3; ++i) {
    k[i].mvKeysUn.resize(4);
    k[i].stereoinfo_.vuright_ = {-1.f, 10.f, -1.f, 3.f};
    k[i].stereoinfo_.vdepth_ = {-1.f, 4.f, -1.f, 50.f};
    k[i].scalepyrinfo_.vinvlevelsigma2_ = {1.f};
    k[i].ftimestamp_ = 0.5 * i;
    std::memset(&k[i].ns.s, 0, sizeof(VieoNavState));
    k[i].ns.s.q[0] = 1;
  }
  k[1].prev = &k[0]; k[2].prev = &k[1];
  m[0].X[0] = 1; m[0].X[1] = 2; m[0].X[2] = 3; m[0].obs_ = {{&k[0], 1}, {&k[2], 3}};
  m[1].X[0] = 4; m[1].X[1] = 5; m[1].X[2] = 6; m[1].obs_ = {{&k[1], 0}, {&k[2], 1}, {&k[0], 2}};
  std::vector<KeyFrame*> local = {&k[1], &k[2]}, fixed = {&k[0]};
  std::vector<MapPoint*> mps = {&m[0], &m[1]};
  auto W = FlattenLocalWindow<KeyFrame, MapPoint>(local, fixed, mps, 20.f, [](KeyFrame*) { return VieoImuPreint{}; });
  if (W.kfs.size() != 3 || W.mps.size() != 2 || W.edge_state.size() != 5 || W.imu_i.size() != 2) return 1;
  if (W.state_flags[0] != 2 || W.state_flags[2] != (1 | 2 | 4)) return 2;      // k[0] is fixed and carries V / Bias
  if (W.edge_point[0] != 0 || W.edge_point[2] != 1 || W.edge_state[0] != 2 || W.edge_state[1] != 1) return 3;
  if (!(W.edge_flags[0] & VIEO_EDGE_STEREO) || !(W.edge_flags[0] & VIEO_EDGE_CLOSE) || (W.edge_flags[1] & VIEO_EDGE_CLOSE)) return 4;
  const double gw[3] = {0, 0, -9.81};
  VieoBaProblem pb = W.problem(gw, 1.0, 1.0, false, false);
  if (pb.n_states != 3 || pb.n_edges != 5 || pb.n_imu != 2) return 5;
  VieoBaResult res{};
  res.accepted = 1;
  std::vector<VieoNavState> so(W.states);
  so[0].p[0] = 7;
  std::vector<double> po = {9, 9, 9, 8, 8, 8};
  std::vector<uint8_t> er = {0, 1, 0, 0, 0};
  WriteBackLocalWindow(W, 2, res, so, po, er);
  if (k[1].ns.s.p[0] != 7 || m[0].X[0] != 9 || m[1].X[2] != 8 || m[0].updated != 1 || k[2].erased != 1) return 6;
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  const std::string dir = argv[1];
  if (int rc = window_templates_check()) {
    std::fprintf(stderr, "window templates check failed: %d\n", rc);
    return 5;
  }
  if (int rc = essential_graph_templates_check()) {
    std::fprintf(stderr, "essential-graph templates check failed: %d\n", rc);
    return 5;
  }
  if (std::string(argv[1]) == "--templates-only") {
    std::printf("HOST_TEMPLATES_OK\n");
    return 0;
  }
  try {
    // 1. ORBextractor::operator() — keypoints, descriptors, monoIndex, public pyramid
    {
      auto meta = rd<int32_t>(dir, "orb_meta.i32");  // w, h, nfeatures, nlevels, lapping0, lapping1 (-1: none)
      auto img = rd<uint8_t>(dir, "orb_img.u8");
      ORBextractor ex(meta[2], 1.2f, meta[3], 20, 7);
      cv::Mat image(meta[1], meta[0], img.data());
      cv::Mat mask, desc;
      std::vector<cv::KeyPoint> kps;
      std::vector<int> lap = {meta[4], meta[5]};
      const int mono = ex(image, mask, kps, desc, meta[4] >= 0 ? &lap : nullptr);
      std::vector<float> flat;
      for (const cv::KeyPoint& k : kps) {
        flat.insert(flat.end(), {k.pt.x, k.pt.y, k.size, k.angle, k.response});
        float oct;
        std::memcpy(&oct, &k.octave, 4);
        flat.push_back(oct);
      }
      wr(dir, "orb_kps.out", flat.data(), flat.size());
      wr(dir, "orb_desc.out", desc.data, (size_t)desc.rows * 32);
      const int32_t info[3] = {mono, (int32_t)kps.size(), ex.GetLevels()};
      wr(dir, "orb_info.out", info, 3);
      wr(dir, "orb_level3.out", ex.mvImagePyramid[3].data, (size_t)ex.mvImagePyramid[3].rows * ex.mvImagePyramid[3].cols);
      cv::Mat empty;
      if (ex(empty, mask, kps, desc) != -1) return 3;
    }
    // 2. IMUPreIntegratorBase::PreIntegration
    {
      auto smp = rd<double>(dir, "imu_samples.f64");  // n x 7 (t, a, w)
      auto par = rd<double>(dir, "imu_par.f64");      // ti, tj, bg(3), ba(3), sigma2(4)
      IMUPreintegrator::SetParam(&par[8], 1, 200.0);
      std::list<IMUSample> data;
      for (size_t i = 0; i + 6 < smp.size(); i += 7) data.push_back({smp[i], {smp[i + 1], smp[i + 2], smp[i + 3]}, {smp[i + 4], smp[i + 5], smp[i + 6]}});
      IMUPreintegrator pre;
      const int st = pre.PreIntegration(par[0], par[1], &par[2], &par[5], data.begin(), data.end());
      std::vector<double> o;
      o.insert(o.end(), pre.mRij, pre.mRij + 9); o.insert(o.end(), pre.mvij, pre.mvij + 3); o.insert(o.end(), pre.mpij, pre.mpij + 3);
      o.insert(o.end(), pre.mSigmaijPRV, pre.mSigmaijPRV + 81); o.insert(o.end(), pre.mSigmaij, pre.mSigmaij + 81);
      o.insert(o.end(), pre.mJgpij, pre.mJgpij + 9); o.insert(o.end(), pre.mJapij, pre.mJapij + 9);
      o.insert(o.end(), pre.mJgvij, pre.mJgvij + 9); o.insert(o.end(), pre.mJavij, pre.mJavij + 9);
      o.insert(o.end(), pre.mJgRij, pre.mJgRij + 9);
      o.push_back(pre.mdeltatij); o.push_back((double)st);
      wr(dir, "imu.out", o.data(), o.size());
    }
    // 3. Optimizer::PoseOptimization through the flatten template on a stand-in Frame
    {
      auto cam = rd<VieoCamera>(dir, "po_cam.bin");
      auto ns = rd<VieoNavState>(dir, "po_state.bin");
      auto X = rd<double>(dir, "po_Xw.f64");
      auto ob = rd<float>(dir, "po_obs.f32");
      auto w = rd<float>(dir, "po_w.f32");
      auto fl = rd<uint8_t>(dir, "po_flags.u8");
      const int E = (int)fl.size();
      Frame F;
      F.N = E + 5;  // a few keypoints without a map point
      std::vector<MapPoint> mps(E);
      F.mvpMapPoints.assign(F.N, nullptr); F.mvKeysUn.resize(F.N); F.stereoinfo_.vuright_.assign(F.N, -1.f);
      F.mvbOutlier.assign(F.N, true);
      F.scalepyrinfo_.vinvlevelsigma2_ = {1.f, 0.f};
      // invSigma2 values are arbitrary floats in the dump: one table entry per edge keeps them exact
      F.scalepyrinfo_.vinvlevelsigma2_.assign(w.begin(), w.end());
      for (int e = 0; e < E; ++e) {
        const int i = e + (e >= 3 ? 2 : 0) + (e >= 40 ? 3 : 0);  // holes at 3, 4 and three more from 42 on
        mps[e].X[0] = X[3 * e]; mps[e].X[1] = X[3 * e + 1]; mps[e].X[2] = X[3 * e + 2];
        F.mvpMapPoints[i] = &mps[e];
        F.mvKeysUn[i].pt.x = ob[3 * e]; F.mvKeysUn[i].pt.y = ob[3 * e + 1]; F.mvKeysUn[i].octave = e;
        F.stereoinfo_.vuright_[i] = (fl[e] & VIEO_EDGE_STEREO) ? ob[3 * e + 2] : -1.f;
      }
      F.ns.s = ns[0];
      const int inl = PoseOptimizationVisual(&F, cam[0]);
      std::vector<uint8_t> outl;
      for (int i = 0; i < F.N; ++i)
        if (F.mvpMapPoints[i]) outl.push_back(F.mvbOutlier[i] ? 1 : 0);
      wr(dir, "po_outlier.out", outl.data(), outl.size());
      wr(dir, "po_state.out", &F.ns.s, 1);
      const int32_t r[1] = {inl};
      wr(dir, "po_inliers.out", r, 1);
    }
    // 4. LocalBA::Run and the asynchronous Begin / End on the dumped window
    {
      auto cam = rd<VieoCamera>(dir, "ba_cam.bin");
      auto st = rd<VieoNavState>(dir, "ba_states.bin");
      auto sf = rd<uint8_t>(dir, "ba_state_flags.u8");
      auto pts = rd<double>(dir, "ba_points.f64");
      auto es = rd<int32_t>(dir, "ba_edge_state.i32");
      auto ep = rd<int32_t>(dir, "ba_edge_point.i32");
      auto ob = rd<float>(dir, "ba_obs.f32");
      auto w = rd<float>(dir, "ba_w.f32");
      auto ef = rd<uint8_t>(dir, "ba_edge_flags.u8");
      auto ii = rd<int32_t>(dir, "ba_imu_i.i32");
      auto ij = rd<int32_t>(dir, "ba_imu_j.i32");
      auto pre = rd<VieoImuPreint>(dir, "ba_preint.bin");
      auto dt = rd<double>(dir, "ba_dt.f64");
      auto par = rd<double>(dir, "ba_par.f64");  // gw(3), inv_sigma_bg2, inv_sigma_ba2
      VieoBaProblem pb{};
      pb.n_states = (int)st.size(); pb.n_points = (int)pts.size() / 3; pb.n_edges = (int)es.size(); pb.n_imu = (int)ii.size();
      pb.states = st.data(); pb.state_flags = sf.data(); pb.points = pts.data(); pb.edge_state = es.data(); pb.edge_point = ep.data();
      pb.obs = ob.data(); pb.inv_sigma2 = w.data(); pb.edge_flags = ef.data(); pb.imu_i = ii.data(); pb.imu_j = ij.data();
      pb.preint = pre.data(); pb.imu_dt_kf = dt.data();
      pb.gw[0] = par[0]; pb.gw[1] = par[1]; pb.gw[2] = par[2]; pb.inv_sigma_bg2 = par[3]; pb.inv_sigma_ba2 = par[4];
      LocalBA ba(64, 4096, 32768, 32);
      std::vector<VieoNavState> so(st.size());
      std::vector<double> po(pts.size()), chi(es.size());
      std::vector<uint8_t> er(es.size());
      VieoBaResult res{};
      bool stop = false;
      ba.Run(pb, cam[0], &stop, so.data(), po.data(), chi.data(), er.data(), res);
      wr(dir, "ba_states.out", so.data(), so.size());
      wr(dir, "ba_points.out", po.data(), po.size());
      wr(dir, "ba_erase.out", er.data(), er.size());
      wr(dir, "ba_res.out", &res, 1);
      std::vector<VieoNavState> so2(st.size());
      std::vector<double> po2(pts.size());
      std::vector<uint8_t> er2(es.size());
      VieoBaResult res2{};
      ba.Begin(pb, cam[0], &stop);
      ba.End(so2.data(), po2.data(), nullptr, er2.data(), res2);
      if (std::memcmp(so.data(), so2.data(), sizeof(VieoNavState) * so.size()) || std::memcmp(po.data(), po2.data(), 8 * po.size()) ||
          er != er2) return 4;
    }
    // 5. Optimizer::OptimizeSim3 through the collection / write-back template on stand-in keyframes.  Keyframe 1 holds
    //    map point i at keypoint i, the match is a second map point observed in keyframe 2 at keypoint i; both keyframes'
    //    camera poses are the identity, so the dumped camera-frame positions are the map points' world positions.  One
    //    matched map point is bad and one match is empty: neither may reach the optimiser.
    {
      auto cam = rd<VieoCamera>(dir, "s3_cam.bin");
      auto X1 = rd<double>(dir, "s3_X1.f64");
      auto X2 = rd<double>(dir, "s3_X2.f64");
      auto o1 = rd<float>(dir, "s3_obs1.f32");
      auto o2 = rd<float>(dir, "s3_obs2.f32");
      auto oc1 = rd<int32_t>(dir, "s3_oct1.i32");
      auto oc2 = rd<int32_t>(dir, "s3_oct2.i32");
      auto isig = rd<float>(dir, "s3_invsigma2.f32");
      auto par = rd<double>(dir, "s3_par.f64");  // q12(4), t12(3), s12, th2, bFixScale
      const int M = (int)oc1.size();
      KeyFrame k1, k2;
      std::vector<MapPoint> m1(M + 2), m2(M + 2);
      std::vector<MapPoint*> vpMatches1(M + 2, nullptr);
      k1.scalepyrinfo_.vinvlevelsigma2_ = isig; k2.scalepyrinfo_.vinvlevelsigma2_ = isig;
      k1.mvKeysUn.resize(M + 2); k2.mvKeysUn.resize(M + 2); k1.mvpMapPoints.resize(M + 2);
      for (int i = 0; i < M + 2; ++i) {
        const int j = std::min(i, M - 1);
        for (int c = 0; c < 3; ++c) { m1[i].X[c] = X1[3 * j + c]; m2[i].X[c] = X2[3 * j + c]; }
        k1.mvKeysUn[i].pt.x = o1[2 * j]; k1.mvKeysUn[i].pt.y = o1[2 * j + 1]; k1.mvKeysUn[i].octave = oc1[j];
        k2.mvKeysUn[i].pt.x = o2[2 * j]; k2.mvKeysUn[i].pt.y = o2[2 * j + 1]; k2.mvKeysUn[i].octave = oc2[j];
        k1.mvpMapPoints[i] = &m1[i];
        vpMatches1[i] = &m2[i];
      }
      m2[M].bad = true;             // a bad matched point: skipped (:2767)
      vpMatches1[M + 1] = nullptr;  // no match: skipped (:2754)
      const float I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, z3[3] = {0, 0, 0};
      double q12[4] = {par[0], par[1], par[2], par[3]}, t12[3] = {par[4], par[5], par[6]}, s12 = par[7];
      MapPoint* base = m2.data();
      const int nIn = OptimizeSim3<KeyFrame, MapPoint>(&k1, &k2, vpMatches1, I3, z3, I3, z3, q12, t12, s12, (float)par[8], par[9] != 0,
                                                       cam[0], [&](MapPoint* p) { return (int)(p - base); });
      std::vector<uint8_t> kept(M + 2);
      for (int i = 0; i < M + 2; ++i) kept[i] = vpMatches1[i] != nullptr;
      if (!kept[M]) return 6;  // the skipped bad match is not touched by the write-back
      const double o[9] = {(double)nIn, q12[0], q12[1], q12[2], q12[3], t12[0], t12[1], t12[2], s12};
      wr(dir, "s3.out", o, 9);
      wr(dir, "s3_keep.out", kept.data(), (size_t)M);
    }
    // 6. Optimizer::OptimizeEssentialGraph through the collection / write-back templates on the hand-made map: the flattened
    //    graph and the results are written out; the test feeds the same arrays to the ctypes path and compares the bytes.
    {
      PgWorld W;
      auto G = CollectEssentialGraph(&W.map, &W.kf[2], &W.kf[23], W.NonCorrected, W.Corrected, W.LoopConnections, (char)2);
      wr(dir, "pg_Scw.bin", G.vScw.data(), G.vScw.size());
      wr(dir, "pg_fixed.u8", G.fixed.data(), G.fixed.size());
      wr(dir, "pg_ei.i32", G.edge_i.data(), G.edge_i.size());
      wr(dir, "pg_ej.i32", G.edge_j.data(), G.edge_j.size());
      wr(dir, "pg_meas.bin", G.Sji.data(), G.Sji.size());
      wr(dir, "pg_info.f64", G.info.data(), G.info.size());
      std::vector<float> Pw;
      std::vector<int32_t> ref;
      for (auto& m : W.mp) {
        if (m.bad) continue;
        Pw.insert(Pw.end(), m.X, m.X + 3);
        ref.push_back(m.mnCorrectedByKF == 23 ? (int)m.mnCorrectedReference : (int)m.ref->nid_);
      }
      wr(dir, "pg_Pw.f32", Pw.data(), Pw.size());
      wr(dir, "pg_ref.i32", ref.data(), ref.size());
      VieoPoseGraphStats st{};
      const int its = OptimizeAndWriteBackEssentialGraph(&W.map, G, &W.kf[23], true, st);
      if (its < 1 || W.kf[9].pose_sets != 0 || W.kf[3].pose_sets != 1 || W.mp[13].updated != 0 || W.mp[12].updated != 1) return 7;
      wr(dir, "pg_Scw.out", G.vScw.data(), G.vScw.size());
      std::vector<double> T;
      for (auto& k : W.kf) { for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) T.push_back(k.Rcw[3 * r + c]); T.push_back(k.tcw[r]); } }
      wr(dir, "pg_Tcw.out", T.data(), T.size());
      std::vector<float> Po;
      for (auto& m : W.mp) if (!m.bad) Po.insert(Po.end(), m.X, m.X + 3);
      wr(dir, "pg_Pw.out", Po.data(), Po.size());
      wr(dir, "pg_stats.out", &st, 1);
    }
    std::printf("HOST_SHIM_GPU_OK %s\n", vieo_version());
  } catch (const std::exception& e) {
    std::fprintf(stderr, "exception: %s\n", e.what());
    return 1;
  }
  return 0;
}
