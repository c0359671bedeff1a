// orb_b200_matcher.cpp — the ORBmatcher members of the hot path, for the reference's build.
//
// Compiled against the reference's own, UNMODIFIED include/ORBmatcher.h (all eleven methods stay declared there). A
// maintainer deletes the bodies of ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:2083-2103),
// ORBmatcher::SearchForInitialization (:573-717), ORBmatcher::SearchByProjection(Frame&, const Frame&, float, bool)
// (:1710-1860, the per-frame tracking search) and ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, float)
// (:72-169, the local-map search), both ORBmatcher::SearchByBoW (:247-420, :729-880) and ORBmatcher::SearchForTriangulation
// (:884-1100) from src/ORBmatcher.cc and adds this file; the four methods that write into live MapPoint / KeyFrame objects
// (SearchByProjection keyframe / loop variants, Fuse x2, SearchBySim3) keep their CPU bodies in ORBmatcher.cc, and so do
// RadiusByViewingCos (:171), CheckDistEpipolarLine (:205) and ComputeThreeMaxima (:2035). (This repository's test build of the reference does the same without touching the
// source: it weakens the two symbols in the compiled ORBmatcher.o, see INTEGRATION.md.)
#include <algorithm>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "ORBmatcher.h"   // the reference's header (include/ORBmatcher.h)

#include "../../include/orb_b200.h"

namespace ORB_SLAM2 {

namespace {
struct MatcherCache {   // one C-ABI matcher per host thread (ORBmatcher objects are created per call in the reference)
  orb_matcher* m = nullptr;
  int cap = 0;
  ~MatcherCache() { orb_matcher_destroy(m); }
  orb_matcher* get(int n) {
    if (!m || cap < n) {
      orb_matcher_destroy(m);
      m = nullptr;
      cap = std::max(n, 4096);
      if (orb_matcher_create(0, 1, cap, &m) != ORB_OK)
        throw std::runtime_error(std::string("ORBmatcher (liborb_b200): ") + orb_last_error());
    }
    return m;
  }
};
thread_local MatcherCache t_matcher;

// DBoW2::FeatureVector (node id -> feature indices) inverted: the node id of every feature, -1 = none
std::vector<int32_t> node_of_features(const DBoW2::FeatureVector& fv, size_t n) {
  std::vector<int32_t> node(n, -1);
  for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it)
    for (size_t k = 0; k < it->second.size(); k++)
      if (it->second[k] < n) node[it->second[k]] = (int32_t)it->first;
  return node;
}

std::vector<unsigned char> rows_of(const cv::Mat& descriptors, size_t n) {
  std::vector<unsigned char> d(n * 32);
  for (size_t i = 0; i < n; i++) std::memcpy(&d[i * 32], descriptors.ptr((int)i), 32);
  return d;
}

void pack(const Frame& F, std::vector<float>& xy, std::vector<int32_t>& oct, std::vector<float>& ang, std::vector<unsigned char>& desc) {
  const size_t n = F.mvKeysUn.size();
  xy.resize(2 * n); oct.resize(n); ang.resize(n); desc.resize(32 * n);
  for (size_t i = 0; i < n; i++) {
    xy[2 * i] = F.mvKeysUn[i].pt.x; xy[2 * i + 1] = F.mvKeysUn[i].pt.y;
    oct[i] = F.mvKeysUn[i].octave; ang[i] = F.mvKeysUn[i].angle;
    std::memcpy(&desc[32 * i], F.mDescriptors.ptr((int)i), 32);
  }
}
}  // namespace

// src/ORBmatcher.cc:2083
int ORBmatcher::DescriptorDistance(const cv::Mat& a, const cv::Mat& b) { return orb_descriptor_distance(a.ptr(0), b.ptr(0)); }

// src/ORBmatcher.cc:573 (called at src/Tracking.cc:926 with windowSize 100)
int ORBmatcher::SearchForInitialization(Frame& F1, Frame& F2, std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12,
                                        int windowSize) {
  const int n1 = (int)F1.mvKeysUn.size(), n2 = (int)F2.mvKeysUn.size();
  vnMatches12 = std::vector<int>(n1, -1);
  if (n1 == 0 || n2 == 0) return 0;
  std::vector<float> xy1, xy2, a1, a2;
  std::vector<int32_t> o1, o2;
  std::vector<unsigned char> d1, d2;
  pack(F1, xy1, o1, a1, d1);
  pack(F2, xy2, o2, a2, d2);
  orb_frame_view v1 = {n1, xy1.data(), o1.data(), a1.data(), d1.data()};
  orb_frame_view v2 = {n2, xy2.data(), o2.data(), a2.data(), d2.data()};
  // the grid of F2 spans Frame's static undistorted image bounds (src/Frame.cc:45-48, 590-670)
  orb_match_params mp = {mfNNratio, mbCheckOrientation ? 1 : 0, windowSize, 0, Frame::mnMinX, Frame::mnMaxX, Frame::mnMinY, Frame::mnMaxY};
  static_assert(sizeof(cv::Point2f) == 8, "cv::Point2f must be two floats");
  int nmatches = 0;
  const int st = orb_search_for_initialization(t_matcher.get(std::max(n1, n2)), &v1, &v2, &mp, reinterpret_cast<float*>(vbPrevMatched.data()),
                                               vnMatches12.data(), &nmatches, nullptr, nullptr);
  if (st != ORB_OK) throw std::runtime_error(std::string("ORBmatcher::SearchForInitialization (liborb_b200): ") + orb_last_error());
  return nmatches;
}

// src/ORBmatcher.cc:1710-1860 (called at src/Tracking.cc:1047 TrackWithMotionModel): every map point of the last frame is
// projected with the current pose and takes the best keypoint in its window that no earlier map point holds; rotation
// histogram at the end. One call of the library's host-memory entry point does the projection, the grid, the scan and the
// in-order commit on the GPU.
int ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono) {
  // bForward / bBackward exactly as :1717-1730 (the caller's own cv::Mat arithmetic)
  const cv::Mat Rcw = CurrentFrame.mTcw.rowRange(0, 3).colRange(0, 3);
  const cv::Mat tcw = CurrentFrame.mTcw.rowRange(0, 3).col(3);
  const cv::Mat twc = -Rcw.t() * tcw;
  const cv::Mat Rlw = LastFrame.mTcw.rowRange(0, 3).colRange(0, 3);
  const cv::Mat tlw = LastFrame.mTcw.rowRange(0, 3).col(3);
  const cv::Mat tlc = Rlw * twc + tlw;
  const bool bForward = tlc.at<float>(2) > CurrentFrame.mb && !bMono;
  const bool bBackward = -tlc.at<float>(2) > CurrentFrame.mb && !bMono;

  const int nc = CurrentFrame.N, nl = LastFrame.N;
  if (nc == 0 || nl == 0) return 0;
  static_assert(sizeof(cv::KeyPoint) == sizeof(orb_keypoint), "cv::KeyPoint must be the 28-byte record of orb_keypoint");
  std::vector<unsigned char> curDesc((size_t)nc * 32), occupied(nc, 0), flags(nl, 0), mpDesc((size_t)nl * 32, 0);
  std::vector<float> Xw((size_t)nl * 3, 0.f);
  for (int i = 0; i < nc; i++) {
    std::memcpy(&curDesc[(size_t)i * 32], CurrentFrame.mDescriptors.ptr(i), 32);
    MapPoint* p = CurrentFrame.mvpMapPoints[i];
    occupied[i] = p && p->Observations() > 0;                       // :1795-1797
  }
  for (int i = 0; i < nl; i++) {
    MapPoint* p = LastFrame.mvpMapPoints[i];
    if (!p || LastFrame.mvbOutlier[i]) continue;                    // :1736-1740
    const cv::Mat x3Dw = p->GetWorldPos();
    Xw[3 * i] = x3Dw.at<float>(0); Xw[3 * i + 1] = x3Dw.at<float>(1); Xw[3 * i + 2] = x3Dw.at<float>(2);
    const cv::Mat d = p->GetDescriptor();
    std::memcpy(&mpDesc[(size_t)i * 32], d.ptr(0), 32);
    flags[i] = (unsigned char)(1 | (p->Observations() > 0 ? 2 : 0));
  }
  float Tcw[16];
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) Tcw[4 * r + c] = CurrentFrame.mTcw.at<float>(r, c);
  const float cam4[4] = {Frame::fx, Frame::fy, Frame::cx, Frame::cy};
  const float bounds4[4] = {Frame::mnMinX, Frame::mnMaxX, Frame::mnMinY, Frame::mnMaxY};
  orb_last_frame_search a;
  std::memset(&a, 0, sizeof a);
  a.n_cur = nc;
  a.cur_keypoints_un = reinterpret_cast<const orb_keypoint*>(CurrentFrame.mvKeysUn.data());
  a.cur_descriptors = curDesc.data();
  a.cur_uright = CurrentFrame.mvuRight.empty() ? nullptr : CurrentFrame.mvuRight.data();
  a.cur_occupied = occupied.data();
  a.bounds4 = bounds4;
  a.n_last = nl;
  a.last_keypoints = reinterpret_cast<const orb_keypoint*>(LastFrame.mvKeys.data());
  a.last_world_pos = Xw.data();
  a.last_mp_flags = flags.data();
  a.last_mp_descriptors = mpDesc.data();
  a.Tcw = Tcw;
  a.direction = bForward ? 1 : (bBackward ? 2 : 0);
  a.cam4 = cam4;
  a.mbf = CurrentFrame.mbf;
  a.th = th;
  a.scale_factors = CurrentFrame.mvScaleFactors.data();
  a.nlevels = (int)CurrentFrame.mvScaleFactors.size();
  a.th_dist = TH_HIGH;
  a.nn_ratio = mfNNratio;
  a.check_orientation = mbCheckOrientation ? 1 : 0;
  std::vector<int32_t> matchOfKp(nc, -1);
  int nmatches = 0;
  const int st = orb_search_by_projection_last_frame(0, &a, matchOfKp.data(), &nmatches);
  if (st != ORB_OK) throw std::runtime_error(std::string("ORBmatcher::SearchByProjection (liborb_b200): ") + orb_last_error());
  for (int i = 0; i < nc; i++)
    if (matchOfKp[i] >= 0) CurrentFrame.mvpMapPoints[i] = LastFrame.mvpMapPoints[matchOfKp[i]];   // :1813
  return nmatches;
}

// src/ORBmatcher.cc:72-169 (called at src/Tracking.cc:1463 SearchLocalPoints): every local map point that Frame::isInFrustum
// marked visible takes the best keypoint of its predicted level (or the one below) inside its window, unless an observed map
// point already holds it; ratio test only when best and second best share a level.
int ORBmatcher::SearchByProjection(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const float th) {
  const bool bFactor = th != 1.0;
  const int nc = F.N, nq = (int)vpMapPoints.size();
  if (nc == 0 || nq == 0) return 0;
  std::vector<orb_proj_query> q(nq);
  std::vector<unsigned char> qdesc((size_t)nq * 32, 0), curDesc((size_t)nc * 32), occupied(nc, 0);
  std::memset(q.data(), 0, (size_t)nq * sizeof(orb_proj_query));
  for (int i = 0; i < nq; i++) {
    MapPoint* pMP = vpMapPoints[i];
    if (!pMP->mbTrackInView || pMP->isBad()) continue;              // :80-84: flags stay 0, the query is skipped
    const int level = pMP->mnTrackScaleLevel;
    float r = RadiusByViewingCos(pMP->mTrackViewCos);               // :88
    if (bFactor) r *= th;
    q[i].u = pMP->mTrackProjX; q[i].v = pMP->mTrackProjY; q[i].ur = pMP->mTrackProjXR;
    q[i].radius = r * F.mvScaleFactors[level];                      // :95
    q[i].min_level = level - 1; q[i].max_level = level;             // :96
    q[i].flags = 1 | (pMP->Observations() > 0 ? 2 : 0);
    const cv::Mat d = pMP->GetDescriptor();
    std::memcpy(&qdesc[(size_t)i * 32], d.ptr(0), 32);
  }
  for (int i = 0; i < nc; i++) {
    std::memcpy(&curDesc[(size_t)i * 32], F.mDescriptors.ptr(i), 32);
    MapPoint* p = F.mvpMapPoints[i];
    occupied[i] = p && p->Observations() > 0;                       // :116-118
  }
  const float bounds4[4] = {Frame::mnMinX, Frame::mnMaxX, Frame::mnMinY, Frame::mnMaxY};
  const orb_search_params sp = {ORB_SEARCH_RATIO_LEVEL, TH_HIGH, mfNNratio, 0};
  std::vector<int32_t> matchOfKp(nc, -1);
  int nmatches = 0;
  const int st = orb_search_by_projection_host(0, nc, reinterpret_cast<const orb_keypoint*>(F.mvKeysUn.data()), curDesc.data(),
                                               F.mvuRight.empty() ? nullptr : F.mvuRight.data(), occupied.data(), bounds4, nq, q.data(),
                                               qdesc.data(), &sp, matchOfKp.data(), &nmatches);
  if (st != ORB_OK) throw std::runtime_error(std::string("ORBmatcher::SearchByProjection (liborb_b200): ") + orb_last_error());
  for (int i = 0; i < nc; i++)
    if (matchOfKp[i] >= 0) F.mvpMapPoints[i] = vpMapPoints[matchOfKp[i]];   // :161
  return nmatches;
}

// src/ORBmatcher.cc:247-420 (Tracking::TrackReferenceKeyFrame, Relocalization): features of the keyframe that hold a good map
// point are matched, inside their vocabulary node, against the frame's features; ratio test, one map point per frame feature,
// rotation histogram.
int ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches) {
  const std::vector<MapPoint*> vpMapPointsKF = pKF->GetMapPointMatches();
  vpMapPointMatches = std::vector<MapPoint*>(F.N, static_cast<MapPoint*>(NULL));
  const int n1 = (int)vpMapPointsKF.size(), n2 = F.N;
  if (n1 == 0 || n2 == 0) return 0;
  std::vector<unsigned char> usable(n1, 0);
  for (int i = 0; i < n1; i++) usable[i] = vpMapPointsKF[i] && !vpMapPointsKF[i]->isBad();   // :283-287
  const std::vector<int32_t> node1 = node_of_features(pKF->mFeatVec, n1), node2 = node_of_features(F.mFeatVec, n2);
  const std::vector<unsigned char> d1 = rows_of(pKF->mDescriptors, n1), d2 = rows_of(F.mDescriptors, n2);
  const orb_search_params sp = {ORB_SEARCH_RATIO, TH_LOW, mfNNratio, mbCheckOrientation ? 1 : 0};
  std::vector<int32_t> ofKp(n2, -1), ofQ(n1, -1);
  int nmatches = 0;
  const int st = orb_search_by_bow_host(0, n1, reinterpret_cast<const orb_keypoint*>(pKF->mvKeysUn.data()), d1.data(), node1.data(),
                                        usable.data(), n2, reinterpret_cast<const orb_keypoint*>(F.mvKeysUn.data()), d2.data(), node2.data(),
                                        nullptr, &sp, ofKp.data(), ofQ.data(), &nmatches);
  if (st != ORB_OK) throw std::runtime_error(std::string("ORBmatcher::SearchByBoW (liborb_b200): ") + orb_last_error());
  for (int i = 0; i < n2; i++)
    if (ofKp[i] >= 0) vpMapPointMatches[i] = vpMapPointsKF[ofKp[i]];   // :343
  return nmatches;
}

// src/ORBmatcher.cc:729-880 (LoopClosing::ComputeSim3): map points of keyframe 1 against map points of keyframe 2.
int ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12) {
  const std::vector<MapPoint*> vpMapPoints1 = pKF1->GetMapPointMatches();
  const std::vector<MapPoint*> vpMapPoints2 = pKF2->GetMapPointMatches();
  vpMatches12 = std::vector<MapPoint*>(vpMapPoints1.size(), static_cast<MapPoint*>(NULL));
  const int n1 = (int)vpMapPoints1.size(), n2 = (int)vpMapPoints2.size();
  if (n1 == 0 || n2 == 0) return 0;
  std::vector<unsigned char> usable(n1, 0), notCandidate(n2, 0);
  for (int i = 0; i < n1; i++) usable[i] = vpMapPoints1[i] && !vpMapPoints1[i]->isBad();            // :770-774
  for (int i = 0; i < n2; i++) notCandidate[i] = !vpMapPoints2[i] || vpMapPoints2[i]->isBad();      // :788-793
  const std::vector<int32_t> node1 = node_of_features(pKF1->mFeatVec, n1), node2 = node_of_features(pKF2->mFeatVec, n2);
  const std::vector<unsigned char> d1 = rows_of(pKF1->mDescriptors, n1), d2 = rows_of(pKF2->mDescriptors, n2);
  const orb_search_params sp = {ORB_SEARCH_RATIO, TH_LOW - 1, mfNNratio, mbCheckOrientation ? 1 : 0};   // :813 tests bestDist1 < TH_LOW
  std::vector<int32_t> ofKp(n2, -1), ofQ(n1, -1);
  int nmatches = 0;
  const int st = orb_search_by_bow_host(0, n1, reinterpret_cast<const orb_keypoint*>(pKF1->mvKeysUn.data()), d1.data(), node1.data(),
                                        usable.data(), n2, reinterpret_cast<const orb_keypoint*>(pKF2->mvKeysUn.data()), d2.data(),
                                        node2.data(), notCandidate.data(), &sp, ofKp.data(), ofQ.data(), &nmatches);
  if (st != ORB_OK) throw std::runtime_error(std::string("ORBmatcher::SearchByBoW (liborb_b200): ") + orb_last_error());
  for (int i = 0; i < n1; i++)
    if (ofQ[i] >= 0) vpMatches12[i] = vpMapPoints2[ofQ[i]];   // :818
  return nmatches;
}

// src/ORBmatcher.cc:884-1100 (LocalMapping::CreateNewMapPoints): features of keyframe 1 without a map point against features of
// keyframe 2 without one, inside their vocabulary node, best distance <= TH_LOW, epipole and epipolar-line tests.
int ORBmatcher::SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, cv::Mat F12, std::vector<std::pair<size_t, size_t> >& vMatchedPairs,
                                       const bool bOnlyStereo) {
  // the epipole: keyframe 1's camera centre in keyframe 2's image, :893-901 (the caller's own cv::Mat arithmetic)
  cv::Mat Cw = pKF1->GetCameraCenter();
  cv::Mat R2w = pKF2->GetRotation();
  cv::Mat t2w = pKF2->GetTranslation();
  cv::Mat C2 = R2w * Cw + t2w;
  const float invz = 1.0f / C2.at<float>(2);
  orb_triangulation_pair pr;
  pr.ex = pKF2->fx * C2.at<float>(0) * invz + pKF2->cx;
  pr.ey = pKF2->fy * C2.at<float>(1) * invz + pKF2->cy;
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) pr.F12[3 * r + c] = F12.at<float>(r, c);
  pr.only_stereo = bOnlyStereo ? 1 : 0;
  vMatchedPairs.clear();
  const int n1 = pKF1->N, n2 = pKF2->N;
  if (n1 == 0 || n2 == 0) return 0;
  std::vector<unsigned char> has1(n1, 0), has2(n2, 0);
  for (int i = 0; i < n1; i++) has1[i] = pKF1->GetMapPoint(i) != NULL;    // :928-933
  for (int i = 0; i < n2; i++) has2[i] = pKF2->GetMapPoint(i) != NULL;    // :963-967
  const std::vector<int32_t> node1 = node_of_features(pKF1->mFeatVec, n1), node2 = node_of_features(pKF2->mFeatVec, n2);
  const std::vector<unsigned char> d1 = rows_of(pKF1->mDescriptors, n1), d2 = rows_of(pKF2->mDescriptors, n2);
  std::vector<int32_t> m12(n1, -1);
  int nmatches = 0;
  const int st = orb_search_for_triangulation_host(0, n1, reinterpret_cast<const orb_keypoint*>(pKF1->mvKeysUn.data()), d1.data(),
                                                   node1.data(), has1.data(), pKF1->mvuRight.data(), n2,
                                                   reinterpret_cast<const orb_keypoint*>(pKF2->mvKeysUn.data()), d2.data(), node2.data(),
                                                   has2.data(), pKF2->mvuRight.data(), &pr, pKF2->mvScaleFactors.data(),
                                                   pKF2->mvLevelSigma2.data(), (int)pKF2->mvScaleFactors.size(), mbCheckOrientation ? 1 : 0,
                                                   m12.data(), &nmatches);
  if (st != ORB_OK) throw std::runtime_error(std::string("ORBmatcher::SearchForTriangulation (liborb_b200): ") + orb_last_error());
  vMatchedPairs.reserve(nmatches);
  for (int i = 0; i < n1; i++)
    if (m12[i] >= 0) vMatchedPairs.push_back(std::make_pair((size_t)i, (size_t)m12[i]));   // :1089-1094
  return nmatches;
}

}  // namespace ORB_SLAM2
