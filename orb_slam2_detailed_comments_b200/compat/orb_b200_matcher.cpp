// orb_b200_matcher.cpp — the ORBmatcher members of the hot path, for the reference's build.
//
// Compiled against the reference's own, UNMODIFIED include/ORBmatcher.h (all eleven methods stay declared there). A
// maintainer deletes the bodies of ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:2083-2103),
// ORBmatcher::SearchForInitialization (:573-717), ORBmatcher::SearchByProjection(Frame&, const Frame&, float, bool)
// (:1710-1860, the per-frame tracking search) and ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, float)
// (:72-169, the local-map search) from src/ORBmatcher.cc and adds this file; the other seven methods keep their CPU
// bodies in ORBmatcher.cc (RadiusByViewingCos :171 stays there too and is used here). (This repository's test build of the reference does the same without touching the
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

}  // namespace ORB_SLAM2
