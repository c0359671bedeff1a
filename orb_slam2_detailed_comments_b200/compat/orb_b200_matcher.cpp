// orb_b200_matcher.cpp — the two ORBmatcher members of the hot path, for the reference's build.
//
// Compiled against the reference's own, UNMODIFIED include/ORBmatcher.h (all eleven methods stay declared there). A
// maintainer deletes the bodies of ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:2083-2103) and
// ORBmatcher::SearchForInitialization (:573-717) from src/ORBmatcher.cc and adds this file; the other nine methods keep
// their CPU bodies in ORBmatcher.cc. (This repository's test build of the reference does the same without touching the
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

}  // namespace ORB_SLAM2
