// orb_b200_frame.cpp — Frame::ComputeStereoMatches (src/Frame.cc:831-1082) for the reference's build.
//
// Compiled against the reference's own, UNMODIFIED include/Frame.h. The stereo constructor (src/Frame.cc:121-158) keeps
// its shape: two threads call Frame::ExtractORB -> ORBextractor::operator() (orb_b200_extractor.cpp), are joined, and
// ComputeStereoMatches() is called. Here that member pairs the two extractors' device-resident results with ONE call
// (orb_stereo_match: row-band candidates, best Hamming, 11x11 SAD slide with parabola fit, median cut - all on the GPU)
// instead of walking mvImagePyramid on the host. A maintainer deletes the body at src/Frame.cc:831-1082 and adds this
// file (this repository's test build of the reference weakens the symbol in the compiled Frame.o instead).
#include <stdexcept>
#include <string>
#include <vector>

#include "Frame.h"   // the reference's header (include/Frame.h)

#include "orb_b200_registry.hpp"

namespace ORB_SLAM2 {

void Frame::ComputeStereoMatches() {
  mvuRight = std::vector<float>(N, -1.0f);   // :836-837
  mvDepth = std::vector<float>(N, -1.0f);
  if (N == 0) return;
  orb_b200_compat::ExtractorEntry* L = orb_b200_compat::entry_of(mpORBextractorLeft);
  orb_b200_compat::ExtractorEntry* R = orb_b200_compat::entry_of(mpORBextractorRight);
  if (!L || !R) throw std::runtime_error("Frame::ComputeStereoMatches (liborb_b200): extractors are not registered");
  // The stereo constructor calls this member at src/Frame.cc:157, BEFORE it sets mb = mbf / fx (:196): the reference reads
  // an uninitialised `mb` here (minZ = mb, maxD = mbf / minZ, src/Frame.cc:847-851). The drop-in uses the value the
  // constructor is about to store, computed from this frame's own calibration matrix.
  const float baseline = mbf / mK.at<float>(0, 0);
  int n = 0;
  const int st = orb_stereo_match(L->handle, R->handle, mbf, baseline, mvuRight.data(), mvDepth.data(), &n);
  if (st != ORB_OK) throw std::runtime_error(std::string("Frame::ComputeStereoMatches (liborb_b200): ") + orb_last_error());
  if (n != N) throw std::runtime_error("Frame::ComputeStereoMatches (liborb_b200): keypoint count changed since ExtractORB");
}

}  // namespace ORB_SLAM2
