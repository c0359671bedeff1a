// orb_b200_extractor.cpp — replaces src/ORBextractor.cc of the reference in its build.
//
// Compiled against the reference's own, UNMODIFIED include/ORBextractor.h: it defines the two members of
// ORB_SLAM2::ORBextractor that the rest of the reference calls - the constructor (src/ORBextractor.cc:469-571, called at
// src/Tracking.cc:175-188) and operator() (src/ORBextractor.cc:1533-1649, called at src/Frame.cc:437,443) - on top of the
// C ABI of liborb_b200.so. Frame.cc, Tracking.cc and every header stay as they are; the getters of the header read the
// tables filled here and Frame::ComputeStereoMatches finds mvImagePyramid as before.
// There is no CPU path: without a usable GPU the constructor throws.
#include <cassert>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>

#include "ORBextractor.h"   // the reference's header (include/ORBextractor.h)

#include "orb_b200_registry.hpp"

namespace ORB_SLAM2 {

using orb_b200_compat::ExtractorEntry;

ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST)
    : nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), iniThFAST(_iniThFAST), minThFAST(_minThFAST) {
  static_assert(sizeof(cv::KeyPoint) == sizeof(orb_keypoint), "cv::KeyPoint must be the 28-byte record of orb_keypoint");
  orb_params p = {_nfeatures, _scaleFactor, _nlevels, _iniThFAST, _minThFAST};
  ExtractorEntry* e = new ExtractorEntry();
  // max_batch 2: any extractor can serve as the left eye of Frame::ComputeStereoMatches (orb_stereo_match)
  if (orb_create(&p, 0, 2, &e->handle) != ORB_OK) {
    delete e;
    throw std::runtime_error(std::string("ORBextractor (liborb_b200): ") + orb_last_error());
  }
  mvScaleFactor.resize(nlevels); mvInvScaleFactor.resize(nlevels);
  mvLevelSigma2.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
  mnFeaturesPerLevel.resize(nlevels);
  orb_get_scale_tables(e->handle, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(), mvInvLevelSigma2.data(),
                       mnFeaturesPerLevel.data());
  mvImagePyramid.resize(nlevels);
  e->views.resize(nlevels);
  std::lock_guard<std::mutex> lock(orb_b200_compat::registry_mutex());
  ExtractorEntry*& slot = orb_b200_compat::registry()[this];
  if (slot) {   // an extractor that lived at this address before (its destructor is an empty inline function)
    orb_destroy(slot->handle);
    delete slot;
  }
  slot = e;
}

void ORBextractor::operator()(cv::InputArray _image, cv::InputArray /*_mask: ignored by the reference too*/,
                              std::vector<cv::KeyPoint>& _keypoints, cv::OutputArray _descriptors) {
  if (_image.empty()) return;   // :1537
  cv::Mat image = _image.getMat();
  assert(image.type() == CV_8UC1);   // :1543
  ExtractorEntry* e = orb_b200_compat::entry_of(this);
  if (!e) throw std::runtime_error("ORBextractor (liborb_b200): extractor is not registered");
  const int cap = orb_max_keypoints_for_size(e->handle, image.cols, image.rows);
  e->kps.resize(cap > 0 ? cap : 1);
  e->desc.resize((size_t)(cap > 0 ? cap : 1) * 32);
  int n = 0;
  // ORB_B200_COMPAT_PYRAMID=0: no host copy of the pyramid (mvImagePyramid stays empty). The reference reads it only in
  // Frame::ComputeStereoMatches, whose drop-in body works on the device-resident pyramid: 0.15 instead of 0.25 ms per call.
  static const bool wantViews = [] { const char* v = std::getenv("ORB_B200_COMPAT_PYRAMID"); return !(v && v[0] == '0'); }();
  const int st = orb_extract(e->handle, image.data, image.cols, image.rows, (size_t)image.step, e->kps.data(), cap > 0 ? cap : 1, &n,
                             e->desc.data(), wantViews ? e->views.data() : nullptr);
  if (st != ORB_OK) throw std::runtime_error(std::string("ORBextractor::operator() (liborb_b200): ") + orb_last_error());
  _keypoints.clear();
  _keypoints.resize(n);
  if (n) std::memcpy(static_cast<void*>(_keypoints.data()), e->kps.data(), (size_t)n * sizeof(orb_keypoint));
  if (n == 0) {
    _descriptors.release();   // :1572
  } else {
    _descriptors.create(n, 32, CV_8U);   // :1577
    cv::Mat d = _descriptors.getMat();
    for (int i = 0; i < n; i++) std::memcpy(d.ptr(i), &e->desc[(size_t)i * 32], 32);
  }
  // mvImagePyramid (include/ORBextractor.h:162): views with >= 19 readable border pixels around them, valid until the next
  // call - what Frame::ComputeStereoMatches slices at src/Frame.cc:967,1003
  for (int l = 0; l < nlevels; l++)
    mvImagePyramid[l] = wantViews ? cv::Mat(e->views[l].height, e->views[l].width, CV_8UC1, e->views[l].data, (size_t)e->views[l].step) : cv::Mat();
}

}  // namespace ORB_SLAM2
