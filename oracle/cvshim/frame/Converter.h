// oracle/cvshim/frame/Converter.h — TEST INFRASTRUCTURE ONLY.
//
// Stands in for the reference's include/Converter.h when src/Frame.cc is compiled for oracle/_ref: the real
// header pulls in Eigen and g2o, neither of which is in this image, and Frame.cc uses exactly one function of
// it (Frame::ComputeBoW, Frame.cc:711), which the pinned path never calls. Placed on an include path that is
// searched before $(REFERENCE)/include (see oracle/Makefile).
#pragma once
#include <vector>
#include <opencv2/core/core.hpp>
namespace ORB_SLAM2 {
class Converter {
 public:
  static std::vector<cv::Mat> toDescriptorVector(const cv::Mat& Descriptors);
};
}  // namespace ORB_SLAM2
