// Compile/link probe for the C++ adapter (no OpenCV in this image: stand-in types).
// argv[1] == "run" additionally extracts + matches two frames read from stdin-free synthetic data.
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "../orb_slam2_detailed_comments_b200/compat/orb_b200_compat.hpp"

int main(int argc, char** argv) {
  using namespace ORB_SLAM2;
  orbcv::Mat a, b;
  a.create(1, 32); b.create(1, 32);
  for (int i = 0; i < 32; i++) { a.data[i] = (unsigned char)(i * 7); b.data[i] = (unsigned char)(i * 7 ^ 0x11); }
  const int d = ORBmatcher::DescriptorDistance(a, b);
  std::printf("distance %d\n", d);
  if (d != 64) return 2;
  if (argc > 1 && std::string(argv[1]) == "run") {
    const int W = 640, H = 480;
    orbcv::Mat img; img.create(H, W);
    unsigned s = 12345u;
    for (int y = 0; y < H; y++)
      for (int x = 0; x < W; x++) {
        s = s * 1664525u + 1013904223u;
        int v = ((x / 24 + y / 24) & 1) * 120 + 60 + (int)((s >> 24) & 15);
        img.data[y * W + x] = (unsigned char)v;
      }
    ORBextractor ex(1000, 1.2f, 8, 20, 7);
    FrameLike F1, F2;
    ex(img, orbcv::Mat(), F1.mvKeysUn, F1.mDescriptors);
    ex(img, orbcv::Mat(), F2.mvKeysUn, F2.mDescriptors);
    F1.mnMaxX = F2.mnMaxX = W; F1.mnMaxY = F2.mnMaxY = H;
    std::vector<orbcv::Point2f> prev(F1.mvKeysUn.size());
    for (size_t i = 0; i < prev.size(); i++) prev[i] = F1.mvKeysUn[i].pt;
    std::vector<int> m12;
    ORBmatcher matcher(0.9f, true);
    int n = matcher.SearchForInitialization(F1, F2, prev, m12, 100);
    std::printf("keypoints %zu matches %d pyramid0 %dx%d\n", F1.mvKeysUn.size(), n, ex.mvImagePyramid[0].cols, ex.mvImagePyramid[0].rows);
    if (F1.mvKeysUn.empty() || n <= 0) return 3;
    if (argc > 2) {   // what the adapter returned, for the value comparison with the oracle (tests/test_compat_build.py)
      static_assert(sizeof(orbcv::KeyPoint) == 28, "cv::KeyPoint layout");
      FILE* f = std::fopen(argv[2], "wb");
      if (!f) return 4;
      const int nk = (int)F1.mvKeysUn.size(), nm = (int)m12.size();
      std::fwrite(&nk, 4, 1, f);
      std::fwrite(F1.mvKeysUn.data(), 28, nk, f);
      for (int i = 0; i < nk; i++) std::fwrite(F1.mDescriptors.data + (size_t)i * 32, 1, 32, f);
      std::fwrite(&n, 4, 1, f);
      std::fwrite(&nm, 4, 1, f);
      std::fwrite(m12.data(), 4, nm, f);
      std::fclose(f);
    }
    // latency of the drop-in call itself (operator() of the adapter: keypoints, descriptors and mvImagePyramid views)
    for (int i = 0; i < 20; i++) ex(img, orbcv::Mat(), F2.mvKeysUn, F2.mDescriptors);
    const auto t0 = std::chrono::steady_clock::now();
    const int reps = 200;
    for (int i = 0; i < reps; i++) ex(img, orbcv::Mat(), F2.mvKeysUn, F2.mDescriptors);
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / reps;
    std::printf("adapter operator() %dx%d, %zu keypoints: %.3f ms per call\n", W, H, F2.mvKeysUn.size(), ms);
  } else {
    try { ORBextractor ex(1000, 1.2f, 8, 20, 7); std::printf("extractor created\n"); }
    catch (const std::exception& e) { std::printf("no device: %s\n", e.what()); }
  }
  return 0;
}
