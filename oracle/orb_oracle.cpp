// oracle/orb_oracle.cpp — CPU ORACLE. TEST INFRASTRUCTURE ONLY.
//
// A from-scratch CPU restatement of the reference's ORB front-end hot path
// (electech6/ORB_SLAM2_detailed_comments): ORBextractor::operator() and
// ORBmatcher::DescriptorDistance / SearchForInitialization. Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library; the product (orb_slam2_detailed_comments_b200/) never does.
//
// PARITY PINNING: the reference holds no golden vectors, tests or fixtures for this path.
// The oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, run here:
//   * oracle/_ref/liborbref.so = the reference's own src/ORBextractor.cc, src/ORBmatcher.cc and
//     src/Frame.cc compiled unmodified (oracle/Makefile target `ref`, wrapper oracle/ref_wrap.cpp,
//     OpenCV stand-in oracle/cvshim/): tests/test_oracle_vs_ref.py and test_oracle_vs_ref_matcher.py
//     require bit-identical keypoints, descriptors, pyramids, matches, grids, mvuRight / mvDepth;
//   * every OpenCV primitive restated below (resize INTER_LINEAR, copyMakeBorder REFLECT_101,
//     FAST 9/16 + NMS, GaussianBlur 7x7 s=2, fastAtan2, undistortPoints, gemm order) is checked
//     bit-exactly against python cv2 4.13.0 (tests/test_oracle_vs_cv2.py, test_oracle_frame.py) -
//     these same functions serve the OpenCV calls of _ref;
//   * a cv2-driven restatement of the control flow (tests/cv2_reference.py) and golden vectors.
// The pixel arithmetic therefore is "the reference's code + OpenCV 4.13 primitives".
//
// Canonicalised non-determinism: DistributeOctTree sorts (size, node pointer) pairs
// (ORBextractor.cc:926); ties are broken by heap address in the reference. Canonical
// rule used here and by the CUDA path: addresses grow with creation order, i.e. among
// equal sizes the LATER-created node is expanded first.
//
// Build: g++ -O3 -march=x86-64-v3 -ffp-contract=off -shared -fPIC (see oracle/Makefile).
#include <algorithm>
#include <atomic>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <list>
#include <thread>
#include <utility>
#include <map>
#include <vector>

#include "../include/orb_pattern_data.h"

namespace {

typedef unsigned char u8;

const int kPatch = 31;       // ORBextractor.cc:79
const int kHalfPatch = 15;   // ORBextractor.cc:80
const int kEdge = 19;        // ORBextractor.cc:81

struct KeyPoint {  // binary layout of cv::KeyPoint (28 bytes)
  float x, y, size, angle, response;
  int octave, class_id;
};

// round-half-to-even of a float, like cvRound (SSE cvtss2si under default rounding).
inline int cv_round(float v) { return (int)lrintf(v); }
inline int cv_round(double v) { return (int)lrint(v); }

struct Image {
  int w = 0, h = 0, step = 0;
  std::vector<u8> buf;
  u8* origin = nullptr;  // pixel (0,0); for bordered images this sits kEdge rows/cols inside buf
  const u8* row(int y) const { return origin + (ptrdiff_t)y * step; }
  u8* row(int y) { return origin + (ptrdiff_t)y * step; }
};

// ---------------------------------------------------------------- parameters
// ORBextractor::ORBextractor, ORBextractor.cc:469-571
struct Params {
  int nfeatures, nlevels, iniTh, minTh;
  double scaleFactor;  // member is double, ctor argument float (ORBextractor.h:93,207)
  std::vector<float> scale, invScale, sigma2, invSigma2;
  std::vector<int> perLevel;
  int umax[kHalfPatch + 1];

  Params(int nf, float sf, int nl, int ini, int mn)
      : nfeatures(nf), nlevels(nl), iniTh(ini), minTh(mn), scaleFactor(sf) {
    scale.assign(nl, 1.0f);
    sigma2.assign(nl, 1.0f);
    for (int i = 1; i < nl; i++) {
      scale[i] = (float)(scale[i - 1] * scaleFactor);  // float*double -> float  (:488)
      sigma2[i] = scale[i] * scale[i];
    }
    invScale.resize(nl);
    invSigma2.resize(nl);
    for (int i = 0; i < nl; i++) {
      invScale[i] = 1.0f / scale[i];
      invSigma2[i] = 1.0f / sigma2[i];
    }
    perLevel.assign(nl, 0);
    float factor = (float)(1.0f / scaleFactor);
    float want = nf * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nl));
    int sum = 0;
    for (int l = 0; l < nl - 1; l++) {
      perLevel[l] = cv_round(want);
      sum += perLevel[l];
      want *= factor;
    }
    perLevel[nl - 1] = std::max(nf - sum, 0);
    // quarter-circle row extents of the orientation patch (:542-570)
    int vmax = (int)std::floor(kHalfPatch * std::sqrt(2.f) / 2 + 1);
    int vmin = (int)std::ceil(kHalfPatch * std::sqrt(2.f) / 2);
    const double hp2 = kHalfPatch * kHalfPatch;
    for (int v = 0; v <= kHalfPatch; v++) umax[v] = 0;
    for (int v = 0; v <= vmax; v++) umax[v] = cv_round(std::sqrt(hp2 - v * v));
    for (int v = kHalfPatch, v0 = 0; v >= vmin; --v) {
      while (umax[v0] == umax[v0 + 1]) ++v0;
      umax[v] = v0;
      ++v0;
    }
  }
};

// ---------------------------------------------------------------- primitives
// copyMakeBorder(..., BORDER_REFLECT_101): gfedcb|abcdefgh|gfedcba
inline int reflect101(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) {
    if (p < 0) p = -p;
    else p = 2 * (n - 1) - p;
  }
  return p;
}

// fill the kEdge-wide frame around an image whose interior is already written
void fill_border(Image& im) {
  const int b = kEdge;
  for (int y = 0; y < im.h; y++) {
    u8* r = im.row(y);
    for (int k = 1; k <= b; k++) {
      r[-k] = r[reflect101(-k, im.w)];
      r[im.w - 1 + k] = r[reflect101(im.w - 1 + k, im.w)];
    }
  }
  for (int k = 1; k <= b; k++) {
    memcpy(im.row(-k) - b, im.row(reflect101(-k, im.h)) - b, im.w + 2 * b);
    memcpy(im.row(im.h - 1 + k) - b, im.row(reflect101(im.h - 1 + k, im.h)) - b, im.w + 2 * b);
  }
}

void alloc_bordered(Image& im, int w, int h) {
  im.w = w;
  im.h = h;
  im.step = w + 2 * kEdge;
  im.buf.assign((size_t)im.step * (h + 2 * kEdge), 0);
  im.origin = im.buf.data() + (size_t)kEdge * im.step + kEdge;
}

// cv::resize(..., INTER_LINEAR) for 8-bit single channel, OpenCV 4.x fixed point:
// 11-bit horizontal/vertical coefficients, the (>>4, *c >>16, +2 >>2) vertical pass.
void resize_linear_u8(const u8* src, int sw, int sh, int sstep, u8* dst, int dw, int dh, int dstep) {
  std::vector<int> sx0(dw), sy0(dh);
  std::vector<short> cx(2 * dw), cy(2 * dh);
  auto taps = [](int dsize, int ssize, int* s0, short* c) {
    double scale = 1.0 / ((double)dsize / ssize);
    for (int d = 0; d < dsize; d++) {
      float f = (float)((d + 0.5) * scale - 0.5);
      int s = (int)std::floor(f);
      f -= s;
      if (s < 0) { s = 0; f = 0.f; }
      if (s >= ssize - 1) { s = ssize - 1; f = 0.f; }
      s0[d] = s;
      c[2 * d] = (short)cv_round((1.f - f) * 2048.f);
      c[2 * d + 1] = (short)cv_round(f * 2048.f);
    }
  };
  taps(dw, sw, sx0.data(), cx.data());
  taps(dh, sh, sy0.data(), cy.data());
  std::vector<int> h0(dw), h1(dw);
  for (int y = 0; y < dh; y++) {
    const u8* r0 = src + (ptrdiff_t)sy0[y] * sstep;
    const u8* r1 = src + (ptrdiff_t)std::min(sy0[y] + 1, sh - 1) * sstep;
    for (int x = 0; x < dw; x++) {
      int a = sx0[x], b = std::min(a + 1, sw - 1);
      h0[x] = r0[a] * cx[2 * x] + r0[b] * cx[2 * x + 1];
      h1[x] = r1[a] * cx[2 * x] + r1[b] * cx[2 * x + 1];
    }
    u8* d = dst + (ptrdiff_t)y * dstep;
    int c0 = cy[2 * y], c1 = cy[2 * y + 1];
    for (int x = 0; x < dw; x++) {
      int v = (((c0 * (h0[x] >> 4)) >> 16) + ((c1 * (h1[x] >> 4)) >> 16) + 2) >> 2;
      d[x] = (u8)std::min(std::max(v, 0), 255);
    }
  }
}

// FAST-9/16 corner score: the largest threshold at which p is still a corner
// (cv::FAST response). Circle offsets (dx,dy) in OpenCV's order.
const int kCircle[16][2] = {{0, 3},  {1, 3},   {2, 2},   {3, 1},   {3, 0},  {3, -1}, {2, -2}, {1, -3},
                            {0, -3}, {-1, -3}, {-2, -2}, {-3, -1}, {-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}};

inline int fast_score(const u8* p, int step) {
  int d[16 + 9];
  int c = p[0];
  for (int k = 0; k < 16; k++) d[k] = c - p[kCircle[k][1] * step + kCircle[k][0]];
  for (int k = 16; k < 25; k++) d[k] = d[k - 16];
  int best_dark = INT_MIN, best_bright = INT_MIN;
  for (int s = 0; s < 16; s++) {
    int mn = d[s], mx = d[s];
    for (int k = 1; k < 9; k++) {
      mn = std::min(mn, d[s + k]);
      mx = std::max(mx, d[s + k]);
    }
    best_dark = std::max(best_dark, mn);       // all 9 darker than centre by > mn-1
    best_bright = std::max(best_bright, -mx);  // all 9 brighter
  }
  return std::max(best_dark, best_bright) - 1;
}

// cv::FAST(img, kps, threshold, nonmaxSuppression=true): detections in row-major order.
// The score map restricted to [3,w-4]x[3,h-4] of the SUB-IMAGE; 0 outside / non-corner.
struct FastHit { int x, y, score; };

// Cheap necessary condition for "corner at threshold t": every 9-arc contains one pixel of
// each antipodal pair, so all 8 pairs must hold a pixel of the arc's class (same pruning idea
// as OpenCV's scalar FAST loop). Keeps the CPU baseline honest; never changes a result.
inline bool fast_maybe_corner(const u8* p, int step, int t) {
  const int lo = p[0] - t, hi = p[0] + t;
  auto cls = [&](int k) {
    int v = p[kCircle[k][1] * step + kCircle[k][0]];
    return (v < lo ? 1 : 0) | (v > hi ? 2 : 0);
  };
  int d = cls(0) | cls(8);
  if (!d) return false;
  d &= cls(4) | cls(12);
  if (!d) return false;
  d &= cls(2) | cls(10);
  d &= cls(6) | cls(14);
  if (!d) return false;
  d &= cls(1) | cls(9);
  d &= cls(3) | cls(11);
  d &= cls(5) | cls(13);
  d &= cls(7) | cls(15);
  return d != 0;
}

void fast_detect(const u8* img, int w, int h, int step, int threshold, bool nms, std::vector<FastHit>& out) {
  out.clear();
  if (w < 7 || h < 7) return;
  static thread_local std::vector<int> S;
  S.assign((size_t)w * h, 0);
  bool any = false;
  for (int y = 3; y < h - 3; y++)
    for (int x = 3; x < w - 3; x++) {
      const u8* p = img + (ptrdiff_t)y * step + x;
      if (!fast_maybe_corner(p, step, threshold)) continue;
      int s = fast_score(p, step);
      if (s >= threshold) { S[(size_t)y * w + x] = s; any = true; }
    }
  if (!any) return;
  for (int y = 3; y < h - 3; y++)
    for (int x = 3; x < w - 3; x++) {
      int s = S[(size_t)y * w + x];
      if (s < threshold || s == 0) continue;
      bool keep = true;
      if (nms)
        for (int dy = -1; dy <= 1 && keep; dy++)
          for (int dx = -1; dx <= 1; dx++)
            if ((dx || dy) && !(s > S[(size_t)(y + dy) * w + x + dx])) { keep = false; break; }
      if (keep) out.push_back({x, y, s});
    }
}

// cv::GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101), 8-bit fixed-point path of OpenCV 4.x.
void gauss7_u8(const u8* src, int w, int h, int sstep, u8* dst, int dstep) {
  static const int K[7] = {18, 34, 48, 56, 48, 34, 18};
  std::vector<uint16_t> H((size_t)w * h);
  for (int y = 0; y < h; y++) {
    const u8* r = src + (ptrdiff_t)y * sstep;
    uint16_t* o = &H[(size_t)y * w];
    for (int x = 0; x < w; x++) {
      if (x >= 3 && x < w - 3) {
        o[x] = (uint16_t)(18 * (r[x - 3] + r[x + 3]) + 34 * (r[x - 2] + r[x + 2]) + 48 * (r[x - 1] + r[x + 1]) + 56 * r[x]);
      } else {
        int acc = 0;
        for (int i = 0; i < 7; i++) acc += K[i] * r[reflect101(x + i - 3, w)];
        o[x] = (uint16_t)acc;
      }
    }
  }
  for (int y = 0; y < h; y++) {
    u8* d = dst + (ptrdiff_t)y * dstep;
    const uint16_t* rr[7];
    for (int j = 0; j < 7; j++) rr[j] = &H[(size_t)reflect101(y + j - 3, h) * w];
    for (int x = 0; x < w; x++) {
      uint32_t acc = 32768u + 18u * (rr[0][x] + rr[6][x]) + 34u * (rr[1][x] + rr[5][x]) + 48u * (rr[2][x] + rr[4][x]) +
                     56u * rr[3][x];
      d[x] = (u8)(acc >> 16);
    }
  }
}

// cv::fastAtan2(y, x): degrees in [0,360), float polynomial, no FMA (-ffp-contract=off).
float fast_atan2_deg(float y, float x) {
  const float s = (float)(180.0 / 3.14159265358979323846);
  const float p1 = 0.9997878412794807f * s, p3 = -0.3258083974640975f * s;
  const float p5 = 0.1555786518463281f * s, p7 = -0.04432655554792128f * s;
  float ax = std::fabs(x), ay = std::fabs(y), a, c, c2;
  if (ax >= ay) {
    c = ay / (ax + (float)DBL_EPSILON);
    c2 = c * c;
    a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  } else {
    c = ax / (ay + (float)DBL_EPSILON);
    c2 = c * c;
    a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  }
  if (x < 0) a = 180.f - a;
  if (y < 0) a = 360.f - a;
  return a;
}

// IC_Angle, ORBextractor.cc:94-141
float ic_angle(const Image& im, float px, float py, const int* umax) {
  int m01 = 0, m10 = 0;
  const u8* c = im.row(cv_round(py)) + cv_round(px);
  for (int u = -kHalfPatch; u <= kHalfPatch; ++u) m10 += u * c[u];
  for (int v = 1; v <= kHalfPatch; ++v) {
    int vsum = 0, d = umax[v];
    for (int u = -d; u <= d; ++u) {
      int lo = c[u + v * im.step], hi = c[u - v * im.step];
      vsum += lo - hi;
      m10 += u * (lo + hi);
    }
    m01 += v * vsum;
  }
  return fast_atan2_deg((float)m01, (float)m10);
}

// computeOrbDescriptor, ORBextractor.cc:153-204
void brief_descriptor(const u8* img, int step, float px, float py, float angle_deg, u8* desc) {
  const float factorPI = (float)(3.14159265358979323846 / 180.f);
  float ang = angle_deg * factorPI;
  float a = cosf(ang), b = sinf(ang);
  const u8* c = img + (ptrdiff_t)cv_round(py) * step + cv_round(px);
  const signed char* pat = ORB_BIT_PATTERN_31;
  for (int i = 0; i < 32; i++) {
    int val = 0;
    for (int bit = 0; bit < 8; bit++, pat += 4) {
      float x0 = pat[0], y0 = pat[1], x1 = pat[2], y1 = pat[3];
      int t0 = c[cv_round(x0 * b + y0 * a) * step + cv_round(x0 * a - y0 * b)];
      int t1 = c[cv_round(x1 * b + y1 * a) * step + cv_round(x1 * a - y1 * b)];
      val |= (t0 < t1) << bit;
    }
    desc[i] = (u8)val;
  }
}

// ---------------------------------------------------------------- quadtree
struct Cand { float x, y; int score; };  // coordinates relative to the 16-px detection border

struct QNode {
  int x0, x1, y0, y1;       // [x0,x1) x [y0,y1)
  std::vector<int> keys;    // indices into the candidate array, in candidate order
  bool frozen = false;
  long seq = 0;             // creation counter = canonical stand-in for the heap address
  std::list<QNode>::iterator self;
};

// ExtractorNode::DivideNode, ORBextractor.cc:602-674
void divide(const QNode& n, const std::vector<Cand>& c, QNode out[4]) {
  int hx = (int)std::ceil((float)(n.x1 - n.x0) / 2);
  int hy = (int)std::ceil((float)(n.y1 - n.y0) / 2);
  int mx = n.x0 + hx, my = n.y0 + hy;
  out[0].x0 = n.x0; out[0].x1 = mx;   out[0].y0 = n.y0; out[0].y1 = my;
  out[1].x0 = mx;   out[1].x1 = n.x1; out[1].y0 = n.y0; out[1].y1 = my;
  out[2].x0 = n.x0; out[2].x1 = mx;   out[2].y0 = my;   out[2].y1 = n.y1;
  out[3].x0 = mx;   out[3].x1 = n.x1; out[3].y0 = my;   out[3].y1 = n.y1;
  for (int k : n.keys) {
    bool left = c[k].x < (float)mx, top = c[k].y < (float)my;
    out[left ? (top ? 0 : 2) : (top ? 1 : 3)].keys.push_back(k);
  }
  for (int q = 0; q < 4; q++) out[q].frozen = out[q].keys.size() == 1;
}

// ORBextractor::DistributeOctTree, ORBextractor.cc:688-1033. Returns candidate indices in
// final list order (front to back).
std::vector<int> distribute_quadtree(const std::vector<Cand>& c, int minX, int maxX, int minY, int maxY, int N,
                                     int* tie_sensitive) {
  if (tie_sensitive) *tie_sensitive = 0;
  const int nIni = (int)std::round((float)(maxX - minX) / (maxY - minY));
  const float hX = (float)(maxX - minX) / nIni;
  std::list<QNode> nodes;
  if (nIni <= 0) return std::vector<int>();  // portrait strips: the reference divides by zero here
  std::vector<QNode*> roots(nIni);
  long counter = 0;
  for (int i = 0; i < nIni; i++) {
    QNode r;
    r.x0 = (int)(hX * (float)i);
    r.x1 = (int)(hX * (float)(i + 1));
    r.y0 = 0;
    r.y1 = maxY - minY;
    // roots are appended at the back; give them descending seq so that "list order ==
    // descending seq" holds for them too (they never take part in the sorted phase).
    r.seq = -(long)i - 1;
    nodes.push_back(r);
    roots[i] = &nodes.back();
  }
  for (int k = 0; k < (int)c.size(); k++) roots[(int)(c[k].x / hX)]->keys.push_back(k);
  for (auto it = nodes.begin(); it != nodes.end();) {
    if (it->keys.size() == 1) { it->frozen = true; ++it; }
    else if (it->keys.empty()) it = nodes.erase(it);
    else ++it;
  }
  typedef std::pair<int, QNode*> SizedNode;
  std::vector<SizedNode> expandable;
  auto emit_children = [&](QNode ch[4], int* nToExpand) {
    for (int q = 0; q < 4; q++) {
      if (ch[q].keys.empty()) continue;
      ch[q].seq = counter++;
      nodes.push_front(ch[q]);
      nodes.front().self = nodes.begin();
      if (ch[q].keys.size() > 1) {
        if (nToExpand) ++*nToExpand;
        expandable.push_back(SizedNode((int)ch[q].keys.size(), &nodes.front()));
      }
    }
  };
  bool done = false;
  while (!done) {
    int prev = (int)nodes.size();
    int nToExpand = 0;
    expandable.clear();
    for (auto it = nodes.begin(); it != nodes.end();) {
      if (it->frozen) { ++it; continue; }
      QNode ch[4];
      divide(*it, c, ch);
      emit_children(ch, &nToExpand);
      it = nodes.erase(it);
    }
    if ((int)nodes.size() >= N || (int)nodes.size() == prev) {
      done = true;
    } else if ((int)nodes.size() + nToExpand * 3 > N) {
      while (!done) {
        prev = (int)nodes.size();
        std::vector<SizedNode> todo = expandable;
        expandable.clear();
        // canonical (size, creation order) ascending; walked from the back
        std::sort(todo.begin(), todo.end(), [](const SizedNode& a, const SizedNode& b) {
          return a.first != b.first ? a.first < b.first : a.second->seq < b.second->seq;
        });
        for (int j = (int)todo.size() - 1; j >= 0; j--) {
          QNode ch[4];
          divide(*todo[j].second, c, ch);
          emit_children(ch, nullptr);
          nodes.erase(todo[j].second->self);
          if ((int)nodes.size() >= N) {
            // the walk stopped at a member of a run of equal sizes: where the node count crosses N inside
            // that run depends on the order of its members (each split gains 0..3 nodes), i.e. on the tie
            // rule. Conservative: flagged whenever the stopping node's run has more than one member
            // (pinned against the reference under plain malloc, tests/test_oracle_vs_ref.py).
            if (tie_sensitive && ((j > 0 && todo[j - 1].first == todo[j].first) ||
                                  (j + 1 < (int)todo.size() && todo[j + 1].first == todo[j].first)))
              *tie_sensitive = 1;
            break;
          }
        }
        if ((int)nodes.size() >= N || (int)nodes.size() == prev) done = true;
      }
    }
  }
  std::vector<int> result;
  result.reserve(nodes.size());
  for (const QNode& n : nodes) {
    int best = n.keys[0];
    for (size_t k = 1; k < n.keys.size(); k++)
      if (c[n.keys[k]].score > c[best].score) best = n.keys[k];
    result.push_back(best);
  }
  return result;
}

// ---------------------------------------------------------------- extractor
struct LevelDebug {
  std::vector<Cand> cand;     // candidates entering the quadtree (relative coords)
  std::vector<int> kept;      // indices into cand, in list order
  int cells = 0, fallback_cells = 0, tie_sensitive = 0;
};

struct Extractor {
  Params p;
  std::vector<Image> pyr;     // bordered pyramid (mvImagePyramid)
  std::vector<Image> blur;    // blurred clones (compact)
  std::vector<LevelDebug> dbg;
  std::vector<KeyPoint> kps;
  std::vector<u8> desc;
  Extractor(int nf, float sf, int nl, int ini, int mn) : p(nf, sf, nl, ini, mn), pyr(nl), blur(nl), dbg(nl) {}

  // ORBextractor::ComputePyramid, ORBextractor.cc:1655-1724
  void compute_pyramid(const u8* img, int w, int h, int step) {
    for (int l = 0; l < p.nlevels; l++) {
      float s = p.invScale[l];
      int lw = cv_round((float)w * s), lh = cv_round((float)h * s);
      alloc_bordered(pyr[l], lw, lh);
      if (l == 0) {
        for (int y = 0; y < h; y++) memcpy(pyr[0].row(y), img + (ptrdiff_t)y * step, w);
      } else {
        resize_linear_u8(pyr[l - 1].origin, pyr[l - 1].w, pyr[l - 1].h, pyr[l - 1].step, pyr[l].origin, lw, lh,
                         pyr[l].step);
      }
      fill_border(pyr[l]);
    }
  }

  // ORBextractor::ComputeKeyPointsOctTree, ORBextractor.cc:1037-1184 (per level)
  void detect_level(int l, std::vector<KeyPoint>& out) {
    const Image& im = pyr[l];
    LevelDebug& D = dbg[l];
    D = LevelDebug();
    out.clear();
    const int minBX = kEdge - 3, minBY = minBX;
    const int maxBX = im.w - kEdge + 3, maxBY = im.h - kEdge + 3;
    const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
    const int nCols = (int)(width / 30.f), nRows = (int)(height / 30.f);
    if (nCols <= 0 || nRows <= 0) return;  // reference divides by zero here; we define "no keypoints"
    const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
    std::vector<FastHit> hits;
    for (int i = 0; i < nRows; i++) {
      const float iniY = (float)(minBY + i * hCell);
      float maxY = iniY + hCell + 6;
      if (iniY >= maxBY - 3) continue;
      if (maxY > maxBY) maxY = (float)maxBY;
      for (int j = 0; j < nCols; j++) {
        const float iniX = (float)(minBX + j * wCell);
        float maxX = iniX + wCell + 6;
        if (iniX >= maxBX - 6) continue;
        if (maxX > maxBX) maxX = (float)maxBX;
        const u8* sub = im.row((int)iniY) + (int)iniX;
        int sw = (int)maxX - (int)iniX, sh = (int)maxY - (int)iniY;
        D.cells++;
        fast_detect(sub, sw, sh, im.step, p.iniTh, true, hits);
        if (hits.empty()) {
          fast_detect(sub, sw, sh, im.step, p.minTh, true, hits);
          if (!hits.empty()) D.fallback_cells++;
        }
        for (const FastHit& hct : hits)
          D.cand.push_back({(float)(hct.x + j * wCell), (float)(hct.y + i * hCell), hct.score});
      }
    }
    if (D.cand.empty()) return;  // the reference would dereference an empty list; nothing to keep
    D.kept = distribute_quadtree(D.cand, minBX, maxBX, minBY, maxBY, p.perLevel[l], &D.tie_sensitive);
    const int patch = (int)(kPatch * p.scale[l]);
    for (int k : D.kept) {
      KeyPoint kp;
      kp.x = D.cand[k].x + minBX;
      kp.y = D.cand[k].y + minBY;
      kp.size = (float)patch;
      kp.response = (float)D.cand[k].score;
      kp.octave = l;
      kp.class_id = -1;
      kp.angle = ic_angle(im, kp.x, kp.y, p.umax);
      out.push_back(kp);
    }
  }

  // ORBextractor::operator(), ORBextractor.cc:1533-1649
  int run(const u8* img, int w, int h, int step) {
    kps.clear();
    desc.clear();
    if (!img || w <= 0 || h <= 0) return 0;
    compute_pyramid(img, w, h, step);
    std::vector<std::vector<KeyPoint>> all(p.nlevels);
    for (int l = 0; l < p.nlevels; l++) detect_level(l, all[l]);
    for (int l = 0; l < p.nlevels; l++) {
      blur[l].w = blur[l].h = 0;
      if (all[l].empty()) continue;
      const Image& im = pyr[l];
      Image& B = blur[l];
      B.w = im.w; B.h = im.h; B.step = im.w;
      B.buf.assign((size_t)im.w * im.h, 0);
      B.origin = B.buf.data();
      gauss7_u8(im.origin, im.w, im.h, im.step, B.origin, B.step);
      // the blurred clone has no border memory of its own in the reference either; samples
      // stay inside because keypoints are >= 19 px from the edge and the pattern reaches 18.
      size_t off = desc.size();
      desc.resize(off + all[l].size() * 32);
      for (size_t i = 0; i < all[l].size(); i++)
        brief_descriptor(B.origin, B.step, all[l][i].x, all[l][i].y, all[l][i].angle, &desc[off + i * 32]);
      if (l != 0) {
        float sc = p.scale[l];
        for (KeyPoint& kp : all[l]) { kp.x *= sc; kp.y *= sc; }
      }
      kps.insert(kps.end(), all[l].begin(), all[l].end());
    }
    return (int)kps.size();
  }
};

// ---------------------------------------------------------------- matcher
// ORBmatcher::DescriptorDistance, ORBmatcher.cc:2083-2103 (SWAR popcount on 8 x 32 bit)
inline int hamming256(const u8* a, const u8* b) {
  uint32_t pa[8], pb[8];
  memcpy(pa, a, 32);
  memcpy(pb, b, 32);
  int dist = 0;
  for (int i = 0; i < 8; i++) {
    uint32_t v = pa[i] ^ pb[i];
    v = v - ((v >> 1) & 0x55555555);
    v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
    dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
  }
  return dist;
}

inline int hamming256_popcnt(const u8* a, const u8* b) {
  uint64_t pa[4], pb[4];
  memcpy(pa, a, 32);
  memcpy(pb, b, 32);
  return __builtin_popcountll(pa[0] ^ pb[0]) + __builtin_popcountll(pa[1] ^ pb[1]) +
         __builtin_popcountll(pa[2] ^ pb[2]) + __builtin_popcountll(pa[3] ^ pb[3]);
}

// ORBmatcher::ComputeThreeMaxima, ORBmatcher.cc:2035-2077
void three_maxima(const int* histo, int L, int& i1, int& i2, int& i3) {
  int m1 = 0, m2 = 0, m3 = 0;
  i1 = i2 = i3 = -1;
  for (int i = 0; i < L; i++) {
    int s = histo[i];
    if (s > m1) { m3 = m2; m2 = m1; m1 = s; i3 = i2; i2 = i1; i1 = i; }
    else if (s > m2) { m3 = m2; m2 = s; i3 = i2; i2 = i; }
    else if (s > m3) { m3 = s; i3 = i; }
  }
  if (m2 < 0.1f * (float)m1) { i2 = -1; i3 = -1; }
  else if (m3 < 0.1f * (float)m1) { i3 = -1; }
}

struct FrameView {  // the slice of Frame the matcher reads
  int n;
  const float* xy;      // n x 2 (mvKeysUn[i].pt)
  const int* octave;    // n
  const float* angle;   // n
  const u8* desc;       // n x 32
};

struct Grid {  // Frame::AssignFeaturesToGrid / PosInGrid, Frame.cc:399-423, 682-698
  float minX, minY, invW, invH;
  std::vector<int> cell[64][48];
  Grid(const FrameView& f, float mnMinX, float mnMaxX, float mnMinY, float mnMaxY) {
    minX = mnMinX; minY = mnMinY;
    invW = 64.f / (mnMaxX - mnMinX);
    invH = 48.f / (mnMaxY - mnMinY);
    for (int i = 0; i < f.n; i++) {
      int gx = (int)std::round((f.xy[2 * i] - minX) * invW);
      int gy = (int)std::round((f.xy[2 * i + 1] - minY) * invH);
      if (gx < 0 || gx >= 64 || gy < 0 || gy >= 48) continue;
      cell[gx][gy].push_back(i);
    }
  }
  // Frame::GetFeaturesInArea, Frame.cc:590-670 (this fork: circular window)
  void query(const FrameView& f, float x, float y, float r, int minLevel, int maxLevel, std::vector<int>& out) const {
    out.clear();
    int cx0 = std::max(0, (int)std::floor((x - minX - r) * invW));
    if (cx0 >= 64) return;
    int cx1 = std::min(63, (int)std::ceil((x - minX + r) * invW));
    if (cx1 < 0) return;
    int cy0 = std::max(0, (int)std::floor((y - minY - r) * invH));
    if (cy0 >= 48) return;
    int cy1 = std::min(47, (int)std::ceil((y - minY + r) * invH));
    if (cy1 < 0) return;
    const bool check = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = cx0; ix <= cx1; ix++)
      for (int iy = cy0; iy <= cy1; iy++)
        for (int idx : cell[ix][iy]) {
          if (check) {
            if (f.octave[idx] < minLevel) continue;
            if (maxLevel >= 0 && f.octave[idx] > maxLevel) continue;
          }
          float dx = f.xy[2 * idx] - x, dy = f.xy[2 * idx + 1] - y;
          if (dx * dx + dy * dy < r * r) out.push_back(idx);
        }
  }
};

// ORBmatcher::SearchForInitialization, ORBmatcher.cc:573-717.
// mode 0: reference-faithful (octave-0 rows, grid window of radius `window` around prev[i1]);
// mode 1: brute force (every row against every descriptor of F2, index order).
int search_for_initialization(const FrameView& F1, const FrameView& F2, const float bounds[4], float* prev,
                              int* m12, int window, float nnratio, int checkOri, int mode, int* bestOut,
                              int* secondOut) {
  const int TH_LOW = 50, HISTO = 30;  // ORBmatcher.cc:49-51
  int nmatches = 0;
  for (int i = 0; i < F1.n; i++) m12[i] = -1;
  std::vector<int> hist[HISTO];
  const float factor = HISTO / 360.0f;  // this fork (ORBmatcher.cc:585-586)
  std::vector<int> matchedDist(F2.n, INT_MAX), m21(F2.n, -1);
  Grid* grid = mode == 0 ? new Grid(F2, bounds[0], bounds[1], bounds[2], bounds[3]) : nullptr;
  std::vector<int> cands;
  if (mode == 1) { cands.resize(F2.n); for (int i = 0; i < F2.n; i++) cands[i] = i; }
  for (int i1 = 0; i1 < F1.n; i1++) {
    if (bestOut) { bestOut[i1] = INT_MAX; secondOut[i1] = INT_MAX; }
    if (mode == 0) {
      if (F1.octave[i1] > 0) continue;
      grid->query(F2, prev[2 * i1], prev[2 * i1 + 1], (float)window, 0, 0, cands);
    }
    if (cands.empty()) continue;
    const u8* d1 = F1.desc + (size_t)i1 * 32;
    int best = INT_MAX, second = INT_MAX, bestIdx = -1;
    for (int i2 : cands) {
      int dist = hamming256(d1, F2.desc + (size_t)i2 * 32);
      if (matchedDist[i2] <= dist) continue;
      if (dist < best) { second = best; best = dist; bestIdx = i2; }
      else if (dist < second) second = dist;
    }
    if (bestOut) { bestOut[i1] = best; secondOut[i1] = second; }
    if (best <= TH_LOW && best < (float)second * nnratio) {
      if (m21[bestIdx] >= 0) { m12[m21[bestIdx]] = -1; nmatches--; }
      m12[i1] = bestIdx;
      m21[bestIdx] = i1;
      matchedDist[bestIdx] = best;
      nmatches++;
      if (checkOri) {
        float rot = F1.angle[i1] - F2.angle[bestIdx];
        if (rot < 0.0) rot += 360.0f;
        int bin = (int)std::round(rot * factor);
        if (bin == HISTO) bin = 0;
        if (bin >= 0 && bin < HISTO) hist[bin].push_back(i1);
      }
    }
  }
  if (checkOri) {
    int cnt[HISTO], a, b, c;
    for (int i = 0; i < HISTO; i++) cnt[i] = (int)hist[i].size();
    three_maxima(cnt, HISTO, a, b, c);
    for (int i = 0; i < HISTO; i++) {
      if (i == a || i == b || i == c) continue;
      for (int idx : hist[i])
        if (m12[idx] >= 0) { m12[idx] = -1; nmatches--; }
    }
  }
  if (prev)
    for (int i1 = 0; i1 < F1.n; i1++)
      if (m12[i1] >= 0) { prev[2 * i1] = F2.xy[2 * m12[i1]]; prev[2 * i1 + 1] = F2.xy[2 * m12[i1] + 1]; }
  delete grid;
  return nmatches;
}

// ---------------------------------------------------------------- stereo
// Frame::ComputeStereoMatches, Frame.cc:831-1082. Uses the pyramids the two extractors hold from
// their last run (mpORBextractorLeft/Right->mvImagePyramid). Returns the number of stereo points.
int compute_stereo_matches(const Extractor& EL, const Extractor& ER, const KeyPoint* kpL, int N, const u8* descL,
                           const KeyPoint* kpR, int Nr, const u8* descR, float mbf, float mb, float* uRight,
                           float* depth) {
  const int TH_HIGH = 100, TH_LOW = 50;
  for (int i = 0; i < N; i++) { uRight[i] = -1.0f; depth[i] = -1.0f; }
  const int thOrbDist = (TH_HIGH + TH_LOW) / 2;
  const int nRows = EL.pyr[0].h;
  std::vector<std::vector<int>> rowIdx(nRows);
  for (int iR = 0; iR < Nr; iR++) {
    const float kpY = kpR[iR].y;
    const float r = 2.0f * EL.p.scale[kpR[iR].octave];
    const int maxr = (int)std::ceil(kpY + r), minr = (int)std::floor(kpY - r);
    for (int yi = minr; yi <= maxr; yi++)
      if (yi >= 0 && yi < nRows) rowIdx[yi].push_back(iR);   // the reference does not bounds-check
  }
  const float minZ = mb, minD = 0, maxD = mbf / minZ;
  std::vector<std::pair<int, int>> distIdx;
  for (int iL = 0; iL < N; iL++) {
    const int levelL = kpL[iL].octave;
    const float vL = kpL[iL].y, uL = kpL[iL].x;
    const int row = (int)vL;
    if (row < 0 || row >= nRows) continue;
    const std::vector<int>& cands = rowIdx[row];
    if (cands.empty()) continue;
    const float minU = uL - maxD, maxU = uL - minD;
    if (maxU < 0) continue;
    int bestDist = TH_HIGH, bestIdxR = 0;
    for (int iR : cands) {
      if (kpR[iR].octave < levelL - 1 || kpR[iR].octave > levelL + 1) continue;
      const float uR = kpR[iR].x;
      if (uR >= minU && uR <= maxU) {
        const int d = hamming256(descL + (size_t)iL * 32, descR + (size_t)iR * 32);
        if (d < bestDist) { bestDist = d; bestIdxR = iR; }
      }
    }
    if (bestDist >= thOrbDist) continue;
    const float uR0 = kpR[bestIdxR].x;
    const float sf = EL.p.invScale[levelL];
    const float scaleduL = std::round(kpL[iL].x * sf), scaledvL = std::round(kpL[iL].y * sf);
    const float scaleduR0 = std::round(uR0 * sf);
    const int w = 5, L = 5;
    const Image& IL = EL.pyr[levelL];
    const Image& IR = ER.pyr[levelL];
    const int cvL = (int)scaledvL, cuL = (int)scaleduL, cuR = (int)scaleduR0;
    const float iniu = scaleduR0 - L - w, endu = scaleduR0 + L + w + 1;
    if (iniu < 0 || endu >= IR.w) continue;
    int best = INT_MAX, bestinc = 0;
    float vd[2 * 5 + 1];
    const float cL = (float)IL.row(cvL)[cuL];
    for (int inc = -L; inc <= L; inc++) {
      const float cR = (float)IR.row(cvL)[cuR + inc];
      float dist = 0;
      for (int dy = -w; dy <= w; dy++)
        for (int dx = -w; dx <= w; dx++) {
          const float a = (float)IL.row(cvL + dy)[cuL + dx] - cL;
          const float b = (float)IR.row(cvL + dy)[cuR + inc + dx] - cR;
          dist += std::fabs(a - b);
        }
      if (dist < best) { best = (int)dist; bestinc = inc; }
      vd[L + inc] = dist;
    }
    if (bestinc == -L || bestinc == L) continue;
    const float d1 = vd[L + bestinc - 1], d2 = vd[L + bestinc], d3 = vd[L + bestinc + 1];
    const float deltaR = (d1 - d3) / (2.0f * (d1 + d3 - 2.0f * d2));
    if (deltaR < -1 || deltaR > 1) continue;
    float bestuR = EL.p.scale[levelL] * ((float)scaleduR0 + (float)bestinc + deltaR);
    float disparity = uL - bestuR;
    if (disparity >= minD && disparity < maxD) {
      if (disparity <= 0) { disparity = 0.01; bestuR = uL - 0.01; }
      depth[iL] = mbf / disparity;
      uRight[iL] = bestuR;
      distIdx.push_back(std::make_pair(best, iL));
    }
  }
  if (distIdx.empty()) return 0;   // the reference indexes an empty vector here
  std::sort(distIdx.begin(), distIdx.end());
  const float median = (float)distIdx[distIdx.size() / 2].first;
  const float thDist = 1.5f * 1.4f * median;
  int kept = (int)distIdx.size();
  for (int i = (int)distIdx.size() - 1; i >= 0; i--) {
    if (distIdx[i].first < thDist) break;
    uRight[distIdx[i].second] = -1;
    depth[distIdx[i].second] = -1;
    kept--;
  }
  return kept;
}

// ---------------------------------------------------------------- undistortion + grid
// cv::undistortPoints(src, dst, K, dist, Mat(), K) as called by Frame::UndistortKeyPoints
// (Frame.cc:724-776) and Frame::ComputeImageBounds (:779-829): OpenCV's iterative inverse of the
// (k1,k2,p1,p2,k3) model, 5 fixed iterations in double (default TermCriteria(COUNT,5,0.01)),
// re-projected with P = K. cam = {fx, fy, cx, cy, k1, k2, p1, p2, k3} (floats, like mK/mDistCoef).
void undistort_point(const float* cam, float u, float v, float& ou, float& ov) {
  const double fx = cam[0], fy = cam[1], cx = cam[2], cy = cam[3];
  const double k1 = cam[4], k2 = cam[5], p1 = cam[6], p2 = cam[7], k3 = cam[8];
  const double ifx = 1. / fx, ify = 1. / fy;
  double x = ((double)u - cx) * ifx, y = ((double)v - cy) * ify;
  const double x0 = x, y0 = y;
  for (int j = 0; j < 5; j++) {
    const double r2 = x * x + y * y;
    const double icdist = (1 + ((0 * r2 + 0) * r2 + 0) * r2) / (1 + ((k3 * r2 + k2) * r2 + k1) * r2);
    if (icdist < 0) { x = ((double)u - cx) * ifx; y = ((double)v - cy) * ify; break; }
    const double dX = 2 * p1 * x * y + p2 * (r2 + 2 * x * x) + 0 * r2 + 0 * r2 * r2;
    const double dY = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y + 0 * r2 + 0 * r2 * r2;
    x = (x0 - dX) * icdist;
    y = (y0 - dY) * icdist;
  }
  const double xx = fx * x + 0 * y + cx, yy = 0 * x + fy * y + cy, ww = 1. / (0 * x + 0 * y + 1.);
  ou = (float)(xx * ww);
  ov = (float)(yy * ww);
}

// Frame::ComputeImageBounds
void image_bounds(const float* cam, int w, int h, float* b /* minX maxX minY maxY */) {
  if (cam[4] != 0.0f) {
    float x[4], y[4];
    const float px[4] = {0.f, (float)w, 0.f, (float)w}, py[4] = {0.f, 0.f, (float)h, (float)h};
    for (int i = 0; i < 4; i++) undistort_point(cam, px[i], py[i], x[i], y[i]);
    b[0] = std::min(x[0], x[2]); b[1] = std::max(x[1], x[3]);
    b[2] = std::min(y[0], y[1]); b[3] = std::max(y[2], y[3]);
  } else {
    b[0] = 0.f; b[1] = (float)w; b[2] = 0.f; b[3] = (float)h;
  }
}

}  // namespace

// ==================================================================== C ABI (ctypes)
// ------------------------------------------------------------------------------------------
// Ordered candidate-set matchers: the map points / keyframe features are visited
// in call order, every one picks its best (and second best) candidate among the keypoints that
// are not occupied yet, and an accepted match occupies its keypoint for the later ones.
// ------------------------------------------------------------------------------------------
struct ProjQuery {   // one projected map point (32 bytes, same layout as orb_proj_query)
  float u, v;        // projection in the current frame
  float radius;      // search window, also the tolerance of the right-image coordinate
  float ur;          // projected right-image coordinate
  float angle;       // angle of the source keypoint (rotation histogram)
  int minLevel, maxLevel;
  int flags;         // bit0: usable; bit1: its map point has Observations() > 0 (occupies the keypoint)
};
enum { SEARCH_BEST = 0, SEARCH_RATIO_LEVEL = 1, SEARCH_RATIO = 2 };

struct OrderedSearch {
  const FrameView& F;
  const float* uright;          // mvuRight (may be null = monocular)
  std::vector<u8> occupied;     // F.mvpMapPoints[idx] && Observations() > 0
  int mode, th, checkOri;
  float ratio;
  int* matchOfKp;               // F.mvpMapPoints as query indices
  int* matchOfQuery;
  int nmatches = 0;
  std::vector<int> hist[30], histQ[30];

  // the candidate loop + acceptance of one query; `cands` in the reference's visiting order
  void visit(int q, const std::vector<int>& cands, const u8* d, float angle, float ur, float radius, bool occupies, bool stereoCheck) {
    int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
    for (int idx : cands) {
      if (occupied[idx]) continue;
      if (stereoCheck && uright && uright[idx] > 0) {
        const float er = std::fabs(ur - uright[idx]);
        if (er > radius) continue;
      }
      const int dist = hamming256(d, F.desc + (size_t)idx * 32);
      if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = F.octave[idx]; bestIdx = idx; }
      else if (dist < bestDist2) { bestLevel2 = F.octave[idx]; bestDist2 = dist; }
    }
    if (bestDist > th) return;
    if (mode == SEARCH_RATIO_LEVEL && bestLevel == bestLevel2 && bestDist > ratio * bestDist2) return;   // ORBmatcher.cc:158
    if (mode == SEARCH_RATIO && !((float)bestDist < ratio * (float)bestDist2)) return;                   // ORBmatcher.cc:336
    matchOfKp[bestIdx] = q;
    matchOfQuery[q] = bestIdx;
    if (occupies) occupied[bestIdx] = 1;
    nmatches++;
    if (checkOri) {
      float rot = angle - F.angle[bestIdx];
      if (rot < 0.0) rot += 360.0f;
      int bin = (int)std::round(rot * (30 / 360.0f));
      if (bin == 30) bin = 0;
      if (bin >= 0 && bin < 30) { hist[bin].push_back(bestIdx); histQ[bin].push_back(q); }
    }
  }
  int finish() {
    if (checkOri) {
      int cnt[30], a, b, c;
      for (int i = 0; i < 30; i++) cnt[i] = (int)hist[i].size();
      three_maxima(cnt, 30, a, b, c);
      for (int i = 0; i < 30; i++) {
        if (i == a || i == b || i == c) continue;
        for (size_t j = 0; j < hist[i].size(); j++) { matchOfKp[hist[i][j]] = -1; matchOfQuery[histQ[i][j]] = -1; nmatches--; }
      }
    }
    return nmatches;
  }
};

// ORBmatcher::SearchByProjection, local map (ORBmatcher.cc:72-169, mode SEARCH_RATIO_LEVEL, th = TH_HIGH, no
// orientation check) and last frame (:1710-1860, mode SEARCH_BEST), on queries prepared by the caller.
int search_by_projection(const FrameView& F, const float* uright, const float bounds[4], const u8* occupied0,
                         const ProjQuery* Q, const u8* qdesc, int nq, int mode, int th, float ratio, int checkOri,
                         int* matchOfKp, int* matchOfQuery) {
  OrderedSearch S{F, uright, std::vector<u8>(occupied0, occupied0 + F.n), mode, th, checkOri, ratio, matchOfKp, matchOfQuery};
  for (int i = 0; i < F.n; i++) matchOfKp[i] = -1;
  Grid grid(F, bounds[0], bounds[1], bounds[2], bounds[3]);
  std::vector<int> cands;
  for (int q = 0; q < nq; q++) {
    matchOfQuery[q] = -1;
    if (!(Q[q].flags & 1)) continue;
    grid.query(F, Q[q].u, Q[q].v, Q[q].radius, Q[q].minLevel, Q[q].maxLevel, cands);
    if (cands.empty()) continue;
    S.visit(q, cands, qdesc + (size_t)q * 32, Q[q].angle, Q[q].ur, Q[q].radius, (Q[q].flags & 2) != 0, true);
  }
  return S.finish();
}

// The projection part of ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono), ORBmatcher.cc:1734-1775.
// x3Dc = Rcw*x3Dw + tcw follows cv::gemm's 3x3 float special case: float products and sums in index order,
// then the addend (checked against cv2.gemm in tests/test_oracle_search.py). dir: 0 = +-1 level, 1 = forward, 2 = backward.
void project_last_frame(int n1, const float* Xw, const u8* mpFlags, const int* octave1, const float* angle1, const float* Tcw,
                        const float* cam4, const float* bounds4, float mbf, float th, const float* scaleFactors, int dir,
                        ProjQuery* out) {
  const float fx = cam4[0], fy = cam4[1], cx = cam4[2], cy = cam4[3];
  for (int i = 0; i < n1; i++) {
    ProjQuery& q = out[i];
    q = ProjQuery{0.f, 0.f, 0.f, 0.f, angle1[i], 0, -1, 0};
    if (!(mpFlags[i] & 1)) continue;
    const float* X = Xw + 3 * (size_t)i;
    float c[3];
    for (int r = 0; r < 3; r++) {
      const float t = Tcw[4 * r] * X[0] + Tcw[4 * r + 1] * X[1] + Tcw[4 * r + 2] * X[2];
      c[r] = t + Tcw[4 * r + 3];
    }
    const float xc = c[0], yc = c[1];
    const float invzc = (float)(1.0 / c[2]);
    if (invzc < 0) continue;
    const float u = fx * xc * invzc + cx;
    const float v = fy * yc * invzc + cy;
    if (u < bounds4[0] || u > bounds4[1]) continue;
    if (v < bounds4[2] || v > bounds4[3]) continue;
    const int oct = octave1[i];
    q.u = u; q.v = v;
    q.radius = th * scaleFactors[oct];
    q.ur = u - mbf * invzc;
    if (dir == 1) { q.minLevel = oct; q.maxLevel = -1; }
    else if (dir == 2) { q.minLevel = 0; q.maxLevel = oct; }
    else { q.minLevel = oct - 1; q.maxLevel = oct + 1; }
    q.flags = 1 | (mpFlags[i] & 2);
  }
}

// ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...), ORBmatcher.cc:247-420. The DBoW2 FeatureVectors are given as
// the node id of every feature (-1 = none): FeatureVector = std::map<node, indices in feature order>.
// usable1[i]: KF feature i has a good map point. Output: matchOfKp = vpMapPointMatches as KF feature indices.
// unusable2 (may be null): frame-side features that are no candidates — SearchByBoW(KeyFrame*, KeyFrame*), :729-880, skips
// keyframe-2 features without a good map point (and tests bestDist1 < TH_LOW: pass th = TH_LOW - 1).
int search_by_bow(const FrameView& KF, const int* node1, const u8* usable1, const FrameView& F, const int* node2, const u8* unusable2,
                  int th, float ratio, int checkOri, int* matchOfKp, int* matchOfQuery) {
  OrderedSearch S{F, nullptr, unusable2 ? std::vector<u8>(unusable2, unusable2 + F.n) : std::vector<u8>(F.n, 0), SEARCH_RATIO, th,
                  checkOri, ratio, matchOfKp, matchOfQuery};
  for (int i = 0; i < F.n; i++) matchOfKp[i] = -1;
  for (int i = 0; i < KF.n; i++) matchOfQuery[i] = -1;
  std::map<int, std::vector<int>> fv1, fv2;
  for (int i = 0; i < KF.n; i++) if (node1[i] >= 0) fv1[node1[i]].push_back(i);
  for (int i = 0; i < F.n; i++) if (node2[i] >= 0) fv2[node2[i]].push_back(i);
  auto it1 = fv1.begin(), it2 = fv2.begin();
  while (it1 != fv1.end() && it2 != fv2.end()) {
    if (it1->first == it2->first) {
      for (int i1 : it1->second) {
        if (!usable1[i1]) continue;
        S.visit(i1, it2->second, KF.desc + (size_t)i1 * 32, KF.angle[i1], 0.f, 0.f, true, false);
      }
      ++it1; ++it2;
    } else if (it1->first < it2->first) it1 = fv1.lower_bound(it2->first);
    else it2 = fv2.lower_bound(it1->first);
  }
  return S.finish();
}

// ------------------------------------------------------------------------------------------
// Input stage and MapPoint::ComputeDistinctiveDescriptors
// ------------------------------------------------------------------------------------------
// cv::cvtColor(..., CV_RGB2GRAY / CV_BGR2GRAY / CV_RGBA2GRAY / CV_BGRA2GRAY) on 8-bit images as Tracking.cc:250-276,
// 310-324, 369-383 call it. OpenCV 4.x: 15-bit coefficients RY15 = 9798, GY15 = 19235, BY15 = 3735, rounded
// (checked bit-exactly against cv2.cvtColor in tests/test_oracle_input.py).
void cvt_gray(const u8* src, int w, int h, size_t step, int channels, int blueFirst, u8* dst, size_t dstep) {
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      const u8* p = src + (size_t)y * step + (size_t)x * channels;
      const int r = blueFirst ? p[2] : p[0], g = p[1], b = blueFirst ? p[0] : p[2];
      dst[(size_t)y * dstep + x] = (u8)((r * 9798 + g * 19235 + b * 3735 + (1 << 14)) >> 15);
    }
}

// cv::remap(src, dst, mapx, mapy, INTER_LINEAR) with CV_32FC1 maps and the default BORDER_CONSTANT(0), as
// Examples/Stereo/stereo_euroc.cc:181-188 calls it. OpenCV's fixed point: coordinates rounded to 1/32 px, 2x2
// weights from the 32x32 table of initInterTab2D (shorts scaled by 2^15, repaired to sum 2^15), result
// (sum + 2^14) >> 15 (checked bit-exactly against cv2.remap).
struct RemapTab {
  short w[32][32][4];
  RemapTab() {
    float t1[32][2];
    for (int i = 0; i < 32; i++) { const float x = (float)i * (1.f / 32); t1[i][0] = 1.f - x; t1[i][1] = x; }
    for (int i = 0; i < 32; i++)
      for (int j = 0; j < 32; j++) {
        short* it = w[i][j];
        int isum = 0;
        for (int k1 = 0; k1 < 2; k1++)
          for (int k2 = 0; k2 < 2; k2++) {
            const float v = t1[i][k1] * t1[j][k2];
            long r = lrintf(v * 32768.f);
            r = std::min(32767L, std::max(-32768L, r));   // saturate_cast<short>
            it[k1 * 2 + k2] = (short)r;
            isum += (int)r;
          }
        if (isum != 32768) {   // initInterTab2D's repair with ksize = 2: only element [3] is inspected
          const int diff = isum - 32768;
          it[3] = (short)(it[3] - diff);
        }
      }
  }
};
void remap_linear(const u8* src, int sw, int sh, size_t sstep, const float* mapx, const float* mapy, int dw, int dh, u8* dst,
                  size_t dstep) {
  static const RemapTab tab;
  for (int y = 0; y < dh; y++)
    for (int x = 0; x < dw; x++) {
      const int sx = (int)lrintf(mapx[(size_t)y * dw + x] * 32.f), sy = (int)lrintf(mapy[(size_t)y * dw + x] * 32.f);
      const int X = std::min(32767, std::max(-32768, sx >> 5)), Y = std::min(32767, std::max(-32768, sy >> 5));
      const short* wt = tab.w[sy & 31][sx & 31];
      int acc = 0;
      for (int k1 = 0; k1 < 2; k1++)
        for (int k2 = 0; k2 < 2; k2++) {
          const int yy = Y + k1, xx = X + k2;
          const int p = (yy >= 0 && yy < sh && xx >= 0 && xx < sw) ? src[(size_t)yy * sstep + xx] : 0;
          acc += p * wt[k1 * 2 + k2];
        }
      dst[(size_t)y * dstep + x] = (u8)std::min(255, std::max(0, (acc + (1 << 14)) >> 15));
    }
}

// MapPoint::ComputeDistinctiveDescriptors, MapPoint.cc:365-448: the observation whose median distance to all
// observations (itself included) is smallest; first wins.
int distinctive_descriptor(const u8* desc, int n) {
  int bestMedian = INT_MAX, bestIdx = 0;
  std::vector<int> d(n);
  for (int i = 0; i < n; i++) {
    for (int j = 0; j < n; j++) d[j] = hamming256(desc + (size_t)i * 32, desc + (size_t)j * 32);
    std::sort(d.begin(), d.end());
    const int median = d[(int)(0.5 * (n - 1))];
    if (median < bestMedian) { bestMedian = median; bestIdx = i; }
  }
  return bestIdx;
}

// ORBmatcher::SearchForTriangulation, ORBmatcher.cc:884-1100 (with CheckDistEpipolarLine :205-227). FeatureVectors
// as per-feature node ids; ur1 / ur2 = mvuRight (may be null = monocular); hasMp = "GetMapPoint(idx) != NULL".
struct TriPair { float F12[9]; float ex, ey; int onlyStereo; };
static bool check_dist_epipolar_line(float x1, float y1, float x2, float y2, int oct2, const float* F12, const float* sigma2) {
  const float a = x1 * F12[0] + y1 * F12[3] + F12[6];
  const float b = x1 * F12[1] + y1 * F12[4] + F12[7];
  const float c = x1 * F12[2] + y1 * F12[5] + F12[8];
  const float num = a * x2 + b * y2 + c;
  const float den = a * a + b * b;
  if (den == 0) return false;
  const float dsqr = num * num / den;
  return dsqr < 3.84 * sigma2[oct2];
}
int search_for_triangulation(const FrameView& K1, const int* node1, const u8* hasMp1, const float* ur1, const FrameView& K2,
                             const int* node2, const u8* hasMp2, const float* ur2, const TriPair& P, const float* sf,
                             const float* sigma2, int checkOri, int* matches12) {
  const int TH_LOW = 50;
  int nmatches = 0;
  std::vector<u8> matched2(K2.n, 0);
  for (int i = 0; i < K1.n; i++) matches12[i] = -1;
  std::vector<int> hist[30];
  std::map<int, std::vector<int>> fv1, fv2;
  for (int i = 0; i < K1.n; i++) if (node1[i] >= 0) fv1[node1[i]].push_back(i);
  for (int i = 0; i < K2.n; i++) if (node2[i] >= 0) fv2[node2[i]].push_back(i);
  auto it1 = fv1.begin(), it2 = fv2.begin();
  while (it1 != fv1.end() && it2 != fv2.end()) {
    if (it1->first == it2->first) {
      for (int idx1 : it1->second) {
        if (hasMp1[idx1]) continue;
        const bool stereo1 = ur1 && ur1[idx1] >= 0;
        if (P.onlyStereo && !stereo1) continue;
        int bestDist = TH_LOW, bestIdx2 = -1;
        for (int idx2 : it2->second) {
          if (matched2[idx2] || hasMp2[idx2]) continue;
          const bool stereo2 = ur2 && ur2[idx2] >= 0;
          if (P.onlyStereo && !stereo2) continue;
          const int dist = hamming256(K1.desc + (size_t)idx1 * 32, K2.desc + (size_t)idx2 * 32);
          if (dist > TH_LOW || dist > bestDist) continue;
          if (!stereo1 && !stereo2) {
            const float distex = P.ex - K2.xy[2 * idx2], distey = P.ey - K2.xy[2 * idx2 + 1];
            if (distex * distex + distey * distey < 100 * sf[K2.octave[idx2]]) continue;
          }
          if (check_dist_epipolar_line(K1.xy[2 * idx1], K1.xy[2 * idx1 + 1], K2.xy[2 * idx2], K2.xy[2 * idx2 + 1], K2.octave[idx2],
                                       P.F12, sigma2)) {
            bestIdx2 = idx2;
            bestDist = dist;
          }
        }
        if (bestIdx2 >= 0) {
          matches12[idx1] = bestIdx2;
          matched2[bestIdx2] = 1;
          nmatches++;
          if (checkOri) {
            float rot = K1.angle[idx1] - K2.angle[bestIdx2];
            if (rot < 0.0) rot += 360.0f;
            int bin = (int)std::round(rot * (30 / 360.0f));
            if (bin == 30) bin = 0;
            if (bin >= 0 && bin < 30) hist[bin].push_back(idx1);
          }
        }
      }
      ++it1; ++it2;
    } else if (it1->first < it2->first) it1 = fv1.lower_bound(it2->first);
    else it2 = fv2.lower_bound(it1->first);
  }
  if (checkOri) {
    int cnt[30], a, b, c;
    for (int i = 0; i < 30; i++) cnt[i] = (int)hist[i].size();
    three_maxima(cnt, 30, a, b, c);
    for (int i = 0; i < 30; i++) {
      if (i == a || i == b || i == c) continue;
      for (int idx1 : hist[i]) { matches12[idx1] = -1; nmatches--; }
    }
  }
  return nmatches;
}

extern "C" {

void* orc_create(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh) {
  return new Extractor(nfeatures, scaleFactor, nlevels, iniTh, minTh);
}
void orc_destroy(void* h) { delete (Extractor*)h; }

void orc_tables(void* h, float* scale, float* invScale, float* sigma2, float* invSigma2, int* perLevel, int* umax) {
  Extractor* e = (Extractor*)h;
  for (int l = 0; l < e->p.nlevels; l++) {
    scale[l] = e->p.scale[l]; invScale[l] = e->p.invScale[l];
    sigma2[l] = e->p.sigma2[l]; invSigma2[l] = e->p.invSigma2[l];
    perLevel[l] = e->p.perLevel[l];
  }
  for (int v = 0; v <= kHalfPatch; v++) umax[v] = e->p.umax[v];
}

// returns N; keypoints (28-byte records) and N x 32 descriptors are copied up to cap
int orc_extract(void* h, const u8* img, int w, int hgt, int step, void* kps, u8* desc, int cap) {
  Extractor* e = (Extractor*)h;
  int n = e->run(img, w, hgt, step);
  int m = std::min(n, cap);
  if (kps && m > 0) memcpy(kps, e->kps.data(), (size_t)m * sizeof(KeyPoint));
  if (desc && m > 0) memcpy(desc, e->desc.data(), (size_t)m * 32);
  return n;
}

// level geometry + pointers into the last run's stage outputs
void orc_level_info(void* h, int l, int* w, int* hgt, int* step) {
  Extractor* e = (Extractor*)h;
  *w = e->pyr[l].w; *hgt = e->pyr[l].h; *step = e->pyr[l].step;
}
// copies the BORDERED level ((w+38) x (h+38), tightly packed)
void orc_level_copy(void* h, int l, u8* dst) {
  Extractor* e = (Extractor*)h;
  memcpy(dst, e->pyr[l].buf.data(), e->pyr[l].buf.size());
}
int orc_blur_copy(void* h, int l, u8* dst) {  // w x h tightly packed; returns 0 if level had no keypoints
  Extractor* e = (Extractor*)h;
  if (e->blur[l].w == 0) return 0;
  memcpy(dst, e->blur[l].buf.data(), e->blur[l].buf.size());
  return 1;
}
int orc_level_candidates(void* h, int l, int* xs, int* ys, int* score, int cap) {  // image coords of the level
  Extractor* e = (Extractor*)h;
  const LevelDebug& D = e->dbg[l];
  int n = (int)D.cand.size();
  for (int i = 0; i < std::min(n, cap); i++) {
    xs[i] = (int)D.cand[i].x + (kEdge - 3); ys[i] = (int)D.cand[i].y + (kEdge - 3); score[i] = D.cand[i].score;
  }
  return n;
}
int orc_level_kept(void* h, int l, int* idx, int cap) {
  Extractor* e = (Extractor*)h;
  const LevelDebug& D = e->dbg[l];
  int n = (int)D.kept.size();
  for (int i = 0; i < std::min(n, cap); i++) idx[i] = D.kept[i];
  return n;
}
void orc_level_stats(void* h, int l, int* cells, int* fallback, int* tie_sensitive) {
  Extractor* e = (Extractor*)h;
  *cells = e->dbg[l].cells; *fallback = e->dbg[l].fallback_cells; *tie_sensitive = e->dbg[l].tie_sensitive;
}

// ---- primitives, exposed so that tests can pin each one against cv2
void orc_resize_linear(const u8* src, int sw, int sh, int sstep, u8* dst, int dw, int dh, int dstep) {
  resize_linear_u8(src, sw, sh, sstep, dst, dw, dh, dstep);
}
void orc_border101(const u8* src, int w, int h, int sstep, u8* dst /* (w+38)x(h+38) packed */) {
  Image im;
  alloc_bordered(im, w, h);
  for (int y = 0; y < h; y++) memcpy(im.row(y), src + (ptrdiff_t)y * sstep, w);
  fill_border(im);
  memcpy(dst, im.buf.data(), im.buf.size());
}
int orc_fast(const u8* img, int w, int h, int step, int threshold, int nms, int* xs, int* ys, int* score, int cap) {
  std::vector<FastHit> hits;
  fast_detect(img, w, h, step, threshold, nms != 0, hits);
  for (int i = 0; i < std::min((int)hits.size(), cap); i++) { xs[i] = hits[i].x; ys[i] = hits[i].y; score[i] = hits[i].score; }
  return (int)hits.size();
}
void orc_fast_score_map(const u8* img, int w, int h, int step, int* out /* w*h, 0 in the 3-px frame */) {
  for (int i = 0; i < w * h; i++) out[i] = 0;
  for (int y = 3; y < h - 3; y++)
    for (int x = 3; x < w - 3; x++) out[y * w + x] = fast_score(img + (ptrdiff_t)y * step + x, step);
}
void orc_gauss7(const u8* src, int w, int h, int sstep, u8* dst, int dstep) { gauss7_u8(src, w, h, sstep, dst, dstep); }
float orc_fast_atan2(float y, float x) { return fast_atan2_deg(y, x); }
void orc_fast_atan2_many(const float* y, const float* x, float* out, int n) {
  for (int i = 0; i < n; i++) out[i] = fast_atan2_deg(y[i], x[i]);
}
float orc_ic_angle(const u8* img, int step, int px, int py, void* h) {
  Image im; im.origin = (u8*)img; im.step = step;
  return ic_angle(im, (float)px, (float)py, ((Extractor*)h)->p.umax);
}
void orc_brief(const u8* img, int step, float px, float py, float angle, u8* desc) {
  brief_descriptor(img, step, px, py, angle, desc);
}
// quadtree alone: candidates in relative coords; returns kept count, indices in list order
int orc_quadtree(const float* xs, const float* ys, const int* score, int n, int minX, int maxX, int minY, int maxY,
                 int N, int* kept, int cap, int* tie_sensitive) {
  std::vector<Cand> c(n);
  for (int i = 0; i < n; i++) c[i] = {xs[i], ys[i], score[i]};
  if (n == 0) return 0;
  std::vector<int> r = distribute_quadtree(c, minX, maxX, minY, maxY, N, tie_sensitive);
  for (int i = 0; i < std::min((int)r.size(), cap); i++) kept[i] = r[i];
  return (int)r.size();
}

int orc_stereo_matches(void* hL, void* hR, const void* kpL, int nL, const u8* descL, const void* kpR, int nR,
                       const u8* descR, float mbf, float mb, float* uRight, float* depth) {
  return compute_stereo_matches(*(Extractor*)hL, *(Extractor*)hR, (const KeyPoint*)kpL, nL, descL, (const KeyPoint*)kpR, nR,
                                descR, mbf, mb, uRight, depth);
}

// Frame::UndistortKeyPoints: only pt changes; identity when k1 == 0
void orc_undistort_keypoints(const void* kin, int n, const float* cam9, void* kout) {
  const KeyPoint* a = (const KeyPoint*)kin;
  KeyPoint* o = (KeyPoint*)kout;
  for (int i = 0; i < n; i++) {
    o[i] = a[i];
    if (cam9[4] != 0.0f) undistort_point(cam9, a[i].x, a[i].y, o[i].x, o[i].y);
  }
}
void orc_image_bounds(const float* cam9, int w, int h, float* bounds4) { image_bounds(cam9, w, h, bounds4); }

// Frame::AssignFeaturesToGrid (Frame.cc:399-423) as CSR: cell = ix*48 + iy, items in insertion order
void orc_assign_grid(const void* kps, int n, const float* bounds4, int* cellStart /*3073*/, int* items /*n*/) {
  const KeyPoint* k = (const KeyPoint*)kps;
  std::vector<float> xy(2 * (size_t)n);
  std::vector<int> oct(n, 0);
  for (int i = 0; i < n; i++) { xy[2 * i] = k[i].x; xy[2 * i + 1] = k[i].y; }
  FrameView f{n, xy.data(), oct.data(), nullptr, nullptr};
  Grid g(f, bounds4[0], bounds4[1], bounds4[2], bounds4[3]);
  int pos = 0;
  for (int ix = 0; ix < 64; ix++)
    for (int iy = 0; iy < 48; iy++) {
      cellStart[ix * 48 + iy] = pos;
      for (int idx : g.cell[ix][iy]) items[pos++] = idx;
    }
  cellStart[64 * 48] = pos;
}

// Frame::GetFeaturesInArea (Frame.cc:590-670): indices in the reference's order
int orc_features_in_area(const void* kps, int n, const float* bounds4, float x, float y, float r, int minLevel,
                         int maxLevel, int* out, int cap) {
  const KeyPoint* k = (const KeyPoint*)kps;
  std::vector<float> xy(2 * (size_t)n);
  std::vector<int> oct(n);
  for (int i = 0; i < n; i++) { xy[2 * i] = k[i].x; xy[2 * i + 1] = k[i].y; oct[i] = k[i].octave; }
  FrameView f{n, xy.data(), oct.data(), nullptr, nullptr};
  Grid g(f, bounds4[0], bounds4[1], bounds4[2], bounds4[3]);
  std::vector<int> res;
  g.query(f, x, y, r, minLevel, maxLevel, res);
  for (int i = 0; i < std::min((int)res.size(), cap); i++) out[i] = res[i];
  return (int)res.size();
}

// ---- matcher
int orc_hamming(const u8* a, const u8* b) { return hamming256(a, b); }
void orc_hamming_matrix(const u8* a, int na, const u8* b, int nb, int* out) {
  for (int i = 0; i < na; i++)
    for (int j = 0; j < nb; j++) out[(size_t)i * nb + j] = hamming256(a + (size_t)i * 32, b + (size_t)j * 32);
}
void orc_three_maxima(const int* histo, int L, int* out3) { three_maxima(histo, L, out3[0], out3[1], out3[2]); }

int orc_search_for_initialization(int n1, const float* xy1, const int* oct1, const float* ang1, const u8* desc1,
                                  int n2, const float* xy2, const int* oct2, const float* ang2, const u8* desc2,
                                  const float* bounds4, float* prevMatched, int* matches12, int windowSize,
                                  float nnratio, int checkOri, int mode, int* bestOut, int* secondOut) {
  FrameView F1{n1, xy1, oct1, ang1, desc1}, F2{n2, xy2, oct2, ang2, desc2};
  return search_for_initialization(F1, F2, bounds4, prevMatched, matches12, windowSize, nnratio, checkOri, mode,
                                   bestOut, secondOut);
}

// ---- ordered candidate-set matchers
static FrameView view_of(const void* kps, int n, const u8* desc, std::vector<float>& xy, std::vector<int>& oct, std::vector<float>& ang) {
  const KeyPoint* k = (const KeyPoint*)kps;
  xy.resize(2 * (size_t)n); oct.resize(n); ang.resize(n);
  for (int i = 0; i < n; i++) { xy[2 * i] = k[i].x; xy[2 * i + 1] = k[i].y; oct[i] = k[i].octave; ang[i] = k[i].angle; }
  return FrameView{n, xy.data(), oct.data(), ang.data(), desc};
}
int orc_search_by_projection(const void* kpsUn, int n, const u8* desc, const float* uright, const float* bounds4,
                             const u8* occupied0, const void* queries, const u8* qdesc, int nq, int mode, int th,
                             float ratio, int checkOri, int* matchOfKp, int* matchOfQuery) {
  std::vector<float> xy, ang; std::vector<int> oct;
  FrameView F = view_of(kpsUn, n, desc, xy, oct, ang);
  return search_by_projection(F, uright, bounds4, occupied0, (const ProjQuery*)queries, qdesc, nq, mode, th, ratio, checkOri,
                              matchOfKp, matchOfQuery);
}
void orc_project_last_frame(int n1, const float* Xw, const u8* mpFlags, const void* kps1, const float* Tcw, const float* cam4,
                            const float* bounds4, float mbf, float th, const float* scaleFactors, int dir, void* out) {
  const KeyPoint* k = (const KeyPoint*)kps1;
  std::vector<int> oct(n1); std::vector<float> ang(n1);
  for (int i = 0; i < n1; i++) { oct[i] = k[i].octave; ang[i] = k[i].angle; }
  project_last_frame(n1, Xw, mpFlags, oct.data(), ang.data(), Tcw, cam4, bounds4, mbf, th, scaleFactors, dir, (ProjQuery*)out);
}
int orc_search_by_bow(const void* kps1, int n1, const u8* desc1, const int* node1, const u8* usable1, const void* kps2, int n2,
                      const u8* desc2, const int* node2, const u8* unusable2, int th, float ratio, int checkOri, int* matchOfKp,
                      int* matchOfQuery) {
  std::vector<float> xy1, ang1, xy2, ang2; std::vector<int> oct1, oct2;
  FrameView A = view_of(kps1, n1, desc1, xy1, oct1, ang1), B = view_of(kps2, n2, desc2, xy2, oct2, ang2);
  return search_by_bow(A, node1, usable1, B, node2, unusable2, th, ratio, checkOri, matchOfKp, matchOfQuery);
}

int orc_search_for_triangulation(const void* kps1, int n1, const u8* desc1, const int* node1, const u8* hasMp1, const float* ur1,
                                 const void* kps2, int n2, const u8* desc2, const int* node2, const u8* hasMp2, const float* ur2,
                                 const void* pair, const float* sf, const float* sigma2, int checkOri, int* matches12) {
  std::vector<float> xy1, ang1, xy2, ang2; std::vector<int> oct1, oct2;
  FrameView A = view_of(kps1, n1, desc1, xy1, oct1, ang1), B = view_of(kps2, n2, desc2, xy2, oct2, ang2);
  return search_for_triangulation(A, node1, hasMp1, ur1, B, node2, hasMp2, ur2, *(const TriPair*)pair, sf, sigma2, checkOri, matches12);
}

// ---- input stage / map point descriptors
void orc_cvt_gray(const u8* src, int w, int h, size_t step, int channels, int blueFirst, u8* dst) {
  cvt_gray(src, w, h, step, channels, blueFirst, dst, (size_t)w);
}
void orc_remap_linear(const u8* src, int sw, int sh, size_t sstep, const float* mapx, const float* mapy, int dw, int dh, u8* dst) {
  remap_linear(src, sw, sh, sstep, mapx, mapy, dw, dh, dst, (size_t)dw);
}
void orc_distinctive_descriptors(const u8* desc, const int* offsets, int nPoints, int* bestIdx) {
  for (int p = 0; p < nPoints; p++) {
    const int n = offsets[p + 1] - offsets[p];
    bestIdx[p] = n > 0 ? distinctive_descriptor(desc + (size_t)offsets[p] * 32, n) : -1;
  }
}

// All-pairs keyframe matching count (config 5): for keyframe pair (i,j), the number of rows
// of i whose 2-NN in j passes best<=TH_LOW and best < ratio*second (no dedup, no histogram).
void orc_allpairs_counts(const u8* desc, int nKF, int nDesc, float nnratio, int rowBegin, int rowEnd, int* out) {
  for (int i = rowBegin; i < rowEnd; i++)
    for (int j = 0; j < nKF; j++) {
      int cnt = 0;
      for (int a = 0; a < nDesc; a++) {
        int best = INT_MAX, second = INT_MAX;
        const u8* da = desc + ((size_t)i * nDesc + a) * 32;
        for (int b = 0; b < nDesc; b++) {
          int d = hamming256_popcnt(da, desc + ((size_t)j * nDesc + b) * 32);
          if (d < best) { second = best; best = d; }
          else if (d < second) second = d;
        }
        if (best <= 50 && best < (float)second * nnratio) cnt++;
      }
      out[(size_t)(i - rowBegin) * nKF + j] = cnt;
    }
}

// ---- multi-threaded drivers for the CPU baseline (one frame / one pair per thread)
// returns total keypoints; images are B contiguous w*h frames
long orc_extract_batch_mt(int nfeatures, float sf, int nlevels, int ini, int mn, const u8* imgs, int B, int w, int h,
                          int nthreads, int* counts) {
  std::atomic<int> next(0);
  std::atomic<long> total(0);
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; t++)
    th.emplace_back([&]() {
      Extractor e(nfeatures, sf, nlevels, ini, mn);
      for (;;) {
        int i = next.fetch_add(1);
        if (i >= B) break;
        int n = e.run(imgs + (size_t)i * w * h, w, h, w);
        if (counts) counts[i] = n;
        total += n;
      }
    });
  for (auto& t : th) t.join();
  return total.load();
}

// brute-force 2-NN matching of P pairs (pair p: A = desc[(2p)*n..], B = desc[(2p+1)*n..]); swar=1 uses the
// reference's SWAR popcount, 0 the popcnt instruction. Returns total matches.
long orc_match_batch_mt(const u8* desc, const float* angle, int P, int n, float nnratio, int nthreads, int* nmatch) {
  std::atomic<int> next(0);
  std::atomic<long> total(0);
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; t++)
    th.emplace_back([&]() {
      std::vector<int> m12(n), oct(n, 0);
      std::vector<float> xy(2 * (size_t)n, 0.f);
      float bounds[4] = {0, 1, 0, 1};
      for (;;) {
        int p = next.fetch_add(1);
        if (p >= P) break;
        FrameView F1{n, xy.data(), oct.data(), angle + (size_t)(2 * p) * n, desc + (size_t)(2 * p) * n * 32};
        FrameView F2{n, xy.data(), oct.data(), angle + (size_t)(2 * p + 1) * n, desc + (size_t)(2 * p + 1) * n * 32};
        int nm = search_for_initialization(F1, F2, bounds, nullptr, m12.data(), 0, nnratio, 1, 1, nullptr, nullptr);
        if (nmatch) nmatch[p] = nm;
        total += nm;
      }
    });
  for (auto& t : th) t.join();
  return total.load();
}

int orc_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
