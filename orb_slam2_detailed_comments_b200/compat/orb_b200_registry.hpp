// orb_b200_registry.hpp — shared by the drop-in translation units (orb_b200_extractor.cpp, orb_b200_matcher.cpp,
// orb_b200_frame.cpp). The reference's class declarations stay UNMODIFIED, so there is no member to hold the C-ABI handle:
// the handle of an ORB_SLAM2::ORBextractor is kept in a registry keyed by the object's address.
#pragma once
#include <mutex>
#include <unordered_map>
#include <vector>

#include "../../include/orb_b200.h"

namespace orb_b200_compat {

struct ExtractorEntry {
  orb_extractor* handle = nullptr;
  std::vector<orb_keypoint> kps;       // staging of operator() (an extractor is not re-entrant, like the reference's)
  std::vector<unsigned char> desc;
  std::vector<orb_level_view> views;
};

inline std::mutex& registry_mutex() { static std::mutex m; return m; }
inline std::unordered_map<const void*, ExtractorEntry*>& registry() {
  static std::unordered_map<const void*, ExtractorEntry*> r;
  return r;
}
// The reference declares `~ORBextractor(){}` inline (include/ORBextractor.h:105), so destruction cannot be observed: an
// entry lives until an extractor is constructed at the same address again (or the process ends).
inline ExtractorEntry* entry_of(const void* extractor) {
  std::lock_guard<std::mutex> lock(registry_mutex());
  auto it = registry().find(extractor);
  return it == registry().end() ? nullptr : it->second;
}

}  // namespace orb_b200_compat
