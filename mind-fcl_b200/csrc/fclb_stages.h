// Stage schedule of the chunked *_host pipelines (fclb_engine.cu, fclb_bvh.cu, fclb_collide_api.cu, fclb_scene_api.cu).
// Plain C++ (no CUDA) so that the CPU tests can compile it: tests/test_stage_schedule.py.
#pragma once
#include <algorithm>
#include <cstddef>
#include <vector>

namespace fclb {

// Stage sizes of a chunked host pipeline.  taper > 0 (copy-bound calls: the call ends one stage's compute + copy-out after
// the last upload): `chunk` queries while two or more stages are left, then halving down to `taper`.  head > 0
// (compute-bound calls: nothing runs before the first upload is in): the first stage has `head` queries and the stages
// double up to `chunk`.
inline void stageSizes(size_t n, size_t chunk, size_t head, size_t taper, std::vector<size_t>& begin, std::vector<size_t>& size) {
  begin.clear();
  size.clear();
  if (chunk == 0) chunk = 1;
  size_t ramp = head ? std::min(head, chunk) : chunk;
  for (size_t b = 0; b < n;) {
    const size_t rem = n - b;
    size_t m = ramp;
    if (ramp < chunk) {
      ramp = std::min(chunk, 2 * ramp);
      if (rem < m + m / 2) m = rem;  // no sliver at the end
    } else if (taper && rem < 2 * chunk) {
      m = std::max(taper, (rem / 2 + 4095) / 4096 * 4096);
      if (rem < m + taper) m = rem;
    }
    m = std::min(m, rem);
    begin.push_back(b);
    size.push_back(m);
    b += m;
  }
}

}  // namespace fclb
