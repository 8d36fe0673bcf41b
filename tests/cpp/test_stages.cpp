// Prints the stage schedule of the chunked host pipelines for "n chunk head taper" quadruples given on the command line:
// one line per case, stage sizes separated by blanks.  Driver of tests/test_stage_schedule.py (CPU).
#include <cstdio>
#include <cstdlib>

#include "fclb_stages.h"

int main(int argc, char** argv) {
  for (int i = 1; i + 3 < argc; i += 4) {
    std::vector<size_t> begin, size;
    fclb::stageSizes(std::strtoull(argv[i], nullptr, 10), std::strtoull(argv[i + 1], nullptr, 10),
                     std::strtoull(argv[i + 2], nullptr, 10), std::strtoull(argv[i + 3], nullptr, 10), begin, size);
    size_t at = 0;
    for (size_t k = 0; k < size.size(); k++) {
      if (begin[k] != at) return 2;  // stages must tile [0, n) in order
      at += size[k];
      std::printf(k ? " %zu" : "%zu", size[k]);
    }
    std::printf("\n");
  }
  return 0;
}
