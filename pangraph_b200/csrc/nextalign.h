// K6: map_variations (banded affine alignment of a sequence against a block consensus) on the GPU.  See nextalign.cu.
#pragma once
#include "nextalign_host.h"

#include <cstdint>
#include <vector>

namespace pgmm {
namespace na {

struct Problem {
  const char *ref, *qry;  // ASCII, borrowed for the call
  int32_t rlen, qlen, mean_shift, band_width;
};
struct Stats {
  double kernel_ms = 0;
  uint64_t cells = 0, problems = 0, launches = 0;  // band cells asked for (every attempt counts), problem attempts
};

// map_variations for every problem (PG/align/map_variations.rs:39-80): band_width + extra_band_width, up to max_attempts
// attempts with the band doubled while the traceback touches the band boundary (align.rs:52-63).
void run_batch(const std::vector<Problem> &probs, int extra_band_width, int max_attempts, std::vector<Edit> &edits, Stats *stats);

}  // namespace na
}  // namespace pgmm
