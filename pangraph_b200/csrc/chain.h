// Host stage: RMQ chaining of sorted anchors (reference: minimap2/lchain.c:250-368).
#pragma once
#include <cstdint>
#include <vector>

#include "flag_sort.h"

namespace pgmm {

struct ChainParams {
  int max_dist;        // opt->max_gap
  int max_dist_inner;  // opt->rmq_inner_dist
  int bw;              // opt->bw
  int max_chn_skip;    // opt->max_chain_skip
  int cap_rmq_size;    // opt->rmq_size_cap
  int min_cnt;         // opt->min_cnt
  int min_sc;          // opt->min_chain_score
  float pen_gap;       // chain_gap_scale * 0.01 * k   (map.c:273)
  float pen_skip;      // chain_skip_scale * 0.01 * k
};

// The score fill of mg_lchain_rmq restarts from an empty tree wherever the target/strand changes or two neighbouring
// anchors are more than max_dist apart on the target (lchain.c:295-312 evicts everything): the anchor array falls into
// independent SEGMENTS.  The device fills them (one warp each, chain_fill.cu); the host version below is the arbiter
// for the segments the device hands back (equal priorities inside an RMQ window -- the reference then follows the
// shape of its balanced tree -- or windows beyond the device limits).
struct ChainSeg {
  int64_t start, end;
};
void chain_find_segments(const ChainParams &cp, const U128 *a, int64_t n, std::vector<ChainSeg> &segs);

// Scores f, predecessors p (index into a, -1 = none) and peak scores v of the anchors [seg_start, seg_end) with the
// reference's balanced range-minimum tree; t is scratch of n int32 that must not hold a value >= 1 the caller cares about.
void chain_fill_host(const ChainParams &cp, const U128 *a, int64_t n, int64_t seg_start, int64_t seg_end, int32_t *f, int32_t *p,
                     int32_t *v, int32_t *t);

// Backtracking and compaction (lchain.c:27-111): a = sorted anchors in, anchors of the kept chains out;
// u[i] = score<<32 | n_anchors.  v and t (n int32 each) are overwritten.
void chain_backtrack(const ChainParams &cp, std::vector<U128> &a, const int32_t *f, const int32_t *p, int32_t *v, int32_t *t,
                     std::vector<uint64_t> &u);

// One query's sorted anchors on their way through the device fill (chain_fill.cu).  After the backend's chain_fill,
// f/p/v point at n results each (p: index into a, -1 = none) that stay valid until the backend's next chain_fill.
struct ChainFillJob {
  const U128 *a = nullptr;
  int64_t n = 0;
  std::vector<ChainSeg> segs;    // from chain_find_segments
  std::vector<uint8_t> redo;     // per segment: 0 = filled by the device, else the reason the host must fill it
  int32_t *f = nullptr, *p = nullptr, *v = nullptr;
};

struct ChainFillStats {
  uint64_t anchors = 0, segments = 0, redo_segments = 0, redo_anchors = 0;
  int launches = 0;
  double kernel_ms = 0;
  uint64_t iterations = 0, batches = 0;  // K4p: fixed-point iterations summed over the fills (batches)
};

// segments + host fill + backtrack in one call (stage tests, CPU-only seam)
void chain_rmq(const ChainParams &cp, std::vector<U128> &a, std::vector<uint64_t> &u);

}  // namespace pgmm
