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

// a: anchors sorted by x (in), anchors of the kept chains, chain after chain (out); u[i] = score<<32 | n_anchors.
void chain_rmq(const ChainParams &cp, std::vector<U128> &a, std::vector<uint64_t> &u);

}  // namespace pgmm
