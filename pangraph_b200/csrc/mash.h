// K7: shared-minimizer counts for pangraph's mash distance on the GPU.
// Reference (PG = packages/pangraph/src): PG/distance/mash/mash_distance.rs:9-68 over minimizer.rs:49-160.
#pragma once
#include <cstdint>

namespace pgmm {
namespace mash {

struct Stats {
  double upload_ms = 0, sketch_ms = 0, sort_ms = 0, pair_ms = 0;  // device time of the four stages (CUDA events)
  uint64_t bases = 0, tiles = 0, minimizers = 0, unique_keys = 0, shared_values = 0, launches = 0;
};

// counts = n x n, row-major.  counts[i][i] = number of distinct minimizer values of sequence i; counts[i][j] = counts[j][i] =
// number of distinct values sequences i and j have in common -- the matrix mash_distance.rs:28-48 accumulates.
// -> 0; -1: k outside 1..31 or w outside 1..255 (minimizer.rs:53-54); -2: n < 1; -3: 2k + bits(n) > 64 (the sort key packs
// value and sequence into one 64-bit word); -4: a sequence of 2^31 bases or more (its positions do not fit the reference's
// id << 32 | locus << 1 packing either); -5: the incidence bitmap does not fit the device.
int shared_counts(const char *const *seqs, const int64_t *lens, int n, int k, int w, uint32_t *counts, Stats *st);

}  // namespace mash
}  // namespace pgmm
