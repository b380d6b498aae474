// Host side of the guide-tree path: distances from K7's counts, neighbour joining, Newick in and out, the
// balanced re-rooting.  Reference (PG = packages/pangraph/src): PG/distance/mash/mash_distance.rs:50-67,
// PG/tree/neighbor_joining.rs:16-101, PG/tree/newick.rs:11-62,162-280, PG/tree/balance.rs:4-18, PG/tree/clade.rs:49-71.
//
// A tree over n leaves is two arrays: leaves are 0..n-1, internal node n + t has children left[t], right[t], the root is
// 2n - 2 (n == 1: the single leaf is the root and the arrays are empty).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace pgmm {
namespace gt {

struct Tree {
  int n = 0;
  std::vector<int32_t> left, right;
};

// mash_distance.rs:50-67.  -> 0, or 1 + i when sequence i has no minimizer of its own (the reference panics)
int distances_from_counts(const uint32_t *counts, int n, double *dist);

// neighbor_joining.rs:16-35 on an n x n row-major matrix.  -> 0; -1: n < 2 (the reference indexes nodes[1]); -2: a NaN met
// while looking for the smallest Q (ndarray-stats' UndefinedOrder)
int neighbor_joining(const double *dist, int n, Tree &out);

// newick.rs:43-62.  Leaves are numbered in order of appearance.  -> true, or false with the reference's message in `err`
bool parse_newick(const std::string &text, Tree &out, std::vector<std::string> &leaf_names, std::string &err);
// newick.rs:11-38 (internal nodes carry no label)
std::string to_newick(const Tree &t, const std::vector<std::string> &leaf_names);
// balance.rs:4-18: same leaves left to right, bisected
Tree balance(const Tree &t);
// clade.rs:49-71: left subtree, right subtree, node
std::vector<int32_t> postorder(const Tree &t);

}  // namespace gt
}  // namespace pgmm
