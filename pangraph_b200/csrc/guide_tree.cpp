#include "guide_tree.h"

#include <cmath>
#include <cstring>
#include <limits>

#include "text_util.h"

namespace pgmm {
namespace gt {

int distances_from_counts(const uint32_t *counts, int n, double *dist) {
  for (int i = 0; i < n; ++i)
    if (counts[(size_t)i * n + i] == 0) return 1 + i;  // "no self-hit found for sequence i"
  for (int i = 0; i < n; ++i) {
    const double self = (double)counts[(size_t)i * n + i];
    dist[(size_t)i * n + i] = 0.0;
    for (int j = i + 1; j < n; ++j) {  // the SMALLER index's own count is the denominator (mash_distance.rs:60)
      const double d = 1.0 - (double)counts[(size_t)i * n + j] / self;
      dist[(size_t)i * n + j] = d, dist[(size_t)j * n + i] = d;
    }
  }
  return 0;
}

namespace {

// The distance matrix of the taxa still alive, kept dense in the top-left m x m corner of an n x n block (stride n): removing
// a taxon moves the rows and columns behind it up, as ndarray's remove_index does.
struct Work {
  int n, m;
  std::vector<double> d, col_sum, row_sum;
  double &at(int r, int c) { return d[(size_t)r * n + c]; }
};

// One row's sum the way ndarray 0.16.1 folds a contiguous lane (numeric_util::unrolled_fold): eight interleaved partial
// sums over blocks of eight, combined (0+4) (1+5) (2+6) (3+7) in that order, then the tail one element at a time.
double fold_lane(const double *x, int len) {
  double part[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int at = 0;
  for (; at + 8 <= len; at += 8)
    for (int u = 0; u < 8; ++u) part[u] += x[at + u];
  double s = 0.0;
  s += part[0] + part[4];
  s += part[1] + part[5];
  s += part[2] + part[6];
  s += part[3] + part[7];
  for (; at < len; ++at) s += x[at];
  return s;
}

// create_Q_matrix + argmin (neighbor_joining.rs:46-71) without materialising Q: sum_axis(Axis(0)) adds the rows one after
// the other, sum_axis(Axis(1)) folds each row; Q[r][c] = ((m - 2) D[r][c] - col_sum[c]) - row_sum[r]; the diagonal is +inf;
// the first strictly smaller element in row-major order wins.
int closest_pair(Work &w, int &pi, int &pj) {
  const int m = w.m;
  for (int c = 0; c < m; ++c) w.col_sum[c] = 0.0;
  for (int r = 0; r < m; ++r) {
    const double *row = &w.at(r, 0);
    for (int c = 0; c < m; ++c) w.col_sum[c] += row[c];
    w.row_sum[r] = fold_lane(row, m);
  }
  const double scale = (double)m - 2.0;
  double best = std::numeric_limits<double>::infinity();  // Q[0][0]
  int br = 0, bc = 0;
  for (int r = 0; r < m; ++r) {
    const double *row = &w.at(r, 0);
    const double rs = w.row_sum[r];
    for (int c = 0; c < m; ++c) {
      if (c == r) continue;
      const double q = (scale * row[c] - w.col_sum[c]) - rs;
      if (std::isnan(q)) return -2;
      if (q < best) best = q, br = r, bc = c;
    }
  }
  pi = br < bc ? br : bc, pj = br < bc ? bc : br;
  return 0;
}

}  // namespace

int neighbor_joining(const double *dist, int n, Tree &out) {
  out = Tree();
  if (n < 2) return -1;
  out.n = n;
  Work w;
  w.n = n, w.m = n;
  w.d.assign(dist, dist + (size_t)n * n);
  w.col_sum.resize((size_t)n), w.row_sum.resize((size_t)n);
  std::vector<int32_t> alive((size_t)n);
  for (int i = 0; i < n; ++i) alive[(size_t)i] = i;
  std::vector<double> joined((size_t)n);
  while (w.m > 2) {
    int i, j;
    if (int rc = closest_pair(w, i, j)) return rc;
    const int m = w.m;
    out.left.push_back(alive[(size_t)i]), out.right.push_back(alive[(size_t)j]);
    alive[(size_t)i] = n + (int32_t)out.left.size() - 1;
    alive.erase(alive.begin() + j);
    // dist(): distances of the new node to everything (neighbor_joining.rs:73-80), written over row and column i
    const double dij = w.at(i, j);
    for (int c = 0; c < m; ++c) joined[(size_t)c] = 0.5 * ((w.at(i, c) + w.at(j, c)) - dij);
    for (int c = 0; c < m; ++c) w.at(i, c) = joined[(size_t)c], w.at(c, i) = joined[(size_t)c];
    w.at(i, i) = 0.0;
    for (int r = 0; r < m; ++r)  // drop column j
      std::memmove(&w.at(r, j), &w.at(r, j + 1), (size_t)(m - 1 - j) * sizeof(double));
    for (int r = j; r + 1 < m; ++r)  // drop row j
      std::memcpy(&w.at(r, 0), &w.at(r + 1, 0), (size_t)(m - 1) * sizeof(double));
    --w.m;
  }
  out.left.push_back(alive[0]), out.right.push_back(alive[1]);
  return 0;
}

std::vector<int32_t> postorder(const Tree &t) {
  std::vector<int32_t> order;
  if (t.n < 1) return order;
  std::vector<std::pair<int32_t, bool>> stack;
  stack.push_back({(int32_t)(t.n == 1 ? 0 : 2 * t.n - 2), false});
  while (!stack.empty()) {
    const auto [v, expanded] = stack.back();
    stack.pop_back();
    if (expanded || v < t.n) {
      order.push_back(v);
      continue;
    }
    stack.push_back({v, true});
    stack.push_back({t.right[(size_t)(v - t.n)], false});
    stack.push_back({t.left[(size_t)(v - t.n)], false});
  }
  return order;
}

std::string to_newick(const Tree &t, const std::vector<std::string> &leaf_names) {
  std::string s;
  if (t.n < 1) return ";";
  // (left subtree) , (right subtree) ) -- an explicit stack of "what to print next"
  struct Item {
    int32_t node;
    char lit;  // non-zero: print this character instead of a node
  };
  std::vector<Item> stack;
  stack.push_back({(int32_t)(t.n == 1 ? 0 : 2 * t.n - 2), 0});
  while (!stack.empty()) {
    const Item it = stack.back();
    stack.pop_back();
    if (it.lit) s.push_back(it.lit);
    else if (it.node < t.n) s += leaf_names[(size_t)it.node];
    else {
      s.push_back('(');
      stack.push_back({0, ')'});
      stack.push_back({t.right[(size_t)(it.node - t.n)], 0});
      stack.push_back({0, ','});
      stack.push_back({t.left[(size_t)(it.node - t.n)], 0});
    }
  }
  s.push_back(';');
  return s;
}

Tree balance(const Tree &t) {
  Tree out;
  out.n = t.n;
  if (t.n < 2) return out;
  std::vector<int32_t> tips;
  for (int32_t v : postorder(t))
    if (v < t.n) tips.push_back(v);
  // bisect(tips[lo..hi)): children first, so a node's number is larger than its children's and the root comes last
  struct Frame {
    int lo, hi, stage;
    int32_t l;
  };
  std::vector<Frame> stack;
  std::vector<int32_t> ret;
  stack.push_back({0, (int)tips.size(), 0, -1});
  while (!stack.empty()) {
    Frame &f = stack.back();
    if (f.hi - f.lo <= 1) {
      ret.push_back(tips[(size_t)f.lo]);
      stack.pop_back();
      continue;
    }
    const int mid = f.lo + (f.hi - f.lo) / 2;
    if (f.stage == 0) {
      f.stage = 1;
      stack.push_back({f.lo, mid, 0, -1});
    } else if (f.stage == 1) {
      f.l = ret.back(), ret.pop_back();
      f.stage = 2;
      stack.push_back({mid, f.hi, 0, -1});
    } else {
      const int32_t r = ret.back();
      ret.pop_back();
      out.left.push_back(f.l), out.right.push_back(r);
      ret.push_back(t.n + (int32_t)out.left.size() - 1);
      stack.pop_back();
    }
  }
  return out;
}

namespace {

struct Newick {
  const std::string &in;
  size_t pos = 0;
  std::string err;
  std::vector<std::string> &names;
  Tree &tree;
  // nodes while parsing: >= 0 leaf number, < 0: -(1 + index into the internal arrays); renumbered at the end
  std::vector<int32_t> il, ir;

  Newick(const std::string &s, std::vector<std::string> &n, Tree &t) : in(s), names(n), tree(t) {}
  bool eof() const { return pos >= in.size(); }
  // the reference walks chars of a UTF-8 string; positions in its messages are byte offsets and every character it tests
  // for is ASCII, so bytes behave the same except for non-ASCII whitespace (char::is_whitespace, text_util.h)
  size_t ws_len() const { return text::ws_len(in.data(), pos, in.size()); }
  size_t char_len() const { return text::char_len(in.data(), pos, in.size()); }
  void skip_ws() {
    while (!eof()) {
      const size_t n = ws_len();
      if (!n) break;
      pos += n;
    }
  }
  bool peek_is(char c) const { return !eof() && in[pos] == c; }

  bool parse_name(std::string &name) {  // newick.rs:226-262; false = no name
    skip_ws();
    name.clear();
    if (peek_is('\'')) {
      ++pos;
      for (;;) {
        if (eof()) return false;
        if (in[pos] == '\'') {
          ++pos;
          if (peek_is('\'')) name.push_back('\''), ++pos;
          else break;
        } else {
          const size_t n = char_len();
          name.append(in, pos, n), pos += n;
        }
      }
      return true;
    }
    while (!eof()) {
      const char c = in[pos];
      if (c == '(' || c == ')' || c == ',' || c == ':' || c == ';' || ws_len()) break;
      const size_t n = char_len();
      name.append(in, pos, n), pos += n;
    }
    return !name.empty();
  }
  bool parse_branch_length() {  // newick.rs:264-281
    skip_ws();
    if (peek_is(':')) {
      ++pos;
      skip_ws();
      const size_t start = pos;
      while (!eof() && ((in[pos] >= '0' && in[pos] <= '9') || in[pos] == '.' || in[pos] == 'e' || in[pos] == 'E' || in[pos] == '+' ||
                        in[pos] == '-'))
        ++pos;
      if (pos == start) {
        err = "Newick: expected a number after ':' at position " + std::to_string(pos);
        return false;
      }
    }
    return true;
  }
  // newick.rs:184-224 with the recursion unrolled onto a stack of open parentheses (a caterpillar tree over thousands of
  // genomes nests as deep as it has leaves)
  bool parse_tree(int32_t &root) {
    std::vector<std::vector<int32_t>> open;  // children collected so far for each open '('
    for (;;) {
      skip_ws();
      int32_t done;
      if (peek_is('(')) {
        ++pos;
        skip_ws();
        open.emplace_back();
        continue;
      }
      {
        std::string name;
        if (!parse_name(name)) {
          err = "Newick: leaf without a name at position " + std::to_string(pos);
          return false;
        }
        if (!parse_branch_length()) return false;
        done = (int32_t)names.size();
        names.push_back(name);
      }
      // a finished subtree: hand it to the enclosing '(' and close as many of them as end here
      for (;;) {
        if (open.empty()) {
          root = done;
          return true;
        }
        open.back().push_back(done);
        skip_ws();
        if (peek_is(',')) {
          ++pos;
          skip_ws();
          break;  // next sibling
        }
        if (peek_is(')')) ++pos;
        else if (!eof()) {
          const size_t n = char_len();
          err = "Newick: expected ')' or ',' at position " + std::to_string(pos) + ", found '" + in.substr(pos, n) + "'";
          return false;
        } else {
          err = "Newick: unexpected end of input, expected ')'";
          return false;
        }
        std::string label;
        parse_name(label);  // internal labels are read and dropped
        if (!parse_branch_length()) return false;
        const std::vector<int32_t> kids = open.back();
        open.pop_back();
        if (kids.size() != 2) {
          err = "Newick: internal node has " + std::to_string(kids.size()) + " children; only strictly bifurcating trees are supported";
          return false;
        }
        il.push_back(kids[0]), ir.push_back(kids[1]);
        done = -(int32_t)il.size();
      }
    }
  }
};

}  // namespace

bool parse_newick(const std::string &text, Tree &out, std::vector<std::string> &leaf_names, std::string &err) {
  out = Tree();
  leaf_names.clear();
  Newick p(text, leaf_names, out);
  p.skip_ws();
  if (p.eof()) {
    err = "Newick input is empty";
    return false;
  }
  int32_t root = 0;
  if (!p.parse_tree(root)) {
    err = p.err;
    return false;
  }
  p.skip_ws();
  if (p.peek_is(';')) ++p.pos;
  p.skip_ws();
  if (!p.eof()) {
    err = "Newick: unexpected trailing content at position " + std::to_string(p.pos) + ": '" + text.substr(p.pos) + "'";
    return false;
  }
  const int n = (int)leaf_names.size();
  out.n = n;
  // internal nodes were closed children-first, so creation order already puts the root last
  auto renumber = [n](int32_t v) { return v >= 0 ? v : n + (-v - 1); };
  for (size_t t = 0; t < p.il.size(); ++t) out.left.push_back(renumber(p.il[t])), out.right.push_back(renumber(p.ir[t]));
  (void)root;
  return true;
}

}  // namespace gt
}  // namespace pgmm
