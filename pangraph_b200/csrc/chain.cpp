// Host stage: anchor chaining with a range-minimum tree, as selected by every preset pangraph can reach (asm5/10/20
// set MM_F_RMQ).  Behaviour follows the reference's mg_lchain_rmq and helpers
// (packages/minimap2-sys/minimap2/lchain.c:250-368, :232-248, :27-111) and its balanced RMQ tree (krmq.h), including the
// tie rules of the subtree-minimum pointers, which depend on the tree's rotation history and are observable in chain
// scores.  The tree here is index-based over a node pool instead of pointer-and-macro based.
//
// This stage is sequential per query by construction (each anchor's score depends on its predecessors'); it runs on
// host threads, one query per thread, overlapped across the queries of a batch.
#include "chain.h"

#include <algorithm>
#include <cassert>
#include <cstring>
#include <set>
#include <cstdio>
#include <cstdlib>
#include <ctime>

namespace pgmm {

namespace {

constexpr int32_t NIL = -1;

struct Node {
  int32_t y;
  int64_t i;
  double pri;
  int32_t ch[2];
  int32_t smin;  // node holding the minimum priority of this subtree (history-dependent among ties)
  double smin_pri;  // its priority, cached so that comparisons need no second hop
  int8_t bal;
  uint32_t size;
};

// pool shared by the two trees of one chaining run; slot 0 is the scratch "fake root" used while erasing
struct Pool {
  std::vector<Node> nd;
  std::vector<int32_t> free_;
  Pool() { nd.resize(1); }
  int32_t alloc() {
    if (!free_.empty()) {
      int32_t k = free_.back();
      free_.pop_back();
      return k;
    }
    nd.emplace_back();
    return (int32_t)nd.size() - 1;
  }
  void release(int32_t k) { free_.push_back(k); }
};

constexpr int kMaxDepth = 64;

struct MinTree {
  Pool &P;
  int32_t root = NIL;
  explicit MinTree(Pool &p) : P(p) {}

  Node &N(int32_t k) { return P.nd[k]; }
  uint32_t size() { return root == NIL ? 0 : N(root).size; }
  uint32_t child_size(int32_t k, int d) { return N(k).ch[d] == NIL ? 0 : N(N(k).ch[d]).size; }
  static int cmp_key(int32_t y, int64_t i, const Node &b) {
    return y < b.y ? -1 : y > b.y ? 1 : (i > b.i) - (i < b.i);
  }
  bool lt(int32_t a, int32_t b) { return N(a).pri < N(b).pri; }

  // subtree minimum of p given the children it is about to have: left child's minimum beats p on ties, right child's
  // minimum beats both on ties
  void pull_min(int32_t p, int32_t l, int32_t r) {
    Node &np = N(p);
    int32_t s = p;
    double sp = np.pri;
    if (l != NIL && !(sp < N(l).smin_pri)) s = N(l).smin, sp = N(l).smin_pri;
    if (r != NIL && !(sp < N(r).smin_pri)) s = N(r).smin, sp = N(r).smin_pri;
    np.smin = s, np.smin_pri = sp;
  }

  int32_t rotate_single(int32_t p, int dir) {
    const int opp = 1 - dir;
    const int32_t q = N(p).ch[opp], s = N(p).smin;
    const double s_pri = N(p).smin_pri;
    const uint32_t size_p = N(p).size;
    N(p).size -= N(q).size - child_size(q, dir);
    N(q).size = size_p;
    pull_min(p, N(p).ch[dir], N(q).ch[dir]);
    N(q).smin = s, N(q).smin_pri = s_pri;
    N(p).ch[opp] = N(q).ch[dir];
    N(q).ch[dir] = p;
    return q;
  }

  int32_t rotate_double(int32_t p, int dir) {
    const int opp = 1 - dir;
    const int32_t q = N(p).ch[opp], r = N(q).ch[dir], s = N(p).smin;
    const double s_pri = N(p).smin_pri;
    const uint32_t size_r_dir = child_size(r, dir);
    N(r).size = N(p).size;
    N(p).size -= N(q).size - size_r_dir;
    N(q).size -= size_r_dir + 1;
    pull_min(p, N(p).ch[dir], N(r).ch[dir]);
    pull_min(q, N(q).ch[opp], N(r).ch[opp]);
    N(r).smin = s, N(r).smin_pri = s_pri;
    N(p).ch[opp] = N(r).ch[dir];
    N(r).ch[dir] = p;
    N(q).ch[dir] = N(r).ch[opp];
    N(r).ch[opp] = q;
    const int b1 = dir == 0 ? +1 : -1;
    if (N(r).bal == b1) N(q).bal = 0, N(p).bal = (int8_t)-b1;
    else if (N(r).bal == 0) N(q).bal = N(p).bal = 0;
    else N(q).bal = (int8_t)b1, N(p).bal = 0;
    N(r).bal = 0;
    return r;
  }

  void insert(int32_t x) {
    uint8_t turns[kMaxDepth];
    int32_t path[kMaxDepth];
    int32_t anchor = root, anchor_parent = NIL;  // deepest node on the path with a non-zero balance, and its parent
    int32_t p = root, q = NIL;
    int top = 0, path_len = 0, which = 0;
    const int32_t xy = N(x).y;
    const int64_t xi = N(x).i;
    while (p != NIL) {
      const int c = cmp_key(xy, xi, N(p));
      assert(c != 0);
      if (N(p).bal != 0) anchor_parent = q, anchor = p, top = 0;
      turns[top++] = (uint8_t)(which = c > 0);
      path[path_len++] = p;
      q = p, p = N(p).ch[which];
    }
    Node &nx = N(x);
    nx.bal = 0, nx.size = 1, nx.ch[0] = nx.ch[1] = NIL, nx.smin = x, nx.smin_pri = nx.pri;
    if (q == NIL) root = x;
    else N(q).ch[which] = x;
    if (anchor == NIL) return;
    for (int k = 0; k < path_len; ++k) ++N(path[k]).size;
    for (int k = path_len - 1; k >= 0; --k) {
      pull_min(path[k], N(path[k]).ch[0], N(path[k]).ch[1]);
      if (N(path[k]).smin != x) break;
    }
    top = 0;
    for (p = anchor; p != x; p = N(p).ch[turns[top]], ++top) N(p).bal += turns[top] == 0 ? -1 : +1;
    if (N(anchor).bal > -2 && N(anchor).bal < 2) return;
    which = N(anchor).bal < 0;
    const int b1 = which == 0 ? +1 : -1;
    q = N(anchor).ch[1 - which];
    int32_t r;
    if (N(q).bal == b1) {
      r = rotate_single(anchor, which);
      N(q).bal = N(anchor).bal = 0;
    } else r = rotate_double(anchor, which);
    if (anchor_parent == NIL) root = r;
    else N(anchor_parent).ch[anchor != N(anchor_parent).ch[0]] = r;
  }

  int32_t find(int32_t y, int64_t i) {
    int32_t p = root;
    while (p != NIL) {
      const int c = cmp_key(y, i, N(p));
      if (c < 0) p = N(p).ch[0];
      else if (c > 0) p = N(p).ch[1];
      else break;
    }
    return p;
  }

  // removes the node with key (y,i); returns its slot or NIL
  int32_t erase(int32_t y, int64_t i) {
    if (root == NIL) return NIL;
    int32_t path[kMaxDepth];
    uint8_t dir[kMaxDepth];
    const int32_t FAKE = 0;
    N(FAKE) = N(root);
    N(FAKE).ch[0] = root, N(FAKE).ch[1] = NIL;
    int d = 0, c = -1;
    int32_t p = FAKE;
    while (c != 0) {
      const int which = c > 0;
      dir[d] = (uint8_t)which;
      path[d++] = p;
      p = N(p).ch[which];
      if (p == NIL) return NIL;
      c = cmp_key(y, i, N(p));
    }
    for (int k = 1; k < d; ++k) --N(path[k]).size;
    if (N(p).ch[1] == NIL) {
      N(path[d - 1]).ch[dir[d - 1]] = N(p).ch[0];
    } else {
      int32_t q = N(p).ch[1];
      if (N(q).ch[0] == NIL) {  // the right child is the successor
        N(q).ch[0] = N(p).ch[0];
        N(q).bal = N(p).bal;
        N(path[d - 1]).ch[dir[d - 1]] = q;
        path[d] = q, dir[d++] = 1;
        N(q).size = N(p).size - 1;
      } else {  // successor = leftmost node of the right subtree
        int32_t r;
        const int e = d++;
        for (;;) {
          dir[d] = 0;
          path[d++] = q;
          r = N(q).ch[0];
          if (N(r).ch[0] == NIL) break;
          q = r;
        }
        N(r).ch[0] = N(p).ch[0];
        N(q).ch[0] = N(r).ch[1];
        N(r).ch[1] = N(p).ch[1];
        N(r).bal = N(p).bal;
        N(path[e - 1]).ch[dir[e - 1]] = r;
        path[e] = r, dir[e] = 1;
        for (int k = e + 1; k < d; ++k) --N(path[k]).size;
        N(r).size = N(p).size - 1;
      }
    }
    for (int k = d - 1; k >= 0; --k) pull_min(path[k], N(path[k]).ch[0], N(path[k]).ch[1]);
    while (--d > 0) {
      const int32_t q = path[d];
      const int which = dir[d], other = 1 - which;
      int b1 = 1, b2 = 2;
      if (which) b1 = -b1, b2 = -b2;
      N(q).bal += (int8_t)b1;
      if (N(q).bal == b1) break;
      else if (N(q).bal == b2) {
        const int32_t r = N(q).ch[other];
        if (N(r).bal == -b1) {
          N(path[d - 1]).ch[dir[d - 1]] = rotate_double(q, which);
        } else {
          N(path[d - 1]).ch[dir[d - 1]] = rotate_single(q, which);
          if (N(r).bal == 0) {
            N(r).bal = (int8_t)-b1;
            N(q).bal = (int8_t)b1;
            break;
          } else N(r).bal = N(q).bal = 0;
        }
      }
    }
    root = N(FAKE).ch[0];
    return p;
  }

  // minimum-priority node with lo <= key <= hi (closed), first strict minimum met on the two root-to-leaf walks
  int32_t range_min(int32_t lo_y, int64_t lo_i, int32_t hi_y, int64_t hi_i) {
    if (root == NIL) return NIL;
    int32_t path[2][kMaxDepth];
    int pc[2][kMaxDepth], plen[2] = {0, 0};
    int32_t p = root;
    while (p != NIL) {
      const int c = cmp_key(lo_y, lo_i, N(p));
      path[0][plen[0]] = p, pc[0][plen[0]++] = c;
      if (c < 0) p = N(p).ch[0];
      else if (c > 0) p = N(p).ch[1];
      else break;
    }
    p = root;
    while (p != NIL) {
      const int c = cmp_key(hi_y, hi_i, N(p));
      path[1][plen[1]] = p, pc[1][plen[1]++] = c;
      if (c < 0) p = N(p).ch[0];
      else if (c > 0) p = N(p).ch[1];
      else break;
    }
    int k;
    for (k = 0; k < plen[0] && k < plen[1]; ++k)
      if (path[0][k] == path[1][k] && pc[0][k] <= 0 && pc[1][k] >= 0) break;
    if (k == plen[0] || k == plen[1]) return NIL;
    const int lca = k;
    int32_t best = path[0][lca];
    double best_pri = N(best).pri;
    for (k = lca + 1; k < plen[0]; ++k)
      if (pc[0][k] <= 0) {
        const Node &nn = N(path[0][k]);
        if (nn.pri < best_pri) best = path[0][k], best_pri = nn.pri;
        if (nn.ch[1] != NIL && N(nn.ch[1]).smin_pri < best_pri) best = N(nn.ch[1]).smin, best_pri = N(nn.ch[1]).smin_pri;
      }
    for (k = lca + 1; k < plen[1]; ++k)
      if (pc[1][k] >= 0) {
        const Node &nn = N(path[1][k]);
        if (nn.pri < best_pri) best = path[1][k], best_pri = nn.pri;
        if (nn.ch[0] != NIL && N(nn.ch[0]).smin_pri < best_pri) best = N(nn.ch[0]).smin, best_pri = N(nn.ch[0]).smin_pri;
      }
    return best;
  }

  // greatest node <= (y,i), or NIL
  int32_t floor_node(int32_t y, int64_t i) {
    int32_t p = root, l = NIL;
    while (p != NIL) {
      const int c = cmp_key(y, i, N(p));
      if (c < 0) p = N(p).ch[0];
      else if (c > 0) l = p, p = N(p).ch[1];
      else return p;
    }
    return l;
  }

  // in-order cursor walking towards smaller keys
  struct Cursor {
    int32_t stack[kMaxDepth];
    int top = -1;
  };
  void seek(int32_t node, Cursor &c) {
    c.top = -1;
    int32_t p = root;
    const int32_t y = N(node).y;
    const int64_t i = N(node).i;
    while (p != NIL) {
      c.stack[++c.top] = p;
      const int k = cmp_key(y, i, N(p));
      if (k < 0) p = N(p).ch[0];
      else if (k > 0) p = N(p).ch[1];
      else break;
    }
  }
  bool prev(Cursor &c) {
    if (c.top < 0) return false;
    int32_t p = N(c.stack[c.top]).ch[0];
    if (p != NIL) {
      for (; p != NIL; p = N(p).ch[1]) c.stack[++c.top] = p;
      return true;
    }
    do {
      p = c.stack[c.top--];
    } while (c.top >= 0 && p == N(c.stack[c.top]).ch[0]);
    return c.top >= 0;
  }
};

// log2 approximation used by the chaining gap cost (mmpriv.h:118-126); only meaningful for x >= 2
inline float fast_log2(float x) {
  union {
    float f;
    uint32_t i;
  } z = {x};
  float log_2 = (float)(int)(((z.i >> 23) & 255) - 128);
  z.i &= ~(255u << 23);
  z.i += 127u << 23;
  log_2 += (-0.34484843f * z.f + 2.02466578f) * z.f - 0.67487759f;
  return log_2;
}

// score of extending the chain ending at aj with ai (lchain.c:232-248)
inline int32_t link_score(const U128 &ai, const U128 &aj, float pen_gap, float pen_skip, bool *exact, int32_t *width) {
  const int32_t dq = (int32_t)ai.y - (int32_t)aj.y, dr = (int32_t)(ai.x - aj.x);
  const int32_t dd = dr > dq ? dr - dq : dq - dr, dg = dr < dq ? dr : dq;
  const int32_t q_span = (int32_t)(aj.y >> 32 & 0xff);
  int32_t sc = q_span < dg ? q_span : dg;
  *width = dd;
  if (exact) *exact = dd == 0 && dg <= q_span;
  if (dd || dq > q_span) {
    const float lin_pen = pen_gap * (float)dd + pen_skip * (float)dg;
    const float log_pen = dd >= 1 ? fast_log2((float)(dd + 1)) : 0.0f;
    sc -= (int)(lin_pen + .5f * log_pen);
  }
  return sc;
}

// where a chain ending at z[k] stops when walked backwards (lchain.c:9-25)
// (zx = the chain end's score, zy = its anchor)
int64_t chain_stop(int32_t max_drop, int32_t zx, int64_t zy, const int32_t *f, const int32_t *p, int32_t *t) {
  int64_t i = zy, end_i = -1, max_i = i;
  int32_t max_s = 0;
  if (i < 0 || t[i] != 0) return i;
  do {
    t[i] = 2;
    end_i = i = p[i];
    const int32_t s = i < 0 ? zx : zx - f[i];
    if (s > max_s) max_s = s, max_i = i;
    else if (max_s - s > max_drop) break;
  } while (i >= 0 && t[i] == 0);
  for (i = zy; i >= 0 && i != end_i; i = p[i]) t[i] = 0;
  return max_i;
}

}  // namespace

void chain_find_segments(const ChainParams &cp, const U128 *a, int64_t n, std::vector<ChainSeg> &segs) {
  segs.clear();
  if (n == 0) return;
  const int64_t max_dist = cp.max_dist < cp.bw ? cp.bw : cp.max_dist;
  int64_t s = 0;
  for (int64_t i = 1; i < n; ++i)
    if (a[i].x >> 32 != a[i - 1].x >> 32 || a[i].x > a[i - 1].x + (uint64_t)max_dist) segs.push_back(ChainSeg{s, i}), s = i;
  segs.push_back(ChainSeg{s, n});
}

void chain_fill_host(const ChainParams &cp, const U128 *a, int64_t n, int64_t seg_start, int64_t seg_end, int32_t *f, int32_t *p,
                     int32_t *v, int32_t *t) {
  int max_dist = cp.max_dist, max_dist_inner = cp.max_dist_inner;
  const int bw = cp.bw;
  if (max_dist < bw) max_dist = bw;
  if (max_dist_inner <= 0 || max_dist_inner >= max_dist) max_dist_inner = 0;
  Pool pool;
  MinTree outer(pool);
  int64_t i0 = seg_start, st = seg_start, st_inner = seg_start;
  // The second ("inner") tree of the reference is only ever searched for a floor element and walked backwards in key
  // order (lchain.c:323-348): what it returns depends on the SET of anchors it holds, not on its shape.  That set is
  // the index window [st_inner, i0): a small window is kept as a sorted array of (y, index) (insertions and evictions
  // are short memmoves), a large one (repeats) is mirrored in an ordered set.
  constexpr int64_t kMirrorOn = 1024, kMirrorOff = 256;
  std::set<std::pair<int32_t, int64_t>> mirror;
  bool mirrored = false;
  std::vector<std::pair<int32_t, int64_t>> near;  // ascending, the members of the window while !mirrored

  for (int64_t i = seg_start; i < seg_end; ++i) {
    int64_t max_j = -1;
    const int32_t q_span = (int32_t)(a[i].y >> 32 & 0xff);
    int32_t max_f = q_span;
    // anchors sharing one target position enter the trees together, once the position changes (lchain.c:280-293)
    if (i0 < i && a[i0].x != a[i].x) {
      for (int64_t j = i0; j < i; ++j) {
        const int32_t q = pool.alloc();
        Node &nq = pool.nd[q];
        nq.y = (int32_t)a[j].y, nq.i = j;
        nq.pri = -(f[j] + 0.5 * cp.pen_gap * ((int32_t)a[j].x + (int32_t)a[j].y));
        outer.insert(q);
        if (max_dist_inner > 0) {
          const std::pair<int32_t, int64_t> key((int32_t)a[j].y, j);
          if (mirrored) mirror.insert(key);
          else near.insert(std::upper_bound(near.begin(), near.end(), key), key);
        }
      }
      i0 = i;
    }
    // evict anchors that left the window, changed target/strand, or overflow the size cap (:295-312)
    while (st < i && (a[i].x >> 32 != a[st].x >> 32 || a[i].x > a[st].x + max_dist || (int64_t)outer.size() > cp.cap_rmq_size)) {
      const int32_t q = outer.erase((int32_t)a[st].y, st);
      if (q != NIL) pool.release(q);
      ++st;
    }
    if (max_dist_inner > 0) {
      while (st_inner < i && (a[i].x >> 32 != a[st_inner].x >> 32 || a[i].x > a[st_inner].x + max_dist_inner ||
                              std::max<int64_t>(0, i0 - st_inner) > cp.cap_rmq_size)) {
        if (st_inner < i0) {
          const std::pair<int32_t, int64_t> key((int32_t)a[st_inner].y, st_inner);
          if (mirrored) mirror.erase(key);
          else near.erase(std::lower_bound(near.begin(), near.end(), key));
        }
        ++st_inner;
      }
      const int64_t in_size = std::max<int64_t>(0, i0 - st_inner);
      if (!mirrored && in_size > kMirrorOn) {
        mirror.clear();
        mirror.insert(near.begin(), near.end());
        near.clear();
        mirrored = true;
      } else if (mirrored && in_size < kMirrorOff) {
        near.assign(mirror.begin(), mirror.end());
        mirror.clear();
        mirrored = false;
      }
    }
    // best predecessor by priority inside the query window, then a bounded scan of the near neighbourhood (:313-351)
    const int32_t qn = outer.range_min((int32_t)a[i].y - max_dist, INT32_MAX, (int32_t)a[i].y, 0);
    if (qn != NIL) {
      int32_t width, n_skip = 0;
      bool exact;
      int64_t j = pool.nd[qn].i;
      int32_t sc = f[j] + link_score(a[i], a[j], cp.pen_gap, cp.pen_skip, &exact, &width);
      if (width <= bw && sc > max_f) max_f = sc, max_j = j;
      if (!exact && max_dist_inner > 0 && i0 > st_inner && (int32_t)a[i].y > 0) {
        const int32_t y_hi = (int32_t)a[i].y - 1, y_lo = (int32_t)a[i].y - max_dist_inner;
        // visits window members in descending (y, index) order from the floor of (y_hi, n), like the reference's cursor
        const auto visit = [&](int64_t jj) -> bool {
          j = jj;
          sc = f[j] + link_score(a[i], a[j], cp.pen_gap, cp.pen_skip, nullptr, &width);
          if (width <= bw) {
            if (sc > max_f) {
              max_f = sc, max_j = j;
              if (n_skip > 0) --n_skip;
            } else if (t[j] == (int32_t)i) {
              if (++n_skip > cp.max_chn_skip) return false;
            }
            if (p[j] >= 0) t[p[j]] = (int32_t)i;
          }
          return true;
        };
        if (mirrored) {
          auto it = mirror.upper_bound(std::make_pair(y_hi, n));
          while (it != mirror.begin()) {
            --it;
            if (it->first < y_lo) break;
            if (!visit(it->second)) break;
          }
        } else {
          auto it = std::upper_bound(near.begin(), near.end(), std::make_pair(y_hi, n));
          while (it != near.begin()) {
            --it;
            if (it->first < y_lo) break;
            if (!visit(it->second)) break;
          }
        }
      }
    }
    f[i] = max_f, p[i] = (int32_t)max_j;
    v[i] = max_j >= 0 && v[max_j] > max_f ? v[max_j] : max_f;
  }
}

void chain_backtrack(const ChainParams &cp, std::vector<U128> &a, const int32_t *f, const int32_t *p, int32_t *v, int32_t *t,
                     std::vector<uint64_t> &u) {
  u.clear();
  const int64_t n = (int64_t)a.size();
  if (n == 0) return;
  const int bw = cp.bw;
  static const bool tr__ = getenv("PGMM_TRACE_BT") != nullptr;
  const auto now__ = []() { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec * 1e3 + t.tv_nsec * 1e-6; };
  double t0__ = now__(), t1__;
#define BT_MARK(what) if (tr__) { t1__ = now__(); fprintf(stderr, "[bt] %-10s %.1f ms\n", what, t1__ - t0__); t0__ = t1__; }
  // ---- backtrack (lchain.c:27-76): best end points first, each anchor used once ----
  const int32_t min_sc = cp.min_sc, min_cnt = cp.min_cnt, max_drop = bw;
  // chain ends as score << 32 | anchor, sorted by score like the reference's (score, anchor) pairs: the unstable radix
  // sort's moves depend on the key digits only, so packing the pair into 8 bytes leaves the permutation unchanged
  std::vector<uint64_t> z;
  {
    int64_t n_z = 0;
    for (int64_t i = 0; i < n; ++i) n_z += f[i] >= min_sc;
    z.resize((size_t)n_z);
    uint64_t *zp = z.data();
    for (int64_t i = 0; i < n; ++i)
      if (f[i] >= min_sc) *zp++ = (uint64_t)(uint32_t)f[i] << 32 | (uint64_t)i;
  }
  if (z.empty()) {
    a.clear();
    return;
  }
  BT_MARK("z build");
  if (min_sc >= 0) flag_sort(z.data(), z.data() + z.size(), [](uint64_t v) { return v >> 32; });
  else {  // negative scores sign-extend into the upper key bytes (lchain.c:41): sort the reference's 16-byte pairs, then pack
    std::vector<U128> zz(z.size());
    for (size_t k = 0; k < z.size(); ++k) zz[k] = U128{(uint64_t)(int64_t)(int32_t)(z[k] >> 32), (uint64_t)(uint32_t)z[k]};
    flag_sort_128x(zz.data(), zz.data() + zz.size());
    for (size_t k = 0; k < z.size(); ++k) z[k] = (uint64_t)(uint32_t)(int32_t)zz[k].x << 32 | zz[k].y;
  }
  BT_MARK("z sort");
  std::fill(t, t + n, 0);
  int64_t n_v = 0;
  for (int64_t k = (int64_t)z.size() - 1; k >= 0; --k) {
    const int32_t zx = (int32_t)(z[k] >> 32);
    const int64_t zy = (int64_t)(uint32_t)z[k];
    if (t[zy] != 0) continue;
    const int64_t n_v0 = n_v;
    const int64_t end_i = chain_stop(max_drop, zx, zy, f, p, t);
    int64_t i;
    for (i = zy; i != end_i; i = p[i]) v[n_v++] = (int32_t)i, t[i] = 1;
    const int32_t sc = i < 0 ? zx : zx - f[i];
    if (sc >= min_sc && n_v > n_v0 && n_v - n_v0 >= min_cnt) u.push_back((uint64_t)sc << 32 | (uint64_t)(n_v - n_v0));
    else n_v = n_v0;
  }
  BT_MARK("walk");
  if (u.empty()) {
    a.clear();
    return;
  }
  // ---- compact (lchain.c:78-111): anchors of each chain in forward order, chains ordered by first target position ----
  const int64_t n_u = (int64_t)u.size();
  std::vector<U128> b((size_t)n_v), w((size_t)n_u);
  for (int64_t i = 0, k = 0; i < n_u; ++i) {
    const int64_t k0 = k, ni = (int32_t)u[i];
    for (int64_t j = 0; j < ni; ++j) b[k++] = a[v[k0 + (ni - j - 1)]];
  }
  for (int64_t i = 0, k = 0; i < n_u; ++i) {
    w[i].x = b[k].x, w[i].y = (uint64_t)k << 32 | (uint64_t)i;
    k += (int32_t)u[i];
  }
  flag_sort_128x(w.data(), w.data() + n_u);
  std::vector<uint64_t> u2((size_t)n_u);
  a.resize((size_t)n_v);
  for (int64_t i = 0, k = 0; i < n_u; ++i) {
    const int32_t j = (int32_t)w[i].y, cnt = (int32_t)u[j];
    u2[i] = u[j];
    memcpy(&a[k], &b[w[i].y >> 32], (size_t)cnt * sizeof(U128));
    k += cnt;
  }
  u.swap(u2);
  BT_MARK("compact");
#undef BT_MARK
}

void chain_rmq(const ChainParams &cp, std::vector<U128> &a, std::vector<uint64_t> &u) {
  u.clear();
  const int64_t n = (int64_t)a.size();
  if (n == 0) return;
  std::vector<int32_t> f(n), p(n), t(n, 0), v(n);
  std::vector<ChainSeg> segs;
  chain_find_segments(cp, a.data(), n, segs);
  for (const ChainSeg &sg : segs) chain_fill_host(cp, a.data(), n, sg.start, sg.end, f.data(), p.data(), v.data(), t.data());
  chain_backtrack(cp, a, f.data(), p.data(), v.data(), t.data(), u);
}

}  // namespace pgmm
