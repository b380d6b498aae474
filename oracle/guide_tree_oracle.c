/* TEST INFRASTRUCTURE ONLY -- never linked into the product (pangraph_b200/), only tests/, __graft_entry__.smoke() and
 * bench.py's CPU legs may load this.
 *
 * Plain-C restatement of the reference's guide-tree path, PG = /root/reference/packages/pangraph/src:
 *   orc_mash_hash       PG/distance/mash/hash.rs:3-12
 *   orc_mash_sketch     PG/distance/mash/minimizer.rs:49-160   (the emitted list: order and repetitions included)
 *   orc_mash_distance   PG/distance/mash/mash_distance.rs:9-68
 *   orc_nj_q_matrix     PG/tree/neighbor_joining.rs:46-62
 *   orc_nj_dist         PG/tree/neighbor_joining.rs:73-80
 *   orc_nj_tree         PG/tree/neighbor_joining.rs:16-35, 82-101
 *
 * Pinned by the reference's own unit vectors (tests/test_oracle_guide_tree.py): hash.rs:20-27, minimizer.rs:190-211,
 * mash_distance.rs:91-151, neighbor_joining.rs:111-151 and the two trees of its disabled tests (:203-288).
 *
 * PARITY UNPINNED for one detail of neighbour joining: the order in which the row / column sums of the distance matrix
 * are accumulated lives in the third-party crate ndarray 0.16.1 (Cargo.lock), which is not vendored in /root/reference.
 * Restated here from that crate's published source: sum_axis over the axis with the smallest stride (the columns of a
 * row: Axis(1)) folds each lane with eight interleaved partial sums (numeric_util::unrolled_fold), sum_axis over the other
 * axis adds the rows one after the other.  The reference's test matrices are small integers, for which every order gives
 * the same bits, so they cannot tell the orders apart.  ndarray-stats 0.6.0 argmin: first strictly smaller element in
 * row-major order.  Everything else on this path is integer work and exact IEEE division / subtraction.
 *
 * Integer overflow: hash.rs uses plain `+` on u64, which wraps in a release build (the shipped binary) and traps in a
 * debug build when k > 28 or so; unsigned C arithmetic is the release behaviour. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define U64MAX (~(uint64_t)0)

uint64_t orc_mash_hash(uint64_t x, uint64_t mask) { /* hash.rs:3-12 */
  x = (~x + (x << 21)) & mask;
  x = x ^ (x >> 24);
  x = (x + (x << 3) + (x << 8)) & mask;
  x = x ^ (x >> 14);
  x = (x + (x << 2) + (x << 4)) & mask;
  x = x ^ (x >> 28);
  x = (x + (x << 31)) & mask;
  return x;
}

static int mash_code(unsigned char c) { /* minimizer.rs:163-181: A C G T/U in either case, everything else 4 */
  switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': case 'U': case 'u': return 3;
    default: return 4;
  }
}

typedef struct {
  uint64_t value, position;
} mzr_t;

typedef struct {
  mzr_t *a;
  int64_t n, cap;
} mzr_vec_t;

static void mzr_push(mzr_vec_t *v, mzr_t m) {
  if (v->n == v->cap) {
    v->cap = v->cap ? 2 * v->cap : 256;
    v->a = (mzr_t *)realloc(v->a, (size_t)v->cap * sizeof(mzr_t));
  }
  v->a[v->n++] = m;
}

/* minimizer.rs:49-160, statement by statement */
static void mash_sketch(const char *seq, int64_t len, uint64_t id, int k, int w, mzr_vec_t *out) {
  const mzr_t none = {U64MAX, U64MAX};
  uint64_t fwd = 0, rev = 0;
  const uint64_t mask = ((uint64_t)1 << (2 * k)) - 1, shift = 2 * (uint64_t)(k - 1);
  mzr_t min = none;
  mzr_t *window = (mzr_t *)malloc((size_t)w * sizeof(mzr_t));
  for (int i = 0; i < w; ++i) window[i] = none;
  uint64_t l = 0;
  int bi = 0, mi = 0;
  for (int64_t at = 0; at < len; ++at) {
    const uint64_t locus = (uint64_t)at + 1;
    const uint64_t c = (uint64_t)mash_code((unsigned char)seq[at]);
    mzr_t nw = none;
    if (c >= 4) l = 0;
    else {
      fwd = ((fwd << 2) | c) & mask;
      rev = (rev >> 2) | ((3 ^ c) << shift);
      l += 1;
      if (l >= (uint64_t)k) {
        const uint64_t pos = (id << 32) | (locus << 1);
        if (fwd <= rev) nw.value = orc_mash_hash(fwd, mask), nw.position = pos;
        else nw.value = orc_mash_hash(rev, mask), nw.position = pos | 1;
      }
    }
    window[bi] = nw;
    if (l == (uint64_t)(w + k - 1) && min.value != U64MAX) {
      for (int i = bi + 1; i < w; ++i)
        if (min.value == window[i].value && min.position != window[i].position) mzr_push(out, window[i]);
      for (int i = 0; i <= bi; ++i)
        if (min.value == window[i].value && min.position != window[i].position) mzr_push(out, window[i]);
    }
    if (nw.value < min.value) {
      if (l >= (uint64_t)(w + k) && min.value != U64MAX) mzr_push(out, min);
      min = nw, mi = bi;
    } else if (bi == mi) {
      if (l >= (uint64_t)(w + k - 1) && min.value != U64MAX) mzr_push(out, min);
      min.value = U64MAX; /* the position stays */
      for (int i = bi + 1; i < w; ++i)
        if (window[i].value < min.value) mi = i, min = window[i];
      for (int i = 0; i <= bi; ++i)
        if (window[i].value < min.value) mi = i, min = window[i];
      if (l >= (uint64_t)(w + k - 1) && min.value != U64MAX) {
        for (int i = bi + 1; i < w; ++i)
          if (min.value == window[i].value && min.position != window[i].position) mzr_push(out, window[i]);
        for (int i = 0; i <= bi; ++i)
          if (min.value == window[i].value && min.position != window[i].position) mzr_push(out, window[i]);
      }
    }
    if (++bi >= w) bi = 0;
  }
  if (min.value != U64MAX) mzr_push(out, min);
  free(window);
}

/* -> number of minimizers (0 = the reference's "No minimizers found" error); the first `cap` are stored */
int64_t orc_mash_sketch(const char *seq, int64_t len, uint64_t id, int k, int w, uint64_t *value, uint64_t *position, int64_t cap) {
  if (k < 1 || k >= 32 || w < 1 || w >= 256) return -1; /* the reference's assert!s */
  mzr_vec_t v = {0, 0, 0};
  mash_sketch(seq, len, id, k, w, &v);
  for (int64_t i = 0; i < v.n && i < cap; ++i) value[i] = v.a[i].value, position[i] = v.a[i].position;
  free(v.a);
  return v.n;
}

static int cmp_value(const void *a, const void *b) {
  const mzr_t *x = (const mzr_t *)a, *y = (const mzr_t *)b;
  return x->value < y->value ? -1 : x->value > y->value;
}
static int cmp_u32(const void *a, const void *b) {
  const uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
  return x < y ? -1 : x > y;
}

/* mash_distance.rs:9-68.  out = n x n doubles (n >= 1); -> 0, -(1+i) when sequence i has no minimizer (the reference
 * panics: .expect("no minimizer found ...")), -1000000 on n == 0 (array![[]] fails neighbour joining's shape assert). */
int orc_mash_distance(const char *const *seqs, const int64_t *lens, int n, int k, int w, double *out) {
  if (n <= 0) return -1000000;
  mzr_vec_t all = {0, 0, 0};
  for (int i = 0; i < n; ++i) {
    const int64_t before = all.n;
    mash_sketch(seqs[i], lens[i], (uint64_t)i, k, w, &all);
    if (all.n == before) {
      free(all.a);
      return -(1 + i);
    }
  }
  /* sorted_by_key is stable; only the set of sequences behind each value is looked at, so any order by value will do */
  qsort(all.a, (size_t)all.n, sizeof(mzr_t), cmp_value);
  for (int64_t i = 0; i < (int64_t)n * n; ++i) out[i] = 0.0;
  uint32_t *hits = (uint32_t *)malloc((size_t)(all.n ? all.n : 1) * sizeof(uint32_t));
  int64_t l = 0, r = 0;
  while (l < all.n) {
    while (r < all.n && all.a[r].value == all.a[l].value) ++r;
    int64_t m = 0;
    for (int64_t i = l; i < r; ++i) hits[m++] = (uint32_t)(all.a[i].position >> 32);
    qsort(hits, (size_t)m, sizeof(uint32_t), cmp_u32);
    int64_t u = 0;
    for (int64_t i = 0; i < m; ++i)
      if (i == 0 || hits[i] != hits[i - 1]) hits[u++] = hits[i];
    for (int64_t i = 0; i < u; ++i)
      for (int64_t j = i; j < u; ++j) out[(int64_t)hits[i] * n + hits[j]] += 1.0;
    l = r;
  }
  free(hits), free(all.a);
  for (int i = 0; i < n; ++i) {
    if (!(out[(int64_t)i * n + i] > 0.)) return -(1 + i); /* "no self-hit found" */
    for (int j = i + 1; j < n; ++j) {
      out[(int64_t)i * n + j] = 1.0 - out[(int64_t)i * n + j] / out[(int64_t)i * n + i];
      out[(int64_t)j * n + i] = out[(int64_t)i * n + j];
    }
    out[(int64_t)i * n + i] = 0.0;
  }
  return 0;
}

/* ndarray 0.16.1 numeric_util::unrolled_fold with + over one contiguous lane */
static double lane_sum(const double *xs, int n) {
  double acc = 0.0, p[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  while (n >= 8) {
    for (int i = 0; i < 8; ++i) p[i] = p[i] + xs[i];
    xs += 8, n -= 8;
  }
  acc = acc + (p[0] + p[4]);
  acc = acc + (p[1] + p[5]);
  acc = acc + (p[2] + p[6]);
  acc = acc + (p[3] + p[7]);
  for (int i = 0; i < n && i < 7; ++i) acc = acc + xs[i];
  return acc;
}

/* neighbor_joining.rs:46-62: Q = (n - 2) D - sum_0 (broadcast along rows) - sum_1 (broadcast along columns), diagonal = inf.
 * D and Q are dense n x n, row-major. */
void orc_nj_q_matrix(const double *D, int n, double *Q) {
  double *s0 = (double *)calloc((size_t)n, sizeof(double)), *s1 = (double *)calloc((size_t)n, sizeof(double));
  for (int r = 0; r < n; ++r) /* sum_axis(Axis(0)): res = res + row, row after row */
    for (int c = 0; c < n; ++c) s0[c] = s0[c] + D[(int64_t)r * n + c];
  for (int r = 0; r < n; ++r) s1[r] = lane_sum(D + (int64_t)r * n, n); /* sum_axis(Axis(1)): one lane per row */
  const double f = (double)n - 2.0;
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < n; ++c) Q[(int64_t)r * n + c] = r == c ? INFINITY : (f * D[(int64_t)r * n + c] - s0[c]) - s1[r];
  free(s0), free(s1);
}

/* neighbor_joining.rs:73-80 */
void orc_nj_dist(const double *D, int n, int i, int j, double *dn) {
  for (int c = 0; c < n; ++c) dn[c] = 0.5 * ((D[(int64_t)i * n + c] + D[(int64_t)j * n + c]) - D[(int64_t)i * n + j]);
}

/* neighbor_joining.rs:16-35, 82-101.  Leaves are 0..n-1, the node made by the t-th join is n + t, the root is 2n - 2;
 * left[t], right[t] are the children of node n + t.  -> 0, -1 when n < 2 (the reference indexes nodes[1] and panics),
 * -2 when a Q matrix holds a NaN (argmin's UndefinedOrder). */
int orc_nj_tree(const double *D0, int n, int32_t *left, int32_t *right) {
  if (n < 2) return -1;
  double *D = (double *)malloc((size_t)n * n * sizeof(double)), *Q = (double *)malloc((size_t)n * n * sizeof(double));
  double *dn = (double *)malloc((size_t)n * sizeof(double));
  int32_t *nodes = (int32_t *)malloc((size_t)n * sizeof(int32_t));
  memcpy(D, D0, (size_t)n * n * sizeof(double));
  for (int i = 0; i < n; ++i) nodes[i] = i;
  int m = n, made = 0, rc = 0;
  while (m > 2) {
    orc_nj_q_matrix(D, m, Q);
    int bi = 0, bj = 0; /* argmin: first strictly smaller in row-major order, starting from element (0, 0) */
    double best = Q[0];
    for (int r = 0; r < m && !rc; ++r)
      for (int c = 0; c < m; ++c) {
        const double q = Q[(int64_t)r * m + c];
        if (isnan(q) || isnan(best)) {
          rc = -2;
          break;
        }
        if (q < best) best = q, bi = r, bj = c;
      }
    if (rc) break;
    int i = bi < bj ? bi : bj, j = bi < bj ? bj : bi;
    left[made] = nodes[i], right[made] = nodes[j];
    nodes[i] = n + made, ++made;
    memmove(nodes + j, nodes + j + 1, (size_t)(m - 1 - j) * sizeof(int32_t));
    orc_nj_dist(D, m, i, j, dn);
    for (int c = 0; c < m; ++c) D[(int64_t)i * m + c] = dn[c];
    for (int r = 0; r < m; ++r) D[(int64_t)r * m + i] = dn[r];
    D[(int64_t)i * m + i] = 0.0;
    /* remove row and column j */
    int64_t o = 0;
    for (int r = 0; r < m; ++r)
      if (r != j)
        for (int c = 0; c < m; ++c)
          if (c != j) Q[o++] = D[(int64_t)r * m + c];
    --m;
    memcpy(D, Q, (size_t)m * m * sizeof(double));
  }
  if (!rc) left[made] = nodes[0], right[made] = nodes[1];
  free(D), free(Q), free(dn), free(nodes);
  return rc;
}
