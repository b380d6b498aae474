// K4: the score fill of RMQ chaining on the GPU.
//
// Reference: the main loop of mg_lchain_rmq (minimap2/lchain.c:276-358) with comput_sc_simple (:232-248) and
// mg_log2 (mmpriv.h:118-126).  The reference walks the anchors in target order and keeps two balanced trees keyed by
// query position: one answers "smallest priority among the anchors at most max_dist behind on both axes", the other is
// walked backwards over the near neighbourhood (rmq_inner_dist) with the max_chain_skip early exit.
//
// What the trees hold is always a contiguous index window of the sorted anchor array, [st, i0) and [st_inner, i0),
// so no tree exists here: one warp owns one independent segment of the array (chain.h) and, per anchor,
//   * advances the two window starts with one ballot each (the eviction conditions are monotone in the index),
//   * scans the outer window for the smallest priority inside the query range (coalesced loads of y and of an
//     order-preserving 64-bit image of the double priority, two redux.sync to reduce),
//   * when the best predecessor is not an exact diagonal neighbour, gathers the near window, puts it in descending
//     (query position, index) order (usually it already is; otherwise a bitonic sort in shared memory) and evaluates
//     32 candidates at a time; the reference's sequential "skip" counter is replayed over two ballot masks.
// The anchor loop itself is sequential by nature (every score depends on earlier ones); parallelism comes from the
// 32 lanes inside a step and from the independent segments, queries and rounds in flight.
//
// The one thing a tree-free scan cannot reproduce is the reference's choice among EQUAL priorities inside an RMQ window,
// which follows the rotation history of its AVL tree.  Such a segment is flagged and handed back to the host
// (chain_fill_host); so are windows beyond kRing / kInnerCap.  Everything else is bit-identical.
#include "chain_fill.h"

#include <algorithm>
#include <cstring>

namespace pgmm {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int kScanUnroll = 8;

// Read once from global memory at kernel start: values that arrive as kernel parameters are re-fetched from the constant
// bank (and the shared-memory base / lane id re-derived from special registers) in front of almost every use, and on a
// one-warp dependent chain each of those fetches is exposed latency.  `zero` is always 0: adding it makes the
// shared-memory base and the lane id ordinary register values.
struct ChainDevParams {
  int max_dist, max_dist_inner, bw, max_skip, cap;
  float pen_gap, pen_skip;
  int zero;
  double half_pen;  // 0.5 * (double)pen_gap, the factor of lchain.c:285
};

// mmpriv.h:118-126, every operation rounded separately like the host build (-ffp-contract=off)
__device__ __forceinline__ float fast_log2_dev(float x) {
  uint32_t zi = __float_as_uint(x);
  float log_2 = (float)(int)(((zi >> 23) & 255) - 128);
  zi &= ~(255u << 23);
  zi += 127u << 23;
  const float zf = __uint_as_float(zi);
  float t = __fadd_rn(__fmul_rn(-0.34484843f, zf), 2.02466578f);
  t = __fadd_rn(__fmul_rn(t, zf), -0.67487759f);
  return __fadd_rn(log_2, t);
}

// lchain.c:232-248
__device__ __forceinline__ int link_score_dev(int xi, int yi, int xj, int yj, int q_span, float pen_gap, float pen_skip, bool &exact,
                                              int &width) {
  const int dq = yi - yj, dr = xi - xj;
  const int dd = dr > dq ? dr - dq : dq - dr, dg = dr < dq ? dr : dq;
  int sc = q_span < dg ? q_span : dg;
  width = dd;
  exact = dd == 0 && dg <= q_span;
  if (dd || dq > q_span) {
    const float lin_pen = __fadd_rn(__fmul_rn(pen_gap, (float)dd), __fmul_rn(pen_skip, (float)dg));
    const float log_pen = dd >= 1 ? fast_log2_dev((float)(dd + 1)) : 0.0f;
    sc -= (int)__fadd_rn(lin_pen, __fmul_rn(.5f, log_pen));
  }
  return sc;
}

// Everything about an anchor that depends on COORDINATES only is computed here, one thread per anchor, before the
// sequential fill: the split coordinates, the start of its target-position group (when the group's predecessors become
// visible, lchain.c:280-293), where both windows start once the evictions of lchain.c:295-312 have run (the conditions
// are monotone in the index, so each start is a lower bound found by binary search), and the score of linking it to
// the anchor right before it -- the predecessor the fill picks on most steps.
// AUX[i] = (group start, outer window start, inner window start, link score << 2 | exact << 1 | width <= bw).
__global__ void chain_prep_kernel(const U128 *__restrict__ a, int n, const int *__restrict__ seg_starts, int n_segs,
                                  const ChainDevParams *__restrict__ Pp, int *__restrict__ X, int *__restrict__ Y, uint8_t *__restrict__ QS,
                                  int4 *__restrict__ AUX, int *__restrict__ SEG) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const ChainDevParams P = *Pp;
  const U128 v = a[i];
  const int x = (int)(uint32_t)v.x, y = (int)(uint32_t)v.y;
  X[i] = x, Y[i] = y, QS[i] = (uint8_t)(v.y >> 32 & 0xff);
  int lo = 0, hi = n_segs;  // the segment this anchor belongs to
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (seg_starts[mid] <= i) lo = mid;
    else hi = mid;
  }
  const int s = seg_starts[lo];
  if (SEG) SEG[i] = lo;
  // first index in [s, i] whose target coordinate is at least `bound`
  const auto lower = [&](int bound) {
    int l = s, h = i;  // a[i] itself always qualifies
    while (l < h) {
      const int mid = (l + h) >> 1;
      if ((int)(uint32_t)a[mid].x >= bound) h = mid;
      else l = mid + 1;
    }
    return l;
  };
  int i0 = i, scp = 0;
  if (i > s) {
    const U128 pv = a[i - 1];
    const int px = (int)(uint32_t)pv.x;
    if (px == x) i0 = lower(x);
    bool exact;
    int width;
    const int sc = link_score_dev(x, y, px, (int)(uint32_t)pv.y, (int)(pv.y >> 32 & 0xff), P.pen_gap, P.pen_skip, exact, width);
    scp = sc * 4 + (exact ? 2 : 0) + (width <= P.bw ? 1 : 0);
  }
  // x - a[c].x <= d  <=>  a[c].x >= x - d (no overflow: coordinates are below 2^31 and d is a distance)
  const int st = min(i, max(lower(x - P.max_dist), i0 - P.cap));
  const int sti = P.max_dist_inner > 0 ? min(i, max(lower(x - P.max_dist_inner), i0 - P.cap)) : s;
  AUX[i] = make_int4(i0, st, sti, scp);
}

// Shared memory is addressed through 32-bit shared-space addresses computed once (plain pointers into dynamic shared
// memory made the compiler rebuild the shared window base from SR_CgaCtaId in front of every access).
__device__ __forceinline__ int lds32(unsigned a) {
  int v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ unsigned long long lds64(unsigned a) {
  unsigned long long v;
  asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts32(unsigned a, int v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts64(unsigned a, unsigned long long v) { asm volatile("st.shared.b64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); }

// segs[b] = (first anchor, end, first anchor of the query, unused); one warp per segment.
//
// Everything the loop WRITES and later reads back -- priority keys, scores, peak scores, predecessors, the walk's
// visit stamps -- lives in a shared-memory ring of kRing slots indexed by anchor number (the windows never reach further
// back than that, or the segment is handed to the host), so the dependent chain of one step never waits for L2.  The
// read-only coordinates stay in global memory (L1-resident after the first touch); the warp fetches its next 32
// anchors one batch ahead and broadcasts them by shuffle; results leave in coalesced batches of 32.
__global__ void __launch_bounds__(32) chain_fill_kernel(const int *__restrict__ X, const int *__restrict__ Y, const uint8_t *__restrict__ QS,
                                                        const int4 *__restrict__ AUX, const int4 *__restrict__ segs, const ChainDevParams *__restrict__ Pp,
                                                        int *__restrict__ F, int *__restrict__ PP, int *__restrict__ V, int *__restrict__ seg_flag) {
  constexpr int R = ChainEngine::kRing, M = R - 1;
  extern __shared__ unsigned long long smem_u64[];
  const ChainDevParams P = *Pp;
  const unsigned sm = (unsigned)__cvta_generic_to_shared(smem_u64) + (unsigned)P.zero;
  const unsigned a_pri = sm;                                   // [R] u64 priority keys
  const unsigned a_keys = a_pri + 8 * R;                       // [kInnerCap] u64 walk order
  const unsigned a_f = a_keys + 8 * ChainEngine::kInnerCap;    // [R] score
  const unsigned a_v = a_f + 4 * R;                            // [R] peak score
  const unsigned a_p = a_v + 4 * R;                            // [R] predecessor (batch-absolute index, -1 = none)
  const unsigned a_stamp = a_p + 4 * R;                        // [R] last step whose walk marked this anchor
  const unsigned a_stage = a_stamp + 4 * R;                    // [32][8] this batch's anchors: x, y, i0, st | sti, link, q_span, -
  const int lane = (int)threadIdx.x + P.zero;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int4 sg = segs[blockIdx.x];
  const int s = sg.x, e = sg.y, qbase = sg.z;
  const unsigned long long tr0 = trace::begin();
  const int max_dist = P.max_dist, max_dist_inner = P.max_dist_inner, bw = P.bw, max_skip = P.max_skip;
  const float pen_gap = P.pen_gap, pen_skip = P.pen_skip;
  const double half_pen = P.half_pen;
  int i0 = s, flag = ChainEngine::DONE;
  // smallest key of the visible window [st, i0) regardless of query position: its holder, the holder's query position,
  // whether the key is shared
  unsigned long long cb_key = ~0ull, last_key = ~0ull;
  int cb_j = -1, cb_y = 0;  // cb_j: holder, -1 = empty window, -2 = must be recomputed
  bool cb_tie = false;
  int last_y = 0, last_f = 0, last_v = 0;  // the previous anchor, still in registers
  for (int k = lane; k < R; k += 32) sts32(a_stamp + 4 * k, -1);
  __syncwarp();

  int nx = 0, ny = 0, nq = 0;
  int4 na = make_int4(0, 0, 0, 0);
  if (s + lane < e) nx = X[s + lane], ny = Y[s + lane], nq = QS[s + lane], na = AUX[s + lane];
  for (int ib = s; ib < e && flag == ChainEngine::DONE; ib += 32) {
    // stage this batch (one 32-byte record per anchor, read back as two broadcast 16-byte loads per step), fetch the next
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_stage + 32 * lane), "r"(nx), "r"(ny), "r"(na.x), "r"(na.y) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_stage + 32 * lane + 16), "r"(na.z), "r"(na.w), "r"(nq), "r"(ny - max_dist) : "memory");
    __syncwarp();
    if (ib + 32 + lane < e) nx = X[ib + 32 + lane], ny = Y[ib + 32 + lane], nq = QS[ib + 32 + lane], na = AUX[ib + 32 + lane];
    const int nb = min(32, e - ib);
    // records are loaded two steps ahead, so that nothing ever waits for the load it was just issued behind
    int4 ra, rb, ra2, rb2;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(ra.x), "=r"(ra.y), "=r"(ra.z), "=r"(ra.w) : "r"(a_stage));
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(rb.x), "=r"(rb.y), "=r"(rb.z), "=r"(rb.w) : "r"(a_stage + 16));
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(ra2.x), "=r"(ra2.y), "=r"(ra2.z), "=r"(ra2.w) : "r"(a_stage + 32));
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(rb2.x), "=r"(rb2.y), "=r"(rb2.z), "=r"(rb2.w) : "r"(a_stage + 48));
    for (int t = 0; t < nb; ++t) {
      const int i = ib + t;
      const int xi = ra.x, yi = ra.y, i0n = ra.z, st = ra.w, sti = rb.x, link = rb.y, qsi = rb.z, ylo = rb.w;
      ra = ra2, rb = rb2;
      {
        const unsigned nxt = a_stage + 32 * ((t + 2) & 31);
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(ra2.x), "=r"(ra2.y), "=r"(ra2.z), "=r"(ra2.w) : "r"(nxt));
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(rb2.x), "=r"(rb2.y), "=r"(rb2.z), "=r"(rb2.w) : "r"(nxt + 16));
      }
      int max_f = qsi, max_j = -1;
      if (i0n != i0) {  // the previous target position's anchors become visible together (:280-293)
        if (i0 == i - 1) {
          if (last_key < cb_key) cb_key = last_key, cb_j = i - 1, cb_y = last_y, cb_tie = false;
          else if (last_key == cb_key) cb_tie = true;
        } else {
#pragma unroll 1
          for (int jj = i0; jj < i; ++jj) {
            const unsigned long long k = lds64(a_pri + 8 * (jj & M));
            if (k < cb_key) cb_key = k, cb_j = jj, cb_y = Y[jj], cb_tie = false;
            else if (k == cb_key) cb_tie = true;
          }
        }
        i0 = i0n;
      }
      if (i - st >= R) {  // the ring no longer covers the window
        flag = ChainEngine::WINDOW;
        break;
      }
      // range minimum over [st, i0) with query position in (yi - max_dist, yi]; at yi itself only the query's first
      // anchor qualifies (the closed upper key is (yi, 0), :314).
      //
      // Fast path: the smallest key of the whole window, ignoring the query range, is carried from step to step (a newly
      // visible anchor is folded in; it is recomputed when its holder is evicted).  Along a chain that holder is the
      // newest anchor, and when it is unique and inside the query range it is the answer.  Otherwise the window is scanned.
      if ((unsigned)cb_j < (unsigned)st) cb_j = -2;  // the holder was evicted
      if (st >= i0) cb_tie = false, cb_j = -1, cb_key = ~0ull;
      int j = -1;
      bool fast = cb_j != -2 && !cb_tie;
      if (fast && cb_j >= 0) {
        if (cb_y > ylo && (cb_y < yi || (cb_y == yi && cb_j == qbase))) j = cb_j;
        else fast = false;
      }
      if (!fast) {  // branch-free scan, eight independent loads in flight; both the filtered and the unfiltered minimum
        unsigned long long bk = ~0ull, uk = ~0ull;
        int bj = -1, uj = -1;
        bool tie = false, utie = false;
#pragma unroll 1
        for (int jb = st + lane; jb < i0 + lane; jb += 32 * kScanUnroll) {
          int yv[kScanUnroll];
          unsigned long long kv[kScanUnroll];
#pragma unroll
          for (int u = 0; u < kScanUnroll; ++u) {
            const int jc = min(jb + 32 * u, i0 - 1);
            yv[u] = Y[jc], kv[u] = lds64(a_pri + 8 * (jc & M));
          }
#pragma unroll
          for (int u = 0; u < kScanUnroll; ++u) {
            const int jj = jb + 32 * u, y = yv[u];
            const bool in = jj < i0, ok = in && y > ylo && (y < yi || (y == yi && jj == qbase));
            const unsigned long long k0 = in ? kv[u] : ~0ull, k = ok ? k0 : ~0ull;
            tie = k < bk ? false : (ok && k == bk ? true : tie);
            bj = k < bk ? jj : bj;
            bk = k < bk ? k : bk;
            utie = k0 < uk ? false : (in && k0 == uk ? true : utie);
            uj = k0 < uk ? jj : uj;
            uk = k0 < uk ? k0 : uk;
          }
        }
        {  // refresh the carried minimum
          const unsigned hi = (unsigned)(uk >> 32);
          const unsigned mh = __reduce_min_sync(FULL, hi);
          const unsigned ml = __reduce_min_sync(FULL, hi == mh ? (unsigned)uk : 0xffffffffu);
          cb_key = (unsigned long long)mh << 32 | ml;
          const unsigned hm = __ballot_sync(FULL, uj >= 0 && uk == cb_key);
          cb_tie = __popc(hm) > 1 || __any_sync(FULL, utie && uk == cb_key);
          cb_j = hm ? __shfl_sync(FULL, uj, __ffs(hm) - 1) : -1;
          cb_y = cb_j >= 0 ? Y[cb_j] : 0;
        }
        const unsigned hi = (unsigned)(bk >> 32);
        const unsigned mh = __reduce_min_sync(FULL, hi);
        const unsigned ml = __reduce_min_sync(FULL, hi == mh ? (unsigned)bk : 0xffffffffu);
        const unsigned long long gk = (unsigned long long)mh << 32 | ml;
        if (gk != ~0ull) {
          const unsigned hm = __ballot_sync(FULL, bj >= 0 && bk == gk);
          if (__popc(hm) > 1 || __any_sync(FULL, tie && bk == gk)) {
            flag = ChainEngine::TIE;
            break;
          }
          j = __shfl_sync(FULL, bj, __ffs(hm) - 1);
        }
      }
      if (j >= 0) {
        bool exact, wok;
        int sc;
        if (j == i - 1) sc = last_f + (link >> 2), exact = link & 2, wok = link & 1;  // scored by the prep kernel
        else {
          int width;
          sc = lds32(a_f + 4 * (j & M)) + link_score_dev(xi, yi, X[j], Y[j], QS[j], pen_gap, pen_skip, exact, width);
          wok = width <= bw;
        }
        if (wok && sc > max_f) max_f = sc, max_j = j;
        if (!exact && max_dist_inner > 0 && i0 > sti && yi > 0) {
          // near neighbourhood (:319-348): members of [sti, i0) with query position in [yi - max_dist_inner, yi - 1],
          // visited in descending (query position, index) order
          const int y_hi = yi - 1, y_lo = yi - max_dist_inner;
          int cnt = 0;
#pragma unroll 1
          for (int base = i0 - 1; base >= sti; base -= 32) {
            const int j2 = base - lane;
            bool c = false;
            int y = 0;
            if (j2 >= sti) y = Y[j2], c = y >= y_lo && y <= y_hi;
            const unsigned m = __ballot_sync(FULL, c);
            const int pos = cnt + __popc(m & lt_mask);
            if (c && pos < ChainEngine::kInnerCap) sts64(a_keys + 8 * pos, (unsigned long long)(unsigned)y << 32 | (unsigned)(j2 - s));
            cnt += __popc(m);
          }
          if (cnt > ChainEngine::kInnerCap) {
            flag = ChainEngine::INNER;
            break;
          }
          __syncwarp();
          bool unsorted = false;
#pragma unroll 1
          for (int k = lane; k + 1 < cnt; k += 32) unsorted |= lds64(a_keys + 8 * k) < lds64(a_keys + 8 * k + 8);
          if (__any_sync(FULL, unsorted)) {
            int n2 = 32;
            while (n2 < cnt) n2 <<= 1;
            for (int k = cnt + lane; k < n2; k += 32) sts64(a_keys + 8 * k, 0);  // smallest: pads end up behind every member
            __syncwarp();
#pragma unroll 1
            for (int k = 2; k <= n2; k <<= 1)
#pragma unroll 1
              for (int d = k >> 1; d > 0; d >>= 1) {
#pragma unroll 1
                for (int w = lane; w < n2; w += 32) {
                  const int u = w ^ d;
                  if (u > w) {
                    const unsigned long long ka = lds64(a_keys + 8 * w), kb = lds64(a_keys + 8 * u);
                    const bool desc = (w & k) == 0;
                    if (desc ? ka < kb : ka > kb) sts64(a_keys + 8 * w, kb), sts64(a_keys + 8 * u, ka);
                  }
                }
                __syncwarp();
              }
          }
          int n_skip = 0;
#pragma unroll 1
          for (int c0 = 0; c0 < cnt; c0 += 32) {
            const int k = c0 + lane;
            int j2 = -1, sc2 = INT32_MIN;
            bool ok = false;
            if (k < cnt) {
              j2 = s + (int)(unsigned)lds64(a_keys + 8 * k);
              bool ex2;
              int w2;
              sc2 = lds32(a_f + 4 * (j2 & M)) + link_score_dev(xi, yi, X[j2], Y[j2], QS[j2], pen_gap, pen_skip, ex2, w2);
              ok = w2 <= bw;
              const int pj = lds32(a_p + 4 * (j2 & M));
              // "a predecessor of something already seen in this walk" (:344); one outside the near window is never visited
              if (ok && pj >= sti) sts32(a_stamp + 4 * (pj & M), i);
            }
            __syncwarp();
            // every writer of an anchor's stamp sits earlier in the visiting order (a predecessor has a smaller query position)
            const bool marked = ok && lds32(a_stamp + 4 * (j2 & M)) == i;
            // running maximum before each lane's turn -- only needed when some candidate beats the score so far
            int incl = INT32_MIN;
            unsigned um = 0, im;
            if (__any_sync(FULL, ok && sc2 > max_f)) {
              incl = ok ? sc2 : INT32_MIN;
#pragma unroll
              for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_up_sync(FULL, incl, d);
                if (lane >= d) incl = max(incl, o);
              }
              int excl = __shfl_up_sync(FULL, incl, 1);
              if (lane == 0) excl = INT32_MIN;
              excl = max(excl, max_f);
              const bool upd = ok && sc2 > excl;
              um = __ballot_sync(FULL, upd), im = __ballot_sync(FULL, ok && !upd && marked);
            } else im = __ballot_sync(FULL, marked);
            // the skip counter, event by event (:338-343)
            int brk = -1;
            if (um == 0) {
              const int room = max_skip - n_skip;  // increments that still fit
              if (__popc(im) > room) {
                unsigned m2 = im;
                for (int r = 0; r < room; ++r) m2 &= m2 - 1;
                brk = __ffs(m2) - 1, n_skip = max_skip + 1;  // increment number room+1 trips the limit
              } else n_skip += __popc(im);
            } else {
              unsigned evs = um | im;
              while (evs) {
                const int b = __ffs(evs) - 1;
                evs &= evs - 1;
                if (um >> b & 1) {
                  if (n_skip > 0) --n_skip;
                } else if (++n_skip > max_skip) {
                  brk = b;
                  break;
                }
              }
            }
            const int last = brk >= 0 ? brk : 31;
            const int best = um ? max(max_f, __shfl_sync(FULL, incl, last)) : max_f;
            if (best > max_f) {
              const unsigned wm = __ballot_sync(FULL, ok && sc2 == best);  // the first lane reaching it made the last update
              max_j = __shfl_sync(FULL, j2, __ffs(wm) - 1);
              max_f = best;
            }
            if (brk >= 0) break;
          }
        }
      }
      {
        const double sum = __dadd_rn((double)max_f, __dmul_rn(half_pen, (double)(xi + yi)));  // lchain.c:285: pri = -sum
        const long long b = __double_as_longlong(sum) ^ (long long)0x8000000000000000ull;
        last_key = b < 0 ? ~(unsigned long long)b : (unsigned long long)b | 0x8000000000000000ull;  // same order, unsigned
      }
      int vv = max_f;
      if (max_j >= 0) {
        const int vm = max_j == i - 1 ? last_v : lds32(a_v + 4 * (max_j & M));
        if (vm > max_f) vv = vm;
      }
      last_y = yi, last_f = max_f, last_v = vv;
      if (lane == 0) {
        const unsigned o = 4 * (i & M);
        sts32(a_f + o, max_f), sts32(a_p + o, max_j), sts32(a_v + o, vv);
        sts64(a_pri + 2 * o, last_key);
      }
      __syncwarp();
    }
    if (flag != ChainEngine::DONE) break;
    if (lane < nb) {  // this batch's results, coalesced
      const int i = ib + lane;
      const unsigned o = 4 * (i & M);
      const int pj = lds32(a_p + o);
      F[i] = lds32(a_f + o), V[i] = lds32(a_v + o), PP[i] = pj >= 0 ? pj - qbase : -1;
    }
  }
  if (lane == 0) seg_flag[blockIdx.x] = flag, trace::emit(5, tr0, tr0, (unsigned)(e - s));
}


// =====================================================================================================================
// K4p: the same score fill as a PARALLEL FIXED-POINT iteration (default; the one-warp-per-segment kernel above is kept
// as the PGMM_K4_SERIAL=1 variant).
//
// In mg_lchain_rmq the decision of anchor i (its predecessor p[i]) is a pure function of the coordinates and of (f, p)
// of EARLIER anchors; f and v then follow from p alone: f[i] = f[p[i]] + link(p[i], i) (q_span for a chain start),
// v[i] = max(v[p[i]], f[i]).  So the fill is the unique fixed point of
//     decide:    p'  = D(f, p)      -- every anchor at once, one warp each, reading the previous iterate
//     propagate: f,v = prefix sums / prefix maxima along the predecessor forest p'  (pointer jumping, log2(n) rounds)
// (unique by induction on the anchor index).  Starting from "every anchor links to the one right before it", a 5-Mbp
// query of the benchmark reaches the fixed point in 6 iterations and a real E. coli query in 15 -- instead of a dependent
// chain of 190 000 steps on one warp.  Everything below the first anchor of a segment whose decision changed in an
// iteration is final (its inputs are), so later iterations only touch what lies behind that frontier.
//
// decide answers the range minimum from block minima (32 anchors per block) plus the ragged ends of the window; when
// the smallest key of the whole window is unique and inside the query range it IS the reference's answer, otherwise the
// window is scanned with the range filter.  Equal smallest priorities inside the range still cannot be arbitrated
// without the reference's tree: the anchor is marked and its segment goes to the host arbiter, as before.
// =====================================================================================================================
struct ParArgs {
  const int *X, *Y;
  const uint8_t *QS;
  const int4 *AUX;      // i0, st, sti, link with the previous anchor
  const int *SEG;       // segment of every anchor
  const int4 *segs;     // per segment: start, end, first anchor of its query, -
  int *F, *P, *V, *Pn, *G;
  int *PX;              // exported predecessors (relative to the query's first anchor)
  unsigned long long *KEY;
  int4 *BM;             // per block of 32 anchors: key lo, key hi, holder (absolute index), tie
  int4 *JA, *JB;        // pointer-jumping state (ancestor, sum, prefix maximum, -), two copies
  int *frontier;        // [2][n_segs] first anchor whose decision changed in iteration t (t & 1), INT_MAX = none
  unsigned char *TIE;   // per anchor: the last evaluation found equal smallest priorities inside the range
  int *seg_flag;        // per segment: ChainEngine::INNER when a walk exceeded the device limits
  int *changed;         // [kMaxIter + 2] decisions changed per iteration ([0] preset to 1)
  int n, n_segs;
};

constexpr int kMaxIter = 64;        // iterations before the segments that still change go to the host arbiter
constexpr int kParWarps = 4;        // warps (= anchors in flight) per CTA of the decide kernel
constexpr int kMarkBits = 4096;     // anchors of a near window that can carry a walk mark

__device__ __forceinline__ unsigned long long pri_key(int f, int x, int y, double half_pen) {
  const double sum = __dadd_rn((double)f, __dmul_rn(half_pen, (double)(x + y)));  // lchain.c:285: pri = -sum
  const long long b = __double_as_longlong(sum) ^ (long long)0x8000000000000000ull;
  return b < 0 ? ~(unsigned long long)b : (unsigned long long)b | 0x8000000000000000ull;  // same order, unsigned
}

// iteration 0: every anchor whose left neighbour is visible and gives a positive link starts out chained to it
__global__ void par_init_kernel(ParArgs A, const ChainDevParams *__restrict__ Pp) {
  const int bw = Pp->bw;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < A.n; i += gridDim.x * blockDim.x) {
    const int4 aux = A.AUX[i];
    int p = -1, g = A.QS[i];
    const int link = aux.w;
    (void)bw;
    if (aux.x == i && aux.y < i && (link & 1) && (link >> 2) > 0) p = i - 1, g = link >> 2;
    A.Pn[i] = p, A.G[i] = g, A.P[i] = -2, A.TIE[i] = 0;
  }
}

// start of the propagation of iteration t: anchors below their segment's frontier keep their final values
__global__ void par_jump_init_kernel(ParArgs A, int t) {
  if (t > 0 && A.changed[t] == 0) return;  // nothing changed in this iteration: the iterate is the fixed point
  const int *fr = A.frontier + (size_t)(t & 1) * A.n_segs;
  int *fr_next = A.frontier + (size_t)((t + 1) & 1) * A.n_segs;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < A.n; i += gridDim.x * blockDim.x) {
    if (i < A.n_segs) fr_next[i] = INT32_MAX;
    const int sg = A.SEG[i];
    if (t > 0 && i < fr[sg]) A.JA[i] = make_int4(-1, A.F[i], A.V[i], 0);
    else {
      const int g = A.G[i];
      A.JA[i] = make_int4(A.Pn[i], g, g, 0);
    }
  }
}

// one round of pointer jumping: F[i] = F[J] + S, V[i] = max(V[J], F[J] + D)
__global__ void par_jump_kernel(ParArgs A, int t, int flip) {
  if (t > 0 && A.changed[t] == 0) return;  // nothing changed in this iteration: the iterate is the fixed point
  const int4 *src = flip ? A.JB : A.JA;
  int4 *dst = flip ? A.JA : A.JB;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < A.n; i += gridDim.x * blockDim.x) {
    int4 a = src[i];
    if (a.x >= 0) {
      const int4 b = src[a.x];
      a.z = max(b.z, b.y + a.z), a.y += b.y, a.x = b.x;
    }
    dst[i] = a;
  }
}

// end of iteration t: scores, peak scores, priority keys and the block minima the next decide reads
__global__ void par_finalize_kernel(ParArgs A, const ChainDevParams *__restrict__ Pp, int t, int flip) {
  if (t > 0 && A.changed[t] == 0) return;  // nothing changed in this iteration: the iterate is the fixed point
  const double half_pen = Pp->half_pen;
  const int4 *src = flip ? A.JB : A.JA;
  const int lane = threadIdx.x & 31;
  const int n32 = (A.n + 31) & ~31;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n32; i += gridDim.x * blockDim.x) {
    unsigned long long key = ~0ull;
    if (i < A.n) {
      const int4 a = src[i];
      A.F[i] = a.y, A.V[i] = a.z, A.P[i] = A.Pn[i];
      key = pri_key(a.y, A.X[i], A.Y[i], half_pen);
      A.KEY[i] = key;
    }
    const unsigned hi = (unsigned)(key >> 32);
    const unsigned mh = __reduce_min_sync(FULL, hi);
    const unsigned ml = __reduce_min_sync(FULL, hi == mh ? (unsigned)key : 0xffffffffu);
    const unsigned long long mk = (unsigned long long)mh << 32 | ml;
    const unsigned hm = __ballot_sync(FULL, key == mk && i < A.n);
    if (lane == 0) A.BM[i >> 5] = make_int4((int)ml, (int)mh, hm ? i + __ffs(hm) - 1 : -1, __popc(hm) > 1);
  }
}

// The decision of every anchor from the previous iterate (lchain.c:313-351), one warp per anchor.
__global__ void __launch_bounds__(kParWarps * 32) par_decide_kernel(ParArgs A, const ChainDevParams *__restrict__ Pp, int t) {
  if (A.changed[t - 1] == 0) return;  // the fixed point was reached
  __shared__ unsigned long long s_keys[kParWarps][ChainEngine::kInnerCap];
  __shared__ unsigned s_mark[kParWarps][kMarkBits / 32];
  const ChainDevParams P = *Pp;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  unsigned long long *keys = s_keys[wid];
  unsigned *mark = s_mark[wid];
  const int max_dist = P.max_dist, max_dist_inner = P.max_dist_inner, bw = P.bw, max_skip = P.max_skip;
  const float pen_gap = P.pen_gap, pen_skip = P.pen_skip;
  const int *fr_prev = A.frontier + (size_t)((t - 1) & 1) * A.n_segs;
  int *fr_cur = A.frontier + (size_t)(t & 1) * A.n_segs;
  const int *__restrict__ X = A.X, *__restrict__ Y = A.Y, *__restrict__ F = A.F, *__restrict__ PO = A.P;
  const uint8_t *__restrict__ QS = A.QS;
  const unsigned long long *__restrict__ KEY = A.KEY;
  int n_changed = 0;
  for (int i = blockIdx.x * kParWarps + wid; i < A.n; i += gridDim.x * kParWarps) {
    const int sgi = A.SEG[i];
    if (t > 1 && i <= fr_prev[sgi]) continue;  // final already (its inputs did not change in the last iteration)
    const int4 aux = A.AUX[i];
    const int i0 = aux.x, st = aux.y, sti = aux.z, link = aux.w;
    const int xi = X[i], yi = Y[i], qsi = QS[i], ylo = yi - max_dist;
    const int qbase = A.segs[sgi].z;
    int max_f = qsi, max_j = -1;
    int j = -1;
    bool tie_here = false;
    if (st < i0) {
      // ---- smallest key of the whole window from block minima and the ragged ends ----
      const int b0 = (st + 31) >> 5, b1 = i0 >> 5;
      const bool blocks = b0 < b1;
      const int nh = blocks ? (b0 << 5) - st : i0 - st, nb = blocks ? b1 - b0 : 0, nt = blocks ? i0 - (b1 << 5) : 0;
      unsigned long long uk = ~0ull;
      int uj = -1;
      bool utie = false;
      for (int c = lane; c < nh + nb + nt; c += 32) {
        unsigned long long k;
        int h;
        bool tf = false;
        if (c < nh) h = st + c, k = KEY[h];
        else if (c < nh + nb) {
          const int4 bm = A.BM[b0 + c - nh];
          k = (unsigned long long)(unsigned)bm.y << 32 | (unsigned)bm.x, h = bm.z, tf = bm.w != 0;
        } else h = (b1 << 5) + (c - nh - nb), k = KEY[h];
        if (k < uk) uk = k, uj = h, utie = tf;
        else if (k == uk) utie = true;
      }
      bool fast;
      {
        const unsigned hi = (unsigned)(uk >> 32);
        const unsigned mh = __reduce_min_sync(FULL, hi);
        const unsigned ml = __reduce_min_sync(FULL, hi == mh ? (unsigned)uk : 0xffffffffu);
        const unsigned long long gk = (unsigned long long)mh << 32 | ml;
        const unsigned hm = __ballot_sync(FULL, uj >= 0 && uk == gk);
        const bool gt = __popc(hm) > 1 || __any_sync(FULL, utie && uk == gk);
        const int gj = hm ? __shfl_sync(FULL, uj, __ffs(hm) - 1) : -1;
        fast = !gt && gj >= 0;
        if (fast) {
          const int yh = Y[gj];
          if (yh > ylo && (yh < yi || (yh == yi && gj == qbase))) j = gj;
          else fast = false;
        }
      }
      if (!fast) {  // range-filtered scan of the window, eight independent loads in flight
        unsigned long long bk = ~0ull;
        int bj = -1;
        bool tie = false;
#pragma unroll 1
        for (int jb = st + lane; jb < i0 + lane; jb += 32 * kScanUnroll) {
          int yv[kScanUnroll];
          unsigned long long kv[kScanUnroll];
#pragma unroll
          for (int u = 0; u < kScanUnroll; ++u) {
            const int jc = min(jb + 32 * u, i0 - 1);
            yv[u] = Y[jc], kv[u] = KEY[jc];
          }
#pragma unroll
          for (int u = 0; u < kScanUnroll; ++u) {
            const int jj = jb + 32 * u, y = yv[u];
            const bool ok = jj < i0 && y > ylo && (y < yi || (y == yi && jj == qbase));
            const unsigned long long k = ok ? kv[u] : ~0ull;
            tie = k < bk ? false : (ok && k == bk ? true : tie);
            bj = k < bk ? jj : bj;
            bk = k < bk ? k : bk;
          }
        }
        const unsigned hi = (unsigned)(bk >> 32);
        const unsigned mh = __reduce_min_sync(FULL, hi);
        const unsigned ml = __reduce_min_sync(FULL, hi == mh ? (unsigned)bk : 0xffffffffu);
        const unsigned long long gk = (unsigned long long)mh << 32 | ml;
        if (gk != ~0ull) {
          const unsigned hm = __ballot_sync(FULL, bj >= 0 && bk == gk);
          tie_here = __popc(hm) > 1 || __any_sync(FULL, tie && bk == gk);
          // a lane's bj is its first (lowest) holder; the lowest lane does not hold the lowest index, so take the minimum
          j = __reduce_min_sync(FULL, (bj >= 0 && bk == gk) ? bj : INT32_MAX);
        }
      }
    }
    bool overflow = false;
    if (j >= 0) {
      bool exact, wok;
      int sc;
      if (j == i - 1) sc = F[j] + (link >> 2), exact = link & 2, wok = link & 1;  // scored by the prep kernel
      else {
        int width;
        sc = F[j] + link_score_dev(xi, yi, X[j], Y[j], QS[j], pen_gap, pen_skip, exact, width);
        wok = width <= bw;
      }
      if (wok && sc > max_f) max_f = sc, max_j = j;
      if (!exact && max_dist_inner > 0 && i0 > sti && yi > 0) {
        // near neighbourhood (:319-348): members of [sti, i0) with query position in [yi - max_dist_inner, yi - 1],
        // visited in descending (query position, index) order
        const int y_hi = yi - 1, y_lo = yi - max_dist_inner;
        int cnt = 0;
        if (i0 - sti > kMarkBits) overflow = true;
#pragma unroll 1
        for (int base = i0 - 1; base >= sti && !overflow; base -= 32) {
          const int j2 = base - lane;
          bool c = false;
          int y = 0;
          if (j2 >= sti) y = Y[j2], c = y >= y_lo && y <= y_hi;
          const unsigned m = __ballot_sync(FULL, c);
          const int pos = cnt + __popc(m & lt_mask);
          if (c && pos < ChainEngine::kInnerCap) keys[pos] = (unsigned long long)(unsigned)y << 32 | (unsigned)j2;
          cnt += __popc(m);
        }
        if (cnt > ChainEngine::kInnerCap) overflow = true;
        if (!overflow) {
          for (int k = lane; k < (i0 - sti + 31) / 32; k += 32) mark[k] = 0;
          __syncwarp();
          bool unsorted = false;
#pragma unroll 1
          for (int k = lane; k + 1 < cnt; k += 32) unsorted |= keys[k] < keys[k + 1];
          if (__any_sync(FULL, unsorted)) {
            int n2 = 32;
            while (n2 < cnt) n2 <<= 1;
            for (int k = cnt + lane; k < n2; k += 32) keys[k] = 0;  // smallest: pads end up behind every member
            __syncwarp();
#pragma unroll 1
            for (int k = 2; k <= n2; k <<= 1)
#pragma unroll 1
              for (int d = k >> 1; d > 0; d >>= 1) {
#pragma unroll 1
                for (int w = lane; w < n2; w += 32) {
                  const int u = w ^ d;
                  if (u > w) {
                    const unsigned long long ka = keys[w], kb = keys[u];
                    const bool desc = (w & k) == 0;
                    if (desc ? ka < kb : ka > kb) keys[w] = kb, keys[u] = ka;
                  }
                }
                __syncwarp();
              }
          }
          int n_skip = 0;
#pragma unroll 1
          for (int c0 = 0; c0 < cnt; c0 += 32) {
            const int k = c0 + lane;
            int j2 = -1, sc2 = INT32_MIN;
            bool ok = false;
            if (k < cnt) {
              j2 = (int)(unsigned)keys[k];
              bool ex2;
              int w2;
              sc2 = F[j2] + link_score_dev(xi, yi, X[j2], Y[j2], QS[j2], pen_gap, pen_skip, ex2, w2);
              ok = w2 <= bw;
              const int pj = PO[j2];
              // "a predecessor of something already seen in this walk" (:344); one outside the near window is never visited
              if (ok && pj >= sti) atomicOr(&mark[(pj - sti) >> 5], 1u << ((pj - sti) & 31));
            }
            __syncwarp();
            // every writer of an anchor's mark sits earlier in the visiting order (a predecessor has a smaller query position)
            const bool marked = ok && (mark[(j2 - sti) >> 5] >> ((j2 - sti) & 31) & 1u);
            int incl = INT32_MIN;
            unsigned um = 0, im;
            if (__any_sync(FULL, ok && sc2 > max_f)) {
              incl = ok ? sc2 : INT32_MIN;
#pragma unroll
              for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_up_sync(FULL, incl, d);
                if (lane >= d) incl = max(incl, o);
              }
              int excl = __shfl_up_sync(FULL, incl, 1);
              if (lane == 0) excl = INT32_MIN;
              excl = max(excl, max_f);
              const bool upd = ok && sc2 > excl;
              um = __ballot_sync(FULL, upd), im = __ballot_sync(FULL, ok && !upd && marked);
            } else im = __ballot_sync(FULL, marked);
            int brk = -1;
            if (um == 0) {
              const int room = max_skip - n_skip;
              if (__popc(im) > room) {
                unsigned m2 = im;
                for (int r = 0; r < room; ++r) m2 &= m2 - 1;
                brk = __ffs(m2) - 1, n_skip = max_skip + 1;
              } else n_skip += __popc(im);
            } else {
              unsigned evs = um | im;
              while (evs) {
                const int b = __ffs(evs) - 1;
                evs &= evs - 1;
                if (um >> b & 1) {
                  if (n_skip > 0) --n_skip;
                } else if (++n_skip > max_skip) {
                  brk = b;
                  break;
                }
              }
            }
            const int last = brk >= 0 ? brk : 31;
            const int best = um ? max(max_f, __shfl_sync(FULL, incl, last)) : max_f;
            if (best > max_f) {
              const unsigned wm = __ballot_sync(FULL, ok && sc2 == best);
              max_j = __shfl_sync(FULL, j2, __ffs(wm) - 1);
              max_f = best;
            }
            if (brk >= 0) break;
            __syncwarp();
          }
        }
      }
    }
    if (lane == 0) {
      A.TIE[i] = tie_here ? 1 : 0;
      if (overflow) atomicMax(&A.seg_flag[sgi], (int)ChainEngine::INNER);
      A.Pn[i] = max_j;
      A.G[i] = max_j >= 0 ? max_f - F[max_j] : max_f;
      if (max_j != PO[i]) {
        atomicMin(&fr_cur[sgi], i);
        ++n_changed;
      }
    }
  }
  if (lane == 0 && n_changed) atomicAdd(&A.changed[t], n_changed);
}

// results in the caller's terms: predecessors relative to the query's first anchor, ties folded into the segment flags
__global__ void par_export_kernel(ParArgs A, int t_last) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < A.n; i += gridDim.x * blockDim.x) {
    const int sg = A.SEG[i];
    const int p = A.P[i];
    A.PX[i] = p >= 0 ? p - A.segs[sg].z : -1;
    if (A.TIE[i]) atomicMax(&A.seg_flag[sg], (int)ChainEngine::TIE);
    // a segment whose decisions still changed in the last iteration that ran has not converged: the host fills it
    if (A.frontier[(size_t)(t_last & 1) * A.n_segs + sg] != INT32_MAX && i == A.segs[sg].x) atomicMax(&A.seg_flag[sg], (int)ChainEngine::WINDOW);
  }
}

}  // namespace

void trace_attach_chain(PgmmCtaTraceRec *buf, unsigned long long *cnt, unsigned long long cap) { trace::attach(buf, cnt, cap); }

constexpr size_t kFillSmem = (size_t)ChainEngine::kRing * (8 + 4 * 4) + (size_t)ChainEngine::kInnerCap * 8 + 32 * 32;
static_assert(kFillSmem <= 227 * 1024, "the ring must fit the shared memory of one SM");
static_assert((ChainEngine::kRing & (ChainEngine::kRing - 1)) == 0, "ring slots are addressed with a mask");

ChainEngine::ChainEngine() {
  static const bool once = [] {
    PGMM_CUDA(cudaFuncSetAttribute(chain_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
    return true;
  }();
  (void)once;
  PGMM_CUDA(cudaEventCreate(&ev0_));
  PGMM_CUDA(cudaEventCreate(&ev1_));
}
ChainEngine::~ChainEngine() {
  if (ev0_) cudaEventDestroy(ev0_);
  if (ev1_) cudaEventDestroy(ev1_);
}

void ChainEngine::run(const ChainParams &cp, std::vector<ChainFillJob> &jobs, cudaStream_t st, ChainFillStats *stats) {
  size_t n_total = 0, n_segs = 0;
  for (const ChainFillJob &j : jobs) n_total += (size_t)j.n, n_segs += j.segs.size();
  if (n_total == 0) return;
  if (n_total >= (size_t)INT32_MAX) PGMM_FATAL("chain fill: %zu anchors in one batch exceed the 31-bit index space", n_total);
  static const bool serial = getenv("PGMM_K4_SERIAL") != nullptr && atoi(getenv("PGMM_K4_SERIAL")) != 0;
  ChainDevParams P;
  P.max_dist = cp.max_dist < cp.bw ? cp.bw : cp.max_dist;
  P.max_dist_inner = (cp.max_dist_inner <= 0 || cp.max_dist_inner >= P.max_dist) ? 0 : cp.max_dist_inner;
  P.bw = cp.bw, P.max_skip = cp.max_chn_skip, P.cap = cp.cap_rmq_size, P.pen_gap = cp.pen_gap, P.pen_skip = cp.pen_skip;
  P.zero = 0, P.half_pen = 0.5 * (double)cp.pen_gap;

  // staging: anchors query after query.  Segment list 1 (serial kernel: longest first so that the long ones start
  // first), the kernels' parameter block, the segment starts in ascending order, segment list 2 (ascending, what SEG[]
  // indexes: start, end, first anchor of the query)
  U128 *ha = h_a_.ensure(n_total);
  const size_t n_start4 = (n_segs + 3) / 4;
  const size_t o_par = n_segs, o_starts = n_segs + 3, o_asc = n_segs + 3 + n_start4, n_hs = o_asc + n_segs;
  int4 *hs = h_segs_.ensure(n_hs);
  static_assert(sizeof(ChainDevParams) <= 3 * sizeof(int4), "parameter block");
  struct Ref {
    int job, seg;
    int64_t len;
  };
  std::vector<Ref> order;
  order.reserve(n_segs);
  std::vector<size_t> base(jobs.size());
  size_t off = 0;
  int64_t seg_max = 1;
  for (size_t q = 0; q < jobs.size(); ++q) {
    ChainFillJob &j = jobs[q];
    base[q] = off;
    if (j.n) memcpy(ha + off, j.a, (size_t)j.n * sizeof(U128));
    off += (size_t)j.n;
    j.redo.assign(j.segs.size(), 0);
    for (size_t k = 0; k < j.segs.size(); ++k) {
      order.push_back(Ref{(int)q, (int)k, j.segs[k].end - j.segs[k].start});
      seg_max = std::max(seg_max, j.segs[k].end - j.segs[k].start);
    }
  }
  std::vector<Ref> asc = order;  // (query, segment) in ascending anchor order
  if (serial) std::stable_sort(order.begin(), order.end(), [](const Ref &a, const Ref &b) { return a.len > b.len; });
  for (size_t k = 0; k < n_segs; ++k) {
    const Ref &r = order[k];
    const ChainSeg &sg = jobs[r.job].segs[r.seg];
    hs[k] = make_int4((int)(base[r.job] + sg.start), (int)(base[r.job] + sg.end), (int)base[r.job], 0);
    const Ref &ra = asc[k];
    const ChainSeg &sa = jobs[ra.job].segs[ra.seg];
    hs[o_asc + k] = make_int4((int)(base[ra.job] + sa.start), (int)(base[ra.job] + sa.end), (int)base[ra.job], 0);
  }
  d_a_.ensure(n_total), d_x_.ensure(n_total), d_y_.ensure(n_total), d_qs_.ensure(n_total), d_f_.ensure(3 * n_total);
  d_segs_.ensure(n_hs), d_flag_.ensure(n_segs), d_aux_.ensure(n_total);
  memcpy(hs + o_par, &P, sizeof(P));
  {
    int *starts = (int *)(hs + o_starts);
    size_t k = 0;
    for (size_t q = 0; q < jobs.size(); ++q)
      for (const ChainSeg &sg : jobs[q].segs) starts[k++] = (int)(base[q] + sg.start);  // ascending by construction
  }
  int32_t *dF = d_f_.p, *dP = d_f_.p + n_total, *dV = d_f_.p + 2 * n_total;
  PGMM_CUDA(cudaMemcpyAsync(d_a_.p, ha, n_total * sizeof(U128), cudaMemcpyHostToDevice, st));
  PGMM_CUDA(cudaMemcpyAsync(d_segs_.p, hs, n_hs * sizeof(int4), cudaMemcpyHostToDevice, st));
  PGMM_CUDA(cudaEventRecord(ev0_, st));
  const ChainDevParams *dP_ = (const ChainDevParams *)(d_segs_.p + o_par);
  int launches = 0, iters = 0;
  if (serial) {
    chain_prep_kernel<<<(unsigned)((n_total + 255) / 256), 256, 0, st>>>(d_a_.p, (int)n_total, (const int *)(d_segs_.p + o_starts), (int)n_segs, dP_,
                                                                         d_x_.p, d_y_.p, d_qs_.p, d_aux_.p, nullptr);
    // PGMM_K4_SMEM_KB: shared memory a fill warp asks for at least (a large value keeps other CTAs off its SM)
    static const size_t fill_smem = getenv("PGMM_K4_SMEM_KB") ? std::max(kFillSmem, (size_t)atoi(getenv("PGMM_K4_SMEM_KB")) * 1024) : kFillSmem;
    chain_fill_kernel<<<(unsigned)n_segs, 32, fill_smem, st>>>(d_x_.p, d_y_.p, d_qs_.p, d_aux_.p, d_segs_.p, dP_, dF, dP, dV, d_flag_.p);
    PGMM_CUDA(cudaGetLastError());
    launches = 2;
  } else {
    // ---- K4p: parallel fixed-point iteration ----
    const size_t n = n_total, nb = (n + 31) / 32;
    // one slab: SEG, P, Pn, G (int n each) | KEY (u64 n) | BM (int4 nb) | JA, JB (int4 n each) | frontier (2 n_segs) |
    // changed (kMaxIter + 2) | TIE (n bytes)
    const size_t words = 4 * n + 2 * n + 4 * nb + 8 * n + 2 * n_segs + (size_t)kMaxIter + 2 + (n + 3) / 4 + 64;
    int *slab = d_par_.ensure(words);
    ParArgs A;
    A.X = d_x_.p, A.Y = d_y_.p, A.QS = d_qs_.p, A.AUX = d_aux_.p, A.segs = d_segs_.p + o_asc;
    size_t w = 0;
    int *SEG = slab + w;
    w += n;
    A.SEG = SEG, A.P = slab + w, w += n, A.Pn = slab + w, w += n, A.G = slab + w, w += n;
    w = (w + 3) & ~(size_t)3;
    A.KEY = (unsigned long long *)(slab + w), w += 2 * n;
    w = (w + 3) & ~(size_t)3;
    A.BM = (int4 *)(slab + w), w += 4 * nb;
    A.JA = (int4 *)(slab + w), w += 4 * n;
    A.JB = (int4 *)(slab + w), w += 4 * n;
    A.frontier = slab + w, w += 2 * n_segs;
    A.changed = slab + w, w += (size_t)kMaxIter + 2;
    A.TIE = (unsigned char *)(slab + w);
    A.F = dF, A.V = dV, A.PX = dP, A.seg_flag = d_flag_.p;
    A.n = (int)n, A.n_segs = (int)n_segs;
    int rounds = 0;
    while ((int64_t(1) << rounds) < seg_max) ++rounds;  // pointer-jumping rounds: 2^rounds >= longest possible chain
    const unsigned g1 = (unsigned)std::min<size_t>((n + 255) / 256, 148 * 8);
    const unsigned gd = (unsigned)std::min<size_t>((n + kParWarps - 1) / kParWarps, 148 * 32);
    chain_prep_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_a_.p, (int)n, (const int *)(d_segs_.p + o_starts), (int)n_segs, dP_, d_x_.p, d_y_.p,
                                                                   d_qs_.p, d_aux_.p, SEG);
    PGMM_CUDA(cudaMemsetAsync(d_flag_.p, 0, n_segs * sizeof(int), st));
    PGMM_CUDA(cudaMemsetAsync(A.changed, 0, ((size_t)kMaxIter + 2) * sizeof(int), st));
    PGMM_CUDA(cudaMemsetAsync(A.frontier, 0x7f, 2 * n_segs * sizeof(int), st));
    {
      const int one = 1;  // changed[0] = 1: the first decide always runs
      h_changed_.ensure(kMaxIter + 2)[0] = one;
      PGMM_CUDA(cudaMemcpyAsync(A.changed, h_changed_.p, sizeof(int), cudaMemcpyHostToDevice, st));
    }
    par_init_kernel<<<g1, 256, 0, st>>>(A, dP_);
    launches += 2;
    const auto propagate = [&](int t) {
      par_jump_init_kernel<<<g1, 256, 0, st>>>(A, t);
      for (int r = 0; r < rounds; ++r) par_jump_kernel<<<g1, 256, 0, st>>>(A, t, r & 1);
      par_finalize_kernel<<<g1, 256, 0, st>>>(A, dP_, t, rounds & 1);
      launches += rounds + 2;
    };
    propagate(0);
    int t = 0, t_last = 0;
    int *hc = h_changed_.p;
    while (t < kMaxIter) {
      const int t_to = std::min(kMaxIter, t + (t == 0 ? 8 : 8));
      for (int k = t + 1; k <= t_to; ++k) {
        par_decide_kernel<<<gd, kParWarps * 32, 0, st>>>(A, dP_, k);
        ++launches;
        propagate(k);
      }
      PGMM_CUDA(cudaGetLastError());
      PGMM_CUDA(cudaMemcpyAsync(hc, A.changed, ((size_t)kMaxIter + 2) * sizeof(int), cudaMemcpyDeviceToHost, st));
      PGMM_CUDA(cudaStreamSynchronize(st));
      t_last = t_to;
      bool done = false;
      for (int k = t + 1; k <= t_to; ++k)
        if (hc[k] == 0) {
          t_last = k, done = true;
          break;
        }
      t = t_to;
      if (done) break;
    }
    iters = t_last;
    par_export_kernel<<<g1, 256, 0, st>>>(A, t_last);
    ++launches;
    PGMM_CUDA(cudaGetLastError());
  }
  PGMM_CUDA(cudaEventRecord(ev1_, st));
  int32_t *hfpv = h_fpv_.ensure(3 * n_total);
  int32_t *hflag = h_flag_.ensure(n_segs);
  PGMM_CUDA(cudaMemcpyAsync(hfpv, d_f_.p, 3 * n_total * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  PGMM_CUDA(cudaMemcpyAsync(hflag, d_flag_.p, n_segs * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  PGMM_CUDA(cudaStreamSynchronize(st));
  for (size_t q = 0; q < jobs.size(); ++q) {
    ChainFillJob &j = jobs[q];
    j.f = hfpv + base[q], j.p = hfpv + n_total + base[q], j.v = hfpv + 2 * n_total + base[q];
  }
  uint64_t redo_segs = 0, redo_anchors = 0;
  const std::vector<Ref> &flag_order = serial ? order : asc;
  for (size_t k = 0; k < n_segs; ++k)
    if (hflag[k] != DONE) {
      jobs[flag_order[k].job].redo[flag_order[k].seg] = (uint8_t)hflag[k];
      ++redo_segs, redo_anchors += (uint64_t)flag_order[k].len;
    }
  if (stats) {
    float ms = 0;
    PGMM_CUDA(cudaEventElapsedTime(&ms, ev0_, ev1_));
    stats->anchors += n_total, stats->segments += n_segs, stats->redo_segments += redo_segs, stats->redo_anchors += redo_anchors;
    stats->launches += launches, stats->kernel_ms += ms, stats->iterations += (uint64_t)iters, stats->batches += 1;
  }
}

}  // namespace pgmm
