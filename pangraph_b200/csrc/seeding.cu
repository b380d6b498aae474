// K1-K3 on the GPU (see seeding.h).  All stages are streaming integer work over arrays in HBM: one thread per base,
// per k-mer event, per minimizer, per seed or per anchor; ordering-sensitive outputs are produced by prefix sums and
// order-preserving scatters, so the arrays that reach the host are element-for-element what the reference builds.
//
// Exactness notes (reference = packages/minimap2-sys/minimap2/):
//  * mm_sketch (sketch.c:77-143) is a sequential scan with a ring of w slots.  Its state after any position is a
//    pure function of the last w "ring events" (every position except a skipped palindromic k-mer): the running
//    minimum is always the NEWEST of the smallest values in the window.  So every event can decide on its own which
//    elements the reference would emit at that step (first-window duplicates, displaced minimum, minimum leaving
//    the window + duplicates of its successor, final minimum).  Each element is emitted at most once and emission
//    order equals position order, hence a flag array + ordered compaction reproduces the output array.
//  * k-mer registers are not cleared by an ambiguous base, only the run length is; the palindrome test therefore
//    looks at the last k unambiguous bases, walking over ambiguous ones.
#include "seeding.h"

#include <atomic>
#include <cub/cub.cuh>

#include <algorithm>
#include <cstring>

namespace pgmm {

uint64_t g_seed_launches = 0;
std::atomic<uint64_t> g_anchor_sorted_device{0}, g_anchor_sorted_host{0};  // queries whose anchor order came from the device / from the host replay

namespace {

constexpr uint64_t U64MAX = ~0ull;
constexpr int TPB = 256;
inline unsigned nblk(uint64_t n) { return (unsigned)((n + TPB - 1) / TPB); }

struct SeqView {
  const uint8_t *codes;
  const uint64_t *starts;  // [n] first base of each sequence in codes
  const uint64_t *vstart;  // [n+1] first position of each sequence in the virtual concatenation
  int n;
  uint64_t total;
};

// largest s with v[s] <= p  (v[0] = 0, v ascending, p < v[n])
__device__ __forceinline__ int upper_seq(const uint64_t *v, int n, uint64_t p) {
  int lo = 0, hi = n;  // invariant: v[lo] <= p < v[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (v[mid] <= p) lo = mid;
    else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ uint64_t hash64(uint64_t key, uint64_t mask) {  // sketch.c:28-38
  key = (~key + (key << 21)) & mask;
  key = key ^ key >> 24;
  key = ((key + (key << 3)) + (key << 8)) & mask;
  key = key ^ key >> 14;
  key = ((key + (key << 2)) + (key << 4)) & mask;
  key = key ^ key >> 28;
  key = (key + (key << 31)) & mask;
  return key;
}

// per base: forward / reverse k-mer over the last k unambiguous bases of the sequence, palindrome test, hash
__global__ void kmer_kernel(SeqView v, int k, uint8_t *__restrict__ ev, uint64_t *__restrict__ hv, int32_t *__restrict__ rmark) {
  const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= v.total) return;
  const int s = upper_seq(v.vstart, v.n, p);
  const uint64_t first = v.vstart[s];
  const uint8_t *seq = v.codes + v.starts[s];
  const int64_t i = (int64_t)(p - first);
  const int c = seq[i];
  int32_t mark = p == first ? (int32_t)first - 1 : INT32_MIN;
  uint8_t is_event = 1;
  uint64_t h = 0;
  if (c < 4) {
    uint64_t k0 = 0, k1 = 0;
    int m = 0;
    for (int64_t j = i; j >= 0 && m < k; --j) {
      const int cj = seq[j];
      if (cj < 4) {
        k0 |= (uint64_t)cj << (2 * m);
        k1 |= (uint64_t)(3 ^ cj) << (2 * (k - 1 - m));
        ++m;
      }
    }
    if (k0 == k1) is_event = 0;  // "symmetric k-mer": the reference skips it before touching the ring (sketch.c:108)
    else {
      const int z = k0 < k1 ? 0 : 1;
      const uint64_t mask = (1ULL << 2 * k) - 1;
      h = hash64(z ? k1 : k0, mask) << 1 | (uint64_t)z;
    }
  } else mark = (int32_t)p;
  ev[p] = is_event, hv[p] = h, rmark[p] = mark;
}

// per base that is a ring event: its slot in the event stream
__global__ void event_kernel(SeqView v, int k, const uint8_t *__restrict__ ev, const uint64_t *__restrict__ hv,
                             const uint32_t *__restrict__ ev_incl, const int32_t *__restrict__ rpos,
                             uint64_t *__restrict__ EX, uint64_t *__restrict__ EP, int32_t *__restrict__ EL) {
  const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= v.total || !ev[p]) return;
  const int s = upper_seq(v.vstart, v.n, p);
  const int c = v.codes[v.starts[s] + (p - v.vstart[s])];
  const uint32_t e = ev_incl[p] - 1;
  int32_t l = 0;
  if (c < 4) {
    const int32_t r = rpos[p];
    l = (int32_t)(ev_incl[p] - (r >= 0 ? ev_incl[r] : 0));  // non-palindromic k-mers since the last ambiguous base
  }
  const uint64_t h = hv[p];
  EX[e] = (c < 4 && l >= k) ? ((h >> 1) << 8 | (uint64_t)k) : U64MAX;
  EP[e] = p << 1 | (h & 1);
  EL[e] = l;
}

// per ring event: mark what the reference's scan would append at this step (sketch.c:116-142)
__global__ void select_kernel(SeqView v, int w, int k, uint32_t n_ev, const uint32_t *__restrict__ ev_incl,
                              const uint64_t *__restrict__ EX, const uint64_t *__restrict__ EP,
                              const int32_t *__restrict__ EL, uint32_t *__restrict__ flag) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_ev) return;
  const uint64_t p = EP[e] >> 1;
  const int s = upper_seq(v.vstart, v.n, p);
  const int64_t E0 = v.vstart[s] > 0 ? ev_incl[v.vstart[s] - 1] : 0;  // first event of this sequence
  const int64_t Eend = ev_incl[v.vstart[s + 1] - 1];                   // one past its last event
  const int64_t ee = e;
  // newest-of-the-smallest over [e-w, e-1] (before this step) and over [e-w+1, e] (after it)
  int64_t pm = -1, nm = -1;
  uint64_t xpm = U64MAX, xnm = U64MAX;
  for (int64_t j = max(E0, ee - w); j < ee; ++j) {
    const uint64_t x = EX[j];
    if (xpm >= x) xpm = x, pm = j;
    if (j > ee - w && xnm >= x) xnm = x, nm = j;
  }
  const uint64_t xe = EX[e];
  if (xnm >= xe) xnm = xe, nm = ee;
  const int32_t le = EL[e];
  if (le == w + k - 1 && xpm != U64MAX)  // first full window: duplicates of the current minimum
    for (int64_t j = max(E0, ee - w + 1); j < ee; ++j)
      if (EX[j] == xpm && j != pm) flag[j] = 1;
  if (xe <= xpm) {  // a new minimum displaces the old one
    if (le >= w + k && xpm != U64MAX) flag[pm] = 1;
  } else if (pm == ee - w) {  // the old minimum leaves the window
    if (le >= w + k - 1 && xpm != U64MAX) flag[pm] = 1;
    if (le >= w + k - 1 && xnm != U64MAX)
      for (int64_t j = max(E0, ee - w + 1); j <= ee; ++j)
        if (EX[j] == xnm && j != nm) flag[j] = 1;
  }
  if (ee == Eend - 1 && xnm != U64MAX) flag[nm] = 1;  // the minimum standing at the end of the sequence
}

__global__ void scatter_mz_kernel(SeqView v, uint32_t n_ev, const uint32_t *__restrict__ flag, const uint32_t *__restrict__ fpos,
                                  const uint64_t *__restrict__ EX, const uint64_t *__restrict__ EP,
                                  uint64_t *__restrict__ mx, uint64_t *__restrict__ my) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_ev || !flag[e]) return;
  const uint64_t p = EP[e] >> 1;
  const int s = upper_seq(v.vstart, v.n, p);
  const uint32_t o = fpos[e];
  mx[o] = EX[e];
  my[o] = (uint64_t)s << 32 | (uint64_t)(uint32_t)(p - v.vstart[s]) << 1 | (EP[e] & 1);
}

__global__ void mz_offsets_kernel(SeqView v, uint32_t n_ev, const uint32_t *__restrict__ ev_incl, const uint32_t *__restrict__ fpos,
                                  uint32_t n_flag_total, uint64_t *__restrict__ mz_off) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s > v.n) return;
  if (s == v.n) {
    mz_off[s] = n_flag_total;
    return;
  }
  const uint64_t E0 = v.vstart[s] > 0 ? ev_incl[v.vstart[s] - 1] : 0;
  mz_off[s] = E0 < n_ev ? fpos[E0] : n_flag_total;
}


// ---------------------------------------------------------------------------------------------------------------
// K1 fused (odd k, the case of every pangraph preset: -k 19): sketch of one tile of one sequence in ONE kernel.
// For odd k no k-mer equals its reverse complement, so every position is a ring event and event index = position: the
// event stream of the general path above collapses onto the sequence itself and a tile plus halo holds everything a
// decision needs.  A CTA stages the tile's bases with 128-bit loads, every thread rolls the two k-mers over its stretch of
// positions (one shift / mask per base instead of a k-base walk back per position), hashes, and leaves key and run length
// in shared memory; the second phase takes the select_kernel decisions (same rules, sketch.c:116-142) from shared memory;
// the third compacts the tile's minimizers in order (block scan) to the front of the tile's own slice of a position-sized
// buffer and publishes their number.  A scan over the tile counts and one gather kernel put the tiles end to end.
// Three launches and one host synchronisation instead of twelve and four.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kTile = 4096, kTileThreads = 256, kTileItems = kTile / kTileThreads;

struct SketchTile {
  uint32_t seq, t0;  // sequence and first position of the tile inside it
};

__global__ void __launch_bounds__(kTileThreads) sketch_tile_kernel(SeqView v, const SketchTile *__restrict__ tiles, int w, int k,
                                                                    uint64_t *__restrict__ TX, uint64_t *__restrict__ TY,
                                                                    uint32_t *__restrict__ counts) {
  extern __shared__ __align__(16) uint8_t sk_smem[];
  const SketchTile tl = tiles[blockIdx.x];
  const int tid = threadIdx.x;
  const int64_t L = (int64_t)(v.vstart[tl.seq + 1] - v.vstart[tl.seq]);
  const uint8_t *seq = v.codes + v.starts[tl.seq];
  const int64_t t0 = tl.t0, t1 = min(L, t0 + (int64_t)kTile);
  // events (= positions) whose key is needed: the tile, w before it (windows) and w after it (later events flag tile positions)
  const int64_t e_lo = max((int64_t)0, t0 - w), e_hi = min(L, t1 + w);
  const int n_ev = (int)(e_hi - e_lo);
  // bases needed: back to where the run length of the first event is decided (w + k + 1 positions)
  const int64_t c_lo = max((int64_t)0, e_lo - (w + k + 1));
  const int n_codes = (int)(e_hi - c_lo);
  // shared memory (every part starts on a 16-byte boundary): keys [kTile + 2w], run length | strand << 15 [kTile + 2w],
  // flags [kTile + 2w], bases [kTile + 3w + k + 48]
  const int cap_ev = kTile + 2 * w;
  uint64_t *X = (uint64_t *)sk_smem;
  uint16_t *EL = (uint16_t *)(sk_smem + (size_t)cap_ev * 8);
  uint8_t *F = sk_smem + (size_t)cap_ev * 8 + ((size_t)cap_ev * 2 + 15) / 16 * 16;
  uint8_t *CDraw = F + ((size_t)cap_ev + 15) / 16 * 16;

  // ---- stage the bases with 128-bit loads: from the 16-byte boundary at or before the first base needed (the code
  // buffers are cudaMalloc blocks, so that boundary lies inside them); CD[i] is the base at position c_lo + i ----
  const uint8_t *src = seq + c_lo;
  const int pre = (int)((uintptr_t)src & 15);
  const uint8_t *CD = CDraw + pre;
  {
    const int n16 = (pre + n_codes) / 16;
    const uint4 *src16 = (const uint4 *)(src - pre);
    for (int i = tid; i < n16; i += kTileThreads) ((uint4 *)CDraw)[i] = __ldg(src16 + i);
    for (int i = 16 * n16 + tid; i < pre + n_codes; i += kTileThreads) CDraw[i] = src[i - pre];
    for (int i = tid; i < cap_ev; i += kTileThreads) F[i] = 0;
  }
  __syncthreads();

  // ---- phase 1: key and run length of every event in [e_lo, e_hi) ----
  {
    const int per = (n_ev + kTileThreads - 1) / kTileThreads;
    const int a0 = tid * per, a1 = min(n_ev, a0 + per);
    if (a0 < a1) {
      const uint64_t mask = (1ull << 2 * k) - 1;
      const int shift1 = 2 * (k - 1), cap_l = w + k + 1;
      // run length of unambiguous bases ending right before the first event of this stretch (exact up to w + k + 1)
      int64_t pos = e_lo + a0;  // position of the first event
      int l = 0;
      for (int64_t j = pos - 1; j >= c_lo && l < cap_l; --j) {
        if (CD[j - c_lo] > 3) break;
        ++l;
      }
      // the two k-mers over the k - 1 bases before the stretch (only used when they are all unambiguous)
      uint64_t k0 = 0, k1 = 0;
      for (int64_t j = max(c_lo, pos - (k - 1)); j < pos; ++j) {
        const int c = CD[j - c_lo] & 3;
        k0 = (k0 << 2 | (uint64_t)c) & mask;
        k1 = k1 >> 2 | (uint64_t)(3 ^ c) << shift1;
      }
      for (int a = a0; a < a1; ++a, ++pos) {
        const int c = CD[pos - c_lo];
        uint64_t x = U64MAX;
        int z = 0;
        if (c < 4) {
          l = l < cap_l ? l + 1 : l;
          k0 = (k0 << 2 | (uint64_t)c) & mask;
          k1 = k1 >> 2 | (uint64_t)(3 ^ c) << shift1;
          if (l >= k) {
            z = k0 < k1 ? 0 : 1;
            x = hash64(z ? k1 : k0, mask) << 8 | (uint64_t)k;
          }
        } else l = 0;
        X[a] = x, EL[a] = (uint16_t)(l | z << 15);
      }
    }
  }
  __syncthreads();

  // ---- phase 2: what the reference's scan appends at each event (select_kernel's rules on shared memory) ----
  {
    const int64_t d_lo = t0, d_hi = e_hi;  // deciding events: the tile's own and the w after it
    const int n_dec = (int)(d_hi - d_lo);
    for (int q = tid; q < n_dec; q += kTileThreads) {
      const int64_t ee = d_lo + q;          // position of the event
      const int a = (int)(ee - e_lo);        // its slot
      const int lo_prev = (int)(max((int64_t)0, ee - w) - e_lo), lo_new = (int)(max((int64_t)0, ee - w + 1) - e_lo);
      const int jm = (int)(ee - w - e_lo);  // slot of the event that leaves the window at this step (negative: none yet)
      // newest-of-the-smallest over [ee-w, ee-1] (before this step) and over [ee-w+1, ee] (after it)
      int pm = -1, nm = -1;
      uint64_t xpm = U64MAX, xnm = U64MAX;
      for (int j = lo_prev; j < a; ++j) {
        const uint64_t x = X[j];
        if (xpm >= x) xpm = x, pm = j;
        if (j > jm && xnm >= x) xnm = x, nm = j;
      }
      const uint64_t xe = X[a];
      if (xnm >= xe) xnm = xe, nm = a;
      const int le = EL[a] & 0x7fff;
      if (le == w + k - 1 && xpm != U64MAX)
        for (int j = lo_new; j < a; ++j)
          if (X[j] == xpm && j != pm) F[j] = 1;
      if (xe <= xpm) {
        if (le >= w + k && xpm != U64MAX) F[pm] = 1;
      } else if (pm == jm) {
        if (le >= w + k - 1 && xpm != U64MAX) F[pm] = 1;
        if (le >= w + k - 1 && xnm != U64MAX)
          for (int j = lo_new; j <= a; ++j)
            if (X[j] == xnm && j != nm) F[j] = 1;
      }
      if (ee == L - 1 && xnm != U64MAX) F[nm] = 1;
    }
  }
  __syncthreads();

  // ---- phase 3: the tile's flagged positions, in order, to the front of the tile's slice ----
  {
    typedef cub::BlockScan<int, kTileThreads> Scan;
    __shared__ typename Scan::TempStorage scan_tmp;
    const int base = (int)(t0 - e_lo);  // slot of the tile's first position
    const int n_own = (int)(t1 - t0);
    int cnt = 0;
#pragma unroll
    for (int i = 0; i < kTileItems; ++i) {
      const int j = tid * kTileItems + i;
      cnt += j < n_own && F[base + j];
    }
    int off, total;
    Scan(scan_tmp).ExclusiveSum(cnt, off, total);
    const uint64_t g0 = v.vstart[tl.seq] + (uint64_t)t0;  // the tile's slice of the position-sized buffers
#pragma unroll
    for (int i = 0; i < kTileItems; ++i) {
      const int j = tid * kTileItems + i;
      if (j < n_own && F[base + j]) {
        TX[g0 + off] = X[base + j];
        TY[g0 + off] = (uint64_t)tl.seq << 32 | (uint64_t)(uint32_t)(t0 + j) << 1 | (uint64_t)(EL[base + j] >> 15);
        ++off;
      }
    }
    if (tid == 0) counts[blockIdx.x] = (uint32_t)total;
  }
}

// tile t's minimizers move from the front of its slice to their final place
__global__ void sketch_gather_kernel(SeqView v, const SketchTile *__restrict__ tiles, const uint32_t *__restrict__ counts,
                                     const uint32_t *__restrict__ prefix, const uint64_t *__restrict__ TX, const uint64_t *__restrict__ TY,
                                     uint64_t *__restrict__ mx, uint64_t *__restrict__ my) {
  const SketchTile tl = tiles[blockIdx.x];
  const uint64_t g0 = v.vstart[tl.seq] + (uint64_t)tl.t0;
  const uint32_t n = counts[blockIdx.x], o = prefix[blockIdx.x];
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) mx[o + i] = TX[g0 + i], my[o + i] = TY[g0 + i];
}

// ---------------- anchor order on the device ----------------
// x of every anchor, its query as a second key, and after the two stable sorts: does any query hold two anchors with the same x?
__global__ void anchor_keys_kernel(uint64_t n, const U128 *__restrict__ a, const uint64_t *__restrict__ qa_off, int nq,
                                   uint64_t *__restrict__ kx, uint32_t *__restrict__ kq, uint32_t *__restrict__ perm) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int lo = 0, hi = nq;  // qa_off[lo] <= i < qa_off[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (qa_off[mid] <= i) lo = mid;
    else hi = mid;
  }
  kx[i] = a[i].x, kq[i] = (uint32_t)lo, perm[i] = (uint32_t)i;
}
__global__ void gather_u32_kernel(uint64_t n, const uint32_t *__restrict__ src, const uint32_t *__restrict__ perm, uint32_t *__restrict__ dst) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[perm[i]];
}
__global__ void anchor_gather_kernel(uint64_t n, const U128 *__restrict__ a, const uint32_t *__restrict__ perm, const uint32_t *__restrict__ kq,
                                     U128 *__restrict__ out, uint32_t *__restrict__ tie) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const U128 v = a[perm[i]];
  out[i] = v;
  if (i > 0 && kq[i - 1] == kq[i] && a[perm[i - 1]].x == v.x) tie[kq[i]] = 1;
}

struct MaxOp {
  __device__ __forceinline__ int32_t operator()(int32_t a, int32_t b) const { return a > b ? a : b; }
};

// ---------------- index ----------------
__global__ void split_key_kernel(uint64_t n, const uint64_t *__restrict__ mx, uint64_t *__restrict__ key) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) key[i] = mx[i] >> 8;
}

// position of `key` in the sorted distinct keys, or -1

// ---------------- query-side occurrence filter (seed.c:5-28) ----------------
__global__ void mzflt_prepare_kernel(uint64_t n, const uint64_t *__restrict__ my, uint32_t *__restrict__ qid, uint32_t *__restrict__ idx) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) qid[i] = (uint32_t)(my[i] >> 32), idx[i] = (uint32_t)i;
}
__global__ void mzflt_heads_kernel(uint64_t n, const uint64_t *__restrict__ sx, const uint32_t *__restrict__ sq, uint32_t *__restrict__ head) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) head[i] = (i == 0 || sx[i] != sx[i - 1] || sq[i] != sq[i - 1]) ? 1u : 0u;
}
__global__ void mzflt_runstart_kernel(uint64_t n, const uint32_t *__restrict__ head, const uint32_t *__restrict__ run_incl,
                                      uint32_t *__restrict__ run_start, uint32_t n_runs) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && head[i]) run_start[run_incl[i] - 1] = (uint32_t)i;
  if (i == 0) run_start[n_runs] = (uint32_t)n;
}
__global__ void mzflt_mark_kernel(uint64_t n, const uint32_t *__restrict__ run_incl, const uint32_t *__restrict__ run_start,
                                  const uint32_t *__restrict__ sq, const uint32_t *__restrict__ sidx,
                                  const uint64_t *__restrict__ mz_off, int32_t q_occ_max, float q_occ_frac,
                                  uint32_t *__restrict__ keep) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t run = run_incl[i] - 1;
  const int32_t cnt = (int32_t)(run_start[run + 1] - run_start[run]);
  const uint32_t q = sq[i];
  const uint64_t nq = mz_off[q + 1] - mz_off[q];
  if ((int64_t)nq <= q_occ_max) return;  // the reference does not filter short queries at all
  if (cnt > q_occ_max && (float)cnt > __fmul_rn((float)nq, q_occ_frac)) keep[sidx[i]] = 0;
}

// generic ordered compaction helpers
__global__ void fill_u32_kernel(uint64_t n, uint32_t *a, uint32_t v) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = v;
}
__global__ void compact_mz_kernel(uint64_t n, const uint32_t *__restrict__ keep, const uint32_t *__restrict__ kpos,
                                  const uint64_t *__restrict__ mx, const uint64_t *__restrict__ my,
                                  uint64_t *__restrict__ fx, uint64_t *__restrict__ fy) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && keep[i]) fx[kpos[i]] = mx[i], fy[kpos[i]] = my[i];
}
// offsets of per-sequence segments after a compaction: new_off[s] = kpos[old_off[s]]
__global__ void remap_offsets_kernel(int n, const uint64_t *__restrict__ old_off, const uint32_t *__restrict__ kpos, uint64_t n_old,
                                     uint64_t n_new, uint64_t *__restrict__ new_off) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s > n) return;
  const uint64_t o = old_off[s];
  new_off[s] = o < n_old ? kpos[o] : n_new;
}

// ---------------- the hash table of the index (index.c:81-98: mm_idx_get is a khash lookup) ----------------
__device__ __forceinline__ uint64_t ht_slot(uint64_t key, uint64_t mask) {
  key *= 0x9E3779B97F4A7C15ull;  // minimizer hashes are already mixed (sketch.c:28-38); one multiply spreads their low bits
  return (key ^ key >> 29) & mask;
}
__global__ void ht_insert_kernel(uint64_t n_keys, const uint64_t *__restrict__ keys, unsigned long long *__restrict__ ht_key,
                                 uint32_t *__restrict__ ht_rank, uint64_t mask) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_keys) return;
  const unsigned long long key = keys[i];
  for (uint64_t s = ht_slot(key, mask);; s = (s + 1) & mask)
    if (atomicCAS(ht_key + s, ~0ull, key) == ~0ull) {  // keys are distinct: the first empty slot is ours
      ht_rank[s] = (uint32_t)i;
      return;
    }
}
// rank of `key` among the index's distinct keys, -1 if absent
__device__ __forceinline__ int64_t ht_find(const uint64_t *__restrict__ ht_key, const uint32_t *__restrict__ ht_rank, uint64_t mask, uint64_t key) {
  for (uint64_t s = ht_slot(key, mask);; s = (s + 1) & mask) {
    const uint64_t k = ht_key[s];
    if (k == key) return (int64_t)ht_rank[s];
    if (k == ~0ull) return -1;
  }
}

// ---------------- seed lookup (seed.c:30-52, index.c:81-98) ----------------
__global__ void lookup_kernel(uint64_t n, const uint64_t *__restrict__ fx, const uint64_t *__restrict__ fy, const uint64_t *__restrict__ f_off,
                              const uint64_t *__restrict__ ht_key, const uint32_t *__restrict__ ht_rank, uint64_t ht_mask,
                              const uint32_t *__restrict__ key_off,
                              uint32_t *__restrict__ s_n, uint32_t *__restrict__ s_off, uint32_t *__restrict__ s_tandem,
                              uint32_t *__restrict__ s_has) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t h = fx[i] >> 8;
  const int64_t kx = ht_find(ht_key, ht_rank, ht_mask, h);  // the hash probe
  uint32_t t = 0, off = 0;
  if (kx >= 0) off = key_off[kx], t = key_off[kx + 1] - off;
  const uint32_t q = (uint32_t)(fy[i] >> 32);
  uint32_t tandem = 0;
  if (i > f_off[q] && fx[i - 1] >> 8 == h) tandem = 1;
  if (i + 1 < f_off[q + 1] && fx[i + 1] >> 8 == h) tandem = 1;
  s_n[i] = t, s_off[i] = off, s_tandem[i] = tandem, s_has[i] = t > 0;
}

struct Seed {  // mm_seed_t without the pointer (mmpriv.h:41-47)
  uint32_t n, off;
  uint32_t q_pos;          // lastPos<<1 | strand of the query minimizer
  uint32_t q_span : 8, tandem : 1, flt : 1;
  uint32_t qid;
};

__global__ void compact_seed_kernel(uint64_t n, const uint32_t *__restrict__ s_has, const uint32_t *__restrict__ spos,
                                    const uint32_t *__restrict__ s_n, const uint32_t *__restrict__ s_off,
                                    const uint32_t *__restrict__ s_tandem, const uint64_t *__restrict__ fx,
                                    const uint64_t *__restrict__ fy, Seed *__restrict__ seeds) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !s_has[i]) return;
  Seed s;
  s.n = s_n[i], s.off = s_off[i], s.q_pos = (uint32_t)fy[i], s.q_span = (uint32_t)(fx[i] & 0xff), s.tandem = s_tandem[i], s.flt = 0;
  s.qid = (uint32_t)(fy[i] >> 32);
  seeds[spos[i]] = s;
}

// high-occurrence streak thinning (seed.c:56-96): one thread per streak of consecutive seeds with n > max_occ
__global__ void seed_select_kernel(uint64_t n, Seed *__restrict__ seeds, const uint64_t *__restrict__ sd_off, const int *__restrict__ qlens,
                                   int max_occ, int max_max_occ, int dist, int streak_mode) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Seed s = seeds[i];
  const uint64_t q0 = sd_off[s.qid], q1 = sd_off[s.qid + 1];
  if (!streak_mode) {  // seed.c:108-112
    if ((int64_t)s.n > max_occ) seeds[i].flt = 1;
    return;
  }
  if (q1 - q0 <= 1) return;  // n == 0 || n == 1: nothing is filtered
  if ((int64_t)s.n <= max_occ) return;
  if (i > q0 && (int64_t)seeds[i - 1].n > max_occ) return;  // not the first seed of its streak
  uint64_t en = i + 1;
  while (en < q1 && (int64_t)seeds[en].n > max_occ) ++en;
  const int32_t ps = i == q0 ? 0 : (int32_t)(seeds[i - 1].q_pos >> 1);
  const int32_t pe = en == q1 ? qlens[s.qid] : (int32_t)(seeds[en].q_pos >> 1);
  int32_t max_high_occ = (int32_t)((double)(pe - ps) / dist + .499);
  // keep the max_high_occ seeds with the smallest (n, index) -- what the reference's bounded max-heap leaves
  uint64_t best[128];
  int nb = 0;
  if (max_high_occ > 128) max_high_occ = 128;
  for (uint64_t j = i; j < en; ++j) seeds[j].flt = 1;
  if (max_high_occ > 0) {
    for (uint64_t j = i; j < en; ++j) {
      const uint64_t key = (uint64_t)seeds[j].n << 32 | (uint32_t)(j - q0);
      if (nb < max_high_occ) {
        int t = nb++;
        while (t > 0 && best[t - 1] > key) best[t] = best[t - 1], --t;
        best[t] = key;
      } else if (key < best[nb - 1]) {
        int t = nb - 1;
        while (t > 0 && best[t - 1] > key) best[t] = best[t - 1], --t;
        best[t] = key;
      }
    }
    for (int t = 0; t < nb; ++t) seeds[q0 + (uint32_t)best[t]].flt = 0;
  }
  for (uint64_t j = i; j < en; ++j)
    if ((int64_t)seeds[j].n > max_max_occ) seeds[j].flt = 1;
}

__global__ void seed_flags_kernel(uint64_t n, const Seed *__restrict__ seeds, uint32_t *__restrict__ used, uint32_t *__restrict__ flt) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) used[i] = !seeds[i].flt, flt[i] = seeds[i].flt;
}
__global__ void compact_used_kernel(uint64_t n, const Seed *__restrict__ seeds, const uint32_t *__restrict__ used, const uint32_t *__restrict__ upos,
                                    Seed *__restrict__ useds, uint64_t *__restrict__ mini_pos, uint32_t *__restrict__ u_n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !used[i]) return;
  const Seed s = seeds[i];
  const uint32_t o = upos[i];
  useds[o] = s;
  mini_pos[o] = (uint64_t)s.q_span << 32 | s.q_pos >> 1;
  u_n[o] = s.n;
}
__global__ void compact_flt_kernel(uint64_t n, const Seed *__restrict__ seeds, const uint32_t *__restrict__ flt, const uint32_t *__restrict__ fpos,
                                   Seed *__restrict__ flts) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && flt[i]) flts[fpos[i]] = seeds[i];
}
// bases covered by filtered seeds (seed.c:113-128): sum over maximal runs of (last end - first start)
__global__ void rep_len_kernel(uint64_t n, const Seed *__restrict__ flts, int32_t *__restrict__ rep_len) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Seed s = flts[i];
  const int32_t en = (int32_t)(s.q_pos >> 1) + 1, st = en - (int32_t)s.q_span;
  bool head = true, tail = true;
  if (i > 0 && flts[i - 1].qid == s.qid) head = st > (int32_t)(flts[i - 1].q_pos >> 1) + 1;
  if (i + 1 < n && flts[i + 1].qid == s.qid) {
    const Seed t = flts[i + 1];
    tail = ((int32_t)(t.q_pos >> 1) + 1 - (int32_t)t.q_span) > en;
  }
  const int32_t d = (tail ? en : 0) - (head ? st : 0);
  if (d) atomicAdd(&rep_len[s.qid], d);
}

// ---------------- anchor expansion with the all-vs-all skips (map.c:78-100,168-201) ----------------
__global__ void anchor_kernel(uint64_t n_slots, uint64_t n_used, const Seed *__restrict__ useds, const uint64_t *__restrict__ a_off,
                              const uint64_t *__restrict__ pos, const int *__restrict__ qlens, const int32_t *__restrict__ q_rank,
                              const int32_t *__restrict__ t_rank, const uint32_t *__restrict__ t_len, int64_t flag,
                              uint64_t *__restrict__ ax, uint64_t *__restrict__ ay, uint32_t *__restrict__ akeep) {
  const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_slots) return;
  uint64_t lo = 0, hi = n_used;  // a_off[lo] <= j < a_off[hi]
  while (hi - lo > 1) {
    const uint64_t mid = (lo + hi) >> 1;
    if (a_off[mid] <= j) lo = mid;
    else hi = mid;
  }
  const Seed s = useds[lo];
  const uint64_t r = pos[(uint64_t)s.off + (j - a_off[lo])];
  const uint32_t rid = (uint32_t)(r >> 32);
  const int qlen = qlens[s.qid];
  bool skip = false, is_self = false;
  const int32_t qr = q_rank[s.qid];
  if ((flag & (MM_F_NO_DIAG | MM_F_NO_DUAL)) && qr != INT32_MIN) {  // INT32_MIN: the query has no name (map.c:81)
    const int32_t tr = t_rank[rid];
    if ((flag & MM_F_NO_DIAG) && qr == tr && (int)t_len[rid] == qlen) {
      if ((uint32_t)r >> 1 == (s.q_pos >> 1)) skip = true;
      if ((r & 1) == (s.q_pos & 1)) is_self = true;
    }
    if ((flag & MM_F_NO_DUAL) && qr > tr) skip = true;
  }
  const uint32_t rpos = (uint32_t)r >> 1;
  uint64_t x, y;
  if ((r & 1) == (s.q_pos & 1)) {
    x = (r & 0xffffffff00000000ULL) | rpos;
    y = (uint64_t)s.q_span << 32 | s.q_pos >> 1;
  } else {
    x = 1ULL << 63 | (r & 0xffffffff00000000ULL) | rpos;
    y = (uint64_t)s.q_span << 32 | (uint32_t)(qlen - (int32_t)((s.q_pos >> 1) + 1 - s.q_span) - 1);
  }
  if (s.tandem) y |= SEED_TANDEM;
  if (is_self) y |= SEED_SELF;
  ax[j] = x, ay[j] = y, akeep[j] = !skip;
}
__global__ void compact_anchor_kernel(uint64_t n, const uint32_t *__restrict__ akeep, const uint32_t *__restrict__ apos,
                                      const uint64_t *__restrict__ ax, const uint64_t *__restrict__ ay, U128 *__restrict__ out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && akeep[i]) out[apos[i]] = U128{ax[i], ay[i]};
}
// first anchor slot of each query: slot offset of its first used seed
__global__ void query_slot_kernel(int nq, const uint64_t *__restrict__ u_off, const uint64_t *__restrict__ a_off, uint64_t n_used,
                                  uint64_t n_slots, uint64_t *__restrict__ q_slot) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q > nq) return;
  const uint64_t u = u_off[q];
  q_slot[q] = u < n_used ? a_off[u] : n_slots;
}

}  // namespace

struct SeedEngine::Impl {
  DevBuf<uint8_t> temp;  // cub scratch
  // sketch workspace
  DevBuf<uint8_t> ev;
  DevBuf<uint64_t> hv, EX, EP, vstart, starts;
  DevBuf<int32_t> rmark, rpos, EL;
  DevBuf<uint32_t> ev_incl, flag, fpos;
  DevBuf<SketchTile> tiles;
  // index workspace
  DevBuf<uint64_t> key_in, key_out, val_out, rle_keys;
  DevBuf<uint32_t> rle_cnt, cnt_sorted, n_runs;
  // collect workspace
  DevBuf<uint32_t> u32[16];
  DevBuf<uint64_t> u64[10];
  DevBuf<Seed> seeds, useds, flts;
  DevBuf<U128> anchors, anchors_sorted;
  DevBuf<int32_t> rep_len, q_rank;
  DevBuf<int> qlens;
  PinBuf<U128> h_anchors;  // pinned landing zones for the two large device->host copies
  PinBuf<uint64_t> h_mini;

  // Small device->host results (counts, offsets) land in PINNED memory: a cudaMemcpyAsync into pageable memory is
  // staged by the driver and the calling thread spins until the stream gets there -- with dozens of rounds in flight
  // that was 85 ms of host CPU per round (more than half of everything the host did).
  PinBuf<uint64_t> pinned;
  size_t pin_used = 0;
  void pin_reset(size_t words) { pinned.ensure(words + 64), pin_used = 0; }
  template <class T>
  T *pin(size_t n) {
    const size_t w = (n * sizeof(T) + 7) / 8 + 1;
    if (pin_used + w > pinned.cap) PGMM_FATAL("pinned scratch of the seeding stage is too small (%zu + %zu > %zu words)", pin_used, w, pinned.cap);
    T *p = (T *)(pinned.p + pin_used);
    pin_used += w;
    return p;
  }

  void *tmp(size_t bytes) { return temp.ensure(bytes + 256); }

  template <class In, class Out>
  void excl_sum(const In *in, Out *out, uint64_t n, cudaStream_t st) {
    size_t bytes = 0;
    PGMM_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int64_t)n, st));
    PGMM_CUDA(cub::DeviceScan::ExclusiveSum(tmp(bytes), bytes, in, out, (int64_t)n, st));
  }
  template <class In, class Out>
  void incl_sum(const In *in, Out *out, uint64_t n, cudaStream_t st) {
    size_t bytes = 0;
    PGMM_CUDA(cub::DeviceScan::InclusiveSum(nullptr, bytes, in, out, (int64_t)n, st));
    PGMM_CUDA(cub::DeviceScan::InclusiveSum(tmp(bytes), bytes, in, out, (int64_t)n, st));
  }
  // exclusive scan of 0/1 flags + total (synchronises the stream)
  uint64_t scan_flags(const uint32_t *f, uint32_t *pos, uint64_t n, cudaStream_t st) {
    if (n == 0) return 0;
    excl_sum(f, pos, n, st);
    uint32_t *last = pin<uint32_t>(2);
    PGMM_CUDA(cudaMemcpyAsync(last, pos + n - 1, 4, cudaMemcpyDeviceToHost, st));
    PGMM_CUDA(cudaMemcpyAsync(last + 1, f + n - 1, 4, cudaMemcpyDeviceToHost, st));
    PGMM_CUDA(cudaStreamSynchronize(st));
    return (uint64_t)last[0] + last[1];
  }
};

SeedEngine::SeedEngine() : impl_(new Impl) {}
SeedEngine::~SeedEngine() { delete impl_; }

void SeedEngine::sketch(const uint8_t *d_codes, const std::vector<uint64_t> &starts, const std::vector<int> &lens, int w, int k,
                        DeviceSeqSet &set, cudaStream_t st) {
  Impl &m = *impl_;
  const int n = (int)lens.size();
  size_t n_tiles_max = 0;  // tiles of the fused path (their table and their prefix sums pass through the pinned scratch)
  for (int i = 0; i < n; ++i) n_tiles_max += ((size_t)lens[i] + kTile - 1) / kTile;
  m.pin_reset(5 * ((size_t)n + 2) + 2 * n_tiles_max + 64);
  set.n = n, set.d_codes = d_codes, set.h_starts = starts, set.h_lens = lens;
  set.h_vstart.assign(n + 1, 0);
  for (int i = 0; i < n; ++i) set.h_vstart[i + 1] = set.h_vstart[i] + (uint64_t)lens[i];
  set.total = set.h_vstart[n];
  set.n_mz = 0;
  set.h_mz_off.assign(n + 1, 0);
  if (set.total >= (1ull << 31)) PGMM_FATAL("a batch of %llu bases exceeds the 2^31 positions one sketch launch indexes", (unsigned long long)set.total);
  if (w <= 0 || w >= 256 || k <= 0 || k > 28) PGMM_FATAL("minimizer parameters out of range: w=%d k=%d (need 0<w<256, 0<k<=28)", w, k);
  {
    uint64_t *ps = m.pin<uint64_t>(n + 1), *pv = m.pin<uint64_t>(n + 1);
    int *pl = m.pin<int>(n + 1);
    memcpy(ps, starts.data(), (size_t)n * 8), memcpy(pv, set.h_vstart.data(), ((size_t)n + 1) * 8), memcpy(pl, lens.data(), (size_t)n * 4);
    PGMM_CUDA(cudaMemcpyAsync(set.starts.ensure(n + 1), ps, n * 8, cudaMemcpyHostToDevice, st));
    PGMM_CUDA(cudaMemcpyAsync(set.vstart.ensure(n + 1), pv, (n + 1) * 8, cudaMemcpyHostToDevice, st));
    PGMM_CUDA(cudaMemcpyAsync(set.lens.ensure(n + 1), pl, n * 4, cudaMemcpyHostToDevice, st));
  }
  set.mz_off.ensure(n + 1);
  const uint64_t N = set.total;
  if (N == 0) {
    PGMM_CUDA(cudaMemsetAsync(set.mz_off.p, 0, (n + 1) * 8, st));
    PGMM_CUDA(cudaStreamSynchronize(st));
    return;
  }
  SeqView v{d_codes, set.starts.p, set.vstart.p, n, N};
  static const bool no_fused = getenv("PGMM_SKETCH_GENERAL") != nullptr && atoi(getenv("PGMM_SKETCH_GENERAL")) != 0;
  if ((k & 1) && !no_fused) {
    // ---- fused path (odd k): tiles of kTile positions, never across a sequence boundary ----
    std::vector<uint32_t> first_tile((size_t)n + 1);
    size_t n_tiles = 0;
    for (int i = 0; i < n; ++i) first_tile[i] = (uint32_t)n_tiles, n_tiles += ((size_t)lens[i] + kTile - 1) / kTile;
    first_tile[n] = (uint32_t)n_tiles;
    SketchTile *ht = m.pin<SketchTile>(n_tiles + 1);
    for (int i = 0; i < n; ++i)
      for (uint32_t t = first_tile[i], o = 0; t < first_tile[i + 1]; ++t, o += kTile) ht[t] = SketchTile{(uint32_t)i, o};
    m.tiles.ensure(n_tiles + 1), m.flag.ensure(n_tiles + 2), m.fpos.ensure(n_tiles + 2);
    m.EX.ensure(N), m.EP.ensure(N);
    PGMM_CUDA(cudaMemcpyAsync(m.tiles.p, ht, n_tiles * sizeof(SketchTile), cudaMemcpyHostToDevice, st));
    const int cap_ev = kTile + 2 * w;
    const size_t smem = (size_t)cap_ev * 8 + ((size_t)cap_ev * 2 + 15) / 16 * 16 + ((size_t)cap_ev + 15) / 16 * 16 + (size_t)kTile + 3 * w + k + 48;
    static const bool attr = [] {
      PGMM_CUDA(cudaFuncSetAttribute(sketch_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      return true;
    }();
    (void)attr;
    ++g_seed_launches, sketch_tile_kernel<<<(unsigned)n_tiles, kTileThreads, smem, st>>>(v, m.tiles.p, w, k, m.EX.p, m.EP.p, m.flag.p);
    PGMM_CUDA(cudaGetLastError());
    PGMM_CUDA(cudaMemsetAsync(m.flag.p + n_tiles, 0, 4, st));  // one more element: its exclusive sum is the total
    m.excl_sum(m.flag.p, m.fpos.p, n_tiles + 1, st);
    uint32_t *hp = m.pin<uint32_t>(n_tiles + 2);
    PGMM_CUDA(cudaMemcpyAsync(hp, m.fpos.p, (n_tiles + 1) * 4, cudaMemcpyDeviceToHost, st));
    PGMM_CUDA(cudaStreamSynchronize(st));
    const uint64_t n_mz = hp[n_tiles];
    set.n_mz = n_mz;
    set.mx.ensure(n_mz + 1), set.my.ensure(n_mz + 1);
    ++g_seed_launches, sketch_gather_kernel<<<(unsigned)n_tiles, 128, 0, st>>>(v, m.tiles.p, m.flag.p, m.fpos.p, m.EX.p, m.EP.p, set.mx.p, set.my.p);
    PGMM_CUDA(cudaGetLastError());
    for (int i = 0; i <= n; ++i) set.h_mz_off[i] = hp[first_tile[i]];
    uint64_t *p_off = m.pin<uint64_t>(n + 1);
    memcpy(p_off, set.h_mz_off.data(), ((size_t)n + 1) * 8);
    PGMM_CUDA(cudaMemcpyAsync(set.mz_off.p, p_off, (n + 1) * 8, cudaMemcpyHostToDevice, st));
    return;
  }
  m.ev.ensure(N), m.hv.ensure(N), m.rmark.ensure(N), m.rpos.ensure(N), m.ev_incl.ensure(N);
  ++g_seed_launches, kmer_kernel<<<nblk(N), TPB, 0, st>>>(v, k, m.ev.p, m.hv.p, m.rmark.p);
  m.incl_sum(m.ev.p, m.ev_incl.p, N, st);
  {
    size_t bytes = 0;
    PGMM_CUDA(cub::DeviceScan::InclusiveScan(nullptr, bytes, m.rmark.p, m.rpos.p, MaxOp(), (int64_t)N, st));
    PGMM_CUDA(cub::DeviceScan::InclusiveScan(m.tmp(bytes), bytes, m.rmark.p, m.rpos.p, MaxOp(), (int64_t)N, st));
  }
  uint32_t *p_n_ev = m.pin<uint32_t>(1);
  PGMM_CUDA(cudaMemcpyAsync(p_n_ev, m.ev_incl.p + N - 1, 4, cudaMemcpyDeviceToHost, st));
  PGMM_CUDA(cudaStreamSynchronize(st));
  const uint32_t n_ev = *p_n_ev;
  if (n_ev == 0) {
    PGMM_CUDA(cudaMemsetAsync(set.mz_off.p, 0, (n + 1) * 8, st));
    PGMM_CUDA(cudaStreamSynchronize(st));
    return;
  }
  m.EX.ensure(n_ev), m.EP.ensure(n_ev), m.EL.ensure(n_ev), m.flag.ensure(n_ev), m.fpos.ensure(n_ev);
  ++g_seed_launches, event_kernel<<<nblk(N), TPB, 0, st>>>(v, k, m.ev.p, m.hv.p, m.ev_incl.p, m.rpos.p, m.EX.p, m.EP.p, m.EL.p);
  PGMM_CUDA(cudaMemsetAsync(m.flag.p, 0, (size_t)n_ev * 4, st));
  ++g_seed_launches, select_kernel<<<nblk(n_ev), TPB, 0, st>>>(v, w, k, n_ev, m.ev_incl.p, m.EX.p, m.EP.p, m.EL.p, m.flag.p);
  PGMM_CUDA(cudaGetLastError());
  const uint64_t n_mz = m.scan_flags(m.flag.p, m.fpos.p, n_ev, st);
  set.n_mz = n_mz;
  set.mx.ensure(n_mz + 1), set.my.ensure(n_mz + 1);
  ++g_seed_launches, scatter_mz_kernel<<<nblk(n_ev), TPB, 0, st>>>(v, n_ev, m.flag.p, m.fpos.p, m.EX.p, m.EP.p, set.mx.p, set.my.p);
  ++g_seed_launches, mz_offsets_kernel<<<nblk(n + 1), TPB, 0, st>>>(v, n_ev, m.ev_incl.p, m.fpos.p, (uint32_t)n_mz, set.mz_off.p);
  PGMM_CUDA(cudaGetLastError());
  uint64_t *p_off = m.pin<uint64_t>(n + 1);
  PGMM_CUDA(cudaMemcpyAsync(p_off, set.mz_off.p, (n + 1) * 8, cudaMemcpyDeviceToHost, st));
  PGMM_CUDA(cudaStreamSynchronize(st));
  memcpy(set.h_mz_off.data(), p_off, ((size_t)n + 1) * 8);
}

void SeedEngine::build_index(DeviceIndex &idx, const std::vector<uint32_t> &lens, const std::vector<int32_t> &name_rank, cudaStream_t st) {
  Impl &m = *impl_;
  DeviceSeqSet &set = idx.seqs;
  const uint64_t n = set.n_mz;
  const int ns = set.n;
  m.pin_reset((size_t)ns + 64);
  {
    uint32_t *pl = m.pin<uint32_t>(ns + 1);
    int32_t *pr = m.pin<int32_t>(ns + 1);
    memcpy(pl, lens.data(), (size_t)ns * 4), memcpy(pr, name_rank.data(), (size_t)ns * 4);
    PGMM_CUDA(cudaMemcpyAsync(idx.seq_len.ensure(ns + 1), pl, ns * 4, cudaMemcpyHostToDevice, st));
    PGMM_CUDA(cudaMemcpyAsync(idx.name_rank.ensure(ns + 1), pr, ns * 4, cudaMemcpyHostToDevice, st));
  }
  idx.n_keys = 0;
  idx.keys.ensure(n + 1), idx.key_off.ensure(n + 2), idx.pos.ensure(n + 1), idx.occ_sorted.ensure(n + 1);
  if (n == 0) {
    PGMM_CUDA(cudaStreamSynchronize(st));
    return;
  }
  // minimizers arrive ordered by (sequence, position); a stable sort on the hash therefore leaves every key's
  // positions ascending, which is the order mm_idx_get hands out (index.c:245-255)
  m.key_in.ensure(n), m.key_out.ensure(n);
  ++g_seed_launches, split_key_kernel<<<nblk(n), TPB, 0, st>>>(n, set.mx.p, m.key_in.p);
  {
    size_t bytes = 0;
    PGMM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, m.key_in.p, m.key_out.p, set.my.p, idx.pos.p, (int64_t)n, 0, 2 * idx.k, st));
    PGMM_CUDA(cub::DeviceRadixSort::SortPairs(m.tmp(bytes), bytes, m.key_in.p, m.key_out.p, set.my.p, idx.pos.p, (int64_t)n, 0, 2 * idx.k, st));
  }
  m.rle_cnt.ensure(n + 1), m.n_runs.ensure(4);
  {
    size_t bytes = 0;
    PGMM_CUDA(cub::DeviceRunLengthEncode::Encode(nullptr, bytes, m.key_out.p, idx.keys.p, m.rle_cnt.p, m.n_runs.p, (int64_t)n, st));
    PGMM_CUDA(cub::DeviceRunLengthEncode::Encode(m.tmp(bytes), bytes, m.key_out.p, idx.keys.p, m.rle_cnt.p, m.n_runs.p, (int64_t)n, st));
  }
  uint32_t *p_n_keys = m.pin<uint32_t>(1);
  PGMM_CUDA(cudaMemcpyAsync(p_n_keys, m.n_runs.p, 4, cudaMemcpyDeviceToHost, st));
  PGMM_CUDA(cudaStreamSynchronize(st));
  const uint32_t n_keys = *p_n_keys;
  idx.n_keys = n_keys;
  {  // the probe table: the smallest power of two >= 2 n_keys slots
    uint64_t slots = 16;
    while (slots < 2ull * n_keys) slots <<= 1;
    idx.ht_mask = slots - 1;
    idx.ht_key.ensure(slots), idx.ht_rank.ensure(slots);
    PGMM_CUDA(cudaMemsetAsync(idx.ht_key.p, 0xff, slots * 8, st));
    if (n_keys) ++g_seed_launches, ht_insert_kernel<<<nblk(n_keys), TPB, 0, st>>>(n_keys, idx.keys.p, (unsigned long long *)idx.ht_key.p, idx.ht_rank.p, idx.ht_mask);
    PGMM_CUDA(cudaGetLastError());
  }
  PGMM_CUDA(cudaMemsetAsync(m.rle_cnt.p + n_keys, 0, 4, st));
  m.excl_sum(m.rle_cnt.p, idx.key_off.p, (uint64_t)n_keys + 1, st);
  {
    size_t bytes = 0;
    PGMM_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, bytes, m.rle_cnt.p, idx.occ_sorted.p, (int64_t)n_keys, 0, 32, st));
    PGMM_CUDA(cub::DeviceRadixSort::SortKeys(m.tmp(bytes), bytes, m.rle_cnt.p, idx.occ_sorted.p, (int64_t)n_keys, 0, 32, st));
  }
  PGMM_CUDA(cudaStreamSynchronize(st));
}

int32_t SeedEngine::cal_max_occ(const DeviceIndex &idx, float f, cudaStream_t st) {
  if (f <= 0.) return INT32_MAX;
  const size_t n = idx.n_keys;
  if (n == 0) return 1;  // the reference reads an empty array here; an empty index never reaches a lookup
  const uint32_t kk = (uint32_t)((1. - f) * n);
  thread_local PinBuf<uint32_t> pinned_word;
  uint32_t *v = pinned_word.ensure(2);
  PGMM_CUDA(cudaMemcpyAsync(v, idx.occ_sorted.p + (kk < n ? kk : n - 1), 4, cudaMemcpyDeviceToHost, st));
  PGMM_CUDA(cudaStreamSynchronize(st));
  return (int32_t)(*v + 1);
}

void SeedEngine::collect(const DeviceIndex &idx, const DeviceSeqSet &qs, const std::vector<int32_t> &q_name_rank, const mm_mapopt_t &opt,
                         std::vector<QuerySeeds> &out, cudaStream_t st) {
  Impl &m = *impl_;
  const int nq = qs.n;
  out.assign(nq, QuerySeeds());
  uint64_t n = qs.n_mz;
  if (n == 0 || idx.n_keys == 0) return;
  m.pin_reset(5 * ((size_t)nq + 2) + 64);
  {
    int32_t *pr = m.pin<int32_t>(nq + 1);
    memcpy(pr, q_name_rank.data(), (size_t)nq * 4);
    PGMM_CUDA(cudaMemcpyAsync(m.q_rank.ensure(nq + 1), pr, nq * 4, cudaMemcpyHostToDevice, st));
  }
  const uint64_t *fx = qs.mx.p, *fy = qs.my.p, *f_off = qs.mz_off.p;

  // ---- query-side occurrence filter (seed.c:5-28); only queries with more than mid_occ minimizers are affected ----
  bool need_flt = false;
  if (opt.q_occ_frac > 0.0f && opt.mid_occ > 0)
    for (int q = 0; q < nq; ++q) need_flt |= (int64_t)(qs.h_mz_off[q + 1] - qs.h_mz_off[q]) > opt.mid_occ;
  if (need_flt) {
    uint32_t *qid = m.u32[0].ensure(n), *idxv = m.u32[1].ensure(n), *sq = m.u32[2].ensure(n), *sidx = m.u32[3].ensure(n);
    uint32_t *qid2 = m.u32[4].ensure(n), *idx2 = m.u32[5].ensure(n), *head = m.u32[6].ensure(n), *run_incl = m.u32[7].ensure(n);
    uint32_t *keep = m.u32[8].ensure(n), *kpos = m.u32[9].ensure(n);
    uint64_t *sx = m.u64[0].ensure(n), *sx2 = m.u64[1].ensure(n);
    ++g_seed_launches, mzflt_prepare_kernel<<<nblk(n), TPB, 0, st>>>(n, qs.my.p, qid, idxv);
    // sort by x, carrying (query id, index); then stable by query id: equal (query, x) become adjacent
    size_t bytes = 0;
    PGMM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, qs.mx.p, sx, idxv, idx2, (int64_t)n, 0, 64, st));
    PGMM_CUDA(cub::DeviceRadixSort::SortPairs(m.tmp(bytes), bytes, qs.mx.p, sx, idxv, idx2, (int64_t)n, 0, 64, st));
    PGMM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, qs.mx.p, sx, qid, qid2, (int64_t)n, 0, 64, st));
    PGMM_CUDA(cub::DeviceRadixSort::SortPairs(m.tmp(bytes), bytes, qs.mx.p, sx, qid, qid2, (int64_t)n, 0, 64, st));
    int qbits = 1;
    while ((1 << qbits) < nq) ++qbits;
    PGMM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, qid2, sq, idx2, sidx, (int64_t)n, 0, qbits, st));
    PGMM_CUDA(cub::DeviceRadixSort::SortPairs(m.tmp(bytes), bytes, qid2, sq, idx2, sidx, (int64_t)n, 0, qbits, st));
    PGMM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, qid2, sq, sx, sx2, (int64_t)n, 0, qbits, st));
    PGMM_CUDA(cub::DeviceRadixSort::SortPairs(m.tmp(bytes), bytes, qid2, sq, sx, sx2, (int64_t)n, 0, qbits, st));
    ++g_seed_launches, mzflt_heads_kernel<<<nblk(n), TPB, 0, st>>>(n, sx2, sq, head);
    m.incl_sum(head, run_incl, n, st);
    uint32_t *p_n_runs = m.pin<uint32_t>(1);
    PGMM_CUDA(cudaMemcpyAsync(p_n_runs, run_incl + n - 1, 4, cudaMemcpyDeviceToHost, st));
    PGMM_CUDA(cudaStreamSynchronize(st));
    const uint32_t n_runs = *p_n_runs;
    uint32_t *run_start = m.u32[10].ensure((uint64_t)n_runs + 2);
    ++g_seed_launches, mzflt_runstart_kernel<<<nblk(n), TPB, 0, st>>>(n, head, run_incl, run_start, n_runs);
    ++g_seed_launches, fill_u32_kernel<<<nblk(n), TPB, 0, st>>>(n, keep, 1u);
    ++g_seed_launches, mzflt_mark_kernel<<<nblk(n), TPB, 0, st>>>(n, run_incl, run_start, sq, sidx, qs.mz_off.p, opt.mid_occ, opt.q_occ_frac, keep);
    PGMM_CUDA(cudaGetLastError());
    const uint64_t n_keep = m.scan_flags(keep, kpos, n, st);
    if (n_keep != n) {
      uint64_t *cx = m.u64[2].ensure(n_keep + 1), *cy = m.u64[3].ensure(n_keep + 1), *c_off = m.u64[4].ensure(nq + 2);
      ++g_seed_launches, compact_mz_kernel<<<nblk(n), TPB, 0, st>>>(n, keep, kpos, qs.mx.p, qs.my.p, cx, cy);
      ++g_seed_launches, remap_offsets_kernel<<<nblk(nq + 1), TPB, 0, st>>>(nq, qs.mz_off.p, kpos, n, n_keep, c_off);
      fx = cx, fy = cy, f_off = c_off, n = n_keep;
      if (n == 0) return;
    }
  }

  // ---- index probe and seed list (seed.c:30-52) ----
  uint32_t *s_n = m.u32[0].ensure(n), *s_off = m.u32[1].ensure(n), *s_tandem = m.u32[2].ensure(n), *s_has = m.u32[3].ensure(n);
  uint32_t *spos = m.u32[4].ensure(n);
  ++g_seed_launches, lookup_kernel<<<nblk(n), TPB, 0, st>>>(n, fx, fy, f_off, idx.ht_key.p, idx.ht_rank.p, idx.ht_mask, idx.key_off.p, s_n, s_off, s_tandem, s_has);
  PGMM_CUDA(cudaGetLastError());
  const uint64_t n_seed = m.scan_flags(s_has, spos, n, st);
  if (n_seed == 0) return;
  Seed *seeds = m.seeds.ensure(n_seed);
  uint64_t *sd_off = m.u64[5].ensure(nq + 2);
  ++g_seed_launches, compact_seed_kernel<<<nblk(n), TPB, 0, st>>>(n, s_has, spos, s_n, s_off, s_tandem, fx, fy, seeds);
  ++g_seed_launches, remap_offsets_kernel<<<nblk(nq + 1), TPB, 0, st>>>(nq, f_off, spos, n, n_seed, sd_off);

  // ---- high-occurrence thinning (seed.c:56-96,105-112) ----
  const int streak_mode = opt.occ_dist > 0 && opt.max_max_occ > opt.mid_occ;
  ++g_seed_launches, seed_select_kernel<<<nblk(n_seed), TPB, 0, st>>>(n_seed, seeds, sd_off, qs.lens.p, opt.mid_occ, opt.max_max_occ, opt.occ_dist, streak_mode);
  uint32_t *used = m.u32[5].ensure(n_seed), *flt = m.u32[6].ensure(n_seed), *upos = m.u32[7].ensure(n_seed), *fpos = m.u32[8].ensure(n_seed);
  ++g_seed_launches, seed_flags_kernel<<<nblk(n_seed), TPB, 0, st>>>(n_seed, seeds, used, flt);
  PGMM_CUDA(cudaGetLastError());
  const uint64_t n_used = m.scan_flags(used, upos, n_seed, st);
  const uint64_t n_flt = m.scan_flags(flt, fpos, n_seed, st);

  // ---- repeat length from the filtered seeds (seed.c:113-128) ----
  int32_t *rep_len = m.rep_len.ensure(nq + 1);
  PGMM_CUDA(cudaMemsetAsync(rep_len, 0, (nq + 1) * 4, st));
  if (n_flt > 0) {
    Seed *flts = m.flts.ensure(n_flt);
    ++g_seed_launches, compact_flt_kernel<<<nblk(n_seed), TPB, 0, st>>>(n_seed, seeds, flt, fpos, flts);
    ++g_seed_launches, rep_len_kernel<<<nblk(n_flt), TPB, 0, st>>>(n_flt, flts, rep_len);
  }
  int32_t *h_rep = m.pin<int32_t>(nq + 1);
  PGMM_CUDA(cudaMemcpyAsync(h_rep, rep_len, nq * 4, cudaMemcpyDeviceToHost, st));

  uint64_t *h_u_off = m.pin<uint64_t>(nq + 1), *h_a_off = m.pin<uint64_t>(nq + 1);
  memset(h_u_off, 0, ((size_t)nq + 1) * 8), memset(h_a_off, 0, ((size_t)nq + 1) * 8);
  uint64_t *h_mini = nullptr;
  U128 *h_anchors = nullptr, *a_sorted = nullptr;
  const U128 *a_orig = nullptr;
  uint32_t *a_tie = nullptr, *h_tie = nullptr;
  if (n_used > 0) {
    // ---- used seeds, their query positions (mini_pos) and the anchor slots they expand to ----
    Seed *useds = m.useds.ensure(n_used);
    uint64_t *mini_pos = m.u64[6].ensure(n_used), *u_off = m.u64[7].ensure(nq + 2), *a_off = m.u64[8].ensure(n_used + 1);
    uint32_t *u_n = m.u32[9].ensure(n_used + 1);
    ++g_seed_launches, compact_used_kernel<<<nblk(n_seed), TPB, 0, st>>>(n_seed, seeds, used, upos, useds, mini_pos, u_n);
    ++g_seed_launches, remap_offsets_kernel<<<nblk(nq + 1), TPB, 0, st>>>(nq, sd_off, upos, n_seed, n_used, u_off);
    PGMM_CUDA(cudaMemsetAsync(u_n + n_used, 0, 4, st));
    m.excl_sum(u_n, a_off, n_used + 1, st);
    uint64_t *p_n_slots = m.pin<uint64_t>(1);
    PGMM_CUDA(cudaMemcpyAsync(p_n_slots, a_off + n_used, 8, cudaMemcpyDeviceToHost, st));
    PGMM_CUDA(cudaMemcpyAsync(h_u_off, u_off, (nq + 1) * 8, cudaMemcpyDeviceToHost, st));
    h_mini = m.h_mini.ensure(n_used);
    PGMM_CUDA(cudaMemcpyAsync(h_mini, mini_pos, n_used * 8, cudaMemcpyDeviceToHost, st));
    PGMM_CUDA(cudaStreamSynchronize(st));
    const uint64_t n_slots = *p_n_slots;
    if (n_slots >= (1ull << 32)) PGMM_FATAL("%llu anchors in one batch exceed the 2^32 slots of the expansion pass", (unsigned long long)n_slots);
    if (n_slots > 0) {
      uint64_t *ax = m.u64[0].ensure(n_slots), *ay = m.u64[1].ensure(n_slots), *q_slot = m.u64[9].ensure(nq + 2);
      uint32_t *akeep = m.u32[10].ensure(n_slots), *apos = m.u32[11].ensure(n_slots);
      ++g_seed_launches, anchor_kernel<<<nblk(n_slots), TPB, 0, st>>>(n_slots, n_used, useds, a_off, idx.pos.p, qs.lens.p, m.q_rank.p, idx.name_rank.p,
                                                   idx.seq_len.p, opt.flag, ax, ay, akeep);
      ++g_seed_launches, query_slot_kernel<<<nblk(nq + 1), TPB, 0, st>>>(nq, u_off, a_off, n_used, n_slots, q_slot);
      PGMM_CUDA(cudaGetLastError());
      const uint64_t n_anchor = m.scan_flags(akeep, apos, n_slots, st);
      uint64_t *qa_off = m.u64[2].ensure(nq + 2);
      ++g_seed_launches, remap_offsets_kernel<<<nblk(nq + 1), TPB, 0, st>>>(nq, q_slot, apos, n_slots, n_anchor, qa_off);
      PGMM_CUDA(cudaMemcpyAsync(h_a_off, qa_off, (nq + 1) * 8, cudaMemcpyDeviceToHost, st));
      if (n_anchor > 0) {
        U128 *anchors = m.anchors.ensure(n_anchor);
        ++g_seed_launches, compact_anchor_kernel<<<nblk(n_slots), TPB, 0, st>>>(n_slots, akeep, apos, ax, ay, anchors);
        h_anchors = m.h_anchors.ensure(n_anchor);
        // Worth it for small batches only: two random 19-mers of a 5-Mbp genome coincide often enough (a few dozen pairs) that a
        // large query practically always holds a tie somewhere, and a tie anywhere sends the whole query to the host replay.
        static const uint64_t dev_sort_max = getenv("PGMM_DEVICE_ANCHOR_SORT_MAX") ? (uint64_t)atoll(getenv("PGMM_DEVICE_ANCHOR_SORT_MAX")) : 65536;
        if (n_anchor <= dev_sort_max) {
          // The reference sorts every query's anchors by target position with an UNSTABLE in-place radix sort (map.c:202):
          // the order it leaves among equal positions is observable, everything else is not.  So: stable sort on the device
          // (by position, then by query), look for equal neighbours; a query without any gets its anchors back sorted and
          // the host skips its replay of the reference's sort, a query with ties gets them in collection order as before.
          uint64_t *kx = m.u64[3].ensure(n_anchor), *kx2 = m.u64[4].ensure(n_anchor);
          uint32_t *kq = m.u32[12].ensure(n_anchor), *kq2 = m.u32[13].ensure(n_anchor), *pm = m.u32[14].ensure(n_anchor), *pm2 = m.u32[15].ensure(n_anchor);
          a_tie = m.u32[8].ensure(nq + 1);
          a_sorted = m.anchors_sorted.ensure(n_anchor);
          PGMM_CUDA(cudaMemsetAsync(a_tie, 0, ((size_t)nq + 1) * 4, st));
          ++g_seed_launches, anchor_keys_kernel<<<nblk(n_anchor), TPB, 0, st>>>(n_anchor, anchors, qa_off, nq, kx, kq, pm);
          size_t bytes = 0;
          PGMM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, kx, kx2, pm, pm2, (int64_t)n_anchor, 0, 64, st));
          PGMM_CUDA(cub::DeviceRadixSort::SortPairs(m.tmp(bytes), bytes, kx, kx2, pm, pm2, (int64_t)n_anchor, 0, 64, st));
          if (nq > 1) {  // anchors of all queries went through one sort: bring every query's back together (stable, by query)
            int qbits = 1;
            while ((1 << qbits) < nq) ++qbits;
            ++g_seed_launches, gather_u32_kernel<<<nblk(n_anchor), TPB, 0, st>>>(n_anchor, kq, pm2, kq2);  // query of each sorted anchor
            PGMM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, kq2, kq, pm2, pm, (int64_t)n_anchor, 0, qbits, st));
            PGMM_CUDA(cub::DeviceRadixSort::SortPairs(m.tmp(bytes), bytes, kq2, kq, pm2, pm, (int64_t)n_anchor, 0, qbits, st));
            // (kq: queries ascending = the segments of qa_off; pm: the permutation)
          } else {
            PGMM_CUDA(cudaMemsetAsync(kq, 0, n_anchor * 4, st));
            uint32_t *t = pm; pm = pm2, pm2 = t;
          }
          ++g_seed_launches, anchor_gather_kernel<<<nblk(n_anchor), TPB, 0, st>>>(n_anchor, anchors, pm, kq, a_sorted, a_tie);
          PGMM_CUDA(cudaGetLastError());
          h_tie = m.pin<uint32_t>(nq + 1);
          PGMM_CUDA(cudaMemcpyAsync(h_tie, a_tie, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
          PGMM_CUDA(cudaMemcpyAsync(h_anchors, a_sorted, n_anchor * sizeof(U128), cudaMemcpyDeviceToHost, st));
          a_orig = anchors;
        } else PGMM_CUDA(cudaMemcpyAsync(h_anchors, anchors, n_anchor * sizeof(U128), cudaMemcpyDeviceToHost, st));
      }
    }
  }
  PGMM_CUDA(cudaStreamSynchronize(st));
  bool refetch = false;
  for (int q = 0; q < nq && h_tie; ++q)
    if (h_tie[q] && h_a_off[q + 1] > h_a_off[q]) {  // equal positions in this query: the host needs the collection order
      PGMM_CUDA(cudaMemcpyAsync(h_anchors + h_a_off[q], a_orig + h_a_off[q], (h_a_off[q + 1] - h_a_off[q]) * sizeof(U128), cudaMemcpyDeviceToHost, st));
      refetch = true;
    }
  if (refetch) PGMM_CUDA(cudaStreamSynchronize(st));
  for (int q = 0; q < nq; ++q) {
    QuerySeeds &o = out[q];
    o.rep_len = h_rep[q];
    if (h_mini) o.mini_pos.assign(h_mini + h_u_off[q], h_mini + h_u_off[q + 1]);
    if (h_anchors) o.a.assign(h_anchors + h_a_off[q], h_anchors + h_a_off[q + 1]);
    o.sorted = h_tie != nullptr && h_tie[q] == 0;
    if (!o.a.empty()) ++(o.sorted ? g_anchor_sorted_device : g_anchor_sorted_host);
  }
}

}  // namespace pgmm
