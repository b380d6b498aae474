// K7: the minimizer sketch behind pangraph's mash distance -- the parts shared by the CUDA kernel (mash.cu)
// and the host (the CPU emulation of the kernel's schedule in tests/mash_emul.cpp): tile geometry, the rolling k-mers of one
// thread's stretch of positions, and the decision one position takes.
//
// Reference (PG = packages/pangraph/src): PG/distance/mash/minimizer.rs:49-160 (minimizers_sketch), hash.rs:3-12.
//
// minimizers_sketch is a sequential scan with a ring of w slots, the same shape as minimap2's mm_sketch (K1, seeding.cu)
// with three differences that matter here: the running minimum is the OLDEST of the smallest values in the window (strict <
// when a new value arrives and when the ring is rescanned, minimizer.rs:108,121,128), a k-mer equal to its reverse complement
// is an ordinary event (fwd <= rev picks the forward strand, :80), and the value is the bare hash.  Every position writes one
// ring slot, so the state after position e is a pure function of the values of positions [e-w+1, e]: position e can decide on
// its own what the scan appends at that step (first-window duplicates :93-104, a displaced minimum :108-112, a minimum leaving
// the window and the duplicates of its successor :113-146, the minimum standing at the end :153-155).
//
// With the oldest-wins rule the duplicates are NEWER than the minimum they duplicate, are appended before it, and are
// appended again each time an older duplicate leaves: the reference's list is neither in position order nor free of
// repetitions.  Its only consumer, mash_distance (mash_distance.rs:16-48), sorts the list by value and reduces every group
// of equal values to the set of sequences behind it -- so what K7 has to reproduce is the SET of (value, position) pairs
// the scan appends, which is what the flags below are.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define MASH_HD __host__ __device__ __forceinline__
#else
#define MASH_HD inline
#endif

namespace pgmm {
namespace mash {

constexpr uint64_t kNone = ~0ull;
constexpr int kTile = 4096, kTileThreads = 256, kTileItems = kTile / kTileThreads;
constexpr int kMaxK = 31, kMaxW = 255;  // minimizer.rs:53-54: assert!(k < 32), assert!(w < 256)

MASH_HD uint64_t hash(uint64_t x, uint64_t mask) {  // hash.rs:3-12
  x = (~x + (x << 21)) & mask;
  x = x ^ (x >> 24);
  x = (x + (x << 3) + (x << 8)) & mask;
  x = x ^ (x >> 14);
  x = (x + (x << 2) + (x << 4)) & mask;
  x = x ^ (x >> 28);
  x = (x + (x << 31)) & mask;
  return x;
}

// minimizer.rs:163-181: A C G T/U in either case are 0..3, every other byte is 4
MASH_HD int code(uint8_t c) {
  const uint8_t u = c & 0xdf;  // only 'A' and 'a' give 'A', and so on
  return u == 'A' ? 0 : u == 'C' ? 1 : u == 'G' ? 2 : (u == 'T' || u == 'U') ? 3 : 4;
}

// One tile = kTile consecutive positions [t0, t1) of one sequence.  Deciding positions: the tile's own and the w after it
// (they flag positions up to w behind themselves).  Values are needed from w before the tile; bases from a further
// w + k + 1 back, where the run length of the first value is decided.
struct Tile {
  int64_t L, t0, t1, e_lo, e_hi, c_lo;
  int n_ev, n_codes;
};
MASH_HD Tile tile_of(int64_t L, int64_t t0, int w, int k) {
  Tile t;
  t.L = L, t.t0 = t0, t.t1 = t0 + kTile < L ? t0 + kTile : L;
  t.e_lo = t0 - w > 0 ? t0 - w : 0;
  t.e_hi = t.t1 + w < L ? t.t1 + w : L;
  t.c_lo = t.e_lo - (w + k + 1) > 0 ? t.e_lo - (w + k + 1) : 0;
  t.n_ev = (int)(t.e_hi - t.e_lo), t.n_codes = (int)(t.e_hi - t.c_lo);
  return t;
}
MASH_HD int cap_ev(int w) { return kTile + 2 * w; }                 // slots of X / EL / P / S / F
MASH_HD int cap_codes(int w, int k) { return kTile + 3 * w + k + 48; }  // staged bases (+ alignment slack)

// The tile's working set in shared memory (byte offsets, every part on a 16-byte boundary): values X[u64], run length |
// strand << 15 EL[u16], block-prefix / block-suffix minima P, S [u16], flags F[u8], staged bases CD[u8].
struct Layout {
  uint32_t x, el, p, s, f, cd, total;
};
MASH_HD Layout layout_of(int w, int k) {
  const uint32_t ce = (uint32_t)cap_ev(w), r16 = (ce * 2 + 15) / 16 * 16;
  Layout l;
  l.x = 0, l.el = ce * 8, l.p = l.el + r16, l.s = l.p + r16, l.f = l.s + r16, l.cd = l.f + (ce + 15) / 16 * 16;
  l.total = l.cd + ((uint32_t)cap_codes(w, k) + 15) / 16 * 16;
  return l;
}

// Thread `tid` of `n_threads`: value (or kNone) and run length | strand << 15 of its stretch of the slots [0, n_ev).
// CD[i] = code of position c_lo + i.
MASH_HD void roll(const Tile &t, int w, int k, int tid, int n_threads, const uint8_t *CD, uint64_t *X, uint16_t *EL) {
  const int per = (t.n_ev + n_threads - 1) / n_threads;
  const int a0 = tid * per, a1 = a0 + per < t.n_ev ? a0 + per : t.n_ev;
  if (a0 >= a1) return;
  const uint64_t mask = (1ull << 2 * k) - 1;
  const int shift1 = 2 * (k - 1), cap_l = w + k + 1;
  int64_t pos = t.e_lo + a0;
  int l = 0;  // unambiguous bases right before the stretch (exact up to w + k + 1)
  for (int64_t j = pos - 1; j >= t.c_lo && l < cap_l; --j) {
    if (CD[j - t.c_lo] > 3) break;
    ++l;
  }
  uint64_t fwd = 0, rev = 0;  // over the k - 1 bases before the stretch (used only when they are all unambiguous)
  for (int64_t j = pos - (k - 1) > t.c_lo ? pos - (k - 1) : t.c_lo; j < pos; ++j) {
    const uint64_t c = CD[j - t.c_lo] & 3;
    fwd = (fwd << 2 | c) & mask;
    rev = rev >> 2 | (3 ^ c) << shift1;
  }
  for (int a = a0; a < a1; ++a, ++pos) {
    const uint64_t c = CD[pos - t.c_lo];
    uint64_t x = kNone;
    int z = 0;
    if (c < 4) {
      l = l < cap_l ? l + 1 : l;
      fwd = (fwd << 2 | c) & mask;
      rev = rev >> 2 | (3 ^ c) << shift1;
      if (l >= k) {
        z = fwd <= rev ? 0 : 1;
        x = hash(z ? rev : fwd, mask);
      }
    } else l = 0;
    X[a] = x, EL[a] = (uint16_t)(l | z << 15);
  }
}

// Sliding-window minima in two steps (van Herk / Gil-Werman): the slots are cut into blocks of w; inside each block P[a] is
// the oldest of the smallest over [block start, a] and S[a] over [a, block end].  A window of w slots is the tail of one block
// plus the head of the next, so its minimum is one comparison (window_min) instead of a scan of the window.  One task per
// (block, direction).  Slots need 13 bits (kTile + 2 kMaxW < 8192); the bits above carry what window_min would otherwise
// have to recompute: kDup = the smallest value occurs more than once in the scanned range, kWhole (P only) = the window
// ending at this slot lies inside the block (its last slot, or any slot of block 0, where windows are cut by the start).
constexpr uint16_t kSlotMask = 0x1fff, kWhole = 0x4000, kDup = 0x8000;
MASH_HD void scan_blocks(const Tile &t, int w, int tid, int n_threads, const uint64_t *X, uint16_t *P, uint16_t *S) {
  const int nblk = (t.n_ev + w - 1) / w;
  for (int task = tid; task < 2 * nblk; task += n_threads) {
    const int lo = (task >> 1) * w, hi = lo + w < t.n_ev ? lo + w : t.n_ev;
    if (task & 1) {  // from the block's newest slot down: an equal value further down is older and wins
      int best = hi - 1;
      uint64_t xb = X[best];
      uint16_t dup = 0;
      S[best] = (uint16_t)best;
      for (int a = hi - 2; a >= lo; --a) {
        const uint64_t x = X[a];
        if (x <= xb) dup = x == xb ? kDup : 0, xb = x, best = a;
        S[a] = (uint16_t)(best | dup);
      }
    } else {
      int best = lo;
      uint64_t xb = X[lo];
      uint16_t dup = 0;
      const uint16_t whole0 = lo == 0 ? kWhole : 0;
      P[lo] = (uint16_t)(lo | whole0 | (w == 1 ? kWhole : 0));
      for (int a = lo + 1; a < hi; ++a) {
        const uint64_t x = X[a];
        if (x < xb) xb = x, best = a, dup = 0;
        else if (x == xb) dup = kDup;
        P[a] = (uint16_t)(best | dup | whole0 | (a == lo + w - 1 ? kWhole : 0));
      }
    }
  }
}
// Slot of the oldest of the smallest values over the w slots ending at a (fewer at the start of the sequence); `dup` =
// that value occurs at another slot of the window as well.
MASH_HD int window_min(int a, int w, const uint64_t *X, const uint16_t *P, const uint16_t *S, bool &dup) {
  const uint16_t pa = P[a];
  const int p = pa & kSlotMask;
  if (pa & kWhole) {
    dup = (pa & kDup) != 0;
    return p;
  }
  const uint16_t sl = S[a - w + 1];  // the window starts in the block before
  const int s = sl & kSlotMask;
  const uint64_t xp = X[p], xs = X[s];
  dup = xp == xs || (((xp < xs ? pa : sl) & kDup) != 0);
  return xp < xs ? p : s;
}

// Thread `tid`: the decisions of its share of the deciding positions [t0, e_hi); F[slot] = 1 for every position the
// reference's scan appends (several threads may set the same flag).
MASH_HD void decide(const Tile &t, int w, int k, int tid, int n_threads, const uint64_t *X, const uint16_t *EL, const uint16_t *P,
                    const uint16_t *S, uint8_t *F) {
  const int n_dec = (int)(t.e_hi - t.t0);
  for (int q = tid; q < n_dec; q += n_threads) {
    const int64_t ee = t.t0 + q;             // the position
    const int a = (int)(ee - t.e_lo);        // its slot
    const int lo_new = (int)((ee - w + 1 > 0 ? ee - w + 1 : 0) - t.e_lo);
    const int jm = (int)(ee - w - t.e_lo);   // slot of the position that leaves the window at this step (negative: none yet)
    // oldest of the smallest over [ee-w, ee-1] (the scan's minimum before this step) and over [ee-w+1, ee] (after it)
    int pm = -1;
    uint64_t xpm = kNone;
    bool dup_pm = false, dup_nm = false;
    if (a > 0) pm = window_min(a - 1, w, X, P, S, dup_pm), xpm = X[pm];
    const int nm = window_min(a, w, X, P, S, dup_nm);
    const uint64_t xnm = X[nm], xe = X[a];
    const int le = EL[a] & 0x7fff;
    // first full window: the other positions holding the minimum's value (the slot that just left the window held no
    // value -- the run was k - 1 long there -- so there are none unless the old window had the value twice or this slot has it)
    if (le == w + k - 1 && xpm != kNone && (dup_pm || xe == xpm))
      for (int j = lo_new; j <= a; ++j)
        if (X[j] == xpm && j != pm) F[j] = 1;
    if (xe < xpm) {  // a smaller value displaces the minimum
      if (le >= w + k && xpm != kNone) F[pm] = 1;
    } else if (xpm != kNone && pm == jm) {  // the minimum leaves the window
      if (le >= w + k - 1) {
        F[pm] = 1;
        if (xnm != kNone && dup_nm)  // (the scan below finds nothing unless the value occurs twice in the window)
          for (int j = lo_new; j <= a; ++j)
            if (X[j] == xnm && j != nm) F[j] = 1;
      }
    }
    if (ee == t.L - 1 && xnm != kNone) F[nm] = 1;  // the minimum standing at the end of the sequence
  }
}

}  // namespace mash
}  // namespace pgmm
