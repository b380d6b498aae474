// TEST ONLY: the schedule of K7's tile kernel (pangraph_b200/csrc/mash.cu) replayed serially on the CPU -- the same
// mash_core.h functions, called for thread 0..255 of every tile, with the kernel's buffers on the heap.  It checks the
// tile / halo arithmetic and the per-position decisions against the sequential restatement of the reference without a GPU.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../pangraph_b200/csrc/mash_core.h"

using namespace pgmm::mash;

// -> number of flagged positions; value[i], position[i] (id << 32 | locus << 1 | strand, locus 1-based) in position order
extern "C" __attribute__((visibility("default"))) int64_t mash_emul_sketch(const char *seq, int64_t L, uint64_t id, int k, int w,
                                                                           uint64_t *value, uint64_t *position, int64_t cap) {
  if (k < 1 || k > kMaxK || w < 1 || w > kMaxW) return -1;
  const Layout lay = layout_of(w, k);
  std::vector<uint64_t> smem((size_t)lay.total / 8 + 1);  // the kernel's shared memory, same offsets
  uint8_t *sm = (uint8_t *)smem.data();
  uint64_t *X = (uint64_t *)(sm + lay.x);
  uint16_t *EL = (uint16_t *)(sm + lay.el), *P = (uint16_t *)(sm + lay.p), *S = (uint16_t *)(sm + lay.s);
  uint8_t *F = sm + lay.f, *CD = sm + lay.cd;
  int64_t n = 0;
  for (int64_t t0 = 0; t0 < L; t0 += kTile) {
    const Tile t = tile_of(L, t0, w, k);
    if (t.n_ev > cap_ev(w) || t.n_codes > cap_codes(w, k)) return -2;
    memset(sm, 0x55, lay.total);  // nothing may depend on what an earlier tile left behind
    const int pre = (int)(t.c_lo & 15);  // the kernel's staging starts at a 16-byte boundary of the source
    for (int i = 0; i < t.n_codes; ++i) CD[(size_t)(pre + i)] = (uint8_t)code((uint8_t)seq[t.c_lo + i]);
    memset(F, 0, (size_t)cap_ev(w));
    for (int tid = 0; tid < kTileThreads; ++tid) roll(t, w, k, tid, kTileThreads, CD + pre, X, EL);
    for (int tid = 0; tid < kTileThreads; ++tid) scan_blocks(t, w, tid, kTileThreads, X, P, S);
    for (int tid = 0; tid < kTileThreads; ++tid) decide(t, w, k, tid, kTileThreads, X, EL, P, S, F);
    const int base = (int)(t.t0 - t.e_lo);
    for (int j = 0; j < (int)(t.t1 - t.t0); ++j)
      if (F[base + j]) {
        if (n < cap) {
          value[n] = X[base + j];
          position[n] = id << 32 | (uint64_t)(t.t0 + j + 1) << 1 | (uint64_t)(EL[base + j] >> 15);
        }
        ++n;
      }
  }
  return n;
}
