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
  std::vector<uint64_t> X((size_t)cap_ev(w));
  std::vector<uint16_t> EL((size_t)cap_ev(w));
  std::vector<uint8_t> F((size_t)cap_ev(w)), CD((size_t)cap_codes(w, k));
  int64_t n = 0;
  for (int64_t t0 = 0; t0 < L; t0 += kTile) {
    const Tile t = tile_of(L, t0, w, k);
    if (t.n_ev > cap_ev(w) || t.n_codes > cap_codes(w, k)) return -2;
    for (int i = 0; i < t.n_codes; ++i) CD[(size_t)i] = (uint8_t)code((uint8_t)seq[t.c_lo + i]);
    std::fill(F.begin(), F.end(), 0);
    std::fill(X.begin(), X.end(), 0x5555555555555555ull);  // slots past n_ev must never be read
    for (int tid = 0; tid < kTileThreads; ++tid) roll(t, w, k, tid, kTileThreads, CD.data(), X.data(), EL.data());
    for (int tid = 0; tid < kTileThreads; ++tid) decide(t, w, k, tid, kTileThreads, X.data(), EL.data(), F.data());
    const int base = (int)(t.t0 - t.e_lo);
    for (int j = 0; j < (int)(t.t1 - t.t0); ++j)
      if (F[(size_t)(base + j)]) {
        if (n < cap) {
          value[n] = X[(size_t)(base + j)];
          position[n] = id << 32 | (uint64_t)(t.t0 + j + 1) << 1 | (uint64_t)(EL[(size_t)(base + j)] >> 15);
        }
        ++n;
      }
  }
  return n;
}
