// K1-K3: minimizer sketch, index build, seed lookup and anchor expansion on the GPU.
// Replaces, for pangraph's path: mm_sketch (minimap2/sketch.c:77-143), mm_idx_str's index construction
// (index.c:213-265,408-456), mm_idx_cal_max_occ (index.c:186-207), mm_seed_mz_flt / mm_collect_matches /
// mm_seed_select (seed.c) and collect_seed_hits without its final sort (map.c:168-204).
#pragma once
#include <atomic>
#include <cstdint>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "mapper.h"
#include "pgmm_cuda.h"

namespace pgmm {

extern uint64_t g_seed_launches;  // kernels launched by the sketch / index / seeding stages (ours + CUB)
extern std::atomic<uint64_t> g_anchor_sorted_device, g_anchor_sorted_host;  // queries by where their anchor order came from

// A set of sequences whose coded bases (0..4) sit in one device buffer, plus their minimizers.
struct DeviceSeqSet {
  int n = 0;
  uint64_t total = 0;              // sum of lengths = positions of the virtual concatenation
  const uint8_t *d_codes = nullptr;  // not owned
  std::vector<uint64_t> h_starts, h_vstart;
  std::vector<int> h_lens;
  DevBuf<uint64_t> starts, vstart;
  DevBuf<int> lens;
  DevBuf<uint64_t> mx, my;         // minimizers: x = hash<<8|span, y = seq<<32 | lastPos<<1 | strand, in emission order
  DevBuf<uint64_t> mz_off;         // [n+1] first minimizer of each sequence
  uint64_t n_mz = 0;
  std::vector<uint64_t> h_mz_off;
};

// The lookup structure over the target minimizers: distinct hashes ascending, each with its ascending position list.
struct DeviceIndex {
  int w = 0, k = 0;
  DeviceSeqSet seqs;
  DevBuf<uint64_t> keys;        // [n_keys]
  DevBuf<uint32_t> key_off;     // [n_keys+1] range of each key in pos
  DevBuf<uint64_t> pos;         // [n_mz]
  uint64_t n_keys = 0;
  DevBuf<uint32_t> seq_len;     // [n]
  DevBuf<int32_t> name_rank;    // [n] rank of each target name in strcmp order (MM_F_NO_DUAL / MM_F_NO_DIAG)
  DevBuf<uint32_t> occ_sorted;  // [n_keys] occurrence counts ascending (for mid_occ)
  // the probe structure (the reference's khash, index.c:81-98): open addressing, linear probing, load factor <= 1/2;
  // a slot holds the key's rank in `keys` (its position list is pos[key_off[rank] .. key_off[rank+1]))
  DevBuf<uint64_t> ht_key;      // [ht_mask + 1], ~0 = empty
  DevBuf<uint32_t> ht_rank;     // [ht_mask + 1]
  uint64_t ht_mask = 0;
};

class SeedEngine {
 public:
  struct Impl;
  SeedEngine();
  ~SeedEngine();
  // K1: sketches lens.size() sequences; sequence i occupies d_codes[starts[i] .. starts[i]+lens[i])
  void sketch(const uint8_t *d_codes, const std::vector<uint64_t> &starts, const std::vector<int> &lens, int w, int k,
              DeviceSeqSet &set, cudaStream_t st);
  // K2: sort + run-length encode idx.seqs' minimizers
  void build_index(DeviceIndex &idx, const std::vector<uint32_t> &lens, const std::vector<int32_t> &name_rank, cudaStream_t st);
  // (1-f) quantile of the occurrence counts, plus one (index.c:186-207)
  static int32_t cal_max_occ(const DeviceIndex &idx, float f, cudaStream_t st);
  // K3: occurrence filter, index probe, streak thinning, anchor expansion for every query of qs
  void collect(const DeviceIndex &idx, const DeviceSeqSet &qs, const std::vector<int32_t> &q_name_rank, const mm_mapopt_t &opt,
               std::vector<QuerySeeds> &out, cudaStream_t st);

 private:
  Impl *impl_;
};

}  // namespace pgmm
