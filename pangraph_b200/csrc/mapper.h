// Host side of the mapping pipeline: what the reference does per query inside mm_map (map.c:227-374), re-organised
// so that the data-parallel stages (sketch, index probe, anchor expansion, DP) run as whole-batch GPU launches and the
// inherently sequential ones (anchor sort order, chaining, region bookkeeping, CIGAR stitching) run on host threads.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "../../include/pgmm_b200.h"
#include "chain.h"
#include "flag_sort.h"
#include "ksw_extd2.h"

namespace pgmm {

constexpr uint64_t SEED_LONG_JOIN = 1ULL << 40, SEED_IGNORE = 1ULL << 41, SEED_TANDEM = 1ULL << 42, SEED_SELF = 1ULL << 43;

// Host view of the indexed (target) sequences.
struct TargetSet {
  int k = 0, w = 0;
  std::vector<std::string> names;
  std::vector<uint32_t> lens;
  std::vector<uint64_t> offs;   // offset of each sequence in codes
  std::vector<uint8_t> codes;   // 0..4 per base, all sequences concatenated
};

// One batch of queries: ASCII in, coded forward and reverse-complement copies kept on both sides.
// std::vector<uint8_t>::resize zero-fills; the 20 MB query buffer of a 5-Mbp pair is overwritten right away
template <class T>
struct DefaultInitAlloc : std::allocator<T> {
  template <class U>
  struct rebind {
    using other = DefaultInitAlloc<U>;
  };
  template <class U, class... Args>
  void construct(U *p, Args &&...args) {
    if constexpr (sizeof...(Args) == 0) ::new ((void *)p) U;
    else ::new ((void *)p) U(std::forward<Args>(args)...);
  }
};

struct QueryBatch {
  int n = 0;
  std::vector<const char *> seqs, names;
  std::vector<int> lens;
  std::vector<uint64_t> base;  // query i: forward codes at codes[base[i] .. +len), reverse complement right after
  std::vector<uint8_t, DefaultInitAlloc<uint8_t>> codes;
  bool from_targets = false;   // the queries ARE the indexed sequences (pangraph's all-vs-all round): seqs is unused and
                               // the device builds its query buffer from the resident target codes
};

// What the seeding stage hands to the host for one query (map.c:168-204 minus the final sort).
struct QuerySeeds {
  std::vector<U128> a;             // anchors in collection order (seed-major, hits ascending)
  std::vector<uint64_t> mini_pos;  // q_span<<32 | q_pos of the seeds that were used (seed.c:125)
  int rep_len = 0;                 // bases covered by filtered, repetitive seeds (seed.c:113-128)
  // true: `a` arrives sorted by target position and holds no two anchors with the same one, so it IS what the reference's
  // radix_sort_128x (map.c:202) leaves -- the order of equal keys under that unstable sort is the only thing a stable device
  // sort cannot reproduce, and there are none.  false: collection order, the exact replay on the host sorts it.
  bool sorted = false;
};

struct DpStats {
  uint64_t jobs = 0, cells = 0, waves = 0, seq_bytes = 0;
  uint64_t fam_cells[3] = {0, 0, 0}, fam_bases[3] = {0, 0, 0}, fam_launches[3] = {0, 0, 0};  // per DP kernel family
  double fam_ms[3] = {0, 0, 0};
  int launches = 0;
  double kernel_ms = 0;
  // wall-clock phases of map_batch (ms): encode, seeding (GPU), anchor sort + chain + plan (host), DP waves (GPU incl.
  // copies), host work between waves, final filters + output
  double t_encode = 0, t_seed = 0, t_chain = 0, t_dp = 0, t_stitch = 0, t_final = 0;
  // t_chain split: anchor sort + segments (host), score fill (device, incl. copies), segments redone on the host +
  // backtrack + hit skeletons + first DP plan (host)
  double t_chain_sort = 0, t_chain_fill = 0, t_chain_rest = 0;
  ChainFillStats chain;
};

// The device stages.  The product has exactly one implementation (CUDA, cuda_backend.cu); a second one exists only in
// the CPU-only test library, where it forwards to the reference's own C so that the host logic can be checked without
// a GPU (tests/hostlogic_backend.cpp).
struct Backend {
  virtual ~Backend() {}
  virtual void begin_batch(const TargetSet &ts, const QueryBatch &qb) = 0;
  virtual void seed_batch(const TargetSet &ts, const QueryBatch &qb, const mm_mapopt_t &opt, std::vector<QuerySeeds> &out) = 0;
  // score fill of the chaining stage for every query of the batch (chain.h); segments the device hands back are
  // marked in jobs[i].redo and filled by the caller with chain_fill_host
  virtual void chain_fill(const ChainParams &cp, std::vector<ChainFillJob> &jobs) = 0;
  // q_off indexes qb.codes, t_off indexes ts.codes
  virtual void run_dp(std::vector<KswJob> &jobs, const KswScoring &sc, KswBatchResult &res) = 0;
  // the chaining results (ChainFillJob::f/p/v) are no longer needed: whatever the seeding and chaining stages hold may go
  virtual void end_chain() {}
  virtual void end_batch() {}
  DpStats stats;
};

// Host CPU time by phase (thread CPU clocks, summed over every thread that works for the path): [0] encode, [1] seeding
// (host side), [2] anchor sort + segments, [3] chain fill (host side), [4] host refill + backtrack + hits + plan,
// [5] DP waves: job lists, hand-over, result slices (round side), [6] DP service workers, [7] stitching + next plan,
// [8] final filters and output, [9] index build (host side)
constexpr int kCpuPhases = 10;
void cpu_phase_add(int phase, uint64_t ns);
void cpu_phase_read(double *ms_out, bool reset);
struct CpuScope {
  int phase;
  uint64_t t0;
  static uint64_t now();
  explicit CpuScope(int p) : phase(p), t0(now()) {}
  ~CpuScope() { cpu_phase_add(phase, now() - t0); }
};

// Maps every query of the batch; n_regs[i]/regs[i] are malloc()-owned like mm_map's result.
void map_batch(Backend &be, const TargetSet &ts, QueryBatch &qb, const mm_mapopt_t &opt, int *n_regs, mm_reg1_t **regs,
               int n_threads);

// ASCII -> 0..4 (A,C,G,T/U in either case -> 0..3, everything else 4; sketch.c:9-26)
extern const uint8_t kNt4[256];
void encode_queries(QueryBatch &qb, const TargetSet &ts, int n_threads);

// --- pieces exposed for stage-level tests ---
// stop_at > 0: return as soon as the score reaches it (the end coordinates are then meaningless)
int ll_local_score(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat, int gapo, int gape,
                   int *qe, int *te, int stop_at = 0);

}  // namespace pgmm
