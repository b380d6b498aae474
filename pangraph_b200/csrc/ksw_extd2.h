// K5: batched dual-affine banded extension/global DP with traceback on the GPU.
// Replaces ksw_extd2_sse (reference: packages/minimap2-sys/minimap2/ksw2_extd2_sse.c:34-401) for pangraph's
// alignment path; one CTA per DP problem, anti-diagonal wavefront, int8x4 packed difference recurrence.
#pragma once
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

namespace pgmm {

// flag bits of a DP job: the low bits are the reference's KSW_EZ_* (ksw2.h:8-14)
enum : int32_t {
  KSW_RIGHT = 0x02,
  KSW_APPROX_MAX = 0x08,
  KSW_EXTZ_ONLY = 0x40,
  KSW_REV_CIGAR = 0x80,
  KSW_JOB_REVSEQ = 0x10000,  // ours: read query and target windows back to front (left extension, align.c:710-712)
};

constexpr int32_t KSW_NEG_INF = -0x40000000;

struct KswJob {
  uint64_t q_off;    // first base of the query window in the query code buffer
  uint64_t t_off;    // first base of the target window in the target code buffer
  uint64_t p_off;    // byte offset of this job's traceback matrix in the traceback arena
  uint64_t cig_off;  // u32 offset of this job's CIGAR scratch (capacity qlen+tlen+2)
  uint64_t scr_off;  // byte offset of global DP state when it does not fit shared memory, else ~0
  int32_t qlen, tlen;
  int32_t w, zdrop, end_bonus, flag;
};

struct KswOut {  // ksw_extz_t minus the pointers (ksw2.h:31-40)
  int32_t max, zdropped, max_q, max_t, mqe, mqe_t, mte, mte_q, score, reach_end, n_cigar;
  int32_t n_diag;    // anti-diagonals actually evaluated (profiling aid; not part of the reference's result)
  uint32_t cig_pos;  // where the kernel put this problem's CIGAR in the wave's packed output
  // mm_test_zdrop's scan of the CIGAR (align.c:47-68), filled for first-pass fills only: the largest score drop along
  // the path and the window (target from/to, query from/to) where it happens
  int32_t zd_max, zd_t0, zd_t1, zd_q0, zd_q1;
};

struct KswScoring {  // mm_mapopt_t a,b,sc_ambi -> ksw_gen_simple_mat (align.c:9-22); q,e,q2,e2
  int8_t sc_mch, sc_mis, sc_ambi;
  int8_t q, e, q2, e2;
};

// Bytes of traceback a job needs, its row stride, and its DP-state footprint (shared memory or global scratch).
struct KswGeom {
  int32_t T;        // padded target length (multiple of 16)
  int32_t n_col16;  // row stride of the traceback matrix in bytes
  int64_t p_bytes;
  int64_t state_bytes;
};
KswGeom ksw_geometry(int qlen, int tlen, int w, int flag);

// Launches the DP for `jobs` (device-resident descriptors are built inside). qcodes/tcodes are device pointers to
// 0..4 coded bases. Results land in host vectors. Blocking on `stream`.
struct KswBatchResult {
  std::vector<KswOut> out;
  std::vector<uint32_t> cigar;      // concatenated, job i at cig_start[i] .. + out[i].n_cigar
  std::vector<uint64_t> cig_start;
  uint64_t cells = 0;               // in-band cells actually asked for (for the GCUPS / roofline figures)
  int launches = 0;
  float kernel_ms = 0.f;            // CUDA-event time of the DP kernels alone (all launch classes of a wave, concurrent)
  // per kernel family (0 = K5 generic band-limited, 1 = K5a small first-pass fills, 2 = K5b wide unbanded fills):
  // CUDA-event time of every launch on the stream it is launched on, cells, bases read, launches
  float fam_ms[3] = {0.f, 0.f, 0.f};
  uint64_t fam_cells[3] = {0, 0, 0}, fam_bases[3] = {0, 0, 0};
  int fam_launches[3] = {0, 0, 0};
};

class KswEngine {
 public:
  KswEngine();
  ~KswEngine();
  // Runs all jobs (fields p_off/cig_off/scr_off are assigned here). Jobs are split into memory-bounded waves.
  void run(std::vector<KswJob> &jobs, const uint8_t *d_qcodes, const uint8_t *d_tcodes, const KswScoring &sc,
           KswBatchResult &res, cudaStream_t stream);
  size_t arena_budget_bytes = size_t(24) << 30;  // traceback arena per wave

 private:
  struct Impl;
  Impl *impl_;
};

}  // namespace pgmm
