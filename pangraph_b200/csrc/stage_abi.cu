// Stage-level C-ABI entry points: each runs ONE device stage on host buffers so that tests can compare it with the
// oracle in isolation.  Declared in include/pgmm_b200.h.
#include "../../include/pgmm_b200.h"
#include "ksw_extd2.h"
#include "pgmm_cuda.h"

#include <cstring>
#include <vector>

using namespace pgmm;

extern "C" int pgmm_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

extern "C" int pgmm_ksw_extd2_batch(int n, const int32_t *qlen, const int32_t *tlen, const uint64_t *q_off,
                                    const uint64_t *t_off, const uint8_t *qcodes, uint64_t q_total,
                                    const uint8_t *tcodes, uint64_t t_total, const int32_t *w, const int32_t *zdrop,
                                    const int32_t *end_bonus, const int32_t *flag, int a, int b, int sc_ambi, int q,
                                    int e, int q2, int e2, int32_t *out_ez, uint32_t *out_cigar, uint64_t cigar_cap,
                                    uint64_t *out_cig_start, double *out_kernel_ms, uint64_t arena_budget_bytes) {
  require_device();
  std::vector<KswJob> jobs(n);
  for (int i = 0; i < n; ++i) {
    KswJob &j = jobs[i];
    memset(&j, 0, sizeof(j));
    j.q_off = q_off[i], j.t_off = t_off[i], j.qlen = qlen[i], j.tlen = tlen[i];
    j.w = w[i], j.zdrop = zdrop[i], j.end_bonus = end_bonus[i], j.flag = flag[i];
  }
  DevBuf<uint8_t> dq, dt;
  dq.ensure(q_total + 16), dt.ensure(t_total + 16);
  cudaStream_t st;
  PGMM_CUDA(cudaStreamCreate(&st));
  PGMM_CUDA(cudaMemcpyAsync(dq.p, qcodes, q_total, cudaMemcpyHostToDevice, st));
  PGMM_CUDA(cudaMemcpyAsync(dt.p, tcodes, t_total, cudaMemcpyHostToDevice, st));
  KswScoring sc;
  sc.sc_mch = (int8_t)(a < 0 ? -a : a), sc.sc_mis = (int8_t)(b > 0 ? -b : b), sc.sc_ambi = (int8_t)sc_ambi;
  sc.q = (int8_t)q, sc.e = (int8_t)e, sc.q2 = (int8_t)q2, sc.e2 = (int8_t)e2;
  KswEngine eng;
  if (arena_budget_bytes) eng.arena_budget_bytes = arena_budget_bytes;
  KswBatchResult res;
  eng.run(jobs, dq.p, dt.p, sc, res, st);
  PGMM_CUDA(cudaStreamDestroy(st));
  if (res.cigar.size() > cigar_cap) return -1;
  for (int i = 0; i < n; ++i) {
    memcpy(out_ez + 11 * (size_t)i, &res.out[i], 11 * sizeof(int32_t));
    out_cig_start[i] = res.cig_start[i];
  }
  if (!res.cigar.empty()) memcpy(out_cigar, res.cigar.data(), res.cigar.size() * sizeof(uint32_t));
  if (out_kernel_ms) *out_kernel_ms = res.kernel_ms;
  return 0;
}
