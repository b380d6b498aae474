// Stage-level C-ABI entry points: each runs ONE device stage on host buffers so that tests can compare it with the
// oracle in isolation.  Declared in include/pgmm_b200.h.
#include "../../include/pgmm_b200.h"
#include "chain_fill.h"
#include "fasta.h"
#include "guide_tree.h"
#include "ksw_extd2.h"
#include "mash.h"
#include "nextalign.h"
#include "pgmm_cuda.h"

#include <algorithm>
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

// K4 alone: chains n sorted anchors (2 uint64 each) like mg_lchain_rmq.  The device fills the scores; segments it hands
// back are filled by the host arbiter unless host_redo == 0, in which case -2 is returned when any segment was handed
// back.  Returns n_u; u[] gets score<<32|count per chain, xy the compacted anchors, n_a_out their number;
// out_fpv (3n int32, optional) gets f, p, v as filled; seg_stats (optional) = {segments, redone segments, redone anchors}.
extern "C" int64_t pgmm_chain_rmq(uint64_t *xy, int64_t n, int max_dist, int max_dist_inner, int bw, int max_skip, int cap,
                                  int min_cnt, int min_sc, float pen_gap, float pen_skip, uint64_t *u, int64_t *n_a_out,
                                  int32_t *out_fpv, int64_t *seg_stats, int host_redo) {
  require_device();
  std::vector<U128> a((U128 *)xy, (U128 *)xy + n);
  ChainParams cp{max_dist, max_dist_inner, bw, max_skip, cap, min_cnt, min_sc, pen_gap, pen_skip};
  std::vector<ChainFillJob> jobs(1);
  jobs[0].a = a.data(), jobs[0].n = n;
  chain_find_segments(cp, a.data(), n, jobs[0].segs);
  cudaStream_t st;
  PGMM_CUDA(cudaStreamCreate(&st));
  ChainEngine eng;
  ChainFillStats cs;
  eng.run(cp, jobs, st, &cs);
  PGMM_CUDA(cudaStreamDestroy(st));
  if (seg_stats) seg_stats[0] = (int64_t)cs.segments, seg_stats[1] = (int64_t)cs.redo_segments, seg_stats[2] = (int64_t)cs.redo_anchors;
  *n_a_out = 0;
  if (n == 0) return 0;
  if (cs.redo_segments && !host_redo) return -2;
  std::vector<int32_t> t((size_t)n, 0);
  for (size_t k = 0; k < jobs[0].segs.size(); ++k)
    if (jobs[0].redo[k]) chain_fill_host(cp, a.data(), n, jobs[0].segs[k].start, jobs[0].segs[k].end, jobs[0].f, jobs[0].p, jobs[0].v, t.data());
  if (out_fpv) {
    memcpy(out_fpv, jobs[0].f, (size_t)n * 4);
    memcpy(out_fpv + n, jobs[0].p, (size_t)n * 4);
    memcpy(out_fpv + 2 * n, jobs[0].v, (size_t)n * 4);
  }
  std::vector<uint64_t> uu;
  chain_backtrack(cp, a, jobs[0].f, jobs[0].p, jobs[0].v, t.data(), uu);
  memcpy(xy, a.data(), a.size() * 16);
  memcpy(u, uu.data(), uu.size() * 8);
  *n_a_out = (int64_t)a.size();
  return (int64_t)uu.size();
}

// ---- CTA trace (pgmm_cuda.h): begin allocates a device buffer of `capacity` records and attaches it to the DP and
// chaining kernels; end detaches, copies up to max_n records (8 uint64 each: t0, t1, t2, kernel | block << 32,
// smid | aux << 32 ... as laid out in PgmmCtaTraceRec) and returns how many were produced.
static PgmmCtaTraceRec *g_trace_buf = nullptr;
static unsigned long long *g_trace_cnt = nullptr;
static unsigned long long g_trace_cap = 0;
extern "C" int pgmm_cta_trace_begin(uint64_t capacity) {
  require_device();
  if (g_trace_buf) return -1;
  PGMM_CUDA(cudaMalloc((void **)&g_trace_buf, capacity * sizeof(PgmmCtaTraceRec)));
  PGMM_CUDA(cudaMalloc((void **)&g_trace_cnt, sizeof(unsigned long long)));
  PGMM_CUDA(cudaMemset(g_trace_cnt, 0, sizeof(unsigned long long)));
  g_trace_cap = capacity;
  trace_attach_ksw(g_trace_buf, g_trace_cnt, capacity);
  trace_attach_chain(g_trace_buf, g_trace_cnt, capacity);
  return 0;
}
extern "C" int64_t pgmm_cta_trace_end(void *out, uint64_t max_n) {
  if (!g_trace_buf) return -1;
  PGMM_CUDA(cudaDeviceSynchronize());
  trace_attach_ksw(nullptr, nullptr, 0);
  trace_attach_chain(nullptr, nullptr, 0);
  unsigned long long n = 0;
  PGMM_CUDA(cudaMemcpy(&n, g_trace_cnt, sizeof(n), cudaMemcpyDeviceToHost));
  const unsigned long long take = std::min<unsigned long long>(std::min<unsigned long long>(n, g_trace_cap), max_n);
  if (take) PGMM_CUDA(cudaMemcpy(out, g_trace_buf, take * sizeof(PgmmCtaTraceRec), cudaMemcpyDeviceToHost));
  cudaFree(g_trace_buf), cudaFree(g_trace_cnt);
  g_trace_buf = nullptr, g_trace_cnt = nullptr;
  return (int64_t)n;
}

// ---- map_variations (Part 4) ----
template <class T>
static T *dup_array(const T *src, size_t n) {
  T *p = (T *)malloc(std::max<size_t>(1, n) * sizeof(T));
  if (n) memcpy(p, src, n * sizeof(T));
  return p;
}

extern "C" int pgmm_map_variations_batch(int n, const char *const *refs, const int32_t *ref_lens, const char *const *qrys,
                                         const int32_t *qry_lens, const int32_t *mean_shift, const int32_t *band_width,
                                         int extra_band_width, int max_alignment_attempts, pgmm_edit_t **out, double *stats) {
  require_device();
  std::vector<na::Problem> probs((size_t)std::max(0, n));
  for (int i = 0; i < n; ++i) probs[i] = na::Problem{refs[i], qrys[i], ref_lens[i], qry_lens[i], mean_shift[i], band_width[i]};
  std::vector<na::Edit> edits;
  na::Stats st;
  na::run_batch(probs, extra_band_width, max_alignment_attempts, edits, &st);
  pgmm_edit_t *res = (pgmm_edit_t *)calloc((size_t)std::max(1, n), sizeof(pgmm_edit_t));
  for (int i = 0; i < n; ++i) {
    const na::Edit &e = edits[i];
    pgmm_edit_t &r = res[i];
    r.status = e.status, r.hit_boundary = e.hit_boundary, r.attempts = e.attempts, r.band_width = e.band_width, r.score = e.score;
    r.n_sub = (int32_t)e.sub_pos.size(), r.n_del = (int32_t)e.del_pos.size(), r.n_ins = (int32_t)e.ins_pos.size();
    r.sub_pos = dup_array(e.sub_pos.data(), e.sub_pos.size()), r.sub_chr = dup_array(e.sub_chr.data(), e.sub_chr.size());
    r.del_pos = dup_array(e.del_pos.data(), e.del_pos.size()), r.del_len = dup_array(e.del_len.data(), e.del_len.size());
    r.ins_pos = dup_array(e.ins_pos.data(), e.ins_pos.size()), r.ins_len = dup_array(e.ins_len.data(), e.ins_len.size());
    r.ins_seq = dup_array(e.ins_seq.data(), e.ins_seq.size());
  }
  *out = res;
  if (stats) stats[0] = st.kernel_ms, stats[1] = (double)st.cells, stats[2] = (double)st.problems, stats[3] = (double)st.launches;
  return 0;
}

extern "C" void pgmm_edits_free(pgmm_edit_t *edits, int n) {
  if (!edits) return;
  for (int i = 0; i < n; ++i) {
    free(edits[i].sub_pos), free(edits[i].sub_chr), free(edits[i].del_pos), free(edits[i].del_len);
    free(edits[i].ins_pos), free(edits[i].ins_len), free(edits[i].ins_seq);
  }
  free(edits);
}

// ---------------- Part 5: the guide tree ----------------
extern "C" int pgmm_mash_distance(int n, const char *const *seqs, const int64_t *lens, int k, int w, double *dist, double *stats) {
  if (n < 1) return -2;
  std::vector<uint32_t> counts((size_t)n * n);
  mash::Stats st;
  if (const int rc = mash::shared_counts(seqs, lens, n, k, w, counts.data(), &st)) return rc;
  if (stats) {
    stats[0] = st.upload_ms, stats[1] = st.sketch_ms, stats[2] = st.sort_ms, stats[3] = st.pair_ms, stats[4] = (double)st.bases;
    stats[5] = (double)st.tiles, stats[6] = (double)st.minimizers, stats[7] = (double)st.unique_keys, stats[8] = (double)st.shared_values;
    stats[9] = (double)st.launches;
  }
  return gt::distances_from_counts(counts.data(), n, dist);
}

extern "C" int pgmm_nj_tree(int n, const double *dist, int32_t *left, int32_t *right) {
  gt::Tree t;
  if (const int rc = gt::neighbor_joining(dist, n, t)) return rc;
  std::copy(t.left.begin(), t.left.end(), left), std::copy(t.right.begin(), t.right.end(), right);
  return 0;
}

extern "C" int pgmm_newick_parse(const char *text, int32_t *n_leaves, char **names, int64_t *names_bytes, int32_t **left, int32_t **right,
                                 char *err, int err_cap) {
  gt::Tree t;
  std::vector<std::string> leaf_names;
  std::string msg;
  *n_leaves = 0, *names = nullptr, *names_bytes = 0, *left = nullptr, *right = nullptr;
  if (!gt::parse_newick(text ? text : "", t, leaf_names, msg)) {
    if (err && err_cap > 0) snprintf(err, (size_t)err_cap, "%s", msg.c_str());
    return -1;
  }
  std::string joined;
  for (const std::string &s : leaf_names) joined += s, joined.push_back('\0');
  *n_leaves = t.n, *names_bytes = (int64_t)joined.size();
  *names = dup_array(joined.data(), joined.size());
  *left = dup_array(t.left.data(), t.left.size()), *right = dup_array(t.right.data(), t.right.size());
  return 0;
}

static gt::Tree tree_of(int n, const int32_t *left, const int32_t *right) {
  gt::Tree t;
  t.n = n;
  if (n > 1) t.left.assign(left, left + (n - 1)), t.right.assign(right, right + (n - 1));
  return t;
}

extern "C" char *pgmm_newick_write(int n, const char *const *names, const int32_t *left, const int32_t *right) {
  std::vector<std::string> leaf_names;
  for (int i = 0; i < n; ++i) leaf_names.emplace_back(names[i]);
  const std::string s = gt::to_newick(tree_of(n, left, right), leaf_names);
  return dup_array(s.c_str(), s.size() + 1);
}

extern "C" int pgmm_tree_balance(int n, const int32_t *left, const int32_t *right, int32_t *out_left, int32_t *out_right) {
  if (n < 1) return -1;
  const gt::Tree b = gt::balance(tree_of(n, left, right));
  std::copy(b.left.begin(), b.left.end(), out_left), std::copy(b.right.begin(), b.right.end(), out_right);
  return 0;
}

extern "C" int pgmm_tree_postorder(int n, const int32_t *left, const int32_t *right, int32_t *order) {
  if (n < 1) return -1;
  const std::vector<int32_t> o = gt::postorder(tree_of(n, left, right));
  std::copy(o.begin(), o.end(), order);
  return (int)o.size();
}

extern "C" void pgmm_free(void *p) { free(p); }

// ---------------- Part 6: FASTA input ----------------
static int fasta_out(bool ok, const std::vector<fasta::Record> &rs, const std::string &msg, pgmm_fasta_record_t **recs, int64_t *n_recs,
                     char *err, int err_cap) {
  *recs = nullptr, *n_recs = 0;
  if (!ok) {
    if (err && err_cap > 0) snprintf(err, (size_t)err_cap, "%s", msg.c_str());
    return -1;
  }
  pgmm_fasta_record_t *out = (pgmm_fasta_record_t *)calloc(std::max<size_t>(1, rs.size()), sizeof(pgmm_fasta_record_t));
  for (size_t i = 0; i < rs.size(); ++i) {
    out[i].name = dup_array(rs[i].name.c_str(), rs[i].name.size() + 1);
    out[i].desc = rs[i].has_desc ? dup_array(rs[i].desc.c_str(), rs[i].desc.size() + 1) : nullptr;
    out[i].seq = dup_array(rs[i].seq.c_str(), rs[i].seq.size() + 1);
    out[i].len = (int64_t)rs[i].seq.size(), out[i].index = rs[i].index;
  }
  *recs = out, *n_recs = (int64_t)rs.size();
  return 0;
}

extern "C" int pgmm_fasta_read_files(int n_paths, const char *const *paths, const char *alphabet, pgmm_fasta_record_t **recs,
                                     int64_t *n_recs, char *err, int err_cap) {
  std::vector<std::string> ps;
  for (int i = 0; i < n_paths; ++i) ps.emplace_back(paths[i]);
  std::vector<fasta::Record> rs;
  std::string msg;
  const bool ok = fasta::read_files(ps, alphabet, rs, msg);
  return fasta_out(ok, rs, msg, recs, n_recs, err, err_cap);
}

extern "C" int pgmm_fasta_read_buffer(const char *data, int64_t n, const char *alphabet, pgmm_fasta_record_t **recs, int64_t *n_recs,
                                      char *err, int err_cap) {
  std::vector<fasta::Record> rs;
  std::string msg;
  const bool ok = fasta::read_buffer(data, (size_t)std::max<int64_t>(0, n), alphabet, rs, msg);
  return fasta_out(ok, rs, msg, recs, n_recs, err, err_cap);
}

extern "C" void pgmm_fasta_free(pgmm_fasta_record_t *recs, int64_t n_recs) {
  if (!recs) return;
  for (int64_t i = 0; i < n_recs; ++i) free(recs[i].name), free(recs[i].desc), free(recs[i].seq);
  free(recs);
}
