// CPU-only test seam for the HOST logic of pangraph_b200 (anchor order, chaining, hit bookkeeping, DP scheduling and
// CIGAR stitching in mapper.cpp / chain.cpp).  It implements pgmm::Backend by forwarding the DEVICE stages to the
// reference's own C (oracle/_ref/libmm2ref.so, opened with dlopen): mm_sketch + mm_seed_mz_flt + mm_collect_matches
// for seeding and ksw_extd2_sse for the DP.  This file is compiled ONLY into tests/_build/libpgmm_hostlogic.so by
// tests/build_hostlogic.py; the product library has no such backend and no CPU path.
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../pangraph_b200/csrc/mapper.h"

using namespace pgmm;

namespace {

struct ref_mm128_t { uint64_t x, y; };
struct ref_mm128_v { size_t n, m; ref_mm128_t *a; };
struct ref_seed_t {  // mmpriv.h:41-47
  uint32_t n, q_pos;
  uint32_t q_span : 31, flt : 1;
  uint32_t seg_id : 31, is_tandem : 1;
  const uint64_t *cr;
};
struct ref_ez_t {  // ksw2.h:31-40
  uint32_t max : 31, zdropped : 1;
  int max_q, max_t, mqe, mqe_t, mte, mte_q, score, m_cigar, n_cigar, reach_end;
  uint32_t *cigar;
};

struct RefLib {
  void *h = nullptr;
  mm_idx_t *(*idx_str)(int, int, int, int, int, const char **, const char **);
  void (*idx_destroy)(mm_idx_t *);
  void (*mapopt_update)(mm_mapopt_t *, const mm_idx_t *);
  void (*sketch)(void *, const char *, int, int, int, uint32_t, int, ref_mm128_v *);
  void (*mz_flt)(void *, ref_mm128_v *, int32_t, float);
  ref_seed_t *(*collect)(void *, int *, int, int, int, int, const mm_idx_t *, const ref_mm128_v *, int64_t *, int *, int *, uint64_t **);
  void (*extd2)(void *, int, const uint8_t *, int, const uint8_t *, int8_t, const int8_t *, int8_t, int8_t, int8_t, int8_t, int, int, int, int, ref_ez_t *);
  explicit RefLib(const char *path) {
    h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!h) { fprintf(stderr, "dlopen %s: %s\n", path, dlerror()); abort(); }
#define SYM(field, name) *(void **)(&field) = dlsym(h, name); if (!field) { fprintf(stderr, "missing %s\n", name); abort(); }
    SYM(idx_str, "mm_idx_str") SYM(idx_destroy, "mm_idx_destroy") SYM(mapopt_update, "mm_mapopt_update")
    SYM(sketch, "mm_sketch") SYM(mz_flt, "mm_seed_mz_flt") SYM(collect, "mm_collect_matches") SYM(extd2, "ksw_extd2_sse")
#undef SYM
  }
};

struct RefBackend : Backend {
  RefLib &L;
  mm_idx_t *mi;
  const TargetSet *ts = nullptr;
  const QueryBatch *qb = nullptr;
  RefBackend(RefLib &l, mm_idx_t *m) : L(l), mi(m) {}
  void begin_batch(const TargetSet &t, const QueryBatch &q) override { ts = &t, qb = &q; }
  void seed_batch(const TargetSet &t, const QueryBatch &q, const mm_mapopt_t &opt, std::vector<QuerySeeds> &out) override {
    out.assign(q.n, QuerySeeds());
    for (int i = 0; i < q.n; ++i) {
      const int qlen = q.lens[i];
      if (qlen == 0) continue;
      ref_mm128_v mv = {0, 0, nullptr};
      L.sketch(nullptr, q.seqs[i], qlen, mi->w, mi->k, 0, mi->flag & 1, &mv);
      if (opt.q_occ_frac > 0.0f) L.mz_flt(nullptr, &mv, opt.mid_occ, opt.q_occ_frac);
      int n_m, rep_len, n_mini_pos;
      int64_t n_a;
      uint64_t *mini_pos;
      ref_seed_t *m = L.collect(nullptr, &n_m, qlen, opt.mid_occ, opt.max_max_occ, opt.occ_dist, mi, &mv, &n_a, &rep_len, &n_mini_pos, &mini_pos);
      QuerySeeds &qs = out[i];
      qs.rep_len = rep_len;
      qs.mini_pos.assign(mini_pos, mini_pos + n_mini_pos);
      // anchor expansion with the all-vs-all skips, restated from map.c:78-100 and :168-204 (no sort here)
      const char *qname = q.names[i];
      for (int s = 0; s < n_m; ++s) {
        const ref_seed_t *sd = &m[s];
        for (uint32_t k = 0; k < sd->n; ++k) {
          const uint64_t r = sd->cr[k];
          const int32_t rpos = (uint32_t)r >> 1;
          bool is_self = false, skip = false;
          if (qname && (opt.flag & (MM_F_NO_DIAG | MM_F_NO_DUAL))) {
            const int cmp = strcmp(qname, mi->seq[r >> 32].name);
            if ((opt.flag & MM_F_NO_DIAG) && cmp == 0 && (int)mi->seq[r >> 32].len == qlen) {
              if ((uint32_t)r >> 1 == (sd->q_pos >> 1)) skip = true;
              if ((r & 1) == (sd->q_pos & 1)) is_self = true;
            }
            if ((opt.flag & MM_F_NO_DUAL) && cmp > 0) skip = true;
          }
          if (skip) continue;
          U128 p;
          if ((r & 1) == (sd->q_pos & 1)) {
            p.x = (r & 0xffffffff00000000ULL) | (uint32_t)rpos;
            p.y = (uint64_t)sd->q_span << 32 | sd->q_pos >> 1;
          } else {
            p.x = 1ULL << 63 | (r & 0xffffffff00000000ULL) | (uint32_t)rpos;
            p.y = (uint64_t)sd->q_span << 32 | (uint32_t)(qlen - ((sd->q_pos >> 1) + 1 - sd->q_span) - 1);
          }
          p.y |= (uint64_t)sd->seg_id << 48;
          if (sd->is_tandem) p.y |= SEED_TANDEM;
          if (is_self) p.y |= SEED_SELF;
          qs.a.push_back(p);
        }
      }
      free(m), free(mini_pos), free(mv.a);
    }
  }
  // no device here: every segment is handed back to the product's host arbiter (chain_fill_host)
  std::vector<std::vector<int32_t>> fpv;
  void chain_fill(const ChainParams &, std::vector<ChainFillJob> &jobs) override {
    fpv.assign(jobs.size(), {});
    for (size_t q = 0; q < jobs.size(); ++q) {
      ChainFillJob &j = jobs[q];
      fpv[q].assign((size_t)j.n * 3 + 1, 0);
      j.f = fpv[q].data(), j.p = j.f + j.n, j.v = j.p + j.n;
      j.redo.assign(j.segs.size(), 1);
    }
  }
  void run_dp(std::vector<KswJob> &jobs, const KswScoring &sc, KswBatchResult &res) override {
    const size_t n = jobs.size();
    res.out.assign(n, KswOut{});
    res.cig_start.assign(n + 1, 0);
    res.cigar.clear();
    int8_t mat[25];
    for (int i = 0; i < 4; ++i) {
      for (int j = 0; j < 4; ++j) mat[i * 5 + j] = i == j ? sc.sc_mch : sc.sc_mis;
      mat[i * 5 + 4] = (int8_t)-abs(sc.sc_ambi);
    }
    for (int j = 0; j < 5; ++j) mat[20 + j] = (int8_t)-abs(sc.sc_ambi);
    for (size_t i = 0; i < n; ++i) {
      const KswJob &j = jobs[i];
      std::vector<uint8_t> qv(qb->codes.begin() + j.q_off, qb->codes.begin() + j.q_off + j.qlen);
      std::vector<uint8_t> tv(ts->codes.begin() + j.t_off, ts->codes.begin() + j.t_off + j.tlen);
      if (j.flag & KSW_JOB_REVSEQ) {
        for (int a = 0, b = j.qlen - 1; a < b; ++a, --b) std::swap(qv[a], qv[b]);
        for (int a = 0, b = j.tlen - 1; a < b; ++a, --b) std::swap(tv[a], tv[b]);
      }
      ref_ez_t ez;
      memset(&ez, 0, sizeof(ez));
      L.extd2(nullptr, j.qlen, qv.data(), j.tlen, tv.data(), 5, mat, sc.q, sc.e, sc.q2, sc.e2, j.w, j.zdrop, j.end_bonus, j.flag & 0xffff, &ez);
      KswOut &o = res.out[i];
      o.max = ez.max, o.zdropped = ez.zdropped, o.max_q = ez.max_q, o.max_t = ez.max_t, o.mqe = ez.mqe, o.mqe_t = ez.mqe_t;
      o.mte = ez.mte, o.mte_q = ez.mte_q, o.score = ez.score, o.reach_end = ez.reach_end, o.n_cigar = ez.n_cigar;
      o.zd_max = -1;  // no device-side mm_test_zdrop scan here: the host does it
      res.cig_start[i] = res.cigar.size();
      res.cigar.insert(res.cigar.end(), ez.cigar, ez.cigar + ez.n_cigar);
      free(ez.cigar);
      res.cells += (uint64_t)j.qlen * j.tlen;
    }
    res.cig_start[n] = res.cigar.size();
  }
};

}  // namespace

// Maps all sequences against all (pangraph's call pattern) with the HOST logic under test; the index and the device
// stages come from the reference library at `ref_so`.
extern "C" __attribute__((visibility("default"))) int pgmm_hostlogic_map_all(
    const char *ref_so, int n, const char **seqs, const char **names, const mm_idxopt_t *io, mm_mapopt_t *mo, int n_threads,
    int *n_regs, mm_reg1_t **regs) {
  static RefLib *L = nullptr;
  if (!L) L = new RefLib(ref_so);
  mm_idx_t *mi = L->idx_str(io->w, io->k, io->flag & 1, io->bucket_bits, n, seqs, names);
  if (!mi) return -1;
  L->mapopt_update(mo, mi);
  TargetSet ts;
  ts.k = mi->k, ts.w = mi->w;
  uint64_t off = 0;
  for (int i = 0; i < n; ++i) {
    ts.names.push_back(names[i]);
    ts.lens.push_back((uint32_t)strlen(seqs[i]));
    ts.offs.push_back(off);
    off += ts.lens.back();
  }
  ts.codes.resize(off + 16);
  for (int i = 0; i < n; ++i)
    for (uint32_t j = 0; j < ts.lens[i]; ++j) ts.codes[ts.offs[i] + j] = kNt4[(uint8_t)seqs[i][j]];
  QueryBatch qb;
  qb.n = n;
  for (int i = 0; i < n; ++i) qb.seqs.push_back(seqs[i]), qb.names.push_back(names[i]), qb.lens.push_back((int)ts.lens[i]);
  RefBackend be(*L, mi);
  map_batch(be, ts, qb, *mo, n_regs, regs, n_threads);
  L->idx_destroy(mi);
  return 0;
}

// ---- host building blocks exposed one at a time (tests/test_host_stages.py) ----
extern "C" __attribute__((visibility("default"))) void pgmm_test_flag_sort_128x(uint64_t *xy, size_t n) {
  flag_sort_128x((U128 *)xy, (U128 *)xy + n);
}
extern "C" __attribute__((visibility("default"))) void pgmm_test_flag_sort_64(uint64_t *v, size_t n) { flag_sort_64(v, v + n); }

// chains anchors (already sorted); returns n_u, writes u[] and the compacted anchors back into xy (2 words each)
extern "C" __attribute__((visibility("default"))) int64_t pgmm_test_chain_rmq(uint64_t *xy, int64_t n, int max_dist, int max_dist_inner,
                                                                              int bw, int max_skip, int cap, int min_cnt, int min_sc,
                                                                              float pen_gap, float pen_skip, uint64_t *u, int64_t *n_a_out) {
  std::vector<U128> a((U128 *)xy, (U128 *)xy + n);
  std::vector<uint64_t> uu;
  ChainParams cp{max_dist, max_dist_inner, bw, max_skip, cap, min_cnt, min_sc, pen_gap, pen_skip};
  chain_rmq(cp, a, uu);
  memcpy(xy, a.data(), a.size() * 16);
  memcpy(u, uu.data(), uu.size() * 8);
  *n_a_out = (int64_t)a.size();
  return (int64_t)uu.size();
}

extern "C" __attribute__((visibility("default"))) int pgmm_test_ll_score(int qlen, const uint8_t *q, int tlen, const uint8_t *t,
                                                                         const int8_t *mat, int gapo, int gape, int *qe, int *te) {
  return ll_local_score(qlen, q, tlen, t, mat, gapo, gape, qe, te);
}

// encode_queries in its two modes: from ASCII, and from the resident target codes (pgmm_map_self).  Returns 0 when they agree.
extern "C" __attribute__((visibility("default"))) int pgmm_test_encode_modes(int n, const char *const *seqs, const int *lens, int n_threads) {
  TargetSet ts;
  uint64_t sum = 0;
  for (int i = 0; i < n; ++i) {
    ts.lens.push_back((uint32_t)lens[i]), ts.offs.push_back(sum), ts.names.push_back("");
    sum += (uint64_t)lens[i];
  }
  ts.codes.resize(sum + 64);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < lens[i]; ++j) ts.codes[ts.offs[i] + j] = kNt4[(uint8_t)seqs[i][j]];
  QueryBatch a, b;
  a.n = b.n = n;
  a.seqs.assign(seqs, seqs + n), b.seqs.assign(n, nullptr);
  a.lens.assign(lens, lens + n), b.lens.assign(lens, lens + n);
  a.names.assign(n, nullptr), b.names.assign(n, nullptr);
  b.from_targets = true;
  encode_queries(a, ts, n_threads);
  encode_queries(b, ts, n_threads);
  if (a.base != b.base) return 1;
  // from the resident codes only the reverse complements are built (the forward strand is read in place from ts.codes)
  for (int i = 0; i < n; ++i) {
    const uint64_t L = (uint64_t)lens[i];
    if (memcmp(a.codes.data() + a.base[i], ts.codes.data() + ts.offs[i], L) != 0) return 3;                    // ASCII path, forward
    if (memcmp(a.codes.data() + a.base[i] + L, b.codes.data() + b.base[i] + L, L) != 0) return 2;             // both paths, reverse
  }
  return 0;
}

// the product's host arbiter on every segment: f, p, v (n int32 each) of sorted anchors
extern "C" __attribute__((visibility("default"))) void pgmm_test_chain_fill_host(const uint64_t *xy, int64_t n, int max_dist, int max_dist_inner,
                                                                                 int bw, int max_skip, int cap, float pen_gap, float pen_skip,
                                                                                 int32_t *f, int32_t *p, int32_t *v) {
  ChainParams cp{max_dist, max_dist_inner, bw, max_skip, cap, 0, 0, pen_gap, pen_skip};
  std::vector<ChainSeg> segs;
  chain_find_segments(cp, (const U128 *)xy, n, segs);
  std::vector<int32_t> t((size_t)n + 1, 0);
  for (const ChainSeg &sg : segs) chain_fill_host(cp, (const U128 *)xy, n, sg.start, sg.end, f, p, v, t.data());
}
