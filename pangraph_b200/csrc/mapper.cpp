// Host side of the mapping pipeline (see mapper.h).  Reference behaviour restated here, with citations into
// packages/minimap2-sys/minimap2/ (C/): map.c:227-374 (per-query driver), hit.c (chains -> hits, filters, order,
// mapq), esterr.c (divergence estimate), align.c (DP window selection, stitching, z-drop splitting, inversion rescue,
// CIGAR clean-up), ksw2_ll_sse.c (striped local score used by the inversion tests).
//
// What is different from the reference is the control structure: mm_align1 runs one DP call after another; here every
// region of every query first *plans* all of its DP windows (they depend on the anchors only), the whole batch of
// windows goes to the GPU as one wave, and regions are *finished* from the results.  Second-pass fills, split-off
// regions and inversion rescues form further waves until nothing is pending.
#include "mapper.h"

#include <algorithm>
#include <atomic>
#include <cassert>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <malloc.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <functional>
#include <thread>
#include <memory_resource>
#include <unordered_map>
#include <ctime>

namespace pgmm {

// Rounds run concurrently on many host threads; keep large vectors on the heap (no mmap/munmap per wave, whose page
// faults serialise on the process's memory map) and never hand memory back to the kernel between waves.
static const bool g_malloc_tuned = [] {
  mallopt(M_MMAP_THRESHOLD, 1 << 30);
  mallopt(M_TRIM_THRESHOLD, INT32_MAX);
  mallopt(M_TOP_PAD, 64 << 20);
  return true;
}();

const uint8_t kNt4[256] = {
#define R4 4, 4, 4, 4
#define R16 R4, R4, R4, R4
    0, 1, 2, 3, R4, R4, R4, R16, R16, R16,                                              // 0..63
    4, 0, 4, 1, 4, 4, 4, 2, R4, R4, 4, 4, 4, 4, 3, 3, 4, 4, R4, R4,                      // 64..95  (A C G T U)
    4, 0, 4, 1, 4, 4, 4, 2, R4, R4, 4, 4, 4, 4, 3, 3, 4, 4, R4, R4,                      // 96..127
    R16, R16, R16, R16, R16, R16, R16, R16
#undef R16
#undef R4
};

static std::atomic<uint64_t> g_cpu_ns[kCpuPhases];
void cpu_phase_add(int phase, uint64_t ns) { g_cpu_ns[phase].fetch_add(ns, std::memory_order_relaxed); }
void cpu_phase_read(double *ms_out, bool reset) {
  for (int i = 0; i < kCpuPhases; ++i) ms_out[i] = (double)(reset ? g_cpu_ns[i].exchange(0) : g_cpu_ns[i].load()) * 1e-6;
}
uint64_t CpuScope::now() {
  timespec t;
  clock_gettime(CLOCK_THREAD_CPUTIME_ID, &t);
  return (uint64_t)t.tv_sec * 1000000000ull + (uint64_t)t.tv_nsec;
}

static void parallel_chunks(uint64_t total, int n_threads, const std::function<void(uint64_t, uint64_t)> &fn) {
  const uint64_t kChunk = 1 << 20;
  const uint64_t n_chunks = (total + kChunk - 1) / kChunk;
  if (n_threads <= 1 || n_chunks <= 1) {
    fn(0, total);
    return;
  }
  std::atomic<uint64_t> next(0);
  std::vector<std::thread> th;
  const int nt = (int)std::min<uint64_t>(n_chunks, (uint64_t)n_threads);
  for (int t = 0; t < nt; ++t)
    th.emplace_back([&]() {
      for (;;) {
        const uint64_t c = next.fetch_add(1);
        if (c >= n_chunks) break;
        fn(c * kChunk, std::min(total, (c + 1) * kChunk));
      }
    });
  for (auto &t : th) t.join();
}

void encode_queries(QueryBatch &qb, const TargetSet &ts, int n_threads) {
  qb.base.resize(qb.n);
  uint64_t tot = 0;
  std::vector<uint64_t> vstart(qb.n + 1, 0);
  for (int i = 0; i < qb.n; ++i) qb.base[i] = tot, tot += 2ull * qb.lens[i], vstart[i + 1] = vstart[i] + (uint64_t)qb.lens[i];
  qb.codes.resize(tot + 16);
  memset(qb.codes.data() + tot, 0, 16);  // (the buffer itself is not zero-filled: every base is written below)
  // forward codes and their reverse complement per query (align.c:969-975), chunked over all bases of the batch
  parallel_chunks(vstart[qb.n], n_threads, [&](uint64_t lo, uint64_t hi) {
    CpuScope cpu_scope(0);
    int i = (int)(std::upper_bound(vstart.begin(), vstart.end(), lo) - vstart.begin()) - 1;
    for (uint64_t p = lo; p < hi;) {
      while (vstart[i + 1] <= p) ++i;
      const uint64_t L = (uint64_t)qb.lens[i], j0 = p - vstart[i], j1 = std::min(L, hi - vstart[i]);
      uint8_t *f = qb.codes.data() + qb.base[i], *r = f + L;
      if (qb.from_targets) {  // codes are there already: only the reverse complement is built, eight bases at a time
        const uint8_t *s = ts.codes.data() + ts.offs[i];  // (the forward strand is read in place: map_batch points q0[0] here)
        uint64_t j = j0;
        for (; j + 8 <= j1; j += 8) {
          uint64_t x;
          memcpy(&x, s + j, 8);
          x = __builtin_bswap64(x);                                // base j lands in the last byte
          const uint64_t amb = (x & 0x0404040404040404ull) >> 2;   // 1 where the code is 4
          x = (x ^ 0x0303030303030303ull) ^ (amb * 3);             // 0..3 -> 3 - c, 4 stays 4
          memcpy(r + (L - 8 - j), &x, 8);
        }
        for (; j < j1; ++j) {
          const uint8_t c = s[j];
          r[L - 1 - j] = c < 4 ? 3 - c : 4;
        }
      } else {
        const uint8_t *s = (const uint8_t *)qb.seqs[i];
        for (uint64_t j = j0; j < j1; ++j) {
          const uint8_t c = kNt4[s[j]];
          f[j] = c, r[L - 1 - j] = c < 4 ? 3 - c : 4;
        }
      }
      p = vstart[i] + j1;
    }
  });
}

namespace {

constexpr int32_t PARENT_UNSET = -1, PARENT_TMP_PRI = -2;

inline uint32_t roundup_pow2(uint32_t x) {
  --x;
  x |= x >> 1, x |= x >> 2, x |= x >> 4, x |= x >> 8, x |= x >> 16;
  return ++x;
}
inline float fast_log2f(float x) {  // mmpriv.h:118-126
  union {
    float f;
    uint32_t i;
  } z = {x};
  float log_2 = (float)(int)(((z.i >> 23) & 255) - 128);
  z.i &= ~(255u << 23);
  z.i += 127u << 23;
  log_2 += (-0.34484843f * z.f + 2.02466578f) * z.f - 0.67487759f;
  return log_2;
}
inline uint32_t wang32(uint32_t key) {  // khash.h:400-409
  key += ~(key << 15);
  key ^= (key >> 10);
  key += (key << 3);
  key ^= (key >> 6);
  key += ~(key << 11);
  key ^= (key >> 16);
  return key;
}
inline uint32_t x31_hash(const char *s) {  // khash.h:383-388
  uint32_t h = (uint32_t)*s;
  if (h)
    for (++s; *s; ++s) h = (h << 5) - h + (uint32_t)*s;
  return h;
}
inline uint64_t mix64(uint64_t key) {  // hit.c:40-50
  key = (~key + (key << 21));
  key = key ^ key >> 24;
  key = ((key + (key << 3)) + (key << 8));
  key = key ^ key >> 14;
  key = ((key + (key << 2)) + (key << 4));
  key = key ^ key >> 28;
  key = (key + (key << 31));
  return key;
}

using Ez = KswOut;
inline Ez ez_reset() {
  Ez e{};
  e.max_q = e.max_t = e.mqe_t = e.mte_q = -1;
  e.score = e.mqe = e.mte = KSW_NEG_INF;
  return e;
}

// A DP problem is a pure function of its windows and parameters, and the same window comes up again when a hit is split
// at a z-drop and its remainder is planned anew (the fills behind the split and the right extension are the ones the
// original hit already asked for, speculatively).  Every query keeps the results it has seen, keyed by the job itself.
struct DpKey {
  uint64_t q_off, t_off;
  int32_t qlen, tlen, w, zdrop, end_bonus, flag;
  bool operator==(const DpKey &o) const {
    return q_off == o.q_off && t_off == o.t_off && qlen == o.qlen && tlen == o.tlen && w == o.w && zdrop == o.zdrop &&
           end_bonus == o.end_bonus && flag == o.flag;
  }
};
struct DpKeyHash {
  size_t operator()(const DpKey &k) const {
    uint64_t h = k.q_off * 0x9E3779B97F4A7C15ull;
    h ^= (k.t_off + 0x7F4A7C15ull) * 0xC2B2AE3D27D4EB4Full;
    h ^= ((uint64_t)(uint32_t)k.qlen << 32 | (uint32_t)k.tlen) * 0x165667B19E3779F9ull;
    h ^= ((uint64_t)(uint32_t)k.w << 32 | (uint32_t)k.flag) + ((uint64_t)(uint32_t)k.zdrop << 17) + (uint32_t)k.end_bonus;
    h ^= h >> 29;
    return (size_t)(h * 0xBF58476D1CE4E5B9ull);
  }
};
struct DpCached {
  int8_t zcode = -1;  // mm_test_zdrop's verdict on this result when it is a first-pass fill (-1 = not computed yet)
  int job = -1;  // >= 0: submitted in the current wave under this index, result not in yet
  KswOut ez;
  const uint32_t *cigar = nullptr;  // inside a wave result the query keeps alive (QCtx::wave_results)
};

struct DpCall {   // one DP window and, once its wave has run, its result
  int job = -1;   // index in the query's job list of the current wave; -1 = not submitted
  int wave = -1;  // the wave `job` belongs to
  struct DpCached *entry = nullptr;  // the query's result-cache entry of this problem (node addresses are stable), if it has one
  Ez ez = ez_reset();
  // the CIGAR stays in the wave's result buffer instead of being copied per problem; the buffer lives as long as the query
  // (QCtx::wave_results holds one reference per wave -- a reference per problem was 70 000 atomic increments per round on a
  // counter shared by every round of the merged wave)
  const uint32_t *cigar = nullptr;
};

struct Fill {
  int i;  // anchor index relative to as1
  int32_t rs, qs, re, qe;
  int bw1;
  DpCall pass1, pass2;
  int code = 0;        // mm_test_zdrop verdict on pass 1
  bool spec2 = false;  // the exact second pass was queued together with the first one
};

struct Region {
  mm_reg1_t r;
  bool has_p = false;
  uint32_t capacity = 0, n_ambi = 0;
  int32_t dp_score = 0, dp_max = 0, dp_max2 = 0;
  std::vector<uint32_t> cigar;
  enum State { NEW, WAIT1, WAIT2, WAIT_INV, DONE } state = NEW;
  bool inv_checked = false;
  // plan of mm_align1
  int32_t rid = 0, rev = 0, as1 = 0, cnt1 = 0, rs = 0, qs = 0, re = 0, qe = 0, rs0 = 0, qs0 = 0, re0 = 0, qe0 = 0;
  bool has_left = false, has_right = false;
  DpCall left, right;
  std::vector<Fill> fills;
  // inversion rescue in flight (mm_align1_inv)
  DpCall inv;
  int32_t inv_q_off = 0, inv_t_off = 0, inv_ql = 0, inv_tl = 0;
  Region() { memset(&r, 0, sizeof(r)); }
};

struct QCtx {
  int qi = 0, qlen = 0;
  const char *qname = nullptr;
  const uint8_t *q0[2] = {nullptr, nullptr};
  uint64_t qbase = 0;
  uint32_t hash = 0;
  std::vector<U128> a;
  int n_a = 0, rep_len = 0;
  std::vector<uint64_t> mini_pos;
  std::vector<std::unique_ptr<Region>> regs;
  std::vector<KswJob> jobs;  // this query's share of the next wave
  size_t job_base = 0;
  bool pending = false;
  // every DP result of this query so far; its nodes come from a bump allocator that is released with the query (23 000
  // windows per 5-Mbp query: one malloc and one free each otherwise)
  std::pmr::monotonic_buffer_resource dp_pool{1 << 20};
  std::pmr::unordered_map<DpKey, DpCached, DpKeyHash> dp_cache{&dp_pool};
  std::vector<DpCached *> wave_entries;                      // cache entry of jobs[i] of the wave being assembled / in flight
  std::vector<std::shared_ptr<const KswBatchResult>> wave_results;  // every wave result this query's calls point into
  uint64_t dp_reused = 0;
  int wave_id = 0, done_wave = -1;  // the wave being assembled; the last one whose results are in
};

struct Mapper {
  const TargetSet &ts;
  const QueryBatch &qb;
  const mm_mapopt_t &opt;
  int8_t mat[25];
  // fills at least this long get their exact second pass queued with the first one (0 = never; PGMM_SPEC_FILL_LEN)
  int spec_fill_len = 400;
  bool dp_reuse = true;  // PGMM_DP_REUSE=0: every planned window is computed again
  int spec_depth = 16;   // generations of split-off remainders planned ahead (PGMM_SPEC_DEPTH, 0 = off)
  Mapper(const TargetSet &t, const QueryBatch &q, const mm_mapopt_t &o) : ts(t), qb(q), opt(o) {
    if (const char *e = getenv("PGMM_SPEC_FILL_LEN")) spec_fill_len = atoi(e);
    if (const char *e = getenv("PGMM_DP_REUSE")) dp_reuse = atoi(e) != 0;
    if (const char *e = getenv("PGMM_SPEC_DEPTH")) spec_depth = atoi(e);
    // ksw_gen_simple_mat(5, mat, a, b, sc_ambi), align.c:9-22
    const int a = opt.a < 0 ? -opt.a : opt.a, b = opt.b > 0 ? -opt.b : opt.b, amb = opt.sc_ambi > 0 ? -opt.sc_ambi : opt.sc_ambi;
    for (int i = 0; i < 4; ++i) {
      for (int j = 0; j < 4; ++j) mat[i * 5 + j] = (int8_t)(i == j ? a : b);
      mat[i * 5 + 4] = (int8_t)amb;
    }
    for (int j = 0; j < 5; ++j) mat[20 + j] = (int8_t)amb;
  }
  const uint8_t *tseq(int rid, int32_t pos) const { return ts.codes.data() + ts.offs[rid] + pos; }

  // ---------------- chains -> hits (hit.c:8-88) ----------------
  static void set_coor(mm_reg1_t &r, int32_t qlen, const U128 *a) {
    const int32_t k = r.as, q_span = (int32_t)(a[k].y >> 32 & 0xff);
    r.rev = a[k].x >> 63;
    r.rid = (int32_t)(a[k].x << 1 >> 33);
    r.rs = (int32_t)a[k].x + 1 > q_span ? (int32_t)a[k].x + 1 - q_span : 0;
    r.re = (int32_t)a[k + r.cnt - 1].x + 1;
    if (!r.rev) {
      r.qs = (int32_t)a[k].y + 1 - q_span;
      r.qe = (int32_t)a[k + r.cnt - 1].y + 1;
    } else {
      r.qs = qlen - ((int32_t)a[k + r.cnt - 1].y + 1);
      r.qe = qlen - ((int32_t)a[k].y + 1 - q_span);
    }
    // fuzzy lengths, hit.c:8-21
    r.mlen = r.blen = 0;
    if (r.cnt <= 0) return;
    r.mlen = r.blen = (int32_t)(a[r.as].y >> 32 & 0xff);
    for (int i = r.as + 1; i < r.as + r.cnt; ++i) {
      const int span = (int)(a[i].y >> 32 & 0xff);
      const int tl = (int32_t)a[i].x - (int32_t)a[i - 1].x, ql = (int32_t)a[i].y - (int32_t)a[i - 1].y;
      r.blen += tl > ql ? tl : ql;
      r.mlen += tl > span && ql > span ? span : tl < ql ? tl : ql;
    }
  }

  void gen_regs(QCtx &q, const std::vector<uint64_t> &u) const {
    const int n_u = (int)u.size();
    if (n_u == 0) return;
    std::vector<U128> z(n_u);
    for (int i = 0, k = 0; i < n_u; ++i) {
      const uint32_t h = (uint32_t)mix64((mix64(q.a[k].x) + mix64(q.a[k].y)) ^ q.hash);
      z[i].x = u[i] ^ h;
      z[i].y = (uint64_t)k << 32 | (uint32_t)(int32_t)u[i];
      k += (int32_t)u[i];
    }
    flag_sort_128x(z.data(), z.data() + n_u);
    std::reverse(z.begin(), z.end());  // larger score first (hit.c:70-71 swaps ends pairwise: the same permutation)
    for (int i = 0; i < n_u; ++i) {
      auto R = std::make_unique<Region>();
      mm_reg1_t &ri = R->r;
      ri.id = i, ri.parent = PARENT_UNSET;
      ri.score = ri.score0 = (int32_t)(z[i].x >> 32);
      ri.hash = (uint32_t)z[i].x;
      ri.cnt = (int32_t)z[i].y;
      ri.as = (int32_t)(z[i].y >> 32);
      ri.div = -1.0f;
      set_coor(ri, q.qlen, q.a.data());
      q.regs.push_back(std::move(R));
    }
  }

  // ---------------- divergence estimate (esterr.c:7-64) ----------------
  static int32_t fwd_qpos(int32_t qlen, const U128 &a) {
    int32_t x = (int32_t)a.y;
    const int32_t q_span = (int32_t)(a.y >> 32 & 0xff);
    if (a.x >> 63) x = qlen - 1 - (x + 1 - q_span);
    return x;
  }
  void est_err(QCtx &q) const {
    const int32_t n = (int32_t)q.mini_pos.size();
    if (n == 0) return;
    uint64_t sum_k = 0;
    for (int i = 0; i < n; ++i) sum_k += q.mini_pos[i] >> 32 & 0xff;
    const float avg_k = (float)sum_k / n;
    for (auto &R : q.regs) {
      mm_reg1_t &r = R->r;
      r.div = -1.0f;
      if (r.cnt == 0) continue;
      const U128 *a = q.a.data();
      int32_t st = -1;
      {
        const int32_t x = fwd_qpos(q.qlen, r.rev ? a[r.as + r.cnt - 1] : a[r.as]);
        int32_t L = 0, Rr = n - 1;
        while (L <= Rr) {
          const int32_t m = (int32_t)(((uint64_t)L + Rr) >> 1), y = (int32_t)q.mini_pos[m];
          if (y < x) L = m + 1;
          else if (y > x) Rr = m - 1;
          else {
            st = m;
            break;
          }
        }
      }
      if (st < 0) continue;
      int32_t en = st, k = 1, n_match = 1;
      const int32_t l_ref = (int32_t)ts.lens[r.rid];
      for (int32_t j = st + 1; j < n && k < r.cnt; ++j) {
        const int32_t x = fwd_qpos(q.qlen, r.rev ? a[r.as + r.cnt - 1 - k] : a[r.as + k]);
        if (x == (int32_t)q.mini_pos[j]) ++k, en = j, ++n_match;
      }
      int32_t n_tot = en - st + 1;
      if (r.qs > avg_k && r.rs > avg_k) ++n_tot;
      if (q.qlen - r.qs > avg_k && l_ref - r.re > avg_k) ++n_tot;
      r.div = n_match >= n_tot ? 0.0f : (float)(1.0 - pow((double)n_match / n_tot, 1.0 / avg_k));
    }
  }

  // drops anchors that no hit references (hit.c:311-329)
  static int squeeze_a(QCtx &q) {
    const int n = (int)q.regs.size();
    std::vector<uint64_t> aux(n);
    for (int i = 0; i < n; ++i) aux[i] = (uint64_t)q.regs[i]->r.as << 32 | (uint32_t)i;
    flag_sort_64(aux.data(), aux.data() + n);
    int as = 0;
    for (int i = 0; i < n; ++i) {
      mm_reg1_t &r = q.regs[(int32_t)aux[i]]->r;
      if (r.as != as) {
        memmove(&q.a[as], &q.a[r.as], (size_t)r.cnt * 16);
        r.as = as;
      }
      as += r.cnt;
    }
    return as;
  }

  // ---------------- CIGAR bookkeeping (align.c:291-314) ----------------
  static void append_cigar(Region &R, const uint32_t *cig, int n) {
    if (n == 0) return;
    if (!R.has_p) {
      R.capacity = roundup_pow2((uint32_t)n + 6);
      R.has_p = true;
    } else if (R.cigar.size() + n + 6 > R.capacity) {
      R.capacity = roundup_pow2((uint32_t)(R.cigar.size() + n + 6));
    }
    if (!R.cigar.empty() && (R.cigar.back() & 0xf) == (cig[0] & 0xf)) {
      R.cigar.back() += cig[0] >> 4 << 4;
      R.cigar.insert(R.cigar.end(), cig + 1, cig + n);
    } else R.cigar.insert(R.cigar.end(), cig, cig + n);
  }

  // indel left-alignment, merging of xIyDzI runs, removal of a leading gap (align.c:91-167)
  static void fix_cigar(Region &R, const uint8_t *qseq, const uint8_t *tseq, int *qshift, int *tshift) {
    mm_reg1_t &r = R.r;
    std::vector<uint32_t> &c = R.cigar;
    int32_t toff = 0, qoff = 0;
    bool to_shrink = false;
    *qshift = *tshift = 0;
    uint32_t n = (uint32_t)c.size();
    if (n <= 1) return;
    for (uint32_t k = 0; k < n; ++k) {
      const uint32_t op = c[k] & 0xf, len = c[k] >> 4;
      if (len == 0) to_shrink = true;
      if (op == MM_CIGAR_MATCH) toff += len, qoff += len;
      else if (op == MM_CIGAR_INS || op == MM_CIGAR_DEL) {
        if (k > 0 && k < n - 1 && (c[k - 1] & 0xf) == 0 && (c[k + 1] & 0xf) == 0) {
          int l;
          const int prev_len = (int)(c[k - 1] >> 4);
          if (op == MM_CIGAR_INS) {
            for (l = 0; l < prev_len; ++l)
              if (qseq[qoff - 1 - l] != qseq[qoff + len - 1 - l]) break;
          } else {
            for (l = 0; l < prev_len; ++l)
              if (tseq[toff - 1 - l] != tseq[toff + len - 1 - l]) break;
          }
          if (l > 0) c[k - 1] -= (uint32_t)l << 4, c[k + 1] += (uint32_t)l << 4, qoff -= l, toff -= l;
          if (l == prev_len) to_shrink = true;
        }
        if (op == MM_CIGAR_INS) qoff += len;
        else toff += len;
      } else if (op == 3) toff += len;
    }
    assert(qoff == r.qe - r.qs && toff == r.re - r.rs);
    for (uint32_t k = 0; k + 2 < n; ++k) {
      if ((c[k] & 0xf) > 0 && (c[k] & 0xf) + (c[k + 1] & 0xf) == 3) {
        uint32_t l, s[3] = {0, 0, 0};
        for (l = k; l < n; ++l) {
          const uint32_t op = c[l] & 0xf;
          if (op == MM_CIGAR_INS || op == MM_CIGAR_DEL || c[l] >> 4 == 0) s[op] += c[l] >> 4;
          else break;
        }
        if (s[1] > 0 && s[2] > 0 && l - k > 2) {
          c[k] = s[1] << 4 | MM_CIGAR_INS;
          c[k + 1] = s[2] << 4 | MM_CIGAR_DEL;
          for (k += 2; k < l; ++k) c[k] &= 0xf;
          to_shrink = true;
        }
        k = l;
      }
    }
    if (to_shrink) {
      uint32_t l = 0;
      for (uint32_t k = 0; k < n; ++k)
        if (c[k] >> 4 != 0) c[l++] = c[k];
      n = l;
      l = 0;
      for (uint32_t k = 0; k < n; ++k)
        if (k == n - 1 || (c[k] & 0xf) != (c[k + 1] & 0xf)) c[l++] = c[k];
        else c[k + 1] += c[k] >> 4 << 4;
      n = l;
    }
    if ((c[0] & 0xf) == MM_CIGAR_INS || (c[0] & 0xf) == MM_CIGAR_DEL) {
      const int32_t l = (int32_t)(c[0] >> 4);
      if ((c[0] & 0xf) == MM_CIGAR_INS) {
        if (r.rev) r.qe -= l;
        else r.qs += l;
        *qshift = l;
      } else r.rs += l, *tshift = l;
      --n;
      memmove(c.data(), c.data() + 1, (size_t)n * 4);
    }
    c.resize(n);
  }

  // mlen / blen / n_ambi / dp_max from the final CIGAR (align.c:240-289)
  void update_extra(Region &R, const uint8_t *qseq, const uint8_t *tseq) const {
    if (!R.has_p) return;
    mm_reg1_t &r = R.r;
    int qshift, tshift;
    int32_t toff = 0, qoff = 0;
    double s = 0.0, max = 0.0;
    const int q = (int8_t)opt.q, e = (int8_t)opt.e;
    fix_cigar(R, qseq, tseq, &qshift, &tshift);
    qseq += qshift, tseq += tshift;
    r.blen = r.mlen = 0;
    for (uint32_t c : R.cigar) {
      const uint32_t op = c & 0xf, len = c >> 4;
      if (op == MM_CIGAR_MATCH) {
        int n_ambi = 0, n_diff = 0;
        uint32_t l = 0;
        // eight equal unambiguous bases at a time: the score only climbs through them (s >= 0 on entry, every step adds
        // the match score), so the clamp never acts and the running maximum ends at the last one -- the same doubles as
        // base by base (all of them integers plus a few float-precision gap terms, far below 2^53)
        if (mat[0] > 0)
          for (; l + 8 <= len; l += 8) {
            uint64_t wq, wt;
            memcpy(&wq, qseq + qoff + l, 8), memcpy(&wt, tseq + toff + l, 8);
            if (wq != wt || (wq & 0x0404040404040404ull)) {  // a mismatch or an ambiguous base: this block goes base by base
              for (uint32_t j = l; j < l + 8; ++j) {
                const int cq = qseq[qoff + j], ct = tseq[toff + j];
                if (ct > 3 || cq > 3) ++n_ambi;
                else if (ct != cq) ++n_diff;
                s += mat[ct * 5 + cq];
                if (s < 0) s = 0;
                else max = max > s ? max : s;
              }
              continue;
            }
            s += 8.0 * mat[0];
            max = max > s ? max : s;
          }
        for (; l < len; ++l) {
          const int cq = qseq[qoff + l], ct = tseq[toff + l];
          if (ct > 3 || cq > 3) ++n_ambi;
          else if (ct != cq) ++n_diff;
          s += mat[ct * 5 + cq];
          if (s < 0) s = 0;
          else max = max > s ? max : s;
        }
        r.blen += len - n_ambi, r.mlen += len - (n_ambi + n_diff), R.n_ambi += n_ambi;
        toff += len, qoff += len;
      } else if (op == MM_CIGAR_INS || op == MM_CIGAR_DEL) {
        int n_ambi = 0;
        const uint8_t *sq = op == MM_CIGAR_INS ? qseq + qoff : tseq + toff;
        for (uint32_t l = 0; l < len; ++l)
          if (sq[l] > 3) ++n_ambi;
        r.blen += len - n_ambi, R.n_ambi += n_ambi;
        s -= q + (double)e * fast_log2f(1.0f + (float)len);  // log_gap is always on (no MM_F_SR)
        if (s < 0) s = 0;
        if (op == MM_CIGAR_INS) qoff += len;
        else toff += len;
      } else if (op == 3) toff += len;
    }
    R.dp_max = (int32_t)(max + .499);
    assert(qoff == r.qe - r.qs && toff == r.re - r.rs);
  }

  // ---------------- seed clean-up before DP (align.c:372-498) ----------------
  static void long_gaps(const U128 *a, int as1, int cnt1, int min_gap, std::vector<int> &K) {
    K.clear();
    for (int i = 1; i < cnt1; ++i) {
      const int gap = (int)(((int32_t)a[as1 + i].y - a[as1 + i - 1].y) - ((int32_t)a[as1 + i].x - a[as1 + i - 1].x));
      if (gap < -min_gap || gap > min_gap) K.push_back(i);
    }
    if (K.size() <= 1) K.clear();
  }
  static void filter_bad_seeds(U128 *a, int as1, int cnt1, int min_gap, int diff_thres, int max_ext_len, int max_ext_cnt) {
    std::vector<int> K;
    long_gaps(a, as1, cnt1, min_gap, K);
    const int n = (int)K.size();
    if (n == 0) return;
    int max = 0, max_st = -1, max_en = -1;
    for (int k = 0;; ++k) {
      if (k == n || k >= max_en) {
        if (max_en > 0)
          for (int i = K[max_st]; i < K[max_en]; ++i) a[as1 + i].y |= SEED_IGNORE;
        max = 0, max_st = max_en = -1;
        if (k == n) break;
      }
      const int i = K[k];
      int gap = ((int32_t)a[as1 + i].y - (int32_t)a[as1 + i - 1].y) - (int32_t)(a[as1 + i].x - a[as1 + i - 1].x);
      int n_ins = 0, n_del = 0, max_diff = 0, max_diff_l = -1;
      if (gap > 0) n_ins += gap;
      else n_del += -gap;
      const int qs = (int32_t)a[as1 + i - 1].y, rs = (int32_t)a[as1 + i - 1].x;
      for (int l = k + 1; l < n && l <= k + max_ext_cnt; ++l) {
        const int j = K[l];
        if ((int32_t)a[as1 + j].y - qs > max_ext_len || (int32_t)a[as1 + j].x - rs > max_ext_len) break;
        gap = ((int32_t)a[as1 + j].y - (int32_t)a[as1 + j - 1].y) - (int32_t)(a[as1 + j].x - a[as1 + j - 1].x);
        if (gap > 0) n_ins += gap;
        else n_del += -gap;
        const int diff = n_ins + n_del - abs(n_ins - n_del);
        if (max_diff < diff) max_diff = diff, max_diff_l = l;
      }
      if (max_diff > diff_thres && max_diff > max) max = max_diff, max_st = k, max_en = max_diff_l;
    }
  }
  static void filter_bad_seeds_alt(U128 *a, int as1, int cnt1, int min_gap, int max_ext) {
    std::vector<int> K;
    long_gaps(a, as1, cnt1, min_gap, K);
    const int n = (int)K.size();
    for (int k = 0; k < n;) {
      const int i = K[k];
      int l;
      int gap1 = ((int32_t)a[as1 + i].y - (int32_t)a[as1 + i - 1].y) - ((int32_t)a[as1 + i].x - (int32_t)a[as1 + i - 1].x);
      int re1 = (int32_t)a[as1 + i].x, qe1 = (int32_t)a[as1 + i].y;
      gap1 = gap1 > 0 ? gap1 : -gap1;
      for (l = k + 1; l < n; ++l) {
        const int j = K[l];
        if ((int32_t)a[as1 + j].y - qe1 > max_ext || (int32_t)a[as1 + j].x - re1 > max_ext) break;
        int gap2 = ((int32_t)a[as1 + j].y - (int32_t)a[as1 + j - 1].y) - (int32_t)(a[as1 + j].x - a[as1 + j - 1].x);
        const int q_span_pre = (int)(a[as1 + j - 1].y >> 32 & 0xff);
        const int rs2 = (int32_t)a[as1 + j - 1].x + q_span_pre, qs2 = (int32_t)a[as1 + j - 1].y + q_span_pre;
        const int m = rs2 - re1 < qs2 - qe1 ? rs2 - re1 : qs2 - qe1;
        gap2 = gap2 > 0 ? gap2 : -gap2;
        if (m > gap1 + gap2) break;
        re1 = (int32_t)a[as1 + j].x, qe1 = (int32_t)a[as1 + j].y;
        gap1 = gap2;
      }
      if (l > k + 1) {
        const int end = K[l - 1];
        for (int j = K[k]; j < end; ++j) a[as1 + j].y |= SEED_IGNORE;
        a[as1 + end].y |= SEED_LONG_JOIN;
      }
      k = l;
    }
  }
  static void fix_bad_ends(const mm_reg1_t &r, const U128 *a, int bw, int min_match, int32_t *as, int32_t *cnt) {
    *as = r.as, *cnt = r.cnt;
    if (r.cnt < 3) return;
    int32_t m, l;
    m = l = (int32_t)(a[r.as].y >> 32 & 0xff);
    for (int32_t i = r.as + 1; i < r.as + r.cnt - 1; ++i) {
      const int32_t q_span = (int32_t)(a[i].y >> 32 & 0xff);
      if (a[i].y & SEED_LONG_JOIN) break;
      const int32_t lr = (int32_t)a[i].x - (int32_t)a[i - 1].x, lq = (int32_t)a[i].y - (int32_t)a[i - 1].y;
      const int32_t mn = lr < lq ? lr : lq, mx = lr > lq ? lr : lq;
      if (mx - mn > l >> 1) *as = i;
      l += mn;
      m += mn < q_span ? mn : q_span;
      if (l >= bw << 1 || (m >= min_match && m >= bw) || m >= r.mlen >> 1) break;
    }
    *cnt = r.as + r.cnt - *as;
    m = l = (int32_t)(a[r.as + r.cnt - 1].y >> 32 & 0xff);
    for (int32_t i = r.as + r.cnt - 2; i > *as; --i) {
      const int32_t q_span = (int32_t)(a[i + 1].y >> 32 & 0xff);
      if (a[i + 1].y & SEED_LONG_JOIN) break;
      const int32_t lr = (int32_t)a[i + 1].x - (int32_t)a[i].x, lq = (int32_t)a[i + 1].y - (int32_t)a[i].y;
      const int32_t mn = lr < lq ? lr : lq, mx = lr > lq ? lr : lq;
      if (mx - mn > l >> 1) *cnt = i + 1 - *as;
      l += mn;
      m += mn < q_span ? mn : q_span;
      if (l >= bw << 1 || (m >= min_match && m >= bw) || m >= r.mlen >> 1) break;
    }
  }

  // ---------------- DP window submission ----------------
  // registers one DP window; windows above max_sw_mat are answered on the spot like mm_align_pair does (align.c:326-328)
  void submit(QCtx &q, DpCall &c, int strand, int32_t qstart, int32_t ql, int rid, int32_t tstart, int32_t tl, int w,
              int zdrop, int end_bonus, int flag) const {
    c.ez = ez_reset();
    c.cigar = nullptr;
    c.job = -1, c.entry = nullptr;
    if (opt.max_sw_mat > 0 && (int64_t)tl * ql > opt.max_sw_mat) {
      c.ez.zdropped = 1;
      return;
    }
    if (ql <= 0 || tl <= 0) return;
    KswJob j;
    memset(&j, 0, sizeof(j));
    j.q_off = q.qbase + (strand ? (uint64_t)q.qlen : 0) + (uint64_t)qstart;
    j.t_off = ts.offs[rid] + (uint64_t)tstart;
    j.qlen = ql, j.tlen = tl, j.w = w, j.zdrop = zdrop, j.end_bonus = end_bonus, j.flag = flag;
    c.wave = q.wave_id;
    if (dp_reuse) {
      const DpKey key{j.q_off, j.t_off, ql, tl, w, zdrop, end_bonus, flag};
      if (q.dp_cache.bucket_count() < 4096) q.dp_cache.reserve(32768);  // one hash operation per window, no rehash on the way
      const auto ins = q.dp_cache.try_emplace(key);
      DpCached &e = ins.first->second;
      c.entry = &e;
      if (!ins.second) {
        ++q.dp_reused;
        if (e.job >= 0) c.job = e.job, q.pending = true;  // asked for earlier in this very wave: share the slot
        else c.ez = e.ez, c.cigar = e.cigar;
        return;
      }
      e.job = (int)q.jobs.size();
      q.wave_entries.push_back(&e);
    }
    c.job = (int)q.jobs.size();
    q.jobs.push_back(j);
    q.pending = true;
  }
  // the results of the wave that just ran become reusable (before any hit of the query looks at them)
  static void publish_wave(QCtx &q, const std::shared_ptr<const KswBatchResult> &res) {
    q.wave_results.push_back(res);
    for (size_t k = 0; k < q.wave_entries.size(); ++k) {
      DpCached &e = *q.wave_entries[k];
      const size_t g = q.job_base + k;
      e.job = -1, e.ez = res->out[g], e.cigar = res->cigar.data() + res->cig_start[g];
    }
    q.wave_entries.clear();
  }
  // a call is ready when nothing was submitted for it or the wave it went into has run
  static bool ready(const QCtx &q, const DpCall &c) { return c.job < 0 || c.wave == q.done_wave; }
  static bool region_ready(const QCtx &q, const Region &R) {
    if (!ready(q, R.left) || !ready(q, R.right) || !ready(q, R.inv)) return false;
    for (const Fill &f : R.fills)
      if (!ready(q, f.pass1) || !ready(q, f.pass2)) return false;
    return true;
  }
  // mm_test_zdrop on a first-pass result; the verdict is remembered with the result (it is a function of the same inputs)
  int fill_code(QCtx &q, const Region &R, const Fill &f, DpCall &p1) const {
    DpCached *e = p1.entry;
    if (e && e->zcode >= 0) return e->zcode;
    if (p1.ez.zd_max < 0) zdrop_scan(q.q0[R.rev] + f.qs, tseq(R.rid, f.rs), p1.cigar, p1.ez.n_cigar, p1.ez);
    const int code = test_zdrop(q.q0[R.rev] + f.qs, tseq(R.rid, f.rs), p1.ez);
    if (e) e->zcode = (int8_t)code;
    return code;
  }
  // the result of a call that is not pending: its own, or -- for a call planned in this very iteration -- nothing yet
  static bool resolved(const DpCall &c) { return c.job < 0; }

  // Where will this hit be split?  Everything finish_region needs for that -- the z-drop verdict of every fill and the
  // exact pass of the first fill that drops -- may already be known from the hit the region was split off (their
  // windows coincide), long before the region's own new windows (its left extension, usually its first fill) have
  // run.  A fill whose result is still pending is taken not to drop.  Returns true and the split-off hit when the
  // answer is "it will be split"; exact second passes that are missing are queued on the way (speculatively: their
  // results only land in the query's result cache).
  bool predict_split(QCtx &q, Region &R, mm_reg1_t &r2) const {
    const U128 *a = q.a.data();
    static const bool trace = getenv("PGMM_TRACE") != nullptr;
    if (trace) {
      int n_pend = 0, n_code = 0, n_drop = 0;
      for (Fill &f : R.fills) {
        if (!resolved(f.pass1)) { ++n_pend; continue; }
        const int code = fill_code(q, R, f, f.pass1);
        n_code += code != 0, n_drop += f.pass1.ez.zdropped != 0;
        if (code) fprintf(stderr, "[pgmm trace]   fill i=%d len %d x %d code %d spec2 %d pass2 resolved %d zdropped %d\n", f.i, f.qe - f.qs, f.re - f.rs, code, (int)f.spec2, (int)resolved(f.pass2), f.pass2.ez.zdropped);
      }
      fprintf(stderr, "[pgmm trace]   predict: %zu fills, %d pending, %d fail the z-drop test, %d first passes dropped\n", R.fills.size(), n_pend, n_code, n_drop);
    }
    for (Fill &f : R.fills) {
      DpCall &p1 = f.pass1;
      if (!resolved(p1)) continue;
      const int code = fill_code(q, R, f, p1);
      const Ez *ez = &p1.ez;
      DpCall tmp;
      if (code != 0) {
        if (f.spec2 && resolved(f.pass2)) ez = &f.pass2.ez;
        else if (f.spec2) return false;
        else {
          submit(q, tmp, R.rev, f.qs, f.qe - f.qs, R.rid, f.rs, f.re - f.rs, f.bw1, code == 2 ? opt.zdrop_inv : opt.zdrop, -1, 0);
          if (!resolved(tmp)) return false;  // queued now (or pending): known after the next wave
          ez = &tmp.ez;
        }
      }
      if (ez->zdropped) {
        int j;
        for (j = f.i - 1; j >= 0; --j)
          if ((int32_t)a[R.as1 + j].x <= f.rs + ez->max_t) break;
        if (j < 0) j = 0;
        if (R.cnt1 - (j + 1) < opt.min_cnt) return false;
        mm_reg1_t r = R.r;
        memset(&r2, 0, sizeof(r2));
        split_reg(r, r2, R.as1 + j + 1 - r.as, q.qlen, a);
        if (code == 2) r2.split_inv = 1;
        return r2.cnt > 0;
      }
    }
    return false;
  }
  // Plans the remainders a waiting hit is going to leave behind, generation after generation, so that their new
  // windows (left extensions, first fills) run in the NEXT wave instead of one wave per generation.  Nothing of this
  // touches the real state: the plans live in throw-away regions, the anchor flags they set are put back, and what
  // the windows compute reaches the real remainders through the result cache when they are planned for real.
  void speculate_remainders(QCtx &q, Region &R0) const {
    if (!dp_reuse || spec_depth <= 0 || R0.fills.empty()) return;
    mm_reg1_t r2;
    static const bool trace = getenv("PGMM_TRACE") != nullptr;
    if (!predict_split(q, R0, r2)) {
      if (trace) fprintf(stderr, "[pgmm trace] speculate: hit as=%d cnt=%d: no split predicted (%zu fills, %zu jobs queued)\n", R0.r.as, R0.r.cnt, R0.fills.size(), q.jobs.size());
      return;
    }
    if (trace) fprintf(stderr, "[pgmm trace] speculate: hit as=%d cnt=%d -> remainder as=%d cnt=%d\n", R0.r.as, R0.r.cnt, r2.as, r2.cnt);
    const int32_t lo = R0.r.as, hi = R0.r.as + R0.r.cnt;
    std::vector<uint64_t> saved((size_t)(hi - lo));
    for (int32_t i = lo; i < hi; ++i) saved[i - lo] = q.a[i].y;
    for (int depth = 0; depth < spec_depth; ++depth) {
      Region S;
      S.r = r2;
      plan_region(q, S);
      mm_reg1_t r3;
      if (S.fills.empty() || !predict_split(q, S, r3)) break;
      if (trace) fprintf(stderr, "[pgmm trace] speculate:   depth %d: remainder as=%d cnt=%d -> as=%d cnt=%d (%zu jobs queued)\n", depth, r2.as, r2.cnt, r3.as, r3.cnt, q.jobs.size());
      r2 = r3;
    }
    for (int32_t i = lo; i < hi; ++i) q.a[i].y = saved[i - lo];
  }

  static void collect(const QCtx &q, DpCall &c, const std::shared_ptr<const KswBatchResult> &res) {
    if (c.job < 0) return;
    const size_t g = q.job_base + (size_t)c.job;
    c.ez = res->out[g];
    c.cigar = res->cigar.data() + res->cig_start[g];
    c.job = -1;
  }

  // ---------------- mm_align1, first half: every DP window of a hit follows from its anchors (align.c:575-805) ----
  void plan_region(QCtx &q, Region &R) const {
    mm_reg1_t &r = R.r;
    U128 *a = q.a.data();
    const int qlen = q.qlen;
    R.fills.clear();
    R.has_left = R.has_right = false;
    if (r.cnt == 0) {
      R.state = Region::DONE;
      return;
    }
    const int32_t rid = (int32_t)(a[r.as].x << 1 >> 33), rev = (int32_t)(a[r.as].x >> 63);
    const int32_t tlen_full = (int32_t)ts.lens[rid];
    R.rid = rid, R.rev = rev;
    int bw = (int)(opt.bw * 1.5 + 1.), bw_long = (int)(opt.bw_long * 1.5 + 1.);
    if (bw_long < bw) bw_long = bw;
    int32_t as1, cnt1;
    if (!(opt.flag & MM_F_NO_END_FLT)) fix_bad_ends(r, a, opt.bw, opt.min_chain_score * 2, &as1, &cnt1);
    else as1 = r.as, cnt1 = r.cnt;
    filter_bad_seeds(a, as1, cnt1, 10, 40, opt.max_gap >> 1, 10);
    filter_bad_seeds_alt(a, as1, cnt1, 30, opt.max_gap >> 1);
    const int half_k = ts.k >> 1;  // mm_adjust_minier without HPC, align.c:355-370
    int32_t rs = (int32_t)a[as1].x - half_k, qs = (int32_t)a[as1].y - half_k;
    int32_t re = (int32_t)a[as1 + cnt1 - 1].x - half_k, qe = (int32_t)a[as1 + cnt1 - 1].y - half_k;
    R.as1 = as1, R.cnt1 = cnt1;

    // how far the two end extensions may reach (align.c:632-696)
    int32_t rs0, qs0, re0, qe0, rs1, qs1, re1, qe1, l;
    rs0 = (int32_t)a[r.as].x + 1 - (int32_t)(a[r.as].y >> 32 & 0xff);
    qs0 = (int32_t)a[r.as].y + 1 - (int32_t)(a[r.as].y >> 32 & 0xff);
    if (rs0 < 0) rs0 = 0;
    rs1 = qs1 = 0;
    l = 0;
    for (int32_t i = r.as - 1; i >= 0 && a[i].x >> 32 == a[r.as].x >> 32; --i) {
      const int32_t x = (int32_t)a[i].x + 1 - (int32_t)(a[i].y >> 32 & 0xff);
      const int32_t y = (int32_t)a[i].y + 1 - (int32_t)(a[i].y >> 32 & 0xff);
      if (x < rs0 && y < qs0) {
        if (++l > opt.min_cnt) {
          l = rs0 - x > qs0 - y ? rs0 - x : qs0 - y;
          rs1 = rs0 - l, qs1 = qs0 - l;
          if (rs1 < 0) rs1 = 0;
          break;
        }
      }
    }
    if (qs > 0 && rs > 0) {
      l = qs < opt.max_gap ? qs : opt.max_gap;
      qs1 = qs1 > qs - l ? qs1 : qs - l;
      qs0 = qs0 < qs1 ? qs0 : qs1;
      l += l * opt.a > opt.q ? (l * opt.a - opt.q) / opt.e : 0;
      l = l < opt.max_gap ? l : opt.max_gap;
      l = l < rs ? l : rs;
      rs1 = rs1 > rs - l ? rs1 : rs - l;
      rs0 = rs0 < rs1 ? rs0 : rs1;
      rs0 = rs0 < rs ? rs0 : rs;
    } else rs0 = rs, qs0 = qs;
    re0 = (int32_t)a[r.as + r.cnt - 1].x + 1;
    qe0 = (int32_t)a[r.as + r.cnt - 1].y + 1;
    re1 = tlen_full, qe1 = qlen;
    l = 0;
    for (int32_t i = r.as + r.cnt; i < q.n_a && a[i].x >> 32 == a[r.as].x >> 32; ++i) {
      const int32_t x = (int32_t)a[i].x + 1, y = (int32_t)a[i].y + 1;
      if (x > re0 && y > qe0) {
        if (++l > opt.min_cnt) {
          l = x - re0 > y - qe0 ? x - re0 : y - qe0;
          re1 = re0 + l, qe1 = qe0 + l;
          break;
        }
      }
    }
    if (qe < qlen && re < tlen_full) {
      l = qlen - qe < opt.max_gap ? qlen - qe : opt.max_gap;
      qe1 = qe1 < qe + l ? qe1 : qe + l;
      qe0 = qe0 > qe1 ? qe0 : qe1;
      l += l * opt.a > opt.q ? (l * opt.a - opt.q) / opt.e : 0;
      l = l < opt.max_gap ? l : opt.max_gap;
      l = l < tlen_full - re ? l : tlen_full - re;
      re1 = re1 < re + l ? re1 : re + l;
      re0 = re0 > re1 ? re0 : re1;
    } else re0 = re, qe0 = qe;
    if (a[r.as].y & SEED_SELF) {
      int max_ext = r.qs > r.rs ? r.qs - r.rs : r.rs - r.qs;
      if (r.rs - rs0 > max_ext) rs0 = r.rs - max_ext;
      if (r.qs - qs0 > max_ext) qs0 = r.qs - max_ext;
      max_ext = r.qe > r.re ? r.qe - r.re : r.re - r.qe;
      if (re0 - r.re > max_ext) re0 = r.re + max_ext;
      if (qe0 - r.qe > max_ext) qe0 = r.qe + max_ext;
    }
    assert(re0 > rs0);
    R.rs = rs, R.qs = qs, R.re = re, R.qe = qe, R.rs0 = rs0, R.qs0 = qs0, R.re0 = re0, R.qe0 = qe0;

    if (qs > 0 && rs > 0) {  // left extension: both windows read back to front (align.c:702-722)
      R.has_left = true;
      submit(q, R.left, rev, qs0, qs - qs0, rid, rs0, rs - rs0, bw, r.split_inv ? opt.zdrop_inv : opt.zdrop, opt.end_bonus,
             KSW_EXTZ_ONLY | KSW_RIGHT | KSW_REV_CIGAR | KSW_JOB_REVSEQ);
    }
    {  // gap fills between anchors (align.c:726-757): first pass of every fill, approximate max, effectively unbanded
      int32_t frs = rs, fqs = qs;
      {  // how many fills there will be: the vector is sized once (a Fill is ~200 bytes)
        size_t n_fill = 0;
        int32_t crs = rs, cqs = qs;
        for (int32_t i = 1; i < cnt1; ++i) {
          if ((a[as1 + i].y & (SEED_IGNORE | SEED_TANDEM)) && i != cnt1 - 1) continue;
          const int32_t fre = (int32_t)a[as1 + i].x - half_k, fqe = (int32_t)a[as1 + i].y - half_k;
          if (i == cnt1 - 1 || (a[as1 + i].y & SEED_LONG_JOIN) || (fqe - cqs >= opt.min_ksw_len && fre - crs >= opt.min_ksw_len)) ++n_fill, crs = fre, cqs = fqe;
        }
        R.fills.reserve(n_fill);
      }
      for (int32_t i = 1; i < cnt1; ++i) {
        if ((a[as1 + i].y & (SEED_IGNORE | SEED_TANDEM)) && i != cnt1 - 1) continue;
        const int32_t fre = (int32_t)a[as1 + i].x - half_k, fqe = (int32_t)a[as1 + i].y - half_k;
        if (i == cnt1 - 1 || (a[as1 + i].y & SEED_LONG_JOIN) || (fqe - fqs >= opt.min_ksw_len && fre - frs >= opt.min_ksw_len)) {
          Fill f;
          f.i = i, f.rs = frs, f.qs = fqs, f.re = fre, f.qe = fqe, f.bw1 = bw_long;
          if (a[as1 + i].y & SEED_LONG_JOIN) f.bw1 = fqe - fqs > fre - frs ? fqe - fqs : fre - frs;
          R.fills.push_back(std::move(f));
          frs = fre, fqs = fqe;
        }
      }
      for (Fill &f : R.fills)
        submit(q, f.pass1, rev, f.qs, f.qe - f.qs, rid, f.rs, f.re - f.rs, f.bw1, opt.zdrop, -1, KSW_APPROX_MAX);
      // A long fill spans a big indel or an inversion and all but always fails mm_test_zdrop: its exact second pass
      // (align.c:757-759) is queued together with the first one instead of one wave later -- speculative like the end
      // extensions, and only when both z-drop thresholds agree, so that the pass does not depend on the verdict.
      if (opt.zdrop == opt.zdrop_inv && spec_fill_len > 0)
        for (Fill &f : R.fills)
          if (std::max(f.qe - f.qs, f.re - f.rs) >= spec_fill_len) {
            submit(q, f.pass2, rev, f.qs, f.qe - f.qs, rid, f.rs, f.re - f.rs, f.bw1, opt.zdrop, -1, 0);
            f.spec2 = true;
          }
    }
    if (qe < qe0 && re < re0) {  // right extension; speculative: unused if a fill z-drops (align.c:789-805)
      R.has_right = true;
      submit(q, R.right, rev, qe, qe0 - qe, rid, re, re0 - re, bw, opt.zdrop, opt.end_bonus, KSW_EXTZ_ONLY);
    }
    R.state = Region::WAIT1;
  }

  // mm_test_zdrop (align.c:33-89) in two halves.  The scan of the first-pass CIGAR (largest score drop along the path
  // and where it happens) comes back from the DP kernel with the result (KswOut::zd_*); a backend that does not fill it
  // in (the CPU test seam) sets zd_max < 0 and the scan is done here.  The decision, including the striped local
  // alignment that looks for an inversion inside the dropped window, is made on the host.
  void zdrop_scan(const uint8_t *qseq, const uint8_t *tseq, const uint32_t *cigar, int n_cigar, Ez &ez) const {
    int32_t score = 0, max = INT32_MIN, max_i = -1, max_j = -1, i = 0, j = 0, max_zdrop = 0;
    int pos[2][2] = {{-1, -1}, {-1, -1}};
    const auto upd = [&](int32_t sc, int ii, int jj) {
      if (sc < max) {
        const int li = ii - max_i, lj = jj - max_j, diff = li > lj ? li - lj : lj - li, z = max - sc - diff * opt.e;
        if (z > max_zdrop) {
          max_zdrop = z;
          pos[0][0] = max_i, pos[0][1] = ii;
          pos[1][0] = max_j, pos[1][1] = jj;
        }
      } else max = sc, max_i = ii, max_j = jj;
    };
    for (int ci = 0; ci < n_cigar; ++ci) {
      const uint32_t c = cigar[ci];
      const uint32_t op = c & 0xf, len = c >> 4;
      if (op == MM_CIGAR_MATCH) {
        for (uint32_t l = 0; l < len; ++l) {
          score += mat[tseq[i + l] * 5 + qseq[j + l]];
          upd(score, i + (int)l, j + (int)l);
        }
        i += len, j += len;
      } else if (op == MM_CIGAR_INS || op == MM_CIGAR_DEL || op == 3) {
        score -= opt.q + opt.e * (int)len;
        if (op == MM_CIGAR_INS) j += len;
        else i += len;
        upd(score, i, j);
      }
    }
    ez.zd_max = max_zdrop, ez.zd_t0 = pos[0][0], ez.zd_t1 = pos[0][1], ez.zd_q0 = pos[1][0], ez.zd_q1 = pos[1][1];
  }
  int test_zdrop(const uint8_t *qseq, const uint8_t *tseq, const Ez &ez) const {
    const int max_zdrop = ez.zd_max;
    const int q_len = ez.zd_q1 - ez.zd_q0, t_len = ez.zd_t1 - ez.zd_t0;
    if (!(opt.flag & (MM_F_SPLICE | MM_F_SR | MM_F_FOR_ONLY | MM_F_REV_ONLY)) && max_zdrop > opt.zdrop_inv && q_len < opt.max_gap &&
        t_len < opt.max_gap) {
      std::vector<uint8_t> q2((size_t)(q_len > 0 ? q_len : 0));
      for (int k = 0; k < q_len; ++k) {
        const int c = qseq[ez.zd_q1 - k - 1];
        q2[k] = (uint8_t)(c >= 4 ? 4 : 3 - c);
      }
      int q_off, t_off;
      // only "does the local score reach both thresholds" matters here (align.c:62-64): the scan stops as soon as it does
      const int need = std::max(std::max(opt.min_chain_score * opt.a, opt.min_dp_max), 1);
      const int sc = ll_local_score(q_len, q2.data(), t_len, tseq + ez.zd_t0, mat, opt.q, opt.e, &q_off, &t_off, need);
      if (sc >= opt.min_chain_score * opt.a && sc >= opt.min_dp_max) return 2;
    }
    return max_zdrop > opt.zdrop ? 1 : 0;
  }

  // after the first wave: collect results, test every fill, queue the exact second passes (align.c:757-759)
  // returns whether a second wave is needed for this hit
  bool after_pass1(QCtx &q, Region &R, const std::shared_ptr<const KswBatchResult> &res) const {
    if (R.has_left) collect(q, R.left, res);
    if (R.has_right) collect(q, R.right, res);
    bool any = false;
    for (Fill &f : R.fills)
      if (f.spec2) collect(q, f.pass2, res);  // all of them: their slots belong to this wave
    bool behind = false;  // behind a fill that is known to drop: this hit never uses these fills, its remainders will
    for (Fill &f : R.fills) {
      collect(q, f.pass1, res);
      if (behind) {
        if (!dp_reuse || spec_depth <= 0) continue;  // (collected all the same: a slot left pending would never be ready)
        if (fill_code(q, R, f, f.pass1) != 0 && !f.spec2) {  // exact pass for the remainders' benefit (result cache only)
          DpCall tmp;
          const int code = fill_code(q, R, f, f.pass1);
          submit(q, tmp, R.rev, f.qs, f.qe - f.qs, R.rid, f.rs, f.re - f.rs, f.bw1, code == 2 ? opt.zdrop_inv : opt.zdrop, -1, 0);
        }
        continue;
      }
      f.code = fill_code(q, R, f, f.pass1);
      if (f.code != 0 && !f.spec2) {
        submit(q, f.pass2, R.rev, f.qs, f.qe - f.qs, R.rid, f.rs, f.re - f.rs, f.bw1, f.code == 2 ? opt.zdrop_inv : opt.zdrop, -1, 0);
        any |= f.pass2.job >= 0;
      }
      if ((f.code ? f.pass2.job < 0 && f.pass2.ez.zdropped : f.pass1.ez.zdropped)) behind = true;  // later fills can never be used
    }
    R.state = Region::WAIT2;  // finishing happens in one place: right away when nothing is pending, else after the second wave
    return any;
  }

  // splits a hit at its n-th anchor (hit.c:106-123)
  static void split_reg(mm_reg1_t &r, mm_reg1_t &r2, int n, int qlen, const U128 *a) {
    if (n <= 0 || n >= r.cnt) return;
    r2 = r;
    r2.id = -1;
    r2.sam_pri = 0;
    r2.p = nullptr;
    r2.split_inv = 0;
    r2.cnt = r.cnt - n;
    r2.score = (int32_t)(r.score * ((float)r2.cnt / r.cnt) + .499);
    r2.as = r.as + n;
    if (r.parent == r.id) r2.parent = PARENT_TMP_PRI;
    set_coor(r2, qlen, a);
    r.cnt -= r2.cnt;
    r.score -= r2.score;
    set_coor(r, qlen, a);
    r.split |= 1, r2.split |= 2;
  }

  // mm_align1, second half: stitch the CIGAR in order, truncate and split on z-drop (align.c:715-827).
  // Returns the split-off hit (cnt > 0) if there is one.
  mm_reg1_t finish_region(QCtx &q, Region &R, const std::shared_ptr<const KswBatchResult> &res) const {
    mm_reg1_t &r = R.r;
    mm_reg1_t r2;
    memset(&r2, 0, sizeof(r2));
    const U128 *a = q.a.data();
    const int qlen = q.qlen;
    for (Fill &f : R.fills) collect(q, f.pass2, res);
    int32_t rs1, qs1, re1, qe1;
    bool dropped = false;
    if (R.has_left) {
      const Ez &ez = R.left.ez;
      if (ez.n_cigar > 0) {
        append_cigar(R, R.left.cigar, ez.n_cigar);
        R.dp_score += ez.max;
      }
      rs1 = R.rs - (ez.reach_end ? ez.mqe_t + 1 : ez.max_t + 1);
      qs1 = R.qs - (ez.reach_end ? R.qs - R.qs0 : ez.max_q + 1);
    } else rs1 = R.rs, qs1 = R.qs;
    re1 = R.rs, qe1 = R.qs;
    if (R.cnt1 > 1) re1 = R.re, qe1 = R.qe;  // the last anchor is always visited by the fill loop
    for (Fill &f : R.fills) {
      const DpCall &c = f.code ? f.pass2 : f.pass1;
      const Ez &ez = c.ez;
      if (ez.n_cigar > 0) append_cigar(R, c.cigar, ez.n_cigar);
      if (ez.zdropped) {
        if (!R.has_p) R.has_p = true, R.capacity = roundup_pow2(6);
        int j;
        for (j = f.i - 1; j >= 0; --j)
          if ((int32_t)a[R.as1 + j].x <= f.rs + ez.max_t) break;
        dropped = true;
        if (j < 0) j = 0;
        R.dp_score += ez.max;
        re1 = f.rs + (ez.max_t + 1);
        qe1 = f.qs + (ez.max_q + 1);
        if (R.cnt1 - (j + 1) >= opt.min_cnt) {
          split_reg(r, r2, R.as1 + j + 1 - r.as, qlen, a);
          if (f.code == 2) r2.split_inv = 1;
        }
        break;
      } else R.dp_score += ez.score;
    }
    if (!dropped && R.has_right) {
      const Ez &ez = R.right.ez;
      if (ez.n_cigar > 0) {
        append_cigar(R, R.right.cigar, ez.n_cigar);
        R.dp_score += ez.max;
      }
      re1 = R.re + (ez.reach_end ? ez.mqe_t + 1 : ez.max_t + 1);
      qe1 = R.qe + (ez.reach_end ? R.qe0 - R.qe : ez.max_q + 1);
    }
    assert(qe1 <= qlen);
    r.rs = rs1, r.re = re1;
    if (!R.rev) r.qs = qs1, r.qe = qe1;
    else r.qs = qlen - qe1, r.qe = qlen - qs1;
    if (R.has_p) update_extra(R, q.q0[r.rev] + qs1, tseq(R.rid, rs1));
    R.fills.clear();
    R.fills.shrink_to_fit();
    return r2;
  }

  // ---------------- inversion rescue between a hit and its split-off neighbour (align.c:830-885) ----------------
  // returns true if a DP window was queued for the inversion candidate
  bool plan_inversion(QCtx &q, const Region &R1, const Region &R2, Region &Rinv) const {
    const mm_reg1_t &r1 = R1.r, &r2 = R2.r;
    if (!(r1.split & 1) || !(r2.split & 2)) return false;
    if (r1.id != r1.parent && r1.parent != PARENT_TMP_PRI) return false;
    if (r2.id != r2.parent && r2.parent != PARENT_TMP_PRI) return false;
    if (r1.rid != r2.rid || r1.rev != r2.rev) return false;
    const int ql = r1.rev ? r1.qs - r2.qe : r2.qs - r1.qe, tl = r2.rs - r1.re;
    if (ql < opt.min_chain_score || ql > opt.max_gap) return false;
    if (tl < opt.min_chain_score || tl > opt.max_gap) return false;
    // the candidate lies on the opposite strand of the flanking hits
    const int strand = r1.rev ? 0 : 1;
    const int32_t qstart = r1.rev ? r2.qe : q.qlen - r2.qs;
    const uint8_t *qseq = q.q0[strand] + qstart, *tsq = tseq(r1.rid, r1.re);
    std::vector<uint8_t> qr(qseq, qseq + ql), tr(tsq, tsq + tl);
    std::reverse(qr.begin(), qr.end());
    std::reverse(tr.begin(), tr.end());
    int q_off, t_off;
    const int score = ll_local_score(ql, qr.data(), tl, tr.data(), mat, opt.q, opt.e, &q_off, &t_off);
    if (score < opt.min_dp_max) return false;
    q_off = ql - (q_off + 1), t_off = tl - (t_off + 1);
    Rinv.inv_q_off = q_off, Rinv.inv_t_off = t_off, Rinv.inv_ql = ql, Rinv.inv_tl = tl;
    Rinv.rid = r1.rid, Rinv.rev = strand;
    submit(q, Rinv.inv, strand, qstart + q_off, ql - q_off, r1.rid, r1.re + t_off, tl - t_off, (int)(opt.bw * 1.5), opt.zdrop, -1,
           KSW_EXTZ_ONLY);
    // geometry needed to finish: remember the flanks
    Rinv.rs0 = r1.re, Rinv.qs0 = qstart;
    Rinv.r.rid = r1.rid;
    Rinv.r.rev = !r1.rev;
    Rinv.qe0 = r2.qe, Rinv.re0 = r2.qs;  // r2->qe and r2->qs, used for the query coordinates below
    return true;
  }
  bool finish_inversion(QCtx &q, Region &Rinv, const std::shared_ptr<const KswBatchResult> &res) const {
    collect(q, Rinv.inv, res);
    const Ez &ez = Rinv.inv.ez;
    if (ez.n_cigar == 0) return false;
    mm_reg1_t &ri = Rinv.r;
    const int32_t rid = ri.rid;
    const uint32_t rev = ri.rev;
    memset(&ri, 0, sizeof(ri));
    append_cigar(Rinv, Rinv.inv.cigar, ez.n_cigar);
    Rinv.dp_score = ez.max;
    ri.id = -1, ri.parent = PARENT_UNSET, ri.inv = 1, ri.rev = rev, ri.rid = rid, ri.div = -1.0f;
    if (ri.rev == 0) {
      ri.qs = Rinv.qe0 + Rinv.inv_q_off;
      ri.qe = ri.qs + ez.max_q + 1;
    } else {
      ri.qe = Rinv.re0 - Rinv.inv_q_off;
      ri.qs = ri.qe - (ez.max_q + 1);
    }
    ri.rs = Rinv.rs0 + Rinv.inv_t_off;
    ri.re = ri.rs + ez.max_t + 1;
    update_extra(Rinv, q.q0[Rinv.rev] + Rinv.qs0 + Rinv.inv_q_off, tseq(rid, Rinv.rs0 + Rinv.inv_t_off));
    return true;
  }

  // ---------------- filters, order, mapq (hit.c:290-309,188-218,396-466; align.c:887-960) ----------------
  void filter_regs(QCtx &q) const {
    std::vector<std::unique_ptr<Region>> keep;
    for (auto &R : q.regs) {
      const mm_reg1_t &r = R->r;
      bool flt = false;
      if (!r.inv && !r.seg_split && r.cnt < opt.min_cnt) flt = true;
      if (R->has_p) {
        if (r.mlen < opt.min_chain_score) flt = true;
        else if (R->dp_max < opt.min_dp_max) flt = true;
        else if (r.qs > q.qlen * opt.max_clip_ratio && q.qlen - r.qe > q.qlen * opt.max_clip_ratio) flt = true;
      }
      if (!flt) keep.push_back(std::move(R));
    }
    q.regs.swap(keep);
  }
  static double event_identity(const Region &R) {
    if (!R.has_p) return -1.0f;
    int32_t n_gap = 0, n_gapo = 0;
    for (uint32_t c : R.cigar) {
      const int32_t op = c & 0xf, len = c >> 4;
      if (op == MM_CIGAR_INS || op == MM_CIGAR_DEL) ++n_gapo, n_gap += len;
    }
    return (double)R.r.mlen / (R.r.blen + (int32_t)R.n_ambi - n_gap + n_gapo);
  }
  void update_dp_max(QCtx &q) const {
    const int n = (int)q.regs.size();
    const float frac = opt.rank_frac;
    const int a = opt.a, b = opt.b;
    int32_t max = -1, max2 = -1, max_i = -1;
    if (n < 2) return;
    for (int i = 0; i < n; ++i) {
      const Region &R = *q.regs[i];
      if (!R.has_p) continue;
      if (R.dp_max > max) max2 = max, max = R.dp_max, max_i = i;
      else if (R.dp_max > max2) max2 = R.dp_max;
    }
    if (max_i < 0 || max < 0 || max2 < 0) return;
    if (q.regs[max_i]->r.qe - q.regs[max_i]->r.qs < (double)q.qlen * frac) return;
    if (max2 < (double)max * frac) return;
    double div = 1. - event_identity(*q.regs[max_i]);
    if (div < 0.02) div = 0.02;
    double b2 = 0.5 / div;
    if (b2 * a < b) b2 = (double)a / b;
    for (int i = 0; i < n; ++i) {
      Region &R = *q.regs[i];
      if (!R.has_p) continue;
      int32_t n_gap = 0, n_gapo = 0;
      double gap_cost = 0.0;
      for (uint32_t c : R.cigar) {
        const int32_t op = c & 0xf, len = c >> 4;
        if (op == MM_CIGAR_INS || op == MM_CIGAR_DEL) {
          gap_cost += b2 + (double)fast_log2f(1.0f + (float)len);
          ++n_gapo, n_gap += len;
        }
      }
      const int32_t n_mis = R.r.blen + (int32_t)R.n_ambi - R.r.mlen - n_gap;
      R.dp_max = (int32_t)(a * (R.r.mlen - b2 * n_mis - gap_cost) + .499);
      if (R.dp_max < 0) R.dp_max = 0;
    }
  }
  void hit_sort(QCtx &q) const {
    const int n = (int)q.regs.size();
    if (n <= 1) return;
    std::vector<U128> aux;
    for (int i = 0; i < n; ++i) {
      const Region &R = *q.regs[i];
      if (R.r.inv || R.r.cnt > 0) {
        const int score = R.has_p ? R.dp_max : R.r.score;
        aux.push_back(U128{(uint64_t)score << 32 | R.r.hash, (uint64_t)i});
      }
    }
    flag_sort_128x(aux.data(), aux.data() + aux.size());
    std::vector<std::unique_ptr<Region>> out;
    for (int i = (int)aux.size() - 1; i >= 0; --i) out.push_back(std::move(q.regs[aux[i].y]));
    q.regs.swap(out);
  }
  void set_mapq(QCtx &q) const {
    const int n = (int)q.regs.size();
    if (n == 0) return;
    const float q_coef = 40.0f;
    const int match_sc = opt.a, min_chain_sc = opt.min_chain_score;
    int64_t sum_sc = 0;
    for (auto &R : q.regs)
      if (R->r.parent == R->r.id) sum_sc += R->r.score;
    const float uniq_ratio = (float)sum_sc / (sum_sc + q.rep_len);
    for (auto &Rp : q.regs) {
      Region &R = *Rp;
      mm_reg1_t &r = R.r;
      if (r.inv) r.mapq = 0;
      else if (r.parent == r.id) {
        int mapq;
        const float pen_s1 = (r.score > 100 ? 1.0f : 0.01f * r.score) * uniq_ratio;
        float pen_cm = r.cnt > 10 ? 1.0f : 0.1f * r.cnt;
        pen_cm = pen_s1 < pen_cm ? pen_s1 : pen_cm;
        const int subsc = r.subsc > min_chain_sc ? r.subsc : min_chain_sc;
        if (R.has_p && R.dp_max2 > 0 && R.dp_max > 0) {
          const float identity = (float)r.mlen / r.blen;
          const float x = (float)R.dp_max2 * subsc / R.dp_max / r.score0;
          mapq = (int)(identity * pen_cm * q_coef * (1.0f - x * x) * logf((float)R.dp_max / match_sc));
          const int mapq_alt = (int)(6.02f * identity * identity * (R.dp_max - R.dp_max2) / match_sc + .499f);
          mapq = mapq < mapq_alt ? mapq : mapq_alt;
        } else {
          const float x = (float)subsc / r.score0;
          if (R.has_p) {
            const float identity = (float)r.mlen / r.blen;
            mapq = (int)(identity * pen_cm * q_coef * (1.0f - x) * logf((float)R.dp_max / match_sc));
          } else mapq = (int)(pen_cm * q_coef * (1.0f - x) * logf((float)r.score));
        }
        mapq -= (int)(4.343f * logf((float)(r.n_sub + 1)) + .499f);
        mapq = mapq > 0 ? mapq : 0;
        r.mapq = mapq < 60 ? mapq : 60;
        if (R.has_p && R.dp_max > R.dp_max2 && r.mapq == 0) r.mapq = 1;
      } else r.mapq = 0;
    }
    // an inversion takes the lower mapq of its neighbours along the target (hit.c:396-419)
    if (n < 3) return;
    bool any_inv = false;
    for (auto &R : q.regs) any_inv |= R->r.inv;
    if (!any_inv) return;
    std::vector<U128> aux;
    for (int i = 0; i < n; ++i)
      if (q.regs[i]->r.parent == i || q.regs[i]->r.parent < 0)
        aux.push_back(U128{(uint64_t)q.regs[i]->r.rid << 32 | (uint32_t)q.regs[i]->r.rs, (uint64_t)i});
    flag_sort_128x(aux.data(), aux.data() + aux.size());
    for (int i = 1; i < (int)aux.size() - 1; ++i) {
      mm_reg1_t &inv = q.regs[aux[i].y]->r;
      if (inv.inv) {
        const mm_reg1_t &l = q.regs[aux[i - 1].y]->r, &rr = q.regs[aux[i + 1].y]->r;
        inv.mapq = l.mapq < rr.mapq ? l.mapq : rr.mapq;
      }
    }
  }
};

void parallel_for(int n, int n_threads, const std::function<void(int)> &fn);

}  // namespace

// Striped local alignment score with eight 16-bit lanes (Farrar's layout; ksw2_ll_sse.c:37-152).  The result depends
// on the striping (ties for the end coordinates, saturation), so the layout is kept: vector j, lane l <-> query
// position j + l*slen.  Written with the SSE2 intrinsics on x86-64 hosts and lane by lane elsewhere.
#if defined(__SSE2__)
int ll_local_score(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat, int gapo, int gape,
                   int *qe, int *te, int stop_at) {
  *qe = *te = -1;
  const int slen = (qlen + 7) / 8;
  if (slen == 0) {
    if (tlen > 0) *te = tlen - 1;
    return 0;
  }
  std::vector<uint64_t> buf((size_t)slen * 9 * 2 + 2);
  __m128i *prof = (__m128i *)(((uintptr_t)buf.data() + 15) & ~(uintptr_t)15), *H0 = prof + (size_t)slen * 5, *H1 = H0 + slen, *E = H1 + slen, *Hmax = E + slen;
  {
    int16_t *t = (int16_t *)prof;
    for (int a = 0; a < 5; ++a)
      for (int j = 0; j < slen; ++j)
        for (int l = 0; l < 8; ++l) {
          const int k = j + l * slen;
          *t++ = (int16_t)(k >= qlen ? 0 : mat[a * 5 + query[k]]);
        }
  }
  const __m128i zero = _mm_setzero_si128(), goe = _mm_set1_epi16((short)(gapo + gape)), ge = _mm_set1_epi16((short)gape);
  for (int j = 0; j < slen; ++j) H0[j] = E[j] = Hmax[j] = zero;
  int gmax = 0;
  for (int i = 0; i < tlen; ++i) {
    const __m128i *S = prof + (size_t)target[i] * slen;
    __m128i f = zero, rowmax = zero;
    __m128i h = _mm_slli_si128(H0[slen - 1], 2);  // the last vector of the previous row, moved up one lane
    for (int j = 0; j < slen; ++j) {
      h = _mm_adds_epi16(h, S[j]);
      __m128i e = E[j];
      h = _mm_max_epi16(h, e);
      h = _mm_max_epi16(h, f);
      rowmax = _mm_max_epi16(rowmax, h);
      H1[j] = h;
      h = _mm_subs_epu16(h, goe);
      e = _mm_max_epi16(_mm_subs_epu16(e, ge), h);
      E[j] = e;
      f = _mm_max_epi16(_mm_subs_epu16(f, ge), h);
      h = H0[j];
    }
    for (int k = 0, done = 0; k < 8 && !done; ++k) {  // lazy F: propagate horizontal gaps across the stripes
      f = _mm_slli_si128(f, 2);
      for (int j = 0; j < slen; ++j) {
        h = _mm_max_epi16(H1[j], f);
        H1[j] = h;
        h = _mm_subs_epu16(h, goe);
        f = _mm_subs_epu16(f, ge);
        if (!_mm_movemask_epi8(_mm_cmpgt_epi16(f, h))) {
          done = 1;
          break;
        }
      }
    }
    __m128i m = _mm_max_epi16(rowmax, _mm_srli_si128(rowmax, 8));
    m = _mm_max_epi16(m, _mm_srli_si128(m, 4));
    m = _mm_max_epi16(m, _mm_srli_si128(m, 2));
    const int imax = _mm_extract_epi16(m, 0);
    if (imax >= gmax) {
      gmax = imax, *te = i;
      if (stop_at > 0 && gmax >= stop_at) return gmax;  // the caller only asks whether the score reaches stop_at
      memcpy(Hmax, H1, (size_t)slen * sizeof(__m128i));
    }
    std::swap(H0, H1);
  }
  const uint16_t *H8 = (const uint16_t *)Hmax;
  for (int i = 0; i < slen * 8; ++i)
    if ((int)H8[i] == gmax) *qe = i / 8 + i % 8 * slen;
  return gmax;
}
#else
// striped local alignment score with 16-bit lanes, restated lane by lane (ksw2_ll_sse.c:37-152): the result depends on
// the striped layout (ties for the end positions, saturation), so the layout is kept: vector j, lane l <-> query j+l*slen
int ll_local_score(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat, int gapo, int gape,
                   int *qe, int *te, int stop_at) {
  (void)stop_at;
  typedef int16_t v8 __attribute__((vector_size(16)));
  *qe = *te = -1;
  const int slen = (qlen + 7) / 8;
  if (slen == 0) {  // empty query: the reference's loops run over zero vectors; every row maximum is 0
    if (tlen > 0) *te = tlen - 1;
    return 0;
  }
  std::vector<v8> prof((size_t)slen * 5), H0(slen), H1(slen), E(slen), Hmax(slen);
  for (int a = 0; a < 5; ++a)
    for (int j = 0; j < slen; ++j)
      for (int l = 0; l < 8; ++l) {
        const int k = j + l * slen;
        prof[(size_t)a * slen + j][l] = k >= qlen ? 0 : mat[a * 5 + query[k]];
      }
  const v8 zero = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int j = 0; j < slen; ++j) H0[j] = E[j] = Hmax[j] = zero;
  const uint16_t goe = (uint16_t)(gapo + gape), ge = (uint16_t)gape;
  const auto adds = [](v8 a, v8 b) {  // signed saturating add
    v8 r;
    for (int l = 0; l < 8; ++l) {
      int s = (int)a[l] + (int)b[l];
      r[l] = (int16_t)(s > 32767 ? 32767 : s < -32768 ? -32768 : s);
    }
    return r;
  };
  const auto subsu = [](v8 a, uint16_t b) {  // unsigned saturating subtract
    v8 r;
    for (int l = 0; l < 8; ++l) {
      const uint16_t x = (uint16_t)a[l];
      r[l] = (int16_t)(uint16_t)(x > b ? x - b : 0);
    }
    return r;
  };
  const auto smax = [](v8 a, v8 b) {
    v8 r;
    for (int l = 0; l < 8; ++l) r[l] = a[l] > b[l] ? a[l] : b[l];
    return r;
  };
  const auto shl1 = [](v8 a) {  // move every lane up by one, lane 0 <- 0
    v8 r;
    r[0] = 0;
    for (int l = 1; l < 8; ++l) r[l] = a[l - 1];
    return r;
  };
  int gmax = 0;
  v8 *h0 = H0.data(), *h1 = H1.data();
  for (int i = 0; i < tlen; ++i) {
    v8 e, h, f = zero, max = zero;
    const v8 *S = prof.data() + (size_t)target[i] * slen;
    h = shl1(h0[slen - 1]);
    for (int j = 0; j < slen; ++j) {
      h = adds(h, S[j]);
      e = E[j];
      h = smax(h, e);
      h = smax(h, f);
      max = smax(max, h);
      h1[j] = h;
      h = subsu(h, goe);
      e = subsu(e, ge);
      e = smax(e, h);
      E[j] = e;
      f = subsu(f, ge);
      f = smax(f, h);
      h = h0[j];
    }
    for (int k = 0; k < 8; ++k) {  // lazy F
      bool done = false;
      f = shl1(f);
      for (int j = 0; j < slen; ++j) {
        h = h1[j];
        h = smax(h, f);
        h1[j] = h;
        h = subsu(h, goe);
        f = subsu(f, ge);
        bool any = false;
        for (int l = 0; l < 8; ++l) any |= f[l] > h[l];
        if (!any) {
          done = true;
          break;
        }
      }
      if (done) break;
    }
    int imax = max[0];
    for (int l = 1; l < 8; ++l) imax = imax > max[l] ? imax : max[l];
    if (imax >= gmax) {
      gmax = imax, *te = i;
      memcpy(Hmax.data(), h1, (size_t)slen * sizeof(v8));
    }
    std::swap(h0, h1);
  }
  for (int i = 0; i < slen * 8; ++i)
    if ((int)(uint16_t)Hmax[i / 8][i % 8] == gmax) *qe = i / 8 + i % 8 * slen;
  return gmax;
}

#endif

namespace {
void parallel_for(int n, int n_threads, const std::function<void(int)> &fn) {
  if (n_threads <= 1 || n <= 1) {
    for (int i = 0; i < n; ++i) fn(i);
    return;
  }
  std::atomic<int> next(0);
  std::vector<std::thread> th;
  const int nt = std::min(n, n_threads);
  for (int t = 0; t < nt; ++t)
    th.emplace_back([&]() {
      for (;;) {
        const int i = next.fetch_add(1);
        if (i >= n) break;
        fn(i);
      }
    });
  for (auto &t : th) t.join();
}
}  // namespace

void map_batch(Backend &be, const TargetSet &ts, QueryBatch &qb, const mm_mapopt_t &opt, int *n_regs, mm_reg1_t **regs,
               int n_threads) {
  for (int i = 0; i < qb.n; ++i) n_regs[i] = 0, regs[i] = nullptr;
  if (qb.n == 0) return;
  // the flag set pangraph uses is the one this pipeline implements; anything else has no path here
  const int64_t need = MM_F_CIGAR | MM_F_RMQ | MM_F_ALL_CHAINS | MM_F_NO_LJOIN;
  const int64_t unsupported = MM_F_SPLICE | MM_F_SR | MM_F_FRAG_MODE | MM_F_HEAP_SORT | MM_F_QSTRAND | MM_F_EQX | MM_F_FOR_ONLY |
                              MM_F_REV_ONLY | MM_F_INDEPEND_SEG;
  if ((opt.flag & need) != need || (opt.flag & unsupported) || opt.sdust_thres > 0 || opt.split_prefix) {
    fprintf(stderr, "[pgmm_b200] fatal: mapping flags 0x%llx are outside pangraph's path (need CIGAR|RMQ|ALL_CHAINS|NO_LJOIN, "
                    "no splice/sr/qstrand/eqx); there is no fallback implementation\n", (long long)opt.flag);
    abort();
  }
  if (opt.q == opt.q2 && opt.e == opt.e2) {
    fprintf(stderr, "[pgmm_b200] fatal: single-affine gap scoring (q==q2, e==e2) is not on pangraph's path\n");
    abort();
  }
  Mapper M(ts, qb, opt);
  const auto now = []() {
    timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return t.tv_sec * 1e3 + t.tv_nsec * 1e-6;
  };
  const auto cpu_ms = []() {  // CPU time of the whole process (meaningful per phase when one round runs alone)
    timespec t;
    clock_gettime(CLOCK_PROCESS_CPUTIME_ID, &t);
    return t.tv_sec * 1e3 + t.tv_nsec * 1e-6;
  };
  const bool trace_cpu = getenv("PGMM_TRACE") != nullptr;
  double cpu0 = cpu_ms();
  const auto cpu_mark = [&](const char *what) {
    if (!trace_cpu) return;
    const double c = cpu_ms();
    fprintf(stderr, "[pgmm trace] cpu %-28s %8.1f ms\n", what, c - cpu0);
    cpu0 = c;
  };
  double t0 = now(), t1;
  {
    CpuScope cpu_scope(0);  // (the helper threads account for themselves; this is the caller's share)
    encode_queries(qb, ts, std::min(n_threads, 4));
  }
  cpu_mark("encode");
  t1 = now(), be.stats.t_encode += t1 - t0, t0 = t1;
  std::vector<QuerySeeds> seeds;
  {
    CpuScope cpu_scope(1);
    be.begin_batch(ts, qb);
    be.seed_batch(ts, qb, opt, seeds);
  }
  t1 = now(), be.stats.t_seed += t1 - t0, t0 = t1;
  cpu_mark("seeding (host side)");

  std::vector<QCtx> Q(qb.n);
  const float pen_gap = (float)(opt.chain_gap_scale * 0.01 * ts.k), pen_skip = (float)(opt.chain_skip_scale * 0.01 * ts.k);
  ChainParams cp{opt.max_gap, opt.rmq_inner_dist, opt.bw, opt.max_chain_skip, opt.rmq_size_cap, opt.min_cnt, opt.min_chain_score,
                 pen_gap, pen_skip};

  // ---- per query: anchor order (host), chain scores (device), backtrack + hit skeletons + first DP plan (host) ----
  const bool trace = getenv("PGMM_TRACE") != nullptr;
  std::vector<ChainFillJob> cjobs(qb.n);
  parallel_for(qb.n, n_threads, [&](int i) {
    CpuScope cpu_scope(2);
    QCtx &q = Q[i];
    q.qi = i, q.qlen = qb.lens[i], q.qname = qb.names[i], q.qbase = qb.base[i];
    q.q0[0] = qb.from_targets ? ts.codes.data() + ts.offs[i] : qb.codes.data() + q.qbase, q.q0[1] = qb.codes.data() + q.qbase + q.qlen;
    if (q.qlen == 0) return;
    if (opt.max_qlen > 0 && q.qlen > opt.max_qlen) return;
    uint32_t h = q.qname && !(opt.flag & MM_F_NO_HASH_NAME) ? x31_hash(q.qname) : 0;  // map.c:246-248
    h ^= wang32((uint32_t)q.qlen) + wang32((uint32_t)opt.seed);
    q.hash = wang32(h);
    q.a.swap(seeds[i].a);
    q.mini_pos.swap(seeds[i].mini_pos);
    q.rep_len = seeds[i].rep_len;
    if (!seeds[i].sorted) flag_sort_128x(q.a.data(), q.a.data() + q.a.size());  // map.c:202 (unless the device proved the order unique)
    ChainFillJob &cj = cjobs[i];
    cj.a = q.a.data(), cj.n = (int64_t)q.a.size();
    chain_find_segments(cp, cj.a, cj.n, cj.segs);
    if (const char *dump = getenv("PGMM_DUMP_ANCHORS")) {  // profiling aid: sorted anchors of every query, 16 bytes each
      char path[4096];
      snprintf(path, sizeof(path), "%s.%d", dump, i);
      if (FILE *fp = fopen(path, "wb")) {
        fwrite(q.a.data(), sizeof(U128), q.a.size(), fp);
        fclose(fp);
      }
    }
  });
  const double tc0 = now();
  be.stats.t_chain_sort += tc0 - t0;
  cpu_mark("anchor sort + segments");
  {
    CpuScope cpu_scope(3);
    be.chain_fill(cp, cjobs);
  }
  cpu_mark("chain fill (host side)");
  const double tc1 = now();
  be.stats.t_chain_fill += tc1 - tc0;
  parallel_for(qb.n, n_threads, [&](int i) {
    CpuScope cpu_scope(4);
    QCtx &q = Q[i];
    ChainFillJob &cj = cjobs[i];
    if (cj.n == 0) return;
    double c0 = now(), c1;
    std::vector<int32_t> t((size_t)cj.n, 0);
    int64_t n_redo = 0;
    for (size_t k = 0; k < cj.segs.size(); ++k)
      if (cj.redo[k]) {
        chain_fill_host(cp, cj.a, cj.n, cj.segs[k].start, cj.segs[k].end, cj.f, cj.p, cj.v, t.data());
        n_redo += cj.segs[k].end - cj.segs[k].start;
      }
    c1 = now();
    const double t_redo = c1 - c0;
    c0 = c1;
    std::vector<uint64_t> u;
    chain_backtrack(cp, q.a, cj.f, cj.p, cj.v, t.data(), u);
    c1 = now();
    const double t_bt = c1 - c0;
    c0 = c1;
    M.gen_regs(q, u);
    M.est_err(q);
    q.n_a = q.regs.empty() ? 0 : Mapper::squeeze_a(q);
    c1 = now();
    const double t_regs = c1 - c0;
    c0 = c1;
    for (auto &R : q.regs) M.plan_region(q, *R);
    c1 = now();
    int64_t seg_max = 0;
    for (const ChainSeg &sg : cj.segs) seg_max = std::max(seg_max, sg.end - sg.start);
    if (trace)
      fprintf(stderr, "[pgmm trace] query %d: %lld anchors in %zu segments (largest %lld), host refill %.1f ms (%lld anchors), backtrack %.1f ms, regs %.1f ms, "
              "plan %.1f ms (%zu hits, %zu jobs)\n", i, (long long)cj.n, cj.segs.size(), (long long)seg_max, t_redo, (long long)n_redo, t_bt, t_regs, c1 - c0,
              q.regs.size(), q.jobs.size());
  });
  be.stats.t_chain_rest += now() - tc1;
  be.end_chain();
  cpu_mark("backtrack + regs + plan");

  t1 = now(), be.stats.t_chain += t1 - t0, t0 = t1;
  // ---- DP waves ----
  // a hit with nothing pending after its first wave is finished in that wave instead of waiting for the next one
  // (hits only depend on their own DP results and on the finished hit in front of them); PGMM_EARLY_FINISH=0 disables
  const bool early_finish = getenv("PGMM_EARLY_FINISH") == nullptr || atoi(getenv("PGMM_EARLY_FINISH")) != 0;
  KswScoring sc;
  sc.sc_mch = (int8_t)(opt.a < 0 ? -opt.a : opt.a), sc.sc_mis = (int8_t)(opt.b > 0 ? -opt.b : opt.b), sc.sc_ambi = (int8_t)opt.sc_ambi;
  sc.q = (int8_t)opt.q, sc.e = (int8_t)opt.e, sc.q2 = (int8_t)opt.q2, sc.e2 = (int8_t)opt.e2;
  std::vector<KswJob> jobs;
  for (;;) {
    std::unique_ptr<CpuScope> cpu_wave(new CpuScope(5));
    auto res_mut = std::make_shared<KswBatchResult>();
    KswBatchResult &res = *res_mut;
    const std::shared_ptr<const KswBatchResult> res_sp = res_mut;
    jobs.clear();
    bool any_pending = false;
    for (QCtx &q : Q) {
      q.job_base = jobs.size();
      jobs.insert(jobs.end(), q.jobs.begin(), q.jobs.end());
      q.jobs.clear();
      q.done_wave = q.wave_id++;  // what was assembled so far runs now; new windows join the next wave
      any_pending |= q.pending;
      q.pending = false;
    }
    bool any_waiting = false;
    for (QCtx &q : Q)
      for (auto &R : q.regs) any_waiting |= R->state != Region::DONE;
    if (!any_waiting) break;
    t1 = now(), be.stats.t_stitch += t1 - t0, t0 = t1;
    cpu_mark("  wave: collect jobs");
    if (trace_cpu) {
      int64_t big = 0, n_big = 0;
      for (const KswJob &j : jobs) {
        big = std::max<int64_t>(big, (int64_t)j.qlen * j.tlen);
        n_big += (int64_t)j.qlen + j.tlen > 1500;
      }
      fprintf(stderr, "[pgmm trace]   wave: %zu jobs, %lld with more than 1500 anti-diagonals, largest %lld cells\n", jobs.size(), (long long)n_big,
              (long long)big);
    }
    if (!jobs.empty()) {
      be.run_dp(jobs, sc, res);
      cpu_mark("  wave: run_dp (host side)");
      t1 = now(), be.stats.t_dp += t1 - t0, t0 = t1;
      be.stats.jobs += jobs.size(), be.stats.cells += res.cells, be.stats.waves += 1;
      for (const KswJob &j : jobs) be.stats.seq_bytes += (uint64_t)j.qlen + j.tlen;
      be.stats.launches += res.launches, be.stats.kernel_ms += res.kernel_ms;
      for (int f = 0; f < 3; ++f)
        be.stats.fam_ms[f] += res.fam_ms[f], be.stats.fam_cells[f] += res.fam_cells[f], be.stats.fam_bases[f] += res.fam_bases[f],
            be.stats.fam_launches[f] += res.fam_launches[f];
    } else {
      res.out.clear(), res.cigar.clear(), res.cig_start.clear();
    }
    cpu_wave.reset();
    parallel_for(qb.n, n_threads, [&](int qi) {
      CpuScope cpu_scope(7);
      QCtx &q = Q[qi];
      if (!jobs.empty()) Mapper::publish_wave(q, res_sp);
      double s0 = now(), s1, tr_p1 = 0, tr_fin = 0, tr_plan = 0;
      for (int pass = 0; pass < 64; ++pass) {
      for (size_t k = 0; k < q.regs.size(); ++k) {
        Region &R = *q.regs[k];
        if (!Mapper::region_ready(q, R)) continue;  // waits for windows of the next wave
        switch (R.state) {
          case Region::WAIT1: {
            s0 = now();
            const bool second_wave = M.after_pass1(q, R, res_sp);
            tr_p1 += now() - s0;
            if (second_wave || !early_finish) break;
          }
          // fall through: nothing is pending for this hit, it is finished in the same wave
          case Region::WAIT2: {
            s0 = now();
            mm_reg1_t r2 = M.finish_region(q, R, res_sp);
            tr_fin += now() - s0;
            R.state = Region::DONE;
            if (r2.cnt > 0) {  // the split-off remainder is aligned next, right after its parent (align.c:1004)
              auto N = std::make_unique<Region>();
              N->r = r2;
              q.regs.insert(q.regs.begin() + k + 1, std::move(N));
            }
            // inversion rescue between the previous hit and this one (align.c:1005-1010)
            if (k > 0 && q.regs[k]->r.split_inv && !(opt.flag & MM_F_NO_INV)) {
              auto I = std::make_unique<Region>();
              if (M.plan_inversion(q, *q.regs[k - 1], *q.regs[k], *I)) {
                I->state = Region::WAIT_INV;
                q.regs.insert(q.regs.begin() + k + 1, std::move(I));
                ++k;  // the inversion slot sits between this hit and its remainder
              }
            }
            break;
          }
          case Region::WAIT_INV:
            if (M.finish_inversion(q, R, res_sp)) R.state = Region::DONE;
            else {
              q.regs.erase(q.regs.begin() + k);
              --k;
            }
            break;
          default:
            break;
        }
      }
      // newly created remainders plan their windows now; what the result cache cannot answer joins the next wave, and
      // so do the new windows of the remainders these hits are going to leave behind in turn
      s0 = now();
      bool go_on = false;
      for (size_t k = 0; k < q.regs.size(); ++k) {
        Region &R = *q.regs[k];
        if (R.state != Region::NEW) continue;
        M.plan_region(q, R);
        if (R.state != Region::WAIT1) continue;
        M.speculate_remainders(q, R);
        go_on |= M.dp_reuse && Mapper::region_ready(q, R);  // everything it asked for is known already
      }
      s1 = now(), tr_plan += s1 - s0;
      if (!go_on) break;
      }
      if (getenv("PGMM_TRACE") && tr_p1 + tr_fin + tr_plan > 1.0)
        fprintf(stderr, "[pgmm trace] query %d wave: test_zdrop+queue %.1f ms, finish %.1f ms, plan %.1f ms\n", qi, tr_p1, tr_fin, tr_plan);
    });
    cpu_mark("  wave: stitch + next plan");
  }
  be.end_batch();
  t1 = now(), be.stats.t_stitch += t1 - t0, t0 = t1;
  cpu_mark("  last wave: results");

  // ---- final filters, order, mapq, and the malloc()-owned result the boundary promises (minimap.h:353-366) ----
  parallel_for(qb.n, n_threads, [&](int qi) {
    CpuScope cpu_scope(8);
    QCtx &q = Q[qi];
    if (q.regs.empty()) return;
    M.filter_regs(q);
    if (!opt.split_prefix && q.qlen >= opt.rank_min_len) {
      M.update_dp_max(q);
      M.filter_regs(q);
    }
    M.hit_sort(q);
    M.set_mapq(q);
    const int n = (int)q.regs.size();
    // the reference returns realloc(regs, 0 bytes) for an empty list; callers only look at n_regs
    mm_reg1_t *out = (mm_reg1_t *)malloc(sizeof(mm_reg1_t) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; ++i) {
      Region &R = *q.regs[i];
      out[i] = R.r;
      out[i].p = nullptr;
      if (R.has_p) {
        const uint32_t cap = std::max<uint32_t>(R.capacity, (uint32_t)R.cigar.size() + 6);
        mm_extra_t *p = (mm_extra_t *)calloc(cap, 4);
        p->capacity = R.capacity, p->dp_score = R.dp_score, p->dp_max = R.dp_max, p->dp_max2 = R.dp_max2;
        p->n_ambi = R.n_ambi, p->trans_strand = 0, p->n_cigar = (uint32_t)R.cigar.size();
        memcpy(p->cigar, R.cigar.data(), R.cigar.size() * 4);
        out[i].p = p;
      }
    }
    n_regs[qi] = n, regs[qi] = out;
  });
  t1 = now(), be.stats.t_final += t1 - t0;
  cpu_mark("final");
}

}  // namespace pgmm

// replaces align.c:887-917 at the boundary (packages/minimap2/src/map.rs:323)
extern "C" double mm_event_identity(const mm_reg1_t *r) {
  if (r->p == nullptr) return -1.0f;
  int32_t n_gap = 0, n_gapo = 0;
  for (uint32_t i = 0; i < r->p->n_cigar; ++i) {
    const int32_t op = r->p->cigar[i] & 0xf, len = r->p->cigar[i] >> 4;
    if (op == MM_CIGAR_INS || op == MM_CIGAR_DEL) ++n_gapo, n_gap += len;
  }
  return (double)r->mlen / (r->blen + (int32_t)r->p->n_ambi - n_gap + n_gapo);
}
