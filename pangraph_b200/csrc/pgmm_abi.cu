// The C-ABI of libpgmm_b200.so, part 1 and 2 of include/pgmm_b200.h: the minimap2-sys boundary (index, map) backed by
// the CUDA stages, and the batched entry point.  Reference call sites: packages/minimap2/src/index.rs:40-47,
// map.rs:390, buf.rs:17,36; reference implementations replaced: minimap2/index.c:408-456, map.c:376-381.
#include "../../include/pgmm_b200.h"

#include <algorithm>
#include <cstring>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "chain_fill.h"
#include "chain_service.h"
#include "dp_service.h"
#include "ksw_extd2.h"
#include "mapper.h"
#include "pgmm_cuda.h"
#include "seeding.h"

using namespace pgmm;

namespace {

struct Stats {
  double seed_ms = 0, dp_kernel_ms = 0, total_ms = 0, index_ms = 0;
  uint64_t dp_seq_bytes = 0, dp_jobs = 0, dp_cells = 0, dp_waves = 0, bases_mapped = 0, bases_indexed = 0, batches = 0, launches = 0;
  double t_encode = 0, t_seed = 0, t_chain = 0, t_dp = 0, t_stitch = 0, t_final = 0;
  double fam_ms[3] = {0, 0, 0};
  uint64_t fam_cells[3] = {0, 0, 0}, fam_bases[3] = {0, 0, 0}, fam_launches[3] = {0, 0, 0};
  double t_chain_sort = 0, t_chain_fill = 0, t_chain_rest = 0, chain_kernel_ms = 0;
  uint64_t chain_anchors = 0, chain_segments = 0, chain_redo_segments = 0, chain_redo_anchors = 0, chain_launches = 0;
  uint64_t chain_iterations = 0, chain_batches = 0;
};
Stats g_stats;
std::mutex g_stats_mu;

double now_ms() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

// Everything that hangs off mm_idx_t::h: read-only once built, so any number of mapping calls may share it
struct PgmmIndex {
  TargetSet ts;
  DeviceIndex didx;
  DevBuf<uint8_t> d_tcodes;
  std::vector<std::string> sorted_names;  // distinct target names in strcmp order
  std::vector<int32_t> t_rank;
};

// One execution context = one CUDA stream + the engines' workspaces.  Calls check a context out for their duration;
// concurrent callers (the reference maps queries from a rayon pool, one mm_tbuf_t each) run on different streams.
struct DeviceCtx {
  SeedEngine seeder;
  ChainEngine chainer;
  std::unique_ptr<KswEngine> ksw;  // only when the DP service is off (its 29 streams and its arena are per engine)
  DeviceSeqSet qset;
  cudaStream_t stream = nullptr;
  DeviceCtx() {
    // PGMM_CTX_STREAMS=k: the contexts share k streams (a device runs kernels from at most 32 hardware queues and
    // streams beyond that alias onto them; every kernel of a context is short, so sharing a stream costs little)
    static const int shared = getenv("PGMM_CTX_STREAMS") ? atoi(getenv("PGMM_CTX_STREAMS")) : 0;
    // The kernels of the seeding and chaining stages are tiny and sit on a round's critical path (a chain fill is ~190
    // dependent launches of a few microseconds); the thousands of small-fill CTAs a DP wave queues are not.  When an SM
    // slot frees up the block scheduler serves the highest stream priority first, so these streams get the highest
    // (PGMM_CTX_PRIORITY=0: default priority, i.e. first come first served behind a wave's CTA backlog).
    int prio_lo = 0, prio_hi = 0;
    PGMM_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    static const bool high = getenv("PGMM_CTX_PRIORITY") == nullptr || atoi(getenv("PGMM_CTX_PRIORITY")) != 0;
    const int prio = high ? prio_hi : prio_lo;
    if (shared > 0) {
      static std::mutex mu;
      static std::vector<cudaStream_t> pool;
      static int next = 0;
      std::lock_guard<std::mutex> g(mu);
      if ((int)pool.size() < shared) {
        PGMM_CUDA(cudaStreamCreateWithPriority(&stream, cudaStreamNonBlocking, prio));
        pool.push_back(stream);
      } else stream = pool[next++ % shared];
    } else
    PGMM_CUDA(cudaStreamCreateWithPriority(&stream, cudaStreamNonBlocking, prio));
    if (!DpService::enabled()) {
      const char *e = getenv("PGMM_ARENA_GB");
      const double gb = e ? atof(e) : 8.0;
      ksw.reset(new KswEngine);
      ksw->arena_budget_bytes = (size_t)(gb * (1ull << 30));
    }
  }
};
class CtxPool {
 public:
  static CtxPool &get() {
    static CtxPool *p = new CtxPool;
    return *p;
  }
  DeviceCtx *acquire() {
    std::unique_lock<std::mutex> g(mu_);
    for (;;) {
      if (!idle_.empty()) {
        DeviceCtx *c = idle_.back();
        idle_.pop_back();
        return c;
      }
      if (n_ < max_) {
        ++n_;
        g.unlock();
        return new DeviceCtx;
      }
      cv_.wait(g);
    }
  }
  void release(DeviceCtx *c) {
    {
      std::lock_guard<std::mutex> g(mu_);
      idle_.push_back(c);
    }
    cv_.notify_one();
  }

 private:
  CtxPool() {
    const char *e = getenv("PGMM_CONTEXTS");
    max_ = e ? atoi(e) : 16;
    if (max_ < 1) max_ = 1;
  }
  std::mutex mu_;
  std::condition_variable cv_;
  std::vector<DeviceCtx *> idle_;
  int n_ = 0, max_ = 16;
};
struct CtxLease {
  DeviceCtx *c;
  CtxLease() : c(CtxPool::get().acquire()) {}
  ~CtxLease() { CtxPool::get().release(c); }
  DeviceCtx *operator->() { return c; }
};

int host_threads() {
  static int n = [] {
    const char *e = getenv("PGMM_THREADS");
    int v = e ? atoi(e) : (int)std::thread::hardware_concurrency();
    return v > 0 ? v : 1;
  }();
  return n;
}

// strcmp order without strings on the device: targets get odd ranks 2*i+1 (i = index among the distinct sorted target
// names); a query equal to a target shares its rank, any other query gets the even rank between its neighbours
int32_t rank_of(const std::vector<std::string> &sorted, const char *name) {
  const auto it = std::lower_bound(sorted.begin(), sorted.end(), name,
                                   [](const std::string &a, const char *b) { return strcmp(a.c_str(), b) < 0; });
  const int32_t i = (int32_t)(it - sorted.begin());
  if (it != sorted.end() && strcmp(it->c_str(), name) == 0) return 2 * i + 1;
  return 2 * i;
}

// builds the device query buffer (forward codes, then their reverse complement, per sequence) from resident target codes
__global__ void self_query_kernel(const uint8_t *__restrict__ tcodes, const uint64_t *__restrict__ t_off, const uint64_t *__restrict__ vstart,
                                  int n, uint64_t total, uint8_t *__restrict__ qcodes) {
  const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= total) return;
  int lo = 0, hi = n;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (vstart[mid] <= p) lo = mid;
    else hi = mid;
  }
  const uint64_t i = p - vstart[lo], L = vstart[lo + 1] - vstart[lo];
  const uint8_t c = tcodes[t_off[lo] + i];
  uint8_t *q = qcodes + 2 * vstart[lo];
  q[i] = c, q[L + (L - 1 - i)] = c < 4 ? 3 - c : 4;
}

// The seeding and chaining stages need an execution context (their workspaces are sized for a whole round: ~100 bytes
// per base); the DP waves -- most of a round's latency -- only need the round's query codes.  So a round holds a context
// from begin_batch to end_chain only and a dozen contexts serve several dozen rounds in flight.
struct CudaBackend : Backend {
  PgmmIndex &ix;
  std::unique_ptr<CtxLease> lease;
  DevBuf<uint8_t> d_qcodes;  // forward + reverse-complement codes of the round's queries (what the DP kernels read)
  CudaBackend(PgmmIndex &i) : ix(i) {}
  DeviceCtx &cx() {
    if (!lease) lease.reset(new CtxLease);
    return *lease->c;
  }
  double t_seed = 0;
  void begin_batch(const TargetSet &ts, const QueryBatch &qb) override {
    const size_t nbytes = qb.codes.size();
    d_qcodes.ensure(nbytes + 64);
    DeviceCtx &c = cx();
    if (qb.from_targets) {  // nothing crosses the bus: the query buffer is derived from the resident target codes
      const DeviceSeqSet &s = ix.didx.seqs;
      if (s.total > 0)
        self_query_kernel<<<(unsigned)((s.total + 255) / 256), 256, 0, c.stream>>>(ix.d_tcodes.p, s.starts.p, s.vstart.p, s.n, s.total, d_qcodes.p);
      PGMM_CUDA(cudaGetLastError());
      ++g_seed_launches;
      (void)ts;
    } else PGMM_CUDA(cudaMemcpyAsync(d_qcodes.p, qb.codes.data(), nbytes, cudaMemcpyHostToDevice, c.stream));
  }
  void seed_batch(const TargetSet &ts, const QueryBatch &qb, const mm_mapopt_t &opt, std::vector<QuerySeeds> &out) override {
    const double t0 = now_ms();
    DeviceCtx &c = cx();
    std::vector<uint64_t> starts(qb.base.begin(), qb.base.end());
    std::vector<int> lens(qb.lens.begin(), qb.lens.end());
    c.seeder.sketch(d_qcodes.p, starts, lens, ts.w, ts.k, c.qset, c.stream);
    std::vector<int32_t> q_rank(qb.n);
    // skip_seed only looks at names when the query has one (map.c:81): INT32_MIN marks "no name"
    for (int i = 0; i < qb.n; ++i) q_rank[i] = qb.names[i] ? rank_of(ix.sorted_names, qb.names[i]) : INT32_MIN;
    c.seeder.collect(ix.didx, c.qset, q_rank, opt, out, c.stream);
    t_seed += now_ms() - t0;
  }
  std::shared_ptr<std::vector<int32_t>> chain_keep;  // f / p / v of this round when they came from the chain service
  void chain_fill(const ChainParams &cp, std::vector<ChainFillJob> &jobs) override {
    if (ChainService::enabled()) ChainService::get().run(cp, jobs, chain_keep, &stats.chain);
    else {
      DeviceCtx &c = cx();
      c.chainer.run(cp, jobs, c.stream, &stats.chain);
    }
  }
  void end_chain() override {
    if (DpService::enabled()) lease.reset();  // without the service the context's own DP engine runs the waves
  }
  void run_dp(std::vector<KswJob> &jobs, const KswScoring &sc, KswBatchResult &res) override {
    if (!DpService::enabled()) {
      DeviceCtx &c = cx();
      c.ksw->run(jobs, d_qcodes.p, ix.d_tcodes.p, sc, res, c.stream);
    } else DpService::get().run(jobs, d_qcodes.p, ix.d_tcodes.p, sc, res);
  }
};

void map_with_index(const mm_idx_t *mi, int n, const int *lens, const char *const *seqs, const char *const *names,
                    const mm_mapopt_t *opt, int *n_regs, mm_reg1_t **regs) {
  require_device();
  PgmmIndex *ix = (PgmmIndex *)mi->h;
  if (!ix) PGMM_FATAL("mm_idx_t was not created by libpgmm_b200 (no device index attached)");
  const double t0 = now_ms();
  QueryBatch qb;
  std::vector<int> self_lens;
  std::vector<const char *> self_names;
  if (seqs == nullptr) {  // pgmm_map_self: the queries are the indexed sequences
    n = (int)ix->ts.lens.size();
    qb.from_targets = true;
    for (int i = 0; i < n; ++i) self_lens.push_back((int)ix->ts.lens[i]), self_names.push_back(mi->seq[i].name);
    lens = self_lens.data();
    qb.seqs.assign(n, nullptr);
    qb.names = self_names;
  } else {
    qb.seqs.assign(seqs, seqs + n);
    if (names) qb.names.assign(names, names + n);
    else qb.names.assign(n, nullptr);
  }
  qb.n = n;
  qb.lens.assign(lens, lens + n);
  CudaBackend be(*ix);
  map_batch(be, ix->ts, qb, *opt, n_regs, regs, host_threads());
  std::lock_guard<std::mutex> sl(g_stats_mu);
  g_stats.total_ms += now_ms() - t0, g_stats.seed_ms += be.t_seed, g_stats.dp_kernel_ms += be.stats.kernel_ms;
  g_stats.dp_jobs += be.stats.jobs, g_stats.dp_cells += be.stats.cells, g_stats.dp_waves += be.stats.waves;
  g_stats.launches += be.stats.launches, g_stats.batches += 1, g_stats.dp_seq_bytes += be.stats.seq_bytes;
  g_stats.t_encode += be.stats.t_encode, g_stats.t_seed += be.stats.t_seed, g_stats.t_chain += be.stats.t_chain;
  g_stats.t_dp += be.stats.t_dp, g_stats.t_stitch += be.stats.t_stitch, g_stats.t_final += be.stats.t_final;
  for (int f = 0; f < 3; ++f)
    g_stats.fam_ms[f] += be.stats.fam_ms[f], g_stats.fam_cells[f] += be.stats.fam_cells[f], g_stats.fam_bases[f] += be.stats.fam_bases[f],
        g_stats.fam_launches[f] += be.stats.fam_launches[f];
  g_stats.t_chain_sort += be.stats.t_chain_sort, g_stats.t_chain_fill += be.stats.t_chain_fill, g_stats.t_chain_rest += be.stats.t_chain_rest;
  g_stats.chain_kernel_ms += be.stats.chain.kernel_ms, g_stats.chain_anchors += be.stats.chain.anchors;
  g_stats.chain_segments += be.stats.chain.segments, g_stats.chain_redo_segments += be.stats.chain.redo_segments;
  g_stats.chain_redo_anchors += be.stats.chain.redo_anchors, g_stats.chain_launches += (uint64_t)be.stats.chain.launches;
  g_stats.chain_iterations += be.stats.chain.iterations, g_stats.chain_batches += be.stats.chain.batches;
  for (int i = 0; i < n; ++i) g_stats.bases_mapped += lens[i];
}

}  // namespace

extern "C" {

mm_idx_t *pgmm_idx_upload(int n, const char **seq, const char **name) {
  if (n <= 0) return nullptr;
  require_device();
  mm_idx_t *mi = (mm_idx_t *)calloc(1, sizeof(mm_idx_t));
  mi->flag = name == nullptr ? MM_I_NO_NAME : 0;
  mi->n_seq = (uint32_t)n;
  mi->seq = (mm_idx_seq_t *)calloc((size_t)n, sizeof(mm_idx_seq_t));
  PgmmIndex *ix = new PgmmIndex;
  mi->h = ix;
  CtxLease cx;
  TargetSet &ts = ix->ts;
  uint64_t sum = 0;
  for (int i = 0; i < n; ++i) {
    const size_t len = strlen(seq[i]);
    if (name && name[i]) mi->seq[i].name = strdup(name[i]);
    mi->seq[i].offset = sum, mi->seq[i].len = (uint32_t)len, mi->seq[i].is_alt = 0;
    ts.names.push_back(name && name[i] ? name[i] : "");
    ts.lens.push_back((uint32_t)len);
    ts.offs.push_back(sum);
    sum += len;
  }
  ts.codes.resize(sum + 64);
  {
    std::vector<std::thread> th;
    const int nt = std::min(host_threads(), n);
    for (int t = 0; t < nt; ++t)
      th.emplace_back([&, t]() {
        for (int i = t; i < n; i += nt) {
          uint8_t *d = ts.codes.data() + ts.offs[i];
          const uint8_t *s = (const uint8_t *)seq[i];
          for (uint32_t j = 0; j < ts.lens[i]; ++j) d[j] = kNt4[s[j]];
        }
      });
    for (auto &t : th) t.join();
  }
  ix->d_tcodes.ensure(sum + 64);
  PGMM_CUDA(cudaMemcpyAsync(ix->d_tcodes.p, ts.codes.data(), sum, cudaMemcpyHostToDevice, cx->stream));
  // name ranks for the all-vs-all skips
  ix->sorted_names = ts.names;
  std::sort(ix->sorted_names.begin(), ix->sorted_names.end(), [](const std::string &a, const std::string &b) { return strcmp(a.c_str(), b.c_str()) < 0; });
  ix->sorted_names.erase(std::unique(ix->sorted_names.begin(), ix->sorted_names.end()), ix->sorted_names.end());
  ix->t_rank.resize(n);
  for (int i = 0; i < n; ++i) ix->t_rank[i] = rank_of(ix->sorted_names, ts.names[i].c_str());
  PGMM_CUDA(cudaStreamSynchronize(cx->stream));
  return mi;
}

void pgmm_idx_build(mm_idx_t *mi, int w, int k, int bucket_bits) {
  require_device();
  PgmmIndex *ix = (PgmmIndex *)mi->h;
  if (!ix) PGMM_FATAL("mm_idx_t was not created by libpgmm_b200");
  CtxLease cx;
  CpuScope cpu_scope(9);
  const double t0 = now_ms();
  if (bucket_bits < 0) bucket_bits = 14;
  if (k * 2 < bucket_bits) bucket_bits = k * 2;
  if (w < 1) w = 1;
  mi->w = w, mi->k = k, mi->b = bucket_bits;
  TargetSet &ts = ix->ts;
  ts.k = k, ts.w = w;
  // K1 + K2
  std::vector<int> lens(ts.lens.begin(), ts.lens.end());
  ix->didx.w = w, ix->didx.k = k;
  cx->seeder.sketch(ix->d_tcodes.p, ts.offs, lens, w, k, ix->didx.seqs, cx->stream);
  cx->seeder.build_index(ix->didx, ts.lens, ix->t_rank, cx->stream);
  uint64_t sum = 0;
  for (uint32_t l : ts.lens) sum += l;
  std::lock_guard<std::mutex> sl(g_stats_mu);
  g_stats.index_ms += now_ms() - t0, g_stats.bases_indexed += sum;
}

mm_idx_t *mm_idx_str(int w, int k, int is_hpc, int bucket_bits, int n, const char **seq, const char **name) {
  if (n <= 0) return nullptr;
  if (is_hpc) PGMM_FATAL("homopolymer-compressed minimizers (MM_I_HPC) are outside pangraph's path; no implementation here");
  mm_idx_t *mi = pgmm_idx_upload(n, seq, name);
  pgmm_idx_build(mi, w, k, bucket_bits);
  return mi;
}

int pgmm_set_device(int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) return -1;
  if (bound_device() >= 0 && bound_device() != device) return -2;  // already computing on another device
  bound_device() = device;
  PGMM_CUDA(cudaSetDevice(device));
  return 0;
}

void pgmm_map_self(const mm_idx_t *mi, const mm_mapopt_t *opt, int *n_regs, mm_reg1_t **regs) {
  map_with_index(mi, 0, nullptr, nullptr, nullptr, opt, n_regs, regs);
}

void mm_mapopt_update(mm_mapopt_t *opt, const mm_idx_t *mi) {
  if ((opt->flag & MM_F_SPLICE_FOR) || (opt->flag & MM_F_SPLICE_REV)) opt->flag |= MM_F_SPLICE;
  if (opt->mid_occ <= 0) {
    require_device();
    PgmmIndex *ix = (PgmmIndex *)mi->h;
    if (!ix) PGMM_FATAL("mm_idx_t was not created by libpgmm_b200");
    CtxLease cx;
    opt->mid_occ = SeedEngine::cal_max_occ(ix->didx, opt->mid_occ_frac, cx->stream);
    if (opt->mid_occ < opt->min_mid_occ) opt->mid_occ = opt->min_mid_occ;
    if (opt->max_mid_occ > opt->min_mid_occ && opt->mid_occ > opt->max_mid_occ) opt->mid_occ = opt->max_mid_occ;
  }
  if (opt->bw_long < opt->bw) opt->bw_long = opt->bw;
}

void mm_idx_destroy(mm_idx_t *mi) {
  if (mi == nullptr) return;
  delete (PgmmIndex *)mi->h;
  if (mi->seq) {
    for (uint32_t i = 0; i < mi->n_seq; ++i) free(mi->seq[i].name);
    free(mi->seq);
  }
  free(mi);
}

struct mm_tbuf_s {
  void *km;
  int rep_len, frag_gap;
};

mm_tbuf_t *mm_tbuf_init(void) { return (mm_tbuf_t *)calloc(1, sizeof(mm_tbuf_s)); }
void mm_tbuf_destroy(mm_tbuf_t *b) { free(b); }

mm_reg1_t *mm_map(const mm_idx_t *mi, int l_seq, const char *seq, int *n_regs, mm_tbuf_t *, const mm_mapopt_t *opt, const char *name) {
  mm_reg1_t *regs = nullptr;
  const char *names[1] = {name};
  map_with_index(mi, 1, &l_seq, &seq, name ? names : nullptr, opt, n_regs, &regs);
  return regs;
}

void pgmm_map_batch(const mm_idx_t *mi, int n, const int *lens, const char *const *seqs, const char *const *names, const mm_mapopt_t *opt,
                    int *n_regs, mm_reg1_t **regs) {
  map_with_index(mi, n, lens, seqs, names, opt, n_regs, regs);
}

// stats: [0] total_ms [1] seed_ms [2] dp_kernel_ms [3] index_ms [4] dp_jobs [5] dp_cells [6] dp_waves [7] bases_mapped
//        [8] bases_indexed [9] batches [10] kernel launches (all engines) ... [33..41] chaining: sort ms, fill ms, rest ms,
//        fill kernel ms, anchors, segments, segments redone on the host, their anchors, launches
//        [42] fixed-point iterations of the chain fill summed over its batches [43] those batches
//        [44..53] host CPU ms by phase (mapper.h: cpu_phase_*)
//        [54] queries whose anchors came back from the device sorted (no equal target positions), [55] queries sorted by the host replay
void pgmm_get_stats(double *out, int n, int reset) {
  std::lock_guard<std::mutex> sl(g_stats_mu);
  double cpu[kCpuPhases];
  cpu_phase_read(cpu, reset != 0);
  const double v[46 + kCpuPhases] = {g_stats.total_ms, g_stats.seed_ms, g_stats.dp_kernel_ms, g_stats.index_ms, (double)g_stats.dp_jobs,
                        (double)g_stats.dp_cells, (double)g_stats.dp_waves, (double)g_stats.bases_mapped,
                        (double)g_stats.bases_indexed, (double)g_stats.batches, (double)g_stats.launches + (double)pgmm::g_seed_launches + (double)g_stats.chain_launches,
                        g_stats.t_encode, g_stats.t_seed, g_stats.t_chain, g_stats.t_dp, g_stats.t_stitch, g_stats.t_final,
                        (double)pgmm::h2d_bytes(), (double)pgmm::d2h_bytes(), (double)g_stats.dp_seq_bytes,
                        (double)pgmm::DevicePool::misses(),
                        g_stats.fam_ms[0], (double)g_stats.fam_cells[0], (double)g_stats.fam_bases[0], (double)g_stats.fam_launches[0],
                        g_stats.fam_ms[1], (double)g_stats.fam_cells[1], (double)g_stats.fam_bases[1], (double)g_stats.fam_launches[1],
                        g_stats.fam_ms[2], (double)g_stats.fam_cells[2], (double)g_stats.fam_bases[2], (double)g_stats.fam_launches[2],
                        g_stats.t_chain_sort, g_stats.t_chain_fill, g_stats.t_chain_rest, g_stats.chain_kernel_ms,
                        (double)g_stats.chain_anchors, (double)g_stats.chain_segments, (double)g_stats.chain_redo_segments,
                        (double)g_stats.chain_redo_anchors, (double)g_stats.chain_launches,
                        (double)g_stats.chain_iterations, (double)g_stats.chain_batches,
                        cpu[0], cpu[1], cpu[2], cpu[3], cpu[4], cpu[5], cpu[6], cpu[7], cpu[8], cpu[9],
                        (double)pgmm::g_anchor_sorted_device.load(), (double)pgmm::g_anchor_sorted_host.load()};
  for (int i = 0; i < n && i < 46 + kCpuPhases; ++i) out[i] = v[i];
  if (reset) pgmm::g_seed_launches = 0, pgmm::h2d_bytes() = 0, pgmm::d2h_bytes() = 0, pgmm::DevicePool::misses() = 0;
  if (reset) pgmm::g_anchor_sorted_device = 0, pgmm::g_anchor_sorted_host = 0;
  if (reset) g_stats = Stats();
}


// ---- stage-level entry points (include/pgmm_b200.h part 3) ----

// K1 alone: minimizers of n ASCII sequences. out_x/out_y need capacity for all of them (cap entries); out_off[n+1].
int pgmm_sketch(int n, const char *const *seqs, const int *lens, int w, int k, uint64_t *out_x, uint64_t *out_y, uint64_t cap,
                uint64_t *out_off) {
  require_device();
  std::vector<uint64_t> starts(n);
  std::vector<int> l(lens, lens + n);
  uint64_t sum = 0;
  for (int i = 0; i < n; ++i) starts[i] = sum, sum += (uint64_t)lens[i];
  std::vector<uint8_t> codes(sum + 64);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < lens[i]; ++j) codes[starts[i] + j] = kNt4[(uint8_t)seqs[i][j]];
  cudaStream_t st;
  PGMM_CUDA(cudaStreamCreate(&st));
  DevBuf<uint8_t> d;
  d.ensure(sum + 64);
  PGMM_CUDA(cudaMemcpyAsync(d.p, codes.data(), sum, cudaMemcpyHostToDevice, st));
  SeedEngine eng;
  DeviceSeqSet set;
  eng.sketch(d.p, starts, l, w, k, set, st);
  int rc = 0;
  if (set.n_mz > cap) rc = -1;
  else if (set.n_mz > 0) {
    PGMM_CUDA(cudaMemcpyAsync(out_x, set.mx.p, set.n_mz * 8, cudaMemcpyDeviceToHost, st));
    PGMM_CUDA(cudaMemcpyAsync(out_y, set.my.p, set.n_mz * 8, cudaMemcpyDeviceToHost, st));
  }
  PGMM_CUDA(cudaStreamSynchronize(st));
  for (int i = 0; i <= n; ++i) out_off[i] = set.h_mz_off[i];
  PGMM_CUDA(cudaStreamDestroy(st));
  return rc;
}

// K1+K3 alone: the anchors (before the host's sort), used-seed positions and repeat length of every query against mi.
// out_anchor: 2 uint64 per anchor; out_n[3*i..] = {n_anchors, n_mini_pos, rep_len} of query i.
int pgmm_collect_seeds(const mm_idx_t *mi, int n, const int *lens, const char *const *seqs, const char *const *names,
                       const mm_mapopt_t *opt, uint64_t *out_anchor, uint64_t anchor_cap, uint64_t *out_mini, uint64_t mini_cap,
                       int64_t *out_n) {
  require_device();
  PgmmIndex *ix = mi ? (PgmmIndex *)mi->h : nullptr;
  if (!ix) PGMM_FATAL("mm_idx_t was not created by libpgmm_b200 (no device index attached)");
  QueryBatch qb;
  qb.n = n;
  qb.seqs.assign(seqs, seqs + n);
  if (names) qb.names.assign(names, names + n);
  else qb.names.assign(n, nullptr);
  qb.lens.assign(lens, lens + n);
  encode_queries(qb, ix->ts, host_threads());
  CudaBackend be(*ix);
  be.begin_batch(ix->ts, qb);
  std::vector<QuerySeeds> out;
  be.seed_batch(ix->ts, qb, *opt, out);
  uint64_t na = 0, nm = 0;
  for (int i = 0; i < n; ++i) {
    if (na + out[i].a.size() > anchor_cap || nm + out[i].mini_pos.size() > mini_cap) return -1;
    memcpy(out_anchor + 2 * na, out[i].a.data(), out[i].a.size() * 16);
    memcpy(out_mini + nm, out[i].mini_pos.data(), out[i].mini_pos.size() * 8);
    out_n[3 * i] = (int64_t)out[i].a.size(), out_n[3 * i + 1] = (int64_t)out[i].mini_pos.size(), out_n[3 * i + 2] = out[i].rep_len;
    na += out[i].a.size(), nm += out[i].mini_pos.size();
  }
  return 0;
}

}  // extern "C"
