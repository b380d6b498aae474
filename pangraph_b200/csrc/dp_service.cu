// Cross-round batching of the DP waves (dp_service.h).
#include "dp_service.h"

#include <algorithm>
#include <climits>
#include <condition_variable>
#include <cstdio>
#include <ctime>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>

#include "mapper.h"
#include "pgmm_cuda.h"

namespace pgmm {

namespace {
struct Request {
  std::vector<KswJob> jobs;   // this part's jobs (offsets relative to d_q / d_t)
  std::vector<int> where;     // their positions in the caller's job list
  const uint8_t *d_q, *d_t;
  KswScoring sc;
  std::shared_ptr<const KswBatchResult> merged;  // set by the worker
  size_t base = 0;                               // first job of this request inside the merged wave
  bool stats_owner = false, done = false;
};
}  // namespace

// Three lanes with their own workers.  A merged wave lasts as long as its longest problem, so problems are batched with
// their likes: SHORT (the thousands of small fills, < 1 ms each), MEDIUM (one CTA for a few ms: mid-size fills, end
// extensions that cannot run long) and LONG (a fill across an inversion, a 10 kbp end extension: one CTA for tens of
// milliseconds).  A round waits for all of its parts; a late wave that only carries a 3-ms problem no longer waits for
// somebody else's 40-ms one.
constexpr int kLanes = 3;
struct DpService::Impl {
  std::mutex mu;
  std::condition_variable cv_work[kLanes], cv_done;
  std::deque<Request *> lane[kLanes];
  std::vector<std::thread> workers;
  size_t arena_bytes[kLanes] = {0, 0, 0}, max_jobs = 0;

  void worker(int L) {
    std::deque<Request *> &pending = lane[L];
    require_device();
    cudaStream_t stream;
    {
      int prio_lo = 0, prio_hi = 0;  // copies and markers of a wave: urgent, tiny
      PGMM_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
      PGMM_CUDA(cudaStreamCreateWithPriority(&stream, cudaStreamNonBlocking, prio_hi));
    }
    KswEngine eng;
    eng.arena_budget_bytes = arena_bytes[L];
    std::vector<Request *> batch;
    std::vector<KswJob> merged_jobs;
    for (;;) {
      batch.clear();
      {
        std::unique_lock<std::mutex> g(mu);
        cv_work[L].wait(g, [&] { return !pending.empty(); });
        // everything pending that scores like the first request (rounds of one build all do)
        const KswScoring sc = pending.front()->sc;
        size_t n = 0;
        for (auto it = pending.begin(); it != pending.end();) {
          if (memcmp(&(*it)->sc, &sc, sizeof(sc)) == 0 && (batch.empty() || n + (*it)->jobs.size() <= max_jobs)) {
            n += (*it)->jobs.size();
            batch.push_back(*it);
            it = pending.erase(it);
          } else ++it;
        }
      }
      static const bool trace = getenv("PGMM_TRACE") != nullptr;
      CpuScope cpu_scope(6);
      timespec ts0, ts1, ts2;
      clock_gettime(CLOCK_MONOTONIC, &ts0);
      merged_jobs.clear();
      for (Request *r : batch) {
        r->base = merged_jobs.size();
        for (KswJob j : r->jobs) {  // absolute device addresses: the kernels add the offsets to a null base
          j.q_off += (uint64_t)(uintptr_t)r->d_q, j.t_off += (uint64_t)(uintptr_t)r->d_t;
          merged_jobs.push_back(j);
        }
      }
      auto res = std::make_shared<KswBatchResult>();
      clock_gettime(CLOCK_MONOTONIC, &ts1);
      eng.run(merged_jobs, nullptr, nullptr, batch.front()->sc, *res, stream);
      clock_gettime(CLOCK_MONOTONIC, &ts2);
      if (trace)
        fprintf(stderr, "[pgmm trace] dp batch (%s lane): %zu rounds, %zu jobs, merge %.1f ms, run %.1f ms (kernels %.1f ms, %d launches)\n", L == 0 ? "short" : L == 1 ? "medium" : "long", batch.size(),
                merged_jobs.size(), (ts1.tv_sec - ts0.tv_sec) * 1e3 + (ts1.tv_nsec - ts0.tv_nsec) * 1e-6,
                (ts2.tv_sec - ts1.tv_sec) * 1e3 + (ts2.tv_nsec - ts1.tv_nsec) * 1e-6, res->kernel_ms, res->launches);
      {
        std::lock_guard<std::mutex> g(mu);
        for (size_t k = 0; k < batch.size(); ++k) batch[k]->merged = res, batch[k]->stats_owner = k == 0, batch[k]->done = true;
      }
      cv_done.notify_all();
    }
  }
};

DpService::DpService() : impl_(new Impl) {
  // A batch lasts as long as its longest problem, and a worker runs one batch at a time: enough workers that a new
  // wave rarely waits for a running batch (each owns one stream per size class and one arena).
  const char *e = getenv("PGMM_DP_WORKERS");  // "<short>x<medium>x<long>"
  int n[kLanes] = {6, 8, 10};
  if (e) sscanf(e, "%dx%dx%d", &n[0], &n[1], &n[2]);
  for (int L = 0; L < kLanes; ++L) n[L] = std::max(1, n[L]);
  // Traceback arena budget per worker: PGMM_DP_ARENA_GB for the short lane, a third / half of it for the medium / long
  // lane -- scaled down so that all arenas together never exceed PGMM_DP_ARENA_FRAC (default 0.4) of the memory that is
  // free on the device when the service starts.  Arenas grow on demand up to their budget (KswEngine::run); a wave that
  // needs more is split.
  e = getenv("PGMM_DP_ARENA_GB");
  double gb = e ? atof(e) : 6.0;
  e = getenv("PGMM_DP_ARENA_FRAC");
  const double frac = e ? std::min(0.9, std::max(0.01, atof(e))) : 0.4;
  require_device();
  size_t free_b = 0, total_b = 0;
  PGMM_CUDA(cudaMemGetInfo(&free_b, &total_b));
  const double all_gb = gb * n[0] + gb / 3 * n[1] + gb / 2 * n[2], cap_gb = frac * (double)free_b / (double)(1ull << 30);
  if (all_gb > cap_gb) gb *= cap_gb / all_gb;
  if (gb * (double)(1ull << 30) / 3 < (double)(64u << 20))
    PGMM_FATAL("DP service: %.1f GB free on the device leave less than 64 MB of traceback arena per worker (budget = %.2f of free "
               "memory over %d+%d+%d workers; see PGMM_DP_ARENA_FRAC / PGMM_DP_WORKERS)", (double)free_b / (1ull << 30), frac, n[0], n[1], n[2]);
  impl_->arena_bytes[0] = (size_t)(gb * (double)(1ull << 30));
  impl_->arena_bytes[1] = (size_t)(gb / 3 * (double)(1ull << 30));
  impl_->arena_bytes[2] = (size_t)(gb / 2 * (double)(1ull << 30));
  if (getenv("PGMM_TRACE"))
    fprintf(stderr, "[pgmm trace] dp service: %.1f of %.1f GB free, arena budgets %.2f / %.2f / %.2f GB x %d / %d / %d workers\n",
            (double)free_b / (1ull << 30), (double)total_b / (1ull << 30), gb, gb / 3, gb / 2, n[0], n[1], n[2]);
  e = getenv("PGMM_DP_MAX_JOBS");
  impl_->max_jobs = e ? (size_t)atoll(e) : (size_t)300000;
  for (int L = 0; L < kLanes; ++L)
    for (int i = 0; i < n[L]; ++i) impl_->workers.emplace_back([this, L] { impl_->worker(L); });
  for (auto &t : impl_->workers) t.detach();
}

DpService &DpService::get() {
  static DpService *s = new DpService;  // lives as long as the process
  return *s;
}

bool DpService::enabled() {
  static const bool on = [] {
    const char *e = getenv("PGMM_DP_SERVICE");
    return e == nullptr || atoi(e) != 0;
  }();
  return on;
}

// Lane of a problem, from its shape alone.  Anything but a small fill occupies one CTA for its whole duration:
// ~0.65 ms per million cells when the band cannot bind (K5b), up to ~3 us per anti-diagonal for a band-limited end
// extension (K5), which stops early at a z-drop or runs through all qlen + tlen anti-diagonals.
static int lane_of(const KswJob &j) {
  const int64_t rows = (int64_t)j.qlen + j.tlen;
  const int64_t front = std::min<int64_t>(std::min(j.qlen, j.tlen), j.w < 0 ? INT32_MAX : (int64_t)j.w + 1);
  if (!(rows > 1500 && front > 300)) return 0;
  const bool banded = j.w >= 0 && j.w < std::max(j.qlen, j.tlen);
  const double worst_ms = banded ? (double)rows * 0.003 : (double)j.qlen * (double)j.tlen * 0.65e-6;
  return worst_ms <= 8.0 ? 1 : 2;
}

void DpService::run(const std::vector<KswJob> &jobs, const uint8_t *d_q, const uint8_t *d_t, const KswScoring &sc, KswBatchResult &res) {
  static const bool split = getenv("PGMM_DP_NO_SPLIT") == nullptr;
  Request part[kLanes];
  for (int L = 0; L < kLanes; ++L) part[L].d_q = d_q, part[L].d_t = d_t, part[L].sc = sc;
  const size_t n = jobs.size();
  for (size_t i = 0; i < n; ++i) {
    Request &r = part[split ? lane_of(jobs[i]) : 0];
    r.jobs.push_back(jobs[i]), r.where.push_back((int)i);
  }
  static const bool trace = getenv("PGMM_TRACE") != nullptr;
  timespec w0, w1;
  clock_gettime(CLOCK_MONOTONIC, &w0);
  {
    std::unique_lock<std::mutex> g(impl_->mu);
    for (int L = 0; L < kLanes; ++L) {
      if (part[L].jobs.empty()) part[L].done = true;
      else impl_->lane[L].push_back(&part[L]), impl_->cv_work[L].notify_one();
    }
    impl_->cv_done.wait(g, [&] { return part[0].done && part[1].done && part[2].done; });
  }
  if (trace) {
    clock_gettime(CLOCK_MONOTONIC, &w1);
    fprintf(stderr, "[pgmm trace] dp wave: %zu short, %zu medium, %zu long jobs, waited %.1f ms\n", part[0].jobs.size(), part[1].jobs.size(),
            part[2].jobs.size(), (w1.tv_sec - w0.tv_sec) * 1e3 + (w1.tv_nsec - w0.tv_nsec) * 1e-6);
  }
  // this round's slices of the merged waves, back in the caller's job order (copied by the round's own thread)
  res.out.assign(n, KswOut{});
  res.cig_start.assign(n + 1, 0);
  res.cells = 0, res.launches = 0, res.kernel_ms = 0.f;
  for (int f = 0; f < 3; ++f) res.fam_ms[f] = 0.f, res.fam_cells[f] = 0, res.fam_bases[f] = 0, res.fam_launches[f] = 0;
  std::vector<const uint32_t *> src(n, nullptr);
  size_t words = 0;
  for (int L = 0; L < kLanes; ++L) {
    const Request &r = part[L];
    if (!r.merged) continue;
    const KswBatchResult &m = *r.merged;
    for (size_t k = 0; k < r.where.size(); ++k) {
      const size_t i = (size_t)r.where[k];
      res.out[i] = m.out[r.base + k];
      src[i] = m.cigar.data() + m.cig_start[r.base + k];
      words += (size_t)res.out[i].n_cigar;
    }
    if (r.stats_owner) {  // a merged wave's counters are booked once
      res.cells += m.cells, res.launches += m.launches, res.kernel_ms += m.kernel_ms;
      for (int f = 0; f < 3; ++f)
        res.fam_ms[f] += m.fam_ms[f], res.fam_cells[f] += m.fam_cells[f], res.fam_bases[f] += m.fam_bases[f], res.fam_launches[f] += m.fam_launches[f];
    }
  }
  res.cigar.resize(words);
  size_t pos = 0;
  for (size_t i = 0; i < n; ++i) {
    const size_t nc = (size_t)res.out[i].n_cigar;
    res.cig_start[i] = pos;
    if (nc) memcpy(res.cigar.data() + pos, src[i], nc * sizeof(uint32_t));
    pos += nc;
  }
  res.cig_start[n] = pos;
}

}  // namespace pgmm
