// Cross-round batching of the chaining score fill (chain_service.h).
#include "chain_service.h"

#include <algorithm>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <deque>
#include <mutex>
#include <thread>

#include "chain_fill.h"
#include "pgmm_cuda.h"

namespace pgmm {

// requests are merged when they chain with the same parameters (field by field: padding bytes carry no meaning)
static bool same_params(const ChainParams &a, const ChainParams &b) {
  return a.max_dist == b.max_dist && a.max_dist_inner == b.max_dist_inner && a.bw == b.bw && a.max_chn_skip == b.max_chn_skip &&
         a.cap_rmq_size == b.cap_rmq_size && a.min_cnt == b.min_cnt && a.min_sc == b.min_sc && a.pen_gap == b.pen_gap &&
         a.pen_skip == b.pen_skip;
}

namespace {
struct Request {
  ChainParams cp;
  std::vector<ChainFillJob> *jobs;
  std::shared_ptr<std::vector<int32_t>> store;
  ChainFillStats stats;
  bool done = false;
};
}  // namespace

struct ChainService::Impl {
  std::mutex mu;
  std::condition_variable cv_work, cv_done;
  std::deque<Request *> pending;
  std::vector<std::thread> workers;
  size_t max_anchors = 0;

  void worker() {
    require_device();
    cudaStream_t stream;
    PGMM_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    ChainEngine eng;
    std::vector<Request *> batch;
    std::vector<ChainFillJob> merged;
    std::vector<std::pair<int, int>> origin;  // (request, job) of every merged job
    static const bool trace = getenv("PGMM_TRACE") != nullptr;
    for (;;) {
      batch.clear(), merged.clear(), origin.clear();
      {
        std::unique_lock<std::mutex> g(mu);
        cv_work.wait(g, [&] { return !pending.empty(); });
        const ChainParams cp = pending.front()->cp;
        size_t n = 0;
        for (auto it = pending.begin(); it != pending.end();) {
          size_t na = 0;
          for (const ChainFillJob &j : *(*it)->jobs) na += (size_t)j.n;
          if (same_params((*it)->cp, cp) && (batch.empty() || n + na <= max_anchors)) {
            n += na;
            batch.push_back(*it);
            it = pending.erase(it);
          } else ++it;
        }
      }
      timespec t0, t1;
      clock_gettime(CLOCK_MONOTONIC, &t0);
      for (size_t r = 0; r < batch.size(); ++r)
        for (size_t k = 0; k < batch[r]->jobs->size(); ++k) {
          const ChainFillJob &j = (*batch[r]->jobs)[k];
          ChainFillJob m;
          m.a = j.a, m.n = j.n, m.segs = j.segs;
          merged.push_back(std::move(m));
          origin.emplace_back((int)r, (int)k);
        }
      ChainFillStats st;
      eng.run(batch.front()->cp, merged, stream, &st);
      // every request gets its own copy of its f / p / v (the engine's staging is reused by the next batch)
      for (size_t r = 0; r < batch.size(); ++r) {
        size_t na = 0;
        for (const ChainFillJob &j : *batch[r]->jobs) na += (size_t)j.n;
        batch[r]->store = std::make_shared<std::vector<int32_t>>(3 * na + 1);
      }
      std::vector<size_t> used(batch.size(), 0);
      for (size_t m = 0; m < merged.size(); ++m) {
        Request &rq = *batch[origin[m].first];
        ChainFillJob &dst = (*rq.jobs)[origin[m].second];
        int32_t *base = rq.store->data() + used[origin[m].first];
        const size_t n = (size_t)dst.n;
        dst.f = base, dst.p = base + n, dst.v = base + 2 * n;
        if (n) {
          memcpy(dst.f, merged[m].f, n * 4);
          memcpy(dst.p, merged[m].p, n * 4);
          memcpy(dst.v, merged[m].v, n * 4);
        }
        dst.redo = merged[m].redo;
        used[origin[m].first] += 3 * n;
      }
      if (trace) {
        clock_gettime(CLOCK_MONOTONIC, &t1);
        fprintf(stderr, "[pgmm trace] chain batch: %zu rounds, %llu anchors, %.1f ms (kernels %.1f ms)\n", batch.size(),
                (unsigned long long)st.anchors, (t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6, st.kernel_ms);
      }
      {
        std::lock_guard<std::mutex> g(mu);
        for (size_t r = 0; r < batch.size(); ++r) {
          batch[r]->stats = ChainFillStats();
          if (r == 0) batch[r]->stats = st;  // the merged launch is booked once
          batch[r]->done = true;
        }
      }
      cv_done.notify_all();
    }
  }
};

ChainService::ChainService() : impl_(new Impl) {
  const char *e = getenv("PGMM_CHAIN_MAX_ANCHORS");
  impl_->max_anchors = e ? (size_t)atoll(e) : (size_t)8 << 20;
  e = getenv("PGMM_CHAIN_WORKERS");
  const int n = e ? std::max(1, atoi(e)) : 3;
  for (int i = 0; i < n; ++i) impl_->workers.emplace_back([this] { impl_->worker(); });
  for (auto &t : impl_->workers) t.detach();
}

ChainService &ChainService::get() {
  static ChainService *s = new ChainService;  // lives as long as the process
  return *s;
}

bool ChainService::enabled() {
  static const bool on = [] {
    const char *e = getenv("PGMM_CHAIN_SERVICE");
    return e != nullptr && atoi(e) != 0;
  }();
  return on;
}

void ChainService::run(const ChainParams &cp, std::vector<ChainFillJob> &jobs, std::shared_ptr<std::vector<int32_t>> &keep, ChainFillStats *stats) {
  Request r;
  r.cp = cp, r.jobs = &jobs;
  size_t na = 0;
  for (const ChainFillJob &j : jobs) na += (size_t)j.n;
  if (na == 0) {
    for (ChainFillJob &j : jobs) j.redo.assign(j.segs.size(), 0);
    return;
  }
  {
    std::unique_lock<std::mutex> g(impl_->mu);
    impl_->pending.push_back(&r);
    impl_->cv_work.notify_one();
    impl_->cv_done.wait(g, [&] { return r.done; });
  }
  keep = r.store;
  if (stats) {
    stats->anchors += r.stats.anchors, stats->segments += r.stats.segments, stats->redo_segments += r.stats.redo_segments;
    stats->redo_anchors += r.stats.redo_anchors, stats->launches += r.stats.launches, stats->kernel_ms += r.stats.kernel_ms;
  }
}

}  // namespace pgmm
