// Cross-round batching of the chaining score fill (K4), the same idea as dp_service.h: a K4 launch lasts as long as
// its longest segment (~100-200 ms, one warp per segment) and, launched per round, holds one of the device's 32 hardware
// queues for that long.  Rounds hand their chaining jobs to this service; a few workers merge whatever is pending into
// one launch and hand every round its slice of f / p / v.
#pragma once
#include <memory>
#include <vector>

#include "chain.h"

namespace pgmm {

class ChainService {
 public:
  static ChainService &get();
  // Opt-in (PGMM_CHAIN_SERVICE=1).  Measured on one B200, 64 rounds in flight: the fill's latency per round doubles
  // (253 ms instead of 125: a launch now lasts as long as the longest segment of ANY of its rounds and a round first waits
  // for a free worker), the DP waves get faster by as much (438 ms instead of 640 per round) and the throughput stays
  // where it was (0.62 Gbp/s); device memory per worker grows with the batch.  Off by default until a round's stages
  // overlap better.
  static bool enabled();
  // Fills jobs[i].f/p/v/redo like ChainEngine::run (blocking).  The arrays live in *keep until the caller drops it.
  void run(const ChainParams &cp, std::vector<ChainFillJob> &jobs, std::shared_ptr<std::vector<int32_t>> &keep, ChainFillStats *stats);

 private:
  ChainService();
  struct Impl;
  Impl *impl_;
};

}  // namespace pgmm
