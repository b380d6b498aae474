// Cross-round batching of the DP waves.
//
// A B200 executes kernels from at most 32 hardware work queues, and work that shares a queue is serialised: with one
// set of launches per round and wave, throughput grew linearly with CUDA_DEVICE_MAX_CONNECTIONS and stopped at 32
// (bench sweeps in DESIGN.md section 5) while most SMs idled.  The remedy is the one the reference's threading model
// suggests ("mm_map must enqueue into a shared batcher and block"): rounds hand their DP waves to this
// service, a few worker threads merge whatever is pending into ONE wave (one launch per size class for all rounds,
// one traceback arena instead of one per round) and hand every round its slice of the result.
#pragma once
#include <memory>
#include <vector>

#include "ksw_extd2.h"

namespace pgmm {

class DpService {
 public:
  static DpService &get();
  static bool enabled();  // PGMM_DP_SERVICE=0 keeps every round on its own engine
  // Same contract as KswEngine::run (blocking); q_off / t_off of the jobs are relative to d_q / d_t.
  void run(const std::vector<KswJob> &jobs, const uint8_t *d_q, const uint8_t *d_t, const KswScoring &sc, KswBatchResult &res);

 private:
  DpService();
  struct Impl;
  Impl *impl_;
};

}  // namespace pgmm
