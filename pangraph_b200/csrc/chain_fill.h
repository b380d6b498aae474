// K4: the score fill of RMQ chaining on the GPU (reference: mg_lchain_rmq's main loop, minimap2/lchain.c:276-358).
#pragma once
#include <cstdint>
#include <vector>

#include <cuda_runtime.h>

#include "chain.h"
#include "pgmm_cuda.h"

namespace pgmm {

class ChainEngine {
 public:
  // limits beyond which a segment is handed back to the host (chain_fill_host)
  // anchors between the oldest window member and the current one.  Synthetic 1 %-divergent genomes need ~800 (10 kbp of
  // target at one anchor per ~13 bp); a real E. coli pair reaches 2835 in its repeats (18 % of its anchors sit in segments
  // that overflow a ring of 2048; measured with the oracle on the CPU), hence 4096 (105 KB of shared memory per warp).
  static constexpr int kRing = 4096;
  static constexpr int kInnerCap = 1024;    // candidates of one near-neighbourhood walk
  enum Redo : uint8_t { DONE = 0, TIE = 1, WINDOW = 2, INNER = 3 };

  ChainEngine();
  ~ChainEngine();
  // uploads the anchors of all jobs, fills every segment, downloads f/p/v and the per-segment redo flags
  void run(const ChainParams &cp, std::vector<ChainFillJob> &jobs, cudaStream_t st, ChainFillStats *stats = nullptr);

 private:
  DevBuf<U128> d_a_;
  DevBuf<int32_t> d_x_, d_y_, d_f_, d_flag_;
  DevBuf<uint8_t> d_qs_;
  DevBuf<int4> d_segs_, d_aux_;
  DevBuf<int> d_par_;      // state of the parallel fixed-point fill (K4p)
  PinBuf<int> h_changed_;
  PinBuf<U128> h_a_;
  PinBuf<int32_t> h_fpv_, h_flag_;
  PinBuf<int4> h_segs_;
  cudaEvent_t ev0_ = nullptr, ev1_ = nullptr;
};

}  // namespace pgmm
