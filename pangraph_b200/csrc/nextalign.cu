// K6: map_variations on the GPU -- the banded affine alignment (i32 scores, i8 paths) pangraph runs for every
// (block, node) pair after a merge (PG/pangraph/reweave.rs:40-94 -> PG/align/map_variations.rs:39-80 ->
// PG/align/nextclade/align/align.rs:33-75, score_matrix.rs, backtrace.rs).  Embarrassingly parallel over
// problems: one warp per problem, a batch of problems per launch.
//
// Wavefront inside a problem.  In band coordinates (nextalign_core.h) a cell (ri, K) reads (ri-1, K), (ri, K-1) and
// (ri-1, K+1).  Lane j owns the m = max(2, ceil(W/32)) band columns [j m, j m + m) and, in macro step s, walks them left to
// right for row ri = s - j: its left neighbour finished the same row one macro step earlier (carry by shuffle), and the first
// column of its right neighbour -- which is one row behind -- is written in micro step 0 of the same macro step and read in
// micro step m-1.  Every lane is busy in every micro step (no idle half like a plain anti-diagonal sweep), the row state
// (score and query-gap score per band column) lives in shared memory and is updated in place, paths stream to HBM as one
// byte per band cell.  The cell update itself is the reference's, statement by statement (na::cell).
// The cells outside the band (forced stripe ends, clamped stripes) are chains that one lane walks after the wavefront; the
// traceback is walked by the whole warp (32 cells of the diagonal per step while the path stays on it).
#include "nextalign.h"

#include "nextalign_core.h"
#include "pgmm_cuda.h"

#include <algorithm>
#include <cstring>

namespace pgmm {
namespace na {

namespace {

struct Job {
  int32_t rlen, qlen, ms, bw;
  uint64_t r_off, q_off;     // codes
  uint64_t band_off, edge_off, tail_off, runs_off;
};
struct Out {
  int32_t score, hit_boundary, status;
  int32_t n_runs;
};

__global__ void __launch_bounds__(32) nextalign_kernel(const Job *__restrict__ jobs, const uint8_t *__restrict__ codes, Params p,
                                                       uint8_t *__restrict__ band, uint8_t *__restrict__ edge, uint8_t *__restrict__ tail,
                                                       uint32_t *__restrict__ runs, Out *__restrict__ outs) {
  extern __shared__ int32_t na_smem[];
  const Job job = jobs[blockIdx.x];
  const int lane = threadIdx.x;
  Geom g;
  g.rlen = job.rlen, g.qlen = job.qlen, g.ms = job.ms, g.bw = job.bw, g.W = 2 * job.bw + 1;
  const int W = g.W, m = max(2, (W + 31) / 32);
  int32_t *S = na_smem, *QG = na_smem + (W + 2);
  const uint8_t *R = codes + job.r_off, *Q = codes + job.q_off;
  uint8_t *B = band + job.band_off, *E = edge + job.edge_off, *T = tail + job.tail_off;
  uint32_t *RUNS = runs + job.runs_off;
  const unsigned full = 0xffffffffu;

  // ---- wavefront over the in-band cells ----
  int32_t lastS = 0, lastRG = kNoAlign;  // score / running reference-gap score after this lane's last cell of its previous row
  const int lanes_used = (W + m - 1) / m, n_macro = g.rlen + lanes_used;
  for (int s = 0; s < n_macro; ++s) {
    const int ri = s - lane;
    const int32_t inS = __shfl_up_sync(full, lastS, 1);
    int32_t rg = __shfl_up_sync(full, lastRG, 1);
    if (lane == 0) rg = kNoAlign;  // a row starts with ref_gaps = NO_ALIGN (score_matrix.rs:86)
    int32_t curS = inS;
    const bool row_ok = ri >= 0 && ri <= g.rlen;
    const int q0 = g.b(ri) + lane * m;  // query column of this lane's first band column
    const int beg = row_ok ? g.begin(ri) : 0, en = row_ok ? g.end(ri) : 0;
    const int rc = (row_ok && ri > 0) ? R[ri - 1] : 0;
    const RowGeom rw = (row_ok && ri > 0) ? row_geom(g, ri) : RowGeom{0, 0, 0, 0};
    for (int c = 0; c < m; ++c) {
      const int K = lane * m + c, qpos = q0 + c;
      if (row_ok && K < W && qpos >= beg && qpos < en) {
        CellOut o;
        if (ri == 0) o.S = row0_score(p, qpos), o.path = row0_path(qpos), o.qry_gaps = kNoAlign, o.ref_gaps = rg;
        else {
          CellIn in;
          in.diagS = S[K], in.leftS = curS, in.ref_gaps = rg, in.upS = S[K + 1], in.qry_gaps = QG[K + 1];
          in.qc = qpos > 0 ? Q[qpos - 1] : 0, in.rc = rc;
          o = cell(g, p, rw, ri, qpos, in);
        }
        S[K] = o.S, QG[K] = o.qry_gaps, rg = o.ref_gaps, curS = o.S;
        B[(int64_t)ri * W + K] = (uint8_t)o.path;
      }
      __syncwarp();
    }
    if (row_ok) lastS = curS, lastRG = rg;
  }
  __syncwarp();

  // ---- the two chains outside the band and the final score (lane 0 computes, all lanes learn the score) ----
  int32_t final_score = 0;
  {
    // last in-band cell of the last row: start of the tail along the last row
    const int t0 = g.end_unforced(g.rlen);         // first query column right of the last row's own stripe
    const int K_last = t0 - 1 - g.b(g.rlen);       // band column of (rlen, t0 - 1); >= W when the band lies left of the matrix
    const int owner = (K_last >= 0 && K_last < W) ? K_last / m : 0;
    const int32_t ownS = __shfl_sync(full, lastS, owner), ownRG = __shfl_sync(full, lastRG, owner);
    if (lane == 0) {
      if (g.rlen == 0) final_score = row0_score(p, g.qlen);
      else {
        bool have = false;
        // (1) chain down the right edge: rows whose band lies right of the matrix, b(ri) > qlen
        const int i1 = g.qlen + g.ms + g.bw;  // b(i1) == qlen
        if (i1 < g.rlen) {
          int32_t upS, qg;
          if (i1 >= 0) upS = S[0], qg = QG[0];                  // (i1, qlen) is the in-band cell K = 0 (row 0: its closed form)
          else upS = row0_score(p, g.qlen), qg = kNoAlign;      // row 0 spans the whole matrix (forced begin)
          for (int ri = max(i1, 0) + 1; ri <= g.rlen; ++ri) {
            CellIn in;
            in.diagS = row0_score(p, g.qlen - 1);  // only read by (1, qlen) under the forced full row 0
            in.leftS = 0, in.ref_gaps = kNoAlign, in.upS = upS, in.qry_gaps = qg;
            in.qc = g.qlen > 0 ? Q[g.qlen - 1] : 0, in.rc = R[ri - 1];
            const CellOut o = cell(g, p, row_geom(g, ri), ri, g.qlen, in);
            E[ri] = (uint8_t)o.path, upS = o.S, qg = o.qry_gaps;
          }
          final_score = upS, have = true;
        }
        // (2) chain along the last row right of its stripe (forced end)
        if (!have && t0 <= g.qlen) {
          int32_t leftS, rgc;
          if (K_last >= 0 && K_last < W) leftS = ownS, rgc = ownRG;
          else leftS = col0_score(p, g.rlen), rgc = kNoAlign;   // the stripe was clamped to the single cell (rlen, 0)
          const RowGeom rw_last = row_geom(g, g.rlen);
          for (int qpos = t0; qpos <= g.qlen; ++qpos) {
            CellIn in;
            in.diagS = col0_score(p, g.rlen - 1);  // only read for qpos == 1 below a stripe clamped to column 0
            in.leftS = leftS, in.ref_gaps = rgc, in.upS = 0, in.qry_gaps = kNoAlign;
            in.qc = Q[qpos - 1], in.rc = R[g.rlen - 1];
            const CellOut o = cell(g, p, rw_last, g.rlen, qpos, in);
            T[qpos] = (uint8_t)o.path, leftS = o.S, rgc = o.ref_gaps;
          }
          final_score = leftS, have = true;
        }
        if (!have) final_score = S[g.qlen - g.b(g.rlen)];
      }
    }
  }
  __threadfence_block();
  __syncwarp();

  // ---- traceback (backtrace.rs:37-85), whole warp ----
  const auto pa = [&](int ri, int qpos) {
    return path_at(g, ri, qpos, [&](int64_t i) { return (int)__ldcg(B + i); }, [&](int r) { return (int)__ldcg(E + r); },
                   [&](int q) { return (int)__ldcg(T + q); });
  };
  int r_pos = g.rlen, q_pos = g.qlen, current = 0, hb = 0, status = 0, n = 0;
  uint32_t last = 0;
  const auto push = [&](uint32_t op, uint32_t len) {
    if (last != 0 && (last & 3u) == op) last += len << 2;
    else {
      if (last != 0 && lane == 0) RUNS[n] = last;
      n += last != 0;
      last = len << 2 | op;
    }
  };
  while (r_pos > 0 || q_pos > 0) {
    if (current == 0) {
      const int rr = r_pos - lane, qq = q_pos - lane;
      const bool valid = rr >= 0 && qq >= 0 && (rr > 0 || qq > 0) && g.exists(rr, qq);
      const int o = valid ? pa(rr, qq) : 0;
      const unsigned stop = __ballot_sync(full, !(o & kMatch)), bnd = __ballot_sync(full, (o & kBoundary) != 0);
      const int run = stop ? __ffs(stop) - 1 : 32;  // leading cells that continue the diagonal
      if (run > 0) {
        if (bnd & (run == 32 ? full : ((1u << run) - 1u))) hb = 1;
        push(0, (uint32_t)run), r_pos -= run, q_pos -= run;
      }
      if (run == 32 || !(r_pos > 0 || q_pos > 0)) continue;
      if (!__shfl_sync(full, (int)valid, run)) {  // the walk left the band: cannot happen for paths the fill produced
        status = -3;
        break;
      }
      const int oo = __shfl_sync(full, o, run);
      if (oo & kBoundary) hb = 1;
      const int op = walk_step(oo, current, r_pos, q_pos);
      if (op < 0) {
        status = -2;
        break;
      }
      push((uint32_t)op, 1);
    } else {
      if (!g.exists(r_pos, q_pos)) {
        status = -3;
        break;
      }
      const int oo = pa(r_pos, q_pos);
      if (oo & kBoundary) hb = 1;
      const int op = walk_step(oo, current, r_pos, q_pos);
      if (op < 0) {
        status = -2;
        break;
      }
      push((uint32_t)op, 1);
    }
  }
  if (last != 0) {
    if (lane == 0) RUNS[n] = last;
    ++n;
  }
  if (lane == 0) {
    Out o;
    o.score = final_score, o.hit_boundary = hb, o.status = status, o.n_runs = n;
    outs[blockIdx.x] = o;
  }
}

}  // namespace

constexpr int kMaxW = 24000;  // 2 x (W + 2) x 4 bytes of shared memory per problem

void run_batch(const std::vector<Problem> &probs, int extra_band_width, int max_attempts, std::vector<Edit> &edits, Stats *stats) {
  require_device();
  const size_t n = probs.size();
  edits.assign(n, Edit{});
  if (n == 0) return;
  static const bool once = [] {
    PGMM_CUDA(cudaFuncSetAttribute(nextalign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    return true;
  }();
  (void)once;
  Params p;  // NextalignParams::default() with map_variations' overrides (params.rs:143-175)
  p.ext = 0, p.gopen = 6, p.mismatch = 1, p.match = 3, p.left_free = 1, p.right_free = 1, p.left_align = 1;

  // codes of every sequence, one buffer
  std::vector<uint64_t> r_off(n), q_off(n);
  uint64_t total = 0;
  for (size_t i = 0; i < n; ++i) r_off[i] = total, total += (uint64_t)probs[i].rlen, q_off[i] = total, total += (uint64_t)probs[i].qlen;
  PinBuf<uint8_t> h_codes;
  uint8_t *hc = h_codes.ensure(total + 16);
  std::vector<int> bw(n), attempt(n, 1);
  std::vector<size_t> active;
  for (size_t i = 0; i < n; ++i) {
    const Problem &pr = probs[i];
    // align_nuc_simplestripe rejects a query shorter than min_length (1); to_nuc_seq rejects unknown characters
    if (pr.qlen < 1 || !encode(pr.ref, pr.rlen, hc + r_off[i]) || !encode(pr.qry, pr.qlen, hc + q_off[i])) {
      edits[i].status = -1;
      continue;
    }
    bw[i] = pr.band_width + extra_band_width;
    active.push_back(i);
  }
  cudaStream_t st;
  PGMM_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  DevBuf<uint8_t> d_codes, d_band, d_edge, d_tail;
  DevBuf<uint32_t> d_runs;
  DevBuf<Job> d_jobs;
  DevBuf<Out> d_outs;
  PinBuf<Job> h_jobs;
  PinBuf<Out> h_outs;
  PinBuf<uint32_t> h_runs;
  d_codes.ensure(total + 16);
  PGMM_CUDA(cudaMemcpyAsync(d_codes.p, hc, total, cudaMemcpyHostToDevice, st));
  cudaEvent_t ev0, ev1;
  PGMM_CUDA(cudaEventCreate(&ev0));
  PGMM_CUDA(cudaEventCreate(&ev1));
  size_t free_b = 0, total_b = 0;
  PGMM_CUDA(cudaMemGetInfo(&free_b, &total_b));
  const uint64_t budget = std::max<uint64_t>((uint64_t)1 << 28, (uint64_t)(free_b * 0.5));  // path bytes of one launch

  while (!active.empty()) {
    std::vector<size_t> next;
    size_t pos = 0;
    while (pos < active.size()) {
      // one launch: as many problems as fit the path budget
      uint64_t band_b = 0, edge_b = 0, tail_b = 0, runs_w = 0;
      size_t end = pos, smem = 0;
      Job *hj = h_jobs.ensure(active.size() - pos);
      while (end < active.size()) {
        const size_t i = active[end];
        const Problem &pr = probs[i];
        const int64_t W = 2 * (int64_t)bw[i] + 1;
        if (W > kMaxW) {  // wider than the kernel's row state: report instead of guessing (no fallback)
          edits[i].status = -4, edits[i].attempts = attempt[i], edits[i].band_width = bw[i];
          active.erase(active.begin() + (long)end);
          continue;
        }
        const uint64_t bb = ((uint64_t)(pr.rlen + 1) * (uint64_t)W + 255) / 256 * 256;
        if (end > pos && band_b + bb > budget) break;
        Job &j = hj[end - pos];
        j.rlen = pr.rlen, j.qlen = pr.qlen, j.ms = pr.mean_shift, j.bw = bw[i];
        j.r_off = r_off[i], j.q_off = q_off[i];
        j.band_off = band_b, band_b += bb;
        j.edge_off = edge_b, edge_b += ((uint64_t)pr.rlen + 1 + 255) / 256 * 256;
        j.tail_off = tail_b, tail_b += ((uint64_t)pr.qlen + 1 + 255) / 256 * 256;
        j.runs_off = runs_w, runs_w += (uint64_t)pr.rlen + pr.qlen + 2;
        smem = std::max(smem, (size_t)(2 * (W + 2) * 4));
        ++end;
      }
      const size_t nl = end - pos;
      if (nl == 0) break;
      d_band.ensure(band_b + 256), d_edge.ensure(edge_b + 256), d_tail.ensure(tail_b + 256), d_runs.ensure(runs_w + 16);
      d_jobs.ensure(nl), d_outs.ensure(nl);
      PGMM_CUDA(cudaMemcpyAsync(d_jobs.p, hj, nl * sizeof(Job), cudaMemcpyHostToDevice, st));
      PGMM_CUDA(cudaEventRecord(ev0, st));
      nextalign_kernel<<<(unsigned)nl, 32, smem, st>>>(d_jobs.p, d_codes.p, p, d_band.p, d_edge.p, d_tail.p, d_runs.p, d_outs.p);
      PGMM_CUDA(cudaGetLastError());
      PGMM_CUDA(cudaEventRecord(ev1, st));
      Out *ho = h_outs.ensure(nl);
      uint32_t *hr = h_runs.ensure(runs_w + 16);
      PGMM_CUDA(cudaMemcpyAsync(ho, d_outs.p, nl * sizeof(Out), cudaMemcpyDeviceToHost, st));
      PGMM_CUDA(cudaMemcpyAsync(hr, d_runs.p, runs_w * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
      PGMM_CUDA(cudaStreamSynchronize(st));
      if (stats) {
        float ms = 0;
        PGMM_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
        stats->kernel_ms += ms, stats->launches += 1;
      }
      for (size_t k = 0; k < nl; ++k) {
        const size_t i = active[pos + k];
        const Problem &pr = probs[i];
        if (stats) stats->cells += (uint64_t)(pr.rlen + 1) * (uint64_t)(2 * bw[i] + 1), stats->problems += 1;
        Edit &e = edits[i];
        e.status = ho[k].status, e.hit_boundary = ho[k].hit_boundary, e.attempts = attempt[i], e.band_width = bw[i], e.score = ho[k].score;
        if (e.status != 0) continue;
        if (e.hit_boundary && attempt[i] < max_attempts) {  // align.rs:55-63: double the band (at least |mean_shift|, at least 1)
          const int ams = pr.mean_shift < 0 ? -pr.mean_shift : pr.mean_shift;
          bw[i] = std::max(2 * bw[i], std::max(1, ams));
          ++attempt[i];
          next.push_back(i);
          continue;
        }
        edit_from_runs(pr.ref, pr.rlen, pr.qry, pr.qlen, hr + hj[k].runs_off, ho[k].n_runs, e);
      }
      pos = end;
    }
    active.swap(next);
  }
  cudaEventDestroy(ev0), cudaEventDestroy(ev1);
  cudaStreamDestroy(st);
}

}  // namespace na
}  // namespace pgmm
