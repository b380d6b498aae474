// K6: the banded affine alignment behind pangraph's map_variations -- the parts shared by the CUDA kernel
// (nextalign.cu) and the host: band geometry, the update of one cell, where a traceback byte lives.
//
// Reference (PG = packages/pangraph/src): PG/align/nextclade/align/band_2d.rs:36-54 (simple_stripes),
// score_matrix.rs:23-200 (the cell update, statement by statement), backtrace.rs:17-98.
//
// Geometry.  Row ri of the reference's Band2d holds query columns [begin(ri), end(ri)) with
//   b(ri) = -mean_shift - band_width + ri,  begin = clamp(b, 0, qlen),  end = clamp(b + W, 1, qlen + 1),  W = 2 band_width + 1,
// and stripes[0].begin = 0, stripes[ref_len].end = qlen + 1 forced.  A cell (ri, qpos) has BAND COLUMN K = qpos - b(ri); the
// three cells it reads -- (ri-1, qpos-1), (ri, qpos-1), (ri-1, qpos) -- have band columns K, K-1, K+1.  Cells with 0 <= K < W are
// "in band" and are computed by the wavefront kernel; they only ever read in-band cells.  The clamps and the two forced stripe
// ends create cells outside the band, all of them closed forms or one-directional chains:
//   * row 0 left of the band (forced begin): initial values;
//   * (ri, 0) when the whole band lies left of the matrix (b + W <= 0): first-column values;
//   * (ri, qlen) when the whole band lies right of the matrix (b > qlen): a chain down the right edge ("edge" array);
//   * the last row right of the band (forced end): a chain along the last row ("tail" array).
// The chains are walked by one thread after the wavefront (they read nothing but their own predecessor).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define NA_HD __host__ __device__ __forceinline__
#else
#define NA_HD inline
#endif

namespace pgmm {
namespace na {

constexpr int kMatch = 1, kRefGapMatrix = 2, kQryGapMatrix = 4, kRefGapExtend = 8, kQryGapExtend = 16, kBoundary = 32;
constexpr int32_t kNoAlign = -1000000000;

struct Params {  // NextalignParams as map_variations sets them (align/nextclade/align/params.rs:143-175, map_variations.rs:45-49)
  int32_t ext, gopen, mismatch, match;  // penalty_gap_extend, penalty_gap_open (flat gap_open_close), penalty_mismatch, score_match
  int32_t left_free, right_free, left_align;
};

struct Geom {
  int32_t rlen, qlen, ms, bw, W;
  NA_HD int b(int ri) const { return -ms - bw + ri; }
  NA_HD int begin(int ri) const {
    if (ri == 0) return 0;
    const int v = b(ri);
    return v < 0 ? 0 : v > qlen ? qlen : v;
  }
  NA_HD int end_unforced(int ri) const {
    const int v = b(ri) + W;
    return v < 1 ? 1 : v > qlen + 1 ? qlen + 1 : v;
  }
  NA_HD int end(int ri) const { return ri == rlen ? qlen + 1 : end_unforced(ri); }
  NA_HD bool exists(int ri, int qpos) const { return ri >= 0 && ri <= rlen && qpos >= begin(ri) && qpos < end(ri); }
};

// code + 1 is the set of bases (T=1 A=2 C=4 G=8), N = 14; inputs never hold gaps
NA_HD bool nuc_match(int x, int y) { return ((x + 1) & (y + 1)) != 0; }

struct CellIn {
  int32_t diagS, leftS, ref_gaps, upS, qry_gaps;  // only read where the reference reads them
  int qc, rc;                                     // query / reference code of the cell (qpos - 1, ri - 1)
};
struct CellOut {
  int32_t S, ref_gaps, qry_gaps;
  int path;
};

// The stripe limits a cell of row ri looks at: computed once per row, not per cell.
struct RowGeom {
  int beg, beg1, end1, end2;  // begin(ri), begin(ri-1), end(ri-1), end(ri-2) (unused for ri == 1)
};
NA_HD RowGeom row_geom(const Geom &g, int ri) {
  RowGeom r;
  r.beg = g.begin(ri), r.beg1 = g.begin(ri - 1), r.end1 = g.end(ri - 1), r.end2 = ri >= 2 ? g.end(ri - 2) : 0;
  return r;
}

// One cell with ri >= 1 (score_matrix.rs:92-195).  `in.ref_gaps` is the running value of the row (NO_ALIGN at its start).
NA_HD CellOut cell(const Geom &g, const Params &p, const RowGeom &rw, int ri, int qpos, const CellIn &in) {
  CellOut o;
  int tmp_path = 0, origin = 0;
  int32_t score = kNoAlign, tmp_score;
  o.ref_gaps = in.ref_gaps, o.qry_gaps = kNoAlign;
  if (qpos == 0) {
    tmp_path = kQryGapExtend, origin = kQryGapMatrix;
    if (p.left_free) score = 0;
    else score = -p.gopen - (ri - 1) * p.ext;  // scores[(ri-1, 0)] - ext unrolled: row 1 opens, every further row extends
  } else {
    const int beg = rw.beg, beg1 = rw.beg1, end1 = rw.end1;
    if (qpos > beg1 && qpos - 1 < end1) {
      if (in.qc == 14 || in.rc == 14) score = in.diagS + p.match - 1;
      else if (nuc_match(in.qc, in.rc)) score = in.diagS + p.match;
      else score = in.diagS - p.mismatch;
      origin = kMatch;
    } else if (ri < g.rlen && qpos < g.qlen) tmp_path |= kBoundary;
    if (qpos > beg) {
      int32_t r_gap_extend, r_gap_open;
      if (ri != g.rlen || !p.right_free) r_gap_extend = in.ref_gaps - p.ext, r_gap_open = in.leftS - p.gopen;
      else r_gap_extend = in.ref_gaps, r_gap_open = in.leftS;
      if (r_gap_extend >= r_gap_open && qpos > beg + 1) tmp_score = r_gap_extend, tmp_path += kRefGapExtend;
      else tmp_score = r_gap_open;
      o.ref_gaps = tmp_score;
      if (score - p.left_align < tmp_score) score = tmp_score, origin = kRefGapMatrix;
    } else if (ri < g.rlen && qpos < g.qlen) tmp_path |= kBoundary;
    if (qpos < end1) {
      int32_t q_gap_extend, q_gap_open;
      if (qpos != g.qlen || !p.right_free) q_gap_extend = in.qry_gaps - p.ext, q_gap_open = in.upS - p.gopen;
      else q_gap_extend = in.qry_gaps, q_gap_open = in.upS;
      // (ri == 1: qry_gaps is still NO_ALIGN, the extension can never win, stripes[ri - 2] is not looked at)
      if (q_gap_extend >= q_gap_open && ri >= 2 && qpos < rw.end2) tmp_score = q_gap_extend, tmp_path += kQryGapExtend;
      else tmp_score = q_gap_open;
      o.qry_gaps = tmp_score;
      if (score - p.left_align < tmp_score) score = tmp_score, origin = kQryGapMatrix;
    } else if (qpos < g.qlen && ri < g.rlen) tmp_path |= kBoundary;  // (qry_gaps[qpos] = NO_ALIGN: it never held anything else)
  }
  o.S = score, o.path = tmp_path + origin;
  return o;
}

// Row 0 (score_matrix.rs:62-81) and the first column are closed forms.
NA_HD int32_t row0_score(const Params &p, int qpos) { return (p.left_free || qpos == 0) ? 0 : -p.gopen - (qpos - 1) * p.ext; }
NA_HD int row0_path(int qpos) { return qpos == 0 ? 0 : kRefGapExtend + kRefGapMatrix; }
NA_HD int32_t col0_score(const Params &p, int ri) { return (p.left_free || ri == 0) ? 0 : -p.gopen - (ri - 1) * p.ext; }

// Where the traceback byte of an existing cell lives: the band matrix [ri * W + K], the right-edge chain [ri], the tail of
// the last row [qpos], or a closed form.
template <class BandAt, class EdgeAt, class TailAt>
NA_HD int path_at(const Geom &g, int ri, int qpos, BandAt band, EdgeAt edge, TailAt tail) {
  if (ri == 0) return row0_path(qpos);
  const int K = qpos - g.b(ri);
  if (K >= 0 && K < g.W) return band((int64_t)ri * g.W + K);
  if (qpos == 0) return kQryGapExtend + kQryGapMatrix;  // band left of the matrix
  if (K < 0) return edge(ri);                           // band right of the matrix: qpos == qlen
  return tail(qpos);                                    // last row, right of the band
}

// One step of backtrace.rs:41-85 from (r_pos, q_pos) with the cell's byte: returns the op (0 = match column, 1 = query base
// against a reference gap, 2 = reference base against a query gap, -1 = the reference's unreachable!()) and moves.
NA_HD int walk_step(int origin, int &current, int &r_pos, int &q_pos) {
  if ((origin & kMatch) && current == 0) {
    --q_pos, --r_pos;
    return 0;
  }
  if (((origin & kRefGapMatrix) && current == 0) || current == kRefGapMatrix) {
    --q_pos;
    current = (origin & kRefGapExtend) ? kRefGapMatrix : 0;
    return 1;
  }
  if (((origin & kQryGapMatrix) && current == 0) || current == kQryGapMatrix) {
    --r_pos;
    current = (origin & kQryGapExtend) ? kQryGapMatrix : 0;
    return 2;
  }
  return -1;
}

}  // namespace na
}  // namespace pgmm
