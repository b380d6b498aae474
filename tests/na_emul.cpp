// TEST-ONLY: a serial emulation of nextalign.cu's kernel -- the same lane / macro-step / micro-step schedule, the same
// in-place row state, the same chains and the same 32-wide traceback walk, built from the product's nextalign_core.h and
// nextalign_host.cpp -- so that the band-coordinate formulation can be checked against the oracle without a GPU
// (tests/test_nextalign_emul.py).  Not part of the product.
#include "../pangraph_b200/csrc/nextalign_core.h"
#include "../pangraph_b200/csrc/nextalign_host.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

using namespace pgmm::na;

namespace {
struct AttemptOut {
  int32_t score, hit_boundary, status;
  std::vector<uint32_t> runs;
};

AttemptOut attempt(const uint8_t *R, int rlen, const uint8_t *Q, int qlen, int ms, int bw, const Params &p) {
  Geom g;
  g.rlen = rlen, g.qlen = qlen, g.ms = ms, g.bw = bw, g.W = 2 * bw + 1;
  const int W = g.W, m = std::max(2, (W + 31) / 32);
  std::vector<int32_t> S(W + 2, 123456789), QG(W + 2, 123456789);  // poison: must never be read before written
  std::vector<uint8_t> B((size_t)(rlen + 1) * W, 0xee), E(rlen + 1, 0xee), T(qlen + 1, 0xee);
  int32_t lastS[32], lastRG[32];
  for (int l = 0; l < 32; ++l) lastS[l] = 0, lastRG[l] = kNoAlign;
  const int lanes_used = (W + m - 1) / m, n_macro = rlen + lanes_used;
  for (int s = 0; s < n_macro; ++s) {
    int32_t inS[32], rg[32], curS[32];
    for (int l = 0; l < 32; ++l) inS[l] = l ? lastS[l - 1] : lastS[0], rg[l] = l ? lastRG[l - 1] : kNoAlign, curS[l] = inS[l];
    for (int c = 0; c < m; ++c)
      for (int lane = 0; lane < 32; ++lane) {
        const int ri = s - lane;
        const bool row_ok = ri >= 0 && ri <= rlen;
        const int K = lane * m + c, qpos = g.b(ri) + K;
        if (row_ok && K < W && qpos >= g.begin(ri) && qpos < g.end(ri)) {
          CellOut o;
          if (ri == 0) o.S = row0_score(p, qpos), o.path = row0_path(qpos), o.qry_gaps = kNoAlign, o.ref_gaps = rg[lane];
          else {
            CellIn in;
            in.diagS = S[K], in.leftS = curS[lane], in.ref_gaps = rg[lane], in.upS = S[K + 1], in.qry_gaps = QG[K + 1];
            in.qc = qpos > 0 ? Q[qpos - 1] : 0, in.rc = R[ri - 1];
            o = cell(g, p, row_geom(g, ri), ri, qpos, in);
          }
          S[K] = o.S, QG[K] = o.qry_gaps, rg[lane] = o.ref_gaps, curS[lane] = o.S;
          B[(size_t)ri * W + K] = (uint8_t)o.path;
        }
      }
    for (int lane = 0; lane < 32; ++lane) {
      const int ri = s - lane;
      if (ri >= 0 && ri <= rlen) lastS[lane] = curS[lane], lastRG[lane] = rg[lane];
    }
  }
  AttemptOut out;
  int32_t final_score = 0;
  {
    const int t0 = g.end_unforced(rlen), K_last = t0 - 1 - g.b(rlen);
    const int owner = (K_last >= 0 && K_last < W) ? K_last / m : 0;
    const int32_t ownS = lastS[owner], ownRG = lastRG[owner];
    if (rlen == 0) final_score = row0_score(p, qlen);
    else {
      bool have = false;
      const int i1 = qlen + ms + bw;
      if (i1 < rlen) {
        int32_t upS, qg;
        if (i1 >= 0) upS = S[0], qg = QG[0];
        else upS = row0_score(p, qlen), qg = kNoAlign;
        for (int ri = std::max(i1, 0) + 1; ri <= rlen; ++ri) {
          CellIn in;
          in.diagS = row0_score(p, qlen - 1), in.leftS = 0, in.ref_gaps = kNoAlign, in.upS = upS, in.qry_gaps = qg;
          in.qc = qlen > 0 ? Q[qlen - 1] : 0, in.rc = R[ri - 1];
          const CellOut o = cell(g, p, row_geom(g, ri), ri, qlen, in);
          E[ri] = (uint8_t)o.path, upS = o.S, qg = o.qry_gaps;
        }
        final_score = upS, have = true;
      }
      if (!have && t0 <= qlen) {
        int32_t leftS, rgc;
        if (K_last >= 0 && K_last < W) leftS = ownS, rgc = ownRG;
        else leftS = col0_score(p, rlen), rgc = kNoAlign;
        const RowGeom rw_last = row_geom(g, rlen);
        for (int qpos = t0; qpos <= qlen; ++qpos) {
          CellIn in;
          in.diagS = col0_score(p, rlen - 1), in.leftS = leftS, in.ref_gaps = rgc, in.upS = 0, in.qry_gaps = kNoAlign;
          in.qc = Q[qpos - 1], in.rc = R[rlen - 1];
          const CellOut o = cell(g, p, rw_last, rlen, qpos, in);
          T[qpos] = (uint8_t)o.path, leftS = o.S, rgc = o.ref_gaps;
        }
        final_score = leftS, have = true;
      }
      if (!have) final_score = S[qlen - g.b(rlen)];
    }
  }
  const auto pa = [&](int ri, int qpos) {
    return path_at(g, ri, qpos, [&](int64_t i) { return (int)B[(size_t)i]; }, [&](int r) { return (int)E[r]; }, [&](int q) { return (int)T[q]; });
  };
  int r_pos = rlen, q_pos = qlen, current = 0, hb = 0, status = 0;
  uint32_t last = 0;
  const auto push = [&](uint32_t op, uint32_t len) {
    if (last != 0 && (last & 3u) == op) last += len << 2;
    else {
      if (last != 0) out.runs.push_back(last);
      last = len << 2 | op;
    }
  };
  while (r_pos > 0 || q_pos > 0) {
    if (current == 0) {
      int o[32];
      bool valid[32];
      unsigned stop = 0, bnd = 0;
      for (int lane = 0; lane < 32; ++lane) {
        const int rr = r_pos - lane, qq = q_pos - lane;
        valid[lane] = rr >= 0 && qq >= 0 && (rr > 0 || qq > 0) && g.exists(rr, qq);
        o[lane] = valid[lane] ? pa(rr, qq) : 0;
        if (!(o[lane] & kMatch)) stop |= 1u << lane;
        if (o[lane] & kBoundary) bnd |= 1u << lane;
      }
      const int run = stop ? __builtin_ffs((int)stop) - 1 : 32;
      if (run > 0) {
        if (bnd & (run == 32 ? 0xffffffffu : ((1u << run) - 1u))) hb = 1;
        push(0, (uint32_t)run), r_pos -= run, q_pos -= run;
      }
      if (run == 32 || !(r_pos > 0 || q_pos > 0)) continue;
      if (!valid[run]) { status = -3; break; }
      const int oo = o[run];
      if (oo & kBoundary) hb = 1;
      const int op = walk_step(oo, current, r_pos, q_pos);
      if (op < 0) { status = -2; break; }
      push((uint32_t)op, 1);
    } else {
      if (!g.exists(r_pos, q_pos)) { status = -3; break; }
      const int oo = pa(r_pos, q_pos);
      if (oo & kBoundary) hb = 1;
      const int op = walk_step(oo, current, r_pos, q_pos);
      if (op < 0) { status = -2; break; }
      push((uint32_t)op, 1);
    }
  }
  if (last != 0) out.runs.push_back(last);
  out.score = final_score, out.hit_boundary = hb, out.status = status;
  return out;
}
}  // namespace

// map_variations through the emulated kernel; outputs like orc_map_variations.  Returns the status.
extern "C" __attribute__((visibility("default"))) int na_emul_map_variations(
    const char *ref, int rlen, const char *qry, int qlen, int mean_shift, int band_width, int extra_band_width, int max_attempts, int32_t *n_sub,
    int32_t *sub_pos, char *sub_chr, int32_t *n_del, int32_t *del_pos, int32_t *del_len, int32_t *n_ins, int32_t *ins_pos, int32_t *ins_len,
    char *ins_seq, int32_t *hit_boundary, int32_t *attempts, int32_t *score) {
  Params p;
  p.ext = 0, p.gopen = 6, p.mismatch = 1, p.match = 3, p.left_free = 1, p.right_free = 1, p.left_align = 1;
  std::vector<uint8_t> R((size_t)rlen + 1), Q((size_t)qlen + 1);
  if (qlen < 1 || !encode(ref, rlen, R.data()) || !encode(qry, qlen, Q.data())) return -1;
  int bw = band_width + extra_band_width, att = 1;
  AttemptOut a = attempt(R.data(), rlen, Q.data(), qlen, mean_shift, bw, p);
  while (a.status == 0 && a.hit_boundary && att < max_attempts) {
    const int ams = mean_shift < 0 ? -mean_shift : mean_shift;
    bw = std::max(2 * bw, std::max(1, ams));
    ++att;
    a = attempt(R.data(), rlen, Q.data(), qlen, mean_shift, bw, p);
  }
  *hit_boundary = a.hit_boundary, *attempts = att, *score = a.score;
  if (a.status != 0) return a.status;
  Edit e;
  edit_from_runs(ref, rlen, qry, qlen, a.runs.data(), (int64_t)a.runs.size(), e);
  *n_sub = (int32_t)e.sub_pos.size(), *n_del = (int32_t)e.del_pos.size(), *n_ins = (int32_t)e.ins_pos.size();
  memcpy(sub_pos, e.sub_pos.data(), e.sub_pos.size() * 4), memcpy(sub_chr, e.sub_chr.data(), e.sub_chr.size());
  memcpy(del_pos, e.del_pos.data(), e.del_pos.size() * 4), memcpy(del_len, e.del_len.data(), e.del_len.size() * 4);
  memcpy(ins_pos, e.ins_pos.data(), e.ins_pos.size() * 4), memcpy(ins_len, e.ins_len.data(), e.ins_len.size() * 4);
  memcpy(ins_seq, e.ins_seq.data(), e.ins_seq.size());
  return 0;
}
