// Host half of the path, above the aligner: what pangraph does with the hits of one alignment round.
// Mirrors, in C++ behind a C-ABI, the reference's Rust:
//   align_with_minimap2_lib / Alignment::from_minimap_paf_obj  packages/pangraph/src/align/minimap2_lib/align_with_minimap2_lib.rs:15-121
//   split_matches, keep_groups, generate_subalignment, side_patches  packages/pangraph/src/pangraph/split_matches.rs:13-237
//   add_flanking_indel, cigar_matches_len, cigar_total_len            packages/pangraph/src/align/bam/cigar.rs:14-96
//   alignment_energy2                                                  packages/pangraph/src/align/energy.rs:37-54
//   filter_matches, is_match_compatible, update_intervals              packages/pangraph/src/pangraph/graph_merging.rs:187-242
//   the alignment part of self_merge                                   packages/pangraph/src/pangraph/graph_merging.rs:95-121
// Same names, same argument meaning, same error conditions (returned as negative codes instead of eyre reports).
#include "../../include/pgmm_b200.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <numeric>
#include <string>
#include <vector>

namespace {

enum : uint32_t { OP_M = 0, OP_I = 1, OP_D = 2, OP_N = 3, OP_S = 4, OP_H = 5, OP_P = 6, OP_EQ = 7, OP_X = 8 };

struct Aln {  // owning image of pgmm_alignment_t
  pgmm_alignment_t a;
  std::vector<uint32_t> cigar;
};

inline bool is_match_op(uint32_t op) { return op == OP_M || op == OP_EQ || op == OP_X; }

Aln from_c(const pgmm_alignment_t &c) {
  Aln x;
  x.a = c;
  x.cigar.assign(c.cigar, c.cigar + c.n_cigar);
  x.a.cigar = nullptr;
  return x;
}
pgmm_alignment_t to_c(const Aln &x) {
  pgmm_alignment_t c = x.a;
  c.n_cigar = (uint32_t)x.cigar.size();
  c.cigar = (uint32_t *)malloc(sizeof(uint32_t) * (x.cigar.size() ? x.cigar.size() : 1));
  memcpy(c.cigar, x.cigar.data(), x.cigar.size() * sizeof(uint32_t));
  return c;
}
int emit(const std::vector<Aln> &v, pgmm_alignment_t **out, size_t *n_out) {
  *n_out = v.size();
  *out = (pgmm_alignment_t *)malloc(sizeof(pgmm_alignment_t) * (v.size() ? v.size() : 1));
  for (size_t i = 0; i < v.size(); ++i) (*out)[i] = to_c(v[i]);
  return 0;
}

// (start_index, end_index) of the CIGAR groups to keep (split_matches.rs:32-92)
int keep_groups(const std::vector<uint32_t> &cigar, uint64_t thr, std::vector<std::pair<size_t, size_t>> &groups) {
  bool have_start = false, have_last = false;
  size_t g_start = 0, last_match = 0;
  uint64_t M_sum = 0, I_sum = 0, D_sum = 0;
  for (size_t i = 0; i < cigar.size(); ++i) {
    const uint32_t op = cigar[i] & 0xf;
    const uint64_t len = cigar[i] >> 4;
    if (!have_start) {
      if (!is_match_op(op)) continue;
      g_start = i, have_start = true;
    }
    if (is_match_op(op)) {
      M_sum += len, I_sum = 0, D_sum = 0;
      last_match = i, have_last = true;
    } else if (op == OP_I) I_sum += len;
    else if (op == OP_D) D_sum += len;
    else return -3;  // "Unexpected CIGAR operation" (soft/hard clip, pad, skip)
    if (std::max(I_sum, D_sum) >= thr) {
      if (have_start && have_last && M_sum >= thr) groups.emplace_back(g_start, last_match);
      have_start = have_last = false;
      M_sum = I_sum = D_sum = 0;
    }
  }
  if (have_start && have_last && M_sum >= thr) groups.emplace_back(g_start, last_match);
  return 0;
}

uint64_t matches_len(const std::vector<uint32_t> &c) {
  uint64_t s = 0;
  for (uint32_t x : c)
    if (is_match_op(x & 0xf)) s += x >> 4;
  return s;
}
uint64_t total_len(const std::vector<uint32_t> &c) {
  uint64_t s = 0;
  for (uint32_t x : c) s += x >> 4;
  return s;
}

// the sub-alignment spanning CIGAR elements [g0, g1] (split_matches.rs:95-185)
Aln subalignment(const Aln &aln, size_t g0, size_t g1) {
  uint64_t q_beg = 0, q_end = 0, r_beg = 0, r_end = 0, qp = 0, rp = 0;
  for (size_t i = 0; i < aln.cigar.size(); ++i) {
    const uint32_t op = aln.cigar[i] & 0xf;
    const uint64_t len = aln.cigar[i] >> 4;
    if (i == g0) q_beg = qp, r_beg = rp;
    if (is_match_op(op) || op == OP_I) qp += len;
    if (is_match_op(op) || op == OP_D) rp += len;
    if (i == g1) q_end = qp, r_end = rp;
  }
  Aln s;
  s.a = aln.a;
  s.a.ref_start = aln.a.ref_start + r_beg, s.a.ref_end = aln.a.ref_start + r_end;
  if (!aln.a.reverse) s.a.qry_start = aln.a.qry_start + q_beg, s.a.qry_end = aln.a.qry_start + q_end;
  else s.a.qry_start = aln.a.qry_end - q_end, s.a.qry_end = aln.a.qry_end - q_beg;
  s.cigar.assign(aln.cigar.begin() + g0, aln.cigar.begin() + g1 + 1);
  s.a.matches = matches_len(s.cigar);
  s.a.length = total_len(s.cigar);
  return s;  // quality, orientation, divergence and align are inherited
}

// extend or add an insertion/deletion before the first match op on one side (bam/cigar.rs:60-96)
void add_flanking_indel(std::vector<uint32_t> &c, uint32_t kind, uint64_t add_len, bool leading) {
  long replace = -1;
  const long n = (long)c.size();
  for (long k = 0; k < n; ++k) {
    const long i = leading ? k : n - 1 - k;
    const uint32_t op = c[i] & 0xf;
    if (is_match_op(op)) break;
    if (op == kind) replace = i;  // the reference keeps overwriting: the last one met before the match wins
  }
  if (replace >= 0) c[replace] = (uint32_t)(((c[replace] >> 4) + add_len) << 4) | kind;
  else c.insert(leading ? c.begin() : c.end(), (uint32_t)(add_len << 4) | kind);
}

// absorb short overhangs into the alignment (split_matches.rs:189-237)
void side_patches(Aln &aln, uint64_t thr) {
  pgmm_alignment_t &a = aln.a;
  std::vector<uint32_t> &ops = aln.cigar;
  {
    const uint64_t rs = a.ref_start, re = a.ref_end, rL = a.ref_len;
    if (rs > 0 && rs < thr) a.ref_start = 0, a.length += rs, add_flanking_indel(ops, OP_D, rs, true);
    if (re < rL && rL - re < thr) a.ref_end = rL, a.length += rL - re, add_flanking_indel(ops, OP_D, rL - re, false);
  }
  {
    const uint64_t qs = a.qry_start, qe = a.qry_end, qL = a.qry_len;
    if (qs > 0 && qs < thr) a.qry_start = 0, a.length += qs, add_flanking_indel(ops, OP_I, qs, !a.reverse);
    if (qe < qL && qL - qe < thr) a.qry_end = qL, a.length += qL - qe, add_flanking_indel(ops, OP_I, qL - qe, (bool)a.reverse);
  }
}

int split_matches(const Aln &aln, const pgmm_alignment_args_t &args, std::vector<Aln> &out) {
  std::vector<std::pair<size_t, size_t>> groups;
  const int rc = keep_groups(aln.cigar, args.indel_len_threshold, groups);
  if (rc) return rc;
  for (auto &g : groups) {
    Aln s = subalignment(aln, g.first, g.second);
    side_patches(s, args.indel_len_threshold);
    out.push_back(std::move(s));
  }
  return 0;
}

double energy2(const pgmm_alignment_t &a, const pgmm_alignment_args_t &args) {  // energy.rs:37-54
  const uint64_t L = a.matches;
  const double M = (a.has_divergence ? a.divergence : 0.0) * (double)L;
  int C = 4;
  if (a.qry_start == 0) C -= 1;
  if (a.qry_end == a.qry_len) C -= 1;
  if (a.ref_start == 0) C -= 1;
  if (a.ref_end == a.ref_len) C -= 1;
  return -(double)L + (double)C * args.alpha + M * args.beta;
}

bool no_overlap(const std::vector<std::pair<uint64_t, uint64_t>> &v, uint64_t s, uint64_t e) {
  for (auto &iv : v)
    if (iv.second > s && iv.first < e) return false;  // Interval::has_overlap_with, utils/interval.rs
  return true;
}

void filter_matches(const std::vector<Aln> &alns, const pgmm_alignment_args_t &args, std::vector<Aln> &out) {
  std::vector<std::pair<double, size_t>> keyed;
  for (size_t i = 0; i < alns.size(); ++i) {
    const double e = energy2(alns[i].a, args);
    if (e < 0.0) keyed.emplace_back(e, i);
  }
  std::stable_sort(keyed.begin(), keyed.end(), [](const std::pair<double, size_t> &x, const std::pair<double, size_t> &y) { return x.first < y.first; });
  std::map<uint64_t, std::vector<std::pair<uint64_t, uint64_t>>> accepted;  // one interval list per block id
  for (auto &ke : keyed) {
    const pgmm_alignment_t &a = alns[ke.second].a;
    const bool ref_ok = no_overlap(accepted[a.ref_name], a.ref_start, a.ref_end);
    const bool qry_ok = no_overlap(accepted[a.qry_name], a.qry_start, a.qry_end);
    if (ref_ok && qry_ok) {
      out.push_back(alns[ke.second]);
      accepted[a.ref_name].emplace_back(a.ref_start, a.ref_end);
      accepted[a.qry_name].emplace_back(a.qry_start, a.qry_end);
    }
  }
}

int align_blocks(int n, const uint64_t *ids, const char *const *consensus, const pgmm_alignment_args_t &args, std::vector<Aln> &out) {
  const char *preset;
  switch (args.sensitivity) {
    case 5: preset = "asm5"; break;
    case 10: preset = "asm10"; break;
    case 20: preset = "asm20"; break;
    default: return -1;  // "Unknown sensitivity preset"
  }
  if (n < 0 || (n > 0 && (!ids || !consensus))) return -2;
  if (n == 0) return 0;
  // BTreeMap<BlockId, _> iterates in ascending id order (align_with_minimap2_lib.rs:19-22)
  std::vector<int> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return ids[a] < ids[b]; });
  std::vector<std::string> names(n);
  std::vector<const char *> seqs(n), name_ptrs(n);
  for (int i = 0; i < n; ++i) {
    char buf[32];
    snprintf(buf, sizeof buf, "%llu", (unsigned long long)ids[order[i]]);
    names[i] = buf;
    seqs[i] = consensus[order[i]];
  }
  for (int i = 0; i < n; ++i) name_ptrs[i] = names[i].c_str();
  // Minimap2Args { x: preset, k, c: true, X: true, s: max(l - 10, 5), bucket_bits: 14 }  (:49-57; options_args.rs:273-331)
  mm_idxopt_t io;
  mm_mapopt_t mo;
  if (mm_set_opt(nullptr, &io, &mo) != 0 || mm_set_opt(preset, &io, &mo) != 0) return -1;
  if (args.kmer_length > 0) io.k = (short)args.kmer_length;
  mo.flag |= MM_F_OUT_CG | MM_F_CIGAR;
  const int64_t l = (int64_t)args.indel_len_threshold - 10;
  mo.min_dp_max = (int)(l < 5 ? 5 : l);
  mo.flag |= MM_F_ALL_CHAINS | MM_F_NO_DIAG | MM_F_NO_DUAL | MM_F_NO_LJOIN;
  io.bucket_bits = 14;
  if (mm_check_opt(&io, &mo) != 0) return -4;
  mm_idx_t *mi = pgmm_idx_upload(n, seqs.data(), name_ptrs.data());
  if (!mi) return -5;  // "minimap2: failed to create index"
  pgmm_idx_build(mi, io.w, io.k, io.bucket_bits);
  mm_mapopt_update(&mo, mi);
  std::vector<int> n_regs(n, 0);
  std::vector<mm_reg1_t *> regs(n, nullptr);
  pgmm_map_self(mi, &mo, n_regs.data(), regs.data());
  int rc = 0;
  for (int q = 0; q < n; ++q) {
    for (int j = 0; j < n_regs[q]; ++j) {
      const mm_reg1_t &r = regs[q][j];
      if (!r.p) {
        rc = -6;  // "Unable to find CIGAR string in the result"
        continue;
      }
      Aln x;
      memset(&x.a, 0, sizeof(x.a));
      x.a.qry_name = ids[order[q]], x.a.qry_len = strlen(seqs[q]);
      x.a.qry_start = (uint64_t)r.qs, x.a.qry_end = (uint64_t)r.qe;
      x.a.ref_name = ids[order[r.rid]], x.a.ref_len = mi->seq[r.rid].len;
      x.a.ref_start = (uint64_t)r.rs, x.a.ref_end = (uint64_t)r.re;
      x.a.matches = (uint64_t)r.mlen, x.a.length = (uint64_t)r.blen, x.a.quality = r.mapq;
      x.a.reverse = r.rev;
      x.a.has_divergence = 1, x.a.divergence = 1.0 - mm_event_identity(&r);  // PAF "de" (map.rs:321-325)
      x.a.align = (double)r.p->dp_score;                                      // PAF "AS"
      x.cigar.assign(r.p->cigar, r.p->cigar + r.p->n_cigar);
      out.push_back(std::move(x));
    }
    for (int j = 0; j < n_regs[q]; ++j) free(regs[q][j].p);
    free(regs[q]);
  }
  mm_idx_destroy(mi);
  return rc;
}

}  // namespace

extern "C" {

void pgmm_alignment_args_default(pgmm_alignment_args_t *a) {
  a->indel_len_threshold = 100, a->alpha = 100.0, a->beta = 10.0, a->sensitivity = 10, a->kmer_length = 0;
}

int pgmm_align_with_minimap2_lib(int n_blocks, const uint64_t *block_ids, const char *const *consensus, const pgmm_alignment_args_t *args,
                                 pgmm_alignment_t **out, size_t *n_out) {
  std::vector<Aln> v;
  const int rc = align_blocks(n_blocks, block_ids, consensus, *args, v);
  if (rc) {
    *out = nullptr, *n_out = 0;
    return rc;
  }
  return emit(v, out, n_out);
}

int pgmm_split_matches(const pgmm_alignment_t *aln, const pgmm_alignment_args_t *args, pgmm_alignment_t **out, size_t *n_out) {
  std::vector<Aln> v;
  const int rc = split_matches(from_c(*aln), *args, v);
  if (rc) {
    *out = nullptr, *n_out = 0;
    return rc;
  }
  return emit(v, out, n_out);
}

double pgmm_alignment_energy2(const pgmm_alignment_t *aln, const pgmm_alignment_args_t *args) { return energy2(*aln, *args); }

int pgmm_filter_matches(const pgmm_alignment_t *alns, size_t n, const pgmm_alignment_args_t *args, pgmm_alignment_t **out, size_t *n_out) {
  std::vector<Aln> in, v;
  for (size_t i = 0; i < n; ++i) in.push_back(from_c(alns[i]));
  filter_matches(in, *args, v);
  return emit(v, out, n_out);
}

int pgmm_find_filtered_matches(int n_blocks, const uint64_t *block_ids, const char *const *consensus, const pgmm_alignment_args_t *args,
                               pgmm_alignment_t **out, size_t *n_out) {
  *out = nullptr, *n_out = 0;
  std::vector<Aln> found, split, kept;
  int rc = align_blocks(n_blocks, block_ids, consensus, *args, found);
  if (rc) return rc;
  for (const Aln &m : found) {
    if (m.a.qry_name == m.a.ref_name) continue;  // self-alignments are dropped after being computed (graph_merging.rs:108)
    rc = split_matches(m, *args, split);
    if (rc) return rc;
  }
  filter_matches(split, *args, kept);
  return emit(kept, out, n_out);
}

void pgmm_alignments_free(pgmm_alignment_t *alns, size_t n) {
  if (!alns) return;
  for (size_t i = 0; i < n; ++i) free(alns[i].cigar);
  free(alns);
}

}  // extern "C"
