// Host half of K6: see nextalign_host.h.  Reference: PG/align/nextclade/align/insertions_strip.rs:48-98,
// PG/align/nextclade/analyze/nuc_changes.rs:19-70, PG/align/nextclade/align_with_nextclade.rs:44-67, PG/align/map_variations.rs:58-79.
#include "nextalign_host.h"

#include <cstring>

namespace pgmm {
namespace na {

bool encode(const char *s, int64_t n, uint8_t *out) {
  static const struct Table {
    int8_t t[256];
    Table() {
      memset(t, -1, sizeof(t));
      const char *abc = "TAWCYMHGKRDSBVN";
      for (int i = 0; abc[i]; ++i) t[(unsigned char)abc[i]] = (int8_t)i;
    }
  } tab;
  for (int64_t i = 0; i < n; ++i) {
    const int8_t c = tab.t[(unsigned char)s[i]];
    if (c < 0) return false;
    out[i] = (uint8_t)c;
  }
  return true;
}

void edit_from_runs(const char *ref, int32_t rlen, const char *qry, int32_t qlen, const uint32_t *runs, int64_t n_runs, Edit &e) {
  e.sub_pos.clear(), e.sub_chr.clear(), e.del_pos.clear(), e.del_len.clear(), e.ins_pos.clear(), e.ins_len.clear(), e.ins_seq.clear();
  int32_t ri = 0, qi = 0;              // next reference / query base
  // insertions_strip: an insertion collects query bases while the reference shows gaps; it is closed by the next column
  // that holds a reference base and sits after the reference base that precedes it (-1: before the first one)
  int32_t cur_ins = 0;
  // find_nuc_changes over the stripped query (one column per reference base)
  int64_t n_del = 0, del_at = -1, a_start = -1, a_end = -1;
  bool before = true;
  const auto close_ins = [&] {
    if (cur_ins) e.ins_len.push_back(cur_ins), cur_ins = 0;
  };
  for (int64_t k = n_runs - 1; k >= 0; --k) {
    const uint32_t op = runs[k] & 3u;
    const int32_t len = (int32_t)(runs[k] >> 2);
    if (op == 1) {  // reference gap
      if (cur_ins == 0) e.ins_pos.push_back(ri);  // = (index of the preceding reference base) + 1
      e.ins_seq.append(qry + qi, (size_t)len);
      cur_ins += len, qi += len;
    } else if (op == 0) {  // query base on reference base
      close_ins();
      if (before) a_start = ri, before = false;
      else if (n_del > 0) e.del_pos.push_back((int32_t)del_at), e.del_len.push_back((int32_t)n_del), n_del = 0;
      for (int32_t j = 0; j < len; ++j)
        if (qry[qi + j] != ref[ri + j]) e.sub_pos.push_back(ri + j), e.sub_chr.push_back(qry[qi + j]);
      ri += len, qi += len;
      a_end = ri;
    } else {  // query gap
      close_ins();
      if (!before) {
        if (n_del == 0) del_at = ri;
        n_del += len;
      }
      ri += len;
    }
  }
  close_ins();
  (void)qlen;
  // terminal gaps are not deletions for find_nuc_changes; align_with_nextclade appends them (after the sorted inner ones)
  if (a_start >= 0 && a_end >= 0) {
    if (a_start > 0) e.del_pos.push_back(0), e.del_len.push_back((int32_t)a_start);
    if (a_end < rlen) e.del_pos.push_back((int32_t)a_end), e.del_len.push_back((int32_t)(rlen - a_end));
  } else e.del_pos.push_back(0), e.del_len.push_back(rlen);
}

}  // namespace na
}  // namespace pgmm
