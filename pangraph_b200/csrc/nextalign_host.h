// Host half of K6 (map_variations): base codes, and the conversion of the traceback's run list into the reference's Edit
// (insertions_strip + find_nuc_changes + the terminal deletions of align_with_nextclade + map_variations' position shift).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace pgmm {
namespace na {

struct Edit {
  std::vector<int32_t> sub_pos;
  std::string sub_chr;
  std::vector<int32_t> del_pos, del_len;  // inner deletions ascending, then the leading, then the trailing one
  std::vector<int32_t> ins_pos, ins_len;  // position AFTER the insertion (map_variations.rs:73), ascending
  std::string ins_seq;                    // the inserted bases, insertion after insertion
  int32_t status = 0, hit_boundary = 0, attempts = 0, band_width = 0, score = 0;
};

// to_nuc (PG/align/nextclade/alphabet/nuc.rs:100-121): 0..14 for "TAWCYMHGKRDSBVN"; returns false on any other character
// (a gap in an input sequence is rejected too: the aligner is only ever given ungapped sequences)
bool encode(const char *s, int64_t n, uint8_t *out);

// runs: (len << 2 | op) in walk order (back to front); op 0 = match column, 1 = query base against a reference gap,
// 2 = reference base against a query gap
void edit_from_runs(const char *ref, int32_t rlen, const char *qry, int32_t qlen, const uint32_t *runs, int64_t n_runs, Edit &e);

}  // namespace na
}  // namespace pgmm
