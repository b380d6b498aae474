// FASTA input as `pangraph build` reads it (the input half).  Reference (PG = packages/pangraph/src):
// PG/io/fasta.rs:51-225 (FastaReader::read / read_many / from_paths), PG/io/compression.rs:48-69 (compression by extension).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace pgmm {
namespace fasta {

struct Record {  // FastaRecord (fasta.rs:17-24)
  std::string name, desc, seq;
  bool has_desc = false;
  int64_t index = 0;
};

// The default alphabet (Alphabet::DnaWithoutGap, fasta.rs:273-277) and the one with '-' (:279-283)
extern const char *const kDna;
extern const char *const kDnaWithGap;

// read_many over one buffer: every record until the end of input (or until the first record that is_empty(), as the
// reference's loop does).  `alphabet` = the accepted characters after upper-casing.  -> true, or false with the reference's
// message chain in `err`.
bool read_buffer(const char *data, size_t n, const char *alphabet, std::vector<Record> &out, std::string &err);
// from_paths + read_many: the files one after the other, separated by a newline; ".gz" is inflated (any number of members),
// ".bz2" / ".xz" / ".zst" are refused (this build links zlib only)
bool read_files(const std::vector<std::string> &paths, const char *alphabet, std::vector<Record> &out, std::string &err);

}  // namespace fasta
}  // namespace pgmm
