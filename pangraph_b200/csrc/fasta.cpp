#include "fasta.h"

#include <zlib.h>

#include <algorithm>
#include <cerrno>
#include <cstring>

#include "text_util.h"

namespace pgmm {
namespace fasta {

const char *const kDna = "ACGTYRWSKMDVHBN";
const char *const kDnaWithGap = "ACGTYRWSKMDVHBN-";

namespace {

// The reader's state between records (fasta.rs:51-58): the header line of the next record once it has been seen.
struct Reader {
  const char *data;
  size_t n, at = 0;
  bool accept[256];
  std::string pending;  // trimmed header line found while reading the previous record's sequence ("" = none)
  size_t n_chars = 0;
  int64_t index = 0;

  Reader(const char *d, size_t len, const char *alphabet) : data(d), n(len) {
    std::fill(accept, accept + 256, false);
    for (const char *a = alphabet; *a; ++a) accept[(unsigned char)*a] = true;
  }
  // read_line + trim: false at the end of input
  bool next_line(size_t &b, size_t &e) {
    if (at >= n) return false;
    const char *nl = (const char *)memchr(data + at, '\n', n - at);
    const size_t stop = nl ? (size_t)(nl - data) + 1 : n;
    size_t tb, te;
    text::trim(data + at, stop - at, tb, te);
    b = at + tb, e = at + te;
    at = stop;
    n_chars += e - b;
    return true;
  }
  // fasta.rs:131-225.  -> 1 a record, 0 the end of input (record left cleared), -1 error
  int read(Record &r, std::string &err) {
    r = Record();
    if (pending.empty()) {
      for (;;) {
        size_t b, e;
        if (!next_line(b, e)) {
          if (index > 0 || n_chars == 0) return 0;  // (an empty or all-whitespace input is allowed: no records)
          err = "FASTA input is incorrectly formatted: expected at least one FASTA record starting with character '>', but none found";
          return -1;
        }
        if (e > b && data[b] == '>') {
          pending.assign(data + b, e - b);
          break;
        }
      }
    }
    const size_t sp = pending.find(' ', 1);
    if (sp == std::string::npos) r.name = pending.substr(1);
    else {
      r.name = pending.substr(1, sp - 1), r.desc = pending.substr(sp + 1);
      r.has_desc = !r.desc.empty();
    }
    r.index = index++;
    pending.clear();
    size_t b, e;
    while (next_line(b, e)) {
      if (e > b && data[b] == '>') {
        pending.assign(data + b, e - b);
        break;
      }
      const size_t base = r.seq.size();
      r.seq.resize(base + (e - b));
      char *dst = &r.seq[0] + base;
      for (size_t i = b; i < e; ++i) {
        unsigned char c = (unsigned char)data[i];
        if (c >= 'a' && c <= 'z') c = (unsigned char)(c - 32);
        if (!accept[c]) {
          const size_t len = text::char_len(data, i, e);
          err = "When processing sequence #" + std::to_string(index) + ": \">" + r.name + (r.has_desc ? " " + r.desc : "") +
                "\": FASTA input is incorrect: character \"" + std::string(data + i, len) + "\" is not in the alphabet";
          return -1;
        }
        dst[i - b] = (char)c;
      }
    }
    return 1;
  }
};

bool is_empty(const Record &r) { return r.name.empty() && r.seq.empty() && !r.has_desc && r.index == 0; }

std::string lower_extension(const std::string &path) {
  const size_t slash = path.find_last_of('/');
  const size_t dot = path.find_last_of('.');
  if (dot == std::string::npos || (slash != std::string::npos && dot < slash) || dot + 1 >= path.size()) return "";
  std::string ext = path.substr(dot + 1);
  for (char &c : ext) c = (char)((c >= 'A' && c <= 'Z') ? c + 32 : c);
  return ext;
}

}  // namespace

bool read_buffer(const char *data, size_t n, const char *alphabet, std::vector<Record> &out, std::string &err) {
  out.clear();
  Reader rd(data, n, alphabet ? alphabet : kDna);
  for (;;) {
    Record r;
    const int rc = rd.read(r, err);
    if (rc < 0) return false;
    if (rc == 0 || is_empty(r)) return true;  // read_many stops at the first empty record (fasta.rs:118-124)
    out.push_back(std::move(r));
  }
}

bool read_files(const std::vector<std::string> &paths, const char *alphabet, std::vector<Record> &out, std::string &err) {
  std::string all;
  for (size_t f = 0; f < paths.size(); ++f) {
    const std::string ext = lower_extension(paths[f]);
    if (ext == "bz2" || ext == "xz" || ext == "zst") {
      err = "When opening file '" + paths[f] + "': ." + ext + " input is not supported by this build (gzip or plain text only)";
      return false;
    }
    gzFile fp = gzopen(paths[f].c_str(), "rb");  // transparent for files that are not gzip
    if (!fp) {
      err = "When opening file '" + paths[f] + "': " + strerror(errno);
      return false;
    }
    gzbuffer(fp, 1 << 20);
    char buf[1 << 16];
    int got;
    while ((got = gzread(fp, buf, sizeof(buf))) > 0) all.append(buf, (size_t)got);
    if (got < 0) {
      int code = 0;
      err = "While decompressing file '" + paths[f] + "': " + gzerror(fp, &code);
      gzclose(fp);
      return false;
    }
    gzclose(fp);
    if (f + 1 < paths.size()) all.push_back('\n');  // Concat::with_delimiter(readers, "\n") (fasta.rs:103)
  }
  return read_buffer(all.data(), all.size(), alphabet, out, err);
}

}  // namespace fasta
}  // namespace pgmm
