// Byte-level helpers for the text formats on the host (FASTA, Newick): Rust's char::is_whitespace and str::trim on UTF-8 bytes.
#pragma once
#include <cstddef>
#include <string>

namespace pgmm {
namespace text {

// bytes of the White_Space character starting at s[pos] (0 = not whitespace)
inline size_t ws_len(const char *s, size_t pos, size_t n) {
  const unsigned char c = (unsigned char)s[pos];
  if (c == ' ' || (c >= 0x09 && c <= 0x0d)) return 1;
  const auto at = [&](size_t k) { return pos + k < n ? (unsigned char)s[pos + k] : 0u; };
  if (c == 0xc2 && (at(1) == 0x85 || at(1) == 0xa0)) return 2;
  if (c == 0xe1 && at(1) == 0x9a && at(2) == 0x80) return 3;
  if (c == 0xe2) {
    const unsigned d = at(1), e = at(2);
    if (d == 0x80 && ((e >= 0x80 && e <= 0x8a) || e == 0xa8 || e == 0xa9 || e == 0xaf)) return 3;
    if (d == 0x81 && e == 0x9f) return 3;
  }
  if (c == 0xe3 && at(1) == 0x80 && at(2) == 0x80) return 3;
  return 0;
}
// bytes of the UTF-8 character starting at s[pos]
inline size_t char_len(const char *s, size_t pos, size_t n) {
  const unsigned char c = (unsigned char)s[pos];
  const size_t len = c < 0x80 ? 1 : c < 0xe0 ? 2 : c < 0xf0 ? 3 : 4;
  return pos + len <= n ? len : n - pos;
}
// str::trim: [begin, end) without the whitespace at both ends
inline void trim(const char *s, size_t n, size_t &begin, size_t &end) {
  begin = 0, end = n;
  while (begin < end) {
    const size_t w = ws_len(s, begin, end);
    if (!w) break;
    begin += w;
  }
  while (end > begin) {  // step back over one character (1-3 bytes) if it is whitespace
    size_t w = 0;
    for (size_t back = 1; back <= 3 && back <= end - begin; ++back)
      if (ws_len(s, end - back, end) == back) {
        w = back;
        break;
      }
    if (!w) break;
    end -= w;
  }
}

}  // namespace text
}  // namespace pgmm
