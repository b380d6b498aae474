// In-place MSD "American flag" radix sort with the reference's exact permutation behaviour.
//
// The reference sorts anchors, chain end points and hit keys with radix_sort_128x / radix_sort_64
// (packages/minimap2-sys/minimap2/ksort.h:101-151; instantiated in misc.c:155-162): 8-bit digits from the top byte
// down, cycle-leader scatter inside each pass, insertion sort for buckets of <= 64 elements.  The sort is NOT stable
// and the order it leaves among equal keys is observable downstream (chain backtracking, anchor order),
// so this routine reproduces the same sequence of element moves rather than just "a sorted array".
#pragma once
#include <cstddef>
#include <cstdint>

namespace pgmm {

template <class T, class Key>
inline void insertion_sort_by(T *beg, T *end, Key key) {
  for (T *i = beg + 1; i < end; ++i) {
    if (key(*i) < key(*(i - 1))) {
      T tmp = *i;
      T *j = i;
      for (; j > beg && key(tmp) < key(*(j - 1)); --j) *j = *(j - 1);
      *j = tmp;
    }
  }
}

template <class T, class Key>
void flag_sort_pass(T *beg, T *end, int shift, Key key) {
  constexpr int kDigits = 256;
  constexpr ptrdiff_t kSmall = 64;
  T *head[kDigits], *tail[kDigits];
  size_t count[kDigits] = {0};
  uint64_t all_or = 0, all_and = ~0ull;
  for (T *i = beg; i != end; ++i) {
    const uint64_t k = key(*i);
    ++count[(k >> shift) & 0xff], all_or |= k, all_and &= k;
  }
  // A digit every key of the range shares puts the whole range into one bucket: the reference's pass moves nothing and
  // recurses on the same range with the next digit.  Such passes (the zero bytes above a chain score, the bytes between
  // the strand / target id and the position of an anchor) are skipped in one step.
  if (const uint64_t varies = all_or ^ all_and; ((varies >> shift) & 0xff) == 0) {
    int s2 = shift;
    while (s2 > 0 && ((varies >> s2) & 0xff) == 0) s2 -= 8;
    if (s2 < 0) s2 = 0;
    if (((varies >> s2) & 0xff) != 0) flag_sort_pass(beg, end, s2, key);
    return;
  }
  {
    T *p = beg;
    for (int d = 0; d < kDigits; ++d) head[d] = p, p += count[d], tail[d] = p;
  }
  // cycle-leader scatter: take the first unplaced element of the lowest unfinished bucket and push it (and whatever
  // it displaces) home until an element that belongs to this bucket comes back
  for (int d = 0; d < kDigits;) {
    if (head[d] == tail[d]) {
      ++d;
      continue;
    }
    int home = (int)((key(*head[d]) >> shift) & 0xff);
    if (home == d) {
      ++head[d];
      continue;
    }
    T carry = *head[d];
    do {
      T displaced = *head[home];
      *head[home]++ = carry;
      carry = displaced;
      home = (int)((key(carry) >> shift) & 0xff);
    } while (home != d);
    *head[d]++ = carry;
  }
  if (shift == 0) return;
  const int next = shift > 8 ? shift - 8 : 0;
  T *b = beg;
  for (int d = 0; d < kDigits; ++d) {
    T *e = tail[d];
    if (e - b > kSmall) flag_sort_pass(b, e, next, key);
    else if (e - b > 1) insertion_sort_by(b, e, key);
    b = e;
  }
}

// key(T) must return uint64_t
template <class T, class Key>
inline void flag_sort(T *beg, T *end, Key key) {
  if (end - beg <= 64) insertion_sort_by(beg, end, key);
  else flag_sort_pass(beg, end, 56, key);
}

struct U128 {
  uint64_t x, y;
};
inline void flag_sort_128x(U128 *beg, U128 *end) {
  flag_sort(beg, end, [](const U128 &a) { return a.x; });
}
inline void flag_sort_64(uint64_t *beg, uint64_t *end) {
  flag_sort(beg, end, [](uint64_t a) { return a; });
}

}  // namespace pgmm
