// mm_stdsort.h -- a step-for-step replay of libstdc++'s std::sort (introsort: median-of-3 quicksort to
// depth 2*lg(n), heapsort fallback, final insertion sort; bits/stl_algo.h, bits/stl_heap.h, unchanged since
// GCC 4.x) for keys compared by their high 32 bits only.
//
// Why: the reference sorts a read's minimizers with the UNSTABLE std::sort by hash and then std::unique
// keeps the first of every equal-hash run (src/map/include/computeMap.hpp:292-295).  When a read contains
// the same hash twice with different strands, which copy survives -- and therefore strandQ in the L2
// strand vote (slidingMap.hpp:244) -- is decided by this exact permutation.  Only such reads (a handful per
// million) are replayed, one thread per read, on the device.
#pragma once
#include "mm_platform.h"

namespace mm {
namespace stdsort {

MM_HD bool lt(uint64_t a, uint64_t b) { return (uint32_t)(a >> 32) < (uint32_t)(b >> 32); }
MM_HD void swp(uint64_t* a, int64_t i, int64_t j) { uint64_t t = a[i]; a[i] = a[j]; a[j] = t; }

MM_HD void push_heap(uint64_t* f, int64_t hole, int64_t top, uint64_t v) {
  int64_t parent = (hole - 1) / 2;
  while (hole > top && lt(f[parent], v)) { f[hole] = f[parent]; hole = parent; parent = (hole - 1) / 2; }
  f[hole] = v;
}
MM_HD void adjust_heap(uint64_t* f, int64_t hole, int64_t len, uint64_t v) {
  const int64_t top = hole;
  int64_t child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (lt(f[child], f[child - 1])) child--;
    f[hole] = f[child]; hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    f[hole] = f[child - 1]; hole = child - 1;
  }
  push_heap(f, hole, top, v);
}
MM_HD void heap_sort(uint64_t* f, int64_t len) {          // __partial_sort(first, last, last)
  if (len >= 2) {
    int64_t parent = (len - 2) / 2;
    for (;;) { uint64_t v = f[parent]; adjust_heap(f, parent, len, v); if (parent == 0) break; parent--; }
  }
  int64_t last = len;
  while (last > 1) { --last; uint64_t v = f[last]; f[last] = f[0]; adjust_heap(f, 0, last, v); }
}
MM_HD void move_median_to_first(uint64_t* a, int64_t r, int64_t x, int64_t y, int64_t z) {
  if (lt(a[x], a[y])) {
    if (lt(a[y], a[z])) swp(a, r, y);
    else if (lt(a[x], a[z])) swp(a, r, z);
    else swp(a, r, x);
  } else if (lt(a[x], a[z])) swp(a, r, x);
  else if (lt(a[y], a[z])) swp(a, r, z);
  else swp(a, r, y);
}
MM_HD int64_t unguarded_partition(uint64_t* a, int64_t first, int64_t last, int64_t pivot) {
  for (;;) {
    while (lt(a[first], a[pivot])) ++first;
    --last;
    while (lt(a[pivot], a[last])) --last;
    if (!(first < last)) return first;
    swp(a, first, last);
    ++first;
  }
}
MM_HD void unguarded_linear_insert(uint64_t* a, int64_t last) {
  uint64_t v = a[last]; int64_t next = last - 1;
  while (lt(v, a[next])) { a[last] = a[next]; last = next; --next; }
  a[last] = v;
}
MM_HD void insertion_sort(uint64_t* a, int64_t first, int64_t last) {
  if (first == last) return;
  for (int64_t i = first + 1; i != last; ++i) {
    if (lt(a[i], a[first])) { uint64_t v = a[i]; for (int64_t j = i; j > first; --j) a[j] = a[j - 1]; a[first] = v; }
    else unguarded_linear_insert(a, i);
  }
}
// std::sort(a, a+n, lessByHash)
MM_HD void sort(uint64_t* a, int64_t n) {
  if (n <= 0) return;
  int lg = 0; for (int64_t t = n; t > 1; t >>= 1) lg++;
  // explicit stack for the recursive call on [cut,last)
  int64_t sf[128], sl[128]; int sd[128]; int sp = 0;
  sf[0] = 0; sl[0] = n; sd[0] = lg * 2; sp = 1;
  while (sp > 0) {
    --sp;
    int64_t first = sf[sp], last = sl[sp]; int depth = sd[sp];
    while (last - first > 16) {
      if (depth == 0) { heap_sort(a + first, last - first); break; }
      --depth;
      int64_t mid = first + (last - first) / 2;
      move_median_to_first(a, first, first + 1, mid, last - 1);
      int64_t cut = unguarded_partition(a, first + 1, last, first);
      // the real code recurses into [cut,last) FIRST and then loops on [first,cut); the two ranges are
      // disjoint, so deferring the right part on a stack gives the same final array
      if (sp < 128) { sf[sp] = cut; sl[sp] = last; sd[sp] = depth; sp++; }
      last = cut;
    }
  }
  if (n > 16) { insertion_sort(a, 0, 16); for (int64_t i = 16; i != n; ++i) unguarded_linear_insert(a, i); }
  else insertion_sort(a, 0, n);
}

}  // namespace stdsort
}  // namespace mm
