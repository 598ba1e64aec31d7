// mm_prims.h -- global sort / scan / run-length primitives.
// Device build: CUB (the CUDA toolkit's own primitives; library code, used only for whole-array sorts and
// scans around the hand-written kernels).  MM_HOST_EMU build: the std:: equivalents.
#pragma once
#include "mm_platform.h"

#ifdef MM_HOST_EMU
#include <algorithm>
#include <numeric>
#else
#include <cub/cub.cuh>
#endif

namespace mm {

template <class TI, class TO>
struct CastOp {
  MM_HD TO operator()(const TI& v) const { return (TO)v; }
};

struct Prims {
  Runtime& rt;
  DevBuf<char> tmp;
  explicit Prims(Runtime& r) : rt(r) {}

  // stable LSD radix sort of (key,value) pairs on key bits [0,end_bit)
  template <class K, class V>
  void sort_pairs(const K* kin, K* kout, const V* vin, V* vout, int64_t n, int end_bit = sizeof(K) * 8) {
    if (n <= 0) return;
#ifdef MM_HOST_EMU
    std::vector<int64_t> idx(n);
    std::iota(idx.begin(), idx.end(), 0);
    K mask = end_bit >= (int)sizeof(K) * 8 ? ~K(0) : ((K(1) << end_bit) - 1);
    std::stable_sort(idx.begin(), idx.end(), [&](int64_t a, int64_t b) { return (kin[a] & mask) < (kin[b] & mask); });
    std::vector<K> ks(n); std::vector<V> vs(n);
    for (int64_t i = 0; i < n; i++) { ks[i] = kin[idx[i]]; vs[i] = vin[idx[i]]; }
    std::copy(ks.begin(), ks.end(), kout); std::copy(vs.begin(), vs.end(), vout);
    rt.launches++;
#else
    size_t bytes = 0;
    MM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, kin, kout, vin, vout, n, 0, end_bit, rt.stream));
    tmp.ensure(bytes);
    MM_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, kin, kout, vin, vout, n, 0, end_bit, rt.stream));
    rt.launches += 1 + (end_bit + 7) / 8;
#endif
  }
  template <class K>
  void sort_keys(const K* kin, K* kout, int64_t n, int end_bit = sizeof(K) * 8) {
    if (n <= 0) return;
#ifdef MM_HOST_EMU
    std::vector<K> ks(kin, kin + n);
    K mask = end_bit >= (int)sizeof(K) * 8 ? ~K(0) : ((K(1) << end_bit) - 1);
    std::stable_sort(ks.begin(), ks.end(), [&](K a, K b) { return (a & mask) < (b & mask); });
    std::copy(ks.begin(), ks.end(), kout);
    rt.launches++;
#else
    size_t bytes = 0;
    MM_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, bytes, kin, kout, n, 0, end_bit, rt.stream));
    tmp.ensure(bytes);
    MM_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, bytes, kin, kout, n, 0, end_bit, rt.stream));
    rt.launches += 1 + (end_bit + 7) / 8;
#endif
  }
  // sort keys inside each segment [off[s], off[s+1])
  template <class K>
  void segmented_sort_keys(const K* kin, K* kout, int64_t n, int64_t nseg, const int64_t* off) {
    if (n <= 0 || nseg <= 0) return;
#ifdef MM_HOST_EMU
    std::copy(kin, kin + n, kout);
    for (int64_t s = 0; s < nseg; s++) std::sort(kout + off[s], kout + off[s + 1]);
    rt.launches++;
#else
    // CUB takes int item/segment counts: walk the segments in slices below 2^30 items
    std::vector<int64_t> h((size_t)nseg + 1);
    d2h(rt, h.data(), off, sizeof(int64_t) * h.size());
    int64_t s0 = 0;
    while (s0 < nseg) {
      int64_t s1 = s0 + 1;
      while (s1 < nseg && h[(size_t)s1 + 1] - h[(size_t)s0] < ((int64_t)1 << 30) && s1 - s0 < ((int64_t)1 << 30)) s1++;
      int64_t items = h[(size_t)s1] - h[(size_t)s0];
      if (items >= ((int64_t)1 << 31)) throw Error(-34, "one read has more than 2^31 seed hits");
      if (items > 0) {
        size_t bytes = 0;
        MM_CUDA(cub::DeviceSegmentedSort::SortKeys(nullptr, bytes, kin, kout, (int)items, (int)(s1 - s0), off + s0, off + s0 + 1, rt.stream));
        tmp.ensure(bytes);
        MM_CUDA(cub::DeviceSegmentedSort::SortKeys(tmp.p, bytes, kin, kout, (int)items, (int)(s1 - s0), off + s0, off + s0 + 1, rt.stream));
      }
      s0 = s1;
    }
    rt.launches += 3;
#endif
  }
  // out[i] = sum_{j<i} in[j], i in [0,n), accumulated in TO (call with n = count+1 and in[count] = 0 to get
  // the total in out[count])
  template <class TI, class TO>
  void exclusive_sum(const TI* in, TO* out, int64_t n) {
    if (n <= 0) return;
#ifdef MM_HOST_EMU
    TO acc = 0;
    for (int64_t i = 0; i < n; i++) { TO v = (TO)in[i]; out[i] = acc; acc += v; }
    rt.launches++;
#else
    cub::TransformInputIterator<TO, CastOp<TI, TO>, const TI*> it(in, CastOp<TI, TO>());
    size_t bytes = 0;
    MM_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, it, out, n, rt.stream));
    tmp.ensure(bytes);
    MM_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, it, out, n, rt.stream));
    rt.launches += 1;
#endif
  }
  // run-length encode a sorted array; *n_runs (device scalar) receives the number of runs
  template <class K>
  void rle(const K* in, K* uniq, int32_t* counts, int64_t* n_runs_dev, int64_t n) {
#ifdef MM_HOST_EMU
    int64_t r = 0;
    for (int64_t i = 0; i < n;) {
      int64_t j = i;
      while (j < n && in[j] == in[i]) j++;
      uniq[r] = in[i]; counts[r] = (int32_t)(j - i); r++; i = j;
    }
    *n_runs_dev = r;
    rt.launches++;
#else
    if (n <= 0) { dev_memset(rt, n_runs_dev, 0, sizeof(int64_t)); return; }
    size_t bytes = 0;
    MM_CUDA(cub::DeviceRunLengthEncode::Encode(nullptr, bytes, in, uniq, counts, n_runs_dev, n, rt.stream));
    tmp.ensure(bytes);
    MM_CUDA(cub::DeviceRunLengthEncode::Encode(tmp.p, bytes, in, uniq, counts, n_runs_dev, n, rt.stream));
    rt.launches += 1;
#endif
  }
  template <class T>
  void reduce_max(const T* in, T* out_dev, int64_t n) {
#ifdef MM_HOST_EMU
    T m = in[0];
    for (int64_t i = 1; i < n; i++) m = std::max(m, in[i]);
    *out_dev = m;
    rt.launches++;
#else
    size_t bytes = 0;
    MM_CUDA(cub::DeviceReduce::Max(nullptr, bytes, in, out_dev, n, rt.stream));
    tmp.ensure(bytes);
    MM_CUDA(cub::DeviceReduce::Max(tmp.p, bytes, in, out_dev, n, rt.stream));
    rt.launches += 1;
#endif
  }
  template <class T>
  void reduce_sum(const T* in, T* out_dev, int64_t n) {
#ifdef MM_HOST_EMU
    T m = 0;
    for (int64_t i = 0; i < n; i++) m += in[i];
    *out_dev = m;
    rt.launches++;
#else
    size_t bytes = 0;
    MM_CUDA(cub::DeviceReduce::Sum(nullptr, bytes, in, out_dev, n, rt.stream));
    tmp.ensure(bytes);
    MM_CUDA(cub::DeviceReduce::Sum(tmp.p, bytes, in, out_dev, n, rt.stream));
    rt.launches += 1;
#endif
  }
};

}  // namespace mm
