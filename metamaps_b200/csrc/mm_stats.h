// mm_stats.h -- host-side statistics of the mapper (C++, glibc libm: the float/double mix below is what
// fixes the printed identities, so it must run on the host exactly like the reference does).
//
// Replaces skch::Stat (reference src/map/include/map_stats.hpp:44-256) and the Boost.Math calls under it.
// Boost is not vendored by the reference and its version is unpinned; the two functions needed are restated
// from their published definitions:
//   quantile(complement(binomial(n,p), q))  -- default policy integer_round_outwards: smallest integer x
//                                              with P(X > x) <= q
//   pdf(binomial(n,p), k)                   -- binomial pmf (saddle-point form, C. Loader 2000)
// The GPU never evaluates these: per-sketch-size tables (minimumHits[s], acceptMin[s]) are built here and
// uploaded, and K6 has its own device pmf (mm_mapq.h).
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>
#include <algorithm>

namespace mm { namespace stats {

inline double stirlerr(double n) {
  const double S0 = 1.0 / 12, S1 = 1.0 / 360, S2 = 1.0 / 1260, S3 = 1.0 / 1680, S4 = 1.0 / 1188;
  if (n <= 15.0) {
    double lf = 0; for (int i = 2; i <= (int)n; i++) lf += std::log((double)i);
    return lf - (n + 0.5) * std::log(n) + n - 0.918938533204672741780329736406;
  }
  double nn = n * n;
  if (n > 500) return (S0 - S1 / nn) / n;
  if (n > 80) return (S0 - (S1 - S2 / nn) / nn) / n;
  if (n > 35) return (S0 - (S1 - (S2 - S3 / nn) / nn) / nn) / n;
  return (S0 - (S1 - (S2 - (S3 - S4 / nn) / nn) / nn) / nn) / n;
}
inline double bd0(double x, double np) {
  if (std::fabs(x - np) < 0.1 * (x + np)) {
    double v = (x - np) / (x + np), s = (x - np) * v, ej = 2 * x * v;
    v = v * v;
    for (int j = 1; j < 1000; j++) { ej *= v; double s1 = s + ej / (2 * j + 1); if (s1 == s) return s1; s = s1; }
    return s;
  }
  return x * std::log(x / np) + np - x;
}
inline double binom_pmf(int k, int n, double p) {
  if (k < 0 || k > n) return 0.0;
  if (p <= 0.0) return k == 0 ? 1.0 : 0.0;
  if (p >= 1.0) return k == n ? 1.0 : 0.0;
  if (n == 0) return 1.0;
  double q = 1.0 - p;
  if (k == 0) return std::exp(n * (p < 0.1 ? std::log1p(-p) : std::log(q)));
  if (k == n) return std::exp(n * std::log(p));
  double x = k, N = n;
  double lc = stirlerr(N) - stirlerr(x) - stirlerr(N - x) - bd0(x, N * p) - bd0(N - x, N * q);
  double lf = 1.837877066409345483560659472811 + std::log(x) + std::log1p(-x / N);
  return std::exp(lc - 0.5 * lf);
}
// P(X > k): terms summed downwards from far in the upper tail with the pmf recurrence
inline double binom_sf(int k, int n, double p) {
  if (k < 0) return 1.0;
  if (k >= n) return 0.0;
  if (p <= 0) return 0.0;
  if (p >= 1) return 1.0;
  double mean = n * p, sd = std::sqrt(n * p * (1 - p));
  if (k + 1 >= mean) {
    int hi = (int)std::min((double)n, std::floor(std::max(mean, (double)k) + 14 * sd + 14));
    if (hi <= k) return 0.0;
    double pm = binom_pmf(hi, n, p), s = 0, r = (1 - p) / p;
    for (int j = hi; j > k; j--) { s += pm; pm *= (double)j / (double)(n - j + 1) * r; }
    return std::min(1.0, s);
  }
  int lo = (int)std::max(0.0, std::ceil(std::min(mean, (double)k) - 14 * sd - 14));
  double pm = binom_pmf(lo, n, p), s = 0, r = p / (1 - p);
  for (int j = lo; j <= k; j++) { s += pm; pm *= (double)(n - j) / (double)(j + 1) * r; }
  return std::max(0.0, 1.0 - s);
}
// smallest integer x with P(X > x) <= q
inline int binom_quantile_upper(int n, double p, double q) {
  if (p <= 0.0) return 0;
  if (p >= 1.0) return n;
  if (q <= 0.0) return n;
  if (q >= 1.0) return 0;
  double mean = n * p, sd = std::sqrt(n * p * (1 - p));
  int x = (int)std::min((double)n, std::floor(mean + 14 * sd + 14));
  double tail = 0.0;
  double pm = binom_pmf(x, n, p), r = (1 - p) / p;
  while (x > 0) {
    double t2 = tail + pm;                 // P(X > x-1)
    if (t2 <= q) { tail = t2; pm *= (double)x / (double)(n - x + 1) * r; x--; } else break;
  }
  return x;
}

inline float j2md(float j, int k) {                          // map_stats.hpp:44-54
  if (j == 0) return 1.0;
  if (j == 1) return 0.0;
  float d = (-1.0 / k) * log(2.0 * j / (1 + j));
  return d;
}
inline float md2j(float d, int k) { float j = 1.0 / (2.0 * exp(k * d) - 1.0); return j; }   // :62-66
inline float md_lower_bound(float d, int s, int k, float ci) {                                // :79-111
  float q2 = (1.0 - ci) / 2;
  int x = binom_quantile_upper(s, (double)md2j(d, k), (double)q2);
  float jaccard = float(x) / s;
  return j2md(jaccard, k);
}
inline int estimateMinimumHits(int s, int k, float pi) {                                      // :120-131
  float md = 1.0 - pi / 100.0;
  float j = md2j(md, k);
  return (int)ceil(1.0 * s * j);
}
inline float identity_upper(int shared, int s, int k) {
  float d = j2md(1.0 * shared / s, k);
  float dl = md_lower_bound(d, s, k, 0.9);
  return 100.0 * (1.0 - dl);
}
inline int estimateMinimumHitsRelaxed(int s, int k, float pi) {                               // :142-167
  int first = estimateMinimumHits(s, k, pi);
  int relaxed = first;
  for (int i = first; i >= 0; i--) {
    if (identity_upper(i, s, k) >= pi) relaxed = i; else break;
  }
  return relaxed;
}
// smallest shared count whose identity upper bound passes the filter of computeMap.hpp:415.
// identity_upper is non-decreasing in `shared` (each step of shared moves the jaccard by 1/s, far more than
// any rounding), so this is a threshold; s+1 means "never".
inline int acceptMinimum(int s, int k, float pi) {
  int first = std::min(estimateMinimumHits(s, k, pi), s);
  if (first < 0) first = 0;
  if (identity_upper(first, s, k) >= pi) {
    int x = first;
    while (x > 0 && identity_upper(x - 1, s, k) >= pi) x--;
    return x;
  }
  for (int x = first + 1; x <= s; x++) if (identity_upper(x, s, k) >= pi) return x;
  return s + 1;
}
inline void identity(int shared, int s, int k, float* nuc, float* upper) {                    // computeMap.hpp:405-411
  float md = j2md(1.0 * shared / s, k);
  float lb = md_lower_bound(md, s, k, 0.9);
  *nuc = 100 * (1 - md);
  *upper = 100 * (1 - lb);
}
inline void identity_only(int shared, int s, int k, float* nuc) {                             // computeMap.hpp:403-408
  float md = j2md(1.0 * shared / s, k);
  *nuc = 100 * (1 - md);
}
inline double estimate_pvalue(int s, int k, int alphabet, float identity_, int lenQ, uint64_t lenR) {  // :179-213
  double kmerSpace = pow(alphabet, k);
  double pX, pY; pX = pY = 1. / (1. + kmerSpace / lenQ);
  double r = pX * pY / (pX + pY - pX * pY);
  int x = estimateMinimumHitsRelaxed(s, k, identity_);
  double cc = (x == 0) ? 1.0 : binom_sf(x - 1, s, r);
  return lenR * cc;
}
inline int recommendedWindowSize(double pcut, int k, int alphabet, float identity_, int lenQ, uint64_t lenR) {  // :226-256
  std::vector<int> cand{1, 2, 5};
  for (int i = 10; i < lenQ; i += 10) cand.push_back(i);
  int opt = 0;
  for (int e : cand) { if (estimate_pvalue(e, k, alphabet, identity_, lenQ, lenR) <= pcut) { opt = e; break; } }
  if (opt == 0) opt = cand.back();   // the reference leaves this uninitialised; unreachable for sane inputs
  int w = 2.0 * lenQ / opt;
  return std::min(std::max(w, 1), lenQ);
}

// per-sketch-size tables, extended lazily
struct Tables {
  int k = 0; float pi = 0;
  std::vector<int32_t> minHits, acceptMin;     // index s; entry 0 unused
  void extend(int k_, float pi_, int smax) {
    if (k_ != k || pi_ != pi) { k = k_; pi = pi_; minHits.assign(1, 0); acceptMin.assign(1, 1); }
    int old = (int)minHits.size();
    if (smax + 1 <= old) return;
    minHits.resize((size_t)smax + 1); acceptMin.resize((size_t)smax + 1);
#pragma omp parallel for schedule(dynamic, 64)
    for (int s = old; s <= smax; s++) {
      minHits[(size_t)s] = estimateMinimumHitsRelaxed(s, k, pi);
      acceptMin[(size_t)s] = acceptMinimum(s, k, pi);
    }
  }
};

}}  // namespace mm::stats
