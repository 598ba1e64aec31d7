// oracle/mm_refmath.h -- TEST INFRASTRUCTURE ONLY (oracle); never linked into the product.
//
// CPU restatement of the Boost.Math calls MetaMaps makes on the hot path.  Boost is a
// third-party dependency that is NOT vendored under /root/reference and whose version is
// unpinned (configure.ac:17,50 only takes --with-boost=<path>), so the published
// definitions are restated here:
//   * binomial pdf        -- mapWrap.h:340      boost::math::pdf(binomial_distribution<>(n,p), k)
//   * binomial upper quantile -- map_stats.hpp:88  quantile(complement(binomial(n,p), q))
//       default policy integer_round_outwards => an upper quantile is rounded UP:
//       the smallest integer x with sf(x) = P(X > x) <= q.
//   * binomial cdf / complement cdf -- map_stats.hpp:204, fEM.h:1107
//   * poisson pdf (fEM.h:1101), chi_squared(1) cdf (fEM.h:1064)
// Pinned against SciPy's Boost-backed ufuncs (_binom_pmf/_binom_isf/_binom_sf) by
// tests/golden/boost_binomial.json (generator: tests/golden/make_boost_golden.py).
#pragma once
#include <cmath>
#include <algorithm>

namespace mmref {

// Saddle-point binomial pmf (C. Loader, "Fast and accurate computation of binomial
// probabilities", 2000).  Accurate to a few ulp for all n used here.
inline double stirlerr(double n) {
  const double S0 = 1.0 / 12, S1 = 1.0 / 360, S2 = 1.0 / 1260, S3 = 1.0 / 1680, S4 = 1.0 / 1188;
  if (n <= 15.0) {
    // exact log-factorial minus Stirling's leading terms
    double lf = 0; for (int i = 2; i <= (int)n; i++) lf += std::log((double)i);
    return lf - (n + 0.5) * std::log(n) + n - 0.918938533204672741780329736406; // 0.5*log(2*pi)
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
    for (int j = 1; j < 1000; j++) {
      ej *= v;
      double s1 = s + ej / (2 * j + 1);
      if (s1 == s) return s1;
      s = s1;
    }
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
  double lf = 1.837877066409345483560659472811 + std::log(x) + std::log1p(-x / N); // log(2*pi)
  return std::exp(lc - 0.5 * lf);
}
// P(X <= k)
inline double binom_cdf(int k, int n, double p) {
  if (k < 0) return 0.0;
  if (k >= n) return 1.0;
  // sum the smaller side for accuracy
  double mean = n * p;
  if (k < mean) { double s = 0; for (int j = 0; j <= k; j++) s += binom_pmf(j, n, p); return std::min(1.0, s); }
  double s = 0; for (int j = n; j > k; j--) s += binom_pmf(j, n, p); return std::max(0.0, 1.0 - s);
}
// P(X > k)
inline double binom_sf(int k, int n, double p) {
  if (k < 0) return 1.0;
  if (k >= n) return 0.0;
  double mean = n * p;
  if (k >= mean) { double s = 0; for (int j = n; j > k; j--) s += binom_pmf(j, n, p); return std::min(1.0, s); }
  double s = 0; for (int j = 0; j <= k; j++) s += binom_pmf(j, n, p); return std::max(0.0, 1.0 - s);
}
// smallest integer x with sf(x) <= q   (Boost: quantile(complement(binomial(n,p), q)))
inline double binom_quantile_upper(int n, double p, double q) {
  if (p <= 0.0) return 0;
  if (p >= 1.0) return n;
  if (q <= 0.0) return n;
  if (q >= 1.0) return 0;
  // walk down from n accumulating the upper tail
  double tail = 0.0;            // sf(x) for the current x (starts at x = n)
  int x = n;
  while (x > 0) {
    double t2 = tail + binom_pmf(x, n, p);   // sf(x-1)
    if (t2 <= q) { tail = t2; x--; } else break;
  }
  return x;
}
inline double poisson_pmf(int k, double lambda) {
  if (lambda <= 0) return k == 0 ? 1.0 : 0.0;
  if (k == 0) return std::exp(-lambda);
  return std::exp(-stirlerr(k) - bd0(k, lambda)) / std::sqrt(6.283185307179586476925286766559 * k);
}
inline double chi2_1df_cdf(double x) { return x <= 0 ? 0.0 : std::erf(std::sqrt(x / 2.0)); }

} // namespace mmref
