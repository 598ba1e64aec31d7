// mm_mapq.h -- K6: mapping qualities.
// Replaces mapWrap::addMappingQualities + likelihood_observed_set_sizes (reference src/map/mapWrap.h:215-356):
// per read, idmax = exp(-(1 - max_i identity_i)); n = len-k+1; E = round(idmax^k * n); p = E/(2n-E);
// L_i = BinomialPmf(shared_i; sketch_i, p); mapq_i = L_i / sum_j L_j.
// Boost's pdf is replaced by the saddle-point pmf (C. Loader 2000) in double precision on the device;
// parity bar for this column is 1e-6 absolute (BASELINE.json north_star).
#pragma once
#include "mm_platform.h"
#include <cmath>

namespace mm {

MM_HD double d_stirlerr(double n) {
  const double S0 = 1.0 / 12, S1 = 1.0 / 360, S2 = 1.0 / 1260, S3 = 1.0 / 1680, S4 = 1.0 / 1188;
  if (n <= 15.0) {
    double lf = 0; for (int i = 2; i <= (int)n; i++) lf += log((double)i);
    return lf - (n + 0.5) * log(n) + n - 0.918938533204672741780329736406;
  }
  double nn = n * n;
  if (n > 500) return (S0 - S1 / nn) / n;
  if (n > 80) return (S0 - (S1 - S2 / nn) / nn) / n;
  if (n > 35) return (S0 - (S1 - (S2 - S3 / nn) / nn) / nn) / n;
  return (S0 - (S1 - (S2 - (S3 - S4 / nn) / nn) / nn) / nn) / n;
}
MM_HD double d_bd0(double x, double np) {
  if (fabs(x - np) < 0.1 * (x + np)) {
    double v = (x - np) / (x + np), s = (x - np) * v, ej = 2 * x * v;
    v = v * v;
    for (int j = 1; j < 1000; j++) { ej *= v; double s1 = s + ej / (2 * j + 1); if (s1 == s) return s1; s = s1; }
    return s;
  }
  return x * log(x / np) + np - x;
}
MM_HD double d_binom_pmf(int k, int n, double p) {
  if (k < 0 || k > n) return 0.0;
  if (p <= 0.0) return k == 0 ? 1.0 : 0.0;
  if (p >= 1.0) return k == n ? 1.0 : 0.0;
  if (n == 0) return 1.0;
  double q = 1.0 - p;
  if (k == 0) return exp(n * (p < 0.1 ? log1p(-p) : log(q)));
  if (k == n) return exp(n * log(p));
  double x = k, N = n;
  double lc = d_stirlerr(N) - d_stirlerr(x) - d_stirlerr(N - x) - d_bd0(x, N * p) - d_bd0(N - x, N * q);
  double lf = 1.837877066409345483560659472811 + log(x) + log1p(-x / N);
  return exp(lc - 0.5 * lf);
}

struct MapqFn {
  const double* identity; const int32_t* shared; const int32_t* sketch; const int32_t* readLen; const int64_t* readOff; int k;
  double* mapq; int32_t* status;
  MM_HD void operator()(int64_t r) const {
    int64_t b = ldg(readOff + r), e = ldg(readOff + r + 1);
    if (e <= b) { status[r] = 0; return; }
    double maxid = -1;
    for (int64_t m = b; m < e; m++) { double v = ldg(identity + m); if (v > maxid) maxid = v; }
    maxid = exp(-(1 - maxid));
    int n_kmers = ldg(readLen + r) - k + 1;
    double surv = pow(maxid, (double)k);
    double E = round(surv * n_kmers);
    double U = n_kmers + (n_kmers - E);
    double p = E / U;
    double sum = 0;
    for (int64_t m = b; m < e; m++) { double l = d_binom_pmf(ldg(shared + m), ldg(sketch + m), p); mapq[m] = l; sum += l; }
    if (!(sum > 0)) { status[r] = 1; return; }
    for (int64_t m = b; m < e; m++) mapq[m] = mapq[m] / sum;
    status[r] = 0;
  }
};

}  // namespace mm
