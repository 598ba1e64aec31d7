// mm_em.h -- K7 (EM iterations) and K8 (final posterior pass).
// Replaces the loop of meta::doEM (reference src/meta/fEM.h:491-661), the per-read posterior of
// getMappingLocations (:350-363) and the final pass + getBestMapping (:693-716, :217-232).
//
// Data layout: mappings are read-major (the order of the mappings file).  weight[m] = (1/nloc[m]) * mapq[m]
// is fixed across iterations, so an iteration is
//   E-step  (read-major)   rsum[r] = sum_m f[taxon[m]] * weight[m];  ll += log rsum[r]
//   M-step  (taxon-major)  fnext[t] = f[t] * sum_{m in t} weight[m] / rsum[read[m]]
// The taxon-major copy (weightT, readT, sorted once by taxon) turns the taxon-count reduction into a
// segmented sum over contiguous memory instead of M scattered atomics.
// Algorithmic bytes per iteration: 12 B/mapping read-major + 12 B/mapping taxon-major + 8 B gather.
#pragma once
#include "mm_prims.h"
#include "mm_sketch.h"
#include <cmath>

namespace mm {

struct EmPrepFn {
  const double* mapq; const double* nloc; const int64_t* readOff; int64_t n_reads; double* weight; int32_t* readOf; uint32_t* iota;
  MM_HD void operator()(int64_t m) const {
    weight[m] = (1 / ldg(nloc + m)) * ldg(mapq + m);
    readOf[m] = (int32_t)(upper_bound_idx(readOff, n_reads + 1, m) - 1);
    iota[m] = (uint32_t)m;
  }
};
struct EmPermuteFn {
  const uint32_t* perm; const double* weight; const int32_t* readOf; double* weightT; int32_t* readT;
  MM_HD void operator()(int64_t e) const { uint32_t m = ldg(perm + e); weightT[e] = ldg(weight + m); readT[e] = ldg(readOf + m); }
};
struct EmFillFn { double* f; double v; MM_HD void operator()(int64_t t) const { f[t] = v; } };

struct EmReadSumFn {
  const int32_t* taxon; const double* weight; const int64_t* readOff; const double* f; double* rsum; double* logsum;
  MM_HD void operator()(int64_t r) const {
    double s = 0;
    for (int64_t m = ldg(readOff + r), e = ldg(readOff + r + 1); m < e; m++) s += ldg(f + ldg(taxon + m)) * ldg(weight + m);
    rsum[r] = s; logsum[r] = log(s);
  }
};
// one item = EM_TILE consecutive taxon-major entries; partial sums flushed when the taxon changes
static const int EM_TILE = 32;
struct EmTaxonSumFn {
  const uint32_t* taxonT; const double* weightT; const int32_t* readT; const double* rsum; double* acc; int64_t M;
  MM_HD void operator()(int64_t tile) const {
    int64_t b = tile * EM_TILE, e = b + EM_TILE; if (e > M) e = M;
    uint32_t cur = ldg(taxonT + b); double s = 0;
    for (int64_t i = b; i < e; i++) {
      uint32_t t = ldg(taxonT + i);
      if (t != cur) { atomic_add(acc + cur, s); cur = t; s = 0; }
      s += ldg(weightT + i) / ldg(rsum + ldg(readT + i));
    }
    atomic_add(acc + cur, s);
  }
};
struct EmScaleFn {      // fnext[t] = f[t] * acc[t]   (acc then holds fnext for the all-reduce)
  const double* f; double* acc;
  MM_HD void operator()(int64_t t) const { acc[t] = ldg(f + t) * acc[t]; }
};
struct EmNormFn {
  const double* fnext; const double* total; double* f;
  MM_HD void operator()(int64_t t) const { f[t] = ldg(fnext + t) / ldg(total); }
};
struct EmFinalFn {      // fEM.h:693-716 + getBestMapping :217-232 (first maximum wins)
  const int32_t* taxon; const double* weight; const int64_t* readOff; const double* f; double* posterior; int64_t* best;
  MM_HD void operator()(int64_t r) const {
    int64_t b = ldg(readOff + r), e = ldg(readOff + r + 1);
    double s = 0;
    for (int64_t m = b; m < e; m++) s += ldg(f + ldg(taxon + m)) * ldg(weight + m);
    double maxp = 0; int64_t bi = b;
    for (int64_t m = b; m < e; m++) {
      double p = (ldg(f + ldg(taxon + m)) * ldg(weight + m)) / s;
      posterior[m] = p;
      if (m == b || p > maxp) { maxp = p; bi = m; }
    }
    best[r] = bi;
  }
};

}  // namespace mm
