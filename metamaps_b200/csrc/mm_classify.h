// mm_classify.h -- the classify stage on the device: accepted mappings -> identity -> K6 mapping quality -> nLoc -> K7/K8 EM.
//
// Replaces, on arrays that never leave HBM between the stages:
//   Map::doL2Mapping's nucIdentity                       computeMap.hpp:403-408, map_stats.hpp:44-54
//   mapWrap::addMappingQualities                         mapWrap.h:215-323 (+ likelihood_observed_set_sizes :332-356)
//   meta::getMappingLocations (nLoc, posterior)          fEM.h:234-373
//   meta::doEM loop, final pass, getBestMapping          fEM.h:491-661, :693-716, :217-232
//
// Segmented per-read work (K6, K7, K8) runs with a GROUP of G lanes per read (G = 4 / 8 / 32, picked from the mean number
// of mappings per read): the lanes stride over the read's mappings (coalesced), reductions are xor-butterflies
// (`__shfl_xor_sync`), so every lane of the group ends with the same value and the order of the sum is fixed.
// One EM round = ONE kernel (E-step: per-read likelihood sum, log-likelihood, per-mapping posterior accumulated into taxon
// sums kept in shared memory, flushed once per CTA) + one single-CTA kernel that normalises f and applies the reference's
// stopping rule (fEM.h:636) ON THE DEVICE: the host looks at the `done` flag every few rounds instead of reading the
// log-likelihood back after each one; rounds launched past the stopping round do nothing.
// Algorithmic bytes per round: 12 A (u32 taxon + f64 weight) + 4 per read + 16 T.
//
// The host-emulation build (tests/_emu) runs the same per-read functors with G = 1.
#pragma once
#include "mm_mapq.h"
#include "mm_prims.h"
#include "mm_stats.h"
#include <algorithm>
#include <cmath>
#include <functional>
#include <vector>

namespace mm {

// ---- lane groups ----------------------------------------------------------------------------------------------------
template <int G>
struct Grp {
  MM_HD static int lane() {
#if defined(__CUDA_ARCH__)
    return (int)(threadIdx.x & (G - 1));
#else
    return 0;
#endif
  }
  MM_HD static double sum(double v) {
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
#endif
    return v;
  }
  MM_HD static double max(double v) {
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) { const double u = __shfl_xor_sync(0xffffffffu, v, o); v = u > v ? u : v; }
#endif
    return v;
  }
  // (largest p, smallest index among equal p): getBestMapping keeps the FIRST maximum (fEM.h:217-232)
  MM_HD static void argmax_first(double& p, int64_t& i) {
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
      const double p2 = __shfl_xor_sync(0xffffffffu, p, o); const long long i2 = __shfl_xor_sync(0xffffffffu, (long long)i, o);
      if (p2 > p || (p2 == p && i2 < i)) { p = p2; i = i2; }
    }
#endif
  }
  MM_HD static int any(int v) {
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, o);
#endif
    return v;
  }
};
#ifdef MM_HOST_EMU
static const int GRP_HOST = 1;
#endif
inline int64_t round_up32(int64_t n) { return (n + 31) / 32 * 32; }

// ---- the accepted mappings of a batch, device resident ---------------------------------------------------------------
struct MapTable {
  DevBuf<int32_t> read, seq, pos, shared, sketch, strand;      // per mapping, reference order: read, then (contig, position)
  int64_t n = 0; int parts = 0; bool sorted = true;
  void reserve(Runtime& rt, int64_t want) {
    if ((size_t)want <= read.cap) return;
    read.grow(rt, (size_t)want, (size_t)n); seq.grow(rt, (size_t)want, (size_t)n); pos.grow(rt, (size_t)want, (size_t)n);
    shared.grow(rt, (size_t)want, (size_t)n); sketch.grow(rt, (size_t)want, (size_t)n); strand.grow(rt, (size_t)want, (size_t)n);
  }
};

// candidate -> mapping compaction (the lines of reportReadMappings with --all, computeMap.hpp:546-588)
struct MapAppendFn {
  const int32_t* oAccept; const int64_t* accIdx; const int32_t* cRead; const int32_t* cSeq; const int32_t* oPos; const int32_t* oShared;
  const int32_t* oVotes; const int32_t* sOf; int32_t seqBase; int32_t readBase; int64_t base;
  int32_t* mRead; int32_t* mSeq; int32_t* mPos; int32_t* mShared; int32_t* mSketch; int32_t* mStrand;
  MM_HD void operator()(int64_t c) const {
    if (!ldg(oAccept + c)) return;
    const int64_t d = base + ldg(accIdx + c); const int32_t r = ldg(cRead + c);
    mRead[d] = r + readBase; mSeq[d] = ldg(cSeq + c) + seqBase; mPos[d] = ldg(oPos + c); mShared[d] = ldg(oShared + c); mSketch[d] = ldg(sOf + r);
    mStrand[d] = ldg(oVotes + c) > 0 ? 1 : -1;                         // computeMap.hpp:438
  }
};
struct GatherI32Fn { const uint32_t* perm; const int32_t* in; int32_t* out; MM_HD void operator()(int64_t i) const { out[i] = ldg(in + ldg(perm + i)); } };
struct IotaU32Fn { uint32_t* a; MM_HD void operator()(int64_t i) const { a[i] = (uint32_t)i; } };

// ---- identity (computeMap.hpp:403-408 through map_stats.hpp:44-54) ---------------------------------------------------
// float nucIdentity = 100 * (1 - j2md(1.0 * shared / s, k)) with j2md = (-1.0 / k) * log(2.0 * j / (1 + j)) evaluated in
// double and rounded to float.  Every operation but the log is IEEE-exact on both sides; the device log may differ from
// glibc's by an ulp of the DOUBLE, which changes the float only when the product sits within a few double-ulps of a
// float rounding boundary.  Those (about one mapping in 10^7) are flagged and recomputed by the host with glibc, like every
// identity outside the range where the 6-significant-digit round trip of the text file (mapWrap.h:229, fEM.h:297) has a
// closed form.  So the arrays are what the reference's files hold, bit for bit, without a host pass over the batch.
struct IdentityFn {
  const int32_t* shared; const int32_t* sketch; int k; float* identity; double* parsed; unsigned long long* nFix; int32_t* fixList; int64_t fixCap;
  const int32_t* mask;       // optional: entries with mask[m] == 0 are skipped (identity 0)
  MM_HD void operator()(int64_t m) const {
    if (mask && !ldg(mask + m)) { identity[m] = 0; parsed[m] = 0; return; }
    const int32_t sh = ldg(shared + m), s = ldg(sketch + m);
    const float j = (float)(1.0 * sh / s);
    float md; bool unsure = false;
    if (j == 0) md = 1.0f;
    else if (j == 1) md = 0.0f;
    else {
      const double v = (-1.0 / k) * log(2.0 * j / (1 + j));
      md = (float)v;
      unsure = (float)(v * (1.0 + 0x1p-47)) != (float)(v * (1.0 - 0x1p-47));
    }
    const float id = 100 * (1 - md);
    const double x = (double)id;
    double p;
    if (x >= 10.0 && x < 99.99995) p = rint(x * 1e4) / 1e4;       // "%.6g" of [10,100) = 4 decimals; x*1e4 is exact in double
    else if (x == 100.0) p = 100.0;
    else { p = x; unsure = true; }
    identity[m] = id; parsed[m] = p;
    if (unsure) { const unsigned long long slot = atomic_add_u64(nFix, 1ull); if ((int64_t)slot < fixCap) fixList[slot] = (int32_t)m; }
  }
};

// ---- read groups -----------------------------------------------------------------------------------------------------
struct GroupHeadFn { const int32_t* mRead; int32_t* head; int64_t n; MM_HD void operator()(int64_t m) const { head[m] = (m < n && (m == 0 || ldg(mRead + m) != ldg(mRead + m - 1))) ? 1 : 0; } };
struct GroupScatterFn {
  const int32_t* mRead; const int32_t* head; const int64_t* gidx; int64_t n; int64_t* grpOff; int32_t* grpRead; int32_t* mGrp;
  MM_HD void operator()(int64_t m) const {
    if (m == n) { grpOff[ldg(gidx + n)] = n; return; }
    const int64_t g = ldg(gidx + m) + ldg(head + m) - 1;
    mGrp[m] = (int32_t)g;
    if (ldg(head + m)) { grpOff[g] = m; grpRead[g] = ldg(mRead + m); }
  }
};
struct GroupLenFn { const int32_t* grpRead; const int32_t* readLen; int32_t* out; MM_HD void operator()(int64_t g) const { out[g] = ldg(readLen + ldg(grpRead + g)); } };

// ---- K6, one group of G lanes per read -------------------------------------------------------------------------------
template <int G>
struct MapqGroupFn {
  const double* parsed; double div;      // identity = parsed / div: column 10 / 100 (mapWrap.h:229), or a fraction given as such (div = 1)
  const int32_t* shared; const int32_t* sketch; const int32_t* grpLen; const int64_t* grpOff; int64_t nGroups; int k;
  double* mapq; int32_t* status;
  MM_HD void operator()(int64_t item) const {       // no lane leaves before the last shuffle of its warp
    const int64_t g = item / G; const int lane = Grp<G>::lane();
    const bool on = g < nGroups;
    const int64_t b = on ? ldg(grpOff + g) : 0, e = on ? ldg(grpOff + g + 1) : 0;
    double maxid = -1;
    for (int64_t m = b + lane; m < e; m += G) { const double v = ldg(parsed + m) / div; if (v > maxid) maxid = v; }
    maxid = Grp<G>::max(maxid);
    const bool live = on && e > b;
    maxid = exp(-(1 - maxid));
    const int n_kmers = (live ? ldg(grpLen + g) : k) - k + 1;
    const double surv = pow(maxid, (double)k);
    const double E = round(surv * n_kmers);
    const double U = n_kmers + (n_kmers - E);
    const double p = E / U;
    double sum = 0;
    for (int64_t m = b + lane; m < e; m += G) { const double l = d_binom_pmf(ldg(shared + m), ldg(sketch + m), p); mapq[m] = l; sum += l; }
    sum = Grp<G>::sum(sum);
    if (!live) { if (on && lane == 0) status[g] = 0; return; }
    if (!(sum > 0)) { if (lane == 0) status[g] = 1; return; }                                   // the reference asserts (mapWrap.h:298)
    for (int64_t m = b + lane; m < e; m += G) mapq[m] = mapq[m] / sum;
    if (lane == 0) status[g] = 0;
  }
};

// ---- nLoc (fEM.h:324-348) --------------------------------------------------------------------------------------------
// taxon t's contigs sorted by length: lens[start[t] .. start[t+1]), csum = prefix sums over the whole lens array
struct NlocFn {
  const int32_t* mSeq; const int32_t* mGrp; const int64_t* grpOff; const int32_t* grpLen; const double* mapq;
  const int64_t* contigLen; const int32_t* contigTaxon; int32_t nContigs; const int64_t* lens; const int64_t* start; const int64_t* csum;
  int32_t* tax; double* nloc; double* weight; int32_t* bad;
  MM_HD void operator()(int64_t m) const {
    const int32_t sq = ldg(mSeq + m);
    if (sq < 0 || sq >= nContigs) { *bad = 1; tax[m] = 0; if (nloc) nloc[m] = 1; if (weight) weight[m] = 0; return; }
    const int32_t t = ldg(contigTaxon + sq); const int32_t g = ldg(mGrp + m);
    const int64_t L = ldg(grpLen + g);
    const int64_t b = ldg(start + t), e = ldg(start + t + 1);
    int64_t lo = b, hi = e;
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (ldg(lens + mid) < L) lo = mid + 1; else hi = mid; }      // contigs at least as long as the read
    const int64_t nBig = e - lo;
    int64_t v = (ldg(csum + e) - ldg(csum + lo)) - nBig * (L - 1);
    if (lo != b) {          // shorter contigs of the taxon count once each if this read maps to them (fEM.h:337-345)
      const int64_t m0 = ldg(grpOff + g), m1 = ldg(grpOff + g + 1);
      for (int64_t x = m0; x < m1; x++) {
        const int32_t sx = ldg(mSeq + x);
        if (sx < 0 || sx >= nContigs || ldg(contigTaxon + sx) != t || ldg(contigLen + sx) >= L) continue;
        bool first = true;
        for (int64_t y = m0; y < x; y++) if (ldg(mSeq + y) == sx) { first = false; break; }
        if (first) v++;
      }
    }
    tax[m] = t;
    if (nloc) nloc[m] = (double)v;
    if (weight) weight[m] = (1 / (double)v) * ldg(mapq + m);          // (1/nLoc) * mapQ, fixed over the EM rounds (fEM.h:353)
  }
};
struct EmWeightFn {     // weight from caller-supplied nloc / mapq arrays (mm_em_run)
  const double* mapq; const double* nloc; double* weight;
  MM_HD void operator()(int64_t m) const { weight[m] = (1 / ldg(nloc + m)) * ldg(mapq + m); }
};

// ---- K7 --------------------------------------------------------------------------------------------------------------
// device-side loop state: [0] ll of the previous round, [1] rounds done, [2] done flag, [3] a read with a non-positive likelihood
struct EmState { double llPrev; int32_t iters; int32_t done; int32_t bad; int32_t pad; };

// one read: likelihood sum over its mappings (fEM.h:350-363); every lane of the group returns the same s
template <int G>
MM_HD double em_read_sum(const int32_t* tax, const double* w, const double* f, int64_t b, int64_t e, int lane) {
  double s = 0;
  for (int64_t m = b + lane; m < e; m += G) s += ldg(f + ldg(tax + m)) * ldg(w + m);
  return Grp<G>::sum(s);
}

#ifndef MM_HOST_EMU
// One EM round, E-step + taxon accumulation + log-likelihood.  acc[0..T) += posterior sums, acc[T] += sum_r log s_r.
// Taxon sums go to shared memory first: `copies` private copies of the T accumulators per CTA (warp w uses copy
// w % copies; with enough copies a warp owns one and atomics never contend across warps), flushed with one global atomic
// per taxon and copy at the end.  copies == 0 (T too large for shared memory): global atomics directly.
// (Measured and dropped, config 4 shape, 20 M mappings: warp-owned copies updated with a plain load / add / store behind a
//  __match_any_sync guard + four mappings in flight per lane = 0.217 ms per round against 0.152 ms with these atomics.)
template <int G>
__global__ void __launch_bounds__(256) em_round_kernel(const int32_t* __restrict__ tax, const double* __restrict__ w, const int64_t* __restrict__ grpOff,
                                                       int64_t nGroups, const double* __restrict__ f, double* acc, int32_t T, int copies, EmState* st) {
  extern __shared__ double sacc[];
  __shared__ double llPart[8];
  if (st->done) return;
  for (int i = threadIdx.x; i < copies * T; i += blockDim.x) sacc[i] = 0.0;
  __syncthreads();
  const int lane = threadIdx.x & (G - 1);
  double* my = copies ? sacc + (size_t)((threadIdx.x >> 5) % copies) * T : acc;
  const int64_t perGrid = ((int64_t)gridDim.x * blockDim.x) / G;
  const int64_t trips = (nGroups + perGrid - 1) / perGrid;
  int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
  double ll = 0; int bad = 0;
  for (int64_t t = 0; t < trips; t++, g += perGrid) {           // same trip count on every lane: the shuffles stay converged
    const bool on = g < nGroups;
    const int64_t b = on ? grpOff[g] : 0, e = on ? grpOff[g + 1] : 0;
    const double s = em_read_sum<G>(tax, w, f, b, e, lane);
    if (e > b) {
      if (!(s > 0) || !(s <= 1.7976931348623157e308)) bad = 1;
      else {
        if (lane == 0) ll += log(s);
        for (int64_t m = b + lane; m < e; m += G) { const int32_t tx = __ldg(tax + m); atomicAdd(my + tx, (__ldg(f + tx) * __ldg(w + m)) / s); }
      }
    }
  }
  // log-likelihood: warp butterfly, then one atomic per CTA
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ll += __shfl_xor_sync(0xffffffffu, ll, o);
  if ((threadIdx.x & 31) == 0) llPart[threadIdx.x >> 5] = ll;
  if (bad) st->bad = 1;
  __syncthreads();
  if (threadIdx.x == 0) { double s = 0; for (int i = 0; i < (int)(blockDim.x >> 5); i++) s += llPart[i]; atomicAdd(acc + T, s); }
  for (int i = threadIdx.x; i < T; i += blockDim.x) {
    double s = 0;
    for (int c = 0; c < copies; c++) s += sacc[(size_t)c * T + i];
    if (copies && s != 0.0) atomicAdd(acc + i, s);
  }
}
// M-step normalisation (fEM.h:606-615) + stopping rule (fEM.h:624-640), one CTA.  acc[0..T) = f_next (all-reduced), acc[T] = ll.
__global__ void __launch_bounds__(1024) em_finish_kernel(double* acc, double* f, int32_t T, EmState* st, double* llHist, int32_t llCap, int32_t maxIter) {
  __shared__ double part[32]; __shared__ double total;
  if (st->done) return;
  double s = 0;
  for (int i = threadIdx.x; i < T; i += blockDim.x) s += acc[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) { double t = 0; for (int i = 0; i < (int)(blockDim.x >> 5); i++) t += part[i]; total = t; }
  __syncthreads();
  for (int i = threadIdx.x; i < T; i += blockDim.x) f[i] = acc[i] / total;
  if (threadIdx.x == 0) {
    const double ll = acc[T]; const int32_t it = st->iters;
    if (llHist && it < llCap) llHist[it] = ll;
    int done = 0;
    if (it > 0 && maxIter <= 0) { const double diff = ll - st->llPrev, rel = ll / st->llPrev; if (diff <= 1 && (1 - rel) < 0.0001) done = 1; }
    if (!(ll == ll)) { st->bad = 1; done = 1; }                                        // NaN can never satisfy the rule above
    st->iters = it + 1; st->llPrev = ll;
    if (maxIter > 0 && it + 1 >= maxIter) done = 1;
    if (st->bad) done = 1;
    st->done = done;
  }
}
#endif

// host-emulation / reference form of the same round (one item per read; also what the G-lane kernels are tested against)
struct EmRoundSeqFn {
  const int32_t* tax; const double* w; const int64_t* grpOff; const double* f; double* acc; int32_t T; EmState* st;
  MM_HD void operator()(int64_t g) const {
    if (st->done) return;
    const int64_t b = ldg(grpOff + g), e = ldg(grpOff + g + 1);
    if (e <= b) return;
    const double s = em_read_sum<1>(tax, w, f, b, e, 0);
    if (!(s > 0) || !(s <= 1.7976931348623157e308)) { st->bad = 1; return; }
    atomic_add(acc + T, log(s));
    for (int64_t m = b; m < e; m++) { const int32_t tx = ldg(tax + m); atomic_add(acc + tx, (ldg(f + tx) * ldg(w + m)) / s); }
  }
};
struct EmFinishSeqFn {
  double* acc; double* f; int32_t T; EmState* st; double* llHist; int32_t llCap; int32_t maxIter;
  MM_HD void operator()(int64_t) const {
    if (st->done) return;
    double total = 0;
    for (int32_t i = 0; i < T; i++) total += acc[i];
    for (int32_t i = 0; i < T; i++) f[i] = acc[i] / total;
    const double ll = acc[T]; const int32_t it = st->iters;
    if (llHist && it < llCap) llHist[it] = ll;
    int done = 0;
    if (it > 0 && maxIter <= 0) { const double diff = ll - st->llPrev, rel = ll / st->llPrev; if (diff <= 1 && (1 - rel) < 0.0001) done = 1; }
    if (!(ll == ll)) { st->bad = 1; done = 1; }
    st->iters = it + 1; st->llPrev = ll;
    if (maxIter > 0 && it + 1 >= maxIter) done = 1;
    if (st->bad) done = 1;
    st->done = done;
  }
};
struct EmFillFn { double* f; double v; MM_HD void operator()(int64_t t) const { f[t] = v; } };

// ---- K8: posterior per mapping + first maximum per read (fEM.h:693-716, :217-232) --------------------------------------
template <int G>
struct EmFinalGroupFn {
  const int32_t* tax; const double* w; const int64_t* grpOff; int64_t nGroups; const double* f; double* posterior; int64_t* best;
  MM_HD void operator()(int64_t item) const {
    const int64_t g = item / G; const int lane = Grp<G>::lane();
    const bool on = g < nGroups;
    const int64_t b = on ? ldg(grpOff + g) : 0, e = on ? ldg(grpOff + g + 1) : 0;
    const double s = em_read_sum<G>(tax, w, f, b, e, lane);
    double maxp = -1; int64_t bi = 0x7fffffffffffffffll;
    for (int64_t m = b + lane; m < e; m += G) {
      const double p = (ldg(f + ldg(tax + m)) * ldg(w + m)) / s;
      posterior[m] = p;
      if (p > maxp || !(maxp >= 0)) { maxp = p; bi = m; }          // strict >: the first maximum of this lane's subsequence
    }
    Grp<G>::argmax_first(maxp, bi);
    if (on && lane == 0) best[g] = e > b ? bi : b;
  }
};


// ---- host orchestration ------------------------------------------------------------------------------------------------
struct Taxonomy {       // mm_classify_setup: contig lengths / taxa (global contig ids) + each taxon's contigs sorted by length
  DevBuf<int64_t> contigLen, lens, start, csum; DevBuf<int32_t> contigTaxon; int32_t nContigs = 0, T = 0; bool set = false;
  void upload(Runtime& rt, const int64_t* contig_len, const int32_t* contig_taxon, int32_t n_contigs, int32_t T_) {
    std::vector<int64_t> st((size_t)T_ + 1, 0);
    for (int32_t c = 0; c < n_contigs; c++) {
      if (contig_taxon[c] < 0 || contig_taxon[c] >= T_) throw Error(-22, "contig taxon out of range");
      st[(size_t)contig_taxon[c] + 1]++;
    }
    for (int32_t t = 0; t < T_; t++) st[(size_t)t + 1] += st[(size_t)t];
    std::vector<int64_t> ln((size_t)n_contigs), fill(st.begin(), st.end() - 1), cs((size_t)n_contigs + 1, 0);
    for (int32_t c = 0; c < n_contigs; c++) ln[(size_t)fill[(size_t)contig_taxon[c]]++] = contig_len[c];
    for (int32_t t = 0; t < T_; t++) std::sort(ln.begin() + st[(size_t)t], ln.begin() + st[(size_t)t + 1]);
    for (int32_t c = 0; c < n_contigs; c++) cs[(size_t)c + 1] = cs[(size_t)c] + ln[(size_t)c];
    contigLen.ensure((size_t)n_contigs + 1); contigTaxon.ensure((size_t)n_contigs + 1); lens.ensure((size_t)n_contigs + 1);
    start.ensure((size_t)T_ + 1); csum.ensure((size_t)n_contigs + 1);
    h2d(rt, contigLen.p, contig_len, 8 * (size_t)n_contigs); h2d(rt, contigTaxon.p, contig_taxon, 4 * (size_t)n_contigs);
    h2d(rt, lens.p, ln.data(), 8 * (size_t)n_contigs); h2d(rt, start.p, st.data(), 8 * ((size_t)T_ + 1)); h2d(rt, csum.p, cs.data(), 8 * ((size_t)n_contigs + 1));
    rt.sync();
    nContigs = n_contigs; T = T_; set = true;
  }
};

struct Classifier {
  Runtime& rt; Prims& pr;
  MapTable tab; Taxonomy taxo;
  DevBuf<float> id32; DevBuf<double> parsed, mapq, nloc, w, f, acc, post, llHist;
  DevBuf<int32_t> head, mGrp, grpRead, grpLen, status, tax, fixList, bad, tmpI; DevBuf<uint32_t> keyA, keyB, permA, permB;
  DevBuf<int64_t> gidx, grpOff, best; DevBuf<unsigned long long> cnt; DevBuf<EmState> st;
  int64_t nGroups = 0; int32_t iters = 0; int64_t nFix = 0; double emMs = 0; int32_t lastT = 0;
  // several read batches in one table (mm_classify_next_batch): read indices are readBase + index in the batch
  DevBuf<int32_t> readLenAll; int64_t readBase = 0, readsSeen = 0;
  std::function<void(double*, size_t)> allreduce;       // sum over ranks of a device buffer (NCCL or the host transport); empty = single rank
  bool hostTransport = false;

  Classifier(Runtime& r, Prims& p) : rt(r), pr(p) {}

  static int group_size(int64_t M, int64_t groups) {
#ifdef MM_HOST_EMU
    (void)M; (void)groups; return 1;
#else
    const double avg = groups > 0 ? (double)M / (double)groups : 0;
    return avg > 12 ? 32 : avg > 3 ? 8 : 4;
#endif
  }

  // groups = runs of equal read index in mRead[0..n) (non-decreasing): grpOff, grpRead, mGrp; one host sync for the count
  void build_groups(const int32_t* mRead, int64_t n) {
    head.ensure((size_t)n + 2); gidx.ensure((size_t)n + 2); mGrp.ensure((size_t)n + 1);
    foreach(rt, n + 1, GroupHeadFn{mRead, head.p, n});
    pr.exclusive_sum<int32_t, int64_t>(head.p, gidx.p, n + 1);
    d2h(rt, &nGroups, gidx.p + n, sizeof(int64_t));
    grpOff.ensure((size_t)nGroups + 2); grpRead.ensure((size_t)nGroups + 1); grpLen.ensure((size_t)nGroups + 1);
    foreach(rt, n + 1, GroupScatterFn{mRead, head.p, gidx.p, n, grpOff.p, grpRead.p, mGrp.p});
  }

  // nucIdentity + its 6-significant-digit round trip for n (shared, sketch) pairs on the device; the (rare) unsure ones through glibc
  void identity(const int32_t* shared, const int32_t* sketch, int64_t n, int k, const int32_t* mask = nullptr) {
    id32.ensure((size_t)n + 1); parsed.ensure((size_t)n + 1); cnt.ensure(2); fixList.ensure(4096);
    dev_memset(rt, cnt.p, 0, sizeof(unsigned long long));
    foreach(rt, n, IdentityFn{shared, sketch, k, id32.p, parsed.p, cnt.p, fixList.p, (int64_t)fixList.cap, mask});
  }
  // after a host sync point: settle the flagged identities (count read by the caller together with its other scalars)
  void identity_fixups(const int32_t* shared, const int32_t* sketch, int64_t n, int k, unsigned long long flagged, const int32_t* mask = nullptr);

  template <int G>
  void mapq_t(const double* idArr, double div, const int32_t* shared, const int32_t* sketch, int k) {
    foreach(rt, round_up32(nGroups * G), MapqGroupFn<G>{idArr, div, shared, sketch, grpLen.p, grpOff.p, nGroups, k, mapq.p, status.p});
  }
  void run_mapq(const double* idArr, double div, const int32_t* shared, const int32_t* sketch, int64_t M, int k) {
    mapq.ensure((size_t)M + 1); status.ensure((size_t)nGroups + 1);
    const int G = group_size(M, nGroups);
#ifdef MM_HOST_EMU
    (void)G; mapq_t<1>(idArr, div, shared, sketch, k);
#else
    if (G == 32) mapq_t<32>(idArr, div, shared, sketch, k); else if (G == 8) mapq_t<8>(idArr, div, shared, sketch, k); else mapq_t<4>(idArr, div, shared, sketch, k);
#endif
  }

  // EM over M mappings in nG groups (grpOffDev: nG + 1 offsets), weights in w.p, taxa in taxDev.  Leaves f, post, best on the device.
  void run_em(const int32_t* taxDev, const int64_t* grpOffDev, int64_t nG, int64_t M, int32_t T, int32_t maxIter, int32_t llCap) {
    f.ensure((size_t)T); acc.ensure((size_t)T + 1); post.ensure((size_t)M + 1); best.ensure((size_t)nG + 1); st.ensure(1);
    llHist.ensure((size_t)(llCap > 0 ? llCap : 1));
    foreach(rt, T, EmFillFn{f.p, 1.0 / (double)T});                       // fEM.h:491-495
    dev_memset(rt, st.p, 0, sizeof(EmState));
    lastT = T;
    const int G = group_size(M, nG);
    EmState hs; memset(&hs, 0, sizeof hs);
    const int32_t HARD_CAP = 1000000;                                    // the reference has no cap (fEM.h:636); a diverging input must not hang the call
    int check = hostTransport ? 1 : 4;
    if (const char* e = getenv("MM_EM_CHECK")) { int v = atoi(e); if (v >= 1) check = v; }
#ifndef MM_HOST_EMU
    int copies = 0; size_t smem = 0; int grid = 1;
    {
      const size_t per = (size_t)T * sizeof(double);
      if (per <= 48 * 1024) { copies = (int)((48 * 1024) / per); if (copies > 8) copies = 8; if (copies < 1) copies = 1; }
      else if (per <= 200 * 1024) copies = 1;
      smem = (size_t)copies * per;
      int perSm = smem ? (int)((220 * 1024) / (smem + 1024)) : 8; if (perSm > 8) perSm = 8; if (perSm < 1) perSm = 1;
      const int64_t need = (nG * G + 255) / 256;
      grid = (int)std::min<int64_t>(std::max<int64_t>(need, 1), (int64_t)rt.sm_count * perSm);
      if (smem > 48 * 1024) {
        if (G == 32 && rt.first((const void*)em_round_kernel<32>)) MM_CUDA(cudaFuncSetAttribute(em_round_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        if (G == 8 && rt.first((const void*)em_round_kernel<8>)) MM_CUDA(cudaFuncSetAttribute(em_round_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        if (G == 4 && rt.first((const void*)em_round_kernel<4>)) MM_CUDA(cudaFuncSetAttribute(em_round_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      }
    }
#endif
    {
      StageTimer tm(rt, &emMs);
      while (true) {
        int burst = check;
        if (maxIter > 0) burst = std::min(check, maxIter - hs.iters);
        for (int i = 0; i < burst; i++) {
          dev_memset(rt, acc.p, 0, sizeof(double) * ((size_t)T + 1));
#ifdef MM_HOST_EMU
          foreach(rt, nG, EmRoundSeqFn{taxDev, w.p, grpOffDev, f.p, acc.p, T, st.p});
#else
          if (G == 32) em_round_kernel<32><<<grid, 256, smem, rt.stream>>>(taxDev, w.p, grpOffDev, nG, f.p, acc.p, T, copies, st.p);
          else if (G == 8) em_round_kernel<8><<<grid, 256, smem, rt.stream>>>(taxDev, w.p, grpOffDev, nG, f.p, acc.p, T, copies, st.p);
          else em_round_kernel<4><<<grid, 256, smem, rt.stream>>>(taxDev, w.p, grpOffDev, nG, f.p, acc.p, T, copies, st.p);
          MM_CUDA(cudaGetLastError());
          rt.launches++;
#endif
          if (allreduce) allreduce(acc.p, (size_t)T + 1);                  // taxon sums + log-likelihood (slot T) over the ranks
#ifdef MM_HOST_EMU
          foreach(rt, 1, EmFinishSeqFn{acc.p, f.p, T, st.p, llHist.p, llCap, maxIter});
#else
          em_finish_kernel<<<1, 1024, 0, rt.stream>>>(acc.p, f.p, T, st.p, llHist.p, llCap, maxIter);
          MM_CUDA(cudaGetLastError());
          rt.launches++;
#endif
        }
        d2h(rt, &hs, st.p, sizeof(EmState));                               // one small read per burst of rounds
        if (hs.done || hs.iters >= HARD_CAP) break;
      }
      iters = hs.iters;
      if (!hs.bad) {
#ifdef MM_HOST_EMU
        foreach(rt, nG, EmFinalGroupFn<1>{taxDev, w.p, grpOffDev, nG, f.p, post.p, best.p});
#else
        if (G == 32) foreach(rt, round_up32(nG * 32), EmFinalGroupFn<32>{taxDev, w.p, grpOffDev, nG, f.p, post.p, best.p});
        else if (G == 8) foreach(rt, round_up32(nG * 8), EmFinalGroupFn<8>{taxDev, w.p, grpOffDev, nG, f.p, post.p, best.p});
        else foreach(rt, round_up32(nG * 4), EmFinalGroupFn<4>{taxDev, w.p, grpOffDev, nG, f.p, post.p, best.p});
#endif
      }
    }
    if (hs.bad) throw Error(-22, "EM: a read's likelihood sum is not positive (all its mapping qualities are 0, or an nLoc is 0); the reference asserts here (fEM.h:357)");
    if (!hs.done) throw Error(-34, "EM did not reach the reference's stopping rule within 1000000 rounds");
  }
};
inline void Classifier::identity_fixups(const int32_t* shared, const int32_t* sketch, int64_t n, int k, unsigned long long flagged, const int32_t* mask) {
  nFix = (int64_t)flagged;
  if (!flagged) return;
  std::vector<int32_t> list;
  if ((int64_t)flagged > (int64_t)fixList.cap) {      // more than the list holds (identities outside [10,100), e.g. a very low --pi): redo all on the host
    std::vector<int32_t> hm;
    if (mask) { hm.resize((size_t)n); d2h(rt, hm.data(), mask, 4 * (size_t)n); }
    list.reserve((size_t)n);
    for (int64_t i = 0; i < n; i++) if (!mask || hm[(size_t)i]) list.push_back((int32_t)i);
    std::vector<int32_t> hs((size_t)n), hk((size_t)n);
    d2h(rt, hs.data(), shared, 4 * (size_t)n); d2h(rt, hk.data(), sketch, 4 * (size_t)n);
    std::vector<float> hid((size_t)n, 0.f); std::vector<double> hp((size_t)n, 0.0);
    for (int32_t m : list) {
      float id; stats::identity_only(hs[(size_t)m], hk[(size_t)m], k, &id);
      const double x = (double)id; double pz;
      if (x >= 10.0 && x < 99.99995) pz = rint(x * 1e4) / 1e4;
      else { char buf[64]; snprintf(buf, sizeof buf, "%.6g", x); pz = strtod(buf, nullptr); }
      hid[(size_t)m] = id; hp[(size_t)m] = pz;
    }
    h2d(rt, id32.p, hid.data(), 4 * (size_t)n); h2d(rt, parsed.p, hp.data(), 8 * (size_t)n); rt.sync();
    return;
  }
  list.resize((size_t)flagged); d2h(rt, list.data(), fixList.p, 4 * (size_t)flagged);
  for (size_t i = 0; i < list.size(); i++) {
    const int64_t m = list[i];
    int32_t a, b;
    d2h(rt, &a, shared + m, 4); d2h(rt, &b, sketch + m, 4);
    float id; stats::identity_only(a, b, k, &id);
    const double x = (double)id; double pz;
    if (x >= 10.0 && x < 99.99995) pz = rint(x * 1e4) / 1e4;
    else { char buf[64]; snprintf(buf, sizeof buf, "%.6g", x); pz = strtod(buf, nullptr); }
    h2d(rt, id32.p + m, &id, 4); h2d(rt, parsed.p + m, &pz, 8);
    rt.sync();
  }
}

}  // namespace mm
