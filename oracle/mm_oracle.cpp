// oracle/mm_oracle.cpp -- TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the MetaMaps hot path (map + EM), used as the parity checker by tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline leg.  The product (metamaps_b200/csrc)
// never includes, links or calls anything in this directory.
//
// Each function cites the reference file:line (relative to /root/reference/src) it restates.
// Pinned against (a) the unmodified reference compiled with oracle/boost_shim (oracle/_ref:
// libmm_refharness.so function-level, metamaps binary end-to-end) and (b) the golden vectors in
// tests/golden (example-output KAT, SciPy/Boost binomial vectors).
//
// Plain scalar C++ on purpose: clarity over speed.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <deque>
#include <limits>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "mm_refmath.h"

namespace {

// ---------------------------------------------------------------- murmur3 (common/murmur3.h:226-303)
inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
inline uint64_t fmix64(uint64_t k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
  return k;
}
// MurmurHash3_x64_128, returns the first 32 bits of the digest (commonFunc.hpp:71-81)
uint32_t murmur_low32(const uint8_t* data, int len, uint32_t seed) {
  const int nblocks = len / 16;
  uint64_t h1 = seed, h2 = seed;
  const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
  for (int i = 0; i < nblocks; i++) {
    uint64_t k1, k2;
    memcpy(&k1, data + i * 16, 8);
    memcpy(&k2, data + i * 16 + 8, 8);
    k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
    h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
    k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2;
    h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
  }
  const uint8_t* tail = data + nblocks * 16;
  uint64_t k1 = 0, k2 = 0;
  int rem = len & 15;
  for (int i = rem - 1; i >= 8; i--) k2 ^= (uint64_t)tail[i] << ((i - 8) * 8);
  if (rem > 8) { k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2; }
  for (int i = std::min(rem, 8) - 1; i >= 0; i--) k1 ^= (uint64_t)tail[i] << (i * 8);
  if (rem > 0) { k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1; }
  h1 ^= (uint64_t)len; h2 ^= (uint64_t)len;
  h1 += h2; h2 += h1;
  h1 = fmix64(h1); h2 = fmix64(h2);
  h1 += h2;
  return (uint32_t)h1;
}

struct Mini { uint32_t hash; int32_t seqId; int32_t wpos; int32_t strand; };
inline bool same(const Mini& a, const Mini& b) {
  return a.hash == b.hash && a.seqId == b.seqId && a.wpos == b.wpos && a.strand == b.strand;
}

// commonFunc.hpp:92-175  (upper-casing :57-66, reverse complement :38-55, seed 42 :33)
void add_minimizers(std::vector<Mini>& out, const char* seq_in, int len, int k, int w, int seqId) {
  std::string seq(seq_in, seq_in + len);
  for (auto& c : seq) if (c > 96 && c < 123) c -= 32;
  std::string rev(len, 'N');
  for (int i = 0; i < len; i++) {
    char b = seq[i];
    switch (b) { case 'A': b = 'T'; break; case 'C': b = 'G'; break; case 'G': b = 'C'; break; case 'T': b = 'A'; break; default: break; }
    rev[len - i - 1] = b;
  }
  std::deque<std::pair<Mini, int>> Q;
  for (int i = 0; i < len - k + 1; i++) {
    int win = i - w + 1;
    uint32_t hf = murmur_low32((const uint8_t*)seq.data() + i, k, 42);
    uint32_t hb = murmur_low32((const uint8_t*)rev.data() + len - i - k, k, 42);
    if (hb == hf) continue;                                  // :130 symmetric k-mer: nothing happens
    uint32_t cur = std::min(hf, hb);
    int strand = hf < hb ? 1 : -1;
    while (!Q.empty() && Q.front().second <= i - w) Q.pop_front();          // :139
    while (!Q.empty() && Q.back().first.hash >= cur) Q.pop_back();          // :144 (ties: newest wins)
    Q.push_back({Mini{cur, seqId, 0, strand}, i});
    if (win >= 0) {
      if (out.empty() || !same(out.back(), Q.front().first)) {              // :157
        Q.front().first.wpos = win;
        out.push_back(Q.front().first);
      }
    }
  }
}

// ---------------------------------------------------------------- statistics (map_stats.hpp)
inline float j2md(float j, int k) {                       // :44-54
  if (j == 0) return 1.0;
  if (j == 1) return 0.0;
  float d = (-1.0 / k) * log(2.0 * j / (1 + j));
  return d;
}
inline float md2j(float d, int k) { float j = 1.0 / (2.0 * exp(k * d) - 1.0); return j; }  // :62-66
inline float md_lower_bound(float d, int s, int k, float ci) {       // :79-111 (USE_BOOST branch)
  float q2 = (1.0 - ci) / 2;
  int x = (int)mmref::binom_quantile_upper(s, (double)md2j(d, k), (double)q2);
  float jaccard = float(x) / s;
  return j2md(jaccard, k);
}
inline int estimateMinimumHits(int s, int k, float pi) {             // :120-131
  float md = 1.0 - pi / 100.0;
  float j = md2j(md, k);
  return (int)ceil(1.0 * s * j);
}
int estimateMinimumHitsRelaxed(int s, int k, float pi) {             // :142-167
  int first = estimateMinimumHits(s, k, pi);
  int relaxed = first;
  for (int i = first; i >= 0; i--) {
    float jaccard = 1.0 * i / s;
    float d = j2md(jaccard, k);
    float dl = md_lower_bound(d, s, k, 0.9);
    float idu = 100.0 * (1.0 - dl);
    if (idu >= pi) relaxed = i; else break;
  }
  return relaxed;
}
double estimate_pvalue(int s, int k, int alphabet, float identity, int lenQ, uint64_t lenR) {  // :179-213
  double kmerSpace = pow(alphabet, k);
  double pX, pY; pX = pY = 1. / (1. + kmerSpace / lenQ);
  double r = pX * pY / (pX + pY - pX * pY);
  int x = estimateMinimumHitsRelaxed(s, k, identity);
  double cc = (x == 0) ? 1.0 : mmref::binom_sf(x - 1, s, r);
  return lenR * cc;
}
int recommendedWindowSize(double pcut, int k, int alphabet, float identity, int lenQ, uint64_t lenR) {  // :226-256
  std::vector<int> cand{1, 2, 5};
  for (int i = 10; i < lenQ; i += 10) cand.push_back(i);
  int opt = 0;
  for (int e : cand) { if (estimate_pvalue(e, k, alphabet, identity, lenQ, lenR) <= pcut) { opt = e; break; } }
  int w = 2.0 * lenQ / opt;
  return std::min(std::max(w, 1), lenQ);
}

// ---------------------------------------------------------------- index (winSketch.hpp)
struct Index {
  int k, w;
  std::vector<Mini> mi;                                       // minimizerIndex, (seqId,wpos) order :129
  std::unordered_map<uint32_t, std::vector<Mini>> lookup;     // minimizerPosLookupIndex :119
  std::vector<int> contigLen;
  int freqThreshold = std::numeric_limits<int>::max();
  std::map<int, int> hist;

  void add(const char* seq, int len) {                        // build_and_store_index :252-345 (no memory limit)
    int id = (int)contigLen.size();
    contigLen.push_back(len);
    if (len < w || len < k) return;
    std::vector<Mini> v;
    add_minimizers(v, seq, len, k, w, id);
    for (auto& e : v) lookup[e.hash].push_back(e);
    mi.insert(mi.end(), v.begin(), v.end());
  }
  void freq_hist() {                                          // computeFreqHist :452-495
    if (lookup.empty()) return;
    for (auto& e : lookup) hist[(int)e.second.size()] += 1;
    int64_t total = (int64_t)lookup.size();
    int64_t toIgnore = total * 0.001f / 100;                  // float percentageThreshold = 0.001 :85
    int64_t sum = 0;
    for (auto it = hist.rbegin(); it != hist.rend(); ++it) {
      sum += it->second;
      if (sum < toIgnore) freqThreshold = it->first;
      else if (sum == toIgnore) { freqThreshold = it->first; break; }
      else break;
    }
  }
  // searchIndex :506-517  lower_bound on (seqId,wpos)
  int64_t search(int seqId, int wpos) const {
    int64_t lo = 0, hi = (int64_t)mi.size();
    while (lo < hi) {
      int64_t m = (lo + hi) / 2;
      if (std::make_pair(mi[m].seqId, mi[m].wpos) < std::make_pair(seqId, wpos)) lo = m + 1; else hi = m;
    }
    return lo;
  }
};

// ---------------------------------------------------------------- SlideMapper (slidingMap.hpp)
struct SlideMapper {
  struct V { int wposQ, strandQ, wposR, strandR; };
  static const int NA = std::numeric_limits<int>::max();
  std::map<uint32_t, V> M;
  std::map<uint32_t, V>::iterator pivot;
  int shared = 0, s;
  SlideMapper(const std::vector<Mini>& q, int s_) : s(s_) {         // init :114-131
    for (int i = 0; i < s; i++) M.emplace_hint(M.end(), q[i].hash, V{q[i].wpos, q[i].strand, NA, 0});
    pivot = std::next(M.begin(), s - 1);
  }
  void insert_ref(const Mini& m) {                                   // :139-164, :263-287
    int status;
    auto f = M.find(m.hash);
    if (f == M.end()) { M[m.hash] = V{NA, 0, m.wpos, m.strand}; status = 1; }
    else { status = (f->second.wposR == NA) ? 2 : 3; f->second.wposR = m.wpos; f->second.strandR = m.strand; }
    if (m.hash <= pivot->first) {
      if (status == 2) shared += 1;
      else if (status == 1) {
        if (pivot->second.wposQ != NA && pivot->second.wposR != NA) shared -= 1;
        --pivot;
      }
    }
  }
  void delete_ref(const Mini& m) {                                   // :170-219, :294-316
    int status; bool pivotDel = false;
    auto f = M.find(m.hash);
    if (f->second.wposR == m.wpos) {
      if (f->second.wposQ == NA) {
        if (f == pivot) {
          ++pivot;
          if (pivot->second.wposQ != NA && pivot->second.wposR != NA) shared += 1;
          pivotDel = true;
        }
        M.erase(f);
        status = 1;
      } else { f->second.wposR = NA; status = 2; }
    } else status = 3;
    if (!pivotDel && m.hash <= pivot->first) {
      if (status == 2) shared -= 1;
      else if (status == 1) {
        ++pivot;
        if (pivot->second.wposQ != NA && pivot->second.wposR != NA) shared += 1;
      }
    }
  }
  void stats(int& votes, int& uniqRef) {                             // computeStatistics :232-254
    int n = 0; votes = uniqRef = 0;
    for (auto& e : M) {
      n++;
      if (n <= s && e.second.wposQ != NA && e.second.wposR != NA) votes += e.second.strandQ * e.second.strandR;
      if (e.second.wposR != NA) uniqRef++;
    }
  }
};

struct L1Cand { int seqId, start, end; };
struct L2Res { int seqId, meanOptimalPos, shared; int64_t optStart, optEnd; int strandVotes; int valid; };

struct ReadMap {
  std::vector<Mini> q; int s = 0; int minimumHits = 0; int64_t nHits = 0;
  std::vector<L1Cand> l1; std::vector<L2Res> l2;
};

// computeMap.hpp:277-336 (doL1Mapping) + :346-386 (computeL1CandidateRegions)
void do_l1(const Index& ix, const char* seq, int len, float pi, ReadMap& R) {
  add_minimizers(R.q, seq, len, ix.k, ix.w, 0);
  std::sort(R.q.begin(), R.q.end(), [](const Mini& a, const Mini& b) { return a.hash < b.hash; });   // :292 (same libstdc++ introsort)
  auto ue = std::unique(R.q.begin(), R.q.end(), [](const Mini& a, const Mini& b) { return a.hash == b.hash; });
  R.s = (int)std::distance(R.q.begin(), ue);
  if (R.s == 0) return;
  struct Hit { int seqId, wpos, strand; };
  std::vector<Hit> hits;
  for (int i = 0; i < R.s; i++) {
    auto f = ix.lookup.find(R.q[i].hash);
    if (f != ix.lookup.end() && (int64_t)f->second.size() < (int64_t)ix.freqThreshold)
      for (auto& e : f->second) hits.push_back(Hit{e.seqId, e.wpos, e.strand});
  }
  R.nHits = (int64_t)hits.size();
  int minimumHits = estimateMinimumHitsRelaxed(R.s, ix.k, pi);
  R.minimumHits = minimumHits;
  if (minimumHits < 1) minimumHits = 1;
  std::sort(hits.begin(), hits.end(), [](const Hit& a, const Hit& b) {
    return std::tie(a.seqId, a.wpos, a.strand) < std::tie(b.seqId, b.wpos, b.strand); });
  for (size_t i = 0; i < hits.size(); i++) {
    if (hits.size() - i >= (size_t)minimumHits) {
      const Hit& a = hits[i]; const Hit& b = hits[i + minimumHits - 1];
      if (b.seqId == a.seqId && b.wpos - a.wpos < len) {
        L1Cand c{a.seqId, std::max(0, b.wpos - len + 1), a.wpos};
        if (!R.l1.empty() && c.seqId == R.l1.back().seqId && R.l1.back().end >= c.start)
          R.l1.back().end = std::max(c.end, R.l1.back().end);
        else R.l1.push_back(c);
      }
    }
  }
}

// computeMap.hpp:460-538 + MIIteratorL2.hpp:54-96
void do_l2(const Index& ix, int len, ReadMap& R) {
  for (auto& c : R.l1) {
    L2Res o{}; o.seqId = c.seqId;
    int64_t fs = ix.search(c.seqId, c.start);
    int cmw = len - (ix.w - 1) - (ix.k - 1);
    int64_t fe = ix.search(c.seqId, ix.mi[fs].wpos + cmw);
    int64_t le = ix.search(c.seqId, c.end + len);
    SlideMapper sm(R.q, R.s);
    int64_t beg = fs, end = fe; int sw_pos = ix.mi[beg].wpos;
    for (int64_t j = beg; j < end; j++) sm.insert_ref(ix.mi[j]);
    int64_t pb = beg, pe = end;
    int bpos = 0, lpos = 0; bool any = false;
    while (end < le) {
      if (pb != beg) sm.delete_ref(ix.mi[pb]);
      if (pe != end) sm.insert_ref(ix.mi[pe]);
      if (sm.shared > o.shared) { o.shared = sm.shared; o.optStart = beg; o.optEnd = end; bpos = lpos = ix.mi[beg].wpos; any = true; }
      else if (sm.shared == o.shared) { lpos = ix.mi[beg].wpos; }
      pb = beg; pe = end;
      int beginPos = sw_pos, lastPos = sw_pos + cmw - 1;
      int adv = std::min(ix.mi[beg + 1].wpos - beginPos, ix.mi[end].wpos - lastPos);
      sw_pos += adv;
      if (adv == ix.mi[pb + 1].wpos - beginPos) beg++;
      if (adv == ix.mi[pe].wpos - lastPos) end++;
    }
    // The reference leaves beginOptimalPos uninitialised when no window beats 0 shared; such a
    // candidate has shared == 0 and never passes the identity filter, so it is flagged invalid here.
    o.valid = any ? 1 : 0;
    o.meanOptimalPos = (bpos + lpos) / 2;
    if (any) {                                              // strand: computeMap.hpp:431-438
      SlideMapper s2(R.q, R.s);
      for (int64_t j = o.optStart; j < o.optEnd; j++) s2.insert_ref(ix.mi[j]);
      int votes, uniq; s2.stats(votes, uniq);
      o.strandVotes = votes;
    }
    R.l2.push_back(o);
  }
}

} // namespace

// ==================================================================== C interface (ctypes)
extern "C" {

uint32_t mmo_hash(const char* kmer, int k) { return murmur_low32((const uint8_t*)kmer, k, 42); }

int64_t mmo_minimizers(const char* seq, int len, int k, int w, uint32_t* hash, int32_t* wpos, int32_t* strand, int64_t cap) {
  std::vector<Mini> v;
  if (!(len < w || len < k)) add_minimizers(v, seq, len, k, w, 0);
  for (int64_t i = 0; i < (int64_t)v.size() && i < cap; i++) { hash[i] = v[i].hash; wpos[i] = v[i].wpos; strand[i] = v[i].strand; }
  return (int64_t)v.size();
}

int mmo_min_hits_relaxed(int s, int k, float pi) { return estimateMinimumHitsRelaxed(s, k, pi); }
int mmo_recommended_window(double p, int k, int alphabet, float pi, int lenQ, uint64_t lenR) { return recommendedWindowSize(p, k, alphabet, pi, lenQ, lenR); }
double mmo_estimate_pvalue(int s, int k, int alphabet, float pi, int lenQ, uint64_t lenR) { return estimate_pvalue(s, k, alphabet, pi, lenQ, lenR); }
// computeMap.hpp:405-411: identity and its 90 % CI upper bound from (shared, s)
void mmo_identity(int shared, int s, int k, float* nucIdentity, float* upper) {
  float md = j2md(1.0 * shared / s, k);
  float lb = md_lower_bound(md, s, k, 0.9);
  *nucIdentity = 100 * (1 - md);
  *upper = 100 * (1 - lb);
}
double mmo_binom_pmf(int k, int n, double p) { return mmref::binom_pmf(k, n, p); }
double mmo_binom_quantile_upper(int n, double p, double q) { return mmref::binom_quantile_upper(n, p, q); }
double mmo_binom_sf(int k, int n, double p) { return mmref::binom_sf(k, n, p); }

void* mmo_index_build(const char* seqs, const int64_t* offsets, int n, int k, int w) {
  Index* ix = new Index(); ix->k = k; ix->w = w;
  for (int i = 0; i < n; i++) ix->add(seqs + offsets[i], (int)(offsets[i + 1] - offsets[i]));
  ix->freq_hist();
  return ix;
}
void mmo_index_free(void* p) { delete (Index*)p; }
int64_t mmo_index_size(void* p) { return (int64_t)((Index*)p)->mi.size(); }
int64_t mmo_index_unique(void* p) { return (int64_t)((Index*)p)->lookup.size(); }
int mmo_index_freq_threshold(void* p) { return ((Index*)p)->freqThreshold; }
void mmo_index_get(void* p, uint32_t* hash, int32_t* seqId, int32_t* wpos, int32_t* strand) {
  Index* ix = (Index*)p;
  for (size_t i = 0; i < ix->mi.size(); i++) { hash[i] = ix->mi[i].hash; seqId[i] = ix->mi[i].seqId; wpos[i] = ix->mi[i].wpos; strand[i] = ix->mi[i].strand; }
}

// Map one read.  Outputs (caller-sized by cap): L1 candidates and their L2 results.
// info[0]=sketch size s, info[1]=minimumHits (before the max(1,.) clamp), info[2]=#L1 candidates, info[3]=#seed hits
int mmo_map_read(void* p, const char* seq, int len, float pi, int64_t* info,
                 int32_t* c_seq, int32_t* c_start, int32_t* c_end,
                 int32_t* l2_pos, int32_t* l2_shared, int32_t* l2_votes, int32_t* l2_valid,
                 int64_t* l2_optStart, int64_t* l2_optEnd, int cap) {
  Index* ix = (Index*)p;
  ReadMap R;
  do_l1(*ix, seq, len, pi, R);
  if (R.s > 0) do_l2(*ix, len, R);
  info[0] = R.s; info[1] = R.minimumHits; info[2] = (int64_t)R.l1.size(); info[3] = R.nHits;
  for (int i = 0; i < (int)R.l1.size() && i < cap; i++) {
    c_seq[i] = R.l1[i].seqId; c_start[i] = R.l1[i].start; c_end[i] = R.l1[i].end;
    l2_pos[i] = R.l2[i].meanOptimalPos; l2_shared[i] = R.l2[i].shared; l2_votes[i] = R.l2[i].strandVotes;
    l2_valid[i] = R.l2[i].valid; l2_optStart[i] = R.l2[i].optStart; l2_optEnd[i] = R.l2[i].optEnd;
  }
  return (int)R.l1.size();
}
// query sketch after sort+unique (computeMap.hpp:292-298): the s surviving minimizers
int mmo_read_sketch(const char* seq, int len, int k, int w, uint32_t* hash, int32_t* wpos, int32_t* strand, int cap) {
  std::vector<Mini> q;
  add_minimizers(q, seq, len, k, w, 0);
  std::sort(q.begin(), q.end(), [](const Mini& a, const Mini& b) { return a.hash < b.hash; });
  auto ue = std::unique(q.begin(), q.end(), [](const Mini& a, const Mini& b) { return a.hash == b.hash; });
  int s = (int)std::distance(q.begin(), ue);
  for (int i = 0; i < s && i < cap; i++) { hash[i] = q[i].hash; wpos[i] = q[i].wpos; strand[i] = q[i].strand; }
  return s;
}

// mapWrap.h:215-323 + :332-356.  identity[] = column 10 / 100 as re-parsed from text.
// Returns 0, or -1 if the likelihood sum is zero (reference: assert(likelihood_sum > 0) :298).
int mmo_mapq(const double* identity, const int32_t* shared, const int32_t* sketch, int n, int readLen, int k, double* mapq) {
  double maxid = -1;
  for (int i = 0; i < n; i++) if (identity[i] > maxid) maxid = identity[i];
  maxid = exp(-(1 - maxid));
  int n_kmers = readLen - k + 1;
  double sum = 0;
  for (int i = 0; i < n; i++) {
    double surv = std::pow(maxid, k);
    double E = std::round(surv * n_kmers);
    double U = n_kmers + (n_kmers - E);
    mapq[i] = mmref::binom_pmf(shared[i], sketch[i], E / U);
    sum += mapq[i];
  }
  if (!(sum > 0)) return -1;
  for (int i = 0; i < n; i++) mapq[i] /= sum;
  return 0;
}

// fEM.h:491-661 (EM loop) + :234-373 (per-read posterior) + :693-716 (final pass) on arrays:
//   taxon[m]  index into f, mapq[m] = column 14, nloc[m] = mappingLocations_per_taxonID for that read/taxon,
//   read_off[r..r+1] delimits the mappings of read r (reads with >= 1 mapping only).
// max_iter <= 0: run to the reference's stopping rule (:636); max_iter > 0: exactly max_iter rounds (the
// reference has no cap; BASELINE config 4 benchmarks a fixed round count).  Returns the rounds run.
int mmo_em(const int32_t* taxon, const double* mapq, const double* nloc, const int64_t* read_off, int64_t n_reads, int T,
           int max_iter, double* f, double* posterior, int64_t* best, double* ll_hist, int ll_cap) {
  std::vector<double> fcur(T, 1.0 / (double)T), fn(T);
  double ll_last = 0; int it = 0; bool cont = true;
  while (cont) {
    std::fill(fn.begin(), fn.end(), 0.0);
    double ll = 0;
    for (int64_t r = 0; r < n_reads; r++) {
      double tot = 0;
      for (int64_t m = read_off[r]; m < read_off[r + 1]; m++) tot += fcur[taxon[m]] * (1 / (double)nloc[m]) * mapq[m];
      for (int64_t m = read_off[r]; m < read_off[r + 1]; m++) fn[taxon[m]] += (fcur[taxon[m]] * (1 / (double)nloc[m]) * mapq[m]) / tot;
      ll += log(tot);
    }
    double sum = 0; for (double v : fn) sum += v;
    for (double& v : fn) v /= sum;
    if (it < ll_cap) ll_hist[it] = ll;
    if (it > 0 && max_iter <= 0) {
      double diff = ll - ll_last, rel = ll / ll_last;
      if (diff <= 1 && (1 - rel) < 0.0001) cont = false;
    }
    fcur = fn; it++; ll_last = ll;
    if (max_iter > 0 && it >= max_iter) cont = false;
  }
  for (int64_t r = 0; r < n_reads; r++) {
    double tot = 0;
    for (int64_t m = read_off[r]; m < read_off[r + 1]; m++) tot += fcur[taxon[m]] * (1 / (double)nloc[m]) * mapq[m];
    double maxp = 0; int64_t bi = read_off[r];
    for (int64_t m = read_off[r]; m < read_off[r + 1]; m++) {
      posterior[m] = (fcur[taxon[m]] * (1 / (double)nloc[m]) * mapq[m]) / tot;
      if (m == read_off[r] || posterior[m] > maxp) { maxp = posterior[m]; bi = m; }   // getBestMapping :217-232
    }
    best[r] = bi;
  }
  for (int t = 0; t < T; t++) f[t] = fcur[t];
  return it;
}

} // extern "C"
