// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY.
// Thin C interface over the UNMODIFIED reference headers (compiled from /root/reference/src where
// they lie, with oracle/boost_shim standing in for Boost).  `#define private public` exposes the
// private L1/L2 methods of skch::Map and the index of skch::Sketch so that the restatement in
// mm_oracle.cpp can be pinned function by function.  Built into oracle/_ref/libmm_refharness.so.
// the same standard headers mash_map.cpp pulls in before mapWrap.h (mash_map.cpp:7-16)
#include <iostream>
#include <ctime>
#include <cmath>
#include <chrono>
#include <functional>
#include <fstream>
#include <exception>
#include <stdexcept>
#include <assert.h>
#include <cstring>
#include <cstdint>
#include <sstream>
#include <set>
#include <map>
#include <unordered_map>
#include <deque>
#include <vector>
#include <algorithm>
#include <regex>
#define private public
#define protected public
#include "map/mapWrap.h"
#undef private
#undef protected

namespace {
struct RefState {
  skch::Parameters param;
  skch::Sketch* sketch = nullptr;
  skch::Map* map = nullptr;
};
}

extern "C" {

uint32_t mmr_hash(const char* kmer, int k) { return skch::CommonFunc::getHash(kmer, k); }

int64_t mmr_minimizers(const char* seq, int len, int k, int w, uint32_t* hash, int32_t* wpos, int32_t* strand, int64_t cap) {
  std::vector<skch::MinimizerInfo> v;
  std::string s(seq, seq + len);
  if (!(len < w || len < k)) skch::CommonFunc::addMinimizers(v, &s[0], len, k, w, 4, 0);
  for (int64_t i = 0; i < (int64_t)v.size() && i < cap; i++) { hash[i] = v[i].hash; wpos[i] = v[i].wpos; strand[i] = v[i].strand; }
  return (int64_t)v.size();
}
int mmr_min_hits_relaxed(int s, int k, float pi) { return skch::Stat::estimateMinimumHitsRelaxed(s, k, pi); }
int mmr_recommended_window(double p, int k, int alphabet, float pi, int lenQ, uint64_t lenR) { return skch::Stat::recommendedWindowSize(p, k, alphabet, pi, lenQ, lenR); }
void mmr_identity(int shared, int s, int k, float* nucIdentity, float* upper) {
  float md = skch::Stat::j2md(1.0 * shared / s, k);
  float lb = skch::Stat::md_lower_bound(md, s, k, 0.9);
  *nucIdentity = 100 * (1 - md);
  *upper = 100 * (1 - lb);
}
double mmr_likelihood(int k, int n_kmers, double identity, int sketch, int inter) { return mapWrap::likelihood_observed_set_sizes(k, n_kmers, identity, sketch, inter); }

// Build the reference index from a FASTA file (no memory limit) and keep a Map object whose
// constructor saw no query files, so its private L1/L2 methods can be driven read by read.
void* mmr_index_build(const char* fasta, int k, int w, int minReadLen, float pi) {
  RefState* st = new RefState();
  skch::Parameters& p = st->param;
  p.kmerSize = k; p.windowSize = w; p.minReadLength = minReadLen; p.alphabetSize = 4;
  p.referenceSize = 0; p.percentageIdentity = pi; p.p_value = 1e-3; p.threads = 1;
  p.refSequences = {std::string(fasta)}; p.outFileName = ""; p.reportAll = true; p.maximumMemory = 0;
  std::function<void(skch::Sketch*, size_t)> noop = [](skch::Sketch*, size_t) {};
  st->sketch = new skch::Sketch(p, 0, &noop);
  p.outFileName = "/dev/null";
  st->map = new skch::Map(p, *st->sketch);
  return st;
}
void mmr_index_free(void* h) { RefState* st = (RefState*)h; delete st->map; delete st->sketch; delete st; }
int64_t mmr_index_size(void* h) { return (int64_t)((RefState*)h)->sketch->minimizerIndex.size(); }
int64_t mmr_index_unique(void* h) { return (int64_t)((RefState*)h)->sketch->minimizerPosLookupIndex.size(); }
int mmr_index_freq_threshold(void* h) { return ((RefState*)h)->sketch->getFreqThreshold(); }
void mmr_index_get(void* h, uint32_t* hash, int32_t* seqId, int32_t* wpos, int32_t* strand) {
  auto& mi = ((RefState*)h)->sketch->minimizerIndex;
  for (size_t i = 0; i < mi.size(); i++) { hash[i] = mi[i].hash; seqId[i] = mi[i].seqId; wpos[i] = mi[i].wpos; strand[i] = mi[i].strand; }
}

int mmr_map_read(void* h, const char* seq, int len, int64_t* info,
                 int32_t* c_seq, int32_t* c_start, int32_t* c_end,
                 int32_t* l2_pos, int32_t* l2_shared, int32_t* l2_votes, int32_t* l2_valid,
                 int64_t* l2_optStart, int64_t* l2_optEnd, int cap) {
  RefState* st = (RefState*)h;
  typedef skch::Sketch::MI_Type MinVec;
  skch::QueryMetaData<MinVec> Q;
  std::string s(seq, seq + len);
  Q.seq = &s[0]; Q.len = len; Q.seqCounter = 0; Q.sketchSize = 0;
  std::vector<skch::Map::L1_candidateLocus_t> l1;
  st->map->doL1Mapping(Q, l1);
  info[0] = Q.sketchSize;
  info[1] = Q.sketchSize > 0 ? skch::Stat::estimateMinimumHitsRelaxed(Q.sketchSize, st->param.kmerSize, st->param.percentageIdentity) : 0;
  info[2] = (int64_t)l1.size(); info[3] = -1;
  auto base = st->sketch->minimizerIndex.begin();
  for (int i = 0; i < (int)l1.size() && i < cap; i++) {
    skch::Map::L2_mapLocus_t l2 = {};
    st->map->computeL2MappedRegions(Q, l1[i], l2);
    c_seq[i] = l1[i].seqId; c_start[i] = l1[i].rangeStartPos; c_end[i] = l1[i].rangeEndPos;
    l2_shared[i] = l2.sharedSketchSize; l2_valid[i] = l2.sharedSketchSize > 0;
    l2_pos[i] = l2.meanOptimalPos;
    l2_votes[i] = 0; l2_optStart[i] = 0; l2_optEnd[i] = 0;
    if (l2.sharedSketchSize > 0) {
      l2_optStart[i] = l2.optimalStart - base; l2_optEnd[i] = l2.optimalEnd - base;
      skch::SlideMapper<skch::QueryMetaData<MinVec>> sm(Q);
      sm.insert_ref(l2.optimalStart, l2.optimalEnd);
      int votes, uniq; sm.computeStatistics(votes, uniq);
      l2_votes[i] = votes;
    }
  }
  return (int)l1.size();
}
int mmr_read_sketch(void* h, const char* seq, int len, uint32_t* hash, int32_t* wpos, int32_t* strand, int cap) {
  RefState* st = (RefState*)h;
  typedef skch::Sketch::MI_Type MinVec;
  skch::QueryMetaData<MinVec> Q;
  std::string s(seq, seq + len);
  Q.seq = &s[0]; Q.len = len; Q.seqCounter = 0; Q.sketchSize = 0;
  std::vector<skch::Map::L1_candidateLocus_t> l1;
  st->map->doL1Mapping(Q, l1);
  for (int i = 0; i < Q.sketchSize && i < cap; i++) { hash[i] = Q.minimizerTableQuery[i].hash; wpos[i] = Q.minimizerTableQuery[i].wpos; strand[i] = Q.minimizerTableQuery[i].strand; }
  return Q.sketchSize;
}
} // extern "C"
