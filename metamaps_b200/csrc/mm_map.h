// mm_map.h -- K3 (read sketch), K4 (L1 candidate regions), K5 (L2 sliding MinHash).
//
// Replaces skch::Map::doL1Mapping / computeL1CandidateRegions / computeL2MappedRegions / doL2Mapping
// (reference src/map/include/computeMap.hpp:277-538), SlideMapper (slidingMap.hpp) and MIIteratorL2
// (MIIteratorL2.hpp:54-96).
//
// L2 restated.  SlideMapper keeps an ordered map over  Q (the read's s distinct minimizer hashes)  union
// W (the distinct hashes of the reference minimizers inside the current super-window) with a pivot on its
// s-th smallest key, and counts the keys <= pivot present on both sides.  Writing q_1<...<q_s for Q and
// cnt[g] for the number of distinct W-only hashes falling between q_g and q_{g+1} ("gap g"):
//     rank of q_i in the union   F(i) = i + sum_{g<i} cnt[g]
//     istar = max{ i : F(i) <= s }            (how many query minimizers are inside the bottom-s)
//     sharedSketchElements = #{ i <= istar : q_i in W }
// Only the per-gap counts matter, never the order of W-only hashes inside a gap.  Each window shift inserts
// and/or deletes one reference minimizer, which moves istar by at most one.  K5 therefore runs in three
// phases per candidate: (A) classify every reference minimizer of the candidate's span against the read
// sketch (binary search -> "match i" or "gap g"), (B) replay the reference's exact evaluate-then-advance
// loop over those codes with O(1) work per shift, (C) strand vote over the optimal window.
#pragma once
#include "mm_index.h"
#include "mm_stats.h"
#include "mm_stdsort.h"

namespace mm {

static const uint32_t CODE_MATCH = 0x80000000u, CODE_DUP = 0x40000000u, CODE_IDX = 0x3FFFFFFFu;

// ---------------------------------------------------------------------------------------------- K3
struct ReadKeyFn {          // entry e of the batch sketch -> (read << 32 | hash)
  const uint32_t* hash; const int64_t* seqOff; int32_t n_reads; uint64_t* key;
  MM_HD void operator()(int64_t e) const {
    int64_t r = upper_bound_idx(seqOff, (int64_t)n_reads + 1, e) - 1;
    key[e] = ((uint64_t)r << 32) | ldg(hash + e);
  }
};
struct HeadFlagFn {
  const uint64_t* key; int32_t* head; int64_t n;
  MM_HD void operator()(int64_t e) const { head[e] = (e < n && (e == 0 || ldg(key + e) != ldg(key + e - 1))) ? 1 : 0; }
};
struct UniqueScatterFn {    // std::unique keeps the first element of every equal-hash run (computeMap.hpp:295)
  const uint64_t* key; const uint32_t* ws; const int32_t* head; const int64_t* idx; uint32_t* qHash; uint8_t* qStrand;
  MM_HD void operator()(int64_t e) const {
    if (ldg(head + e)) { int64_t d = ldg(idx + e); qHash[d] = (uint32_t)ldg(key + e); qStrand[d] = (uint8_t)(ldg(ws + e) & 1u); }
  }
};
struct ReadSketchOffFn {    // sorting is within reads, so read r still owns [seqOff[r], seqOff[r+1])
  const int64_t* seqOff; const int64_t* idx; int64_t* qOff; int32_t* sOf; int32_t n_reads;
  MM_HD void operator()(int64_t r) const {
    int64_t a = ldg(idx + ldg(seqOff + r));
    qOff[r] = a;
    if (r < n_reads) sOf[r] = (int32_t)(ldg(idx + ldg(seqOff + r + 1)) - a);
  }
};

// Reads in which one hash occurs with both strands: the surviving copy depends on std::sort's permutation.
struct AmbigDetectFn {
  const uint64_t* key; const uint32_t* ws; const int32_t* head; int32_t* ambig; unsigned long long* count; int32_t* list; int64_t cap;
  MM_HD void operator()(int64_t e) const {
    if (e == 0 || ldg(head + e)) return;
    if (((ldg(ws + e) ^ ldg(ws + e - 1)) & 1u) == 0) return;
    int32_t r = (int32_t)(ldg(key + e) >> 32);
    if (atomic_cas_u32((uint32_t*)ambig + r, 0u, 1u) == 0u) {
      unsigned long long s = atomic_add_u64(count, 1ull);
      if ((int64_t)s < cap) list[s] = r;
    }
  }
};
// one item per ambiguous read: replay std::sort on the read's minimizers in emission order, then take the
// strand of the first element of every equal-hash run (std::unique)
struct AmbigResolveFn {
  const int32_t* list; const uint32_t* hash; const uint32_t* ws; const int64_t* seqOff; uint64_t* scratch;
  const int64_t* qOff; const uint32_t* qHash; uint8_t* qStrand;
  MM_HD void operator()(int64_t i) const {
    int32_t r = ldg(list + i);
    int64_t b = ldg(seqOff + r), n = ldg(seqOff + r + 1) - b;
    uint64_t* a = scratch + b;
    for (int64_t j = 0; j < n; j++) a[j] = ((uint64_t)ldg(hash + b + j) << 32) | ldg(ws + b + j);
    stdsort::sort(a, n);
    int64_t q = ldg(qOff + r);
    for (int64_t j = 0; j < n; j++) {
      if (j == 0 || (uint32_t)(a[j] >> 32) != (uint32_t)(a[j - 1] >> 32)) { qStrand[q] = (uint8_t)(a[j] & 1u); q++; }
    }
  }
};

// ---------------------------------------------------------------------------------------------- K4
struct ProbeFn {            // computeMap.hpp:307-321
  const Slot* table; uint32_t mask; const uint32_t* qHash; int32_t freqThreshold; int32_t* hitCnt; int64_t* hitStart; int64_t n;
  MM_HD void operator()(int64_t i) const {
    if (i >= n) { hitCnt[i] = 0; return; }
    int64_t st = 0;
    uint32_t c = table_find(table, mask, ldg(qHash + i), &st);
    if (c != 0 && (int64_t)c < (int64_t)freqThreshold) { hitCnt[i] = (int32_t)c; hitStart[i] = st; }
    else hitCnt[i] = 0;
  }
};
struct GatherHitsFn {
  const int32_t* hitCnt; const int64_t* hitStart; const int64_t* hitOff; const uint64_t* posKey; uint64_t* hits;
  MM_HD void operator()(int64_t i) const {
    int32_t c = ldg(hitCnt + i);
    if (!c) return;
    int64_t s = ldg(hitStart + i), d = ldg(hitOff + i);
    for (int32_t j = 0; j < c; j++) hits[d + j] = ldg(posKey + s + j);
  }
};
struct ReadHitOffFn {
  const int64_t* qOff; const int64_t* hitOff; int64_t* readHitOff;
  MM_HD void operator()(int64_t r) const { readHitOff[r] = ldg(hitOff + ldg(qOff + r)); }
};

// computeL1CandidateRegions (computeMap.hpp:346-386) over the read's sorted hits.  pass 0 counts, pass 1 writes.
struct CandidateFn {
  const uint64_t* hits; const int64_t* readHitOff; const int32_t* sOf; const int32_t* readLen;
  const int32_t* minHitsTab; int pass;
  int32_t* candCnt; const int64_t* candOff;
  int32_t* cRead; int32_t* cSeq; int32_t* cStart; int32_t* cEnd;
  MM_HD void operator()(int64_t r) const {
    int32_t s = ldg(sOf + r);
    int32_t n = 0;
    if (s > 0) {
      int64_t b = ldg(readHitOff + r), e = ldg(readHitOff + r + 1);
      int32_t mh = ldg(minHitsTab + s); if (mh < 1) mh = 1;
      int32_t len = ldg(readLen + r);
      int64_t out = pass ? ldg(candOff + r) : 0;
      int32_t lastSeq = -1, lastStart = 0, lastEnd = 0;
      for (int64_t i = b; i + mh - 1 < e; i++) {
        uint64_t ka = ldg(hits + i), kb = ldg(hits + i + mh - 1);
        int32_t sa = (int32_t)(ka >> 32), sb = (int32_t)(kb >> 32);
        int32_t wa = (int32_t)((uint32_t)ka >> 1), wb = (int32_t)((uint32_t)kb >> 1);
        if (sa == sb && wb - wa < len) {
          int32_t st = wb - len + 1; if (st < 0) st = 0;
          if (n > 0 && sa == lastSeq && lastEnd >= st) { if (wa > lastEnd) lastEnd = wa; }
          else {
            if (n > 0 && pass) { cRead[out + n - 1] = (int32_t)r; cSeq[out + n - 1] = lastSeq; cStart[out + n - 1] = lastStart; cEnd[out + n - 1] = lastEnd; }
            n++; lastSeq = sa; lastStart = st; lastEnd = wa;
          }
        }
      }
      if (n > 0 && pass) { cRead[out + n - 1] = (int32_t)r; cSeq[out + n - 1] = lastSeq; cStart[out + n - 1] = lastStart; cEnd[out + n - 1] = lastEnd; }
    }
    if (!pass) candCnt[r] = n;
  }
};

// ---------------------------------------------------------------------------------------------- K5
MM_HD int64_t search_index(const uint32_t* miWs, const int64_t* contigStart, int32_t seq, int64_t wpos) {   // winSketch.hpp:506-517
  int64_t lo = ldg(contigStart + seq), hi = ldg(contigStart + seq + 1);
  while (lo < hi) { int64_t m = (lo + hi) >> 1; if ((int64_t)(ldg(miWs + m) >> 1) < wpos) lo = m + 1; else hi = m; }
  return lo;
}
struct L2SetupFn {          // computeMap.hpp:465-480
  const uint32_t* miWs; const int64_t* contigStart; const int32_t* cRead; const int32_t* cSeq; const int32_t* cStart; const int32_t* cEnd;
  const int32_t* readLen; const int32_t* sOf; int k, w;
  int64_t* beg0; int64_t* fe; int64_t* le; int32_t* spanN; int64_t n;
  MM_HD void operator()(int64_t c) const {
    if (c >= n) { spanN[c] = 0; return; }
    int32_t r = ldg(cRead + c), sq = ldg(cSeq + c); int32_t len = ldg(readLen + r);
    int64_t b = search_index(miWs, contigStart, sq, ldg(cStart + c));
    int32_t cmw = len - (w - 1) - (k - 1);
    int64_t e = search_index(miWs, contigStart, sq, (int64_t)(ldg(miWs + b) >> 1) + cmw);
    int64_t l = search_index(miWs, contigStart, sq, (int64_t)ldg(cEnd + c) + len);
    beg0[c] = b; fe[c] = e; le[c] = l;
    spanN[c] = (int32_t)(l > b ? l - b : 0);
  }
};
// per-candidate sweep state: (s+1) gap counters + s match bits, in 32-bit words
struct StWordsFn {
  const int32_t* cRead; const int32_t* sOf; int32_t* stWords; int64_t n; int32_t cntBytes;
  MM_HD void operator()(int64_t c) const {
    if (c >= n) { stWords[c] = 0; return; }
    int32_t s = ldg(sOf + ldg(cRead + c));
    stWords[c] = ((s + 1) * cntBytes + 3) / 4 + (s + 31) / 32;
  }
};

// phase A: one item per reference minimizer of a candidate span
struct L2ClassifyFn {
  const uint32_t* miHash; const uint32_t* miWs; const uint32_t* dupBits;
  const int64_t* evOff; int64_t cand0, nCand; int64_t evBase;   // candidates [cand0, cand0+nCand), events relative to evBase
  const int64_t* beg0; const int32_t* cRead; const uint32_t* qHash; const int64_t* qOff; const int32_t* sOf;
  uint2* ev;
  MM_HD void operator()(int64_t t) const {
    int64_t c = cand0 + upper_bound_idx(evOff + cand0, nCand + 1, t + evBase) - 1;
    int64_t j = ldg(beg0 + c) + (t + evBase - ldg(evOff + c));
    int32_t r = ldg(cRead + c); int32_t s = ldg(sOf + r);
    const uint32_t* q = qHash + ldg(qOff + r);
    uint32_t h = ldg(miHash + j);
    int32_t lo = 0, hi = s;
    while (lo < hi) { int32_t m = (lo + hi) >> 1; if (ldg(q + m) < h) lo = m + 1; else hi = m; }
    uint32_t code = (lo < s && ldg(q + lo) == h) ? (CODE_MATCH | (uint32_t)(lo + 1)) : (uint32_t)lo;
    if ((ldg(dupBits + (j >> 5)) >> (j & 31)) & 1u) code |= CODE_DUP;
    ev[t] = make_uint2(code, ldg(miWs + j));
  }
};

MM_HD uint64_t dup_links(const uint32_t* dupIdx, const uint64_t* dupLinks, int64_t n_dup, int64_t j) {
  int64_t p = lower_bound_idx(dupIdx, n_dup, (uint32_t)j);
  return (p < n_dup && ldg(dupIdx + p) == (uint32_t)j) ? ldg(dupLinks + p) : 0ull;
}

// phase B: the evaluate-then-advance loop of computeL2MappedRegions (computeMap.hpp:482-533).
// CntT = uint16_t when every span of the pass has fewer than 65535 minimizers (a gap count can never
// exceed the span size), uint32_t otherwise.
template <class CntT>
struct L2SweepFn {
  const uint2* ev; const int64_t* evOff; int64_t evBase; uint32_t* state; const int64_t* stOff; int64_t stBase; int64_t cand0;
  const int64_t* beg0; const int64_t* fe; const int64_t* le; const int32_t* cRead; const int32_t* sOf; const int32_t* readLen;
  const uint32_t* dupIdx; const uint64_t* dupLinks; int64_t n_dup; int k, w;
  int32_t* oShared; int32_t* oPos; int32_t* oValid; int64_t* oOptS; int64_t* oOptE; int32_t* oIstar;

  struct St { CntT* cnt; uint32_t* mb; int32_t s, istar, C, shared; };
  MM_HD static void ins(St& z, uint32_t code) {
    if (code & CODE_MATCH) {
      int32_t i = (int32_t)(code & CODE_IDX);
      z.mb[(i - 1) >> 5] |= 1u << ((i - 1) & 31);
      if (i <= z.istar) z.shared++;
    } else {
      int32_t g = (int32_t)(code & CODE_IDX);
      if (g >= z.s) return;                   // above the largest query hash: can never enter the bottom-s
      z.cnt[g]++;
      if (g < z.istar) {
        z.C++;
        if (z.istar + z.C > z.s) {            // q_istar drops out of the bottom-s
          z.C -= z.cnt[z.istar - 1];
          if ((z.mb[(z.istar - 1) >> 5] >> ((z.istar - 1) & 31)) & 1u) z.shared--;
          z.istar--;
        }
      }
    }
  }
  MM_HD static void del(St& z, uint32_t code) {
    if (code & CODE_MATCH) {
      int32_t i = (int32_t)(code & CODE_IDX);
      z.mb[(i - 1) >> 5] &= ~(1u << ((i - 1) & 31));
      if (i <= z.istar) z.shared--;
    } else {
      int32_t g = (int32_t)(code & CODE_IDX);
      if (g >= z.s) return;
      z.cnt[g]--;
      if (g < z.istar) z.C--;
      if (z.istar < z.s && z.istar + 1 + z.C + (int32_t)z.cnt[z.istar] <= z.s) {   // q_{istar+1} enters
        z.C += z.cnt[z.istar];
        z.istar++;
        if ((z.mb[(z.istar - 1) >> 5] >> ((z.istar - 1) & 31)) & 1u) z.shared++;
      }
    }
  }
  MM_HD void operator()(int64_t ci) const {
    int64_t c = cand0 + ci;
    int32_t r = ldg(cRead + c); int32_t s = ldg(sOf + r); int32_t len = ldg(readLen + r);
    int64_t b0 = ldg(beg0 + c);
    const uint2* e = ev + (ldg(evOff + c) - evBase) - b0;        // e[j] for index position j
    uint32_t* stp = state + (ldg(stOff + c) - stBase);
    St z; z.cnt = (CntT*)stp; z.mb = stp + ((s + 1) * (int32_t)sizeof(CntT) + 3) / 4; z.s = s; z.istar = s; z.C = 0; z.shared = 0;
    int64_t beg = b0, end = ldg(fe + c), last = ldg(le + c);
    int32_t cmw = len - (w - 1) - (k - 1);
    // slidemap.insert_ref(sw_beg, sw_end) (computeMap.hpp:488); a hash already present is only revised
    for (int64_t j = beg; j < end; j++) {
      uint32_t code = e[j].x;
      if (code & CODE_DUP) { uint64_t l = dup_links(dupIdx, dupLinks, n_dup, j); uint32_t pd = (uint32_t)(l >> 32); if (pd && j - (int64_t)pd >= beg) continue; }
      ins(z, code);
    }
    int32_t best = 0, bpos = 0, lpos = 0, valid = 0, bistar = s; int64_t optS = 0, optE = 0;
    int64_t pb = beg, pe = end;
    int32_t sw_pos = (int32_t)(e[beg].y >> 1);
    while (end < last) {
      if (pb != beg) {                                   // delete_ref(prev_beg) (slidingMap.hpp:170-219)
        uint32_t code = e[pb].x; bool noop = false;
        if (code & CODE_DUP) { uint64_t l = dup_links(dupIdx, dupLinks, n_dup, pb); uint32_t nd = (uint32_t)l; if (nd && pb + (int64_t)nd < pe) noop = true; }
        if (!noop) del(z, code);
      }
      if (pe != end) {                                   // insert_ref(prev_end) (slidingMap.hpp:139-164)
        uint32_t code = e[pe].x; bool noop = false;
        if (code & CODE_DUP) { uint64_t l = dup_links(dupIdx, dupLinks, n_dup, pe); uint32_t pd = (uint32_t)(l >> 32); if (pd && pe - (int64_t)pd >= beg) noop = true; }
        if (!noop) ins(z, code);
      }
      int32_t wb = (int32_t)(e[beg].y >> 1);
      if (z.shared > best) { best = z.shared; optS = beg; optE = end; bpos = lpos = wb; valid = 1; bistar = z.istar; }
      else if (z.shared == best) lpos = wb;
      pb = beg; pe = end;
      int32_t nb = (int32_t)(e[beg + 1].y >> 1) - sw_pos;            // MIIteratorL2::next (MIIteratorL2.hpp:74-96)
      int32_t ne = (int32_t)(e[end].y >> 1) - (sw_pos + cmw - 1);
      int32_t adv = nb < ne ? nb : ne;
      sw_pos += adv;
      if (adv == nb) beg++;
      if (adv == ne) end++;
    }
    oShared[c] = best; oPos[c] = (bpos + lpos) / 2; oValid[c] = valid; oOptS[c] = optS; oOptE[c] = optE; oIstar[c] = bistar;
  }
};

// phase C: strand vote of the optimal window (computeMap.hpp:431-438, slidingMap.hpp:232-254)
struct L2StrandFn {
  const uint2* ev; const int64_t* evOff; int64_t evBase; int64_t cand0; const int64_t* beg0;
  const int32_t* cRead; const int64_t* qOff; const uint8_t* qStrand;
  const uint32_t* dupIdx; const uint64_t* dupLinks; int64_t n_dup;
  const int32_t* oValid; const int64_t* oOptS; const int64_t* oOptE; const int32_t* oIstar; int32_t* oVotes;
  MM_HD void operator()(int64_t ci) const {
    int64_t c = cand0 + ci;
    int32_t votes = 0;
    if (ldg(oValid + c)) {
      const uint2* e = ev + (ldg(evOff + c) - evBase) - ldg(beg0 + c);
      const uint8_t* qs = qStrand + ldg(qOff + ldg(cRead + c));
      int64_t a = ldg(oOptS + c), b = ldg(oOptE + c); int32_t istar = ldg(oIstar + c);
      for (int64_t j = a; j < b; j++) {
        uint2 v = e[j];
        if (!(v.x & CODE_MATCH)) continue;
        int32_t i = (int32_t)(v.x & CODE_IDX);
        if (i > istar) continue;
        // the map keeps the strand of the LAST inserted occurrence of a hash
        if (v.x & CODE_DUP) { uint64_t l = dup_links(dupIdx, dupLinks, n_dup, j); uint32_t nd = (uint32_t)l; if (nd && j + (int64_t)nd < b) continue; }
        int32_t sq = ldg(qs + i - 1) ? 1 : -1, sr = (v.y & 1u) ? 1 : -1;
        votes += sq * sr;
      }
    }
    oVotes[c] = votes;
  }
};
struct AcceptFn {           // computeMap.hpp:415 through the per-s threshold table
  const int32_t* cRead; const int32_t* sOf; const int32_t* acceptTab; const int32_t* oShared; const int32_t* oValid;
  int32_t* oAccept; int32_t* readMapped;
  MM_HD void operator()(int64_t c) const {
    int32_t r = ldg(cRead + c); int32_t s = ldg(sOf + r);
    int32_t a = (ldg(oValid + c) && ldg(oShared + c) >= ldg(acceptTab + s)) ? 1 : 0;
    oAccept[c] = a;
    if (a) readMapped[r] = 1;
  }
};

struct MapStats { double ms[8]; int64_t counters[8]; };

struct Mapper {
  Runtime& rt; Prims& pr; Sketcher& sk;
  stats::Tables tabs;
  DevBuf<int32_t> dMinHits, dAccept; int tabUploaded = 0; int tabK = 0; float tabPi = 0;
  // per batch (kept until the next batch for the fetch calls)
  SeqBatch batch; SketchOut rs;
  int32_t n_reads = 0; int64_t n_q = 0, n_hits = 0, n_cand = 0;
  DevBuf<int32_t> readLen, sOf, head, hitCnt, candCnt, cRead, cSeq, cStart, cEnd, spanN, stWords;
  DevBuf<int32_t> oShared, oPos, oValid, oIstar, oVotes, oAccept, readMapped;
  DevBuf<int64_t> qOff, idx, hitStart, hitOff, readHitOff, candOff, beg0, fe, le, evOff, stOff, oOptS, oOptE, scalar;
  DevBuf<uint64_t> key, key2, hits, hits2;
  DevBuf<uint32_t> ws2, qHash, state; DevBuf<uint8_t> qStrand; DevBuf<uint2> ev;
  DevBuf<int32_t> ambig, ambigList; DevBuf<unsigned long long> ambigCount; int64_t n_ambig = 0;
  std::vector<int32_t> h_effLen;
  MapStats st;
  int64_t evBudget = (int64_t)1 << 28;       // span elements classified per L2 pass (8 B each)

  Mapper(Runtime& r, Prims& p, Sketcher& s) : rt(r), pr(p), sk(s) { memset(&st, 0, sizeof(st)); }

  void ensure_tables(int k, float pi, int smax) {
    if (smax < 16) smax = 16;
    bool fresh = (k != tabK || pi != tabPi);
    if (fresh || smax + 1 > tabUploaded) {
      int want = fresh ? smax : std::max(smax, tabUploaded * 2);
      tabs.extend(k, pi, want);
      dMinHits.ensure(tabs.minHits.size()); dAccept.ensure(tabs.acceptMin.size());
      h2d(rt, dMinHits.p, tabs.minHits.data(), sizeof(int32_t) * tabs.minHits.size());
      h2d(rt, dAccept.p, tabs.acceptMin.data(), sizeof(int32_t) * tabs.acceptMin.size());
      tabUploaded = (int)tabs.minHits.size(); tabK = k; tabPi = pi;
    }
  }

  // `batch` must already be loaded (Sketcher::load)
  void run(const Index& ix, float pi, int32_t minReadLen, int64_t* summary /*6*/) {
    memset(&st, 0, sizeof(st));
    const int k = ix.k, w = ix.w;
    n_reads = batch.n_seqs;
    // reads shorter than w, k or -m are skipped (computeMap.hpp:137): hide them from K1 by zeroing their length
    h_effLen.assign((size_t)n_reads, 0);
    int64_t nShort = 0, basesOk = 0; int32_t maxLen = 0;
    std::vector<int32_t> saveLen = batch.h_len;
    for (int32_t i = 0; i < n_reads; i++) {
      int32_t L = batch.h_len[i];
      if (L < w || L < k || L < minReadLen) { nShort++; batch.h_len[i] = 0; }
      else { h_effLen[i] = L; basesOk += L; if (L > maxLen) maxLen = L; }
    }
    readLen.ensure((size_t)n_reads + 1); h2d(rt, readLen.p, h_effLen.data(), sizeof(int32_t) * (size_t)n_reads);
    {
      StageTimer t(rt, &st.ms[0]);
      sk.run(batch, k, w, rs);
    }
    batch.h_len = saveLen;
    // ---- K3: sort by (read, hash), unique
    int64_t nm = rs.n_total;
    sOf.ensure((size_t)n_reads + 1); qOff.ensure((size_t)n_reads + 2);
    {
      StageTimer t(rt, &st.ms[1]);
      key.ensure((size_t)nm + 1); key2.ensure((size_t)nm + 1); ws2.ensure((size_t)nm + 1); head.ensure((size_t)nm + 2); idx.ensure((size_t)nm + 2);
      foreach(rt, nm, ReadKeyFn{rs.hash.p, rs.seqOff.p, n_reads, key.p});
      int bits = 33; while (bits < 64 && ((int64_t)1 << (bits - 32)) < n_reads) bits++;
      pr.sort_pairs<uint64_t, uint32_t>(key.p, key2.p, rs.ws.p, ws2.p, nm, bits);
      foreach(rt, nm + 1, HeadFlagFn{key2.p, head.p, nm});
      pr.exclusive_sum<int32_t, int64_t>(head.p, idx.p, nm + 1);
      d2h(rt, &n_q, idx.p + nm, sizeof(int64_t));
      qHash.ensure((size_t)n_q + 1); qStrand.ensure((size_t)n_q + 1);
      foreach(rt, nm, UniqueScatterFn{key2.p, ws2.p, head.p, idx.p, qHash.p, qStrand.p});
      foreach(rt, (int64_t)n_reads + 1, ReadSketchOffFn{rs.seqOff.p, idx.p, qOff.p, sOf.p, n_reads});
      {   // duplicate hashes with both strands: settle the survivor like std::sort + std::unique would
        ambig.ensure((size_t)n_reads + 1); ambigList.ensure(4096); ambigCount.ensure(1);
        dev_memset(rt, ambig.p, 0, sizeof(int32_t) * ((size_t)n_reads + 1));
        dev_memset(rt, ambigCount.p, 0, sizeof(unsigned long long));
        foreach(rt, nm, AmbigDetectFn{key2.p, ws2.p, head.p, ambig.p, ambigCount.p, ambigList.p, 4096});
        unsigned long long na = 0; d2h(rt, &na, ambigCount.p, sizeof(na));
        if (na > 4096) {
          ambigList.ensure((size_t)na);
          dev_memset(rt, ambig.p, 0, sizeof(int32_t) * ((size_t)n_reads + 1));
          dev_memset(rt, ambigCount.p, 0, sizeof(unsigned long long));
          foreach(rt, nm, AmbigDetectFn{key2.p, ws2.p, head.p, ambig.p, ambigCount.p, ambigList.p, (int64_t)na});
        }
        n_ambig = (int64_t)na;
        if (na) foreach(rt, (int64_t)na, AmbigResolveFn{ambigList.p, rs.hash.p, rs.ws.p, rs.seqOff.p, key.p, qOff.p, qHash.p, qStrand.p}, 128, 16);
      }
      // minimumHits[s] / acceptMin[s] tables up to the largest sketch of the batch (host, map_stats.hpp)
      int32_t maxS = 0;
      if (n_reads > 0) { DevBuf<int32_t> m; m.ensure(1); pr.reduce_max<int32_t>(sOf.p, m.p, n_reads); d2h(rt, &maxS, m.p, sizeof(int32_t)); }
      ensure_tables(k, pi, maxS);
    }
    // ---- K4: probe, gather, sort, candidate regions
    candOff.ensure((size_t)n_reads + 2); candCnt.ensure((size_t)n_reads + 2);
    {
      StageTimer t(rt, &st.ms[2]);
      hitCnt.ensure((size_t)n_q + 2); hitStart.ensure((size_t)n_q + 2); hitOff.ensure((size_t)n_q + 2);
      foreach(rt, n_q + 1, ProbeFn{ix.table.p, ix.tableMask, qHash.p, ix.freqThreshold, hitCnt.p, hitStart.p, n_q});
      pr.exclusive_sum<int32_t, int64_t>(hitCnt.p, hitOff.p, n_q + 1);
      d2h(rt, &n_hits, hitOff.p + n_q, sizeof(int64_t));
      hits.ensure((size_t)n_hits + 1); hits2.ensure((size_t)n_hits + 1); readHitOff.ensure((size_t)n_reads + 2);
      foreach(rt, n_q, GatherHitsFn{hitCnt.p, hitStart.p, hitOff.p, ix.posKey.p, hits.p});
      foreach(rt, (int64_t)n_reads + 1, ReadHitOffFn{qOff.p, hitOff.p, readHitOff.p});
      pr.segmented_sort_keys<uint64_t>(hits.p, hits2.p, n_hits, n_reads, readHitOff.p);
      CandidateFn cf{hits2.p, readHitOff.p, sOf.p, readLen.p, dMinHits.p, 0, candCnt.p, candOff.p, nullptr, nullptr, nullptr, nullptr};
      foreach(rt, n_reads, cf);
      dev_memset(rt, candCnt.p + n_reads, 0, sizeof(int32_t));
      pr.exclusive_sum<int32_t, int64_t>(candCnt.p, candOff.p, (int64_t)n_reads + 1);
      d2h(rt, &n_cand, candOff.p + n_reads, sizeof(int64_t));
      cRead.ensure((size_t)n_cand + 1); cSeq.ensure((size_t)n_cand + 1); cStart.ensure((size_t)n_cand + 1); cEnd.ensure((size_t)n_cand + 1);
      cf.pass = 1; cf.cRead = cRead.p; cf.cSeq = cSeq.p; cf.cStart = cStart.p; cf.cEnd = cEnd.p;
      foreach(rt, n_reads, cf);
    }
    // ---- K5
    int64_t totalEv = 0;
    int32_t cntBytes = 2;
    oShared.ensure((size_t)n_cand + 1); oPos.ensure((size_t)n_cand + 1); oValid.ensure((size_t)n_cand + 1); oIstar.ensure((size_t)n_cand + 1);
    oVotes.ensure((size_t)n_cand + 1); oAccept.ensure((size_t)n_cand + 1); oOptS.ensure((size_t)n_cand + 1); oOptE.ensure((size_t)n_cand + 1);
    readMapped.ensure((size_t)n_reads + 1); dev_memset(rt, readMapped.p, 0, sizeof(int32_t) * ((size_t)n_reads + 1));
    if (n_cand > 0) {
      beg0.ensure((size_t)n_cand + 1); fe.ensure((size_t)n_cand + 1); le.ensure((size_t)n_cand + 1);
      spanN.ensure((size_t)n_cand + 2); stWords.ensure((size_t)n_cand + 2); evOff.ensure((size_t)n_cand + 2); stOff.ensure((size_t)n_cand + 2);
      {
        StageTimer t(rt, &st.ms[3]);
        foreach(rt, n_cand + 1, L2SetupFn{ix.miWs.p, ix.contigStart.p, cRead.p, cSeq.p, cStart.p, cEnd.p, readLen.p, sOf.p, k, w,
                                           beg0.p, fe.p, le.p, spanN.p, n_cand});
        pr.exclusive_sum<int32_t, int64_t>(spanN.p, evOff.p, n_cand + 1);
      }
      std::vector<int64_t> hEv((size_t)n_cand + 1), hSt((size_t)n_cand + 1);
      d2h(rt, hEv.data(), evOff.p, sizeof(int64_t) * hEv.size());
      totalEv = hEv[(size_t)n_cand];
      // a gap counter never exceeds the number of minimizers in the span: 16 bits unless some span is huge
      for (int64_t c = 0; c < n_cand; c++) if (hEv[(size_t)c + 1] - hEv[(size_t)c] >= 65535) { cntBytes = 4; break; }
      foreach(rt, n_cand + 1, StWordsFn{cRead.p, sOf.p, stWords.p, n_cand, cntBytes});
      pr.exclusive_sum<int32_t, int64_t>(stWords.p, stOff.p, n_cand + 1);
      d2h(rt, hSt.data(), stOff.p, sizeof(int64_t) * hSt.size());
      int64_t c0 = 0;
      while (c0 < n_cand) {          // passes bounded by the event budget
        int64_t c1 = c0 + 1;
        while (c1 < n_cand && hEv[(size_t)c1 + 1] - hEv[(size_t)c0] <= evBudget) c1++;
        int64_t nEv = hEv[(size_t)c1] - hEv[(size_t)c0], nSt = hSt[(size_t)c1] - hSt[(size_t)c0], nc = c1 - c0;
        ev.ensure((size_t)nEv + 1); state.ensure((size_t)nSt + 1);
        dev_memset(rt, state.p, 0, sizeof(uint32_t) * (size_t)nSt);
        {
          StageTimer t(rt, &st.ms[4]);
          foreach(rt, nEv, L2ClassifyFn{ix.miHash.p, ix.miWs.p, ix.dupBits.p, evOff.p, c0, nc, hEv[(size_t)c0], beg0.p, cRead.p,
                                        qHash.p, qOff.p, sOf.p, ev.p});
        }
        {
          StageTimer t(rt, &st.ms[5]);
          if (cntBytes == 2)
            foreach(rt, nc, L2SweepFn<uint16_t>{ev.p, evOff.p, hEv[(size_t)c0], state.p, stOff.p, hSt[(size_t)c0], c0, beg0.p, fe.p, le.p, cRead.p, sOf.p,
                                                readLen.p, ix.dupIdx.p, ix.dupLinks.p, ix.n_dup, k, w, oShared.p, oPos.p, oValid.p, oOptS.p, oOptE.p, oIstar.p},
                    128, 16);
          else
            foreach(rt, nc, L2SweepFn<uint32_t>{ev.p, evOff.p, hEv[(size_t)c0], state.p, stOff.p, hSt[(size_t)c0], c0, beg0.p, fe.p, le.p, cRead.p, sOf.p,
                                                readLen.p, ix.dupIdx.p, ix.dupLinks.p, ix.n_dup, k, w, oShared.p, oPos.p, oValid.p, oOptS.p, oOptE.p, oIstar.p},
                    128, 16);
        }
        {
          StageTimer t(rt, &st.ms[6]);
          foreach(rt, nc, L2StrandFn{ev.p, evOff.p, hEv[(size_t)c0], c0, beg0.p, cRead.p, qOff.p, qStrand.p, ix.dupIdx.p, ix.dupLinks.p, ix.n_dup,
                                     oValid.p, oOptS.p, oOptE.p, oIstar.p, oVotes.p}, 128, 16);
        }
        c0 = c1;
      }
      foreach(rt, n_cand, AcceptFn{cRead.p, sOf.p, dAccept.p, oShared.p, oValid.p, oAccept.p, readMapped.p});
    }
    scalar.ensure(4);
    int64_t nMap = 0, nReadsMapped = 0;
    if (n_cand > 0) {
      DevBuf<int32_t> red; red.ensure(2);
      pr.reduce_sum<int32_t>(oAccept.p, red.p, n_cand);
      pr.reduce_sum<int32_t>(readMapped.p, red.p + 1, n_reads);
      int32_t h[2]; d2h(rt, h, red.p, sizeof(h)); nMap = h[0]; nReadsMapped = h[1];
    }
    rt.sync();
    st.counters[0] = n_q; st.counters[1] = n_hits; st.counters[2] = n_cand; st.counters[3] = totalEv; st.counters[4] = nMap;
    st.counters[5] = rs.n_total; st.counters[6] = basesOk; st.counters[7] = batch.n_exc;
    summary[0] = n_reads; summary[1] = nShort; summary[2] = n_cand; summary[3] = nMap; summary[4] = nReadsMapped; summary[5] = basesOk;
  }
};

}  // namespace mm
