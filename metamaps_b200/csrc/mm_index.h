// mm_index.h -- K2: the GPU-resident reference index.
//
// Replaces skch::Sketch (reference src/map/include/winSketch.hpp):
//   minimizerIndex           (:129)  -> miHash[] / miWs[]  SoA in (seqId,wpos) order, 8 B per minimizer,
//                                       + contigStart[] (first entry of each contig)
//   minimizerPosLookupIndex  (:119)  -> open-addressed table of 16-byte slots {hash, count, start} over
//                                       posKey[] (CSR, entries of one hash contiguous, (seqId,wpos) order)
//   computeFreqHist          (:452-495) -> freqThreshold
//   searchIndex              (:506-517) -> contigStart[seq] + lower_bound over the contig's wpos
//
// Extra, not in the reference: a "dup" side structure.  SlideMapper (slidingMap.hpp:139-219) keeps ONE
// entry per hash, so when the same hash occurs twice inside a contig the L2 window must treat it with set
// semantics.  dupBits marks minimizers that share their hash with another minimizer of the same contig and
// dupIdx/dupPrev/dupNext give the distance (in index entries) to the previous/next such occurrence.
#pragma once
#include "mm_sketch.h"
#include <algorithm>

namespace mm {

struct Slot { uint32_t key, count, startLo, startHi; };

// minimizer hashes are biased towards small values (they are window minima), so spread them first
MM_HD uint32_t slot_of(uint32_t h, uint32_t mask) { return (uint32_t)(((uint64_t)h * 0x9E3779B97F4A7C15ull) >> 32) & mask; }

// probe one hash: returns count (0 = absent) and the CSR start
MM_HD uint32_t table_find(const Slot* table, uint32_t mask, uint32_t h, int64_t* start) {
  uint32_t s = slot_of(h, mask);
  for (;;) {
#if defined(__CUDA_ARCH__)
    uint4 v = __ldg(reinterpret_cast<const uint4*>(table + s));
    uint32_t key = v.x, cnt = v.y, lo = v.z, hi = v.w;
#else
    uint32_t key = table[s].key, cnt = table[s].count, lo = table[s].startLo, hi = table[s].startHi;
#endif
    if (cnt == 0) return 0;
    if (key == h) { *start = (int64_t)(((uint64_t)hi << 32) | lo); return cnt; }
    s = (s + 1) & mask;
  }
}

struct IotaFn { uint32_t* v; MM_HD void operator()(int64_t i) const { v[i] = (uint32_t)i; } };

struct TableInsertFn {
  Slot* table; uint32_t mask; const uint32_t* uniq; const int32_t* counts; const int64_t* starts;
  MM_HD void operator()(int64_t u) const {
    uint32_t h = ldg(uniq + u); uint32_t c = (uint32_t)ldg(counts + u); int64_t st = ldg(starts + u);
    uint32_t s = slot_of(h, mask);
    for (;;) {
      uint32_t old = atomic_cas_u32(&table[s].count, 0u, c);   // hashes are unique: claiming a slot is enough
      if (old == 0) { table[s].key = h; table[s].startLo = (uint32_t)st; table[s].startHi = (uint32_t)((uint64_t)st >> 32); return; }
      s = (s + 1) & mask;
    }
  }
};

// CSR entry e (hash order, then index order) -> (seqId << 32) | (wpos << 1 | strandbit): sorts like the
// reference's MinimizerMetaData::operator< (base_types.hpp:100-103; REV = -1 < FWD = 1)
struct PosKeyFn {
  const uint32_t* sortedPos; const uint32_t* miWs; const int64_t* contigStart; int32_t n_contigs; uint64_t* posKey;
  uint16_t* posSeq16;      // optional 2-byte copy of the contig id (the L1 contig filter streams this instead of posKey)
  MM_HD void operator()(int64_t e) const {
    uint32_t p = ldg(sortedPos + e);
    int64_t sq = upper_bound_idx(contigStart, (int64_t)n_contigs + 1, (int64_t)p) - 1;
    posKey[e] = ((uint64_t)sq << 32) | ldg(miWs + p);
    if (posSeq16) posSeq16[e] = (uint16_t)sq;
  }
};

// pass 0: mark + count dups; pass 1: also append (p, prev, next)
struct DupFn {
  const uint32_t* sortedHash; const uint32_t* sortedPos; const uint64_t* posKey; int64_t n;
  uint32_t* dupBits; unsigned long long* count; int pass;
  uint32_t* outIdx; uint64_t* outLinks;
  MM_HD void operator()(int64_t e) const {
    uint32_t h = ldg(sortedHash + e); uint32_t sq = (uint32_t)(ldg(posKey + e) >> 32); uint32_t p = ldg(sortedPos + e);
    uint32_t prev = 0, next = 0;
    if (e > 0 && ldg(sortedHash + e - 1) == h && (uint32_t)(ldg(posKey + e - 1) >> 32) == sq) prev = p - ldg(sortedPos + e - 1);
    if (e + 1 < n && ldg(sortedHash + e + 1) == h && (uint32_t)(ldg(posKey + e + 1) >> 32) == sq) next = ldg(sortedPos + e + 1) - p;
    if (prev | next) {
      unsigned long long s = atomic_add_u64(count, 1ull);
      if (pass == 0) atomic_or_u32(dupBits + (p >> 5), 1u << (p & 31));
      else { outIdx[s] = p; outLinks[s] = ((uint64_t)prev << 32) | next; }
    }
  }
};

// dupRB[i] = {dupBits word i, number of duplicate bits set before word i}: the link record of minimizer j is
// dupLinks[rank + popc(bits below j)] -- two dependent loads where a binary search over dupIdx took ~17 (it was 22 % of the
// instructions K5a issued, almost all with a single active lane)
MM_HD uint32_t popc32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return (uint32_t)__popc(x);
#else
  return (uint32_t)__builtin_popcount(x);
#endif
}
struct DupPopcFn { const uint32_t* bits; int32_t* out; MM_HD void operator()(int64_t i) const { out[i] = (int32_t)popc32(ldg(bits + i)); } };
struct DupRBFn {
  const uint32_t* bits; const int64_t* rank; uint2* out;
  MM_HD void operator()(int64_t i) const { out[i] = make_uint2(ldg(bits + i), (uint32_t)ldg(rank + i)); }
};

// bit 31 of a slot's count: the hash is over-frequent in the WHOLE (sharded) reference, see Index::keepUnique; such a
// count compares above every threshold, so the L1 probe skips the hash without knowing about the flag
static const uint32_t SLOT_GLOBAL_FREQ = 0x80000000u;
struct LookupFn {
  const Slot* table; uint32_t mask; const uint32_t* hashes; int32_t* counts;
  MM_HD void operator()(int64_t i) const { int64_t st; counts[i] = (int32_t)(table_find(table, mask, ldg(hashes + i), &st) & ~SLOT_GLOBAL_FREQ); }
};
struct GatherStrideFn {     // out[j] = in[j * step]
  const uint32_t* in; int64_t step; uint32_t* out;
  MM_HD void operator()(int64_t j) const { out[j] = ldg(in + j * step); }
};
struct LowerBoundFn {       // out[j] = first index of the sorted array with a[i] >= key[j]
  const uint32_t* a; int64_t n; const uint32_t* key; int64_t* out;
  MM_HD void operator()(int64_t j) const { out[j] = lower_bound_idx(a, n, ldg(key + j)); }
};
struct RunSumFn {           // gc[u] = sum of the values of run u
  const uint32_t* val; const int64_t* runStart; const int32_t* runLen; uint32_t* gc;
  MM_HD void operator()(int64_t u) const {
    int64_t b = ldg(runStart + u); uint64_t acc = 0;
    for (int32_t i = 0; i < ldg(runLen + u); i++) acc += ldg(val + b + i);
    gc[u] = acc > 0x7FFFFFFFull ? 0x7FFFFFFFu : (uint32_t)acc;
  }
};
struct GlobalCountFn {      // local unique hash lo+i -> its count in the merged (all shards) list of this hash range
  const uint32_t* uHash; int64_t lo; const uint32_t* mergedHash; const uint32_t* mergedCount; int64_t nMerged; uint32_t* out;
  MM_HD void operator()(int64_t i) const {
    const uint32_t h = ldg(uHash + lo + i);
    const int64_t p = lower_bound_idx(mergedHash, nMerged, h);
    out[lo + i] = ldg(mergedCount + p);           // present by construction
  }
};
struct FlagFreqFn {         // mark the slots of the hashes whose global count reaches the threshold
  Slot* table; uint32_t mask; const uint32_t* uHash; const uint32_t* gCount; uint32_t threshold;
  MM_HD void operator()(int64_t i) const {
    if (ldg(gCount + i) < threshold) return;
    const uint32_t h = ldg(uHash + i);
    uint32_t s = slot_of(h, mask);
    while (table[s].key != h || table[s].count == 0) s = (s + 1) & mask;
    table[s].count |= SLOT_GLOBAL_FREQ;
  }
};

struct Index {
  Runtime& rt; Prims& pr; Sketcher& sk;
  int k, w;
  bool finalized = false;
  // positional arrays
  DevBuf<uint32_t> miHash, miWs; int64_t n = 0;
  std::vector<int64_t> h_contigStart{0}; std::vector<int32_t> h_contigLen;
  DevBuf<int64_t> contigStart; DevBuf<int32_t> contigLen; int32_t n_contigs = 0;
  // lookup
  DevBuf<Slot> table; uint32_t tableMask = 0;
  DevBuf<uint64_t> posKey;
  DevBuf<uint16_t> posSeq16; bool hasSeq16 = false;     // contig id of every CSR entry when n_contigs <= 65536
  int64_t n_unique = 0; int32_t freqThreshold = 0x7fffffff;
  // contig-range shard of a larger reference: the sorted unique hashes and their local counts are kept after finalize() so
  // that mm_index_sync_threshold can derive the occurrence threshold of the WHOLE reference (winSketch.hpp:452-495)
  bool keepUnique = false; int32_t firstContig = 0; bool globalSynced = false; int32_t globalThreshold = 0x7fffffff;
  // The reference under --maxmemory builds one Sketch object chunk after chunk WITHOUT resetting its occurrence histogram or its
  // threshold (only minimizerIndex, minimizerPosLookupIndex and metadata are cleared, winSketch.hpp:302-304): chunk N's histogram is
  // added to those of chunks 0..N-1 (:458-459), the walk uses chunk N's own unique count (:465-466), and freqThreshold keeps its
  // previous value when the first bucket already overshoots (:471-488).  carryHist / carryThreshold = that state before this chunk,
  // histCum = after it (mm_index_set_freq_carry / mm_index_get_freq_hist).
  std::vector<std::pair<uint32_t, int64_t>> carryHist, histOwn, histCum; int32_t carryThreshold = 0x7fffffff;
  // histCum / freqThreshold from histOwn + the carry (finalize, and again whenever the carry is set on a finalized index: ranks
  // that build their chunks concurrently learn the earlier chunks' histograms afterwards)
  void apply_carry() {
    if (n == 0) { histCum = carryHist; freqThreshold = carryThreshold; return; }      // computeFreqHist does nothing on an empty chunk (:455)
    histCum = carryHist;
    histCum.insert(histCum.end(), histOwn.begin(), histOwn.end());
    std::sort(histCum.begin(), histCum.end());
    { std::vector<std::pair<uint32_t, int64_t>> m; for (auto& p : histCum) { if (!m.empty() && m.back().first == p.first) m.back().second += p.second; else m.push_back(p); } histCum.swap(m); }
    float percentageThreshold = 0.001f;
    int64_t toIgnore = (int64_t)(n_unique * percentageThreshold / 100);
    int64_t sum = 0;
    freqThreshold = carryThreshold;
    for (int64_t b = (int64_t)histCum.size() - 1; b >= 0; b--) {
      sum += histCum[(size_t)b].second;
      if (sum < toIgnore) freqThreshold = (int32_t)histCum[(size_t)b].first;
      else if (sum == toIgnore) { freqThreshold = (int32_t)histCum[(size_t)b].first; break; }
      else break;
    }
  }
  DevBuf<uint32_t> uHash; DevBuf<int32_t> uCnt;
  // dups
  DevBuf<uint32_t> dupBits, dupIdx; DevBuf<uint64_t> dupLinks; DevBuf<uint2> dupRB; int64_t n_dup = 0;
  int64_t total_bases = 0;

  Index(Runtime& r, Prims& p, Sketcher& s, int k_, int w_) : rt(r), pr(p), sk(s), k(k_), w(w_) {}

  void add(SeqBatch& B) {
    if (finalized) throw Error(-22, "index already finalized");
    SketchOut out;
    sk.run(B, k, w, out);
    if ((uint64_t)(n + out.n_total) >= 0xFFFFFFF0ull) throw Error(-34, "index shard exceeds 2^32 minimizers: split the reference into shards");
    miHash.grow(rt, (size_t)(n + out.n_total), (size_t)n); miWs.grow(rt, (size_t)(n + out.n_total), (size_t)n);
    d2d(rt, miHash.p + n, out.hash.p, sizeof(uint32_t) * (size_t)out.n_total);
    d2d(rt, miWs.p + n, out.ws.p, sizeof(uint32_t) * (size_t)out.n_total);
    std::vector<int64_t> off((size_t)B.n_seqs + 1);
    d2h(rt, off.data(), out.seqOff.p, sizeof(int64_t) * off.size());
    off[B.n_seqs] = out.n_total;
    for (int32_t i = 0; i < B.n_seqs; i++) {
      h_contigLen.push_back(B.h_len[i]);
      h_contigStart.push_back(n + off[i + 1]);
    }
    n += out.n_total; n_contigs += B.n_seqs; total_bases += B.total_bases;
    rt.sync();
  }

  void finalize() {
    if (finalized) return;
    contigStart.ensure(h_contigStart.size()); h2d(rt, contigStart.p, h_contigStart.data(), sizeof(int64_t) * h_contigStart.size());
    contigLen.ensure(h_contigLen.size() + 1); h2d(rt, contigLen.p, h_contigLen.data(), sizeof(int32_t) * h_contigLen.size());
    finalized = true;
    if (n == 0) apply_carry();
    if (n == 0) {          // no minimizers (no contigs, or all shorter than w / k): every array exists, zeroed, so that save / load / map work
      table.ensure(1024); dev_memset(rt, table.p, 0, sizeof(Slot) * 1024); tableMask = 1023;
      miHash.ensure(1); miWs.ensure(1); posKey.ensure(1); hasSeq16 = n_contigs <= 65536; if (hasSeq16) posSeq16.ensure(8);
      dupBits.ensure(2); dev_memset(rt, dupBits.p, 0, sizeof(uint32_t) * 2); dupIdx.ensure(1); dupLinks.ensure(1);
      build_dup_rank();
      return;
    }
    DevBuf<uint32_t> iota, sortedHash, sortedPos, uniq, cntSorted, cntUniq;
    DevBuf<int32_t> counts, cntRuns; DevBuf<int64_t> starts, nRuns;
    iota.ensure((size_t)n); sortedHash.ensure((size_t)n); sortedPos.ensure((size_t)n);
    foreach(rt, n, IotaFn{iota.p});
    pr.sort_pairs<uint32_t, uint32_t>(miHash.p, sortedHash.p, iota.p, sortedPos.p, n);
    iota.release();
    uniq.ensure((size_t)n); counts.ensure((size_t)n + 1); nRuns.ensure(2);
    pr.rle<uint32_t>(sortedHash.p, uniq.p, counts.p, nRuns.p, n);
    d2h(rt, &n_unique, nRuns.p, sizeof(int64_t));
    starts.ensure((size_t)n_unique + 1);
    dev_memset(rt, counts.p + n_unique, 0, sizeof(int32_t));
    pr.exclusive_sum<int32_t, int64_t>(counts.p, starts.p, n_unique + 1);

    // frequency threshold (computeFreqHist, winSketch.hpp:452-495): histogram of occurrences-per-hash,
    // walked from the most frequent bucket while the cumulative count stays within 0.001 % of the hashes
    {
      cntSorted.ensure((size_t)n_unique); cntUniq.ensure((size_t)n_unique); cntRuns.ensure((size_t)n_unique + 1);
      pr.sort_keys<uint32_t>((const uint32_t*)counts.p, cntSorted.p, n_unique);
      pr.rle<uint32_t>(cntSorted.p, cntUniq.p, cntRuns.p, nRuns.p, n_unique);
      int64_t nb = 0; d2h(rt, &nb, nRuns.p, sizeof(int64_t));
      std::vector<uint32_t> hv((size_t)nb); std::vector<int32_t> hc((size_t)nb);
      d2h(rt, hv.data(), cntUniq.p, sizeof(uint32_t) * (size_t)nb);
      d2h(rt, hc.data(), cntRuns.p, sizeof(int32_t) * (size_t)nb);
      // this chunk's histogram on top of the carried one (empty unless the host walks reference chunks)
      histOwn.clear();
      for (int64_t b = 0; b < nb; b++) histOwn.emplace_back(hv[(size_t)b], (int64_t)hc[(size_t)b]);
      apply_carry();
      cntSorted.release(); cntUniq.release(); cntRuns.release();
    }

    // load factor <= 0.5.  (MM_TABLE_MULT=4 halves it: measured on config 2, the probe kernel went 2.90 -> 2.71 ms for 10 GB more
    // table -- its 135 B of DRAM traffic per probe are 64-byte bursts around random 16-byte slots, not probe-sequence length)
    uint64_t mult = 2; if (const char* e = getenv("MM_TABLE_MULT")) { int v = atoi(e); if (v >= 2 && v <= 16) mult = (uint64_t)v; }
    uint64_t slots = 1024; while (slots < (uint64_t)n_unique * mult) slots <<= 1;
    if (slots > (1ull << 32)) { slots = 1024; while (slots < (uint64_t)n_unique * 2) slots <<= 1; }
    if (slots > (1ull << 32)) throw Error(-34, "hash table too large");
    table.ensure((size_t)slots); dev_memset(rt, table.p, 0, sizeof(Slot) * (size_t)slots);
    tableMask = (uint32_t)(slots - 1);
    foreach(rt, n_unique, TableInsertFn{table.p, tableMask, uniq.p, counts.p, starts.p});
    if (keepUnique) { std::swap(uHash.p, uniq.p); std::swap(uHash.cap, uniq.cap); std::swap(uCnt.p, counts.p); std::swap(uCnt.cap, counts.cap); }
    uniq.release(); counts.release(); starts.release();

    posKey.ensure((size_t)n);
    hasSeq16 = n_contigs <= 65536;
    if (hasSeq16) posSeq16.ensure((size_t)n + 8);
    foreach(rt, n, PosKeyFn{sortedPos.p, miWs.p, contigStart.p, n_contigs, posKey.p, hasSeq16 ? posSeq16.p : nullptr});

    dupBits.ensure((size_t)(n / 32 + 2)); dev_memset(rt, dupBits.p, 0, sizeof(uint32_t) * (size_t)(n / 32 + 2));
    DevBuf<unsigned long long> cnt; cnt.ensure(1); dev_memset(rt, cnt.p, 0, sizeof(unsigned long long));
    foreach(rt, n, DupFn{sortedHash.p, sortedPos.p, posKey.p, n, dupBits.p, cnt.p, 0, nullptr, nullptr});
    unsigned long long nd = 0; d2h(rt, &nd, cnt.p, sizeof(nd));
    n_dup = (int64_t)nd;
    if (n_dup) {
      DevBuf<uint32_t> tIdx; DevBuf<uint64_t> tLinks;
      tIdx.ensure((size_t)n_dup); tLinks.ensure((size_t)n_dup);
      dev_memset(rt, cnt.p, 0, sizeof(unsigned long long));
      foreach(rt, n, DupFn{sortedHash.p, sortedPos.p, posKey.p, n, dupBits.p, cnt.p, 1, tIdx.p, tLinks.p});
      dupIdx.ensure((size_t)n_dup); dupLinks.ensure((size_t)n_dup);
      pr.sort_pairs<uint32_t, uint64_t>(tIdx.p, dupIdx.p, tLinks.p, dupLinks.p, n_dup);
      rt.sync();
    }
    build_dup_rank();
    rt.sync();
  }
  // dupRB from dupBits (finalize, and after mm_index_load: the file format keeps dupBits / dupIdx / dupLinks only)
  void build_dup_rank() {
    const int64_t nw = n / 32 + 2;
    dupRB.ensure((size_t)nw);
    DevBuf<int32_t> pc; DevBuf<int64_t> rk; pc.ensure((size_t)nw + 1); rk.ensure((size_t)nw + 1);
    foreach(rt, nw, DupPopcFn{dupBits.p, pc.p});
    pr.exclusive_sum<int32_t, int64_t>(pc.p, rk.p, nw);
    foreach(rt, nw, DupRBFn{dupBits.p, rk.p, dupRB.p});
    rt.sync();
  }

  int64_t device_bytes() const {
    return (int64_t)(miHash.bytes() + miWs.bytes() + table.bytes() + posKey.bytes() + posSeq16.bytes() + dupBits.bytes() + dupRB.bytes() + dupIdx.bytes() +
                     dupLinks.bytes() + contigStart.bytes() + contigLen.bytes() + uHash.bytes() + uCnt.bytes());
  }
};

}  // namespace mm
