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
#ifdef MM_HOST_EMU
static long long g_emu_rebuild_calls = 0, g_emu_rebuild_elems = 0;
static int g_emu_prune_shift[2] = {0, 0};   // test hook: moves the prune pass's upper / lower ranks off their estimates (exactness must not depend on them)
static long long g_emu_sweep_iters = 0;      // test hook: events applied by the banded sweep (how much the window skipping saves)
#endif

// event code of one reference minimizer of a span: bit 31 = its hash is in the read sketch (idx = 1-based query rank),
// else idx = gap (number of query hashes below it); bits 30/29 = its insertion / deletion does not change the window's
// hash SET because another copy of the same hash is inside the window at that moment (SlideMapper keeps one entry per hash)
static const uint32_t CODE_MATCH = 0x80000000u, CODE_INS_NOP = 0x40000000u, CODE_DEL_NOP = 0x20000000u, CODE_DUP = 0x10000000u, CODE_IDX = 0x0FFFFFFFu;
// (CODE_DUP: the minimizer shares its hash with another one of the same contig; the banded sweep then consults the dup links
//  when it has to rebuild its state from a window)
// K5b state of one candidate: s+1 gap counters (cntBytes each) + s match bits + one spare word, in 32-bit words
MM_HD int32_t sweep_cnt_words(int32_t s, int32_t cntBytes) { return ((s + 1) * cntBytes + 3) / 4; }
MM_HD int32_t sweep_state_words(int32_t s, int32_t cntBytes) { return sweep_cnt_words(s, cntBytes) + (s + 31) / 32 + 1; }

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
  const uint64_t* key; const uint32_t* ws; const int32_t* head; const int64_t* idx; uint32_t* qHash; uint8_t* qStrand; int32_t* qRead;
  MM_HD void operator()(int64_t e) const {
    if (ldg(head + e)) {
      int64_t d = ldg(idx + e); uint64_t k = ldg(key + e);
      qHash[d] = (uint32_t)k; qStrand[d] = (uint8_t)(ldg(ws + e) & 1u); qRead[d] = (int32_t)(k >> 32);
    }
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

#ifndef MM_HOST_EMU
// Device fast path of K3: one CTA per read.  The read's minimizers (hash, wpos|strand) are radix-sorted by hash in shared
// memory (stable, like the global sort it replaces), equal-hash runs are reduced to their first element (std::unique,
// computeMap.hpp:295) and written, compacted, to the read's own slice of tHash/tStrand; reads with a hash that occurs
// on both strands are listed for the std::sort replay (AmbigResolveFn).  Three instantiations cover reads of up to
// 1024 / 2048 / 6144 minimizers; a longer read sends the whole batch through the global-sort path.
template <int ITEMS>
struct K3Block {
  typedef cub::BlockLoad<uint32_t, 256, ITEMS, cub::BLOCK_LOAD_WARP_TRANSPOSE> Load;
  typedef cub::BlockRadixSort<uint32_t, 256, ITEMS, uint32_t> Sort;
  union Temp { typename Load::TempStorage load; typename Sort::TempStorage sort; };
};
template <int ITEMS>
__global__ void __launch_bounds__(256) read_sketch_block_kernel(const uint32_t* hash, const uint32_t* ws, const int64_t* seqOff, int32_t n_reads, int32_t nLo,
                                                                int32_t nHi, uint32_t* tHash, uint8_t* tStrand, int32_t* sOf, unsigned long long* ambCount,
                                                                int32_t* ambList, int64_t ambCap, unsigned long long* readCursor) {
  typedef K3Block<ITEMS> B;
  typedef cub::BlockScan<int32_t, 256> Scan;
  extern __shared__ __align__(16) unsigned char dyn[];
  typename B::Temp& tmp = *reinterpret_cast<typename B::Temp*>(dyn);
  __shared__ typename Scan::TempStorage scanTmp;
  __shared__ uint32_t lastKey[256], lastVal[256];
  __shared__ int smR0;
  // reads are handed out eight at a time (readCursor): the reads of one size class are scattered over the batch and a static
  // stride leaves some CTAs with several of the long ones
  for (int32_t rr = 0;; rr++) {
    if ((rr & 7) == 0) {
      __syncthreads();
      if (threadIdx.x == 0) smR0 = (int)atomicAdd(readCursor, 8ull);
      __syncthreads();
    }
    const int32_t r = smR0 + (rr & 7);
    if (smR0 >= n_reads) break;
    if (r >= n_reads) continue;
    const int64_t b = seqOff[r]; const int32_t n = (int32_t)(seqOff[r + 1] - b);
    if (n <= nLo || n > nHi) continue;
    uint32_t keys[ITEMS], vals[ITEMS];
    typename B::Load(tmp.load).Load(hash + b, keys, n, 0xFFFFFFFFu);      // blocked arrangement, emission order; padding sorts last
    __syncthreads();
    typename B::Load(tmp.load).Load(ws + b, vals, n, 0xFFFFFFFFu);
    __syncthreads();
    typename B::Sort(tmp.sort).Sort(keys, vals);                            // stable LSD radix sort by hash
    lastKey[threadIdx.x] = keys[ITEMS - 1]; lastVal[threadIdx.x] = vals[ITEMS - 1];
    __syncthreads();
    uint32_t pk = threadIdx.x > 0 ? lastKey[threadIdx.x - 1] : 0u, pv = threadIdx.x > 0 ? lastVal[threadIdx.x - 1] : 0u;
    const int32_t i0 = (int32_t)threadIdx.x * ITEMS;
    uint32_t headMask = 0; int32_t cnt = 0; int amb = 0;
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
      const int32_t i = i0 + j;
      const bool real = i < n;
      const bool head = real && (i == 0 || keys[j] != pk);
      amb |= (real && !head && ((vals[j] ^ pv) & 1u)) ? 1 : 0;             // same hash, other strand (AmbigDetectFn)
      headMask |= head ? (1u << j) : 0u; cnt += head ? 1 : 0;
      pk = keys[j]; pv = vals[j];
    }
    int32_t base = 0, total = 0;
    Scan(scanTmp).ExclusiveSum(cnt, base, total);
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
      if (headMask & (1u << j)) { tHash[b + base] = keys[j]; tStrand[b + base] = (uint8_t)(vals[j] & 1u); base++; }
    }
    const int anyAmb = __syncthreads_or(amb);
    if (threadIdx.x == 0) {
      sOf[r] = total;
      if (anyAmb) { const unsigned long long slot = atomicAdd(ambCount, 1ull); if ((int64_t)slot < ambCap) ambList[slot] = r; }
    }
    __syncthreads();
  }
}
// reads outside every class: sketch size 0 for empty reads, overflow flag for oversize ones
struct K3EdgeFn {
  const int64_t* seqOff; int32_t nMax; int32_t* sOf; unsigned long long* overflow;
  MM_HD void operator()(int64_t r) const {
    const int64_t n = ldg(seqOff + r + 1) - ldg(seqOff + r);
    if (n == 0) sOf[r] = 0;
    else if (n > nMax) { sOf[r] = 0; atomic_add_u64(overflow, 1ull); }
  }
};
// the reads' compacted sketches -> one dense array (qOff = prefix sum of the sketch sizes)
__global__ void __launch_bounds__(256) read_sketch_gather_kernel(const uint32_t* tHash, const uint8_t* tStrand, const int64_t* seqOff, const int64_t* qOff,
                                                                 const int32_t* sOf, int32_t n_reads, uint32_t* qHash, uint8_t* qStrand, int32_t* qRead) {
  const int lane = threadIdx.x & 31;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_reads; r += nw) {
    const int64_t src = seqOff[r], dst = qOff[r]; const int32_t s = sOf[r];
    for (int32_t i = lane; i < s; i += 32) { qHash[dst + i] = tHash[src + i]; qStrand[dst + i] = tStrand[src + i]; qRead[dst + i] = (int32_t)r; }
  }
}
#endif

struct AmbigResolveBigFn {  // items beyond *count are empty
  AmbigResolveFn f; const unsigned long long* count;
  MM_HD void operator()(int64_t i) const { if ((unsigned long long)i < *count) f(i); }
};
#ifndef MM_HOST_EMU
// Device fast path of the std::sort replay: one CTA per ambiguous read, the read's minimizers staged in shared memory by
// the whole warp, the replay itself run by lane 0 (it is a sequential algorithm; what matters is that its ~10^5 dependent
// loads and stores hit shared memory instead of L2).  Reads beyond the shared-memory budget use AmbigResolveFn.
__global__ void __launch_bounds__(32) ambig_resolve_smem_kernel(AmbigResolveFn f, int32_t maxN, int32_t* big, unsigned long long* bigCount) {
  extern __shared__ __align__(8) uint64_t sa[];
  const int32_t r = f.list[blockIdx.x];
  const int64_t b = f.seqOff[r]; const int64_t n = f.seqOff[r + 1] - b;
  if (n > maxN) { if (threadIdx.x == 0) { const unsigned long long s_ = atomicAdd(bigCount, 1ull); big[s_] = r; } return; }
  for (int64_t j = threadIdx.x; j < n; j += 32) sa[j] = ((uint64_t)__ldg(f.hash + b + j) << 32) | __ldg(f.ws + b + j);
  __syncwarp();
  if (threadIdx.x == 0) {
    stdsort::sort(sa, n);
    int64_t q = f.qOff[r];
    for (int64_t j = 0; j < n; j++) {
      if (j == 0 || (uint32_t)(sa[j] >> 32) != (uint32_t)(sa[j - 1] >> 32)) { f.qStrand[q] = (uint8_t)(sa[j] & 1u); q++; }
    }
  }
}
#endif

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
#ifndef MM_HOST_EMU
// Device fast path of the probe: persistent CTAs; batches of 2048 probe keys are staged into shared memory by the
// TMA engine (cp.async.bulk global -> shared, completion on an mbarrier), double-buffered, so the next batch lands
// while the current one is being probed.  The probes themselves are independent random 16-byte slot reads; each thread
// keeps four of them in flight.
MM_DEV uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
MM_DEV void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
MM_DEV void tma_load_1d(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)), "l"(src), "r"(bytes),
               "r"(smem_addr(bar))
               : "memory");
}
MM_DEV void mbar_wait(unsigned long long* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "MM_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra MM_DONE_%=;\n"
      "bra MM_WAIT_%=;\n"
      "MM_DONE_%=:\n"
      "}\n" ::"r"(smem_addr(bar)),
      "r"(parity)
      : "memory");
}
static const int PROBE_TILE = 2048;
__global__ void __launch_bounds__(256) l1_probe_tma_kernel(const Slot* table, uint32_t mask, const uint32_t* qHash, int32_t freqThreshold, int32_t* hitCnt,
                                                           int64_t* hitStart, int64_t n) {
  __shared__ __align__(128) uint32_t keys[2][PROBE_TILE];
  __shared__ __align__(8) unsigned long long bar[2];
  const int64_t nTiles = (n + PROBE_TILE - 1) / PROBE_TILE;
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1); mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto tileBytes = [&](int64_t tile) { int64_t c = n - tile * PROBE_TILE; if (c > PROBE_TILE) c = PROBE_TILE; return (uint32_t)(((c * 4 + 15) / 16) * 16); };
  int64_t tile = blockIdx.x; int stage = 0; uint32_t phase0 = 0, phase1 = 0;
  if (tile < nTiles && threadIdx.x == 0) tma_load_1d(keys[0], qHash + tile * PROBE_TILE, tileBytes(tile), &bar[0]);
  for (; tile < nTiles; tile += gridDim.x) {
    const int64_t next = tile + gridDim.x;
    if (next < nTiles && threadIdx.x == 0) tma_load_1d(keys[stage ^ 1], qHash + next * PROBE_TILE, tileBytes(next), &bar[stage ^ 1]);
    if (stage == 0) { mbar_wait(&bar[0], phase0); phase0 ^= 1; } else { mbar_wait(&bar[1], phase1); phase1 ^= 1; }
    const int64_t base = tile * PROBE_TILE;
    const int32_t cnt = (int32_t)((n - base) < PROBE_TILE ? (n - base) : PROBE_TILE);
    for (int32_t i0 = threadIdx.x; i0 < cnt; i0 += 4 * blockDim.x) {
      uint32_t h[4]; uint4 v[4]; uint32_t sl[4]; bool act[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int32_t i = i0 + u * (int32_t)blockDim.x;
        act[u] = i < cnt;
        h[u] = act[u] ? keys[stage][i] : 0u;
        sl[u] = slot_of(h[u], mask);
        if (act[u]) v[u] = __ldg(reinterpret_cast<const uint4*>(table + sl[u]));
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        if (!act[u]) continue;
        const int32_t i = i0 + u * (int32_t)blockDim.x;
        uint4 w = v[u]; uint32_t s_ = sl[u];
        while (w.y != 0 && w.x != h[u]) { s_ = (s_ + 1) & mask; w = __ldg(reinterpret_cast<const uint4*>(table + s_)); }
        int32_t c = 0;
        if (w.y != 0 && (int64_t)w.y < (int64_t)freqThreshold) { c = (int32_t)w.y; hitStart[base + i] = (int64_t)(((uint64_t)w.w << 32) | w.z); }
        hitCnt[base + i] = c;
      }
    }
    __syncthreads();          // the whole CTA is done with this stage before it is refilled
    stage ^= 1;
  }
}
#endif

// Hits are written as ONE 64-bit sort key  read << (seqBits+wsBits) | seqId << wsBits | (wpos<<1|strand), so that a
// single radix sort over the used bits orders them per read by (seqId, wpos, strand) -- the order std::sort gives
// the reference (computeMap.hpp:352).
struct HitKeyLayout { int seqBits, wsBits; };
inline HitKeyLayout hit_key_layout(const Index& ix) {
  HitKeyLayout lay; lay.seqBits = 1; lay.wsBits = 2;
  int64_t nc_ = ix.n_contigs > 1 ? ix.n_contigs : 2; while (((int64_t)1 << lay.seqBits) < nc_) lay.seqBits++;
  int64_t ml = 2; for (int32_t l : ix.h_contigLen) if (l > ml) ml = l;
  while (((int64_t)1 << (lay.wsBits - 1)) < ml) lay.wsBits++;
  return lay;
}
struct GatherHitsFn {
  const int32_t* hitCnt; const int64_t* hitStart; const int64_t* hitOff; const int32_t* qRead; const uint64_t* posKey; uint64_t* hits; HitKeyLayout lay;
  MM_HD void operator()(int64_t i) const {
    int32_t c = ldg(hitCnt + i);
    if (!c) return;
    int64_t s = ldg(hitStart + i), d = ldg(hitOff + i);
    uint64_t hi = (uint64_t)ldg(qRead + i) << (lay.seqBits + lay.wsBits);
    for (int32_t j = 0; j < c; j++) {
      uint64_t pk = ldg(posKey + s + j);
      hits[d + j] = hi | ((pk >> 32) << lay.wsBits) | (pk & 0xFFFFFFFFull);
    }
  }
};
struct ReadHitOffFn {
  const int64_t* qOff; const int64_t* hitOff; int64_t* readHitOff;
  MM_HD void operator()(int64_t r) const { readHitOff[r] = ldg(hitOff + ldg(qOff + r)); }
};

#ifndef MM_HOST_EMU
// Device fast path of the gather: one CTA per read.  Pass 1 counts the read's seed hits per contig in shared memory
// (16-bit counters, contig id folded into HF bins); pass 2 keeps only the hits on contigs that collected at least
// minimumHits of them -- fewer can never satisfy computeMap.hpp:358-362, so dropping them changes nothing downstream
// but shrinks the sort by the spurious 32-bit-hash collisions (7000 -> ~900 hits per read on the 12 Gbp DB).
// Survivors are appended as composite sort keys at a global cursor; the radix sort restores read order.
__global__ void __launch_bounds__(256) l1_filter_gather_kernel(const int32_t* hitCnt, const int64_t* hitStart, const int64_t* qOff, const int32_t* sOf,
                                                               const int32_t* minHitsTab, const uint64_t* posKey, HitKeyLayout lay, int32_t n_reads,
                                                               uint32_t binMask, unsigned long long* cursor, uint64_t* hitsOut, int32_t* keptPerRead) {
  extern __shared__ uint32_t bins[];            // (binMask+1)/2 words, two 16-bit counters per word
  __shared__ unsigned int smTotal, smPos; __shared__ unsigned long long smBase;
  for (int32_t r = blockIdx.x; r < n_reads; r += gridDim.x) {
    const int32_t s = sOf[r];
    const int64_t q0 = qOff[r], q1 = qOff[r + 1];
    if (s == 0 || q1 <= q0) { if (threadIdx.x == 0) keptPerRead[r] = 0; continue; }
    int32_t mh = minHitsTab[s]; if (mh < 1) mh = 1;
    for (uint32_t i = threadIdx.x; i < (binMask + 1) / 2; i += blockDim.x) bins[i] = 0;
    if (threadIdx.x == 0) { smTotal = 0; smPos = 0; }
    __syncthreads();
    bool saturated = false;
    for (int64_t q = q0 + threadIdx.x; q < q1; q += blockDim.x) {
      const int32_t c = hitCnt[q]; const int64_t st = hitStart[q];
      for (int32_t j = 0; j < c; j++) {
        const uint32_t b = (uint32_t)(__ldg(posKey + st + j) >> 32) & binMask;
        const uint32_t old = atomicAdd(&bins[b >> 1], (b & 1u) ? 0x10000u : 1u);
        if ((((b & 1u) ? (old >> 16) : (old & 0xFFFFu)) & 0xFFFFu) >= 0xFFF0u) saturated = true;
      }
    }
    const int anySat = __syncthreads_or(saturated ? 1 : 0);      // absurdly deep pile-up: keep everything for this read
    // survivors = sum of the bins that reached minimumHits (no second pass over the position lists)
    unsigned int local = 0;
    if (anySat) {              // counters wrapped: everything is kept, so the total is simply the read's hit count
      for (int64_t q = q0 + threadIdx.x; q < q1; q += blockDim.x) local += (unsigned int)hitCnt[q];
    } else {
      for (uint32_t i = threadIdx.x; i < (binMask + 1) / 2; i += blockDim.x) {
        const uint32_t w = bins[i], lo = w & 0xFFFFu, hi = w >> 16;
        local += (lo >= (uint32_t)mh) ? lo : 0u;
        local += (hi >= (uint32_t)mh) ? hi : 0u;
      }
    }
    if (local) atomicAdd(&smTotal, local);
    __syncthreads();
    if (threadIdx.x == 0) { smBase = atomicAdd(cursor, (unsigned long long)smTotal); keptPerRead[r] = (int32_t)smTotal; }
    __syncthreads();
    const uint64_t hi = (uint64_t)r << (lay.seqBits + lay.wsBits);
    for (int64_t q = q0 + threadIdx.x; q < q1; q += blockDim.x) {
      const int32_t c = hitCnt[q]; const int64_t st = hitStart[q];
      for (int32_t j = 0; j < c; j++) {
        const uint64_t pk = __ldg(posKey + st + j);
        const uint32_t b = (uint32_t)(pk >> 32) & binMask;
        const uint32_t v = (bins[b >> 1] >> ((b & 1u) * 16)) & 0xFFFFu;
        if (anySat || v >= (uint32_t)mh) {
          const unsigned int p = atomicAdd(&smPos, 1u);
          hitsOut[smBase + p] = hi | ((pk >> 32) << lay.wsBits) | (pk & 0xFFFFFFFFull);
        }
      }
    }
    __syncthreads();
  }
}
#endif

#ifndef MM_HOST_EMU
// Same filter, streaming the index's 2-byte contig-id side array (Index::posSeq16) instead of the 8-byte position keys:
// pass 1 reads the contig id of every hit once (a 7-entry list is 14 B, one sector), counts it into the bins and parks it in
// a shared-memory cache at the hit's rank within the read (hitOff, the prefix sum of the per-query hit counts); pass 2
// re-reads the ids from that cache and fetches the 8-byte position key of the survivors only.
__global__ void __launch_bounds__(256) l1_filter_gather16_kernel(const int32_t* hitCnt, const int64_t* hitStart, const int64_t* hitOff, const int64_t* qOff,
                                                                 const int32_t* sOf, const int32_t* minHitsTab, const uint16_t* posSeq16, const uint64_t* posKey,
                                                                 HitKeyLayout lay, int32_t n_reads, uint32_t binMask, unsigned long long* cursor,
                                                                 uint64_t* hitsOut, int32_t* keptPerRead, uint32_t cacheCap) {
  extern __shared__ uint32_t bins[];            // (binMask+1)/2 words of 16-bit counters, then cacheCap 16-bit contig ids
  uint16_t* cache = reinterpret_cast<uint16_t*>(bins + (binMask + 1) / 2);
  __shared__ unsigned int smTotal, smPos; __shared__ unsigned long long smBase;
  for (int32_t r = blockIdx.x; r < n_reads; r += gridDim.x) {
    const int32_t s = sOf[r];
    const int64_t q0 = qOff[r], q1 = qOff[r + 1];
    if (s == 0 || q1 <= q0) { if (threadIdx.x == 0) keptPerRead[r] = 0; continue; }
    int32_t mh = minHitsTab[s]; if (mh < 1) mh = 1;
    const int64_t h0 = hitOff[q0]; const int64_t nHits = hitOff[q1] - h0;
    if (nHits == 0) { if (threadIdx.x == 0) keptPerRead[r] = 0; continue; }
    const bool anySat = nHits >= 0xFFF0;        // a 16-bit bin could wrap: keep everything for this read
    for (uint32_t i = threadIdx.x; i < (binMask + 1) / 2; i += blockDim.x) bins[i] = 0;
    if (threadIdx.x == 0) { smTotal = 0; smPos = 0; }
    __syncthreads();
    // Warp-cooperative walk over the position lists: each warp takes 32 queries at a time (lane i holds query i's list start and
    // length), then the lanes stride over the FLATTENED hits of those 32 lists; a lane finds the list its hit belongs to with a
    // 5-step search over the warp's prefix sums (shuffles).  Every lane does the same number of trips whatever the list
    // lengths, and consecutive lanes read consecutive entries of a list.
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nWarps = blockDim.x >> 5;
    unsigned int local = 0;
    for (int pass = 0; pass < 2; pass++) {
      if (pass == 1) {
        if (!anySat) {
          __syncthreads();
          // survivors = sum of the bins that reached minimumHits
          for (uint32_t i = threadIdx.x; i < (binMask + 1) / 2; i += blockDim.x) {
            const uint32_t w = bins[i], lo = w & 0xFFFFu, hi = w >> 16;
            local += (lo >= (uint32_t)mh) ? lo : 0u;
            local += (hi >= (uint32_t)mh) ? hi : 0u;
          }
        } else if (threadIdx.x == 0) local = (unsigned int)nHits;
        if (local) atomicAdd(&smTotal, local);
        __syncthreads();
        if (threadIdx.x == 0) { smBase = atomicAdd(cursor, (unsigned long long)smTotal); keptPerRead[r] = (int32_t)smTotal; }
        __syncthreads();
        if (smTotal == 0) break;
      } else if (anySat) continue;                         // counters could wrap: no counting pass, everything is kept
      const uint64_t hiKey = (uint64_t)r << (lay.seqBits + lay.wsBits);
      for (int64_t qb = q0 + (int64_t)wid * 32; qb < q1; qb += (int64_t)nWarps * 32) {
        const int64_t q = qb + lane;
        const int32_t c = q < q1 ? hitCnt[q] : 0;
        const int64_t st = c ? hitStart[q] : 0;
        int32_t incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        const int32_t ex = incl - c;
        const int32_t T = __shfl_sync(0xffffffffu, incl, 31);
        const uint32_t rankBase = (uint32_t)(hitOff[qb] - h0);
        // U trips at a time: the U searches first (shuffles only), then the U loads back to back, then the shared-memory updates --
        // with one load per trip the kernel ran at one DRAM round trip per warp and trip (2.2 TB/s of 64-byte bursts)
        constexpr int U = 4;
        for (int32_t t0 = 0; t0 < T; t0 += 32 * U) {
          int64_t idx[U]; uint32_t rank[U]; bool ok[U];
#pragma unroll
          for (int u = 0; u < U; u++) {
            const int32_t t = t0 + 32 * u + lane;
            int32_t j = 0;
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
              const int32_t cand = j + step;
              const int32_t v = __shfl_sync(0xffffffffu, ex, cand & 31);
              if (cand < 32 && v <= t) j = cand;
            }
            const int64_t stj = __shfl_sync(0xffffffffu, st, j);
            const int32_t exj = __shfl_sync(0xffffffffu, ex, j);
            ok[u] = t < T; idx[u] = stj + (t - exj); rank[u] = rankBase + (uint32_t)t;
          }
          if (pass == 0) {
            uint32_t sq[U];
#pragma unroll
            for (int u = 0; u < U; u++) sq[u] = ok[u] ? (uint32_t)__ldg(posSeq16 + idx[u]) : 0u;
#pragma unroll
            for (int u = 0; u < U; u++) if (ok[u]) {
              if (rank[u] < cacheCap) cache[rank[u]] = (uint16_t)sq[u];
              const uint32_t b_ = sq[u] & binMask;
              atomicAdd(&bins[b_ >> 1], (b_ & 1u) ? 0x10000u : 1u);
            }
          } else {
            bool keep[U]; uint64_t pk[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
              keep[u] = ok[u] && anySat;
              if (ok[u] && !anySat) {
                const uint32_t sq = (rank[u] < cacheCap) ? (uint32_t)cache[rank[u]] : (uint32_t)__ldg(posSeq16 + idx[u]);
                const uint32_t b_ = sq & binMask;
                keep[u] = ((bins[b_ >> 1] >> ((b_ & 1u) * 16)) & 0xFFFFu) >= (uint32_t)mh;
              }
            }
#pragma unroll
            for (int u = 0; u < U; u++) pk[u] = keep[u] ? __ldg(posKey + idx[u]) : 0ull;
#pragma unroll
            for (int u = 0; u < U; u++) {
              const unsigned int m = __ballot_sync(0xffffffffu, keep[u]);
              if (m) {                                      // one cursor bump per warp and trip
                unsigned int base = 0;
                if (lane == __ffs(m) - 1) base = atomicAdd(&smPos, (unsigned int)__popc(m));
                base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
                if (keep[u]) hitsOut[smBase + base + __popc(m & ((1u << lane) - 1u))] = hiKey | ((pk[u] >> 32) << lay.wsBits) | (pk[u] & 0xFFFFFFFFull);
              }
            }
          }
        }
      }
    }
    __syncthreads();
  }
}
#endif

#ifndef MM_HOST_EMU
// K4 in ONE kernel: probe + contig filter + survivor gather, one CTA per read.  The read's sorted sketch is walked in chunks of
// PF_CHUNK keys: the chunk is staged into shared memory by the TMA engine (cp.async.bulk + mbarrier, like l1_probe_tma_kernel),
// every thread keeps four random 16-byte slot reads in flight, and the (start, count) of every probe stays in shared memory for
// the counting walk over the position lists (pass 1 of l1_filter_gather16_kernel, same flattened walk) -- so the per-probe
// hitCnt / hitStart / hitOff arrays (20 B per probe written, read twice) and the global prefix sum between the two kernels
// are gone; the rank of a hit within the read comes from per-chunk prefix sums in shared memory.  The 8-byte (start, count)
// pairs are spilled to `probeOut` for pass 2 (written and re-read by the same CTA a few microseconds apart: L2).
// Survivors are appended at a global cursor; if the output buffer is too small the kernel still counts (cursor[0] = survivors
// needed, cursor[2] = 1) and the host re-runs it with a larger one.  cursor[1] += all seed hits (the H of SURVEY 8d).
template <int PF_CHUNK>
__global__ void __launch_bounds__(256) l1_probe_filter_kernel(const Slot* table, uint32_t mask, int32_t freqThreshold, const uint32_t* qHash, const int64_t* qOff,
                                                              const int32_t* sOf, const int32_t* minHitsTab, const uint16_t* posSeq16, const uint64_t* posKey,
                                                              HitKeyLayout lay, int32_t n_reads, uint32_t binMask, unsigned long long* cursor, uint64_t* hitsOut,
                                                              unsigned long long hitsCap, int32_t* keptPerRead, uint32_t cacheCap, uint2* probeOut,
                                                              unsigned long long* segStart) {
  extern __shared__ __align__(128) uint32_t bins[];            // bins | contig-id cache | chunk keys | chunk counts | chunk starts
  const uint32_t binWords = (binMask + 1) / 2;
  uint16_t* cache = reinterpret_cast<uint16_t*>(bins + binWords);
  uint32_t* keys = bins + binWords + cacheCap / 2;
  uint32_t* pCnt = keys + PF_CHUNK + 8;
  uint32_t* pStart = pCnt + PF_CHUNK;
  __shared__ __align__(8) unsigned long long bar;
  __shared__ unsigned int smTotal, smPos, grpBase[PF_CHUNK / 32 + 1]; __shared__ unsigned long long smBase; __shared__ int smSkip;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nWarps = blockDim.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  uint32_t phase = 0;
  __shared__ int smRead;
  for (;;) {                                     // reads are handed out one at a time (cursor[3]): their hit counts differ 30-fold
    __syncthreads();
    if (threadIdx.x == 0) smRead = (int)atomicAdd(cursor + 3, 1ull);
    __syncthreads();
    const int32_t r = smRead;
    if (r >= n_reads) break;
    const int32_t s = sOf[r];
    const int64_t q0 = qOff[r], q1 = qOff[r + 1];
    if (s == 0 || q1 <= q0) { if (threadIdx.x == 0) keptPerRead[r] = 0; continue; }
    int32_t mh = minHitsTab[s]; if (mh < 1) mh = 1;
    for (uint32_t i = threadIdx.x; i < binWords; i += blockDim.x) bins[i] = 0;
    if (threadIdx.x == 0) { smTotal = 0; smPos = 0; smSkip = 0; }
    __syncthreads();
    const uint64_t hiKey = (uint64_t)r << (lay.seqBits + lay.wsBits);
    uint32_t running = 0;                        // hits of the chunks before the current one (block-uniform)
    bool anySat = false;
    // group bases of the chunk in shared memory: hits before each group of 32 probes, and the chunk's total
    auto chunk_bases = [&](int32_t nchunk) -> uint32_t {
      const int32_t nG = (nchunk + 31) >> 5;
      for (int32_t gi = wid; gi < nG; gi += nWarps) {
        const int32_t i = gi * 32 + lane;
        uint32_t v = i < nchunk ? pCnt[i] : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) grpBase[gi + 1] = v;
      }
      __syncthreads();
      if (wid == 0) {                              // 32 groups at most: one warp scans them
        uint32_t v = lane < nG ? grpBase[lane + 1] : 0u, incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
        if (lane < nG) grpBase[lane + 1] = running + incl;
        if (lane == 0) grpBase[0] = running;
      }
      __syncthreads();
      return grpBase[nG] - running;
    };
    // the flattened walk over the position lists of the chunk's probes (see l1_filter_gather16_kernel); pass 0 counts, pass 1 gathers
    auto walk = [&](int pass, int32_t nchunk) {
      const int32_t nG = (nchunk + 31) >> 5;
      for (int32_t gi = wid; gi < nG; gi += nWarps) {
        const int32_t i = gi * 32 + lane;
        const int32_t c = i < nchunk ? (int32_t)pCnt[i] : 0;
        const int64_t st = c ? (int64_t)pStart[i] : 0;
        int32_t incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        const int32_t ex = incl - c;
        const int32_t T = __shfl_sync(0xffffffffu, incl, 31);
        const uint32_t rankBase = grpBase[gi];
        constexpr int U = 4;
        for (int32_t t0 = 0; t0 < T; t0 += 32 * U) {
          int64_t idx[U]; uint32_t rank[U]; bool ok[U];
#pragma unroll
          for (int u = 0; u < U; u++) {
            const int32_t t = t0 + 32 * u + lane;
            int32_t j = 0;
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
              const int32_t cand = j + step;
              const int32_t v = __shfl_sync(0xffffffffu, ex, cand & 31);
              if (cand < 32 && v <= t) j = cand;
            }
            const int64_t stj = __shfl_sync(0xffffffffu, st, j);
            const int32_t exj = __shfl_sync(0xffffffffu, ex, j);
            ok[u] = t < T; idx[u] = stj + (t - exj); rank[u] = rankBase + (uint32_t)t;
          }
          if (pass == 0) {
            uint32_t sq[U];
#pragma unroll
            for (int u = 0; u < U; u++) sq[u] = ok[u] ? (uint32_t)__ldg(posSeq16 + idx[u]) : 0u;
#pragma unroll
            for (int u = 0; u < U; u++) if (ok[u]) {
              if (rank[u] < cacheCap) cache[rank[u]] = (uint16_t)sq[u];
              const uint32_t b_ = sq[u] & binMask;
              atomicAdd(&bins[b_ >> 1], (b_ & 1u) ? 0x10000u : 1u);
            }
          } else {
            bool keep[U]; uint64_t pk[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
              keep[u] = ok[u] && anySat;
              if (ok[u] && !anySat) {
                const uint32_t sq = (rank[u] < cacheCap) ? (uint32_t)cache[rank[u]] : (uint32_t)__ldg(posSeq16 + idx[u]);
                const uint32_t b_ = sq & binMask;
                keep[u] = ((bins[b_ >> 1] >> ((b_ & 1u) * 16)) & 0xFFFFu) >= (uint32_t)mh;
              }
            }
#pragma unroll
            for (int u = 0; u < U; u++) pk[u] = keep[u] ? __ldg(posKey + idx[u]) : 0ull;
#pragma unroll
            for (int u = 0; u < U; u++) {
              const unsigned int m = __ballot_sync(0xffffffffu, keep[u]);
              if (m) {
                unsigned int base = 0;
                if (lane == __ffs(m) - 1) base = atomicAdd(&smPos, (unsigned int)__popc(m));
                base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
                if (keep[u]) hitsOut[smBase + base + __popc(m & ((1u << lane) - 1u))] = hiKey | ((pk[u] >> 32) << lay.wsBits) | (pk[u] & 0xFFFFFFFFull);
              }
            }
          }
        }
      }
    };
    // ---- pass 0, chunk by chunk: stage keys, probe, count
    for (int64_t c0 = q0; c0 < q1; c0 += PF_CHUNK) {
      const int32_t nchunk = (int32_t)((q1 - c0) < PF_CHUNK ? (q1 - c0) : PF_CHUNK);
      const int64_t a0 = c0 & ~(int64_t)3; const int32_t shift = (int32_t)(c0 - a0);
      if (threadIdx.x == 0) tma_load_1d(keys, qHash + a0, (uint32_t)(((shift + nchunk + 3) & ~3) * 4), &bar);
      mbar_wait(&bar, phase); phase ^= 1;
      for (int32_t i0 = threadIdx.x; i0 < nchunk; i0 += 4 * blockDim.x) {
        uint32_t h[4]; uint4 v[4]; uint32_t sl[4]; bool act[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int32_t i = i0 + u * (int32_t)blockDim.x;
          act[u] = i < nchunk;
          h[u] = act[u] ? keys[shift + i] : 0u;
          sl[u] = slot_of(h[u], mask);
          if (act[u]) v[u] = __ldg(reinterpret_cast<const uint4*>(table + sl[u]));
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          if (!act[u]) continue;
          const int32_t i = i0 + u * (int32_t)blockDim.x;
          uint4 w = v[u]; uint32_t s_ = sl[u];
          while (w.y != 0 && w.x != h[u]) { s_ = (s_ + 1) & mask; w = __ldg(reinterpret_cast<const uint4*>(table + s_)); }
          uint32_t c = 0;
          if (w.y != 0 && (int64_t)w.y < (int64_t)freqThreshold) c = w.y;        // computeMap.hpp:314 (flagged counts compare above everything)
          pCnt[i] = c; pStart[i] = w.z;                                            // CSR start < n < 2^32
          probeOut[c0 + i] = make_uint2(w.z, c);
        }
      }
      __syncthreads();
      const uint32_t tot = chunk_bases(nchunk);
      walk(0, nchunk);
      running += tot;
      __syncthreads();                             // the chunk arrays are free again
    }
    const uint32_t nHits = running;
    if (nHits == 0) { if (threadIdx.x == 0) keptPerRead[r] = 0; __syncthreads(); continue; }
    anySat = nHits >= 0xFFF0u;                      // a 16-bit bin may have wrapped: keep everything for this read
    unsigned int local = 0;
    if (!anySat) {
      for (uint32_t i = threadIdx.x; i < binWords; i += blockDim.x) {
        const uint32_t w = bins[i], lo = w & 0xFFFFu, hi = w >> 16;
        local += (lo >= (uint32_t)mh) ? lo : 0u;
        local += (hi >= (uint32_t)mh) ? hi : 0u;
      }
    } else if (threadIdx.x == 0) local = nHits;
    if (local) atomicAdd(&smTotal, local);
    __syncthreads();
    if (threadIdx.x == 0) {
      smBase = atomicAdd(cursor, (unsigned long long)smTotal); keptPerRead[r] = (int32_t)smTotal; segStart[r] = smBase;
      atomicAdd(cursor + 1, (unsigned long long)nHits);
      if (smBase + smTotal > hitsCap) { smSkip = 1; cursor[2] = 1ull; }
    }
    __syncthreads();
    if (smTotal != 0 && !smSkip) {
      // ---- pass 1: the (start, count) pairs come back from probeOut
      running = 0;
      for (int64_t c0 = q0; c0 < q1; c0 += PF_CHUNK) {
        const int32_t nchunk = (int32_t)((q1 - c0) < PF_CHUNK ? (q1 - c0) : PF_CHUNK);
        for (int32_t i = threadIdx.x; i < nchunk; i += blockDim.x) { const uint2 pc = probeOut[c0 + i]; pStart[i] = pc.x; pCnt[i] = pc.y; }
        __syncthreads();
        const uint32_t tot = chunk_bases(nchunk);
        walk(1, nchunk);
        running += tot;
        __syncthreads();
      }
    }
    __syncthreads();
  }
}
#endif

#ifndef MM_HOST_EMU
// The fused kernel leaves every read's survivors in one contiguous segment (at segStart[r], in no particular order, reads in the order
// their CTAs finished).  The order the candidate stage needs -- (read, contig, position) -- is a sort WITHIN each segment plus a move of
// the segment to readHitOff[r] (the prefix sum of the kept counts): a bitonic sort in shared memory by the group that owns the read
// (a warp for up to SEG_SORT_WARP_CAP keys, a CTA for up to SEG_SORT_CTA_CAP; keys are unique, so any sort gives the radix sort's result)
// instead of seven passes of the device-wide radix sort over all 52 key bits.  Reads beyond the CTA capacity raise `overflow`: the
// host then runs the radix sort after all.
static const int SEG_SORT_WARP_CAP = 1024, SEG_SORT_CTA_CAP = 8192;
// Bitonic network with every compare-exchange ascending (the first step of a merge pairs i with i ^ (k - 1), the others with i ^ j): the
// keys beyond n then behave like +infinity at the end of the array, so a pair whose upper index is >= n is skipped and nothing is padded.
template <bool CTA>
__device__ __forceinline__ void seg_sort_one(const uint64_t* in, uint64_t* out, uint64_t* buf, int32_t n, int lane, int width) {
  int32_t P = 2; while (P < n) P <<= 1;
  for (int32_t i = lane; i < n; i += width) buf[i] = in[i];
  if (CTA) __syncthreads(); else __syncwarp();
  int lh = 0;                                                  // log2 of half the merge size
  for (int32_t k = 2; k <= P; k <<= 1, lh++) {
    const int32_t h = k >> 1;
    for (int32_t t = lane; t < (P >> 1); t += width) {       // flip step
      const int32_t base = (t >> lh) << (lh + 1), off = t & (h - 1), i = base + off, x = base + (k - 1 - off);
      if (x < n) { const uint64_t a = buf[i], b = buf[x]; if (a > b) { buf[i] = b; buf[x] = a; } }
    }
    if (CTA) __syncthreads(); else __syncwarp();
    for (int32_t j = h >> 1; j > 0; j >>= 1) {
      for (int32_t t = lane; t < (P >> 1); t += width) {
        const int32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), x = i | j;
        if (x < n) { const uint64_t a = buf[i], b = buf[x]; if (a > b) { buf[i] = b; buf[x] = a; } }
      }
      if (CTA) __syncthreads(); else __syncwarp();
    }
  }
  for (int32_t i = lane; i < n; i += width) out[i] = buf[i];
  if (CTA) __syncthreads(); else __syncwarp();
}
__global__ void __launch_bounds__(256) l1_sort_segments_warp_kernel(const uint64_t* in, const unsigned long long* segStart, const int32_t* kept, const int64_t* readHitOff,
                                                                    int32_t n_reads, uint64_t* out, int32_t* bigList, unsigned long long* bigCount, unsigned long long* readCursor) {
  extern __shared__ __align__(16) uint64_t segbuf[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint64_t* buf = segbuf + (size_t)wid * SEG_SORT_WARP_CAP;
  for (;;) {                                     // reads are handed out eight at a time: their survivor counts differ widely
    unsigned long long r0 = 0;
    if (lane == 0) r0 = atomicAdd(readCursor, 8ull);
    r0 = __shfl_sync(0xffffffffu, r0, 0);
    if (r0 >= (unsigned long long)n_reads) break;
    for (int32_t r = (int32_t)r0; r < n_reads && r < (int32_t)r0 + 8; r++) {
      const int32_t n = kept[r];
      if (n <= 0) continue;
      if (n > SEG_SORT_WARP_CAP) { if (lane == 0) bigList[atomicAdd(bigCount, 1ull)] = r; continue; }
      seg_sort_one<false>(in + segStart[r], out + readHitOff[r], buf, n, lane, 32);
    }
  }
}
__global__ void __launch_bounds__(256) l1_sort_segments_cta_kernel(const uint64_t* in, const unsigned long long* segStart, const int32_t* kept, const int64_t* readHitOff,
                                                                   uint64_t* out, const int32_t* bigList, const unsigned long long* bigCount, int32_t* overflow) {
  extern __shared__ __align__(16) uint64_t segbuf[];
  const unsigned long long nb = *bigCount;
  for (unsigned long long b = blockIdx.x; b < nb; b += gridDim.x) {
    const int32_t r = bigList[b], n = kept[r];
    if (n > SEG_SORT_CTA_CAP) { if (threadIdx.x == 0) *overflow = 1; continue; }
    seg_sort_one<true>(in + segStart[r], out + readHitOff[r], segbuf, n, (int)threadIdx.x, (int)blockDim.x);
  }
}
#endif

// computeL1CandidateRegions (computeMap.hpp:346-386), one item per sorted hit:
//   hit i opens a candidate iff hits i and i+minimumHits-1 lie on the same contig less than a read length apart;
//   consecutive such candidates are merged while prev.end >= start (ends are non-decreasing, so "prev" is simply
//   the previous flagged hit).
struct HitDecode {
  HitKeyLayout lay;
  MM_HD int32_t read(uint64_t k) const { return (int32_t)(k >> (lay.seqBits + lay.wsBits)); }
  MM_HD int32_t seq(uint64_t k) const { return (int32_t)((k >> lay.wsBits) & ((1ull << lay.seqBits) - 1)); }
  MM_HD int32_t wpos(uint64_t k) const { return (int32_t)((k & ((1ull << lay.wsBits) - 1)) >> 1); }
};
struct HitFlagFn {
  const uint64_t* hits; const int64_t* readHitOff; const int32_t* sOf; const int32_t* readLen; const int32_t* minHitsTab; HitDecode dec;
  int32_t* flag; int64_t n;
  MM_HD void operator()(int64_t i) const {
    int32_t f = 0;
    if (i < n) {
      uint64_t ka = ldg(hits + i);
      int32_t r = dec.read(ka);
      int32_t mh = ldg(minHitsTab + ldg(sOf + r)); if (mh < 1) mh = 1;
      int64_t j = i + mh - 1;
      if (j < ldg(readHitOff + r + 1)) {
        uint64_t kb = ldg(hits + j);
        if (dec.seq(ka) == dec.seq(kb) && dec.wpos(kb) - dec.wpos(ka) < ldg(readLen + r)) f = 1;
      }
    }
    flag[i] = f;
  }
};
struct HitCompactFn {
  const int32_t* flag; const int64_t* fidx; int64_t* flagged;
  MM_HD void operator()(int64_t i) const { if (ldg(flag + i)) flagged[ldg(fidx + i)] = i; }
};
MM_HD void hit_candidate(const uint64_t* hits, const int32_t* sOf, const int32_t* readLen, const int32_t* minHitsTab, const HitDecode& dec, int64_t i,
                         int32_t* r, int32_t* seq, int32_t* start, int32_t* end) {
  uint64_t ka = ldg(hits + i);
  *r = dec.read(ka); *seq = dec.seq(ka); *end = dec.wpos(ka);
  int32_t mh = ldg(minHitsTab + ldg(sOf + *r)); if (mh < 1) mh = 1;
  int32_t st = dec.wpos(ldg(hits + i + mh - 1)) - ldg(readLen + *r) + 1;
  *start = st < 0 ? 0 : st;
}
struct LocusHeadFn {
  const uint64_t* hits; const int64_t* flagged; const int32_t* sOf; const int32_t* readLen; const int32_t* minHitsTab; HitDecode dec;
  int32_t* head; int64_t n;
  MM_HD void operator()(int64_t f) const {
    int32_t h = 0;
    if (f < n) {
      int32_t r, sq, st, en; hit_candidate(hits, sOf, readLen, minHitsTab, dec, ldg(flagged + f), &r, &sq, &st, &en);
      h = 1;
      if (f > 0) {
        int32_t r2, sq2, st2, en2; hit_candidate(hits, sOf, readLen, minHitsTab, dec, ldg(flagged + f - 1), &r2, &sq2, &st2, &en2);
        if (r2 == r && sq2 == sq && en2 >= st) h = 0;
      }
    }
    head[f] = h;
  }
};
struct LocusWriteFn {
  const uint64_t* hits; const int64_t* flagged; const int32_t* sOf; const int32_t* readLen; const int32_t* minHitsTab; HitDecode dec;
  const int32_t* head; const int64_t* lidx; int64_t n;
  int32_t* cRead; int32_t* cSeq; int32_t* cStart; int32_t* cEnd; int32_t* candCnt;
  int64_t* cHitLo; int64_t* cHitHi;                            // the candidate's L1 hits are hits[cHitLo, cHitHi) (K5's prune pass estimates matches from their number)
  MM_HD void operator()(int64_t f) const {
    int32_t r, sq, st, en; hit_candidate(hits, sOf, readLen, minHitsTab, dec, ldg(flagged + f), &r, &sq, &st, &en);
    int64_t l = ldg(lidx + f) + (ldg(head + f) ? 0 : -1);      // lidx = exclusive scan of head
    if (ldg(head + f)) { cRead[l] = r; cSeq[l] = sq; cStart[l] = st; cHitLo[l] = ldg(flagged + f); atomic_add(candCnt + r, 1); }
    if (f + 1 == n || ldg(head + f + 1)) {                      // last flagged hit of the group carries the largest end
      cEnd[l] = en;
      int32_t mh = ldg(minHitsTab + ldg(sOf + r)); if (mh < 1) mh = 1;
      cHitHi[l] = ldg(flagged + f) + mh;
    }
  }
};

#ifndef MM_HOST_EMU
// computeL1CandidateRegions for one read per warp, over the read's sorted hits (same rules as HitFlagFn / LocusHeadFn / LocusWriteFn above,
// which the host emulation and MM_L1_CAND=legacy still run): hit i opens a window iff hits i and i + minimumHits - 1 lie on one contig
// less than a read length apart; consecutive windows merge while prev.end >= start.  The warp walks the hits 32 at a time, carrying the
// last flagged hit across chunks, and writes the read's k-th candidate to scratch slot readHitOff[r] + k; CandGatherFn moves them to
// their final places once the per-read counts are scanned.  One pass over the hits instead of four kernels and two scans over 28 M flags.
__global__ void __launch_bounds__(256) l1_candidates_warp_kernel(const uint64_t* hits, const int64_t* readHitOff, const int32_t* sOf, const int32_t* readLen,
                                                                 const int32_t* minHitsTab, HitDecode dec, int32_t n_reads, int32_t* tSeq, int32_t* tStart, int32_t* tEnd,
                                                                 int32_t* tLo, int32_t* tHi, int32_t* candCnt, unsigned long long* readCursor) {
  const int lane = threadIdx.x & 31;
  for (;;) {
    unsigned long long r0 = 0;
    if (lane == 0) r0 = atomicAdd(readCursor, 8ull);
    r0 = __shfl_sync(0xffffffffu, r0, 0);
    if (r0 >= (unsigned long long)n_reads) break;
    for (int32_t r = (int32_t)r0; r < n_reads && r < (int32_t)r0 + 8; r++) {
      const int64_t b = readHitOff[r];
      const int32_t n = (int32_t)(readHitOff[r + 1] - b);
      if (n <= 0) { if (lane == 0) candCnt[r] = 0; continue; }
      const int32_t len = readLen[r];
      int32_t mh = minHitsTab[sOf[r]]; if (mh < 1) mh = 1;
      bool havePrev = false; int32_t prevSeq = 0, prevEnd = 0, nCand = 0;
      for (int32_t i0 = 0; i0 < n; i0 += 32) {
        const int32_t i = i0 + lane, j = i + mh - 1;
        bool flag = false; int32_t seqA = 0, en = 0, st = 0;
        if (j < n) {
          const uint64_t ka = hits[b + i], kb = hits[b + j];
          seqA = dec.seq(ka); en = dec.wpos(ka);
          flag = seqA == dec.seq(kb) && dec.wpos(kb) - en < len;
          st = dec.wpos(kb) - len + 1; if (st < 0) st = 0;
        }
        const unsigned fm = __ballot_sync(0xffffffffu, flag);
        if (fm == 0) continue;
        const unsigned below = fm & ((1u << lane) - 1u);
        const int p = below ? 31 - __clz(below) : 0;
        int32_t pSeq = __shfl_sync(0xffffffffu, seqA, p), pEn = __shfl_sync(0xffffffffu, en, p);
        bool hasP = below != 0;
        if (!hasP) { pSeq = prevSeq; pEn = prevEnd; hasP = havePrev; }
        const bool head = flag && !(hasP && pSeq == seqA && pEn >= st);
        const unsigned hm = __ballot_sync(0xffffffffu, head);
        const int32_t k = nCand + __popc(hm & ((2u << lane) - 1u)) - 1;       // heads up to and including this lane: the group this flagged hit belongs to
        const unsigned above = lane < 31 ? (fm & ~((2u << lane) - 1u)) : 0u;
        const bool lastHere = flag && (above == 0 || ((hm >> (__ffs(above) - 1)) & 1u));
        if (head) { tSeq[b + k] = seqA; tStart[b + k] = st; tLo[b + k] = i; }
        if (lastHere) { tEnd[b + k] = en; tHi[b + k] = i + mh; }          // a later chunk's hits of the same group overwrite this
        const int top = 31 - __clz(fm);
        prevSeq = __shfl_sync(0xffffffffu, seqA, top); prevEnd = __shfl_sync(0xffffffffu, en, top); havePrev = true;
        nCand += __popc(hm);
        __syncwarp();
      }
      if (lane == 0) candCnt[r] = nCand;
    }
  }
}
struct CandGatherFn {       // candidate c -> (read, k-th of the read) -> its scratch slot
  const int64_t* candOff; int32_t n_reads; const int64_t* readHitOff; const int32_t* tSeq; const int32_t* tStart; const int32_t* tEnd; const int32_t* tLo; const int32_t* tHi;
  int32_t* cRead; int32_t* cSeq; int32_t* cStart; int32_t* cEnd; int64_t* cHitLo; int64_t* cHitHi;
  MM_HD void operator()(int64_t c) const {
    const int64_t r = upper_bound_idx(candOff, (int64_t)n_reads + 1, c) - 1;
    const int64_t src = ldg(readHitOff + r) + (c - ldg(candOff + r));
    cRead[c] = (int32_t)r; cSeq[c] = ldg(tSeq + src); cStart[c] = ldg(tStart + src); cEnd[c] = ldg(tEnd + src);
    cHitLo[c] = (int64_t)ldg(tLo + src); cHitHi[c] = (int64_t)ldg(tHi + src);
  }
};
#endif

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
  const int64_t* cHitLo; const int64_t* cHitHi; int32_t* cHits;
  MM_HD void operator()(int64_t c) const {
    if (c >= n) { spanN[c] = 0; return; }
    { const int64_t nh = ldg(cHitHi + c) - ldg(cHitLo + c); cHits[c] = nh < 0 ? 0 : nh > 0x3fffffff ? 0x3fffffff : (int32_t)nh; }
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
    stWords[c] = sweep_state_words(s, cntBytes);
  }
};

MM_HD uint64_t dup_links(const uint2* dupRB, const uint64_t* dupLinks, int64_t n_dup, int64_t j) {
  (void)n_dup;
  const uint2 rb = ldg(dupRB + (j >> 5));                  // {dupBits word, duplicates before it}
  if (!((rb.x >> (j & 31)) & 1u)) return 0ull;
  return ldg(dupLinks + rb.y + popc32(rb.x & ((1u << (j & 31)) - 1u)));
}
// Which of a duplicated minimizer's two events are no-ops.  Element j enters the window at the step where
// sw_pos = wpos[j]-cmw+1 (or at once if j < fe) and leaves at the step where sw_pos = wpos[j+1]; inside one step the
// reference deletes before it inserts (computeMap.hpp:500-505).
MM_HD uint32_t dup_event_flags(const uint32_t* miWs, const uint2* dupRB, const uint64_t* dupLinks, int64_t n_dup, int64_t j, int64_t b0, int64_t fe,
                               int64_t last, int32_t cmw) {
  const uint64_t l = dup_links(dupRB, dupLinks, n_dup, j);
  const uint32_t pd = (uint32_t)(l >> 32), nd = (uint32_t)l;
  uint32_t f = 0;
  const int64_t wj = (int64_t)(ldg(miWs + j) >> 1);
  if (pd) {
    const int64_t p = j - (int64_t)pd;
    if (p >= b0 && (j < fe || (int64_t)(ldg(miWs + p + 1) >> 1) > wj - cmw + 1)) f |= CODE_INS_NOP;      // the earlier copy is still inside
  }
  if (nd) {
    const int64_t q = j + (int64_t)nd;
    if (q < last && (q < fe || (int64_t)(ldg(miWs + q) >> 1) - cmw + 1 < (int64_t)(ldg(miWs + j + 1) >> 1))) f |= CODE_DEL_NOP;   // a later copy already entered
  }
  return f;
}

// ---------------------------------------------------------------------------------------------- K5 prune
// Which window starts can hold the optimum?  computeL2MappedRegions keeps the FIRST window with the maximal shared count and the
// LAST one that ties with it (computeMap.hpp:510-533), so a window start whose windows cannot reach the count T of some window
// the loop does evaluate need not be visited at all.  Bounds from counts alone, with the read sketch q_1 < ... < q_s, F(i) =
// i + #{distinct W-only hashes below q_i} and istar = max{i : F(i) <= s} (see "L2 restated" in DESIGN.md):
//   lower: if i + (W-only ELEMENTS with gap <= i) <= s then istar >= i, and shared >= the window's non-duplicated matches of rank <= i;
//   upper: if i + (non-duplicated W-only elements with gap < i) > s then istar < i, and shared <= the window's matches of rank < i
//          (without the premise: <= all its matches).
// The ranks i come from where istar is expected (prune_thresholds: the L1 hit count of the candidate estimates the matches of its
// best window); exactness does not depend on them -- a premise that fails leaves the trivial bound.  K5a counts the seven
// indicators per group of 32 consecutive span elements while it classifies them; prefix sums over the groups give every window's
// counts up to a group at either end (rounded to the safe side).  The lower bound is taken on the first window of every group of
// window starts (an evaluated window: its end lies inside the span), the upper bound on the union of a group's windows; the
// sweep (K5b) covers the hull of the groups whose upper bound reaches the best lower bound.
static const int PR_GMAX = 512;            // most groups per candidate the prune pass can take (MM_PRUNE_GMAX; default 256 = 8192 span elements:
                                           // pruning the longer spans too visits 7 % fewer window starts but K5b takes 5.1 instead of 4.6 ms, call AK)
static const int PR_NS = 7;                // series: see PruneView
struct PruneThr { int32_t i0, i1, i1b; };  // upper-bound rank; lower-bound ranks (i1 from the estimate, i1b < i1 holds whatever matches)
MM_HD int32_t prune_istar_est(int32_t s, int32_t win, int32_t m) {
  if (m < 0) m = 0;
  if (m > win) m = win;
  const int64_t e = ((int64_t)s * s) / ((int64_t)s + (win - m) + 1);      // F(i) ~ i (1 + (win - m) / s) = s
  return (int32_t)e;
}
MM_HD PruneThr prune_thresholds(int32_t s, int32_t win, int32_t mest) {
  PruneThr t;
  const int32_t e0 = prune_istar_est(s, win, 0), e1 = prune_istar_est(s, win, mest);
  // spread of istar around the estimate: a binomial count of W-only hashes below q_istar, divided by the slope of F; plus what
  // the rounding of the window's ends to groups of 32 elements can move
  const float p = s > 0 ? (float)e1 / (float)s : 0.f, nw = (float)(win - (mest < win ? mest : win));
  const float sd = sqrtf(nw * p * (1.f - p) + 1.f) / (1.f + nw / (float)(s > 0 ? s : 1));
  // measured on config-2-like data (emulation): 2 sd + 24 above / 2 sd - 4 below sweeps 19 % of the window starts; 3 sd + 20 both ways 23 %
  const int32_t m2 = (int32_t)(2.f * sd);
  t.i0 = e1 + m2 + 24; t.i1 = e1 - m2 + 4; t.i1b = e0 - m2 - 12;
#ifdef MM_HOST_EMU
  t.i0 += g_emu_prune_shift[0]; t.i1 += g_emu_prune_shift[1]; t.i1b += g_emu_prune_shift[1];
#endif
  if (t.i1b > t.i1) t.i1b = t.i1;
  return t;
}
// per-group indicator counts, 6 bits each (0..32): word 0 = A | B << 6 | C << 12 | D << 18 | M << 24, word 1 = C2 | D2 << 6 with
//   A  non-duplicated W-only elements with gap < i0        B  matches with rank < i0          M  all matches
//   C  W-only elements with gap <= i1                      D  non-duplicated matches with rank <= i1      (C2, D2: the same for i1b)
// (W-only: not in the read sketch and not in gap s, which SlideMapper never counts; duplicated: CODE_DUP)
MM_HD void prune_count(uint32_t code, int32_t s, const PruneThr& th, uint32_t& w0, uint32_t& w1) {
  const bool isM = (code & CODE_MATCH) != 0, dup = (code & CODE_DUP) != 0;
  const int32_t idx = (int32_t)(code & CODE_IDX);
  const bool v1 = isM || (idx < s && !dup);                  // counts for A / B
  const bool v2 = isM ? !dup : idx < s;                      // counts for C / D and C2 / D2
  const uint32_t f0 = (v1 && idx < th.i0) ? 1u : 0u, f1 = (v2 && idx <= th.i1) ? 1u : 0u, f2 = (v2 && idx <= th.i1b) ? 1u : 0u;
  const int sh = isM ? 6 : 0;
  w0 += (f0 << sh) + (f1 << (12 + sh)) + (isM ? (1u << 24) : 0u);
  w1 += f2 << sh;
}
MM_HD uint32_t prune_series_of(uint32_t w0, uint32_t w1, int q) { return q < 5 ? (w0 >> (6 * q)) & 63u : (w1 >> (6 * (q - 5))) & 63u; }      // q = 0..6: A, B, C, D, M, C2, D2
// Series q = 0..6: A, B, C, D, M, C2, D2.  Acc::rng(q, ga, gb) = sum of series q over groups [ga, gb) (from prefix sums over the groups);
// H[g] = the last group whose first element lies at or below pos[32 g] + cmw, so that the first element E(32 g) at or beyond that
// position (the end of the first window that starts at 32 g) has 32 H <= E <= 32 (H + 1)
struct PrunePrefix16 {       // P[q * ld + g] = sum over groups [0, g)
  const uint16_t* P; int32_t ld;
  MM_HD int32_t rng(int q, int32_t ga, int32_t gb) const { return gb > ga ? (int32_t)P[q * ld + gb] - (int32_t)P[q * ld + ga] : 0; }
};
struct PrunePrefixPair {     // two series per word: P[(q >> 1) * ld + g], series q in the half q & 1
  const uint32_t* P; int32_t ld;
  MM_HD int32_t rng(int q, int32_t ga, int32_t gb) const {
    if (gb <= ga) return 0;
    const int sh = 16 * (q & 1);
    return (int32_t)((P[(q >> 1) * ld + gb] >> sh) & 0xffffu) - (int32_t)((P[(q >> 1) * ld + ga] >> sh) & 0xffffu);
  }
};
template <class Acc>
struct PruneView {
  Acc acc; const uint16_t* H; int32_t nG, s; PruneThr th;
  // lower bound of the shared count of the window [32 g, E(32 g)): it lies inside groups [g, h] and contains groups [g, h)
  MM_HD int32_t lb(int32_t g) const {
    const int32_t h = H[g];
    if (h + 1 >= nG) return 0;                               // E(32 g) may be the end of the span: that window is not evaluated
    if (th.i1 >= 0 && th.i1 <= s && th.i1 + acc.rng(2, g, h + 1) <= s) return acc.rng(3, g, h);
    if (th.i1b >= 0 && th.i1b <= s && th.i1b + acc.rng(5, g, h + 1) <= s) return acc.rng(6, g, h);
    return 0;
  }
  // upper bound over every window that starts in group g: they contain groups [g + 1, H[g]) and lie inside groups [g, H[g + 1] + 1)
  MM_HD int32_t ub(int32_t g) const {
    const int32_t hLo = H[g];
    int32_t hUp = nG;
    if (g + 1 < nG) { hUp = (int32_t)H[g + 1] + 1; if (hUp > nG) hUp = nG; }
    if (th.i0 + acc.rng(0, g + 1, hLo) > s) return acc.rng(1, g, hUp);
    return acc.rng(4, g, hUp);
  }
};
MM_HD int32_t prune_window_group(const uint32_t* pos, int32_t stride, int32_t nG, int32_t g, int32_t cmw) {      // pos[stride * g]: position of group g's first element
  const uint32_t target = pos[stride * g] + (uint32_t)cmw;
  int32_t lo = g, hi = nG - 1;
  while (lo < hi) { const int32_t m = (lo + hi + 1) >> 1; if (pos[stride * m] <= target) lo = m; else hi = m - 1; }
  return lo;
}
#ifdef MM_HOST_EMU
// serial form of what the device's K5a does besides classifying: the window starts [swB0, swB1) of a candidate worth sweeping
struct L2PruneFn {
  const uint2* ev; const int64_t* evOff; int64_t evBase; int64_t cand0;
  const int64_t* beg0; const int64_t* fe; const int32_t* cRead; const int32_t* sOf; const int32_t* readLen; const int32_t* cHits; int k, w;
  int32_t* swB0; int32_t* swB1; int32_t gmax;                // indexed by the candidate's number within the pass
  void operator()(int64_t ci) const {
    const int64_t c = cand0 + ci;
    const int32_t n = (int32_t)(evOff[c + 1] - evOff[c]), nG = (n + 31) >> 5;
    swB0[ci] = 0; swB1[ci] = 0x7fffffff;
    if (nG < 1 || nG > gmax) return;
    const int32_t r = cRead[c], s = sOf[r], cmw = readLen[r] - (w - 1) - (k - 1);
    const PruneThr th = prune_thresholds(s, (int32_t)(fe[c] - beg0[c]), cHits ? cHits[c] : 0);
    const uint2* e = ev + (evOff[c] - evBase);
    const int32_t ld = nG + 1;
    std::vector<uint16_t> P((size_t)PR_NS * ld, 0), H((size_t)nG); std::vector<uint32_t> pos((size_t)nG);
    for (int32_t g = 0; g < nG; g++) {
      uint32_t w0 = 0, w1 = 0;
      for (int32_t j = 32 * g; j < n && j < 32 * g + 32; j++) prune_count(e[j].x, s, th, w0, w1);
      pos[g] = e[32 * g].y >> 1;
      for (int q = 0; q < PR_NS; q++) P[(size_t)q * ld + g + 1] = (uint16_t)(P[(size_t)q * ld + g] + prune_series_of(w0, w1, q));
    }
    for (int32_t g = 0; g < nG; g++) H[g] = (uint16_t)prune_window_group(pos.data(), 1, nG, g, cmw);
    const PruneView<PrunePrefix16> v{PrunePrefix16{P.data(), ld}, H.data(), nG, s, th};
    int32_t T = 0;
    for (int32_t g = 0; g < nG; g++) { const int32_t l = v.lb(g); if (l > T) T = l; }
    if (T <= 0) return;
    int32_t gF = nG, gL = -1;
    for (int32_t g = 0; g < nG; g++) if (v.ub(g) >= T) { if (g < gF) gF = g; if (g > gL) gL = g; }
    if (gL >= gF) { swB0[ci] = 32 * gF; swB1[ci] = 32 * (gL + 1); }
  }
};
#else
// one warp per candidate: prefix sums of the seven series over the groups (two series per word, lane l takes groups [J l, J l + J)),
// the window extents, the best lower bound, the hull of the groups whose upper bound reaches it
static const int PRUNE_WARPS = 8;
struct L2PruneArgs {
  const uint4* grp; const int64_t* evOff; int64_t evBase; int64_t cand0;
  const int64_t* beg0; const int64_t* fe; const int32_t* cRead; const int32_t* sOf; const int32_t* readLen; const int32_t* cHits; int k, w;
  int32_t* swB0; int32_t* swB1;
};
MM_HD size_t prune_warp_words(int32_t gmax) { return (size_t)4 * (gmax + 1) + (size_t)gmax + (size_t)(gmax + 1) / 2; }      // P pairs, positions, H (16-bit)
__global__ void __launch_bounds__(PRUNE_WARPS * 32) l2_prune_warp_kernel(L2PruneArgs a, int64_t nCand, int32_t gmax) {
  extern __shared__ __align__(16) uint32_t prsm[];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int32_t LD = gmax + 1;
  uint32_t* P = prsm + (size_t)wid * prune_warp_words(gmax);
  uint32_t* POS = P + (size_t)4 * LD;
  uint16_t* H = reinterpret_cast<uint16_t*>(POS + gmax);
  for (int64_t ci = (int64_t)blockIdx.x * PRUNE_WARPS + wid; ci < nCand; ci += (int64_t)gridDim.x * PRUNE_WARPS) {
    const int64_t c = a.cand0 + ci;
    const int32_t n = (int32_t)(a.evOff[c + 1] - a.evOff[c]), nG = (n + 31) >> 5;
    int32_t B0 = 0, B1 = 0x7fffffff;
    if (nG >= 1 && nG <= gmax) {
      const int32_t r = a.cRead[c], s = a.sOf[r], cmw = a.readLen[r] - (a.w - 1) - (a.k - 1);
      const PruneThr th = prune_thresholds(s, (int32_t)(a.fe[c] - a.beg0[c]), a.cHits ? a.cHits[c] : 0);
      const uint4* G = a.grp + (((a.evOff[c] - a.evBase) >> 5) + ci);
      const int32_t J = (nG + 31) >> 5, g0 = J * lane;             // lane l owns groups [J l, J l + J)
      auto unpack = [](const uint4& v, uint32_t* o) {
        o[0] = (v.x & 63u) | (((v.x >> 6) & 63u) << 16);           // A, B
        o[1] = ((v.x >> 12) & 63u) | (((v.x >> 18) & 63u) << 16);  // C, D
        o[2] = ((v.x >> 24) & 63u) | ((v.y & 63u) << 16);          // M, C2
        o[3] = (v.y >> 6) & 63u;                                   // D2
      };
      uint32_t tot[4] = {0, 0, 0, 0};
      for (int32_t j = 0; j < J; j++) {
        const int32_t g = g0 + j;
        if (g < nG) {
          const uint4 v = __ldg(G + g); POS[g] = v.z;
          uint32_t o[4]; unpack(v, o);
#pragma unroll
          for (int q = 0; q < 4; q++) tot[q] += o[q];
        }
      }
      uint32_t run[4];
#pragma unroll
      for (int q = 0; q < 4; q++) {
        uint32_t inc = tot[q];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += x; }
        run[q] = inc - tot[q];
        if (lane == 0) P[q * LD] = 0;
      }
      for (int32_t j = 0; j < J; j++) {                            // second pass over the lane's own records (L1)
        const int32_t g = g0 + j;
        if (g < nG) {
          uint32_t o[4]; unpack(__ldg(G + g), o);
#pragma unroll
          for (int q = 0; q < 4; q++) { run[q] += o[q]; P[q * LD + g + 1] = run[q]; }
        }
      }
      __syncwarp();
      for (int32_t g = lane; g < nG; g += 32) H[g] = (uint16_t)prune_window_group(POS, 1, nG, g, cmw);
      __syncwarp();
      const PruneView<PrunePrefixPair> pv{PrunePrefixPair{P, LD}, H, nG, s, th};
      int32_t T = 0;
      for (int32_t g = lane; g < nG; g += 32) { const int32_t l = pv.lb(g); T = l > T ? l : T; }
      T = __reduce_max_sync(0xffffffffu, T);
      if (T > 0) {
        int32_t gF = nG, gL = -1;
        for (int32_t g = lane; g < nG; g += 32) if (pv.ub(g) >= T) { gF = g < gF ? g : gF; gL = g > gL ? g : gL; }
        gF = __reduce_min_sync(0xffffffffu, gF); gL = __reduce_max_sync(0xffffffffu, gL);
        if (gL >= gF) { B0 = 32 * gF; B1 = 32 * (gL + 1); }
      }
      __syncwarp();
    }
    if (lane == 0) { a.swB0[ci] = B0; a.swB1[ci] = B1; }
  }
}
#endif

// phase A: one item per reference minimizer of a candidate span
struct L2ClassifyFn {
  const uint32_t* miHash; const uint32_t* miWs; const uint32_t* dupBits;
  const int64_t* evOff; int64_t cand0, nCand; int64_t evBase;   // candidates [cand0, cand0+nCand), events relative to evBase
  const int64_t* beg0; const int32_t* cRead; const uint32_t* qHash; const int64_t* qOff; const int32_t* sOf;
  uint2* ev;
  const int64_t* fe; const int64_t* le; const int32_t* readLen; const uint2* dupRB; const uint64_t* dupLinks; int64_t n_dup; int k, w;
  // the device kernel also counts the prune pass's indicators: one record per group of 32 span elements, candidate c's groups start
  // at grp_base(c) (null: no pruning); cHits: L1 hits of the candidate (estimate of its best window's matches)
  const int32_t* cHits; uint4* grp; int32_t gmax;          // gmax: groups per candidate the prune pass takes (<= PR_GMAX)
  MM_HD int64_t grp_base(int64_t c) const { return ((ldg(evOff + c) - evBase) >> 5) + (c - cand0); }
  MM_HD void operator()(int64_t t) const {
    int64_t c = cand0 + upper_bound_idx(evOff + cand0, nCand + 1, t + evBase) - 1;
    int64_t j = ldg(beg0 + c) + (t + evBase - ldg(evOff + c));
    int32_t r = ldg(cRead + c); int32_t s = ldg(sOf + r);
    const uint32_t* q = qHash + ldg(qOff + r);
    uint32_t h = ldg(miHash + j);
    int32_t lo = 0, hi = s;
    while (lo < hi) { int32_t m = (lo + hi) >> 1; if (ldg(q + m) < h) lo = m + 1; else hi = m; }
    uint32_t code = (lo < s && ldg(q + lo) == h) ? (CODE_MATCH | (uint32_t)(lo + 1)) : (uint32_t)lo;
    if ((ldg(dupBits + (j >> 5)) >> (j & 31)) & 1u)
      code |= CODE_DUP | dup_event_flags(miWs, dupRB, dupLinks, n_dup, j, ldg(beg0 + c), ldg(fe + c), ldg(le + c), ldg(readLen + r) - (w - 1) - (k - 1));
    ev[t] = make_uint2(code, ldg(miWs + j));
  }
};

#ifndef MM_HOST_EMU
// Device fast path of phase A: each CTA takes a contiguous run of candidates (candidates are ordered by read, so the read
// sketch staged in shared memory is reused by the read's other candidates).  Next to the sketch sits a 2048-bucket index
// on the top 11 hash bits (first sketch rank of every bucket), so the rank search of a reference minimizer is a lookup
// plus a short binary search inside one bucket instead of log2(s) steps.  Same codes as L2ClassifyFn.  2048 buckets (CLS_BUCKET_BITS = 11): 4096 / 8192
// buckets measured 7.3 / 10.1 ms against 6.4 (the per-read fill and the lost residency outweigh the shorter searches).
// PRUNE: the warp's 32 lanes hold one group of 32 span elements: two warp-wide adds of packed 6-bit fields count the prune pass's
// indicators (prune_count); lane 0 writes the group's record {count words 0 and 1, position of its first element, -} for
// l2_prune_warp_kernel.  (Deciding inside this kernel cost four barriers per candidate: 8.0 ms against 4.9 without pruning.)
template <bool PRUNE, int CLS_BUCKET_BITS>
__global__ void __launch_bounds__(128) l2_classify_smem_kernel(L2ClassifyFn a, int32_t perCta) {
  constexpr int CLS_BUCKETS = 1 << CLS_BUCKET_BITS;
  extern __shared__ uint32_t smq[];
  __shared__ uint16_t bstart[CLS_BUCKETS + 2];
  const int lane = threadIdx.x & 31;
  int32_t curRead = -1, s = 0;
  const int64_t ciEnd = ((int64_t)blockIdx.x + 1) * perCta < a.nCand ? ((int64_t)blockIdx.x + 1) * perCta : a.nCand;
  for (int64_t ci = (int64_t)blockIdx.x * perCta; ci < ciEnd; ci++) {
    const int64_t c = a.cand0 + ci;
    const int32_t r = a.cRead[c];
    if (r != curRead) {
      __syncthreads();                                   // everyone is done with the previous sketch
      curRead = r; s = a.sOf[r];
      const uint32_t* q = a.qHash + a.qOff[r];
      for (int32_t i = threadIdx.x; i < s; i += blockDim.x) smq[i] = __ldg(q + i);
      __syncthreads();
      for (int32_t i = threadIdx.x; i < s; i += blockDim.x) {       // bucket b starts at the first rank whose hash is in bucket >= b
        const int32_t bPrev = i > 0 ? (int32_t)(smq[i - 1] >> (32 - CLS_BUCKET_BITS)) : -1;
        const int32_t bCur = (int32_t)(smq[i] >> (32 - CLS_BUCKET_BITS));
        for (int32_t bb = bPrev + 1; bb <= bCur; bb++) bstart[bb] = (uint16_t)i;
      }
      // the buckets above the largest hash start at s: minimizer hashes crowd the bottom of the range, so this run is most of the table --
      // filled by the whole CTA (left to the thread that holds rank s it was ~1800 serial stores behind a barrier at every read change)
      for (int32_t bb = (s > 0 ? (int32_t)(smq[s - 1] >> (32 - CLS_BUCKET_BITS)) : -1) + 1 + (int32_t)threadIdx.x; bb <= CLS_BUCKETS; bb += blockDim.x) bstart[bb] = (uint16_t)s;
      __syncthreads();
    }
    const int64_t b0 = a.beg0[c];
    const int64_t e0 = a.evOff[c] - a.evBase; const int32_t n = (int32_t)(a.evOff[c + 1] - a.evOff[c]);
    const int64_t fe = a.fe[c], le = a.le[c]; const int32_t cmw = a.readLen[r] - (a.w - 1) - (a.k - 1);
    const int32_t nG = (n + 31) >> 5;
    const bool doPrune = PRUNE && nG >= 1 && nG <= a.gmax;         // CTA-uniform
    PruneThr th{0, 0, 0};
    if (doPrune) th = prune_thresholds(s, (int32_t)(fe - b0), a.cHits ? a.cHits[c] : 0);
    const int64_t gb = PRUNE ? a.grp_base(c) : 0;
    auto rank_code = [&](uint32_t h) -> uint32_t {
      const uint32_t bk = h >> (32 - CLS_BUCKET_BITS);
      int32_t lo = bstart[bk], hi = bstart[bk + 1];
      while (lo < hi) { const int32_t m = (lo + hi) >> 1; if (smq[m] < h) lo = m + 1; else hi = m; }
      return (lo < s && smq[lo] == h) ? (CODE_MATCH | (uint32_t)(lo + 1)) : (uint32_t)lo;
    };
    // four elements per thread and trip: the twelve global loads are issued before the first search needs one
    for (int32_t t0 = threadIdx.x; PRUNE ? ((t0 & ~31) < n) : (t0 < n); t0 += 4 * blockDim.x) {      // PRUNE: warp-uniform bound, the ballots below need the whole warp
      uint32_t h[4], wsv[4], db[4]; bool on[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int32_t t = t0 + u * (int32_t)blockDim.x; on[u] = t < n;
        const int64_t j = b0 + (on[u] ? t : 0);
        h[u] = __ldg(a.miHash + j); wsv[u] = __ldg(a.miWs + j); db[u] = (__ldg(a.dupBits + (j >> 5)) >> (j & 31)) & 1u;
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int32_t t = t0 + u * (int32_t)blockDim.x;
        uint32_t code = 0;
        if (on[u]) {
          code = rank_code(h[u]);
          if (db[u]) code |= CODE_DUP | dup_event_flags(a.miWs, a.dupRB, a.dupLinks, a.n_dup, b0 + t, b0, fe, le, cmw);
          a.ev[e0 + t] = make_uint2(code, wsv[u]);
        }
        if (PRUNE && doPrune && (t & ~31) < n) {         // warp-uniform: t & ~31 is the group's first element
          uint32_t w0 = 0, w1 = 0;                         // the lane's own indicators, 6-bit fields; one warp-wide add per word gives the group's counts (<= 32 each)
          if (on[u]) prune_count(code, s, th, w0, w1);
          w0 = __reduce_add_sync(0xffffffffu, w0); w1 = __reduce_add_sync(0xffffffffu, w1);
          if (lane == 0) a.grp[gb + (t >> 5)] = make_uint4(w0, w1, wsv[u] >> 1, 0u);
        }
      }
    }
  }
}
#endif


// phase B: the evaluate-then-advance loop of computeL2MappedRegions (computeMap.hpp:482-533).
// The gap counters cnt[] and the match bits mb[] live wherever the caller puts them:
//   * L2SweepFn<uint16_t|uint32_t>  -- global memory (any sketch size; also what the host-emulation tests run);
//   * l2_sweep_smem_kernel          -- shared memory, 8-bit counters (device fast path; a counter that would
//                                      pass 255 sends the candidate back to the global-memory functor).
// Both call l2_sweep_one, so the logic that is checked against the oracle on the CPU is the logic the GPU runs.
struct L2SweepArgs {
  const uint2* ev; const int64_t* evOff; int64_t evBase; int64_t cand0;
  const int64_t* beg0; const int64_t* fe; const int64_t* le; const int32_t* cRead; const int32_t* sOf; const int32_t* readLen;
  const uint2* dupRB; const uint64_t* dupLinks; int64_t n_dup; int k, w;
  int32_t* oShared; int32_t* oPos; int32_t* oValid; int64_t* oOptS; int64_t* oOptE; int32_t* oIstar;
};

// Branch-free updates: every lane of a warp runs the same instruction stream whatever the event is.  cnt[] has
// s+1 counters; slot s ("above the largest query hash", never consulted) doubles as the dummy target of events
// that do not move a counter.  mb[] has one spare word after the s match bits for the same purpose.
template <class CntT, bool CHECK_OVF>
struct SweepState {
  CntT* cnt; uint32_t* mb; int32_t s, istar, C, shared, junkBit; bool ovf;
  MM_HD uint32_t bit(int32_t i0) const { return (mb[i0 >> 5] >> (i0 & 31)) & 1u; }      // i0 = 0-based query index
  MM_HD void ins(uint32_t code) {
    const int32_t isM = (int32_t)(code >> 31), idx = (int32_t)(code & CODE_IDX);
    const bool gv = !isM && idx < s;                       // a gap counter really moves
    const int32_t gi = gv ? idx : s;
    const CntT v = cnt[gi];
    if (CHECK_OVF) ovf = ovf || (gv && v == (CntT)~(CntT)0);
    cnt[gi] = (CntT)(v + 1);
    const int32_t below = (gv && idx < istar) ? 1 : 0;
    C += below;
    const int32_t im1 = istar > 0 ? istar - 1 : 0;
    const int32_t cprev = (int32_t)cnt[im1];               // after the increment above (g may be istar-1)
    const int32_t mprev = (int32_t)bit(im1);
    const int32_t dec = (below && istar + C > s) ? 1 : 0;  // q_istar drops out of the bottom-s
    C -= dec ? cprev : 0;
    shared -= dec & mprev;
    istar -= dec;
    const int32_t mi = isM ? idx - 1 : junkBit;
    mb[mi >> 5] |= (uint32_t)isM << (mi & 31);
    shared += (isM && idx <= istar) ? 1 : 0;
  }
  MM_HD void del(uint32_t code) {
    const int32_t isM = (int32_t)(code >> 31), idx = (int32_t)(code & CODE_IDX);
    const bool gv = !isM && idx < s;
    const int32_t gi = gv ? idx : s;
    cnt[gi] = (CntT)(cnt[gi] - 1);
    C -= (gv && idx < istar) ? 1 : 0;
    const int32_t ccur = (int32_t)cnt[istar];              // istar <= s: always a valid slot
    const int32_t adv = (gv && istar < s && istar + 1 + C + ccur <= s) ? 1 : 0;   // q_{istar+1} enters
    C += adv ? ccur : 0;
    istar += adv;
    const int32_t im1 = istar > 0 ? istar - 1 : 0;
    shared += adv & (int32_t)bit(im1);
    const int32_t mi = isM ? idx - 1 : junkBit;
    mb[mi >> 5] &= ~((uint32_t)isM << (mi & 31));
    shared -= (isM && idx <= istar) ? 1 : 0;
  }
};

// one candidate; cnt/mb must be zero on entry.  Returns false if an 8-bit counter overflowed (CHECK_OVF only).
template <class CntT, bool CHECK_OVF>
MM_HD bool l2_sweep_one(const L2SweepArgs& a, int64_t c, CntT* cnt, uint32_t* mb) {
  const int32_t r = ldg(a.cRead + c), s = ldg(a.sOf + r), len = ldg(a.readLen + r);
  const int64_t b0 = ldg(a.beg0 + c);
  const uint2* e = a.ev + (ldg(a.evOff + c) - a.evBase);             // e[j]: index position b0 + j
  SweepState<CntT, CHECK_OVF> z; z.cnt = cnt; z.mb = mb; z.s = s; z.istar = s; z.C = 0; z.shared = 0; z.ovf = false;
  z.junkBit = 32 * ((s + 31) / 32);
  const uint32_t NOP = (uint32_t)s;                                  // "gap s": touches only the dummy slots
  int32_t beg = 0, end = (int32_t)(ldg(a.fe + c) - b0); const int32_t last = (int32_t)(ldg(a.le + c) - b0);
  const int32_t cmw = len - (a.w - 1) - (a.k - 1);
  // slidemap.insert_ref(sw_beg, sw_end) (computeMap.hpp:488); a hash already present is only revised
  for (int32_t j = 0; j < end; j++) {
    const uint32_t code = ldg(&e[j].x);
    z.ins((code & CODE_INS_NOP) ? NOP : (code & (CODE_MATCH | CODE_IDX)));
  }
  int32_t best = 0, bpos = 0, lpos = 0, valid = 0, bistar = s, optS = 0, optE = 0;
  // the two event streams are kept two elements ahead in registers (the loads have two iterations to land)
  uint2 evBeg = ldg(e + beg), evBeg1 = ldg(e + beg + 1), evBeg2 = ldg(e + beg + 2);
  uint2 evEnd = ldg(e + end), evEnd1 = ldg(e + end + 1), evEnd2 = ldg(e + end + 2);
  int32_t sw_pos = (int32_t)(evBeg.y >> 1);
  uint32_t delCode = NOP, insCode = NOP;
  while (end < last) {
    z.del((delCode & CODE_DEL_NOP) ? NOP : (delCode & (CODE_MATCH | CODE_IDX)));      // delete_ref(prev_beg) (slidingMap.hpp:170-219)
    z.ins((insCode & CODE_INS_NOP) ? NOP : (insCode & (CODE_MATCH | CODE_IDX)));      // insert_ref(prev_end) (slidingMap.hpp:139-164)
    if (CHECK_OVF && z.ovf) return false;
    const int32_t wb = (int32_t)(evBeg.y >> 1);
    const bool better = z.shared > best;
    lpos = (z.shared >= best) ? wb : lpos;
    if (better) { best = z.shared; optS = beg; optE = end; bpos = wb; valid = 1; bistar = z.istar; }
    const int32_t nb = (int32_t)(evBeg1.y >> 1) - sw_pos;            // MIIteratorL2::next (MIIteratorL2.hpp:74-96)
    const int32_t ne = (int32_t)(evEnd.y >> 1) - (sw_pos + cmw - 1);
    const int32_t adv = nb < ne ? nb : ne;
    sw_pos += adv;
    delCode = NOP; insCode = NOP;
    if (adv == nb) { delCode = evBeg.x; evBeg = evBeg1; evBeg1 = evBeg2; beg++; evBeg2 = ldg(e + beg + 2); }
    if (adv == ne) { insCode = evEnd.x; evEnd = evEnd1; evEnd1 = evEnd2; end++; evEnd2 = ldg(e + end + 2); }
  }
  a.oShared[c] = best; a.oPos[c] = (bpos + lpos) / 2; a.oValid[c] = valid; a.oOptS[c] = b0 + optS; a.oOptE[c] = b0 + optE; a.oIstar[c] = bistar;
  return true;
}

// global-memory state; item ci -> candidate cand0 + (list ? list[ci] : ci)
template <class CntT>
struct L2SweepFn {
  L2SweepArgs a; uint32_t* state; const int64_t* stOff; int64_t stBase; const int32_t* list;
  MM_HD void operator()(int64_t ci) const {
    int64_t c = a.cand0 + (list ? (int64_t)ldg(list + ci) : ci);
    int32_t s = ldg(a.sOf + ldg(a.cRead + c));
    uint32_t* stp = state + (ldg(stOff + c) - stBase);
    l2_sweep_one<CntT, false>(a, c, (CntT*)stp, stp + sweep_cnt_words(s, (int32_t)sizeof(CntT)));
  }
};
struct SweepKeyFn {     // sort key: descending span length (spanN given) or sketch size: similar work inside a warp, longest first
  const int32_t* cRead; const int32_t* sOf; const int32_t* spanN; int64_t cand0; uint32_t* key; uint32_t* val;
  MM_HD void operator()(int64_t ci) const {
    const uint32_t v = spanN ? (uint32_t)ldg(spanN + cand0 + ci) : (uint32_t)ldg(sOf + ldg(cRead + cand0 + ci));
    key[ci] = 0xFFFFFFFFu - v; val[ci] = (uint32_t)ci;
  }
};

#ifndef MM_HOST_EMU
// Shared-memory fast path.  Each WARP owns a fixed slice of the CTA's shared memory and pulls "tiles" from a global
// counter: a tile = up to 32 consecutive candidates (in descending-sketch-size order) whose gap counters + match bits
// fit the slice; lane t sweeps candidate order[tileStart+t].  No block-level barrier: a slow candidate only delays its
// own warp.  Per-candidate word counts are odd, so the lanes' regions start on different banks.
static const int SWEEP_WARPS_MAX = 8;                // warps per CTA (1 CTA per SM); MM_SWEEP_WARPS overrides
static const int SWEEP_SMEM_WORDS = 56 * 1024;       // 224 KB of the SM's 227 KB, split evenly between the warps
__global__ void __launch_bounds__(SWEEP_WARPS_MAX * 32) l2_sweep_smem_kernel(L2SweepArgs a, const uint32_t* order, const int32_t* tileStart, int32_t nTiles,
                                                                         const int32_t* localOff, unsigned int* tileCounter, int32_t* redo,
                                                                         unsigned long long* redoCount, int32_t sliceWords) {
  extern __shared__ uint32_t sm[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t* slice = sm + wid * sliceWords;
  for (;;) {
    int32_t tile = 0;
    if (lane == 0) tile = (int32_t)atomicAdd(tileCounter, 1u);
    tile = __shfl_sync(0xffffffffu, tile, 0);
    if (tile >= nTiles) break;
    const int32_t t0 = tileStart[tile], t1 = tileStart[tile + 1];
    const int32_t i = t0 + lane;
    if (i < t1) {
      const int64_t c = a.cand0 + (int64_t)order[i];
      const int32_t s = a.sOf[a.cRead[c]];
      uint32_t* my = slice + localOff[i];
      const int32_t cw = sweep_cnt_words(s, 1), words = sweep_state_words(s, 1);
      for (int32_t j = 0; j < words; j++) my[j] = 0;
      const bool ok = l2_sweep_one<uint8_t, true>(a, c, (uint8_t*)my, my + cw);
      if (!ok) { unsigned long long slot = atomicAdd(redoCount, 1ull); redo[slot] = (int32_t)order[i]; }
    }
    __syncwarp();
  }
}
#endif


// ---- banded sweep ---------------------------------------------------------------------------------------------------
// The state of SlideMapper at any moment is a function of the window's CONTENT alone:
//     istar = max{ i : i + #{distinct W-only hashes with gap < i} <= s },   shared = #{ matches with rank <= istar }
// and one event moves istar by at most one.  So the sweep only ever consults the gap counters / match bits next to istar.
// l2_sweep_band keeps them for a BAND of BW consecutive gaps [lo, lo+BW) around istar; W-only events below the band
// only move C (the number of W-only hashes below q_istar), matches below it only move `shared`, events above it are
// ignored.  When istar reaches the edge of the band the state is REBUILT from the window's events around the new istar
// (two simple scans of the ~s events currently in the window).  The first window is built the same way, after a
// 64-bin histogram of the gap indices has located istar.  Per-candidate state: BW + BW/8 + 8 bytes whatever the sketch
// size (the full-state sweep above needs s + s/8), which is what lets 3-4x more candidates be resident per SM.
//
// The loop applies ONE event per iteration (the reference's step "delete prev_beg, insert prev_end, evaluate" becomes
// delete -> [insert] -> evaluate, with the evaluation skipped between the two halves of a step that does both).
// Events come through an EvSrc: DirectEv reads the global array (host emulation, tests); RingEv (device) streams both
// event cursors through per-lane shared-memory rings filled by cp.async, several iterations ahead of their use.
MM_HD int32_t band_bins(int32_t BW) { return BW >= 128 ? 64 : BW / 2; }
struct DirectEv {
  const uint2* e;
  MM_HD void init(int32_t, int32_t) {}
  MM_HD uint2 fetch(int32_t /*stream: 0 = beg cursor, 1 = end cursor*/, int32_t j) { return ldg(e + j); }
};

template <class Ev>
struct BandSweep {
  static constexpr int RB_UNROLL = 16;                      // event codes in flight per lane in rebuild()'s scan (8: 10 % of K5b's stall samples sat on these loads)
  const L2SweepArgs& a; const uint2* e; Ev& ev;
  uint8_t* cnt; uint32_t* mb; int32_t BW;                  // cnt[BW+1], mb[BW/32+1]: the last entries are write-only dummies
  int64_t b0; int32_t s, sh, lo, istar, C, shared; bool fail, bad;

  MM_HD uint32_t bit(int32_t rel) const { return (mb[rel >> 5] >> (rel & 31)) & 1u; }
  // element j of the span counts for a window starting at `beg` iff no earlier copy of its hash is inside that window
  MM_HD bool first_copy(uint32_t code, int32_t j, int32_t beg) const {
    if (!(code & CODE_DUP)) return true;
    const uint64_t l = dup_links(a.dupRB, a.dupLinks, a.n_dup, b0 + j);
    const int64_t pd = (int64_t)(l >> 32);
    return !(pd && (int64_t)j - pd >= (int64_t)beg);
  }
  // State of the window [beg, end) with the band placed around `center` (center < 0: locate istar first with a coarse
  // histogram of the gap indices).  Sets `bad` when istar turns out to lie outside the band (only possible when the
  // caller guessed the center).  Event codes are fetched eight at a time so that the loads overlap.
  MM_HD void rebuild(int32_t beg, int32_t end, int32_t center, int32_t bias) {
#ifdef MM_HOST_EMU
    g_emu_rebuild_calls++; g_emu_rebuild_elems += (center < 0 ? 2 : 1) * (long long)(end - beg);
#endif
    bad = false;
    const uint32_t SKIP = (uint32_t)s;                       // "W-only hash in gap s": ignored everywhere
    if (center < 0) {
      uint16_t* H = reinterpret_cast<uint16_t*>(cnt);       // NB coarse bins of 2^sh gaps (2*NB bytes <= BW)
      const int32_t NB = band_bins(BW);
      for (int32_t b = 0; b < NB; b++) H[b] = 0;
      for (int32_t j0 = beg; j0 < end; j0 += 8) {
        uint32_t cd[8];
#pragma unroll
        for (int u = 0; u < 8; u++) cd[u] = (j0 + u < end) ? ldg(&e[j0 + u].x) : SKIP;
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const uint32_t code = cd[u];
          if ((code & CODE_MATCH) || !first_copy(code, j0 + u, beg)) continue;
          const int32_t g = (int32_t)(code & CODE_IDX);
          if (g < s) H[g >> sh]++;
        }
      }
      int32_t acc = 0, bb = 0;
      for (int32_t b = 0; b < NB; b++) {                     // F(b << sh) = (b << sh) + #{gaps below}: increasing in b
        const int32_t ib = b << sh;
        if (ib > s || ib + acc > s) break;
        bb = b; acc += (int32_t)H[b];
      }
      center = (bb << sh) + (1 << sh) / 2; bias = 0;
    }
    int32_t l0 = center - BW / 2 + bias;
    if (l0 > s + 1 - BW) l0 = s + 1 - BW;
    if (l0 < 0) l0 = 0;
    lo = l0;
    for (int32_t i = 0; i < (BW + 4) / 4; i++) reinterpret_cast<uint32_t*>(cnt)[i] = 0;
    for (int32_t i = 0; i <= BW / 32; i++) mb[i] = 0;
    int32_t Cb = 0, Sb = 0;
    for (int32_t j0 = beg; j0 < end; j0 += RB_UNROLL) {
      uint32_t cd[RB_UNROLL];
#pragma unroll
      for (int u = 0; u < RB_UNROLL; u++) cd[u] = (j0 + u < end) ? ldg(&e[j0 + u].x) : SKIP;
#pragma unroll
      for (int u = 0; u < RB_UNROLL; u++) {
        const uint32_t code = cd[u];
        const bool cntd = first_copy(code, j0 + u, beg);
        const int32_t isM = (int32_t)(code >> 31), idx = (int32_t)(code & CODE_IDX);
        const int32_t rel = idx - isM - lo;                  // match of rank idx -> bit idx-1; W-only hash -> gap idx
        const bool below = cntd && rel < 0;
        Sb += (below && isM) ? 1 : 0; Cb += (below && !isM) ? 1 : 0;
        const bool inb = cntd && (uint32_t)rel < (uint32_t)BW;
        // straight-line: the update that does not apply lands in the dummy slot
        const int32_t cslot = (inb && !isM && idx < s) ? rel : BW;
        const uint32_t v = cnt[cslot];
        if (cslot != BW && v == 255) fail = true;
        cnt[cslot] = (uint8_t)(v + 1);
        const int32_t mslot = (inb && isM) ? rel : BW;
        mb[mslot >> 5] |= 1u << (mslot & 31);
      }
    }
    int32_t i = lo; C = Cb; shared = Sb;
    if (lo > 0 && lo + Cb > s) bad = true;                   // istar is below the band
    while (i < s && i - lo < BW && i + 1 + C + (int32_t)cnt[i - lo] <= s) { C += (int32_t)cnt[i - lo]; i++; shared += (int32_t)bit(i - 1 - lo); }
    if (i < s && i - lo >= BW) bad = true;                   // istar is above the band
    istar = i;
  }
#if defined(__CUDA_ARCH__)
  // Warp-cooperative form of rebuild(beg, end, center, bias) for lane `l` of a converged warp (all 32 lanes call it): the
  // lanes stride over lane l's window (coalesced 8-byte events) and update ITS band with shared-memory atomics; lane l then
  // walks to istar.  30 strides instead of one lane scanning ~s events while 31 lanes wait.
  MM_DEV void rebuild_coop(int l, int32_t beg_, int32_t end_, int32_t center, int32_t bias) {
    const int lane = threadIdx.x & 31;
    int32_t l0 = center - BW / 2 + bias;
    if (l0 > s + 1 - BW) l0 = s + 1 - BW;
    if (l0 < 0) l0 = 0;
    // lane l's view, broadcast
    const int32_t lo_l = __shfl_sync(0xffffffffu, l0, l), s_l = __shfl_sync(0xffffffffu, s, l);
    const int32_t beg_l = __shfl_sync(0xffffffffu, beg_, l), end_l = __shfl_sync(0xffffffffu, end_, l);
    const unsigned long long e_l = __shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)e, l);
    const unsigned long long cnt_l = __shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)cnt, l);
    const unsigned long long mb_l = __shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)mb, l);
    const long long b0_l = __shfl_sync(0xffffffffu, (long long)b0, l);
    const uint2* el = reinterpret_cast<const uint2*>((uintptr_t)e_l);
    uint32_t* cw = reinterpret_cast<uint32_t*>((uintptr_t)cnt_l);
    uint32_t* mw = reinterpret_cast<uint32_t*>((uintptr_t)mb_l);
    for (int32_t i = lane; i < (BW + 4) / 4; i += 32) cw[i] = 0;
    for (int32_t i = lane; i <= BW / 32; i += 32) mw[i] = 0;
    __syncwarp();
    int32_t Cb = 0, Sb = 0, ovf = 0;
    for (int32_t j = beg_l + lane; j < end_l; j += 32) {
      const uint32_t code = __ldg(&el[j].x);
      bool cntd = true;
      if (code & CODE_DUP) {
        const uint64_t lk = dup_links(a.dupRB, a.dupLinks, a.n_dup, b0_l + j);
        const int64_t pd = (int64_t)(lk >> 32);
        cntd = !(pd && (int64_t)j - pd >= (int64_t)beg_l);
      }
      const int32_t isM = (int32_t)(code >> 31), idx = (int32_t)(code & CODE_IDX);
      const int32_t rel = idx - isM - lo_l;
      const bool below = cntd && rel < 0;
      Sb += (below && isM) ? 1 : 0; Cb += (below && !isM) ? 1 : 0;
      const bool inb = cntd && (uint32_t)rel < (uint32_t)BW;
      if (inb && !isM && idx < s_l) {
        const uint32_t old = atomicAdd(cw + (rel >> 2), 1u << (8 * (rel & 3)));
        if (((old >> (8 * (rel & 3))) & 0xffu) == 255u) ovf = 1;
      }
      if (inb && isM) atomicOr(mw + (rel >> 5), 1u << (rel & 31));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { Cb += __shfl_xor_sync(0xffffffffu, Cb, o); Sb += __shfl_xor_sync(0xffffffffu, Sb, o); ovf |= __shfl_xor_sync(0xffffffffu, ovf, o); }
    __syncwarp();
    if (lane == l) {
      bad = false; lo = l0;
      if (ovf) fail = true;
      int32_t i = lo; C = Cb; shared = Sb;
      if (lo > 0 && lo + Cb > s) bad = true;
      while (i < s && i - lo < BW && i + 1 + C + (int32_t)cnt[i - lo] <= s) { C += (int32_t)cnt[i - lo]; i++; shared += (int32_t)bit(i - 1 - lo); }
      if (i < s && i - lo >= BW) bad = true;
      istar = i;
    }
    __syncwarp();
  }
#endif
  // one insertion (isDel = 0) or deletion (isDel = 1) of the reference minimizer with event code `code`
  // (SlideMapper::insert_ref / delete_ref, slidingMap.hpp:139-219).  Straight-line: an update that does not apply
  // lands in the dummy slots.  An insertion can push q_istar out of the bottom-s, a deletion can let q_{istar+1} in;
  // both consult one gap counter and one match bit: rank istar-1 for the insertion, rank istar for the deletion.
  MM_HD void apply(uint32_t code, int32_t isDel) {
    // the slots the move test may consult depend on istar alone: both are loaded up front, off the event's decode chain
    const int32_t t0 = (istar > 0 ? istar - 1 : 0) - lo, t1 = istar - lo;       // insertion consults rank istar-1, deletion rank istar
    const int32_t c0 = (int32_t)cnt[t0], c1 = (int32_t)cnt[t1];
    const uint32_t w0 = mb[t0 >> 5], w1 = mb[t1 >> 5];
    const bool live = (code & (isDel ? CODE_DEL_NOP : CODE_INS_NOP)) == 0;
    const bool isM = (int32_t)code < 0;
    const int32_t idx = (int32_t)(code & CODE_IDX);
    const int32_t d = 1 - 2 * isDel;
    const bool gv = live && !isM && idx < s;                 // a W-only hash in gap idx
    const uint32_t grel = (uint32_t)(idx - lo);
    const bool gin = gv && grel < (uint32_t)BW;
    const int32_t gslot = gin ? (int32_t)grel : BW;
    const int32_t v = (int32_t)cnt[gslot];
    if (gin && !isDel && v == 255) fail = true;
    cnt[gslot] = (uint8_t)(v + d);
    const bool below = gv && idx < istar;
    C += below ? d : 0;
    const int32_t trel = isDel ? t1 : t0;                    // inside the band by the out_of_band() invariant
    const int32_t cX = (isDel ? c1 : c0) + ((gin && gslot == trel) ? d : 0);   // after the update above (the gap may be the consulted one)
    const int32_t mX = (int32_t)(((isDel ? w1 : w0) >> (trel & 31)) & 1u);
    const bool move = isDel ? (gv && istar < s && istar + 1 + C + cX <= s) : (below && istar + C > s);
    C -= move ? d * cX : 0; shared -= move ? d * mX : 0; istar -= move ? d : 0;
    const bool mv = live && isM;                             // a hash of the read sketch, rank idx
    const int32_t mrel = idx - 1 - lo;
    const bool min_ = mv && (uint32_t)mrel < (uint32_t)BW;
    const int32_t mslot = min_ ? mrel : BW;
    const uint32_t mbit = 1u << (mslot & 31);
    const uint32_t word = mb[mslot >> 5];
    mb[mslot >> 5] = isDel ? (word & ~mbit) : (word | mbit);
    shared += (mv && (mrel < 0 || (min_ && idx <= istar))) ? d : 0;
  }
  MM_HD bool out_of_band() const { return (istar <= lo && lo > 0) || (istar >= lo + BW - 1 && lo + BW <= s); }
};

// Result of one SEGMENT of a candidate's sweep.  Because the state is a function of the window alone, a long candidate
// is cut into segments by the window's first element ("beg" in [B0, B1)); every segment rebuilds its first window and
// sweeps on its own, and L2BandMergeFn combines them: the optimum is the FIRST window with the maximal count, the
// reported last position the LAST window with that count (computeMap.hpp:510-533), whichever segments they fall in.
struct BandPart { int32_t shared, bpos, lpos, optS, optE, istar, any, fail; };
// window starts per segment.  With the prune pass most items are a few hundred window starts and the longest single item is the kernel's critical
// path (one lane, ~1000 cycles per event): 256 / 512 / 1024 / 2048 / 4096 = 5.3 / 4.4 / 4.3 / 5.3 / 7.7 ms on config 2 (unpruned: 4096 was best, 12.2 ms)
static const int BAND_SEG_DEFAULT = 1024;

// one segment; cnt: BW+1 (+3 pad) bytes, mb: BW/32+1 words.  part.fail: the candidate must go through the full-state
// sweep instead (sketch too large for the band, a gap counter passing 255, a span beyond 65534 elements).
// On the device the 32 lanes of a warp sweep 32 items in lock step: every lane of the warp must call this function
// (`valid` = the lane has an item), and the loop's continuation test is a warp vote, which re-converges the lanes after
// every iteration (without it, lanes that part ways in one of the data-dependent branches keep looping in separate groups
// and every group pays the full instruction stream).
#if defined(__CUDA_ARCH__)
#define MM_WARP_ANY(x) __any_sync(0xffffffffu, (x))
#else
#define MM_WARP_ANY(x) (x)
#endif
template <class Ev>
MM_HD void l2_sweep_band(const L2SweepArgs& a, bool valid, int64_t c, int32_t B0, int32_t B1, uint8_t* cnt, uint32_t* mb, int32_t BW, Ev& ev, BandPart& out) {
  out.shared = 0; out.bpos = 0; out.lpos = 0; out.optS = 0; out.optE = 0; out.istar = 0; out.any = 0; out.fail = 0;
  if (!valid) c = a.cand0;                                   // harmless loads; the lane never becomes active
  const int32_t r = ldg(a.cRead + c), s = ldg(a.sOf + r), len = ldg(a.readLen + r);
  const int64_t b0 = ldg(a.beg0 + c);
  const uint2* e = a.ev + (ldg(a.evOff + c) - a.evBase);
  const int32_t last = (int32_t)(ldg(a.le + c) - b0);
  const int32_t cmw = len - (a.w - 1) - (a.k - 1);
  out.istar = s;
  int32_t sh = 0; while ((band_bins(BW) << sh) < s + 1) sh++;
  bool fail = valid && (((2 << sh) > BW && s + 1 > BW) || last >= 65535);   // a coarse bin must fit the band with margins
  bool active = valid && !fail;
  // The window whose first element is B0, after every event of that step: all elements below wpos[B0] + cmw are inside
  // (for B0 = 0 this is the reference's first window, computeMap.hpp:465-480).
  int32_t beg = B0, end = last;
  if (active) {
    if (B0 == 0) end = (int32_t)(ldg(a.fe + c) - b0);
    else if (B0 < last) {
      const int32_t lim = (int32_t)(ldg(&e[B0].y) >> 1) + cmw;
      int32_t l = B0, h = last;
      while (l < h) { const int32_t m = (l + h) >> 1; if ((int32_t)(ldg(&e[m].y) >> 1) < lim) l = m + 1; else h = m; }
      end = l;
    }
    active = end < last;                                     // else the reference's loop is over before this window is evaluated
  }
  BandSweep<Ev> z{a, e, ev, cnt, mb, BW, b0, s, sh, 0, 0, 0, 0, false, false};
  if (active) {
    // slidemap.insert_ref(sw_beg, sw_end) (computeMap.hpp:488).  With every window minimizer W-only, F(i) ~ i*(1 + nW/s):
    // try the band around that istar first (one scan); if the guess misses, locate istar with the histogram (two scans).
    z.rebuild(beg, end, (int32_t)(((int64_t)s * s) / (s + (end - beg) + 1)), 0);
    if (z.bad || z.out_of_band()) z.rebuild(beg, end, -1, 0);
    if (z.fail || z.bad) { fail = true; active = false; }
  }
  int32_t best = 0, bpos = 0, lpos = 0, bistar = s, optS = 0, optE = 0, any = 0;
  // The ORDER of the events (which cursor moves next, MIIteratorL2::next, MIIteratorL2.hpp:74-96) depends on the positions alone,
  // not on the band state.  The loop is software-pipelined on that: while event k is applied to the state, event k+1 is
  // already being merged out of the two streams, so the two dependency chains (merge, apply) overlap instead of adding up.
  // merge side (one event ahead): evBeg / evBeg1 / evEnd, sw_pos, mbeg / mend;  p*: the event about to be applied.
  uint2 evBeg = make_uint2(0, 0), evBeg1 = evBeg, evEnd = evBeg;
  if (active) {
    ev.init(beg + 1, end);
    evBeg = ldg(e + beg); evBeg1 = ev.fetch(0, beg + 1); evEnd = ev.fetch(1, end);
  }
  int32_t sw_pos = (int32_t)(evBeg.y >> 1);
  int32_t mbeg = beg, mend = end;
  uint32_t pCode = 0; int32_t pIsDel = 0, pWb = 0; bool pEval = false;
  auto merge_step = [&]() {
    const int32_t nb = (int32_t)(evBeg1.y >> 1) - sw_pos;      // MIIteratorL2::next
    const int32_t ne = (int32_t)(evEnd.y >> 1) - (sw_pos + cmw - 1);
    const int32_t isDel = nb <= ne ? 1 : 0;
    pEval = nb != ne;                                           // a step that deletes AND inserts is evaluated after the insert
    sw_pos += isDel ? nb : ne;
    pCode = isDel ? evBeg.x : evEnd.x; pIsDel = isDel;
    mbeg += isDel; mend += 1 - isDel;
    const uint2 nx = ev.fetch(1 - isDel, isDel ? mbeg + 1 : mend);
    evBeg = isDel ? evBeg1 : evBeg; evBeg1 = isDel ? nx : evBeg1; evEnd = isDel ? evEnd : nx;
    pWb = (int32_t)(evBeg.y >> 1);                              // first element of the window after this event
  };
  if (active) {                                                 // the first window is evaluated before any event (computeMap.hpp:496-510)
    lpos = sw_pos; any = 1;
    if (z.shared > 0) { best = z.shared; bpos = sw_pos; optS = beg; optE = end; bistar = z.istar; }
    merge_step();
  }
#ifdef MM_BAND_DEBUG
  int dbgIters = 0, dbgWin = end - beg, dbgReb = 0;
#endif
  while (MM_WARP_ANY(active)) {
    if (active) {
#ifdef MM_BAND_DEBUG
      dbgIters++;
#endif
      const uint32_t code = pCode; const int32_t isDel = pIsDel, wb = pWb; const bool doEval = pEval;
#ifdef MM_HOST_EMU
      g_emu_sweep_iters++;
#endif
      beg += isDel; end += 1 - isDel;                           // the window this event produces
      const bool more = end < last && beg < B1;
      if (more) merge_step();                                   // event k+1: independent of the band state
      z.apply(code, isDel);
#if !defined(__CUDA_ARCH__)
      if (z.out_of_band() && !z.fail && more) {
#ifdef MM_BAND_DEBUG
        dbgReb += end - beg;
#endif
        z.rebuild(beg, end, z.istar, z.istar <= z.lo ? -BW / 4 : BW / 4);   // keep drifting room on the side it left
        if (z.bad) z.fail = true;
      }
#endif
      active = !z.fail && more;
      if (active && doEval) {
        const bool better = z.shared > best;
        lpos = (z.shared >= best) ? wb : lpos;
        if (better) { best = z.shared; optS = beg; optE = end; bpos = wb; bistar = z.istar; }
      }
    }
#if defined(__CUDA_ARCH__)
    // istar left the band on some lanes: the whole warp rebuilds their states one after the other
    unsigned need = __ballot_sync(0xffffffffu, active && z.out_of_band());
    while (need) {
      const int l = __ffs(need) - 1; need &= need - 1;
      z.rebuild_coop(l, beg, end, z.istar, z.istar <= z.lo ? -BW / 4 : BW / 4);
      if ((int)(threadIdx.x & 31) == l && (z.bad || z.fail)) { z.fail = true; active = false; }
    }
#endif
  }
  fail = fail || z.fail;
  out.shared = best; out.bpos = bpos; out.lpos = lpos; out.optS = optS; out.optE = optE; out.istar = bistar; out.any = any; out.fail = fail ? 1 : 0;
#ifdef MM_BAND_DEBUG
  { ::g_band_dbg_cur[0] = dbgIters; ::g_band_dbg_cur[1] = dbgWin; ::g_band_dbg_cur[2] = dbgReb; }
#endif
}
MM_HD int32_t band_state_words(int32_t BW) { return (BW + 4) / 4 + BW / 32 + 1; }

// work items: a candidate's window start ("beg") runs over [0, span - window) (the loop ends when the window's end reaches the
// end of the span), cut into segments of `seg`; itemOff = prefix sum of the segment counts over the pass
MM_HD int32_t band_beg_range(int32_t span, int32_t win) { const int32_t v = span - win; return v > 0 ? v : 0; }
// the window starts a candidate's items cover: the whole range, or the hull the prune pass left (swB0 / swB1, indexed like nSeg)
MM_HD void band_clip(const int32_t* swB0, const int32_t* swB1, int64_t ci, int32_t range, int32_t& lo, int32_t& hi, bool& open) {
  lo = 0; hi = range; open = true;                           // open: the last segment runs to the end of the span (the window may shrink there)
  if (swB0) {
    const int32_t b0 = ldg(swB0 + ci), b1 = ldg(swB1 + ci);
    if (b0 > 0 && b0 < range) lo = b0;
    if (b1 < range && b1 > lo) { hi = b1; open = false; }
  }
}
struct BandSegCountFn {
  const int32_t* spanN; const int64_t* beg0; const int64_t* fe; int64_t cand0, nCand; int32_t seg; int32_t* nSeg;
  const int32_t* swB0; const int32_t* swB1; unsigned long long* tot;     // tot[0] += window starts to sweep, tot[1] += window starts in all
  MM_HD void operator()(int64_t ci) const {
    if (ci >= nCand) { nSeg[ci] = 0; return; }
    const int64_t c = cand0 + ci;
    const int32_t range = band_beg_range(ldg(spanN + c), (int32_t)(ldg(fe + c) - ldg(beg0 + c)));
    int32_t lo, hi; bool open;
    band_clip(swB0, swB1, ci, range, lo, hi, open);
    const int32_t n = (hi - lo + seg - 1) / seg;
    nSeg[ci] = n < 1 ? 1 : n;
    if (tot) {
#if defined(__CUDA_ARCH__)
      // one pair of atomics per warp (one per candidate made this 30-instruction functor a 0.4 ms kernel)
      const unsigned m = __activemask();
      const unsigned a_ = __reduce_add_sync(m, (unsigned)(hi - lo)), b_ = __reduce_add_sync(m, (unsigned)range);
      if ((int)(threadIdx.x & 31) == __ffs(m) - 1) { atomic_add_u64(tot, (unsigned long long)a_); atomic_add_u64(tot + 1, (unsigned long long)b_); }
#else
      atomic_add_u64(tot, (unsigned long long)(hi - lo)); atomic_add_u64(tot + 1, (unsigned long long)range);
#endif
    }
  }
};
struct BandItemFn {         // item -> candidate, window starts [B0, B1) + sort key (wide-band class first, then descending work)
  const int64_t* itemOff; int64_t nCand; const int32_t* spanN; const int64_t* beg0; const int64_t* fe; int64_t cand0; int32_t seg;
  uint32_t* key; uint32_t* val; int32_t* itemCand; int32_t* itemB0; int32_t* itemB1;
  const int32_t* cRead; const int32_t* sOf; int32_t wideFrom; unsigned long long* nWide;
  const int32_t* swB0; const int32_t* swB1;
  MM_HD void operator()(int64_t i) const {
    const int64_t ci = upper_bound_idx(itemOff, nCand + 1, i) - 1;
    const int32_t sg = (int32_t)(i - ldg(itemOff + ci));
    const int64_t c = cand0 + ci;
    const int32_t win = (int32_t)(ldg(fe + c) - ldg(beg0 + c));
    int32_t lo, hi; bool open;
    band_clip(swB0, swB1, ci, band_beg_range(ldg(spanN + c), win), lo, hi, open);
    const bool lastSeg = i + 1 == ldg(itemOff + ci + 1);
    const int32_t B0 = lo + sg * seg;
    itemCand[i] = (int32_t)ci; itemB0[i] = B0; itemB1[i] = lastSeg ? (open ? 0x7fffffff : hi) : B0 + seg;
    int32_t rest = hi - B0; if (rest > seg) rest = seg; if (rest < 0) rest = 0;
    // ~2 loop iterations per window start (one delete, one insert), and a rebuild scan of the window (cheaper per element)
    // large sketches drift further than the narrow band tolerates: they run in the wide-band instantiation, sorted first
    const bool wide = ldg(sOf + ldg(cRead + c)) >= wideFrom;
    if (wide) atomic_add_u64(nWide, 1ull);
    uint32_t work = (uint32_t)(2 * rest + win / 4); if (work > 0x7FFFFFFFu) work = 0x7FFFFFFFu;
    key[i] = (wide ? 0u : 0x80000000u) | (0x7FFFFFFFu - work); val[i] = (uint32_t)i;
#ifdef MM_BAND_DEBUG
    if (i < 1000000) ::g_band_dbg_key[i] = (int)work;
#endif
  }
};
struct L2BandMergeFn {      // candidate ci: combine its segments (in time order) into the final outputs
  L2SweepArgs a; const BandPart* parts; const int64_t* itemOff; int32_t* redo; unsigned long long* redoCount;
  MM_HD void operator()(int64_t ci) const {
    const int64_t c = a.cand0 + ci;
    const int32_t s = ldg(a.sOf + ldg(a.cRead + c));
    const int64_t b0 = ldg(a.beg0 + c);
    int32_t best = 0, bpos = 0, lpos = 0, valid = 0, bistar = s, optS = 0, optE = 0, fail = 0;
    for (int64_t i = ldg(itemOff + ci); i < ldg(itemOff + ci + 1); i++) {
      const BandPart p = parts[i];
      fail |= p.fail;
      if (!p.any) continue;
      if (p.shared > best) { best = p.shared; bpos = p.bpos; optS = p.optS; optE = p.optE; bistar = p.istar; valid = 1; }
      if (p.shared >= best) lpos = p.lpos;
    }
    if (fail) { unsigned long long slot = atomic_add_u64(redoCount, 1ull); redo[slot] = (int32_t)ci; return; }
    a.oShared[c] = best; a.oPos[c] = (bpos + lpos) / 2; a.oValid[c] = valid; a.oOptS[c] = b0 + optS; a.oOptE[c] = b0 + optE; a.oIstar[c] = bistar;
  }
};

// global-memory band state (host emulation and tests): item i of the pass
struct L2SweepBandFn {
  L2SweepArgs a; uint32_t* state; int32_t BW; const int32_t* itemCand; const int32_t* itemB0; const int32_t* itemB1; BandPart* parts;
  MM_HD void operator()(int64_t i) const {
    const int64_t c = a.cand0 + ldg(itemCand + i);
    uint32_t* stp = state + i * band_state_words(BW);
    DirectEv ev{a.ev + (ldg(a.evOff + c) - a.evBase)};
    BandPart p;
#ifdef MM_BAND_DEBUG
    ::g_band_dbg_cur[0] = ::g_band_dbg_cur[1] = ::g_band_dbg_cur[2] = 0;
#endif
    l2_sweep_band<DirectEv>(a, true, c, ldg(itemB0 + i), ldg(itemB1 + i), (uint8_t*)stp, stp + (BW + 4) / 4, BW, ev, p);
#ifdef MM_BAND_DEBUG
    if (i < 1000000) { ::g_band_dbg_n = i; ::g_band_dbg_it[i] = ::g_band_dbg_cur[0]; ::g_band_dbg_win[i] = ::g_band_dbg_cur[1]; ::g_band_dbg_reb[i] = ::g_band_dbg_cur[2]; }
#endif
    parts[i] = p;
  }
};

#ifndef MM_HOST_EMU
// Device event source: two per-lane rings of R 16-byte pairs (2 events each) in shared memory, laid out [stream][pair][lane].
// A pair is refilled by cp.async the moment the cursor leaves it, i.e. 2R-2 events ahead of its use.
template <int R>
struct RingEv {
  const uint2* gbase;       // candidate's events, rounded down to a 16-byte boundary
  int32_t par;              // 0/1: the candidate's first event is the second half of its pair
  uint32_t sb;              // shared-memory byte address of this lane's slot 0 of stream 0
  MM_DEV RingEv(const uint2* evArray, int64_t off, uint4* ring) : gbase(evArray + (off & ~(int64_t)1)), par((int32_t)(off & 1)), sb((uint32_t)__cvta_generic_to_shared(ring)) {}
  MM_DEV void load_pair(int32_t stream, int32_t pair) {
    const uint32_t dst = sb + (uint32_t)((stream * R + (pair & (R - 1))) * 512);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gbase + 2 * (int64_t)pair));
    asm volatile("cp.async.commit_group;" ::);
  }
  MM_DEV void init(int32_t jb, int32_t je) {
    const int32_t pb = (jb + par) >> 1, pe = (je + par) >> 1;
#pragma unroll
    for (int i = 0; i < R; i++) load_pair(0, pb + i);
#pragma unroll
    for (int i = 0; i < R; i++) load_pair(1, pe + i);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  // the frontier of a stream moved to element j (one past the previous one)
  MM_DEV uint2 fetch(int32_t stream, int32_t j) {
    const int32_t jj = j + par;
    if ((jj & 1) == 0) {                      // entered a new pair: the one behind it is dead, reuse its slot
      load_pair(stream, (jj >> 1) + R - 1);
      asm volatile("cp.async.wait_group %0;" ::"n"(R - 1));
    }
    const uint32_t addr = sb + (uint32_t)((stream * R + ((jj >> 1) & (R - 1))) * 512 + (jj & 1) * 8);
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
  }
};

// Each warp owns WARP_WORDS of shared memory (32 band states + the rings) and pulls tiles of 32 consecutive work items
// of `order` (descending work: lanes of a tile run loops of similar length) from a global counter.
template <int BW, int R, int MINB>
__global__ void __launch_bounds__(MINB == 1 ? 512 : 384, MINB) l2_sweep_band_kernel(L2SweepArgs a, const uint32_t* order, int64_t nItems, const int32_t* itemCand,
                                                               const int32_t* itemB0, const int32_t* itemB1, BandPart* parts, unsigned int* tileCounter) {
  extern __shared__ __align__(16) uint32_t sm[];
  constexpr int ST = ((BW + 4) / 4 + BW / 32 + 1) | 1;       // odd word count: lanes start on different banks
  constexpr int WARP_WORDS = 2 * R * 32 * 4 + 32 * ST;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t* wbase = sm + (size_t)wid * WARP_WORDS;
  uint4* ring = reinterpret_cast<uint4*>(wbase) + lane;
  uint32_t* my = wbase + 2 * R * 32 * 4 + lane * ST;
  const int64_t nTiles = (nItems + 31) / 32;
  for (;;) {
    int64_t tile = 0;
    if (lane == 0) tile = (int64_t)atomicAdd(tileCounter, 1u);
    tile = __shfl_sync(0xffffffffu, tile, 0);
    if (tile >= nTiles) break;
    const int64_t t = tile * 32 + lane;
    const bool valid = t < nItems;
    const int64_t i = valid ? (int64_t)order[t] : 0;
    const int64_t c = a.cand0 + (valid ? itemCand[i] : 0);
    RingEv<R> ev(a.ev, a.evOff[c] - a.evBase, ring);
    BandPart p;
    l2_sweep_band<RingEv<R>>(a, valid, c, valid ? itemB0[i] : 0, valid ? itemB1[i] : 0, (uint8_t*)my, my + (BW + 4) / 4, BW, ev, p);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (valid) parts[i] = p;
    __syncwarp();
  }
}
#endif

// phase C: strand vote of the optimal window (computeMap.hpp:431-438, slidingMap.hpp:232-254)
struct L2StrandFn {
  const uint2* ev; const int64_t* evOff; int64_t evBase; int64_t cand0; const int64_t* beg0;
  const int32_t* cRead; const int64_t* qOff; const uint8_t* qStrand;
  const uint2* dupRB; const uint64_t* dupLinks; int64_t n_dup; const uint32_t* dupBits;
  const int32_t* oValid; const int64_t* oOptS; const int64_t* oOptE; const int32_t* oIstar; int32_t* oVotes;
  // contribution of index position j to the vote of a window [.., b) whose bottom-s holds query ranks <= istar
  MM_HD int32_t vote_of(const uint2* e, const uint8_t* qs, int64_t j, int64_t b, int32_t istar) const {
    uint2 v = e[j];
    if (!(v.x & CODE_MATCH)) return 0;
    int32_t i = (int32_t)(v.x & CODE_IDX);
    if (i > istar) return 0;
    // the map keeps the strand of the LAST inserted occurrence of a hash
    if ((ldg(dupBits + (j >> 5)) >> (j & 31)) & 1u) { uint64_t l = dup_links(dupRB, dupLinks, n_dup, j); uint32_t nd = (uint32_t)l; if (nd && j + (int64_t)nd < b) return 0; }
    int32_t sq = ldg(qs + i - 1) ? 1 : -1, sr = (v.y & 1u) ? 1 : -1;
    return sq * sr;
  }
  MM_HD void operator()(int64_t ci) const {
    int64_t c = cand0 + ci;
    int32_t votes = 0;
    if (ldg(oValid + c)) {
      const uint2* e = ev + (ldg(evOff + c) - evBase) - ldg(beg0 + c);
      const uint8_t* qs = qStrand + ldg(qOff + ldg(cRead + c));
      int64_t a = ldg(oOptS + c), b = ldg(oOptE + c); int32_t istar = ldg(oIstar + c);
      for (int64_t j = a; j < b; j++) votes += vote_of(e, qs, j, b, istar);
    }
    oVotes[c] = votes;
  }
};
#ifndef MM_HOST_EMU
// Device fast path of phase C: one warp per candidate, lanes stride over the optimal window (coalesced 8-byte events).
__global__ void __launch_bounds__(256) l2_strand_warp_kernel(L2StrandFn f, int64_t nc) {
  const int lane = threadIdx.x & 31;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t ci = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; ci < nc; ci += nw) {
    const int64_t c = f.cand0 + ci;
    int32_t votes = 0;
    if (f.oValid[c]) {
      const uint2* e = f.ev + (f.evOff[c] - f.evBase) - f.beg0[c];
      const uint8_t* qs = f.qStrand + f.qOff[f.cRead[c]];
      const int64_t a = f.oOptS[c], b = f.oOptE[c]; const int32_t istar = f.oIstar[c];
      for (int64_t j = a + lane; j < b; j += 32) votes += f.vote_of(e, qs, j, b, istar);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) votes += __shfl_xor_sync(0xffffffffu, votes, o);
    if (lane == 0) f.oVotes[c] = votes;
  }
}
#endif
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

struct MapStats { double ms[16]; int64_t counters[16]; };

struct Mapper {
  Runtime& rt; Prims& pr; Sketcher& sk;
  stats::Tables tabs;
  DevBuf<int32_t> dMinHits, dAccept; int tabUploaded = 0; int tabK = 0; float tabPi = 0;
  // per batch (kept until the next batch for the fetch calls)
  SeqBatch batch; SketchOut rs;          // `batch`: the default input slot (mm_map_batch); run() takes any loaded batch
  int32_t n_reads = 0; int64_t n_q = 0, n_hits = 0, n_cand = 0; int lastK = 16;
  DevBuf<int32_t> readLen, sOf, head, hitCnt, candCnt, cRead, cSeq, cStart, cEnd, spanN, stWords;
  DevBuf<int32_t> oShared, oPos, oValid, oIstar, oVotes, oAccept, readMapped;
  DevBuf<int64_t> qOff, idx, hitStart, hitOff, readHitOff, candOff, beg0, fe, le, evOff, stOff, oOptS, oOptE;
  DevBuf<uint64_t> key, key2, hits, hits2;
  DevBuf<uint32_t> ws2, qHash, state; DevBuf<uint8_t> qStrand, tStrand; DevBuf<uint2> ev;
  DevBuf<int32_t> ambig, ambigList, ambigBig, red, swTile, swLocal, swRedo; DevBuf<unsigned long long> scal, scal2; int64_t n_ambig = 0;
  DevBuf<uint32_t> swKey, swKey2, swVal, swOrder; int32_t maxSketch = 0;
  DevBuf<int32_t> qRead, hflag, lhead, keptPerRead; DevBuf<int64_t> hfidx, flagged, lidx; int64_t n_hits_all = 0;
  std::vector<uint32_t> hk; std::vector<int32_t> tileStartH, localOffH;
  std::vector<int32_t> h_effLen;
  MapStats st;
  int sweepBand = 256, sweepRing = 4, sweepMode = 0, sweepSeg = BAND_SEG_DEFAULT, sweepWideFrom = 0x7fffffff;   // wide-band class: measured neutral on config 2, off unless MM_SWEEP_WIDE_FROM is set
  DevBuf<int64_t> itemOff; DevBuf<int32_t> itemCand, itemB0, itemB1, segCnt, swB0, swB1, cHits; DevBuf<int64_t> cHitLo, cHitHi; DevBuf<BandPart> bandParts; std::vector<int64_t> hEvSpan;
  // K5a also bounds every window's shared count and K5b sweeps only the window starts that can hold the optimum (l2 prune, PruneView);
  // MM_SWEEP_PRUNE=0 sweeps every window start as the reference does
  DevBuf<uint4> grpSum; bool sweepPrune = true, prunedPass = false; int32_t pruneGmax = 256;
  DevBuf<unsigned long long> segStart; DevBuf<int32_t> segBig, candTmp; bool segSorted = false;
  DevBuf<uint2> probeOut;      // (CSR start, count) of every probe of the batch: l1_probe_filter_kernel's spill between its two passes
  int64_t evBudget = (int64_t)1 << 30;       // span elements classified per L2 pass (8 B each: at most 8.6 GB of scratch)

  Mapper(Runtime& r, Prims& p, Sketcher& s) : rt(r), pr(p), sk(s) {
    memset(&st, 0, sizeof(st));
    if (const char* e = getenv("MM_EV_BUDGET")) { long long v = atoll(e); if (v > 0) evBudget = v; }   // tests: force several L2 passes
    // K5b variants (tests and A/B measurements): band width, event-ring depth, and which sweep runs first
    if (const char* e = getenv("MM_SWEEP_BAND")) { int v = atoi(e); if (v == 64 || v == 128 || v == 256) sweepBand = v; }
    if (const char* e = getenv("MM_SWEEP_RING")) { int v = atoi(e); if (v == 2 || v == 4 || v == 8) sweepRing = v; }
    if (const char* e = getenv("MM_SWEEP_WIDE_FROM")) { int v = atoi(e); if (v >= 1) sweepWideFrom = v; }
    if (const char* e = getenv("MM_SWEEP_SEG")) { int v = atoi(e); if (v >= 64) sweepSeg = v; }
    if (const char* e = getenv("MM_SWEEP_PRUNE")) sweepPrune = atoi(e) != 0;
    if (const char* e = getenv("MM_PRUNE_GMAX")) { int v = atoi(e); if (v >= 1 && v <= PR_GMAX) pruneGmax = v; }
    if (const char* e = getenv("MM_SWEEP")) sweepMode = !strcmp(e, "full") ? 1 : !strcmp(e, "global") ? 2 : 0;
  }

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

  // `batch` must already be loaded (Sketcher::load, or prepare + pack_async + finish_pack)
  void run(const Index& ix, SeqBatch& batch, float pi, int32_t minReadLen, int64_t* summary /*6*/) {
    sketch_reads(ix.k, ix.w, batch, minReadLen);
    map_sketched(ix, pi, summary);
  }
  // K0/K1/K3 of a read batch: leaves readLen, sOf, qOff, qHash, qStrand, qRead, n_q, maxSketch (and, pending on the side stream,
  // the std::sort replay of the ambiguous reads).  A contig-sharded rank runs this on ITS block of the reads only; the sketches
  // of all blocks are then all-gathered (mm_map_batch_sharded_dev) before map_sketched.
  int64_t nShort_ = 0, basesOk_ = 0, nExc_ = 0, nMinimizers_ = 0; bool ambigPending_ = false;
  void sketch_reads(int k, int w, SeqBatch& batch, int32_t minReadLen) {
    memset(&st, 0, sizeof(st));
    lastK = k;
    n_reads = batch.n_seqs;
    // reads shorter than w, k or -m are skipped (computeMap.hpp:137): hide them from K1 by zeroing their length
    h_effLen.assign((size_t)n_reads, 0);
    int64_t nShort = 0, basesOk = 0;
    std::vector<int32_t> saveLen = batch.h_len;
    for (int32_t i = 0; i < n_reads; i++) {
      int32_t L = batch.h_len[i];
      if (L < w || L < k || L < minReadLen) { nShort++; batch.h_len[i] = 0; }
      else { h_effLen[i] = L; basesOk += L; }
    }
    readLen.ensure((size_t)n_reads + 1); h2d(rt, readLen.p, h_effLen.data(), sizeof(int32_t) * (size_t)n_reads);
    {
      StageTimer t(rt, &st.ms[0]);
      sk.chunkMs = &st.ms[10];
      sk.run(batch, k, w, rs);
      sk.chunkMs = nullptr;
    }
    batch.h_len = saveLen;
    nShort_ = nShort; basesOk_ = basesOk; nExc_ = batch.n_exc;
    // ---- K3: sort by (read, hash), unique
    int64_t nm = rs.n_total;
    nMinimizers_ = nm;
    sOf.ensure((size_t)n_reads + 1); qOff.ensure((size_t)n_reads + 2);
    bool ambigPending = false;
    {
      StageTimer t(rt, &st.ms[1]);
      bool blockPath = false;
      unsigned long long na = 0; int32_t maxS = 0;
      ambig.ensure((size_t)n_reads + 1); ambigList.ensure(4096); scal.ensure(16);
      key.ensure((size_t)nm + 1);                     // scratch of the std::sort replay
#ifndef MM_HOST_EMU
      if (!getenv("MM_K3_GLOBAL") && nm > 0) {
        // per-read block sort (read_sketch_block_kernel); scal[0] = ambiguous reads, [1] = max sketch, [2] = oversize reads, [3] = n_q
        ws2.ensure((size_t)nm + 1); tStrand.ensure((size_t)nm + 1);
        dev_memset(rt, scal.p, 0, sizeof(unsigned long long) * 12);             // [8..10] = the read cursors of the three size classes
        if (rt.first((const void*)read_sketch_block_kernel<24>))
          MM_CUDA(cudaFuncSetAttribute(read_sketch_block_kernel<24>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K3Block<24>::Temp)));
        int ambCap = (int)ambigList.cap;
        int grid = n_reads < rt.sm_count * 8 ? n_reads : rt.sm_count * 8;
        // the three size classes touch disjoint reads: the two larger ones run on the side stream next to the small one
        // (each alone leaves SMs idle in its tail; MM_K3_SERIAL=1 = one after the other on the main stream)
        const bool fork = !getenv("MM_K3_SERIAL");
        cudaStream_t s2 = fork ? rt.side : rt.stream;
        if (fork) { MM_CUDA(cudaEventRecord(evFork(), rt.stream)); MM_CUDA(cudaStreamWaitEvent(s2, evFork(), 0)); }
        read_sketch_block_kernel<24><<<grid, 256, sizeof(K3Block<24>::Temp), s2>>>(rs.hash.p, rs.ws.p, rs.seqOff.p, n_reads, 2048, 6144, ws2.p, tStrand.p,
                                                                                   sOf.p, scal.p, ambigList.p, ambCap, scal.p + 8);
        read_sketch_block_kernel<8><<<grid, 256, sizeof(K3Block<8>::Temp), s2>>>(rs.hash.p, rs.ws.p, rs.seqOff.p, n_reads, 1024, 2048, ws2.p, tStrand.p, sOf.p,
                                                                                 scal.p, ambigList.p, ambCap, scal.p + 9);
        if (fork) MM_CUDA(cudaEventRecord(evJoin(), s2));
        read_sketch_block_kernel<4><<<grid, 256, sizeof(K3Block<4>::Temp), rt.stream>>>(rs.hash.p, rs.ws.p, rs.seqOff.p, n_reads, 0, 1024, ws2.p, tStrand.p, sOf.p,
                                                                                        scal.p, ambigList.p, ambCap, scal.p + 10);
        if (fork) MM_CUDA(cudaStreamWaitEvent(rt.stream, evJoin(), 0));
        MM_CUDA(cudaGetLastError());
        rt.launches += 3;
        foreach(rt, n_reads, K3EdgeFn{rs.seqOff.p, 6144, sOf.p, scal.p + 2});
        dev_memset(rt, sOf.p + n_reads, 0, sizeof(int32_t));
        pr.exclusive_sum<int32_t, int64_t>(sOf.p, qOff.p, (int64_t)n_reads + 1);
        pr.reduce_max<int32_t>(sOf.p, (int32_t*)(scal.p + 1), n_reads);
        d2d(rt, scal.p + 3, qOff.p + n_reads, sizeof(int64_t));
        unsigned long long hs[4] = {0, 0, 0, 0}; d2h(rt, hs, scal.p, sizeof(hs));
        if (hs[2] == 0 && hs[0] <= (unsigned long long)ambCap) {
          blockPath = true;
          na = hs[0]; maxS = (int32_t)(hs[1] & 0xffffffffu); n_q = (int64_t)hs[3];
          qHash.ensure((size_t)n_q + 1); qStrand.ensure((size_t)n_q + 1); qRead.ensure((size_t)n_q + 1);
          int g2 = (int)(((int64_t)n_reads + 7) / 8); if (g2 > rt.sm_count * 8) g2 = rt.sm_count * 8;
          read_sketch_gather_kernel<<<g2, 256, 0, rt.stream>>>(ws2.p, tStrand.p, rs.seqOff.p, qOff.p, sOf.p, n_reads, qHash.p, qStrand.p, qRead.p);
          MM_CUDA(cudaGetLastError());
          rt.launches++;
        }
      }
#endif
      if (!blockPath) {       // global sort by (read, hash): host emulation, oversize reads, MM_K3_GLOBAL
        key2.ensure((size_t)nm + 1); ws2.ensure((size_t)nm + 1); head.ensure((size_t)nm + 2); idx.ensure((size_t)nm + 2);
        foreach(rt, nm, ReadKeyFn{rs.hash.p, rs.seqOff.p, n_reads, key.p});
        int bits = 33; while (bits < 64 && ((int64_t)1 << (bits - 32)) < n_reads) bits++;
        pr.sort_pairs<uint64_t, uint32_t>(key.p, key2.p, rs.ws.p, ws2.p, nm, bits);
        foreach(rt, nm + 1, HeadFlagFn{key2.p, head.p, nm});
        pr.exclusive_sum<int32_t, int64_t>(head.p, idx.p, nm + 1);
        d2h(rt, &n_q, idx.p + nm, sizeof(int64_t));
        qHash.ensure((size_t)n_q + 1); qStrand.ensure((size_t)n_q + 1); qRead.ensure((size_t)n_q + 1);
        foreach(rt, nm, UniqueScatterFn{key2.p, ws2.p, head.p, idx.p, qHash.p, qStrand.p, qRead.p});
        foreach(rt, (int64_t)n_reads + 1, ReadSketchOffFn{rs.seqOff.p, idx.p, qOff.p, sOf.p, n_reads});
        // duplicate hashes with both strands: settle the survivor like std::sort + std::unique would
        dev_memset(rt, ambig.p, 0, sizeof(int32_t) * ((size_t)n_reads + 1));
        dev_memset(rt, scal.p, 0, sizeof(unsigned long long) * 4);
        foreach(rt, nm, AmbigDetectFn{key2.p, ws2.p, head.p, ambig.p, scal.p, ambigList.p, 4096});
        if (n_reads > 0) pr.reduce_max<int32_t>(sOf.p, (int32_t*)(scal.p + 1), n_reads);
        unsigned long long hs[2] = {0, 0}; d2h(rt, hs, scal.p, sizeof(hs));
        na = hs[0]; maxS = (int32_t)(hs[1] & 0xffffffffu);
        if (na > 4096) {
          ambigList.ensure((size_t)na);
          dev_memset(rt, ambig.p, 0, sizeof(int32_t) * ((size_t)n_reads + 1));
          dev_memset(rt, scal.p, 0, sizeof(unsigned long long));
          foreach(rt, nm, AmbigDetectFn{key2.p, ws2.p, head.p, ambig.p, scal.p, ambigList.p, (int64_t)na});
        }
      }
      {
        n_ambig = (int64_t)na;
        if (na) {
          // The replay is one slow sequential thread per read and only the strand vote (K5c) needs its result:
          // run it on the side stream, rejoin before K5c.
#ifndef MM_HOST_EMU
          MM_CUDA(cudaEventRecord(evFork(), rt.stream));
          MM_CUDA(cudaStreamWaitEvent(rt.side, evFork(), 0));
#endif
          AmbigResolveFn rf{ambigList.p, rs.hash.p, rs.ws.p, rs.seqOff.p, key.p, qOff.p, qHash.p, qStrand.p};
#ifndef MM_HOST_EMU
          {   // shared-memory replay for reads of up to 6000 minimizers; the (rare) longer ones through the global-memory functor
            if (rt.first((const void*)ambig_resolve_smem_kernel)) MM_CUDA(cudaFuncSetAttribute(ambig_resolve_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 48000));
            ambigBig.ensure((size_t)na + 1); scal2.ensure(2);
            MM_CUDA(cudaMemsetAsync(scal2.p, 0, sizeof(unsigned long long), rt.side));
            ambig_resolve_smem_kernel<<<(unsigned)na, 32, 48000, rt.side>>>(rf, 6000, ambigBig.p, scal2.p);
            MM_CUDA(cudaGetLastError());
            rt.launches++;
            rf.list = ambigBig.p;                                   // no host round trip: the functor checks the device-side count
            foreach(rt, (int64_t)na, AmbigResolveBigFn{rf, scal2.p}, 128, 16, true);
          }
#else
          foreach(rt, (int64_t)na, rf, 128, 16, true);
#endif
#ifndef MM_HOST_EMU
          MM_CUDA(cudaEventRecord(evJoin(), rt.side));
#endif
          ambigPending = true;
        }
        maxSketch = maxS;
      }
    }
    ambigPending_ = ambigPending;
  }
  // the std::sort replay of sketch_reads (side stream) must be complete before qStrand is read by anything but K5c
  void join_sketch() {
    if (ambigPending_) {
#ifndef MM_HOST_EMU
      MM_CUDA(cudaStreamWaitEvent(rt.stream, evJoin(), 0));
#endif
      ambigPending_ = false;
    }
  }
  // K4 + K5 over the sketches sketch_reads (or the all-gather of several ranks' sketch_reads) left in this object
  void map_sketched(const Index& ix, float pi, int64_t* summary /*6*/) {
    const int k = ix.k, w = ix.w;
    bool ambigPending = ambigPending_;
    const int64_t nShort = nShort_, basesOk = basesOk_;
    // minimumHits[s] / acceptMin[s] tables up to the largest sketch of the batch (host, map_stats.hpp)
    ensure_tables(k, pi, maxSketch);
    // ---- K4: probe, gather, sort, candidate regions
    candOff.ensure((size_t)n_reads + 2); candCnt.ensure((size_t)n_reads + 2);
    const HitKeyLayout lay = hit_key_layout(ix);
    int readBits = 1; while (((int64_t)1 << readBits) < (n_reads > 1 ? n_reads : 2)) readBits++;
    if (readBits + lay.seqBits + lay.wsBits > 64) throw Error(-34, "read batch too large for the 64-bit hit key: map fewer reads per call");
    HitDecode dec{lay};
    {
      StageTimer t(rt, &st.ms[2]);
      readHitOff.ensure((size_t)n_reads + 2);
      bool fusedDone = false;
#ifndef MM_HOST_EMU
      {
        // K4 in one kernel (l1_probe_filter_kernel) whenever the 2-byte contig ids exist; MM_L1_FUSED=0 = the two-kernel route below
        static const bool fusedOff = [] { const char* e = getenv("MM_L1_FUSED"); return e && atoi(e) == 0; }();
        const char* lf = getenv("MM_L1_FILTER"); const bool legacy = lf && !strcmp(lf, "legacy");
        if (!fusedOff && !legacy && ix.hasSeq16 && n_q > 0 && n_reads > 0 && ix.n < ((int64_t)1 << 32)) {
          uint32_t binsN = 64; while (binsN < (uint32_t)ix.n_contigs && binsN < 32768u) binsN <<= 1;
          static int cacheCap = 0;        // contig ids cached between the passes: 8192 of the ~7000 hits of a read (12288: one CTA fewer per SM, 10.9 ms; 4096: 10.8; 8192: 8.6)
          if (!cacheCap) { const char* e = getenv("MM_L1_CACHE"); cacheCap = e ? atoi(e) : 8192; if (cacheCap < 0 || cacheCap > 14336) cacheCap = 8192; cacheCap &= ~7; }
          static const int pfChunk = [] { const char* e = getenv("MM_L1_CHUNK"); return (e && atoi(e) == 1024) ? 1024 : 512; }();
          const size_t smem = (size_t)binsN * 2 + (size_t)cacheCap * 2 + (size_t)(3 * pfChunk + 8) * 4;
          if (rt.first((const void*)l1_probe_filter_kernel<512>)) {
            MM_CUDA(cudaFuncSetAttribute(l1_probe_filter_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
            MM_CUDA(cudaFuncSetAttribute(l1_probe_filter_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
          }
          int perSm = (int)((220 * 1024) / (smem + 2048)); if (perSm > 8) perSm = 8; if (perSm < 1) perSm = 1;
          if (const char* e = getenv("MM_L1_CTAS")) { int v = atoi(e); if (v >= 1 && v <= perSm) perSm = v; }
          const int grid = n_reads < rt.sm_count * perSm ? n_reads : rt.sm_count * perSm;
          probeOut.ensure((size_t)n_q + 8); keptPerRead.ensure((size_t)n_reads + 2); scal.ensure(8); segStart.ensure((size_t)n_reads + 2);
          if (hits.cap < (size_t)n_q + 1) hits.ensure((size_t)n_q + 1);          // first guess: one survivor per sketch element; grow-only
          unsigned long long hs[3] = {0, 0, 0};
          for (int attempt = 0; attempt < 2; attempt++) {
            dev_memset(rt, scal.p, 0, sizeof(unsigned long long) * 4);
            StageTimer tk(rt, &st.ms[13]);                                          // the fused kernel proper
            if (pfChunk == 512)
              l1_probe_filter_kernel<512><<<grid, 256, smem, rt.stream>>>(ix.table.p, ix.tableMask, ix.freqThreshold, qHash.p, qOff.p, sOf.p, dMinHits.p, ix.posSeq16.p,
                                                                          ix.posKey.p, lay, n_reads, binsN - 1, scal.p, hits.p, (unsigned long long)hits.cap,
                                                                          keptPerRead.p, (uint32_t)cacheCap, probeOut.p, segStart.p);
            else
              l1_probe_filter_kernel<1024><<<grid, 256, smem, rt.stream>>>(ix.table.p, ix.tableMask, ix.freqThreshold, qHash.p, qOff.p, sOf.p, dMinHits.p, ix.posSeq16.p,
                                                                           ix.posKey.p, lay, n_reads, binsN - 1, scal.p, hits.p, (unsigned long long)hits.cap,
                                                                           keptPerRead.p, (uint32_t)cacheCap, probeOut.p, segStart.p);
            MM_CUDA(cudaGetLastError());
            rt.launches++;
            tk.stop();
            d2h(rt, hs, scal.p, sizeof(hs));
            if (!hs[2]) break;
            hits.ensure((size_t)hs[0] + 1);                                        // the survivors did not fit: now they do
          }
          if (hs[2]) throw Error(-12, "l1_probe_filter_kernel: survivor buffer overflow after growing it");
          n_hits = (int64_t)hs[0]; n_hits_all = (int64_t)hs[1];
          dev_memset(rt, keptPerRead.p + n_reads, 0, sizeof(int32_t));
          pr.exclusive_sum<int32_t, int64_t>(keptPerRead.p, readHitOff.p, (int64_t)n_reads + 1);
          hits2.ensure((size_t)n_hits + 1);
          fusedDone = true; segSorted = true;
        }
      }
#endif
      if (!fusedDone) {
      hitCnt.ensure((size_t)n_q + 2); hitStart.ensure((size_t)n_q + 2); hitOff.ensure((size_t)n_q + 2);
#ifndef MM_HOST_EMU
      if (n_q > 0) {
        int64_t tiles = (n_q + PROBE_TILE - 1) / PROBE_TILE;
        int grid = (int)(tiles < (int64_t)rt.sm_count * 4 ? tiles : (int64_t)rt.sm_count * 4);
        l1_probe_tma_kernel<<<grid, 256, 0, rt.stream>>>(ix.table.p, ix.tableMask, qHash.p, ix.freqThreshold, hitCnt.p, hitStart.p, n_q);
        MM_CUDA(cudaGetLastError());
        rt.launches++;
      }
      dev_memset(rt, hitCnt.p + n_q, 0, sizeof(int32_t));
#else
      foreach(rt, n_q + 1, ProbeFn{ix.table.p, ix.tableMask, qHash.p, ix.freqThreshold, hitCnt.p, hitStart.p, n_q});
#endif
      pr.exclusive_sum<int32_t, int64_t>(hitCnt.p, hitOff.p, n_q + 1);
      d2h(rt, &n_hits, hitOff.p + n_q, sizeof(int64_t));
      n_hits_all = n_hits;
#ifndef MM_HOST_EMU
      if (n_hits > 0 && n_reads > 0) {
        // filtered gather (see l1_filter_gather_kernel): survivors only, unordered by read
        uint32_t binsN = 64; while (binsN < (uint32_t)ix.n_contigs && binsN < 32768u) binsN <<= 1;
        hits.ensure((size_t)n_hits + 1); keptPerRead.ensure((size_t)n_reads + 2); scal.ensure(4);
        dev_memset(rt, scal.p, 0, sizeof(unsigned long long));
        if (rt.first((const void*)l1_filter_gather_kernel)) MM_CUDA(cudaFuncSetAttribute(l1_filter_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        const char* lf = getenv("MM_L1_FILTER"); const bool legacy = lf && !strcmp(lf, "legacy");
        if (ix.hasSeq16 && !legacy) {
          if (rt.first((const void*)l1_filter_gather16_kernel)) MM_CUDA(cudaFuncSetAttribute(l1_filter_gather16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
          static int cacheCap = 0;
          if (!cacheCap) { const char* e = getenv("MM_L1_CACHE"); cacheCap = e ? atoi(e) : 12288; if (cacheCap < 0 || cacheCap > 14336) cacheCap = 12288; }
          size_t smem = (size_t)binsN * 2 + (size_t)cacheCap * 2;
          int perSm = (int)((220 * 1024) / (smem + 1024)); if (perSm > 8) perSm = 8; if (perSm < 1) perSm = 1;
          int grid = n_reads < rt.sm_count * perSm ? n_reads : rt.sm_count * perSm;
          // (one list per group of 8 lanes instead of the flattened walk -- no search, direct ranks -- was measured on config 2:
          //  the L1 stage went 10.0 -> 15.5 ms: with 7.5 entries per list on average too many lanes idle; removed)
          l1_filter_gather16_kernel<<<grid, 256, smem, rt.stream>>>(hitCnt.p, hitStart.p, hitOff.p, qOff.p, sOf.p, dMinHits.p, ix.posSeq16.p, ix.posKey.p, lay, n_reads,
                                                                    binsN - 1, scal.p, hits.p, keptPerRead.p, (uint32_t)cacheCap);
        } else {
          int grid = n_reads < rt.sm_count * 8 ? n_reads : rt.sm_count * 8;
          l1_filter_gather_kernel<<<grid, 256, binsN * 2, rt.stream>>>(hitCnt.p, hitStart.p, qOff.p, sOf.p, dMinHits.p, ix.posKey.p, lay, n_reads, binsN - 1,
                                                                       scal.p, hits.p, keptPerRead.p);
        }
        MM_CUDA(cudaGetLastError());
        rt.launches++;
        unsigned long long kept = 0; d2h(rt, &kept, scal.p, sizeof(kept));
        n_hits = (int64_t)kept;
        dev_memset(rt, keptPerRead.p + n_reads, 0, sizeof(int32_t));
        pr.exclusive_sum<int32_t, int64_t>(keptPerRead.p, readHitOff.p, (int64_t)n_reads + 1);
        hits2.ensure((size_t)n_hits + 1);
      } else
#endif
      {
        hits.ensure((size_t)n_hits + 1); hits2.ensure((size_t)n_hits + 1);
        foreach(rt, n_q, GatherHitsFn{hitCnt.p, hitStart.p, hitOff.p, qRead.p, ix.posKey.p, hits.p, lay});
        foreach(rt, (int64_t)n_reads + 1, ReadHitOffFn{qOff.p, hitOff.p, readHitOff.p});
      }
      }   // !fusedDone
    }
    {
      StageTimer t(rt, &st.ms[3]);
      bool sorted = false;
#ifndef MM_HOST_EMU
      static const bool segOff = [] { const char* e = getenv("MM_L1_SEGSORT"); return e && atoi(e) == 0; }();
      if (segSorted && !segOff && n_hits > 0) {            // the fused kernel left one segment per read: sort inside the segments (l1_sort_segments_*)
        segBig.ensure((size_t)n_reads + 1);
        dev_memset(rt, scal.p + 4, 0, sizeof(unsigned long long) * 3);       // [4] = big reads, [5] = read cursor, [6] = overflow flag
        const size_t smW = (size_t)8 * SEG_SORT_WARP_CAP * 8, smC = (size_t)SEG_SORT_CTA_CAP * 8;
        if (rt.first((const void*)l1_sort_segments_warp_kernel)) {
          MM_CUDA(cudaFuncSetAttribute(l1_sort_segments_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smW));
          MM_CUDA(cudaFuncSetAttribute(l1_sort_segments_cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smC));
        }
        l1_sort_segments_warp_kernel<<<rt.sm_count * 3, 256, smW, rt.stream>>>(hits.p, segStart.p, keptPerRead.p, readHitOff.p, n_reads, hits2.p, segBig.p, scal.p + 4, scal.p + 5);
        l1_sort_segments_cta_kernel<<<rt.sm_count * 3, 256, smC, rt.stream>>>(hits.p, segStart.p, keptPerRead.p, readHitOff.p, hits2.p, segBig.p, scal.p + 4, (int32_t*)(scal.p + 6));
        MM_CUDA(cudaGetLastError());
        rt.launches += 2;
        unsigned long long ovf = 0; d2h(rt, &ovf, scal.p + 6, sizeof(ovf));
        sorted = (ovf & 0xffffffffull) == 0;               // a read with more survivors than a CTA sorts: the radix sort below does the whole batch
      }
#endif
      segSorted = false;
      if (!sorted) pr.sort_keys<uint64_t>(hits.p, hits2.p, n_hits, readBits + lay.seqBits + lay.wsBits);
    }
    {
      StageTimer t(rt, &st.ms[4]);
      n_cand = 0;
      dev_memset(rt, candCnt.p, 0, sizeof(int32_t) * ((size_t)n_reads + 1));
      bool candDone = false;
#ifndef MM_HOST_EMU
      static const bool candLegacy = [] { const char* e = getenv("MM_L1_CAND"); return e && !strcmp(e, "legacy"); }();
      if (!candLegacy && n_hits > 0 && n_hits < ((int64_t)1 << 31)) {          // one warp per read over its sorted hits (l1_candidates_warp_kernel)
        candTmp.ensure((size_t)5 * ((size_t)n_hits + 2));
        int32_t* tSeq = candTmp.p; int32_t* tStart = tSeq + n_hits + 2; int32_t* tEnd = tStart + n_hits + 2; int32_t* tLo = tEnd + n_hits + 2; int32_t* tHi = tLo + n_hits + 2;
        scal.ensure(8); dev_memset(rt, scal.p + 7, 0, sizeof(unsigned long long));
        l1_candidates_warp_kernel<<<rt.sm_count * 8, 256, 0, rt.stream>>>(hits2.p, readHitOff.p, sOf.p, readLen.p, dMinHits.p, dec, n_reads, tSeq, tStart, tEnd, tLo, tHi,
                                                                          candCnt.p, scal.p + 7);
        MM_CUDA(cudaGetLastError());
        rt.launches++;
        pr.exclusive_sum<int32_t, int64_t>(candCnt.p, candOff.p, (int64_t)n_reads + 1);
        d2h(rt, &n_cand, candOff.p + n_reads, sizeof(int64_t));
        cRead.ensure((size_t)n_cand + 1); cSeq.ensure((size_t)n_cand + 1); cStart.ensure((size_t)n_cand + 1); cEnd.ensure((size_t)n_cand + 1);
        cHitLo.ensure((size_t)n_cand + 1); cHitHi.ensure((size_t)n_cand + 1);
        if (n_cand > 0) foreach(rt, n_cand, CandGatherFn{candOff.p, n_reads, readHitOff.p, tSeq, tStart, tEnd, tLo, tHi, cRead.p, cSeq.p, cStart.p, cEnd.p, cHitLo.p, cHitHi.p});
        candDone = true;
      }
#endif
      if (!candDone && n_hits > 0) {
        hflag.ensure((size_t)n_hits + 2); hfidx.ensure((size_t)n_hits + 2);
        foreach(rt, n_hits + 1, HitFlagFn{hits2.p, readHitOff.p, sOf.p, readLen.p, dMinHits.p, dec, hflag.p, n_hits});
        pr.exclusive_sum<int32_t, int64_t>(hflag.p, hfidx.p, n_hits + 1);
        int64_t nFlag = 0; d2h(rt, &nFlag, hfidx.p + n_hits, sizeof(int64_t));
        if (nFlag > 0) {
          flagged.ensure((size_t)nFlag + 1); lhead.ensure((size_t)nFlag + 2); lidx.ensure((size_t)nFlag + 2);
          foreach(rt, n_hits, HitCompactFn{hflag.p, hfidx.p, flagged.p});
          foreach(rt, nFlag + 1, LocusHeadFn{hits2.p, flagged.p, sOf.p, readLen.p, dMinHits.p, dec, lhead.p, nFlag});
          pr.exclusive_sum<int32_t, int64_t>(lhead.p, lidx.p, nFlag + 1);
          d2h(rt, &n_cand, lidx.p + nFlag, sizeof(int64_t));
          cRead.ensure((size_t)n_cand + 1); cSeq.ensure((size_t)n_cand + 1); cStart.ensure((size_t)n_cand + 1); cEnd.ensure((size_t)n_cand + 1);
          cHitLo.ensure((size_t)n_cand + 1); cHitHi.ensure((size_t)n_cand + 1);
          foreach(rt, nFlag, LocusWriteFn{hits2.p, flagged.p, sOf.p, readLen.p, dMinHits.p, dec, lhead.p, lidx.p, nFlag,
                                          cRead.p, cSeq.p, cStart.p, cEnd.p, candCnt.p, cHitLo.p, cHitHi.p});
        }
      }
      if (!candDone) pr.exclusive_sum<int32_t, int64_t>(candCnt.p, candOff.p, (int64_t)n_reads + 1);
      cRead.ensure((size_t)n_cand + 1); cSeq.ensure((size_t)n_cand + 1); cStart.ensure((size_t)n_cand + 1); cEnd.ensure((size_t)n_cand + 1);
      cHitLo.ensure((size_t)n_cand + 1); cHitHi.ensure((size_t)n_cand + 1);
    }
    // ---- K5
    int64_t totalEv = 0, smemSwept = 0;
    int32_t cntBytes = 2;
    oShared.ensure((size_t)n_cand + 1); oPos.ensure((size_t)n_cand + 1); oValid.ensure((size_t)n_cand + 1); oIstar.ensure((size_t)n_cand + 1);
    oVotes.ensure((size_t)n_cand + 1); oAccept.ensure((size_t)n_cand + 1); oOptS.ensure((size_t)n_cand + 1); oOptE.ensure((size_t)n_cand + 1);
    readMapped.ensure((size_t)n_reads + 1); dev_memset(rt, readMapped.p, 0, sizeof(int32_t) * ((size_t)n_reads + 1));
    if (n_cand > 0) {
      beg0.ensure((size_t)n_cand + 1); fe.ensure((size_t)n_cand + 1); le.ensure((size_t)n_cand + 1); cHits.ensure((size_t)n_cand + 1);
      spanN.ensure((size_t)n_cand + 2); stWords.ensure((size_t)n_cand + 2); evOff.ensure((size_t)n_cand + 2); stOff.ensure((size_t)n_cand + 2);
      {
        StageTimer t(rt, &st.ms[5]);
        foreach(rt, n_cand + 1, L2SetupFn{ix.miWs.p, ix.contigStart.p, cRead.p, cSeq.p, cStart.p, cEnd.p, readLen.p, sOf.p, k, w,
                                           beg0.p, fe.p, le.p, spanN.p, n_cand, cHitLo.p, cHitHi.p, cHits.p});
        pr.exclusive_sum<int32_t, int64_t>(spanN.p, evOff.p, n_cand + 1);
      }
      // the usual case is ONE pass over all candidates: only the total and the largest span come to the host then; the full
      // offset array is fetched when the event budget forces several passes
      std::vector<int64_t>& hEv = hEvSpan;
      red.ensure(2);
      pr.reduce_max<int32_t>(spanN.p, red.p, n_cand);
      d2d(rt, scal.p, evOff.p + n_cand, sizeof(int64_t));
      d2d(rt, scal.p + 1, red.p, sizeof(int32_t));
      unsigned long long hs2[2] = {0, 0}; d2h(rt, hs2, scal.p, sizeof(hs2));
      totalEv = (int64_t)hs2[0];
      const int32_t maxSpan = (int32_t)(hs2[1] & 0xffffffffu);
      if (totalEv <= evBudget) { hEv.assign((size_t)n_cand + 1, 0); hEv[(size_t)n_cand] = totalEv; if (maxSpan >= 65535) cntBytes = 4; }
      else { hEv.resize((size_t)n_cand + 1); d2h(rt, hEv.data(), evOff.p, sizeof(int64_t) * hEv.size()); }
      // a gap counter never exceeds the number of minimizers in the span: 16 bits unless some span is huge
      if (totalEv > evBudget) for (int64_t c = 0; c < n_cand; c++) if (hEv[(size_t)c + 1] - hEv[(size_t)c] >= 65535) { cntBytes = 4; break; }
      int64_t c0 = 0;
      while (c0 < n_cand) {          // passes bounded by the event budget
        int64_t c1 = c0 + 1;
        if (totalEv <= evBudget) c1 = n_cand;
        while (c1 < n_cand && hEv[(size_t)c1 + 1] - hEv[(size_t)c0] <= evBudget) c1++;
        int64_t nEv = hEv[(size_t)c1] - hEv[(size_t)c0], nc = c1 - c0;
        ev.ensure((size_t)nEv + 64);      // the sweeps prefetch a few events past the end of a span
        {
          StageTimer t(rt, &st.ms[6]);
          bool prune = sweepPrune && sweepMode == 0 && nc < ((int64_t)1 << 31);
#ifndef MM_HOST_EMU
          const bool fast = (int64_t)maxSketch * 4 <= 200 * 1024 && maxSketch < 65535;      // bucket starts are 16-bit ranks
          prune = prune && fast;                           // the prune pass is part of the shared-memory kernel
#endif
          if (prune) { swB0.ensure((size_t)nc + 1); swB1.ensure((size_t)nc + 1); grpSum.ensure((size_t)(nEv >> 5) + (size_t)nc + 4); }
          prunedPass = prune;
          L2ClassifyFn cf{ix.miHash.p, ix.miWs.p, ix.dupBits.p, evOff.p, c0, nc, hEv[(size_t)c0], beg0.p, cRead.p, qHash.p, qOff.p, sOf.p, ev.p,
                          fe.p, le.p, readLen.p, ix.dupRB.p, ix.dupLinks.p, ix.n_dup, k, w, cHits.p, prune ? grpSum.p : nullptr, pruneGmax};
#ifndef MM_HOST_EMU
          if (fast) {
            // contiguous runs of candidates per CTA, ~8 waves of CTAs so that uneven runs average out
            int64_t g = (int64_t)rt.sm_count * 64; if (g > nc) g = nc;
            int32_t perCta = (int32_t)((nc + g - 1) / g); g = (nc + perCta - 1) / perCta;
            auto launch = [&](auto kern) {
              if (rt.first((const void*)kern)) MM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
              kern<<<(int)g, 128, (size_t)(maxSketch > 0 ? maxSketch : 1) * 4, rt.stream>>>(cf, perCta);
            };
            if (prune) launch(l2_classify_smem_kernel<true, 11>); else launch(l2_classify_smem_kernel<false, 11>);
            MM_CUDA(cudaGetLastError());
            rt.launches++;
          } else
#endif
            foreach(rt, nEv, cf);
        }
        if (prunedPass) {
          StageTimer tp(rt, &st.ms[12]);
#ifdef MM_HOST_EMU
          foreach(rt, nc, L2PruneFn{ev.p, evOff.p, hEv[(size_t)c0], c0, beg0.p, fe.p, cRead.p, sOf.p, readLen.p, cHits.p, k, w, swB0.p, swB1.p, pruneGmax});
#else
          int32_t gmax = (maxSpan + 31) >> 5; if (gmax > pruneGmax) gmax = pruneGmax; if (gmax < 1) gmax = 1;
          const size_t smem = prune_warp_words(gmax) * 4 * PRUNE_WARPS;
          if (rt.first((const void*)l2_prune_warp_kernel)) MM_CUDA(cudaFuncSetAttribute(l2_prune_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(prune_warp_words(PR_GMAX) * 4 * PRUNE_WARPS)));
          int64_t g = (nc + PRUNE_WARPS - 1) / PRUNE_WARPS; if (g > (int64_t)rt.sm_count * 8) g = (int64_t)rt.sm_count * 8;
          l2_prune_warp_kernel<<<(int)g, PRUNE_WARPS * 32, smem, rt.stream>>>(L2PruneArgs{grpSum.p, evOff.p, hEv[(size_t)c0], c0, beg0.p, fe.p, cRead.p, sOf.p, readLen.p, cHits.p, k, w, swB0.p, swB1.p}, nc, gmax);
          MM_CUDA(cudaGetLastError());
          rt.launches++;
#endif
        }
        L2SweepArgs sa{ev.p, evOff.p, hEv[(size_t)c0], c0, beg0.p, fe.p, le.p, cRead.p, sOf.p, readLen.p, ix.dupRB.p, ix.dupLinks.p, ix.n_dup, k, w,
                       oShared.p, oPos.p, oValid.p, oOptS.p, oOptE.p, oIstar.p};
        {
          StageTimer t(rt, &st.ms[7]);
          smemSwept += sweep_pass(sa, nc, cntBytes);
        }
        if (ambigPending) {
#ifndef MM_HOST_EMU
          MM_CUDA(cudaStreamWaitEvent(rt.stream, evJoin(), 0));
#endif
          ambigPending = false;
        }
        {
          StageTimer t(rt, &st.ms[8]);
          L2StrandFn sf{ev.p, evOff.p, hEv[(size_t)c0], c0, beg0.p, cRead.p, qOff.p, qStrand.p, ix.dupRB.p, ix.dupLinks.p, ix.n_dup, ix.dupBits.p,
                        oValid.p, oOptS.p, oOptE.p, oIstar.p, oVotes.p};
#ifndef MM_HOST_EMU
          {
            int64_t g = (nc + 7) / 8; if (g > (int64_t)rt.sm_count * 8) g = (int64_t)rt.sm_count * 8;
            l2_strand_warp_kernel<<<(int)g, 256, 0, rt.stream>>>(sf, nc);
            MM_CUDA(cudaGetLastError());
            rt.launches++;
          }
#else
          foreach(rt, nc, sf, 128, 16);
#endif
        }
        c0 = c1;
      }
    }
    if (ambigPending) {
#ifndef MM_HOST_EMU
      MM_CUDA(cudaStreamWaitEvent(rt.stream, evJoin(), 0));
#endif
    }
    int64_t nMap = 0, nReadsMapped = 0;
    {
      StageTimer t(rt, &st.ms[9]);
      if (n_cand > 0) {
        foreach(rt, n_cand, AcceptFn{cRead.p, sOf.p, dAccept.p, oShared.p, oValid.p, oAccept.p, readMapped.p});
        red.ensure(2);
        pr.reduce_sum<int32_t>(oAccept.p, red.p, n_cand);
        pr.reduce_sum<int32_t>(readMapped.p, red.p + 1, n_reads);
        int32_t h[2]; d2h(rt, h, red.p, sizeof(h)); nMap = h[0]; nReadsMapped = h[1];
      }
    }
    rt.sync();
    st.counters[0] = n_q; st.counters[1] = n_hits_all; st.counters[10] = n_hits; st.counters[2] = n_cand; st.counters[3] = totalEv; st.counters[4] = nMap;
    st.counters[5] = nMinimizers_; st.counters[6] = basesOk; st.counters[7] = nExc_; st.counters[8] = n_ambig; st.counters[9] = smemSwept;
    ambigPending_ = false;
    summary[0] = n_reads; summary[1] = nShort; summary[2] = n_cand; summary[3] = nMap; summary[4] = nReadsMapped; summary[5] = basesOk;
  }

  // K5b over the candidates [sa.cand0, sa.cand0+nc) of one pass.  Returns how many were swept by the fast path.
  //   device: banded sweep in shared memory (l2_sweep_band_kernel); MM_SWEEP=full selects the older full-state
  //           shared-memory kernel instead, MM_SWEEP=global sends everything through the global-memory functor;
  //   host emulation: the same banded logic with its state in global memory (L2SweepBandFn).
  // Whatever a fast path declines (oversize sketches, counter overflow) goes through L2SweepFn (full state, global memory).
  int64_t sweep_pass(const L2SweepArgs& sa, int64_t nc, int32_t cntBytes) {
    int64_t done_fast = 0;
    const int32_t* redoList = nullptr; int64_t nRedo = nc;        // default: everything through the global-memory functor
    const int BAND = sweepBand, MODE = sweepMode;                 // MODE 0 band, 1 full-state smem, 2 global only
    swRedo.ensure((size_t)nc + 1); scal.ensure(8);
    if (MODE == 0 && nc > 0 && nc < ((int64_t)1 << 31) && maxSketch < (1 << 20)) {
      // work items = segments of candidates
      segCnt.ensure((size_t)nc + 2); itemOff.ensure((size_t)nc + 2);
      const int32_t* cb0 = prunedPass ? swB0.p : nullptr; const int32_t* cb1 = prunedPass ? swB1.p : nullptr;
      dev_memset(rt, scal.p, 0, sizeof(unsigned long long) * 6);          // scal[0] = redo count, [1] = tile counter, [2] = wide-band items, [4] / [5] = window starts to sweep / in all
      foreach(rt, nc + 1, BandSegCountFn{spanN.p, beg0.p, fe.p, sa.cand0, nc, sweepSeg, segCnt.p, cb0, cb1, scal.p + 4});
      pr.exclusive_sum<int32_t, int64_t>(segCnt.p, itemOff.p, nc + 1);
      int64_t nItems = 0; d2h(rt, &nItems, itemOff.p + nc, sizeof(int64_t));
      itemCand.ensure((size_t)nItems); itemB0.ensure((size_t)nItems); itemB1.ensure((size_t)nItems); bandParts.ensure((size_t)nItems);
      swKey.ensure((size_t)nItems); swKey2.ensure((size_t)nItems); swVal.ensure((size_t)nItems); swOrder.ensure((size_t)nItems);
      foreach(rt, nItems, BandItemFn{itemOff.p, nc, spanN.p, beg0.p, fe.p, sa.cand0, sweepSeg, swKey.p, swVal.p, itemCand.p, itemB0.p, itemB1.p,
                                     cRead.p, sOf.p, sweepWideFrom, scal.p + 2, cb0, cb1});
#ifdef MM_HOST_EMU
      state.ensure((size_t)nItems * band_state_words(BAND) + 1);
      foreach(rt, nItems, L2SweepBandFn{sa, state.p, BAND, itemCand.p, itemB0.p, itemB1.p, bandParts.p});
#else
      pr.sort_pairs<uint32_t, uint32_t>(swKey.p, swKey2.p, swVal.p, swOrder.p, nItems, 32);
      unsigned long long nWide = 0; d2h(rt, &nWide, scal.p + 2, sizeof(nWide));
      {
        StageTimer tb(rt, &st.ms[11]);                                      // the sweep kernel proper
        if (nWide > 0) { if (BAND == 256) launch_band_t<512, 8>(sa, swOrder.p, (int64_t)nWide); else launch_band_t<256, 8>(sa, swOrder.p, (int64_t)nWide); }
        if ((int64_t)nWide < nItems) {
          if (nWide > 0) dev_memset(rt, scal.p + 1, 0, sizeof(unsigned long long));       // fresh tile counter
          launch_band(sa, swOrder.p + nWide, nItems - (int64_t)nWide);
        }
      }
      st.counters[11] += nItems;
#endif
      foreach(rt, nc, L2BandMergeFn{sa, bandParts.p, itemOff.p, swRedo.p, scal.p});
      unsigned long long nr[6] = {0, 0, 0, 0, 0, 0}; d2h(rt, nr, scal.p, sizeof(nr));
      nRedo = (int64_t)nr[0]; redoList = swRedo.p; done_fast = nc - nRedo;
      st.counters[12] += (int64_t)nr[4]; st.counters[13] += (int64_t)nr[5];
    }
#ifndef MM_HOST_EMU
    if (MODE == 1 && nc >= 1 && nc < ((int64_t)1 << 31) && maxSketch < (1 << 20)) {
      // order: descending sketch size, so that the lanes of a warp run similar loops
      swKey.ensure((size_t)nc); swKey2.ensure((size_t)nc); swVal.ensure((size_t)nc); swOrder.ensure((size_t)nc);
      foreach(rt, nc, SweepKeyFn{cRead.p, sOf.p, nullptr, sa.cand0, swKey.p, swVal.p});
      pr.sort_pairs<uint32_t, uint32_t>(swKey.p, swKey2.p, swVal.p, swOrder.p, nc, 32);
      dev_memset(rt, scal.p, 0, sizeof(unsigned long long) * 2);          // scal[0] = redo count, scal[1] = tile counter
      {
        static int SWEEP_WARPS = 0;
        if (!SWEEP_WARPS) {
          const char* ev_ = getenv("MM_SWEEP_WARPS");
          SWEEP_WARPS = ev_ ? atoi(ev_) : 8;
          if (SWEEP_WARPS < 1 || SWEEP_WARPS > SWEEP_WARPS_MAX) SWEEP_WARPS = 8;
        }
        if (rt.first((const void*)l2_sweep_smem_kernel)) MM_CUDA(cudaFuncSetAttribute(l2_sweep_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SWEEP_SMEM_WORDS * 4));
        const int32_t SLICE = SWEEP_SMEM_WORDS / SWEEP_WARPS;
        hk.resize((size_t)nc);
        d2h(rt, hk.data(), swKey2.p, sizeof(uint32_t) * (size_t)nc);
        // tiles over the descending-s order; candidates whose state does not fit a warp slice go to the global path
        tileStartH.clear(); localOffH.assign((size_t)nc, 0);
        int64_t firstFit = 0;
        auto wordsOf = [](uint32_t s_) { return sweep_state_words((int32_t)s_, 1) | 1; };
        while (firstFit < nc && wordsOf(0xFFFFFFFFu - hk[(size_t)firstFit]) > SLICE) firstFit++;
        int64_t i = firstFit;
        while (i < nc) {
          tileStartH.push_back((int32_t)i);
          int32_t used = 0, cnt = 0;
          while (i < nc && cnt < 32) {
            int32_t wds = wordsOf(0xFFFFFFFFu - hk[(size_t)i]);
            if (used + wds > SLICE) break;
            localOffH[(size_t)i] = used; used += wds; cnt++; i++;
          }
        }
        tileStartH.push_back((int32_t)nc);
        int32_t nTiles = (int32_t)tileStartH.size() - 1;
        if (nTiles > 0) {
          swTile.ensure(tileStartH.size()); swLocal.ensure((size_t)nc);
          h2d(rt, swTile.p, tileStartH.data(), sizeof(int32_t) * tileStartH.size());
          h2d(rt, swLocal.p, localOffH.data(), sizeof(int32_t) * (size_t)nc);
          if (firstFit > 0) {          // candidates that fit no slice are pre-loaded into the redo list
            d2d(rt, swRedo.p, swOrder.p, sizeof(int32_t) * (size_t)firstFit);
            unsigned long long ff = (unsigned long long)firstFit; h2d(rt, scal.p, &ff, sizeof(ff));
          }
          int grid = (nTiles + SWEEP_WARPS - 1) / SWEEP_WARPS; if (grid > rt.sm_count) grid = rt.sm_count;
          l2_sweep_smem_kernel<<<grid, SWEEP_WARPS * 32, SWEEP_WARPS * SLICE * 4, rt.stream>>>(sa, swOrder.p, swTile.p, nTiles, swLocal.p,
                                                                                               (unsigned int*)(scal.p + 1), swRedo.p, scal.p, SLICE);
          MM_CUDA(cudaGetLastError());
          rt.launches++;
          unsigned long long nr = 0; d2h(rt, &nr, scal.p, sizeof(nr));
          nRedo = (int64_t)nr; redoList = swRedo.p;
          done_fast = nc - nRedo;
        }
      }
    }
#endif
    if (nRedo > 0) {
      // full state in global memory for the rest
      stWords.ensure((size_t)n_cand + 2); stOff.ensure((size_t)n_cand + 2);
      foreach(rt, n_cand + 1, StWordsFn{cRead.p, sOf.p, stWords.p, n_cand, cntBytes});
      pr.exclusive_sum<int32_t, int64_t>(stWords.p, stOff.p, n_cand + 1);
      int64_t stRange[2];
      d2h(rt, &stRange[0], stOff.p + sa.cand0, sizeof(int64_t));
      d2h(rt, &stRange[1], stOff.p + sa.cand0 + nc, sizeof(int64_t));
      int64_t nSt = stRange[1] - stRange[0];
      state.ensure((size_t)nSt + 1);
      dev_memset(rt, state.p, 0, sizeof(uint32_t) * (size_t)nSt);
      if (cntBytes == 2) foreach(rt, nRedo, L2SweepFn<uint16_t>{sa, state.p, stOff.p, stRange[0], redoList}, 128, 16);
      else foreach(rt, nRedo, L2SweepFn<uint32_t>{sa, state.p, stOff.p, stRange[0], redoList}, 128, 16);
    }
    return done_fast;
  }

#ifndef MM_HOST_EMU
  template <int BW, int R, int MINB = 1>
  void launch_band_t(const L2SweepArgs& sa, const uint32_t* order, int64_t nc) {
    constexpr int ST = ((BW + 4) / 4 + BW / 32 + 1) | 1;
    constexpr int WARP_BYTES = (2 * R * 32 * 4 + 32 * ST) * 4;
    static int warps = 0, ctasPerSm = 1;
    if (!warps) {
      int total = (220 * 1024) / WARP_BYTES;                 // warps that fit one SM's shared memory
      if (total > 32) total = 32;
      ctasPerSm = MINB;                                      // MINB = 2: two CTAs of 12 warps (the register cap of launch_bounds(384, 2) makes them fit)
      if (total > 16 * ctasPerSm) total = 16 * ctasPerSm;
      if (MINB == 2 && total > 24) total = 24;
      warps = total / ctasPerSm;
      if (const char* e = getenv("MM_SWEEP_WARPS")) { int v = atoi(e); if (v >= 1 && v <= 16) warps = v; }
    }
    if (rt.first((const void*)l2_sweep_band_kernel<BW, R, MINB>)) MM_CUDA(cudaFuncSetAttribute(l2_sweep_band_kernel<BW, R, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, warps * WARP_BYTES));
    int64_t tiles = (nc + 31) / 32;
    int64_t g = (tiles + warps - 1) / warps; if (g > (int64_t)rt.sm_count * ctasPerSm) g = (int64_t)rt.sm_count * ctasPerSm;
    l2_sweep_band_kernel<BW, R, MINB><<<(int)g, warps * 32, (size_t)warps * WARP_BYTES, rt.stream>>>(sa, order, nc, itemCand.p, itemB0.p, itemB1.p, bandParts.p,
                                                                                               (unsigned int*)(scal.p + 1));
    MM_CUDA(cudaGetLastError());
    rt.launches++;
  }
  void launch_band(const L2SweepArgs& sa, const uint32_t* order, int64_t nc) {
    const int BAND = sweepBand, RING = sweepRing;
    static const bool two = getenv("MM_SWEEP_2CTA") != nullptr;        // A/B: 24 warps per SM under an 85-register cap
    if (BAND == 64) launch_band_t<64, 4>(sa, order, nc);
    else if (BAND == 128 && RING == 4 && two) launch_band_t<128, 4, 2>(sa, order, nc);
    else if (BAND == 128 && RING == 4) launch_band_t<128, 4>(sa, order, nc);
    else if (BAND == 128) launch_band_t<128, 8>(sa, order, nc);
    else if (RING == 2) launch_band_t<256, 2>(sa, order, nc);
    else if (RING == 4) launch_band_t<256, 4>(sa, order, nc);
    else launch_band_t<256, 8>(sa, order, nc);
  }
#endif

#ifndef MM_HOST_EMU
  cudaEvent_t evF = nullptr, evJ = nullptr;
  cudaEvent_t evFork() { if (!evF) MM_CUDA(cudaEventCreateWithFlags(&evF, cudaEventDisableTiming)); return evF; }
  cudaEvent_t evJoin() { if (!evJ) MM_CUDA(cudaEventCreateWithFlags(&evJ, cudaEventDisableTiming)); return evJ; }
#endif
};

}  // namespace mm
