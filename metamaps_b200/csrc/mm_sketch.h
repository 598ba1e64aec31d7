// mm_sketch.h -- K0 (2-bit pack) and K1 (winnowed minimizers).
//
// Replaces CommonFunc::addMinimizers (reference src/map/include/commonFunc.hpp:92-175).
//
// Device data layout of a sequence batch (SeqBatch):
//   packed[]   uint32 words, 16 bases per word, 2 bits per base (A=0 C=1 T=2 G=3), base j of sequence i at
//              word wordOff[i] + j/16, bits 2*(j%16); every sequence starts on a word boundary
//   excPos[] / excByte[]   sorted side list of the bases that are not A/C/G/T after upper-casing
//              (global base index = 16*word + slot, upper-cased byte).  The reference hashes such bytes
//              verbatim (commonFunc.hpp:44-51,106), so K1 patches them back in when a k-mer overlaps one.
//
// K1 work decomposition: a sequence with npos = len-k+1 k-mer positions is cut into chunks of CH
// positions; one logical thread owns a chunk and replays the reference's monotone deque over
// [c0 - 2(w-1), c1) -- the 2(w-1) halo makes the deque front exact for every step >= c0-(w-1), which is
// all that the emit rule `front != last emitted` (commonFunc.hpp:157) can look at.  Emitted minimizers go
// to a slab indexed by k-mer position (worst case one per position) and are then compacted in order.
#pragma once
#include "mm_murmur.h"
#include "mm_prims.h"

namespace mm {

template <class T>
MM_HD int64_t upper_bound_idx(const T* a, int64_t n, T v) {   // first index with a[i] > v
  int64_t lo = 0, hi = n;
  while (lo < hi) { int64_t m = (lo + hi) >> 1; if (ldg(a + m) <= v) lo = m + 1; else hi = m; }
  return lo;
}
template <class T>
MM_HD int64_t lower_bound_idx(const T* a, int64_t n, T v) {   // first index with a[i] >= v
  int64_t lo = 0, hi = n;
  while (lo < hi) { int64_t m = (lo + hi) >> 1; if (ldg(a + m) < v) lo = m + 1; else hi = m; }
  return lo;
}

static const uint32_t MM_TOMB = 0xFFFFFFFFu;

// ------------------------------------------------------------------------------------------- K0 pack
struct PackFn {
  const uint8_t* asc; const int64_t* ascOff;      // ASCII bytes + byte offset of each sequence
  const int64_t* wordOff; const int32_t* len; int32_t n_seqs;
  uint32_t* packed;
  unsigned long long* excCount; uint64_t* excPos; uint8_t* excByte; int64_t excCap;
  int64_t total_words;
  // item = (group of 256 consecutive words, lane): the lane packs words lane, lane + 32, ... of the group, so a warp writes 32
  // consecutive words per trip and the sequence a word belongs to is searched once per 8 words (then walked forward)
  static constexpr int TRIPS = 8;
  MM_HD void operator()(int64_t item) const {
    int64_t t = (item >> 5) * (32 * TRIPS) + (item & 31);
    if (t >= total_words) return;
    int64_t sq = upper_bound_idx(wordOff, (int64_t)n_seqs + 1, t) - 1;
    for (int i = 0; i < TRIPS && t < total_words; i++, t += 32) {
      while (ldg(wordOff + sq + 1) <= t) sq++;         // t < total_words = wordOff[n_seqs]: stops at the word's sequence
      pack_word(t, sq);
    }
  }
  MM_HD void pack_word(int64_t t, int64_t sq) const {
    int64_t lw = t - ldg(wordOff + sq);
    int32_t L = ldg(len + sq);
    const uint8_t* src = asc + ldg(ascOff + sq) + lw * 16;
    int32_t nb = L - (int32_t)(lw * 16); if (nb > 16) nb = 16;
    // the 16 bytes are fetched as the (at most five) aligned 32-bit words that hold them -- every such word contains at least
    // one byte of the sequence, so the loads stay inside the caller's allocation -- and realigned with funnel shifts
    const uintptr_t addr = (uintptr_t)src; const int mis = (int)(addr & 3);
    const uint32_t* wp = reinterpret_cast<const uint32_t*>(addr - mis);
    const int nWords = (mis + nb + 3) / 4;
    uint32_t wv[5];
#pragma unroll
    for (int i = 0; i < 5; i++) wv[i] = i < nWords ? ldg(wp + i) : 0u;
    uint32_t word = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const uint32_t four = mis ? ((wv[q] >> (8 * mis)) | (wv[q + 1] << (32 - 8 * mis))) : wv[q];
      const int nbq = nb - 4 * q;                      // bytes of this group that belong to the sequence
      if (nbq <= 0) break;
      // four bytes at a time: upper-case, 2-bit codes and the A/C/G/T test as word operations; a group with any other byte
      // (N, IUPAC codes, >= 0x80) takes the byte loop below, which also files the exceptions
      const uint32_t vmask = nbq >= 4 ? 0xFFFFFFFFu : ((1u << (8 * nbq)) - 1u);
      const uint32_t x = four & vmask;
      const uint32_t lower = (x + 0x1f1f1f1fu) & ~(x + 0x05050505u) & 0x80808080u;      // bytes in 'a'..'z' (valid while no byte >= 0x80)
      const uint32_t u4 = x - (lower >> 2);
      const uint32_t cd = (u4 >> 1) & 0x03030303u;
      uint32_t rec = 0;                                // the ASCII letters the codes stand for
#if defined(__CUDA_ARCH__)
      rec = __byte_perm(0x47544341u, 0u, (cd & 0x3u) | ((cd >> 4) & 0x30u) | ((cd >> 8) & 0x300u) | ((cd >> 12) & 0x3000u));
#else
      for (int b4 = 0; b4 < 4; b4++) rec |= code_to_ascii((cd >> (8 * b4)) & 3u) << (8 * b4);
#endif
      if (!(x & 0x80808080u) && ((rec ^ u4) & vmask) == 0) {
        word |= ((cd | (cd >> 6) | (cd >> 12) | (cd >> 18)) & 0xFFu) << (8 * q);
        continue;
      }
#pragma unroll
      for (int b4 = 0; b4 < 4; b4++) {
        const int b = 4 * q + b4;
        if (b < nb) {
          uint32_t u = upper_ascii((four >> (8 * b4)) & 0xffu);
          word |= ((u >> 1) & 3u) << (2 * b);
          if (!is_acgt_upper(u)) {
            unsigned long long slot = atomic_add_u64(excCount, 1ull);
            if ((int64_t)slot < excCap) { excPos[slot] = (uint64_t)t * 16 + b; excByte[slot] = (uint8_t)u; }
          }
        }
      }
    }
    packed[t] = word;
  }
};

// ------------------------------------------------------------------------------------------- K1 sketch
struct SketchArgs {
  const uint32_t* packed; const int64_t* wordOff; const int32_t* len; int32_t n_seqs;
  const int64_t* chunkOff;     // n_seqs+1: first chunk of each sequence
  const int64_t* posOff;       // n_seqs+1: first slab slot of each sequence (prefix of npos)
  const uint64_t* excPos; const uint8_t* excByte; int64_t n_exc;
  int k, w, CH;
  uint32_t* slabHash; uint32_t* slabWs;   // one slot per k-mer position
  int32_t* chunkCount;
  // overflow handling of the in-register deque (see SketchChunkFn)
  unsigned long long* ovfCount; int64_t* ovfList; int64_t ovfCap;
  const int64_t* redoList;     // REDO instantiations: item i processes chunk redoList[i]
  uint32_t* gdq; int gdqCap;   // global deque storage, 2*gdqCap words per item
  // chunks the block-minimum kernel (SketchBlockMinFn) hands back to the deque kernel
  unsigned long long* bailCount; int64_t* bailList; int64_t bailCap;
};

// K16: compile-time k = 16 (the reference's default and maximum): constant shifts and masks, one full Murmur3 block
template <bool GLOBALDQ, bool K16 = false, bool REDO = GLOBALDQ>
struct SketchChunkFn {
  SketchArgs a;
  MM_HD void operator()(int64_t item) const {
    const int DQL = 32;                       // local deque capacity (power of two)
    uint32_t lh[GLOBALDQ ? 1 : DQL], lp[GLOBALDQ ? 1 : DQL];
    uint32_t *dqh, *dqp; uint32_t dmask;
    if (GLOBALDQ) { dqh = a.gdq + item * 2 * (int64_t)a.gdqCap; dqp = dqh + a.gdqCap; dmask = (uint32_t)a.gdqCap - 1; }
    else { dqh = lh; dqp = lp; dmask = DQL - 1; }
    int64_t chunk = REDO ? ldg(a.redoList + item) : item;

    int64_t sq = upper_bound_idx(a.chunkOff, (int64_t)a.n_seqs + 1, chunk) - 1;
    int32_t L = ldg(a.len + sq);
    int32_t npos = L - (K16 ? 16 : a.k) + 1;
    int32_t c0 = (int32_t)(chunk - ldg(a.chunkOff + sq)) * a.CH;
    int32_t c1 = c0 + a.CH; if (c1 > npos) c1 = npos;
    const int k = K16 ? 16 : a.k, w = a.w;
    int32_t start = c0 - 2 * (w - 1); if (start < 0) start = 0;
    int32_t shadow = c0 - (w - 1);            // steps >= shadow have an exact deque front
    const uint32_t* pw = a.packed + ldg(a.wordOff + sq);
    uint64_t gbase = (uint64_t)ldg(a.wordOff + sq) * 16;
    int64_t ec = a.n_exc ? lower_bound_idx(a.excPos, a.n_exc, gbase + (uint64_t)start) : 0;
    uint64_t nextExc = (a.n_exc && ec < a.n_exc) ? ldg(a.excPos + ec) : ~0ull;

    // sliding ASCII windows: forward k-mer bytes 0..k-1 in (f0,f1); reverse-complement k-mer in (r0,r1)
    uint64_t f0 = 0, f1 = 0, r0 = 0, r1 = 0;
    const int insShift = ((k - 1) & 7) * 8; const bool insHi = (k - 1) >= 8;
    const uint64_t rmask0 = k >= 8 ? ~0ull : ((1ull << (8 * k)) - 1);
    const uint64_t rmask1 = k >= 16 ? ~0ull : (k <= 8 ? 0ull : ((1ull << (8 * (k - 8))) - 1));
    uint32_t word = 0;
    int32_t j = start;                         // next base to push
    if ((j & 15) != 0) word = ldg(pw + (j >> 4));
    int32_t dhead = 0, dsize = 0;
    int64_t carry = -1;                        // deque-front position at the last emitting step, -1 = none
    int32_t cnt = 0;
    int64_t slab0 = ldg(a.posOff + sq) + c0;
    bool overflow = false;

    for (int32_t i = start - (k - 1); i < c1; i++) {
      // push base j = i + k - 1 (the first k-1 iterations only warm the windows up)
      if ((j & 15) == 0) word = ldg(pw + (j >> 4));
      uint32_t u = code_to_ascii((word >> (2 * (j & 15))) & 3u);
      uint32_t cu = code_to_ascii(((word >> (2 * (j & 15))) & 3u) ^ 2u);
      if (gbase + (uint64_t)j == nextExc) {
        u = ldg(a.excByte + ec); cu = u; ec++;
        nextExc = ec < a.n_exc ? ldg(a.excPos + ec) : ~0ull;
      }
      j++;
      f0 = (f0 >> 8) | (f1 << 56); f1 >>= 8;
      if (insHi) f1 |= (uint64_t)u << insShift; else f0 |= (uint64_t)u << insShift;
      r1 = ((r1 << 8) | (r0 >> 56)) & rmask1; r0 = ((r0 << 8) | cu) & rmask0;
      if (i < start) continue;

      uint32_t hf = murmur_kmer(f0, f1, k), hb = murmur_kmer(r0, r1, k);
      if (hf == hb) continue;                                  // commonFunc.hpp:130
      uint32_t cur = hf < hb ? hf : hb;
      uint32_t sbit = hf < hb ? 1u : 0u;
      while (dsize > 0 && (int32_t)(dqp[dhead & dmask] >> 1) <= i - w) { dhead++; dsize--; }        // :139
      while (dsize > 0 && dqh[(dhead + dsize - 1) & dmask] >= cur) dsize--;                          // :144
      if (dsize > (int32_t)dmask) { overflow = true; break; }
      dqh[(dhead + dsize) & dmask] = cur; dqp[(dhead + dsize) & dmask] = ((uint32_t)i << 1) | sbit; dsize++;
      if (i >= w - 1 && i >= shadow) {
        uint32_t fp = dqp[dhead & dmask];
        int64_t fpos = (int64_t)(fp >> 1);
        if (i >= c0 && fpos != carry) {                        // :157 (the wpos==0 quirk is fixed up later)
          a.slabHash[slab0 + cnt] = dqh[dhead & dmask];
          a.slabWs[slab0 + cnt] = ((uint32_t)(i - w + 1) << 1) | (fp & 1u);
          cnt++;
        }
        carry = fpos;
      }
    }
    if (overflow) {
      if (!GLOBALDQ) {
        unsigned long long s = atomic_add_u64(a.ovfCount, 1ull);
        if ((int64_t)s < a.ovfCap) a.ovfList[s] = chunk;
      }
      cnt = 0;
    }
    a.chunkCount[chunk] = cnt;
  }
};

// K1, fast path.  The deque of commonFunc.hpp:139-147 only ever answers one question: which k-mer of the window
// [i-w+1, i] has the smallest hash, the newest one on ties (`>=` pops equal hashes, :144).  Positions are cut into blocks
// of w; with S[j] = minimum of block positions j..w-1 of the PREVIOUS block (one backward pass when a block completes) and
// P = running minimum of the current block, the window ending at block position j is min(S[j+1], P) -- three compares per
// k-mer, no data-dependent loops, and a halo of one block (w positions) instead of 2(w-1).  The thread's w {hash, pos<<1|strand}
// pairs live in shared memory (A[j*stride], conflict-free).  A chunk that meets what this formulation does not cover -- a k-mer
// with hashFwd == hashBwd (skipped without expiring the deque, :130) or a non-ACGT byte -- is handed to SketchChunkFn
// through the bail list; that is 0.2 % of the chunks of a random sequence (palindromic 16-mers, 4^-8 per position).
template <bool K16>
struct SketchBlockMinFn {
  SketchArgs a;
  MM_HD void bail(int64_t chunk) const {
    unsigned long long s = atomic_add_u64(a.bailCount, 1ull);
    if ((int64_t)s < a.bailCap) a.bailList[s] = chunk;
    a.chunkCount[chunk] = 0;
  }
  MM_HD void operator()(int64_t chunk, uint2* A, int stride) const {
    int64_t sq = upper_bound_idx(a.chunkOff, (int64_t)a.n_seqs + 1, chunk) - 1;
    const int k = K16 ? 16 : a.k, w = a.w;
    int32_t npos = ldg(a.len + sq) - k + 1;
    int32_t c0 = (int32_t)(chunk - ldg(a.chunkOff + sq)) * a.CH;      // CH >= w (host)
    int32_t c1 = c0 + a.CH; if (c1 > npos) c1 = npos;
    const int32_t start = c0 > 0 ? c0 - w : 0;
    const int64_t wo = ldg(a.wordOff + sq);
    const uint32_t* pw = a.packed + wo;
    if (a.n_exc) {
      uint64_t gbase = (uint64_t)wo * 16;
      int64_t ec = lower_bound_idx(a.excPos, a.n_exc, gbase + (uint64_t)start);
      if (ec < a.n_exc && ldg(a.excPos + ec) < gbase + (uint64_t)(c1 + k - 1)) { bail(chunk); return; }
    }
    uint64_t f0 = 0, f1 = 0, r0 = 0, r1 = 0;
    const int insShift = ((k - 1) & 7) * 8; const bool insHi = (k - 1) >= 8;
    const uint64_t rmask0 = k >= 8 ? ~0ull : ((1ull << (8 * k)) - 1);
    const uint64_t rmask1 = k >= 16 ? ~0ull : (k <= 8 ? 0ull : ((1ull << (8 * (k - 8))) - 1));
    const uint2 INF = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
    bool halo = c0 > 0;
    if (!halo) for (int j = 1; j < w; j++) A[j * stride] = INF;        // no previous block
    uint2 P = INF;
    uint32_t carry = 0xFFFFFFFFu;              // pos<<1|strand of the window minimum at the previous step
    int jj = 0;
    int32_t j = start;                         // next base to push
    uint32_t word = 0;
    if ((j & 15) != 0) word = ldg(pw + (j >> 4));
    int32_t cnt = 0;
    const int64_t slab0 = ldg(a.posOff + sq) + c0;
    bool bailed = false;
    for (int32_t i = start - (k - 1); i < c1; i++) {
      if ((j & 15) == 0) word = ldg(pw + (j >> 4));
      const uint32_t code = (word >> (2 * (j & 15))) & 3u;
      const uint32_t u = code_to_ascii(code), cu = code_to_ascii(code ^ 2u);
      j++;
      f0 = (f0 >> 8) | (f1 << 56); f1 >>= 8;
      if (insHi) f1 |= (uint64_t)u << insShift; else f0 |= (uint64_t)u << insShift;
      r1 = ((r1 << 8) | (r0 >> 56)) & rmask1; r0 = ((r0 << 8) | cu) & rmask0;
      if (i < start) continue;
      const uint32_t hf = K16 ? murmur_kmer16(f0, f1) : murmur_kmer(f0, f1, k), hb = K16 ? murmur_kmer16(r0, r1) : murmur_kmer(r0, r1, k);
      if (hf == hb) { bailed = true; break; }
      const uint2 key = make_uint2(hf < hb ? hf : hb, ((uint32_t)i << 1) | (hf < hb ? 1u : 0u));
      if (key.x <= P.x) P = key;
      if (!halo && i >= w - 1) {
        uint2 F = P;
        if (jj + 1 < w) { const uint2 S = A[(jj + 1) * stride]; if (!(P.x <= S.x)) F = S; }
        if (F.y != carry) {                    // commonFunc.hpp:157 (the wpos==0 quirk is fixed up later)
          a.slabHash[slab0 + cnt] = F.x;
          a.slabWs[slab0 + cnt] = ((uint32_t)(i - w + 1) << 1) | (F.y & 1u);
          cnt++;
        }
        carry = F.y;
      }
      A[jj * stride] = key;
      if (++jj == w) {                         // block complete: suffix minima in place, newest wins ties
        jj = 0;
        uint2 run = key;
        for (int q = w - 2; q >= 0; q--) { const uint2 v = A[q * stride]; if (v.x < run.x) run = v; A[q * stride] = run; }
        if (halo) { carry = run.y; halo = false; }
        P = INF;
      }
    }
    if (bailed) { bail(chunk); return; }
    a.chunkCount[chunk] = cnt;
  }
};
#ifndef MM_HOST_EMU
template <bool K16>
__global__ void __launch_bounds__(128) sketch_blockmin_kernel(int64_t n, SketchBlockMinFn<K16> f) {
  extern __shared__ uint2 mm_k1_smem[];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) f(i, mm_k1_smem + threadIdx.x, 128);
}
// the same with the w pairs in local memory (MM_SKETCH_SCRATCH=local): no shared-memory carve-out to agree on with kernels
// of other streams that are resident at the same time
template <bool K16>
__global__ void __launch_bounds__(128) sketch_blockmin_local_kernel(int64_t n, SketchBlockMinFn<K16> f) {
  uint2 A[32];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) f(i, A, 1);
}
#endif

// gather the per-chunk slab segments into dense arrays
struct SketchCompactFn {
  const int64_t* chunkOff; const int64_t* posOff; int32_t n_seqs; int CH;
  const int32_t* chunkCount; const int64_t* chunkOutOff;
  const uint32_t* slabHash; const uint32_t* slabWs;
  uint32_t* outHash; uint32_t* outWs;
  MM_HD void operator()(int64_t chunk) const {
    int32_t n = ldg(chunkCount + chunk);
    if (n == 0) return;
    int64_t sq = upper_bound_idx(chunkOff, (int64_t)n_seqs + 1, chunk) - 1;
    int64_t src = ldg(posOff + sq) + (chunk - ldg(chunkOff + sq)) * CH;
    int64_t dst = ldg(chunkOutOff + chunk);
    for (int32_t i = 0; i < n; i++) { outHash[dst + i] = slabHash[src + i]; outWs[dst + i] = slabWs[src + i]; }
  }
};
#ifndef MM_HOST_EMU
// the same gather, one warp per 32 consecutive chunks: their output range is contiguous, so lane l takes output o0 + l,
// finds its chunk among the 32 with a shuffle search and copies one element -- coalesced writes, 60-byte runs of reads
__global__ void __launch_bounds__(256) sketch_compact_warp_kernel(SketchCompactFn f, int64_t chunks) {
  const int lane = threadIdx.x & 31;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t base = ((((int64_t)blockIdx.x * blockDim.x) + threadIdx.x) >> 5) * 32; base < chunks; base += nw * 32) {
    const int64_t c = base + lane;
    int32_t n = 0; int64_t src = 0;
    const int64_t dst = ldg(f.chunkOutOff + (c < chunks ? c : chunks));
    if (c < chunks) {
      n = ldg(f.chunkCount + c);
      if (n) {
        int64_t sq = upper_bound_idx(f.chunkOff, (int64_t)f.n_seqs + 1, c) - 1;
        src = ldg(f.posOff + sq) + (c - ldg(f.chunkOff + sq)) * f.CH;
      }
    }
    const int64_t d0 = __shfl_sync(0xffffffffu, dst, 0);
    const int32_t len = (int32_t)(__shfl_sync(0xffffffffu, dst + n, 31) - d0);
    const int32_t rel = (int32_t)(dst - d0);
    for (int32_t o0 = 0; o0 < len; o0 += 32) {
      const int32_t o = o0 + lane;
      int j = 0;                                   // last chunk of the group that starts at or before output o
#pragma unroll
      for (int step = 16; step; step >>= 1) { const int32_t r = __shfl_sync(0xffffffffu, rel, j + step); if (r <= o) j += step; }
      const int64_t sj = __shfl_sync(0xffffffffu, src, j);
      const int32_t rj = __shfl_sync(0xffffffffu, rel, j);
      if (o < len) { f.outHash[d0 + o] = f.slabHash[sj + (o - rj)]; f.outWs[d0 + o] = f.slabWs[sj + (o - rj)]; }
    }
  }
}
#endif
struct SeqOutOffFn {
  const int64_t* chunkOff; const int64_t* chunkOutOff; int64_t* seqOutOff;
  MM_HD void operator()(int64_t sq) const { seqOutOff[sq] = ldg(chunkOutOff + ldg(chunkOff + sq)); }
};

// The reference compares the deque front with minimizerIndex.back() on ALL four fields, and a front that
// has never been emitted still carries wpos 0 (commonFunc.hpp:150-162).  So while the last emitted
// minimizer is the very first one with wpos 0, a new front with the same hash and strand is NOT emitted.
// Such fronts directly follow entry 0 in the emitted list; this pass tombstones them.
struct SketchQuirkFn {
  const int64_t* seqOutOff; uint32_t* outHash; uint32_t* outWs; uint32_t* anyTomb;
  MM_HD void operator()(int64_t sq) const {
    int64_t b = ldg(seqOutOff + sq), e = ldg(seqOutOff + sq + 1);
    if (e - b < 2) return;
    uint32_t h0 = outHash[b], w0 = outWs[b];
    if ((w0 >> 1) != 0) return;
    bool any = false;
    for (int64_t i = b + 1; i < e; i++) {
      if (outHash[i] == h0 && (outWs[i] & 1u) == (w0 & 1u)) { outWs[i] = MM_TOMB; any = true; }
      else break;
    }
    if (any) atomic_or_u32(anyTomb, 1u);
  }
};
struct TombKeepFn {
  const uint32_t* ws; int32_t* keep; int64_t n;
  MM_HD void operator()(int64_t i) const { keep[i] = (i < n && ws[i] != MM_TOMB) ? 1 : 0; }
};
struct TombScatterFn {
  const uint32_t* inHash; const uint32_t* inWs; const int64_t* newIdx; uint32_t* outHash; uint32_t* outWs;
  MM_HD void operator()(int64_t i) const {
    if (inWs[i] != MM_TOMB) { int64_t d = ldg(newIdx + i); outHash[d] = inHash[i]; outWs[d] = inWs[i]; }
  }
};
struct RemapOffFn {
  const int64_t* newIdx; int64_t* off;
  MM_HD void operator()(int64_t i) const { off[i] = ldg(newIdx + off[i]); }
};

// ------------------------------------------------------------------------------------------- host side
struct SeqBatch {
  int32_t n_seqs = 0;
  int64_t total_words = 0, total_bases = 0;
  DevBuf<uint8_t> asc; DevBuf<int64_t> ascOff;
  DevBuf<uint32_t> packed; DevBuf<int64_t> wordOff; DevBuf<int32_t> len;
  DevBuf<uint64_t> excPos, excPos2; DevBuf<uint8_t> excByte, excByte2; DevBuf<unsigned long long> excCount;
  int64_t n_exc = 0;
  std::vector<int64_t> h_wordOff; std::vector<int32_t> h_len;
};

struct SketchOut {
  int64_t n_total = 0; int32_t n_seqs = 0;
  DevBuf<uint32_t> hash, ws; DevBuf<int64_t> seqOff;   // seqOff: n_seqs+1
};

struct Sketcher {
  Runtime& rt; Prims& pr;
  // scratch
  DevBuf<int64_t> chunkOff, posOff, chunkOutOff, ovfList, bailList, newIdx;
  DevBuf<int32_t> chunkCount, keep;
  DevBuf<uint32_t> slabHash, slabWs, tmpHash, tmpWs, gdq, flag;
  DevBuf<unsigned long long> ovfCount;
  double* chunkMs = nullptr;      // optional accumulator: device time of the K1 kernel proper (SketchChunkFn)
  Sketcher(Runtime& r, Prims& p) : rt(r), pr(p) {}

  // ASCII (host or device) -> packed batch.  asc_dev != nullptr: data already on the device.
  void load(SeqBatch& B, const char* seqs_host, const void* asc_dev, const int64_t* offsets, int32_t n) {
    prepare(B, offsets, n);
    const uint8_t* asc = (const uint8_t*)asc_dev;
    if (!asc) {
      B.asc.ensure((size_t)B.total_bases + 16);
      h2d(rt, B.asc.p, seqs_host + offsets[0], (size_t)B.total_bases);
      asc = B.asc.p - offsets[0];
    }
    pack_async(B, asc);
    finish_pack(B, asc);
  }
  // host-side layout of the batch + the small offset tables on the device (issued on rt.stream)
  void prepare(SeqBatch& B, const int64_t* offsets, int32_t n) {
    B.n_seqs = n;
    B.h_wordOff.assign((size_t)n + 1, 0); B.h_len.assign((size_t)n, 0);
    int64_t words = 0;
    for (int32_t i = 0; i < n; i++) {
      int64_t L = offsets[i + 1] - offsets[i];
      if (L < 0 || L > 0x7fffffff) throw Error(-34, "sequence length out of range (offset_t is int in the reference)");
      B.h_len[i] = (int32_t)L; B.h_wordOff[i] = words; words += (L + 15) / 16;
    }
    B.h_wordOff[n] = words; B.total_words = words; B.total_bases = offsets[n] - offsets[0];
    B.ascOff.ensure((size_t)n + 1); h2d(rt, B.ascOff.p, offsets, sizeof(int64_t) * ((size_t)n + 1));
    B.wordOff.ensure((size_t)n + 1); h2d(rt, B.wordOff.p, B.h_wordOff.data(), sizeof(int64_t) * ((size_t)n + 1));
    B.len.ensure((size_t)n + 1); h2d(rt, B.len.p, B.h_len.data(), sizeof(int32_t) * (size_t)n);
    B.packed.ensure((size_t)words + 4);
    B.excCount.ensure(1);
    size_t cap = B.excPos.cap; if (cap < 4096) cap = 4096;
    B.excPos.ensure(cap); B.excByte.ensure(cap);
  }
  // K0 on rt.stream; `asc` (indexed with the caller's offsets) may be device memory or mapped pinned host memory, in which
  // case the kernel pulls the bytes over PCIe itself.  No host synchronisation.
  void pack_async(SeqBatch& B, const uint8_t* asc, int ctas_per_sm = 8) {
    dev_memset(rt, B.excCount.p, 0, sizeof(unsigned long long));
    PackFn f{asc, B.ascOff.p, B.wordOff.p, B.len.p, B.n_seqs, B.packed.p, B.excCount.p, B.excPos.p, B.excByte.p, (int64_t)B.excPos.cap, B.total_words};
#ifndef MM_HOST_EMU
    // K0 of the NEXT batch is resident (one CTA per SM, PCIe-bound) while this batch's kernels run; an SM keeps the L1 /
    // shared-memory split it was given when it was last idle, so K0 asks for the all-shared split its co-residents need
    if (rt.first((const void*)foreach_kernel<PackFn>)) {
      const char* e = getenv("MM_K0_CARVEOUT"); const int pct = e ? atoi(e) : 100;
      if (pct >= 0) MM_CUDA(cudaFuncSetAttribute(foreach_kernel<PackFn>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    }
#endif
    foreach(rt, (B.total_words + 32 * PackFn::TRIPS - 1) / (32 * PackFn::TRIPS) * 32, f, 256, ctas_per_sm);
  }
  // the non-ACGT side list: count it (host sync on rt.stream), re-pack with a larger list if it overflowed, sort by position
  void finish_pack(SeqBatch& B, const uint8_t* asc) {
    unsigned long long c = 0; d2h(rt, &c, B.excCount.p, sizeof(c));
    B.n_exc = (int64_t)c;
    if (B.n_exc > (int64_t)B.excPos.cap) {      // list overflowed: rerun with room for all of them
      B.excPos.ensure((size_t)B.n_exc + 64); B.excByte.ensure((size_t)B.n_exc + 64);
      pack_async(B, asc);
      d2h(rt, &c, B.excCount.p, sizeof(c));
      B.n_exc = (int64_t)c;
    }
    if (B.n_exc > 1) {        // atomics fill the list in arbitrary order: sort by position
      B.excPos2.ensure((size_t)B.n_exc); B.excByte2.ensure((size_t)B.n_exc);
      pr.sort_pairs<uint64_t, uint8_t>(B.excPos.p, B.excPos2.p, B.excByte.p, B.excByte2.p, B.n_exc);
      d2d(rt, B.excPos.p, B.excPos2.p, sizeof(uint64_t) * (size_t)B.n_exc);
      d2d(rt, B.excByte.p, B.excByte2.p, (size_t)B.n_exc);
    }
  }

  template <class F>
  void launch_blockmin(const F& f, int64_t chunks, int w, int gridCap) {
#ifdef MM_HOST_EMU
    (void)gridCap;
    uint2 A[32];
    for (int64_t c = 0; c < chunks; c++) { (void)w; f(c, A, 1); }
    rt.launches++;
#else
    int64_t need = (chunks + 127) / 128, maxg = (int64_t)rt.sm_count * gridCap;
    int grid = (int)(need < maxg ? need : maxg);
    const char* scr = getenv("MM_SKETCH_SCRATCH");
    if (scr && !strcmp(scr, "local")) sketch_blockmin_local_kernel<<<grid, 128, 0, rt.stream>>>(chunks, f);
    else sketch_blockmin_kernel<<<grid, 128, (size_t)w * 128 * sizeof(uint2), rt.stream>>>(chunks, f);
    MM_CUDA(cudaGetLastError());
    rt.launches++;
#endif
  }

  // K1 over a packed batch.  Sequences with len < w or len < k get no minimizers (winSketch.hpp:258).
  void run(const SeqBatch& B, int k, int w, SketchOut& out) {
    int CH = 256;                            // tunables are read per call so that tests can walk through them
    { const char* e = getenv("MM_SKETCH_CH"); if (e) CH = atoi(e); if (CH < 32 || CH > 4096) CH = 256; }
    int32_t n = B.n_seqs;
    std::vector<int64_t> hChunk((size_t)n + 1), hPos((size_t)n + 1);
    int64_t chunks = 0, pos = 0;
    for (int32_t i = 0; i < n; i++) {
      hChunk[i] = chunks; hPos[i] = pos;
      int64_t L = B.h_len[i];
      if (L >= w && L >= k) { int64_t np = L - k + 1; chunks += (np + CH - 1) / CH; pos += np; }
    }
    hChunk[n] = chunks; hPos[n] = pos;
    out.n_seqs = n; out.n_total = 0;
    out.seqOff.ensure((size_t)n + 2);
    if (chunks == 0) { dev_memset(rt, out.seqOff.p, 0, sizeof(int64_t) * ((size_t)n + 1)); return; }
    chunkOff.ensure((size_t)n + 1); posOff.ensure((size_t)n + 1);
    h2d(rt, chunkOff.p, hChunk.data(), sizeof(int64_t) * ((size_t)n + 1));
    h2d(rt, posOff.p, hPos.data(), sizeof(int64_t) * ((size_t)n + 1));
    slabHash.ensure((size_t)pos); slabWs.ensure((size_t)pos);
    chunkCount.ensure((size_t)chunks + 1); chunkOutOff.ensure((size_t)chunks + 1);
    ovfCount.ensure(2); dev_memset(rt, ovfCount.p, 0, 2 * sizeof(unsigned long long));      // [0] deque overflows, [1] bails
    int64_t ovfCap = 1 << 16; ovfList.ensure((size_t)ovfCap);
    int fastPath = 1;
    { const char* e = getenv("MM_SKETCH_BLOCKMIN"); if (e) fastPath = atoi(e); }
    const bool fast = fastPath && w <= 32 && w <= CH;
    int64_t bailCap = 0;
    if (fast) {
      bailCap = chunks / 16 + 4096;          // more bails than this (low-complexity input): one deque pass over everything
      const char* e = getenv("MM_SKETCH_BAILCAP"); if (e && atoi(e) > 0) bailCap = atoi(e);
      bailList.ensure((size_t)bailCap);
    }
    SketchArgs a{B.packed.p, B.wordOff.p, B.len.p, n, chunkOff.p, posOff.p, B.excPos.p, B.excByte.p, B.n_exc,
                 k, w, CH, slabHash.p, slabWs.p, chunkCount.p, ovfCount.p, ovfList.p, ovfCap, nullptr, nullptr, 0,
                 ovfCount.p + 1, bailList.p, bailCap};
    static int gridCap = 0;
    if (!gridCap) { const char* e = getenv("MM_SKETCH_CTAS"); gridCap = e ? atoi(e) : (1 << 20); if (gridCap < 1) gridCap = 16; }   // one chunk per thread: the hardware block scheduler balances the tail
    auto deque_pass = [&]() {
      if (k == 16) foreach(rt, chunks, SketchChunkFn<false, true>{a}, 128, gridCap);
      else foreach(rt, chunks, SketchChunkFn<false>{a}, 128, gridCap);
    };
    {
      StageTimer t(rt, chunkMs);
      if (fast) {
        if (k == 16) launch_blockmin(SketchBlockMinFn<true>{a}, chunks, w, gridCap);
        else launch_blockmin(SketchBlockMinFn<false>{a}, chunks, w, gridCap);
      } else deque_pass();
    }
    unsigned long long cnts[2] = {0, 0}; d2h(rt, cnts, ovfCount.p, sizeof(cnts));
    if (getenv("MM_SKETCH_DEBUG")) fprintf(stderr, "[k1] fast=%d chunks=%lld bailed=%llu w=%d CH=%d\n", (int)fast, (long long)chunks, cnts[1], w, CH);
    if (cnts[1]) {   // chunks with a skipped k-mer or a non-ACGT byte: the reference's deque, step by step
      if ((int64_t)cnts[1] > bailCap) deque_pass();
      else {
        SketchArgs a1 = a; a1.redoList = bailList.p;
        if (k == 16) foreach(rt, (int64_t)cnts[1], SketchChunkFn<false, true, true>{a1}, 128, gridCap);
        else foreach(rt, (int64_t)cnts[1], SketchChunkFn<false, false, true>{a1}, 128, gridCap);
      }
      d2h(rt, cnts, ovfCount.p, sizeof(cnts));
    }
    unsigned long long novf = cnts[0];
    if (novf) {   // deque longer than 32 entries: replay those chunks with a w-entry deque in global memory
      if ((int64_t)novf > ovfCap) throw Error(-34, "too many deque overflows in one batch");
      int cap = 64; while (cap < w + 1) cap <<= 1;
      gdq.ensure((size_t)novf * 2 * cap);
      SketchArgs a2 = a; a2.redoList = ovfList.p; a2.gdq = gdq.p; a2.gdqCap = cap;
      foreach(rt, (int64_t)novf, SketchChunkFn<true>{a2}, 128, 16);
    }
    dev_memset(rt, chunkCount.p + chunks, 0, sizeof(int32_t));
    pr.exclusive_sum<int32_t, int64_t>(chunkCount.p, chunkOutOff.p, chunks + 1);
    int64_t total = 0; d2h(rt, &total, chunkOutOff.p + chunks, sizeof(int64_t));
    out.hash.ensure((size_t)total + 1); out.ws.ensure((size_t)total + 1);
    {
      SketchCompactFn cf{chunkOff.p, posOff.p, n, CH, chunkCount.p, chunkOutOff.p, slabHash.p, slabWs.p, out.hash.p, out.ws.p};
#ifndef MM_HOST_EMU
      const char* e = getenv("MM_SKETCH_COMPACT");
      if (!(e && !strcmp(e, "thread"))) {
        int64_t g = (chunks + 255) / 256; if (g > (int64_t)rt.sm_count * 16) g = (int64_t)rt.sm_count * 16;
        sketch_compact_warp_kernel<<<(int)g, 256, 0, rt.stream>>>(cf, chunks);
        MM_CUDA(cudaGetLastError());
        rt.launches++;
      } else
#endif
      foreach(rt, chunks, cf);
    }
    foreach(rt, (int64_t)n + 1, SeqOutOffFn{chunkOff.p, chunkOutOff.p, out.seqOff.p});
    flag.ensure(1); dev_memset(rt, flag.p, 0, sizeof(uint32_t));
    foreach(rt, n, SketchQuirkFn{out.seqOff.p, out.hash.p, out.ws.p, flag.p});
    uint32_t anyTomb = 0; d2h(rt, &anyTomb, flag.p, sizeof(uint32_t));
    if (anyTomb) {
      keep.ensure((size_t)total + 1); newIdx.ensure((size_t)total + 1);
      foreach(rt, total + 1, TombKeepFn{out.ws.p, keep.p, total});
      pr.exclusive_sum<int32_t, int64_t>(keep.p, newIdx.p, total + 1);
      tmpHash.ensure((size_t)total); tmpWs.ensure((size_t)total);
      d2d(rt, tmpHash.p, out.hash.p, sizeof(uint32_t) * (size_t)total);
      d2d(rt, tmpWs.p, out.ws.p, sizeof(uint32_t) * (size_t)total);
      foreach(rt, total, TombScatterFn{tmpHash.p, tmpWs.p, newIdx.p, out.hash.p, out.ws.p});
      foreach(rt, (int64_t)n + 1, RemapOffFn{newIdx.p, out.seqOff.p});
      d2h(rt, &total, newIdx.p + total, sizeof(int64_t));
    }
    out.n_total = total;
  }
};

}  // namespace mm
