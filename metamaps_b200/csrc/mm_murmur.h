// mm_murmur.h -- k-mer hashing, bit-exact with the reference.
//
// Reference: CommonFunc::getHash (src/map/include/commonFunc.hpp:71-81) = the first 4 bytes of
// MurmurHash3_x64_128(kmer, k, seed 42) (src/common/murmur3.h:226-303; public-domain algorithm by
// A. Appleby), i.e. the low 32 bits of h1.  The reference restricts k <= 16 (parseCmdArgs.hpp:62), so a
// k-mer is at most one 16-byte block: the k ASCII bytes live in two little-endian 64-bit words
// (b0 = bytes 0..7, b1 = bytes 8..15, unused bytes zero).
#pragma once
#include "mm_platform.h"

namespace mm {

MM_HD uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
MM_HD uint64_t fmix64(uint64_t k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdULL;
  k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL;
  k ^= k >> 33;
  return k;
}

// low 32 bits of MurmurHash3_x64_128 over the k bytes held in (b0,b1), seed 42
MM_HD uint32_t murmur_kmer(uint64_t b0, uint64_t b1, int k) {
  const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
  uint64_t h1 = 42, h2 = 42;
  if (k == 16) {                       // one full block, no tail (murmur3.h:243-253)
    uint64_t k1 = b0, k2 = b1;
    k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
    h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
    k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2;
    h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
  } else {                             // tail only (murmur3.h:258-284)
    if (k > 8) { uint64_t k2 = b1; k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2; }
    if (k > 0) { uint64_t k1 = b0; k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1; }
  }
  h1 ^= (uint64_t)k; h2 ^= (uint64_t)k;
  h1 += h2; h2 += h1;
  h1 = fmix64(h1); h2 = fmix64(h2);
  h1 += h2;
  return (uint32_t)h1;
}

// 2-bit code <-> ASCII.  code = (upper(c) >> 1) & 3 : A=0 C=1 T=2 G=3 ; complement = code ^ 2.
MM_HD uint32_t code_to_ascii(uint32_t c) { return (0x47544341u >> (8 * c)) & 0xffu; }
MM_HD uint32_t upper_ascii(uint32_t c) { return (c > 96 && c < 123) ? c - 32 : c; }   // commonFunc.hpp:57-66
MM_HD bool is_acgt_upper(uint32_t u) { return u == 'A' || u == 'C' || u == 'G' || u == 'T'; }
// reverseComplement leaves every byte other than A,C,G,T untouched (commonFunc.hpp:44-51)
MM_HD uint32_t comp_ascii(uint32_t u) {
  return u == 'A' ? 'T' : u == 'C' ? 'G' : u == 'G' ? 'C' : u == 'T' ? 'A' : u;
}

}  // namespace mm
