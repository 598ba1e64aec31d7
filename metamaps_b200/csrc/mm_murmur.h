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

// k = 16 on the device: the same arithmetic on explicit 32-bit halves.  K1 is bound by the half-rate IMAD pipe (ncu: fmaheavy
// 88 % busy), and left to itself the compiler folds each rotate-after-multiply into a second wide multiply and spends IMADs
// on moves and adds; written out, a 64 x 64 -> 64 multiply is three IMADs, a rotate two funnel shifts (ALU pipe).
#if defined(__CUDA_ARCH__)
struct U64h { uint32_t lo, hi; };
__device__ __forceinline__ U64h mulc64(U64h a, uint32_t clo, uint32_t chi) {
  U64h r; uint32_t t;
  asm("{ .reg .u64 w; mul.wide.u32 w, %3, %4; mov.b64 {%0, %2}, w; mad.lo.u32 %1, %3, %5, %2; mad.lo.u32 %1, %6, %4, %1; }"
      : "=&r"(r.lo), "=&r"(r.hi), "=&r"(t) : "r"(a.lo), "r"(clo), "r"(chi), "r"(a.hi));
  return r;
}
__device__ __forceinline__ U64h rotl_lt32(U64h a, int n) {            // 0 < n < 32
  U64h r; r.hi = __funnelshift_l(a.lo, a.hi, n); r.lo = __funnelshift_l(a.hi, a.lo, n); return r;
}
__device__ __forceinline__ U64h add64h(U64h a, U64h b) {
  U64h r; asm("add.cc.u32 %0, %2, %4; addc.u32 %1, %3, %5;" : "=&r"(r.lo), "=&r"(r.hi) : "r"(a.lo), "r"(a.hi), "r"(b.lo), "r"(b.hi)); return r;
}
__device__ __forceinline__ U64h mul5add(U64h a, uint32_t c) {         // a * 5 + c
  U64h r; uint32_t t;
  asm("{ .reg .u64 w, cc; mov.b64 cc, {%4, %5}; mad.wide.u32 w, %3, 5, cc; mov.b64 {%0, %2}, w; mad.lo.u32 %1, %6, 5, %2; }"
      : "=&r"(r.lo), "=&r"(r.hi), "=&r"(t) : "r"(a.lo), "r"(c), "r"(0u), "r"(a.hi));
  return r;
}
__device__ __forceinline__ U64h fmix64h_but_last(U64h k) {            // fmix64 without its final k ^= k >> 33
  k.lo ^= k.hi >> 1; k = mulc64(k, 0xed558ccdu, 0xff51afd7u);
  k.lo ^= k.hi >> 1; k = mulc64(k, 0x1a85ec53u, 0xc4ceb9feu);
  return k;
}
#endif
MM_HD uint32_t murmur_kmer16(uint64_t b0, uint64_t b1) {
#if defined(__CUDA_ARCH__)
  const uint32_t c1l = 0x114253d5u, c1h = 0x87c37b91u, c2l = 0x2745937fu, c2h = 0x4cf5ad43u;
  U64h k1{(uint32_t)b0, (uint32_t)(b0 >> 32)}, k2{(uint32_t)b1, (uint32_t)(b1 >> 32)};
  k1 = mulc64(k1, c1l, c1h); k1 = rotl_lt32(k1, 31); k1 = mulc64(k1, c2l, c2h);
  U64h h1{42u ^ k1.lo, k1.hi};
  h1 = rotl_lt32(h1, 27); h1 = add64h(h1, U64h{42u, 0u}); h1 = mul5add(h1, 0x52dce729u);
  k2 = mulc64(k2, c2l, c2h); k2 = rotl_lt32(U64h{k2.hi, k2.lo}, 1); k2 = mulc64(k2, c1l, c1h);       // rotl 33 = swap halves, rotl 1
  U64h h2{42u ^ k2.lo, k2.hi};
  h2 = rotl_lt32(h2, 31); h2 = add64h(h2, h1); h2 = mul5add(h2, 0x38495ab5u);
  h1.lo ^= 16u; h2.lo ^= 16u;
  h1 = add64h(h1, h2); h2 = add64h(h2, h1);
  h1 = fmix64h_but_last(h1); h2 = fmix64h_but_last(h2);
  return (h1.lo ^ (h1.hi >> 1)) + (h2.lo ^ (h2.hi >> 1));
#else
  return murmur_kmer(b0, b1, 16);
#endif
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
