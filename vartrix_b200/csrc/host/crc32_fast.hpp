// crc32_fast.hpp -- CRC-32 (IEEE 802.3, reflected, as in gzip / BGZF trailers) with carry-less multiplication.
//
// htslib verifies the CRC of every BGZF block it inflates, and so does this host (bam_reader.hpp); with zlib's table-driven
// crc32 that check is 11 % of a staging thread's time.  This is the folding scheme of Gopal et al., "Fast CRC Computation for
// Generic Polynomials Using PCLMULQDQ Instruction" (Intel, 2009), for the reflected polynomial 0x1DB710641: four 128-bit lanes
// are folded 512 bits ahead per step, then into one lane, then reduced to 64 and (Barrett) to 32 bits.  The constants are
// x^(512+64), x^512, x^(128+64), x^128, x^96 mod P and P, floor(x^64 / P) in the bit-reflected representation.
// Chosen at run time (__builtin_cpu_supports); everything else falls through to zlib.  tests/test_host_io_cpu.py checks it
// against zlib on random buffers of every length class.
#pragma once
#include <zlib.h>

#include <cstddef>
#include <cstdint>

#if defined(__x86_64__) && (defined(__GNUC__) || defined(__clang__))
#include <immintrin.h>
#define VTX_CRC_CLMUL 1
#else
#define VTX_CRC_CLMUL 0
#endif

namespace vtx_crc {

#if VTX_CRC_CLMUL
#define VTX_CLMUL_FN __attribute__((target("pclmul,sse4.1"))) inline
VTX_CLMUL_FN __m128i ld(const uint8_t* q) { return _mm_loadu_si128(reinterpret_cast<const __m128i*>(q)); }
VTX_CLMUL_FN __m128i fold(__m128i acc, __m128i k, __m128i next)            // acc moved ahead by the distance k encodes, plus the data there
{
    return _mm_xor_si128(_mm_xor_si128(_mm_clmulepi64_si128(acc, k, 0x00), _mm_clmulepi64_si128(acc, k, 0x11)), next);
}
// raw register in, raw register out (no 0xFFFFFFFF conditioning); len >= 64 and a multiple of 16
VTX_CLMUL_FN uint32_t fold_clmul(const uint8_t* p, size_t len, uint32_t reg)
{
    const __m128i k_512 = _mm_set_epi64x(0x01c6e41596, 0x0154442bd4);      // lo: x^(512+64), hi: x^512
    const __m128i k_128 = _mm_set_epi64x(0x00ccaa009e, 0x01751997d0);      // lo: x^(128+64), hi: x^128
    const __m128i k_96 = _mm_set_epi64x(0, 0x0163cd6124);
    const __m128i k_poly = _mm_set_epi64x(0x01f7011641, 0x01db710641);     // lo: P, hi: mu
    __m128i a = _mm_xor_si128(ld(p), _mm_cvtsi32_si128(int(reg))), b = ld(p + 16), c = ld(p + 32), d = ld(p + 48);
    p += 64; len -= 64;
    while (len >= 64) {
        a = fold(a, k_512, ld(p)); b = fold(b, k_512, ld(p + 16)); c = fold(c, k_512, ld(p + 32)); d = fold(d, k_512, ld(p + 48));
        p += 64; len -= 64;
    }
    a = fold(a, k_128, b); a = fold(a, k_128, c); a = fold(a, k_128, d);
    while (len >= 16) { a = fold(a, k_128, ld(p)); p += 16; len -= 16; }
    // 128 -> 64 bits
    const __m128i mask32 = _mm_setr_epi32(~0, 0, ~0, 0);
    __m128i t = _mm_clmulepi64_si128(a, k_128, 0x10);                        // low half times x^128
    a = _mm_xor_si128(_mm_srli_si128(a, 8), t);
    t = _mm_srli_si128(a, 4);
    a = _mm_xor_si128(_mm_clmulepi64_si128(_mm_and_si128(a, mask32), k_96, 0x00), t);
    // Barrett reduction to 32 bits
    t = _mm_clmulepi64_si128(_mm_and_si128(a, mask32), k_poly, 0x10);
    t = _mm_clmulepi64_si128(_mm_and_si128(t, mask32), k_poly, 0x00);
    return uint32_t(_mm_extract_epi32(_mm_xor_si128(a, t), 1));
}
inline bool have_clmul()
{
    static const bool ok = __builtin_cpu_supports("pclmul") && __builtin_cpu_supports("sse4.1");
    return ok;
}
#endif

// crc32 of p[0, len), same value as zlib's crc32(crc32(0, NULL, 0), p, len)
inline uint32_t crc32_of(const uint8_t* p, size_t len)
{
#if VTX_CRC_CLMUL
    if (len >= 64 && have_clmul()) {
        const size_t body = len & ~size_t(15);
        const uint32_t reg = fold_clmul(p, body, 0xFFFFFFFFu);               // zlib's interface hides the conditioning; here it is explicit
        const uint32_t partial = ~reg;                                       // = crc32 of the first `body` bytes
        return body == len ? partial : uint32_t(crc32(partial, p + body, uInt(len - body)));
    }
#endif
    return uint32_t(crc32(crc32(0L, Z_NULL, 0), p, uInt(len)));
}

}  // namespace vtx_crc
