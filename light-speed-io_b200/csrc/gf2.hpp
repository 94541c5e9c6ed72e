// gf2.hpp -- host arithmetic in GF(2)[x]/P for CRC-32C (P = 0x1EDC6F41, reflected 0x82F63B78).
// Reflected representation: bit 31 of a word is the coefficient of x^0.
#pragma once

#include <cstdint>

namespace lsio {
namespace gf2 {

constexpr uint32_t kPoly = 0x82F63B78u;

inline uint32_t mul_x(uint32_t b) { return (b >> 1) ^ (kPoly & (0u - (b & 1u))); }
inline uint32_t div_x(uint32_t b) { return (b & 0x80000000u) ? (((b ^ kPoly) << 1) | 1u) : (b << 1); }
inline uint32_t multmodp(uint32_t a, uint32_t b) {
    uint32_t p = 0;
    for (uint32_t m = 0x80000000u; m; m >>= 1) {
        if (a & m) p ^= b;
        b = mul_x(b);
    }
    return p;
}
// x^n mod P; n may be negative (x is invertible because P(0) = 1).
inline uint32_t xpow(int64_t n) {
    uint32_t r = 0x80000000u;
    if (n < 0) {
        for (int64_t i = 0; i < -n; i++) r = div_x(r);
        return r;
    }
    uint32_t base = 0x40000000u;
    while (n) {
        if (n & 1) r = multmodp(r, base);
        base = multmodp(base, base);
        n >>= 1;
    }
    return r;
}
// crc32c(A || B) from the finalised crc32c(A), crc32c(B) and len(B): used when one range was
// read as several segments (reads above max_read_bytes are split).
inline uint32_t crc32c_combine(uint32_t crc_a, uint32_t crc_b, uint64_t len_b) {
    if (len_b == 0) return crc_a;
    return multmodp(xpow((int64_t)(8 * len_b)), crc_a) ^ crc_b;
}

}  // namespace gf2
}  // namespace lsio
