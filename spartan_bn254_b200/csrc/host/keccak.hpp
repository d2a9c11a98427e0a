// FIPS-202 SHA3-256 / SHAKE256 for the host-side generator derivation (reference commitments.rs:31-62 uses
// the sha3 crate).  Header only.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>

namespace sbn {
namespace keccak {

inline uint64_t rotl(uint64_t x, int n) { return (x << n) | (x >> (64 - n)); }

inline void permute(uint64_t a[25]) {
    static const uint64_t RC[24] = {
        0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x000000000000808bULL,
        0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008aULL, 0x0000000000000088ULL,
        0x0000000080008009ULL, 0x000000008000000aULL, 0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL,
        0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
        0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
    static const int RHO[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
    for (int round = 0; round < 24; round++) {
        uint64_t c[5], d[5], b[25];
        for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
        for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ rotl(c[(x + 1) % 5], 1);
        for (int i = 0; i < 25; i++) a[i] ^= d[i % 5];
        for (int x = 0; x < 5; x++)
            for (int y = 0; y < 5; y++) {
                int i = x + 5 * y;
                int r = RHO[i];
                uint64_t v = r ? rotl(a[i], r) : a[i];
                b[y + 5 * ((2 * x + 3 * y) % 5)] = v;        // pi: (x, y) -> (y, 2x + 3y)
            }
        for (int y = 0; y < 5; y++)
            for (int x = 0; x < 5; x++) a[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
        a[0] ^= RC[round];
    }
}

inline void sponge(const uint8_t* in, size_t len, size_t rate, uint8_t dom, uint8_t* out, size_t outlen) {
    uint64_t a[25] = {0};
    uint8_t block[200];
    auto absorb_block = [&](const uint8_t* p) {
        for (size_t i = 0; i < rate / 8; i++) {
            uint64_t w = 0;
            for (int j = 0; j < 8; j++) w |= (uint64_t)p[8 * i + j] << (8 * j);
            a[i] ^= w;
        }
        permute(a);
    };
    while (len >= rate) { absorb_block(in); in += rate; len -= rate; }
    std::memset(block, 0, sizeof block);
    std::memcpy(block, in, len);
    block[len] ^= dom;
    block[rate - 1] ^= 0x80;
    absorb_block(block);
    while (outlen) {
        size_t k = outlen < rate ? outlen : rate;
        for (size_t i = 0; i < k; i++) out[i] = (uint8_t)(a[i / 8] >> (8 * (i % 8)));
        out += k; outlen -= k;
        if (outlen) permute(a);
    }
}
inline void sha3_256(const uint8_t* in, size_t len, uint8_t out[32]) { sponge(in, len, 136, 0x06, out, 32); }
inline void shake256(const uint8_t* in, size_t len, uint8_t* out, size_t outlen) { sponge(in, len, 136, 0x1f, out, outlen); }

}  // namespace keccak
}  // namespace sbn
