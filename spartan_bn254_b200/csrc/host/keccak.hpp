// FIPS-202 SHA3-256 / SHAKE256 for the host-side generator derivation (reference commitments.rs:31-62 uses
// the sha3 crate).  Header only.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>

namespace sbn {
namespace keccak {

inline uint64_t rotl(uint64_t x, int n) { return (x << n) | (x >> (64 - n)); }

// Keccak-f[1600], 24 rounds, with the 25 lanes in local variables and theta / rho / pi / chi written out (the index
// arithmetic of the textbook loops -- (2x + 3y) % 5 per lane per round -- made a permutation ~2 us; a proof runs ~7 000 of them
// on the Fiat-Shamir critical path).  Lane a[x + 5 y] is named by its row letter (b g k m s for y = 0..4) and column vowel
// (a e i o u for x = 0..4), as in the reference implementations.
inline void permute(uint64_t a[25]) {
    static const uint64_t RC[24] = {
        0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x000000000000808bULL,
        0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008aULL, 0x0000000000000088ULL,
        0x0000000080008009ULL, 0x000000008000000aULL, 0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL,
        0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
        0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
    uint64_t Aba = a[0], Abe = a[1], Abi = a[2], Abo = a[3], Abu = a[4];
    uint64_t Aga = a[5], Age = a[6], Agi = a[7], Ago = a[8], Agu = a[9];
    uint64_t Aka = a[10], Ake = a[11], Aki = a[12], Ako = a[13], Aku = a[14];
    uint64_t Ama = a[15], Ame = a[16], Ami = a[17], Amo = a[18], Amu = a[19];
    uint64_t Asa = a[20], Ase = a[21], Asi = a[22], Aso = a[23], Asu = a[24];
    for (int round = 0; round < 24; round++) {
        // theta
        const uint64_t Ca = Aba ^ Aga ^ Aka ^ Ama ^ Asa, Ce = Abe ^ Age ^ Ake ^ Ame ^ Ase, Ci = Abi ^ Agi ^ Aki ^ Ami ^ Asi,
                       Co = Abo ^ Ago ^ Ako ^ Amo ^ Aso, Cu = Abu ^ Agu ^ Aku ^ Amu ^ Asu;
        const uint64_t Da = Cu ^ rotl(Ce, 1), De = Ca ^ rotl(Ci, 1), Di = Ce ^ rotl(Co, 1), Do = Ci ^ rotl(Cu, 1), Du = Co ^ rotl(Ca, 1);
        // rho + pi: B[y][(2x + 3y) % 5] = rotl(A[x][y] ^ D[x], r[x][y]); then chi row by row
        uint64_t B0, B1, B2, B3, B4;
        B0 = Aba ^ Da; B1 = rotl(Age ^ De, 44); B2 = rotl(Aki ^ Di, 43); B3 = rotl(Amo ^ Do, 21); B4 = rotl(Asu ^ Du, 14);
        const uint64_t Eba = B0 ^ (~B1 & B2) ^ RC[round], Ebe = B1 ^ (~B2 & B3), Ebi = B2 ^ (~B3 & B4), Ebo = B3 ^ (~B4 & B0), Ebu = B4 ^ (~B0 & B1);
        B0 = rotl(Abo ^ Do, 28); B1 = rotl(Agu ^ Du, 20); B2 = rotl(Aka ^ Da, 3); B3 = rotl(Ame ^ De, 45); B4 = rotl(Asi ^ Di, 61);
        const uint64_t Ega = B0 ^ (~B1 & B2), Ege = B1 ^ (~B2 & B3), Egi = B2 ^ (~B3 & B4), Ego = B3 ^ (~B4 & B0), Egu = B4 ^ (~B0 & B1);
        B0 = rotl(Abe ^ De, 1); B1 = rotl(Agi ^ Di, 6); B2 = rotl(Ako ^ Do, 25); B3 = rotl(Amu ^ Du, 8); B4 = rotl(Asa ^ Da, 18);
        const uint64_t Eka = B0 ^ (~B1 & B2), Eke = B1 ^ (~B2 & B3), Eki = B2 ^ (~B3 & B4), Eko = B3 ^ (~B4 & B0), Eku = B4 ^ (~B0 & B1);
        B0 = rotl(Abu ^ Du, 27); B1 = rotl(Aga ^ Da, 36); B2 = rotl(Ake ^ De, 10); B3 = rotl(Ami ^ Di, 15); B4 = rotl(Aso ^ Do, 56);
        const uint64_t Ema = B0 ^ (~B1 & B2), Eme = B1 ^ (~B2 & B3), Emi = B2 ^ (~B3 & B4), Emo = B3 ^ (~B4 & B0), Emu = B4 ^ (~B0 & B1);
        B0 = rotl(Abi ^ Di, 62); B1 = rotl(Ago ^ Do, 55); B2 = rotl(Aku ^ Du, 39); B3 = rotl(Ama ^ Da, 41); B4 = rotl(Ase ^ De, 2);
        const uint64_t Esa = B0 ^ (~B1 & B2), Ese = B1 ^ (~B2 & B3), Esi = B2 ^ (~B3 & B4), Eso = B3 ^ (~B4 & B0), Esu = B4 ^ (~B0 & B1);
        Aba = Eba; Abe = Ebe; Abi = Ebi; Abo = Ebo; Abu = Ebu;
        Aga = Ega; Age = Ege; Agi = Egi; Ago = Ego; Agu = Egu;
        Aka = Eka; Ake = Eke; Aki = Eki; Ako = Eko; Aku = Eku;
        Ama = Ema; Ame = Eme; Ami = Emi; Amo = Emo; Amu = Emu;
        Asa = Esa; Ase = Ese; Asi = Esi; Aso = Eso; Asu = Esu;
    }
    a[0] = Aba; a[1] = Abe; a[2] = Abi; a[3] = Abo; a[4] = Abu;
    a[5] = Aga; a[6] = Age; a[7] = Agi; a[8] = Ago; a[9] = Agu;
    a[10] = Aka; a[11] = Ake; a[12] = Aki; a[13] = Ako; a[14] = Aku;
    a[15] = Ama; a[16] = Ame; a[17] = Ami; a[18] = Amo; a[19] = Amu;
    a[20] = Asa; a[21] = Ase; a[22] = Asi; a[23] = Aso; a[24] = Asu;
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
