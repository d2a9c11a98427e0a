// Merlin transcript (merlin 3.0: STROBE-128 over Keccak-f[1600]) for the host mirrors of the reference's Fiat-Shamir layer
// (reference transcript.rs; third-party `merlin`).  Host only; the state is a plain 203-byte struct so bindings can own it.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include "keccak.hpp"

namespace sbn {
namespace merlin {

struct State {
    uint8_t st[200];
    uint8_t pos, pos_begin, cur_flags;
};

static constexpr int kRate = 166;
static constexpr uint8_t FLAG_I = 1, FLAG_A = 2, FLAG_C = 4, FLAG_M = 16, FLAG_K = 32;

inline void permute(State& s) {
    alignas(8) uint64_t lanes[25];
    std::memcpy(lanes, s.st, 200);          // little-endian host
    keccak::permute(lanes);
    std::memcpy(s.st, lanes, 200);
}
inline void run_f(State& s) {
    s.st[s.pos] ^= s.pos_begin;
    s.st[s.pos + 1] ^= 0x04;
    s.st[kRate + 1] ^= 0x80;
    permute(s);
    s.pos = 0;
    s.pos_begin = 0;
}
inline void absorb(State& s, const uint8_t* data, size_t n) {
    while (n) {                              // rate-sized runs: the byte-at-a-time form was a fifth of a transcript append
        const size_t take = n < (size_t)(kRate - s.pos) ? n : (size_t)(kRate - s.pos);
        uint8_t* d = s.st + s.pos;
        for (size_t i = 0; i < take; i++) d[i] ^= data[i];
        s.pos = (uint8_t)(s.pos + take);
        data += take;
        n -= take;
        if (s.pos == kRate) run_f(s);
    }
}
inline void squeeze(State& s, uint8_t* out, size_t n) {
    for (size_t i = 0; i < n; i++) {
        out[i] = s.st[s.pos];
        s.st[s.pos++] = 0;
        if (s.pos == kRate) run_f(s);
    }
}
inline void begin_op(State& s, uint8_t flags, bool more) {
    if (more) return;
    const uint8_t old_begin = s.pos_begin;
    s.pos_begin = (uint8_t)(s.pos + 1);
    s.cur_flags = flags;
    const uint8_t hdr[2] = {old_begin, flags};
    absorb(s, hdr, 2);
    if ((flags & (FLAG_C | FLAG_K)) && s.pos != 0) run_f(s);
}
inline void meta_ad(State& s, const uint8_t* d, size_t n, bool more) { begin_op(s, FLAG_M | FLAG_A, more); absorb(s, d, n); }
inline void ad(State& s, const uint8_t* d, size_t n, bool more) { begin_op(s, FLAG_A, more); absorb(s, d, n); }

inline void append_message(State& s, const uint8_t* label, size_t llen, const uint8_t* msg, size_t mlen) {
    meta_ad(s, label, llen, false);
    const uint8_t len4[4] = {(uint8_t)mlen, (uint8_t)(mlen >> 8), (uint8_t)(mlen >> 16), (uint8_t)(mlen >> 24)};
    meta_ad(s, len4, 4, true);
    ad(s, msg, mlen, false);
}
inline void challenge_bytes(State& s, const uint8_t* label, size_t llen, uint8_t* out, size_t n) {
    meta_ad(s, label, llen, false);
    const uint8_t len4[4] = {(uint8_t)n, (uint8_t)(n >> 8), (uint8_t)(n >> 16), (uint8_t)(n >> 24)};
    meta_ad(s, len4, 4, true);
    begin_op(s, FLAG_I | FLAG_A | FLAG_C, false);
    squeeze(s, out, n);
}
inline void init(State& s, const uint8_t* label, size_t llen) {
    std::memset(&s, 0, sizeof s);
    const uint8_t hdr[6] = {1, kRate + 2, 1, 0, 1, 96};
    std::memcpy(s.st, hdr, 6);
    std::memcpy(s.st + 6, "STROBEv1.0.2", 12);
    permute(s);
    meta_ad(s, (const uint8_t*)"Merlin v1.0", 11, false);
    append_message(s, (const uint8_t*)"dom-sep", 7, label, llen);
}

}  // namespace merlin
}  // namespace sbn
