// Signed fixed-window recoding with a run-time window width, shared by the tabulated-sum commit (mult_kernels.cuh) and
// its host test (tests/host/digits_host_test.cpp): s = sum_k d_k 2^(k c) with |d_k| <= 2^(c-1), W = ceil(255 / c) digits.
// A raw c-bit window above 2^(c-1) borrows from the next one (d = raw - 2^c, carry 1); the top window of a scalar below
// r < 2^254 never exceeds 2^(c-1), so no carry leaves the last digit.
#pragma once
#include "fp.cuh"

namespace sbn {

// Digit k of the canonical little-endian limbs l[8].  `carry` is the borrow of digit k - 1 on entry (0 for k = 0) and of
// digit k on return.  Returns |d_k| with bit 31 set for a negative digit; 0 for a zero digit.
SBN_HD uint32_t signed_window_digit(const uint32_t* l, int k, int c, uint32_t& carry) {
    const int bit = k * c, limb = bit >> 5, off = bit & 31;
    uint32_t raw = l[limb] >> off;
    if (off + c > 32 && limb + 1 < 8) raw |= l[limb + 1] << (32 - off);
    uint32_t d = (raw & ((1u << c) - 1)) + carry;
    carry = 0;
    uint32_t neg = 0;
    if (d > (1u << (c - 1))) { d = (1u << c) - d; carry = 1; neg = 1u << 31; }
    return d ? (d | neg) : 0;
}

}  // namespace sbn
