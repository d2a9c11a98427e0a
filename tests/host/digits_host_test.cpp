// Host build of the signed-window recoding of the tabulated-sum commit (spartan_bn254_b200/csrc/digits.cuh): prints, for
// every scalar given as 64 hex digits on the command line and every window width 8..16, the signed digits, one line each:
//   <c> <W> d_0 d_1 ... d_{W-1} <carry out>
// tests/test_host_arith.py recomposes sum d_k 2^(k c) in Python and compares with the scalar.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include "../../spartan_bn254_b200/csrc/digits.cuh"

int main(int argc, char** argv) {
    for (int a = 1; a < argc; a++) {
        uint32_t l[8] = {0};
        const char* h = argv[a];
        if (strlen(h) != 64) return 2;
        for (int i = 0; i < 8; i++) {          // hex is big-endian: the last 8 digits are limb 0
            unsigned v = 0;
            sscanf(h + 56 - 8 * i, "%8x", &v);
            l[i] = v;
        }
        for (int c = 8; c <= 16; c++) {
            const int W = (254 + c) / c;
            uint32_t carry = 0;
            printf("%d %d", c, W);
            for (int k = 0; k < W; k++) {
                const uint32_t d = sbn::signed_window_digit(l, k, c, carry);
                printf(" %lld", (d >> 31) ? -(long long)(d & 0x7fffffffu) : (long long)d);
            }
            printf(" %u\n", carry);
        }
    }
    return 0;
}
