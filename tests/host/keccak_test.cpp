// CPU test of the product's FIPS-202 implementation against known answers.
#include <cstdio>
#include <cstring>
#include <string>
#include "../../spartan_bn254_b200/csrc/host/keccak.hpp"
static std::string hex(const uint8_t* p, size_t n) { static const char* d = "0123456789abcdef"; std::string s; for (size_t i = 0; i < n; i++) { s += d[p[i] >> 4]; s += d[p[i] & 15]; } return s; }
int main(int argc, char** argv) {
    // prints sha3_256(msg) and shake256(msg, 64 B) for msg = argv[1] repeated argv[2] times
    std::string msg;
    int rep = argc > 2 ? atoi(argv[2]) : 1;
    for (int i = 0; i < rep; i++) msg += argc > 1 ? argv[1] : "";
    uint8_t h[32], x[200];
    sbn::keccak::sha3_256((const uint8_t*)msg.data(), msg.size(), h);
    sbn::keccak::shake256((const uint8_t*)msg.data(), msg.size(), x, 200);
    printf("%s\n%s\n", hex(h, 32).c_str(), hex(x, 200).c_str());
    return 0;
}
