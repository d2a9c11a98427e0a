// C++ host-side mirror of the reference's Hyrax / Pedersen interface over the libsbn254 C ABI.
// (The reference is Rust; this image has no Rust toolchain, so the compiled-language host mirror is C++.)
// Same names and argument meaning as the reference; reference preconditions (assert!/panic) surface as
// std::logic_error, library failures as std::runtime_error.  Header only; link with -lsbn254.
//
//   MultiCommitGens      commitments.rs:17-114     DotProductProofGens  nizk/mod.rs:404-415
//   PolyCommitmentGens   hyrax.rs:20-31            DensePolynomial      hyrax.rs:155-324
//   GroupElement         group.rs:20,135-175
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/sbn254.h"
#include "keccak.hpp"

namespace sbn {
namespace host {

inline void check(int status, const char* what) {
    if (status != SBN_OK) throw std::runtime_error(std::string(what) + ": " + sbn_strerror(status));
}

class Context {
public:
    explicit Context(int device = 0) { check(sbn_ctx_create(device, &h_), "sbn_ctx_create (no CPU fallback: a CUDA device is required)"); }
    ~Context() { if (h_) sbn_ctx_destroy(h_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    sbn_ctx* get() const { return h_; }
private:
    sbn_ctx* h_ = nullptr;
};

// math.rs:11-15
inline size_t log_2(size_t n) { if (n == 0) throw std::logic_error("log_2(0)"); size_t l = 0; while (n >>= 1) l++; return l; }
// hyrax.rs:371-373
inline std::pair<size_t, size_t> compute_factored_lens(size_t ell) { return {ell / 2, ell - ell / 2}; }

struct GroupElement {        // affine + infinity flag, ABI layout
    sbn_g1a p{};
    uint8_t inf = 1;
    static GroupElement generator() {     // (1, 2) in Montgomery form
        GroupElement g;
        const uint64_t one[4] = {0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL};
        const uint64_t two[4] = {0xa6ba871b8b1e1b3aULL, 0x14f1d651eb8e167bULL, 0xccdd46def0f28c58ULL, 0x1c14ef83340fbe5eULL};
        std::memcpy(g.p.x, one, 32); std::memcpy(g.p.y, two, 32); g.inf = 0;
        return g;
    }
    // group.rs:171-175.  A length mismatch yields the identity (unwrap_or_default).
    static GroupElement msm_affine(const Context& ctx, const std::vector<sbn_fr>& scalars, const std::vector<sbn_g1a>& points) {
        GroupElement r;
        if (scalars.size() != points.size()) return r;
        check(sbn_msm(ctx.get(), points.data(), nullptr, scalars.data(), scalars.size(), &r.p, &r.inf), "sbn_msm");
        return r;
    }
};

class MultiCommitGens {
public:
    size_t n = 0;
    std::vector<sbn_g1a> G;
    sbn_g1a h{};
    const Context* ctx = nullptr;

    // commitments.rs:31-62: SHAKE256(label || compress(G)) -> (n+1) x 64 B -> from_uniform_bytes (group.rs:110-132)
    static MultiCommitGens create(const Context& ctx, size_t n, const std::string& label) {
        std::vector<uint8_t> seed(label.begin(), label.end());
        uint8_t comp[32] = {1};                   // compress(G): x = 1 LE, y = 2 is the smaller root -> no flag
        seed.insert(seed.end(), comp, comp + 32);
        std::vector<uint8_t> xof(64 * (n + 1));
        keccak::shake256(seed.data(), seed.size(), xof.data(), xof.size());
        std::vector<uint64_t> canon(4 * (n + 1));
        for (size_t i = 0; i <= n; i++) uniform_bytes_to_scalar(&xof[64 * i], &canon[4 * i]);
        std::vector<sbn_fr> mont(n + 1);
        check(sbn_fr_from_canonical(ctx.get(), canon.data(), n + 1, mont.data()), "sbn_fr_from_canonical");
        std::vector<sbn_g1a> pts(n + 1);
        std::vector<uint8_t> inf(n + 1);
        GroupElement g = GroupElement::generator();
        check(sbn_g1_scalar_mul_batch(ctx.get(), &g.p, mont.data(), n + 1, pts.data(), inf.data()), "sbn_g1_scalar_mul_batch");
        MultiCommitGens out;
        out.n = n; out.ctx = &ctx;
        out.G.assign(pts.begin(), pts.begin() + n);
        out.h = pts[n];
        return out;
    }
    // commitments.rs:101-114
    static MultiCommitGens from_generators(const Context& ctx, std::vector<sbn_g1a> G, const sbn_g1a& h) {
        MultiCommitGens out; out.n = G.size(); out.G = std::move(G); out.h = h; out.ctx = &ctx; return out;
    }
    // commitments.rs:78-98
    std::pair<MultiCommitGens, MultiCommitGens> split_at(size_t mid) const {
        MultiCommitGens a = from_generators(*ctx, std::vector<sbn_g1a>(G.begin(), G.begin() + mid), h);
        MultiCommitGens b = from_generators(*ctx, std::vector<sbn_g1a>(G.begin() + mid, G.end()), h);
        return {std::move(a), std::move(b)};
    }
    // commitments.rs:64-76
    MultiCommitGens scale(const sbn_fr& s) const {
        std::vector<sbn_g1a> out(n);
        std::vector<uint8_t> inf(n);
        check(sbn_g1_scale_points(ctx->get(), G.data(), nullptr, n, &s, out.data(), inf.data()), "sbn_g1_scale_points");
        return from_generators(*ctx, std::move(out), h);
    }
    // resident copy + window tables, created on first use
    sbn_bases* device_bases() const {
        if (!bases_) {
            sbn_bases* b = nullptr;
            check(sbn_bases_create(ctx->get(), G.data(), nullptr, n, &h, &b), "sbn_bases_create");
            bases_ = std::shared_ptr<sbn_bases>(b, [](sbn_bases* p) { sbn_bases_destroy(p); });
        }
        return bases_.get();
    }
    // <[Scalar] as Commitments>::commit, commitments.rs:144-154
    GroupElement commit(const std::vector<sbn_fr>& scalars, const sbn_fr& blind) const {
        if (scalars.size() != n) throw std::logic_error("assert_eq!(gens_n.n, self.len())");
        GroupElement r;
        check(sbn_commit(ctx->get(), device_bases(), scalars.data(), n, &blind, &r.p, &r.inf), "sbn_commit");
        return r;
    }

private:
    mutable std::shared_ptr<sbn_bases> bases_;
    static bool below_r(const uint64_t v[4]) {
        static const uint64_t r[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
        for (int i = 3; i >= 0; i--) { if (v[i] < r[i]) return true; if (v[i] > r[i]) return false; }
        return false;
    }
    static void le32(const uint8_t* b, uint64_t v[4]) {
        for (int i = 0; i < 4; i++) { v[i] = 0; for (int j = 0; j < 8; j++) v[i] |= (uint64_t)b[8 * i + j] << (8 * j); }
    }
    static void uniform_bytes_to_scalar(const uint8_t chunk[64], uint64_t out[4]) {
        uint8_t hsh[32];
        keccak::sha3_256(chunk, 64, hsh);
        le32(hsh, out);
        if (below_r(out)) return;
        uint8_t buf[72];
        std::memcpy(buf, "fallback", 8);
        std::memcpy(buf + 8, chunk, 64);
        keccak::sha3_256(buf, 72, hsh);
        le32(hsh, out);
        if (below_r(out)) return;
        out[0] = 1; out[1] = out[2] = out[3] = 0;
    }
};

// nizk/mod.rs:404-415
struct DotProductProofGens {
    size_t n;
    MultiCommitGens gens_n, gens_1;
    DotProductProofGens(const Context& ctx, size_t n_, const std::string& label) : n(n_) {
        auto parts = MultiCommitGens::create(ctx, n_ + 1, label).split_at(n_);
        gens_n = std::move(parts.first);
        gens_1 = std::move(parts.second);
    }
};

// hyrax.rs:20-31
struct PolyCommitmentGens {
    DotProductProofGens gens;
    PolyCommitmentGens(const Context& ctx, size_t num_vars, const std::string& label)
        : gens(ctx, size_t(1) << compute_factored_lens(num_vars).second, label) {}
};

// hyrax.rs:38-42
struct PolyCommitment {
    std::vector<sbn_g1a> C;
    std::vector<uint8_t> inf;
};

// hyrax.rs:155-324
class DensePolynomial {
public:
    explicit DensePolynomial(std::vector<sbn_fr> Z) : Z_(std::move(Z)), len_(Z_.size()), num_vars_(len_ ? log_2(len_) : 0) {}
    size_t get_num_vars() const { return num_vars_; }
    size_t len() const { return len_; }

    // hyrax.rs:253-281
    PolyCommitment commit_inner(const std::vector<sbn_fr>& blinds, const MultiCommitGens& gens) const {
        const size_t L_size = blinds.size();
        if (L_size == 0 || (len_ / L_size) * L_size != len_) throw std::logic_error("assert_eq!(L_size * R_size, self.Z.len())");
        const size_t R_size = len_ / L_size;
        if (gens.n != R_size) throw std::logic_error("assert_eq!(gens_n.n, self.len())");
        bool zero = true;
        for (const auto& b : blinds) zero = zero && !(b.l[0] | b.l[1] | b.l[2] | b.l[3]);
        PolyCommitment out;
        out.C.resize(L_size);
        out.inf.resize(L_size);
        check(sbn_hyrax_commit(gens.ctx->get(), gens.device_bases(), Z_.data(), L_size, R_size, zero ? nullptr : blinds.data(),
                               out.C.data(), out.inf.data()), "sbn_hyrax_commit");
        return out;
    }
    // hyrax.rs:283-308: `blinds` empty = random_tape None (zero blinds); otherwise the tape's L_size scalars
    std::pair<PolyCommitment, std::vector<sbn_fr>> commit(const PolyCommitmentGens& gens, std::vector<sbn_fr> blinds = {}) const {
        if (len_ != (size_t(1) << num_vars_)) throw std::logic_error("assert_eq!(n, ell.pow2())");
        auto lens = compute_factored_lens(num_vars_);
        const size_t L_size = size_t(1) << lens.first;
        if (blinds.empty()) blinds.assign(L_size, sbn_fr{});
        if (blinds.size() != L_size) throw std::logic_error("blinds.len() == L_size");
        return {commit_inner(blinds, gens.gens.gens_n), blinds};
    }
    // hyrax.rs:311-324
    std::vector<sbn_fr> bound(const Context& ctx, const std::vector<sbn_fr>& L) const {
        auto lens = compute_factored_lens(num_vars_);
        const size_t L_size = size_t(1) << lens.first, R_size = size_t(1) << lens.second;
        if (L.size() != L_size) throw std::logic_error("L.len() == L_size");
        std::vector<sbn_fr> out(R_size);
        check(sbn_bound(ctx.get(), Z_.data(), L.data(), L_size, R_size, out.data()), "sbn_bound");
        return out;
    }

private:
    std::vector<sbn_fr> Z_;
    size_t len_, num_vars_;
};

}  // namespace host
}  // namespace sbn
