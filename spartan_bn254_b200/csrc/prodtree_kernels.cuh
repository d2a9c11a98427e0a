// Kernels of the product-layer argument (sm_100a): product circuits and the batched cubic sumcheck.
//
//   reference product_tree.rs:21-57     ProductCircuit::compute_layer / new   -> k_product_layer
//   reference hyrax.rs:355-369          EqPolynomial::evals                   -> k_eq_expand
//   reference sumcheck.rs:201-271       round evaluations of A * B * C at 0, 2, 3, one triple per instance
//                                                                             -> k_cubic_eval_batched
//   reference sumcheck.rs:293-306       bound_poly_var_top on every table     -> k_bind_top_batched
//
// Pure Fr arithmetic over tables that halve every round.  The evaluation does six Montgomery products per 192 B it
// streams (32 B per product; the multiplier sustains 6.9e10 products/s = 2.2 TB/s of operands, a third of HBM), so its
// first rounds are bound by the IMAD pipe, not by HBM (1.41 ms for 12 x 2^21 triples = 78 % of that bound, 1.5 TB/s);
// the bind (one product per 96 B) is HBM-bound; the last rounds of every layer are launch-latency-bound.
#pragma once
#include "opening_kernels.cuh"

namespace sbn {

// next[i] = cur[i] * cur[half + i]: a layer's (left | right) halves multiplied element-wise give the next layer, which
// is again stored as (left | right) (product_tree.rs:26-33 splits the products at len / 4).
__global__ void k_product_layer(const Fr* __restrict__ cur, size_t half, Fr* __restrict__ next) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= half) return;
    store_fr(next + i, fp_mul(load_fr(cur + i), load_fr(cur + half + i)));
}

// One doubling step of EqPolynomial::evals: out[2i + 1] = in[i] * r_j, out[2i] = in[i] - out[2i + 1].
__global__ void k_eq_expand(const Fr* __restrict__ in, size_t size, const Fr* __restrict__ rj, Fr* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= size) return;
    const Fr s = load_fr(in + i);
    const Fr hi = fp_mul(s, load_fr(rj));
    store_fr(out + 2 * i + 1, hi);
    store_fr(out + 2 * i, fp_sub(s, hi));
}

struct CubicTriple { const Fr *A, *B, *C; };

// partial[(inst * 3 + e) * gridDim.x + block] = this block's share of sum_i A_t[i] B_t[i] C_t[i] at t = 0, 2, 3 with
// X_t = lo + t (hi - lo); blockIdx.y = instance.
__global__ void __launch_bounds__(kDotThreads)
k_cubic_eval_batched(const CubicTriple* __restrict__ triples, size_t half, Fr* __restrict__ partial) {
    __shared__ Fr sm[kDotThreads];
    const CubicTriple t = triples[blockIdx.y];
    Fr e0 = Fr::zero(), e2 = Fr::zero(), e3 = Fr::zero();
    for (size_t i = (size_t)blockIdx.x * kDotThreads + threadIdx.x; i < half; i += (size_t)gridDim.x * kDotThreads) {
        const Fr a0 = load_fr(t.A + i), a1 = load_fr(t.A + half + i);
        const Fr b0 = load_fr(t.B + i), b1 = load_fr(t.B + half + i);
        const Fr c0 = load_fr(t.C + i), c1 = load_fr(t.C + half + i);
        const Fr a2 = fp_sub(fp_add(a1, a1), a0), b2 = fp_sub(fp_add(b1, b1), b0), c2 = fp_sub(fp_add(c1, c1), c0);
        const Fr a3 = fp_sub(fp_add(a2, a1), a0), b3 = fp_sub(fp_add(b2, b1), b0), c3 = fp_sub(fp_add(c2, c1), c0);
        // six products, inlined: three independent chains of two give the scheduler something to overlap
        e0 = fp_add(e0, fp_mul(fp_mul(a0, b0), c0));
        e2 = fp_add(e2, fp_mul(fp_mul(a2, b2), c2));
        e3 = fp_add(e3, fp_mul(fp_mul(a3, b3), c3));
    }
    Fr* out = partial + (size_t)blockIdx.y * 3 * gridDim.x;
    e0 = block_sum_fr(e0, sm, kDotThreads);
    if (threadIdx.x == 0) store_fr(out + blockIdx.x, e0);
    __syncthreads();
    e2 = block_sum_fr(e2, sm, kDotThreads);
    if (threadIdx.x == 0) store_fr(out + gridDim.x + blockIdx.x, e2);
    __syncthreads();
    e3 = block_sum_fr(e3, sm, kDotThreads);
    if (threadIdx.x == 0) store_fr(out + 2 * gridDim.x + blockIdx.x, e3);
}

// Round j >= 1 of a batched cubic sumcheck in ONE kernel: the bind of round j - 1 and the evaluation of round j.  A round is
// Fiat-Shamir-sequential -- evaluation, host (transcript, challenge), bind, next evaluation -- and on that chain every launch
// costs its launch-to-start latency as well as its run time: with the bind folded into the next evaluation a round is one
// kernel instead of two.  h = the evaluation's half length; the tables still have 4 h entries (the previous challenge r is
// pending).  Thread i binds entries i and i + h of every table of its instance,
//     T'[i] = T[i] + r (T[i + 2h] - T[i]),   T'[i + h] = T[i + h] + r (T[i + 3h] - T[i + h]),
// stores them in place (a thread writes only what it alone reads; the upper half is never written) and evaluates the cubic on
// (T'[i], T'[i + h]).  The eq table shared by the parallel instances (inst < P) is read by all of them, so its bound copy
// goes to the other ping-pong buffer, written by instance 0 only; a sequential instance owns its C table and binds it in place.
__global__ void __launch_bounds__(kDotThreads)
k_bind_eval_batched(const CubicTriple* __restrict__ triples, int P, const Fr* __restrict__ eq_in, Fr* __restrict__ eq_out, size_t h,
                    const Fr r, Fr* __restrict__ partial) {
    __shared__ Fr sm[kDotThreads];
    const int inst = blockIdx.y;
    const CubicTriple t = triples[inst];
    const bool shared_c = inst < P;
    Fr* A = const_cast<Fr*>(t.A);
    Fr* B = const_cast<Fr*>(t.B);
    const Fr* Cin = shared_c ? eq_in : t.C;
    Fr* Cout = shared_c ? eq_out : const_cast<Fr*>(t.C);
    const bool write_c = !shared_c || inst == 0;
    const size_t hp = 2 * h;
    Fr e0 = Fr::zero(), e2 = Fr::zero(), e3 = Fr::zero();
    for (size_t i = (size_t)blockIdx.x * kDotThreads + threadIdx.x; i < h; i += (size_t)gridDim.x * kDotThreads) {
        // plain loads: these tables are rewritten by this very kernel (by this very thread), not read-only data
        auto ld = [](const Fr* p) {
            const uint4* q = reinterpret_cast<const uint4*>(p);
            const uint4 u = q[0], v = q[1];
            Fr x;
            x.l[0] = u.x; x.l[1] = u.y; x.l[2] = u.z; x.l[3] = u.w; x.l[4] = v.x; x.l[5] = v.y; x.l[6] = v.z; x.l[7] = v.w;
            return x;
        };
        auto bound = [&](const Fr* T, size_t k) {
            const Fr lo = ld(T + k), hi = ld(T + k + hp);
            return fp_add(lo, fp_mul(r, fp_sub(hi, lo)));
        };
        const Fr a0 = bound(A, i), a1 = bound(A, i + h);
        const Fr b0 = bound(B, i), b1 = bound(B, i + h);
        const Fr c0 = bound(Cin, i), c1 = bound(Cin, i + h);
        store_fr(A + i, a0); store_fr(A + i + h, a1);
        store_fr(B + i, b0); store_fr(B + i + h, b1);
        if (write_c) { store_fr(Cout + i, c0); store_fr(Cout + i + h, c1); }
        const Fr a2 = fp_sub(fp_add(a1, a1), a0), b2 = fp_sub(fp_add(b1, b1), b0), c2 = fp_sub(fp_add(c1, c1), c0);
        const Fr a3 = fp_sub(fp_add(a2, a1), a0), b3 = fp_sub(fp_add(b2, b1), b0), c3 = fp_sub(fp_add(c2, c1), c0);
        e0 = fp_add(e0, fp_mul(fp_mul(a0, b0), c0));
        e2 = fp_add(e2, fp_mul(fp_mul(a2, b2), c2));
        e3 = fp_add(e3, fp_mul(fp_mul(a3, b3), c3));
    }
    Fr* out = partial + (size_t)blockIdx.y * 3 * gridDim.x;
    e0 = block_sum_fr(e0, sm, kDotThreads);
    if (threadIdx.x == 0) store_fr(out + blockIdx.x, e0);
    __syncthreads();
    e2 = block_sum_fr(e2, sm, kDotThreads);
    if (threadIdx.x == 0) store_fr(out + gridDim.x + blockIdx.x, e2);
    __syncthreads();
    e3 = block_sum_fr(e3, sm, kDotThreads);
    if (threadIdx.x == 0) store_fr(out + 2 * gridDim.x + blockIdx.x, e3);
}

// T[i] <- T[i] + r (T[half + i] - T[i]) for every table of the list; blockIdx.y = table.
__global__ void k_bind_top_batched(Fr* const* __restrict__ tables, size_t half, const Fr* __restrict__ r) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= half) return;
    Fr* T = tables[blockIdx.y];
    const Fr lo = load_fr(T + i), hi = load_fr(T + half + i);
    store_fr(T + i, fp_add(lo, fp_mul(load_fr(r), fp_sub(hi, lo))));
}

// The same with the challenge passed by value (the in-library round loop has it on the host: no 32-byte copy per round)
__global__ void k_bind_top_batched_v(Fr* const* __restrict__ tables, size_t half, const Fr r) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= half) return;
    Fr* T = tables[blockIdx.y];
    const Fr lo = load_fr(T + i), hi = load_fr(T + half + i);
    store_fr(T + i, fp_add(lo, fp_mul(r, fp_sub(hi, lo))));
}

// out[t] = tables[t][0]: the final values of a sumcheck whose tables have been bound down to one entry (sumcheck.rs:309-327)
__global__ void k_gather_first(Fr* const* __restrict__ tables, int ntables, Fr* __restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < ntables) store_fr(out + t, load_fr(tables[t]));
}

// Derefs (sparse_mlpoly_full.rs:245-257 deref_mem, :292-297 Derefs::new, hyrax.rs:237-247 merge): segment s < batch gathers
// mem_rx[row_addr[s][i]], segment batch + s gathers mem_ry[col_addr[s][i]]; the tail up to the next power of two is zero.
__global__ void k_derefs_gather(const Fr* __restrict__ mem_rx, const Fr* __restrict__ mem_ry, const uint32_t* __restrict__ row_addr,
                                const uint32_t* __restrict__ col_addr, size_t batch, size_t N, size_t total, Fr* __restrict__ Z) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const size_t seg = i / N, k = i - seg * N;
    Fr v = Fr::zero();
    if (seg < batch) v = load_fr(mem_rx + row_addr[seg * N + k]);
    else if (seg < 2 * batch) v = load_fr(mem_ry + col_addr[(seg - batch) * N + k]);
    store_fr(Z + i, v);
}

// Hash layer of the memory-checking argument (sparse_mlpoly_full.rs:745-798 build_hash_layer):
//     h(addr, val, ts) = ts * r_hash^2 + val * r_hash + addr - r_multiset_check .
// Addresses and timestamps are small integers kept as uint32; they enter Montgomery form through one product each:
// montmul(c, R^2) = c R for an address, montmul(c, r_hash^2 R^2) = c r_hash^2 R for a timestamp (`rh2_R2`).
struct HashParams { Fr rh, rh2, rh2_R2, r_ms; };

__device__ __forceinline__ Fr fr_R2() {
    Fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = FrParams::R2(i);
    return r;
}
__device__ __forceinline__ Fr fr_from_u32(uint32_t c) {
    Fr v = Fr::zero();
    v.l[0] = c;
    return v;
}

// init[i] = h(i, mem[i], 0), audit[i] = h(i, mem[i], audit_ts[i])                       (:758-770)
__global__ void k_hash_mem(const Fr* __restrict__ mem, const uint32_t* __restrict__ audit_ts, size_t M, HashParams hp,
                           Fr* __restrict__ init_out, Fr* __restrict__ audit_out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const Fr base = fp_sub(fp_add(fp_mul(load_fr(mem + i), hp.rh), fp_mul(fr_from_u32((uint32_t)i), fr_R2())), hp.r_ms);
    store_fr(init_out + i, base);
    store_fr(audit_out + i, fp_add(base, fp_mul(fr_from_u32(audit_ts[i]), hp.rh2_R2)));
}

// read[j] = h(addr[j], mem[addr[j]], read_ts[j]), write[j] = h(addr[j], mem[addr[j]], read_ts[j] + 1)   (:775-793);
// mem[addr[j]] is the derefs value of the operation (deref_mem, :245-251).
__global__ void k_hash_ops(const Fr* __restrict__ mem, const uint32_t* __restrict__ addr, const uint32_t* __restrict__ read_ts,
                           size_t N, HashParams hp, Fr* __restrict__ read_out, Fr* __restrict__ write_out) {
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    const uint32_t a = addr[j];
    Fr v = fp_add(fp_mul(load_fr(mem + a), hp.rh), fp_mul(fr_from_u32(a), fr_R2()));
    v = fp_sub(fp_add(v, fp_mul(fr_from_u32(read_ts[j]), hp.rh2_R2)), hp.r_ms);
    store_fr(read_out + j, v);
    store_fr(write_out + j, fp_add(v, hp.rh2));
}

// dst[i] = Montgomery form of the integer src[i] (DensePolynomial::from_usize, hyrax.rs:249-251)
__global__ void k_u32_to_fr(const uint32_t* __restrict__ src, size_t n, Fr* __restrict__ dst) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    store_fr(dst + i, fp_mul(fr_from_u32(src[i]), fr_R2()));
}

// partial[block] = sum_i a[i] b[i] c[i]   (DotProductCircuit::evaluate, product_tree.rs:81-86)
__global__ void __launch_bounds__(kDotThreads)
k_fr_triple_dot(const Fr* __restrict__ a, const Fr* __restrict__ b, const Fr* __restrict__ c, size_t n, Fr* __restrict__ partial) {
    __shared__ Fr sm[kDotThreads];
    Fr acc = Fr::zero();
    for (size_t i = (size_t)blockIdx.x * kDotThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kDotThreads)
        acc = fp_add(acc, fp_mul(fp_mul(load_fr(a + i), load_fr(b + i)), load_fr(c + i)));
    acc = block_sum_fr(acc, sm, kDotThreads);
    if (threadIdx.x == 0) store_fr(partial + blockIdx.x, acc);
}

// SparseMatPolynomial::evaluate_with_tables (sparse_mlpoly_full.rs:103-108): partial[block] of
// sum_i val[i] * mem_rx[row[i]] * mem_ry[col[i]]
__global__ void __launch_bounds__(kDotThreads)
k_sparse_eval(const Fr* __restrict__ val, const uint32_t* __restrict__ row, const uint32_t* __restrict__ col,
              const Fr* __restrict__ mem_rx, const Fr* __restrict__ mem_ry, size_t n, Fr* __restrict__ partial) {
    __shared__ Fr sm[kDotThreads];
    Fr acc = Fr::zero();
    for (size_t i = (size_t)blockIdx.x * kDotThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kDotThreads) {
        const Fr v = load_fr(val + i);
        if (v.is_zero()) continue;                       // padding entries
        acc = fp_add(acc, fp_mul(fp_mul(v, load_fr(mem_rx + row[i])), load_fr(mem_ry + col[i])));
    }
    acc = block_sum_fr(acc, sm, kDotThreads);
    if (threadIdx.x == 0) store_fr(partial + blockIdx.x, acc);
}

// Sparse matrix times vector, compressed-row form: out[i] = sum_m coeff_m * sum_{k in [ptr_m[i], ptr_m[i+1])} val_m[k] * vec[idx_m[k]]
// (SparseMatPolynomial::multiply_vec, sparse_mlpoly.rs:77-87, on a row-sorted copy; compute_eval_table_sparse, :145-160, on a
// column-sorted copy with vec = eq(rx)).  Up to three matrices are combined in one pass (r_A A + r_B B + r_C C).
struct SpMat { const uint32_t* ptr; const uint32_t* idx; const Fr* val; };
static constexpr uint32_t kSpmvHeavy = 2048;      // rows longer than this (the constant-1 column of an R1CS transpose) get a block

__global__ void k_spmv(SpMat m0, SpMat m1, SpMat m2, Fr c0, Fr c1, Fr c2, int nm, int use_coeff, const Fr* __restrict__ vec, size_t n,
                       Fr* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const SpMat mats[3] = {m0, m1, m2};
    const Fr coeffs[3] = {c0, c1, c2};
    Fr total = Fr::zero();
    for (int m = 0; m < nm; m++) {
        Fr acc = Fr::zero();
        const uint32_t k0 = mats[m].ptr[i], k1 = mats[m].ptr[i + 1];
        if (k1 - k0 > kSpmvHeavy) continue;                // a whole block sums this row in k_spmv_heavy
        for (uint32_t k = k0; k < k1; k++)
            acc = fp_add(acc, fr_mul_call(load_fr(mats[m].val + k), load_fr(vec + mats[m].idx[k])));
        if (use_coeff) acc = fr_mul_call(acc, coeffs[m]);
        total = fp_add(total, acc);
    }
    store_fr(out + i, total);
}

// out[row] += coeff * sum_k val[k] * vec[idx[k]] for the heavy rows of one matrix.  The constant-1 column of an R1CS transpose
// holds one entry per constraint (2^20 at keyless scale): one block walking it is 4096 dependent gather + product steps per
// thread (8 ms, most of the second sumcheck's set-up).  kSpmvSplit blocks per heavy row write partial sums, a second launch
// adds them up.
static constexpr int kSpmvSplit = 128;
__global__ void __launch_bounds__(kDotThreads)
k_spmv_heavy_partial(SpMat m, const uint32_t* __restrict__ heavy_rows, const Fr* __restrict__ vec, Fr* __restrict__ part) {
    __shared__ Fr sm[kDotThreads];
    const uint32_t row = heavy_rows[blockIdx.y];
    Fr acc = Fr::zero();
    for (uint32_t k = m.ptr[row] + blockIdx.x * kDotThreads + threadIdx.x; k < m.ptr[row + 1]; k += kSpmvSplit * kDotThreads)
        acc = fp_add(acc, fp_mul(load_fr(m.val + k), load_fr(vec + m.idx[k])));
    acc = block_sum_fr(acc, sm, kDotThreads);
    if (threadIdx.x == 0) store_fr(part + (size_t)blockIdx.y * kSpmvSplit + blockIdx.x, acc);
}
__global__ void __launch_bounds__(kDotThreads)
k_spmv_heavy_final(const uint32_t* __restrict__ heavy_rows, const Fr* __restrict__ part, Fr coeff, int use_coeff, Fr* __restrict__ out) {
    __shared__ Fr sm[kDotThreads];
    const uint32_t row = heavy_rows[blockIdx.x];
    Fr acc = Fr::zero();
    for (int i = threadIdx.x; i < kSpmvSplit; i += kDotThreads) acc = fp_add(acc, load_fr(part + (size_t)blockIdx.x * kSpmvSplit + i));
    acc = block_sum_fr(acc, sm, kDotThreads);
    if (threadIdx.x == 0) {
        if (use_coeff) acc = fp_mul(acc, coeff);
        store_fr(out + row, fp_add(load_fr(out + row), acc));
    }
}

}  // namespace sbn
