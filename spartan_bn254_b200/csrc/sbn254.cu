// libsbn254 C ABI (include/sbn254.h): contexts, resident generator sets, the batched Hyrax commit
// pipeline and its stream plumbing.  Kernels live in msm_kernels.cuh / opening_kernels.cuh.
#define SBN_HOST_FAST_FP 1      // host code of this library multiplies on 4 x 64-bit limbs (fp.cuh); the host TESTS do not define it
#include "../../include/sbn254.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "msm_kernels.cuh"
#include "opening_kernels.cuh"
#include "ba_kernels.cuh"
#include "small_kernels.cuh"
#include "mult_kernels.cuh"
#include "prodtree_kernels.cuh"
#include "transcript_kernels.cuh"
#include "host/keccak.hpp"
#include "host/merlin.hpp"

using namespace sbn;

static_assert(sizeof(Fr) == sizeof(sbn_fr), "Fr layout");
static_assert(sizeof(Affine) == sizeof(sbn_g1a), "affine layout");
static_assert(sizeof(XYZZ) == 128, "xyzz layout");

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

struct HostBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaEvent_t busy = nullptr;      // the last H2D copy that read this buffer (asynchronous calls return before it is done)
};

struct sbn_ctx {
    int device = 0;
    cudaStream_t compute = nullptr, copy = nullptr;
    std::mutex mu;
    std::string last_error;
    long chunk_rows = 0;    // 0 = auto (see commit_chunk_rows)
    long dedup_generators = 1;   // merge equal generators of a set created afterwards (k_aggregate_rows)
    long first_chunk_rows = 0;   // host path: rows of the short first chunk; 0 = auto (an eighth of a chunk, measured best)
    long window_bits = 0;
    long task_cap = 0;      // 0 = auto: 2.5 x the mean bucket occupancy
    long ba_rounds = -1;    // batched-affine pre-reduction rounds before the XYZZ accumulation (0..3); -1 = auto
    long ba_batch = 0;      // pairs per thread in a round; 0 = auto
    long leaf_m = 0;        // buckets per leaf thread of the two-level reduction; 0 = auto
    uint64_t launches = 0, h2d = 0, d2h = 0;
    uint64_t pool_flushes = 0;         // times an allocation failure emptied the buffer pool (dev_malloc)
    uint64_t mult_fallbacks = 0;       // digit-multiple tables that could not be allocated: those sets use the bucket pipeline
    // grow-only workspaces
    struct Slot {          // one in-flight chunk of rows: private workspace
        DevBuf entries, tstart, tasks, partials, heavy, pairs;
        DevBuf pts[3], prefix, other, wtot, winv;      // batched-affine rounds (ba_kernels.cuh)
        DevBuf zagg;                                   // per-group scalar sums when the generator set has duplicates
    } slots[4];
    // Pipeline streams.  The latency-bound stages (sort, split-bucket fold, bucket reduction) run on a HIGH priority
    // stream and the IMAD-bound accumulation on LOW priority ones, so that while chunk i accumulates, the blocks of
    // sort(i+1) and reduce(i-1) are placed first as accumulation blocks retire and fill its idle issue slots.
    cudaStream_t hi = nullptr, lo[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t fork = nullptr, join_hi = nullptr;
    DevBuf totals, dZ, dblinds, dC, dinf, scratch0, scratch1, scratch2;
    DevBuf mtotals[2];                 // row totals of the tabulated-sum path, one per workspace set
    uint64_t mult_calls = 0, host_calls = 0;
    DevBuf hZ[3], hC[3], hI[3], hB[3];   // staging of sbn_hyrax_commit_async, one set per call in turn
    // Completion of the last call that used a workspace set (tabulated-sum path) / a staging set: the next call that takes the
    // set waits for it on the device, whatever stream the caller put it on (with two alternating caller streams the wait is
    // already implied by stream order; with three, or with one stream per commit, it is what keeps the sets apart).
    cudaEvent_t set_done[2] = {nullptr, nullptr}, hset_done[3] = {nullptr, nullptr, nullptr};
    HostBuf stage_pin[2][2];           // pinned ring for pageable host scalars: [workspace set][chunk parity]
    HostBuf out_pin[3];                // pinned landing buffers for the results of asynchronous calls with pageable outputs
    int force_set = -1, last_was_mult = 0;
    DevBuf spmv_part;                  // partial sums of the heavy rows of a sparse matrix-vector product
    DevBuf scan;                       // two words: largest bit length of a sample / of all scalars (k_max_bits)
    DevBuf zkeep;                      // z of the last sbn_sumcheck_begin_r1cs, reused by _begin_quad_r1cs(z = NULL)
    size_t zkeep_len = 0;
    DevBuf tabpart;                    // per-block partial sums of the tabulated few-row commit (small_kernels.cuh)
    long mult_max_mb = 6144;           // largest digit-multiple table built for a commit's generator set (MiB); 0 = none
    long mult_rounds = 0;              // batched-affine rounds of the tabulated-sum path; 0 = auto (mult_rounds_for)
    long mult_min_rows = 256;          // commits of at least this many rows take the tabulated-sum path
    long mult_streams = 2;             // chunks of the tabulated-sum path in flight (streams / workspaces), 1..4
    long ba_minb = 3;                  // finish pass register target: resident CTAs per SM (3: 80 registers, 4: 64)
    long ba_prefetch = 0;              // round 1 (measured: loses 5 %): entries read two pairs ahead, table points prefetched into L2 one pair ahead
    long finish_smem_kb = 0;           // dynamic shared memory requested by the finish pass: caps its resident blocks so that a prefix pass of the other stream co-resides
    long prefix_smem_kb = 0;
    long sum_wpr = 0;                  // warps per row of the final row sums (1, 2, 4); 0 = by the number of points left
    long ablate = 0;                   // PROFILING ONLY (results are wrong when non-zero): bit mask of skipped launches of the tabulated-sum path
    long fused_rounds = 1;             // sbn_bsumcheck_prove: the bind of a round rides in the next round's evaluation kernel (one launch per round on the Fiat-Shamir chain)
    long host_normalize = 1;           // the few points of a short commitment / a bullet round are normalised (XYZZ -> affine) on the host: a 30 us one-warp dependency chain on the device, a few us on a host core
    long bsc_device = 0;               // product-layer sumchecks (transcript_kernels.cuh; measured no faster than the host loop, kept as an option): 1 = the short last rounds of a layer run in one block with the Merlin transcript on the device, the long ones through the host loop; 2 = every round on the device; 0 = host loop only
    long small_scalar_path = 1;        // commits without blinds scan their scalars' bit length and use a short window schedule when it is small
    long small_scalar_hits = 0;        // commits that took it
    long mult_layout = 1;              // 1: position-major lists, separate passes (default); 2: prefix passes fused into the previous round (measured: 2.69 vs 2.65 ms); 0: row-major (round 1)
    long tab_max_mb = 3072;            // largest digit-multiple table built for an opening's generator set (MiB); 0 = none
    // Pool of released table-sized device buffers (product circuits, resident polynomials, sumcheck tables): a proof
    // allocates and releases ~5 GB of them, and cudaFree costs ~35 ms per 268 MB buffer (574 ms per keyless-scale proof).
    // Everything that touches them is ordered on `compute`, so a released buffer can be handed out again without a sync.
    std::vector<std::pair<size_t, void*>> mem_pool;
    size_t mem_pool_bytes = 0;
    std::unordered_map<void*, size_t> pool_live;   // size class of every buffer handed out by pool_alloc
    uint8_t *ev_pin = nullptr, *ev_pin_dev = nullptr;   // mapped pinned: the round evaluations of the in-library sumcheck loops
    uint8_t* small_pin = nullptr;      // mapped pinned staging of the short-commitment path (scalars in, points out)
    uint8_t* small_pin_dev = nullptr;
    long small_commit_path = 1;        // 0: short generator sets go through the general pipeline (test hook)
    // last-commit profile
    std::vector<cudaEvent_t> ev_pool;
    float prof_ms[4] = {0, 0, 0, 0};
    int prof_launches[4] = {0, 0, 0, 0};
};

struct sbn_bases {
    sbn_ctx* ctx = nullptr;
    size_t n = 0;      // generators without h
    int has_g1 = 0;    // extra generator g1 (DotProductProofGens::gens_1.G[0]) in column n
    int n1 = 0;        // table columns: n + has_g1 + 1, h is the last one
    int c = 0, W = 0, nb = 0;
    Affine* table = nullptr;   // W * n1 affine points
    Affine* orig = nullptr;    // the n + has_g1 + 1 generators as given (the explicit-folding bullet path reads them)
    // duplicate generators merged (k_aggregate_rows): n1 counts DISTINCT non-identity points, n_cols the given columns
    int dedup = 0, n_cols = 0, n_big = 0;
    uint32_t *gptr = nullptr, *gcols = nullptr, *gbig = nullptr;
    Affine* small = nullptr;   // short sets (n_cols <= kSmallMaxCols): every digit multiple of every generator (small_kernels.cuh)
    // every digit multiple of the n1 table columns for many-row commits (mult_kernels.cuh), built on the first such commit
    Affine* mult = nullptr;
    int mc = 0, mW = 0, mult_tried = 0, mult_fails = 0;
    // the same kind of table for SMALL scalars: msW windows of msc bits cover values below 2^(msW * msc - 1) (encode-time commits)
    Affine* mult_s = nullptr;
    int msc = 0, msW = 0, max_group = 1;
};

#define SBN_CUDA(ctx, call)                                                                      \
    do {                                                                                         \
        cudaError_t _e = (call);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            (ctx)->last_error = std::string(#call) + ": " + cudaGetErrorString(_e);              \
            return _e == cudaErrorMemoryAllocation ? SBN_ERR_OOM : SBN_ERR_CUDA;                 \
        }                                                                                        \
    } while (0)

// Entry of an API call: this thread talks to the context's device, and whatever non-sticky error an earlier runtime call of
// this thread left behind (another context's teardown, the caller's own CUDA code) is dropped, so that the
// cudaGetLastError() after this call's launches reports this call's errors only.
#define SBN_ENTER(ctx)                                  \
    do {                                                \
        SBN_CUDA(ctx, cudaSetDevice((ctx)->device));    \
        cudaGetLastError();                             \
    } while (0)

#define SBN_TRY(expr)            \
    do {                         \
        int _s = (expr);         \
        if (_s != SBN_OK) return _s; \
    } while (0)

static constexpr size_t kEvPinBytes = 16384;
static constexpr size_t kPoolMinBytes = size_t(1) << 20, kPoolMaxBytes = size_t(48) << 30;

static void pool_flush(sbn_ctx* ctx) {
    for (auto& e : ctx->mem_pool) cudaFree(e.second);
    ctx->mem_pool.clear();
    ctx->mem_pool_bytes = 0;
}
// cudaMalloc for everything that is not pooled.  The pool may be sitting on tens of GB of released buffers: when the
// allocator says no, the sticky error is cleared, the pool is emptied (after the work that may still read its buffers has
// drained) and the allocation is tried once more, so that a long-lived prover does not see a spurious SBN_ERR_OOM.
static cudaError_t dev_malloc(sbn_ctx* ctx, void** out, size_t bytes) {
    cudaError_t e = cudaMalloc(out, bytes);
    if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        if (!ctx->mem_pool.empty()) {
            cudaDeviceSynchronize();
            pool_flush(ctx);
            ctx->pool_flushes++;
            e = cudaMalloc(out, bytes);
            if (e == cudaErrorMemoryAllocation) cudaGetLastError();
        }
    }
    if (e != cudaSuccess) *out = nullptr;
    return e;
}
template <class T>
static cudaError_t dev_malloc(sbn_ctx* ctx, T** out, size_t bytes) { return dev_malloc(ctx, (void**)out, bytes); }

static int ensure(sbn_ctx* ctx, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap) return SBN_OK;
    if (b.p) {
        SBN_CUDA(ctx, cudaDeviceSynchronize());
        SBN_CUDA(ctx, cudaFree(b.p));
        b.p = nullptr;
        b.cap = 0;
    }
    size_t want = bytes + bytes / 8;
    SBN_CUDA(ctx, dev_malloc(ctx, &b.p, want));
    b.cap = want;
    return SBN_OK;
}
// Size class of a pooled buffer: table-sized requests recur with exact sizes (powers of two times 32 B); small ones (the
// per-layer scratch of the sumchecks and the bullet reduction) are rounded up to a power of two so that they recur too --
// a cudaMalloc / cudaFree pair costs about a millisecond next to several GB of live tables, and a proof makes hundreds.
static size_t pool_class(size_t bytes) {
    if (bytes >= kPoolMinBytes) return bytes;
    size_t c = 256;
    while (c < bytes) c <<= 1;
    return c;
}
static cudaError_t pool_alloc(sbn_ctx* ctx, void** out, size_t bytes) {
    const size_t cls = pool_class(bytes);
    for (size_t i = ctx->mem_pool.size(); i-- > 0;)
        if (ctx->mem_pool[i].first == cls) {
            *out = ctx->mem_pool[i].second;
            ctx->mem_pool_bytes -= cls;
            ctx->mem_pool.erase(ctx->mem_pool.begin() + i);
            ctx->pool_live[*out] = cls;
            return cudaSuccess;
        }
    cudaError_t e = dev_malloc(ctx, out, cls);
    if (e == cudaSuccess) ctx->pool_live[*out] = cls;
    return e;
}
template <class T>
static cudaError_t pool_alloc(sbn_ctx* ctx, T** out, size_t bytes) { return pool_alloc(ctx, (void**)out, bytes); }
static void pool_free(sbn_ctx* ctx, void* p, size_t bytes) {
    if (!p) return;
    const size_t cls = pool_class(bytes);
    ctx->pool_live.erase(p);
    if (ctx->mem_pool_bytes + cls <= kPoolMaxBytes) {
        ctx->mem_pool.emplace_back(cls, p);
        ctx->mem_pool_bytes += cls;
    } else {
        cudaFree(p);
    }
}
// Release of a pool_alloc'ed buffer whose size the owner did not keep.
static void pool_release(sbn_ctx* ctx, void* p) {
    if (!p) return;
    auto it = ctx->pool_live.find(p);
    if (it == ctx->pool_live.end()) { cudaFree(p); return; }
    const size_t cls = it->second;
    pool_free(ctx, p, cls);
}

static void release(DevBuf& b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
}

extern "C" const char* sbn_strerror(int s) {
    switch (s) {
        case SBN_OK: return "ok";
        case SBN_ERR_ARG: return "invalid argument";
        case SBN_ERR_SHAPE: return "shape precondition violated";
        case SBN_ERR_CUDA: return "CUDA error";
        case SBN_ERR_OOM: return "out of device memory";
        case SBN_ERR_UNSUPPORTED: return "unsupported";
        default: return "unknown status";
    }
}
extern "C" const char* sbn_last_cuda_error(const sbn_ctx* ctx) { return ctx ? ctx->last_error.c_str() : ""; }
extern "C" int sbn_version(void) { return 1; }
// host utility: Keccak-f[1600] for the host mirrors' Merlin transcript (200-byte little-endian state, in place)
extern "C" void sbn_keccak_f1600(uint64_t* state) { sbn::keccak::permute(state); }
// host utilities: the Merlin transcript of the host mirrors (state = 203 bytes owned by the caller)
extern "C" void sbn_merlin_init(void* state, const uint8_t* label, size_t llen) { sbn::merlin::init(*(sbn::merlin::State*)state, label, llen); }
extern "C" void sbn_merlin_append(void* state, const uint8_t* label, size_t llen, const uint8_t* msg, size_t mlen) {
    sbn::merlin::append_message(*(sbn::merlin::State*)state, label, llen, msg, mlen);
}
extern "C" void sbn_merlin_append_many(void* state, const uint8_t* label, size_t llen, const uint8_t* msgs, size_t mlen, size_t count) {
    for (size_t i = 0; i < count; i++) sbn::merlin::append_message(*(sbn::merlin::State*)state, label, llen, msgs + i * mlen, mlen);
}
extern "C" void sbn_merlin_challenge(void* state, const uint8_t* label, size_t llen, uint8_t* out, size_t n) {
    sbn::merlin::challenge_bytes(*(sbn::merlin::State*)state, label, llen, out, n);
}

// host utility: GroupElement::compress (group.rs:135-140) of n affine Montgomery points -- ark's compressed short-Weierstrass
// encoding: x as 32 little-endian bytes, bit 7 of byte 31 set when y > p - y, bit 6 (x = 0) for the identity
extern "C" int sbn_g1_compress(const sbn_g1a* pts, const uint8_t* inf, size_t n, uint8_t* out) {
    if ((n && !pts) || !out) return SBN_ERR_ARG;
    for (size_t i = 0; i < n; i++) {
        uint8_t* o = out + 32 * i;
        Fq xm, ym;
        memcpy(xm.l, pts[i].x, 32);
        memcpy(ym.l, pts[i].y, 32);
        if ((inf && inf[i]) || (xm.is_zero() && ym.is_zero())) {
            memset(o, 0, 32);
            o[31] = 0x40;
            continue;
        }
        const Fq x = fp_from_mont(xm), y = fp_from_mont(ym), ny = fp_neg(y);    // y != 0 on a prime-order curve
        memcpy(o, x.l, 32);
        bool greater = false;
        for (int k = 7; k >= 0; k--)
            if (y.l[k] != ny.l[k]) { greater = y.l[k] > ny.l[k]; break; }
        if (greater) o[31] |= 0x80;
    }
    return SBN_OK;
}
// host utilities: Scalar <-> canonical little-endian integers for short vectors (scalar.rs:75-95); long ones change form
// on the device (sbn_fr_to_canonical / sbn_fr_from_canonical)
extern "C" int sbn_fr_to_canonical_host(const sbn_fr* in, size_t n, uint64_t* canon) {
    if (n && (!in || !canon)) return SBN_ERR_ARG;
    for (size_t i = 0; i < n; i++) {
        Fr v;
        memcpy(v.l, &in[i], 32);
        v = fp_from_mont(v);
        memcpy(canon + 4 * i, v.l, 32);
    }
    return SBN_OK;
}
extern "C" int sbn_fr_from_canonical_host(const uint64_t* canon, size_t n, sbn_fr* out) {
    if (n && (!canon || !out)) return SBN_ERR_ARG;
    for (size_t i = 0; i < n; i++) {
        Fr v;
        memcpy(v.l, canon + 4 * i, 32);
        v = fp_to_mont(v);             // any 256-bit value: the product with R^2 reduces it
        memcpy(&out[i], v.l, 32);
    }
    return SBN_OK;
}
// PolyCommitment::append_to_transcript's share loop (hyrax.rs:46-50): compress and append n points under one label
extern "C" int sbn_merlin_append_points(void* state, const uint8_t* label, size_t llen, const sbn_g1a* pts, const uint8_t* inf, size_t n) {
    if (!state || !label || (n && !pts)) return SBN_ERR_ARG;
    uint8_t buf[32];
    for (size_t i = 0; i < n; i++) {
        sbn_g1_compress(pts + i, inf ? inf + i : nullptr, 1, buf);
        sbn::merlin::append_message(*(sbn::merlin::State*)state, label, llen, buf, 32);
    }
    return SBN_OK;
}

extern "C" int sbn_ctx_create(int device, sbn_ctx** out) {
    if (!out) return SBN_ERR_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return SBN_ERR_CUDA;
    sbn_ctx* ctx = new (std::nothrow) sbn_ctx();
    if (!ctx) return SBN_ERR_OOM;
    ctx->device = device;
    bool ok = cudaSetDevice(device) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->compute, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->copy, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreateWithFlags(&ctx->fork, cudaEventDisableTiming) == cudaSuccess;
    int prio_least = 0, prio_greatest = 0;
    ok = ok && cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest) == cudaSuccess &&
         cudaStreamCreateWithPriority(&ctx->hi, cudaStreamNonBlocking, prio_greatest) == cudaSuccess &&
         cudaStreamCreateWithPriority(&ctx->lo[0], cudaStreamNonBlocking, prio_least) == cudaSuccess &&
         cudaStreamCreateWithPriority(&ctx->lo[1], cudaStreamNonBlocking, prio_least) == cudaSuccess &&
         cudaStreamCreateWithPriority(&ctx->lo[2], cudaStreamNonBlocking, prio_least) == cudaSuccess &&
         cudaStreamCreateWithPriority(&ctx->lo[3], cudaStreamNonBlocking, prio_least) == cudaSuccess &&
         cudaEventCreateWithFlags(&ctx->join_hi, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
        delete ctx;
        return SBN_ERR_CUDA;
    }
    // Bench / test / tuning hooks (scripts/sweep_mult.sh).  They go through sbn_ctx_set, so an out-of-range value is
    // ignored exactly as the call would reject it (the batched-affine workspaces are sized for >= 4 pairs per thread).
    static const char* const kEnvHooks[][2] = {{"SBN_MULT_MAX_MB", "mult_max_mb"}, {"SBN_MULT_ROUNDS", "mult_rounds"},
                                               {"SBN_BA_BATCH", "ba_batch"},       {"SBN_BA_ROUNDS", "ba_rounds"},
                                               {"SBN_MULT_STREAMS", "mult_streams"}};
    for (auto& hk : kEnvHooks)
        if (const char* e = getenv(hk[0])) sbn_ctx_set(ctx, hk[1], atol(e));
    *out = ctx;
    return SBN_OK;
}

extern "C" int sbn_ctx_destroy(sbn_ctx* ctx) {
    if (!ctx) return SBN_ERR_ARG;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->compute);
    cudaStreamSynchronize(ctx->copy);
    for (DevBuf* b : {&ctx->totals, &ctx->dZ, &ctx->dblinds, &ctx->dC, &ctx->dinf, &ctx->scratch0, &ctx->scratch1,
                      &ctx->scratch2, &ctx->tabpart, &ctx->zkeep, &ctx->scan, &ctx->spmv_part, &ctx->mtotals[0], &ctx->mtotals[1], &ctx->hZ[0], &ctx->hZ[1], &ctx->hZ[2], &ctx->hC[0], &ctx->hC[1], &ctx->hC[2], &ctx->hI[0], &ctx->hI[1],
                      &ctx->hI[2], &ctx->hB[0], &ctx->hB[1], &ctx->hB[2]})
        release(*b);
    for (cudaEvent_t e : {ctx->set_done[0], ctx->set_done[1], ctx->hset_done[0], ctx->hset_done[1], ctx->hset_done[2]})
        if (e) cudaEventDestroy(e);
    for (cudaStream_t st : {ctx->hi, ctx->lo[0], ctx->lo[1], ctx->lo[2], ctx->lo[3]})
        if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
    for (auto& sl : ctx->slots)
        for (DevBuf* b : {&sl.entries, &sl.tstart, &sl.tasks, &sl.partials, &sl.heavy, &sl.pairs, &sl.pts[0], &sl.pts[1], &sl.pts[2],
                          &sl.prefix, &sl.other, &sl.wtot, &sl.winv, &sl.zagg})
            release(*b);
    pool_flush(ctx);
    cudaEventDestroy(ctx->fork);
    if (ctx->join_hi) cudaEventDestroy(ctx->join_hi);
    for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
    cudaStreamDestroy(ctx->compute);
    cudaStreamDestroy(ctx->copy);
    if (ctx->small_pin) cudaFreeHost(ctx->small_pin);
    if (ctx->ev_pin) cudaFreeHost(ctx->ev_pin);
    for (auto& a : ctx->stage_pin) for (auto& r : a) { if (r.p) cudaFreeHost(r.p); if (r.busy) cudaEventDestroy(r.busy); }
    for (auto& r : ctx->out_pin) if (r.p) cudaFreeHost(r.p);
    delete ctx;
    cudaGetLastError();
    return SBN_OK;
}

extern "C" int sbn_ctx_synchronize(sbn_ctx* ctx) {
    if (!ctx) return SBN_ERR_ARG;
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_ENTER(ctx);
    SBN_CUDA(ctx, cudaStreamSynchronize(ctx->copy));
    SBN_CUDA(ctx, cudaStreamSynchronize(ctx->compute));
    return SBN_OK;
}

extern "C" int sbn_ctx_set(sbn_ctx* ctx, const char* key, long value) {
    if (!ctx || !key) return SBN_ERR_ARG;
    std::lock_guard<std::mutex> g(ctx->mu);
    if (!strcmp(key, "small_commit_path")) {
        ctx->small_commit_path = value ? 1 : 0;
    } else if (!strcmp(key, "mult_max_mb")) {
        if (value < 0) return SBN_ERR_ARG;
        ctx->mult_max_mb = value;
    } else if (!strcmp(key, "mult_rounds")) {
        if (value < 0 || value > 12) return SBN_ERR_ARG;
        ctx->mult_rounds = value;
    } else if (!strcmp(key, "mult_min_rows")) {
        if (value < 1) return SBN_ERR_ARG;
        ctx->mult_min_rows = value;
    } else if (!strcmp(key, "mult_streams")) {
        if (value < 1 || value > 4) return SBN_ERR_ARG;
        ctx->mult_streams = value;
    } else if (!strcmp(key, "ba_minb")) {
        if (value != 3 && value != 4) return SBN_ERR_ARG;
        ctx->ba_minb = value;
    } else if (!strcmp(key, "finish_smem_kb") || !strcmp(key, "prefix_smem_kb")) {
        if (value < 0 || value > 200) return SBN_ERR_ARG;
        (key[0] == 'f' ? ctx->finish_smem_kb : ctx->prefix_smem_kb) = value;
        const int bytes = (int)value << 10;
        cudaSetDevice(ctx->device);
        if (key[0] == 'f') {
            cudaFuncSetAttribute(k_bat_finish<true, 3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
            cudaFuncSetAttribute(k_bat_finish<true, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
            cudaFuncSetAttribute(k_bat_finish<true, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
            cudaFuncSetAttribute(k_bat_finish<true, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
            cudaFuncSetAttribute(k_bat_finish<false, 3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
            cudaFuncSetAttribute(k_bat_finish<false, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        } else {
            cudaFuncSetAttribute(k_bat_prefix<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
            cudaFuncSetAttribute(k_bat_prefix<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
            cudaFuncSetAttribute(k_bat_prefix<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        }
        if (cudaGetLastError() != cudaSuccess) return SBN_ERR_CUDA;
    } else if (!strcmp(key, "sum_wpr")) {
        if (value != 0 && value != 1 && value != 2 && value != 4) return SBN_ERR_ARG;
        ctx->sum_wpr = value;
    } else if (!strcmp(key, "ablate")) {
        ctx->ablate = value;
    } else if (!strcmp(key, "small_scalar_path")) {
        ctx->small_scalar_path = value ? 1 : 0;
    } else if (!strcmp(key, "mult_layout")) {
        if (value < 0 || value > 2) return SBN_ERR_ARG;
        ctx->mult_layout = value;
    } else if (!strcmp(key, "fused_rounds")) {
        ctx->fused_rounds = value ? 1 : 0;
    } else if (!strcmp(key, "host_normalize")) {
        ctx->host_normalize = value ? 1 : 0;
    } else if (!strcmp(key, "bsc_device")) {
        ctx->bsc_device = value < 0 || value > 2 ? 0 : value;
    } else if (!strcmp(key, "ba_prefetch")) {
        ctx->ba_prefetch = value ? 1 : 0;
    } else if (!strcmp(key, "l2_fetch")) {      // cudaLimitMaxL2FetchGranularity: the table gathers are random 64 B reads
        if (value != 32 && value != 64 && value != 128) return SBN_ERR_ARG;
        if (cudaSetDevice(ctx->device) != cudaSuccess ||
            cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)value) != cudaSuccess) {
            cudaGetLastError();
            return SBN_ERR_CUDA;
        }
    } else if (!strcmp(key, "tab_max_mb")) {
        if (value < 0) return SBN_ERR_ARG;
        ctx->tab_max_mb = value;
    } else if (!strcmp(key, "dedup_generators")) {
        ctx->dedup_generators = value ? 1 : 0;
    } else if (!strcmp(key, "first_chunk_rows")) {
        if (value < 0) return SBN_ERR_ARG;
        ctx->first_chunk_rows = value;
    } else if (!strcmp(key, "chunk_rows")) {
        if (value < 0) return SBN_ERR_ARG;
        ctx->chunk_rows = value;
    } else if (!strcmp(key, "task_cap")) {
        if (value < 0 || value > kMaxTaskCap) return SBN_ERR_ARG;
        ctx->task_cap = value;
    } else if (!strcmp(key, "ba_rounds")) {
        if (value < -1 || value > 3) return SBN_ERR_ARG;
        ctx->ba_rounds = value;
    } else if (!strcmp(key, "ba_batch")) {
        if (value != 0 && (value < 4 || value > 256)) return SBN_ERR_ARG;
        ctx->ba_batch = value;
    } else if (!strcmp(key, "leaf_m")) {
        if (value < 0 || (value & (value - 1))) return SBN_ERR_ARG;
        ctx->leaf_m = value;
    } else if (!strcmp(key, "window_bits")) {
        if (value != 0 && (value < kMinWindowBits || value > kMaxWindowBits)) return SBN_ERR_ARG;
        ctx->window_bits = value;
    } else {
        return SBN_ERR_ARG;
    }
    return SBN_OK;
}

extern "C" int sbn_ctx_counters(sbn_ctx* ctx, uint64_t* launches, uint64_t* h2d, uint64_t* d2h, int reset) {
    if (!ctx) return SBN_ERR_ARG;
    std::lock_guard<std::mutex> g(ctx->mu);
    if (launches) *launches = ctx->launches;
    if (h2d) *h2d = ctx->h2d;
    if (d2h) *d2h = ctx->d2h;
    if (reset) ctx->launches = ctx->h2d = ctx->d2h = 0;
    return SBN_OK;
}

extern "C" int sbn_ctx_memory_stats(sbn_ctx* ctx, uint64_t out[8]) {
    if (!ctx || !out) return SBN_ERR_ARG;
    std::lock_guard<std::mutex> g(ctx->mu);
    out[0] = ctx->mem_pool_bytes;
    out[1] = ctx->pool_flushes;
    out[2] = ctx->mult_fallbacks;
    out[3] = ctx->mem_pool.size();
    out[4] = (uint64_t)ctx->small_scalar_hits;
    out[5] = out[6] = out[7] = 0;
    return SBN_OK;
}

extern "C" int sbn_ctx_last_commit_profile(sbn_ctx* ctx, float ms[4], int launches[4]) {
    if (!ctx) return SBN_ERR_ARG;
    std::lock_guard<std::mutex> g(ctx->mu);
    for (int i = 0; i < 4; i++) {
        if (ms) ms[i] = ctx->prof_ms[i];
        if (launches) launches[i] = ctx->prof_launches[i];
    }
    return SBN_OK;
}

extern "C" int sbn_host_alloc(void** out, size_t bytes) {
    if (!out) return SBN_ERR_ARG;
    return cudaHostAlloc(out, bytes, cudaHostAllocDefault) == cudaSuccess ? SBN_OK : SBN_ERR_OOM;
}
extern "C" int sbn_host_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? SBN_OK : SBN_ERR_CUDA; }

// Streams for the asynchronous entry points, for callers without CUDA bindings of their own (the Rust shim): a non-blocking
// stream on the context's device, its synchronisation and its release.
extern "C" int sbn_stream_create(sbn_ctx* ctx, void** stream_out) {
    if (!ctx || !stream_out) return SBN_ERR_ARG;
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_ENTER(ctx);
    cudaStream_t st = nullptr;
    SBN_CUDA(ctx, cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    *stream_out = (void*)st;
    return SBN_OK;
}
extern "C" int sbn_stream_synchronize(sbn_ctx* ctx, void* stream) {
    if (!ctx || !stream) return SBN_ERR_ARG;
    // not under the context's mutex: other threads may issue calls on other streams while this one waits
    if (cudaSetDevice(ctx->device) != cudaSuccess) return SBN_ERR_CUDA;
    return cudaStreamSynchronize((cudaStream_t)stream) == cudaSuccess ? SBN_OK : SBN_ERR_CUDA;
}
extern "C" int sbn_stream_destroy(sbn_ctx* ctx, void* stream) {
    if (!ctx || !stream) return SBN_ERR_ARG;
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_ENTER(ctx);
    SBN_CUDA(ctx, cudaStreamSynchronize((cudaStream_t)stream));
    SBN_CUDA(ctx, cudaStreamDestroy((cudaStream_t)stream));
    return SBN_OK;
}

// ------------------------------------------------------------------------------------------------
// window selection: Fq multiplications per row = W * n1 * 10 (mixed adds) + 2 * nb * 14 * 1.3 (reduce)
// ------------------------------------------------------------------------------------------------
static int choose_window(size_t n1) {
    int best = kMinWindowBits;
    double best_cost = 1e300;
    for (int c = kMinWindowBits; c <= kMaxWindowBits; c++) {
        double W = msm_num_windows(c), nb = double(1 << (c - 1));
        double cost = W * double(n1) * 10.0 + 2.0 * nb * 14.0 * 1.3;
        if (cost < best_cost) { best_cost = cost; best = c; }
    }
    return best;
}

// Natural buckets (Poisson around the mean occupancy) stay whole; only genuinely heavy ones -- the top
// window of a 254-bit scalar has few distinct digits, derefs-style inputs repeat scalars -- are split.
static int task_cap_for(const sbn_ctx* ctx, const sbn_bases* b, int ba);

// ------------------------------------------------------------------------------------------------
// bases
// ------------------------------------------------------------------------------------------------
static int bases_create(sbn_ctx* ctx, const sbn_g1a* G, const uint8_t* G_inf, size_t n, const sbn_g1a* g1, const sbn_g1a* h,
                        sbn_bases** out) {
    if (!ctx || !G || !h || !out) return SBN_ERR_ARG;
    *out = nullptr;
    if (n == 0 || n > (1u << 24)) return SBN_ERR_SHAPE;
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_ENTER(ctx);
    sbn_bases* b = new (std::nothrow) sbn_bases();
    if (!b) return SBN_ERR_OOM;
    b->ctx = ctx;
    b->n = n;
    b->has_g1 = g1 ? 1 : 0;
    b->n1 = (int)n + 1 + b->has_g1;
    b->n_cols = b->n1;
    // group equal points (byte-equal affine coordinates); identity points drop out
    std::vector<uint32_t> gptr, gcols, gbig;
    std::vector<sbn_g1a> distinct;
    if (!g1 && ctx->dedup_generators) {
        std::unordered_map<std::string, uint32_t> ids;
        std::vector<std::vector<uint32_t>> members;
        auto is_identity = [&](const sbn_g1a* p, size_t j) {
            if (j < n && G_inf && G_inf[j]) return true;
            static const sbn_g1a zero = {};
            return memcmp(p, &zero, sizeof(sbn_g1a)) == 0;
        };
        for (size_t j = 0; j <= n; j++) {
            const sbn_g1a* p = j < n ? &G[j] : h;
            if (is_identity(p, j)) continue;
            std::string key((const char*)p, sizeof(sbn_g1a));
            auto it = ids.find(key);
            if (it == ids.end()) {
                it = ids.emplace(key, (uint32_t)members.size()).first;
                members.emplace_back();
                distinct.push_back(*p);
            }
            members[it->second].push_back((uint32_t)j);
        }
        if (!members.empty() && members.size() * 10 <= (size_t)b->n1 * 8) {       // at least a fifth of the columns merge away
            b->dedup = 1;
            gptr.push_back(0);
            for (size_t gi = 0; gi < members.size(); gi++) {
                for (uint32_t cidx : members[gi]) gcols.push_back(cidx);
                gptr.push_back((uint32_t)gcols.size());
                if (members[gi].size() > kAggBig) gbig.push_back((uint32_t)gi);
            }
            b->n1 = (int)members.size();
            b->n_big = (int)gbig.size();
            for (auto& mv : members) b->max_group = std::max<int>(b->max_group, (int)mv.size());
        }
    }
    b->c = ctx->window_bits ? (int)ctx->window_bits : choose_window((size_t)b->n1);
    b->W = msm_num_windows(b->c);
    b->nb = 1 << (b->c - 1);
    if ((uint64_t)b->W * b->n1 >= (1ull << 31)) { delete b; return SBN_ERR_SHAPE; }

    Affine* dbases = nullptr;
    uint8_t* dinf = nullptr;
    auto fail = [&](int code) {
        if (dbases) cudaFree(dbases);
        if (dinf) cudaFree(dinf);
        if (b->table) cudaFree(b->table);
        if (b->orig) cudaFree(b->orig);
        if (b->small) cudaFree(b->small);
        for (uint32_t* p : {b->gptr, b->gcols, b->gbig}) if (p) cudaFree(p);
        delete b;
        return code;
    };
    cudaError_t e;
    if ((e = dev_malloc(ctx, &b->orig, sizeof(Affine) * b->n_cols)) != cudaSuccess) {
        ctx->last_error = std::string("sbn_bases_create cudaMalloc: ") + cudaGetErrorString(e);
        return fail(SBN_ERR_OOM);
    }
    if ((e = dev_malloc(ctx, &dbases, sizeof(Affine) * b->n1)) != cudaSuccess ||
        (e = dev_malloc(ctx, &dinf, b->n1)) != cudaSuccess ||
        (e = dev_malloc(ctx, &b->table, sizeof(Affine) * (size_t)b->W * b->n1)) != cudaSuccess) {
        ctx->last_error = std::string("sbn_bases_create cudaMalloc: ") + cudaGetErrorString(e);
        return fail(SBN_ERR_OOM);
    }
    std::vector<uint8_t> inf_host(std::max(b->n1, b->n_cols), 0);
    if (G_inf && !b->dedup) memcpy(inf_host.data(), G_inf, n);
    // the generators as given (identity flags folded into the (0, 0) encoding)
    {
        std::vector<sbn_g1a> given(b->n_cols);
        memcpy(given.data(), G, sizeof(sbn_g1a) * n);
        if (G_inf) for (size_t j = 0; j < n; j++) if (G_inf[j]) memset(&given[j], 0, sizeof(sbn_g1a));
        if (g1) given[n] = *g1;
        given[b->n_cols - 1] = *h;
        if ((e = cudaMemcpy(b->orig, given.data(), sizeof(Affine) * b->n_cols, cudaMemcpyHostToDevice)) != cudaSuccess) {
            ctx->last_error = std::string("sbn_bases_create upload: ") + cudaGetErrorString(e);
            return fail(SBN_ERR_CUDA);
        }
    }
    if (b->dedup) {
        if ((e = dev_malloc(ctx, &b->gptr, gptr.size() * sizeof(uint32_t))) != cudaSuccess ||
            (e = dev_malloc(ctx, &b->gcols, gcols.size() * sizeof(uint32_t))) != cudaSuccess ||
            (e = dev_malloc(ctx, &b->gbig, std::max<size_t>(1, gbig.size()) * sizeof(uint32_t))) != cudaSuccess ||
            (e = cudaMemcpy(b->gptr, gptr.data(), gptr.size() * sizeof(uint32_t), cudaMemcpyHostToDevice)) != cudaSuccess ||
            (e = cudaMemcpy(b->gcols, gcols.data(), gcols.size() * sizeof(uint32_t), cudaMemcpyHostToDevice)) != cudaSuccess ||
            (gbig.size() && (e = cudaMemcpy(b->gbig, gbig.data(), gbig.size() * sizeof(uint32_t), cudaMemcpyHostToDevice)) != cudaSuccess) ||
            (e = cudaMemcpyAsync(dbases, distinct.data(), sizeof(Affine) * b->n1, cudaMemcpyHostToDevice, ctx->compute)) != cudaSuccess ||
            (e = cudaMemcpyAsync(dinf, inf_host.data(), b->n1, cudaMemcpyHostToDevice, ctx->compute)) != cudaSuccess) {
            ctx->last_error = std::string("sbn_bases_create upload: ") + cudaGetErrorString(e);
            return fail(SBN_ERR_CUDA);
        }
    } else if ((e = cudaMemcpyAsync(dbases, G, sizeof(Affine) * n, cudaMemcpyHostToDevice, ctx->compute)) != cudaSuccess ||
        (g1 && (e = cudaMemcpyAsync(dbases + n, g1, sizeof(Affine), cudaMemcpyHostToDevice, ctx->compute)) != cudaSuccess) ||
        (e = cudaMemcpyAsync(dbases + (b->n1 - 1), h, sizeof(Affine), cudaMemcpyHostToDevice, ctx->compute)) != cudaSuccess ||
        (e = cudaMemcpyAsync(dinf, inf_host.data(), b->n1, cudaMemcpyHostToDevice, ctx->compute)) != cudaSuccess) {
        ctx->last_error = std::string("sbn_bases_create upload: ") + cudaGetErrorString(e);
        return fail(SBN_ERR_CUDA);
    }
    ctx->h2d += sizeof(Affine) * (b->n1 + b->n_cols) + b->n1;
    k_build_tables<<<(b->n1 + 63) / 64, 64, 0, ctx->compute>>>(dbases, dinf, b->n1, b->c, b->W, b->table);
    ctx->launches++;
    if ((e = cudaGetLastError()) != cudaSuccess || (e = cudaStreamSynchronize(ctx->compute)) != cudaSuccess) {
        ctx->last_error = std::string("k_build_tables: ") + cudaGetErrorString(e);
        return fail(SBN_ERR_CUDA);
    }
    cudaFree(dbases);
    cudaFree(dinf);
    // Digit-multiple tables (small_kernels.cuh): always for a short set (a few ms, <= 4 MiB); for an opening's set (the one
    // created with gens_1's generator, whose commits are single rows and row pairs on the Fiat-Shamir critical path) up to
    // tab_max_mb -- 2 GiB at 8192 generators, built once per generator set in ~0.1 s.
    const size_t tab_bytes = (size_t)kSmallW * b->n_cols * kSmallD * sizeof(Affine);
    if (b->n_cols <= kSmallMaxCols || (b->has_g1 && tab_bytes <= ((size_t)ctx->tab_max_mb << 20))) {
        const size_t entries = (size_t)kSmallW * b->n_cols * kSmallD;
        if ((e = dev_malloc(ctx, &b->small, entries * sizeof(Affine))) != cudaSuccess) {
            ctx->last_error = std::string("sbn_bases_create cudaMalloc: ") + cudaGetErrorString(e);
            return fail(SBN_ERR_OOM);
        }
        k_build_small_table<<<(kSmallW * b->n_cols + 31) / 32, 32, 0, ctx->compute>>>(b->orig, b->n_cols, b->small);
        ctx->launches++;
        if ((e = cudaGetLastError()) != cudaSuccess || (e = cudaStreamSynchronize(ctx->compute)) != cudaSuccess) {
            ctx->last_error = std::string("k_build_small_table: ") + cudaGetErrorString(e);
            return fail(SBN_ERR_CUDA);
        }
    }
    *out = b;
    return SBN_OK;
}

extern "C" int sbn_bases_create(sbn_ctx* ctx, const sbn_g1a* G, const uint8_t* G_inf, size_t n, const sbn_g1a* h,
                                sbn_bases** out) {
    return bases_create(ctx, G, G_inf, n, nullptr, h, out);
}
extern "C" int sbn_bases_create_ext(sbn_ctx* ctx, const sbn_g1a* G, const uint8_t* G_inf, size_t n, const sbn_g1a* g1,
                                    const sbn_g1a* h, sbn_bases** out) {
    if (!g1) return SBN_ERR_ARG;
    return bases_create(ctx, G, G_inf, n, g1, h, out);
}

extern "C" int sbn_bases_destroy(sbn_bases* b) {
    if (!b) return SBN_ERR_ARG;
    {
        std::lock_guard<std::mutex> g(b->ctx->mu);
        cudaSetDevice(b->ctx->device);
        cudaStreamSynchronize(b->ctx->compute);
        if (b->table) cudaFree(b->table);
        if (b->orig) cudaFree(b->orig);
        if (b->small) cudaFree(b->small);
        if (b->mult) cudaFree(b->mult);
        if (b->mult_s) cudaFree(b->mult_s);
        for (uint32_t* p : {b->gptr, b->gcols, b->gbig}) if (p) cudaFree(p);
    }
    delete b;
    return SBN_OK;
}
extern "C" size_t sbn_bases_len(const sbn_bases* b) { return b ? b->n : 0; }
extern "C" int sbn_bases_window_bits(const sbn_bases* b) { return b ? b->c : 0; }
// window width and size of the digit-multiple table many-row commits sum over (0 when none has been built, mult_kernels.cuh)
extern "C" int sbn_bases_mult_table(const sbn_bases* b, int* window_bits, uint64_t* bytes) {
    if (!b) return SBN_ERR_ARG;
    if (window_bits) *window_bits = b->mult ? b->mc : 0;
    if (bytes) *bytes = b->mult ? ((uint64_t)b->mW * b->n1 << (b->mc - 1)) * sizeof(Affine) : 0;
    return SBN_OK;
}

// ------------------------------------------------------------------------------------------------
// commit pipeline
// ------------------------------------------------------------------------------------------------
// Batched-affine rounds for a chunk of `rows` rows.  Measured on B200 (scripts/sweep_sort.py ... ba_rounds): one round
// takes 8-13 % off the accumulation once a chunk holds more than ~8 M list entries (below that the three extra launches
// per round cost what they save); a second round adds ~4 % at >= 4096 generators; a third loses to the padding.
static int ba_rounds_for(const sbn_ctx* ctx, const sbn_bases* b, size_t rows) {
    if (ctx->ba_rounds >= 0) return (int)ctx->ba_rounds;
    const double entries = double(rows) * b->W * b->n1;
    if (entries < 8e6) return 0;
    return (b->n1 >= 4096 && entries >= 32e6) ? 2 : 1;
}

static int task_cap_for(const sbn_ctx* ctx, const sbn_bases* b, int ba) {
    if (ctx->task_cap) return (int)ctx->task_cap;
    double mean = double(b->W) * b->n1 / b->nb;
    if (ba) {    // tasks count points after the batched-affine rounds
        mean = mean / double(1 << ba) + 0.5;
        return std::max(8, std::min(kMaxTaskCap, (int)(2.5 * mean + 0.5)));
    }
    int cap = (int)(2.5 * mean + 0.5);
    return std::max(32, std::min(kMaxTaskCap, cap));
}

static cudaEvent_t get_event(sbn_ctx* ctx, size_t idx) {
    while (ctx->ev_pool.size() <= idx) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
        ctx->ev_pool.push_back(e);
    }
    return ctx->ev_pool[idx];
}

template <int C>
static void launch_sort(const Fr* Z, const Fr* blinds, int R, int n1, int cap, int align_log, uint32_t E, uint32_t max_tasks,
                        uint32_t max_heavy, uint32_t* entries, uint32_t* tstart, Task* tasks, uint32_t* heavy, int rows,
                        cudaStream_t s) {
    // Rows of >= 1024 scalars get one 1024-thread CTA: the same number of resident threads per SM as four 256-thread
    // CTAs, but a quarter of the rows in flight, so the rows being scattered (E * 4 bytes each) stay inside L2 and their
    // 32-byte sectors fill up before they are evicted (2.4x faster at 8192 generators).
    if (n1 >= 1024)
        k_sort_row<C, 1024><<<rows, 1024, 0, s>>>(Z, blinds, R, n1, n1 - 1, 1, cap, align_log, E, max_tasks, max_heavy, entries,
                                                  tstart, tasks, heavy);
    else
        k_sort_row<C, 256><<<rows, 256, 0, s>>>(Z, blinds, R, n1, n1 - 1, 1, cap, align_log, E, max_tasks, max_heavy, entries,
                                                tstart, tasks, heavy);
}

static int dispatch_sort(int c, const Fr* Z, const Fr* blinds, int R, int n1, int cap, int align_log, uint32_t E,
                         uint32_t max_tasks, uint32_t max_heavy, uint32_t* entries, uint32_t* tstart, Task* tasks, uint32_t* heavy,
                         int rows, cudaStream_t s) {
    switch (c) {
#define SBN_CASE(CC) case CC: launch_sort<CC>(Z, blinds, R, n1, cap, align_log, E, max_tasks, max_heavy, entries, tstart, tasks, heavy, rows, s); return SBN_OK;
        SBN_CASE(4) SBN_CASE(5) SBN_CASE(6) SBN_CASE(7) SBN_CASE(8) SBN_CASE(9) SBN_CASE(10) SBN_CASE(11)
        SBN_CASE(12) SBN_CASE(13)
#undef SBN_CASE
        default: return SBN_ERR_UNSUPPORTED;
    }
}

// Per-commit launch plan shared by the stages of one chunk.
struct ChunkPlan {
    uint32_t E;          // row stride of the entry array (padded for aligned bucket starts when ba > 0)
    uint32_t Epts;       // list elements per row that the XYZZ accumulation sees: E >> ba
    uint32_t max_tasks, max_heavy;
    int cap, ba;
};
static ChunkPlan chunk_plan(const sbn_ctx* ctx, const sbn_bases* b, size_t rows) {
    ChunkPlan p;
    p.ba = ba_rounds_for(ctx, b, rows);
    const uint64_t E0 = (uint64_t)b->W * (uint64_t)b->n1;
    const uint64_t A = 1ull << p.ba;
    p.E = (uint32_t)((E0 + (uint64_t)b->nb * (A - 1) + A - 1) / A * A);
    p.Epts = p.E >> p.ba;
    p.cap = task_cap_for(ctx, b, p.ba);
    p.max_tasks = (uint32_t)msm_max_tasks(p.Epts, b->nb, p.cap);
    p.max_heavy = (uint32_t)msm_max_heavy(p.Epts, p.cap);
    return p;
}

// Stage timing: two timing events bracket the launches of one stage on the stream they run on.
struct StageMarks {
    sbn_ctx* ctx;
    size_t& ev_idx;
    std::vector<int>& ev_stage;
    void mark(int stage, cudaStream_t st) {
        cudaEvent_t e = get_event(ctx, ev_idx++);
        if (e) cudaEventRecord(e, st);
        ev_stage.push_back(stage);
    }
};

static int stage_sort(sbn_ctx* ctx, const sbn_bases* b, sbn_ctx::Slot& sl, const Fr* dZ_chunk, const Fr* dblinds_chunk, int rows,
                      int R, cudaStream_t st, StageMarks& m) {
    const ChunkPlan p = chunk_plan(ctx, b, (size_t)rows);
    m.mark(-1, st);
    if (p.ba) SBN_CUDA(ctx, cudaMemsetAsync(sl.entries.p, 0xff, (size_t)rows * p.E * sizeof(uint32_t), st));   // NULL padding
    if (b->dedup) {     // scalars of equal generators are summed first; the blind joins the group of h
        k_aggregate_rows<<<rows, kAggThreads, 0, st>>>(dZ_chunk, dblinds_chunk, R, b->n_cols, b->gptr, b->gcols, b->n1, b->gbig,
                                                       b->n_big, (Fr*)sl.zagg.p);
        ctx->launches += 1;
        dZ_chunk = (const Fr*)sl.zagg.p;
        dblinds_chunk = nullptr;
        R = b->n1;
    }
    SBN_TRY(dispatch_sort(b->c, dZ_chunk, dblinds_chunk, R, b->n1, p.cap, p.ba, p.E, p.max_tasks, p.max_heavy,
                          (uint32_t*)sl.entries.p, (uint32_t*)sl.tstart.p, (Task*)sl.tasks.p, (uint32_t*)sl.heavy.p, rows, st));
    m.mark(0, st);
    ctx->launches += 1;
    SBN_CUDA(ctx, cudaGetLastError());
    return SBN_OK;
}

static int ba_pairs_per_thread(const sbn_ctx* ctx, size_t npairs) {
    if (ctx->ba_batch) return (int)ctx->ba_batch;
    return (int)std::max<size_t>(4, std::min<size_t>(32, npairs / 150000));   // keep >= ~150k threads in a round
}

// Pairs per thread of a round of the tabulated-sum path.  Both passes of a round hold kBatSlots = 148 SMs x 4 resident
// blocks (k_bat_prefix: 58 registers; k_bat_finish at its 64-register target); the grid is sized to fill a whole number of
// such waves -- a 1.3-wave grid leaves a third of the machine idle for the length of a block -- with at most ~20 pairs per
// thread (longer chains lose more to the tail than they save on the per-thread inversion bookkeeping) and at least 4.
static int bat_pairs_per_thread(const sbn_ctx* ctx, size_t npairs) {
    if (ctx->ba_batch) return (int)ctx->ba_batch;
    const size_t wave = (size_t)148 * (ctx->ba_minb == 4 ? 4 : 3) * kBaThreads;     // pairs of one wave at one pair per thread
    const size_t waves = std::max<size_t>(1, (npairs + wave * 20 - 1) / (wave * 20));
    const size_t B = (npairs + wave * waves - 1) / (wave * waves);
    return (int)std::max<size_t>(4, B);
}

static int stage_accumulate(sbn_ctx* ctx, const sbn_bases* b, sbn_ctx::Slot& sl, int rows, cudaStream_t st, StageMarks& m) {
    const ChunkPlan p = chunk_plan(ctx, b, (size_t)rows);
    const size_t threads = (size_t)rows * p.max_tasks;
    const unsigned acc_blocks = (unsigned)((threads + kAccThreads - 1) / kAccThreads);
    m.mark(-1, st);
    if (p.ba == 0) {
        k_accumulate<<<acc_blocks, kAccThreads, 0, st>>>(b->table, (const uint32_t*)sl.entries.p, (const uint32_t*)sl.tstart.p,
                                                         (const Task*)sl.tasks.p, (XYZZ*)sl.partials.p, rows, b->nb, p.E,
                                                         p.max_tasks);
        ctx->launches += 1;
    } else {
        // batched-affine rounds over the flat pair arrays of the chunk (rows concatenate: p.E is a multiple of 2^ba)
        const Affine* in = nullptr;
        for (int k = 0; k < p.ba; k++) {
            const size_t npairs = ((size_t)rows * p.E) >> (k + 1);
            const int B = ba_pairs_per_thread(ctx, npairs);
            const unsigned blocks = (unsigned)((npairs + (size_t)kBaThreads * B - 1) / ((size_t)kBaThreads * B));
            const size_t nwarps = (size_t)blocks * kBaThreads / 32;
            Affine* out = (Affine*)sl.pts[k].p;
            if (k == 0) {
                k_ba_prefix<true><<<blocks, kBaThreads, 0, st>>>((const uint32_t*)sl.entries.p, b->table, nullptr, npairs, B,
                                                                 (Fq*)sl.prefix.p, (Fq*)sl.other.p, (Fq*)sl.wtot.p);
                k_ba_invert<<<(unsigned)((nwarps + 63) / 64), 64, 0, st>>>((const Fq*)sl.wtot.p, nwarps, (Fq*)sl.winv.p);
                k_ba_finish<true><<<blocks, kBaThreads, 0, st>>>((const uint32_t*)sl.entries.p, b->table, nullptr, npairs, B,
                                                                 (const Fq*)sl.prefix.p, (const Fq*)sl.other.p,
                                                                 (const Fq*)sl.winv.p, out);
            } else {
                k_ba_prefix<false><<<blocks, kBaThreads, 0, st>>>(nullptr, nullptr, in, npairs, B, (Fq*)sl.prefix.p,
                                                                  (Fq*)sl.other.p, (Fq*)sl.wtot.p);
                k_ba_invert<<<(unsigned)((nwarps + 63) / 64), 64, 0, st>>>((const Fq*)sl.wtot.p, nwarps, (Fq*)sl.winv.p);
                k_ba_finish<false><<<blocks, kBaThreads, 0, st>>>(nullptr, nullptr, in, npairs, B, (const Fq*)sl.prefix.p,
                                                                  (const Fq*)sl.other.p, (const Fq*)sl.winv.p, out);
            }
            in = out;
            ctx->launches += 3;
        }
        k_accumulate_pts<<<acc_blocks, kAccThreads, 0, st>>>(in, p.Epts, (const uint32_t*)sl.tstart.p, (const Task*)sl.tasks.p,
                                                             (XYZZ*)sl.partials.p, rows, b->nb, p.max_tasks);
        ctx->launches += 1;
    }
    m.mark(1, st);
    SBN_CUDA(ctx, cudaGetLastError());
    return SBN_OK;
}

// Buckets per leaf thread (m) of the two-level reduction, by a small cost model fitted on B200: the leaf kernel is the
// slower of its dependency chain (2m additions, ~7 us each when a warp runs alone) and its share of the multiplier
// (28 Montgomery products per bucket at ~6.9e10 /s, ~75 % efficient); the top kernel is a pure dependency chain of
// 3e + 12 additions (e = pairs per lane, ~8.4 us each).
static int reduce_leaf_m(const sbn_ctx* ctx, int nb, int rows) {
    const int m_min = std::max(4, nb / 256);
    if (ctx->leaf_m) return (int)std::min<long>(nb, std::max<long>(ctx->leaf_m, m_min));
    int best = m_min;
    double best_t = 1e300;
    for (int m = m_min; m <= std::min(nb, 32); m *= 2) {
        const double leaf = std::max(2.0 * m * 7e-6, double(rows) * nb * 28.0 / 6.9e10 / 0.75);
        const int tpr = nb / m, e = tpr >= 32 ? tpr / 32 : 1;
        const double top = (3.0 * e + 12.0) * 8.4e-6;
        if (leaf + top < best_t) { best_t = leaf + top; best = m; }
    }
    return best;
}

static int stage_reduce(sbn_ctx* ctx, const sbn_bases* b, sbn_ctx::Slot& sl, int rows, XYZZ* totals_chunk, cudaStream_t st,
                        StageMarks& mk) {
    const ChunkPlan p = chunk_plan(ctx, b, (size_t)rows);
    XYZZ* partials = (XYZZ*)sl.partials.p;
    const uint32_t* tstart = (const uint32_t*)sl.tstart.p;
    mk.mark(-1, st);
    {   // fold the partials of split buckets; persistent grid, one warp per row at a time
        const int warps_needed = rows;
        const int blocks = std::max(1, std::min(148 * 4, (warps_needed * 32 + kHeavyThreads - 1) / kHeavyThreads));
        k_combine_heavy<<<blocks, kHeavyThreads, 0, st>>>(partials, tstart, (const uint32_t*)sl.heavy.p, rows, b->nb, p.max_tasks,
                                                          p.max_heavy);
    }
    {   // two-level reduction
        const int m = reduce_leaf_m(ctx, b->nb, rows);
        const int tpr = b->nb / m;
        int log_m = 0;
        while ((1 << log_m) < m) log_m++;
        const size_t threads = (size_t)rows * tpr;
        k_reduce_leaf<<<(unsigned)((threads + kLeafThreads - 1) / kLeafThreads), kLeafThreads, 0, st>>>(
            partials, tstart, rows, b->nb, m, p.max_tasks, (LeafPair*)sl.pairs.p);
        k_reduce_top<<<(rows * 32 + kTopThreads - 1) / kTopThreads, kTopThreads, 0, st>>>((const LeafPair*)sl.pairs.p, rows, tpr,
                                                                                         log_m, totals_chunk);
    }
    mk.mark(2, st);
    ctx->launches += 3;
    SBN_CUDA(ctx, cudaGetLastError());
    return SBN_OK;
}

// Rows per pipeline chunk.  Measured on B200 (scripts/sweep_dev.py, sweep_e2e.py): chunks of 1024 rows are as fast as
// any; below ~512 rows the per-chunk launches and the tail of each accumulation grid start to cost.  Chunks bound the
// workspace (entries + partial sums: ~0.5 MB per row at 1024 generators, ~5 MB per row at 8192) and let the host path
// overlap the H2D copy of chunk i+1 with the kernels of chunk i.
static size_t commit_chunk_rows(const sbn_ctx* ctx, size_t L) {
    if (ctx->chunk_rows > 0) return std::min<size_t>(L, (size_t)ctx->chunk_rows);
    return std::min<size_t>(L, 1024);
}

static int ensure_commit_workspace(sbn_ctx* ctx, const sbn_bases* b, size_t chunk, size_t L) {
    // a short chunk may run without the batched-affine rounds: size for both plans
    const ChunkPlan plan = chunk_plan(ctx, b, chunk), plain = chunk_plan(ctx, b, 1);
    const size_t E = std::max(plan.E, plain.E);
    const size_t max_tasks = std::max(plan.max_tasks, plain.max_tasks);
    const size_t max_heavy = std::max(plan.max_heavy, plain.max_heavy);
    if (max_tasks >= (1u << 24)) return SBN_ERR_SHAPE;
    const size_t nslots = (L > chunk || L > chunk / 8) ? 2 : 1;    // the host path may add a short first chunk
    for (size_t i = 0; i < nslots; i++) {
        auto& sl = ctx->slots[i];
        SBN_TRY(ensure(ctx, sl.entries, chunk * E * sizeof(uint32_t)));
        SBN_TRY(ensure(ctx, sl.tstart, chunk * (b->nb + 1) * sizeof(uint32_t)));
        SBN_TRY(ensure(ctx, sl.tasks, chunk * max_tasks * sizeof(Task)));
        SBN_TRY(ensure(ctx, sl.partials, chunk * max_tasks * sizeof(XYZZ)));
        SBN_TRY(ensure(ctx, sl.heavy, chunk * (max_heavy + 1) * sizeof(uint32_t)));
        SBN_TRY(ensure(ctx, sl.pairs, chunk * (size_t)(b->nb / std::max(4, b->nb / 256) + 1) * sizeof(LeafPair)));
        if (b->dedup) SBN_TRY(ensure(ctx, sl.zagg, chunk * (size_t)b->n1 * sizeof(Fr)));
        const int ba = std::max(plan.ba, plain.ba);
        if (ba) {
            const size_t np1 = chunk * E / 2;                       // pairs of the first round
            for (int k = 0; k < ba; k++) SBN_TRY(ensure(ctx, sl.pts[k], (np1 >> k) * sizeof(Affine)));
            SBN_TRY(ensure(ctx, sl.prefix, np1 * sizeof(Fq)));
            const size_t nthreads = np1 / 4 + 2 * kBaThreads;       // >= 4 pairs per thread
            SBN_TRY(ensure(ctx, sl.other, nthreads * sizeof(Fq)));
            SBN_TRY(ensure(ctx, sl.wtot, (nthreads / 32 + 1) * sizeof(Fq)));
            SBN_TRY(ensure(ctx, sl.winv, (nthreads / 32 + 1) * sizeof(Fq)));
        }
    }
    SBN_TRY(ensure(ctx, ctx->totals, L * sizeof(XYZZ)));
    return SBN_OK;
}

static void collect_profile(sbn_ctx* ctx, const std::vector<int>& ev_stage) {
    for (int i = 0; i < 4; i++) { ctx->prof_ms[i] = 0; ctx->prof_launches[i] = 0; }
    for (size_t i = 1; i < ev_stage.size() && i < ctx->ev_pool.size(); i++) {
        int st = ev_stage[i];
        if (st < 0) continue;
        float ms = 0;
        if (cudaEventElapsedTime(&ms, ctx->ev_pool[i - 1], ctx->ev_pool[i]) == cudaSuccess) {
            ctx->prof_ms[st] += ms;
            ctx->prof_launches[st] += 1;
        }
    }
}

static int check_commit_shape(const sbn_bases* b, size_t L, size_t R) {
    if (L == 0 || R == 0) return SBN_ERR_SHAPE;
    if (R != b->n) return SBN_ERR_SHAPE;            // commitments.rs:146 assert_eq!(gens_n.n, self.len())
    if (L > (1u << 24)) return SBN_ERR_SHAPE;
    return SBN_OK;
}

// The pipeline shared by all entry points.  Rows are cut into chunks that alternate between two workspaces.  Issue order
//   hi : sort(0) sort(1) | wait acc(0) | reduce(0) sort(2) | wait acc(1) | reduce(1) sort(3) | ...
//   lo : wait sort(i) | acc(i)                       (two low-priority streams alternate so tails overlap heads)
// so reduce(i-1) and sort(i+1) run underneath acc(i).  When `host_Z` is given each chunk's H2D copy is issued on the
// copy stream and handed to the sort by an event.  `main` is the stream the caller's inputs are ordered on and on which
// the normalisation runs.
// Few rows over a tabulated generator set: two launches on the caller's stream (small_kernels.cuh).
static int tab_commit(sbn_ctx* ctx, const sbn_bases* b, const Fr* dZ, const Fr* host_Z, size_t L, size_t R, const Fr* dblinds,
                      Affine* dC, uint8_t* dinf, cudaStream_t main, bool normalize) {
    const unsigned nblk = (unsigned)((R + 1 + kTabScalarsPerBlock - 1) / kTabScalarsPerBlock);
    SBN_TRY(ensure(ctx, ctx->tabpart, L * nblk * sizeof(XYZZ)));
    SBN_TRY(ensure(ctx, ctx->totals, L * sizeof(XYZZ)));
    if (host_Z) {
        SBN_CUDA(ctx, cudaMemcpyAsync((void*)dZ, host_Z, L * R * sizeof(Fr), cudaMemcpyHostToDevice, main));
        ctx->h2d += L * R * sizeof(Fr);
    }
    XYZZ* totals = (XYZZ*)ctx->totals.p;
    k_tab_commit_partial<<<dim3(nblk, (unsigned)L), kTabThreads, 0, main>>>(dZ, dblinds, (int)R, b->n_cols, b->small,
                                                                           (XYZZ*)ctx->tabpart.p);
    k_tab_commit_final<<<(unsigned)L, kTabThreads, 0, main>>>((const XYZZ*)ctx->tabpart.p, (int)nblk, totals);
    ctx->launches += 2;
    if (normalize) {
        k_normalize<<<1, 64, 0, main>>>(totals, (int)L, dC, dinf);
        ctx->launches++;
    }
    SBN_CUDA(ctx, cudaGetLastError());
    return SBN_OK;
}

// Digit-multiple table of a commit's generator set (mult_kernels.cuh), built on the first commit of many rows: the largest
// window width whose table fits mult_max_mb, provided the tabulated sum then costs clearly less than the bucket method
// (W_m batched-affine additions of ~6.5 products against W additions of ~8 plus the bucket reduction).
static void mult_try_build(sbn_ctx* ctx, sbn_bases* b) {
    if (ctx->mult_max_mb <= 0) return;     // not a decision about this set: a later, larger budget may still build one
    const double cost_cur = b->W * 8.0 + 2.0 * b->nb / double(b->n1) * 14.0;
    for (int c = kMultMaxBits; c >= 8; c--) {
        const int W = msm_num_windows(c);
        const uint64_t entries = (uint64_t)W * b->n1 << (c - 1);
        if (entries >= (1ull << 31) || entries * sizeof(Affine) > ((uint64_t)ctx->mult_max_mb << 20)) continue;
        if (W * 6.5 >= 0.9 * cost_cur) { b->mult_tried = 1; return; }      // smaller windows only cost more: deliberate "no table"
        if (dev_malloc(ctx, &b->mult, entries * sizeof(Affine)) != cudaSuccess) {
            // The budget allowed the table but the device could not hold it (dev_malloc has already emptied the pool and
            // retried).  Not remembered in mult_tried: the next many-row commit tries again.  Visible to the caller
            // through sbn_ctx_counters2 / sbn_last_cuda_error instead of a silent switch to the slower pipeline.
            b->mult = nullptr;
            ctx->mult_fallbacks++;
            if (++b->mult_fails >= 3) b->mult_tried = 1;      // stop asking for tens of GB on every commit
            ctx->last_error = "digit-multiple table: cudaMalloc of " + std::to_string(entries * sizeof(Affine)) +
                              " bytes failed; this commit runs through the bucket pipeline";
            return;
        }
        const uint64_t threads = (uint64_t)W * b->n1 * ((1u << (c - 1)) / kMultChunk);
        k_mult_fill<<<(unsigned)((threads + 63) / 64), 64, 0, ctx->compute>>>(b->table, b->n1, c, W, b->mult);   // window 0 of the tables = the bases
        ctx->launches++;
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->compute);
        if (e != cudaSuccess) {
            cudaFree(b->mult);
            b->mult = nullptr;
            ctx->mult_fallbacks++;
            ctx->last_error = std::string("digit-multiple table build: ") + cudaGetErrorString(e);
            return;
        }
        b->mc = c;
        b->mW = W;
        b->mult_tried = 1;
        return;
    }
    b->mult_tried = 1;       // no window width fits the budget
}

// Per-row bit lengths of an L x R matrix of device scalars; `rowbits` stays empty when a strided sample of the scalars shows
// none that is small -- every prove-time commit (eq-table values, witnesses) leaves after one tiny launch and a 4-byte copy.
static constexpr int kSmallScalarBits = 64;
static int scan_row_bits(sbn_ctx* ctx, const Fr* dZ, size_t L, size_t R, cudaStream_t st, std::vector<uint32_t>& rowbits) {
    rowbits.clear();
    const size_t n = L * R, count = std::min<size_t>(n, 8192), stride = n / count;
    SBN_TRY(ensure(ctx, ctx->scan, (L + 1) * sizeof(uint32_t)));
    uint32_t* d = (uint32_t*)ctx->scan.p;
    uint32_t lo = 0xffffffffu;
    SBN_CUDA(ctx, cudaMemsetAsync(d, 0xff, sizeof(uint32_t), st));
    k_min_bits_sample<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(dZ, stride, count, d);
    ctx->launches++;
    SBN_CUDA(ctx, cudaMemcpyAsync(&lo, d, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    SBN_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->d2h += 4;
    if (lo > (uint32_t)kSmallScalarBits) return SBN_OK;
    k_row_max_bits<<<(unsigned)L, 256, 0, st>>>(dZ, (int)R, d + 1);
    ctx->launches++;
    rowbits.resize(L);
    SBN_CUDA(ctx, cudaMemcpyAsync(rowbits.data(), d + 1, L * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    SBN_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->d2h += L * 4;
    return SBN_OK;
}

// Table of the digit multiples of the first few windows only: Ws windows of cs bits with Ws * cs - 1 >= bits (signed digits:
// the top window must not carry out), the fewest windows whose table fits the budget.  Kept per generator set and rebuilt only
// when a later commit needs more bits than it covers.
static const Affine* mult_small_table(sbn_ctx* ctx, sbn_bases* b, int bits) {
    bits = std::max(bits, 1);
    if (b->mult_s && b->msW * b->msc - 1 >= bits) return b->mult_s;
    for (int Ws = 1; Ws <= 6; Ws++) {
        int cs = (bits + 1 + Ws - 1) / Ws;
        if (cs < 8) cs = 8;                                   // k_mult_fill works in runs of 128 multiples
        if (cs > kMultMaxBits) continue;
        const uint64_t entries = (uint64_t)Ws * b->n1 << (cs - 1);
        if (entries >= (1ull << 31) || entries * sizeof(Affine) > ((uint64_t)ctx->mult_max_mb << 20)) continue;
        if (b->mult && Ws * 2 > b->mW) return nullptr;        // the full table's schedule is not much longer: keep to it
        if (b->mult_s) {
            cudaStreamSynchronize(ctx->compute);
            cudaFree(b->mult_s);
            b->mult_s = nullptr;
        }
        if (dev_malloc(ctx, &b->mult_s, entries * sizeof(Affine)) != cudaSuccess) { b->mult_s = nullptr; return nullptr; }
        const uint64_t threads = (uint64_t)Ws * b->n1 * ((1u << (cs - 1)) / kMultChunk);
        k_mult_fill<<<(unsigned)((threads + 63) / 64), 64, 0, ctx->compute>>>(b->table, b->n1, cs, Ws, b->mult_s);
        ctx->launches++;
        if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(ctx->compute) != cudaSuccess) {
            cudaFree(b->mult_s);
            b->mult_s = nullptr;
            return nullptr;
        }
        b->msc = cs;
        b->msW = Ws;
        return b->mult_s;
    }
    return nullptr;
}

static void launch_sum_rows(sbn_ctx* ctx, int rows, cudaStream_t st, const Fq* px, const Fq* py, uint32_t cnt, uint32_t rp, int nrows,
                            XYZZ* totals) {
    (void)nrows;
    (void)cnt;
    const int wpr = ctx->sum_wpr > 0 ? (int)ctx->sum_wpr : 1;      // more warps per row only add shuffle-tree levels: 2.55 / 2.61 / 2.73 ms at 1 / 2 / 4
    const unsigned rows_per_block = (unsigned)(kMultSumThreads / 32 / wpr);
    const unsigned blocks = ((unsigned)rows + rows_per_block - 1) / rows_per_block;
    if (wpr == 4) k_mult_sum_rows_t<4><<<blocks, kMultSumThreads, 0, st>>>(px, py, cnt, rp, rows, totals);
    else if (wpr == 2) k_mult_sum_rows_t<2><<<blocks, kMultSumThreads, 0, st>>>(px, py, cnt, rp, rows, totals);
    else k_mult_sum_rows_t<1><<<blocks, kMultSumThreads, 0, st>>>(px, py, cnt, rp, rows, totals);
}

// memcpy of a chunk of scalars on four host threads (one thread moves ~10 GB/s; a 16 MB chunk in ~0.5 ms on four)
static void host_copy_mt(void* dst, const void* src, size_t bytes) {
    const int nt = bytes >= (size_t(4) << 20) ? 4 : 1;
    if (nt == 1) { memcpy(dst, src, bytes); return; }
    const size_t part = (bytes / nt + 4095) & ~size_t(4095);
    std::thread th[3];
    for (int t = 1; t < nt; t++) {
        const size_t off = std::min(bytes, part * t), len = std::min(bytes - off, part);
        th[t - 1] = std::thread([=] { memcpy((char*)dst + off, (const char*)src + off, len); });
    }
    memcpy(dst, src, std::min(bytes, part));
    for (int t = 1; t < nt; t++) th[t - 1].join();
}

static int mult_rounds_for(uint32_t used) {
    int r = 1;
    // leave ~400-800 points per row to the XYZZ sum: with its additions inlined the sum kernel is cheap enough that a sixth
    // round (three more launches and a 26 us inversion for 257 instead of 514 points) no longer pays: 2.55 -> 2.52 ms at cfg1
    while (r < 10 && (used >> (r + 1)) >= 384) r++;
    return r;
}

// Many rows as sums of tabulated digit multiples.  Chunks of rows alternate between the two workspaces and the two
// low-priority streams, so that one chunk's inversion launches (a 27 us dependency per round) and kernel tails run under
// the other chunk's additions; a commit that would be one chunk is cut in two for the same reason.  With host scalars the
// copy of chunk i + 1 is issued on the copy stream before the kernels of chunk i.
static int mult_commit(sbn_ctx* ctx, const sbn_bases* b, const Affine* mtable, int mtc, int mtW, const Fr* dZ, const Fr* host_Z,
                       size_t L, size_t R, const Fr* dblinds, Affine* dC, uint8_t* dinf, cudaStream_t main,
                       std::vector<int>& ev_stage, bool normalize) {
    size_t chunk = commit_chunk_rows(ctx, L);
    const size_t ns_max = (size_t)ctx->mult_streams;
    if (ctx->chunk_rows <= 0 && L >= 512 && L <= chunk) chunk = (L + ns_max - 1) / ns_max;    // at least one chunk per stream
    std::vector<size_t> sched;
    {
        size_t done = 0;
        const size_t first = ctx->first_chunk_rows > 0 ? std::min<size_t>((size_t)ctx->first_chunk_rows, chunk) : chunk / 4;
        // a short first chunk gets the kernels started under the rest of the copy -- for a BLOCKING call; asynchronous calls
        // (force_set >= 0) are pipelined against each other by the caller, and equal chunks keep every launch at full size
        const bool short_first = (host_Z && ctx->force_set < 0) || ctx->first_chunk_rows > 0;
        if (short_first && L > first && chunk >= 8 && first > 0) { sched.push_back(first); done = first; }
        while (done < L) { size_t cr = std::min(chunk, L - done); sched.push_back(cr); done += cr; }
    }
    const size_t nchunks = sched.size();
    const int c = mtc, W = mtW;
    const int Rk = b->dedup ? b->n1 : (int)R;              // scalars per row the entries kernel sees
    const uint32_t used = (uint32_t)W * (uint32_t)(Rk + 1);
    int rounds = mult_rounds_for(used);
    if (ctx->mult_rounds > 0) rounds = (int)std::min<long>(ctx->mult_rounds, 12);
    while (rounds > 1 && (used >> rounds) == 0) rounds--;
    const uint32_t stride = (used + (1u << rounds) - 1) >> rounds << rounds;
    const bool tr = ctx->mult_layout != 0;                 // position-major lists: rows padded to a multiple of 32
    const size_t chunk_pad = tr ? (chunk + 31) / 32 * 32 : chunk;
    const size_t np1 = chunk_pad * (size_t)stride / 2;
    const size_t ns = std::max<size_t>(1, std::min(ns_max, nchunks));
    // Two sets of workspaces / streams, taken in turn by consecutive calls: a caller that issues independent commits on two
    // streams of its own (sbn_hyrax_commit_device is asynchronous) gets the tail of one commit -- the short last rounds, the
    // row sums, the normalisation: ~0.3 ms of a mostly idle GPU -- underneath the head of the next.  Within one caller stream
    // nothing changes (stream order).
    const size_t set = ns > 2 ? 0 : (ctx->force_set >= 0 ? (size_t)ctx->force_set : (host_Z ? 0 : (size_t)(ctx->mult_calls++ & 1)));
    ctx->last_was_mult = 1;
    const size_t sb = 2 * set;
    for (size_t k = 0; k < ns; k++) {
        auto& sl = ctx->slots[sb + k];
        SBN_TRY(ensure(ctx, sl.entries, chunk_pad * (size_t)stride * sizeof(uint32_t)));
        SBN_TRY(ensure(ctx, sl.pts[0], np1 * sizeof(Affine)));
        SBN_TRY(ensure(ctx, sl.pts[1], (np1 / 2 + 1) * sizeof(Affine)));
        SBN_TRY(ensure(ctx, sl.prefix, (np1 + np1 / 2 + 1) * sizeof(Fq)));       // fused rounds: this round's and the next one's
        // threads of a round: >= 4 pairs each in the separate-pass layouts; the fused rounds give a thread 2^(rounds - 1) pairs
        // of round 1, which is fewer than 4 for short rows (few windows x few distinct generators)
        const size_t nthreads = np1 / std::min<size_t>(4, size_t(1) << (rounds - 1)) + 2 * kBaThreads;
        SBN_TRY(ensure(ctx, sl.other, 2 * nthreads * sizeof(Fq)));
        SBN_TRY(ensure(ctx, sl.wtot, 2 * (nthreads / 32 + 1) * sizeof(Fq)));
        SBN_TRY(ensure(ctx, sl.winv, 2 * (nthreads / 32 + 1) * sizeof(Fq)));
        if (b->dedup) SBN_TRY(ensure(ctx, sl.zagg, chunk * (size_t)b->n1 * sizeof(Fr)));
    }
    SBN_TRY(ensure(ctx, ctx->mtotals[set], L * sizeof(XYZZ)));
    XYZZ* totals = (XYZZ*)ctx->mtotals[set].p;
    size_t ev_idx = 0;
    StageMarks marks{ctx, ev_idx, ev_stage};
    const size_t sync_base = 3 * nchunks + 8 + set * (nchunks + 8);      // hand-off events live after the profiling events, per set
    if (!get_event(ctx, sync_base + nchunks + 4)) { ctx->last_error = "cudaEventCreate failed"; return SBN_ERR_CUDA; }
    std::vector<size_t> row0(nchunks, 0);
    for (size_t i = 1; i < nchunks; i++) row0[i] = row0[i - 1] + sched[i - 1];
    // Pageable host scalars (a Rust Vec<Scalar>; cudaMemcpyAsync stages them through the driver at ~12 GB/s, synchronously) go
    // through a ring of two pinned buffers per workspace set instead, filled by four host threads: the copy of chunk i + 1 is
    // then a real asynchronous H2D that runs under the kernels of chunk i.
    bool pageable = false;
    if (host_Z) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, host_Z) != cudaSuccess) { cudaGetLastError(); pageable = true; }
        else pageable = at.type == cudaMemoryTypeUnregistered;
    }
    auto issue_copy = [&](size_t ci) -> int {
        const size_t bytes = sched[ci] * R * sizeof(Fr);
        const void* src = host_Z + row0[ci] * R;
        if (pageable) {
            HostBuf& ring = ctx->stage_pin[set][ci & 1];
            if (ring.cap < bytes) {
                if (ring.p) { SBN_CUDA(ctx, cudaDeviceSynchronize()); cudaFreeHost(ring.p); ring.p = nullptr; ring.cap = 0; }
                SBN_CUDA(ctx, cudaHostAlloc(&ring.p, bytes + bytes / 8, cudaHostAllocDefault));
                ring.cap = bytes + bytes / 8;
            }
            // the last H2D out of this slot -- two chunks ago in this call, or a chunk of the asynchronous call that had this
            // workspace set before -- must be done before the host overwrites it
            if (ring.busy) SBN_CUDA(ctx, cudaEventSynchronize(ring.busy));
            host_copy_mt(ring.p, src, bytes);
            src = ring.p;
        }
        SBN_CUDA(ctx, cudaMemcpyAsync((void*)(dZ + row0[ci] * R), src, bytes, cudaMemcpyHostToDevice, ctx->copy));
        ctx->h2d += sched[ci] * R * sizeof(Fr);
        SBN_CUDA(ctx, cudaEventRecord(get_event(ctx, sync_base + ci), ctx->copy));
        if (pageable) {
            HostBuf& ring = ctx->stage_pin[set][ci & 1];
            if (!ring.busy) SBN_CUDA(ctx, cudaEventCreateWithFlags(&ring.busy, cudaEventDisableTiming));
            SBN_CUDA(ctx, cudaEventRecord(ring.busy, ctx->copy));
        }
        return SBN_OK;
    };
    SBN_CUDA(ctx, cudaEventRecord(ctx->fork, main));
    for (size_t k = 0; k < ns; k++) {
        SBN_CUDA(ctx, cudaStreamWaitEvent(ctx->lo[sb + k], ctx->fork, 0));
        // the previous user of this workspace set (both sets when more than two chunks are in flight) may sit on another stream
        for (size_t q = 0; q < 2; q++)
            if ((q == set || ns > 2) && ctx->set_done[q]) SBN_CUDA(ctx, cudaStreamWaitEvent(ctx->lo[sb + k], ctx->set_done[q], 0));
    }
    auto mark_done = [&]() -> int {
        for (size_t q = 0; q < 2; q++) {
            if (q != set && ns <= 2) continue;
            if (!ctx->set_done[q]) SBN_CUDA(ctx, cudaEventCreateWithFlags(&ctx->set_done[q], cudaEventDisableTiming));
            SBN_CUDA(ctx, cudaEventRecord(ctx->set_done[q], main));
        }
        return SBN_OK;
    };
    if (host_Z) {
        SBN_CUDA(ctx, cudaStreamWaitEvent(ctx->copy, ctx->fork, 0));
        SBN_TRY(issue_copy(0));
    }
    for (size_t ci = 0; ci < nchunks; ci++) {
        const int rows = (int)sched[ci];
        auto& sl = ctx->slots[sb + ci % ns];
        cudaStream_t st = ctx->lo[sb + ci % ns];
        if (host_Z) {
            if (ci + 1 < nchunks) SBN_TRY(issue_copy(ci + 1));
            SBN_CUDA(ctx, cudaStreamWaitEvent(st, get_event(ctx, sync_base + ci), 0));
        }
        const Fr* zc = dZ + row0[ci] * R;
        const Fr* bc = dblinds ? dblinds + row0[ci] : nullptr;
        marks.mark(-1, st);
        if (b->dedup) {
            k_aggregate_rows<<<rows, kAggThreads, 0, st>>>(zc, bc, (int)R, b->n_cols, b->gptr, b->gcols, b->n1, b->gbig, b->n_big,
                                                           (Fr*)sl.zagg.p);
            ctx->launches++;
            zc = (const Fr*)sl.zagg.p;
            bc = nullptr;
        }
        if (tr) {
            const uint32_t rp = (uint32_t)((rows + 31) / 32 * 32);
            const size_t ethreads = (size_t)rp * ((size_t)(Rk + 1) + (stride - used));
            k_mult_entries_t<<<(unsigned)((ethreads + 255) / 256), 256, 0, st>>>(zc, bc, Rk, b->n1, c, W, stride, rows, rp,
                                                                                 (uint32_t*)sl.entries.p);
            ctx->launches++;
            marks.mark(0, st);
            if (ctx->mult_layout == 2) {
                // fused rounds: thread = (row, run of B consecutive pair positions), B = 2^(rounds - k) in round k
                const uint32_t npb = stride >> rounds;                           // position blocks = points left per row at the end
                const size_t nwarps = (size_t)npb * (rp / 32);
                const unsigned blocks = (unsigned)((nwarps + kBaThreads / 32 - 1) / (kBaThreads / 32));
                const size_t nthr = (size_t)blocks * kBaThreads;
                Fq* pre[2] = {(Fq*)sl.prefix.p, (Fq*)sl.prefix.p + np1};
                Fq* oth[2] = {(Fq*)sl.other.p, (Fq*)sl.other.p + nthr};
                Fq* tot[2] = {(Fq*)sl.wtot.p, (Fq*)sl.wtot.p + blocks};
                Fq* inv[2] = {(Fq*)sl.winv.p, (Fq*)sl.winv.p + blocks};
                const long ab = ctx->ablate;
                const bool m4 = ctx->ba_minb == 4;
                if (!(ab & 1))
                    k_bat2_prefix1<<<blocks, kBaThreads, 0, st>>>((const uint32_t*)sl.entries.p, mtable, npb, rp, 1 << (rounds - 1), pre[0],
                                                                  oth[0], tot[0]);
                if (!(ab & 4)) k_ba_invert<<<(blocks + 63) / 64, 64, 0, st>>>(tot[0], blocks, inv[0]);
                ctx->launches += 2;
                const Fq *inx = nullptr, *iny = nullptr;
                for (int k = 0; k < rounds; k++) {
                    const int B = 1 << (rounds - 1 - k), cur = k & 1, nxt = cur ^ 1, asc = (k & 1) == 0;
                    const bool emit = k + 1 < rounds;
                    const size_t npairs = ((size_t)rp * stride) >> (k + 1);
                    Fq* outx = (Fq*)sl.pts[k & 1].p;
                    Fq* outy = outx + npairs;
                    void (*fn)(const uint32_t*, const Affine*, const Fq*, const Fq*, uint32_t, uint32_t, int, int, const Fq*, const Fq*,
                               const Fq*, Fq*, Fq*, Fq*, Fq*, Fq*);
                    if (k == 0) fn = emit ? (m4 ? k_bat2_round<true, true, 4> : k_bat2_round<true, true, 3>)
                                          : (m4 ? k_bat2_round<true, false, 4> : k_bat2_round<true, false, 3>);
                    else fn = emit ? (m4 ? k_bat2_round<false, true, 4> : k_bat2_round<false, true, 3>)
                                   : (m4 ? k_bat2_round<false, false, 4> : k_bat2_round<false, false, 3>);
                    if (!(ab & (k == 0 ? 16 : 32)))
                        fn<<<blocks, kBaThreads, 0, st>>>((const uint32_t*)sl.entries.p, mtable, inx, iny, npb, rp, B, asc, pre[cur], oth[cur],
                                                          inv[cur], outx, outy, pre[nxt], oth[nxt], tot[nxt]);
                    ctx->launches++;
                    if (emit) {
                        if (!(ab & 4)) k_ba_invert<<<(blocks + 63) / 64, 64, 0, st>>>(tot[nxt], blocks, inv[nxt]);
                        ctx->launches++;
                    }
                    inx = outx;
                    iny = outy;
                }
                if (!(ctx->ablate & 8))
                    launch_sum_rows(ctx, rows, st, 
                        inx, iny, npb, rp, rows, totals + row0[ci]);
                ctx->launches++;
                marks.mark(1, st);
                SBN_CUDA(ctx, cudaGetLastError());
                continue;
            }
            const Fq *inx = nullptr, *iny = nullptr;
            for (int k = 0; k < rounds; k++) {
                const size_t npairs = ((size_t)rp * stride) >> (k + 1);
                const int B = bat_pairs_per_thread(ctx, npairs);
                const unsigned blocks = (unsigned)((npairs + (size_t)kBaThreads * B - 1) / ((size_t)kBaThreads * B));
                Fq* outx = (Fq*)sl.pts[k & 1].p;
                Fq* outy = outx + npairs;
                const bool pf = ctx->ba_prefetch != 0, m4 = ctx->ba_minb == 4;
                const long ab = ctx->ablate;
                const size_t fsm = (size_t)ctx->finish_smem_kb << 10, psm = (size_t)ctx->prefix_smem_kb << 10;
                if (k == 0) {
                    auto pre = pf ? k_bat_prefix<true, true> : k_bat_prefix<true, false>;
                    if (!(ab & 1))
                        pre<<<blocks, kBaThreads, psm, st>>>((const uint32_t*)sl.entries.p, mtable, nullptr, nullptr, npairs, rp, B,
                                                           (Fq*)sl.prefix.p, (Fq*)sl.other.p, (Fq*)sl.wtot.p);
                    if (!(ab & 4)) k_ba_invert<<<(blocks + 63) / 64, 64, 0, st>>>((const Fq*)sl.wtot.p, blocks, (Fq*)sl.winv.p);
                    auto fin = m4 ? (pf ? k_bat_finish<true, 4, true> : k_bat_finish<true, 4, false>)
                                  : (pf ? k_bat_finish<true, 3, true> : k_bat_finish<true, 3, false>);
                    if (!(ab & 16))
                        fin<<<blocks, kBaThreads, fsm, st>>>((const uint32_t*)sl.entries.p, mtable, nullptr, nullptr, npairs, rp, B,
                                                           (const Fq*)sl.prefix.p, (const Fq*)sl.other.p, (const Fq*)sl.winv.p, outx, outy);
                } else {
                    if (!(ab & 2))
                        k_bat_prefix<false><<<blocks, kBaThreads, psm, st>>>(nullptr, nullptr, inx, iny, npairs, rp, B, (Fq*)sl.prefix.p,
                                                                           (Fq*)sl.other.p, (Fq*)sl.wtot.p);
                    if (!(ab & 4)) k_ba_invert<<<(blocks + 63) / 64, 64, 0, st>>>((const Fq*)sl.wtot.p, blocks, (Fq*)sl.winv.p);
                    auto fin = m4 ? k_bat_finish<false, 4> : k_bat_finish<false, 3>;
                    if (!(ab & 32))
                        fin<<<blocks, kBaThreads, fsm, st>>>(nullptr, nullptr, inx, iny, npairs, rp, B, (const Fq*)sl.prefix.p,
                                                           (const Fq*)sl.other.p, (const Fq*)sl.winv.p, outx, outy);
                }
                inx = outx;
                iny = outy;
                ctx->launches += 3;
            }
            if (!(ctx->ablate & 8))
                launch_sum_rows(ctx, rows, st, 
                    inx, iny, stride >> rounds, rp, rows, totals + row0[ci]);
            ctx->launches++;
            marks.mark(1, st);
            SBN_CUDA(ctx, cudaGetLastError());
            continue;
        }
        const unsigned ethreads = (unsigned)(Rk + 1) + (stride - used);
        k_mult_entries<<<dim3((ethreads + 255) / 256, (unsigned)rows), 256, 0, st>>>(zc, bc, Rk, b->n1, c, W, stride,
                                                                                    (uint32_t*)sl.entries.p);
        ctx->launches++;
        marks.mark(0, st);
        const Affine* in = nullptr;
        for (int k = 0; k < rounds; k++) {
            const size_t npairs = ((size_t)rows * stride) >> (k + 1);
            const int B = ba_pairs_per_thread(ctx, npairs);
            const unsigned blocks = (unsigned)((npairs + (size_t)kBaThreads * B - 1) / ((size_t)kBaThreads * B));
            const size_t nwarps = (size_t)blocks * kBaThreads / 32;
            Affine* out = (Affine*)sl.pts[k & 1].p;
            if (k == 0) {
                k_ba_prefix<true><<<blocks, kBaThreads, 0, st>>>((const uint32_t*)sl.entries.p, mtable, nullptr, npairs, B,
                                                                 (Fq*)sl.prefix.p, (Fq*)sl.other.p, (Fq*)sl.wtot.p);
                k_ba_invert<<<(unsigned)((nwarps + 63) / 64), 64, 0, st>>>((const Fq*)sl.wtot.p, nwarps, (Fq*)sl.winv.p);
                auto fin = ctx->ba_minb == 4 ? (ctx->ba_prefetch ? k_ba_finish<true, 4, true> : k_ba_finish<true, 4, false>)
                                             : (ctx->ba_prefetch ? k_ba_finish<true, 3, true> : k_ba_finish<true, 3, false>);
                fin<<<blocks, kBaThreads, 0, st>>>((const uint32_t*)sl.entries.p, mtable, nullptr, npairs, B, (const Fq*)sl.prefix.p,
                                                   (const Fq*)sl.other.p, (const Fq*)sl.winv.p, out);
            } else {
                k_ba_prefix<false><<<blocks, kBaThreads, 0, st>>>(nullptr, nullptr, in, npairs, B, (Fq*)sl.prefix.p,
                                                                  (Fq*)sl.other.p, (Fq*)sl.wtot.p);
                k_ba_invert<<<(unsigned)((nwarps + 63) / 64), 64, 0, st>>>((const Fq*)sl.wtot.p, nwarps, (Fq*)sl.winv.p);
                auto fin = ctx->ba_minb == 4 ? k_ba_finish<false, 4, false> : k_ba_finish<false, 3, false>;
                fin<<<blocks, kBaThreads, 0, st>>>(nullptr, nullptr, in, npairs, B, (const Fq*)sl.prefix.p, (const Fq*)sl.other.p,
                                                   (const Fq*)sl.winv.p, out);
            }
            in = out;
            ctx->launches += 3;
        }
        k_mult_sum_rows<<<(unsigned)((rows * 32 + kMultSumThreads - 1) / kMultSumThreads), kMultSumThreads, 0, st>>>(
            in, stride >> rounds, rows, totals + row0[ci]);
        ctx->launches++;
        marks.mark(1, st);
        SBN_CUDA(ctx, cudaGetLastError());
    }
    for (size_t k = 0; k < ns; k++) {
        cudaEvent_t e = get_event(ctx, sync_base + nchunks + k);
        SBN_CUDA(ctx, cudaEventRecord(e, ctx->lo[sb + k]));
        SBN_CUDA(ctx, cudaStreamWaitEvent(main, e, 0));
    }
    if (!normalize) return mark_done();
    marks.mark(-1, main);
    k_normalize<<<(unsigned)((L + 63) / 64), 64, 0, main>>>(totals, (int)L, dC, dinf);
    ctx->launches++;
    marks.mark(3, main);
    SBN_CUDA(ctx, cudaGetLastError());
    return mark_done();
}

static int run_commit_inner(sbn_ctx* ctx, const sbn_bases* b, const Fr* dZ, const Fr* host_Z, size_t L, size_t R,
                            const Fr* dblinds, Affine* dC, uint8_t* dinf, cudaStream_t main, std::vector<int>& ev_stage,
                            bool normalize);
static int run_commit_general(sbn_ctx* ctx, const sbn_bases* b, const Fr* dZ, const Fr* host_Z, size_t L, size_t R,
                              const Fr* dblinds, Affine* dC, uint8_t* dinf, cudaStream_t main, std::vector<int>& ev_stage,
                              bool normalize);
// Every commit goes through here.  The pipeline forks `main` into the copy / hi / lo streams and joins them at the very
// end; an early error return skips that join, so the streams are drained here before the caller releases or reuses the
// buffers they may still be reading (the pool hands buffers out again in `compute` order only).
static int run_commit(sbn_ctx* ctx, const sbn_bases* b, const Fr* dZ, const Fr* host_Z, size_t L, size_t R,
                      const Fr* dblinds, Affine* dC, uint8_t* dinf, cudaStream_t main, std::vector<int>& ev_stage,
                      bool normalize = true) {
    const int rc = run_commit_inner(ctx, b, dZ, host_Z, L, R, dblinds, dC, dinf, main, ev_stage, normalize);
    if (rc != SBN_OK) {
        cudaDeviceSynchronize();
        cudaGetLastError();
    }
    return rc;
}
static int run_commit_inner(sbn_ctx* ctx, const sbn_bases* b, const Fr* dZ, const Fr* host_Z, size_t L, size_t R,
                            const Fr* dblinds, Affine* dC, uint8_t* dinf, cudaStream_t main, std::vector<int>& ev_stage,
                            bool normalize) {
    // Small scalars (no blinds, scalars already on the device).  The rows are classified by the bit length of their largest
    // scalar; runs of at least mult_min_rows small rows are committed over a table of just the few windows that cover them,
    // the other runs through the general path -- comb_ops is four fifths addresses and timestamps below 2^21 and one fifth
    // matrix coefficients of any size (sparse_mlpoly_full.rs:176-196).
    if (ctx->small_scalar_path && !dblinds && !host_Z && !b->has_g1 && ctx->mult_max_mb > 0 && L >= (size_t)ctx->mult_min_rows &&
        L * R >= (size_t(1) << 16)) {
        std::vector<uint32_t> rowbits;
        SBN_TRY(scan_row_bits(ctx, dZ, L, R, main, rowbits));
        if (!rowbits.empty()) {
            int extra = 0;
            while ((1 << extra) < b->max_group) extra++;                      // merged generators add their scalars
            std::vector<uint8_t> small(L);
            for (size_t i = 0; i < L; i++) small[i] = rowbits[i] + (uint32_t)extra <= (uint32_t)kSmallScalarBits;
            struct Run { size_t row0, n; bool small; };
            std::vector<Run> runs;
            for (size_t i = 0; i < L;) {
                size_t j = i;
                while (j < L && small[j] == small[i]) j++;
                bool sm = small[i] && (j - i) >= (size_t)ctx->mult_min_rows;   // short small runs join their neighbours
                if (!runs.empty() && runs.back().small == sm) runs.back().n += j - i;
                else runs.push_back({i, j - i, sm});
                i = j;
            }
            uint32_t bits = 0;
            bool any = false;
            for (auto& r : runs)
                if (r.small) { any = true; for (size_t i = r.row0; i < r.row0 + r.n; i++) bits = std::max(bits, rowbits[i]); }
            const Affine* t = any ? mult_small_table(ctx, const_cast<sbn_bases*>(b), (int)bits + extra) : nullptr;
            if (t) {
                ctx->small_scalar_hits++;
                for (auto& r : runs) {
                    const Fr* z = dZ + r.row0 * R;
                    // every run re-uses the profiling events from index 0: its marks go to a scratch list, and a commit that
                    // was split into runs reports no stage profile (ev_stage stays empty)
                    std::vector<int> scratch_marks;
                    if (r.small)
                        SBN_TRY(mult_commit(ctx, b, t, b->msc, b->msW, z, nullptr, r.n, R, nullptr, dC + r.row0, dinf + r.row0, main,
                                            scratch_marks, normalize));
                    else
                        SBN_TRY(run_commit_general(ctx, b, z, nullptr, r.n, R, nullptr, dC + r.row0, dinf + r.row0, main,
                                                   scratch_marks, normalize));
                }
                return SBN_OK;
            }
        }
    }
    return run_commit_general(ctx, b, dZ, host_Z, L, R, dblinds, dC, dinf, main, ev_stage, normalize);
}

static int run_commit_general(sbn_ctx* ctx, const sbn_bases* b, const Fr* dZ, const Fr* host_Z, size_t L, size_t R,
                              const Fr* dblinds, Affine* dC, uint8_t* dinf, cudaStream_t main, std::vector<int>& ev_stage,
                              bool normalize) {
    if (L >= (size_t)ctx->mult_min_rows && !b->has_g1 && ctx->mult_max_mb > 0) {
        if (!b->mult_tried) mult_try_build(ctx, const_cast<sbn_bases*>(b));
        if (b->mult) return mult_commit(ctx, b, b->mult, b->mc, b->mW, dZ, host_Z, L, R, dblinds, dC, dinf, main, ev_stage, normalize);
    }
    if (b->small && ctx->small_commit_path && L <= (size_t)kTabMaxRows && R + 1 <= (size_t)b->n_cols)
        return tab_commit(ctx, b, dZ, host_Z, L, R, dblinds, dC, dinf, main, normalize);
    const size_t chunk = commit_chunk_rows(ctx, L);
    // chunk schedule: equal chunks; when the scalars come from the host the first chunk is a short one (an eighth: 4.85 ms
    // end to end at 1024 x 1024 against 5.06 with a quarter and 5.08 with none) so the kernels start after a short copy and
    // the remaining copies hide behind them
    std::vector<size_t> sched;
    {
        size_t done = 0;
        const size_t first = ctx->first_chunk_rows > 0 ? std::min<size_t>((size_t)ctx->first_chunk_rows, chunk) : chunk / 8;
        if (host_Z && L > first && chunk >= 8 && first > 0) { sched.push_back(first); done = first; }
        while (done < L) { size_t c = std::min(chunk, L - done); sched.push_back(c); done += c; }
    }
    const size_t nchunks = sched.size();
    std::vector<size_t> row0(nchunks, 0);
    for (size_t i = 1; i < nchunks; i++) row0[i] = row0[i - 1] + sched[i - 1];
    XYZZ* totals = (XYZZ*)ctx->totals.p;
    size_t ev_idx = 0;
    StageMarks marks{ctx, ev_idx, ev_stage};
    const size_t sync_base = 6 * nchunks + 8;       // hand-off events live after the profiling events: 3 per chunk
    auto sync_event = [&](size_t ci, int kind) { return get_event(ctx, sync_base + 3 * ci + kind); };   // 0 copied, 1 sorted, 2 accumulated
    if (!sync_event(nchunks, 0)) { ctx->last_error = "cudaEventCreate failed"; return SBN_ERR_CUDA; }
    SBN_CUDA(ctx, cudaEventRecord(ctx->fork, main));
    for (cudaStream_t st : {ctx->hi, ctx->lo[0], ctx->lo[1]}) SBN_CUDA(ctx, cudaStreamWaitEvent(st, ctx->fork, 0));
    if (host_Z) SBN_CUDA(ctx, cudaStreamWaitEvent(ctx->copy, ctx->fork, 0));

    auto issue_sort = [&](size_t ci) -> int {
        auto& sl = ctx->slots[ci & 1];
        const int rows = (int)sched[ci];
        if (host_Z) {
            SBN_CUDA(ctx, cudaMemcpyAsync((void*)(dZ + row0[ci] * R), host_Z + row0[ci] * R, (size_t)rows * R * sizeof(Fr),
                                          cudaMemcpyHostToDevice, ctx->copy));
            ctx->h2d += (size_t)rows * R * sizeof(Fr);
            SBN_CUDA(ctx, cudaEventRecord(sync_event(ci, 0), ctx->copy));
            SBN_CUDA(ctx, cudaStreamWaitEvent(ctx->hi, sync_event(ci, 0), 0));
        }
        SBN_TRY(stage_sort(ctx, b, sl, dZ + row0[ci] * R, dblinds ? dblinds + row0[ci] : nullptr, rows, (int)R, ctx->hi, marks));
        SBN_CUDA(ctx, cudaEventRecord(sync_event(ci, 1), ctx->hi));
        return SBN_OK;
    };
    SBN_TRY(issue_sort(0));
    if (nchunks > 1) SBN_TRY(issue_sort(1));
    for (size_t ci = 0; ci < nchunks; ci++) {
        auto& sl = ctx->slots[ci & 1];
        const int rows = (int)sched[ci];
        cudaStream_t lo = ctx->lo[ci & 1];
        SBN_CUDA(ctx, cudaStreamWaitEvent(lo, sync_event(ci, 1), 0));
        SBN_TRY(stage_accumulate(ctx, b, sl, rows, lo, marks));
        SBN_CUDA(ctx, cudaEventRecord(sync_event(ci, 2), lo));
        SBN_CUDA(ctx, cudaStreamWaitEvent(ctx->hi, sync_event(ci, 2), 0));
        SBN_TRY(stage_reduce(ctx, b, sl, rows, totals + row0[ci], ctx->hi, marks));
        if (ci + 2 < nchunks) SBN_TRY(issue_sort(ci + 2));
    }
    SBN_CUDA(ctx, cudaEventRecord(ctx->join_hi, ctx->hi));     // hi has waited on every accumulation
    SBN_CUDA(ctx, cudaStreamWaitEvent(main, ctx->join_hi, 0));
    if (!normalize) return SBN_OK;     // caller consumes the XYZZ row totals in ctx->totals
    marks.mark(-1, main);
    k_normalize<<<(unsigned)((L + 63) / 64), 64, 0, main>>>(totals, (int)L, dC, dinf);
    ctx->launches++;
    marks.mark(3, main);
    SBN_CUDA(ctx, cudaGetLastError());
    return SBN_OK;
}

extern "C" int sbn_hyrax_commit_device(sbn_ctx* ctx, const sbn_bases* b, const void* dZ, size_t L, size_t R,
                                       const void* dblinds, void* dC_out, void* dinf_out, void* stream_) {
    if (!ctx || !b || !dZ || !dC_out || b->ctx != ctx) return SBN_ERR_ARG;
    SBN_TRY(check_commit_shape(b, L, R));
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_ENTER(ctx);
    cudaStream_t stream = stream_ ? (cudaStream_t)stream_ : ctx->compute;
    const size_t chunk = commit_chunk_rows(ctx, L);
    SBN_TRY(ensure_commit_workspace(ctx, b, chunk, L));
    std::vector<int> ev_stage;
    SBN_TRY(run_commit(ctx, b, (const Fr*)dZ, nullptr, L, R, (const Fr*)dblinds, (Affine*)dC_out, (uint8_t*)dinf_out,
                       stream, ev_stage));
    if (!stream_) {   // context stream: resolve the stage timings now
        SBN_CUDA(ctx, cudaStreamSynchronize(stream));
        collect_profile(ctx, ev_stage);
    } else {          // caller's stream stays asynchronous; no profile for this call
        for (int i = 0; i < 4; i++) { ctx->prof_ms[i] = 0; ctx->prof_launches[i] = 0; }
    }
    return SBN_OK;
}

// Short generator set, few rows: one launch over the tabulated digit multiples, operands through mapped pinned memory.
// XYZZ -> affine on the host (the library's host-side field code: 4 x 64-bit limbs, safegcd inverse): for the one to four
// points of a short commitment or a bullet round, where the device-side normalisation is a 30 us chain on one warp
static void host_normalize(const XYZZ& v, sbn_g1a* out, uint8_t* inf) {
    const Affine a = xyzz_to_affine<MulInline>(v);
    memcpy(out, &a, sizeof(Affine));
    *inf = v.is_identity() ? 1 : 0;
}

static int small_commit(sbn_ctx* ctx, const sbn_bases* b, const sbn_fr* Z, size_t L, size_t R, const sbn_fr* blinds,
                        sbn_g1a* C_out, uint8_t* inf_out) {
    const size_t off_blind = (size_t)kSmallMaxRows * kSmallMaxCols * sizeof(Fr), off_out = off_blind + kSmallMaxRows * sizeof(Fr),
                 off_inf = off_out + kSmallMaxRows * sizeof(Affine), off_raw = (off_inf + kSmallMaxRows + 15) & ~size_t(15),
                 total = off_raw + kSmallMaxRows * sizeof(XYZZ);
    const bool host_norm = ctx->host_normalize != 0;
    if (!ctx->small_pin) {
        SBN_CUDA(ctx, cudaHostAlloc((void**)&ctx->small_pin, total, cudaHostAllocMapped));
        SBN_CUDA(ctx, cudaHostGetDevicePointer((void**)&ctx->small_pin_dev, ctx->small_pin, 0));
    }
    memcpy(ctx->small_pin, Z, L * R * sizeof(Fr));
    if (blinds) memcpy(ctx->small_pin + off_blind, blinds, L * sizeof(Fr));
    uint8_t* d = ctx->small_pin_dev;
    k_small_commit<<<(unsigned)L, 32 * ((unsigned)R + 1), 0, ctx->compute>>>((const Fr*)d, blinds ? (const Fr*)(d + off_blind) : nullptr,
                                                                             (int)R, b->n_cols, b->small, (Affine*)(d + off_out),
                                                                             d + off_inf, host_norm ? (XYZZ*)(d + off_raw) : nullptr);
    ctx->launches += 1;
    SBN_CUDA(ctx, cudaGetLastError());
    SBN_CUDA(ctx, cudaStreamSynchronize(ctx->compute));
    if (host_norm) {
        for (size_t i = 0; i < L; i++) {
            XYZZ v;
            memcpy(&v, ctx->small_pin + off_raw + i * sizeof(XYZZ), sizeof(XYZZ));
            host_normalize(v, C_out + i, inf_out + i);
        }
    } else {
        memcpy(C_out, ctx->small_pin + off_out, L * sizeof(Affine));
        memcpy(inf_out, ctx->small_pin + off_inf, L);
    }
    ctx->h2d += L * R * sizeof(Fr) + (blinds ? L * sizeof(Fr) : 0);
    ctx->d2h += L * (sizeof(Affine) + 1);
    for (int i = 0; i < 4; i++) { ctx->prof_ms[i] = 0; ctx->prof_launches[i] = 0; }
    return SBN_OK;
}

// Host-pointer commit.  `async_stream` == nullptr: the synchronous call of the ABI (results are in C_out / inf_out on return).
// Otherwise everything -- the chunked H2D copies, the kernels, the D2H copy of the commitments -- is ordered on the caller's
// stream and the call returns at once; the staging buffers come in two sets taken in turn, like the workspaces, so that two
// commits issued on two streams overlap: the copy of one under the kernels of the other.
static int hyrax_commit_host(sbn_ctx* ctx, const sbn_bases* b, const sbn_fr* Z, size_t L, size_t R, const sbn_fr* blinds,
                             sbn_g1a* C_out, uint8_t* inf_out, cudaStream_t async_stream) {
    if (!ctx || !b || !Z || !C_out || !inf_out || b->ctx != ctx) return SBN_ERR_ARG;
    SBN_TRY(check_commit_shape(b, L, R));
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_ENTER(ctx);
    if (b->small && b->n_cols <= kSmallMaxCols && ctx->small_commit_path && L <= (size_t)kSmallMaxRows) {
        if (async_stream) SBN_CUDA(ctx, cudaStreamSynchronize(async_stream));      // the short path is synchronous on the context stream
        return small_commit(ctx, b, Z, L, R, blinds, C_out, inf_out);
    }
    const bool async = async_stream != nullptr;
    // asynchronous calls take the three staging sets and the two workspace sets in turn
    const int hset = async ? (int)(ctx->host_calls % 3) : 0, wset = async ? (int)(ctx->host_calls & 1) : 0;
    if (async) ctx->host_calls++;
    cudaStream_t main = async ? async_stream : ctx->compute;
    if (async && ctx->hset_done[hset]) SBN_CUDA(ctx, cudaStreamWaitEvent(main, ctx->hset_done[hset], 0));
    DevBuf &bZ = async ? ctx->hZ[hset] : ctx->dZ, &bC = async ? ctx->hC[hset] : ctx->dC, &bI = async ? ctx->hI[hset] : ctx->dinf,
           &bB = async ? ctx->hB[hset] : ctx->dblinds;
    const size_t chunk = commit_chunk_rows(ctx, L);
    SBN_TRY(ensure_commit_workspace(ctx, b, chunk, L));
    SBN_TRY(ensure(ctx, bZ, L * R * sizeof(Fr)));
    SBN_TRY(ensure(ctx, bC, L * sizeof(Affine)));
    SBN_TRY(ensure(ctx, bI, L));
    Fr* dbl = nullptr;
    if (blinds) {
        SBN_TRY(ensure(ctx, bB, L * sizeof(Fr)));
        dbl = (Fr*)bB.p;
        SBN_CUDA(ctx, cudaMemcpyAsync(dbl, blinds, L * sizeof(Fr), cudaMemcpyHostToDevice, main));
        ctx->h2d += L * sizeof(Fr);
    }
    // Host scalars that look small (a strided sample on the host) are uploaded in one piece, so that the device-side row
    // classification of the small-scalar schedule sees them; everything else streams chunk by chunk under the kernels.
    const Fr* host_Z = (const Fr*)Z;
    if (ctx->small_scalar_path && !blinds && L >= (size_t)ctx->mult_min_rows && L * R >= (size_t(1) << 16) && ctx->mult_max_mb > 0 &&
        !b->has_g1) {
        const size_t n = L * R, count = 256, stride = n / count;
        bool any_small = false;
        for (size_t i = 0; i < count && !any_small; i++) {
            Fr v;
            memcpy(v.l, &Z[i * stride], 32);
            v = fp_from_mont(v);
            any_small = (v.l[2] | v.l[3] | v.l[4] | v.l[5] | v.l[6] | v.l[7]) == 0;
        }
        if (any_small) {
            SBN_CUDA(ctx, cudaMemcpyAsync(bZ.p, Z, n * sizeof(Fr), cudaMemcpyHostToDevice, main));
            ctx->h2d += n * sizeof(Fr);
            host_Z = nullptr;
        }
    }
    std::vector<int> ev_stage;
    ctx->force_set = async ? wset : -1;
    ctx->last_was_mult = 0;
    const int rc = run_commit(ctx, b, (const Fr*)bZ.p, host_Z, L, R, dbl, (Affine*)bC.p, (uint8_t*)bI.p, main, ev_stage);
    ctx->force_set = -1;
    SBN_TRY(rc);
    // A D2H copy into PAGEABLE memory blocks the host until everything before it on the stream has run -- an asynchronous call
    // with a Rust Vec as its output would wait for its own commit.  Such outputs land in a pinned buffer of the staging set and
    // a host function, in stream order, moves them on; the caller's stream synchronisation covers it.
    bool out_pageable = false;
    if (async) {
        cudaPointerAttributes at;
        for (const void* q : {(const void*)C_out, (const void*)inf_out}) {
            if (cudaPointerGetAttributes(&at, q) != cudaSuccess) { cudaGetLastError(); out_pageable = true; }
            else if (at.type == cudaMemoryTypeUnregistered) out_pageable = true;
        }
    }
    if (out_pageable) {
        HostBuf& oc = ctx->out_pin[hset];
        const size_t need = L * sizeof(Affine) + L;
        if (oc.cap < need) {
            if (oc.p) { SBN_CUDA(ctx, cudaDeviceSynchronize()); cudaFreeHost(oc.p); oc.p = nullptr; oc.cap = 0; }
            SBN_CUDA(ctx, cudaHostAlloc(&oc.p, need, cudaHostAllocDefault));
            oc.cap = need;
        }
        struct OutJob { const void *s0, *s1; void *d0, *d1; size_t n0, n1; };
        OutJob* job = new (std::nothrow) OutJob{oc.p, (const char*)oc.p + L * sizeof(Affine), C_out, inf_out, L * sizeof(Affine), L};
        if (!job) return SBN_ERR_OOM;
        cudaError_t ce = cudaMemcpyAsync(oc.p, bC.p, L * sizeof(Affine), cudaMemcpyDeviceToHost, main);
        if (ce == cudaSuccess) ce = cudaMemcpyAsync((char*)oc.p + L * sizeof(Affine), bI.p, L, cudaMemcpyDeviceToHost, main);
        if (ce == cudaSuccess)
            ce = cudaLaunchHostFunc(main, [](void* u) {
                OutJob* j = (OutJob*)u;
                memcpy(j->d0, j->s0, j->n0);
                memcpy(j->d1, j->s1, j->n1);
                delete j;
            }, job);
        if (ce != cudaSuccess) { delete job; SBN_CUDA(ctx, ce); }
    } else {
        SBN_CUDA(ctx, cudaMemcpyAsync(C_out, bC.p, L * sizeof(Affine), cudaMemcpyDeviceToHost, main));
        SBN_CUDA(ctx, cudaMemcpyAsync(inf_out, bI.p, L, cudaMemcpyDeviceToHost, main));
    }
    ctx->d2h += L * sizeof(Affine) + L;
    if (async) {
        if (!ctx->hset_done[hset]) SBN_CUDA(ctx, cudaEventCreateWithFlags(&ctx->hset_done[hset], cudaEventDisableTiming));
        SBN_CUDA(ctx, cudaEventRecord(ctx->hset_done[hset], main));
    }
    if (!async || !ctx->last_was_mult) {      // only the tabulated-sum path keeps per-call workspaces: everything else completes here
        SBN_CUDA(ctx, cudaStreamSynchronize(main));
        if (!async) collect_profile(ctx, ev_stage);
    }
    return SBN_OK;
}

extern "C" int sbn_hyrax_commit(sbn_ctx* ctx, const sbn_bases* b, const sbn_fr* Z, size_t L, size_t R, const sbn_fr* blinds,
                                sbn_g1a* C_out, uint8_t* inf_out) {
    return hyrax_commit_host(ctx, b, Z, L, R, blinds, C_out, inf_out, nullptr);
}
extern "C" int sbn_hyrax_commit_async(sbn_ctx* ctx, const sbn_bases* b, const sbn_fr* Z, size_t L, size_t R, const sbn_fr* blinds,
                                      sbn_g1a* C_out, uint8_t* inf_out, void* stream) {
    if (!stream) return SBN_ERR_ARG;
    return hyrax_commit_host(ctx, b, Z, L, R, blinds, C_out, inf_out, (cudaStream_t)stream);
}

// One process, k GPUs: the rows of one commit divided into k contiguous blocks, block i committed by context i (one
// context per device, each with its own resident copy of the generator set) from a host thread of its own -- the calls
// run concurrently, the blocks land in the caller's C_out / inf_out side by side, and there is no inter-GPU traffic at
// all (hyrax.rs:259-265: rows are independent; the Rayon fan-out of the reference becomes a fan-out over devices).
extern "C" int sbn_hyrax_commit_multi(sbn_ctx* const* ctxs, const sbn_bases* const* bases, size_t k, const sbn_fr* Z, size_t L,
                                      size_t R, const sbn_fr* blinds, sbn_g1a* C_out, uint8_t* inf_out) {
    if (!ctxs || !bases || k == 0 || !Z || !C_out || !inf_out) return SBN_ERR_ARG;
    for (size_t i = 0; i < k; i++)
        if (!ctxs[i] || !bases[i] || bases[i]->ctx != ctxs[i]) return SBN_ERR_ARG;
    if (L == 0 || R == 0) return SBN_ERR_SHAPE;
    const size_t kk = std::min(k, L);
    std::vector<int> rc(kk, SBN_OK);
    std::vector<std::thread> workers;
    size_t row0 = 0;
    for (size_t i = 0; i < kk; i++) {
        const size_t n = L / kk + (i < L % kk ? 1 : 0);
        workers.emplace_back([=, &rc] {
            rc[i] = sbn_hyrax_commit(ctxs[i], bases[i], Z + row0 * R, n, R, blinds ? blinds + row0 : nullptr, C_out + row0,
                                     inf_out + row0);
        });
        row0 += n;
    }
    for (auto& w : workers) w.join();
    for (size_t i = 0; i < kk; i++)
        if (rc[i] != SBN_OK) return rc[i];
    return SBN_OK;
}

// ------------------------------------------------------------------------------------------------
// small helpers shared by the remaining entry points
// ------------------------------------------------------------------------------------------------
static int upload(sbn_ctx* ctx, DevBuf& buf, const void* host, size_t bytes) {
    SBN_TRY(ensure(ctx, buf, bytes));
    SBN_CUDA(ctx, cudaMemcpyAsync(buf.p, host, bytes, cudaMemcpyHostToDevice, ctx->compute));
    ctx->h2d += bytes;
    return SBN_OK;
}
static int download(sbn_ctx* ctx, void* host, const void* dev, size_t bytes) {
    SBN_CUDA(ctx, cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->compute));
    ctx->d2h += bytes;
    return SBN_OK;
}

extern "C" int sbn_g1_scalar_mul_batch(sbn_ctx* ctx, const sbn_g1a* P, const sbn_fr* s, size_t n, sbn_g1a* out,
                                       uint8_t* inf_out) {
    if (!ctx || !P || !s || !out || !inf_out) return SBN_ERR_ARG;
    if (n == 0) return SBN_OK;
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_ENTER(ctx);
    SBN_TRY(upload(ctx, ctx->scratch0, P, sizeof(Affine)));
    SBN_TRY(upload(ctx, ctx->scratch1, s, n * sizeof(Fr)));
    SBN_TRY(ensure(ctx, ctx->dC, n * sizeof(Affine)));
    SBN_TRY(ensure(ctx, ctx->dinf, n));
    k_scalar_mul<<<(unsigned)((n + 63) / 64), 64, 0, ctx->compute>>>((const Affine*)ctx->scratch0.p, nullptr, 0,
                                                                      (const Fr*)ctx->scratch1.p, 1, (int)n,
                                                                      (Affine*)ctx->dC.p, (uint8_t*)ctx->dinf.p);
    ctx->launches++;
    SBN_CUDA(ctx, cudaGetLastError());
    SBN_TRY(download(ctx, out, ctx->dC.p, n * sizeof(Affine)));
    SBN_TRY(download(ctx, inf_out, ctx->dinf.p, n));
    SBN_CUDA(ctx, cudaStreamSynchronize(ctx->compute));
    return SBN_OK;
}

extern "C" int sbn_g1_scale_points(sbn_ctx* ctx, const sbn_g1a* P, const uint8_t* inf, size_t n, const sbn_fr* s,
                                   sbn_g1a* out, uint8_t* inf_out) {
    if (!ctx || !P || !s || !out || !inf_out) return SBN_ERR_ARG;
    if (n == 0) return SBN_OK;
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_ENTER(ctx);
    SBN_TRY(upload(ctx, ctx->scratch0, P, n * sizeof(Affine)));
    SBN_TRY(upload(ctx, ctx->scratch1, s, sizeof(Fr)));
    const uint8_t* dinf_in = nullptr;
    if (inf) {
        SBN_TRY(upload(ctx, ctx->scratch2, inf, n));
        dinf_in = (const uint8_t*)ctx->scratch2.p;
    }
    SBN_TRY(ensure(ctx, ctx->dC, n * sizeof(Affine)));
    SBN_TRY(ensure(ctx, ctx->dinf, n));
    k_scalar_mul<<<(unsigned)((n + 63) / 64), 64, 0, ctx->compute>>>((const Affine*)ctx->scratch0.p, dinf_in, 1,
                                                                      (const Fr*)ctx->scratch1.p, 0, (int)n,
                                                                      (Affine*)ctx->dC.p, (uint8_t*)ctx->dinf.p);
    ctx->launches++;
    SBN_CUDA(ctx, cudaGetLastError());
    SBN_TRY(download(ctx, out, ctx->dC.p, n * sizeof(Affine)));
    SBN_TRY(download(ctx, inf_out, ctx->dinf.p, n));
    SBN_CUDA(ctx, cudaStreamSynchronize(ctx->compute));
    return SBN_OK;
}

static int fr_convert(sbn_ctx* ctx, const void* in, size_t n, int to_mont, void* out) {
    if (!ctx || !in || !out) return SBN_ERR_ARG;
    if (n == 0) return SBN_OK;
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_ENTER(ctx);
    SBN_TRY(upload(ctx, ctx->scratch0, in, n * sizeof(Fr)));
    SBN_TRY(ensure(ctx, ctx->scratch1, n * sizeof(Fr)));
    k_fr_convert<<<(unsigned)((n + 127) / 128), 128, 0, ctx->compute>>>((const Fr*)ctx->scratch0.p, (int)n, to_mont,
                                                                         (Fr*)ctx->scratch1.p);
    ctx->launches++;
    SBN_CUDA(ctx, cudaGetLastError());
    SBN_TRY(download(ctx, out, ctx->scratch1.p, n * sizeof(Fr)));
    SBN_CUDA(ctx, cudaStreamSynchronize(ctx->compute));
    return SBN_OK;
}
extern "C" int sbn_fr_from_canonical(sbn_ctx* ctx, const uint64_t* canon, size_t n, sbn_fr* out) {
    return fr_convert(ctx, canon, n, 1, out);
}
extern "C" int sbn_fr_to_canonical(sbn_ctx* ctx, const sbn_fr* in, size_t n, uint64_t* canon) {
    return fr_convert(ctx, in, n, 0, canon);
}

extern "C" int sbn_microbench(sbn_ctx* ctx, int kind, double* per_second) {
    if (!ctx || !per_second) return SBN_ERR_ARG;
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_ENTER(ctx);
    SBN_TRY(ensure(ctx, ctx->scratch0, 4096));
    cudaDeviceProp prop;
    SBN_CUDA(ctx, cudaGetDeviceProperties(&prop, ctx->device));
    const int blocks = prop.multiProcessorCount * 4, threads = 256;
    cudaEvent_t e0, e1;
    SBN_CUDA(ctx, cudaEventCreate(&e0));
    SBN_CUDA(ctx, cudaEventCreate(&e1));
    double best = 0;
    for (int rep = 0; rep < 4; rep++) {
        int iters = kind == 3 ? 2000 : 20000;
        SBN_CUDA(ctx, cudaEventRecord(e0, ctx->compute));
        switch (kind) {
            case 0: k_microbench<0><<<blocks, threads, 0, ctx->compute>>>((uint32_t*)ctx->scratch0.p, iters, 12345u + rep); break;
            case 1: k_microbench<1><<<blocks, threads, 0, ctx->compute>>>((uint32_t*)ctx->scratch0.p, iters, 12345u + rep); break;
            case 2: k_microbench<2><<<blocks, threads, 0, ctx->compute>>>((uint32_t*)ctx->scratch0.p, iters, 12345u + rep); break;
            case 3: k_microbench_fqmul<<<blocks, threads, 0, ctx->compute>>>((Fq*)ctx->scratch0.p, iters, 12345u + rep); break;
            default: cudaEventDestroy(e0); cudaEventDestroy(e1); return SBN_ERR_ARG;
        }
        ctx->launches++;
        SBN_CUDA(ctx, cudaEventRecord(e1, ctx->compute));
        SBN_CUDA(ctx, cudaEventSynchronize(e1));
        float ms = 0;
        SBN_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
        double ops = double(blocks) * threads * double(iters) * (kind == 3 ? 4.0 : 32.0) * (kind == 2 ? 2.0 : 1.0);
        if (rep > 0) best = std::max(best, ops / (ms * 1e-3));
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *per_second = best;
    return SBN_OK;
}

// ------------------------------------------------------------------------------------------------
// a6 / a5: single MSMs
// ------------------------------------------------------------------------------------------------
extern "C" int sbn_msm(sbn_ctx* ctx, const sbn_g1a* points, const uint8_t* inf, const sbn_fr* scalars, size_t n,
                       sbn_g1a* out, uint8_t* inf_out) {
    if (!ctx || !out || !inf_out || (n && (!points || !scalars))) return SBN_ERR_ARG;
    if (n > (1u << 26)) return SBN_ERR_SHAPE;
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_ENTER(ctx);
    if (n == 0) {   // empty sum = identity
        memset(out, 0, sizeof(*out));
        *inf_out = 1;
        return SBN_OK;
    }
    cudaStream_t st = ctx->compute;
    SBN_TRY(upload(ctx, ctx->scratch0, points, n * sizeof(Affine)));
    SBN_TRY(upload(ctx, ctx->scratch1, scalars, n * sizeof(Fr)));
    const uint8_t* dinf_in = nullptr;
    if (inf) {
        SBN_TRY(upload(ctx, ctx->scratch2, inf, n));
        dinf_in = (const uint8_t*)ctx->scratch2.p;
    }
    const unsigned blocks = (unsigned)((n + kSmallThreads - 1) / kSmallThreads);
    SBN_TRY(ensure(ctx, ctx->totals, (blocks + 1) * sizeof(XYZZ)));
    SBN_TRY(ensure(ctx, ctx->dC, sizeof(Affine)));
    SBN_TRY(ensure(ctx, ctx->dinf, 1));
    XYZZ* part = (XYZZ*)ctx->totals.p;
    k_msm_naive<<<dim3(blocks, 1), kSmallThreads, 0, st>>>((const Affine*)ctx->scratch0.p, dinf_in, 0,
                                                           (const Fr*)ctx->scratch1.p, 0, (int)n, 1, part);
    k_points_sum<<<1, kSmallThreads, 0, st>>>(part, (int)blocks, part + blocks, 1);
    k_combine<<<1, 1, 0, st>>>(part + blocks, 1, 1, (Affine*)ctx->dC.p, (uint8_t*)ctx->dinf.p);
    ctx->launches += 3;
    SBN_CUDA(ctx, cudaGetLastError());
    SBN_TRY(download(ctx, out, ctx->dC.p, sizeof(Affine)));
    SBN_TRY(download(ctx, inf_out, ctx->dinf.p, 1));
    SBN_CUDA(ctx, cudaStreamSynchronize(st));
    return SBN_OK;
}

extern "C" int sbn_commit(sbn_ctx* ctx, const sbn_bases* bases, const sbn_fr* scalars, size_t n, const sbn_fr* blind,
                          sbn_g1a* out, uint8_t* inf_out) {
    if (!bases) return SBN_ERR_ARG;
    return sbn_hyrax_commit(ctx, bases, scalars, 1, n, blind, out, inf_out);
}

// ------------------------------------------------------------------------------------------------
// a11: bound
// ------------------------------------------------------------------------------------------------
static int bound_device(sbn_ctx* ctx, const Fr* dZ, const Fr* dL, size_t L, size_t R, Fr* dout, cudaStream_t st) {
    const unsigned bx = (unsigned)((R + 127) / 128);
    unsigned slices = (unsigned)std::max<size_t>(1, std::min<size_t>(L, 1184 / std::max(1u, bx)));
    const int rows_per_slice = (int)((L + slices - 1) / slices);
    slices = (unsigned)((L + rows_per_slice - 1) / rows_per_slice);
    SBN_TRY(ensure(ctx, ctx->scratch2, (size_t)slices * R * sizeof(Fr)));
    k_bound_partial<<<dim3(bx, slices), 128, 0, st>>>(dZ, dL, (int)L, (int)R, rows_per_slice, (Fr*)ctx->scratch2.p);
    k_fr_colsum<<<bx, 128, 0, st>>>((const Fr*)ctx->scratch2.p, (int)slices, (int)R, dout);
    ctx->launches += 2;
    SBN_CUDA(ctx, cudaGetLastError());
    return SBN_OK;
}

extern "C" int sbn_bound(sbn_ctx* ctx, const sbn_fr* Z, const sbn_fr* Lv, size_t L, size_t R, sbn_fr* LZ_out) {
    if (!ctx || !Z || !Lv || !LZ_out) return SBN_ERR_ARG;
    if (L == 0 || R == 0 || L > (1u << 24) || R > (1u << 24)) return SBN_ERR_SHAPE;
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_ENTER(ctx);
    SBN_TRY(upload(ctx, ctx->dZ, Z, L * R * sizeof(Fr)));
    SBN_TRY(upload(ctx, ctx->scratch0, Lv, L * sizeof(Fr)));
    SBN_TRY(ensure(ctx, ctx->scratch1, R * sizeof(Fr)));
    SBN_TRY(bound_device(ctx, (const Fr*)ctx->dZ.p, (const Fr*)ctx->scratch0.p, L, R, (Fr*)ctx->scratch1.p, ctx->compute));
    SBN_TRY(download(ctx, LZ_out, ctx->scratch1.p, R * sizeof(Fr)));
    SBN_CUDA(ctx, cudaStreamSynchronize(ctx->compute));
    return SBN_OK;
}

// ------------------------------------------------------------------------------------------------
// resident polynomial
// ------------------------------------------------------------------------------------------------
struct sbn_poly {
    sbn_ctx* ctx = nullptr;
    Fr* Z = nullptr;
    size_t len = 0;
};

// H2D copy of a large host buffer, synchronous.  From pageable memory (a Rust Vec, a numpy array) cudaMemcpy stages through the
// driver at 4-12 GB/s; here 8 MiB pieces go through two pinned buffers filled by four host threads, so the memcpy of piece
// i + 1 runs under the DMA of piece i.  Pinned sources take the direct copy.
static cudaError_t h2d_sync(sbn_ctx* ctx, void* dst, const void* src, size_t bytes, cudaStream_t st) {
    static constexpr size_t kPiece = size_t(8) << 20;
    bool pageable = bytes >= 2 * kPiece;
    if (pageable) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, src) != cudaSuccess) cudaGetLastError();
        else pageable = at.type == cudaMemoryTypeUnregistered;
    }
    if (!pageable) {
        cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st);
        return e != cudaSuccess ? e : cudaStreamSynchronize(st);
    }
    for (int k = 0; k < 2; k++) {
        HostBuf& ring = ctx->stage_pin[0][k];
        if (ring.cap < kPiece) {
            if (ring.p) { cudaDeviceSynchronize(); cudaFreeHost(ring.p); ring.p = nullptr; ring.cap = 0; }
            cudaError_t e = cudaHostAlloc(&ring.p, kPiece, cudaHostAllocDefault);
            if (e != cudaSuccess) return e;
            ring.cap = kPiece;
        }
        if (!ring.busy) { cudaError_t e = cudaEventCreateWithFlags(&ring.busy, cudaEventDisableTiming); if (e != cudaSuccess) return e; }
    }
    size_t off = 0;
    for (int k = 0; off < bytes; k ^= 1) {
        HostBuf& ring = ctx->stage_pin[0][k];
        const size_t n = std::min(kPiece, bytes - off);
        cudaError_t e = cudaEventSynchronize(ring.busy);          // the last DMA out of this buffer (an unrecorded event is complete)
        if (e != cudaSuccess) return e;
        host_copy_mt(ring.p, (const char*)src + off, n);
        if ((e = cudaMemcpyAsync((char*)dst + off, ring.p, n, cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
        if ((e = cudaEventRecord(ring.busy, st)) != cudaSuccess) return e;
        off += n;
    }
    return cudaStreamSynchronize(st);
}

extern "C" int sbn_poly_upload(sbn_ctx* ctx, const sbn_fr* Z, size_t len, sbn_poly** out) {
    if (!ctx || !Z || !out) return SBN_ERR_ARG;
    *out = nullptr;
    if (len == 0 || len > (size_t(1) << 32)) return SBN_ERR_SHAPE;
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_ENTER(ctx);
    sbn_poly* p = new (std::nothrow) sbn_poly();
    if (!p) return SBN_ERR_OOM;
    p->ctx = ctx;
    p->len = len;
    if (pool_alloc(ctx, &p->Z, len * sizeof(Fr)) != cudaSuccess) { delete p; ctx->last_error = "sbn_poly_upload: cudaMalloc failed"; return SBN_ERR_OOM; }
    if (h2d_sync(ctx, p->Z, Z, len * sizeof(Fr), ctx->compute) != cudaSuccess) {
        cudaGetLastError();
        pool_release(ctx, p->Z);
        delete p;
        ctx->last_error = "sbn_poly_upload: copy failed";
        return SBN_ERR_CUDA;
    }
    ctx->h2d += len * sizeof(Fr);
    *out = p;
    return SBN_OK;
}

extern "C" int sbn_poly_destroy(sbn_poly* p) {
    if (!p) return SBN_ERR_ARG;
    {
        std::lock_guard<std::mutex> g(p->ctx->mu);
        cudaSetDevice(p->ctx->device);
        pool_free(p->ctx, p->Z, p->len * sizeof(Fr));      // stream-ordered reuse, see sbn_ctx::mem_pool
    }
    delete p;
    return SBN_OK;
}

// rows [first, first + L) of the resident polynomial read as rows of R scalars; blinds: one per committed row
static int poly_commit_rows(sbn_ctx* ctx, const sbn_bases* b, const sbn_poly* poly, size_t first, size_t L, size_t R, const sbn_fr* blinds,
                            sbn_g1a* C_out, uint8_t* inf_out) {
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_ENTER(ctx);
    const size_t chunk = commit_chunk_rows(ctx, L);
    SBN_TRY(ensure_commit_workspace(ctx, b, chunk, L));
    SBN_TRY(ensure(ctx, ctx->dC, L * sizeof(Affine)));
    SBN_TRY(ensure(ctx, ctx->dinf, L));
    Fr* dbl = nullptr;
    if (blinds) {
        SBN_TRY(upload(ctx, ctx->dblinds, blinds, L * sizeof(Fr)));
        dbl = (Fr*)ctx->dblinds.p;
    }
    std::vector<int> ev_stage;
    SBN_TRY(run_commit(ctx, b, poly->Z + first * R, nullptr, L, R, dbl, (Affine*)ctx->dC.p, (uint8_t*)ctx->dinf.p, ctx->compute, ev_stage));
    SBN_TRY(download(ctx, C_out, ctx->dC.p, L * sizeof(Affine)));
    SBN_TRY(download(ctx, inf_out, ctx->dinf.p, L));
    SBN_CUDA(ctx, cudaStreamSynchronize(ctx->compute));
    collect_profile(ctx, ev_stage);
    return SBN_OK;
}

extern "C" int sbn_poly_commit(sbn_ctx* ctx, const sbn_bases* b, const sbn_poly* poly, size_t L, size_t R, const sbn_fr* blinds,
                               sbn_g1a* C_out, uint8_t* inf_out) {
    if (!ctx || !b || !poly || !C_out || !inf_out || b->ctx != ctx || poly->ctx != ctx) return SBN_ERR_ARG;
    SBN_TRY(check_commit_shape(b, L, R));
    if (L * R != poly->len) return SBN_ERR_SHAPE;               // hyrax.rs:258
    return poly_commit_rows(ctx, b, poly, 0, L, R, blinds, C_out, inf_out);
}

// A block of rows of the same commitment: rows are independent (hyrax.rs:259-265), so k ranks (one process per GPU, each
// holding the polynomial) commit k blocks and exchange 65 bytes per row.
extern "C" int sbn_poly_commit_rows(sbn_ctx* ctx, const sbn_bases* b, const sbn_poly* poly, size_t first_row, size_t n_rows, size_t R,
                                    const sbn_fr* blinds, sbn_g1a* C_out, uint8_t* inf_out) {
    if (!ctx || !b || !poly || !C_out || !inf_out || b->ctx != ctx || poly->ctx != ctx) return SBN_ERR_ARG;
    SBN_TRY(check_commit_shape(b, n_rows, R));
    if (R == 0 || poly->len % R != 0 || first_row > poly->len / R || n_rows > poly->len / R - first_row) return SBN_ERR_SHAPE;
    return poly_commit_rows(ctx, b, poly, first_row, n_rows, R, blinds, C_out, inf_out);
}

extern "C" int sbn_poly_bound(sbn_ctx* ctx, const sbn_poly* poly, const sbn_fr* Lv, size_t L, size_t R, sbn_fr* LZ_out) {
    if (!ctx || !poly || !Lv || !LZ_out || poly->ctx != ctx) return SBN_ERR_ARG;
    if (L == 0 || R == 0 || L * R != poly->len) return SBN_ERR_SHAPE;
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_ENTER(ctx);
    SBN_TRY(upload(ctx, ctx->scratch0, Lv, L * sizeof(Fr)));
    SBN_TRY(ensure(ctx, ctx->scratch1, R * sizeof(Fr)));
    SBN_TRY(bound_device(ctx, poly->Z, (const Fr*)ctx->scratch0.p, L, R, (Fr*)ctx->scratch1.p, ctx->compute));
    SBN_TRY(download(ctx, LZ_out, ctx->scratch1.p, R * sizeof(Fr)));
    SBN_CUDA(ctx, cudaStreamSynchronize(ctx->compute));
    return SBN_OK;
}

// ------------------------------------------------------------------------------------------------
// a14: bullet reduction with device-resident G, a, b
// ------------------------------------------------------------------------------------------------
struct sbn_bullet {
    sbn_ctx* ctx = nullptr;
    const sbn_bases* bases = nullptr;
    size_t n0 = 0;              // original length
    size_t n = 0;               // current length of a, b (and of G on the folding path)
    bool fast = false;          // table-based rounds (Q = q * g1); otherwise G is folded explicitly
    Affine* G = nullptr;        // folding path only
    uint8_t* Ginf = nullptr;
    Fr *a = nullptr, *b = nullptr;
    Affine* QH = nullptr;       // folding path: [Q, H, Q, H]
    Fr* scal = nullptr;         // [c_L, c_R, blind_L, blind_R, u, u_inv, q, tmp]
    XYZZ* partial = nullptr;    // folding path: 2 x blocks
    XYZZ* terms = nullptr;      // folding path: 2 groups x 3
    Fr* frpart = nullptr;       // 2 x blocks
    Affine* outp = nullptr;     // 2
    uint8_t* outinf = nullptr;  // 2
    Fr* coef[2] = {nullptr, nullptr};   // fast path: ping-pong coefficient vectors, length n0 / n
    int coef_cur = 0;
    Fr* rows = nullptr;         // fast path: 2 x (n0 + 1) expanded scalars
    unsigned max_blocks = 0;
};

static void bullet_free(sbn_bullet* st) {
    for (void* p : {(void*)st->G, (void*)st->Ginf, (void*)st->a, (void*)st->b, (void*)st->QH, (void*)st->scal,
                    (void*)st->partial, (void*)st->terms, (void*)st->frpart, (void*)st->outp, (void*)st->outinf,
                    (void*)st->coef[0], (void*)st->coef[1], (void*)st->rows})
        pool_release(st->ctx, p);
    delete st;
}

static const unsigned kBulletDotBlocks = 64;

extern "C" int sbn_bullet_begin(sbn_ctx* ctx, const sbn_bases* bases, const sbn_g1a* Q, const sbn_fr* q_scalar, const sbn_fr* a,
                                const sbn_fr* b, size_t n, const sbn_fr* blind, sbn_g1a* Gamma_out, uint8_t* Gamma_inf,
                                sbn_bullet** out) {
    if (!ctx || !bases || !a || !b || !blind || !Gamma_out || !Gamma_inf || !out || bases->ctx != ctx) return SBN_ERR_ARG;
    if ((Q == nullptr) == (q_scalar == nullptr)) return SBN_ERR_ARG;        // exactly one description of Q
    *out = nullptr;
    if (n == 0 || (n & (n - 1)) || n != bases->n) return SBN_ERR_SHAPE;     // bullet.rs:42-47
    const bool fast = q_scalar != nullptr;
    if (fast && !bases->has_g1) return SBN_ERR_ARG;
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_ENTER(ctx);
    cudaStream_t s = ctx->compute;
    sbn_bullet* st = new (std::nothrow) sbn_bullet();
    if (!st) return SBN_ERR_OOM;
    st->ctx = ctx;
    st->bases = bases;
    st->n0 = st->n = n;
    st->fast = fast;
    st->max_blocks = (unsigned)((n / 2 + kSmallThreads - 1) / kSmallThreads) + 1;
    bool ok = pool_alloc(ctx, &st->a, n * sizeof(Fr)) == cudaSuccess && pool_alloc(ctx, &st->b, n * sizeof(Fr)) == cudaSuccess &&
              pool_alloc(ctx, &st->scal, 8 * sizeof(Fr)) == cudaSuccess &&
              pool_alloc(ctx, &st->frpart, 2 * (kBulletDotBlocks + 1) * sizeof(Fr)) == cudaSuccess &&
              pool_alloc(ctx, &st->outp, 2 * sizeof(Affine)) == cudaSuccess && pool_alloc(ctx, &st->outinf, 2) == cudaSuccess;
    if (ok && fast)
        ok = pool_alloc(ctx, &st->coef[0], n * sizeof(Fr)) == cudaSuccess && pool_alloc(ctx, &st->coef[1], n * sizeof(Fr)) == cudaSuccess &&
             pool_alloc(ctx, &st->rows, 2 * (n + 1) * sizeof(Fr)) == cudaSuccess;
    if (ok && !fast)
        ok = pool_alloc(ctx, &st->G, n * sizeof(Affine)) == cudaSuccess && pool_alloc(ctx, &st->Ginf, n) == cudaSuccess &&
             pool_alloc(ctx, &st->QH, 4 * sizeof(Affine)) == cudaSuccess &&
             pool_alloc(ctx, &st->partial, 2 * (st->max_blocks + 1) * sizeof(XYZZ)) == cudaSuccess &&
             pool_alloc(ctx, &st->terms, 6 * sizeof(XYZZ)) == cudaSuccess;
    if (!ok) { bullet_free(st); ctx->last_error = "sbn_bullet_begin: cudaMalloc failed"; return SBN_ERR_OOM; }
    // a failed commit may leave kernels of the pipeline streams in flight: drain them before the buffers return to the pool
    auto fail = [&](int code) { cudaDeviceSynchronize(); cudaGetLastError(); bullet_free(st); return code; };
#define BCUDA(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { ctx->last_error = std::string(#call) + ": " + cudaGetErrorString(_e); return fail(SBN_ERR_CUDA); } } while (0)
    BCUDA(cudaMemcpyAsync(st->a, a, n * sizeof(Fr), cudaMemcpyHostToDevice, s));
    BCUDA(cudaMemcpyAsync(st->b, b, n * sizeof(Fr), cudaMemcpyHostToDevice, s));
    BCUDA(cudaMemcpyAsync(st->scal + 2, blind, sizeof(Fr), cudaMemcpyHostToDevice, s));
    ctx->h2d += 2 * n * sizeof(Fr) + sizeof(Fr);
    // <a, b>
    k_fr_dot<<<dim3(kBulletDotBlocks, 1), kDotThreads, 0, s>>>(st->a, 0, st->b, 0, (int)n, st->frpart);
    k_fr_sum<<<1, kDotThreads, 0, s>>>(st->frpart, (int)kBulletDotBlocks, st->scal, 1);
    ctx->launches += 2;
    std::vector<int> ev_stage;
    int rc;
    if (fast) {
        // Gamma = one row [a, <a,b> * q] over (G, g1) with blind on h               (bullet.rs:57-59)
        const sbn_fr one_mont = {{0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL}};
        BCUDA(cudaMemcpyAsync(st->scal + 6, q_scalar, sizeof(Fr), cudaMemcpyHostToDevice, s));
        BCUDA(cudaMemcpyAsync(st->coef[0], &one_mont, sizeof(Fr), cudaMemcpyHostToDevice, s));
        k_bullet_row_single<<<(unsigned)((n + 128) / 128), 128, 0, s>>>(st->a, (int)n, st->scal, st->scal + 6, st->rows);
        ctx->launches++;
        rc = ensure_commit_workspace(ctx, bases, 2, 2);
        if (rc != SBN_OK) return fail(rc);
        rc = run_commit(ctx, bases, st->rows, nullptr, 1, n + 1, st->scal + 2, st->outp, st->outinf, s, ev_stage);
        if (rc != SBN_OK) return fail(rc);
    } else {
        // G <- the resident generators (window 0 of the tables), H = h
        BCUDA(cudaMemcpyAsync(st->G, bases->orig, n * sizeof(Affine), cudaMemcpyDeviceToDevice, s));
        BCUDA(cudaMemsetAsync(st->Ginf, 0, n, s));
        for (int k = 0; k < 2; k++) {
            BCUDA(cudaMemcpyAsync(st->QH + 2 * k, Q, sizeof(Affine), cudaMemcpyHostToDevice, s));
            BCUDA(cudaMemcpyAsync(st->QH + 2 * k + 1, bases->orig + (bases->n_cols - 1), sizeof(Affine), cudaMemcpyDeviceToDevice, s));
        }
        ctx->h2d += sizeof(Affine);
        // Gamma = MSM(a, G) + blind * H (one row through the table pipeline) + <a, b> * Q
        rc = ensure_commit_workspace(ctx, bases, 1, 1);
        if (rc != SBN_OK) return fail(rc);
        rc = run_commit(ctx, bases, st->a, nullptr, 1, n, st->scal + 2, nullptr, nullptr, s, ev_stage, false);
        if (rc != SBN_OK) return fail(rc);
        k_scalar_mul_terms<<<1, 32, 0, s>>>(st->QH, st->scal, 1, st->terms + 1, 1, 0);
        BCUDA(cudaMemcpyAsync(st->terms, ctx->totals.p, sizeof(XYZZ), cudaMemcpyDeviceToDevice, s));
        k_combine<<<1, 1, 0, s>>>(st->terms, 1, 2, st->outp, st->outinf);
        ctx->launches += 2;
    }
    BCUDA(cudaGetLastError());
    BCUDA(cudaMemcpyAsync(Gamma_out, st->outp, sizeof(Affine), cudaMemcpyDeviceToHost, s));
    BCUDA(cudaMemcpyAsync(Gamma_inf, st->outinf, 1, cudaMemcpyDeviceToHost, s));
    BCUDA(cudaStreamSynchronize(s));
    ctx->d2h += sizeof(Affine) + 1;
    *out = st;
    return SBN_OK;
}

extern "C" int sbn_bullet_round(sbn_bullet* st, const sbn_fr* blind_L, const sbn_fr* blind_R, sbn_g1a* L_out, uint8_t* L_inf,
                                sbn_g1a* R_out, uint8_t* R_inf) {
    if (!st || !blind_L || !blind_R || !L_out || !L_inf || !R_out || !R_inf) return SBN_ERR_ARG;
    if (st->n < 2) return SBN_ERR_SHAPE;
    sbn_ctx* ctx = st->ctx;
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_ENTER(ctx);
    cudaStream_t s = ctx->compute;
    const long n2 = (long)(st->n / 2);
    bool host_norm = false;
    const unsigned dot_blocks = (unsigned)std::min<long>(kBulletDotBlocks, (n2 + kDotThreads - 1) / kDotThreads);
    SBN_CUDA(ctx, cudaMemcpyAsync(st->scal + 2, blind_L, sizeof(Fr), cudaMemcpyHostToDevice, s));
    SBN_CUDA(ctx, cudaMemcpyAsync(st->scal + 3, blind_R, sizeof(Fr), cudaMemcpyHostToDevice, s));
    // set 0: c_L = <a_L, b_R>, set 1: c_R = <a_R, b_L>                                     (bullet.rs:70-71)
    k_fr_dot<<<dim3(dot_blocks, 2), kDotThreads, 0, s>>>(st->a, n2, st->b + n2, -n2, (int)n2, st->frpart);
    k_fr_sum<<<2, kDotThreads, 0, s>>>(st->frpart, (int)dot_blocks, st->scal, 1);
    ctx->launches += 2;
    if (st->fast) {
        // rows 0 / 1 = L / R over the original generators, see opening_kernels.cuh        (bullet.rs:75-76)
        const size_t n0 = st->n0;
        k_bullet_expand<<<(unsigned)((n0 + 128) / 128), 128, 0, s>>>(st->a, st->coef[st->coef_cur], (int)n0, (int)st->n, st->scal,
                                                                     st->scal + 6, st->rows);
        ctx->launches++;
        std::vector<int> ev_stage;
        SBN_TRY(ensure_commit_workspace(ctx, st->bases, 2, 2));
        // the two-row sum over the opening's table leaves its XYZZ totals in ctx->totals; they are normalised on the host
        const sbn_bases* bb = st->bases;
        host_norm = ctx->host_normalize && bb->small && ctx->small_commit_path && n0 + 2 <= (size_t)bb->n_cols &&
                    !(2 >= (size_t)ctx->mult_min_rows && !bb->has_g1 && ctx->mult_max_mb > 0);
        SBN_TRY(run_commit(ctx, st->bases, st->rows, nullptr, 2, n0 + 1, st->scal + 2, st->outp, st->outinf, s, ev_stage, !host_norm));
    } else {
        const unsigned blocks = (unsigned)((n2 + kSmallThreads - 1) / kSmallThreads);
        // set 0: L = <a_L, G_R>, set 1: R = <a_R, G_L>
        k_msm_naive<<<dim3(blocks, 2), kSmallThreads, 0, s>>>(st->G + n2, st->Ginf + n2, -n2, st->a, n2, (int)n2, 1, st->partial);
        k_points_sum<<<2, kSmallThreads, 0, s>>>(st->partial, (int)blocks, st->terms, 3);
        // terms: [L_msm, c_L Q, blind_L H, R_msm, c_R Q, blind_R H];  scalars [c_L, blind_L, c_R, blind_R]
        SBN_CUDA(ctx, cudaMemcpyAsync(st->scal + 7, st->scal + 1, sizeof(Fr), cudaMemcpyDeviceToDevice, s));   // c_R
        SBN_CUDA(ctx, cudaMemcpyAsync(st->scal + 1, st->scal + 2, sizeof(Fr), cudaMemcpyDeviceToDevice, s));   // blind_L
        SBN_CUDA(ctx, cudaMemcpyAsync(st->scal + 2, st->scal + 7, sizeof(Fr), cudaMemcpyDeviceToDevice, s));   // c_R
        k_scalar_mul_terms<<<4, 32, 0, s>>>(st->QH, st->scal, 4, st->terms + 1, 3, 2);
        k_combine<<<1, 2, 0, s>>>(st->terms, 2, 3, st->outp, st->outinf);
        ctx->launches += 4;
    }
    SBN_CUDA(ctx, cudaGetLastError());
    sbn_g1a pts[2];
    uint8_t infs[2];
    if (host_norm) {
        XYZZ tot[2];
        SBN_CUDA(ctx, cudaMemcpyAsync(tot, ctx->totals.p, 2 * sizeof(XYZZ), cudaMemcpyDeviceToHost, s));
        SBN_CUDA(ctx, cudaStreamSynchronize(s));
        for (int i = 0; i < 2; i++) host_normalize(tot[i], &pts[i], &infs[i]);
        ctx->d2h += 2 * sizeof(XYZZ);
    } else {
        SBN_CUDA(ctx, cudaMemcpyAsync(pts, st->outp, 2 * sizeof(Affine), cudaMemcpyDeviceToHost, s));
        SBN_CUDA(ctx, cudaMemcpyAsync(infs, st->outinf, 2, cudaMemcpyDeviceToHost, s));
        SBN_CUDA(ctx, cudaStreamSynchronize(s));
        ctx->d2h += 2 * sizeof(Affine) + 2;
    }
    ctx->h2d += 2 * sizeof(Fr);
    *L_out = pts[0]; *L_inf = infs[0];
    *R_out = pts[1]; *R_inf = infs[1];
    return SBN_OK;
}

extern "C" int sbn_bullet_fold(sbn_bullet* st, const sbn_fr* u, const sbn_fr* u_inv) {
    if (!st || !u || !u_inv) return SBN_ERR_ARG;
    if (st->n < 2) return SBN_ERR_SHAPE;
    sbn_ctx* ctx = st->ctx;
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_ENTER(ctx);
    cudaStream_t s = ctx->compute;
    const int n2 = (int)(st->n / 2);
    SBN_CUDA(ctx, cudaMemcpyAsync(st->scal + 4, u, sizeof(Fr), cudaMemcpyHostToDevice, s));
    SBN_CUDA(ctx, cudaMemcpyAsync(st->scal + 5, u_inv, sizeof(Fr), cudaMemcpyHostToDevice, s));
    if (st->fast) {
        const int len = (int)(st->n0 / st->n);
        k_coef_update<<<(len + 127) / 128, 128, 0, s>>>(st->coef[st->coef_cur], len, st->scal + 4, st->coef[st->coef_cur ^ 1]);
        st->coef_cur ^= 1;
    } else {
        k_fold_points<<<(n2 + kSmallThreads - 1) / kSmallThreads, kSmallThreads, 0, s>>>(st->G, st->Ginf, n2, st->scal + 4);
    }
    k_fold_scalars<<<(n2 + 127) / 128, 128, 0, s>>>(st->a, st->b, n2, st->scal + 4);
    ctx->launches += 2;
    ctx->h2d += 2 * sizeof(Fr);
    SBN_CUDA(ctx, cudaGetLastError());
    if (!st->fast) SBN_CUDA(ctx, cudaStreamSynchronize(s));
    st->n = n2;
    return SBN_OK;
}

extern "C" int sbn_bullet_end(sbn_bullet* st, sbn_fr* a_hat, sbn_fr* b_hat, sbn_g1a* g_hat, uint8_t* g_hat_inf) {
    if (!st || !a_hat || !b_hat || !g_hat || !g_hat_inf) return SBN_ERR_ARG;
    if (st->n != 1) return SBN_ERR_SHAPE;                      // bullet.rs:110-112 asserts
    sbn_ctx* ctx = st->ctx;
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_ENTER(ctx);
    cudaStream_t s = ctx->compute;
    const Affine* gsrc = st->G;
    const uint8_t* gisrc = st->Ginf;
    if (st->fast) {   // g_hat = MSM(coef, G): one more row over the tables
        const size_t n0 = st->n0;
        k_bullet_row_single<<<(unsigned)((n0 + 128) / 128), 128, 0, s>>>(st->coef[st->coef_cur], (int)n0, nullptr, nullptr, st->rows);
        ctx->launches++;
        std::vector<int> ev_stage;
        SBN_TRY(ensure_commit_workspace(ctx, st->bases, 2, 2));
        SBN_TRY(run_commit(ctx, st->bases, st->rows, nullptr, 1, n0 + 1, nullptr, st->outp, st->outinf, s, ev_stage));
        gsrc = st->outp;
        gisrc = st->outinf;
    }
    SBN_CUDA(ctx, cudaMemcpyAsync(a_hat, st->a, sizeof(Fr), cudaMemcpyDeviceToHost, s));
    SBN_CUDA(ctx, cudaMemcpyAsync(b_hat, st->b, sizeof(Fr), cudaMemcpyDeviceToHost, s));
    SBN_CUDA(ctx, cudaMemcpyAsync(g_hat, gsrc, sizeof(Affine), cudaMemcpyDeviceToHost, s));
    SBN_CUDA(ctx, cudaMemcpyAsync(g_hat_inf, gisrc, 1, cudaMemcpyDeviceToHost, s));
    SBN_CUDA(ctx, cudaStreamSynchronize(s));
    ctx->d2h += 2 * sizeof(Fr) + sizeof(Affine) + 1;
    if (*g_hat_inf) memset(g_hat, 0, sizeof(*g_hat));
    return SBN_OK;
}

// sbn_bullet_end plus DotProductProofLog's delta = d * g_hat + r_delta * h (nizk/mod.rs:497-500).  On the table path
// g_hat = <coef, G>, so delta = <d * coef, G> + r_delta * h is a second row of the same commit over the resident tables
// instead of a 254-step double-and-add chain on the one-off point g_hat (~3 ms).
extern "C" int sbn_bullet_end_delta(sbn_bullet* st, const sbn_fr* d, const sbn_fr* r_delta, sbn_fr* a_hat, sbn_fr* b_hat,
                                    sbn_g1a* g_hat, uint8_t* g_hat_inf, sbn_g1a* delta, uint8_t* delta_inf) {
    if (!st || !d || !r_delta || !a_hat || !b_hat || !g_hat || !g_hat_inf || !delta || !delta_inf) return SBN_ERR_ARG;
    if (st->n != 1) return SBN_ERR_SHAPE;                      // bullet.rs:110-112 asserts
    if (!st->fast) return SBN_ERR_UNSUPPORTED;                 // the explicit-folding path holds no coefficients
    sbn_ctx* ctx = st->ctx;
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_ENTER(ctx);
    cudaStream_t s = ctx->compute;
    const size_t n0 = st->n0;
    Fr blinds[2];
    blinds[0] = Fr::zero();
    memcpy(&blinds[1], r_delta, sizeof(Fr));
    SBN_CUDA(ctx, cudaMemcpyAsync(st->scal + 2, blinds, 2 * sizeof(Fr), cudaMemcpyHostToDevice, s));   // the rounds' blind slots
    SBN_CUDA(ctx, cudaMemcpyAsync(st->scal + 7, d, sizeof(Fr), cudaMemcpyHostToDevice, s));
    const unsigned blocks = (unsigned)((n0 + 128) / 128);
    k_bullet_row_single<<<blocks, 128, 0, s>>>(st->coef[st->coef_cur], (int)n0, nullptr, nullptr, st->rows);
    k_bullet_row_scaled<<<blocks, 128, 0, s>>>(st->coef[st->coef_cur], (int)n0, st->scal + 7, st->rows + (n0 + 1));
    ctx->launches += 2;
    std::vector<int> ev_stage;
    SBN_TRY(ensure_commit_workspace(ctx, st->bases, 2, 2));
    SBN_TRY(run_commit(ctx, st->bases, st->rows, nullptr, 2, n0 + 1, st->scal + 2, st->outp, st->outinf, s, ev_stage));
    Affine pts[2];
    uint8_t infs[2];
    SBN_CUDA(ctx, cudaMemcpyAsync(a_hat, st->a, sizeof(Fr), cudaMemcpyDeviceToHost, s));
    SBN_CUDA(ctx, cudaMemcpyAsync(b_hat, st->b, sizeof(Fr), cudaMemcpyDeviceToHost, s));
    SBN_CUDA(ctx, cudaMemcpyAsync(pts, st->outp, 2 * sizeof(Affine), cudaMemcpyDeviceToHost, s));
    SBN_CUDA(ctx, cudaMemcpyAsync(infs, st->outinf, 2, cudaMemcpyDeviceToHost, s));
    SBN_CUDA(ctx, cudaStreamSynchronize(s));
    ctx->d2h += 2 * sizeof(Fr) + 2 * (sizeof(Affine) + 1);
    memcpy(g_hat, &pts[0], sizeof(Affine));
    memcpy(delta, &pts[1], sizeof(Affine));
    *g_hat_inf = infs[0];
    *delta_inf = infs[1];
    if (*g_hat_inf) memset(g_hat, 0, sizeof(*g_hat));
    if (*delta_inf) memset(delta, 0, sizeof(*delta));
    return SBN_OK;
}

extern "C" int sbn_bullet_destroy(sbn_bullet* st) {
    if (!st) return SBN_ERR_ARG;
    {
        std::lock_guard<std::mutex> g(st->ctx->mu);
        cudaSetDevice(st->ctx->device);
        cudaStreamSynchronize(st->ctx->compute);
    }
    bullet_free(st);
    return SBN_OK;
}

// ------------------------------------------------------------------------------------------------
// a16: sumcheck rounds with the four tables resident
// ------------------------------------------------------------------------------------------------
struct sbn_sumcheck {
    sbn_ctx* ctx = nullptr;
    size_t len = 0;
    Fr* T[4] = {nullptr, nullptr, nullptr, nullptr};
    int ntables = 4;         // 4: cubic with additive term (tau, Az, Bz, Cz); 2: quadratic (z, ABC)
    Fr* partial = nullptr;   // 3 x blocks
    Fr* out = nullptr;       // 3 evals + r
    unsigned blocks = 0;
};

static void sumcheck_free(sbn_sumcheck* st) {
    for (Fr* t : st->T) pool_release(st->ctx, t);
    pool_release(st->ctx, st->partial);
    pool_release(st->ctx, st->out);
    delete st;
}

static int sumcheck_begin(sbn_ctx* ctx, const sbn_fr* const* src, int ntables, size_t len, sbn_sumcheck** out);

extern "C" int sbn_sumcheck_begin(sbn_ctx* ctx, const sbn_fr* tau, const sbn_fr* Az, const sbn_fr* Bz, const sbn_fr* Cz,
                                  size_t len, sbn_sumcheck** out) {
    if (!ctx || !tau || !Az || !Bz || !Cz || !out) return SBN_ERR_ARG;
    const sbn_fr* src[4] = {tau, Az, Bz, Cz};
    return sumcheck_begin(ctx, src, 4, len, out);
}
extern "C" int sbn_sumcheck_begin_quad(sbn_ctx* ctx, const sbn_fr* z, const sbn_fr* ABC, size_t len, sbn_sumcheck** out) {
    if (!ctx || !z || !ABC || !out) return SBN_ERR_ARG;
    const sbn_fr* src[4] = {z, ABC, nullptr, nullptr};
    return sumcheck_begin(ctx, src, 2, len, out);
}

static int sumcheck_begin(sbn_ctx* ctx, const sbn_fr* const* src, int ntables, size_t len, sbn_sumcheck** out) {
    *out = nullptr;
    if (len < 1 || (len & (len - 1)) || len > (1u << 28)) return SBN_ERR_SHAPE;
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_ENTER(ctx);
    sbn_sumcheck* st = new (std::nothrow) sbn_sumcheck();
    if (!st) return SBN_ERR_OOM;
    st->ctx = ctx;
    st->len = len;
    st->blocks = 592;
    st->ntables = ntables;
    bool ok = pool_alloc(ctx, &st->partial, 3 * st->blocks * sizeof(Fr)) == cudaSuccess && pool_alloc(ctx, &st->out, 4 * sizeof(Fr)) == cudaSuccess;
    for (int k = 0; k < ntables && ok; k++) ok = pool_alloc(ctx, &st->T[k], len * sizeof(Fr)) == cudaSuccess;
    if (!ok) { sumcheck_free(st); ctx->last_error = "sbn_sumcheck_begin: cudaMalloc failed"; return SBN_ERR_OOM; }
    for (int k = 0; k < ntables; k++) {
        if (cudaMemcpyAsync(st->T[k], src[k], len * sizeof(Fr), cudaMemcpyHostToDevice, ctx->compute) != cudaSuccess) {
            sumcheck_free(st);
            ctx->last_error = "sbn_sumcheck_begin: upload failed";
            return SBN_ERR_CUDA;
        }
    }
    ctx->h2d += ntables * len * sizeof(Fr);
    SBN_CUDA(ctx, cudaStreamSynchronize(ctx->compute));
    *out = st;
    return SBN_OK;
}

extern "C" int sbn_sumcheck_round_eval(sbn_sumcheck* st, sbn_fr* e0, sbn_fr* e2, sbn_fr* e3) {
    if (!st || !e0 || !e2 || (st->ntables == 4 && !e3)) return SBN_ERR_ARG;
    if (st->len < 2) return SBN_ERR_SHAPE;
    sbn_ctx* ctx = st->ctx;
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_ENTER(ctx);
    cudaStream_t s = ctx->compute;
    const int half = (int)(st->len / 2);
    const unsigned blocks = (unsigned)std::min<size_t>(st->blocks, (half + kDotThreads - 1) / kDotThreads);
    const int nevals = st->ntables == 4 ? 3 : 2;
    if (st->ntables == 4) k_sumcheck_eval<<<blocks, kDotThreads, 0, s>>>(st->T[0], st->T[1], st->T[2], st->T[3], half, st->partial);
    else k_sumcheck_eval_quad<<<blocks, kDotThreads, 0, s>>>(st->T[0], st->T[1], half, st->partial);
    k_fr_sum<<<nevals, kDotThreads, 0, s>>>(st->partial, (int)blocks, st->out, 1);
    ctx->launches += 2;
    SBN_CUDA(ctx, cudaGetLastError());
    sbn_fr host[3];
    SBN_CUDA(ctx, cudaMemcpyAsync(host, st->out, nevals * sizeof(Fr), cudaMemcpyDeviceToHost, s));
    SBN_CUDA(ctx, cudaStreamSynchronize(s));
    ctx->d2h += nevals * sizeof(Fr);
    *e0 = host[0]; *e2 = host[1];
    if (nevals == 3) *e3 = host[2];
    return SBN_OK;
}

extern "C" int sbn_sumcheck_bind(sbn_sumcheck* st, const sbn_fr* r) {
    if (!st || !r) return SBN_ERR_ARG;
    if (st->len < 2) return SBN_ERR_SHAPE;
    sbn_ctx* ctx = st->ctx;
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_ENTER(ctx);
    cudaStream_t s = ctx->compute;
    const int half = (int)(st->len / 2);
    SBN_CUDA(ctx, cudaMemcpyAsync(st->out + 3, r, sizeof(Fr), cudaMemcpyHostToDevice, s));
    k_bind_top<<<(half + 127) / 128, 128, 0, s>>>(st->T[0], st->T[1], st->T[2], st->T[3], half, st->out + 3);
    ctx->launches += 1;
    ctx->h2d += sizeof(Fr);
    SBN_CUDA(ctx, cudaGetLastError());
    SBN_CUDA(ctx, cudaStreamSynchronize(s));
    st->len = half;
    return SBN_OK;
}

extern "C" int sbn_sumcheck_end(sbn_sumcheck* st, sbn_fr finals[4]) {
    if (!st || !finals) return SBN_ERR_ARG;
    if (st->len != 1) return SBN_ERR_SHAPE;
    sbn_ctx* ctx = st->ctx;
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_ENTER(ctx);
    for (int k = 0; k < 4; k++) {
        if (k < st->ntables) SBN_CUDA(ctx, cudaMemcpyAsync(&finals[k], st->T[k], sizeof(Fr), cudaMemcpyDeviceToHost, ctx->compute));
        else memset(&finals[k], 0, sizeof(sbn_fr));
    }
    SBN_CUDA(ctx, cudaStreamSynchronize(ctx->compute));
    ctx->d2h += st->ntables * sizeof(Fr);
    return SBN_OK;
}

extern "C" int sbn_sumcheck_destroy(sbn_sumcheck* st) {
    if (!st) return SBN_ERR_ARG;
    {
        std::lock_guard<std::mutex> g(st->ctx->mu);
        cudaSetDevice(st->ctx->device);
        cudaStreamSynchronize(st->ctx->compute);
    }
    sumcheck_free(st);
    return SBN_OK;
}

// ------------------------------------------------------------------------------------------------
// f1: product circuits and the batched cubic sumcheck of the product layer
// ------------------------------------------------------------------------------------------------
struct sbn_prodcircuit {
    sbn_ctx* ctx = nullptr;
    size_t len = 0;
    int num_layers = 0;
    Fr* buf = nullptr;               // layer l = (left | right) of len >> l scalars at buf + off[l]
    std::vector<size_t> off;
};

extern "C" int sbn_prodcircuit_create(sbn_ctx* ctx, const sbn_fr* poly, size_t len, sbn_prodcircuit** out) {
    if (!ctx || !poly || !out) return SBN_ERR_ARG;
    *out = nullptr;
    if (len < 2 || (len & (len - 1)) || len > (size_t(1) << 30)) return SBN_ERR_SHAPE;
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_ENTER(ctx);
    sbn_prodcircuit* pc = new (std::nothrow) sbn_prodcircuit();
    if (!pc) return SBN_ERR_OOM;
    pc->ctx = ctx;
    pc->len = len;
    size_t total = 0;
    for (size_t n = len; n >= 2; n >>= 1) { pc->off.push_back(total); total += n; pc->num_layers++; }
    if (pool_alloc(ctx, &pc->buf, total * sizeof(Fr)) != cudaSuccess) { delete pc; ctx->last_error = "sbn_prodcircuit_create: cudaMalloc failed"; return SBN_ERR_OOM; }
    cudaStream_t s = ctx->compute;
    cudaError_t e = cudaMemcpyAsync(pc->buf, poly, len * sizeof(Fr), cudaMemcpyHostToDevice, s);
    ctx->h2d += len * sizeof(Fr);
    for (int l = 0; l + 1 < pc->num_layers && e == cudaSuccess; l++) {
        const size_t half = (len >> l) / 2;
        k_product_layer<<<(unsigned)((half + 127) / 128), 128, 0, s>>>(pc->buf + pc->off[l], half, pc->buf + pc->off[l + 1]);
        ctx->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) {
        ctx->last_error = std::string("sbn_prodcircuit_create: ") + cudaGetErrorString(e);
        pool_free(ctx, pc->buf, total * sizeof(Fr));
        delete pc;
        return SBN_ERR_CUDA;
    }
    *out = pc;
    return SBN_OK;
}

extern "C" int sbn_prodcircuit_evaluate(sbn_prodcircuit* pc, sbn_fr* out) {
    if (!pc || !out) return SBN_ERR_ARG;
    sbn_ctx* ctx = pc->ctx;
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_ENTER(ctx);
    Fr top[2];
    SBN_CUDA(ctx, cudaMemcpyAsync(top, pc->buf + pc->off[pc->num_layers - 1], 2 * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->compute));
    SBN_CUDA(ctx, cudaStreamSynchronize(ctx->compute));
    ctx->d2h += 2 * sizeof(Fr);
    const Fr v = fp_mul(top[0], top[1]);                 // product_tree.rs:59-64
    memcpy(out, &v, sizeof(Fr));
    return SBN_OK;
}

extern "C" size_t sbn_prodcircuit_num_layers(const sbn_prodcircuit* pc) { return pc ? (size_t)pc->num_layers : 0; }

extern "C" int sbn_prodcircuit_destroy(sbn_prodcircuit* pc) {
    if (!pc) return SBN_ERR_ARG;
    {
        std::lock_guard<std::mutex> g(pc->ctx->mu);
        cudaSetDevice(pc->ctx->device);
        pool_free(pc->ctx, pc->buf, (2 * pc->len - 2) * sizeof(Fr));
    }
    delete pc;
    return SBN_OK;
}

struct sbn_bsumcheck {
    sbn_ctx* ctx = nullptr;
    size_t P = 0, S = 0, len = 0, T0 = 0;
    std::vector<Fr*> A, B, C;        // per instance (C of a parallel instance = the shared eq table)
    Fr* eq[2] = {nullptr, nullptr};  // ping-pong buffers of EqPolynomial::evals; eq[cur] is poly_C_par
    int eq_cur = 0;
    Fr* seq = nullptr;               // 3 * S tables of the sequential instances
    CubicTriple* d_triples = nullptr;
    Fr** d_tables = nullptr;
    int ntables = 0;
    Fr* partial = nullptr;
    Fr* out = nullptr;               // 3 * (P + S) evaluations, then r
    unsigned max_blocks = 0;
};

static void bsumcheck_free(sbn_bsumcheck* st) {
    pool_free(st->ctx, st->eq[0], st->T0 * sizeof(Fr));
    pool_free(st->ctx, st->eq[1], st->T0 * sizeof(Fr));
    pool_free(st->ctx, st->seq, 3 * st->S * st->T0 * sizeof(Fr));
    for (void* p : {(void*)st->d_triples, (void*)st->d_tables, (void*)st->partial, (void*)st->out})
        pool_release(st->ctx, p);
    delete st;
}

// seq_dev != nullptr: the 3 * S tables of the sequential instances are device pointers (copied device-to-device)
static int bsumcheck_begin(sbn_ctx* ctx, sbn_prodcircuit* const* circuits, size_t P, size_t layer_id, const sbn_fr* rand,
                           size_t n_rand, const sbn_fr* const* seqA, const sbn_fr* const* seqB, const sbn_fr* const* seqC,
                           const Fr* const* seq_dev, size_t S, sbn_bsumcheck** out) {
    if (!ctx || !out || !circuits || P == 0 || (n_rand && !rand) || (S && !seq_dev && (!seqA || !seqB || !seqC))) return SBN_ERR_ARG;
    *out = nullptr;
    if (n_rand > 30 || P + S > 4096) return SBN_ERR_SHAPE;
    const size_t T = size_t(1) << n_rand;                      // table length: poly_C_par = eq(rand) has 2^|rand| entries
    for (size_t i = 0; i < P; i++) {
        if (!circuits[i] || circuits[i]->ctx != ctx) return SBN_ERR_ARG;
        if (layer_id >= (size_t)circuits[i]->num_layers) return SBN_ERR_SHAPE;
        if ((circuits[i]->len >> layer_id) != 2 * T) return SBN_ERR_SHAPE;      // product_tree.rs:272 assert_eq!(poly_C_par.len(), len / 2)
    }
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_ENTER(ctx);
    cudaStream_t s = ctx->compute;
    sbn_bsumcheck* st = new (std::nothrow) sbn_bsumcheck();
    if (!st) return SBN_ERR_OOM;
    st->ctx = ctx;
    st->P = P;
    st->S = S;
    st->len = T;
    const size_t n = P + S;
    st->max_blocks = (unsigned)std::max<size_t>(1, std::min<size_t>((T / 2 + kDotThreads - 1) / kDotThreads, std::max<size_t>(16, 1184 / n)));
    st->T0 = T;
    bool ok = pool_alloc(ctx, &st->eq[0], T * sizeof(Fr)) == cudaSuccess && pool_alloc(ctx, &st->eq[1], T * sizeof(Fr)) == cudaSuccess &&
              pool_alloc(ctx, &st->d_triples, n * sizeof(CubicTriple)) == cudaSuccess &&
              pool_alloc(ctx, &st->d_tables, (2 * P + 1 + 3 * S) * sizeof(Fr*)) == cudaSuccess &&
              pool_alloc(ctx, &st->partial, 3 * n * st->max_blocks * sizeof(Fr)) == cudaSuccess &&
              pool_alloc(ctx, &st->out, (3 * n + n_rand + 2) * sizeof(Fr)) == cudaSuccess &&
              (S == 0 || pool_alloc(ctx, &st->seq, 3 * S * T * sizeof(Fr)) == cudaSuccess);
    if (!ok) { bsumcheck_free(st); ctx->last_error = "sbn_bsumcheck_begin: cudaMalloc failed"; return SBN_ERR_OOM; }
    auto fail = [&](const char* what, cudaError_t e) {
        ctx->last_error = std::string(what) + ": " + cudaGetErrorString(e);
        bsumcheck_free(st);
        return SBN_ERR_CUDA;
    };
    cudaError_t e;
    // poly_C_par = EqPolynomial::new(rand).evals()                                  (hyrax.rs:355-369)
    Fr* rdev = st->out + 3 * n + 2;
    const Fr one = Fr::one();
    if ((e = cudaMemcpyAsync(st->eq[0], &one, sizeof(Fr), cudaMemcpyHostToDevice, s)) != cudaSuccess) return fail("eq seed", e);
    if (n_rand && (e = cudaMemcpyAsync(rdev, rand, n_rand * sizeof(Fr), cudaMemcpyHostToDevice, s)) != cudaSuccess) return fail("rand upload", e);
    ctx->h2d += (n_rand + 1) * sizeof(Fr);
    for (size_t j = 0; j < n_rand; j++) {
        const size_t size = size_t(1) << j;
        k_eq_expand<<<(unsigned)((size + 127) / 128), 128, 0, s>>>(st->eq[st->eq_cur], size, rdev + j, st->eq[st->eq_cur ^ 1]);
        st->eq_cur ^= 1;
        ctx->launches++;
    }
    for (size_t k = 0; k < S; k++) {
        for (int w = 0; w < 3; w++) {
            const void* src = seq_dev ? (const void*)seq_dev[3 * k + w] : (const void*)(w == 0 ? seqA[k] : w == 1 ? seqB[k] : seqC[k]);
            if (!src) { bsumcheck_free(st); return SBN_ERR_ARG; }
            if ((e = cudaMemcpyAsync(st->seq + (3 * k + w) * T, src, T * sizeof(Fr),
                                     seq_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s)) != cudaSuccess)
                return fail("sequential instance upload", e);
        }
        if (!seq_dev) ctx->h2d += 3 * T * sizeof(Fr);
    }
    std::vector<CubicTriple> triples(n);
    std::vector<Fr*> tables;
    for (size_t i = 0; i < P; i++) {
        Fr* layer = circuits[i]->buf + circuits[i]->off[layer_id];
        st->A.push_back(layer);                 // left_vec[layer_id]
        st->B.push_back(layer + T);             // right_vec[layer_id]
        st->C.push_back(st->eq[st->eq_cur]);
        tables.push_back(layer);
        tables.push_back(layer + T);
    }
    tables.push_back(st->eq[st->eq_cur]);
    for (size_t k = 0; k < S; k++) {
        for (int w = 0; w < 3; w++) tables.push_back(st->seq + (3 * k + w) * T);
        st->A.push_back(st->seq + (3 * k) * T);
        st->B.push_back(st->seq + (3 * k + 1) * T);
        st->C.push_back(st->seq + (3 * k + 2) * T);
    }
    for (size_t i = 0; i < n; i++) triples[i] = CubicTriple{st->A[i], st->B[i], st->C[i]};
    st->ntables = (int)tables.size();
    if ((e = cudaMemcpyAsync(st->d_triples, triples.data(), n * sizeof(CubicTriple), cudaMemcpyHostToDevice, s)) != cudaSuccess ||
        (e = cudaMemcpyAsync(st->d_tables, tables.data(), tables.size() * sizeof(Fr*), cudaMemcpyHostToDevice, s)) != cudaSuccess ||
        (e = cudaGetLastError()) != cudaSuccess || (e = cudaStreamSynchronize(s)) != cudaSuccess)
        return fail("sbn_bsumcheck_begin", e);
    *out = st;
    return SBN_OK;
}

extern "C" int sbn_bsumcheck_begin(sbn_ctx* ctx, sbn_prodcircuit* const* circuits, size_t P, size_t layer_id, const sbn_fr* rand,
                                   size_t n_rand, const sbn_fr* const* seqA, const sbn_fr* const* seqB,
                                   const sbn_fr* const* seqC, size_t S, sbn_bsumcheck** out) {
    return bsumcheck_begin(ctx, circuits, P, layer_id, rand, n_rand, seqA, seqB, seqC, nullptr, S, out);
}

extern "C" int sbn_bsumcheck_begin_resident(sbn_ctx* ctx, sbn_prodcircuit* const* circuits, size_t P, size_t layer_id,
                                            const sbn_fr* rand, size_t n_rand, const sbn_poly* const* seq_polys,
                                            const size_t* seq_offsets, size_t S, sbn_bsumcheck** out) {
    if (S && (!seq_polys || !seq_offsets)) return SBN_ERR_ARG;
    if (n_rand > 30) return SBN_ERR_SHAPE;
    const size_t T = size_t(1) << n_rand;
    std::vector<const Fr*> dev(3 * S);
    for (size_t i = 0; i < 3 * S; i++) {
        if (!seq_polys[i] || seq_polys[i]->ctx != ctx) return SBN_ERR_ARG;
        if (seq_offsets[i] > seq_polys[i]->len || T > seq_polys[i]->len - seq_offsets[i]) return SBN_ERR_SHAPE;
        dev[i] = seq_polys[i]->Z + seq_offsets[i];
    }
    return bsumcheck_begin(ctx, circuits, P, layer_id, rand, n_rand, nullptr, nullptr, nullptr, S ? dev.data() : nullptr, S, out);
}

extern "C" int sbn_bsumcheck_round_eval(sbn_bsumcheck* st, sbn_fr* evals) {
    if (!st || !evals) return SBN_ERR_ARG;
    if (st->len < 2) return SBN_ERR_SHAPE;
    sbn_ctx* ctx = st->ctx;
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_ENTER(ctx);
    cudaStream_t s = ctx->compute;
    const size_t half = st->len / 2, n = st->P + st->S;
    const unsigned blocks = (unsigned)std::min<size_t>(st->max_blocks, (half + kDotThreads - 1) / kDotThreads);
    k_cubic_eval_batched<<<dim3(blocks, (unsigned)n), kDotThreads, 0, s>>>(st->d_triples, half, st->partial);
    k_fr_sum<<<(unsigned)(3 * n), kDotThreads, 0, s>>>(st->partial, (int)blocks, st->out, 1);
    ctx->launches += 2;
    SBN_CUDA(ctx, cudaGetLastError());
    SBN_CUDA(ctx, cudaMemcpyAsync(evals, st->out, 3 * n * sizeof(Fr), cudaMemcpyDeviceToHost, s));
    SBN_CUDA(ctx, cudaStreamSynchronize(s));
    ctx->d2h += 3 * n * sizeof(Fr);
    return SBN_OK;
}

extern "C" int sbn_bsumcheck_bind(sbn_bsumcheck* st, const sbn_fr* r) {
    if (!st || !r) return SBN_ERR_ARG;
    if (st->len < 2) return SBN_ERR_SHAPE;
    sbn_ctx* ctx = st->ctx;
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_ENTER(ctx);
    cudaStream_t s = ctx->compute;
    const size_t half = st->len / 2, n = st->P + st->S;
    Fr* rdev = st->out + 3 * n;
    SBN_CUDA(ctx, cudaMemcpyAsync(rdev, r, sizeof(Fr), cudaMemcpyHostToDevice, s));
    k_bind_top_batched<<<dim3((unsigned)((half + 127) / 128), (unsigned)st->ntables), 128, 0, s>>>(st->d_tables, half, rdev);
    ctx->launches += 1;
    ctx->h2d += sizeof(Fr);
    SBN_CUDA(ctx, cudaGetLastError());
    st->len = half;
    return SBN_OK;
}

extern "C" int sbn_bsumcheck_end(sbn_bsumcheck* st, sbn_fr* A_final, sbn_fr* B_final, sbn_fr* C_final) {
    if (!st || !A_final || !B_final || !C_final) return SBN_ERR_ARG;
    if (st->len != 1) return SBN_ERR_SHAPE;
    sbn_ctx* ctx = st->ctx;
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_ENTER(ctx);
    cudaStream_t s = ctx->compute;
    const size_t n = st->P + st->S;
    for (size_t i = 0; i < n; i++) {
        SBN_CUDA(ctx, cudaMemcpyAsync(A_final + i, st->A[i], sizeof(Fr), cudaMemcpyDeviceToHost, s));
        SBN_CUDA(ctx, cudaMemcpyAsync(B_final + i, st->B[i], sizeof(Fr), cudaMemcpyDeviceToHost, s));
    }
    // C_final[0] = poly_C_par[0], then one per sequential instance            (sumcheck.rs:309-327)
    SBN_CUDA(ctx, cudaMemcpyAsync(C_final, st->eq[st->eq_cur], sizeof(Fr), cudaMemcpyDeviceToHost, s));
    for (size_t k = 0; k < st->S; k++)
        SBN_CUDA(ctx, cudaMemcpyAsync(C_final + 1 + k, st->C[st->P + k], sizeof(Fr), cudaMemcpyDeviceToHost, s));
    SBN_CUDA(ctx, cudaStreamSynchronize(s));
    ctx->d2h += (2 * n + 1 + st->S) * sizeof(Fr);
    return SBN_OK;
}

// Framing of one sumcheck round's transcript traffic for the device-side transcript (transcript_kernels.cuh): a recording run
// of exactly the operations csrc/host/merlin.hpp performs for UniPoly::append_to_transcript (unipoly.rs:119-127) followed by
// challenge_scalar (transcript.rs:56-67), from the cursor (pos, pos_begin), with zeros in place of the coefficient bytes.
static void bsc_make_frame(int pos, int pos_begin, BscFrame* f) {
    memset(f, 0, sizeof *f);
    size_t n = 0, chunk_start = 0;
    int chunk_pos = pos;
    auto close_chunk = [&](int fbegin) {
        const int c = f->nchunks++;
        f->off[c] = (uint16_t)chunk_start;
        f->len[c] = (uint16_t)(n - chunk_start);
        f->pos[c] = (uint16_t)chunk_pos;
        f->f_begin[c] = (uint16_t)fbegin;
        chunk_start = n;
    };
    auto run_f = [&]() { close_chunk(pos_begin); pos = 0; pos_begin = 0; chunk_pos = 0; };
    auto absorb = [&](const uint8_t* d, size_t len) {
        for (size_t i = 0; i < len; i++) {
            f->stream[n++] = d ? d[i] : 0;
            if (++pos == sbn::merlin::kRate) run_f();
        }
    };
    auto begin_op = [&](uint8_t flags) {
        const uint8_t hdr[2] = {(uint8_t)pos_begin, flags};
        pos_begin = pos + 1;
        absorb(hdr, 2);
        if ((flags & (sbn::merlin::FLAG_C | sbn::merlin::FLAG_K)) && pos != 0) run_f();
    };
    auto meta_len = [&](uint32_t v) { const uint8_t b[4] = {(uint8_t)v, (uint8_t)(v >> 8), (uint8_t)(v >> 16), (uint8_t)(v >> 24)}; absorb(b, 4); };
    auto append_message = [&](const char* label, const char* msg, size_t mlen) -> size_t {
        begin_op(sbn::merlin::FLAG_M | sbn::merlin::FLAG_A);
        absorb((const uint8_t*)label, strlen(label));
        meta_len((uint32_t)mlen);
        begin_op(sbn::merlin::FLAG_A);
        const size_t at = n;
        absorb((const uint8_t*)msg, mlen);
        return at;
    };
    append_message("poly", "UniPoly_begin", 13);
    for (int k = 0; k < 4; k++) f->coeff_off[k] = (uint16_t)append_message("coeff", nullptr, 32);
    append_message("poly", "UniPoly_end", 11);
    begin_op(sbn::merlin::FLAG_M | sbn::merlin::FLAG_A);
    absorb((const uint8_t*)"challenge_nextround", 19);
    meta_len(64);
    begin_op(sbn::merlin::FLAG_I | sbn::merlin::FLAG_A | sbn::merlin::FLAG_C);        // ends in F: the squeeze starts at position 0
    if (n > chunk_start) close_chunk(0xffff);
    f->total = (uint16_t)n;
}

// ---- host-side field helpers of the round loop below (Montgomery Fr; the same fp.cuh code compiled for the host)
namespace {
struct FrHost {
    static Fr small(uint32_t v) { Fr x = Fr::zero(); x.l[0] = v; return fp_to_mont(x); }
    static Fr mul_small(const Fr& a, int k) { Fr r = a; for (int i = 1; i < k; i++) r = fp_add(r, a); return r; }
    // from_le_bytes_mod_order of 64 bytes (transcript.rs:56-67): lo + hi * 2^256, the Montgomery form of 2^256 being R^2 mod r
    static Fr from_wide(const uint8_t* b) {
        Fr lo, hi, r2;
        memcpy(lo.l, b, 32);
        memcpy(hi.l, b + 32, 32);
        for (int i = 0; i < 8; i++) r2.l[i] = FrParams::R2(i);
        return fp_add(fp_to_mont(lo), fp_mul(fp_to_mont(hi), r2));
    }
};
}  // namespace

// n challenge scalars in one call, in Montgomery form: RandomTape::random_vector (random.rs:24-31) / challenge_vector
// (transcript.rs:69-73) -- the 1024 row blinds of R1CSProof::commit_poly were 1024 round trips through the binding.
extern "C" void sbn_merlin_challenge_scalars(void* state, const uint8_t* label, size_t llen, size_t n, sbn_fr* out) {
    sbn::merlin::State& tr = *(sbn::merlin::State*)state;
    for (size_t i = 0; i < n; i++) {
        uint8_t wide[64];
        sbn::merlin::challenge_bytes(tr, label, llen, wide, 64);
        const Fr r = FrHost::from_wide(wide);
        memcpy(out + i, &r, sizeof(Fr));
    }
}

// SumcheckInstanceProof::prove_cubic_batched (sumcheck.rs:165-330) for one layer with the round loop inside the library:
// per round the batched evaluation (:201-271), the combination with `coeffs` (:273-275), UniPoly::from_evals
// (unipoly.rs:28-59), the transcript append (unipoly.rs:119-127), the challenge (transcript.rs:56-67) and the bind
// (:293-306); then the final values (:309-327).  `merlin` is the 203-byte transcript state the sbn_merlin_* calls drive.
// polys: num_rounds x 4 coefficients, lowest degree first; r_out: the challenges; claim_out: the claim after the last round.
extern "C" int sbn_bsumcheck_prove(sbn_bsumcheck* st, void* merlin, const sbn_fr* claim, const sbn_fr* coeffs, size_t num_rounds,
                                   sbn_fr* polys, sbn_fr* r_out, sbn_fr* claim_out, sbn_fr* A_final, sbn_fr* B_final,
                                   sbn_fr* C_final) {
    if (!st || !merlin || !claim || !coeffs || !polys || !r_out || !claim_out || !A_final || !B_final || !C_final) return SBN_ERR_ARG;
    if (st->len != (size_t(1) << num_rounds)) return SBN_ERR_SHAPE;
    sbn_ctx* ctx = st->ctx;
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_ENTER(ctx);
    cudaStream_t s = ctx->compute;
    sbn::merlin::State& tr = *(sbn::merlin::State*)merlin;
    const size_t n = st->P + st->S;
    static const Fr two_inv = fp_inv(FrHost::small(2)), six_inv = fp_inv(FrHost::small(6));
    std::vector<Fr> ev(3 * n), cf(n);
    memcpy(cf.data(), coeffs, n * sizeof(Fr));
    Fr e;
    memcpy(e.l, claim, sizeof(Fr));
    if (3 * n * sizeof(Fr) > kEvPinBytes) return SBN_ERR_SHAPE;
    if (!ctx->ev_pin) {
        SBN_CUDA(ctx, cudaHostAlloc((void**)&ctx->ev_pin, kEvPinBytes, cudaHostAllocMapped));
        SBN_CUDA(ctx, cudaHostGetDevicePointer((void**)&ctx->ev_pin_dev, ctx->ev_pin, 0));
    }
    auto append_scalar = [&](const Fr& v) {
        const Fr c = fp_from_mont(v);
        sbn::merlin::append_message(tr, (const uint8_t*)"coeff", 5, (const uint8_t*)c.l, 32);
    };
    // Rounds whose tables are still long go through the host loop below (evaluation launch, synchronisation, transcript on the
    // host, bind launch: ~55 us a round, of which the kernels are the larger part); once half a table is kBscTailHalf entries
    // or fewer, ONE block finishes the layer with the transcript on the device (~20 us a round, no launches).  Measured at
    // 12 x 2^22 (scripts/exp_bsc.py): a one-warp transcript step costs ~16 us on the GPU (Keccak-f and Montgomery products are
    // latency chains there), so for the long rounds the device-side transcript (bsc_device = 2) is no faster than the host's.
    size_t host_rounds = num_rounds;
    if (ctx->bsc_device && n <= (size_t)kBscTailMaxInst && num_rounds <= 30) {
        host_rounds = 0;
        if (ctx->bsc_device == 1)
            while (host_rounds < num_rounds && ((st->len >> host_rounds) / 2) > (size_t)kBscTailHalf) host_rounds++;
    }
    // fused rounds (prodtree_kernels.cuh, k_bind_eval_batched): the bind of a round rides in the next round's evaluation
    // kernel, so a round is one launch on the Fiat-Shamir chain instead of two; the shared eq table ping-pongs between its two
    // buffers and the last round's bind is a launch of its own
    const bool fused = ctx->fused_rounds && host_rounds == num_rounds;
    bool pending = false;
    Fr r_prev = Fr::zero();
    Fr* const eq_baked = st->eq[st->eq_cur];
    int eq_now = st->eq_cur;
    // (Measured with timers around the three parts of a round at 12 x 2^22: launches 6 us, host 4 us, waiting for the round's
    // kernels 54 us, of which their own run time averages 36 -- the long rounds of the upper layers.)
    for (size_t j = 0; j < host_rounds; j++) {
        const size_t half = st->len / 2;
        const unsigned blocks = (unsigned)std::min<size_t>(st->max_blocks, (half + kDotThreads - 1) / kDotThreads);
        // The 3 n evaluations land in mapped pinned memory: no copy call per round, and a one-block evaluation (the many short
        // rounds at the end of every layer) writes them itself instead of through a second launch.
        Fr* ev_dev = (Fr*)ctx->ev_pin_dev;
        Fr* ev_to = blocks == 1 ? ev_dev : st->partial;
        if (pending) {
            k_bind_eval_batched<<<dim3(blocks, (unsigned)n), kDotThreads, 0, s>>>(st->d_triples, (int)st->P, st->eq[eq_now], st->eq[eq_now ^ 1],
                                                                                 half, r_prev, ev_to);
            eq_now ^= 1;
        } else {
            k_cubic_eval_batched<<<dim3(blocks, (unsigned)n), kDotThreads, 0, s>>>(st->d_triples, half, ev_to);
        }
        ctx->launches += 1;
        if (blocks > 1) {
            k_fr_sum<<<(unsigned)(3 * n), kDotThreads, 0, s>>>(st->partial, (int)blocks, ev_dev, 1);
            ctx->launches += 1;
        }
        SBN_CUDA(ctx, cudaGetLastError());
        SBN_CUDA(ctx, cudaStreamSynchronize(s));
        memcpy(ev.data(), ctx->ev_pin, 3 * n * sizeof(Fr));
        ctx->d2h += 3 * n * sizeof(Fr);
        Fr comb[3] = {Fr::zero(), Fr::zero(), Fr::zero()};
        for (size_t i = 0; i < n; i++)
            for (int k = 0; k < 3; k++) comb[k] = fp_add(comb[k], fp_mul(ev[3 * i + k], cf[i]));
        // evaluations at 0, 1, 2, 3 -> coefficients (unipoly.rs:28-59, the cubic branch)
        const Fr e0 = comb[0], e1 = fp_sub(e, comb[0]), e2 = comb[1], e3 = comb[2];
        const Fr d = e0;
        const Fr a = fp_mul(six_inv, fp_sub(fp_add(fp_sub(e3, FrHost::mul_small(e2, 3)), FrHost::mul_small(e1, 3)), e0));
        const Fr b = fp_mul(two_inv, fp_sub(fp_add(fp_sub(FrHost::mul_small(e0, 2), FrHost::mul_small(e1, 5)), FrHost::mul_small(e2, 4)), e3));
        const Fr c = fp_sub(fp_sub(fp_sub(e1, d), a), b);
        const Fr poly[4] = {d, c, b, a};
        sbn::merlin::append_message(tr, (const uint8_t*)"poly", 4, (const uint8_t*)"UniPoly_begin", 13);
        for (int k = 0; k < 4; k++) append_scalar(poly[k]);
        sbn::merlin::append_message(tr, (const uint8_t*)"poly", 4, (const uint8_t*)"UniPoly_end", 11);
        uint8_t wide[64];
        sbn::merlin::challenge_bytes(tr, (const uint8_t*)"challenge_nextround", 19, wide, 64);
        const Fr r = FrHost::from_wide(wide);
        if (fused && j + 1 < host_rounds) {
            pending = true;                       // bound by the next round's kernel
            r_prev = r;
        } else {
            if (eq_now != st->eq_cur) {           // the tables' own eq pointer is the other buffer: bring the 2 * half live entries home
                SBN_CUDA(ctx, cudaMemcpyAsync(eq_baked, st->eq[eq_now], 2 * half * sizeof(Fr), cudaMemcpyDeviceToDevice, s));
                eq_now = st->eq_cur;
            }
            k_bind_top_batched_v<<<dim3((unsigned)((half + 127) / 128), (unsigned)st->ntables), 128, 0, s>>>(st->d_tables, half, r);
            ctx->launches += 1;
            SBN_CUDA(ctx, cudaGetLastError());
            pending = false;
        }
        st->len = half;
        // e = poly(r)
        Fr acc = poly[0], power = r;
        for (int k = 1; k < 4; k++) {
            acc = fp_add(acc, fp_mul(power, poly[k]));
            power = fp_mul(power, r);
        }
        e = acc;
        memcpy(polys + 4 * j, poly, 4 * sizeof(Fr));
        memcpy(r_out + j, &r, sizeof(Fr));
    }
    if (host_rounds < num_rounds) {
        const size_t j0 = host_rounds, nr = num_rounds - host_rounds;      // rounds j0 .. num_rounds - 1 run on the device
        // The whole layer on the device: transcript state, claim, coefficients and the framing of a round's transcript traffic
        // go up once; per round the evaluation, the transcript step (k_bsc_round) and the bind are three launches in stream
        // order with no host round trip, and once the tables are short ONE block finishes every remaining round
        // (k_bsc_tail); polynomials, challenges, final values and the transcript state come back in one copy.
        const size_t nt = (size_t)st->ntables;
        const size_t sw = (sizeof(BscState) + sizeof(Fr) - 1) / sizeof(Fr);
        const size_t words = sw + n + 4 * nr + nr + 1 + nt;
        Fr* buf = nullptr;
        if (pool_alloc(ctx, &buf, words * sizeof(Fr)) != cudaSuccess) { ctx->last_error = "sbn_bsumcheck_prove: cudaMalloc failed"; return SBN_ERR_OOM; }
        BscState* dstate = (BscState*)buf;
        Fr *dcf = buf + sw, *dpolys = dcf + n, *dr = dpolys + 4 * nr, *drcur = dr + nr, *dfin = drcur + 1;
        std::vector<Fr> up(sw + n);
        memset(up.data(), 0, sw * sizeof(Fr));
        BscState* hs = (BscState*)up.data();
        memcpy(hs->merlin, &tr, sizeof(sbn::merlin::State));
        hs->claim = e;
        bsc_make_frame(tr.pos, tr.pos_begin, &hs->frame[0]);
        bsc_make_frame(64, 0, &hs->frame[1]);
        memcpy(up.data() + sw, cf.data(), n * sizeof(Fr));
        auto fail = [&](int code) { cudaStreamSynchronize(s); cudaGetLastError(); pool_release(ctx, buf); return code; };
        if (cudaMemcpyAsync(buf, up.data(), (sw + n) * sizeof(Fr), cudaMemcpyHostToDevice, s) != cudaSuccess) return fail(SBN_ERR_CUDA);
        ctx->h2d += (sw + n) * sizeof(Fr);
        BscConst kc{two_inv, six_inv};
        bool tail = false;
        for (size_t j = 0; j < nr; j++) {
            const size_t half = st->len / 2;
            if (half <= (size_t)kBscTailHalf) {
                auto tailk = n <= 16 ? k_bsc_tail<512> : k_bsc_tail<1024>;
                tailk<<<1, (unsigned)(32 * n), 0, s>>>(dstate, st->d_triples, st->d_tables, st->ntables, (int)n, (int)st->len, j == 0 ? 1 : 0,
                                                           dcf, kc, dpolys + 4 * j, dr + j, dfin);
                ctx->launches += 1;
                st->len = 1;
                tail = true;
                break;
            }
            const unsigned blocks = (unsigned)std::min<size_t>(st->max_blocks, (half + kDotThreads - 1) / kDotThreads);
            k_cubic_eval_batched<<<dim3(blocks, (unsigned)n), kDotThreads, 0, s>>>(st->d_triples, half, st->partial);
            k_bsc_round<<<1, 128, 0, s>>>(dstate, st->partial, (int)blocks, dcf, (int)n, j == 0 ? 1 : 0, kc, dpolys + 4 * j, dr + j, drcur);
            k_bind_top_batched<<<dim3((unsigned)((half + 127) / 128), (unsigned)st->ntables), 128, 0, s>>>(st->d_tables, half, drcur);
            ctx->launches += 3;
            st->len = half;
        }
        if (!tail) {
            k_gather_first<<<(unsigned)((st->ntables + 127) / 128), 128, 0, s>>>(st->d_tables, st->ntables, dfin);
            ctx->launches += 1;
        }
        if (cudaGetLastError() != cudaSuccess) { ctx->last_error = "sbn_bsumcheck_prove: launch failed"; return fail(SBN_ERR_CUDA); }
        std::vector<Fr> down(words);
        if (cudaMemcpyAsync(down.data(), buf, words * sizeof(Fr), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
            cudaStreamSynchronize(s) != cudaSuccess) {
            ctx->last_error = std::string("sbn_bsumcheck_prove: ") + cudaGetErrorString(cudaGetLastError());
            return fail(SBN_ERR_CUDA);
        }
        ctx->d2h += words * sizeof(Fr);
        pool_release(ctx, buf);
        const BscState* ds = (const BscState*)down.data();
        memcpy(&tr, ds->merlin, sizeof(sbn::merlin::State));
        memcpy(claim_out, &ds->claim, sizeof(Fr));
        memcpy(polys + 4 * j0, down.data() + sw + n, 4 * nr * sizeof(Fr));
        memcpy(r_out + j0, down.data() + sw + n + 4 * nr, nr * sizeof(Fr));
        const Fr* fin = down.data() + sw + n + 5 * nr + 1;
        for (size_t i = 0; i < st->P; i++) {
            memcpy(A_final + i, &fin[2 * i], sizeof(Fr));
            memcpy(B_final + i, &fin[2 * i + 1], sizeof(Fr));
        }
        memcpy(C_final, &fin[2 * st->P], sizeof(Fr));
        for (size_t k = 0; k < st->S; k++) {
            memcpy(A_final + st->P + k, &fin[2 * st->P + 1 + 3 * k], sizeof(Fr));
            memcpy(B_final + st->P + k, &fin[2 * st->P + 2 + 3 * k], sizeof(Fr));
            memcpy(C_final + 1 + k, &fin[2 * st->P + 3 + 3 * k], sizeof(Fr));
        }
        return SBN_OK;
    }
    memcpy(claim_out, &e, sizeof(Fr));
    // final values: element 0 of every table, gathered by one kernel and one copy
    std::vector<Fr> fin(st->ntables);
    Fr* gbuf = st->partial;              // scratch of 3 * n * max_blocks >= ntables scalars
    k_gather_first<<<(unsigned)((st->ntables + 127) / 128), 128, 0, s>>>(st->d_tables, st->ntables, gbuf);
    ctx->launches += 1;
    SBN_CUDA(ctx, cudaGetLastError());
    SBN_CUDA(ctx, cudaMemcpyAsync(fin.data(), gbuf, st->ntables * sizeof(Fr), cudaMemcpyDeviceToHost, s));
    SBN_CUDA(ctx, cudaStreamSynchronize(s));
    ctx->d2h += st->ntables * sizeof(Fr);
    // table order: (A_i, B_i) per parallel instance, eq, then (A, B, C) per sequential instance
    for (size_t i = 0; i < st->P; i++) {
        memcpy(A_final + i, &fin[2 * i], sizeof(Fr));
        memcpy(B_final + i, &fin[2 * i + 1], sizeof(Fr));
    }
    memcpy(C_final, &fin[2 * st->P], sizeof(Fr));
    for (size_t k = 0; k < st->S; k++) {
        memcpy(A_final + st->P + k, &fin[2 * st->P + 1 + 3 * k], sizeof(Fr));
        memcpy(B_final + st->P + k, &fin[2 * st->P + 2 + 3 * k], sizeof(Fr));
        memcpy(C_final + 1 + k, &fin[2 * st->P + 3 + 3 * k], sizeof(Fr));
    }
    return SBN_OK;
}

extern "C" int sbn_bsumcheck_destroy(sbn_bsumcheck* st) {
    if (!st) return SBN_ERR_ARG;
    {
        std::lock_guard<std::mutex> g(st->ctx->mu);
        cudaSetDevice(st->ctx->device);
        cudaStreamSynchronize(st->ctx->compute);
    }
    bsumcheck_free(st);
    return SBN_OK;
}

// ------------------------------------------------------------------------------------------------
// f2: derefs built on the device (address vectors resident, eq tables generated in HBM)
// ------------------------------------------------------------------------------------------------
struct sbn_addrs {
    sbn_ctx* ctx = nullptr;
    size_t batch = 0, N = 0;
    uint32_t *row = nullptr, *col = nullptr;
    uint32_t max_row = 0, max_col = 0;
    // memory-checking timestamps (AddrTimestamps::new, sparse_mlpoly_full.rs:212-243): read_ts per operation, audit_ts per cell
    size_t num_cells = 0;
    uint32_t *read_ts[2] = {nullptr, nullptr}, *audit_ts[2] = {nullptr, nullptr};
};

extern "C" int sbn_addrs_upload(sbn_ctx* ctx, const uint32_t* row_addrs, const uint32_t* col_addrs, size_t batch, size_t N,
                                sbn_addrs** out) {
    if (!ctx || !row_addrs || !col_addrs || !out) return SBN_ERR_ARG;
    *out = nullptr;
    if (batch == 0 || N == 0 || batch * N > (size_t(1) << 31)) return SBN_ERR_SHAPE;
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_ENTER(ctx);
    sbn_addrs* a = new (std::nothrow) sbn_addrs();
    if (!a) return SBN_ERR_OOM;
    a->ctx = ctx;
    a->batch = batch;
    a->N = N;
    for (size_t i = 0; i < batch * N; i++) {           // the bound checked against the eq-table sizes at every gather
        a->max_row = std::max(a->max_row, row_addrs[i]);
        a->max_col = std::max(a->max_col, col_addrs[i]);
    }
    const size_t bytes = batch * N * sizeof(uint32_t);
    if (dev_malloc(ctx, &a->row, bytes) != cudaSuccess || dev_malloc(ctx, &a->col, bytes) != cudaSuccess ||
        cudaMemcpyAsync(a->row, row_addrs, bytes, cudaMemcpyHostToDevice, ctx->compute) != cudaSuccess ||
        cudaMemcpyAsync(a->col, col_addrs, bytes, cudaMemcpyHostToDevice, ctx->compute) != cudaSuccess ||
        cudaStreamSynchronize(ctx->compute) != cudaSuccess) {
        if (a->row) cudaFree(a->row);
        if (a->col) cudaFree(a->col);
        delete a;
        ctx->last_error = "sbn_addrs_upload failed";
        return SBN_ERR_CUDA;
    }
    ctx->h2d += 2 * bytes;
    *out = a;
    return SBN_OK;
}

extern "C" int sbn_addrs_destroy(sbn_addrs* a) {
    if (!a) return SBN_ERR_ARG;
    {
        std::lock_guard<std::mutex> g(a->ctx->mu);
        cudaSetDevice(a->ctx->device);
        cudaStreamSynchronize(a->ctx->compute);
        cudaFree(a->row);
        cudaFree(a->col);
        for (int k = 0; k < 2; k++) {
            if (a->read_ts[k]) cudaFree(a->read_ts[k]);
            if (a->audit_ts[k]) cudaFree(a->audit_ts[k]);
        }
    }
    delete a;
    return SBN_OK;
}

extern "C" int sbn_addrs_set_timestamps(sbn_addrs* a, const uint32_t* row_read_ts, const uint32_t* row_audit_ts,
                                        const uint32_t* col_read_ts, const uint32_t* col_audit_ts, size_t num_cells) {
    if (!a || !row_read_ts || !row_audit_ts || !col_read_ts || !col_audit_ts) return SBN_ERR_ARG;
    if (num_cells == 0 || (num_cells & (num_cells - 1)) || a->max_row >= num_cells || a->max_col >= num_cells) return SBN_ERR_SHAPE;
    sbn_ctx* ctx = a->ctx;
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_ENTER(ctx);
    const uint32_t* src_r[2] = {row_read_ts, col_read_ts};
    const uint32_t* src_a[2] = {row_audit_ts, col_audit_ts};
    const size_t ops_bytes = a->batch * a->N * sizeof(uint32_t), mem_bytes = num_cells * sizeof(uint32_t);
    for (int k = 0; k < 2; k++) {
        if (!a->read_ts[k]) SBN_CUDA(ctx, dev_malloc(ctx, &a->read_ts[k], ops_bytes));
        if (a->audit_ts[k] && a->num_cells != num_cells) { cudaFree(a->audit_ts[k]); a->audit_ts[k] = nullptr; }
        if (!a->audit_ts[k]) SBN_CUDA(ctx, dev_malloc(ctx, &a->audit_ts[k], mem_bytes));
        SBN_CUDA(ctx, cudaMemcpyAsync(a->read_ts[k], src_r[k], ops_bytes, cudaMemcpyHostToDevice, ctx->compute));
        SBN_CUDA(ctx, cudaMemcpyAsync(a->audit_ts[k], src_a[k], mem_bytes, cudaMemcpyHostToDevice, ctx->compute));
    }
    SBN_CUDA(ctx, cudaStreamSynchronize(ctx->compute));
    ctx->h2d += 2 * (ops_bytes + mem_bytes);
    a->num_cells = num_cells;
    return SBN_OK;
}

// eq(r) evaluations into `dst` (2^n scalars), using `tmp` (same size) as the other ping-pong buffer; r already on device
static Fr* eq_evals_device(sbn_ctx* ctx, const Fr* r_dev, size_t n, Fr* dst, Fr* tmp, cudaStream_t s) {
    Fr* buf[2] = {(n & 1) ? tmp : dst, (n & 1) ? dst : tmp};       // an even/odd number of steps ends in dst
    const Fr one = Fr::one();
    cudaMemcpyAsync(buf[0], &one, sizeof(Fr), cudaMemcpyHostToDevice, s);
    int cur = 0;
    for (size_t j = 0; j < n; j++) {
        const size_t size = size_t(1) << j;
        k_eq_expand<<<(unsigned)((size + 127) / 128), 128, 0, s>>>(buf[cur], size, r_dev + j, buf[cur ^ 1]);
        cur ^= 1;
        ctx->launches++;
    }
    return buf[cur];
}

static int derefs_commit_rows(sbn_ctx* ctx, const sbn_bases* b, const sbn_addrs* addrs, const sbn_fr* rx, size_t nx,
                              const sbn_fr* ry, size_t ny, size_t row0, size_t nrows, bool all_rows, sbn_g1a* C_out,
                              uint8_t* inf_out, sbn_poly** poly_out) {
    if (!ctx || !b || !addrs || !rx || !ry || !C_out || !inf_out || b->ctx != ctx || addrs->ctx != ctx) return SBN_ERR_ARG;
    if (nx == 0 || ny == 0 || nx > 30 || ny > 30) return SBN_ERR_SHAPE;
    if (addrs->max_row >= (size_t(1) << nx) || addrs->max_col >= (size_t(1) << ny)) return SBN_ERR_SHAPE;   // would index past mem_rx / mem_ry
    // comb = merge(row_ops_val ++ col_ops_val) zero-padded to a power of two (hyrax.rs:237-247); Hyrax shape hyrax.rs:292
    const size_t used = 2 * addrs->batch * addrs->N;
    size_t len = 1;
    int ell = 0;
    while (len < used) { len <<= 1; ell++; }
    const size_t Lfull = size_t(1) << (ell / 2), R = len / Lfull;
    if (all_rows) { row0 = 0; nrows = Lfull; }
    if (nrows == 0 || row0 > Lfull || nrows > Lfull - row0) return SBN_ERR_SHAPE;
    const size_t L = nrows;                               // rows committed by this call (a rank's block of the Hyrax matrix)
    SBN_TRY(check_commit_shape(b, L, R));
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_ENTER(ctx);
    cudaStream_t s = ctx->compute;
    sbn_poly* p = new (std::nothrow) sbn_poly();
    if (!p) return SBN_ERR_OOM;
    p->ctx = ctx;
    p->len = len;
    if (pool_alloc(ctx, &p->Z, len * sizeof(Fr)) != cudaSuccess) { delete p; ctx->last_error = "sbn_derefs_commit: cudaMalloc failed"; return SBN_ERR_OOM; }
    // the commit reads p->Z on the pipeline streams, which join `compute` only at its end: drain them before the release
    auto fail = [&](int code) { cudaDeviceSynchronize(); cudaGetLastError(); pool_free(ctx, p->Z, len * sizeof(Fr)); delete p; return code; };
    const size_t tx = size_t(1) << nx, ty = size_t(1) << ny;
    int rc;
    if ((rc = ensure(ctx, ctx->scratch0, 2 * tx * sizeof(Fr))) != SBN_OK || (rc = ensure(ctx, ctx->scratch1, 2 * ty * sizeof(Fr))) != SBN_OK ||
        (rc = ensure(ctx, ctx->scratch2, (nx + ny) * sizeof(Fr))) != SBN_OK)
        return fail(rc);
    Fr* rdev = (Fr*)ctx->scratch2.p;
    if (cudaMemcpyAsync(rdev, rx, nx * sizeof(Fr), cudaMemcpyHostToDevice, s) != cudaSuccess ||
        cudaMemcpyAsync(rdev + nx, ry, ny * sizeof(Fr), cudaMemcpyHostToDevice, s) != cudaSuccess) {
        ctx->last_error = "sbn_derefs_commit: upload failed";
        return fail(SBN_ERR_CUDA);
    }
    ctx->h2d += (nx + ny) * sizeof(Fr);
    const Fr* mem_rx = eq_evals_device(ctx, rdev, nx, (Fr*)ctx->scratch0.p, (Fr*)ctx->scratch0.p + tx, s);
    const Fr* mem_ry = eq_evals_device(ctx, rdev + nx, ny, (Fr*)ctx->scratch1.p, (Fr*)ctx->scratch1.p + ty, s);
    k_derefs_gather<<<(unsigned)((len + 255) / 256), 256, 0, s>>>(mem_rx, mem_ry, addrs->row, addrs->col, addrs->batch, addrs->N, len, p->Z);
    ctx->launches++;
    if (cudaGetLastError() != cudaSuccess) { ctx->last_error = "k_derefs_gather launch failed"; return fail(SBN_ERR_CUDA); }
    const size_t chunk = commit_chunk_rows(ctx, L);
    if ((rc = ensure_commit_workspace(ctx, b, chunk, L)) != SBN_OK || (rc = ensure(ctx, ctx->dC, L * sizeof(Affine))) != SBN_OK ||
        (rc = ensure(ctx, ctx->dinf, L)) != SBN_OK)
        return fail(rc);
    std::vector<int> ev_stage;
    if ((rc = run_commit(ctx, b, p->Z + row0 * R, nullptr, L, R, nullptr, (Affine*)ctx->dC.p, (uint8_t*)ctx->dinf.p, s, ev_stage)) != SBN_OK)
        return fail(rc);
    if ((rc = download(ctx, C_out, ctx->dC.p, L * sizeof(Affine))) != SBN_OK || (rc = download(ctx, inf_out, ctx->dinf.p, L)) != SBN_OK)
        return fail(rc);
    if (cudaStreamSynchronize(s) != cudaSuccess) { ctx->last_error = "sbn_derefs_commit: synchronize failed"; return fail(SBN_ERR_CUDA); }
    collect_profile(ctx, ev_stage);
    if (poly_out) *poly_out = p;        // the derefs polynomial stays resident for its opening (DerefsEvalProof)
    else { pool_free(ctx, p->Z, len * sizeof(Fr)); delete p; }
    return SBN_OK;
}

extern "C" int sbn_derefs_commit(sbn_ctx* ctx, const sbn_bases* b, const sbn_addrs* addrs, const sbn_fr* rx, size_t nx,
                                 const sbn_fr* ry, size_t ny, sbn_g1a* C_out, uint8_t* inf_out, sbn_poly** poly_out) {
    return derefs_commit_rows(ctx, b, addrs, rx, nx, ry, ny, 0, 0, true, C_out, inf_out, poly_out);
}
// Multi-GPU form: the whole derefs polynomial is built (it is needed for the opening), rows [row0, row0 + nrows) of its
// Hyrax matrix are committed; the caller gathers the row blocks of the ranks (hyrax.rs:259-265: rows are independent).
extern "C" int sbn_derefs_commit_rows(sbn_ctx* ctx, const sbn_bases* b, const sbn_addrs* addrs, const sbn_fr* rx, size_t nx,
                                      const sbn_fr* ry, size_t ny, size_t row0, size_t nrows, sbn_g1a* C_out, uint8_t* inf_out,
                                      sbn_poly** poly_out) {
    return derefs_commit_rows(ctx, b, addrs, rx, nx, ry, ny, row0, nrows, false, C_out, inf_out, poly_out);
}

extern "C" size_t sbn_poly_len(const sbn_poly* p) { return p ? p->len : 0; }

extern "C" int sbn_poly_download(sbn_ctx* ctx, const sbn_poly* p, sbn_fr* out) {
    if (!ctx || !p || !out || p->ctx != ctx) return SBN_ERR_ARG;
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_ENTER(ctx);
    SBN_TRY(download(ctx, out, p->Z, p->len * sizeof(Fr)));
    SBN_CUDA(ctx, cudaStreamSynchronize(ctx->compute));
    return SBN_OK;
}

// ------------------------------------------------------------------------------------------------
// f1 (cont.): hash layer + product circuits of one side of the memory-checking network, built on the device
// ------------------------------------------------------------------------------------------------
static sbn_prodcircuit* prodcircuit_alloc(sbn_ctx* ctx, size_t len) {
    sbn_prodcircuit* pc = new (std::nothrow) sbn_prodcircuit();
    if (!pc) return nullptr;
    pc->ctx = ctx;
    pc->len = len;
    size_t total = 0;
    for (size_t n = len; n >= 2; n >>= 1) { pc->off.push_back(total); total += n; pc->num_layers++; }
    if (pool_alloc(ctx, &pc->buf, total * sizeof(Fr)) != cudaSuccess) { delete pc; return nullptr; }
    return pc;
}

static void prodcircuit_build_upper(sbn_ctx* ctx, sbn_prodcircuit* pc, cudaStream_t s) {
    for (int l = 0; l + 1 < pc->num_layers; l++) {
        const size_t half = (pc->len >> l) / 2;
        k_product_layer<<<(unsigned)((half + 127) / 128), 128, 0, s>>>(pc->buf + pc->off[l], half, pc->buf + pc->off[l + 1]);
        ctx->launches++;
    }
}

extern "C" int sbn_hashlayer_build(sbn_ctx* ctx, const sbn_addrs* a, int side, const sbn_fr* r, size_t nr, const sbn_fr* r_hash,
                                   const sbn_fr* r_multiset_check, sbn_prodcircuit** circuits_out) {
    if (!ctx || !a || !r || !r_hash || !r_multiset_check || !circuits_out || a->ctx != ctx || (side != 0 && side != 1)) return SBN_ERR_ARG;
    if (!a->read_ts[side] || !a->audit_ts[side]) return SBN_ERR_ARG;                 // sbn_addrs_set_timestamps first
    if (nr == 0 || nr > 30 || (size_t(1) << nr) != a->num_cells || a->N < 2 || (a->N & (a->N - 1)) || a->num_cells < 2) return SBN_ERR_SHAPE;
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_ENTER(ctx);
    cudaStream_t s = ctx->compute;
    const size_t M = a->num_cells, N = a->N, B = a->batch, ncirc = 2 + 2 * B;
    for (size_t i = 0; i < ncirc; i++) circuits_out[i] = nullptr;
    auto fail = [&](int code) {
        for (size_t i = 0; i < ncirc; i++)
            if (circuits_out[i]) {
                pool_free(ctx, circuits_out[i]->buf, (2 * circuits_out[i]->len - 2) * sizeof(Fr));
                delete circuits_out[i];
                circuits_out[i] = nullptr;
            }
        return code;
    };
    for (size_t i = 0; i < ncirc; i++) {
        const bool mem_sized = i == 0 || i == ncirc - 1;
        circuits_out[i] = prodcircuit_alloc(ctx, mem_sized ? M : N);
        if (!circuits_out[i]) { ctx->last_error = "sbn_hashlayer_build: cudaMalloc failed"; return fail(SBN_ERR_OOM); }
    }
    int rc;
    if ((rc = ensure(ctx, ctx->scratch0, 2 * M * sizeof(Fr))) != SBN_OK || (rc = ensure(ctx, ctx->scratch2, nr * sizeof(Fr))) != SBN_OK)
        return fail(rc);
    Fr* rdev = (Fr*)ctx->scratch2.p;
    if (cudaMemcpyAsync(rdev, r, nr * sizeof(Fr), cudaMemcpyHostToDevice, s) != cudaSuccess) return fail(SBN_ERR_CUDA);
    ctx->h2d += nr * sizeof(Fr);
    const Fr* mem = eq_evals_device(ctx, rdev, nr, (Fr*)ctx->scratch0.p, (Fr*)ctx->scratch0.p + M, s);      // eval_table
    HashParams hp;
    memcpy(&hp.rh, r_hash, sizeof(Fr));
    memcpy(&hp.r_ms, r_multiset_check, sizeof(Fr));
    hp.rh2 = fp_mul(hp.rh, hp.rh);
    Fr R2;
    for (int i = 0; i < 8; i++) R2.l[i] = FrParams::R2(i);
    hp.rh2_R2 = fp_mul(hp.rh2, R2);
    k_hash_mem<<<(unsigned)((M + 127) / 128), 128, 0, s>>>(mem, a->audit_ts[side], M, hp, circuits_out[0]->buf, circuits_out[ncirc - 1]->buf);
    const uint32_t* addr = side == 0 ? a->row : a->col;
    for (size_t k = 0; k < B; k++)
        k_hash_ops<<<(unsigned)((N + 127) / 128), 128, 0, s>>>(mem, addr + k * N, a->read_ts[side] + k * N, N, hp,
                                                               circuits_out[1 + k]->buf, circuits_out[1 + B + k]->buf);
    ctx->launches += 1 + B;
    for (size_t i = 0; i < ncirc; i++) prodcircuit_build_upper(ctx, circuits_out[i], s);
    if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess) {
        ctx->last_error = "sbn_hashlayer_build: kernel failure";
        return fail(SBN_ERR_CUDA);
    }
    return SBN_OK;
}

extern "C" int sbn_prodcircuit_download_layer(sbn_prodcircuit* pc, size_t layer, sbn_fr* out) {
    if (!pc || !out) return SBN_ERR_ARG;
    if (layer >= (size_t)pc->num_layers) return SBN_ERR_SHAPE;
    sbn_ctx* ctx = pc->ctx;
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_ENTER(ctx);
    SBN_TRY(download(ctx, out, pc->buf + pc->off[layer], (pc->len >> layer) * sizeof(Fr)));
    SBN_CUDA(ctx, cudaStreamSynchronize(ctx->compute));
    return SBN_OK;
}

// DensePolynomial::evaluate (hyrax.rs:217-222) of a resident polynomial, or of a 2^nr-entry segment of it (the hash layer
// evaluates the individual polynomials merged into derefs / comb_ops / comb_mem, sparse_mlpoly_full.rs:907-976):
// <Z[offset ..], eq(r)> with the eq table built in HBM.
extern "C" int sbn_poly_evaluate(sbn_ctx* ctx, const sbn_poly* poly, size_t offset, const sbn_fr* r, size_t nr, sbn_fr* out) {
    if (!ctx || !poly || !r || !out || poly->ctx != ctx) return SBN_ERR_ARG;
    if (nr == 0 || nr > 30) return SBN_ERR_SHAPE;
    const size_t n = size_t(1) << nr;
    if (offset > poly->len || n > poly->len - offset) return SBN_ERR_SHAPE;
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_ENTER(ctx);
    cudaStream_t s = ctx->compute;
    const unsigned blocks = (unsigned)std::min<size_t>(592, (n + kDotThreads - 1) / kDotThreads);
    SBN_TRY(ensure(ctx, ctx->scratch0, 2 * n * sizeof(Fr)));
    SBN_TRY(ensure(ctx, ctx->scratch1, (blocks + 1) * sizeof(Fr)));
    SBN_TRY(ensure(ctx, ctx->scratch2, nr * sizeof(Fr)));
    SBN_CUDA(ctx, cudaMemcpyAsync(ctx->scratch2.p, r, nr * sizeof(Fr), cudaMemcpyHostToDevice, s));
    ctx->h2d += nr * sizeof(Fr);
    const Fr* eq = eq_evals_device(ctx, (const Fr*)ctx->scratch2.p, nr, (Fr*)ctx->scratch0.p, (Fr*)ctx->scratch0.p + n, s);
    Fr* partial = (Fr*)ctx->scratch1.p;
    k_fr_dot<<<dim3(blocks, 1), kDotThreads, 0, s>>>(poly->Z + offset, 0, eq, 0, (int)n, partial);
    k_fr_sum<<<1, kDotThreads, 0, s>>>(partial, (int)blocks, partial + blocks, 1);
    ctx->launches += 2;
    SBN_CUDA(ctx, cudaGetLastError());
    SBN_TRY(download(ctx, out, partial + blocks, sizeof(Fr)));
    SBN_CUDA(ctx, cudaStreamSynchronize(s));
    return SBN_OK;
}

// `count` evaluations at the same point of the equally long segments starting at offset0 + i * stride: the eq table is
// built once and the dot products run as one launch (HashLayerProof::prove evaluates the 2 b segments of derefs, the 5 b
// of comb_ops and the 2 of comb_mem at one point each, sparse_mlpoly_full.rs:935-976)
extern "C" int sbn_poly_evaluate_strided(sbn_ctx* ctx, const sbn_poly* poly, size_t offset0, size_t stride, size_t count,
                                         const sbn_fr* r, size_t nr, sbn_fr* out) {
    if (!ctx || !poly || !r || !out || poly->ctx != ctx) return SBN_ERR_ARG;
    if (nr == 0 || nr > 30 || count == 0 || count > 4096) return SBN_ERR_SHAPE;
    const size_t n = size_t(1) << nr;
    if (offset0 > poly->len || n > poly->len - offset0 || (count > 1 && (stride == 0 || (count - 1) > (poly->len - offset0 - n) / stride)))
        return SBN_ERR_SHAPE;
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_ENTER(ctx);
    cudaStream_t s = ctx->compute;
    const unsigned blocks = (unsigned)std::min<size_t>(std::max<size_t>(1, 1184 / count), (n + kDotThreads - 1) / kDotThreads);
    SBN_TRY(ensure(ctx, ctx->scratch0, 2 * n * sizeof(Fr)));
    SBN_TRY(ensure(ctx, ctx->scratch1, ((size_t)blocks * count + count) * sizeof(Fr)));
    SBN_TRY(ensure(ctx, ctx->scratch2, nr * sizeof(Fr)));
    SBN_CUDA(ctx, cudaMemcpyAsync(ctx->scratch2.p, r, nr * sizeof(Fr), cudaMemcpyHostToDevice, s));
    ctx->h2d += nr * sizeof(Fr);
    const Fr* eq = eq_evals_device(ctx, (const Fr*)ctx->scratch2.p, nr, (Fr*)ctx->scratch0.p, (Fr*)ctx->scratch0.p + n, s);
    Fr* partial = (Fr*)ctx->scratch1.p;
    Fr* res = partial + (size_t)blocks * count;
    k_fr_dot<<<dim3(blocks, (unsigned)count), kDotThreads, 0, s>>>(poly->Z + offset0, (long)stride, eq, 0, (int)n, partial);
    k_fr_sum<<<(unsigned)count, kDotThreads, 0, s>>>(partial, (int)blocks, res, 1);
    ctx->launches += 2;
    SBN_CUDA(ctx, cudaGetLastError());
    SBN_TRY(download(ctx, out, res, count * sizeof(Fr)));
    SBN_CUDA(ctx, cudaStreamSynchronize(s));
    return SBN_OK;
}

// ------------------------------------------------------------------------------------------------
// f4 building blocks: the dense representation of a multi-sparse-matrix commitment, resident
// ------------------------------------------------------------------------------------------------
// comb_ops = merge(row.ops_addr, row.read_ts, col.ops_addr, col.read_ts, val) and comb_mem = row.audit_ts ++ col.audit_ts
// (sparse_mlpoly_full.rs:155-170), built in HBM from the resident addresses / timestamps and the host `val` (batch x N).
extern "C" int sbn_spark_comb_polys(sbn_ctx* ctx, const sbn_addrs* a, const sbn_fr* val, sbn_poly** comb_ops, sbn_poly** comb_mem) {
    if (!ctx || !a || !val || !comb_ops || !comb_mem || a->ctx != ctx) return SBN_ERR_ARG;
    if (!a->read_ts[0] || !a->audit_ts[0]) return SBN_ERR_ARG;
    *comb_ops = *comb_mem = nullptr;
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_ENTER(ctx);
    cudaStream_t s = ctx->compute;
    const size_t seg = a->batch * a->N, used = 5 * seg;
    size_t len = 1;
    while (len < used) len <<= 1;
    sbn_poly* po = new (std::nothrow) sbn_poly();
    sbn_poly* pm = new (std::nothrow) sbn_poly();
    if (!po || !pm) { delete po; delete pm; return SBN_ERR_OOM; }
    po->ctx = pm->ctx = ctx;
    po->len = len;
    pm->len = 2 * a->num_cells;
    auto fail = [&](int code) {
        pool_free(ctx, po->Z, po->len * sizeof(Fr));
        pool_free(ctx, pm->Z, pm->len * sizeof(Fr));
        delete po; delete pm;
        return code;
    };
    if (pool_alloc(ctx, &po->Z, po->len * sizeof(Fr)) != cudaSuccess || pool_alloc(ctx, &pm->Z, pm->len * sizeof(Fr)) != cudaSuccess) {
        ctx->last_error = "sbn_spark_comb_polys: cudaMalloc failed";
        return fail(SBN_ERR_OOM);
    }
    const uint32_t* src[4] = {a->row, a->read_ts[0], a->col, a->read_ts[1]};
    for (int k = 0; k < 4; k++) k_u32_to_fr<<<(unsigned)((seg + 255) / 256), 256, 0, s>>>(src[k], seg, po->Z + k * seg);
    for (int k = 0; k < 2; k++)
        k_u32_to_fr<<<(unsigned)((a->num_cells + 255) / 256), 256, 0, s>>>(a->audit_ts[k], a->num_cells, pm->Z + k * a->num_cells);
    ctx->launches += 6;
    if (cudaMemcpyAsync(po->Z + 4 * seg, val, seg * sizeof(Fr), cudaMemcpyHostToDevice, s) != cudaSuccess ||
        cudaMemsetAsync(po->Z + used, 0, (len - used) * sizeof(Fr), s) != cudaSuccess ||
        cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess) {
        ctx->last_error = "sbn_spark_comb_polys failed";
        return fail(SBN_ERR_CUDA);
    }
    ctx->h2d += seg * sizeof(Fr);
    *comb_ops = po;
    *comb_mem = pm;
    return SBN_OK;
}

// sum_i A[offA + i] B[offB + i] C[offC + i], i < n, over resident polynomials (DotProductCircuit::evaluate,
// product_tree.rs:81-86; SparseMatPolynomial::evaluate_with_tables, sparse_mlpoly_full.rs:103-108)
extern "C" int sbn_poly_triple_dot(sbn_ctx* ctx, const sbn_poly* A, size_t offA, const sbn_poly* B, size_t offB, const sbn_poly* Cp,
                                   size_t offC, size_t n, sbn_fr* out) {
    if (!ctx || !A || !B || !Cp || !out || A->ctx != ctx || B->ctx != ctx || Cp->ctx != ctx) return SBN_ERR_ARG;
    if (n == 0 || offA > A->len || n > A->len - offA || offB > B->len || n > B->len - offB || offC > Cp->len || n > Cp->len - offC)
        return SBN_ERR_SHAPE;
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_ENTER(ctx);
    cudaStream_t s = ctx->compute;
    const unsigned blocks = (unsigned)std::min<size_t>(592, (n + kDotThreads - 1) / kDotThreads);
    SBN_TRY(ensure(ctx, ctx->scratch1, (blocks + 1) * sizeof(Fr)));
    Fr* partial = (Fr*)ctx->scratch1.p;
    k_fr_triple_dot<<<blocks, kDotThreads, 0, s>>>(A->Z + offA, B->Z + offB, Cp->Z + offC, n, partial);
    k_fr_sum<<<1, kDotThreads, 0, s>>>(partial, (int)blocks, partial + blocks, 1);
    ctx->launches += 2;
    SBN_CUDA(ctx, cudaGetLastError());
    SBN_TRY(download(ctx, out, partial + blocks, sizeof(Fr)));
    SBN_CUDA(ctx, cudaStreamSynchronize(s));
    return SBN_OK;
}

// SparseMatPolynomial::multi_evaluate (sparse_mlpoly_full.rs:110-118) for the batch behind `addrs`, with the values read from
// the val segment of the resident comb_ops: out[s] = sum_i val_s[i] * eq(rx)[row_s[i]] * eq(ry)[col_s[i]].
extern "C" int sbn_spark_evaluate(sbn_ctx* ctx, const sbn_addrs* a, const sbn_poly* comb_ops, const sbn_fr* rx, size_t nx,
                                  const sbn_fr* ry, size_t ny, sbn_fr* out) {
    if (!ctx || !a || !comb_ops || !rx || !ry || !out || a->ctx != ctx || comb_ops->ctx != ctx) return SBN_ERR_ARG;
    if (nx == 0 || ny == 0 || nx > 30 || ny > 30) return SBN_ERR_SHAPE;
    if (a->max_row >= (size_t(1) << nx) || a->max_col >= (size_t(1) << ny) || comb_ops->len < 5 * a->batch * a->N) return SBN_ERR_SHAPE;
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_ENTER(ctx);
    cudaStream_t s = ctx->compute;
    const size_t tx = size_t(1) << nx, ty = size_t(1) << ny, N = a->N, B = a->batch;
    const unsigned blocks = (unsigned)std::min<size_t>(592, (N + kDotThreads - 1) / kDotThreads);
    SBN_TRY(ensure(ctx, ctx->scratch0, 2 * tx * sizeof(Fr)));
    SBN_TRY(ensure(ctx, ctx->scratch1, 2 * ty * sizeof(Fr)));
    SBN_TRY(ensure(ctx, ctx->scratch2, (nx + ny + (blocks + 1) * B) * sizeof(Fr)));
    Fr* rdev = (Fr*)ctx->scratch2.p;
    Fr* partial = rdev + nx + ny;
    SBN_CUDA(ctx, cudaMemcpyAsync(rdev, rx, nx * sizeof(Fr), cudaMemcpyHostToDevice, s));
    SBN_CUDA(ctx, cudaMemcpyAsync(rdev + nx, ry, ny * sizeof(Fr), cudaMemcpyHostToDevice, s));
    ctx->h2d += (nx + ny) * sizeof(Fr);
    const Fr* mem_rx = eq_evals_device(ctx, rdev, nx, (Fr*)ctx->scratch0.p, (Fr*)ctx->scratch0.p + tx, s);
    const Fr* mem_ry = eq_evals_device(ctx, rdev + nx, ny, (Fr*)ctx->scratch1.p, (Fr*)ctx->scratch1.p + ty, s);
    for (size_t k = 0; k < B; k++)
        k_sparse_eval<<<blocks, kDotThreads, 0, s>>>(comb_ops->Z + (4 * B + k) * N, a->row + k * N, a->col + k * N, mem_rx, mem_ry, N,
                                                     partial + k * blocks);
    k_fr_sum<<<(unsigned)B, kDotThreads, 0, s>>>(partial, (int)blocks, partial + B * blocks, 1);
    ctx->launches += B + 1;
    SBN_CUDA(ctx, cudaGetLastError());
    SBN_TRY(download(ctx, out, partial + B * blocks, B * sizeof(Fr)));
    SBN_CUDA(ctx, cudaStreamSynchronize(s));
    return SBN_OK;
}

// ------------------------------------------------------------------------------------------------
// R1CS-sat helpers: resident sparse matrices (compressed rows) and eq tables
// ------------------------------------------------------------------------------------------------
struct sbn_spmat {
    sbn_ctx* ctx = nullptr;
    size_t n = 0, nnz = 0, ncols = 0;
    uint32_t *ptr = nullptr, *idx = nullptr;
    Fr* val = nullptr;
    uint32_t* heavy = nullptr;      // rows with more than kSpmvHeavy entries
    size_t nheavy = 0;
};

extern "C" int sbn_spmat_upload(sbn_ctx* ctx, const uint32_t* ptr, const uint32_t* idx, const sbn_fr* val, size_t n, size_t nnz,
                                size_t ncols, sbn_spmat** out) {
    if (!ctx || !ptr || !out || (nnz && (!idx || !val))) return SBN_ERR_ARG;
    *out = nullptr;
    if (n == 0 || n > (size_t(1) << 30) || nnz > (size_t(1) << 31) || ptr[0] != 0 || ptr[n] != nnz) return SBN_ERR_SHAPE;
    for (size_t i = 0; i < n; i++) if (ptr[i] > ptr[i + 1]) return SBN_ERR_SHAPE;
    for (size_t k = 0; k < nnz; k++) if (idx[k] >= ncols) return SBN_ERR_SHAPE;
    std::vector<uint32_t> heavy_rows;
    for (size_t i = 0; i < n; i++) if (ptr[i + 1] - ptr[i] > kSpmvHeavy) heavy_rows.push_back((uint32_t)i);
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_ENTER(ctx);
    sbn_spmat* m = new (std::nothrow) sbn_spmat();
    if (!m) return SBN_ERR_OOM;
    m->ctx = ctx; m->n = n; m->nnz = nnz; m->ncols = ncols;
    cudaStream_t s = ctx->compute;
    bool ok = dev_malloc(ctx, &m->ptr, (n + 1) * sizeof(uint32_t)) == cudaSuccess &&
              dev_malloc(ctx, &m->idx, std::max<size_t>(1, nnz) * sizeof(uint32_t)) == cudaSuccess &&
              dev_malloc(ctx, &m->val, std::max<size_t>(1, nnz) * sizeof(Fr)) == cudaSuccess &&
              cudaMemcpyAsync(m->ptr, ptr, (n + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, s) == cudaSuccess &&
              (nnz == 0 || (cudaMemcpyAsync(m->idx, idx, nnz * sizeof(uint32_t), cudaMemcpyHostToDevice, s) == cudaSuccess &&
                            cudaMemcpyAsync(m->val, val, nnz * sizeof(Fr), cudaMemcpyHostToDevice, s) == cudaSuccess)) &&
              cudaStreamSynchronize(s) == cudaSuccess;
    if (ok && !heavy_rows.empty()) {
        m->nheavy = heavy_rows.size();
        ok = dev_malloc(ctx, &m->heavy, m->nheavy * sizeof(uint32_t)) == cudaSuccess &&
             cudaMemcpy(m->heavy, heavy_rows.data(), m->nheavy * sizeof(uint32_t), cudaMemcpyHostToDevice) == cudaSuccess &&
             ensure(ctx, ctx->spmv_part, m->nheavy * (size_t)kSpmvSplit * sizeof(Fr)) == SBN_OK;      // so that spmv_device never allocates
    }
    if (!ok) {
        if (m->heavy) cudaFree(m->heavy);
        if (m->ptr) cudaFree(m->ptr);
        if (m->idx) cudaFree(m->idx);
        if (m->val) cudaFree(m->val);
        delete m;
        ctx->last_error = "sbn_spmat_upload failed";
        return SBN_ERR_CUDA;
    }
    ctx->h2d += (n + 1 + nnz) * sizeof(uint32_t) + nnz * sizeof(Fr);
    *out = m;
    return SBN_OK;
}

extern "C" int sbn_spmat_destroy(sbn_spmat* m) {
    if (!m) return SBN_ERR_ARG;
    {
        std::lock_guard<std::mutex> g(m->ctx->mu);
        cudaSetDevice(m->ctx->device);
        cudaStreamSynchronize(m->ctx->compute);
        cudaFree(m->ptr); cudaFree(m->idx); cudaFree(m->val);
        if (m->heavy) cudaFree(m->heavy);
    }
    delete m;
    return SBN_OK;
}

// out[i] = sum_m coeffs[m] * (M_m vec)[i]   (coeffs == NULL: plain sum); 1 <= nm <= 3 matrices of equal shape
// out = M vec (nm = 1, coeffs NULL) or sum_m coeffs[m] M_m vec, everything in HBM
static void spmv_device(sbn_ctx* ctx, const sbn_spmat* const* mats, size_t nm, const sbn_fr* coeffs, const Fr* dvec, Fr* dout,
                        cudaStream_t s) {
    const size_t n = mats[0]->n;
    SpMat sm[3];
    Fr c[3];
    for (size_t m = 0; m < 3; m++) {
        const sbn_spmat* src = mats[m < nm ? m : 0];
        sm[m] = SpMat{src->ptr, src->idx, src->val};
        if (coeffs && m < nm) memcpy(&c[m], &coeffs[m], sizeof(Fr)); else c[m] = Fr::zero();
    }
    k_spmv<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(sm[0], sm[1], sm[2], c[0], c[1], c[2], (int)nm, coeffs ? 1 : 0, dvec, n, dout);
    ctx->launches++;
    for (size_t m = 0; m < nm; m++)
        if (mats[m]->nheavy) {
            k_spmv_heavy_partial<<<dim3(kSpmvSplit, (unsigned)mats[m]->nheavy), kDotThreads, 0, s>>>(sm[m], mats[m]->heavy, dvec,
                                                                                                      (Fr*)ctx->spmv_part.p);
            k_spmv_heavy_final<<<(unsigned)mats[m]->nheavy, kDotThreads, 0, s>>>(mats[m]->heavy, (const Fr*)ctx->spmv_part.p, c[m],
                                                                                 coeffs ? 1 : 0, dout);
            ctx->launches += 2;
        }
}

extern "C" int sbn_spmat_mulvec(sbn_ctx* ctx, const sbn_spmat* const* mats, const sbn_fr* coeffs, size_t nm, const sbn_fr* vec,
                                size_t veclen, sbn_fr* out) {
    if (!ctx || !mats || !vec || !out || nm < 1 || nm > 3) return SBN_ERR_ARG;
    for (size_t m = 0; m < nm; m++) {
        if (!mats[m] || mats[m]->ctx != ctx) return SBN_ERR_ARG;
        if (mats[m]->n != mats[0]->n || mats[m]->ncols > veclen) return SBN_ERR_SHAPE;
    }
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_ENTER(ctx);
    cudaStream_t s = ctx->compute;
    const size_t n = mats[0]->n;
    SBN_TRY(upload(ctx, ctx->scratch0, vec, veclen * sizeof(Fr)));
    SBN_TRY(ensure(ctx, ctx->scratch1, n * sizeof(Fr)));
    spmv_device(ctx, mats, nm, coeffs, (const Fr*)ctx->scratch0.p, (Fr*)ctx->scratch1.p, s);
    SBN_CUDA(ctx, cudaGetLastError());
    SBN_TRY(download(ctx, out, ctx->scratch1.p, n * sizeof(Fr)));
    SBN_CUDA(ctx, cudaStreamSynchronize(s));
    return SBN_OK;
}

// The table set-ups of the two R1CS-sat sumchecks with nothing but z, tau / rx and three coefficients crossing the bus:
//   phase 1 (r1csproof.rs:268-290)  eq(tau), A z, B z, C z   -> the four tables of the cubic sumcheck
//   phase 2 (r1csproof.rs:378-410)  z and r_A A^T eq(rx) + r_B B^T eq(rx) + r_C C^T eq(rx)   -> the two tables of the quadratic one
// (the host-table variants copy 4 x 32 MB up after copying 3 x 32 MB down at 2^20 constraints).
static sbn_sumcheck* sumcheck_alloc(sbn_ctx* ctx, int ntables, size_t len) {
    sbn_sumcheck* st = new (std::nothrow) sbn_sumcheck();
    if (!st) return nullptr;
    st->ctx = ctx;
    st->len = len;
    st->blocks = 592;
    st->ntables = ntables;
    bool ok = pool_alloc(ctx, &st->partial, 3 * st->blocks * sizeof(Fr)) == cudaSuccess && pool_alloc(ctx, &st->out, 4 * sizeof(Fr)) == cudaSuccess;
    for (int k = 0; k < ntables && ok; k++) ok = pool_alloc(ctx, &st->T[k], len * sizeof(Fr)) == cudaSuccess;
    if (!ok) { sumcheck_free(st); ctx->last_error = "sbn_sumcheck_begin: cudaMalloc failed"; return nullptr; }
    return st;
}

static int sumcheck_begin_r1cs_impl(sbn_ctx* ctx, const sbn_spmat* const* mats, const sbn_fr* z, const sbn_poly* vars,
                                    const sbn_fr* tail, size_t n_tail, size_t zlen, const sbn_fr* tau, size_t n_tau, sbn_sumcheck** out);
extern "C" int sbn_sumcheck_begin_r1cs(sbn_ctx* ctx, const sbn_spmat* const* mats, const sbn_fr* z, size_t zlen, const sbn_fr* tau,
                                       size_t n_tau, sbn_sumcheck** out) {
    if (!z) return SBN_ERR_ARG;
    return sumcheck_begin_r1cs_impl(ctx, mats, z, nullptr, nullptr, 0, zlen, tau, n_tau, out);
}
// The same with z = (vars, tail, 0, ..., 0) assembled ON THE DEVICE from the witness polynomial that sbn_poly_upload already
// holds (r1csproof.rs:255-265 builds z = [vars, 1, inputs] on the host): the 2^21 x 32 B upload from pageable memory and the
// host-side assembly of a keyless-scale proof disappear (5-7 ms); `tail` = the n_tail scalars that follow the variables
// (the constant 1 and the public inputs).
extern "C" int sbn_sumcheck_begin_r1cs_resident(sbn_ctx* ctx, const sbn_spmat* const* mats, const sbn_poly* vars, const sbn_fr* tail,
                                                size_t n_tail, size_t zlen, const sbn_fr* tau, size_t n_tau, sbn_sumcheck** out) {
    if (!vars || vars->ctx != ctx || (n_tail && !tail)) return SBN_ERR_ARG;
    if (vars->len + n_tail > zlen) return SBN_ERR_SHAPE;
    return sumcheck_begin_r1cs_impl(ctx, mats, nullptr, vars, tail, n_tail, zlen, tau, n_tau, out);
}
static int sumcheck_begin_r1cs_impl(sbn_ctx* ctx, const sbn_spmat* const* mats, const sbn_fr* z, const sbn_poly* vars,
                                    const sbn_fr* tail, size_t n_tail, size_t zlen, const sbn_fr* tau, size_t n_tau, sbn_sumcheck** out) {
    if (!ctx || !mats || !tau || !out) return SBN_ERR_ARG;
    *out = nullptr;
    if (n_tau == 0 || n_tau > 28) return SBN_ERR_SHAPE;
    const size_t len = size_t(1) << n_tau;
    for (int m = 0; m < 3; m++) {
        if (!mats[m] || mats[m]->ctx != ctx) return SBN_ERR_ARG;
        if (mats[m]->n != len || mats[m]->ncols > zlen) return SBN_ERR_SHAPE;
    }
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_ENTER(ctx);
    cudaStream_t s = ctx->compute;
    SBN_TRY(ensure(ctx, ctx->zkeep, zlen * sizeof(Fr)));
    ctx->zkeep_len = 0;
    SBN_TRY(ensure(ctx, ctx->scratch2, n_tau * sizeof(Fr)));
    sbn_sumcheck* st = sumcheck_alloc(ctx, 4, len);
    if (!st) return SBN_ERR_OOM;
    auto fail = [&](cudaError_t e) { ctx->last_error = std::string("sbn_sumcheck_begin_r1cs: ") + cudaGetErrorString(e); sumcheck_free(st); return SBN_ERR_CUDA; };
    cudaError_t e;
    if (z) {
        if ((e = cudaMemcpyAsync(ctx->zkeep.p, z, zlen * sizeof(Fr), cudaMemcpyHostToDevice, s)) != cudaSuccess) return fail(e);
        ctx->h2d += zlen * sizeof(Fr);
    } else {
        Fr* dz = (Fr*)ctx->zkeep.p;
        if ((e = cudaMemcpyAsync(dz, vars->Z, vars->len * sizeof(Fr), cudaMemcpyDeviceToDevice, s)) != cudaSuccess) return fail(e);
        if (n_tail && (e = cudaMemcpyAsync(dz + vars->len, tail, n_tail * sizeof(Fr), cudaMemcpyHostToDevice, s)) != cudaSuccess) return fail(e);
        if ((e = cudaMemsetAsync(dz + vars->len + n_tail, 0, (zlen - vars->len - n_tail) * sizeof(Fr), s)) != cudaSuccess) return fail(e);
        ctx->h2d += n_tail * sizeof(Fr);
    }
    if ((e = cudaMemcpyAsync(ctx->scratch2.p, tau, n_tau * sizeof(Fr), cudaMemcpyHostToDevice, s)) != cudaSuccess) return fail(e);
    ctx->h2d += n_tau * sizeof(Fr);
    eq_evals_device(ctx, (const Fr*)ctx->scratch2.p, n_tau, st->T[0], st->T[1], s);      // T[1] is scratch until A z lands in it
    for (int m = 0; m < 3; m++) spmv_device(ctx, mats + m, 1, nullptr, (const Fr*)ctx->zkeep.p, st->T[1 + m], s);
    if ((e = cudaGetLastError()) != cudaSuccess || (e = cudaStreamSynchronize(s)) != cudaSuccess) return fail(e);
    ctx->zkeep_len = zlen;             // phase 2 of the same proof takes z from here (z = NULL)
    *out = st;
    return SBN_OK;
}

extern "C" int sbn_sumcheck_begin_quad_r1cs(sbn_ctx* ctx, const sbn_spmat* const* mats_t, const sbn_fr* coeffs, const sbn_fr* rx,
                                            size_t n_rx, const sbn_fr* z, size_t zlen, sbn_sumcheck** out) {
    if (!ctx || !mats_t || !coeffs || !rx || !out) return SBN_ERR_ARG;
    *out = nullptr;
    if (!z && ctx->zkeep_len != zlen) return SBN_ERR_ARG;      // z = NULL: the z of the preceding sbn_sumcheck_begin_r1cs
    if (n_rx == 0 || n_rx > 28 || zlen < 2 || (zlen & (zlen - 1)) || zlen > (1u << 28)) return SBN_ERR_SHAPE;
    const size_t veclen = size_t(1) << n_rx;
    for (int m = 0; m < 3; m++) {
        if (!mats_t[m] || mats_t[m]->ctx != ctx) return SBN_ERR_ARG;
        if (mats_t[m]->n != zlen || mats_t[m]->ncols > veclen) return SBN_ERR_SHAPE;
    }
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_ENTER(ctx);
    cudaStream_t s = ctx->compute;
    SBN_TRY(ensure(ctx, ctx->scratch0, 2 * veclen * sizeof(Fr)));
    SBN_TRY(ensure(ctx, ctx->scratch2, n_rx * sizeof(Fr)));
    sbn_sumcheck* st = sumcheck_alloc(ctx, 2, zlen);
    if (!st) return SBN_ERR_OOM;
    auto fail = [&](cudaError_t e) { ctx->last_error = std::string("sbn_sumcheck_begin_quad_r1cs: ") + cudaGetErrorString(e); sumcheck_free(st); return SBN_ERR_CUDA; };
    cudaError_t e;
    if ((e = cudaMemcpyAsync(st->T[0], z ? (const void*)z : (const void*)ctx->zkeep.p, zlen * sizeof(Fr),
                             z ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, s)) != cudaSuccess) return fail(e);
    if ((e = cudaMemcpyAsync(ctx->scratch2.p, rx, n_rx * sizeof(Fr), cudaMemcpyHostToDevice, s)) != cudaSuccess) return fail(e);
    ctx->h2d += ((z ? zlen : 0) + n_rx) * sizeof(Fr);
    const Fr* eq = eq_evals_device(ctx, (const Fr*)ctx->scratch2.p, n_rx, (Fr*)ctx->scratch0.p, (Fr*)ctx->scratch0.p + veclen, s);
    spmv_device(ctx, mats_t, 3, coeffs, eq, st->T[1], s);
    if ((e = cudaGetLastError()) != cudaSuccess || (e = cudaStreamSynchronize(s)) != cudaSuccess) return fail(e);
    *out = st;
    return SBN_OK;
}

// EqPolynomial::evals (hyrax.rs:355-369) computed in HBM and copied back: out holds 2^n scalars
extern "C" int sbn_eq_evals(sbn_ctx* ctx, const sbn_fr* r, size_t n, sbn_fr* out) {
    if (!ctx || !out || (n && !r)) return SBN_ERR_ARG;
    if (n > 28) return SBN_ERR_SHAPE;
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_ENTER(ctx);
    cudaStream_t s = ctx->compute;
    const size_t len = size_t(1) << n;
    SBN_TRY(ensure(ctx, ctx->scratch0, 2 * len * sizeof(Fr)));
    SBN_TRY(ensure(ctx, ctx->scratch2, std::max<size_t>(1, n) * sizeof(Fr)));
    if (n) SBN_CUDA(ctx, cudaMemcpyAsync(ctx->scratch2.p, r, n * sizeof(Fr), cudaMemcpyHostToDevice, s));
    const Fr* eq = eq_evals_device(ctx, (const Fr*)ctx->scratch2.p, n, (Fr*)ctx->scratch0.p, (Fr*)ctx->scratch0.p + len, s);
    SBN_CUDA(ctx, cudaGetLastError());
    SBN_TRY(download(ctx, out, eq, len * sizeof(Fr)));
    SBN_CUDA(ctx, cudaStreamSynchronize(s));
    return SBN_OK;
}
