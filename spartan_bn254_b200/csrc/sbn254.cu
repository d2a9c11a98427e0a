// libsbn254 C ABI (include/sbn254.h): contexts, resident generator sets, the batched Hyrax commit
// pipeline and its stream plumbing.  Kernels live in msm_kernels.cuh / opening_kernels.cuh.
#include "../../include/sbn254.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "msm_kernels.cuh"
#include "opening_kernels.cuh"

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

struct sbn_ctx {
    int device = 0;
    cudaStream_t compute = nullptr, copy = nullptr;
    std::mutex mu;
    std::string last_error;
    long chunk_rows = 512;
    long window_bits = 0;
    long task_cap = 0;      // 0 = auto: 2.5 x the mean bucket occupancy
    long reduce_m = 16;     // buckets per reduction thread
    uint64_t launches = 0, h2d = 0, d2h = 0;
    // grow-only workspaces
    struct Slot {          // one in-flight chunk of rows: private workspace + stream
        cudaStream_t stream = nullptr;
        cudaEvent_t done = nullptr;
        DevBuf entries, tstart, tasks, partials;
    } slots[2];
    cudaEvent_t fork = nullptr;
    DevBuf totals, dZ, dblinds, dC, dinf, scratch0, scratch1, scratch2;
    // last-commit profile
    std::vector<cudaEvent_t> ev_pool;
    float prof_ms[4] = {0, 0, 0, 0};
    int prof_launches[4] = {0, 0, 0, 0};
};

struct sbn_bases {
    sbn_ctx* ctx = nullptr;
    size_t n = 0;      // generators without h
    int n1 = 0;        // n + 1
    int c = 0, W = 0, nb = 0;
    Affine* table = nullptr;   // W * n1 affine points
};

#define SBN_CUDA(ctx, call)                                                                      \
    do {                                                                                         \
        cudaError_t _e = (call);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            (ctx)->last_error = std::string(#call) + ": " + cudaGetErrorString(_e);              \
            return _e == cudaErrorMemoryAllocation ? SBN_ERR_OOM : SBN_ERR_CUDA;                 \
        }                                                                                        \
    } while (0)

#define SBN_TRY(expr)            \
    do {                         \
        int _s = (expr);         \
        if (_s != SBN_OK) return _s; \
    } while (0)

static int ensure(sbn_ctx* ctx, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap) return SBN_OK;
    if (b.p) {
        SBN_CUDA(ctx, cudaDeviceSynchronize());
        SBN_CUDA(ctx, cudaFree(b.p));
        b.p = nullptr;
        b.cap = 0;
    }
    size_t want = bytes + bytes / 8;
    SBN_CUDA(ctx, cudaMalloc(&b.p, want));
    b.cap = want;
    return SBN_OK;
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
    for (auto& sl : ctx->slots)
        ok = ok && cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
        delete ctx;
        return SBN_ERR_CUDA;
    }
    *out = ctx;
    return SBN_OK;
}

extern "C" int sbn_ctx_destroy(sbn_ctx* ctx) {
    if (!ctx) return SBN_ERR_ARG;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->compute);
    cudaStreamSynchronize(ctx->copy);
    for (DevBuf* b : {&ctx->totals, &ctx->dZ, &ctx->dblinds, &ctx->dC, &ctx->dinf, &ctx->scratch0, &ctx->scratch1,
                      &ctx->scratch2})
        release(*b);
    for (auto& sl : ctx->slots) {
        cudaStreamSynchronize(sl.stream);
        for (DevBuf* b : {&sl.entries, &sl.tstart, &sl.tasks, &sl.partials}) release(*b);
        cudaStreamDestroy(sl.stream);
        cudaEventDestroy(sl.done);
    }
    cudaEventDestroy(ctx->fork);
    for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
    cudaStreamDestroy(ctx->compute);
    cudaStreamDestroy(ctx->copy);
    delete ctx;
    return SBN_OK;
}

extern "C" int sbn_ctx_synchronize(sbn_ctx* ctx) {
    if (!ctx) return SBN_ERR_ARG;
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_CUDA(ctx, cudaSetDevice(ctx->device));
    SBN_CUDA(ctx, cudaStreamSynchronize(ctx->copy));
    SBN_CUDA(ctx, cudaStreamSynchronize(ctx->compute));
    return SBN_OK;
}

extern "C" int sbn_ctx_set(sbn_ctx* ctx, const char* key, long value) {
    if (!ctx || !key) return SBN_ERR_ARG;
    std::lock_guard<std::mutex> g(ctx->mu);
    if (!strcmp(key, "chunk_rows")) {
        if (value < 1) return SBN_ERR_ARG;
        ctx->chunk_rows = value;
    } else if (!strcmp(key, "task_cap")) {
        if (value < 0 || value > kMaxTaskCap) return SBN_ERR_ARG;
        ctx->task_cap = value;
    } else if (!strcmp(key, "reduce_m")) {
        if (value < 1 || (value & (value - 1))) return SBN_ERR_ARG;
        ctx->reduce_m = value;
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
static int task_cap_for(const sbn_ctx* ctx, const sbn_bases* b);

// ------------------------------------------------------------------------------------------------
// bases
// ------------------------------------------------------------------------------------------------
extern "C" int sbn_bases_create(sbn_ctx* ctx, const sbn_g1a* G, const uint8_t* G_inf, size_t n, const sbn_g1a* h,
                                sbn_bases** out) {
    if (!ctx || !G || !h || !out) return SBN_ERR_ARG;
    *out = nullptr;
    if (n == 0 || n > (1u << 24)) return SBN_ERR_SHAPE;
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_CUDA(ctx, cudaSetDevice(ctx->device));
    sbn_bases* b = new (std::nothrow) sbn_bases();
    if (!b) return SBN_ERR_OOM;
    b->ctx = ctx;
    b->n = n;
    b->n1 = (int)n + 1;
    b->c = ctx->window_bits ? (int)ctx->window_bits : choose_window(n + 1);
    b->W = msm_num_windows(b->c);
    b->nb = 1 << (b->c - 1);
    if ((uint64_t)b->W * b->n1 >= (1ull << 31)) { delete b; return SBN_ERR_SHAPE; }

    Affine* dbases = nullptr;
    uint8_t* dinf = nullptr;
    auto fail = [&](int code) {
        if (dbases) cudaFree(dbases);
        if (dinf) cudaFree(dinf);
        if (b->table) cudaFree(b->table);
        delete b;
        return code;
    };
    cudaError_t e;
    if ((e = cudaMalloc(&dbases, sizeof(Affine) * b->n1)) != cudaSuccess ||
        (e = cudaMalloc(&dinf, b->n1)) != cudaSuccess ||
        (e = cudaMalloc(&b->table, sizeof(Affine) * (size_t)b->W * b->n1)) != cudaSuccess) {
        ctx->last_error = std::string("sbn_bases_create cudaMalloc: ") + cudaGetErrorString(e);
        return fail(SBN_ERR_OOM);
    }
    std::vector<uint8_t> inf_host(b->n1, 0);
    if (G_inf) memcpy(inf_host.data(), G_inf, n);
    if ((e = cudaMemcpyAsync(dbases, G, sizeof(Affine) * n, cudaMemcpyHostToDevice, ctx->compute)) != cudaSuccess ||
        (e = cudaMemcpyAsync(dbases + n, h, sizeof(Affine), cudaMemcpyHostToDevice, ctx->compute)) != cudaSuccess ||
        (e = cudaMemcpyAsync(dinf, inf_host.data(), b->n1, cudaMemcpyHostToDevice, ctx->compute)) != cudaSuccess) {
        ctx->last_error = std::string("sbn_bases_create upload: ") + cudaGetErrorString(e);
        return fail(SBN_ERR_CUDA);
    }
    ctx->h2d += sizeof(Affine) * b->n1 + b->n1;
    k_build_tables<<<(b->n1 + 63) / 64, 64, 0, ctx->compute>>>(dbases, dinf, b->n1, b->c, b->W, b->table);
    ctx->launches++;
    if ((e = cudaGetLastError()) != cudaSuccess || (e = cudaStreamSynchronize(ctx->compute)) != cudaSuccess) {
        ctx->last_error = std::string("k_build_tables: ") + cudaGetErrorString(e);
        return fail(SBN_ERR_CUDA);
    }
    cudaFree(dbases);
    cudaFree(dinf);
    *out = b;
    return SBN_OK;
}

extern "C" int sbn_bases_destroy(sbn_bases* b) {
    if (!b) return SBN_ERR_ARG;
    {
        std::lock_guard<std::mutex> g(b->ctx->mu);
        cudaSetDevice(b->ctx->device);
        cudaStreamSynchronize(b->ctx->compute);
        if (b->table) cudaFree(b->table);
    }
    delete b;
    return SBN_OK;
}
extern "C" size_t sbn_bases_len(const sbn_bases* b) { return b ? b->n : 0; }
extern "C" int sbn_bases_window_bits(const sbn_bases* b) { return b ? b->c : 0; }

// ------------------------------------------------------------------------------------------------
// commit pipeline
// ------------------------------------------------------------------------------------------------
static int task_cap_for(const sbn_ctx* ctx, const sbn_bases* b) {
    if (ctx->task_cap) return (int)ctx->task_cap;
    double mean = double(b->W) * b->n1 / b->nb;
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
static void launch_sort(const Fr* Z, const Fr* blinds, int R, int cap, uint32_t E, uint32_t max_tasks, uint32_t* entries,
                        uint32_t* tstart, Task* tasks, int rows, cudaStream_t s) {
    k_sort_row<C><<<rows, kSortThreads, 0, s>>>(Z, blinds, R, 1, cap, E, max_tasks, entries, tstart, tasks);
}

static int dispatch_sort(int c, const Fr* Z, const Fr* blinds, int R, int cap, uint32_t E, uint32_t max_tasks,
                         uint32_t* entries, uint32_t* tstart, Task* tasks, int rows, cudaStream_t s) {
    switch (c) {
#define SBN_CASE(CC) case CC: launch_sort<CC>(Z, blinds, R, cap, E, max_tasks, entries, tstart, tasks, rows, s); return SBN_OK;
        SBN_CASE(4) SBN_CASE(5) SBN_CASE(6) SBN_CASE(7) SBN_CASE(8) SBN_CASE(9) SBN_CASE(10) SBN_CASE(11)
        SBN_CASE(12) SBN_CASE(13)
#undef SBN_CASE
        default: return SBN_ERR_UNSUPPORTED;
    }
}

// Runs the first three stages for `rows` rows in pipeline slot `sl`; dZ_chunk points at the chunk's first row.
static int commit_chunk(sbn_ctx* ctx, const sbn_bases* b, sbn_ctx::Slot& sl, const Fr* dZ_chunk, const Fr* dblinds_chunk,
                        int rows, int R, XYZZ* totals_chunk, size_t& ev_idx, std::vector<int>& ev_stage) {
    const uint32_t E = (uint32_t)b->W * (uint32_t)b->n1;
    const int cap = task_cap_for(ctx, b);
    const uint32_t max_tasks = (uint32_t)msm_max_tasks(E, b->nb, cap);
    uint32_t* entries = (uint32_t*)sl.entries.p;
    uint32_t* tstart = (uint32_t*)sl.tstart.p;
    Task* tasks = (Task*)sl.tasks.p;
    XYZZ* partials = (XYZZ*)sl.partials.p;
    cudaStream_t stream = sl.stream;
    auto mark = [&](int stage) {
        cudaEvent_t e = get_event(ctx, ev_idx++);
        if (e) cudaEventRecord(e, stream);
        ev_stage.push_back(stage);
    };
    mark(-1);
    SBN_TRY(dispatch_sort(b->c, dZ_chunk, dblinds_chunk, R, cap, E, max_tasks, entries, tstart, tasks, rows, stream));
    mark(0);
    const size_t threads = (size_t)rows * max_tasks;
    k_accumulate<<<(unsigned)((threads + kAccThreads - 1) / kAccThreads), kAccThreads, 0, stream>>>(
        b->table, entries, tstart, tasks, partials, rows, b->nb, E, max_tasks);
    mark(1);
    int m = std::min((int)ctx->reduce_m, b->nb);
    int tpr = std::min(kRedThreads, b->nb / m);
    int rows_per_block = kRedThreads / tpr;
    k_reduce<<<(rows + rows_per_block - 1) / rows_per_block, kRedThreads, kRedThreads * sizeof(XYZZ), stream>>>(
        partials, tstart, rows, b->nb, tpr, max_tasks, totals_chunk);
    mark(2);
    ctx->launches += 3;
    SBN_CUDA(ctx, cudaGetLastError());
    return SBN_OK;
}

static int ensure_commit_workspace(sbn_ctx* ctx, const sbn_bases* b, size_t chunk, size_t L) {
    const size_t E = (size_t)b->W * b->n1;
    const size_t max_tasks = msm_max_tasks(E, b->nb, task_cap_for(ctx, b));
    if (max_tasks >= (1u << 24)) return SBN_ERR_SHAPE;
    const size_t nslots = L > chunk ? 2 : 1;
    for (size_t i = 0; i < nslots; i++) {
        auto& sl = ctx->slots[i];
        SBN_TRY(ensure(ctx, sl.entries, chunk * E * sizeof(uint32_t)));
        SBN_TRY(ensure(ctx, sl.tstart, chunk * (b->nb + 1) * sizeof(uint32_t)));
        SBN_TRY(ensure(ctx, sl.tasks, chunk * max_tasks * sizeof(Task)));
        SBN_TRY(ensure(ctx, sl.partials, chunk * max_tasks * sizeof(XYZZ)));
    }
    SBN_TRY(ensure(ctx, ctx->totals, L * sizeof(XYZZ)));
    return SBN_OK;
}

static void collect_profile(sbn_ctx* ctx, const std::vector<int>& ev_stage) {
    for (int i = 0; i < 4; i++) { ctx->prof_ms[i] = 0; ctx->prof_launches[i] = 0; }
    for (size_t i = 1; i < ev_stage.size(); i++) {
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

// The pipeline shared by both entry points.  Chunks of rows alternate between two slots (streams with
// private workspaces) so one chunk's latency-bound reduction overlaps the next chunk's accumulation; when
// `host_Z` is given each chunk's H2D copy is issued on the copy stream and handed over by an event.
// `main` is the stream the caller's inputs are ordered on and on which the normalisation runs.
static int run_commit(sbn_ctx* ctx, const sbn_bases* b, const Fr* dZ, const Fr* host_Z, size_t L, size_t R,
                      const Fr* dblinds, Affine* dC, uint8_t* dinf, cudaStream_t main, std::vector<int>& ev_stage) {
    const size_t chunk = std::min<size_t>(L, (size_t)ctx->chunk_rows);
    const size_t nchunks = (L + chunk - 1) / chunk;
    XYZZ* totals = (XYZZ*)ctx->totals.p;
    size_t ev_idx = 0;
    const size_t handoff_base = 4 * nchunks + 8;    // copy->compute events live after the profiling events
    SBN_CUDA(ctx, cudaEventRecord(ctx->fork, main));
    const size_t nslots = nchunks > 1 ? 2 : 1;
    for (size_t i = 0; i < nslots; i++) SBN_CUDA(ctx, cudaStreamWaitEvent(ctx->slots[i].stream, ctx->fork, 0));
    if (host_Z) SBN_CUDA(ctx, cudaStreamWaitEvent(ctx->copy, ctx->fork, 0));
    for (size_t ci = 0, row0 = 0; row0 < L; row0 += chunk, ci++) {
        auto& sl = ctx->slots[ci % nslots];
        int rows = (int)std::min(chunk, L - row0);
        if (host_Z) {
            SBN_CUDA(ctx, cudaMemcpyAsync((void*)(dZ + row0 * R), host_Z + row0 * R, (size_t)rows * R * sizeof(Fr),
                                          cudaMemcpyHostToDevice, ctx->copy));
            ctx->h2d += (size_t)rows * R * sizeof(Fr);
            cudaEvent_t copied = get_event(ctx, handoff_base + ci);
            if (!copied) { ctx->last_error = "cudaEventCreate failed"; return SBN_ERR_CUDA; }
            SBN_CUDA(ctx, cudaEventRecord(copied, ctx->copy));
            SBN_CUDA(ctx, cudaStreamWaitEvent(sl.stream, copied, 0));
        }
        SBN_TRY(commit_chunk(ctx, b, sl, dZ + row0 * R, dblinds ? dblinds + row0 : nullptr, rows, (int)R, totals + row0,
                             ev_idx, ev_stage));
    }
    for (size_t i = 0; i < nslots; i++) {
        SBN_CUDA(ctx, cudaEventRecord(ctx->slots[i].done, ctx->slots[i].stream));
        SBN_CUDA(ctx, cudaStreamWaitEvent(main, ctx->slots[i].done, 0));
    }
    {
        cudaEvent_t e = get_event(ctx, ev_idx++);
        if (e) cudaEventRecord(e, main);
        ev_stage.push_back(-1);
    }
    k_normalize<<<(unsigned)((L + 63) / 64), 64, 0, main>>>(totals, (int)L, dC, dinf);
    ctx->launches++;
    {
        cudaEvent_t e = get_event(ctx, ev_idx++);
        if (e) cudaEventRecord(e, main);
        ev_stage.push_back(3);
    }
    SBN_CUDA(ctx, cudaGetLastError());
    return SBN_OK;
}

extern "C" int sbn_hyrax_commit_device(sbn_ctx* ctx, const sbn_bases* b, const void* dZ, size_t L, size_t R,
                                       const void* dblinds, void* dC_out, void* dinf_out, void* stream_) {
    if (!ctx || !b || !dZ || !dC_out || b->ctx != ctx) return SBN_ERR_ARG;
    SBN_TRY(check_commit_shape(b, L, R));
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t stream = stream_ ? (cudaStream_t)stream_ : ctx->compute;
    const size_t chunk = std::min<size_t>(L, (size_t)ctx->chunk_rows);
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

extern "C" int sbn_hyrax_commit(sbn_ctx* ctx, const sbn_bases* b, const sbn_fr* Z, size_t L, size_t R, const sbn_fr* blinds,
                                sbn_g1a* C_out, uint8_t* inf_out) {
    if (!ctx || !b || !Z || !C_out || !inf_out || b->ctx != ctx) return SBN_ERR_ARG;
    SBN_TRY(check_commit_shape(b, L, R));
    std::lock_guard<std::mutex> g(ctx->mu);
    SBN_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t chunk = std::min<size_t>(L, (size_t)ctx->chunk_rows);
    SBN_TRY(ensure_commit_workspace(ctx, b, chunk, L));
    SBN_TRY(ensure(ctx, ctx->dZ, L * R * sizeof(Fr)));
    SBN_TRY(ensure(ctx, ctx->dC, L * sizeof(Affine)));
    SBN_TRY(ensure(ctx, ctx->dinf, L));
    Fr* dbl = nullptr;
    if (blinds) {
        SBN_TRY(ensure(ctx, ctx->dblinds, L * sizeof(Fr)));
        dbl = (Fr*)ctx->dblinds.p;
        SBN_CUDA(ctx, cudaMemcpyAsync(dbl, blinds, L * sizeof(Fr), cudaMemcpyHostToDevice, ctx->compute));
        ctx->h2d += L * sizeof(Fr);
    }
    std::vector<int> ev_stage;
    SBN_TRY(run_commit(ctx, b, (const Fr*)ctx->dZ.p, (const Fr*)Z, L, R, dbl, (Affine*)ctx->dC.p, (uint8_t*)ctx->dinf.p,
                       ctx->compute, ev_stage));
    SBN_CUDA(ctx, cudaMemcpyAsync(C_out, ctx->dC.p, L * sizeof(Affine), cudaMemcpyDeviceToHost, ctx->compute));
    SBN_CUDA(ctx, cudaMemcpyAsync(inf_out, ctx->dinf.p, L, cudaMemcpyDeviceToHost, ctx->compute));
    ctx->d2h += L * sizeof(Affine) + L;
    SBN_CUDA(ctx, cudaStreamSynchronize(ctx->compute));
    collect_profile(ctx, ev_stage);
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
    SBN_CUDA(ctx, cudaSetDevice(ctx->device));
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
    SBN_CUDA(ctx, cudaSetDevice(ctx->device));
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
    SBN_CUDA(ctx, cudaSetDevice(ctx->device));
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
    SBN_CUDA(ctx, cudaSetDevice(ctx->device));
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
// not yet implemented entry points
// ------------------------------------------------------------------------------------------------
extern "C" int sbn_msm(sbn_ctx*, const sbn_g1a*, const uint8_t*, const sbn_fr*, size_t, sbn_g1a*, uint8_t*) { return SBN_ERR_UNSUPPORTED; }
extern "C" int sbn_commit(sbn_ctx*, const sbn_bases*, const sbn_fr*, size_t, const sbn_fr*, sbn_g1a*, uint8_t*) { return SBN_ERR_UNSUPPORTED; }
extern "C" int sbn_bound(sbn_ctx*, const sbn_fr*, const sbn_fr*, size_t, size_t, sbn_fr*) { return SBN_ERR_UNSUPPORTED; }
extern "C" int sbn_bullet_begin(sbn_ctx*, const sbn_bases*, const sbn_g1a*, const sbn_fr*, const sbn_fr*, size_t, const sbn_fr*, sbn_g1a*, uint8_t*, sbn_bullet**) { return SBN_ERR_UNSUPPORTED; }
extern "C" int sbn_bullet_round(sbn_bullet*, const sbn_fr*, const sbn_fr*, sbn_g1a*, uint8_t*, sbn_g1a*, uint8_t*) { return SBN_ERR_UNSUPPORTED; }
extern "C" int sbn_bullet_fold(sbn_bullet*, const sbn_fr*, const sbn_fr*) { return SBN_ERR_UNSUPPORTED; }
extern "C" int sbn_bullet_end(sbn_bullet*, sbn_fr*, sbn_fr*, sbn_g1a*, uint8_t*) { return SBN_ERR_UNSUPPORTED; }
extern "C" int sbn_bullet_destroy(sbn_bullet*) { return SBN_ERR_UNSUPPORTED; }
extern "C" int sbn_sumcheck_begin(sbn_ctx*, const sbn_fr*, const sbn_fr*, const sbn_fr*, const sbn_fr*, size_t, sbn_sumcheck**) { return SBN_ERR_UNSUPPORTED; }
extern "C" int sbn_sumcheck_round_eval(sbn_sumcheck*, sbn_fr*, sbn_fr*, sbn_fr*) { return SBN_ERR_UNSUPPORTED; }
extern "C" int sbn_sumcheck_bind(sbn_sumcheck*, const sbn_fr*) { return SBN_ERR_UNSUPPORTED; }
extern "C" int sbn_sumcheck_end(sbn_sumcheck*, sbn_fr*) { return SBN_ERR_UNSUPPORTED; }
extern "C" int sbn_sumcheck_destroy(sbn_sumcheck*) { return SBN_ERR_UNSUPPORTED; }
