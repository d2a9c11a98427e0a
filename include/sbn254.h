/* libsbn254 -- B200-native (sm_100a) backend for the Hyrax commit / opening hot path of
 * Antiparadox/Spartan-BN254.  Flat C ABI: plain pointers and sizes, opaque handles, no unwinding.
 *
 * The reference has no FFI of its own (100% Rust); each entry point below replaces one Rust call
 * site, cited as file:line under the reference tree.  INTEGRATION.md shows the `extern "C"` shim a
 * maintainer adds on the Rust side.
 *
 * Data layout (identical to ark-ff / ark-ec in-memory values, so the shim passes slices as-is):
 *   sbn_fr   Fr element, 4 x u64 little-endian limbs, MONTGOMERY form (reference scalar.rs:15)
 *   sbn_g1a  affine G1 point, x then y, each 4 x u64 LE limbs Montgomery Fq; the identity is
 *            carried out-of-band by an `inf` byte array (1 = identity, coordinates then ignored
 *            on input and written as zero on output)  (reference group.rs:20, commitments.rs:24)
 *
 * Return value: 0 on success, negative sbn_status on error; sbn_strerror() names it.  Preconditions
 * the reference enforces with assert!/panic (hyrax.rs:258,290,295; commitments.rs:134,146;
 * bullet.rs:42-47) come back as SBN_ERR_SHAPE -- the shim turns non-zero into panic!.
 * There is no CPU fallback: without a usable CUDA device every call fails with SBN_ERR_CUDA.
 *
 * Threading: a context serialises its own calls with an internal mutex (reference call sites are
 * single threaded except rayon fan-out inside commit_inner, which this library replaces wholesale).
 */
#ifndef SBN254_H
#define SBN254_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint64_t l[4]; } sbn_fr;
typedef struct { uint64_t x[4], y[4]; } sbn_g1a;

typedef struct sbn_ctx sbn_ctx;        /* one CUDA device + streams + workspace            */
typedef struct sbn_bases sbn_bases;    /* generator set resident in HBM with window tables */
typedef struct sbn_bullet sbn_bullet;  /* device-resident state of one bullet reduction    */
typedef struct sbn_sumcheck sbn_sumcheck;
typedef struct sbn_prodcircuit sbn_prodcircuit;
typedef struct sbn_bsumcheck sbn_bsumcheck;
typedef struct sbn_addrs sbn_addrs;
typedef struct sbn_spmat sbn_spmat;

typedef enum {
    SBN_OK = 0,
    SBN_ERR_ARG = -1,      /* null pointer / bad handle                */
    SBN_ERR_SHAPE = -2,    /* size precondition of the reference broken */
    SBN_ERR_CUDA = -3,     /* CUDA runtime error (see sbn_last_cuda_error) */
    SBN_ERR_OOM = -4,
    SBN_ERR_UNSUPPORTED = -5
} sbn_status;

const char* sbn_strerror(int status);
const char* sbn_last_cuda_error(const sbn_ctx* ctx);   /* text of the last CUDA failure, "" if none */
int sbn_version(void);                                  /* 100 * major + minor */

/* ---- context ------------------------------------------------------------------------------- */
int sbn_ctx_create(int device, sbn_ctx** out);
int sbn_ctx_destroy(sbn_ctx* ctx);
int sbn_ctx_synchronize(sbn_ctx* ctx);
/* Tunables: "chunk_rows" (rows per pipeline chunk; 0 = auto), "window_bits" (0 = auto, applies to bases created
 * afterwards), "task_cap" (max entries one accumulation thread sums; fuller buckets are split; 0 = auto), "leaf_m"
 * (buckets per leaf thread of the two-level bucket reduction; 0 = auto), "ba_rounds" / "ba_batch" (batched-affine rounds of
 * the bucket pipeline and pairs per thread; -1 / 0 = auto), "dedup_generators" (merge equal generators of sets created
 * afterwards; default 1).
 * Resident tables of digit multiples (HBM for speed; all results are identical with or without them):
 *   "mult_max_mb"   budget (MiB, default 6144) of the table d * 2^(k c) * G_j that commits of many rows sum over instead of
 *                   sorting into buckets; the widest window that fits is used, 0 disables the path
 *   "mult_min_rows" commits of at least this many rows (default 256) take that path and build the table on first use
 *   "tab_max_mb"    budget (MiB, default 3072) of the 8-bit-window table of an opening's generator set
 *                   (sbn_bases_create_ext): single rows and row pairs become sums of table points; 0 disables
 *   "small_commit_path" 0 sends short generator sets and few-row commits through the general pipeline (test hook)
 *   "small_scalar_path" 1 (default): a commit without blinds classifies its rows by the bit length of their largest scalar
 *                   and commits runs of small rows (comb_ops' addresses and timestamps, sparse_mlpoly_full.rs:176-196) over
 *                   a table of just the windows that cover them; 0 disables the scan
 * Tuning of the tabulated-sum path (mult_kernels.cuh; every setting gives identical results):
 *   "mult_layout"   1 (default) position-major entry lists -- a warp is 32 rows at one table column; 0 row-major (round 1)
 *   "mult_streams"  chunks in flight, 1..4 (default 2);  "mult_rounds" batched-affine rounds, 0 = auto
 *   "ba_minb"       register target of the finish pass: 3 (80 registers, default) or 4 (64) resident blocks per SM
 *   "ba_prefetch"   round 1 reads its entries two pairs ahead and prefetches the table points into L2 (default 0: loses)
 *   "bsc_device"    sbn_bsumcheck_prove: 0 (default) the transcript runs on the host between the kernels of a round; 1 the
 *                   short last rounds of a layer run in ONE block with the Merlin transcript on the device; 2 every round
 *                   (csrc/transcript_kernels.cuh; identical proofs and transcript states; measured no faster -- a one-warp
 *                   Keccak + Montgomery step is ~16 us on the GPU)
 *   "fused_rounds"  sbn_bsumcheck_prove: 1 (default) the bind of a round rides in the next round's evaluation kernel (one launch
 *                   per round on the Fiat-Shamir chain); 0 separate bind launches
 *   "host_normalize" 1 (default): the one to four points of a short commitment / a bullet round come back as XYZZ and are
 *                   normalised on the host (a 30 us one-warp inversion chain on the device, a few us on a host core); 0 on the device
 *   "finish_smem_kb" / "prefix_smem_kb"  dynamic shared memory requested per block to cap co-residency (default 0)
 *   "l2_fetch"      cudaLimitMaxL2FetchGranularity (32 / 64 / 128; measured to change nothing on B200)
 *   "ablate"        PROFILING ONLY: bit mask of skipped launches (1 prefix round 1, 2 prefix rounds >= 2, 4 inversions,
 *                   8 row sums, 16 finish round 1, 32 finish rounds >= 2); results are WRONG when non-zero */
int sbn_ctx_set(sbn_ctx* ctx, const char* key, long value);
/* Counters since creation / last reset: kernels launched by this library, bytes copied H2D / D2H. */
int sbn_ctx_counters(sbn_ctx* ctx, uint64_t* kernel_launches, uint64_t* h2d_bytes, uint64_t* d2h_bytes, int reset);
/* Memory behaviour since creation: out[0] bytes of released device buffers held by the context's pool, out[1] times an
 * allocation failure emptied that pool and retried, out[2] digit-multiple tables ("mult_max_mb") the device could not hold
 * -- those generator sets ran through the bucket pipeline and sbn_last_cuda_error says so --, out[3] buffers in the pool,
 * out[4] commits that took the small-scalar schedule ("small_scalar_path"), out[5..7] reserved (0).
 * (No reference counterpart: the Rust prover's Vec allocations cannot fail softly; a GPU backend's can.) */
int sbn_ctx_memory_stats(sbn_ctx* ctx, uint64_t out[8]);
/* Device time (ms, CUDA events on the library's compute stream) of the kernels of the last
 * sbn_hyrax_commit* call, per stage: [0] digit decomposition + bucket sort, [1] bucket accumulation,
 * [2] bucket reduction, [3] affine normalisation; and the number of launches per stage. */
int sbn_ctx_last_commit_profile(sbn_ctx* ctx, float ms[4], int launches[4]);

/* pinned host memory for callers that want overlapped copies (optional) */
int sbn_host_alloc(void** out, size_t bytes);
int sbn_host_free(void* p);
/* Streams for the asynchronous entry points (sbn_hyrax_commit_async, sbn_hyrax_commit_device), for callers without CUDA bindings
 * of their own (the Rust shim): a non-blocking stream on ctx's device.  sbn_stream_synchronize does not take the context's
 * lock, so one thread can wait for a commit while another issues the next one on another stream. */
int sbn_stream_create(sbn_ctx* ctx, void** stream_out);
int sbn_stream_synchronize(sbn_ctx* ctx, void* stream);
int sbn_stream_destroy(sbn_ctx* ctx, void* stream);

/* ---- generators: MultiCommitGens kept resident (commitments.rs:17-27, G_affine + h_affine) ---- */
/* Uploads n affine generators G and the blinding generator h, and precomputes the fixed-base window
 * tables 2^(k*c) * G_j used by every later commit.  inf may be NULL (no identity among the bases). */
int sbn_bases_create(sbn_ctx* ctx, const sbn_g1a* G, const uint8_t* G_inf, size_t n, const sbn_g1a* h,
                     sbn_bases** out);
/* Same with one extra generator g1 kept resident beside G and h: DotProductProofGens (nizk/mod.rs:404-415) splits
 * MultiCommitGens::new(n + 1) into gens_n = (G[0..n), h) and gens_1 = (G[n], h); passing g1 = G[n] lets the opening
 * (sbn_bullet_begin with q_scalar) run every group operation of the bullet reduction on the window tables. */
int sbn_bases_create_ext(sbn_ctx* ctx, const sbn_g1a* G, const uint8_t* G_inf, size_t n, const sbn_g1a* g1,
                         const sbn_g1a* h, sbn_bases** out);
int sbn_bases_destroy(sbn_bases* bases);
size_t sbn_bases_len(const sbn_bases* bases);                 /* n (without h) */
int sbn_bases_window_bits(const sbn_bases* bases);
/* Window width and size of the digit-multiple table that commits of many rows sum over (0 / 0 when none has been built:
 * the table is built by the first commit of at least `mult_min_rows` rows, within `mult_max_mb` -- sbn_ctx_set). */
int sbn_bases_mult_table(const sbn_bases* b, int* window_bits, uint64_t* bytes);

/* ---- a7/a9: DensePolynomial::commit_inner (hyrax.rs:253-281), R1CSProof::commit_poly
 *      (r1csproof.rs:210-237).  C_i = sum_j Z[i*R_size + j] * G_j + blinds[i] * h  for i < L_size.
 * Requires R_size == sbn_bases_len(bases) (commitments.rs:146 assert).  blinds NULL = zeros
 * (hyrax.rs:301-305).  Outputs are AFFINE (what append_to_transcript needs, hyrax.rs:44-52). */
int sbn_hyrax_commit(sbn_ctx* ctx, const sbn_bases* bases, const sbn_fr* Z, size_t L_size, size_t R_size,
                     const sbn_fr* blinds, sbn_g1a* C_out, uint8_t* inf_out);
/* The same call, asynchronous: the chunked H2D copies of Z, the kernels and the D2H copy of the commitments are ordered on
 * `stream` (a cudaStream_t, not 0) and the call returns at once; the caller synchronises the stream before reading C_out /
 * inf_out.  Z, C_out and inf_out should be pinned (sbn_host_alloc) for the copies to be asynchronous.  The library keeps three
 * sets of staging buffers and two sets of workspaces; consecutive calls take them in turn and every set is handed from one call
 * to the next by a device-side event, so the calls may sit on any streams: commits issued on two or three streams overlap, the
 * copy of one under the kernels of the others (a prover that commits several polynomials -- comb_ops and comb_mem at encode time,
 * sparse_mlpoly_full.rs:183-184 -- issues them back to back).  Generator sets without a digit-multiple table complete before
 * the call returns. */
int sbn_hyrax_commit_async(sbn_ctx* ctx, const sbn_bases* bases, const sbn_fr* Z, size_t L_size, size_t R_size,
                           const sbn_fr* blinds, sbn_g1a* C_out, uint8_t* inf_out, void* stream);
/* Same with every buffer already in device memory of ctx's device (Z, blinds, C_out, inf_out are
 * device pointers); runs asynchronously on `stream` (a cudaStream_t, 0 = the context's stream). */
int sbn_hyrax_commit_device(sbn_ctx* ctx, const sbn_bases* bases, const void* dZ, size_t L_size, size_t R_size,
                            const void* dblinds, void* dC_out, void* dinf_out, void* stream);

/* a7 on k GPUs of one process (sbn_ctx_create(n_gpus) became one context per device plus this call).
 * Replaces the Rayon fan-out of reference hyrax.rs:259-265: the L rows are cut into k contiguous blocks, block i is
 * committed by ctxs[i] over bases[i] (the same generators, created once per context) from its own host thread, and the
 * blocks are written side by side into C_out / inf_out -- no inter-GPU exchange.  Same arguments, results and errors as
 * sbn_hyrax_commit; the first non-zero status of any block is returned. */
int sbn_hyrax_commit_multi(sbn_ctx* const* ctxs, const sbn_bases* const* bases, size_t k, const sbn_fr* Z, size_t L_size,
                           size_t R_size, const sbn_fr* blinds, sbn_g1a* C_out, uint8_t* inf_out);

/* ---- a6: GroupElement::msm_affine / vartime_multiscalar_mul (group.rs:143-175).
 * One variable-base MSM over caller-supplied points.  The reference swallows a length mismatch into
 * the identity (`unwrap_or_default`); here the caller passes one n, so that case cannot arise. */
int sbn_msm(sbn_ctx* ctx, const sbn_g1a* points, const uint8_t* inf, const sbn_fr* scalars, size_t n,
            sbn_g1a* out, uint8_t* inf_out);
/* a5: <[Scalar] as Commitments>::commit (commitments.rs:144-154) against resident bases:
 * sum_j s[j] G_j + blind * h, n == sbn_bases_len. */
int sbn_commit(sbn_ctx* ctx, const sbn_bases* bases, const sbn_fr* scalars, size_t n, const sbn_fr* blind,
               sbn_g1a* out, uint8_t* inf_out);

/* ---- batched fixed-point scalar multiplication: out[i] = s[i] * P  (group.rs:121,130 generator() *
 *      scalar for generator derivation; commitments.rs:64-76 MultiCommitGens::scale) */
int sbn_g1_scalar_mul_batch(sbn_ctx* ctx, const sbn_g1a* P, const sbn_fr* s, size_t n, sbn_g1a* out, uint8_t* inf_out);
/* out[i] = s * P[i] */
int sbn_g1_scale_points(sbn_ctx* ctx, const sbn_g1a* P, const uint8_t* inf, size_t n, const sbn_fr* s,
                        sbn_g1a* out, uint8_t* inf_out);

/* ---- a11: DensePolynomial::bound (hyrax.rs:311-324): LZ[i] = sum_j L[j] * Z[j*R_size + i] */
int sbn_bound(sbn_ctx* ctx, const sbn_fr* Z, const sbn_fr* L, size_t L_size, size_t R_size, sbn_fr* LZ_out);

/* ---- resident polynomial: the evaluation vector of a DensePolynomial (hyrax.rs:155-160) uploaded once and used by
 *      both its commitment (prove/encode time) and its opening (`bound`), so the scalars cross PCIe once. */
typedef struct sbn_poly sbn_poly;
int sbn_poly_upload(sbn_ctx* ctx, const sbn_fr* Z, size_t len, sbn_poly** out);
int sbn_poly_destroy(sbn_poly* poly);
int sbn_poly_commit(sbn_ctx* ctx, const sbn_bases* bases, const sbn_poly* poly, size_t L_size, size_t R_size,
                    const sbn_fr* blinds, sbn_g1a* C_out, uint8_t* inf_out);
/* Rows [first_row, first_row + n_rows) of the same commitment (blinds: one per committed row, NULL = zeros): the block one
 * rank of a row-sharded R1CSProof::commit_poly (r1csproof.rs:210-237) contributes; rows are independent (hyrax.rs:259-265). */
int sbn_poly_commit_rows(sbn_ctx* ctx, const sbn_bases* bases, const sbn_poly* poly, size_t first_row, size_t n_rows,
                         size_t R_size, const sbn_fr* blinds, sbn_g1a* C_out, uint8_t* inf_out);                    /* = sbn_hyrax_commit */
int sbn_poly_bound(sbn_ctx* ctx, const sbn_poly* poly, const sbn_fr* L, size_t L_size, size_t R_size, sbn_fr* LZ_out);

/* ---- a14: BulletReductionProof::prove (nizk/bullet.rs:24-126) with a, b (and the generators) resident on device.
 * The Fiat-Shamir transcript stays on the host: each round returns (L, R), the caller appends them,
 * draws u and hands it back.  `bases` supplies G[0..n) and H = h.  Q is given either
 *   - as q_scalar (Q = q_scalar * g1, bases made by sbn_bases_create_ext) -- the form the reference's only caller
 *     uses (nizk/mod.rs:480-485: Q = gens_1.scale(r).G[0]); every MSM of the reduction then runs on the window
 *     tables and the generators are never folded (their fold coefficients are carried as scalars), or
 *   - as an arbitrary point Q (q_scalar NULL): generators are folded explicitly as in bullet.rs:85-89.
 *   begin : Gamma = MSM(a,G) + <a,b> Q + blind H                                         (bullet.rs:57-59)
 *   round : L = MSM(a_L,G_R) + c_L Q + blind_L H, R = MSM(a_R,G_L) + c_R Q + blind_R H    (bullet.rs:70-76)
 *   fold  : G <- u^-1 G_L + u G_R, a <- u a_L + u^-1 a_R, b <- u^-1 b_L + u b_R           (bullet.rs:85-102)
 *   end   : a_hat, b_hat, g_hat                                                         (bullet.rs:110-112) */
int sbn_bullet_begin(sbn_ctx* ctx, const sbn_bases* bases, const sbn_g1a* Q, const sbn_fr* q_scalar, const sbn_fr* a,
                     const sbn_fr* b, size_t n, const sbn_fr* blind, sbn_g1a* Gamma_out, uint8_t* Gamma_inf, sbn_bullet** out);
int sbn_bullet_round(sbn_bullet* st, const sbn_fr* blind_L, const sbn_fr* blind_R,
                     sbn_g1a* L_out, uint8_t* L_inf, sbn_g1a* R_out, uint8_t* R_inf);
int sbn_bullet_fold(sbn_bullet* st, const sbn_fr* u, const sbn_fr* u_inv);
int sbn_bullet_end(sbn_bullet* st, sbn_fr* a_hat, sbn_fr* b_hat, sbn_g1a* g_hat, uint8_t* g_hat_inf);
/* sbn_bullet_end together with DotProductProofLog::prove's delta = d * g_hat + r_delta * h (nizk/mod.rs:497-500), computed
 * as a second row over the resident tables (g_hat = <coef, G>).  Table path only (begin with q_scalar): SBN_ERR_UNSUPPORTED
 * otherwise. */
int sbn_bullet_end_delta(sbn_bullet* st, const sbn_fr* d, const sbn_fr* r_delta, sbn_fr* a_hat, sbn_fr* b_hat, sbn_g1a* g_hat,
                         uint8_t* g_hat_inf, sbn_g1a* delta, uint8_t* delta_inf);
int sbn_bullet_destroy(sbn_bullet* st);

/* ---- a16: R1CS-sat sumcheck round (sumcheck.rs:501-530 evaluation, :551-554 + hyrax.rs:195-203 bind).
 * Four tables (tau, Az, Bz, Cz) of `len` scalars stay on device across rounds. */
int sbn_sumcheck_begin(sbn_ctx* ctx, const sbn_fr* tau, const sbn_fr* Az, const sbn_fr* Bz, const sbn_fr* Cz,
                       size_t len, sbn_sumcheck** out);
/* phase 2 of the R1CS-sat proof (sumcheck.rs:657-811, loop :690-699): two tables (z, ABC), comb = z * ABC,
 * evaluations at 0 and 2 only (round_eval leaves e3 untouched; it may be NULL). */
int sbn_sumcheck_begin_quad(sbn_ctx* ctx, const sbn_fr* z, const sbn_fr* ABC, size_t len, sbn_sumcheck** out);
int sbn_sumcheck_round_eval(sbn_sumcheck* st, sbn_fr* e0, sbn_fr* e2, sbn_fr* e3);
int sbn_sumcheck_bind(sbn_sumcheck* st, const sbn_fr* r);
int sbn_sumcheck_end(sbn_sumcheck* st, sbn_fr finals[4]);   /* the length-1 tables (unused slots zero) */
int sbn_sumcheck_destroy(sbn_sumcheck* st);

/* ---- f1: the product layer of the Spark argument, pure Fr.
 * Product circuit = ProductCircuit::new (product_tree.rs:39-57): every layer of pairwise products of a 2^k-entry
 * polynomial, resident on device; evaluate = ProductCircuit::evaluate (:59-64). */
int sbn_prodcircuit_create(sbn_ctx* ctx, const sbn_fr* poly, size_t len, sbn_prodcircuit** out);
int sbn_prodcircuit_evaluate(sbn_prodcircuit* pc, sbn_fr* out);
size_t sbn_prodcircuit_num_layers(const sbn_prodcircuit* pc);
int sbn_prodcircuit_destroy(sbn_prodcircuit* pc);
/* Batched cubic sumcheck = the table work of SumcheckInstanceProof::prove_cubic_batched (sumcheck.rs:165-330) as
 * ProductCircuitEvalProofBatched::prove drives it (product_tree.rs:251-392): P "parallel" instances (left_vec[layer],
 * right_vec[layer] of each circuit) share poly_C = EqPolynomial(rand).evals() (built on device, hyrax.rs:355-369);
 * S "sequential" instances (dot-product circuits, :292-305) bring their own three tables of 2^n_rand scalars.
 * The transcript stays on the host: round_eval returns (e0, e2, e3) per instance in instance order (parallel first),
 * the caller combines them with its coefficients, draws r_j and calls bind; end returns the length-1 tables:
 * A_final / B_final per instance, C_final[0] = poly_C_par[0] followed by one entry per sequential instance.
 * The circuits' layer tables are bound IN PLACE, as the reference consumes them. */
int sbn_bsumcheck_begin(sbn_ctx* ctx, sbn_prodcircuit* const* circuits, size_t P, size_t layer_id, const sbn_fr* rand,
                        size_t n_rand, const sbn_fr* const* seqA, const sbn_fr* const* seqB, const sbn_fr* const* seqC,
                        size_t S, sbn_bsumcheck** out);
/* Same, with the 3 * S tables of the sequential instances given as segments of resident polynomials
 * (seq_polys[3k + w], seq_offsets[3k + w]; w = 0 left, 1 right, 2 weight): they are copied device to device. */
int sbn_bsumcheck_begin_resident(sbn_ctx* ctx, sbn_prodcircuit* const* circuits, size_t P, size_t layer_id, const sbn_fr* rand,
                                 size_t n_rand, const sbn_poly* const* seq_polys, const size_t* seq_offsets, size_t S,
                                 sbn_bsumcheck** out);
int sbn_bsumcheck_round_eval(sbn_bsumcheck* st, sbn_fr* evals /* (P + S) x 3 */);
int sbn_bsumcheck_bind(sbn_bsumcheck* st, const sbn_fr* r);
int sbn_bsumcheck_end(sbn_bsumcheck* st, sbn_fr* A_final, sbn_fr* B_final, sbn_fr* C_final);
/* One layer's whole round loop inside the library (SumcheckInstanceProof::prove_cubic_batched, sumcheck.rs:165-330): per
 * round the batched evaluation, the combination with `coeffs` (P + S scalars, :273-275), UniPoly::from_evals
 * (unipoly.rs:28-59), the transcript append (unipoly.rs:119-127), the challenge "challenge_nextround"
 * (transcript.rs:56-67) and the bind; then the final values as sbn_bsumcheck_end returns them.  `merlin` is the transcript
 * state of the sbn_merlin_* calls.  polys: num_rounds x 4 coefficients, lowest degree first; r_out: the challenges. */
int sbn_bsumcheck_prove(sbn_bsumcheck* st, void* merlin, const sbn_fr* claim, const sbn_fr* coeffs, size_t num_rounds,
                        sbn_fr* polys, sbn_fr* r_out, sbn_fr* claim_out, sbn_fr* A_final, sbn_fr* B_final, sbn_fr* C_final);
int sbn_bsumcheck_destroy(sbn_bsumcheck* st);

/* ---- f2: the derefs polynomial built on the device.
 * The address vectors of the Spark commitment are fixed at encode time (sparse_mlpoly_full.rs:204-243); they are
 * uploaded once (batch x N row addresses, batch x N column addresses).  At prove time only rx / ry cross the bus:
 * sbn_derefs_commit builds mem_rx = eq(rx), mem_ry = eq(ry) in HBM (:1713-1718), gathers
 * comb = merge(mem_rx[row_s] ..., mem_ry[col_s] ...) (:245-257, :292-297; zero-padded to a power of two, hyrax.rs:237-247)
 * and commits it with zero blinds (:301-304 -> hyrax.rs:283-308) over `bases` (R_size = 2^(ell - ell/2) generators).
 * C_out / inf_out hold 2^(ell/2) points.  poly_out (may be NULL) receives the resident polynomial for its opening. */
int sbn_addrs_upload(sbn_ctx* ctx, const uint32_t* row_addrs, const uint32_t* col_addrs, size_t batch, size_t N,
                     sbn_addrs** out);
int sbn_addrs_destroy(sbn_addrs* addrs);
int sbn_derefs_commit(sbn_ctx* ctx, const sbn_bases* bases, const sbn_addrs* addrs, const sbn_fr* rx, size_t nx,
                      const sbn_fr* ry, size_t ny, sbn_g1a* C_out, uint8_t* inf_out, sbn_poly** poly_out);
/* Memory-checking timestamps of the same commitment (AddrTimestamps::new, sparse_mlpoly_full.rs:212-243): read_ts per
 * operation (batch x N) and audit_ts per memory cell (num_cells, a power of two) for the row and the column side. */
int sbn_addrs_set_timestamps(sbn_addrs* addrs, const uint32_t* row_read_ts, const uint32_t* row_audit_ts,
                             const uint32_t* col_read_ts, const uint32_t* col_audit_ts, size_t num_cells);
/* Layers::new for one side (0 = row over eq(rx), 1 = column over eq(ry); sparse_mlpoly_full.rs:745-841): hashes
 * h(addr, val, ts) = ts r_hash^2 + val r_hash + addr - r_multiset_check of the init / read / write / audit sets and
 * builds their product circuits, all on the device.  circuits_out receives 2 + 2 * batch handles in the order of
 * ProductLayer: init, read_vec[0..batch), write_vec[0..batch), audit. */
int sbn_hashlayer_build(sbn_ctx* ctx, const sbn_addrs* addrs, int side, const sbn_fr* r, size_t nr, const sbn_fr* r_hash,
                        const sbn_fr* r_multiset_check, sbn_prodcircuit** circuits_out);
int sbn_prodcircuit_download_layer(sbn_prodcircuit* pc, size_t layer, sbn_fr* out /* len >> layer scalars */);
/* DensePolynomial::evaluate (hyrax.rs:217-222) of the 2^nr evaluations poly[offset .. offset + 2^nr) at the point r. */
int sbn_poly_evaluate(sbn_ctx* ctx, const sbn_poly* poly, size_t offset, const sbn_fr* r, size_t nr, sbn_fr* out);
/* The same for `count` equally long segments starting at offset0 + i * stride, at one point: one eq table, one launch
 * (the evaluations of HashLayerProof::prove, sparse_mlpoly_full.rs:935-976).  out: count scalars. */
int sbn_poly_evaluate_strided(sbn_ctx* ctx, const sbn_poly* poly, size_t offset0, size_t stride, size_t count, const sbn_fr* r,
                              size_t nr, sbn_fr* out);
/* comb_ops = merge(row.ops_addr, row.read_ts, col.ops_addr, col.read_ts, val) (zero-padded to a power of two) and
 * comb_mem = row.audit_ts ++ col.audit_ts (sparse_mlpoly_full.rs:155-170) as resident polynomials, built from the
 * resident addresses / timestamps and the host `val` (batch x N Montgomery scalars). */
int sbn_spark_comb_polys(sbn_ctx* ctx, const sbn_addrs* addrs, const sbn_fr* val, sbn_poly** comb_ops, sbn_poly** comb_mem);
/* SparseMatPolynomial::multi_evaluate (sparse_mlpoly_full.rs:110-118) of the batch behind `addrs`, values read from the val
 * segment of the resident comb_ops: out[s] = sum_i val_s[i] * eq(rx)[row_s[i]] * eq(ry)[col_s[i]], s < batch. */
int sbn_spark_evaluate(sbn_ctx* ctx, const sbn_addrs* addrs, const sbn_poly* comb_ops, const sbn_fr* rx, size_t nx,
                       const sbn_fr* ry, size_t ny, sbn_fr* out);
/* sum_{i < n} A[offA + i] * B[offB + i] * C[offC + i] over resident polynomials (DotProductCircuit::evaluate,
 * product_tree.rs:81-86). */
int sbn_poly_triple_dot(sbn_ctx* ctx, const sbn_poly* A, size_t offA, const sbn_poly* B, size_t offB, const sbn_poly* C,
                        size_t offC, size_t n, sbn_fr* out);
/* Multi-GPU form of sbn_derefs_commit: builds the whole polynomial, commits rows [row0, row0 + nrows) of its Hyrax matrix
 * (C_out / inf_out hold nrows points); ranks gather their blocks. */
int sbn_derefs_commit_rows(sbn_ctx* ctx, const sbn_bases* bases, const sbn_addrs* addrs, const sbn_fr* rx, size_t nx,
                           const sbn_fr* ry, size_t ny, size_t row0, size_t nrows, sbn_g1a* C_out, uint8_t* inf_out,
                           sbn_poly** poly_out);
size_t sbn_poly_len(const sbn_poly* poly);
int sbn_poly_download(sbn_ctx* ctx, const sbn_poly* poly, sbn_fr* out);

/* ---- R1CS-sat helpers (r1csproof.rs:285, :380): a sparse matrix in compressed-row form resident on the device, and
 * out[i] = sum_m coeffs[m] * (M_m * vec)[i] for up to three matrices of the same shape (coeffs NULL = plain sum).
 * multiply_vec (sparse_mlpoly.rs:77-87) uses the row-sorted copy with vec = z; compute_eval_table_sparse (:145-160) the
 * column-sorted copy with vec = eq(rx) and coeffs = (r_A, r_B, r_C). */
int sbn_spmat_upload(sbn_ctx* ctx, const uint32_t* ptr /* n + 1 */, const uint32_t* idx, const sbn_fr* val, size_t n,
                     size_t nnz, size_t ncols, sbn_spmat** out);
int sbn_spmat_destroy(sbn_spmat* m);
int sbn_spmat_mulvec(sbn_ctx* ctx, const sbn_spmat* const* mats, const sbn_fr* coeffs, size_t nm, const sbn_fr* vec,
                     size_t veclen, sbn_fr* out);
/* The two sumcheck set-ups of R1CSProof::prove with the tables built in HBM from the resident matrices, so that only z,
 * the challenge vector and three coefficients cross the bus:
 *   begin_r1cs       r1csproof.rs:268-290: tables eq(tau), A z, B z, C z; mats = {A, B, C} by row, 2^n_tau rows each
 *   begin_quad_r1cs  r1csproof.rs:378-410: tables z and r_A A^T e + r_B B^T e + r_C C^T e with e = eq(rx); mats_t = the
 *                    column-sorted copies (zlen rows each), coeffs = (r_A, r_B, r_C)
 * begin_quad_r1cs accepts z = NULL: the z of the preceding begin_r1cs on this context (same zlen) is still resident.
 * The states are driven by sbn_sumcheck_round_eval / _bind / _end / _destroy like the host-table ones. */
int sbn_sumcheck_begin_r1cs(sbn_ctx* ctx, const sbn_spmat* const* mats, const sbn_fr* z, size_t zlen, const sbn_fr* tau,
                            size_t n_tau, sbn_sumcheck** out);
int sbn_sumcheck_begin_quad_r1cs(sbn_ctx* ctx, const sbn_spmat* const* mats_t, const sbn_fr* coeffs, const sbn_fr* rx,
                                 size_t n_rx, const sbn_fr* z, size_t zlen, sbn_sumcheck** out);
/* begin_r1cs with z = (vars, tail, 0 ... 0) assembled on the device from the resident witness polynomial (sbn_poly_upload of the
 * variables, the same handle R1CSProof::commit_poly committed, r1csproof.rs:210-237,255-265): `tail` = the n_tail scalars that
 * follow the variables (the constant 1 and the public inputs), zlen = 2 * num_vars.  Nothing table-sized crosses the bus. */
int sbn_sumcheck_begin_r1cs_resident(sbn_ctx* ctx, const sbn_spmat* const* mats, const sbn_poly* vars, const sbn_fr* tail,
                                     size_t n_tail, size_t zlen, const sbn_fr* tau, size_t n_tau, sbn_sumcheck** out);
/* EqPolynomial::evals (hyrax.rs:355-369): the 2^n evaluations of eq(r, .), computed on the device. */
int sbn_eq_evals(sbn_ctx* ctx, const sbn_fr* r, size_t n, sbn_fr* out);

/* ---- utilities used by tests / harnesses */
/* Keccak-f[1600] on a 25-lane little-endian state, in place (host only): the permutation under the Merlin transcript of
 * the host mirrors (transcript.rs; merlin 3.0 = STROBE-128). */
void sbn_keccak_f1600(uint64_t* state);
/* The Merlin transcript itself (merlin 3.0 = STROBE-128; transcript.rs:37-80 wraps it): `state` is 203 caller-owned bytes.
 * append_many appends `count` messages of `mlen` bytes under the same label (append_scalars, transcript.rs:46-52). */
void sbn_merlin_init(void* state, const uint8_t* label, size_t label_len);
void sbn_merlin_append(void* state, const uint8_t* label, size_t label_len, const uint8_t* msg, size_t msg_len);
void sbn_merlin_append_many(void* state, const uint8_t* label, size_t label_len, const uint8_t* msgs, size_t msg_len, size_t count);
void sbn_merlin_challenge(void* state, const uint8_t* label, size_t label_len, uint8_t* out, size_t n);
/* n consecutive challenge_scalar calls with one label (transcript.rs:56-73, random.rs:24-31), results in Montgomery form */
void sbn_merlin_challenge_scalars(void* state, const uint8_t* label, size_t label_len, size_t n, sbn_fr* out);
/* GroupElement::compress (group.rs:135-140) on the host: n affine Montgomery points -> n x 32 bytes of ark's compressed
 * encoding; and the share loop of PolyCommitment::append_to_transcript (hyrax.rs:46-50): compress + append each point. */
int sbn_g1_compress(const sbn_g1a* pts, const uint8_t* inf, size_t n, uint8_t* out);
int sbn_merlin_append_points(void* state, const uint8_t* label, size_t label_len, const sbn_g1a* pts, const uint8_t* inf, size_t n);
int sbn_fr_from_canonical(sbn_ctx* ctx, const uint64_t* canon /* n x 4 */, size_t n, sbn_fr* out);
int sbn_fr_to_canonical(sbn_ctx* ctx, const sbn_fr* in, size_t n, uint64_t* canon);
/* the same conversions on the host, for the handful of scalars a sumcheck round exchanges */
int sbn_fr_to_canonical_host(const sbn_fr* in, size_t n, uint64_t* canon);
int sbn_fr_from_canonical_host(const uint64_t* canon, size_t n, sbn_fr* out);
/* integer-multiply microbenchmark: returns achieved 32-bit multiply-add results per second for
 * kind 0 = IMAD (mad.lo), 1 = IMAD.HI, 2 = IMAD.WIDE (a 64-bit result counted as 2), 3 = Montgomery Fq
 * multiplications per second (not x264). */
int sbn_microbench(sbn_ctx* ctx, int kind, double* per_second);

#ifdef __cplusplus
}
#endif
#endif
