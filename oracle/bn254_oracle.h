/* TEST INFRASTRUCTURE ONLY -- CPU oracle for the Hyrax-commit / opening hot path of
 * Antiparadox/Spartan-BN254.  Plain C restatement; see bn254_oracle.c for the per-function
 * reference citations.  PARITY STATUS: unpinned by reference fixtures (the reference has no
 * golden vectors and cannot be built here: no Rust toolchain, arkworks 0.5 not vendored);
 * pinned mathematically (unique affine result) + cross-checked against oracle/pymodel.py and
 * public BN254 constants.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library. */
#ifndef BN254_ORACLE_H
#define BN254_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint64_t l[4]; } ofp;                 /* Montgomery form, LE limbs (ark-ff layout) */
typedef struct { ofp x, y; } og1a;                      /* affine, Montgomery; identity = (0,0)+inf  */
typedef struct { ofp X, Y, Z; } og1j;                   /* Jacobian; Z==0 is the identity            */

/* field helpers (mod selects 0 = Fq, 1 = Fr) */
void orc_fp_from_u64x4(int mod, const uint64_t canon[4], ofp* out);  /* canonical -> Montgomery */
void orc_fp_to_u64x4(int mod, const ofp* in, uint64_t canon[4]);     /* Montgomery -> canonical */
void orc_fp_mul(int mod, const ofp* a, const ofp* b, ofp* out);
void orc_fp_add(int mod, const ofp* a, const ofp* b, ofp* out);
void orc_fp_sub(int mod, const ofp* a, const ofp* b, ofp* out);
int  orc_fp_inv(int mod, const ofp* a, ofp* out);                    /* 0 if a==0 */

/* group */
void orc_g1_generator(og1a* out);
int  orc_g1_on_curve(const og1a* p, uint8_t inf);
void orc_g1_add_affine(const og1a* a, uint8_t ainf, const og1a* b, uint8_t binf, og1a* out, uint8_t* oinf);
void orc_g1_scalar_mul(const og1a* p, uint8_t inf, const ofp* s_mont, og1a* out, uint8_t* oinf);
void orc_g1_compress(const og1a* p, uint8_t inf, uint8_t out[32]);

/* MSM (group.rs:171-175).  algo 0 = naive double-and-add sum, 1 = signed-window Pippenger
 * (restating the published arkworks VariableBaseMSM algorithm). */
void orc_msm(const og1a* pts, const uint8_t* inf /*may be NULL*/, const ofp* scalars_mont, size_t n,
             int algo, og1a* out, uint8_t* oinf);

/* generators (commitments.rs:31-62 + group.rs:110-132) -- n+1 points: G[0..n), h = index n.
 * kinds (may be NULL): 0 primary, 1 fallback, 2 one. */
void orc_gen_scalars(const uint8_t* label, size_t label_len, size_t n, ofp* scalars_mont, uint8_t* kinds);
void orc_multi_commit_gens(const uint8_t* label, size_t label_len, size_t n, og1a* out /* n+1 */);

/* Hyrax (hyrax.rs:253-281): L_size row commitments; blinds NULL = zeros.  threads<=0 -> all cores. */
void orc_hyrax_commit(const og1a* G, const og1a* h, const ofp* Z, size_t L_size, size_t R_size,
                      const ofp* blinds, int threads, og1a* C_out, uint8_t* inf_out);
/* hyrax.rs:311-324 */
void orc_bound(const ofp* Z, const ofp* L, size_t L_size, size_t R_size, int threads, ofp* LZ_out);
/* hyrax.rs:355-369 */
void orc_eq_evals(const ofp* r, size_t ell, ofp* out /* 2^ell */);

/* nizk/bullet.rs:24-126 with caller-supplied challenges u[lg n] (transcript lives on the host).
 * Outputs: L_out/R_out lg n points each, Gamma, a_hat, b_hat, g_hat, blind_hat. */
void orc_bullet_prove(const og1a* Q, const og1a* G, size_t n, const og1a* H, const ofp* a, const ofp* b,
                      const ofp* blind, const ofp* blinds_L, const ofp* blinds_R, const ofp* u,
                      og1a* L_out, uint8_t* L_inf, og1a* R_out, uint8_t* R_inf,
                      og1a* Gamma, uint8_t* Gamma_inf, ofp* a_hat, ofp* b_hat,
                      og1a* g_hat, uint8_t* g_hat_inf, ofp* blind_hat);

/* sumcheck.rs:501-530 cubic round evaluation + hyrax.rs:195-203 bind (a16) */
void orc_sumcheck_cubic_eval(const ofp* A, const ofp* B, const ofp* C, const ofp* D, size_t len,
                             ofp* e0, ofp* e2, ofp* e3);
void orc_sumcheck_quad_eval(const ofp* Z, const ofp* ABC, size_t len, ofp* e0, ofp* e2);   /* sumcheck.rs:690-699 */
void orc_bind_top(ofp* Z, size_t len, const ofp* r);   /* in place: first len/2 entries valid */

/* hashes (third-party sha3 0.10 in the reference: FIPS-202) */
void orc_sha3_256(const uint8_t* in, size_t len, uint8_t out[32]);
void orc_shake256(const uint8_t* in, size_t len, uint8_t* out, size_t outlen);

/* Merlin transcript (third-party merlin 3.0; STROBE-128/Keccak-f[1600]) */
typedef struct { uint8_t st[200]; uint8_t pos, pos_begin, cur_flags; } orc_transcript;
void orc_transcript_new(orc_transcript* t, const uint8_t* label, size_t len);
void orc_transcript_append(orc_transcript* t, const uint8_t* label, size_t llen, const uint8_t* msg, size_t mlen);
void orc_transcript_challenge(orc_transcript* t, const uint8_t* label, size_t llen, uint8_t* out, size_t outlen);
void orc_transcript_challenge_scalar(orc_transcript* t, const uint8_t* label, size_t llen, ofp* out_mont);


/* Hyrax opening proof (hyrax.rs:65-151 PolyEvalProof, nizk/mod.rs:418-567 DotProductProofLog,
 * nizk/bullet.rs BulletReductionProof) with Merlin transcripts.  The random tape is a transcript the caller
 * seeds (the reference seeds it from OsRng, random.rs:15-23). */
#define ORC_MAX_LG 32
typedef struct {
    size_t lg_n;
    og1a L[ORC_MAX_LG], R[ORC_MAX_LG];
    uint8_t L_inf[ORC_MAX_LG], R_inf[ORC_MAX_LG];
    og1a delta, beta;
    uint8_t delta_inf, beta_inf;
    ofp z1, z2;
} orc_eval_proof;
/* gens: G[n] + h (gens_n) and G1 = gens_1.G[0]; n = 2^(ell - ell/2).  blinds may be NULL (zeros), blind_Zr may be NULL. */
void orc_poly_eval_prove(const ofp* Z, size_t ell, const ofp* blinds, const ofp* r, const ofp* Zr, const ofp* blind_Zr,
                         const og1a* G, const og1a* h, const og1a* G1, orc_transcript* transcript, orc_transcript* tape,
                         orc_eval_proof* proof, og1a* C_Zr_prime, uint8_t* C_Zr_prime_inf);
/* returns 1 when the proof verifies (hyrax.rs:118-137) */
int orc_poly_eval_verify(const orc_eval_proof* proof, size_t ell, const ofp* r, const og1a* C_Zr, uint8_t C_Zr_inf,
                         const og1a* comm, const uint8_t* comm_inf, const og1a* G, const og1a* h, const og1a* G1,
                         orc_transcript* transcript);

/* CPU baseline of the end-to-end prove: the table-sized phases of SNARK::prove (snark.rs:428-484) for a synthetic R1CS of
 * the keyless shape with 2^s constraints, on `threads` host threads (0 = all); seconds[] has orc_prove_workload_phases()
 * entries.  A restatement of the WORK (same algorithms, operation counts and round-to-round dependencies), not a prover.
 * derefs_rows_done > 0 commits only that many derefs rows and scales that phase up (returns 1 then, 0 otherwise, < 0 on a
 * bad argument). */
int orc_prove_workload(int s, int threads, size_t derefs_rows_done, double* seconds, double* gens_seconds);
int orc_prove_workload_phases(void);
const char* orc_prove_workload_phase_name(int i);

#ifdef __cplusplus
}
#endif
#endif
