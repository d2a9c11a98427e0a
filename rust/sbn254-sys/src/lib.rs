//! Thin FFI crate over `include/sbn254.h`.  NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no
//! Rust toolchain and no crates.io access; this is the source a maintainer of Spartan-BN254 drops in.
//!
//! Layout contract: `ark_bn254::Fr` / `Fq` are `Fp256<MontBackend<_, 4>>` whose only field is
//! `BigInt<4>([u64; 4])` in Montgomery form, so a `&[Scalar]` (newtype over `Fr`, reference scalar.rs:15)
//! is passed as `*const SbnFr` without conversion.  `G1Affine { x, y, infinity }` is not `repr(C)`, so
//! generators are marshalled field by field ONCE per generator set (see `Bases::new`).
#![allow(non_camel_case_types)]
use ark_bn254::{Fq, Fr, G1Affine};
use ark_ff::{BigInt, PrimeField};
use std::os::raw::{c_char, c_int, c_long, c_void};

#[repr(C)] #[derive(Clone, Copy, Default)] pub struct SbnFr { pub l: [u64; 4] }
#[repr(C)] #[derive(Clone, Copy, Default)] pub struct SbnG1a { pub x: [u64; 4], pub y: [u64; 4] }
#[repr(C)] pub struct sbn_ctx { _p: [u8; 0] }
#[repr(C)] pub struct sbn_bases { _p: [u8; 0] }
#[repr(C)] pub struct sbn_bullet { _p: [u8; 0] }
#[repr(C)] pub struct sbn_sumcheck { _p: [u8; 0] }
#[repr(C)] pub struct sbn_poly { _p: [u8; 0] }
#[repr(C)] pub struct sbn_prodcircuit { _p: [u8; 0] }
#[repr(C)] pub struct sbn_bsumcheck { _p: [u8; 0] }
#[repr(C)] pub struct sbn_addrs { _p: [u8; 0] }
#[repr(C)] pub struct sbn_spmat { _p: [u8; 0] }

extern "C" {
    pub fn sbn_strerror(status: c_int) -> *const c_char;
    pub fn sbn_ctx_create(device: c_int, out: *mut *mut sbn_ctx) -> c_int;
    pub fn sbn_ctx_destroy(ctx: *mut sbn_ctx) -> c_int;
    pub fn sbn_ctx_set(ctx: *mut sbn_ctx, key: *const c_char, value: c_long) -> c_int;
    pub fn sbn_bases_create(ctx: *mut sbn_ctx, g: *const SbnG1a, g_inf: *const u8, n: usize, h: *const SbnG1a,
                            out: *mut *mut sbn_bases) -> c_int;
    pub fn sbn_bases_destroy(b: *mut sbn_bases) -> c_int;
    pub fn sbn_bases_mult_table(b: *const sbn_bases, window_bits: *mut c_int, bytes: *mut u64) -> c_int;
    pub fn sbn_hyrax_commit(ctx: *mut sbn_ctx, b: *const sbn_bases, z: *const SbnFr, l_size: usize, r_size: usize,
                            blinds: *const SbnFr, c_out: *mut SbnG1a, inf_out: *mut u8) -> c_int;
    pub fn sbn_hyrax_commit_async(ctx: *mut sbn_ctx, b: *const sbn_bases, z: *const SbnFr, l_size: usize, r_size: usize,
                                  blinds: *const SbnFr, c_out: *mut SbnG1a, inf_out: *mut u8, stream: *mut c_void) -> c_int;
    pub fn sbn_hyrax_commit_multi(ctxs: *const *mut sbn_ctx, bases: *const *const sbn_bases, k: usize, z: *const SbnFr,
                                  l_size: usize, r_size: usize, blinds: *const SbnFr, c_out: *mut SbnG1a, inf_out: *mut u8) -> c_int;
    pub fn sbn_ctx_memory_stats(ctx: *mut sbn_ctx, out: *mut u64) -> c_int;
    pub fn sbn_host_alloc(out: *mut *mut c_void, bytes: usize) -> c_int;
    pub fn sbn_host_free(p: *mut c_void) -> c_int;
    pub fn sbn_stream_create(ctx: *mut sbn_ctx, stream_out: *mut *mut c_void) -> c_int;
    pub fn sbn_stream_synchronize(ctx: *mut sbn_ctx, stream: *mut c_void) -> c_int;
    pub fn sbn_stream_destroy(ctx: *mut sbn_ctx, stream: *mut c_void) -> c_int;
    pub fn sbn_msm(ctx: *mut sbn_ctx, pts: *const SbnG1a, inf: *const u8, s: *const SbnFr, n: usize,
                   out: *mut SbnG1a, inf_out: *mut u8) -> c_int;
    pub fn sbn_commit(ctx: *mut sbn_ctx, b: *const sbn_bases, s: *const SbnFr, n: usize, blind: *const SbnFr,
                      out: *mut SbnG1a, inf_out: *mut u8) -> c_int;
    pub fn sbn_bound(ctx: *mut sbn_ctx, z: *const SbnFr, l: *const SbnFr, l_size: usize, r_size: usize,
                     lz_out: *mut SbnFr) -> c_int;
    pub fn sbn_bases_create_ext(ctx: *mut sbn_ctx, g: *const SbnG1a, g_inf: *const u8, n: usize, g1: *const SbnG1a,
                                h: *const SbnG1a, out: *mut *mut sbn_bases) -> c_int;
    pub fn sbn_bullet_begin(ctx: *mut sbn_ctx, b: *const sbn_bases, q: *const SbnG1a, q_scalar: *const SbnFr, a: *const SbnFr, bv: *const SbnFr,
                            n: usize, blind: *const SbnFr, gamma: *mut SbnG1a, gamma_inf: *mut u8,
                            out: *mut *mut sbn_bullet) -> c_int;
    pub fn sbn_bullet_round(st: *mut sbn_bullet, blind_l: *const SbnFr, blind_r: *const SbnFr, l_out: *mut SbnG1a,
                            l_inf: *mut u8, r_out: *mut SbnG1a, r_inf: *mut u8) -> c_int;
    pub fn sbn_bullet_fold(st: *mut sbn_bullet, u: *const SbnFr, u_inv: *const SbnFr) -> c_int;
    pub fn sbn_bullet_end(st: *mut sbn_bullet, a_hat: *mut SbnFr, b_hat: *mut SbnFr, g_hat: *mut SbnG1a,
                          g_hat_inf: *mut u8) -> c_int;
    // nizk/mod.rs:497-500: delta = d * g_hat + r_delta * h returned with g_hat (a second row over the resident tables)
    pub fn sbn_bullet_end_delta(st: *mut sbn_bullet, d: *const SbnFr, r_delta: *const SbnFr, a_hat: *mut SbnFr, b_hat: *mut SbnFr,
                                g_hat: *mut SbnG1a, g_hat_inf: *mut u8, delta: *mut SbnG1a, delta_inf: *mut u8) -> c_int;
    pub fn sbn_bullet_destroy(st: *mut sbn_bullet) -> c_int;
    // resident polynomials (hyrax.rs:217-222, 283-324)
    pub fn sbn_poly_upload(ctx: *mut sbn_ctx, z: *const SbnFr, len: usize, out: *mut *mut sbn_poly) -> c_int;
    pub fn sbn_poly_destroy(p: *mut sbn_poly) -> c_int;
    pub fn sbn_poly_commit(ctx: *mut sbn_ctx, b: *const sbn_bases, p: *const sbn_poly, l_size: usize, r_size: usize,
                           blinds: *const SbnFr, c_out: *mut SbnG1a, inf_out: *mut u8) -> c_int;
    pub fn sbn_poly_commit_rows(ctx: *mut sbn_ctx, b: *const sbn_bases, p: *const sbn_poly, first_row: usize, n_rows: usize, r_size: usize,
                                blinds: *const SbnFr, c_out: *mut SbnG1a, inf_out: *mut u8) -> c_int;
    pub fn sbn_poly_bound(ctx: *mut sbn_ctx, p: *const sbn_poly, l: *const SbnFr, l_size: usize, r_size: usize, lz_out: *mut SbnFr) -> c_int;
    pub fn sbn_poly_evaluate(ctx: *mut sbn_ctx, p: *const sbn_poly, offset: usize, r: *const SbnFr, nr: usize, out: *mut SbnFr) -> c_int;
    pub fn sbn_poly_evaluate_strided(ctx: *mut sbn_ctx, p: *const sbn_poly, offset0: usize, stride: usize, count: usize,
                                     r: *const SbnFr, nr: usize, out: *mut SbnFr) -> c_int;
    pub fn sbn_poly_triple_dot(ctx: *mut sbn_ctx, a: *const sbn_poly, off_a: usize, b: *const sbn_poly, off_b: usize,
                               c: *const sbn_poly, off_c: usize, n: usize, out: *mut SbnFr) -> c_int;
    // R1CS-sat sumcheck rounds (sumcheck.rs:501-530, 690-699) and their inputs (r1csproof.rs:285, 380)
    pub fn sbn_sumcheck_begin(ctx: *mut sbn_ctx, tau: *const SbnFr, az: *const SbnFr, bz: *const SbnFr, cz: *const SbnFr,
                              len: usize, out: *mut *mut sbn_sumcheck) -> c_int;
    pub fn sbn_sumcheck_begin_quad(ctx: *mut sbn_ctx, z: *const SbnFr, abc: *const SbnFr, len: usize, out: *mut *mut sbn_sumcheck) -> c_int;
    // the same two set-ups with the tables built in HBM from the resident matrices (r1csproof.rs:268-290, 378-410)
    pub fn sbn_sumcheck_begin_r1cs(ctx: *mut sbn_ctx, mats: *const *const sbn_spmat, z: *const SbnFr, zlen: usize, tau: *const SbnFr,
                                   n_tau: usize, out: *mut *mut sbn_sumcheck) -> c_int;
    pub fn sbn_sumcheck_begin_r1cs_resident(ctx: *mut sbn_ctx, mats: *const *const sbn_spmat, vars: *const sbn_poly, tail: *const SbnFr,
                                            n_tail: usize, zlen: usize, tau: *const SbnFr, n_tau: usize, out: *mut *mut sbn_sumcheck) -> c_int;
    pub fn sbn_sumcheck_begin_quad_r1cs(ctx: *mut sbn_ctx, mats_t: *const *const sbn_spmat, coeffs: *const SbnFr, rx: *const SbnFr,
                                        n_rx: usize, z: *const SbnFr, zlen: usize, out: *mut *mut sbn_sumcheck) -> c_int;
    pub fn sbn_sumcheck_round_eval(st: *mut sbn_sumcheck, e0: *mut SbnFr, e2: *mut SbnFr, e3: *mut SbnFr) -> c_int;
    pub fn sbn_sumcheck_bind(st: *mut sbn_sumcheck, r: *const SbnFr) -> c_int;
    pub fn sbn_sumcheck_end(st: *mut sbn_sumcheck, finals: *mut SbnFr) -> c_int;
    pub fn sbn_sumcheck_destroy(st: *mut sbn_sumcheck) -> c_int;
    pub fn sbn_spmat_upload(ctx: *mut sbn_ctx, ptr: *const u32, idx: *const u32, val: *const SbnFr, n: usize, nnz: usize,
                            ncols: usize, out: *mut *mut sbn_spmat) -> c_int;
    pub fn sbn_spmat_destroy(m: *mut sbn_spmat) -> c_int;
    pub fn sbn_spmat_mulvec(ctx: *mut sbn_ctx, mats: *const *const sbn_spmat, coeffs: *const SbnFr, nm: usize, vec: *const SbnFr,
                            veclen: usize, out: *mut SbnFr) -> c_int;
    pub fn sbn_eq_evals(ctx: *mut sbn_ctx, r: *const SbnFr, n: usize, out: *mut SbnFr) -> c_int;
    // Spark: addresses / timestamps resident, derefs, hash layer, product layer (sparse_mlpoly_full.rs, product_tree.rs)
    pub fn sbn_addrs_upload(ctx: *mut sbn_ctx, row: *const u32, col: *const u32, batch: usize, n: usize, out: *mut *mut sbn_addrs) -> c_int;
    pub fn sbn_addrs_set_timestamps(a: *mut sbn_addrs, row_read: *const u32, row_audit: *const u32, col_read: *const u32,
                                    col_audit: *const u32, num_cells: usize) -> c_int;
    pub fn sbn_addrs_destroy(a: *mut sbn_addrs) -> c_int;
    pub fn sbn_spark_comb_polys(ctx: *mut sbn_ctx, a: *const sbn_addrs, val: *const SbnFr, comb_ops: *mut *mut sbn_poly,
                                comb_mem: *mut *mut sbn_poly) -> c_int;
    pub fn sbn_spark_evaluate(ctx: *mut sbn_ctx, a: *const sbn_addrs, comb_ops: *const sbn_poly, rx: *const SbnFr, nx: usize,
                              ry: *const SbnFr, ny: usize, out: *mut SbnFr) -> c_int;
    pub fn sbn_derefs_commit(ctx: *mut sbn_ctx, b: *const sbn_bases, a: *const sbn_addrs, rx: *const SbnFr, nx: usize, ry: *const SbnFr,
                             ny: usize, c_out: *mut SbnG1a, inf_out: *mut u8, poly_out: *mut *mut sbn_poly) -> c_int;
    pub fn sbn_derefs_commit_rows(ctx: *mut sbn_ctx, b: *const sbn_bases, a: *const sbn_addrs, rx: *const SbnFr, nx: usize,
                                  ry: *const SbnFr, ny: usize, row0: usize, nrows: usize, c_out: *mut SbnG1a, inf_out: *mut u8,
                                  poly_out: *mut *mut sbn_poly) -> c_int;
    pub fn sbn_hashlayer_build(ctx: *mut sbn_ctx, a: *const sbn_addrs, side: c_int, r: *const SbnFr, nr: usize, r_hash: *const SbnFr,
                               r_multiset_check: *const SbnFr, circuits_out: *mut *mut sbn_prodcircuit) -> c_int;
    pub fn sbn_prodcircuit_create(ctx: *mut sbn_ctx, poly: *const SbnFr, len: usize, out: *mut *mut sbn_prodcircuit) -> c_int;
    pub fn sbn_prodcircuit_evaluate(pc: *mut sbn_prodcircuit, out: *mut SbnFr) -> c_int;
    pub fn sbn_prodcircuit_destroy(pc: *mut sbn_prodcircuit) -> c_int;
    pub fn sbn_bsumcheck_begin_resident(ctx: *mut sbn_ctx, circuits: *const *mut sbn_prodcircuit, p: usize, layer_id: usize,
                                        rand: *const SbnFr, n_rand: usize, seq_polys: *const *const sbn_poly,
                                        seq_offsets: *const usize, s: usize, out: *mut *mut sbn_bsumcheck) -> c_int;
    pub fn sbn_bsumcheck_round_eval(st: *mut sbn_bsumcheck, evals: *mut SbnFr) -> c_int;
    pub fn sbn_bsumcheck_bind(st: *mut sbn_bsumcheck, r: *const SbnFr) -> c_int;
    pub fn sbn_bsumcheck_end(st: *mut sbn_bsumcheck, a_final: *mut SbnFr, b_final: *mut SbnFr, c_final: *mut SbnFr) -> c_int;
    // one layer's whole round loop against the caller's Merlin state (only meaningful when the transcript is the
    // library's own sbn_merlin_* state; a Rust prover keeps merlin::Transcript and drives round_eval / bind instead)
    pub fn sbn_bsumcheck_prove(st: *mut sbn_bsumcheck, merlin: *mut c_void, claim: *const SbnFr, coeffs: *const SbnFr,
                               num_rounds: usize, polys: *mut SbnFr, r_out: *mut SbnFr, claim_out: *mut SbnFr,
                               a_final: *mut SbnFr, b_final: *mut SbnFr, c_final: *mut SbnFr) -> c_int;
    pub fn sbn_bsumcheck_destroy(st: *mut sbn_bsumcheck) -> c_int;
}

fn check(status: c_int, what: &str) {
    // The reference's preconditions are assert!/panic (hyrax.rs:258,290,295; commitments.rs:146): keep that.
    if status != 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(sbn_strerror(status)) }.to_string_lossy().into_owned();
        panic!("{what}: libsbn254 status {status} ({msg})");
    }
}

#[inline] fn fq_limbs(v: &Fq) -> [u64; 4] { (v.0).0 }          // Montgomery limbs as stored by ark-ff
#[inline] fn fq_from_limbs(l: [u64; 4]) -> Fq { ark_ff::Fp(BigInt(l), core::marker::PhantomData) }

pub fn to_abi(p: &G1Affine) -> (SbnG1a, u8) {
    if p.infinity { (SbnG1a::default(), 1) } else { (SbnG1a { x: fq_limbs(&p.x), y: fq_limbs(&p.y) }, 0) }
}
pub fn from_abi(p: &SbnG1a, inf: u8) -> G1Affine {
    if inf != 0 { G1Affine::identity() } else { G1Affine::new_unchecked(fq_from_limbs(p.x), fq_from_limbs(p.y)) }
}

pub struct Context(pub *mut sbn_ctx);
unsafe impl Send for Context {}
unsafe impl Sync for Context {}            // the library serialises calls per context
impl Context {
    pub fn new(device: i32) -> Self { let mut h = std::ptr::null_mut(); check(unsafe { sbn_ctx_create(device, &mut h) }, "sbn_ctx_create"); Context(h) }
}
impl Drop for Context { fn drop(&mut self) { unsafe { sbn_ctx_destroy(self.0) }; } }

/// A `MultiCommitGens` resident on the GPU (upload once per generator set; commitments.rs:17-27).
pub struct Bases { pub h: *mut sbn_bases, pub n: usize }
unsafe impl Send for Bases {}
unsafe impl Sync for Bases {}
impl Bases {
    pub fn new(ctx: &Context, g_affine: &[G1Affine], h_affine: &G1Affine) -> Self {
        let (pts, inf): (Vec<_>, Vec<_>) = g_affine.iter().map(to_abi).unzip();
        let (h, _) = to_abi(h_affine);
        let mut out = std::ptr::null_mut();
        check(unsafe { sbn_bases_create(ctx.0, pts.as_ptr(), inf.as_ptr(), pts.len(), &h, &mut out) }, "sbn_bases_create");
        Bases { h: out, n: pts.len() }
    }
}
impl Drop for Bases { fn drop(&mut self) { unsafe { sbn_bases_destroy(self.h) }; } }

/// Drop-in body of `DensePolynomial::commit_inner` (hyrax.rs:253-281): `z` is the row-major evaluation
/// vector, `blinds.len()` = L_size.  `Fr` slices are passed as-is (Montgomery limbs).
pub fn hyrax_commit(ctx: &Context, bases: &Bases, z: &[Fr], blinds: &[Fr]) -> Vec<G1Affine> {
    let l_size = blinds.len();
    let r_size = z.len() / l_size;
    assert_eq!(l_size * r_size, z.len());
    let mut c = vec![SbnG1a::default(); l_size];
    let mut inf = vec![0u8; l_size];
    let zero_blinds = blinds.iter().all(|b| b.is_zero_vartime());
    let bp = if zero_blinds { std::ptr::null() } else { blinds.as_ptr() as *const SbnFr };
    check(unsafe { sbn_hyrax_commit(ctx.0, bases.h, z.as_ptr() as *const SbnFr, l_size, r_size, bp, c.as_mut_ptr(), inf.as_mut_ptr()) },
          "sbn_hyrax_commit");
    c.iter().zip(inf.iter()).map(|(p, i)| from_abi(p, *i)).collect()
}

/// The same commit over several GPUs of this process (`Context` per device, `Bases` per context): contiguous row blocks,
/// one host thread per device inside the library, no inter-GPU traffic (the Rayon fan-out of hyrax.rs:259-265).
pub fn hyrax_commit_multi(ctxs: &[&Context], bases: &[&Bases], z: &[Fr], blinds: &[Fr]) -> Vec<G1Affine> {
    assert_eq!(ctxs.len(), bases.len());
    let l_size = blinds.len();
    let r_size = z.len() / l_size;
    assert_eq!(l_size * r_size, z.len());
    let cp: Vec<*mut sbn_ctx> = ctxs.iter().map(|c| c.0).collect();
    let bp: Vec<*const sbn_bases> = bases.iter().map(|b| b.h as *const sbn_bases).collect();
    let mut c = vec![SbnG1a::default(); l_size];
    let mut inf = vec![0u8; l_size];
    let zero_blinds = blinds.iter().all(|b| b.is_zero_vartime());
    let blp = if zero_blinds { std::ptr::null() } else { blinds.as_ptr() as *const SbnFr };
    check(unsafe { sbn_hyrax_commit_multi(cp.as_ptr(), bp.as_ptr(), cp.len(), z.as_ptr() as *const SbnFr, l_size, r_size, blp,
                                          c.as_mut_ptr(), inf.as_mut_ptr()) }, "sbn_hyrax_commit_multi");
    c.iter().zip(inf.iter()).map(|(p, i)| from_abi(p, *i)).collect()
}

/// Drop-in body of `GroupElement::msm_affine` (group.rs:171-175).  arkworks' `G1Projective::msm` returns `Err` when the
/// slices differ in length and the reference swallows it with `unwrap_or_default()` (group.rs:156,173): the identity.  The
/// same here, without touching the GPU.  An empty MSM is the identity too.
pub fn msm_affine(ctx: &Context, scalars: &[Fr], points: &[G1Affine]) -> G1Affine {
    if scalars.len() != points.len() || scalars.is_empty() {
        return G1Affine::identity();
    }
    let (pts, inf): (Vec<_>, Vec<_>) = points.iter().map(to_abi).unzip();
    let mut out = SbnG1a::default();
    let mut out_inf = 0u8;
    check(unsafe { sbn_msm(ctx.0, pts.as_ptr(), inf.as_ptr(), scalars.as_ptr() as *const SbnFr, scalars.len(), &mut out, &mut out_inf) },
          "sbn_msm");
    from_abi(&out, out_inf)
}

/// `<[Scalar] as Commitments>::commit` (commitments.rs:144-154) over a resident generator set.
pub fn commit(ctx: &Context, bases: &Bases, scalars: &[Fr], blind: &Fr) -> G1Affine {
    assert_eq!(bases.n, scalars.len());                                  // commitments.rs:146
    let mut out = SbnG1a::default();
    let mut out_inf = 0u8;
    check(unsafe { sbn_commit(ctx.0, bases.h, scalars.as_ptr() as *const SbnFr, scalars.len(), blind as *const Fr as *const SbnFr,
                              &mut out, &mut out_inf) }, "sbn_commit");
    from_abi(&out, out_inf)
}

/// Drop-in body of `DensePolynomial::bound` (hyrax.rs:311-324): LZ[i] = sum_j L[j] * Z[j * R_size + i].
pub fn bound(ctx: &Context, z: &[Fr], l: &[Fr]) -> Vec<Fr> {
    let l_size = l.len();
    let r_size = z.len() / l_size;
    assert_eq!(l_size * r_size, z.len());
    let mut out = vec![SbnFr::default(); r_size];
    check(unsafe { sbn_bound(ctx.0, z.as_ptr() as *const SbnFr, l.as_ptr() as *const SbnFr, l_size, r_size, out.as_mut_ptr()) },
          "sbn_bound");
    out.into_iter().map(|v| ark_ff::Fp(BigInt(v.l), core::marker::PhantomData)).collect()
}

/// `BulletReductionProof::prove` (nizk/bullet.rs:24-126) with a, b, the folding coefficients and the generator tables
/// resident on the GPU.  The caller keeps the transcript: per round it appends L, R, draws u and hands it back.
///
/// ```ignore
/// let mut red = BulletReduction::begin(&ctx, &bases_ext, &q, None, a, b, &blind);   // Gamma = red.gamma (bullet.rs:57)
/// for (bl, br) in blinds_vec {                                                      // bullet.rs:63
///     let (l, r) = red.round(bl, br);                                               // :75-76
///     transcript.append_point(b"L", &l.compress()); transcript.append_point(b"R", &r.compress());
///     let u = transcript.challenge_scalar(b"u");                                    // :81-83
///     red.fold(&u, &u.invert().unwrap());                                           // :85-104
/// }
/// let (a_hat, b_hat, g_hat) = red.end();                                            // :108-125
/// ```
pub struct BulletReduction { st: *mut sbn_bullet, pub gamma: G1Affine }
impl BulletReduction {
    pub fn begin(ctx: &Context, bases_ext: &Bases, q: &G1Affine, q_scalar: Option<&Fr>, a: &[Fr], b: &[Fr], blind: &Fr) -> Self {
        assert_eq!(a.len(), b.len());                                    // bullet.rs:42-47
        assert!(a.len().is_power_of_two());
        let (qa, _) = to_abi(q);
        let mut st = std::ptr::null_mut();
        let mut gamma = SbnG1a::default();
        let mut ginf = 0u8;
        let qs = q_scalar.map(|s| s as *const Fr as *const SbnFr).unwrap_or(std::ptr::null());
        check(unsafe { sbn_bullet_begin(ctx.0, bases_ext.h, &qa, qs, a.as_ptr() as *const SbnFr, b.as_ptr() as *const SbnFr, a.len(),
                                        blind as *const Fr as *const SbnFr, &mut gamma, &mut ginf, &mut st) }, "sbn_bullet_begin");
        BulletReduction { st, gamma: from_abi(&gamma, ginf) }
    }
    pub fn round(&mut self, blind_l: &Fr, blind_r: &Fr) -> (G1Affine, G1Affine) {
        let (mut l, mut r, mut li, mut ri) = (SbnG1a::default(), SbnG1a::default(), 0u8, 0u8);
        check(unsafe { sbn_bullet_round(self.st, blind_l as *const Fr as *const SbnFr, blind_r as *const Fr as *const SbnFr,
                                        &mut l, &mut li, &mut r, &mut ri) }, "sbn_bullet_round");
        (from_abi(&l, li), from_abi(&r, ri))
    }
    pub fn fold(&mut self, u: &Fr, u_inv: &Fr) {
        check(unsafe { sbn_bullet_fold(self.st, u as *const Fr as *const SbnFr, u_inv as *const Fr as *const SbnFr) }, "sbn_bullet_fold");
    }
    pub fn end(self) -> (Fr, Fr, G1Affine) {
        let (mut a, mut b, mut g, mut gi) = (SbnFr::default(), SbnFr::default(), SbnG1a::default(), 0u8);
        check(unsafe { sbn_bullet_end(self.st, &mut a, &mut b, &mut g, &mut gi) }, "sbn_bullet_end");
        (ark_ff::Fp(BigInt(a.l), core::marker::PhantomData), ark_ff::Fp(BigInt(b.l), core::marker::PhantomData), from_abi(&g, gi))
    }
}
impl Drop for BulletReduction { fn drop(&mut self) { unsafe { sbn_bullet_destroy(self.st) }; } }

/// A generator set with gens_1's point addressable (DotProductProofGens, nizk/mod.rs:412-415): what the opening uses.
impl Bases {
    pub fn new_ext(ctx: &Context, g_affine: &[G1Affine], g1: &G1Affine, h_affine: &G1Affine) -> Self {
        let (pts, inf): (Vec<_>, Vec<_>) = g_affine.iter().map(to_abi).unzip();
        let (h, _) = to_abi(h_affine);
        let (g1a, _) = to_abi(g1);
        let mut out = std::ptr::null_mut();
        check(unsafe { sbn_bases_create_ext(ctx.0, pts.as_ptr(), inf.as_ptr(), pts.len(), &g1a, &h, &mut out) }, "sbn_bases_create_ext");
        Bases { h: out, n: pts.len() }
    }
}

/// Pinned host memory for a polynomial's evaluations: `sbn_hyrax_commit` copies from pageable memory at roughly half the rate
/// (bench.py `e2e.pageable_value`); a prover that builds its Z vectors in this buffer gets the pinned figure.
pub struct PinnedScalars { ptr: *mut Fr, len: usize }
impl PinnedScalars {
    pub fn new(len: usize) -> Self {
        let mut p: *mut c_void = std::ptr::null_mut();
        check(unsafe { sbn_host_alloc(&mut p, len * std::mem::size_of::<Fr>()) }, "sbn_host_alloc");
        unsafe { std::ptr::write_bytes(p as *mut u8, 0, len * std::mem::size_of::<Fr>()) };       // all-zero limbs = Fr::zero()
        PinnedScalars { ptr: p as *mut Fr, len }
    }
    pub fn as_mut_slice(&mut self) -> &mut [Fr] { unsafe { std::slice::from_raw_parts_mut(self.ptr, self.len) } }
    pub fn as_slice(&self) -> &[Fr] { unsafe { std::slice::from_raw_parts(self.ptr, self.len) } }
}
impl Drop for PinnedScalars { fn drop(&mut self) { unsafe { sbn_host_free(self.ptr as *mut c_void) }; } }

/// A stream of the library's own for the asynchronous commits (no CUDA bindings needed on the Rust side).
pub struct Stream<'a> { ctx: &'a Context, h: *mut c_void }
impl<'a> Stream<'a> {
    pub fn new(ctx: &'a Context) -> Self {
        let mut h: *mut c_void = std::ptr::null_mut();
        check(unsafe { sbn_stream_create(ctx.0, &mut h) }, "sbn_stream_create");
        Stream { ctx, h }
    }
    pub fn synchronize(&self) { check(unsafe { sbn_stream_synchronize(self.ctx.0, self.h) }, "sbn_stream_synchronize"); }
}
impl<'a> Drop for Stream<'a> { fn drop(&mut self) { unsafe { sbn_stream_destroy(self.ctx.0, self.h) }; } }

/// A commit in flight (`hyrax_commit_async`): the commitments are valid after `wait()`.  Holds its inputs borrowed, so the
/// scalars cannot be dropped or changed while the copies run.
pub struct PendingCommit<'a> { stream: &'a Stream<'a>, c: Vec<SbnG1a>, inf: Vec<u8>, _z: &'a [Fr], _b: Option<&'a [Fr]> }
impl<'a> PendingCommit<'a> {
    pub fn wait(self) -> Vec<G1Affine> {
        self.stream.synchronize();
        self.c.iter().zip(self.inf.iter()).map(|(p, i)| from_abi(p, *i)).collect()
    }
}

/// `DensePolynomial::commit_inner` (hyrax.rs:253-281), asynchronous: returns at once; several commits issued on two or three
/// streams overlap (the copy of one under the kernels of the others).  `z` and `blinds` may be ordinary (pageable) vectors:
/// the library stages them through pinned buffers and lands the results through a stream-ordered host function.
pub fn hyrax_commit_async<'a>(ctx: &'a Context, bases: &Bases, z: &'a [Fr], blinds: Option<&'a [Fr]>, l_size: usize,
                              stream: &'a Stream<'a>) -> PendingCommit<'a> {
    assert!(l_size > 0 && z.len() % l_size == 0, "assert_eq!(L_size * R_size, self.Z.len())");          // hyrax.rs:258
    if let Some(b) = blinds { assert_eq!(b.len(), l_size); }
    let r_size = z.len() / l_size;
    let mut c = vec![SbnG1a::default(); l_size];
    let mut inf = vec![0u8; l_size];
    let bp = blinds.map(|b| b.as_ptr() as *const SbnFr).unwrap_or(std::ptr::null());
    check(unsafe { sbn_hyrax_commit_async(ctx.0, bases.h, z.as_ptr() as *const SbnFr, l_size, r_size, bp, c.as_mut_ptr(),
                                          inf.as_mut_ptr(), stream.h) }, "sbn_hyrax_commit_async");
    PendingCommit { stream, c, inf, _z: z, _b: blinds }
}

trait IsZeroVartime { fn is_zero_vartime(&self) -> bool; }
impl IsZeroVartime for Fr { fn is_zero_vartime(&self) -> bool { self.into_bigint().0 == [0u64; 4] } }

#[allow(dead_code)] fn _unused(_: *mut c_void) {}
