//! Dumps byte-level outputs of the reference (Antiparadox/Spartan-BN254) for the seeded inputs that
//! `tests/golden/make_golden.py` uses, so that the oracle of this repository can be PINNED to the real reference:
//!
//!   * generators of both labels (`MultiCommitGens::new`, commitments.rs:31-62 / group.rs:110-132): affine coordinates and
//!     compressed bytes of the first 16 points and of `h`, plus how many of 1025 are distinct;
//!   * `DotProductProofGens::new(8, ..)` (nizk/mod.rs:412-415): gens_n, gens_1, h;
//!   * the 4 x 8 Hyrax commit of make_golden.py with zero blinds (`DensePolynomial::commit(gens, None)`, hyrax.rs:283-308)
//!     and the commit of one row with a blind (`<[Scalar] as Commitments>::commit`, commitments.rs:144-154);
//!   * `EqPolynomial::evals`, `compute_factored_evals`, `DensePolynomial::bound` (hyrax.rs:311-324, 355-383);
//!   * `GroupElement::msm_affine` of the 5 G known-answer test (group.rs:313-321) and a seeded 8-point MSM;
//!   * a Merlin transcript of the reference's own traits (transcript.rs:38-108): protocol name, scalars, a point, then
//!     `challenge_scalar` / `challenge_vector`.
//!
//! Everything here is deterministic (no `RandomTape`: it seeds itself from `OsRng`, random.rs:15-22, so proofs with blinds
//! are not reproducible byte for byte; `tests/test_snark.py` pins those through the verifier instead).
//! Scalars are written as canonical little-endian hex (`Scalar::to_bytes`, scalar.rs:75-84), points as affine
//! (x, y) canonical big-endian hex like tests/golden/hyrax_golden.json, compressed points as hex of the 32 bytes.
use ark_ec::{AffineRepr, CurveGroup};
use ark_ff::{BigInteger, PrimeField};
use merlin::Transcript;
use serde_json::{json, Value};
use spartan_bn254::commitments::{Commitments, MultiCommitGens};
use spartan_bn254::group::GroupElement;
use spartan_bn254::hyrax::{DensePolynomial, EqPolynomial, PolyCommitmentGens};
use spartan_bn254::nizk::DotProductProofGens;
use spartan_bn254::scalar::Scalar;
use spartan_bn254::transcript::{AppendToTranscript, ProofTranscript};

/// SplitMix64 exactly as oracle/pymodel.py: four outputs = one scalar, LE limbs, reduced mod r.
struct SplitMix64(u64);
impl SplitMix64 {
    fn next(&mut self) -> u64 {
        self.0 = self.0.wrapping_add(0x9E3779B97F4A7C15);
        let mut z = self.0;
        z = (z ^ (z >> 30)).wrapping_mul(0xBF58476D1CE4E5B9);
        z = (z ^ (z >> 27)).wrapping_mul(0x94D049BB133111EB);
        z ^ (z >> 31)
    }
    fn scalar(&mut self) -> Scalar {
        let mut bytes = [0u8; 32];
        for i in 0..4 {
            bytes[8 * i..8 * i + 8].copy_from_slice(&self.next().to_le_bytes());
        }
        Scalar(ark_bn254::Fr::from_le_bytes_mod_order(&bytes))
    }
}

fn hex_le(s: &Scalar) -> String {
    // canonical value as 0x + 64 hex digits, most significant first (the format of hyrax_golden.json)
    let mut be = s.to_bytes();
    be.reverse();
    format!("0x{}", be.iter().map(|b| format!("{:02x}", b)).collect::<String>())
}
fn fq_hex(v: &ark_bn254::Fq) -> String {
    let be = v.into_bigint().to_bytes_be();
    format!("0x{}", be.iter().map(|b| format!("{:02x}", b)).collect::<String>())
}
fn point(p: &GroupElement) -> Value {
    let a = p.0.into_affine();
    if a.is_zero() {
        Value::Null
    } else {
        json!([fq_hex(&a.x().unwrap()), fq_hex(&a.y().unwrap())])
    }
}
fn compressed(p: &GroupElement) -> String {
    p.compress().as_bytes().iter().map(|b| format!("{:02x}", b)).collect()
}

fn gens_fixture(label: &'static [u8]) -> Value {
    let g = MultiCommitGens::new(16, label);
    let big = MultiCommitGens::new(1024, label);
    let mut distinct: Vec<String> = big.G.iter().map(compressed).collect();
    distinct.push(compressed(&big.h));
    distinct.sort();
    distinct.dedup();
    json!({
        "points": g.G.iter().map(point).collect::<Vec<_>>(),
        "compressed": g.G.iter().map(compressed).collect::<Vec<_>>(),
        "h": point(&g.h),
        "distinct_of_1025": distinct.len(),
    })
}

fn main() {
    let out_path = std::env::args().nth(1).unwrap_or_else(|| "reference_fixtures.json".into());
    let mut out = serde_json::Map::new();

    // ---- generators
    let mut gens = serde_json::Map::new();
    gens.insert("gens_r1cs_sat".into(), gens_fixture(b"gens_r1cs_sat"));
    gens.insert("gens_r1cs_eval".into(), gens_fixture(b"gens_r1cs_eval"));
    out.insert("generators".into(), Value::Object(gens));
    let d = DotProductProofGens::new(8, b"gens_r1cs_eval");
    out.insert("dotproduct_gens_8".into(), json!({
        "gens_n": d.gens_n.G.iter().map(point).collect::<Vec<_>>(), "gens_1": point(&d.gens_1.G[0]), "h": point(&d.gens_n.h),
    }));

    // ---- the 4 x 8 commit of make_golden.py: SplitMix64(7), 32 scalars, last row zero; zero blinds here
    let mut rng = SplitMix64(7);
    let mut z: Vec<Scalar> = (0..32).map(|_| rng.scalar()).collect();
    for j in 0..8 { z[24 + j] = Scalar::zero(); }
    let blind1 = { let _b0 = rng.scalar(); rng.scalar() };          // make_golden.py draws blinds[0], blinds[1] next
    let pc_gens = PolyCommitmentGens::new(5, b"gens_r1cs_eval");
    let poly = DensePolynomial::new(z.clone());
    let (comm, _blinds) = poly.commit(&pc_gens, None);
    let row1_blinded = z[8..16].commit(&blind1, &pc_gens.gens.gens_n);
    out.insert("hyrax_commit_4x8_zero_blinds".into(), json!({
        "Z": z.iter().map(hex_le).collect::<Vec<_>>(),
        "C": comm.C.iter().map(point).collect::<Vec<_>>(),
        "C_compressed": comm.C.iter().map(compressed).collect::<Vec<_>>(),
        "row1_blind": hex_le(&blind1), "row1_commit_with_blind": point(&row1_blinded),
    }));
    // PolyCommitment::append_to_transcript (hyrax.rs:44-52) followed by a challenge
    let mut t = Transcript::new(b"fixture");
    comm.append_to_transcript(b"poly_commitment", &mut t);
    out.insert("transcript_after_commitment_challenge".into(), json!(hex_le(&t.challenge_scalar(b"c"))));

    // ---- eq tables and bound (seeded point r, 5 variables)
    let r: Vec<Scalar> = (0..5).map(|_| rng.scalar()).collect();
    let eq = EqPolynomial::new(r.clone());
    let (lv, rv) = eq.compute_factored_evals();
    out.insert("bound_4x8".into(), json!({
        "r": r.iter().map(hex_le).collect::<Vec<_>>(),
        "eq_evals": eq.evals().iter().map(hex_le).collect::<Vec<_>>(),
        "L": lv.iter().map(hex_le).collect::<Vec<_>>(), "R": rv.iter().map(hex_le).collect::<Vec<_>>(),
        "LZ": poly.bound(&lv).iter().map(hex_le).collect::<Vec<_>>(),
        "evaluate": hex_le(&poly.evaluate(&r)),
    }));

    // ---- MSM: the reference's own 5 G test (group.rs:313-321) and a seeded 8-point one over gens_n
    let g = GroupElement::generator();
    let ga = g.0.into_affine();
    let five = GroupElement::msm_affine(&[Scalar::from_u64(2), Scalar::from_u64(3)], &[ga, ga]);
    let s8: Vec<Scalar> = (0..8).map(|_| rng.scalar()).collect();
    let msm8 = GroupElement::msm_affine(&s8, &pc_gens.gens.gens_n.G_affine);
    out.insert("msm".into(), json!({
        "five_G": point(&five), "scalars8": s8.iter().map(hex_le).collect::<Vec<_>>(), "msm8_over_gens_n": point(&msm8),
        "compress_G": compressed(&g), "compress_identity": compressed(&GroupElement::identity()),
    }));

    // ---- Merlin through the reference's ProofTranscript (transcript.rs:38-76)
    let mut t = Transcript::new(b"fixture-transcript");
    t.append_protocol_name(b"protocol");
    t.append_scalar(b"s", &s8[0]);
    t.append_scalars(b"v", &s8[1..4]);
    t.append_point(b"p", &msm8.compress());
    let c1 = t.challenge_scalar(b"c1");
    let cv = t.challenge_vector(b"cv", 3);
    out.insert("transcript".into(), json!({"c1": hex_le(&c1), "cv": cv.iter().map(hex_le).collect::<Vec<_>>()}));

    std::fs::write(&out_path, serde_json::to_string_pretty(&Value::Object(out)).unwrap()).unwrap();
    eprintln!("wrote {out_path}");
}
