"""Pins the oracle to bytes produced by the REAL reference, when they exist.

`rust/fixtures` (a cargo crate with the reference as a path dependency; not buildable in this image: no Rust toolchain) writes
tests/golden/reference_fixtures.json for the seeded inputs of tests/golden/make_golden.py.  When that file is present,
`test_oracle_matches_reference_fixtures` checks the C oracle (the checker of every GPU parity test) against it value by value
-- generators, Hyrax commits, eq tables, `bound`, `evaluate`, MSMs, point compression, Merlin challenges.  Until then the
parity of this repository is UNPINNED by the reference and the test is skipped with that message.

`test_fixture_loader_on_emulated_fixture` runs the same checks on a fixture of the same layout produced by the independent
Python big-int model (oracle/pymodel.py), so the loader and the layout are exercised on every CI run."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURE = os.path.join(HERE, "golden", "reference_fixtures.json")


def hx(v):
    return "0x%064x" % v


def h2i(s):
    return int(s, 16)


def pt(p):
    return None if p is None else [hx(p[0]), hx(p[1])]


def emulate_fixture():
    """The layout rust/fixtures/src/main.rs writes, filled in by oracle/pymodel.py."""
    import pymodel as pm
    import hashlib  # noqa: F401
    out = {}
    gens = {}
    for label in (b"gens_r1cs_sat", b"gens_r1cs_eval"):
        G16, h16 = pm.multi_commit_gens(16, label)
        sc = pm.gen_scalars(1024, label)
        gens[label.decode()] = {"points": [pt(p) for p in G16], "compressed": [pm.compress(p).hex() for p in G16],
                                "h": pt(h16), "distinct_of_1025": len(set(s for s, _ in sc))}
    out["generators"] = gens
    (gn, h), (g1, _) = pm.dotproduct_gens(8, b"gens_r1cs_eval")
    out["dotproduct_gens_8"] = {"gens_n": [pt(p) for p in gn], "gens_1": pt(g1[0]), "h": pt(h)}
    rng = pm.SplitMix64(7)
    Z = [rng.scalar() for _ in range(32)]
    for j in range(8):
        Z[24 + j] = 0
    rng.scalar()
    blind1 = rng.scalar()
    C = pm.hyrax_commit(Z, [0, 0, 0, 0], (gn, h))
    out["hyrax_commit_4x8_zero_blinds"] = {
        "Z": [hx(z) for z in Z], "C": [pt(c) for c in C], "C_compressed": [pm.compress(c).hex() for c in C],
        "row1_blind": hx(blind1), "row1_commit_with_blind": pt(pm.commit_row(Z[8:16], blind1, (gn, h)))}
    r = [rng.scalar() for _ in range(5)]
    Lv, Rv = pm.factored_evals(r)
    eq = pm.eq_evals(r)
    out["bound_4x8"] = {"r": [hx(x) for x in r], "eq_evals": [hx(x) for x in eq], "L": [hx(x) for x in Lv], "R": [hx(x) for x in Rv],
                        "LZ": [hx(x) for x in pm.bound(Z, Lv, 5)], "evaluate": hx(sum(a * b for a, b in zip(eq, Z)) % pm.R)}
    s8 = [rng.scalar() for _ in range(8)]
    out["msm"] = {"five_G": pt(pm.msm([2, 3], [pm.G, pm.G])), "scalars8": [hx(x) for x in s8],
                  "msm8_over_gens_n": pt(pm.msm(s8, gn)), "compress_G": pm.compress(pm.G).hex(),
                  "compress_identity": pm.compress(None).hex()}
    return out, s8, C


def check_fixture(fx, orc, transcripts=True):
    """Every value of the fixture against the C oracle (oracle/bn254_oracle.c through oracle/oracle.py)."""
    # generators: MultiCommitGens::new(16, label), and the number of distinct points among the 1025 of new(1024, label)
    for label, g in fx["generators"].items():
        G, h = orc.multi_commit_gens(label.encode(), 16)
        got = orc.points_to_ints(G, [0] * 16)
        assert got == [(h2i(p[0]), h2i(p[1])) for p in g["points"]], "generators of %s" % label
        assert orc.points_to_ints(h.reshape(1, 8), [0])[0] == (h2i(g["h"][0]), h2i(g["h"][1]))
        assert [orc.compress(G[i], 0).hex() for i in range(16)] == g["compressed"]
        Gb, hb = orc.multi_commit_gens(label.encode(), 1024)
        pts = np.concatenate([Gb, hb.reshape(1, 8)]).view(np.uint8).reshape(1025, 64)
        assert len({bytes(p) for p in pts}) == g["distinct_of_1025"]
    d = fx["dotproduct_gens_8"]
    Gn, hh, G1 = orc.dotproduct_gens(b"gens_r1cs_eval", 8)
    assert orc.points_to_ints(Gn, [0] * 8) == [(h2i(p[0]), h2i(p[1])) for p in d["gens_n"]]
    assert orc.points_to_ints(G1.reshape(1, 8), [0])[0] == (h2i(d["gens_1"][0]), h2i(d["gens_1"][1]))
    assert orc.points_to_ints(hh.reshape(1, 8), [0])[0] == (h2i(d["h"][0]), h2i(d["h"][1]))
    # DensePolynomial::commit with zero blinds, one blinded row
    c = fx["hyrax_commit_4x8_zero_blinds"]
    Z = orc.to_mont([h2i(z) for z in c["Z"]])
    C, inf = orc.hyrax_commit(Gn, hh, Z, 4, 8, None)
    exp = [None if p is None else (h2i(p[0]), h2i(p[1])) for p in c["C"]]
    assert orc.points_to_ints(C, inf) == exp
    assert [orc.compress(C[i], inf[i]).hex() for i in range(4)] == c["C_compressed"]
    assert inf[3] == 1
    bl = orc.to_mont([h2i(c["row1_blind"])])
    C1, inf1 = orc.hyrax_commit(Gn, hh, Z[8:16], 1, 8, bl)
    assert orc.points_to_ints(C1, inf1)[0] == (h2i(c["row1_commit_with_blind"][0]), h2i(c["row1_commit_with_blind"][1]))
    # eq tables, bound, evaluate
    b = fx["bound_4x8"]
    r = orc.to_mont([h2i(x) for x in b["r"]])
    assert orc.from_mont(orc.eq_evals(r)) == [h2i(x) for x in b["eq_evals"]]
    Lv = orc.eq_evals(r[:2])
    assert orc.from_mont(Lv) == [h2i(x) for x in b["L"]]
    assert orc.from_mont(orc.eq_evals(r[2:])) == [h2i(x) for x in b["R"]]
    assert orc.from_mont(orc.bound(Z, Lv, 4, 8)) == [h2i(x) for x in b["LZ"]]
    assert orc.from_mont(orc.evaluate(Z, r).reshape(1, 4))[0] == h2i(b["evaluate"])
    # MSM, compression
    m = fx["msm"]
    g = orc.generator()
    two_pts = np.stack([g, g])
    out, oinf = orc.msm(two_pts, np.zeros(2, dtype=np.uint8), orc.to_mont([2, 3]))
    assert orc.points_to_ints(out.reshape(1, 8), [oinf])[0] == (h2i(m["five_G"][0]), h2i(m["five_G"][1]))
    s8 = orc.to_mont([h2i(x) for x in m["scalars8"]])
    out8, oinf8 = orc.msm(Gn, np.zeros(8, dtype=np.uint8), s8)
    assert orc.points_to_ints(out8.reshape(1, 8), [oinf8])[0] == (h2i(m["msm8_over_gens_n"][0]), h2i(m["msm8_over_gens_n"][1]))
    assert orc.compress(g, 0).hex() == m["compress_G"]
    assert orc.compress(np.zeros(8, dtype=np.uint64), 1).hex() == m["compress_identity"]
    if not transcripts:
        return
    # Merlin through transcript.rs: PolyCommitment::append_to_transcript + challenge, then the scalar / point appends
    t = orc.Transcript(b"fixture")
    t.append_message(b"poly_commitment", b"poly_commitment_begin")
    for i in range(4):
        t.append_message(b"poly_commitment_share", orc.compress(C[i], inf[i]))
    t.append_message(b"poly_commitment", b"poly_commitment_end")
    assert orc.from_mont(t.challenge_scalar(b"c").reshape(1, 4))[0] == h2i(fx["transcript_after_commitment_challenge"])
    t = orc.Transcript(b"fixture-transcript")
    t.append_message(b"protocol-name", b"protocol")
    canon = [h2i(x) for x in m["scalars8"]]
    t.append_message(b"s", canon[0].to_bytes(32, "little"))
    for v in canon[1:4]:
        t.append_message(b"v", v.to_bytes(32, "little"))
    t.append_message(b"p", orc.compress(out8, oinf8))
    assert orc.from_mont(t.challenge_scalar(b"c1").reshape(1, 4))[0] == h2i(fx["transcript"]["c1"])
    assert [orc.from_mont(t.challenge_scalar(b"cv").reshape(1, 4))[0] for _ in range(3)] == [h2i(x) for x in fx["transcript"]["cv"]]


def test_fixture_loader_on_emulated_fixture(orc):
    fx, _, _ = emulate_fixture()
    check_fixture(fx, orc, transcripts=False)      # the Python model has no Merlin; test_transcript.py covers that against Merlin's vector


def test_oracle_matches_reference_fixtures(orc):
    if not os.path.exists(FIXTURE):
        pytest.skip("tests/golden/reference_fixtures.json absent: parity is UNPINNED by the reference until rust/fixtures is run "
                    "on a machine with cargo (cd rust/fixtures && cargo run --release -- ../../tests/golden/reference_fixtures.json)")
    check_fixture(json.load(open(FIXTURE)), orc)
