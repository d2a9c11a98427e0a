"""CPU tests of the Python Merlin transcript used by the host mirror."""


def test_merlin_published_vector():
    from spartan_bn254_b200.transcript import Transcript
    t = Transcript(b"test protocol")
    t.append_message(b"some label", b"some data")
    assert t.challenge_bytes(b"challenge", 32).hex() == "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"


def test_matches_c_oracle_transcript(orc):
    from spartan_bn254_b200.transcript import Transcript, RandomTape
    a = Transcript(b"snark_proof")
    b = orc.Transcript(b"snark_proof")
    msg = bytes(range(256)) * 3                    # crosses the 166-byte STROBE rate several times
    a.append_message(b"poly_commitment_share", msg)
    b.append_message(b"poly_commitment_share", msg)
    for _ in range(3):
        assert a.challenge_bytes(b"c", 64) == b.challenge_bytes(b"c", 64)
    s = a.challenge_scalar(b"challenge_nextround")
    assert orc.from_mont(b.challenge_scalar(b"challenge_nextround")) == [s]
    tape = RandomTape(b"proof", 12345)
    ot = orc.Transcript(b"proof")
    ot.append_message(b"init_randomness", (12345).to_bytes(32, "little"))
    assert orc.from_mont(ot.challenge_scalar(b"d")) == [tape.random_scalar(b"d")]


def test_native_and_python_transcripts_agree():
    """The library's host-side Merlin (csrc/host/merlin.hpp) and the pure-Python STROBE produce the same challenges on a
    random sequence of operations."""
    import random
    from spartan_bn254_b200.transcript import Transcript, PyTranscript
    a, b = Transcript(b"agree"), Transcript(b"agree", native=False)
    assert isinstance(b, PyTranscript) and not isinstance(a, PyTranscript)
    rnd = random.Random(1)
    for _ in range(200):
        k = rnd.randrange(4)
        if k == 0:
            m = bytes(rnd.randrange(256) for _ in range(rnd.randrange(0, 400)))
            a.append_message(b"lab", m); b.append_message(b"lab", m)
        elif k == 1:
            v = [rnd.randrange(1 << 250) for _ in range(rnd.randrange(1, 9))]
            a.append_scalars(b"sc", v); b.append_scalars(b"sc", v)
        elif k == 2:
            assert a.challenge_scalar(b"c") == b.challenge_scalar(b"c")
        else:
            n = rnd.randrange(1, 200)
            assert a.challenge_bytes(b"cb", n) == b.challenge_bytes(b"cb", n)


def test_native_point_compression_and_append_points(orc):
    """sbn_g1_compress / sbn_merlin_append_points (host code of the library) against GroupElement::compress restated in
    Python (group.rs:135-140) and against the oracle's transcript fed point by point: multiples of G (both signs of y occur),
    the identity by flag and by the (0, 0) encoding."""
    import ctypes as C
    import numpy as np
    from spartan_bn254_b200.hyrax import GroupElement
    from spartan_bn254_b200.lib import load_library
    from spartan_bn254_b200.transcript import Transcript, PyTranscript
    lib = load_library()
    G, h = orc.multi_commit_gens(b"gens_r1cs_eval", 24)
    pts = np.concatenate([G, h.reshape(1, 8), np.zeros((2, 8), dtype=np.uint64)])
    neg = pts[:6].copy()                         # -P: y -> p - y flips the sign bit
    P = 21888242871839275222246405745257275088696311157297823662689037894645226208583
    for row in neg:
        y = sum(int(v) << (64 * i) for i, v in enumerate(row[4:]))
        ny = (P - y) % P
        row[4:] = [(ny >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]
    pts = np.ascontiguousarray(np.concatenate([pts, neg]))
    inf = np.zeros(pts.shape[0], dtype=np.uint8)
    inf[25] = 1                                  # flagged identity; row 26 is the (0, 0) encoding without a flag
    out = np.zeros((pts.shape[0], 32), dtype=np.uint8)
    assert lib.sbn_g1_compress(pts.ctypes.data_as(C.c_void_p), inf.ctypes.data_as(C.c_void_p), C.c_size_t(pts.shape[0]),
                               out.ctypes.data_as(C.c_void_p)) == 0
    exp = [GroupElement(p, i or not p.any()).compress_py() for p, i in zip(pts, inf)]
    assert [bytes(o) for o in out] == exp
    assert [GroupElement(p, i or not p.any()).compress() for p, i in zip(pts, inf)] == exp      # the single-point binding
    assert {e[31] >> 7 for e in exp} == {0, 1} and exp[25][31] == 0x40 and exp[26][31] == 0x40
    inf[26] = 1
    a, b, c = Transcript(b"pts"), PyTranscript(b"pts"), orc.Transcript(b"pts")
    a.append_points(b"poly_commitment_share", pts, inf)
    b.append_points(b"poly_commitment_share", pts, inf)
    for e in exp:
        c.append_message(b"poly_commitment_share", e)
    ch = a.challenge_bytes(b"c", 64)
    assert ch == b.challenge_bytes(b"c", 64) == c.challenge_bytes(b"c", 64)


def test_host_side_scalar_conversions_and_canonical_append(orc):
    """sbn_fr_{to,from}_canonical_host (the library's host code behind fr_vec_to_ints / fr_vec_from_ints for short vectors)
    against the Python big-int forms and the oracle's Montgomery conversion; append_scalars_canonical feeds the transcript
    the same bytes as append_scalars."""
    import random
    import numpy as np
    from spartan_bn254_b200.hyrax import (R_MOD, _to_canonical_host, fr_from_int, fr_to_int, fr_vec_from_ints, fr_vec_to_ints)
    from spartan_bn254_b200.transcript import Transcript, PyTranscript
    rnd = random.Random(9)
    vals = [0, 1, 2, R_MOD - 1, R_MOD - 2, 1 << 253] + [rnd.randrange(R_MOD) for _ in range(60)]
    m = fr_vec_from_ints(vals)                                   # 66 < 128 elements: host path of the library
    assert np.array_equal(m, orc.to_mont(vals))
    assert all(np.array_equal(m[i], fr_from_int(v)) for i, v in enumerate(vals))
    assert fr_vec_to_ints(m) == vals == [fr_to_int(x) for x in m]
    canon = _to_canonical_host(m)
    assert [int.from_bytes(canon[i].tobytes(), "little") for i in range(len(vals))] == vals
    a, b, c = Transcript(b"sc"), Transcript(b"sc"), PyTranscript(b"sc")
    a.append_scalars(b"a", vals)
    b.append_scalars_canonical(b"a", canon)
    c.append_scalars_canonical(b"a", canon)
    ch = a.challenge_bytes(b"c", 64)
    assert ch == b.challenge_bytes(b"c", 64) == c.challenge_bytes(b"c", 64)


def test_challenge_scalars_in_one_call_match_the_loop():
    """sbn_merlin_challenge_scalars (RandomTape::random_vector, random.rs:24-31, in one library call, Montgomery form out) against
    n separate challenge_scalar calls on a transcript in the same state -- and the states afterwards."""
    from spartan_bn254_b200.transcript import Transcript, RandomTape
    from spartan_bn254_b200.hyrax import fr_vec_to_ints
    a, b = Transcript(b"t"), Transcript(b"t")
    if not hasattr(a, "_lib"):
        pytest.skip("libsbn254 not built")
    for t in (a, b):
        t.append_scalar(b"x", 12345)
    want = a.challenge_scalars(b"poly_blinds", 37)
    got = fr_vec_to_ints(b.challenge_scalars_mont(b"poly_blinds", 37))
    assert got == want
    assert a.challenge_scalar(b"next") == b.challenge_scalar(b"next")
    ta, tb = RandomTape(b"tape", 7), RandomTape(b"tape", 7)
    assert fr_vec_to_ints(tb.random_vector_mont(b"v", 5)) == ta.random_vector(b"v", 5)
