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
