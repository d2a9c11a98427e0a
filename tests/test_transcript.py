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
