"""Merlin transcript (STROBE-128 over Keccak-f[1600]) for the host-side mirror of the reference's
Fiat-Shamir layer (reference transcript.rs, random.rs; third-party merlin 3.0).  The transcript is host
logic in the reference too -- the GPU backend only needs the challenges it produces -- so this is a plain
Python restatement used by the Python host mirror and its tests.  Check value: merlin's published test
vector (tests/test_transcript.py)."""

import ctypes as _ctypes

_MASK = (1 << 64) - 1
_RC = [0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000, 0x000000000000808B,
       0x0000000080000001, 0x8000000080008081, 0x8000000000008009, 0x000000000000008A, 0x0000000000000088,
       0x0000000080008009, 0x000000008000000A, 0x000000008000808B, 0x800000000000008B, 0x8000000000008089,
       0x8000000000008003, 0x8000000000008002, 0x8000000000000080, 0x000000000000800A, 0x800000008000000A,
       0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008]
_RHO = [0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14]
R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001


def _rol(x, n):
    return ((x << n) | (x >> (64 - n))) & _MASK if n else x


def keccak_f(a):
    """In place on a list of 25 lanes (x + 5 y)."""
    for rc in _RC:
        c = [a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20] for x in range(5)]
        d = [c[(x + 4) % 5] ^ _rol(c[(x + 1) % 5], 1) for x in range(5)]
        a = [a[i] ^ d[i % 5] for i in range(25)]
        b = [0] * 25
        for x in range(5):
            for y in range(5):
                b[y + 5 * ((2 * x + 3 * y) % 5)] = _rol(a[x + 5 * y], _RHO[x + 5 * y])
        a = [b[i] ^ (~b[(i % 5 + 1) % 5 + 5 * (i // 5)] & _MASK & b[(i % 5 + 2) % 5 + 5 * (i // 5)]) for i in range(25)]
        a[0] ^= rc
    return a


_native = False


def _native_keccak():
    """sbn_keccak_f1600 from libsbn254.so when the library has been built; the Python permutation above otherwise (the
    transcript is host logic and needs no GPU)."""
    global _native
    if _native is False:
        try:
            from .lib import load_library
            fn = load_library().sbn_keccak_f1600
            fn.restype = None
            _native = fn
        except Exception:
            _native = None
    return _native


_R = 166
_FLAG_I, _FLAG_A, _FLAG_C, _FLAG_T, _FLAG_M, _FLAG_K = 1, 2, 4, 8, 16, 32


def _native_merlin():
    """libsbn254's host-side Merlin (csrc/host/merlin.hpp) when the library has been built."""
    try:
        from .lib import load_library
        lib = load_library()
        for f in (lib.sbn_merlin_init, lib.sbn_merlin_append, lib.sbn_merlin_append_many, lib.sbn_merlin_challenge):
            f.restype = None
        return lib
    except Exception:
        return None


class Transcript:
    """Merlin transcript.  With libsbn254 built, the STROBE state lives in a 203-byte buffer driven by the library's host
    code (a keyless-scale proof appends ~60 000 scalars); otherwise the pure-Python STROBE below runs (PyTranscript)."""

    def __new__(cls, label, native=True):
        if cls is Transcript and (not native or _native_merlin() is None):
            return object.__new__(PyTranscript)
        return object.__new__(cls)

    def __init__(self, label, native=True):
        self._lib = _native_merlin()
        self._st = _ctypes.create_string_buffer(208)
        self._c64 = _ctypes.create_string_buffer(64)          # challenge_scalar's bytes: one buffer, not one per call
        self._lib.sbn_merlin_init(self._st, label, _ctypes.c_size_t(len(label)))

    def append_message(self, label, message):
        message = bytes(message)
        self._lib.sbn_merlin_append(self._st, label, _ctypes.c_size_t(len(label)), message, _ctypes.c_size_t(len(message)))

    def challenge_bytes(self, label, n):
        out = _ctypes.create_string_buffer(n)
        self._lib.sbn_merlin_challenge(self._st, label, _ctypes.c_size_t(len(label)), out, _ctypes.c_size_t(n))
        return out.raw

    # ---- ProofTranscript (reference transcript.rs:37-80); scalars are canonical Python ints here
    def append_protocol_name(self, name):
        self.append_message(b"protocol-name", name)

    def append_scalar(self, label, s):
        self.append_message(label, int(s).to_bytes(32, "little"))          # scalar.rs:75-84

    def append_scalars(self, label, scalars):
        data = b"".join(int(s).to_bytes(32, "little") for s in scalars)
        self._lib.sbn_merlin_append_many(self._st, label, _ctypes.c_size_t(len(label)), data, _ctypes.c_size_t(32),
                                         _ctypes.c_size_t(len(data) // 32))

    def append_scalars_canonical(self, label, canon):
        """append_scalars for a uint64[n, 4] array of canonical little-endian values (no Python integers in between)."""
        data = canon.tobytes()
        self._lib.sbn_merlin_append_many(self._st, label, _ctypes.c_size_t(len(label)), data, _ctypes.c_size_t(32),
                                         _ctypes.c_size_t(len(data) // 32))

    def append_point(self, label, compressed32):
        self.append_message(label, compressed32)

    def append_points(self, label, xy, inf):
        """One append_point per affine Montgomery point (uint64[n, 8] + infinity bytes): the share loop of
        PolyCommitment::append_to_transcript (hyrax.rs:46-50), compressed (group.rs:135-140) by the library's host code."""
        import numpy as _np
        xy = _np.ascontiguousarray(xy, dtype=_np.uint64).reshape(-1, 8)
        inf = _np.ascontiguousarray(inf, dtype=_np.uint8).reshape(-1)
        self._lib.sbn_merlin_append_points(self._st, label, _ctypes.c_size_t(len(label)), xy.ctypes.data_as(_ctypes.c_void_p),
                                           inf.ctypes.data_as(_ctypes.c_void_p), _ctypes.c_size_t(xy.shape[0]))

    def challenge_scalar(self, label):
        self._lib.sbn_merlin_challenge(self._st, label, _ctypes.c_size_t(len(label)), self._c64, _ctypes.c_size_t(64))
        return int.from_bytes(self._c64.raw, "little") % R_MOD                     # transcript.rs:56-67

    def challenge_scalars(self, label, n):
        return [self.challenge_scalar(label) for _ in range(n)]

    def challenge_scalars_mont(self, label, n):
        """n challenge scalars as a Montgomery uint64[n, 4] array, in one library call (sbn_merlin_challenge_scalars)."""
        import numpy as _np
        out = _np.empty((n, 4), dtype=_np.uint64)
        f = getattr(self._lib, "sbn_merlin_challenge_scalars", None) if getattr(self, "_lib", None) is not None else None
        if f is None:
            from .hyrax import fr_vec_from_ints
            return fr_vec_from_ints(self.challenge_scalars(label, n))
        f.restype = None
        f(self._st, label, _ctypes.c_size_t(len(label)), _ctypes.c_size_t(n), out.ctypes.data_as(_ctypes.c_void_p))
        return out


class PyTranscript(Transcript):
    """The same transcript in pure Python (no library needed)."""

    def __init__(self, label, native=False):
        self.st = bytearray(200)
        self.st[0:6] = bytes([1, _R + 2, 1, 0, 1, 96])
        self.st[6:18] = b"STROBEv1.0.2"
        self._permute()
        self.pos = 0
        self.pos_begin = 0
        self.cur_flags = 0
        self._meta_ad(b"Merlin v1.0", False)
        self.append_message(b"dom-sep", label)

    # ---- STROBE
    def _permute(self):
        f = _native_keccak()
        if f is not None:           # libsbn254's host-side Keccak-f (a product-layer proof draws thousands of challenges)
            buf = (_ctypes.c_uint64 * 25).from_buffer(self.st)
            f(buf)
            return
        lanes = [int.from_bytes(self.st[8 * i: 8 * i + 8], "little") for i in range(25)]
        lanes = keccak_f(lanes)
        for i, l in enumerate(lanes):
            self.st[8 * i: 8 * i + 8] = l.to_bytes(8, "little")

    def _run_f(self):
        self.st[self.pos] ^= self.pos_begin
        self.st[self.pos + 1] ^= 0x04
        self.st[_R + 1] ^= 0x80
        self._permute()
        self.pos = 0
        self.pos_begin = 0

    def _absorb(self, data):
        for b in data:
            self.st[self.pos] ^= b
            self.pos += 1
            if self.pos == _R:
                self._run_f()

    def _squeeze(self, n):
        out = bytearray()
        for _ in range(n):
            out.append(self.st[self.pos])
            self.st[self.pos] = 0
            self.pos += 1
            if self.pos == _R:
                self._run_f()
        return bytes(out)

    def _begin_op(self, flags, more):
        if more:
            assert self.cur_flags == flags
            return
        old_begin = self.pos_begin
        self.pos_begin = self.pos + 1
        self.cur_flags = flags
        self._absorb(bytes([old_begin, flags]))
        if flags & (_FLAG_C | _FLAG_K) and self.pos != 0:
            self._run_f()

    def _meta_ad(self, data, more):
        self._begin_op(_FLAG_M | _FLAG_A, more)
        self._absorb(data)

    def _ad(self, data, more):
        self._begin_op(_FLAG_A, more)
        self._absorb(data)

    def _prf(self, n, more):
        self._begin_op(_FLAG_I | _FLAG_A | _FLAG_C, more)
        return self._squeeze(n)

    # ---- Merlin
    def append_message(self, label, message):
        self._meta_ad(label, False)
        self._meta_ad(len(message).to_bytes(4, "little"), True)
        self._ad(message, False)

    def challenge_bytes(self, label, n):
        self._meta_ad(label, False)
        self._meta_ad(n.to_bytes(4, "little"), True)
        return self._prf(n, False)

    # ---- ProofTranscript (reference transcript.rs:37-80); scalars are canonical Python ints here
    def append_protocol_name(self, name):
        self.append_message(b"protocol-name", name)

    def append_scalar(self, label, s):
        self.append_message(label, int(s).to_bytes(32, "little"))          # scalar.rs:75-84

    def append_scalars(self, label, scalars):
        for s in scalars:
            self.append_scalar(label, s)

    def append_scalar(self, label, s):
        self.append_message(label, int(s).to_bytes(32, "little"))

    def append_point(self, label, compressed32):
        self.append_message(label, compressed32)

    def append_scalars_canonical(self, label, canon):
        data = canon.tobytes()
        for i in range(len(data) // 32):
            self.append_message(label, data[32 * i: 32 * i + 32])

    def append_points(self, label, xy, inf):
        from .hyrax import GroupElement
        for p, i in zip(xy.reshape(-1, 8), inf.reshape(-1)):
            self.append_point(label, GroupElement(p, i).compress())

    def challenge_scalar(self, label):
        return int.from_bytes(self.challenge_bytes(label, 64), "little") % R_MOD   # transcript.rs:56-67

    def challenge_scalars(self, label, n):
        return [self.challenge_scalar(label) for _ in range(n)]


class RandomTape:
    """random.rs:10-31.  The reference seeds the tape from OsRng; here the seed scalar is injected so a proof
    can be reproduced (and compared byte for byte with the CPU prover in the tests)."""

    def __init__(self, name, seed_scalar):
        self.tape = Transcript(name)
        self.tape.append_scalar(b"init_randomness", seed_scalar)

    def random_scalar(self, label):
        return self.tape.challenge_scalar(label)

    def random_vector(self, label, n):
        return self.tape.challenge_scalars(label, n)

    def random_vector_mont(self, label, n):
        """random_vector as a Montgomery uint64[n, 4] array (one library call for the whole vector)."""
        return self.tape.challenge_scalars_mont(label, n)
