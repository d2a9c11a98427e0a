"""GPU parity tests AT THE BENCHMARKED CONFIGURATIONS (-m gpu): the exact paths bench.py times -- cfg1 with the 70000 MiB
table budget (bench.py's default: 17-bit windows, 64.5 GB table) and with 36000 MiB (16-bit windows, 34 GB), the keyless derefs shape 4096 x 8192 with and without a table, the encode-time
8192 x 8192 commit of small scalars -- compared with the CPU oracle on >= 64 sampled rows each, through the device-pointer
C-ABI entry point bench.py calls (sbn_hyrax_commit_device).  Each test owns its context: the tables are tens of GB."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _commit_device(ctx, bases, Z, L, R):
    import torch
    dev = torch.device("cuda", 0)
    dZ = torch.from_numpy(np.ascontiguousarray(Z).view(np.int64)).to(dev)
    dC = torch.empty((L, 8), dtype=torch.int64, device=dev)
    dinf = torch.empty((L,), dtype=torch.uint8, device=dev)
    ctx.hyrax_commit_device(bases, dZ.data_ptr(), L, R, 0, dC.data_ptr(), dinf.data_ptr(), stream=0)
    torch.cuda.synchronize()
    return dC.cpu().numpy().view(np.uint64), dinf.cpu().numpy()


def _check_rows(orc, G, h, Z, L, R, C, inf, nrows=64, must_include=()):
    rows = sorted(set(int(x) for x in np.linspace(0, L - 1, nrows)) | set(must_include))
    zs = np.ascontiguousarray(Z.reshape(L, R, 4)[rows].reshape(-1, 4))
    C_ref, inf_ref = orc.hyrax_commit(G, h, zs, len(rows), R, None, threads=0)
    assert np.array_equal(C[rows], C_ref), "affine limbs differ on sampled rows"
    assert np.array_equal(inf[rows], inf_ref), "infinity flags differ on sampled rows"
    return rows


@pytest.mark.parametrize("gens_kind,table_mb,want_bits", [("distinct", 70000, 17), ("distinct", 36000, 16), ("ref", 36000, 16)])
def test_cfg1_at_bench_table_budget(orc, gens_kind, table_mb, want_bits):
    """bench.py's default: 1024 x 1024, --table-mb 70000 => 17-bit windows; 36000 => 16-bit windows (round-2 records)."""
    from spartan_bn254_b200 import Context, synth
    from spartan_bn254_b200.hyrax import MultiCommitGens
    ctx = Context(0)
    try:
        ctx.set("mult_max_mb", table_mb)
        L = R = 1024
        if gens_kind == "distinct":
            G, h = synth.distinct_generators(ctx, R)
        else:
            g = MultiCommitGens.new(R, b"gens_r1cs_eval", ctx)
            G, h = g.G, g.h
        bases = ctx.bases(G, h)
        Z = synth.uniform_scalars(11, L * R)
        Z[5 * R:6 * R] = 0                      # an all-zero row: identity commitment
        C, inf = _commit_device(ctx, bases, Z, L, R)
        bits, nbytes = bases.mult_table()
        # distinct generators: 17-bit windows (64.5 GB) / 16-bit (34 GB); the reference set merges to 333 distinct points and
        # affords 17 bits under either budget
        assert bits == want_bits if gens_kind == "distinct" else bits >= want_bits, "benchmarked window width: want %d, got %d" % (want_bits, bits)
        _check_rows(orc, G, h, Z, L, R, C, inf, 64, must_include=(5,))
        assert inf[5] == 1
        # the same commit through the host-pointer entry point (the e2e leg) and with the row-major layout of round 1
        C2, inf2 = ctx.hyrax_commit(bases, Z, L, R, None)
        assert np.array_equal(C2, C) and np.array_equal(inf2, inf)
        ctx.set("mult_layout", 0)
        C3, inf3 = _commit_device(ctx, bases, Z, L, R)
        assert np.array_equal(C3, C) and np.array_equal(inf3, inf)
        assert ctx.memory_stats()["mult_table_fallbacks"] == 0
        bases.close()
    finally:
        ctx.close()


@pytest.mark.parametrize("table_mb", [50000, 0])
def test_cfg2_keyless_derefs_shape(orc, table_mb):
    """BASELINE configs[2]: 4096 x 8192, derefs-style scalars (last quarter of the rows zero), with the 43 GB table
    (bench.py --workload cfg2_4096x8192 --table-mb 50000, and the strong-scaling leg) and through the bucket pipeline."""
    from spartan_bn254_b200 import Context, synth
    ctx = Context(0)
    try:
        ctx.set("mult_max_mb", table_mb)
        L, R = 4096, 8192
        G, h = synth.distinct_generators(ctx, R)
        bases = ctx.bases(G, h)
        Z = synth.derefs_scalars(25)
        C, inf = _commit_device(ctx, bases, Z, L, R)
        bits, _ = bases.mult_table()
        assert (bits > 0) == (table_mb > 0)
        _check_rows(orc, G, h, Z, L, R, C, inf, 64, must_include=(0, 3071, 3072, 4095))
        assert inf[3072:].all() and not inf[:3072].any()      # sparse_mlpoly_full.rs:295: the padding rows commit to the identity
        bases.close()
    finally:
        ctx.close()


def test_enc_8192x8192_small_scalars(orc):
    """comb_ops of the keyless encode (sparse_mlpoly_full.rs:155-196): 8192 x 8192 values below 2^21."""
    from spartan_bn254_b200 import Context, synth
    ctx = Context(0)
    try:
        L = R = 8192
        G, h = synth.distinct_generators(ctx, R)
        bases = ctx.bases(G, h)
        Z = ctx.fr_from_canonical(synth.small_scalars_canonical(8, L * R))
        C, inf = _commit_device(ctx, bases, Z, L, R)
        _check_rows(orc, G, h, Z, L, R, C, inf, 64)
        assert ctx.memory_stats()["small_scalar_commits"] == 1, "21-bit scalars must take the short window schedule"
        ctx.set("small_scalar_path", 0)                    # and the general path gives the same points
        C2, inf2 = _commit_device(ctx, bases, Z, L, R)
        assert np.array_equal(C, C2) and np.array_equal(inf, inf2)
        bases.close()
    finally:
        ctx.close()


@pytest.mark.parametrize("gens_kind", ["distinct", "ref"])
def test_mixed_small_and_full_rows(orc, gens_kind):
    """comb_ops-like layout (sparse_mlpoly_full.rs:176-196): runs of rows of small values (addresses, timestamps) around a
    run of full-size ones (matrix coefficients), plus an isolated small row and an all-zero row; the row classification must
    send each run through its own schedule and reproduce the oracle everywhere.  With the reference's generators two thirds of
    the columns merge into one point, whose scalars add up (the schedule must allow for the extra bits)."""
    from spartan_bn254_b200 import Context, synth
    from spartan_bn254_b200.hyrax import MultiCommitGens
    ctx = Context(0)
    try:
        L, R = 1024, 512
        if gens_kind == "distinct":
            G, h = synth.distinct_generators(ctx, R)
        else:
            g = MultiCommitGens.new(R, b"gens_r1cs_eval", ctx)
            G, h = g.G, g.h
        bases = ctx.bases(G, h)
        small = ctx.fr_from_canonical(synth.small_scalars_canonical(3, L * R)).reshape(L, R, 4)
        full = synth.uniform_scalars(4, L * R).reshape(L, R, 4)
        Z = small.copy()
        Z[400:700] = full[400:700]          # a run of full-size rows
        Z[100] = full[100]                  # one full-size row inside a small run
        Z[800] = 0
        Z[900, 7] = full[900, 7]            # a single large scalar makes its row large
        Z = np.ascontiguousarray(Z.reshape(L * R, 4))
        C, inf = _commit_device(ctx, bases, Z, L, R)
        assert ctx.memory_stats()["small_scalar_commits"] == 1
        rows = _check_rows(orc, G, h, Z, L, R, C, inf, 96, must_include=(99, 100, 101, 399, 400, 699, 700, 800, 899, 900, 901, 1023))
        assert inf[800] == 1 and len(rows) >= 96
        bases.close()
    finally:
        ctx.close()


@pytest.mark.parametrize("nstreams", [2, 3, 8])
def test_independent_commits_on_two_caller_streams(orc, nstreams):
    """bench.py's device-resident leg issues consecutive, independent commits on two alternating caller streams; the library
    takes its two workspace / stream sets in turn so that they overlap.  Eight commits of different inputs in flight must each
    equal the commit of the same input issued alone, and the oracle on sampled rows.  With three streams, or one stream per
    commit, consecutive users of a workspace set sit on DIFFERENT streams: the library's own completion events keep them apart."""
    import torch
    from spartan_bn254_b200 import Context, synth
    ctx = Context(0)
    try:
        ctx.set("mult_max_mb", 2048)
        L, R = 512, 128
        dev = torch.device("cuda", 0)
        G, h = synth.distinct_generators(ctx, R)
        bases = ctx.bases(G, h)
        Zs = [synth.uniform_scalars(40 + i, L * R) for i in range(8)]
        dZ = [torch.from_numpy(z.view(np.int64)).to(dev) for z in Zs]
        dC = [torch.empty((L, 8), dtype=torch.int64, device=dev) for _ in range(8)]
        dinf = [torch.empty((L,), dtype=torch.uint8, device=dev) for _ in range(8)]
        streams = [torch.cuda.Stream(device=dev) for _ in range(nstreams)]
        torch.cuda.synchronize()
        for i in range(8):
            ctx.hyrax_commit_device(bases, dZ[i].data_ptr(), L, R, 0, dC[i].data_ptr(), dinf[i].data_ptr(),
                                    stream=streams[i % nstreams].cuda_stream)
        torch.cuda.synchronize()
        assert bases.mult_table()[0] > 0
        for i in range(8):
            C = dC[i].cpu().numpy().view(np.uint64)
            inf = dinf[i].cpu().numpy()
            C1, inf1 = _commit_device(ctx, bases, Zs[i], L, R)
            assert np.array_equal(C, C1) and np.array_equal(inf, inf1), "commit %d in flight differs from the same commit alone" % i
            if i in (0, 7):
                _check_rows(orc, G, h, Zs[i], L, R, C, inf, 16)
        bases.close()
    finally:
        ctx.close()


@pytest.mark.parametrize("nstreams", [2, 3])
def test_async_host_commits_on_two_streams(orc, nstreams):
    """sbn_hyrax_commit_async (bench.py's e2e leg: three streams taken in turn): six host-pointer commits from pinned buffers
    issued on two alternating streams and on three (three staging sets, two workspace sets, each handed on by an event),
    with and without blinds, over a generator set with a digit-multiple table and over one without (which completes inside the
    call); every result equals the synchronous sbn_hyrax_commit of the same input."""
    import torch
    from spartan_bn254_b200 import Context, synth
    ctx = Context(0)
    try:
        L, R = 512, 128
        dev = torch.device("cuda", 0)
        G, h = synth.distinct_generators(ctx, R)
        streams = [torch.cuda.Stream(device=dev) for _ in range(nstreams)]
        for table_mb in (2048, 0):
            ctx.set("mult_max_mb", table_mb)
            bases = ctx.bases(G, h)
            Zs = [synth.uniform_scalars(60 + i, L * R) for i in range(6)]
            bl = synth.uniform_scalars(70, L)
            pin = [torch.from_numpy(z.view(np.int64)).pin_memory() for z in Zs]
            pbl = torch.from_numpy(bl.view(np.int64)).pin_memory()
            outC = [torch.empty((L, 8), dtype=torch.int64).pin_memory() for _ in range(6)]
            outI = [torch.empty((L,), dtype=torch.uint8).pin_memory() for _ in range(6)]
            for i in range(6):
                ctx.hyrax_commit_raw_async(bases, pin[i].data_ptr(), L, R, pbl.data_ptr() if i % 3 == 2 else 0, outC[i].data_ptr(),
                                           outI[i].data_ptr(), streams[i % nstreams].cuda_stream)
            torch.cuda.synchronize()
            for i in range(6):
                C, inf = ctx.hyrax_commit(bases, Zs[i], L, R, bl if i % 3 == 2 else None)
                assert np.array_equal(outC[i].numpy().view(np.uint64), C) and np.array_equal(outI[i].numpy(), inf), (table_mb, i)
            _check_rows(orc, G, h, Zs[0], L, R, outC[0].numpy().view(np.uint64), outI[0].numpy(), 16)
            bases.close()
    finally:
        ctx.close()


def test_async_host_commits_from_pageable_memory(orc):
    """sbn_hyrax_commit_async with PAGEABLE inputs and outputs (a Rust Vec<Scalar> / Vec<G1Affine>): the scalars go through
    the pinned ring of the workspace set, the results through a pinned landing buffer and a stream-ordered host function.
    Nine commits on three streams, the ring slots and landing buffers reused three times over; every result equals the
    blocking call's."""
    import torch
    from spartan_bn254_b200 import Context, synth
    ctx = Context(0)
    try:
        ctx.set("mult_max_mb", 2048)
        L, R = 512, 128
        dev = torch.device("cuda", 0)
        G, h = synth.distinct_generators(ctx, R)
        bases = ctx.bases(G, h)
        # streams of the library's own (sbn_stream_create / _synchronize / _destroy: what a Rust caller without CUDA bindings uses)
        lib_streams = [ctx.stream_create() for _ in range(3)]

        class _S:
            def __init__(self, h):
                self.cuda_stream = h
        streams = [_S(h) for h in lib_streams]
        Zs = [synth.uniform_scalars(80 + i, L * R) for i in range(9)]
        outC = [np.zeros((L, 8), dtype=np.uint64) for _ in range(9)]
        outI = [np.full((L,), 7, dtype=np.uint8) for _ in range(9)]
        for i in range(9):
            ctx.hyrax_commit_raw_async(bases, Zs[i].ctypes.data, L, R, 0, outC[i].ctypes.data, outI[i].ctypes.data,
                                       streams[i % 3].cuda_stream)
        for h in lib_streams:
            ctx.stream_synchronize(h)
        assert bases.mult_table()[0] > 0
        for i in range(9):
            C, inf = ctx.hyrax_commit(bases, Zs[i], L, R, None)
            assert np.array_equal(outC[i], C) and np.array_equal(outI[i], inf), i
        for h in lib_streams:
            ctx.stream_destroy(h)
        _check_rows(orc, G, h, Zs[8], L, R, outC[8], outI[8], 16)
        bases.close()
    finally:
        ctx.close()
