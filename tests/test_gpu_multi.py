"""Multi-GPU paths on real devices (-m gpu).  One-process / k-context commit (sbn_hyrax_commit_multi) runs on a single
GPU too (two contexts on device 0); the two-device and NCCL tests skip on a one-GPU box (run them with gpurun --gpus 2)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("devices", [(0, 0), (0, 1), (0, 0, 0)])
def test_commit_multi_matches_single_context(orc, devices):
    """sbn_hyrax_commit_multi: k contexts, contiguous row blocks, one host thread per context."""
    from spartan_bn254_b200 import Context, synth
    from spartan_bn254_b200.lib import hyrax_commit_multi
    if max(devices) >= _ngpu():
        pytest.skip("needs %d GPUs" % (max(devices) + 1))
    L, R = 70, 64                     # 70 rows over 2 or 3 contexts: uneven blocks
    ctxs = [Context(d) for d in devices]
    try:
        G, h = synth.distinct_generators(ctxs[0], R)
        bases = [c.bases(G, h) for c in ctxs]
        Z = synth.uniform_scalars(5, L * R)
        Z[3 * R:4 * R] = 0
        blinds = synth.uniform_scalars(6, L)
        C, inf = hyrax_commit_multi(ctxs, bases, Z, L, R, blinds)
        C1, inf1 = ctxs[0].hyrax_commit(bases[0], Z, L, R, blinds)
        assert np.array_equal(C, C1) and np.array_equal(inf, inf1)
        Co, info = orc.hyrax_commit(G, h, Z, L, R, blinds)
        assert np.array_equal(C, Co) and np.array_equal(inf, info)
        # zero blinds, many rows (each block of >= 256 rows takes the tabulated-sum path of its own context)
        L2 = 600
        Z2 = synth.uniform_scalars(7, L2 * R)
        C2, inf2 = hyrax_commit_multi(ctxs, bases, Z2, L2, R, None)
        Co2, info2 = orc.hyrax_commit(G, h, Z2, L2, R, None, threads=0)
        assert np.array_equal(C2, Co2) and np.array_equal(inf2, info2)
        for b in bases:
            b.close()
    finally:
        for c in ctxs:
            c.close()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _nccl_worker(rank, world, port, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import torch
    import torch.distributed as dist
    from spartan_bn254_b200 import Context, synth
    from spartan_bn254_b200.hyrax import MultiCommitGens
    from spartan_bn254_b200.parallel import make_all_gather
    from spartan_bn254_b200.sparse_mlpoly import SparkAddresses
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    ctx = Context(rank)
    batch, N, nx, ny = 3, 4096, 10, 10
    rng = np.random.default_rng(5)
    row = rng.integers(0, 1 << nx, size=(batch, N), dtype=np.uint32)
    col = rng.integers(0, 1 << ny, size=(batch, N), dtype=np.uint32)
    spark = SparkAddresses(ctx, 1 << nx, row, col)
    used = 2 * batch * N
    ell = (used - 1).bit_length()
    R = 1 << (ell - ell // 2)
    gens = MultiCommitGens.new(R, b"gens_r1cs_eval", ctx)
    rx, ry = synth.uniform_scalars(21, nx), synth.uniform_scalars(22, ny)
    C, inf, poly = spark.derefs_commit(gens, rx, ry, shard=(rank, world, make_all_gather(dev)))
    np.save(os.path.join(out_dir, f"C{rank}.npy"), C)
    np.save(os.path.join(out_dir, f"inf{rank}.npy"), inf)
    if rank == 0:
        C1, inf1, poly1 = spark.derefs_commit(gens, rx, ry)
        np.save(os.path.join(out_dir, "C_single.npy"), C1)
        np.save(os.path.join(out_dir, "inf_single.npy"), inf1)
        poly1.close()
    # R1CSProof::commit_poly sharded the same way (sbn_poly_commit_rows): a resident 64 x 64 witness polynomial with blinds
    from spartan_bn254_b200.hyrax import DensePolynomial
    Lw = Rw = 64
    gw = MultiCommitGens.new(Rw, b"gens_r1cs_sat", ctx)
    Zw, bw = synth.uniform_scalars(31, Lw * Rw), synth.uniform_scalars(32, Lw)
    pw = DensePolynomial(Zw)
    pw.resident(ctx)
    cw = pw.commit_inner(bw, gw, shard=(rank, world, make_all_gather(dev)))
    np.save(os.path.join(out_dir, f"W{rank}.npy"), cw.C)
    np.save(os.path.join(out_dir, f"Winf{rank}.npy"), cw.inf)
    if rank == 0:
        c1 = pw.commit_inner(bw, gw)
        np.save(os.path.join(out_dir, "W_single.npy"), c1.C)
        np.save(os.path.join(out_dir, "Winf_single.npy"), c1.inf)
    dist.barrier()
    poly.close()
    spark.close()
    ctx.close()
    dist.destroy_process_group()


def test_nccl_sharded_derefs_commit_matches_one_gpu(tmp_path):
    """sbn_derefs_commit_rows on two GPUs + the NCCL all-gather of the row blocks == the one-GPU commitment, including the
    identity rows of the zero-padded last quarter (which no rank computes)."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_nccl_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    C1, inf1 = np.load(tmp_path / "C_single.npy"), np.load(tmp_path / "inf_single.npy")
    assert inf1[3 * len(inf1) // 4:].all() and not inf1[: 3 * len(inf1) // 4].any()
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"C{r}.npy"), C1)
        assert np.array_equal(np.load(tmp_path / f"inf{r}.npy"), inf1)
        assert np.array_equal(np.load(tmp_path / f"W{r}.npy"), np.load(tmp_path / "W_single.npy"))
        assert np.array_equal(np.load(tmp_path / f"Winf{r}.npy"), np.load(tmp_path / "Winf_single.npy"))


def test_poly_commit_rows_blocks_equal_the_whole(orc):
    """sbn_poly_commit_rows: three uneven row blocks of a resident polynomial (with blinds, and a block long enough for the
    tabulated-sum path) put side by side equal sbn_poly_commit and the oracle; out-of-range blocks are refused."""
    from spartan_bn254_b200 import Context, synth
    from spartan_bn254_b200.lib import Poly, SbnError
    ctx = Context(0)
    try:
        ctx.set("mult_max_mb", 2048)
        L, R = 700, 64
        G, h = synth.distinct_generators(ctx, R)
        bases = ctx.bases(G, h)
        Z, bl = synth.uniform_scalars(41, L * R), synth.uniform_scalars(42, L)
        poly = Poly(ctx, Z)
        C, inf = poly.commit(bases, L, R, bl)
        parts = [poly.commit_rows(bases, f, n, R, bl[f:f + n]) for f, n in ((0, 300), (300, 1), (301, 399))]
        assert np.array_equal(np.concatenate([p[0] for p in parts]), C)
        assert np.array_equal(np.concatenate([p[1] for p in parts]), inf)
        Co, info = orc.hyrax_commit(G, h, Z, L, R, bl, threads=0)
        assert np.array_equal(C, Co) and np.array_equal(inf, info)
        with pytest.raises(SbnError):
            poly.commit_rows(bases, 650, 51, R, None)
        poly.close()
        bases.close()
    finally:
        ctx.close()
