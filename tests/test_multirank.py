"""CPU tests of the N > 1 host logic (gloo, world_size 2 and 3): row partition + all-gather of the commitment
vector reproduce the single-process result.  The per-rank commit is the oracle here (no GPU in this
container); on GPUs bench.py drives the same sharding with the CUDA commit and NCCL."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, L, R, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "oracle"))
    import torch.distributed as dist
    import oracle as orc
    from spartan_bn254_b200 import synth, parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    G, h = orc.multi_commit_gens(b"gens_r1cs_eval", R)
    Z = synth.derefs_scalars((L * R).bit_length() - 1)
    blinds = synth.uniform_scalars(4, L)
    Zl, first, n = parallel.local_slice(Z, L, R, world, rank)
    fn = lambda z, nl, r, b: orc.hyrax_commit(G, h, z, nl, r, b, threads=1)
    C, inf = parallel.commit_sharded(fn, Zl, n, L, R, blinds[first:first + n])
    np.save(os.path.join(out_dir, f"C{rank}.npy"), C)
    np.save(os.path.join(out_dir, f"inf{rank}.npy"), inf)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,L", [(2, 16), (3, 16)])
def test_sharded_commit_matches_single_process(orc, tmp_path, world, L):
    from spartan_bn254_b200 import synth
    R = 8
    port = _free_port()
    mp.spawn(_worker, args=(world, port, L, R, str(tmp_path)), nprocs=world, join=True)
    G, h = orc.multi_commit_gens(b"gens_r1cs_eval", R)
    Z = synth.derefs_scalars((L * R).bit_length() - 1)
    blinds = synth.uniform_scalars(4, L)
    C, inf = orc.hyrax_commit(G, h, Z, L, R, blinds)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"C{r}.npy"), C)
        assert np.array_equal(np.load(tmp_path / f"inf{r}.npy"), inf)


def test_shard_rows_partition():
    from spartan_bn254_b200.parallel import shard_rows
    for L in (1, 7, 8, 1024, 4097):
        for world in (1, 2, 3, 8):
            spans = [shard_rows(L, world, r) for r in range(world)]
            assert spans[0][0] == 0
            for (f0, n0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + n0 == f1
            assert spans[-1][0] + spans[-1][1] == L
            assert max(n for _, n in spans) - min(n for _, n in spans) <= 1


def _commit_inner_worker(rank, world, port, L, R, out_dir):
    """DensePolynomial.commit_inner(shard=...) -- the row-sharded R1CSProof::commit_poly -- with the resident polynomial
    replaced by a stand-in whose commit_rows is the oracle: what runs here is the host logic (block bounds, blinds slice,
    all-gather of the 65-byte row records over gloo, reassembly)."""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "oracle"))
    import torch.distributed as dist
    import oracle as orc
    from spartan_bn254_b200 import synth
    from spartan_bn254_b200.hyrax import DensePolynomial
    from spartan_bn254_b200.parallel import make_all_gather
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    G, h = orc.multi_commit_gens(b"gens_r1cs_sat", R)
    Z = synth.uniform_scalars(21, L * R)
    blinds = synth.uniform_scalars(22, L)

    class Gens:            # the two things commit_inner asks of a MultiCommitGens
        n = R

        @staticmethod
        def device_bases():
            return None

    class ResidentStandIn:
        def commit_rows(self, bases, first, n, R_size, bl):
            return orc.hyrax_commit(G, h, np.ascontiguousarray(Z.reshape(L, R, 4)[first:first + n]).reshape(-1, 4), n, R_size, bl, threads=1)

    poly = DensePolynomial(Z)
    poly._poly = ResidentStandIn()
    comm = poly.commit_inner(blinds, Gens, shard=(rank, world, make_all_gather()))
    np.save(os.path.join(out_dir, f"C{rank}.npy"), comm.C)
    np.save(os.path.join(out_dir, f"inf{rank}.npy"), comm.inf)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_commit_inner_host_logic(orc, tmp_path):
    from spartan_bn254_b200 import synth
    world, L, R = 2, 16, 8
    mp.spawn(_commit_inner_worker, args=(world, _free_port(), L, R, str(tmp_path)), nprocs=world, join=True)
    G, h = orc.multi_commit_gens(b"gens_r1cs_sat", R)
    C, inf = orc.hyrax_commit(G, h, synth.uniform_scalars(21, L * R), L, R, synth.uniform_scalars(22, L))
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"C{r}.npy"), C)
        assert np.array_equal(np.load(tmp_path / f"inf{r}.npy"), inf)
