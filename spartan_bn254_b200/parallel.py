"""Row sharding of a Hyrax commit across ranks (one process per GPU).

Row commitments are independent (reference hyrax.rs:259-265 maps rows in parallel), so rank g commits the
contiguous row block [g*L/k, (g+1)*L/k) of the row-major evaluation vector and the only exchange is an
all-gather of the commitment vector (64 B + 1 flag byte per row) -- NCCL on GPUs, gloo in the CPU tests."""
import numpy as np


def shard_rows(L_size, world, rank):
    """-> (first_row, n_rows) of this rank; the first L % world ranks take one extra row."""
    base, extra = divmod(L_size, world)
    n = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, n


def local_slice(Z, L_size, R_size, world, rank):
    """Rows of this rank out of the full row-major Z (uint64[L*R, 4])."""
    first, n = shard_rows(L_size, world, rank)
    Z = np.ascontiguousarray(Z, dtype=np.uint64).reshape(L_size, R_size, 4)
    return Z[first:first + n].reshape(n * R_size, 4), first, n


def commit_sharded(commit_fn, Z_local, n_local, L_size, R_size, blinds_local=None, group=None, device=None):
    """Commits this rank's rows with `commit_fn(Z_local, n_local, R_size, blinds_local) -> (C, inf)` and
    all-gathers the per-row results; returns (C uint64[L,8], inf uint8[L]) identical on every rank."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if n_local:
        C, inf = commit_fn(Z_local, n_local, R_size, blinds_local)
    else:
        C, inf = np.zeros((0, 8), dtype=np.uint64), np.zeros(0, dtype=np.uint8)
    if world == 1:
        return C, inf
    max_rows = -(-L_size // world)
    buf = np.zeros((max_rows, 9), dtype=np.int64)            # 8 limbs + flag, padded to the largest shard
    buf[:n_local, :8] = C.view(np.int64)
    buf[:n_local, 8] = inf
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    Cs, infs = [], []
    for r in range(world):
        _, n = shard_rows(L_size, world, r)
        a = out[r].cpu().numpy()
        Cs.append(a[:n, :8].view(np.uint64))
        infs.append(a[:n, 8].astype(np.uint8))
    return np.concatenate(Cs), np.concatenate(infs)


def make_all_gather(device=None, group=None):
    """-> all_gather(array) returning the list of every rank's (equal-shaped) numpy array: the only exchange of a sharded
    prove (the commitment row blocks).  NCCL when `device` is a CUDA device, gloo on CPU."""
    import torch
    import torch.distributed as dist

    def all_gather(arr):
        a = np.ascontiguousarray(arr)
        t = torch.from_numpy(a.view(np.uint8).reshape(-1).copy())
        if device is not None:
            t = t.to(device)
        out = [torch.empty_like(t) for _ in range(dist.get_world_size(group))]
        dist.all_gather(out, t, group=group)
        return [o.cpu().numpy().view(a.dtype).reshape(a.shape) for o in out]

    return all_gather
