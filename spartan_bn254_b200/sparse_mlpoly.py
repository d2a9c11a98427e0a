"""Host mirror of the pieces of the reference's Spark commitment (sparse_mlpoly_full.rs) that sit next to the GPU entry
points: the address / timestamp vectors fixed at encode time, the derefs commitment and the memory-checking layers.

  reference sparse_mlpoly_full.rs:204-258   AddrTimestamps::{new, deref}
  reference sparse_mlpoly_full.rs:292-304   Derefs::{new, commit}
  reference sparse_mlpoly_full.rs:731-841   ProductLayer, Layers::new
  reference sparse_mlpoly_full.rs:853-866   PolyEvalNetwork::new
"""
import numpy as np

from .parallel import shard_rows

from .lib import Addrs
from .product_tree import ProductCircuit


class AddrTimestamps:
    """Addresses of one side (uint32[batch, N]) with their read / audit timestamps.  The reference walks the operations
    one by one (:220-236); a stable sort gives the same counts: the read timestamp of an operation is the number of
    earlier operations (over all instances, in order) on the same cell."""

    def __init__(self, num_cells, ops_addr):
        ops = np.ascontiguousarray(ops_addr, dtype=np.uint32)
        assert ops.ndim == 2 and (ops < num_cells).all()
        flat = ops.reshape(-1).astype(np.int64)
        order = np.argsort(flat, kind="stable")
        sorted_addr = flat[order]
        start = np.flatnonzero(np.concatenate(([True], sorted_addr[1:] != sorted_addr[:-1])))
        group_start = np.repeat(start, np.diff(np.concatenate((start, [len(flat)]))))
        read = np.empty(len(flat), dtype=np.uint32)
        read[order] = (np.arange(len(flat)) - group_start).astype(np.uint32)
        self.num_cells = num_cells
        self.ops_addr = ops
        self.read_ts = read.reshape(ops.shape)
        self.audit_ts = np.bincount(flat, minlength=num_cells).astype(np.uint32)


class SparkAddresses:
    """Row and column AddrTimestamps of a multi-sparse-matrix commitment, resident on the GPU."""

    def __init__(self, ctx, num_cells, row_addr, col_addr):
        self.ctx = ctx
        self.row = AddrTimestamps(num_cells, row_addr)
        self.col = AddrTimestamps(num_cells, col_addr)
        self.gpu = Addrs(ctx, self.row.ops_addr, self.col.ops_addr)
        self.gpu.set_timestamps(self.row.read_ts, self.row.audit_ts, self.col.read_ts, self.col.audit_ts)
        self.batch = self.gpu.batch

    def derefs_commit(self, gens_n, rx, ry, shard=None):
        """dense.deref(mem_rx, mem_ry) + derefs.commit(gens_derefs) (:1720-1724): returns (C, inf, resident polynomial).
        shard = (rank, world, all_gather): this rank commits its contiguous block of rows and `all_gather(array)` returns the
        list of every rank's block (rows are independent, hyrax.rs:259-265; no other collective on the data path)."""
        if shard is None or shard[1] == 1:
            return self.gpu.derefs_commit(gens_n.device_bases(), rx, ry)
        rank, world, all_gather = shard
        L = self.gpu.derefs_rows()
        # Only the rows that hold values are divided: the merged polynomial is zero-padded to a power of two
        # (sparse_mlpoly_full.rs:295), so with three matrices the last quarter of the rows is all-zero and commits to the
        # identity -- contiguous blocks of ALL rows would hand the last ranks nothing to do.
        used = 2 * self.gpu.batch * self.gpu.N
        R = (1 << (max(0, (used - 1).bit_length()))) // L
        live = min(L, -(-used // R))
        first, n = shard_rows(live, world, rank)
        n_max = -(-live // world)
        C = np.zeros((n_max, 8), dtype=np.uint64)
        inf = np.ones(n_max, dtype=np.uint8)
        poly = None
        if n:
            Cb, infb, poly = self.gpu.derefs_commit(gens_n.device_bases(), rx, ry, rows=(first, n))
            C[:n], inf[:n] = Cb, infb
        Cs, infs = all_gather(C), all_gather(inf)
        C_all = np.zeros((L, 8), dtype=np.uint64)
        inf_all = np.ones(L, dtype=np.uint8)                 # rows past `live`: identity
        for r in range(world):
            f, m = shard_rows(live, world, r)
            C_all[f:f + m], inf_all[f:f + m] = Cs[r][:m], infs[r][:m]
        return C_all, inf_all, poly

    def close(self):
        self.gpu.close()


class ProductLayer:
    def __init__(self, circuits, batch):
        self.init = circuits[0]
        self.read_vec = circuits[1:1 + batch]
        self.write_vec = circuits[1 + batch:1 + 2 * batch]
        self.audit = circuits[1 + 2 * batch]

    def all(self):
        return [self.init] + self.read_vec + self.write_vec + [self.audit]


class Layers:
    """Layers::new (:800-841): hash layer + product circuits of one side, built on the device."""

    def __init__(self, spark, side, r, r_mem_check):
        gpu = spark.gpu.hashlayer(side, r, r_mem_check[0], r_mem_check[1])
        wrapped = []
        for g in gpu:
            pc = ProductCircuit.__new__(ProductCircuit)
            pc.gpu, pc.len, pc.num_layers = g, g.len, g.num_layers
            wrapped.append(pc)
        self.prod_layer = ProductLayer(wrapped, spark.batch)


class PolyEvalNetwork:
    """PolyEvalNetwork::new (:853-866).  rx / ry are the (equalised) evaluation points whose eq tables are the memories."""

    def __init__(self, spark, rx, ry, r_mem_check):
        self.row_layers = Layers(spark, 0, rx, r_mem_check)
        self.col_layers = Layers(spark, 1, ry, r_mem_check)
