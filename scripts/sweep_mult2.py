"""Round-2 sweep of the tabulated-sum commit (cfg1, device-resident): streams x chunk rows x pairs per thread x register
target x prefetch x L2 fetch granularity.  Every configuration's commitments are compared with the first one's."""
import itertools, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from spartan_bn254_b200 import Context, synth

L, R = 1024, 1024
table_mb = int(sys.argv[1]) if len(sys.argv) > 1 else 36000
ctx = Context(0)
dev = torch.device("cuda", 0)
G, h = synth.distinct_generators(ctx, R)
ctx.set("mult_max_mb", table_mb)
bases = ctx.bases(G, h)
zs = [torch.from_numpy(synth.uniform_scalars(1 + i, L * R).view(np.int64)).to(dev) for i in range(6)]
dC = torch.empty((L, 8), dtype=torch.int64, device=dev); dinf = torch.empty((L,), dtype=torch.uint8, device=dev)
stream = torch.cuda.current_stream()
ref = None


def run(tag, n=20):
    global ref
    for i in range(3):
        ctx.hyrax_commit_device(bases, zs[i % 6].data_ptr(), L, R, 0, dC.data_ptr(), dinf.data_ptr(), stream=stream.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(n):
        ctx.hyrax_commit_device(bases, zs[i % 6].data_ptr(), L, R, 0, dC.data_ptr(), dinf.data_ptr(), stream=stream.cuda_stream)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    ctx.hyrax_commit_device(bases, zs[0].data_ptr(), L, R, 0, dC.data_ptr(), dinf.data_ptr(), stream=stream.cuda_stream)
    torch.cuda.synchronize()
    out = (dC.cpu().numpy().copy(), dinf.cpu().numpy().copy())
    ok = True
    if ref is None:
        ref = out
    else:
        ok = np.array_equal(ref[0], out[0]) and np.array_equal(ref[1], out[1])
    print(f"{tag}: {ms:.3f} ms/commit {L*R/ms/1e3:.1f} Mpts/s table={bases.mult_table()} same={ok}", flush=True)
    return ms


ctx.set("mult_layout", 1); run("layout 1: position-major, separate prefix passes")
ctx.set("mult_layout", 2); run("layout 2: fused rounds, minb=3")
ctx.set("ba_minb", 4); run("layout 2: fused rounds, minb=4")
ctx.set("ba_minb", 3)
for streams, chunk in ((1, 0), (2, 256), (4, 256), (2, 1024)):
    ctx.set("mult_streams", streams); ctx.set("chunk_rows", chunk)
    run(f"layout 2 streams={streams} chunk={chunk}")
ctx.set("mult_streams", 2); ctx.set("chunk_rows", 0)
for r in (5, 7):
    ctx.set("mult_rounds", r)
    run(f"layout 2 rounds={r}")
ctx.set("mult_rounds", 0)
for ab, what in ((1, "prefix<1>"), (4, "invert"), (8, "sum_rows"), (16, "round 1"), (32, "rounds 2-6")):
    ctx.set("ablate", ab)
    ref = None
    run(f"layout 2 WITHOUT {what}")
ctx.set("ablate", 0)
