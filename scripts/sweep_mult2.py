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


for rounds in (6, 5, 4, 3):
    for wpr in (1, 2, 4):
        ctx.set("mult_rounds", rounds); ctx.set("sum_wpr", wpr)
        run(f"rounds={rounds} warps per row={wpr}")
ctx.set("mult_rounds", 0); ctx.set("sum_wpr", 0)
run("auto")
