"""Do consecutive independent commits overlap when the caller alternates two streams?  (cfg1, device-resident)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from spartan_bn254_b200 import Context, synth

L, R = 1024, 1024
ctx = Context(0)
dev = torch.device("cuda", 0)
G, h = synth.distinct_generators(ctx, R)
ctx.set("mult_max_mb", 36000)
bases = ctx.bases(G, h)
zs = [torch.from_numpy(synth.uniform_scalars(1 + i, L * R).view(np.int64)).to(dev) for i in range(6)]
dC = [torch.empty((L, 8), dtype=torch.int64, device=dev) for _ in range(2)]
dinf = [torch.empty((L,), dtype=torch.uint8, device=dev) for _ in range(2)]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]


def run(tag, nstreams, n=40):
    ref = None
    for rep in range(2):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(torch.cuda.current_stream())
        for s in streams[:nstreams]:
            s.wait_event(e0)
        for i in range(n):
            k = i % nstreams
            ctx.hyrax_commit_device(bases, zs[i % 6].data_ptr(), L, R, 0, dC[k].data_ptr(), dinf[k].data_ptr(), stream=streams[k].cuda_stream)
        for s in streams[:nstreams]:
            torch.cuda.current_stream().wait_stream(s)
        e1.record(torch.cuda.current_stream())
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
    # correctness of the last two commits (i = n - 2, n - 1) against single-stream results
    outs = [(dC[k].cpu().numpy().copy(), dinf[k].cpu().numpy().copy()) for k in range(nstreams)]
    print(f"{tag}: {ms:.3f} ms/commit {L*R/ms/1e3:.1f} Mpts/s", flush=True)
    return outs


if len(sys.argv) > 1 and sys.argv[1] == "ablate":
    run("two caller streams, everything", 2)
    for ab, what in ((1, "prefix<1>"), (2, "prefix rounds >= 2"), (3, "all prefix"), (4, "invert"), (8, "sum_rows"), (16, "finish<1>"),
                     (32, "finish rounds >= 2"), (15, "all but finish + entries"), (63, "all but entries")):
        ctx.set("ablate", ab)
        run(f"two caller streams WITHOUT {what}", 2)
    ctx.set("ablate", 0)
    for layout in (2, 1):
        ctx.set("mult_layout", layout)
        run(f"two caller streams, layout {layout}", 2)
    for mb in (4, 3):
        ctx.set("ba_minb", mb)
        run(f"two caller streams, minb {mb}", 2)
    for r in (4, 5, 6):
        ctx.set("mult_rounds", r)
        run(f"two caller streams, rounds {r}", 2)
    ctx.set("mult_rounds", 0)
    for B in (8, 12, 16, 24):
        ctx.set("ba_batch", B)
        run(f"two caller streams, pairs per thread {B}", 2)
    sys.exit(0)
if len(sys.argv) > 1 and sys.argv[1] == "smem":
    run("two caller streams, default", 2)
    for pk, fk in ((57, 0), (75, 0), (110, 0), (0, 50), (0, 75), (57, 50), (75, 75), (110, 75)):
        ctx.set("prefix_smem_kb", pk); ctx.set("finish_smem_kb", fk)
        run(f"two caller streams, prefix smem {pk} KB, finish smem {fk} KB", 2)
    sys.exit(0)
one = run("one caller stream", 1)
two = run("two alternating caller streams", 2)
# reference results for inputs (n-2) % 6 and (n-1) % 6 on one stream
n = 40
exp = []
for i in (n - 2, n - 1):
    ctx.hyrax_commit_device(bases, zs[i % 6].data_ptr(), L, R, 0, dC[0].data_ptr(), dinf[0].data_ptr(), stream=streams[0].cuda_stream)
    torch.cuda.synchronize()
    exp.append((dC[0].cpu().numpy().copy(), dinf[0].cpu().numpy().copy()))
print("two-stream results equal the one-stream ones:", all(np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) for a, b in zip(two, exp)))
