"""Device-resident sweep: wall time (CUDA events) of hyrax_commit_device vs chunk_rows / leaf_m / window."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from spartan_bn254_b200 import Context, synth

L = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
R = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
cs = [0] if len(sys.argv) < 4 else [int(x) for x in sys.argv[3].split(",")]
chunks = [L, L // 2, L // 4, L // 8] if len(sys.argv) < 5 else [int(x) for x in sys.argv[4].split(",")]
ms_ = [16, 32] if len(sys.argv) < 6 else [int(x) for x in sys.argv[5].split(",")]
ctx = Context(0)
dev = torch.device("cuda", 0)
G, h = synth.distinct_generators(ctx, R)
zs = [torch.from_numpy(synth.uniform_scalars(1 + i, L * R).view(np.int64)).to(dev) for i in range(4)]
dC = torch.empty((L, 8), dtype=torch.int64, device=dev); dinf = torch.empty((L,), dtype=torch.uint8, device=dev)
stream = torch.cuda.current_stream()
for c in cs:
    ctx.set("window_bits", c)
    bases = ctx.bases(G, h)
    for chunk in chunks:
        for m in ms_:
            ctx.set("chunk_rows", max(1, chunk)); ctx.set("leaf_m", m)
            for i in range(3):
                ctx.hyrax_commit_device(bases, zs[i % 4].data_ptr(), L, R, 0, dC.data_ptr(), dinf.data_ptr(), stream=stream.cuda_stream)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 10
            e0.record(stream)
            for i in range(n):
                ctx.hyrax_commit_device(bases, zs[i % 4].data_ptr(), L, R, 0, dC.data_ptr(), dinf.data_ptr(), stream=stream.cuda_stream)
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            print(f"c={bases.window_bits} chunk={chunk} m={m}: {ms:.3f} ms/commit  {L*R/ms/1e3:.1f} Mpts/s", flush=True)
