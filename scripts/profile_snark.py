"""Host-side profile of one warm SNARK.prove (cProfile): where the Python mirror spends the time the kernels do not.
Usage: profile_snark.py [log2_constraints=20] [out=gpurun_out/snark_profile.txt]"""
import cProfile, io, os, pstats, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import bench_snark
from spartan_bn254_b200.r1csproof import SNARK
from spartan_bn254_b200.transcript import Transcript

k = int(sys.argv[1]) if len(sys.argv) > 1 else 20
outp = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "snark_profile.txt")
keep = {}
res = bench_snark.run(k, quiet=True, keep=keep, keep_instance=True)
inst, comm, decomm, vars_m, input_m, gens, ctx = keep["instance"]
pr = cProfile.Profile()
pr.enable()
SNARK.prove(inst, comm, decomm, vars_m, input_m, gens, Transcript(b"snark"), 1)
ctx.synchronize()
pr.disable()
s = io.StringIO()
st = pstats.Stats(pr, stream=s)
st.sort_stats("tottime").print_stats(45)
st.sort_stats("cumulative").print_stats(70)
os.makedirs(os.path.dirname(outp), exist_ok=True)
open(outp, "w").write(s.getvalue())
print(s.getvalue()[:6000])
