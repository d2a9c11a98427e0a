#!/usr/bin/env python
"""Benchmark of the Hyrax row-commitment MSM hot path (BASELINE.json metric: BN254 Hyrax-commit MSM
G1 points/s).

One step = one Hyrax commit (DensePolynomial::commit_inner, reference hyrax.rs:253-267) of a synthetic
L x R polynomial per GPU over resident generators.  Default workload = BASELINE.json configs[1]: 2^20
scalars as 1024 rows x 1024 generators on one B200; with N GPUs every rank commits its own 1024-row block
of an (N*1024) x 1024 polynomial (rows are independent -- weak scaling, no data-path collective; the
commitment vector is all-gathered over NCCL at the end of every step, reference hyrax.rs:259-265 collects
the rows the same way).

The generators' digit-multiple table (mult_kernels.cuh: every d * 2^(kc) * G_j, built once per generator set and kept in
HBM like the bases themselves) is sized by --table-mb (default 70000 MiB: c = 17 at 1025 generators, 64.5 GB; 36000 gives
c = 16 / 34 GB, the library default 6144 c = 13 / 5.4 GB -- reported as value_default_budget); --table-mb 0 times the bucket
pipeline instead.  The headline table is released before the strong-scaling and prove legs, which run under 36000 MiB.

  value  points/s with scalars already resident in HBM (device-pointer C-ABI entry point)
  e2e    the same metric through sbn_hyrax_commit with PINNED HOST buffers: H2D of the scalars and D2H
         of the commitments inside the timed region
  --impl reference : the CPU restatement of the reference's path (oracle port; the Rust reference cannot be
         built offline) on all host threads, bounded sample per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "bn254_hyrax_commit_msm_points_per_s"
UNIT = "points/s"
WORKLOADS = {
    # name: (rows per GPU, generators)
    "cfg1_1024x1024": (1024, 1024),
    "cfg2_4096x4096": (4096, 4096),
    "cfg2_4096x8192": (4096, 8192),
    "enc_8192x8192": (8192, 8192),      # comb_ops of the keyless encode (sparse_mlpoly_full.rs:183); use with --scalars small
}
A_ADDS_PER_POINT = {1024: 26.0, 2048: 24.0, 4096: 22.0, 8192: 21.0}
IMAD_PER_FQMUL = 264       # 8x8 (lo+hi) products + Montgomery reduction, 32-bit limbs
FQMUL_PER_MIXED_ADD = 10   # XYZZ madd-2008-s: 8M + 2S


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg1_1024x1024", choices=sorted(WORKLOADS))
    ap.add_argument("--gens", default="distinct", choices=["distinct", "ref"],
                    help="distinct: k_j*G random (throughput headline); ref: the reference's degenerate MultiCommitGens")
    ap.add_argument("--scalars", default="uniform", choices=["uniform", "derefs", "small"],
                    help="uniform mod r | derefs-style gathers with zero rows | small: 21-bit values (comb_ops-like)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: every GPU commits the workload's rows (default); strong: the workload's rows are divided "
                         "across the GPUs (BASELINE configs[2]: the keyless derefs commitment sharded by rows)")
    ap.add_argument("--table-mb", type=int, default=70000,
                    help="budget (MiB) of the digit-multiple table of the commit's generator set (mult_kernels.cuh): the widest "
                         "window whose table fits is tabulated once and kept resident; 0 = bucket pipeline only; the "
                         "library's own default is 6144")
    ap.add_argument("--caller-streams", type=int, default=2, choices=[1, 2],
                    help="consecutive (independent) commits of the device-resident leg are issued on this many alternating caller "
                         "streams: with 2 the tail of one commit runs under the head of the next (sbn_hyrax_commit_device is asynchronous)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong", action="store_true",
                    help="skip the strong-scaling leg (BASELINE configs[2]: 4096 x 8192 derefs-shaped commit divided by rows across the GPUs)")
    ap.add_argument("--no-parity", action="store_true", help="skip the post-timing check of sampled commitments against the oracle")
    ap.add_argument("--no-prove", action="store_true",
                    help="skip the second half of BASELINE.json's metric: the keyless-shaped end-to-end prove time")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        for ts, line in self.samples:
            if ts < t0 - 0.05 or ts > t1 + 0.15:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args, rank, world):
    """CPU restatement of the reference path (oracle port) on the host cores -- rank 0 only."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import oracle as orc
    from spartan_bn254_b200 import synth
    orc.build()
    rows_per_gpu, R = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    G, h = orc.multi_commit_gens(b"bench-gens", R)              # any valid generators: cost is scalar-driven
    Z = synth.uniform_scalars(1, rows_per_gpu * R)
    # one untimed pass over `cores` rows sizes the step: every row of the workload when K steps of that fit ~150 s,
    # otherwise the largest whole multiple of the thread count that does (stated in `sample`)
    t0 = time.perf_counter()
    orc.hyrax_commit(G, h, Z[: cores * R], cores, R, None, threads=0)
    t_row = time.perf_counter() - t0                            # seconds per batch of `cores` rows (one per thread)
    per_full = t_row * rows_per_gpu / cores
    sample_rows = rows_per_gpu
    if per_full * (args.steps + min(args.warmup, 1)) > 150.0:
        sample_rows = max(cores, int(150.0 / (args.steps + 1) / t_row) * cores)
        sample_rows = min(sample_rows, rows_per_gpu)
    Z = Z[: sample_rows * R]
    for _ in range(min(args.warmup, 1)):
        orc.hyrax_commit(G, h, Z, sample_rows, R, None, threads=0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.hyrax_commit(G, h, Z, sample_rows, R, None, threads=0)
    dt = time.perf_counter() - t0
    value = args.steps * sample_rows * R / dt
    sample = (f"{sample_rows} of {rows_per_gpu} rows x {R} generators per step" +
              (" (every row of the workload)" if sample_rows == rows_per_gpu else " (bounded sample)") +
              ", uniform scalars, zero blinds")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32x8 Montgomery (CPU: u64x4)", "data": "synthetic",
        "config": {"workload": args.workload, "rows_per_gpu": rows_per_gpu, "generators": R,
                   "note": "CPU restatement of hyrax.rs:253-267 -> commitments.rs:144-154 -> signed-window Pippenger; "
                           "the Rust reference cannot be built offline (no cargo, arkworks not vendored)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "points_per_s_per_thread": value / cores,
                         "reference_published_points_per_s_per_thread": 2.0e5,
                         "published_note": "arkworks on ONE M2-Max thread (BASELINE.md 1: 166.2 s for the keyless derefs commitment); "
                                           "the C port is slower per thread than arkworks -- quote any GPU/CPU ratio beside both"},
        "same_config": sample_rows == rows_per_gpu,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
def cpu_baseline(R, rows_per_gpu):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as orc
    from spartan_bn254_b200 import synth
    orc.build()
    cores = os.cpu_count() or 1
    G, h = orc.multi_commit_gens(b"bench-gens", R)
    batch = max(cores, 64)
    Z = synth.uniform_scalars(1, batch * R)
    done, t0 = 0, time.perf_counter()
    while True:
        orc.hyrax_commit(G, h, Z, batch, R, None, threads=0)
        done += batch
        dt = time.perf_counter() - t0
        if dt > 10.0 or done >= 4 * rows_per_gpu:
            break
    return {"value": done * R / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{done} rows x {R} generators in {dt:.1f} s on {cores} host threads (oracle/bn254_oracle.c, "
                      f"threads across rows as rayon does at hyrax.rs:259)"}


def cpu_prove_baseline(gpu_prove):
    """Same-host CPU figure for the end-to-end prove (BASELINE configs[4]): oracle/bn254_oracle.c orc_prove_workload runs the
    table-sized phases of SNARK::prove at the keyless shape on every host thread -- the reference's algorithms and operation
    counts with real field / group arithmetic, rounds chained through a Merlin transcript; a restatement of the WORK, not a
    verifying prover (the Rust reference cannot be built here).  Also composes the "hooks only" estimate: the phases the
    SURVEY 8(b) boundary replaces (commit_inner, the openings' bound / MSMs / bullet reduction) at their GPU times, every
    other phase at its CPU time -- what the drop-in delivers under an otherwise unchanged CPU prover."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as orc
    orc.build()
    cores = os.cpu_count() or 1
    log_cons = 20
    try:
        avail_gb = [int(l.split()[1]) for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0] / 1e6
    except Exception:
        avail_gb = 64.0
    if avail_gb < 24:
        log_cons = 18                                   # ~8 GiB of tables at 2^20 constraints
    t0 = time.perf_counter()
    r = orc.prove_workload(log_cons, threads=0, derefs_rows=0)
    wall = time.perf_counter() - t0
    ph = r["phases"]
    out = {"seconds": r["seconds"], "unit": "s", "cores": cores, "kind": "port", "log2_constraints": log_cons,
           "phases_s": {k: round(v, 4) for k, v in ph.items()},
           "generator_derivation_s": r["gens_seconds"], "wall_s_with_setup": wall,
           "sample": f"every phase at 2^{log_cons} constraints (nnz padded to 2^{log_cons + 2}), every derefs row, once, on {cores} host threads",
           "note": "CPU restatement of the WORK of SNARK::prove (snark.rs:428-484): same algorithms, operation counts and round-to-round "
                   "dependencies, synthetic tables; it does not assemble or verify a proof"}
    if log_cons == 20 and gpu_prove.get("phases_ms"):
        g = gpu_prove["phases_ms"]
        try:
            gpu_hooks = (g["sat"]["witness_commit_ms"] + g["sat"]["witness_opening_ms"] + g["eval"]["eq_tables+derefs+derefs_commitment_ms"]) / 1e3
            # the three hash-layer openings run with the evaluations in one GPU phase: charge the whole phase
            gpu_hooks += g["eval"]["network_proof.hash_layer(evaluations + 3 openings)_ms"] / 1e3
            h2d = (1 << 30) / 25e9 + (1 << 25) / 25e9      # derefs Z (1 GiB) and the witness (32 MiB) cross PCIe from a CPU prover
            cpu_rest = sum(v for k, v in ph.items() if k not in ("witness_commit", "witness_opening", "derefs_commit", "hash_layer_openings"))
            out["hooks_only_estimate_s"] = cpu_rest + gpu_hooks + h2d
            out["hooks_only_note"] = ("COMPOSED, not run: CPU time of the phases the 8(b) hooks do not touch (%.2f s) + GPU time of the commit / "
                                      "opening phases (%.3f s) + the H2D copy of their inputs at 25 GB/s (%.3f s)" % (cpu_rest, gpu_hooks, h2d))
        except KeyError:
            pass
    return out


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from spartan_bn254_b200 import Context, synth
    from spartan_bn254_b200.hyrax import MultiCommitGens

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    ctx = Context(local_rank)
    L, R = WORKLOADS[args.workload]
    if args.scaling == "strong":
        if L % world:
            raise SystemExit("--scaling strong needs the row count to divide by the number of GPUs")
        L //= world             # contiguous row block of this rank (hyrax.rs:259-265: rows are independent)

    # ---- generators (resident for the whole run) and synthetic scalars
    if args.gens == "distinct":
        G, h = synth.distinct_generators(ctx, R)
    else:
        g = MultiCommitGens.new(R, b"gens_r1cs_eval", ctx)
        G, h = g.G, g.h
    ctx.set("mult_max_mb", args.table_mb)
    bases = ctx.bases(G, h)
    nbuf = max(2, -(-(160 << 20) // (L * R * 32)) + 1)      # rotate inputs: total > 126 MiB L2
    nbuf = min(nbuf, 8)
    if L * R * 32 > (1 << 30):
        nbuf = 1                                            # one 2 GiB input already exceeds L2 many times over
    host_bufs, dev_bufs, page_bufs = [], [], []
    for i in range(nbuf):
        seed = 1 + 131 * rank + i
        if args.scalars == "uniform":
            z = synth.uniform_scalars(seed, L * R)
        elif args.scalars == "small":
            z = ctx.fr_from_canonical(synth.small_scalars_canonical(seed, L * R))
        else:
            z = synth.derefs_scalars((L * R).bit_length() - 1, seed_table=2 + seed, seed_addr=3 + seed)
        page_bufs.append(z)                                 # pageable numpy memory: what a Rust Vec<Scalar> is
        t = torch.from_numpy(z.view(np.int64)).pin_memory()
        host_bufs.append(t)
        dev_bufs.append(t.to(dev, non_blocking=False))
    # two output buffers: with N > 1 the all-gather of step i runs on NCCL's stream underneath the commit of step i + 1
    dCs = [torch.empty((L, 8), dtype=torch.int64, device=dev) for _ in range(2)]
    dinfs = [torch.empty((L,), dtype=torch.uint8, device=dev) for _ in range(2)]
    dC, dinf = dCs[0], dinfs[0]
    gathers = [torch.empty((world * L, 8), dtype=torch.int64, device=dev) for _ in range(2)] if world > 1 else None
    pending = [None, None]
    hC = torch.empty((L, 8), dtype=torch.int64).pin_memory()
    hinf = torch.empty((L,), dtype=torch.uint8).pin_memory()
    stream = torch.cuda.current_stream()
    # Caller streams of the device-resident leg.  Step i runs on cstreams[i % n] with output buffer i & 1: the commits are
    # independent, so with two streams the library (two workspace sets, taken in turn) overlaps one commit's tail with the next
    # one's head.  The timed events sit on `stream`, which every caller stream forks from and joins back into.
    cstreams = [torch.cuda.Stream(device=dev) for _ in range(args.caller_streams)] if args.caller_streams > 1 else [stream]

    def fork():
        if len(cstreams) > 1:
            for cs in cstreams:
                cs.wait_stream(stream)

    def step_device(i, b=None):
        k = i & 1
        cs = cstreams[i % len(cstreams)]
        with torch.cuda.stream(cs):
            if pending[k] is not None:          # the gather of step i - 2 still reads this output buffer
                pending[k].wait()
                pending[k] = None
            ctx.hyrax_commit_device(b or bases, dev_bufs[i % nbuf].data_ptr(), L, R, 0, dCs[k].data_ptr(), dinfs[k].data_ptr(),
                                    stream=cs.cuda_stream)
            if world > 1:
                pending[k] = dist.all_gather_into_tensor(gathers[k], dCs[k], async_op=True)     # one NCCL kernel, no per-rank copies

    def drain():
        for k in range(2):
            if pending[k] is not None:
                with torch.cuda.stream(cstreams[k % len(cstreams)]):
                    pending[k].wait()        # the caller stream waits for the collective: it is inside the timed region
                pending[k] = None
        if len(cstreams) > 1:
            for cs in cstreams:
                stream.wait_stream(cs)

    def step_e2e(i):
        ctx.hyrax_commit_raw(bases, host_bufs[i % nbuf].data_ptr(), L, R, 0, hC.data_ptr(), hinf.data_ptr())

    # the same call, asynchronous, on three caller streams taken in turn: the library has three staging sets and two workspace
    # sets, so the H2D copy of step i + 2 runs while steps i and i + 1 compute (each set is handed on by a device-side event)
    NE = 3
    estreams = [torch.cuda.Stream(device=dev) for _ in range(NE)]
    hCs = [torch.empty((L, 8), dtype=torch.int64).pin_memory() for _ in range(NE)]
    hinfs = [torch.empty((L,), dtype=torch.uint8).pin_memory() for _ in range(NE)]

    def step_e2e_async(i):
        k = i % NE
        ctx.hyrax_commit_raw_async(bases, host_bufs[i % nbuf].data_ptr(), L, R, 0, hCs[k].data_ptr(), hinfs[k].data_ptr(),
                                   estreams[k].cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_device(steps, warm, b=None):
        for i in range(warm):
            step_device(i, b)
        drain()
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        fork()
        for i in range(steps):
            step_device(warm + i, b)
        drain()
        a1.record(stream)
        barrier()
        tt = torch.tensor([a0.elapsed_time(a1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    # ---- integer roofline denominator, measured live (MEASURED_PEAKS.json has no integer-pipe figure)
    peak_imad = ctx.microbench(0)

    # ---- the first commit builds the digit-multiple table of the generator set (one-off, outside the timed region)
    barrier()
    t0 = time.perf_counter()
    step_device(0)
    drain()
    barrier()
    first_call_s = time.perf_counter() - t0

    # ---- device-resident timing (library defaults: two chunks on two streams)
    for i in range(args.warmup):
        step_device(i)
    drain()
    barrier()
    ctx.counters(reset=True)
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    w0 = time.time()
    e0.record(stream)
    fork()
    for i in range(args.steps):
        step_device(args.warmup + i)
    drain()
    e1.record(stream)
    barrier()
    w1 = time.time()
    ms = e0.elapsed_time(e1)
    launches = ctx.counters()["kernel_launches"]
    clocks = sampler.stop(w0, w1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    points_per_step = L * R * world
    value = points_per_step * args.steps / (ms_max * 1e-3)
    last_buf = (args.warmup + args.steps - 1) % nbuf
    dC, dinf = dCs[(args.warmup + args.steps - 1) & 1], dinfs[(args.warmup + args.steps - 1) & 1]
    table_build_s = max(0.0, first_call_s - ms_max / args.steps * 1e-3)

    # ---- parity of the TIMED configuration: sampled rows of the last timed commit against the CPU oracle (the oracle is the
    #      checker here, never the thing measured; rank 0 only)
    parity = None
    if not args.no_parity and rank == 0:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle as orc
        orc.build()
        nchk = min(L, 32)
        rows = sorted(set(int(x) for x in np.linspace(0, L - 1, nchk)))
        zs = page_bufs[last_buf].reshape(L, R, 4)[rows].reshape(-1, 4)
        C_ref, inf_ref = orc.hyrax_commit(G, h, np.ascontiguousarray(zs), len(rows), R, None, threads=0)
        C_gpu = dC.cpu().numpy().view(np.uint64)[rows]
        inf_gpu = dinf.cpu().numpy()[rows]
        ok = bool(np.array_equal(C_gpu, C_ref) and np.array_equal(inf_gpu, inf_ref))
        parity = {"parity_checked": ok, "rows_checked": len(rows),
                  "what": "rows of the last timed commit (device-resident path, this table budget) vs oracle.hyrax_commit, bit-exact affine limbs + infinity flags"}
        if not ok:
            raise SystemExit("bench.py: the timed commit differs from the oracle on sampled rows -- refusing to print a number")

    # ---- the same K commits on ONE caller stream (strict stream order between consecutive commits), for comparison
    single_stream = None
    if len(cstreams) > 1:
        saved = list(cstreams)
        cstreams[:] = [stream]
        n1 = max(3, min(args.steps, 100))
        ms1 = timed_device(n1, args.warmup)
        cstreams[:] = saved
        single_stream = {"value": points_per_step * n1 / (ms1 * 1e-3), "unit": UNIT, "ms_per_step": ms1 / n1, "steps": n1,
                         "note": "consecutive commits on one caller stream: each waits for the previous one's last kernel"}

    # ---- stage profile on the library's own stream (CUDA events inside the library): the first `prof_rows` rows as ONE
    #      chunk, so the stages run back to back and each event pair brackets exactly one launch set of that stage
    #      (a single chunk of every row would need > 100 GB of workspace at 8192 x 8192)
    prof_rows = L if L * R <= (1 << 24) else max(1024, (1 << 24) // R)
    ctx.set("chunk_rows", prof_rows)
    ctx.set("mult_streams", 1)
    prof_ms = []
    for _ in range(5):
        ctx.hyrax_commit_device(bases, dev_bufs[0].data_ptr(), prof_rows, R, 0, dC.data_ptr(), dinf.data_ptr(), stream=0)
        prof_ms.append(ctx.last_commit_profile())
    prof = prof_ms[-1]
    acc_ms = sorted(p["accumulate"]["ms"] for p in prof_ms[1:])[len(prof_ms[1:]) // 2]      # median of the warm ones
    ctx.set("mult_streams", 2)
    ctx.set("chunk_rows", 0)      # back to the library default (auto)

    # ---- end to end: host buffers through the host-pointer C ABI (library defaults: a short first chunk, then chunks; the
    #      H2D copy of chunk i+1 overlaps the kernels of chunk i).  Pinned buffers first, then PAGEABLE ones (what an
    #      unmodified Rust caller's Vec<Scalar> is; sbn_host_alloc exists for callers that can allocate pinned)
    def timed_e2e(fn, steps):
        for i in range(args.warmup):
            fn(i)
        barrier()
        tt0 = time.perf_counter()
        for i in range(steps):
            fn(args.warmup + i)
        torch.cuda.synchronize()
        tt = torch.tensor([time.perf_counter() - tt0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return points_per_step * steps / float(tt.item())

    e2e_sync_value = timed_e2e(step_e2e, max(3, min(args.steps, 100)))
    e2e_value = timed_e2e(step_e2e_async, args.steps) if len(cstreams) > 1 else e2e_sync_value
    # the last asynchronous commit's commitments, read back on the host, against the device-resident leg's result for the
    # same input buffer is checked below (e2e_parity)
    e2e_last = (args.warmup + args.steps - 1)
    e2e_parity = None
    if len(cstreams) > 1:
        kk = e2e_last % NE
        ctx.hyrax_commit_device(bases, dev_bufs[e2e_last % nbuf].data_ptr(), L, R, 0, dCs[0].data_ptr(), dinfs[0].data_ptr(), stream=0)
        torch.cuda.synchronize()
        e2e_parity = bool(torch.equal(dCs[0].cpu(), hCs[kk]) and torch.equal(dinfs[0].cpu(), hinfs[kk]))
        if not e2e_parity:
            raise SystemExit("bench.py: the asynchronous host-pointer commit differs from the device-resident one")
    pC = np.empty((L, 8), dtype=np.uint64)
    pinf = np.empty((L,), dtype=np.uint8)

    def step_pageable(i):
        ctx.hyrax_commit_raw(bases, page_bufs[i % nbuf].ctypes.data, L, R, 0, pC.ctypes.data, pinf.ctypes.data)

    e2e_pageable = timed_e2e(step_pageable, max(3, min(args.steps, 50)))
    # pageable inputs AND outputs through the asynchronous call on the same three streams: the host threads that stage call
    # i + 1's scalars into the pinned ring run while call i's kernels do
    pCs = [np.empty((L, 8), dtype=np.uint64) for _ in range(NE)]
    pinfs = [np.empty((L,), dtype=np.uint8) for _ in range(NE)]

    def step_pageable_async(i):
        k = i % NE
        ctx.hyrax_commit_raw_async(bases, page_bufs[i % nbuf].ctypes.data, L, R, 0, pCs[k].ctypes.data, pinfs[k].ctypes.data,
                                   estreams[k].cuda_stream)

    e2e_pageable_async = None
    if len(cstreams) > 1:
        e2e_pageable_async = timed_e2e(step_pageable_async, max(3, min(args.steps, 50)))
        last = args.warmup + max(3, min(args.steps, 50)) - 1
        step_pageable(last)                                   # the same input through the blocking call, into pC / pinf
        if not (np.array_equal(pCs[last % NE], pC) and np.array_equal(pinfs[last % NE], pinf)):
            raise SystemExit("bench.py: the asynchronous pageable commit differs from the blocking one")

    # ---- the headline generator set's table (64 GB at the default budget) is released before the other legs build theirs
    mult_bits, mult_bytes = bases.mult_table()
    bases_window_bits = bases.window_bits
    torch.cuda.synchronize()
    bases.close()
    other_budget = min(args.table_mb, 36000)     # strong / prove legs: the budget their records were taken with

    # ---- the same commit under the LIBRARY's default table budget (6144 MiB: c = 13 at 1025 generators) -- the headline above
    #      uses --table-mb; a deployment that cannot spare that much HBM gets this number
    default_budget = None
    if args.table_mb != 6144 and L * R <= (1 << 24):
        ctx.set("mult_max_mb", 6144)
        bases_d = ctx.bases(G, h)
        nd = max(3, min(args.steps, 50))
        ms_d = timed_device(nd, args.warmup, bases_d)
        bits_d, bytes_d = bases_d.mult_table()
        default_budget = {"value": points_per_step * nd / (ms_d * 1e-3), "unit": UNIT, "ms_per_step": ms_d / nd,
                          "table_budget_mb": 6144, "table_window_bits": bits_d, "table_bytes": bytes_d, "steps": nd}
        bases_d.close()
    ctx.set("mult_max_mb", other_budget)

    # ---- the same commit on the REFERENCE's generator set (MultiCommitGens::new, commitments.rs:31-62): about two thirds of
    #      those generators are the same point, which the library merges (k_aggregate_rows), so the real prover's commits
    #      run faster than the distinct-generator headline above; reported separately, as SURVEY.md 8(d) asks
    ref_gens = None
    if args.gens == "distinct" and args.workload == "cfg1_1024x1024":
        gref = MultiCommitGens.new(R, b"gens_r1cs_eval", ctx)
        bases_ref = ctx.bases(gref.G, gref.h)
        distinct_pts = len({bytes(p) for p in np.concatenate([gref.G, gref.h.reshape(1, 8)]).view(np.uint8).reshape(R + 1, 64)})
        nr = max(3, min(args.steps, 50))
        tref = timed_device(nr, args.warmup, bases_ref)
        ref_gens = {"value": points_per_step * nr / (tref * 1e-3), "unit": UNIT,
                    "ms_per_step": tref / nr, "generators": R + 1, "distinct_points": distinct_pts,
                    "note": "MultiCommitGens::new(R, b\"gens_r1cs_eval\"): equal generators are merged by summing their scalars"}
        bases_ref.close()

    # ---- BASELINE configs[2], strong scaling: the keyless derefs shape (4096 rows x 8192 generators) divided by rows across
    #      the GPUs, every rank commits its block and the blocks are all-gathered.  Run at every N (N = 1 is the efficiency
    #      denominator), so the driver's --gpus N sweep carries it.
    strong = None
    if not args.no_strong and args.workload == "cfg1_1024x1024" and args.scaling == "weak" and 4096 % world == 0:
        strong = run_strong(ctx, synth, torch, dist, dev, stream, rank, world, args)

    # ---- "keyless prove time (s)": SNARK::prove of a synthetic keyless-shaped R1CS (2^20 constraints) through the GPU path,
    #      derefs commitment sharded by rows across the ranks (scripts/bench_snark.py).  The proof is not checked here -- the
    #      oracle is test infrastructure: tests/test_snark.py verifies the same keyless-scale proof with the CPU verifier.
    prove = None
    if not args.no_prove and args.workload == "cfg1_1024x1024":
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        import bench_snark
        res = bench_snark.run(20, verify=False, quiet=True, ctx_in=ctx)
        if res is not None:
            prove = {"seconds": res["ms"]["prove.SNARK_total"] / 1e3, "n_gpus": world,
                     "verified": "tests/test_snark.py::test_keyless_scale_proof_is_accepted (-m gpu) checks this proof with the CPU verifier",
                     "shape": "synthetic satisfiable R1CS, 2^20 constraints / variables, nnz padded to 2^22 (keyless shape)",
                     "first_call_seconds": res["ms"]["prove.first_call(cold kernels and workspaces)"] / 1e3,
                     "encode_seconds": res["ms"]["encode(dense representation + comb_ops/comb_mem commitments)"] / 1e3,
                     "phases_ms": res["prove_phases_ms"], "note": res["prove_total_note"],
                     "reference_published_s": 208.8,
                     "reference_published_note": "README of the reference: M2 Max, RAYON_NUM_THREADS=1 (other hardware; the same-host "
                                                 "figure is cpu_baseline below)"}
            if rank == 0 and not args.no_cpu_baseline:
                prove["cpu_baseline"] = cpu_prove_baseline(prove)

    if rank == 0:
        path_bits = mult_bits or bases_window_bits            # window width of the path that was TIMED
        W = (254 + path_bits) // path_bits
        Wb = (254 + bases_window_bits) // bases_window_bits    # the bucket method's window count (SURVEY 8(d) accounting)
        if args.scalars == "small":     # 21-bit values: only the windows that can hold a non-zero digit count as work
            W = min(W, -(-22 // path_bits))
            Wb = min(Wb, -(-22 // bases_window_bits))
        # one XYZZ mixed addition per (scalar, window) pair of the BUCKET method = Wb * 10 * 264 IMAD per point
        alg_imad_acc = float(prof_rows) * R * Wb * FQMUL_PER_MIXED_ADD * IMAD_PER_FQMUL
        formula = alg_imad_acc / (acc_ms * 1e-3) if acc_ms > 0 else 0.0
        # what the timed path EXECUTES: tabulated sum = one batched-affine addition (6 products) per (scalar, window) entry of
        # the table's own window width; bucket pipeline = one XYZZ mixed addition (10 products) per entry
        per_entry = 6 if mult_bits else FQMUL_PER_MIXED_ADD
        exec_imad_stage = float(prof_rows) * (R + 1) * W * per_entry * IMAD_PER_FQMUL
        # An executed fraction exists only where the work per entry is known: the tabulated-sum path with every row on the
        # full-width schedule.  The bucket pipeline mixes 6-product batched-affine and 10-product XYZZ additions in proportions
        # the library chooses per chunk, and a commit split into small-scalar runs has no single stage time.
        executed = exec_imad_stage / (acc_ms * 1e-3) if (acc_ms > 0 and mult_bits and args.scalars != "small") else None
        exec_imad_step = float(L) * (R + 1) * W * per_entry * IMAD_PER_FQMUL
        step_s = ms_max / args.steps * 1e-3
        traffic, traffic_note = stage_traffic(args.workload, mult_bits)
        hbm_peak, hbm_src = hbm_peak_gbs()
        step_alg = A_ADDS_PER_POINT.get(R, 26.0) * FQMUL_PER_MIXED_ADD * IMAD_PER_FQMUL
        if args.scalars == "small":
            step_alg = (Wb + 2.0 * (1 << (bases_window_bits - 1)) / R) * FQMUL_PER_MIXED_ADD * IMAD_PER_FQMUL
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "u32x8 Montgomery (integer, IMAD.WIDE carry chains)", "data": "synthetic",
            "config": {"workload": args.workload, "rows_per_gpu": L, "generators": R, "window_bits": path_bits,
                       "path": "tabulated sum over the resident digit-multiple table" if mult_bits else "bucket pipeline",
                       "bucket_pipeline_window_bits": bases_window_bits,
                       "table_budget_mb": args.table_mb, "table_bytes": mult_bytes, "table_build_s": table_build_s,
                       "other_legs_table_budget_mb": other_budget,
                       "table_note": "built once per generator set by the first commit of >= 256 rows, outside the timed region; "
                                     "value_default_budget is the same commit under the library's default 6144 MiB",
                       "gens": args.gens, "scalars": args.scalars, "blinds": "zero (derefs-style, hyrax.rs:301-305)",
                       "caller_streams": len(cstreams),
                       "caller_streams_note": "independent commits issued on alternating caller streams; the library takes its two workspace "
                                              "sets in turn, so one commit's tail overlaps the next one's head (roofline.frac is measured on "
                                              "ONE commit on ONE stream and does not benefit)",
                       "l2": f"inputs rotated over {nbuf} buffers ({nbuf * L * R * 32 >> 20} MiB > 126 MiB L2)",
                       "points_counted": "L x R scalar-base pairs per GPU per step",
                       "collective": ("NCCL all_gather of the commitment vector per step, issued asynchronously: it runs under the "
                                      "next step's commit (two output buffers) and is waited for inside the timed region")
                       if world > 1 else "none (1 GPU)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": L * R * 32, "d2h_bytes_per_step": L * 65,
                    "host_memory": "pinned",
                    "call": ("sbn_hyrax_commit_async on three caller streams taken in turn (three staging sets, two workspace sets in the "
                             "library): the H2D copy of step i + 2 runs under the kernels of steps i and i + 1; every step's H2D and "
                             "D2H are inside the timed region") if len(cstreams) > 1 else "sbn_hyrax_commit",
                    "synchronous_value": e2e_sync_value, "results_match_device_leg": e2e_parity,
                    "pageable_value": e2e_pageable, "pageable_async_value": e2e_pageable_async,
                    "pageable_note": "the same call from pageable numpy buffers (a Rust Vec<Scalar>); sbn_host_alloc gives callers pinned memory"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {
                "bound": "int32 multiply-add (IMAD pipe); not hbm, not tensor: modular integer arithmetic",
                "kernel": ("accumulation stage as a batched-affine sum tree over tabulated digit multiples (k_bat_prefix / "
                           "k_ba_invert / k_bat_finish per round + k_mult_sum_rows_t)") if mult_bits else
                          ("bucket accumulation stage (k_accumulate; with batched-affine rounds: k_ba_prefix / k_ba_invert / "
                           "k_ba_finish + k_accumulate_pts)"),
                "achieved": executed / 1e12 if executed else None, "peak": peak_imad / 1e12, "unit": "TIMAD/s",
                "frac": executed / peak_imad if (executed and peak_imad) else None,
                "frac_note": f"EXECUTED work of the stage: entries x {per_entry} Fq products x 264 IMAD / stage time / peak "
                             f"(entries = rows x (R + 1) x {W} windows of {path_bits} bits)",
                "executed_imad_per_launch_set": exec_imad_stage,
                "kernel_ms_per_launch_set": acc_ms, "rows_per_launch_set": prof_rows,
                "launch_set_note": "median of 4 warm commits of prof_rows rows as ONE chunk on ONE stream, CUDA events inside the "
                                   "library around the stage's launches",
                "whole_step_executed_frac": exec_imad_step / step_s / peak_imad if (executed and peak_imad) else None,
                "algorithmic_vs_formula": formula / peak_imad if (peak_imad and acc_ms > 0) else None,
                "algorithmic_vs_formula_note": "SURVEY 8(d)'s accounting (bucket method: W x 10 x 264 IMAD per point at the bucket "
                                               "window width) over the stage time; NOT a utilisation -- the path does less work than "
                                               "the formula assumes (wider window, 6-product additions), so this can exceed 1",
                "whole_step_vs_formula": (value / world) * step_alg / peak_imad if peak_imad else None,
                "traffic": traffic, "traffic_note": traffic_note,
                "hbm_frac": (traffic / (acc_ms * 1e-3) / 1e9 / hbm_peak) if (traffic and acc_ms > 0 and hbm_peak) else None,
                "hbm_peak_GBps": hbm_peak, "hbm_peak_source": hbm_src,
                "digit_multiple_table": {"window_bits": mult_bits, "bytes": mult_bytes,
                                         "note": "d * 2^(k c) * G_j for every window k, generator j and digit d <= 2^(c-1), "
                                                 "resident in HBM, built once per generator set"} if mult_bits else None,
                "peak_source": "measured live on this GPU: independent mad.lo.u32 streams (sbn_microbench kind 0); "
                               "MEASURED_PEAKS.json has no integer-pipe figure",
                "hbm": {"algorithmic_bytes_per_step": L * R * 32 + L * 64,
                        "achieved_GBps": (L * R * 32 + L * 64) / step_s / 1e9},
            },
            "stage_ms": {k: v["ms"] for k, v in prof.items()}, "stage_rows": prof_rows,
            "memory": ctx.memory_stats(),
        }
        if parity is not None:
            line.update(parity)
        if single_stream is not None:
            line["value_single_caller_stream"] = single_stream
        if default_budget is not None:
            line["value_default_budget"] = default_budget
        if ref_gens is not None:
            line["reference_generators"] = ref_gens
        if strong is not None:
            line["strong_cfg2_4096x8192"] = strong
        if prove is not None:
            line["keyless_prove"] = prove
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(R, L)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def source_sha16():
    """Hash of the CUDA sources: profiles/roofline_traffic.json is stamped with it, so a DRAM-traffic figure measured on other
    kernels is never printed as if it were current."""
    import hashlib
    hsh = hashlib.sha256()
    csrc = os.path.join(ROOT, "spartan_bn254_b200", "csrc")
    for d, _, files in sorted(os.walk(csrc)):
        for f in sorted(files):
            hsh.update(open(os.path.join(d, f), "rb").read())
    return hsh.hexdigest()[:16]


def stage_traffic(workload, mult_bits):
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        rec = json.load(open(tp)).get(workload, {})
    except Exception:
        return None, "profiles/roofline_traffic.json missing"
    if rec.get("source_sha16") != source_sha16():
        return None, ("stale: profiles/roofline_traffic.json was measured on other kernel sources (sha16 %s, now %s); re-run "
                      "scripts/update_traffic.py on an ncu launch list" % (rec.get("source_sha16"), source_sha16()))
    if rec.get("table_window_bits") != mult_bits:
        return None, "profiles/roofline_traffic.json was measured with another table width"
    return rec.get("accumulate_stage_dram_bytes_per_commit"), rec.get("source")


def hbm_peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json (driver-written copy bandwidth)"
    except Exception:
        return 6459.0, "fallback: the pool's measured copy bandwidth as recorded in round 1 (MEASURED_PEAKS.json absent)"


def run_strong(ctx, synth, torch, dist, dev, stream, rank, world, args):
    import numpy as np
    Ls, Rs = 4096 // world, 8192
    Gs, hs = synth.distinct_generators(ctx, Rs)
    ctx.set("mult_max_mb", max(args.table_mb, 50000) if args.table_mb else 0)     # c = 13 over 8193 generators is 43 GB
    bs = ctx.bases(Gs, hs)
    z = torch.from_numpy(synth.uniform_scalars(77 + rank, Ls * Rs).view(np.int64)).to(dev)
    dCs = torch.empty((Ls, 8), dtype=torch.int64, device=dev)
    dis = torch.empty((Ls,), dtype=torch.uint8, device=dev)
    gat = torch.empty((world * Ls, 8), dtype=torch.int64, device=dev) if world > 1 else None

    def step():
        ctx.hyrax_commit_device(bs, z.data_ptr(), Ls, Rs, 0, dCs.data_ptr(), dis.data_ptr(), stream=stream.cuda_stream)
        if world > 1:
            dist.all_gather_into_tensor(gat, dCs)

    for _ in range(2):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    n = 5
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record(stream)
    for _ in range(n):
        step()
    a1.record(stream)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    tt = torch.tensor([a0.elapsed_time(a1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms = float(tt.item()) / n
    bits, nbytes = bs.mult_table()
    bs.close()
    ctx.set("mult_max_mb", min(args.table_mb, 36000))
    return {"ms_per_commit": ms, "value": 4096 * Rs / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "rows_per_gpu": Ls,
            "generators": Rs, "scaling": "strong", "steps": n, "table_window_bits": bits, "table_bytes": nbytes,
            "inputs": "2 GiB / n_gpus of uniform scalars per GPU: larger than L2, not rotated",
            "note": "BASELINE configs[2] (keyless derefs shape) divided by rows across the GPUs + all_gather of the blocks; "
                    "time is the max over ranks"}


if __name__ == "__main__":
    main()
