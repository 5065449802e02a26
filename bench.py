#!/usr/bin/env python
"""bench.py -- Lurk-fib prove throughput (trace rows / s) on B200, next to the CPU prover.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--log-rows L] [--impl ours|reference]

Workload (BASELINE.json configs[1]): Lair `fib` of the reference's demo toplevel
(src/lair/mod.rs:65-83) with n = G * 2^L - 1, i.e. exactly 2^L trace rows per shard and one shard
per GPU (weak scaling); 12 chips (Entrypoint, 4 funcs of which fib is non-empty, 6 memory tables,
byte table); sphinx default FRI parameters (100 queries, 16 proof-of-work bits).  A step is one
complete prove of the record: trace generation, 3 commitments, quotient, openings, FRI, queries.

  value  rows/s with the flattened record already resident in HBM (lsk_shard_commit + lsk_shard_prove)
  e2e    rows/s through lsk_prove with the record in HOST memory (upload + proof download inside)
  roofline      the kernel family with the largest share of the step, from CUDA events recorded by
                the library around every launch during K further steps of the same workload
  cpu_baseline  the CPU oracle prover (restated reference; the Rust reference cannot be built
                offline) on this box's host cores

`--impl reference` times that CPU prover alone (rank 0 only under torchrun).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "lurk_fib_prove_trace_rows_per_sec"
UNIT = "rows/s"


MODMUL_PER_PERM16 = 8 * 16 * 4 + 13 * (4 + 16)      # x^7 = 4 products; 8 full rounds, 13 partial rounds (s-box + diagonal)
MODMUL_PEAK = 11.97 * 148 * 1.965e9                 # measured Montgomery products/s (profiles/r01_fp64_probe.jsonl)


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def ncu_summary():
    """Per-kernel figures read off the committed `ncu --set full` capture (profiles/ncu_summary.json):
    DRAM bytes per launch of the dominant kernel and the utilisation of the pipe that bounds it."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_summary.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def workload_name(log_rows, world, workload="fib"):
    if workload == "cons":
        return (f"Lurk-shaped cons_chain(n={world}*2^{log_rows}-1): per shard 2^{log_rows} rows of the 655-wide hash4 chip "
                f"(Poseidon2 wide witness) + 2^{log_rows} x 37 caller rows, {world} shard(s), FRI 100 queries / 16 PoW bits")
    return (f"Lair fib(n={world}*2^{log_rows}-1) of demo_toplevel: 2^{log_rows} trace rows x 18 cols per shard, "
            f"{world} shard(s), 12 chips, FRI 100 queries / 16 PoW bits")


def workload_program(workload):
    """(toplevel, entry func name) -- fib: src/lair/mod.rs:65-83 (benches/fib.rs shape); cons: lurk_b200/programs.py"""
    if workload == "cons":
        from lurk_b200.programs import hash_toplevel
        return hash_toplevel(), "cons_chain"
    from lurk_b200.lair import demo_toplevel
    return demo_toplevel(), "fib"


# ------------------------------------------------------------------------------------------------
# CPU prover (oracle) -- cpu_baseline leg and --impl reference
# ------------------------------------------------------------------------------------------------
class CpuProver:
    def __init__(self, workload="fib"):
        from oracle import pyoracle
        from lurk_b200.air import build_chip_vector, serialize_machine
        from lurk_b200.lair import FuncChip
        pyoracle.build()
        self.o = pyoracle
        top, entry = workload_program(workload)
        fc = FuncChip.from_name(entry, top)
        self.func = fc.func.index
        self.nout = fc.func.output_size
        blob = serialize_machine(build_chip_vector(fc), fc.func.input_size + fc.func.output_size)
        self.ot = pyoracle.OracleToplevel(top.serialize())
        self.m = pyoracle.OracleMachine(self.ot, self.func, blob, 100, 16)
        self.cores = os.cpu_count() or 1

    def record(self, log_rows):
        rec = self.o.OracleRecord(self.ot)
        rec.execute(self.func, [(1 << log_rows) - 1], self.nout)
        return rec

    def time_prove(self, rec, log_rows):
        t = time.perf_counter()
        self.m.prove(rec, 1 << log_rows)
        return time.perf_counter() - t


def cpu_baseline(log_rows, budget_s=25.0, workload="fib"):
    """Oracle prover on a bounded sample: the largest 2^l <= 2^log_rows rows whose prove fits the budget."""
    cp = CpuProver(workload)
    la, lb = (14, 16) if workload == "fib" else (8, 10)
    ta = cp.time_prove(cp.record(la), la)
    tb = cp.time_prove(cp.record(lb), lb)
    per_row = max((tb - ta) / ((1 << lb) - (1 << la)), 1e-7)
    fixed = max(ta - per_row * (1 << la), 0.0)
    l = lb
    while l < log_rows and fixed + per_row * (1 << (l + 1)) <= budget_s:
        l += 1
    t = cp.time_prove(cp.record(l), l) if l > lb else tb
    return {"value": (1 << l) / t, "unit": UNIT, "cores": cp.cores, "kind": "port",
            "sample": f"oracle CPU prover (C++, std::thread x{cp.cores}), one prove of {workload} at 2^{l} rows, {t:.2f} s"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = args.gpus
    cp = CpuProver(args.workload)
    steps, warmup = args.steps, args.warmup
    la, lb = (14, 16) if args.workload == "fib" else (8, 10)
    ta = cp.time_prove(cp.record(la), la)
    tb = cp.time_prove(cp.record(lb), lb)
    per_row = max((tb - ta) / ((1 << lb) - (1 << la)), 1e-7)
    fixed = max(ta - per_row * (1 << la), 0.0)
    budget = 150.0 / max(1, steps + warmup)
    l = la
    while l < args.log_rows and fixed + per_row * (1 << (l + 1)) <= budget:
        l += 1
    rec = cp.record(l)
    for _ in range(warmup):
        cp.time_prove(rec, l)
    times = [cp.time_prove(rec, l) for _ in range(steps)]
    total = sum(times)
    value = steps * (1 << l) / total
    sample = f"oracle CPU prover (C++, std::thread x{cp.cores}), each step one prove of {args.workload} at 2^{l} rows"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": 1e3 * total / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": workload_name(args.log_rows, world, args.workload), "log_rows": args.log_rows,
                   "sample_log_rows": l},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cp.cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "restated CPU baseline (oracle port), not the Rust reference binary: cargo and the un-vendored "
                "sphinx/Plonky3 git dependencies are unavailable offline (SURVEY.md F1/F3)",
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region: an NVML polling thread (every 10 ms; in-process, so it
    has samples even when the region lasts a fraction of a second), `nvidia-smi -lms` when NVML is not importable."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []
        self.nvml = None
        self.samples = []        # (sm MHz, reasons bitmask)
        self.max_mhz = None
        self.stop_flag = threading.Event()

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.device)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons

            def poll():
                while not self.stop_flag.is_set():
                    try:
                        self.samples.append((float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)), int(reasons(h))))
                    except Exception:
                        pass
                    self.stop_flag.wait(0.01)
            self.nvml = pynvml
            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.nvml:
            self.stop_flag.set()
            self.thread.join(timeout=1)
            sm = [s for s, _ in self.samples]
            mask = 0
            for _, r in self.samples:
                mask |= r
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_mhz, "samples": len(sm),
                    "reasons": sorted(name for bit, name in self.REASONS if mask & bit), "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi"}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from lurk_b200 import _native as N
    from lurk_b200.execute import QueryRecord, Shard, ShardingConfig, execute_by_name
    from lurk_b200.machine import assemble_proof, new_machine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: liblurkstark has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = N.Context(local)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)

    L = args.log_rows
    top, entry = workload_program(args.workload)
    m = new_machine(top, entry, ctx=ctx)
    pk, vk = m.setup()
    q = QueryRecord.new(top)
    t0 = time.perf_counter()
    K = args.shards_per_gpu
    column = args.partition == "column"   # every shard is proved by a GROUP of ranks together
    g = (args.gpus_per_shard or world) if column else 1
    if world % g:
        raise SystemExit("--gpus-per-shard must divide --gpus")
    ngroups = world // g
    nshards = ngroups * K
    out = execute_by_name(top, entry, [nshards * (1 << L) - 1], q)
    q.finalize()                       # row-ordered tables: each rank uploads only its shard's slice
    t_exec = time.perf_counter() - t0
    q.pin()                            # the record is the step's input: keep it in pinned host memory
    cfg = ShardingConfig(1 << L)
    shards = Shard(q, 0, cfg).shard(cfg)
    assert len(shards) == nshards, (len(shards), world, K)
    shards_of = None
    if column:
        from lurk_b200.parallel import group_shards, join_device_group
        mine_all = [shards[rank // g + j * ngroups] for j in range(K)]   # group j proves shards j, j + ngroups, ...
        whole = Shard(q, 0, cfg)
        shards_of = group_shards(nshards, ngroups, g)
        if world > 1:
            join_device_group(ctx, g)  # the library splits each stage over the group and all-gathers (NCCL, own stream)
    else:
        mine_all = [shards[rank + j * world] for j in range(K)]   # rank r proves shards r, r + G, ...
    mine = mine_all[0]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from lurk_b200.parallel import all_gather_roots, gather_shard_proofs

    def step_resident():
        mine_roots = {s.index: m.commit_shard(s) for s in mine_all}
        if ngroups > 1:
            roots = all_gather_roots(mine_roots, nshards, shards_of)
        else:
            roots = np.stack([mine_roots[i] for i in range(K)])
        return {s.index: m.prove_shard(s, roots) for s in mine_all}

    def step_e2e():
        if column and ngroups == 1:
            return m.prove(pk, whole).words     # every rank uploads the record and ends with the same proof
        if world == 1:
            return m.prove(pk, mine).words      # lsk_prove: upload, all K shards, one proof
        m.release_record()             # force the host -> device upload of the record
        parts = step_resident()
        # the proof of the job is the concatenation of the shard proofs: collect them on rank 0
        parts = gather_shard_proofs(parts, nshards, 0, shards_of)
        return assemble_proof(parts).words if rank == 0 else None

    def timed(step, steps, profile):
        total_ms = 0.0
        launches = 0
        prof_acc = {}
        for _ in range(steps):
            flush.fill_(1)             # evict L2 between iterations (256 MiB write, outside the timed pair)
            barrier()
            if profile:
                ctx.profile_enable(True)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            step()
            b.record(stream)
            barrier()
            ms = torch.tensor([a.elapsed_time(b)], device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            total_ms += float(ms)
            if profile:
                for k, v in ctx.profile_read().items():
                    acc = prof_acc.setdefault(k, {"ms": 0.0, "bytes": 0.0, "launches": 0, "work": 0.0})
                    for f in acc:
                        acc[f] += v[f]
                ctx.profile_enable(False)
        return total_ms, prof_acc

    for _ in range(max(args.warmup, 3)):
        step_resident()
    proof = step_e2e()
    verify_ms = None
    if rank == 0:                      # machine.verify (host C++ verifier of the library) accepts the job's proof
        t0 = time.perf_counter()
        m.verify(vk, proof)
        verify_ms = 1e3 * (time.perf_counter() - t0)
    step_resident()                    # re-upload the record so that it is resident for the timed steps
    barrier()

    clocks = ClockSampler(local)
    clocks.start()
    res_ms, _ = timed(step_resident, args.steps, False)
    e2e_ms, _ = timed(step_e2e, args.steps, False)
    h2d, d2h = m.last_transfer_bytes()
    step_resident()
    # the same K steps again with the library's per-kernel CUDA events on (they cost ~2 us per
    # launch, so the headline numbers above are taken without them)
    prof_ms, prof = timed(step_resident, args.steps, True)
    clk = clocks.stop()

    rows = nshards * (1 << L)
    value = rows * args.steps / (res_ms / 1e3)
    e2e_value = rows * args.steps / (e2e_ms / 1e3)

    if rank == 0:
        peak, peak_kind = measured_peaks()
        kernels = {}
        top_k, top_ms = None, -1.0
        for k, v in prof.items():
            gbs = v["bytes"] / (v["ms"] * 1e6) if v["ms"] > 0 else 0.0
            kernels[k] = {"ms_per_step": v["ms"] / args.steps, "launches_per_step": v["launches"] // args.steps,
                          "algorithmic_GBps": gbs}
            if v.get("work"):      # hash kernels: width-16 Poseidon2 permutations, 772 modular products each
                kernels[k]["poseidon2_perm_per_s"] = v["work"] / (v["ms"] * 1e-3)
            if v["ms"] > top_ms:
                top_k, top_ms = k, v["ms"]
        tv = prof[top_k]
        achieved = tv["bytes"] / (tv["ms"] * 1e6)
        gpu_ms = sum(v["ms"] for v in prof.values())
        ncu = ncu_summary().get(top_k, {})
        roofline = {"kernel": top_k, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": ncu.get("dram_bytes_per_launch"),
                    "traffic_source": ncu.get("source"), "pipe": ncu.get("pipe"),
                    "peak_source": f"MEASURED_PEAKS.json ({peak_kind})",
                    "bytes_per_launch": tv["bytes"] / max(1, tv["launches"]),
                    "ms_per_launch": tv["ms"] / max(1, tv["launches"]),
                    "share_of_step_gpu_time": tv["ms"] / gpu_ms if gpu_ms else None,
                    "profiled_ms_per_step": prof_ms / args.steps, "kernel_ms_per_step": gpu_ms / args.steps}
        if tv.get("work"):
            # the bound that actually applies (Poseidon2 is integer-pipe bound): modular products per
            # second, live, against the measured peak of a dependent-free Montgomery product stream on this chip
            # (tools/fp64_probe.cu "mont_int": 11.97 products/clk/SM x 148 SMs x 1.965 GHz, profiles/r01_fp64_probe.jsonl)
            mm = tv["work"] * MODMUL_PER_PERM16 / (tv["ms"] * 1e-3)
            roofline["compute"] = {"bound": "integer multiply pipe", "achieved": mm / 1e12, "peak": MODMUL_PEAK / 1e12,
                                   "unit": "T modmul/s", "frac": mm / MODMUL_PEAK,
                                   "perm_per_s": tv["work"] / (tv["ms"] * 1e-3)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": res_ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if column and ngroups == 1 else "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": workload_name(L, nshards, args.workload), "log_rows": L, "shards": nshards,
                       "partition": ("column-sharded: %d group(s) of %d GPUs, every shard proved by one group together (LDE by "
                                     "column block, Merkle / quotient / reduced openings by row range, NCCL all-gathers)"
                                     % (ngroups, g))
                       if column else "one shard per GPU at a time (rank r proves shards r, r+G, ...)",
                       "rows_per_step": rows, "l2": "flushed between timed steps (256 MiB write)",
                       "executor_s": round(t_exec, 3), "output_word0": int(out[0]), "proof_words": int(len(proof)),
                       "proof_verified_ms": verify_ms,
                       "compiled_constraint_programs": int(N.lib().lsk_machine_jit_chips(m.h))},
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(sum(v["launches"] for v in prof.values())),
            "clocks": clk,
            "roofline": roofline,
            "kernels": kernels,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(L, workload=args.workload)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--log-rows", type=int, default=20, help="log2 trace rows per shard (= per GPU)")
    ap.add_argument("--shards-per-gpu", type=int, default=1,
                    help="shards each GPU proves per step (configs[3]: --log-rows 22 --shards-per-gpu 4 = 2^24 rows)")
    ap.add_argument("--partition", default="shard", choices=["shard", "column"],
                    help="multi-GPU mode: shard = one shard per GPU (weak scaling, the default); column = all GPUs prove "
                         "each of --shards-per-gpu shards together (strong scaling; configs[3])")
    ap.add_argument("--gpus-per-shard", type=int, default=0,
                    help="with --partition column: GPUs that prove one shard together (default: all of them); "
                         "--gpus 8 --gpus-per-shard 2 proves 4 shards at a time, two GPUs each")
    ap.add_argument("--workload", default="fib", choices=["fib", "cons"],
                    help="fib: BASELINE configs[1] (default); cons: Lurk-shaped hash-chip load (configs[2] substitute)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
