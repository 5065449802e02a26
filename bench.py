#!/usr/bin/env python
"""bench.py -- Lurk prove throughput (trace rows / s) on B200, next to the CPU prover.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload lurk-fib|lurk-lcs|lurk-sum|fib|cons]
                    [--log-rows L] [--shards-per-gpu S] [--impl ours|reference]

Workloads.  The default is the reference's own fib bench (benches/fib.rs:36-43,71-163) for real: the Lurk program
`(letrec ((fib ...)) (fib n))` read from its source text, evaluated by `lurk_main` of the 39-func Lurk toplevel
(src/core/eval_direct.rs, ingress.rs, misc.rs restated as data) and proved on the 48-chip machine `new_machine`
builds for it (Entrypoint, 39 funcs among them hash3/4/5 and the u64 chips, 6 memories, bytes), with n chosen so
that the tallest chip (`eval`, 10 rows per fib step) has S * 2^L rows: BASELINE.json configs[1] at the default
L = 20, S = 1; the north star's "fib at 2^24 rows on one GPU" is --log-rows 22 --shards-per-gpu 4 (four shards of
the reference's default size, src/lair/execute.rs:226-241).  lurk-lcs (benches/lcs.rs:27-46; configs[2] is
--log-rows 22) and lurk-sum (benches/sum.rs:36-47) likewise.  `fib` is the Lair demo func of src/lair/mod.rs:65-83
(12 chips, the round-1 line), `cons` a hash-chip load.  FRI: sphinx defaults (100 queries, 16 proof-of-work bits).
A step is one complete prove of the record: trace generation, 3 commitments, quotient, openings, FRI, queries.

  value  rows/s with the flattened record already resident in HBM (lsk_shard_commit + lsk_shard_prove)
  e2e    rows/s through lsk_prove with the record in pinned HOST memory (upload + proof download inside)
  roofline      the kernel family with the largest share of the step, from CUDA events recorded by the library
                around every launch during K further steps of the same workload
  cpu_baseline  the CPU oracle prover (restated reference; the Rust reference cannot be built offline, SURVEY.md
                F1/F3) on this box's host cores, on a bounded sample of the same workload; the GPU proves the same
                sample and `config.parity_vs_oracle` says whether the two proofs are word-for-word equal
  N > 1  Lurk workloads: every rank proves its own record (independent proofs, no exchange; weak scaling), and the
         line carries a `strong` sub-record: ONE record proved by all N GPUs together (column-sharded device
         group), compared word for word with the single-GPU proof.  `fib` / `cons`: one reference shard per GPU
         with an all-gather of the 8-word roots (Shard::shard), or --partition column.

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


LURK_PROGRAMS = {"lurk-fib": "fib", "lurk-lcs": "lcs", "lurk-sum": "sum"}
REF_SITE = {"lurk-fib": "benches/fib.rs:36-43", "lurk-lcs": "benches/lcs.rs:27-46", "lurk-sum": "benches/sum.rs:36-47"}


class Workload:
    """One of the bench programs: how to build its toplevel, its record for a given size, and what to call it.
    `arg` is the program's size parameter (fib's n, the element count of sum, the prefix length of lcs, the Lair
    fib / cons_chain argument); rows(arg) = rows of the tallest chip."""

    def __init__(self, name):
        self.name = name
        self.lurk = name in LURK_PROGRAMS
        if self.lurk:
            from lurk_b200.lurk_eval import build_lurk_toplevel
            self.top, self.entry = build_lurk_toplevel(), "lurk_main"
            self.tallest = "eval"
        elif name == "cons":
            from lurk_b200.programs import hash_toplevel
            self.top, self.entry = hash_toplevel(), "cons_chain"
            self.tallest = "cons_chain"
        else:
            from lurk_b200.lair import demo_toplevel
            self.top, self.entry = demo_toplevel(), "fib"
            self.tallest = "fib"
        self.func = self.top.func_by_name(self.entry)

    def arg_for(self, log_rows, shards=1):
        if self.lurk:
            from lurk_b200.bench_programs import arg_for_rows
            return arg_for_rows(LURK_PROGRAMS[self.name], log_rows, shards)
        return (shards << log_rows) - 1          # Lair fib(n) / cons_chain(n): n + 1 rows

    def describe(self, arg, log_rows, shards, replicas):
        if self.lurk:
            prog = LURK_PROGRAMS[self.name]
            what = {"fib": f"(fib {arg})", "sum": f"sum of a {arg}-element list", "lcs": f"lcs of two {arg}-character strings"}[prog]
            s = (f"Lurk {what} [{REF_SITE[self.name]}] through lurk_main: 48-chip machine (Entrypoint, 39 funcs incl. "
                 f"hash3/4/5 and u64 chips, 6 memories, bytes), tallest chip `eval` <= {shards} x 2^{log_rows} rows, "
                 f"{shards} shard(s) of 2^{log_rows}")
        elif self.name == "cons":
            s = (f"Lurk-shaped cons_chain(n={arg}): per shard 2^{log_rows} rows of the 655-wide hash4 chip "
                 f"(Poseidon2 wide witness) + 2^{log_rows} x 37 caller rows, {shards} shard(s)")
        else:
            s = (f"Lair fib(n={arg}) of demo_toplevel [src/lair/mod.rs:65-83]: 2^{log_rows} trace rows x 18 cols per "
                 f"shard, {shards} shard(s), 12 chips")
        if replicas > 1:
            s += f"; {replicas} independent records, one per GPU"
        return s + "; FRI 100 queries / 16 PoW bits"

    # -- product side: QueryRecord of liblurkstark's native executor ------------------------------------------
    def record(self, arg):
        from lurk_b200.execute import QueryRecord, execute_by_name
        if self.lurk:
            from lurk_b200.bench_programs import setup
            inp, q, _ = setup(LURK_PROGRAMS[self.name], arg, self.top)
        else:
            inp, q = [arg], QueryRecord.new(self.top)
        out = execute_by_name(self.top, self.entry, inp, q)
        return q, [int(v) for v in out]

    def chip_rows(self, q):
        from lurk_b200 import _native as N
        rows = {f.name: int(N.lib().lsk_record_num_queries(q.h, f.index)) for f in self.top.funcs}
        return {k: v for k, v in rows.items() if v}

    def machine_blob(self):
        from lurk_b200 import host
        return host.chip_vector_blob(self.top.surface_image(), self.func.index)


# ------------------------------------------------------------------------------------------------
# CPU prover (oracle) -- cpu_baseline leg and --impl reference.  The only place bench.py touches oracle/.
# ------------------------------------------------------------------------------------------------
class CpuProver:
    def __init__(self, wl: Workload):
        from oracle import pyoracle
        pyoracle.build()
        self.o = pyoracle
        self.wl = wl
        self.ot = pyoracle.OracleToplevel(wl.top.serialize())
        self.m = pyoracle.OracleMachine(self.ot, wl.func.index, wl.machine_blob(), 100, 16)
        self.cores = os.cpu_count() or 1

    def record(self, arg):
        wl = self.wl
        rec = self.o.OracleRecord(self.ot)
        if wl.lurk:
            from lurk_b200.bench_programs import source
            from lurk_b200.reader import read
            from lurk_b200.zstore import ZStore
            z = ZStore()
            e = read(z, source(LURK_PROGRAMS[wl.name], arg))
            for name, width in (("hash3", 24), ("hash4", 32), ("hash5", 40)):
                if z.hashes[width]:
                    rec.inject_inv_queries(wl.top.func_by_name(name).index, [list(k) for k in z.hashes[width]],
                                           [list(v) for v in z.hashes[width].values()])
            inp = [e.tag] + [0] * 7 + list(e.digest) + [0] * 8
        else:
            inp = [arg]
        rec.execute(wl.func.index, inp, wl.func.output_size)
        return rec

    def rows(self, rec):
        return int(rec.num_queries(self.wl.top.func_by_name(self.wl.tallest).index))

    def prove(self, rec, shard_size=1 << 22):
        t = time.perf_counter()
        proof = self.m.prove(rec, shard_size)
        return time.perf_counter() - t, proof

    def sample_arg(self, log_rows, budget_s):
        """The largest sample (arg for 2^l rows, l <= log_rows) whose prove is expected to fit the budget, from two
        small timed proves (fixed cost + cost per row)."""
        la, lb = (14, 16) if self.wl.name == "fib" else (9, 12)
        la, lb = min(la, log_rows), min(lb, log_rows)
        ra = self.record(self.wl.arg_for(la))
        ta, _ = self.prove(ra)
        if lb == la:
            return la
        rb = self.record(self.wl.arg_for(lb))
        tb, _ = self.prove(rb)
        na, nb = self.rows(ra), self.rows(rb)
        per_row = max((tb - ta) / max(nb - na, 1), 1e-7)
        fixed = max(ta - per_row * na, 0.0)
        l = lb
        while l < log_rows and fixed + per_row * (1 << (l + 1)) <= budget_s:
            l += 1
        return l

    def describe(self):
        return f"oracle CPU prover (C++, std::thread x{self.cores})"


def cpu_baseline(wl, log_rows, budget_s=25.0):
    """Oracle prover on a bounded sample of the workload.  -> (cpu_baseline object, sample arg, the oracle's proof)"""
    cp = CpuProver(wl)
    l = cp.sample_arg(log_rows, budget_s)
    arg = wl.arg_for(l)
    rec = cp.record(arg)
    rows = cp.rows(rec)
    t, proof = cp.prove(rec)
    return ({"value": rows / t, "unit": UNIT, "cores": cp.cores, "kind": "port",
             "sample": f"{cp.describe()}, one prove of {wl.name} with {rows} rows of `{wl.tallest}` (2^{l} sample), {t:.2f} s"},
            arg, proof)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = args.gpus
    wl = Workload(args.workload)
    cp = CpuProver(wl)
    steps, warmup = args.steps, args.warmup
    l = cp.sample_arg(args.log_rows, 150.0 / max(1, steps + warmup))
    arg = wl.arg_for(l)
    rec = cp.record(arg)
    rows = cp.rows(rec)
    for _ in range(warmup):
        cp.prove(rec)
    total = sum(cp.prove(rec)[0] for _ in range(steps))
    value = steps * rows / total
    sample = f"{cp.describe()}, each step one prove of {wl.name} with {rows} rows of `{wl.tallest}` (2^{l} sample)"
    full_arg = wl.arg_for(args.log_rows, args.shards_per_gpu)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": 1e3 * total / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": wl.describe(full_arg, args.log_rows, args.shards_per_gpu, world if wl.lurk else 1),
                   "log_rows": args.log_rows, "sample_log_rows": l, "sample_arg": arg},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cp.cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "restated CPU baseline (oracle port: scalar C++, no SIMD), not the Rust reference binary: cargo and the "
                "un-vendored sphinx/Plonky3 git dependencies are unavailable offline (SURVEY.md F1/F3)",
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
    from lurk_b200.execute import Shard, ShardingConfig
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
    wl = Workload(args.workload)
    top, entry = wl.top, wl.entry
    m = new_machine(top, entry, ctx=ctx)
    if rank == 0 and wl.name.startswith("lurk") and int(N.lib().lsk_machine_jit_chips(m.h)) == 0:
        # a Lurk machine without compiled constraint programs proves correctly but ~25 % slower (every chip interpreted):
        # the cubin cache is stale -- an edit of field.cuh / quotient_kernel.cuh invalidates it; build() refills it
        print("bench.py: WARNING: no compiled constraint programs in lurk_b200/jit_cache for this library; run "
              "`python -m lurk_b200.jit_warm` (or __graft_entry__.build())", file=sys.stderr, flush=True)
    pk, vk = m.setup()
    K = args.shards_per_gpu
    column = args.partition == "column"   # every shard is proved by a GROUP of ranks together
    replicas = wl.lurk and world > 1 and not column   # every rank proves its own record: independent proofs
    g = (args.gpus_per_shard or world) if column else 1
    if world % g:
        raise SystemExit("--gpus-per-shard must divide --gpus")
    ngroups = 1 if replicas else world // g
    nshards = ngroups * K                 # shards of ONE record
    arg = wl.arg_for(L, nshards)
    t0 = time.perf_counter()
    q, out = wl.record(arg)
    q.finalize()                       # row-ordered tables: each rank uploads only its shard's slice
    t_exec = time.perf_counter() - t0
    chip_rows = wl.chip_rows(q)
    q.pin()                            # the record is the step's input: keep it in pinned host memory
    cfg = ShardingConfig(1 << L)
    shards = Shard(q, 0, cfg).shard(cfg)
    if len(shards) != nshards:
        raise SystemExit(f"{wl.name}({arg}) gives {len(shards)} shards of 2^{L} rows, expected {nshards}")
    whole = Shard(q, 0, cfg)
    shards_of = None
    if column:
        from lurk_b200.parallel import group_shards, join_device_group
        mine_all = [shards[rank // g + j * ngroups] for j in range(K)]   # group j proves shards j, j + ngroups, ...
        shards_of = group_shards(nshards, ngroups, g)
        if world > 1:
            join_device_group(ctx, g)  # the library splits each stage over the group and all-gathers (NCCL, own stream)
    elif replicas:
        mine_all = list(shards)
    else:
        mine_all = [shards[rank + j * world] for j in range(K)]   # rank r proves shards r, r + G, ...
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
            roots = np.stack([mine_roots[i] for i in range(nshards)])
        return {s.index: m.prove_shard(s, roots) for s in mine_all}

    def step_e2e():
        if replicas or world == 1 or (column and ngroups == 1):
            return m.prove(pk, whole).words     # lsk_prove: upload, all shards of the record, one proof
        m.release_record()             # force the host -> device upload of the record
        parts = step_resident()
        # the proof of the job is the concatenation of the shard proofs: collect them on rank 0
        parts = gather_shard_proofs(parts, nshards, 0, shards_of)
        return assemble_proof(parts).words if rank == 0 else None

    def timed(step, steps, profile):
        total_ms = 0.0
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

    nrec = world if replicas else 1
    rows = nrec * nshards * (1 << L)
    value = rows * args.steps / (res_ms / 1e3)
    e2e_value = rows * args.steps / (e2e_ms / 1e3)

    # ---- N > 1, Lurk workloads: the SAME record proved by all GPUs together (strong scaling) ----
    strong = None
    if replicas:
        from lurk_b200.parallel import join_device_group
        m.release_record()
        join_device_group(ctx, world)
        ctx.profile_enable(False)
        grouped = m.prove(pk, whole).words
        same = torch.tensor([int(np.array_equal(grouped, proof))], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        m.prove(pk, whole)
        strong_ms, _ = timed(lambda: m.prove(pk, whole), args.steps, False)
        info = ctx.comm_info()
        strong = {"value": nshards * (1 << L) * args.steps / (strong_ms / 1e3), "unit": UNIT, "n_gpus": world,
                  "ms_per_step": strong_ms / args.steps, "scaling": "strong", "rows_per_step": nshards * (1 << L),
                  "what": "ONE record (host buffers in, proof out: the e2e call) proved by all GPUs as one device group",
                  "proof_equals_single_gpu_proof": bool(int(same)),
                  "nccl_bytes_received_per_rank": info["bytes_received"]}
        ctx.comm_destroy()

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
            # second, live, against the peak of a dependency-free Montgomery product stream on this chip -- a figure
            # this repo measured itself (tools/fp64_probe.cu "mont_int": 11.97 products/clk/SM x 148 SMs x 1.965 GHz,
            # profiles/r01_fp64_probe.jsonl with the probe's SASS mix beside it), not a driver-written one
            mm = tv["work"] * MODMUL_PER_PERM16 / (tv["ms"] * 1e-3)
            roofline["compute"] = {"bound": "integer multiply pipe", "achieved": mm / 1e12, "peak": MODMUL_PEAK / 1e12,
                                   "unit": "T modmul/s", "frac": mm / MODMUL_PEAK, "peak_source": "self-measured (tools/fp64_probe.cu)",
                                   "perm_per_s": tv["work"] / (tv["ms"] * 1e-3)}
        hash_ms = sum(v["ms"] for k, v in prof.items() if v.get("work"))
        hash_work = sum(v["work"] for k, v in prof.items() if v.get("work"))
        if hash_ms > 0:
            roofline["all_hash_kernels"] = {"ms_per_step": hash_ms / args.steps, "perm_per_s": hash_work / (hash_ms * 1e-3),
                                            "modmul_frac_of_self_measured_peak": hash_work * MODMUL_PER_PERM16 / (hash_ms * 1e-3) / MODMUL_PEAK}
        if column:
            partition = ("column-sharded: %d group(s) of %d GPUs, every shard proved by one group together (LDE by column block, "
                         "Merkle / quotient / reduced openings by row range, NCCL)" % (ngroups, g))
        elif replicas:
            partition = "independent records: every rank proves its own record (all of its shards), no exchange"
        else:
            partition = "one shard per GPU at a time (rank r proves shards r, r+G, ...), all-gather of the 8-word roots"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": res_ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if column and ngroups == 1 else "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": wl.describe(arg, L, nshards, nrec), "program_arg": arg, "log_rows": L, "shards": nshards,
                       "records": nrec, "partition": partition,
                       "rows_per_step": rows, "rows_definition": "padded height of the tallest chip x shards (x records)",
                       "rows_tallest_chip": max(chip_rows.values()), "rows_all_chips": sum(chip_rows.values()),
                       "chips_with_rows": len(chip_rows),
                       "l2": "flushed between timed steps (256 MiB write)",
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
        if strong:
            line["strong"] = strong
        if world == 1 and not args.no_cpu_baseline:
            # the CPU prover on a bounded sample; the GPU proves the same sample: word-for-word parity of the two proofs
            cb, sarg, oproof = cpu_baseline(wl, L)
            qs, _ = wl.record(sarg)
            gproof = m.prove(pk, Shard.new(qs)).words
            line["cpu_baseline"] = cb
            line["config"]["parity_vs_oracle"] = bool(np.array_equal(gproof, oproof))
            line["config"]["parity_sample"] = f"{wl.name} arg {sarg}: GPU proof ({len(gproof)} words) vs the oracle's proof of the same record"
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--log-rows", type=int, default=20, help="log2 rows of the tallest chip per shard")
    ap.add_argument("--shards-per-gpu", type=int, default=1,
                    help="shards each GPU proves per step (--log-rows 22 --shards-per-gpu 4 = 2^24 rows in reference-size shards)")
    ap.add_argument("--partition", default="shard", choices=["shard", "column"],
                    help="multi-GPU mode: shard = independent records (Lurk workloads) / one shard per GPU (fib, cons); column = "
                         "all GPUs prove each of --shards-per-gpu shards together (strong scaling; configs[3])")
    ap.add_argument("--gpus-per-shard", type=int, default=0,
                    help="with --partition column: GPUs that prove one shard together (default: all of them); "
                         "--gpus 8 --gpus-per-shard 2 proves 4 shards at a time, two GPUs each")
    ap.add_argument("--workload", default="lurk-fib", choices=["lurk-fib", "lurk-lcs", "lurk-sum", "fib", "cons"],
                    help="lurk-fib: benches/fib.rs through lurk_main (default, BASELINE configs[1]); lurk-lcs: benches/lcs.rs "
                         "(configs[2] with --log-rows 22); lurk-sum: benches/sum.rs; fib: the Lair demo func; cons: hash-chip load")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
