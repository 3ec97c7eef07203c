#!/usr/bin/env python
"""Benchmark of the simplex hot path (BASELINE.json metric: "batched LP solves/sec at
1/2/4/8 B200; pivots/sec on single large LP").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c5|c2|c3|c4]
                  [--numerics exact|fast]

A STEP is one pass of the hot path over one batch of synthetic input: every LP of the batch
is solved from scratch on the device.

Workloads
  c5 (default)  BASELINE configs[4]: LPs of m=64 x n=128 (lowered 192x448), the batch the
                multi-GPU metric is quoted on.  A FIXED total batch (--total-lps) is sharded
                by contiguous LP id over the N ranks: STRONG scaling.  The full 262 144-LP
                batch is hours of GPU time; the default total (9472 = 8 x 1184) keeps a step
                at ~25 s on one GPU, so the driver's 25-step run fits its per-N time limit.
  c2            BASELINE configs[1]: 4096 LPs of m=32 x n=64 (lowered 96x224) on one GPU; with
                N > 1 every rank gets its own 4096 (weak scaling, as in round 1).
  c3, c4        BASELINE configs[2], configs[3]: ONE large LP (dense 2000x4000 packing LP;
                sparse transportation-style 20k x 50k), metric pivots/s over a fixed pivot
                prefix (--prefix).  A single LP does not shard (serially dependent pivots):
                with N > 1 every rank runs a replica and the value is the sum.

`value`   units/s with the inputs already resident in HBM: device time of the solve (CUDA
          events on the launch stream), summed over the K steps, max over ranks.
`e2e`     the same through the C ABI with HOST buffers, measured in the SAME K steps: pinned
          host -> device copy of the step's inputs, solve, device -> host copy of the results,
          and (N > 1) the final gather of per-LP results, wall clock, max over ranks.
`parity`  every run checks a strided sample of its own results against the CPU oracle
          (status, pivot count, pivot trace hash, objective bits).
`--numerics fast` (c2, c5 only) runs the OPT-IN fast-numerics kernel instead (dz_fast.cu): a
separately reported line whose `parity` block is agreement with the oracle to 1e-9, not bit parity.
`--impl reference` times the reference's CPU algorithm on the box's host cores (the literal
oracle restatement oracle/dzo.cpp: the Rust crate cannot be built in this image).
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import statistics
import struct
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

BATCHED = {
    # name: (m, n, generator name, default total LPs, LPs per core per CPU step, oracle parity sample)
    "c5": (64, 128, "config5", 9472, 1, 3),
    "c2": (32, 64, "config2", 4096, 2, 8),
}
SINGLE = {"c3": "dense packing LP m=2000 x n=4000 (lowered 6000x14000)",
          "c4": "sparse transportation-style LP, 10000 supply + 10000 demand rows, 50000 arcs x 20 nnz "
                "(m=20k x n=50k, lowered 70000x170000, nnz 2.17 M)"}


def metric_of(workload: str) -> tuple[str, str]:
    return ("pivots/sec on single large LP", "pivots/s") if workload in SINGLE else ("batched LP solves/sec", "LP/s")


def _bits(x: float) -> str:
    return struct.pack("<d", float(x)).hex()


def _host_cores() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def _lib_sha16() -> str:
    from dantzig_b200 import _capi

    return hashlib.sha256(open(_capi.LIB_PATH, "rb").read()).hexdigest()[:16]


KERNEL_FILES = {  # traffic.json tag -> the source file that defines its kernel
    "c5": "dz_kernel.cu", "c2": "dz_kernel.cu", "c5-core": "dz_core.cu", "c3": "dz_grid.cu", "c4": "dz_grid.cu",
    "c2_fast": "dz_fast.cu", "c5_fast": "dz_fast.cu", "c5_fast_dmma": "dz_fast.cu",
}


def _src_sha16(tag: str | None = None) -> str:
    """Hash of kernel SOURCES: what identifies a build across machines (two nvcc builds of the same
    sources are not byte-identical).  With a traffic.json tag: the files that hold that kernel's
    device code (its .cu and the shared device helpers); without: everything under dantzig_b200/csrc
    plus the public header."""
    h = hashlib.sha256()
    csrc = os.path.join(ROOT, "dantzig_b200", "csrc")
    if tag is None:
        files = sorted(os.listdir(csrc)) + [os.path.join("..", "..", "include", "dantzig_b200.h")]
    else:
        files = [KERNEL_FILES[tag], "dz_device.cuh"]
    for f in files:
        h.update(open(os.path.join(csrc, f), "rb").read())
    return h.hexdigest()[:16]


# --------------------------------------------------------------------------- workloads
def batched_workload(name: str, count: int, first: int):
    from dantzig_b200 import generate

    return getattr(generate, BATCHED[name][2])(count, first=first)


def single_workload(name: str):
    """(template, theta[1, n_theta], lower() -> oracle Lowered) of the one large LP."""
    from dantzig_b200 import Template, generate
    from dantzig_b200.model import model_from_theta

    if name == "c3":
        w = generate.packing(1, 2000, 4000)
        t = Template(w.structure)
        return t, w.theta[:1], (lambda: __import__("oracle.dzo_py", fromlist=["x"]).lower(
            model_from_theta(w.structure, w.theta[0])))
    model = generate.transportation_model(0, 10000, 10000, 50000, 10)
    t = Template(model)
    return t, t.pack_theta(model)[None, :], (lambda: __import__("oracle.dzo_py", fromlist=["x"]).lower(model))


# --------------------------------------------------------------------------- CPU arm
def _cpu_solve_range(args):
    """Worker: regenerate LPs [first, first+count) and solve them with the LITERAL oracle (the
    reference's own arithmetic: two fresh dense LUs per pivot, linalg.rs:8-10)."""
    name, first, count = args
    from dantzig_b200.model import model_from_theta
    from oracle import dzo_py

    w = batched_workload(name, count, first)
    piv = 0
    for i in range(count):
        piv += dzo_py.lower(model_from_theta(w.structure, w.theta[i])).solve(dzo_py.LITERAL).pivots
    return count, piv


def cpu_batched_pass(name: str, per_core: int, cores: int, pool) -> tuple[float, int, int]:
    jobs = [(name, i * per_core, per_core) for i in range(cores)]
    t0 = time.perf_counter()
    res = pool.map(_cpu_solve_range, jobs)
    return time.perf_counter() - t0, sum(r[0] for r in res), sum(r[1] for r in res)


def cpu_single_prefix(name: str, prefix: int) -> tuple[float, int]:
    """The one large LP on one host core.  The reference's literal arithmetic cannot follow
    these sizes (a dense 6000^2 LU twice per pivot is minutes per pivot; config 4's dense
    lowering would need 95 GB), so the CPU leg is the oracle's sparse-row variant: the
    reference's operations in the reference's order with exact-zero work skipped."""
    from oracle import dzo_py

    _, _, lower = single_workload(name)
    lo = lower()
    t0 = time.perf_counter()
    r = lo.solve(dzo_py.SPARSE, max_pivots=prefix)
    return time.perf_counter() - t0, int(r.pivots)


def run_reference(args) -> None:
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import multiprocessing as mp

    from oracle import dzo_py

    dzo_py.build()
    metric, unit = metric_of(args.workload)
    cores = _host_cores()
    t_tot, units = 0.0, 0
    if args.workload in BATCHED:
        m, n, _, _, per_core, _ = BATCHED[args.workload]
        with mp.get_context("fork").Pool(cores) as pool:
            for _ in range(args.warmup):
                cpu_batched_pass(args.workload, per_core, cores, pool)
            for _ in range(args.steps):
                dt, nlp, _ = cpu_batched_pass(args.workload, per_core, cores, pool)
                t_tot += dt
                units += nlp
        used = cores
        sample = (f"{per_core * cores} LPs per step ({per_core} per core, one solve per core via a process pool), "
                  f"literal oracle (two fresh dense LUs per pivot)")
        wl = f"{args.workload}: sample of the m={m} x n={n} batch, all <= rows, nonneg vars, seed 1234"
        kind_note = "reference = literal C++ restatement of simplex.rs + linalg.rs (oracle/dzo.cpp, -O2 -ffp-contract=off)"
    else:
        prefix = min(args.prefix, 60)
        for _ in range(min(args.warmup, 1)):
            cpu_single_prefix(args.workload, prefix)
        for _ in range(args.steps):
            dt, piv = cpu_single_prefix(args.workload, prefix)
            t_tot += dt
            units += piv
        used = 1
        sample = f"first {prefix} pivots per step, one core, sparse-row oracle (the literal dense algorithm cannot run this size)"
        wl = f"{args.workload}: {SINGLE[args.workload]}"
        kind_note = "reference = sparse-row C++ restatement (oracle/dzo.cpp DZO_SPARSE): the reference's operations and order, exact-zero work skipped"
    value = units / t_tot
    print(json.dumps({
        "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps,
        "higher_is_better": True, "scaling": "strong" if args.workload == "c5" else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl, "note": kind_note + "; the Rust crate cannot be built in this image (no cargo)"},
        "cpu_baseline": {"value": value, "unit": unit, "cores": used, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows: list[list[str]] = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def mark(self) -> int:
        return len(self.rows)

    def stop(self, start: int = 0, end: int | None = None) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        rows = self.rows[start:end] or self.rows
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks() -> dict:
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def recorded_traffic(tag: str, units: float, prefix: int | None = None):
    """DRAM bytes per launch from the committed ncu capture of this workload's kernel -- only if
    that capture was taken on a build of the very sources that are loaded now (profiles/traffic.json
    holds bytes per LP, or per launch of a pivot prefix, with the source hash and the hash of the
    library binary the capture ran on)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        for e in t.get("captures", []):
            same = e.get("src_sha16") == _src_sha16(tag) or e.get("lib_sha16") == _lib_sha16()
            if same and e.get("tag") == tag:
                if e.get("prefix") is not None:
                    return e["dram_bytes_per_unit"] if prefix == e["prefix"] else None
                return e["dram_bytes_per_unit"] * units
    except Exception:
        pass
    return None


# --------------------------------------------------------------------------- GPU arm
def run_gpu(args) -> None:
    import torch

    from dantzig_b200 import Batch, Template, measure_fp64_peak
    from dantzig_b200.model import model_from_theta
    from dantzig_b200.sharding import gather_results, shard_range
    from oracle import dzo_py  # the checker (parity sample) and the cpu_baseline leg only

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    single = args.workload in SINGLE
    metric, unit = metric_of(args.workload)

    if single:
        tmpl, theta, lower = single_workload(args.workload)
        total, lo, hi = world, rank, rank + 1          # one replica per rank
        scaling = "weak"
        batch = Batch(tmpl, 1, device=local, max_pivots=args.prefix)
        wname = f"{args.workload}: {SINGLE[args.workload]}, first {args.prefix} pivots"
        structure = None
    else:
        m, n, _, default_total, _, n_check = BATCHED[args.workload]
        strong = args.workload == "c5"
        total = (args.total_lps or default_total) if strong else world * (args.total_lps or default_total)
        lo, hi = shard_range(total, rank, world)
        w = batched_workload(args.workload, hi - lo, lo)
        theta, structure = w.theta, w.structure
        tmpl = Template(structure)
        scaling = "strong" if strong else "weak"
        batch = Batch(tmpl, hi - lo, device=local, basis_home=args.basis_home, worker_warps=args.worker_warps,
                      ctas_per_sm=args.ctas_per_sm, numerics=args.numerics)
        wname = (f"{w.name}: {total} independent LPs in total ({hi - lo} on this rank), m={m} n={n} "
                 f"(lowered {tmpl.m}x{tmpl.n_int}), all <= rows, nonneg vars, seed 1234")
    units_rank = hi - lo
    pinned = torch.empty(theta.shape, dtype=torch.float64).pin_memory()
    pinned.numpy()[...] = theta
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    warm = max(args.warmup, 3)
    batch.upload_ptr(pinned.data_ptr())
    for _ in range(warm):
        batch.solve()
        batch.sync()
    sampler = ClockSampler(local) if rank == 0 else None
    time.sleep(0.3)
    barrier()
    c0 = sampler.mark() if sampler else 0
    kernel_ms, e2e_s = [], 0.0
    res = None
    launches = 0
    for _ in range(args.steps):
        flush.zero_()                          # evict the previous step's working set from L2
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        batch.upload_ptr(pinned.data_ptr())    # H2D of the step's inputs from pinned host memory
        batch.solve()
        res = batch.download(light=True)       # D2H of status / objective / values + sync
        if dist is not None:                   # the one collective: final gather of per-LP results
            gather_results({"status": res.status, "objective": res.objective, "pivots": res.pivots},
                           total, dist, dev)
        torch.cuda.synchronize()
        e2e_s += time.perf_counter() - t0
        kernel_ms.append(batch.kernel_ms())    # cudaEvents on the launch stream, around the solve only
        launches += batch.launches()
    barrier()
    c1 = sampler.mark() if sampler else 0
    pivots_rank = int(res.pivots.sum())
    per_step_units = pivots_rank if single else units_rank
    t = torch.tensor([float(sum(kernel_ms)) * 1e-3, e2e_s, float(per_step_units)], dtype=torch.float64, device=dev)
    tmax, tsum = t.clone(), t.clone()
    if dist is not None:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    units_all = float(tsum[2].item())
    value = units_all * args.steps / float(tmax[0].item())
    e2e_value = units_all * args.steps / float(tmax[1].item())
    full = gather_results({"status": res.status}, total, dist, dev)
    h2d = theta.nbytes
    d2h = sum(a.nbytes for a in (res.status, res.pivots, res.n_primal, res.trace_hash, res.objective,
                                 res.values, res.work))
    clocks = sampler.stop(c0, c1) if sampler else None

    # ---- parity of this very run against the oracle (rank 0's shard, strided sample) ----
    parity = None
    if rank == 0:
        if single:
            cap = min(args.prefix, args.parity_prefix)
            b2 = Batch(tmpl, 1, device=local, max_pivots=cap)
            b2.upload(theta)
            b2.solve()
            r2 = b2.download(light=True)
            b2.close()
            o = lower().solve(dzo_py.SPARSE, max_pivots=cap)
            ok = (int(r2.status[0]), int(r2.pivots[0]), int(r2.trace_hash[0]), _bits(r2.objective[0])) == (
                o.status, o.pivots, o.trace_hash, _bits(o.objective))
            parity = {"checked": 1, "mismatches": 0 if ok else 1,
                      "what": f"status, pivot count, trace hash, objective bits of the first {cap} pivots vs the sparse-row oracle"}
        elif args.numerics == "fast":
            # opt-in fast numerics: NOT bit parity.  Agreement with the exact-skip oracle on the sample:
            # same status, objective within 1e-9 relative where both are optimal, pivot-count deltas.
            idx = list(range(0, units_rank, max(1, units_rank // n_check)))[:n_check]
            bad, rel_max, deltas = 0, 0.0, []
            for i in idx:
                o = dzo_py.lower(model_from_theta(structure, theta[i])).solve(dzo_py.SKIP)
                deltas.append(int(res.pivots[i]) - int(o.pivots))
                if int(res.status[i]) != o.status:
                    bad += 1
                elif o.status == 0:
                    rel = abs(res.objective[i] - o.objective) / max(1.0, abs(o.objective))
                    rel_max = max(rel_max, rel)
                    bad += rel > 1e-9
            parity = {"checked": len(idx), "mismatches": int(bad), "max_rel_objective_error": rel_max,
                      "pivot_count_deltas": deltas,
                      "what": "FAST numerics (opt-in, not bit parity): status equal and objective within 1e-9 "
                              "relative vs the exact-skip oracle; pivot-count deltas listed"}
        else:
            idx = list(range(0, units_rank, max(1, units_rank // n_check)))[:n_check]
            bad = 0
            for i in idx:
                o = dzo_py.lower(model_from_theta(structure, theta[i])).solve(dzo_py.SKIP)
                bad += (int(res.status[i]), int(res.pivots[i]), int(res.trace_hash[i]), _bits(res.objective[i])) != (
                    o.status, o.pivots, o.trace_hash, _bits(o.objective))
            parity = {"checked": len(idx), "mismatches": int(bad),
                      "what": "status, pivot count, trace hash, objective bits vs the exact-skip oracle"}

    if rank == 0:
        info = batch.launch_info()
        ms_step = float(tmax[0].item()) * 1e3 / args.steps
        flops_exec = float(res.work[:, :4].sum())         # executed (zero-skipped) flops, this rank, one step
        M, Nn = tmpl.m, tmpl.n_int - tmpl.m
        peaks = measured_peaks()
        hbm_peak = peaks.get("hbm_gbs", peaks.get("hbm_GBs", 6650.0))
        if single:
            # HBM-bound phases of the single-LP kernel: the working core is cleared and filled
            # once per solve (16 B per double), pricing streams 12 B per priced entry, the
            # vectors 48 B per row/column per pivot.  Counted by the kernel itself (work[4..]).
            core_doubles = float(res.work[0, 4])
            alg_bytes = 16.0 * core_doubles + 12.0 * float(res.work[0, 2]) / 2.0 + 48.0 * (M + Nn) * pivots_rank
            ach = alg_bytes / (ms_step * 1e-3) / 1e9
            roof = {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                    "traffic": recorded_traffic(args.workload, 1, args.prefix),
                    "algorithmic_bytes_per_launch": alg_bytes,
                    "note": "algorithmic bytes = 16 B x doubles of working core cleared+filled over all solves "
                            "+ 12 B x priced entries + 48 B x (m_int + n_nonbasic) x pivots; the kernel is bound by "
                            "dependent latency (grid barriers, ordered subtraction chains), not by bandwidth",
                    "elimination_steps": float(res.work[0, 5]), "grid_wide_steps": float(res.work[0, 6]),
                    "flops_executed_per_launch": flops_exec}
        else:
            mul_sub, fma = measure_fp64_peak(local)
            ach_tf = flops_exec / (ms_step * 1e-3) / 1e12
            nnz_n = tmpl.nnz * Nn / tmpl.n_int
            flops_literal = pivots_rank * (4.0 / 3.0 * M ** 3 + 4.0 * M ** 2 + 2.0 * nnz_n)
            alg_bytes = h2d + d2h
            fast = args.numerics == "fast"
            peak_tf = (fma if fast else mul_sub) / 1e3
            roof = {"bound": "fp64", "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s",
                    "frac": ach_tf / peak_tf,
                    "traffic": recorded_traffic(args.workload + ("_fast" if fast else ""), units_rank),
                    "note": ("FAST numerics: fused FP64 vector work on a shared-memory-resident k x k block per LP; "
                             "peak = fused DFMA rate measured live by dz_measure_fp64_peak; achieved = executed flops "
                             "(2 k^3 per Gauss-Jordan inversion + solves + pricing + updates) / kernel time; the "
                             "elimination is bound by shared-memory bandwidth and barrier latency, see DESIGN.md 4.5"
                             if fast else
                             "the exact path is un-fused FP64 vector work (no tensor/HBM bound applies); peak = "
                             "un-fused DMUL+DSUB rate measured live by dz_measure_fp64_peak (fused DFMA rate "
                             f"{fma / 1e3:.1f} TFLOP/s); achieved = EXECUTED flops (exact-zero work skipped) / kernel time"),
                    "flops_executed_per_launch": flops_exec,
                    "flops_literal_reference_per_launch": flops_literal,
                    "hbm": {"achieved": alg_bytes / (ms_step * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                            "frac": alg_bytes / (ms_step * 1e-3) / 1e9 / hbm_peak,
                            "algorithmic_bytes_per_launch": alg_bytes}}
        line = {
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": wname,
                "numerics": args.numerics + (" (OPT-IN, not the parity path: one factorisation per pivot reused for "
                                             "BTRAN, FMA; results agree with the reference to rounding only)"
                                             if args.numerics == "fast" else
                                             " (default: the reference's operations in the reference's order)"),
                "sharding": ("one replica per rank (a single LP does not shard)" if single else
                             f"contiguous LP-id ranges of the {total}-LP batch per rank; no data-path collective; "
                             "all_gather of status/objective/pivots inside the e2e region"),
                "l2": f"256 MB flush before every timed step (inputs {theta.nbytes / 1e6:.0f} MB)",
                "launch": info, "lib_sha16": _lib_sha16(), "src_sha16": _src_sha16(),
                "status_hist": np.bincount(full["status"], minlength=5).tolist(),
                "pivots_per_step_rank0": pivots_rank,
            },
            "roofline": roof,
            "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches,
            "clocks": clocks,
            "parity": parity,
        }
        if world == 1:
            line["cpu_baseline"] = cpu_baseline(args)
        print(json.dumps(line))
    batch.close()
    if dist is not None:
        dist.destroy_process_group()


def cpu_baseline(args) -> dict:
    """The oracle on the box's host cores, on a bounded sample of the same workload (reported
    baseline, not the target)."""
    import multiprocessing as mp

    from oracle import dzo_py

    dzo_py.build()
    metric, unit = metric_of(args.workload)
    if args.workload in SINGLE:
        prefix = min(args.prefix, 60)
        dt, piv = cpu_single_prefix(args.workload, prefix)
        return {"value": piv / dt, "unit": unit, "cores": 1, "kind": "port",
                "sample": f"first {piv} pivots, sparse-row oracle on one core, {dt:.1f} s (the reference's literal "
                          "dense algorithm cannot run this size)"}
    cores = _host_cores()
    per_core = BATCHED[args.workload][4] * (4 if args.workload == "c2" else 1)
    with mp.get_context("fork").Pool(cores) as pool:
        dt, nlp, piv = cpu_batched_pass(args.workload, per_core, cores, pool)
    return {"value": nlp / dt, "unit": unit, "cores": cores, "kind": "port",
            "sample": f"first {nlp} LPs of the batch ({per_core} per core), literal oracle (two dense LUs per pivot), "
                      f"{dt:.1f} s wall, {piv} pivots"}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c5", choices=["c5", "c2", "c3", "c4"])
    ap.add_argument("--total-lps", type=int, default=0, help="c5: total batch (fixed as N grows); c2: LPs per rank")
    ap.add_argument("--prefix", type=int, default=200, help="c3/c4: pivots per step")
    ap.add_argument("--parity-prefix", type=int, default=40, help="c3/c4: pivots compared with the oracle")
    ap.add_argument("--basis-home", type=int, default=0)
    ap.add_argument("--worker-warps", type=int, default=0)
    ap.add_argument("--ctas-per-sm", type=int, default=0)
    ap.add_argument("--numerics", default="exact", choices=["exact", "fast"],
                    help="fast = the opt-in fast-numerics kernel (batched workloads only), reported separately")
    args = ap.parse_args()
    if args.numerics == "fast" and args.workload in SINGLE:
        ap.error("--numerics fast covers the batched workloads (c2, c5) only")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
