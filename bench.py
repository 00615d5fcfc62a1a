#!/usr/bin/env python
"""Benchmark of the exact-inference hot path (see the contract in the task brief).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload grid10x10|asia_1m|dag50]
                    [--rows R] [--impl b200|reference]

A *step* is one pass of the hot path over one batch of synthetic evidence rows:
`rows` independent exact-inference queries (same query variables, same evidence
variables, different observed states) per GPU.  Multi-GPU = one process per GPU
(torchrun), evidence rows sharded across ranks (weak scaling: `rows` per GPU), the only
collective is the final gather of the posteriors on rank 0 (NCCL), inside the timed step.

Rank 0 prints ONE JSON line:
  value      rows/s over all GPUs with evidence codes already resident in HBM
  e2e        the same metric through the C ABI with HOST (pinned) buffers: H2D of the
             evidence codes, every kernel, D2H of the posteriors, per step
  roofline   algorithmic HBM bytes of the step kernels / their device time vs measured peak
  cpu_baseline  the CPU oracle (numpy port of the reference algorithm) on a bounded sample

`--impl reference` times the CPU arm instead (the oracle port of the reference's
variable elimination, one process per host core, bounded sample per step).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "exact-inference queries/sec"
UNIT = "queries/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="grid10x10")
    ap.add_argument("--rows", type=int, default=0, help="evidence rows per GPU per step (0 = workload default)")
    ap.add_argument("--cpu-rows", type=int, default=0, help="rows of the CPU sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dump", default="", help="write per-launch timings (JSON) here")
    return ap.parse_args()


# ------------------------------------------------------------------ clocks sampling
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU through NVML every ~2 ms.  The thread
    is started early (NVML init takes longer than the timed region); `begin()` / `end()`
    bracket the timed region and only samples taken in between are reported."""

    def __init__(self, index: int):
        self.index = index
        self.sm, self.reasons_seen = [], 0
        self.max_sm = None
        self._stop = threading.Event()
        self._ready = threading.Event()
        self._active = False
        self._ok = False
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()
        self._ready.wait(timeout=20)

    def _run(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self._ok = True
            self._ready.set()
            while not self._stop.is_set():
                if self._active:
                    self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                    try:
                        self.reasons_seen |= int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                    except Exception:
                        self.reasons_seen |= int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                self._stop.wait(0.002)
        except Exception:
            self._ok = False
            self._ready.set()

    def begin(self):
        self.sm, self.reasons_seen = [], 0
        self._active = True

    def end(self):
        self._active = False

    def close(self):
        self._stop.set()
        self._thread.join(timeout=6)

    def summary(self):
        if not self._ok or not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_sm, "reasons": [], "samples": 0}
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                "hw_power_brake_slowdown": 0x80}
        reasons = [n for n, b in bits.items() if self.reasons_seen & b]
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.max_sm, "reasons": reasons,
                "samples": len(self.sm)}


# --------------------------------------------------------------------- CPU baseline
_CPU_STATE = {}  # per worker process: workload -> (compiled net, dense oracle net, evidence names, order)


def _cpu_state(workload):
    if workload not in _CPU_STATE:
        from oracle import ve_oracle
        from sorobn_b200 import planner, workloads

        wl = workloads.WORKLOADS[workload]()
        bn = wl.build()
        net = bn._compiled
        dn = ve_oracle.dense_from_pandas(bn.P, bn.parents, bn.nodes)
        plan = planner.build_plan(net, [net.index[q] for q in wl.query], [net.index[e] for e in wl.evidence])
        _CPU_STATE[workload] = (wl, net, dn, [net.names[v] for v in plan.order])
    return _CPU_STATE[workload]


def _cpu_worker(args):
    """Answer rows [lo, hi) with the CPU oracle; network / plan setup is cached per process
    (the warm-up map pays for it), so the timed map measures inference only."""
    workload, codes, lo, hi = args
    from oracle import ve_oracle

    wl, net, dn, order = _cpu_state(workload)
    t = time.perf_counter()
    acc = 0.0
    for b in range(lo, hi):
        ev = {v: net.domains[net.index[v]][codes[i, b]] for i, v in enumerate(wl.evidence)}
        acc += float(ve_oracle.query(dn, *wl.query, event=ev, order=order)[1].reshape(-1)[0])
    return time.perf_counter() - t, acc


def _cpu_warm(workload):
    _cpu_state(workload)
    time.sleep(0.2)
    return os.getpid()


def cpu_rate(workload: str, codes: np.ndarray, n_rows: int, n_procs: int):
    """Queries/s of the CPU oracle (numpy restatement of bayes_net.py:739-794, same
    min-fill order as the device program) on `n_rows` rows with `n_procs` processes."""
    import multiprocessing as mp

    n_rows = min(n_rows, codes.shape[1])
    if n_procs <= 1:
        _cpu_state(workload)
        t0 = time.perf_counter()
        _cpu_worker((workload, codes, 0, n_rows))
        return n_rows / (time.perf_counter() - t0)
    bounds = np.linspace(0, n_rows, n_procs + 1).astype(int)
    jobs = [(workload, codes, int(bounds[i]), int(bounds[i + 1])) for i in range(n_procs) if bounds[i + 1] > bounds[i]]
    # one single-threaded worker per core: numpy's own thread pools would oversubscribe
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "NUMEXPR_NUM_THREADS"):
        os.environ[var] = "1"
    ctx = mp.get_context("spawn")
    with ctx.Pool(len(jobs)) as pool:
        # warm every worker (imports, network build, plan): chunksize 1 and as many tasks as
        # workers, each sleeping briefly so that no worker takes two
        pool.map(_cpu_warm, [workload] * len(jobs), chunksize=1)
        t0 = time.perf_counter()
        pool.map(_cpu_worker, jobs, chunksize=1)
        dt = time.perf_counter() - t0
    return n_rows / dt


def effective_cores() -> int:
    """Host cores this process may actually use: the smallest of the CPU count, the scheduler
    affinity mask and the cgroup CPU quota (containers often expose 128 CPUs with a quota of 8)."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            with open(path) as f:
                parts = f.read().split()
            if path.endswith("cpu.max"):
                if parts[0] != "max":
                    n = min(n, max(1, int(int(parts[0]) / int(parts[1]))))
            else:
                quota = int(parts[0])
                if quota > 0:
                    with open("/sys/fs/cgroup/cpu/cpu.cfs_period_us") as g:
                        n = min(n, max(1, quota // int(g.read().split()[0])))
            break
        except Exception:
            continue
    return max(1, n)


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload: str):
    """Per-launch DRAM traffic of the dominant kernel from the committed ncu capture."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                return json.load(f).get(workload)
        except Exception:
            return None
    return None


# ------------------------------------------------------------------ reference arm
def run_reference(args, rank, world):
    if rank != 0:
        return
    from sorobn_b200 import workloads

    wl = workloads.WORKLOADS[args.workload]()
    bn = wl.build()
    cores = effective_cores()
    sample = args.cpu_rows or {"grid10x10": 512 * cores, "asia_1m": 20000 * cores, "dag50": 256 * cores}.get(args.workload, 256 * cores)
    codes = wl.codes(bn, sample, seed=0)
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_rate(args.workload, codes, max(cores, sample // 8), cores)
    rates = [cpu_rate(args.workload, codes, sample, cores) for _ in range(max(1, args.steps))]
    value = float(np.mean(rates))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sample / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl.name, "description": wl.description, "rows_per_step": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} evidence rows of the same workload per step, {cores} processes "
                                   "(numpy oracle port of the reference's variable elimination, min-fill order)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------- B200 arm
def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from sorobn_b200 import engine, planner, workloads

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world > 1
    if distributed and not dist.is_initialized():
        # stdout carries the one JSON line: NCCL prints its version banner to fd 1 when the
        # communicator is created, so fd 1 points at stderr until that has happened
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    wl = workloads.WORKLOADS[args.workload]()
    bn = wl.build(device=local_rank)
    net = bn._compiled
    plan = planner.build_plan(net, [net.index[q] for q in wl.query], [net.index[e] for e in wl.evidence])
    prog = engine.Program(plan, device=local_rank)
    rows = args.rows or wl.default_rows
    prog.reserve(rows)
    reserved = prog.info()["reserved_rows"]
    assert reserved >= rows, f"scratch for {rows} rows does not fit (got {reserved})"

    n_ev, Q = prog.n_ev, prog.Q
    codes_host = engine.PinnedArray((max(n_ev, 1), rows), np.uint8)
    codes_host.array[:n_ev] = wl.codes(bn, rows, seed=1000 + rank)
    out_host = engine.PinnedArray((Q, rows), np.float32)
    d_ev = torch.from_numpy(codes_host.array).to(dev)
    d_out = torch.empty((Q, rows), dtype=torch.float32, device=dev)
    gathered = torch.empty((world, Q, rows), dtype=torch.float32, device=dev) if distributed and rank == 0 else None
    stream = torch.cuda.current_stream().cuda_stream

    step_bytes = plan.bytes_per_row() * rows
    flush = None
    if step_bytes < 512e6:  # working set could sit in the 126 MB L2: flush between steps
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def device_step():
        prog.run_device(d_ev.data_ptr(), rows, rows, d_out.data_ptr(), rows, stream)
        if distributed:
            dist.gather(d_out, list(gathered.unbind(0)) if rank == 0 else None, dst=0)

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local_rank)
    for _ in range(args.warmup):
        device_step()
    barrier()
    launches0 = prog.info()["launches"]

    clocks.begin()
    if True:
        if flush is None:
            start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            start.record()
            for _ in range(args.steps):
                device_step()
            end.record()
            barrier()
            dev_ms = start.elapsed_time(end)
        else:
            dev_ms = 0.0
            for _ in range(args.steps):
                flush.fill_(1)
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                barrier()
                s.record()
                device_step()
                e.record()
                barrier()
                dev_ms += s.elapsed_time(e)
        launches = prog.info()["launches"] - launches0

        # ---- end to end through the C ABI with host buffers ----------------------------
        for _ in range(max(1, args.warmup // 2)):
            prog.run(codes_host.array[:n_ev], rows, out=out_host.array)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            prog.run(codes_host.array[:n_ev], rows, out=out_host.array)
        barrier()
        e2e_s = time.perf_counter() - t0
    clocks.end()
    clock_summary = clocks.summary()
    clocks.close()

    t = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_s = float(t[0]), float(t[1])

    # correctness guard of the timed outputs: every posterior sums to one
    sums = d_out.sum(dim=0)
    ok = bool(torch.isfinite(sums).all() and ((sums - 1).abs() < 1e-4).all())
    same = np.allclose(out_host.array, d_out.cpu().numpy(), rtol=0, atol=0)

    if rank == 0:
        ms_per_step = dev_ms / args.steps
        total_rows = rows * world
        value = total_rows / (ms_per_step * 1e-3)
        e2e_value = total_rows / (e2e_s / args.steps)

        # per-launch timings of one step (CUDA events around every launch, same stream)
        prof = prog.profile(d_ev.data_ptr(), rows, rows, d_out.data_ptr(), rows, stream)
        sb = plan.step_bytes_per_row()
        kernel_ms = float(sum(ms for ms, st in zip(prof[:-1], plan.steps) if st.kind == planner.KIND_BATCHED))
        kernel_bytes = float(sum(sb)) * rows
        peak, peak_src = measured_peak()
        achieved = kernel_bytes / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else 0.0
        roofline = {
            "bound": "hbm", "kernel": "sbn_step_tiled (+ sbn_step_batched on sum-out-only steps)", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "peak_source": peak_src,
            "algorithmic_bytes_per_step": kernel_bytes, "kernel_ms_per_step": kernel_ms,
            "launches_per_step": int(sum(1 for st in plan.steps if st.kind == planner.KIND_BATCHED)),
            "whole_step_frac": (plan.bytes_per_row() * rows / (ms_per_step * 1e-3) / 1e9) / peak,
            "traffic": (ncu_traffic(wl.name) or {}).get("dram_bytes_per_step") if rows == wl.default_rows else None,
            "traffic_detail": ncu_traffic(wl.name) if rows == wl.default_rows else None,
        }
        if args.dump:
            with open(args.dump, "w") as f:
                json.dump({"workload": wl.name, "rows": rows, "step_ms": [float(x) for x in prof],
                           "step_bytes_per_row": sb,
                           "steps": [{"kind": st.kind, "cx": st.cx, "cards": list(st.cards),
                                      "inputs": [("B" if fct.batched else "t") + str(int(np.prod([net.card[v] for v in fct.vars])) if fct.vars else 1)
                                                 + (f"e{len(fct.ev)}" if fct.ev else "") for fct, _, _ in st.inputs]}
                                     for st in plan.steps]}, f)

        cpu = None
        if not args.no_cpu_baseline:
            n_cpu = args.cpu_rows or {"grid10x10": 8192, "asia_1m": 100000, "dag50": 4096}.get(wl.name, 4096)
            codes_cpu = np.ascontiguousarray(codes_host.array[:n_ev, :n_cpu])
            rate = cpu_rate(wl.name, codes_cpu, n_cpu, 1)
            cpu = {"value": rate, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": f"first {n_cpu} evidence rows of the same batch, single process "
                             "(oracle/ve_oracle.py: numpy port of the reference's variable elimination, "
                             "same min-fill order)"}

        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": wl.name, "description": wl.description, "rows_per_gpu": rows,
                "global_rows": total_rows, "parallelism": f"rows sharded x{world}; NCCL gather of posteriors",
                "elimination_steps": len(plan.steps), "max_factor_entries_per_row": plan.max_factor_per_row(),
                "algorithmic_bytes_per_row": plan.bytes_per_row(),
                "l2": ("flushed (256 MB write) between timed steps" if flush is not None else
                       f"not flushed: each step streams {step_bytes / 1e9:.2f} GB of factors >> 126 MB L2"),
            },
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(n_ev * rows) * world,
                    "d2h_bytes_per_step": int(Q * rows * 4) * world, "ms_per_step": 1e3 * e2e_s / args.steps,
                    "api": "sbn_program_run_host (pinned host buffers)"},
            "gpu_launches": int(launches) * world,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "clocks": clock_summary,
            "checks": {"posteriors_sum_to_one": ok, "host_path_equals_device_path": bool(same)},
        }
        print(json.dumps(line), flush=True)

    if distributed:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
