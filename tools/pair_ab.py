"""A/B of the fused launches (csrc/sbn_pair.cu: a step and its consumer as one kernel) against one launch per
step on one workload, device-resident codes, CUDA-event timing of 10 replays after 3 warm-ups.
   python tools/pair_ab.py grid10x10 [rows]            PAIR_STEPS=1: also print the per-step profile of both"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from sorobn_b200 import engine, planner, workloads  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "grid10x10"
wl = workloads.WORKLOADS[name]()
rows = int(sys.argv[2]) if len(sys.argv) > 2 else wl.default_rows
bn = wl.build()
net = bn._compiled
plan = planner.build_plan(net, [net.index[q] for q in wl.query], [net.index[e] for e in wl.evidence])
prog = engine.Program(plan)
prog.reserve(rows)
codes = wl.codes(bn, rows, seed=1000)
d_ev = torch.from_numpy(codes).cuda()
d_out = torch.empty((prog.Q, rows), dtype=torch.float32, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
res = {}
for mode, label in ((11, "paired"), (10, "single")):
    prog.set_tiled(mode)
    for _ in range(3):
        prog.run_device(d_ev.data_ptr(), rows, rows, d_out.data_ptr(), rows, stream)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        prog.run_device(d_ev.data_ptr(), rows, rows, d_out.data_ptr(), rows, stream)
    e.record()
    torch.cuda.synchronize()
    res[label] = d_out.cpu().numpy().copy()
    print(f"{name} rows={rows} {label:9s} {s.elapsed_time(e) / 10:8.3f} ms/step  info={prog.info()}")
d = np.abs(res["paired"] - res["single"]) / np.maximum(res["single"], 1e-30)
print("max rel diff paired vs single", float(d.max()), "finite", bool(np.isfinite(res["paired"]).all()))
if os.environ.get("PAIR_STEPS"):
    # per-step device times (plain launches): a fused launch shows up as (time, ~0) on its two steps
    for mode, label in ((11, "paired"), (10, "single")):
        prog.set_tiled(mode)
        ms = prog.profile(d_ev.data_ptr(), rows, rows, d_out.data_ptr(), rows, stream)
        ms = prog.profile(d_ev.data_ptr(), rows, rows, d_out.data_ptr(), rows, stream)
        print(label, "per-step us:", " ".join(f"{i}:{float(x) * 1000:.0f}" for i, x in enumerate(ms) if x > 0.001))
