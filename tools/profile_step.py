"""Run ONE pass of a workload's device program with plain launches (no CUDA graph), for
ncu:   ncu ... python tools/profile_step.py grid10x10 [rows] [passes]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from sorobn_b200 import engine, planner, workloads  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "grid10x10"
wl = workloads.WORKLOADS[name]()
rows = int(sys.argv[2]) if len(sys.argv) > 2 else wl.default_rows
passes = int(sys.argv[3]) if len(sys.argv) > 3 else 1
bn = wl.build()
net = bn._compiled
plan = planner.build_plan(net, [net.index[q] for q in wl.query], [net.index[e] for e in wl.evidence])
prog = engine.Program(plan)
prog.set_graph(False)
if os.environ.get("SBN_PLAIN"):
    prog.set_tiled(False)
if os.environ.get("SBN_CHAIN"):
    prog.set_tiled(7)
codes = wl.codes(bn, rows, seed=1000)
d_ev = torch.from_numpy(codes).cuda()
d_out = torch.empty((prog.Q, rows), dtype=torch.float32, device="cuda")
for _ in range(passes):
    prog.run_device(d_ev.data_ptr(), rows, rows, d_out.data_ptr(), rows, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
s = d_out.sum(dim=0)
print("rows", rows, "sum min/max", float(s.min()), float(s.max()), "launches", prog.info()["launches"])
