"""Does ordering the evidence rows by their codes help a workload?  (rows are independent, so any
order gives the same posteriors; lanes of a warp that share evidence codes gather the same table
entries.)   python tools/sorted_rows_ab.py dag50 [rows]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from sorobn_b200 import engine, planner, workloads  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "dag50"
wl = workloads.WORKLOADS[name]()
rows = int(sys.argv[2]) if len(sys.argv) > 2 else wl.default_rows
bn = wl.build()
net = bn._compiled
plan = planner.build_plan(net, [net.index[q] for q in wl.query], [net.index[e] for e in wl.evidence])
prog = engine.Program(plan)
prog.reserve(rows)
codes = wl.codes(bn, rows, seed=1000)
# evidence columns gathered by the biggest tables first
weight = np.zeros(len(wl.evidence))
for st in plan.steps:
    if st.kind != planner.KIND_BATCHED:
        continue
    for f, _, _ in st.inputs:
        if not f.batched and f.ev:
            size = int(np.prod([net.card[v] for v in f.vars])) * int(np.prod([c for _, _, c in f.ev]))
            for col, _, _ in f.ev:
                weight[col] += size
order_cols = np.argsort(-weight)
print("columns by gathered table size:", [(int(c), int(weight[c])) for c in order_cols[:8]])
stream = torch.cuda.current_stream().cuda_stream
d_out = torch.empty((prog.Q, rows), dtype=torch.float32, device="cuda")


def time_codes(c, label):
    d_ev = torch.from_numpy(np.ascontiguousarray(c)).cuda()
    for _ in range(3):
        prog.run_device(d_ev.data_ptr(), rows, rows, d_out.data_ptr(), rows, stream)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        prog.run_device(d_ev.data_ptr(), rows, rows, d_out.data_ptr(), rows, stream)
    e.record()
    torch.cuda.synchronize()
    print(f"{name} rows={rows} {label:28s} {s.elapsed_time(e) / 10:8.3f} ms/step")
    return d_out.cpu().numpy().copy()


base = time_codes(codes, "as generated")
for k in (2, 4, 6, len(order_cols)):
    keys = [codes[c] for c in order_cols[:k]][::-1]  # lexsort: last key is primary
    perm = np.lexsort(keys)
    out = time_codes(codes[:, perm], f"sorted by top {k} columns")
    assert np.allclose(out, base[:, perm], rtol=1e-6, atol=1e-30)
