"""Per-step table of a workload's device program: shape, algorithmic bytes, time, GB/s.
   python tools/step_table.py grid10x10 [rows] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from sorobn_b200 import engine, planner, workloads  # noqa: E402

if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "grid10x10"
    wl = workloads.WORKLOADS[name]()
    rows = int(sys.argv[2]) if len(sys.argv) > 2 else wl.default_rows
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    bn = wl.build()
    net = bn._compiled
    plan = planner.build_plan(net, [net.index[q] for q in wl.query], [net.index[e] for e in wl.evidence])
    prog = engine.Program(plan)
    codes = wl.codes(bn, rows, seed=1000)
    d_ev = torch.from_numpy(codes).cuda()
    d_out = torch.empty((prog.Q, rows), dtype=torch.float32, device="cuda")
    ms = None
    for r in range(reps + 1):
        t = prog.profile(d_ev.data_ptr(), rows, rows, d_out.data_ptr(), rows, torch.cuda.current_stream().cuda_stream)
        if r:
            ms = t if ms is None else np.minimum(ms, t)
    per = plan.step_bytes_per_row()
    card = net.card
    tot = 0.0
    print(f"{'step':>4} {'kind':>5} {'n_in':>4} {'out':>7} {'elim':>5} {'in sizes':<34} {'B/row':>8} {'us':>8} {'GB/s':>7}")
    for i, st in enumerate(plan.steps):
        outn = int(np.prod(st.cards)) if len(st.cards) else 1
        eliml = st.cx
        ins = ",".join(("b" if f.batched else "t") + str(int(np.prod([card[v] for v in f.vars]))) + ("e%d" % len(f.ev) if f.ev else "")
                       for f, _, _ in st.inputs)
        b = per[i] if i < len(per) else 0
        gbs = b * rows / (ms[i] * 1e-3) / 1e9 if ms[i] > 0 else 0
        tot += ms[i]
        print(f"{i:>4} {('bat' if st.kind else 'flat'):>5} {len(st.inputs):>4} {outn:>7} {eliml:>5} {ins:<34} {b:>8} {ms[i]*1e3:>8.1f} {gbs:>7.0f}")
    print("normalise us", ms[-1] * 1e3, "total ms", tot + ms[-1])
