"""CPU-only dump of a workload's device program: per step the output scope, inputs (with the
slot each one reads and which earlier step produced it), tile class of every input
(U = no tile axis, A = axis 0, B = axis 1, C = both) and the consumer of the output.
   python tools/plan_dump.py grid10x10"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from sorobn_b200 import planner, workloads  # noqa: E402

if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "grid10x10"
    wl = workloads.WORKLOADS[name]()
    bn = wl.build()
    net = bn._compiled
    plan = planner.build_plan(net, [net.index[q] for q in wl.query], [net.index[e] for e in wl.evidence])
    card = net.card
    writer = {}
    consumer = {}
    for i, st in enumerate(plan.steps):
        for f, _, _ in st.inputs:
            if f.is_slot:
                consumer[writer[f.buf]] = i
        writer[st.out_slot] = i
    writer = {}
    for i, st in enumerate(plan.steps):
        ins = []
        for f, es, ss in st.inputs:
            n = int(np.prod([card[v] for v in f.vars])) if f.vars else 1
            cls = "C" if (len(ss) > 1 and ss[0] and ss[1]) else "A" if (ss and ss[0]) else "B" if (len(ss) > 1 and ss[1]) else "U"
            src = f"s{writer[f.buf]}" if f.is_slot else f"T{f.buf}"
            ins.append(f"{'b' if f.batched else 't'}{n}{'e%d' % len(f.ev) if f.ev else ''}:{cls}<{src}")
        writer[st.out_slot] = i
        print(f"{i:>3} {'bat ' if st.kind else 'flat'} out={int(np.prod(st.cards)) if st.cards else 1:>5} cards={list(st.cards)} "
              f"elim={list(st.ecards)} -> s{consumer.get(i, 'post')}  | " + "  ".join(ins))
    print("bytes/row", plan.bytes_per_row(), "scratch floats/row", plan.scratch_floats_per_row())
