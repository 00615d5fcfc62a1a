"""BASELINE.json configs[4]: Gibbs sampling on the 100-node 5-state grid, 10k chains x 10k
iterations per GPU (one chain per evidence row).  Prints one JSON line; wall-clock through
the C ABI with host buffers (evidence H2D, kernel, frequencies D2H)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from sorobn_b200 import engine, workloads  # noqa: E402

if __name__ == "__main__":
    chains = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
    wl = workloads.grid10x10()
    bn = wl.build()
    net = bn._compiled
    nonevents = sorted(set(bn.nodes) - set(wl.evidence))
    sampler = engine.GibbsSampler(net, [net.index[q] for q in wl.query], [net.index[e] for e in wl.evidence],
                                  [net.index[v] for v in nonevents])
    codes = wl.codes(bn, chains, seed=3)
    sampler.run(codes, chains, 100, 1)  # warm-up
    times = []
    for rep in range(3):
        t = time.perf_counter()
        freq = sampler.run(codes, chains, iters, 1234 + rep)
        times.append(time.perf_counter() - t)
    exact = bn.query_many(*wl.query, events=wl.events(chains, seed=3, bn=bn)).to_numpy()
    err = float(np.abs(freq.T - exact).mean())
    dt = min(times)
    print(json.dumps({"workload": "grid10x10 gibbs", "chains": chains, "iterations": iters, "seconds": dt,
                      "variable_updates_per_s": chains * iters / dt, "chains_x_iterations": chains * iters,
                      "mean_abs_error_vs_exact": err, "n_gpus": 1}))
