"""How does the CPU oracle arm scale with processes on this box?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from sorobn_b200 import workloads

if __name__ == "__main__":
    wl = workloads.grid10x10(); bn = wl.build()
    codes = wl.codes(bn, 32768, seed=0)
    print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), "effective", bench.effective_cores())
    for path in ("/sys/fs/cgroup/cpu.max", "/proc/loadavg"):
        try:
            print(path, open(path).read().strip())
        except Exception as e:
            print(path, "n/a", e)
    for procs in (1, 8, 32, 64, 128):
        if procs > (os.cpu_count() or 1): break
        n = 256 * procs
        t = time.time(); r = bench.cpu_rate("grid10x10", codes, n, procs); print(procs, "procs", round(r), "q/s", round(r / procs), "per proc", round(time.time() - t, 1), "s wall", flush=True)
