"""Multi-rank host logic on the CPU: two gloo processes shard a batch of queries, each
answers its slice (with the CPU oracle standing in for the GPU engine -- tests may use
it as the checker), rank 0 gathers; the result must equal the unsharded answer."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from sorobn_b200 import sharding


def test_row_shard_partitions_exactly():
    for n in (0, 1, 2, 7, 8, 100_000, 100_003):
        for world in (1, 2, 3, 8):
            slices = [sharding.row_shard(n, r, world) for r in range(world)]
            assert slices[0].start == 0 and slices[-1].stop == n
            assert all(a.stop == b.start for a, b in zip(slices, slices[1:]))
            sizes = [s.stop - s.start for s in slices]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.row_shard(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_rows, out_path):
    import torch.distributed as dist

    from oracle import ve_oracle
    from sorobn_b200 import workloads

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        wl = workloads.asia_1m()
        bn = wl.build()
        net = bn._compiled
        dn = ve_oracle.dense_from_pandas(bn.P, bn.parents, bn.nodes)
        codes = wl.codes(bn, n_rows, seed=11)  # identical on every rank (seeded)

        def run_fn(local_codes, n_local):
            out = np.zeros((2, n_local), dtype=np.float32)
            for b in range(n_local):
                ev = {v: net.domains[net.index[v]][local_codes[i, b]] for i, v in enumerate(wl.evidence)}
                out[:, b] = ve_oracle.query(dn, *wl.query, event=ev)[1].reshape(-1)
            return out

        got = sharding.run_sharded(codes, n_rows, run_fn)
        if rank == 0:
            full = run_fn(codes, n_rows)
            assert got.shape == (2, n_rows)
            assert np.array_equal(got.numpy(), full)
            np.save(out_path, got.numpy())
        else:
            assert got is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_rows", [1, 2, 37])
def test_two_rank_gloo_sharding_matches_single_rank(tmp_path, n_rows):
    world = 2
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(world, _free_port(), n_rows, out), nprocs=world, join=True)
    got = np.load(out)
    assert got.shape == (2, n_rows)
    assert np.allclose(got.sum(axis=0), 1.0, atol=1e-6)


class _OracleProgram:
    """Stand-in for engine.Program on the CPU ranks (tests may use the oracle as the engine)."""

    def __init__(self):
        from oracle import ve_oracle
        from sorobn_b200 import workloads

        self.wl = workloads.asia_1m()
        self.bn = self.wl.build()
        self.net = self.bn._compiled
        self.dn = ve_oracle.dense_from_pandas(self.bn.P, self.bn.parents, self.bn.nodes)
        self.ve = ve_oracle

    def run(self, codes, n):
        out = np.zeros((2, n), dtype=np.float32)
        for b in range(n):
            ev = {v: self.net.domains[self.net.index[v]][codes[i, b]] for i, v in enumerate(self.wl.evidence)}
            out[:, b] = self.ve.query(self.dn, *self.wl.query, event=ev)[1].reshape(-1)
        return out


def _sharded_program_worker(rank, world, port, out_path):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        prog = _OracleProgram()
        counts = [9, 6]  # ragged: rank 1 has fewer rows than rows_max
        sp = sharding.ShardedProgram(prog, Q=2, n_ev=4, rows_max=max(counts), dst=0)
        codes = prog.wl.codes(prog.bn, counts[rank], seed=100 + rank)  # every rank has its own rows
        for _ in range(2):  # buffers are reused between calls
            got = sp.run_host(codes, counts[rank], counts=counts)
        parts = sp.run_host(codes, counts[rank], counts=counts, blocks=True)  # per-rank views, no host concatenation
        if rank == 0:
            want = np.concatenate([prog.run(prog.wl.codes(prog.bn, counts[r], seed=100 + r), counts[r])
                                   for r in range(world)], axis=1)
            assert got.shape == (2, sum(counts)) and np.array_equal(got, want)
            assert [p.shape for p in parts] == [(2, c) for c in counts] and np.array_equal(np.concatenate(parts, axis=1), want)
            np.save(out_path, got)
        else:
            assert got is None and parts is None
    finally:
        dist.destroy_process_group()


def test_sharded_program_gathers_ragged_rank_batches(tmp_path):
    """The torchrun path bench.py --gpus N drives (sharding.ShardedProgram), over gloo with the
    oracle as the per-rank engine: every rank answers its own rows, rank 0 gets them in rank order."""
    out = str(tmp_path / "sp.npy")
    mp.spawn(_sharded_program_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    assert got.shape == (2, 15) and np.allclose(got.sum(axis=0), 1.0, atol=1e-6)
