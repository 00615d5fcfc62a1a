"""Row sharding of a batch of independent queries over the GPUs of one box.

Evidence rows are independent (each is one `BayesNet.query` call in the reference,
/root/reference/sorobn/bayes_net.py:796), so the hot path has no data-path collective:
rank r answers a contiguous slice of the rows on its own GPU.  The only exchange is the
final gather of the posteriors on the destination rank (NCCL on GPUs; the CPU tests run
the same code over gloo with world_size 2).
"""
from __future__ import annotations

import numpy as np

__all__ = ["row_shard", "gather_rows", "run_sharded"]


def row_shard(n_rows: int, rank: int, world: int) -> slice:
    """Contiguous, balanced slice of `n_rows` for `rank` (sizes differ by at most one;
    ranks beyond n_rows get empty slices)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank {rank} / world {world}")
    base, extra = divmod(int(n_rows), world)
    lo = rank * base + min(rank, extra)
    return slice(lo, lo + base + (1 if rank < extra else 0))


def gather_rows(local, n_rows: int, group=None, dst: int = 0, device=None):
    """Gather per-rank posteriors [Q, n_local] (row slices in rank order) into
    [Q, n_rows] on `dst` (a global rank, which must belong to `group`); other ranks get None.  `local` is a numpy array or a torch
    tensor; the collective runs on whatever backend `group` uses."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    t = torch.as_tensor(local)
    if device is not None:
        t = t.to(device)
    Q = t.shape[0]
    widest = row_shard(n_rows, 0, world)
    width = widest.stop - widest.start
    padded = torch.zeros((Q, width), dtype=t.dtype, device=t.device)
    padded[:, : t.shape[1]] = t
    # `dst` is a GLOBAL rank (what dist.gather takes); inside a sub-group the group-local rank differs
    is_dst = dist.get_rank() == dst
    bufs = [torch.empty_like(padded) for _ in range(world)] if is_dst else None
    dist.gather(padded, bufs, dst=dst, group=group)
    if not is_dst:
        return None
    parts = []
    for r in range(world):
        sl = row_shard(n_rows, r, world)
        parts.append(bufs[r][:, : sl.stop - sl.start])
    return torch.cat(parts, dim=1)


def run_sharded(codes: np.ndarray, n_rows: int, run_fn, group=None, dst: int = 0, device=None):
    """Answer `n_rows` queries across the ranks of `group`.

    codes  : uint8 [n_ev, n_rows] evidence codes, identical on every rank (each rank only
             reads its slice).
    run_fn : (codes_slice [n_ev, n_local], n_local) -> posterior [Q, n_local]; in
             production `engine.Program.run` / `run_device`, in the CPU tests a stand-in.
    Returns [Q, n_rows] on `dst`, None elsewhere.
    """
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sl = row_shard(n_rows, rank, world)
    n_local = sl.stop - sl.start
    local_codes = np.ascontiguousarray(codes[:, sl])
    if n_local > 0:
        local = run_fn(local_codes, n_local)
    else:
        local = None
    # every rank needs Q to build its (possibly empty) contribution
    import torch

    q = torch.tensor([0 if local is None else int(local.shape[0])], dtype=torch.int64,
                     device=device if device is not None else "cpu")
    dist.all_reduce(q, op=dist.ReduceOp.MAX, group=group)
    Q = int(q.item())
    if local is None:
        local = np.zeros((Q, 0), dtype=np.float32)
    return gather_rows(local, n_rows, group=group, dst=dst, device=device)
