"""N > 1 path on CPU: two gloo ranks each own a contiguous block of sweep points, solve it
independently (the CPU oracle stands in for the per-rank solver here) and rank 0 gathers the blocks;
the gathered result must equal the single-process result."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cedarsim.jl_b200 import circuits, multi  # noqa: E402
from cedarsim.jl_b200.flat import params_matrix  # noqa: E402


def _sweep():
    fc = circuits.two_resistor()
    r1, r2 = np.meshgrid(np.arange(100, 1101, 100.0), np.arange(100, 701, 100.0), indexing="ij")
    return fc, params_matrix([r1.ravel(order="F"), r2.ravel(order="F")])


def _worker(rank, world, port, q):
    from oracle import orc
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fc, P = _sweep()
    B = P.shape[1]
    lo, hi = multi.partition(B, world, rank)
    x, _, st, _ = orc.dc(fc, np.ascontiguousarray(P[:, lo:hi]))
    sizes = [multi.partition(B, world, r)[1] - multi.partition(B, world, r)[0] for r in range(world)]
    full = multi.gather_blocks(torch.from_numpy(x), world, rank, sizes, axis=1)
    if rank == 0:
        q.put(full.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_partition_covers_all_points():
    for B, W in ((77, 2), (16384, 8), (5, 8), (1, 1)):
        blocks = [multi.partition(B, W, r) for r in range(W)]
        assert blocks[0][0] == 0 and blocks[-1][1] == B
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(W - 1))


def test_two_rank_gather_matches_single_process():
    from oracle import orc
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    full = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    fc, P = _sweep()
    x, _, _, _ = orc.dc(fc, P)
    assert full.shape == x.shape and np.array_equal(full, x)
