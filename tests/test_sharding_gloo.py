"""Multi-GPU host logic on CPU: two gloo ranks, per-rank seed offsets (45 * rank), no data-path
collective; the only collectives are the timing reduction of bench.py and an optional gather."""
import hashlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _digest(arrs):
    h = hashlib.sha256()
    for k in sorted(arrs):
        h.update(np.ascontiguousarray(arrs[k]).tobytes())
    return np.frombuffer(h.digest()[:8], dtype=np.int64)[0]


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ofdg_b200 as o
    a = o.ParamStream(7, seed_offset=45 * rank).generate(6).arrays()
    mine = torch.tensor([_digest(a)], dtype=torch.int64)
    allv = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allv, mine)
    # whole-job time = max over ranks (bench.py)
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # optional gather of the finished blobs to the training rank (ofdg_b200.gather_blobs; small stand-ins for the three blobs here)
    blobs = [torch.full((2, 3, 4, 5), float(rank)), torch.full((2, 3, 4, 5), 10.0 + rank), torch.full((2, 2, 4, 5), 20.0 + rank)]
    got = o.gather_blobs(blobs, dst=0)
    if rank == 0:
        assert [tuple(g.shape) for g in got] == [(2 * world, 3, 4, 5), (2 * world, 3, 4, 5), (2 * world, 2, 4, 5)]
        assert float(got[2][2 * (world - 1), 0, 0, 0]) == 20.0 + world - 1
        out.put(([int(v.item()) for v in allv], float(t.item()), [float(got[0][2 * r, 0, 0, 0]) for r in range(world)]))
    else:
        assert got is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding():
    sys.path.insert(0, ROOT)
    import ofdg_b200 as o
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    digests, tmax, gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert digests[0] != digests[1], "ranks must generate different samples"
    # a shard is a pure function of (mode, seed offset, sample index): any process reproduces it
    for r in range(2):
        assert digests[r] == int(_digest(o.ParamStream(7, seed_offset=45 * r).generate(6).arrays()))
    assert tmax == 2.0 and gathered == [0.0, 1.0]
