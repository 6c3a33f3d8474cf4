"""The N>1 path of bench.py on CPU: world_size-2 gloo processes shard the canonical job stream
with no data-path collective, and the timing reduction is max-over-ranks / sum-of-units."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    from llmq_b200.fixtures import make_jobs

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    jobs = make_jobs(10, 1024, prompt_tokens=4, start=rank, stride=world)
    units = torch.tensor([float(len(jobs))])
    t = torch.tensor([1.0 + rank])  # rank 1 is the slow one
    dist.barrier()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(units, op=dist.ReduceOp.SUM)
    ids = [None] * world
    dist.all_gather_object(ids, [j["id"] for j in jobs])
    if rank == 0:
        out.put((t.item(), units.item(), ids))
    dist.destroy_process_group()


def test_two_rank_sharding_and_reduction():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 500
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    t, units, ids = q.get(timeout=120)
    [p.join(60) for p in procs]
    assert t == 2.0 and units == 10.0  # max over ranks; whole-job units
    flat = sorted(ids[0] + ids[1])
    assert flat == [f"job-{i:07d}" for i in range(10)] and not set(ids[0]) & set(ids[1])
