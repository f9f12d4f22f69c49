"""CPU, world_size 2 over gloo: the multi-GPU path has no data-path collective -- every rank derives
its shard from the same list of read lengths -- so what is tested is that the shards are disjoint,
complete and balanced, that per-rank synthetic slabs are reproducible from (seed, read_id) alone, and
that the only exchanges bench.py performs (barrier + MAX of the step time) work."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp
import torch


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nanoreviser_b200 import synth, workqueue
    lengths = synth.read_lengths("cfg3", 600)
    mine = workqueue.shard_for_rank(lengths.tolist(), rank, world)
    # every rank generates only its own reads; a read depends on (seed, read_id) alone
    b = synth.make_batch([int(lengths[i]) for i in mine[:3]], seed=5, first_id=0)
    digest = int(np.frombuffer(b.bases[:64].tobytes(), np.uint8).sum())
    load = torch.tensor([float(sum(int(lengths[i]) for i in mine))], dtype=torch.float64)
    gathered = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, load)
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.barrier()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)           # bench.py: max-over-ranks step time
    q.put((rank, mine, [float(g.item()) for g in gathered], float(t.item()), digest))
    dist.destroy_process_group()


def test_two_rank_sharding_over_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    shards = [r[1] for r in res]
    assert sorted(shards[0] + shards[1]) == list(range(600)) and not set(shards[0]) & set(shards[1])
    loads = res[0][2]
    assert loads == res[1][2] and max(loads) / (sum(loads) / 2) < 1.01
    assert res[0][3] == res[1][3] == 2.0
