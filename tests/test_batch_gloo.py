"""CPU: the N > 1 path of the batch driver (bce_b200/batch.py) with world_size 2 over gloo.
The path shards by independent inputs (replicas only); the one collective is the stats gather."""
import os
import socket
import sys
from pathlib import Path

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bce_b200 import batch

    sizes = [100 + 7 * i for i in range(7)]          # 7 inputs over 2 ranks: 4 + 3

    def work(i):
        # stands in for "compress input i on this rank's GPU"
        return sizes[i], sizes[i] // 3, sizes[i] * 2, 1.0 + i, 2.0 + i, f"archive-{i}"

    res = batch.run_batch(len(sizes), work, device="cpu")
    q.put((rank, sorted(res.outputs), [vars(r) for r in res.per_rank], vars(res.total)))
    dist.barrier()
    dist.destroy_process_group()


def test_round_robin_shard_is_a_partition():
    from bce_b200 import batch
    for world in (1, 2, 4, 8):
        seen = sorted(i for r in range(world) for i in batch.shard(64, r, world))
        assert seen == list(range(64))
        assert max(len(batch.shard(64, r, world)) for r in range(world)) == 64 // world


def test_two_ranks_gloo_gather_stats():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got.sort()
    sizes = [100 + 7 * i for i in range(7)]
    assert got[0][1] == [0, 2, 4, 6] and got[1][1] == [1, 3, 5]
    # every rank sees the same gathered table
    assert got[0][2] == got[1][2]
    per_rank = got[0][2]
    assert [r["inputs"] for r in per_rank] == [4, 3]
    assert per_rank[0]["bytes_in"] == sum(sizes[0::2]) and per_rank[1]["bytes_in"] == sum(sizes[1::2])
    total = got[0][3]
    assert total["inputs"] == 7 and total["bytes_in"] == sum(sizes)
    assert total["gpu_ms"] == max(r["gpu_ms"] for r in per_rank)      # ranks run concurrently
