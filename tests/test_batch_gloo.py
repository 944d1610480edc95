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


def _file_worker(rank, world, port, q, tmp):
    sys.path.insert(0, str(ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bce_b200 import batch
    paths = [os.path.join(tmp, f"in{i}") for i in range(9)]          # in4 does not exist, in7 is empty

    def fake_compress(data, cfg):
        # stands in for the GPU front end + host coders: the test is about the queue, the failures, the manifest
        import time
        time.sleep(0.01 * (1 + rank))                                  # the ranks run at different speeds
        return b"BCE" + bytes([int(data[0])]) + bytes(len(data) // 4), 2.5

    res = batch.compress_files(paths, os.path.join(tmp, "out"), fake_compress, device="cpu")
    q.put((rank, [vars(f) for f in res.files], [vars(r) for r in res.per_rank], vars(res.total)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_compress_files_shared_queue_and_failures(tmp_path):
    """configs[4] shape on the CPU: 9 files, 2 workers pulling from the shared queue, one archive per file, a missing
    and an empty file recorded as failures while the batch goes on, the manifest gathered on every rank."""
    for i in range(9):
        if i == 4:
            continue
        (tmp_path / f"in{i}").write_bytes(bytes([i]) * (0 if i == 7 else 1000 + 100 * i))
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_file_worker, args=(r, world, port, q, str(tmp_path))) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0][1] == got[1][1] and got[0][2] == got[1][2]            # every rank holds the same manifest and stats
    files = got[0][1]
    assert [f["index"] for f in files] == list(range(9))               # every file exactly once
    assert {f["rank"] for f in files} == {0, 1}                         # both workers pulled from the queue
    bad = [f for f in files if not f["ok"]]
    assert [f["index"] for f in bad] == [4, 7]
    assert "No such file" in bad[0]["error"] or "FileNotFound" in bad[0]["error"]
    assert "Error loading file" in bad[1]["error"]
    for f in files:
        if f["ok"]:
            arc = (tmp_path / "out" / f"in{f['index']}.bce").read_bytes()
            assert arc[:4] == b"BCE" + bytes([f["index"]]) and len(arc) == f["bytes_out"]
            assert f["bytes_in"] == 1000 + 100 * f["index"] and f["gpu_ms"] == 2.5
    total = got[0][3]
    assert total["inputs"] == 7 and total["failed"] == 2
    assert total["bytes_in"] == sum(1000 + 100 * i for i in range(9) if i not in (4, 7))
    assert sorted(p.name for p in (tmp_path / "out").iterdir()) == sorted(f"in{i}.bce" for i in range(9) if i not in (4, 7))


def test_single_process_queue_without_a_process_group(tmp_path):
    from bce_b200 import batch
    for i in range(3):
        (tmp_path / f"f{i}").write_bytes(b"x" * (10 + i))
    res = batch.compress_files([str(tmp_path / f"f{i}") for i in range(3)], str(tmp_path / "o"),
                               lambda d, cfg: (bytes(d[:2]), 1.0))
    assert [f.ok for f in res.files] == [True] * 3 and res.total.inputs == 3 and res.total.failed == 0
    # two files in flight on one rank: two compressors, two worker threads, every file still exactly once
    import threading
    seen = []
    def make(tag):
        def fn(d, cfg):
            seen.append((tag, threading.get_ident()))
            return bytes([tag]) + bytes(d[:1]), 1.0
        return fn
    for i in range(3, 11):
        (tmp_path / f"f{i}").write_bytes(b"y" * (10 + i))
    res = batch.compress_files([str(tmp_path / f"f{i}") for i in range(11)], str(tmp_path / "o2"), [make(1), make(2)])
    assert [f.index for f in res.files] == list(range(11)) and all(f.ok for f in res.files)
    assert res.total.inputs == 11 and len(seen) == 11 and len({t for _, t in seen}) == 2


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
