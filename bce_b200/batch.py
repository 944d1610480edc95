"""Batch mode: independent inputs sharded over the GPUs of one node, one archive per input.

The path does not shard inside an input (suffix order is global), so multi-GPU
means replicas: rank r of W takes inputs r, r+W, r+2W, ... and nothing but a small stats
vector ever crosses NVLink (one all_gather at the end; NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Sequence

import torch
import torch.distributed as dist


@dataclass
class RankStats:
    inputs: int = 0
    bytes_in: int = 0
    bytes_out: int = 0
    counts: int = 0
    gpu_ms: float = 0.0
    wall_ms: float = 0.0

    def as_tensor(self, device) -> torch.Tensor:
        return torch.tensor([self.inputs, self.bytes_in, self.bytes_out, self.counts, self.gpu_ms, self.wall_ms],
                            dtype=torch.float64, device=device)

    @staticmethod
    def from_tensor(t: torch.Tensor) -> "RankStats":
        v = t.tolist()
        return RankStats(int(v[0]), int(v[1]), int(v[2]), int(v[3]), float(v[4]), float(v[5]))


@dataclass
class BatchResult:
    per_rank: list = field(default_factory=list)
    outputs: dict = field(default_factory=dict)      # index -> whatever `work` returned (this rank only)

    @property
    def total(self) -> RankStats:
        t = RankStats()
        for r in self.per_rank:
            t.inputs += r.inputs
            t.bytes_in += r.bytes_in
            t.bytes_out += r.bytes_out
            t.counts += r.counts
            t.gpu_ms = max(t.gpu_ms, r.gpu_ms)        # ranks run concurrently: the job takes the max
            t.wall_ms = max(t.wall_ms, r.wall_ms)
        return t


def shard(num_inputs: int, rank: int, world: int) -> list:
    """Indices of the inputs rank `rank` processes (round robin: equal counts +-1)."""
    return list(range(rank, num_inputs, world))


def gather_stats(local: RankStats, device="cpu") -> list:
    """All ranks learn every rank's stats (the only collective on this path)."""
    if not (dist.is_available() and dist.is_initialized()):
        return [local]
    world = dist.get_world_size()
    mine = local.as_tensor(device)
    out = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(out, mine)
    return [RankStats.from_tensor(t.cpu()) for t in out]


def run_batch(num_inputs: int, work: Callable[[int], tuple], device="cpu") -> BatchResult:
    """`work(i)` processes input i on this rank's GPU and returns
    (bytes_in, bytes_out, counts, gpu_ms, wall_ms, output)."""
    rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    local = RankStats()
    res = BatchResult()
    for i in shard(num_inputs, rank, world):
        b_in, b_out, counts, gpu_ms, wall_ms, out = work(i)
        local.inputs += 1
        local.bytes_in += b_in
        local.bytes_out += b_out
        local.counts += counts
        local.gpu_ms += gpu_ms
        local.wall_ms += wall_ms
        res.outputs[i] = out
    res.per_rank = gather_stats(local, device)
    return res
