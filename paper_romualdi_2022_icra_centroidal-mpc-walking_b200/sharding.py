"""Multi-GPU plumbing of the batched solve: instances are independent, so ranks own contiguous index ranges and the only
exchange is one all_gather of the per-instance results (objective, status, iterations) at the end of a batch
(SURVEY.md 8(e)).  torch.distributed is the transport (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(total: int, rank: int, world: int) -> tuple[int, int]:
    """contiguous range [lo, hi) of instance indices owned by `rank`; sizes differ by at most one"""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_results(obj: torch.Tensor, status: torch.Tensor, iters: torch.Tensor) -> torch.Tensor:
    """(B, 3) float64: objective, status, iterations of every instance of the local shard"""
    return torch.stack([obj.double(), status.double(), iters.double()], dim=1)


def gather_results(local: torch.Tensor, world: int, sizes: list[int] | None = None) -> torch.Tensor:
    """all_gather of the packed local results; `sizes` = rows per rank when the shards are uneven (padded exchange)"""
    if world == 1:
        return local
    n = max(sizes) if sizes else local.shape[0]
    buf = local
    if local.shape[0] < n:
        buf = torch.zeros(n, local.shape[1], dtype=local.dtype, device=local.device)
        buf[: local.shape[0]] = local
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    if sizes:
        out = [o[:s] for o, s in zip(out, sizes)]
    return torch.cat(out, dim=0)


class AsyncGather:
    """The gather of a batch's results overlapped with the solve of the next batch: the all_gather is enqueued on a side
    stream behind an event of the producing stream, so a rank that finishes a batch early starts the next one instead of
    waiting for the slowest rank at every batch (the ranks only meet in wait()).  CUDA / NCCL only; with world == 1 or on
    CPU tensors it degenerates to gather_results."""

    def __init__(self, world: int, sizes: list[int] | None = None):
        self.world, self.sizes, self.pending = world, sizes, []
        self.stream = torch.cuda.Stream() if (world > 1 and torch.cuda.is_available()) else None

    def submit(self, local: torch.Tensor):
        if self.stream is None or not local.is_cuda:
            self.pending.append(gather_results(local, self.world, self.sizes))
            return
        ready = torch.cuda.Event()
        ready.record()                                   # on the stream that produced `local`
        local.record_stream(self.stream)
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            self.pending.append(gather_results(local, self.world, self.sizes))

    def wait(self) -> list[torch.Tensor]:
        """makes the current stream wait for every submitted gather; returns the gathered results in submission order"""
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)
        out, self.pending = self.pending, []
        return out
