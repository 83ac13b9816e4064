"""Data-parallel plumbing: the only collective of the path is the gradient average of one flat bf16
buffer per optimizer step (reference: Lightning DDP, train.py:461-474 -> 467.7 MB for tv2o-medium).

One process per GPU (torchrun), NCCL over NVLink 5 / NVSwitch.  The buffer is reduced in a few large
buckets on a side stream so the first buckets (token-level stack + lm_head, whose gradients are
complete first) overlap the event-level stack's backward.  `average_` also runs over gloo on CPU
tensors so the bucket logic is testable without GPUs.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def bucket_ranges(numel: int, seg_bounds: List[int], bucket_elems: int = 64 * 1024 * 1024) -> List[Tuple[int, int]]:
    """Split [0, numel) at the given segment boundaries, then into buckets of <= bucket_elems."""
    cuts = sorted(set([0, numel] + [b for b in seg_bounds if 0 < b < numel]))
    out = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        s = a
        while s < b:
            e = min(b, s + bucket_elems)
            out.append((s, e))
            s = e
    return out


def average_(flat: torch.Tensor, start: int, end: int, group=None):
    """In-place average of flat[start:end] over the process group (NCCL: native AVG; gloo: SUM then scale)."""
    world = dist.get_world_size(group)
    if world == 1:
        return
    view = flat[start:end]
    if dist.get_backend(group) == "nccl":
        dist.all_reduce(view, op=dist.ReduceOp.AVG, group=group)
    else:
        dist.all_reduce(view, op=dist.ReduceOp.SUM, group=group)
        view.div_(world)


class GradSync:
    """Overlapped gradient averaging for the fused trainer: call `ready(start, end)` as soon as a slice of
    the flat gradient buffer is final, `wait()` before the optimizer."""

    def __init__(self, flat: torch.Tensor, group=None, bucket_elems: int = 64 * 1024 * 1024):
        self.flat, self.group, self.bucket = flat, group, bucket_elems
        self.stream = torch.cuda.Stream(device=flat.device) if flat.is_cuda else None

    def ready(self, start: int, end: int):
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return
        if self.stream is None:
            for a, b in bucket_ranges(end - start, [], self.bucket):
                average_(self.flat, start + a, start + b, self.group)
            return
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            for a, b in bucket_ranges(end - start, [], self.bucket):
                average_(self.flat, start + a, start + b, self.group)

    def wait(self):
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)
