"""Data-parallel training step of the hot path (BASELINE.json config 5): the one collective on this path.

The reference trains in a single process (`trainRGB.py:158-255`; `torch.nn.DataParallel` is commented out, :374).
Sharding the batch by image over the ranks leaves exactly one exchange step per iteration: the gradients.
`GradientAllReduce` reproduces the single-process update on every rank:

    loss.backward()            per rank, on its shard (mean over the local images)
    GradientAllReduce(...)()   bucketed all-reduce (SUM) over NCCL / NVLink, divided by the world size
                               == gradient of the mean over the GLOBAL batch (equal shards);
                               then the reference's value clip `grad.clamp_(-5, 5)` (`trainRGB.py:190-195`),
                               applied AFTER the reduction -- clipping before it would change the update
    optimizer.step()           identical state on every rank (replicated parameters and Adam moments)

Buckets are flat fp32 buffers (default 32 MiB: NVSwitch all-reduce cost is latency-, not link-bound, SURVEY.md
section 8e), filled in reverse parameter order (the order backward produces gradients) and reduced with
asynchronous collectives that overlap each other; the copy-back fuses the 1/world scale and the clip.
Works with any `torch.distributed` backend (NCCL on the B200 box, gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class GradientAllReduce:
    def __init__(self, params, bucket_bytes: int = 32 << 20, clip_value: float | None = 5.0, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.clip_value = clip_value
        self.group = group
        self.buckets: list[list[torch.nn.Parameter]] = []
        cur, cur_bytes = [], 0
        for p in reversed(self.params):
            nbytes = p.numel() * p.element_size()
            if cur and (cur_bytes + nbytes > bucket_bytes or p.dtype != cur[0].dtype or p.device != cur[0].device):
                self.buckets.append(cur)
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += nbytes
        if cur:
            self.buckets.append(cur)
        self._flat: list[torch.Tensor | None] = [None] * len(self.buckets)
        self._view_cache: list[list[torch.Tensor]] = [[] for _ in self.buckets]

    @property
    def bytes_per_step(self) -> int:
        return sum(p.numel() * p.element_size() for p in self.params)

    def _views(self, i: int):
        """flat buffer of bucket i and one view per parameter (shaped like it)"""
        bucket = self.buckets[i]
        n = sum(p.numel() for p in bucket)
        if self._flat[i] is None or self._flat[i].numel() != n:
            flat = torch.empty(n, dtype=bucket[0].dtype, device=bucket[0].device)
            views, off = [], 0
            for p in bucket:
                views.append(flat[off:off + p.numel()].view_as(p))
                off += p.numel()
            self._flat[i] = flat
            self._view_cache[i] = views
        return self._flat[i], self._view_cache[i]

    def __call__(self) -> None:
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        works = []
        for i, bucket in enumerate(self.buckets):
            flat, views = self._views(i)
            have = [(v, p.grad) for v, p in zip(views, bucket) if p.grad is not None]
            for v, p in zip(views, bucket):
                if p.grad is None:
                    v.zero_()            # a rank that did not touch p still takes part in the reduction
            if have:                     # one multi-tensor copy instead of a launch per parameter
                torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
            works.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
                         if world > 1 else None)
        for i, bucket in enumerate(self.buckets):
            if works[i] is not None:
                works[i].wait()
            flat, views = self._views(i)
            if world > 1:
                flat.mul_(1.0 / world)
            if self.clip_value is not None:
                flat.clamp_(-self.clip_value, self.clip_value)
            for v, p in zip(views, bucket):
                if p.grad is None:
                    p.grad = v.clone()
            torch._foreach_copy_([p.grad for p in bucket], views)
