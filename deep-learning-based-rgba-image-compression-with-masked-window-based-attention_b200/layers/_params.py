"""Cache of the kernel-ready parameter blocks (`gdn_prepare` / `mwa_prepare` outputs) per module.

The block is a plain attribute, NOT a registered buffer: the state_dict keys must stay exactly the reference's.

When is it rebuilt?
* whenever a parameter it depends on changed version (`_version` bumps on every in-place update through the
  parameter itself: optimizer step, `load_state_dict`, init), moved (`data_ptr`, device) or the block size changed;
* on EVERY call while autograd is recording for one of the parameters (training: the modules call
  `_refresh_if_training()` from their forward) or while the current stream is being captured into a CUDA graph: the
  prepare kernel is a few microseconds, and this way a replayed graph re-derives the block from the weights it finds at
  replay time instead of baking in the ones of the capture;
* after `invalidate()`, which the modules call from `_apply` (`.to()`, `.cuda()`, `.half()` ...) and
  `_load_from_state_dict`.

What the key canNOT see: writes through `p.data` (`p.data.copy_()`, `p.data.mul_()`, `dist.broadcast(p.data)`, EMA
weight swaps) change neither `_version` nor `data_ptr`.  In inference mode call `module.invalidate_param_block()` (or
`mwa_b200.invalidate_param_blocks(model)`) after such a write; in training mode the block is rebuilt every call anyway.

Streams: the block remembers an event recorded behind its prepare kernel; a cache hit on another stream waits on it.
"""
from __future__ import annotations

import torch


class ParamBlock:
    def __init__(self):
        self.key = None
        self.blk: torch.Tensor | None = None
        self.ready: torch.cuda.Event | None = None
        self.stream = None

    def invalidate(self) -> None:
        self.key = None

    def __deepcopy__(self, memo):
        return ParamBlock()          # a copied module prepares its own block (CUDA events do not copy)

    def get(self, tensors, nbytes: int, fill):
        """tensors: iterable of parameters the block depends on; fill(blk) runs the prepare kernel on the current stream."""
        live = [t for t in tensors if t is not None]
        dev = live[0].device
        key = tuple((t.data_ptr(), t._version, str(t.device)) if t is not None else None for t in tensors)
        capturing = torch.cuda.is_current_stream_capturing()
        cur = torch.cuda.current_stream(dev)
        fresh = self.blk is None or self.blk.numel() != nbytes or self.blk.device != dev
        if fresh or capturing or key != self.key:
            if fresh:
                self.blk = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            elif self.ready is not None and self.stream != cur and not capturing:
                cur.wait_event(self.ready)              # the old contents may still be in use / being written elsewhere
            fill(self.blk)
            self.key = key
            self.stream = cur
            if not capturing:
                self.ready = torch.cuda.Event()
                self.ready.record(cur)
        elif self.stream != cur and self.ready is not None:
            cur.wait_event(self.ready)
        return self.blk


class ParamBlockOwner:
    """Mixin for the modules that own a ParamBlock in `self._blk`: drops the cached block whenever the parameters are
    replaced or re-loaded behind the cache key's back."""

    def invalidate_param_block(self) -> None:
        self._blk.invalidate()

    def _refresh_if_training(self) -> None:
        """call from the module's forward, OUTSIDE the autograd Function (inside it grad mode is off): when autograd is
        recording for one of this module's parameters the block is rebuilt on this call, whatever the cache key says"""
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            self._blk.invalidate()

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        if "_blk" in self.__dict__:
            self._blk.invalidate()
        return out

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        if "_blk" in self.__dict__:
            self._blk.invalidate()


def invalidate_param_blocks(model: torch.nn.Module) -> int:
    """Drop every cached parameter block below `model` (after writes through `.data`).  Returns how many were dropped."""
    n = 0
    for m in model.modules():
        if isinstance(m, ParamBlockOwner):
            m.invalidate_param_block()
            n += 1
        for name in ("_img", "_dimg"):                 # prepared weight images of the convolutions (layers/conv.py)
            img = m.__dict__.get(name)
            if img is not None and hasattr(img, "invalidate"):
                img.invalidate()
                n += 1
    return n
