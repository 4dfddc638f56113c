"""Cache of the kernel-ready parameter blocks (`gdn_prepare` / `mwa_prepare` outputs) per module.

The block is rebuilt whenever a parameter changed (`_version` bumps on every in-place update:
optimizer step, load_state_dict, init) or moved (`data_ptr`, device).  It is a plain attribute,
NOT a registered buffer: the state_dict keys must stay exactly the reference's.
"""
from __future__ import annotations

import torch


class ParamBlock:
    def __init__(self):
        self.key = None
        self.blk: torch.Tensor | None = None

    def get(self, tensors, nbytes: int, fill):
        """tensors: iterable of parameters the block depends on; fill(blk) runs the prepare kernel."""
        key = tuple((t.data_ptr(), t._version, str(t.device)) if t is not None else None for t in tensors)
        if self.blk is None or key != self.key or self.blk.numel() != nbytes:
            dev = next(t.device for t in tensors if t is not None)
            blk = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            fill(blk)
            self.blk, self.key = blk, key
        return self.blk
