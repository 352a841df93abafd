"""Host -> device staging for streaming inference: the copy of batch i+1 overlaps the compute of
batch i (one side stream, two static device slots), results are read back on the compute stream.

The reference feeds the model from a DataLoader and copies synchronously (`X[feature].to(device)`,
trainer.py:350-351, :470-472); on a B200 the 54.7 MB of fp32 crops per 8-window batch cost ~1 ms
of PCIe time per step if they are not overlapped.  PyTorch is used for streams, events and the
copies only.
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable, Optional

import torch


class HostPrefetcher:
    def __init__(self, device: torch.device, depth: int = 2):
        self.device = torch.device(device)
        self.depth = depth
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._slots = [None] * depth                  # dict of static device buffers per slot
        self._ready = [torch.cuda.Event() for _ in range(depth)]
        self._free = [None] * depth                   # recorded on the compute stream after a slot was consumed

    def _stage(self, slot: int, host: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        flat = self._slots[slot]
        if flat is None:
            flat = self._slots[slot] = {}
        bufs = {}
        for k, v in host.items():                      # grow-only flat device buffers, viewed at the batch's shape
            f = flat.get(k)
            if f is None or f.dtype != v.dtype or f.numel() < v.numel():
                f = flat[k] = torch.empty(max(v.numel(), 1), dtype=v.dtype, device=self.device)
            bufs[k] = f[:v.numel()].view(v.shape)
        with torch.cuda.stream(self.copy_stream):
            if self._free[slot] is not None:
                self.copy_stream.wait_event(self._free[slot])      # the previous user of this slot is done
            for k, v in host.items():
                bufs[k].copy_(v, non_blocking=True)
            self._ready[slot].record(self.copy_stream)
        return bufs

    def run(self, host_batches: Iterable[Dict[str, torch.Tensor]], fn: Callable[[Dict[str, torch.Tensor]], torch.Tensor],
            sink: Optional[Callable[[int, torch.Tensor], None]] = None) -> int:
        """For every pinned host batch: stage it (ahead of time), run ``fn`` on the device copy on the
        current stream, hand the result to ``sink(i, out)``.  Returns the number of batches."""
        main = torch.cuda.current_stream(self.device)
        it = iter(host_batches)
        pending = []
        n = 0
        try:
            pending.append(self._stage(0, next(it)))
        except StopIteration:
            return 0
        while pending:
            slot = n % self.depth
            nxt = None
            try:
                nxt = next(it)
            except StopIteration:
                pass
            if nxt is not None:
                pending.append(self._stage((n + 1) % self.depth, nxt))   # overlaps fn(batch n)
            bufs = pending.pop(0)
            main.wait_event(self._ready[slot])
            out = fn(dict(bufs))
            ev = torch.cuda.Event()
            ev.record(main)
            self._free[slot] = ev
            if sink is not None:
                sink(n, out)
            n += 1
        return n
