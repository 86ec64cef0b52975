"""Host → device hand-off of batches, one batch ahead of the training step.

The reference moves every batch with a synchronous `data.to(device)` inside the step
(/root/reference/training_scripts/train_class_simple.py:330-340; `DataLoader(num_workers <= 1, pin_memory=False)`,
/root/reference/src/UCF_VIT/dataloaders/datamodule.py:245-248,515-522), so the 154 MB of a ViT-B batch cross PCIe
while the GPU idles.  `DevicePrefetcher` wraps any iterable of host batches:

    for data, label in DevicePrefetcher(train_dataloader, device):
        loss = step(data, label)

Batches are staged through page-locked buffers (skipped when the loader already pins) and copied on a side stream
into one of `depth` device slots while the previous step runs; the consumer's stream waits on the slot's event, and
a slot is only overwritten after the consumer has moved past it.  A yielded batch stays valid until `depth - 1`
further batches have been requested.  CUDA only.
"""
from typing import Any, Iterable, List

import torch


class _Slot:
    def __init__(self):
        self.pinned: List[torch.Tensor] = []
        self.device: List[torch.Tensor] = []
        self.ready = torch.cuda.Event()        # H2D copies of this slot have been enqueued up to here
        self.consumed = torch.cuda.Event()     # the consumer's stream is done with the slot's tensors
        self.used = False


def _flatten(batch):
    """(tensors, rebuild) for a tensor, or a (nested) tuple / list / dict of tensors and pass-through objects."""
    leaves: List[Any] = []

    def walk(x):
        if isinstance(x, torch.Tensor):
            leaves.append(x)
            return ("t", len(leaves) - 1)
        if isinstance(x, (tuple, list)):
            return ("s", type(x), [walk(v) for v in x])
        if isinstance(x, dict):
            return ("d", [(k, walk(v)) for k, v in x.items()])
        return ("o", x)

    spec = walk(batch)

    def rebuild(ts, node=spec):
        kind = node[0]
        if kind == "t":
            return ts[node[1]]
        if kind == "s":
            vals = [rebuild(ts, n) for n in node[2]]
            return tuple(vals) if node[1] is tuple else node[1](vals)
        if kind == "d":
            return {k: rebuild(ts, n) for k, n in node[1]}
        return node[1]

    return leaves, rebuild


class DevicePrefetcher:
    def __init__(self, loader: Iterable, device, depth: int = 2):
        if depth < 2:
            raise ValueError("DevicePrefetcher needs depth >= 2 (one slot in use, one in flight)")
        self.loader = loader
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DevicePrefetcher moves batches to a CUDA device; there is no CPU path")
        self.depth = depth
        self.h2d_bytes = 0                      # bytes copied so far (bench.py's e2e accounting)

    def __len__(self):
        return len(self.loader)

    def _issue(self, slot: _Slot, batch, copy_stream):
        leaves, rebuild = _flatten(batch)
        if slot.used:
            slot.ready.synchronize()            # the staging buffers may be rewritten only after their copies ran
        if len(slot.device) != len(leaves):
            slot.pinned, slot.device = [None] * len(leaves), [None] * len(leaves)
        for i, t in enumerate(leaves):          # (re)allocate on the consumer's stream, before any copy is queued
            d = slot.device[i]
            if d is None or d.shape != t.shape or d.dtype != t.dtype:
                slot.device[i] = torch.empty(t.shape, dtype=t.dtype, device=self.device)
                slot.pinned[i] = None
            if not t.is_cuda and not t.is_pinned():
                if slot.pinned[i] is None:
                    slot.pinned[i] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                slot.pinned[i].copy_(t)
        with torch.cuda.stream(copy_stream):
            if slot.used:
                copy_stream.wait_event(slot.consumed)
            else:
                copy_stream.wait_stream(torch.cuda.current_stream(self.device))
            for i, t in enumerate(leaves):
                src = t if (t.is_cuda or t.is_pinned()) else slot.pinned[i]
                slot.device[i].copy_(src, non_blocking=True)
                if not t.is_cuda:
                    self.h2d_bytes += t.numel() * t.element_size()
            slot.ready.record(copy_stream)
        slot.used = True
        return rebuild(slot.device)

    def __iter__(self):
        copy_stream = torch.cuda.Stream(device=self.device)
        slots = [_Slot() for _ in range(self.depth)]
        it = iter(self.loader)
        pending = []                            # (slot, batch on the device), oldest first
        nxt = 0
        try:
            for _ in range(self.depth - 1):
                pending.append((slots[nxt], self._issue(slots[nxt], next(it), copy_stream)))
                nxt = (nxt + 1) % self.depth
        except StopIteration:
            it = None
        prev = None
        while pending:
            if prev is not None:
                prev.consumed.record(torch.cuda.current_stream(self.device))
            if it is not None:
                try:
                    pending.append((slots[nxt], self._issue(slots[nxt], next(it), copy_stream)))
                    nxt = (nxt + 1) % self.depth
                except StopIteration:
                    it = None
            slot, dev_batch = pending.pop(0)
            torch.cuda.current_stream(self.device).wait_event(slot.ready)
            prev = slot
            yield dev_batch
