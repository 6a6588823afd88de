"""Host -> device input prefetch on a copy stream (the device side of what the reference's DataLoader does with
``pin_memory`` + ``non_blocking=True`` copies, /root/reference/src/utils/vqvae.py:547-552): the upload of the NEXT
batch overlaps the training step of the current one instead of sitting in front of it on the compute stream.

    pre = DevicePrefetcher(device)
    h = pre.upload(first_pinned_batch)
    for batch in ...:
        x = pre.take(h)                    # the compute stream waits for that copy only
        h = pre.upload(next_pinned_batch)  # runs under the step below
        step(x)

PyTorch plumbing only (streams, events, device buffers); no arithmetic.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch


class DevicePrefetcher:
    def __init__(self, device, depth: int = 2):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self.depth = depth
        self.slots: List[Optional[torch.Tensor]] = [None] * depth
        self.next = 0

    def upload(self, host: torch.Tensor) -> Tuple[torch.Tensor, torch.cuda.Event]:
        """enqueue the copy of a (pinned) host tensor on the copy stream; returns a handle for `take`"""
        k = self.next
        self.next = (k + 1) % self.depth
        buf = self.slots[k]
        if buf is None or buf.shape != host.shape or buf.dtype != host.dtype:
            buf = self.slots[k] = torch.empty(host.shape, dtype=host.dtype, device=self.device)
        # the slot's previous consumer was enqueued on the compute stream before this call: do not overwrite under it
        guard = torch.cuda.Event()
        guard.record(torch.cuda.current_stream(self.device))
        self.stream.wait_event(guard)
        with torch.cuda.stream(self.stream):
            buf.copy_(host, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.stream)
        return buf, done

    def take(self, handle: Tuple[torch.Tensor, torch.cuda.Event]) -> torch.Tensor:
        """the uploaded tensor, valid for work enqueued on the current stream from here on"""
        buf, done = handle
        torch.cuda.current_stream(self.device).wait_event(done)
        return buf
