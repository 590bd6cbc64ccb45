"""Double-buffered host -> device input pipeline.

The reference feeds its training loop from a ``DataLoader(pin_memory=True)`` and moves each batch with
``x.to(device, non_blocking=True)`` (reference train.py:150-165, 207); the copy of the next batch can then overlap the
step on the current one.  ``HostPrefetcher`` is that overlap made explicit: batch k+1 is copied on a dedicated copy
stream while step k runs on the compute stream; ``next()`` hands out device tensors the compute stream may use."""
from collections import deque

import torch


class HostPrefetcher:
    def __init__(self, batches, device, depth=2):
        """batches: iterable of tuples / lists of (pinned) CPU tensors; depth: batches in flight (>= 1)."""
        self.it = iter(batches)
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.queue = deque()
        self.depth = max(1, int(depth))
        for _ in range(self.depth - 1):
            self._issue()

    def _issue(self):
        try:
            batch = next(self.it)
        except StopIteration:
            return
        self.stream.wait_stream(torch.cuda.current_stream(self.device))   # (re-used device blocks: ordered after their last use)
        with torch.cuda.stream(self.stream):
            dev = tuple(t.to(self.device, non_blocking=True) for t in batch)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.queue.append((dev, ev))

    def __iter__(self):
        return self

    def __next__(self):
        if not self.queue:
            self._issue()
        if not self.queue:
            raise StopIteration
        dev, ev = self.queue.popleft()
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        for t in dev:
            t.record_stream(cur)        # allocated on the copy stream, consumed on the compute stream
        self._issue()                   # the next batch travels while this one is being consumed
        return dev
