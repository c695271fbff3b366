"""Host -> device input staging for the training loop.

The reference loop copies each batch with `.to(device, non_blocking=True)` right before the forward
(run_pretrain_distributed_gpt3.py:103-107); with pinned host memory that copy can run on the DMA engine while the
previous step is still computing.  `DevicePrefetcher` does exactly that: `submit()` enqueues the copies of the NEXT
batch on a side stream, `take()` hands the oldest staged batch to the current stream.  `ops.clip_normalize` can be
applied to a staged uint8 clip for the device-side input tail."""
import collections

import torch


class DevicePrefetcher:
    def __init__(self, device, depth=2):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.depth = depth
        self._q = collections.deque()

    def submit(self, *host_tensors):
        """Start copying one batch (pinned CPU tensors) to the device; returns immediately."""
        if len(self._q) >= self.depth:
            raise RuntimeError("DevicePrefetcher: take() a batch before submitting more")
        with torch.cuda.stream(self.stream):
            dev = [t.to(self.device, non_blocking=True) for t in host_tensors]
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._q.append((dev, ev))

    def take(self):
        """The oldest staged batch, ordered after its copy on the current stream."""
        dev, ev = self._q.popleft()
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        for t in dev:
            t.record_stream(cur)  # allocated on the copy stream, consumed on the compute stream
        return dev

    def __len__(self):
        return len(self._q)
