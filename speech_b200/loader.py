"""Input pipeline for the GPU step (batch assembly / H2D).

The reference assembles every minibatch on the training thread inside `model.loss(batch)`
(`collate` -> `zero_pad_concat` -> `.cuda()`, speech/models/ctc_model.py:26-27,42-53,
model.py:135-141), so padding and the host-to-device copy sit on the critical path of every step.
`BatchPrefetcher` moves both off it: a worker thread collates batch i+1 into pinned memory and
copies it to the device on its own CUDA stream while the GPU is still busy with step i; the
training loop receives `StagedBatch` objects, which `model.loss` / `model.infer` accept in place of
the reference's `(inputs, labels)` pair.

    loader = speech.loader.make_loader(...)            # the reference's loader, unchanged
    for batch in BatchPrefetcher(model, loader):
        loss = model.loss(batch)
"""
import queue
import threading

import torch


class StagedBatch:
    """A collated minibatch whose inputs already are (or are on their way to be) on the device."""
    __slots__ = ("x", "y", "x_lens", "y_lens", "ready")

    def __init__(self, x, y, x_lens, y_lens, ready):
        self.x, self.y, self.x_lens, self.y_lens, self.ready = x, y, x_lens, y_lens, ready

    def tensors(self):
        """(x, y, x_lens, y_lens) for use on the CURRENT stream."""
        if self.ready is not None:
            cur = torch.cuda.current_stream()
            cur.wait_event(self.ready)
            self.x.record_stream(cur)
            self.ready = None
        return [self.x, self.y, self.x_lens, self.y_lens]


class BatchPrefetcher:
    def __init__(self, model, batches, depth=1):
        """model: a speech_b200 model on a CUDA device (its `collate` defines the batch layout);
        batches: iterable of reference-style batches; depth: staged batches kept ahead."""
        self.model = model
        self.device = next(model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("BatchPrefetcher needs the model on a CUDA device")
        self.batches = batches
        self.q = queue.Queue(maxsize=max(1, depth))
        self.error = None
        self.stop = False
        self.thread = threading.Thread(target=self._work, daemon=True)
        self.thread.start()

    def _work(self):
        try:
            torch.cuda.set_device(self.device)
            stream = torch.cuda.Stream(device=self.device)
            for batch in self.batches:
                if self.stop:
                    break
                with torch.cuda.stream(stream):
                    x, y, x_lens, y_lens = self.model.collate(*batch)
                    ev = torch.cuda.Event()
                    ev.record(stream)
                self._put(StagedBatch(x, y, x_lens, y_lens, ev))
        except BaseException as e:      # surfaced in the consumer thread
            self.error = e
        self._put(None)

    def _put(self, item):
        while not self.stop:
            try:
                self.q.put(item, timeout=0.1)
                return
            except queue.Full:
                continue

    def __iter__(self):
        while True:
            item = self.q.get()
            if item is None:
                if self.error is not None:
                    raise self.error
                return
            yield item

    def close(self):
        self.stop = True
