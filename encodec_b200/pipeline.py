"""Host-to-host streaming around ``EncodecModel.forward``: upload, codec and download overlap across batches.

The reference is driven from host tensors (``inference.py:142,238-240``, ``compress.py:51``: ``model(x)`` on a batch that
was loaded on the CPU and whose results go back to the CPU). On a B200 one batch of BASELINE config 2 costs ~44 ms of
kernels plus ~2.5 ms of PCIe traffic each way; a caller that serialises copy-in, forward and copy-out on one stream pays
for all three. ``HostPipeline`` keeps ``depth`` batches in flight on three CUDA streams -- copy-in, compute (the caller's
current stream) and copy-out -- with events between them, so that the link time of batch ``i+1`` / ``i-1`` hides behind
the kernels of batch ``i``. Nothing here is numerical: results are exactly those of ``model(x)``.
"""
from __future__ import annotations

import typing as tp

import torch


class HostPipeline:
    """``for audio, codes in HostPipeline(model).run(batches)`` with ``batches`` an iterable of pinned host tensors
    ``[B, C, T]`` fp32. Yields pinned host tensors ``(audio [B, C, T] fp32, codes [B, n_q, T_f] int64)`` in order; a
    yielded pair is valid until ``depth`` further batches have been submitted (the buffers are a ring)."""

    def __init__(self, model, depth: int = 2):
        if depth < 1:
            raise ValueError("depth must be >= 1")
        p = next(model.parameters())
        if not p.is_cuda:
            raise RuntimeError("encodec_b200: HostPipeline needs the model on a CUDA device (no CPU fallback)")
        self.model = model
        self.device = p.device
        self.depth = depth
        self._in_stream = torch.cuda.Stream(self.device)
        self._out_stream = torch.cuda.Stream(self.device)
        self._slots: tp.List[dict] = [dict() for _ in range(depth)]

    def _slot_buffers(self, slot: dict, x_host: torch.Tensor):
        if slot.get("shape") != tuple(x_host.shape):
            slot.clear()
            slot["shape"] = tuple(x_host.shape)
            # allocated from the copy-in stream's pool: a block recycled from the compute stream could still be in use
            # by the previous batch's kernels when the upload (which does not wait for them) starts writing it
            with torch.cuda.stream(self._in_stream):
                slot["x_dev"] = torch.empty(x_host.shape, dtype=torch.float32, device=self.device)
            slot["x_dev"].record_stream(torch.cuda.current_stream(self.device))
            slot["in_free"] = None     # recorded on the compute stream once forward has consumed x_dev
            slot["out_done"] = None    # recorded on the copy-out stream once the host buffers hold the results
            slot["audio_host"] = None
            slot["codes_host"] = None
        return slot

    @torch.no_grad()
    def run(self, batches: tp.Iterable[torch.Tensor],
            after_forward: tp.Optional[tp.Callable[[torch.Tensor, torch.Tensor], None]] = None
            ) -> tp.Iterator[tp.Tuple[torch.Tensor, torch.Tensor]]:
        """``after_forward(audio, codes)``, if given, runs on the compute stream right after each forward (e.g. the
        multi-GPU gather of ``encodec_b200.dist``)."""
        compute = torch.cuda.current_stream(self.device)
        pending: tp.List[dict] = []
        for i, x_host in enumerate(batches):
            if x_host.is_cuda or x_host.dtype != torch.float32:
                raise ValueError("HostPipeline.run expects fp32 host tensors")
            if not x_host.is_pinned():
                x_host = x_host.pin_memory()
            if len(pending) == self.depth:           # the slot we are about to reuse: hand its results out first
                done = pending.pop(0)
                done["out_done"].synchronize()
                yield done["audio_host"], done["codes_host"]
            slot = self._slot_buffers(self._slots[i % self.depth], x_host)
            with torch.cuda.stream(self._in_stream):
                if slot["in_free"] is not None:
                    self._in_stream.wait_event(slot["in_free"])
                slot["x_dev"].copy_(x_host, non_blocking=True)
                in_ready = torch.cuda.Event()
                in_ready.record(self._in_stream)
            compute.wait_event(in_ready)
            audio, codes, _, _ = self.model(slot["x_dev"])
            if after_forward is not None:
                after_forward(audio, codes)
            slot["in_free"] = torch.cuda.Event()
            slot["in_free"].record(compute)
            done_ev = torch.cuda.Event()
            done_ev.record(compute)
            if slot["audio_host"] is None or slot["audio_host"].shape != audio.shape or slot["codes_host"].shape != codes.shape:
                slot["audio_host"] = torch.empty(audio.shape, dtype=audio.dtype).pin_memory()
                slot["codes_host"] = torch.empty(codes.shape, dtype=codes.dtype).pin_memory()
            with torch.cuda.stream(self._out_stream):
                self._out_stream.wait_event(done_ev)
                slot["audio_host"].copy_(audio, non_blocking=True)
                slot["codes_host"].copy_(codes, non_blocking=True)
                audio.record_stream(self._out_stream)
                codes.record_stream(self._out_stream)
                slot["out_done"] = torch.cuda.Event()
                slot["out_done"].record(self._out_stream)
            pending.append(slot)
        for done in pending:
            done["out_done"].synchronize()
            yield done["audio_host"], done["codes_host"]

    def join(self):
        """Make the caller's current stream wait for every copy issued so far (for device-side timing)."""
        compute = torch.cuda.current_stream(self.device)
        compute.wait_stream(self._in_stream)
        compute.wait_stream(self._out_stream)
