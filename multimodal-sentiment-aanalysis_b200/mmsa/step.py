"""One training step of the fusion path -- forward, CE + contrastive loss, backward -- as a
callable over STATIC device buffers, optionally captured in a CUDA graph.

This is the loop body of Trainer.py:59-79 (`outputs, contrastive_loss = model(eeg, eye, pps, labels)`;
`loss = CE(outputs, labels) + w * contrastive_loss`; `loss.backward()`), minus the optimiser.  The
step launches ~150 kernels of 5-300 us each; issued one by one from Python they are launch-bound,
so the steady-state loop replays a captured graph instead (no tracing compiler involved: the graph
is the recorded sequence of the library's own kernel launches).

Each *slot* owns one set of input buffers (text, image, labels) and one captured graph; two slots
let the host->device copy of step i+1 overlap the kernels of step i."""
from __future__ import annotations

from typing import Callable, List, Optional

import torch

from . import ops

Tensor = torch.Tensor


class _Slot:
    __slots__ = ("text", "image", "labels", "loss", "logits", "graph", "grads")


class TrainStep:
    def __init__(self, model, batch: int, text_len: int, regions: int = 49, text_dim: int = 768,
                 image_dim: int = 2048, feature_dtype: torch.dtype = torch.bfloat16, n_slots: int = 1,
                 use_graph: bool = True, post_backward: Optional[Callable[[], None]] = None,
                 device: Optional[torch.device] = None):
        self.model = model
        self.device = device or next(model.parameters()).device
        self.use_graph = use_graph
        self.post_backward = post_backward
        self.slots: List[_Slot] = []
        for _ in range(n_slots):
            s = _Slot()
            s.text = torch.zeros((batch, text_len, text_dim), device=self.device, dtype=feature_dtype)
            s.image = torch.zeros((batch, regions, image_dim), device=self.device, dtype=feature_dtype)
            s.labels = torch.zeros((batch,), device=self.device, dtype=torch.int64)
            s.loss = None
            s.logits = None
            s.graph = None
            s.grads = None
            self.slots.append(s)
        self._pool = None
        self._one = torch.ones((), device=self.device, dtype=torch.float32)
        self.launches_per_step = 0

    # the arithmetic of one step; every op below is a kernel of libmmsa.so
    def _body(self, s: _Slot) -> None:
        self.model.prepare_step()                                            # bf16 operand copies of the weights
        logits, closs = self.model(s.text, s.image, None, s.labels)          # Trainer.py:60
        loss = ops.cross_entropy(logits, s.labels, extra=closs)              # Trainer.py:68-71 (CE + w * contrastive)
        loss.backward(gradient=self._one)                                    # Trainer.py:79 (seed given: no fill launch)
        if self.post_backward is not None:
            self.post_backward()                                             # e.g. gradient all-reduce
        s.loss, s.logits = loss.detach(), logits.detach()

    def warmup(self, iters: int = 3) -> None:
        """Eager iterations on a side stream (allocator warm-up, lazy module state) before capture."""
        st = torch.cuda.Stream(device=self.device)
        st.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(st):
            for _ in range(iters):
                for s in self.slots:
                    self.model.zero_grad(set_to_none=True)
                    self._body(s)
        torch.cuda.current_stream(self.device).wait_stream(st)
        torch.cuda.synchronize(self.device)

    def _capture_stream(self):
        """The main chain is captured on a high-priority stream: its kernel nodes keep that priority in the graph, so
        when a main-chain kernel and a weight-gradient GEMM of the helper stream (default priority, ops._side_stream) are
        both pending, the SMs go to the critical path first and the weight gradients fill what is left."""
        if getattr(self, "_cap_stream", None) is None:
            self._cap_stream = torch.cuda.Stream(device=self.device, priority=ops.critical_priority())
        return self._cap_stream

    def capture(self) -> None:
        from . import _lib
        if not self.use_graph:
            return
        params = [p for p in self.model.parameters()]
        for s in self.slots:
            # fresh .grad tensors per slot, allocated from the graph pool during capture, so that a
            # replay WRITES this slot's gradients instead of accumulating onto another slot's
            self.model.zero_grad(set_to_none=True)
            g = torch.cuda.CUDAGraph()
            n0 = _lib.launch_count()
            with torch.cuda.graph(g, pool=self._pool, stream=self._capture_stream()):
                self._body(s)
            self.launches_per_step = _lib.launch_count() - n0
            if self._pool is None:
                self._pool = g.pool()
            s.graph = g
            s.grads = [p.grad for p in params]
        torch.cuda.synchronize(self.device)

    def run(self, slot: int = 0) -> Tensor:
        """Run one step on the inputs currently in slot `slot`; returns the (device) loss."""
        s = self.slots[slot]
        if s.graph is not None:
            s.graph.replay()
            if len(self.slots) > 1:          # point .grad at the tensors this slot's graph writes
                for p, g in zip(self.model.parameters(), s.grads):
                    p.grad = g
        else:
            from . import _lib
            n0 = _lib.launch_count()
            self.model.zero_grad(set_to_none=True)
            self._body(s)
            self.launches_per_step = _lib.launch_count() - n0
        return s.loss
