"""Fused global-norm clip + AdamW over flat parameter arenas.

Replaces the pair the reference's trainers run after every backward --
    torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1.0)        (Trainer.py:80, MultiTaskTrainer.py:205)
    optim.AdamW(model.parameters(), lr=1e-4, weight_decay=0.01).step()      (Trainer.py:19-21,81)
-- ~50 foreach launches over ~50 tensors, with three launches per parameter group: a multi-tensor gradient pack, the
sum of squares (mmsa_sumsq) and one clip+AdamW kernel (mmsa_clip_adamw) that reads p, g, m, v once and writes p, m, v
once (28 B per parameter: HBM-bound).  The parameters of a group are re-homed into ONE fp32 arena (`p.data` becomes a
view of it; Parameter identity, names and state_dict are unchanged), as are both moment buffers.

It is a torch.optim.Optimizer, so `ReduceLROnPlateau(optimizer, ...)` (Trainer.py:28) and `add_param_group`
(Trainer.py:24-26, the trainer's own contrastive weight) keep working; the global norm is taken over ALL groups, as
clip_grad_norm_ over the same parameters would.  When the gradients already live in one flat buffer in parameter order
(mmsa.dist.GradAllReducer after its all-reduce), that buffer is used in place and the pack disappears."""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch

from . import _lib

Tensor = torch.Tensor


def arena_layout(params, align: int = 64):
    """Offsets (in elements) of `params` inside one flat fp32 arena, each start aligned to `align` elements (256 B: the
    bf16 operand casts, TMA tensor maps and 16-byte vector loads all want aligned bases), and the arena length.  The
    gaps stay zero in p, g, m and v, which the update maps to zero again.  mmsa.dist.GradAllReducer packs gradients with
    the same layout, so its all-reduced buffer is consumed in place."""
    offs, off = [], 0
    for p in params:
        offs.append(off)
        off += (p.numel() + align - 1) // align * align
    return offs, off


class FusedClipAdamW(torch.optim.Optimizer):
    def __init__(self, params: Iterable, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, max_norm: Optional[float] = 1.0):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        self.max_norm = max_norm
        self._arenas = {}            # id(group) -> dict(p, m, v, g, views, key)
        self._step = 0
        super().__init__(params, defaults)

    # ------------------------------------------------------------------ arenas
    def _arena(self, group) -> dict:
        plist: List[torch.nn.Parameter] = [p for p in group["params"] if p.requires_grad]
        key = tuple(id(p) for p in plist)
        a = self._arenas.get(id(group))
        if a is not None and a["key"] == key:
            return a
        if not plist:
            a = {"key": key, "params": [], "n": 0}
            self._arenas[id(group)] = a
            return a
        dev = plist[0].device
        if not plist[0].is_cuda:
            raise _lib.MmsaError("mmsa.FusedClipAdamW: parameters must live on a CUDA device (no CPU fallback)")
        offs, n = arena_layout(plist)
        flat = torch.zeros(n, device=dev, dtype=torch.float32)
        views = []
        for p, off in zip(plist, offs):
            if p.dtype != torch.float32:
                raise TypeError("mmsa.FusedClipAdamW: fp32 master parameters expected")
            v = flat[off:off + p.numel()].view_as(p)
            v.copy_(p.data)
            p.data = v                                   # re-home the parameter into the arena
            views.append(v)
        old = a or {}
        m = torch.zeros(n, device=dev, dtype=torch.float32)
        v2 = torch.zeros(n, device=dev, dtype=torch.float32)
        if old.get("n") and old["key"] == key[:len(old["key"])]:      # same leading parameters: keep their moments
            m[:old["n"]].copy_(old["m"]); v2[:old["n"]].copy_(old["v"])
        a = {"key": key, "params": plist, "n": n, "offs": offs, "p": flat, "m": m, "v": v2,
             "g": torch.zeros(n, device=dev, dtype=torch.float32),
             "gviews": None, "sq": torch.zeros(1, device=dev, dtype=torch.float32),
             "partials": torch.empty(512, device=dev, dtype=torch.float32)}
        a["gviews"] = [a["g"][off:off + p.numel()].view_as(p) for p, off in zip(plist, offs)]
        self._arenas[id(group)] = a
        return a

    def _flat_grad(self, a: dict) -> Tensor:
        """the group's gradients as one flat fp32 tensor in parameter order (in place when they already are one)."""
        plist = a["params"]
        grads = [p.grad for p in plist]
        first = grads[0]
        if first is not None and first._base is not None and first._base.dtype == torch.float32 and first._base.ndim == 1:
            base, start, ok = first._base, first.storage_offset(), True
            for p, g, off in zip(plist, grads, a["offs"]):
                if g is None or g._base is not base or g.storage_offset() != start + off or not g.is_contiguous():
                    ok = False
                    break
            if ok and start + a["n"] <= base.numel():
                return base[start:start + a["n"]]
        live_dst = [d for d, g in zip(a["gviews"], grads) if g is not None]
        live_src = [g for g in grads if g is not None]
        if len(live_src) != len(grads):
            a["g"].zero_()                               # parameters without a gradient this step take a zero update
        if live_src:
            torch._foreach_copy_(live_dst, live_src)     # multi-tensor pack (plumbing)
        return a["g"]

    # ------------------------------------------------------------------ step
    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        st = torch.cuda.current_stream().cuda_stream
        self._step += 1
        work = []
        for group in self.param_groups:
            a = self._arena(group)
            if a["n"]:
                work.append((group, a, self._flat_grad(a)))
        if not work:
            return loss
        # global gradient norm over every group (clip_grad_norm_ semantics)
        total_sq = None
        for group, a, g in work:
            _lib.call("mmsa_sumsq", g.data_ptr(), a["n"], a["partials"].data_ptr(), 512, a["sq"].data_ptr(), st)
            total_sq = a["sq"] if total_sq is None else total_sq.add_(a["sq"])
        max_norm = float(self.max_norm) if self.max_norm is not None else 3.0e38
        for group, a, g in work:
            b1, b2 = group["betas"]
            _lib.call("mmsa_clip_adamw", a["p"].data_ptr(), g.data_ptr(), a["m"].data_ptr(), a["v"].data_ptr(), a["n"],
                      total_sq.data_ptr(), max_norm, float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                      float(group["weight_decay"]), self._step, st)
        # the kernels write the parameters behind autograd's back (no version bump): mark the cached bf16 operand
        # copies stale so that the next forward (or model.prepare_step()) re-casts them, into the same buffers
        from . import ops
        ops.bump_weights_epoch()
        return loss
