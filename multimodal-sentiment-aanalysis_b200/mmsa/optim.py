"""Fused global-norm clip + AdamW over flat parameter arenas.

Replaces the pair the reference's trainers run after every backward --
    torch.nn.utils.clip_grad_norm_(self.model.parameters(), 1.0)            (Trainer.py:80, MultiTaskTrainer.py:205)
    optim.AdamW(model.parameters(), lr=1e-4, weight_decay=0.01).step()      (Trainer.py:19-21,81)
-- ~50 foreach launches over ~50 tensors, with three launches per parameter group: a multi-tensor gradient pack, the
sum of squares (mmsa_sumsq) and one clip+AdamW kernel (mmsa_clip_adamw) that reads p, g, m, v once and writes p, m, v
once (28 B per parameter: HBM-bound).  The parameters of a group are re-homed into ONE fp32 arena (`p.data` becomes a
view of it; Parameter identity, names and state_dict are unchanged), as are both moment buffers.

Semantics follow the reference's call pattern exactly:
* **What is clipped.**  `clip_grad_norm_(self.model.parameters(), 1.0)` covers the MODEL's parameters only -- not the
  trainer's own `contrastive_weight`, which joins the optimiser later through `add_param_group` (Trainer.py:24-26) and
  whose gradient (= the contrastive loss value, order 1-10 at T = 0.01) would otherwise dominate the norm.  Every group
  carries a `clip` flag: groups given to the constructor default to True, groups added later default to False (pass
  `{"params": ..., "clip": True}` to include one).  The norm is taken over the clip groups; only their gradients are
  scaled.
* **Parameters without a gradient are skipped** (no weight decay, no moment decay, no step count), as torch.optim.AdamW
  does -- e.g. the whole `valence_head` under contract="single".  Step counts are per parameter, as in torch.
* **State** lives where torch keeps it: `optimizer.state[p] = {"step", "exp_avg", "exp_avg_sq"}` (the moments are views
  of the arenas), so `state_dict()` / `load_state_dict()` round-trip and are interchangeable with torch.optim.AdamW's
  for the same parameter order.  Arenas are built eagerly (constructor / `add_param_group`) and keyed by group index.

**CUDA graphs**: construct the optimiser BEFORE capturing a step graph -- the constructor moves `p.data` into the arena,
and a graph captured earlier would keep reading the old storage.  `step()` raises if a parameter's storage was moved
away from its arena afterwards (e.g. `model.to(...)`).

It is a torch.optim.Optimizer, so `ReduceLROnPlateau(optimizer, ...)` (Trainer.py:28) and `add_param_group` keep
working.  When the gradients already live in one flat buffer in parameter order (mmsa.dist.GradAllReducer after its
all-reduce), that buffer is used in place and the pack disappears."""
from __future__ import annotations

from typing import Iterable, List, Optional, Tuple

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
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, clip=True)
        self.max_norm = max_norm
        self._arenas: List[dict] = []        # one per param group, same index
        self._constructing = True
        super().__init__(params, defaults)
        self._constructing = False

    # ------------------------------------------------------------------ arenas
    def add_param_group(self, param_group) -> None:
        if not self._constructing and isinstance(param_group, dict):
            param_group.setdefault("clip", False)      # e.g. the trainer's own contrastive weight (Trainer.py:24-26)
        super().add_param_group(param_group)
        self._arenas.append(self._build_arena(self.param_groups[-1]))

    def _build_arena(self, group) -> dict:
        plist: List[torch.nn.Parameter] = list(group["params"])
        if not plist:
            return {"params": [], "n": 0}
        dev = plist[0].device
        for p in plist:
            if not p.is_cuda:
                raise _lib.MmsaError("mmsa.FusedClipAdamW: parameters must live on a CUDA device (no CPU fallback)")
            if p.dtype != torch.float32:
                raise TypeError("mmsa.FusedClipAdamW: fp32 master parameters expected")
            if p.device != dev:
                raise _lib.MmsaError("mmsa.FusedClipAdamW: one device per parameter group")
        offs, n = arena_layout(plist)
        flat = torch.zeros(n, device=dev, dtype=torch.float32)
        views = []
        with torch.no_grad():
            for p, off in zip(plist, offs):
                v = flat[off:off + p.numel()].view_as(p)
                v.copy_(p.data)
                p.data = v                               # re-home the parameter into the arena
                views.append(v)
        a = {"params": plist, "n": n, "offs": offs, "p": flat, "views": views,
             "m": torch.zeros(n, device=dev, dtype=torch.float32),
             "v": torch.zeros(n, device=dev, dtype=torch.float32),
             "g": torch.zeros(n, device=dev, dtype=torch.float32),
             "steps": [0] * len(plist),
             "sq": torch.zeros(1, device=dev, dtype=torch.float32),
             "partials": torch.empty(512, device=dev, dtype=torch.float32)}
        a["ends"] = offs[1:] + [n]
        a["gviews"] = [a["g"][off:off + p.numel()].view_as(p) for p, off in zip(plist, offs)]
        a["mviews"] = [a["m"][off:off + p.numel()].view_as(p) for p, off in zip(plist, offs)]
        a["vviews"] = [a["v"][off:off + p.numel()].view_as(p) for p, off in zip(plist, offs)]
        return a

    def _flat_grad(self, a: dict) -> Tuple[Tensor, List[bool]]:
        """the group's gradients as one flat fp32 tensor in parameter order (in place when they already are one), and
        which parameters have a gradient this step."""
        plist = a["params"]
        grads = [p.grad for p in plist]
        has = [g is not None for g in grads]
        first = grads[0]
        if first is not None and first._base is not None and first._base.dtype == torch.float32 and first._base.ndim == 1:
            base, start, ok = first._base, first.storage_offset(), True
            for g, off in zip(grads, a["offs"]):
                if g is None or g._base is not base or g.storage_offset() != start + off or not g.is_contiguous():
                    ok = False
                    break
            if ok and start + a["n"] <= base.numel():
                return base[start:start + a["n"]], has
        live_dst = [d for d, g in zip(a["gviews"], grads) if g is not None]
        live_src = [g for g in grads if g is not None]
        if len(live_src) != len(grads):
            a["g"].zero_()                               # ranges without a gradient: zero for the norm, skipped by the update
        if live_src:
            torch._foreach_copy_(live_dst, live_src)     # multi-tensor pack (plumbing)
        return a["g"], has

    def _ensure_state(self, a: dict, i: int) -> None:
        p = a["params"][i]
        st = self.state[p]
        if "exp_avg" not in st:
            st["step"] = torch.tensor(0.0, dtype=torch.float32)
            st["exp_avg"] = a["mviews"][i]
            st["exp_avg_sq"] = a["vviews"][i]

    # ------------------------------------------------------------------ step
    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        work = []
        for gi, group in enumerate(self.param_groups):
            a = self._arenas[gi]
            if not a["n"]:
                continue
            for p, v in zip(a["params"], a["views"]):
                if p.data_ptr() != v.data_ptr():
                    raise _lib.MmsaError("mmsa.FusedClipAdamW: a parameter's storage left its arena (model.to(...) or "
                                         "p.data = ... after the optimiser was built); rebuild the optimiser")
            g, has = self._flat_grad(a)
            if any(has):
                work.append((group, a, g, has))
        if not work:
            return loss
        dev = work[0][1]["p"].device
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream(dev).cuda_stream
            # global gradient norm over the clip groups (clip_grad_norm_(model.parameters()) semantics)
            total_sq = None
            for group, a, g, has in work:
                if not group.get("clip", True) or self.max_norm is None:
                    continue
                _lib.call("mmsa_sumsq", g.data_ptr(), a["n"], a["partials"].data_ptr(), 512, a["sq"].data_ptr(), st)
                total_sq = a["sq"] if total_sq is None else total_sq.add_(a["sq"])
            if total_sq is None:
                total_sq = work[0][1]["sq"].zero_()
            for group, a, g, has in work:
                clip = group.get("clip", True) and self.max_norm is not None
                max_norm = float(self.max_norm) if clip else 3.0e38
                b1, b2 = group["betas"]
                steps, offs, ends = a["steps"], a["offs"], a["ends"]
                n = len(has)
                i = 0
                while i < n:                         # runs of neighbouring parameters with a gradient and equal step counts
                    if not has[i]:
                        i += 1
                        continue
                    j = i
                    while j + 1 < n and has[j + 1] and steps[j + 1] == steps[i]:
                        j += 1
                    for k in range(i, j + 1):
                        steps[k] += 1
                        self._ensure_state(a, k)
                    lo, hi = offs[i], ends[j]
                    _lib.call("mmsa_clip_adamw", a["p"].data_ptr() + 4 * lo, g.data_ptr() + 4 * lo,
                              a["m"].data_ptr() + 4 * lo, a["v"].data_ptr() + 4 * lo, hi - lo, total_sq.data_ptr(),
                              max_norm, float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                              float(group["weight_decay"]), steps[i], st)
                    i = j + 1
        # the kernels write the parameters behind autograd's back (no version bump): mark the cached bf16 operand
        # copies stale so that the next forward (or model.prepare_step()) re-casts them, into the same buffers
        from . import ops
        ops.bump_weights_epoch()
        return loss

    # ------------------------------------------------------------------ checkpoints
    def state_dict(self):
        for a in self._arenas:
            for i, p in enumerate(a.get("params", [])):
                if p in self.state and "step" in self.state[p]:
                    self.state[p]["step"] = torch.tensor(float(a["steps"][i]), dtype=torch.float32)
        return super().state_dict()

    @torch.no_grad()
    def load_state_dict(self, state_dict) -> None:
        clips = [g.get("clip", True) for g in self.param_groups]
        super().load_state_dict(state_dict)              # replaces self.state (fresh tensors) and the group dicts
        for g, c in zip(self.param_groups, clips):       # a torch.optim.AdamW checkpoint carries no clip flags
            g.setdefault("clip", c)
        for a in self._arenas:
            if not a["n"]:
                continue
            a["m"].zero_(); a["v"].zero_()
            for i, p in enumerate(a["params"]):
                st = self.state.get(p)
                if not st or "exp_avg" not in st:
                    a["steps"][i] = 0
                    continue
                a["mviews"][i].copy_(st["exp_avg"])
                a["vviews"][i].copy_(st["exp_avg_sq"])
                a["steps"][i] = int(round(float(st["step"])))
                st["exp_avg"], st["exp_avg_sq"] = a["mviews"][i], a["vviews"][i]
                st["step"] = torch.tensor(float(a["steps"][i]), dtype=torch.float32)
