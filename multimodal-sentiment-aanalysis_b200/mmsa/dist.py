"""Data parallelism for the fusion path: one process per GPU, torch.distributed (NCCL over
NVLink/NVSwitch on the GPU box, gloo in CPU tests).

The path shards by batch.  Only three exchanges exist:
  1. all-gather of the column-side contrastive embeddings [B,E] -> [B_g,E] and labels [B] -> [B_g];
  2. reduce-scatter (sum) of the gradient w.r.t. the gathered embeddings back to their owners;
  3. all-reduce (mean) of the replicated parameter gradients (flat buckets, overlappable).
BatchNorm statistics stay per shard (DDP-without-SyncBN semantics)."""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist
from torch.autograd import Function

Tensor = torch.Tensor


class AllGatherRows(Function):
    """y = cat_r x_r along dim 0; backward = reduce-scatter(sum) of dy to the owning rank."""

    @staticmethod
    def forward(ctx, x: Tensor, group):
        ctx.group = group
        world = dist.get_world_size(group)
        ctx.rows = x.shape[0]
        x = x.contiguous()
        out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), device=x.device, dtype=x.dtype)
        dist.all_gather_into_tensor(out, x, group=group)
        return out

    @staticmethod
    def backward(ctx, dy: Tensor):
        dy = dy.contiguous()
        dx = torch.empty((ctx.rows,) + tuple(dy.shape[1:]), device=dy.device, dtype=dy.dtype)
        if dist.get_backend(ctx.group) == "gloo":     # gloo has no reduce_scatter_tensor
            tmp = dy.clone()
            dist.all_reduce(tmp, group=ctx.group)
            r = dist.get_rank(ctx.group)
            dx.copy_(tmp[r * ctx.rows:(r + 1) * ctx.rows])
        else:
            dist.reduce_scatter_tensor(dx, dy, group=ctx.group)
        return dx, None


def all_gather_rows(x: Tensor, group=None) -> Tensor:
    return AllGatherRows.apply(x, group)


def gather_labels(labels: Tensor, group=None) -> Tensor:
    world = dist.get_world_size(group)
    out = torch.empty((world * labels.shape[0],), device=labels.device, dtype=labels.dtype)
    dist.all_gather_into_tensor(out, labels.contiguous(), group=group)
    return out


def sharded_infonce(f1: Tensor, f2: Tensor, labels: Tensor, temperature, group=None, fast: bool = False) -> Tensor:
    """InfoNCE (MultimodalModel.py:232-260) over the GLOBAL batch with rows sharded by rank:
    this rank owns rows [rank*B, (rank+1)*B) of the B_g x B_g similarity matrix and all its columns.
    Returns the local mean over this rank's rows; averaging parameter gradients over ranks then
    yields the gradient of the global-batch mean."""
    from . import ops
    rank = dist.get_rank(group)
    f2_all = all_gather_rows(f2, group)
    labels_all = gather_labels(labels, group)
    return ops.infonce(f1, f2_all, labels, temperature, labels_cols=labels_all, row_offset=rank * f1.shape[0], fast=fast)


def _sharded_two_view(z1: Tensor, z2: Tensor, labels: Optional[Tensor], temperature: float, kind: int, group=None) -> Tensor:
    """SupCon (train.py:16-40) / NT-Xent (ME-MHACL/train.py:47-66) over the GLOBAL stacked batch [all first views; all
    second views] with rows sharded by rank: both views are all-gathered (reduce-scatter in the backward), this rank
    scores its B first-view rows (global offset rank*B) and its B second-view rows (offset Bg + rank*B) against all 2*Bg
    columns, and returns the mean over its 2B rows -- averaging gradients over ranks then gives the global-batch mean."""
    from . import ops
    rank = dist.get_rank(group)
    B = z1.shape[0]
    z1_all = all_gather_rows(z1, group)
    z2_all = all_gather_rows(z2, group)
    Bg = z1_all.shape[0]
    z_all = ops._StackFn.apply(z1_all, z2_all)
    if labels is not None:
        lab_all = gather_labels(labels.view(-1), group)
        lab_cols = torch.cat([lab_all, lab_all])
        lab_rows = labels.view(-1)
    else:
        lab_cols = lab_rows = None
    a = ops.ContrastiveFn.apply(z1, z_all, lab_rows, lab_cols, None, float(temperature), kind, rank * B, 2 * B, False, False)
    b = ops.ContrastiveFn.apply(z2, z_all, lab_rows, lab_cols, None, float(temperature), kind, Bg + rank * B, 2 * B, False, False)
    return a + b


def sharded_supcon(z1: Tensor, z2: Tensor, labels: Tensor, temperature: float = 0.1, group=None) -> Tensor:
    from ._lib import LOSS_SUPCON
    return _sharded_two_view(z1, z2, labels, temperature, LOSS_SUPCON, group)


def sharded_ntxent(z1: Tensor, z2: Tensor, temperature: float = 0.5, group=None) -> Tensor:
    from ._lib import LOSS_NTXENT
    return _sharded_two_view(z1, z2, None, temperature, LOSS_NTXENT, group)


def shard_contrastive(model, group=None):
    """Switch a MultimodalTransformerModel to the batch-sharded contrastive loss."""
    model.dp_group = group if group is not None else dist.group.WORLD
    return model


class GradAllReducer:
    """Flat-bucket mean all-reduce of parameter gradients (the 10.18 M fusion parameters = 40.7 MB
    fp32 at E=768).  Gradients are packed into one arena per bucket so NCCL sees few large
    messages (launch-latency-bound over NVSwitch, not link-bound)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], group=None, bucket_mb: float = 64.0):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.group = group
        self.bucket_bytes = int(bucket_mb * (1 << 20))
        self._arenas: Optional[List[Tensor]] = None
        self._flat: Optional[Tensor] = None
        self._flat_views: List[Tensor] = []
        self._flat_key = None
        self._slices = None

    def _plan(self):
        arenas, slices, cur, cur_n = [], [], [], 0
        for p in self.params:
            n = p.numel()
            if cur and (cur_n + n) * 4 > self.bucket_bytes:
                arenas.append(cur_n)
                cur, cur_n = [], 0
            slices.append((len(arenas), cur_n, n))
            cur.append(p)
            cur_n += n
        arenas.append(cur_n)
        dev = self.params[0].device
        self._arenas = [torch.zeros(n, device=dev, dtype=torch.float32) for n in arenas]
        self._slices = slices

    @torch.no_grad()
    def step(self):
        """Average the parameter gradients over the group.  NCCL: ONE coalesced group call over the gradient tensors
        where they lie (ncclGroupStart/End: no packing copies), reduction op AVG (no division kernel).  Other
        backends (gloo, CPU tests): packed arenas, SUM, divide."""
        world = dist.get_world_size(self.group)
        if dist.get_backend(self.group) == "nccl":
            # ONE flat fp32 arena: a multi-tensor pack (torch._foreach_copy_, plumbing), ONE all-reduce with op AVG (no
            # division kernel; a single large message is launch-latency-optimal over NVSwitch), and .grad re-pointed at
            # the arena views (no unpack).
            live = [p for p in self.params if p.grad is not None]
            if not live:
                return
            key = tuple(id(p) for p in live)
            if self._flat is None or self._flat_key != key:
                from .optim import arena_layout          # same (256-byte aligned) layout as FusedClipAdamW's arenas
                offs, total = arena_layout(live)
                self._flat = torch.zeros(total, device=live[0].device, dtype=torch.float32)
                self._flat_views = [self._flat[off:off + p.numel()].view_as(p) for p, off in zip(live, offs)]
                self._flat_key = key
            torch._foreach_copy_(self._flat_views, [p.grad for p in live])
            dist.all_reduce(self._flat, op=dist.ReduceOp.AVG, group=self.group)
            for p, v in zip(live, self._flat_views):
                p.grad = v
            return
        if self._arenas is None:
            self._plan()
        for p, (a, off, n) in zip(self.params, self._slices):
            view = self._arenas[a][off:off + n]
            if p.grad is None:
                view.zero_()
            else:
                view.copy_(p.grad.reshape(-1))
        for arena in self._arenas:
            dist.all_reduce(arena, group=self.group)
        for p, (a, off, n) in zip(self.params, self._slices):
            g = self._arenas[a][off:off + n].view_as(p)
            if p.grad is None:
                p.grad = (g / world).clone()
            else:
                p.grad.copy_(g).div_(world)
