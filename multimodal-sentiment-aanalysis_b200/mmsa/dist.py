"""Data parallelism for the fusion path: one process per GPU, torch.distributed (NCCL over
NVLink/NVSwitch on the GPU box, gloo in CPU tests).

The path shards by batch.  Only three exchanges exist:
  1. all-gather of the column-side contrastive embeddings [B,E] -> [B_g,E] and labels [B] -> [B_g];
  2. reduce-scatter (sum) of the gradient w.r.t. the gathered embeddings back to their owners;
  3. all-reduce (mean) of the replicated parameter gradients (flat buckets, overlappable).
BatchNorm statistics stay per shard (DDP-without-SyncBN semantics)."""
from __future__ import annotations

import os
from typing import Iterable, List, Optional, Tuple

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


def sharded_infonce(f1: Tensor, f2: Tensor, labels: Tensor, temperature, group=None, fast: bool = False,
                    weight: Optional[Tensor] = None) -> Tensor:
    """InfoNCE (MultimodalModel.py:232-260) over the GLOBAL batch with rows sharded by rank:
    this rank owns rows [rank*B, (rank+1)*B) of the B_g x B_g similarity matrix and all its columns.
    Returns the local mean over this rank's rows; averaging parameter gradients over ranks then
    yields the gradient of the global-batch mean."""
    from . import ops
    rank = dist.get_rank(group)
    f2_all = all_gather_rows(f2, group)
    labels_all = gather_labels(labels, group)
    return ops.infonce(f1, f2_all, labels, temperature, labels_cols=labels_all, row_offset=rank * f1.shape[0], fast=fast,
                       weight=weight)


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
    a = ops.ContrastiveFn.apply(z1, z_all, lab_rows, lab_cols, None, float(temperature), kind, rank * B, 2 * B, False, False,
                                None)
    b = ops.ContrastiveFn.apply(z2, z_all, lab_rows, lab_cols, None, float(temperature), kind, Bg + rank * B, 2 * B, False,
                                False, None)
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
    fp32 at E=768).  Gradients are packed into one flat arena so NCCL sees few large messages
    (launch-latency-bound over NVSwitch, not link-bound).

    `overlap=True` (or MMSA_DP_OVERLAP=1) splits the arena in two buckets by the order in which the gradients become
    ready (recorded on the first backward through post-accumulate-grad hooks): the EARLY bucket -- everything but the
    last `late_frac` of the bytes, i.e. all but the input-projection weights, whose gradients come last -- is packed
    and all-reduced on a communication stream the moment its last gradient lands, under the remaining backward
    kernels; `step()` then only reduces the LATE bucket.  Results are identical to the one-bucket form (same arena,
    same element-wise mean; checked at N = 2 on B200: same loss to the last bit).  Works under CUDA-graph capture: the
    hooks run during capture and the communication stream forks from / joins the capturing stream.
    OFF by default: in this first form the two buckets came out as 4 + 3 arena ranges (7 all-reduces instead of 1) and
    the NCCL kernels compete with the persistent GEMMs for SMs -- 2.085 ms against 1.900 ms per step at N = 2.  What it
    needs next is an arena laid out in landing order (one range per bucket) and a CTA cap on the NCCL kernels."""

    def __init__(self, params: Iterable[torch.nn.Parameter], group=None, bucket_mb: float = 64.0,
                 overlap: Optional[bool] = None, late_frac: float = 0.25):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.group = group
        self.bucket_bytes = int(bucket_mb * (1 << 20))
        self._arenas: Optional[List[Tensor]] = None
        self._flat: Optional[Tensor] = None
        self._flat_views: List[Tensor] = []
        self._flat_key = None
        self._slices = None
        self.overlap = (os.environ.get("MMSA_DP_OVERLAP", "0") == "1") if overlap is None else bool(overlap)
        self.late_frac = float(late_frac)
        self._order: List[int] = []                    # parameter indices in the order their gradients landed
        self._recording = True                         # first backward: the hooks only record that order
        self._fired: List[bool] = [False] * len(self.params)
        self._early: Optional[List[int]] = None        # parameter indices of the early bucket (None: not planned yet)
        self._early_spans: List[Tuple[int, int]] = []
        self._late_spans: List[Tuple[int, int]] = []
        self._pending_early = 0
        self._early_launched = False
        self._comm = None
        self.overlapped_steps = 0                      # steps whose early bucket went out under the backward (tests)
        if self.overlap:
            for i, p in enumerate(self.params):
                p.register_post_accumulate_grad_hook(lambda _p, i=i: self._on_grad(i))

    # ---- overlap mode ----
    def _comm_stream(self, device):
        if self._comm is None:
            self._comm = torch.cuda.Stream(device=device, priority=-1)
        return self._comm

    def _reduce_spans(self, spans, world: int):
        avg = dist.get_backend(self.group) == "nccl"
        for lo, hi in spans:
            seg = self._flat if (lo == 0 and hi == self._flat.numel()) else self._flat[lo:hi]
            if avg:
                dist.all_reduce(seg, op=dist.ReduceOp.AVG, group=self.group)
            else:                                      # gloo (CPU tests): no AVG
                dist.all_reduce(seg, group=self.group)
                seg.div_(world)

    def _on_grad(self, i: int):
        if self._fired[i]:
            return
        self._fired[i] = True
        p = self.params[i]
        if self._recording:
            self._order.append(i)
            return
        if self._early is None:                        # no two-bucket plan (yet): step() reduces everything
            return
        if p.is_cuda:                                  # whatever stream this gradient was accumulated on
            self._comm_stream(p.device).wait_stream(torch.cuda.current_stream(p.device))
        if i in self._early_set:
            self._pending_early -= 1
            if self._pending_early == 0 and not self._early_launched:
                self._launch_early()

    @torch.no_grad()
    def _launch_early(self):
        world = dist.get_world_size(self.group)
        ps = [self.params[i] for i in self._early]
        views = [self._view_of[i] for i in self._early]
        if self._flat.is_cuda:
            with torch.cuda.stream(self._comm_stream(self._flat.device)):
                torch._foreach_copy_(views, [p.grad for p in ps])
                self._reduce_spans(self._early_spans, world)
        else:
            torch._foreach_copy_(views, [p.grad for p in ps])
            self._reduce_spans(self._early_spans, world)
        self._early_launched = True

    def _plan_overlap(self, live: List[torch.nn.Parameter], offs: List[int]):
        """early / late split from the recorded landing order; spans = maximal contiguous arena ranges per bucket."""
        idx_of = {id(p): i for i, p in enumerate(self.params)}
        live_idx = [idx_of[id(p)] for p in live]
        # every rank must cut the arena at the same place: rank 0's landing order is the one that counts
        order_t = torch.full((len(self.params),), -1, dtype=torch.int64, device=self._flat.device)
        order_t[:len(self._order)] = torch.tensor(self._order, dtype=torch.int64)
        src = dist.get_global_rank(self.group, 0) if self.group is not None else 0
        dist.broadcast(order_t, src=src, group=self.group)
        live_set = set(live_idx)
        order = [i for i in order_t.tolist() if i >= 0 and i in live_set]
        if sorted(order) != sorted(live_idx):          # hooks did not see every gradient: stay with one bucket
            self._early = None
            return
        total = sum(self.params[i].numel() for i in order)
        pos = {i: k for k, i in enumerate(live_idx)}    # position in the arena (parameter order)
        late, acc = [], 0
        for i in reversed(order):
            n = self.params[i].numel()
            if late and acc + n > self.late_frac * total:
                break
            late.append(i)
            acc += n
        # the cut usually falls inside a group of gradients that land together (one autograd node returning many of
        # them); drop the stragglers of that group that are not arena neighbours of the rest, so that the late bucket
        # stays one contiguous range (every extra range is one more small all-reduce in the exposed tail)
        while len(late) > 1:
            k = pos[late[-1]]
            others = {pos[j] for j in late[:-1]}
            if (k - 1) in others or (k + 1) in others:
                break
            late.pop()
        late_set = set(late)
        self._early = [i for i in order if i not in late_set]
        self._early_set = set(self._early)
        self._view_of = {i: v for i, v in zip(live_idx, self._flat_views)}
        ext = {i: (off, off + self.params[i].numel()) for i, off in zip(live_idx, offs)}

        def spans(members):
            out: List[List[int]] = []
            for i in live_idx:                          # arena (parameter) order; alignment gaps ride along
                if i not in members:
                    continue
                lo, hi = ext[i]
                prev = live_idx[live_idx.index(i) - 1] if live_idx.index(i) > 0 else None
                if out and prev is not None and prev in members:
                    out[-1][1] = hi
                else:
                    out.append([lo, hi])
            return [(a, b) for a, b in out]
        self._early_spans = spans(self._early_set)
        self._late_spans = spans(late_set)
        if not self._early:
            self._early = None
        if os.environ.get("MMSA_DP_DEBUG") and dist.get_rank(self.group) == 0:
            import sys
            nb = lambda idx: sum(self.params[i].numel() for i in idx) * 4 / 1e6
            print(f"GradAllReducer: early {len(self._early or [])} tensors {nb(self._early or []):.1f} MB in {len(self._early_spans)} "
                  f"range(s); late {len(late)} tensors {nb(late):.1f} MB in {len(self._late_spans)} range(s)", file=sys.stderr, flush=True)

    def _early_view_ids(self):
        return {id(self._view_of[i]) for i in self._early}

    def _reset_step(self):
        self._recording = False
        self._fired = [False] * len(self.params)
        self._pending_early = len(self._early) if self._early is not None else 0
        self._early_launched = False

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
        """Average the parameter gradients over the group.  NCCL (and overlap mode on any backend): one flat arena,
        multi-tensor pack, all-reduce with op AVG (gloo: SUM and divide), .grad re-pointed at the arena views.  gloo
        without overlap (CPU tests of the bucket bookkeeping): packed per-bucket arenas, SUM, divide."""
        world = dist.get_world_size(self.group)
        if dist.get_backend(self.group) == "nccl" or self.overlap:
            # ONE flat fp32 arena: a multi-tensor pack (torch._foreach_copy_, plumbing), all-reduce with op AVG (no
            # division kernel; large messages are launch-latency-optimal over NVSwitch), and .grad re-pointed at
            # the arena views (no unpack).
            live = [p for p in self.params if p.grad is not None]
            if not live:
                self._reset_step()
                return
            key = tuple(id(p) for p in live)
            if self._flat is None or self._flat_key != key:
                from .optim import arena_layout          # same (256-byte aligned) layout as FusedClipAdamW's arenas
                offs, total = arena_layout(live)
                self._flat = torch.zeros(total, device=live[0].device, dtype=torch.float32)
                self._flat_views = [self._flat[off:off + p.numel()].view_as(p) for p, off in zip(live, offs)]
                self._flat_key = key
                self._early = None
                if self.overlap and self._order:
                    self._plan_overlap(live, offs)
                self._early_launched = False             # this step's gradients are not in the (new) arena yet
            cuda = self._flat.is_cuda
            if self._early is not None and self._early_launched:
                # the early bucket is already in flight on the communication stream: add the late one and join
                late = [p for p, v in zip(live, self._flat_views) if id(v) not in self._early_view_ids()]
                late_views = [v for v in self._flat_views if id(v) not in self._early_view_ids()]
                if cuda:
                    cur = torch.cuda.current_stream(self._flat.device)
                    comm = self._comm_stream(self._flat.device)
                    comm.wait_stream(cur)
                    with torch.cuda.stream(comm):
                        if late:
                            torch._foreach_copy_(late_views, [p.grad for p in late])
                            self._reduce_spans(self._late_spans, world)
                    cur.wait_stream(comm)
                elif late:
                    torch._foreach_copy_(late_views, [p.grad for p in late])
                    self._reduce_spans(self._late_spans, world)
                self.overlapped_steps += 1
            else:
                if cuda and self._comm is not None:      # hooks may have queued waits on it: keep it joined
                    torch.cuda.current_stream(self._flat.device).wait_stream(self._comm)
                torch._foreach_copy_(self._flat_views, [p.grad for p in live])
                self._reduce_spans([(0, self._flat.numel())], world)
            for p, v in zip(live, self._flat_views):
                p.grad = v
            self._reset_step()
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


class ArenaGradReducer:
    """Parameter-gradient mean all-reduce with the gradients LANDING IN the communication arena and leaving in buckets
    under the rest of the backward (the default reducer of bench.py at N > 1).

    * One flat fp32 arena in PARAMETER order (mmsa.optim.arena_layout, the layout FusedClipAdamW consumes in place), built
      up front.  It is registered as the kernels' gradient SINK (mmsa.ops.set_grad_sink): the fusion core's backward has
      its weight-gradient GEMMs / LayerNorm reductions write straight into the arena views and returns no tensor for
      those parameters -- the 81 MB multi-tensor pack of the one-bucket form disappears (the ~0.8 M parameters of the [B,*]
      tail still arrive through autograd and are packed, 3 MB).
    * Default (`early_buckets=False`): `step()` sends the whole arena in ONE ncclAllReduce (op AVG) after the last weight
      gradient, joins, and points `.grad` at the arena views.
    * `early_buckets=True` (MMSA_DP_BUCKETS=1): the core reports each group of gradients the moment its kernels are ENQUEUED
      (`bucket_done(params, stream)`: tail first, then block e2p, block p2e, the input projections).  A group is a
      contiguous arena range, so it is one all-reduce, issued on a communication stream that waits for the producing stream
      -- inside the captured graph a parallel branch under the remaining dgrad / attention / wgrad kernels; `step()` then
      reduces only what is left.  Measured SLOWER than the single all-reduce at N = 2 and N = 8 on B200 / NVSwitch
      (profiles/r02_dp_matrix.md): the NCCL kernels take SMs from persistent GEMMs that own the whole chip, and capping
      NCCL's CTAs costs more than it saves -- hence not the default.
    Results equal the one-bucket form (same element-wise mean).  Gradient ACCUMULATION over several backwards is not
    supported in this mode (a second backward overwrites the arena): use GradAllReducer for that."""

    def __init__(self, params: Iterable[torch.nn.Parameter], group=None, register_sink: bool = True,
                 early_buckets: Optional[bool] = None):
        from .optim import arena_layout
        # early_buckets=False (MMSA_DP_BUCKETS=0): gradients still land in the arena (no pack), but everything leaves in
        # ONE all-reduce from step() -- for fabrics / world sizes where NCCL kernels beside the GEMMs cost more than the
        # exposed tail saves
        self.early_buckets = (os.environ.get("MMSA_DP_BUCKETS", "0") == "1") if early_buckets is None else bool(early_buckets)
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.group = group
        dev = self.params[0].device
        self.offs, total = arena_layout(self.params)
        self._flat = torch.zeros(total, device=dev, dtype=torch.float32)
        self.views = [self._flat[o:o + p.numel()].view_as(p) for p, o in zip(self.params, self.offs)]
        self._index = {id(p): i for i, p in enumerate(self.params)}
        self._ends = self.offs[1:] + [total]
        self._reduced = [False] * len(self.params)      # this step: already inside an all-reduce that was launched
        self._direct = [False] * len(self.params)       # this step: written into the arena by the kernels
        self._comm = None
        self.launched: List[Tuple[int, int]] = []       # arena ranges all-reduced this step, in launch order (tests)
        self.early_bytes = 0                            # bytes that went out before step() in the last step (tests)
        if register_sink:
            from . import ops
            ops.set_grad_sink(self)
        if self._flat.is_cuda:      # autograd may accumulate a gradient on a side stream: the communication stream waits for it
            for p in self.params:
                p.register_post_accumulate_grad_hook(self._on_accumulate)

    def _on_accumulate(self, p) -> None:
        self._comm_stream().wait_stream(torch.cuda.current_stream(p.device))

    # ---- sink protocol (called by mmsa.ops) ----
    def sink(self, p) -> Optional[Tensor]:
        i = self._index.get(id(p))
        if i is None:
            return None
        self._direct[i] = True
        return self.views[i]

    def _comm_stream(self):
        if self._comm is None and self._flat.is_cuda:
            self._comm = torch.cuda.Stream(device=self._flat.device, priority=-1)
        return self._comm

    def _all_reduce(self, lo: int, hi: int):
        seg = self._flat if (lo == 0 and hi == self._flat.numel()) else self._flat[lo:hi]
        if dist.get_backend(self.group) == "nccl":
            dist.all_reduce(seg, op=dist.ReduceOp.AVG, group=self.group)
        else:                                           # gloo (CPU tests): no AVG
            dist.all_reduce(seg, group=self.group)
            seg.div_(dist.get_world_size(self.group))
        self.launched.append((lo, hi))

    @torch.no_grad()
    def _flush(self, idx: List[int], streams) -> None:
        """all-reduce the arena ranges of parameters `idx` (sorted; neighbours merge into one call) on the communication
        stream, after everything enqueued so far on `streams`; autograd-delivered gradients among them are packed first"""
        idx = [i for i in idx if not self._reduced[i]]
        if not idx:
            return
        comm = self._comm_stream()
        if comm is not None:
            for st in streams:
                if st is not None:
                    comm.wait_stream(st)
        ctx = torch.cuda.stream(comm) if comm is not None else _NullCtx()
        with ctx:
            src, dst = [], []
            for i in idx:
                g = self.params[i].grad
                if not self._direct[i]:
                    if g is None:
                        self.views[i].zero_()           # a parameter without gradient on this rank counts as zero
                    elif g.data_ptr() != self.views[i].data_ptr():
                        src.append(g); dst.append(self.views[i])
            if src:
                torch._foreach_copy_(dst, src)
            lo, hi = self.offs[idx[0]], self._ends[idx[0]]
            for a, b in zip(idx, idx[1:]):
                if b == a + 1:
                    hi = self._ends[b]
                else:
                    self._all_reduce(lo, hi)
                    lo, hi = self.offs[b], self._ends[b]
            self._all_reduce(lo, hi)
        for i in idx:
            self._reduced[i] = True

    def bucket_done(self, params, stream=None) -> None:
        """the gradients of `params` have just been enqueued (on `stream`, default: the current stream): send their arena
        range(s) off.  Every parameter BEFORE them in landing order whose gradient autograd has already delivered (the
        tail) rides along on the first call."""
        idx = sorted({self._index[id(p)] for p in params if id(p) in self._index})
        if not idx:                                     # another model's backward (e.g. a single-GPU yardstick): not ours
            return
        if not self.early_buckets:
            return
        cur = torch.cuda.current_stream(self._flat.device) if self._flat.is_cuda else None
        if not self.launched:                           # first call of the step: the autograd-delivered tail goes too
            tail = [i for i, p in enumerate(self.params) if not self._direct[i] and p.grad is not None]
            self._flush(sorted(tail), (cur,))
        before = len(self.launched)
        self._flush(idx, (stream, cur))
        self.early_bytes += sum((hi - lo) * 4 for lo, hi in self.launched[before:])

    @torch.no_grad()
    def step(self):
        cur = torch.cuda.current_stream(self._flat.device) if self._flat.is_cuda else None
        early = sum((hi - lo) * 4 for lo, hi in self.launched)
        rest = [i for i in range(len(self.params)) if not self._reduced[i] and (self._direct[i] or self.params[i].grad is not None)]
        self._flush(rest, (cur,))
        if self._comm is not None:
            cur.wait_stream(self._comm)
        for i, p in enumerate(self.params):
            if self._reduced[i]:
                p.grad = self.views[i]
        self.early_bytes = early
        self.last_launched = list(self.launched)
        self.launched = []
        self._reduced = [False] * len(self.params)
        self._direct = [False] * len(self.params)


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


# --------------------------------------------------------------------------------------------- parity yardstick
def emulate_data_parallel_step(model, text_all: Tensor, image_all: Tensor, labels_all: Tensor, world: int):
    """The N-rank data-parallel step evaluated on ONE device with the same kernels -- the yardstick `bench.py` (dp_parity)
    and the multi-rank tests hold the sharded run against.  Rank r's loss is CE over ITS shard (BatchNorm statistics per
    shard: DDP-without-SyncBN semantics) plus the InfoNCE of its row block [r*B, (r+1)*B) against ALL columns of the global
    batch (MultimodalModel.py:232-260 with the global diagonal); averaging parameter gradients over ranks is the gradient
    of mean_r(loss_r).  `model` must be a bidirectional single-contract MultimodalTransformerModel with dp_group None
    (typically a fresh copy of the replicated weights: the per-shard tails update its BatchNorm running statistics).
    Returns (mean loss, [per-rank contrastive values], {parameter name: gradient})."""
    from . import ops
    assert model.wiring == "bidirectional" and model.contract == "single" and model.dp_group is None
    Bg = text_all.shape[0]
    assert Bg % world == 0
    B = Bg // world
    model.zero_grad(set_to_none=True)
    model.prepare_step()
    fast = model.compute_dtype == torch.bfloat16
    f0, fv, e1, e2, _, _ = ops.fusion_core(model._cd(text_all), model._cd(image_all), model.eeg_net.proj.weight,
                                     model.eeg_net.proj.bias, model.eye_net.proj.weight, model.eye_net.proj.bias,
                                     model.num_heads, model.cross_attn_e2p.kernel_params(),
                                     model.cross_attn_p2e.kernel_params())
    total, contrastive = None, []
    for r in range(world):
        sl = slice(r * B, (r + 1) * B)
        lab = labels_all[sl].contiguous()
        arousal, _ = model._tail(f0[sl], fv[sl], (f0[sl], e1[sl], e2[sl]))
        c = ops.infonce(e1[sl], e2, lab, model.temperature, labels_cols=labels_all.contiguous(), row_offset=r * B, fast=fast)
        loss_r = ops.cross_entropy(arousal, lab) + (model.contrastive_weight * c).sum()
        contrastive.append(c.detach())
        total = loss_r if total is None else total + loss_r
    total = total / world
    total.backward()
    model._drop.commit(text_all.device)
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    return total.detach(), contrastive, grads
