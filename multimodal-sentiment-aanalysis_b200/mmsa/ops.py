"""torch.autograd.Functions over the C ABI.  Every forward/backward below is a sequence of
mmsa_* calls; PyTorch supplies storage, streams and the autograd tape only.

Storage mode ("compute dtype", cd): activations of the token streams are fp32 (parity mode, 1e-5) or
bf16 (performance mode, fp32 accumulation, fp32 master parameters).  Everything that crosses an
autograd boundary on the small [B,*] side (pooled features, fused vector, logits, losses) is fp32 in
both modes; inside a Function the GEMM operands are cd and the GEMM outputs fp32 (include/mmsa.h)."""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import kernels as K
from ._lib import (ACT_GELU, ACT_NONE, ACT_RELU, ACT_SIGMOID, BN_ONLY, BN_THEN_GELU, LOSS_INFONCE, LOSS_NTXENT,
                   LOSS_SUPCON, RELU_THEN_BN)

Tensor = torch.Tensor

# ------------------------------------------------------------------------------------------ weight copies
# fp32 master weight -> bf16 operand copy.  prepare_weights() refreshes every copy of a model in ONE
# launch (call it once per step, after the optimiser update); _w() falls back to a single cast when a
# parameter was modified since (tensor._version), re-homed (data_ptr) or never prepared.
# The copy is an attribute OF THE PARAMETER (`_mmsa_wc` = (version, epoch, data_ptr, copy)): it dies with the
# parameter and can never be served to another tensor that happens to reuse a Python id.  Writes through `p.data`
# (dist.broadcast(p.data), an optimiser kernel) bump neither the version nor the address: optimisers call
# bump_weights_epoch(), and model.prepare_step() re-casts every copy unconditionally once per step.
_WEPOCH = 0


def bump_weights_epoch() -> None:
    """Mark every cached operand copy stale.  Called by optimisers that update parameters with their own kernels
    (no autograd version bump); the copies are re-cast INTO THE SAME buffers, so addresses captured in a CUDA graph
    stay valid."""
    global _WEPOCH
    _WEPOCH += 1


def _wc_get(p: Tensor, dtype: torch.dtype):
    ent = getattr(p, "_mmsa_wc", None)
    if ent is None or ent[3].shape != p.shape or ent[3].device != p.device or ent[3].dtype != dtype:
        return None
    return ent


def prepare_weights(params: Sequence[Tensor], dtype: torch.dtype) -> None:
    if dtype == torch.float32:
        return
    srcs, dsts = [], []
    for p in params:
        if p.ndim < 2 or not p.is_cuda:
            continue
        ent = _wc_get(p, dtype)
        buf = ent[3] if ent is not None else torch.empty(p.shape, device=p.device, dtype=dtype)
        srcs.append(p.detach())
        dsts.append(buf)
        p._mmsa_wc = (p._version, _WEPOCH, p.data_ptr(), buf)
    K.cast_multi(srcs, dsts)


def _w(w: Tensor, dtype: torch.dtype) -> Tensor:
    """operand copy of a weight in the compute dtype; identity in fp32 mode."""
    if dtype == torch.float32:
        wd = w.detach()
        return wd if wd.is_contiguous() else wd.contiguous()
    ent = _wc_get(w, dtype)
    if ent is not None and ent[0] == w._version and ent[1] == _WEPOCH and ent[2] == w.data_ptr():
        return ent[3]
    wd = w.detach()
    wd = wd if wd.is_contiguous() else wd.contiguous()
    if ent is not None:                           # stale: refresh in place (stable address)
        K.cast_multi([wd], [ent[3]])
        w._mmsa_wc = (w._version, _WEPOCH, w.data_ptr(), ent[3])
        return ent[3]
    return K.cast(wd, dtype)


# Gradient sink (data parallelism, mmsa.dist.ArenaGradReducer): when set, the fusion core's backward asks it for a
# preallocated fp32 destination per parameter (`sink(p)` -> tensor or None), has its kernels write the gradient THERE,
# returns no tensor for that parameter (autograd leaves p.grad alone; the sink's step() points it at the arena view)
# and reports each finished group with `bucket_done(params, stream)` so that the all-reduce can start under the rest of
# the backward.
_GRAD_SINK = None


def set_grad_sink(sink) -> None:
    global _GRAD_SINK
    _GRAD_SINK = sink


OVERLAP_WGRAD = True      # block weight gradients on a second stream (FusionCoreFn.backward)
OVERLAP_BLOCKS = os.environ.get("MMSA_OVERLAP_BLOCKS", "0") == "1"     # forward of block p2e beside block e2p (probe)
OVERLAP_TAIL = True       # tail weight gradients (SeqFn.backward) and the contrastive branch (model.forward) on a second stream


# Linear + BatchNorm1d(+act+dropout) blocks of the [B,*] tail as one launch when the batch fits one CTA (<= 256 rows);
# MMSA_LINEAR_BN=0 keeps the two-kernel form (GEMM, then mmsa_bn_act_fwd) for A/B timing
FUSE_LINEAR_BN = os.environ.get("MMSA_LINEAR_BN", "1") != "0"


def set_overlap(on: bool) -> None:
    """Switch every second-stream overlap on or off.  Off = one kernel at a time on the caller's stream: what
    bench.py's per-launch profiling pass needs (a kernel timed while another one shares the SMs is not a kernel time)."""
    global OVERLAP_WGRAD, OVERLAP_TAIL
    OVERLAP_WGRAD = OVERLAP_TAIL = bool(on)
_SIDE_STREAMS: Dict[int, "torch.cuda.Stream"] = {}


def critical_priority() -> int:
    """CUDA stream priority of the chains that feed the loss and its backward (-1 = high; the weight-gradient helper
    stream keeps the default priority 0, so pending critical-path kernels get SMs first).  MMSA_STREAM_PRIO=0 turns
    the distinction off (A/B measurements)."""
    return -1 if os.environ.get("MMSA_STREAM_PRIO", "1") != "0" else 0


def _side_stream(device) -> "torch.cuda.Stream":
    """one helper stream per device for work that hangs off the main chain (tail weight gradients)."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    st = _SIDE_STREAMS.get(idx)
    if st is None:
        st = torch.cuda.Stream(device=device)
        _SIDE_STREAMS[idx] = st
    return st


def _c(t: Tensor) -> Tensor:
    return t if t.is_contiguous() else t.contiguous()


class CastFn(Function):
    """dtype change with a gradient (fp32 <-> compute dtype); mmsa_cast both ways."""

    @staticmethod
    def forward(ctx, x, dtype):
        ctx.src = x.dtype
        return K.cast(_c(x), dtype)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        return K.cast(_c(dy), ctx.src), None


def cast(x: Tensor, dtype: torch.dtype) -> Tensor:
    return x if x.dtype == dtype else CastFn.apply(x, dtype)


# ------------------------------------------------------------------------------------------ Linear
class LinearFn(Function):
    """nn.Linear (MultimodalModel.py:86,172-198): y = x W^T + b over the last dim, operands in `cd`.
    x may be split in two feature segments (x, x2) to avoid materialising a concat."""

    @staticmethod
    def forward(ctx, x, x2, w, b, cd, out_fp32: bool):
        lead = x.shape[:-1]
        x2d = K.cast(_c(x).view(-1, x.shape[-1]), cd)
        x22d = None if x2 is None else K.cast(_c(x2).view(-1, x2.shape[-1]), cd)
        wc = _w(w, cd)
        od = torch.float32 if out_fp32 else cd
        y = K.linear_fwd(x2d, wc, None if b is None else b.detach(), x2=x22d, out_dtype=od)
        ctx.save_for_backward(x2d, x22d, wc)
        ctx.has_bias = b is not None
        ctx.in_shapes = (x.shape, None if x2 is None else x2.shape)
        ctx.in_dtypes = (x.dtype, None if x2 is None else x2.dtype)
        ctx.cd = cd
        return y.view(*lead, w.shape[0])

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x2d, x22d, wc = ctx.saved_tensors
        cd = ctx.cd
        dy2d = K.cast(_c(dy).view(-1, dy.shape[-1]), cd)
        Kx = x2d.shape[1]
        dx = dx2 = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = K.linear_dgrad(dy2d, wc[:, :Kx], out_dtype=ctx.in_dtypes[0]).view(ctx.in_shapes[0])
        if x22d is not None and ctx.needs_input_grad[1]:
            dx2 = K.linear_dgrad(dy2d, wc[:, Kx:], out_dtype=ctx.in_dtypes[1]).view(ctx.in_shapes[1])
        if ctx.needs_input_grad[2] or (ctx.has_bias and ctx.needs_input_grad[3]):
            want_w = ctx.needs_input_grad[2]
            want_b = ctx.has_bias and ctx.needs_input_grad[3]
            if x22d is None:
                dw, db = K.linear_wgrad(dy2d, x2d, want_bias=want_b, want_weight=want_w)
            else:
                dw = torch.empty((wc.shape[0], wc.shape[1]), device=dy.device, dtype=torch.float32)
                _, db = K.linear_wgrad(dy2d, x2d, dw=dw[:, :Kx], want_bias=want_b, want_weight=True)
                K.linear_wgrad(dy2d, x22d, dw=dw[:, Kx:], want_bias=False, want_weight=True)
                if not want_w:
                    dw = None
        return dx, dx2, dw, db, None, None


def linear(x: Tensor, w: Tensor, b: Optional[Tensor], x2: Optional[Tensor] = None, out_fp32: bool = False,
           cd: Optional[torch.dtype] = None) -> Tensor:
    return LinearFn.apply(x, x2, w, b, cd or x.dtype, out_fp32)


# ------------------------------------------------------------------------------------------ cross block
class _BlockState:
    """Tensors one CrossModalTransformer block keeps between forward and backward."""
    __slots__ = ("q_in", "kv_in", "v_in", "Qp", "KVp", "O", "lse", "A", "G", "mean", "rstd", "w_in", "w_out", "w_gate",
                 "gamma", "B", "Lq", "Lk", "E", "H", "pooled")


def _block_fwd(q_in: Tensor, kv_in: Tensor, B: int, Lq: int, Lk: int, H: int, in_w, in_b, out_w, out_b, gate_w,
               gate_b, ln_w, ln_b, eps: float = 1e-5, pooled: bool = False, v_in: Optional[Tensor] = None,
               keep: Optional[list] = None):
    """CrossModalTransformer.forward (MultimodalModel.py:124-149) on flattened [B*L, E] activations.
    MHA in-projection (q rows of in_proj_weight; packed k,v rows), attention core, out-projection,
    gate GEMM over the two operands [q | attn] (no concat), fused sigmoid/blend/LayerNorm.
    v_in: the `value` tensor when it is not the `key` tensor (forward(query, key, value), MultimodalModel.py:124): K and V
    are then projected by two N = E GEMMs into the two halves of the same packed buffer instead of one N = 2E GEMM.
    pooled=False -> (y [B*Lq,E], state); pooled=True -> ((mean_l y, mean_l q) fp32 [B,E] each, state)."""
    cd = q_in.dtype
    E = q_in.shape[1]
    st = _BlockState()
    st.w_in, st.w_out, st.w_gate = _w(in_w, cd), _w(out_w, cd), _w(gate_w, cd)
    in_b = in_b.detach()
    st.Qp = K.linear_fwd(q_in, st.w_in[:E], in_b[:E])
    if v_in is None:
        st.KVp = K.linear_fwd(kv_in, st.w_in[E:], in_b[E:])
    else:
        st.KVp = torch.empty((kv_in.shape[0], 2 * E), device=kv_in.device, dtype=cd)
        K.linear_fwd(kv_in, st.w_in[E:2 * E], in_b[E:2 * E], out=st.KVp[:, :E])
        K.linear_fwd(v_in, st.w_in[2 * E:], in_b[2 * E:], out=st.KVp[:, E:])
    st.O, st.lse = K.attn_fwd(st.Qp, st.KVp[:, :E], st.KVp[:, E:], B, H, Lq, Lk, E // H)
    st.A = K.linear_fwd(st.O, st.w_out, out_b.detach())
    gate_pre = K.linear_fwd(q_in, st.w_gate, gate_b.detach(), x2=st.A)
    st.gamma = ln_w.detach()
    st.pooled = pooled
    if pooled:
        st.G, st.mean, st.rstd, py, pq, pq_lp = K.gate_ln_pool_fwd(gate_pre, q_in, st.A, st.gamma, ln_b.detach(), eps, B, Lq,
                                                                   want_q_lp=True)
        out = (py, pq, pq_lp)      # pq_lp: the pooled query in the compute dtype (operand copy for the tail's first Linear)
    else:
        st.G, out, st.mean, st.rstd = K.gate_ln_fwd(gate_pre, q_in, st.A, st.gamma, ln_b.detach(), eps)
    st.q_in, st.kv_in, st.v_in = q_in, kv_in, v_in
    st.B, st.Lq, st.Lk, st.E, st.H = B, Lq, Lk, E, H
    if keep is not None:
        keep.append(gate_pre)          # temporaries of a block run on a helper stream stay alive until the caller joins
    return out, st


def _block_bwd(st: _BlockState, dy: Tensor, *, dpooled_q: Optional[Tensor] = None, dq_add: Optional[Tensor] = None,
               dkv_residual: Optional[Tensor] = None, need_dq: bool = True, need_dkv: bool = True, need_w: bool = True,
               wgrad_stream=None, keep: Optional[list] = None, sinks: Optional[Sequence[Optional[Tensor]]] = None):
    """Backward of _block_fwd.  dy is [B*Lq,E] (cd) for a plain block, or the fp32 [B,E] gradient of the
    pooled output for a pooled block (dpooled_q: fp32 [B,E] gradient of the pooled query stream).
    For a block with a separate value tensor (st.v_in) dkv is the pair (dkey, dvalue).
    Returns (dq, dkv, grads) with grads = (d_in_w, d_in_b, d_out_w, d_out_b, d_gate_w, d_gate_b, d_ln_w, d_ln_b).
    Every accumulation of gradients w.r.t. q and kv happens in GEMM residual epilogues / the LN kernel;
    bias gradients fall out of the wgrad GEMMs (ones-tile MMA).
    sinks: optional preallocated fp32 destinations for the eight parameter gradients (same order as `grads`; None
    entries are allocated here) -- the data-parallel gradient arena (set_grad_sink)."""
    E, B, Lq, Lk, H = st.E, st.B, st.Lq, st.Lk, st.H
    dev = dy.device
    sk = list(sinks) if sinks is not None else [None] * 8

    def buf(i, shape):
        return sk[i] if sk[i] is not None else torch.empty(shape, device=dev, dtype=torch.float32)
    # The gate's two input gradients share their A operand: dq = dgate W_g[:, :E] + r0 and dA = dgate W_g[:, E:] + r1.  With r0
    # and r1 written as the two column halves of ONE [M, 2E] buffer they are ONE N = 2E GEMM (residual = that buffer, output
    # [dq | dA]) instead of two N = E ones: 32768 x 1536 x 768 runs at 1 190 TFLOP/s against 1 040 for the 768-wide shape.
    fuse_gate_dgrad = need_dq and st.q_in.dtype == torch.bfloat16
    if st.pooled:
        r0, r1, dgate, d_ln_w, d_ln_b = K.gate_ln_pool_bwd(dy, dpooled_q, dq_add, st.G, st.q_in, st.A, st.gamma,
                                                           st.mean, st.rstd, B, Lq, dgamma=sk[6], dbeta=sk[7],
                                                           packed_parts=fuse_gate_dgrad)
    else:
        r0, r1, dgate, d_ln_w, d_ln_b = K.gate_ln_bwd(dy, 0, st.G, st.q_in, st.A, st.gamma, st.mean, st.rstd,
                                                      dq_add=dq_add, dgamma=sk[6], dbeta=sk[7], packed_parts=fuse_gate_dgrad)
    d_gate_w = d_gate_b = d_out_w = d_out_b = d_in_w = d_in_b = None
    # The weight-gradient GEMMs are leaves of the backward chain and their split-K clusters cover 108 of the 148 SMs:
    # they are enqueued on a second stream (joined by the caller, `wgrad_stream` is not None then), so the dgrad /
    # attention / LayerNorm kernels of the main chain fill the SMs they leave idle.
    main = torch.cuda.current_stream(dev)

    def on_side(fn):
        if wgrad_stream is None:
            fn()
            return
        wgrad_stream.wait_stream(main)
        with torch.cuda.stream(wgrad_stream):
            fn()

    if need_w:
        d_gate_w = buf(4, (E, 2 * E))
        d_gate_b = buf(5, (E,))

        def gate_wgrads():       # dW_gate = dgate^T [q | attn]: one launch over both input tensors (no concat)
            K.linear_wgrad2(dgate, st.q_in, st.A, dw=d_gate_w, db=d_gate_b)
        on_side(gate_wgrads)
    if fuse_gate_dgrad:
        r01 = r0._base if r0._base is not None else None          # the packed [M, 2E] buffer behind the two views
        both = K.linear_dgrad(dgate, st.w_gate, residual=r01)       # [dq_acc | dA]
        dq_acc, dA = both[:, :E], both[:, E:]
    else:
        dA = K.linear_dgrad(dgate, st.w_gate[:, E:], residual=r1)
        dq_acc = K.linear_dgrad(dgate, st.w_gate[:, :E], residual=r0) if need_dq else None
    if need_w:
        d_out_w = buf(2, (E, E))
        d_out_b = buf(3, (E,))
        on_side(lambda: K.linear_wgrad(dA, st.O, dw=d_out_w, db=d_out_b))
    dO = K.linear_dgrad(dA, st.w_out)
    dQp = torch.empty_like(st.Qp)
    dKVp = torch.empty_like(st.KVp)
    K.attn_bwd(st.Qp, st.KVp[:, :E], st.KVp[:, E:], st.O, dO, st.lse, B, H, Lq, Lk, E // H,
               dQp, dKVp[:, :E], dKVp[:, E:])
    if need_w:
        d_in_w = buf(0, (3 * E, E))
        d_in_b = buf(1, (3 * E,))

        def in_wgrads():
            K.linear_wgrad(dQp, st.q_in, dw=d_in_w[:E], db=d_in_b[:E])
            if st.v_in is None:
                K.linear_wgrad(dKVp, st.kv_in, dw=d_in_w[E:], db=d_in_b[E:])
            else:
                K.linear_wgrad(dKVp[:, :E], st.kv_in, dw=d_in_w[E:2 * E], db=d_in_b[E:2 * E])
                K.linear_wgrad(dKVp[:, E:], st.v_in, dw=d_in_w[2 * E:], db=d_in_b[2 * E:])
        on_side(in_wgrads)
    dq = K.linear_dgrad(dQp, st.w_in[:E], residual=dq_acc) if need_dq else None
    if st.v_in is None:
        dkv = K.linear_dgrad(dKVp, st.w_in[E:], residual=dkv_residual) if need_dkv else None
    else:
        dkv = (K.linear_dgrad(dKVp[:, :E], st.w_in[E:2 * E]) if need_dkv else None,
               K.linear_dgrad(dKVp[:, E:], st.w_in[2 * E:]) if need_dkv else None)
    if keep is not None:
        keep.extend((dgate, dA, dQp, dKVp, r0, r1))     # read by the side stream: alive until the join
    return dq, dkv, (d_in_w, d_in_b, d_out_w, d_out_b, d_gate_w, d_gate_b, d_ln_w, d_ln_b)


class CrossBlockFn(Function):
    """One CrossModalTransformer block, query [B,Lq,E] x key [B,Lk,E] (x value [B,Lk,E], None = the key tensor)
    -> [B,Lq,E]."""

    @staticmethod
    def forward(ctx, query, kv, value, in_w, in_b, out_w, out_b, gate_w, gate_b, ln_w, ln_b, num_heads: int):
        B, Lq, E = query.shape
        Lk = kv.shape[1]
        v2d = None if value is None else _c(value).view(B * Lk, E)
        y, st = _block_fwd(_c(query).view(B * Lq, E), _c(kv).view(B * Lk, E), B, Lq, Lk, num_heads,
                           in_w, in_b, out_w, out_b, gate_w, gate_b, ln_w, ln_b, v_in=v2d)
        ctx.st = st
        return y.view(B, Lq, E)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        st = ctx.st
        need_w = any(ctx.needs_input_grad[3:11])
        need_kv = ctx.needs_input_grad[1] or (st.v_in is not None and ctx.needs_input_grad[2])
        dq, dkv, g = _block_bwd(st, K.cast(_c(dy), st.q_in.dtype).view(st.B * st.Lq, st.E),
                                need_dq=ctx.needs_input_grad[0], need_dkv=need_kv, need_w=need_w)
        ctx.st = None
        dq = None if dq is None else dq.view(st.B, st.Lq, st.E)
        dval = None
        if st.v_in is not None:
            dkv, dval = dkv
            dval = None if (dval is None or not ctx.needs_input_grad[2]) else dval.view(st.B, st.Lk, st.E)
            if not ctx.needs_input_grad[1]:
                dkv = None
        dkv = None if dkv is None else dkv.view(st.B, st.Lk, st.E)
        g = tuple(gi if need else None for gi, need in zip(g, ctx.needs_input_grad[3:11]))
        return (dq, dkv, dval) + g + (None,)


def cross_block(query, kv, in_w, in_b, out_w, out_b, gate_w, gate_b, ln_w, ln_b, num_heads: int,
                value: Optional[Tensor] = None) -> Tensor:
    return CrossBlockFn.apply(query, kv, value, in_w, in_b, out_w, out_b, gate_w, gate_b, ln_w, ln_b, num_heads)


class FusionCoreFn(Function):
    """The sequence part of the re-skinned path in ONE autograd node (SURVEY.md section 8(d) composition):
        t = Linear(text), v = Linear(image)                         (Subnetwork.proj pattern, :86)
        t2 = Block_e2p(query=t, kv=v), v2 = Block_p2e(query=v, kv=t) (:287-297, bidirectional)
        returns mean_tokens(t), mean_tokens(v), mean_tokens(t2), mean_tokens(v2)   each fp32 [B,E]
        (+ the first two again in the compute dtype -- operand copies for the tail's first Linear; None in fp32 mode)
    Keeping it one node lets every gradient accumulation on the big [B,L,E] tensors happen inside
    GEMM epilogues instead of autograd's add kernels; t2 / v2 are never written (the LayerNorm kernel
    pools them in fp32 registers) and the pooled-output backward never materialises a [B,L,E] broadcast."""

    @staticmethod
    def forward(ctx, text, image, wt, bt, wi, bi, num_heads, *bp):
        p1, p2 = bp[:8], bp[8:]
        cd = text.dtype
        B, L, Dt = text.shape
        R, Di = image.shape[1], image.shape[2]
        text2d, image2d = _c(text).view(B * L, Dt), _c(image).view(B * R, Di)
        wtc, wic = _w(wt, cd), _w(wi, cd)
        t = K.linear_fwd(text2d, wtc, bt.detach())
        v = K.linear_fwd(image2d, wic, bi.detach())
        # The two blocks only share their inputs: block p2e (image queries: 12 544 rows, GEMMs of two waves each) runs on the
        # helper stream beside block e2p (32 768 rows), so one kernel's partial last wave and launch gap fill with the other's CTAs
        main = torch.cuda.current_stream(text.device)
        side = _side_stream(text.device) if (OVERLAP_BLOCKS and OVERLAP_WGRAD) else None
        keep: list = []
        if side is not None:
            side.wait_stream(main)
            with torch.cuda.stream(side):
                (e2, fv, fv_lp), st2 = _block_fwd(v, t, B, R, L, num_heads, *p2, pooled=True, keep=keep)
        (e1, f0, f0_lp), st1 = _block_fwd(t, v, B, L, R, num_heads, *p1, pooled=True)
        if side is not None:
            main.wait_stream(side)
            keep.clear()
        else:
            (e2, fv, fv_lp), st2 = _block_fwd(v, t, B, R, L, num_heads, *p2, pooled=True)
        ctx.st = (st1, st2, text2d, image2d)
        ctx.dims = (B, L, R)
        ctx.plist = (wt, bt, wi, bi) + tuple(bp)          # the Parameter objects: keys of the gradient sink
        ctx.set_materialize_grads(False)                  # the operand copies never have a gradient: no zero fills
        if f0_lp is not None:
            ctx.mark_non_differentiable(f0_lp, fv_lp)
        return f0, fv, e1, e2, f0_lp, fv_lp

    @staticmethod
    @once_differentiable
    def backward(ctx, df0, dfv, de1, de2, _df0_lp=None, _dfv_lp=None):
        st1, st2, text2d, image2d = ctx.st
        ctx.st = None
        if df0 is None and dfv is None and de1 is None and de2 is None:
            ctx.plist = None
            return (None,) * 23
        zero = lambda: torch.zeros((ctx.dims[0], st1.E), device=text2d.device, dtype=torch.float32)   # an unused output
        df0, dfv, de1, de2 = (zero() if x is None else K.cast(_c(x), torch.float32) for x in (df0, dfv, de1, de2))
        need_w1 = any(ctx.needs_input_grad[7:15])
        need_w2 = any(ctx.needs_input_grad[15:23])
        # data-parallel gradient arena: destinations for the parameter gradients (None: allocate and hand to autograd)
        sink = _GRAD_SINK
        plist = ctx.plist
        ctx.plist = None
        need_p = (ctx.needs_input_grad[2:6]) + tuple(ctx.needs_input_grad[7:23])

        def dst(j):
            return sink.sink(plist[j]) if (sink is not None and need_p[j] and plist[j].is_leaf) else None
        d_proj = [dst(j) for j in range(4)]
        d_b1 = [dst(4 + j) for j in range(8)] if need_w1 else [None] * 8
        d_b2 = [dst(12 + j) for j in range(8)] if need_w2 else [None] * 8
        # block e2p: grad wrt t (as query, + pooled f0 broadcast), grad wrt v (as key/value)
        main = torch.cuda.current_stream(df0.device)
        side = _side_stream(df0.device) if OVERLAP_WGRAD else None
        keep: list = []
        dt1, dv1, g1 = _block_bwd(st1, de1, dpooled_q=df0, need_w=need_w1, wgrad_stream=side, keep=keep, sinks=d_b1)
        if sink is not None and need_w1:
            sink.bucket_done(plist[4:12], side)           # block e2p's gradients are enqueued: their all-reduce may start
        # block p2e: grad wrt v (as query, + pooled fv broadcast + dv1), grad wrt t (as kv, + dt1)
        dv_tot, dt_tot, g2 = _block_bwd(st2, de2, dpooled_q=dfv, dq_add=dv1, dkv_residual=dt1, need_w=need_w2,
                                        wgrad_stream=side, keep=keep, sinks=d_b2)
        if sink is not None and need_w2:
            sink.bucket_done(plist[12:20], side)
        dwt = dbt = dwi = dbi = None
        if ctx.needs_input_grad[2] or ctx.needs_input_grad[3]:
            dwt, dbt = K.linear_wgrad(dt_tot, text2d, dw=d_proj[0], db=d_proj[1])
        if ctx.needs_input_grad[4] or ctx.needs_input_grad[5]:
            dwi, dbi = K.linear_wgrad(dv_tot, image2d, dw=d_proj[2], db=d_proj[3])
        if side is not None:
            main.wait_stream(side)
        keep.clear()
        if sink is not None:
            sink.bucket_done(plist[0:4], main)
        # gradients that went into the sink are not handed to autograd (the sink's step() publishes them as .grad)
        proj_g = [None if d is not None else g for d, g in zip(d_proj, (dwt, dbt, dwi, dbi))]
        g1 = tuple(None if d is not None else g for d, g in zip(d_b1, g1))
        g2 = tuple(None if d is not None else g for d, g in zip(d_b2, g2))
        return (None, None) + tuple(proj_g) + (None,) + g1 + g2


def fusion_core(text, image, wt, bt, wi, bi, num_heads, block1: Sequence[Tensor], block2: Sequence[Tensor]):
    return FusionCoreFn.apply(text, image, wt, bt, wi, bi, num_heads, *block1, *block2)


# ------------------------------------------------------------------------------------------ attention (generic)
class AttnFn(Function):
    """Attention core on projected q,k,v given as [B*L, H*D] (possibly strided column views)."""

    @staticmethod
    def forward(ctx, q, k, v, B, H, Lq, Lk):
        D = q.shape[1] // H
        o, lse = K.attn_fwd(q, k, v, B, H, Lq, Lk, D)
        ctx.save_for_backward(q, k, v, o, lse)
        ctx.dims = (B, H, Lq, Lk, D)
        return o

    @staticmethod
    @once_differentiable
    def backward(ctx, do):
        q, k, v, o, lse = ctx.saved_tensors
        B, H, Lq, Lk, D = ctx.dims
        dq = torch.empty((B * Lq, H * D), device=q.device, dtype=q.dtype)
        dk = torch.empty((B * Lk, H * D), device=q.device, dtype=q.dtype)
        dv = torch.empty((B * Lk, H * D), device=q.device, dtype=q.dtype)
        K.attn_bwd(q, k, v, o, K.cast(_c(do), q.dtype), lse, B, H, Lq, Lk, D, dq, dk, dv)
        return dq, dk, dv, None, None, None, None


class AttnDropFn(Function):
    """Attention core with dropout on the PROBABILITIES (nn.MultiheadAttention(dropout=p) in training mode --
    nn.TransformerEncoderLayer(dropout=0.3), MultimodalModel.py:89-95).  The mask is not stored: the backward re-draws it
    from the same Philox position, a snapshot of the module's device-resident stream state taken at forward time (the
    state itself moves on at the end of the forward pass); parity tests inject an explicit [B,H,Lq,Lk] mask instead."""

    @staticmethod
    def forward(ctx, q, k, v, B, H, Lq, Lk, p: float, drop, name: str):
        D = q.shape[1] // H
        mask, seed, off = drop.next(name, (B, H, Lq, Lk))
        state = None if mask is not None else drop.state(q.device).clone()
        o, lse = K.attn_dropout_fwd(q, k, v, B, H, Lq, Lk, D, p, mask, seed, off, state)
        ctx.save_for_backward(q, k, v, o, lse, mask, state)
        ctx.cfg = (B, H, Lq, Lk, D, p, seed, off)
        return o

    @staticmethod
    @once_differentiable
    def backward(ctx, do):
        q, k, v, o, lse, mask, state = ctx.saved_tensors
        B, H, Lq, Lk, D, p, seed, off = ctx.cfg
        dq = torch.empty((B * Lq, H * D), device=q.device, dtype=q.dtype)
        dk = torch.empty((B * Lk, H * D), device=q.device, dtype=q.dtype)
        dv = torch.empty((B * Lk, H * D), device=q.device, dtype=q.dtype)
        K.attn_dropout_bwd(q, k, v, o, K.cast(_c(do), q.dtype), lse, B, H, Lq, Lk, D, dq, dk, dv, p, mask, seed, off, state)
        return dq, dk, dv, None, None, None, None, None, None, None


def self_attention(x: Tensor, in_w, in_b, out_w, out_b, num_heads: int, dropout_p: float = 0.0,
                   drop: Optional["_DropoutState"] = None, name: str = "self_attn") -> Tensor:
    """nn.MultiheadAttention self-attention on x:[B,S,E] (batch-major), packed in-projection
    (ME-MHACL/model.py:71, MultimodalModel.py:397); dropout_p > 0: dropout on the attention probabilities."""
    B, S, E = x.shape
    qkv = linear(x.reshape(B * S, E), in_w, in_b)
    if dropout_p > 0.0:
        o = AttnDropFn.apply(qkv[:, :E], qkv[:, E:2 * E], qkv[:, 2 * E:], B, num_heads, S, S, dropout_p, drop, name)
    else:
        o = AttnFn.apply(qkv[:, :E], qkv[:, E:2 * E], qkv[:, 2 * E:], B, num_heads, S, S)
    return linear(o, out_w, out_b).view(B, S, E)


# ------------------------------------------------------------------------------------------ pooling
class PoolFn(Function):
    @staticmethod
    def forward(ctx, x, is_max: bool):
        B, L, E = x.shape
        y, arg = K.pool_fwd(_c(x).view(B * L, E), B, L, is_max)
        ctx.dims = (B, L, E, is_max)
        ctx.arg = arg
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        B, L, E, is_max = ctx.dims
        return K.pool_bwd(_c(dy), B, L, is_max, ctx.arg).view(B, L, E), None


def mean_pool(x):
    return PoolFn.apply(x, False)


def max_pool(x):
    return PoolFn.apply(x, True)


class L2NormFn(Function):
    """F.normalize(x, dim=-1) (MultimodalModel.py:388-390)."""

    @staticmethod
    def forward(ctx, x):
        x = _c(x)
        y, norm = K.l2norm_fwd(x)
        ctx.save_for_backward(y, norm)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        y, norm = ctx.saved_tensors
        return K.l2norm_bwd(y, norm, K.cast(_c(dy), torch.float32), None)


def l2_normalize(x):
    return L2NormFn.apply(x)


def stack_tokens(feats: Sequence[Tensor]) -> Tensor:
    """[B,E] x S -> [B,S,E]; a device copy (torch.stack), no arithmetic."""
    return torch.stack(list(feats), dim=1)


# ------------------------------------------------------------------------------------------ modality weights
class ModalConcatFn(Function):
    """softmax over S modality logits + weighted concat (MultimodalModel.py:175,299-306); all fp32."""

    @staticmethod
    def forward(ctx, logits, *slots):
        slots = [K.cast(_c(s), torch.float32) for s in slots]
        w, fused = K.modal_concat_fwd(K.cast(_c(logits), torch.float32), slots, torch.float32)
        ctx.save_for_backward(w, *slots)
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(w)
        return fused, w

    @staticmethod
    @once_differentiable
    def backward(ctx, dfused, _dw):
        w, *slots = ctx.saved_tensors
        if dfused is None:
            return (None,) * (1 + len(slots))
        need = list(ctx.needs_input_grad[1:])
        dslots, dlogits = K.modal_concat_bwd(K.cast(_c(dfused), torch.float32), w, slots, need, torch.float32)
        return (dlogits,) + tuple(dslots)


def modal_concat(logits, slots: Sequence[Tensor]):
    return ModalConcatFn.apply(logits, *slots)


class ModalHeadFn(Function):
    """`attention_weights` (Linear(K,Hd) - GELU - Linear(Hd,S) - Softmax, MultimodalModel.py:171-176) applied to the raw
    features [raw_a | raw_b], and the weighted concat of the S feature slots (:299-306), as ONE autograd node over three
    kernels: the first Linear (GEMM, two A operands, no concat), then GELU + second Linear + softmax + weighting + concat in
    one launch (mmsa_modal_head_fwd), which also writes the fused vector in the compute dtype for the fusion MLP's first
    GEMM.  Backward: one launch for d(slots), d(logits), d(h_pre) (mmsa_modal_head_bwd), then the weight gradients (second
    stream) and the two input gradients of the first Linear."""

    @staticmethod
    def forward(ctx, raw_a, raw_b, a_lp, b_lp, w1, b1, w2, b2, cd, *slots):
        xa = a_lp if (a_lp is not None and a_lp.dtype == cd) else K.cast(K.cast(_c(raw_a), torch.float32), cd)
        xb = None
        if raw_b is not None:
            xb = b_lp if (b_lp is not None and b_lp.dtype == cd) else K.cast(K.cast(_c(raw_b), torch.float32), cd)
        w1c, w2c = _w(w1, cd), _w(w2, cd)
        h_pre = K.linear_fwd(xa, w1c, None if b1 is None else b1.detach(), x2=xb, out_dtype=torch.float32)
        slots32 = [K.cast(_c(s), torch.float32) for s in slots]
        hg, w, fused, fused_lp = K.modal_head_fwd(h_pre, _c(w2c), None if b2 is None else b2.detach(), slots32, cd)
        ctx.save_for_backward(xa, xb, w1c, w2c, h_pre, hg, w, *slots32)
        ctx.cfg = (cd, raw_a.shape, None if raw_b is None else raw_b.shape, b1 is not None, b2 is not None)
        ctx.set_materialize_grads(False)     # fused_lp / w never have a gradient: no zero fills for them
        ctx.mark_non_differentiable(w)
        if fused_lp is not None:
            ctx.mark_non_differentiable(fused_lp)
        return fused, fused_lp, w

    @staticmethod
    @once_differentiable
    def backward(ctx, dfused, _dlp, _dw):
        xa, xb, w1c, w2c, h_pre, hg, w, *slots32 = ctx.saved_tensors
        cd, shp_a, shp_b, has_b1, has_b2 = ctx.cfg
        if dfused is None:
            return (None,) * (9 + len(slots32))
        need_slots = list(ctx.needs_input_grad[9:])
        dslots, dlogits, dh = K.modal_head_bwd(K.cast(_c(dfused), torch.float32), w, slots32, need_slots, h_pre, _c(w2c), cd)
        main = torch.cuda.current_stream(dfused.device)
        side = _side_stream(dfused.device) if OVERLAP_TAIL else None
        Ka = xa.shape[1]
        dw1 = db1 = dw2 = db2 = None
        want_w1, want_b1 = ctx.needs_input_grad[4], has_b1 and ctx.needs_input_grad[5]
        want_w2, want_b2 = ctx.needs_input_grad[6], has_b2 and ctx.needs_input_grad[7]
        if side is not None:
            side.wait_stream(main)
        with torch.cuda.stream(side if side is not None else main):        # leaves of the backward chain
            if want_w2 or want_b2:
                dw2, db2 = K.linear_wgrad(dlogits, hg, want_bias=want_b2, want_weight=want_w2)
            if want_w1 or want_b1:
                dw1 = torch.empty(w1c.shape, device=dh.device, dtype=torch.float32)
                _, db1 = K.linear_wgrad(dh, xa, dw=dw1[:, :Ka], want_bias=want_b1, want_weight=True)
                if xb is not None:
                    K.linear_wgrad(dh, xb, dw=dw1[:, Ka:], want_bias=False, want_weight=True)
                if not want_w1:
                    dw1 = None
        da = db = None
        if ctx.needs_input_grad[0]:
            da = K.linear_dgrad(dh, w1c[:, :Ka], out_dtype=torch.float32).view(shp_a)
        if xb is not None and ctx.needs_input_grad[1]:
            db = K.linear_dgrad(dh, w1c[:, Ka:], out_dtype=torch.float32).view(shp_b)
        if side is not None:          # (deferring this join to the end of the backward pass was tried: no gain, and a tensor
            main.wait_stream(side)    #  kept alive for the helper stream makes AccumulateGrad clone it on the main stream)
        return (da, db, None, None, dw1, db1, dw2, db2, None) + tuple(dslots)


def modal_head(raw_a: Tensor, raw_b: Optional[Tensor], seq: nn.Sequential, slots: Sequence[Tensor], cd: torch.dtype,
               a_lp: Optional[Tensor] = None, b_lp: Optional[Tensor] = None):
    """-> (fused fp32 [B,S*E], fused in the compute dtype or None, softmax weights [B,S]); `seq` must be
    Linear - GELU - Linear - Softmax (the reference's attention_weights)."""
    l1, l2 = seq[0], seq[2]
    return ModalHeadFn.apply(raw_a, raw_b, a_lp, b_lp, l1.weight, l1.bias, l2.weight, l2.bias, cd, *slots)


def is_modal_head(seq: nn.Sequential) -> bool:
    m = list(seq.children())
    return (len(m) == 4 and isinstance(m[0], nn.Linear) and isinstance(m[1], nn.GELU) and isinstance(m[2], nn.Linear)
            and isinstance(m[3], nn.Softmax) and m[2].out_features <= 4 and m[0].out_features <= 256)


# ------------------------------------------------------------------------------------------ nn.Sequential chains
class _DropoutState:
    """Philox stream of the in-kernel dropout of one module; parity tests may inject explicit keep masks.

    The stream POSITION lives in device memory (`state(device)` = int64 {seed, position}): the kernels add it to the
    relative offset they are launched with, and `commit()` -- called by the owning module at the end of every forward --
    enqueues one mmsa_rng_advance that moves the position past what the pass drew.  Both are kernel nodes, so a CUDA
    graph captured around the pass draws a new mask on every replay, from the same stream an eager loop would draw
    (seed/offset passed by value alone would be frozen into the captured nodes)."""

    def __init__(self, seed: int = 0x5EED):
        self.seed = seed
        self.offset = 0                      # draws since the last commit (relative to the device position)
        self.mask_provider = None
        self._state: Dict[torch.device, Tensor] = {}
        self._position0 = 0                  # where a state tensor created later (first use on a device) starts

    def state(self, device) -> Tensor:
        device = torch.device(device)
        st = self._state.get(device)
        if st is None:
            st = torch.tensor([self.seed, self._position0], dtype=torch.int64, device=device)
            self._state[device] = st
        return st

    def reseed(self, seed: int, position: int = 0) -> None:
        self.seed, self.offset, self._position0 = int(seed), 0, int(position)
        for st in self._state.values():
            st.copy_(torch.tensor([self.seed, int(position)], dtype=torch.int64))

    def next(self, name: str, shape) -> Tuple[Optional[Tensor], int, int]:
        mask = self.mask_provider(name, tuple(shape)) if self.mask_provider else None
        off = self.offset
        n = 1
        for s in shape:
            n *= int(s)
        self.offset += n
        return mask, self.seed, off

    def commit(self, device) -> None:
        """end of a forward pass: move the device-resident position past this pass's draws"""
        if self.offset and torch.device(device).type == "cuda":
            K.rng_advance(self.state(device), self.offset)
            self.offset = 0


def _plan(seq: nn.Sequential):
    """nn.Sequential -> list of kernel steps.  Recognised layouts: Linear[-BatchNorm1d-GELU[-Dropout]]
    (MultimodalModel.py:179-225), Linear[-ReLU-BatchNorm1d[-Dropout]] (ME-MHACL/model.py:82-97),
    Linear-ReLU[-Dropout] (:105-109), Linear-GELU (:172-173), Linear-ReLU-BatchNorm1d (MultimodalModel.py:402),
    trailing Linear; a trailing Softmax is left to the caller (fused into modal_concat)."""
    mods = list(seq.children())
    steps, i, n = [], 0, len(mods)
    while i < n:
        m = mods[i]
        nxt = mods[i + 1:i + 4]
        if isinstance(m, nn.Linear):
            steps.append(("linear", m))
            i += 1
        elif isinstance(m, nn.BatchNorm1d):
            if len(nxt) >= 1 and isinstance(nxt[0], nn.GELU):
                if len(nxt) >= 2 and isinstance(nxt[1], nn.Dropout):
                    steps.append(("bn_act", m, BN_THEN_GELU, nxt[1], i + 2)); i += 3
                else:
                    steps.append(("bn_act", m, BN_THEN_GELU, None, -1)); i += 2
            else:
                steps.append(("bn_act", m, BN_ONLY, None, -1)); i += 1
        elif isinstance(m, nn.ReLU) and len(nxt) >= 1 and isinstance(nxt[0], nn.BatchNorm1d):
            if len(nxt) >= 2 and isinstance(nxt[1], nn.Dropout):
                steps.append(("bn_act", nxt[0], RELU_THEN_BN, nxt[1], i + 2)); i += 3
            else:
                steps.append(("bn_act", nxt[0], RELU_THEN_BN, None, -1)); i += 2
        elif isinstance(m, nn.GELU):
            steps.append(("act", ACT_GELU)); i += 1
        elif isinstance(m, nn.ReLU):
            steps.append(("act", ACT_RELU)); i += 1
        elif isinstance(m, nn.Dropout):
            steps.append(("dropout", m, i)); i += 1
        elif isinstance(m, nn.Softmax):
            i += 1
        else:
            raise NotImplementedError(f"mmsa: no kernel mapping for {type(m).__name__}")
    return steps


class SeqFn(Function):
    """One nn.Sequential of the reference's head / fusion layouts as ONE autograd node.
    Input x (and optional second feature segment x2) and output are fp32; inside, GEMM operands are
    `cd`, GEMM outputs fp32, and the elementwise kernel that feeds the next GEMM writes `cd` directly."""

    @staticmethod
    def forward(ctx, x, x2, seq, steps, drop: _DropoutState, name: str, cd, x_lp, x2_lp, want_lp: bool, *params):
        # x_lp / x2_lp: the same values already in the compute dtype (written by the producing kernel): no cast launch;
        # want_lp: also return the output in the compute dtype (from the last kernel of the chain) for the next chain
        training = seq.training
        pi = 0
        tape = []
        cur = K.cast(_c(x), torch.float32)
        a = x_lp if (x_lp is not None and x_lp.dtype == cd) else K.cast(cur, cd)
        a2 = None if x2 is None else (x2_lp if (x2_lp is not None and x2_lp.dtype == cd)
                                      else K.cast(K.cast(_c(x2), torch.float32), cd))
        out_lp = None
        n = len(steps)
        fused_next = False
        for si, st in enumerate(steps):
            if fused_next:                   # this BatchNorm step ran inside the previous Linear's launch
                fused_next = False
                continue
            last = si == n - 1
            nxt_is_linear = (not last) and steps[si + 1][0] == "linear"
            out_dt = cd if nxt_is_linear else torch.float32
            kind = st[0]
            if kind == "linear":
                lin = st[1]
                w, b = params[pi], (params[pi + 1] if lin.bias is not None else None)
                pidx = pi
                pi += 2 if lin.bias is not None else 1
                if a is None:
                    a = K.cast(cur, cd)
                wc = _w(w, cd)
                if (FUSE_LINEAR_BN and not last and steps[si + 1][0] == "bn_act" and a2 is None and K.linear_bn_act_ok(a, wc)):
                    # Linear + BatchNorm(+act+dropout) in ONE launch: a CTA owns all batch rows of its output columns, so the
                    # column statistics never leave it (mmsa_linear_bn_act_fwd); the tape is the two-kernel path's
                    bn, order, dmod, didx = steps[si + 1][1:5]
                    gamma, beta = params[pi], params[pi + 1]
                    bn_pidx = pi
                    pi += 2
                    last2 = si + 1 == n - 1
                    nxt2_is_linear = (not last2) and steps[si + 2][0] == "linear"
                    out_dt2 = cd if nxt2_is_linear else torch.float32
                    p = dmod.p if (dmod is not None and training) else 0.0
                    mask, seed, off = drop.next(f"{name}.{didx}", (a.shape[0], wc.shape[0])) if p > 0 else (None, 0, 0)
                    nbt = bn.num_batches_tracked if (training and bn.track_running_stats) else None
                    lp_here = want_lp and last2 and out_dt2 == torch.float32 and cd == torch.bfloat16
                    res = K.linear_bn_act_fwd(a, wc, None if b is None else b.detach(), gamma.detach(), beta.detach(),
                                              bn.running_mean, bn.running_var, 0.1 if bn.momentum is None else bn.momentum,
                                              bn.eps, training, order, p, mask, seed, off, out_dt2,
                                              rng_state=(drop.state(a.device) if (p > 0 and mask is None) else None),
                                              want_lp=lp_here, num_batches_tracked=nbt)
                    z, y, mean, rstd, mask = res[:5]
                    if lp_here:
                        out_lp = res[5]
                    tape.append(("linear", a, None, wc, pidx, lin.bias is not None))
                    tape.append(("bn_act", z, gamma.detach(), beta.detach(), mean, rstd, training, order, p, mask, bn_pidx))
                    cur, a, a2 = (None, y, None) if nxt2_is_linear else (y, None, None)
                    fused_next = True
                    continue
                z = K.linear_fwd(a, wc, None if b is None else b.detach(), x2=a2, out_dtype=torch.float32)
                tape.append(("linear", a, a2, wc, pidx, lin.bias is not None))
                cur, a, a2 = z, (K.cast(z, cd) if nxt_is_linear else None), None
            elif kind == "bn_act":
                bn, order, dmod, didx = st[1], st[2], st[3], st[4]
                gamma, beta = params[pi], params[pi + 1]
                pidx = pi
                pi += 2
                p = dmod.p if (dmod is not None and training) else 0.0
                mask, seed, off = drop.next(f"{name}.{didx}", cur.shape) if p > 0 else (None, 0, 0)
                nbt = bn.num_batches_tracked if (training and bn.track_running_stats) else None   # the kernel counts
                lp_here = want_lp and last and out_dt == torch.float32 and cd == torch.bfloat16
                res = K.bn_act_fwd(cur, gamma.detach(), beta.detach(), bn.running_mean, bn.running_var,
                                   0.1 if bn.momentum is None else bn.momentum, bn.eps, training, order,
                                   p, mask, seed, off, out_dt,
                                   rng_state=(drop.state(cur.device) if (p > 0 and mask is None) else None), want_lp=lp_here,
                                   num_batches_tracked=nbt)
                y, mean, rstd, mask = res[:4]
                if lp_here:
                    out_lp = res[4]
                tape.append(("bn_act", cur, gamma.detach(), beta.detach(), mean, rstd, training, order, p, mask, pidx))
                cur, a = (None, y) if nxt_is_linear else (y, None)
            elif kind == "act":
                y = K.act_fwd(cur, st[1], out_dt)
                tape.append(("act", cur, st[1]))
                cur, a = (None, y) if nxt_is_linear else (y, None)
            elif kind == "dropout":
                dmod, didx = st[1], st[2]
                if training and dmod.p > 0:
                    mask, seed, off = drop.next(f"{name}.{didx}", cur.shape)
                    y, mask = K.dropout(cur, dmod.p, mask, mask is not None, seed, off, out_dt,
                                        rng_state=(drop.state(cur.device) if mask is None else None))
                    tape.append(("dropout", dmod.p, mask))
                    cur, a = (None, y) if nxt_is_linear else (y, None)
                elif nxt_is_linear:
                    a, cur = K.cast(cur, cd), None
        ctx.tape = tape
        ctx.cd = cd
        ctx.set_materialize_grads(False)     # the compute-dtype copy never has a gradient: no zero fill for it
        ctx.n_params = len(params)
        ctx.has_x2 = x2 is not None
        ctx.in_dtypes = (x.dtype, None if x2 is None else x2.dtype)
        ctx.in_shapes = (x.shape, None if x2 is None else x2.shape)
        if want_lp:
            if out_lp is None and cd != torch.float32:
                out_lp = K.cast(cur, cd)
            if out_lp is not None:
                ctx.mark_non_differentiable(out_lp)
            return cur, out_lp
        return cur

    @staticmethod
    @once_differentiable
    def backward(ctx, dy, _dlp=None):
        tape, cd = ctx.tape, ctx.cd
        ctx.tape = None
        if dy is None:
            return (None,) * (10 + ctx.n_params)
        grads: List[Optional[Tensor]] = [None] * ctx.n_params
        need = ctx.needs_input_grad[10:]
        d = K.cast(_c(dy), torch.float32)          # fp32 unless an elementwise backward wrote cd for a GEMM
        main = torch.cuda.current_stream(dy.device)
        side = _side_stream(dy.device)
        keep: List[Optional[Tensor]] = []
        used_side = False
        dx = dx2 = None
        db_fp32 = None                             # bias gradient handed down by a BatchNorm backward
        for ti in range(len(tape) - 1, -1, -1):
            op = tape[ti]
            prev_is_linear = ti > 0 and tape[ti - 1][0] == "linear"
            out_dt = cd if prev_is_linear else torch.float32
            kind = op[0]
            if kind == "linear":
                _, a, a2, wc, pidx, has_b = op
                d_cd = K.cast(d, cd)
                Kx = a.shape[1]
                want_w = need[pidx]
                want_b = has_b and need[pidx + 1]
                if want_b and db_fp32 is not None:     # column sums taken on fp32 values by bn_act_bwd
                    grads[pidx + 1] = db_fp32
                    want_b = False
                db_fp32 = None
                if want_w or want_b:
                    # weight gradients are leaves of the backward chain: they go to a second stream and overlap the
                    # dgrad / BatchNorm chain (all of it small latency-bound kernels); joined before returning
                    wst = side if OVERLAP_TAIL else main
                    if wst is side:
                        side.wait_stream(main)
                    with torch.cuda.stream(wst):
                        if a2 is None:
                            dw, db = K.linear_wgrad(d_cd, a, want_bias=want_b, want_weight=want_w)
                        else:
                            dw = torch.empty(wc.shape, device=d.device, dtype=torch.float32)
                            _, db = K.linear_wgrad(d_cd, a, dw=dw[:, :Kx], want_bias=want_b, want_weight=True)
                            K.linear_wgrad(d_cd, a2, dw=dw[:, Kx:], want_bias=False, want_weight=True)
                    keep.extend((d_cd, a, a2, dw, db))     # alive until the join: no reuse of their memory by `main` meanwhile
                    used_side = used_side or wst is side
                    grads[pidx] = dw if want_w else None
                    if has_b and want_b:
                        grads[pidx + 1] = db
                earlier_needs = ti > 0 and any(need[:pidx])
                if ti == 0:
                    if ctx.needs_input_grad[0]:
                        dx = K.linear_dgrad(d_cd, wc[:, :Kx], out_dtype=torch.float32)
                    if a2 is not None and ctx.needs_input_grad[1]:
                        dx2 = K.linear_dgrad(d_cd, wc[:, Kx:], out_dtype=torch.float32)
                elif earlier_needs or ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
                    d = K.linear_dgrad(d_cd, wc, out_dtype=torch.float32)
                else:
                    break
            elif kind == "bn_act":
                _, xz, gamma, beta, mean, rstd, training, order, p, mask, pidx = op
                d, dgamma, dbeta, dbp = K.bn_act_bwd(xz, K.cast(d, torch.float32), gamma, beta, mean, rstd, training,
                                                     order, p, mask, out_dt)
                db_fp32 = dbp if prev_is_linear else None
                grads[pidx] = dgamma if need[pidx] else None
                grads[pidx + 1] = dbeta if need[pidx + 1] else None
            elif kind == "act":
                d = K.act_bwd(op[1], K.cast(d, torch.float32), op[2], out_dt)
            elif kind == "dropout":
                d, _ = K.dropout(K.cast(d, torch.float32), op[1], op[2], True, 0, 0, out_dt)
        if used_side:
            main.wait_stream(side)
            keep.clear()
        if tape and tape[0][0] != "linear" and ctx.needs_input_grad[0]:
            dx = K.cast(d, torch.float32)
        if dx is not None:
            dx = K.cast(dx, ctx.in_dtypes[0]).view(ctx.in_shapes[0])
        if dx2 is not None:
            dx2 = K.cast(dx2, ctx.in_dtypes[1]).view(ctx.in_shapes[1])
        return (dx, dx2, None, None, None, None, None, None, None, None) + tuple(grads)


def sequential(x: Tensor, seq: nn.Sequential, drop: _DropoutState, name: str, cd: torch.dtype,
               x2: Optional[Tensor] = None, x_lp: Optional[Tensor] = None, x2_lp: Optional[Tensor] = None,
               want_lp: bool = False):
    """Run an nn.Sequential (parameter container) on the CUDA kernels; fp32 in, fp32 out.
    x_lp / x2_lp: optional copies of x / x2 in the compute dtype (saves the cast launch); want_lp: return
    (out, out in the compute dtype or None) -- the copy comes out of the chain's last kernel."""
    steps = getattr(seq, "_mmsa_plan", None)
    if steps is None:
        steps = _plan(seq)
        object.__setattr__(seq, "_mmsa_plan", steps)
    params: List[Tensor] = []
    for st in steps:
        if st[0] == "linear":
            params.append(st[1].weight)
            if st[1].bias is not None:
                params.append(st[1].bias)
        elif st[0] == "bn_act":
            params.extend([st[1].weight, st[1].bias])
    return SeqFn.apply(x, x2, seq, steps, drop, name, cd, x_lp, x2_lp, want_lp, *params)


# ------------------------------------------------------------------------------------------ losses
class CrossEntropyFn(Function):
    """nn.CrossEntropyLoss() (mean) on fp32 logits (Trainer.py:17,68)."""

    @staticmethod
    def forward(ctx, logits, labels, extra):
        logits = _c(logits.float())
        loss, pred = K.ce_fwd(logits, labels, None if extra is None else _c(extra.float()).view(-1))
        ctx.save_for_backward(logits, labels)
        ctx.pred = pred
        ctx.extra_shape = None if extra is None else extra.shape
        return loss.view(())

    @staticmethod
    @once_differentiable
    def backward(ctx, dloss):
        logits, labels = ctx.saved_tensors
        dloss = _c(dloss.float()).view(1)
        dextra = None
        if ctx.extra_shape is not None and ctx.needs_input_grad[2]:
            dextra = dloss.expand(ctx.extra_shape) if len(ctx.extra_shape) else dloss.view(())
        return K.ce_bwd(logits, labels, dloss), None, dextra


def cross_entropy(logits, labels, extra: Optional[Tensor] = None):
    """mean softmax cross-entropy; extra (a few fp32 device values, e.g. the weighted contrastive loss of Trainer.py:68-71):
    returns CE + extra.sum() from the same launch (d extra = d loss)."""
    return CrossEntropyFn.apply(logits, labels, extra)


class ContrastiveFn(Function):
    """normalize -> cosine block f1n f2n^T -> row reductions (InfoNCE / SupCon / NT-Xent).
    f1:[B,E] local rows, f2:[Bg,E] columns (the gathered global batch under data parallelism).
    The small [B,E] embeddings are processed in fp32 whatever the storage mode: with T=0.01 a bf16
    cosine (4e-3) would move the scaled logits by 0.4.  `fast` (bf16 performance mode) runs the three
    GEMMs on the tensor cores with split-bf16 operands (hi.hi + hi.lo + lo.hi over a 3x longer reduction,
    ~2^-16 relative, fp32 accumulation); otherwise they run on the exact-fp32 CUDA-core GEMM (1e-5 parity mode)."""

    @staticmethod
    def forward(ctx, f1, f2, labels_r, labels_c, temperature, temperature_const, kind, row_offset, denom, same, fast,
                weight):
        # weight (a learnable fp32 scalar or None): returns weight * loss, shape (1,) -- `self.contrastive_weight * loss`
        # (MultimodalModel.py:315-317) from the loss kernels themselves instead of a multiply and its three backward launches
        a = K.cast(_c(f1.detach()), torch.float32)
        n1, norm1 = K.l2norm_fwd(a)
        if same:
            n2, norm2 = n1, norm1
        else:
            n2, norm2 = K.l2norm_fwd(K.cast(_c(f2.detach()), torch.float32))
        n1_row = n2_row = None
        if fast:
            n1_col, n1_row = K.split3(n1, col_side="a", row_side="b")       # A of the forward, B of dn2 = G^T n1
            n2_col, n2_row = K.split3(n2, col_side="b", row_side="b")       # B of the forward, B of dn1 = G n2
            sim = K.linear_fwd(n1_col, n2_col, None, out_dtype=torch.float32)
        else:
            sim = K.linear_fwd(n1, n2, None, out_dtype=torch.float32)
        tptr = None if temperature is None else temperature.detach().float().view(1)
        wptr = None if weight is None else weight.detach().float().view(1)
        loss, stats, raw = K.contrastive_fwd(kind, sim, labels_r, labels_c, tptr, temperature_const, row_offset, denom, wptr)
        ctx.save_for_backward(n1, norm1, n2, norm2, sim, stats, labels_r, labels_c, tptr, n1_row, n2_row, wptr, raw)
        ctx.cfg = (temperature_const, kind, row_offset, denom, same, f1.dtype, f2.dtype, fast)
        ctx.wshape = None if weight is None else weight.shape
        return loss.view(()) if weight is None else loss.view(weight.shape)

    @staticmethod
    @once_differentiable
    def backward(ctx, dloss):
        n1, norm1, n2, norm2, sim, stats, labels_r, labels_c, tptr, n1_row, n2_row, wptr, raw = ctx.saved_tensors
        tconst, kind, row_offset, denom, same, d1, d2, fast = ctx.cfg
        G, dtemp, dweight = K.contrastive_bwd(kind, sim, labels_r, labels_c, tptr, tconst, row_offset, denom, stats,
                                              _c(dloss.float()).view(1), torch.float32, weight=wptr, loss_raw=raw,
                                              want_dweight=wptr is not None and ctx.needs_input_grad[11])
        if fast:
            g_col, g_row = K.split3(G, col_side="a", row_side="a")
            dn1 = K.linear_dgrad(g_col, n2_row, out_dtype=torch.float32)
            dn2, _ = K.linear_wgrad(g_row, n1_row, want_bias=False)
        else:
            dn1 = K.linear_dgrad(G, n2)
            dn2, _ = K.linear_wgrad(G, n1, want_bias=False)
        if same:
            df1 = K.cast(K.l2norm_bwd(n1, norm1, dn1, dn2), d1)
            df2 = None
        else:
            df1 = K.cast(K.l2norm_bwd(n1, norm1, dn1, None), d1)
            df2 = K.cast(K.l2norm_bwd(n2, norm2, dn2, None), d2) if ctx.needs_input_grad[1] else None
        dT = dtemp.view(()) if (tptr is not None and ctx.needs_input_grad[4]) else None
        if dweight is not None:
            dweight = dweight.view(ctx.wshape)
        return df1, df2, None, None, dT, None, None, None, None, None, None, dweight


def infonce(f1: Tensor, f2: Tensor, labels: Tensor, temperature, labels_cols: Optional[Tensor] = None,
            row_offset: int = 0, fast: bool = False, weight: Optional[Tensor] = None) -> Tensor:
    """MultimodalTransformerModel.compute_contrastive_loss (MultimodalModel.py:232-260).
    `labels_cols`/`row_offset` describe a row block of a batch-sharded similarity matrix.
    weight (one-element fp32 tensor, e.g. the model's learnable contrastive_weight): returns weight * loss with the
    weight's shape (MultimodalModel.py:315-317), computed and differentiated inside the loss kernels."""
    same = f1 is f2
    lc = labels if labels_cols is None else labels_cols
    t_tensor = temperature if isinstance(temperature, Tensor) else None
    t_const = 0.0 if t_tensor is not None else float(temperature)
    return ContrastiveFn.apply(f1, f2, labels, lc, t_tensor, t_const, LOSS_INFONCE, row_offset, f1.shape[0], same, fast,
                               weight)


def supcon(z1: Tensor, z2: Tensor, labels: Tensor, temperature: float = 0.1) -> Tensor:
    """train.py:16-40 contrastive_loss: SupCon over the 2B stacked views."""
    z = _StackFn.apply(z1, z2)
    lab = torch.cat([labels.view(-1), labels.view(-1)])
    return ContrastiveFn.apply(z, z, lab, lab, None, float(temperature), LOSS_SUPCON, 0, z.shape[0], True, False, None)


def ntxent(z1: Tensor, z2: Tensor, temperature: float = 0.5) -> Tensor:
    """ME-MHACL/train.py:47-66 contrastive_loss: NT-Xent over the 2N stacked views."""
    z = _StackFn.apply(z1, z2)
    return ContrastiveFn.apply(z, z, None, None, None, float(temperature), LOSS_NTXENT, 0, z.shape[0], True, False, None)


class _StackFn(Function):
    """cat([z1, z2], 0) as two device-to-device copies (plumbing; no arithmetic)."""

    @staticmethod
    def forward(ctx, z1, z2):
        n = z1.shape[0]
        z = torch.empty((2 * n, z1.shape[1]), device=z1.device, dtype=z1.dtype)
        z[:n].copy_(z1)
        z[n:].copy_(z2)
        ctx.n = n
        return z

    @staticmethod
    def backward(ctx, dz):
        return dz[:ctx.n], dz[ctx.n:]


# ------------------------------------------------------------------------------------------ encoder tail (section 8(f) rank 2)
class AddLNFn(Function):
    """y = LayerNorm(x + r) (post-norm residual of nn.TransformerEncoderLayer; r=None: plain LayerNorm)."""

    @staticmethod
    def forward(ctx, x, r, gamma, beta, eps: float):
        x2 = _c(x).view(-1, x.shape[-1])
        r2 = None if r is None else _c(r).view(-1, x.shape[-1])
        y, mean, rstd = K.add_ln_fwd(x2, r2, gamma.detach(), beta.detach(), eps)
        ctx.save_for_backward(x2, r2, gamma.detach(), mean, rstd)
        ctx.shape = x.shape
        return y.view(x.shape)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x2, r2, gamma, mean, rstd = ctx.saved_tensors
        du, dgamma, dbeta = K.add_ln_bwd(K.cast(_c(dy), x2.dtype).view(-1, x2.shape[1]), x2, r2, gamma, mean, rstd)
        du = du.view(ctx.shape)
        return du, (du if r2 is not None else None), dgamma, dbeta, None


def add_layer_norm(x: Tensor, r: Optional[Tensor], ln: nn.LayerNorm) -> Tensor:
    return AddLNFn.apply(x, r, ln.weight, ln.bias, ln.eps)


class AddRowsFn(Function):
    """x + pe[:, :L] broadcast over the batch (PositionalEncoding.forward, MultimodalModel.py:19-20)."""

    @staticmethod
    def forward(ctx, x, pe, L: int):
        return K.add_rows(_c(x).view(-1, x.shape[-1]), pe, L).view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        return dy, None, None


class ActDropFn(Function):
    """activation followed by dropout on a token stream: fp32 pre-activation in, compute dtype out
    (the `dropout(activation(linear1(x)))` of nn.TransformerEncoderLayer's feed-forward block)."""

    @staticmethod
    def forward(ctx, z, act: int, p: float, training: bool, drop: _DropoutState, name: str, cd):
        z = _c(z)
        p_eff = p if training else 0.0
        if p_eff > 0:
            h = K.act_fwd(z, act, torch.float32)
            mask, seed, off = drop.next(name, z.shape)
            y, mask = K.dropout(h, p_eff, mask, mask is not None, seed, off, cd,
                                rng_state=(drop.state(h.device) if mask is None else None))
        else:
            y, mask = K.act_fwd(z, act, cd), None
        ctx.save_for_backward(z, mask)
        ctx.cfg = (act, p_eff, cd)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        z, mask = ctx.saved_tensors
        act, p_eff, cd = ctx.cfg
        d = K.cast(_c(dy), torch.float32)
        if p_eff > 0:
            d, _ = K.dropout(d, p_eff, mask, True, 0, 0, torch.float32)
        return K.act_bwd(z, d, act, torch.float32), None, None, None, None, None, None


class DropFn(Function):
    """stand-alone dropout on a token stream (dropout1 / dropout2 of nn.TransformerEncoderLayer)."""

    @staticmethod
    def forward(ctx, x, p: float, drop: _DropoutState, name: str):
        mask, seed, off = drop.next(name, x.shape)
        y, mask = K.dropout(K.cast(_c(x), torch.float32), p, mask, mask is not None, seed, off, x.dtype,
                            rng_state=(drop.state(x.device) if mask is None else None))
        ctx.save_for_backward(mask)
        ctx.p = p
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        (mask,) = ctx.saved_tensors
        d, _ = K.dropout(K.cast(_c(dy), torch.float32), ctx.p, mask, True, 0, 0, dy.dtype)
        return d, None, None, None


def token_dropout(x: Tensor, p: float, training: bool, drop: _DropoutState, name: str) -> Tensor:
    return DropFn.apply(x, p, drop, name) if (training and p > 0) else x


def encoder_layer(x: Tensor, layer: nn.TransformerEncoderLayer, drop: _DropoutState, name: str, cd) -> Tensor:
    """One post-norm nn.TransformerEncoderLayer (norm_first=False, ReLU) on x:[B,S,E] in the compute dtype:
        x = norm1(x + dropout1(self_attn(x)));  x = norm2(x + dropout2(linear2(dropout(relu(linear1(x))))))
    In training mode the self-attention also drops attention PROBABILITIES (MultiheadAttention(dropout=0.3)), as torch's
    scaled_dot_product_attention(dropout_p=...) does inside TransformerEncoderLayer._sa_block."""
    a = layer.self_attn
    training = layer.training
    B, S, E = x.shape
    sa = self_attention(x, a.in_proj_weight, a.in_proj_bias, a.out_proj.weight, a.out_proj.bias, a.num_heads,
                        dropout_p=(float(a.dropout) if training else 0.0), drop=drop, name=name + ".self_attn.dropout")
    sa = token_dropout(sa, layer.dropout1.p, training, drop, name + ".dropout1")
    x = add_layer_norm(x, sa, layer.norm1)
    z = linear(x.reshape(B * S, E), layer.linear1.weight, layer.linear1.bias, out_fp32=True, cd=cd)
    h = ActDropFn.apply(z, ACT_RELU, layer.dropout.p, training, drop, name + ".dropout", cd)
    ff = linear(h, layer.linear2.weight, layer.linear2.bias, cd=cd).view(B, S, E)
    ff = token_dropout(ff, layer.dropout2.p, training, drop, name + ".dropout2")
    return add_layer_norm(x, ff, layer.norm2)
