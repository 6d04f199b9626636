"""Thin tensor-level wrappers over the C ABI: each function takes torch CUDA tensors, allocates the
outputs (caller-allocates convention of include/mmsa.h) and enqueues the kernels on the current
stream.  No arithmetic happens in Python/PyTorch here."""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import ctypes
import torch

from . import _lib
from ._lib import F32, BF16, call

Tensor = torch.Tensor


def dt(t_or_dtype) -> int:
    d = t_or_dtype.dtype if isinstance(t_or_dtype, Tensor) else t_or_dtype
    if d == torch.float32:
        return F32
    if d == torch.bfloat16:
        return BF16
    raise TypeError(f"mmsa: unsupported dtype {d}")


import threading

_TLS = threading.local()      # .dev: device of the tensors of the call being marshalled on THIS thread (set by _check)


def _stream() -> int:
    """current stream of the device the call's tensors live on (kernels launch on the CURRENT device, which _check has
    verified to be that device).  Thread-local: nn.DataParallel drives one replica per thread."""
    return torch.cuda.current_stream(getattr(_TLS, "dev", None)).cuda_stream


def _p(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _check(*ts: Optional[Tensor]):
    dev = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise _lib.MmsaError("mmsa: tensors must live on a CUDA device (no CPU fallback)")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise _lib.MmsaError(f"mmsa: operands on different devices ({dev} and {t.device})")
    if dev is not None:
        if dev.index != torch.cuda.current_device():
            raise _lib.MmsaError(f"mmsa: tensors live on {dev} but the current CUDA device is "
                                 f"cuda:{torch.cuda.current_device()}; wrap the call in torch.cuda.device({dev.index})")
        _TLS.dev = dev


def cast(x: Tensor, dtype: torch.dtype) -> Tensor:
    """dtype conversion through mmsa_cast (fp32 <-> bf16)."""
    if x.dtype == dtype:
        return x
    _check(x)
    x = x.contiguous()
    y = torch.empty_like(x, dtype=dtype)
    call("mmsa_cast", x.data_ptr(), dt(x), y.data_ptr(), dt(dtype), x.numel(), _stream())
    return y


def split3(x: Tensor, col_side: Optional[str] = None, row_side: Optional[str] = None):
    """fp32 [R,C] -> split-bf16 operands (mmsa_split3): col-concatenated [R,3C] and/or row-concatenated [3R,C];
    side "a" = (hi,hi,lo), side "b" = (hi,lo,hi).  Returns (col or None, row or None)."""
    _check(x)
    assert x.dtype == torch.float32 and x.ndim == 2 and x.stride(1) == 1
    R, C = x.shape
    col = torch.empty((R, 3 * C), device=x.device, dtype=torch.bfloat16) if col_side else None
    row = torch.empty((3 * R, C), device=x.device, dtype=torch.bfloat16) if row_side else None
    call("mmsa_split3", x.data_ptr(), R, C, x.stride(0), _p(col), 1 if col_side == "b" else 0, _p(row),
         1 if row_side == "b" else 0, _stream())
    return col, row


def linear_fwd(x: Tensor, w: Tensor, bias: Optional[Tensor], *, x2: Optional[Tensor] = None,
               residual: Optional[Tensor] = None, act: int = 0, out_dtype: Optional[torch.dtype] = None,
               out: Optional[Tensor] = None) -> Tensor:
    """y[M,N] = act(cat[x, x2] @ w.T + bias + residual); x:[M,K] (row stride x.stride(0)), w:[N,K(+K2)]."""
    _check(x, w, bias, x2, residual)
    M, K = x.shape
    N = w.shape[0]
    K2 = 0 if x2 is None else x2.shape[1]
    assert w.shape[1] == K + K2 and x.stride(1) == 1 and w.stride(1) == 1
    od = out_dtype or x.dtype
    y = out if out is not None else torch.empty((M, N), device=x.device, dtype=od)
    call("mmsa_linear_fwd", dt(x), M, N, K, K2, x.data_ptr(), x.stride(0), _p(x2), 0 if x2 is None else x2.stride(0),
         w.data_ptr(), w.stride(0), _p(bias), _p(residual), 0 if residual is None else residual.stride(0), act,
         y.data_ptr(), y.stride(0), dt(od), _stream())
    return y


def linear_dgrad(dy: Tensor, w: Tensor, *, residual: Optional[Tensor] = None,
                 out_dtype: Optional[torch.dtype] = None, out: Optional[Tensor] = None) -> Tensor:
    """dx[M,K] = dy[M,N] @ w[N,K] (+ residual)."""
    _check(dy, w, residual)
    M, N = dy.shape
    K = w.shape[1]
    assert w.shape[0] == N and dy.stride(1) == 1 and w.stride(1) == 1
    od = out_dtype or dy.dtype
    dx = out if out is not None else torch.empty((M, K), device=dy.device, dtype=od)
    call("mmsa_linear_dgrad", dt(dy), M, N, K, dy.data_ptr(), dy.stride(0), w.data_ptr(), w.stride(0),
         _p(residual), 0 if residual is None else residual.stride(0), dx.data_ptr(), dx.stride(0), dt(od), _stream())
    return dx


def linear_wgrad(dy: Tensor, x: Tensor, *, dw: Optional[Tensor] = None, db: Optional[Tensor] = None,
                 want_bias: bool = True, want_weight: bool = True) -> Tuple[Optional[Tensor], Optional[Tensor]]:
    """dw[N,K] = dy[M,N].T @ x[M,K] (fp32); db[N] = dy.sum(0) (fp32).  `dw` may be a strided view
    (a column block of a wider weight gradient)."""
    _check(dy, x, dw)
    M, N = dy.shape
    K = x.shape[1]
    assert x.shape[0] == M and dy.stride(1) == 1 and x.stride(1) == 1
    if want_weight and dw is None:
        dw = torch.empty((N, K), device=dy.device, dtype=torch.float32)
    if want_bias and db is None:
        db = torch.empty((N,), device=dy.device, dtype=torch.float32)
    if not want_bias:
        db = None
    if M == 0:
        if dw is not None:
            dw.zero_()
        if db is not None:
            db.zero_()
        return dw, db
    nbytes = _lib.load().mmsa_linear_wgrad_workspace(dt(dy), M, N, K)
    ws = torch.empty((nbytes // 4,), device=dy.device, dtype=torch.float32)
    call("mmsa_linear_wgrad", dt(dy), M, N, K, dy.data_ptr(), dy.stride(0), x.data_ptr(), x.stride(0),
         _p(dw) if want_weight else None, 0 if dw is None else dw.stride(0), _p(db), ws.data_ptr(), _stream())
    return (dw if want_weight else None), db


def linear_wgrad2(dy: Tensor, x: Tensor, x2: Tensor, *, dw: Optional[Tensor] = None, db: Optional[Tensor] = None,
                  want_bias: bool = True) -> Tuple[Tensor, Optional[Tensor]]:
    """dw[N, K+K2] = dy.T @ cat[x, x2] (fp32) without the concat; db[N] = dy.sum(0) (mmsa_linear_wgrad2)."""
    _check(dy, x, x2, dw, db)
    M, N = dy.shape
    K1, K2 = x.shape[1], x2.shape[1]
    assert x.shape[0] == M and x2.shape[0] == M and dy.stride(1) == 1 and x.stride(1) == 1 and x2.stride(1) == 1
    if dw is None:
        dw = torch.empty((N, K1 + K2), device=dy.device, dtype=torch.float32)
    if want_bias and db is None:
        db = torch.empty((N,), device=dy.device, dtype=torch.float32)
    if not want_bias:
        db = None
    nbytes = _lib.load().mmsa_linear_wgrad_workspace(dt(dy), M, N, K1 + K2)
    ws = torch.empty((nbytes // 4,), device=dy.device, dtype=torch.float32)
    call("mmsa_linear_wgrad2", dt(dy), M, N, K1, K2, dy.data_ptr(), dy.stride(0), x.data_ptr(), x.stride(0), x2.data_ptr(),
         x2.stride(0), dw.data_ptr(), dw.stride(0), _p(db), ws.data_ptr(), _stream())
    return dw, db


def attn_fwd(q: Tensor, k: Tensor, v: Tensor, B: int, H: int, Lq: int, Lk: int, D: int) -> Tuple[Tensor, Tensor]:
    """q:[B*Lq, >=H*D] views, k/v:[B*Lk, ...] views (row strides may differ) -> o:[B*Lq,H*D], lse:[B,H,Lq]."""
    _check(q, k, v)
    o = torch.empty((B * Lq, H * D), device=q.device, dtype=q.dtype)
    lse = torch.empty((B, H, Lq), device=q.device, dtype=torch.float32)
    call("mmsa_attn_fwd", dt(q), B, H, Lq, Lk, D, q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0),
         v.data_ptr(), v.stride(0), o.data_ptr(), o.stride(0), lse.data_ptr(), _stream())
    return o, lse


def attn_bwd(q, k, v, o, dout, lse, B, H, Lq, Lk, D, dq: Tensor, dk: Tensor, dv: Tensor):
    _check(q, k, v, o, dout, lse, dq, dk, dv)
    delta = torch.empty((B, H, Lq), device=q.device, dtype=torch.float32)
    call("mmsa_attn_bwd", dt(q), B, H, Lq, Lk, D, q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0),
         v.data_ptr(), v.stride(0), o.data_ptr(), o.stride(0), dout.data_ptr(), dout.stride(0), lse.data_ptr(),
         delta.data_ptr(), dq.data_ptr(), dq.stride(0), dk.data_ptr(), dk.stride(0), dv.data_ptr(), dv.stride(0),
         _stream())


def attn_dropout_fwd(q: Tensor, k: Tensor, v: Tensor, B: int, H: int, Lq: int, Lk: int, D: int, p: float,
                     keep_mask: Optional[Tensor], seed: int, offset: int, rng_state: Optional[Tensor]):
    """attention core with dropout on the probabilities (mmsa_attn_dropout_fwd); keep_mask uint8 [B,H,Lq,Lk] or None (Philox)."""
    _check(q, k, v, keep_mask, rng_state)
    o = torch.empty((B * Lq, H * D), device=q.device, dtype=q.dtype)
    lse = torch.empty((B, H, Lq), device=q.device, dtype=torch.float32)
    call("mmsa_attn_dropout_fwd", dt(q), B, H, Lq, Lk, D, q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0),
         v.data_ptr(), v.stride(0), o.data_ptr(), o.stride(0), lse.data_ptr(), float(p), _p(keep_mask), seed, offset,
         _p(rng_state), _stream())
    return o, lse


def attn_dropout_bwd(q, k, v, o, dout, lse, B, H, Lq, Lk, D, dq: Tensor, dk: Tensor, dv: Tensor, p: float,
                     keep_mask: Optional[Tensor], seed: int, offset: int, rng_state: Optional[Tensor]):
    _check(q, k, v, o, dout, lse, dq, dk, dv, keep_mask, rng_state)
    delta = torch.empty((B, H, Lq), device=q.device, dtype=torch.float32)
    call("mmsa_attn_dropout_bwd", dt(q), B, H, Lq, Lk, D, q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0),
         v.data_ptr(), v.stride(0), o.data_ptr(), o.stride(0), dout.data_ptr(), dout.stride(0), lse.data_ptr(),
         delta.data_ptr(), dq.data_ptr(), dq.stride(0), dk.data_ptr(), dk.stride(0), dv.data_ptr(), dv.stride(0),
         float(p), _p(keep_mask), seed, offset, _p(rng_state), _stream())


def gate_ln_fwd(gate_pre: Tensor, q: Tensor, attn: Tensor, gamma: Tensor, beta: Tensor, eps: float,
                want_y: bool = True):
    _check(gate_pre, q, attn, gamma, beta)
    M, E = q.shape
    g = torch.empty_like(q)
    y = torch.empty_like(q) if want_y else None
    mean = torch.empty((M,), device=q.device, dtype=torch.float32)
    rstd = torch.empty((M,), device=q.device, dtype=torch.float32)
    call("mmsa_gate_ln_fwd", dt(q), M, E, gate_pre.data_ptr(), q.data_ptr(), attn.data_ptr(), gamma.data_ptr(),
         beta.data_ptr(), float(eps), g.data_ptr(), _p(y), mean.data_ptr(), rstd.data_ptr(), _stream())
    return g, y, mean, rstd


def gate_ln_bwd(dy: Tensor, rows_per_sample: int, g, q, attn, gamma, mean, rstd, *,
                dq_bcast: Optional[Tensor] = None, bcast_rows: int = 0, dq_add: Optional[Tensor] = None,
                dgamma: Optional[Tensor] = None, dbeta: Optional[Tensor] = None, packed_parts: bool = False):
    """packed_parts: dq_part and dattn_part are returned as the two column halves of ONE [M, 2E] buffer (row stride 2E)."""
    _check(dy, g, q, attn)
    M, E = q.shape
    if packed_parts:
        parts = torch.empty((M, 2 * E), device=q.device, dtype=q.dtype)
        dq_part, dattn_part = parts[:, :E], parts[:, E:]
    else:
        dq_part = torch.empty_like(q)
        dattn_part = torch.empty_like(q)
    dgate_pre = torch.empty_like(q)
    dgamma = dgamma if dgamma is not None else torch.empty((E,), device=q.device, dtype=torch.float32)
    dbeta = dbeta if dbeta is not None else torch.empty((E,), device=q.device, dtype=torch.float32)
    nblk = _lib.load().mmsa_gate_ln_bwd_blocks(M)
    partials = torch.empty((nblk, 2, E), device=q.device, dtype=torch.float32)
    call("mmsa_gate_ln_bwd", dt(q), M, E, dy.data_ptr(), rows_per_sample, g.data_ptr(), q.data_ptr(), attn.data_ptr(),
         gamma.data_ptr(), mean.data_ptr(), rstd.data_ptr(), _p(dq_bcast), bcast_rows, _p(dq_add),
         dq_part.data_ptr(), dattn_part.data_ptr(), dq_part.stride(0), dgate_pre.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(),
         partials.data_ptr(), _stream())
    return dq_part, dattn_part, dgate_pre, dgamma, dbeta


def gate_ln_pool_fwd(gate_pre: Tensor, q: Tensor, attn: Tensor, gamma: Tensor, beta: Tensor, eps: float, B: int, L: int,
                     want_q_lp: bool = True):
    """Fused sigmoid-gate + blend + LayerNorm + token mean-pool (y is never written).
    -> g [M,E], mean [M], rstd [M], pooled_y [B,E] fp32, pooled_q [B,E] fp32, pooled_q_lp [B,E] (q.dtype) or None."""
    _check(gate_pre, q, attn, gamma, beta)
    M, E = q.shape
    assert M == B * L
    g = torch.empty_like(q)
    mean = torch.empty((M,), device=q.device, dtype=torch.float32)
    rstd = torch.empty((M,), device=q.device, dtype=torch.float32)
    py = torch.empty((B, E), device=q.device, dtype=torch.float32)
    pq = torch.empty((B, E), device=q.device, dtype=torch.float32)
    pq_lp = torch.empty((B, E), device=q.device, dtype=q.dtype) if (want_q_lp and q.dtype != torch.float32) else None
    call("mmsa_gate_ln_pool_fwd", dt(q), B, L, E, gate_pre.data_ptr(), q.data_ptr(), attn.data_ptr(), gamma.data_ptr(),
         beta.data_ptr(), float(eps), g.data_ptr(), mean.data_ptr(), rstd.data_ptr(), py.data_ptr(), pq.data_ptr(),
         _p(pq_lp), _stream())
    return g, mean, rstd, py, pq, pq_lp


def gate_ln_pool_bwd(dpy: Tensor, dpq: Optional[Tensor], dq_add: Optional[Tensor], g, q, attn, gamma, mean, rstd,
                     B: int, L: int, dgamma: Optional[Tensor] = None, dbeta: Optional[Tensor] = None,
                     packed_parts: bool = False):
    _check(dpy, dpq, dq_add, g, q, attn)
    M, E = q.shape
    assert dpy.dtype == torch.float32 and (dpq is None or dpq.dtype == torch.float32)
    if packed_parts:
        parts = torch.empty((M, 2 * E), device=q.device, dtype=q.dtype)
        dq_part, dattn_part = parts[:, :E], parts[:, E:]
    else:
        dq_part = torch.empty_like(q)
        dattn_part = torch.empty_like(q)
    dgate_pre = torch.empty_like(q)
    dgamma = dgamma if dgamma is not None else torch.empty((E,), device=q.device, dtype=torch.float32)
    dbeta = dbeta if dbeta is not None else torch.empty((E,), device=q.device, dtype=torch.float32)
    nblk = _lib.load().mmsa_gate_ln_bwd_blocks(M)
    partials = torch.empty((nblk, 2, E), device=q.device, dtype=torch.float32)
    call("mmsa_gate_ln_pool_bwd", dt(q), B, L, E, dpy.data_ptr(), _p(dpq), _p(dq_add), g.data_ptr(), q.data_ptr(),
         attn.data_ptr(), gamma.data_ptr(), mean.data_ptr(), rstd.data_ptr(), dq_part.data_ptr(), dattn_part.data_ptr(),
         dq_part.stride(0), dgate_pre.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(), partials.data_ptr(), _stream())
    return dq_part, dattn_part, dgate_pre, dgamma, dbeta


def pool_fwd(x: Tensor, B: int, L: int, is_max: bool = False):
    _check(x)
    E = x.shape[-1]
    y = torch.empty((B, E), device=x.device, dtype=x.dtype)
    arg = torch.empty((B, E), device=x.device, dtype=torch.int32) if is_max else None
    call("mmsa_pool_fwd", dt(x), B, L, E, x.data_ptr(), int(is_max), y.data_ptr(), _p(arg), _stream())
    return y, arg


def pool_bwd(dy: Tensor, B: int, L: int, is_max: bool = False, argmax: Optional[Tensor] = None):
    _check(dy)
    E = dy.shape[-1]
    dx = torch.empty((B * L, E), device=dy.device, dtype=dy.dtype)
    call("mmsa_pool_bwd", dt(dy), B, L, E, dy.data_ptr(), int(is_max), _p(argmax), dx.data_ptr(), _stream())
    return dx


def _ptr_array(ts: Sequence[Optional[Tensor]]):
    arr = (ctypes.c_void_p * len(ts))()
    for i, t in enumerate(ts):
        arr[i] = None if t is None else t.data_ptr()
    return arr


def cast_multi(srcs: Sequence[Tensor], dsts: Sequence[Tensor]) -> None:
    """One launch: dsts[i] <- srcs[i] (fp32 -> bf16 weight copies of a step)."""
    if not srcs:
        return
    _check(*srcs, *dsts)
    n = len(srcs)
    a1, a2 = _ptr_array(srcs), _ptr_array(dsts)
    numel = (ctypes.c_int64 * n)(*[t.numel() for t in srcs])
    call("mmsa_cast_multi", n, ctypes.cast(a1, ctypes.c_void_p), ctypes.cast(a2, ctypes.c_void_p),
         ctypes.cast(numel, ctypes.c_void_p), dt(srcs[0]), dt(dsts[0]), _stream())


# ---- [B,*] tail: elementwise kernels read fp32 (GEMM outputs) and write `out_dtype` (GEMM operands) ----
def modal_concat_fwd(logits: Tensor, slots: Sequence[Tensor], out_dtype: torch.dtype):
    """logits [B,S] fp32, slots S x [B,E] fp32 -> w [B,S] fp32, fused [B,S*E] (out_dtype)."""
    _check(logits, *slots)
    assert logits.dtype == torch.float32 and all(s.dtype == torch.float32 for s in slots)
    B, E = slots[0].shape
    S = len(slots)
    w = torch.empty((B, S), device=logits.device, dtype=torch.float32)
    fused = torch.empty((B, S * E), device=logits.device, dtype=out_dtype)
    arr = _ptr_array(slots)
    call("mmsa_modal_concat_fwd", dt(out_dtype), B, E, S, logits.data_ptr(), ctypes.cast(arr, ctypes.c_void_p),
         w.data_ptr(), fused.data_ptr(), _stream())
    return w, fused


def modal_concat_bwd(dfused: Tensor, w: Tensor, slots: Sequence[Tensor], need: Sequence[bool], out_dtype: torch.dtype):
    """dfused [B,S*E] fp32 -> dslots fp32 (where needed), dlogits [B,S] (out_dtype)."""
    _check(dfused, w, *slots)
    assert dfused.dtype == torch.float32
    B, E = slots[0].shape
    S = len(slots)
    dslots = [torch.empty_like(s) if n else None for s, n in zip(slots, need)]
    dlogits = torch.empty((B, S), device=dfused.device, dtype=out_dtype)
    a1, a2 = _ptr_array(slots), _ptr_array(dslots)
    call("mmsa_modal_concat_bwd", dt(out_dtype), B, E, S, dfused.data_ptr(), w.data_ptr(),
         ctypes.cast(a1, ctypes.c_void_p), ctypes.cast(a2, ctypes.c_void_p), dlogits.data_ptr(), _stream())
    return dslots, dlogits


def modal_head_fwd(h_pre: Tensor, w2: Tensor, b2: Optional[Tensor], slots: Sequence[Tensor], cd: torch.dtype):
    """GELU -> Linear(Hd,S) -> softmax -> weighted concat in one launch (mmsa_modal_head_fwd).
    h_pre [B,Hd] fp32, w2 [S,Hd] in `cd`, slots S x [B,E] fp32 -> hg [B,Hd] cd, w [B,S] fp32, fused [B,S*E] fp32,
    fused_lp [B,S*E] cd (None in fp32 mode)."""
    _check(h_pre, w2, b2, *slots)
    assert h_pre.dtype == torch.float32 and w2.dtype == cd and w2.is_contiguous() and all(s.dtype == torch.float32 for s in slots)
    B, Hd = h_pre.shape
    E, S = slots[0].shape[1], len(slots)
    hg = torch.empty((B, Hd), device=h_pre.device, dtype=cd)
    w = torch.empty((B, S), device=h_pre.device, dtype=torch.float32)
    fused = torch.empty((B, S * E), device=h_pre.device, dtype=torch.float32)
    fused_lp = torch.empty((B, S * E), device=h_pre.device, dtype=cd) if cd != torch.float32 else None
    arr = _ptr_array(slots)
    call("mmsa_modal_head_fwd", dt(cd), B, E, S, Hd, h_pre.data_ptr(), w2.data_ptr(), _p(b2), ctypes.cast(arr, ctypes.c_void_p),
         hg.data_ptr(), w.data_ptr(), fused.data_ptr(), _p(fused_lp), _stream())
    return hg, w, fused, fused_lp


def modal_head_bwd(dfused: Tensor, w: Tensor, slots: Sequence[Tensor], need: Sequence[bool], h_pre: Tensor, w2: Tensor,
                   cd: torch.dtype):
    """-> dslots (fp32 where needed), dlogits [B,S] cd, dh_pre [B,Hd] cd (mmsa_modal_head_bwd)."""
    _check(dfused, w, h_pre, w2, *slots)
    assert dfused.dtype == torch.float32
    B, Hd = h_pre.shape
    E, S = slots[0].shape[1], len(slots)
    dslots = [torch.empty_like(s) if n else None for s, n in zip(slots, need)]
    dlogits = torch.empty((B, S), device=dfused.device, dtype=cd)
    dh = torch.empty((B, Hd), device=dfused.device, dtype=cd)
    a1, a2 = _ptr_array(slots), _ptr_array(dslots)
    call("mmsa_modal_head_bwd", dt(cd), B, E, S, Hd, dfused.data_ptr(), w.data_ptr(), ctypes.cast(a1, ctypes.c_void_p),
         ctypes.cast(a2, ctypes.c_void_p), h_pre.data_ptr(), w2.data_ptr(), dlogits.data_ptr(), dh.data_ptr(), _stream())
    return dslots, dlogits, dh


def act_fwd(x: Tensor, act: int, out_dtype: torch.dtype) -> Tensor:
    _check(x)
    assert x.dtype == torch.float32
    y = torch.empty(x.shape, device=x.device, dtype=out_dtype)
    call("mmsa_act_fwd", dt(out_dtype), x.numel(), x.data_ptr(), act, y.data_ptr(), _stream())
    return y


def act_bwd(x: Tensor, dy: Tensor, act: int, out_dtype: torch.dtype) -> Tensor:
    _check(x, dy)
    assert x.dtype == torch.float32 and dy.dtype == torch.float32
    dx = torch.empty(x.shape, device=x.device, dtype=out_dtype)
    call("mmsa_act_bwd", dt(out_dtype), x.numel(), x.data_ptr(), dy.data_ptr(), act, dx.data_ptr(), _stream())
    return dx


def bn_act_fwd(x: Tensor, gamma, beta, running_mean, running_var, momentum: float, eps: float, training: bool,
               order: int, dropout_p: float, keep_mask: Optional[Tensor], seed: int, offset: int,
               out_dtype: torch.dtype, rng_state: Optional[Tensor] = None, want_lp: bool = False,
               num_batches_tracked: Optional[Tensor] = None):
    """-> y, save_mean, save_rstd, keep_mask[, y_lp]; want_lp (fp32 out only): also a bf16 copy of y from the same launch.
    num_batches_tracked (int64 scalar on the device): incremented by the kernel in training mode."""
    _check(x, gamma, beta, num_batches_tracked)
    assert num_batches_tracked is None or num_batches_tracked.dtype == torch.int64
    assert x.dtype == torch.float32
    B, N = x.shape
    y = torch.empty((B, N), device=x.device, dtype=out_dtype)
    y_lp = torch.empty((B, N), device=x.device, dtype=torch.bfloat16) if (want_lp and out_dtype == torch.float32) else None
    save_mean = torch.empty((N,), device=x.device, dtype=torch.float32)
    save_rstd = torch.empty((N,), device=x.device, dtype=torch.float32)
    mask_given = keep_mask is not None
    if training and dropout_p > 0 and keep_mask is None:
        keep_mask = torch.empty((B, N), device=x.device, dtype=torch.uint8)
    call("mmsa_bn_act_fwd", dt(out_dtype), B, N, order, x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), _p(running_mean),
         _p(running_var), _p(num_batches_tracked), float(momentum), float(eps), int(training), float(dropout_p), _p(keep_mask),
         int(mask_given), seed, offset, _p(rng_state), y.data_ptr(), _p(y_lp), save_mean.data_ptr(), save_rstd.data_ptr(),
         _stream())
    if want_lp:
        return y, save_mean, save_rstd, keep_mask, y_lp
    return y, save_mean, save_rstd, keep_mask


def linear_bn_act_ok(x: Tensor, w: Tensor) -> bool:
    """whether mmsa_linear_bn_act_fwd takes this Linear (bf16 operands, batch <= 256, N % 8 == 0, K % 8 == 0, aligned)"""
    if x.dtype != torch.bfloat16 or w.dtype != torch.bfloat16 or x.dim() != 2 or w.dim() != 2:
        return False
    if x.stride(1) != 1 or w.stride(1) != 1 or x.data_ptr() % 16 or w.data_ptr() % 16:
        return False
    return bool(_lib.load().mmsa_linear_bn_act_supported(x.shape[0], w.shape[0], x.shape[1], x.stride(0), w.stride(0)))


def linear_bn_act_fwd(x: Tensor, w: Tensor, bias: Optional[Tensor], gamma, beta, running_mean, running_var, momentum: float,
                      eps: float, training: bool, order: int, dropout_p: float, keep_mask: Optional[Tensor], seed: int,
                      offset: int, out_dtype: torch.dtype, rng_state: Optional[Tensor] = None, want_lp: bool = False,
                      num_batches_tracked: Optional[Tensor] = None):
    """Linear + BatchNorm block in one launch (mmsa_linear_bn_act_fwd): -> z [M,N] fp32 (= x w^T + bias), then exactly what
    bn_act_fwd(z, ...) returns: y, save_mean, save_rstd, keep_mask[, y_lp]."""
    _check(x, w, bias, gamma, beta, num_batches_tracked)
    assert w.shape[1] == x.shape[1]
    assert num_batches_tracked is None or num_batches_tracked.dtype == torch.int64
    M, Kd = x.shape
    N = w.shape[0]
    z = torch.empty((M, N), device=x.device, dtype=torch.float32)
    y = torch.empty((M, N), device=x.device, dtype=out_dtype)
    y_lp = torch.empty((M, N), device=x.device, dtype=torch.bfloat16) if (want_lp and out_dtype == torch.float32) else None
    save_mean = torch.empty((N,), device=x.device, dtype=torch.float32)
    save_rstd = torch.empty((N,), device=x.device, dtype=torch.float32)
    mask_given = keep_mask is not None
    if training and dropout_p > 0 and keep_mask is None:
        keep_mask = torch.empty((M, N), device=x.device, dtype=torch.uint8)
    call("mmsa_linear_bn_act_fwd", M, N, Kd, x.data_ptr(), x.stride(0), w.data_ptr(), w.stride(0), _p(bias), gamma.data_ptr(),
         beta.data_ptr(), _p(running_mean), _p(running_var), _p(num_batches_tracked), float(momentum), float(eps),
         int(training), order, float(dropout_p), _p(keep_mask), int(mask_given), seed, offset, _p(rng_state), z.data_ptr(),
         dt(out_dtype), y.data_ptr(), _p(y_lp), save_mean.data_ptr(), save_rstd.data_ptr(), _stream())
    if want_lp:
        return z, y, save_mean, save_rstd, keep_mask, y_lp
    return z, y, save_mean, save_rstd, keep_mask


def bn_act_bwd(x, dy, gamma, beta, save_mean, save_rstd, training: bool, order: int, dropout_p: float, keep_mask,
               out_dtype: torch.dtype):
    _check(x, dy)
    assert x.dtype == torch.float32 and dy.dtype == torch.float32
    B, N = x.shape
    dx = torch.empty((B, N), device=x.device, dtype=out_dtype)
    dgamma = torch.empty((N,), device=x.device, dtype=torch.float32)
    dbeta = torch.empty((N,), device=x.device, dtype=torch.float32)
    dbias_prev = torch.empty((N,), device=x.device, dtype=torch.float32)
    call("mmsa_bn_act_bwd", dt(out_dtype), B, N, order, x.data_ptr(), dy.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
         save_mean.data_ptr(), save_rstd.data_ptr(), int(training), float(dropout_p), _p(keep_mask), dx.data_ptr(),
         dgamma.data_ptr(), dbeta.data_ptr(), dbias_prev.data_ptr(), _stream())
    return dx, dgamma, dbeta, dbias_prev


def dropout(x: Tensor, p: float, keep_mask: Optional[Tensor], mask_given: bool, seed: int, offset: int,
            out_dtype: torch.dtype, rng_state: Optional[Tensor] = None):
    _check(x, keep_mask)
    assert x.dtype == torch.float32
    y = torch.empty(x.shape, device=x.device, dtype=out_dtype)
    if keep_mask is None:
        keep_mask = torch.empty(x.shape, device=x.device, dtype=torch.uint8)
    call("mmsa_dropout", dt(out_dtype), x.numel(), x.data_ptr(), float(p), keep_mask.data_ptr(), int(mask_given), seed,
         offset, _p(rng_state), y.data_ptr(), _stream())
    return y, keep_mask


def rng_advance(rng_state: Tensor, n: int) -> None:
    """rng_state[1] += n on the current stream (a kernel node: replays of a captured graph keep advancing)."""
    _check(rng_state)
    assert rng_state.dtype == torch.int64 and rng_state.numel() >= 2
    call("mmsa_rng_advance", rng_state.data_ptr(), int(n), _stream())


def ce_fwd(logits: Tensor, labels: Tensor, addend: Optional[Tensor] = None):
    """-> loss [1] (= mean CE + sum(addend)), pred [B]; addend: up to 64 fp32 device values."""
    _check(logits, labels, addend)
    assert addend is None or (addend.dtype == torch.float32 and addend.is_contiguous() and addend.numel() <= 64)
    B, C = logits.shape
    loss = torch.empty((1,), device=logits.device, dtype=torch.float32)
    pred = torch.empty((B,), device=logits.device, dtype=torch.int64)
    row = torch.empty((B,), device=logits.device, dtype=torch.float32)
    call("mmsa_ce_fwd", B, C, logits.data_ptr(), labels.data_ptr(), _p(addend), 0 if addend is None else addend.numel(),
         loss.data_ptr(), pred.data_ptr(), row.data_ptr(), _stream())
    return loss, pred


def ce_bwd(logits: Tensor, labels: Tensor, dloss: Tensor, out_dtype: torch.dtype = torch.float32) -> Tensor:
    _check(logits, labels, dloss)
    B, C = logits.shape
    dlogits = torch.empty((B, C), device=logits.device, dtype=out_dtype)
    call("mmsa_ce_bwd", dt(out_dtype), B, C, logits.data_ptr(), labels.data_ptr(), dloss.data_ptr(), dlogits.data_ptr(),
         _stream())
    return dlogits


def l2norm_fwd(x: Tensor):
    _check(x)
    B, E = x.shape
    y = torch.empty_like(x)
    norm = torch.empty((B,), device=x.device, dtype=torch.float32)
    call("mmsa_l2norm_fwd", dt(x), B, E, x.data_ptr(), y.data_ptr(), norm.data_ptr(), _stream())
    return y, norm


def l2norm_bwd(y: Tensor, norm: Tensor, dy1: Tensor, dy2: Optional[Tensor]) -> Tensor:
    _check(y, norm, dy1, dy2)
    B, E = y.shape
    dx = torch.empty_like(y)
    call("mmsa_l2norm_bwd", dt(y), B, E, y.data_ptr(), norm.data_ptr(), dy1.data_ptr(), _p(dy2), dx.data_ptr(),
         _stream())
    return dx


def contrastive_fwd(kind: int, sim: Tensor, labels_rows, labels_cols, temperature: Optional[Tensor],
                    temperature_const: float, row_offset: int, denom: int, weight: Optional[Tensor] = None):
    """-> loss [1], row stats [B,4], unweighted loss [1] or None.  weight (fp32 device scalar): loss = weight * mean, from
    the same launch; the unweighted mean is kept for d weight."""
    _check(sim, labels_rows, labels_cols, temperature, weight)
    assert weight is None or (weight.dtype == torch.float32 and weight.numel() == 1)
    B, Bg = sim.shape
    stats = torch.empty((B, 4), device=sim.device, dtype=torch.float32)
    row_loss = torch.empty((B,), device=sim.device, dtype=torch.float32)
    loss = torch.empty((1,), device=sim.device, dtype=torch.float32)
    raw = torch.empty((1,), device=sim.device, dtype=torch.float32) if weight is not None else None
    call("mmsa_contrastive_fwd", kind, B, Bg, row_offset, sim.data_ptr(), _p(labels_rows), _p(labels_cols),
         _p(temperature), float(temperature_const), denom, _p(weight), stats.data_ptr(), row_loss.data_ptr(), loss.data_ptr(),
         _p(raw), _stream())
    return loss, stats, raw


def contrastive_bwd(kind: int, sim: Tensor, labels_rows, labels_cols, temperature: Optional[Tensor],
                    temperature_const: float, row_offset: int, denom: int, stats: Tensor, dloss: Tensor,
                    g_dtype: torch.dtype, weight: Optional[Tensor] = None, loss_raw: Optional[Tensor] = None,
                    want_dweight: bool = False):
    """-> G [B,Bg], dtemp [1], dweight [1] or None (weight / loss_raw as in contrastive_fwd)."""
    _check(sim, stats, dloss, weight, loss_raw)
    B, Bg = sim.shape
    G = torch.empty((B, Bg), device=sim.device, dtype=g_dtype)
    dtemp_rows = torch.empty((B,), device=sim.device, dtype=torch.float32)
    dtemp = torch.empty((1,), device=sim.device, dtype=torch.float32)
    dweight = torch.empty((1,), device=sim.device, dtype=torch.float32) if (want_dweight and weight is not None) else None
    call("mmsa_contrastive_bwd", kind, B, Bg, row_offset, sim.data_ptr(), _p(labels_rows), _p(labels_cols),
         _p(temperature), float(temperature_const), denom, stats.data_ptr(), dloss.data_ptr(), _p(weight), _p(loss_raw),
         G.data_ptr(), dt(g_dtype), dtemp_rows.data_ptr(), dtemp.data_ptr(), _p(dweight), _stream())
    return G, dtemp, dweight


# ---- encoder tail: residual add + LayerNorm, positional table ----
def add_ln_fwd(x: Tensor, r: Optional[Tensor], gamma: Tensor, beta: Tensor, eps: float):
    _check(x, r, gamma, beta)
    M, E = x.shape
    y = torch.empty_like(x)
    mean = torch.empty((M,), device=x.device, dtype=torch.float32)
    rstd = torch.empty((M,), device=x.device, dtype=torch.float32)
    call("mmsa_add_ln_fwd", dt(x), M, E, x.data_ptr(), _p(r), gamma.data_ptr(), beta.data_ptr(), float(eps), y.data_ptr(),
         mean.data_ptr(), rstd.data_ptr(), _stream())
    return y, mean, rstd


def add_ln_bwd(dy: Tensor, x: Tensor, r: Optional[Tensor], gamma: Tensor, mean: Tensor, rstd: Tensor):
    _check(dy, x, r)
    M, E = x.shape
    du = torch.empty_like(x)
    dgamma = torch.empty((E,), device=x.device, dtype=torch.float32)
    dbeta = torch.empty((E,), device=x.device, dtype=torch.float32)
    nblk = _lib.load().mmsa_gate_ln_bwd_blocks(M)
    partials = torch.empty((nblk, 2, E), device=x.device, dtype=torch.float32)
    call("mmsa_add_ln_bwd", dt(x), M, E, dy.data_ptr(), x.data_ptr(), _p(r), gamma.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
         du.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(), partials.data_ptr(), _stream())
    return du, dgamma, dbeta


def add_rows(x: Tensor, pe: Tensor, L: int) -> Tensor:
    """y[m,:] = x[m,:] + pe[m % L,:]; pe fp32 [>=L, E] contiguous."""
    _check(x, pe)
    M, E = x.shape
    assert pe.dtype == torch.float32 and pe.is_contiguous() and pe.shape[-1] == E and pe.shape[-2] >= L
    y = torch.empty_like(x)
    call("mmsa_add_rows", dt(x), M, E, L, x.data_ptr(), pe.data_ptr(), y.data_ptr(), _stream())
    return y
