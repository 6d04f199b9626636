"""Kernel-level parity (-m gpu): every C-ABI entry point against a plain PyTorch fp32/fp64 statement
of the same op on identical seeded inputs.  Tolerances (parity_util.rel_err, relative to tensor
scale): fp32 storage 1e-5, bf16 storage 2e-2 (BASELINE.json north_star)."""
import math
import os

import pytest
import torch
import torch.nn.functional as F

from parity_util import check_close, rel_err, report_summary

pytestmark = pytest.mark.gpu

TOL = {torch.float32: 1e-5, torch.bfloat16: 2e-2}


def _k():
    from mmsa import kernels
    return kernels


def _rand(shape, dtype, dev, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dev).to(dtype)


# ---------------------------------------------------------------------------------------- GEMMs
GEMM_SHAPES = [
    (256, 768, 768),      # one E x E projection, 2 m-tiles
    (392, 768, 2048),     # image projection at B=8 (ragged M: 392 = 3*128 + 8)
    (1000, 1536, 768),    # packed K/V projection, ragged M
    (64, 256, 2304),      # fusion.0 (small M)
    (37, 64, 1536),       # tiny ragged everything
    (300, 3, 128),        # class head, N = 3 (CUDA-core path)
    (128, 128, 72),       # K not a multiple of 64 (TMA zero fill)
    (4352, 768, 768),     # 2-CTA (cta_group::2) 256 x 256 tiles, 17 pair-rows
    (5000, 1536, 264),    # 2-CTA, ragged M (5000 = 19*256 + 136: the peer CTA of the last pair is half empty) and ragged K
]


@pytest.fixture(params=[0, 1], ids=["auto", "tcgen05"])
def gemm_engine(request):
    """bf16 products below 0.13 GFLOP take the mma.sync cluster kernel (gemm_small.cu) by default; engine 1 keeps
    them on the tcgen05 kernel so that both engines see the ragged shapes."""
    from mmsa import _lib
    lib = _lib.load()
    lib.mmsa_debug_gemm_engine(request.param)
    yield request.param
    lib.mmsa_debug_gemm_engine(0)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_linear_fwd(cuda_device, gemm_engine, dtype, M, N, K):
    if dtype == torch.float32 and gemm_engine:
        pytest.skip("engine switch only affects bf16 storage")
    k = _k()
    x = _rand((M, K), dtype, cuda_device, 1)
    w = _rand((N, K), dtype, cuda_device, 2, 1 / math.sqrt(K))
    b = _rand((N,), torch.float32, cuda_device, 3)
    r = _rand((M, N), dtype, cuda_device, 4)
    y = k.linear_fwd(x, w, b, residual=r)
    ref = x.double() @ w.double().T + b.double() + r.double()
    assert rel_err(y, ref) <= TOL[dtype]
    # fp32 output from bf16 operands (logits path)
    y32 = k.linear_fwd(x, w, b, out_dtype=torch.float32)
    assert y32.dtype == torch.float32
    assert rel_err(y32, x.double() @ w.double().T + b.double()) <= (1e-5 if dtype == torch.float32 else 2e-3)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_linear_fwd_split_operand(cuda_device, dtype):
    """gate GEMM: cat[q, attn] @ Wg^T without the concat (MultimodalModel.py:147)."""
    k = _k()
    M, E = (520 if dtype == torch.float32 else 4500), 768       # bf16: 2-CTA tiles with the two-tensor A operand
    q = _rand((M, E), dtype, cuda_device, 1)
    a = _rand((M, E), dtype, cuda_device, 2)
    w = _rand((E, 2 * E), dtype, cuda_device, 3, 0.02)
    b = _rand((E,), torch.float32, cuda_device, 4)
    y = k.linear_fwd(q, w, b, x2=a)
    ref = torch.cat([q, a], 1).double() @ w.double().T + b.double()
    assert rel_err(y, ref) <= TOL[dtype]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("M,N,K", [(256, 768, 768), (392, 1536, 768), (77, 3, 128), (640, 768, 1536), (4200, 768, 768),
                                   (6272, 1536, 768)])
def test_linear_dgrad(cuda_device, gemm_engine, dtype, M, N, K):
    if dtype == torch.float32 and gemm_engine:
        pytest.skip("engine switch only affects bf16 storage")
    k = _k()
    dy = _rand((M, N), dtype, cuda_device, 1)
    wfull = _rand((N, K + 64), dtype, cuda_device, 2, 1 / math.sqrt(N))
    w = wfull[:, 64:]                      # column block of a wider weight (ld != K)
    r = _rand((M, K), dtype, cuda_device, 3)
    dx = k.linear_dgrad(dy, w, residual=r)
    ref = dy.double() @ w.double() + r.double()
    assert rel_err(dx, ref) <= TOL[dtype]
    dx0 = k.linear_dgrad(dy, w, out_dtype=torch.float32)          # no residual: the skinny (N <= 8) kernel in bf16 mode
    assert rel_err(dx0, dy.double() @ w.double()) <= (1e-5 if dtype == torch.float32 else 2e-3)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("M,N,K", [(1024, 768, 768), (392, 768, 2048), (5000, 1536, 768), (300, 3, 128), (64, 256, 2304),
                                   (256, 256, 2304), (100, 24, 40),
                                   # 2-CTA 256x256 tiles + workspace split-K (tokens >= 4096, dW rows % 256 == 0): ragged token
                                   # count (TMA zero fill of the last k-block), ragged dW columns, the configs[1] shape
                                   (4100, 768, 2048), (8192, 256, 520), (32768, 768, 768)])
def test_linear_wgrad(cuda_device, gemm_engine, dtype, M, N, K):
    if dtype == torch.float32 and gemm_engine:
        pytest.skip("engine switch only affects bf16 storage")
    k = _k()
    dy = _rand((M, N), dtype, cuda_device, 1)
    x = _rand((M, K), dtype, cuda_device, 2)
    dwfull = torch.zeros((N, K + 8), device=cuda_device)
    dw, db = k.linear_wgrad(dy, x, dw=dwfull[:, 8:])
    ref_w = dy.double().T @ x.double()
    ref_b = dy.double().sum(0)
    tol = 1e-5 if dtype == torch.float32 else 2e-3     # bf16 operands, fp32 accumulate and output
    assert rel_err(dw, ref_w) <= tol
    assert rel_err(db, ref_b) <= tol
    assert float(dwfull[:, :8].abs().max()) == 0.0
    dw2, db2 = k.linear_wgrad(dy, x)                   # deterministic: slice-ordered sums, no atomics
    assert torch.equal(dw2, dw) and torch.equal(db2, db)


@pytest.mark.parametrize("M,N,K,K2", [(100, 256, 2304, 0), (256, 64, 768, 768), (100, 37, 72, 0), (33, 130, 8, 0),
                                      (256, 128, 256, 0), (50, 768, 768, 768), (1, 16, 4096, 0)])
@pytest.mark.parametrize("act", [0, 1, 2, 3])
def test_linear_small_engine(cuda_device, M, N, K, K2, act):
    """gemm_small.cu (mma.sync + cluster split-K): ragged edges, two-tensor A operand, activations, bf16 and fp32
    outputs, and bit-equality of repeated launches (slice-ordered DSMEM sum)."""
    k = _k()
    x = _rand((M, K), torch.bfloat16, cuda_device, 1)
    x2 = _rand((M, K2), torch.bfloat16, cuda_device, 2) if K2 else None
    w = _rand((N, K + K2), torch.bfloat16, cuda_device, 3, 1 / math.sqrt(K + K2))
    b = _rand((N,), torch.float32, cuda_device, 4)
    r = _rand((M, N), torch.bfloat16, cuda_device, 5)
    xa = torch.cat([x, x2], 1) if K2 else x
    pre = xa.double() @ w.double().T + b.double() + r.double()
    ref = [pre, torch.sigmoid(pre), F.gelu(pre), torch.relu(pre)][act]
    y = k.linear_fwd(x, w, b, residual=r, x2=x2, act=act)
    assert rel_err(y, ref) <= 2e-2
    y32 = k.linear_fwd(x, w, b, residual=r, x2=x2, act=act, out_dtype=torch.float32)
    assert rel_err(y32, ref) <= 2e-3
    assert torch.equal(y32, k.linear_fwd(x, w, b, residual=r, x2=x2, act=act, out_dtype=torch.float32))


# ---------------------------------------------------------------------------------------- attention
def _attn_ref(q, k, v, B, H, Lq, Lk, D):
    qd = q.double().view(B, Lq, H, D).transpose(1, 2)
    kd = k.double().view(B, Lk, H, D).transpose(1, 2)
    vd = v.double().view(B, Lk, H, D).transpose(1, 2)
    s = (qd / math.sqrt(D)) @ kd.transpose(-1, -2)
    p = torch.softmax(s, -1)
    o = (p @ vd).transpose(1, 2).reshape(B * Lq, H * D)
    return o, torch.logsumexp(s, -1)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,H,Lq,Lk,D", [(3, 12, 64, 49, 64), (2, 12, 49, 128, 64), (2, 4, 1, 1, 64),
                                         (5, 8, 3, 3, 32), (1, 12, 130, 70, 64), (2, 12, 49, 512, 64),
                                         (2, 12, 512, 49, 64), (3, 12, 128, 49, 64), (2, 2, 300, 64, 64)])
def test_attention(cuda_device, dtype, B, H, Lq, Lk, D):
    k_ = _k()
    E = H * D
    q = _rand((B * Lq, E), dtype, cuda_device, 1)
    kv = _rand((B * Lk, 2 * E), dtype, cuda_device, 2)        # packed K|V, row stride 2E
    kk, vv = kv[:, :E], kv[:, E:]
    do = _rand((B * Lq, E), dtype, cuda_device, 3)
    o, lse = k_.attn_fwd(q, kk, vv, B, H, Lq, Lk, D)
    qr, kr, vr = (t.detach().double().requires_grad_(True) for t in (q, kk, vv))
    o_ref, lse_ref = _attn_ref(qr, kr, vr, B, H, Lq, Lk, D)
    assert rel_err(o, o_ref) <= TOL[dtype]
    assert rel_err(lse, lse_ref) <= (1e-5 if dtype == torch.float32 else 2e-3)
    o_ref.backward(do.double())
    dq = torch.empty_like(q)
    dkv = torch.empty_like(kv)
    k_.attn_bwd(q, kk, vv, o, do, lse, B, H, Lq, Lk, D, dq, dkv[:, :E], dkv[:, E:])
    if Lk == 1:     # softmax over one key is constant: dq = dk = 0 exactly; compare against the scale of dv
        assert float(dq.float().abs().max()) <= TOL[dtype] * float(vr.grad.abs().max())
        assert float(dkv[:, :E].float().abs().max()) <= TOL[dtype] * float(vr.grad.abs().max())
    else:
        assert rel_err(dq, qr.grad) <= TOL[dtype]
        assert rel_err(dkv[:, :E], kr.grad) <= TOL[dtype]
    assert rel_err(dkv[:, E:], vr.grad) <= TOL[dtype]


def test_attention_engines_agree(cuda_device):
    """bf16: the tcgen05/TMA engine (forward + fused backward) vs the mma.sync engine vs the CUDA-core engine."""
    from mmsa import _lib
    k_ = _k()
    lib = _lib.load()
    for (B, H, Lq, Lk) in [(2, 12, 128, 49), (2, 12, 49, 128), (1, 4, 256, 33)]:
        D = 64
        E = H * D
        q = _rand((B * Lq, E), torch.bfloat16, cuda_device, 1)
        kv = _rand((B * Lk, 2 * E), torch.bfloat16, cuda_device, 2)
        do = _rand((B * Lq, E), torch.bfloat16, cuda_device, 3)
        outs = {}
        for name, eng, simt in (("tcgen05", 0, 0), ("mma.sync", 1, 0), ("simt", 0, 1)):
            lib.mmsa_debug_attention_engine(eng)
            lib.mmsa_debug_force_simt_attention(simt)
            try:
                o, lse = k_.attn_fwd(q, kv[:, :E], kv[:, E:], B, H, Lq, Lk, D)
                dq = torch.empty_like(q)
                dkv = torch.empty_like(kv)
                k_.attn_bwd(q, kv[:, :E], kv[:, E:], o, do, lse, B, H, Lq, Lk, D, dq, dkv[:, :E], dkv[:, E:])
            finally:
                lib.mmsa_debug_attention_engine(0)
                lib.mmsa_debug_force_simt_attention(0)
            outs[name] = (o, lse, dq, dkv)
        for other in ("mma.sync", "simt"):
            for a, b, tol in zip(outs["tcgen05"], outs[other], (2e-2, 2e-3, 2e-2, 2e-2)):
                assert rel_err(a, b) <= tol, (other, Lq, Lk)


# ---------------------------------------------------------------------------------------- gate + LN
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("M,E,L", [(8 * 64, 768, 64), (5 * 49, 768, 49), (20, 256, 1), (3 * 7, 1024, 7)])
def test_gate_ln(cuda_device, dtype, M, E, L):
    k = _k()
    gp = _rand((M, E), dtype, cuda_device, 1)
    q = _rand((M, E), dtype, cuda_device, 2)
    a = _rand((M, E), dtype, cuda_device, 3)
    gamma = _rand((E,), torch.float32, cuda_device, 4) * 0.1 + 1
    beta = _rand((E,), torch.float32, cuda_device, 5) * 0.1
    g, y, mean, rstd = k.gate_ln_fwd(gp, q, a, gamma, beta, 1e-5)

    def ref(gp_, q_, a_, gm, bt):
        gg = torch.sigmoid(gp_)
        u = gg * q_ + (1 - gg) * a_
        return F.layer_norm(u, (E,), gm, bt, 1e-5)

    ins = [t.detach().double().requires_grad_(True) for t in (gp, q, a, gamma, beta)]
    y_ref = ref(*ins)
    assert rel_err(y, y_ref) <= TOL[dtype]
    assert rel_err(g, torch.sigmoid(gp.double())) <= TOL[dtype]
    # backward, token-pooled form: dy is [M/L, E] broadcast over tokens / L, plus a pooled grad for q
    B = M // L
    dyp = _rand((B, E), dtype, cuda_device, 6)
    dqb = _rand((B, E), dtype, cuda_device, 7)
    dqa = _rand((M, E), dtype, cuda_device, 8)
    (y_ref.view(B, L, E).mean(1) * dyp.double()).sum().backward()
    dq_part, da_part, dgp, dgamma, dbeta = k.gate_ln_bwd(dyp, L, g, q, a, gamma, mean, rstd, dq_bcast=dqb, bcast_rows=L,
                                                       dq_add=dqa)
    g_ = g.double()
    du_q = ins[1].grad      # = du*g
    extra = dqb.double().repeat_interleave(L, 0) / L + dqa.double()
    tol = TOL[dtype]
    assert rel_err(dq_part, du_q + extra) <= tol
    assert rel_err(da_part, ins[2].grad) <= tol
    why_dgate = ("bf16 storage: the gate g is ROUNDED to bf16 before g(1-g) and (q-attn) are formed (so that forward and "
                 "backward see the same g); g(1-g) near saturation carries that 2^-9 relative error of g at full size")
    check_close(f"gate_ln[{dtype}]", "dgate_pre (un-pooled bwd)", dgp, ins[0].grad, tol,
                fixed_bar=(tol if dtype == torch.float32 else 4e-2), why=why_dgate)
    assert rel_err(dgamma, ins[3].grad) <= (1e-5 if dtype == torch.float32 else 2e-2)
    assert rel_err(dbeta, ins[4].grad) <= (1e-5 if dtype == torch.float32 else 2e-2)
    # fused forward/backward with the token mean-pool (no y in HBM, fp32 pooled outputs and gradients)
    g2, mean2, rstd2, py, pq, pq_lp = k.gate_ln_pool_fwd(gp, q, a, gamma, beta, 1e-5, B, L)
    assert torch.equal(g2, g) and rel_err(mean2, mean) <= 1e-6 and rel_err(rstd2, rstd) <= 1e-6
    # (bf16: g is rounded to storage precision before the blend, so that forward and backward agree)
    assert py.dtype == torch.float32 and rel_err(py, y_ref.view(B, L, E).mean(1)) <= TOL[dtype]
    assert rel_err(pq, q.double().view(B, L, E).mean(1)) <= 1e-5
    if dtype == torch.bfloat16:
        assert pq_lp.dtype == dtype and rel_err(pq_lp, pq) <= 4e-3
    dq2, da2, dgp2, dgamma2, dbeta2 = k.gate_ln_pool_bwd(dyp.float(), dqb.float(), dqa, g, q, a, gamma, mean, rstd, B, L)
    assert rel_err(dq2, du_q + extra) <= tol and rel_err(da2, ins[2].grad) <= tol
    check_close(f"gate_ln[{dtype}]", "dgate_pre (pooled bwd)", dgp2, ins[0].grad, tol,
                fixed_bar=(tol if dtype == torch.float32 else 4e-2), why=why_dgate)
    assert rel_err(dgamma2, ins[3].grad) <= (1e-5 if dtype == torch.float32 else 2e-2)
    assert rel_err(dbeta2, ins[4].grad) <= (1e-5 if dtype == torch.float32 else 2e-2)
    # plain form
    dy = _rand((M, E), dtype, cuda_device, 9)
    for t in ins:
        t.grad = None
    ref(*ins).backward(dy.double())
    dq_part, da_part, dgp, dgamma, dbeta = k.gate_ln_bwd(dy, 0, g, q, a, gamma, mean, rstd)
    assert rel_err(dq_part, ins[1].grad) <= tol and rel_err(da_part, ins[2].grad) <= tol
    assert rel_err(dgamma, ins[3].grad) <= (1e-5 if dtype == torch.float32 else 2e-2)


# ---------------------------------------------------------------------------------------- pooling / concat
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,L,E", [(8, 64, 768), (5, 49, 768), (6, 3, 256), (2, 512, 768), (3, 1, 256)])
def test_pool(cuda_device, dtype, B, L, E):
    k = _k()
    x = _rand((B * L, E), dtype, cuda_device, 1)
    y, _ = k.pool_fwd(x, B, L)
    assert rel_err(y, x.double().view(B, L, E).mean(1)) <= TOL[dtype]
    ym, arg = k.pool_fwd(x, B, L, is_max=True)
    vals, idx = x.float().view(B, L, E).max(1)
    assert torch.equal(ym.float(), vals)
    assert torch.equal(arg.long(), idx)
    dy = _rand((B, E), dtype, cuda_device, 2)
    dx = k.pool_bwd(dy, B, L)
    assert rel_err(dx, (dy.double() / L).repeat_interleave(L, 0)) <= TOL[dtype]
    dxm = k.pool_bwd(dy, B, L, True, arg)
    ref = torch.zeros(B, L, E, dtype=torch.float64, device=cuda_device).scatter_(1, idx.unsqueeze(1), dy.double().unsqueeze(1))
    assert rel_err(dxm, ref.view(B * L, E)) <= TOL[dtype]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_modal_concat(cuda_device, dtype):
    k = _k()
    B, E, S = 37, 768, 3
    logits = _rand((B, S), torch.float32, cuda_device, 1)          # fp32 in (GEMM outputs), `dtype` out
    slots = [_rand((B, E), torch.float32, cuda_device, 2 + i) for i in range(S)]
    w, fused = k.modal_concat_fwd(logits, slots, dtype)
    assert fused.dtype == dtype
    lr = logits.detach().double().requires_grad_(True)
    sr = [s.detach().double().requires_grad_(True) for s in slots]
    wr = torch.softmax(lr, 1)
    fr = torch.cat([sr[i] * wr[:, i:i + 1] for i in range(S)], 1)
    assert rel_err(w, wr) <= 1e-5 and rel_err(fused, fr) <= TOL[dtype]
    df = _rand((B, S * E), torch.float32, cuda_device, 9)
    fr.backward(df.double())
    dslots, dlog = k.modal_concat_bwd(df, w, slots, [True] * S, dtype)
    for i in range(S):
        assert dslots[i].dtype == torch.float32 and rel_err(dslots[i], sr[i].grad) <= 1e-5
    assert dlog.dtype == dtype and rel_err(dlog, lr.grad) <= TOL[dtype]


# ---------------------------------------------------------------------------------------- BN / dropout / act
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("order", [0, 1])
@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("B", [48, 256, 300])       # <= 256 rows: register-resident kernels; above: the looping ones
def test_bn_act(cuda_device, dtype, order, training, B):
    k = _k()
    N, p = 200, 0.3
    x = _rand((B, N), torch.float32, cuda_device, 1)               # fp32 in (GEMM output), `dtype` out
    gamma = _rand((N,), torch.float32, cuda_device, 2) * 0.1 + 1
    beta = _rand((N,), torch.float32, cuda_device, 3) * 0.1
    rm = _rand((N,), torch.float32, cuda_device, 4) * 0.1
    rv = _rand((N,), torch.float32, cuda_device, 5).abs() + 0.5
    keep = (torch.rand(B, N, generator=torch.Generator().manual_seed(6)) >= p).to(torch.uint8).to(cuda_device)
    rm_k, rv_k = rm.clone(), rv.clone()
    y, mean, rstd, mask = k.bn_act_fwd(x, gamma, beta, rm_k, rv_k, 0.1, 1e-5, training, order, p if training else 0.0,
                                       keep if training else None, 0, 0, dtype)
    assert y.dtype == dtype
    xr, gr, br = (t.detach().double().requires_grad_(True) for t in (x, gamma, beta))
    rm_r, rv_r = rm.double().clone(), rv.double().clone()
    h = F.relu(xr) if order == 1 else xr
    h = F.batch_norm(h, rm_r, rv_r, gr, br, training, 0.1, 1e-5)
    if order == 0:
        h = F.gelu(h)
    if training:
        h = h * keep.double() / (1 - p)
    assert rel_err(y, h) <= TOL[dtype]
    if training:
        assert rel_err(rm_k, rm_r) <= 1e-5 and rel_err(rv_k, rv_r) <= 1e-5
    dy = _rand((B, N), torch.float32, cuda_device, 7)
    h.backward(dy.double())
    dx, dg, db, dbp = k.bn_act_bwd(x, dy, gamma, beta, mean, rstd, training, order, p if training else 0.0,
                              mask if training else None, dtype)
    assert dx.dtype == dtype and rel_err(dx, xr.grad) <= TOL[dtype]
    assert rel_err(dg, gr.grad) <= 1e-5 and rel_err(db, br.grad) <= 1e-5    # batch reductions run on fp32 values
    # column sums of dx taken before rounding (bias gradient of the Linear in front; ~0 under training BN)
    assert float((dbp.double() - xr.grad.sum(0)).abs().max()) <= 1e-5 * float(xr.grad.abs().max()) * B


@pytest.mark.parametrize("M,N,K", [(256, 256, 2304), (256, 128, 256), (256, 128, 128), (6, 128, 128), (200, 64, 136),
                                   (33, 24, 64), (1, 16, 8), (256, 256, 4096)])
@pytest.mark.parametrize("order", [0, 1, 2])
@pytest.mark.parametrize("training", [True, False])
def test_linear_bn_act_one_launch(cuda_device, M, N, K, order, training):
    """mmsa_linear_bn_act_fwd (Linear + BatchNorm1d + act + dropout in one launch, MultimodalModel.py:179-199) against the
    two-launch form it replaces (mmsa_linear_fwd, then mmsa_bn_act_fwd) AND against float64 torch: z, y (fp32 with a bf16
    copy, and bf16), batch statistics, running statistics, num_batches_tracked, the Philox keep mask (identical bits: same
    indexing) and an injected mask.  Shapes: the tail's layers at B = 256 (K = 2304 splits the reduction over a cluster of 6,
    K = 4096 over 8), a tiny batch, ragged K (136: zero-filled tail block), N not a multiple of the 16-column CTA tile, M = 1."""
    k = _k()
    if not training and order == 2:
        pytest.skip("same code path as order 0 in eval mode")
    p = 0.3
    x = _rand((M, K), torch.bfloat16, cuda_device, 1)
    w = _rand((N, K), torch.bfloat16, cuda_device, 2, 1 / math.sqrt(K))
    bias = _rand((N,), torch.float32, cuda_device, 3, 0.2)
    gamma = _rand((N,), torch.float32, cuda_device, 4) * 0.1 + 1
    beta = _rand((N,), torch.float32, cuda_device, 5) * 0.1
    rm = _rand((N,), torch.float32, cuda_device, 6) * 0.1
    rv = _rand((N,), torch.float32, cuda_device, 7).abs() + 0.5
    assert k.linear_bn_act_ok(x, w)
    state = torch.tensor([1234, 1000], dtype=torch.int64, device=cuda_device)
    # two-launch form
    rm_a, rv_a, nbt_a = rm.clone(), rv.clone(), torch.zeros((), dtype=torch.int64, device=cuda_device)
    z_a = k.linear_fwd(x, w, bias, out_dtype=torch.float32)
    y_a, mean_a, rstd_a, mask_a, ylp_a = k.bn_act_fwd(z_a, gamma, beta, rm_a, rv_a, 0.1, 1e-5, training, order,
                                                     p if training else 0.0, None, 0, 7, torch.float32, rng_state=state,
                                                     want_lp=True, num_batches_tracked=nbt_a)
    # one launch
    rm_b, rv_b, nbt_b = rm.clone(), rv.clone(), torch.zeros((), dtype=torch.int64, device=cuda_device)
    z_b, y_b, mean_b, rstd_b, mask_b, ylp_b = k.linear_bn_act_fwd(x, w, bias, gamma, beta, rm_b, rv_b, 0.1, 1e-5, training, order,
                                                                 p if training else 0.0, None, 0, 7, torch.float32,
                                                                 rng_state=state, want_lp=True, num_batches_tracked=nbt_b)
    torch.cuda.synchronize()
    assert rel_err(z_b, z_a) <= 1e-5
    if training:
        assert mask_b is not None and torch.equal(mask_b, mask_a)
        assert rel_err(rm_b, rm_a) <= 1e-5 and rel_err(rv_b, rv_a) <= 1e-5
        assert int(nbt_a) == 1 and int(nbt_b) == 1
    else:
        assert int(nbt_b) == 0 and torch.equal(rm_b, rm) and torch.equal(rv_b, rv)
    # float64 statement of the block on the same bf16 operands
    zr = x.double() @ w.double().t() + bias.double()
    rm_r, rv_r = rm.double().clone(), rv.double().clone()
    h = F.relu(zr) if order == 1 else zr
    if training and M == 1:
        h = (h - h) * gamma.double() + beta.double()          # one sample: batch variance 0 (torch refuses this case)
    else:
        h = F.batch_norm(h, rm_r, rv_r, gamma.double(), beta.double(), training, 0.1, 1e-5)
    if order == 0:
        h = F.gelu(h)
    if training:
        h = h * mask_b.double() / (1 - p)
    tol = 1e-5 if M > 8 else 1e-4                             # tiny batches: 1/sqrt(var) amplifies the fp32 rounding of z
    assert rel_err(z_b, zr) <= 1e-5
    assert rel_err(y_b, h) <= tol, (rel_err(y_b, h), rel_err(y_a, h))
    assert rel_err(mean_b, mean_a) <= 1e-5 and rel_err(rstd_b, rstd_a) <= (1e-5 if M > 8 else 1e-3)
    assert ylp_b.dtype == torch.bfloat16 and torch.equal(ylp_b, y_b.to(torch.bfloat16))
    # bf16 output + an injected keep mask
    if training:
        keep = (torch.rand(M, N, generator=torch.Generator().manual_seed(8)) >= p).to(torch.uint8).to(cuda_device)
        res = k.linear_bn_act_fwd(x, w, bias, gamma, beta, rm.clone(), rv.clone(), 0.1, 1e-5, True, order, p, keep, 0, 0,
                                  torch.bfloat16)
        ref = k.bn_act_fwd(z_a, gamma, beta, rm.clone(), rv.clone(), 0.1, 1e-5, True, order, p, keep, 0, 0, torch.bfloat16)
        assert res[1].dtype == torch.bfloat16 and rel_err(res[1], ref[0]) <= 2e-2
        assert res[4] is keep
    # no bias
    z_n, y_n = k.linear_bn_act_fwd(x, w, None, gamma, beta, rm.clone(), rv.clone(), 0.1, 1e-5, training, order, 0.0, None, 0, 0,
                                   torch.float32)[:2]
    assert rel_err(z_n, zr - bias.double()) <= 1e-5


def test_linear_bn_act_shapes_outside_the_one_launch_form(cuda_device):
    k = _k()
    x = _rand((300, 128), torch.bfloat16, cuda_device, 1)
    w = _rand((64, 128), torch.bfloat16, cuda_device, 2)
    assert not k.linear_bn_act_ok(x, w)                                        # batch > 256 rows
    assert not k.linear_bn_act_ok(x[:64], _rand((12, 128), torch.bfloat16, cuda_device, 3))       # N % 8 != 0
    assert not k.linear_bn_act_ok(x[:64].float(), w.float())                   # fp32 parity mode keeps the exact kernels
    assert k.linear_bn_act_ok(x[:256], w)


def test_dropout_rng(cuda_device):
    """in-kernel Philox: keep rate ~ 1-p, deterministic for a fixed (seed, offset), scaled by 1/(1-p)."""
    k = _k()
    x = torch.ones(4096, 64, device=cuda_device)
    y1, m1 = k.dropout(x, 0.3, None, False, 123, 0, torch.float32)
    y2, m2 = k.dropout(x, 0.3, None, False, 123, 0, torch.float32)
    y3, m3 = k.dropout(x, 0.3, None, False, 124, 0, torch.float32)
    assert torch.equal(m1, m2) and not torch.equal(m1, m3)
    assert abs(float(m1.float().mean()) - 0.7) < 0.01
    assert torch.allclose(y1, m1.float() / 0.7)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("act", [2, 3])
def test_act(cuda_device, dtype, act):
    k = _k()
    x = _rand((1000, 67), torch.float32, cuda_device, 1)
    y = k.act_fwd(x, act, dtype)
    xr = x.detach().double().requires_grad_(True)
    ref = F.gelu(xr) if act == 2 else F.relu(xr)
    assert rel_err(y, ref) <= TOL[dtype]
    dy = _rand((1000, 67), torch.float32, cuda_device, 2)
    ref.backward(dy.double())
    assert rel_err(k.act_bwd(x, dy, act, dtype), xr.grad) <= TOL[dtype]


# ---------------------------------------------------------------------------------------- losses
def test_cross_entropy(cuda_device):
    k = _k()
    B, C = 257, 3
    logits = _rand((B, C), torch.float32, cuda_device, 1, 3.0)
    labels = torch.randint(0, C, (B,), generator=torch.Generator().manual_seed(2)).to(cuda_device)
    loss, pred = k.ce_fwd(logits, labels)
    lr = logits.detach().double().requires_grad_(True)
    ref = F.cross_entropy(lr, labels)
    assert rel_err(loss, ref) <= 1e-5
    assert torch.equal(pred, logits.argmax(1))
    ref.backward()
    dl = k.ce_bwd(logits, labels, torch.ones(1, device=cuda_device))
    assert rel_err(dl, lr.grad) <= 1e-5


@pytest.mark.parametrize("B", [1, 256, 20000])       # 20000: past the one-launch form (two launches, same arithmetic)
def test_cross_entropy_with_addend(cuda_device, B):
    """loss = CE + sum(addend) from the CE launch itself (the trainer's CE + w * contrastive, Trainer.py:68-71), and through
    autograd: d addend = d loss."""
    import mmsa
    k = _k()
    C = 3
    logits = _rand((B, C), torch.float32, cuda_device, 3, 3.0)
    labels = torch.randint(0, C, (B,), generator=torch.Generator().manual_seed(4)).to(cuda_device)
    add = torch.tensor([0.75, -2.5], device=cuda_device)
    loss, pred = k.ce_fwd(logits, labels, add)
    ref = F.cross_entropy(logits.double(), labels) + add.double().sum()
    assert rel_err(loss, ref) <= 1e-5
    assert torch.equal(pred, logits.argmax(1))
    lg = logits.clone().requires_grad_(True)
    extra = add.clone().requires_grad_(True)
    out = mmsa.cross_entropy(lg, labels, extra=extra)
    assert rel_err(out, ref) <= 1e-5
    (2.0 * out).backward()
    lr = logits.double().requires_grad_(True)
    (2.0 * F.cross_entropy(lr, labels)).backward()
    assert rel_err(lg.grad, lr.grad) <= 1e-5
    assert torch.equal(extra.grad, torch.full_like(add, 2.0))


def _oracle():
    from oracle import fusion_oracle as O
    return O


@pytest.mark.parametrize("B,E,same,T", [(32, 256, True, 0.01), (64, 768, False, 0.01), (48, 128, False, 0.2),
                                        (20, 256, True, 0.5)])
def test_infonce_op(cuda_device, B, E, same, T):
    """ops.infonce vs oracle (MultimodalModel.py:232-260), including dT and the max-subtraction gradient."""
    from mmsa import ops
    O = _oracle()
    g = torch.Generator().manual_seed(B)
    f1 = torch.randn(B, E, generator=g)
    f2 = f1 if same else (f1 * 0.7 + 0.3 * torch.randn(B, E, generator=g))
    labels = torch.randint(0, 3, (B,), generator=g)
    temp = torch.tensor(T)
    a = f1.clone().double().requires_grad_(True)
    b = a if same else f2.clone().double().requires_grad_(True)
    tt = temp.clone().double().requires_grad_(True)
    ref = O.infonce(a, b, labels, tt)
    ref.backward()
    x = f1.clone().to(cuda_device).requires_grad_(True)
    y = x if same else f2.clone().to(cuda_device).requires_grad_(True)
    tg = temp.clone().to(cuda_device).requires_grad_(True)
    out = ops.infonce(x, y, labels.to(cuda_device), tg)
    out.backward()
    # the reference's own fp32 evaluation: exp(s/T) at T = 0.01 amplifies the fp32 rounding of a cosine 100x, so its
    # gradients sit up to ~1e-4 from their float64 values; bar = max(1e-5, 3 x that), every excess over 1e-5 is logged
    a32 = f1.clone().requires_grad_(True)
    b32 = a32 if same else f2.clone().requires_grad_(True)
    t32 = temp.clone().requires_grad_(True)
    O.infonce(a32, b32, labels, t32).backward()
    name = f"infonce_op[B={B},E={E},same={same},T={T}]"
    assert rel_err(out, ref) <= 1e-5
    check_close(name, "d feat1", x.grad, a.grad, 1e-5, a32.grad)
    if not same:
        check_close(name, "d feat2", y.grad, b.grad, 1e-5, b32.grad)
    check_close(name, "d temperature", tg.grad, tt.grad, 1e-5, t32.grad)


@pytest.mark.parametrize("fast", [False, True])
def test_infonce_weight_folded_into_the_loss_kernels(cuda_device, fast):
    """ops.infonce(..., weight=w) == w * ops.infonce(...) (`self.contrastive_weight * loss`, MultimodalModel.py:315-317):
    value with the weight's shape, and the gradients w.r.t. both feature sets, the temperature and the weight."""
    from mmsa import ops
    B, E, T = 48, 256, 0.07
    g = torch.Generator().manual_seed(11)
    f1 = torch.randn(B, E, generator=g)
    f2 = f1 * 0.7 + 0.3 * torch.randn(B, E, generator=g)
    labels = torch.randint(0, 3, (B,), generator=g).to(cuda_device)
    outs = []
    for folded in (False, True):
        x, y = f1.clone().to(cuda_device).requires_grad_(True), f2.clone().to(cuda_device).requires_grad_(True)
        tg = torch.tensor(T, device=cuda_device, requires_grad=True)
        w = torch.tensor([1.75], device=cuda_device, requires_grad=True)
        if folded:
            out = ops.infonce(x, y, labels, tg, fast=fast, weight=w)
        else:
            out = w * ops.infonce(x, y, labels, tg, fast=fast)
        assert out.shape == (1,)
        (3.0 * out.sum()).backward()
        outs.append((out.detach(), x.grad, y.grad, tg.grad, w.grad))
    for a, b, what in zip(outs[0], outs[1], ("loss", "d feat1", "d feat2", "d temperature", "d weight")):
        assert b is not None and b.shape == a.shape, what
        assert rel_err(b, a) <= 1e-6, what
    # a weight that does not require a gradient gets none
    x = f1.clone().to(cuda_device).requires_grad_(True)
    w = torch.tensor([0.5], device=cuda_device)
    ops.infonce(x, x, labels, T, weight=w).sum().backward()
    assert x.grad is not None and w.grad is None


@pytest.mark.parametrize("B,Bg,E,T", [(64, 64, 768, 0.01), (256, 256, 768, 0.07), (128, 512, 768, 0.01), (40, 40, 128, 0.2)])
def test_infonce_split_bf16_tensor_core(cuda_device, B, Bg, E, T):
    """bf16-mode InfoNCE: the three GEMMs run on tcgen05 with split-bf16 operands (hi.hi + hi.lo + lo.hi,
    mmsa_split3).  Against the float64 oracle the loss must stay within 1e-4 and the gradients within 2e-3
    (bf16-mode bar is 2e-2); rows may be a block of a wider gathered batch (Bg > B)."""
    from mmsa import ops
    O = _oracle()
    g = torch.Generator().manual_seed(B + Bg)
    f1 = torch.randn(B, E, generator=g)
    f2 = torch.randn(Bg, E, generator=g)
    f2[:B] = f1 * 0.7 + 0.3 * f2[:B]
    lab_c = torch.randint(0, 3, (Bg,), generator=g)
    lab_r = lab_c[:B].clone()
    a, b = f1.clone().double().requires_grad_(True), f2.clone().double().requires_grad_(True)
    tt = torch.tensor(T, dtype=torch.float64, requires_grad=True)
    ref = O.infonce(a, b, lab_r, tt, labels2=lab_c, row_offset=0)
    ref.backward()
    x, y = f1.clone().to(cuda_device).requires_grad_(True), f2.clone().to(cuda_device).requires_grad_(True)
    tg = torch.tensor(T, device=cuda_device, requires_grad=True)
    out = ops.infonce(x, y, lab_r.to(cuda_device), tg, labels_cols=lab_c.to(cuda_device), row_offset=0, fast=True)
    out.backward()
    assert rel_err(out, ref) <= 1e-4
    assert rel_err(x.grad, a.grad) <= 2e-3 and rel_err(y.grad, b.grad) <= 2e-3
    assert rel_err(tg.grad, tt.grad) <= 2e-3


@pytest.mark.parametrize("B,Bg,row_offset,T", [(4096, 4096, 0, 0.01), (2048, 4096, 2048, 0.07), (1024, 8192, 3072, 0.01),
                                               (1024, 8192, 7168, 0.07), (8192, 8192, 0, 0.01)])
def test_infonce_global_batch_sizes(cuda_device, B, Bg, row_offset, T):
    """BASELINE.json configs[2] / configs[3]: InfoNCE over a GLOBAL batch of 4096 / 8192 (E_c = 768), as the whole square
    on one GPU and as the row block one rank owns (2048 of 4096 = N=2; 1024 of 8192 = N=8, global diagonal at
    row_offset + i), in both engines -- split-bf16 tcgen05 GEMMs (`fast`, the bf16-mode path that bench.py runs) and the
    exact-fp32 CUDA-core GEMMs (parity mode) -- against the float64 oracle (MultimodalModel.py:232-260): loss, d feat1,
    d feat2 (the gathered side), d temperature.  Strict bar 1e-5 for the exact engine (noise-floor clause against the
    oracle's own fp32 evaluation, logged), 2e-3 for the split-bf16 engine (bf16-mode bar is 2e-2)."""
    from mmsa import ops
    O = _oracle()
    E = 768
    g = torch.Generator().manual_seed(B + Bg + row_offset)
    f2 = torch.randn(Bg, E, generator=g)
    f1 = 0.7 * f2[row_offset:row_offset + B] + 0.3 * torch.randn(B, E, generator=g)      # positives correlate with their pair
    lab_c = torch.randint(0, 3, (Bg,), generator=g)
    lab_r = lab_c[row_offset:row_offset + B].clone()
    refs = {}
    for dt in (torch.float64, torch.float32):
        a, b = f1.clone().to(dt).requires_grad_(True), f2.clone().to(dt).requires_grad_(True)
        tt = torch.tensor(T, dtype=dt, requires_grad=True)
        loss = O.infonce(a, b, lab_r, tt, labels2=lab_c, row_offset=row_offset)
        loss.backward()
        refs[dt] = (loss.detach(), a.grad, b.grad, tt.grad)
    r64, r32 = refs[torch.float64], refs[torch.float32]
    for fast in (True, False):
        x, y = f1.clone().to(cuda_device).requires_grad_(True), f2.clone().to(cuda_device).requires_grad_(True)
        tg = torch.tensor(T, device=cuda_device, requires_grad=True)
        out = ops.infonce(x, y, lab_r.to(cuda_device), tg, labels_cols=lab_c.to(cuda_device), row_offset=row_offset, fast=fast)
        out.backward()
        name = f"infonce_global[B={B},Bg={Bg},off={row_offset},T={T},{'split-bf16 tcgen05' if fast else 'exact fp32'}]"
        got = (out, x.grad, y.grad, tg.grad)
        errs = {}
        for nm, gt, w64, w32 in zip(("loss", "d feat1", "d feat2 (gathered)", "d temperature"), got, r64, r32):
            strict = 1e-5 if (not fast or nm == "loss") else 2e-3
            if fast and nm == "loss":
                strict = 1e-4
            errs[nm] = check_close(name, nm, gt, w64, strict, w32)
        report_summary(name, 1e-5 if not fast else 2e-3, errs, note="float64 oracle; noise = the oracle's own fp32 evaluation")


def test_split3_layout(cuda_device):
    """mmsa_split3: hi + lo reproduces x to 2^-16 and the three copies land where the GEMMs expect them."""
    from mmsa import kernels as K
    x = torch.randn(37, 24, device=cuda_device)
    col, row = K.split3(x, col_side="a", row_side="b")
    hi = x.bfloat16()
    lo = (x - hi.float()).bfloat16()
    assert torch.equal(col[:, :24], hi) and torch.equal(col[:, 24:48], hi) and torch.equal(col[:, 48:], lo)
    assert torch.equal(row[:37], hi) and torch.equal(row[37:74], lo) and torch.equal(row[74:], hi)
    assert float((hi.float() + lo.float() - x).abs().max()) <= float(x.abs().max()) * 2.0 ** -16


def test_infonce_edge_cases(cuda_device):
    """rows with no positives, duplicate rows (ties in the row max), single class."""
    from mmsa import ops
    O = _oracle()
    g = torch.Generator().manual_seed(7)
    f = torch.randn(12, 64, generator=g)
    f[5] = f[2]                                   # exact duplicate -> tie on the row max
    for labels in (torch.arange(12), torch.zeros(12, dtype=torch.long), torch.tensor([0, 1, 2] * 4)):
        a = f.clone().double().requires_grad_(True)
        ref = O.infonce(a, a, labels, torch.tensor(0.05, dtype=torch.float64))
        ref.backward()
        x = f.clone().to(cuda_device).requires_grad_(True)
        out = ops.infonce(x, x, labels.to(cuda_device), 0.05)
        out.backward()
        assert rel_err(out, ref) <= 1e-5
        # labels = arange has no positives: every true gradient is < 1e-7 (the loss is flat at
        # -log(1e-12)), so the error is measured against the natural gradient scale 1/(B T |f|) ~ 0.2
        assert rel_err(x.grad, a.grad, floor=1e-3) <= 2e-4


def test_sharded_infonce_rows(cuda_device):
    """row block + global diagonal: two half-batches reproduce the full-batch loss and gradients."""
    from mmsa import ops
    O = _oracle()
    g = torch.Generator().manual_seed(11)
    B, E = 32, 128
    f1, f2 = torch.randn(B, E, generator=g), torch.randn(B, E, generator=g)
    labels = torch.randint(0, 3, (B,), generator=g)
    a, b = f1.clone().double().requires_grad_(True), f2.clone().double().requires_grad_(True)
    ref = O.infonce(a, b, labels, torch.tensor(0.07, dtype=torch.float64))
    ref.backward()
    x, y = f1.clone().to(cuda_device).requires_grad_(True), f2.clone().to(cuda_device).requires_grad_(True)
    lab = labels.to(cuda_device)
    h = B // 2
    total = 0
    for r in range(2):
        total = total + 0.5 * ops.infonce(x[r * h:(r + 1) * h], y, lab[r * h:(r + 1) * h], 0.07, labels_cols=lab,
                                          row_offset=r * h)
    total.backward()
    assert rel_err(total, ref) <= 1e-5
    assert rel_err(x.grad, a.grad) <= 1e-4 and rel_err(y.grad, b.grad) <= 1e-4


def test_supcon_ntxent(cuda_device):
    from mmsa import ops
    O = _oracle()
    g = torch.Generator().manual_seed(3)
    B, D = 24, 128
    z1, z2 = torch.randn(B, D, generator=g), torch.randn(B, D, generator=g)
    labels = torch.randint(0, 2, (B,), generator=g)
    for name in ("supcon", "ntxent"):
        a, b = z1.clone().double().requires_grad_(True), z2.clone().double().requires_grad_(True)
        ref = O.supcon(a, b, labels, 0.1) if name == "supcon" else O.ntxent(a, b, 0.5)
        ref.backward()
        x, y = z1.clone().to(cuda_device).requires_grad_(True), z2.clone().to(cuda_device).requires_grad_(True)
        out = ops.supcon(x, y, labels.to(cuda_device), 0.1) if name == "supcon" else ops.ntxent(x, y, 0.5)
        out.backward()
        assert rel_err(out, ref) <= 1e-5, name
        assert rel_err(x.grad, a.grad) <= 1e-4 and rel_err(y.grad, b.grad) <= 1e-4, name


def test_clip_adamw(cuda_device):
    """fused global-norm clip + AdamW vs torch clip_grad_norm_ + AdamW (Trainer.py:19-21,80-81)."""
    from mmsa import _lib
    n = 100_003
    p0 = _rand((n,), torch.float32, cuda_device, 1)
    g0 = _rand((n,), torch.float32, cuda_device, 2, 0.05)
    ref_p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([ref_p], lr=1e-4, weight_decay=0.01)
    p, m, v = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    partials = torch.empty(256, device=cuda_device)
    sq = torch.empty(1, device=cuda_device)
    st = torch.cuda.current_stream().cuda_stream
    for step in range(1, 4):
        grad = g0 * step
        ref_p.grad = grad.clone()
        torch.nn.utils.clip_grad_norm_([ref_p], 1.0)
        opt.step()
        _lib.call("mmsa_sumsq", grad.data_ptr(), n, partials.data_ptr(), 256, sq.data_ptr(), st)
        _lib.call("mmsa_clip_adamw", p.data_ptr(), grad.data_ptr(), m.data_ptr(), v.data_ptr(), n, sq.data_ptr(),
                  1.0, 1e-4, 0.9, 0.999, 1e-8, 0.01, step, st)
    assert rel_err(p, ref_p) <= 1e-6


def test_not_sm100_message():
    """the device gate exists and reports (on the B200 box it passes)."""
    from mmsa import _lib
    assert _lib.load().mmsa_check_device() == 0


def test_fused_clip_adamw_optimizer(cuda_device):
    """mmsa.FusedClipAdamW (flat arenas, one clip+AdamW kernel per run of parameters) vs the reference's own call pattern
    (Trainer.py:19-26,80-81): torch.optim.AdamW over the model parameters + add_param_group(the trainer's weight), and
    clip_grad_norm_ over the MODEL parameters only (the added group is neither counted in the norm nor scaled).  A parameter
    whose grad is None is skipped on both sides (no weight decay, no moment decay, no step count), an lr change goes
    through param_groups (ReduceLROnPlateau, Trainer.py:28), and the state round-trips through state_dict() in both
    directions between the two optimisers."""
    import mmsa
    g = torch.Generator().manual_seed(5)
    shapes = [(256, 768), (768,), (3, 128), (128, 128), (1,), ()]
    base = [torch.randn(s, generator=g) * 0.1 for s in shapes]
    ours = [torch.nn.Parameter(b.clone().to(cuda_device)) for b in base]
    ref = [torch.nn.Parameter(b.clone().to(cuda_device)) for b in base]
    extra_o, extra_r = torch.nn.Parameter(torch.ones(1, device=cuda_device)), torch.nn.Parameter(torch.ones(1, device=cuda_device))
    opt_o = mmsa.FusedClipAdamW(ours, lr=1e-3, weight_decay=0.01, max_norm=1.0)
    opt_r = torch.optim.AdamW(ref, lr=1e-3, weight_decay=0.01)
    opt_o.add_param_group({"params": [extra_o], "lr": 1e-3})
    opt_r.add_param_group({"params": [extra_r], "lr": 1e-3})
    assert opt_o.param_groups[0]["clip"] is True and opt_o.param_groups[1]["clip"] is False
    ids = [id(p) for p in ours]

    def one_step(step, oo, orr, po_all, pr_all, model_ref):
        for i, (po, pr) in enumerate(zip(po_all, pr_all)):
            if (i == 2 and step % 2 == 0) or (i == 3 and step < 3):     # parameters without a gradient this step
                po.grad, pr.grad = None, None
                continue
            scale = 40.0 if i == len(po_all) - 1 else 0.5 * step        # the trainer's weight: a large gradient (the loss value)
            gr = torch.randn(po.shape, generator=g).to(cuda_device) * scale
            po.grad, pr.grad = gr.clone(), gr.clone()
        torch.nn.utils.clip_grad_norm_(model_ref, 1.0)                  # Trainer.py:80: model.parameters() only
        orr.step()
        oo.step()
        for k, (po, pr) in enumerate(zip(po_all, pr_all)):
            assert rel_err(po, pr) <= 2e-6, (step, k)

    for step in range(1, 6):
        if step == 4:
            for grp in opt_o.param_groups + opt_r.param_groups:
                grp["lr"] *= 0.1
        one_step(step, opt_o, opt_r, ours + [extra_o], ref + [extra_r], ref)
    assert [id(p) for p in ours] == ids            # Parameter identity survives the re-homing into the arena
    # per-parameter step counts as torch keeps them (parameter 3 skipped two steps, parameter 2 every other one)
    sd_o, sd_r = opt_o.state_dict(), opt_r.state_dict()
    assert sorted(sd_o["state"].keys()) == sorted(sd_r["state"].keys())
    for k in sd_r["state"]:
        assert float(sd_o["state"][k]["step"]) == float(sd_r["state"][k]["step"]), k
        assert rel_err(sd_o["state"][k]["exp_avg"], sd_r["state"][k]["exp_avg"]) <= 2e-6
        assert rel_err(sd_o["state"][k]["exp_avg_sq"], sd_r["state"][k]["exp_avg_sq"]) <= 2e-6
    # resume: fresh optimisers over clones of the parameters, each loading the OTHER implementation's state_dict
    ours2 = [torch.nn.Parameter(p.detach().clone()) for p in ours + [extra_o]]
    ref2 = [torch.nn.Parameter(p.detach().clone()) for p in ours + [extra_o]]
    opt_o2 = mmsa.FusedClipAdamW(ours2[:-1], lr=1e-3, weight_decay=0.01, max_norm=1.0)
    opt_o2.add_param_group({"params": [ours2[-1]], "lr": 1e-3})
    opt_r2 = torch.optim.AdamW(ref2[:-1], lr=1e-3, weight_decay=0.01)
    opt_r2.add_param_group({"params": [ref2[-1]], "lr": 1e-3})
    opt_o2.load_state_dict(sd_r)
    opt_r2.load_state_dict(sd_o)
    assert opt_o2.param_groups[0]["lr"] == pytest.approx(1e-4) and opt_o2.param_groups[1].get("clip") is False
    for step in range(6, 9):
        one_step(step, opt_o2, opt_r2, ours2, ref2, ref2[:-1])
    # gradients that already live in one flat buffer in parameter order (GradAllReducer) are used in place
    from mmsa.optim import arena_layout
    offs, total = arena_layout(ours)
    flat = torch.zeros(total, device=cuda_device)
    for p, off in zip(ours, offs):
        p.grad = flat[off:off + p.numel()].view_as(p)
    assert opt_o._flat_grad(opt_o._arenas[0])[0].data_ptr() == flat.data_ptr()
    # a parameter moved out of its arena after construction is an error, not a silent stale update
    ours[0].data = ours[0].data.clone()
    with pytest.raises(mmsa._lib.MmsaError):
        opt_o.step()


def test_sharded_supcon_ntxent_rows(cuda_device):
    """SupCon / NT-Xent row blocks of a batch-sharded run (mmsa.dist._sharded_two_view's kernel calls), two emulated ranks in
    one process: the per-rank means average to the oracle's global loss, gradients of local rows and gathered columns add up."""
    from mmsa import ops
    from mmsa._lib import LOSS_NTXENT, LOSS_SUPCON
    O = _oracle()
    g = torch.Generator().manual_seed(9)
    Bg, D, world = 32, 128, 2
    B = Bg // world
    z1, z2 = torch.randn(Bg, D, generator=g), torch.randn(Bg, D, generator=g)
    labels = torch.randint(0, 3, (Bg,), generator=g)
    for kind, name in ((LOSS_SUPCON, "supcon"), (LOSS_NTXENT, "ntxent")):
        a, b = z1.clone().double().requires_grad_(True), z2.clone().double().requires_grad_(True)
        ref = O.supcon(a, b, labels, 0.1) if name == "supcon" else O.ntxent(a, b, 0.5)
        ref.backward()
        x, y = z1.clone().to(cuda_device).requires_grad_(True), z2.clone().to(cuda_device).requires_grad_(True)
        z_all = torch.cat([x, y])                           # what the two all-gathers + stack produce on every rank
        lab = labels.to(cuda_device)
        lab_cols = torch.cat([lab, lab]) if name == "supcon" else None
        T = 0.1 if name == "supcon" else 0.5
        total = 0
        for r in range(world):
            sl = slice(r * B, (r + 1) * B)
            lr = lab[sl] if name == "supcon" else None
            la = ops.ContrastiveFn.apply(x[sl], z_all, lr, lab_cols, None, T, kind, r * B, 2 * B, False, False, None)
            lb = ops.ContrastiveFn.apply(y[sl], z_all, lr, lab_cols, None, T, kind, Bg + r * B, 2 * B, False, False, None)
            total = total + (la + lb) / world
        total.backward()
        assert rel_err(total, ref) <= 1e-5, name
        assert rel_err(x.grad, a.grad) <= 5e-5 and rel_err(y.grad, b.grad) <= 5e-5, name


def test_sharded_two_view_losses_single_rank_group(cuda_device, tmp_path):
    """mmsa.dist.sharded_supcon / sharded_ntxent through a real (1-rank NCCL) process group == the single-GPU losses."""
    import torch.distributed as dist
    from mmsa import dist as mdist, ops
    created = False
    if not dist.is_initialized():
        dist.init_process_group("nccl", init_method=f"file://{tmp_path}/rendezvous", rank=0, world_size=1,
                                device_id=cuda_device)            # file rendezvous: no port to collide on
        created = True
    try:
        g = torch.Generator().manual_seed(10)
        z1, z2 = torch.randn(24, 128, generator=g).to(cuda_device), torch.randn(24, 128, generator=g).to(cuda_device)
        labels = torch.randint(0, 2, (24,), generator=g).to(cuda_device)
        for fn_s, fn_1, args in ((mdist.sharded_supcon, ops.supcon, (labels, 0.1)), (mdist.sharded_ntxent, ops.ntxent, (0.5,))):
            a1, a2 = z1.clone().requires_grad_(True), z2.clone().requires_grad_(True)
            b1, b2 = z1.clone().requires_grad_(True), z2.clone().requires_grad_(True)
            ls = fn_s(a1, a2, *args)
            l1 = fn_1(b1, b2, *args)
            ls.backward(); l1.backward()
            assert rel_err(ls, l1) <= 1e-6
            assert rel_err(a1.grad, b1.grad) <= 1e-5 and rel_err(a2.grad, b2.grad) <= 1e-5
    finally:
        if created:
            dist.destroy_process_group()


@pytest.mark.parametrize("D", [32, 64])
def test_attention_probability_dropout_philox_redraw(cuda_device, D):
    """mmsa_attn_dropout_fwd/bwd: the backward RE-DRAWS the Philox keep mask of the probabilities instead of storing it.
    With V = per-head identity rows the forward output IS the dropped probability matrix, which yields the mask the kernel
    drew; the Philox backward must equal the explicit-mask backward given that mask, the kept probabilities must be the
    undropped ones scaled by 1/(1-p), and the keep rate must be 1-p within 4 sigma."""
    k = _k()
    B, H, Lq, Lk, p = 4, 2, 16, 16, 0.3
    E = H * D
    q, kk = _rand((B * Lq, E), torch.float32, cuda_device, 1), _rand((B * Lk, E), torch.float32, cuda_device, 2)
    v = torch.zeros(B * Lk, E, device=cuda_device)
    for j in range(Lk):
        for h in range(H):
            v[j::Lk, h * D + j] = 1.0
    state = torch.tensor([99, 5], dtype=torch.int64, device=cuda_device)
    o, lse = k.attn_dropout_fwd(q, kk, v, B, H, Lq, Lk, D, p, None, 0, 7, state)
    o0, _ = k.attn_dropout_fwd(q, kk, v, B, H, Lq, Lk, D, 0.0, None, 0, 0, None)
    o_val, _ = k.attn_dropout_fwd(q, kk, v, B, H, Lq, Lk, D, p, None, 99, 12, None)      # by-value (seed, offset) = state + 7
    assert torch.equal(o, o_val)
    Pd = o.view(B, Lq, H, D)[..., :Lk].permute(0, 2, 1, 3)
    P0 = o0.view(B, Lq, H, D)[..., :Lk].permute(0, 2, 1, 3)
    mask = (Pd != 0).to(torch.uint8).contiguous()
    n = mask.numel()
    assert abs(float(mask.float().mean()) - (1 - p)) <= 4 * (p * (1 - p) / n) ** 0.5
    assert float((Pd - P0 * mask / (1 - p)).abs().max()) <= 1e-6
    do = _rand((B * Lq, E), torch.float32, cuda_device, 3)
    got = [torch.empty_like(t) for t in (q, kk, v)]
    want = [torch.empty_like(t) for t in (q, kk, v)]
    k.attn_dropout_bwd(q, kk, v, o, do, lse, B, H, Lq, Lk, D, *got, p, None, 0, 7, state)
    k.attn_dropout_bwd(q, kk, v, o, do, lse, B, H, Lq, Lk, D, *want, p, mask, 0, 0, None)
    for a, b in zip(got, want):
        assert torch.equal(a, b)


def test_wgrad_a_from_tmem_probe(cuda_device):
    """MMSA_WGRAD_A_TMEM=1 (probe, off by default): the 2-CTA weight gradient with its A tile transposed into tensor memory by
    the epilogue warps (tcgen05.st) and A-from-TMEM MMAs.  The switch is read once per process, so the check runs in a child
    process: result against a float64 product, incl. the bias gradient of the ones-tile MMA and a ragged token count."""
    import subprocess
    import sys
    code = r"""
import sys, torch
sys.path.insert(0, %r)
from mmsa import kernels as K
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(3)
for (T, N, Kd) in ((4100, 768, 768), (8192, 256, 520)):
    dy = torch.randn(T, N, generator=g).to(dev).bfloat16(); x = torch.randn(T, Kd, generator=g).to(dev).bfloat16()
    dw, db = K.linear_wgrad(dy, x)
    rw = dy.double().T @ x.double(); rb = dy.double().sum(0)
    ew = float((dw.double() - rw).abs().max() / rw.abs().max()); eb = float((db.double() - rb).abs().max() / rb.abs().max())
    assert ew <= 2e-3 and eb <= 2e-3, (T, N, Kd, ew, eb)
print("ok")
""" % os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "multimodal-sentiment-aanalysis_b200")
    env = dict(os.environ, MMSA_WGRAD_A_TMEM="1")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("M,N,K1,K2", [(4352, 768, 768, 768),      # one 2-CTA launch, B operand from two tensor maps
                                       (4100, 256, 512, 256),      # ragged tokens, unequal segments (split on a tile boundary)
                                       (300, 768, 768, 768),       # too few tokens for the pair path: two launches on dw's column blocks
                                       (4352, 768, 264, 768)])     # first segment not a multiple of the tile width: fallback
def test_linear_wgrad_two_inputs(cuda_device, dtype, M, N, K1, K2):
    """mmsa_linear_wgrad2: dW = dy^T [x | x2] without materialising the concat (the gate's Linear(2E, E), MultimodalModel.py:147)."""
    k = _k()
    dy = _rand((M, N), dtype, cuda_device, 1)
    x = _rand((M, K1), dtype, cuda_device, 2)
    x2 = _rand((M, K2), dtype, cuda_device, 3)
    dw, db = k.linear_wgrad2(dy, x, x2)
    ref_w = dy.double().T @ torch.cat([x, x2], 1).double()
    tol = 1e-5 if dtype == torch.float32 else 2e-3
    assert rel_err(dw, ref_w) <= tol
    assert rel_err(db, dy.double().sum(0)) <= tol
    dw2, _ = k.linear_wgrad2(dy, x, x2, want_bias=False)
    assert torch.equal(dw2, dw)
