"""Probe (GPU box): tcgen05/TMA attention forward + fused backward vs a float64 reference and the mma.sync engine."""
import os, sys, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-sentiment-aanalysis_b200"))
import torch
from mmsa import kernels as K, _lib
dev = torch.device("cuda:0")
lib = _lib.load()

def ref(q, k, v, do, B, H, Lq, Lk, D):
    q, k, v = (t.double().clone().requires_grad_(True) for t in (q, k, v))
    qd = q.view(B, Lq, H, D).transpose(1, 2); kd = k.view(B, Lk, H, D).transpose(1, 2); vd = v.view(B, Lk, H, D).transpose(1, 2)
    s = (qd / math.sqrt(D)) @ kd.transpose(-1, -2)
    o = (torch.softmax(s, -1) @ vd).transpose(1, 2).reshape(B * Lq, H * D)
    o.backward(do.double())
    return o.detach(), torch.logsumexp(s, -1).detach(), q.grad, k.grad, v.grad

def rel(a, b): return float((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-30))

shapes = [(4, 12, 128, 49), (4, 12, 49, 128), (2, 12, 512, 49), (2, 12, 49, 512), (3, 12, 64, 49), (2, 4, 100, 130), (1, 1, 1, 1), (2, 2, 200, 130)]
for (B, H, Lq, Lk) in shapes:
    D = 64; E = H * D
    g = torch.Generator().manual_seed(Lq * 7 + Lk)
    q = torch.randn(B * Lq, E, generator=g).to(dev).bfloat16()
    kv = torch.randn(B * Lk, 2 * E, generator=g).to(dev).bfloat16()
    do = torch.randn(B * Lq, E, generator=g).to(dev).bfloat16()
    k, v = kv[:, :E], kv[:, E:]
    o_ref, lse_ref, dq_ref, dk_ref, dv_ref = ref(q, k, v, do, B, H, Lq, Lk, D)
    line = f"B={B} H={H} Lq={Lq} Lk={Lk}:"
    for eng in (1, 0):
        lib.mmsa_debug_attention_engine(eng)
        o, lse = K.attn_fwd(q, k, v, B, H, Lq, Lk, D)
        dq = torch.full_like(q, float("nan")); dkv = torch.full_like(kv, float("nan"))
        K.attn_bwd(q, k, v, o, do, lse, B, H, Lq, Lk, D, dq, dkv[:, :E], dkv[:, E:])
        torch.cuda.synchronize()
        line += (f" | {'tc ' if eng == 0 else 'mma'} o={rel(o, o_ref):.1e} lse={float((lse.double() - lse_ref).abs().max()):.1e}"
                 f" dq={rel(dq, dq_ref):.1e} dk={rel(dkv[:, :E], dk_ref):.1e} dv={rel(dkv[:, E:], dv_ref):.1e}")
    print(line, flush=True)

B, H, D = 256, 12, 64; E = H * D
for (Lq, Lk) in [(128, 49), (49, 128), (512, 49), (49, 512)]:
    Bb = B if max(Lq, Lk) <= 128 else 128
    q = torch.randn(Bb * Lq, E, device=dev).bfloat16(); do = torch.randn(Bb * Lq, E, device=dev).bfloat16()
    kv = torch.randn(Bb * Lk, 2 * E, device=dev).bfloat16()
    dq = torch.empty_like(q); dkv = torch.empty_like(kv)
    line = f"B={Bb} Lq={Lq} Lk={Lk}:"
    for eng in (1, 0):
        lib.mmsa_debug_attention_engine(eng)
        o, lse = K.attn_fwd(q, kv[:, :E], kv[:, E:], Bb, H, Lq, Lk, D)
        def fwd(): K.attn_fwd(q, kv[:, :E], kv[:, E:], Bb, H, Lq, Lk, D)
        def bwd(): K.attn_bwd(q, kv[:, :E], kv[:, E:], o, do, lse, Bb, H, Lq, Lk, D, dq, dkv[:, :E], dkv[:, E:])
        for name, fn, nbytes in (("fwd", fwd, 2.0 * 64 * Bb * H * (2.0 * Lq + 2.0 * Lk)), ("bwd", bwd, 2.0 * 64 * Bb * H * (4.0 * Lq + 4.0 * Lk))):
            for _ in range(3): fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): fn()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            line += f"  {'tc ' if eng == 0 else 'mma'} {name} {ms*1e3:6.1f} us {nbytes/ms/1e6:5.0f} GB/s"
    print(line, flush=True)
lib.mmsa_debug_attention_engine(0)
