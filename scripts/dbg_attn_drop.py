"""Diagnostic (GPU box): mmsa_attn_dropout_fwd/bwd -- extract the Philox mask through V = identity rows and compare the
re-drawn-mask backward with the explicit-mask backward."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-sentiment-aanalysis_b200"))
import torch
from mmsa import kernels as K
dev = torch.device("cuda:0")
torch.manual_seed(0)
for D in (32, 64):
    B, H, Lq, Lk, p = 3, 2, 9, 9, 0.3
    E = H * D
    q = torch.randn(B * Lq, E, device=dev); k = torch.randn(B * Lk, E, device=dev)
    # V = per-head identity rows: o[i, d] = P'[i, d] for d < Lk
    v = torch.zeros(B * Lk, E, device=dev)
    for b in range(B):
        for j in range(Lk):
            for h in range(H):
                v[b * Lk + j, h * D + j] = 1.0
    state = torch.tensor([99, 5], dtype=torch.int64, device=dev)
    o, lse = K.attn_dropout_fwd(q, k, v, B, H, Lq, Lk, D, p, None, 0, 7, state)
    o0, lse0 = K.attn_dropout_fwd(q, k, v, B, H, Lq, Lk, D, 0.0, None, 0, 0, None)
    Pd = o.view(B, Lq, H, D)[..., :Lk].permute(0, 2, 1, 3)       # [B,H,Lq,Lk] dropped probs
    P0 = o0.view(B, Lq, H, D)[..., :Lk].permute(0, 2, 1, 3)
    mask = (Pd != 0).to(torch.uint8).contiguous()
    print("D", D, "keep rate", float(mask.float().mean()), "scale ok", float((Pd - P0 * mask / (1 - p)).abs().max()))
    do = torch.randn(B * Lq, E, device=dev)
    outs = []
    for mode in ("philox", "explicit"):
        dq, dk, dv = (torch.empty_like(q), torch.empty_like(k), torch.empty_like(v))
        if mode == "philox":
            K.attn_dropout_bwd(q, k, v, o, do, lse, B, H, Lq, Lk, D, dq, dk, dv, p, None, 0, 7, state)
        else:
            K.attn_dropout_bwd(q, k, v, o, do, lse, B, H, Lq, Lk, D, dq, dk, dv, p, mask, 0, 0, None)
        outs.append((dq, dk, dv))
    for a, b, n in zip(outs[0], outs[1], ("dq", "dk", "dv")):
        print("  ", n, float((a - b).abs().max()), float(b.abs().max()))
