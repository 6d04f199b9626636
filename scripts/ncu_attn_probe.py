"""A few attention launches (tcgen05 engines) for an ncu capture (GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-sentiment-aanalysis_b200"))
import torch
from mmsa import kernels as K
dev = torch.device("cuda:0")
B, H, D = 256, 12, 64; E = H * D
for (Lq, Lk) in [(128, 49), (49, 128)]:
    q = torch.randn(B * Lq, E, device=dev).bfloat16(); do = torch.randn(B * Lq, E, device=dev).bfloat16()
    kv = torch.randn(B * Lk, 2 * E, device=dev).bfloat16()
    dq = torch.empty_like(q); dkv = torch.empty_like(kv)
    for _ in range(2):
        o, lse = K.attn_fwd(q, kv[:, :E], kv[:, E:], B, H, Lq, Lk, D)
        K.attn_bwd(q, kv[:, :E], kv[:, E:], o, do, lse, B, H, Lq, Lk, D, dq, dkv[:, :E], dkv[:, E:])
torch.cuda.synchronize()
print("ok")
