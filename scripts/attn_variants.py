"""Probe: forward tcgen05 attention variants (3-D vs flat load maps, TMA vs direct stores)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-sentiment-aanalysis_b200"))
import torch
from mmsa import kernels as K, _lib
dev = torch.device("cuda:0"); lib = _lib.load()
B, H, D = 256, 12, 64; E = H * D
for (Lq, Lk) in [(128, 49), (49, 128)]:
    q = torch.randn(B * Lq, E, device=dev).bfloat16(); kv = torch.randn(B * Lk, 2 * E, device=dev).bfloat16()
    line = f"Lq={Lq} Lk={Lk}:"
    for dbg, name in ((0, "3d+tma"), (1, "flat+tma"), (2, "3d+direct"), (3, "flat+direct")):
        lib.mmsa_debug_attention_engine(dbg << 8)
        for _ in range(3): K.attn_fwd(q, kv[:, :E], kv[:, E:], B, H, Lq, Lk, D)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): K.attn_fwd(q, kv[:, :E], kv[:, E:], B, H, Lq, Lk, D)
        e1.record(); torch.cuda.synchronize()
        line += f"  {name} {e0.elapsed_time(e1)/20*1e3:6.1f} us"
    print(line, flush=True)
lib.mmsa_debug_attention_engine(0)
