"""Probe (GPU box): intrinsic main-loop rate of one CTA (pair) on a lightly loaded chip: 9 output tiles, K = 32768,
for every operand-major combination; reports cycles per 64-deep k-block at 1.965 GHz (ideal: 512 for a 128x256
tile on one SM or a 256x256 tile on a CTA pair)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-sentiment-aanalysis_b200"))
import torch
from mmsa import _lib
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
GK = 32768
for pair in (0, 1):
    GM = 768 if pair else 384
    GN = 768
    for a_mn in (0, 1):
        for b_mn in (0, 1):
            A = torch.randn((GK, GM) if a_mn else (GM, GK), device=dev).bfloat16()
            B = torch.randn((GK, GN) if b_mn else (GN, GK), device=dev).bfloat16()
            C = torch.empty(GM, GN, device=dev)
            def run():
                _lib.call("mmsa_debug_gemm", a_mn, b_mn, GM, GN, GK, A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0),
                          C.data_ptr(), C.stride(0), 1, 512 if pair else 256, st)
            for _ in range(2): run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): run()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print(f"{'pair 256x256' if pair else '1cta 128x256'} A={'MN' if a_mn else 'K '} B={'MN' if b_mn else 'K '}: {ms*1e3:7.1f} us  "
                  f"{ms*1e-3/(GK/64)*1.965e9:6.0f} cyc/k-block", flush=True)
