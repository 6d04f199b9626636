"""Probe: tcgen05 main-loop rate per operand-major combination and tile width, one CTA per tile, long K
(so the epilogue is negligible): C[M,N] fp32 = A B^T, M = 148*128 rows (one tile per SM), K = 8192."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-sentiment-aanalysis_b200"))
import torch
from mmsa import _lib
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
GK = 8192
for bn in (64, 128, 192, 256):
    GM, GN = 148 * 128, bn
    C = torch.empty(GM, GN, device=dev, dtype=torch.float32)
    for a_mn in (0, 1):
        for b_mn in (0, 1):
            A = torch.randn((GK, GM) if a_mn else (GM, GK), device=dev).bfloat16()
            B = torch.randn((GK, GN) if b_mn else (GN, GK), device=dev).bfloat16()
            def run():
                _lib.call("mmsa_debug_gemm", a_mn, b_mn, GM, GN, GK, A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0),
                          C.data_ptr(), C.stride(0), 1, bn, st)
            for _ in range(2): run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): run()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            kb = GK // 64
            ref = (A.float().t() if a_mn else A.float())[:256] @ (B.float() if b_mn else B.float().t())
            err = float((C[:256] - ref).abs().max() / ref.abs().max())
            print(f"BN={bn:3d} A={'MN' if a_mn else 'K '} B={'MN' if b_mn else 'K '}: {ms*1e3:8.1f} us  {2.0*GM*GN*GK/ms/1e9:7.0f} TF/s  "
                  f"{ms*1e6/kb*1.9:6.0f} cyc/k-block (peak {2*bn})  err={err:.1e}", flush=True)
