"""Probe: tcgen05 main-loop rate per operand-major combination at one L2-resident GEMM size
(8192 x 3072 x K, fp32 output through the TMA epilogue), via mmsa_debug_gemm."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-sentiment-aanalysis_b200"))
import torch
from mmsa import _lib
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
GM, GN = 8192, 3072
for GK in (1024, 4096):
    for bn in (192, 256):
        C = torch.empty(GM, GN, device=dev, dtype=torch.float32)
        for a_mn in (0, 1):
            for b_mn in (0, 1):
                A = torch.randn((GK, GM) if a_mn else (GM, GK), device=dev).bfloat16()
                B = torch.randn((GK, GN) if b_mn else (GN, GK), device=dev).bfloat16()
                def run():
                    _lib.call("mmsa_debug_gemm", a_mn, b_mn, GM, GN, GK, A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0),
                              C.data_ptr(), C.stride(0), 1, bn, st)
                for _ in range(3): run()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10): run()
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 10
                print(f"K={GK} BN={bn:3d} A={'MN' if a_mn else 'K '} B={'MN' if b_mn else 'K '}: {ms*1e3:8.1f} us  {2.0*GM*GN*GK/ms/1e9:7.0f} TF/s", flush=True)
