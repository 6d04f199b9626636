"""In-graph latency of the [B,*] tail products on the two bf16 engines (gemm_small.cu vs the tcgen05 kernel):
20 dependent launches of one shape captured in a CUDA graph, replayed; prints us per launch.

    python scripts/small_gemm_probe.py"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "multimodal-sentiment-aanalysis_b200"))
from mmsa import _lib, kernels as K   # noqa: E402

lib = _lib.load()
dev = torch.device("cuda:0")
SHAPES = [("fwd", 256, 256, 2304), ("fwd", 256, 128, 256), ("fwd", 256, 64, 1536), ("fwd", 256, 128, 128),
          ("dgrad", 256, 256, 2304), ("dgrad", 256, 64, 768), ("wgrad", 256, 256, 2304), ("wgrad", 256, 64, 768),
          ("fwd", 256, 256, 768)]
REP = 20


def bench(kind, M, N, Kd):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(M, Kd, generator=g).to(dev).bfloat16()
    w = (torch.randn(N, Kd, generator=g) / math.sqrt(Kd)).to(dev).bfloat16()
    dy = torch.randn(M, N, generator=g).to(dev).bfloat16()
    b = torch.zeros(N, device=dev)
    if kind == "fwd":
        out = torch.empty(M, N, device=dev)
        fn = lambda: K.linear_fwd(x, w, b, out_dtype=torch.float32, out=out)
    elif kind == "dgrad":
        out = torch.empty(M, Kd, device=dev)
        fn = lambda: K.linear_dgrad(dy, w, out_dtype=torch.float32, out=out)
    else:
        dw = torch.empty(N, Kd, device=dev)
        fn = lambda: K.linear_wgrad(dy, x, dw=dw)
    res = {}
    for eng in (0, 1):
        lib.mmsa_debug_gemm_engine(eng)
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            for _ in range(3):
                fn()
            s.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=s):
                for _ in range(REP):
                    fn()
            for _ in range(3):
                gr.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            for _ in range(10):
                gr.replay()
            e1.record(s)
            s.synchronize()
        res[eng] = e0.elapsed_time(e1) * 1e3 / (10 * REP)
    lib.mmsa_debug_gemm_engine(0)
    print(f"{kind:6s} M={M:4d} N={N:4d} K={Kd:5d}   small {res[0]:6.2f} us   tcgen05 {res[1]:6.2f} us")


for sh in SHAPES:
    bench(*sh)
