"""Probe (GPU box): 2-CTA (cta_group::2, 256x256 tiles) vs 1-CTA tcgen05 GEMM on the big activation shapes,
through mmsa_debug_gemm (bn = 0: planner's choice, pair when eligible; bn = -1: force the 1-CTA kernel)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-sentiment-aanalysis_b200"))
import torch
from mmsa import _lib
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
shapes = [(32768, 768, 768), (32768, 1536, 768), (32768, 768, 1536), (12544, 768, 768), (12544, 768, 2048), (4352, 768, 768), (5000, 1536, 264)]
for (M, N, Kd) in shapes:
    for b_mn in (0, 1):
        nset = max(2, int(300e6 // ((M * Kd + M * N * 2) * 2)) + 1)
        As = [torch.randn(M, Kd, device=dev).bfloat16() for _ in range(nset)]
        Bm = (torch.randn((Kd, N) if b_mn else (N, Kd), device=dev) / Kd ** 0.5).bfloat16()
        Cs = [torch.zeros(M, N, device=dev, dtype=torch.float32) for _ in range(nset)]
        ref = (As[0][:512].float() @ (Bm.float() if b_mn else Bm.float().t()))
        ref_tail = (As[0][-300:].float() @ (Bm.float() if b_mn else Bm.float().t()))
        line = f"M={M:6d} N={N:5d} K={Kd:5d} B={'MN' if b_mn else 'K '}:"
        for bn in (-1, 0):
            def run(i):
                j = i % nset
                _lib.call("mmsa_debug_gemm", 0, b_mn, M, N, Kd, As[j].data_ptr(), As[j].stride(0), Bm.data_ptr(), Bm.stride(0),
                          Cs[j].data_ptr(), Cs[j].stride(0), 1, bn, st)
            for i in range(nset): run(i)
            torch.cuda.synchronize()
            err = max(float((Cs[0][:512] - ref).abs().max()), float((Cs[0][-300:] - ref_tail).abs().max())) / float(ref.abs().max())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(20): run(i)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            line += f"  {'pair ' if bn == 0 else '1-cta'} {ms*1e3:7.1f} us {2.0*M*N*Kd/ms/1e9:6.0f} TF/s err={err:.1e}"
            for c in Cs: c.zero_()
        print(line, flush=True)
        del As, Cs
