"""Probe (GPU box): wgrad-shaped GEMMs (both operands MN-major, split-K clusters): 1-CTA 128x256 tiles vs 2-CTA 256x256
tiles (bn = 512 in mmsa_debug_gemm) over the split counts."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-sentiment-aanalysis_b200"))
import torch
from mmsa import _lib
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
# (Nw = dY cols, Kw = X cols, rows)
for (Nw, Kw, rows) in [(768, 768, 32768), (768, 768, 12544), (1536, 768, 32768), (1536, 768, 12544), (768, 2048, 12544)]:
    nset = 3
    dYs = [torch.randn(rows, Nw, device=dev).bfloat16() for _ in range(nset)]
    Xs = [torch.randn(rows, Kw, device=dev).bfloat16() for _ in range(nset)]
    C = torch.zeros(Nw, Kw, device=dev)
    ref = dYs[0].float().t()[:256] @ Xs[0].float()
    line = f"dW[{Nw}x{Kw}] rows={rows}:"
    for bn, sp_list in ((256, (6,)), (128, (2, 3, 4)), (192, (4, 6)), (64, (2,))):
        for sp in sp_list:
            def run(i):
                j = i % nset
                _lib.call("mmsa_debug_gemm", 1, 1, Nw, Kw, rows, dYs[j].data_ptr(), dYs[j].stride(0), Xs[j].data_ptr(), Xs[j].stride(0),
                          C.data_ptr(), C.stride(0), sp, bn, st)
            run(0); torch.cuda.synchronize()
            err = float((C[:256] - ref).abs().max() / ref.abs().max())
            for i in range(3): run(i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(12): run(i)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 12
            line += f"  bn={bn} S={sp} {ms*1e3:5.1f}us {2.0*Nw*Kw*rows/ms/1e9:5.0f}TF e={err:.0e}"
    print(line, flush=True)
