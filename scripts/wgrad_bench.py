"""Probe (GPU box): weight-gradient GEMMs of the configs[1] step through the C ABI (mmsa_linear_wgrad), rotating operand sets
larger than L2; MMSA_WGRAD_PAIR_OFF=1 selects the 1-CTA cluster split-K path for an A/B."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-sentiment-aanalysis_b200"))
import torch
from mmsa import kernels as K
dev = torch.device("cuda:0")
for (T, N, Kd) in [(32768, 768, 768), (12544, 768, 768), (32768, 1536, 768), (12544, 1536, 768), (12544, 768, 2048)]:
    nset = max(2, int(300e6 // ((T * N + T * Kd) * 2)) + 1)
    dys = [torch.randn(T, N, device=dev).bfloat16() for _ in range(nset)]
    xs = [torch.randn(T, Kd, device=dev).bfloat16() for _ in range(nset)]
    dw = torch.empty(N, Kd, device=dev); db = torch.empty(N, device=dev)
    for i in range(3): K.linear_wgrad(dys[i % nset], xs[i % nset], dw=dw, db=db)
    torch.cuda.synchronize()
    ref = dys[2 % nset].float().T @ xs[2 % nset].float()
    err = float((dw - ref).abs().max() / ref.abs().max())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20): K.linear_wgrad(dys[i % nset], xs[i % nset], dw=dw, db=db)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"wgrad tokens={T:6d} dW[{N}x{Kd}]  {ms*1e3:7.1f} us  {2.0*T*N*Kd/ms/1e9:6.0f} TF/s  err={err:.1e}", flush=True)
    del dys, xs
