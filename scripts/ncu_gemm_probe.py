"""A handful of representative GEMM launches for an `ncu --set full` capture (GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-sentiment-aanalysis_b200"))
import torch
from mmsa import kernels as K
dev = torch.device("cuda:0")
M, E = 32768, 768
x = torch.randn(M, E, device=dev).bfloat16()
dy = torch.randn(M, E, device=dev).bfloat16()
w = (torch.randn(E, E, device=dev) / E ** 0.5).bfloat16()
b = torch.randn(E, device=dev)
for _ in range(2):
    K.linear_fwd(x, w, b)            # 2-CTA, K-major B
    K.linear_dgrad(dy, w)            # 2-CTA, MN-major B
    K.linear_wgrad(dy, x)            # split-K cluster, both MN-major
torch.cuda.synchronize()
print("ok")
