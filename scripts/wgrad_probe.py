import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-sentiment-aanalysis_b200"))
import torch
from mmsa import kernels as K
dev = torch.device("cuda:0")
M, N, Kd = 32768, 768, 768
bias = os.environ.get("BIAS", "1") == "1"
xs = [torch.randn(M, Kd, device=dev).bfloat16() for _ in range(3)]
dys = [torch.randn(M, N, device=dev).bfloat16() for _ in range(3)]
def run(i): K.linear_wgrad(dys[i % 3], xs[i % 3], want_bias=bias)
for i in range(3): run(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(20): run(i)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"splits={os.environ.get('MMSA_WGRAD_SPLITS')} bn={os.environ.get('MMSA_WGRAD_BN')} bias={bias}: {ms*1e3:.1f} us {2.0*M*N*Kd/ms/1e9:.0f} TF/s")
