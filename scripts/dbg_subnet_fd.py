"""Diagnostic (GPU box): finite-difference check of mmsa.Subnetwork in train mode with the Philox stream reset before every
forward, per dropout kind (attention probabilities / token dropouts)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-sentiment-aanalysis_b200"))
import torch, mmsa
dev = torch.device("cuda:0")
torch.manual_seed(7)
m = mmsa.Subnetwork(64, feat_dim=64, num_layers=1, nhead=2).to(dev).train()
x = torch.randn(6, 9, 64, device=dev)
wgt = torch.randn(6, 9, 64, device=dev)
def run(pa, pt):
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout): mod.p = pt
        if isinstance(mod, torch.nn.MultiheadAttention): mod.dropout = pa
    def f(xx):
        m._drop.reseed(99, position=5)
        return (m(xx) * wgt).sum()
    xg = x.clone().requires_grad_(True)
    f(xg).backward()
    d = torch.randn_like(x); d /= d.norm()
    an = float((xg.grad.double() * d.double()).sum())
    for eps in (1e-1, 3e-2, 1e-2, 3e-3):
        with torch.no_grad():
            fd = (f(x + eps * d).double() - f(x - eps * d).double()) / (2 * eps)
        print(f"pa={pa} pt={pt} eps={eps}: fd={float(fd):+.5f} an={an:+.5f}")
run(0.0, 0.0); run(0.3, 0.0); run(0.0, 0.3); run(0.3, 0.3)
