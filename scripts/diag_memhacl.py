import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import torch
from parity_util import O, rel_err
import mmsa
g = torch.load(os.path.join(ROOT, "tests", "golden", "memhacl.pt"))
dev = "cuda:0"
feats = [f.to(dev) for f in g["feats"]]
enc2 = mmsa.MultiModalEncoder(variant="max")
enc2.load_state_dict(g["max"]["state_dict"], strict=True)
enc2 = enc2.to(dev).train()
xs = [f.clone().requires_grad_(True) for f in feats]
y2 = enc2(*xs)
y2.square().sum().backward()
print("out", rel_err(y2, g["max"]["out"]))
for i, (x, d) in enumerate(zip(xs, g["max"]["dfeats"])):
    print("dfeat", i, rel_err(x.grad, d))
# float64 oracle of the same thing
p = {k: v.double() for k, v in g["max"]["state_dict"].items() if v.is_floating_point()}
xs64 = [f.double().cpu().clone().requires_grad_(True) for f in g["feats"]]
y64 = O.memhacl_fusion(xs64, p, num_heads=8, variant="max", training=True)
y64.square().sum().backward()
print("vs fp64: out ours", rel_err(y2, y64), "golden32", rel_err(g["max"]["out"], y64))
for i in range(3):
    print("vs fp64: dfeat", i, "ours", rel_err(xs[i].grad, xs64[i].grad), "golden32", rel_err(g["max"]["dfeats"][i], xs64[i].grad))
