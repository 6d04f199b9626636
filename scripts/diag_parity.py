"""Diagnostic (GPU box): per-tensor error of the CUDA path against the fp64 oracle, beside the fp32
oracle's own error against fp64 (the noise floor of ANY fp32 implementation of the same maths)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import torch
from parity_util import O, build_model, oracle_step, rel_err, grad_err, zero_grad_bias_keys
import mmsa

def run(batch, L, dtype, temp=None):
    cd = torch.float32 if dtype == "fp32" else torch.bfloat16
    cfg = O.FusionConfig(embed_dim=768, num_heads=12, wiring="bidirectional", contract="single", valence=False)
    params, buffers = O.init_params(cfg, seed=0)
    if temp is not None:
        params["temperature"] = torch.tensor(temp)
    inputs, labels = O.synth_inputs(cfg, batch, L=L, R=49, seed=1234)
    l64, o64, g64 = oracle_step(cfg, params, inputs, labels, dtype=torch.float64)
    l32, o32, g32 = oracle_step(cfg, params, inputs, labels, dtype=torch.float32)
    model = build_model(cfg, params, buffers, cd, "cuda:0")
    text, image = (x.cuda() for x in inputs)
    lab = labels.cuda()
    logits, closs = model(text, image, None, lab)
    loss = mmsa.cross_entropy(logits, lab) + closs.sum()
    loss.backward()
    zk = zero_grad_bias_keys(g64.keys())
    rows = [("logits", rel_err(logits, o64.arousal), rel_err(o32.arousal, o64.arousal)),
            ("loss", rel_err(loss, l64), rel_err(l32, l64))]
    for k, p in model.named_parameters():
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        rows.append(("grad:" + k, grad_err(k, g, g64[k], g64, zk), grad_err(k, g32[k], g64[k], g64, zk)))
    print(f"== B={batch} L={L} {dtype} T={temp}  loss={float(loss):.6f} oracle64={float(l64):.6f}")
    for n, a, b in sorted(rows, key=lambda r: -r[1])[:12]:
        print(f"   {n:55s} ours={a:.3e}  oracle_fp32={b:.3e}")

if __name__ == "__main__":
    for dtype in ("fp32", "bf16"):
        for (b, L) in ((4, 64), (8, 128)):
            for T in (None, 0.07):
                run(b, L, dtype, T)
