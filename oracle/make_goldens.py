"""Generate tests/golden/*.pt from the IMPORTED REFERENCE (run here, where /root/reference is mounted).

The reference has no golden vectors of its own, so these fixtures are the pin: outputs and
gradients of the reference's own MultimodalTransformerModel / MultiModalEncoder / ProjectionHead /
Classifier classes at their native sizes, on seeded inputs, with the out-of-scope modality encoders
replaced by nn.Identity() (features in) and dropout p set to 0 (RNG streams cannot be matched).

    python oracle/make_goldens.py        # rewrites tests/golden/

Gradients are stored as (L2 norm, sum, 64 sampled elements) per tensor to keep the fixture small."""
from __future__ import annotations

import importlib.util
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/MML_ZYC"
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)


def grad_digest(g: torch.Tensor):
    flat = g.detach().reshape(-1)
    n = flat.numel()
    idx = torch.linspace(0, n - 1, steps=min(64, n)).long()
    return {"norm": float(flat.double().norm()), "sum": float(flat.double().sum()), "idx": idx, "vals": flat[idx].clone(),
            "absmax": float(flat.abs().max())}


def _zero_dropout(m: nn.Module):
    for mod in m.modules():
        if isinstance(mod, nn.Dropout):
            mod.p = 0.0


def native_fusion():
    sys.path.insert(0, REF)
    import MultimodalModel as R
    from oracle import fusion_oracle as O
    torch.manual_seed(0)
    m = R.MultimodalTransformerModel()
    m.eeg_net, m.eye_net, m.pps_net = nn.Identity(), nn.Identity(), nn.Identity()
    _zero_dropout(m)
    m.train()
    # move the temperature off its init so InfoNCE is not in the pure-underflow regime for one case
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    torch.save({"state_dict": sd0, "keys": list(sd0.keys())}, os.path.join(OUT, "native_params.pt"))
    cfg = O.FusionConfig()
    for B, temp in ((2, 0.01), (20, 0.01), (20, 0.5), (64, 0.07)):
        m.load_state_dict(sd0)
        with torch.no_grad():
            m.temperature.fill_(temp)
        xs, labels = O.synth_inputs(cfg, B, seed=100 + B)
        val_labels = (labels + 1) % 3
        m.zero_grad()
        a, v, c0, c1, c2 = m(*xs, labels=(labels, val_labels))
        loss = F.cross_entropy(a, labels) + F.cross_entropy(v, val_labels) + c0.sum() + c1.sum() + c2.sum()
        loss.backward()
        case = {
            "B": B, "temperature": temp, "inputs": [x.clone() for x in xs], "labels": labels, "val_labels": val_labels,
            "arousal": a.detach().clone(), "valence": v.detach().clone(),
            "contrastive": [c.detach().clone() for c in (c0, c1, c2)], "loss": loss.detach().clone(),
            "grads": {k: grad_digest(p.grad) for k, p in m.named_parameters() if p.grad is not None},
            "buffers_after": {k: b.clone() for k, b in m.named_buffers()},
        }
        # the same step evaluated by the reference in float64 (m.double()): the measure of the fp32
        # reference's OWN rounding error, which bounds what any fp32 implementation can match
        import copy
        m64 = copy.deepcopy(m).double()
        m64.load_state_dict({k: (v.double() if v.is_floating_point() else v) for k, v in sd0.items()})
        with torch.no_grad():
            m64.temperature.fill_(temp)
        m64.zero_grad()
        a64, v64, d0, d1, d2 = m64(*[x.double() for x in xs], labels=(labels, val_labels))
        loss64 = F.cross_entropy(a64, labels) + F.cross_entropy(v64, val_labels) + d0.sum() + d1.sum() + d2.sum()
        loss64.backward()
        case["ref64"] = {"arousal": a64.detach().clone(), "valence": v64.detach().clone(),
                         "contrastive": [c.detach().clone() for c in (d0, d1, d2)], "loss": loss64.detach().clone(),
                         "grads": {k: grad_digest(p.grad) for k, p in m64.named_parameters() if p.grad is not None}}
        # eval-mode logits (Tester.py:53 path; running stats after one train step)
        m.eval()
        with torch.no_grad():
            ea, ev = m(*xs)
        m.train()
        case["eval_arousal"], case["eval_valence"] = ea.clone(), ev.clone()
        torch.save(case, os.path.join(OUT, f"native_case_B{B}_T{temp}.pt"))
        print("native", B, temp, float(loss), [float(c) for c in (c0, c1, c2)])


def _load_memhacl():
    spec = importlib.util.spec_from_file_location("memhacl_model", os.path.join(REF, "ME-MHACL", "model.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def memhacl():
    """ME-MHACL fusion tail (both variants), ProjectionHead, Classifier, SupCon, NT-Xent."""
    sys.path.insert(0, REF)
    import MultimodalModel as R
    M = _load_memhacl()
    out = {}
    g = torch.Generator().manual_seed(77)
    B = 24
    feats = [torch.randn(B, 256, generator=g) for _ in range(3)]
    labels = torch.randint(0, 2, (B,), generator=g)

    # mean variant: ME-MHACL/model.py MultiModalEncoder with identity encoders
    torch.manual_seed(1)
    enc = M.MultiModalEncoder()
    # the reference forward unsqueezes eye/phy for its Conv1d encoders, so it cannot take features
    # directly: run its own multihead_attn module through the fusion tail exactly as model.py:68-74 does
    xs = [f.clone().requires_grad_(True) for f in feats]
    x = torch.stack(xs, dim=0)
    a, _ = enc.multihead_attn(x, x, x)
    y = a.mean(dim=0)
    y.square().sum().backward()
    out["mean"] = {"state_dict": {k: v.clone() for k, v in enc.multihead_attn.state_dict().items()},
                   "out": y.detach().clone(), "dfeats": [t.grad.clone() for t in xs],
                   "grads": {k: grad_digest(p.grad) for k, p in enc.multihead_attn.named_parameters()}}

    # max variant: MultimodalModel.py MultiModalEncoder (normalize, max-pool, fusion_mlp)
    torch.manual_seed(2)
    enc2 = R.MultiModalEncoder()
    enc2.eeg_net, enc2.eye_net, enc2.pps_net = nn.Identity(), nn.Identity(), nn.Identity()
    enc2.train()
    xs = [f.clone().requires_grad_(True) for f in feats]
    y2 = enc2(*xs)
    y2.square().sum().backward()
    out["max"] = {"state_dict": {k: v.clone() for k, v in enc2.state_dict().items()},
                  "out": y2.detach().clone(), "dfeats": [t.grad.clone() for t in xs],
                  "grads": {k: grad_digest(p.grad) for k, p in enc2.named_parameters() if p.grad is not None}}

    # ProjectionHead + Classifier (dropout 0)
    torch.manual_seed(3)
    ph = M.ProjectionHead()
    _zero_dropout(ph)
    ph.train()
    sd_ph = {k: v.clone() for k, v in ph.state_dict().items()}
    xin = feats[0].clone().requires_grad_(True)
    z = ph(xin)
    z.square().sum().backward()
    out["projection"] = {"state_dict": sd_ph, "x": feats[0].clone(), "out": z.detach().clone(), "dx": xin.grad.clone(),
                         "grads": {k: grad_digest(p.grad) for k, p in ph.named_parameters()}}
    torch.manual_seed(4)
    cl = R.Classifier()
    _zero_dropout(cl)
    cl.train()
    xin = feats[1].clone().requires_grad_(True)
    oa, ov = cl(xin)
    (F.cross_entropy(oa, labels) + F.cross_entropy(ov, labels)).backward()
    out["classifier"] = {"state_dict": {k: v.clone() for k, v in cl.state_dict().items()}, "x": feats[1].clone(),
                         "out_a": oa.detach().clone(), "out_v": ov.detach().clone(), "dx": xin.grad.clone(),
                         "grads": {k: grad_digest(p.grad) for k, p in cl.named_parameters()}}

    # losses: the reference functions live in scripts that load data at import -> exec only the function source
    def grab(path, name):
        src = open(path, encoding="utf-8").read()
        start = src.index(f"def {name}(")
        end = src.index("\n\n\n", start) if "\n\n\n" in src[start:] else len(src)
        ns = {"torch": torch, "F": F}
        body = src[start:end]
        exec(body, ns)
        return ns[name]
    supcon = grab(os.path.join(REF, "train.py"), "contrastive_loss")
    src = open(os.path.join(REF, "ME-MHACL", "train.py"), encoding="utf-8").read()
    start = src.index("def contrastive_loss(")
    end = src.index("# 1) Pre-training loop")
    ns = {"torch": torch, "F": F}
    exec(src[start:end], ns)
    ntxent = ns["contrastive_loss"]
    z1 = torch.randn(B, 128, generator=g)
    z2 = torch.randn(B, 128, generator=g)
    a1, a2 = z1.clone().requires_grad_(True), z2.clone().requires_grad_(True)
    l1 = supcon(a1, a2, labels, 0.1)
    l1.backward()
    out["supcon"] = {"z1": z1, "z2": z2, "labels": labels, "loss": l1.detach().clone(), "dz1": a1.grad.clone(), "dz2": a2.grad.clone()}
    a1, a2 = z1.clone().requires_grad_(True), z2.clone().requires_grad_(True)
    l2 = ntxent(a1, a2, 0.5)
    l2.backward()
    out["ntxent"] = {"z1": z1, "z2": z2, "loss": l2.detach().clone(), "dz1": a1.grad.clone(), "dz2": a2.grad.clone()}
    out["feats"] = feats
    out["labels"] = labels
    torch.save(out, os.path.join(OUT, "memhacl.pt"))
    print("memhacl", float(l1), float(l2))


def subnetwork():
    """Subnetwork encoder tail (MultimodalModel.py:83-105) from the imported reference, eval mode (dropout off; the
    attention-probability dropout of its MultiheadAttention cannot be matched by any other RNG), gradients kept."""
    sys.path.insert(0, REF)
    import MultimodalModel as R
    torch.manual_seed(11)
    m = R.Subnetwork(38)
    m.eval()
    g = torch.Generator().manual_seed(12)
    x = torch.randn(20, 38, generator=g)
    xin = x.clone().requires_grad_(True)
    y = m(xin)
    # a random linear functional of the output: sum(y^2) is constant behind a LayerNorm (its gradient is rounding noise)
    wgt = torch.randn(y.shape, generator=g)
    (y * wgt).sum().backward()
    out = {"state_dict": {k: v.clone() for k, v in m.state_dict().items()}, "x": x, "wgt": wgt, "out": y.detach().clone(),
           "dx": xin.grad.clone(), "grads": {k: grad_digest(p.grad) for k, p in m.named_parameters()}}
    torch.save(out, os.path.join(OUT, "subnetwork.pt"))
    print("subnetwork", float(y.detach().abs().mean()))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ["native", "memhacl", "subnetwork"]
    if "native" in which:
        native_fusion()
    if "memhacl" in which:
        memhacl()
    if "subnetwork" in which:
        subnetwork()
