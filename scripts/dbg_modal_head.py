"""Diagnostic (GPU box): native-wiring golden cases with the fused modality-head kernel vs the unfused chain (B = 2 is
chaotic through its two-sample BatchNorm stack; B = 20 agrees to 1e-7)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-sentiment-aanalysis_b200")); sys.path.insert(0, ROOT)
import torch, mmsa
from mmsa import ops
dev = torch.device("cuda:0")
GOLD = os.path.join(ROOT, "tests", "golden")
gp = torch.load(os.path.join(GOLD, "native_params.pt"))
for case in ("native_case_B2_T0.01.pt", "native_case_B20_T0.01.pt"):
    c = torch.load(os.path.join(GOLD, case))
    outs = {}
    for mode in ("fused", "unfused"):
        orig = ops.is_modal_head
        if mode == "unfused":
            ops.is_modal_head = lambda s: False
        model = mmsa.MultimodalTransformerModel().set_dropout(0.0)
        model.load_state_dict(gp["state_dict"], strict=True)
        with torch.no_grad(): model.temperature.fill_(c["temperature"])
        model = model.to(dev).train()
        xs = [x.to(dev) for x in c["inputs"]]
        labels, vlabels = c["labels"].to(dev), c["val_labels"].to(dev)
        a, v, c0, c1, c2 = model(*xs, labels=(labels, vlabels))
        loss = mmsa.cross_entropy(a, labels) + mmsa.cross_entropy(v, vlabels) + c0.sum() + c1.sum() + c2.sum()
        loss.backward()
        outs[mode] = (a.detach(), v.detach(), {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None})
        ops.is_modal_head = orig
    r64 = c["ref64"]
    def rel(x, r): return float((x.double().cpu() - r.double().cpu()).abs().max() / max(float(r.double().abs().max()), 1e-12))
    print(case, "arousal fused/unfused vs ref64", rel(outs["fused"][0], r64["arousal"]), rel(outs["unfused"][0], r64["arousal"]), "ref32", rel(c["arousal"], r64["arousal"]))
    print(case, "valence fused/unfused vs ref64", rel(outs["fused"][1], r64["valence"]), rel(outs["unfused"][1], r64["valence"]), "ref32", rel(c["valence"], r64["valence"]))
    worst = []
    for k in outs["fused"][2]:
        worst.append((rel(outs["fused"][2][k], outs["unfused"][2][k]), k))
    worst.sort(reverse=True)
    print("  fused vs unfused grads worst:", worst[:6])
