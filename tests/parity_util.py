"""Shared parity harness: build the drop-in model and the oracle from the same parameters and
seeded inputs, run fwd+bwd on both, and report per-tensor errors.

Error metric (used for every tolerance in tests/): for a tensor x against reference r,
    err = max|x - r| / max(max|r|, floor)
i.e. relative to the tensor's scale (elementwise relative error is meaningless where r crosses 0)."""
from __future__ import annotations

import os
import sys
from typing import Dict, Optional

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "multimodal-sentiment-aanalysis_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

from oracle import fusion_oracle as O  # noqa: E402


# ---------------------------------------------------------------------------------------- strict-bar exceptions report
# north_star's bars are 1e-5 (fp32) and 2e-2 (bf16) relative.  A few quantities are ill-conditioned enough that the
# REFERENCE'S OWN evaluation at the working precision misses its float64 value by more than that (InfoNCE at T = 0.01,
# tiny-batch BatchNorm); those pass through the documented noise-floor clause  err <= noise_mult x reference deviation.
# Every such pass (and every fixed softer bar in the kernel tests) is LOGGED here, so the softness is never silent:
# one JSON line per tensor in $MMSA_PARITY_REPORT (default gpurun_out/parity_report.jsonl), turned into
# profiles/rNN_parity_report.md by scripts/make_parity_report.py.
import json as _json

REPORT_PATH = os.environ.get("MMSA_PARITY_REPORT", os.path.join(ROOT, "gpurun_out", "parity_report.jsonl"))


def _report(rec: Dict) -> None:
    try:
        os.makedirs(os.path.dirname(REPORT_PATH), exist_ok=True)
        with open(REPORT_PATH, "a") as f:
            f.write(_json.dumps(rec) + "\n")
    except OSError:
        pass


def report_summary(test: str, strict: float, errs: Dict[str, float], noise: Optional[Dict[str, float]] = None,
                   bar: Optional[Dict[str, float]] = None, full: bool = False, note: str = "") -> None:
    """log one test case: how many tensors were checked, the worst error, and every tensor above the strict bar"""
    noise = noise or {}
    bar = bar or {}
    worst = max(errs, key=lambda k: errs[k]) if errs else None
    _report({"kind": "summary", "test": test, "strict": strict, "checked": len(errs), "worst": worst,
             "worst_err": errs.get(worst), "over_strict": sum(1 for v in errs.values() if v > strict), "note": note})
    for k, e in errs.items():
        if e > strict or full:
            _report({"kind": "tensor", "test": test, "tensor": k, "err": e, "strict": strict, "noise": noise.get(k),
                     "bar": bar.get(k), "over_strict": e > strict})


def check_close(test: str, name: str, got: torch.Tensor, want64: torch.Tensor, strict: float,
                want_same_precision: Optional[torch.Tensor] = None, noise_mult: float = 3.0,
                fixed_bar: Optional[float] = None, why: str = "") -> float:
    """assert rel_err(got, want64) <= max(strict, noise_mult x rel_err(reference at the same precision, want64)) (or
    <= fixed_bar when given); whenever the strict bar alone would not have passed, the exception is logged."""
    err = rel_err(got, want64)
    noise = rel_err(want_same_precision, want64) if want_same_precision is not None else None
    bar = fixed_bar if fixed_bar is not None else max(strict, noise_mult * (noise or 0.0))
    if err > strict:
        _report({"kind": "tensor", "test": test, "tensor": name, "err": err, "strict": strict, "noise": noise, "bar": bar,
                 "over_strict": True, "why": why})
    assert err <= bar, f"{test}: {name}: err {err:.3e} > bar {bar:.3e} (strict {strict:.1e}, reference noise {noise})"
    return err


def rel_err(x: torch.Tensor, r: torch.Tensor, floor: float = 1e-12) -> float:
    x = x.detach().double().cpu().reshape(-1)
    r = r.detach().double().cpu().reshape(-1)
    if x.numel() == 0:
        return 0.0
    return float((x - r).abs().max() / max(float(r.abs().max()), floor))


def zero_grad_bias_keys(names):
    """Biases of a Linear that feeds a training-mode BatchNorm1d (fusion.0/.4, arousal_head.0,
    valence_head.0/.4/.8/.12; MultimodalModel.py:179-225).  BatchNorm subtracts the batch mean, so
    dL/db is EXACTLY zero in exact arithmetic; both the reference and the kernels return rounding
    noise of the order eps * |dz|.  A relative error against noise is meaningless, so these are
    checked for magnitude against the gradient scale of the same layer's weight instead."""
    out = set()
    names = set(names)
    for n in names:
        if not n.endswith(".bias"):
            continue
        stem, idx = n[:-5].rsplit(".", 1) if "." in n[:-5] else (n[:-5], "")
        if not idx.isdigit():
            continue
        head = stem.split(".")[-1]
        if head in ("fusion", "arousal_head", "valence_head") and f"{stem}.{int(idx) + 1}.weight" in names:
            out.add(n)
    return out


def grad_err(name: str, g: torch.Tensor, ref: torch.Tensor, ref_grads: Dict[str, torch.Tensor], zero_keys) -> float:
    """rel_err for a parameter gradient; the mathematically-zero biases (zero_grad_bias_keys) are
    measured as max|g| / max|dW of the same Linear|."""
    if name in zero_keys:
        w = ref_grads[name[:-5] + ".weight"]
        return float(g.detach().double().abs().max().cpu()) / max(float(w.abs().max()), 1e-30)
    return rel_err(g, ref)


def oracle_step(cfg: O.FusionConfig, params: Dict[str, torch.Tensor], inputs, labels, dtype=torch.float32,
                gathered=None, autocast_bf16: bool = False):
    """Oracle fwd+bwd of the Trainer.py:60-79 step: loss = CE(arousal) + sum(contrastive).
    autocast_bf16: evaluate under torch.autocast(cpu, bfloat16) with fp32 parameters -- the reference's
    arithmetic as PyTorch itself runs it in bf16 (matmuls in bf16, reductions / norms / losses in fp32)."""
    p = {k: v.detach().clone().to(dtype).requires_grad_(True) for k, v in params.items()}
    xs = tuple(x.to(dtype) for x in inputs)
    if autocast_bf16:
        with torch.autocast(device_type="cpu", dtype=torch.bfloat16):
            loss, out = O.trainer_loss(cfg, p, xs, labels, training=True, gathered=gathered)
        loss = loss.float()
    else:
        loss, out = O.trainer_loss(cfg, p, xs, labels, training=True, gathered=gathered)
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in p.items()}
    return loss.detach(), out, grads


def build_model(cfg: O.FusionConfig, params, buffers, compute_dtype, device):
    import mmsa
    model = mmsa.MultimodalTransformerModel(
        num_classes=cfg.num_classes, embed_dim=cfg.embed_dim, num_heads=cfg.num_heads, wiring=cfg.wiring,
        text_dim=cfg.text_dim, image_dim=cfg.image_dim, contract=cfg.contract, compute_dtype=compute_dtype,
        valence=cfg.valence)
    sd = {k: v.clone() for k, v in params.items()}
    sd.update({k: v.clone() for k, v in buffers.items()})
    missing, unexpected = model.load_state_dict(sd, strict=True)
    model.set_dropout(0.0)
    return model.to(device).train()


def run_fusion_parity(batch: int = 4, L: int = 64, R: int = 49, dtype: str = "fp32", tol: float = 1e-5,
                      device: str = "cuda:0", embed_dim: int = 768, num_heads: int = 12, text_dim: int = 768,
                      image_dim: int = 2048, seed: int = 0, noise_mult: float = 3.0,
                      temperature: Optional[float] = None, name: Optional[str] = None, full_report: bool = False) -> Dict:
    """Bidirectional (text+image) fusion step on the GPU vs the CPU oracle on identical inputs.

    The yardstick is the oracle evaluated in float64.  A tensor passes when its error against that is
    within `tol` (1e-5 fp32 / 2e-2 bf16, BASELINE.json north_star), or -- for the few quantities where
    the REFERENCE'S OWN evaluation at the same precision (fp32; bf16 = torch CPU autocast) is further than
    tol/noise_mult from its float64 value (tiny-batch BatchNorm, InfoNCE at small T: condition numbers
    of 1e2-1e3) -- within noise_mult x that deviation: no implementation at that precision with a
    different summation order can do better."""
    import mmsa
    cd = torch.float32 if dtype == "fp32" else torch.bfloat16
    cfg = O.FusionConfig(embed_dim=embed_dim, num_heads=num_heads, wiring="bidirectional", text_dim=text_dim,
                         image_dim=image_dim, contract="single", valence=False)
    params, buffers = O.init_params(cfg, seed=seed)
    if temperature is not None:
        params["temperature"] = torch.tensor(float(temperature))
    inputs, labels = O.synth_inputs(cfg, batch, L=L, R=R, seed=1234 + seed)
    o_loss, o_out, o_grads = oracle_step(cfg, params, inputs, labels, dtype=torch.float64)
    # the reference arithmetic at the SAME working precision (fp32, or bf16 autocast): its deviation from
    # the float64 value is the noise floor any implementation at that precision shares
    n_loss, n_out, n_grads = oracle_step(cfg, params, inputs, labels, dtype=torch.float32,
                                         autocast_bf16=(dtype == "bf16"))

    model = build_model(cfg, params, buffers, cd, device)
    text, image = (x.to(device) for x in inputs)
    lab = labels.to(device)
    logits, closs = model(text, image, None, lab)
    ce = mmsa.cross_entropy(logits, lab)
    loss = ce + closs.sum()
    loss.backward()
    torch.cuda.synchronize()

    zero_keys = zero_grad_bias_keys(o_grads.keys())
    errs = {"logits": rel_err(logits, o_out.arousal), "loss": rel_err(loss, o_loss)}
    noise = {"logits": rel_err(n_out.arousal, o_out.arousal), "loss": rel_err(n_loss, o_loss)}
    for k, prm in model.named_parameters():
        g = prm.grad if prm.grad is not None else torch.zeros_like(prm)
        errs["grad:" + k] = grad_err(k, g, o_grads[k], o_grads, zero_keys)
        noise["grad:" + k] = grad_err(k, n_grads[k], o_grads[k], o_grads, zero_keys)
    excess = {k: errs[k] / max(tol, noise_mult * noise[k]) for k in errs}
    worst = max(excess, key=lambda k: excess[k])
    report_summary(name or f"fusion_step[{dtype},B={batch},L={L},T={temperature}]", tol, errs, noise,
                   {k: max(tol, noise_mult * noise[k]) for k in errs}, full=full_report,
                   note=f"noise = the oracle's own {'bf16-autocast' if dtype == 'bf16' else 'fp32'} deviation from float64; bar = max(strict, {noise_mult} x noise)")
    labels_equal = bool(torch.equal(logits.argmax(1).cpu(), o_out.arousal.argmax(1)))
    top2 = o_out.arousal.detach().topk(2, dim=1).values
    # labels: bit-exact wherever the oracle's own top-2 margin exceeds the precision bar on the logits (a sample whose
    # two best logits differ by less than the allowed logit error has no defined label at that precision)
    margin = (top2[:, 0] - top2[:, 1])
    sure = margin > tol * float(o_out.arousal.detach().abs().max())
    pred, want = logits.argmax(1).cpu(), o_out.arousal.argmax(1)
    labels_equal_sure = bool(torch.equal(pred[sure], want[sure]))
    _report({"kind": "labels", "test": name or f"fusion_step[{dtype},B={batch},L={L},T={temperature}]",
             "samples": int(batch), "mismatches": int((pred != want).sum()), "below_margin": int((~sure).sum()),
             "mismatches_above_margin": int((pred[sure] != want[sure]).sum()), "min_margin": float(margin.min())})
    return {"ok": excess[worst] <= 1.0 and labels_equal_sure, "labels_equal_sure": labels_equal_sure, "max_rel": errs[worst], "worst": worst, "errs": errs,
            "noise": noise, "excess": excess[worst],
            "failing": {k: (errs[k], noise[k]) for k in errs if excess[k] > 1.0},
            "loss": float(loss.detach()), "oracle_loss": float(o_loss), "labels_equal": labels_equal,
            "logit_margin": float((top2[:, 0] - top2[:, 1]).min())}
