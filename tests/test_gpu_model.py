"""Model-level parity (-m gpu): the drop-in modules, called the way Trainer.py / Tester.py /
MultiTaskTrainer.py call the reference model, against the CPU oracle and the committed goldens
(generated from the imported reference by oracle/make_goldens.py).
Tolerances: fp32 1e-5, bf16 2e-2 (parity_util.rel_err; BASELINE.json north_star)."""
import os

import pytest
import torch
import torch.nn.functional as F

from parity_util import O, build_model, check_close, oracle_step, rel_err, run_fusion_parity, zero_grad_bias_keys

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _digest_err(flat: torch.Tensor, dig) -> float:
    """error of a gradient against a stored digest (64 sampled elements, tensor absmax as scale).
    The scale is floored at 1e-12: with O(1) parameters and losses a gradient that small is an
    underflow residue (dL/dT of the InfoNCE at T = 0.01 is 1e-22 in the reference), not a value."""
    scale = max(dig["absmax"], 1e-12)
    return float((flat[dig["idx"]].double() - dig["vals"].double()).abs().max()) / scale


def _check_digest(grad: torch.Tensor, dig32, dig64, tol: float, name: str, noise_mult: float = 3.0, test: str = ""):
    """ours vs the reference's float64 gradient, within max(tol, noise_mult x the fp32 reference's own deviation); an error
    above the strict bar is logged to the parity report."""
    flat = grad.detach().reshape(-1).cpu()
    dig64 = dig64 if dig64 is not None else dig32
    noise = float((dig32["vals"].double() - dig64["vals"].double()).abs().max()) / max(dig64["absmax"], 1e-12)
    err = _digest_err(flat, dig64)
    if err > tol and test:
        from parity_util import _report
        _report({"kind": "tensor", "test": test, "tensor": "grad:" + name + " (64 sampled elements)", "err": err, "strict": tol,
                 "noise": noise, "bar": max(tol, noise_mult * noise), "over_strict": True,
                 "why": "tiny-batch BatchNorm stack" if noise_mult > 3.0 else ""})
    assert err <= max(tol, noise_mult * noise), f"{name}: sampled grad err {err:.3e} (fp32 reference noise {noise:.3e})"
    n_err = abs(float(flat.double().norm()) - dig64["norm"]) / max(dig64["norm"], 1e-12)
    n_noise = abs(dig32["norm"] - dig64["norm"]) / max(dig64["norm"], 1e-12)
    assert n_err <= max(tol * 10, noise_mult * n_noise), f"{name}: grad norm err {n_err:.3e}"


# ------------------------------------------------------------------ re-skinned (text + image) path
@pytest.mark.parametrize("batch,L", [(8, 64), (6, 128), (4, 512)])
def test_fusion_step_fp32(cuda_device, batch, L):
    rep = run_fusion_parity(batch=batch, L=L, R=49, dtype="fp32", tol=1e-5)
    assert rep["labels_equal"]
    assert rep["ok"], (rep["worst"], rep["max_rel"], rep["failing"])


@pytest.mark.parametrize("batch,L,temp", [(32, 64, None), (16, 128, None), (16, 128, 0.07)])
def test_fusion_step_bf16(cuda_device, batch, L, temp):
    # bf16: within 2e-2 of the float64 oracle, or no worse than torch's own bf16 autocast of the reference
    rep = run_fusion_parity(batch=batch, L=L, R=49, dtype="bf16", tol=2e-2, temperature=temp, noise_mult=1.0)
    assert rep["labels_equal"]
    assert rep["ok"], (rep["worst"], rep["max_rel"], rep["logit_margin"], rep["failing"])


def test_fusion_eval_and_contracts(cuda_device):
    """Tester.py:53 (`outputs = model(eeg, eye, pps)` under no_grad/eval) and the return arities."""
    cfg = O.FusionConfig(embed_dim=768, num_heads=12, wiring="bidirectional", contract="single", valence=False)
    params, buffers = O.init_params(cfg, seed=1)
    inputs, labels = O.synth_inputs(cfg, 6, L=64, R=49, seed=5)
    model = build_model(cfg, params, buffers, torch.float32, cuda_device)
    text, image = (x.to(cuda_device) for x in inputs)
    out = model(text, image, None, labels.to(cuda_device))
    assert isinstance(out, tuple) and len(out) == 2 and out[0].shape == (6, 3) and out[1].shape == (1,)
    bns = [m for m in model.modules() if isinstance(m, torch.nn.BatchNorm1d) and m in list(model.fusion) + list(model.arousal_head)]
    assert bns and all(int(m.num_batches_tracked) == 1 for m in bns)      # counted by the BatchNorm kernel itself
    model.eval()
    with torch.no_grad():
        logits = model(text, image, None)
    assert all(int(m.num_batches_tracked) == 1 for m in bns)              # eval mode does not count
    assert logits.shape == (6, 3) and logits.dtype == torch.float32
    p = {k: v for k, v in params.items()}
    p.update(model.state_dict())                       # running stats after the one train step
    p = {k: v.detach().cpu() for k, v in p.items()}
    ref = O.fusion_forward(cfg, p, inputs, None, training=False)
    assert rel_err(logits, ref.arousal) <= 1e-5
    probs = torch.softmax(logits, dim=1)               # Tester.py:54
    assert torch.isfinite(probs).all()


# ------------------------------------------------------------------ native wiring vs goldens from the reference
@pytest.mark.parametrize("case", ["native_case_B2_T0.01.pt", "native_case_B20_T0.01.pt", "native_case_B20_T0.5.pt",
                                  "native_case_B64_T0.07.pt"])
def test_native_against_reference_goldens(cuda_device, case):
    import mmsa
    gp = torch.load(os.path.join(GOLD, "native_params.pt"))
    c = torch.load(os.path.join(GOLD, case))
    model = mmsa.MultimodalTransformerModel().set_dropout(0.0)
    assert list(model.state_dict().keys()) == gp["keys"], "state_dict keys must match the reference's"
    model.load_state_dict(gp["state_dict"], strict=True)
    with torch.no_grad():
        model.temperature.fill_(c["temperature"])
    model = model.to(cuda_device).train()
    xs = [x.to(cuda_device) for x in c["inputs"]]
    labels, vlabels = c["labels"].to(cuda_device), c["val_labels"].to(cuda_device)
    a, v, c0, c1, c2 = model(*xs, labels=(labels, vlabels))       # MultiTaskTrainer.py:199 contract
    assert a.shape == c["arousal"].shape and c0.shape == (1,)
    loss = mmsa.cross_entropy(a, labels) + mmsa.cross_entropy(v, vlabels) + c0.sum() + c1.sum() + c2.sum()
    loss.backward()
    r64 = c["ref64"]        # the reference evaluated in float64; c[...] itself is the fp32 reference

    # B = 2: every BatchNorm divides by a TWO-sample standard deviation (rstd = 1/sqrt(var + 1e-5) with var ~ 1e-6 where the
    # two samples nearly agree), so fp32 rounding anywhere upstream is amplified ~1e4x per layer and valence_head stacks four
    # of them: the fp32 reference itself sits ~1e-3 from its float64 value there, and two fp32 evaluations that differ only
    # in summation order (this library's fused and unfused tail, scripts/dbg_modal_head.py) differ from each other by
    # 2-5 x that.  The clause is 3 x the reference's deviation (10 x for the B = 2 case, whose "noise" is a single draw of a
    # heavy-tailed quantity); every excess over 1e-5 is logged.  At B = 20 / 64 the same checks hold at the 3 x clause.
    nm = 10.0 if c["B"] < 8 else 3.0

    def close(got, want32, want64, tol=1e-5, name=""):
        check_close(f"native_goldens[{case}]", name, got, want64, tol, want32, noise_mult=nm,
                    why="tiny-batch BatchNorm stack (see test comment)" if c["B"] < 8 else "")
        return True
    assert close(a, c["arousal"], r64["arousal"], name="arousal logits") and close(v, c["valence"], r64["valence"], name="valence logits")
    for i, (got, w32, w64) in enumerate(zip((c0, c1, c2), c["contrastive"], r64["contrastive"])):
        assert close(got, w32, w64, name=f"contrastive[{i}]")
    assert close(loss, c["loss"], r64["loss"], name="loss")
    assert torch.equal(a.argmax(1).cpu(), c["arousal"].argmax(1))
    zero_keys = zero_grad_bias_keys(c["grads"].keys())
    for k, prm in model.named_parameters():
        if k in zero_keys:      # exactly-zero gradient (bias in front of BatchNorm): magnitude check
            assert float(prm.grad.abs().max()) <= 1e-5 * c["grads"][k[:-5] + ".weight"]["absmax"], k
        elif k in c["grads"]:
            _check_digest(prm.grad, c["grads"][k], r64["grads"].get(k), 1e-5, k, noise_mult=nm, test=f"native_goldens[{case}]")
    # running statistics: the golden holds the fp32 reference only, so the float64 yardstick comes from the
    # oracle (bit-identical to the reference in fp32, test_cpu_oracle_and_abi).  With B = 2 every BatchNorm
    # divides by a two-sample standard deviation, which amplifies fp32 rounding layer by layer (valence_head
    # stacks four): the reference's own fp32 buffers sit up to ~1e-4 from the exact value there
    p64 = {k: v.double() for k, v in gp["state_dict"].items() if v.is_floating_point()}
    p64["temperature"] = torch.tensor(c["temperature"], dtype=torch.float64)
    b64 = {k: v.double() if v.is_floating_point() else v.clone() for k, v in gp["state_dict"].items()
           if "running_" in k or "num_batches" in k}
    O.fusion_forward(O.FusionConfig(), p64, [x.double() for x in c["inputs"]], c["labels"],
                     training=True, buffers=b64)
    for k, b in model.named_buffers():
        if b.is_floating_point():
            want32 = c["buffers_after"][k].float()
            check_close(f"native_goldens[{case}]", "buffer:" + k, b.float(), b64[k], 1e-5, want32, noise_mult=nm,
                        why="tiny-batch BatchNorm stack" if c["B"] < 8 else "")
        else:
            assert torch.equal(b.cpu(), c["buffers_after"][k]), k
    model.eval()
    with torch.no_grad():
        ea, ev = model(*xs)
    # eval mode reads the running statistics written by the train step above; with B = 2 those carry
    # the amplified noise of the 2-sample variance
    etol = 1e-5 if c["B"] >= 8 else nm * max(rel_err(c["arousal"], r64["arousal"]), rel_err(c["valence"], r64["valence"]), 1e-5)
    assert rel_err(ea, c["eval_arousal"]) <= etol and rel_err(ev, c["eval_valence"]) <= etol


def test_single_contract_native(cuda_device):
    """Trainer.py:60 `outputs, contrastive_loss = model(eeg, eye, pps, labels)` with 1-D labels."""
    import mmsa
    model = mmsa.MultimodalTransformerModel(contract="single").set_dropout(0.0).to(cuda_device).train()
    cfg = O.FusionConfig()
    xs, labels = O.synth_inputs(cfg, 16, seed=9)
    out, closs = model(*[x.to(cuda_device) for x in xs], labels.to(cuda_device))
    assert out.shape == (16, 3) and closs.shape == (1,)
    crit = torch.nn.CrossEntropyLoss()                 # the trainer's own criterion works on our logits
    w = torch.nn.Parameter(torch.ones(1, device=cuda_device))
    loss = crit(out, labels.to(cuda_device)) + w * closs
    loss.backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
    # the single-task contract never runs valence_head (Trainer.py:60 has one logits tensor), so its
    # parameters keep grad None, which clip_grad_norm_ / AdamW skip; everything else must be finite
    for n, p in model.named_parameters():
        if n.startswith("valence_head."):
            assert p.grad is None, n
        else:
            assert p.grad is not None and torch.isfinite(p.grad).all(), n
    assert w.grad is not None


def test_requires_grad_toggling(cuda_device):
    """MultiTaskTrainer.py:50-177 freezes sub-modules by attribute name; frozen params get no grad."""
    import mmsa
    model = mmsa.MultimodalTransformerModel().set_dropout(0.0).to(cuda_device).train()
    for p in model.cross_attn_e2p.parameters():
        p.requires_grad = False
    for p in model.fusion.parameters():
        p.requires_grad = False
    cfg = O.FusionConfig()
    xs, labels = O.synth_inputs(cfg, 8, seed=3)
    a, v, c0, c1, c2 = model(*[x.to(cuda_device) for x in xs], labels=(labels.to(cuda_device), labels.to(cuda_device)))
    (mmsa.cross_entropy(a, labels.to(cuda_device)) + c0.sum()).backward()
    assert all(p.grad is None for p in model.cross_attn_e2p.parameters())
    assert all(p.grad is None for p in model.fusion.parameters())
    assert all(p.grad is not None for p in model.arousal_head.parameters())
    assert model.temperature.grad is not None and model.contrastive_weight.grad is not None


# ------------------------------------------------------------------ ME-MHACL pieces vs goldens from the reference
def test_memhacl_against_reference_goldens(cuda_device):
    import mmsa
    g = torch.load(os.path.join(GOLD, "memhacl.pt"))
    feats = [f.to(cuda_device) for f in g["feats"]]
    labels = g["labels"].to(cuda_device)
    # mean variant (ME-MHACL/model.py:68-74)
    enc = mmsa.MultiModalEncoder(variant="mean")
    enc.multihead_attn.load_state_dict(g["mean"]["state_dict"])
    enc = enc.to(cuda_device).train()
    xs = [f.clone().requires_grad_(True) for f in feats]
    y = enc(*xs)
    y.square().sum().backward()
    assert rel_err(y, g["mean"]["out"]) <= 1e-5
    for x, d in zip(xs, g["mean"]["dfeats"]):
        assert rel_err(x.grad, d) <= 1e-5
    for k, p in enc.multihead_attn.named_parameters():
        _check_digest(p.grad, g["mean"]["grads"][k], None, 5e-5, k)
    # max variant (MultimodalModel.py:388-406)
    enc2 = mmsa.MultiModalEncoder(variant="max")
    enc2.load_state_dict(g["max"]["state_dict"], strict=True)
    enc2 = enc2.to(cuda_device).train()
    xs = [f.clone().requires_grad_(True) for f in feats]
    y2 = enc2(*xs)
    y2.square().sum().backward()
    assert rel_err(y2, g["max"]["out"]) <= 1e-5
    # the golden input gradients are the reference's fp32 values, which sit 2e-4 from the float64
    # evaluation themselves (L2-normalise + max-pool + BatchNorm backward): judge against float64,
    # bounded below by the reference's own deviation
    p64 = {k: v.double() for k, v in g["max"]["state_dict"].items() if v.is_floating_point()}
    xs64 = [f.double().clone().requires_grad_(True) for f in g["feats"]]
    O.memhacl_fusion(xs64, p64, num_heads=8, variant="max", training=True).square().sum().backward()
    for x, d, x64 in zip(xs, g["max"]["dfeats"], xs64):
        assert rel_err(x.grad, x64.grad) <= max(2e-5, 3.0 * rel_err(d, x64.grad))
    # ProjectionHead (ME-MHACL/model.py:82-97) and Classifier
    ph = mmsa.ProjectionHead()
    ph.load_state_dict(g["projection"]["state_dict"], strict=True)
    for m in ph.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    ph = ph.to(cuda_device).train()
    x = g["projection"]["x"].to(cuda_device).requires_grad_(True)
    z = ph(x)
    z.square().sum().backward()
    assert rel_err(z, g["projection"]["out"]) <= 1e-5 and rel_err(x.grad, g["projection"]["dx"]) <= 1e-5
    for k, p in ph.named_parameters():
        _check_digest(p.grad, g["projection"]["grads"][k], None, 5e-5, k)
    cl = mmsa.Classifier()
    cl.load_state_dict(g["classifier"]["state_dict"], strict=True)
    for m in cl.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    cl = cl.to(cuda_device).train()
    x = g["classifier"]["x"].to(cuda_device).requires_grad_(True)
    oa, ov = cl(x)
    (mmsa.cross_entropy(oa, labels) + mmsa.cross_entropy(ov, labels)).backward()
    assert rel_err(oa, g["classifier"]["out_a"]) <= 1e-5 and rel_err(ov, g["classifier"]["out_v"]) <= 1e-5
    assert rel_err(x.grad, g["classifier"]["dx"]) <= 1e-5
    # SupCon (train.py:16-40) and NT-Xent (ME-MHACL/train.py:47-66)
    for name in ("supcon", "ntxent"):
        c = g[name]
        z1, z2 = c["z1"].to(cuda_device).requires_grad_(True), c["z2"].to(cuda_device).requires_grad_(True)
        loss = mmsa.supcon(z1, z2, c["labels"].to(cuda_device), 0.1) if name == "supcon" else mmsa.ntxent(z1, z2, 0.5)
        loss.backward()
        assert rel_err(loss, c["loss"]) <= 1e-5, name
        assert rel_err(z1.grad, c["dz1"]) <= 2e-5 and rel_err(z2.grad, c["dz2"]) <= 2e-5, name


def test_no_fallback_on_cpu_tensors(cuda_device):
    """CPU tensors must fail loudly, never silently run a PyTorch path."""
    import mmsa
    from mmsa._lib import MmsaError
    model = mmsa.MultimodalTransformerModel().train()          # parameters left on the CPU
    cfg = O.FusionConfig()
    xs, labels = O.synth_inputs(cfg, 4, seed=1)
    with pytest.raises((MmsaError, RuntimeError)):
        model(*xs, labels=(labels, labels))


# ------------------------------------------------------------------ BASELINE.json full sizes: size-independent properties
def _full_step(model, text, image, labels):
    import mmsa
    model.zero_grad(set_to_none=True)
    logits, closs = model(text, image, None, labels)
    loss = mmsa.cross_entropy(logits, labels) + closs.sum()
    loss.backward()
    return logits.detach(), loss.detach(), {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}


def test_step_loss_assembly_matches_trainer_form(cuda_device):
    """mmsa.TrainStep folds `CE + w * contrastive` (Trainer.py:68-71) into the CE launch and seeds backward with a static
    one: same loss and same gradients as the trainer's own expression, and unused outputs of the fused nodes (logits only,
    no contrastive term) still back-propagate."""
    import mmsa
    cfg = O.FusionConfig(embed_dim=256, num_heads=4, wiring="bidirectional", contract="single", valence=False)
    params, buffers = O.init_params(cfg, seed=2)
    inputs, labels = O.synth_inputs(cfg, 16, L=32, R=49, seed=6)
    text, image = (x.to(cuda_device) for x in inputs)
    labels = labels.to(cuda_device)
    model = build_model(cfg, params, buffers, torch.float32, cuda_device).set_dropout(0.0)
    _, loss_a, grads_a = _full_step(model, text, image, labels)
    buffers2 = {k: v.clone() for k, v in buffers.items()}
    model2 = build_model(cfg, params, buffers2, torch.float32, cuda_device).set_dropout(0.0)
    logits, closs = model2(text, image, None, labels)
    loss_b = mmsa.cross_entropy(logits, labels, extra=closs)
    loss_b.backward(gradient=torch.ones((), device=cuda_device))
    assert rel_err(loss_b, loss_a) <= 1e-6
    for k, p in model2.named_parameters():
        if k in grads_a:
            assert rel_err(p.grad, grads_a[k]) <= 1e-5, k
    # only the classifier output is used: the contrastive features' gradients are absent (None, not zero-filled)
    model2.zero_grad(set_to_none=True)
    logits = model2(text, image, None)
    mmsa.cross_entropy(logits, labels).backward()
    assert model2.eeg_net.proj.weight.grad is not None and torch.isfinite(model2.eeg_net.proj.weight.grad).all()
    assert model2.temperature.grad is None


@pytest.mark.parametrize("dtype,tol,noise_mult", [("fp32", 1e-5, 3.0), ("bf16", 2e-2, 1.0)])
@pytest.mark.parametrize("batch,L", [(256, 128), (64, 512)])
def test_full_size_oracle(cuda_device, batch, L, dtype, tol, noise_mult):
    """BASELINE.json configs[1] EXACTLY (B=256, L=128) and a configs[4]-shaped slice (B=64, L=512) against the float64
    CPU oracle: logits, loss, predicted labels and every parameter gradient.  M = B*L = 32768 rows puts the 2-CTA (PAIR)
    256x256 GEMM tiles, the 6-way cluster split-K weight gradients and multi-tile attention of the benchmarked step under
    the oracle (the small-shape tests stay below the M >= 4096 switch).  Bars: 1e-5 (fp32 mode) / 2e-2 (bf16 mode), or the
    noise floor of the reference's own arithmetic at that precision; every tensor's error goes to the parity report."""
    rep = run_fusion_parity(batch=batch, L=L, R=49, dtype=dtype, tol=tol, noise_mult=noise_mult, seed=3,
                            temperature=0.07 if dtype == "bf16" else None,
                            name=f"full_size_oracle[{dtype},B={batch},L={L}]", full_report=True)
    print(f"full-size oracle parity ({dtype}, B={batch}, L={L}): worst {rep['worst']} err {rep['max_rel']:.3e}, "
          f"loss {rep['loss']:.6f} vs {rep['oracle_loss']:.6f}")
    if dtype == "fp32":
        assert rep["labels_equal"], "predicted labels differ from the oracle's"
    assert rep["labels_equal_sure"], "predicted labels differ from the oracle's above the logit error bar"
    assert rep["ok"], (rep["worst"], rep["max_rel"], rep["failing"])


@pytest.mark.parametrize("batch,L", [(256, 128), (64, 512)])
def test_full_size_properties(cuda_device, batch, L):
    """configs[1] (B=256, L=128) and a configs[4]-shaped slice (L=512), size-independent properties on top of
    test_full_size_oracle -- (1) determinism: two bf16 steps on the same inputs are bit-identical (no atomics, fixed
    reduction orders, split-K in split order); (2) the bf16 tensor-core path agrees with this library's own exact-fp32
    CUDA-core path (itself oracle-checked at small sizes) within the bf16 bar of 2e-2 on loss and logits (5e-2 on every
    parameter-gradient norm); (3) predicted labels agree wherever the fp32 logit margin exceeds the bf16 error bar;
    (4) sample independence: the logits of the first 32 samples do not change when they are run in eval mode alone
    or inside the full batch (BatchNorm running statistics, no cross-sample leakage through tiles / padding)."""
    cfg = O.FusionConfig(embed_dim=768, num_heads=12, wiring="bidirectional", contract="single", valence=False)
    params, buffers = O.init_params(cfg, seed=3)
    params["temperature"] = torch.tensor(0.07)
    inputs, labels = O.synth_inputs(cfg, batch, L=L, R=49, seed=77)
    text32, image32 = (x.to(cuda_device) for x in inputs)
    lab = labels.to(cuda_device)
    m16 = build_model(cfg, params, buffers, torch.bfloat16, cuda_device)
    t16, i16 = text32.bfloat16(), image32.bfloat16()
    lg_a, loss_a, g_a = _full_step(m16, t16, i16, lab)
    m16b = build_model(cfg, params, buffers, torch.bfloat16, cuda_device)
    lg_b, loss_b, g_b = _full_step(m16b, t16, i16, lab)
    assert torch.equal(lg_a, lg_b) and torch.equal(loss_a, loss_b)
    for k in g_a:
        assert torch.equal(g_a[k], g_b[k]), f"non-deterministic gradient {k}"
    m32 = build_model(cfg, params, buffers, torch.float32, cuda_device)
    lg32, loss32, g32 = _full_step(m32, t16.float(), i16.float(), lab)       # same (bf16-representable) inputs
    assert torch.isfinite(loss_a) and rel_err(loss_a, loss32) <= 2e-2
    assert rel_err(lg_a, lg32) <= 2e-2
    zero_keys = zero_grad_bias_keys(g32.keys())
    bad, worst = [], (0.0, 0.0)
    for k in g32:
        if k in zero_keys:
            continue
        # gradients pass through ~10 bf16-rounded layers (2^-9 per rounding); bias / LayerNorm-affine / temperature
        # gradients are additionally long, strongly cancelling sums over the batch.  Bars: tensor norm within 5e-2,
        # element-wise error (relative to the tensor's largest element) within 2e-1
        n16, n32 = float(g_a[k].double().norm()), float(g32[k].double().norm())
        en, ee = abs(n16 - n32) / max(n32, 1e-12), rel_err(g_a[k], g32[k])
        worst = (max(worst[0], en), max(worst[1], ee))
        # (a 3-element bias gradient that sums to zero by the softmax constraint gets the element bar for its norm too)
        if en > (5e-2 if g32[k].numel() > 8 else 2e-1) or ee > 2e-1:
            bad.append((k, round(en, 4), round(ee, 4)))
    print(f"full-size bf16 vs fp32 (B={batch}, L={L}): worst norm err {worst[0]:.3e}, worst element err {worst[1]:.3e}")
    assert not bad, bad
    top2 = lg32.topk(2, dim=1).values
    margin = top2[:, 0] - top2[:, 1]
    sure = margin > 4e-2 * float(lg32.abs().max())
    assert torch.equal(lg_a.argmax(1)[sure], lg32.argmax(1)[sure])
    m32.eval()
    with torch.no_grad():
        full = m32(t16.float(), i16.float(), None)
        head = m32(t16[:32].float().contiguous(), i16[:32].float().contiguous(), None)
    assert rel_err(head, full[:32]) <= 1e-5


# ------------------------------------------------------------------ encoder tail
def test_subnetwork_against_reference_golden(cuda_device):
    """mmsa.Subnetwork (proj -> +PE -> 2 x post-norm TransformerEncoderLayer -> LayerNorm, MultimodalModel.py:83-105) against
    the golden from the imported reference (eval mode), fp32 1e-5, every gradient."""
    import mmsa
    g = torch.load(os.path.join(GOLD, "subnetwork.pt"))
    m = mmsa.Subnetwork(38)
    assert list(m.state_dict().keys()) == list(g["state_dict"].keys())
    m.load_state_dict(g["state_dict"], strict=True)
    m = m.to(cuda_device).eval()
    x = g["x"].to(cuda_device).requires_grad_(True)
    y = m(x)
    assert y.shape == g["out"].shape
    (y * g["wgt"].to(cuda_device)).sum().backward()
    assert rel_err(y, g["out"]) <= 1e-5
    assert rel_err(x.grad, g["dx"]) <= 2e-5
    for k, prm in m.named_parameters():
        _check_digest(prm.grad, g["grads"][k], None, 5e-5, k)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 2e-2)])
def test_subnetwork_token_sequences(cuda_device, dtype, tol):
    """generalised to L > 1 (text tokens get a self-attention stage): [B, L, 768] -> [B, L, 256] against the float64 oracle."""
    import mmsa
    torch.manual_seed(21)
    m = mmsa.Subnetwork(768, feat_dim=256, num_layers=2, nhead=4, compute_dtype=dtype)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    m = m.to(cuda_device).eval()
    B, L = 6, 24
    x = torch.randn(B, L, 768)
    p64 = {k: v.double().requires_grad_(k != "pos_encoder.pe") for k, v in sd.items()}
    x64 = x.double().requires_grad_(True)
    wgt = torch.randn(B, L, 256)           # random linear functional: sum(y^2) is constant behind the final LayerNorm
    want = O.subnetwork(x64, p64)
    (want * wgt.double()).sum().backward()
    xg = x.to(cuda_device).requires_grad_(True)
    y = m(xg)
    (y * wgt.to(cuda_device)).sum().backward()
    assert y.shape == (B, L, 256) and y.dtype == torch.float32
    assert rel_err(y, want.detach()) <= tol
    # gradients: fp32 5e-5 element-wise; bf16 passes two layers of rounded activations and ReLU masks that flip near zero:
    # tensor norm within 5e-2, element-wise (relative to the tensor's largest element) within 2e-1, as in the full-size test
    def ok(got, want):
        if dtype == torch.float32:
            return rel_err(got, want) <= 5e-5
        n_got, n_want = float(got.double().norm()), float(want.double().norm())
        return abs(n_got - n_want) <= 5e-2 * n_want and rel_err(got, want) <= 2e-1
    assert ok(xg.grad, x64.grad)
    for k, prm in m.named_parameters():
        assert ok(prm.grad, p64[k].grad), k


# ------------------------------------------------------------------ dropout under CUDA-graph replay
def test_dropout_stream_advances_under_graph_replay(cuda_device):
    """TrainStep(use_graph=True) with Dropout(0.3) active: the Philox position lives in device memory and is moved by a
    captured kernel node, so (1) two replays of the SAME captured graph on the SAME inputs draw different keep-masks
    (different losses), and (2) the graph replays draw from the same stream as an eager loop started at the same position
    (losses bit-identical step by step: all kernels are deterministic and no parameter changes between steps)."""
    import mmsa
    from mmsa.step import TrainStep
    cfg = O.FusionConfig(embed_dim=768, num_heads=12, wiring="bidirectional", contract="single", valence=False)
    params, buffers = O.init_params(cfg, seed=2)
    inputs, labels = O.synth_inputs(cfg, 32, L=64, R=49, seed=11)

    def losses(use_graph: bool):
        model = build_model(cfg, params, buffers, torch.bfloat16, cuda_device).set_dropout(0.3)
        step = TrainStep(model, 32, 64, feature_dtype=torch.bfloat16, n_slots=1, use_graph=use_graph, device=cuda_device)
        s = step.slots[0]
        s.text.copy_(inputs[0]); s.image.copy_(inputs[1]); s.labels.copy_(labels)
        step.warmup(2)
        step.capture()
        model._drop.reseed(1234, position=0)            # both runs start from the same stream position
        torch.cuda.synchronize()
        out = []
        for _ in range(4):
            out.append(float(step.run(0).item()))
        pos = int(model._drop.state(cuda_device)[1].item())
        return out, pos

    eager, pos_e = losses(False)
    graph, pos_g = losses(True)
    assert len(set(eager)) == 4, f"eager steps drew identical masks: {eager}"
    assert len(set(graph)) == 4, f"graph replays drew identical masks (frozen RNG): {graph}"
    assert eager == graph, (eager, graph)
    assert pos_e == pos_g == 4 * 32 * (256 + 128 + 128)    # fusion.3, fusion.7, arousal_head.3 draws per step


def test_bn_act_dropout_keep_rate(cuda_device):
    """statistics of the fused BatchNorm + GELU + Philox dropout kernel (the path the model runs, not the stand-alone
    dropout kernel): keep-rate within 4 sigma of 1 - p, kept values scaled by 1/(1-p), masks of consecutive stream
    positions uncorrelated, and the device-resident state reproduces the by-value offset."""
    from mmsa import kernels as K
    from mmsa._lib import BN_THEN_GELU
    B, N, p = 2048, 256, 0.3
    x = torch.randn(B, N, device=cuda_device)
    gamma, beta = torch.ones(N, device=cuda_device), torch.zeros(N, device=cuda_device)
    state = torch.tensor([77, 0], dtype=torch.int64, device=cuda_device)
    y0, mean, rstd, _ = K.bn_act_fwd(x, gamma, beta, None, None, 0.1, 1e-5, True, BN_THEN_GELU, 0.0, None, 0, 0, torch.float32)
    y1, _, _, m1 = K.bn_act_fwd(x, gamma, beta, None, None, 0.1, 1e-5, True, BN_THEN_GELU, p, None, 0, 0, torch.float32, rng_state=state)
    K.rng_advance(state, B * N)
    y2, _, _, m2 = K.bn_act_fwd(x, gamma, beta, None, None, 0.1, 1e-5, True, BN_THEN_GELU, p, None, 0, 0, torch.float32, rng_state=state)
    y3, _, _, m3 = K.bn_act_fwd(x, gamma, beta, None, None, 0.1, 1e-5, True, BN_THEN_GELU, p, None, 77, B * N, torch.float32)
    assert torch.equal(m2, m3) and torch.equal(y2, y3)            # state {seed 77, position B*N} == by-value (77, B*N)
    n = B * N
    sigma = (p * (1 - p) / n) ** 0.5
    for m in (m1, m2):
        assert abs(float(m.float().mean()) - (1 - p)) <= 4 * sigma
    both = float((m1 & m2).float().mean())
    assert abs(both - (1 - p) ** 2) <= 6 * sigma                  # independent draws
    kept = m1.bool()
    assert torch.allclose(y1[kept], y0[kept] / (1 - p), rtol=1e-6, atol=1e-7) and float(y1[~kept].abs().max()) == 0.0
    col_rate = m1.float().mean(0)                                 # no column (feature unit) is systematically dropped
    assert float((col_rate - (1 - p)).abs().max()) <= 6 * (p * (1 - p) / B) ** 0.5


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 2e-2)])
def test_subnetwork_train_mode_dropout(cuda_device, dtype, tol):
    """TRAIN mode of mmsa.Subnetwork (MultimodalModel.py:89-95, TransformerEncoderLayer(dropout=0.3)): all four dropouts of
    a layer -- attention probabilities, dropout1, dropout, dropout2 -- with explicit keep masks, against the float64 oracle
    given the same masks (the oracle's mask positions are pinned against torch's own layer on the CPU)."""
    import mmsa
    torch.manual_seed(31)
    E, H, B, L, p = 256, 4, 5, 12, 0.3
    m = mmsa.Subnetwork(768, feat_dim=E, num_layers=2, nhead=H, compute_dtype=dtype)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    m = m.to(cuda_device).train()
    g = torch.Generator().manual_seed(8)
    keep = {}
    for i in range(2):
        pre = f"transformer.layers.{i}."
        keep[pre + "self_attn.dropout"] = torch.rand(B, H, L, L, generator=g) >= p
        keep[pre + "dropout1"] = torch.rand(B, L, E, generator=g) >= p
        keep[pre + "dropout"] = torch.rand(B * L, 3 * E, generator=g) >= p
        keep[pre + "dropout2"] = torch.rand(B, L, E, generator=g) >= p
    seen = []

    def provider(name, shape):
        seen.append(name)
        k = keep[name]
        assert tuple(k.shape) == tuple(shape) or k.numel() == int(torch.tensor(shape).prod()), (name, shape)
        return k.reshape(shape).to(torch.uint8).to(cuda_device).contiguous()
    m._drop.mask_provider = provider
    x = torch.randn(B, L, 768)
    wgt = torch.randn(B, L, E)
    xg = x.to(cuda_device).requires_grad_(True)
    y = m(xg)
    (y * wgt.to(cuda_device)).sum().backward()
    assert len(seen) == 8, seen
    sc = 1.0 / (1.0 - p)
    masks = [{"attn": keep[f"transformer.layers.{i}.self_attn.dropout"].double() * sc,
              "dropout1": keep[f"transformer.layers.{i}.dropout1"].double() * sc,
              "dropout": keep[f"transformer.layers.{i}.dropout"].double().view(B, L, 3 * E) * sc,
              "dropout2": keep[f"transformer.layers.{i}.dropout2"].double() * sc} for i in range(2)]
    p64 = {k: v.double().requires_grad_(k != "pos_encoder.pe") for k, v in sd.items()}
    x64 = x.double().requires_grad_(True)
    want = O.subnetwork(x64, p64, num_heads=H, masks=masks)
    (want * wgt.double()).sum().backward()
    assert rel_err(y, want.detach()) <= tol

    def ok(got, ref):
        if dtype == torch.float32:
            return rel_err(got, ref) <= 5e-5
        n_got, n_ref = float(got.double().norm()), float(ref.double().norm())
        return abs(n_got - n_ref) <= 5e-2 * n_ref and rel_err(got, ref) <= 2e-1
    assert ok(xg.grad, x64.grad)
    for k, prm in m.named_parameters():
        assert ok(prm.grad, p64[k].grad), k


def test_subnetwork_train_mode_philox_masks_consistent(cuda_device):
    """in-kernel Philox masks (no provider): the attention backward RE-DRAWS the probability mask from the stream position
    snapshotted at forward time.  With the stream reset to the same position before every forward the module is a fixed
    function of x, so a central finite difference along a random direction must match <grad, direction>; two forwards
    WITHOUT the reset differ (the position moves on); the keep rate of the output is sane."""
    import mmsa
    torch.manual_seed(7)
    m = mmsa.Subnetwork(64, feat_dim=64, num_layers=1, nhead=2).to(cuda_device).train()
    x = torch.randn(6, 9, 64, device=cuda_device, dtype=torch.float32)
    wgt = torch.randn(6, 9, 64, device=cuda_device)

    def f(xx):
        m._drop.reseed(99, position=5)
        return (m(xx) * wgt).sum()
    xg = x.clone().requires_grad_(True)
    f(xg).backward()
    d = torch.randn_like(x)
    d /= d.norm()
    eps = 1e-2
    with torch.no_grad():
        fd = (f(x + eps * d).double() - f(x - eps * d).double()) / (2 * eps)
    an = float((xg.grad.double() * d.double()).sum())
    assert abs(float(fd) - an) <= 2e-2 * max(abs(an), 1.0), (float(fd), an)
    with torch.no_grad():
        y1 = m(x)
        y2 = m(x)
    assert not torch.equal(y1, y2)
    m.eval()
    with torch.no_grad():
        assert torch.equal(m(x), m(x))


@pytest.mark.parametrize("E,H,Lq,Lk,dtype,tol", [(256, 4, 1, 1, torch.float32, 1e-5), (768, 12, 20, 49, torch.float32, 1e-5),
                                                  (768, 12, 20, 49, torch.bfloat16, 2e-2)])
def test_cross_modal_transformer_distinct_key_value(cuda_device, E, H, Lq, Lk, dtype, tol):
    """CrossModalTransformer.forward(query, key, value) with key IS NOT value (MultimodalModel.py:124; the signature allows
    it even though every reference call site passes one tensor): output and all gradients, incl. separate dkey / dvalue and
    the k / v row blocks of in_proj_weight, against the float64 oracle.  Native shape [B,E] (Lq = Lk = 1) and token shapes."""
    import mmsa
    torch.manual_seed(E + Lq)
    blk = mmsa.CrossModalTransformer(E, H)
    with torch.no_grad():
        blk.multihead_attn.in_proj_bias.normal_(0, 0.1)
    sd = {k: v.clone() for k, v in blk.state_dict().items()}
    blk = blk.to(cuda_device)
    B = 6
    shp = (lambda L: (B, E) if (Lq == 1 and Lk == 1) else (B, L, E))
    q, k, v = torch.randn(shp(Lq)), torch.randn(shp(Lk)), torch.randn(shp(Lk))
    wgt = torch.randn(shp(Lq))
    xs = [ops_cast(t, dtype, cuda_device) for t in (q, k, v)]
    y = blk(*xs)
    (y.float() * wgt.to(cuda_device)).sum().backward()
    p64 = {"b." + kk: vv.double().requires_grad_(True) for kk, vv in sd.items()}
    x64 = [t.to(dtype).double().requires_grad_(True) for t in (q, k, v)]          # the same (storage-rounded) inputs
    u = [t if t.ndim == 3 else t.unsqueeze(1) for t in x64]
    want = O.cross_block(u[0], u[1], p64, "b.", H, value=u[2])
    want = want if q.ndim == 3 else want.squeeze(1)
    (want * wgt.double()).sum().backward()
    assert rel_err(y, want.detach()) <= tol

    def ok(got, ref):
        if dtype == torch.float32:
            return rel_err(got, ref) <= 5e-5
        return rel_err(got, ref) <= 5e-2
    for name, a, b in zip(("dquery", "dkey", "dvalue"), xs, x64):
        if Lk == 1 and name == "dkey":      # soft-max over ONE key is 1 whatever the score: dL/dkey is exactly zero
            assert float(b.grad.abs().max()) == 0.0 and float(a.grad.abs().max()) <= 1e-6 * float(x64[2].grad.abs().max())
            continue
        assert ok(a.grad, b.grad), name
    for kk, prm in blk.named_parameters():
        ref = p64["b." + kk].grad
        if Lk == 1 and kk.startswith("multihead_attn.in_proj"):     # the q and k row blocks get exactly-zero gradients
            n3 = ref.shape[0] // 3
            assert float(prm.grad[:2 * n3].abs().max()) <= 1e-6 * float(ref.abs().max())
            assert ok(prm.grad[2 * n3:], ref[2 * n3:]), kk
            continue
        assert ok(prm.grad, ref), kk
    # the one-tensor call (key is value) still takes the packed K/V path and agrees with passing an equal COPY as value
    blk.zero_grad()
    y_same = blk(xs[0].detach(), xs[1].detach(), xs[1].detach())
    y_copy = blk(xs[0].detach(), xs[1].detach(), xs[1].detach().clone())
    assert rel_err(y_copy, y_same) <= (1e-6 if dtype == torch.float32 else 1e-2)


def ops_cast(t, dtype, device):
    return t.to(device).to(dtype).requires_grad_(True)
