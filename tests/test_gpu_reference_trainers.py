"""The reference's OWN, UNMODIFIED trainers driving the drop-in module on the GPU (-m gpu).

north_star: "keeps the reference's nn.Module forward/backward signatures so it drops into Trainer.py/Tester.py unchanged".
`__graft_entry__.build()` stages Trainer.py, Tester.py and dataLoader/MultiTaskTrainer.py from /root/reference into the
git-ignored baseline/_ref/ (which travels to the GPU box with the snapshot); this test imports them from there -- with empty
stand-ins for matplotlib / seaborn, which the box does not have and the code paths used here never call -- and runs
    Trainer.train_epoch / Trainer.test / Trainer.early_stop (Trainer.py:42-122: contract `model(eeg,eye,pps,labels) ->
        (logits, contrastive_loss)`, its own AdamW + add_param_group + clip_grad_norm_, torch.save of the state_dict),
    Tester.load_model / Tester.evaluate (Tester.py:29-68: `model(eeg,eye,pps) -> logits`, checkpoint round trip),
    MultiTaskTrainer phases (dataLoader/MultiTaskTrainer.py:179-233,347-467: 5-tuple contract, requires_grad toggling by
        sub-module name, per-phase optimisers)
over mmsa.FeatureBatches in the two batch formats.  Skipped when baseline/_ref is absent."""
import importlib.util
import os
import sys
import types

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref", "MML_ZYC")
have_ref = os.path.exists(os.path.join(REF, "Trainer.py"))


def _load(rel: str, name: str):
    for mod in ("matplotlib", "matplotlib.pyplot", "seaborn"):         # never called on the paths below
        if mod not in sys.modules:
            try:
                __import__(mod)
            except Exception:
                sys.modules[mod] = types.ModuleType(mod)
                if mod == "matplotlib":
                    sys.modules[mod].pyplot = types.ModuleType("matplotlib.pyplot")
                    sys.modules["matplotlib.pyplot"] = sys.modules[mod].pyplot
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _features(n: int, L: int, seed: int):
    """synthetic, LEARNABLE features: the class shifts a few text and image channels"""
    g = torch.Generator().manual_seed(seed)
    labels = torch.randint(0, 3, (n,), generator=g)
    text = torch.randn(n, L, 768, generator=g)
    image = torch.randn(n, 49, 2048, generator=g)
    text[:, :, :48] += (labels.float()[:, None, None] - 1.0) * 1.5
    image[:, :, :128] += (labels.float()[:, None, None] - 1.0) * 1.0
    return text, image, labels


@pytest.mark.skipif(not have_ref, reason="baseline/_ref not staged (run __graft_entry__.build() where /root/reference is mounted)")
def test_reference_trainer_and_tester_drive_the_dropin(cuda_device, tmp_path, monkeypatch):
    import mmsa
    T = _load("Trainer.py", "ref_Trainer")
    Te = _load("Tester.py", "ref_Tester")
    monkeypatch.chdir(tmp_path)                                  # Trainer.early_stop writes best_model.pth into the cwd
    torch.manual_seed(0)
    text, image, labels = _features(96, 32, 1)
    train = mmsa.FeatureBatches(text, image, labels, batch_size=32, fmt="dict")
    test = mmsa.FeatureBatches(*_features(64, 32, 2), batch_size=32, fmt="dict")
    model = mmsa.MultimodalTransformerModel(num_classes=3, embed_dim=768, num_heads=12, wiring="bidirectional",
                                            contract="single", compute_dtype=torch.bfloat16, temperature=0.07)
    trainer = T.Trainer(model, train, test, device="cuda")       # builds its own AdamW + add_param_group (Trainer.py:19-26)
    hist = []
    for epoch in range(1, 6):
        avg_loss, avg_ce, avg_con, acc = trainer.train_epoch(epoch)
        hist.append((avg_loss, avg_ce, avg_con, acc))
        assert all(map(lambda v: v == v and abs(v) != float("inf"), (avg_loss, avg_ce, avg_con)))
    assert hist[-1][0] < hist[0][0], f"training loss did not decrease over 5 epochs: {hist}"
    t_loss, t_ce, t_con, t_acc = trainer.test()                  # eval-mode call WITH labels (Trainer.py:128-170)
    assert t_loss == t_loss and 0.0 <= t_acc <= 1.0
    assert trainer.contrastive_weight.grad is not None
    assert trainer.early_stop(t_loss) is False and os.path.exists("best_model.pth")      # Trainer.py:108-111
    # Tester: fresh module, checkpoint through Tester.load_model (Tester.py:29-35), eval forward without labels (:53)
    fresh = mmsa.MultimodalTransformerModel(num_classes=3, embed_dim=768, num_heads=12, wiring="bidirectional",
                                            contract="single", compute_dtype=torch.bfloat16, temperature=0.07)
    tester = Te.Tester(fresh, test, device="cuda")
    tester.load_model("best_model.pth")
    for (k, a), (_, b) in zip(model.state_dict().items(), fresh.state_dict().items()):
        assert torch.equal(a, b), k
    res = tester.evaluate(verbose=False)
    assert res["predictions"].shape == (64,) and res["probabilities"].shape == (64, 3)
    assert abs(res["accuracy"] - t_acc) < 1e-9                   # the same weights, the same eval batches
    # a DataParallel-style checkpoint ('module.' prefix) goes through Tester.load_model's own stripping (Tester.py:32-33)
    torch.save({"module." + k: v for k, v in model.state_dict().items()}, "dp.pth")
    tester.load_model("dp.pth")
    one = tester.predict_single({"eeg": text[0], "eye": image[0], "pps": torch.zeros(1)})     # Tester.py:112-127
    assert one["probabilities"].shape == (3,)


@pytest.mark.skipif(not have_ref, reason="baseline/_ref not staged (run __graft_entry__.build() where /root/reference is mounted)")
def test_reference_multitask_trainer_drives_the_dropin(cuda_device, tmp_path, monkeypatch):
    """The live trainer (main.py:9,21): native three-modality wiring with mmsa.Subnetwork encoders in the eeg/eye/pps slots
    (the reference's eye_net / pps_net are Subnetworks, MultimodalModel.py:165-166), train mode -> all dropouts active."""
    import mmsa
    M = _load(os.path.join("dataLoader", "MultiTaskTrainer.py"), "ref_MultiTaskTrainer")
    monkeypatch.chdir(tmp_path)
    torch.manual_seed(1)
    g = torch.Generator().manual_seed(3)
    n = 96
    a_lab, v_lab = torch.randint(0, 3, (n,), generator=g), torch.randint(0, 3, (n,), generator=g)
    x0, x1, x2 = torch.randn(n, 64, generator=g), torch.randn(n, 38, generator=g), torch.randn(n, 230, generator=g)
    x0[:, :8] += (a_lab.float()[:, None] - 1) * 2.0
    x2[:, :16] += (v_lab.float()[:, None] - 1) * 2.0
    mk = lambda lo, hi: mmsa.FeatureBatches(x0[lo:hi], x1[lo:hi], a_lab[lo:hi], batch_size=32, third=x2[lo:hi],
                                            valence_labels=v_lab[lo:hi], fmt="tuple")
    train, test = mk(0, 64), mk(64, 96)
    enc = [mmsa.Subnetwork(64), mmsa.Subnetwork(38), mmsa.Subnetwork(230)]
    model = mmsa.MultimodalTransformerModel(num_classes=3, wiring="native", contract="multitask", encoders=enc)
    tr = M.MultiTaskTrainer(model, train, test, device="cuda")
    before = {k: v.detach().clone() for k, v in model.named_parameters()}
    m1 = tr.train_epoch_phase_eeg(1)            # only eeg_net trainable, loss = c_loss1 (MultiTaskTrainer.py:179-233)
    moved = {k for k, v in model.named_parameters() if not torch.equal(v.detach().cpu(), before[k].cpu())}
    assert moved and all(k.startswith("eeg_net.") for k in moved), sorted(moved)[:5]
    m2a = tr.train_epoch_phase2(1)              # fusion modules + arousal head, loss = CE(arousal) (:347-406)
    m2b = tr.train_epoch_phase2(2)
    m3 = tr.train_epoch_phase3(1)               # valence head (:408-467)
    ev = tr.evaluate("test")                    # eval forward with labels, 5-tuple (:469-515)
    for d in (m1, m2a, m2b, m3, ev):
        assert all(v == v and abs(v) != float("inf") for v in d.values()), d
    assert 0.0 <= ev["a_acc"] <= 1.0 and 0.0 <= ev["v_acc"] <= 1.0
    assert tr.early_stopping(ev["loss"]) is False and os.path.exists("best_model.pth")     # :517-527
    fresh = mmsa.MultimodalTransformerModel(num_classes=3, wiring="native", contract="multitask",
                                            encoders=[mmsa.Subnetwork(64), mmsa.Subnetwork(38), mmsa.Subnetwork(230)])
    fresh.load_state_dict(torch.load("best_model.pth"))
