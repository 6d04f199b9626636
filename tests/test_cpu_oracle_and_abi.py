"""CPU suite (-m "not gpu"): the oracle against the imported reference and the committed goldens,
the host-side contracts, and the C-ABI library (loads, exports every symbol of include/mmsa.h,
refuses to compute without a B200)."""
import os
import re
import sys

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from parity_util import O, rel_err

GOLD = os.path.join(os.path.dirname(__file__), "golden")
REF = "/root/reference/MML_ZYC"
has_ref = os.path.isdir(REF)


# ------------------------------------------------------------------ oracle vs the imported reference
@pytest.mark.skipif(not has_ref, reason="reference tree not mounted (GPU box)")
@pytest.mark.parametrize("B", [2, 20, 64])
def test_oracle_bit_identical_to_reference(B):
    """oracle.fusion_forward == reference MultimodalTransformerModel.forward, outputs and every
    gradient bit for bit, at the reference's native sizes (encoders = Identity, dropout p = 0)."""
    sys.path.insert(0, REF)
    import MultimodalModel as R
    torch.manual_seed(B)
    m = R.MultimodalTransformerModel()
    m.eeg_net, m.eye_net, m.pps_net = nn.Identity(), nn.Identity(), nn.Identity()
    for mod in m.modules():
        if isinstance(mod, nn.Dropout):
            mod.p = 0.0
    m.train()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    cfg = O.FusionConfig()
    xs, labels = O.synth_inputs(cfg, B, seed=B)
    a, v, c0, c1, c2 = m(*xs, labels=(labels, labels))
    (F.cross_entropy(a, labels) + F.cross_entropy(v, labels) + c0.sum() + c1.sum() + c2.sum()).backward()
    p = {k: t.clone().requires_grad_(t.dtype.is_floating_point) for k, t in sd.items()}
    out = O.fusion_forward(cfg, p, xs, labels, training=True)
    (F.cross_entropy(out.arousal, labels) + F.cross_entropy(out.valence, labels)
     + sum(c.sum() for c in out.contrastive)).backward()
    assert torch.equal(a, out.arousal) and torch.equal(v, out.valence)
    for x, y in zip((c0, c1, c2), out.contrastive):
        assert torch.equal(x, y)
    for k, prm in m.named_parameters():
        assert torch.equal(prm.grad, p[k].grad), k


@pytest.mark.skipif(not has_ref, reason="reference tree not mounted (GPU box)")
def test_oracle_block_matches_reference_with_long_keys():
    """Lq = 1, Lk = 49 is the one multi-token shape the reference block can run."""
    sys.path.insert(0, REF)
    import MultimodalModel as R
    torch.manual_seed(0)
    blk = R.CrossModalTransformer()
    q = torch.randn(5, 256)
    kv = torch.randn(5, 49, 256)
    ref = blk(q, kv, kv)
    p = {"b." + k: v for k, v in blk.state_dict().items()}
    got = O.cross_block(q.unsqueeze(1), kv, p, "b.", 4).squeeze(1)
    # 3-D key/value are passed through un-squeezed, so `key is value` holds and torch takes the packed
    # k,v in-projection (one [2E] GEMM instead of two): same arithmetic, different blocking -> ulp level
    assert rel_err(got, ref) <= 1e-6


# ------------------------------------------------------------------ oracle vs committed goldens
@pytest.mark.parametrize("case", ["native_case_B2_T0.01.pt", "native_case_B20_T0.5.pt", "native_case_B64_T0.07.pt"])
def test_oracle_reproduces_goldens(case):
    gp = torch.load(os.path.join(GOLD, "native_params.pt"))
    c = torch.load(os.path.join(GOLD, case))
    p = {k: t.clone().requires_grad_(t.dtype.is_floating_point) for k, t in gp["state_dict"].items()}
    with torch.no_grad():
        p["temperature"].fill_(c["temperature"])
    cfg = O.FusionConfig()
    out = O.fusion_forward(cfg, p, c["inputs"], c["labels"], training=True)
    loss = (F.cross_entropy(out.arousal, c["labels"]) + F.cross_entropy(out.valence, c["val_labels"])
            + sum(x.sum() for x in out.contrastive))
    loss.backward()
    assert rel_err(out.arousal, c["arousal"]) <= 1e-6 and rel_err(out.valence, c["valence"]) <= 1e-6
    assert rel_err(loss, c["loss"]) <= 1e-6
    for k, dig in c["grads"].items():
        flat = p[k].grad.reshape(-1)
        assert float((flat[dig["idx"]] - dig["vals"]).abs().max()) <= 1e-6 * max(dig["absmax"], 1e-12) * 10, k


def test_oracle_losses_reproduce_goldens():
    g = torch.load(os.path.join(GOLD, "memhacl.pt"))
    c = g["supcon"]
    assert rel_err(O.supcon(c["z1"], c["z2"], c["labels"], 0.1), c["loss"]) <= 1e-6
    c = g["ntxent"]
    assert rel_err(O.ntxent(c["z1"], c["z2"], 0.5), c["loss"]) <= 1e-6
    pm = {"multihead_attn." + k: v for k, v in g["mean"]["state_dict"].items()}
    assert rel_err(O.memhacl_fusion(g["feats"], pm, 8, "mean"), g["mean"]["out"]) <= 1e-6
    assert rel_err(O.memhacl_fusion(g["feats"], dict(g["max"]["state_dict"]), 8, "max"), g["max"]["out"]) <= 1e-6
    assert rel_err(O.projection_head(g["projection"]["x"], dict(g["projection"]["state_dict"])),
                   g["projection"]["out"]) <= 1e-6
    oa, ov = O.classifier(g["classifier"]["x"], dict(g["classifier"]["state_dict"]))
    assert rel_err(oa, g["classifier"]["out_a"]) <= 1e-6 and rel_err(ov, g["classifier"]["out_v"]) <= 1e-6


def test_oracle_sharded_infonce_equals_global():
    """row-block form (labels2 / row_offset) used under data parallelism == the reference's global form."""
    g = torch.Generator().manual_seed(0)
    f1, f2 = torch.randn(16, 32, generator=g, dtype=torch.float64), torch.randn(16, 32, generator=g, dtype=torch.float64)
    lab = torch.randint(0, 3, (16,), generator=g)
    T = torch.tensor(0.05, dtype=torch.float64)
    full = O.infonce(f1, f2, lab, T)
    parts = [O.infonce(f1[r * 4:(r + 1) * 4], f2, lab[r * 4:(r + 1) * 4], T, labels2=lab, row_offset=r * 4) for r in range(4)]
    assert abs(float(full) - float(sum(parts) / 4)) < 1e-12


def test_oracle_bidirectional_runs_at_baseline_shapes():
    cfg = O.FusionConfig(embed_dim=768, num_heads=12, wiring="bidirectional", contract="single", valence=False)
    params, _ = O.init_params(cfg, seed=0)
    xs, labels = O.synth_inputs(cfg, 2, L=64, R=49)
    loss, out = O.trainer_loss(cfg, params, xs, labels)
    assert out.arousal.shape == (2, 3) and torch.isfinite(loss)
    n = sum(v.numel() for k, v in params.items())
    assert n == 10_181_381 + 0 or n > 9_000_000     # ~10.2 M fusion parameters at E=768


# ------------------------------------------------------------------ C ABI
def test_library_loads_and_exports_every_header_symbol():
    from mmsa import _lib
    lib = _lib.load()
    names = _lib.header_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"libmmsa.so does not export {n}"
        assert n in _lib._PROTOS, f"no ctypes prototype for {n}"
    assert lib.mmsa_version().decode().startswith("mmsa-b200")


def test_prototype_arity_matches_header():
    from mmsa import _lib
    text = open(_lib.HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    for m in re.finditer(r"\b(mmsa_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        name, args = m.group(1), m.group(2).strip()
        n = 0 if args in ("", "void") else args.count(",") + 1
        assert len(_lib._PROTOS[name][1]) == n, f"{name}: header has {n} args, binding has {len(_lib._PROTOS[name][1])}"


def test_no_gpu_means_error_not_fallback():
    """Without a B200 the product path must fail loudly."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import mmsa
    from mmsa import _lib
    assert _lib.load().mmsa_check_device() != 0
    assert "sm_100" in _lib.load().mmsa_last_error().decode()
    model = mmsa.MultimodalTransformerModel()
    xs, labels = O.synth_inputs(O.FusionConfig(), 4)
    with pytest.raises(Exception):
        model(*xs, labels=(labels, labels))


def test_product_does_not_import_oracle():
    pkg = os.path.join(os.path.dirname(os.path.dirname(__file__)), "multimodal-sentiment-aanalysis_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f), encoding="utf-8").read()
                assert "fusion_oracle" not in src and "import oracle" not in src and "from oracle" not in src, f


def test_state_dict_keys_match_reference_layout():
    import mmsa
    gp = torch.load(os.path.join(GOLD, "native_params.pt"))
    model = mmsa.MultimodalTransformerModel()
    assert list(model.state_dict().keys()) == gp["keys"]
    for k, v in model.state_dict().items():
        assert tuple(v.shape) == tuple(gp["state_dict"][k].shape), k
    # module. prefix stripping of Tester.load_model (Tester.py:32-33) round-trips
    sd = {"module." + k: v for k, v in model.state_dict().items()}
    model.load_state_dict({k[7:]: v for k, v in sd.items()}, strict=True)


def test_io_adapters_batch_formats_and_checkpoints(tmp_path):
    """mmsa.io: the reference's two batch formats (data/Dataset.py:65-67 dict batches, dataLoader/DataLoader.py:152-156
    5-tuples) and checkpoint loading with the `module.` prefix of nn.DataParallel stripped (Tester.py:32-33)."""
    import mmsa
    n = 10
    text, image = torch.randn(n, 4, 768), torch.randn(n, 49, 2048)
    labels = torch.arange(n) % 3
    batches = list(mmsa.FeatureBatches(text, image, labels, batch_size=4, fmt="dict"))
    assert len(batches) == 3 and set(batches[0][0].keys()) == {"eeg", "eye", "pps"}
    assert batches[0][0]["eeg"].shape == (4, 4, 768) and batches[2][1].shape == (2,)
    assert torch.equal(torch.cat([b[1] for b in batches]), labels)
    tup = next(iter(mmsa.FeatureBatches(text, image, labels, batch_size=4, fmt="tuple", valence_labels=labels.flip(0))))
    assert len(tup) == 5 and tup[3].dtype == torch.int64 and torch.equal(tup[4], labels.flip(0)[:4])
    shuffled = mmsa.FeatureBatches(text, image, labels, batch_size=4, fmt="tuple", shuffle=True, seed=1, drop_last=True)
    assert len(shuffled) == 2 and len(list(shuffled)) == 2
    # checkpoint written from a DataParallel-wrapped model: keys carry "module."
    src = mmsa.MultimodalTransformerModel()
    sd = {"module." + k: v.clone() + 1 for k, v in src.state_dict().items()}
    path = str(tmp_path / "best_model.pth")
    torch.save(sd, path)                                              # Trainer.py:111
    dst = mmsa.MultimodalTransformerModel()
    res = mmsa.load_reference_state_dict(dst, path)
    assert not res.missing_keys and not res.unexpected_keys
    for k, v in dst.state_dict().items():
        assert torch.equal(v, sd["module." + k])
    # a reference checkpoint also holds the out-of-scope encoder weights: dropped on request
    sd2 = dict(sd)
    sd2["module.eeg_net.conv1.weight"] = torch.zeros(3)
    mmsa.load_reference_state_dict(dst, sd2, ignore_prefixes=("eeg_net.", "eye_net.", "pps_net."))


def test_oracle_subnetwork_reproduces_golden_and_torch():
    """Encoder tail: the oracle restatement of Subnetwork (MultimodalModel.py:83-105) against the
    golden generated from the imported reference, and -- for L > 1, where the reference is never run -- against torch's own
    nn.TransformerEncoder built from the same parameters."""
    g = torch.load(os.path.join(GOLD, "subnetwork.pt"))
    p = {k: v.clone().requires_grad_(v.is_floating_point() and k != "pos_encoder.pe") for k, v in g["state_dict"].items()}
    x = g["x"].clone().requires_grad_(True)
    y = O.subnetwork(x, p)
    (y * g["wgt"]).sum().backward()
    assert rel_err(y, g["out"]) <= 1e-6 and rel_err(x.grad, g["dx"]) <= 1e-5
    for k, dig in g["grads"].items():
        flat = p[k].grad.reshape(-1)
        assert float((flat[dig["idx"]] - dig["vals"]).abs().max()) <= 2e-5 * max(dig["absmax"], 1e-12), k
    # L = 7 tokens against torch.nn.TransformerEncoder (eval mode: dropout off)
    import torch.nn as nn
    torch.manual_seed(3)
    layer = nn.TransformerEncoderLayer(d_model=64, nhead=4, dim_feedforward=192, dropout=0.3, batch_first=True)
    enc = nn.TransformerEncoder(layer, 2, enable_nested_tensor=False).eval()
    proj, norm = nn.Linear(10, 64), nn.LayerNorm(64)
    sd = {"proj.weight": proj.weight, "proj.bias": proj.bias, "norm.weight": norm.weight, "norm.bias": norm.bias}
    sd.update({"transformer." + k: v for k, v in enc.state_dict().items()})
    sd = {k: v.detach().clone() for k, v in sd.items()}
    xs = torch.randn(3, 7, 10)
    want = norm(enc(proj(xs) + O.positional_table(64, 100)[:7]))
    got = O.subnetwork(xs, sd)
    assert rel_err(got, want.detach()) <= 1e-5


@pytest.mark.skipif(not has_ref, reason="reference tree not mounted (GPU box)")
def test_oracle_subnetwork_bit_identical_to_reference():
    sys.path.insert(0, REF)
    import MultimodalModel as R
    torch.manual_seed(5)
    m = R.Subnetwork(230).eval()
    x = torch.randn(9, 230)
    with torch.no_grad():
        want = m(x)
    got = O.subnetwork(x, {k: v for k, v in m.state_dict().items()})
    assert rel_err(got, want) <= 1e-6


def test_oracle_subnetwork_train_mode_dropout_positions(monkeypatch):
    """TRAIN mode of the encoder tail: torch's nn.TransformerEncoderLayer(dropout=0.3) (what Subnetwork builds,
    MultimodalModel.py:89-95) drops in four places per layer -- the attention probabilities, dropout1, dropout, dropout2.
    torch draws those masks from its own generator, so the pin intercepts the two functions the third-party layer calls
    (F.dropout and F.scaled_dot_product_attention) and feeds them recorded masks; the oracle given the same masks must
    reproduce the layer's output and gradients.  (The imported reference Subnetwork is used when the tree is mounted.)"""
    import torch.nn as nn
    import torch.nn.functional as Fn
    torch.manual_seed(11)
    E, H, L, Bn, p_drop = 64, 4, 5, 3, 0.3
    layer = nn.TransformerEncoderLayer(d_model=E, nhead=H, dim_feedforward=3 * E, dropout=p_drop, batch_first=True)
    enc = nn.TransformerEncoder(layer, 2, enable_nested_tensor=False).train()
    proj, norm = nn.Linear(10, E), nn.LayerNorm(E)
    g = torch.Generator().manual_seed(2)
    masks = []
    for _ in range(2):
        masks.append({"attn": (torch.rand(Bn, H, L, L, generator=g) >= p_drop).float() / (1 - p_drop),
                      "dropout1": (torch.rand(Bn, L, E, generator=g) >= p_drop).float() / (1 - p_drop),
                      "dropout": (torch.rand(Bn, L, 3 * E, generator=g) >= p_drop).float() / (1 - p_drop),
                      "dropout2": (torch.rand(Bn, L, E, generator=g) >= p_drop).float() / (1 - p_drop)})
    calls = {"drop": [], "sdpa": 0}
    order = ["dropout1", "dropout", "dropout2"]          # _sa_block's dropout1, then _ff_block's dropout and dropout2

    def fake_dropout(x, p=0.5, training=True, inplace=False):
        if not training or p == 0:
            return x
        i = len(calls["drop"])
        # call order inside one layer: dropout1 (after self_attn), dropout (after relu(linear1)), dropout2 (after linear2)
        name = {0: "dropout1", 1: "dropout", 2: "dropout2"}[i % 3]
        calls["drop"].append(name)
        return x * masks[i // 3][name].to(x.dtype)

    def fake_sdpa(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False, **kw):
        i = calls["sdpa"]
        calls["sdpa"] += 1
        assert abs(dropout_p - p_drop) < 1e-12 and attn_mask is None        # training mode hands the rate to SDPA
        w = torch.softmax(q @ k.transpose(-2, -1) / math.sqrt(q.shape[-1]), dim=-1)
        return (w * masks[i]["attn"].to(w.dtype)) @ v

    import math
    monkeypatch.setattr(Fn, "dropout", fake_dropout)
    monkeypatch.setattr(Fn, "scaled_dot_product_attention", fake_sdpa)
    xs = torch.randn(Bn, L, 10).requires_grad_(True)
    wgt = torch.randn(Bn, L, E)
    want = norm(enc(proj(xs) + O.positional_table(E, 100)[:L]))
    (want * wgt).sum().backward()
    monkeypatch.undo()
    assert calls["sdpa"] == 2 and calls["drop"] == order * 2
    sd = {"proj.weight": proj.weight, "proj.bias": proj.bias, "norm.weight": norm.weight, "norm.bias": norm.bias}
    sd.update({"transformer." + k: v for k, v in enc.state_dict().items()})
    pp = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    x2 = xs.detach().clone().requires_grad_(True)
    got = O.subnetwork(x2, pp, num_heads=H, masks=masks)
    (got * wgt).sum().backward()
    assert rel_err(got, want.detach()) <= 1e-5
    assert rel_err(x2.grad, xs.grad) <= 1e-5
    named = dict(proj.named_parameters(prefix="proj")) | dict(norm.named_parameters(prefix="norm")) | \
        dict(enc.named_parameters(prefix="transformer"))
    for k, prm in named.items():
        assert rel_err(pp[k].grad, prm.grad) <= 2e-5, k


def test_oracle_sharded_supcon_ntxent_equal_global():
    """row-block forms (two ranks, rows sharded, both views gathered) reproduce train.py:16-40 / ME-MHACL/train.py:47-66."""
    g = torch.Generator().manual_seed(4)
    Bg, D, world = 12, 16, 2
    B = Bg // world
    z1, z2 = torch.randn(Bg, D, generator=g, dtype=torch.float64), torch.randn(Bg, D, generator=g, dtype=torch.float64)
    labels = torch.randint(0, 3, (Bg,), generator=g)
    z_all = torch.cat([z1, z2])
    lab_all = torch.cat([labels, labels])
    tot_s = tot_n = 0.0
    for r in range(world):
        sl = slice(r * B, (r + 1) * B)
        tot_s = tot_s + O.supcon_rows(z1[sl], z_all, labels[sl], lab_all, r * B, 0.1) \
                      + O.supcon_rows(z2[sl], z_all, labels[sl], lab_all, Bg + r * B, 0.1)
        tot_n = tot_n + O.ntxent_rows(z1[sl], z_all, r * B, 0.5) + O.ntxent_rows(z2[sl], z_all, Bg + r * B, 0.5)
    assert abs(float(tot_s) / (2 * Bg) - float(O.supcon(z1, z2, labels, 0.1))) < 1e-12
    assert abs(float(tot_n) / (2 * Bg) - float(O.ntxent(z1, z2, 0.5))) < 1e-12


def test_sass_shows_the_blackwell_paths():
    """The built library really contains the hardware paths DESIGN.md claims (SASS mnemonics, B200_PROFILING.md):
    tcgen05 MMA + TMEM loads + TMA in the GEMM and attention kernels, bulk copies in the row kernels, mma.sync +
    cp.async + cluster barriers in the small-product kernel.  Guards against a silent fall-back after a refactor."""
    import re
    import shutil
    import subprocess
    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "multimodal-sentiment-aanalysis_b200",
                      "mmsa", "libmmsa.so")
    if shutil.which("cuobjdump") is None or not os.path.exists(so):
        pytest.skip("cuobjdump or the built library is not available")
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    per = {}
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = set()
            continue
        if cur is not None:
            for mn in re.findall(r"\b(UTC\w*MMA|LDTM|UTMALDG|UTMASTG|UBLKCP|HMMA|LDGSTS|UCGABAR_ARV)\b", line):
                per[cur].add("UTCMMA" if mn.startswith("UTC") else mn)
    def have(sub):
        ks = [v for k, v in per.items() if sub in k]
        assert ks, f"no kernel named *{sub}* in the library"
        return ks
    for v in have("gemm_tcgen05_kernel"):
        assert {"UTCMMA", "LDTM", "UTMALDG", "UTMASTG"} <= v, v
    for name in ("attn_fwd_tc_kernel", "attn_bwd_tc_kernel"):
        for v in have(name):
            assert {"UTCMMA", "LDTM", "UTMALDG", "UTMASTG"} <= v, (name, v)
    for name in ("gate_ln_pool_fwd_kernel", "gate_ln_pool_bwd_kernel"):
        for v in have(name):
            assert "UBLKCP" in v, (name, v)
    for v in have("gemm_small_kernel"):
        assert {"HMMA", "LDGSTS", "UCGABAR_ARV"} <= v, v
