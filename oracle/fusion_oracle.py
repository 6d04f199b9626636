"""CPU oracle for the fusion / ME-MHACL / contrastive hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(`multimodal-sentiment-aanalysis_b200/`) may import this file; only `tests/`,
`__graft_entry__.smoke()` and the baseline legs of `bench.py` (`cpu_baseline`,
`--impl reference`, and `torch_eager`: the same functions on `cuda`, i.e. PyTorch
eager running the reference arithmetic on the B200 as the kernel-to-beat) use it,
and only as the checker / baseline -- never on the measured product path.

It is a plain-PyTorch (CPU, fp32 or fp64) functional restatement of the
reference arithmetic, parametrised in (E, H, L, R, input widths, wiring) so
that it also runs at BASELINE.json's text/image sizes where the reference's
own `CrossModalTransformer.forward` cannot (its gate concatenates on dim=1,
MultimodalModel.py:147, which is the feature axis only when Lq == 1).

Parity pin: the reference holds no golden vectors or tests for this path
(SURVEY.md section 4), so the oracle is pinned against the *imported reference itself*
at the reference's native sizes (tests/test_cpu_oracle_and_abi.py::
test_oracle_bit_identical_to_reference, run where /root/reference is mounted),
against committed fixtures generated from the imported reference by
oracle/make_goldens.py (tests/golden/*.pt), and -- for the training-mode dropout
positions of the encoder layer -- against torch's own nn.TransformerEncoderLayer
with intercepted masks (test_oracle_subnetwork_train_mode_dropout_positions).

Every function cites the reference lines it follows
(paths relative to /root/reference/MML_ZYC/).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Params = Dict[str, Tensor]


# --------------------------------------------------------------------------
# torch.nn.MultiheadAttention, need_weights=True branch
# (third-party: torch/nn/functional.py multi_head_attention_forward, reached from
#  MultimodalModel.py:139-143 and ME-MHACL/model.py:71 / MultimodalModel.py:397)
# --------------------------------------------------------------------------
def mha_core(q: Tensor, k: Tensor, v: Tensor, num_heads: int, prob_mask: Optional[Tensor] = None) -> Tensor:
    """q:[Lq,B,E] k,v:[Lk,B,E] (already projected, seq-first) -> [Lq*B, E] pre-out-proj.
    prob_mask: pre-scaled keep mask [B*H, Lq, Lk] multiplied onto the soft-max probabilities (training-mode
    MultiheadAttention(dropout=p): torch applies dropout to the attention weights, functional.py
    multi_head_attention_forward -> scaled_dot_product_attention(dropout_p))."""
    Lq, B, E = q.shape
    Lk = k.shape[0]
    d = E // num_heads
    qh = q.reshape(Lq, B * num_heads, d).transpose(0, 1)
    kh = k.reshape(Lk, B * num_heads, d).transpose(0, 1)
    vh = v.reshape(Lk, B * num_heads, d).transpose(0, 1)
    # scale is applied to q *before* the bmm (functional.py need_weights branch)
    q_scaled = qh * math.sqrt(1.0 / float(d))
    w = torch.bmm(q_scaled, kh.transpose(-2, -1))
    w = torch.softmax(w, dim=-1)
    if prob_mask is not None:
        w = w * prob_mask
    o = torch.bmm(w, vh)
    return o.transpose(0, 1).contiguous().view(Lq * B, E)


def mha_cross(query: Tensor, key: Tensor, value: Tensor, in_w: Tensor, in_b: Tensor,
              out_w: Tensor, out_b: Tensor, num_heads: int) -> Tensor:
    """batch_first cross attention as nn.MultiheadAttention(E,H,batch_first=True) runs it
    when called from CrossModalTransformer.forward (MultimodalModel.py:139-143): key and
    value are distinct tensor objects after the unsqueeze at :134-137, so torch takes the
    three-way `w.chunk(3)` in-projection."""
    B, Lq, E = query.shape
    q = query.transpose(1, 0)
    k = key.transpose(1, 0)
    v = value.transpose(1, 0)
    w_q, w_k, w_v = in_w.chunk(3)
    b_q, b_k, b_v = in_b.chunk(3)
    qp = F.linear(q, w_q, b_q)
    kp = F.linear(k, w_k, b_k)
    vp = F.linear(v, w_v, b_v)
    o = mha_core(qp, kp, vp, num_heads)
    o = F.linear(o, out_w, out_b).view(Lq, B, E)
    return o.transpose(1, 0)


def mha_self_seq_first(x: Tensor, in_w: Tensor, in_b: Tensor, out_w: Tensor, out_b: Tensor,
                       num_heads: int, prob_mask: Optional[Tensor] = None) -> Tensor:
    """self attention, seq-first [L,B,E] (ME-MHACL/model.py:71, MultimodalModel.py:397):
    packed in-projection `linear(x, W, b)` then split in q,k,v order."""
    L, B, E = x.shape
    proj = F.linear(x, in_w, in_b)
    proj = proj.unflatten(-1, (3, E)).unsqueeze(0).transpose(0, -2).squeeze(-2).contiguous()
    o = mha_core(proj[0], proj[1], proj[2], num_heads, prob_mask)
    return F.linear(o, out_w, out_b).view(L, B, E)


# --------------------------------------------------------------------------
# CrossModalTransformer (MultimodalModel.py:108-149)
# --------------------------------------------------------------------------
def cross_block(query: Tensor, kv: Tensor, p: Params, prefix: str, num_heads: int,
                value: Optional[Tensor] = None) -> Tensor:
    """query:[B,Lq,E] kv:[B,Lk,E] -> [B,Lq,E]  (value: a separate value tensor, forward(query, key, value) :124;
    every reference call site passes the key tensor again, :287-297).
    MHA (:139-143) -> gate = sigmoid(Linear(2E,E)(cat[q, attn])) (:147, restated with the
    concat on the LAST dim so Lq > 1 works; identical when Lq == 1) ->
    g*q + (1-g)*attn (:148) -> LayerNorm(E) (:149)."""
    attn = mha_cross(query, kv, kv if value is None else value,
                     p[prefix + "multihead_attn.in_proj_weight"], p[prefix + "multihead_attn.in_proj_bias"],
                     p[prefix + "multihead_attn.out_proj.weight"], p[prefix + "multihead_attn.out_proj.bias"],
                     num_heads)
    gate = torch.sigmoid(F.linear(torch.cat([query, attn], dim=-1),
                                  p[prefix + "gate.0.weight"], p[prefix + "gate.0.bias"]))
    out = gate * query + (1 - gate) * attn
    E = query.shape[-1]
    return F.layer_norm(out, (E,), p[prefix + "norm.weight"], p[prefix + "norm.bias"], 1e-5)


# --------------------------------------------------------------------------
# Linear -> BatchNorm1d -> GELU -> Dropout chains
# (fusion :179-189, arousal_head :192-199, valence_head :200-225)
# and Linear -> ReLU -> BatchNorm1d -> Dropout chains (ProjectionHead, ME-MHACL/model.py:82-97)
# --------------------------------------------------------------------------
def batchnorm1d(x: Tensor, p: Params, prefix: str, training: bool, eps: float = 1e-5,
                momentum: float = 0.1, buffers: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """nn.BatchNorm1d on [B,N] through the same third-party op the reference reaches
    (torch.nn.functional.batch_norm): train mode = batch mean / biased variance
    y = (x-mean)/sqrt(var+eps)*w+b; running stats (when `buffers` is given) updated with the
    UNBIASED variance and momentum 0.1; eval mode = running stats."""
    w, b = p[prefix + "weight"], p[prefix + "bias"]
    src = buffers if buffers is not None else p
    rm = src.get(prefix + "running_mean")
    rv = src.get(prefix + "running_var")
    if training:
        if buffers is not None:
            buffers[prefix + "num_batches_tracked"] += 1
            return F.batch_norm(x, rm, rv, w, b, True, momentum, eps)
        return F.batch_norm(x, None, None, w, b, True, momentum, eps)
    return F.batch_norm(x, rm, rv, w, b, False, momentum, eps)


def mlp_chain(x: Tensor, p: Params, prefix: str, plan: Sequence[Tuple[str, int]], training: bool,
              dropout_masks: Optional[List[Tensor]] = None,
              buffers: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """Run an nn.Sequential restated as a plan of (kind, index) steps, kind in
    {linear, bn, gelu, relu, dropout}.  GELU is the exact-erf form (nn.GELU() default).
    Dropout: identity unless `dropout_masks` supplies pre-scaled keep masks (one per dropout)."""
    mi = 0
    for kind, idx in plan:
        key = f"{prefix}{idx}."
        if kind == "linear":
            x = F.linear(x, p[key + "weight"], p[key + "bias"])
        elif kind == "bn":
            x = batchnorm1d(x, p, key, training, buffers=buffers)
        elif kind == "gelu":
            x = F.gelu(x)
        elif kind == "relu":
            x = F.relu(x)
        elif kind == "dropout":
            if dropout_masks is not None and training:
                x = x * dropout_masks[mi]
            mi += 1
        else:
            raise ValueError(kind)
    return x


FUSION_PLAN = [("linear", 0), ("bn", 1), ("gelu", 2), ("dropout", 3),
               ("linear", 4), ("bn", 5), ("gelu", 6), ("dropout", 7)]
AROUSAL_PLAN = [("linear", 0), ("bn", 1), ("gelu", 2), ("dropout", 3), ("linear", 4)]
VALENCE_PLAN = [("linear", 0), ("bn", 1), ("gelu", 2), ("dropout", 3),
                ("linear", 4), ("bn", 5), ("gelu", 6), ("dropout", 7),
                ("linear", 8), ("bn", 9), ("gelu", 10), ("dropout", 11),
                ("linear", 12), ("bn", 13), ("gelu", 14), ("dropout", 15),
                ("linear", 16)]
PROJECTION_PLAN = [("linear", 0), ("relu", 1), ("bn", 2), ("dropout", 3),
                   ("linear", 4), ("relu", 5), ("bn", 6), ("dropout", 7), ("linear", 8)]


# --------------------------------------------------------------------------
# Contrastive losses
# --------------------------------------------------------------------------
def infonce(feat1: Tensor, feat2: Tensor, labels: Tensor, temperature: Tensor,
            labels2: Optional[Tensor] = None, row_offset: int = 0) -> Tensor:
    """MultimodalTransformerModel.compute_contrastive_loss (MultimodalModel.py:232-260).
    `labels2` / `row_offset` generalise it to a row block of a batch-sharded similarity
    matrix: rows are local samples (global index row_offset+i), columns the gathered
    global batch; the reference's fill_diagonal_(0) (:241) then clears the GLOBAL diagonal."""
    f1 = F.normalize(feat1, dim=1)                                   # :234
    f2 = F.normalize(feat2, dim=1)                                   # :235
    sim = torch.mm(f1, f2.t()) / temperature                         # :237
    l2 = labels if labels2 is None else labels2
    pos_mask = torch.eq(labels.unsqueeze(1), l2.unsqueeze(0)).to(sim.dtype)   # :240
    n_rows = pos_mask.shape[0]
    idx = torch.arange(n_rows)
    cols = idx + row_offset
    ok = cols < pos_mask.shape[1]
    pos_mask[idx[ok], cols[ok]] = 0                                  # :241 fill_diagonal_(0)
    sim = sim - torch.max(sim, dim=1, keepdim=True)[0]               # :245 (grad flows through max)
    exp_sim = torch.exp(sim)                                         # :248
    pos_sim = (exp_sim * pos_mask).sum(1)                            # :251
    all_sim = exp_sim.sum(1)                                         # :254
    loss = -torch.log((pos_sim + 1e-12) / (all_sim + 1e-12))         # :257
    return loss.mean()                                               # :260


def supcon(z1: Tensor, z2: Tensor, labels: Tensor, temperature: float = 0.1) -> Tensor:
    """train.py:16-40 contrastive_loss (SupCon over the 2B stacked views, no max subtraction)."""
    z1 = F.normalize(z1, dim=1)
    z2 = F.normalize(z2, dim=1)
    z = torch.cat([z1, z2], dim=0)
    sim = torch.matmul(z, z.T) / temperature
    lab = labels.view(-1, 1)
    lab = torch.cat([lab, lab], dim=0)
    mask = torch.eq(lab, lab.T).to(sim.dtype)
    self_mask = torch.eye(mask.size(0), dtype=torch.bool)
    mask = mask.masked_fill(self_mask, 0)
    sim_exp = torch.exp(sim)
    sim_exp = sim_exp.masked_fill(self_mask, 0)
    sim_sum = sim_exp.sum(dim=1, keepdim=True)
    log_prob = sim - torch.log(sim_sum + 1e-8)
    loss = -(mask * log_prob).sum(dim=1) / (mask.sum(dim=1) + 1e-8)
    return loss.mean()


def ntxent(z1: Tensor, z2: Tensor, temperature: float = 0.5) -> Tensor:
    """ME-MHACL/train.py:47-66 contrastive_loss (SimCLR NT-Xent)."""
    n = z1.size(0)
    z = torch.cat([z1, z2], dim=0)
    z = F.normalize(z, dim=1)
    sim = torch.matmul(z, z.T)
    mask = torch.eye(2 * n, dtype=torch.bool)
    sim = sim.masked_fill(mask, -9e15)
    sim = sim / temperature
    targets = torch.cat([torch.arange(n, 2 * n), torch.arange(0, n)], dim=0)
    return F.cross_entropy(sim, targets)


def supcon_rows(z_rows: Tensor, z_all: Tensor, labels_rows: Tensor, labels_all: Tensor, row_offset: int,
                temperature: float = 0.1) -> Tensor:
    """Row block of train.py:16-40 under batch sharding: `z_all` [2Bg, D] is the stacked global batch (all first
    views, then all second views), `z_rows` the local rows whose global indices start at `row_offset`.  Returns the SUM
    of the row losses (the caller divides by the number of rows it averages over)."""
    zr = F.normalize(z_rows, dim=1)
    za = F.normalize(z_all, dim=1)
    sim = torch.matmul(zr, za.T) / temperature
    mask = torch.eq(labels_rows.view(-1, 1), labels_all.view(1, -1)).to(sim.dtype)
    self_mask = torch.zeros_like(mask, dtype=torch.bool)
    idx = torch.arange(z_rows.shape[0])
    self_mask[idx, idx + row_offset] = True
    mask = mask.masked_fill(self_mask, 0)
    sim_exp = torch.exp(sim).masked_fill(self_mask, 0)
    log_prob = sim - torch.log(sim_exp.sum(dim=1, keepdim=True) + 1e-8)
    return (-(mask * log_prob).sum(dim=1) / (mask.sum(dim=1) + 1e-8)).sum()


def ntxent_rows(z_rows: Tensor, z_all: Tensor, row_offset: int, temperature: float = 0.5) -> Tensor:
    """Row block of ME-MHACL/train.py:47-66 under batch sharding (same conventions as supcon_rows): cross-entropy of each
    local row against its partner view at global index (i + Bg) mod 2Bg, diagonal masked.  Returns the SUM over rows."""
    zr = F.normalize(z_rows, dim=1)
    za = F.normalize(z_all, dim=1)
    sim = torch.matmul(zr, za.T)
    n2 = z_all.shape[0]
    idx = torch.arange(z_rows.shape[0])
    gi = idx + row_offset
    mask = torch.zeros_like(sim, dtype=torch.bool)
    mask[idx, gi] = True
    sim = sim.masked_fill(mask, -9e15) / temperature
    return F.cross_entropy(sim, (gi + n2 // 2) % n2, reduction="sum")


# --------------------------------------------------------------------------
# MultimodalTransformerModel.forward hot path (MultimodalModel.py:262-322)
# --------------------------------------------------------------------------
@dataclass
class FusionConfig:
    """Shape/wiring description of one instance of the hot path.

    wiring == "native":  the reference's own wiring (three modality feature vectors [B,E];
        query is always modality 0, :287-297; three self-contrast losses, :271-284).
    wiring == "bidirectional": BASELINE.json's text+image re-skin (SURVEY.md section 8(d) canonical
        composition): text tokens [B,L,D_t] and image regions [B,R,D_i] are projected to E
        (Subnetwork.proj pattern, :86), block e2p is query=text/kv=image, block p2e is
        query=image/kv=text, token mean-pool (:76 pattern) feeds the three fused slots,
        modality weights read the RAW pooled features (:299-301), InfoNCE(f1,f2,labels)."""
    embed_dim: int = 256
    num_heads: int = 4
    num_classes: int = 3
    wiring: str = "native"
    text_dim: int = 768
    image_dim: int = 2048
    contract: str = "multitask"      # "multitask" = 5-tuple (MultiTaskTrainer.py:199), "single" = Trainer.py:60
    valence: bool = True


@dataclass
class FusionOutputs:
    arousal: Tensor
    valence: Optional[Tensor]
    contrastive: List[Tensor]
    feats: Dict[str, Tensor] = field(default_factory=dict)


def fusion_forward(cfg: FusionConfig, p: Params, inputs: Sequence[Tensor],
                   labels: Optional[Tensor] = None, training: bool = True,
                   dropout_masks: Optional[Dict[str, List[Tensor]]] = None,
                   buffers: Optional[Dict[str, Tensor]] = None,
                   gathered: Optional[Tuple[Tensor, Tensor, int]] = None) -> FusionOutputs:
    """The hot path of MultimodalTransformerModel.forward, starting from per-modality FEATURES
    (the modality encoders eeg_net/eye_net/pps_net, :264-266, are out of scope).

    native:        inputs = (f_eeg[B,E], f_eye[B,E], f_pps[B,E])
    bidirectional: inputs = (text[B,L,D_t], image[B,R,D_i])
    labels: arousal labels [B] int64 (the reference contrasts on labels[0], :273).
    gathered: optional (f2_global, labels_global, row_offset) for the batch-sharded InfoNCE."""
    H = cfg.num_heads
    dm = dropout_masks or {}
    out_feats: Dict[str, Tensor] = {}
    contrastive: List[Tensor] = []
    cw = p["contrastive_weight"]
    T = p["temperature"]

    if cfg.wiring == "native":
        f0, f1r, f2r = [x if x.ndim == 2 else x.squeeze(1) for x in inputs]
        if labels is not None:                                       # :271-284
            for f in (f0, f1r, f2r):
                contrastive.append(infonce(f, f, labels, T))
        e1 = cross_block(f0.unsqueeze(1), f1r.unsqueeze(1), p, "cross_attn_e2p.", H).squeeze(1)   # :287
        e2 = cross_block(f0.unsqueeze(1), f2r.unsqueeze(1), p, "cross_attn_p2e.", H).squeeze(1)   # :293
        raw = torch.cat([f0, f1r, f2r], dim=1)                       # :300
        slots = (f0, e1, e2)
    elif cfg.wiring == "bidirectional":
        text, image = inputs[0], inputs[1]
        t = F.linear(text, p["eeg_net.proj.weight"], p["eeg_net.proj.bias"])      # :86 pattern
        v = F.linear(image, p["eye_net.proj.weight"], p["eye_net.proj.bias"])
        t2 = cross_block(t, v, p, "cross_attn_e2p.", H)              # text <- image
        v2 = cross_block(v, t, p, "cross_attn_p2e.", H)              # image <- text
        f0, fv = t.mean(1), v.mean(1)                                # :76 pooling pattern
        e1, e2 = t2.mean(1), v2.mean(1)
        raw = torch.cat([f0, fv], dim=1)
        slots = (f0, e1, e2)
        if labels is not None:
            if gathered is None:
                contrastive.append(infonce(e1, e2, labels, T))
            else:
                g2, gl, off = gathered
                contrastive.append(infonce(e1, g2, labels, T, labels2=gl, row_offset=off))
    else:
        raise ValueError(cfg.wiring)
    out_feats["slot0"], out_feats["slot1"], out_feats["slot2"] = slots

    h = F.gelu(F.linear(raw, p["attention_weights.0.weight"], p["attention_weights.0.bias"]))     # :171-176
    w = torch.softmax(F.linear(h, p["attention_weights.2.weight"], p["attention_weights.2.bias"]), dim=1)
    fused = torch.cat([slots[0] * w[:, 0:1], slots[1] * w[:, 1:2], slots[2] * w[:, 2:3]], dim=1)  # :302-306
    out_feats["weights"] = w
    fused = mlp_chain(fused, p, "fusion.", FUSION_PLAN, training, dm.get("fusion"), buffers)      # :309
    out_feats["fused"] = fused
    arousal = mlp_chain(fused, p, "arousal_head.", AROUSAL_PLAN, training, dm.get("arousal_head"), buffers)  # :312
    valence = None
    if cfg.valence:
        valence = mlp_chain(fused, p, "valence_head.", VALENCE_PLAN, training, dm.get("valence_head"), buffers)  # :313
    contrastive = [cw * c for c in contrastive]                      # :315-317 -> shape (1,)
    return FusionOutputs(arousal, valence, contrastive, out_feats)


def trainer_loss(cfg: FusionConfig, p: Params, inputs, labels, **kw) -> Tuple[Tensor, FusionOutputs]:
    """Trainer.py:60-71 step loss for the single-task contract, with the trainer-owned weight
    folded to 1: CE(arousal_logits, labels) + sum(contrastive terms)."""
    out = fusion_forward(cfg, p, inputs, labels, **kw)
    ce = F.cross_entropy(out.arousal, labels)                        # Trainer.py:17,68
    loss = ce
    for c in out.contrastive:
        loss = loss + c.sum()
    return loss, out


# --------------------------------------------------------------------------
# ME-MHACL pieces (rows J, K, M of SURVEY.md section 8a)
# --------------------------------------------------------------------------
def memhacl_fusion(feats: Sequence[Tensor], p: Params, num_heads: int = 8, variant: str = "mean",
                   training: bool = True, buffers=None) -> Tensor:
    """MultiModalEncoder.forward fusion tail.
    variant "mean": ME-MHACL/model.py:68-74 (stack dim 0 -> self-MHA -> mean over modalities).
    variant "max":  MultimodalModel.py:388-406 (L2-normalise inputs, self-MHA, max over
                    modalities, fusion_mlp = Linear-ReLU-BN)."""
    if variant == "max":
        feats = [F.normalize(f, dim=-1) for f in feats]
    x = torch.stack(list(feats), dim=0)
    a = mha_self_seq_first(x, p["multihead_attn.in_proj_weight"], p["multihead_attn.in_proj_bias"],
                           p["multihead_attn.out_proj.weight"], p["multihead_attn.out_proj.bias"], num_heads)
    if variant == "mean":
        return a.mean(dim=0)
    fused = a.max(dim=0)[0]
    return mlp_chain(fused, p, "fusion_mlp.", [("linear", 0), ("relu", 1), ("bn", 2)], training, None, buffers)


def projection_head(x: Tensor, p: Params, training: bool = True, dropout_masks=None, buffers=None) -> Tensor:
    """ProjectionHead.forward (ME-MHACL/model.py:82-97; MultimodalModel.py:414-429)."""
    return mlp_chain(x, p, "net.", PROJECTION_PLAN, training, dropout_masks, buffers)


def classifier(x: Tensor, p: Params, training: bool = True, dropout_mask: Optional[Tensor] = None):
    """Classifier.forward (ME-MHACL/model.py:105-119; MultimodalModel.py:437-451)."""
    h = F.relu(F.linear(x, p["shared.0.weight"], p["shared.0.bias"]))
    if dropout_mask is not None and training:
        h = h * dropout_mask
    return (F.linear(h, p["fc_arousal.weight"], p["fc_arousal.bias"]),
            F.linear(h, p["fc_valence.weight"], p["fc_valence.bias"]))


# --------------------------------------------------------------------------
# Parameter construction with the reference initialisers (for sizes the reference cannot build)
# --------------------------------------------------------------------------
def init_params(cfg: FusionConfig, seed: int = 0, dtype=torch.float32) -> Tuple[Params, Dict[str, Tensor]]:
    """Parameters + BN buffers with the reference's key names and initialisers
    (nn.MultiheadAttention xavier-uniform in_proj / zero biases; nn.Linear default
    kaiming-uniform(a=sqrt 5); LN/BN weight 1 bias 0; temperature 0.01; contrastive_weight 1)."""
    import torch.nn as nn
    g = torch.random.get_rng_state()
    torch.manual_seed(seed)
    E, H, C = cfg.embed_dim, cfg.num_heads, cfg.num_classes
    mods: Dict[str, nn.Module] = {}
    if cfg.wiring == "bidirectional":
        mods["eeg_net.proj"] = nn.Linear(cfg.text_dim, E)
        mods["eye_net.proj"] = nn.Linear(cfg.image_dim, E)
        raw_w = 2 * E
    else:
        raw_w = 3 * E
    for name in ("cross_attn_e2p", "cross_attn_p2e"):
        mods[name + ".multihead_attn"] = nn.MultiheadAttention(E, H, batch_first=True)
        mods[name + ".gate.0"] = nn.Linear(2 * E, E)
        mods[name + ".norm"] = nn.LayerNorm(E)
    mods["attention_weights.0"] = nn.Linear(raw_w, 64)
    mods["attention_weights.2"] = nn.Linear(64, 3)
    mods["fusion.0"] = nn.Linear(3 * E, 256); mods["fusion.1"] = nn.BatchNorm1d(256)
    mods["fusion.4"] = nn.Linear(256, 128); mods["fusion.5"] = nn.BatchNorm1d(128)
    mods["arousal_head.0"] = nn.Linear(128, 128); mods["arousal_head.1"] = nn.BatchNorm1d(128)
    mods["arousal_head.4"] = nn.Linear(128, C)
    if cfg.valence:
        dims = [(128, 256), (256, 256), (256, 128), (128, 64)]
        for i, (a, b) in enumerate(dims):
            mods[f"valence_head.{4 * i}"] = nn.Linear(a, b)
            mods[f"valence_head.{4 * i + 1}"] = nn.BatchNorm1d(b)
        mods["valence_head.16"] = nn.Linear(64, C)
    params: Params = {}
    buffers: Dict[str, Tensor] = {}
    for mname, m in mods.items():
        for k, v in m.state_dict().items():
            full = f"{mname}.{k}"
            if "running_" in k or "num_batches" in k:
                buffers[full] = v.clone()
            else:
                params[full] = v.detach().clone().to(dtype)
    params["contrastive_weight"] = torch.ones(1, dtype=dtype)
    params["temperature"] = torch.tensor(0.01, dtype=dtype)
    torch.random.set_rng_state(g)
    return params, buffers


def synth_inputs(cfg: FusionConfig, batch: int, L: int = 64, R: int = 49, seed: int = 1234,
                 dtype=torch.float32):
    """Seeded synthetic inputs of SURVEY.md section 8(d): N(0,1) features, labels randint(0,3)."""
    g = torch.Generator().manual_seed(seed)
    if cfg.wiring == "bidirectional":
        xs = (torch.randn(batch, L, cfg.text_dim, generator=g, dtype=dtype),
              torch.randn(batch, R, cfg.image_dim, generator=g, dtype=dtype))
    else:
        xs = tuple(torch.randn(batch, cfg.embed_dim, generator=g, dtype=dtype) for _ in range(3))
    labels = torch.randint(0, cfg.num_classes, (batch,), generator=g)
    return xs, labels


# --------------------------------------------------------------------------
# Subnetwork encoder tail (MultimodalModel.py:83-105)
# --------------------------------------------------------------------------
def positional_table(d_model: int, max_len: int = 100, dtype=torch.float32) -> Tensor:
    """PositionalEncoding.__init__ (MultimodalModel.py:9-17): [max_len, d_model]."""
    half = torch.arange(0, d_model, 2, dtype=torch.float32)
    angle = torch.arange(max_len, dtype=torch.float32)[:, None] * torch.exp(half * (-math.log(10000.0) / d_model))[None, :]
    pe = torch.stack([torch.sin(angle), torch.cos(angle)], dim=2).reshape(max_len, -1)[:, :d_model].contiguous()
    return pe.to(dtype)


def encoder_layer(x: Tensor, p: Params, prefix: str, num_heads: int,
                  masks: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """One nn.TransformerEncoderLayer as Subnetwork builds it (MultimodalModel.py:89-95: post-norm, ReLU, batch_first):
        x = norm1(x + dropout1(self_attn(x)));  x = norm2(x + dropout2(linear2(dropout(relu(linear1(x))))))
    (third-party torch nn/modules/transformer.py TransformerEncoderLayer._sa_block / _ff_block, norm_first=False branch).
    masks=None: dropout off.  Training mode: `masks` holds PRE-SCALED keep masks for the four dropouts of the layer --
    "attn" [B,H,L,L] on the attention probabilities (MultiheadAttention(dropout=0.3)), "dropout1" [B,L,E], "dropout"
    [B,L,3E], "dropout2" [B,L,E]; tests/test_cpu_oracle_and_abi.py pins these four positions against torch's own layer."""
    B, L, E = x.shape
    m = masks or {}
    pm = m["attn"].reshape(B * num_heads, L, L) if "attn" in m else None
    sa = mha_self_seq_first(x.transpose(0, 1), p[prefix + "self_attn.in_proj_weight"], p[prefix + "self_attn.in_proj_bias"],
                            p[prefix + "self_attn.out_proj.weight"], p[prefix + "self_attn.out_proj.bias"],
                            num_heads, pm).transpose(0, 1)
    if "dropout1" in m:
        sa = sa * m["dropout1"]
    x = F.layer_norm(x + sa, (E,), p[prefix + "norm1.weight"], p[prefix + "norm1.bias"], 1e-5)
    h = torch.relu(F.linear(x, p[prefix + "linear1.weight"], p[prefix + "linear1.bias"]))
    if "dropout" in m:
        h = h * m["dropout"]
    ff = F.linear(h, p[prefix + "linear2.weight"], p[prefix + "linear2.bias"])
    if "dropout2" in m:
        ff = ff * m["dropout2"]
    return F.layer_norm(x + ff, (E,), p[prefix + "norm2.weight"], p[prefix + "norm2.bias"], 1e-5)


def subnetwork(x: Tensor, p: Params, num_heads: int = 4, num_layers: int = 2,
               masks: Optional[Sequence[Dict[str, Tensor]]] = None) -> Tensor:
    """Subnetwork.forward (MultimodalModel.py:99-105); x:[B,D] (one token, as the reference feeds it) or [B,L,D]
    (generalised to L tokens).  masks: per-layer dropout masks (see encoder_layer) for training mode, None = dropout off."""
    squeeze = x.ndim == 2
    if squeeze:
        x = x.unsqueeze(1)                                           # :102
    h = F.linear(x, p["proj.weight"], p["proj.bias"])               # :102
    pe = p["pos_encoder.pe"] if "pos_encoder.pe" in p else positional_table(h.shape[-1], 100, h.dtype).unsqueeze(0)
    h = h + pe[:, :h.shape[1]].to(h.dtype)                           # :103 (PositionalEncoding.forward :19-20)
    for i in range(num_layers):
        h = encoder_layer(h, p, f"transformer.layers.{i}.", num_heads, None if masks is None else masks[i])     # :104
    h = F.layer_norm(h, (h.shape[-1],), p["norm.weight"], p["norm.bias"], 1e-5)   # :105
    return h.squeeze(1) if squeeze else h
