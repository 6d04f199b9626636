"""Drop-in nn.Modules for the reference's fusion / ME-MHACL path.

Same class names, constructor defaults, attribute names, parameter names/shapes (state_dict keys)
and forward signatures as /root/reference/MML_ZYC/MultimodalModel.py and ME-MHACL/model.py, so
Trainer.py:60, Tester.py:53 and dataLoader/MultiTaskTrainer.py:199 can call them unchanged.  The
sub-modules (nn.MultiheadAttention, nn.Linear, nn.LayerNorm, nn.BatchNorm1d ...) are kept only as
PARAMETER CONTAINERS (so initialisation and checkpoints match the reference); their own forward is
never called -- every forward here runs the hand-written sm_100a kernels through mmsa.ops, and
fails loudly off-GPU (no fallback)."""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import kernels as K
from . import ops

Tensor = torch.Tensor


_DropoutState = ops._DropoutState


def run_sequential(x: Tensor, seq: nn.Sequential, drop: _DropoutState, name: str, cd: torch.dtype,
                   x2: Optional[Tensor] = None) -> Tensor:
    """Execute an nn.Sequential of the reference's head/fusion layouts on the CUDA kernels
    (ops.sequential: one autograd node per Sequential; fp32 in, fp32 out, `cd` GEMM operands)."""
    return ops.sequential(x, seq, drop, name, cd, x2=x2)


class FeatureProjection(nn.Module):
    """`Subnetwork.proj = nn.Linear(input_dim, feat_dim)` (MultimodalModel.py:86,102): the projection
    GEMM that maps BERT tokens [B,L,768] / ResNet-50 regions [B,49,2048] to the fusion width."""

    def __init__(self, input_dim: int, feat_dim: int = 256):
        super().__init__()
        self.proj = nn.Linear(input_dim, feat_dim)

    def forward(self, x: Tensor) -> Tensor:
        return ops.linear(x, self.proj.weight, self.proj.bias)


def _prepare(module: nn.Module, cd: torch.dtype) -> None:
    """Refresh the compute-dtype operand copies of every weight matrix of `module` in one launch."""
    ops.prepare_weights([p for p in module.parameters() if p.ndim >= 2], cd)


class PositionalEncoding(nn.Module):
    """MultimodalModel.py:8-20: the sinusoidal table as a buffer `pe` [1, max_len, d_model] (parameter container; the add
    runs in mmsa_add_rows)."""

    def __init__(self, d_model: int, max_len: int = 5000):
        super().__init__()
        import math
        # angle[pos, i] = pos / 10000^(2i/d); even columns take the sine, odd columns the cosine (:11-15)
        half = torch.arange(0, d_model, 2, dtype=torch.float32)
        angle = torch.arange(max_len, dtype=torch.float32)[:, None] * torch.exp(half * (-math.log(10000.0) / d_model))[None, :]
        pe = torch.stack([torch.sin(angle), torch.cos(angle)], dim=2).reshape(max_len, -1)[:, :d_model].contiguous()
        self.register_buffer("pe", pe.unsqueeze(0))


class Subnetwork(nn.Module):
    """Drop-in for MultimodalModel.Subnetwork (MultimodalModel.py:83-105), the encoder tail in front of the fusion path
    (SURVEY.md section 8(f) rank 2): proj -> + positional table -> num_layers x post-norm TransformerEncoderLayer(d, nhead,
    ff = 3d, ReLU, dropout 0.3) -> LayerNorm.  Same attribute names and state_dict keys.  Train mode applies all four
    dropouts of a layer (attention probabilities, dropout1, dropout, dropout2) with in-kernel Philox masks.  The reference feeds [B, input_dim]
    (one token); [B, L, input_dim] runs the same layers over L tokens (text tokens get a self-attention stage before the
    cross-attention).  torch modules are parameter containers only; every op runs the sm_100a kernels."""

    def __init__(self, input_dim: int, feat_dim: int = 256, num_layers: int = 2, nhead: int = 4,
                 compute_dtype: torch.dtype = torch.float32):
        super().__init__()
        self.compute_dtype = compute_dtype
        self.proj = nn.Linear(input_dim, feat_dim)
        self.pos_encoder = PositionalEncoding(feat_dim, max_len=100)
        layer = nn.TransformerEncoderLayer(d_model=feat_dim, nhead=nhead, dim_feedforward=feat_dim * 3, dropout=0.3,
                                           batch_first=True)
        self.transformer = nn.TransformerEncoder(layer, num_layers, enable_nested_tensor=False)
        self.norm = nn.LayerNorm(feat_dim)
        self._drop = _DropoutState()

    def set_dropout(self, p: float) -> "Subnetwork":
        """every dropout of the encoder tail: the layers' three nn.Dropout modules and the attention-probability rate"""
        for m in self.modules():
            if isinstance(m, nn.Dropout):
                m.p = p
            if isinstance(m, nn.MultiheadAttention):
                m.dropout = p
        return self

    def forward(self, x: Tensor) -> Tensor:
        cd = self.compute_dtype
        squeeze = x.ndim == 2
        if squeeze:
            x = x.unsqueeze(1)
        B, L, D = x.shape
        if L > self.pos_encoder.pe.shape[1]:
            raise ValueError(f"mmsa.Subnetwork: sequence length {L} exceeds the positional table ({self.pos_encoder.pe.shape[1]})")
        if not x.is_floating_point():
            x = x.float()
        h = ops.linear(ops.cast(x.contiguous(), cd).view(B * L, D), self.proj.weight, self.proj.bias)
        h = ops.AddRowsFn.apply(h, self.pos_encoder.pe[0].contiguous(), L).view(B, L, -1)
        for i, layer in enumerate(self.transformer.layers):
            h = ops.encoder_layer(h, layer, self._drop, f"transformer.layers.{i}", cd)
        h = ops.add_layer_norm(h, None, self.norm)
        h = ops.cast(h, torch.float32)
        self._drop.commit(h.device)
        return h.squeeze(1) if squeeze else h


class CrossModalTransformer(nn.Module):
    """MultimodalModel.py:108-149.  forward(query, key, value) with 2-D [B,E] or 3-D [B,L,E] inputs (key and value may be
    one tensor, as at every reference call site, or two); the gate concat runs on the feature axis, so Lq > 1 works
    (identical to the reference for Lq == 1)."""

    def __init__(self, embed_dim: int = 256, num_heads: int = 4):
        super().__init__()
        self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.multihead_attn = nn.MultiheadAttention(embed_dim=embed_dim, num_heads=num_heads, batch_first=True)
        self.gate = nn.Sequential(nn.Linear(embed_dim * 2, embed_dim), nn.Sigmoid())
        self.norm = nn.LayerNorm(embed_dim)

    def kernel_params(self) -> Tuple[Tensor, ...]:
        a = self.multihead_attn
        return (a.in_proj_weight, a.in_proj_bias, a.out_proj.weight, a.out_proj.bias,
                self.gate[0].weight, self.gate[0].bias, self.norm.weight, self.norm.bias)

    def forward(self, query: Tensor, key: Tensor, value: Tensor) -> Tensor:
        squeeze = query.ndim == 2
        if squeeze:
            query = query.unsqueeze(1)
        if key.ndim == 2:
            key = key.unsqueeze(1)
        if value.ndim == 2:
            value = value.unsqueeze(1)
        # every reference call site passes ONE tensor as key and value (MultimodalModel.py:287-297): K and V then come out
        # of one packed N = 2E GEMM; a distinct value tensor takes two N = E GEMMs into the same packed buffer
        same = key.data_ptr() == value.data_ptr() and key.shape == value.shape and key.stride() == value.stride()
        out = ops.cross_block(query, key, *self.kernel_params(), self.num_heads, value=None if same else value)
        return out.squeeze(1) if squeeze else out


class MultimodalTransformerModel(nn.Module):
    """Drop-in for MultimodalModel.MultimodalTransformerModel (MultimodalModel.py:152-322).

    wiring="native":  the reference's three-modality wiring at its own sizes.  The modality encoders
        (eeg_net/eye_net/pps_net, :164-166) are outside the hot path: pass them in `encoders`, or
        leave the default nn.Identity() and feed [B,E] features.
    wiring="bidirectional": BASELINE.json's text+image re-skin -- eeg_net/eye_net become the
        projection GEMMs over BERT tokens [B,L,text_dim] and ResNet regions [B,R,image_dim]; the
        third positional input (pps) is accepted and ignored.
    contract="multitask": forward(..., labels=(arousal, valence)) -> 5-tuple (MultiTaskTrainer.py:199).
    contract="single":    forward(x0, x1, x2, labels[B]) -> (logits, contrastive_loss) and
                          forward(x0, x1, x2) -> logits   (Trainer.py:60, Tester.py:53)."""

    def __init__(self, num_classes: int = 3, temperature: float = 0.01, *, embed_dim: int = 256, num_heads: int = 4,
                 wiring: str = "native", text_dim: int = 768, image_dim: int = 2048, contract: str = "multitask",
                 compute_dtype: torch.dtype = torch.float32, encoders: Optional[Sequence[nn.Module]] = None,
                 valence: bool = True):
        super().__init__()
        assert wiring in ("native", "bidirectional") and contract in ("multitask", "single")
        E = embed_dim
        self.embed_dim, self.num_heads, self.wiring, self.contract = E, num_heads, wiring, contract
        self.compute_dtype = compute_dtype
        if wiring == "native":
            enc = list(encoders) if encoders is not None else [nn.Identity(), nn.Identity(), nn.Identity()]
            self.eeg_net, self.eye_net, self.pps_net = enc
            raw_width = 3 * E
        else:
            self.eeg_net = FeatureProjection(text_dim, E)
            self.eye_net = FeatureProjection(image_dim, E)
            self.pps_net = nn.Identity()
            raw_width = 2 * E
        self.cross_attn_e2p = CrossModalTransformer(E, num_heads)
        self.cross_attn_p2e = CrossModalTransformer(E, num_heads)
        self.attention_weights = nn.Sequential(nn.Linear(raw_width, 64), nn.GELU(), nn.Linear(64, 3), nn.Softmax(dim=1))
        self.fusion = nn.Sequential(
            nn.Linear(E * 3, 256), nn.BatchNorm1d(256, eps=1e-5), nn.GELU(), nn.Dropout(0.3),
            nn.Linear(256, 128), nn.BatchNorm1d(128, eps=1e-5), nn.GELU(), nn.Dropout(0.3))
        self.arousal_head = nn.Sequential(
            nn.Linear(128, 128), nn.BatchNorm1d(128), nn.GELU(), nn.Dropout(0.3), nn.Linear(128, num_classes))
        if valence:
            self.valence_head = nn.Sequential(
                nn.Linear(128, 256), nn.BatchNorm1d(256), nn.GELU(), nn.Dropout(0.3),
                nn.Linear(256, 256), nn.BatchNorm1d(256), nn.GELU(), nn.Dropout(0.3),
                nn.Linear(256, 128), nn.BatchNorm1d(128), nn.GELU(), nn.Dropout(0.3),
                nn.Linear(128, 64), nn.BatchNorm1d(64), nn.GELU(), nn.Dropout(0.3),
                nn.Linear(64, num_classes))
        else:
            self.valence_head = None
        self.contrastive_weight = nn.Parameter(torch.ones(1))
        self.temperature = nn.Parameter(torch.tensor(temperature))
        self._drop = _DropoutState()
        # data-parallel contrastive sharding (set by mmsa.dist.shard_contrastive)
        self.dp_group = None
        self.overlap_contrastive = True      # contrastive branch on a second stream (text+image wiring)
        self._side_stream = None

    # -- helpers ------------------------------------------------------------------------------------
    def set_dropout(self, p: float) -> "MultimodalTransformerModel":
        for m in self.modules():
            if isinstance(m, nn.Dropout):
                m.p = p
        return self

    def _cd(self, x: Tensor) -> Tensor:
        if not x.is_floating_point():
            x = x.float()
        return ops.cast(x.contiguous(), self.compute_dtype)

    def compute_contrastive_loss(self, feat1: Tensor, feat2: Tensor, labels: Tensor) -> Tensor:
        """MultimodalModel.py:232-260."""
        return ops.infonce(feat1, feat2, labels, self.temperature)

    def prepare_step(self) -> None:
        """Call once per training step (after the optimiser update): one launch re-casts every fp32
        master weight matrix to its bf16 operand copy.  Without it the copies are refreshed lazily,
        one cast per weight, whenever a parameter's version changed."""
        _prepare(self, self.compute_dtype)

    def _tail(self, raw_a: Tensor, raw_b: Optional[Tensor], slots: Sequence[Tensor], a_lp: Optional[Tensor] = None,
              b_lp: Optional[Tensor] = None):
        """modality weights (:171-176, :299-301) -> weighted concat (:302-306) -> fusion (:309) -> heads (:312-313).
        Everything here is [B,*]: fp32 at the autograd boundaries, `compute_dtype` GEMM operands inside.  The chain is
        latency-bound (a few dozen launches of a few microseconds each), so operand copies in the compute dtype are passed
        from producer to consumer (a_lp / b_lp: the pooled features as the LayerNorm kernel wrote them) instead of being
        re-cast, and the modality-weight MLP tail + softmax + concat run as one kernel."""
        cd = self.compute_dtype
        if ops.is_modal_head(self.attention_weights) and len(slots) <= 4:
            fused, fused_lp, w = ops.modal_head(raw_a, raw_b, self.attention_weights, slots, cd, a_lp=a_lp, b_lp=b_lp)
        else:
            logits3 = run_sequential(raw_a, self.attention_weights, self._drop, "attention_weights", cd, x2=raw_b)
            fused, w = ops.modal_concat(logits3, slots)
            fused_lp = None
        fused, fused2_lp = ops.sequential(fused, self.fusion, self._drop, "fusion", cd, x_lp=fused_lp, want_lp=True)
        arousal = ops.sequential(fused, self.arousal_head, self._drop, "arousal_head", cd, x_lp=fused2_lp)
        valence = None
        if self.valence_head is not None and self.contract == "multitask":
            valence = ops.sequential(fused, self.valence_head, self._drop, "valence_head", cd, x_lp=fused2_lp)
        return arousal, valence

    # -- forward ------------------------------------------------------------------------------------
    def forward(self, eeg: Tensor, eye: Tensor, pps: Optional[Tensor] = None, labels=None):
        out = self._forward(eeg, eye, pps, labels)
        self._drop.commit(self.temperature.device)       # one kernel node: the Philox position moves on every (re)play
        return out

    def _forward(self, eeg: Tensor, eye: Tensor, pps: Optional[Tensor] = None, labels=None):
        if labels is not None and self.contract == "multitask":
            con_labels = labels[0]                                   # MultimodalModel.py:273
        else:
            con_labels = labels
        if con_labels is not None:
            con_labels = con_labels.contiguous().long()
        contrastive: List[Tensor] = []
        side = None
        if self.wiring == "native":
            f0 = self._cd(self.eeg_net(eeg))
            f1 = self._cd(self.eye_net(eye))
            f2 = self._cd(self.pps_net(pps))
            if con_labels is not None:                               # :271-284
                for f in (f0, f1, f2):                               # weighted (:315-317) inside the loss kernels
                    contrastive.append(ops.infonce(f, f, con_labels, self.temperature, weight=self.contrastive_weight))
            e1 = self.cross_attn_e2p(f0, f1, f1)                     # :287
            e2 = self.cross_attn_p2e(f0, f2, f2)                     # :293
            f32 = torch.float32
            raw_a, raw_b = torch.cat([ops.cast(f, f32) for f in (f0, f1, f2)], dim=1), None   # :300 (tiny [B,3E] copy)
            slots = (ops.cast(f0, f32), ops.cast(e1, f32), ops.cast(e2, f32))
        else:
            text, image = self._cd(eeg), self._cd(eye)
            f0, fv, e1, e2, f0_lp, fv_lp = ops.fusion_core(
                text, image, self.eeg_net.proj.weight, self.eeg_net.proj.bias, self.eye_net.proj.weight,
                self.eye_net.proj.bias, self.num_heads, self.cross_attn_e2p.kernel_params(),
                self.cross_attn_p2e.kernel_params())
            raw_a, raw_b = f0, fv
            slots = (f0, e1, e2)
            if con_labels is not None:
                fast = self.compute_dtype == torch.bfloat16      # split-bf16 tensor-core GEMMs (fp32-accurate)
                # The contrastive branch (normalise, similarity GEMM, row reductions, and under data parallelism the
                # embedding all-gather) and the classifier tail are independent chains of small, latency-bound kernels
                # hanging off the same pooled features: the branch is enqueued on a second stream, so both chains (and
                # their backward halves, which autograd replays on the stream of the forward) overlap -- also as two
                # parallel branches of the captured CUDA graph.
                main = torch.cuda.current_stream(e1.device)
                if self.overlap_contrastive and ops.OVERLAP_TAIL:
                    if self._side_stream is None:
                        self._side_stream = torch.cuda.Stream(device=e1.device, priority=ops.critical_priority())   # feeds the loss
                    side = self._side_stream
                    side.wait_stream(main)
                with torch.cuda.stream(side if side is not None else main):
                    cw = self.contrastive_weight                     # :315-317, applied inside the loss kernels -> shape (1,)
                    if self.dp_group is not None:
                        from . import dist as mdist
                        contrastive.append(mdist.sharded_infonce(e1, e2, con_labels, self.temperature, self.dp_group, fast=fast,
                                                                 weight=cw))
                    else:
                        contrastive.append(ops.infonce(e1, e2, con_labels, self.temperature, fast=fast, weight=cw))
                    if self.contract == "multitask":
                        contrastive.append(ops.infonce(e2, e1, con_labels, self.temperature, fast=fast, weight=cw))
        arousal, valence = self._tail(raw_a, raw_b, slots, *((f0_lp, fv_lp) if self.wiring == "bidirectional" else ()))
        if side is not None:                                         # join the contrastive branch
            cur = torch.cuda.current_stream(arousal.device)
            cur.wait_stream(side)
            for c in contrastive:
                c.record_stream(cur)
        if self.contract == "single":
            if labels is None:
                return arousal                                       # Tester.py:53
            total = contrastive[0]
            for c in contrastive[1:]:
                total = total + c
            return arousal, total                                    # Trainer.py:60
        if labels is None:
            return arousal, valence                                  # :319-320
        while len(contrastive) < 3:
            contrastive.append(torch.zeros(1, device=arousal.device))
        return arousal, valence, contrastive[0], contrastive[1], contrastive[2]   # :321-322


class MultiModalEncoder(nn.Module):
    """ME-MHACL fusion tail (ME-MHACL/model.py:47-74; variant MultimodalModel.py:357-406): three
    modality features -> 3-token self-attention (8 heads) -> mean (or max + fusion_mlp) over modalities.
    The Conv1d/BiLSTM modality encoders are outside the hot path: pass them in `encoders` or feed features."""

    def __init__(self, feat_dim: int = 256, num_heads: int = 8, variant: str = "mean",
                 encoders: Optional[Sequence[nn.Module]] = None, compute_dtype: torch.dtype = torch.float32):
        super().__init__()
        assert variant in ("mean", "max")
        self.feat_dim, self.num_heads, self.variant = feat_dim, num_heads, variant
        self.compute_dtype = compute_dtype
        enc = list(encoders) if encoders is not None else [nn.Identity(), nn.Identity(), nn.Identity()]
        self.eeg_net, self.eye_net, self.pps_net = enc
        self.multihead_attn = nn.MultiheadAttention(embed_dim=feat_dim, num_heads=num_heads, batch_first=False)
        if variant == "max":
            self.fusion_mlp = nn.Sequential(nn.Linear(feat_dim, feat_dim), nn.ReLU(), nn.BatchNorm1d(feat_dim))
        self._drop = _DropoutState()

    def forward(self, eeg: Tensor, eye: Tensor, pps: Tensor, labels=None) -> Tensor:
        cd = self.compute_dtype
        feats = [ops.cast(n(x).float().contiguous(), cd) for n, x in
                 ((self.eeg_net, eeg), (self.eye_net, eye), (self.pps_net, pps))]
        if self.variant == "max":                                    # MultimodalModel.py:388-390
            feats = [ops.l2_normalize(f) for f in feats]
        x = ops.stack_tokens(feats)                                  # [B,3,E]
        a = self.multihead_attn
        y = ops.self_attention(x, a.in_proj_weight, a.in_proj_bias, a.out_proj.weight, a.out_proj.bias, self.num_heads)
        if self.variant == "mean":
            return ops.cast(ops.mean_pool(y), torch.float32)         # ME-MHACL/model.py:73
        fused = ops.max_pool(y)                                      # MultimodalModel.py:401
        out = run_sequential(fused, self.fusion_mlp, self._drop, "fusion_mlp", cd)
        self._drop.commit(out.device)
        return out


class ProjectionHead(nn.Module):
    """ME-MHACL/model.py:77-97 (Linear-ReLU-BN-Dropout x2 + Linear)."""

    def __init__(self, in_dim: int = 256, hidden_dim: int = 256, out_dim: int = 128,
                 compute_dtype: torch.dtype = torch.float32):
        super().__init__()
        self.compute_dtype = compute_dtype
        self.net = nn.Sequential(
            nn.Linear(in_dim, hidden_dim), nn.ReLU(inplace=True), nn.BatchNorm1d(hidden_dim), nn.Dropout(0.5),
            nn.Linear(hidden_dim, out_dim), nn.ReLU(inplace=True), nn.BatchNorm1d(out_dim), nn.Dropout(0.5),
            nn.Linear(out_dim, out_dim))
        self._drop = _DropoutState()

    def forward(self, x: Tensor) -> Tensor:
        out = run_sequential(x, self.net, self._drop, "net", self.compute_dtype)
        self._drop.commit(out.device)
        return out


class Classifier(nn.Module):
    """ME-MHACL/model.py:100-119 / MultimodalModel.py:432-451 (num_out = 2 or 3)."""

    def __init__(self, in_dim: int = 256, hidden_dim: int = 128, num_out: int = 3,
                 compute_dtype: torch.dtype = torch.float32):
        super().__init__()
        self.compute_dtype = compute_dtype
        self.shared = nn.Sequential(nn.Linear(in_dim, hidden_dim), nn.ReLU(inplace=True), nn.Dropout(0.5))
        self.fc_arousal = nn.Linear(hidden_dim, num_out)
        self.fc_valence = nn.Linear(hidden_dim, num_out)
        self._drop = _DropoutState()

    def forward(self, x: Tensor):
        cd = self.compute_dtype
        h = run_sequential(x, self.shared, self._drop, "shared", cd)
        out_a = ops.linear(h, self.fc_arousal.weight, self.fc_arousal.bias, out_fp32=True, cd=cd)
        out_v = ops.linear(h, self.fc_valence.weight, self.fc_valence.bias, out_fp32=True, cd=cd)
        self._drop.commit(out_a.device)
        return out_a, out_v
