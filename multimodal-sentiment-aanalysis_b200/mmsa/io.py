"""Batch-format and checkpoint adapters either side of the hot path.

* The reference's loaders yield `({'eeg': x0, 'eye': x1, 'pps': x2}, labels)` dict batches (data/Dataset.py:65-67,
  consumed by Trainer.py:51-56) or 5-tuples `(eeg, eye, pps, arousal, valence)` (dataLoader/DataLoader.py:152-156,
  consumed by MultiTaskTrainer.py:186-195).  `FeatureBatches` yields either format from in-memory feature tensors
  (BERT tokens [N,L,768], ResNet regions [N,49,2048]) staged in pinned host memory.
* Checkpoints are bare `state_dict()`s saved with torch.save (Trainer.py:111,262), possibly from nn.DataParallel
  (keys prefixed `module.`, stripped by Tester.py:32-33).  `load_reference_state_dict` applies one to a drop-in module."""
from __future__ import annotations

from typing import Dict, Iterator, Optional, Tuple, Union

import torch

Tensor = torch.Tensor


def strip_module_prefix(state_dict: Dict[str, Tensor]) -> Dict[str, Tensor]:
    """Tester.py:32-33: `{k.replace('module.', ''): v}` for checkpoints written from nn.DataParallel."""
    return {(k[len("module."):] if k.startswith("module.") else k): v for k, v in state_dict.items()}


def load_reference_state_dict(model: torch.nn.Module, source: Union[str, Dict[str, Tensor]], strict: bool = True,
                              ignore_prefixes: Tuple[str, ...] = ()):
    """Load a reference checkpoint (path or dict) into a drop-in module.  `ignore_prefixes` drops the out-of-scope
    encoder weights (e.g. ("eeg_net.", "eye_net.", "pps_net.")) when the drop-in is fed precomputed features."""
    sd = torch.load(source, map_location="cpu") if isinstance(source, str) else source
    sd = strip_module_prefix(sd)
    if ignore_prefixes:
        sd = {k: v for k, v in sd.items() if not k.startswith(tuple(ignore_prefixes))}
    return model.load_state_dict(sd, strict=strict)


class FeatureBatches:
    """Iterate mini-batches of precomputed features in one of the reference's two batch formats.

    format="dict":  ({'eeg': text, 'eye': image, 'pps': third}, labels)          (Trainer.py:51-56)
    format="tuple": (text, image, third, arousal_labels, valence_labels)         (MultiTaskTrainer.py:186-195)
    Every batch is gathered (torch.index_select(..., out=...)) into one of `n_staging` preallocated PINNED staging
    buffers, rotated per batch, so the trainers' `.to(device)` copies (Trainer.py:53-56) read page-locked memory -- an
    index with a tensor alone would return a fresh pageable copy.  A yielded batch stays valid until `n_staging - 1`
    further batches have been drawn (the trainers consume each batch before asking for the next)."""

    def __init__(self, text: Tensor, image: Tensor, labels: Tensor, batch_size: int = 64, third: Optional[Tensor] = None,
                 valence_labels: Optional[Tensor] = None, fmt: str = "dict", shuffle: bool = False, seed: int = 0,
                 drop_last: bool = False, n_staging: int = 3):
        assert fmt in ("dict", "tuple") and text.shape[0] == image.shape[0] == labels.shape[0]
        self.pinned = torch.cuda.is_available()
        self.text, self.image = text.contiguous(), image.contiguous()
        self.third = (third if third is not None else torch.zeros(text.shape[0], 1)).contiguous()
        self.labels = labels.long().contiguous()
        self.valence = (valence_labels if valence_labels is not None else labels).long().contiguous()
        self.batch_size, self.fmt, self.shuffle, self.drop_last = batch_size, fmt, shuffle, drop_last
        self._gen = torch.Generator().manual_seed(seed)
        self._staging = []
        for _ in range(max(1, n_staging)):
            bufs = []
            for src in (self.text, self.image, self.third, self.labels, self.valence):
                b = torch.empty((batch_size,) + tuple(src.shape[1:]), dtype=src.dtype)
                bufs.append(b.pin_memory() if self.pinned else b)
            self._staging.append(bufs)
        self._turn = 0

    @property
    def dataset(self):
        """`len(loader.dataset)` = number of samples, as MultiTaskTrainer.py:225,281,337,398,459,504 reads it off a
        torch DataLoader."""
        return range(self.text.shape[0])

    def __len__(self) -> int:
        n = self.text.shape[0]
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def __iter__(self) -> Iterator:
        n = self.text.shape[0]
        order = torch.randperm(n, generator=self._gen) if self.shuffle else torch.arange(n)
        for i in range(len(self)):
            idx = order[i * self.batch_size:(i + 1) * self.batch_size]
            bufs = self._staging[self._turn]
            self._turn = (self._turn + 1) % len(self._staging)
            k = idx.numel()
            t, im, th, la, va = (torch.index_select(src, 0, idx, out=buf[:k])
                                 for src, buf in zip((self.text, self.image, self.third, self.labels, self.valence), bufs))
            if self.fmt == "dict":
                yield {"eeg": t, "eye": im, "pps": th}, la
            else:
                yield t, im, th, la, va
