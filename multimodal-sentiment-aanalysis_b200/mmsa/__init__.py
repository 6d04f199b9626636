"""mmsa -- B200 (sm_100a) hot path of the multimodal fusion / ME-MHACL / contrastive model.

Python host mirroring the reference's nn.Module interface (MultimodalModel.py, ME-MHACL/model.py)
over a C-ABI CUDA library (include/mmsa.h).  Importing the package does not need a GPU; running
any op does, and raises if the extension is missing (no fallback)."""
from . import _lib
from .model import (Classifier, CrossModalTransformer, FeatureProjection, MultiModalEncoder,
                    MultimodalTransformerModel, PositionalEncoding, ProjectionHead, Subnetwork)
from .ops import cross_entropy, infonce, ntxent, supcon
from .optim import FusedClipAdamW
from .io import FeatureBatches, load_reference_state_dict, strip_module_prefix

__all__ = ["MultimodalTransformerModel", "CrossModalTransformer", "FeatureProjection", "MultiModalEncoder",
           "ProjectionHead", "Classifier", "Subnetwork", "PositionalEncoding", "cross_entropy", "infonce", "supcon", "ntxent", "FusedClipAdamW",
           "FeatureBatches", "load_reference_state_dict", "strip_module_prefix", "_lib"]
