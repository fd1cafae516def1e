"""deepsvg_b200: B200-native (sm_100a) implementation of DeepSVG's hot path -- the hierarchical SVG transformer (V)AE
train step (`SVGTransformer.forward` -> `SVGLoss` -> backward) behind the reference's own module API.

    from deepsvg_b200 import SVGTransformer, SVGLoss, Hierarchical
    model = SVGTransformer(model_cfg).cuda();  loss_fn = SVGLoss(model_cfg)

See DESIGN.md / INTEGRATION.md.  There is no CPU fallback: the CUDA library must be present.
"""
from .config import (Hierarchical, HierarchicalSelfMatching, OneStageOneShot, SketchRNN, Sketchformer,  # noqa: F401
                     _DefaultConfig)
from .data import PackedBatch, pack_icons, pack_tensors  # noqa: F401
from .loss import SVGLoss  # noqa: F401
from .model import SVGTransformer  # noqa: F401
from .optim import FusedAdamW  # noqa: F401

__all__ = ["SVGTransformer", "SVGLoss", "Hierarchical", "OneStageOneShot", "HierarchicalSelfMatching", "SketchRNN",
           "Sketchformer", "_DefaultConfig", "FusedAdamW", "PackedBatch", "pack_icons", "pack_tensors"]
