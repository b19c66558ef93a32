"""dynamont_b200 — B200-native implementation of Dynamont's segmentation hot path (basic-mode banded
forward-backward HMM, posterior decoding and training statistics) behind the reference's operator API."""
from .aligner import Aligner, PoreType, pore_type  # noqa: F401
from .parallel import MultiDeviceAligner  # noqa: F401
from . import frontend  # noqa: F401  (batched dynamont-resquiggle front end: jobs -> GPU -> CSV writer)

__all__ = ["Aligner", "PoreType", "pore_type", "MultiDeviceAligner"]
