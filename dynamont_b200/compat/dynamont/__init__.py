"""Drop-in for the reference's ``dynamont`` package root (src/dynamont/__init__.py:8): put ``dynamont_b200/compat`` on
``sys.path`` ahead of (or instead of) the reference's compiled extension and ``from dynamont import Aligner, PoreType``
resolves to the B200 implementation — the reference's own front ends (segmentation/segment.py:34-45,161,
segmentation/utils.py:154-191, train.py) run unchanged."""
from dynamont._dynamont import Aligner, PoreType, pore_type  # noqa: F401

__version__ = "b200"
__all__ = ["Aligner", "PoreType", "__version__"]
