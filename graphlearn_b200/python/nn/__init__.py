"""graphlearn/python/nn/__init__.py: Data, Dataset, SubGraph, HeteroSubGraph (+ the tf / pytorch sub-packages)."""
from ...nn import Data, Dataset, HeteroSubGraph  # noqa: F401
from ...data.values import SubGraph  # noqa: F401
from . import pytorch, tf  # noqa: F401
