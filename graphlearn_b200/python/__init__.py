"""Import-path compatibility with the reference's package layout: scripts written as

    import graphlearn as gl
    import graphlearn.python.nn.tf as tfg          # or graphlearn.python.nn.pytorch as thg

keep working after replacing ``graphlearn`` by ``graphlearn_b200`` - ``graphlearn_b200.python.nn.tf`` and ``.pytorch`` re-export
the same flat name sets as graphlearn/python/nn/{tf,pytorch}/__init__.py (PyTorch modules under both)."""
from .. import *  # noqa: F401,F403
from .. import nn  # noqa: F401
