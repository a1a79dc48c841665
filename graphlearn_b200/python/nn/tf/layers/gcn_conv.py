"""graphlearn/python/nn/tf/layers/gcn_conv.py"""
from .....nn import GCNConv  # noqa: F401
