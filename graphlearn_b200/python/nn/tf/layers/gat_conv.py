"""graphlearn/python/nn/tf/layers/gat_conv.py"""
from .....nn import GATConv  # noqa: F401
