"""graphlearn/python/nn/tf/layers/hetero_conv.py"""
from .....nn import HeteroConv  # noqa: F401
