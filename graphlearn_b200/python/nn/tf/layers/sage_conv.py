"""graphlearn/python/nn/tf/layers/sage_conv.py"""
from .....nn import SAGEConv  # noqa: F401
