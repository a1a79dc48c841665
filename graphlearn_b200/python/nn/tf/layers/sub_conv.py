"""graphlearn/python/nn/tf/layers/sub_conv.py"""
from .....nn import SubConv  # noqa: F401
