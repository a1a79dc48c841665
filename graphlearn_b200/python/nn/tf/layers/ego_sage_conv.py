"""graphlearn/python/nn/tf/layers/ego_sage_conv.py"""
from .....nn import EgoSAGEConv  # noqa: F401
