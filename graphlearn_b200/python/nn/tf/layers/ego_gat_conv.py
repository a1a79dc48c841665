"""graphlearn/python/nn/tf/layers/ego_gat_conv.py"""
from .....nn import EgoGATConv  # noqa: F401
