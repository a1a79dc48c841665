"""graphlearn/python/nn/tf/layers/ego_rgcn_conv.py"""
from .....nn import EgoRGCNConv  # noqa: F401
