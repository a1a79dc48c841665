"""graphlearn/python/nn/tf/layers/ego_layer.py"""
from .....nn import EgoLayer, EgoConv  # noqa: F401
