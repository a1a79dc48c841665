"""graphlearn/python/nn/tf/layers/ego_gin_conv.py"""
from .....nn import EgoGINConv  # noqa: F401
