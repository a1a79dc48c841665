"""graphlearn/python/nn/tf/layers/linear_layer.py"""
from .....nn import LinearLayer  # noqa: F401
