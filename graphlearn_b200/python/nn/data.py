"""graphlearn/python/nn/data.py"""
from ...nn.data import Data  # noqa: F401
