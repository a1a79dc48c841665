"""graphlearn/python/nn/dataset.py"""
from ...nn.dataset import Dataset  # noqa: F401
