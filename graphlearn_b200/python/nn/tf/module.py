"""graphlearn/python/nn/tf/module.py"""
from torch.nn import Module  # noqa: F401
