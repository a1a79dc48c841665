"""graphlearn/python/nn/hetero_subgraph.py"""
from ...nn.hetero import HeteroSubGraph  # noqa: F401
