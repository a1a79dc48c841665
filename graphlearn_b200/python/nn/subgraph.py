"""graphlearn/python/nn/subgraph.py"""
from ...data.values import SubGraph  # noqa: F401
