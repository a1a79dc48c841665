"""Edge traversal sampler (graphlearn/python/sampler/edge_sampler.py)."""
from .node_sampler import EdgeSampler  # noqa: F401
