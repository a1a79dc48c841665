"""Edge traversal sampler (graphlearn/python/sampler/edge_sampler.py)."""
from .node_sampler import ByOrderEdgeSampler, EdgeSampler, RandomEdgeSampler, ShuffleEdgeSampler  # noqa: F401
