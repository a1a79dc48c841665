"""Small helpers shared by the Python layers (reference: graphlearn/python/utils.py)."""
from __future__ import annotations

import enum


class Mask(enum.Enum):
    """Dataset split tag; a masked table is registered under the type name
    ``MASK<TAG>_<type>`` exactly like the reference (graphlearn/python/utils.py:44-63)."""
    NONE = 0
    TRAIN = 1
    TEST = 2
    VAL = 3


def get_mask_type(raw_type, mask=Mask.NONE):
    if mask == Mask.NONE or mask is None:
        return raw_type
    return "MASK{}_{}".format(mask.name, raw_type)


_STRATEGY_TO_OP = {
    "random": "RandomSampler",
    "random_without_replacement": "RandomWithoutReplacementSampler",
    "topk": "TopkSampler",
    "in_degree": "InDegreeSampler",
    "edge_weight": "EdgeWeightSampler",
    "full": "FullSampler",
    "node_weight": "NodeWeightNegativeSampler",
}


def strategy2op(strategy, op_type="Sampler"):
    """'edge_weight' -> 'EdgeWeightSampler', ('in_degree','NegativeSampler') -> 'InDegreeNegativeSampler'."""
    words = strategy.split("_")
    return "".join(w.capitalize() for w in words) + op_type


def ensure_list(x):
    if x is None:
        return []
    if isinstance(x, (list, tuple)):
        return list(x)
    return [x]


def deprecated(date, old, instead):
    """Decorator that prints an API-change notice on every call (graphlearn/python/utils.py:22-32)."""
    import functools

    def log_decorator(func):
        @functools.wraps(func)
        def wrapper(*args, **kwargs):
            print("[WARNING] %s will not be supported after %s, please use %s instead." % (old, date, instead))
            return func(*args, **kwargs)
        return wrapper
    return log_decorator
