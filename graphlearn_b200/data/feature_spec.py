"""Feature specification used by the feature encoders
(reference: graphlearn/python/data/feature_spec.py)."""
from __future__ import annotations


class DenseSpec(object):
    def __init__(self, is_float):
        self.is_float = is_float
        self.dimension = 1


class SparseSpec(object):
    def __init__(self, bucket_size, dimension, need_hash):
        self.bucket_size = bucket_size
        self.dimension = dimension
        self.need_hash = need_hash


class MultivalSpec(object):
    def __init__(self, bucket_size, dimension, delimiter=","):
        self.bucket_size = bucket_size
        self.dimension = dimension
        self.delimiter = delimiter


class DynamicSparseSpec(object):
    """Embedding column without a fixed bucket count (PAI-TF dynamic embedding in the reference,
    feature_spec.py); accepted for script parity, ``nn.FeatureEncoder`` needs a bucket size."""

    def __init__(self, dimension, need_hash=True):
        self.bucket_size = None
        self.dimension = dimension
        self.need_hash = need_hash


class DynamicMultivalSpec(object):
    def __init__(self, dimension, delimiter=","):
        self.bucket_size = None
        self.dimension = dimension
        self.delimiter = delimiter


class FeatureSpec(object):
    def __init__(self, size, weighted=False, labeled=False, timestamped=False):
        self.size = size
        self.weighted, self.labeled, self.timestamped = weighted, labeled, timestamped
        self.int_specs, self.float_specs, self.string_specs = [], [], []
        self.specs = []

    def append_dense(self, is_float=True):
        s = DenseSpec(is_float)
        (self.float_specs if is_float else self.int_specs).append(s)
        self.specs.append(s)

    def append_sparse(self, bucket_size, dimension, need_hash):
        s = SparseSpec(bucket_size, dimension, need_hash)
        # hashed strings are stored as int attributes by the loader
        self.int_specs.append(s)
        self.specs.append(s)

    def append_multival(self, bucket_size, dimension, delimiter=","):
        s = MultivalSpec(bucket_size, dimension, delimiter)
        self.string_specs.append(s)
        self.specs.append(s)

    @property
    def num_int(self):
        return len(self.int_specs)

    @property
    def num_float(self):
        return len(self.float_specs)

    @property
    def num_string(self):
        return len(self.string_specs)
