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


class DynamicSparseSpec(SparseSpec):
    """Embedding column without a fixed bucket count (PAI-TF dynamic embedding variables in the reference,
    feature_spec.py:28-31): every distinct key owns a row (``nn.DynamicEmbedding``)."""

    def __init__(self, dimension, need_hash=True):
        super().__init__(None, dimension, need_hash)


class DynamicMultivalSpec(MultivalSpec):
    """Multi-value string column over an unbounded vocabulary (feature_spec.py:43-46)."""

    def __init__(self, dimension, delimiter=","):
        super().__init__(None, dimension, delimiter)


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

    def append_sparse(self, bucket_size, dimension, need_hash=False):
        """bucket_size None = dynamic vocabulary: hashed (int) keys stay int attributes, raw strings are string
        attributes (feature_spec.py:96-107)."""
        if bucket_size is not None:
            s = SparseSpec(bucket_size, dimension, need_hash)
            self.int_specs.append(s)                  # hashed strings are stored as int attributes by the loader
        else:
            s = DynamicSparseSpec(dimension, need_hash)
            (self.int_specs if need_hash else self.string_specs).append(s)
        self.specs.append(s)

    def append_multival(self, bucket_size, dimension, delimiter=","):
        s = MultivalSpec(bucket_size, dimension, delimiter) if bucket_size is not None else DynamicMultivalSpec(dimension, delimiter)
        self.string_specs.append(s)
        self.specs.append(s)

    @property
    def dimension(self):
        """width of the encoded feature vector (feature_spec.py:93-94)"""
        return sum(int(s.dimension) for s in self.specs)

    @property
    def num_int(self):
        return len(self.int_specs)

    @property
    def num_float(self):
        return len(self.float_specs)

    @property
    def num_string(self):
        return len(self.string_specs)
