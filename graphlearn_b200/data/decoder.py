"""Data-source schema description (``gl.Decoder``).

Same constructor and properties as the reference
(graphlearn/python/data/decoder.py:26-260): column order is
id | (src, dst), weight, label, timestamp, attributes; the attribute column is
one delimiter-joined string whose fields are typed by ``attr_types``.
"""
from __future__ import annotations

from .feature_spec import FeatureSpec

_TYPE_CODE = {"int": 0, "float": 1, "string": 2}


class Decoder(object):
    def __init__(self, weighted=False, labeled=False, timestamped=False, attr_types=None, attr_delimiter=":",
                 attr_dims=None):
        self._weighted = bool(weighted)
        self._labeled = bool(labeled)
        self._timestamped = bool(timestamped)
        self._attr_types = list(attr_types) if attr_types else []
        if attr_types is not None and not isinstance(attr_types, (list, tuple)):
            raise ValueError("attr_types for Decoder must be a list, got {}.".format(type(attr_types)))
        self._attr_delimiter = attr_delimiter
        self._attr_dims = list(attr_dims) if attr_dims else []
        self._int_attr_num = 0
        self._float_attr_num = 0
        self._string_attr_num = 0
        self._fspec = None
        self._parsed = [self.parse(t) for t in self._attr_types]
        for name, bucket, multival in self._parsed:
            if name == "int":
                self._int_attr_num += 1
            elif name == "float":
                self._float_attr_num += 1
            elif multival or bucket is None:
                self._string_attr_num += 1
            else:                       # hashed string becomes an int attribute
                self._int_attr_num += 1
        self._attributed = len(self._attr_types) > 0

    @staticmethod
    def parse(attr_type):
        if isinstance(attr_type, (tuple, list)):
            name = attr_type[0]
            bucket = attr_type[1] if len(attr_type) >= 2 else None
            multival = attr_type[2] if len(attr_type) >= 3 else False
        else:
            name, bucket, multival = attr_type, None, False
        if name not in _TYPE_CODE:
            raise ValueError("attr type must be one of int/float/string, got %r" % (name,))
        if multival and name != "string":
            raise ValueError("multi-value attribute must be string type.")
        return name, bucket, multival

    # ---- properties (reference names)
    @property
    def has_property(self):
        return self._weighted or self._labeled or self._timestamped or self._attributed

    weighted = property(lambda self: self._weighted)
    labeled = property(lambda self: self._labeled)
    timestamped = property(lambda self: self._timestamped)
    attributed = property(lambda self: self._attributed)
    attr_types = property(lambda self: self._attr_types)
    attr_delimiter = property(lambda self: self._attr_delimiter)
    attr_dims = property(lambda self: self._attr_dims)
    int_attr_num = property(lambda self: self._int_attr_num)
    float_attr_num = property(lambda self: self._float_attr_num)
    string_attr_num = property(lambda self: self._string_attr_num)

    @property
    def data_format(self):
        # attributed << 4 | timestamped << 3 | labeled << 2 | weighted << 1 (decoder.py:199-203)
        return int(self._weighted * 2 + self._labeled * 4 + self._timestamped * 8 + self._attributed * 16)

    @property
    def feature_spec(self):
        if self._fspec is None:
            self._build_feature_spec()
        return self._fspec

    # ---- native-loader view
    def loader_schema(self):
        """(type codes, hash buckets) for csrc/host_loader.cpp."""
        codes, buckets = [], []
        for name, bucket, multival in self._parsed:
            codes.append(_TYPE_CODE[name])
            buckets.append(int(bucket) if (name == "string" and bucket and not multival) else 0)
        return codes, buckets

    def _build_feature_spec(self):
        n = len(self._attr_types)
        dims = self._attr_dims or [None] * n
        if len(dims) != n:
            raise ValueError("The size of attr_dims must be equal with attr_types.")
        spec = FeatureSpec(n, self._weighted, self._labeled, self._timestamped)
        for (name, bucket, multival), dim in zip(self._parsed, dims):
            if multival:
                spec.append_multival(bucket, dim, ",")
            elif dim:
                assert name in ("int", "string"), "Must assign an attr_dim with None for {}".format(name)
                spec.append_sparse(bucket, dim, name == "int")
            else:
                assert name in ("int", "float") and bucket is None, \
                    "Must assign an attr_dim for {}, and bucket_size should None.".format(name)
                spec.append_dense(name == "float")
        self._fspec = spec

    def format_attrs(self, int_attrs, float_attrs, string_attrs):
        if int_attrs is not None:
            int_attrs = int_attrs.reshape(-1, self._int_attr_num)
        if float_attrs is not None:
            float_attrs = float_attrs.reshape(-1, self._float_attr_num)
        if string_attrs is not None:
            string_attrs = string_attrs.reshape(-1, self._string_attr_num)
        return int_attrs, float_attrs, string_attrs
