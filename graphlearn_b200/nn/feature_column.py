"""Feature columns, groups and the spec-driven handler - the column-level API of the reference's feature encoding
(graphlearn/python/nn/tf/data/feature_column.py:34-310, feature_handler.py:31-214) as torch Modules.

* ``NumericColumn``                 continuous value, optional normaliser
* ``EmbeddingColumn``               categorical id -> embedding row (``need_hash``: hashed into the buckets first)
* ``FusedEmbeddingColumn``          several categorical features of one embedding width share ONE table (one lookup)
* ``SparseEmbeddingColumn``         multi-value string ("a,b,c") -> sum of the token embeddings
* ``DynamicEmbeddingColumn`` / ``DynamicSparseEmbeddingColumn``   the same over an unbounded vocabulary (no bucket size;
  rows are created when a key is first seen: ``nn.DynamicEmbedding``)
* ``FeatureGroup``                  applies column i to input i and concatenates
* ``FeatureHandler``                builds the groups from a ``FeatureSpec`` and encodes a ``Data`` object (float / int /
  string attributes) exactly in the reference's order: floats, per-feature int columns, fused int columns, strings

``PartitionableColumn`` exists for script parity: the reference partitions big TF variables over parameter servers; here
big tables are sharded over GPUs by ``nn.ShardedEmbedding`` and a column's table stays one Parameter.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch
import torch.nn as nn

from ..data.feature_spec import (DenseSpec, DynamicMultivalSpec, DynamicSparseSpec, FeatureSpec, MultivalSpec, SparseSpec)
from .feature import DynamicEmbedding, _hash_bucket, _hash_token


def _as_list_of_str(x) -> List[str]:
    if isinstance(x, torch.Tensor):
        return [str(v) for v in x.reshape(-1).tolist()]
    import numpy as np
    out = []
    for v in np.asarray(x, dtype=object).reshape(-1).tolist():
        out.append(v.decode("utf-8") if isinstance(v, bytes) else str(v))
    return out


class FeatureColumn(nn.Module):
    """Base class: a column turns one raw feature ([batch]) into a dense tensor ([batch] or [batch, dim])."""

    def __init__(self, name: str = ""):
        super().__init__()
        self.name = name

    @property
    def output_dim(self) -> int:
        raise NotImplementedError


class PartitionableColumn(FeatureColumn):
    """Parity shim for the reference's variable partitioners (feature_column.py:75-98)."""

    def _partitioner(self, partitioner="min_max"):
        return None


class NumericColumn(FeatureColumn):
    def __init__(self, name: str, normalizer_func: Optional[Callable[[torch.Tensor], torch.Tensor]] = None):
        super().__init__(name)
        self._norm = normalizer_func

    @property
    def output_dim(self):
        return 1

    def forward(self, x):
        x = torch.as_tensor(x).float()
        return self._norm(x) if self._norm is not None else x


class EmbeddingColumn(PartitionableColumn):
    def __init__(self, name: str, bucket_size: int, dimension: int, need_hash: bool = False):
        super().__init__(name)
        self.bucket_size, self.dimension, self.need_hash = int(bucket_size), int(dimension), bool(need_hash)
        self.table = nn.Embedding(self.bucket_size, self.dimension)

    @property
    def output_dim(self):
        return self.dimension

    def forward(self, x):
        x = torch.as_tensor(x).to(torch.int64).to(self.table.weight.device)
        idx = _hash_bucket(x, self.bucket_size) if self.need_hash else x.clamp(0, self.bucket_size - 1)
        return self.table(idx)


class DynamicEmbeddingColumn(PartitionableColumn):
    """``is_string``: inputs are strings (hashed to 64-bit keys on the host), else integer keys."""

    def __init__(self, name: str, dimension: int, is_string: bool = False):
        super().__init__(name)
        self.dimension, self.is_string = int(dimension), bool(is_string)
        self.table = DynamicEmbedding(self.dimension)

    @property
    def output_dim(self):
        return self.dimension

    def forward(self, x):
        if self.is_string:
            keys = torch.tensor([_hash_token(t) & 0x7FFFFFFFFFFFFFFF for t in _as_list_of_str(x)], dtype=torch.int64)
        else:
            keys = torch.as_tensor(x).to(torch.int64).reshape(-1)
        return self.table(keys)


class FusedEmbeddingColumn(PartitionableColumn):
    """forward(list of [batch] id tensors, one per fused feature) -> [batch, len(list) * dimension]."""

    def __init__(self, name: str, bucket_list: Sequence[int], dimension: int):
        super().__init__(name)
        self.bucket_list, self.dimension = [int(b) for b in bucket_list], int(dimension)
        offs, tot = [], 0
        for b in self.bucket_list:
            offs.append(tot)
            tot += b
        self.register_buffer("offsets", torch.tensor(offs, dtype=torch.int64))
        self.register_buffer("limits", torch.tensor(self.bucket_list, dtype=torch.int64))
        self.table = nn.Embedding(tot, self.dimension)

    @property
    def output_dim(self):
        return self.dimension * len(self.bucket_list)

    def forward(self, xs):
        if isinstance(xs, torch.Tensor):
            ids = xs.to(torch.int64).reshape(-1, len(self.bucket_list))
        else:
            ids = torch.stack([torch.as_tensor(x).to(torch.int64).reshape(-1) for x in xs], 1)
        ids = ids.to(self.table.weight.device)
        ids = torch.minimum(ids.clamp(min=0), self.limits - 1) + self.offsets
        e = self.table(ids)                                      # one lookup for all fused features
        return e.reshape(e.size(0), -1)


class _TokenBag(PartitionableColumn):
    def __init__(self, name: str, dimension: int, delimiter: str):
        super().__init__(name)
        self.dimension, self.delimiter = int(dimension), delimiter

    @property
    def output_dim(self):
        return self.dimension

    def _tokens(self, x):
        keys, seg = [], []
        rows = _as_list_of_str(x)
        for r, s in enumerate(rows):
            for t in s.split(self.delimiter):
                if t:
                    keys.append(_hash_token(t) & 0x7FFFFFFFFFFFFFFF)
                    seg.append(r)
        return len(rows), torch.tensor(keys, dtype=torch.int64), torch.tensor(seg, dtype=torch.int64)


class SparseEmbeddingColumn(_TokenBag):
    def __init__(self, name: str, bucket_size: int, dimension: int, delimiter: str = ","):
        super().__init__(name, dimension, delimiter)
        self.bucket_size = int(bucket_size)
        self.table = nn.Embedding(self.bucket_size, self.dimension)

    def forward(self, x):
        n, keys, seg = self._tokens(x)
        dev = self.table.weight.device
        out = torch.zeros(n, self.dimension, device=dev, dtype=self.table.weight.dtype)
        if keys.numel():
            out = out.index_add(0, seg.to(dev), self.table((keys % self.bucket_size).to(dev)))
        return out


class DynamicSparseEmbeddingColumn(_TokenBag):
    def __init__(self, name: str, dimension: int, delimiter: str = ","):
        super().__init__(name, dimension, delimiter)
        self.table = DynamicEmbedding(self.dimension)

    def forward(self, x):
        n, keys, seg = self._tokens(x)
        dev = self.table.weight.device
        out = torch.zeros(n, self.dimension, device=dev, dtype=self.table.weight.dtype)
        if keys.numel():
            out = out.index_add(0, seg.to(dev), self.table(keys))
        return out


class FeatureGroup(nn.Module):
    """A list of columns applied position-wise (feature_handler.py:31-74)."""

    def __init__(self, feature_column_list: Sequence[FeatureColumn]):
        super().__init__()
        self.columns = nn.ModuleList(list(feature_column_list))

    def __len__(self):
        return len(self.columns)

    def __bool__(self):
        return len(self.columns) > 0

    def __getitem__(self, i):
        return self.columns[i]

    @property
    def output_dim(self):
        return sum(c.output_dim for c in self.columns)

    def forward(self, x_list):
        if isinstance(x_list, torch.Tensor):
            x_list = [x_list[..., i] for i in range(x_list.shape[-1])]
        elif not isinstance(x_list, (list, tuple)):
            import numpy as np
            arr = np.asarray(x_list, dtype=object)
            x_list = [arr[..., i] for i in range(arr.shape[-1])]
        if len(x_list) != len(self.columns):
            raise ValueError("%d feature columns, but got %d inputs." % (len(self.columns), len(x_list)))
        outs = []
        for c, x in zip(self.columns, x_list):
            o = c(x)
            outs.append(o.unsqueeze(-1) if isinstance(c, NumericColumn) else o)
        dev = next((o.device for o in outs if o.device.type != "cpu"), outs[0].device)
        return torch.cat([o.to(dev) for o in outs], -1)


class FeatureHandler(nn.Module):
    """Encode the attributes of a ``Data``-like object (``float_attrs [n, F]``, ``int_attrs [n, I]``, ``string_attrs [n, S]``)
    as described by a ``FeatureSpec`` (feature_handler.py:77-214)."""

    def __init__(self, name: str, feature_spec: FeatureSpec, fuse_embedding: bool = True):
        super().__init__()
        self.name, self._fspec, self._fuse = name, feature_spec, fuse_embedding
        self._int_mapping: List[int] = []
        fused = {}                                                 # dimension -> (indices, buckets)
        int_cols: List[FeatureColumn] = []
        for i, spec in enumerate(feature_spec.int_specs):
            if isinstance(spec, DynamicSparseSpec) or (isinstance(spec, SparseSpec) and spec.bucket_size is None):
                int_cols.append(DynamicEmbeddingColumn("dynamic_int_emb_%d" % i, spec.dimension, is_string=False))
                self._int_mapping.append(i)
            elif isinstance(spec, SparseSpec):
                if spec.need_hash or not fuse_embedding:
                    int_cols.append(EmbeddingColumn("emb_%d" % i, spec.bucket_size, spec.dimension, spec.need_hash))
                    self._int_mapping.append(i)
                elif spec.dimension:
                    idx, buckets = fused.setdefault(int(spec.dimension), ([], []))
                    idx.append(i)
                    buckets.append(int(spec.bucket_size))
            else:                                                  # an integer treated like a float
                int_cols.append(NumericColumn("sparse_as_dense_%d" % i))
                self._int_mapping.append(i)
        self._fused_index = [idx for _, (idx, _) in fused.items()]
        self._float_fg = FeatureGroup([NumericColumn("dense_%d" % i) for i in range(len(feature_spec.float_specs))])
        self._int_fg = FeatureGroup(int_cols)
        self._fused_int_fg = FeatureGroup([FusedEmbeddingColumn("fused_emb_%d" % d, b, d) for d, (_, b) in fused.items()])
        str_cols: List[FeatureColumn] = []
        for i, spec in enumerate(feature_spec.string_specs):
            if isinstance(spec, DynamicMultivalSpec):
                str_cols.append(DynamicSparseEmbeddingColumn("dynamic_sparse_emb_%d" % i, spec.dimension, spec.delimiter))
            elif isinstance(spec, MultivalSpec):
                str_cols.append(SparseEmbeddingColumn("sparse_emb_%d" % i, spec.bucket_size, spec.dimension, spec.delimiter))
            elif isinstance(spec, (DynamicSparseSpec, SparseSpec)):
                str_cols.append(DynamicEmbeddingColumn("dynamic_str_emb_%d" % i, spec.dimension, is_string=True))
            else:
                raise ValueError("unsupported string feature spec %r" % (spec,))
        self._string_fg = FeatureGroup(str_cols)

    @property
    def output_dim(self) -> int:
        return self._float_fg.output_dim + self._int_fg.output_dim + self._fused_int_fg.output_dim + self._string_fg.output_dim

    def forward(self, data) -> torch.Tensor:
        outs = []
        fa, ia, sa = getattr(data, "float_attrs", None), getattr(data, "int_attrs", None), getattr(data, "string_attrs", None)
        if self._float_fg:
            fa = torch.as_tensor(fa).float()
            outs.append(self._float_fg(fa.reshape(-1, fa.shape[-1])))
        if self._int_fg or self._fused_int_fg:
            ia = torch.as_tensor(ia).to(torch.int64)
            ia = ia.reshape(-1, ia.shape[-1])
        if self._int_fg:
            outs.append(self._int_fg([ia[:, i] for i in self._int_mapping]))
        if self._fused_int_fg:
            outs.append(self._fused_int_fg([[ia[:, i] for i in idx] for idx in self._fused_index]))
        if self._string_fg:
            import numpy as np
            sarr = np.asarray(sa, dtype=object)
            outs.append(self._string_fg(sarr.reshape(-1, sarr.shape[-1])))
        if not outs:
            return torch.zeros(0)
        dev = next((o.device for o in outs if o.device.type != "cpu"), outs[0].device)
        return torch.cat([o.to(dev) for o in outs], -1)
