"""Feature encoding (graphlearn/python/nn/tf/data/feature_column.py:34-299,
feature_handler.py:77-214): float attributes pass through, int / hashed-string attributes are
embedded (optionally hashed into buckets), same-dimension embeddings are FUSED into one table,
multi-value strings become a summed sparse embedding.  Output = concat of all pieces."""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn

from ..data.feature_spec import DenseSpec, FeatureSpec, MultivalSpec, SparseSpec


def _hash_token(tok: str) -> int:
    """Process-independent string hash (FNV-1a 64): Python's ``hash()`` is salted per process, which would map the
    same token to different buckets on different ranks and between training and serving."""
    h = 0xcbf29ce484222325
    for b in tok.encode("utf-8"):
        h = ((h ^ b) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    return h


def _hash_bucket(x: torch.Tensor, buckets: int) -> torch.Tensor:
    """deterministic integer hash -> [0, buckets)."""
    h = (x.to(torch.int64) * 2654435761) & 0x7FFFFFFFFFFFFFFF
    h = h ^ (h >> 29)
    return h % buckets


class DynamicEmbedding(nn.Module):
    """Embedding table over an UNBOUNDED id space (the reference's DynamicEmbeddingColumn / PAI-TF
    ``get_embedding_variable``, graphlearn/python/nn/tf/data/feature_column.py:160-190,281-310): no bucket size, every
    distinct key owns a row that is created the first time the key is seen.

    Keys are kept sorted in a device tensor (lookup = ``searchsorted``); unseen keys of a batch are appended and the
    weight matrix grows by doubling, so the optimizer sees an ordinary dense ``nn.Parameter`` (rebuilt only on
    growth - call ``optimizer_state_dict_hook``-free optimisers such as SGD / Adagrad-style sparse updates, or re-create
    the optimizer group when ``grew`` is set).  During evaluation unseen keys map to the zero vector."""

    def __init__(self, dim: int, initial_capacity: int = 1024, init_std: float = 0.05):
        super().__init__()
        self.embedding_dim = int(dim)
        self.init_std = float(init_std)
        self.weight = nn.Parameter(torch.zeros(int(initial_capacity), self.embedding_dim))
        self.register_buffer("keys", torch.zeros(0, dtype=torch.int64))          # sorted
        self.register_buffer("rows", torch.zeros(0, dtype=torch.int64))          # row of keys[i]
        self.grew = False

    @property
    def num_keys(self) -> int:
        return int(self.keys.numel())

    def _insert(self, new_keys: torch.Tensor):
        n_old, n_new = self.num_keys, int(new_keys.numel())
        need = n_old + n_new
        if need > self.weight.size(0):
            cap = max(need, 2 * self.weight.size(0))
            w = torch.zeros(cap, self.embedding_dim, device=self.weight.device, dtype=self.weight.dtype)
            w[:self.weight.size(0)] = self.weight.data
            self.weight = nn.Parameter(w)
            self.grew = True
        with torch.no_grad():
            self.weight.data[n_old:need].normal_(0.0, self.init_std)
        keys = torch.cat([self.keys, new_keys])
        rows = torch.cat([self.rows, torch.arange(n_old, need, device=keys.device)])
        order = torch.argsort(keys)
        self.keys, self.rows = keys[order], rows[order]

    def forward(self, ids: torch.Tensor) -> torch.Tensor:
        flat = ids.reshape(-1).to(torch.int64).to(self.weight.device).contiguous()
        if self.keys.device != flat.device:
            self.keys, self.rows = self.keys.to(flat.device), self.rows.to(flat.device)
        if self.training:
            uniq = torch.unique(flat)
            if self.num_keys:
                pos = torch.searchsorted(self.keys, uniq).clamp_(max=self.num_keys - 1)
                uniq = uniq[self.keys[pos] != uniq]
            if uniq.numel():
                self._insert(uniq)
        if self.num_keys == 0:
            return torch.zeros(tuple(ids.shape) + (self.embedding_dim,), device=flat.device)
        pos = torch.searchsorted(self.keys, flat).clamp_(max=self.num_keys - 1)
        hit = self.keys[pos] == flat
        out = self.weight[self.rows[pos]] * hit.unsqueeze(-1).to(self.weight.dtype)
        return out.reshape(tuple(ids.shape) + (self.embedding_dim,))


class FeatureEncoder(nn.Module):
    def __init__(self, spec: FeatureSpec, fuse_embedding: bool = True):
        super().__init__()
        self.spec = spec
        self.int_specs = list(spec.int_specs)
        self.n_float = spec.num_float
        self.dense_int = [i for i, s in enumerate(self.int_specs) if isinstance(s, DenseSpec)]
        self.sparse_int = [i for i, s in enumerate(self.int_specs) if isinstance(s, SparseSpec)]
        # fused tables: one nn.Embedding per embedding dim, rows = sum of bucket sizes (+ offsets)
        self.fuse = fuse_embedding
        self.groups = {}
        self.tables = nn.ModuleDict()
        self.dynamic = nn.ModuleDict()          # columns without a bucket size: unbounded key space (DynamicEmbedding)
        for i in self.sparse_int:
            s = self.int_specs[i]
            if not s.bucket_size:
                self.dynamic[str(i)] = DynamicEmbedding(int(s.dimension))
                continue
            key = str(s.dimension) if fuse_embedding else "%d_%d" % (s.dimension, i)
            self.groups.setdefault(key, []).append(i)
        self.offsets = {}
        for key, cols in self.groups.items():
            off, total = [], 0
            for i in cols:
                off.append(total)
                total += int(self.int_specs[i].bucket_size)
            self.offsets[key] = off
            self.tables[key] = nn.Embedding(total, int(self.int_specs[cols[0]].dimension))
        self.multival = nn.ModuleList([nn.EmbeddingBag(int(s.bucket_size), int(s.dimension), mode="sum")
                                       for s in spec.string_specs if isinstance(s, MultivalSpec)])

    @property
    def output_dim(self):
        d = self.n_float + len(self.dense_int)
        for key, cols in self.groups.items():
            d += len(cols) * self.tables[key].embedding_dim
        d += sum(m.embedding_dim for m in self.dynamic.values())
        d += sum(m.embedding_dim for m in self.multival)
        return d

    def forward(self, float_attrs: Optional[torch.Tensor] = None, int_attrs: Optional[torch.Tensor] = None,
                string_attrs=None) -> torch.Tensor:
        parts: List[torch.Tensor] = []
        if self.n_float and float_attrs is not None:
            parts.append(float_attrs.float().reshape(-1, self.n_float))
        if int_attrs is not None and self.int_specs:
            ia = int_attrs.reshape(-1, len(self.int_specs))
            if self.dense_int:
                parts.append(ia[:, self.dense_int].float())
            for key, cols in self.groups.items():
                idx = []
                for c, off in zip(cols, self.offsets[key]):
                    s = self.int_specs[c]
                    v = _hash_bucket(ia[:, c], int(s.bucket_size)) if s.need_hash else ia[:, c].clamp(0, int(s.bucket_size) - 1)
                    idx.append(v + off)
                emb = self.tables[key](torch.stack(idx, 1))              # [n, cols, dim]  one fused lookup
                parts.append(emb.reshape(emb.size(0), -1))
            for i, m in self.dynamic.items():
                parts.append(m(ia[:, int(i)]))
        if self.multival and string_attrs is not None:
            for j, m in enumerate(self.multival):
                spec = [s for s in self.spec.string_specs if isinstance(s, MultivalSpec)][j]
                flat, offs = [], [0]
                for row in string_attrs[:, j]:
                    toks = [t for t in str(row).split(spec.delimiter) if t]
                    flat.extend(_hash_token(t) % int(spec.bucket_size) for t in toks)
                    offs.append(len(flat))
                dev = m.weight.device
                parts.append(m(torch.tensor(flat, dtype=torch.long, device=dev), torch.tensor(offs[:-1], dtype=torch.long, device=dev)))
        return torch.cat(parts, 1) if parts else torch.zeros(0)
