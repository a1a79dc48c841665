"""Temporal Graph Network (memory + temporal attention) and its batch loader.

Capability parity with graphlearn/examples/pytorch/tgn/{train_and_eval,temporal_batch_loader}.py,
which drive PyG's ``TGNMemory`` / ``TransformerConv`` from GSL temporal queries.  PyG is not
part of this image, so the model is written out in plain PyTorch:

  memory      one state row per node + last-update time; pending raw messages are aggregated with
              "last message wins" and folded in by a GRU cell (IdentityMessage + LastAggregator)
  embedding   2-head attention over the k most recent interactions before the event
              (edge features = time encoding of the age of the interaction ++ its message vector)
  decoder     MLP link predictor on (src, dst) embeddings

All tensors live on the graph's device; the sampler side is the temporal GSL query
``E(events) -> outV/inV -> outE/inE.sample(k).by("topk")`` whose time filter (edges strictly before
the event) and most-recent-first order are applied inside the sampling kernel.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..nn.conv import TimeEncoder
from ..nn.sparse_conv import segment_softmax


class TGNMemory(nn.Module):
    def __init__(self, num_nodes: int, raw_msg_dim: int, memory_dim: int, time_dim: int):
        super().__init__()
        self.num_nodes, self.raw_msg_dim, self.memory_dim = num_nodes, raw_msg_dim, memory_dim
        self.time_enc = TimeEncoder(time_dim)
        self.gru = nn.GRUCell(2 * memory_dim + raw_msg_dim + time_dim, memory_dim)
        self.register_buffer("memory", torch.zeros(num_nodes, memory_dim))
        self.register_buffer("last_update", torch.zeros(num_nodes, dtype=torch.long))
        # pending (not yet folded) message per node: the most recent event the node took part in
        self.register_buffer("p_valid", torch.zeros(num_nodes, dtype=torch.bool))
        self.register_buffer("p_other", torch.zeros(num_nodes, dtype=torch.long))
        self.register_buffer("p_t", torch.zeros(num_nodes, dtype=torch.long))
        self.register_buffer("p_raw", torch.zeros(num_nodes, raw_msg_dim))

    def reset_state(self):
        self.memory.zero_(); self.last_update.zero_(); self.p_valid.zero_()

    def detach(self):
        self.memory.detach_()

    def _updated(self, n_id: torch.Tensor):
        """memory / last_update of ``n_id`` after folding their pending message (no state change)."""
        mem, lu = self.memory[n_id], self.last_update[n_id]
        has = self.p_valid[n_id]
        if bool(has.any()):
            idx = n_id[has]
            t = self.p_t[idx]
            rel = (t - self.last_update[idx]).to(mem.dtype)
            msg = torch.cat([self.memory[idx], self.memory[self.p_other[idx]], self.p_raw[idx], self.time_enc(rel)], -1)
            new = self.gru(msg, self.memory[idx])
            mem = mem.clone()
            mem[has] = new
            lu = lu.clone()
            lu[has] = t
        return mem, lu

    def forward(self, n_id: torch.Tensor):
        return self._updated(n_id)

    @torch.no_grad()
    def _fold(self, n_id: torch.Tensor):
        mem, lu = self._updated(n_id)
        self.memory[n_id] = mem.detach()
        self.last_update[n_id] = lu
        self.p_valid[n_id] = False

    def update_state(self, src: torch.Tensor, dst: torch.Tensor, t: torch.Tensor, raw_msg: torch.Tensor):
        """Fold the pending messages of the nodes of this batch, then store the batch's events as
        their new pending messages (a node that occurs several times keeps its LAST event)."""
        n_id = torch.unique(torch.cat([src, dst]))
        if self.training:
            # keep the graph of this fold alive until detach(): gradients reach the GRU through the
            # NEXT batch's forward(), exactly like the reference's training loop
            mem, lu = self._updated(n_id)
            self.memory = self.memory.clone()
            self.memory[n_id] = mem
            self.last_update[n_id] = lu
            self.p_valid[n_id] = False
        else:
            self._fold(n_id)
        node = torch.cat([src, dst])
        other = torch.cat([dst, src])
        tt = torch.cat([t, t])
        raw = torch.cat([raw_msg, raw_msg])
        pos = torch.arange(node.numel(), device=node.device)
        last = torch.full((self.num_nodes,), -1, dtype=torch.long, device=node.device)
        last.scatter_reduce_(0, node, pos, reduce="amax", include_self=True)
        keep = last[node] == pos
        nk = node[keep]
        self.p_valid[nk] = True
        self.p_other[nk] = other[keep]
        self.p_t[nk] = tt[keep]
        self.p_raw[nk] = raw[keep].to(self.p_raw.dtype)


class TemporalAttention(nn.Module):
    """Multi-head attention of a node over its recent interactions with edge features
    (the role of PyG's TransformerConv in the reference's GraphAttentionEmbedding)."""

    def __init__(self, in_dim: int, out_dim: int, edge_dim: int, heads: int = 2, dropout: float = 0.1):
        super().__init__()
        assert out_dim % heads == 0
        self.h, self.c = heads, out_dim // heads
        self.q = nn.Linear(in_dim, out_dim)
        self.k = nn.Linear(in_dim, out_dim)
        self.v = nn.Linear(in_dim, out_dim)
        self.e = nn.Linear(edge_dim, out_dim, bias=False)
        self.skip = nn.Linear(in_dim, out_dim)
        self.dropout = dropout

    def forward(self, x, edge_index, edge_attr):
        j, i = edge_index[0], edge_index[1]          # message j -> i
        n = x.size(0)
        e = self.e(edge_attr).view(-1, self.h, self.c)
        q = self.q(x).view(n, self.h, self.c)[i]
        k = self.k(x).view(n, self.h, self.c)[j] + e
        v = self.v(x).view(n, self.h, self.c)[j] + e
        score = (q * k).sum(-1) / math.sqrt(self.c)            # [E, h]
        alpha = torch.stack([segment_softmax(score[:, h], i, n) for h in range(self.h)], 1)
        alpha = F.dropout(alpha, self.dropout, self.training)
        out = torch.zeros(n, self.h, self.c, device=x.device, dtype=x.dtype)
        out.index_add_(0, i, alpha.unsqueeze(-1) * v)
        return out.reshape(n, -1) + self.skip(x)


class LinkPredictor(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.lin_src, self.lin_dst, self.lin_final = nn.Linear(dim, dim), nn.Linear(dim, dim), nn.Linear(dim, 1)

    def forward(self, z_src, z_dst):
        return self.lin_final(F.relu(self.lin_src(z_src) + self.lin_dst(z_dst)))


class TGN(nn.Module):
    def __init__(self, num_nodes: int, msg_dim: int, memory_dim: int = 100, time_dim: int = 100,
                 embedding_dim: int = 100):
        super().__init__()
        self.memory = TGNMemory(num_nodes, msg_dim, memory_dim, time_dim)
        self.gnn = TemporalAttention(memory_dim, embedding_dim, msg_dim + time_dim)
        self.link_pred = LinkPredictor(embedding_dim)
        self.register_buffer("assoc", torch.zeros(num_nodes, dtype=torch.long))

    def forward(self, batch: "TemporalBatch"):
        """-> (pos_logit [B,1], neg_logit [B,1]); call ``update(batch)`` afterwards."""
        n_id = batch.n_id
        self.assoc[n_id] = torch.arange(n_id.numel(), device=n_id.device)
        z, last_update = self.memory(n_id)
        rel_t = (last_update[batch.edge_index[0]] - batch.nbr_t).to(z.dtype)
        edge_attr = torch.cat([self.memory.time_enc(rel_t), batch.nbr_msg], -1)
        z = self.gnn(z, batch.edge_index, edge_attr)
        s = z[self.assoc[batch.src]]
        return self.link_pred(s, z[self.assoc[batch.pos_dst]]), self.link_pred(s, z[self.assoc[batch.neg_dst]])

    def loss(self, batch):
        pos, neg = self(batch)
        return F.binary_cross_entropy_with_logits(pos, torch.ones_like(pos)) + \
            F.binary_cross_entropy_with_logits(neg, torch.zeros_like(neg))

    def update(self, batch):
        self.memory.update_state(batch.src, batch.pos_dst, batch.t, batch.msg)


class TemporalBatch(object):
    """One chronological batch of events + the induced recent-interaction graph
    (temporal_batch_loader.py:26-51)."""

    def __init__(self, src, pos_dst, neg_dst, t, msg, n_id, edge_index, nbr_t, nbr_msg):
        self.src, self.pos_dst, self.neg_dst, self.t, self.msg = src, pos_dst, neg_dst, t, msg
        self.n_id, self.edge_index, self.nbr_t, self.nbr_msg = n_id, edge_index, nbr_t, nbr_msg
        self.num_events = int(src.numel())


class TemporalBatchLoader(object):
    """Iterates ``TemporalBatch`` objects for one pass over the events of edge type ``source``.

    GSL plan (temporal_batch_loader.py:71-82): event edges in insertion (= chronological) order,
    one random negative destination per event, and for src / pos_dst / neg_dst the ``nbr_size`` most
    recent interactions strictly before the event time.  Node ids must be dense in [0, num_nodes)."""

    def __init__(self, graph, source: str, num_nodes: int, batch_size: int, nbr_size: int, msg_dim: int,
                 interaction: str = "interaction"):
        from ..gsl.dataset import Dataset
        self.nbr_size, self.msg_dim, self.num_nodes = nbr_size, msg_dim, num_nodes
        ev = graph.E(source).batch(batch_size).alias("event")
        s = ev.outV().alias("src")
        d = ev.inV().alias("pos_dst")
        n = s.outNeg(interaction).sample(1).by("random").alias("neg_dst")
        s.outE(interaction).sample(nbr_size).by("topk").alias("src_nbr")
        d.inE(interaction).sample(nbr_size).by("topk").alias("dst_nbr")
        n.inE(interaction).sample(nbr_size).by("topk").alias("neg_nbr")
        self._ds = Dataset(ev.values())

    def __iter__(self):
        from .. import errors
        while True:
            try:
                r = self._ds.next()
            except errors.OutOfRangeError:
                return
            yield self._induce(r)

    def _induce(self, r) -> TemporalBatch:
        e = r["event"]
        src, pos_dst = e.tensor("src_ids").reshape(-1), e.tensor("dst_ids").reshape(-1)
        neg_dst = r["neg_dst"].ids_t.reshape(-1)
        t = e.tensor("timestamps").reshape(-1)
        msg = e.tensor("float_attrs").reshape(-1, self.msg_dim)
        centers = torch.cat([src, pos_dst, neg_dst])
        k = self.nbr_size
        nbr = torch.cat([r[a].tensor("dst_ids").reshape(-1, k) for a in ("src_nbr", "dst_nbr", "neg_nbr")])
        nt = torch.cat([r[a].tensor("timestamps").reshape(-1, k) for a in ("src_nbr", "dst_nbr", "neg_nbr")])
        nm = torch.cat([r[a].tensor("float_attrs").reshape(-1, k, self.msg_dim) for a in ("src_nbr", "dst_nbr", "neg_nbr")])
        # one neighbour row per distinct node: the LAST occurrence wins, like the reference's
        # ``topo[src], topo[pos_dst], topo[neg_dst] = ...`` assignment (temporal_batch_loader.py:97-99)
        pos = torch.arange(centers.numel(), device=centers.device)
        last = torch.full((self.num_nodes,), -1, dtype=torch.long, device=centers.device)
        last.scatter_reduce_(0, centers, pos, reduce="amax", include_self=True)
        uniq = torch.unique(centers)
        rows = last[uniq]
        nbr, nt, nm = nbr[rows], nt[rows], nm[rows]
        mask = nbr >= 0
        ctr = uniq[:, None].expand(-1, k)[mask]
        nb = nbr[mask]
        n_id = torch.unique(torch.cat([uniq, nb]))
        assoc = torch.zeros(self.num_nodes, dtype=torch.long, device=n_id.device)
        assoc[n_id] = torch.arange(n_id.numel(), device=n_id.device)
        edge_index = torch.stack([assoc[nb], assoc[ctr]])
        return TemporalBatch(src, pos_dst, neg_dst, t, msg, n_id, edge_index, nt[mask], nm[mask])
