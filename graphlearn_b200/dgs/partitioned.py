"""Partitioned streaming service (D1: ``Partitioner`` + ``PartitionRouter``, src/common/partitioner.h,
partition_router.h): the vertex id space is hash-partitioned (``vid % P``) over P sample stores - one per GPU of the
box (or P logical partitions on one device).  Updates are routed to the partition that owns their SOURCE vertex
(edges) / the vertex itself (features); a query walks the plan hop by hop, sending every frontier vertex to its
owner and stitching the answers back in request order - the reference's sampling-worker -> serving-worker
forwarding through Kafka becomes an index_select / index_copy per partition."""
from __future__ import annotations

import time
from typing import Dict, List, Optional, Sequence

import torch

from .plan import QueryPlan
from .service import AdaptiveRateLimiter, DynamicGraphService


class Partitioner(object):
    def __init__(self, num_partitions: int):
        self.P = int(num_partitions)

    def __call__(self, vids: torch.Tensor) -> torch.Tensor:
        return vids.abs() % self.P


class PartitionedGraphService(object):
    def __init__(self, schema: Dict[str, dict], num_partitions: int = 2, devices: Optional[Sequence] = None):
        self.schema = schema
        self.partitioner = Partitioner(num_partitions)
        P = self.partitioner.P
        devices = list(devices) if devices else [None] * P
        assert len(devices) == P
        self.parts: List[DynamicGraphService] = [DynamicGraphService(schema, device=d) for d in devices]
        self.device = self.parts[0].device          # answers are stitched on partition 0's device
        self.queries: Dict[int, QueryPlan] = {}
        self.limiter = AdaptiveRateLimiter()
        self.ingested = 0
        self.served = 0

    def install_query(self, qid: int, plan: QueryPlan):
        for p in self.parts:
            p.install_query(qid, plan)
        self.queries[qid] = plan

    # ------------------------------------------------------------------ ingest
    def apply_updates(self, batch: dict):
        P = self.partitioner.P
        for etype, rec in batch.get("edges", {}).items():
            src = torch.as_tensor(rec["src"])
            owner = self.partitioner(src)
            for p in range(P):
                m = owner == p
                if bool(m.any()):
                    sub = {k: torch.as_tensor(v)[m] for k, v in rec.items() if v is not None}
                    self.parts[p].apply_updates({"edges": {etype: sub}})
            self.ingested += int(src.numel())
        for vt, rec in batch.get("vertices", {}).items():
            vid = torch.as_tensor(rec["id"])
            owner = self.partitioner(vid)
            for p in range(P):
                m = owner == p
                if bool(m.any()):
                    self.parts[p].apply_updates({"vertices": {vt: {k: torch.as_tensor(v)[m] for k, v in rec.items()}}})

    # ------------------------------------------------------------------ serve
    def _features(self, vtype: str, vids: torch.Tensor):
        d = self.parts[0].vstores[vtype]
        if d.feat is None:
            return None
        flat = vids.reshape(-1)
        out = torch.zeros(flat.numel(), d.feat.size(1), device=self.device)
        owner = self.partitioner(flat.clamp(min=0))
        for p, part in enumerate(self.parts):
            vs = part.vstores[vtype]
            m = (owner == p) & (flat >= 0) & (flat < vs.n)
            if bool(m.any()):
                out[m] = vs.feat[flat[m].to(vs.device)].to(self.device)
        return out.reshape(tuple(vids.shape) + (-1,))

    def run_query(self, qid: int, vids) -> dict:
        t0 = time.perf_counter()
        plan = self.queries[qid]
        src = torch.as_tensor(list(vids) if not isinstance(vids, torch.Tensor) else vids, dtype=torch.int64).to(self.device)
        out = {"src": src, "hops": [], "nodes": {}}
        cur_ids = {0: (src, plan.source_type)}
        for nid in plan.topo_order():
            node = plan.nodes[nid]
            if node.kind == "SOURCE":
                continue
            cur, cur_type = cur_ids[node.parent]
            flat = cur.reshape(-1)
            if node.kind == "VERTEX_SAMPLER":
                out["nodes"][nid] = {"kind": node.kind, "ids": flat, "features": self._features(node.vtype or cur_type, flat)}
                continue
            k = node.fanout
            nbr = torch.full((flat.numel(), k), -1, dtype=torch.int64, device=self.device)
            ts = torch.full((flat.numel(), k), -(2 ** 62), dtype=torch.int64, device=self.device)
            w = torch.zeros((flat.numel(), k), device=self.device)
            owner = self.partitioner(flat.clamp(min=0))
            for p, part in enumerate(self.parts):                  # PartitionRouter: frontier vertex -> owner
                st = part.stores[node.etype]
                m = (owner == p) & (flat >= 0) & (flat < st.n)
                if bool(m.any()):
                    a, b, c = st.lookup(flat[m].to(st.device), k)
                    nbr[m], ts[m], w[m] = a.to(self.device), b.to(self.device), c.to(self.device)
            dst_type = self.schema["edges"][node.etype]["dst"]
            rec = {"kind": node.kind, "edge_type": node.etype, "ids": nbr, "timestamps": ts, "weights": w,
                   "features": self._features(dst_type, nbr)}
            out["nodes"][nid] = rec
            out["hops"].append(rec)
            cur_ids[nid] = (nbr, dst_type)
        self.limiter.record((time.perf_counter() - t0) * 1e3)
        self.served += int(src.numel())
        return out

    def checkpoint(self) -> dict:
        return {"parts": [p.checkpoint() for p in self.parts], "ingested": self.ingested}

    def restore(self, ck: dict):
        for p, c in zip(self.parts, ck["parts"]):
            p.restore(c)
        self.ingested = ck["ingested"]

    def stats(self) -> dict:
        return {"ingested": self.ingested, "served": self.served, "partitions": [p.stats() for p in self.parts]}
