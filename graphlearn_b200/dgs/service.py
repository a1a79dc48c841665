from __future__ import annotations

import time
from typing import Dict, List, Optional, Sequence

import torch


class SampleStore(object):
    """Fixed-capacity, timestamp-ordered neighbour samples per source vertex, in device memory."""

    def __init__(self, num_vertices: int, capacity: int, device, feat_dim: int = 0):
        self.n, self.K = int(num_vertices), int(capacity)
        self.device = torch.device(device)
        device = self.device
        self.nbr = torch.full((self.n, self.K), -1, dtype=torch.int64, device=device)
        self.ts = torch.full((self.n, self.K), -(2 ** 62), dtype=torch.int64, device=device)
        self.w = torch.zeros((self.n, self.K), dtype=torch.float32, device=device)
        self.count = torch.zeros(self.n, dtype=torch.int64, device=device)
        self.feat = torch.zeros((self.n, feat_dim), dtype=torch.float32, device=device) if feat_dim else None
        self.feat_ts = torch.full((self.n,), -(2 ** 62), dtype=torch.int64, device=device) if feat_dim else None

    def ensure(self, n: int):
        """grow the tables (doubling) so that vertex ids < n are addressable - the reference's KV store
        has no fixed vertex universe, the HBM tables emulate that by amortised growth"""
        if n <= self.n:
            return
        new_n = max(n, 2 * self.n)
        grow = new_n - self.n

        def ext(t, fill):
            pad = torch.full((grow,) + tuple(t.shape[1:]), fill, dtype=t.dtype, device=t.device)
            return torch.cat([t, pad])
        self.nbr, self.ts, self.w = ext(self.nbr, -1), ext(self.ts, -(2 ** 62)), ext(self.w, 0)
        self.count = ext(self.count, 0)
        if self.feat is not None:
            self.feat, self.feat_ts = ext(self.feat, 0), ext(self.feat_ts, -(2 ** 62))
        self.n = new_n

    MAX_VERTEX_ID = 1 << 31        # growth cap of the HBM tables: ids come from the network (HTTP ingest)

    def _valid_rows(self, key, *cols):
        """drop records whose key id is negative (would wrap to another vertex's row) or absurdly large (would make
        ``ensure`` allocate without bound)"""
        if key.numel() == 0:
            return (key,) + cols
        ok = (key >= 0) & (key < self.MAX_VERTEX_ID)
        if bool(ok.all()):
            return (key,) + cols
        return (key[ok],) + tuple(None if c is None else c.to(key.device)[ok] for c in cols)

    def apply_edges(self, src: torch.Tensor, dst: torch.Tensor, ts: torch.Tensor, w: Optional[torch.Tensor] = None):
        """TopK-by-timestamp: an incoming edge replaces the OLDEST kept sample of its source if it is
        newer.  Processed in timestamp order; duplicates of one source inside a batch are resolved in
        rounds (each round applies at most one update per source)."""
        src, dst, ts = src.to(self.device), dst.to(self.device), ts.to(self.device)
        src, dst, ts, w = self._valid_rows(src, dst, ts, w)
        if src.numel():
            self.ensure(int(src.max().item()) + 1)
        w = torch.ones_like(ts, dtype=torch.float32) if w is None else w.to(self.device).float()
        order = torch.argsort(ts, stable=True)
        src, dst, ts, w = src[order], dst[order], ts[order], w[order]
        if self.device.type == "cuda" and self.K <= 64:
            # one launch per record batch (csrc/dgs.cu): sort by (src, ts), the thread on the last record of every
            # source folds that source's <= K newest records into its row
            from ..parallel.runtime import native
            o2 = torch.argsort(src, stable=True)
            native().dgs_apply_edges(self.nbr, self.ts, self.w, self.count, src[o2].contiguous(), dst[o2].contiguous(),
                                     ts[o2].contiguous(), w[o2].contiguous())
            return
        pending = torch.ones_like(src, dtype=torch.bool)
        while bool(pending.any()):
            idx = pending.nonzero().flatten()
            s = src[idx]
            # first pending occurrence of every source in this round
            uniq, inv = torch.unique(s, return_inverse=True)
            first = torch.full((uniq.numel(),), idx.numel(), dtype=torch.int64, device=self.device)
            first.scatter_reduce_(0, inv, torch.arange(idx.numel(), device=self.device), "amin")
            sel = idx[first]
            vs, vd, vt, vw = src[sel], dst[sel], ts[sel], w[sel]
            slot = torch.argmin(self.ts[vs], dim=1)                    # oldest (or empty) slot
            newer = vt > self.ts[vs, slot]
            vs2, sl2 = vs[newer], slot[newer]
            was_empty = self.nbr[vs2, sl2] < 0
            self.nbr[vs2, sl2] = vd[newer]
            self.ts[vs2, sl2] = vt[newer]
            self.w[vs2, sl2] = vw[newer]
            self.count[vs2] += was_empty.to(torch.int64)
            pending[sel] = False

    def apply_vertices(self, vid: torch.Tensor, ts: torch.Tensor, feat: torch.Tensor):
        """latest-version vertex sampler: keep the feature row with the largest timestamp."""
        if self.feat is None:
            return
        vid, ts, feat = vid.to(self.device), ts.to(self.device), feat.to(self.device).float()
        vid, ts, feat = self._valid_rows(vid, ts, feat)
        if vid.numel():
            self.ensure(int(vid.max().item()) + 1)
        order = torch.argsort(ts, stable=True)
        vid, ts, feat = vid[order], ts[order], feat[order]
        newer = ts >= self.feat_ts[vid]
        # later rows of the (sorted) batch win
        self.feat[vid[newer]] = feat[newer]
        self.feat_ts.scatter_reduce_(0, vid[newer], ts[newer], "amax")

    def expire(self, before_ts: int) -> int:
        """TTL: drop every kept sample older than ``before_ts`` (the reference stores edge samples in a RocksDB
        ``DBWithTTL``, sample_store.h:71-170, option ``sample-store.ttl-hours``).  Returns the number dropped."""
        old = (self.ts < int(before_ts)) & (self.nbr >= 0)
        n = int(old.sum().item())
        if n:
            self.nbr[old] = -1
            self.ts[old] = -(2 ** 62)
            self.w[old] = 0
            self.count -= old.sum(1)
        return n

    def lookup(self, vids: torch.Tensor, k: int):
        """most recent k samples of each vertex: (nbr [B,k], ts [B,k], w [B,k]); -1 padded."""
        if self.device.type == "cuda" and self.K <= 64:
            from ..parallel.runtime import native
            n_, t_, w_ = native().dgs_lookup(self.nbr, self.ts, self.w, vids.to(self.device).to(torch.int64), int(k))
            return n_, t_, w_
        vids = torch.where(vids < self.n, vids, torch.zeros_like(vids)) if vids.numel() else vids
        t, order = torch.sort(self.ts[vids], dim=1, descending=True)
        order = order[:, :k]
        return (torch.gather(self.nbr[vids], 1, order), t[:, :k], torch.gather(self.w[vids], 1, order))

    # ---- key-level access (the reference's KV surface: sample_store.h GetEdgesByPrefix / GetVertex / Delete*ByPrefix)
    def get(self, vid: int):
        """kept samples of ONE vertex, most recent first: (nbr [c], ts [c], w [c]) with c = number of kept samples."""
        if not 0 <= int(vid) < self.n:
            e = torch.zeros(0, dtype=torch.int64, device=self.device)
            return e, e.clone(), torch.zeros(0, device=self.device)
        ok = self.nbr[vid] >= 0
        t, order = torch.sort(self.ts[vid][ok], descending=True)
        return self.nbr[vid][ok][order], t, self.w[vid][ok][order]

    def get_vertex(self, vid: int):
        """(feature row, its timestamp) of one vertex, or None when nothing was stored for it."""
        if self.feat is None or not 0 <= int(vid) < self.n or int(self.feat_ts[vid]) == -(2 ** 62):
            return None
        return self.feat[vid].clone(), int(self.feat_ts[vid])

    def delete(self, vids: torch.Tensor) -> int:
        """drop every kept sample (and the stored feature version) of the given vertices; returns the number of samples
        dropped (DeleteEdgesByPrefix / DeleteVerticesByPrefix: a vertex left the subscribed set or was deleted upstream)."""
        vids = torch.as_tensor(vids, dtype=torch.int64).to(self.device).reshape(-1)
        vids = torch.unique(vids[(vids >= 0) & (vids < self.n)])
        if vids.numel() == 0:
            return 0
        dropped = int((self.nbr[vids] >= 0).sum().item())
        self.nbr[vids] = -1
        self.ts[vids] = -(2 ** 62)
        self.w[vids] = 0
        self.count[vids] = 0
        if self.feat is not None:
            self.feat[vids] = 0
            self.feat_ts[vids] = -(2 ** 62)
        return dropped

    def state_dict(self):
        return {k: getattr(self, k).clone() for k in ("nbr", "ts", "w", "count") } | \
               ({"feat": self.feat.clone(), "feat_ts": self.feat_ts.clone()} if self.feat is not None else {})

    def load_state_dict(self, sd):
        self.ensure(int(sd["nbr"].size(0)))
        # rows the tables grew by since the snapshot must not survive a restore
        self.nbr.fill_(-1); self.ts.fill_(-(2 ** 62)); self.w.zero_(); self.count.zero_()
        if self.feat is not None:
            self.feat.zero_(); self.feat_ts.fill_(-(2 ** 62))
        for k, v in sd.items():
            getattr(self, k)[:v.size(0)].copy_(v)


from .plan import PlanNode, QueryPlan  # noqa: E402,F401


class AdaptiveRateLimiter(object):
    """If fewer than 99 % of recent query latencies meet the P99 target, divide the ingest
    concurrency by 3; after a stable window multiply it by 1.3 (adaptive_rate_limiter.cc:52-87)."""

    def __init__(self, target_ms: float = 20.0, max_concurrency: int = 64, stable_windows: int = 3):
        self.target_ms, self.max_c = target_ms, max_concurrency
        self.concurrency = max_concurrency
        self._lat: List[float] = []
        self._stable = 0
        self._need = stable_windows

    def record(self, latency_ms: float):
        self._lat.append(latency_ms)

    def tick(self) -> int:
        if self._lat:
            ok = sum(1 for x in self._lat if x <= self.target_ms) / len(self._lat)
            if ok < 0.99:
                self.concurrency = max(1, self.concurrency // 3)
                self._stable = 0
            else:
                self._stable += 1
                if self._stable >= self._need:
                    self.concurrency = min(self.max_c, max(self.concurrency + 1, int(self.concurrency * 1.3)))
                    self._stable = 0
        self._lat = []
        return self.concurrency


class DynamicGraphService(object):
    def __init__(self, schema: Dict[str, dict], device=None):
        """schema: {"vertices": {type: {"count": n, "feat_dim": d}}, "edges": {etype: {"src": t, "dst": t}}}"""
        self.device = torch.device(device or ("cuda" if torch.cuda.is_available() else "cpu"))
        self.schema = schema
        self.stores: Dict[str, SampleStore] = {}
        self.vstores: Dict[str, SampleStore] = {}
        self.queries: Dict[int, QueryPlan] = {}
        self.limiter = AdaptiveRateLimiter()
        self.ingested = 0
        self.served = 0
        for vt, info in schema["vertices"].items():
            self.vstores[vt] = SampleStore(info["count"], 1, self.device, feat_dim=info.get("feat_dim", 0))

    def install_query(self, qid: int, plan: QueryPlan):
        """Allocates sampler state for every (edge type, capacity) the plan needs - like the
        reference, only what a query subscribes to is kept."""
        for etype, k in plan.hops:
            e = self.schema["edges"][etype]
            n = self.schema["vertices"][e["src"]]["count"]
            cur = self.stores.get(etype)
            if cur is None or cur.K < k:
                # the existing store may have grown past the schema's vertex count (ensure()): size the wider one to
                # whatever is larger and back-fill only the rows / slots that exist
                st = SampleStore(max(n, cur.n) if cur is not None else n, k, self.device)
                if cur is not None:                       # back-fill existing samples
                    st.nbr[:cur.n, :cur.K], st.ts[:cur.n, :cur.K], st.w[:cur.n, :cur.K] = cur.nbr, cur.ts, cur.w
                    st.count[:cur.n].copy_(cur.count)
                self.stores[etype] = st
        self.queries[qid] = plan

    def apply_updates(self, batch: dict):
        """batch: {"edges": {etype: {"src","dst","ts"[,"weight"]}}, "vertices": {vtype: {"id","ts","feat"}}}"""
        for etype, rec in batch.get("edges", {}).items():
            if etype in self.stores:
                self.stores[etype].apply_edges(torch.as_tensor(rec["src"]), torch.as_tensor(rec["dst"]),
                                               torch.as_tensor(rec["ts"]),
                                               None if rec.get("weight") is None else torch.as_tensor(rec["weight"]))
                self.ingested += len(rec["src"])
        for vt, rec in batch.get("vertices", {}).items():
            self.vstores[vt].apply_vertices(torch.as_tensor(rec["id"]), torch.as_tensor(rec["ts"]),
                                            torch.as_tensor(rec["feat"]))

    def run_query(self, qid: int, vids: Sequence[int]) -> dict:
        """Batched inference-time lookup.  Walks the plan tree (QueryExecutor, query_executor.cc:42-125):
        every EDGE_SAMPLER node returns ids [B*prod(k of its ancestors), k] (-1 padded) + timestamps +
        weights and the latest features of the returned vertices; VERTEX_SAMPLER nodes return the
        latest feature rows of their input vertices.  ``out["hops"]`` lists the edge nodes in plan
        order (chain plans: hop i), ``out["nodes"][plan_node_id]`` has every node."""
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
                vs = self.vstores[node.vtype or cur_type]
                ok = (flat >= 0) & (flat < vs.n)
                feat = vs.feat[torch.where(ok, flat, torch.zeros_like(flat))] if vs.feat is not None else None
                if feat is not None:
                    feat = torch.where(ok[:, None], feat, torch.zeros_like(feat))
                out["nodes"][nid] = {"kind": node.kind, "ids": flat, "features": feat}
                continue
            st = self.stores[node.etype]
            nbr, ts, w = st.lookup(flat.clamp(min=0), node.fanout)
            ok = ((flat >= 0) & (flat < st.n))[:, None]
            nbr = torch.where(ok, nbr, torch.full_like(nbr, -1))
            ts = torch.where(ok, ts, torch.full_like(ts, -(2 ** 62)))      # unknown vertices answer like empty ones
            w = torch.where(ok, w, torch.zeros_like(w))
            dst_type = self.schema["edges"][node.etype]["dst"]
            vs = self.vstores[dst_type]
            feat = None
            if vs.feat is not None:
                okv = (nbr >= 0) & (nbr < vs.n)
                feat = vs.feat[torch.where(okv, nbr, torch.zeros_like(nbr))] * okv.unsqueeze(-1)
            rec = {"kind": node.kind, "edge_type": node.etype, "ids": nbr, "timestamps": ts, "weights": w, "features": feat}
            out["nodes"][nid] = rec
            out["hops"].append(rec)
            cur_ids[nid] = (nbr, dst_type)
        self.limiter.record((time.perf_counter() - t0) * 1e3)
        self.served += int(src.numel())
        return out

    def expire(self, before_ts: int) -> int:
        """apply the sample TTL to every edge store (timestamps are whatever unit the records use)"""
        return sum(st.expire(before_ts) for st in self.stores.values())

    def delete_vertices(self, vtype: str, ids) -> int:
        """a vertex was deleted upstream: drop its feature version and the samples of every edge type that starts at it"""
        n = self.vstores[vtype].delete(ids) if vtype in self.vstores else 0
        for et, info in self.schema["edges"].items():
            if info["src"] == vtype and et in self.stores:
                n += self.stores[et].delete(ids)
        return n

    def delete_edges(self, etype: str, src_ids) -> int:
        """drop the kept samples of ``etype`` of the given source vertices"""
        return self.stores[etype].delete(src_ids) if etype in self.stores else 0

    def stats(self) -> dict:
        return {"ingested": self.ingested, "served": self.served, "queries": sorted(self.queries),
                "concurrency": self.limiter.concurrency,
                "stores": {k: {"vertices": v.n, "capacity": v.K, "filled": int((v.count > 0).sum())} for k, v in self.stores.items()}}

    def checkpoint(self) -> dict:
        return {"stores": {k: v.state_dict() for k, v in self.stores.items()},
                "vstores": {k: v.state_dict() for k, v in self.vstores.items()}, "ingested": self.ingested}

    def restore(self, ck: dict):
        for k, v in ck["stores"].items():
            self.stores[k].load_state_dict(v)
        for k, v in ck["vstores"].items():
            self.vstores[k].load_state_dict(v)
        self.ingested = ck["ingested"]
