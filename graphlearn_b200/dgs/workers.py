"""Decoupled sampling / serving workers of the streaming service (D5 SubscriptionTable, D7 SamplingActor,
D8 ServingActor, D9 channels).

Reference dataflow (dynamic_graph_service/src/service/sampling_actor.act.cc:146-260, serving_actor.act.cc,
src/core/storage/subscription_table.h, src/service/channel/{record_poller,sample_publisher}.cc):

    data loaders --(Kafka dl2spl, partitioned by src vid)--> SAMPLING workers: apply the update to the samplers,
    look the touched (vertex, op) up in the subscription table, publish the new sample to every subscribed SERVING
    worker (Kafka spl2srv) and - because the new neighbour is now reachable from that worker's sources - send a
    subscription rule for (neighbour, downstream op) to the neighbour's owner, which back-fills the neighbour's
    current samples.  A SERVING worker therefore holds a complete k-hop cache of the sources it owns and answers
    ``/infer`` from local state only.

Here the same three roles exist with HBM tables instead of RocksDB and log channels instead of Kafka:

* :class:`LogChannel`         partitioned append-only log with consumer offsets, retention and optional on-disk
                              segments (replay after a restore from the offsets stored in the checkpoint)
* :class:`SubscriptionTable`  one int64 bitmask per (edge-sampler op, vertex): bit w = serving worker w subscribes.
                              Ops fed by SOURCE are subscribed implicitly by the serving partitioner (vid % S).
* :class:`SamplingWorker`     owns the sample stores of the vertices ``vid % P == p``; ``poll()`` drains its ingest
                              partition, applies the batch with the segmented top-k kernel (csrc/dgs.cu), publishes the
                              newest-first rows of touched+subscribed vertices and pushes downstream rules
* :class:`ServingWorker`      per-op row cache ``[V, fanout]`` filled by published rows; ``run_query`` is one row gather
                              per plan node - no selection, no cross-worker traffic
* :class:`StreamingCluster`   wires P sampling x S serving workers, routes rules, exposes produce / pump / query /
                              checkpoint / restore and the barrier test ("everything produced before the barrier has
                              been sampled AND published").
"""
from __future__ import annotations

import os
import threading
from typing import Dict, List, Optional, Sequence

import torch

from .plan import QueryPlan
from .service import DynamicGraphService

_NEG_TS = -(2 ** 62)


# --------------------------------------------------------------------------------------------------- channels
class LogChannel(object):
    """A topic: ``partitions`` append-only logs of record batches.  Offsets count batches; ``poll`` never removes,
    ``truncate`` applies retention.  With ``path`` every batch is also written as ``<path>/p<k>/<offset>.pt`` so a
    restarted consumer can replay from a checkpointed offset."""

    def __init__(self, name: str, partitions: int, path: Optional[str] = None, max_batches: int = 1 << 16):
        self.name, self.P, self.path, self.max_batches = name, int(partitions), path, int(max_batches)
        self._logs: List[List[dict]] = [[] for _ in range(self.P)]
        self._base = [0] * self.P                 # offset of _logs[p][0]
        self._records = [0] * self.P              # records ever produced per partition
        self._cv = threading.Condition()
        if path:
            for p in range(self.P):
                os.makedirs(os.path.join(path, "p%d" % p), exist_ok=True)
            self._recover()

    def _recover(self):
        for p in range(self.P):
            d = os.path.join(self.path, "p%d" % p)
            offs = sorted(int(f[:-3]) for f in os.listdir(d) if f.endswith(".pt"))
            if os.path.exists(os.path.join(d, "BASE")):               # retention point survives an empty log
                with open(os.path.join(d, "BASE")) as f:
                    self._base[p] = int(f.read().strip() or 0)
            if offs:
                self._base[p] = offs[0]
                self._logs[p] = [self._load_segment(os.path.join(d, "%d.pt" % o)) for o in offs]
                self._records[p] = sum(int(b.get("_n", 0)) for b in self._logs[p])

    @staticmethod
    def _load_segment(path: str) -> dict:
        with open(path, "rb") as f:
            head = f.read(6)
        if head == b"GLBR1\x00":
            from .file_loader import decode_record_batch
            with open(path, "rb") as f:
                return decode_record_batch(f.read())
        return torch.load(path, weights_only=True)

    def end_offset(self, p: int) -> int:
        return self._base[p] + len(self._logs[p])

    def produce(self, p: int, batch: dict, n_records: int = 0) -> int:
        with self._cv:
            if len(self._logs[p]) >= self.max_batches:
                raise BufferError("channel %s[%d] is full (%d batches): consumer too slow" % (self.name, p, self.max_batches))
            off = self.end_offset(p)
            batch = dict(batch, _n=int(n_records))
            self._logs[p].append(batch)
            self._records[p] += int(n_records)
            if self.path:
                tmp = os.path.join(self.path, "p%d" % p, "%d.tmp" % off)
                if "edges" in batch or "vertices" in batch:          # record batches: columnar wire format (file_loader.py)
                    from .file_loader import encode_record_batch
                    with open(tmp, "wb") as f:
                        f.write(encode_record_batch(batch))
                else:                                                # published sample rows (device tensors)
                    torch.save(batch, tmp)
                os.replace(tmp, os.path.join(self.path, "p%d" % p, "%d.pt" % off))
            self._cv.notify_all()
            return off

    def poll(self, p: int, offset: int, max_batches: int = 64, timeout: Optional[float] = None) -> List[dict]:
        with self._cv:
            if timeout and offset >= self.end_offset(p):
                self._cv.wait(timeout)
            if offset < self._base[p]:
                raise LookupError("offset %d of %s[%d] fell out of retention (base %d)" % (offset, self.name, p, self._base[p]))
            lo = offset - self._base[p]
            return self._logs[p][lo:lo + max_batches]

    def truncate(self, p: int, before_offset: int):
        with self._cv:
            drop = max(0, min(before_offset - self._base[p], len(self._logs[p])))
            if self.path:
                for o in range(self._base[p], self._base[p] + drop):
                    f = os.path.join(self.path, "p%d" % p, "%d.pt" % o)
                    if os.path.exists(f):
                        os.remove(f)
            del self._logs[p][:drop]
            self._base[p] += drop
            if self.path:
                with open(os.path.join(self.path, "p%d" % p, "BASE"), "w") as f:
                    f.write(str(self._base[p]))


# --------------------------------------------------------------------------------------------------- subscriptions
class SubscriptionTable(object):
    """(op, vertex) -> bitmask of subscribed serving workers (<= 63 workers).  Device resident, grows with the ids."""

    def __init__(self, plan: QueryPlan, num_serving: int, device, n_hint: int = 1024):
        assert num_serving <= 63
        self.S, self.device = int(num_serving), torch.device(device)
        self.root_ops = {n.id for n in plan.nodes.values() if n.kind != "SOURCE" and n.parent == 0}
        self.mask: Dict[int, torch.Tensor] = {n.id: torch.zeros(n_hint, dtype=torch.int64, device=self.device)
                                              for n in plan.nodes.values() if n.kind != "SOURCE" and n.id not in self.root_ops}

    def _ensure(self, op: int, n: int):
        m = self.mask[op]
        if n > m.numel():
            self.mask[op] = torch.cat([m, torch.zeros(max(n, 2 * m.numel()) - m.numel(), dtype=torch.int64, device=self.device)])

    def subscribers(self, op: int, vids: torch.Tensor) -> torch.Tensor:
        """bitmask per vertex"""
        if op in self.root_ops:                                   # SOURCE-fed ops: the serving partitioner decides
            return torch.ones_like(vids) << (vids % self.S)
        m = self.mask[op]
        ok = vids < m.numel()
        return torch.where(ok, m[torch.where(ok, vids, torch.zeros_like(vids))], torch.zeros_like(vids))

    def subscribe(self, op: int, vids: torch.Tensor, worker: int) -> torch.Tensor:
        """returns the vids that were NOT subscribed by ``worker`` before (the ones to back-fill)"""
        if op in self.root_ops or vids.numel() == 0:
            return vids[:0]
        vids = torch.unique(vids[vids >= 0])
        if vids.numel() == 0:
            return vids
        self._ensure(op, int(vids.max().item()) + 1)
        bit = 1 << int(worker)
        m = self.mask[op]
        new = (m[vids] & bit) == 0
        fresh = vids[new]
        m[fresh] |= bit
        return fresh

    def state_dict(self):
        return {str(k): v.clone() for k, v in self.mask.items()}

    def load_state_dict(self, sd):
        for k, v in sd.items():
            self.mask[int(k)] = v.to(self.device).clone()


# --------------------------------------------------------------------------------------------------- workers
class SamplingWorker(object):
    def __init__(self, wid: int, cluster: "StreamingCluster", device=None):
        self.wid, self.c = wid, cluster
        self.svc = DynamicGraphService(cluster.schema, device=device)
        self.device = self.svc.device
        self.offset = 0                    # next ingest batch to consume
        self.subs: Optional[SubscriptionTable] = None
        self.applied_records = 0
        self.published_rows = 0

    def install(self, plan: QueryPlan):
        self.svc.install_query(0, plan)
        self.plan = plan
        self.subs = SubscriptionTable(plan, self.c.S, self.device)
        self.ops_of_etype: Dict[str, List[int]] = {}
        for n in plan.edge_nodes():
            self.ops_of_etype.setdefault(n.etype, []).append(n.id)
        self.vops_of_vtype: Dict[str, List[int]] = {}
        for n in plan.nodes.values():
            if n.kind == "VERTEX_SAMPLER":
                self.vops_of_vtype.setdefault(n.vtype or self._type_at(n.parent), []).append(n.id)

    def _type_at(self, nid: int) -> str:
        n = self.plan.nodes[nid]
        if n.kind == "SOURCE":
            return self.plan.source_type
        return self.c.schema["edges"][n.etype]["dst"]

    # ---- ingest
    def poll(self, max_batches: int = 64) -> int:
        got = self.c.ingest.poll(self.wid, self.offset, max_batches)
        for b in got:
            self._apply(b)
        self.offset += len(got)
        return len(got)

    def _apply(self, batch: dict):
        self.svc.apply_updates(batch)
        self.applied_records += int(batch.get("_n", 0))
        for etype, rec in batch.get("edges", {}).items():
            touched = torch.unique(torch.as_tensor(rec["src"]).to(self.device))
            touched = touched[touched >= 0]
            for op in self.ops_of_etype.get(etype, []):
                self._publish_edges(op, touched, self.subs.subscribers(op, touched))
        for vt, rec in batch.get("vertices", {}).items():
            touched = torch.unique(torch.as_tensor(rec["id"]).to(self.device))
            touched = touched[touched >= 0]
            for op in self.vops_of_vtype.get(vt, []):
                self._publish_vertices(op, vt, touched, self.subs.subscribers(op, touched))

    # ---- publish (SamplePublisher): rows go to the serving workers whose bit is set
    def _publish_edges(self, op: int, vids: torch.Tensor, mask: torch.Tensor):
        if vids.numel() == 0 or not bool((mask != 0).any()):
            return
        node = self.plan.nodes[op]
        st = self.svc.stores[node.etype]
        for w in range(self.c.S):
            sel = vids[(mask >> w) & 1 == 1]
            if sel.numel() == 0:
                continue
            nbr, ts, wt = st.lookup(sel, node.fanout)             # newest-first rows: exactly what a query returns
            self.c.publish.produce(w, {"op": op, "kind": "E", "vid": sel, "nbr": nbr, "ts": ts, "w": wt}, int(sel.numel()))
            self.published_rows += int(sel.numel())
            self._downstream(op, nbr, w)

    def _publish_vertices(self, op: int, vtype: str, vids: torch.Tensor, mask: torch.Tensor):
        vs = self.svc.vstores[vtype]
        if vs.feat is None or vids.numel() == 0 or not bool((mask != 0).any()):
            return
        for w in range(self.c.S):
            sel = vids[((mask >> w) & 1 == 1) & (vids < vs.n)]
            if sel.numel() == 0:
                continue
            self.c.publish.produce(w, {"op": op, "kind": "V", "vid": sel, "feat": vs.feat[sel], "fts": vs.feat_ts[sel]},
                                   int(sel.numel()))
            self.published_rows += int(sel.numel())

    def _downstream(self, op: int, nbr: torch.Tensor, worker: int):
        """CollectDownstreamSubsRules: the neighbours just published to ``worker`` must be tracked for every child op"""
        kids = self.plan.nodes[op].children
        if not kids:
            return
        ids = torch.unique(nbr[nbr >= 0])
        if ids.numel():
            for kid in kids:
                self.c.route_rules(kid, ids, worker)

    # ---- rules (UpdateSubsRules): subscribe + back-fill the current samples of newly subscribed vertices
    def update_rules(self, op: int, vids: torch.Tensor, worker: int):
        vids = vids.to(self.device)
        fresh = self.subs.subscribe(op, vids, worker)
        if fresh.numel() == 0:
            return
        node = self.plan.nodes[op]
        bit = torch.full_like(fresh, 1 << worker)
        if node.kind == "EDGE_SAMPLER":
            st = self.svc.stores[node.etype]
            fresh = fresh[fresh < st.n]
            self._publish_edges(op, fresh, bit[:fresh.numel()])
        else:
            self._publish_vertices(op, node.vtype or self._type_at(node.parent), fresh, bit)

    def checkpoint(self) -> dict:
        return {"svc": self.svc.checkpoint(), "subs": self.subs.state_dict(), "offset": self.offset,
                "applied": self.applied_records}

    def restore(self, ck: dict):
        self.svc.restore(ck["svc"])
        self.subs.load_state_dict(ck["subs"])
        self.offset, self.applied_records = int(ck["offset"]), int(ck["applied"])


class ServingWorker(object):
    """Holds, per plan node, the rows published for the vertices reachable from the sources it owns."""

    def __init__(self, wid: int, cluster: "StreamingCluster", device=None):
        self.wid, self.c = wid, cluster
        self.device = torch.device(device or ("cuda" if torch.cuda.is_available() else "cpu"))
        self.offset = 0
        self.rows: Dict[int, Dict[str, torch.Tensor]] = {}
        self.served = 0

    def install(self, plan: QueryPlan):
        self.plan = plan
        for n in plan.nodes.values():
            if n.kind == "EDGE_SAMPLER":
                self.rows[n.id] = self._alloc_e(1024, n.fanout)
            elif n.kind == "VERTEX_SAMPLER":
                vt = n.vtype or self.c.type_at(plan, n.parent)
                self.rows[n.id] = self._alloc_v(1024, int(self.c.schema["vertices"][vt].get("feat_dim", 0)))

    def _alloc_e(self, n, k):
        d = self.device
        return {"nbr": torch.full((n, k), -1, dtype=torch.int64, device=d), "ts": torch.full((n, k), _NEG_TS, dtype=torch.int64, device=d),
                "w": torch.zeros((n, k), device=d)}

    def _alloc_v(self, n, dim):
        return {"feat": torch.zeros((n, dim), device=self.device), "fts": torch.full((n,), _NEG_TS, dtype=torch.int64, device=self.device)}

    def _ensure(self, op: int, n: int):
        t = self.rows[op]
        cur = next(iter(t.values())).size(0)
        if n <= cur:
            return
        new_n = max(n, 2 * cur)
        fresh = self._alloc_e(new_n, t["nbr"].size(1)) if "nbr" in t else self._alloc_v(new_n, t["feat"].size(1))
        for k, v in t.items():
            fresh[k][:cur] = v
        self.rows[op] = fresh

    def poll(self, max_batches: int = 256) -> int:
        got = self.c.publish.poll(self.wid, self.offset, max_batches)
        for b in got:
            vid = b["vid"].to(self.device)
            self._ensure(b["op"], int(vid.max().item()) + 1)
            t = self.rows[b["op"]]
            if b["kind"] == "E":
                t["nbr"][vid], t["ts"][vid], t["w"][vid] = b["nbr"].to(self.device), b["ts"].to(self.device), b["w"].to(self.device)
            else:
                newer = b["fts"].to(self.device) >= t["fts"][vid]
                t["feat"][vid[newer]] = b["feat"].to(self.device)[newer]
                t["fts"][vid[newer]] = b["fts"].to(self.device)[newer]
        self.offset += len(got)
        return len(got)

    def _gather(self, t: torch.Tensor, ids: torch.Tensor, fill):
        ok = (ids >= 0) & (ids < t.size(0))
        out = t[torch.where(ok, ids, torch.zeros_like(ids))]
        return torch.where(ok.view((-1,) + (1,) * (out.dim() - 1)), out, torch.full_like(out, fill))

    def run_query(self, vids) -> dict:
        """same answer layout as DynamicGraphService.run_query, from local rows only"""
        plan = self.plan
        src = torch.as_tensor(list(vids) if not isinstance(vids, torch.Tensor) else vids, dtype=torch.int64).to(self.device)
        out = {"src": src, "hops": [], "nodes": {}}
        cur = {0: src}
        for nid in plan.topo_order():
            n = plan.nodes[nid]
            if n.kind == "SOURCE":
                continue
            flat = cur[n.parent].reshape(-1)
            t = self.rows[nid]
            if n.kind == "VERTEX_SAMPLER":
                out["nodes"][nid] = {"kind": n.kind, "ids": flat, "features": self._gather(t["feat"], flat, 0.)}
                continue
            rec = {"kind": n.kind, "edge_type": n.etype, "ids": self._gather(t["nbr"], flat, -1),
                   "timestamps": self._gather(t["ts"], flat, _NEG_TS), "weights": self._gather(t["w"], flat, 0.), "features": None}
            out["nodes"][nid] = rec
            out["hops"].append(rec)
            cur[nid] = rec["ids"]
        self.served += int(src.numel())
        return out

    def checkpoint(self) -> dict:
        return {"rows": {str(op): {k: v.clone() for k, v in t.items()} for op, t in self.rows.items()}, "offset": self.offset}

    def restore(self, ck: dict):
        for op, t in ck["rows"].items():
            self.rows[int(op)] = {k: v.to(self.device).clone() for k, v in t.items()}
        self.offset = int(ck["offset"])


# --------------------------------------------------------------------------------------------------- cluster
class StreamingCluster(object):
    """P sampling workers x S serving workers on one box (one device each, or logical partitions of one device)."""

    def __init__(self, schema: Dict[str, dict], num_sampling: int = 2, num_serving: int = 2, sampling_devices: Optional[Sequence] = None,
                 serving_devices: Optional[Sequence] = None, log_dir: Optional[str] = None):
        self.schema, self.P, self.S = schema, int(num_sampling), int(num_serving)
        self.ingest = LogChannel("dl2spl", self.P, os.path.join(log_dir, "dl2spl") if log_dir else None)
        self.publish = LogChannel("spl2srv", self.S, os.path.join(log_dir, "spl2srv") if log_dir else None)
        sd = list(sampling_devices) if sampling_devices else [None] * self.P
        vd = list(serving_devices) if serving_devices else [None] * self.S
        self.sampling = [SamplingWorker(p, self, sd[p]) for p in range(self.P)]
        self.serving = [ServingWorker(s, self, vd[s]) for s in range(self.S)]
        self.plan: Optional[QueryPlan] = None
        self.produced = 0

    def type_at(self, plan: QueryPlan, nid: int) -> str:
        n = plan.nodes[nid]
        return plan.source_type if n.kind == "SOURCE" else self.schema["edges"][n.etype]["dst"]

    def install_query(self, plan: QueryPlan):
        self.plan = plan
        for w in self.sampling:
            w.install(plan)
        for w in self.serving:
            w.install(plan)

    # ---- data loader side: split a record batch by the owner of its key and append to the ingest partitions
    def produce(self, batch: dict) -> int:
        n_total = 0
        parts: List[dict] = [{} for _ in range(self.P)]
        cnt = [0] * self.P
        for kind, key in (("edges", "src"), ("vertices", "id")):
            for name, rec in batch.get(kind, {}).items():
                k = torch.as_tensor(rec[key])
                owner = k.abs() % self.P
                n_total += int(k.numel())
                for p in range(self.P):
                    m = owner == p
                    c = int(m.sum())
                    if c:
                        parts[p].setdefault(kind, {})[name] = {f: torch.as_tensor(v)[m] for f, v in rec.items() if v is not None}
                        cnt[p] += c
        for p in range(self.P):
            if cnt[p]:
                self.ingest.produce(p, parts[p], cnt[p])
        self.produced += n_total
        return n_total

    def route_rules(self, op: int, vids: torch.Tensor, worker: int):
        owner = vids.abs() % self.P
        for p in range(self.P):
            sel = vids[owner == p]
            if sel.numel():
                self.sampling[p].update_rules(op, sel, worker)

    def pump(self, max_rounds: int = 1 << 20) -> int:
        """drive every worker until all channels are drained; returns the number of batches moved"""
        moved = 0
        for _ in range(max_rounds):
            n = sum(w.poll() for w in self.sampling) + sum(w.poll() for w in self.serving)
            moved += n
            if n == 0:
                break
        return moved

    def run_query(self, vids) -> dict:
        """route every source to the serving worker that owns it and stitch the answers in request order"""
        src = torch.as_tensor(list(vids) if not isinstance(vids, torch.Tensor) else vids, dtype=torch.int64)
        if self.S == 1:
            return self.serving[0].run_query(src)
        owner = src.abs() % self.S
        outs, idxs = [], []
        for s in range(self.S):
            idx = (owner == s).nonzero().flatten()
            if idx.numel():
                outs.append(self.serving[s].run_query(src[idx]))
                idxs.append(idx)
        return _stitch(self.plan, src, outs, idxs, self.serving[0].device)

    def barrier_ready(self) -> bool:
        return (all(w.offset >= self.ingest.end_offset(w.wid) for w in self.sampling) and
                all(w.offset >= self.publish.end_offset(w.wid) for w in self.serving))

    def checkpoint(self) -> dict:
        return {"sampling": [w.checkpoint() for w in self.sampling], "serving": [w.checkpoint() for w in self.serving],
                "produced": self.produced}

    def restore(self, ck: dict):
        for w, c in zip(self.sampling, ck["sampling"]):
            w.restore(c)
        for w, c in zip(self.serving, ck["serving"]):
            w.restore(c)
        self.produced = int(ck["produced"])

    def stats(self) -> dict:
        return {"produced": self.produced,
                "sampling": [{"applied": w.applied_records, "published_rows": w.published_rows, "offset": w.offset} for w in self.sampling],
                "serving": [{"served": w.served, "offset": w.offset} for w in self.serving]}


def _stitch(plan: QueryPlan, src, outs, idxs, device):
    B = int(src.numel())
    res = {"src": src.to(device), "hops": [], "nodes": {}}
    mult = {0: 1}
    for nid in plan.topo_order():
        n = plan.nodes[nid]
        if n.kind == "SOURCE":
            continue
        rows_per_src = mult[n.parent]
        rec = None
        for o, idx in zip(outs, idxs):
            part = o["nodes"][nid]
            if rec is None:
                rec = {k: (torch.zeros((B * rows_per_src,) + tuple(v.shape[1:]), dtype=v.dtype, device=device)
                           if isinstance(v, torch.Tensor) else v) for k, v in part.items()}
            pos = (idx.to(device)[:, None] * rows_per_src + torch.arange(rows_per_src, device=device)[None, :]).reshape(-1)
            for k, v in part.items():
                if isinstance(v, torch.Tensor):
                    rec[k][pos] = v.to(device)
        res["nodes"][nid] = rec
        if n.kind == "EDGE_SAMPLER":
            res["hops"].append(rec)
            mult[nid] = rows_per_src * n.fanout
    return res
