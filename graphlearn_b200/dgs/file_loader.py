"""Record ingestion from files (D13: dataloader SDK + ``file_loader`` app).

Pattern file (dynamic_graph_service/dataloader/apps/file_loader/loader.cc:31-50):
    #VERTEX:user,vid,timestamp,feature
    #EDGE:u2i,src,dst,timestamp,weight
data lines start with the vertex / edge type name followed by the fields in pattern order; list
attributes are ``:``-separated.  Records are batched per type (``RecordBatchBuilder``) and handed to
``DynamicGraphService.apply_updates`` - the Kafka hop of the reference (producer -> topic ->
RecordPoller) is an in-process queue here."""
from __future__ import annotations

from typing import Dict, List

import numpy as np


class RecordBatchBuilder(object):
    """Accumulates vertex / edge records and emits the service's batch dict (D2 record batches)."""

    def __init__(self):
        self.edges: Dict[str, Dict[str, list]] = {}
        self.vertices: Dict[str, Dict[str, list]] = {}
        self.size = 0

    def add_edge(self, etype, src, dst, ts, weight=1.0):
        e = self.edges.setdefault(etype, {"src": [], "dst": [], "ts": [], "weight": []})
        e["src"].append(int(src)); e["dst"].append(int(dst)); e["ts"].append(int(ts)); e["weight"].append(float(weight))
        self.size += 1

    def add_vertex(self, vtype, vid, ts, feat):
        v = self.vertices.setdefault(vtype, {"id": [], "ts": [], "feat": []})
        v["id"].append(int(vid)); v["ts"].append(int(ts)); v["feat"].append([float(x) for x in feat])
        self.size += 1

    def finish(self) -> dict:
        out = {"edges": {k: {f: np.asarray(v) for f, v in e.items()} for k, e in self.edges.items()},
               "vertices": {k: {"id": np.asarray(v["id"]), "ts": np.asarray(v["ts"]),
                                "feat": np.asarray(v["feat"], dtype=np.float32)} for k, v in self.vertices.items() if v["id"]}}
        self.__init__()
        return out


class FileLoader(object):
    def __init__(self, pattern_file: str, schema, delimiter: str = ",", list_delimiter: str = ":",
                 batch_size: int = 4096, reverse_edges: Dict[str, str] = None):
        self.schema, self.delim, self.ldelim, self.batch_size = schema, delimiter, list_delimiter, batch_size
        self.patterns: Dict[str, tuple] = {}
        self.reverse_edges = reverse_edges or {}      # etype -> reversed etype fed with (dst, src)
        with open(pattern_file) as f:
            for line in f:
                parts = line.strip().split(delimiter)
                if not parts or not parts[0].startswith("#"):
                    continue
                kind, name = parts[0][1:].split(":")
                if kind not in ("VERTEX", "EDGE"):
                    continue
                self.patterns[name] = (kind, parts[1:])

    def _feat(self, fields: Dict[str, str], attrs) -> List[float]:
        out: List[float] = []
        for a in attrs:
            if a.name in ("timestamp", "weight") or a.name not in fields:
                continue
            v = fields[a.name]
            out.extend(float(x) for x in (v.split(self.ldelim) if a.is_list else [v]) if x != "")
        return out

    def load(self, path: str, service, adaptive: bool = False) -> int:
        """Stream one file into the service; returns the number of records applied.  ``adaptive=True`` scales the
        ingest batch size with the service's ``AdaptiveRateLimiter`` (record-polling concurrency in the reference:
        the limiter shrinks ingest when query latencies miss their P99 target and restores it when stable)."""
        b, n = RecordBatchBuilder(), 0
        base_bs = self.batch_size
        with open(path) as f:
            for line in f:
                parts = line.rstrip("\n").split(self.delim)
                pat = self.patterns.get(parts[0])
                if pat is None or len(parts) - 1 != len(pat[1]):
                    continue
                kind, names = pat
                fields = dict(zip(names, parts[1:]))
                ts = int(fields.get("timestamp", 0))
                if kind == "VERTEX":
                    b.add_vertex(parts[0], fields[names[0]], ts, self._feat(fields, self.schema.vertex_attrs.get(parts[0], [])))
                else:
                    s, d = fields[names[0]], fields[names[1]]
                    w = float(fields.get("weight", 1.0))
                    b.add_edge(parts[0], s, d, ts, w)
                    if parts[0] in self.reverse_edges:
                        b.add_edge(self.reverse_edges[parts[0]], d, s, ts, w)
                n += 1
                if b.size >= self.batch_size:
                    service.apply_updates(b.finish())
                    if adaptive and hasattr(service, "limiter"):
                        c = service.limiter.tick()
                        self.batch_size = max(1, base_bs * c // max(service.limiter.max_c, 1))
        if b.size:
            service.apply_updates(b.finish())
        self.batch_size = base_bs
        return n


class GroupProducer(object):
    """The data-loader SDK's producer (dataloader/include/dataloader/group_producer.h: ``AddVertex`` / ``AddEdge`` /
    ``FlushAll`` batch records per DATA PARTITION and produce one Kafka message per full batch; ``dataloader.h``
    ``SetBarrier``).  Here the sink is a :class:`~graphlearn_b200.dgs.workers.StreamingCluster` (its ingest ``LogChannel``
    is the topic) or a plain service (``apply_updates``): records are grouped by the owner of their key, a partition's
    builder is flushed when it holds ``max_batch_size`` records, and ``set_barrier`` marks "everything produced so far"
    on the coordinator."""

    def __init__(self, sink, max_batch_size: int = 4096, num_partitions: int = None):
        self.sink, self.max_batch = sink, int(max_batch_size)
        self.P = int(num_partitions or getattr(sink, "P", 1))
        self.builders = [RecordBatchBuilder() for _ in range(self.P)]
        self.produced = 0

    def _part(self, key: int) -> int:
        return abs(int(key)) % self.P

    def add_vertex(self, vtype, vid, ts, feat):
        p = self._part(vid)
        self.builders[p].add_vertex(vtype, vid, ts, feat)
        if self.builders[p].size >= self.max_batch:
            self.flush(p)

    def add_edge(self, etype, src, dst, ts, weight=1.0):
        p = self._part(src)
        self.builders[p].add_edge(etype, src, dst, ts, weight)
        if self.builders[p].size >= self.max_batch:
            self.flush(p)

    def flush(self, p: int):
        b = self.builders[p]
        if b.size == 0:
            return
        n = b.size
        batch = b.finish()
        if hasattr(self.sink, "ingest") and hasattr(self.sink, "produced"):      # StreamingCluster: straight into the partition's log
            self.sink.ingest.produce(p, batch, n)
            self.sink.produced += n
        elif hasattr(self.sink, "produce"):
            self.sink.produce(batch)
        else:
            self.sink.apply_updates(batch)
        self.produced += n

    def flush_all(self):
        for p in range(self.P):
            self.flush(p)

    def set_barrier(self, coordinator, name: str):
        """flush, then ask the coordinator to report READY once everything produced so far is sampled and published"""
        self.flush_all()
        coordinator.set_barrier(name)
