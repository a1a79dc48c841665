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
                 batch_size: int = 4096, reverse_edges: Dict[str, str] = None, native: bool = True):
        """``native``: parse with the C++ record parser (``csrc/host_loader.cpp parse_records``: columnar batches straight from
        the file, ~two orders of magnitude faster than the line-by-line Python path below, which stays as the oracle and the
        fallback when the extension is not built)."""
        self.schema, self.delim, self.ldelim, self.batch_size = schema, delimiter, list_delimiter, batch_size
        self.native = native
        self.patterns: Dict[str, tuple] = {}
        self.reverse_edges = reverse_edges or {}      # etype -> reversed etype fed with (dst, src)
        with open(pattern_file) as f:
            for line in f:
                parts = line.strip().split(delimiter)
                if not parts or not parts[0].startswith("#"):
                    continue
                head = parts[0][1:].split(":")
                if len(head) != 2 or head[0] not in ("VERTEX", "EDGE"):
                    continue                                  # a comment line
                kind, name = head
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
        if self.native and len(self.delim) == 1 and len(self.ldelim) == 1:
            C = _native_or_none()
            if C is not None:
                return self._load_native(C, path, service, adaptive)
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


    def _specs(self):
        names = list(self.patterns)
        kinds, n_fields, ts_idx, w_idx, feat_off, feat_field, feat_is_list = [], [], [], [], [0], [], []
        for name in names:
            kind, fields = self.patterns[name]
            kinds.append(0 if kind == "VERTEX" else 1)
            n_fields.append(len(fields))
            ts_idx.append(fields.index("timestamp") if "timestamp" in fields else -1)
            w_idx.append(fields.index("weight") if kind == "EDGE" and "weight" in fields else -1)
            if kind == "VERTEX":
                for a in self.schema.vertex_attrs.get(name, []):
                    if a.name in ("timestamp", "weight") or a.name not in fields:
                        continue
                    feat_field.append(fields.index(a.name))
                    feat_is_list.append(1 if a.is_list else 0)
            feat_off.append(len(feat_field))
        return names, kinds, n_fields, ts_idx, w_idx, feat_off, feat_field, feat_is_list

    def _load_native(self, C, path: str, service, adaptive: bool) -> int:
        names, kinds, n_fields, ts_idx, w_idx, feat_off, feat_field, feat_is_list = self._specs()
        base_bs, off, total = self.batch_size, 0, 0
        while True:
            out = C.parse_records(path, off, max(1, self.batch_size), self.delim, self.ldelim, names, kinds, n_fields, ts_idx, w_idx,
                                  feat_off, feat_field, feat_is_list, 8 << 20)
            off, n, _, eof = (int(x) for x in out[-1].tolist())
            if n:
                batch = {"edges": {}, "vertices": {}}
                for t, name in enumerate(names):
                    a, b_, c, d = out[4 * t:4 * t + 4]
                    if a.numel() == 0:
                        continue
                    if kinds[t] == 0:
                        batch["vertices"][name] = {"id": a.numpy(), "ts": b_.numpy(), "feat": c.numpy()}
                    else:
                        batch["edges"].setdefault(name, []).append((a, b_, c, d))
                        if name in self.reverse_edges:
                            batch["edges"].setdefault(self.reverse_edges[name], []).append((b_, a, c, d))
                batch["edges"] = {k: {"src": np.concatenate([p[0].numpy() for p in v]), "dst": np.concatenate([p[1].numpy() for p in v]),
                                      "ts": np.concatenate([p[2].numpy() for p in v]),
                                      "weight": np.concatenate([p[3].numpy() for p in v]).astype(np.float64)}
                                  for k, v in batch["edges"].items()}
                service.apply_updates(batch)
                total += n
                if adaptive and hasattr(service, "limiter"):
                    c_ = service.limiter.tick()
                    self.batch_size = max(1, base_bs * c_ // max(service.limiter.max_c, 1))
            if eof or (n == 0 and off == 0 and eof):
                break
        self.batch_size = base_bs
        return total


def _native_or_none():
    try:
        from ..parallel.runtime import native
        C = native()
        return C if hasattr(C, "parse_records") else None
    except Exception:  # noqa: BLE001  (extension not built: the Python path is complete)
        return None


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


# ---------------------------------------------------------------------------------------------------- D2 wire format
_MAGIC = b"GLBR1\x00"
_DT = {"i8": np.int64, "f4": np.float32, "f8": np.float64, "i4": np.int32}


def encode_record_batch(batch: dict) -> bytes:
    """Record batch -> bytes.  The reference ships FlatBuffers ``RecordBatch`` tables (fbs/record.fbs) through Kafka; the
    batches here are columnar dicts, so the wire form is columnar too: a JSON header (kind / type name / column / dtype /
    shape / byte range) followed by the raw little-endian column buffers - zero parsing on the consumer side
    (``np.frombuffer`` views).  Used for HTTP ingest bodies and for persisted ``LogChannel`` segments."""
    import json
    cols, blobs, off = [], [], 0
    for kind in ("edges", "vertices"):
        for name, rec in batch.get(kind, {}).items():
            for col, arr in rec.items():
                if arr is None:
                    continue
                a = np.ascontiguousarray(arr.detach().cpu().numpy() if hasattr(arr, "detach") else np.asarray(arr))
                if a.dtype.kind == "i" and a.dtype != np.int64:
                    a = a.astype(np.int64)
                if a.dtype.kind == "f" and a.dtype not in (np.float32, np.float64):
                    a = a.astype(np.float32)
                code = {np.dtype(np.int64): "i8", np.dtype(np.float32): "f4", np.dtype(np.float64): "f8"}[a.dtype]
                b = a.tobytes()
                cols.append({"k": kind, "t": name, "c": col, "d": code, "s": list(a.shape), "o": off, "n": len(b)})
                blobs.append(b)
                off += len(b)
    head = json.dumps({"cols": cols, "n": int(batch.get("_n", 0))}).encode()
    return _MAGIC + len(head).to_bytes(4, "little") + head + b"".join(blobs)


def decode_record_batch(buf: bytes) -> dict:
    """inverse of :func:`encode_record_batch`; columns are numpy views of ONE writable copy of ``buf`` (torch wants writable
    memory), pass a ``bytearray`` to decode in place"""
    import json
    if not isinstance(buf, bytearray):
        buf = bytearray(buf)
    if bytes(buf[:len(_MAGIC)]) != _MAGIC:
        raise ValueError("not a graphlearn_b200 record batch")
    p = len(_MAGIC)
    hl = int.from_bytes(buf[p:p + 4], "little")
    head = json.loads(bytes(buf[p + 4:p + 4 + hl]).decode())
    base = p + 4 + hl
    out: dict = {}
    for c in head["cols"]:
        if c["d"] not in _DT or c["k"] not in ("edges", "vertices"):
            raise ValueError("bad column descriptor in record batch")
        a = np.frombuffer(buf, dtype=_DT[c["d"]], count=int(np.prod(c["s"])) if c["s"] else 1, offset=base + c["o"]).reshape(c["s"])
        out.setdefault(c["k"], {}).setdefault(c["t"], {})[c["c"]] = a
    if head.get("n"):
        out["_n"] = int(head["n"])
    return out
