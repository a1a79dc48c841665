"""GSL client of the streaming service (D14).

The reference ships a Java client (dynamic_graph_service/gsl_client: ``Graph.connect(addr)``, ``g.V(vtype).feed(source)
.properties(1).alias("seed").outV("u2i").sample(15).by("topk_by_timestamp").properties(1).alias("hop1").values()``,
``install`` / ``run`` (+ async forms), ``checkBarrier``, ``getSchema``, ``getQuery`` and the ``EgoGraph`` / ``EgoTensor``
converters that turn a query result into per-hop model inputs - Graph.java, Traversal.java, Query.java,
predict/EgoGraph.java, predict/EgoTensor.java).  This module is that client for Python programs: the same fluent
traversal produces the reference's install-query JSON (``plan_nodes`` with ``kind / type / links / params / filter``),
talks to ``dgs/http_server.py`` over HTTP, and decodes results into numpy arrays keyed by alias.

Only the standard library and numpy are needed on the client side (no torch, no GPU).
"""
from __future__ import annotations

import json
import urllib.error
import urllib.parse
import urllib.request
from concurrent.futures import Future, ThreadPoolExecutor
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np

STRATEGIES = {"topk_by_timestamp": 0}      # the streaming samplers keep the k most recent edges per vertex


class UserException(Exception):
    """Misuse of the traversal API or an error answered by the service (exception/UserException.java)."""


class Status(object):
    OK, NOT_READY, ERROR = "OK", "NOT_READY_ERROR", "ERROR"

    def __init__(self, code: str = "OK", message: str = ""):
        self.code, self.message = code, message

    def ok(self) -> bool:
        return self.code == Status.OK

    def __repr__(self):
        return "Status(%s%s)" % (self.code, (": " + self.message) if self.message else "")


class DataSource(object):
    """Feeds seed vertex ids to a query, ``batch`` ids per ``next()`` (DataSource.java)."""

    def __init__(self, vids: Iterable[int], batch: int = 1):
        self._vids = [int(v) for v in vids]
        self._batch, self._pos = max(1, int(batch)), 0

    def has_next(self) -> bool:
        return self._pos < len(self._vids)

    def next(self) -> List[int]:
        if not self.has_next():
            raise StopIteration
        out = self._vids[self._pos:self._pos + self._batch]
        self._pos += len(out)
        return out

    def seek(self, pos: int = 0):
        self._pos = int(pos)


class _Node(object):
    def __init__(self, nid, kind, vtype, etype=None, fanout=0, strategy=0, versions=1, dst_vtype=None):
        self.id, self.kind, self.vtype, self.etype = nid, kind, vtype, etype
        self.fanout, self.strategy, self.versions = fanout, strategy, versions
        self.dst_vtype = dst_vtype          # vertex type reached through an edge sampler
        self.links: List[int] = []
        self.alias: Optional[str] = None

    def to_json(self) -> dict:
        params = [{"key": "vtype", "value": self.vtype}]
        if self.kind == "EDGE_SAMPLER":
            params += [{"key": "etype", "value": self.etype}, {"key": "fanout", "value": self.fanout},
                       {"key": "strategy", "value": self.strategy}]
        else:
            params.append({"key": "versions", "value": self.versions})
        extra = {"alias": self.alias} if self.alias else {}
        return {**extra, "id": self.id, "kind": self.kind, "type": "EDGE" if self.kind == "EDGE_SAMPLER" else "VERTEX",
                "links": [{"node": n, "src_output": 1 if self.kind == "EDGE_SAMPLER" else 0, "dst_input": 0} for n in self.links],
                "params": params,
                "filter": {"weighted": self.kind == "EDGE_SAMPLER", "labeled": False, "attributed": self.kind == "VERTEX_SAMPLER"}}


class Query(object):
    """A finished traversal: the plan (reference JSON), its aliases and the data source (Query.java)."""

    def __init__(self, nodes: List[_Node], source: Optional[DataSource], priority: int = 0):
        self.nodes, self.source, self.priority = nodes, source, priority
        self.id: Optional[int] = None
        self.server_ids: Dict[int, int] = {n.id: n.id for n in nodes}       # plan id -> id used in the service's answers

    def feed(self, source: DataSource):
        self.source = source
        return self

    def aliases(self) -> Dict[str, int]:
        return {n.alias: n.id for n in self.nodes if n.alias}

    def to_json(self) -> dict:
        d = {"priority": self.priority, "query_plan": {"plan_nodes": [n.to_json() for n in self.nodes]}}
        if self.id is not None:
            d["query_id"] = self.id
        return d

    @staticmethod
    def from_json(d: dict) -> "Query":
        nodes = []
        for n in d["query_plan"]["plan_nodes"]:
            p = {x["key"]: x["value"] for x in n.get("params", [])}
            node = _Node(n["id"], n["kind"], p.get("vtype"), p.get("etype"), p.get("fanout", 0), p.get("strategy", 0), p.get("versions", 1))
            node.links = [l["node"] for l in n.get("links", [])]
            node.alias = n.get("alias")
            nodes.append(node)
        q = Query(nodes, None, d.get("priority", 0))
        q.id = d.get("query_id")
        return q


class Traversal(object):
    """Fluent plan builder (Traversal.java): V -> [properties] -> outV -> sample -> by -> [properties] -> ... -> values."""

    def __init__(self, graph: "Graph", nodes: List[_Node], cur: _Node, cur_vtype: int, source: Optional[DataSource] = None):
        # _cur is the POSITION of the traversal: the SOURCE node or the edge sampler that led here
        self._g, self._nodes, self._cur, self._vtype, self._source = graph, nodes, cur, cur_vtype, source

    def feed(self, source: DataSource) -> "Traversal":
        if self._cur.kind != "SOURCE":
            raise UserException("feed() belongs right after V()")
        self._source = source
        return self

    def _new(self, kind, **kw) -> _Node:
        n = _Node(len(self._nodes), kind, **kw)
        self._nodes.append(n)
        return n

    def outV(self, etype: str) -> "Traversal":
        sch = self._g.get_schema()
        if etype not in sch["edge_id"]:
            raise UserException("unknown edge type %r" % etype)
        src_t, dst_t = sch["relations"][etype]
        if sch["vertex_id"][src_t] != self._vtype:
            raise UserException("edge %r starts at %r, the traversal is at vertex type %d" % (etype, src_t, self._vtype))
        n = self._new("EDGE_SAMPLER", vtype=self._vtype, etype=sch["edge_id"][etype], fanout=1, dst_vtype=sch["vertex_id"][dst_t])
        self._cur.links.append(n.id)
        return Traversal(self._g, self._nodes, n, sch["vertex_id"][dst_t], self._source)

    def sample(self, fanout: int) -> "Traversal":
        if self._cur.kind != "EDGE_SAMPLER":
            raise UserException("sample() follows outV()")
        if int(fanout) <= 0:
            raise UserException("fanout must be positive")
        self._cur.fanout = int(fanout)
        return self

    def by(self, strategy: str) -> "Traversal":
        if self._cur.kind != "EDGE_SAMPLER":
            raise UserException("by() follows sample()")
        if strategy not in STRATEGIES:
            raise UserException("unknown strategy %r (known: %s)" % (strategy, ", ".join(STRATEGIES)))
        self._cur.strategy = STRATEGIES[strategy]
        return self

    def properties(self, versions: int = 1, *keys: str) -> "Traversal":
        """Fetch the latest ``versions`` property versions of the vertices at the current position."""
        n = self._new("VERTEX_SAMPLER", vtype=self._vtype, versions=int(versions))
        self._cur.links.append(n.id)
        return self

    def alias(self, name: str) -> "Traversal":
        """Name the vertices of the current position; ``Value[name]`` merges the position's record with its properties."""
        if any(n.alias == name for n in self._nodes if n is not self._cur):
            raise UserException("duplicate alias %r" % name)
        self._cur.alias = name
        return self

    def values(self) -> Query:
        if self._source is None:
            raise UserException("no DataSource: call feed() after V()")
        for n in self._nodes:
            if n.kind == "EDGE_SAMPLER" and n.fanout <= 0:
                raise UserException("edge sampler %d has no fan-out" % n.id)
        return Query(self._nodes, self._source)


class Value(object):
    """Result of one ``run``: per alias the ids (and timestamps / weights / features when the node carries them) as numpy
    arrays (Value.java + ValueBuilder.java decode the FlatBuffers answer; here the answer is JSON)."""

    def __init__(self, query: Query, raw: dict):
        self.raw = raw
        self.src = np.asarray(raw.get("src", []), dtype=np.int64)
        self._by_node: Dict[int, dict] = {}
        for nid, rec in raw.get("nodes", {}).items():
            dec = {}
            for k, v in rec.items():
                if v is None or isinstance(v, str):
                    dec[k] = v
                else:
                    dec[k] = np.asarray(v, dtype=np.float32 if k in ("features", "weights") else np.int64)
            self._by_node[int(nid)] = dec
        self._alias = {a: query.server_ids.get(n, n) for a, n in query.aliases().items()}
        self._query = query

    def node(self, plan_id: int) -> dict:
        return self._by_node[self._query.server_ids.get(plan_id, plan_id)]

    def __getitem__(self, alias: str) -> dict:
        pos = self._query.aliases()[alias]
        by_id = {n.id: n for n in self._query.nodes}
        rec = dict(self._by_node.get(self._alias[alias], {})) if by_id[pos].kind != "SOURCE" else {"ids": self.src}
        for c in by_id[pos].links:                               # the position's property node, when it has one
            if by_id[c].kind == "VERTEX_SAMPLER" and rec.get("features") is None:
                f = self.node(c).get("features")
                if f is not None:
                    rec["features"] = f.reshape(tuple(rec["ids"].shape) + (f.shape[-1],))
        return rec

    def aliases(self) -> List[str]:
        return list(self._alias)

    def ego_graph(self) -> "EgoGraph":
        return EgoGraph(self._query, self)


class EgoGraph(object):
    """Hop-structured view of a chain / tree query result (predict/EgoGraph.java): hop 0 = the seeds, hop i = the vertices
    reached by the i-th edge sampler on the path, with their latest features when the plan asked for properties."""

    def __init__(self, query: Query, value: Value):
        by_id = {n.id: n for n in query.nodes}
        self.vtypes: List[int] = [by_id[0].vtype]
        self.ids: List[np.ndarray] = [value.src.reshape(-1)]
        self.features: List[Optional[np.ndarray]] = [None]
        self.fanouts: List[int] = []
        # features of the seeds: a VERTEX_SAMPLER child of the source
        for c in by_id[0].links:
            if by_id[c].kind == "VERTEX_SAMPLER":
                self.features[0] = value.node(c).get("features")
        cur = by_id[0]
        while True:
            nxt = [by_id[c] for c in cur.links if by_id[c].kind == "EDGE_SAMPLER"]
            if not nxt:
                break
            cur = nxt[0]                                           # the first edge path (trees: query the others by alias)
            rec = value.node(cur.id)
            self.vtypes.append(cur.dst_vtype if cur.dst_vtype is not None else -1)
            self.ids.append(rec["ids"].reshape(-1))
            f = rec.get("features")
            self.features.append(None if f is None else f.reshape(-1, f.shape[-1]))
            self.fanouts.append(cur.fanout)

    def num_hops(self) -> int:
        return len(self.ids) - 1

    def get_vids(self, hop: int) -> np.ndarray:
        return self.ids[hop]

    def get_vtype(self, hop: int) -> int:
        return self.vtypes[hop]

    def hop_tensors(self, feat_dim: Optional[int] = None) -> List[np.ndarray]:
        """Model inputs (predict/EgoTensor.java): one dense [n_hop, d] float matrix per hop, zeros where a vertex is
        missing (-1 padding) or carries no features."""
        out = []
        for ids, f in zip(self.ids, self.features):
            d = feat_dim if feat_dim is not None else (f.shape[-1] if f is not None else 0)
            x = np.zeros((ids.shape[0], d), dtype=np.float32)
            if f is not None and d > 0:
                w = min(d, f.shape[-1])
                x[:, :w] = f.reshape(ids.shape[0], -1)[:, :w]
                x[ids < 0] = 0
            out.append(x)
        return out


class Graph(object):
    """Connection to a service front end (Graph.java / impl/GraphImpl.java)."""

    def __init__(self, server_addr: str, timeout: float = 30.0, workers: int = 4, admin_token: str = ""):
        if "://" not in server_addr:
            server_addr = "http://" + server_addr
        self._base, self._timeout = server_addr.rstrip("/"), timeout
        self._token = admin_token          # sent as ``Authorization: Bearer`` on admin calls when the service asks for one
        self._schema: Optional[dict] = None
        self._pool = ThreadPoolExecutor(max_workers=workers)
        self._query: Optional[Query] = None

    @staticmethod
    def connect(server_addr: str, **kw) -> "Graph":
        return Graph(server_addr, **kw)

    # ---- transport
    def _http(self, method: str, path: str, body: Optional[dict] = None, **params) -> dict:
        url = self._base + path + (("?" + urllib.parse.urlencode(params)) if params else "")
        data = json.dumps(body).encode() if body is not None else (b"" if method == "POST" else None)
        headers = {"Content-Type": "application/json"}
        if self._token and path.startswith("/admin/"):
            headers["Authorization"] = "Bearer " + self._token
        req = urllib.request.Request(url, data=data, method=method, headers=headers)
        try:
            with urllib.request.urlopen(req, timeout=self._timeout) as r:
                return json.loads(r.read() or b"{}")
        except urllib.error.HTTPError as e:
            try:
                msg = json.loads(e.read()).get("error", "")
            except Exception:  # noqa: BLE001
                msg = str(e)
            raise UserException("%s %s -> %d %s" % (method, path, e.code, msg)) from None
        except urllib.error.URLError as e:
            raise UserException("cannot reach %s: %s" % (self._base, e.reason)) from None

    # ---- schema
    def get_schema(self) -> dict:
        """{"raw": reference schema JSON, "vertex_id": name -> vtype, "edge_id": name -> etype, "relations": edge name ->
        (src vertex name, dst vertex name)}"""
        if self._schema is None:
            raw = self._http("GET", "/admin/schema")
            vname = {v["vtype"]: v["name"] for v in raw.get("vertex_defs", [])}
            ename = {e["etype"]: e["name"] for e in raw.get("edge_defs", [])}
            self._schema = {"raw": raw, "vertex_id": {n: t for t, n in vname.items()}, "edge_id": {n: t for t, n in ename.items()},
                            "relations": {ename[r["etype"]]: (vname[r["src_vtype"]], vname[r["dst_vtype"]])
                                          for r in raw.get("edge_relation_defs", [])}}
        return self._schema

    # ---- traversal
    def V(self, vtype: str) -> Traversal:
        sch = self.get_schema()
        if vtype not in sch["vertex_id"]:
            raise UserException("unknown vertex type %r" % vtype)
        src = _Node(0, "SOURCE", sch["vertex_id"][vtype])
        return Traversal(self, [src], src, sch["vertex_id"][vtype])

    # ---- install / run
    def install(self, query: Query) -> Status:
        try:
            ans = self._http("POST", "/admin/init", query.to_json())
        except UserException as e:
            return Status(Status.ERROR, str(e))
        query.id = int(ans["query_id"])
        if "node_ids" in ans:
            query.server_ids = {int(k): int(v) for k, v in ans["node_ids"].items()}
        self._query = query
        return Status()

    def install_async(self, query: Query) -> "Future[Status]":
        return self._pool.submit(self.install, query)

    def run(self, query: Query, vids: Optional[Sequence[int]] = None) -> Value:
        if query.id is None:
            raise UserException("the query is not installed")
        if vids is None:
            if query.source is None or not query.source.has_next():
                raise UserException("the query's DataSource is exhausted")
            vids = query.source.next()
        raw = self._http("GET", "/infer", qid=query.id, vid=",".join(str(int(v)) for v in vids))
        return Value(query, raw)

    def run_async(self, query: Query, vids: Optional[Sequence[int]] = None) -> "Future[Value]":
        if vids is None and query.source is not None and query.source.has_next():
            vids = query.source.next()                      # draw in the caller's order, not the pool's
        return self._pool.submit(self.run, query, vids)

    def check_barrier(self, name: str) -> Status:
        st = self._http("GET", "/admin/barrier/status", name=name).get("status")
        return Status() if st == "READY" else Status(Status.NOT_READY, str(st))

    def get_query(self, qid: Optional[int] = None) -> Query:
        """The query registered on the service (``qid`` None: the most recent one)."""
        params = {} if qid is None else {"qid": int(qid)}
        d = self._http("GET", "/admin/query", **params)
        q = Query.from_json(d)
        if "node_ids" in d:
            q.server_ids = {int(k): int(v) for k, v in d["node_ids"].items()}
        return q

    def stats(self) -> dict:
        return self._http("GET", "/admin/stats")

    # ---- data-loader side (the reference's dataloader SDK talks to the same service: records + barriers)
    def load_file(self, pattern_file: str, data_file: str, reverse_edges: Optional[Dict[str, str]] = None) -> int:
        """ask the service to bulk-load a record file IT can read (pattern-file format of the file-loader app)"""
        body = {"pattern": pattern_file, "data": data_file}
        if reverse_edges:
            body["reverse_edges"] = reverse_edges
        return int(self._http("POST", "/admin/load", body)["records"])

    def ingest(self, batch: dict) -> int:
        """push one record batch: {"edges": {etype: {"src", "dst", "ts"[, "weight"]}}, "vertices": {vtype: {"id", "ts", "feat"}}}"""
        def plain(x):
            if isinstance(x, dict):
                return {k: plain(v) for k, v in x.items()}
            return x.tolist() if hasattr(x, "tolist") else x
        return int(self._http("POST", "/admin/ingest", plain(batch))["ingested"])

    def set_barrier(self, name: str, after_records: Optional[int] = None):
        """mark "everything produced so far" (``after_records`` None = the service's current ingest count)"""
        params = {"name": name}
        if after_records is not None:
            params["produced"] = int(after_records)
        self._http("POST", "/admin/barrier/set", None, **params)

    def close(self):
        self._pool.shutdown(wait=False)


class HttpSink(object):
    """Sink of a ``GroupProducer`` that talks to a REMOTE service: every flushed record batch becomes one ``/admin/ingest`` call."""

    P = 1

    def __init__(self, graph: Graph):
        self.graph = graph

    def apply_updates(self, batch: dict):
        self.graph.ingest(batch)

    def set_barrier(self, name: str):
        self.graph.set_barrier(name)


def data_loader(dgs_host: str, max_batch_size: int = 4096, admin_token: str = ""):
    """The data-loader SDK entry point (dataloader/dataloader.h ``Initialize(dgs_host)`` + ``GroupProducer``): connects to the
    service, reads its schema and returns ``(producer, graph)`` - ``producer.add_vertex / add_edge`` batch records per type and
    flush them to the service when a batch is full, ``producer.flush_all()``, ``producer.set_barrier(producer.sink, name)``
    marks "everything produced so far", ``graph.check_barrier(name)`` polls it."""
    from .file_loader import GroupProducer            # numpy-only module
    g = Graph.connect(dgs_host, admin_token=admin_token)
    info = g._http("GET", "/admin/init-info/dataloader")          # Initialize(dgs_host): endpoints, partitions, schema
    g.loader_info = info
    g.get_schema()
    return GroupProducer(HttpSink(g), max_batch_size=max_batch_size, num_partitions=1), g
