"""Client side of server mode: a graph handle without a graph."""
from __future__ import annotations

from multiprocessing.connection import Client
from typing import List, Sequence, Tuple

from .. import errors
from ..data import values as V_
from .server import AUTHKEY


def _parse(addr) -> Tuple[str, int]:
    if isinstance(addr, (tuple, list)):
        return str(addr[0]), int(addr[1])
    host, port = str(addr).rsplit(":", 1)
    return host, int(port)


class _Conn(object):
    def __init__(self, addr, authkey=AUTHKEY, connect_timeout: float = None):
        """A server that is still loading its graph does not listen yet: retry the connection until ``gl.set_timeout`` seconds
        have passed (the reference's channel manager re-resolves broken channels every second, channel_manager.cc:151-170)."""
        import time
        from .. import config as _config
        self.addr = _parse(addr)
        deadline = time.time() + float(connect_timeout if connect_timeout is not None else (_config.get().timeout or 60))
        while True:
            try:
                self._c = Client(self.addr, authkey=authkey)
                break
            except (ConnectionRefusedError, ConnectionResetError, FileNotFoundError):
                if time.time() >= deadline:
                    raise errors.UnavailableError("graph server %s:%d is not reachable" % self.addr)
                time.sleep(0.2)

    def call(self, *req):
        self._c.send(tuple(req))
        status, payload = self._c.recv()
        if status == "error":
            raise errors.InternalError("graph server %s:%d: %s" % (self.addr + (payload,)))
        return status, payload

    def close(self):
        try:
            self._c.close()
        except Exception:
            pass


def decode_value(d: dict, graph=None):
    sp = d.get("sparse")
    attrs = dict(int_attrs=d.get("int_attrs"), float_attrs=d.get("float_attrs"), string_attrs=d.get("string_attrs"),
                 weights=d.get("weights"), labels=d.get("labels"), timestamps=d.get("timestamps"))
    if sp is not None and d["kind"] == "nodes":
        v = V_.SparseNodes(d["ids"], sp["offsets"], tuple(sp["dense_shape"]), d["type"], graph=graph, **attrs)
        v._inited = True
        return v
    if sp is not None:
        v = V_.SparseEdges(d["src_ids"], d["src_type"], d["dst_ids"], d["dst_type"], d["edge_type"], sp["offsets"],
                           tuple(sp["dense_shape"]), edge_ids=d["edge_ids"], graph=graph, **attrs)
        v._inited = True
        return v
    if d["kind"] == "nodes":
        v = V_.Nodes(d["ids"], d["type"], int_attrs=d.get("int_attrs"), float_attrs=d.get("float_attrs"),
                     string_attrs=d.get("string_attrs"), weights=d.get("weights"), labels=d.get("labels"),
                     timestamps=d.get("timestamps"), shape=d["shape"], graph=graph)
    else:
        v = V_.Edges(d["src_ids"], d["src_type"], d["dst_ids"], d["dst_type"], d["edge_type"], d["edge_ids"],
                     int_attrs=d.get("int_attrs"), float_attrs=d.get("float_attrs"), string_attrs=d.get("string_attrs"),
                     weights=d.get("weights"), labels=d.get("labels"), timestamps=d.get("timestamps"), shape=d["shape"], graph=graph)
    v._inited = True            # everything the decoder declares came with the batch: no lazy lookups on the client
    return v


class RemoteGraphClient(object):
    """Connections of ONE client to the servers it owns.  With S servers and C clients, client i owns the servers
    {s : s % C == i} (or server i % S when there are fewer servers than clients) - the reference's round-robin deal."""

    def __init__(self, servers: Sequence, client_id: int = 0, client_count: int = 1):
        servers = [s for s in (servers.split(",") if isinstance(servers, str) else servers) if s]
        if not servers:
            raise ValueError("server mode needs at least one server address")
        S, C = len(servers), max(1, int(client_count))
        mine = [s for i, s in enumerate(servers) if i % C == client_id % C] if S >= C else [servers[client_id % S]]
        self.conns: List[_Conn] = [_Conn(a) for a in mine]
        self.meta = self.conns[0].call("meta")[1]

    def stop(self):
        for c in self.conns:
            try:
                c.call("stop")
            except Exception:
                pass
            c.close()
        self.conns = []


class RemoteDataset(object):
    """``gl.Dataset`` of a client: the query runs on the servers, ``next()`` pulls one whole batch; when a server
    reports the end of its epoch the next server is asked, after the last one the epoch ends here as well
    (graphlearn/python/gsl/dag_dataset.py:84-96)."""

    def __init__(self, graph, dag, window=10):
        self._graph, self._dag = graph, dag
        self._conns = graph._remote.conns
        d = dag.to_def()
        self._ids = [c.call("run_dag", d, int(window))[1] for c in self._conns]
        self._cur = 0
        self.epoch = 0

    def next(self):
        while self._cur < len(self._conns):
            status, payload = self._conns[self._cur].call("next", self._ids[self._cur])
            if status == "ok":
                from ..gsl.dataset import DagValues
                res = DagValues({a: decode_value(v, self._graph) for a, v in payload.items()})
                f = self._dag.value_func
                return f(res) if f is not None else res
            self._cur += 1
        self._cur = 0
        self.epoch += 1
        raise errors.OutOfRangeError("out of range")

    __next__ = next

    def __iter__(self):
        return self

    def close(self):
        pass
