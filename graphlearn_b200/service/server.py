"""GraphServer: executes shipped GSL queries on the local graph and streams batches to clients."""
from __future__ import annotations

import threading
from multiprocessing.connection import Listener
from typing import Dict, Tuple

import numpy as np

from .. import errors
from ..data import values as V_

AUTHKEY = b"graphlearn_b200"


def _np(x):
    return None if x is None else np.asarray(x)


def encode_value(v, no_property: bool = False) -> dict:
    """Nodes / Edges -> plain dict of numpy arrays (ids + every attribute the decoder declares: the reference ships the
    LookupNodes / LookupEdges results inside the tape, dag_node.py:558-564,595-610)."""
    if isinstance(v, V_.SubGraph):
        raise errors.UnimplementedError("server mode streams Nodes / Edges values (sub-graph queries run in worker mode)")
    sparse = None
    if isinstance(v, (V_.SparseNodes, V_.SparseEdges)):      # 'full' neighbourhoods: values + per-row counts + dense shape
        sparse = {"offsets": _np(v.offsets), "dense_shape": tuple(v.dense_shape)}
    if isinstance(v, V_.Edges):
        d = {"kind": "edges", "edge_type": v.edge_type, "src_type": v.src_type, "dst_type": v.dst_type, "shape": tuple(v.shape),
             "src_ids": _np(v.src_ids), "dst_ids": _np(v.dst_ids), "edge_ids": _np(v.edge_ids)}
        dec = v._get_decoder()
    elif isinstance(v, V_.Nodes):
        d = {"kind": "nodes", "type": v.type, "shape": tuple(v.shape), "ids": _np(v.ids)}
        dec = v._get_decoder()
    else:
        raise errors.UnimplementedError("cannot ship %r" % (type(v),))
    if sparse is not None:
        d["sparse"] = sparse
    if not no_property:
        if dec.float_attr_num:
            d["float_attrs"] = _np(v.float_attrs)
        if dec.int_attr_num:
            d["int_attrs"] = _np(v.int_attrs)
        if dec.string_attr_num:
            d["string_attrs"] = _np(v.string_attrs)
        if dec.weighted:
            d["weights"] = _np(v.weights)
        if dec.labeled:
            d["labels"] = _np(v.labels)
        if dec.timestamped:
            d["timestamps"] = _np(v.timestamps)
    return d


class GraphServer(object):
    def __init__(self, graph, address: Tuple[str, int] = ("127.0.0.1", 0), client_count: int = 1, authkey: bytes = AUTHKEY,
                 server_index: int = 0, server_count: int = 1):
        graph._check_inited()
        self.g = graph
        self._listener = Listener(address, authkey=authkey)
        self.address = self._listener.address
        # how many clients will say STOP to THIS server: the same round-robin deal the clients apply
        # (service/client.py RemoteGraphClient: S >= C -> client c owns servers {s : s % C == c}; else client c uses server c % S)
        S, C = max(1, int(server_count)), max(1, int(client_count))
        self.client_count = 1 if S >= C else len([c for c in range(C) if c % S == int(server_index)])
        if S == 1:
            self.client_count = C
        self._stops = 0
        self._lock = threading.Lock()
        self._done = threading.Event()
        self._threads = []
        self._accept = threading.Thread(target=self._accept_loop, daemon=True)

    def start(self):
        self._accept.start()
        return self

    def wait_for_close(self, timeout=None):
        """Blocks until every client has said STOP (the reference's stop protocol: fs_coordinator.cc:110-118)."""
        self._done.wait(timeout)
        try:
            self._listener.close()
        except Exception:
            pass
        for t in list(self._threads):          # let the connection threads close their datasets before the process exits
            t.join(timeout=10)

    # ------------------------------------------------------------------ internals
    def _accept_loop(self):
        while not self._done.is_set():
            try:
                conn = self._listener.accept()
            except Exception:
                return
            t = threading.Thread(target=self._serve, args=(conn,), daemon=True)
            t.start()
            self._threads.append(t)

    def _meta(self):
        g = self.g
        topo = g.get_topology()
        return {"edges": {e: (topo.get_src_type(e), topo.get_dst_type(e)) for e in topo.edge_types()},
                "node_types": sorted(g.store.nodes.keys()),
                "node_decoders": dict(g.get_node_decoders()), "edge_decoders": dict(g.get_edge_decoders()),
                "undirected": list(g.undirected_edges), "stats": g.get_stats()}

    def _serve(self, conn):
        from ..gsl.dag import Dag
        from ..gsl.dataset import Dataset
        datasets: Dict[int, Tuple[object, Dataset]] = {}
        try:
            while True:
                try:
                    req = conn.recv()
                except (EOFError, OSError):
                    return
                op = req[0]
                try:
                    if op == "meta":
                        conn.send(("ok", self._meta()))
                    elif op == "run_dag":
                        _, dag_def, window = req
                        with self._lock:                       # queries are built against the shared graph object
                            dag = Dag.from_def(self.g, dag_def)
                            ds = Dataset(dag, window=window)
                        datasets[len(datasets) + 1] = (dag, ds)
                        conn.send(("ok", len(datasets)))
                    elif op == "next":
                        dag, ds = datasets[req[1]]
                        try:
                            with self._lock:
                                vals = ds.next()
                                payload = {a: encode_value(v, bool(dag.get_node(a).params.get("no_property"))) for a, v in vals.items()}
                            conn.send(("ok", payload))
                        except errors.OutOfRangeError:
                            conn.send(("eoe", ds.epoch))
                    elif op == "lookup_nodes":
                        with self._lock:
                            conn.send(("ok", encode_value(self.g.lookup_nodes(req[1], req[2]))))
                    elif op == "stats":
                        conn.send(("ok", self.g.get_stats()))
                    elif op == "stop":
                        conn.send(("ok", None))
                        with self._lock:
                            self._stops += 1
                            if self._stops >= self.client_count:
                                self._done.set()
                        return
                    else:
                        conn.send(("error", "unknown request %r" % (op,)))
                except Exception as e:          # the server survives a bad request; the client gets the message
                    conn.send(("error", "%s: %s" % (type(e).__name__, e)))
        finally:
            for _, ds in datasets.values():
                try:
                    ds.close()
                except Exception:
                    pass
            conn.close()
