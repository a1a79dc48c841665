"""``gl.Graph`` - the user entry point (API parity with
graphlearn/python/graph.py:38-1119).

    g = gl.Graph()
    g.node(path, "item", decoder=gl.Decoder(labeled=True, attr_types=["float"] * 100))
     .edge(path, ("item", "item", "sim"), decoder=gl.Decoder(weighted=True), directed=False)
     .init()
    q = g.V("item").batch(512).shuffle(traverse=True).alias("src") \
         .outV("sim").sample(25).by("random").alias("h1").values()

``init()`` does not start servers: it loads the sources with the native loader,
hash-partitions them over the ranks of the box (one process per GPU, SPMD) and
builds the HBM shards (store/graph_store.py).  "worker mode" of the reference
== running the same script under torchrun; "server mode" (separate sampler
servers) has no equivalent because sampling is a device kernel.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch

from . import config as _config
from . import errors
from .data import values as V_
from .data.decoder import Decoder
from .ops import gather as G
from .ops import sampling as S
from .parallel import runtime as _rt
from .store.graph_store import ORIGIN, REVERSED, GraphStore, Source, Topology
from .utils import Mask, get_mask_type

NODE = 0
EDGE_SRC = 1
EDGE_DST = 2


def _as_tensor(x, device, dtype=torch.int64):
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=dtype)
    return torch.as_tensor(np.asarray(x), dtype=dtype).to(device)


class Graph(object):
    def __init__(self):
        self._node_sources: List[Source] = []
        self._edge_sources: List[Source] = []
        self._node_decoders: Dict[str, Decoder] = {}
        self._edge_decoders: Dict[str, Decoder] = {}
        self._topology = Topology()
        self._undirected_edges: List[str] = []
        self._store: Optional[GraphStore] = None
        self._rt = None
        self._datasets = []
        self._inited = False
        self._with_vineyard = False
        self._node_views: List[tuple] = []

    # ------------------------------------------------------------------ description
    def node(self, source, node_type, decoder=None, option=None, mask=Mask.NONE):
        """Register a node source: a path (file, dir or comma list) or an in-memory dict
        {ids, weights, labels, timestamps, int_attrs, float_attrs, string_attrs}."""
        if not isinstance(source, (str, dict)):
            raise ValueError("source for node() must be a string or a dict of arrays.")
        if not isinstance(node_type, str):
            raise ValueError("node_type for node() must be a string.")
        decoder = decoder or Decoder()
        if not isinstance(decoder, Decoder):
            raise ValueError("decoder must be an instance of Decoder, got {}".format(type(decoder)))
        t = get_mask_type(node_type, mask)
        self._node_decoders[t] = decoder
        if isinstance(source, dict):
            self._node_sources.append(Source("node", None, t, decoder, option=option, data=source))
        else:
            self._node_sources.append(Source("node", source, t, decoder, option=option))
        return self

    def node_view(self, node_type, mask=Mask.NONE, seed=0, nsplit=1, split_range=(0, 1)):
        """Virtual masked view of an existing node type: the nodes whose ``hash(id, seed) % nsplit`` falls in
        ``[split_range[0], split_range[1])`` become ``V(node_type, mask=mask)`` - train / val / test splits
        without extra id files and without loading anything twice (graphlearn/python/graph.py:243-262; the
        reference offers this for its Vineyard backend only, here it works for every source)."""
        if not any(s.types == node_type for s in self._node_sources):
            raise ValueError('Node type "%s" doesn\'t exist.' % (node_type,))
        if mask == Mask.NONE:
            raise ValueError("node_view() needs a TRAIN / VAL / TEST mask")
        mt = get_mask_type(node_type, mask)
        self._node_decoders[mt] = self._node_decoders[node_type]
        self._node_views.append((node_type, mt, int(seed), int(nsplit), (int(split_range[0]), int(split_range[1]))))
        return self

    def edge(self, source, edge_type, decoder=None, directed=True, option=None, mask=Mask.NONE):
        """edge_type = (src_type, dst_type, edge_type).  ``directed=False`` also adds the reversed
        rows: into the same type when src_type == dst_type, otherwise as '<edge_type>_reverse'
        (graphlearn/python/graph.py:357-381)."""
        if not isinstance(source, (str, dict)):
            raise ValueError("source for edge() must be a string or a dict of arrays.")
        if not isinstance(edge_type, tuple) or len(edge_type) != 3:
            raise ValueError("edge_type for edge() must be a tuple of (src_type, dst_tye, edge_type).")
        decoder = decoder or Decoder()
        if not isinstance(decoder, Decoder):
            raise ValueError("decoder must be an instance of Decoder, got {}".format(type(decoder)))
        st, dt, et = edge_type
        met = get_mask_type(et, mask)
        self._edge_decoders[met] = decoder
        self._topology.add(met, st, dt)

        self._add_edge_source(source, (st, dt, met), decoder, ORIGIN, option)
        if not directed:
            self.add_reverse_edges((st, dt, met) if st == dt else (st, dt, et), source, decoder, option)
        return self

    def _add_edge_source(self, source, types, decoder, direction, option):
        if isinstance(source, dict):
            self._edge_sources.append(Source("edge", None, types, decoder, direction, option, data=source))
        else:
            self._edge_sources.append(Source("edge", source, types, decoder, direction, option))     # comma lists / dirs: the loader expands them

    def add_reverse_edges(self, edge_type, source, decoder, option=None):
        """Load ``source`` a second time with the end points swapped (graph.py:357-381): into the same edge type when
        src_type == dst_type, otherwise as ``<edge_type>_reverse``.  ``edge(..., directed=False)`` calls this."""
        st, dt, et = edge_type
        raw = et[len("MASK"):].split("_", 1)[1] if et.startswith("MASK") and "_" in et else et
        if raw not in self._undirected_edges:
            self._undirected_edges.append(raw)
        if st != dt:
            rt_ = et + "_reverse"
            self._edge_decoders[rt_] = decoder
            self._topology.add(rt_, dt, st)
            self._add_edge_source(source, (dt, st, rt_), decoder, REVERSED, option)
        else:
            self._add_edge_source(source, (st, dt, et), decoder, REVERSED, option)
        return self

    @property
    def undirected_edges(self):
        return self._undirected_edges

    def is_directed(self, edge_type):
        return edge_type not in self._undirected_edges

    # ------------------------------------------------------------------ lifecycle
    def init(self, task_index=0, task_count=1, cluster="", job_name="", device=None, **kwargs):
        """Build the sharded graph on this rank's GPU.  ``task_index/task_count/cluster/job_name``
        are accepted for script compatibility; the rank/world come from torchrun env vars."""
        if self._inited:
            return self
        # ---- server mode (graph.py:452-494 of the reference): cluster = {"server": "h:p,h:p", "client_count": C}
        spec = None
        if cluster:
            import json as _json
            spec = _json.loads(cluster) if isinstance(cluster, str) else dict(cluster)
        self._remote = None
        self._server = None
        ep_dir = None
        if spec and job_name in ("server", "client") and not spec.get("server"):
            # FS-tracker server mode: no host list, the servers publish "ip:port" under <tracker>/endpoints/<index> and the
            # clients read them (the reference's FS naming engine, fs_naming_engine.cc:80-153)
            import os as _os
            ep_dir = _os.path.join(_os.path.abspath(spec.get("tracker") or kwargs.get("tracker") or "/tmp/graphlearn"), "endpoints")
            _os.makedirs(ep_dir, exist_ok=True)
        if spec and job_name == "client":
            from .service import RemoteGraphClient
            if ep_dir is not None:
                import os as _os
                import time as _time
                S = int(spec.get("server_count", 1))
                t0 = _time.time()
                while len([f for f in _os.listdir(ep_dir) if f.isdigit()]) < S:
                    if _time.time() - t0 > float(_config.get().timeout or 60) * 5:
                        raise errors.DeadlineExceededError("only %d of %d servers published an endpoint under %s" %
                                                           (len(_os.listdir(ep_dir)), S, ep_dir))
                    _time.sleep(0.1)
                spec["server"] = ",".join(open(_os.path.join(ep_dir, str(i))).read().strip() for i in range(S))
            self._remote = RemoteGraphClient(spec["server"], client_id=int(task_index), client_count=int(spec.get("client_count", 1)))
            meta = self._remote.meta
            self._node_decoders, self._edge_decoders = dict(meta["node_decoders"]), dict(meta["edge_decoders"])
            self._topology = Topology()
            for et, (st, dt) in meta["edges"].items():
                self._topology.add(et, st, dt)
            self._undirected_edges = list(meta["undirected"])
            self._remote_node_types = set(meta["node_types"])
            self._inited = True
            return self
        import os as _os
        if kwargs.get("hosts") and int(task_count or 1) <= 1:
            task_count = len([h for h in str(kwargs["hosts"]).split(",") if h.strip()])      # rpc tracker: one host per worker
        if int(task_count or 1) > 1 and "WORLD_SIZE" not in _os.environ and not (spec and job_name):
            # the reference's worker-mode launch: N plain processes, each with its task index and a shared tracker directory
            # (or a host list) - no torchrun.  Form the process group from those arguments.
            _rt.bootstrap_cluster(int(task_index), int(task_count), kwargs.get("tracker"), kwargs.get("hosts"), device)
        self._rt = _rt.init(device=device)
        self._store = GraphStore(self._rt)
        self._store.node_decoders = self._node_decoders
        self._store.edge_decoders = self._edge_decoders
        self._store.build(self._node_sources, self._edge_sources)
        for base, mt, seed, nsplit, rng in self._node_views:
            self._store.add_node_view(base, mt, seed, nsplit, rng)
        cap = int(_config.get().local_node_cache_capacity)
        if cap > 0:      # reference: set_local_node_cache_capacity -> LFU cache of remote node attrs
            self._store.build_feature_caches(cap)
        self._topology = self._store.topology
        self._inited = True
        if spec and job_name == "server":
            from .service import GraphServer
            from .service.client import _parse
            import os as _os
            addrs = [a for a in str(spec.get("server") or "").split(",") if a]
            S = len(addrs) or int(spec.get("server_count", 1))
            if S > 1 and self._rt.world == 1:
                # several server PROCESSES that are not one process group (the reference's launch: plain processes + tracker):
                # every server holds the whole graph and TRAVERSES its hash share of the ids; look-ups / sampling of any id
                # are local.  (Under torchrun the servers form one sharded group instead.)
                self._traverse_shard = (int(task_index), S)
            address = _parse(addrs[int(task_index) % len(addrs)]) if addrs else ("127.0.0.1", 0)
            self._server = GraphServer(self, address=address, client_count=int(spec.get("client_count", 1)),
                                       server_index=int(task_index), server_count=S).start()
            if ep_dir is not None:
                host, port = self._server.address[0], self._server.address[1]
                tmp = _os.path.join(ep_dir, ".%d.tmp" % int(task_index))
                with open(tmp, "w") as f:
                    f.write("%s:%d" % (host, port))
                _os.replace(tmp, _os.path.join(ep_dir, str(int(task_index))))
        return self

    # ---- the reference's deployment entry points (graph.py:439-511); ``init`` picks one of them from its arguments
    def deploy_in_local_mode(self, task_index=0):
        """one process, one shard (``g.init()`` without a cluster)"""
        return self.init(task_index=task_index)

    def deploy_in_worker_mode(self, tracker=None, hosts=None, task_index=0, task_count=1):
        """every process owns a shard AND trains.  Under ``torchrun`` the ranks already know each other (a mismatching task count
        is an error); launched as N plain processes (the reference's way) the tracker directory or the host list is the
        rendezvous (``parallel/runtime.py:bootstrap_cluster``)."""
        import os as _os
        n = len([h for h in hosts.split(",") if h.strip()]) if hosts else int(task_count)
        if "WORLD_SIZE" in _os.environ and n != int(_os.environ["WORLD_SIZE"]):
            raise ValueError("worker mode with %d tasks inside a %s-process torchrun job" % (n, _os.environ["WORLD_SIZE"]))
        return self.init(task_index=task_index, task_count=n, tracker=tracker, hosts=hosts)

    def deploy_in_server_mode(self, task_index, cluster, job_name):
        """``job_name`` 'server': build this shard and serve it; 'client': connect to the servers of ``cluster``"""
        if job_name not in ("server", "client"):
            raise ValueError("job_name must be 'server' or 'client'")
        if not isinstance(cluster, (dict, str)):
            raise ValueError("cluster must be dict or json string.")
        return self.init(task_index=task_index, cluster=cluster, job_name=job_name)

    def node_attributes(self, node_type, attrs, n_int=0, n_float=0, n_string=0):
        """Keep only the named attribute columns of a node source (the reference offers this for Vineyard property graphs,
        graph.py:265-277): ``attrs`` are 0-based positions (or 'a<i>' names) into the source's attribute list, in the order
        int, float, string; the decoder is rewritten accordingly."""
        return self._select_attrs(self._node_sources, self._node_decoders, node_type, attrs, n_int, n_float, n_string, "Node")

    def edge_attributes(self, edge_type, attrs, n_int=0, n_float=0, n_string=0):
        return self._select_attrs(self._edge_sources, self._edge_decoders, edge_type, attrs, n_int, n_float, n_string, "edge")

    def _select_attrs(self, sources, decoders, type_name, attrs, n_int, n_float, n_string, what):
        srcs = [s for s in sources if (s.types if isinstance(s.types, str) else s.types[2]) == type_name]
        if not srcs:
            raise ValueError('%s type "%s" doesn\'t exist.' % (what, type_name))
        dec = decoders[type_name]
        idx = [int(a[1:]) if isinstance(a, str) and a[:1] == "a" and a[1:].isdigit() else int(a) for a in attrs]
        if len(idx) != n_int + n_float + n_string:
            raise ValueError("%d attributes selected, but n_int + n_float + n_string = %d" % (len(idx), n_int + n_float + n_string))
        from .store.graph_store import storage_kinds
        kinds = storage_kinds(dec)
        for i, t in zip(idx, ["int"] * n_int + ["float"] * n_float + ["string"] * n_string):
            have = kinds[i] if 0 <= i < len(kinds) else None
            if have != t:
                raise ValueError("attribute %d of %s is stored as %r, selected as %r" % (i, type_name, have, t))
        new = Decoder(weighted=dec.weighted, labeled=dec.labeled, timestamped=dec.timestamped,
                      attr_types=[dec.attr_types[i] for i in idx], attr_delimiter=dec.attr_delimiter)
        for s in srcs:
            s.use_attrs = list(idx)
            s.decoder_full, s.decoder = dec, new
        decoders[type_name] = new
        return self

    def vineyard(self, handle, nodes=None, edges=None):
        """Vineyard object stores are not part of this runtime: load the graph from files / arrays instead."""
        raise errors.UnimplementedError("vineyard sources are not supported: use node() / edge() with files or in-memory arrays")

    def init_vineyard(self, server_index=None, worker_index=None, worker_count=None, standalone=False):
        raise ValueError("Not a vineyard graph")

    @property
    def remote(self) -> bool:
        """True for a server-mode CLIENT handle (no local graph; queries run on the servers)."""
        return getattr(self, "_remote", None) is not None

    def close(self):
        if self.remote:
            self._remote.stop()
            self._remote = None
            self._inited = False
            return
        for ds in self._datasets:
            try:
                ds.close()
            except Exception:
                pass
        self._inited = False

    def wait_for_close(self):
        """Server role: serve until every client has stopped (server_impl / fs_coordinator stop protocol); SPMD worker
        ranks simply synchronise."""
        if getattr(self, "_server", None) is not None:
            self._server.wait_for_close()
            self._server = None
        if self._rt is not None:
            self._rt.barrier()

    def add_dataset(self, ds):
        self._datasets.append(ds)

    # ------------------------------------------------------------------ introspection
    @property
    def store(self) -> GraphStore:
        self._check_inited()
        return self._store

    @property
    def runtime(self):
        return self._rt

    @property
    def device(self):
        return self._rt.device if self._rt else torch.device("cpu")

    def _check_inited(self):
        if not self._inited:
            raise errors.FailedPreconditionError("Graph is not initialised; call g.init() first.")

    def get_topology(self):
        return self._topology

    def get_node_decoder(self, node_type):
        d = self._node_decoders.get(node_type)
        if d is None:
            # masked / unmasked fallback
            for k, v in self._node_decoders.items():
                if k.endswith("_" + node_type) and k.startswith("MASK"):
                    return v
            return Decoder()
        return d

    def get_edge_decoder(self, edge_type):
        return self._edge_decoders.get(edge_type, Decoder())

    def get_node_decoders(self):
        return self._node_decoders

    def get_edge_decoders(self):
        return self._edge_decoders

    def get_stats(self):
        """{type: [count on rank 0, count on rank 1, ...]} (GetStats)."""
        self._check_inited()
        if self.remote:                                   # server-mode client: ask the servers this client talks to
            return dict(self._remote.conns[0].call("stats")[1])
        return dict(self._store.stats)

    server_get_stats = get_stats

    # ------------------------------------------------------------------ id helpers
    def _table(self, node_type):
        self._check_inited()
        if node_type not in self._store.nodes:
            raise errors.NotFoundError("unknown node type %r" % (node_type,))
        return self._store.nodes[node_type]

    def _csr(self, edge_type):
        self._check_inited()
        if edge_type not in self._store.edges:
            raise errors.NotFoundError("unknown edge type %r" % (edge_type,))
        return self._store.edges[edge_type]

    def to_vids(self, node_type, ids):
        t = self._table(node_type)
        return t.idmap.to_vid(_as_tensor(ids, self.device))

    def to_ids(self, node_type, vids):
        return self._table(node_type).idmap.to_id(vids)

    # ------------------------------------------------------------------ lookups (O18)
    def lookup_nodes(self, node_type, ids, vids=None) -> V_.Nodes:
        """Attributes / weights / labels / timestamps of nodes; unknown ids get the configured
        defaults (graphlearn/src/core/graph/storage/memory_node_storage.cc:88-138)."""
        tab = self._table(node_type)
        dec = self.get_node_decoder(node_type)
        cfg = _config.get()
        ids_t = _as_tensor(ids, self.device)
        shape = tuple(ids_t.shape)
        flat = ids_t.reshape(-1)
        v = vids.reshape(-1) if vids is not None else tab.idmap.to_vid(flat)
        rt = self._rt
        out = V_.Nodes(ids_t, node_type, shape=shape, graph=self, vids=v)
        out._inited = True
        if dec.float_attr_num > 0 and tab.feats is not None:
            pull = G.gather_rows_dedup if cfg.dedup_feature_pull else G.gather_rows
            out._t["float_attrs"] = pull(rt, tab.feats, tab.feat_desc, v, tab.float_dim, fill=cfg.default_float_attribute)
        if dec.int_attr_num > 0 and tab.ints is not None:
            out._t["int_attrs"] = G.gather_any(rt, tab.ints, v, fill=cfg.default_int_attribute)
        if dec.labeled and tab.labels is not None:
            out._t["labels"] = G.gather_any(rt, tab.labels, v, fill=cfg.default_label)
        if dec.weighted and tab.weights is not None:
            out._t["weights"] = G.gather_any(rt, tab.weights, v, fill=cfg.default_weight)
        if dec.timestamped and tab.timestamps is not None:
            out._t["timestamps"] = G.gather_any(rt, tab.timestamps, v, fill=cfg.default_timestamp)
        if dec.string_attr_num > 0 and tab.strings is not None:
            out._t["string_attrs"] = self._lookup_strings(tab, v)
        return out

    def _lookup_strings(self, tab, vids):
        return tab.lookup_strings(vids, _config.get().default_string_attribute)

    def lookup_edges(self, edge_type, src_ids, edge_ids, src_vids=None) -> V_.Edges:
        csr = self._csr(edge_type)
        dec = self.get_edge_decoder(edge_type)
        cfg = _config.get()
        src_t = _as_tensor(src_ids, self.device)
        eid_t = _as_tensor(edge_ids, self.device)
        shape = tuple(eid_t.shape)
        W = self._rt.world
        sv = src_vids.reshape(-1) if src_vids is not None else self._table(csr.src_type).idmap.to_vid(src_t.reshape(-1))
        if sv.numel() != eid_t.numel() and sv.numel() > 0:
            sv = sv.reshape(-1, 1).expand(-1, eid_t.numel() // sv.numel()).reshape(-1)
        # edge rows live on the SOURCE's owner at position edge_id: address them as vid = eid * W + owner
        key = torch.where(eid_t.reshape(-1) >= 0, eid_t.reshape(-1) * W + (sv.clamp(min=0) % W),
                          torch.full_like(sv, -1))
        out = V_.Edges(src_t, csr.src_type, None, csr.dst_type, edge_type, eid_t, shape=shape, graph=self)
        out._inited = True
        rt = self._rt
        if dec.weighted and csr.weights is not None:
            out._t["weights"] = G.gather_any(rt, csr.weights, key, fill=cfg.default_weight)
        if dec.labeled and csr.labels is not None:
            out._t["labels"] = G.gather_any(rt, csr.labels, key, fill=cfg.default_label)
        if dec.timestamped and csr.ts is not None:
            out._t["timestamps"] = G.gather_any(rt, csr.ts, key, fill=cfg.default_timestamp)
        if dec.float_attr_num > 0 and csr.float_attrs is not None:
            out._t["float_attrs"] = G.gather_any(rt, csr.float_attrs, key, fill=cfg.default_float_attribute)
        if dec.int_attr_num > 0 and csr.int_attrs is not None:
            out._t["int_attrs"] = G.gather_any(rt, csr.int_attrs, key, fill=cfg.default_int_attribute)
        if dec.string_attr_num > 0 and (W > 1 or getattr(csr, "strings", None) is not None):
            # edge strings sit on the host of the SOURCE's owner at position edge_id: key = eid * W + owner
            kk = key.cpu().numpy()
            n_str = dec.string_attr_num
            strings = getattr(csr, "strings", None)

            def local_rows(req):
                e = req // W
                out_ = np.full((len(req), n_str), cfg.default_string_attribute, dtype=object)
                if strings is not None:
                    ok = (req >= 0) & (e < len(strings))
                    if ok.any():
                        out_[ok] = np.asarray(strings, dtype=object).reshape(len(strings), -1)[e[ok]]
                return out_

            if W == 1:
                arr = local_rows(kk)
            else:
                arr = np.full((len(kk), n_str), cfg.default_string_attribute, dtype=object)
                owner = np.where(kk >= 0, kk % W, -1)
                asked = rt.all_gather_object([kk[owner == o] for o in range(W)])
                got = rt.all_gather_object([local_rows(np.asarray(asked[q][rt.rank], dtype=np.int64)) for q in range(W)])
                for o in range(W):
                    m = owner == o
                    if m.any():
                        arr[m] = got[o][rt.rank]
            out._t["string_attrs"] = arr
        return out

    def get_nodes(self, node_type, ids, offsets=None, shape=None):
        """Construct Nodes / SparseNodes bound to this graph (attributes are fetched lazily)."""
        if offsets is None:
            return V_.Nodes(ids, node_type, shape=shape, graph=self)
        return V_.SparseNodes(ids, offsets, shape, node_type, graph=self)

    def find_edge_ids(self, edge_type, src_ids, dst_ids):
        """edge id of every (src, dst) pair (first match in the source's adjacency row, -1 when the edge does not
        exist).  The reference cannot look up attributes of edges given only by their end points; this resolves
        them with one full-row fetch + match (collective on the portable multi-rank path)."""
        csr = self._csr(edge_type)
        s = _as_tensor(src_ids, self.device).reshape(-1)
        d = _as_tensor(dst_ids, self.device).reshape(-1)
        sv = self._table(csr.src_type).idmap.to_vid(s)
        dv = self._table(csr.dst_type).idmap.to_vid(d)
        vals, eids, offs = S.sample_full(csr, sv, cap=0, want_eids=True)
        counts = offs[1:] - offs[:-1]
        seg = torch.repeat_interleave(torch.arange(s.numel(), device=self.device), counts)
        hit = vals == dv[seg]
        out = torch.full((s.numel(),), -1, dtype=torch.int64, device=self.device)
        # first match wins: scatter in reverse order so that earlier positions overwrite later ones
        idx = torch.nonzero(hit).flatten().flip(0)
        out[seg[idx]] = eids[idx]
        return out

    def get_edges(self, edge_type, src_ids, dst_ids, edge_ids=None, offsets=None, shape=None):
        st, dt = self._topology.get_src_type(edge_type), self._topology.get_dst_type(edge_type)
        if offsets is None:
            return V_.Edges(src_ids, st, dst_ids, dt, edge_type, edge_ids, shape=shape, graph=self)
        return V_.SparseEdges(src_ids, st, dst_ids, dt, edge_type, offsets, shape, edge_ids=edge_ids, graph=self)

    # ------------------------------------------------------------------ degrees / aggregation
    def out_degrees(self, ids, edge_type):
        csr = self._csr(edge_type)
        v = self.to_vids(csr.src_type, ids)
        d = S.get_degrees(csr, v.reshape(-1))
        return d.reshape(v.shape).cpu().numpy()

    def in_degrees(self, ids, edge_type):
        """The reference returns Unimplemented for in-degrees (degree_getter.cc:48-53); here the
        in-edge CSR is built on demand."""
        csr = self._csr(edge_type)
        rev = self._store.reverse_csr(edge_type)
        v = self.to_vids(csr.dst_type, ids)
        d = S.get_degrees(rev, v.reshape(-1))
        return d.reshape(v.shape).cpu().numpy()

    def aggregate_nodes(self, node_type, ids, func="sum", k=0, offsets=None, vids=None):
        """Server-side segment aggregation of float attributes (Aggregator operators O16)."""
        tab = self._table(node_type)
        v = vids if vids is not None else tab.idmap.to_vid(_as_tensor(ids, self.device).reshape(-1))
        return G.gather_agg(self._rt, tab.feats, tab.feat_desc, v, tab.float_dim, func, offsets=offsets, k=k)

    # ------------------------------------------------------------------ GSL entry points
    def V(self, t, feed=None, node_from=NODE, mask=Mask.NONE):
        from .gsl.dag import Dag
        from .gsl.dag_node import TraverseVertexDagNode
        if feed is not None:
            raise NotImplementedError("`feed` is not supported for V() yet (same as the reference).")
        self._check_inited()
        dag = Dag(self)
        if node_from == NODE:
            node_type = get_mask_type(t, mask)
            if node_type not in (self._remote_node_types if self.remote else self._store.nodes):
                raise ValueError("node type %r not in graph" % (node_type,))
            params = {"node_from": NODE, "node_type": node_type}
            out_type = node_type
            base_type = t
        else:
            et = get_mask_type(t, mask)
            if not self._topology.is_exist(et):
                raise ValueError("edge type %r not in graph" % (et,))
            params = {"node_from": node_from, "edge_type": et}
            out_type = self._topology.get_src_type(et) if node_from == EDGE_SRC else self._topology.get_dst_type(et)
            base_type = out_type
        node = TraverseVertexDagNode(dag, op_name="GetNodes", params=params)
        node.set_output_type(out_type, base_type)
        dag.root = node
        return node

    def E(self, edge_type, feed=None, reverse=False, mask=Mask.NONE):
        from .gsl.dag import Dag
        from .gsl.dag_node import TraverseSourceEdgeDagNode
        if feed is not None:
            raise NotImplementedError("`feed` is not supported for E() yet (same as the reference).")
        self._check_inited()
        et = get_mask_type(edge_type, mask)
        if not self._topology.is_exist(et):
            raise ValueError("edge type %r not in graph" % (et,))
        dag = Dag(self)
        node = TraverseSourceEdgeDagNode(dag, op_name="GetEdges", params={"edge_type": et, "reverse": reverse})
        dag.root = node
        return node

    def SubGraph(self, seed_type, nbr_type, batch_size=64, strategy="random_node", num_nbrs=None, need_dist=False):
        """Root of a subgraph-sampling query (graphlearn/python/graph.py:629-667)."""
        from .gsl.dag import Dag
        from .gsl.dag_node import SubGraphDagNode
        self._check_inited()
        dag = Dag(self)
        node = SubGraphDagNode(dag, params={"seed_type": seed_type, "nbr_type": nbr_type, "batch_size": batch_size,
                                            "strategy": strategy, "num_nbrs": list(num_nbrs or []),
                                            "need_dist": need_dist})
        dag.root = node
        return node

    # ------------------------------------------------------------------ imperative samplers (P4)
    def node_sampler(self, t, batch_size=64, strategy="by_order", node_from=NODE, mask=Mask.NONE):
        from .sampler.node_sampler import NodeSampler
        return NodeSampler(self, get_mask_type(t, mask) if node_from == NODE else t, batch_size, strategy, node_from)

    def edge_sampler(self, edge_type, batch_size=64, strategy="by_order", mask=Mask.NONE):
        from .sampler.edge_sampler import EdgeSampler
        return EdgeSampler(self, get_mask_type(edge_type, mask), batch_size, strategy)

    def neighbor_sampler(self, meta_path, expand_factor, strategy="random"):
        from .sampler.neighbor_sampler import FullNeighborSampler, NeighborSampler
        if strategy == "full":
            return FullNeighborSampler(self, meta_path, expand_factor)
        return NeighborSampler(self, meta_path, expand_factor, strategy)

    def negative_sampler(self, object_type, expand_factor, strategy="random", conditional=False, **kwargs):
        from .sampler.negative_sampler import ConditionalNegativeSampler, NegativeSampler
        if conditional:
            return ConditionalNegativeSampler(self, object_type, expand_factor, strategy, **kwargs)
        return NegativeSampler(self, object_type, expand_factor, strategy)

    def subgraph_sampler(self, seed_type=None, nbr_type=None, batch_size=64, strategy="random_node", num_nbrs=None,
                         need_dist=False):
        """``seed_type`` defaults to the source node type of ``nbr_type`` (the reference's own tests call
        ``g.subgraph_sampler(nbr_type="relation")``)."""
        from .sampler.subgraph_sampler import SubGraphSampler
        if isinstance(nbr_type, (list, tuple)):                  # g.subgraph_sampler('relation', [10, 10]): the class-level signature
            num_nbrs, nbr_type = list(nbr_type), None
        if nbr_type is None and seed_type is not None and seed_type in self._store.edges:
            seed_type, nbr_type = None, seed_type
        if nbr_type is None:
            raise ValueError("subgraph_sampler needs the neighbour (edge) type")
        if seed_type is None:
            seed_type = self._store.edges[nbr_type].src_type
        return SubGraphSampler(self, seed_type, nbr_type, batch_size, strategy, num_nbrs, need_dist)

    def search(self, node_type, inputs, option):
        """KNN search over a node type's float attributes (KnnOperator)."""
        from .ops import knn
        tab = self._table(node_type)
        q = _as_tensor(inputs, self.device, torch.float32)
        ids, dist = knn.search(self._rt, tab, q, option.k, metric=_config.get().knn_metric)
        return ids.cpu().numpy(), dist.cpu().numpy()

    def get_client(self):
        return self
