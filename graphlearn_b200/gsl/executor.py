"""Executes a GSL query plan on the device, one batch per call.

This replaces the reference's server-side continuous DAG execution
(DagScheduler -> DagNodeRunner -> OpRunner -> Tape,
graphlearn/src/core/runner/dag_scheduler.cc:45-86, dag_node_runner.cc:32-98):
the plan is walked in topological (construction) order and every traversal is
one kernel launch over the sharded store; the per-alias results are ``Nodes``
/ ``Edges`` objects backed by device tensors.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from .. import config as _config
from .. import errors
from ..data import values as V_
from ..ops import negative as NEG
from ..ops import rng as rng_ops
from ..ops import sampling as S
from ..ops import subgraph as SUB
from ..ops import walk as WALK
from .dag_node import (DagNode, FakeNode, SubGraphDagNode, TraverseEdgeDagNode, TraverseNegVertexDagNode,
                       TraverseSourceEdgeDagNode, TraverseVertexDagNode)
from .iterators import SeedIterator

NODE, EDGE_SRC, EDGE_DST = 0, 1, 2


class _Out(object):
    """Per-node execution result: ids (+ vids in the node type's base table) or an edge batch."""
    __slots__ = ("ids", "vids", "shape", "value", "src_ids", "src_vids", "eids")

    def __init__(self, ids=None, vids=None, shape=None, value=None):
        self.ids, self.vids, self.shape, self.value = ids, vids, shape, value
        self.src_ids = self.src_vids = self.eids = None


class QueryExecutor(object):
    def __init__(self, dag, drop_last: bool = False, seed: Optional[int] = None, sync_epoch: Optional[bool] = None):
        if not dag.is_ready():
            raise ValueError("query is not ready: end it with .values()")
        self.dag = dag
        self.g = dag.graph
        self.store = self.g.store
        self.rt = self.g.runtime
        self.rng = rng_ops.DeviceRng(self.rt, _config.get().seed if seed is None else seed)
        self.drop_last = drop_last
        self._iter: Optional[SeedIterator] = None
        self._order = self._toposort()
        self._salt = 0
        if sync_epoch is None:
            # auto: on whenever the ops of this query are collectives - always on the portable path; on the peer
            # kernel path only id translation of non-dense id spaces and host-side string attributes are
            collective = not (self.rt.is_cuda and _config.get().use_peer_kernels)
            for tab in self.store.nodes.values():
                collective = collective or tab.strings is not None or tab.idmap.collective
            for csr in self.store.edges.values():
                collective = collective or getattr(csr, "strings", None) is not None
            sync_epoch = self.rt.world > 1 and collective
        self.sync_epoch = bool(sync_epoch) and self.rt.world > 1

    # ------------------------------------------------------------------ plan
    def _toposort(self):
        """Topological order of the DAG.  Besides its upstream a node may depend on a SIBLING branch:
        ``filter(alias)`` and ``where(alias, condition)`` read another node's output (the reference wires these
        as extra DagEdges, dag_node.py:236-304), so edges = upstream + filter target + condition target."""
        nodes, seen = [], set()

        def collect(n: DagNode):
            if id(n) in seen:
                return
            seen.add(id(n))
            nodes.append(n)
            for d in n.downstreams:
                collect(d)

        collect(self.dag.root)

        def deps(n: DagNode):
            out = [n.upstream] if n.upstream is not None else []
            if n._filter is not None:
                out.append(n._filter)
            if isinstance(n.params.get("dst_node"), DagNode):
                out.append(n.params["dst_node"])
            return [d for d in out if id(d) in seen]

        order, done = [], set()

        def visit(n: DagNode, stack=()):
            if id(n) in done:
                return
            if id(n) in stack:
                raise errors.InvalidArgumentError("cyclic filter / where dependency in the GSL query")
            for d in deps(n):
                visit(d, stack + (id(n),))
            done.add(id(n))
            order.append(n)

        for n in nodes:
            visit(n)
        return order

    # ------------------------------------------------------------------ roots
    def _ensure_root_iter(self):
        """Create the seed iterator of the root (this rank's nodes / edges of the root type) on first use."""
        if self._iter is not None:
            return
        node, r, dev = self.dag.root, self.rt.rank, self.rt.device
        p = node.params
        seed = _config.get().seed + 17 * r
        if isinstance(node, SubGraphDagNode):
            tab = self.store.nodes[p["seed_type"]]
            self._iter = SeedIterator(tab.n_local, int(p["batch_size"]), "shuffle" if "random" in p["strategy"] else "by_order",
                                      dev, seed=seed)
        elif isinstance(node, TraverseSourceEdgeDagNode) or p.get("node_from", NODE) != NODE:
            csr = self.store.edges[p["edge_type"]]
            n_edges = csr.n_edges
            self._edge_sel = None
            shard = getattr(self.g, "_traverse_shard", None)
            if shard is not None and n_edges > 0:
                # replicated multi-server mode: this server traverses the edges whose SOURCE id hashes to it
                pos = csr.insertion_pos()
                src_ids = self.g.to_ids(csr.src_type, csr._row_of_edge[pos] * self.rt.world + r)
                self._edge_sel = (src_ids.abs() % shard[1] == shard[0]).nonzero().flatten()
                n_edges = int(self._edge_sel.numel())
            self._iter = SeedIterator(n_edges, int(p.get("batch_size", 64)), p.get("strategy", "by_order"), dev,
                                      seed=seed, drop_last=self.drop_last)
        else:
            tab = self.store.nodes[node.type]
            rows = tab.present.nonzero().flatten() if tab.present is not None else torch.arange(tab.n_local, device=dev)
            shard = getattr(self.g, "_traverse_shard", None)
            if shard is not None and rows.numel() > 0:
                ids_ = tab.idmap.to_id(rows * self.rt.world + r) if not tab.idmap.dense else rows * self.rt.world + r
                rows = rows[ids_.abs() % shard[1] == shard[0]]         # this server's share of the node ids
            self._rows = rows
            self._iter = SeedIterator(int(rows.numel()), int(p.get("batch_size", 64)), p.get("strategy", "by_order"), dev,
                                      seed=seed, drop_last=self.drop_last)

    def _sync_epoch_end(self):
        """Multi-rank lock step: when ANY rank has exhausted its shard, every rank ends the epoch now (ranks that
        still had batches drop them, like ``DistributedSampler(drop_last=True)``).  Needed whenever the query's
        ops are collectives (portable path; non-dense id maps / string lookups also on GPUs): a rank that stopped
        iterating would leave the others blocked inside an all-to-all."""
        import torch.distributed as dist
        self._ensure_root_iter()
        flag = torch.tensor([1 if self._iter.has_next() else 0], dtype=torch.int32, device=self.rt.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            self._iter.end_epoch()
            raise errors.OutOfRangeError("end of epoch (synchronised across ranks)")

    def _root_nodes(self, node: DagNode) -> _Out:
        p = node.params
        bs = int(p.get("batch_size", 64))
        strategy = p.get("strategy", "by_order")
        W, r = self.rt.world, self.rt.rank
        if p.get("node_from", NODE) == NODE:
            tab = self.store.nodes[node.type]
            self._ensure_root_iter()
            if _config.get().actor_enabled and W > 1 and strategy in ("shuffle", "by_order"):
                vids = self._next_dispatched(node, bs, strategy)       # gl.enable_actor(): balanced batch dispatch
            else:
                idx = self._iter.next_index()
                rows = self._rows[idx]
                vids = rows * W + r
            ids = tab.idmap.to_id(vids) if not tab.idmap.dense else vids
            # vids must live in the BASE table (edges index the unmasked node type)
            base = node._base_type
            vtype = node.type
            if base != node.type and base in self.store.nodes:
                # masked roots (V(t, mask=...)) only choose WHICH ids are traversed; labels and
                # attributes are looked up in the base table (graph.py:582-588: set_path(t, ...))
                bvids = self.store.nodes[base].idmap.to_vid(ids)
                vtype, vids = base, bvids
            else:
                bvids = vids
            out = _Out(ids=ids, vids=bvids, shape=(int(ids.numel()),))
            out.value = V_.Nodes(ids, vtype, shape=out.shape, graph=self.g, vids=vids)
            return out
        # nodes from edge end points
        csr = self.store.edges[p["edge_type"]]
        self._ensure_root_iter()
        # traversal follows the INSERTION order of the edges (the reference's edge id = insertion
        # index, memory_edge_storage.cc; GetEdges by_order walks edge ids) - chronological event
        # files therefore yield chronological batches although the CSR is row-major
        nxt = self._iter.next_index()
        idx = csr.insertion_pos()[nxt if getattr(self, "_edge_sel", None) is None else self._edge_sel[nxt]]
        if p["node_from"] == EDGE_SRC:
            vids = csr._row_of_edge[idx] * W + r
            t = csr.src_type
        else:
            vids = csr.indices.local[idx]
            t = csr.dst_type
        ids = self.g.to_ids(t, vids)
        out = _Out(ids=ids, vids=vids, shape=(int(ids.numel()),))
        out.value = V_.Nodes(ids, t, shape=out.shape, graph=self.g, vids=vids)
        return out

    def _next_dispatched(self, node: DagNode, bs: int, strategy: str) -> torch.Tensor:
        """Root batches under ``gl.enable_actor()``: every epoch the ranks pool their seed batches and a data-size
        aware plan (engine/dispatch.py; reference: actor/runner/tape_dispatcher.cc:61-175) hands every rank the same
        number of batches with balanced first-hop work.  Collective at epoch boundaries only."""
        from ..engine.dispatch import BalancedSeedDispatcher
        W, r, dev = self.rt.world, self.rt.rank, self.rt.device
        if getattr(self, "_planned", None) is None:
            rows = self._rows
            if strategy == "shuffle":
                g = torch.Generator(device="cpu").manual_seed(_config.get().seed * 7919 + 31 * r + self._iter.epoch)
                rows = rows[torch.randperm(int(rows.numel()), generator=g).to(dev)]
            vids = rows * W + r
            weights = None
            for d in node.downstreams:                     # work estimate: out-degree along the first sampled edge type
                et = d.params.get("edge_type")
                tab = self.store.nodes[node.type]
                if et in tab.out_degrees:
                    weights = tab.out_degrees[et][rows].float() + 1.0
                    break
            self._planned = BalancedSeedDispatcher(self.rt, bs).plan(vids, weights)
            self._plan_iter = SeedIterator(int(self._planned.numel()), bs, "by_order", dev)
        try:
            idx = self._plan_iter.next_index()
        except errors.OutOfRangeError:
            self._planned = None
            self._iter.epoch += 1
            raise
        return self._planned[idx]

    def _root_edges(self, node: DagNode) -> _Out:
        p = node.params
        csr = self.store.edges[p["edge_type"]]
        W, r = self.rt.world, self.rt.rank
        bs = int(p.get("batch_size", 64))
        self._ensure_root_iter()
        # traversal follows the INSERTION order of the edges (the reference's edge id = insertion
        # index, memory_edge_storage.cc; GetEdges by_order walks edge ids) - chronological event
        # files therefore yield chronological batches although the CSR is row-major
        nxt = self._iter.next_index()
        idx = csr.insertion_pos()[nxt if getattr(self, "_edge_sel", None) is None else self._edge_sel[nxt]]
        src_v = csr._row_of_edge[idx] * W + r
        dst_v = csr.indices.local[idx]
        src_ids = self.g.to_ids(csr.src_type, src_v)
        dst_ids = self.g.to_ids(csr.dst_type, dst_v)
        st, dt = csr.src_type, csr.dst_type
        if p.get("reverse"):
            src_ids, dst_ids, src_v, dst_v, st, dt = dst_ids, src_ids, dst_v, src_v, dt, st
        out = _Out(ids=dst_ids, vids=dst_v, shape=(int(idx.numel()),))
        out.src_ids, out.src_vids, out.eids = src_ids, src_v, idx
        out.value = V_.Edges(src_ids, st, dst_ids, dt, p["edge_type"], idx, shape=out.shape, graph=self.g,
                             src_vids=(csr._row_of_edge[idx] * W + r))
        return out

    # ------------------------------------------------------------------ traversals
    def _sample(self, node: DagNode, up: _Out, results) -> _Out:
        p = node.params
        et = p["edge_type"]
        direction = p.get("direction", "out")
        strategy = p.get("strategy", "random")
        k = int(p.get("neighbor_count", 1))
        if direction == "in" and (et + "_reverse") in self.store.edges:
            # an undirected heterogeneous edge type stores its mirrored rows as '<type>_reverse': the reference defines
            # inV / inE(type) as outV / outE(type + '_reverse') (gsl/dag_node.py:470-492), results carry that type name
            et, direction = et + "_reverse", "out"
        if direction == "in":
            csr = self.store.reverse_csr(et)
            dst_t = self.store.edges[et].src_type
        else:
            csr = self.store.edges[et]
            dst_t = csr.dst_type
        if strategy == "in_degree" and direction == "out":
            self.store.ensure_indegree_weights(et)
        src_v = up.vids.reshape(-1)
        self._salt += 1
        want_edges = p.get("emit") == "edges"
        if strategy == "full":
            cap = k if k > 0 else 0
            vals, eids, offs = S.sample_full(csr, src_v, cap=cap, want_eids=True)
            counts = (offs[1:] - offs[:-1])
            ids = self.g.to_ids(dst_t, vals)
            B = int(src_v.numel())
            maxd = int(counts.max().item()) if B > 0 else 0
            out = _Out(ids=ids, vids=vals, shape=(int(ids.numel()),))
            if want_edges:
                src_ids = torch.repeat_interleave(up.ids.reshape(-1), counts)
                out.value = V_.SparseEdges(src_ids, up.value.type if hasattr(up.value, "type") else None, ids, dst_t, et,
                                           counts, (B, maxd), edge_ids=eids, graph=self.g)
            else:
                out.value = V_.SparseNodes(ids, counts, (B, maxd), dst_t, graph=self.g, vids=vals)
            return out
        fmode, fvals = S.FILTER_NONE, None
        if node._filter is not None:
            f = results[id(node._filter)]
            fv = f.vids.reshape(-1)
            if fv.numel() != src_v.numel():
                fv = fv.reshape(-1, 1).expand(-1, src_v.numel() // max(fv.numel(), 1)).reshape(-1)
            fmode, fvals = S.FILTER_ID, fv
        elif csr.timestamped and self._root_ts is not None and p.get("temporal", True):
            # temporal roots constrain EVERY downstream traversal to edges before the root element's
            # timestamp (dag_node.py:357-364,387-392); values are expanded by fan-out to align with src ids
            ts = self._root_ts.reshape(-1).to(torch.int64)
            n = int(src_v.numel())
            if ts.numel() > 0 and n % ts.numel() == 0:
                fmode, fvals = S.FILTER_TS, ts.reshape(-1, 1).expand(-1, n // ts.numel()).reshape(-1)
        nbr, eid = S.sample_neighbors(csr, src_v, k, strategy, fmode, fvals, want_eids=True, rng=self.rng,
                                      salt=self._salt)
        B = int(src_v.numel())
        ids = self.g.to_ids(dst_t, nbr)
        out = _Out(ids=ids, vids=nbr, shape=(B, k))
        if want_edges:
            src_ids = up.ids.reshape(-1, 1).expand(B, k)
            out.src_ids, out.src_vids, out.eids = src_ids, src_v, eid
            topo = self.g.get_topology()
            st = topo.get_src_type(et) if direction == "out" else topo.get_dst_type(et)
            # attribute rows live on the owner of the edge's ORIGINAL source: for in-edges that is the
            # sampled neighbour (eid = position in the forward CSR there)
            owner_v = nbr if direction == "in" else src_v.reshape(-1, 1).expand(B, k)
            out.value = V_.Edges(src_ids, st, ids, dst_t, et, eid, shape=(B, k), graph=self.g, src_vids=owner_v)
        else:
            out.value = V_.Nodes(ids, dst_t, shape=(B, k), graph=self.g, vids=nbr)
        return out

    def _negative(self, node: DagNode, up: _Out, results) -> _Out:
        p = node.params
        k = int(p.get("neighbor_count", 1))
        strategy = p.get("strategy", "random")
        self._salt += 1
        src_v = up.vids.reshape(-1)
        B = int(src_v.numel())
        gen = self.rng.torch_generator(self._salt)
        if p.get("conditional"):
            dst = results[id(p["dst_node"])]
            neg = NEG.conditional_negative(self.store, p["edge_type"], src_v, dst.vids.reshape(-1), k, strategy,
                                           p["condition"], gen)
            dst_t = self.store.edges[p["edge_type"]].dst_type
        elif "node_type" in p and "edge_type" not in p:
            dst_t = p["node_type"]
            neg = NEG.node_weight_negative(self.store, dst_t, src_v, k, gen)
        else:
            et = p["edge_type"]
            dst_t = self.store.edges[et].dst_type if p.get("direction", "out") == "out" else self.store.edges[et].src_type
            neg = NEG.edge_negative(self.store, et, src_v, k, strategy, gen, direction=p.get("direction", "out"),
                                    rng=self.rng, salt=self._salt)
        ids = self.g.to_ids(dst_t, neg)
        out = _Out(ids=ids, vids=neg, shape=(B, k))
        out.value = V_.Nodes(ids, dst_t, shape=(B, k), graph=self.g, vids=neg)
        return out

    def _walk(self, node: DagNode, up: _Out) -> _Out:
        p = node.params
        csr = self.store.edges[p["edge_type"]]
        self._salt += 1
        walks = WALK.random_walk(csr, up.vids.reshape(-1), int(p["walk_len"]), p["p"], p["q"], rng=self.rng,
                                 salt=self._salt)
        ids = self.g.to_ids(csr.dst_type, walks)
        out = _Out(ids=ids, vids=walks, shape=tuple(walks.shape))
        out.value = V_.Nodes(ids, csr.dst_type, shape=tuple(walks.shape), graph=self.g, vids=walks)
        return out

    def _subgraph(self, node: DagNode, up: Optional[_Out]) -> _Out:
        p = node.params
        et = p["nbr_type"]
        csr = self.store.edges[et]
        if up is None:
            # root SubGraph: seeds from a node iterator over the seed type
            self._ensure_root_iter()
            idx = self._iter.next_index()
            seeds = idx * self.rt.world + self.rt.rank
            src = dst = None
        elif p.get("from_edges"):
            src, dst = up.src_vids.reshape(-1), up.vids.reshape(-1)
            seeds = torch.cat([src, dst])
        else:
            seeds = up.vids.reshape(-1)
            src = dst = None
        sg = SUB.induce_subgraph(self.store, et, seeds, p.get("num_nbrs") or [], need_dist=p.get("need_dist", False),
                                 src=src, dst=dst, rng=self.rng)
        ids = self.g.to_ids(csr.src_type, sg["nodes"])
        nodes = V_.Nodes(ids, csr.src_type, graph=self.g, vids=sg["nodes"])
        edges = None
        if sg.get("eids") is not None:
            edges = V_.Edges(ids[sg["row"]], csr.src_type, ids[sg["col"]], csr.dst_type, et, sg["eids"], graph=self.g,
                             src_vids=sg["nodes"][sg["row"]])
        kw = {}
        if sg.get("dist_to_src") is not None:
            kw = {"dist_to_src": sg["dist_to_src"].cpu().numpy(), "dist_to_dst": sg["dist_to_dst"].cpu().numpy()}
        out = _Out(ids=ids, vids=sg["nodes"], shape=(int(ids.numel()),))
        out.value = V_.SubGraph(torch.stack([sg["row"], sg["col"]]), nodes, edges, **kw)
        return out

    # ------------------------------------------------------------------ run one batch
    _root_ts = None

    def run(self) -> Dict[str, object]:
        if self.sync_epoch:
            self._sync_epoch_end()
        results: Dict[int, _Out] = {}
        for node in self._order:
            up = results.get(id(node.upstream)) if node.upstream is not None else None
            if node is self.dag.root:
                if isinstance(node, SubGraphDagNode):
                    res = self._subgraph(node, None)
                elif isinstance(node, TraverseSourceEdgeDagNode):
                    res = self._root_edges(node)
                else:
                    res = self._root_nodes(node)
            elif isinstance(node, FakeNode):
                which = node.params["which"]
                if which == "src":
                    res = _Out(ids=up.src_ids, vids=(up.src_vids.reshape(-1, 1).expand(up.shape).reshape(up.shape)
                                                     if up.src_vids is not None and len(up.shape) == 2 else up.src_vids),
                               shape=up.shape)
                else:
                    res = _Out(ids=up.ids, vids=up.vids, shape=up.shape)
                res.value = V_.Nodes(res.ids, node.type, shape=res.shape, graph=self.g, vids=res.vids)
            elif isinstance(node, SubGraphDagNode):
                res = self._subgraph(node, up)
            elif node.op_name == "RandomWalk":
                res = self._walk(node, up)
            elif isinstance(node, TraverseNegVertexDagNode) or node.params.get("negative"):
                res = self._negative(node, up, results)
            elif node.op_name == "Sampler":
                res = self._sample(node, up, results)
            else:
                raise errors.UnimplementedError("unknown GSL node %r" % (node.op_name,))
            results[id(node)] = res
            if node is self.dag.root:
                self._root_ts = None
                if isinstance(node, TraverseSourceEdgeDagNode):
                    timed = self.store.edges[node.params["edge_type"]].timestamped
                elif isinstance(node, SubGraphDagNode):
                    timed = False
                else:
                    tab = self.store.nodes.get(node.type)
                    timed = tab is not None and tab.timestamps is not None and node.params.get("node_from", NODE) == NODE
                if timed and hasattr(res.value, "tensor"):
                    self._root_ts = res.value.tensor("timestamps")
        self.rng.advance(1)
        out = {}
        for alias in self.dag.list_alias():
            out[alias] = results[id(self.dag.get_node(alias))].value
        return out

    # ------------------------------------------------------------------ state
    @property
    def epoch(self):
        return self._iter.epoch if self._iter is not None else 0

    def state_dict(self):
        return {"iter": None if self._iter is None else self._iter.state_dict(), "rng": self.rng.state_dict()}

    def load_state_dict(self, sd):
        if sd.get("iter") is not None:
            self._ensure_root_iter()
            self._iter.load_state_dict(sd["iter"])
        self.rng.load_state_dict(sd["rng"])
