"""GSL query -> static sampling plan -> captured CUDA graph.

The reference compiles a GSL chain once into a DagDef, ships it to the servers and executes it
continuously into a bounded queue of tapes (graphlearn/python/gsl/dag_node.py:164-305,
gsl/dag_dataset.py:29-97, src/core/runner/dag_scheduler.cc:45-86, src/core/dag/tape.cc:110-153).
Here the same chain

    g.V(t).batch(B).shuffle(traverse=True).alias(a0).outV(e).sample(k1).by(s1).alias(a1)...values()

is lowered to a :class:`SamplingPlan` (root type, batch size, traversal mode, hop list with edge
type / direction / fan-out / strategy) and executed by :class:`CompiledQuery`:

* seed traversal (GetNodes) stays on the host - a cursor over a per-epoch permutation of the rank's
  own rows - and every batch is staged through pinned memory;
* the hop chain is ONE captured CUDA graph per ring slot (seed H2D copy -> K1 sampling kernel per hop
  writing straight into the slot's pre-allocated hop buffers -> RNG advance): no Python per hop, no
  allocation, no host sync (the Tape / TapeStore ring, R5-R7);
* ``FastSageTrainer.from_query`` consumes the same plan: the hop chain becomes the sampling branch
  of the training-step graph.

Queries that do not fit (negatives, filters, full/sparse sampling, edge traversals, sub-graphs,
non-dense id spaces) keep using the interpreter in ``gsl/executor.py``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import torch

from .. import config as _config
from .. import errors
from ..data import values as V_
from ..ops import rng as rng_ops
from ..ops import sampling as S
from .dag_node import (DagNode, FakeNode, SubGraphDagNode, TraverseEdgeDagNode, TraverseNegVertexDagNode,
                       TraverseSourceEdgeDagNode, TraverseVertexDagNode)
from .iterators import SeedIterator

NODE = 0
_KERNEL_STRATEGIES = ("random", "random_without_replacement", "topk", "edge_weight", "in_degree")


@dataclass
class HopSpec:
    edge_type: str
    direction: str          # 'out' | 'in'
    k: int
    strategy: str
    alias: Optional[str]
    dst_type: str


@dataclass
class SamplingPlan:
    root_type: str           # (possibly masked) node type the seeds are traversed from
    base_type: str           # unmasked type: id space of the edges / attribute table
    root_alias: Optional[str]
    batch_size: int
    traverse: str            # 'by_order' | 'shuffle' | 'random'
    hops: List[HopSpec] = field(default_factory=list)

    @property
    def fanouts(self):
        return [h.k for h in self.hops]

    def describe(self) -> str:
        s = "V(%s).batch(%d).%s" % (self.root_type, self.batch_size, self.traverse)
        for h in self.hops:
            s += ".%sV(%s).sample(%d).by(%s)" % (h.direction, h.edge_type, h.k, h.strategy)
        return s


def compile_query(dag) -> Optional[SamplingPlan]:
    """Lower a ready GSL query to a static plan, or return None when it needs the interpreter."""
    if not dag.is_ready():
        raise ValueError("query is not ready: end it with .values()")
    root: DagNode = dag.root
    if root is None or isinstance(root, (SubGraphDagNode, TraverseSourceEdgeDagNode, TraverseEdgeDagNode)):
        return None
    p = root.params
    if p.get("node_from", NODE) != NODE or "batch_size" not in p:
        return None
    plan = SamplingPlan(root_type=root.type, base_type=root._base_type or root.type, root_alias=root.get_alias(),
                        batch_size=int(p["batch_size"]), traverse=p.get("strategy", "by_order"))
    node = root
    while True:
        downs = node.downstreams
        if not downs:
            break
        if len(downs) != 1:
            return None                                    # branching queries: interpreter
        d = downs[0]
        if isinstance(d, (TraverseNegVertexDagNode, TraverseEdgeDagNode, SubGraphDagNode, FakeNode)):
            return None
        if not isinstance(d, TraverseVertexDagNode) or d.op_name != "Sampler":
            return None
        dp = d.params
        if d._filter is not None or dp.get("emit") == "edges":
            return None
        strategy = dp.get("strategy", "random")
        if strategy not in _KERNEL_STRATEGIES or "neighbor_count" not in dp:
            return None
        plan.hops.append(HopSpec(dp["edge_type"], dp.get("direction", "out"), int(dp["neighbor_count"]), strategy,
                                 d.get_alias(), d.type))
        node = d
    return plan


class CompiledQuery(object):
    """Executes a :class:`SamplingPlan` as a captured CUDA graph over a ring of pre-allocated hop buffers."""

    def __init__(self, graph, plan: SamplingPlan, depth: int = 2, drop_last: bool = False, seed: Optional[int] = None):
        self.g, self.plan = graph, plan
        self.store, self.rt = graph.store, graph.runtime
        rt, dev = self.rt, self.rt.device
        assert rt.is_cuda and _config.get().use_peer_kernels
        self.B = plan.batch_size
        self.depth = max(1, min(int(depth), 8))
        self.rng = rng_ops.DeviceRng(rt, _config.get().seed if seed is None else seed)
        tab = self.store.nodes[plan.root_type]
        base = self.store.nodes[plan.base_type]
        rows = tab.present.nonzero().flatten() if tab.present is not None else torch.arange(tab.n_local, device=dev)
        vids = rows * rt.world + rt.rank
        if plan.base_type != plan.root_type:
            vids = base.idmap.to_vid(tab.idmap.to_id(vids))
        self.seed_vids = vids.cpu()                         # host copy: the traversal runs on the host
        self.iter = SeedIterator(int(vids.numel()), self.B, plan.traverse, "cpu", seed=_config.get().seed + 17 * rt.rank,
                                 drop_last=drop_last)
        self.csrs = []
        for h in plan.hops:
            if h.direction == "in":
                csr = self.store.reverse_csr(h.edge_type)
            else:
                csr = self.store.edges[h.edge_type]
                if h.strategy == "in_degree":
                    self.store.ensure_indegree_weights(h.edge_type)
            self.csrs.append(csr)
        # ring of batches: pinned seed staging + device hop buffers + one graph per slot
        n = [self.B]
        for h in plan.hops:
            n.append(n[-1] * h.k)
        self.n = n
        self.h_seeds = [torch.full((self.B,), -1, dtype=torch.int64).pin_memory() for _ in range(self.depth)]
        self.hops = [[torch.zeros(m, dtype=torch.int64, device=dev) for m in n] for _ in range(self.depth)]
        self.events = [torch.cuda.Event() for _ in range(self.depth)]
        self.graphs: List[Optional[torch.cuda.CUDAGraph]] = [None] * self.depth
        self._stream = torch.cuda.Stream()
        self._slot = 0
        self._capture()

    def _run_chain(self, slot: int):
        hops = self.hops[slot]
        hops[0].copy_(self.h_seeds[slot], non_blocking=True)
        cur = hops[0]
        for i, (h, csr) in enumerate(zip(self.plan.hops, self.csrs)):
            S.sample_neighbors(csr, cur, h.k, h.strategy, want_eids=False, rng=self.rng, salt=i + 1, out=hops[i + 1])
            cur = hops[i + 1]
        self.rng.advance()

    def _capture(self):
        s = self._stream
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self._run_chain(0)                              # warm-up (also validates the strategies)
        s.synchronize()
        pool = None
        for i in range(self.depth):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool, stream=s):
                self._run_chain(i)
            pool = g.pool() if pool is None else pool
            self.graphs[i] = g
        torch.cuda.synchronize()

    # ------------------------------------------------------------------ iteration
    def next_seeds(self) -> torch.Tensor:
        """host vids of the next seed batch (raises OutOfRangeError at the end of an epoch)"""
        if _config.get().actor_enabled and self.rt.world > 1 and self.plan.traverse in ("shuffle", "by_order"):
            return self._next_dispatched()
        idx = self.iter.next_index()
        return self.seed_vids[idx]

    def _next_dispatched(self) -> torch.Tensor:
        """``gl.enable_actor()``: at every epoch start the ranks pool their seed batches and the balanced plan of
        engine/dispatch.py gives each rank the same number of (work-balanced) batches - one collective per epoch."""
        from ..engine.dispatch import BalancedSeedDispatcher
        if getattr(self, "_planned", None) is None:
            vids = self.seed_vids
            if self.plan.traverse == "shuffle":
                g = torch.Generator(device="cpu").manual_seed(_config.get().seed * 7919 + 31 * self.rt.rank + self.iter.epoch)
                vids = vids[torch.randperm(int(vids.numel()), generator=g)]
            dv = vids.to(self.rt.device)
            weights = None
            tab = self.store.nodes[self.plan.base_type]
            et = self.plan.hops[0].edge_type
            if self.plan.hops[0].direction == "out" and et in tab.out_degrees:
                weights = tab.out_degrees[et][dv // self.rt.world].float() + 1.0
            self._planned = BalancedSeedDispatcher(self.rt, self.B).plan(dv, weights).cpu()
            self._plan_iter = SeedIterator(int(self._planned.numel()), self.B, "by_order", "cpu")
        try:
            idx = self._plan_iter.next_index()
        except errors.OutOfRangeError:
            self._planned = None
            self.iter.epoch += 1
            raise
        return self._planned[idx]

    def launch(self):
        """Stage the next seed batch and replay the hop-chain graph of the next ring slot.  Returns (slot, n_real)."""
        seeds = self.next_seeds()
        slot = self._slot
        self._slot = (slot + 1) % self.depth
        self.events[slot].synchronize()                     # the slot's previous batch has been produced
        hs = self.h_seeds[slot]
        m = int(seeds.numel())
        hs[:m].copy_(seeds)
        if m < self.B:
            hs[m:].fill_(-1)                                # invalid ids: the samplers emit the default neighbour
        self._stream.wait_stream(torch.cuda.current_stream())   # consumers of this slot's previous batch are enqueued
        with torch.cuda.stream(self._stream):
            self.graphs[slot].replay()
            self.events[slot].record(self._stream)
        return slot, m

    def values(self, slot: int, m: int):
        """alias -> Nodes over the slot's buffers (consumer stream waits for the producer event)"""
        torch.cuda.current_stream().wait_event(self.events[slot])
        out = {}
        hops = self.hops[slot]
        plan = self.plan
        if plan.root_alias:
            ids = hops[0][:m]
            out[plan.root_alias] = V_.Nodes(ids, plan.base_type, shape=(m,), graph=self.g, vids=ids)
        rows = m
        for i, h in enumerate(plan.hops):
            if h.alias:
                ids = hops[i + 1][:rows * h.k].view(rows, h.k)
                out[h.alias] = V_.Nodes(ids, h.dst_type, shape=(rows, h.k), graph=self.g, vids=ids)
            rows *= h.k
        return out

    @property
    def epoch(self):
        return self.iter.epoch

    def state_dict(self):
        return {"iter": self.iter.state_dict(), "rng": self.rng.state_dict()}

    def load_state_dict(self, sd):
        self.iter.load_state_dict(sd["iter"])
        self.rng.load_state_dict(sd["rng"])


def compilable(graph, plan: Optional[SamplingPlan]) -> bool:
    """The compiled path needs the CUDA peer kernels and dense id spaces (vid == id, no collective translation)."""
    if plan is None or not plan.hops:
        return False
    rt = graph.runtime
    if rt is None or not rt.is_cuda or not _config.get().use_peer_kernels:
        return False
    store = graph.store
    types = {plan.root_type, plan.base_type} | {h.dst_type for h in plan.hops}
    for t in types:
        if t not in store.nodes or not store.nodes[t].idmap.dense:
            return False
    return all(h.edge_type in store.edges for h in plan.hops)
