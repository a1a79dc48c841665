"""``gl.Dataset``: iterate a GSL query batch by batch.

API parity with graphlearn/python/gsl/dag_dataset.py:29-155: ``next()`` returns a
dict-like object indexed by alias and raises ``OutOfRangeError`` once per epoch.

The reference keeps `window` batches in flight between sampler servers and the
trainer (TapeStore + client prefetch threads).  Here sampling is a handful of
device kernels on a side CUDA stream; ``window`` bounds how many batches are
produced ahead of the consumer (device-side ring of ready batches, R6/R9).
"""
from __future__ import annotations

import collections

import torch

from .. import errors
from .compile import CompiledQuery, compilable, compile_query
from .executor import QueryExecutor


class DagValues(dict):
    """alias -> Nodes / Edges / SubGraph of one batch."""

    def process(self):
        return self


class Dataset(object):
    def __new__(cls, dag, *args, **kwargs):
        if getattr(dag.graph, "remote", False):             # server-mode client: the query runs on the graph servers
            from ..service import RemoteDataset
            return RemoteDataset(dag.graph, dag, kwargs.get("window", args[0] if args else 10))
        return super().__new__(cls)

    def __init__(self, dag, window=10, keep_alive_rounds=1, drop_last=False, prefetch=True, sync_epoch=None):
        """``sync_epoch``: end the epoch on every rank as soon as one rank runs out of seeds (None = automatic: on
        for multi-rank runs whose sampling ops are collectives, i.e. the portable path)."""
        self._dag = dag
        self._graph = dag.graph
        self._window = max(1, int(window))
        # static multi-hop chains run as a captured CUDA graph over a ring of `window` pre-allocated batches
        # (gsl/compile.py); everything else is interpreted node by node (gsl/executor.py)
        self.plan = compile_query(dag)
        self._compiled = None
        if prefetch and compilable(self._graph, self.plan):
            self._compiled = CompiledQuery(self._graph, self.plan, depth=min(max(self._window, 2), 8), drop_last=drop_last)
        self._exec = None if self._compiled is not None else QueryExecutor(dag, drop_last=drop_last, sync_epoch=sync_epoch)
        self._ring = collections.deque()
        self._pending_eoe = False
        self._prefetch = bool(prefetch) and self._graph.runtime.is_cuda
        self._stream = torch.cuda.Stream() if self._prefetch else None
        self._graph.add_dataset(self)
        self._closed = False

    def _produce_one(self):
        """Produce one batch (on the sampling stream) and push (values, event) to the ring."""
        if self._pending_eoe:
            return False
        if self._compiled is not None:
            try:
                slot, m = self._compiled.launch()
            except errors.OutOfRangeError:
                self._pending_eoe = True
                return False
            self._ring.append(((slot, m), None))
            return True
        try:
            if self._stream is not None:
                self._stream.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(self._stream):
                    vals = self._exec.run()
                    ev = torch.cuda.Event()
                    ev.record(self._stream)
            else:
                vals, ev = self._exec.run(), None
        except errors.OutOfRangeError:
            self._pending_eoe = True
            return False
        self._ring.append((vals, ev))
        return True

    def next(self):
        if self._closed:
            raise errors.FailedPreconditionError("dataset is closed")
        if not self._ring:
            self._produce_one()
        if not self._ring:
            self._pending_eoe = False          # epoch boundary consumed; next call starts a new epoch
            raise errors.OutOfRangeError("out of range")
        vals, ev = self._ring.popleft()
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)
        # keep the pipeline `window` deep (sampling of later batches overlaps the consumer's compute); the compiled
        # ring re-uses slot buffers, so one slot stays reserved for the batch the consumer is holding
        ahead = (self._compiled.depth - 2) if self._compiled is not None else (self._window - 1)
        while len(self._ring) < ahead and self._produce_one():
            pass
        if self._compiled is not None:
            vals = self._compiled.values(*vals)
        res = DagValues(vals)
        f = self._dag.value_func
        return f(res) if f is not None else res

    __next__ = next

    def __iter__(self):
        return self

    @property
    def compiled(self) -> bool:
        """True when the query runs as a captured CUDA graph (static sampling plan)."""
        return self._compiled is not None

    @property
    def epoch(self):
        return (self._compiled or self._exec).epoch

    def state_dict(self):
        return (self._compiled or self._exec).state_dict()

    def load_state_dict(self, sd):
        self._ring.clear()
        self._pending_eoe = False
        (self._compiled or self._exec).load_state_dict(sd)

    def close(self):
        self._closed = True
        self._ring.clear()
