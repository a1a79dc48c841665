"""``nn.Dataset``: turn GSL query results into model inputs (graphlearn/python/nn/dataset.py:82-181,
nn/pytorch/data/dataset.py:31-98).  ``TorchDataset`` is the ``torch.utils.data.IterableDataset``
adapter; because batches are born on the GPU there are no DataLoader worker processes (the
reference turns every worker into a sampling client, pyg_dataloader.py:42-117)."""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch

from .. import errors
from ..gsl.dataset import Dataset as _GslDataset
from .data import BatchGraph, Data, EgoGraph, TemporalGraph


class Dataset(object):
    def __init__(self, query, window=10, induce_func: Optional[Callable] = None, batch_size: int = 1, drop_last: bool = False,
                 **kwargs):
        """``batch_size`` / ``drop_last`` only matter for ``get_subgraphs_v2`` (how many SubGraph samples form one batch,
        nn/dataset.py:44-58 of the reference); GSL batches are sized by the query's ``.batch(n)``."""
        self._query = query
        self._ds = _GslDataset(query, window=window, **kwargs)
        self._induce = induce_func
        self.batch_size, self.drop_last = int(batch_size), bool(drop_last)

    def __iter__(self):
        """one epoch of ``get_data_dict()`` results"""
        while True:
            try:
                yield self.get_data_dict()
            except errors.OutOfRangeError:
                return

    @property
    def iterator(self):
        return iter(self)

    def get_subgraphs(self, inducer):
        """``inducer.induce_func(query result) -> (positive subgraphs, negative subgraphs | None)`` for the next GSL batch
        (nn/dataset.py:124-136; ``nn.SubGraphInducer``)"""
        return inducer.induce_func(self._ds.next())

    def get_subgraphs_v2(self, processor):
        """``batch_size`` processed SubGraph samples of a ``g.SubGraph(...)`` query (nn/dataset.py:138-157;
        ``nn.SubGraphProcessor``); raises OutOfRangeError at the end of the epoch - a short tail is returned unless
        ``drop_last``"""
        if getattr(self, "_sg_epoch_end", False):          # the previous call returned the short tail of the epoch
            self._sg_epoch_end = False
            raise errors.OutOfRangeError("OutOfRange")
        rets = []
        alias = self._query.list_alias()[-1]
        while len(rets) < self.batch_size:
            try:
                rets.append(processor.process_func(self._ds.next()[alias]))
            except errors.OutOfRangeError:
                if rets and not self.drop_last:
                    self._sg_epoch_end = True
                    return rets
                raise
        return rets

    @property
    def raw(self):
        return self._ds

    def next(self):
        return self._ds.next()

    def get_data_dict(self) -> dict:
        """alias -> Data (flattened device tensors) of the next batch."""
        res = self._ds.next()
        return {k: Data.from_values(v) for k, v in res.items() if hasattr(v, "_t")}

    def _default_neighbors(self, source: str) -> List[str]:
        """hop aliases below ``source`` when every hop has exactly ONE positive downstream (the reference's
        ``get_egograph(source, neighbors=None)``, nn/tf/data/dataset.py:95-150); edge hops (outE) are skipped - their vertex
        ends (inV / outV) are the neighbours"""
        from ..gsl.dag_node import TraverseEdgeDagNode
        pre = self._query.get_node(source)
        out: List[str] = []
        recepts = pre.pos_downstreams
        while recepts:
            if len(recepts) > 1:
                raise ValueError("Can't automatically find neighbors for {}, which has multiple downstreams. You should assign "
                                 "specific neighbors for {}.".format(pre.get_alias(), source))
            cur = recepts[0]
            if not isinstance(cur, TraverseEdgeDagNode):
                out.append(cur.get_alias())
            pre, recepts = cur, cur.pos_downstreams
        return out

    def get_egograph(self, source: str, neighbors: Optional[Sequence[str]] = None, nbr_nums: Optional[Sequence[int]] = None,
                     res=None) -> EgoGraph:
        """EgoGraph rooted at alias `source` with hop aliases `neighbors` (in hop order; None = follow the single positive
        downstream chain of `source` in the query)."""
        if neighbors is None:
            neighbors = self._default_neighbors(source)
        res = res if res is not None else self._ds.next()
        src = Data.from_values(res[source])
        hops = [Data.from_values(res[a]) for a in neighbors]
        if nbr_nums is None:
            nbr_nums = [int(res[a].shape[-1]) for a in neighbors]
        return EgoGraph(src, hops, nbr_nums=nbr_nums)

    def get_temporalgraph(self, source: str, nbr_edges: Sequence[str], nbr_nodes: Sequence[str], res=None,
                          time_dim: int = 16) -> TemporalGraph:
        """TemporalGraph rooted at alias `source` (a timestamped node or edge root); hop i is described by the edge
        alias ``nbr_edges[i]`` (timestamps / edge features) and the node alias ``nbr_nodes[i]``."""
        res = res if res is not None else self._ds.next()
        src = Data.from_values(res[source])
        src_t = res[source].tensor("timestamps").reshape(-1)
        nodes = [Data.from_values(res[a]) for a in nbr_nodes]
        edges = [Data.from_values(res[a]) for a in nbr_edges]
        nbr_t = [res[a].tensor("timestamps").reshape(-1) for a in nbr_edges]
        nums = [int(res[a].shape[-1]) for a in nbr_edges]
        return TemporalGraph(src, src_t, nodes, nbr_t, edges, nums, time_dim=time_dim)

    def get_batchgraph(self, alias: str, additional_keys=()) -> BatchGraph:
        res = self._ds.next()
        g = res[alias]
        graphs = g if isinstance(g, (list, tuple)) else [g]
        if self._induce is not None:
            graphs = self._induce(res)
        return BatchGraph.from_graphs(graphs, additional_keys)

    def state_dict(self):
        return self._ds.state_dict()

    def load_state_dict(self, sd):
        self._ds.load_state_dict(sd)


class TorchDataset(torch.utils.data.IterableDataset):
    """Iterate one epoch of a query; every item is a dict alias -> Data (or the output of
    ``transform``).  Use with ``DataLoader(ds, batch_size=None)`` or iterate directly."""

    def __init__(self, query, window=10, transform: Optional[Callable] = None, length: Optional[int] = None,
                 induce_func: Optional[Callable] = None, graph=None, cluster=None):
        """``induce_func`` is the reference's name for ``transform`` (nn/pytorch/data/dataset.py:31-60: data dict -> list of
        sub-graphs).  ``graph`` / ``cluster``: lazy client-mode initialisation - the dataset connects ``graph`` to the servers
        of ``cluster`` (default: ``nn.get_cluster_spec()``) on first iteration, as DataLoader workers do in the reference."""
        super().__init__()
        self._query, self._window = query, window
        self._graph, self._cluster = graph, cluster
        self._lazy = graph is not None
        if self._lazy:
            from .utils import is_server_launched
            if not is_server_launched():
                raise RuntimeError("graph learn server should be running firstly when using lazy init dataset")
        self._nn = None if self._lazy else Dataset(query, window=window)
        self._transform = transform if transform is not None else induce_func
        self._length = length
        self._as_dict = False
        self._client_id = 0

    def lazy_init(self) -> bool:
        return self._lazy

    @property
    def client_id(self) -> int:
        return self._client_id

    @client_id.setter
    def client_id(self, value):
        if not isinstance(value, int):
            raise ValueError("client_id must be an int")
        self._client_id = value

    def as_dict(self):
        """yield plain dicts of tensors instead of ``Data`` objects (for torch's default collate; dataset.py:86-93)"""
        self._as_dict = True
        return self

    def _ensure(self):
        if self._nn is None:
            from .utils import get_cluster_spec
            self._graph.init(cluster=self._cluster or get_cluster_spec(), job_name="client", task_index=self._client_id)
            q = self._query(self._graph) if callable(self._query) else self._query
            self._nn = Dataset(q, window=self._window)

    def __iter__(self):
        self._ensure()
        n = 0
        while self._length is None or n < self._length:
            try:
                d = self._nn.get_data_dict()
            except errors.OutOfRangeError:
                return
            n += 1
            if self._transform:
                yield self._transform(d)
            elif self._as_dict:
                names = {"ints": "int_attrs", "floats": "float_attrs", "strings": "string_attrs"}      # the reference's field names
                yield {k: {names.get(a, a): b for a, b in v.__dict__.items() if isinstance(b, torch.Tensor)} for k, v in d.items()}
            else:
                yield d


class SubGraphData(object):
    """Minimal stand-in for a PyG ``Data`` (x, edge_index, y [, extra]) - PyG is not part of this
    image; models in ``graphlearn_b200.models`` consume (x, edge_index) directly."""

    def __init__(self, x, edge_index, y=None, **kw):
        self.x, self.edge_index, self.y = x, edge_index, y
        for k, v in kw.items():
            setattr(self, k, v)

    @property
    def num_nodes(self):
        return int(self.x.size(0))

    def to(self, device):
        for k, v in list(self.__dict__.items()):
            if isinstance(v, torch.Tensor):
                setattr(self, k, v.to(device))
        return self


class Batch(SubGraphData):
    """Concatenation of SubGraphData items with node offsets (``Batch.from_data_list``)."""

    @staticmethod
    def from_data_list(items: Sequence[SubGraphData]) -> "Batch":
        off, xs, eis, ys, batch = 0, [], [], [], []
        for i, d in enumerate(items):
            xs.append(d.x); eis.append(d.edge_index + off)
            if d.y is not None:
                ys.append(d.y.reshape(-1))
            batch.append(torch.full((d.num_nodes,), i, dtype=torch.long, device=d.x.device))
            off += d.num_nodes
        return Batch(torch.cat(xs), torch.cat(eis, 1), torch.cat(ys) if ys else None, batch=torch.cat(batch),
                     num_graphs=len(items))


class PyGDataLoader(object):
    """Loader over a ``TorchDataset`` whose ``transform`` (the reference's ``induce_func``) returns a
    list of subgraphs per GSL batch; yields one collated ``Batch`` per GSL batch
    (graphlearn/python/nn/pytorch/data/pyg_dataloader.py:42-117: batch_size is forced to 1 there because
    the GSL query already batches; ``length`` fixes the number of iterations per epoch so that all
    DDP ranks take the same number of steps).  No worker processes: batches are born on the GPU."""

    def __init__(self, dataset: TorchDataset, length: Optional[int] = None, collate_fn: Optional[Callable] = None, **_ignored):
        self.dataset, self.length = dataset, length
        self.collate_fn = collate_fn or Batch.from_data_list

    def __iter__(self):
        n = 0
        for item in self.dataset:
            if self.length is not None and n >= self.length:
                return
            n += 1
            yield self.collate_fn(item) if isinstance(item, (list, tuple)) else item

    def __len__(self):
        return self.length if self.length is not None else 0


class Collater(object):
    """Collate the item list of one GSL batch (pyg_dataloader.py:47-68): sub-graph items become one ``Batch``, tensors are
    stacked."""

    def collate(self, batch):
        if isinstance(batch, (list, tuple)) and len(batch) == 1 and isinstance(batch[0], (list, tuple)):
            batch = batch[0]                                      # DataLoader(batch_size=1) wraps the GSL batch once more
        elem = batch[0]
        if isinstance(elem, SubGraphData):
            return Batch.from_data_list(batch)
        if isinstance(elem, torch.Tensor):
            return torch.stack(list(batch))
        raise TypeError("PyGDataLoader found invalid type: {}".format(type(elem)))

    def __call__(self, batch):
        return self.collate(batch)


def worker_init_fn(worker_id):
    """DataLoader worker hook of the reference (pyg_dataloader.py:42-45: each worker becomes GL client
    ``worker_id + rank * num_client``).  Batches are sampled on the GPU of the owning process here, so worker processes are
    not used; the hook only records the id for scripts that read it."""
    from .utils import get_num_client, get_rank
    info = torch.utils.data.get_worker_info()
    if info is not None:
        info.dataset.client_id = worker_id + get_rank() * get_num_client()


class TemporalData(object):
    """A batch of timestamped events: ``src, dst, t, msg`` (the PyG ``TemporalData`` the reference's TGN example consumes)."""

    def __init__(self, src, dst, t, msg=None, **kw):
        self.src, self.dst, self.t, self.msg = src, dst, t, msg
        for k, v in kw.items():
            setattr(self, k, v)

    @property
    def num_events(self):
        return int(self.src.numel())

    def to(self, device):
        for k, v in list(self.__dict__.items()):
            if isinstance(v, torch.Tensor):
                setattr(self, k, v.to(device))
        return self


class TemporalDataset(TorchDataset):
    """Edge-rooted query -> stream of ``TemporalData`` (temporal_dataset.py:31-47): the alias ``event_name`` must name a
    timestamped ``E(...)`` traversal; ``msg`` carries the edge weights (or float attributes when the edges have them)."""

    def __init__(self, query, window=10, event_name="event", length: Optional[int] = None):
        def induce(data_dict_raw):
            if event_name not in data_dict_raw:
                raise ValueError("Event name {} not exist.".format(event_name))
            ev = data_dict_raw[event_name]
            msg = ev.tensor("float_attrs")
            if not isinstance(msg, torch.Tensor):
                msg = ev.tensor("weights")
            return TemporalData(ev.tensor("src_ids").reshape(-1), ev.tensor("dst_ids").reshape(-1),
                                ev.tensor("timestamps").reshape(-1), msg if isinstance(msg, torch.Tensor) else None)
        super().__init__(query, window=window, length=length)
        self._induce = induce

    def __iter__(self):
        n = 0
        while self._length is None or n < self._length:
            try:
                res = self._nn.next()
            except errors.OutOfRangeError:
                return
            n += 1
            yield self._induce(res)


class TemporalDataLoader(object):
    """Iterates a ``TemporalDataset`` as is (temporal_dataloader.py:31-45: batch size, shuffling and collation are fixed by
    the GSL query, the loader only forwards)."""

    def __init__(self, dataset: TemporalDataset, **_ignored):
        assert isinstance(dataset, TemporalDataset), "TemporalDataLoader only accepts GraphLearn TemporalDataset"
        self.dataset = dataset

    def __iter__(self):
        return iter(self.dataset)
