"""Process-level helpers of the reference's PyTorch data layer (graphlearn/python/nn/pytorch/data/utils.py:31-150) and the
TF trainers' end-of-training barrier (nn/tf/utils/sync_barrier.py:28-75).

In the reference every DDP trainer starts a GL *server* process next to itself (``launch_server``), the trainers all-reduce
their servers' ``ip:port`` (``bootstrap``) into the cluster spec, and DataLoader workers become GL clients.  Here sampling is a
device kernel inside each rank, so none of this is needed for worker mode - but scripts written against the reference call
these functions, and server mode (``service/``) is a real deployment: ``launch_server`` starts a ``GraphServer`` for this rank
and ``get_cluster_spec`` carries real addresses.
"""
from __future__ import annotations

import os
import socket
import time
import warnings
from typing import Optional

import torch
import torch.distributed as dist

SERVER_LAUNCHED = False
CLUSTER_SPEC = None
WORLD_SIZE = None
RANK = None
NUM_CLIENT = None
STATS_DICT = []


def get_world_size() -> int:
    global WORLD_SIZE
    if WORLD_SIZE is None:
        WORLD_SIZE = dist.get_world_size() if dist.is_available() and dist.is_initialized() else int(os.getenv("WORLD_SIZE", 1))
    return WORLD_SIZE


def get_rank() -> int:
    global RANK
    if RANK is None:
        RANK = dist.get_rank() if dist.is_available() and dist.is_initialized() else int(os.getenv("RANK", 0))
    return RANK


def get_num_client() -> int:
    global NUM_CLIENT
    if NUM_CLIENT is None:
        NUM_CLIENT = int(os.getenv("GL_NUM_CLIENT", 1))
    return NUM_CLIENT


def set_client_num(n: int):
    assert isinstance(n, int), "client_num should be int, not {}".format(type(n))
    global NUM_CLIENT
    if NUM_CLIENT is not None:
        warnings.warn("graph learn client number has been configured")
    else:
        NUM_CLIENT = n


def _free_port(host: str) -> int:
    s = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
    try:
        s.bind((host, 0))
        return s.getsockname()[1]
    finally:
        s.close()


def _local_ip() -> str:
    try:
        ip = socket.gethostbyname(socket.gethostname())
    except OSError:
        ip = ""
    return ip or "127.0.0.1"         # container hostnames may not resolve


def bootstrap(world_size: int, rank: int) -> str:
    """-> ``"ip:port,ip:port,..."``: one freshly reserved server address per rank, exchanged with one all-reduce
    (utils.py:82-116)."""
    ip = _local_ip()
    port = _free_port(ip)
    if not (dist.is_available() and dist.is_initialized()):
        return "%s:%d" % (ip, port)
    t = torch.zeros(world_size, 5, dtype=torch.int32)
    t[rank] = torch.tensor([int(x) for x in ip.split(".")] + [port], dtype=torch.int32)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return ",".join("%d.%d.%d.%d:%d" % tuple(r) for r in t.cpu().tolist())


def get_cluster_spec() -> dict:
    global CLUSTER_SPEC
    if CLUSTER_SPEC is None:
        world, rank = get_world_size(), get_rank()
        CLUSTER_SPEC = {"server": bootstrap(world, rank), "client_count": world * get_num_client()}
    return CLUSTER_SPEC


def launch_server(g, cluster: Optional[dict] = None, task_index: Optional[int] = None):
    """Start this rank's graph server (utils.py:123-140).  The server lives on daemon threads of this process
    (``service.GraphServer``); its statistics are available from ``get_counts()`` right away."""
    global SERVER_LAUNCHED
    if SERVER_LAUNCHED:
        raise RuntimeError("duplicate server launch detected")
    if cluster is None:
        cluster, task_index = get_cluster_spec(), get_rank()
    elif task_index is None:
        raise UserWarning("task_index should be explicitly defined when cluster defined by user")
    g.init(cluster=cluster, job_name="server", task_index=task_index)
    SERVER_LAUNCHED = True
    STATS_DICT.append(dict(g.server_get_stats()))
    return g


def get_counts() -> dict:
    return STATS_DICT[0]


def is_server_launched() -> bool:
    return SERVER_LAUNCHED


class Config(object):
    """Model-layer switches of the reference (nn/tf/config.py): ``conf.training`` (dropout / attention dropout of the layers that
    read it - torch modules normally use ``module.train()`` / ``.eval()``, the flag is honoured by ``models.SparseGNN`` and the
    EgoDataLoader example), the embedding partitioning knobs (kept for scripts; big tables are sharded over GPUs by
    ``nn.ShardedEmbedding``) and ``emb_live_steps`` (``DynamicEmbedding`` rows never expire here)."""

    def __init__(self):
        self.training = True
        self.partitioner = "min_max"
        self.emb_max_partitions = None
        self.emb_min_slice_size = 128 * 1024
        self.emb_live_steps = None


conf = Config()


def _reset_for_tests():
    global SERVER_LAUNCHED, CLUSTER_SPEC, WORLD_SIZE, RANK, NUM_CLIENT
    SERVER_LAUNCHED, CLUSTER_SPEC, WORLD_SIZE, RANK, NUM_CLIENT = False, None, None, None, None
    del STATS_DICT[:]


class SyncBarrierHook(object):
    """End-of-training barrier (sync_barrier.py:28-75: workers enqueue a token into a shared queue on the parameter server and
    wait until all ``num_worker`` tokens are there, so no worker tears the cluster down while others still train).  Here the
    queue is a ``torch.distributed`` store counter; without a process group the hook is a no-op.  Use ``hook.end()`` after the
    training loop, or as a context manager around it."""

    _instances = 0          # hooks are created in the same order on every rank: the sequence number keys the store counter

    def __init__(self, num_worker: Optional[int] = None, is_chief: Optional[bool] = None, timeout_s: float = 3600.0,
                 key: Optional[str] = None):
        self._n = num_worker if num_worker is not None else get_world_size()
        self._chief = is_chief if is_chief is not None else get_rank() == 0
        SyncBarrierHook._instances += 1
        self._timeout, self._done = timeout_s, False
        self._key = key or "glb_sync_barrier_%d" % SyncBarrierHook._instances

    def begin(self):
        return self

    def after_create_session(self, session=None, coord=None):
        return None

    def end(self, session=None):
        if self._done:
            return
        self._done = True
        if not (dist.is_available() and dist.is_initialized()) or self._n <= 1:
            return
        store = dist.distributed_c10d._get_default_store()
        store.add(self._key, 1)
        t0 = time.time()
        while int(store.add(self._key, 0)) < self._n:
            if time.time() - t0 > self._timeout:
                raise TimeoutError("SyncBarrierHook: %d of %d workers finished" % (int(store.add(self._key, 0)), self._n))
            time.sleep(0.05)

    def __enter__(self):
        return self.begin()

    def __exit__(self, *exc):
        if exc[0] is None:
            self.end()
        return False
