"""Cluster description helpers (graphlearn/python/cluster.py:26-101).

The reference derives (cluster spec, job name, task index) from ``TF_CONFIG`` or
CLI flags for its server/client processes.  On a single 8-GPU box the cluster is
the torchrun world; this helper returns the same triple from RANK / WORLD_SIZE
(or from ``TF_CONFIG`` when present) so launch scripts keep working."""
from __future__ import annotations

import json
import os


def get_cluster(cluster_spec=None, job_name=None, task_index=None):
    tf = os.environ.get("TF_CONFIG")
    if cluster_spec is None and tf:
        cfg = json.loads(tf)
        cluster = cfg.get("cluster", {})
        task = cfg.get("task", {})
        workers = cluster.get("worker", []) + cluster.get("chief", [])
        spec = {"server_count": len(cluster.get("ps", [])) or len(workers), "client_count": len(workers)}
        return spec, task.get("type", "worker"), int(task.get("index", 0))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    spec = cluster_spec or {"server_count": world, "client_count": world}
    return spec, job_name or "worker", rank if task_index is None else task_index


def get_cluster_spec(port=None, addr=None, gl_rank=None, world_size=None):
    """PyTorch-side helper of the reference (nn/pytorch/data/utils.py:64-110): every DDP rank
    all-reduces its sampler server's ip:port to build the cluster spec.  There are no sampler
    servers here - sampling is a device kernel inside each rank - so the spec only names the ranks."""
    world = int(world_size if world_size is not None else os.environ.get("WORLD_SIZE", "1"))
    host = addr or os.environ.get("MASTER_ADDR", "127.0.0.1")
    return {"server": ",".join("%s:%d" % (host, 0) for _ in range(world)), "client_count": world}


def launch_server(graph, cluster=None, task_index=0):
    """Reference: starts a GL server next to each trainer (nn/pytorch/data/utils.py:112-138).  Builds this rank's shard
    and - when ``cluster`` names server addresses ({"server": "host:port,...", "client_count": C}) - starts a
    ``service.GraphServer`` on this task's address (daemon threads; ``graph.wait_for_close()`` blocks until every client
    has stopped).  Without a cluster spec the call only initialises the graph (worker mode)."""
    return graph.init(task_index=task_index, cluster=cluster or "", job_name="server" if cluster else "")
