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
