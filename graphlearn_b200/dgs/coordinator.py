"""Coordinator duties that survive without Kafka / k8s (D12): periodic + on-demand checkpoints with
retention (``SubServiceCheckpointManager``, dynamic_graph_service/python/coordinator/checkpoint.py:135-241) and global
barriers over the ingest stream (``GlobalBarrierMonitor``, barrier.py:85-168: a barrier is released
once every record produced BEFORE it has been applied - here: once the ingested-record counter
reaches the value captured when the barrier was set)."""
from __future__ import annotations

import json
import os
import threading
import time
from typing import Dict, Optional

import torch


class CheckpointManager(object):
    def __init__(self, service, path: str, keep: int = 3):
        self.service, self.path, self.keep = service, path, int(keep)
        os.makedirs(path, exist_ok=True)
        self._lock = threading.Lock()
        self._timer: Optional[threading.Thread] = None
        self._stop = threading.Event()

    def _list(self):
        return sorted(int(f.split(".")[1]) for f in os.listdir(self.path) if f.startswith("ckpt.") and f.endswith(".pt"))

    def save(self) -> int:
        with self._lock:
            cid = (self._list() or [0])[-1] + 1
            tmp = os.path.join(self.path, "tmp.%d" % cid)
            torch.save({"state": self.service.checkpoint(), "queries": {q: _plan_to_dict(p) for q, p in self.service.queries.items()},
                        "time": time.time()}, tmp)
            os.replace(tmp, os.path.join(self.path, "ckpt.%d.pt" % cid))     # atomic publish
            for old in self._list()[:-self.keep]:
                os.remove(os.path.join(self.path, "ckpt.%d.pt" % old))
            with open(os.path.join(self.path, "LATEST"), "w") as f:
                json.dump({"id": cid, "ingested": self.service.ingested}, f)
            return cid

    # ---- backup-engine surface (the reference keeps RocksDB BackupEngine backups per store partition and restores a chosen
    # backup id: sample_store.h StorePartitionBackupInfo / Backup() / PartitionedDB::Restore)
    def list_backups(self):
        """[{"id", "time", "bytes"}] of the retained checkpoints, oldest first"""
        out = []
        for cid in self._list():
            f = os.path.join(self.path, "ckpt.%d.pt" % cid)
            out.append({"id": cid, "time": os.path.getmtime(f), "bytes": os.path.getsize(f)})
        return out

    def purge(self, keep: Optional[int] = None) -> int:
        """delete all but the newest ``keep`` checkpoints; returns how many were removed"""
        keep = self.keep if keep is None else int(keep)
        with self._lock:
            old = self._list()[:-keep] if keep > 0 else self._list()
            for cid in old:
                os.remove(os.path.join(self.path, "ckpt.%d.pt" % cid))
            return len(old)

    def restore_latest(self) -> Optional[int]:
        ids = self._list()
        return self.restore(ids[-1]) if ids else None

    def restore(self, cid: int) -> int:
        """restore the service (stores + installed queries) from checkpoint ``cid``"""
        f = os.path.join(self.path, "ckpt.%d.pt" % int(cid))
        if not os.path.exists(f):
            raise FileNotFoundError("no checkpoint %d under %s (have %s)" % (cid, self.path, self._list()))
        ids = [int(cid)]
        ck = torch.load(f, weights_only=True)
        from .plan import PlanNode, QueryPlan
        for q, d in ck["queries"].items():
            plan = QueryPlan(d["source"])
            for n in d["nodes"]:
                plan.add(PlanNode(**n))
            self.service.install_query(int(q), plan)
        self.service.restore(ck["state"])
        return ids[-1]

    def start_periodic(self, interval_s: float):
        def loop():
            while not self._stop.wait(interval_s):
                self.save()
        self._timer = threading.Thread(target=loop, daemon=True)
        self._timer.start()

    def stop(self):
        self._stop.set()


def _plan_to_dict(plan) -> dict:
    return {"source": plan.source_type,
            "nodes": [{"nid": n.id, "kind": n.kind, "vtype": n.vtype, "etype": n.etype, "fanout": n.fanout,
                       "versions": n.versions, "parent": n.parent} for i, n in sorted(plan.nodes.items()) if i != 0]}


class BarrierMonitor(object):
    def __init__(self, service):
        self.service = service
        self._barriers: Dict[str, int] = {}

    def set(self, name: str, produced: Optional[int] = None):
        """``produced``: number of records produced so far by the data loaders (default: everything
        already applied, i.e. the barrier is immediately ready)."""
        self._barriers[name] = int(self.service.ingested if produced is None else produced)

    def status(self, name: str) -> str:
        if name not in self._barriers:
            return "NOT_SET"
        return "READY" if self.service.ingested >= self._barriers[name] else "PRODUCED"


# ------------------------------------------------------------------------------------------ worker registry (D12)
TERMINATED, REGISTERED, STARTED = 0, 1, 2


class WorkerRegistry(object):
    """Per sub-service worker state machine TERMINATED -> REGISTERED -> STARTED
    (dynamic_graph_service/python/coordinator/state_manager.py:30-89).  Sampling workers are NOT independent: a
    registration that arrives while every peer is already registered means one of them restarted, so the whole group
    is reset and has to register again (they exchange subscription rules and must agree on a checkpoint); serving
    workers are independent and re-register alone."""

    def __init__(self, name: str, num_workers: int, independent: bool):
        self.name, self.n, self.independent = name, int(num_workers), bool(independent)
        self.addr = [""] * self.n
        self.state = [TERMINATED] * self.n
        self.beat = [0.0] * self.n
        self._lock = threading.Lock()

    def register(self, wid: int, addr: str = "") -> bool:
        if not 0 <= wid < self.n:
            return False
        with self._lock:
            if not self.independent and all(s >= REGISTERED for s in self.state):
                self.addr, self.state = [""] * self.n, [TERMINATED] * self.n
            self.addr[wid], self.state[wid], self.beat[wid] = addr, REGISTERED, time.time()
        return True

    def set_started(self, wid: int):
        with self._lock:
            self.state[wid], self.beat[wid] = STARTED, time.time()

    def heartbeat(self, wid: int):
        with self._lock:
            self.beat[wid] = time.time()

    def reap(self, timeout_s: float):
        """workers that missed their heartbeat for ``timeout_s`` go back to TERMINATED; returns their ids"""
        now, dead = time.time(), []
        with self._lock:
            for i in range(self.n):
                if self.state[i] != TERMINATED and now - self.beat[i] > timeout_s:
                    self.state[i] = TERMINATED
                    dead.append(i)
        return dead

    def all_registered(self) -> bool:
        return all(s >= REGISTERED for s in self.state)

    def all_started(self) -> bool:
        return all(s >= STARTED for s in self.state)


class Coordinator(object):
    """Control plane of a :class:`~graphlearn_b200.dgs.workers.StreamingCluster`
    (python/coordinator/coordinator.py:80-137): workers register and fetch their init info (schema, installed
    query, partition counts, the checkpoint to restore from), report STARTED, and the coordinator drives consistent
    cluster checkpoints (taken at a barrier so the sampling state, the subscription tables and the serving caches
    belong to the same ingest offsets) and named barriers."""

    def __init__(self, cluster, meta_dir: Optional[str] = None, keep: int = 3):
        self.cluster = cluster
        self.sampling = WorkerRegistry("SamplingWorker", cluster.P, independent=False)
        self.serving = WorkerRegistry("ServingWorker", cluster.S, independent=True)
        self.meta_dir, self.keep = meta_dir, int(keep)
        self._barriers: Dict[str, dict] = {}
        if meta_dir:
            os.makedirs(meta_dir, exist_ok=True)

    def _group(self, kind: str) -> WorkerRegistry:
        return self.sampling if kind == "sampling" else self.serving

    def register_worker(self, kind: str, wid: int, addr: str = "") -> dict:
        if not self._group(kind).register(wid, addr):
            raise ValueError("invalid %s worker id %d" % (kind, wid))
        c = self.cluster
        return {"schema": c.schema, "query_plan": None if c.plan is None else _plan_to_dict(c.plan),
                "num_sampling": c.P, "num_serving": c.S, "restore_from": self.latest_checkpoint()}

    def report_started(self, kind: str, wid: int):
        self._group(kind).set_started(wid)

    def ready(self) -> bool:
        return self.sampling.all_started() and self.serving.all_started()

    # ---- barriers over the whole pipeline
    def set_barrier(self, name: str):
        c = self.cluster
        self._barriers[name] = {"ingest": [c.ingest.end_offset(p) for p in range(c.P)]}

    def barrier_status(self, name: str) -> str:
        b = self._barriers.get(name)
        if b is None:
            return "NOT_SET"
        c = self.cluster
        if any(w.offset < b["ingest"][w.wid] for w in c.sampling):
            return "PRODUCED"
        if any(w.offset < c.publish.end_offset(w.wid) for w in c.serving):
            return "SAMPLED"
        return "READY"

    # ---- checkpoints
    def _ids(self):
        if not self.meta_dir:
            return []
        return sorted(int(f.split(".")[1]) for f in os.listdir(self.meta_dir) if f.startswith("cluster.") and f.endswith(".pt"))

    def latest_checkpoint(self) -> Optional[int]:
        ids = self._ids()
        return ids[-1] if ids else None

    def checkpoint(self) -> int:
        assert self.meta_dir, "coordinator has no meta_dir"
        self.cluster.pump()                                     # quiesce: sampled AND published
        cid = (self._ids() or [0])[-1] + 1
        tmp = os.path.join(self.meta_dir, "tmp.%d" % cid)
        torch.save({"state": self.cluster.checkpoint(), "plan": None if self.cluster.plan is None else _plan_to_dict(self.cluster.plan)}, tmp)
        os.replace(tmp, os.path.join(self.meta_dir, "cluster.%d.pt" % cid))
        for old in self._ids()[:-self.keep]:
            os.remove(os.path.join(self.meta_dir, "cluster.%d.pt" % old))
        # everything before the checkpointed offsets can leave the logs
        for w in self.cluster.sampling:
            self.cluster.ingest.truncate(w.wid, w.offset)
        for w in self.cluster.serving:
            self.cluster.publish.truncate(w.wid, w.offset)
        return cid

    def restore_latest(self) -> Optional[int]:
        cid = self.latest_checkpoint()
        if cid is None:
            return None
        ck = torch.load(os.path.join(self.meta_dir, "cluster.%d.pt" % cid), weights_only=True)
        if ck["plan"] is not None:
            from .plan import PlanNode, QueryPlan
            plan = QueryPlan(ck["plan"]["source"])
            for n in ck["plan"]["nodes"]:
                plan.add(PlanNode(**n))
            self.cluster.install_query(plan)
        self.cluster.restore(ck["state"])
        return cid
