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

    def restore_latest(self) -> Optional[int]:
        ids = self._list()
        if not ids:
            return None
        ck = torch.load(os.path.join(self.path, "ckpt.%d.pt" % ids[-1]), weights_only=True)
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
