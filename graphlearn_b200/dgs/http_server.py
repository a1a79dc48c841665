"""HTTP front end (D11 + the admin surface of D12): the reference serves ``/infer?qid=&vid=`` from a
seastar httpd on every serving worker (dynamic_graph_service/src/service/service.cc, docs/en/dgs) and
installs queries / triggers checkpoints / sets barriers through the coordinator's admin HTTP API.
Here one ``ThreadingHTTPServer`` fronts a ``DynamicGraphService``; query work is serialised on the
device stream by a lock.

  GET  /infer?qid=0&vid=12[,13,...]   -> {"src": [...], "nodes": {plan_node_id: {...}}}
  POST /admin/init                     body = install-query JSON (reference format)   -> {"query_id": n}
  POST /admin/ingest                   body = {"edges": {...}, "vertices": {...}}     (testing / small feeds)
  POST /admin/load                     body = {"pattern": path, "data": path[, "reverse_edges": {etype: reversed etype}]}: bulk-load a
                                       record file that the SERVICE can read (the file-loader app; native record parser)
  POST /admin/checkpoint               -> {"checkpoint_id": n}
  POST /admin/barrier/set?name=x[&produced=n]     GET /admin/barrier/status?name=x
  GET  /admin/stats                    liveness + counters (k8s probes)
  GET  /admin/schema                   the graph schema (reference JSON)        GET /admin/query[?qid=n]  an installed query
  GET  /admin/init-info/dataloader     what a data loader needs: ingest endpoints, data partition count, schema
The Python GSL client (``dgs/client.py``, the role of the reference's Java client) speaks exactly this surface.
"""
from __future__ import annotations

import json
import threading
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer
from urllib.parse import parse_qs, urlparse

import torch

from .coordinator import BarrierMonitor, CheckpointManager
from .plan import QueryPlan


def _jsonable(x):
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().tolist()
    if isinstance(x, dict):
        return {str(k): _jsonable(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [_jsonable(v) for v in x]
    return x


class HttpFrontEnd(object):
    def __init__(self, service, schema=None, checkpoint_dir: str = "", host: str = "127.0.0.1", port: int = 0,
                 admin_token: str = ""):
        """``admin_token``: when set, every POST /admin/* call must carry ``Authorization: Bearer <token>`` (the admin surface
        installs queries, ingests records and reads server-side files; ``/infer`` and the read-only GETs stay open)."""
        self.service, self.schema = service, schema
        self.admin_token = admin_token or ""
        self.ckpt = CheckpointManager(service, checkpoint_dir) if checkpoint_dir else None
        self.barriers = BarrierMonitor(service)
        self._lock = threading.Lock()
        self.installed = {}                 # qid -> install-query JSON (+ node_ids), answered by /admin/query
        front = self

        class Handler(BaseHTTPRequestHandler):
            def log_message(self, *a):      # quiet
                pass

            def _send(self, code, obj):
                body = json.dumps(obj).encode()
                self.send_response(code)
                self.send_header("Content-Type", "application/json")
                self.send_header("Content-Length", str(len(body)))
                self.end_headers()
                self.wfile.write(body)

            def _body(self):
                n = int(self.headers.get("Content-Length", 0) or 0)
                return json.loads(self.rfile.read(n) or b"{}") if n else {}

            def do_GET(self):
                u = urlparse(self.path)
                q = parse_qs(u.query)
                try:
                    if u.path == "/infer":
                        vids = [int(v) for v in q["vid"][0].split(",")]
                        with front._lock:
                            res = front.service.run_query(int(q["qid"][0]), vids)
                        self._send(200, {"src": _jsonable(res["src"]), "nodes": _jsonable(res["nodes"])})
                    elif u.path == "/admin/stats":
                        self._send(200, front.service.stats())
                    elif u.path == "/admin/barrier/status":
                        self._send(200, {"status": front.barriers.status(q["name"][0])})
                    elif u.path == "/admin/schema":
                        self._send(200, front.schema.raw if front.schema is not None else {})
                    elif u.path == "/admin/init-info/dataloader":
                        # what a data loader needs before producing (the reference hands out the Kafka brokers / topic /
                        # partition count + the schema, coordinator/http_service.py:95-103): here records go to /admin/ingest
                        # (or /admin/load), one data partition per service process
                        parts = int(getattr(front.service, "P", 1))
                        self._send(200, {"downstream": {"transport": "http", "ingest": "/admin/ingest", "load": "/admin/load"},
                                         "data_partition_num": parts, "schema": front.schema.raw if front.schema is not None else {}})
                    elif u.path == "/admin/query":
                        if not front.installed:
                            self._send(404, {"error": "no query installed"})
                        else:
                            qid = int(q["qid"][0]) if "qid" in q else max(front.installed)
                            self._send(200, front.installed[qid])
                    else:
                        self._send(404, {"error": "unknown path"})
                except Exception as e:  # noqa: BLE001
                    self._send(400, {"error": repr(e)})

            def do_POST(self):
                u = urlparse(self.path)
                q = parse_qs(u.query)
                if front.admin_token and u.path.startswith("/admin/"):
                    import hmac
                    got = self.headers.get("Authorization", "")
                    if not hmac.compare_digest(got, "Bearer " + front.admin_token):
                        self._body()
                        self._send(401, {"error": "admin token required"})
                        return
                try:
                    if u.path == "/admin/init":
                        d = self._body()
                        plan = QueryPlan.from_json(d, front.schema)
                        qid = int(d.get("query_id", len(front.service.queries)))
                        with front._lock:
                            front.service.install_query(qid, plan)
                        node_ids = {str(k): int(v) for k, v in getattr(plan, "json_ids", {}).items()}
                        front.installed[qid] = {**d, "query_id": qid, "node_ids": node_ids}
                        self._send(200, {"query_id": qid, "node_ids": node_ids})
                    elif u.path == "/admin/ingest":
                        with front._lock:
                            front.service.apply_updates(self._body())
                        self._send(200, {"ingested": front.service.ingested})
                    elif u.path == "/admin/load":
                        from .file_loader import FileLoader
                        d = self._body()
                        fl = FileLoader(d["pattern"], front.schema, reverse_edges=d.get("reverse_edges"))
                        with front._lock:
                            n = fl.load(d["data"], front.service)
                        self._send(200, {"records": n, "ingested": front.service.ingested})
                    elif u.path == "/admin/checkpoint":
                        with front._lock:
                            cid = front.ckpt.save()
                        self._send(200, {"checkpoint_id": cid})
                    elif u.path == "/admin/barrier/set":
                        front.barriers.set(q["name"][0], int(q["produced"][0]) if "produced" in q else None)
                        self._send(200, {"ok": True})
                    else:
                        self._send(404, {"error": "unknown path"})
                except Exception as e:  # noqa: BLE001
                    self._send(400, {"error": repr(e)})

        self.httpd = ThreadingHTTPServer((host, port), Handler)
        self.port = self.httpd.server_address[1]
        self._thread = None

    def start(self):
        self._thread = threading.Thread(target=self.httpd.serve_forever, daemon=True)
        self._thread.start()
        return self

    def stop(self):
        self.httpd.shutdown()
        self.httpd.server_close()
