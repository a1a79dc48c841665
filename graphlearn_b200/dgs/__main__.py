"""Service entry point: ``python -m graphlearn_b200.dgs --schema schema.json [--options opt.yml] ...``.

One process = one streaming service on one GPU (or the CPU) behind the HTTP front end - the unit the Helm chart in
``deploy/dgs`` schedules (the reference starts ``dgs_service`` binaries per worker role from its chart:
dynamic_graph_service/k8s/charts/dgs/templates/{coordinator,sampling,serving}).  At start-up the process restores the
newest checkpoint of ``--checkpoint-dir`` (queries included), optionally installs a query and bulk-loads record files,
then serves until SIGTERM / SIGINT, writing a final checkpoint on the way out.
"""
from __future__ import annotations

import argparse
import json
import signal
import sys
import threading


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="python -m graphlearn_b200.dgs", description=__doc__.split("\n\n")[0])
    ap.add_argument("--schema", required=True, help="graph schema JSON (reference format: attr_defs / vertex_defs / edge_defs / edge_relation_defs)")
    ap.add_argument("--options", default="", help="YAML option file (http-port, checkpoint.path / keep, sample-store.ttl-hours, ...)")
    ap.add_argument("--host", default="0.0.0.0")
    ap.add_argument("--port", type=int, default=None, help="HTTP port (default: options http-port, 0 = ephemeral)")
    ap.add_argument("--device", default=None, help="cuda[:i] | cpu (default: cuda when available)")
    ap.add_argument("--capacity", type=int, default=1 << 16, help="initial rows of the vertex tables (they grow on demand)")
    ap.add_argument("--feat-dim", action="append", default=[], metavar="VTYPE=DIM", help="feature width of a vertex type (repeatable)")
    ap.add_argument("--checkpoint-dir", default="", help="restore from / checkpoint into this directory")
    ap.add_argument("--checkpoint-interval", type=float, default=0.0, help="seconds between periodic checkpoints (0 = only on demand / exit)")
    ap.add_argument("--install-query", default="", help="install-query JSON to install at start-up")
    ap.add_argument("--load", nargs=2, action="append", default=[], metavar=("PATTERN", "DATA"), help="bulk-load a record file described by a pattern file (repeatable)")
    ap.add_argument("--admin-token-file", default="", help="file holding the bearer token required on POST /admin/* (default: $GLB_DGS_ADMIN_TOKEN, empty = open)")
    ap.add_argument("--port-file", default="", help="write the bound port here once the service answers (scripts / probes)")
    a = ap.parse_args(argv)

    from . import CheckpointManager, DynamicGraphService, FileLoader, HttpFrontEnd, Options, QueryPlan, Schema
    schema = Schema.from_json(a.schema)
    opt = Options.from_yaml(a.options) if a.options else Options()
    dims = {}
    for kv in a.feat_dim:
        k, _, v = kv.partition("=")
        dims[k] = int(v)
    svc = DynamicGraphService(schema.to_service_schema(capacity=a.capacity, feat_dims=dims), device=a.device)
    ckpt_dir = a.checkpoint_dir or ""
    import os
    token = open(a.admin_token_file).read().strip() if a.admin_token_file else os.environ.get("GLB_DGS_ADMIN_TOKEN", "")
    front = HttpFrontEnd(svc, schema, checkpoint_dir=ckpt_dir, host=a.host,
                         port=a.port if a.port is not None else int(opt.get("http-port", 0) or 0), admin_token=token)
    if front.ckpt is not None:
        front.ckpt.keep = int(opt.get("checkpoint.keep", 3))
        cid = front.ckpt.restore_latest()
        if cid is not None:
            print("restored checkpoint %d (%d records ingested before it)" % (cid, svc.ingested), flush=True)
    if a.install_query:
        with open(a.install_query) as f:
            d = json.load(f)
        qid = int(d.get("query_id", len(svc.queries)))
        if qid not in svc.queries:
            svc.install_query(qid, QueryPlan.from_json(d, schema))
    for pattern, data in a.load:
        n = FileLoader(pattern, schema).load(data, svc)
        print("loaded %d records from %s" % (n, data), flush=True)
    front.start()
    if front.ckpt is not None and a.checkpoint_interval > 0:
        front.ckpt.start_periodic(a.checkpoint_interval)
    print("serving on %s:%d (device %s)" % (a.host, front.port, svc.device), flush=True)
    if a.port_file:
        with open(a.port_file, "w") as f:
            f.write(str(front.port))
    stop = threading.Event()
    for sig in (signal.SIGTERM, signal.SIGINT):
        signal.signal(sig, lambda *_: stop.set())
    stop.wait()
    front.stop()
    if front.ckpt is not None:
        front.ckpt.stop()
        print("final checkpoint %d" % front.ckpt.save(), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
