"""u2i end to end, deployment shaped (the reference's DGS tutorial: python/data/u2i/u2i_generator.py -> file loader -> service
-> Java client -> TF-Serving; docs/en/dgs/tutorial):

  1. generate one synthetic user-click history and write it TWICE: as static tables for offline training (the TSV dialect)
     and as a stream of DGS records (+ schema / pattern files),
  2. train a 2-hop bipartite-style EgoGraphSAGE offline on the static snapshot (user -> clicked items -> co-clicked items),
  3. start the streaming service as its OWN PROCESS (`python -m graphlearn_b200.dgs`),
  4. connect with the GSL client, build + install the 2-hop query with the fluent API (the installed query decides which
     sampler states exist, so it comes before the data), have the service bulk-load the record file through the native
     record parser, set a barrier and wait for it,
  5. answer requests: client.run -> EgoGraph.hop_tensors -> model -> which item category the user clicks most.

    python examples/u2i_online_pipeline.py            (CPU; add --device cuda for the service on a GPU)
"""
import argparse
import json
import os
import signal
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch
import torch.nn.functional as F

from common import sys as _sys  # noqa: F401  (repo root on sys.path)

import graphlearn_b200 as gl
from graphlearn_b200 import models
from graphlearn_b200 import nn as glnn
from graphlearn_b200.dgs import client as C

FAN = [5, 3]
DIM = 8


def generate(d, n_user=200, n_item=120, n_cat=4, clicks=12, seed=0):
    """users prefer one item category; items of a category share a feature centre; i2i = consecutive clicks of a user"""
    rs = np.random.RandomState(seed)
    cat_of_item = rs.randint(0, n_cat, n_item)
    centres = rs.randn(n_cat, DIM) * 2
    item_x = centres[cat_of_item] + 0.5 * rs.randn(n_item, DIM)
    pref = rs.randint(0, n_cat, n_user)
    user_x = 0.1 * rs.randn(n_user, DIM)
    by_cat = [np.where(cat_of_item == c)[0] for c in range(n_cat)]
    u2i, i2i, t = [], [], 1000
    for u in range(n_user):
        prev = None
        for _ in range(clicks):
            c = pref[u] if rs.rand() < 0.85 else rs.randint(0, n_cat)
            i = int(rs.choice(by_cat[c]))
            t += 1
            u2i.append((u, i, t))
            if prev is not None and prev != i:
                i2i.append((prev, i, t))
            prev = i
    os.makedirs(d, exist_ok=True)
    fmt = lambda x: ":".join("%.4f" % v for v in x)  # noqa: E731
    # ---- static tables (offline training)
    with open(d + "/user.tsv", "w") as f:
        f.write("id:int64\tlabel:int32\tfeature:string\n" + "".join("%d\t%d\t%s\n" % (u, pref[u], fmt(user_x[u])) for u in range(n_user)))
    with open(d + "/item.tsv", "w") as f:
        f.write("id:int64\tfeature:string\n" + "".join("%d\t%s\n" % (i, fmt(item_x[i])) for i in range(n_item)))
    with open(d + "/u2i.tsv", "w") as f:
        f.write("src_id:int64\tdst_id:int64\ttimestamp:int64\n" + "".join("%d\t%d\t%d\n" % e for e in u2i))
    with open(d + "/i2i.tsv", "w") as f:
        f.write("src_id:int64\tdst_id:int64\ttimestamp:int64\n" + "".join("%d\t%d\t%d\n" % e for e in i2i))
    # ---- streaming records + service configuration (the reference's schema / pattern formats)
    schema = {"attr_defs": [{"type": 0, "name": "timestamp", "value_type": "INT64"}, {"type": 1, "name": "feature", "value_type": "FLOAT32_LIST"}],
              "vertex_defs": [{"vtype": 0, "name": "user", "attr_types": [0, 1]}, {"vtype": 1, "name": "item", "attr_types": [0, 1]}],
              "edge_defs": [{"etype": 2, "name": "u2i", "attr_types": [0]}, {"etype": 3, "name": "i2i", "attr_types": [0]}],
              "edge_relation_defs": [{"etype": 2, "src_vtype": 0, "dst_vtype": 1}, {"etype": 3, "src_vtype": 1, "dst_vtype": 1}]}
    json.dump(schema, open(d + "/schema.json", "w"))
    with open(d + "/pattern", "w") as f:
        f.write("#VERTEX:user,vid,timestamp,feature\n#VERTEX:item,vid,timestamp,feature\n#EDGE:u2i,src,dst,timestamp\n#EDGE:i2i,src,dst,timestamp\n")
    with open(d + "/records", "w") as f:
        for u in range(n_user):
            f.write("user,%d,0,%s\n" % (u, fmt(user_x[u])))
        for i in range(n_item):
            f.write("item,%d,0,%s\n" % (i, fmt(item_x[i])))
        events = sorted([("u2i",) + e for e in u2i] + [("i2i",) + e for e in i2i], key=lambda e: e[3])
        for e in events:
            f.write("%s,%d,%d,%d\n" % e)
    return pref, len(u2i) + len(i2i) + n_user + n_item


def train_offline(d, epochs):
    g = gl.Graph()
    g.node(d + "/user.tsv", "user", decoder=gl.Decoder(labeled=True, attr_types=["float"] * DIM))
    g.node(d + "/item.tsv", "item", decoder=gl.Decoder(attr_types=["float"] * DIM))
    g.edge(d + "/u2i.tsv", ("user", "item", "u2i"), decoder=gl.Decoder(timestamped=True))
    g.edge(d + "/i2i.tsv", ("item", "item", "i2i"), decoder=gl.Decoder(timestamped=True))
    g.init(device="cpu")
    # most recent clicks first - the same neighbourhoods the streaming samplers keep (top-k by timestamp)
    q = (g.V("user").batch(64).shuffle(traverse=True).alias("u")
          .outV("u2i").sample(FAN[0]).by("topk").alias("h1").outV("i2i").sample(FAN[1]).by("topk").alias("h2").values())
    ds = glnn.Dataset(q)
    model = models.EgoGraphSAGE(DIM, 32, 4, 2, bf16_activations=False)
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    for _ in range(epochs):
        while True:
            try:
                ego = ds.get_egograph("u", ["h1", "h2"])
            except gl.OutOfRangeError:
                break
            loss = F.cross_entropy(model([h.floats for h in ego.hops()], FAN), ego.src.labels)
            opt.zero_grad(); loss.backward(); opt.step()
    g.close()
    return model.eval()


def start_service(d, device):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""))
    port_file = d + "/port"
    p = subprocess.Popen([sys.executable, "-m", "graphlearn_b200.dgs", "--schema", d + "/schema.json", "--host", "127.0.0.1", "--port", "0",
                          "--device", device, "--capacity", "64", "--feat-dim", "user=%d" % DIM, "--feat-dim", "item=%d" % DIM,
                          "--port-file", port_file], env=env,
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    t0 = time.time()
    while not os.path.exists(port_file) and p.poll() is None and time.time() - t0 < 180:
        time.sleep(0.2)
    if not os.path.exists(port_file):
        raise RuntimeError("service did not start:\n" + p.communicate(timeout=5)[0].decode()[-2000:])
    return p, int(open(port_file).read())


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--epochs", type=int, default=6)
    a = ap.parse_args(argv)
    d = tempfile.mkdtemp(prefix="glb_u2i_")
    pref, n_records = generate(d)
    model = train_offline(d, a.epochs)
    proc, port = start_service(d, a.device)
    try:
        g = C.Graph.connect("127.0.0.1:%d" % port)
        users = list(range(0, 200, 3))
        q = (g.V("user").feed(C.DataSource(users, batch=16)).properties(1).alias("u")
              .outV("u2i").sample(FAN[0]).by("topk_by_timestamp").properties(1).alias("h1")
              .outV("i2i").sample(FAN[1]).by("topk_by_timestamp").properties(1).alias("h2").values())
        assert g.install(q).ok()
        assert g.load_file(d + "/pattern", d + "/records") == n_records
        g.set_barrier("loaded")
        while not g.check_barrier("loaded").ok():
            time.sleep(0.05)
        hit = tot = 0
        while q.source.has_next():
            ego = g.run(q).ego_graph()
            x = [torch.from_numpy(t) for t in ego.hop_tensors(DIM)]
            with torch.no_grad():
                pred = model(x, FAN).argmax(1).numpy()
            seeds = ego.get_vids(0)
            hit += int((pred == pref[seeds]).sum())
            tot += len(seeds)
        acc = hit / tot
        served = g.stats()["served"]
        g.close()
    finally:
        proc.send_signal(signal.SIGTERM)
        proc.communicate(timeout=60)
    print("served %d users through the service process: preferred-category accuracy %.3f (%d records loaded)" % (tot, acc, n_records))
    return acc, tot, served


if __name__ == "__main__":
    main()
