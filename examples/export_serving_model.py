"""Export a trained EgoGraphSAGE for online serving and answer inference requests from the streaming service
(graphlearn/examples/tf/serving/export_serving_model.py + the DGS serving pipeline, docs/en/dgs/intro.md:19-68).

The reference rewrites the TF graph so that the sampler's ``IteratorGetNext`` outputs become placeholders and
saves a ``SavedModel``; a serving worker then turns ``/infer?qid&vid`` into an ego graph of cached samples and
feeds TF-Serving.  Here:
  1. train EgoGraphSAGE with the offline GSL pipeline,
  2. export ``model(x_seed, x_hop1, x_hop2) -> embedding`` as a TorchScript module (the placeholders are the
     per-hop feature tensors),
  3. stream the same edges + vertex features into ``DynamicGraphService``, install the matching 2-hop query,
  4. per request: ``run_query`` -> hop features -> exported module.
   python examples/export_serving_model.py"""
import argparse
import os
import tempfile

import numpy as np
import torch
import torch.nn.functional as F

from common import write_citation_like  # noqa: E402

import graphlearn_b200 as gl
from graphlearn_b200 import models
from graphlearn_b200 import nn as glnn
from graphlearn_b200.dgs import DynamicGraphService, QueryPlan

FAN = [4, 3]


class ServingModule(torch.nn.Module):
    """Fixed fan-outs baked in so that the exported graph has tensor-only inputs."""

    def __init__(self, model):
        super().__init__()
        self.model = model

    def forward(self, x0: torch.Tensor, x1: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
        return self.model([x0, x1, x2], FAN)


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=2)
    ap.add_argument("--nodes", type=int, default=800)
    ap.add_argument("--out", default="")
    a = ap.parse_args(argv)
    d = tempfile.mkdtemp()
    node_f, edge_f, dim, classes = write_citation_like(d, n=a.nodes)
    g = gl.Graph().node(node_f, "i", decoder=gl.Decoder(labeled=True, attr_types=["float"] * dim)) \
        .edge(edge_f, ("i", "i", "e"), decoder=gl.Decoder(weighted=True), directed=True).init(device="cpu")
    # ---- 1. offline training (top-k by weight so that offline and online neighbourhoods have the same shape)
    q = g.V("i").batch(128).shuffle(traverse=True).alias("src") \
         .outV("e").sample(FAN[0]).by("random").alias("h1").outV("e").sample(FAN[1]).by("random").alias("h2").values()
    ds = glnn.Dataset(q)
    model = models.EgoGraphSAGE(dim, 32, classes, 2, bf16_activations=False)
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    for _ in range(a.epochs):
        while True:
            try:
                ego = ds.get_egograph("src", ["h1", "h2"])
            except gl.OutOfRangeError:
                break
            loss = F.cross_entropy(model([h.floats for h in ego.hops()], FAN), ego.src.labels)
            opt.zero_grad(); loss.backward(); opt.step()
    # ---- 2. export
    model.eval()
    ex = (torch.zeros(2, dim), torch.zeros(2 * FAN[0], dim), torch.zeros(2 * FAN[0] * FAN[1], dim))
    scripted = torch.jit.trace(ServingModule(model), ex)
    out = a.out or os.path.join(d, "ego_sage_serving.pt")
    scripted.save(out)
    served = torch.jit.load(out)
    # ---- 3. online store: stream vertices + edges, install the 2-hop query
    svc = DynamicGraphService({"vertices": {"i": {"count": a.nodes, "feat_dim": dim}},
                               "edges": {"e": {"src": "i", "dst": "i"}}}, device="cpu")
    svc.install_query(0, QueryPlan("i").out("e", FAN[0]).out("e", FAN[1]))
    ids = np.arange(a.nodes)
    feats = g.lookup_nodes("i", ids).float_attrs
    labels = g.lookup_nodes("i", ids).labels
    svc.apply_updates({"vertices": {"i": {"id": ids, "ts": np.zeros(a.nodes, dtype=np.int64), "feat": feats}}})
    rows = [l.split("\t") for l in open(edge_f).read().strip().split("\n")[1:]]
    src = np.array([int(r[0]) for r in rows]); dst = np.array([int(r[1]) for r in rows])
    for lo in range(0, len(src), 1000):               # edges arrive as a stream of timestamped records
        sl = slice(lo, lo + 1000)
        svc.apply_updates({"edges": {"e": {"src": src[sl], "dst": dst[sl], "ts": np.arange(lo, lo + len(src[sl]))}}})
    # ---- 4. requests
    req = torch.arange(0, a.nodes, 7)
    res = svc.run_query(0, req)
    x0 = svc.vstores["i"].feat[req]
    x1 = res["hops"][0]["features"].reshape(-1, dim)
    x2 = res["hops"][1]["features"].reshape(-1, dim)
    with torch.no_grad():
        logits = served(x0, x1, x2)
        same = torch.allclose(logits, model([x0, x1, x2], FAN), atol=1e-5)
    acc = float((logits.argmax(1).numpy() == labels[req.numpy()]).mean())
    print("exported %s; served %d requests, accuracy %.3f, scripted == eager: %s" % (out, req.numel(), acc, same))
    return acc, same, out


if __name__ == "__main__":
    main()
