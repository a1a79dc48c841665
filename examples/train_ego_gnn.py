"""Supervised EgoGNN node classification with any ego conv: sage | gat | gin
(graphlearn/examples/tf/{ego_sage,ego_gat}).   python examples/train_ego_gnn.py --model gat"""
import argparse
import tempfile

import torch
import torch.nn.functional as F

from common import write_citation_like  # noqa: E402

import graphlearn_b200 as gl
from graphlearn_b200 import models
from graphlearn_b200 import nn as glnn


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="gat", choices=["sage", "gat", "gin"])
    ap.add_argument("--epochs", type=int, default=3)
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--nodes", type=int, default=1500)
    ap.add_argument("--device", default=None)
    a = ap.parse_args(argv)
    node_f, edge_f, dim, classes = write_citation_like(tempfile.mkdtemp(), n=a.nodes)
    g = gl.Graph().node(node_f, "i", decoder=gl.Decoder(labeled=True, attr_types=["float"] * dim)) \
        .edge(edge_f, ("i", "i", "e"), decoder=gl.Decoder(weighted=True), directed=False).init(device=a.device)
    fan = [8, 4]
    q = g.V("i").batch(a.batch).shuffle(traverse=True).alias("src") \
         .outV("e").sample(fan[0]).by("random").alias("h1") \
         .outV("e").sample(fan[1]).by("random").alias("h2").values()
    ds = glnn.Dataset(q)
    model = models.make_ego_gnn(a.model, dim, 32, classes, 2).to(g.device)
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    acc = 0.0
    for ep in range(a.epochs):
        correct = n = 0
        while True:
            try:
                ego = ds.get_egograph("src", ["h1", "h2"])
            except gl.OutOfRangeError:
                break
            logits = model([h.floats for h in ego.hops()], fan)
            loss = F.cross_entropy(logits, ego.src.labels)
            opt.zero_grad(); loss.backward(); opt.step()
            correct += int((logits.argmax(1) == ego.src.labels).sum()); n += ego.src.labels.numel()
        acc = correct / max(n, 1)
        print("epoch %d train acc %.3f" % (ep, acc))
    return acc


if __name__ == "__main__":
    main()
