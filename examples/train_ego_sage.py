"""Supervised EgoGraphSAGE node classification through the public API (the reference's
examples/tf/ego_sage):  TSV files -> gl.Graph -> GSL 2-hop query -> gl.nn.Dataset -> EgoGraphSAGE.
Runs on CPU or one GPU;  python examples/train_ego_sage.py [--epochs 3]"""
import argparse
import tempfile

import torch
import torch.nn.functional as F

from common import write_citation_like  # noqa: E402  (also fixes sys.path)

import graphlearn_b200 as gl
from graphlearn_b200 import models
from graphlearn_b200 import nn as glnn


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=3)
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--device", default=None)
    args = ap.parse_args(argv)
    d = tempfile.mkdtemp()
    node_f, edge_f, dim, classes = write_citation_like(d)
    g = gl.Graph() \
        .node(node_f, "i", decoder=gl.Decoder(labeled=True, attr_types=["float"] * dim)) \
        .edge(edge_f, ("i", "i", "e"), decoder=gl.Decoder(weighted=True), directed=False) \
        .init(device=args.device)
    fan = [10, 5]
    q = g.V("i").batch(args.batch).shuffle(traverse=True).alias("src") \
         .outV("e").sample(fan[0]).by("random").alias("h1") \
         .outV("e").sample(fan[1]).by("random").alias("h2").values()
    ds = glnn.Dataset(q)
    model = models.EgoGraphSAGE(dim, 64, classes, 2, bf16_activations=False).to(g.device)
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    for ep in range(args.epochs):
        tot, correct, n, steps = 0.0, 0, 0, 0
        while True:
            try:
                ego = ds.get_egograph("src", ["h1", "h2"])
            except gl.OutOfRangeError:
                break
            logits = model([h.floats for h in ego.hops()], fan)
            loss = F.cross_entropy(logits, ego.src.labels)
            opt.zero_grad()
            loss.backward()
            opt.step()
            tot += float(loss); steps += 1
            correct += int((logits.argmax(1) == ego.src.labels).sum()); n += ego.src.labels.numel()
        print("epoch %d  loss %.4f  train acc %.3f" % (ep, tot / max(steps, 1), correct / max(n, 1)))
    return correct / max(n, 1)


if __name__ == "__main__":
    main()
