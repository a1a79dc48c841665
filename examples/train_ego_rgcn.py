"""EgoRGCN over a multi-relation graph (graphlearn/examples/tf/ego_rgcn): one sampled hop per
relation, relation-specific weights with basis decomposition.   python examples/train_ego_rgcn.py"""
import argparse
import os
import tempfile

import torch
import torch.nn.functional as F

from common import write_hetero  # noqa: E402

import graphlearn_b200 as gl
from graphlearn_b200.nn.conv import EgoRGCNConv
from graphlearn_b200.nn.data import Data


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=4)
    ap.add_argument("--device", default=None)
    a = ap.parse_args(argv)
    d, dim, classes = write_hetero(tempfile.mkdtemp())
    g = gl.Graph().node(os.path.join(d, "node.tsv"), "i", decoder=gl.Decoder(labeled=True, attr_types=["float"] * dim))
    for r in ("a", "b"):
        g.edge(os.path.join(d, "edge_%s.tsv" % r), ("i", "i", r), decoder=gl.Decoder(), directed=False)
    g.init(device=a.device)
    k = 5
    src = g.V("i").batch(64).shuffle(traverse=True).alias("src")
    for r in ("a", "b"):
        src.outV(r).sample(k).by("random").alias("nbr_" + r)
    ds = gl.Dataset(src.values())
    conv = EgoRGCNConv(dim, classes, num_relations=2, num_bases=2).to(g.device)
    opt = torch.optim.Adam(conv.parameters(), lr=2e-2)
    acc = 0.0
    for ep in range(a.epochs):
        correct = n = 0
        while True:
            try:
                res = ds.next()
            except gl.OutOfRangeError:
                break
            x = Data.from_values(res["src"])
            nbrs = [Data.from_values(res["nbr_" + r]).floats for r in ("a", "b")]
            logits = conv(x.floats, nbrs, [k, k])
            loss = F.cross_entropy(logits, x.labels)
            opt.zero_grad(); loss.backward(); opt.step()
            correct += int((logits.argmax(1) == x.labels).sum()); n += x.labels.numel()
        acc = correct / max(n, 1)
        print("epoch %d train acc %.3f" % (ep, acc))
    return acc


if __name__ == "__main__":
    main()
