"""Unsupervised GraphSAGE (graphlearn/examples/tf/ego_sage/train_unsupervised.py): positive pairs
from edges, sampled negatives, one shared EgoGraphSAGE encoder for src / dst / neg ego graphs,
sigmoid cross entropy; the learned embeddings are exported in the reference's
``id:int64\\temb:string`` dialect.   python examples/train_unsupervised_sage.py"""
import argparse
import os
import tempfile

import torch

from common import write_citation_like  # noqa: E402

import graphlearn_b200 as gl
from graphlearn_b200 import models
from graphlearn_b200.nn.data import Data
from graphlearn_b200.nn.loss import sigmoid_cross_entropy_loss
from graphlearn_b200.utils.checkpoint import save_embeddings


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=2)
    ap.add_argument("--nodes", type=int, default=1000)
    ap.add_argument("--device", default=None)
    ap.add_argument("--out", default="")
    a = ap.parse_args(argv)
    node_f, edge_f, dim, classes = write_citation_like(tempfile.mkdtemp(), n=a.nodes)
    g = gl.Graph().node(node_f, "i", decoder=gl.Decoder(labeled=True, attr_types=["float"] * dim)) \
        .edge(edge_f, ("i", "i", "e"), decoder=gl.Decoder(weighted=True), directed=False).init(device=a.device)
    fan, neg_num = [5, 3], 4

    def ego(node, prefix):
        h1 = node.outV("e").sample(fan[0]).by("random").alias(prefix + "1")
        h1.outV("e").sample(fan[1]).by("random").alias(prefix + "2")

    e = g.E("e").batch(128).shuffle(traverse=True).alias("edge")
    s = e.outV().alias("s"); ego(s, "s")
    d = e.inV().alias("d"); ego(d, "d")
    n = s.outNeg("e").sample(neg_num).by("in_degree").alias("n"); ego(n, "n")
    ds = gl.Dataset(e.values())
    model = models.EgoGraphSAGE(dim, 32, 16, 2, bf16_activations=False).to(g.device)
    opt = torch.optim.Adam(model.parameters(), lr=5e-3)
    f = lambda v: Data.from_values(v).floats  # noqa: E731
    enc = lambda r, p: model([f(r[p]), f(r[p + "1"]), f(r[p + "2"])], fan)  # noqa: E731
    first = last = None
    for ep in range(a.epochs):
        tot, k = 0.0, 0
        while True:
            try:
                r = ds.next()
            except gl.OutOfRangeError:
                break
            zs, zd, zn = enc(r, "s"), enc(r, "d"), enc(r, "n")
            pos = (zs * zd).sum(-1)
            neg = (zs.unsqueeze(1) * zn.view(zs.size(0), neg_num, -1)).sum(-1)
            loss = sigmoid_cross_entropy_loss(pos, neg.reshape(-1))
            opt.zero_grad(); loss.backward(); opt.step()
            tot += float(loss.detach()); k += 1
        first = tot / k if first is None else first
        last = tot / k
        print("epoch %d loss %.4f" % (ep, last))
    # export embeddings of all nodes
    q = g.V("i").batch(256).alias("s")
    ego(q, "s")
    ds2 = gl.Dataset(q.values())
    ids, embs = [], []
    model.eval()
    with torch.no_grad():
        while True:
            try:
                r = ds2.next()
            except gl.OutOfRangeError:
                break
            ids.append(r["s"].ids_t.reshape(-1)); embs.append(enc(r, "s"))
    out = a.out or os.path.join(tempfile.mkdtemp(), "emb.tsv")
    save_embeddings(out, torch.cat(ids), torch.cat(embs))
    print("saved %d embeddings to %s" % (sum(i.numel() for i in ids), out))
    return first, last, out


if __name__ == "__main__":
    main()
