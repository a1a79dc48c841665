"""DeepWalk / node2vec (graphlearn/examples/tf/node2vec): walks from the resident-walker kernel,
skip-gram pairs, sampled negatives.   python examples/node2vec.py [--p 0.5 --q 2]"""
import argparse
import tempfile

import torch

from common import write_citation_like  # noqa: E402

import graphlearn_b200 as gl
from graphlearn_b200 import models


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--p", type=float, default=1.0)
    ap.add_argument("--q", type=float, default=1.0)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--device", default=None)
    a = ap.parse_args(argv)
    node_f, edge_f, dim, _ = write_citation_like(tempfile.mkdtemp(), n=1000)
    g = gl.Graph().node(node_f, "i", decoder=gl.Decoder(labeled=True, attr_types=["float"] * dim)) \
        .edge(edge_f, ("i", "i", "e"), decoder=gl.Decoder(weighted=True), directed=False).init(device=a.device)
    q = g.V("i").batch(64).shuffle(traverse=True).alias("src") \
         .random_walk("e", 10, p=a.p, q=a.q).alias("walk").values()
    ds = gl.Dataset(q)
    neg = g.negative_sampler("e", 5, "in_degree")
    model = models.Node2Vec(1000, 32, sparse=False).to(g.device)
    opt = torch.optim.Adam(model.parameters(), lr=0.02)
    first = last = None
    for it in range(a.steps):
        try:
            r = ds.next()
        except gl.OutOfRangeError:
            continue
        path = torch.cat([r["src"].ids_t[:, None], r["walk"].ids_t], 1)
        s, d = models.gen_pair(path, 2, 2)
        n = neg.get(s).ids_t
        loss = model(s, d, n)
        opt.zero_grad(); loss.backward(); opt.step()
        first = first if first is not None else float(loss)
        last = float(loss)
    print("node2vec loss %.4f -> %.4f" % (first, last))
    return first, last


if __name__ == "__main__":
    main()
