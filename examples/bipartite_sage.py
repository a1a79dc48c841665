"""User-item recommendation with EgoBipartiteSAGE (graphlearn/examples/tf/ego_bipartite_sage):
heterogeneous u->i->u ego graphs, weighted sampling, sampled + in-batch negatives."""
import tempfile

import torch

from common import write_bipartite  # noqa: E402

import graphlearn_b200 as gl
from graphlearn_b200 import models
from graphlearn_b200.nn.data import Data


def main(steps=40, device=None, conv="sage"):
    """conv="gat": the Taobao-shaped config (2-layer, 4-head bipartite GAT, weighted sampling, in-batch negatives)."""
    d = write_bipartite(tempfile.mkdtemp())
    g = gl.Graph() \
        .node(d + "/user.tsv", "u", decoder=gl.Decoder(attr_types=["float"] * 8)) \
        .node(d + "/item.tsv", "i", decoder=gl.Decoder(weighted=True, attr_types=["float"] * 8)) \
        .edge(d + "/u2i.tsv", ("u", "i", "u2i"), decoder=gl.Decoder(weighted=True), directed=False) \
        .init(device=device)
    q = g.E("u2i").batch(64).shuffle(traverse=True).alias("e").each(lambda e: (
        e.outV().alias("u").each(lambda u: (
            u.outV("u2i").sample(4).by("edge_weight").alias("u1").outV("u2i_reverse").sample(3).by("random").alias("u2"),
            u.outNeg("u2i").sample(5).by("in_degree").alias("neg"))),
        e.inV().alias("i").outV("u2i_reverse").sample(4).by("random").alias("i1")
         .outV("u2i").sample(3).by("edge_weight").alias("i2"))).values()
    ds = gl.Dataset(q)
    model = models.EgoBipartiteSAGE(8, 8, 32, 16, hops=2, conv=conv, num_head=4).to(g.device)
    opt = torch.optim.Adam(model.parameters(), lr=5e-3)
    f = lambda v: Data.from_values(v).floats  # noqa: E731
    first = last = None
    for it in range(steps):
        try:
            r = ds.next()
        except gl.OutOfRangeError:
            continue
        ue, ie = model([f(r["u"]), f(r["u1"]), f(r["u2"])], [f(r["i"]), f(r["i1"]), f(r["i2"])], [4, 3], [4, 3])
        neg_x = f(r["neg"])                                   # raw features of sampled negatives ...
        loss = model.in_batch_negative_loss(ue, ie)           # ... in-batch softmax is the main loss
        opt.zero_grad(); loss.backward(); opt.step()
        first = first if first is not None else float(loss)
        last = float(loss)
    print("bipartite sage loss %.4f -> %.4f (neg batch %s)" % (first, last, tuple(neg_x.shape)))
    return first, last


if __name__ == "__main__":
    import sys
    main(conv="gat" if "--gat" in sys.argv else "sage")
