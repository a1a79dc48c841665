"""Link prediction with SubGraph-based GraphSAGE (graphlearn/examples/tf/sage/train.py + edge_inducer.py):
edges are the seeds, both end points and one sampled negative destination get their full 1-hop
neighbourhoods, an ``EdgeInducer`` turns every (src, dst) pair into an enclosing subgraph
[src, dst, src's nbrs, dst's nbrs], the subgraphs are batched into one edge_index and a sparse SAGE model
embeds src / dst; loss = sigmoid cross entropy of pos vs neg pairs.
   python examples/train_subgraph_sage.py"""
import argparse
import tempfile

import torch

from common import write_citation_like  # noqa: E402

import graphlearn_b200 as gl
from graphlearn_b200 import models
from graphlearn_b200 import nn as glnn
from graphlearn_b200.nn.loss import sigmoid_cross_entropy_loss


class EdgeInducer(glnn.SubGraphInducer):
    """Vectorised form of the reference's per-edge python loop: every pair i becomes one star-pair subgraph; all
    subgraphs of the batch are emitted as ONE batched graph (x, edge_index, src_index, dst_index)."""

    def __init__(self, use_neg=True):
        super().__init__(use_neg=use_neg)

    @staticmethod
    def _pairs(src, dst, src_nbrs, dst_nbrs):
        dev = src.ids_t.device
        B = int(src.ids_t.numel())
        xs, xd = src.tensor("float_attrs").reshape(B, -1), dst.tensor("float_attrs").reshape(B, -1)
        cs = torch.as_tensor(src_nbrs.offsets, device=dev).long()
        cd = torch.as_tensor(dst_nbrs.offsets, device=dev).long()
        xsn, xdn = src_nbrs.tensor("float_attrs"), dst_nbrs.tensor("float_attrs")
        # node order of the batched graph: [all src | all dst | all src nbrs | all dst nbrs]
        x = torch.cat([xs, xd, xsn, xdn])
        src_index = torch.arange(B, device=dev)
        dst_index = B + src_index
        sn = 2 * B + torch.arange(xsn.size(0), device=dev)
        dn = 2 * B + xsn.size(0) + torch.arange(xdn.size(0), device=dev)
        s_of = torch.repeat_interleave(src_index, cs)
        d_of = torch.repeat_interleave(dst_index, cd)
        row = torch.cat([s_of, sn, d_of, dn, src_index, dst_index])          # both directions + the target edge
        col = torch.cat([sn, s_of, dn, d_of, dst_index, src_index])
        return glnn.SubGraphData(x, torch.stack([row, col]), src_index=src_index, dst_index=dst_index)

    def induce_func(self, values):
        pos = self._pairs(values["pos_src"], values["pos_dst"], values["src_hop1"], values["dst_hop1"])
        neg = self._pairs(values["pos_src"], values["neg_dst"], values["src_hop1"], values["neg_hop1"]) if self.use_neg else None
        return pos, neg


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=2)
    ap.add_argument("--nodes", type=int, default=800)
    ap.add_argument("--device", default=None)
    a = ap.parse_args(argv)
    node_f, edge_f, dim, _ = write_citation_like(tempfile.mkdtemp(), n=a.nodes)
    g = gl.Graph().node(node_f, "i", decoder=gl.Decoder(labeled=True, attr_types=["float"] * dim)) \
        .edge(edge_f, ("i", "i", "train"), decoder=gl.Decoder(weighted=True), directed=False).init(device=a.device)
    nbrs = 20
    seed = g.E("train").batch(128).shuffle(traverse=True).alias("seed")
    src = seed.outV().alias("pos_src")
    src.outV("train").sample(nbrs).by("full").alias("src_hop1")
    dst = seed.inV().alias("pos_dst")
    dst.outV("train").sample(nbrs).by("full").alias("dst_hop1")
    src.outNeg("train").sample(1).by("random").alias("neg_dst").outV("train").sample(nbrs).by("full").alias("neg_hop1")
    ds = gl.Dataset(seed.values())
    inducer = EdgeInducer(use_neg=True)
    model = models.SparseGNN("sage", dim, 32, 32, 2).to(g.device)
    opt = torch.optim.Adam(model.parameters(), lr=5e-3)
    first = last = None

    def embed(sg):
        z = model(sg.x, sg.edge_index)
        return z[sg.src_index], z[sg.dst_index]

    for ep in range(a.epochs):
        tot, n = 0.0, 0
        while True:
            try:
                values = ds.next()
            except gl.OutOfRangeError:
                break
            pos, neg = inducer.induce_func(values)
            ps, pd = embed(pos)
            ns, nd = embed(neg)
            loss = sigmoid_cross_entropy_loss((ps * pd).sum(-1), (ns * nd).sum(-1))
            opt.zero_grad(); loss.backward(); opt.step()
            tot += float(loss.detach()); n += 1
        first = tot / n if first is None else first
        last = tot / n
        print("epoch %d loss %.4f" % (ep, last))
    return first, last


if __name__ == "__main__":
    main()
