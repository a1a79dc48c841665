"""UltraGCN recommendation (graphlearn/examples/tf/ultra_gcn): imperative samplers - shuffled u-i
edges, sampled negatives, top-k i2i neighbours of the positive item with their weights, out-degrees -
feed the embedding model's constraint losses.   python examples/ultra_gcn.py"""
import argparse
import os
import tempfile

import numpy as np
import torch

from common import write_bipartite  # noqa: E402

import graphlearn_b200 as gl
from graphlearn_b200 import models


def write_i2i(d, n_item, seed=0):
    rs = np.random.RandomState(seed)
    with open(os.path.join(d, "i2i.tsv"), "w") as f:
        f.write("src_id:int64\tdst_id:int64\tweight:float\n")
        for i in range(n_item):
            for j in rs.choice(n_item, 6, replace=False):
                f.write("%d\t%d\t%.3f\n" % (i, j, rs.rand()))
    return os.path.join(d, "i2i.tsv")


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=3)
    ap.add_argument("--device", default=None)
    a = ap.parse_args(argv)
    n_user, n_item, nbr_num, neg_num = 300, 500, 4, 5
    d = write_bipartite(tempfile.mkdtemp(), n_user, n_item)
    i2i = write_i2i(d, n_item)
    g = gl.Graph().node(d + "/user.tsv", "u", decoder=gl.Decoder(attr_types=["float"] * 8)) \
        .node(d + "/item.tsv", "i", decoder=gl.Decoder(weighted=True, attr_types=["float"] * 8)) \
        .edge(d + "/u2i.tsv", ("u", "i", "u-i"), decoder=gl.Decoder(weighted=True), directed=False) \
        .edge(i2i, ("i", "i", "i-i"), decoder=gl.Decoder(weighted=True), directed=True) \
        .init(device=a.device)
    edge_sampler = g.edge_sampler("u-i", 128, strategy="shuffle")
    neg_sampler = g.negative_sampler("u-i", neg_num, "random")
    nbr_sampler = g.neighbor_sampler("i-i", nbr_num, strategy="topk")
    model = models.UltraGCN(n_user, n_item, 16).to(g.device)
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    dev = g.device
    t = lambda x, dt=None: torch.as_tensor(np.asarray(x), device=dev).to(dt) if dt else torch.as_tensor(np.asarray(x), device=dev)  # noqa: E731
    first = last = None
    for ep in range(a.epochs):
        tot, n = 0.0, 0
        while True:
            try:
                edges = edge_sampler.get()
            except gl.OutOfRangeError:
                break
            neg = neg_sampler.get(edges.src_ids)
            nbrs = nbr_sampler.get(edges.dst_ids)
            loss = model(t(edges.src_ids), t(g.out_degrees(edges.src_ids, "u-i"), torch.float32),
                         t(edges.dst_ids), t(g.out_degrees(edges.dst_ids, "u-i_reverse"), torch.float32),
                         t(nbrs.layer_nodes(1).ids), t(nbrs.layer_edges(1).weights, torch.float32), t(neg.ids))
            loss = loss / len(edges.src_ids)
            opt.zero_grad(); loss.backward(); opt.step()
            tot += float(loss.detach()); n += 1
        first = tot / n if first is None else first
        last = tot / n
        print("epoch %d loss %.4f" % (ep, tot / n))
    # recall@20 of held-in edges as a sanity metric
    with torch.no_grad():
        U = model.user_embeddings(torch.arange(n_user, device=dev))
        I = model.item_embeddings(torch.arange(n_item, device=dev))
        top = (U @ I.t()).topk(20, dim=1).indices.cpu().numpy()
    hit = tot_e = 0
    for line in open(d + "/u2i.tsv").read().strip().split("\n")[1:]:
        u, i = int(line.split("\t")[0]), int(line.split("\t")[1])
        hit += int(i in top[u]); tot_e += 1
    print("train-edge recall@20 %.3f" % (hit / tot_e))
    return first, last, hit / tot_e


if __name__ == "__main__":
    main()
