"""EgoTGAT on a timestamped graph (graphlearn/examples/tf/ego_tgat): the temporal root makes every
hop sample only edges BEFORE the seed's timestamp, most recent first (top-k), and the layer attends
with time-encoded keys.   python examples/train_ego_tgat.py"""
import argparse
import os
import tempfile

import torch
import torch.nn.functional as F

from common import write_temporal_nodes  # noqa: E402

import graphlearn_b200 as gl
from graphlearn_b200.nn.conv import EgoTGATConv


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=4)
    ap.add_argument("--device", default=None)
    a = ap.parse_args(argv)
    d, dim, classes = write_temporal_nodes(tempfile.mkdtemp())
    gl.set_default_neighbor_id(-1)
    gl.set_padding_mode(gl.REPLICATE)
    g = gl.Graph().node(os.path.join(d, "node.tsv"), "i",
                        decoder=gl.Decoder(labeled=True, timestamped=True, attr_types=["float"] * dim)) \
        .edge(os.path.join(d, "edge.tsv"), ("i", "i", "e"), decoder=gl.Decoder(timestamped=True), directed=False) \
        .init(device=a.device)
    k = 6
    src = g.V("i").batch(64).shuffle(traverse=True).alias("src")
    src.outE("e").sample(k).by("topk").alias("e1").inV().alias("h1")
    ds = gl.Dataset(src.values())
    conv = EgoTGATConv(dim, 32, time_dim=8, num_head=2).to(g.device)
    head = torch.nn.Linear(32, classes).to(g.device)
    opt = torch.optim.Adam(list(conv.parameters()) + list(head.parameters()), lr=1e-2)
    acc = 0.0
    for ep in range(a.epochs):
        correct = n = 0
        while True:
            try:
                res = ds.next()
            except gl.OutOfRangeError:
                break
            s, e1, h1 = res["src"], res["e1"], res["h1"]
            t_self = s.tensor("timestamps").reshape(-1)
            t_edge = e1.tensor("timestamps").reshape(-1)
            assert bool(((t_edge < t_self.repeat_interleave(k)) | (t_edge < 0)).all()), "future edge leaked"
            x = s.tensor("float_attrs").reshape(-1, dim)
            xn = h1.tensor("float_attrs").reshape(-1, dim)
            pad = (h1.ids_t.reshape(-1) < 0)
            xn = torch.where(pad[:, None], torch.zeros_like(xn), xn)
            logits = head(F.relu(conv(x, xn, k, t_self, torch.where(pad, t_self.repeat_interleave(k), t_edge))))
            y = s.tensor("labels").reshape(-1)
            loss = F.cross_entropy(logits, y)
            opt.zero_grad(); loss.backward(); opt.step()
            correct += int((logits.argmax(1) == y).sum()); n += y.numel()
        acc = correct / max(n, 1)
        print("epoch %d train acc %.3f" % (ep, acc))
    return acc


if __name__ == "__main__":
    main()
