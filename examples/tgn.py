"""TGN link prediction on a JODIE-shaped event stream (graphlearn/examples/pytorch/tgn).
   python examples/tgn.py [--epochs 3]"""
import argparse
import os
import tempfile

import torch

from common import write_temporal  # noqa: E402

import graphlearn_b200 as gl
from graphlearn_b200 import models


def auc(pos, neg):
    s = torch.cat([pos, neg]).reshape(-1)
    y = torch.cat([torch.ones_like(pos), torch.zeros_like(neg)]).reshape(-1)
    order = torch.argsort(s)
    rank = torch.empty_like(s)
    rank[order] = torch.arange(1, s.numel() + 1, dtype=s.dtype, device=s.device)
    npos, nneg = y.sum(), (1 - y).sum()
    return float((rank[y > 0].sum() - npos * (npos + 1) / 2) / (npos * nneg).clamp(min=1))


def run_epoch(g, model, source, a, num_nodes, opt=None):
    loader = models.TemporalBatchLoader(g, source, num_nodes, a.batch_size, a.nbr_size, a.msg_dim)
    tot, n, aucs = 0.0, 0, []
    for batch in loader:
        if opt is not None:
            opt.zero_grad()
            loss = model.loss(batch)
            model.update(batch)
            loss.backward()
            opt.step()
            model.memory.detach()
            tot += float(loss) * batch.num_events
        else:
            with torch.no_grad():
                pos, neg = model(batch)
                aucs.append(auc(pos.sigmoid(), neg.sigmoid()))
                model.update(batch)
        n += batch.num_events
    return (tot / max(n, 1)) if opt is not None else sum(aucs) / max(len(aucs), 1)


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=3)
    ap.add_argument("--batch_size", type=int, default=100)
    ap.add_argument("--nbr_size", type=int, default=5)
    ap.add_argument("--msg_dim", type=int, default=6)
    ap.add_argument("--device", default=None)
    a = ap.parse_args(argv)
    d, num_nodes = write_temporal(tempfile.mkdtemp(), msg_dim=a.msg_dim)
    gl.set_default_neighbor_id(-1)
    gl.set_padding_mode(gl.REPLICATE)
    dec = gl.Decoder(attr_types=["float"] * a.msg_dim, timestamped=True)
    g = gl.Graph().node(os.path.join(d, "src.tsv"), "src", decoder=gl.Decoder()) \
        .node(os.path.join(d, "dst.tsv"), "dst", decoder=gl.Decoder())
    for name, et in (("events", "interaction"), ("train", "train"), ("val", "val"), ("test", "test")):
        g.edge(os.path.join(d, name + ".tsv"), ("src", "dst", et), decoder=dec, directed=False)
    g.init(device=a.device)
    model = models.TGN(num_nodes, a.msg_dim, 32, 16, 32).to(g.device)
    opt = torch.optim.Adam(model.parameters(), lr=2e-3)
    first = last = val = None
    for ep in range(a.epochs):
        model.train()
        model.memory.reset_state()
        loss = run_epoch(g, model, "train", a, num_nodes, opt)
        model.eval()
        val = run_epoch(g, model, "val", a, num_nodes)
        first = loss if first is None else first
        last = loss
        print("epoch %d loss %.4f val AUC %.3f" % (ep, loss, val))
    return first, last, val


if __name__ == "__main__":
    main()
