"""SEAL-style link prediction (graphlearn/examples/tf/seal): per edge, induce the enclosing
subgraph with DRNL labels (SubGraph sampler, need_dist) and classify it with a sparse GNN."""
import tempfile

import torch

from common import write_citation_like  # noqa: E402

import graphlearn_b200 as gl
from graphlearn_b200 import models
from graphlearn_b200.ops import subgraph as SUB


def main(steps=30, device=None):
    node_f, edge_f, dim, _ = write_citation_like(tempfile.mkdtemp(), n=600)
    g = gl.Graph().node(node_f, "i", decoder=gl.Decoder(labeled=True, attr_types=["float"] * dim)) \
        .edge(edge_f, ("i", "i", "e"), decoder=gl.Decoder(weighted=True), directed=False).init(device=device)
    es = g.edge_sampler("e", batch_size=8, strategy="shuffle")
    model = models.SEAL(dim, 32, num_layers=2).to(g.device)
    opt = torch.optim.Adam(model.parameters(), lr=5e-3)
    losses = []
    for it in range(steps):
        try:
            e = es.get()
        except gl.OutOfRangeError:
            continue
        logits, labels = [], []
        src_v, dst_v = g.to_vids("i", e.src_ids), g.to_vids("i", e.dst_ids)
        neg_v = g.to_vids("i", torch.randint(0, 600, (len(e.src_ids),)))
        for s, d, y in list(zip(src_v, dst_v, [1.0] * len(src_v))) + list(zip(src_v, neg_v, [0.0] * len(src_v))):
            sg = SUB.induce_subgraph(g.store, "e", torch.stack([s, d]), [4], need_dist=True, src=s.view(1), dst=d.view(1))
            x = g.lookup_nodes("i", g.to_ids("i", sg["nodes"]), vids=sg["nodes"]).tensor("float_attrs")
            z = models.drnl_node_labeling(sg["dist_to_src"], sg["dist_to_dst"])
            logits.append(model(x, torch.stack([sg["row"], sg["col"]]), z))
            labels.append(y)
        loss = torch.nn.functional.binary_cross_entropy_with_logits(torch.stack(logits), torch.tensor(labels, device=g.device))
        opt.zero_grad(); loss.backward(); opt.step()
        losses.append(float(loss))
    print("SEAL loss %.4f -> %.4f" % (losses[0], sum(losses[-5:]) / 5))
    return losses


if __name__ == "__main__":
    main()
