"""Shared helpers for the examples: tiny synthetic datasets written in the reference's TSV dialect."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def write_citation_like(d, n=2000, classes=7, dim=32, deg=6, seed=0):
    """Cora-shaped toy: class-clustered features + homophilous edges (node table: label + floats)."""
    os.makedirs(d, exist_ok=True)
    rs = np.random.RandomState(seed)
    y = rs.randint(0, classes, n)
    centers = rs.randn(classes, dim)
    x = centers[y] + 0.8 * rs.randn(n, dim)
    with open(os.path.join(d, "node.tsv"), "w") as f:
        f.write("id:int64\tlabel:int32\tfeature:string\n")
        for i in range(n):
            f.write("%d\t%d\t%s\n" % (i, y[i], ":".join("%.4f" % v for v in x[i])))
    by_class = [np.where(y == c)[0] for c in range(classes)]
    with open(os.path.join(d, "edge.tsv"), "w") as f:
        f.write("src_id:int64\tdst_id:int64\tweight:float\n")
        for i in range(n):
            for _ in range(deg):
                j = rs.choice(by_class[y[i]]) if rs.rand() < 0.8 else rs.randint(0, n)
                f.write("%d\t%d\t%.3f\n" % (i, j, rs.rand() + 0.1))
    return os.path.join(d, "node.tsv"), os.path.join(d, "edge.tsv"), dim, classes


def write_bipartite(d, n_user=300, n_item=500, dim=8, seed=0):
    os.makedirs(d, exist_ok=True)
    rs = np.random.RandomState(seed)
    with open(os.path.join(d, "user.tsv"), "w") as f:
        f.write("id:int64\tfeature:string\n")
        for i in range(n_user):
            f.write("%d\t%s\n" % (i, ":".join("%.3f" % v for v in rs.randn(dim))))
    with open(os.path.join(d, "item.tsv"), "w") as f:
        f.write("id:int64\tweight:float\tfeature:string\n")
        for i in range(n_item):
            f.write("%d\t%.3f\t%s\n" % (i, 1.0 + rs.rand(), ":".join("%.3f" % v for v in rs.randn(dim))))
    with open(os.path.join(d, "u2i.tsv"), "w") as f:
        f.write("src_id:int64\tdst_id:int64\tweight:float\n")
        for u in range(n_user):
            for i in rs.choice(n_item, 8, replace=False):
                f.write("%d\t%d\t%.3f\n" % (u, i, rs.rand() + 0.1))
    return d


def write_temporal(d, n_src=120, n_dst=80, n_events=3000, msg_dim=6, seed=0):
    """JODIE-shaped toy (graphlearn/examples/pytorch/tgn): bipartite src->dst interaction events with a
    timestamp and a message vector; users have a preferred item group so links are predictable.
    Event files are split 70/15/15 by time into train/val/test edge tables."""
    os.makedirs(d, exist_ok=True)
    rs = np.random.RandomState(seed)
    pref = rs.randint(0, 4, n_src)
    ev = []
    for t in range(n_events):
        u = rs.randint(0, n_src)
        group = pref[u] if rs.rand() < 0.85 else rs.randint(0, 4)
        i = n_src + group * (n_dst // 4) + rs.randint(0, n_dst // 4)
        msg = rs.randn(msg_dim) * 0.3
        msg[group % msg_dim] += 1.0
        ev.append((u, i, t + 1, msg))
    with open(os.path.join(d, "src.tsv"), "w") as f:
        f.write("id:int64\n")
        for u in range(n_src):
            f.write("%d\n" % u)
    with open(os.path.join(d, "dst.tsv"), "w") as f:
        f.write("id:int64\n")
        for i in range(n_src, n_src + n_dst):
            f.write("%d\n" % i)

    def dump(name, rows):
        with open(os.path.join(d, name), "w") as f:
            f.write("src_id:int64\tdst_id:int64\ttimestamp:int64\tfeature:string\n")
            for u, i, t, msg in rows:
                f.write("%d\t%d\t%d\t%s\n" % (u, i, t, ":".join("%.4f" % v for v in msg)))
    a, b = int(0.7 * n_events), int(0.85 * n_events)
    dump("events.tsv", ev)
    dump("train.tsv", ev[:a])
    dump("val.tsv", ev[a:b])
    dump("test.tsv", ev[b:])
    return d, n_src + n_dst


def write_hetero(d, n=600, classes=4, dim=16, seed=0):
    """Two relation types over one node type (ego_rgcn): relation "a" is homophilous, "b" is noise."""
    os.makedirs(d, exist_ok=True)
    rs = np.random.RandomState(seed)
    y = rs.randint(0, classes, n)
    centers = rs.randn(classes, dim)
    x = centers[y] + 1.2 * rs.randn(n, dim)
    with open(os.path.join(d, "node.tsv"), "w") as f:
        f.write("id:int64\tlabel:int32\tfeature:string\n")
        for i in range(n):
            f.write("%d\t%d\t%s\n" % (i, y[i], ":".join("%.4f" % v for v in x[i])))
    by_class = [np.where(y == c)[0] for c in range(classes)]
    for name, homo in (("a", 0.9), ("b", 0.0)):
        with open(os.path.join(d, "edge_%s.tsv" % name), "w") as f:
            f.write("src_id:int64\tdst_id:int64\n")
            for i in range(n):
                for _ in range(5):
                    j = rs.choice(by_class[y[i]]) if rs.rand() < homo else rs.randint(0, n)
                    f.write("%d\t%d\n" % (i, j))
    return d, dim, classes


def write_temporal_nodes(d, n=500, classes=3, dim=8, seed=0):
    """Timestamped node + edge tables for ego_tgat: a node's label is the majority label of its
    EARLIER neighbours, so only time-respecting aggregation can predict it."""
    os.makedirs(d, exist_ok=True)
    rs = np.random.RandomState(seed)
    y = rs.randint(0, classes, n)
    centers = rs.randn(classes, dim)
    x = centers[y] + 1.0 * rs.randn(n, dim)
    ts = np.sort(rs.randint(1, 100000, n))
    with open(os.path.join(d, "node.tsv"), "w") as f:
        f.write("id:int64\tlabel:int32\ttimestamp:int64\tfeature:string\n")
        for i in range(n):
            f.write("%d\t%d\t%d\t%s\n" % (i, y[i], ts[i], ":".join("%.4f" % v for v in x[i])))
    by_class = [np.where(y == c)[0] for c in range(classes)]
    with open(os.path.join(d, "edge.tsv"), "w") as f:
        f.write("src_id:int64\tdst_id:int64\ttimestamp:int64\n")
        for i in range(n):
            for _ in range(6):
                j = rs.choice(by_class[y[i]]) if rs.rand() < 0.85 else rs.randint(0, n)
                f.write("%d\t%d\t%d\n" % (i, j, min(ts[i], ts[j]) - 1 if rs.rand() < 0.7 else ts[i] + 5))
    return d, dim, classes
