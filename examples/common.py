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
