"""Closed-form synthetic fixtures: ids encode their own attributes / weights so
that every check is exact (same idea as graphlearn/python/tests/utils.py:44-440)."""
import os

import numpy as np

N_USER, N_ITEM = 40, 60


def user_attr(i):
    return (i, i * 10, "u%d" % i, float(i) / 2.0)          # int, int, string, float


def item_float(i, j):
    return float(i) + j * 0.25


def write_graph(d):
    """users (weighted, labeled, attrs int:int:string:float), items (4 floats),
    u2i edges (weighted; user u buys items (u*k) % N_ITEM for k=1..(u%5)+1, weight = k),
    i2i edges (timestamped + labeled)."""
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "user.tsv"), "w") as f:
        f.write("id:int64\tweight:float\tlabel:int32\tfeature:string\n")
        for i in range(N_USER):
            a = user_attr(i)
            f.write("%d\t%f\t%d\t%d:%d:%s:%f\n" % (i, 1.0 + i, i % 3, a[0], a[1], a[2], a[3]))
    with open(os.path.join(d, "item.tsv"), "w") as f:
        f.write("id:int64\tfeature:string\n")
        for i in range(N_ITEM):
            f.write("%d\t%s\n" % (i, ":".join("%f" % item_float(i, j) for j in range(4))))
    # the same item table again as a DIRECTORY of 3 part files (dealt round-robin to the ranks, N10)
    os.makedirs(os.path.join(d, "item_parts"), exist_ok=True)
    for p in range(3):
        with open(os.path.join(d, "item_parts", "part-%d" % p), "w") as f:
            f.write("id:int64\tfeature:string\n")
            for i in range(p, N_ITEM, 3):
                f.write("%d\t%s\n" % (i, ":".join("%f" % item_float(i, j) for j in range(4))))
    with open(os.path.join(d, "u2i.tsv"), "w") as f:
        f.write("src_id:int64\tdst_id:int64\tweight:float\n")
        for u in range(N_USER):
            for k in range(1, (u % 5) + 2):
                f.write("%d\t%d\t%f\n" % (u, (u * k + k) % N_ITEM, float(k)))
    with open(os.path.join(d, "i2i.tsv"), "w") as f:
        f.write("src_id:int64\tdst_id:int64\tlabel:int32\ttimestamp:int64\n")
        for i in range(N_ITEM):
            for k in range(1, 4):
                f.write("%d\t%d\t%d\t%d\n" % (i, (i + k) % N_ITEM, k, 100 * k + i))
    return d


def u2i_adj():
    adj = {}
    for u in range(N_USER):
        adj[u] = [((u * k + k) % N_ITEM, float(k)) for k in range(1, (u % 5) + 2)]
    return adj


def build_graph(d, undirected_i2i=False):
    import graphlearn_b200 as gl
    g = gl.Graph()
    g.node(os.path.join(d, "user.tsv"), "user",
           decoder=gl.Decoder(weighted=True, labeled=True, attr_types=["int", "int", "string", "float"]))
    g.node(os.path.join(d, "item.tsv"), "item", decoder=gl.Decoder(attr_types=["float"] * 4))
    g.edge(os.path.join(d, "u2i.tsv"), ("user", "item", "buy"), decoder=gl.Decoder(weighted=True))
    g.edge(os.path.join(d, "i2i.tsv"), ("item", "item", "sim"), decoder=gl.Decoder(labeled=True, timestamped=True),
           directed=not undirected_i2i)
    g.init(device="cpu")
    return g
