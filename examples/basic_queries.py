"""Tour of the data-access API on a tiny generated graph - the counterpart of the reference's examples/basic
(gen_test_data.py + query_examples.py + test_local.py): node / edge traversal in every order, multi-hop neighbour sampling
with every strategy, full-neighbour truncation, negative and conditional-negative sampling, attribute / degree / stats
look-ups, sub-graph sampling, random walks and KNN.  Every section returns something checkable; ``main`` returns the dict.

    python examples/basic_queries.py --device cpu            # or cuda; torchrun --nproc-per-node N for worker mode
"""
import argparse
import os
import tempfile

import numpy as np

from common import sys  # noqa: F401  (puts the repo root on sys.path)

import graphlearn_b200 as gl


def gen_files(d, u_count=100, i_count=10):
    """users 0..99 (weighted), items 100..109 (string / int / float attributes), u-i edges to every item (weighted),
    a train subset, labelled entities with 4 floats and a ring of relations (examples/basic/gen_test_data.py)."""
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "user"), "w") as f:
        f.write("id:int64\tweight:float\n" + "".join("%d\t%f\n" % (i, i / 10.0) for i in range(u_count)))
    with open(os.path.join(d, "item"), "w") as f:
        f.write("id:int64\tfeature:string\n")
        for i in range(100, 100 + i_count):
            f.write("%d\t%ss:%d:%f:%f:x\n" % (i, i, i, i * 1.0, i * 10.0))
    with open(os.path.join(d, "u-i"), "w") as f:
        f.write("src_id:int64\tdst_id:int64\tweight:float\n")
        for i in range(u_count):
            for j in range(100, 100 + i_count):
                f.write("%d\t%d\t%f\n" % (i, j, (i + j) * 0.1))
    with open(os.path.join(d, "entity"), "w") as f:
        f.write("id:int64\tlabel:int64\tfeature:string\n")
        for i in range(120):
            f.write("%d\t%d\t%f:%f:%f:%f\n" % (i, i % 5, i * 0.1, i * 0.2, i * 0.3, i * 0.4))
    with open(os.path.join(d, "relation"), "w") as f:
        f.write("src_id:int64\tdst_id:int64\tweight:float\n")
        for i in range(120):
            for k in (1, 2, 3):
                f.write("%d\t%d\t%f\n" % (i, (i + k) % 120, float(k)))
    return d


def build(d, device):
    g = gl.Graph()
    g.node(os.path.join(d, "user"), "user", decoder=gl.Decoder(weighted=True))
    g.node(os.path.join(d, "item"), "item", decoder=gl.Decoder(attr_types=["string", "int", "float", "float", "string"]))
    g.node(os.path.join(d, "entity"), "entity", decoder=gl.Decoder(labeled=True, attr_types=["float"] * 4))
    g.edge(os.path.join(d, "u-i"), ("user", "item", "buy"), decoder=gl.Decoder(weighted=True))
    g.edge(os.path.join(d, "relation"), ("entity", "entity", "relation"), decoder=gl.Decoder(weighted=True), directed=False)
    return g.init(device=device)


def node_iterate(g):
    """V().batch().shuffle(traverse=True): every node exactly once per epoch, OutOfRangeError at the end."""
    ds = gl.Dataset(g.V("user").batch(32).shuffle(traverse=True).alias("u").values())
    seen = []
    for _ in range(2):                                   # two epochs
        epoch = []
        while True:
            try:
                epoch.append(ds.next()["u"].ids)
            except gl.OutOfRangeError:
                break
        seen.append(np.sort(np.concatenate(epoch)))
    return seen


def edge_iterate(g):
    ds = gl.Dataset(g.E("buy").batch(128).alias("e").values())
    n = 0
    while True:
        try:
            e = ds.next()["e"]
            assert ((e.src_ids + e.dst_ids) * 0.1 - e.weights).max() < 1e-3
            n += e.src_ids.size
        except gl.OutOfRangeError:
            return n


def multi_hop(g):
    q = (g.V("entity").batch(8).alias("src")
          .outV("relation").sample(3).by("edge_weight").alias("h1")
          .outV("relation").sample(2).by("random").alias("h2").values())
    r = gl.Dataset(q).next()
    return {"src": r["src"].ids.shape, "h1": r["h1"].ids.shape, "h2": r["h2"].ids.shape, "h1_float": r["h1"].float_attrs.shape}


def strategies(g):
    ids = np.arange(5)
    out = {}
    for s in ("random", "random_without_replacement", "topk", "edge_weight", "in_degree"):
        out[s] = g.neighbor_sampler("buy", expand_factor=4, strategy=s).get(ids).layer_nodes(1).ids
    full = g.neighbor_sampler("buy", expand_factor=3, strategy="full").get(ids).layer_nodes(1)     # truncated to 3 per row
    out["full_offsets"] = np.asarray(full.offsets)
    return out


def negatives(g):
    ids = np.arange(6)
    neg = g.negative_sampler("buy", expand_factor=4, strategy="random").get(ids)
    cond = g.negative_sampler("buy", expand_factor=2, strategy="random", conditional=True, unique=False, int_cols=[0],
                              int_props=[1.0]).get(ids, np.full(6, 104))
    return neg.ids, cond.ids


def lookups(g):
    nodes = g.get_nodes("item", np.array([100, 105]))
    deg = g.out_degrees(np.arange(3), "buy")
    return {"int": nodes.int_attrs.reshape(-1).tolist(), "string": np.asarray(nodes.string_attrs)[:, 0].tolist(), "deg": deg.tolist(),
            "stats": g.get_stats()}


def subgraph_and_walks(g):
    sg = g.subgraph_sampler("entity", "relation", batch_size=8, strategy="random_node").get()
    walks = gl.Dataset(g.V("entity").batch(4).alias("s").random_walk("relation", walk_len=5, p=0.5, q=2.0).alias("w").values()).next()["w"].ids
    return sg.edge_index.shape, walks


def knn(g):
    q = np.array([[0.5, 1.0, 1.5, 2.0]], dtype=np.float32)         # == entity 5
    ids, dist = g.search("entity", q, gl.KnnOption(k=3))
    return ids[0].tolist(), dist[0].tolist()


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default=None)
    ap.add_argument("--data", default="")
    a = ap.parse_args(argv)
    d = gen_files(a.data or tempfile.mkdtemp(prefix="glb_basic_"))
    g = build(d, a.device)
    out = {"node_epochs": node_iterate(g), "edges": edge_iterate(g), "multi_hop": multi_hop(g), "strategies": strategies(g),
           "negatives": negatives(g), "lookups": lookups(g), "subgraph_walks": subgraph_and_walks(g), "knn": knn(g)}
    g.close()
    for k, v in out.items():
        print(k, "->", (v if k not in ("node_epochs", "strategies") else "..."))
    return out


if __name__ == "__main__":
    main()
