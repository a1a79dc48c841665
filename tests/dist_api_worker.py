"""SPMD worker: the PUBLIC API (gl.Graph from TSV files, GSL, samplers, lookups) on >= 2 ranks.
Each rank only stores its hash partition; results must equal the closed-form fixture."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import graphlearn_b200 as gl
from tests import fixtures as fx


def main():
    d = sys.argv[1]
    dev = "cpu" if os.environ.get("GLB_TEST_DEVICE", "") == "cpu" else None
    g = gl.Graph()
    g.node(os.path.join(d, "user.tsv"), "user",
           decoder=gl.Decoder(weighted=True, labeled=True, attr_types=["int", "int", "string", "float"]))
    g.node(os.path.join(d, "item_parts"), "item", decoder=gl.Decoder(attr_types=["float"] * 4))   # directory source
    g.edge(os.path.join(d, "u2i.tsv"), ("user", "item", "buy"), decoder=gl.Decoder(weighted=True))
    g.edge(os.path.join(d, "i2i.tsv"), ("item", "item", "sim"), decoder=gl.Decoder(labeled=True, timestamped=True))
    g.node_view("item", gl.Mask.TRAIN, seed=5, nsplit=3, split_range=(0, 2))
    g.init(device=dev)
    rt = g.runtime
    W, r = rt.world, rt.rank
    assert W >= 2
    st = g.get_stats()
    assert sum(st["user"]) == fx.N_USER and sum(st["item"]) == fx.N_ITEM
    assert sum(st["buy"]) == sum((u % 5) + 1 for u in range(fx.N_USER))
    # lookups of ids owned by OTHER ranks
    ids = np.arange(fx.N_USER)
    n = g.lookup_nodes("user", ids)
    assert np.allclose(n.weights, 1.0 + ids) and (n.labels == ids % 3).all()
    assert np.allclose(n.float_attrs[:, 0], ids / 2.0) and (n.int_attrs[:, 0] == ids).all()
    assert list(n.string_attrs[:, 0]) == ["u%d" % i for i in ids]          # host-side strings of REMOTE owners too
    it = g.lookup_nodes("item", np.arange(fx.N_ITEM))
    assert np.allclose(it.float_attrs[:, 1], np.arange(fx.N_ITEM) + 0.25)
    # neighbour sampling from seeds on any rank
    adj = fx.u2i_adj()
    for strategy in ("random", "topk", "edge_weight", "random_without_replacement"):
        lay = g.neighbor_sampler(["buy", "sim"], [3, 2], strategy=strategy if strategy != "edge_weight" else "random").get(ids)
        n1 = lay.layer_nodes(1).ids
        for u in ids:
            assert set(n1[u].tolist()) <= {x[0] for x in adj[u]}
        n2 = lay.layer_nodes(2).ids
        dd = (n2 - n1.reshape(-1)[:, None]) % fx.N_ITEM
        assert ((dd >= 1) & (dd <= 3)).all()
    assert g.out_degrees(ids, "buy").tolist() == [len(adj[u]) for u in ids]
    # a user-defined sampler (gl.register_sampler): the local rule runs on the OWNER of every source row, the framework
    # partitions the request and stitches the answers (seeds of all ranks' users asked from every rank)

    def lightest(adj_, rows, k, gen):
        end = adj_.indptr[rows + 1]
        pos = end[:, None] - 1 - torch.arange(k, device=rows.device)[None, :]
        return torch.where(pos >= adj_.indptr[rows][:, None], pos, torch.full_like(pos, -1))
    gl.register_sampler("lightest", lightest, overwrite=True)
    light = g.neighbor_sampler("buy", 2, strategy="lightest").get(ids).layer_nodes(1).ids
    for u in ids:
        by_w = sorted(adj[u], key=lambda tw: tw[1])
        pad = gl.get_config().default_neighbor_id
        assert light[u].tolist() == [t for t, _ in by_w[:2]] + [pad] * (2 - min(2, len(by_w))), (u, light[u])
    exp = np.zeros(fx.N_ITEM, int)
    for u in adj:
        for i, _ in adj[u]:
            exp[i] += 1
    assert g.in_degrees(np.arange(fx.N_ITEM), "buy").tolist() == exp.tolist()
    # GSL: every rank traverses ITS OWN users; union over ranks = all users, once per epoch
    q = g.V("user").batch(4).alias("u").outV("buy").sample(2).by("random").alias("i") \
         .outV("sim").sample(2).by("topk").alias("ii").values()
    ds = gl.Dataset(q)
    mine = []
    try:
        while True:
            res = ds.next()
            mine.extend(res["u"].ids.tolist())
            assert res["ii"].float_attrs.shape[-1] == 4
            assert (res["u"].labels == res["u"].ids % 3).all()
    except gl.OutOfRangeError:
        pass
    assert all(u % W == r for u in mine)
    allu = rt.all_gather_object(mine)
    assert sorted(sum(allu, [])) == list(range(fx.N_USER))
    # negative sampling + random walk + aggregation across ranks
    neg = g.negative_sampler("buy", 4, "in_degree").get(ids)
    for u, row in zip(ids, neg.ids):
        assert not (set(row.tolist()) & {x[0] for x in adj[u]})
    nodes = g.get_nodes("item", np.array([[1, 2, 3], [10, 20, 30]]))
    assert np.allclose(nodes.embedding_agg("mean"), nodes.float_attrs.mean(1), atol=1e-5)
    # ---- unequal shards: hash views give the ranks different numbers of seeds; with collective sampling ops the
    # epoch must end on every rank together (Dataset sync_epoch auto-on for the portable path)
    if not rt.is_cuda:
        dsv = gl.Dataset(g.V("item", mask=gl.Mask.TRAIN).batch(3).alias("s").outV("sim").sample(2).by("random").alias("n").values())
        counts = []
        for _ in range(2):
            nb = 0
            try:
                while True:
                    dsv.next(); nb += 1
            except gl.OutOfRangeError:
                pass
            counts.append(nb)
        sizes = rt.all_gather_object(int(g._store.nodes[gl.get_mask_type("item", gl.Mask.TRAIN)].n_local))
        allc = rt.all_gather_object(counts)
        assert sizes[0] != sizes[1], sizes                       # the shards really are unequal ...
        assert all(c == allc[0] for c in allc) and allc[0][0] == -(-min(sizes) // 3), (sizes, allc)   # ... the epochs equal
    # ---- ops whose portable path loops a data-dependent number of collective rounds (lock-step across ranks)
    w = gl.Dataset(g.V("item").batch(5).alias("s").random_walk("sim", 4, p=0.5, q=2.0).alias("w").values()).next()
    steps = np.concatenate([w["s"].ids[:, None], w["w"].ids], 1)
    dd = (steps[:, 1:] - steps[:, :-1]) % fx.N_ITEM
    assert ((dd >= 1) & (dd <= 3)).all()                       # every hop follows a real i -> i+k edge
    sg = gl.Dataset(g.SubGraph("item", "sim", batch_size=6).alias("sg").values()).next()["sg"]
    ei, nid = sg.edge_index, sg.nodes.ids
    assert ei.shape[0] == 2 and all(((nid[c] - nid[r_]) % fx.N_ITEM in (1, 2, 3)) or ((nid[r_] - nid[c]) % fx.N_ITEM in (1, 2, 3))
                                    for r_, c in zip(ei[0], ei[1]))
    e = g.E("buy").batch(6).alias("e")
    s_ = e.outV().alias("src")
    e.inV().alias("dst")
    s_.outNeg("buy").sample(4).by("in_degree").where("dst", condition={"float_cols": [0], "float_props": [0.5], "unique": True}).alias("neg")
    res = gl.Dataset(e.values()).next()
    for u, row in zip(res["src"].ids, res["neg"].ids):
        assert not (set(row.tolist()) & {x[0] for x in adj[int(u)]})
    qv = np.stack([np.arange(4, dtype=np.float32) * 0.25 + i for i in (3.0, 17.0)])     # == item 3 / item 17 features
    knn_ids, _ = g.search("item", qv, gl.KnnOption(k=2))
    assert knn_ids[:, 0].tolist() == [3, 17]
    # N17: replica cache of remote feature rows (partial, then full); results must not change
    for cap in (5, 10 ** 6):
        got = g._store.build_feature_caches(cap)
        n_remote_items = fx.N_ITEM - len([i for i in range(fx.N_ITEM) if i % W == r])
        assert got["item"] == min(cap, n_remote_items), got
        it2 = g.lookup_nodes("item", np.arange(fx.N_ITEM))
        assert np.allclose(it2.float_attrs, it.float_attrs)
        nodes2 = g.get_nodes("item", np.array([[1, 2, 3], [10, 20, 30]]))
        assert np.allclose(nodes2.embedding_agg("mean"), nodes.float_attrs.mean(1), atol=1e-5)
        feats = g._store.nodes["item"].feats
        cm = feats.cache_map
        if cm is None:      # CUDA full replica: per-owner local copies instead of a slot table
            assert sum(t.size(0) for t in feats.replicas) == got["item"]
        else:
            assert int((cm >= 0).sum()) == got["item"] and all(int(v) % W != r for v in torch.nonzero(cm >= 0).flatten())
    # actor-engine capability: balanced batch dispatch (gl.enable_actor) - every rank gets the same number of batches,
    # together they cover whole global rounds of the pooled seed batches, labels / attributes still resolve (remote rows)
    gl.enable_actor()
    dsb = gl.Dataset(g.V("item").batch(4).shuffle(traverse=True).alias("s").outV("sim").sample(2).by("random").alias("n").values())
    mine_b = []
    while True:
        try:
            v = dsb.next()
        except gl.OutOfRangeError:
            break
        assert v["n"].ids.shape == (4, 2) and v["s"].float_attrs.shape == (4, 4)
        assert np.allclose(v["s"].float_attrs[:, 0], v["s"].ids.astype(np.float32))          # item_float(i, 0) == i
        mine_b.append(v["s"].ids.copy())
    counts_b = rt.all_gather_object(len(mine_b))
    allb = np.concatenate([np.concatenate(x) if x else np.zeros(0, np.int64) for x in rt.all_gather_object(mine_b)])
    assert len(set(counts_b)) == 1 and counts_b[0] >= 1, counts_b
    assert len(set(allb.tolist())) == len(allb) and set(allb.tolist()) <= set(range(fx.N_ITEM))
    per_rank = [len([i for i in range(fx.N_ITEM) if i % W == q]) // 4 for q in range(W)]
    assert len(allb) == (sum(per_rank) // W) * W * 4, (len(allb), per_rank)
    rt.barrier()
    if r == 0:
        print("DIST_API_OK world=%d device=%s" % (W, rt.device))
    rt.shutdown()


if __name__ == "__main__":
    main()
