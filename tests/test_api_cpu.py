"""CPU tests of the public API (Graph / GSL / Dataset / samplers) on closed-form
fixtures.  They exercise the portable torch path of every op - the same path that
serves as the oracle for the CUDA kernels."""
import os

import numpy as np
import pytest
import torch

import graphlearn_b200 as gl
from tests import fixtures as fx


@pytest.fixture(scope="module")
def g(tmp_path_factory):
    d = fx.write_graph(str(tmp_path_factory.mktemp("graph")))
    return fx.build_graph(d)


def test_decoder_format_bits():
    d = gl.Decoder(weighted=True, labeled=True, timestamped=True, attr_types=["int", ("string", 10), "float", "string"])
    assert d.data_format == 2 + 4 + 8 + 16
    assert (d.int_attr_num, d.float_attr_num, d.string_attr_num) == (2, 1, 1)
    assert gl.Decoder().data_format == 0


def test_stats_and_topology(g):
    st = g.get_stats()
    assert st["user"] == [fx.N_USER] and st["item"] == [fx.N_ITEM]
    assert st["buy"] == [sum((u % 5) + 1 for u in range(fx.N_USER))]
    topo = g.get_topology()
    assert topo.get_src_type("buy") == "user" and topo.get_dst_type("buy") == "item"


def test_lookup_nodes_closed_form(g):
    ids = np.array([0, 7, 39, 12])
    n = g.lookup_nodes("user", ids)
    assert np.allclose(n.weights, 1.0 + ids)
    assert (n.labels == ids % 3).all()
    assert (n.int_attrs == np.stack([ids, ids * 10], 1)).all()
    assert np.allclose(n.float_attrs[:, 0], ids / 2.0)
    assert list(n.string_attrs[:, 0]) == ["u%d" % i for i in ids]
    it = g.lookup_nodes("item", np.array([[3, 4], [5, 59]]))
    assert it.float_attrs.shape == (2, 2, 4)
    assert np.allclose(it.float_attrs[1, 1], [fx.item_float(59, j) for j in range(4)])


def test_lookup_unknown_id_defaults(g):
    gl.set_default_label(-7)
    n = g.lookup_nodes("user", np.array([5, 1000]))
    assert n.labels[1] == -7 and n.labels[0] == 5 % 3
    assert n.weights[1] == 0.0
    assert (n.float_attrs[1] == 0).all()


def test_node_sampler_epochs(g):
    s = g.node_sampler("user", batch_size=16, strategy="by_order")
    seen = []
    with pytest.raises(gl.OutOfRangeError):
        while True:
            seen.extend(s.get().ids.tolist())
    assert sorted(seen) == list(range(fx.N_USER))
    assert len(s.get().ids) == 16            # next epoch restarts
    sh = g.node_sampler("user", batch_size=10, strategy="shuffle")
    ep = []
    with pytest.raises(gl.OutOfRangeError):
        while True:
            ep.extend(sh.get().ids.tolist())
    assert sorted(ep) == list(range(fx.N_USER)) and ep != list(range(fx.N_USER))
    r = g.node_sampler("user", batch_size=7, strategy="random")
    for _ in range(20):
        assert len(r.get().ids) == 7


def test_edge_sampler(g):
    s = g.edge_sampler("buy", batch_size=8, strategy="by_order")
    e = s.get()
    adj = fx.u2i_adj()
    for u, i in zip(e.src_ids, e.dst_ids):
        assert i in [x[0] for x in adj[u]]
    assert e.weights.shape == (8,)


@pytest.mark.parametrize("strategy", ["random", "random_without_replacement", "topk", "edge_weight", "in_degree"])
def test_neighbor_sampler_membership(g, strategy):
    adj = fx.u2i_adj()
    ids = np.arange(fx.N_USER)
    layers = g.neighbor_sampler("buy", 4, strategy=strategy).get(ids)
    n = layers.layer_nodes(1)
    e = layers.layer_edges(1)
    assert n.ids.shape == (fx.N_USER, 4)
    for u in ids:
        items = [x[0] for x in adj[u]]
        assert set(n.ids[u].tolist()) <= set(items)
        if strategy == "topk":       # rows sorted by weight desc; circular padding
            exp = [x[0] for x in sorted(adj[u], key=lambda t: -t[1])]
            assert n.ids[u].tolist() == [exp[j % len(exp)] for j in range(4)]
        if strategy == "random_without_replacement" and len(items) >= 4:
            assert len(set(e.edge_ids[u].tolist())) == 4
    assert np.allclose(e.weights.shape, (fx.N_USER, 4))


def test_padding_replicate(g):
    gl.set_padding_mode(gl.REPLICATE)
    gl.set_default_neighbor_id(-1)
    layers = g.neighbor_sampler("buy", 6, strategy="topk").get(np.array([0]))     # user 0 has 1 item
    ids = layers.layer_nodes(1).ids[0]
    assert ids[0] == 1 and (ids[1:] == -1).all()


def test_full_sampler_sparse(g):
    adj = fx.u2i_adj()
    layers = g.neighbor_sampler("buy", 0, strategy="full").get(np.array([3, 4, 9]))
    n = layers.layer_nodes(1)
    assert n.offsets.tolist() == [len(adj[3]), len(adj[4]), len(adj[9])]
    parts = [x.ids.tolist() for x in n]
    for u, p in zip([3, 4, 9], parts):
        assert sorted(p) == sorted(x[0] for x in adj[u])
    assert n.dense_shape[0] == 3


def test_edge_weight_distribution(g):
    # user 4 has 5 items with weights 1..5
    layers = g.neighbor_sampler("buy", 2000, strategy="edge_weight").get(np.array([4]))
    ids = layers.layer_nodes(1).ids[0]
    adj = dict(fx.u2i_adj()[4])
    tot = sum(adj.values())
    for item, w in adj.items():
        frac = (ids == item).mean()
        assert abs(frac - w / tot) < 0.05


def test_gsl_two_hop_query(g):
    q = (g.V("user").batch(8).alias("src")
          .outV("buy").sample(3).by("random").alias("h1")
          .outV("sim").sample(2).by("topk").alias("h2").values())
    ds = gl.Dataset(q)
    res = ds.next()
    assert res["src"].ids.shape == (8,)
    assert res["h1"].ids.shape == (8, 3)
    assert res["h2"].ids.shape == (24, 2)
    assert res["h2"].float_attrs.shape == (24, 2, 4)
    h1 = res["h1"].ids.reshape(-1)
    # i2i rows are timestamp-ascending (k = 1, 2, 3): topk(2) = first two
    assert (res["h2"].ids[:, 0] == (h1 + 1) % fx.N_ITEM).all()
    assert (res["h2"].ids[:, 1] == (h1 + 2) % fx.N_ITEM).all()
    assert (res["src"].labels == res["src"].ids % 3).all()
    n = 1
    with pytest.raises(gl.OutOfRangeError):
        while True:
            ds.next(); n += 1
    assert n == fx.N_USER // 8
    assert ds.next()["src"].ids.shape == (8,)        # new epoch


def test_gsl_edges_and_each(g):
    q = (g.E("buy").batch(5).alias("e")
          .each(lambda e: (e.outV().alias("u"), e.inV().alias("i").outV("sim").sample(2).by("random").alias("ii")))
          .values())
    res = gl.Dataset(q).next()
    assert res["e"].src_ids.shape == (5,) and res["e"].weights.shape == (5,)
    assert (res["u"].ids == res["e"].src_ids).all() and (res["i"].ids == res["e"].dst_ids).all()
    assert res["ii"].ids.shape == (5, 2)


def test_gsl_negative_and_filter(g):
    q = (g.E("buy").batch(6).alias("e").each(lambda e: (
        e.outV().alias("u").outNeg("buy").sample(4).by("in_degree").alias("neg"),
        e.inV().alias("i"))).values())
    res = gl.Dataset(q).next()
    adj = fx.u2i_adj()
    for u, negs in zip(res["u"].ids, res["neg"].ids):
        assert not (set(negs.tolist()) & {x[0] for x in adj[u]})
    q2 = g.V("item").batch(10).alias("a").outV("sim").sample(2).by("random").filter("a").alias("b").values()
    r2 = gl.Dataset(q2).next()
    assert (r2["b"].ids != r2["a"].ids[:, None]).all()


def test_random_walk_and_node2vec(g):
    for p, qq in ((1.0, 1.0), (0.25, 4.0)):
        q = g.V("item").batch(12).alias("s").random_walk("sim", 5, p=p, q=qq).alias("w").values()
        res = gl.Dataset(q).next()
        w = res["w"].ids
        assert w.shape == (12, 5)
        cur = res["s"].ids
        for s in range(5):
            d = (w[:, s] - cur) % fx.N_ITEM
            assert ((d >= 1) & (d <= 3)).all()
            cur = w[:, s]


def test_subgraph_induce(g):
    sg = g.subgraph_sampler("item", "sim", batch_size=10, strategy="by_order").get()
    ids = sg.nodes.ids
    assert ids.tolist() == list(range(10))
    ei = sg.edge_index
    pairs = set(zip(ids[ei[0]].tolist(), ids[ei[1]].tolist()))
    for a in range(10):
        for k in (1, 2, 3):
            b = a + k
            if b < 10:
                assert (a, b) in pairs and (b, a) in pairs
    assert ei.shape[1] == 2 * sum(1 for a in range(10) for k in (1, 2, 3) if a + k < 10)


def test_degrees_and_aggregation(g):
    adj = fx.u2i_adj()
    od = g.out_degrees(np.arange(10), "buy")
    assert od.tolist() == [len(adj[u]) for u in range(10)]
    indeg = g.in_degrees(np.arange(fx.N_ITEM), "buy")
    exp = np.zeros(fx.N_ITEM, int)
    for u in adj:
        for i, _ in adj[u]:
            exp[i] += 1
    assert indeg.tolist() == exp.tolist()
    nodes = g.get_nodes("item", np.array([[1, 2, 3], [10, 20, 30]]))
    for func, f in (("sum", np.sum), ("mean", np.mean), ("max", np.max), ("min", np.min)):
        got = nodes.embedding_agg(func)
        ref = f(nodes.float_attrs, axis=1)
        assert np.allclose(got, ref, atol=1e-5), func


def test_undirected_and_masks(tmp_path):
    d = fx.write_graph(str(tmp_path))
    g2 = gl.Graph()
    g2.node(d + "/item.tsv", "item", decoder=gl.Decoder(attr_types=["float"] * 4))
    g2.node(d + "/item.tsv", "item", decoder=gl.Decoder(attr_types=["float"] * 4), mask=gl.Mask.TRAIN)
    g2.edge(d + "/i2i.tsv", ("item", "item", "sim"), decoder=gl.Decoder(labeled=True, timestamped=True), directed=False)
    g2.init(device="cpu")
    assert g2.get_stats()["sim"] == [2 * 3 * fx.N_ITEM]
    assert g2.out_degrees(np.array([5]), "sim").tolist() == [6]
    q = g2.V("item", mask=gl.Mask.TRAIN).batch(4).alias("s").outV("sim").sample(3).by("random").alias("n").values()
    res = gl.Dataset(q).next()
    assert res["s"].type == "item" and res["n"].ids.shape == (4, 3)
    d1 = (res["n"].ids - res["s"].ids[:, None]) % fx.N_ITEM
    assert (np.isin(d1, [1, 2, 3, fx.N_ITEM - 1, fx.N_ITEM - 2, fx.N_ITEM - 3])).all()


def test_in_memory_sources_and_knn():
    g3 = gl.Graph()
    n = 200
    x = np.random.RandomState(0).randn(n, 16).astype(np.float32)
    g3.node({"ids": np.arange(n), "float_attrs": x}, "p", decoder=gl.Decoder(attr_types=["float"] * 16))
    g3.edge({"src_ids": np.arange(n), "dst_ids": (np.arange(n) + 1) % n}, ("p", "p", "next"))
    g3.init(device="cpu")
    ids, dist = g3.search("p", x[:5], gl.KnnOption(k=3))
    assert (ids[:, 0] == np.arange(5)).all() and np.allclose(dist[:, 0], 0, atol=1e-4)


def test_knn_ivfpq_index():
    """IVF-PQ (index_factory.cc 'ivfpq'): product-quantised lists + exact re-rank find the true neighbours of clustered data,
    and the raw quantised scores (refine = 0) approximate the exact distances."""
    import torch
    from graphlearn_b200.ops import knn as K
    rs = np.random.RandomState(3)
    n, d = 3000, 24
    centres = rs.randn(30, d).astype(np.float32) * 4
    x = (centres[rs.randint(0, 30, n)] + 0.3 * rs.randn(n, d)).astype(np.float32)
    g = gl.Graph()
    opt = gl.IndexOption(); opt.index_type = "ivfpq"; opt.nlist = 16; opt.nprobe = 8; opt.m = 6
    g.node({"ids": np.arange(n), "float_attrs": x}, "p", decoder=gl.Decoder(attr_types=["float"] * d), option=opt)
    g.edge({"src_ids": np.arange(n), "dst_ids": (np.arange(n) + 1) % n}, ("p", "p", "next"))
    g.init(device="cpu")
    q = x[:40] + 0.01 * rs.randn(40, d).astype(np.float32)
    ids, dist = g.search("p", q, gl.KnnOption(k=5))
    assert (ids[:, 0] == np.arange(40)).all()
    full = ((q[:, None, :] - x[None, :, :]) ** 2).sum(2)
    ref = np.argsort(full, 1)[:, :5]
    recall = np.mean([len(set(ids[i]) & set(ref[i])) / 5.0 for i in range(40)])
    assert recall > 0.8, recall
    assert np.allclose(dist[:, 0], full[np.arange(40), np.arange(40)], atol=1e-3)
    tab = g._table("p")
    idx = tab._knn_index
    assert isinstance(idx, K.IvfPqIndex) and idx.m == 6 and idx.dsub == 4 and idx.codes.shape == (n, 6)
    assert idx.code_bytes() == n * 6
    # plain PQ scores: within the quantisation error of the true (negated squared) distances
    idx.refine = 0
    s, r = idx.search(torch.from_numpy(q), 5, False)
    true = -torch.from_numpy(full)[torch.arange(40)[:, None], r.clamp(min=0)]
    assert (r >= 0).all() and float((s - true).abs().mean()) < 0.5 * float(true.abs().mean() + 1)
    # inner-product metric goes through the same tables (+ the <q, centroid> term)
    idx2 = K.IvfPqIndex(tab, tab.feats.local, d, 16, 16, 1, m=6)
    s2, r2 = idx2.search(torch.from_numpy(q), 5, False)
    ip = torch.from_numpy(q @ x.T)
    ref2 = ip.topk(5, dim=1).indices
    rec2 = np.mean([len(set(r2[i].tolist()) & set(ref2[i].tolist())) / 5.0 for i in range(40)])
    assert rec2 > 0.6, rec2
    g.close()


def test_non_dense_ids(tmp_path):
    """ids that are not 0..N-1: the id <-> virtual-id map must be transparent."""
    g4 = gl.Graph()
    ids = np.array([1000, 5, 77, 123456789012, 42])
    g4.node({"ids": ids, "labels": np.arange(5)}, "n", decoder=gl.Decoder(labeled=True))
    g4.edge({"src_ids": ids, "dst_ids": np.roll(ids, 1)}, ("n", "n", "e"))
    g4.init(device="cpu")
    lay = g4.neighbor_sampler("e", 2).get(ids)
    assert (lay.layer_nodes(1).ids[:, 0] == np.roll(ids, 1)).all()
    assert (g4.lookup_nodes("n", ids).labels == np.arange(5)).all()


def test_native_loader_byte_range_slices(tmp_path):
    """N10: the union of the `part_count` record ranges of a file is the file, in order, for any
    part count (with / without header, tiny and multi-threaded sizes)."""
    import torch
    from graphlearn_b200.parallel.runtime import native
    C = native()
    for header, n in ((True, 7), (False, 5000), (True, 40000)):
        p = str(tmp_path / ("e_%d_%d.tsv" % (header, n)))
        with open(p, "w") as f:
            if header:
                f.write("src_id:int64\tdst_id:int64\tweight:float\n")
            for i in range(n):
                f.write("%d\t%d\t%.3f\n" % (i * 7, (i * 13) % 1000, 0.5 + (i % 9)))
        full = C.load_table(p, True, True, False, False, [], [], ":", "\t", 4, 0, 1)
        assert full[0].numel() == n
        for parts in (2, 3, 8, 11):
            got = [C.load_table(p, True, True, False, False, [], [], ":", "\t", 4, i, parts) for i in range(parts)]
            for col in range(3):
                assert torch.equal(torch.cat([g[col] for g in got]), full[col]), (header, n, parts, col)


def test_node_view_splits(tmp_path):
    """node_view(): hash splits of one node type act as TRAIN / VAL / TEST masks (disjoint, complete,
    attributes come from the base table)."""
    d = fx.write_graph(str(tmp_path))
    g = gl.Graph()
    g.node(d + "/item.tsv", "item", decoder=gl.Decoder(attr_types=["float"] * 4))
    g.edge(d + "/i2i.tsv", ("item", "item", "sim"), decoder=gl.Decoder(labeled=True, timestamped=True))
    g.node_view("item", gl.Mask.TRAIN, seed=3, nsplit=10, split_range=(0, 7))
    g.node_view("item", gl.Mask.VAL, seed=3, nsplit=10, split_range=(7, 8))
    g.node_view("item", gl.Mask.TEST, seed=3, nsplit=10, split_range=(8, 10))
    g.init(device="cpu")
    got = {}
    for m in (gl.Mask.TRAIN, gl.Mask.VAL, gl.Mask.TEST):
        ds = gl.Dataset(g.V("item", mask=m).batch(16).alias("s").outV("sim").sample(2).by("random").alias("n").values())
        ids = []
        try:
            while True:
                r = ds.next()
                ids.extend(r["s"].ids.tolist())
                assert np.allclose(r["s"].float_attrs[:, 1], np.asarray(r["s"].ids) + 0.25)
        except gl.OutOfRangeError:
            pass
        got[m] = ids
    all_ids = sum(got.values(), [])
    assert sorted(all_ids) == list(range(fx.N_ITEM))
    assert len(got[gl.Mask.TRAIN]) > len(got[gl.Mask.TEST]) > 0


def test_directory_and_comma_list_sources(tmp_path):
    """A node source may be a directory of part files or a comma separated list (slice_reader.h:137-157)."""
    d = fx.write_graph(str(tmp_path))
    for src in (d + "/item_parts", ",".join(d + "/item_parts/part-%d" % p for p in range(3)), "file://" + d + "/item_parts"):
        g = gl.Graph().node(src, "item", decoder=gl.Decoder(attr_types=["float"] * 4)).init(device="cpu")
        n = g.lookup_nodes("item", np.arange(fx.N_ITEM))
        assert np.allclose(n.float_attrs[:, 1], np.arange(fx.N_ITEM) + 0.25) and g.get_stats()["item"] == [fx.N_ITEM]
        g.close()


def _cat_graph(tmp_path, n_user=30, n_item=60):
    """users buy items; every item has an int category (i % 4), a float price band and a string brand."""
    d = str(tmp_path)
    with open(d + "/u.tsv", "w") as f:
        f.write("id:int64\n" + "".join("%d\n" % u for u in range(n_user)))
    with open(d + "/i.tsv", "w") as f:
        f.write("id:int64\tfeature:string\n")
        for i in range(n_item):
            f.write("%d\t%d:%.1f:b%d\n" % (i, i % 4, float(i % 3), i % 5))
    with open(d + "/b.tsv", "w") as f:
        f.write("src_id:int64\tdst_id:int64\n")
        for u in range(n_user):
            for k in range(3):
                f.write("%d\t%d\n" % (u, (u * 2 + k * 7) % n_item))
    g = gl.Graph().node(d + "/u.tsv", "u", decoder=gl.Decoder()) \
        .node(d + "/i.tsv", "i", decoder=gl.Decoder(attr_types=["int", "float", "string"])) \
        .edge(d + "/b.tsv", ("u", "i", "buy"), decoder=gl.Decoder()).init(device="cpu")
    return g


def test_conditional_negative_sampler(tmp_path):
    """O12: negatives share the selected attribute values of the positive dst; true neighbours and the
    positive itself are excluded; batch_share excludes every positive of the batch; unique de-duplicates."""
    g = _cat_graph(tmp_path)
    src = np.arange(30)
    dst = (src * 2) % 60
    nbrs = {u: {(u * 2 + k * 7) % 60 for k in range(3)} for u in src}
    s = g.negative_sampler("buy", 8, "random", conditional=True, int_cols=[0], int_props=[1.0])
    neg = s.get(src, dst).ids
    assert neg.shape == (30, 8)
    assert (neg % 4 == (dst % 4)[:, None]).all()                       # same int category as the positive
    assert all(not (set(neg[b].tolist()) & nbrs[b]) for b in range(30))
    s2 = g.negative_sampler("buy", 6, "random", conditional=True, float_cols=[0], float_props=[0.5],
                            str_cols=[0], str_props=[0.5], unique=True)
    neg2 = s2.get(src, dst).ids
    assert (neg2[:, :3] % 3 == (dst % 3)[:, None]).all()               # float column slots
    assert (neg2[:, 3:] % 5 == (dst % 5)[:, None]).all()               # string column slots
    s3 = g.negative_sampler("buy", 8, "random", conditional=True, int_cols=[0], int_props=[0.5], batch_share=True)
    gl.set_neg_sampling_retry_times(8)
    neg3 = s3.get(src[:6], dst[:6]).ids
    assert not (set(neg3.reshape(-1).tolist()) & set(dst[:6].tolist()))   # no positive of the whole batch


def test_temporal_root_constrains_all_hops_and_in_edges(tmp_path):
    """Temporal roots: every traversal only sees edges strictly before the root element's timestamp, most recent
    first (dag_node.py:357-392, filter.cc:68-82); in-edge samples carry the ORIGINAL edge's id / timestamp / attrs."""
    d = str(tmp_path)
    n_src, n_dst, T = 6, 5, 60
    ev = [(t % n_src, 100 + (t * 3) % n_dst, t + 1, float(t)) for t in range(T)]
    with open(d + "/s.tsv", "w") as f:
        f.write("id:int64\n" + "".join("%d\n" % i for i in range(n_src)))
    with open(d + "/d.tsv", "w") as f:
        f.write("id:int64\n" + "".join("%d\n" % (100 + i) for i in range(n_dst)))
    with open(d + "/e.tsv", "w") as f:
        f.write("src_id:int64\tdst_id:int64\ttimestamp:int64\tfeature:string\n")
        f.write("".join("%d\t%d\t%d\t%.1f\n" % e for e in ev))
    gl.set_default_neighbor_id(-1)
    gl.set_padding_mode(gl.REPLICATE)
    dec = gl.Decoder(timestamped=True, attr_types=["float"])
    g = gl.Graph().node(d + "/s.tsv", "s", decoder=gl.Decoder()).node(d + "/d.tsv", "d", decoder=gl.Decoder()) \
        .edge(d + "/e.tsv", ("s", "d", "ev"), decoder=dec, directed=False).init(device="cpu")
    root = g.E("ev").batch(7).alias("e")
    root.outV().alias("src").outE("ev").sample(3).by("topk").alias("src_hist")
    root.inV().alias("dst").inE("ev").sample(3).by("topk").alias("dst_hist")
    ds = gl.Dataset(root.values())
    seen = []
    try:
        while True:
            r = ds.next()
            e, sh, dh = r["e"], r["src_hist"], r["dst_hist"]
            seen.extend(e.timestamps.tolist())
            for b in range(len(e.src_ids)):
                u, v, t = int(e.src_ids[b]), int(e.dst_ids[b]), int(e.timestamps[b])
                want = [x for x in reversed(ev) if x[0] == u and x[2] < t][:3]
                got = [(int(sh.dst_ids[b, j]), int(sh.timestamps[b, j]), float(sh.float_attrs[b, j, 0]))
                       for j in range(3) if sh.dst_ids[b, j] >= 0]
                assert got == [(x[1], x[2], x[3]) for x in want], (u, t, got, want)
                wantd = [x for x in reversed(ev) if x[1] == v and x[2] < t][:3]
                gotd = [(int(dh.dst_ids[b, j]), int(dh.timestamps[b, j]), float(dh.float_attrs[b, j, 0]))
                        for j in range(3) if dh.dst_ids[b, j] >= 0]
                assert gotd == [(x[0], x[2], x[3]) for x in wantd], (v, t, gotd, wantd)
    except gl.OutOfRangeError:
        pass
    assert seen == sorted(seen) and len(seen) == T          # edges traversed in insertion (= time) order, once


def test_gsl_where_depends_on_sibling_branch(tmp_path):
    """`outNeg(...).where(dst_alias, condition)` reads the output of a SIBLING branch: the executor must
    schedule that branch first (extra DAG edge in the reference, dag_node.py:236-304)."""
    g = _cat_graph(tmp_path)
    e = g.E("buy").batch(10).alias("e")
    s = e.outV().alias("src")                # `src` (and its children) come before `dst` in declaration order
    e.inV().alias("dst")
    s.outNeg("buy").sample(5).by("random").where("dst", condition={"int_cols": [0], "int_props": [1.0]}).alias("neg")
    r = gl.Dataset(e.values()).next()
    neg, dst = r["neg"].ids, r["dst"].ids
    assert neg.shape == (10, 5) and (neg % 4 == (dst % 4)[:, None]).all()
    assert not (neg == dst[:, None]).any()


def test_get_edges_by_endpoints_and_saint_norm(g):
    """`g.get_edges(etype, src, dst)` without edge ids resolves them from the adjacency rows (missing edge -> defaults);
    `nn.compute_norm` consumes a subgraph sampler like the reference's compute_norm."""
    adj = fx.u2i_adj()
    src = np.array([3, 3, 7, 9])
    dst = np.array([adj[3][0][0], adj[3][1][0], adj[7][0][0], 59 if 59 not in [x[0] for x in adj[9]] else 58])
    e = g.get_edges("buy", src, dst)
    assert np.allclose(e.weights[:3], [adj[3][0][1], adj[3][1][1], adj[7][0][1]])
    assert e.edge_ids[3] == -1 and e.weights[3] == gl.get_config().default_weight
    from graphlearn_b200 import nn as glnn
    node_norm, edge_norm = glnn.compute_norm(fx.N_ITEM, sum(g.get_stats()["sim"]), g.subgraph_sampler("item", "sim", batch_size=12),
                                             sample_coverage=3)
    assert node_norm.shape == (fx.N_ITEM,) and edge_norm.shape == (sum(g.get_stats()["sim"]),)
    assert float(node_norm.min()) > 0 and float(edge_norm.max()) <= 1e4


def test_negative_sampler_cache_is_per_graph(tmp_path):
    """Two graphs with the same edge type name in one process: the in-degree negative sampler of the second must not
    see the first one's distribution (the cache used to be a module global keyed by type name only)."""
    import numpy as np
    import graphlearn_b200 as gl

    def build(n_items, hot):
        src = np.arange(200) % 10
        dst = np.full(200, hot)
        g = gl.Graph()
        g.node({"ids": np.arange(10)}, "u", decoder=gl.Decoder())
        g.node({"ids": np.arange(n_items)}, "i", decoder=gl.Decoder())
        g.edge({"src_ids": src, "dst_ids": dst}, ("u", "i", "buy"), decoder=gl.Decoder())
        g.init(device="cpu")
        return g

    ga = build(100, 99)
    na = ga.negative_sampler("buy", expand_factor=4, strategy="in_degree").get(np.array([0, 1, 2]))
    ga.close()
    gb = build(20, 3)
    nb = gb.negative_sampler("buy", expand_factor=4, strategy="in_degree").get(np.array([0, 1, 2]))
    ids = np.asarray(nb.ids).reshape(-1)
    assert ids.max() < 20, ids            # id 99 only exists in graph A


def test_dgs_install_wider_query_after_growth():
    from graphlearn_b200.dgs import DynamicGraphService, QueryPlan
    import torch
    schema = {"vertices": {"u": {"count": 4}, "i": {"count": 4}}, "edges": {"buy": {"src": "u", "dst": "i"}}}
    svc = DynamicGraphService(schema, device="cpu")
    svc.install_query(1, QueryPlan("u").out("buy", 2))
    svc.apply_updates({"edges": {"buy": {"src": [10, 1], "dst": [2, 3], "ts": [5, 6]}}})
    svc.install_query(2, QueryPlan("u").out("buy", 4))          # used to raise a shape mismatch
    res = svc.run_query(2, [10])
    assert 2 in res["hops"][0]["ids"].reshape(-1).tolist()
    # hostile ids are dropped instead of wrapping / exploding the tables
    svc.apply_updates({"edges": {"buy": {"src": [-1, 1 << 40], "dst": [0, 0], "ts": [7, 8]}}})
    assert svc.stores["buy"].n < (1 << 20)


def test_registered_file_system_sources(tmp_path):
    """N9: a custom scheme (here an in-memory 'odps://'-like store) loads through the same Graph API."""
    import io as _io
    import numpy as np
    import graphlearn_b200 as gl
    from graphlearn_b200.io import FileSystem, register_file_system

    blobs = {"/t/node.tsv": b"id:int64\tfeature:string\n" + b"".join(b"%d\t%d.5:1.0\n" % (i, i) for i in range(10)),
             "/t/edges/part-0": b"src_id:int64\tdst_id:int64\n" + b"".join(b"%d\t%d\n" % (i, (i + 1) % 10) for i in range(0, 10, 2)),
             "/t/edges/part-1": b"src_id:int64\tdst_id:int64\n" + b"".join(b"%d\t%d\n" % (i, (i + 1) % 10) for i in range(1, 10, 2))}

    class MemFs(FileSystem):
        def _p(self, path):
            return path.split("://", 1)[1][len("bucket"):]

        def isdir(self, path):
            return self._p(path).rstrip("/") == "/t/edges"

        def listdir(self, path):
            return sorted("mem://bucket" + k for k in blobs if k.startswith("/t/edges/"))

        def open(self, path):
            return _io.BytesIO(blobs[self._p(path)])

    register_file_system("mem", MemFs())
    import os
    os.environ["GLB_SPOOL_DIR"] = str(tmp_path)
    g = gl.Graph().node("mem://bucket/t/node.tsv", "n", decoder=gl.Decoder(attr_types=["float", "float"])) \
        .edge("mem://bucket/t/edges", ("n", "n", "e"), decoder=gl.Decoder()).init(device="cpu")
    assert g.get_stats()["n"] == [10] and g.get_stats()["e"] == [10]
    nb = g.neighbor_sampler("e", expand_factor=1, strategy="random").get(np.arange(10)).layer_nodes(1).ids.reshape(-1)
    assert (nb == (np.arange(10) + 1) % 10).all()
    import pytest
    with pytest.raises(gl.UnimplementedError):
        gl.Graph().node("odps://project/table", "n", decoder=gl.Decoder()).init(device="cpu")


def test_server_mode_clients_drive_queries_on_a_server(tmp_path):
    """Client <-> server decoupling: the server process holds the graph, two clients (no graph, no store) ship GSL
    queries as DagDefs and pull whole batches with attributes; epochs end with OutOfRangeError; stop protocol."""
    import socket
    import threading
    import numpy as np
    fixtures = fx
    import graphlearn_b200 as gl
    d = fixtures.write_graph(str(tmp_path))
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    cluster = {"server": "127.0.0.1:%d" % port, "client_count": 2}
    gs = gl.Graph()
    gs.node(d + "/user.tsv", "user", decoder=gl.Decoder(weighted=True, labeled=True, attr_types=["int", "int", "string", "float"]))
    gs.node(d + "/item.tsv", "item", decoder=gl.Decoder(attr_types=["float"] * 4))
    gs.edge(d + "/u2i.tsv", ("user", "item", "buy"), decoder=gl.Decoder(weighted=True))
    gs.init(cluster=cluster, job_name="server", task_index=0, device="cpu")
    served = threading.Thread(target=gs.wait_for_close, daemon=True)
    served.start()
    results = {}

    def client(cid):
        g = gl.Graph().init(cluster=cluster, job_name="client", task_index=cid)
        assert g.remote
        q = g.V("user").batch(16).alias("u").outV("buy").sample(3).by("topk").alias("i").values()
        ds = gl.Dataset(q, window=2)
        seen, feats = [], None
        while True:
            try:
                v = ds.next()
            except gl.OutOfRangeError:
                break
            seen.append(v["u"].ids)
            assert v["i"].ids.shape == (len(v["u"].ids), 3) and v["i"].float_attrs.shape == (len(v["u"].ids), 3, 4)
            assert v["u"].labels.shape == v["u"].ids.shape
            feats = (v["i"].ids, v["i"].float_attrs)
        results[cid] = (np.concatenate(seen), feats)
        g.close()

    ts = [threading.Thread(target=client, args=(c,)) for c in range(2)]
    [t.start() for t in ts]
    [t.join(60) for t in ts]
    served.join(30)
    assert not served.is_alive(), "server did not shut down after both clients stopped"
    for cid in range(2):
        ids, (iid, fa) = results[cid]
        assert sorted(ids.tolist()) == list(range(fixtures.N_USER))           # every client traverses the server's users once
        ok = iid >= 0
        assert np.allclose(fa[..., 1][ok], iid[ok] + 0.25)                      # item_float(i, 1) = i + 0.25


def test_dedup_before_pull_lookup(tmp_path):
    """gl.set_dedup_feature_pull(True): attribute lookups fetch distinct rows once and expand by index - same answers"""
    import numpy as np
    import graphlearn_b200 as gl
    from graphlearn_b200.ops import gather as G
    from tests import fixtures as fx
    d = fx.write_graph(str(tmp_path))
    g = fx.build_graph(d)
    ids = np.array([[3, 3, 7], [7, 3, 1]])
    a = g.get_nodes("item", ids).float_attrs
    gl.set_dedup_feature_pull(True)
    try:
        b = g.get_nodes("item", ids).float_attrs
        tab = g.store.nodes["item"]
        v = tab.idmap.to_vid(__import__("torch").tensor([5, -1, 5, 2, -1]))
        x = G.gather_rows_dedup(g.runtime, tab.feats, tab.feat_desc, v, tab.float_dim, fill=9.0)
        y = G.gather_rows(g.runtime, tab.feats, tab.feat_desc, v, tab.float_dim, fill=9.0)
    finally:
        gl.set_dedup_feature_pull(False)
    assert np.array_equal(a, b) and bool((x == y).all())
    g.close()


def test_cpp_api_example_builds(tmp_path):
    """L7: a plain C++ program written against include/glb/api.h compiles and links against the in-tree _C.so + libtorch
    (examples/cpp; running it needs a GPU - the same entry points run under tests/test_cpp_api_gpu.py through pybind)."""
    import os
    import shutil
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if shutil.which("g++") is None or shutil.which("python3-config") is None or not os.path.exists(os.path.join(root, "graphlearn_b200", "_C.so")):
        pytest.skip("needs g++, python3-config and the built extension")
    out = str(tmp_path / "sample_and_lookup")
    p = subprocess.run(["make", "-C", os.path.join(root, "examples", "cpp"), "OUT=" + out, "PY=" + sys.executable],
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and os.path.exists(out), (p.stdout + p.stderr)[-3000:]
    sym = subprocess.run(["nm", "-D", "--undefined-only", out], capture_output=True, text=True).stdout
    assert "glb3api5Graph" in sym and "glb3api7Dataset" in sym          # resolved from _C.so at load time


def test_graph_deploy_modes_reverse_edges_and_attribute_selection(tmp_path):
    """Reference Graph methods beyond node/edge/init: add_reverse_edges (what directed=False does), deploy_in_*_mode,
    node_attributes / edge_attributes column selection (decoder rewritten, unselected columns never loaded), vineyard stubs."""
    d = fx.write_graph(str(tmp_path))
    g = gl.Graph()
    g.node(os.path.join(d, "item.tsv"), "item", decoder=gl.Decoder(attr_types=["float"] * 4))
    g.edge(os.path.join(d, "i2i.tsv"), ("item", "item", "sim"), decoder=gl.Decoder(labeled=True, timestamped=True))
    g.add_reverse_edges(("item", "item", "sim"), os.path.join(d, "i2i.tsv"), gl.Decoder(labeled=True, timestamped=True), None)
    assert g.undirected_edges == ["sim"] and not g.is_directed("sim")
    g.node_attributes("item", [2, 0], n_float=2)                   # keep float columns 2 and 0, in that order
    assert g.get_node_decoder("item").float_attr_num == 2
    with pytest.raises(ValueError):
        g.node_attributes("nobody", [0], n_float=1)
    with pytest.raises(ValueError):
        g.node_attributes("item", [0], n_int=1)                    # column 0 is a float
    os.environ["WORLD_SIZE"] = "1"
    try:
        with pytest.raises(ValueError):
            g.deploy_in_worker_mode(hosts="a:1,b:2", task_index=0)      # 2 tasks inside a 1-process torchrun job
    finally:
        del os.environ["WORLD_SIZE"]
    with pytest.raises(ValueError):
        g.deploy_in_server_mode(0, {"server": "127.0.0.1:1"}, "trainer")
    with pytest.raises(gl.UnimplementedError):
        g.vineyard({"server": "x"})
    with pytest.raises(ValueError):
        g.init_vineyard(standalone=True)
    g.deploy_in_local_mode(0) if False else g.init(device="cpu")
    ids = np.array([0, 3, 7])
    fa = g.get_nodes("item", ids).float_attrs
    want = np.array([[fx.item_float(i, 2), fx.item_float(i, 0)] for i in ids], dtype=np.float32)
    assert fa.shape == (3, 2) and np.allclose(fa, want, atol=1e-5)
    # both directions of every edge are stored once add_reverse_edges ran
    assert g.get_stats()["sim"][0] == 2 * sum(1 for _ in open(os.path.join(d, "i2i.tsv"))) - 2
    g.close()


def test_user_defined_sampler(tmp_path):
    """gl.register_sampler (the reference's REGISTER_OPERATOR + .by("xxx"), docs/en/gl/developer/operator.md): a local rule
    returning positions into the shard's CSR is usable from GSL and from the imperative sampler; slots the rule leaves empty
    or points outside the row get the default neighbour id; built-in names cannot be overridden."""
    import torch
    g = fx.build_graph(fx.write_graph(str(tmp_path)))

    def lightest(adj, rows, k, gen):
        # rows are weight-descending: the LAST k positions of a row are its lightest edges (none left: -1)
        end = adj.indptr[rows + 1]
        pos = end[:, None] - 1 - torch.arange(k, device=rows.device)[None, :]
        return torch.where(pos >= adj.indptr[rows][:, None], pos, torch.full_like(pos, -1))
    gl.register_sampler("lightest", lightest)
    try:
        assert "lightest" in gl.registered_samplers()
        with pytest.raises(ValueError):
            gl.register_sampler("lightest", lightest)
        with pytest.raises(ValueError):
            gl.register_sampler("topk", lightest)
        gl.set_default_neighbor_id(-7)
        users = np.array([1, 2, 6])
        nb = g.neighbor_sampler("buy", expand_factor=3, strategy="lightest").get(users).layer_nodes(1)
        adj = fx.u2i_adj()
        for u, row, er in zip(users.tolist(), nb.ids.tolist(), g.neighbor_sampler("buy", 3, strategy="lightest").get(users).layer_edges(1).weights.tolist()):
            by_w = sorted(adj[u], key=lambda tw: tw[1])                # (item, weight) ascending weight
            want = [t for t, _ in by_w[:3]] + [-7] * (3 - min(3, len(by_w)))
            assert row == want, (u, row, want)
        q = g.V("user").batch(4).alias("u").outV("buy").sample(2).by("lightest").alias("i").values()
        res = gl.Dataset(q).next()
        for u, row in zip(res["u"].ids.tolist(), res["i"].ids.tolist()):
            by_w = sorted(adj[u], key=lambda tw: tw[1])
            assert row == ([t for t, _ in by_w[:2]] + [-7] * (2 - min(2, len(by_w))))
        # a rule that points outside its rows cannot fabricate edges
        gl.register_sampler("rogue", lambda adj, rows, k, gen: torch.zeros(rows.numel(), k, dtype=torch.int64) + 10 ** 6)
        assert (g.neighbor_sampler("buy", 2, strategy="rogue").get(users).layer_nodes(1).ids == -7).all()
        with pytest.raises(ValueError):
            g.V("user").batch(2).outV("buy").sample(2).by("nobody")
    finally:
        gl.unregister_sampler("lightest")
        gl.unregister_sampler("rogue")
    g.close()
