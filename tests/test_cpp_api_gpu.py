"""Public C++ API (graphlearn_b200/include/glb/api.h, csrc/api.cpp) exercised through its pybind export:
glb::api::Graph / Query / Dataset must agree with plain torch references on the same data."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cg():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from graphlearn_b200.parallel.runtime import native
    C = native()
    g = torch.Generator().manual_seed(5)
    n_u, n_i, e = 500, 300, 6000
    src = torch.randint(0, n_u, (e,), generator=g)
    dst = torch.randint(0, n_i, (e,), generator=g)
    w = torch.rand(e, generator=g) + 0.1
    fu = torch.randn(n_u, 10, generator=g)
    fi = torch.randint(-4, 5, (n_i, 7), generator=g).float()
    G = C.CppGraph(0, 11)
    G.add_nodes("user", n_u, fu, torch.arange(n_u) % 5)
    G.add_nodes("item", n_i, fi, None, None, True)          # bf16 storage (integers: exact)
    G.add_edges("buy", "user", "item", src, dst, w)
    G.init()
    return C, G, dict(src=src, dst=dst, w=w, fu=fu, fi=fi, n_u=n_u, n_i=n_i)


def _adj(d):
    adj = [[] for _ in range(d["n_u"])]
    for s, t, w in zip(d["src"].tolist(), d["dst"].tolist(), d["w"].tolist()):
        adj[s].append((t, w))
    return adj


def test_cpp_graph_operators(cg):
    C, G, d = cg
    adj = _adj(d)
    ids = torch.arange(0, d["n_u"], 3)
    nodes, edges = G.get_stats()
    assert nodes == {"user": 500, "item": 300} and edges == {"buy": 6000}
    deg = G.get_degree("buy", ids).cpu()
    assert deg.tolist() == [len(adj[i]) for i in ids.tolist()]
    for strat in ("random", "random_without_replacement", "edge_weight", "in_degree"):
        nb = G.sample_neighbors("buy", ids, 4, strat).cpu()
        assert nb.shape == (ids.numel(), 4)
        for i, row in zip(ids.tolist(), nb.tolist()):
            allowed = {t for t, _ in adj[i]}
            assert all((x in allowed) if allowed else x == 0 for x in row), (strat, i, row)
    top = G.sample_neighbors("buy", ids, 2, "topk").cpu()
    for i, row in zip(ids.tolist(), top.tolist()):
        best = sorted(adj[i], key=lambda tw: -tw[1])[:2]
        assert row[:len(best)] == [t for t, _ in best]
    vals, offs = G.full_neighbors("buy", ids)
    vals, offs = vals.cpu(), offs.cpu()
    for n, i in enumerate(ids.tolist()):
        assert sorted(vals[offs[n]:offs[n + 1]].tolist()) == sorted(t for t, _ in adj[i])
    q = torch.tensor([0, 299, 17, -1, 1000])
    f = G.lookup_nodes("item", q).cpu()
    assert torch.equal(f[:3], d["fi"][[0, 299, 17]]) and float(f[3:].abs().max()) == 0.0
    assert G.lookup_labels("user", torch.tensor([7, 499, 600])).cpu().tolist() == [2, 4, -1]
    neg = G.negative_sample("buy", ids, 3, True, False).cpu()
    for i, row in zip(ids.tolist(), neg.tolist()):
        allowed = {t for t, _ in adj[i]}
        assert all(0 <= x < d["n_i"] for x in row)
        if len(allowed) < 100:
            assert not (set(row) & allowed)
    assert G.negative_sample("buy", ids, 3, True, True).shape == (ids.numel(), 3)


def test_cpp_query_dataset_epochs(cg):
    C, G, d = cg
    adj = _adj(d)
    q = C.CppQuery.V("user", "u").batch(64).shuffle(True).outV("buy", 3, "random", "i").with_features(True)
    ds = C.CppDataset(G, q, 3, False)
    assert ds.batches_per_epoch == 8
    for epoch in range(2):
        seen = []
        n_b = 0
        while True:
            b = ds.next()
            if b is None:
                break
            size, ids, feats, labels = b
            torch.cuda.synchronize()
            seeds = ids[0][:size].cpu()
            hop = ids[1].view(64, 3)[:size].cpu()
            for s, row in zip(seeds.tolist(), hop.tolist()):
                allowed = {t for t, _ in adj[s]}
                assert all((x in allowed) if allowed else x == -1 for x in row)
            assert torch.allclose(feats[0][:size].cpu(), d["fu"][seeds], atol=1e-6)
            ok = hop.reshape(-1) >= 0
            assert torch.equal(feats[1].view(64 * 3, -1)[:size * 3].cpu()[ok], d["fi"][hop.reshape(-1)[ok]])
            assert labels[:size].cpu().tolist() == [s % 5 for s in seeds.tolist()]
            seen += seeds.tolist()
            n_b += 1
        assert n_b == 8 and sorted(seen) == list(range(d["n_u"])) and ds.epoch == epoch + 1
    # by_order traversal with drop_last
    ds2 = C.CppDataset(G, C.CppQuery.V("user").batch(128).shuffle(False).outV("buy", 2, "topk"), 2, True)
    got = []
    while True:
        b = ds2.next()
        if b is None:
            break
        got += b[1][0][:b[0]].cpu().tolist()
    assert got == list(range(384))
