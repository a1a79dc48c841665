"""Streaming sampler service (GPU-native DGS analogue) - semantics vs a Python oracle."""
import numpy as np
import torch

from graphlearn_b200.dgs import AdaptiveRateLimiter, DynamicGraphService, QueryPlan


def test_topk_by_timestamp_streaming():
    schema = {"vertices": {"u": {"count": 20, "feat_dim": 3}, "i": {"count": 30, "feat_dim": 2}},
              "edges": {"click": {"src": "u", "dst": "i"}, "sim": {"src": "i", "dst": "i"}}}
    svc = DynamicGraphService(schema, device="cpu")
    svc.install_query(1, QueryPlan("u").out("click", 4).out("sim", 2))
    rs = np.random.RandomState(0)
    oracle = {}
    t = 0
    for b in range(6):
        n = 200
        src, dst = rs.randint(0, 20, n), rs.randint(0, 30, n)
        ts = np.arange(t, t + n); t += n
        svc.apply_updates({"edges": {"click": {"src": src, "dst": dst, "ts": ts}}})
        for s, d, x in zip(src, dst, ts):
            oracle.setdefault(int(s), []).append((int(x), int(d)))
    res = svc.run_query(1, list(range(20)))
    ids, tss = res["hops"][0]["ids"], res["hops"][0]["timestamps"]
    for u in range(20):
        want = sorted(oracle[u], reverse=True)[:4]
        assert [int(x) for x in tss[u]] == [w[0] for w in want]
        assert [int(x) for x in ids[u]] == [w[1] for w in want]
    assert res["hops"][1]["ids"].shape == (80, 2) and bool((res["hops"][1]["ids"] == -1).all())   # no sim edges yet
    # vertex features: latest version wins
    svc.apply_updates({"vertices": {"i": {"id": [5, 5, 6], "ts": [1, 9, 3], "feat": [[1, 1], [2, 2], [3, 3]]}}})
    svc.apply_updates({"vertices": {"i": {"id": [5], "ts": [4], "feat": [[7, 7]]}}})
    assert svc.vstores["i"].feat[5].tolist() == [2.0, 2.0] and svc.vstores["i"].feat[6].tolist() == [3.0, 3.0]
    # checkpoint / restore
    ck = svc.checkpoint()
    svc2 = DynamicGraphService(schema, device="cpu")
    svc2.install_query(1, QueryPlan("u").out("click", 4).out("sim", 2))
    svc2.restore(ck)
    assert torch.equal(svc2.run_query(1, [3])["hops"][0]["ids"], svc.run_query(1, [3])["hops"][0]["ids"])


def test_rate_limiter():
    rl = AdaptiveRateLimiter(target_ms=20, max_concurrency=27, stable_windows=2)
    for _ in range(100):
        rl.record(50.0)
    assert rl.tick() == 9
    for w in range(2):
        for _ in range(100):
            rl.record(1.0)
        c = rl.tick()
    assert c == 11
