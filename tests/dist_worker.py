"""SPMD worker for the multi-rank tests (launched by torchrun; gloo on CPU, NCCL + peer kernels on GPU).
Every check compares the sharded / remote result with an oracle built from all-gathered shards."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from graphlearn_b200.parallel.runtime import init
from graphlearn_b200.store.synthetic import make_sharded_graph
from graphlearn_b200.ops import sampling as S
from graphlearn_b200.ops import gather as G
from graphlearn_b200.ops import comm as COMM
from graphlearn_b200.ops import walk as WALK


def gather_all(t):
    W = dist.get_world_size()
    n = torch.tensor([t.size(0)], device=t.device)
    ns = [torch.zeros_like(n) for _ in range(W)]
    dist.all_gather(ns, n)
    mx = int(max(int(x) for x in ns))
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[:t.size(0)] = t
    outs = [torch.zeros_like(pad) for _ in range(W)]
    dist.all_gather(outs, pad)
    return [o[:int(k)] for o, k in zip(outs, ns)]


def main():
    dev = "cpu" if os.environ.get("GLB_TEST_DEVICE", "") == "cpu" else None
    rt = init(device=dev)
    W, r = rt.world, rt.rank
    assert W >= 2
    N, E, D = 6000, 90000, 100
    nodes, csr = make_sharded_graph(rt, N, E, D, 7, seed=3, weighted=True)
    # ---- global oracle
    ips = gather_all(csr.indptr.local)
    idxs = gather_all(csr.indices.local)
    feats = gather_all(nodes.feats.local[:, :D].float())
    cache_cap = int(os.environ.get("GLB_TEST_CACHE", "0"))
    if cache_cap:      # N17 replica cache: every check below must give identical results through it
        sc = torch.bincount(csr.indices.local.clamp(min=0), minlength=max(nodes.nrows) * W).float()
        dist.all_reduce(sc)
        got = nodes.build_feature_cache(cache_cap, scores=sc)
        assert got == min(cache_cap, N - nodes.n_local), got

    def adj(v):
        o, row = v % W, v // W
        return idxs[o][int(ips[o][row]):int(ips[o][row + 1])]

    g = torch.Generator().manual_seed(100 + r)
    src = torch.randint(0, N, (300,), generator=g).to(rt.device)        # mostly REMOTE seeds
    for strat in ("random", "random_without_replacement", "topk", "edge_weight"):
        nbr, eid = S.sample_neighbors(csr, src, 6, strat)
        for b in range(300):
            a = adj(int(src[b]))
            if a.numel() == 0:
                assert (nbr[b] == 0).all()
            else:
                assert bool(torch.isin(nbr[b], a).all()), (strat, b)
    deg = S.get_degrees(csr, src)
    assert deg.tolist() == [int(adj(int(v)).numel()) for v in src]
    vals, _, offs = S.sample_full(csr, src[:50])
    for b in range(50):
        assert torch.equal(vals[offs[b]:offs[b + 1]], adj(int(src[b])))
    # ---- remote feature rows
    vids = torch.randint(0, N, (500,), generator=g).to(rt.device)
    x = G.gather_rows(rt, nodes.feats, nodes.feat_desc, vids, D)
    ref = torch.stack([feats[int(v) % W][int(v) // W] for v in vids])
    assert torch.allclose(x, ref), "remote gather mismatch"
    lab = G.gather_any(rt, nodes.labels, vids, fill=-1)
    labs = gather_all(nodes.labels.local)
    assert lab.tolist() == [int(labs[int(v) % W][int(v) // W]) for v in vids]
    agg = G.gather_agg(rt, nodes.feats, nodes.feat_desc, vids, D, "mean", k=10)
    assert torch.allclose(agg, ref.view(50, 10, D).mean(1), atol=1e-5)
    # ---- walks over remote rows
    w = WALK.random_walk(csr, src[:64], 5)
    cur = src[:64]
    for s in range(5):
        for b in range(64):
            a = adj(int(cur[b]))
            assert (int(w[b, s]) in a.tolist()) if a.numel() else int(w[b, s]) == 0
        cur = w[:, s]
    # ---- fused SAGE layer reading remote rows (CUDA only) vs oracle
    if rt.is_cuda:
        from graphlearn_b200.ops import sage as SG
        sv = torch.randint(0, N, (257,), generator=g).to(rt.device)
        nv = torch.randint(0, N, (257 * 5,), generator=g).to(rt.device)
        wgt = torch.randn(64, 2 * D, device=rt.device, generator=torch.Generator(device=rt.device).manual_seed(5)) * 0.05
        y = SG.sage_layer(SG.pad_weight(wgt, D, D, "mean"), None, k=5, mode="mean", self_table=nodes, self_vids=sv,
                          nbr_table=nodes, nbr_vids=nv).float()
        xs = torch.stack([feats[int(v) % W][int(v) // W] for v in sv])
        xn = torch.stack([feats[int(v) % W][int(v) // W] for v in nv])
        yref = SG.sage_layer_reference(wgt, None, xs, xn, 5, "mean")
        assert (y - yref).abs().max() < 0.05, float((y - yref).abs().max())
    # ---- gradient all-reduce: peer kernel (GPU) / dist (CPU) vs torch all_reduce, several epochs
    n = 4 * 1000
    ar = COMM.PeerAllReduce(rt, n)
    for it in range(5):
        gbuf = torch.randn(n, device=rt.device, generator=torch.Generator(device=rt.device).manual_seed(it * 10 + r))
        want = gbuf.clone()
        dist.all_reduce(want)
        want /= W
        ar(gbuf, average=True)
        assert torch.allclose(gbuf, want, atol=1e-6), "all-reduce mismatch at iter %d" % it
    ar.check()
    # ---- end-to-end: trainers stay in lock step and learn
    from graphlearn_b200.models.graphsage import EgoGraphSAGE
    torch.manual_seed(0)
    model = EgoGraphSAGE(D, 64, 7, 2).to(rt.device)
    if rt.is_cuda:
        from graphlearn_b200.engine.fast_sage import FastSageTrainer as T
    else:
        from graphlearn_b200.engine.trainer import SageTrainer as T
    tr = T(rt, nodes, csr, model, [5, 3], 128, lr=1e-2)
    if rt.is_cuda:
        tr.seeds.copy_((torch.randint(0, nodes.n_local, (128,), generator=g) * W + r).to(rt.device))
        tr.capture()
    losses = []
    for it in range(30):
        seeds = torch.randint(0, nodes.n_local, (128,), generator=g) * W + r
        l = tr.step(seeds)
        if rt.is_cuda:
            torch.cuda.synchronize()
        losses.append(float(l))
    assert losses[-1] < 0.8 * losses[0], losses
    ps = gather_all(tr.flat_p.reshape(-1, 1))
    for p in ps[1:]:
        assert torch.allclose(p, ps[0], atol=1e-6), "parameters diverged across ranks"
    tr.ar.check()
    rt.barrier()
    if r == 0:
        print("DIST_WORKER_OK world=%d device=%s loss %.3f -> %.3f" % (W, rt.device, losses[0], losses[-1]))
    rt.shutdown()


if __name__ == "__main__":
    main()
