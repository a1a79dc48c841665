"""Kernels added late in round 2 (collected last on purpose: their first driver-side run must not hide the rest of the
suite behind ``-x``).  Every test compares a CUDA kernel with a plain PyTorch oracle of the same op."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _table(rt, x, dtype=torch.float32):
    from graphlearn_b200.store.shards import IdMap, NodeTable
    t = NodeTable(rt, "t", IdMap(rt, torch.arange(x.size(0), device=rt.device), dense=True))
    t.set_float(x, dtype)
    return t


@pytest.mark.parametrize("metric", [0, 1])
@pytest.mark.parametrize("m,d", [(8, 64), (5, 30)])
def test_knn_ivfpq_scan_kernel(metric, m, d):
    """csrc/knn.cu knn_ivfpq_scan_kernel vs the torch ADC oracle on the SAME codes / tables (word and byte code loads,
    padded dims), then the public path: IVF-PQ + exact re-rank finds the true neighbours of clustered data."""
    import graphlearn_b200 as gl
    from graphlearn_b200.ops import knn
    from graphlearn_b200.parallel.runtime import init
    rt = init()
    g = torch.Generator(device=rt.device).manual_seed(11)
    n, B, k = 20000, 70, 10
    centres = torch.randn(50, d, device=rt.device, generator=g) * 4
    x = centres[torch.randint(0, 50, (n,), device=rt.device, generator=g)] + 0.3 * torch.randn(n, d, device=rt.device, generator=g)
    t = _table(rt, x)
    q = x[:B] + 0.01 * torch.randn(B, d, device=rt.device, generator=g)
    idx = knn.IvfPqIndex(t, t.feats.local, d, 32, 8, metric, m=m, refine=0)
    s_k, r_k = idx.search(q, k, True)
    s_o, r_o = idx.search(q, k, False)
    assert torch.allclose(s_k, s_o, rtol=1e-4, atol=1e-3)
    # rows with identical codes tie, so the two top-k id sets may differ at the boundary: every id the kernel returns must
    # appear in the oracle's (4x deeper) list with the score the kernel reported
    s_d, r_d = idx.search(q, 4 * k, False)
    match = r_k[:, :, None] == r_d[:, None, :]
    found = match.any(2) | (r_k < 0)
    assert found.float().mean() > 0.99, float(found.float().mean())
    s_at = torch.gather(s_d, 1, match.float().argmax(2))
    ok = found & (r_k >= 0) & match.any(2)
    assert torch.allclose(s_at[ok], s_k[ok], rtol=1e-4, atol=1e-3)
    # through the public search with the exact re-rank
    opt = gl.IndexOption(); opt.index_type = "gpu_ivfpq"; opt.nlist = 32; opt.nprobe = 16; opt.m = m
    knn.build_index(t, opt)
    ids, dist = knn.search(rt, t, q, k, metric)
    xs = t.feats.local[:, :d].float()
    sc = q @ xs.t() if metric == 1 else -torch.cdist(q, xs) ** 2
    ref_s, ref_i = torch.topk(sc, k, dim=1)
    hit = (ids[:, :, None] == ref_i[:, None, :]).any(2).float().mean()
    # 400 near-equidistant points per cluster: 8 bytes per row cannot rank inside a cluster (0.5 - 0.6 on this data, the portable
    # path gives the same numbers; tests/test_api_cpu.py::test_knn_ivfpq_index checks recall on separable data)
    assert hit > 0.35, float(hit)
    if metric == 0:
        assert (ids[:, 0] == torch.arange(B, device=rt.device)).all()
        assert torch.allclose(dist[:, 0], -ref_s[:, 0], atol=1e-2)


@pytest.mark.parametrize("mode", ["none", "ts", "weight"])
def test_csr_build_kernel_matches_stable_sorts(mode):
    """csrc/csr_build.cu (count -> scan -> scatter -> per-row bitonic sort, warp rows and CTA hub rows) must produce
    EXACTLY the permutation of the two stable global sorts of the portable build, ties included."""
    import graphlearn_b200 as gl
    from graphlearn_b200.parallel.runtime import init
    from graphlearn_b200.store.shards import CsrShard
    rt = init()
    g = torch.Generator(device=rt.device).manual_seed(3)
    n_rows, E = 5000, 400000
    src = torch.randint(0, n_rows, (E,), device=rt.device, generator=g)
    src[:3000] = 7                                  # a hub (CTA sort, not a power of two) ...
    src[3000:3300] = 11                             # ... and a row just above the warp limit
    src[src == 13] = 14                             # an empty row
    dst = torch.randint(0, 1 << 20, (E,), device=rt.device, generator=g)
    ts = torch.randint(-50, 50, (E,), device=rt.device, generator=g) if mode == "ts" else None          # many ties, negative values
    w = None
    if mode == "weight":
        w = torch.randint(-3, 4, (E,), device=rt.device, generator=g).float() * 0.5
        w[::7] = -0.0
    shards = []
    for native in (True, False):
        gl.set_native_csr_build(native)
        shards.append(CsrShard.from_coo(rt, "e", "a", "b", src, dst, n_rows, weights=w, ts=ts))
    a, b = shards
    assert torch.equal(a.indptr.local, b.indptr.local)
    assert torch.equal(a._order, b._order)
    assert torch.equal(a.indices.local, b.indices.local)
    if ts is not None:
        assert torch.equal(a.ts.local, b.ts.local)
    if w is not None:
        assert torch.equal(a.weights.local, b.weights.local) and torch.allclose(a.cumw.local, b.cumw.local)
    # empty input and a single edge
    z = torch.zeros(0, dtype=torch.int64, device=rt.device)
    gl.set_native_csr_build(True)
    e0 = CsrShard.from_coo(rt, "e", "a", "b", z, z, 4)
    assert e0.indptr.local.tolist() == [0, 0, 0, 0, 0]
    one = CsrShard.from_coo(rt, "e", "a", "b", torch.tensor([2], device=rt.device), torch.tensor([9], device=rt.device), 4,
                            weights=torch.tensor([1.5], device=rt.device))
    assert one.indptr.local.tolist() == [0, 0, 0, 1, 1] and one.indices.local.tolist() == [9]
