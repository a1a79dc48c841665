"""Decoder-format matrix: every combination of weighted / labeled / timestamped / attributed for node and
edge tables is loaded from TSV, looked up, traversed (by_order / shuffle / random) and sampled - the role
of the reference's test_{node,edge}_{weighted,labeled,attributed,...}.py family
(graphlearn/python/tests/, closed-form fixtures in tests/utils.py)."""
import itertools
import os

import numpy as np
import pytest

import graphlearn_b200 as gl

N, DEG = 40, 3
FLAGS = list(itertools.product([False, True], repeat=4))      # weighted, labeled, timestamped, attributed


def _w(i):
    return 0.5 + 0.25 * i


def _header(kind, w, l, t, a):
    cols = ["id:int64"] if kind == "node" else ["src_id:int64", "dst_id:int64"]
    cols += (["weight:float"] if w else []) + (["label:int32"] if l else []) + (["timestamp:int64"] if t else []) + \
        (["feature:string"] if a else [])
    return "\t".join(cols)


def _write(d, nflags, eflags):
    w, l, t, a = nflags
    with open(os.path.join(d, "n.tsv"), "w") as f:
        f.write(_header("node", *nflags) + "\n")
        for i in range(N):
            row = [str(i)] + (["%f" % _w(i)] if w else []) + ([str(i % 5)] if l else []) + ([str(1000 + i)] if t else []) + \
                (["%d:%f:s%d" % (i * 2, i / 4.0, i)] if a else [])
            f.write("\t".join(row) + "\n")
    w, l, t, a = eflags
    with open(os.path.join(d, "e.tsv"), "w") as f:
        f.write(_header("edge", *eflags) + "\n")
        for i in range(N):
            for k in range(1, DEG + 1):
                j = (i + k) % N
                row = [str(i), str(j)] + (["%f" % (k + i / 100.0)] if w else []) + ([str((i + j) % 7)] if l else []) + \
                    ([str(10 * k + i)] if t else []) + (["%d:%f" % (i * N + j, k / 2.0)] if a else [])
                f.write("\t".join(row) + "\n")


def _decoder(flags, kind):
    w, l, t, a = flags
    attr = (["int", "float", "string"] if kind == "node" else ["int", "float"]) if a else None
    return gl.Decoder(weighted=w, labeled=l, timestamped=t, attr_types=attr)


@pytest.mark.parametrize("nflags", FLAGS[::3] + [FLAGS[-1]])
@pytest.mark.parametrize("eflags", FLAGS)
def test_format_matrix(tmp_path, nflags, eflags):
    d = str(tmp_path)
    _write(d, nflags, eflags)
    g = gl.Graph().node(d + "/n.tsv", "n", decoder=_decoder(nflags, "node")) \
        .edge(d + "/e.tsv", ("n", "n", "e"), decoder=_decoder(eflags, "edge")).init(device="cpu")
    ids = np.arange(N)
    # ---- node lookups
    nodes = g.lookup_nodes("n", ids)
    w, l, t, a = nflags
    if w:
        assert np.allclose(nodes.weights, [_w(i) for i in ids])
    if l:
        assert (nodes.labels == ids % 5).all()
    if t:
        assert (nodes.timestamps == 1000 + ids).all()
    if a:
        assert (nodes.int_attrs[:, 0] == ids * 2).all() and np.allclose(nodes.float_attrs[:, 0], ids / 4.0)
        assert [s for s in nodes.string_attrs[:, 0]] == ["s%d" % i for i in ids]
    # ---- one epoch of edges by order: every edge exactly once, attributes follow the edge
    ew, el, et_, ea = eflags
    q = g.E("e").batch(16).alias("e").values()
    ds = gl.Dataset(q)
    seen = []
    try:
        while True:
            e = ds.next()["e"]
            s, dd = e.src_ids, e.dst_ids
            k = (dd - s) % N
            if ew:
                assert np.allclose(e.weights, k + s / 100.0, atol=1e-5)
            if el:
                assert (e.labels == (s + dd) % 7).all()
            if et_:
                assert (e.timestamps == 10 * k + s).all()
            if ea:
                assert (e.int_attrs[:, 0] == s * N + dd).all() and np.allclose(e.float_attrs[:, 0], k / 2.0)
            seen.extend(zip(s.tolist(), dd.tolist()))
    except gl.OutOfRangeError:
        pass
    assert sorted(seen) == sorted((i, (i + k) % N) for i in range(N) for k in range(1, DEG + 1))
    # ---- shuffled node traversal covers every node once per epoch, twice over two epochs
    ds2 = gl.Dataset(g.V("n").batch(7).shuffle(traverse=True).alias("v").values())
    for _ in range(2):
        got = []
        try:
            while True:
                got.extend(ds2.next()["v"].ids.tolist())
        except gl.OutOfRangeError:
            pass
        assert sorted(got) == list(range(N))
    # ---- sampling honours the storage order: top-k = heaviest (weighted) / most recent (timestamped) first
    nb = g.neighbor_sampler("e", 2, strategy="topk").get(ids).layer_nodes(1).ids
    if et_:
        assert (nb[:, 0] == (ids + 1) % N).all()          # rows sorted by timestamp ascending: k = 1 first
    elif ew:
        assert (nb[:, 0] == (ids + DEG) % N).all()        # rows sorted by weight descending: k = DEG first
    assert ((nb - ids[:, None]) % N >= 1).all() and ((nb - ids[:, None]) % N <= DEG).all()
    g.close()
