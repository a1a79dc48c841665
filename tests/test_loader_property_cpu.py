"""Property tests of the native TSV parser (csrc/host_loader.cpp) against a Python oracle: random tables with
negative ids, scientific-notation floats, CRLF line ends, missing trailing newline, optional headers and any
part count must parse to exactly the same columns."""
import os

import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

from graphlearn_b200.parallel.runtime import native

floats = st.one_of(st.floats(min_value=-1e6, max_value=1e6, allow_nan=False, width=32),
                   st.sampled_from([0.0, -0.0, 1e-5, 3.5e3, -2.25e-3, 7.0]))
ids = st.integers(min_value=-2 ** 40, max_value=2 ** 40)
words = st.text(alphabet="abcXYZ019_-.", min_size=0, max_size=6)


def _fmt(x, sci):
    return ("%e" % x) if sci else repr(float(np.float32(x)))


@settings(max_examples=40, deadline=None)
@given(rows=st.lists(st.tuples(ids, ids, floats, st.integers(-5, 5), st.integers(0, 10 ** 9), st.integers(-9, 9), floats, words),
                     min_size=0, max_size=60),
       header=st.booleans(), crlf=st.booleans(), trailing_nl=st.booleans(), sci=st.booleans(),
       parts=st.integers(min_value=1, max_value=5), threads=st.integers(min_value=1, max_value=4))
def test_edge_table_roundtrip(tmp_path_factory, rows, header, crlf, trailing_nl, sci, parts, threads):
    d = tmp_path_factory.mktemp("ld")
    p = os.path.join(str(d), "e.tsv")
    nl = "\r\n" if crlf else "\n"
    lines = []
    if header:
        lines.append("src_id:int64\tdst_id:int64\tweight:float\tlabel:int32\ttimestamp:int64\tfeature:string")
    for s, t, w, l, ts, ia, fa, sa in rows:
        lines.append("%d\t%d\t%s\t%d\t%d\t%d:%s:%s" % (s, t, _fmt(w, sci), l, ts, ia, _fmt(fa, sci), sa))
    body = nl.join(lines) + (nl if (trailing_nl and lines) else "")
    with open(p, "w", newline="") as f:
        f.write(body)
    C = native()
    got = [C.load_table(p, True, True, True, True, [0, 1, 2], [], ":", "\t", threads, i, parts) for i in range(parts)]
    cat = lambda j: torch.cat([g[j] for g in got])  # noqa: E731
    n = len(rows)
    assert cat(0).tolist() == [r[0] for r in rows] and cat(1).tolist() == [r[1] for r in rows]
    assert np.allclose(cat(2).numpy(), np.array([np.float32(r[2]) for r in rows], dtype=np.float32), rtol=1e-5, atol=1e-30)
    assert cat(3).tolist() == [r[3] for r in rows] and cat(4).tolist() == [r[4] for r in rows]
    assert cat(5).reshape(-1).tolist() == [r[5] for r in rows]
    assert np.allclose(cat(6).reshape(-1).numpy(), np.array([np.float32(r[6]) for r in rows], dtype=np.float32), rtol=1e-5, atol=1e-30)
    strs = []
    for g in got:
        blob, off = g[7].numpy().tobytes(), g[8].tolist()
        strs += [blob[off[i]:off[i + 1]].decode() for i in range(len(off) - 1)]
    assert strs == [r[7] for r in rows] and len(strs) == n
