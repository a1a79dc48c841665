"""Every example script runs end to end on the CPU (portable path) with tiny synthetic data and
learns something.  The reference's examples have no tests at all; these double as
integration tests of GSL + samplers + nn + models."""
import importlib
import os
import sys

import pytest

EX = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples")


def _run(name, *args, **kw):
    if EX not in sys.path:
        sys.path.insert(0, EX)
    mod = importlib.import_module(name)
    import numpy as np
    import torch
    torch.manual_seed(0)                  # model initialisation: deterministic thresholds below
    np.random.seed(0)
    return mod.main(*args, **kw)


def test_ego_sage_supervised():
    assert _run("train_ego_sage", ["--device", "cpu", "--epochs", "2"]) > 0.8


@pytest.mark.parametrize("kind", ["gat", "gin"])
def test_ego_gnn(kind):
    assert _run("train_ego_gnn", ["--device", "cpu", "--model", kind, "--epochs", "2", "--nodes", "800"]) > 0.7


def test_ego_rgcn():
    assert _run("train_ego_rgcn", ["--device", "cpu", "--epochs", "3"]) > 0.8


def test_ego_tgat_no_future_leak():
    assert _run("train_ego_tgat", ["--device", "cpu", "--epochs", "3"]) > 0.8


def test_gcn_sparse_masks():
    assert _run("train_gcn_sparse", ["--device", "cpu", "--epochs", "2", "--nodes", "800"]) > 0.8


def test_unsupervised_sage_exports_embeddings(tmp_path):
    out = str(tmp_path / "emb.tsv")
    first, last, path = _run("train_unsupervised_sage", ["--device", "cpu", "--epochs", "2", "--nodes", "600", "--out", out])
    assert last < first
    lines = open(path).read().strip().split("\n")
    assert lines[0] == "id:int64\temb:string" and len(lines) == 601


def test_node2vec():
    first, last = _run("node2vec", ["--device", "cpu", "--steps", "40", "--p", "0.5", "--q", "2"])
    assert last < first


def test_bipartite_sage():
    first, last = _run("bipartite_sage", steps=25, device="cpu")
    assert last < first


def test_bipartite_gat_4head():
    first, last = _run("bipartite_sage", steps=25, device="cpu", conv="gat")
    assert last < first


def test_seal():
    out = _run("seal_link_prediction", steps=15, device="cpu")
    assert out is None or out[1] <= out[0] * 1.05


def test_ultra_gcn():
    first, last, recall = _run("ultra_gcn", ["--device", "cpu", "--epochs", "3"])
    assert last < first


def test_tgn_temporal_link_prediction():
    first, last, val_auc = _run("tgn", ["--device", "cpu", "--epochs", "3"])
    assert last < first and val_auc > 0.6


def test_export_serving_model_and_online_inference(tmp_path):
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        acc, same, out = _run("export_serving_model", ["--epochs", "2", "--nodes", "600", "--out", str(tmp_path / "m.pt")])
    assert same and acc > 0.8 and os.path.getsize(out) > 0


def test_subgraph_sage_edge_inducer():
    first, last = _run("train_subgraph_sage", ["--device", "cpu", "--epochs", "2", "--nodes", "500"])
    assert last < first


def test_basic_queries_tour(tmp_path):
    """examples/basic_queries.py (the reference's examples/basic): traversal epochs, edge iteration, multi-hop shapes, every
    sampling strategy, negatives, look-ups, sub-graphs, walks and KNN on the generated toy graph."""
    import numpy as np
    out = _run("basic_queries", ["--device", "cpu", "--data", str(tmp_path)])
    for epoch in out["node_epochs"]:
        assert epoch.tolist() == list(range(100))                     # every user exactly once per epoch
    assert out["edges"] == 1000
    assert out["multi_hop"] == {"src": (8,), "h1": (8, 3), "h2": (24, 2), "h1_float": (8, 3, 4)}
    st = out["strategies"]
    assert (st["topk"] == np.array([[109, 108, 107, 106]] * 5)).all()              # heaviest edges first
    assert all(len(set(r)) == 4 for r in st["random_without_replacement"].tolist())
    assert st["full_offsets"].tolist() == [3] * 5
    neg, cond = out["negatives"]
    assert neg.shape == (6, 4) and cond.shape == (6, 2) and ((cond >= 100) & (cond < 110)).all()
    assert out["lookups"]["int"] == [100, 105] and out["lookups"]["string"] == ["100s", "105s"] and out["lookups"]["deg"] == [10, 10, 10]
    assert out["lookups"]["stats"]["relation"] == [720]                # undirected: both directions stored
    ei_shape, walks = out["subgraph_walks"]
    assert ei_shape[0] == 2 and walks.shape == (4, 5)
    step = np.abs(np.diff(np.concatenate([walks[:, :1], walks], 1)[:, 1:], axis=1)) % 120
    assert np.isin(np.minimum(step, 120 - step), [1, 2, 3]).all()      # every hop follows a ring edge
    assert out["knn"][0][0] == 5 and abs(out["knn"][1][0]) < 1e-5


def test_u2i_online_pipeline_service_process_and_client():
    """examples/u2i_online_pipeline.py: offline training on static tables, the streaming service as its own process, query built
    and installed by the GSL client, record file bulk-loaded through /admin/load (native parser), barrier, online inference on
    EgoGraph hop tensors - the served neighbourhoods must carry enough signal to recover every user's preferred category."""
    acc, n, served = _run("u2i_online_pipeline", ["--epochs", "6"])
    assert n == 67 and served == 67 and acc > 0.85


def test_ego_data_loaders(tmp_path):
    """examples/ego_data_loader.py: the reference's supervised / unsupervised EgoSAGE loaders - one epoch per iteration, ego
    graphs by alias with the hop chain read off the query."""
    if EX not in sys.path:
        sys.path.insert(0, EX)
    import graphlearn_b200 as gl
    from common import write_citation_like
    from ego_data_loader import EgoSAGESupervisedDataLoader, EgoSAGEUnsupervisedDataLoader
    node_f, edge_f, dim, classes = write_citation_like(str(tmp_path), n=300)
    g = gl.Graph().node(node_f, "i", decoder=gl.Decoder(labeled=True, attr_types=["float"] * dim)) \
        .edge(edge_f, ("i", "i", "e"), decoder=gl.Decoder(weighted=True)).init(device="cpu")
    sup = EgoSAGESupervisedDataLoader(g, gl.Mask.NONE, "random", batch_size=64, node_type="i", edge_type="e", nbrs_num=[4, 3])
    for epoch in range(2):
        seen = 0
        for ego in sup:
            assert ego.nbr_nums == [4, 3] and ego.hop_node(1).floats.shape == (ego.src.ids.numel() * 12, dim)
            assert ego.src.labels.shape == ego.src.ids.shape
            seen += ego.src.ids.numel()
        assert seen == 300
    uns = EgoSAGEUnsupervisedDataLoader(g, gl.Mask.NONE, "random", "random", batch_size=32, node_type="i", edge_type="e", nbrs_num=[3], neg_num=2)
    uns.next()
    s, d, n = uns.src_ego, uns.dst_ego, uns.neg_dst_ego
    assert s.src.ids.shape == d.src.ids.shape == (32,) and n.src.ids.numel() == 64
    assert s.nbr_nums == d.nbr_nums == n.nbr_nums == [3] and n.hop_node(0).ids.numel() == 64 * 3
    assert uns["src"].floats.shape == (32, dim) and set(uns.data_dict) >= {"src", "dst", "neg_dst"}
    g.close()
