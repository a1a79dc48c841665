"""Reference arm of bench.py: the UNMODIFIED alibaba/graph-learn build installed in
baseline/_ref, driven through its own public API (gl.Graph / GSL / gl.Dataset in
local or worker mode) on the same metric and config as our arm: sampled-subgraph
train steps/sec, 2-layer GraphSAGE fan-out 25,10, ogbn-products-shaped synthetic.

The reference ships no PyTorch model except a PyG GCN example and PyG is not
installed offline, so - as BASELINE.md prescribes - the model is a plain-PyTorch
GraphSAGE (nn.Linear on cuBLAS, DDP/NCCL for N > 1) fed by the reference's
samplers and feature lookups.  Nothing from graphlearn_b200 is imported here.
"""
from __future__ import annotations

import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def _gen_data(root, n_nodes, n_edges, dim, classes, seed=0):
    """ogbn-products-shaped random graph as the reference's TSV dialect (cached on disk)."""
    import numpy as np
    import pyarrow as pa
    import pyarrow.csv as pacsv
    os.makedirs(root, exist_ok=True)
    done = os.path.join(root, "DONE")
    node_f, edge_f = os.path.join(root, "node.tsv"), os.path.join(root, "edge.tsv")
    if os.path.exists(done):
        return node_f, edge_f
    rs = np.random.RandomState(seed)
    lut = np.array(["%.2f" % (i / 100.0) for i in range(-400, 401)], dtype=object)
    with open(node_f, "w") as f:
        f.write("id:int64\tlabel:int32\tfeature:string\n")
        step = 100_000
        for s in range(0, n_nodes, step):
            e = min(n_nodes, s + step)
            q = np.clip((rs.randn(e - s, dim) * 100).astype(np.int64), -400, 400) + 400
            lab = rs.randint(0, classes, e - s)
            strs = lut[q]
            f.write("".join("%d\t%d\t%s\n" % (s + i, lab[i], ":".join(strs[i])) for i in range(e - s)))
    # skewed out-degree like our synthetic generator
    w = np.exp(rs.randn(n_nodes))
    deg = np.floor(w / w.sum() * n_edges).astype(np.int64)
    rem = n_edges - int(deg.sum())
    deg += np.bincount(rs.randint(0, n_nodes, rem), minlength=n_nodes)
    src = np.repeat(np.arange(n_nodes, dtype=np.int64), deg)
    dst = rs.randint(0, n_nodes, n_edges).astype(np.int64)
    tbl = pa.table({"src_id:int64": src, "dst_id:int64": dst})
    pacsv.write_csv(tbl, edge_f, write_options=pacsv.WriteOptions(delimiter="\t", quoting_style="none"))
    open(done, "w").write("ok")
    return node_f, edge_f


def main(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    try:
        sys.path.insert(0, REF)
        import graphlearn as gl
    except Exception as e:
        if rank == 0:
            print(json.dumps({"impl": "reference", "unavailable": "cannot import baseline/_ref graphlearn: %r" % (e,)}))
        return
    import numpy as np
    import torch
    import torch.distributed as dist
    import torch.nn as nn
    import torch.nn.functional as F

    small = getattr(args, "small", False)
    shape = dict(n_nodes=200_000, n_edges=5_000_000) if small else dict(n_nodes=2_449_029, n_edges=123_718_280)
    dim, classes, hidden, fanouts, B = 100, 47, 256, [25, 10], args.batch
    use_cuda = torch.cuda.is_available()
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local_rank) if use_cuda else torch.device("cpu")
    if use_cuda:
        torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl" if use_cuda else "gloo")
    root = os.environ.get("GLB_REF_DATA", "/tmp/glb_ref_data_%d_%d" % (shape["n_nodes"], shape["n_edges"]))
    t0 = time.time()
    if rank == 0:
        _gen_data(root, dim=dim, classes=classes, **shape)
    if world > 1:
        dist.barrier()
    node_f, edge_f = os.path.join(root, "node.tsv"), os.path.join(root, "edge.tsv")
    gen_s = time.time() - t0

    # ---- the reference's own graph engine (local mode for N=1, worker mode with a FS tracker for N>1)
    t0 = time.time()
    g = gl.Graph() \
        .node(node_f, "i", decoder=gl.Decoder(labeled=True, attr_types=["float"] * dim)) \
        .edge(edge_f, ("i", "i", "e"), decoder=gl.Decoder())
    if world == 1:
        g.init()
    else:
        tracker = os.path.join(root, "tracker_%s" % os.environ.get("MASTER_PORT", "0"))
        if rank == 0:
            import shutil
            shutil.rmtree(tracker, ignore_errors=True)
            os.makedirs(tracker, exist_ok=True)
        dist.barrier()
        g.init(task_index=rank, task_count=world, tracker=tracker)
    load_s = time.time() - t0
    q = g.V("i").batch(B).shuffle(traverse=True).alias("src") \
         .outV("e").sample(fanouts[0]).by("random").alias("h1") \
         .outV("e").sample(fanouts[1]).by("random").alias("h2").values()
    ds = gl.Dataset(q, window=10)

    class SAGE(nn.Module):
        def __init__(self):
            super().__init__()
            self.l1 = nn.Linear(2 * dim, hidden)
            self.l2 = nn.Linear(2 * hidden, classes)

        def conv(self, lin, x, nb, k):
            return lin(torch.cat([x, nb.view(x.size(0), k, -1).mean(1)], 1))

        def forward(self, x0, x1, x2):
            h0 = F.relu(self.conv(self.l1, x0, x1, fanouts[0]))
            h1 = F.relu(self.conv(self.l1, x1, x2, fanouts[1]))
            return self.conv(self.l2, h0, h1, fanouts[0])

    torch.manual_seed(0)
    model = SAGE().to(dev)
    if world > 1:
        model = nn.parallel.DistributedDataParallel(model, device_ids=[local_rank] if use_cuda else None)
    opt = torch.optim.Adam(model.parameters(), lr=3e-3)
    h2d = [0]

    def step():
        while True:
            try:
                r = ds.next()
                break
            except gl.OutOfRangeError:
                continue
        xs = [r["src"].float_attrs.reshape(-1, dim), r["h1"].float_attrs.reshape(-1, dim),
              r["h2"].float_attrs.reshape(-1, dim)]
        y = r["src"].labels.reshape(-1).astype(np.int64)
        h2d[0] = sum(x.nbytes for x in xs) + y.nbytes
        xt = [torch.from_numpy(np.ascontiguousarray(x)).to(dev, non_blocking=True) for x in xs]
        yt = torch.from_numpy(y).to(dev, non_blocking=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=use_cuda):
            logits = model(*xt)
        loss = F.cross_entropy(logits.float(), yt)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return float(loss.item())           # device -> host read of the step result

    for _ in range(max(args.warmup, 3)):
        step()
    if use_cuda:
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    if use_cuda:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    t0 = time.time()
    for _ in range(args.steps):
        loss = step()
    if use_cuda:
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    else:
        ms = (time.time() - t0) * 1e3
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    if rank == 0:
        v = world * args.steps / (ms / 1e3)
        print(json.dumps({
            "impl": "reference",
            "metric": "sampled-subgraph train steps/sec (2-layer GraphSAGE fanout 25,10, ogbn-products-shaped synthetic)",
            "value": v, "unit": "steps/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic (random graph of ogbn-products shape as TSV, random-init weights)",
            "config": {"model": "GraphSAGE-2layer-mean hidden256 (plain PyTorch; PyG unavailable offline)",
                       "global_batch": B * world, "fanout": fanouts, "num_nodes": shape["n_nodes"],
                       "num_edges": shape["n_edges"], "feat_dim": dim,
                       "parallelism": "reference %s mode x%d + DDP" % ("local" if world == 1 else "worker", world),
                       "data_gen_s": round(gen_s, 1), "graph_load_s": round(load_s, 1)},
            "e2e": {"value": v, "unit": "steps/s", "h2d_bytes_per_step": h2d[0], "d2h_bytes_per_step": 4},
            "gpu_launches": 0, "final_loss": loss}))
    try:
        ds.close()
    except Exception:
        pass
    g.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
