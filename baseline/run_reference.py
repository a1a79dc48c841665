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


def _scratch_cwd():
    d = os.path.join(os.environ.get("TMPDIR", "/tmp"), "glb_ref_logs")
    os.makedirs(d, exist_ok=True)
    return d


def _ensure_installed(rank: int) -> bool:
    """baseline/_ref is git-ignored: when the working tree was re-created from the repository it is gone.  The wheel of the
    UNMODIFIED reference that baseline/build_reference.sh produced is kept under baseline/dist/ - install it (offline, no
    dependencies, a few seconds) so that the reference arm survives; other ranks wait for rank 0."""
    marker = os.path.join(REF, "graphlearn", "__init__.py")
    if os.path.exists(marker):
        return True
    import glob
    wheels = sorted(glob.glob(os.path.join(HERE, "dist", "graph_learn-*.whl")))
    if not wheels:
        return False
    if int(os.environ.get("LOCAL_RANK", rank)) == 0:
        import subprocess
        tmp = REF + ".installing"
        subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-deps", "--quiet", "--target", tmp, wheels[-1]],
                       check=False, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        if os.path.exists(os.path.join(tmp, "graphlearn", "__init__.py")) and not os.path.exists(REF):
            os.replace(tmp, REF)                      # publish atomically: waiting ranks never see a half-installed tree
    t0 = time.time()
    while not os.path.exists(marker) and time.time() - t0 < 300:
        time.sleep(0.5)
    return os.path.exists(marker)


def main(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not _ensure_installed(rank):
        if rank == 0:
            print(json.dumps({"impl": "reference", "unavailable": "baseline/_ref is not installed (see DESIGN.md)"}))
        return
    import subprocess
    from multiprocessing import shared_memory
    import numpy as np
    import torch
    import torch.distributed as dist
    import torch.nn as nn
    import torch.nn.functional as F

    small = getattr(args, "small", False)
    if getattr(args, "config", "") == "taobao_gat":
        # the reference's bipartite GAT (graphlearn/examples/tf/ego_bipartite_*) is a TF1 model and TensorFlow is not in
        # this image; a torch re-implementation of it would be OUR model on the reference path, which the arm forbids
        print(json.dumps({"impl": "reference", "unavailable": "reference bipartite GAT is TF1-only (graphlearn/examples/tf); no TensorFlow "
                                                                "offline - only the sampling half could run unmodified"}))
        return
    cfg = getattr(args, "cfg", None) or {"shape": dict(num_nodes=2_449_029, num_edges=123_718_280, feat_dim=100, num_classes=47),
                                         "fanouts": [25, 10], "hidden": 256,
                                         "metric": "sampled-subgraph train steps/sec (2-layer GraphSAGE fanout 25,10, ogbn-products-shaped synthetic)",
                                         "model": "GraphSAGE-2layer-mean hidden256"}
    shape = dict(n_nodes=cfg["shape"]["num_nodes"], n_edges=cfg["shape"]["num_edges"])
    if small:
        shape = dict(n_nodes=200_000, n_edges=5_000_000)
    walk_len = int(cfg.get("walk_len", 0))
    dim, classes, B = cfg["shape"]["feat_dim"], cfg["shape"]["num_classes"], args.batch
    hidden, fanouts = cfg.get("hidden", 256), cfg.get("fanouts", [1])
    use_cuda = torch.cuda.is_available()
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local_rank) if use_cuda else torch.device("cpu")
    if use_cuda:
        torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl" if use_cuda else "gloo")
    root = os.environ.get("GLB_REF_DATA", "/tmp/glb_ref_data_%d_%d" % (shape["n_nodes"], shape["n_edges"]))
    sampler = os.path.join(HERE, "ref_sampler.py")
    root += "_%d" % dim
    common = [sys.executable, sampler, "--root", root, "--nodes", str(shape["n_nodes"]), "--edges",
              str(shape["n_edges"]), "--dim", str(dim), "--classes", str(classes), "--batch", str(B),
              "--fanouts", ",".join(str(f) for f in fanouts)]
    env = dict(os.environ)
    env.pop("PYTHONPATH", None)
    t0 = time.time()
    if rank == 0:
        subprocess.run(common + ["--gen-only"], check=True, env=env, stdout=subprocess.DEVNULL)
    tracker = os.path.join(root, "tracker_%s" % os.environ.get("MASTER_PORT", "0"))
    if world > 1:
        if rank == 0:
            import shutil
            shutil.rmtree(tracker, ignore_errors=True)
            os.makedirs(tracker, exist_ok=True)
        dist.barrier()
    gen_s = time.time() - t0
    # ---- the reference engine lives in its own interpreter (see ref_sampler.py for why)
    # ports of the reference servers: derived from MASTER_PORT (unique per job), kept inside 30000-50007 whatever
    # the launcher chose, and never equal to MASTER_PORT itself
    base_port = 30000 + (int(os.environ.get("MASTER_PORT", "29500")) + 211) % 20000
    hosts = ",".join("127.0.0.1:%d" % (base_port + r) for r in range(world)) if world > 1 else ""
    if walk_len > 0:
        # DeepWalk config: sampling only, timed inside the reference's own process (no tensors to hand over)
        out = subprocess.run(common + ["--rank", str(rank), "--world", str(world), "--tracker", tracker, "--hosts", hosts,
                                       "--walk-len", str(walk_len), "--neg", str(cfg.get("neg", 5)), "--steps", str(args.steps),
                                       "--warmup", str(max(args.warmup, 3))], env=env, cwd=_scratch_cwd(), capture_output=True, text=True)
        line = [l for l in out.stdout.splitlines() if l.startswith("WALK")]
        if not line:
            raise RuntimeError("reference walk run failed: " + out.stderr[-500:])
        dt, load_s = float(line[0].split()[1]), float(line[0].split()[2])
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t[0])
        if rank == 0:
            v = world * B * args.steps / dt
            print(json.dumps({"impl": "reference", "metric": cfg["metric"], "value": v, "unit": "walks/s", "n_gpus": world,
                              "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": dt * 1e3 / args.steps,
                              "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64 ids",
                              "data": "synthetic TSV of the same shape",
                              "config": {"model": "reference RandomWalk op + RandomNegativeSampler via GSL (CPU)", "global_batch": B * world,
                                         "walk_len": walk_len, "num_nodes": shape["n_nodes"], "num_edges": shape["n_edges"],
                                         "data_gen_s": round(gen_s, 1), "graph_load_s": round(load_s, 1)},
                              "e2e": {"value": v, "unit": "walks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                              "gpu_launches": 0}))
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    child = subprocess.Popen(common + ["--rank", str(rank), "--world", str(world), "--tracker", tracker,
                                       "--hosts", hosts],
                             stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True, bufsize=1, env=env,
                             cwd=_scratch_cwd())       # the reference writes its glog files into the cwd
    line = ""
    while not line.startswith("SHM"):
        line = child.stdout.readline()
        if not line:
            raise RuntimeError("reference sampler process died during graph load")
    _, shm_name, slot_bytes, load_s = line.split()
    slot_bytes, load_s = int(slot_bytes), float(load_s)
    shm = shared_memory.SharedMemory(name=shm_name)
    ns = [B]
    for f in fanouts:
        ns.append(ns[-1] * f)
    n_rows = sum(ns)
    fbytes = n_rows * dim * 4
    L = len(fanouts)

    def next_batch():
        while True:
            l = child.stdout.readline()
            if not l:
                raise RuntimeError("reference sampler process died")
            if l.startswith("READY"):
                s = int(l.split()[1])
                break
        off = s * slot_bytes
        x = np.ndarray((n_rows, dim), dtype=np.float32, buffer=shm.buf, offset=off)
        y = np.ndarray((B,), dtype=np.int64, buffer=shm.buf, offset=off + fbytes)
        return s, x, y

    def free_slot(s):
        child.stdin.write("FREE %d\n" % s)
        child.stdin.flush()

    class SAGE(nn.Module):
        """EgoGraphSAGE of depth L: layer l is applied to every adjacent hop pair (ego_gnn.py:58-110)."""

        def __init__(self):
            super().__init__()
            dims = [dim] + [hidden] * (L - 1) + [classes]
            self.lins = nn.ModuleList([nn.Linear(2 * dims[i], dims[i + 1]) for i in range(L)])

        def conv(self, lin, x, nb, k):
            return lin(torch.cat([x, nb.view(x.size(0), k, -1).mean(1)], 1))

        def forward(self, *xs):
            h = list(xs)
            for l in range(L):
                last = l == L - 1
                h = [self.conv(self.lins[l], h[i], h[i + 1], fanouts[i]) for i in range(L - l)]
                if not last:
                    h = [F.relu(t) for t in h]
            return h[0]

    torch.manual_seed(0)
    model = SAGE().to(dev)
    if world > 1:
        model = nn.parallel.DistributedDataParallel(model, device_ids=[local_rank] if use_cuda else None)
    opt = torch.optim.Adam(model.parameters(), lr=3e-3)
    h2d = [0]

    def step():
        s, x, y = next_batch()
        h2d[0] = x.nbytes + y.nbytes
        xt = torch.from_numpy(x).to(dev, non_blocking=False)
        yt = torch.from_numpy(y).to(dev, non_blocking=False)
        free_slot(s)
        xs, o = [], 0
        for n_ in ns:
            xs.append(xt[o:o + n_]); o += n_
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=use_cuda):
            logits = model(*xs)
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
            "metric": cfg["metric"],
            "value": v, "unit": "steps/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic (random graph of ogbn-products shape as TSV, random-init weights)",
            "config": {"model": cfg["model"] + " (plain PyTorch; PyG unavailable offline)",
                       "global_batch": B * world, "fanout": fanouts, "num_nodes": shape["n_nodes"],
                       "num_edges": shape["n_edges"], "feat_dim": dim,
                       "parallelism": "reference %s mode x%d + DDP" % ("local" if world == 1 else "worker", world),
                       "data_gen_s": round(gen_s, 1), "graph_load_s": round(load_s, 1)},
            "e2e": {"value": v, "unit": "steps/s", "h2d_bytes_per_step": h2d[0], "d2h_bytes_per_step": 4},
            "gpu_launches": 0, "final_loss": loss}))
    try:
        child.stdin.write("STOP\n")
        child.stdin.flush()
        child.wait(timeout=60)
    except Exception:
        child.kill()
    shm.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
