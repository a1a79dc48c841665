"""Secondary BASELINE.json configurations (not the headline bench; GPU only, run under gpurun / torchrun):

  sage3     papers100M-shaped (scaled to fit the chosen node count) 3-layer GraphSAGE, fan-out 15,10,5, batch 1024/GPU:
            training steps/s of the CUDA-graph engine (sampling dominated: 15*10*5 = 750 leaves per seed)
  deepwalk  DeepWalk corpus generation: random_walk(length 40) from 8192 seeds per launch + 5 uniform negatives per
            walk position: walks/s and sampled ids/s of the resident-walker kernel over (peer) adjacency rows

  python tools/bench_configs.py sage3 [--nodes 10000000 --edges 150000000 --dim 128 --steps 300]
  torchrun --nproc-per-node 8 tools/bench_configs.py deepwalk
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from graphlearn_b200.parallel.runtime import init  # noqa: E402
from graphlearn_b200.store.synthetic import make_sharded_graph  # noqa: E402


def timed(rt, fn, steps, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize(); rt.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize(); rt.barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=rt.device, dtype=torch.float64)
    if rt.world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["sage3", "deepwalk"])
    ap.add_argument("--nodes", type=int, default=10_000_000)
    ap.add_argument("--edges", type=int, default=150_000_000)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--steps", type=int, default=300)
    a = ap.parse_args()
    rt = init()
    assert rt.is_cuda
    W = rt.world
    nodes, csr = make_sharded_graph(rt, a.nodes, a.edges, a.dim, 172, feature_dtype=torch.bfloat16, seed=1)
    out = {"config": a.what, "n_gpus": W, "num_nodes": a.nodes, "num_edges": a.edges, "feat_dim": a.dim}
    if a.what == "sage3":
        from graphlearn_b200.engine.fast_sage import FastSageTrainer
        from graphlearn_b200.models.graphsage import EgoGraphSAGE
        torch.manual_seed(0)
        model = EgoGraphSAGE(a.dim, 256, 172, 3).to(rt.device)
        tr = FastSageTrainer(rt, nodes, csr, model, [15, 10, 5], 1024)
        tr.seeds.copy_(torch.randint(0, nodes.n_local, (1024,), device=rt.device) * W + rt.rank)
        tr.capture()
        ms = timed(rt, tr.step_device, a.steps)
        out.update(metric="3-layer GraphSAGE fanout 15,10,5 train steps/s (whole job)", value=W * 1e3 / ms, ms_per_step=ms)
    else:
        from graphlearn_b200.ops import rng as rng_ops
        from graphlearn_b200.ops import walk as WK
        from graphlearn_b200.parallel.runtime import native
        B, L, NEG = 8192, 40, 5
        rng = rng_ops.DeviceRng(rt, 7)
        seeds = torch.randint(0, nodes.n_local, (B,), device=rt.device) * W + rt.rank
        off = torch.tensor([0] + [int(x) for x in torch.tensor(nodes.nrows).cumsum(0)], device=rt.device)
        total = int(off[-1])

        def step():
            w = WK.random_walk(csr, seeds, L, rng=rng, salt=1)
            native().negative_sample(csr.desc, w.reshape(-1), NEG, None, off, total, False, 0, 1024, rng.state, 2)
            rng.advance(1)
        ms = timed(rt, step, a.steps)
        out.update(metric="DeepWalk walks/s (length 40, + 5 negatives per position)", value=W * B * 1e3 / ms, ms_per_launch=ms,
                   sampled_ids_per_s=W * B * L * (1 + NEG) * 1e3 / ms)
    if rt.rank == 0:
        print(json.dumps(out))
    rt.shutdown()


if __name__ == "__main__":
    main()
