"""Attribute the gap between the device-only step and the end-to-end step() of FastSageTrainer."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from graphlearn_b200.engine.fast_sage import FastSageTrainer
from graphlearn_b200.models.graphsage import EgoGraphSAGE
from graphlearn_b200.parallel.runtime import init
from graphlearn_b200.store.synthetic import make_sharded_graph

rt = init()
nodes, csr = make_sharded_graph(rt, num_nodes=2_449_029, num_edges=123_718_280, feat_dim=100, num_classes=47, seed=0,
                                feature_dtype=torch.bfloat16)
model = EgoGraphSAGE(100, 256, 47, 2).to(rt.device)
tr = FastSageTrainer(rt, nodes, csr, model, [25, 10], 1024)
N = 2000
seeds = (torch.randint(0, nodes.n_local, (N + 64, 1024))).pin_memory()
tr.seeds.copy_(seeds[0]); tr.capture()
for i in range(20):
    tr.step(seeds[i])
torch.cuda.synchronize()


def timed(name, fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for i in range(N):
        fn(i)
    e1.record(); t_cpu = time.perf_counter() - t0
    torch.cuda.synchronize()
    print("%-44s %.4f ms/step (cpu enqueue %.4f ms/step)" % (name, e0.elapsed_time(e1) / N, t_cpu * 1e3 / N), flush=True)


timed("A device graph only", lambda i: tr.graph.replay())
timed("B e2e graphs alternating, no host work", lambda i: tr._e2e_graphs[i & 1].replay())
ev = [torch.cuda.Event(), torch.cuda.Event()]


def c(i):
    ev[i & 1].synchronize(); tr.graph.replay(); ev[i & 1].record()


timed("C device graph + t-2 event sync pattern", c)


def d(i):
    ev[i & 1].synchronize(); tr._h_seeds2[i & 1].copy_(seeds[i]); tr._e2e_graphs[i & 1].replay(); ev[i & 1].record()


timed("D e2e graphs + sync + host seed copy", d)
timed("E public step()", lambda i: tr.step(seeds[i]))
ev3 = [torch.cuda.Event() for _ in range(4)]


def f(i):
    ev3[i & 3].synchronize(); tr._h_seeds2[i & 1].copy_(seeds[i]); tr._e2e_graphs[i & 1].replay(); ev3[i & 3].record()


timed("F like D but sync on t-4 (diag only)", f)
