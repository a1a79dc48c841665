"""Microbenchmark of the fused SAGE forward kernel (big layer-1 launch) across table size,
rows-per-CTA and dtype; CUDA-event timed, fresh random ids per iteration."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from graphlearn_b200.parallel.runtime import init, native
from graphlearn_b200.ops import sage as SG
from graphlearn_b200.store.shards import IdMap, NodeTable

rt = init(); C = native(); dev = rt.device
M, k, d, n_out = 25600, 10, 100, 256
MODE = int(os.environ.get("GLB_MODE", "0"))
w = torch.randn(n_out, 256, device=dev) * 0.05
img, _ = C.pack_weight_f32(w, 256, False)
bias = torch.zeros(n_out, device=dev)
out = torch.empty(M, n_out, dtype=torch.bfloat16, device=dev)
asave = torch.empty(M, 256, dtype=torch.bfloat16, device=dev)

def run(n_nodes, dt, R, save_a=True, iters=20):
    t = NodeTable(rt, "t", IdMap(rt, torch.arange(n_nodes, device=dev), dense=True))
    t.set_float(torch.randn(n_nodes, d, device=dev), dt)
    sv = [torch.randint(0, n_nodes, (M,), device=dev) for _ in range(iters)]
    nv = [torch.randint(0, n_nodes, (M * k,), device=dev) for _ in range(iters)]
    for i in range(3):
        C.sage_fused_forward(t.feat_desc, sv[i], t.feat_desc, nv[i], M, k, 0, img, bias, 256, n_out, True, True, save_a, R, out, asave if save_a else None, None, MODE)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        C.sage_fused_forward(t.feat_desc, sv[i], t.feat_desc, nv[i], M, k, 0, img, bias, 256, n_out, True, True, save_a, R, out, asave if save_a else None, None, MODE)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / iters
    esz = 4 if dt == torch.float32 else 2
    gb = M * (k + 1) * d * esz / 1e9
    print("nodes=%8d dtype=%s R=%3d save_a=%d : %7.1f us  gather %.2f TB/s" % (n_nodes, "f32" if esz == 4 else "bf16", R, save_a, us, gb / us * 1e6 / 1e3), flush=True)
    del t

for n_nodes in (100_000, 2_449_029):
    for dt in (torch.float32, torch.bfloat16):
        for R in (128, 64, 32, 16):
            run(n_nodes, dt, R)
run(2_449_029, torch.float32, 64, save_a=False)
run(10_000_000, torch.float32, 64)
# plain gather_agg kernel (no GEMM) for comparison: high-occupancy design
from graphlearn_b200.ops import gather as G
for n_nodes in (100_000, 2_449_029):
    t = NodeTable(rt, "t", IdMap(rt, torch.arange(n_nodes, device=dev), dense=True))
    t.set_float(torch.randn(n_nodes, d, device=dev))
    nv = [torch.randint(0, n_nodes, (M * k,), device=dev) for _ in range(20)]
    for i in range(3): G.gather_agg(rt, t.feats, t.feat_desc, nv[i], d, "mean", k=k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20): G.gather_agg(rt, t.feats, t.feat_desc, nv[i], d, "mean", k=k)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / 20
    print("gather_agg kernel nodes=%d: %.1f us  %.2f TB/s" % (n_nodes, us, M * k * d * 4 / 1e9 / us * 1e3), flush=True)
    # torch index_select for reference
    e0.record()
    for i in range(20): t.feats.local[nv[i]].view(M, k, -1).mean(1)
    e1.record(); torch.cuda.synchronize()
    print("torch index+mean nodes=%d: %.1f us" % (n_nodes, e0.elapsed_time(e1) * 1000 / 20), flush=True)
