"""Kernel timeline of ONE replay of the training-step CUDA graph (CUPTI through torch.profiler):
start offset, duration, stream and name of every kernel -> shows which branches really overlap."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from graphlearn_b200.engine.fast_sage import FastSageTrainer
from graphlearn_b200.models.graphsage import EgoGraphSAGE
from graphlearn_b200.parallel.runtime import init
from graphlearn_b200.store.synthetic import make_sharded_graph

rt = init()
nodes, csr = make_sharded_graph(rt, num_nodes=2_449_029, num_edges=123_718_280, feat_dim=100, num_classes=47, seed=0,
                                feature_dtype=torch.bfloat16)
model = EgoGraphSAGE(100, 256, 47, 2).to(rt.device)
tr = FastSageTrainer(rt, nodes, csr, model, [25, 10], 1024)
tr.seeds.copy_(torch.randint(0, nodes.n_local, (1024,), device=rt.device)); tr.capture()
e2e = os.environ.get("GLB_E2E", "0") == "1"
seeds = torch.randint(0, nodes.n_local, (64, 1024)).pin_memory()
for i in range(20):
    tr.step(seeds[i]) if e2e else tr.step_device()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(6):
        tr.step(seeds[20 + i]) if e2e else tr.step_device()
    torch.cuda.synchronize()
evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
# split into replays by the Adam kernel
ends = [i for i, e in enumerate(evs) if "adam_pack" in e.name]
lo, hi = ends[2] + 1, ends[3] + 1
t0 = evs[lo].time_range.start
span = evs[hi - 1].time_range.end - t0
if rt.rank == 0:
    print("timeline of replay 4 (%s), %d kernels, span %.1f us%s" % ("e2e" if e2e else "device", hi - lo, span,
          "" if rt.world == 1 else "  [rank 0 of %d]" % rt.world))
    for e in evs[lo:hi]:
        print("%8.1f +%6.1f us  %s" % (e.time_range.start - t0, e.time_range.end - e.time_range.start, e.name[:70]))
if rt.world > 1:
    # per-rank summary: where the step's time goes on every GPU (adam_pack = fused peer all-reduce + Adam: its duration
    # includes the wait for the slowest rank's gradients)
    dur = {}
    for e in evs[lo:hi]:
        k = e.name.split("(")[0].replace("void ", "").replace("glb::", "")[:28]
        dur[k] = dur.get(k, 0.0) + (e.time_range.end - e.time_range.start)
    rows = rt.all_gather_object((rt.rank, span, dur))
    if rt.rank == 0:
        keys = sorted({k for _, _, d in rows for k in d})
        print("per-rank kernel time (us) in one replay:")
        print("rank  span   " + "  ".join("%-28s" % k for k in keys))
        for r, sp, d in rows:
            print("%4d %6.1f  " % (r, sp) + "  ".join("%-28.1f" % d.get(k, 0.0) for k in keys))
rt.barrier()
