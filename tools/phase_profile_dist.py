"""Average per-kernel device time of eager training steps on N ranks (torchrun), rank 0 prints.
Kernel durations are exact; gaps between kernels are eager-launch artefacts and ignored."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from graphlearn_b200.engine.fast_sage import FastSageTrainer
from graphlearn_b200.models.graphsage import EgoGraphSAGE
from graphlearn_b200.parallel.runtime import init
from graphlearn_b200.store.synthetic import make_sharded_graph

rt = init()
W = rt.world
shape = dict(num_nodes=2_449_029, num_edges=123_718_280)
nodes, csr = make_sharded_graph(rt, feat_dim=100, num_classes=47, seed=0, feature_dtype=torch.bfloat16, **shape)
if W > 1 and os.environ.get("GLB_CACHE", "1") == "1":
    nodes.build_feature_cache(10 ** 9)
model = EgoGraphSAGE(100, 256, 47, 2).to(rt.device)
tr = FastSageTrainer(rt, nodes, csr, model, [25, 10], 1024, use_cuda_graph=False)
g = torch.Generator().manual_seed(rt.rank)
seeds = lambda: torch.randint(0, nodes.n_local, (1024,), generator=g) * W + rt.rank
for _ in range(10):
    tr.step(seeds())
torch.cuda.synchronize(); rt.barrier()
from torch.profiler import profile, ProfilerActivity
NS = 40
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(NS):
        tr.step(seeds())
    torch.cuda.synchronize()
rt.barrier()
if rt.rank == 0:
    agg = collections.OrderedDict()
    evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
    for e in evs:
        a = agg.setdefault(e.name[:60], [0, 0.0])
        a[0] += 1; a[1] += e.device_time
    tot = 0.0
    for k, (n, us) in agg.items():
        print("%-62s x%.1f  %8.1f us/step" % (k, n / NS, us / NS)); tot += us / NS
    print("sum of kernel time per step: %.1f us (world %d)" % (tot, W))
rt.shutdown()
