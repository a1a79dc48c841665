"""One eager (no CUDA graph) training step of the flagship config between
cudaProfilerStart/Stop, for `ncu --profile-from-start off`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from graphlearn_b200.engine.fast_sage import FastSageTrainer
from graphlearn_b200.models.graphsage import EgoGraphSAGE
from graphlearn_b200.parallel.runtime import init
from graphlearn_b200.store.synthetic import make_sharded_graph

small = os.environ.get("GLB_SMALL", "0") == "1"
rt = init()
shape = dict(num_nodes=2_449_029, num_edges=123_718_280) if not small else dict(num_nodes=200_000, num_edges=5_000_000)
fdt = torch.float32 if os.environ.get('GLB_FDT', 'bf16') == 'fp32' else torch.bfloat16
nodes, csr = make_sharded_graph(rt, feat_dim=100, num_classes=47, seed=0, feature_dtype=fdt, **shape)
model = EgoGraphSAGE(100, 256, 47, 2).to(rt.device)
tr = FastSageTrainer(rt, nodes, csr, model, [25, 10], 1024, use_cuda_graph=False)
g = torch.Generator().manual_seed(0)
for _ in range(3):
    tr.step(torch.randint(0, nodes.n_local, (1024,), generator=g))
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
tr.step(torch.randint(0, nodes.n_local, (1024,), generator=g))
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled one step, loss", float(tr.h_loss))
