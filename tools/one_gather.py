import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from graphlearn_b200.parallel.runtime import init, native
from graphlearn_b200.store.shards import IdMap, NodeTable
rt = init(); C = native(); dev = rt.device
M, k, d, n_out = 25600, 10, 100, 256
w = torch.randn(n_out, 256, device=dev) * 0.05
img, _ = C.pack_weight_f32(w, 256, False)
bias = torch.zeros(n_out, device=dev)
out = torch.empty(M, n_out, dtype=torch.bfloat16, device=dev)
asave = torch.empty(M, 256, dtype=torch.bfloat16, device=dev)
n_nodes = 2_449_029
dt = torch.bfloat16 if os.environ.get("GLB_FDT") == "bf16" else torch.float32
t = NodeTable(rt, "t", IdMap(rt, torch.arange(n_nodes, device=dev), dense=True))
t.set_float(torch.randn(n_nodes, d, device=dev), dt)
for it in range(4):
    sv = torch.randint(0, n_nodes, (M,), device=dev); nv = torch.randint(0, n_nodes, (M * k,), device=dev)
    C.sage_fused_forward(t.feat_desc, sv, t.feat_desc, nv, M, k, 0, img, bias, 256, n_out, True, True, True, 128, out, asave, None, int(os.environ.get('GLB_MODE','0')))
torch.cuda.synchronize()
print("ok")
