"""Per-phase clock64 timeline of the fused SAGE kernel (thread 0 of every CTA)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from graphlearn_b200.parallel.runtime import init, native
from graphlearn_b200.store.shards import IdMap, NodeTable
rt = init(); C = native(); dev = rt.device
M, k, d, n_out = 25600, 10, 100, 256
MODE = int(os.environ.get("GLB_MODE", "0"))
w = torch.randn(n_out, 256, device=dev) * 0.05
img, _ = C.pack_weight_f32(w, 256, False)
bias = torch.zeros(n_out, device=dev)
out = torch.empty(M, n_out, dtype=torch.bfloat16, device=dev)
asave = torch.empty(M, 256, dtype=torch.bfloat16, device=dev)
names = ["start", "setup", "ids->smem", "gather", "sync", "W wait", "mma issue", "mma wait", "epilogue", "final sync"]
for dt in (torch.float32, torch.bfloat16):
    n_nodes = 2_449_029
    t = NodeTable(rt, "t", IdMap(rt, torch.arange(n_nodes, device=dev), dense=True))
    t.set_float(torch.randn(n_nodes, d, device=dev), dt)
    for R in (128, 64):
        grid = (M + R - 1) // R
        ts = torch.zeros(grid * 16, dtype=torch.int64, device=dev)
        for it in range(3):
            sv = torch.randint(0, n_nodes, (M,), device=dev); nv = torch.randint(0, n_nodes, (M * k,), device=dev)
            C.sage_fused_forward(t.feat_desc, sv, t.feat_desc, nv, M, k, 0, img, bias, 256, n_out, True, True, True, R, out, asave, ts, MODE)
        torch.cuda.synchronize()
        T = ts.view(grid, 16).cpu().double()
        d_ = (T[:, 1:10] - T[:, 0:9])
        tot = (T[:, 9] - T[:, 0])
        print("dtype=%s R=%d grid=%d  mean cycles per phase (first-wave CTAs):" % (dt, R, grid))
        fw = slice(0, min(grid, 148))
        for i in range(9):
            print("   %-12s %9.0f" % (names[i + 1], d_[fw, i].mean().item()))
        print("   total        %9.0f   (kernel span %.0f cycles)" % (tot[fw].mean().item(), (T[:, 9].max() - T[:, 0].min()).item()))
