"""2-rank microbenchmark: rank 0 gathers random rows that live on rank 1 (peer loads over NVLink)
vs rows that live locally, for several row sizes / dtypes, with the plain gather_rows kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from graphlearn_b200.parallel.runtime import init, native, make_table_desc
rt = init(); C = native(); dev = rt.device
W, r = rt.world, rt.rank
n_rows = 1_000_000
for dim, dt in ((32, torch.float32), (64, torch.float32), (100, torch.float32), (128, torch.float32), (256, torch.float32),
                (100, torch.bfloat16), (128, torch.bfloat16), (256, torch.bfloat16)):
    esz = 4 if dt == torch.float32 else 2
    stride = (dim + (16 // esz) - 1) // (16 // esz) * (16 // esz)
    st = rt.symm_empty((n_rows, stride), dt)
    st.local.normal_()
    torch.cuda.synchronize(); rt.barrier()
    desc = make_table_desc(W, dim, stride, dt, st.nrows, st.ptrs)
    n = 2_000_000
    for where in ("local", "remote"):
        owner = r if where == "local" else (r + 1) % W
        vids = [torch.randint(0, n_rows, (n,), device=dev) * W + owner for _ in range(6)]
        for v in vids[:2]:
            C.gather_rows(desc, v, dt == torch.bfloat16, 0.0)
        torch.cuda.synchronize(); rt.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for v in vids[2:]:
            C.gather_rows(desc, v, dt == torch.bfloat16, 0.0)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 4
        if r == 0:
            print("row %4d B (%s d=%d) %-6s: %7.3f ms  %7.1f GB/s" % (dim * esz, "f32" if esz == 4 else "bf16", dim, where, ms, n * dim * esz / ms / 1e6), flush=True)
        rt.barrier()
rt.shutdown()
