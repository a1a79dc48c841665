"""SPMD worker: sharded embedding table (K9) on 2 ranks - lookups of remote rows and asynchronous-PS style sparse updates
applied on the owning rank (rows touched by both ranks move twice)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from graphlearn_b200.parallel.runtime import init
from graphlearn_b200.nn.embedding import ShardedEmbedding
rt = init(device="cpu")
emb = ShardedEmbedding(rt, 50, 8, lr=0.5)
ids = torch.tensor([1, 2, 3, 4, 10, 11, 49]) if rt.rank == 0 else torch.tensor([1, 7, 8, 49, 0])
before = emb(ids).detach().clone()
out = emb(ids)
out.sum().backward()          # dL/drow = 1 for every looked-up row
rt.barrier()
after = emb(ids).detach()
# row 1 and 49 were looked up by BOTH ranks -> moved by 2 * lr, the others by lr
delta = (before - after)[:, 0]
exp = torch.tensor([1.0, .5, .5, .5, .5, .5, 1.0]) if rt.rank == 0 else torch.tensor([1.0, .5, .5, 1.0, .5])
assert torch.allclose(delta, exp, atol=1e-5), (delta, exp)
rt.barrier()
if rt.rank == 0:
    print("EMB_ALL_OK", flush=True)
rt.shutdown()
