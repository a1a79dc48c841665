"""SPMD worker (2 ranks, gloo): nn/utils.py - bootstrap / get_cluster_spec exchange one address per rank, SyncBarrierHook holds
fast ranks until the slow one is done (twice: keys must not be reused), launch_server starts this rank's GraphServer."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch.distributed as dist
import graphlearn_b200 as gl
from graphlearn_b200 import nn as glnn
from graphlearn_b200.parallel.runtime import init
rt = init(device="cpu")
assert glnn.get_world_size() == 2 and glnn.get_rank() == rt.rank
spec = glnn.get_cluster_spec()
addrs = spec["server"].split(",")
assert len(addrs) == 2 and len(set(addrs)) == 2 and spec["client_count"] == 2, spec
for round_ in range(2):
    hook = glnn.SyncBarrierHook()
    t0 = time.time()
    if rt.rank == 1:
        time.sleep(1.0)
    hook.end()
    waited = time.time() - t0
    assert waited >= 0.9, (round_, rt.rank, waited)          # rank 0 was held back by rank 1
    dist.barrier()
n = 20
g = gl.Graph()
ids = np.arange(n)[np.arange(n) % 2 == rt.rank]
g.node({"ids": ids, "float_attrs": np.stack([ids, ids], 1).astype(np.float32)}, "n", decoder=gl.Decoder(attr_types=["float"] * 2))
g.edge({"src_ids": ids, "dst_ids": (ids + 1) % n}, ("n", "n", "e"))
assert not glnn.is_server_launched()
glnn.launch_server(g)                                         # cluster spec + task index from the helpers above
assert glnn.is_server_launched() and isinstance(glnn.get_counts(), dict)
try:
    glnn.launch_server(g)
    raise SystemExit("second launch must fail")
except RuntimeError:
    pass
dist.barrier()
g.close()
if rt.rank == 0:
    print("NN_UTILS_OK", flush=True)
rt.shutdown()
