"""One of N plain worker processes (no torchrun): g.init(task_index, task_count, tracker=dir) - the reference's FS-tracker
worker-mode launch.  argv: data_dir tracker_dir task_index task_count"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import graphlearn_b200 as gl
from tests import fixtures as fx

d, tracker, idx, cnt = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
    assert k not in os.environ
g = gl.Graph()
g.node(os.path.join(d, "user.tsv"), "user", decoder=gl.Decoder(weighted=True, labeled=True, attr_types=["int", "int", "string", "float"]))
g.node(os.path.join(d, "item.tsv"), "item", decoder=gl.Decoder(attr_types=["float"] * 4))
g.edge(os.path.join(d, "u2i.tsv"), ("user", "item", "buy"), decoder=gl.Decoder(weighted=True))
g.init(task_index=idx, task_count=cnt, tracker=tracker, device="cpu")
rt = g.runtime
assert rt.world == cnt and rt.rank == idx
st = g.get_stats()
assert sum(st["user"]) == fx.N_USER and len(st["user"]) == cnt
ids = np.arange(fx.N_USER)
n = g.lookup_nodes("user", ids)                                     # rows owned by the OTHER process too
assert (n.labels == ids % 3).all()
mine = []
ds = gl.Dataset(g.V("user").batch(7).alias("u").values())
while True:
    try:
        mine += ds.next()["u"].ids.tolist()
    except gl.OutOfRangeError:
        break
assert sorted(mine) == [u for u in range(fx.N_USER) if u % cnt == idx]   # every worker traverses its own shard
rt.barrier()
g.close()
print("TRACKER_WORKER_OK %d" % idx, flush=True)
