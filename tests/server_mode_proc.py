"""One process of a plain-process server-mode job (no torchrun): argv = data_dir tracker_dir job_name task_index out_file.
cluster = {server_count: 2, client_count: 2, tracker}: two INDEPENDENT servers (replicated graph, hash-sharded traversal), two clients."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import graphlearn_b200 as gl

d, tracker, job, idx, out = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), sys.argv[5]
g = gl.Graph()
g.node(os.path.join(d, "user.tsv"), "user", decoder=gl.Decoder(weighted=True, labeled=True, attr_types=["int", "int", "string", "float"]))
g.node(os.path.join(d, "item.tsv"), "item", decoder=gl.Decoder(attr_types=["float"] * 4))
g.edge(os.path.join(d, "u2i.tsv"), ("user", "item", "buy"), decoder=gl.Decoder(weighted=True))
g.init(cluster={"server_count": 2, "client_count": 2, "tracker": tracker}, job_name=job, task_index=idx, device="cpu")
if job == "server":
    g.wait_for_close()
else:
    users, edges, full = [], [], 0
    ds = gl.Dataset(g.V("user").batch(6).alias("u").outV("buy").sample(2).by("topk").alias("i").values())
    while True:
        try:
            r = ds.next()
            users += r["u"].ids.tolist()
            assert r["i"].float_attrs.shape == (len(r["u"].ids), 2, 4)
        except gl.OutOfRangeError:
            break
    de = gl.Dataset(g.E("buy").batch(16).alias("e").values())
    while True:
        try:
            e = de.next()["e"]
            edges += list(zip(e.src_ids.tolist(), e.dst_ids.tolist()))
        except gl.OutOfRangeError:
            break
    fs = gl.Dataset(g.V("user").batch(5).alias("u").outV("buy").sample(0).by("full").alias("n").values()).next()["n"]
    full = [int(x) for x in fs.offsets]
    json.dump({"users": users, "edges": edges, "full": full, "stats": g.get_stats()}, open(out, "w"))
    g.close()
print("SERVER_MODE_PROC_OK", job, idx, flush=True)
