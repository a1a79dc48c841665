"""Sampler process of the reference arm: runs the UNMODIFIED alibaba/graph-learn engine
(baseline/_ref) - gl.Graph + GSL + gl.Dataset - and hands every sampled batch to the trainer
process through POSIX shared memory.

Why a separate process: this image's PyTorch and the reference's bundled gRPC/protobuf/glog
stack crash when loaded into one interpreter (segfault in either import order, also with
-Bsymbolic / RTLD_DEEPBIND), so the reference - which is NumPy-only on its sampling side - gets
its own interpreter.  This is the same process split the reference's PyTorch example uses
(DataLoader workers are the GL clients, graphlearn/python/nn/pytorch/data/pyg_dataloader.py:42-117),
with shared memory instead of pickled queues so that the hand-off costs one memcpy.

protocol (line based, stdin/stdout):  child -> "SHM <name> <slot_bytes>", then per batch
"READY <slot>";  parent -> "FREE <slot>" or "STOP".
"""
from __future__ import annotations

import argparse
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))


def gen_data(root, n_nodes, n_edges, dim, classes, seed=0):
    """ogbn-products-shaped random graph in the reference's TSV dialect (cached on disk)."""
    import numpy as np
    import pyarrow as pa
    import pyarrow.csv as pacsv
    os.makedirs(root, exist_ok=True)
    done = os.path.join(root, "DONE")
    node_f, edge_f = os.path.join(root, "node.tsv"), os.path.join(root, "edge.tsv")
    if os.path.exists(done):
        return node_f, edge_f
    rs = np.random.RandomState(seed)
    lut = np.array(["%.2f" % (i / 100.0) for i in range(-400, 401)], dtype=object)
    with open(node_f, "w") as f:
        f.write("id:int64\tlabel:int32\tfeature:string\n" if dim > 0 else "id:int64\tlabel:int32\n")
        step = 100_000
        for s in range(0, n_nodes, step):
            e = min(n_nodes, s + step)
            lab = rs.randint(0, classes, e - s)
            if dim == 0:
                f.write("".join("%d\t%d\n" % (s + i, lab[i]) for i in range(e - s)))
                continue
            q = np.clip((rs.randn(e - s, dim) * 100).astype(np.int64), -400, 400) + 400
            strs = lut[q]
            f.write("".join("%d\t%d\t%s\n" % (s + i, lab[i], ":".join(strs[i])) for i in range(e - s)))
    w = np.exp(rs.randn(n_nodes))
    deg = np.floor(w / w.sum() * n_edges).astype(np.int64)
    rem = n_edges - int(deg.sum())
    deg += np.bincount(rs.randint(0, n_nodes, rem), minlength=n_nodes)
    src = np.repeat(np.arange(n_nodes, dtype=np.int64), deg)
    dst = rs.randint(0, n_nodes, n_edges).astype(np.int64)
    tbl = pa.table({"src_id:int64": src, "dst_id:int64": dst})
    with open(edge_f, "wb") as f:          # header by hand: pyarrow always quotes column names
        f.write(b"src_id:int64\tdst_id:int64\n")
        pacsv.write_csv(tbl, f, write_options=pacsv.WriteOptions(delimiter="\t", quoting_style="none",
                                                                 include_header=False))
    open(done, "w").write("ok")
    return node_f, edge_f


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--root", required=True)
    ap.add_argument("--nodes", type=int, required=True)
    ap.add_argument("--edges", type=int, required=True)
    ap.add_argument("--dim", type=int, default=100)
    ap.add_argument("--classes", type=int, default=47)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--fanouts", default="25,10")
    ap.add_argument("--rank", type=int, default=0)
    ap.add_argument("--world", type=int, default=1)
    ap.add_argument("--tracker", default="")
    ap.add_argument("--hosts", default="")
    ap.add_argument("--gen-only", action="store_true")
    ap.add_argument("--slots", type=int, default=2)
    ap.add_argument("--walk-len", type=int, default=0, help="> 0: DeepWalk mode (random_walk + negatives), timed in this process")
    ap.add_argument("--neg", type=int, default=5)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    if a.gen_only:
        t0 = time.time()
        gen_data(a.root, a.nodes, a.edges, a.dim, a.classes)
        print("GEN %.1f" % (time.time() - t0), flush=True)
        return
    sys.path.insert(0, os.path.join(HERE, "_ref"))
    import graphlearn as gl            # must be the FIRST native import (it crashes when NumPy is loaded before it)
    import numpy as np
    from multiprocessing import shared_memory
    fans = [int(x) for x in a.fanouts.split(",")]
    node_f, edge_f = os.path.join(a.root, "node.tsv"), os.path.join(a.root, "edge.tsv")
    t0 = time.time()
    g = gl.Graph() \
        .node(node_f, "i", decoder=gl.Decoder(labeled=True, attr_types=["float"] * a.dim) if a.dim > 0 else gl.Decoder(labeled=True)) \
        .edge(edge_f, ("i", "i", "e"), decoder=gl.Decoder())
    if a.world > 1 and a.hosts:
        # RPC tracker with an explicit host list: the FS tracker publishes gethostbyname(hostname), which is
        # empty on boxes whose hostname only resolves to loopback (graphlearn/src/common/base/host.cc:26-50)
        gl.set_tracker_mode(0)
    if a.world == 1:
        g.init()
    else:
        g.init(task_index=a.rank, task_count=a.world, tracker=a.tracker, hosts=a.hosts or None)
    load_s = time.time() - t0
    if a.walk_len > 0:
        # BASELINE config 5: the reference's own walk path (one sharded request per step, random_walk.cc:55-135)
        q = g.V("i").batch(a.batch).shuffle(traverse=True).alias("src").random_walk("e", a.walk_len).alias("walk") \
             .outNeg("e").sample(a.neg).by("random").alias("neg").values()
        ds = gl.Dataset(q, window=10)

        def one():
            while True:
                try:
                    r = ds.next()
                    break
                except gl.OutOfRangeError:
                    continue
            return r["walk"].ids.shape[0] + 0 * r["neg"].ids.shape[0]
        for _ in range(a.warmup):
            one()
        t1 = time.time()
        for _ in range(a.steps):
            one()
        dt = time.time() - t1
        print("WALK %.6f %.1f" % (dt, load_s), flush=True)
        g.close()
        return
    q = g.V("i").batch(a.batch).shuffle(traverse=True).alias("src")
    for i, f in enumerate(fans):
        q = q.outV("e").sample(f).by("random").alias("h%d" % (i + 1))
    q = q.values()
    ds = gl.Dataset(q, window=10)
    B = a.batch
    ns = [B]
    for f in fans:
        ns.append(ns[-1] * f)
    n_rows = sum(ns)
    fbytes = n_rows * a.dim * 4
    slot_bytes = fbytes + B * 8
    shm = shared_memory.SharedMemory(create=True, size=slot_bytes * a.slots)
    print("SHM %s %d %.1f" % (shm.name, slot_bytes, load_s), flush=True)
    free = list(range(a.slots))
    try:
        while True:
            while not free:
                line = sys.stdin.readline()
                if not line or line.startswith("STOP"):
                    raise SystemExit
                if line.startswith("FREE"):
                    free.append(int(line.split()[1]))
            s = free.pop(0)
            while True:
                try:
                    r = ds.next()
                    break
                except gl.OutOfRangeError:
                    continue
            off = s * slot_bytes
            buf = np.ndarray((n_rows, a.dim), dtype=np.float32, buffer=shm.buf, offset=off)
            x0 = r["src"].float_attrs.reshape(-1, a.dim)
            if x0.shape[0] != ns[0]:       # short last batch of an epoch: skip
                free.insert(0, s)
                continue
            buf[:ns[0]] = x0
            o = ns[0]
            for i in range(len(fans)):
                buf[o:o + ns[i + 1]] = r["h%d" % (i + 1)].float_attrs.reshape(-1, a.dim)
                o += ns[i + 1]
            y = np.ndarray((B,), dtype=np.int64, buffer=shm.buf, offset=off + fbytes)
            y[:] = r["src"].labels.reshape(-1)
            print("READY %d" % s, flush=True)
    except SystemExit:
        pass
    finally:
        try:
            g.close()
        except Exception:
            pass
        shm.close()
        shm.unlink()


if __name__ == "__main__":
    main()
