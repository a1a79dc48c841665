"""Graph load time, ours vs the unmodified reference, on the CPU of this box (no GPU needed).
Both load the same TSV files (ogbn-products-shaped synthetic: 100 float attributes per node) and build their
in-memory graph: reference = local mode `g.init()` (loader threads -> UpdateNodes/UpdateEdges -> storages),
ours = native byte-range parser -> columnar tensors -> CSR / feature tables (device = cpu here).
usage: python tools/bench_load.py [data_root]   (default /tmp/glb_ref_small, generated on demand)"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
root = sys.argv[1] if len(sys.argv) > 1 else "/tmp/glb_ref_small"
N, E, D = 200_000, 5_000_000, 100
if not os.path.exists(os.path.join(root, "DONE")):
    subprocess.run([sys.executable, os.path.join(ROOT, "baseline", "ref_sampler.py"), "--root", root, "--nodes", str(N),
                    "--edges", str(E), "--gen-only"], check=True)
node_f, edge_f = os.path.join(root, "node.tsv"), os.path.join(root, "edge.tsv")
mb = (os.path.getsize(node_f) + os.path.getsize(edge_f)) / 1e6

ref_code = r'''
import sys, time
sys.path.insert(0, %r)
import graphlearn as gl
t0 = time.time()
g = gl.Graph().node(%r, node_type="item", decoder=gl.Decoder(labeled=True, attr_types=["float"] * %d, attr_delimiter=":")) \
    .edge(%r, edge_type=("item", "item", "e"), decoder=gl.Decoder(), directed=True).init()
print("REF_INIT %%.2f" %% (time.time() - t0), flush=True)
g.close()
''' % (os.path.join(ROOT, "baseline", "_ref"), node_f, D, edge_f)
ref_s = None
if os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "graphlearn")):
    p = subprocess.run([sys.executable, "-c", ref_code], capture_output=True, text=True, timeout=1800, cwd="/tmp")
    for line in p.stdout.splitlines():
        if line.startswith("REF_INIT"):
            ref_s = float(line.split()[1])
    if ref_s is None:
        print("reference failed:", (p.stdout + p.stderr)[-500:])

import torch  # noqa: E402
import graphlearn_b200 as gl  # noqa: E402
from graphlearn_b200.parallel.runtime import native  # noqa: E402
C = native()
t0 = time.time()
C.load_table(edge_f, True, False, False, False, [], [], ":", "\t", 8, 0, 1)
C.load_table(node_f, False, False, True, False, [1] * D, [], ":", "\t", 8, 0, 1)
parse_s = time.time() - t0
t0 = time.time()
g = gl.Graph().node(node_f, "item", decoder=gl.Decoder(labeled=True, attr_types=["float"] * D)) \
    .edge(edge_f, ("item", "item", "e"), decoder=gl.Decoder(), directed=True).init(device="cpu")
ours_s = time.time() - t0
st = g.get_stats()
print("files: %.0f MB (%d nodes x %d floats, %d edges), %d CPU threads" % (mb, sum(st["item"]), D, sum(st["e"]), os.cpu_count()))
print("ours      parse only %.2f s (%.0f MB/s)   full init %.2f s" % (parse_s, mb / parse_s, ours_s))
if ref_s is not None:
    print("reference full init  %.2f s   -> %.1fx" % (ref_s, ref_s / ours_s))
