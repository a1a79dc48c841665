"""CPU microbenchmark of the streaming-record parsers (dgs/file_loader.py): the native columnar parser
(csrc/host_loader.cpp parse_records) vs the line-by-line Python loader on a generated u2i record file (10 % vertex records with
16 list-valued floats, 90 % weighted edges).  Parser only - the sink discards the batches.

    python tools/bench_record_parser.py [--records 400000]
"""
import argparse
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--records", type=int, default=400000)
    a = ap.parse_args()
    from graphlearn_b200.dgs import FileLoader, Schema
    from graphlearn_b200.parallel.runtime import native
    native()                                                       # load the extension outside the timed region
    schema = Schema({
        "attr_defs": [{"type": 0, "name": "timestamp", "value_type": "INT64"}, {"type": 1, "name": "weight", "value_type": "FLOAT32"},
                      {"type": 2, "name": "feature", "value_type": "FLOAT32_LIST"}],
        "vertex_defs": [{"vtype": 0, "name": "user", "attr_types": [0, 2]}, {"vtype": 1, "name": "item", "attr_types": [0, 2]}],
        "edge_defs": [{"etype": 2, "name": "u2i", "attr_types": [0, 1]}],
        "edge_relation_defs": [{"etype": 2, "src_vtype": 0, "dst_vtype": 1}]})
    d = tempfile.mkdtemp(prefix="glb_parse_")
    with open(d + "/pattern", "w") as f:
        f.write("#VERTEX:user,vid,timestamp,feature\n#VERTEX:item,vid,timestamp,feature\n#EDGE:u2i,src,dst,timestamp,weight\n")
    rs = np.random.RandomState(0)
    with open(d + "/data", "w") as f:
        for i in range(a.records):
            if i % 10 == 0:
                f.write("item,%d,%d,%s\n" % (rs.randint(0, 100000), i, ":".join("%.4f" % x for x in rs.rand(16))))
            else:
                f.write("u2i,%d,%d,%d,%.3f\n" % (rs.randint(0, 100000), rs.randint(0, 100000), i, rs.rand()))
    mb = os.path.getsize(d + "/data") / 1e6

    class Null(object):
        def apply_updates(self, b):
            pass
    for nat in (True, False):
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            n = FileLoader(d + "/pattern", schema, batch_size=65536, native=nat).load(d + "/data", Null())
            best = min(best, time.perf_counter() - t0)
        print("%-6s parser: %d records (%.1f MB) in %.3f s = %.1f MB/s, %.2f M records/s" %
              ("native" if nat else "python", n, mb, best, mb / best, n / best / 1e6))


if __name__ == "__main__":
    main()
