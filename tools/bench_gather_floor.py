"""Random-row read ceiling of HBM3e on this GPU: how fast can ANY kernel fetch n random rows of a given size from a table
much larger than L2?  (The fused GraphSAGE layer reads ~308K random 200-byte rows per step; its roofline is this
number, not the streaming copy bandwidth.)

    python tools/bench_gather_floor.py            # prints one line per (row bytes, index order)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from graphlearn_b200.parallel.runtime import init, native  # noqa: E402

rt = init()
C = native()
dev = rt.device
g = torch.Generator(device=dev).manual_seed(0)
flush = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)
FLUSH_BY = os.environ.get("GLB_FLUSH", "read")     # "read": L2 ends up full of CLEAN lines; "write": full of dirty lines


def timed(fn, reps=7):
    ts = []
    for _ in range(reps):
        if FLUSH_BY == "write":
            flush.zero_()                          # evict L2 (126 MB) - leaves it full of dirty lines that the measured
        else:                                      # kernel has to write back while it reads
            flush.view(torch.int64).max()          # evict by reading: clean lines
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


# streaming reference: contiguous copy of the same number of bytes
src = torch.empty(64 << 20, dtype=torch.uint8, device=dev).random_()
dst = torch.empty_like(src)
t = timed(lambda: dst.copy_(src))
print("streaming copy of 64 MiB: %.1f us -> %.0f GB/s read (+ the same written)" % (t, (64 << 20) / t / 1e3))
t = timed(lambda: src.view(torch.int32).sum())
print("streaming read (int32 sum) of 64 MiB: %.1f us -> %.0f GB/s" % (t, (64 << 20) / t / 1e3))

table_bytes = 512 << 20
print("table %d MiB; L2 flushed (by %s) before every launch; median of 7" % (table_bytes >> 20, FLUSH_BY))
for total, row_bytes, stride_bytes in ((61.6e6, 64, 64), (61.6e6, 128, 128), (61.6e6, 208, 208), (246.4e6, 208, 208), (61.6e6, 208, 256), (61.6e6, 512, 512),
                                       (61.6e6, 4096, 4096), (246.4e6, 4096, 4096)):
    n_table = table_bytes // stride_bytes
    table = torch.empty(n_table, stride_bytes // 2, dtype=torch.bfloat16, device=dev)
    table.view(torch.int16).random_(-100, 100)
    n = int(total // row_bytes)
    out = torch.empty(n, row_bytes // 2, dtype=torch.bfloat16, device=dev)
    for order in ("random", "sorted"):
        idx = torch.randint(0, n_table, (n,), device=dev, generator=g)
        if order == "sorted":
            idx = idx.sort().values
        t = min(timed(lambda: C.gather_copy16(table, idx, row_bytes, out, False, u)) for u in (1, 2, 4, 8))
        trs = [timed(lambda: C.gather_copy16(table, idx, row_bytes, out, True, u)) for u in (1, 2, 4, 8)]
        tr = min(trs)
        C.gather_copy16(table, idx, row_bytes, out, False, 4)
        assert torch.equal(out[:1000].view(torch.int16), table[idx[:1000], :row_bytes // 2].contiguous().view(torch.int16))
        t2 = timed(lambda: torch.index_select(table, 0, idx, out=out), 3) if row_bytes == stride_bytes and row_bytes >= 208 else float("nan")
        print("%5.1f MB in rows of %4d B (stride %4d) %-6s n=%7d: read-only %6.1f us = %5.0f GB/s (loads in flight per thread 1/2/4/8: %s us) | read+write copy %6.1f us = %5.0f GB/s | torch.index_select %6.1f us = %5.0f GB/s"
              % (total / 1e6, row_bytes, stride_bytes, order, n, tr, n * row_bytes / tr / 1e3, "/".join("%.1f" % x for x in trs), t, n * row_bytes / t / 1e3, t2,
                 n * row_bytes / t2 / 1e3))
    del table, out
