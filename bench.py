#!/usr/bin/env python
"""Headline benchmark: sampled-subgraph train steps/sec for 2-layer GraphSAGE
(fan-out 25,10) on an ogbn-products-shaped synthetic graph (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W          # our engine
    python bench.py --impl reference --gpus N ...          # unmodified reference (baseline/_ref)

One JSON line is printed by rank 0.  `value` is the whole-job steps/sec summed
over all N GPUs (each rank trains its own 1024-seed batch per step: weak
scaling), device-timed with CUDA events, max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

PRODUCTS = dict(num_nodes=2_449_029, num_edges=123_718_280, feat_dim=100, num_classes=47)
FANOUTS = [25, 10]
HIDDEN = 256
BATCH = 1024


class ClockSampler:
    """Samples SM clocks / throttle reasons DURING the timed region (NVML in a thread, 20 ms period;
    falls back to polling nvidia-smi)."""

    def __init__(self, index=0, period=0.02):
        self.index, self.period = index, period
        self.sm, self.mx, self.reasons = [], [], set()
        self._stop = threading.Event()
        self._th = None
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:
            self._nvml = None

    def _run_nvml(self):
        n = self._nvml
        bits = {"hw_slowdown": getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM)))
                self.mx.append(float(n.nvmlDeviceGetMaxClockInfo(self._h, n.NVML_CLOCK_SM)))
                r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for name, b in bits.items():
                    if r & b:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(self.period)

    def _run_smi(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.sm.append(float(parts[0])); self.mx.append(float(parts[1]))
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                        if v.lower().startswith("active"):
                            self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self._th = threading.Thread(target=self._run_nvml if self._nvml else self._run_smi, daemon=True)
        self._th.start()

    def stop(self):
        self._stop.set()
        if self._th:
            self._th.join(timeout=6)
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm),
                "source": "nvml" if self._nvml else "nvidia-smi"}


def count_own_launches(trainer):
    """Kernel launches of one step, split into ours (namespace glb::) and library kernels."""
    import torch
    from torch.profiler import ProfilerActivity, profile
    trainer_graph, trainer.graph = trainer.graph, None         # eager step so every launch is visible
    try:
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            trainer._step_body()
            torch.cuda.synchronize()
        own, lib = 0, 0
        names = {}
        times = {}
        for ev in prof.events():
            if ev.device_type is not None and str(ev.device_type).endswith("CUDA") and ev.name and \
                    not ev.name.startswith("Memcpy") and not ev.name.startswith("Memset"):
                key = ev.name.split("(")[0][:60]
                t = getattr(ev, "device_time", None)
                if t is None:
                    t = getattr(ev, "cuda_time", 0.0)
                times[key] = round(times.get(key, 0.0) + float(t), 1)
                if "glb::" in ev.name:
                    own += 1
                    names[key] = names.get(key, 0) + 1
                else:
                    lib += 1
        trainer._kernel_us = times
        return own, lib, names
    finally:
        trainer.graph = trainer_graph


def run_ours(args):
    import torch
    import torch.distributed as dist

    from graphlearn_b200.engine.fast_sage import FastSageTrainer
    from graphlearn_b200.engine.trainer import SageTrainer
    from graphlearn_b200.models.graphsage import EgoGraphSAGE
    from graphlearn_b200.parallel.runtime import init
    from graphlearn_b200.store.synthetic import make_sharded_graph

    rt = init()
    assert rt.is_cuda, "bench.py needs a CUDA device"
    W = rt.world
    shape = dict(PRODUCTS)
    if args.small:
        shape = dict(num_nodes=200_000, num_edges=5_000_000, feat_dim=100, num_classes=47)
    fdt = torch.bfloat16 if args.feature_dtype == "bf16" else torch.float32
    t0 = time.time()
    nodes, csr = make_sharded_graph(rt, feature_dtype=fdt, seed=0, **shape)
    # N17 replica cache of remote feature rows in local HBM (the reference's
    # set_local_node_cache_capacity): -1 = as many remote rows as fit in 25% of the free HBM
    cache_rows = 0
    if W > 1 and args.feature_cache_rows != 0:
        cap = args.feature_cache_rows
        if cap < 0:
            free_b, _ = torch.cuda.mem_get_info()
            cap = int(0.25 * free_b) // (nodes.feats.local.size(1) * nodes.feats.local.element_size())
        scores = None
        if cap < shape["num_nodes"]:
            max_vid = max(int(n) for n in nodes.nrows) * W
            scores = torch.bincount(csr.indices.local.clamp(min=0), minlength=max_vid).float()[:max_vid]
            dist.all_reduce(scores)
        cache_rows = nodes.build_feature_cache(cap, scores=scores)
    torch.cuda.synchronize()
    build_s = time.time() - t0
    torch.manual_seed(0)
    model = EgoGraphSAGE(shape["feat_dim"], HIDDEN, shape["num_classes"], 2).to(rt.device)
    Trainer = SageTrainer if args.engine == "autograd" else FastSageTrainer
    tr = Trainer(rt, nodes, csr, model, FANOUTS, args.batch, lr=3e-3, allreduce=args.allreduce,
                 use_cuda_graph=not args.no_graph)
    # host-side seed stream: each rank traverses (shuffled) its own nodes, like the reference's
    # V().batch().shuffle(traverse=True) root which is unsharded (node_getter.cc:64-92)
    gen = torch.Generator().manual_seed(1234 + rt.rank)
    n_local = nodes.n_local
    total = args.warmup + 2 * args.steps + 8
    seed_rows = torch.randint(0, n_local, (total, args.batch), generator=gen)
    seed_ids = (seed_rows * W + rt.rank).pin_memory()
    tr.seeds.copy_(seed_ids[0])
    tr.capture()
    it = 0
    for _ in range(max(args.warmup, 3)):
        tr.step(seed_ids[it]); it += 1
    torch.cuda.synchronize()
    rt.barrier()

    # ---- device-timed region (kernel path): seeds already resident, K graph replays
    clocks = ClockSampler(rt.local_rank) if rt.rank == 0 else None
    if clocks:
        clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rt.barrier(); torch.cuda.synchronize()
    ev0.record()
    for _ in range(args.steps):
        tr.step_device()
    ev1.record()
    torch.cuda.synchronize(); rt.barrier()
    ms_dev = ev0.elapsed_time(ev1)

    # ---- end-to-end region through the public step(): every call copies that call's seed batch from pinned
    # host memory to the device and one loss back to pinned host memory.  Both copies are nodes of the step's
    # CUDA graph on side branches (the batch staged by call t is trained by call t+1: input prefetch), so the
    # PCIe round trips overlap the compute instead of serialising with it.
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rt.barrier(); torch.cuda.synchronize()
    ev2.record()
    last = None
    for _ in range(args.steps):
        last = tr.step(seed_ids[it]); it += 1
    ev3.record()
    torch.cuda.synchronize(); rt.barrier()
    ms_e2e = ev2.elapsed_time(ev3)
    clk = clocks.stop() if clocks else None
    final_loss = float(last)
    tr.ar.check()

    t = torch.tensor([ms_dev, ms_e2e], device=rt.device, dtype=torch.float64)
    if W > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = float(t[0]), float(t[1])
    own, lib, names = count_own_launches(tr)
    # remote feature traffic of the fused gather (hot path b/c): rows read per step and rank = self rows of both
    # layer-1 segments + their neighbour rows; a row owned by another rank and not replicated locally crosses NVLink
    row_bytes = int(nodes.feats.local.size(1)) * nodes.feats.local.element_size()
    rows_per_step = args.batch * (1 + FANOUTS[0]) + args.batch * FANOUTS[0] * (1 + FANOUTS[1])
    n_total = sum(int(x) for x in nodes.nrows)
    remote_frac = 0.0 if W == 1 else (W - 1) / W * max(0.0, 1.0 - cache_rows / max(n_total - nodes.n_local, 1))
    remote_bytes = rows_per_step * row_bytes * remote_frac
    if rt.rank == 0:
        steps_per_s = W * args.steps / (ms_dev / 1e3)
        e2e_steps_per_s = W * args.steps / (ms_e2e / 1e3)
        out = {
            "metric": "sampled-subgraph train steps/sec (2-layer GraphSAGE fanout 25,10, ogbn-products-shaped synthetic)",
            "value": steps_per_s, "unit": "steps/s", "n_gpus": W, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic (random graph of ogbn-products shape, random-init weights)",
            "impl": "graphlearn_b200",
            "config": {"model": "GraphSAGE-2layer-mean hidden256", "global_batch": args.batch * W, "seq_len": None,
                       "fanout": FANOUTS, "parallelism": "dp%d+graph-partition%d" % (W, W),
                       "num_nodes": shape["num_nodes"], "num_edges": shape["num_edges"],
                       "feat_dim": shape["feat_dim"], "feature_storage": args.feature_dtype,
                       "feature_cache_rows_per_gpu": cache_rows,
                       "feature_cache": ("n/a (single GPU)" if W == 1 else
                                         "none (every remote row is read from its owner's HBM over NVLink inside the fused kernel)"
                                         if cache_rows == 0 else
                                         "replica cache of remote feature rows in local HBM (reference: set_local_node_cache_capacity); "
                                         "topology stays partitioned; --feature-cache-rows 0 disables it"),
                       "remote_feature_bytes_per_step_per_gpu": int(remote_bytes),
                       "remote_feature_GBps_per_gpu": round(remote_bytes / (ms_dev / args.steps * 1e-3) / 1e9, 1),
                       "nvlink_random_row_ceiling_GBps": 477 if row_bytes <= 256 else 580,
                       "l2_policy": "inputs larger than L2: every step gathers ~%d random feature rows from a %.1f GB table"
                                    % (args.batch * (1 + 25 + 250), shape["num_nodes"] * shape["feat_dim"] * (2 if fdt == torch.bfloat16 else 4) / 1e9),
                       "allreduce": tr.ar.backend if W > 1 else "none", "cuda_graph": tr.graph is not None, "engine": args.engine,
                       "graph_build_s": round(build_s, 2)},
            "e2e": {"value": e2e_steps_per_s, "unit": "steps/s", "h2d_bytes_per_step": args.batch * 8,
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": own * args.steps, "own_kernels_per_step": own, "library_kernels_per_step": lib,
            "own_kernel_names": names, "kernel_us_eager_step": getattr(tr, "_kernel_us", {}), "clocks": clk, "final_loss": final_loss,
        }
        print(json.dumps(out))
    rt.barrier()
    rt.shutdown()


def run_reference(args):
    try:
        from baseline import run_reference as rr
    except Exception as e:  # pragma: no cover
        print(json.dumps({"impl": "reference", "unavailable": "baseline runner import failed: %r" % (e,)}))
        return
    rr.main(args)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--feature-dtype", default="bf16", choices=["fp32", "bf16"],
                    help="HBM storage dtype of the float attribute table (compute is bf16 either way; bf16 halves NVLink bytes)")
    ap.add_argument("--allreduce", default="peer", choices=["peer", "nccl"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--engine", default="fast", choices=["fast", "autograd"],
                    help="fast = hand-scheduled fwd/bwd kernel chain; autograd = torch.autograd over the same kernels")
    ap.add_argument("--feature-cache-rows", type=int, default=-1,
                    help="remote feature rows replicated per GPU (N17 cache): -1 auto (25%% of free HBM), 0 off")
    ap.add_argument("--small", action="store_true", help="small graph for quick functional runs")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
