"""Failure detection.  The reference retries gRPC calls with exponential back-off and
re-resolves broken channels (graphlearn/src/service/dist/grpc_client.cc:46-64,
channel_manager.cc:151-170) but has no replica / state transfer: a lost server reloads its shard.
On a single box the failure modes are different: a rank dies or diverges -> the device-side
barriers of the peer kernels stop being signalled.  Those barriers are bounded spins that set an
error flag (csrc/comm.cu); ``check`` turns the flag into an exception so that the job fails fast
and resumes from the last checkpoint (utils/checkpoint.py)."""
from __future__ import annotations

import time

import torch
import torch.distributed as dist


class Watchdog(object):
    def __init__(self, runtime, trainers=(), heartbeat_s: float = 30.0):
        self.rt = runtime
        self.trainers = list(trainers)
        self.heartbeat_s = heartbeat_s
        self._last = time.time()

    def check(self):
        """Raise if any peer-memory collective timed out on this rank."""
        for t in self.trainers:
            ar = getattr(t, "ar", None)
            if ar is not None:
                ar.check()

    def heartbeat(self) -> bool:
        """All ranks exchange a liveness token every `heartbeat_s` (cheap all-reduce)."""
        if self.rt.world == 1 or time.time() - self._last < self.heartbeat_s:
            return True
        self._last = time.time()
        tok = torch.ones(1, device=self.rt.device)
        work = dist.all_reduce(tok, async_op=True)
        work.wait()
        return int(tok.item()) == self.rt.world


class StallDetector(object):
    """Host-side hang detection: the training loop calls ``tick()`` every step; a daemon thread fires
    ``on_stall(seconds_since_last_tick)`` once when no tick arrived for ``timeout_s`` (a peer died inside a
    collective, a kernel spins on a flag that is never set, ...).  The default action dumps every thread's stack
    and exits with code 42 so that a supervisor (torchrun ``--max-restarts``) restarts the job, which then resumes
    from the last checkpoint (utils/checkpoint.py) - the reference's recovery story is the same: a lost server
    process is restarted and reloads its shard."""

    def __init__(self, timeout_s: float = 300.0, on_stall=None, poll_s: float = None):
        import threading
        self.timeout_s = float(timeout_s)
        self.on_stall = on_stall or self._default_action
        self._poll = poll_s if poll_s is not None else max(0.01, min(1.0, self.timeout_s / 4))
        self._last = time.time()
        self._stop = threading.Event()
        self._fired = False
        self._thread = threading.Thread(target=self._run, daemon=True)

    @staticmethod
    def _default_action(idle_s):
        import faulthandler
        import os
        import sys
        sys.stderr.write("[graphlearn_b200] no training step for %.0f s - dumping stacks and exiting (42)\n" % idle_s)
        faulthandler.dump_traceback(file=sys.stderr, all_threads=True)
        os._exit(42)

    def _run(self):
        while not self._stop.wait(self._poll):
            idle = time.time() - self._last
            if idle > self.timeout_s and not self._fired:
                self._fired = True
                self.on_stall(idle)

    def start(self):
        self._last = time.time()
        self._thread.start()
        return self

    def tick(self):
        self._last = time.time()
        self._fired = False

    def stop(self):
        self._stop.set()
