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
