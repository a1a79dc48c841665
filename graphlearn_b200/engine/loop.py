"""Generic training / evaluation / embedding-export loop for arbitrary (dataset, model, loss) -
the PyTorch counterpart of LocalTrainer / DistTrainer
(graphlearn/examples/tf/trainer.py:85-279): epochs end on ``OutOfRangeError``, progress is logged
as steps/sec, checkpoints are periodic, embeddings are dumped in the ``id:int64\\temb:string``
dialect.  Multi-GPU: gradients are averaged with the one-shot peer all-reduce over the flat
gradient buffer (or NCCL), every rank iterates its own shard like the reference's workers."""
from __future__ import annotations

from typing import Callable, Optional

import torch

from .. import errors
from ..ops import comm as comm_ops
from ..parallel.runtime import Runtime
from ..utils.checkpoint import save_checkpoint, save_embeddings
from ..utils.trace import ProgressLogger


class Trainer(object):
    def __init__(self, rt: Runtime, model: torch.nn.Module, dataset, step_fn: Callable, lr: float = 1e-3,
                 optimizer: Optional[torch.optim.Optimizer] = None, allreduce: str = "peer",
                 ckpt_path: str = "", ckpt_every: int = 0, log_every: int = 100):
        """step_fn(model, batch) -> loss tensor.  `dataset.next()` yields batches (gl.Dataset / nn.Dataset)."""
        self.rt, self.model, self.dataset, self.step_fn = rt, model, dataset, step_fn
        self.flat_p, self.flat_g = comm_ops.flatten_module(model)
        if rt.world > 1:            # replicas start from rank 0's initialisation (what DDP does at construction)
            import torch.distributed as dist
            dist.broadcast(self.flat_p, src=0)
        self.opt = optimizer or comm_ops.FlatAdam(self.flat_p, self.flat_g, lr=lr)
        self._flat_opt = optimizer is None
        self.ar = comm_ops.PeerAllReduce(rt, self.flat_g.numel(), backend=allreduce)
        self.ckpt_path, self.ckpt_every = ckpt_path, ckpt_every
        self.progress = ProgressLogger(every=log_every)
        self.global_step = 0

    def _batch(self):
        nxt = getattr(self.dataset, "next")
        return nxt()

    def train_epoch(self, max_steps: Optional[int] = None) -> float:
        """One pass over this rank's shard.  With world > 1 every rank must take the same number of
        steps (collectives inside): pass `max_steps` = min over ranks, as the reference's PyTorch
        example does (examples/pytorch/gcn/train.py:174)."""
        self.model.train()
        tot, n = 0.0, 0
        while max_steps is None or n < max_steps:
            try:
                batch = self._batch()
            except errors.OutOfRangeError:
                break
            self.flat_g.zero_()
            loss = self.step_fn(self.model, batch)
            loss.backward()
            self.ar(self.flat_g, average=True)
            self.opt.step()
            tot += float(loss.detach())
            n += 1
            self.global_step += 1
            self.progress.update(loss.detach())
            if self.ckpt_every and self.ckpt_path and self.global_step % self.ckpt_every == 0:
                self.save(self.ckpt_path)
        return tot / max(n, 1)

    @torch.no_grad()
    def evaluate(self, dataset, metric_fn: Callable) -> float:
        self.model.eval()
        vals, n = 0.0, 0
        while True:
            try:
                batch = dataset.next()
            except errors.OutOfRangeError:
                break
            vals += float(metric_fn(self.model, batch))
            n += 1
        self.model.train()
        return vals / max(n, 1)

    @torch.no_grad()
    def export_embeddings(self, dataset, embed_fn: Callable, path: str, block_max_lines: int = 0):
        """embed_fn(model, batch) -> (ids, emb); written as `<path>.rank<r>`."""
        self.model.eval()
        ids, embs = [], []
        while True:
            try:
                batch = dataset.next()
            except errors.OutOfRangeError:
                break
            i, e = embed_fn(self.model, batch)
            ids.append(i.reshape(-1).cpu())
            embs.append(e.float().cpu())
        self.model.train()
        if ids:
            save_embeddings("%s.rank%d" % (path, self.rt.rank), torch.cat(ids), torch.cat(embs), block_max_lines)

    def save(self, path: str):
        state = {"model": self.flat_p.clone(), "step": self.global_step}
        if self._flat_opt:
            state["opt"] = self.opt.state_dict()
        save_checkpoint(path, extra=state, datasets={"train": self.dataset} if hasattr(self.dataset, "state_dict") else None,
                        rank=self.rt.rank)

    def load(self, path: str) -> bool:
        """Restore what ``save`` wrote for this rank (parameters, optimiser moments, step counter, traversal cursor);
        False when there is no checkpoint."""
        import os
        from ..utils.checkpoint import load_checkpoint
        if not os.path.exists("%s.rank%d" % (path, self.rt.rank)):
            return False
        st = load_checkpoint(path, datasets={"train": self.dataset} if hasattr(self.dataset, "load_state_dict") else None,
                             rank=self.rt.rank, map_location=self.flat_p.device)
        with torch.no_grad():
            self.flat_p.copy_(st["model"].to(self.flat_p.device))
        if self._flat_opt and "opt" in st:
            self.opt.load_state_dict(st["opt"])
        self.global_step = int(st["step"])
        return True
