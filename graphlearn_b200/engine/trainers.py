"""``LocalTrainer`` / ``DistTrainer`` with the method names of the reference's example trainers
(graphlearn/examples/tf/trainer.py:44-408: ``train`` / ``test`` / ``train_and_evaluate`` / ``save_node_embedding`` /
``save_node_embedding_bigdata`` / ``join``; constructor knobs ``ckpt_dir``, ``save_checkpoint_secs``,
``save_checkpoint_steps``, ``profiling``, ``progress_steps``).

The reference builds a TF1 ``MonitoredTrainingSession`` around an iterator and a loss tensor; here the same driver loop runs
``engine.loop.Trainer`` (flat parameters, peer / NCCL gradient all-reduce, fused Adam) over a GSL dataset and a
``step_fn(model, batch) -> loss``:

    trainer = LocalTrainer(ckpt_dir="ckpt", progress_steps=50)
    trainer.train(dataset, model, step_fn, learning_rate=1e-2, epochs=10)
    acc = trainer.test(test_dataset, model, metric_fn)
    trainer.save_node_embedding("emb.txt", save_dataset, model, embed_fn)

``DistTrainer`` is the same loop under ``torchrun``: every rank iterates its own shard (like the reference's workers), epochs
are cut to the shortest rank so that the collectives line up, rank 0 is the chief, and ``join()`` holds finished workers until
all are done (the reference's ``SyncBarrierHook``).  Parameter servers do not exist: ``ps_count`` is accepted and ignored.
"""
from __future__ import annotations

import os
import time
from typing import Callable, Optional

import torch

from .. import errors
from ..utils.trace import StepProfiler, log
from .loop import Trainer


class _ExampleTrainer(object):
    def __init__(self, ckpt_dir: Optional[str] = None, save_checkpoint_secs: Optional[float] = 600,
                 save_checkpoint_steps: Optional[int] = None, profiling: bool = False, progress_steps: int = 10,
                 allreduce: str = "peer"):
        self.ckpt_dir, self.save_checkpoint_secs, self.save_checkpoint_steps = ckpt_dir, save_checkpoint_secs, save_checkpoint_steps
        self.profiling, self.progress_steps, self.allreduce = profiling, progress_steps, allreduce
        self.is_local = True
        self.global_step = 0
        self._trainer: Optional[Trainer] = None

    # ---- helpers
    def _runtime(self):
        from ..parallel.runtime import Runtime
        rt = Runtime.get()
        return rt if rt.initialized else rt.init()         # normally the graph's init() has set the runtime up already

    def _ckpt_path(self):
        return os.path.join(self.ckpt_dir, "model.ckpt") if self.ckpt_dir else ""

    def _epoch_steps(self, dataset, rt) -> Optional[int]:
        """steps every rank can take this epoch (None = run until OutOfRange): the minimum over ranks of the local batch count"""
        if rt.world == 1:
            return None
        n = getattr(dataset, "batches_per_epoch", None)
        n = n() if callable(n) else n
        if n is None:
            return None
        return min(rt.all_gather_object(int(n)))

    def _build(self, dataset, model, step_fn, optimizer, learning_rate):
        tr = self._trainer
        if tr is not None and tr.model is model and tr.dataset is dataset and optimizer is None and tr._flat_opt:
            tr.step_fn = step_fn                      # another train() call on the same job: keep parameters AND Adam moments
            if learning_rate is not None:
                tr.opt.lr = learning_rate
            return tr
        rt = self._runtime()
        if self.ckpt_dir:
            os.makedirs(self.ckpt_dir, exist_ok=True)
        tr = Trainer(rt, model, dataset, step_fn, lr=learning_rate if learning_rate is not None else 1e-3, optimizer=optimizer,
                     allreduce=self.allreduce, ckpt_path=self._ckpt_path(), ckpt_every=self.save_checkpoint_steps or 0,
                     log_every=max(1, self.progress_steps))
        if self.ckpt_dir and tr.load(self._ckpt_path()):
            log.info("restored %s at step %d", self._ckpt_path(), tr.global_step)
        self._trainer = tr
        return tr

    # ---- the reference's surface
    def train(self, dataset, model: torch.nn.Module, step_fn: Callable, optimizer=None, learning_rate: Optional[float] = None,
              epochs: int = 10, **_ignored) -> float:
        """-> mean loss of the last epoch.  Checkpoints every ``save_checkpoint_steps`` steps and/or ``save_checkpoint_secs``
        seconds (and at the end) when ``ckpt_dir`` is set; resumes from it when one exists."""
        tr = self._build(dataset, model, step_fn, optimizer, learning_rate)
        prof = StepProfiler(os.path.join(self.ckpt_dir or ".", "timeline"), start=500, stop=1000, every=100) if self.profiling else None
        last_save, loss = time.time(), float("nan")
        log.info("Start training...")
        for epoch in range(epochs):
            steps = self._epoch_steps(dataset, tr.rt)
            if prof is None and not self.save_checkpoint_secs:
                loss = tr.train_epoch(max_steps=steps)
            else:                                       # step-wise so that the timers / profiler see every step
                tot, n = 0.0, 0
                while steps is None or n < steps:
                    try:
                        if prof is not None:
                            with prof.step(tr.global_step):
                                l = self._one_step(tr)
                        else:
                            l = self._one_step(tr)
                    except errors.OutOfRangeError:
                        break
                    tot += l
                    n += 1
                    if self.ckpt_dir and self.save_checkpoint_secs and time.time() - last_save >= self.save_checkpoint_secs:
                        tr.save(self._ckpt_path())
                        last_save = time.time()
                loss = tot / max(n, 1)
            log.info("End of the epoch %d. loss %.4f", epoch, loss)
        if self.ckpt_dir:
            tr.save(self._ckpt_path())
        self.global_step = tr.global_step
        return loss

    @staticmethod
    def _one_step(tr: Trainer) -> float:
        batch = tr._batch()
        tr.flat_g.zero_()
        loss = tr.step_fn(tr.model, batch)
        loss.backward()
        tr.ar(tr.flat_g, average=True)
        tr.opt.step()
        tr.global_step += 1
        tr.progress.update(loss.detach())
        if tr.ckpt_every and tr.ckpt_path and tr.global_step % tr.ckpt_every == 0:
            tr.save(tr.ckpt_path)
        return float(loss.detach())

    def test(self, dataset, model: torch.nn.Module, metric_fn: Callable, **_ignored) -> float:
        """mean of ``metric_fn(model, batch)`` over one pass of ``dataset`` (averaged over ranks in a distributed job)"""
        tr = self._trainer if self._trainer is not None and self._trainer.model is model else None
        if tr is None:
            tr = Trainer(self._runtime(), model, dataset, lambda m, b: None, allreduce=self.allreduce, log_every=1 << 30)
        v = tr.evaluate(dataset, metric_fn)
        if tr.rt.world > 1:
            v = sum(tr.rt.all_gather_object(float(v))) / tr.rt.world
        log.info("Test metric: %.4f", v)
        return v

    def train_and_evaluate(self, train_dataset, test_dataset, model, step_fn, metric_fn, optimizer=None,
                           learning_rate: Optional[float] = None, epochs: int = 10, **kw):
        loss = self.train(train_dataset, model, step_fn, optimizer, learning_rate, epochs, **kw)
        return loss, self.test(test_dataset, model, metric_fn)

    def save_node_embedding(self, emb_path: str, dataset, model, embed_fn: Callable, block_max_lines: int = 0):
        """``embed_fn(model, batch) -> (ids, embeddings)``; rows ``id \\t v0,v1,...`` in ``<emb_path>.rank<r>``"""
        tr = self._trainer if self._trainer is not None and self._trainer.model is model else \
            Trainer(self._runtime(), model, dataset, lambda m, b: None, allreduce=self.allreduce, log_every=1 << 30)
        tr.export_embeddings(dataset, embed_fn, emb_path, block_max_lines=block_max_lines)

    def save_node_embedding_bigdata(self, emb_path: str, dataset, model, embed_fn: Callable, block_max_lines: int = 100000, **_):
        self.save_node_embedding(emb_path, dataset, model, embed_fn, block_max_lines=block_max_lines)

    def join(self):
        return None


class LocalTrainer(_ExampleTrainer):
    """Single-process training (trainer.py:282-325)."""


class DistTrainer(_ExampleTrainer):
    """Data-parallel training under ``torchrun`` (trainer.py:327-408).  ``cluster_spec / job_name / task_index / worker_count /
    ps_count`` are the reference's constructor arguments: the process group is the cluster here, so they are only checked
    for consistency."""

    def __init__(self, cluster_spec=None, job_name: str = "worker", task_index: Optional[int] = None, worker_count: Optional[int] = None,
                 ps_count: int = 0, **kw):
        super().__init__(**kw)
        from ..nn.utils import SyncBarrierHook, get_rank, get_world_size
        self.is_local = False
        self.cluster_spec, self.job_name, self.ps_count = cluster_spec, job_name, ps_count
        self.task_index = get_rank() if task_index is None else int(task_index)
        self.worker_count = get_world_size() if worker_count is None else int(worker_count)
        self.is_chief = self.task_index == 0
        self.sync_barrier = SyncBarrierHook(self.worker_count, self.is_chief)

    def join(self):
        """block until every worker has finished training (call after ``train``)"""
        self.sync_barrier.end()


class LinkDistTrainer(DistTrainer):
    """Link-prediction driver of the reference's SEAL / ogbl-collab examples (examples/tf/link_trainer.py:38-176): trains on a
    positive/negative edge stream, evaluates Hits@K on held-out positive and negative edges after every epoch, writes scored
    edges with ``predict``."""

    def eval(self, dataset, model, score_fn: Callable):
        """all scores ``score_fn(model, batch) -> [n]`` of one pass over ``dataset`` (numpy)"""
        import numpy as np
        from .. import errors as _errors
        model.eval()
        outs = []
        with torch.no_grad():
            while True:
                try:
                    batch = dataset.next()
                except _errors.OutOfRangeError:
                    break
                outs.append(score_fn(model, batch).detach().float().reshape(-1).cpu().numpy())
        model.train()
        return np.concatenate(outs) if outs else np.zeros(0, dtype=np.float32)

    @staticmethod
    def eval_hits(y_pred_pos, y_pred_neg, k: int) -> dict:
        from ..utils.metrics import hits_at_k
        return {"hits@{}".format(k): hits_at_k(y_pred_pos, y_pred_neg, k)}

    def train_and_eval(self, train_dataset, model, step_fn: Callable, test_dataset, test_neg_dataset, score_fn: Callable,
                       learning_rate: float = 1e-2, epochs: int = 10, hit_K: int = 50, optimizer=None):
        """-> list of per-epoch {"loss", "hits@K"}; ``step_fn(model, batch) -> loss`` consumes whatever ``train_dataset``
        yields (positive + negative edges of one query, or a pair zipped by the caller)"""
        history = []
        for epoch in range(epochs):
            loss = self.train(train_dataset, model, step_fn, optimizer=optimizer, learning_rate=learning_rate, epochs=1)
            pos = self.eval(test_dataset, model, score_fn)
            neg = self.eval(test_neg_dataset, model, score_fn)
            rec = {"loss": loss, **self.eval_hits(pos, neg, hit_K)}
            log.info("Epoch %d: loss %.5f  Test hits@%d: %.4f", epoch, loss, hit_K, rec["hits@{}".format(hit_K)])
            history.append(rec)
        self.join()
        return history

    def predict(self, dataset, model, edge_score_fn: Callable, path: str) -> int:
        """``edge_score_fn(model, batch) -> (src_ids, dst_ids, scores)``; rows ``src \t dst \t score`` in ``<path>.rank<r>``"""
        from .. import errors as _errors
        rt = self._runtime()
        n = 0
        model.eval()
        with open("%s.rank%d" % (path, rt.rank), "w") as f, torch.no_grad():
            f.write("src_id:int64\tdst_id:int64\tscore:float\n")
            while True:
                try:
                    batch = dataset.next()
                except _errors.OutOfRangeError:
                    break
                s, d, sc = edge_score_fn(model, batch)
                for a, b, c in zip(s.reshape(-1).tolist(), d.reshape(-1).tolist(), sc.reshape(-1).tolist()):
                    f.write("%d\t%d\t%.6f\n" % (a, b, c))
                    n += 1
        model.train()
        return n
