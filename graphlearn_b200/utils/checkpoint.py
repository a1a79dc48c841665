"""Checkpoint / resume.

The reference persists only TF model checkpoints and text embedding dumps; sampler state
(traversal cursor, epoch) has empty Save()/Load() stubs
(graphlearn/src/core/operator/graph/node_generator.h:60-66).  Here a checkpoint is complete:
model + optimizer (flat buffers), device RNG (Philox seed, offset), traversal state of every
dataset (epoch, cursor, permutation seed) - enough to resume sampling deterministically.
Embeddings are dumped in the reference's ``id:int64\\temb:string`` dialect
(graphlearn/examples/tf/trainer.py:214-279) by the native writer, so they can be re-ingested as
a node table (e.g. for KNN)."""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch

from ..parallel.runtime import native


def save_checkpoint(path: str, trainer=None, model: Optional[torch.nn.Module] = None, datasets: Optional[Dict] = None,
                    extra: Optional[dict] = None, rank: int = 0):
    """One file per rank (``<path>.rank<r>``): every rank owns its traversal / RNG state."""
    state = {"extra": extra or {}}
    if trainer is not None:
        state["trainer"] = trainer.state_dict()
    if model is not None:
        state["model"] = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    if datasets:
        state["datasets"] = {k: d.state_dict() for k, d in datasets.items()}
    tmp = "%s.rank%d.tmp" % (path, rank)
    torch.save(state, tmp)
    os.replace(tmp, "%s.rank%d" % (path, rank))        # atomic publish


def load_checkpoint(path: str, trainer=None, model: Optional[torch.nn.Module] = None, datasets: Optional[Dict] = None,
                    rank: int = 0, map_location=None) -> dict:
    # payloads are plain dicts / lists / tensors / numbers: refuse to unpickle anything else (a tampered checkpoint
    # must not be able to execute code)
    state = torch.load("%s.rank%d" % (path, rank), map_location=map_location, weights_only=True)
    if trainer is not None and "trainer" in state:
        trainer.load_state_dict(state["trainer"])
    if model is not None and "model" in state:
        model.load_state_dict(state["model"])
    if datasets and "datasets" in state:
        for k, d in datasets.items():
            if k in state["datasets"]:
                d.load_state_dict(state["datasets"][k])
    return state.get("extra", {})


def save_embeddings(path: str, ids: torch.Tensor, emb: torch.Tensor, block_max_lines: int = 0, header: bool = True):
    """``<id>\\t<v0>,<v1>,...`` rows; optionally split into ``<name>_<block>.txt`` files."""
    C = native()
    ids, emb = ids.reshape(-1), emb.reshape(ids.numel(), -1)
    if block_max_lines and ids.numel() > block_max_lines:
        base, ext = os.path.splitext(path)
        for b, s in enumerate(range(0, ids.numel(), block_max_lines)):
            C.save_embeddings("%s_%d%s" % (base, b, ext or ".txt"), ids[s:s + block_max_lines],
                              emb[s:s + block_max_lines], header)
    else:
        C.save_embeddings(path, ids, emb, header)


def embedding_decoder(dim: int):
    """Decoder that re-ingests a dump written by ``save_embeddings`` as a node table."""
    from ..data.decoder import Decoder
    return Decoder(attr_types=["float"] * dim, attr_delimiter=",")
