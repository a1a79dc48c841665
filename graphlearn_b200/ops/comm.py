"""Gradient all-reduce (K8) + fused flat Adam.

``PeerAllReduce`` owns the symmetric staging buffers and cross-GPU flags of the
one-shot peer-memory all-reduce kernel (csrc/comm.cu): every rank pulls all
peers' (scaled) gradients straight from their HBM and reduces in registers -
valid because GNN encoders are a few MB (reference message: 0.74 MB,
graphlearn/examples/pytorch/gcn/train.py:192-193).  ``backend='nccl'`` keeps
the torch.distributed all_reduce as the A/B baseline; on CPU (gloo) it is the
only implementation.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from ..parallel.runtime import MAX_WORLD, Runtime, native

_AR_BLOCKS = 256


class PeerAllReduce:
    def __init__(self, rt: Runtime, numel: int, backend: str = "peer"):
        self.rt = rt
        self.numel = int(numel)
        assert self.numel % 4 == 0, "flat gradient buffers are padded to a multiple of 4"
        self.backend = backend if (rt.is_cuda and rt.world > 1) else "dist"
        if self.backend == "peer":
            self.n_pad = (self.numel + 255) // 256 * 256
            self.stage = rt.symm_empty((2 * self.n_pad,), torch.float32)
            self.flags = rt.symm_empty((_AR_BLOCKS * MAX_WORLD,), torch.int32)
            self.epochs = torch.zeros(_AR_BLOCKS, dtype=torch.int64, device=rt.device)
            self.error = torch.zeros(1, dtype=torch.int32, device=rt.device)
            pad = lambda xs: list(xs) + [0] * (MAX_WORLD - len(xs))  # noqa: E731
            self.desc = torch.tensor([rt.rank, rt.world, self.n_pad] + pad(self.stage.ptrs) + pad(self.flags.ptrs),
                                     dtype=torch.int64)
            rt.barrier()

    def __call__(self, flat_grad: torch.Tensor, average: bool = True):
        W = self.rt.world
        if W == 1:
            return flat_grad
        scale = 1.0 / W if average else 1.0
        if self.backend == "peer":
            native().allreduce_oneshot(self.desc, flat_grad, self.epochs, self.error, scale)
        else:
            dist.all_reduce(flat_grad)
            if average:
                flat_grad.mul_(scale)
        return flat_grad

    def check(self):
        if self.backend == "peer" and int(self.error.item()) != 0:
            raise RuntimeError("peer all-reduce barrier timed out (a rank died or diverged)")


class FlatAdam:
    """Adam over ONE flat fp32 buffer holding every parameter (single kernel,
    device-side step counter => CUDA-graph safe)."""

    def __init__(self, flat_param: torch.Tensor, flat_grad: torch.Tensor, lr=1e-3, betas=(0.9, 0.999), eps=1e-8,
                 weight_decay=0.0):
        self.p, self.g = flat_param, flat_grad
        self.m = torch.zeros_like(flat_param)
        self.v = torch.zeros_like(flat_param)
        self.step_t = torch.zeros(1, dtype=torch.int64, device=flat_param.device)
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay

    def advance(self, rng_state=None):
        """CUDA only: bump the device-resident optimiser step (and RNG offset) - split from ``apply`` so an
        engine can run it on a side stream as soon as its sampling kernels have launched."""
        native().step_advance(rng_state, self.step_t, 1)

    def apply(self):
        """CUDA only: fused Adam over the flat buffers using the already advanced step counter."""
        native().adam_flat(self.p, self.g, self.m, self.v, self.step_t, self.lr, self.betas[0], self.betas[1], self.eps,
                           self.wd)

    def step(self, rng_state=None):
        if self.p.is_cuda:
            self.advance(rng_state)
            self.apply()
            return
        self.step_t += 1
        if rng_state is not None:
            rng_state[1] += 1
        t = float(self.step_t.item())
        b1, b2 = self.betas
        g = self.g + self.wd * self.p if self.wd else self.g
        self.m.mul_(b1).add_(g, alpha=1 - b1)
        self.v.mul_(b2).addcmul_(g, g, value=1 - b2)
        bc1, bc2 = 1 - b1 ** t, 1 - b2 ** t
        denom = (self.v.sqrt() / (bc2 ** 0.5)).add_(self.eps)
        self.p.addcdiv_(self.m, denom, value=-self.lr / bc1)

    def state_dict(self):
        return {"m": self.m.clone(), "v": self.v.clone(), "step": int(self.step_t.item()), "lr": self.lr,
                "betas": self.betas, "eps": self.eps, "wd": self.wd}

    def load_state_dict(self, sd):
        self.m.copy_(sd["m"]); self.v.copy_(sd["v"]); self.step_t.fill_(int(sd["step"]))
        self.lr, self.betas, self.eps, self.wd = sd["lr"], tuple(sd["betas"]), sd["eps"], sd["wd"]


def flatten_module(module: torch.nn.Module):
    """Re-home all parameters (and their .grad) into two flat fp32 buffers
    (every parameter 8-element aligned).  Returns (flat_param, flat_grad)."""
    params = [p for p in module.parameters() if p.requires_grad]
    # every parameter starts on an 8-element (32-byte) boundary: the fused Adam kernel owns 8 consecutive
    # elements per thread and emits one 16-byte chunk of the bf16 weight images from them
    offs, total = [], 0
    for p in params:
        offs.append(total)
        total += (p.numel() + 7) // 8 * 8
    n = total
    dev = params[0].device
    flat_p = torch.zeros(n, dtype=torch.float32, device=dev)
    # 4 spare floats behind the gradients: scratch accumulators (e.g. the loss) that must be zeroed
    # together with the gradients by the single per-step memset of `flat_g.storage`
    g_store = torch.zeros(n + 4, dtype=torch.float32, device=dev)
    flat_g = g_store[:n]
    module._glb_grad_storage = g_store
    with torch.no_grad():
        for p, off in zip(params, offs):
            k = p.numel()
            flat_p[off:off + k].copy_(p.reshape(-1))
            p.data = flat_p[off:off + k].view_as(p)
            p.grad = flat_g[off:off + k].view_as(p)
    module._glb_param_offsets = {id(p): off for p, off in zip(params, offs)}
    return flat_p, flat_g
