"""Hand-scheduled training step for EgoGraphSAGE (any depth, mean/sum aggregation).

Same math as ``models.graphsage.EgoGraphSAGE`` + ``SageTrainer``'s autograd
path, but forward and backward are an explicit, minimal kernel chain over
pre-allocated buffers (static shapes -> CUDA graph friendly):

  K1    sample hop 1..L of the NEXT batch (parallel branch)            (L launches + 1 counter kernel)
  K6+K7 ONE persistent launch per layer: gather / aggregate / tcgen05 GEMM of every hop pair of the layer
        (multi-segment); the top layer's launch also computes softmax-CE, dlogits and the bias gradient
  for l = L..1:  dW_l = dZ_l^T A_l   split-K tcgen05 kernel, fp32 accumulation straight into the flat gradient buffer;
                                     for l < L, dZ_l = relu'(H_l) * (self + nbr/k rows of dA_{l+1}) is computed by its producers
                 dA_l = dZ_l W_l     the forward kernel with a K-major image of W_l^T                      [l > 1]
  K8    ONE kernel: peer-memory all-reduce (world > 1) + Adam + gradient zeroing + bf16 weight images of the next step

9 launches of this repo's kernels (0 library kernels) for the 2-layer flagship instead of ~60 on the autograd path.
Layer l consumes hop pairs (i, i+1) for i in 0..L-l, exactly the EgoGNN
recursion of graphlearn/python/nn/tf/model/ego_gnn.py:58-110.

Scheduling (see ``_step_body`` / ``capture``): the launches above form a DAG, not a chain - independent ones
run on forked streams, which become parallel branches of the captured CUDA graph; under capture the step is
additionally software pipelined: graph replay t trains on the batch that replay t-1 staged (pinned host ->
device) and sampled on a side branch, and the loss leaves through another branch.  Public API: ``capture()``,
``step(seed_ids_host)`` (end to end), ``step_device()`` (no host traffic), ``from_query`` / ``step_query`` (driven by a
compiled GSL query), ``predict``, ``state_dict``.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from ..ops import comm as comm_ops
from ..ops import rng as rng_ops
from ..ops import sage as sage_ops
from ..ops import sampling as S
from ..parallel.runtime import Runtime, local_table_desc, native
from ..store.shards import CsrShard, NodeTable


class FastSageTrainer:
    def __init__(self, rt: Runtime, nodes: NodeTable, csr: CsrShard, model, fanouts: Sequence[int], batch_size: int,
                 lr: float = 3e-3, strategy: str = "random", use_cuda_graph: bool = True, allreduce: str = "peer",
                 seed: int = 0, allreduce_dtype: str = "fp32"):
        self.ar_bf16 = allreduce_dtype == "bf16"      # stage the gradient exchange in bf16 (halves the NVLink bytes)
        assert rt.is_cuda, "FastSageTrainer is the CUDA engine; use SageTrainer for the portable path"
        self.rt, self.nodes, self.csr, self.model = rt, nodes, csr, model
        self.fanouts = list(fanouts)
        self.L = len(self.fanouts)
        assert self.L == model.num_layers
        self.B = int(batch_size)
        self.strategy = strategy
        self.C = native()
        dev = rt.device
        self.rng = rng_ops.DeviceRng(rt, seed)
        self.flat_p, self.flat_g = comm_ops.flatten_module(model)
        if rt.world > 1:            # replicas start from rank 0's initialisation (what DDP does at construction)
            import torch.distributed as dist
            with torch.no_grad():
                dist.broadcast(self.flat_p, src=0)
        self.opt = comm_ops.FlatAdam(self.flat_p, self.flat_g, lr=lr)
        self.ar = comm_ops.PeerAllReduce(rt, self.flat_g.numel(), backend=allreduce)
        self.seeds = torch.zeros(self.B, dtype=torch.int64, device=dev)
        self.g_store = model._glb_grad_storage            # grads + 4 scratch floats, zeroed by ONE memset
        self.loss = self.g_store[self.flat_g.numel():self.flat_g.numel() + 1]
        # hop sizes n_i = B * prod(k_j, j < i)
        self.n = [self.B]
        for k in self.fanouts:
            self.n.append(self.n[-1] * k)
        convs = list(model.convs)
        self.convs = convs
        self.n_split: List[int] = []           # forward images per layer (N-split when W does not fit next to the A tiles)
        self.n_img: List[int] = []             # padded output columns per image
        for c in convs:
            assert c.agg_type in ("mean", "sum"), "fast engine supports mean / sum aggregation"
            kt = c.weight_p.size(1)
            ns, n_img = 1, sage_ops.pad_n(c.out_dim)
            while not sage_ops.smem_fits(kt, n_img) and ns < 8:          # fewest images whose W slice fits next to the A tiles
                ns += 1
                n_img = (-(-c.out_dim // ns) + 31) // 32 * 32
            is_top = c is convs[-1]
            assert sage_ops.smem_fits(kt, n_img) and (ns == 1 or is_top or ns * n_img == c.out_dim), "layer does not fit the fused kernel"
            assert sage_ops.fused_supported(c.in_self, c.in_nbr, min(c.out_dim, n_img), c.agg_type, max(self.fanouts))
            self.n_split.append(ns)
            self.n_img.append(n_img)
        # per layer l (1-based): segments i = 0..L-l, rows concatenated
        self.seg_off: List[List[int]] = []
        self.H: List[Optional[torch.Tensor]] = []      # layer outputs (bf16; last layer fp32 logits)
        self.Hp: List[torch.Tensor] = []               # ... with the rows padded to the image columns
        self.A: List[torch.Tensor] = []                # saved [self || agg] tiles (bf16)
        self.dZ: List[torch.Tensor] = []
        self.dZp: List[torch.Tensor] = []              # dZ with the row padded to the K padding of the dA GEMM
        self.dA: List[Optional[torch.Tensor]] = []
        for l in range(1, self.L + 1):
            c = convs[l - 1]
            segs = self.L - l + 1
            offs = [0]
            for i in range(segs):
                offs.append(offs[-1] + self.n[i])
            rows = offs[-1]
            self.seg_off.append(offs)
            last = l == self.L
            kt = c.weight_p.size(1)
            hw = self.n_split[l - 1] * self.n_img[l - 1] if last else c.out_dim      # padded logits rows (all images)
            self.Hp.append(torch.zeros(rows, hw, dtype=torch.float32 if last else torch.bfloat16, device=dev))
            self.H.append(self.Hp[-1][:, :c.out_dim])
            self.A.append(torch.zeros(rows, kt, dtype=torch.bfloat16, device=dev))
            self.dZp.append(torch.zeros(rows, sage_ops.pad_k(c.out_dim), dtype=torch.bfloat16, device=dev))
            self.dZ.append(self.dZp[-1][:, :c.out_dim])
            self.dA.append(torch.zeros(rows, kt, dtype=torch.bfloat16, device=dev) if l > 1 else None)
        # persistent bf16 weight images (written by the fused Adam kernel at the end of every step)
        self.img: List[torch.Tensor] = []
        self.img_t: List[Optional[torch.Tensor]] = []
        mats = []
        for l in range(1, self.L + 1):
            c = convs[l - 1]
            n_out, kt = c.weight_p.shape
            N = self.n_split[l - 1] * self.n_img[l - 1]
            self.img.append(torch.zeros(kt * N, dtype=torch.bfloat16, device=dev))
            kpad = self.dZp[l - 1].size(1)
            self.img_t.append(torch.zeros((kt + 255) // 256, kpad * 256, dtype=torch.bfloat16, device=dev) if l > 1 else None)
            mats.append([model._glb_param_offsets[id(c.weight_p)], n_out, kt, self.n_img[l - 1], self.img[-1].data_ptr(),
                         self.img_t[-1].data_ptr() if l > 1 else 0, kpad, 256])
        self._mats = torch.tensor(mats, dtype=torch.int64)
        self.loss_out = torch.zeros(1, dtype=torch.float32, device=dev)
        self.repack()
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self._post_loss_hook = None
        self._pre_opt_join = None
        self._skip_opt = False          # tests: stop before the optimiser to inspect the gradients
        # training branches run at high priority: pending CTAs of the whole-SM persistent kernels must win SM slots
        # against the (low-priority) next-batch sampling branch
        self._sides = (torch.cuda.Stream(priority=-1), torch.cuda.Stream(priority=-1))
        self._side_c = torch.cuda.Stream(priority=-1)
        self._ev_pack = torch.cuda.Event()
        self._ev_zero = torch.cuda.Event()
        self.use_graph = bool(use_cuda_graph)
        self._steps = 0
        self.h_seeds = torch.zeros(self.B, dtype=torch.int64).pin_memory()
        self.h_loss = torch.zeros(1, dtype=torch.float32).pin_memory()
        self._h_seeds2 = [torch.zeros(self.B, dtype=torch.int64).pin_memory() for _ in range(2)]
        self._h_loss2 = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(2)]
        # device-side aliases of the pinned host buffers (cudaHostAlloc memory is UVA-mapped)
        self._h_seeds_dev = [self.C.tensor_from_ptr(t.data_ptr(), [self.B], 2, dev.index) for t in self._h_seeds2]
        self._h_loss_dev = [self.C.tensor_from_ptr(t.data_ptr(), [1], 0, dev.index) for t in self._h_loss2]
        self._e2e_done = [torch.cuda.Event() for _ in range(2)]
        for e in self._e2e_done:
            e.record()

    # ------------------------------------------------------------------ GSL front end
    @classmethod
    def from_query(cls, graph, query, model, lr: float = 3e-3, **kw) -> "FastSageTrainer":
        """Train from a ``gl.Graph`` + GSL query (the API the reference trains through,
        graphlearn/python/gsl/dag_node.py:164-305, gsl/dag_dataset.py:29-97):

            q = g.V("i").batch(1024).shuffle(traverse=True).alias("src") \
                 .outV("e").sample(25).by("random").alias("h1").outV("e").sample(10).by("random").alias("h2").values()
            tr = FastSageTrainer.from_query(g, q, model); tr.capture()
            while True:
                try: loss = tr.step_query()
                except gl.OutOfRangeError: break        # end of epoch

        The query is compiled to a static sampling plan (gsl/compile.py); its hop chain becomes the sampling branch of
        the captured training-step graph, its root traversal (batch / shuffle / epochs) drives ``step_query``."""
        from ..gsl.compile import compilable, compile_query
        from ..gsl.iterators import SeedIterator
        plan = compile_query(query)
        if plan is None or not compilable(graph, plan):
            raise ValueError("query cannot be lowered to a static sampling plan (need V().batch()[.shuffle()] followed by "
                             "outV/inV(...).sample(k).by(strategy) hops over dense-id node types on CUDA)")
        ets = {(h.edge_type, h.direction, h.strategy) for h in plan.hops}
        if len(ets) != 1:
            raise ValueError("the fused engine trains homogeneous chains: every hop must use the same edge type and strategy")
        et, direction, strategy = next(iter(ets))
        store = graph.store
        csr = store.reverse_csr(et) if direction == "in" else store.edges[et]
        if strategy == "in_degree" and direction == "out":
            store.ensure_indegree_weights(et)
        nodes = store.nodes[plan.base_type]
        if any(h.dst_type != plan.base_type for h in plan.hops):
            raise ValueError("the fused engine needs src and dst of the sampled edge type to be the root's node type")
        tr = cls(graph.runtime, nodes, csr, model, plan.fanouts, plan.batch_size, lr=lr, strategy=strategy, **kw)
        tr.plan, tr.graph_api = plan, graph
        rt = graph.runtime
        tab = store.nodes[plan.root_type]
        rows = tab.present.nonzero().flatten() if tab.present is not None else torch.arange(tab.n_local, device=rt.device)
        vids = rows * rt.world + rt.rank
        if plan.base_type != plan.root_type:
            vids = nodes.idmap.to_vid(tab.idmap.to_id(vids))
        tr._q_vids = vids.cpu()
        # full batches only: the step graph has static shapes and the loss is normalised by B (the reference's PyTorch
        # loader also truncates every rank to min(count) // batch_size batches, examples/pytorch/gcn/train.py:174)
        from .. import config as _config
        tr._q_iter = SeedIterator(int(vids.numel()), plan.batch_size, plan.traverse, "cpu",
                                  seed=_config.get().seed + 17 * rt.rank, drop_last=True)
        tr._q_epoch, tr._q_order = -1, None
        tr._q_refresh()
        return tr

    def _q_refresh(self):
        """Per-epoch seed order: the root traversal's permutation applied to this rank's vids once per epoch, so that a
        step only slices it (host work per step stays at a few microseconds)."""
        it = self._q_iter
        th = getattr(self, "_q_thread", None)
        if th is not None:
            th.join()
            self._q_thread = None
        nxt = getattr(self, "_q_next", None)
        self._q_next = None
        if it.strategy == "shuffle" and nxt is not None and nxt[0] == it.epoch and it._perm is None:
            it._perm, self._q_order = nxt[1], nxt[2]      # permutation of this epoch, prepared by the background thread
        elif it.strategy == "shuffle":
            self._q_order = self._q_vids[it.prime()]
        elif it.strategy == "by_order":
            self._q_order = self._q_vids
        else:
            self._q_order = None
        self._q_epoch = it.epoch

    def step_query(self) -> torch.Tensor:
        """One training step on the next seed batch of the compiled GSL query (host traversal -> pinned staging ->
        the step graph).  Raises ``OutOfRangeError`` at the end of an epoch, exactly like ``Dataset.next()``."""
        it = self._q_iter
        if self._q_epoch != it.epoch:
            self._q_refresh()
        if self._q_order is None:                       # 'random' traversal: draws with replacement
            return self.step(self._q_vids[it.next_index()])
        lo = it.cursor
        it.next_index()                                  # advances the cursor / raises OutOfRangeError at the epoch end
        if it.strategy == "shuffle" and lo * 2 > it.n and getattr(self, "_q_thread", None) is None and getattr(self, "_q_next", None) is None:
            self._q_prefetch_epoch(it.epoch + 1)         # randperm of millions of rows costs ~20 ms: off the step path
        return self.step(self._q_order[lo:lo + self.B])

    def _q_prefetch_epoch(self, epoch: int):
        """build the NEXT epoch's permutation (same generator seeding as SeedIterator) on a host thread"""
        import threading
        it = self._q_iter

        def work():
            g = torch.Generator()
            g.manual_seed(it.seed * 1000003 + epoch)
            perm = torch.randperm(it.n, generator=g)
            self._q_next = (epoch, perm, self._q_vids[perm])
        self._q_thread = threading.Thread(target=work, daemon=True)
        self._q_thread.start()

    @property
    def epoch(self) -> int:
        return self._q_iter.epoch

    # ------------------------------------------------------------------ helpers
    def sample(self, seeds: torch.Tensor):
        hops = [seeds]
        cur = seeds
        for i, k in enumerate(self.fanouts):
            nbr, _ = S.sample_neighbors(self.csr, cur, k, self.strategy, want_eids=False, rng=self.rng, salt=i + 1)
            cur = nbr.reshape(-1)
            hops.append(cur)
        return hops

    # ------------------------------------------------------------------ the step
    def _step_body(self, pipe=None):
        """One training step as a small DAG of launches.  Independent work runs on forked streams (graph
        branches under capture) so that kernel-boundary latencies overlap instead of adding up:
            sampling: on main (plain schedule) or, under graph capture, the NEXT batch on its own branch, followed by
                      the rng + optimiser step counters
            main    : one persistent launch per layer (all hop pairs of the layer; the top layer also computes the loss)
            side A  : dW_L GEMM   ||   main : dA_L (forward kernel with the W^T image) -> dW_{L-1} (dZ computed on the fly) ...
        everything joins before ONE fused kernel: gradient all-reduce + Adam + grad zeroing + next step's weight images."""
        C, L = self.C, self.L
        main = torch.cuda.current_stream()
        sA, sB = self._sides
        sC = self._side_c
        # no head work: the gradients were zeroed and the bf16 weight images refreshed by the fused Adam kernel of
        # the previous step (``repack()`` covers the very first step and externally modified weights)
        if pipe is None:
            # plain schedule: sample this step's batch, then train on it
            seeds = self.seeds
            hops = self.sample(seeds)
            # counters advance once the sampling kernels have consumed the RNG offset; needed again only by Adam
            self.opt.advance(self.rng.state)
        else:
            # pipelined schedule (graph capture): train on the batch sampled by the PREVIOUS replay while a
            # forked branch stages + samples the NEXT batch (the reference's sampling || training pipeline:
            # dag_scheduler.cc:51-62 / dag_dataset.cc:38-43, here inside one CUDA graph)
            cur, nxt, host_slot = pipe
            hops = self._hops[cur]
            seeds = hops[0]
            sS = self._sample_stream
            sS.wait_stream(main)
            with torch.cuda.stream(sS):
                if host_slot is not None:
                    # zero-copy staging: a tiny copy KERNEL reads the UVA-mapped pinned buffer over PCIe
                    C.copy_i64(self._hops[nxt][0], self._h_seeds_dev[host_slot])
                nb = self._hops[nxt][0]
                for i, k in enumerate(self.fanouts):
                    S.sample_neighbors(self.csr, nb, k, self.strategy, want_eids=False, rng=self.rng, salt=i + 1,
                                       out=self._hops[nxt][i + 1])
                    nb = self._hops[nxt][i + 1]
                self.opt.advance(self.rng.state)
        # ---- forward: ONE persistent launch per layer covering all of its hop-pair segments; the top layer's
        # launch also computes the loss, dlogits and the bias gradient in its epilogue (fused CE)
        top = self.convs[L - 1]
        for l in range(1, L + 1):
            c = self.convs[l - 1]
            last = l == L
            ns = self.n_split[l - 1]
            N = self.n_img[l - 1]
            n_out = c.out_dim if ns == 1 else N
            img = self.img[l - 1]
            fuse_ce = last and ns == 1 and N <= 64
            offs = self.seg_off[l - 1]
            mode = sage_ops.MODE[c.agg_type]
            nseg = L - l + 1
            hb = self.H[l - 1] if ns == 1 else self.Hp[l - 1]
            outs = [hb[offs[i]:offs[i + 1], :n_out] for i in range(nseg)]
            asv = [self.A[l - 1][offs[i]:offs[i + 1]] for i in range(nseg)]
            Ms, ks = [self.n[i] for i in range(nseg)], [self.fanouts[i] for i in range(nseg)]
            ce = []
            if fuse_ce:
                ce = [self.nodes.labels.local, seeds, self.loss, self.dZ[L - 1],
                      top.bias.grad if top.bias is not None else None]
            if l == 1:
                d = self.nodes.feat_desc
                C.sage_fused_multi(d, d, [hops[i] for i in range(nseg)], [hops[i + 1] for i in range(nseg)],
                                   [0] * nseg, [0] * nseg, Ms, ks, outs, asv, mode, img, c.bias, N, n_out,
                                   not last, not last, 0, ce, self.rt.world, ns)
            else:
                po = self.seg_off[l - 2]
                d = local_table_desc(self.H[l - 2])
                C.sage_fused_multi(d, d, [None] * nseg, [None] * nseg, [po[i] for i in range(nseg)],
                                   [po[i + 1] for i in range(nseg)], Ms, ks, outs, asv, mode, img, c.bias, N,
                                   n_out, not last, not last, 0, ce, self.rt.world, ns)
            if last and not fuse_ce:
                # wide / split top layer: the loss runs as its own kernel on the (padded) logits
                C.softmax_ce(self.H[L - 1], self.nodes.labels.local, seeds, self.rt.world, self.loss, self.dZ[L - 1],
                             top.bias.grad if top.bias is not None else None)
        if self._post_loss_hook is not None:
            self._post_loss_hook()          # e2e graph capture: fork the loss D2H here, parallel to the backward
        # ---- backward: all GEMMs on tcgen05 (csrc/sage_bwd.cu), no library kernels
        #   top layer   : dW_L = dZ_L^T A_L with the dense dZ_L written by the fused loss epilogue
        #   layer l < L : dW_l = dZ_l^T A_l where dZ_l = relu'(H_l) * (self + 1/k neighbour rows of dA_{l+1}) is
        #                 computed inside the GEMM's producer warps (never stored for l = 1); bias grads fall out
        #   dA_l = dZ_l . W_l  (l >= 2) on the persistent forward kernel with a K-major image of W_l^T
        for l in range(L, 0, -1):
            c = self.convs[l - 1]
            a = self.A[l - 1]
            if l == L:
                sA.wait_stream(main)
                with torch.cuda.stream(sA):            # dW_L runs beside dA_L
                    C.sage_bwd_dw(self.dZ[l - 1], None, [], [], [], [], [], [], 0, None, a, c.weight_p.grad, None,
                                  c.out_dim, [])
            else:
                nxt = self.convs[l]
                da = self.dA[l]
                offs, po = self.seg_off[l], self.seg_off[l - 1]
                nseg = L - l + 1
                kp_self_n, _ = sage_ops.padded_dims(nxt.in_self, nxt.in_nbr, nxt.agg_type)
                r0 = [po[s] for s in range(nseg)]
                r1 = [po[s + 1] for s in range(nseg)]
                selfs = [da[offs[s]:offs[s + 1]] if s <= L - l - 1 else None for s in range(nseg)]
                nbrs = [da[offs[s - 1]:offs[s]] if s >= 1 else None for s in range(nseg)]
                ks = [self.fanouts[s - 1] if s >= 1 else 1 for s in range(nseg)]
                scales = [((1.0 / self.fanouts[s - 1]) if nxt.agg_type == "mean" else 1.0) if s >= 1 else 1.0
                          for s in range(nseg)]
                C.sage_bwd_dw(None, self.H[l - 1], r0, r1, selfs, nbrs, ks, scales, kp_self_n,
                              self.dZ[l - 1] if l >= 2 else None, a, c.weight_p.grad,
                              c.bias.grad if c.bias is not None else None, c.out_dim, [])
            if l >= 2:
                # dA_l [rows, K_total] = dZ_l [rows, n_out] . W_l, one launch per 256-column block
                dzp = self.dZp[l - 1]
                d = local_table_desc(dzp)
                wt = self.img_t[l - 1]
                da = self.dA[l - 1]
                rows = da.size(0)
                # one launch: CTA b multiplies by image b % n_imgs and writes the columns [256 img, 256 img + 256)
                n_imgs = wt.size(0)
                cols = min(256, da.size(1))
                if l == L:
                    # dW_L (<= 32 whole-SM CTAs for a 1024-row top layer) runs beside this launch: leave it its SMs
                    C.sage_set_max_ctas(max(C.sm_count() - 32, C.sm_count() // 2))
                C.sage_fused_multi(d, d, [None], [None], [0], [0], [rows], [0], [da[:, :cols]], [None], 0, wt, None, 256, cols,
                                   False, True, 0, [], 1, n_imgs)
                C.sage_set_max_ctas(0)
        # ---- join, gradient all-reduce + fused optimiser (Adam + zero grads + next step's weight images)
        main.wait_stream(sA)
        if pipe is not None:
            main.wait_stream(self._sample_stream)
        if self._pre_opt_join is not None:
            main.wait_stream(self._pre_opt_join)       # the loss must have left before the optimiser clears it
        o, ar = self.opt, self.ar
        if self._skip_opt:
            ar(self.flat_g, average=True)
            return
        if ar.backend == "peer":
            # ONE kernel: stage grads (scaled, optionally bf16) -> cross-GPU flag barrier -> pull + reduce every peer's
            # slice in registers -> Adam -> zero grads -> next step's weight images (K8 fused with cast/scale + optimiser)
            C.adam_pack(self.flat_p, self.g_store, o.m, o.v, o.step_t, o.lr, o.betas[0], o.betas[1], o.eps, o.wd,
                        self.loss_out, self._mats, ar.desc, ar.epochs, ar.error, 1.0 / self.rt.world, self.ar_bf16)
        else:
            ar(self.flat_g, average=True)
            C.adam_pack(self.flat_p, self.g_store, o.m, o.v, o.step_t, o.lr, o.betas[0], o.betas[1], o.eps, o.wd,
                        self.loss_out, self._mats, None, None, None, 1.0, False)

    def repack(self):
        """Rebuild the bf16 weight images from the fp32 master weights (after init / load_state_dict / any
        external modification of the parameters)."""
        for l in range(1, self.L + 1):
            c = self.convs[l - 1]
            ns, N = self.n_split[l - 1], self.n_img[l - 1]
            w = c.weight_p.detach().contiguous()
            if ns > 1:
                wpad = torch.zeros(ns * N, w.size(1), device=w.device)
                wpad[:w.size(0)] = w
                self.img[l - 1].copy_(torch.cat([self.C.pack_weight_f32(wpad[i * N:(i + 1) * N].contiguous(), N, False)[0] for i in range(ns)]))
            else:
                self.img[l - 1].copy_(self.C.pack_weight_f32(w, N, False)[0])
            if l > 1:
                self.img_t[l - 1].copy_(self.C.pack_weight_t(c.weight_p.detach().contiguous(), self.dZp[l - 1].size(1), 256))

    # ------------------------------------------------------------------ graph / public step (same API as SageTrainer)
    def capture(self, warmup: int = 3):
        """Warm up eagerly, then capture four graphs (device-only pair + end-to-end pair, one per hop-buffer
        parity); the whole public ``step()`` - H2D of the seeds, sampling, training, D2H of the loss - is then
        ONE graph launch."""
        if not self.use_graph or self.graph is not None:
            return
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._step_body()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.rt.barrier()
        # Four graphs over one memory pool, all software pipelined: graph i trains on hop buffers i and
        # samples the next batch into hop buffers i^1 on a forked branch.
        #   device-only pair : next seeds = whatever hop buffer i^1 holds (re-sampled with a fresh RNG offset)
        #   end-to-end pair  : + next seeds copied from pinned host slot i^1 at the head of the sampling branch
        #                      + loss -> pinned host slot i forked right after the loss kernel
        # No PCIe round trip and no sampling kernel sits on the critical path (tools/diag_e2e.py).
        self._sample_stream = torch.cuda.Stream(priority=0)
        cap_stream = torch.cuda.Stream(priority=-1)
        self._hops = [[torch.zeros(n, dtype=torch.int64, device=self.rt.device) for n in self.n] for _ in range(2)]
        self._hops[0][0] = self.seeds                      # `tr.seeds` stays the handle of buffer 0
        self._seeds_bufs = [self._hops[0][0], self._hops[1][0]]
        self._prime_hops(self.seeds)
        self._dev_graphs, self._e2e_graphs = [], []
        side_out = torch.cuda.Stream()
        pool = None
        for e2e in (False, True):
            for i in range(2):
                gi = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gi, pool=pool, stream=cap_stream):
                    main = torch.cuda.current_stream()
                    if e2e:
                        def hook(i=i, main=main):
                            side_out.wait_stream(main)
                            with torch.cuda.stream(side_out):
                                self._h_loss_dev[i].copy_(self.loss)
                        self._post_loss_hook = hook
                        self._pre_opt_join = side_out
                    self._step_body(pipe=(i, i ^ 1, (i ^ 1) if e2e else None))
                    self._post_loss_hook = None
                    self._pre_opt_join = None
                    if e2e:
                        main.wait_stream(side_out)
                pool = gi.pool() if pool is None else pool
                (self._e2e_graphs if e2e else self._dev_graphs).append(gi)
        self.graph = self._dev_graphs[0]
        self._primed = False
        torch.cuda.synchronize()
        self.rt.barrier()

    def _prime_hops(self, seeds: torch.Tensor):
        """Fill BOTH hop buffers with a sample of `seeds` so that whichever graph replays first finds a batch."""
        hops = self.sample(seeds)
        for j in range(2):
            for dst, src in zip(self._hops[j], hops):
                if dst.data_ptr() != src.data_ptr():
                    dst.copy_(src)

    def step_device(self, next_seeds: Optional[torch.Tensor] = None):
        """Device-only step (no host traffic): graph replay on the resident seed buffers.  ``next_seeds`` (device
        int64 [B]) are the seeds of the batch the replay samples for the FOLLOWING step (software pipeline); without
        them the resident seeds are re-sampled with a fresh RNG offset."""
        if self.graph is not None:
            i = self._steps & 1
            if not self._primed:
                self._prime_hops(self._hops[i][0] if next_seeds is None else next_seeds)
                self._primed = True
            if next_seeds is not None:
                self.C.copy_i64(self._hops[i ^ 1][0], next_seeds)
            self._dev_graphs[i].replay()
        else:
            if next_seeds is not None:
                self.seeds.copy_(next_seeds)
            self._step_body()
        self._steps += 1

    def step(self, seed_ids_host: torch.Tensor) -> torch.Tensor:
        """End-to-end step through the public path: host seed ids -> (pinned staging) -> device,
        train, loss -> pinned host.  Asynchronous: synchronise before reading the returned tensor.
        With CUDA graphs the copies are nodes of the step graph (double-buffered staging)."""
        if self.graph is not None and getattr(self, "_e2e_graphs", None):
            # Pipelined: this call stages `seed_ids_host` (H2D happens inside the graph, in parallel with the
            # training step) and trains on the batch staged by the PREVIOUS call; the returned pinned tensor
            # receives that step's loss.  The very first call also places its batch on the device directly.
            i = self._steps & 1
            if not self._primed:
                self._prime_hops(seed_ids_host.to(self.rt.device))
                self._primed = True
            self._e2e_done[i].synchronize()    # step t-2 (same graph) was the last reader of staging slot i^1
            self._h_seeds2[i ^ 1].copy_(seed_ids_host)
            self._e2e_graphs[i].replay()
            self._e2e_done[i].record()
            self._steps += 1
            self.h_loss = self._h_loss2[i]
            return self.h_loss
        self.h_seeds.copy_(seed_ids_host)
        self.seeds.copy_(self.h_seeds, non_blocking=True)
        self.step_device()
        self.h_loss.copy_(self.loss_out, non_blocking=True)
        return self.h_loss

    def synchronize(self):
        torch.cuda.synchronize()

    @torch.no_grad()
    def predict(self, seeds: torch.Tensor) -> torch.Tensor:
        self.model.eval()
        hops = self.sample(seeds)
        out = self.model.forward_store(self.nodes, hops, self.fanouts)
        self.model.train()
        return out

    def state_dict(self):
        return {"model": self.flat_p.clone(), "opt": self.opt.state_dict(), "rng": self.rng.state_dict(),
                "steps": self._steps}

    def load_state_dict(self, sd):
        self.flat_p.copy_(sd["model"])
        self.repack()
        self.opt.load_state_dict(sd["opt"])
        self.rng.load_state_dict(sd["rng"])
        self._steps = int(sd["steps"])
