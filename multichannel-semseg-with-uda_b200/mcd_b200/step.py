"""The MCD iteration (phase A, B, num_k x C) as one callable, mirroring the reference's inline loop bodies
(adapt_trainer.py:162-212, adapt_mfnet_trainer.py:181-235) over the drop-in modules.

Differences from the reference loop that do NOT change any result (SURVEY.md section 8a16 "legal savings"):
  * phase B back-propagates only into the classifiers: the reference also back-propagates through G and then
    throws those gradients away (`optimizer_g.zero_grad()` at adapt_trainer.py:205 before any use), which is
    ~1 TFLOP of dead work per image pair.  `exact_reference_backward=True` restores the dead work.
  * gradients live in flat per-optimizer buffers (one memset instead of ~250 zero_grad kernels); with
    world_size > 1 the buffers are all-reduced with NCCL in buckets overlapped with wgrad (parallel.GradSync).
  * losses stay on the device; `.item()` is the caller's choice (the reference syncs every phase).
"""
import torch

from . import parallel


def ops_mod():
    from . import ops
    return ops


class MCDStep:
    """method 'MCD' (early fusion): models = (model_g, model_f1, model_f2);
    method 'MFNet': models = (model_g_3ch, model_g_1ch, model_f1, model_f2)."""

    def __init__(self, models, criterion, criterion_d, lr=1e-3, momentum=0.9, weight_decay=2e-5, num_k=4,
                 num_multiply_d_loss=1.0, opt="sgd", exact_reference_backward=False, process_group=None,
                 bucket_mb=25, reuse_target_forward=True, fused_sgd=True, defer_wgrad_reduce=True,
                 logits_dtype=torch.bfloat16):
        from models.model_util import get_optimizer
        self.mfnet = len(models) == 4
        self.gens = list(models[:-2])
        self.f1, self.f2 = models[-2], models[-1]
        self.criterion, self.criterion_d = criterion, criterion_d
        self.num_k, self.mult = num_k, num_multiply_d_loss
        self.exact = exact_reference_backward
        # full-resolution predictions only travel from the heads to the criteria inside the step: bfloat16 halves the
        # traffic of the largest tensors (the modules' drop-in default is fp32, mcd_b200.nn.logits_dtype)
        self.logits_dtype = logits_dtype
        # G is not updated between the phase-B target forward and the first phase-C forward (only optimizer_f
        # steps in between, adapt_trainer.py:195-207), so both forwards are the same computation: run it once,
        # keep its autograd graph for the C[0] backward and let BatchNorm take both momentum updates at once.
        self.reuse_t = reuse_target_forward and not exact_reference_backward
        g_params = [p for m in self.gens for p in m.parameters() if p.requires_grad]
        f_params = [p for p in self.f1.parameters() if p.requires_grad]
        if self.f2 is not self.f1:
            f_params += [p for p in self.f2.parameters() if p.requires_grad]
        self.optimizer_g = get_optimizer(g_params, opt=opt, lr=lr, momentum=momentum, weight_decay=weight_decay)
        self.optimizer_f = get_optimizer(f_params, opt=opt, lr=lr, momentum=momentum, weight_decay=weight_decay)
        self.sync_g = parallel.GradSync(g_params, process_group, bucket_mb)
        self.sync_f = parallel.GradSync(f_params, process_group, bucket_mb)
        self._arena = None
        self.phase_events = None       # a list: (phase name, CUDA event) appended at every phase boundary (bench.py)
        self._packer_g = None
        self._fused_g = None
        self.fused_sgd = fused_sgd     # optimizer_g.step() + weight re-pack as one kernel (plain momentum SGD only)
        # single process + fused SGD: the split-K partial sums of the tcgen05 wgrad kernels are reduced by the
        # optimizer kernel itself (ops.FusedSGD.deferred); with world > 1 the all-reduce needs real gradients
        self.defer_reduce = (fused_sgd and defer_wgrad_reduce and self.sync_g.world == 1
                             and not exact_reference_backward)
        self.world = self.sync_g.world
        if self.world > 1 and hasattr(criterion, "set_process_group"):
            criterion.set_process_group(process_group)   # global sum-of-weights normaliser (DataParallel parity)

    # -- CUDA-graph execution ---------------------------------------------------------------------
    def capture(self, src_imgs, src_lbls, tgt_imgs, warmup=2):
        """Capture the whole iteration (forward, backward, optimizer steps, weight re-packing) in ONE CUDA
        graph: the ~3500 kernel launches of an iteration are then replayed without any host work.  Inputs are
        copied into static buffers by `replay`."""
        from . import abi
        # world > 1: the NCCL bucket all-reduces (side stream, forked / joined with stream waits) and the 4-float
        # normaliser all-reduce are captured as graph nodes too; every rank replays the same sequence.
        dev = src_imgs.device
        if self._packer_g is None and not self._fused_g:
            warmup = max(warmup, 1)        # lazily built host tables (weight re-pack list) need one eager iteration
        self._static = (src_imgs.clone(), src_lbls.clone(), tgt_imgs.clone())
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self(*self._static)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        n0 = abi.launch_count()
        # capture on a HIGH-priority stream: the critical path (forward, dgrad, BatchNorm) then outranks the wgrad
        # kernels that trail on the default-priority side stream when both compete for SMs
        cap_stream = torch.cuda.Stream(dev, priority=-1)
        with torch.cuda.graph(self.graph, stream=cap_stream):
            self._static_out = self(*self._static)
        self.launches_per_replay = abi.launch_count() - n0
        return self

    def replay(self, src_imgs, src_lbls, tgt_imgs):
        s, l, t = self._static
        s.copy_(src_imgs, non_blocking=True)
        l.copy_(src_lbls, non_blocking=True)
        t.copy_(tgt_imgs, non_blocking=True)
        self.graph.replay()
        return self._static_out

    # -- input pipelining: the next batch travels host -> device while the current iteration computes ----------
    def prefetch(self, src_imgs, src_lbls, tgt_imgs):
        """Start the host->device copy of the NEXT batch (pinned host tensors) on a copy stream into staging
        buffers; `replay_prefetched()` consumes it.  The data loader's `.cuda(non_blocking=True)` of the
        reference (adapt_trainer.py:156-160) overlapped with the running iteration."""
        dev = self._static[0].device
        if getattr(self, "_stage", None) is None:
            self._stage = tuple(torch.empty_like(t) for t in self._static)
            self._copy_stream = torch.cuda.Stream(dev)
            self._copied = torch.cuda.Event()
            self._consumed = torch.cuda.Event()
            self._consumed.record(torch.cuda.current_stream(dev))
        cs = self._copy_stream
        cs.wait_event(self._consumed)            # the previous staging contents have been moved into the static inputs
        with torch.cuda.stream(cs):
            for dst, src in zip(self._stage, (src_imgs, src_lbls, tgt_imgs)):
                dst.copy_(src, non_blocking=True)
            self._copied.record(cs)

    def replay_prefetched(self):
        """device->device move of the prefetched batch into the graph's static inputs, then one graph launch."""
        dev = self._static[0].device
        main = torch.cuda.current_stream(dev)
        main.wait_event(self._copied)
        for dst, src in zip(self._static, self._stage):
            dst.copy_(src, non_blocking=True)
        self._consumed.record(main)
        self.graph.replay()
        return self._static_out

    # -- forward helpers -------------------------------------------------------------------------
    def _gen(self, x):
        if not self.mfnet:
            return (self.gens[0](x),)
        # adapt_mfnet_trainer.py:186-187: RGB stream and HHA stream
        return self.gens[0](x[:, :3]), self.gens[1](x[:, 3:])

    def _heads(self, feats):
        return self.f1(*feats), self.f2(*feats)

    def _disc(self, o1, o2):
        d = self.criterion_d(o1, o2)
        return d / self.world if self.world > 1 else d   # SUM all-reduce of gradients => global mean

    def __call__(self, src_imgs, src_lbls, tgt_imgs):
        crit = self.criterion
        from . import ops
        if self._arena is None or self._arena.buf.device != src_imgs.device:
            self._arena = ops.ZeroArena(src_imgs.device)
        prev_arena = ops.set_arena(self._arena)
        from .nn import DirectGrads, logits_dtype
        try:
            self._arena.begin()            # ONE memset for all BatchNorm-statistic / loss accumulators
            with DirectGrads(defer=self.defer_reduce) as self._dg, logits_dtype(self.logits_dtype):
                self._dev = src_imgs.device
                return self._iteration(crit, src_imgs, src_lbls, tgt_imgs)
        finally:
            ops.set_arena(prev_arena)

    def _step_g(self):
        """optimizer_g.step() and the refresh of all packed bf16 weight shadows of G: ONE fused kernel when the
        optimizer is plain momentum SGD (ops.FusedSGD), else optimizer.step() + one multi-tensor re-pack."""
        from .nn import Conv2d
        if self._fused_g is None:
            convs = [m for g in self.gens for m in g.modules() if isinstance(m, Conv2d) and m._packs]
            if self.fused_sgd and ops_mod().FusedSGD.supports(self.optimizer_g):
                self._fused_g = ops_mod().FusedSGD(self.optimizer_g, convs, defer_wgrad_reduce=self.defer_reduce)
            else:
                self._fused_g = False
        if self._fused_g:
            self._fused_g.step()
            return
        self.optimizer_g.step()
        if self._packer_g is None:
            convs = [m for g in self.gens for m in g.modules() if isinstance(m, Conv2d) and m._packs]
            self._packer_g = ops_mod().MultiPacker(convs)
        self._packer_g.repack()

    def _backward(self, loss):
        """loss.backward() + join of the side stream that carries the convolution weight gradients."""
        loss.backward()
        self._dg.join(self._dev)

    def _mark(self, name):
        if self.phase_events is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.phase_events.append((name, e))

    def _iteration(self, crit, src_imgs, src_lbls, tgt_imgs):
        # ---- A: source supervised; updates G, F1, F2
        self._mark("start")
        self.sync_g.zero_and_arm(), self.sync_f.zero_and_arm()
        o1, o2 = self._heads(self._gen(src_imgs))
        loss = crit(o1, src_lbls) + crit(o2, src_lbls)
        self._backward(loss)
        c_loss = loss.detach()
        self.sync_g.wait(), self.sync_f.wait()
        self._step_g(), self.optimizer_f.step()
        self._mark("A")
        # ---- B: classifiers maximise the discrepancy on target; only optimizer_f steps
        self.sync_f.zero_and_arm()
        if self.exact:
            self.sync_g.zero_and_arm(armed=False)
            feats_s, feats_t = self._gen(src_imgs), self._gen(tgt_imgs)
            feats_t_graph = None
        else:
            with torch.no_grad():
                feats_s = self._gen(src_imgs)
            if self.reuse_t:
                from .nn import bn_update_repeat
                with bn_update_repeat(2):
                    feats_t_graph = self._gen(tgt_imgs)
                feats_t = tuple(f.detach() for f in feats_t_graph)
            else:
                feats_t_graph = None
                with torch.no_grad():
                    feats_t = self._gen(tgt_imgs)
        o1, o2 = self._heads(feats_s)
        loss = crit(o1, src_lbls) + crit(o2, src_lbls)
        t1, t2 = self._heads(feats_t)
        loss = loss - self._disc(t1, t2)
        self._backward(loss)
        self.sync_f.wait()
        self.optimizer_f.step()
        self._mark("B")
        # ---- C x num_k: generator minimises the discrepancy; only optimizer_g steps
        #      (classifier gradients of this phase are zeroed before use at adapt_trainer.py:163-164, so unless
        #       exact_reference_backward is set they are not computed at all)
        self.sync_f.disarm()
        if not self.exact:
            for p in self.sync_f.params:
                p.requires_grad_(False)
        for k in range(self.num_k):
            self.sync_g.zero_and_arm()
            feats = feats_t_graph if (k == 0 and feats_t_graph is not None) else self._gen(tgt_imgs)
            feats_t_graph = None
            t1, t2 = self._heads(feats)
            loss = self._disc(t1, t2) * self.mult
            self._backward(loss)
            self.sync_g.wait()
            self._step_g()
            self._mark("C%d" % k)
        if not self.exact:
            for p in self.sync_f.params:
                p.requires_grad_(True)
        d_loss = loss.detach() * (self.world / self.num_k)
        return c_loss, d_loss
