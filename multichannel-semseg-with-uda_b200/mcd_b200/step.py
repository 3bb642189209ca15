"""The MCD iteration (phase A, B, num_k x C) as one callable, mirroring the reference's inline loop bodies over the
drop-in modules:

    early fusion   adapt_trainer.py:162-212                      MCDStep((model_g, model_f1, model_f2), ...)
    MFNet          adapt_mfnet_trainer.py:181-235                MCDStep((model_g_3ch, model_g_1ch, model_f1, model_f2), ...)
    seg + HHA      adapt_multitask_trainer.py:194-262            MCDStep.multitask(model_enc, model_dec, triple=False, ...)
    seg+HHA+bd     adapt_triple_multitask_trainer.py:202-287     MCDStep.multitask(model_enc, model_dec, triple=True, ...)

Differences from the reference loops that do NOT change any result (SURVEY.md section 8a16 "legal savings"):
  * phase B back-propagates only into the classifiers / decoders: the reference also back-propagates through the
    generator / encoder and then throws those gradients away (`optimizer_g.zero_grad()` at adapt_trainer.py:205
    before any use), which is ~1 TFLOP of dead work per image pair.  `exact_reference_backward=True` restores it.
  * the triple trainer computes the depth and boundary losses of phase B and does not use them
    (adapt_triple_multitask_trainer.py:256-276): only the weighted segmentation term is back-propagated; the depth
    decoder still runs forward (no graph) because its train-mode BatchNorm statistics take an update there.
  * gradients live in flat per-optimizer buffers (one memset instead of ~250 zero_grad kernels); with
    world_size > 1 the buffers are all-reduced with NCCL in buckets overlapped with wgrad (parallel.GradSync).
  * losses stay on the device; `.item()` is the caller's choice (the reference syncs every phase).
Parameters that receive no gradient in a phase (unused decoders, adapt_triple_multitask_trainer.py:276; `nmlrgr_dec`
always) are skipped by the optimizer, as torch.optim does for `grad is None` (torch >= 2.0 `zero_grad()`, which is what
the oracle is pinned against; torch 0.4.1 would still apply momentum / weight decay to them).
"""
import torch

from . import parallel


def ops_mod():
    from . import ops
    return ops


def _detach(t):
    """detach a feature structure (tensor / tuple / dict) keeping the IEEE-half originals attached (mcd_b200.ops)."""
    if isinstance(t, dict):
        return {k: _detach(v) for k, v in t.items()}
    if isinstance(t, (tuple, list)):
        return tuple(_detach(v) for v in t)
    d = t.detach()
    tw = getattr(t, "_mcd_h16", None)
    if tw is not None:
        d._mcd_h16 = tw
    return d


def _tree(fn, t):
    """apply fn to every tensor of a (nested) tuple of tensors - the uint8 planes of an InputPipeline batch"""
    if isinstance(t, (tuple, list)):
        return tuple(_tree(fn, v) for v in t)
    return fn(t)


def _tree_copy(dst, src):
    if isinstance(dst, (tuple, list)):
        for d, s_ in zip(dst, src):
            _tree_copy(d, s_)
    else:
        dst.copy_(src, non_blocking=True)


def _first(t):
    return _first(t[0]) if isinstance(t, (tuple, list)) else t


def _streams(x):
    """generator inputs of a batch: a pre-transformed pipeline.Batch carries them, a loader tensor is sliced"""
    return x.streams if hasattr(x, "streams") else None


# ---- the four trainers as "tasks": what differs between them is how features and the three objectives are formed --
class _EarlyFusionTask:
    """adapt_trainer.py:162-212 (one generator) / adapt_mfnet_trainer.py:181-235 (RGB + HHA generators)."""
    a_uses_target = False

    def __init__(self, models, criterion, criterion_d, mult):
        models = [parallel.unwrap(m) for m in models]       # is_data_parallel=True wrappers: MCDStep syncs itself
        self.gens, (self.f1, self.f2) = list(models[:-2]), models[-2:]
        self.clfs = [self.f1] if self.f2 is self.f1 else [self.f1, self.f2]
        self.mfnet = len(self.gens) == 2
        self.criterion, self.criterion_d = criterion, criterion_d
        # adapt_mfnet_trainer.py:233 does NOT multiply the phase-C loss by num_multiply_d_loss (adapt_trainer.py:210 does)
        self.mult = 1.0 if self.mfnet else mult

    def features(self, x):
        st = _streams(x)
        if not self.mfnet:
            return (self.gens[0](x if st is None else st[0]),)
        if st is not None:
            return self.gens[0](st[0]), self.gens[1](st[1])
        return self.gens[0](x[:, :3]), self.gens[1](x[:, 3:])       # adapt_mfnet_trainer.py:186-187

    # The classifier heads only feed the criteria: head + softmax + loss + gradients run as ONE kernel that never
    # materialises the full-resolution logits (mcd_b200/headloss.py) whenever the modules are the reference's plain
    # learned-deconvolution heads and the criteria CrossEntropyLoss2d / Diff2d; anything else takes the module path.
    def _fused_inputs(self, head, feats, n_heads):
        import loss as _loss
        from . import headloss
        if not headloss.enabled():
            return None
        crit_ok = type(self.criterion) is _loss.CrossEntropyLoss2d if n_heads == 1 else type(self.criterion_d) is _loss.Diff2d
        hi = headloss.classifier_inputs(head, feats) if crit_ok else None
        if hi is None:
            return None
        want_dw = any(w.requires_grad for w in hi[1]) and torch.is_grad_enabled()
        if not headloss.fits(n_heads, len(hi[0]), hi[0][0].shape[1], want_dw, False):
            return None
        return hi

    def _ce_head(self, head, feats, lbls, hi=False):
        from . import headloss
        if hi is False:
            hi = self._fused_inputs(head, feats, 1)
        if hi is None:
            return self.criterion(head(*feats), lbls)
        c = self.criterion
        return headloss.head_ce2d(hi[0], hi[1], lbls, c.nll_loss.weight, c.ignore_index, c.size_average)

    def _ce(self, feats, lbls):
        from . import headloss
        if self.f2 is not self.f1:
            ha, hb = self._fused_inputs(self.f1, feats, 1), self._fused_inputs(self.f2, feats, 1)
            # both classifiers on the same score maps and labels: ONE launch (mode 2) when it fits
            if (ha is not None and hb is not None and all(p is q for p, q in zip(ha[0], hb[0]))
                    and headloss.fits(2, len(ha[0]), ha[0][0].shape[1],
                                      torch.is_grad_enabled() and any(w.requires_grad for w in ha[1] + hb[1]), False)):
                c = self.criterion
                return headloss.head_ce2d_pair(ha[0], ha[1], hb[1], lbls, c.nll_loss.weight, c.ignore_index, c.size_average)
            # (ver2 heads: each classifier's own `seg` convolution has already produced its score map - reuse it)
            return self._ce_head(self.f1, feats, lbls, ha) + self._ce_head(self.f2, feats, lbls, hb)
        return self._ce_head(self.f1, feats, lbls) + self._ce_head(self.f2, feats, lbls)

    def loss_a(self, fs, ft, src, lbls, tgt):
        return self._ce(fs, lbls)

    def loss_b(self, fs, ft, src, lbls, tgt):
        return self._ce(fs, lbls) - self.disc(ft)

    def disc(self, ft):
        from . import headloss
        ha, hb = self._fused_inputs(self.f1, ft, 2), self._fused_inputs(self.f2, ft, 2)
        if ha is None or hb is None:
            return self.criterion_d(self.f1(*ft), self.f2(*ft))
        return headloss.head_diff2d(ha[0], ha[1], hb[0], hb[1])


class _MultiTaskTask:
    """adapt_multitask_trainer.py:194-262 (triple=False) / adapt_triple_multitask_trainer.py:202-287 (triple=True)."""
    a_uses_target = True

    def __init__(self, model_enc, model_dec, triple, mult):
        model_enc, model_dec = parallel.unwrap(model_enc), parallel.unwrap(model_dec)
        self.gens, self.clfs = [model_enc], [model_dec]
        self.enc, self.dec, self.triple, self.mult = model_enc, model_dec, triple, mult

    def features(self, x):
        st = _streams(x)
        return self.enc(x[:, :3] if st is None else st[0])

    @staticmethod
    def _rest(x):
        """channels 3.. of the loader tensor (HHA regression target [, boundary map])"""
        return x[:, 3:] if _streams(x) is None else x.aux

    def loss_a(self, fs, ft, src, lbls, tgt):
        rest = self._rest(src)
        if self.triple:
            terms = self.dec.get_loss(fs, lbls, rest[:, :-1], rest[:, -1:], separately_returning=True)
        else:
            terms = self.dec.get_loss(fs, lbls, rest, separately_returning=True)
        return sum(terms) + self.dec.get_depth_loss(ft, self._rest(tgt))

    def loss_b(self, fs, ft, src, lbls, tgt):
        if self.triple:      # :274-276 loss = src_semseg_loss - tgt_discrepancy (depth / boundary terms are dead)
            with torch.no_grad():           # ... but get_loss() ran the depth decoder in train mode: its BatchNorm
                self.dec.depth_forward(fs)  # running statistics take that update (forward only, 25 GF / image)
            return self.dec.get_weighted_semseg_loss(fs, lbls) - self.disc(ft)
        semseg, depth = self.dec.get_loss(fs, lbls, self._rest(src), separately_returning=True)
        return semseg + depth + self.dec.get_depth_loss(ft, self._rest(tgt)) - self.disc(ft)    # adapt_multitask_trainer.py:220

    def disc(self, ft):
        return self.dec.get_cls_descrepancy(ft)


class _SegBDTask(_MultiTaskTask):
    """adapt_segbd_multitask_trainer.py:193-255: segmentation (two MCD classifiers) + boundary decoder whose boundary
    ground truth is derived from the segmentation labels.  Phase A: get_loss(src) [+ scale_bd_loss * a pseudo-boundary
    term on the target features once the trainer is past --boundary_loss_converging_epoch: `pseudo` = "pred_seg"
    (--add_pred_seg_boundary_loss) or "seg2bd" (--use_seg2bd_conv)]; phase B: src_semseg_loss - discrepancy (the
    boundary term get_loss() also evaluates there is dead and has no state).  The target forward of phase A always
    runs (BatchNorm running statistics), without autograd when no pseudo term reads it."""
    a_uses_target = True

    def __init__(self, model_enc, model_dec, mult, pseudo=None, scale_bd_loss=1.0):
        super().__init__(model_enc, model_dec, True, mult)
        assert pseudo in (None, "pred_seg", "seg2bd")
        self.pseudo, self.scale = pseudo, scale_bd_loss
        self.a_target_forward_only = pseudo is None

    def loss_a(self, fs, ft, src, lbls, tgt):
        loss = sum(self.dec.get_loss(fs, lbls, separately_returning=True))
        if self.pseudo == "pred_seg":
            loss = loss + self.dec.get_psuedo_boundary_loss(ft, separately_returning=False) * self.scale
        elif self.pseudo == "seg2bd":
            loss = loss + self.dec.get_boundary_loss_by_extra_conv(ft, separately_returning=False) * self.scale
        return loss

    def loss_b(self, fs, ft, src, lbls, tgt):
        return self.dec.get_weighted_semseg_loss(fs, lbls) - self.disc(ft)


class MCDStep:
    """method 'MCD' (early fusion): models = (model_g, model_f1, model_f2);
    method 'MFNet': models = (model_g_3ch, model_g_1ch, model_f1, model_f2);
    multitask trainers: MCDStep.multitask(model_enc, model_dec, triple=...).
    Call it with (src_imgs, src_lbls, tgt_imgs) exactly as the trainers' loops receive them - or, with
    input_pipeline = mcd_b200.pipeline.InputPipeline(mode), with the loader's DECODED uint8 arrays: src_imgs / tgt_imgs =
    (rgb [N,H,W,3], hha [N,H,W,3][, boundary [N,H,W]]) and src_lbls uint8 [N,H,W]; ToTensor / Normalize / concat /
    ReLabel (transform.py:302-325) then run on the GPU as the first kernels of the (captured) iteration."""

    def __init__(self, models, criterion, criterion_d, lr=1e-3, momentum=0.9, weight_decay=2e-5, num_k=4,
                 num_multiply_d_loss=1.0, opt="sgd", exact_reference_backward=False, process_group=None,
                 bucket_mb=25, reuse_target_forward=True, fused_sgd=True, defer_wgrad_reduce=True,
                 logits_dtype=torch.bfloat16, task=None, input_pipeline=None):
        from models.model_util import get_optimizer
        self.task = task if task is not None else _EarlyFusionTask(models, criterion, criterion_d, num_multiply_d_loss)
        self.gens, self.clfs = self.task.gens, self.task.clfs
        self.mfnet = getattr(self.task, "mfnet", False)
        self.num_k, self.mult = num_k, self.task.mult
        self.exact = exact_reference_backward
        # full-resolution predictions only travel from the heads to the criteria inside the step: bfloat16 halves the
        # traffic of the largest tensors (the modules' drop-in default is fp32, mcd_b200.nn.logits_dtype)
        self.logits_dtype = logits_dtype
        # G is not updated between the phase-B target forward and the first phase-C forward (only optimizer_f
        # steps in between, adapt_trainer.py:195-207), so both forwards are the same computation: run it once,
        # keep its autograd graph for the C[0] backward and let BatchNorm take both momentum updates at once.
        self.reuse_t = reuse_target_forward and not exact_reference_backward
        g_params = [p for m in self.gens for p in m.parameters() if p.requires_grad]
        f_params = [p for m in self.clfs for p in m.parameters() if p.requires_grad]
        self.optimizer_g = get_optimizer(g_params, opt=opt, lr=lr, momentum=momentum, weight_decay=weight_decay)
        self.optimizer_f = get_optimizer(f_params, opt=opt, lr=lr, momentum=momentum, weight_decay=weight_decay)
        self.sync_g = parallel.GradSync(g_params, process_group, bucket_mb)
        self.sync_f = parallel.GradSync(f_params, process_group, bucket_mb)
        self._arena = None
        self.phase_events = None       # a list: (phase name, CUDA event) appended at every phase boundary (bench.py)
        self._packers, self._fused = {}, {}
        self.fused_sgd = fused_sgd     # optimizer.step() + weight re-pack as one kernel (plain momentum SGD only)
        # single process + fused SGD: the split-K partial sums of the tcgen05 wgrad kernels are reduced by the
        # optimizer kernel itself (ops.FusedSGD.deferred); with world > 1 the all-reduce needs real gradients
        # (and a weight gradient may be left in split form only once per step: the multitask trainers run the encoder
        # twice in phase A, source and target, so their gradients accumulate in param.grad instead)
        # a generator that runs its convolutions more than once per forward (FuseDRNSegBase: RGB and HHA through the same
        # trunk) accumulates its weight gradients in param.grad as well, and its BatchNorm layers see two different
        # batches per forward - folding the momentum updates of the shared B / C[0] target forward would reorder them
        shared = any(getattr(m, "_mcd_shared_weights", False) for m in self.gens)
        if shared:
            self.reuse_t = False
        self.defer_reduce = (fused_sgd and defer_wgrad_reduce and self.sync_g.world == 1 and not shared
                             and not exact_reference_backward and not self.task.a_uses_target)
        self.world = self.sync_g.world
        self.group = process_group
        if self.world > 1:
            # every criterion returns this rank's SHARE of the global-batch loss (loss.set_process_group): the
            # SUM-all-reduced gradients then equal the gradients nn.DataParallel computes on the gathered outputs
            import loss as _loss
            _loss.set_process_group(process_group)
        self.graph = None
        self._captured_lr = None
        self.pipeline = input_pipeline

    @classmethod
    def multitask(cls, model_enc, model_dec, triple=True, num_multiply_d_loss=1.0, **kw):
        """adapt_multitask_trainer.py (triple=False) / adapt_triple_multitask_trainer.py (triple=True): the criteria
        live inside model_dec (semseg_criterion / discrepancy_criterion given to get_*multitask_models)."""
        return cls(None, None, None, task=_MultiTaskTask(model_enc, model_dec, triple, num_multiply_d_loss), **kw)

    @classmethod
    def segbd(cls, model_enc, model_dec, num_multiply_d_loss=1.0, pseudo=None, scale_bd_loss=1.0, **kw):
        """adapt_segbd_multitask_trainer.py: get_segbd_multitask_models' encoder + MCDSegBDMultiTaskDecoder."""
        return cls(None, None, None, task=_SegBDTask(model_enc, model_dec, num_multiply_d_loss, pseudo, scale_bd_loss),
                   **kw)

    # -- CUDA-graph execution ---------------------------------------------------------------------
    def capture(self, src_imgs, src_lbls, tgt_imgs, warmup=2):
        """Capture the whole iteration (forward, backward, optimizer steps, weight re-packing) in ONE CUDA
        graph: the kernel launches of an iteration are then replayed without any host work.  Inputs are
        copied into static buffers by `replay`."""
        from . import abi
        # world > 1: the NCCL bucket all-reduces (side stream, forked / joined with stream waits) and the scalar
        # all-reduces of the criteria are captured as graph nodes too; every rank replays the same sequence.
        dev = _first(src_imgs).device
        if not self._fused:
            warmup = max(warmup, 1)        # lazily built host tables (optimizer / re-pack lists) need one eager iteration
        self._static = _tree(torch.clone, (src_imgs, src_lbls, tgt_imgs))
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
        self._captured_lr = self._hyper_snapshot()
        return self

    def _hyper_snapshot(self):
        return tuple((g["lr"], g.get("momentum", 0), g.get("weight_decay", 0))
                     for o in (self.optimizer_g, self.optimizer_f) for g in o.param_groups)

    def _before_replay(self):
        """learning-rate schedules (util.adjust_learning_rate on optimizer_g / optimizer_f every epoch) must reach the
        captured kernels: the fused optimizers read lr / momentum / weight decay from a device tensor that is refreshed
        here; an optimizer torch.optim executes inside the graph has its hyper-parameters baked in at capture."""
        for f in self._fused.values():
            if f:
                f.refresh_hyper()
        if self._hyper_snapshot() != self._captured_lr:
            baked = [name for name in ("g", "f") if not self._fused.get(name)]
            if baked:
                raise RuntimeError("MCDStep: hyper-parameters of optimizer_%s changed after capture() and that optimizer "
                                   "is not the fused SGD (its values are baked into the CUDA graph): call capture() "
                                   "again" % "/".join(baked))
            self._captured_lr = self._hyper_snapshot()

    def replay(self, src_imgs, src_lbls, tgt_imgs):
        _tree_copy(self._static, (src_imgs, src_lbls, tgt_imgs))
        self._before_replay()
        self.graph.replay()
        return self._static_out

    # -- input pipelining: the next batch travels host -> device while the current iteration computes ----------
    def prefetch(self, src_imgs, src_lbls, tgt_imgs):
        """Start the host->device copy of the NEXT batch (pinned host tensors) on a copy stream into staging
        buffers; `replay_prefetched()` consumes it.  The data loader's `.cuda(non_blocking=True)` of the
        reference (adapt_trainer.py:156-160) overlapped with the running iteration."""
        dev = _first(self._static).device
        if getattr(self, "_stage", None) is None:
            self._stage = _tree(torch.empty_like, self._static)
            self._copy_stream = torch.cuda.Stream(dev)
            self._copied = torch.cuda.Event()
            self._consumed = torch.cuda.Event()
            self._consumed.record(torch.cuda.current_stream(dev))
        cs = self._copy_stream
        cs.wait_event(self._consumed)            # the previous staging contents have been moved into the static inputs
        with torch.cuda.stream(cs):
            _tree_copy(self._stage, (src_imgs, src_lbls, tgt_imgs))
            self._copied.record(cs)

    def replay_prefetched(self):
        """device->device move of the prefetched batch into the graph's static inputs, then one graph launch."""
        dev = _first(self._static).device
        main = torch.cuda.current_stream(dev)
        main.wait_event(self._copied)
        _tree_copy(self._static, self._stage)
        self._consumed.record(main)
        self._before_replay()
        self.graph.replay()
        return self._static_out

    # -- one iteration ------------------------------------------------------------------------------
    def __call__(self, src_imgs, src_lbls, tgt_imgs):
        from . import ops
        dev = _first(src_imgs).device
        if self._arena is None or self._arena.buf.device != dev:
            self._arena = ops.ZeroArena(dev)
        prev_arena = ops.set_arena(self._arena)
        from .nn import DirectGrads, logits_dtype
        try:
            self._arena.begin()            # ONE memset for all BatchNorm-statistic / loss accumulators
            with DirectGrads(defer=self.defer_reduce) as self._dg, logits_dtype(self.logits_dtype):
                self._dev = dev
                if self.pipeline is not None:
                    pl = self.pipeline
                    bd = src_imgs[2] if len(src_imgs) > 2 else None
                    src_imgs, tgt_imgs = pl.images(src_imgs[:2], bd), pl.images(tgt_imgs[:2])
                    src_lbls = pl.labels(src_lbls)
                return self._iteration(src_imgs, src_lbls, tgt_imgs)
        finally:
            ops.set_arena(prev_arena)

    def _step(self, which):
        """optimizer.step() and the refresh of the packed 16-bit weight shadows of the stepped modules: ONE fused
        kernel when the optimizer is plain momentum SGD (ops.FusedSGD), else optimizer.step() + one multi-tensor
        re-pack."""
        from .nn import Conv2d
        opt, mods = (self.optimizer_g, self.gens) if which == "g" else (self.optimizer_f, self.clfs)
        if which not in self._fused:
            convs = [m for g in mods for m in g.modules() if isinstance(m, Conv2d) and m._packs]
            if self.fused_sgd and ops_mod().FusedSGD.supports(opt):
                self._fused[which] = ops_mod().FusedSGD(opt, convs,
                                                        defer_wgrad_reduce=self.defer_reduce and which == "g")
            else:
                self._fused[which] = False
        if self._fused[which]:
            self._fused[which].step()
            return
        opt.step()
        if which not in self._packers:
            convs = [m for g in mods for m in g.modules() if isinstance(m, Conv2d) and m._packs]
            self._packers[which] = ops_mod().MultiPacker(convs)
        self._packers[which].repack()

    def _backward(self, loss):
        """loss.backward() + join of the side stream that carries the convolution weight gradients."""
        loss.backward()
        self._dg.join(self._dev)

    def _mark(self, name):
        if self.phase_events is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.phase_events.append((name, e))

    def _global(self, loss):
        """the logged scalar: with world > 1 every criterion returned this rank's share, the sum is the global loss"""
        v = loss.detach().clone()
        if self.world > 1:
            torch.distributed.all_reduce(v, op=torch.distributed.ReduceOp.SUM, group=self.group)
        return v

    def _iteration(self, src_imgs, src_lbls, tgt_imgs):
        task = self.task
        from .nn import bn_update_repeat
        # ---- A: source supervised; updates G (encoder) and F1, F2 (decoder)
        self._mark("start")
        self.sync_g.zero_and_arm(), self.sync_f.zero_and_arm()
        fs = task.features(src_imgs)
        ft = None
        if task.a_uses_target:
            with torch.set_grad_enabled(not getattr(task, "a_target_forward_only", False)):
                ft = task.features(tgt_imgs)
        loss = task.loss_a(fs, ft, src_imgs, src_lbls, tgt_imgs)
        self._backward(loss)
        c_loss = self._global(loss)
        self.sync_g.wait(), self.sync_f.wait()
        self._step("g"), self._step("f")
        self._mark("A")
        # ---- B: classifiers maximise the discrepancy on target; only optimizer_f steps
        self.sync_f.zero_and_arm()
        if self.exact:
            self.sync_g.zero_and_arm(armed=False)
            fs, ft = task.features(src_imgs), task.features(tgt_imgs)
            ft_graph = None
        else:
            with torch.no_grad():
                fs = task.features(src_imgs)
            if self.reuse_t:
                with bn_update_repeat(2):
                    ft_graph = task.features(tgt_imgs)
                ft = _detach(ft_graph)
            else:
                ft_graph = None
                with torch.no_grad():
                    ft = task.features(tgt_imgs)
        loss = task.loss_b(fs, ft, src_imgs, src_lbls, tgt_imgs)
        self._backward(loss)
        self.sync_f.wait()
        self._step("f")
        self._mark("B")
        # ---- C x num_k: generator minimises the discrepancy; only optimizer_g steps
        #      (classifier gradients of this phase are zeroed before use at adapt_trainer.py:163-164, so unless
        #       exact_reference_backward is set they are not computed at all)
        self.sync_f.disarm()
        if not self.exact:
            for p in self.sync_f.params:
                p.requires_grad_(False)
        try:
            for k in range(self.num_k):
                self.sync_g.zero_and_arm()
                feats = ft_graph if (k == 0 and ft_graph is not None) else task.features(tgt_imgs)
                ft_graph = None
                loss = task.disc(feats) * self.mult
                self._backward(loss)
                self.sync_g.wait()
                self._step("g")
                self._mark("C%d" % k)
        finally:
            if not self.exact:
                for p in self.sync_f.params:
                    p.requires_grad_(True)
        d_loss = self._global(loss) / self.num_k
        return c_loss, d_loss
