"""Data-parallel plumbing: one process per GPU, NCCL all-reduce of gradients over NVLink/NVSwitch.

Replaces the reference's single-process nn.DataParallel (models/model_util.py:283-284).  The MCD path shards
over independent source/target image pairs; the only exchange is the gradient all-reduce before each
optimiser step, plus a 4-float all-reduce of the cross-entropy normaliser so that the
loss keeps DataParallel's global `sum w` semantics.

GradSync keeps the gradients of one optimiser in flat fp32 buffers ("buckets", filled in reverse parameter
order = the order backward produces them).  `param.grad` are views into the buckets, so
  * zeroing is one memset per bucket,
  * a bucket is all-reduced (SUM) on a side stream as soon as autograd has accumulated its last gradient,
    overlapping the remaining dgrad / wgrad kernels,
  * no gather / scatter copies are needed.
With world_size == 1 (or no process group) nothing is exchanged: `zero_and_arm` simply drops the gradients
(`p.grad = None`), so autograd hands each freshly computed gradient to the parameter without an accumulation
kernel and nothing needs zeroing.
Works with any torch.distributed backend (tests run it on gloo with CPU tensors).
"""
import torch
import torch.distributed as dist


class _Bucket:
    __slots__ = ("flat", "params", "pending", "work", "event", "streams")

    def __init__(self, flat, params):
        self.flat, self.params = flat, params
        self.pending, self.work, self.event = 0, None, None
        self.streams = []


class GradSync:
    def __init__(self, params, process_group=None, bucket_mb=25):
        self.params = [p for p in params if p.requires_grad]
        self.group = process_group
        self.world = 1
        if dist.is_available() and dist.is_initialized():
            self.world = dist.get_world_size(process_group)
        self.armed = False
        self._got = set()
        self.buckets = []
        self._by_param = {}
        self.flat = self.world > 1
        if not self.flat:
            self.cuda, self.comm_stream = bool(self.params) and self.params[0].is_cuda, None
            return
        cap = int(bucket_mb * (1 << 20) / 4)
        cur, cur_n = [], 0
        for p in reversed(self.params):
            if cur and cur_n + p.numel() > cap:
                self._make_bucket(cur)
                cur, cur_n = [], 0
            cur.append(p)
            cur_n += p.numel()
        if cur:
            self._make_bucket(cur)
        self.cuda = bool(self.params) and self.params[0].is_cuda
        self.comm_stream = torch.cuda.Stream(self.params[0].device) if (self.cuda and self.world > 1) else None
        if self.world > 1:
            for p in self.params:
                p.register_post_accumulate_grad_hook(self._on_grad_ready)

    def _make_bucket(self, params):
        n = sum(p.numel() for p in params)
        flat = torch.zeros(n, dtype=params[0].dtype, device=params[0].device)
        b = _Bucket(flat, list(params))
        off = 0
        for p in params:
            p.grad = flat[off:off + p.numel()].view_as(p)
            self._by_param[p] = b
            p._mcd_sync = self
            off += p.numel()
        self.buckets.append(b)

    # ---------------------------------------------------------------------------------------------
    def zero_and_arm(self, armed=True):
        """zero the flat gradients (re-attaching the views if something replaced .grad) and arm the hooks."""
        for p in self.params:
            p._mcd_written = False
        if not self.flat:
            for p in self.params:
                p.grad = None
            return
        for b in self.buckets:
            b.flat.zero_()
            b.streams = []
            off = 0
            for p in b.params:
                g = p.grad
                if g is None or g.data_ptr() != b.flat.data_ptr() + off * b.flat.element_size():
                    p.grad = b.flat[off:off + p.numel()].view_as(p)
                off += p.numel()
            b.pending, b.work, b.event = len(b.params), None, None
        self._got = set()
        self.armed = armed and self.world > 1

    def disarm(self):
        self.armed = False

    def _on_grad_ready(self, p):
        if not self.armed:
            return
        b = self._by_param[p]
        self._got.add(p)
        b.pending -= 1
        if b.pending == 0:
            self._launch(b)

    def mark_ready(self, p, stream=None):
        """gradient of `p` has been written into its bucket view by a kernel enqueued on `stream` (direct-gradient
        mode of mcd_b200.nn: no autograd accumulation, hence no hook) - same bookkeeping as the hook."""
        if not self.flat:
            return
        b = self._by_param[p]
        if stream is not None and stream not in b.streams:
            b.streams.append(stream)
        self._on_grad_ready(p)

    def _launch(self, b):
        if self.comm_stream is not None:
            self.comm_stream.wait_stream(torch.cuda.current_stream(b.flat.device))
            for st in b.streams:
                self.comm_stream.wait_stream(st)
            with torch.cuda.stream(self.comm_stream):
                b.work = dist.all_reduce(b.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        else:
            b.work = dist.all_reduce(b.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def wait(self):
        """all buckets reduced and visible to the compute stream (call before optimizer.step())."""
        if not self.armed:
            return
        for b in self.buckets:
            if b.pending > 0:          # parameters that received no gradient this phase (e.g. unused decoders)
                self._launch(b)
                b.pending = 0
        for b in self.buckets:
            if b.work is not None:
                b.work.wait()
                b.work = None
        if self.comm_stream is not None:
            torch.cuda.current_stream(self.buckets[0].flat.device).wait_stream(self.comm_stream)
        # parameters no rank produced a gradient for (unused decoders: identical on every rank) keep `grad is None`
        # semantics: the optimizer skips them instead of applying momentum / weight decay to a zero gradient
        for p in self.params:
            if p not in self._got:
                p.grad = None
        self.armed = False


class DataParallel(torch.nn.Module):
    """What `is_data_parallel=True` of the reference's factories returns (models/model_util.py:75-76,96-97,283-284:
    `torch.nn.DataParallel(model)`), for one process per GPU.

    torch.nn.DataParallel replicates a module over the GPUs of ONE process every forward; here every rank of a
    `torchrun` job holds its own replica and its own shard of the batch.  The wrapper keeps what scripts and
    checkpoints see of nn.DataParallel - the `.module` attribute, the `module.` prefix of the state_dict keys, the
    call signature - and makes a hand-written training loop data parallel: with torch.distributed initialised
    (world > 1) rank 0's parameters and buffers are broadcast before the first forward (replica 0 is the one
    nn.DataParallel keeps; by then the script has moved the model to its device, which NCCL needs) and every parameter
    gradient is SUM-all-reduced when autograd delivers it.  The criteria return each rank's
    SHARE of the global-batch loss once `loss.set_process_group()` has been called, so the summed gradients are the
    global-batch gradients (DataParallel's semantics; BatchNorm statistics stay per replica, as there).
    With one process it is a transparent wrapper.  `mcd_b200.step.MCDStep` unwraps it and uses its own bucketed,
    overlapped GradSync instead of the per-parameter collectives of this slow-but-general path."""

    def __init__(self, module, process_group=None):
        super().__init__()
        self.module = module
        self.group = process_group
        self.sync_in_backward = True
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        self._replicated = self.world == 1
        if self.world > 1:
            for prm in module.parameters():
                if prm.requires_grad:
                    prm.register_hook(self._reduce)     # fires once per backward with the gradient of THAT pass

    def _reduce(self, grad):
        if not self.sync_in_backward:
            return None
        g = grad.contiguous().clone()
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group)
        return g

    def replicate(self):
        """rank 0's parameters and buffers to every rank (once; idempotent)"""
        if not self._replicated:
            with torch.no_grad():
                for t in list(self.module.parameters()) + list(self.module.buffers()):
                    dist.broadcast(t, 0, group=self.group)
            self._replicated = True

    def forward(self, *inputs, **kwargs):
        self.replicate()
        return self.module(*inputs, **kwargs)


def unwrap(module):
    """the module inside a DataParallel wrapper (whose own gradient exchange is switched off: the caller takes over)"""
    if isinstance(module, DataParallel):
        module.replicate()
        module.sync_in_backward = False
        return module.module
    return module


def init_from_env(backend=None):
    """torchrun entry: RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* from the environment."""
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if torch.cuda.is_available():
            torch.cuda.set_device(local)
        dist.init_process_group(backend or ("nccl" if torch.cuda.is_available() else "gloo"),
                                rank=rank, world_size=world)
    return rank, local, world


def allreduce_sum_(t, group=None):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t
