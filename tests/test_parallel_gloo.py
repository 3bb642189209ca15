"""Host-side logic of the data-parallel path on CPU: world_size-2 gloo process groups exercise GradSync
(bucketing in reverse parameter order, hook-driven asynchronous all-reduce, parameters that receive no gradient,
flat zeroing, re-arming across phases) exactly as the NCCL path uses it on the GPUs."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "multichannel-semseg-with-uda_b200")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, bucket_mb, q):
    try:
        for p in (PKG, ROOT):
            if p not in sys.path:
                sys.path.insert(0, p)
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                          WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
        from mcd_b200 import parallel
        r, l, w = parallel.init_from_env(backend="gloo")
        assert (r, w) == (rank, world)
        torch.manual_seed(0)                      # identical replicas
        net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4),
                                  torch.nn.Linear(4, 4))   # last layer unused below -> no gradient
        params = list(net.parameters())
        sync = parallel.GradSync(params, None, bucket_mb)
        assert sync.world == world
        ref = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4))
        ref.load_state_dict({k: v.clone() for k, v in net.state_dict().items() if not k.startswith("3.")})
        gen = torch.Generator().manual_seed(1)
        xs = [torch.randn(5, 8, generator=gen) for _ in range(world)]   # every rank knows every shard
        for phase in range(3):                    # re-arm / re-zero across phases
            sync.zero_and_arm()
            out = net[2](net[1](net[0](xs[rank])))
            out.pow(2).sum().backward()
            sync.wait()
            ref.zero_grad()
            sum(ref(x).pow(2).sum() for x in xs).backward()      # SUM over shards == SUM all-reduce
            for p, q_ in zip(params[:4], ref.parameters()):
                assert torch.allclose(p.grad, q_.grad, rtol=1e-5, atol=1e-6), (phase, rank)
            # an unused parameter does not dead-lock the bucket and keeps `grad is None` (the optimizer skips it, as
            # torch.optim does after zero_grad(set_to_none=True)); the other gradients are views of the flat buckets
            assert params[4].grad is None and params[5].grad is None
            assert all(p.grad.data_ptr() >= sync._by_param[p].flat.data_ptr() for p in params[:4])
        # disarmed phase: hooks must not communicate
        sync.zero_and_arm(armed=False)
        net[0](xs[rank]).sum().backward()
        sync.wait()
        assert float(params[0].grad.abs().max()) > 0
        nb = len(sync.buckets)
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok", nb))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "fail: %s\n%s" % (e, traceback.format_exc()), 0))


@pytest.mark.parametrize("bucket_mb", [25, 0.0002])
def test_gradsync_world2_gloo(bucket_mb):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, bucket_mb, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
    assert all(r[1] == "ok" for r in res), res
    if bucket_mb < 1:
        assert res[0][2] > 1      # tiny cap -> several buckets


def test_gradsync_single_process_drops_grads():
    for p in (PKG, ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    from mcd_b200 import parallel
    lin = torch.nn.Linear(4, 4)
    sync = parallel.GradSync(list(lin.parameters()))
    assert sync.world == 1
    sync.zero_and_arm()
    lin(torch.ones(2, 4)).sum().backward()
    sync.wait()
    g = lin.weight.grad.clone()
    sync.zero_and_arm()
    assert lin.weight.grad is None            # single process: gradients are dropped, not zero-filled
    lin(torch.ones(2, 4)).sum().backward()
    assert torch.equal(lin.weight.grad, g)


def _dp_worker(rank, world, port, q):
    try:
        for p in (PKG, ROOT):
            if p not in sys.path:
                sys.path.insert(0, p)
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                          WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
        from mcd_b200 import parallel
        parallel.init_from_env(backend="gloo")
        torch.manual_seed(rank)                   # DIFFERENT initial replicas: rank 0's must win (nn.DataParallel)
        inner = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.BatchNorm1d(16), torch.nn.ReLU(),
                                    torch.nn.Linear(16, 4))
        net = parallel.DataParallel(inner)
        assert net.module is inner and net.world == world
        assert all(k.startswith("module.") for k in net.state_dict())
        torch.manual_seed(0)
        ref = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.BatchNorm1d(16), torch.nn.ReLU(),
                                  torch.nn.Linear(16, 4))
        gen = torch.Generator().manual_seed(1)
        xs = [torch.randn(6, 8, generator=gen) for _ in range(world)]
        inner.eval(), ref.eval()                  # per-replica BatchNorm statistics are DataParallel's too; keep it simple
        with torch.no_grad():
            net(xs[rank])                         # the first forward replicates rank 0's parameters and buffers
        for a, b in zip(inner.state_dict().values(), ref.state_dict().values()):
            assert torch.equal(a, b), "rank 0's replica was not broadcast"
        # every rank's loss is its SHARE of the global objective, as the library's criteria return it
        for accumulate in (False, True):          # second pass WITHOUT zero_grad: gradients accumulate correctly
            if not accumulate:
                net.zero_grad(), ref.zero_grad()
            net(xs[rank]).pow(2).sum().backward()
            sum(ref(x).pow(2).sum() for x in xs).backward()
            for p_, q_ in zip(inner.parameters(), ref.parameters()):
                assert torch.allclose(p_.grad, q_.grad, rtol=1e-5, atol=1e-6), (accumulate, rank)
        # MCDStep takes the exchange over: unwrap() switches the wrapper's collectives off
        assert parallel.unwrap(net) is inner and net.sync_in_backward is False
        net.zero_grad()
        net(xs[rank]).pow(2).sum().backward()
        ref.zero_grad()
        ref(xs[rank]).pow(2).sum().backward()
        for p_, q_ in zip(inner.parameters(), ref.parameters()):
            assert torch.allclose(p_.grad, q_.grad, rtol=1e-5, atol=1e-6)
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "fail: %s\n%s" % (e, traceback.format_exc())))


def test_data_parallel_wrapper_world2_gloo():
    """models.model_util factories with is_data_parallel=True return mcd_b200.parallel.DataParallel wrappers
    (reference models/model_util.py:283-284): replica broadcast, summed gradients == global-batch gradients,
    accumulation across backward passes, `module.` state_dict keys."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
    assert all(r[1] == "ok" for r in res), res


def test_factories_wrap_with_is_data_parallel():
    for p in (PKG, ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import warnings
    from mcd_b200 import parallel
    from models.model_util import get_models, get_multitask_models
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ms = get_models("drn_d_22", 6, 41, is_data_parallel=True)
        enc, dec = get_multitask_models("drn_d_22", 6, 41, is_data_parallel=True)
    assert isinstance(ms, list) and all(isinstance(m, parallel.DataParallel) for m in ms + [enc, dec])
    assert all(k.startswith("module.") for k in ms[0].state_dict()) and hasattr(ms[0].module, "seg")
    assert not hasattr(dec, "get_loss")           # nn.DataParallel does not delegate attributes either
