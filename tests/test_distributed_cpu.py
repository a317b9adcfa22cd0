"""world_size-2 `gloo` tests (CPU) of the multi-GPU host logic: batch sharding + the final gather,
the exact-mode 2-scalar reduction of the corrector, and the bucketed, overlapped gradient averaging used by
data-parallel DSM training.  The same code runs over NCCL on the GPU box (bench.py --gpus N)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from score_based_multimodal_autoencoder_b200 import distributed as D


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(fn, world=2):
    port = _free_port()
    mp.spawn(_entry, args=(world, port, fn), nprocs=world, join=True)


def _entry(rank, world, port, fn):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fn(rank, world)
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions_the_batch():
    for gb in (1, 7, 64, 1024, 1025):
        for world in (1, 2, 3, 8):
            rs = [D.shard_range(gb, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == gb
            assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            sizes = [hi - lo for lo, hi in rs]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


def _gather_and_reduce(rank, world):
    # even and ragged global batches: every rank contributes its slice of arange, the gather restores the order
    for gb in (8, 7):
        full = torch.arange(gb * 6, dtype=torch.float32).view(gb, 1, 2, 3)
        lo, hi = D.shard_range(gb)
        out = D.gather_batch(full[lo:hi].clone(), gb)
        assert torch.equal(out, full)
    # exact-mode corrector reduction: per-shard sums of per-sample norms add up to the full-batch sums
    g = torch.Generator().manual_seed(0)
    grad = torch.randn(10, 5, 8, 8, generator=g)
    noise = torch.randn(10, 5, 8, 8, generator=g)
    lo, hi = D.shard_range(10)
    acc = torch.tensor([grad[lo:hi].flatten(1).norm(dim=1).sum(), noise[lo:hi].flatten(1).norm(dim=1).sum()],
                       dtype=torch.float64)
    D.corrector_allreduce()(acc)
    ref = torch.tensor([grad.flatten(1).norm(dim=1).sum(), noise.flatten(1).norm(dim=1).sum()], dtype=torch.float64)
    assert torch.allclose(acc, ref, rtol=1e-6)


def test_gather_batch_and_corrector_reduction_gloo():
    _run(_gather_and_reduce)


def _grad_reducer(rank, world):
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(s)) for s in [(300, 7), (5,), (4000,), (64, 3, 3, 3), (11,)]]
    red = D.GradReducer(params, bucket_bytes=4096)  # several buckets
    assert len(red.buckets) >= 3 and red.buckets[-1][1] == sum(p.numel() for p in params)
    for it in range(2):  # the flat buffer is reused across steps
        red.begin()
        views = {}
        # backward order = reverse registration order; rank-dependent gradients
        for i in reversed(range(len(params))):
            p = params[i]
            g = torch.full(p.shape, float(i + 1 + it)) * (rank + 1)
            views[i] = red.grad_ready(p, g)
        red.finish()
        assert red.launched == list(range(len(red.buckets)))  # buckets completed front to back
        mean_scale = sum(r + 1 for r in range(world)) / world
        for i, v in views.items():
            assert torch.allclose(v, torch.full(params[i].shape, float(i + 1 + it) * mean_scale))
            assert v.data_ptr() == red.grad_view(params[i]).data_ptr()
    # a parameter that produced no gradient contributes zeros and does not dead-lock its bucket
    red.begin()
    for i in (4, 3, 1, 0):
        red.grad_ready(params[i], torch.ones(params[i].shape))
    red.finish()
    assert torch.equal(red.grad_view(params[2]), torch.zeros(4000))
    assert torch.allclose(red.grad_view(params[0]), torch.ones(300, 7))
    with pytest.raises(RuntimeError):
        red.begin()
        red.grad_ready(params[0], torch.ones(300, 7))
        red.grad_ready(params[0], torch.ones(300, 7))


def test_grad_reducer_bucketed_average_gloo():
    _run(_grad_reducer)


def _grad_reducer_bf16(rank, world):
    params = [torch.nn.Parameter(torch.zeros(s)) for s in [(300, 7), (5,), (4000,)]]
    red = D.GradReducer(params, bucket_bytes=4096, comm_dtype=torch.bfloat16)
    red.begin()
    g = torch.Generator().manual_seed(5)
    full = {i: torch.randn(params[i].shape, generator=g) for i in range(3)}
    for i in reversed(range(3)):
        red.grad_ready(params[i], full[i] * (rank + 1))
    red.finish()
    scale = sum(r + 1 for r in range(world)) / world
    for i in range(3):
        v = red.grad_view(params[i])
        assert v.dtype == torch.float32
        assert torch.allclose(v, full[i] * scale, rtol=2e-2, atol=1e-3)      # bf16 on the wire
        assert not torch.equal(v, full[i] * scale)


def test_grad_reducer_bf16_communication_gloo():
    _run(_grad_reducer_bf16)


def _ddp_wrapper(rank, world):
    torch.manual_seed(rank)  # different initial weights per rank: the wrapper must broadcast rank 0's
    net = torch.nn.Linear(6, 3)
    ddp = D.DataParallelScoreNet(net, bucket_mb=1e-4)
    ws = [torch.empty_like(net.weight) for _ in range(world)]
    dist.all_gather(ws, net.weight.detach())
    assert all(torch.equal(ws[0], w) for w in ws)
    assert set(ddp.state_dict().keys()) == {"weight", "bias"}  # reference key names, no "module." prefix
    assert net._grad_sink is ddp.reducer


def test_data_parallel_wrapper_broadcasts_and_keeps_schema_gloo():
    _run(_ddp_wrapper)


class _SinkFn(torch.autograd.Function):
    """Stand-in for the score net's autograd node: reports the gradient through the reducer like autograd._Plan does."""

    @staticmethod
    def forward(ctx, red, p, scale):
        ctx.red, ctx.p, ctx.scale = red, p, scale
        return (p * scale).sum()

    @staticmethod
    def backward(ctx, dout):
        ctx.red.begin()
        g = ctx.red.grad_ready(ctx.p, torch.full(ctx.p.shape, ctx.scale) * dout)
        ctx.red.finish()
        return None, g, None


def test_grad_reducer_views_survive_zero_grad_in_place_and_accumulation():
    """p.grad aliases the flat bucket buffer after the first backward; a second backward with the gradient
    still alive (set_to_none=False, or accumulation) must not double it."""
    p = torch.nn.Parameter(torch.zeros(5))
    red = D.GradReducer([p])
    opt = torch.optim.SGD([p], lr=0.0)
    seen = []
    for _ in range(3):
        opt.zero_grad(set_to_none=False)
        _SinkFn.apply(red, p, 1.0).backward()
        seen.append(p.grad.clone())
    assert all(torch.equal(g, torch.ones(5)) for g in seen), seen
    # gradient accumulation: two backward passes per step add up
    opt.zero_grad(set_to_none=True)
    _SinkFn.apply(red, p, 2.0).backward()
    _SinkFn.apply(red, p, 3.0).backward()
    assert torch.equal(p.grad, torch.full((5,), 5.0))
    # the default mode still hands out the zero-copy view
    opt.zero_grad(set_to_none=True)
    _SinkFn.apply(red, p, 4.0).backward()
    assert torch.equal(p.grad, torch.full((5,), 4.0))
    assert p.grad.untyped_storage().data_ptr() == red.flat.untyped_storage().data_ptr()
