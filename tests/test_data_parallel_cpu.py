"""CPU (gloo, world_size 2): host-side logic of the data-parallel gradient reduction — flat layout
in backward-completion order, stage slices, averaging — without any GPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import videopose3d_b200 as vp
from videopose3d_b200.data_parallel import GradientReducer, shard_batch, stage_order


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        m = vp.TemporalModelOptimized1f(17, 2, 17, [3, 3, 3], channels=64)
        red = GradientReducer(overlap=True)
        names, spans, stage_spans, total = red.plan_layout(m)
        flat = torch.zeros(total)
        params = dict(m.named_parameters())
        g = torch.Generator().manual_seed(100 + rank)
        local = {}
        for n in names:
            off, numel = spans[n]
            local[n] = torch.randn(numel, generator=g)
            flat[off:off + numel] = local[n]
        # stages arrive in backward order, exactly as the C callback reports them
        for lo, hi in stage_spans:
            red.stage_ready(flat, lo, hi)
        red.finish(flat)
        # expected: mean over ranks of the same generator streams
        for n in names:
            exp = torch.zeros_like(local[n])
            for r in range(world):
                gr = torch.Generator().manual_seed(100 + r)
                vals = {k: torch.randn(spans[k][1], generator=gr) for k in names}
                exp += vals[n]
            exp /= world
            off, numel = spans[n]
            assert torch.allclose(flat[off:off + numel], exp, atol=1e-6), n
            assert flat[off:off + numel].view(params[n].shape).shape == params[n].shape
        assert red.launched == len(stage_spans)
        # ragged last batch: rank r holds r + 1 of the 3 rows -> weights 1/3 and 2/3, not 1/2 each
        red2 = GradientReducer(overlap=False)
        red2.set_step_rows(rank + 1, 3)
        v = torch.full((8,), float(rank + 1))
        red2.reduce_flat(v)
        assert torch.allclose(v, torch.full((8,), (1 * 1 + 2 * 2) / 3.0), atol=1e-6)
        # bf16 wire format: same mean up to bf16 rounding of each contribution
        red3 = GradientReducer(overlap=False, compress="bf16")
        gr = torch.Generator().manual_seed(7 + rank)
        w = torch.randn(1000, generator=gr)
        exp = sum(torch.randn(1000, generator=torch.Generator().manual_seed(7 + r)) for r in range(world)) / world
        red3.reduce_flat(w)
        assert w.dtype == torch.float32 and float((w - exp).abs().max()) <= 2 ** -7
        out[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_gradient_reducer_world2_gloo():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        assert dict(out) == {0: "ok", 1: "ok"}


def test_layout_covers_every_parameter_once():
    m = vp.TemporalModelOptimized1f(17, 2, 17, [3, 3, 3, 3, 3], channels=128)
    red = GradientReducer()
    names, spans, stage_spans, total = red.plan_layout(m)
    assert sorted(names) == sorted(n for n, _ in m.named_parameters())
    assert [n for grp in stage_order(m) for n in grp] == names
    # stages tile the flat buffer, slices are 16-byte aligned
    assert stage_spans[0][0] == 0 and stage_spans[-1][1] == total
    for (lo, hi), (lo2, _) in zip(stage_spans, stage_spans[1:]):
        assert hi == lo2
    assert all(off % 4 == 0 for off, _ in spans.values())
    assert m._learnable_names() == [n for n in m._learnable_names()] and \
        set(m._learnable_names()) == set(names)


def test_shard_batch_partitions_batches():
    world = 8
    for b in range(32):
        owners = [r for r in range(world) if shard_batch(b, r, world)]
        assert owners == [b % world]
