"""Host logic of the multi-GPU path, exercised with world_size 2 over gloo on CPU tensors
(the compositing itself is stubbed with the oracle here: the CUDA kernels need a GPU, the
sharding / packing / gather logic does not)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, r_total, k, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "oracle")]
    import avr_oracle as O
    from avr_b200 import dist as avr_dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        z = torch.sort(0.8 + torch.rand(r_total, k, generator=g), -1).values
        x = torch.cat([torch.sigmoid(torch.randn(r_total, k, 3, generator=g)), torch.relu(torch.randn(r_total, k, 1, generator=g)) * 30], -1)
        lo, hi = avr_dist.shard_bounds(r_total, world, rank)

        def oracle_composite(rgbs, zz, wb, inf, want_w):
            rgb, depth, w = O.composite_rgbs(zz.unsqueeze(0), rgbs.unsqueeze(0), wb, inf)
            return rgb[0], depth[0, :, 0], None

        xs = x[lo:hi].clone().requires_grad_(True)
        rgb, depth, rgb_all, depth_all = avr_dist.composite_sharded(xs, z[lo:hi], True, 1.8, composite_fn=oracle_composite)
        full = O.composite_rgbs(z.unsqueeze(0), x.unsqueeze(0), True)
        ok = torch.equal(rgb_all, full[0][0]) and torch.equal(depth_all, full[1][0, :, 0])
        ok = ok and rgb.shape[0] == hi - lo and not rgb_all.requires_grad
        rgb.sum().backward()                                   # gradients stay local
        ok = ok and xs.grad is not None and xs.grad.shape == xs.shape
        q.put((rank, ok, lo, hi))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("r_total", [64, 65])      # equal and unequal shards
def test_sharded_composite_gathers_full_image(r_total):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, r_total, 12, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _ in res), res
    assert res[0][2] == 0 and res[0][3] == res[1][2] and res[1][3] == r_total


def test_shard_bounds_cover_everything():
    from avr_b200.dist import shard_bounds, shard_bounds_packed

    for n, w in ((1 << 24, 8), (10, 4), (3, 8), (0, 2)):
        spans = [shard_bounds(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
    g = torch.Generator().manual_seed(0)
    counts = torch.randint(8, 257, (5000,), generator=g)
    offsets = torch.zeros(5001, dtype=torch.int64)
    offsets[1:] = torch.cumsum(counts, 0)
    spans = [shard_bounds_packed(offsets, 8, r) for r in range(8)]
    assert spans[0][0] == 0 and spans[-1][1] == 5000 and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    loads = [int(offsets[b] - offsets[a]) for a, b in spans]
    assert max(loads) - min(loads) <= 2 * 256        # balanced by samples to within a ray or two
