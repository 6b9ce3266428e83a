"""Multi-GPU path on real devices (needs >= 2 GPUs; skipped on a single-GPU box): ray sharding,
NCCL all-gather of the outputs, and the all-gather fused into the forward kernel's epilogue
(peer stores into symmetric memory)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q, multicast):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "oracle")]
    import torch.distributed as dist
    from avr_b200 import _lib, dist as avr_dist, ops

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), AVR_GATHER_MULTICAST="1" if multicast else "0")
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        r_total, k = 6000, 96
        g = torch.Generator().manual_seed(0)
        z = torch.sort(0.8 + torch.rand(r_total, k, generator=g), -1).values
        x = torch.cat([torch.sigmoid(torch.randn(r_total, k, 3, generator=g)), torch.relu(torch.randn(r_total, k, 1, generator=g)) * 30], -1)
        lo, hi = avr_dist.shard_bounds(r_total, world, rank)
        xs, zs = x[lo:hi].to(dev), z[lo:hi].to(dev)
        # reference result: every rank composites everything locally
        full_rgb, full_depth, _ = ops.composite(x.to(dev), z.to(dev), True, 1.8, want_w=False)
        # (1) NCCL all-gather of the shard outputs (also async, overlapping other work)
        rgb, depth, rgb_all, depth_all = avr_dist.composite_sharded(xs, zs, True, 1.8)
        ok_nccl = bool(torch.allclose(rgb_all, full_rgb, rtol=2e-6, atol=2e-7) and torch.allclose(depth_all, full_depth, rtol=2e-6, atol=2e-7))
        h = avr_dist.all_gather_outputs(rgb, depth, async_op=True)
        a, b = h.wait()
        ok_nccl = ok_nccl and torch.equal(a, rgb_all) and torch.equal(b, depth_all)
        # (2) gather fused into the kernel epilogue
        fg = avr_dist.FusedGather(hi - lo, dev)
        ok_fused, note = None, fg.error
        if fg.available:
            out = fg.composite_fwd(xs, zs, True, 1.8, want_w=True)
            assert out is not None
            rgb_f, depth_f, w_f = out
            g_rgb, g_depth = fg.finish()
            torch.cuda.synchronize()
            ok_fused = bool(torch.equal(rgb_f, rgb) and torch.equal(depth_f, depth)
                            and torch.equal(g_rgb, rgb_all) and torch.equal(g_depth, depth_all))
            # (2b) six consecutive steps through the two slots with NO cross-rank barrier, uneven shards
            # (6001 rays), a different image every step, the ranks deliberately out of step (odd ranks do
            # extra work before launching, even ranks before reading), against the CPU oracle
            import avr_oracle as O
            r2 = 6001
            z2 = torch.cat([z, z[:1]], 0)
            x2 = torch.cat([x, x[:1]], 0)
            lo2, hi2 = avr_dist.shard_bounds(r2, world, rank)
            fg2 = avr_dist.FusedGather(hi2 - lo2, dev)
            assert fg2.available, fg2.error
            zs2 = z2[lo2:hi2].to(dev)
            for step in range(6):
                xi = x2.clone()
                xi[..., :3] *= 1.0 / (step + 1)
                if rank % 2 == 1:
                    busy = ops.composite(x.to(dev), z.to(dev), True, 1.8, want_w=False)
                out = fg2.composite_fwd(xi[lo2:hi2].to(dev), zs2, True, 1.8, want_w=False)
                assert out is not None
                if rank % 2 == 0:
                    busy = ops.composite(x.to(dev), z.to(dev), True, 1.8, want_w=False)
                a_rgb, a_depth = fg2.finish()
                got_rgb, got_depth = a_rgb.cpu(), a_depth.cpu()      # read before the next launch (the contract)
                want = O.composite_rgbs(z2.unsqueeze(0), xi.unsqueeze(0), True)
                ok_fused = ok_fused and bool(torch.allclose(got_rgb, want[0][0], rtol=1e-5, atol=1e-6)
                                             and torch.allclose(got_depth, want[1][0, :, 0], rtol=1e-5, atol=1e-6)
                                             and torch.equal(got_rgb[lo2:hi2], out[0].cpu()))
            ok_fused = ok_fused and not fg2.timed_out()
            # (3) copy-engine all-gather, double-buffered: three steps through two slots
            pg = avr_dist.PipelinedGather(hi - lo, dev)
            assert pg.available, pg.error
            lib = _lib.load()
            r_loc = hi - lo
            for step in range(3):
                scale = float(step + 1)
                xs2 = xs.clone()
                xs2[..., :3] *= 1.0 / scale                      # a different image every step
                target, n_t, row0 = pg.forward_target()
                o_rgb = torch.empty(r_loc, 3, device=dev)
                o_depth = torch.empty(r_loc, device=dev)
                rc = lib.avr_composite_fwd_gather(xs2.data_ptr(), zs.data_ptr(), r_loc, k, 1, 1.8, None, o_rgb.data_ptr(),
                                                  o_depth.data_ptr(), target, n_t, row0,
                                                  torch.cuda.current_stream(dev).cuda_stream)
                assert rc == 0
                slot = pg.push_async()
                busy = ops.composite(x.to(dev), z.to(dev), True, 1.8, want_w=False)   # work the pushes overlap with
                c_rgb, c_depth = pg.wait(slot)
                x2 = x.clone()
                x2[..., :3] *= 1.0 / scale
                want_rgb, want_depth, _ = ops.composite(x2.to(dev), z.to(dev), True, 1.8, want_w=False)
                torch.cuda.synchronize()
                ok_fused = ok_fused and bool(torch.allclose(c_rgb, want_rgb, rtol=2e-6, atol=2e-7)
                                             and torch.allclose(c_depth, want_depth, rtol=2e-6, atol=2e-7)
                                             and torch.equal(c_rgb[lo:hi], o_rgb) and torch.equal(busy[0], full_rgb))
                dist.barrier()                                   # consumers are done before the slot is reused
        q.put((rank, ok_nccl, ok_fused, note))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("multicast", [False, True])     # per-peer stores / NVSwitch multicast mapping
def test_two_gpu_sharded_composite_and_fused_gather(multicast):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, multicast)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, ok_nccl, ok_fused, note in res:
        assert ok_nccl, f"rank {rank}: NCCL gather mismatch"
        assert ok_fused is not False, f"rank {rank}: fused gather mismatch"
        if ok_fused is None:
            pytest.skip(f"symmetric memory unavailable on this box: {note}")
