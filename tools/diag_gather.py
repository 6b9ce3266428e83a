"""Where does the multi-GPU all-gather of the per-ray outputs spend its time?

    python -m torch.distributed.run --nproc-per-node N tools/diag_gather.py [--rays R]

Times, per rank and as the max over ranks: the copy-engine pushes alone, the cross-rank barrier
alone, pushes + barrier, the host-side enqueue cost of one push call, and an SM push kernel
(torch copy into the peers' buffers) for comparison.  Diagnostic only.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import avr_b200  # noqa: E402
from avr_b200 import _lib, dist as avr_dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=1 << 20)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    lib = avr_b200.load_library()
    fg = avr_dist.FusedGather(a.rays, dev)
    assert fg.available, fg.error
    fg.gathered.normal_()
    st = torch.cuda.current_stream(dev)
    res = {}

    def timed(name, fn, iters=a.iters):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(dev)
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(st)
        for _ in range(iters):
            fn()
        e1.record(st)
        host = (time.perf_counter() - t0) / iters
        torch.cuda.synchronize(dev)
        t = torch.tensor([e0.elapsed_time(e1) / iters, host * 1e3], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[name] = {"gpu_ms": round(float(t[0]), 4), "host_enqueue_ms": round(float(t[1]), 4)}

    def push():
        rc = lib.avr_gather_push_rows(fg._ptr_array, world, rank, rank * a.rays, a.rays, st.cuda_stream)
        assert rc == 0, lib.avr_last_cuda_error()

    peers = [fg.handle.get_buffer(p, (world * a.rays, 4), torch.float32) for p in range(world)]
    mine = fg.gathered[rank * a.rays:(rank + 1) * a.rays]

    def push_sm():
        for d in range(1, world):
            p = (rank + d) % world
            peers[p][rank * a.rays:(rank + 1) * a.rays].copy_(mine)

    packed = torch.randn(a.rays, 4, device=dev)
    out = torch.empty(world * a.rays, 4, device=dev)

    timed("ce_push", push)
    timed("barrier", lambda: fg.handle.barrier())
    timed("ce_push+barrier", lambda: (push(), fg.handle.barrier()))
    timed("torch_copy_push", push_sm)
    timed("nccl_all_gather", lambda: dist.all_gather_into_tensor(out, packed))
    if rank == 0:
        mb = a.rays * 16 * (world - 1) / 1e6
        print(json.dumps({"world": world, "rays_per_rank": a.rays, "MB_out_per_rank": mb, **res}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
