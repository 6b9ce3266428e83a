#!/usr/bin/env python
"""Headline benchmark: fused alpha-compositing forward+backward on synthetic rays.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c2|c5]

A "step" is one forward+backward compositing pass over one batch of rays (BASELINE.json
configs[1]: 2^20 rays x 96 samples fp32 per GPU).  Prints ONE JSON line (rank 0).

* `value`      rays/s over all ranks, inputs already resident in HBM, timed with CUDA events
               on the launching stream, max over ranks.
* `e2e`        the same metric through the host-buffer C-ABI call
               (avr_composite_fwd_bwd_host): pinned host inputs, H2D and D2H inside the timer.
* `roofline`   the dominant kernel (composite_bwd): algorithmic bytes / measured duration
               against MEASURED_PEAKS.json's HBM copy bandwidth.
* `cpu_baseline`  the reference's torch-CPU op sequence (oracle port) timed on this box's host
               cores on a bounded sample of the same workload — reported, not a target.

`--impl reference` times only that CPU path (rank 0; other ranks exit 0).
N > 1: launched by torch.distributed.run, one rank per GPU; rays are sharded (weak scaling:
every rank composites its own 2^20-ray shard) and the only collective is the NCCL all-gather
of the per-ray outputs rgb+depth (16 B/ray), inside the timed step.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

METRIC = "composite_fwd_bwd_rays_per_sec"
UNIT = "rays/s"

# (rays, K) -> DRAM bytes per launch of composite_bwd_span_kernel (ncu, round 1)
NCU_TRAFFIC_BWD = {(1 << 20, 96): 2.030087e9 + 1.566062e9}

WORKLOADS = {
    # name: (rays per GPU, samples per ray, description)
    "c2": (1 << 20, 96, "BASELINE.json configs[1]: composite fwd+bwd, 2^20 rays x 96 samples fp32 per GPU"),
    "c5": (1 << 21, 192, "BASELINE.json configs[4] per-GPU shard at 8 GPUs: 2^21 rays x 192 samples fp32"),
}


def bytes_per_ray(k: int):
    """Algorithmic bytes: fwd reads rgbs 16K + z 4K, writes w 4K + rgb 12 +
    depth 4; bwd reads rgbs 16K + z 4K + g_rgb 12 + g_depth 4, writes d_rgbs 16K."""
    return 24 * k + 16, 36 * k + 16


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


# ------------------------------------------------------------------ clocks sampler
class ClockSampler:
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    REASONS = {
        0x0000000000000004: "sw_power_cap",
        0x0000000000000008: "hw_slowdown",
        0x0000000000000020: "sw_thermal_slowdown",
        0x0000000000000040: "hw_thermal_slowdown",
        0x0000000000000080: "hw_power_brake_slowdown",
    }

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            uuid = torch.cuda.get_device_properties(index).uuid
            try:
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if bits & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.004)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml_unavailable"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------ CPU baseline
def cpu_reference_pass(z, x, g_rgb, g_d):
    """One fwd+bwd of the reference's op sequence (oracle/avr_oracle.py, torch CPU)."""
    import avr_oracle as O

    xx = x.detach().requires_grad_(True)
    rgb, depth, _w = O.composite_rgbs(z, xx, True)
    torch.autograd.backward([rgb, depth], [g_rgb, g_d])
    return xx.grad


def cpu_inputs(rays, k, seed=0):
    g = torch.Generator().manual_seed(seed)
    z = torch.sort(0.8 + torch.rand(1, rays, k, generator=g), -1).values
    x = torch.cat([torch.sigmoid(torch.randn(1, rays, k, 3, generator=g)),
                   torch.relu(torch.randn(1, rays, k, 1, generator=g)) * 30], -1)
    return z, x, torch.randn(1, rays, 3, generator=g), torch.randn(1, rays, 1, generator=g)


def time_cpu(rays, k, steps, warmup):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    args = cpu_inputs(rays, k)
    for _ in range(warmup):
        cpu_reference_pass(*args)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        cpu_reference_pass(*args)
        times.append(time.perf_counter() - t0)
    return times, cores


def run_reference_arm(args, k, desc, emit):
    """--impl reference: the reference's own CPU implementation of the path (its torch op
    sequence, restated in oracle/ because a Python reference cannot travel to the GPU box)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_rays = 1 << 18
    times, cores = time_cpu(sample_rays, k, args.steps, max(args.warmup, 1))
    total = sum(times)
    value = sample_rays * len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "samples_per_ray": k, "sample": f"{sample_rays} rays x {k} samples per step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{len(times)} steps of {sample_rays} rays x {k} samples, torch {torch.__version__} CPU, "
                                   f"{cores} threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------ other configs (short)
def measure_other_configs(dev, peak):
    """BASELINE.json configs 3-5 on one GPU, a few iterations each (CUDA events; inputs larger than
    L2).  ms per launch, rays/s and fraction of the measured HBM roofline from the algorithmic
    bytes of SURVEY.md section 8(d)."""
    from avr_b200 import ops

    g = torch.Generator(device=dev).manual_seed(1)

    def timeit(fn, iters=10, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / iters

    def entry(ms, rays, bytes_per_ray):
        gbs = bytes_per_ray * rays / (ms * 1e-3) / 1e9
        return {"ms": round(ms, 4), "rays_per_s": rays / (ms * 1e-3), "GBps": round(gbs, 1), "hbm_frac": round(gbs / peak, 4)}

    out = {}
    r = 1 << 20
    near, far = torch.tensor([0.8], device=dev), torch.tensor([1.8], device=dev)
    # config 3: importance sampling 64 coarse weights -> 128 fine samples (+ merge with the coarse depths)
    w = torch.rand(1, r, 64, device=dev, generator=g) ** 6
    u, u2 = torch.rand(1, r, 128, device=dev, generator=g), torch.rand(1, r, 128, device=dev, generator=g)
    zc = ops.coarse_sample_raw(near, far, 0, torch.rand(1, r, 64, device=dev, generator=g))
    out["c3_importance_64_to_128_with_merge"] = entry(
        timeit(lambda: ops.importance_sample(w, near, far, u, u2, z_coarse=zc, want_fine=False, want_sorted=True)), r, 2312)
    out["c3_importance_64_to_128_sampling_only"] = entry(
        timeit(lambda: ops.importance_sample(w, near, far, u, u2, want_fine=True)), r, 1800)
    del w, u, u2, zc
    # config 4: ragged rays, counts 8..256 (2^20-ray slice of the 2^22-ray config): composite fwd + bwd, packed
    counts = torch.randint(8, 257, (r,), device=dev, generator=g)
    offsets = torch.zeros(r + 1, dtype=torch.int64, device=dev)
    offsets[1:] = torch.cumsum(counts, 0)
    s = int(offsets[-1])
    zp = torch.sort(0.8 + torch.rand(s, device=dev, generator=g)).values   # ascending everywhere => ascending per ray
    xp = torch.cat([torch.sigmoid(torch.randn(s, 3, device=dev, generator=g)),
                    torch.relu(torch.randn(s, 1, device=dev, generator=g)) * 30], -1)
    xp.requires_grad_(True)

    def packed_step():
        xp.grad = None
        rgb, depth, _ = ops.composite_packed(xp, zp, offsets, True, 1.8, want_w=False)
        torch.autograd.backward([rgb, depth], [rgb, depth])

    out["c4_packed_composite_fwd_bwd_8_to_256"] = dict(entry(timeit(packed_step, iters=5), r, (56 * s) // r + 48), samples=s)
    del xp, zp, offsets, counts
    return out


# ------------------------------------------------------------------ GPU arm
def main():
    # stdout carries exactly ONE JSON line: libraries (NCCL prints its version banner to
    # stdout) are pointed at stderr until the line is written
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="avr_b200", choices=["avr_b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--rays", type=int, default=0, help="override rays per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the short measurements of the other BASELINE.json configs")
    ap.add_argument("--gather", default="fused", choices=["ce", "fused", "nccl", "none"],
                    help="N>1, the all-gather of rgb+depth: 'fused' = the forward kernel forwards every 32 finished rays "
                         "to each peer's symmetric-memory buffer as one coalesced 512-byte store over NVLink, then a "
                         "cross-rank barrier; 'ce' = the kernel packs local rows and the copy engines push them on a side "
                         "stream (double-buffered; starved by the HBM-saturating kernels, kept for comparison); "
                         "'nccl' = all_gather_into_tensor between forward and backward (both fall back to nccl); 'none' = no exchange "
                         "(diagnostic: the compute-only step under the same launch)")
    args = ap.parse_args()
    rays, k, desc = WORKLOADS[args.workload]
    if args.rays:
        rays = args.rays
    if args.impl == "reference":
        run_reference_arm(args, k, desc, emit)
        return
    args.warmup = max(args.warmup, 3)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback "
                         "(use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)

    import __graft_entry__
    if rank == 0:
        __graft_entry__.build()
    if dist is not None:
        dist.barrier()
    import avr_b200
    from avr_b200 import dist as avr_dist

    lib = avr_b200.load_library()
    assert lib.avr_device_check() == 0, lib.avr_last_cuda_error()

    # ---- synthetic inputs, generated on the device (seed = rank), SURVEY.md 8(d) C2
    g = torch.Generator(device=dev).manual_seed(rank)
    z = torch.sort(0.8 + torch.rand(rays, k, device=dev, generator=g), -1).values
    x = torch.cat([torch.sigmoid(torch.randn(rays, k, 3, device=dev, generator=g)),
                   torch.relu(torch.randn(rays, k, 1, device=dev, generator=g)) * 30], -1).contiguous()
    g_rgb = torch.randn(rays, 3, device=dev, generator=g)
    g_d = torch.randn(rays, device=dev, generator=g)
    w = torch.empty(rays, k, device=dev)
    rgb = torch.empty(rays, 3, device=dev)
    depth = torch.empty(rays, device=dev)
    dx = torch.empty_like(x)
    stream = torch.cuda.current_stream(dev)
    sp = stream.cuda_stream

    L, rpt, main_rays = ctypes.c_int(), ctypes.c_int(), ctypes.c_int64()
    span = lib.avr_composite_plan_info(rays, k, x.data_ptr(), z.data_ptr(), ctypes.byref(L), ctypes.byref(rpt),
                                       ctypes.byref(main_rays))
    launches_per_step = 2 * (1 + (1 if (span and main_rays.value < rays) else 0)) if span else 2

    fused = None
    if dist is not None and args.gather in ("fused", "ce"):
        fused = avr_dist.PipelinedGather(rays, dev) if args.gather == "ce" else avr_dist.FusedGather(rays, dev)
        ok = torch.tensor([1 if fused.available else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if not int(ok.item()):
            if rank == 0:
                print(f"[bench] fused gather unavailable ({fused.error}); using NCCL", file=sys.stderr)
            fused = None
    ce = fused is not None and args.gather == "ce"
    gather_mode = "none" if (dist is None or args.gather == "none") else (
        "NCCL all_gather_into_tensor between forward and backward" if fused is None else
        "forward kernel packs local rows into a double-buffered symmetric-memory slot; copy-engine pushes over NVLink + "
        "cross-rank barrier on a side stream, overlapped with backward and the next step (all waited inside the timed region)"
        if ce else ("fused into the forward kernel (512-byte multimem.st through the NVSwitch multicast mapping of the "
                    "symmetric-memory buffers) + barrier" if getattr(fused, "multicast_ptr", 0) else
                    "fused into the forward kernel (coalesced 512-byte peer stores over NVLink into symmetric memory) + barrier"))

    def fwd_fused():
        if ce:
            target, n_t, row0 = fused.forward_target()
            rc = lib.avr_composite_fwd_gather(x.data_ptr(), z.data_ptr(), rays, k, 1, 1.8, w.data_ptr(), rgb.data_ptr(),
                                              depth.data_ptr(), target, n_t, row0, sp)
        else:
            rc = fused.launch(lib, x.data_ptr(), z.data_ptr(), rays, k, 1, 1.8, w.data_ptr(), rgb.data_ptr(),
                              depth.data_ptr(), sp)
        assert rc == 0, (rc, lib.avr_last_cuda_error())
        if ce:
            fused.push_async()

    def fwd():
        if fused is not None:
            return fwd_fused()
        rc = lib.avr_composite_fwd(x.data_ptr(), z.data_ptr(), rays, k, 1, 1.8, w.data_ptr(), rgb.data_ptr(),
                                   depth.data_ptr(), sp)
        assert rc == 0, lib.avr_last_cuda_error()

    def bwd():
        rc = lib.avr_composite_bwd(x.data_ptr(), z.data_ptr(), g_rgb.data_ptr(), g_d.data_ptr(), None, rays, k, 1, 1.8,
                                   dx.data_ptr(), None, sp)
        assert rc == 0, lib.avr_last_cuda_error()

    def step():
        fwd()
        if dist is not None and fused is None and args.gather != "none":  # the path's only exchange: per-ray outputs, 16 B/ray
            # in-stream: an async all-gather cannot co-schedule with the persistent backward
            # kernel (it holds every SM) and measured 2.6x slower than this
            avr_dist.all_gather_outputs(rgb, depth)
        bwd()
        if fused is not None and not ce and not no_barrier:
            fused.finish()

    no_barrier = os.environ.get("AVR_GATHER_NO_BARRIER", "0") == "1"   # diagnostic: kernel-only cost of the fused gather
    for _ in range(args.warmup):
        step()
    if ce:
        fused.wait_all()
    torch.cuda.synchronize(dev)

    # correctness spot-check against the oracle on a few rays (outside the timed region)
    if rank == 0:
        import avr_oracle as O

        pick = torch.arange(0, rays, max(rays // 256, 1), device=dev)[:256]
        want = O.composite_rgbs(z[pick].cpu().unsqueeze(0), x[pick].cpu().unsqueeze(0), True)
        err = (rgb[pick].cpu() - want[0][0]).abs().max().item()
        assert err < 1e-5, f"bench output differs from the oracle: {err}"

    # ---- timed region: events on the launching stream; per-kernel events for the roofline
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize(dev)
    with ClockSampler(local_rank) as clocks:
        t_begin.record(stream)
        for i in range(args.steps):
            ev[i][0].record(stream)
            fwd()
            ev[i][1].record(stream)
            if dist is not None and fused is None and args.gather != "none":
                avr_dist.all_gather_outputs(rgb, depth)
                ev[i][1] = torch.cuda.Event(enable_timing=True)
                ev[i][1].record(stream)
            bwd()
            ev[i][2].record(stream)
            if fused is not None and not ce and not no_barrier:
                fused.finish()
        if ce:
            fused.wait_all()     # every slot's pushes + barrier have completed before the clock stops
        t_end.record(stream)
        torch.cuda.synchronize(dev)
    if dist is not None:
        dist.barrier()
    total_ms = t_begin.elapsed_time(t_end)
    if dist is not None:
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    bwd_ms = sorted(ev[i][1].elapsed_time(ev[i][2]) for i in range(args.steps))
    fwd_ms = sorted(ev[i][0].elapsed_time(ev[i][1]) for i in range(args.steps)) if (dist is None or fused is not None or args.gather == "none") else None
    bwd_avg = sum(bwd_ms) / len(bwd_ms)

    ms_per_step = total_ms / args.steps
    value = world * rays / (ms_per_step * 1e-3)
    fb, bb = bytes_per_ray(k)
    peak, peak_src = measured_peak()
    roofline = {
        "kernel": f"composite_bwd_span_kernel<L={L.value}>" if span else "composite_bwd_ray_kernel",
        "bound": "hbm", "achieved": bb * rays / (bwd_avg * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
        "peak_source": peak_src,
        # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full
        # capture of this kernel at this size (profiles/r01_ncu_span_kernels.md); null for other shapes
        "traffic": NCU_TRAFFIC_BWD.get((rays, k)) if span else None,
        "algorithmic_bytes_per_launch": bb * rays, "avg_launch_ms": bwd_avg, "min_launch_ms": bwd_ms[0],
    }
    roofline["frac"] = roofline["achieved"] / peak
    extra = {}
    if fwd_ms is not None:
        fwd_avg = sum(fwd_ms) / len(fwd_ms)
        extra["roofline_fwd"] = {"kernel": f"composite_fwd_span_kernel<L={L.value},w>" if span else "composite_fwd_ray_kernel",
                                 "bound": "hbm", "achieved": fb * rays / (fwd_avg * 1e-3) / 1e9, "peak": peak,
                                 "unit": "GB/s", "frac": fb * rays / (fwd_avg * 1e-3) / 1e9 / peak,
                                 "avg_launch_ms": fwd_avg, "min_launch_ms": fwd_ms[0]}
        if dist is None:
            extra["step_hbm_frac"] = (fb + bb) * rays / (ms_per_step * 1e-3) / 1e9 / peak

    # ---- end to end through the host-buffer C-ABI call (pinned host memory both ways)
    e2e = None
    if not args.no_e2e:
        hx = torch.empty(rays, k, 4).pin_memory()
        hz = torch.empty(rays, k).pin_memory()
        hx.copy_(x)
        hz.copy_(z)
        hg, hd = g_rgb.cpu().pin_memory(), g_d.cpu().pin_memory()
        o_rgb, o_depth = torch.empty(rays, 3).pin_memory(), torch.empty(rays).pin_memory()
        o_dx = torch.empty(rays, k, 4).pin_memory()
        ws = ctypes.c_void_p()
        assert lib.avr_host_workspace_create(k, 0, ctypes.byref(ws)) == 0, lib.avr_last_cuda_error()

        def host_step():
            rc = lib.avr_composite_fwd_bwd_host(ws, hx.data_ptr(), hz.data_ptr(), hg.data_ptr(), hd.data_ptr(), rays, k,
                                                1, 1.8, o_rgb.data_ptr(), o_depth.data_ptr(), o_dx.data_ptr())
            assert rc == 0, lib.avr_last_cuda_error()

        e2e_steps = max(3, min(args.steps, 10))
        for _ in range(2):
            host_step()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            host_step()          # returns only when the outputs are in host memory
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        assert torch.equal(o_rgb[:4096], rgb[:4096].cpu()), "host path result differs from the device path"
        lib.avr_host_workspace_destroy(ws)
        e2e = {"value": world * rays * e2e_steps / dt, "unit": UNIT,
               "h2d_bytes_per_step": rays * (20 * k + 16), "d2h_bytes_per_step": rays * (16 * k + 16),
               "ms_per_step": 1e3 * dt / e2e_steps, "steps": e2e_steps,
               "api": "avr_composite_fwd_bwd_host (C ABI, pinned host buffers, 3-slot H2D/compute/D2H pipeline)"}

    # ---- the other BASELINE.json configs, briefly (N=1 only; reported beside the headline, never instead of it)
    others = None
    if world == 1 and not args.no_extras:
        try:
            others = measure_other_configs(dev, peak)
        except Exception as exc:  # the headline line must survive a failure here
            others = {"error": f"{type(exc).__name__}: {exc}"}
        # SURVEY 8(f) row 3, the radiance field's front end (tools/bench_field.py: raw C-ABI launches over
        # rotating inputs larger than L2; 2048 rays x 96 samples, conf/default.conf's 512 + 42 wide rows)
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import bench_field
            recs = list(bench_field.run(bench_field.parse_args(["--raw-only", "--iters", "5"]), dev))
            others["field_front_end_2048x96"] = {r["kernel"]: {k: r[k] for k in ("ms", "rows_per_s", "GBps", "hbm_frac")} for r in recs}
        except Exception as exc:
            others["field_front_end_2048x96"] = {"error": f"{type(exc).__name__}: {exc}"}

    # ---- CPU baseline on this box's host cores (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample_rays = 1 << 17
        times, cores = time_cpu(sample_rays, k, steps=5, warmup=1)
        cpu = {"value": sample_rays / min(times), "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"best of {len(times)} fwd+bwd passes over {sample_rays} rays x {k} samples "
                         f"(oracle/avr_oracle.py = the reference's torch-CPU op sequence, {cores} threads)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "rays_per_gpu": rays, "samples_per_ray": k, "outputs": "w, rgb, depth, d_rgbs",
                       "l2_policy": f"inputs larger than L2 ({(20 * k * rays) >> 20} MiB read per pass vs 126 MiB L2)",
                       "kernel_family": "span (TMA bulk-staged blocked scan)" if span else "generic",
                       "samples_per_lane": L.value, "rays_per_tile": rpt.value,
                       "collective": f"all-gather of rgb+depth (16 B/ray) per step: {gather_mode}"},
            "samples_per_sec": value * k,
            "clocks": clocks.summary(),
            "e2e": e2e, "gpu_launches": launches_per_step * args.steps,
            "roofline": roofline, "cpu_baseline": cpu,
        }
        if others is not None:
            line["other_configs"] = others
        line.update(extra)
        emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
