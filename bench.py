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
every rank composites its own 2^20-ray shard) and the only exchange is the all-gather of the
per-ray outputs rgb+depth (16 B/ray), fused into the forward kernel and consumed inside the
timed step; `other_configs` then also carries BASELINE.json configs[4] as stated (2^24 rays x
192 samples split over the ranks, strong scaling).
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

# (rays, K) -> DRAM bytes per launch of composite_bwd_span_kernel<9,3> from the committed ncu --set full
# capture (dram__bytes_read.sum + dram__bytes_write.sum; raw CSV under profiles/)
NCU_TRAFFIC_BWD = {(1 << 20, 96): {"bytes": 2.030070e9 + 1.564075e9, "source": "profiles/r02_ncu_span_final.raw.csv (dram__bytes_read.sum + dram__bytes_write.sum, composite_bwd_span_kernel<9,3,1,0,0>)"}}

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


# ------------------------------------------------------------------ small helpers
def _timeit(fn, dev, iters=10, warm=3):
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


def _entry(ms, rays, bpr, peak):
    gbs = bpr * rays / (ms * 1e-3) / 1e9
    return {"ms": round(ms, 4), "rays_per_s": rays / (ms * 1e-3), "GBps": round(gbs, 1), "hbm_frac": round(gbs / peak, 4)}


def synth_dense(rays, k, dev, seed, chunk=1 << 18):
    """SURVEY.md 8(d) C2 inputs, generated chunk by chunk straight into their final buffers (the
    2^24 x 192 case is 64 GB of inputs: no whole-tensor temporaries)."""
    g = torch.Generator(device=dev).manual_seed(seed)
    z = torch.empty(rays, k, device=dev)
    x = torch.empty(rays, k, 4, device=dev)
    for lo in range(0, rays, chunk):
        hi = min(rays, lo + chunk)
        n = hi - lo
        z[lo:hi] = torch.sort(0.8 + torch.rand(n, k, device=dev, generator=g), -1).values
        x[lo:hi, :, :3] = torch.sigmoid(torch.randn(n, k, 3, device=dev, generator=g))
        x[lo:hi, :, 3] = torch.relu(torch.randn(n, k, device=dev, generator=g)) * 30
    g_rgb = torch.randn(rays, 3, device=dev, generator=g)
    g_d = torch.randn(rays, device=dev, generator=g)
    return z, x, g_rgb, g_d


class CompositeStep:
    """One forward+backward compositing pass over resident inputs through the C ABI, with the
    output all-gather when there is more than one rank."""

    def __init__(self, lib, rays, k, dev, seed, dist=None, gather="fused"):
        from avr_b200 import dist as avr_dist

        self.lib, self.rays, self.k, self.dev, self.dist, self.avr_dist = lib, rays, k, dev, dist, avr_dist
        self.z, self.x, self.g_rgb, self.g_d = synth_dense(rays, k, dev, seed)
        self.w = torch.empty(rays, k, device=dev)
        self.rgb = torch.empty(rays, 3, device=dev)
        self.depth = torch.empty(rays, device=dev)
        self.dx = torch.empty_like(self.x)
        self.stream = torch.cuda.current_stream(dev)
        self.sp = self.stream.cuda_stream
        self.fused, self.ce, self.mode = None, False, "none"
        self.gathered = None
        if dist is not None and gather in ("fused", "ce"):
            fused = avr_dist.PipelinedGather(rays, dev) if gather == "ce" else avr_dist.FusedGather(rays, dev)
            ok = torch.tensor([1 if fused.available else 0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()):
                self.fused, self.ce = fused, gather == "ce"
            elif dist.get_rank() == 0:
                print(f"[bench] fused gather unavailable ({fused.error}); using NCCL", file=sys.stderr)
        if dist is not None and gather != "none":
            if self.fused is None:
                self.mode = "NCCL all_gather_into_tensor between forward and backward"
            elif self.ce:
                self.mode = ("forward kernel packs local rows into a double-buffered symmetric-memory slot; copy-engine pushes "
                             "over NVLink + cross-rank barrier on a side stream (all waited inside the timed region)")
            else:
                how = ("512-byte multimem.st through the NVSwitch multicast mapping" if self.fused.multicast_ptr
                       else "coalesced 512-byte peer stores over NVLink")
                self.mode = (f"fused into the forward kernel ({how} into double-buffered symmetric memory); the kernel's last CTA "
                             "publishes a step flag to every rank and a one-warp wait kernel after backward consumes the flags "
                             "(no cross-rank barrier)")
        self.nccl = dist is not None and self.fused is None and gather != "none"

    def fwd(self):
        lib = self.lib
        if self.fused is not None and self.ce:
            target, n_t, row0 = self.fused.forward_target()
            rc = lib.avr_composite_fwd_gather(self.x.data_ptr(), self.z.data_ptr(), self.rays, self.k, 1, 1.8, self.w.data_ptr(),
                                              self.rgb.data_ptr(), self.depth.data_ptr(), target, n_t, row0, self.sp)
            assert rc == 0, (rc, lib.avr_last_cuda_error())
            self.fused.push_async()
        elif self.fused is not None:
            rc = self.fused.launch(lib, self.x.data_ptr(), self.z.data_ptr(), self.rays, self.k, 1, 1.8, self.w.data_ptr(),
                                   self.rgb.data_ptr(), self.depth.data_ptr(), self.sp)
            assert rc == 0, (rc, lib.avr_last_cuda_error())
        else:
            rc = lib.avr_composite_fwd(self.x.data_ptr(), self.z.data_ptr(), self.rays, self.k, 1, 1.8, self.w.data_ptr(),
                                       self.rgb.data_ptr(), self.depth.data_ptr(), self.sp)
            assert rc == 0, lib.avr_last_cuda_error()

    def exchange(self):
        """NCCL arm only: in-stream (an async all-gather cannot co-schedule with the persistent
        backward kernel — it holds every SM — and measured 2.6x slower than this)."""
        if self.nccl:
            self.gathered = self.avr_dist.all_gather_outputs(self.rgb, self.depth)

    def bwd(self):
        rc = self.lib.avr_composite_bwd(self.x.data_ptr(), self.z.data_ptr(), self.g_rgb.data_ptr(), self.g_d.data_ptr(), None,
                                        self.rays, self.k, 1, 1.8, self.dx.data_ptr(), None, self.sp)
        assert rc == 0, self.lib.avr_last_cuda_error()

    def finish(self):
        """The consumer side of the exchange: after this, the step's gathered image is readable."""
        if self.fused is not None and not self.ce:
            self.gathered = self.fused.finish(self.sp)

    def step(self):
        self.fwd()
        self.exchange()
        self.bwd()
        self.finish()

    @property
    def launches_per_step(self):
        L, rpt, main_rays = ctypes.c_int(), ctypes.c_int(), ctypes.c_int64()
        span = self.lib.avr_composite_plan_info(self.rays, self.k, self.x.data_ptr(), self.z.data_ptr(), ctypes.byref(L),
                                                ctypes.byref(rpt), ctypes.byref(main_rays))
        n = 2 * (1 + (1 if (span and main_rays.value < self.rays) else 0)) if span else 2
        return n + (1 if (self.fused is not None and not self.ce) else 0), span, L.value, rpt.value

    def check_gather(self):
        """Outside the timer: the gathered image of the last step equals an NCCL all-gather of every
        rank's local outputs, bit for bit, and this rank's rows equal its local outputs."""
        if self.dist is None or self.gathered is None:
            return None
        if self.ce:
            self.fused.wait_all()
            self.gathered = self.fused.gathered((self.fused.slot - 1) % self.fused.slots)
        g_rgb, g_depth = self.gathered
        want_rgb, want_depth = self.avr_dist.all_gather_outputs(self.rgb, self.depth)
        ok = torch.equal(g_rgb, want_rgb) and torch.equal(g_depth, want_depth)
        r0 = self.dist.get_rank() * self.rays
        ok = ok and torch.equal(g_rgb[r0:r0 + self.rays], self.rgb)
        if self.fused is not None and not self.ce:
            ok = ok and not self.fused.timed_out()
        t = torch.tensor([1 if ok else 0], device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return bool(int(t.item()))


def timed_steps(cs, steps, warmup, dev, dist, clocks=None):
    """W untimed steps, then exactly `steps` steps between a device-side cross-rank rendezvous +
    event and an event + synchronise; per-kernel events inside.  Returns max-over-ranks total ms and
    the per-step / per-kernel spans of THIS rank."""
    for _ in range(warmup):
        cs.step()
    if cs.ce:
        cs.fused.wait_all()
    torch.cuda.synchronize(dev)
    stream = cs.stream
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(steps)]
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync = torch.zeros(1, device=dev)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize(dev)
    import contextlib
    with (clocks if clocks is not None else contextlib.nullcontext()):
        if dist is not None:
            # device-side rendezvous on the launching stream: every rank's first timed kernel starts
            # when the slowest rank's host gets here, so host start skew is outside the timer
            dist.all_reduce(sync)
        t_begin.record(stream)
        for i in range(steps):
            ev[i][0].record(stream)
            cs.fwd()
            cs.exchange()
            ev[i][1].record(stream)
            cs.bwd()
            ev[i][2].record(stream)
            cs.finish()
            ev[i][3].record(stream)
        if cs.ce:
            cs.fused.wait_all()     # every slot's pushes + barrier have completed before the clock stops
        t_end.record(stream)
        torch.cuda.synchronize(dev)
    if dist is not None:
        dist.barrier()
    total_ms = t_begin.elapsed_time(t_end)
    local_ms = total_ms
    if dist is not None:
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    spans = sorted(ev[i][0].elapsed_time(ev[i][3]) for i in range(steps))
    return {
        "total_ms": total_ms, "local_total_ms": local_ms,
        "fwd_ms": sorted(ev[i][0].elapsed_time(ev[i][1]) for i in range(steps)),
        "bwd_ms": sorted(ev[i][1].elapsed_time(ev[i][2]) for i in range(steps)),
        "wait_ms": sorted(ev[i][2].elapsed_time(ev[i][3]) for i in range(steps)),
        "step_ms_median": spans[len(spans) // 2],
    }


# ------------------------------------------------------------------ other configs (short)
def measure_other_configs(dev, peak, lib):
    """BASELINE.json configs 1, 3, 4, 5 on one GPU, a few iterations each (CUDA events; inputs larger
    than L2).  ms per launch, rays/s and fraction of the measured HBM roofline from the algorithmic
    bytes of SURVEY.md section 8(d)."""
    from avr_b200 import _lib, ops

    g = torch.Generator(device=dev).manual_seed(1)
    out = {}

    def guarded(name, fn):
        try:
            out[name] = fn()
        except Exception as exc:  # the headline line must survive a failure here
            out[name] = {"error": f"{type(exc).__name__}: {exc}"}
        torch.cuda.empty_cache()

    r = 1 << 20
    near, far = torch.tensor([0.8], device=dev), torch.tensor([1.8], device=dev)

    # config 3: importance sampling 64 coarse weights -> 128 fine samples (+ merge with the coarse depths)
    def c3():
        w = torch.rand(1, r, 64, device=dev, generator=g) ** 6
        u, u2 = torch.rand(1, r, 128, device=dev, generator=g), torch.rand(1, r, 128, device=dev, generator=g)
        zc = ops.coarse_sample_raw(near, far, 0, torch.rand(1, r, 64, device=dev, generator=g))
        _lib.dispatch_reset()
        res = {"with_merge": _entry(_timeit(lambda: ops.importance_sample(w, near, far, u, u2, z_coarse=zc, want_fine=False,
                                                                          want_sorted=True), dev), r, 2312, peak)}
        res["with_merge"]["kernel"] = [k for k, v in _lib.dispatch_counters().items() if v]
        res["sampling_only"] = _entry(_timeit(lambda: ops.importance_sample(w, near, far, u, u2, want_fine=True), dev), r, 1800, peak)
        # conf/default.conf's shape: 64 coarse -> 16 importance + 16 depth, merged (renderers.py:252-258)
        u16, n16 = u[..., :16].contiguous(), torch.randn(1, r, 16, device=dev, generator=g)
        res["conf_64_to_16_16_with_merge"] = _entry(_timeit(lambda: ops.importance_sample(
            w, near, far, u16, u16, z_coarse=zc, normals=n16, depth_std=0.01, want_fine=False, want_sorted=True), dev),
            r, 12 * 64 + 12 * 16 + 4 * 16 + 4 * 16 + 8, peak)
        return res
    guarded("c3_importance_64_to_128", c3)

    # the north-star sentence as one number: every kernel of a VolumeRenderer render (sampling + compositing,
    # forward and backward, conf/default.conf's 64 + 16 + 16 samples) over 2^20 rays (tools/bench_dense_pipeline.py)
    def dense():
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_dense_pipeline
        recs = list(bench_dense_pipeline.run(1 << 20, 3, dev, warm=2))
        res = {rec["stage"]: {k: rec[k] for k in ("ms", "GBps", "hbm_frac")} for rec in recs if "stage" in rec}
        res["pipeline"] = {k: v for k, v in recs[-1].items() if k != "pipeline"}
        return res
    guarded("sample_plus_composite_pipeline_2^20_rays_64+32", dense)

    # config 4: the whole ragged pipeline, 2^22 rays with 8..256 samples (tools/bench_packed_pipeline.py)
    def c4():
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_packed_pipeline
        recs = list(bench_packed_pipeline.run(1 << 22, 3, dev, warm=1))
        res = {rec["stage"]: {k: rec[k] for k in ("ms", "GBps", "hbm_frac")} for rec in recs if "stage" in rec}
        res["pipeline"] = {k: v for k, v in recs[-1].items() if k != "pipeline"}
        return res
    guarded("c4_ragged_pipeline_2^22_rays_8_to_256", c4)

    # config 5's per-GPU shard at 8 GPUs (2^21 x 192) and the WHOLE config on this one GPU (2^24 x 192,
    # 129 GB resident) — the single-GPU end of the strong-scaling curve that --gpus N continues
    def c5(rays):
        def run():
            need = rays * 192 * 40 + (6 << 30)
            free, _total = torch.cuda.mem_get_info(dev)
            if free < need:
                return {"skipped": f"needs {need >> 30} GiB, {free >> 30} GiB free"}
            cs = CompositeStep(lib, rays, 192, dev, seed=5)
            t = timed_steps(cs, 3, 2, dev, None)
            fb, bb = bytes_per_ray(192)
            res = _entry(t["total_ms"] / 3, rays, fb + bb, peak)
            res["fwd"] = _entry(sum(t["fwd_ms"]) / 3, rays, fb, peak)
            res["bwd"] = _entry(sum(t["bwd_ms"]) / 3, rays, bb, peak)
            # the last rays live above the 2^32-sample / 48 GB marks: check them against the oracle
            import avr_oracle as O
            pick = torch.arange(rays - 64, rays, device=dev)
            want = O.composite_rgbs(cs.z[pick].cpu().unsqueeze(0), cs.x[pick].cpu().unsqueeze(0), True)
            res["tail_rays_max_err_vs_oracle"] = float((cs.rgb[pick].cpu() - want[0][0]).abs().max())
            return res
        return run
    guarded("c5_shard_2^21x192", c5(1 << 21))
    guarded("c5_whole_2^24x192_one_gpu", c5(1 << 24))

    # config 1 through the drop-in API next to the reference's op sequence on the same GPU
    def c1():
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_dropin
        return bench_dropin.run(dev, iters=10)
    guarded("c1_dropin_volume_renderer", c1)

    # SURVEY 8(f) row 4, the adaptive renderer's LSTM ray march: one persistent kernel vs the reference's loop in torch
    def march():
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_march
        return bench_march.run(dev, iters=10)
    guarded("lstm_ray_march_10_steps_c512", march)

    # SURVEY 8(f) row 3, the radiance field's front end (tools/bench_field.py: raw C-ABI launches over
    # rotating inputs larger than L2; 2048 rays x 96 samples, conf/default.conf's 512 + 42 wide rows)
    def field():
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_field
        recs = list(bench_field.run(bench_field.parse_args(["--raw-only", "--iters", "5"]), dev))
        return {rec["kernel"]: {k: rec[k] for k in ("ms", "rows_per_s", "GBps", "hbm_frac")} for rec in recs}
    guarded("field_front_end_2048x96", field)
    return out


def measure_c5_strong(lib, dev, dist, world, rank, peak, gather):
    """BASELINE.json configs[4] as stated: 2^24 rays x 192 samples split over the N ranks (strong
    scaling), fwd + bwd + the output all-gather, same timing protocol as the headline."""
    total_rays = 1 << 24
    rays = total_rays // world
    need = rays * 192 * 40 + (6 << 30)
    free, _ = torch.cuda.mem_get_info(dev)
    ok = torch.tensor([1 if free >= need else 0], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if not int(ok.item()):
        return {"skipped": f"needs {need >> 30} GiB per GPU"}
    cs = CompositeStep(lib, rays, 192, dev, seed=100 + rank, dist=dist, gather=gather)
    steps = 5
    t = timed_steps(cs, steps, 2, dev, dist)
    gather_ok = cs.check_gather()
    fb, bb = bytes_per_ray(192)
    ms = t["total_ms"] / steps
    return {"workload": "2^24 rays x 192 samples split over the ranks (strong scaling), fwd+bwd + all-gather of rgb+depth",
            "rays_total": total_rays, "rays_per_gpu": rays, "ms_per_step": ms, "rays_per_s": total_rays / (ms * 1e-3),
            "samples_per_s": total_rays * 192 / (ms * 1e-3), "step_ms_median_rank0": t["step_ms_median"],
            "per_gpu_hbm_frac": (fb + bb) * rays / (ms * 1e-3) / 1e9 / peak,
            "fwd_ms": sum(t["fwd_ms"]) / steps, "bwd_ms": sum(t["bwd_ms"]) / steps, "wait_ms": sum(t["wait_ms"]) / steps,
            "gathered_equals_nccl_all_gather": gather_ok, "collective": cs.mode}


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
                         "to each peer's symmetric-memory buffer as one coalesced 512-byte store over NVLink and signals "
                         "completion with per-rank flags (no barrier); 'ce' = the kernel packs local rows and the copy "
                         "engines push them on a side stream (double-buffered; starved by the HBM-saturating kernels, kept "
                         "for comparison); 'nccl' = all_gather_into_tensor between forward and backward (both fall back to "
                         "nccl); 'none' = no exchange (diagnostic: the compute-only step under the same launch)")
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

    lib = avr_b200.load_library()
    assert lib.avr_device_check() == 0, lib.avr_last_cuda_error()
    clocks = ClockSampler(local_rank)     # NVML init happens HERE, long before the timed region

    # ---- synthetic inputs, generated on the device (seed = rank), SURVEY.md 8(d) C2
    cs = CompositeStep(lib, rays, k, dev, seed=rank, dist=dist, gather=args.gather)
    launches_per_step, span, L, rpt = cs.launches_per_step

    # correctness spot-check against the oracle on a few rays (outside the timed region)
    cs.step()
    torch.cuda.synchronize(dev)
    if rank == 0:
        import avr_oracle as O

        pick = torch.arange(0, rays, max(rays // 256, 1), device=dev)[:256]
        want = O.composite_rgbs(cs.z[pick].cpu().unsqueeze(0), cs.x[pick].cpu().unsqueeze(0), True)
        err = (cs.rgb[pick].cpu() - want[0][0]).abs().max().item()
        assert err < 1e-5, f"bench output differs from the oracle: {err}"

    # ---- timed region: events on the launching stream; per-kernel events for the roofline
    t = timed_steps(cs, args.steps, args.warmup, dev, dist, clocks)
    gather_ok = cs.check_gather()
    total_ms = t["total_ms"]
    bwd_ms, fwd_ms = t["bwd_ms"], (t["fwd_ms"] if not cs.nccl else None)
    bwd_avg = sum(bwd_ms) / len(bwd_ms)

    ms_per_step = total_ms / args.steps
    value = world * rays / (ms_per_step * 1e-3)
    fb, bb = bytes_per_ray(k)
    peak, peak_src = measured_peak()
    traffic = NCU_TRAFFIC_BWD.get((rays, k)) if span else None
    roofline = {
        "kernel": f"composite_bwd_span_kernel<L={L}>" if span else "composite_bwd_ray_kernel",
        "bound": "hbm", "achieved": bb * rays / (bwd_avg * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
        "peak_source": peak_src,
        # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full
        # capture of this kernel at this size; null for other shapes
        "traffic": traffic["bytes"] if traffic else None, "traffic_source": traffic["source"] if traffic else None,
        "algorithmic_bytes_per_launch": bb * rays, "avg_launch_ms": bwd_avg, "min_launch_ms": bwd_ms[0],
    }
    roofline["frac"] = roofline["achieved"] / peak
    extra = {"step_ms_median": t["step_ms_median"]}
    if fwd_ms is not None:
        fwd_avg = sum(fwd_ms) / len(fwd_ms)
        extra["roofline_fwd"] = {"kernel": f"composite_fwd_span_kernel<L={L},w>" if span else "composite_fwd_ray_kernel",
                                 "bound": "hbm", "achieved": fb * rays / (fwd_avg * 1e-3) / 1e9, "peak": peak,
                                 "unit": "GB/s", "frac": fb * rays / (fwd_avg * 1e-3) / 1e9 / peak,
                                 "avg_launch_ms": fwd_avg, "min_launch_ms": fwd_ms[0]}
        if dist is None:
            extra["step_hbm_frac"] = (fb + bb) * rays / (ms_per_step * 1e-3) / 1e9 / peak
    if dist is not None:
        extra["multi_gpu"] = {"rank0_total_ms": t["local_total_ms"], "max_over_ranks_total_ms": total_ms,
                              "fwd_ms_avg": sum(t["fwd_ms"]) / args.steps, "bwd_ms_avg": bwd_avg,
                              "wait_ms_avg": sum(t["wait_ms"]) / args.steps,
                              "gathered_equals_nccl_all_gather": gather_ok}

    # ---- end to end through the host-buffer C-ABI call (pinned host memory both ways)
    e2e = None
    if not args.no_e2e:
        hx = torch.empty(rays, k, 4).pin_memory()
        hz = torch.empty(rays, k).pin_memory()
        hx.copy_(cs.x)
        hz.copy_(cs.z)
        hg, hd = cs.g_rgb.cpu().pin_memory(), cs.g_d.cpu().pin_memory()
        o_rgb, o_depth = torch.empty(rays, 3).pin_memory(), torch.empty(rays).pin_memory()
        o_w = torch.empty(rays, k).pin_memory()
        o_dx = torch.empty(rays, k, 4).pin_memory()
        ws = ctypes.c_void_p()
        assert lib.avr_host_workspace_create(k, 0, ctypes.byref(ws)) == 0, lib.avr_last_cuda_error()

        def host_step():
            rc = lib.avr_composite_fwd_bwd_host(ws, hx.data_ptr(), hz.data_ptr(), hg.data_ptr(), hd.data_ptr(), rays, k,
                                                1, 1.8, o_rgb.data_ptr(), o_depth.data_ptr(), o_w.data_ptr(), o_dx.data_ptr())
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
            tt = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        assert torch.equal(o_rgb[:4096], cs.rgb[:4096].cpu()), "host path result differs from the device path"
        lib.avr_host_workspace_destroy(ws)
        e2e = {"value": world * rays * e2e_steps / dt, "unit": UNIT,
               "h2d_bytes_per_step": rays * (20 * k + 16), "d2h_bytes_per_step": rays * (20 * k + 16),
               "ms_per_step": 1e3 * dt / e2e_steps, "steps": e2e_steps,
               "outputs": "rgb, depth, w, d_rgbs (the same work as `value`)",
               "api": "avr_composite_fwd_bwd_host (C ABI, pinned host buffers, 3-slot H2D/compute/D2H pipeline)"}
        del hx, hz, o_dx, o_w

    mode = cs.mode
    # ---- the other BASELINE.json configs, briefly (reported beside the headline, never instead of it)
    others = None
    if not args.no_extras:
        del cs
        torch.cuda.empty_cache()
        if world == 1:
            others = measure_other_configs(dev, peak, lib)
        else:
            try:
                others = {"c5_strong_2^24x192": measure_c5_strong(lib, dev, dist, world, rank, peak, args.gather)}
            except Exception as exc:
                others = {"c5_strong_2^24x192": {"error": f"{type(exc).__name__}: {exc}"}}

    # ---- CPU baseline on this box's host cores (rank 0, N=1 only): the same estimator as --impl reference
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample_rays = 1 << 18
        times, cores = time_cpu(sample_rays, k, steps=5, warmup=1)
        cpu = {"value": sample_rays * len(times) / sum(times), "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"mean of {len(times)} fwd+bwd passes over {sample_rays} rays x {k} samples "
                         f"(oracle/avr_oracle.py = the reference's torch-CPU op sequence, {cores} threads)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "rays_per_gpu": rays, "samples_per_ray": k, "outputs": "w, rgb, depth, d_rgbs",
                       "l2_policy": f"inputs larger than L2 ({(20 * k * rays) >> 20} MiB read per pass vs 126 MiB L2)",
                       "kernel_family": "span (TMA bulk-staged blocked scan)" if span else "generic",
                       "samples_per_lane": L, "rays_per_tile": rpt,
                       "collective": f"all-gather of rgb+depth (16 B/ray) per step: {mode}"},
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
