"""Multi-GPU plumbing: rays are independent, so they shard across ranks with no data-path
collective; the only exchange on this path is the all-gather of the per-ray outputs
(rgb + depth, 16 B/ray), packed as one (R, 4) buffer so it is a single NCCL call.

One process per GPU (torch.distributed, NCCL over NVLink on the GPU box; the same code
runs over gloo on CPU tensors, which is how the host logic is tested).  The reference is
single-process/single-GPU (train.py:238-240); this layer has no counterpart there.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_rays: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous ray range [start, stop) of `rank`; the first n_rays % world ranks get one extra."""
    base, extra = divmod(n_rays, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_bounds_packed(offsets: torch.Tensor, world: int, rank: int) -> Tuple[int, int]:
    """Ray range of `rank` for a packed layout, balanced by SAMPLES (not rays): boundaries
    are the rays whose start offset is nearest below rank/world of the sample stream."""
    total = int(offsets[-1])
    n_rays = offsets.numel() - 1
    targets = torch.tensor([total * rank // world, total * (rank + 1) // world], dtype=offsets.dtype,
                           device=offsets.device)
    cuts = torch.searchsorted(offsets[:-1].contiguous(), targets, right=False)
    start = 0 if rank == 0 else int(cuts[0])
    stop = n_rays if rank == world - 1 else int(cuts[1])
    return start, stop


class GatherHandle:
    """Result of an asynchronous output gather: `wait()` returns (rgb_all (R,3), depth_all (R,))."""

    def __init__(self, work, out, shard_sizes, n_max):
        self._work, self._out, self._sizes, self._n_max = work, out, shard_sizes, n_max

    def wait(self):
        if self._work is not None:
            self._work.wait()       # makes the current stream wait for the collective
            self._work = None
        out, n_max = self._out, self._n_max
        if any(s != n_max for s in self._sizes):
            out = torch.cat([out[r * n_max: r * n_max + s] for r, s in enumerate(self._sizes)], dim=0)
        return out[:, :3], out[:, 3]


def all_gather_outputs(rgb: torch.Tensor, depth: torch.Tensor, group: Optional[dist.ProcessGroup] = None,
                       shard_sizes: Optional[list] = None, async_op: bool = False):
    """Gather every rank's (R_local, 3) rgb and (R_local,) depth into (R, 3) / (R,), ordered by
    rank.  Equal shards take one all_gather_into_tensor; unequal shards are padded to the
    largest (pass `shard_sizes`, the per-rank ray counts).  With `async_op` the collective is
    only enqueued (it overlaps whatever the caller launches next, e.g. the backward kernel)
    and a GatherHandle is returned."""
    world = dist.get_world_size(group)
    packed = torch.cat([rgb.reshape(-1, 3), depth.reshape(-1, 1)], dim=-1).contiguous()
    n_local = packed.shape[0]
    if shard_sizes is None:
        shard_sizes = [n_local] * world
    n_max = max(shard_sizes)
    if n_local < n_max:
        packed = torch.cat([packed, packed.new_zeros(n_max - n_local, 4)], dim=0)
    out = packed.new_empty(world * n_max, 4)
    work = dist.all_gather_into_tensor(out, packed, group=group, async_op=async_op)
    handle = GatherHandle(work if async_op else None, out, shard_sizes, n_max)
    return handle if async_op else handle.wait()


def composite_sharded(rgbs: torch.Tensor, z: torch.Tensor, white_back: bool = True, infinity: float = 1.8,
                      group: Optional[dist.ProcessGroup] = None, composite_fn=None):
    """Each rank composites ITS rays (rgbs (R_local,K,4), z (R_local,K)) and receives the full
    image: returns (rgb_local, depth_local, rgb_all, depth_all).  Gradients flow through the
    local outputs only (each rank owns its rays' gradients; no backward collective)."""
    if composite_fn is None:
        from . import ops

        composite_fn = ops.composite
    rgb, depth, _ = composite_fn(rgbs, z, white_back, infinity, False)
    with torch.no_grad():
        world = dist.get_world_size(group)
        sizes = [None] * world
        n = torch.tensor([rgb.shape[0]], device=rgb.device)
        all_n = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(all_n, n, group=group)
        sizes = [int(t.item()) for t in all_n]
        rgb_all, depth_all = all_gather_outputs(rgb.detach(), depth.detach(), group, sizes)
    return rgb, depth, rgb_all, depth_all


class FusedGather:
    """All-gather of the per-ray outputs fused into the forward compositing kernel, with no
    cross-rank barrier on the step's critical path.

    Every rank owns a symmetric-memory buffer holding TWO slots of ``gathered`` rows — shape
    (R_total, 4) = (r,g,b,depth) per ray, ranks' shards back to back (shards may differ in
    size) — plus a small signal area.  The span kernel's epilogue stores each finished group
    of 32 rays straight into the current slot of EVERY rank's buffer (coalesced 512-byte
    stores to peer memory over NVLink / NVSwitch, or one ``multimem.st`` through the NVSwitch
    multicast mapping), and the last CTA of the launch then publishes the step number into
    word ``rank`` of every rank's flag array (release, system scope).  ``finish()`` enqueues a
    one-warp kernel that waits until all ``world`` words of the LOCAL flag array carry the
    step number (acquire) — so a rank only ever waits for data it is about to read, typically
    long after it arrived, instead of rendezvousing with every peer once per step.

    Contract (same as PipelinedGather): call ``finish()`` once per ``launch()``, and read the
    views it returns on the same stream BEFORE the next ``launch()``.  Steps alternate between
    the two slots: a peer can be at most one step ahead (its step i+2 forward is ordered after
    its own wait for step i+1, i.e. after this rank launched step i+1, i.e. after this rank's
    reads of step i), so it never overwrites rows that are still being read.

    Falls back (``available == False``) when symmetric memory cannot be set up; callers then
    use ``all_gather_outputs`` (NCCL).
    """

    SLOTS = 2
    _SIGNAL_ROWS = 64          # 1 KiB: flags[slot][32] at +0 / +128, done counter at +512, status at +516

    def __init__(self, rays_local: int, device, group: Optional[dist.ProcessGroup] = None):
        import ctypes
        import os

        self.rays_local = rays_local
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.available = False
        self.error = None
        self.step = 0              # launches so far; step i uses slot i % 2 and signal value i + 1
        self._waited = 0           # steps whose wait has been enqueued
        try:
            if self.world > 16:
                raise RuntimeError("the fused gather addresses at most 16 peers")
            import torch.distributed._symmetric_memory as symm_mem

            # shards may be uneven (shard_bounds): every rank needs every rank's row range
            n = torch.tensor([rays_local], dtype=torch.int64, device=device)
            sizes = [torch.zeros_like(n) for _ in range(self.world)]
            dist.all_gather(sizes, n, group=self.group)
            self.sizes = [int(t.item()) for t in sizes]
            self.row0 = sum(self.sizes[:self.rank])
            self.rows_total = sum(self.sizes)
            rows = self.SLOTS * self.rows_total + self._SIGNAL_ROWS
            self.buffer = symm_mem.empty((rows, 4), dtype=torch.float32, device=device)
            self.buffer.zero_()
            self.handle = symm_mem.rendezvous(self.buffer, self.group)
            base = [int(p) for p in self.handle.buffer_ptrs]
            slot_bytes = self.rows_total * 16
            sig = self.SLOTS * slot_bytes
            self._peer_rows = [(ctypes.c_void_p * self.world)(*[b + s * slot_bytes for b in base]) for s in range(self.SLOTS)]
            self._peer_flags = [(ctypes.c_void_p * self.world)(*[b + sig + s * 128 for b in base]) for s in range(self.SLOTS)]
            self._local_flags = [base[self.rank] + sig + s * 128 for s in range(self.SLOTS)]
            self._done_counter = base[self.rank] + sig + 512
            self._status_ptr = base[self.rank] + sig + 516
            # NVSwitch multicast mapping of the same buffers (0 when the box has no NVLS support);
            # opt-in (AVR_GATHER_MULTICAST=1): same step time as the per-peer stores at 2 and 4 GPUs
            mc = int(getattr(self.handle, "multicast_ptr", 0) or 0)
            self.multicast_ptr = mc if os.environ.get("AVR_GATHER_MULTICAST", "0") == "1" else 0
            self._mc_rows = [(ctypes.c_void_p * 1)(self.multicast_ptr + s * slot_bytes) for s in range(self.SLOTS)]
            torch.cuda.synchronize(device)
            self.handle.barrier()          # every rank's buffer is zeroed before anyone signals into it
            torch.cuda.synchronize(device)
            self.available = True
        except Exception as exc:  # pragma: no cover - depends on the box
            self.error = f"{type(exc).__name__}: {exc}"

    # the views of one slot
    def gathered(self, slot: int):
        g = self.buffer[slot * self.rows_total:(slot + 1) * self.rows_total]
        return g[:, :3], g[:, 3]

    def composite_fwd(self, rgbs: torch.Tensor, z: torch.Tensor, white_back: bool = True, infinity: float = 1.8,
                      want_w: bool = True):
        """Forward compositing of this rank's rays + fused gather.  Returns (rgb, depth, w) or
        None if the shape is not eligible for the fused kernel (caller falls back)."""
        from . import _lib
        from .ops import _f32c, _stream

        rgbs, z = _f32c(rgbs), _f32c(z)
        k = z.shape[-1]
        r = z.numel() // k
        if r != self.rays_local:
            raise _lib.AvrError(f"FusedGather was sized for {self.rays_local} rays per rank, got {r}")
        dev = z.device
        w = torch.empty(r, k, dtype=torch.float32, device=dev) if want_w else None
        rgb = torch.empty(r, 3, dtype=torch.float32, device=dev)
        depth = torch.empty(r, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = self.launch(_lib.load(), rgbs.data_ptr(), z.data_ptr(), r, k, int(bool(white_back)), float(infinity),
                             None if w is None else w.data_ptr(), rgb.data_ptr(), depth.data_ptr(), _stream(z))
        if rc == -4:   # AVR_ERR_UNSUPPORTED
            return None
        _lib.check(rc, "avr_composite_fwd_gather_signal")
        return rgb, depth, w

    def launch(self, lib, rgbs_ptr, z_ptr, r, k, white_back, infinity, w_ptr, rgb_ptr, depth_ptr, stream):
        """The raw C-ABI call for the next step (multicast stores when enabled, per-peer stores
        otherwise).  Returns the status; on success the step counter advances."""
        if self._waited != self.step:
            raise RuntimeError("FusedGather.launch: finish() of the previous step was not called")
        slot = self.step % self.SLOTS
        if self.multicast_ptr:
            rows, n_rows, mc = self._mc_rows[slot], 1, 1
        else:
            rows, n_rows, mc = self._peer_rows[slot], self.world, 0
        rc = lib.avr_composite_fwd_gather_signal(rgbs_ptr, z_ptr, r, k, white_back, infinity, w_ptr, rgb_ptr, depth_ptr,
                                                 rows, n_rows, mc, self.row0, self._peer_flags[slot], self.world,
                                                 self.rank, (self.step + 1) & 0xFFFFFFFF, self._done_counter, stream)
        if rc == 0:
            self.step += 1
        return rc

    def finish(self, stream: Optional[int] = None):
        """Stream-ordered wait for the rows of the last launched step from EVERY rank; returns the
        views (rgb_all (R,3), depth_all (R,)) of that step's slot.  No rendezvous: ranks that are
        ahead are not held back."""
        from . import _lib

        if self.step == 0 or self._waited == self.step:
            raise RuntimeError("FusedGather.finish: nothing launched since the last finish()")
        step = self.step - 1
        slot = step % self.SLOTS
        dev = self.buffer.device
        if stream is None:
            stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            _lib.check(_lib.load().avr_gather_wait(self._local_flags[slot], self.world, (step + 1) & 0xFFFFFFFF,
                                                   self._status_ptr, stream), "avr_gather_wait")
        self._waited = self.step
        return self.gathered(slot)

    def timed_out(self) -> bool:
        """True if a wait gave up (a peer never signalled).  Synchronises the device."""
        sig = self.buffer[self.SLOTS * self.rows_total:].view(torch.int32).reshape(-1)
        return bool(int(sig[129].item()))     # status word: byte 516 of the signal area


class PipelinedGather:
    """Copy-engine all-gather of the per-ray outputs, double-buffered so it never sits on the
    critical path.

    Step i: the forward kernel packs this rank's (r,g,b,depth) rows into slot ``i % slots`` of a
    symmetric-memory buffer (local 16-byte stores); a side stream then pushes those rows to
    every peer with ``cudaMemcpyAsync`` over NVLink (copy engines, no SMs) and runs the
    cross-rank barrier, while the caller's stream goes on with backward and with the NEXT step.
    ``wait(slot)`` orders the caller's stream after that slot's pushes + barrier.

    Why: on 8 B200s the pushes take 0.25 ms alone but ~1 ms while the persistent backward
    kernel saturates HBM, and a barrier inside every step pays the slowest rank's jitter; with
    one step of slack both disappear behind compute (tools/diag_gather.py, profiles/).

    Contract for consumers: read slot s (after ``wait(s)``) BEFORE launching the forward of
    the following step on the same stream — a peer's next push into that slot is ordered after
    this rank's following push, hence after the read.
    """

    def __init__(self, rays_local: int, device, group: Optional[dist.ProcessGroup] = None, slots: int = 2):
        import ctypes

        self.rays_local, self.slots = rays_local, slots
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.available, self.error = False, None
        self.slot = 0
        try:
            import torch.distributed._symmetric_memory as symm_mem

            n = torch.tensor([rays_local], dtype=torch.int64, device=device)
            sizes = [torch.zeros_like(n) for _ in range(self.world)]
            dist.all_gather(sizes, n, group=self.group)
            if any(int(t.item()) != rays_local for t in sizes):
                raise RuntimeError("PipelinedGather needs equal shards on every rank (FusedGather takes uneven ones)")
            rows = self.world * rays_local
            self.buffer = symm_mem.empty((slots * rows, 4), dtype=torch.float32, device=device)
            self.handle = symm_mem.rendezvous(self.buffer, self.group)
            base = [int(p) for p in self.handle.buffer_ptrs]
            self._peer_arrays = [(ctypes.c_void_p * self.world)(*[b + s * rows * 16 for b in base]) for s in range(slots)]
            self._local_arrays = [(ctypes.c_void_p * 1)(base[self.rank] + s * rows * 16) for s in range(slots)]
            self._side = torch.cuda.Stream(device=device)
            self._fwd_done = torch.cuda.Event()
            self._slot_done = [torch.cuda.Event() for _ in range(slots)]
            self._pending = [False] * slots
            self.handle.barrier()
            self.available = True
        except Exception as exc:  # pragma: no cover - depends on the box
            self.error = f"{type(exc).__name__}: {exc}"

    def gathered(self, slot: int):
        rows = self.world * self.rays_local
        g = self.buffer[slot * rows:(slot + 1) * rows]
        return g[:, :3], g[:, 3]

    def forward_target(self):
        """(ctypes pointer array, n_peers, row0) for avr_composite_fwd_gather into the current slot;
        orders the caller's stream after the slot's previous pushes."""
        self.wait(self.slot)
        return self._local_arrays[self.slot], 1, self.rank * self.rays_local

    def push_async(self) -> int:
        """Call right after the forward kernel has been enqueued on the current stream.  Returns
        the slot that is now travelling."""
        from . import _lib

        s = self.slot
        dev = self.buffer.device
        self._fwd_done.record(torch.cuda.current_stream(dev))
        self._side.wait_event(self._fwd_done)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().avr_gather_push_rows(self._peer_arrays[s], self.world, self.rank,
                                                        self.rank * self.rays_local, self.rays_local,
                                                        self._side.cuda_stream), "avr_gather_push_rows")
        with torch.cuda.stream(self._side):
            self.handle.barrier()
        self._slot_done[s].record(self._side)
        self._pending[s] = True
        self.slot = (s + 1) % self.slots
        return s

    def wait(self, slot: int):
        """Order the current stream after the pushes + barrier last issued for `slot`."""
        if self._pending[slot]:
            torch.cuda.current_stream(self.buffer.device).wait_event(self._slot_done[slot])
            self._pending[slot] = False
        return self.gathered(slot)

    def wait_all(self):
        for s in range(self.slots):
            self.wait(s)
