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
    """All-gather of the per-ray outputs fused into the forward compositing kernel.

    Every rank owns a symmetric-memory buffer ``gathered`` of shape (world*R_local, 4) =
    (r,g,b,depth) per ray; the span kernel's epilogue stores each finished ray straight into
    row ``rank*R_local + ray`` of EVERY rank's buffer (16-byte stores to peer memory over
    NVLink / NVSwitch), so the exchange rides along with the kernel instead of following it.
    ``finish()`` is the cross-rank barrier after which ``gathered`` is complete everywhere.
    Falls back (``available == False``) when symmetric memory cannot be set up; callers then
    use ``all_gather_outputs`` (NCCL).
    """

    def __init__(self, rays_local: int, device, group: Optional[dist.ProcessGroup] = None):
        import ctypes

        self.rays_local = rays_local
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.available = False
        self.error = None
        try:
            import torch.distributed._symmetric_memory as symm_mem

            self.gathered = symm_mem.empty((self.world * rays_local, 4), dtype=torch.float32, device=device)
            self.handle = symm_mem.rendezvous(self.gathered, self.group)
            ptrs = [int(p) for p in self.handle.buffer_ptrs]
            self._ptr_array = (ctypes.c_void_p * self.world)(*ptrs)
            # NVSwitch multicast mapping of the same buffers (0 when the box has no NVLS support)
            import os

            # opt-in (AVR_GATHER_MULTICAST=1): measured identical to the per-peer stores at 2 and 4 GPUs
            # (profiles/r01_multi_gpu.md) and not yet run at 8, so the per-peer path stays the default
            self.multicast_ptr = int(getattr(self.handle, "multicast_ptr", 0) or 0)
            if os.environ.get("AVR_GATHER_MULTICAST", "0") != "1":
                self.multicast_ptr = 0
            self.handle.barrier()
            self.available = True
        except Exception as exc:  # pragma: no cover - depends on the box
            self.error = f"{type(exc).__name__}: {exc}"

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
        _lib.check(rc, "avr_composite_fwd_gather")
        return rgb, depth, w

    def launch(self, lib, rgbs_ptr, z_ptr, r, k, white_back, infinity, w_ptr, rgb_ptr, depth_ptr, stream):
        """The raw C-ABI call: multicast stores when the box supports them, per-peer stores otherwise."""
        if self.multicast_ptr:
            return lib.avr_composite_fwd_gather_multicast(rgbs_ptr, z_ptr, r, k, white_back, infinity, w_ptr, rgb_ptr,
                                                          depth_ptr, self.multicast_ptr, self.rank * self.rays_local, stream)
        return lib.avr_composite_fwd_gather(rgbs_ptr, z_ptr, r, k, white_back, infinity, w_ptr, rgb_ptr, depth_ptr,
                                            self._ptr_array, self.world, self.rank * self.rays_local, stream)

    def finish(self):
        """Cross-rank barrier (stream-ordered): afterwards every rank's ``gathered`` holds all rays."""
        self.handle.barrier()
        return self.gathered[:, :3], self.gathered[:, 3]


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
