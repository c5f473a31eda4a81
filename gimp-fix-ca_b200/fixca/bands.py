"""fixca.bands -- one-process-per-GPU row banding.

Output row y of the correction pass depends only on (y, params, source rows near the mapped
coordinate) -- fix-ca.c:1091-1329 -- so an image splits into contiguous full-width row bands that are
computed independently; a band needs its own rows plus the halo rows `band_source_rows()` reports.
There is no data-path collective.  When one rank must own the whole frame, `PeerFrame` +
`run_band_into_frame()` make every rank's kernel store its band straight into that rank's memory (peer stores
over NVLink, compute and gather in one kernel); `gather_bands()` is the plain collective form of the same
(NCCL send/recv on GPUs, gloo in the CPU tests).

Nothing here computes pixels: `BandPlan.run()` calls the CUDA library; the CPU tests pass their own
`compute` callable (the oracle) to exercise the sharding logic without a GPU.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import (FixCaParams, band_source_rows, bpc_of, fix_ca_region_dev, fix_ca_region_dev_fanout, split_bands,
               PRECISION_EXACT, frame_alloc, frame_open, frame_close, frame_free)


@dataclass
class BandPlan:
    """What one rank of `world` holds for a height x width image."""

    rank: int
    world: int
    width: int
    height: int
    y1: int           # output rows [y1, y2) this rank produces
    y2: int
    src_lo: int       # source rows [src_lo, src_hi] it must hold (band + halo), inclusive
    src_hi: int

    @property
    def src_rows(self) -> int:
        return self.src_hi - self.src_lo + 1

    @property
    def halo_rows(self) -> int:
        return self.src_rows - (self.y2 - self.y1)


def plan_band(width: int, height: int, params: FixCaParams, rank: int, world: int, y1: int = 0, y2: int | None = None) -> BandPlan:
    y2 = height if y2 is None else y2
    b1, b2 = split_bands(y1, y2, world)[rank]
    if b1 == b2:
        return BandPlan(rank, world, width, height, b1, b2, b1, b1 - 1)
    lo, hi = band_source_rows(width, height, params, b1, b2)
    return BandPlan(rank, world, width, height, b1, b2, lo, hi)


def run_band_device(plan: BandPlan, d_src_ptr: int, src_pitch: int, d_dst_ptr: int, dst_pitch: int, bytes_per_pixel: int,
                    bpc: int, params: FixCaParams, flags: int = PRECISION_EXACT, stream: int = 0) -> None:
    """Launch this rank's band on its GPU: d_src holds rows [src_lo, src_hi], d_dst receives rows [y1, y2)."""
    if plan.y1 == plan.y2:
        return
    fix_ca_region_dev(d_src_ptr, src_pitch, plan.src_lo, plan.src_rows, d_dst_ptr, dst_pitch, plan.y1,
                      plan.width, plan.height, bytes_per_pixel, bpc, params, plan.y1, plan.y2, flags, stream)


class PeerFrame:
    """The whole destination frame on `owner`'s GPU, mapped into every rank of the box (CUDA IPC; peer access
    over NVLink / NVSwitch).  Each rank passes `ptr` as the destination of its band (`run_band_into_frame`):
    the kernel's TMA stores land in the owner's memory, so computing a band and gathering it are one kernel.
    Collective over `group`: every rank constructs it, `sync()`s before the owner reads, and `close()`s it."""

    def __init__(self, rows: int, pitch: int, owner: int = 0, group=None, row0: int = 0):
        import torch.distributed as dist

        # row0: the image row the frame's first row holds (a frame may cover a part of the image: several owners)
        self.rows, self.pitch, self.owner, self.group, self.row0 = rows, pitch, owner, group, row0
        self.rank = dist.get_rank(group)
        box = [None]
        if self.rank == owner:
            self.ptr, handle = frame_alloc(rows * pitch)
            box[0] = handle
        dist.broadcast_object_list(box, src=owner, group=group)
        if self.rank != owner:
            self.ptr = frame_open(box[0])
        self._open = True

    def sync(self) -> None:
        """Every rank's band is in the owner's frame: each rank drains its own stream, then the ranks meet."""
        import torch
        import torch.distributed as dist

        torch.cuda.synchronize()
        dist.barrier(group=self.group)

    def as_tensor(self):
        """Owner only: the frame as a (rows, pitch) uint8 torch tensor (a view of the allocation, no copy)."""
        import torch

        assert self.rank == self.owner

        class _Mem:     # __cuda_array_interface__ view of the raw allocation
            pass

        m = _Mem()
        m.__cuda_array_interface__ = {"shape": (self.rows, self.pitch), "typestr": "|u1", "data": (self.ptr, False),
                                      "version": 3}
        self._keep = m
        return torch.as_tensor(m, device="cuda")

    def close(self) -> None:
        import torch.distributed as dist

        if not self._open:
            return
        self._open = False
        if self.rank != self.owner:
            frame_close(self.ptr)
        dist.barrier(group=self.group)      # nobody maps the frame any more
        if self.rank == self.owner:
            frame_free(self.ptr)


def run_band_into_frame(plan: BandPlan, d_src_ptr: int, src_pitch: int, frame: PeerFrame, bytes_per_pixel: int,
                        bpc: int, params: FixCaParams, flags: int = PRECISION_EXACT, stream: int = 0) -> None:
    """Launch this rank's band with the owner's frame as its destination: rows [y1, y2) are written at their
    place in the whole frame (dst_row0 = 0), over NVLink when the frame lives on another GPU."""
    if plan.y1 == plan.y2:
        return
    fix_ca_region_dev(d_src_ptr, src_pitch, plan.src_lo, plan.src_rows, frame.ptr, frame.pitch, frame.row0,
                      plan.width, plan.height, bytes_per_pixel, bpc, params, plan.y1, plan.y2, flags, stream)


class AllFrames:
    """All-gather form: every rank owns a whole destination frame and maps everybody else's (CUDA IPC; peer access
    over NVLink / NVSwitch).  `run_band_into_all()` makes this rank's kernel store each finished chunk of its band
    into all of them in one launch, so after `sync()` every GPU holds the whole corrected frame.  Collective."""

    def __init__(self, rows: int, pitch: int, group=None):
        import torch.distributed as dist

        self.rows, self.pitch, self.group = rows, pitch, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.own, handle = frame_alloc(rows * pitch)
        handles = [None] * self.world
        dist.all_gather_object(handles, handle, group=group)
        self.ptrs = [self.own if r == self.rank else frame_open(handles[r]) for r in range(self.world)]
        self._open = True

    def sync(self) -> None:
        import torch
        import torch.distributed as dist

        torch.cuda.synchronize()
        dist.barrier(group=self.group)

    def as_tensor(self):
        """This rank's own frame as a (rows, pitch) uint8 torch tensor (a view, no copy)."""
        import torch

        class _Mem:
            pass

        m = _Mem()
        m.__cuda_array_interface__ = {"shape": (self.rows, self.pitch), "typestr": "|u1", "data": (self.own, False),
                                      "version": 3}
        self._keep = m
        return torch.as_tensor(m, device="cuda")

    def close(self) -> None:
        import torch.distributed as dist

        if not self._open:
            return
        self._open = False
        for r, p in enumerate(self.ptrs):
            if r != self.rank:
                frame_close(p)
        dist.barrier(group=self.group)
        frame_free(self.own)


def run_band_into_all(plan: BandPlan, d_src_ptr: int, src_pitch: int, frames: AllFrames, bytes_per_pixel: int,
                      bpc: int, params: FixCaParams, flags: int = PRECISION_EXACT, stream: int = 0) -> None:
    """Launch this rank's band with every rank's frame as a destination (own frame first): rows [y1, y2) land at
    their place in all `world` copies of the frame."""
    if plan.y1 == plan.y2:
        return
    order = [frames.own] + [p for r, p in enumerate(frames.ptrs) if r != frames.rank]
    fix_ca_region_dev_fanout(d_src_ptr, src_pitch, plan.src_lo, plan.src_rows, order, frames.pitch, 0,
                             plan.width, plan.height, bytes_per_pixel, bpc, params, plan.y1, plan.y2, flags, stream)


def gather_bands(band, plan: BandPlan, dst_rank: int = 0, group=None):
    """Reassemble the image on `dst_rank` from every rank's band (a torch tensor of its rows [y1, y2)).

    Bands may differ by one row in height, so each rank sends its own shape; returns the full tensor on
    dst_rank and None elsewhere.  One message per rank: the only inter-GPU traffic of the whole pass."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    rows = torch.tensor([band.shape[0]], dtype=torch.int64, device=band.device)
    all_rows = [torch.zeros_like(rows) for _ in range(world)]
    dist.all_gather(all_rows, rows, group=group)
    sizes = [int(r.item()) for r in all_rows]
    if rank == dst_rank:
        parts = [torch.empty((n,) + tuple(band.shape[1:]), dtype=band.dtype, device=band.device) for n in sizes]
        reqs = []
        for r in range(world):
            if r == rank:
                parts[r].copy_(band)
            elif sizes[r]:
                reqs.append(dist.irecv(parts[r], src=r, group=group))
        for q in reqs:
            q.wait()
        return torch.cat(parts, dim=0)
    if band.shape[0]:
        dist.send(band.contiguous(), dst=dst_rank, group=group)
    return None
