#!/usr/bin/env python
"""bench.py -- throughput of Fix-CA's per-pixel correction pass on B200, in megapixels/s.

Contract (one JSON line on stdout, printed by rank 0):

  python bench.py --gpus N --steps K --warmup W            # the CUDA path (this repo)
  python bench.py --impl reference --gpus N ...            # the reference's own CPU code

A *step* is one pass of the hot path (the reference's fix_ca_region row loop,
fix-ca.c:1091-1333) over one rank's row band.  The default workload is the one
BASELINE.json's metric is quoted on: 12288 x 8192 RGB16 (100.7 MP), Cubic, lateral
blue 3.0 / red -2.0 plus directional shifts, lens at the image centre (SURVEY.md 8(d)
"target").  With N ranks the image is N times taller and every rank owns one contiguous
full-width band of 8192 rows plus the halo rows fixca_band_source_rows() reports (weak
scaling, bands are independent: no data-path collective).

  value      whole-job MP/s, bands resident in HBM, K launches timed with CUDA events on the
             launching stream, max over ranks.
  e2e        the same band through the reference-facing C ABI fixca_cuda_region_ex() with
             pinned HOST buffers: H2D, kernels and D2H all inside the timed region.
  roofline   algorithmic bytes (2 * bytes_per_pixel * pixels) / mean launch
             time, against MEASURED_PEAKS.json's hbm_gbs.
  cpu_baseline  the reference's own fix-ca.c (oracle/_ref, compiled unmodified) on this host's
             cores, rank 0, N=1 only, on a bounded row sample of the same image.

The oracle is used here only for cpu_baseline / --impl reference and for a spot parity
check of the e2e output; the measured CUDA path never touches it.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(ROOT, "gimp-fix-ca_b200"), os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

# ---------------------------------------------------------------------------------------------
# workloads (BASELINE.json configs / SURVEY.md 8(d)); width, rows per rank, channels, dtype
# ---------------------------------------------------------------------------------------------
DIRECTIONAL = dict(x_blue=0.7, x_red=-0.4, y_blue=0.3, y_red=-0.9)
WORKLOADS = {
    # name: (W, H per rank, channels, numpy dtype, interpolation, params, lens ("centre" or (x, y)))
    "target_100mp_rgb16_cubic": (12288, 8192, 3, "u2", 2, dict(blue=3.0, red=-2.0, **DIRECTIONAL), "centre"),
    "cfg2_24mp_rgb8_linear": (6000, 4000, 3, "u1", 1, dict(blue=1.0, red=-1.5), "centre"),
    "cfg3_8k_rgba16_cubic": (7680, 4320, 4, "u2", 2, dict(blue=6.0, red=-2.4), (658, 1280)),
    "cfg4_50mp_rgbf32_cubic": (8192, 6144, 3, "f4", 2, dict(blue=3.0, red=-2.0, **DIRECTIONAL), "centre"),
    "cfg5_4k_rgb8_cubic": (3840, 2160, 3, "u1", 2, dict(blue=1.0, red=-1.5, **DIRECTIONAL), "centre"),
}
# BASELINE configs[4]: a batch of frames per GPU (frames sharded by index, no communication): the step is ONE
# fixca_cuda_frames_dev launch over every frame this rank holds
WORKLOADS["cfg5_batch_4k_rgb8_cubic"] = WORKLOADS["cfg5_4k_rgb8_cubic"]
WORKLOAD_FRAMES = {"cfg5_batch_4k_rgb8_cubic": 128}
DEFAULT_WORKLOAD = "target_100mp_rgb16_cubic"
INTERP_NAME = {0: "none", 1: "linear", 2: "cubic"}


def bpc_of(dt: np.dtype) -> int:
    return -dt.itemsize if dt.kind == "f" else dt.itemsize


def measured_peak():
    """HBM copy bandwidth measured on this pool (driver-written), else the recipe's fallback."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


def ncu_traffic(workload: str, kernel: str):
    """dram read+write bytes per launch of the dominant kernel from the committed ncu capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        e = t.get(workload, {}).get(kernel)
        return None if e is None else float(e["dram_bytes_per_launch"])
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------
# clocks during the timed region
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and clock-event (throttle) reasons of one GPU through NVML while the
    timed region runs."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
               ("sw_power_cap", 0x4), ("hw_power_brake", 0x80), ("sync_boost", 0x10),
               ("applications_clocks_setting", 0x2), ("display_clocks_setting", 0x100))

    def __init__(self, index: int):
        self.samples, self.mask, self.max_mhz, self.err = [], 0, None, None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self.nv, self.err = None, repr(e)

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    self.mask |= nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    self.mask |= nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            except Exception as e:  # pragma: no cover
                self.err = repr(e)
                return
            time.sleep(0.0005)

    def start(self):
        if self.nv is not None:
            self._stop.clear()
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr is not None:
            self._stop.set()
            self._thr.join()
            self._thr = None

    def report(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "samples": len(s),
                "reasons": [n for n, bit in self.REASONS if self.mask & bit],
                **({"error": self.err} if self.err else {})}


# ---------------------------------------------------------------------------------------------
# the reference on host cores (cpu_baseline and --impl reference)
# ---------------------------------------------------------------------------------------------
def cpu_checker():
    import oracle as orc
    return orc.best_checker(), orc


def host_image(orc, h, w, ch, dt, seed):
    """Seeded uniform-noise image; a few rows are drawn and tiled (PCG64 over 300 M samples
    would dominate the run) with a per-tile roll so that rows differ."""
    base_rows = min(h, 256)
    base = orc.synth_image(base_rows, w, ch, dt, seed)
    out = np.empty((h, w, ch), dtype=base.dtype)
    for i, y in enumerate(range(0, h, base_rows)):
        n = min(base_rows, h - y)
        out[y:y + n] = np.roll(base[:n], 7 * i, axis=1)
    return out


def cpu_sample_rows(h, w, cores, mp_per_core_s=5.0, seconds=12.0):
    """Rows of the image to time so that the run costs about `seconds` of wall time."""
    rows = int(seconds * cores * mp_per_core_s * 1e6 / w)
    return max(cores, min(h, rows))


_cpu_dst = {}


def time_cpu(chk, orc, img, p, rows, cores, reps):
    h = img.shape[0]
    y1 = (h - rows) // 2
    dst = _cpu_dst.get(id(img))
    if dst is None:
        dst = _cpu_dst[id(img)] = np.empty_like(img)
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        chk.region(img, p, y1, y1 + rows, dst=dst, threads=cores)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the same pass on this host."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    chk, orc = cpu_checker()
    W, Hr, ch, dts, interp, kw, lens = WORKLOADS[args.workload]
    dt = np.dtype(dts)
    n = max(1, args.gpus)
    H = Hr * n
    lx, ly = (W // 2, H // 2) if lens == "centre" else lens
    p = orc.Params(interpolation=interp, lens_x=float(lx), lens_y=float(ly), **kw)
    cores = os.cpu_count() or 1
    # the sample is rows of rank 0's band; the image buffer only needs those rows + halo, but the
    # reference takes the whole image pointer, so allocate one band (the N-rank image's top band
    # has the same per-pixel cost)
    img = host_image(orc, Hr, W, ch, dts, seed=4)
    p_band = orc.Params(interpolation=interp, lens_x=float(lx), lens_y=float(Hr // 2 if lens == "centre" else ly), **kw)
    budget = max(2.0, min(12.0, 150.0 / max(1, args.steps + args.warmup)))
    rows = cpu_sample_rows(Hr, W, cores, seconds=budget)
    for _ in range(args.warmup):
        time_cpu(chk, orc, img, p_band, rows, cores, 1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        time_cpu(chk, orc, img, p_band, rows, cores, 1)
    per_step = (time.perf_counter() - t0) / max(1, args.steps)
    mps = rows * W / 1e6 / per_step
    sample = "%d of %d rows of one %dx%d band per step, %d threads (one row sub-band each)" % (rows, Hr, W, Hr, cores)
    line = {
        "impl": "reference", "metric": "megapixels/sec (cubic, lateral+directional)" if interp == 2 else "megapixels/sec",
        "value": round(mps, 3), "unit": "MP/s", "n_gpus": n, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(per_step * 1e3, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload, n, "host cores; %s" % sample),
        "cpu_baseline": {"value": round(mps, 3), "unit": "MP/s", "cores": cores, "kind": chk.kind, "sample": sample},
        "e2e": {"value": round(mps, 3), "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=args.out, flush=True)
    return 0


def workload_config(name, n, extra=None):
    W, Hr, ch, dts, interp, kw, lens = WORKLOADS[name]
    c = {"workload": name, "width": W, "rows_per_gpu": Hr, "image": "%dx%d" % (W, Hr * n), "channels": ch,
         "sample": {"u1": "u8", "u2": "u16", "f4": "f32"}[dts], "interpolation": INTERP_NAME[interp],
         "params": kw, "lens": lens, "parallelism": "row bands x%d (+halo rows), no collective" % n,
         "cache": "inputs larger than L2 (band %.0f MB in + %.0f MB out vs 126 MB L2)"
                  % (W * Hr * ch * np.dtype(dts).itemsize / 1e6, W * Hr * ch * np.dtype(dts).itemsize / 1e6)}
    if extra:
        c["note"] = extra
    return c


# ---------------------------------------------------------------------------------------------
# the CUDA path
# ---------------------------------------------------------------------------------------------
def run_cuda_batch(args, torch, dist, fixca, rank, world, local, dev, barrier, max_over_ranks):
    """Frame batches (BASELINE configs[4]): every rank holds `frames` device-resident frames; a step is one launch
    over all of them.  e2e: the host-frame stream API (pinned H2D / kernel / D2H ring) on a few frames."""
    import ctypes
    W, H, ch, dts, interp, kw, lens = WORKLOADS[args.workload]
    F = WORKLOAD_FRAMES[args.workload]
    dt = np.dtype(dts)
    bpp, bpc = ch * dt.itemsize, bpc_of(dt)
    lx, ly = (W // 2, H // 2) if lens == "centre" else lens
    p = fixca.FixCaParams(interpolation=interp, lens_x=float(lx), lens_y=float(ly), **kw)
    flags = fixca.PRECISION_EXACT if args.exact else fixca.PRECISION_FAST
    row_bytes = W * bpp
    pitch = (row_bytes + 127) // 128 * 128
    g = torch.Generator(device=dev)
    g.manual_seed(1000 + rank)
    d_src = torch.randint(0, 256, (F, H, pitch), dtype=torch.uint8, device=dev, generator=g)
    d_dst = torch.empty_like(d_src)
    stream = torch.cuda.current_stream()

    def step():
        fixca.fix_ca_frames_dev(d_src.data_ptr(), pitch, pitch * H, d_dst.data_ptr(), pitch, pitch * H, F, W, H, bpp, bpc, p,
                                flags, stream.cuda_stream)

    sampler = ClockSampler(_nvml_index(local))
    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    n0 = fixca.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    sampler.stop()
    launches = fixca.launch_count() - n0
    kernel = fixca.last_kernel()
    ms_step = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    total_mp_step = W * H * F * world / 1e6
    local_ms = e0.elapsed_time(e1) / args.steps
    alg_bytes = 2.0 * bpp * W * H * F
    achieved = alg_bytes / (local_ms * 1e-3) / 1e9
    peak, peak_src = measured_peak()

    e2e = parity = None
    if not args.no_e2e:
        nf = 16
        h_src = torch.empty((nf, H, row_bytes), dtype=torch.uint8, pin_memory=True)
        h_dst = torch.empty((nf, H, row_bytes), dtype=torch.uint8, pin_memory=True)
        h_src.copy_(d_src[:nf, :, :row_bytes])
        torch.cuda.synchronize()
        vp = ctypes.c_void_p
        srcs = (vp * nf)(*[h_src[k].data_ptr() for k in range(nf)])
        dsts = (vp * nf)(*[h_dst[k].data_ptr() for k in range(nf)])
        L = fixca.load()

        def e2e_step():
            rc = L.fixca_cuda_frames(srcs, dsts, nf, W, H, bpp, bpc, ctypes.byref(p), flags, local)
            if rc:
                raise RuntimeError("fixca_cuda_frames: %d" % rc)

        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        dt_e2e = max_over_ranks(time.perf_counter() - t0)
        barrier()
        e2e = {"value": round(W * H * nf * world * e2e_steps / 1e6 / dt_e2e, 1), "unit": "MP/s",
               "h2d_bytes_per_step": int(nf * H * row_bytes) * world, "d2h_bytes_per_step": int(nf * H * row_bytes) * world,
               "steps": e2e_steps, "ms_per_step": round(dt_e2e / e2e_steps * 1e3, 3),
               "api": "fixca_cuda_frames (%d pinned host frames per step, H2D / kernel / D2H ring), synchronous" % nf}
        if rank == 0 and not args.no_check:
            parity = spot_check(args, h_src[0], h_dst[0], W, H, ch, dt, p, 0, 0, kw, interp, lx, ly)
    if rank == 0:
        cfg = workload_config(args.workload, 1, "arithmetic %s; kernel %s" % (
            "exact FP64 (bit-identical)" if args.exact else "fast FP32 (+-1 LSB of the reference)", kernel))
        cfg.update({"frames_per_gpu": F, "image": "%dx%d x %d frames" % (W, H, F * world),
                    "parallelism": "frames sharded by index over %d GPU(s), no collective" % world,
                    "cache": "inputs larger than L2 (%.0f MB in + %.0f MB out per GPU vs 126 MB L2)"
                             % (F * H * row_bytes / 1e6, F * H * row_bytes / 1e6)})
        cfg.pop("rows_per_gpu", None)
        line = {
            "metric": "megapixels/sec (cubic, lateral+directional)" if interp == 2 else "megapixels/sec",
            "value": round(total_mp_step / (ms_step * 1e-3), 1), "unit": "MP/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": round(ms_step, 5), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64" if args.exact else "f32", "data": "synthetic",
            "config": cfg,
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "traffic": ncu_traffic(args.workload, kernel),
                         "peak_source": peak_src, "kernel": kernel,
                         "algorithmic_bytes_per_launch": int(alg_bytes), "launch_ms": round(local_ms, 5)},
            "cpu_baseline": None, "e2e": e2e, "gpu_launches": int(launches), "clocks": sampler.report(),
        }
        if parity is not None:
            line["parity"] = parity
        print(json.dumps(line), file=args.out, flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_cuda(args):
    import torch
    import torch.distributed as dist

    import fixca

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != max(1, args.gpus) and rank == 0:
        print("bench.py: --gpus %d but WORLD_SIZE %d; using WORLD_SIZE" % (args.gpus, world), file=sys.stderr)
    if not torch.cuda.is_available() or fixca.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _bind_to_gpu_numa_node(_nvml_index(local))     # pinned host buffers should be local to the GPU's root port
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if args.workload in WORKLOAD_FRAMES:
        return run_cuda_batch(args, torch, dist, fixca, rank, world, local, dev, barrier, max_over_ranks)
    W, Hr, ch, dts, interp, kw, lens = WORKLOADS[args.workload]
    dt = np.dtype(dts)
    bpp = ch * dt.itemsize
    bpc = bpc_of(dt)
    H = Hr * world
    lx, ly = (W // 2, H // 2) if lens == "centre" else lens
    p = fixca.FixCaParams(interpolation=interp, lens_x=float(lx), lens_y=float(ly), **kw)
    flags = fixca.PRECISION_EXACT if args.exact else fixca.PRECISION_FAST
    from fixca import bands
    plan = bands.plan_band(W, H, p, rank, world)     # this rank's rows + the halo rows it must hold
    y1, y2, lo, hi, src_rows = plan.y1, plan.y2, plan.src_lo, plan.src_hi, plan.src_rows
    row_bytes = W * bpp
    pitch = (row_bytes + 127) // 128 * 128

    # ---- device-resident band (+halo), synthetic noise ----
    g = torch.Generator(device=dev)
    g.manual_seed(4 + rank)
    if dt.kind == "f":
        d_src = torch.rand((src_rows, pitch // 4), dtype=torch.float32, device=dev, generator=g).view(torch.uint8)
    else:
        d_src = torch.randint(0, 256, (src_rows, pitch), dtype=torch.uint8, device=dev, generator=g)
    d_dst = torch.empty((y2 - y1, pitch), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()

    def step():
        bands.run_band_device(plan, d_src.data_ptr(), pitch, d_dst.data_ptr(), pitch, bpp, bpc, p, flags, stream.cuda_stream)

    sampler = ClockSampler(_nvml_index(local))
    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    n0 = fixca.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    sampler.stop()
    launches = fixca.launch_count() - n0
    kernel = fixca.last_kernel()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_step = ms_total / args.steps
    mp_step = W * (y2 - y1) / 1e6                     # this rank's pixels (equal bands)
    total_mp_step = W * H / 1e6
    value = total_mp_step / (ms_step * 1e-3)
    # roofline of the (single) kernel each step launches: local mean launch time
    local_ms = e0.elapsed_time(e1) / args.steps
    alg_bytes = 2.0 * bpp * W * (y2 - y1)
    achieved = alg_bytes / (local_ms * 1e-3) / 1e9
    peak, peak_src = measured_peak()

    # ---- optional: reassembly of the bands on rank 0 (NCCL send/recv over NVLink), outside the timed steps.
    # Bands are independent, so this is the only inter-GPU traffic the pass can have; it is
    # reported beside the step time to show why it is kept off the measured path.
    gather = None
    if args.gather and world > 1:
        times = []
        for rep in range(4):
            barrier()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record(stream)
            full = bands.gather_bands(d_dst, plan, dst_rank=0)
            g1.record(stream)
            barrier()
            if rep:
                times.append(max_over_ranks(g0.elapsed_time(g1)))
            del full
        gather = {"ms": round(min(times), 3), "bytes_into_rank0": int((world - 1) * (y2 - y1) * pitch),
                  "how": "fixca.bands.gather_bands: one NCCL send/recv per band to rank 0, best of 3"}
        # The same reassembly folded into the pass: rank 0 owns the whole frame, every rank's kernel stores its
        # finished chunks straight into it (TMA stores through the CUDA IPC peer mapping, NVLink): compute + gather
        # in one kernel, timed like a step (barrier both sides, max over ranks).
        frame = bands.PeerFrame(H, pitch, owner=0)

        def peer_step():
            bands.run_band_into_frame(plan, d_src.data_ptr(), pitch, frame, bpp, bpc, p, flags, stream.cuda_stream)

        for _ in range(3):
            peer_step()
        ptimes = []
        for rep in range(6):
            barrier()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record(stream)
            peer_step()
            g1.record(stream)
            barrier()
            ptimes.append(max_over_ranks(g0.elapsed_time(g1)))
        peer_kernel = fixca.last_kernel()
        step()
        full = bands.gather_bands(d_dst, plan, dst_rank=0)
        frame.sync()
        same = None
        if rank == 0:
            same = bool(torch.equal(frame.as_tensor()[:, :row_bytes], full[:, :row_bytes]))
        del full
        frame.close()
        gather["peer_store"] = {"ms": round(min(ptimes), 3), "median_ms": round(sorted(ptimes)[len(ptimes) // 2], 3),
                                "identical_to_nccl_gather": same, "kernel": peer_kernel,
                                "how": "fixca.bands.PeerFrame + run_band_into_frame: every rank's kernel writes its band "
                                       "into rank 0's frame over NVLink (compute + gather, one kernel); max over ranks"}

    # ---- end to end through the host C ABI: pinned host band, H2D + kernels + D2H timed ----
    e2e = None
    parity = None
    if not args.no_e2e:
        h_src = torch.empty((src_rows, row_bytes), dtype=torch.uint8, pin_memory=True)
        h_dst = torch.empty((y2 - y1, row_bytes), dtype=torch.uint8, pin_memory=True)
        h_src.copy_(d_src[:, :row_bytes])
        torch.cuda.synchronize()
        # the ABI addresses whole-image buffers; only rows [lo,hi] of src and [y1,y2) of dst are touched
        src_base = h_src.data_ptr() - lo * row_bytes
        dst_base = h_dst.data_ptr() - y1 * row_bytes

        def e2e_step():
            fixca.fix_ca_region(src_base, dst_base, W, H, bpp, bpc, p, 0, W, y1, y2, True, flags, local)

        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()          # synchronous: returns when dst holds the rows
        torch.cuda.synchronize()
        dt_e2e = max_over_ranks(time.perf_counter() - t0)
        barrier()
        # the floor of any end-to-end number on this host: the same bytes moved by plain copies, upload and
        # download concurrently on two streams, no kernel
        s_up, s_down = torch.cuda.Stream(), torch.cuda.Stream()
        d_in = torch.empty((src_rows, row_bytes), dtype=torch.uint8, device=dev)
        d_out = torch.empty((y2 - y1, row_bytes), dtype=torch.uint8, device=dev)

        def copy_only():
            with torch.cuda.stream(s_up):
                d_in.copy_(h_src, non_blocking=True)
            with torch.cuda.stream(s_down):
                h_dst.copy_(d_out, non_blocking=True)
            s_up.synchronize()
            s_down.synchronize()

        dst_keep = h_dst.clone()
        copy_only()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            copy_only()
        dt_copy = max_over_ranks(time.perf_counter() - t0)
        h_dst.copy_(dst_keep)
        del d_in, d_out, dst_keep
        barrier()
        # the call exactly as the plug-in makes it: pageable buffers (g_new, fix-ca.c:366-367), staged through
        # the library's pinned rings by its copy threads (rank 0 of a single-GPU run only; reported, not the e2e value)
        pageable_ms = None
        if world == 1:
            p_src = np.empty((src_rows, row_bytes), dtype=np.uint8)
            p_dst = np.empty((y2 - y1, row_bytes), dtype=np.uint8)
            p_src[:] = h_src.numpy()
            ps, pd = p_src.ctypes.data - lo * row_bytes, p_dst.ctypes.data - y1 * row_bytes
            fixca.fix_ca_region(ps, pd, W, H, bpp, bpc, p, 0, W, y1, y2, True, flags, local)
            t0 = time.perf_counter()
            for _ in range(2):
                fixca.fix_ca_region(ps, pd, W, H, bpp, bpc, p, 0, W, y1, y2, True, flags, local)
            pageable_ms = round((time.perf_counter() - t0) / 2 * 1e3, 3)
            del p_src, p_dst
        e2e = {"value": round(total_mp_step * e2e_steps / dt_e2e, 1), "unit": "MP/s",
               "copy_only_ms_per_step": round(dt_copy / e2e_steps * 1e3, 3),
               "pageable_caller_ms_per_step": pageable_ms,
               "h2d_bytes_per_step": int(src_rows * row_bytes) * world, "d2h_bytes_per_step": int((y2 - y1) * row_bytes) * world,
               "steps": e2e_steps, "ms_per_step": round(dt_e2e / e2e_steps * 1e3, 3),
               "api": "fixca_cuda_region_ex (host pointers, pinned), synchronous"}
        if rank == 0 and not args.no_check:
            parity = spot_check(args, h_src, h_dst, W, H, ch, dt, p, lo, y1, kw, interp, lx, ly)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline(args)

    if rank == 0:
        line = {
            "metric": "megapixels/sec (cubic, lateral+directional)" if interp == 2 else "megapixels/sec",
            "value": round(value, 1), "unit": "MP/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": round(ms_step, 5), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "f64" if args.exact or dts not in ("u1", "u2", "f4") else "f32",
            "data": "synthetic",
            "config": workload_config(args.workload, world,
                                      "arithmetic %s; kernel %s" % ("exact FP64 (bit-identical)" if args.exact else
                                                                    "fast FP32 (+-1 LSB of the reference)", kernel)),
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "traffic": ncu_traffic(args.workload, kernel),
                         "peak_source": peak_src, "kernel": kernel,
                         "algorithmic_bytes_per_launch": int(alg_bytes), "launch_ms": round(local_ms, 5)},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": sampler.report(),
        }
        if parity is not None:
            line["parity"] = parity
        if gather is not None:
            line["gather"] = gather
        print(json.dumps(line), file=args.out, flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def _bind_to_gpu_numa_node(index: int) -> None:
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(index))
    except Exception:
        pass


def _nvml_index(local: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local])
        except Exception:
            pass
    return local


def spot_check(args, h_src, h_dst, W, H, ch, dt, p, lo, y1, kw, interp, lx, ly):
    """The first rows of rank 0's e2e output against the reference's own code (oracle = checker).
    Rank 0 holds the top band, so its host buffer starts at image row 0; the checker is given an
    H-row view of it and only ever reads the rows output rows [0, rows) need (band independence,
    SURVEY.md 8(e))."""
    try:
        import fixca
        chk, orc = cpu_checker()
        rows = 6
        if lo != 0 or y1 != 0:
            return None
        need_lo, need_hi = fixca.band_source_rows(W, H, p, 0, rows)
        src = h_src.numpy().view(dt).reshape(h_src.shape[0], W, ch)
        got = h_dst.numpy().view(dt).reshape(h_dst.shape[0], W, ch)[:rows]
        if need_hi >= src.shape[0]:
            return None
        want = np.zeros((rows, W, ch), dtype=dt)
        as_strided = np.lib.stride_tricks.as_strided
        P = orc.Params(interpolation=interp, lens_x=float(lx), lens_y=float(ly), **kw)
        chk.region(as_strided(src, shape=(H, W, ch), strides=src.strides), P, 0, rows,
                   dst=as_strided(want, shape=(H, W, ch), strides=want.strides))
        if dt.kind == "f":
            d = float(np.abs(want.astype(np.float64) - got.astype(np.float64)).max())
        else:
            d = int(np.abs(want.astype(np.int64) - got.astype(np.int64)).max())
        return {"checked_rows": rows, "max_abs_diff": d, "checker": chk.kind,
                "tolerance": 0 if args.exact or interp == 0 else (1 if dt.kind != "f" else 1e-6)}
    except Exception as e:  # the check is advisory; never hide the bench line
        return {"error": repr(e)}


def cpu_baseline(args):
    chk, orc = cpu_checker()
    W, Hr, ch, dts, interp, kw, lens = WORKLOADS[args.workload]
    lx, ly = (W // 2, Hr // 2) if lens == "centre" else lens
    P = orc.Params(interpolation=interp, lens_x=float(lx), lens_y=float(ly), **kw)
    cores = os.cpu_count() or 1
    img = host_image(orc, Hr, W, ch, dts, seed=4)
    rows = cpu_sample_rows(Hr, W, cores, seconds=8.0)
    best = time_cpu(chk, orc, img, P, rows, cores, reps=2)
    one = None
    if not args.quick_cpu:
        rows1 = max(8, min(Hr, int(3.0 * 5.0e6 / W)))
        one = rows1 * W / 1e6 / time_cpu(chk, orc, img, P, rows1, 1, reps=1)
    return {"value": round(rows * W / 1e6 / best, 2), "unit": "MP/s", "cores": cores, "kind": chk.kind,
            "sample": "%d of %d rows of the %dx%d image, best of 2, %d threads (one row sub-band each)" % (rows, Hr, W, Hr, cores),
            "single_thread_value": None if one is None else round(one, 2)}


def _stdout_to_stderr():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version
    banner with NCCL_DEBUG=VERSION), so everything but the final line is sent to stderr: fd 1 is pointed at
    fd 2 for the run and the saved descriptor is used for the result."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(saved, "w")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", choices=("cuda", "reference"), default="cuda")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default=DEFAULT_WORKLOAD)
    ap.add_argument("--exact", action="store_true", help="FP64 bit-exact arithmetic instead of fast FP32")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--gather", action="store_true", help="also time the reassembly of the bands on rank 0 (N > 1)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--quick-cpu", action="store_true")
    args = ap.parse_args()
    args.steps = max(1, args.steps)
    args.out = _stdout_to_stderr()
    if args.impl == "reference":
        return run_reference(args)
    return run_cuda(args)


if __name__ == "__main__":
    sys.exit(main())
