#!/usr/bin/env python
"""Developer timing probe (not the contract bench): device-resident kernel time for a few
workloads and tile-height settings, CUDA events on the launching stream."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gimp-fix-ca_b200"))
import torch  # noqa: E402

import fixca  # noqa: E402

KW = dict(blue=3.0, red=-2.0, x_blue=0.7, x_red=-0.4, y_blue=0.3, y_red=-0.9)
PEAK = 6458.7


def run(name, h, w, ch, tdtype, bpc, interp, flags, reps=10, lens=None):
    es = torch.empty((), dtype=tdtype).element_size()
    bpp = ch * es
    pitch = (w * bpp + 127) // 128 * 128
    fixca.reload_tuning()       # callers change FIXCA_* between runs; the library reads them once
    # rotating buffer sets (bench.py's rule): no launch finds its input in the 126 MB L2
    nsets = 1 if h * pitch >= 3 * 126e6 else min(8, 1 + int(-(-2 * 126e6 // (2 * h * pitch))))
    srcs = [torch.randint(0, 255, (h, pitch), dtype=torch.uint8, device="cuda") for _ in range(nsets)]
    dsts = [torch.empty_like(srcs[0]) for _ in range(nsets)]
    lx, ly = lens if lens else (w // 2, h // 2)
    p = fixca.FixCaParams(interpolation=interp, lens_x=lx, lens_y=ly, **KW)
    st = torch.cuda.current_stream().cuda_stream
    turn = [0]

    def call():
        i = turn[0] % nsets
        turn[0] += 1
        fixca.fix_ca_region_dev(srcs[i].data_ptr(), pitch, 0, h, dsts[i].data_ptr(), pitch, 0, w, h, bpp, bpc, p, 0, h, flags, st)
    for _ in range(3 + nsets):
        call()
    torch.cuda.synchronize()
    ts = []
    inner = 10      # launches per event pair: host-side planning/launch latency must not be timed as GPU time
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _i in range(inner):
            call()
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / inner)
    ts.sort()
    best, med = ts[0], ts[len(ts) // 2]
    mp = h * w / 1e6
    gbs = mp * 1e6 * 2 * bpp / (med * 1e-3) / 1e9
    print("%-34s %-28s th=%-3s best %.3f ms  med %.3f ms  %9.0f MP/s  %7.0f GB/s  %.1f%% of %.0f"
          % (name, fixca.last_kernel(), os.environ.get("FIXCA_TILE_H", "auto"), best, med, mp / (med * 1e-3), gbs, 100 * gbs / PEAK, PEAK), flush=True)


def run_batch(name, nf, h, w, ch, tdtype, bpc, interp, flags, reps=5, only_batch=False):
    es = torch.empty((), dtype=tdtype).element_size()
    bpp = ch * es
    pitch = (w * bpp + 127) // 128 * 128
    src = torch.randint(0, 255, (nf, h, pitch), dtype=torch.uint8, device="cuda")
    dst = torch.empty_like(src)
    fixca.reload_tuning()
    p = fixca.FixCaParams(interpolation=interp, lens_x=w // 2, lens_y=h // 2, **KW)
    st = torch.cuda.current_stream().cuda_stream
    batch = lambda: fixca.fix_ca_frames_dev(src.data_ptr(), pitch, pitch * h, dst.data_ptr(), pitch, pitch * h, nf, w, h, bpp, bpc, p, flags, st)

    def loop():
        for k in range(nf):
            fixca.fix_ca_region_dev(src[k].data_ptr(), pitch, 0, h, dst[k].data_ptr(), pitch, 0, w, h, bpp, bpc, p, 0, h, flags, st)
    for label, call in (("one launch", batch), ("per frame", loop))[:1 if only_batch else 2]:
        for _ in range(2):
            call()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); call(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        med = ts[len(ts) // 2]
        mp = nf * h * w / 1e6
        gbs = mp * 1e6 * 2 * bpp / (med * 1e-3) / 1e9
        print("%-30s %-12s %-26s %d frames  med %.3f ms = %.4f ms/frame  %9.0f MP/s  %6.0f GB/s  %.1f%% of %.0f"
              % (name, label, fixca.last_kernel(), nf, med, med / nf, mp / (med * 1e-3), gbs, 100 * gbs / PEAK, PEAK), flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    F, E = fixca.PRECISION_FAST, fixca.PRECISION_EXACT
    if which in ("all", "target"):
        run("100MP rgb16 cubic fast", 8192, 12288, 3, torch.int16, 2, 2, F)
        run("100MP rgb16 cubic exact", 8192, 12288, 3, torch.int16, 2, 2, E, reps=5)
        run("100MP rgb16 linear fast", 8192, 12288, 3, torch.int16, 2, 1, F)
        run("100MP rgb16 none", 8192, 12288, 3, torch.int16, 2, 0, E)
        run("100MP rgb16 cubic fast direct", 8192, 12288, 3, torch.int16, 2, 2, F | fixca.FORCE_DIRECT, reps=5)
    if which in ("all", "cfgs"):
        run("24MP rgb8 linear fast", 4000, 6000, 3, torch.uint8, 1, 1, F)
        run("24MP rgb8 linear exact", 4000, 6000, 3, torch.uint8, 1, 1, E)
        run("24MP rgb8 cubic fast", 4000, 6000, 3, torch.uint8, 1, 2, F)
        run("8K rgba16 cubic fast", 4320, 7680, 4, torch.int16, 2, 2, F, lens=(658, 1280))
        run("8K rgba16 cubic exact", 4320, 7680, 4, torch.int16, 2, 2, E, lens=(658, 1280))
        run("50MP rgb f32 cubic fast", 6144, 8192, 3, torch.float32, -4, 2, F)
        run("50MP rgb f32 cubic exact", 6144, 8192, 3, torch.float32, -4, 2, E)
        run("4K rgb8 cubic fast", 2160, 3840, 3, torch.uint8, 1, 2, F)
    if which == "quick":
        for th in sys.argv[2:] or ["16"]:
            os.environ["FIXCA_TILE_H"] = th
            run("100MP rgb16 cubic fast", 8192, 12288, 3, torch.int16, 2, 2, F)
            run("100MP rgb16 linear fast", 8192, 12288, 3, torch.int16, 2, 1, F)
            run("24MP rgb8 cubic fast", 4000, 6000, 3, torch.uint8, 1, 2, F)
            run("24MP rgb8 linear fast", 4000, 6000, 3, torch.uint8, 1, 1, F)
            run("8K rgba16 cubic fast", 4320, 7680, 4, torch.int16, 2, 2, F, lens=(658, 1280))
            run("50MP rgb f32 cubic fast", 6144, 8192, 3, torch.float32, -4, 2, F)
    if which == "mix":
        run("100MP rgb16 cubic fast", 8192, 12288, 3, torch.int16, 2, 2, F)
        run("100MP rgb16 linear fast", 8192, 12288, 3, torch.int16, 2, 1, F)
        run("24MP rgb8 cubic fast", 4000, 6000, 3, torch.uint8, 1, 2, F)
        run("24MP rgb8 linear fast", 4000, 6000, 3, torch.uint8, 1, 1, F)
        run("4K rgb8 cubic fast", 2160, 3840, 3, torch.uint8, 1, 2, F)
        run("8K rgba16 cubic fast", 4320, 7680, 4, torch.int16, 2, 2, F, lens=(658, 1280))
        run("8K rgba16 linear fast", 4320, 7680, 4, torch.int16, 2, 1, F, lens=(658, 1280))
        run("33MP rgba8 cubic fast", 4320, 7680, 4, torch.uint8, 1, 2, F)
        run("50MP rgb f32 cubic fast", 6144, 8192, 3, torch.float32, -4, 2, F)
        run("50MP rgba f32 cubic fast", 6144, 8192, 4, torch.float32, -4, 2, F)
    if which == "ab":       # the A/B set of the r02 kernel work: RGB8 first, then one of every other layout
        run_batch("4K rgb8 cubic", 64, 2160, 3840, 3, torch.uint8, 1, 2, F)
        run("4K rgb8 cubic fast", 2160, 3840, 3, torch.uint8, 1, 2, F)
        run("24MP rgb8 cubic fast", 4000, 6000, 3, torch.uint8, 1, 2, F)
        run("24MP rgb8 linear fast", 4000, 6000, 3, torch.uint8, 1, 1, F)
        run("24MP rgb8 none", 4000, 6000, 3, torch.uint8, 1, 0, E)
        run("33MP rgba8 cubic fast", 4320, 7680, 4, torch.uint8, 1, 2, F)
        run("100MP rgb16 cubic fast", 8192, 12288, 3, torch.int16, 2, 2, F)
        run("100MP rgb16 linear fast", 8192, 12288, 3, torch.int16, 2, 1, F)
        run("100MP rgb16 none", 8192, 12288, 3, torch.int16, 2, 0, E)
        run("8K rgba16 cubic fast", 4320, 7680, 4, torch.int16, 2, 2, F, lens=(658, 1280))
        run("50MP rgb f32 cubic fast", 6144, 8192, 3, torch.float32, -4, 2, F)
        run("50MP rgba f32 cubic fast", 6144, 8192, 4, torch.float32, -4, 2, F)
    if which == "tlead":    # pass-through tile requested 1 / 2 chunks ahead
        for tl in ("1", "2"):
            os.environ["FIXCA_STREAM_TLEAD"] = tl
            run_batch("4K rgb8 cubic tlead" + tl, 64, 2160, 3840, 3, torch.uint8, 1, 2, F)
            run("4K rgb8 cubic tlead" + tl, 2160, 3840, 3, torch.uint8, 1, 2, F)
            run("24MP rgb8 cubic tlead" + tl, 4000, 6000, 3, torch.uint8, 1, 2, F)
            run("24MP rgb8 linear tlead" + tl, 4000, 6000, 3, torch.uint8, 1, 1, F)
            run("33MP rgba8 cubic tlead" + tl, 4320, 7680, 4, torch.uint8, 1, 2, F)
            run("100MP rgb16 cubic tlead" + tl, 8192, 12288, 3, torch.int16, 2, 2, F)
            run("8K rgba16 cubic tlead" + tl, 4320, 7680, 4, torch.int16, 2, 2, F, lens=(658, 1280))
            run("50MP rgb f32 cubic tlead" + tl, 6144, 8192, 3, torch.float32, -4, 2, F)
            run("50MP rgba f32 cubic tlead" + tl, 6144, 8192, 4, torch.float32, -4, 2, F)
        os.environ.pop("FIXCA_STREAM_TLEAD")
    if which == "u8":       # the 8-bit layouts only (A/B runs of two builds: FIXCA_LIB)
        run_batch("4K rgb8 cubic", 64, 2160, 3840, 3, torch.uint8, 1, 2, F)
        run("4K rgb8 cubic fast", 2160, 3840, 3, torch.uint8, 1, 2, F)
        run("24MP rgb8 cubic fast", 4000, 6000, 3, torch.uint8, 1, 2, F)
        run("24MP rgb8 linear fast", 4000, 6000, 3, torch.uint8, 1, 1, F)
        run("33MP rgba8 cubic fast", 4320, 7680, 4, torch.uint8, 1, 2, F)
        run("24MP rgb8 cubic exact", 4000, 6000, 3, torch.uint8, 1, 2, E)
    if which == "split":    # FIXCA_LIB = the TUNING build: memory pipeline only (1), compute only (4: no loads, 12: no loads, no stores)
        for dbg in ("0", "1", "4", "12"):
            os.environ["FIXCA_STREAM_DEBUG"] = dbg
            run_batch("4K rgb8 cubic dbg" + dbg, 64, 2160, 3840, 3, torch.uint8, 1, 2, F)
            run("24MP rgb8 cubic dbg" + dbg, 4000, 6000, 3, torch.uint8, 1, 2, F)
            run("24MP rgb8 linear dbg" + dbg, 4000, 6000, 3, torch.uint8, 1, 1, F)
            run("24MP rgb8 none dbg" + dbg, 4000, 6000, 3, torch.uint8, 1, 0, E)
            run("33MP rgba8 cubic dbg" + dbg, 4320, 7680, 4, torch.uint8, 1, 2, F)
            run("100MP rgb16 cubic dbg" + dbg, 8192, 12288, 3, torch.int16, 2, 2, F)
        os.environ.pop("FIXCA_STREAM_DEBUG")
    if which == "ctas5":    # FIXCA_LIB = a build with 5-CTA launch bounds: 5 CTAs per SM at depth 1 against the default
        for env in ({}, {"FIXCA_STREAM_CTAS": "5", "FIXCA_STREAM_DEPTH": "1"}, {"FIXCA_STREAM_CTAS": "4", "FIXCA_STREAM_DEPTH": "1"}):
            for k in ("FIXCA_STREAM_CTAS", "FIXCA_STREAM_DEPTH"):
                os.environ.pop(k, None)
            os.environ.update(env)
            os.environ["FIXCA_VERBOSE"] = "1"
            tag = " ".join("%s=%s" % (k[13:], v) for k, v in env.items()) or "default"
            run_batch("4K rgb8 cubic " + tag, 64, 2160, 3840, 3, torch.uint8, 1, 2, F)
            os.environ["FIXCA_VERBOSE"] = "0"
            run("24MP rgb8 cubic " + tag, 4000, 6000, 3, torch.uint8, 1, 2, F)
            run("24MP rgb8 linear " + tag, 4000, 6000, 3, torch.uint8, 1, 1, F)
            run("33MP rgba8 cubic " + tag, 4320, 7680, 4, torch.uint8, 1, 2, F)
    if which == "segs":     # batch launches: rows per CTA (FIXCA_STREAM_SEGS segments per frame)
        for nf in (128, 64):
            for segs in ("0", "1", "2", "3", "4", "6", "9"):
                os.environ["FIXCA_STREAM_SEGS"] = segs
                run_batch("4K rgb8 cubic segs" + segs, nf, 2160, 3840, 3, torch.uint8, 1, 2, F, reps=5, only_batch=True)
        os.environ.pop("FIXCA_STREAM_SEGS")
    if which == "depth":    # pipeline depth sweep of the narrow-strip layouts
        for d in ("0", "2", "3", "4", "6"):
            os.environ["FIXCA_STREAM_DEPTH"] = d
            os.environ["FIXCA_VERBOSE"] = "0"
            for name, args in (("24MP rgb8 none", (4000, 6000, 3, torch.uint8, 1, 0, E)), ("24MP rgb8 linear", (4000, 6000, 3, torch.uint8, 1, 1, F)),
                               ("24MP rgb8 cubic", (4000, 6000, 3, torch.uint8, 1, 2, F)), ("33MP rgba8 cubic", (4320, 7680, 4, torch.uint8, 1, 2, F)),
                               ("33MP rgba8 none", (4320, 7680, 4, torch.uint8, 1, 0, E)),
                               ("50MP rgba f32 cubic", (6144, 8192, 4, torch.float32, -4, 2, F)), ("50MP rgba f32 none", (6144, 8192, 4, torch.float32, -4, 0, E))):
                try:
                    run(name + " D" + d, *args)
                except fixca.FixCaError as e:
                    print(name, "D", d, str(e)[:80])
        os.environ.pop("FIXCA_STREAM_DEPTH")
    if which == "exact8":   # bit-identical 8-bit: the deferred form (repair_patch_kernel behind the streaming kernel) against the in-kernel queues
        for form in ("", "inline"):
            os.environ["FIXCA_EXACT_KERNEL"] = form
            tag = " exact " + (form or "deferred")
            run("24MP rgb8 cubic" + tag, 4000, 6000, 3, torch.uint8, 1, 2, E)
            run("24MP rgb8 linear" + tag, 4000, 6000, 3, torch.uint8, 1, 1, E)
            run("4K rgb8 cubic" + tag, 2160, 3840, 3, torch.uint8, 1, 2, E)
            run("33MP rgba8 cubic" + tag, 4320, 7680, 4, torch.uint8, 1, 2, E)
            run_batch("4K rgb8 cubic" + tag, 32, 2160, 3840, 3, torch.uint8, 1, 2, E, only_batch=True)
        os.environ.pop("FIXCA_EXACT_KERNEL")
        run("24MP rgb8 cubic fast", 4000, 6000, 3, torch.uint8, 1, 2, F)
        run("24MP rgb8 linear fast", 4000, 6000, 3, torch.uint8, 1, 1, F)
        run("4K rgb8 cubic fast", 2160, 3840, 3, torch.uint8, 1, 2, F)
    if which == "small":    # launches dominated by the fixed cost
        run("4K rgb8 cubic fast", 2160, 3840, 3, torch.uint8, 1, 2, F)
        run("4K rgb8 linear fast", 2160, 3840, 3, torch.uint8, 1, 1, F)
        run("24MP rgb8 linear fast", 4000, 6000, 3, torch.uint8, 1, 1, F)
        run("768-row rgb f32 cubic band", 768, 8192, 3, torch.float32, -4, 2, F)
        run("1536-row rgb f32 cubic band", 1536, 8192, 3, torch.float32, -4, 2, F)
        run("1080p rgba8 cubic", 1080, 1920, 4, torch.uint8, 1, 2, F)
        run("100MP rgb16 cubic fast", 8192, 12288, 3, torch.int16, 2, 2, F)
    if which == "naps":     # FIXCA_LIB = the TUNING build: helper warps that nap between barrier tests (debug >> 8 = ns)
        for ns in (0, 64, 128, 256, 512, 1024):
            os.environ["FIXCA_STREAM_DEBUG"] = str(ns << 8)
            run_batch("4K rgb8 cubic naps %d" % ns, 64, 2160, 3840, 3, torch.uint8, 1, 2, F)
            run("24MP rgb8 cubic naps %d" % ns, 4000, 6000, 3, torch.uint8, 1, 2, F)
            run("33MP rgba8 cubic naps %d" % ns, 4320, 7680, 4, torch.uint8, 1, 2, F)
            run("100MP rgb16 cubic naps %d" % ns, 8192, 12288, 3, torch.int16, 2, 2, F)
            run("50MP rgba f32 cubic naps %d" % ns, 6144, 8192, 4, torch.float32, -4, 2, F)
    if which == "rgb8":
        for _ in range(2):
            run("24MP rgb8 cubic fast", 4000, 6000, 3, torch.uint8, 1, 2, F)
            run("24MP rgb8 linear fast", 4000, 6000, 3, torch.uint8, 1, 1, F)
            run("4K rgb8 cubic fast", 2160, 3840, 3, torch.uint8, 1, 2, F)
    if which == "widths":
        for w in (6144, 6000, 5888, 6100):
            run("rgb8 cubic w=%d" % w, 4000, w, 3, torch.uint8, 1, 2, F)
            run("rgb8 linear w=%d" % w, 4000, w, 3, torch.uint8, 1, 1, F)
    if which == "none16":
        run("100MP rgb16 none", 8192, 12288, 3, torch.int16, 2, 0, E)
        run("24MP rgb16 none", 4000, 6000, 3, torch.int16, 2, 0, E)
        run("8MP rgb16 none", 2160, 3840, 3, torch.int16, 2, 0, E)
        run("8K rgba16 none", 4320, 7680, 4, torch.int16, 2, 0, E, lens=(658, 1280))
        run("24MP rgba16 none", 4000, 6000, 4, torch.int16, 2, 0, E)
    if which == "none":
        run("100MP rgb16 none", 8192, 12288, 3, torch.int16, 2, 0, E)
        run("24MP rgb8 none", 4000, 6000, 3, torch.uint8, 1, 0, E)
        run("8K rgba16 none", 4320, 7680, 4, torch.int16, 2, 0, E, lens=(658, 1280))
        run("50MP rgb f32 none", 6144, 8192, 3, torch.float32, -4, 0, E)
        run("50MP rgba f32 none", 6144, 8192, 4, torch.float32, -4, 0, E)
    if which == "batch":
        run_batch("4K rgb8 cubic", 128, 2160, 3840, 3, torch.uint8, 1, 2, F)
        run_batch("4K rgb8 linear", 128, 2160, 3840, 3, torch.uint8, 1, 1, F)
        run_batch("4K rgb16 cubic", 64, 2160, 3840, 3, torch.int16, 2, 2, F)
        run_batch("1080p rgba8 cubic", 256, 1080, 1920, 4, torch.uint8, 1, 2, F)
    if which == "exact":
        run("100MP rgb16 cubic exact", 8192, 12288, 3, torch.int16, 2, 2, E, reps=5)
        run("100MP rgb16 linear exact", 8192, 12288, 3, torch.int16, 2, 1, E, reps=5)
        run("24MP rgb8 linear exact", 4000, 6000, 3, torch.uint8, 1, 1, E)
        run("24MP rgb8 cubic exact", 4000, 6000, 3, torch.uint8, 1, 2, E)
        run("8K rgba16 cubic exact", 4320, 7680, 4, torch.int16, 2, 2, E, lens=(658, 1280))
        run("50MP rgb f32 cubic exact", 6144, 8192, 3, torch.float32, -4, 2, E)
    if which == "x4":       # run with and without FIXCA_STREAM_NOALT=1 (separate processes: plans are cached)
        run("8K rgba16 cubic fast", 4320, 7680, 4, torch.int16, 2, 2, F, lens=(658, 1280))
        run("8K rgba16 linear fast", 4320, 7680, 4, torch.int16, 2, 1, F, lens=(658, 1280))
        run("50MP rgba f32 cubic fast", 6144, 8192, 4, torch.float32, -4, 2, F)
    if which == "narrow":
        for ctas in ("2", "4", "6"):
            os.environ["FIXCA_STREAM_CTAS"] = ctas
            run("24MP rgb8 cubic ctas" + ctas, 4000, 6000, 3, torch.uint8, 1, 2, F)
            run("24MP rgb8 linear ctas" + ctas, 4000, 6000, 3, torch.uint8, 1, 1, F)
            run("4K rgb8 cubic ctas" + ctas, 2160, 3840, 3, torch.uint8, 1, 2, F)
            run("8K rgba16 cubic ctas" + ctas, 4320, 7680, 4, torch.int16, 2, 2, F, lens=(658, 1280))
            run("33MP rgba8 cubic ctas" + ctas, 4320, 7680, 4, torch.uint8, 1, 2, F)
        os.environ["FIXCA_STREAM_DEBUG"] = "1"
        for ctas in ("2", "4"):
            os.environ["FIXCA_STREAM_CTAS"] = ctas
            run("24MP rgb8 cubic NOCOMPUTE ctas" + ctas, 4000, 6000, 3, torch.uint8, 1, 2, F)
            run("8K rgba16 cubic NOCOMPUTE ctas" + ctas, 4320, 7680, 4, torch.int16, 2, 2, F, lens=(658, 1280))
    if which == "pipe":
        for ctas in ("1", "2"):
            for d in ("1", "2", "3", "4", "6", "8"):
                os.environ["FIXCA_STREAM_CTAS"] = ctas
                os.environ["FIXCA_STREAM_DEPTH"] = d
                for dbg in ("0", "1"):
                    os.environ["FIXCA_STREAM_DEBUG"] = dbg
                    try:
                        run("cubic ctas%s D%s dbg%s" % (ctas, d, dbg), 8192, 12288, 3, torch.int16, 2, 2, F, reps=5)
                    except fixca.FixCaError as e:
                        print("ctas", ctas, "D", d, str(e)[:80])
    if which == "nonepipe":
        os.environ["FIXCA_VERBOSE"] = "1"
        run("100MP rgb16 none   default", 8192, 12288, 3, torch.int16, 2, 0, E, reps=5)
        run("100MP rgb16 linear default", 8192, 12288, 3, torch.int16, 2, 1, F, reps=5)
        run("100MP rgb16 cubic  default", 8192, 12288, 3, torch.int16, 2, 2, F, reps=5)
        os.environ["FIXCA_VERBOSE"] = "0"
        for ctas in ("1", "2", "3", "4"):
            for d in ("2", "3", "4", "6", "8"):
                os.environ["FIXCA_STREAM_CTAS"] = ctas
                os.environ["FIXCA_STREAM_DEPTH"] = d
                try:
                    run("none ctas%s D%s" % (ctas, d), 8192, 12288, 3, torch.int16, 2, 0, E, reps=5)
                except fixca.FixCaError as e:
                    print("ctas", ctas, "D", d, str(e)[:80])
    if which == "strip":
        for tw in ("256", "128"):
            os.environ["FIXCA_STRIP_TW"] = tw
            for th in (8, 16, 24, 32, 48):
                os.environ["FIXCA_TILE_H"] = str(th)
                try:
                    run("100MP rgb16 cubic fast tw" + tw, 8192, 12288, 3, torch.int16, 2, 2, F)
                except fixca.FixCaError as e:
                    print("tw", tw, "th", th, e)
        os.environ.pop("FIXCA_STRIP_TW")
        for th in (8, 16, 32):
            os.environ["FIXCA_TILE_H"] = str(th)
            run("100MP rgb16 linear fast", 8192, 12288, 3, torch.int16, 2, 1, F)
            run("24MP rgb8 cubic fast", 4000, 6000, 3, torch.uint8, 1, 2, F)
            run("24MP rgb8 linear fast", 4000, 6000, 3, torch.uint8, 1, 1, F)
            run("8K rgba16 cubic fast", 4320, 7680, 4, torch.int16, 2, 2, F, lens=(658, 1280))
            run("50MP rgb f32 cubic fast", 6144, 8192, 3, torch.float32, -4, 2, F)
        os.environ.pop("FIXCA_TILE_H")
    if which == "sweep":
        for th in (8, 16, 24, 32, 48, 64):
            os.environ["FIXCA_TILE_H"] = str(th)
            try:
                run("100MP rgb16 cubic fast", 8192, 12288, 3, torch.int16, 2, 2, F)
                run("100MP rgb16 cubic exact", 8192, 12288, 3, torch.int16, 2, 2, E, reps=3)
                run("24MP rgb8 cubic fast", 4000, 6000, 3, torch.uint8, 1, 2, F)
            except fixca.FixCaError as e:
                print("th", th, e)
