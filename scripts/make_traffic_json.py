#!/usr/bin/env python
"""Developer helper: rebuild profiles/traffic.json (what bench.py's roofline.traffic reads) from the tracked
ncu summaries profiles/r02_ncu_<workload>_<tag>.md that scripts/evidence_r02.sh produced.
  python scripts/make_traffic_json.py <tag>
The batch workload is captured on 32 frames (profile_one.py) and scaled to the frames one bench launch holds."""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

KERNEL = {  # template arguments of the captured kernel -> the name the library reports
    "unsigned char": "u8", "unsigned short": "u16", "float": "f32",
}


def kernel_name(md_kernel: str, exact: bool) -> str:
    m = re.search(r"(stream|tiled)_kernel<([^,>]+), (\d+), (\d+)", md_kernel)
    fam, ty, ch, interp = m.group(1), m.group(2).strip(), int(m.group(3)), int(m.group(4))
    if fam == "tiled":   # tiled_kernel<Sample, CH, INTERP, Arith>
        return "tiled/%s/%s/%sx%d" % (["none", "linear", "cubic"][interp], "f64" if exact else "f32", KERNEL[ty], ch)
    # stream_kernel<Sample, CH, INTERP, P, TW, ALT, REPAIR, WIDE>
    flags = re.search(r"stream_kernel<[^>]*?, (\d+), (\d+), (\d+)>", md_kernel)
    arith = "f32"
    if flags and int(flags.group(3)):
        arith = "f64+exact"
    elif flags and int(flags.group(2)):     # REPAIR: 1 in-kernel queues, 2 deferred (repair_patch_kernel behind it)
        arith = "f32+f64" if int(flags.group(2)) == 2 else "f32+f64inline"
    return "stream/%s/%s/%sx%d" % (["none", "linear", "cubic"][interp], arith, KERNEL[ty], ch)


def main(tag: str) -> None:
    out = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the ncu "
                       "--set full captures of scripts/evidence_r02.sh (L2 flushed before the captured launch); keyed by "
                       "bench workload and by the kernel name the library reports"}
    for fn in sorted(os.listdir(os.path.join(ROOT, "profiles"))):
        m = re.match(r"r02_ncu_(.+)_%s\.md$" % re.escape(tag), fn)
        if not m:
            continue
        wl = m.group(1)
        exact = wl.endswith("_exact")
        name = wl[:-6] if exact else wl
        txt = open(os.path.join(ROOT, "profiles", fn)).read()
        kern = re.search(r"Kernel: `([^`]+)`", txt).group(1)
        by = float(re.search(r"\| (\d+) bytes / ", txt).group(1))
        note = "profiles/" + fn
        frames = bench.WORKLOAD_FRAMES.get(name, 0)
        if frames > 32:
            note += " (captured on 32 frames: %d B; the bench launch holds %d frames)" % (by, frames)
            by = by * frames / 32
        out.setdefault(name, {})[kernel_name(kern, exact)] = {"dram_bytes_per_launch": int(by), "capture": note}
    with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
        json.dump(out, f, indent=2)
        f.write("\n")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1])
