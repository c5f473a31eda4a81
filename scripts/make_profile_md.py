#!/usr/bin/env python
"""Developer helper: turn gpurun_out/ ncu artefacts into the tracked summaries under profiles/.
  python scripts/make_profile_md.py ncu   <tag> <title> <report.ncu-rep> <out.md>
  python scripts/make_profile_md.py list  <tag> <command> <launches.csv> <out.md>
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "kernel time"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput, % of ncu peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__cycles_elapsed.max", "SM cycles elapsed"),
    ("sm__cycles_active.avg", "SM cycles active (avg)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__block_size", "block size"),
    ("launch__grid_size", "grid size (CTAs)"),
    ("launch__shared_mem_per_block_dynamic", "dynamic shared memory / CTA"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "pipe lsu %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "pipe alu %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "pipe fma %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "pipe fp64 %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "pipe xu %"),
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return [{h: (v, u) for h, u, v in zip(rows[0], rows[1], vals)} for vals in rows[2:]]


def ncu_md(tag, title, rep, out):
    d = raw(rep)[0]
    lines = ["# %s -- %s" % (tag, title), "",
             "Source: `%s` (`ncu --set full --clock-control none --import-source on`, one B200 through gpurun; the" % rep,
             "same command had exited 0 without ncu first).  Kernel: `%s`." % d["Kernel Name"][0], "",
             "| metric | value |", "|---|---|"]
    for k, label in KEYS:
        if k in d:
            v, u = d[k]
            try:
                v = "%.4g" % float(v) if abs(float(v)) < 1e7 else "%.0f" % float(v)
            except ValueError:
                pass
            lines.append("| %s (`%s`) | %s %s |" % (label, k, v, u))
    t = float(d["gpu__time_duration.sum"][0])
    rd, wr = float(d["dram__bytes_read.sum"][0]), float(d["dram__bytes_write.sum"][0])
    unit = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
    by = rd * unit[d["dram__bytes_read.sum"][1]] + wr * unit[d["dram__bytes_write.sum"][1]]
    tus = t * {"us": 1.0, "ms": 1e3, "ns": 1e-3}[d["gpu__time_duration.sum"][1]]
    lines.append("| DRAM read+write per launch / kernel time | %.0f bytes / %.1f us = %.0f GB/s |" % (by, tus, by / tus / 1e3))
    st = sorted(((float(v[0]), h.split("issue_stalled_")[1].split("_per_issue")[0]) for h, v in d.items()
                 if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and v[0] not in ("", "n/a")), reverse=True)
    lines.append("| top stalls (warps per issue) | %s |" % ", ".join("%s %.2f" % (n, v) for v, n in st[:7]))
    open(out, "w").write("\n".join(lines) + "\n")
    print(out, "%.0f bytes" % by)


def list_md(tag, cmd, path, out):
    txt = open(path).read()
    body = txt[txt.index('"ID"'):]
    rows = list(csv.DictReader(io.StringIO(body)))
    lines = ["# %s -- ncu launch list of `%s`" % (tag, cmd), "",
             "`ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv` on one B200 (gpurun); cold-cache,",
             "serialised launch times: compare shares, not absolutes.", "", "| kernel | grid | block | launches | total us | avg us |",
             "|---|---|---|---|---|---|"]
    agg = {}
    for r in rows:
        name = r["Kernel Name"]
        name = name.split("(")[0] if len(name) > 90 else name
        name = name[:110]
        key = (name, r["Grid Size"], r["Block Size"])
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += float(r["Metric Value"]) / 1e3
    tot = sum(a[1] for a in agg.values())
    for (name, g, b), (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append("| `%s` | %s | %s | %d | %.1f | %.1f |" % (name, g, b, n, us, us / n))
    ours = sum(a[1] for k, a in agg.items() if "fixca" in k[0] or "stream_kernel" in k[0])
    lines += ["", "fixca kernels: %.1f us of %.1f us listed (%.1f %%); the rest is torch's synthesis / fill of the input buffers"
              % (ours, tot, 100 * ours / tot), "outside the timed region."]
    lines += ["", "Every fixca launch in order (us): " + ", ".join("%.1f" % (float(r["Metric Value"]) / 1e3) for r in rows
                                                                   if "fixca" in r["Kernel Name"])]
    open(out, "w").write("\n".join(lines) + "\n")
    print(out)


if __name__ == "__main__":
    if sys.argv[1] == "ncu":
        ncu_md(*sys.argv[2:6])
    else:
        list_md(*sys.argv[2:6])
