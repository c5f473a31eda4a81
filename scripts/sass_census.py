#!/usr/bin/env python
"""Developer helper: SASS evidence of the built library (cuobjdump -sass, sm_100a) -> profiles/r02_sass_census.md and
gzipped listings of the two row-loop kernels bench.py measures.
    python scripts/sass_census.py"""
import collections
import gzip
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gimp-fix-ca_b200", "lib", "libfixca_cuda.so")
MNEMONICS = ["UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "UTMAPF", "FFMA", "FMUL", "DFMA", "DMUL", "DADD", "LDS", "STS", "LDG", "PRMT",
             "ACQBULK", "UTMACMDFLUSH", "ELECT"]
KERNELS = {"stream_kernel<unsigned short,3,2,2,256> (headline)": "_ZN5fixca13stream_kernelItLi3ELi2ELi2ELi256ELb0ELi0ELb0EEEvNS_10KernelArgsE14CUtensorMap_stS2_S2_NS_12StreamFanoutE",
           "stream_kernel<unsigned char,3,2,4,256> (cfg5)": "_ZN5fixca13stream_kernelIhLi3ELi2ELi4ELi256ELb0ELi0ELb0EEEvNS_10KernelArgsE14CUtensorMap_stS2_S2_NS_12StreamFanoutE",
           "stream_kernel<unsigned char,3,2,4,256,REPAIR=2> (8-bit EXACT)": "_ZN5fixca13stream_kernelIhLi3ELi2ELi4ELi256ELb0ELi2ELb0EEEvNS_10KernelArgsE14CUtensorMap_stS2_S2_NS_12StreamFanoutE",
           "stream_kernel<unsigned short,3,2,2,256,WIDE> (16-bit EXACT)": "_ZN5fixca13stream_kernelItLi3ELi2ELi2ELi256ELb0ELi0ELb1EEEvNS_10KernelArgsE14CUtensorMap_stS2_S2_NS_12StreamFanoutE"}


def census(text):
    c = collections.Counter()
    for m in re.finditer(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", text, re.M):
        c[m.group(1)] += 1
    return c


def main():
    whole = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    cw = census(whole)
    cols = {}
    for label, sym in KERNELS.items():
        t = subprocess.run(["cuobjdump", "-sass", "-fun", sym, LIB], capture_output=True, text=True).stdout
        cols[label] = census(t)
        tag = "u16x3" if "short" in label else "u8x3"
        with gzip.open(os.path.join(ROOT, "profiles", "r02_sass_stream_%s_cubic.txt.gz" % tag), "wt") as f:
            f.write(t)
    elfs = re.findall(r"ELF file\s+\d+: (\S+)", subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout)
    with open(os.path.join(ROOT, "profiles", "r02_sass_census.md"), "w") as f:
        f.write("# r02 -- SASS mnemonic census of gimp-fix-ca_b200/lib/libfixca_cuda.so (cuobjdump -sass, sm_100a)\n\n")
        f.write("Cubins: %s\n\n" % ", ".join(elfs))
        f.write("| mnemonic | whole library | " + " | ".join(cols) + " |\n|---|---|" + "---|" * len(cols) + "\n")
        for mn in MNEMONICS:
            f.write("| %s | %d | %s |\n" % (mn, cw[mn], " | ".join(str(cols[k][mn]) for k in cols)))
        f.write("\nTMA tensor copies are UTMALDG / UTMASTG, the 1-D bulk copies (chunk records, strip / tiled kernels) UBLKCP, mbarrier "
                "operations SYNCS.  Full listings of the two row-loop kernels: profiles/r02_sass_stream_u16x3_cubic.txt.gz (headline), "
                "profiles/r02_sass_stream_u8x3_cubic.txt.gz (cfg5).\n")
    print(open(os.path.join(ROOT, "profiles", "r02_sass_census.md")).read())


if __name__ == "__main__":
    main()
