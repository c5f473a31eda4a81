#!/usr/bin/env python
"""Developer probe: run named golden cases through two kernel choices and print where they differ.
    python scripts/debug_case.py <case name> [<case name> ...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("gimp-fix-ca_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np  # noqa: E402

import fixca  # noqa: E402
import oracle as orc  # noqa: E402
from helpers import case_image, fx_params, oracle_params  # noqa: E402

cases = {c["name"]: c for c in json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))["suite"]}
chk = orc.best_checker()
for name in sys.argv[1:]:
    c = cases[name]
    img = case_image(c)
    want = chk.region(img, oracle_params(c))
    os.environ.pop("FIXCA_EXACT_KERNEL", None)
    fixca.reload_tuning()
    a = fixca.correct(img, fx_params(fixca, c))
    ka = fixca.last_kernel()
    os.environ["FIXCA_EXACT_KERNEL"] = "tiled"
    fixca.reload_tuning()
    b = fixca.correct(img, fx_params(fixca, c))
    kb = fixca.last_kernel()
    print(name, {k: c[k] for k in c if k not in ("name", "md5")})
    for lab, got, k in (("auto", a, ka), ("tiled", b, kb)):
        d = np.argwhere(got != want)
        print("  %-5s %-28s mismatches %d" % (lab, k, len(d)))
        for y, x, ch in d[:12]:
            print("     y=%d x=%d ch=%d got=%d want=%d" % (y, x, ch, got[y, x, ch], want[y, x, ch]))
        if len(d):
            print("     rows", sorted(set(d[:, 0]))[:20], "cols", sorted(set(d[:, 1]))[:20])
