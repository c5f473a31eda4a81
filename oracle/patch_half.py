#!/usr/bin/env python
"""oracle/patch_half.py -- TEST INFRASTRUCTURE ONLY.

The reference carries its half-precision support as commented-out code ("TODO for another day"):
color_size() would return -2 (fix-ca.c:692-693), get_pixel() would read a `half` (:740-742) and
set_pixel() would store one (:768-770).  This script enables exactly those lines in a scratch copy of
the reference source (plus `typedef _Float16 half;`, the type the comment names), so that the
reference's own row loop defines what "Fix-CA on a 16-bit float image" computes.  The copy is compiled
into oracle/_ref/libfixca_ref_half.so and deleted; no reference source enters this repository.

    python oracle/patch_half.py /root/reference/fix-ca.c /tmp/fix-ca-half.c
"""
import sys


def patch(text: str) -> str:
    edits = [
        # color_size (fix-ca.c:692-693)
        ('\t//if (strstr(str, "half") != NULL)\n\t//\treturn -2; /* IEEE 754 half precision */',
         '\tif (strstr(str, "half") != NULL)\n\t\treturn -2; /* IEEE 754 half precision */'),
        # get_pixel (fix-ca.c:740-742); the cast is needed because ptr is a guchar *
        ('\t//} else if (bpc == -2) {\n\t//\thalf *p = ptr;\n\t//\tret += *p;\n',
         '\t} else if (bpc == -2) {\n\t\thalf *p = (half *)(ptr);\n\t\tret += *p;\n'),
        # set_pixel (fix-ca.c:768-770)
        ('\t//} else if (bpc == -2) {\n\t//\thalf *p = (half *)(dest);\n\t//\t*p = d;\n',
         '\t} else if (bpc == -2) {\n\t\thalf *p = (half *)(dest);\n\t\t*p = d;\n'),
    ]
    for old, new in edits:
        if text.count(old) != 1:
            raise SystemExit("patch_half.py: expected exactly one occurrence of %r" % old[:40])
        text = text.replace(old, new)
    return "typedef _Float16 half;\t/* oracle/patch_half.py */\n" + text


if __name__ == "__main__":
    src, dst = sys.argv[1], sys.argv[2]
    with open(src) as f:
        out = patch(f.read())
    with open(dst, "w") as f:
        f.write(out)
