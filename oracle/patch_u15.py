#!/usr/bin/env python
"""oracle/patch_u15.py -- TEST INFRASTRUCTURE ONLY.

The reference rejects babl's 15-bit unsigned samples: color_size() answers -99 for names containing "u15"
("TODO for another day", fix-ca.c:694-695), and get_pixel() / set_pixel() have no branch for them.  This
script writes the missing lines into a scratch copy of the reference source, in the reference's own pattern
for its other unsigned types, so that the reference's row loop defines what "Fix-CA on a u15 image" computes:

  * color_size():  "u15" -> 15                          (the code include/fixca_cuda.h calls FIXCA_BPC_U15)
  * get_pixel():   bpc == 15:  ret += *(uint16_t *)ptr; ret /= 32768;     (u15: 0 .. 32768 <-> [0.0, 1.0])
  * set_pixel():   bpc == 15:  *(uint16_t *)dest = round(d * 32768);
  * the four places that take the sample size as absolute(bpc) (fix-ca.c:659, :927, :948, :1085) use 2 bytes
    for code 15.

Everything else is the unmodified source.  The copy is compiled into oracle/_ref/libfixca_ref_u15.so and
deleted; no reference source enters this repository.

    python oracle/patch_u15.py /root/reference/fix-ca.c /tmp/fix-ca-u15.c
"""
import re
import sys


def patch(text: str) -> str:
    edits = [
        # color_size (fix-ca.c:694-695)
        ('\tif (strstr(str, "u15") != NULL)\n\t\treturn -99; /* TODO for another day */',
         '\tif (strstr(str, "u15") != NULL)\n\t\treturn 15; /* oracle/patch_u15.py: 15-bit unsigned in 16 bits */'),
        # get_pixel (after the bpc == 2 branch, fix-ca.c:720-723)
        ('\t\tret /= 65535;\n',
         '\t\tret /= 65535;\n\t} else if (bpc == 15) {\n\t\tuint16_t *p = (uint16_t *)(ptr);\n\t\tret += *p;\n\t\tret /= 32768;\n'),
        # set_pixel (after the bpc == 2 branch, fix-ca.c:753-755)
        ('\t\t*p = round(d * 65535);\n',
         '\t\t*p = round(d * 65535);\n\t} else if (bpc == 15) {\n\t\tuint16_t *p = (uint16_t *)(dest);\n\t\t*p = round(d * 32768);\n'),
    ]
    for old, new in edits:
        if text.count(old) != 1:
            raise SystemExit("patch_u15.py: expected exactly one occurrence of %r" % old[:40])
        text = text.replace(old, new)
    # sample size in bytes: absolute(bpc) everywhere (fix-ca.c:659, :927, :948, :1085)
    text, n = re.subn(r"absolute ?\((bpc(?:Img)?)\)", r"u15_sample_bytes(\1)", text)
    if n != 4:
        raise SystemExit("patch_u15.py: expected 4 absolute(bpc) sites, found %d" % n)
    head = ("static int u15_sample_bytes (int bpc) { return bpc == 15 ? 2 : (bpc < 0 ? -bpc : bpc); }"
            "\t/* oracle/patch_u15.py */\n")
    return head + text


if __name__ == "__main__":
    src, dst = sys.argv[1], sys.argv[2]
    with open(src) as f:
        out = patch(f.read())
    with open(dst, "w") as f:
        f.write(out)
