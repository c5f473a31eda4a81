#!/usr/bin/env python
"""oracle/patch_plugin.py -- TEST INFRASTRUCTURE ONLY.

Applies the INTEGRATION.md patch to a scratch copy of the reference plug-in source so that the
reference's own run() / fix_ca() can be driven (behind the fake GIMP of ref_harness.c) with its row
loop replaced by fixca_cuda_region().  The reference source is read where it lies and the patched
text is written to the path given (a temporary file: it is compiled into
oracle/_ref/libfixca_plugin_cuda.so and deleted; no reference source enters this repository).

    python oracle/patch_plugin.py /root/reference/fix-ca.c /tmp/fix-ca-cuda.c
"""
import re
import sys

INCLUDE_BLOCK = r'''
#ifdef HAVE_FIXCA_CUDA
#include <fixca_cuda.h>	/* B200 correction pass; FixCaParams is layout-identical to fixca_params */
static void fixca_progress (int kind, double fraction, void *user)
{
	(void) user;
	if (kind == 0)
		gimp_progress_init (_("Shifting pixel components..."));
	else
		gimp_progress_update (fraction);
}
/* The two whole-image buffers (fix-ca.c:366-367, :648-649) come from pinned memory when a GPU is present, so
 * that the library moves them by DMA without staging copies; g_new stays the fallback. */
static void *fixca_pinned[4];
static guchar *fixca_img_new (gsize n)
{
	void *p = fixca_cuda_device_count () > 0 ? fixca_cuda_host_alloc (n) : NULL;
	int i;
	for (i = 0; p && i < 4; i++)
		if (!fixca_pinned[i]) {
			fixca_pinned[i] = p;
			return p;
		}
	if (p)
		fixca_cuda_host_free (p);
	return g_new (guchar, n);
}
static void fixca_img_free (void *p)
{
	int i;
	for (i = 0; p && i < 4; i++)
		if (fixca_pinned[i] == p) {
			fixca_pinned[i] = NULL;
			fixca_cuda_host_free (p);
			return;
		}
	g_free (p);
}
#define FIXCA_IMG_NEW(n) fixca_img_new (n)
#define FIXCA_IMG_FREE(p) fixca_img_free (p)
#else
#define FIXCA_IMG_NEW(n) g_new (guchar, n)
#define FIXCA_IMG_FREE(p) g_free (p)
#endif
'''

CALL_BLOCK = r'''
#ifdef HAVE_FIXCA_CUDA
	fixca_cuda_set_progress (fixca_progress, NULL);
	/* any failure of the library (no usable GPU, out of memory, a format it declines) falls back to the CPU loop */
	if (x == 0 && width == xImg && fixca_cuda_device_count () > 0 &&
	    fixca_cuda_region (srcImg, destImg, xImg, yImg, bppImg, bpcImg, (const fixca_params *) params,
			       x, x + width, y, y + height, TRUE) == 0)
		;
	else
#endif
'''


PREVIEW_BLOCK = r'''
#ifdef HAVE_FIXCA_CUDA
	/* the pass with show_progress = FALSE (saturation boost and centre lines included) and the 8-bit
	 * down-conversion of the window, on the GPU; only the window's bytes come back */
	b = bpcImg < 0 ? -bpcImg : bpcImg;
	if (fixca_cuda_device_count () > 0 &&
	    fixca_cuda_preview (srcImg, prevImg, xImg, yImg, bppImg, bpcImg, (const fixca_params *) params,
				x, y, width, height) == 0)
		;
	else {
#endif
'''

PREVIEW_END = r'''
#ifdef HAVE_FIXCA_CUDA
	}
#endif
'''


def patch(text: str) -> str:
    # (1) the binding, in front of fix_ca() (fix-ca.c:332), i.e. after the gettext macros it uses
    m = re.search(r'\nstatic int fix_ca \(gint32 drawable_ID', text)
    if not m:
        raise SystemExit("patch_plugin: fix_ca() not found")
    text = text[:m.start()] + "\n" + INCLUDE_BLOCK + text[m.start():]
    # (2) the final-render call in fix_ca() (fix-ca.c:373-374): the one call that passes TRUE
    calls = list(re.finditer(r'\n([ \t]*)fix_ca_region \(srcImg, destImg,[^;]*?TRUE\);', text, re.S))
    if len(calls) != 1:
        raise SystemExit("patch_plugin: expected exactly one final-render call, found %d" % len(calls))
    c = calls[0]
    text = text[:c.start()] + "\n" + CALL_BLOCK.strip("\n") + text[c.start():]
    # (3) the preview call in preview_update() (fix-ca.c:656-657): the one call that passes FALSE
    calls = list(re.finditer(r'\n([ \t]*)fix_ca_region \(srcImg, destImg,[^;]*?FALSE\);', text, re.S))
    if len(calls) != 1:
        raise SystemExit("patch_plugin: expected exactly one preview call, found %d" % len(calls))
    c = calls[0]
    end = text.index("\n\tgimp_preview_draw_buffer (ptr, prevImg", c.end())     # after the conversion loop (:659-671)
    text = (text[:c.start()] + "\n" + PREVIEW_BLOCK.strip("\n") + text[c.start():end] + "\n" + PREVIEW_END.strip("\n") +
            text[end:])
    # (4) the whole-image buffers of fix_ca() and preview_update() (fix-ca.c:366-367, :648-649 and their g_free's)
    text, n_new = re.subn(r'\b(srcImg|destImg)(\s*)= g_new \(guchar, ([^;]*)\);', r'\1\2= FIXCA_IMG_NEW (\3);', text)
    text, n_free = re.subn(r'\bg_free ?\((srcImg|destImg)\);', r'FIXCA_IMG_FREE (\1);', text)
    if n_new != 4 or n_free != 4:
        raise SystemExit("patch_plugin: expected 4 buffer allocations and 4 frees, found %d / %d" % (n_new, n_free))
    return text


if __name__ == "__main__":
    src, dst = sys.argv[1], sys.argv[2]
    with open(src) as f:
        out = patch(f.read())
    with open(dst, "w") as f:
        f.write(out)
