/*
 * fixca_cuda.h -- C ABI of the B200 (sm_100a) implementation of Fix-CA's
 * per-pixel correction pass.
 *
 * This is the drop-in boundary: plain C, plain pointers and sizes.  Each entry
 * point names the piece of the reference plug-in (JoesCat/gimp-fix-ca,
 * fix-ca.c) it replaces or restates.  INTEGRATION.md shows the patch (~50 lines)
 * that makes fix-ca.c call fixca_cuda_region() instead of its CPU row loop.
 *
 * There is no CPU fallback: every compute entry point fails with
 * FIXCA_ERR_CUDA / FIXCA_ERR_NO_DEVICE when no usable GPU is present.
 */
#ifndef FIXCA_CUDA_H
#define FIXCA_CUDA_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define FIXCA_API __declspec(dllexport)
#else
#define FIXCA_API __attribute__((visibility("default")))
#endif

/* ------------------------------------------------------------------------- */
/* Parameter block                                                            */
/* ------------------------------------------------------------------------- */

/*
 * Layout-identical to the reference's FixCaParams (fix-ca.c:70-82; 80 bytes on
 * x86-64), so the plug-in can pass its own struct through a cast.  Values are
 * the post-marshalling ones: lens_x / lens_y are truncated with (int) inside
 * the pass (fix-ca.c:1033-1034); there is no -1,-1 reset here, that lives in
 * the dialog only (fix-ca.c:427-428) -- see fixca_resolve_lens().
 */
typedef struct fixca_params {
	double blue;		/* lateral amount, blue  (fix-ca.c:71) */
	double red;		/* lateral amount, red   (fix-ca.c:72) */
	double lens_x;		/* lens centre           (fix-ca.c:73) */
	double lens_y;		/*                       (fix-ca.c:74) */
	int    update_preview;	/* unused, as in the reference (fix-ca.c:75,223) */
	int    interpolation;	/* 0 None, 1 Linear, 2 Cubic  (fix-ca.c:76,156) */
	double saturation;	/* preview only (fix-ca.c:77); ignored by the pass */
	double x_blue;		/* directional shifts (fix-ca.c:78-81) */
	double x_red;
	double y_blue;
	double y_red;
} fixca_params;

#define FIXCA_INTERP_NONE   0
#define FIXCA_INTERP_LINEAR 1
#define FIXCA_INTERP_CUBIC  2

/* Largest |amount| run() accepts: INPUT_MAX = SOURCE_ROWS/4 (fix-ca.c:64-65,279-292). */
#define FIXCA_INPUT_MAX 30.0

/* `bpc` codes, as produced by color_size() (fix-ca.c:681-711):
 *   1, 2, 4, 8  unsigned integer samples of that many bytes
 *   -4, -8      float, double
 *   -99         unsupported (half, u15, ...)                                  */
#define FIXCA_BPC_UNSUPPORTED (-99)
/* Extensions, the two formats color_size() marks "TODO for another day":
 *   -2   IEEE half (fix-ca.c:692-693, the reference's own commented-out branches)
 *   15   babl "u15": 15-bit unsigned in 16-bit storage, 0..32768 <-> [0,1] (fix-ca.c:694-695 rejects it; here
 *        it gets the arithmetic of the other unsigned types with max = 32768: get_pixel v / 32768, set_pixel
 *        round(d * 32768) after clip_d; None copies the 16-bit storage untouched)               */
#define FIXCA_BPC_HALF (-2)
#define FIXCA_BPC_U15  15

/* ------------------------------------------------------------------------- */
/* Status codes (0 = success; the reference's region returns void, its driver */
/* maps any failure to GIMP_PDB_CALLING_ERROR, fix-ca.c:315-316)               */
/* ------------------------------------------------------------------------- */
#define FIXCA_OK               0
#define FIXCA_ERR_ARG         (-1)	/* NULL pointer, non-positive size, bad rows   */
#define FIXCA_ERR_FORMAT      (-2)	/* bpc / bytes-per-pixel not one the reference handles */
#define FIXCA_ERR_INTERP      (-3)	/* interpolation outside 0..2                  */
#define FIXCA_ERR_REGION      (-4)	/* x1 != 0 or x2 != width: the reference's own
					   column-selection path is broken (SURVEY.md
					   App. D #3); only full-width row bands are defined */
#define FIXCA_ERR_DEGENERATE  (-5)	/* max_dim + amount == 0: scale is infinite and the
					   reference itself indexes out of bounds        */
#define FIXCA_ERR_RANGE       (-6)	/* fixca_check_params(): outside +-FIXCA_INPUT_MAX */
#define FIXCA_ERR_NO_DEVICE   (-7)
#define FIXCA_ERR_CUDA        (-8)	/* see fixca_cuda_last_error()                 */
#define FIXCA_ERR_UNSUPPORTED (-9)	/* a flag this entry point does not take
					   (FIXCA_PREVIEW_OVERLAY on a batch call, ...);
					   u64 Linear/Cubic is computed since r02: the
					   x87 long double steps of fix-ca.c:728-733,
					   759-761 restated in integer arithmetic        */
#define FIXCA_ERR_NOMEM       (-10)

/* ------------------------------------------------------------------------- */
/* Flags                                                                      */
/* ------------------------------------------------------------------------- */
/* Arithmetic of Linear/Cubic (None is always a byte-exact gather):
 *   EXACT  FP64 in the reference's operation order, no FMA contraction:
 *          bit-identical to the reference for every format.  Default.
 *   FAST   FP32 separable weights on raw sample values: within +-1 LSB of the
 *          reference for u8/u16 (2 ulp for float); u32/f64 still run EXACT.   */
#define FIXCA_PRECISION_EXACT  0x0u
#define FIXCA_PRECISION_FAST   0x1u
#define FIXCA_PRECISION_MASK   0x3u
/* Kernel selection, for tests and profiling (default: tiled when it fits). */
#define FIXCA_FORCE_DIRECT     0x10u	/* per-pixel global-memory gather kernel */
#define FIXCA_FORCE_TILED      0x20u	/* fail instead of falling back to direct */
/* The preview epilogue of fix_ca_region(..., show_progress = FALSE) (fix-ca.c:1322-1327): after the
 * correction, saturate() (iff params->saturation != 0, fix-ca.c:922-943) and centerline() (the dashed
 * lens cross and diagonals, fix-ca.c:945-996) are applied to every row written.  Set automatically by
 * fixca_cuda_region*() when show_progress == 0; pass it to fixca_cuda_region_dev() to get the same. */
#define FIXCA_PREVIEW_OVERLAY  0x40u
/* Extension ("correct selection handling"): accept x1 != 0 / x2 != width in
 * fixca_cuda_region_ex().  The reference's arithmetic for a pixel never depends on x1 / x2 (every index in
 * fix-ca.c:1105-1320 is an absolute column), but three buffer-indexing bugs break its own x1 != 0 path
 * (band_adj multiplied by `bytes` twice :1084/:856, the green copy offset in pixels instead of bytes :1098,
 * and fix_ca() handing GEGL a selection-strided rectangle :369-370), so without this flag such calls return
 * FIXCA_ERR_REGION.  With it, columns [x1,x2) of rows [y1,y2) of dst receive exactly what the full-width
 * pass computes for them, and nothing else in dst is written (src and dst stay whole-image buffers,
 * width * height * bytes, as set_data() addresses them, :864-871). */
#define FIXCA_COLUMN_SELECTION 0x80u
/* Device-resident entries only.  The TMA kernels store whole 16-byte units: when width * bytes is not a multiple of
 * 16 they write up to 15 bytes past the end of every destination row.  In the padding of pitched rows that is
 * harmless, in a sub-rectangle view of a wider buffer it would overwrite live pixels -- so without this flag such
 * calls take the per-pixel direct kernel (exact row ends, several times slower).  Pass it when the bytes between
 * width * bytes and the next 16-byte boundary of each destination row are scratch. */
#define FIXCA_PADDING_SCRATCH  0x100u

/* ------------------------------------------------------------------------- */
/* The pass, host buffers: replaces fix_ca_region()                           */
/* ------------------------------------------------------------------------- */
/*
 * Same arguments and meaning as
 *   static void fix_ca_region (guchar *srcPTR, guchar *dstPTR, gint orig_width,
 *       gint orig_height, gint bytes, gint bpc, FixCaParams *params,
 *       gint x1, gint x2, gint y1, gint y2, gboolean show_progress)
 * (fix-ca.c:998-1001; call sites :373-374 and :656-657).  src and dst are
 * caller-owned host arrays of width*height*bytes, tight rows.  Reads src only;
 * writes exactly rows [y1,y2) of dst.  Synchronous.  Internally: pinned
 * staging, H2D of the source rows the band needs, sm_100a kernels, D2H.
 *
 * show_progress != 0 issues the reference's progress sequence through the
 * callback installed with fixca_cuda_set_progress().  show_progress == 0 is
 * the dialog's preview call (fix-ca.c:656-657): as in the reference, the
 * saturation boost (params->saturation) and the centre-line overlay are then
 * drawn into the rows written (FIXCA_PREVIEW_OVERLAY) and no progress is
 * reported.  The HSV conversion behind the saturation boost is libgimpcolor's,
 * restated (it is not part of the reference tree).
 */
FIXCA_API int fixca_cuda_region(const unsigned char *src, unsigned char *dst,
				int width, int height, int bytes, int bpc,
				const fixca_params *params,
				int x1, int x2, int y1, int y2, int show_progress);

/*
 * The middle of preview_update() in one call (fix-ca.c:656-671): fix_ca_region(src, dest, ..., 0, width, y, y + ph,
 * FALSE) -- the pass over the preview's rows with the saturation boost and the centre lines -- followed by the
 * down-conversion of window columns [x, x + pw) to the 8-bit buffer the plug-in hands to
 * gimp_preview_draw_buffer(): every sample through get_pixel() and set_pixel(..., 1) (8-bit drawables: a copy).
 * `prev` is pw * ph * channels bytes, tight rows; only those bytes come back from the GPU (no destImg at all).
 */
FIXCA_API int fixca_cuda_preview(const unsigned char *src, unsigned char *prev, int width, int height,
				 int bytes, int bpc, const fixca_params *params, int x, int y, int pw, int ph);

/* As above with FIXCA_* flags and an explicit CUDA device ordinal (-1 = current). */
FIXCA_API int fixca_cuda_region_ex(const unsigned char *src, unsigned char *dst,
				   int width, int height, int bytes, int bpc,
				   const fixca_params *params,
				   int x1, int x2, int y1, int y2, int show_progress,
				   unsigned flags, int device);

/*
 * Row-banded across several GPUs of one box from one process (the multi-GPU
 * form of fix_ca()'s middle, fix-ca.c:366-377): rows [y1,y2) are split into
 * `ndev` contiguous full-width bands; each device receives its band plus the
 * halo rows fixca_band_source_rows() reports, and writes its rows straight
 * into the caller's dst.  Bands are independent (fix-ca.c:1091-1329), so there
 * is no exchange between devices.  devices == NULL means 0..ndev-1.
 */
FIXCA_API int fixca_cuda_region_multi(const unsigned char *src, unsigned char *dst,
				      int width, int height, int bytes, int bpc,
				      const fixca_params *params, int y1, int y2,
				      unsigned flags, const int *devices, int ndev);

/* ------------------------------------------------------------------------- */
/* The pass, device-resident (benchmarks, pipelines, per-rank bands)          */
/* ------------------------------------------------------------------------- */
/*
 * d_src holds source rows [src_row0, src_row0 + src_rows) of a width x height
 * image, row pitch src_pitch bytes; d_dst receives output rows [y1,y2) at
 * d_dst + (y - dst_row0) * dst_pitch.  For a whole resident image pass
 * src_row0 = dst_row0 = 0, src_rows = height.  The source rows present must
 * cover fixca_band_source_rows(y1,y2) or FIXCA_ERR_ARG is returned.
 * Asynchronous on `stream` (a cudaStream_t, NULL = default stream).  The TMA
 * kernels need both pitches to be multiples of 16 bytes, both pointers 16-byte
 * aligned, and rows that are whole 16-byte units (or FIXCA_PADDING_SCRATCH);
 * otherwise the direct kernel is used.
 * The first call with a new combination of geometry, band and buffers makes a
 * launch plan: it allocates and fills small per-plan tables in device memory
 * on a private stream and waits for them (tens of microseconds); repeated
 * calls take the plan from a per-thread cache and only launch.  Make that
 * first call outside a CUDA stream capture.
 * EXACT Linear / Cubic on 8-bit samples is two launches on `stream` (the streaming kernel, then
 * repair_patch_kernel, which recomputes the near-tie samples the first one queued); the queue
 * between them is device memory kept per calling thread, device and stream (grown to the
 * largest launch seen; fixca_cuda_release() frees the calling thread's).
 */
FIXCA_API int fixca_cuda_region_dev(const void *d_src, size_t src_pitch, int src_row0, int src_rows,
				    void *d_dst, size_t dst_pitch, int dst_row0,
				    int width, int height, int bytes, int bpc,
				    const fixca_params *params, int y1, int y2,
				    unsigned flags, void *stream);

/*
 * A batch of `nframes` equal-sized device-resident frames, same parameters (BASELINE
 * "batch stream of frames": fix-ca.c has no such call; it is fix_ca_region() over every
 * frame, :373-374).  Frame i starts at d_src + i * src_frame_stride (bytes; strides and
 * pitches multiples of 16 for the streaming kernels).  The streaming kernels take the
 * whole batch in ONE launch (a grid layer per frame, long row segments); other kernels
 * are launched per frame.  Asynchronous on `stream`.
 */
FIXCA_API int fixca_cuda_frames_dev(const void *d_src, size_t src_pitch, size_t src_frame_stride,
				    void *d_dst, size_t dst_pitch, size_t dst_frame_stride, int nframes,
				    int width, int height, int bytes, int bpc,
				    const fixca_params *params, unsigned flags, void *stream);

/*
 * Peer frames: the reassembly step of a row-banded image ("reassembly is a P2P/NCCL gather
 * over NVLink only") folded into the pass itself.  One process (the frame's owner) allocates the whole
 * destination frame on its GPU and exports a CUDA IPC handle; every other rank of the box opens the handle
 * (peer access over NVLink / NVSwitch is enabled by the open) and passes the mapped pointer as `d_dst` of
 * fixca_cuda_region_dev() with dst_row0 = 0: its kernel then stores each finished chunk of its band straight
 * into the owner's memory (the same TMA tensor stores, through the peer mapping), so the compute and the gather
 * are ONE kernel and no staging band, second copy or concatenation exists.  The owner sees the rows once the
 * writer's stream has completed and the ranks have met (an event / barrier of the caller's).
 * fix-ca.c has no counterpart (one process, one buffer: destImg, :367).
 *
 * fixca_cuda_frame_alloc: cudaMalloc on the current device + cudaIpcGetMemHandle.
 * fixca_cuda_frame_open:  cudaIpcOpenMemHandle in ANOTHER process (CUDA refuses the exporting process).
 * fixca_cuda_frame_close / _free: undo them (after the ranks have met).
 *
 * All-gather form (every GPU ends up with the whole frame): each rank allocates a frame, opens everybody else's and
 * calls fixca_cuda_region_dev_fanout() with all of them: the streaming kernels store every finished chunk from
 * shared memory into each frame (its own through HBM, the others through NVLink) in the same launch.
 */
#define FIXCA_MAX_FANOUT 8
FIXCA_API int fixca_cuda_region_dev_fanout(const void *d_src, size_t src_pitch, int src_row0, int src_rows,
					   void *const *d_dsts, int ndst, size_t dst_pitch, int dst_row0,
					   int width, int height, int bytes, int bpc,
					   const fixca_params *params, int y1, int y2,
					   unsigned flags, void *stream);
#define FIXCA_IPC_HANDLE_BYTES 64
FIXCA_API int fixca_cuda_frame_alloc(size_t bytes, void **d_frame, unsigned char handle[FIXCA_IPC_HANDLE_BYTES]);
FIXCA_API int fixca_cuda_frame_open(const unsigned char handle[FIXCA_IPC_HANDLE_BYTES], void **d_frame);
FIXCA_API int fixca_cuda_frame_close(void *d_frame);
FIXCA_API int fixca_cuda_frame_free(void *d_frame);

/*
 * A stream of `nframes` equal-sized host frames (tight rows), same parameters:
 * frames are pipelined H2D / kernel / D2H over a ring of pinned staging
 * buffers on `device`.  src_frames[i] / dst_frames[i] are host pointers.
 */
FIXCA_API int fixca_cuda_frames(const unsigned char *const *src_frames, unsigned char *const *dst_frames,
				int nframes, int width, int height, int bytes, int bpc,
				const fixca_params *params, unsigned flags, int device);

/*
 * The same stream of host frames sharded by index over `ndev` GPUs of the box from one process (BASELINE
 * "batch stream of frames sharded across 8xB200"): device i of ndev takes frames i, i + ndev, ... through its own
 * ring and PCIe link on its own worker thread.  Frames are independent, so there is no exchange between
 * devices.  devices == NULL means 0..ndev-1.
 */
FIXCA_API int fixca_cuda_frames_multi(const unsigned char *const *src_frames, unsigned char *const *dst_frames,
				      int nframes, int width, int height, int bytes, int bpc,
				      const fixca_params *params, unsigned flags, const int *devices, int ndev);

/* ------------------------------------------------------------------------- */
/* Host-side logic (no GPU needed)                                            */
/* ------------------------------------------------------------------------- */
/* Inclusive range of source rows that output rows [y1,y2) read: row y itself
 * (green/alpha copy, fix-ca.c:1094-1098) plus the red/blue taps reachable
 * through the affine map (fix-ca.c:1105-1106, 1135-1158, 1204-1256). */
FIXCA_API int fixca_band_source_rows(int width, int height, const fixca_params *params,
				     int y1, int y2, int *src_lo, int *src_hi);

/* Contiguous split of rows [y1,y2) into nbands bands; band i is
 * [band_y1[i], band_y2[i]).  Arrays hold nbands entries. */
FIXCA_API int fixca_split_bands(int y1, int y2, int nbands, int *band_y1, int *band_y2);

/* The dialog's lens reset (fix-ca.c:427-428): a coordinate <= 0 or >= size
 * becomes round(size / 2) (integer division).  The README's "-1,-1 resets to
 * the image centre" behaviour, for callers that want it. */
FIXCA_API void fixca_resolve_lens(int width, int height, double *lens_x, double *lens_y);

/* run()'s non-interactive range check (fix-ca.c:279-295): FIXCA_OK or FIXCA_ERR_RANGE
 * / FIXCA_ERR_INTERP. */
FIXCA_API int fixca_check_params(const fixca_params *params);

/* color_size() (fix-ca.c:681-711): babl format name + bytes per pixel -> bpc code. */
FIXCA_API int fixca_color_size(const char *babl_format_name, int bytes_per_pixel);

/* Extension: color_size() with the reference's commented-out half-precision line
 * (fix-ca.c:692-693) enabled -- "R'G'B' half" etc. give bpc = -2, which every entry point above accepts:
 * samples are IEEE binary16, decoded exactly (`ret += *p`, :740-742), computed like float images and stored
 * with one rounding (`*p = d`, :768-770).  Everything else answers like fixca_color_size(). */
FIXCA_API int fixca_color_size_half(const char *babl_format_name, int bytes_per_pixel);

/* fixca_color_size_half() plus "u15" names (fix-ca.c:694-695) -> FIXCA_BPC_U15 (bytes per pixel 6 or 8). */
FIXCA_API int fixca_color_size_ext(const char *babl_format_name, int bytes_per_pixel);

/* Defaults of fix_ca_params_default (fix-ca.c:85-97). */
FIXCA_API void fixca_params_default(fixca_params *params);

/* ------------------------------------------------------------------------- */
/* Progress, errors, introspection                                            */
/* ------------------------------------------------------------------------- */
/* The reference's progress protocol (fix-ca.c:1022-1023, 1331-1332, 1335-1336):
 * init once, update((y-y1)/(y2-y1)) for every row with (y-y1) % 8 == 0, then
 * update(0.0).  kind: 0 = init, 1 = update. */
typedef void (*fixca_progress_fn)(int kind, double fraction, void *user);
/* The callback belongs to the calling thread: install it on the thread that makes the fixca_cuda_region*()
 * call (the plug-in has one). */
FIXCA_API void fixca_cuda_set_progress(fixca_progress_fn fn, void *user);

/*
 * Pinned host memory for the image buffers fix_ca() allocates with g_new (fix-ca.c:366-367; the preview's at
 * :648-649).  Buffers from here are moved by DMA without staging copies (100 MP RGB16: 13 ms per call instead
 * of 32 ms from pageable memory).  fixca_cuda_host_alloc returns NULL when no GPU is usable or the allocation
 * fails -- the caller keeps its own allocator as the fallback; free with fixca_cuda_host_free only.  Freed
 * buffers are pooled (page-locking costs more than the pass: <= 4 buffers, <= 4 GiB) and handed out again, which
 * is what the dialog's preview needs (two whole-image buffers per refresh); fixca_cuda_release() empties the pool.
 */
FIXCA_API void *fixca_cuda_host_alloc(size_t bytes);
FIXCA_API void  fixca_cuda_host_free(void *p);

/* Re-read the FIXCA_* tuning variables.  They are read once per process; tests and A/B
 * probes that change the environment afterwards call this.  Not for use while other threads are inside
 * the library. */
FIXCA_API void fixca_cuda_reload_tuning(void);

FIXCA_API const char *fixca_cuda_last_error(void);	/* thread-local text of the last failure */
FIXCA_API const char *fixca_strerror(int code);
FIXCA_API int  fixca_cuda_device_count(void);		/* 0 when no driver / no GPU */
/* Name of the kernel variant the last fixca_cuda_region*() call on this thread launched
 * ("tiled/cubic/f32/u16x3", "direct/...", "none/..."), and how many kernels it launched. */
FIXCA_API const char *fixca_cuda_last_kernel(void);
FIXCA_API long fixca_cuda_launch_count(void);		/* kernels launched by this library so far */
/* Wall-clock milliseconds the last fixca_cuda_region() / _ex() call on this thread took (uploads, kernels,
 * downloads): what a plug-in spends inside the call that replaces its row loop. */
FIXCA_API double fixca_cuda_last_call_ms(void);
FIXCA_API void fixca_cuda_release(void);		/* free cached device / pinned buffers */
FIXCA_API const char *fixca_version(void);

#ifdef __cplusplus
}
#endif
#endif /* FIXCA_CUDA_H */
