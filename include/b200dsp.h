/*
 * b200dsp.h — C ABI of libb200dsp.so: B200-native (sm_100a) implementations of FFmpeg's DSP hot paths.
 *
 * Every entry point replaces (and cites) one interface of the reference; paths are relative to the FFmpeg tree.
 * Two levels per path:
 *   drop-in level  : the reference's own signature, HOST pointers.  Used by un-modified callers and by the parity
 *                    tests; each call stages through the device (H2D, kernel, D2H) and is therefore latency-bound.
 *   batched level  : many frames / blocks / transforms per call, DEVICE pointers (or pinned host buffers for the
 *                    *_host variants, which pipeline copies and kernels).  This is where the throughput is.
 * There is no CPU fallback: every function fails with a negative B200_E* code if no CUDA device is usable.
 * No function retains caller pointers after it returns (stream-ordered work is finished or the documented
 * b200_device_sync() point applies).  Contexts are not thread-safe; use one per thread like the reference's.
 */
#ifndef B200DSP_H
#define B200DSP_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200DSP_ABI_VERSION 1

/* error codes: negative errno values, mirroring AVERROR(x) = -(x) (libavutil/error.h) */
#define B200_EINVAL   (-22)
#define B200_ENOMEM   (-12)
#define B200_ENOSYS   (-38)   /* AVERROR(ENOSYS): combination not implemented (the reference's PATCHWELCOME/ENOTSUP cases) */
#define B200_ENODEV   (-19)   /* no CUDA device / CUDA runtime failure */
#define B200_EEXTERNAL (-5)   /* CUDA error during execution (see b200_last_error) */

int         b200_abi_version(void);
const char *b200_last_error(void);          /* thread-local, human readable */

/* ------------------------------------------------------------------------------------------------ device
 * Plays the role of AVHWDeviceContext/AVCUDADeviceContext {CUcontext, CUstream}
 * (libavutil/hwcontext_cuda.h, libavutil/hwcontext_cuda.c:932-956): one device ordinal + one stream on which all
 * work of the contexts created from it is ordered.  `cu_stream` may be a caller-owned CUstream/cudaStream_t
 * (e.g. AVCUDADeviceContext.stream, or torch's current stream) or NULL to let the library create one. */
typedef struct B200Device B200Device;

int   b200_device_open(B200Device **dev, int ordinal, void *cu_stream);
void  b200_device_close(B200Device *dev);
int   b200_device_sync(B200Device *dev);
int   b200_device_ordinal(const B200Device *dev);
void *b200_device_stream(const B200Device *dev);
int   b200_device_sm_count(const B200Device *dev);
/* process-wide default device used by the pointer tables that have no context argument
 * (IDCTDSPContext, MECmpContext, H264QpelContext, HpelDSPContext).  Opened lazily on ordinal 0 if unset. */
int   b200_set_default_device(B200Device *dev);
/* memory helpers so that C callers need no CUDA headers */
void *b200_malloc_device(B200Device *dev, size_t bytes);
void  b200_free_device(B200Device *dev, void *p);
void *b200_malloc_host(size_t bytes);                 /* pinned */
void  b200_free_host(void *p);
int   b200_memcpy_h2d(B200Device *dev, void *dst_dev, const void *src_host, size_t bytes);  /* async on the stream */
int   b200_memcpy_d2h(B200Device *dev, void *dst_host, const void *src_dev, size_t bytes);  /* async on the stream */
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
uint64_t b200_launch_count(void);

/* ------------------------------------------------------------------------------------------------ libswscale
 * Replaces the legacy scaler for AV_PIX_FMT_YUV420P -> packed 8-bit RGB (rgb24, bgr24, rgba, bgra, argb, abgr) and
 * AV_PIX_FMT_YUV420P -> AV_PIX_FMT_YUV420P (scaling):
 *   sws_getContext / sws_init_context ....... libswscale/utils.c:1919,1884 (ff_sws_init_single_context :1137)
 *   sws_setColorspaceDetails ................ libswscale/utils.c:849
 *   sws_scale ............................... libswscale/swscale.h:583, swscale.c:1626
 *   SwsFunc (convert_unscaled / ff_swscale).. libswscale/swscale_internal.h:99-101, swscale.c:263
 * Flags keep the reference's values (libswscale/swscale.h:88-118).  Output is bit-identical to the reference's
 * C path (the one FATE pins with accurate_rnd+bitexact, and yuv2rgb_c_24_rgb without accurate_rnd).  */
#define B200_PIX_FMT_YUV420P 0     /* AV_PIX_FMT_YUV420P, libavutil/pixfmt.h */
#define B200_PIX_FMT_NV12    23    /* AV_PIX_FMT_NV12: source only; src[1] is the interleaved U,V plane, src[2] is ignored */
#define B200_PIX_FMT_NV21    24    /* AV_PIX_FMT_NV21: same with V,U */
#define B200_PIX_FMT_RGB24   2     /* AV_PIX_FMT_RGB24 */
#define B200_PIX_FMT_BGR24   3     /* AV_PIX_FMT_BGR24 */
#define B200_PIX_FMT_ARGB    25    /* AV_PIX_FMT_ARGB  (alpha byte = 255, the source has no alpha plane) */
#define B200_PIX_FMT_RGBA    26    /* AV_PIX_FMT_RGBA  */
#define B200_PIX_FMT_ABGR    27    /* AV_PIX_FMT_ABGR  */
#define B200_PIX_FMT_BGRA    28    /* AV_PIX_FMT_BGRA  */

#define B200_SWS_FAST_BILINEAR 0x1
#define B200_SWS_BILINEAR      0x2
#define B200_SWS_BICUBIC       0x4
#define B200_SWS_X             0x8
#define B200_SWS_POINT         0x10
#define B200_SWS_AREA          0x20
#define B200_SWS_BICUBLIN      0x40
#define B200_SWS_GAUSS         0x80     /* X, GAUSS, SINC, LANCZOS, SPLINE: default parameters (param[] = SWS_PARAM_DEFAULT) */
#define B200_SWS_SINC          0x100
#define B200_SWS_LANCZOS       0x200
#define B200_SWS_SPLINE        0x400
#define B200_SWS_FULL_CHR_H_INT 0x2000
#define B200_SWS_FULL_CHR_H_INP 0x4000   /* packed RGB source: chroma from every pixel instead of every other one */
#define B200_SWS_ACCURATE_RND  0x40000
#define B200_SWS_BITEXACT      0x80000

typedef struct B200SwsContext B200SwsContext;

/* like sws_getContext(); srcFilter/dstFilter/param are not supported (must be the defaults). NULL on failure.
 * Sources: yuv420p, nv12, nv21 (-> packed 8-bit RGB or yuv420p / nv12 / nv21) and rgb24 / bgr24 / rgba / bgra / argb / abgr (-> yuv420p /
 * nv12 / nv21, or, when the size changes and alpha is not carried from source to destination, -> packed 8-bit RGB; the input
 * readers of libswscale/input.c:264-393,1068-1172, hScale16To15_c, and the bgr24ToYv12Wrapper special converter,
 * libswscale/swscale_unscaled.c:2453-2457); a packed RGB source is passed as plane 0 (src[1], src[2] unused).
 * Destinations nv12 / nv21 (what NVENC reads) are yuv420p with the chroma planes interleaved: planarToNv12Wrapper
 * (libswscale/swscale_unscaled.c:147-165) unscaled, yuv2nv12cX_c (libswscale/output.c:495-528) through the scaler; dst[1] is the
 * interleaved plane, dst[2] unused. */
B200SwsContext *b200_sws_getContext(B200Device *dev, int srcW, int srcH, int srcFormat,
                                    int dstW, int dstH, int dstFormat, int flags);
/* The same with SwsContext.src_range / .dst_range (libswscale/swscale.h, AVOption "src_range" / "dst_range"; 0 = limited,
 * 1 = full) set before sws_init_context(), the way vf_scale's in_range / out_range reach the scaler (libswscale/utils.c:
 * 1164-1167).  With a yuv420p destination and different ranges the 15-bit lines are range-converted between the horizontal
 * and the vertical pass (lumRangeToJpeg_c ... chrRangeFromJpeg_c, libswscale/swscale.c:163-209, :577-660) and the same-size
 * case goes through the scaler instead of planarCopyWrapper (utils.c:1623-1626).  For RGB destinations dstRange is ignored
 * and srcRange selects the look-up tables, as in the reference. */
B200SwsContext *b200_sws_getContext_range(B200Device *dev, int srcW, int srcH, int srcFormat, int srcRange,
                                          int dstW, int dstH, int dstFormat, int dstRange, int flags);
/* The same with SwsContext.scaler_params — the `param` argument of sws_getContext() (libswscale/swscale.h; utils.c:312-360):
 * bicubic B / C (default 0 / 0.6), Gaussian exponent (3), Lanczos width (3), the experimental scaler's power (1).  NULL, or
 * 123456 (SWS_PARAM_DEFAULT) in either slot, selects the default. */
B200SwsContext *b200_sws_getContext_params(B200Device *dev, int srcW, int srcH, int srcFormat, int srcRange,
                                           int dstW, int dstH, int dstFormat, int dstRange, int flags, const double *param);
/* sws_getContext's srcFilter / dstFilter (SwsFilter / SwsVector, libswscale/swscale.h:187-204): initFilter convolves the source vector of
 * each direction into that direction's filter bank; of a destination vector only the length counts (the reference never applies it,
 * "FIXME dstFilter", utils.c:384-413); any vector longer than one tap rules out the unscaled special converters (utils.c:1256-1263).
 * NULL pointers = no filter.  The vectors are copied. */
typedef struct B200SwsVector { const double *coeff; int length; } B200SwsVector;
typedef struct B200SwsFilter { const B200SwsVector *lumH, *lumV, *chrH, *chrV; } B200SwsFilter;
B200SwsContext *b200_sws_getContext_filters(B200Device *dev, int srcW, int srcH, int srcFormat, int srcRange, int dstW, int dstH, int dstFormat,
                                            int dstRange, int flags, const B200SwsFilter *srcFilter, const B200SwsFilter *dstFilter,
                                            const double *param);
void b200_sws_freeContext(B200SwsContext *c);
/* like sws_setColorspaceDetails() (libswscale/utils.c:849-1004).  RGB destination: `table` / dstRange are accepted and
 * ignored like the reference.  yuv420p destination: the ranges are stored and the range conversion is re-selected
 * (brightness / contrast / saturation do not apply); a context that was initialised as a plain copy (same size, equal
 * ranges) stays one, as in the reference, where convert_unscaled is chosen once at init.  inv_table != table for
 * yuv -> yuv makes the reference cascade through bgr24: not implemented, B200_ENOSYS. */
int  b200_sws_setColorspaceDetails(B200SwsContext *c, const int inv_table[4], int srcRange,
                                   const int table[4], int dstRange, int brightness, int contrast, int saturation);
/* drop-in for sws_scale(): HOST pointers, strides in bytes (negative allowed), returns output lines.
 * Top-down slice sequences are supported (bands uploaded into a device copy of the picture, lines emitted as soon as
 * their vertical taps are complete, same return values as the reference) from planar / semi-planar and packed RGB sources, into packed RGB and
 * into yuv420p / nv12 / nv21 destinations; a sequence whose first band touches the last line runs bottom-up (the picture is flipped internally like
 * scale_internal, swscale.c:1096-1159; even heights, yuv sources). */
int  b200_sws_scale(B200SwsContext *c, const uint8_t *const srcSlice[], const int srcStride[],
                    int srcSliceY, int srcSliceH, uint8_t *const dst[], const int dstStride[]);
/* SwsFunc-shaped entry (first argument is the context): what a maintainer installs as convert_unscaled */
int  b200_sws_func(void *c, const uint8_t *const src[], const int srcStride[], int srcSliceY, int srcSliceH,
                   uint8_t *const dst[], const int dstStride[]);
/* batched, DEVICE pointers: frame f's plane p starts at src[p] + f*srcFrameStride[p]; rgb at dst + f*dstFrameStride.
 * Asynchronous on the device's stream. */
int  b200_sws_scale_batch_device(B200SwsContext *c, const uint8_t *const src[3], const int srcStride[3],
                                 const int64_t srcFrameStride[3], uint8_t *dst, int dstStride,
                                 int64_t dstFrameStride, int nframes);
/* yuv420p -> yuv420p (context created with dstFormat = B200_PIX_FMT_YUV420P; the reference's yuv2planeX_8_c / yuv2plane1_8_c
 * writers, libswscale/output.c:468-493, and planarCopyWrapper at same size): three destination planes.
 * b200_sws_scale() takes them as dst[0..2] / dstStride[0..2] like the reference; the single-plane batch entry points
 * return B200_EINVAL for such a context and this one returns B200_EINVAL for a packed-RGB context. */
int  b200_sws_scale_batch_device_planar(B200SwsContext *c, const uint8_t *const src[3], const int srcStride[3],
                                        const int64_t srcFrameStride[3], uint8_t *const dst[3], const int dstStride[3],
                                        const int64_t dstFrameStride[3], int nframes);
/* batched, HOST pointers (pinned memory recommended): chunks of frames are copied in, converted and copied back
 * on rotating streams so that H2D, kernels and D2H overlap.  Synchronous: returns when dst is complete. */
int  b200_sws_scale_batch_host(B200SwsContext *c, const uint8_t *const src[3], const int srcStride[3],
                               const int64_t srcFrameStride[3], uint8_t *dst, int dstStride,
                               int64_t dstFrameStride, int nframes);
/* introspection for tests: same layout as the shim used on the reference (16 ints) */
int  b200_sws_info(const B200SwsContext *c, int *out16);
/* diagnostics: which kernels the scaled-path launches since the previous call used (bit 0 two passes through 15-bit line planes,
 * bit 1 fused CUDA-core kernels, bit 2 fused kernels with the tensor-core horizontal pass); reading clears it. */
int  b200_sws_last_path(B200SwsContext *c);
/* which: 0 hLum 1 hChr 2 vLum 3 vChr — host copies of the generated filter tables */
int  b200_sws_get_filter(const B200SwsContext *c, int which, int16_t *filter, int32_t *pos, int cap);

/* host-only: build the set-up tables for a configuration WITHOUT touching a GPU (used by the CPU test tier to check
 * the filter generation against the reference).  info16 as b200_sws_info; filter/pos may be NULL. Returns n or <0. */
int  b200_sws_plan_probe(int srcW, int srcH, int dstW, int dstH, int flags, int which,
                         int16_t *filter, int32_t *pos, int cap, int *info16);
/* the general form: cfg = { srcW, srcH, srcFormat, srcRange, dstW, dstH, dstFormat, dstRange, flags }; details = NULL or the 13
 * ints of a sws_setColorspaceDetails() call made after initialisation { inv_table[4], srcRange, table[4], dstRange,
 * brightness, contrast, saturation }.  info32: [0..15] as b200_sws_info ([12] = vertical chroma shift of the destination),
 * [16] plain-copy context, [17] range conversion (0 none, 1 limited->full, 2 full->limited), [18..21] its luma coefficient,
 * luma offset, chroma coefficient, chroma offset (libswscale/swscale.c:577-624), [22] fast-bilinear horizontal pass,
 * [23] semi-planar source kind, [24] return value of the details call, [25] src_range, [26] dst_range, [27] packed RGB source
 * (bytes per pixel), [28] / [29] horizontal / vertical chroma shift of the source as the scaler sees it, [30] bgr24 -> yv12
 * converter installed, [31] semi-planar destination kind (1 nv12, 2 nv21), [32..40] input_rgb2yuv_table.  The array must hold 48 ints. */
int  b200_sws_plan_probe2(const int cfg[9], const int *details, int which, int16_t *filter, int32_t *pos, int cap, int *info48);
/* host-only: a horizontal filter bank (n outputs x size taps, first-tap positions) regrouped into the operands of the tensor-core
 * horizontal pass (csrc/sws_mma.cuh): ginfo = (ceil(n / 8) + 1) pairs { first source column of the group's window, first chunk index },
 * bfrag = 128 words per chunk (32 lanes x { hi b0, hi b1, lo b0, lo b1 } of mma.m16n8k32's B operand), pitch[2] = staged-line byte
 * pitches for 128- and 64-column tiles.  Returns the number of bfrag words or <0.  Replaces nothing in the reference: the tables
 * hold libswscale's hLumFilter / hLumFilterPos (swscale_internal.h:437-448) in another order. */
int  b200_sws_mma_probe(const int16_t *coef, const int32_t *pos, int n, int size, int32_t *ginfo, int ginfo_cap,
                        uint32_t *bfrag, int bfrag_cap, int *pitch);

/* ------------------------------------------------------------------------------------------------ idctdsp
 * Replaces IDCTDSPContext (libavcodec/idctdsp.h:43-91) as filled by ff_idctdsp_init (libavcodec/idctdsp.c:228-314)
 * for idct_algo = FF_IDCT_SIMPLE / AUTO, 8 bit: ff_simple_idct_{put,add,}_int16_8bit
 * (libavcodec/simple_idct_template.c:329-368) and the clamp helpers (libavcodec/idctdsp.c:73-165).
 * The struct below has the reference's member order so that a maintainer can memcpy / alias it. */
typedef struct B200IDCTDSPContext {
    void (*put_pixels_clamped)(const int16_t *block, uint8_t *pixels, ptrdiff_t line_size);
    void (*put_signed_pixels_clamped)(const int16_t *block, uint8_t *pixels, ptrdiff_t line_size);
    void (*add_pixels_clamped)(const int16_t *block, uint8_t *pixels, ptrdiff_t line_size);
    void (*idct)(int16_t *block);
    void (*idct_put)(uint8_t *dest, ptrdiff_t line_size, int16_t *block);
    void (*idct_add)(uint8_t *dest, ptrdiff_t line_size, int16_t *block);
    uint8_t idct_permutation[64];
    int perm_type;                 /* enum idct_permutation_type: FF_IDCT_PERM_NONE = 0 */
    int mpeg4_studio_profile;
} B200IDCTDSPContext;

/* like ff_idctdsp_init(c, avctx) with avctx->idct_algo / bits_per_raw_sample / lowres passed explicitly.
 * Returns 0, or B200_ENOSYS for algorithms other than simple/auto at 8 bit, lowres 0. */
int  b200_idctdsp_init(B200IDCTDSPContext *c, int idct_algo, int bits_per_raw_sample, int lowres);

#define B200_IDCT      0   /* in place on the coefficient blocks */
#define B200_IDCT_PUT  1
#define B200_IDCT_ADD  2
/* batched, DEVICE pointers.  blocks: nblocks x int16[64] natural order (FF_IDCT_PERM_NONE), 16-byte aligned.
 * Block i goes to dest + dest_off[i] with line size line_size[i] (line_size NULL -> uniform_line_size).
 * For B200_IDCT the result replaces the coefficients and dest/dest_off are ignored.
 * Unlike the reference the coefficient blocks are NOT clobbered by put/add. */
int  b200_idct_batch_device(B200Device *dev, int kind, int16_t *blocks, int64_t nblocks, uint8_t *dest,
                            const int64_t *dest_off, const int32_t *line_size, int uniform_line_size);
/* batched macroblock stream, DEVICE pointers: nframes frames of mb_w x mb_h 4:2:0 macroblocks, 6 blocks each in the
 * decoder's order Y0 Y1 Y2 Y3 Cb Cr (libavcodec/mpegvideo_dec.c:940-1128 put_dct/add_dct calls); destinations are
 * implied: luma (16*mbx + 8*(b&1), 16*mby + 8*(b>>1)), chroma (8*mbx, 8*mby).  Planes of frame f start at
 * plane + f*frame_stride[p]. */
int  b200_idct_mb420_device(B200Device *dev, int kind, const int16_t *blocks, int mb_w, int mb_h, int nframes,
                            uint8_t *const planes[3], const int linesize[3], const int64_t frame_stride[3]);
/* same from HOST memory (copies in/out, pipelined). */
int  b200_idct_mb420_host(B200Device *dev, int kind, const int16_t *blocks, int mb_w, int mb_h, int nframes,
                          uint8_t *const planes[3], const int linesize[3], const int64_t frame_stride[3]);

/* The same table for bits_per_raw_sample 9 / 10 / 12: ff_simple_idct_{,put_,add_}int16_10bit resp. _12bit
 * (libavcodec/idctdsp.c:248-266, libavcodec/simple_idct_template.c:63-104,329-368); uint16 pixels, line sizes in bytes.  Kept as a
 * separate entry until it has run on hardware, then b200_idctdsp_init() takes the depth.  The clamp helpers stay the 8-bit ones
 * like the reference's.  As with the 8-bit batch call the coefficient blocks are NOT clobbered by put / add. */
int  b200_idctdsp_init_hbd(B200IDCTDSPContext *c, int idct_algo, int bits_per_raw_sample, int lowres);
/* batched, DEVICE pointers: as b200_idct_batch_device with depth 10 (9 maps to it) or 12; dest_off / line sizes in bytes, even */
int  b200_idct_hbd_batch_device(B200Device *dev, int depth, int kind, int16_t *blocks, int64_t nblocks, uint8_t *dest,
                                const int64_t *dest_off, const int32_t *line_size, int uniform_line_size);

/* ProresDSPContext (libavcodec/proresdsp.h:28-35) as ff_proresdsp_init(dsp, bits_per_raw_sample) fills it for 10 and 12 bit
 * (libavcodec/proresdsp.c:56-82,102-194): idct_put = dequantise by qmat + inverse transform + bias + clip to [4, 2^bits - 5], the
 * fused "dequant + IDCT" of the ProRes decoder (libavcodec/proresdec.c:559-598).  Same member order as the reference struct;
 * idct_put_bayer (ProRes RAW) is left NULL. */
typedef struct B200ProresDSPContext {
    int idct_permutation_type;             /* FF_IDCT_PERM_NONE */
    uint8_t idct_permutation[64];
    void (*idct_put)(uint16_t *out, ptrdiff_t linesize, int16_t *block, const int16_t *qmat);     /* HOST pointers; block is not clobbered */
    void (*idct_put_bayer)(uint16_t *out, ptrdiff_t linesize, int32_t *block, const int16_t *qmat, const uint16_t *lin_curve);
} B200ProresDSPContext;
int  b200_proresdsp_init(B200ProresDSPContext *c, int bits_per_raw_sample);
/* batched, DEVICE pointers: block i = blocks + 64*i goes, dequantised by qmat[64] (device), to dest + dest_off[i] (bytes; uint16 pixels)
 * with line size line_size[i] or uniform_line_size (bytes, even) */
int  b200_prores_idct_put_batch_device(B200Device *dev, int bits, const int16_t *blocks, int64_t nblocks, const int16_t *qmat,
                                       uint8_t *dest, const int64_t *dest_off, const int32_t *line_size, int uniform_line_size);

/* H.264 in-loop deblocking, 8 bit: the loop-filter members of H264DSPContext (libavcodec/h264dsp.h:48-73) in the reference's order,
 * as ff_h264dsp_init(c, 8, chroma_format_idc) installs them (libavcodec/h264dsp.c:109-132; functions libavcodec/h264dsp_template.c:
 * 103-340).  pix points at the first pixel on the far side of the edge, like the reference; HOST pointers. */
typedef struct B200H264LoopFilterContext {
    void (*v_loop_filter_luma)(uint8_t *pix, ptrdiff_t stride, int alpha, int beta, int8_t *tc0);
    void (*h_loop_filter_luma)(uint8_t *pix, ptrdiff_t stride, int alpha, int beta, int8_t *tc0);
    void (*h_loop_filter_luma_mbaff)(uint8_t *pix, ptrdiff_t stride, int alpha, int beta, int8_t *tc0);
    void (*v_loop_filter_luma_intra)(uint8_t *pix, ptrdiff_t stride, int alpha, int beta);
    void (*h_loop_filter_luma_intra)(uint8_t *pix, ptrdiff_t stride, int alpha, int beta);
    void (*h_loop_filter_luma_mbaff_intra)(uint8_t *pix, ptrdiff_t stride, int alpha, int beta);
    void (*v_loop_filter_chroma)(uint8_t *pix, ptrdiff_t stride, int alpha, int beta, int8_t *tc0);
    void (*h_loop_filter_chroma)(uint8_t *pix, ptrdiff_t stride, int alpha, int beta, int8_t *tc0);
    void (*h_loop_filter_chroma_mbaff)(uint8_t *pix, ptrdiff_t stride, int alpha, int beta, int8_t *tc0);
    void (*v_loop_filter_chroma_intra)(uint8_t *pix, ptrdiff_t stride, int alpha, int beta);
    void (*h_loop_filter_chroma_intra)(uint8_t *pix, ptrdiff_t stride, int alpha, int beta);
    void (*h_loop_filter_chroma_mbaff_intra)(uint8_t *pix, ptrdiff_t stride, int alpha, int beta);
} B200H264LoopFilterContext;
int  b200_h264_loop_filter_init(B200H264LoopFilterContext *c, int bit_depth, int chroma_format_idc);     /* 8, or 9 / 10 / 12 / 14 (uint16 samples) */
/* kinds for the batched call: 0 v_luma, 1 h_luma, 2 h_luma_mbaff, 3-5 their _intra forms, 6 v_chroma, 7 h_chroma, 8 h_chroma_mbaff,
 * 9-11 their _intra forms, 12 / 13 h_chroma / h_chroma_mbaff of 4:2:2 content, 14 / 15 their _intra forms */
/* batched, DEVICE pointers: edge e of kind kinds[e] at pix + pix_off[e] with alpha[e], beta[e] and tc0[4*e .. 4*e+3] (ignored by the
 * intra kinds).  The edges of one call must not read or write each other's pixels (p3..q3 across, the edge's lines along): the
 * decoder's sequential order becomes one call per independent set of edges. */
int  b200_h264_loop_filter_batch_device(B200Device *dev, int64_t nedges, const uint8_t *kinds, uint8_t *pix, const int64_t *pix_off,
                                        ptrdiff_t stride, const uint8_t *alpha, const uint8_t *beta, const int8_t *tc0);
/* the same for 9 / 10 / 12 / 14 bit samples (uint16): pix_off and stride in BYTES; alpha, beta and tc0 are the 8-bit table values the
 * reference's callers pass (the functions scale them to the sample depth, h264dsp_template.c:109-112,241-246) */
int  b200_h264_loop_filter_hbd_batch_device(B200Device *dev, int bit_depth, int64_t nedges, const uint8_t *kinds, uint8_t *pix,
                                            const int64_t *pix_off, ptrdiff_t stride, const uint8_t *alpha, const uint8_t *beta, const int8_t *tc0);

/* mpegvideo inverse quantisers: the members of MPVUnquantDSPContext (libavcodec/mpegvideo_unquantize.h:31-44) as
 * ff_mpv_unquantize_init() installs them (libavcodec/mpegvideo_unquantize.c:50-290), i.e. what runs in front of the IDCT in
 * mpv_reconstruct_mb's put_dct / add_dequant_dct (libavcodec/mpegvideo_dec.c).  The reference functions take the whole
 * MPVContext; B200MpvUnquant carries the fields they read (mpegvideo.h:70-77 y/c_dc_scale, ac_pred, h263_aic, the scan
 * tables; :201-203 intra_matrix / inter_matrix; :258 q_scale_type).  variant 3 is the function the reference installs for
 * AV_CODEC_FLAG_BITEXACT (with mismatch control); variant 2 the plain one. */
#define B200_UNQUANT_MPEG1_INTRA          0
#define B200_UNQUANT_MPEG1_INTER          1
#define B200_UNQUANT_MPEG2_INTRA          2
#define B200_UNQUANT_MPEG2_INTRA_BITEXACT 3
#define B200_UNQUANT_MPEG2_INTER          4
#define B200_UNQUANT_H263_INTRA           5
#define B200_UNQUANT_H263_INTER           6
typedef struct B200MpvUnquant {
    uint16_t intra_matrix[64], inter_matrix[64];     /* raster order after the IDCT permutation, as MPVContext holds them */
    uint8_t  permutated[64], raster_end[64];         /* ScanTable (ff_init_scantable, mpegvideo_unquantize.c:36-48) of the scan in use */
    int32_t  y_dc_scale, c_dc_scale, q_scale_type, h263_aic, ac_pred;
} B200MpvUnquant;
/* batched, DEVICE pointers, in place: block i = blocks + 64*i (int16, 4-byte aligned), its number inside the macroblock
 * blk_n[i] (0-3 luma, 4+ chroma: selects the DC scale; NULL = macroblock stream order, i % 6), quantiser qscale[i] and
 * block_last_index last_index[i] (-1 = no coded coefficient; the H.263 variants need >= 0 like the reference asserts).
 * Coefficients beyond the coded part of the scan are left untouched, as by the reference. */
int  b200_mpv_unquantize_batch_device(B200Device *dev, int variant, const B200MpvUnquant *p, int16_t *blocks, int64_t nblocks,
                                      const uint8_t *blk_n, const uint8_t *qscale, const int8_t *last_index);

/* put_dct / add_dequant_dct of mpv_reconstruct_mb (libavcodec/mpegvideo_dec.c:907-922) for a whole 4:2:0 macroblock stream in one
 * kernel: inverse quantisation (variant, p, qscale[], last_index[] per block as in b200_mpv_unquantize_batch_device, block number
 * inside the macroblock = position in the stream % 6) straight into the simple IDCT put / add of b200_idct_mb420_device — the
 * dequantised coefficients stay in registers.  kind = B200_IDCT_ADD skips blocks whose last_index is < 0, like add_dequant_dct.
 * `blocks` is not modified.  Planes as for b200_idct_mb420_device, 16-byte (luma) / 8-byte (chroma) aligned.  Asynchronous on
 * the device's stream. */
int  b200_mpv_unquant_idct_mb420_device(B200Device *dev, int variant, const B200MpvUnquant *p, int kind, const int16_t *blocks,
                                        const uint8_t *qscale, const int8_t *last_index, int mb_w, int mb_h, int nframes,
                                        uint8_t *const planes[3], const int linesize[3], const int64_t frame_stride[3]);

/* ------------------------------------------------------------------------------------------------ float_dsp
 * Replaces AVFloatDSPContext (libavutil/float_dsp.h:24-210) as avpriv_float_dsp_alloc() fills it with the C functions
 * (libavutil/float_dsp.c:27-141, float_scalarproduct.c:25-33): the element-wise float work around the transforms, e.g.
 * vector_fmul_window after every iMDCT of the AAC decoder (libavcodec/aac/aacdec_dsp_template.c).  Same member order as the
 * reference struct.  The table entries take HOST pointers (copy in, kernel, copy out). */
typedef struct B200FloatDSPContext {
    void   (*vector_fmul)(float *dst, const float *src0, const float *src1, int len);
    void   (*vector_fmac_scalar)(float *dst, const float *src, float mul, int len);
    void   (*vector_dmac_scalar)(double *dst, const double *src, double mul, int len);
    void   (*vector_fmul_scalar)(float *dst, const float *src, float mul, int len);
    void   (*vector_dmul_scalar)(double *dst, const double *src, double mul, int len);
    void   (*vector_fmul_window)(float *dst, const float *src0, const float *src1, const float *win, int len);
    void   (*vector_fmul_add)(float *dst, const float *src0, const float *src1, const float *src2, int len);
    void   (*vector_fmul_reverse)(float *dst, const float *src0, const float *src1, int len);
    void   (*butterflies_float)(float *v1, float *v2, int len);
    float  (*scalarproduct_float)(const float *v1, const float *v2, int len);
    void   (*vector_dmul)(double *dst, const double *src0, const double *src1, int len);
    double (*scalarproduct_double)(const double *v1, const double *v2, size_t len);
} B200FloatDSPContext;
int  b200_float_dsp_init(B200FloatDSPContext *c);        /* like avpriv_float_dsp_alloc(), into caller storage */
/* ops, numbered in the struct's member order */
#define B200_FDSP_VECTOR_FMUL          0
#define B200_FDSP_VECTOR_FMAC_SCALAR   1
#define B200_FDSP_VECTOR_DMAC_SCALAR   2
#define B200_FDSP_VECTOR_FMUL_SCALAR   3
#define B200_FDSP_VECTOR_DMUL_SCALAR   4
#define B200_FDSP_VECTOR_FMUL_WINDOW   5
#define B200_FDSP_VECTOR_FMUL_ADD      6
#define B200_FDSP_VECTOR_FMUL_REVERSE  7
#define B200_FDSP_BUTTERFLIES_FLOAT    8
#define B200_FDSP_SCALARPRODUCT_FLOAT  9
#define B200_FDSP_VECTOR_DMUL          10
#define B200_FDSP_SCALARPRODUCT_DOUBLE 11
/* batched, DEVICE pointers: nvec vectors of len elements (float, or double for the D ops); vector v of each operand starts at
 * base + v*stride (strides in elements; 0 = one vector shared by all, e.g. the window).  Operand roles as in the reference:
 * vector_fmul_window: src2 = window of 2*len, dst gets 2*len; butterflies_float: dst = v1, src0 = v2, both rewritten;
 * the scalar products put one value per vector at dst[v*dst_stride] and, like the C loops, sum left to right (one thread per
 * vector: exact, not fast).  fmac / butterflies read dst; otherwise dst may alias an input only element for element. */
int  b200_float_dsp_batch_device(B200Device *dev, int op, int64_t nvec, int len, void *dst, int64_t dst_stride,
                                 const void *src0, int64_t src0_stride, const void *src1, int64_t src1_stride,
                                 const void *src2, int64_t src2_stride, double mul);

/* H.264 residual transforms, 8 bit: the IDCT members of H264DSPContext (libavcodec/h264dsp.h:81-88) as installed by
 * ff_h264dsp_init(c, 8, chroma_format_idc) (libavcodec/h264dsp.c:66-139): ff_h264_idct_add_8_c, ff_h264_idct8_add_8_c,
 * ff_h264_idct_dc_add_8_c, ff_h264_idct8_dc_add_8_c (libavcodec/h264idct_template.c:33-181).  The add16/add8/add4 wrappers
 * of the reference only dispatch to these four per 4x4/8x8 block and keep working on top of them. */
typedef void (*b200_h264_idct_fn)(uint8_t *dst, int16_t *block, ptrdiff_t stride);   /* block is cleared, like the reference */
typedef struct B200H264IDCTContext {
    b200_h264_idct_fn idct_add;        /* 4x4, dst 4-aligned */
    b200_h264_idct_fn idct8_add;       /* 8x8, dst 8-aligned */
    b200_h264_idct_fn idct_dc_add;     /* 4x4, only block[0] is read (and cleared) */
    b200_h264_idct_fn idct8_dc_add;
} B200H264IDCTContext;
int  b200_h264_idct_init(B200H264IDCTContext *c, int bit_depth, int chroma_format_idc);   /* 8, or 9 / 10 / 12 / 14: uint16 samples and
                                                                                              * int32 coefficients behind the int16_t * argument */
#define B200_H264_IDCT4    0
#define B200_H264_IDCT8    1
#define B200_H264_IDCT4_DC 2
#define B200_H264_IDCT8_DC 3
/* batched, DEVICE pointers, one kind per call: block i = blocks + blk_off[i] (int16 elements, 16-byte aligned blocks of
 * 16 or 64 coefficients; DC kinds touch block[0] only) is transformed, added to the 4x4 / 8x8 pixels at dst + dst_off[i]
 * and cleared.  dst + dst_off[i] and stride must be 4-aligned (8 for the 8x8 kinds), as the reference requires. */
int  b200_h264_idct_batch_device(B200Device *dev, int kind, int64_t n, int16_t *blocks, const int64_t *blk_off,
                                 uint8_t *dst, const int64_t *dst_off, ptrdiff_t stride);
/* the same for 9 / 10 / 12 / 14 bit samples: int32 coefficients (blk_off in int32 elements), uint16 samples (dst_off and stride in BYTES) */
int  b200_h264_idct_hbd_batch_device(B200Device *dev, int bit_depth, int kind, int64_t n, int32_t *blocks, const int64_t *blk_off,
                                     uint8_t *dst, const int64_t *dst_off, ptrdiff_t stride);

/* H.264 explicit weighted prediction, 8 bit: weight_pixels_tab / biweight_pixels_tab of H264DSPContext (libavcodec/h264dsp.h:33-45)
 * as installed by ff_h264dsp_init(c, 8, ...) (libavcodec/h264dsp.c:103-110; functions in h264dsp_template.c:30-99).
 * Index = width: [0] 16, [1] 8, [2] 4, [3] 2. */
typedef void (*b200_h264_weight_func)(uint8_t *block, ptrdiff_t stride, int height, int log2_denom, int weight, int offset);
typedef void (*b200_h264_biweight_func)(uint8_t *dst, uint8_t *src, ptrdiff_t stride, int height, int log2_denom,
                                        int weightd, int weights, int offset);
typedef struct B200H264WeightContext {
    b200_h264_weight_func   weight_pixels_tab[4];
    b200_h264_biweight_func biweight_pixels_tab[4];
} B200H264WeightContext;
int  b200_h264_weight_init(B200H264WeightContext *c, int bit_depth);            /* 8, or 9 / 10 / 12 / 14 (uint16 samples, stride in bytes) */
/* batched, DEVICE pointers.  params: 4 int32 per block (16-byte aligned array): [0] = width index | height << 8 | log2_denom << 16,
 * [1] = weight (the destination weight for biweight), [2] = source weight (biweight), [3] = offset.  src == NULL: weight
 * (in place on dst + dst_off[i]); src != NULL: biweight of dst + dst_off[i] with src + src_off[i].  One stride for both. */
int  b200_h264_weight_batch_device(B200Device *dev, int64_t n, const int32_t *params, uint8_t *dst, const int64_t *dst_off,
                                   const uint8_t *src, const int64_t *src_off, ptrdiff_t stride);
/* the same for 9 / 10 / 12 / 14 bit samples (uint16): offsets and stride in BYTES (even) */
int  b200_h264_weight_hbd_batch_device(B200Device *dev, int bit_depth, int64_t n, const int32_t *params, uint8_t *dst, const int64_t *dst_off,
                                       const uint8_t *src, const int64_t *src_off, ptrdiff_t stride);

/* ------------------------------------------------------------------------------------------------ fdctdsp
 * Replaces FDCTDSPContext (libavcodec/fdctdsp.h:28-31) as filled by ff_fdctdsp_init(c, avctx) (libavcodec/fdctdsp.c:27-45): the forward
 * 8x8 DCT in place on 64 int16 (16-byte aligned), fdct248 = the 2-4-8 variant for interlaced blocks.  dct_algo / bits_per_raw_sample are
 * AVCodecContext's: 9 / 10 bit -> ff_jpeg_fdct_islow_10, else FF_DCT_FASTINT (1) -> ff_fdct_ifast, FF_DCT_FAAN (6) -> ENOSYS (keep the C
 * functions), anything else -> ff_jpeg_fdct_islow_8. */
typedef struct B200FDCTDSPContext {
    void (*fdct)(int16_t *block);       /* HOST pointer */
    void (*fdct248)(int16_t *block);
} B200FDCTDSPContext;
int  b200_fdctdsp_init(B200FDCTDSPContext *c, int dct_algo, int bits_per_raw_sample);
/* batched, DEVICE pointer: n blocks of 64 coefficients transformed in place (blocks 16-byte aligned) */
int  b200_fdct_batch_device(B200Device *dev, int dct_algo, int bits_per_raw_sample, int is248, int16_t *blocks, int64_t n);

/* ------------------------------------------------------------------------------------------------ me_cmp
 * Replaces MECmpContext (libavcodec/me_cmp.h:53-77) as filled by ff_me_cmp_init (libavcodec/me_cmp.c:961-1027) for the
 * SAD / SSE entries: sad[0..1] = pix_abs16_c / pix_abs8_c, sse[0..2] = sse16_c / sse8_c / sse4_c,
 * pix_abs[0..1][0..3] = full / x2 / y2 / xy2 (me_cmp.c:37-385), and hadamard8_diff, vsad, vsse, nsse, median_sad, dct_sad, dct_max,
 * dct264_sad ([0] 16 wide, [1] 8 wide; [4] / [5] the intra forms where upstream has them).  Entries this library does not implement
 * (quant_psnr, bit, rd: they run the encoder's quantiser and VLC tables; w53, w97: the snow wavelets) stay NULL.
 * me_cmp_func (me_cmp.h:47-51): the first argument (MPVEncContext *) is unused by these functions and may be NULL. */
typedef int (*b200_me_cmp_func)(void *c, const uint8_t *blk1, const uint8_t *blk2, ptrdiff_t stride, int h);
typedef struct B200MECmpContext {
    int (*sum_abs_dctelem)(const int16_t *block);
    b200_me_cmp_func sad[6];
    b200_me_cmp_func sse[6];
    b200_me_cmp_func hadamard8_diff[6];
    b200_me_cmp_func dct_sad[6];
    b200_me_cmp_func quant_psnr[6];
    b200_me_cmp_func bit[6];
    b200_me_cmp_func rd[6];
    b200_me_cmp_func vsad[6];
    b200_me_cmp_func vsse[6];
    b200_me_cmp_func nsse[6];
    b200_me_cmp_func w53[6];
    b200_me_cmp_func w97[6];
    b200_me_cmp_func dct_max[6];
    b200_me_cmp_func dct264_sad[6];
    b200_me_cmp_func pix_abs[2][4];
    b200_me_cmp_func median_sad[6];
} B200MECmpContext;
/* like ff_me_cmp_init(c, avctx); avctx->flags passed explicitly (only AV_CODEC_FLAG_BITEXACT matters upstream) */
int  b200_me_cmp_init(B200MECmpContext *c, int codec_flags);

#define B200_MECMP_SAD     0   /* idx 0: 16 wide, 1: 8 wide */
#define B200_MECMP_SSE     1   /* idx 0: 16, 1: 8, 2: 4 wide */
#define B200_MECMP_PIX_ABS 2   /* idx = 4*size(0:16,1:8) + (0 full, 1 x2, 2 y2, 3 xy2) */
#define B200_MECMP_HADAMARD8 3 /* hadamard8_diff (SATD, me_cmp.c:514-562,933-950): idx 0: 16 wide (h = 8 or 16), 1: 8x8 (h ignored);
                                * idx 4 / 5: hadamard8_intra16 / intra8x8 (me_cmp.c:564-612: the block blk1 itself, mean left out) */
#define B200_MECMP_VSAD    4   /* vsad[idx] (me_cmp.c:843-881): idx 0: 16 wide, 1: 8, 4: vsad_intra16, 5: vsad_intra8 */
#define B200_MECMP_VSSE    5   /* vsse[idx] (me_cmp.c:883-931): same indices */
#define B200_MECMP_NSSE    6   /* nsse[idx] (me_cmp.c:387-437): idx 0: 16, 1: 8; weight as set by b200_me_cmp_set_nsse_weight (default 8) */
#define B200_MECMP_MEDIAN_SAD 7 /* median_sad[idx] (me_cmp.c:145-183,292-330): idx 0: 16, 1: 8 */
#define B200_MECMP_DCT_SAD  8   /* dct_sad[idx] (me_cmp.c:614-622,952): sum |fdct(blk1 - blk2)|; idx 0: 16 wide (h = 8 or 16), 1: 8x8 */
#define B200_MECMP_DCT_MAX  9   /* dct_max[idx] (me_cmp.c:678-693,956): max |fdct(blk1 - blk2)| per 8x8 block, added up over the blocks */
#define B200_MECMP_DCT264_SAD 10 /* dct264_sad[idx] (me_cmp.c:624-675,954): the H.264 8x8 integer transform; upstream only in GPL builds */
/* dct_sad / dct_max run FDCTDSPContext.fdct of the encoder context they are handed (s->fdsp.fdct); the table entries cannot read it, so
 * the DCT is chosen here with AVCodecContext.dct_algo's values the way ff_fdctdsp_init does for 8-bit samples (fdctdsp.c:27-45):
 * FF_DCT_FASTINT (1) = ff_fdct_ifast, FF_DCT_FAAN (6) = not implemented (ENOSYS, setting unchanged), anything else = ff_jpeg_fdct_islow_8. */
int  b200_me_cmp_set_dct_algo(int dct_algo);
/* nsse multiplies its noise term by MPVEncContext.c.avctx->nsse_weight, or by 8 when the context argument is NULL (me_cmp.c:407-410).
 * The table entries cannot read the caller's context (its layout is private to the encoder): the weight is set here instead. */
void b200_me_cmp_set_nsse_weight(int weight);
/* sum_abs_dctelem_c (me_cmp.c:105-112) for n blocks of 64 coefficients, DEVICE pointers */
int  b200_sum_abs_dctelem_batch_device(B200Device *dev, const int16_t *blocks, int64_t n, int32_t *out);
/* batched, DEVICE pointers: out[i] = fn(frame1 + off1[i], frame2 + off2[i], stride, h) */
int  b200_me_cmp_batch_device(B200Device *dev, int fn, int idx, const uint8_t *frame1, const uint8_t *frame2,
                              ptrdiff_t stride, int h, const int64_t *off1, const int64_t *off2, int64_t n, int32_t *out);
/* exhaustive search over whole frames, DEVICE pointers: ff_me_search_esa (libavfilter/motion_estimation.c:78-97) driven
 * like vf_mestimate.c:85-127 — for every mb_size x mb_size block of every frame pair: cost of the zero vector first,
 * raster scan of the +-search_param window clipped to the frame, strict '<' (first minimum in raster order wins, the
 * zero vector wins ties against everything).  mb_size in {4, 8, 16}.  out_mv: 2 ints (absolute x, y of the best match)
 * per block in raster order per frame; out_cost: its SAD. */
int  b200_me_esa_device(B200Device *dev, const uint8_t *cur, const uint8_t *ref, int linesize, int width, int height,
                        int64_t frame_stride, int nframes, int mb_size, int search_param, int32_t *out_mv, uint64_t *out_cost);
/* same on HOST buffers (vf_mestimate's frames live in host memory): chunks of frame pairs rotate over three streams, H2D of both luma
 * planes (frame_stride bytes per frame, frame_stride >= linesize * height), the search kernel, D2H of vectors and costs. */
int  b200_me_esa_host(B200Device *dev, const uint8_t *cur, const uint8_t *ref, int linesize, int width, int height,
                      int64_t frame_stride, int nframes, int mb_size, int search_param, int32_t *out_mv, uint64_t *out_cost);

/* libavutil's public block SAD: av_pixelutils_get_sad_fn (libavutil/pixelutils.h:31-52, pixelutils.c:43-111).  Square blocks of
 * 1 << w_bits pixels, w_bits = h_bits = 1 ... 5; NULL for anything else, like the reference (and when no device is set). */
typedef int (*b200_pixelutils_sad_fn)(const uint8_t *src1, ptrdiff_t stride1, const uint8_t *src2, ptrdiff_t stride2);   /* HOST pointers */
b200_pixelutils_sad_fn b200_pixelutils_get_sad_fn(int w_bits, int h_bits, int aligned, void *log_ctx);
/* batched, DEVICE pointers: out[i] = sad(frame1 + off1[i], stride1, frame2 + off2[i], stride2) over a (1 << w_bits)^2 block */
int  b200_pixelutils_sad_batch_device(B200Device *dev, int w_bits, const uint8_t *frame1, ptrdiff_t stride1, const uint8_t *frame2,
                                      ptrdiff_t stride2, const int64_t *off1, const int64_t *off2, int64_t n, int32_t *out);

/* ------------------------------------------------------------------------------------------------ h264qpel / hpeldsp
 * Replaces H264QpelContext (libavcodec/h264qpel.h:27-30) as filled by ff_h264qpel_init(c, 8) (libavcodec/h264qpel.c:50-120)
 * and HpelDSPContext (libavcodec/hpeldsp.h:39-97) as filled by ff_hpeldsp_init (libavcodec/hpeldsp.c:337-352), 8 bit.
 * qpel_mc_func (libavcodec/qpeldsp.h:65-67), op_pixels_func (libavcodec/hpeldsp.h:39-41). */
typedef void (*b200_qpel_mc_func)(uint8_t *dst, const uint8_t *src, ptrdiff_t stride);
typedef void (*b200_op_pixels_func)(uint8_t *block, const uint8_t *pixels, ptrdiff_t line_size, int h);
typedef struct B200H264QpelContext {
    b200_qpel_mc_func put_h264_qpel_pixels_tab[3][16];     /* [0:16x16, 1:8x8, 2:4x4][x + 4*y] */
    b200_qpel_mc_func avg_h264_qpel_pixels_tab[3][16];
} B200H264QpelContext;
typedef struct B200HpelDSPContext {
    b200_op_pixels_func put_pixels_tab[4][4];              /* [0:16, 1:8, 2:4, 3:2 wide][xhalf + 2*yhalf] */
    b200_op_pixels_func avg_pixels_tab[4][4];
    b200_op_pixels_func put_no_rnd_pixels_tab[3][4];       /* only [0..1] filled, like the reference */
    b200_op_pixels_func avg_no_rnd_pixels_tab[4];
} B200HpelDSPContext;
int  b200_h264qpel_init(B200H264QpelContext *c, int bit_depth);    /* bit_depth 8, or 9 / 10 / 12 / 14 (uint16 samples, stride in bytes:
                                                                     * h264qpel_template.c with BIT_DEPTH > 8); else B200_ENOSYS */
int  b200_hpeldsp_init(B200HpelDSPContext *c, int flags);

/* batched motion compensation, DEVICE pointers: operation i interpolates the block at src + src_off[i] into
 * dst + dst_off[i] (same stride for both, like the reference's single stride argument).
 * qpel op byte: bit0 avg, bits1-2 size index (0:16,1:8,2:4), bits3-6 position x+4*y.
 * hpel op byte: bits0-1 table (0 put, 1 avg, 2 put_no_rnd, 3 avg_no_rnd), bits2-3 size index (0:16,1:8,2:4,3:2),
 * bits4-5 xy; h[i] = block height. Destination blocks of one call must not overlap each other.
 * The qpel kernel may load the whole (size+5)^2 window of an operation whatever its position (unused taps never reach the
 * result): keep the planes edge-padded by 3 pixels / rows, as H.264 reference frames are (h264_mb.c mc_dir_part). */
int  b200_h264qpel_batch_device(B200Device *dev, int64_t n, const uint8_t *op, uint8_t *dst, const int64_t *dst_off,
                                const uint8_t *src, const int64_t *src_off, ptrdiff_t stride);
/* the same for 9 / 10 / 12 / 14 bit samples (uint16, native endian): op bytes as above, offsets and stride in BYTES (even) */
int  b200_h264qpel_hbd_batch_device(B200Device *dev, int bit_depth, int64_t n, const uint8_t *op, uint8_t *dst, const int64_t *dst_off,
                                    const uint8_t *src, const int64_t *src_off, ptrdiff_t stride);
int  b200_hpel_batch_device(B200Device *dev, int64_t n, const uint8_t *op, const uint8_t *h, uint8_t *dst,
                            const int64_t *dst_off, const uint8_t *src, const int64_t *src_off, ptrdiff_t stride);
/* HOST buffers, a stream of frames (mc_dir_part for every partition of every macroblock of a picture, h264_mb.c:210-330, collected per
 * frame): frame f occupies [f * frame_bytes, (f + 1) * frame_bytes) of `src` (its reference picture, edge-padded) and of `dst`; its
 * operations are op_begin[f] .. op_begin[f + 1] - 1 (op_begin has nframes + 1 entries), offsets counted from the start of dst / src
 * as in the device entry point, each operation staying inside its own frame.  Chunks of frames rotate over three streams: H2D of the
 * reference and destination pictures and of the lists, the kernel, D2H of the destination pictures. */
int  b200_h264qpel_frames_host(B200Device *dev, int nframes, int64_t frame_bytes, const int64_t *op_begin, const uint8_t *op,
                               uint8_t *dst, const int64_t *dst_off, const uint8_t *src, const int64_t *src_off, ptrdiff_t stride);

/* h264chroma: replaces H264ChromaContext (libavcodec/h264chroma.h:26-31) as filled by ff_h264chroma_init(c, 8)
 * (libavcodec/h264chroma.c:36-65): h264_chroma_mc_func (h264chroma.h:24), bilinear eighth-pel, x, y in 0..7. */
typedef void (*b200_h264_chroma_mc_func)(uint8_t *dst, const uint8_t *src, ptrdiff_t srcStride, int h, int x, int y);
typedef struct B200H264ChromaContext {
    b200_h264_chroma_mc_func put_h264_chroma_pixels_tab[4];   /* [0: 8 wide, 1: 4, 2: 2]; [3] is NULL like the reference's */
    b200_h264_chroma_mc_func avg_h264_chroma_pixels_tab[4];
} B200H264ChromaContext;
int  b200_h264chroma_init(B200H264ChromaContext *c, int bit_depth);   /* 9 ... 16: the uint16 tables (h264chroma.c:45-50); else 8 bit */
/* batched, DEVICE pointers.  op byte: bit0 avg, bits1-2 width index (0:8,1:4,2:2); h[i] = rows; xy[i] = x | y << 3.
 * Like the reference, x == 0 never reads the column right of the block and y == 0 never reads the row below it. */
int  b200_h264chroma_batch_device(B200Device *dev, int64_t n, const uint8_t *op, const uint8_t *h, const uint8_t *xy,
                                  uint8_t *dst, const int64_t *dst_off, const uint8_t *src, const int64_t *src_off,
                                  ptrdiff_t stride);
/* the same for 16-bit samples (any depth above 8): offsets and stride in BYTES (even) */
int  b200_h264chroma_hbd_batch_device(B200Device *dev, int64_t n, const uint8_t *op, const uint8_t *h, const uint8_t *xy,
                                      uint8_t *dst, const int64_t *dst_off, const uint8_t *src, const int64_t *src_off,
                                      ptrdiff_t stride);

/* videodsp: replaces VideoDSPContext (libavcodec/videodsp.h:32-69) as filled by ff_videodsp_init(ctx, 8)
 * (libavcodec/videodsp.c:32-62): emulated_edge_mc copies a block_w x block_h window whose top-left sample is picture
 * position (src_x, src_y) (src points AT that sample, inside or outside the picture) and replicates the border samples of
 * the w x h picture for everything outside it (videodsp_template.c:24-101).  prefetch is the reference's empty function. */
typedef struct B200VideoDSPContext {
    void (*emulated_edge_mc)(uint8_t *dst, const uint8_t *src, ptrdiff_t dst_linesize, ptrdiff_t src_linesize,
                             int block_w, int block_h, int src_x, int src_y, int w, int h);
    void (*prefetch)(const uint8_t *buf, ptrdiff_t stride, int h);
} B200VideoDSPContext;
int  b200_videodsp_init(B200VideoDSPContext *c, int bpc);             /* bpc <= 8: bytes; above: the 16-bit template (uint16 samples,
                                                                        * geometry in pixels, line sizes in bytes), videodsp.c:41-45 */
/* batched, DEVICE pointers: window i = geom[4i..4i+3] = {block_w, block_h, src_x, src_y} of the picture whose sample (0, 0)
 * is src + origin[i] (so one call can span many frames of equal w x h and linesize), written to buf + buf_off[i].
 * geom must be 16-byte aligned.  w == 0 or h == 0 writes nothing, like the reference. */
int  b200_emulated_edge_mc_batch_device(B200Device *dev, int64_t n, uint8_t *buf, const int64_t *buf_off,
                                        ptrdiff_t buf_linesize, const uint8_t *src, const int64_t *origin,
                                        ptrdiff_t src_linesize, const int32_t *geom, int w, int h);
/* the same for 16-bit samples: geom in pixels; buf_off, origin and the line sizes in BYTES (even) */
int  b200_emulated_edge_mc_hbd_batch_device(B200Device *dev, int64_t n, uint8_t *buf, const int64_t *buf_off,
                                            ptrdiff_t buf_linesize, const uint8_t *src, const int64_t *origin,
                                            ptrdiff_t src_linesize, const int32_t *geom, int w, int h);

/* ------------------------------------------------------------------------------------------------ libavutil/tx
 * Replaces av_tx_init / av_tx_uninit / av_tx_fn (libavutil/tx.h:151,202-208) for the types below, with the floating-point operation order
 * of the reference's C codelets (bit-identical results).
 * Inside FFmpeg the plug point is a codelet list with prio FF_TX_PRIO_MAX (libavutil/tx_priv.h:168, tx.c:340-351). */
#define B200_TX_FLOAT_FFT   0     /* AV_TX_FLOAT_FFT  */
#define B200_TX_FLOAT_MDCT  1     /* AV_TX_FLOAT_MDCT */
#define B200_TX_FLOAT_RDFT  6     /* AV_TX_FLOAT_RDFT: forward = real-to-complex (len floats -> len/2+1 complex), inverse =
                                   * complex-to-real; scale: const float *.  Like the reference's ff_tx_rdft_c2r
                                   * (tx_template.c:1655-1724) the host av_tx_fn rewrites its input on the inverse; the
                                   * batched device entry point leaves the input untouched.  AV_TX_REAL_TO_REAL /
                                   * AV_TX_REAL_TO_IMAGINARY are not implemented. */
#define B200_TX_DOUBLE_FFT  2     /* AV_TX_DOUBLE_FFT: AVComplexDouble in / out, power-of-two lengths up to 8192 (libavutil/tx_double.c) */
#define B200_TX_DOUBLE_MDCT 3     /* AV_TX_DOUBLE_MDCT: doubles, scale as const double *, power-of-two lengths */
#define B200_TX_INT32_FFT   4     /* AV_TX_INT32_FFT: AVComplexInt32 in / out, power-of-two lengths (libavutil/tx_int32.c) */
#define B200_TX_INT32_MDCT  5     /* AV_TX_INT32_MDCT: int32 samples, scale as const float *, power-of-two lengths (fixed-point AAC / AC-3) */
#define B200_TX_FLOAT_DCT   9     /* AV_TX_FLOAT_DCT: forward = DCT-II of len points, inverse = DCT-III of 2 * len points (the reference's
                                   * ff_tx_dct_init doubles the length it is given, libavutil/tx_template.c:1844-1848; callers pass N / 2).
                                   * ff_tx_dctII / ff_tx_dctIII (tx_template.c:1874-1968).  Input and output are N floats per transform;
                                   * unlike the reference neither the input is overwritten nor 2 floats of padding are needed. */
#define B200_TX_INPLACE     1     /* AV_TX_INPLACE: AV_TX_FLOAT_FFT only (power-of-two and compound); out == in is allowed (batch: out_step == in_step) */
#define B200_TX_UNALIGNED   2     /* AV_TX_UNALIGNED (accepted, no effect) */
#define B200_TX_FULL_IMDCT  4     /* AV_TX_FULL_IMDCT (libavutil/tx.h:175-180; ff_tx_mdct_inv_full, tx_template.c:1372-1413): inverse
                                   * AV_TX_FLOAT_MDCT only; 2 * len outputs, stride must be sizeof(float) */
typedef struct B200TXContext B200TXContext;
typedef void (*b200_tx_fn)(B200TXContext *s, void *out, void *in, ptrdiff_t stride);   /* av_tx_fn: HOST pointers */
/* like av_tx_init(); uses the process-wide default device.  scale: const float * (MDCT), ignored for FFT.
 * Lengths: powers of two, and for AV_TX_FLOAT_MDCT also 2 * N * 2^k with N = 15, 9, 7, 5 or 3: the compound transform av_tx_init() resolves
 * to for those (ff_tx_mdct_pfa_{15,9,7,5,3}xM_{inv,fwd}_float_c, libavutil/tx_template.c:1471-1599; 15 x M covers the Opus CELT sizes),
 * and for AV_TX_FLOAT_FFT also N * 2^k (2^k = 2 ... 512): the compound transform ff_tx_fft_pfa over fftN_ns and the split-radix
 * transform (libavutil/tx_template.c:948-1080; checkasm lengths 120 / 960 / 1920).  The compound FFT stores out[i * stride] like the
 * reference (stride in bytes, a multiple of 8); AV_TX_INPLACE is accepted for it.
 * Returns 0 or B200_ENOSYS (unsupported type / flags / length), B200_EINVAL, B200_ENODEV. */
int  b200_tx_init(B200TXContext **ctx, b200_tx_fn *tx, int type, int inv, int len, const void *scale, uint64_t flags);
/* host-only: the tables ff_tx_mdct_pfa_init() builds for a 15 x M MDCT, flattened into 32-bit words, for the CPU test tier
 * (layout8: word offsets of in_map, out_map, sub_map, exp, tab_53 + tab_7 + tab_9 (26 floats), cosine tables; then M and log2 M | N << 8).  Returns the word count. */
int  b200_tx_pfa_tables(int inv, int len, float scale, int32_t *words, int cap, int32_t *layout8);
/* host-only: the factor table of ff_tx_dct_init (N rotation factors, then N / 2 butterfly factors), for the CPU test tier */
int  b200_tx_dct_table(int inv, int len, float *tab, int cap);
/* host-only: the tables of an int32 transform flattened into words (layout4: offsets of map, exp, cosine tables; then log2 n) */
int  b200_tx_i32_tables(int type, int inv, int len, float scale, int32_t *words, int cap, int32_t *layout4);
/* host-only: the per-thread plan of the register-resident 512 ... 4096-point kernel (csrc/tx_r16.cu; word w of thread tg at
 * words[w * (n / 16) + tg]), for the CPU test tier, which replays the schedule in numpy and compares it with the checker.
 * Returns the word count, B200_ENOSYS for sizes the kernel does not cover. */
int  b200_tx_r16_plan(int n, int inv, uint32_t *words, int cap);
int  b200_tx_init_device(B200Device *dev, B200TXContext **ctx, b200_tx_fn *tx, int type, int inv, int len,
                         const void *scale, uint64_t flags);
void b200_tx_uninit(B200TXContext **ctx);
/* batched, DEVICE pointers: transform i reads in + i*in_step and writes out + i*out_step (bytes); `stride` as av_tx_fn
 * (bytes between MDCT input samples for the inverse / output samples for the forward transform; unused for FFT). */
int  b200_tx_batch_device(B200TXContext *ctx, void *out, const void *in, ptrdiff_t stride, int64_t count,
                          ptrdiff_t out_step, ptrdiff_t in_step);
/* same on HOST buffers (pinned memory for real overlap): chunks of transforms rotate over three streams, each chunk = one linear H2D
 * of [i * in_step, (i + n) * in_step), the kernels, one linear D2H; steps must be positive multiples of 16 bytes.  What a decoder that
 * keeps its coefficient buffers in host memory calls once per packet batch instead of count x av_tx_fn (aacdec_dsp_template.c:341-343). */
int  b200_tx_batch_host(B200TXContext *ctx, void *out, const void *in, ptrdiff_t stride, int64_t count,
                        ptrdiff_t out_step, ptrdiff_t in_step);

#ifdef __cplusplus
}
#endif
#endif /* B200DSP_H */
