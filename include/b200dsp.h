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
 * Replaces the legacy scaler for AV_PIX_FMT_YUV420P -> AV_PIX_FMT_RGB24:
 *   sws_getContext / sws_init_context ....... libswscale/utils.c:1919,1884 (ff_sws_init_single_context :1137)
 *   sws_setColorspaceDetails ................ libswscale/utils.c:849
 *   sws_scale ............................... libswscale/swscale.h:583, swscale.c:1626
 *   SwsFunc (convert_unscaled / ff_swscale).. libswscale/swscale_internal.h:99-101, swscale.c:263
 * Flags keep the reference's values (libswscale/swscale.h:88-118).  Output is bit-identical to the reference's
 * C path (the one FATE pins with accurate_rnd+bitexact, and yuv2rgb_c_24_rgb without accurate_rnd).  */
#define B200_PIX_FMT_YUV420P 0     /* AV_PIX_FMT_YUV420P, libavutil/pixfmt.h */
#define B200_PIX_FMT_RGB24   2     /* AV_PIX_FMT_RGB24 */

#define B200_SWS_FAST_BILINEAR 0x1
#define B200_SWS_BILINEAR      0x2
#define B200_SWS_BICUBIC       0x4
#define B200_SWS_POINT         0x10
#define B200_SWS_AREA          0x20
#define B200_SWS_BICUBLIN      0x40
#define B200_SWS_FULL_CHR_H_INT 0x2000
#define B200_SWS_ACCURATE_RND  0x40000
#define B200_SWS_BITEXACT      0x80000

typedef struct B200SwsContext B200SwsContext;

/* like sws_getContext(); srcFilter/dstFilter/param are not supported (must be the defaults). NULL on failure. */
B200SwsContext *b200_sws_getContext(B200Device *dev, int srcW, int srcH, int srcFormat,
                                    int dstW, int dstH, int dstFormat, int flags);
void b200_sws_freeContext(B200SwsContext *c);
/* like sws_setColorspaceDetails(); `table`/dstRange are accepted and ignored for RGB output like the reference */
int  b200_sws_setColorspaceDetails(B200SwsContext *c, const int inv_table[4], int srcRange,
                                   const int table[4], int dstRange, int brightness, int contrast, int saturation);
/* drop-in for sws_scale(): HOST pointers, strides in bytes (negative allowed), returns output lines.
 * Only whole-frame calls (srcSliceY == 0, srcSliceH == srcH) are accepted; others return B200_ENOSYS. */
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
/* batched, HOST pointers (pinned memory recommended): chunks of frames are copied in, converted and copied back
 * on rotating streams so that H2D, kernels and D2H overlap.  Synchronous: returns when dst is complete. */
int  b200_sws_scale_batch_host(B200SwsContext *c, const uint8_t *const src[3], const int srcStride[3],
                               const int64_t srcFrameStride[3], uint8_t *dst, int dstStride,
                               int64_t dstFrameStride, int nframes);
/* introspection for tests: same layout as the shim used on the reference (16 ints) */
int  b200_sws_info(const B200SwsContext *c, int *out16);
/* which: 0 hLum 1 hChr 2 vLum 3 vChr — host copies of the generated filter tables */
int  b200_sws_get_filter(const B200SwsContext *c, int which, int16_t *filter, int32_t *pos, int cap);

/* host-only: build the set-up tables for a configuration WITHOUT touching a GPU (used by the CPU test tier to check
 * the filter generation against the reference).  info16 as b200_sws_info; filter/pos may be NULL. Returns n or <0. */
int  b200_sws_plan_probe(int srcW, int srcH, int dstW, int dstH, int flags, int which,
                         int16_t *filter, int32_t *pos, int cap, int *info16);

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

#ifdef __cplusplus
}
#endif
#endif /* B200DSP_H */
