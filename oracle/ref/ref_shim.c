/*
 * ref_shim.c — TEST INFRASTRUCTURE ONLY.
 *
 * Flat C ABI over the UNMODIFIED reference (FFmpeg) C implementation of the five DSP hot paths,
 * linked from the reference sources where they lie (see Makefile in this directory).  It is used
 *   - by tests/ to validate the oracle restatement (oracle/*.c) and the CUDA path against the real thing,
 *   - by scripts/gen_golden.py to produce tests/golden/ fixtures,
 *   - by bench.py's cpu_baseline / --impl reference leg (kind "reference").
 * Nothing in the product path (ffmpeg_b200/) may load this library.
 *
 * Every entry point only forwards to the reference's own table entry; no arithmetic lives here.
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>

#include "config.h"
#include "libavutil/mem.h"
#include "libavutil/log.h"
#include "libavutil/pixfmt.h"
#include "libavutil/tx.h"
#include "libavutil/cpu.h"
#include "libswscale/swscale.h"
#include "libswscale/swscale_internal.h"
#include "libswscale/graph.h"
#include "libavcodec/avcodec.h"
#include "libavcodec/idctdsp.h"
#include "libavcodec/me_cmp.h"
#include "libavcodec/h264qpel.h"
#include "libavcodec/hpeldsp.h"
#include "libavcodec/h264chroma.h"
#include "libavcodec/videodsp.h"
#include "libavcodec/h264dsp.h"
#include "libavfilter/motion_estimation.h"

/* ---- link stubs: the new-API filter graph (sws_scale_frame) is not part of the legacy hot path ---- */
int  ff_sws_pass_aligned_width(const SwsPass *pass, int width) { return width; }
SwsGraph *ff_sws_graph_alloc(void) { return NULL; }
void ff_sws_graph_free(SwsGraph **graph) { if (graph) *graph = NULL; }
int  ff_sws_graph_reinit(SwsGraph *graph, SwsContext *ctx, const SwsFormat *dst, const SwsFormat *src) { return AVERROR(ENOSYS); }
int  ff_sws_graph_run(SwsGraph *graph, const AVFrame *dst, const AVFrame *src) { return AVERROR(ENOSYS); }

#define API __attribute__((visibility("default")))

API int ffref_abi_version(void) { return 1; }
API unsigned ffref_swscale_version(void) { return swscale_version(); }
API int ffref_cpu_flags(void) { return av_get_cpu_flags(); }
API void ffref_set_quiet(void) { av_log_set_level(AV_LOG_ERROR); }

/* ------------------------------------------------------------------ swscale ------------------------------------- */

/* Opaque handle = the reference SwsContext (legacy API: sws_getContext + sws_scale). */
API void *ffref_sws_open_fmt(int srcW, int srcH, int dstW, int dstH, int dstFormat, int flags, int threads);
API void *ffref_sws_open(int srcW, int srcH, int dstW, int dstH, int flags, int threads)
{
    return ffref_sws_open_fmt(srcW, srcH, dstW, dstH, AV_PIX_FMT_RGB24, flags, threads);
}

API void *ffref_sws_open_io(int srcFormat, int srcW, int srcH, int dstFormat, int dstW, int dstH, int flags, int threads);
/* dstFormat: an AVPixelFormat value (packed 8-bit RGB family or yuv420p) */
API void *ffref_sws_open_fmt(int srcW, int srcH, int dstW, int dstH, int dstFormat, int flags, int threads)
{
    return ffref_sws_open_io(AV_PIX_FMT_YUV420P, srcW, srcH, dstFormat, dstW, dstH, flags, threads);
}

API void *ffref_sws_open_io(int srcFormat, int srcW, int srcH, int dstFormat, int dstW, int dstH, int flags, int threads)
{
    SwsContext *c = sws_alloc_context();
    if (!c) return NULL;
    c->src_w = srcW; c->src_h = srcH; c->dst_w = dstW; c->dst_h = dstH;
    c->src_format = srcFormat; c->dst_format = dstFormat;
    c->flags = flags;
    c->threads = threads;
    if (sws_init_context(c, NULL, NULL) < 0) { sws_freeContext(c); return NULL; }
    return c;
}

/* src_range / dst_range set on the context before sws_init_context (SwsContext fields, swscale.h), the way vf_scale's
 * in_range / out_range reach the scaler: with a yuv destination and different ranges the range conversion of
 * swscale.c:163-209 runs between the horizontal and the vertical pass */
API void *ffref_sws_open_range(int srcFormat, int srcW, int srcH, int srcRange, int dstFormat, int dstW, int dstH, int dstRange,
                               int flags, int threads)
{
    SwsContext *c = sws_alloc_context();
    if (!c) return NULL;
    c->src_w = srcW; c->src_h = srcH; c->dst_w = dstW; c->dst_h = dstH;
    c->src_format = srcFormat; c->dst_format = dstFormat;
    c->src_range = srcRange; c->dst_range = dstRange;
    c->flags = flags;
    c->threads = threads;
    if (sws_init_context(c, NULL, NULL) < 0) { sws_freeContext(c); return NULL; }
    return c;
}

/* the same with SwsContext.scaler_params (what sws_getContext copies from its `param` argument) */
API void *ffref_sws_open_params(int srcFormat, int srcW, int srcH, int srcRange, int dstFormat, int dstW, int dstH, int dstRange,
                                int flags, int threads, const double *param)
{
    SwsContext *c = sws_alloc_context();
    if (!c) return NULL;
    c->src_w = srcW; c->src_h = srcH; c->dst_w = dstW; c->dst_h = dstH;
    c->src_format = srcFormat; c->dst_format = dstFormat;
    c->src_range = srcRange; c->dst_range = dstRange;
    c->flags = flags;
    c->threads = threads;
    if (param) { c->scaler_params[0] = param[0]; c->scaler_params[1] = param[1]; }
    if (sws_init_context(c, NULL, NULL) < 0) { sws_freeContext(c); return NULL; }
    return c;
}

/* sws_init_context with srcFilter / dstFilter: four (coefficients, length) pairs each -- lumH, lumV, chrH, chrV; NULL / 0 = no vector */
API void *ffref_sws_open_filters(int srcFormat, int srcW, int srcH, int srcRange, int dstFormat, int dstW, int dstH, int dstRange,
                                 int flags, int threads, const double *const srcCoef[4], const int srcLen[4], const int dstLen[4], const double *param)
{
    SwsContext *c = sws_alloc_context();
    if (!c) return NULL;
    c->src_w = srcW; c->src_h = srcH; c->dst_w = dstW; c->dst_h = dstH;
    c->src_format = srcFormat; c->dst_format = dstFormat;
    c->src_range = srcRange; c->dst_range = dstRange;
    c->flags = flags;
    c->threads = threads;
    if (param) { c->scaler_params[0] = param[0]; c->scaler_params[1] = param[1]; }
    SwsFilter sf = { 0 }, df = { 0 };
    SwsVector **sv[4] = { &sf.lumH, &sf.lumV, &sf.chrH, &sf.chrV }, **dv[4] = { &df.lumH, &df.lumV, &df.chrH, &df.chrV };
    int any_s = 0, any_d = 0;
    for (int k = 0; k < 4; k++) {
        if (srcCoef && srcCoef[k] && srcLen[k] > 0) {
            *sv[k] = sws_allocVec(srcLen[k]);
            memcpy((*sv[k])->coeff, srcCoef[k], sizeof(double) * srcLen[k]);
            any_s = 1;
        }
        if (dstLen && dstLen[k] > 0) {
            *dv[k] = sws_allocVec(dstLen[k]);
            for (int i = 0; i < dstLen[k]; i++) (*dv[k])->coeff[i] = 1.0 / dstLen[k];
            any_d = 1;
        }
    }
    const int ret = sws_init_context(c, any_s ? &sf : NULL, any_d ? &df : NULL);
    for (int k = 0; k < 4; k++) { sws_freeVec(*sv[k]); sws_freeVec(*dv[k]); }
    if (ret < 0) { sws_freeContext(c); return NULL; }
    return c;
}

API void ffref_sws_close(void *h) { sws_freeContext((SwsContext *)h); }

/* colorspace details pass-through (sws_setColorspaceDetails); table index = SWS_CS_* */
API int ffref_sws_set_colorspace(void *h, int src_cs, int src_range, int dst_cs, int dst_range,
                                 int brightness, int contrast, int saturation)
{
    return sws_setColorspaceDetails((SwsContext *)h, sws_getCoefficients(src_cs), src_range,
                                    sws_getCoefficients(dst_cs), dst_range, brightness, contrast, saturation);
}

API int ffref_sws_scale(void *h, const uint8_t *y, int ys, const uint8_t *u, int us, const uint8_t *v, int vs,
                        int srcSliceY, int srcSliceH, uint8_t *dst, int ds)
{
    const uint8_t *src[4] = { y, u, v, NULL };
    int sstr[4] = { ys, us, vs, 0 };
    uint8_t *d[4] = { dst, NULL, NULL, NULL };
    int dstr[4] = { ds, 0, 0, 0 };
    return sws_scale((SwsContext *)h, src, sstr, srcSliceY, srcSliceH, d, dstr);
}

/* planar destination (yuv420p -> yuv420p scaling): three destination planes */
API int ffref_sws_scale_planar(void *h, const uint8_t *y, int ys, const uint8_t *u, int us, const uint8_t *v, int vs,
                               int srcSliceY, int srcSliceH, uint8_t *dy, int dys, uint8_t *du, int dus, uint8_t *dv, int dvs)
{
    const uint8_t *src[4] = { y, u, v, NULL };
    int sstr[4] = { ys, us, vs, 0 };
    uint8_t *d[4] = { dy, du, dv, NULL };
    int dstr[4] = { dys, dus, dvs, 0 };
    return sws_scale((SwsContext *)h, src, sstr, srcSliceY, srcSliceH, d, dstr);
}

/* Introspection of the initialised context, for checking the host-side filter generation. */
API int ffref_sws_info(void *h, int *out /* 16 ints */)
{
    SwsInternal *c = sws_internal((SwsContext *)h);
    out[0] = c->hLumFilterSize; out[1] = c->hChrFilterSize;
    out[2] = c->vLumFilterSize; out[3] = c->vChrFilterSize;
    out[4] = c->chrSrcW; out[5] = c->chrSrcH; out[6] = c->chrDstW; out[7] = c->chrDstH;
    out[8] = c->convert_unscaled != NULL;
    out[9] = c->chrSrcHSubSample; out[10] = c->chrSrcVSubSample;
    out[11] = c->chrDstHSubSample; out[12] = c->chrDstVSubSample;
    out[13] = c->opts.dst_w; out[14] = c->opts.dst_h; out[15] = c->needs_hcscale;
    return 0;
}

/* Range conversion the context ended up with (swscale.c:626-660): out[0] = 0 none, 1 limited -> full (src_range == 0),
 * 2 full -> limited; [1..4] = luma / chroma coefficient and offset narrowed the way the 8-bit line functions narrow them
 * (uint16 / int32, swscale.c:163-209); [5] = an unscaled converter is installed. */
API int ffref_sws_range_info(void *h, int *out)
{
    SwsInternal *c = sws_internal((SwsContext *)h);
    const int on = c->lumConvertRange != NULL;
    out[0] = !on ? 0 : c->opts.src_range ? 2 : 1;
    out[1] = on ? (uint16_t)c->lumConvertRange_coeff : 0; out[2] = on ? (int32_t)c->lumConvertRange_offset : 0;
    out[3] = on ? (uint16_t)c->chrConvertRange_coeff : 0; out[4] = on ? (int32_t)c->chrConvertRange_offset : 0;
    out[5] = c->convert_unscaled != NULL;
    return 0;
}

/* Source side of the initialised context: out[0] = bytes per pixel if the source is packed 8-bit RGB else 0, [1] / [2] =
 * chrSrcHSubSample / chrSrcVSubSample, [3] = an unscaled converter is installed, [4..12] = input_rgb2yuv_table. */
API int ffref_sws_rgb_info(void *h, int *out)
{
    SwsInternal *c = sws_internal((SwsContext *)h);
    const AVPixFmtDescriptor *d = av_pix_fmt_desc_get(c->opts.src_format);
    out[0] = isAnyRGB(c->opts.src_format) ? av_get_bits_per_pixel(d) / 8 : 0;
    out[1] = c->chrSrcHSubSample; out[2] = c->chrSrcVSubSample; out[3] = c->convert_unscaled != NULL;
    for (int i = 0; i < 9; i++) out[4 + i] = c->input_rgb2yuv_table[i];
    return 0;
}

/* which: 0 hLum 1 hChr 2 vLum 3 vChr.  Copies filter (n*size int16) and pos (n int32). Returns n. */
API int ffref_sws_get_filter(void *h, int which, int16_t *filter, int32_t *pos, int cap)
{
    SwsInternal *c = sws_internal((SwsContext *)h);
    const int16_t *f; const int32_t *p; int n, fs;
    switch (which) {
    case 0: f = c->hLumFilter; p = c->hLumFilterPos; n = c->opts.dst_w;    fs = c->hLumFilterSize; break;
    case 1: f = c->hChrFilter; p = c->hChrFilterPos; n = c->chrDstW; fs = c->hChrFilterSize; break;
    case 2: f = c->vLumFilter; p = c->vLumFilterPos; n = c->opts.dst_h;    fs = c->vLumFilterSize; break;
    default:f = c->vChrFilter; p = c->vChrFilterPos; n = c->chrDstH; fs = c->vChrFilterSize; break;
    }
    if (!f || !p) return 0;
    if (n > cap) n = cap;
    if (filter) memcpy(filter, f, (size_t)n * fs * sizeof(int16_t));
    if (pos)    memcpy(pos, p, (size_t)n * sizeof(int32_t));
    return n;
}

/* LUTs: table_rV/gU/gV/bU are pointer tables into yuvTable; export them as byte offsets from yuvTable
 * (gV is an int table), plus the yuvTable bytes themselves.  Returns the size of yuvTable copied. */
API int ffref_sws_get_tables(void *h, int32_t *rV, int32_t *gU, int32_t *gV, int32_t *bU /* 256+2*512 each */,
                             uint8_t *yuvtab, int cap)
{
    SwsInternal *c = sws_internal((SwsContext *)h);
    int n = 256 + 2 * YUVRGB_TABLE_HEADROOM;
    for (int i = 0; i < n; i++) {
        rV[i] = (int32_t)((uint8_t *)c->table_rV[i] - (uint8_t *)c->yuvTable);
        gU[i] = (int32_t)((uint8_t *)c->table_gU[i] - (uint8_t *)c->yuvTable);
        gV[i] = c->table_gV[i];
        bU[i] = (int32_t)((uint8_t *)c->table_bU[i] - (uint8_t *)c->yuvTable);
    }
    int sz = 1024 * 3 + 2 * YUVRGB_TABLE_LUMA_HEADROOM * 3; /* upper bound for 24bpp: see yuv2rgb.c */
    if (sz > cap) sz = cap;
    if (yuvtab && c->yuvTable) memcpy(yuvtab, c->yuvTable, sz);
    return sz;
}

/* The reference's horizontal scaler kernel as installed in the context (hyScale / hcScale). */
API void ffref_sws_hscale(void *h, int chroma, int16_t *dst, int dstW, const uint8_t *src,
                          const int16_t *filter, const int32_t *filterPos, int filterSize)
{
    SwsInternal *c = sws_internal((SwsContext *)h);
    (chroma ? c->hcScale : c->hyScale)(c, dst, dstW, src, filter, filterPos, filterSize);
}

/* ------------------------------------------------------------------ idctdsp ------------------------------------- */

static IDCTDSPContext g_idsp; static int g_idsp_ok;
static void idsp_init(void)
{
    if (g_idsp_ok) return;
    AVCodecContext *avctx = av_mallocz(sizeof(*avctx));
    avctx->idct_algo = FF_IDCT_SIMPLE;
    avctx->bits_per_raw_sample = 8;
    ff_idctdsp_init(&g_idsp, avctx);
    av_free(avctx);
    g_idsp_ok = 1;
}

API int ffref_idct_perm_type(uint8_t *perm64)
{
    idsp_init();
    if (perm64) memcpy(perm64, g_idsp.idct_permutation, 64);
    return g_idsp.perm_type;
}

/* kind: 0 = idct (in place), 1 = idct_put, 2 = idct_add.  blocks are clobbered like the reference does.
 * dest_off[i] = byte offset of block i's top-left in dest. */
API void ffref_idct_batch(int kind, int16_t *blocks, int nblocks, uint8_t *dest, ptrdiff_t line_size,
                          const int64_t *dest_off)
{
    idsp_init();
    for (int i = 0; i < nblocks; i++) {
        int16_t *b = blocks + 64 * (size_t)i;
        if (kind == 0)      g_idsp.idct(b);
        else if (kind == 1) g_idsp.idct_put(dest + dest_off[i], line_size, b);
        else                g_idsp.idct_add(dest + dest_off[i], line_size, b);
    }
}

/* clamp helpers: kind 0 put_pixels_clamped, 1 put_signed_pixels_clamped, 2 add_pixels_clamped */
API void ffref_pixels_clamped(int kind, const int16_t *block, uint8_t *pixels, ptrdiff_t line_size)
{
    idsp_init();
    if (kind == 0) g_idsp.put_pixels_clamped(block, pixels, line_size);
    else if (kind == 1) g_idsp.put_signed_pixels_clamped(block, pixels, line_size);
    else g_idsp.add_pixels_clamped(block, pixels, line_size);
}

/* ------------------------------------------------------------------ me_cmp -------------------------------------- */

static MECmpContext g_mecmp; static int g_mecmp_ok;
static void mecmp_init(void)
{
    if (g_mecmp_ok) return;
    AVCodecContext *avctx = av_mallocz(sizeof(*avctx));
    avctx->flags |= AV_CODEC_FLAG_BITEXACT;
    ff_me_cmp_init(&g_mecmp, avctx);
    av_free(avctx);
    g_mecmp_ok = 1;
}

/* dct_sad / dct_max / dct264_sad read the encoder context: pdsp.diff_pixels_unaligned, fdsp.fdct, sum_abs_dctelem (me_cmp.c:614-693) */
#include "libavcodec/mpegvideoenc.h"
static MPVEncContext *g_mpvenc; static int g_dct_algo;
static MPVEncContext *mecmp_enc(void)
{
    if (!g_mpvenc) {
        AVCodecContext *avctx = av_mallocz(sizeof(*avctx));
        g_mpvenc = av_mallocz(sizeof(*g_mpvenc));
        avctx->dct_algo = g_dct_algo;
        avctx->bits_per_raw_sample = 8;
        ff_fdctdsp_init(&g_mpvenc->fdsp, avctx);
        ff_pixblockdsp_init(&g_mpvenc->pdsp, 8);
        g_mpvenc->sum_abs_dctelem = g_mecmp.sum_abs_dctelem;
        av_free(avctx);
    }
    return g_mpvenc;
}
/* FDCTDSPContext as ff_fdctdsp_init fills it for (dct_algo, bits_per_raw_sample) */
API void ffref_fdct(int dct_algo, int bits, int is248, int16_t *block)
{
    FDCTDSPContext f;
    AVCodecContext *avctx = av_mallocz(sizeof(*avctx));
    LOCAL_ALIGNED_16(int16_t, tmp, [64]);
    avctx->dct_algo = dct_algo;
    avctx->bits_per_raw_sample = bits;
    ff_fdctdsp_init(&f, avctx);
    av_free(avctx);
    memcpy(tmp, block, 128);
    (is248 ? f.fdct248 : f.fdct)(tmp);
    memcpy(block, tmp, 128);
}
API void ffref_me_cmp_set_dct_algo(int algo) { g_dct_algo = algo; av_freep(&g_mpvenc); }

/* fn: 0 sad[idx], 1 sse[idx], 2 pix_abs[idx>>2][idx&3], 3 hadamard8_diff[idx], ..., 8 dct_sad, 9 dct_max, 10 dct264_sad */
API int ffref_me_cmp(int fn, int idx, const uint8_t *blk1, const uint8_t *blk2, ptrdiff_t stride, int h)
{
    mecmp_init();
    if (fn >= 8 && fn <= 10) {
        me_cmp_func g = fn == 8 ? g_mecmp.dct_sad[idx] : fn == 9 ? g_mecmp.dct_max[idx] : g_mecmp.dct264_sad[idx];
        return g ? g(mecmp_enc(), blk1, blk2, stride, h) : -1;
    }
    me_cmp_func f = fn == 0 ? g_mecmp.sad[idx] : fn == 1 ? g_mecmp.sse[idx] : fn == 3 ? g_mecmp.hadamard8_diff[idx] :
                    fn == 4 ? g_mecmp.vsad[idx] : fn == 5 ? g_mecmp.vsse[idx] : fn == 6 ? g_mecmp.nsse[idx] :
                    fn == 7 ? g_mecmp.median_sad[idx] : g_mecmp.pix_abs[idx >> 2][idx & 3];
    if (!f) return -1;
    return f(NULL, blk1, blk2, stride, h);
}
API int ffref_sum_abs_dctelem(const int16_t *block)
{
    mecmp_init();
    return g_mecmp.sum_abs_dctelem(block);
}

/* av_pixelutils_get_sad_fn(bits, bits, aligned = 0): -1 where the reference has no function */
#include "libavutil/pixelutils.h"
API int ffref_pixelutils_sad(int bits, const uint8_t *src1, ptrdiff_t stride1, const uint8_t *src2, ptrdiff_t stride2)
{
    av_pixelutils_sad_fn f = av_pixelutils_get_sad_fn(bits, bits, 0, NULL);
    return f ? f(src1, stride1, src2, stride2) : -1;
}

/* many calls in one go (for the CPU baseline): offsets into two frames */
API void ffref_me_cmp_batch(int fn, int idx, const uint8_t *f1, const uint8_t *f2, ptrdiff_t stride, int h,
                            const int64_t *off1, const int64_t *off2, int n, int32_t *out)
{
    mecmp_init();
    me_cmp_func f = fn == 0 ? g_mecmp.sad[idx] : fn == 1 ? g_mecmp.sse[idx] : fn == 3 ? g_mecmp.hadamard8_diff[idx] : g_mecmp.pix_abs[idx >> 2][idx & 3];
    for (int i = 0; i < n; i++) out[i] = f(NULL, f1 + off1[i], f2 + off2[i], stride, h);
}

/* Exhaustive search exactly as libavfilter/vf_mestimate.c drives ff_me_search_esa: for every mb of the frame
 * (raster order) mv initialised to (x_mb,y_mb); rows [mb_row0, mb_row1) only (to split over threads).
 * out_mv: 2 ints per mb, out_cost: 1 uint64 per mb, indexed by mb_y*b_width+mb_x. */
API void ffref_esa_frame(const uint8_t *cur, const uint8_t *ref, int linesize, int width, int height,
                         int mb_size, int search_param, int mb_row0, int mb_row1, int32_t *out_mv, uint64_t *out_cost)
{
    AVMotionEstContext me = { 0 };
    int log2_mb = 0; while ((1 << log2_mb) < mb_size) log2_mb++;
    int b_width = width >> log2_mb, b_height = height >> log2_mb;
    ff_me_init_context(&me, mb_size, search_param, width, height,
                       0, (b_width - 1) << log2_mb, 0, (b_height - 1) << log2_mb);
    me.data_cur = (uint8_t *)cur; me.data_ref = (uint8_t *)ref; me.linesize = linesize;
    if (mb_row1 > b_height) mb_row1 = b_height;
    for (int mb_y = mb_row0; mb_y < mb_row1; mb_y++)
        for (int mb_x = 0; mb_x < b_width; mb_x++) {
            int x_mb = mb_x << log2_mb, y_mb = mb_y << log2_mb;
            int mv[2] = { x_mb, y_mb };
            uint64_t cost = ff_me_search_esa(&me, x_mb, y_mb, mv);
            out_mv[2 * (mb_y * b_width + mb_x) + 0] = mv[0];
            out_mv[2 * (mb_y * b_width + mb_x) + 1] = mv[1];
            out_cost[mb_y * b_width + mb_x] = cost;
        }
}

/* ------------------------------------------------------------------ h264qpel / hpeldsp -------------------------- */

static H264QpelContext g_qpel; static HpelDSPContext g_hpel; static int g_pel_ok;
static void pel_init(void)
{
    if (g_pel_ok) return;
    ff_h264qpel_init(&g_qpel, 8);
    ff_hpeldsp_init(&g_hpel, AV_CODEC_FLAG_BITEXACT);
    g_pel_ok = 1;
}

/* avg: 0 put 1 avg; size_idx 0:16 1:8 2:4; pos = x + 4*y */
API void ffref_h264qpel(int avg, int size_idx, int pos, uint8_t *dst, const uint8_t *src, ptrdiff_t stride)
{
    pel_init();
    (avg ? g_qpel.avg_h264_qpel_pixels_tab : g_qpel.put_h264_qpel_pixels_tab)[size_idx][pos](dst, src, stride);
}

/* 9 / 10 / 12 / 14 bit tables of ff_h264qpel_init (uint16 samples, stride in bytes) */
API int ffref_h264qpel_hbd(int depth, int avg, int size_idx, int pos, uint8_t *dst, const uint8_t *src, ptrdiff_t stride)
{
    static H264QpelContext c[15]; static int ok[15];
    if (depth != 9 && depth != 10 && depth != 12 && depth != 14) return -1;
    if (!ok[depth]) { ff_h264qpel_init(&c[depth], depth); ok[depth] = 1; }
    (avg ? c[depth].avg_h264_qpel_pixels_tab : c[depth].put_h264_qpel_pixels_tab)[size_idx][pos](dst, src, stride);
    return 0;
}

API void ffref_h264qpel_batch(int n, const uint8_t *op /* n: bit0 avg, bits1-2 size_idx, bits 3-6 pos */,
                              uint8_t *dstbase, const int64_t *dst_off, const uint8_t *srcbase,
                              const int64_t *src_off, ptrdiff_t stride)
{
    pel_init();
    for (int i = 0; i < n; i++) {
        int avg = op[i] & 1, sz = (op[i] >> 1) & 3, pos = (op[i] >> 3) & 15;
        (avg ? g_qpel.avg_h264_qpel_pixels_tab : g_qpel.put_h264_qpel_pixels_tab)[sz][pos]
            (dstbase + dst_off[i], srcbase + src_off[i], stride);
    }
}

/* tab: 0 put 1 avg 2 put_no_rnd 3 avg_no_rnd; size_idx 0:16 1:8 2:4 3:2; xy = xhalf + 2*yhalf */
API int ffref_hpel(int tab, int size_idx, int xy, uint8_t *block, const uint8_t *pixels, ptrdiff_t line_size, int h)
{
    pel_init();
    op_pixels_func f = NULL;
    if (tab == 0) f = g_hpel.put_pixels_tab[size_idx][xy];
    else if (tab == 1) f = g_hpel.avg_pixels_tab[size_idx][xy];
    else if (tab == 2) f = size_idx < 3 ? g_hpel.put_no_rnd_pixels_tab[size_idx][xy] : NULL;
    else f = size_idx == 0 ? g_hpel.avg_no_rnd_pixels_tab[xy] : NULL;
    if (!f) return -1;
    f(block, pixels, line_size, h);
    return 0;
}

/* avg: 0 put 1 avg; idx 0: 8 wide, 1: 4 wide, 2: 2 wide; x, y in 0..7 (eighth-pel) */
API int ffref_h264chroma(int avg, int idx, uint8_t *dst, const uint8_t *src, ptrdiff_t stride, int h, int x, int y)
{
    static H264ChromaContext c; static int ok;
    if (!ok) { ff_h264chroma_init(&c, 8); ok = 1; }
    h264_chroma_mc_func f = (avg ? c.avg_h264_chroma_pixels_tab : c.put_h264_chroma_pixels_tab)[idx];
    if (!f) return -1;
    f(dst, src, stride, h, x, y);
    return 0;
}

API int ffref_h264chroma_hbd(int depth, int avg, int idx, uint8_t *dst, const uint8_t *src, ptrdiff_t stride, int h, int x, int y)
{
    static H264ChromaContext c[17]; static int ok[17];
    if (depth < 9 || depth > 16) return -1;
    if (!ok[depth]) { ff_h264chroma_init(&c[depth], depth); ok[depth] = 1; }
    h264_chroma_mc_func f = (avg ? c[depth].avg_h264_chroma_pixels_tab : c[depth].put_h264_chroma_pixels_tab)[idx];
    if (!f) return -1;
    f(dst, src, stride, h, x, y);
    return 0;
}

API void ffref_emulated_edge_mc_hbd(uint8_t *buf, const uint8_t *src, ptrdiff_t buf_linesize, ptrdiff_t src_linesize,
                                    int block_w, int block_h, int src_x, int src_y, int w, int h)
{
    static VideoDSPContext c; static int ok;
    if (!ok) { ff_videodsp_init(&c, 10); ok = 1; }
    c.emulated_edge_mc(buf, src, buf_linesize, src_linesize, block_w, block_h, src_x, src_y, w, h);
}

API void ffref_emulated_edge_mc(uint8_t *buf, const uint8_t *src, ptrdiff_t buf_linesize, ptrdiff_t src_linesize,
                                int block_w, int block_h, int src_x, int src_y, int w, int h)
{
    static VideoDSPContext c; static int ok;
    if (!ok) { ff_videodsp_init(&c, 8); ok = 1; }
    c.emulated_edge_mc(buf, src, buf_linesize, src_linesize, block_w, block_h, src_x, src_y, w, h);
}

/* kind 0: idct_add (4x4), 1: idct8_add, 2: idct_dc_add, 3: idct8_dc_add; block is cleared like the reference does */
API int ffref_h264_idct(int kind, uint8_t *dst, int16_t *block, ptrdiff_t stride)
{
    static H264DSPContext c; static int ok;
    if (!ok) { ff_h264dsp_init(&c, 8, 1); ok = 1; }
    switch (kind) {
    case 0: c.idct_add(dst, block, stride); return 0;
    case 1: c.idct8_add(dst, block, stride); return 0;
    case 2: c.idct_dc_add(dst, block, stride); return 0;
    case 3: c.idct8_dc_add(dst, block, stride); return 0;
    }
    return -1;
}

/* H.264 weighted prediction: idx 0..3 = widths 16, 8, 4, 2 */
API void ffref_h264_weight(int idx, uint8_t *block, ptrdiff_t stride, int height, int log2_denom, int weight, int offset)
{
    static H264DSPContext c; static int ok;
    if (!ok) { ff_h264dsp_init(&c, 8, 1); ok = 1; }
    c.weight_pixels_tab[idx](block, stride, height, log2_denom, weight, offset);
}
API void ffref_h264_biweight(int idx, uint8_t *dst, uint8_t *src, ptrdiff_t stride, int height, int log2_denom,
                             int weightd, int weights, int offset)
{
    static H264DSPContext c; static int ok;
    if (!ok) { ff_h264dsp_init(&c, 8, 1); ok = 1; }
    c.biweight_pixels_tab[idx](dst, src, stride, height, log2_denom, weightd, weights, offset);
}

/* residual adds of ff_h264dsp_init for 9 / 10 / 12 / 14 bit samples: block holds int32 coefficients behind the int16_t * type */
static H264DSPContext *ref_h264dsp_hbd(int depth);
API int ffref_h264_idct_hbd(int depth, int kind, uint8_t *dst, int32_t *block, ptrdiff_t stride)
{
    H264DSPContext *c = ref_h264dsp_hbd(depth);
    if (!c) return -1;
    switch (kind) {
    case 0: c->idct_add(dst, (int16_t *)block, stride); return 0;
    case 1: c->idct8_add(dst, (int16_t *)block, stride); return 0;
    case 2: c->idct_dc_add(dst, (int16_t *)block, stride); return 0;
    case 3: c->idct8_dc_add(dst, (int16_t *)block, stride); return 0;
    }
    return -1;
}

/* weight / biweight tables of ff_h264dsp_init for 9 / 10 / 12 / 14 bit samples */
static H264DSPContext *ref_h264dsp_hbd(int depth)
{
    static H264DSPContext c[15]; static int ok[15];
    if (depth != 9 && depth != 10 && depth != 12 && depth != 14) return NULL;
    if (!ok[depth]) { ff_h264dsp_init(&c[depth], depth, 1); ok[depth] = 1; }
    return &c[depth];
}
API int ffref_h264_weight_hbd(int depth, int idx, uint8_t *block, ptrdiff_t stride, int height, int log2_denom, int weight, int offset)
{
    H264DSPContext *c = ref_h264dsp_hbd(depth);
    if (!c) return -1;
    c->weight_pixels_tab[idx](block, stride, height, log2_denom, weight, offset);
    return 0;
}
API int ffref_h264_biweight_hbd(int depth, int idx, uint8_t *dst, uint8_t *src, ptrdiff_t stride, int height, int log2_denom,
                                int weightd, int weights, int offset)
{
    H264DSPContext *c = ref_h264dsp_hbd(depth);
    if (!c) return -1;
    c->biweight_pixels_tab[idx](dst, src, stride, height, log2_denom, weightd, weights, offset);
    return 0;
}

/* ------------------------------------------------------------------ tx ------------------------------------------ */

typedef struct { AVTXContext *ctx; av_tx_fn fn; } RefTx;

/* type: AV_TX_FLOAT_FFT=0, AV_TX_FLOAT_MDCT=1 (tx.h enum order) */
API void *ffref_tx_open(int type, int inv, int len, float scale, unsigned flags)
{
    RefTx *t = calloc(1, sizeof(*t));
    if (!t) return NULL;
    if (av_tx_init(&t->ctx, &t->fn, (enum AVTXType)type, inv, len, &scale, flags) < 0) { free(t); return NULL; }
    return t;
}
/* the double types (AV_TX_DOUBLE_FFT = 2, AV_TX_DOUBLE_MDCT = 3, ...): av_tx_init reads the scale as a const double * */
API void *ffref_txd_open(int type, int inv, int len, double scale, unsigned flags)
{
    RefTx *t = calloc(1, sizeof(*t));
    if (!t) return NULL;
    if (av_tx_init(&t->ctx, &t->fn, (enum AVTXType)type, inv, len, &scale, flags) < 0) { free(t); return NULL; }
    return t;
}
API void ffref_tx_close(void *h) { RefTx *t = h; if (t) { av_tx_uninit(&t->ctx); free(t); } }

/* count transforms; in/out advance by in_step/out_step BYTES per transform; stride is the av_tx_fn stride arg */
API void ffref_tx_run(void *h, void *out, void *in, ptrdiff_t stride, int count, ptrdiff_t out_step, ptrdiff_t in_step)
{
    RefTx *t = h;
    for (int i = 0; i < count; i++)
        t->fn(t->ctx, (uint8_t *)out + i * out_step, (uint8_t *)in + i * in_step, stride);
}

/* the codelet tree av_tx_init() resolved to, one "depth name len flags" line per node (the float operation order of a compound
 * transform depends on which decomposition won: tests pin the CUDA plan to the same tree) */
#include "libavutil/tx_priv.h"
static int tx_describe_node(const AVTXContext *s, int depth, char *buf, int cap, int pos)
{
    if (!s || !s->cd_self) return pos;
    int n = snprintf(buf + pos, pos < cap ? cap - pos : 0, "%d %s %d %llx\n", depth, s->cd_self->name, s->len, (unsigned long long)s->flags);
    pos += n > 0 ? n : 0;
    for (int i = 0; i < s->nb_sub; i++) pos = tx_describe_node(&s->sub[i], depth + 1, buf, cap, pos);
    return pos;
}
API int ffref_tx_describe(void *h, char *buf, int cap)
{
    RefTx *t = h;
    if (cap > 0) buf[0] = 0;
    return tx_describe_node(t->ctx, 0, buf, cap, 0);
}

/* ------------------------------------------------------------------ mpegvideo inverse quantisers ---------------- */
#include "libavcodec/mpegvideo.h"
#include "libavcodec/mpegvideodata.h"
#include "libavcodec/mathops.h"
#include "libavcodec/mpegvideo_unquantize.h"

/* variant: 0 mpeg1 intra, 1 mpeg1 inter, 2 mpeg2 intra, 3 mpeg2 intra (bitexact flavour), 4 mpeg2 inter, 5 h263 intra,
 * 6 h263 inter — the members of MPVUnquantDSPContext as ff_mpv_unquantize_init() fills them.  The MPVContext holds only what
 * the functions read; scan tables are built by ff_init_scantable with the identity permutation of the C simple IDCT. */
API int ffref_mpv_unquantize_batch(int variant, const uint16_t *intra_matrix, const uint16_t *inter_matrix, int alternate_scan,
                                   int y_dc_scale, int c_dc_scale, int q_scale_type, int h263_aic, int ac_pred,
                                   int16_t *blocks, int64_t nblocks, const uint8_t *blk_n, const uint8_t *qscale,
                                   const int8_t *last_index)
{
    MPVUnquantDSPContext dsp;
    MPVContext *s = av_mallocz(sizeof(*s));
    uint8_t perm[64];
    if (!s) return -1;
    ff_mpv_unquantize_init(&dsp, variant == 3, q_scale_type);
    for (int i = 0; i < 64; i++) { perm[i] = i; s->intra_matrix[i] = intra_matrix[i]; s->inter_matrix[i] = inter_matrix[i]; }
    ff_init_scantable(perm, &s->intra_scantable, alternate_scan ? ff_alternate_vertical_scan : ff_zigzag_direct);
    ff_init_scantable(perm, &s->inter_scantable, alternate_scan ? ff_alternate_vertical_scan : ff_zigzag_direct);
    s->y_dc_scale = y_dc_scale; s->c_dc_scale = c_dc_scale; s->q_scale_type = q_scale_type;
    s->h263_aic = h263_aic; s->ac_pred = ac_pred;
    void (*fn)(const MPVContext *, int16_t *, int, int) =
        variant == 0 ? dsp.dct_unquantize_mpeg1_intra : variant == 1 ? dsp.dct_unquantize_mpeg1_inter :
        variant == 2 || variant == 3 ? dsp.dct_unquantize_mpeg2_intra : variant == 4 ? dsp.dct_unquantize_mpeg2_inter :
        variant == 5 ? dsp.dct_unquantize_h263_intra : dsp.dct_unquantize_h263_inter;
    for (int64_t b = 0; b < nblocks; b++) {
        const int n = blk_n ? blk_n[b] : (int)(b % 6);
        s->block_last_index[n] = last_index[b];
        fn(s, blocks + 64 * b, n, qscale[b]);
    }
    av_free(s);
    return 0;
}

/* ------------------------------------------------------------------ AVFloatDSPContext ---------------------------- */
#include "libavutil/float_dsp.h"

/* op = member index in AVFloatDSPContext (float_dsp.h:24-210); argument meaning as orc_float_dsp */
API int ffref_float_dsp(int op, void *dst, const void *src0, const void *src1, const void *src2, double mul, int len)
{
    static AVFloatDSPContext *f;
    if (!f) f = avpriv_float_dsp_alloc(0);
    if (!f) return -1;
    switch (op) {
    case 0:  f->vector_fmul(dst, src0, src1, len); return 0;
    case 1:  f->vector_fmac_scalar(dst, src0, (float)mul, len); return 0;
    case 2:  f->vector_dmac_scalar(dst, src0, mul, len); return 0;
    case 3:  f->vector_fmul_scalar(dst, src0, (float)mul, len); return 0;
    case 4:  f->vector_dmul_scalar(dst, src0, mul, len); return 0;
    case 5:  f->vector_fmul_window(dst, src0, src1, src2, len); return 0;
    case 6:  f->vector_fmul_add(dst, src0, src1, src2, len); return 0;
    case 7:  f->vector_fmul_reverse(dst, src0, src1, len); return 0;
    case 8:  f->butterflies_float(dst, (float *)src0, len); return 0;
    case 9:  *(float *)dst = f->scalarproduct_float(src0, src1, len); return 0;
    case 10: f->vector_dmul(dst, src0, src1, len); return 0;
    case 11: *(double *)dst = f->scalarproduct_double(src0, src1, len); return 0;
    }
    return -1;
}

/* ------------------------------------------------------------------ simple IDCT, 10 / 12 bit ---------------------- */
#include "libavcodec/avcodec.h"
#include "libavcodec/idctdsp.h"

/* through ff_idctdsp_init() with bits_per_raw_sample = depth (9, 10 or 12), idct_algo = FF_IDCT_SIMPLE: whatever it installs
 * (idctdsp.c:248-266).  kind 0 idct, 1 idct_put, 2 idct_add; dest = uint16 pixels, line_size in bytes. */
API int ffref_idct_hbd(int depth, int kind, uint8_t *dest, ptrdiff_t line_size, int16_t *block)
{
    static IDCTDSPContext c[13];
    static int done[13];
    if (depth < 9 || depth > 12) return -1;
    if (!done[depth]) {
        AVCodecContext *avctx = av_mallocz(sizeof(*avctx));
        if (!avctx) return -1;
        avctx->bits_per_raw_sample = depth;
        avctx->idct_algo = FF_IDCT_SIMPLE;
        ff_idctdsp_init(&c[depth], avctx);
        av_free(avctx);
        done[depth] = 1;
    }
    if (c[depth].perm_type != FF_IDCT_PERM_NONE) return -2;
    if (kind == 0) c[depth].idct(block);
    else if (kind == 1) c[depth].idct_put(dest, line_size, block);
    else c[depth].idct_add(dest, line_size, block);
    return 0;
}

/* ------------------------------------------------------------------ ProresDSPContext ------------------------------ */
#include "libavcodec/proresdsp.h"

/* ff_proresdsp_init(bits).idct_put: out = uint16 pixels, linesize in bytes */
API int ffref_prores_idct_put(int bits, uint8_t *out, ptrdiff_t linesize, int16_t *block, const int16_t *qmat)
{
    static ProresDSPContext c[2];
    static int done[2];
    if (bits != 10 && bits != 12) return -1;
    const int i = bits == 12;
    if (!done[i]) { ff_proresdsp_init(&c[i], bits); done[i] = 1; }
    if (c[i].idct_permutation_type != FF_IDCT_PERM_NONE) return -2;
    c[i].idct_put((uint16_t *)out, linesize, block, qmat);
    return 0;
}

/* ------------------------------------------------------------------ H.264 deblocking ----------------------------- */
/* the loop-filter members of H264DSPContext, 8 bit; kinds 0-11 from ff_h264dsp_init(c, 8, 1) in the member order of h264dsp.h:48-73,
 * kinds 12-15 = h_loop_filter_chroma, h_loop_filter_chroma_mbaff and their _intra forms from ff_h264dsp_init(c, 8, 2) (4:2:2) */
API int ffref_h264_loop_filter(int kind, uint8_t *pix, ptrdiff_t stride, int alpha, int beta, int8_t *tc0)
{
    static H264DSPContext c420, c422; static int ok;
    if (!ok) { ff_h264dsp_init(&c420, 8, 1); ff_h264dsp_init(&c422, 8, 2); ok = 1; }
    switch (kind) {
    case 0:  c420.v_loop_filter_luma(pix, stride, alpha, beta, tc0); break;
    case 1:  c420.h_loop_filter_luma(pix, stride, alpha, beta, tc0); break;
    case 2:  c420.h_loop_filter_luma_mbaff(pix, stride, alpha, beta, tc0); break;
    case 3:  c420.v_loop_filter_luma_intra(pix, stride, alpha, beta); break;
    case 4:  c420.h_loop_filter_luma_intra(pix, stride, alpha, beta); break;
    case 5:  c420.h_loop_filter_luma_mbaff_intra(pix, stride, alpha, beta); break;
    case 6:  c420.v_loop_filter_chroma(pix, stride, alpha, beta, tc0); break;
    case 7:  c420.h_loop_filter_chroma(pix, stride, alpha, beta, tc0); break;
    case 8:  c420.h_loop_filter_chroma_mbaff(pix, stride, alpha, beta, tc0); break;
    case 9:  c420.v_loop_filter_chroma_intra(pix, stride, alpha, beta); break;
    case 10: c420.h_loop_filter_chroma_intra(pix, stride, alpha, beta); break;
    case 11: c420.h_loop_filter_chroma_mbaff_intra(pix, stride, alpha, beta); break;
    case 12: c422.h_loop_filter_chroma(pix, stride, alpha, beta, tc0); break;
    case 13: c422.h_loop_filter_chroma_mbaff(pix, stride, alpha, beta, tc0); break;
    case 14: c422.h_loop_filter_chroma_intra(pix, stride, alpha, beta); break;
    case 15: c422.h_loop_filter_chroma_mbaff_intra(pix, stride, alpha, beta); break;
    default: return -1;
    }
    return 0;
}


/* the loop-filter members for 9 / 10 / 12 / 14 bit samples: kinds as ffref_h264_loop_filter (0-11 from chroma_format_idc 1, 12-15 from 2) */
API int ffref_h264_loop_filter_hbd(int depth, int kind, uint8_t *pix, ptrdiff_t stride, int alpha, int beta, int8_t *tc0)
{
    static H264DSPContext c420[15], c422[15]; static int ok[15];
    if (depth != 9 && depth != 10 && depth != 12 && depth != 14) return -1;
    if (!ok[depth]) { ff_h264dsp_init(&c420[depth], depth, 1); ff_h264dsp_init(&c422[depth], depth, 2); ok[depth] = 1; }
    H264DSPContext *a = &c420[depth], *b = &c422[depth];
    switch (kind) {
    case 0:  a->v_loop_filter_luma(pix, stride, alpha, beta, tc0); break;
    case 1:  a->h_loop_filter_luma(pix, stride, alpha, beta, tc0); break;
    case 2:  a->h_loop_filter_luma_mbaff(pix, stride, alpha, beta, tc0); break;
    case 3:  a->v_loop_filter_luma_intra(pix, stride, alpha, beta); break;
    case 4:  a->h_loop_filter_luma_intra(pix, stride, alpha, beta); break;
    case 5:  a->h_loop_filter_luma_mbaff_intra(pix, stride, alpha, beta); break;
    case 6:  a->v_loop_filter_chroma(pix, stride, alpha, beta, tc0); break;
    case 7:  a->h_loop_filter_chroma(pix, stride, alpha, beta, tc0); break;
    case 8:  a->h_loop_filter_chroma_mbaff(pix, stride, alpha, beta, tc0); break;
    case 9:  a->v_loop_filter_chroma_intra(pix, stride, alpha, beta); break;
    case 10: a->h_loop_filter_chroma_intra(pix, stride, alpha, beta); break;
    case 11: a->h_loop_filter_chroma_mbaff_intra(pix, stride, alpha, beta); break;
    case 12: b->h_loop_filter_chroma(pix, stride, alpha, beta, tc0); break;
    case 13: b->h_loop_filter_chroma_mbaff(pix, stride, alpha, beta, tc0); break;
    case 14: b->h_loop_filter_chroma_intra(pix, stride, alpha, beta); break;
    case 15: b->h_loop_filter_chroma_mbaff_intra(pix, stride, alpha, beta); break;
    default: return -1;
    }
    return 0;
}
