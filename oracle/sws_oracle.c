/*
 * sws_oracle.c — TEST INFRASTRUCTURE ONLY.  CPU restatement of libswscale's legacy yuv420p -> rgb24 path.
 *
 * Written from the behaviour of the reference (FFmpeg 8.0.git), not copied from it.  Follows:
 *   context set-up .......... libswscale/utils.c:1137-1760 (ff_sws_init_single_context)
 *   filter generation ....... libswscale/utils.c:197-612   (initFilter), :168-175 (get_local_pos)
 *   horizontal FIR .......... libswscale/swscale.c:128-142 (hScale8To15_c)
 *   line scheduling ......... libswscale/swscale.c:412-535 (ff_swscale), libswscale/vscale.c:109-171 (packed_vscale)
 *   vertical FIR + writers .. libswscale/output.c:1789-1939 (yuv2rgb_{X,2,1}_c_template), :1663-1713 (yuv2rgb_write)
 *                             libswscale/output.c:1998-2330 (full-chroma variants, yuv2rgb_write_full)
 *   LUT construction ........ libswscale/yuv2rgb.c:680-703 (fill_table, fill_gv_table), :717-914 (24 bpp case)
 *   unscaled converter ...... libswscale/yuv2rgb.c:68-236 (YUV420FUNC/PUTRGB24), gate swscale_unscaled.c:2426-2431
 *
 * The reference schedules lines through a ring buffer; here whole horizontally-scaled planes are kept, which
 * yields the same bytes because every output line depends only on its own taps (rgb24 has no dither state).
 * filterAlign is 1 (generic C build of the reference: no MMX/NEON/AltiVec alignment padding, utils.c:1675-1710).
 */
#include "oracle.h"
#include <stdlib.h>
#include <math.h>
#include <string.h>
#include <errno.h>

#define HEADROOM      512   /* YUVRGB_TABLE_HEADROOM       swscale_internal.h */
#define LUMA_HEADROOM 512   /* YUVRGB_TABLE_LUMA_HEADROOM  swscale_internal.h */
#define NTAB          (256 + 2 * HEADROOM)
#define YTAB_SIZE     (1024 + 2 * LUMA_HEADROOM)

struct OrcSws {
    int srcW, srcH, dstW, dstH, flags;
    int chrSrcW, chrSrcH, chrDstW, chrDstH;
    int chrDstHSub;            /* 1: one chroma sample per two output pixels, 0: full chroma */
    int unscaled_lut;          /* 1: reference would install yuv2rgb_c_24_rgb / _24_bgr / _32 as convert_unscaled */
    int bpp, ro, go, bo, ao;   /* bytes per pixel and byte positions of R, G, B, A (ao < 0: no alpha byte) */
    int planar;                /* 1: destination is yuv420p (three planes), 0: packed RGB */
    int src_nv;                /* 0: yuv420p source, 1: nv12 (U,V interleaved in plane 1), 2: nv21 (V,U) — nvXXtoUV_c, input.c:921-948 */
    int fast_bilinear, lumXInc, chrXInc;   /* SWS_FAST_BILINEAR: the horizontal pass is ff_hyscale_fast_c / ff_hcscale_fast_c */
    int16_t *hLum, *hChr, *vLum, *vChr;
    int32_t *hLumPos, *hChrPos, *vLumPos, *vChrPos;
    int hLumSize, hChrSize, vLumSize, vChrSize;
    /* yuv->rgb */
    uint8_t ytab[YTAB_SIZE];
    int offR[NTAB], offGU[NTAB], offGV[NTAB], offB[NTAB];
    int y_offset, y_coeff, v2r, v2g, u2g, u2b;   /* int16 values for the full-chroma writer */
    /* yuv -> yuv range conversion on the 15-bit lines between the two passes (swscale.c:163-209, :577-660) */
    int src_range, dst_range;
    int range_conv;            /* 0 none, 1 limited -> full (lum/chrRangeToJpeg_c, clips at 2^15-1), 2 full -> limited (FromJpeg) */
    int lum_coeff, lum_offset, chr_coeff, chr_offset;
    int src_cs[4], dst_cs[4];
    /* packed RGB source (input.c readers -> 16-bit lines -> hScale16To15_c) */
    int src_rgb;               /* 0: yuv source; 3 / 4: bytes per source pixel */
    /* yuv -> yuv with two different matrices: the reference cascades through bgr24 (sws_setColorspaceDetails, utils.c:914-989) */
    struct OrcSws *casc[2];
    int casc_w, casc_h;
    int open_src_fmt, open_dst_fmt, open_flags;      /* as given to the open call */
    double open_param[2];
    int need_alpha;            /* c->needAlpha = isALPHA(src) && isALPHA(dst) (utils.c:1405): the alpha plane goes through the scaler */
    int sro, sgo, sbo;         /* byte positions of R, G, B in a source pixel */
    int chrSrcHSub, chrSrcVSub;
    int rgb2yuv[9];            /* input_rgb2yuv_table: RY GY BY RU GU BU RV GV BV (swscale_internal.h:468-477), 15-bit */
    int bgr24_yv12;            /* 1: reference installs bgr24ToYv12Wrapper (ff_rgb24toyv12_c) as convert_unscaled */
    int dst_nv;                /* 0: three destination planes; 1: nv12, 2: nv21 destination (plane 1 = U,V / V,U interleaved) */
};

static int64_t i64abs(int64_t a) { return a < 0 ? -a : a; }
static int ilog2(unsigned v) { int n = 0; v |= 1; while (v >>= 1) n++; return n; }
static int clip_u8(int a) { return a < 0 ? 0 : a > 255 ? 255 : a; }
static int64_t rounded_div(int64_t a, int64_t b) { return a >= 0 ? (a + (b >> 1)) / b : (a - (b >> 1)) / b; }
static int ceil_rshift(int a, int b) { return -((-a) >> b); }

/* utils.c:168-175 */
static int local_pos(int chr_subsample, int pos)
{
    if (pos == -1 || pos <= -513)
        pos = (128 << chr_subsample) - 128;
    pos += 128;
    return pos >> chr_subsample;
}

/* utils.c:197-612 with srcFilter = dstFilter = NULL, param = defaults, filterAlign = 1 */
/* SwsContext.scaler_params (sws_getContext's `param`, swscale.h): SWS_PARAM_DEFAULT = 123456 selects each kernel's default.
 * The checker is single threaded: orc_sws_open_params() sets them around the open call. */
#define ORC_PARAM_DEFAULT 123456
static double g_param[2] = { ORC_PARAM_DEFAULT, ORC_PARAM_DEFAULT };
/* sws_getContext's srcFilter / dstFilter vectors for the open call in flight ([0] lumH, [1] lumV, [2] chrH, [3] chrV), set by
 * orc_sws_open_filters(); make_filter() is told which one applies */
static const double *g_src_vec[4];
static int g_src_len[4], g_dst_len[4];
static int g_vec_sel = -1;

static int make_filter(int16_t **outFilter, int32_t **outPos, int *outSize, int xInc, int srcW, int dstW,
                       int one, int scaler, int flags, int srcPos, int dstPos)
{
    const int64_t fone = 1LL << (54 - (ilog2(srcW / dstW) < 8 ? ilog2(srcW / dstW) : 8));
    int32_t *pos = calloc((size_t)dstW + 3, sizeof(*pos));
    int64_t *f = NULL;
    int fs, i, j;
    if (!pos) return -ENOMEM;

    if (abs(xInc - 0x10000) < 10 && srcPos == dstPos) {            /* :221-231 identity */
        fs = 1;
        f = calloc(dstW, sizeof(*f));
        for (i = 0; i < dstW; i++) { f[i] = fone; pos[i] = i; }
    } else if (scaler == ORC_SWS_POINT) {                          /* :232-246 */
        int64_t x = ((dstPos * (int64_t)xInc) >> 8) - ((srcPos * 0x8000LL) >> 7);
        fs = 1;
        f = calloc(dstW, sizeof(*f));
        for (i = 0; i < dstW; i++) {
            pos[i] = (int)((x - ((int64_t)(fs - 1) << 15) + (1 << 15)) >> 16);
            f[i] = fone;
            x += xInc;
        }
    } else if ((xInc <= (1 << 16) && scaler == ORC_SWS_AREA) || scaler == ORC_SWS_FAST_BILINEAR) {   /* :247-271 */
        int64_t x = ((dstPos * (int64_t)xInc) >> 8) - ((srcPos * 0x8000LL) >> 7);
        fs = 2;
        f = calloc((size_t)dstW * fs, sizeof(*f));
        for (i = 0; i < dstW; i++) {
            int xx = (int)((x - ((int64_t)(fs - 1) << 15) + (1 << 15)) >> 16);
            pos[i] = xx;
            for (j = 0; j < fs; j++) {
                int64_t coeff = fone - i64abs((int64_t)xx * (1 << 16) - x) * (fone >> 16);
                f[i * fs + j] = coeff < 0 ? 0 : coeff;
                xx++;
            }
            x += xInc;
        }
    } else {                                                       /* :272-383 */
        /* scale_algorithms[] (utils.c:183-195); Lanczos: 2 * param[0] with the default 3 (:278-279) */
        int sizeFactor = scaler == ORC_SWS_BICUBIC ? 4 : scaler == ORC_SWS_BILINEAR ? 2 : scaler == ORC_SWS_AREA ? 1 :
                         scaler == ORC_SWS_GAUSS || scaler == ORC_SWS_X ? 8 : scaler == ORC_SWS_SINC || scaler == ORC_SWS_SPLINE ? 20 :
                         scaler == ORC_SWS_LANCZOS ? (g_param[0] != ORC_PARAM_DEFAULT ? (int)ceil(2 * g_param[0]) : 6) : -1;
        int64_t x;
        if (sizeFactor < 0) { free(pos); return -ENOSYS; }
        if (sizeFactor > 50 || sizeFactor == 0) { free(pos); return -EINVAL; }   /* :282-285 (0 trips the reference's assert) */
        if (xInc <= 1 << 16) fs = 1 + sizeFactor;
        else                 fs = 1 + (int)(((int64_t)sizeFactor * srcW + dstW - 1) / dstW);
        if (fs > srcW - 2) fs = srcW - 2;
        if (fs < 1) fs = 1;
        f = calloc((size_t)dstW * fs, sizeof(*f));
        x = ((dstPos * (int64_t)xInc) >> 7) - ((srcPos * 0x10000LL) >> 7);
        for (i = 0; i < dstW; i++) {
            int xx = (int)((x - (fs - 2) * (1LL << 16)) / (1 << 17));
            pos[i] = xx;
            for (j = 0; j < fs; j++) {
                int64_t d = i64abs(((int64_t)xx * (1 << 17)) - x) << 13;
                int64_t coeff;
                if (xInc > 1 << 16)
                    d = d * dstW / srcW;
                if (scaler == ORC_SWS_BICUBIC) {                   /* B = 0, C = 0.6 in 8.24 */
                    int64_t B = (g_param[0] != ORC_PARAM_DEFAULT ? g_param[0] : 0) * (1 << 24);
                    int64_t C = (g_param[1] != ORC_PARAM_DEFAULT ? g_param[1] : 0.6) * (1 << 24);
                    if (d >= 1LL << 31) {
                        coeff = 0;
                    } else {
                        int64_t dd  = (d * d) >> 30;
                        int64_t ddd = (dd * d) >> 30;
                        if (d < 1LL << 30)
                            coeff = (12 * (1 << 24) - 9 * B - 6 * C) * ddd +
                                    (-18 * (1 << 24) + 12 * B + 6 * C) * dd +
                                    (6 * (1 << 24) - 2 * B) * (1LL << 30);
                        else
                            coeff = (-B - 6 * C) * ddd + (6 * B + 30 * C) * dd +
                                    (-12 * B - 48 * C) * d + (8 * B + 24 * C) * (1LL << 30);
                    }
                    coeff /= (1LL << 54) / fone;
                } else if (scaler == ORC_SWS_AREA) {
                    int64_t d2 = d - (1 << 29);
                    if (d2 * xInc < -(1LL << (29 + 16)))      coeff = 1LL << (30 + 16);
                    else if (d2 * xInc < (1LL << (29 + 16)))  coeff = -d2 * xInc + (1LL << (29 + 16));
                    else                                      coeff = 0;
                    coeff *= fone >> (30 + 16);
                } else if (scaler == ORC_SWS_BILINEAR) {
                    coeff = (1 << 30) - d;
                    if (coeff < 0) coeff = 0;
                    coeff *= fone >> 30;
                } else {                                           /* the kernels evaluated in double precision (:325-368), default parameters */
                    const double fd = d * (1.0 / (1 << 30));
                    if (scaler == ORC_SWS_X) {
                        double c = fd < 1.0 ? cos(fd * M_PI) : -1.0;
                        const double A = g_param[0] != ORC_PARAM_DEFAULT ? g_param[0] : 1.0;
                        c = c < 0.0 ? -pow(-c, A) : pow(c, A);
                        coeff = (c * 0.5 + 0.5) * fone;
                    } else if (scaler == ORC_SWS_GAUSS) {
                        const double p = g_param[0] != ORC_PARAM_DEFAULT ? g_param[0] : 3.0;
                        coeff = exp2(-p * fd * fd) * fone;
                    } else if (scaler == ORC_SWS_SINC) {
                        coeff = (d ? sin(fd * M_PI) / (fd * M_PI) : 1.0) * fone;
                    } else if (scaler == ORC_SWS_LANCZOS) {
                        const double p = g_param[0] != ORC_PARAM_DEFAULT ? g_param[0] : 3.0;
                        coeff = (d ? sin(fd * M_PI) * sin(fd * M_PI / p) / (fd * fd * M_PI * M_PI / p) : 1.0) * fone;
                        if (fd > p) coeff = 0;
                    } else {                                       /* spline: getSplineCoeff(1, 0, p, -p - 1, dist), utils.c:155-167 */
                        double a = 1.0, b = 0.0, c = -2.196152422706632, e = 2.196152422706632 - 1.0, dist = fd;
                        while (dist > 1.0) {
                            const double nb = b + 2.0 * c + 3.0 * e, nc = c + 3.0 * e, ne = -b - 3.0 * c - 6.0 * e;
                            a = 0.0; b = nb; c = nc; e = ne; dist -= 1.0;
                        }
                        coeff = (((e * dist + c) * dist + b) * dist + a) * fone;
                    }
                }
                f[i * fs + j] = coeff;
                xx++;
            }
            x += 2LL * xInc;
        }
    }

    /* srcFilter / dstFilter, utils.c:384-413: the source vector convolved into every row (`filter2[...] += coeff * filter[...]` on an int64:
     * through a double each time), the destination vector only widens the row ("FIXME dstFilter"); the window start moves */
    if (g_vec_sel >= 0 && (g_src_len[g_vec_sel] > 0 || g_dst_len[g_vec_sel] > 0)) {
        const int sl = g_src_len[g_vec_sel], dl = g_dst_len[g_vec_sel];
        const int fs2 = fs + (sl ? sl - 1 : 0) + (dl ? dl - 1 : 0);
        int64_t *f2 = calloc((size_t)dstW * fs2, sizeof(*f2));
        for (i = 0; i < dstW; i++) {
            if (sl) {
                for (int k = 0; k < sl; k++)
                    for (j = 0; j < fs; j++)
                        f2[i * fs2 + k + j] += g_src_vec[g_vec_sel][k] * f[i * fs + j];
            } else
                for (j = 0; j < fs; j++) f2[i * fs2 + j] = f[i * fs + j];
            pos[i] += (fs - 1) / 2 - (fs2 - 1) / 2;
        }
        free(f);
        f = f2; fs = fs2;
    }

    /* size reduction, utils.c:417-457: strip near-zero taps left, count them right */
    int minSize = 0;
    for (i = dstW - 1; i >= 0; i--) {
        int min = fs;
        int64_t cut = 0;
        for (j = 0; j < fs; j++) {
            int k;
            cut += i64abs(f[i * fs]);
            if ((double)cut > 0.002 * (double)fone) break;
            if (i < dstW - 1 && pos[i] >= pos[i + 1]) break;
            for (k = 1; k < fs; k++) f[i * fs + k - 1] = f[i * fs + k];
            f[i * fs + k - 1] = 0;
            pos[i]++;
        }
        cut = 0;
        for (j = fs - 1; j > 0; j--) {
            cut += i64abs(f[i * fs + j]);
            if ((double)cut > 0.002 * (double)fone) break;
            min--;
        }
        if (min > minSize) minSize = min;
    }
    int nfs = minSize;                                             /* filterAlign == 1 */
    if (nfs >= 256) { free(f); free(pos); return -ENOSYS; }        /* reference would cascade (:491-495) */
    int64_t *g = calloc((size_t)dstW * nfs, sizeof(*g));
    for (i = 0; i < dstW; i++)
        for (j = 0; j < nfs; j++)
            g[i * nfs + j] = j >= fs ? 0 : f[i * fs + j];
    free(f);

    /* border folding, utils.c:519-560 */
    for (i = 0; i < dstW; i++) {
        if (pos[i] < 0) {
            for (j = 1; j < nfs; j++) {
                int left = j + pos[i] > 0 ? j + pos[i] : 0;
                g[i * nfs + left] += g[i * nfs + j];
                g[i * nfs + j] = 0;
            }
            pos[i] = 0;
        }
        if (pos[i] + nfs > srcW) {
            int shift = pos[i] + (nfs - srcW < 0 ? nfs - srcW : 0);
            int64_t acc = 0;
            for (j = nfs - 1; j >= 0; j--)
                if (pos[i] + j >= srcW) { acc += g[i * nfs + j]; g[i * nfs + j] = 0; }
            for (j = nfs - 1; j >= 0; j--)
                g[i * nfs + j] = j < shift ? 0 : g[i * nfs + j - shift];
            pos[i] -= shift;
            g[i * nfs + srcW - 1 - pos[i]] += acc;
        }
    }

    /* error-diffused normalisation to `one`, utils.c:568-588 */
    int16_t *out = calloc((size_t)(dstW + 3) * nfs, sizeof(*out));
    for (i = 0; i < dstW; i++) {
        int64_t error = 0, sum = 0;
        for (j = 0; j < nfs; j++) sum += g[i * nfs + j];
        sum = (sum + one / 2) / one;
        if (!sum) sum = 1;
        for (j = 0; j < nfs; j++) {
            int64_t v = g[i * nfs + j] + error;
            int iv = (int)rounded_div(v, sum);
            out[i * nfs + j] = (int16_t)iv;
            error = v - iv * sum;
        }
    }
    free(g);
    pos[dstW] = pos[dstW + 1] = pos[dstW + 2] = pos[dstW - 1];
    for (i = 0; i < nfs; i++) {
        int k = (dstW - 1) * nfs + i;
        out[k + nfs] = out[k + 2 * nfs] = out[k + 3 * nfs] = out[k];
    }
    *outFilter = out; *outPos = pos; *outSize = nfs;
    return 0;
}

/* yuv2rgb.c:717-914 (bpp 24), fill_table :680-692, fill_gv_table :694-703 */
int orc_sws_set_colorspace(OrcSws *s, const int inv_table[4], int fullRange,
                           int brightness, int contrast, int saturation)
{
    const int yoffs = (fullRange ? 384 : 326) + LUMA_HEADROOM;
    int64_t crv = inv_table[0], cbu = inv_table[1], cgu = -inv_table[2], cgv = -inv_table[3];
    int64_t cy = 1 << 16, oy = 0, yb;
    int i;
    if (!fullRange) {
        cy = (cy * 255) / 219;
        oy = 16 << 16;
    } else {
        crv = (crv * 224) / 255; cbu = (cbu * 224) / 255;
        cgu = (cgu * 224) / 255; cgv = (cgv * 224) / 255;
    }
    cy  = (cy * contrast) >> 16;
    crv = (crv * contrast * saturation) >> 32;
    cbu = (cbu * contrast * saturation) >> 32;
    cgu = (cgu * contrast * saturation) >> 32;
    cgv = (cgv * contrast * saturation) >> 32;
    oy -= 256LL * brightness;

#define R16(f) ({ int64_t f_ = (f); int r_ = (int)((f_ + (1 << 15)) >> 16); \
                  (int16_t)(uint16_t)(r_ < -0x7FFF ? 0x8000 : r_ > 0x7FFF ? 0x7FFF : r_); })
    s->y_coeff  = R16(cy  * (1 << 13));
    s->y_offset = R16(oy  * (1 <<  9));
    s->v2r      = R16(crv * (1 << 13));
    s->v2g      = R16(cgv * (1 << 13));
    s->u2g      = R16(cgu * (1 << 13));
    s->u2b      = R16(cbu * (1 << 13));
#undef R16

    int64_t cyd = cy > 1 ? cy : 1;
    crv = ((crv * (1 << 16)) + 0x8000) / cyd;
    cbu = ((cbu * (1 << 16)) + 0x8000) / cyd;
    cgu = ((cgu * (1 << 16)) + 0x8000) / cyd;
    cgv = ((cgv * (1 << 16)) + 0x8000) / cyd;

    yb = -(384 << 16) - LUMA_HEADROOM * cy - oy;
    for (i = 0; i < YTAB_SIZE; i++) {
        s->ytab[i] = (uint8_t)clip_u8((int)((yb + 0x8000) >> 16));
        yb += cy;
    }
    for (i = 0; i < NTAB; i++) {
        int64_t c8 = clip_u8(i - HEADROOM);
        s->offR[i]  = (int)(yoffs - (crv >> 9) + ((c8 * crv) >> 16));
        s->offGU[i] = (int)(yoffs - (cgu >> 9) + ((c8 * cgu) >> 16));
        s->offB[i]  = (int)(yoffs - (cbu >> 9) + ((c8 * cbu) >> 16));
        s->offGV[i] = (int)(-(cgv >> 9) + ((c8 * cgv) >> 16));
    }
    return 0;
}

static const int default_coeffs[4] = { 104597, 132201, 25675, 53279 };  /* yuv2rgb.c:47-59 row SWS_CS_DEFAULT */

/* solve_range_convert (swscale.c:577-589) for an 8-bit destination: src_bits 15, src_shift 7, mult_shift 14
 * (init_range_convert_constants, :591-600).  The line functions take the coefficient as uint16 and the offset as int32. */
static void solve_range(unsigned src_min, unsigned src_max, unsigned dst_min, unsigned dst_max, int *coeff, int *offset)
{
    const int src_shift = 7, mult_shift = 14, total_shift = mult_shift + src_shift;
    const uint64_t src_range = src_max - src_min, dst_range = dst_max - dst_min;
    const uint64_t q = (dst_range << total_shift) / src_range;
    const uint32_t c = (uint32_t)((q + (1u << src_shift) - 1) >> src_shift);                   /* AV_CEIL_RSHIFT */
    const int64_t  o = ((int64_t)dst_max << total_shift) - ((int64_t)src_max << src_shift) * c + (1u << (mult_shift - 1));
    *coeff  = (uint16_t)c;
    *offset = (int32_t)o;
}

/* ff_sws_init_range_convert (swscale.c:626-660) */
static void init_range_convert(OrcSws *s)
{
    s->range_conv = 0;
    if (s->src_range == s->dst_range || !s->planar) return;
    if (s->src_range) {                                              /* full -> limited */
        solve_range(0, 255, 16, 235, &s->lum_coeff, &s->lum_offset);
        solve_range(0, 255, 16, 240, &s->chr_coeff, &s->chr_offset);
        s->range_conv = 2;
    } else {                                                         /* limited -> full */
        solve_range(16, 235, 0, 255, &s->lum_coeff, &s->lum_offset);
        solve_range(16, 240, 0, 255, &s->chr_coeff, &s->chr_offset);
        s->range_conv = 1;
    }
}

/* fill_rgb2yuv_table (utils.c:614-700): always the limited-range matrix ("range = 1 is handled elsewhere", i.e. by the range
 * conversion between the passes); the default BT.601 table gets the hand-rounded constants of :692-702 */
static void fill_rgb2yuv(OrcSws *s, const int table[4])
{
    const int64_t ONE = 65536;
    int64_t vr = table[0], ub = table[1], ug = -table[2], vg = -table[3], cy = ONE * 255 / 219;
    int64_t W = rounded_div(ONE * ONE * ug, ub), V = rounded_div(ONE * ONE * vg, vr), Z = ONE * ONE - W - V;
    int64_t Cy = rounded_div(cy * Z, ONE), Cu = rounded_div(ub * Z, ONE), Cv = rounded_div(vr * Z, ONE);
    int *t = s->rgb2yuv;
    t[0] = (int)-rounded_div((1 << 15) * V, Cy);       t[1] = (int)rounded_div((1 << 15) * ONE * ONE, Cy);  t[2] = (int)-rounded_div((1 << 15) * W, Cy);
    t[3] = (int)rounded_div((1 << 15) * V, Cu);        t[4] = (int)-rounded_div((1 << 15) * ONE * ONE, Cu); t[5] = (int)rounded_div((1 << 15) * (Z + W), Cu);
    t[6] = (int)rounded_div((1 << 15) * (V + Z), Cv);  t[7] = (int)-rounded_div((1 << 15) * ONE * ONE, Cv); t[8] = (int)rounded_div((1 << 15) * W, Cv);
    if (!memcmp(table, default_coeffs, sizeof(default_coeffs))) {
        t[2] =  (int)(0.114 * 219 / 255 * (1 << 15) + 0.5); t[8] = -(int)(0.081 * 224 / 255 * (1 << 15) + 0.5);
        t[5] =  (int)(0.500 * 224 / 255 * (1 << 15) + 0.5); t[1] =  (int)(0.587 * 219 / 255 * (1 << 15) + 0.5);
        t[7] = -(int)(0.419 * 224 / 255 * (1 << 15) + 0.5); t[4] = -(int)(0.331 * 224 / 255 * (1 << 15) + 0.5);
        t[0] =  (int)(0.299 * 219 / 255 * (1 << 15) + 0.5); t[6] =  (int)(0.500 * 224 / 255 * (1 << 15) + 0.5);
        t[3] = -(int)(0.169 * 224 / 255 * (1 << 15) + 0.5);
    }
}

/* sws_setColorspaceDetails (utils.c:849-1004) */
int orc_sws_set_colorspace_details(OrcSws *s, const int inv_table[4], int srcRange, const int table[4], int dstRange,
                                   int brightness, int contrast, int saturation)
{
    if (s->casc[0])                                                  /* utils.c:908-909: a cascaded context hands the call to its main child */
        return orc_sws_set_colorspace_details(s->casc[0], inv_table, srcRange, table, dstRange, brightness, contrast, saturation);
    if (!s->planar) dstRange = 0;                                    /* range_override_needed(dst), utils.c:877-878 */
    if (s->src_rgb) srcRange = 0;                                    /* range_override_needed(src), utils.c:879-880 */
    fill_rgb2yuv(s, table);                                          /* utils.c:1002 */
    memcpy(s->src_cs, inv_table, sizeof(s->src_cs));
    memcpy(s->dst_cs, table, sizeof(s->dst_cs));
    s->src_range = srcRange;
    s->dst_range = dstRange;
    init_range_convert(s);
    if (s->planar && !s->src_rgb) {
        /* utils.c:914-989: yuv -> yuv with different matrices goes through an intermediate bgr24 picture of the smaller of the two
         * sizes: context 0 = source -> bgr24 with these details (the RGB side ignores its half), context 1 = bgr24 -> destination with
         * the ranges set before its initialisation and the details again without brightness / contrast / saturation */
        if (!memcmp(s->src_cs, s->dst_cs, sizeof(s->src_cs))) return 0;
        if (s->casc[0]) return -1;                                   /* "we need to cascade more contexts to compensate" (:985-987) */
        const int big = (int64_t)s->srcW * s->srcH > (int64_t)s->dstW * s->dstH;
        s->casc_w = big ? s->dstW : s->srcW; s->casc_h = big ? s->dstH : s->srcH;
        s->casc[0] = orc_sws_open_params(s->open_src_fmt, s->srcW, s->srcH, 0, ORC_PIX_FMT_BGR24, s->casc_w, s->casc_h, 0, s->open_flags, s->open_param);
        if (!s->casc[0]) return -1;
        orc_sws_set_colorspace_details(s->casc[0], inv_table, srcRange, table, dstRange, brightness, contrast, saturation);
        s->casc[1] = orc_sws_open_params(ORC_PIX_FMT_BGR24, s->casc_w, s->casc_h, srcRange, s->open_dst_fmt, s->dstW, s->dstH, dstRange, s->open_flags, s->open_param);
        if (!s->casc[1]) return -1;
        orc_sws_set_colorspace_details(s->casc[1], inv_table, srcRange, table, dstRange, 0, 1 << 16, 1 << 16);
        return 0;
    }
    if (s->planar) return 0;                                         /* RGB -> yuv: only the rgb2yuv table and the ranges matter */
    return orc_sws_set_colorspace(s, inv_table, srcRange, brightness, contrast, saturation);
}

OrcSws *orc_sws_open(int srcW, int srcH, int dstW, int dstH, int flags)
{
    return orc_sws_open_fmt(srcW, srcH, dstW, dstH, ORC_PIX_FMT_RGB24, flags);
}

/* Byte order of the packed 8-bit RGB outputs.  The 32-bit formats carry alpha = 255 for a source without alpha: the
 * reference adds 255 << abase to every luma-ramp entry (yuv2rgb.c:947-960) and yuv2rgb_write_full stores 255
 * (output.c:2066-2095).  R, G and B are the same values as for rgb24: the 32 bpp ramp is the 24 bpp one shifted
 * (yuv2rgb.c:901-914 vs :947-966), and yuv2rgb_write sums r[Y] + g[Y] + b[Y] (output.c:1676-1695). */
static int set_format(OrcSws *s, int fmt)
{
    s->planar = 0;
    switch (fmt) {
    case ORC_PIX_FMT_YUV420P: s->planar = 1; s->bpp = 1; s->ro = s->go = s->bo = 0; s->ao = -1; break;
    case ORC_PIX_FMT_NV12: case ORC_PIX_FMT_NV21:                   /* as a destination: yuv420p with the chroma planes interleaved */
        s->planar = 1; s->bpp = 1; s->ro = s->go = s->bo = 0; s->ao = -1; break;
    case ORC_PIX_FMT_RGB24: s->bpp = 3; s->ro = 0; s->go = 1; s->bo = 2; s->ao = -1; break;
    case ORC_PIX_FMT_BGR24: s->bpp = 3; s->ro = 2; s->go = 1; s->bo = 0; s->ao = -1; break;
    case ORC_PIX_FMT_RGBA:  s->bpp = 4; s->ro = 0; s->go = 1; s->bo = 2; s->ao = 3;  break;
    case ORC_PIX_FMT_BGRA:  s->bpp = 4; s->ro = 2; s->go = 1; s->bo = 0; s->ao = 3;  break;
    case ORC_PIX_FMT_ARGB:  s->bpp = 4; s->ro = 1; s->go = 2; s->bo = 3; s->ao = 0;  break;
    case ORC_PIX_FMT_ABGR:  s->bpp = 4; s->ro = 3; s->go = 2; s->bo = 1; s->ao = 0;  break;
    default: return -1;
    }
    return 0;
}

OrcSws *orc_sws_open_fmt(int srcW, int srcH, int dstW, int dstH, int dstFormat, int flags)
{
    return orc_sws_open_io(ORC_PIX_FMT_YUV420P, srcW, srcH, dstFormat, dstW, dstH, flags);
}

OrcSws *orc_sws_open_io(int srcFormat, int srcW, int srcH, int dstFormat, int dstW, int dstH, int flags)
{
    return orc_sws_open_range(srcFormat, srcW, srcH, 0, dstFormat, dstW, dstH, 0, flags);
}

OrcSws *orc_sws_open_params(int srcFormat, int srcW, int srcH, int srcRange, int dstFormat, int dstW, int dstH, int dstRange, int flags,
                            const double *param)
{
    if (param) { g_param[0] = param[0]; g_param[1] = param[1]; }
    OrcSws *s = orc_sws_open_range(srcFormat, srcW, srcH, srcRange, dstFormat, dstW, dstH, dstRange, flags);
    g_param[0] = g_param[1] = ORC_PARAM_DEFAULT;
    return s;
}

/* srcFilter / dstFilter as four (coefficients, length) pairs each: lumH, lumV, chrH, chrV; NULL / 0 = no vector */
OrcSws *orc_sws_open_filters(int srcFormat, int srcW, int srcH, int srcRange, int dstFormat, int dstW, int dstH, int dstRange, int flags,
                             const double *const srcCoef[4], const int srcLen[4], const int dstLen[4], const double *param)
{
    for (int k = 0; k < 4; k++) {
        g_src_vec[k] = srcCoef ? srcCoef[k] : NULL;
        g_src_len[k] = srcCoef && srcLen && srcCoef[k] ? srcLen[k] : 0;
        g_dst_len[k] = dstLen ? dstLen[k] : 0;
    }
    OrcSws *s = orc_sws_open_params(srcFormat, srcW, srcH, srcRange, dstFormat, dstW, dstH, dstRange, flags, param);
    for (int k = 0; k < 4; k++) { g_src_vec[k] = NULL; g_src_len[k] = g_dst_len[k] = 0; }
    return s;
}

OrcSws *orc_sws_open_range(int srcFormat, int srcW, int srcH, int srcRange, int dstFormat, int dstW, int dstH, int dstRange, int flags)
{
    if (srcW < 1 || srcH < 1 || dstW < 1 || dstH < 1) return NULL;
    OrcSws *s = calloc(1, sizeof(*s));
    if (!s) return NULL;
    s->open_src_fmt = srcFormat; s->open_dst_fmt = dstFormat; s->open_flags = flags;
    s->open_param[0] = g_param[0]; s->open_param[1] = g_param[1];
    if (set_format(s, srcFormat) == 0 && !s->planar) {              /* packed RGB source: remember its byte layout */
        s->src_rgb = s->bpp; s->sro = s->ro; s->sgo = s->go; s->sbo = s->bo;
    } else if (srcFormat != ORC_PIX_FMT_YUV420P && srcFormat != ORC_PIX_FMT_NV12 && srcFormat != ORC_PIX_FMT_NV21) { free(s); return NULL; }
    if (set_format(s, dstFormat) < 0) { free(s); return NULL; }
    s->dst_nv = dstFormat == ORC_PIX_FMT_NV12 ? 1 : dstFormat == ORC_PIX_FMT_NV21 ? 2 : 0;
    int uses_filter = 0;                                           /* utils.c:1256-1263: usesVFilter || usesHFilter */
    for (int k = 0; k < 4; k++) uses_filter |= g_src_len[k] > 1 || g_dst_len[k] > 1;
    if (!uses_filter && s->src_rgb && !s->planar && srcW == dstW && srcH == dstH) {
        /* same size, packed RGB on both sides: packedCopyWrapper for equal formats (swscale_unscaled.c:2675-2690), else rgbToRgbWrapper
         * (:2001-2060) whenever findRgbConvFn (:1843-1998) returns a function.  For the six 8-bit formats here it always does --
         * shuffle_bytes_* between the 32-bit orders, rgb24tobgr24, rgb32to24 / rgb32tobgr24, rgb24to32 / rgb24tobgr32 -- except under
         * SWS_BITEXACT from 24 bits to AV_PIX_FMT_RGB32 / BGR32 (bgra / rgba on little endian; :1992-1995), which the scaler handles.
         * Every one of them moves R, G, B (and A) of a pixel to their places in the destination order; a destination alpha without a
         * source alpha is 255 (rgb24to32 family, rgb2rgb_template.c; the first byte of each line for argb / abgr, :2030-2036). */
        const int to32 = s->src_rgb == 3 && s->bpp == 4;
        if (!(to32 && s->ao == 3 && (flags & ORC_SWS_BITEXACT))) {
            s->srcW = srcW; s->srcH = srcH; s->dstW = dstW; s->dstH = dstH; s->flags = flags;
            s->chrSrcW = s->chrDstW = srcW; s->chrSrcH = s->chrDstH = srcH;
            s->unscaled_lut = 4;
            return s;
        }
    }
    s->need_alpha = s->src_rgb == 4 && !s->planar && s->bpp == 4;
    s->src_nv = srcFormat == ORC_PIX_FMT_NV12 ? 1 : srcFormat == ORC_PIX_FMT_NV21 ? 2 : 0;
    int algo = flags & 0x7FF;
    if (!algo) { algo = ORC_SWS_BICUBIC; flags |= algo; }           /* utils.c:1209-1217 */
    else if (algo & (algo - 1)) goto fail;
    if (algo == ORC_SWS_FAST_BILINEAR && (srcW < 8 || dstW <= 8)) { /* utils.c:1224-1230 */
        algo = ORC_SWS_BILINEAR;
        flags ^= ORC_SWS_FAST_BILINEAR | algo;
    }
    s->fast_bilinear = algo == ORC_SWS_FAST_BILINEAR && !s->src_rgb;  /* hyscale_fast / hcscale_fast exist for 8-bit input lines only
                                                                    * (swscale.c:675-681); RGB sources are srcBpc 16 (utils.c:1407-1408)
                                                                    * and use the 2-tap filter initFilter builds for the flag */
    if (!s->planar && (dstW & 1)) flags |= ORC_SWS_FULL_CHR_H_INT; /* utils.c:1271-1276 (RGB destinations only) */
    if (!s->planar && s->src_rgb && !(flags & ORC_SWS_FAST_BILINEAR))
        flags |= ORC_SWS_FULL_CHR_H_INT;                           /* utils.c:1277-1285: source chroma is not subsampled */
    s->srcW = srcW; s->srcH = srcH; s->dstW = dstW; s->dstH = dstH; s->flags = flags;
    s->chrDstHSub = (!s->planar && (flags & ORC_SWS_FULL_CHR_H_INT)) ? 0 : 1;      /* utils.c:1359-1360 */
    const int chrDstVSub = s->planar ? 1 : 0;                      /* av_pix_fmt_get_chroma_sub_sample(dstFormat), utils.c:1266 */
    s->chrSrcHSub = 1; s->chrSrcVSub = 1;
    if (s->src_rgb) {                                              /* utils.c:1366-1393: every other pixel for chroma unless asked otherwise */
        s->chrSrcVSub = 0;
        s->chrSrcHSub = !(srcW & 1) && !(flags & ORC_SWS_FULL_CHR_H_INP) &&
                        ((dstW >> s->chrDstHSub) <= (srcW >> 1) || (flags & ORC_SWS_FAST_BILINEAR));
    }
    s->chrSrcW = ceil_rshift(srcW, s->chrSrcHSub); s->chrSrcH = ceil_rshift(srcH, s->chrSrcVSub);
    s->chrDstW = ceil_rshift(dstW, s->chrDstHSub); s->chrDstH = ceil_rshift(dstH, chrDstVSub);
    /* utils.c:1164-1167: the ranges given before initialisation go through sws_setColorspaceDetails */
    orc_sws_set_colorspace_details(s, default_coeffs, srcRange != 0, default_coeffs, dstRange != 0, 0, 1 << 16, 1 << 16);

    /* the unscaled converters are only looked for when no range conversion is due (utils.c:1623-1626) */
    /* bgr24ToYv12Wrapper (swscale_unscaled.c:2453-2457): bgr24 only, not with accurate_rnd, even width */
    if (!uses_filter && s->planar && !s->dst_nv && s->src_rgb == 3 && s->sbo == 0 && srcW == dstW && srcH == dstH && s->src_range == s->dst_range &&
        !(flags & ORC_SWS_ACCURATE_RND) && !(dstW & 1)) {
        s->bgr24_yv12 = 1;
        s->unscaled_lut = 3;
        return s;
    }
    /* planarCopyWrapper (swscale_unscaled.c:2675-2693), planarToNv12Wrapper (:147-165) or, for a semi-planar source,
     * nv12ToPlanarWrapper (:167-188, :2415-2419): all a luma copy plus a chroma copy / interleave / de-interleave.  nv12 <-> nv21
     * has no such converter: it goes through the scaler (same bytes, until sws_setColorspaceDetails changes a range) */
    if (!uses_filter && s->planar && !s->src_rgb && srcW == dstW && srcH == dstH && s->src_range == s->dst_range &&
        !(s->src_nv && s->dst_nv && s->src_nv != s->dst_nv)) {
        s->unscaled_lut = 2;
        return s;
    }
    /* swscale_unscaled.c:2426-2431 through utils.c:1623-1637: only planar yuv420p/422p sources have the LUT converter
     * (packed RGB destination: dst_range was forced to 0 above; isAnyRGB(dst) passes the range test of utils.c:1625) */
    if (!uses_filter && !s->planar && !s->src_nv && !s->src_rgb && srcW == dstW && srcH == dstH && !(flags & ORC_SWS_ACCURATE_RND) && !(dstH & 1)) {
        s->unscaled_lut = 1;
        return s;
    }
    int lum_scaler = algo == ORC_SWS_BICUBLIN ? ORC_SWS_BICUBIC : algo;
    int chr_scaler = algo == ORC_SWS_BICUBLIN ? ORC_SWS_BILINEAR : algo;
    int64_t lumXInc = (((int64_t)srcW << 16) + (dstW >> 1)) / dstW;
    int64_t lumYInc = (((int64_t)srcH << 16) + (dstH >> 1)) / dstH;
    int64_t chrXInc = (((int64_t)s->chrSrcW << 16) + (s->chrDstW >> 1)) / s->chrDstW;
    int64_t chrYInc = (((int64_t)s->chrSrcH << 16) + (s->chrDstH >> 1)) / s->chrDstH;
    s->lumXInc = (int)lumXInc; s->chrXInc = (int)chrXInc;
    g_vec_sel = 0;
    if (make_filter(&s->hLum, &s->hLumPos, &s->hLumSize, (int)lumXInc, srcW, dstW, 1 << 14, lum_scaler, flags,
                    local_pos(0, 0), local_pos(0, 0)) < 0) goto fail;
    g_vec_sel = 2;
    if (make_filter(&s->hChr, &s->hChrPos, &s->hChrSize, (int)chrXInc, s->chrSrcW, s->chrDstW, 1 << 14, chr_scaler, flags,
                    local_pos(s->chrSrcHSub, -513), local_pos(s->chrDstHSub, -513)) < 0) goto fail;
    g_vec_sel = 1;
    if (make_filter(&s->vLum, &s->vLumPos, &s->vLumSize, (int)lumYInc, srcH, dstH, 1 << 12, lum_scaler, flags,
                    local_pos(0, 0), local_pos(0, 0)) < 0) goto fail;
    g_vec_sel = 3;
    if (make_filter(&s->vChr, &s->vChrPos, &s->vChrSize, (int)chrYInc, s->chrSrcH, s->chrDstH, 1 << 12, chr_scaler, flags,
                    local_pos(s->chrSrcVSub, -513), local_pos(chrDstVSub, -513)) < 0) goto fail;
    g_vec_sel = -1;
    return s;
fail:
    g_vec_sel = -1;
    orc_sws_close(s);
    return NULL;
}

void orc_sws_close(OrcSws *s)
{
    if (!s) return;
    orc_sws_close(s->casc[0]); orc_sws_close(s->casc[1]);
    free(s->hLum); free(s->hChr); free(s->vLum); free(s->vChr);
    free(s->hLumPos); free(s->hChrPos); free(s->vLumPos); free(s->vChrPos);
    free(s);
}

int orc_sws_info(const OrcSws *s, int *out)
{
    out[0] = s->hLumSize; out[1] = s->hChrSize; out[2] = s->vLumSize; out[3] = s->vChrSize;
    out[4] = s->chrSrcW; out[5] = s->chrSrcH; out[6] = s->chrDstW; out[7] = s->chrDstH;
    out[8] = s->unscaled_lut; out[9] = 1; out[10] = 1; out[11] = s->chrDstHSub; out[12] = 0;
    out[13] = s->dstW; out[14] = s->dstH; out[15] = 0;
    return 0;
}

/* out[0] range conversion in use (0 none, 1 limited -> full, 2 full -> limited), [1..4] luma coefficient, luma offset,
 * chroma coefficient, chroma offset as the line functions see them, [5] an unscaled converter was installed */
int orc_sws_range_info(const OrcSws *s, int *out)
{
    out[0] = s->range_conv;
    out[1] = s->range_conv ? s->lum_coeff : 0; out[2] = s->range_conv ? s->lum_offset : 0;
    out[3] = s->range_conv ? s->chr_coeff : 0; out[4] = s->range_conv ? s->chr_offset : 0;
    out[5] = s->unscaled_lut != 0;
    return 0;
}

/* out[0] bytes per pixel of a packed RGB source (0: yuv source), [1] / [2] horizontal / vertical chroma shift of the source as
 * scaled, [3] bgr24 -> yv12 converter installed, [4..12] input_rgb2yuv_table */
int orc_sws_rgb_info(const OrcSws *s, int *out)
{
    out[0] = s->src_rgb; out[1] = s->chrSrcHSub; out[2] = s->chrSrcVSub; out[3] = s->bgr24_yv12;
    for (int i = 0; i < 9; i++) out[4 + i] = s->rgb2yuv[i];
    return 0;
}

int orc_sws_get_filter(const OrcSws *s, int which, int16_t *filter, int32_t *pos, int cap)
{
    const int16_t *f; const int32_t *p; int n, fs;
    switch (which) {
    case 0: f = s->hLum; p = s->hLumPos; n = s->dstW;    fs = s->hLumSize; break;
    case 1: f = s->hChr; p = s->hChrPos; n = s->chrDstW; fs = s->hChrSize; break;
    case 2: f = s->vLum; p = s->vLumPos; n = s->dstH;    fs = s->vLumSize; break;
    default:f = s->vChr; p = s->vChrPos; n = s->chrDstH; fs = s->vChrSize; break;
    }
    if (!f) return 0;
    if (n > cap) n = cap;
    if (filter) memcpy(filter, f, (size_t)n * fs * sizeof(int16_t));
    if (pos) memcpy(pos, p, (size_t)n * sizeof(int32_t));
    return n;
}

/* swscale.c:128-142 */
void orc_hscale8to15(int16_t *dst, int dstW, const uint8_t *src, const int16_t *filter,
                     const int32_t *filterPos, int filterSize)
{
    for (int i = 0; i < dstW; i++) {
        int val = 0;
        for (int j = 0; j < filterSize; j++)
            val += (int)src[filterPos[i] + j] * filter[filterSize * i + j];
        val >>= 7;
        dst[i] = (int16_t)(val < 32767 ? val : 32767);
    }
}

/* ff_hyscale_fast_c / ff_hcscale_fast_c (hscale_fast_bilinear.c:27-67): 16.16 stepping, 7-bit blend weights; every output
 * whose left sample is the last source sample (or beyond) is src[srcW-1] * 128 (the fix-up loop at the end of both). */
static void hscale_fast(int16_t *dst, int dstW, const uint8_t *src, int srcW, int xInc, int chroma)
{
    unsigned xpos = 0;
    for (int i = 0; i < dstW; i++, xpos += (unsigned)xInc) {
        const unsigned xx = xpos >> 16, xalpha = (xpos & 0xFFFF) >> 9;
        if (((int64_t)i * xInc) >> 16 >= srcW - 1) dst[i] = (int16_t)(src[srcW - 1] * 128);
        else if (chroma) dst[i] = (int16_t)(src[xx] * (xalpha ^ 127) + src[xx + 1] * xalpha);
        else dst[i] = (int16_t)((src[xx] << 7) + (src[xx + 1] - src[xx]) * xalpha);
    }
}
/* hScale16To15_c (swscale.c:99-125) for an RGB source: sh = 13; the reader's int16 line is read as uint16 */
static void hscale16to15(int16_t *dst, int dstW, const int16_t *src, const int16_t *filter, const int32_t *pos, int fs)
{
    for (int i = 0; i < dstW; i++) {
        int val = 0;
        for (int j = 0; j < fs; j++) val += (int)(uint16_t)src[pos[i] + j] * filter[fs * i + j];
        val >>= 13;
        dst[i] = (int16_t)(val < 32767 ? val : 32767);
    }
}

/* rgb24ToY_c / bgr24ToY_c (input.c:1068-1134) and the 32-bit templates (input.c:264-287, instantiated :390-393 with the
 * coefficients shifted by 8 and S = RGB2YUV_SHIFT + 8): the same value for all six byte orders */
static void rgb_to_y(const OrcSws *s, int16_t *dst, const uint8_t *src, int w)
{
    const int ry = s->rgb2yuv[0], gy = s->rgb2yuv[1], by = s->rgb2yuv[2];
    for (int i = 0; i < w; i++) {
        const uint8_t *p = src + (size_t)i * s->src_rgb;
        dst[i] = (int16_t)((ry * p[s->sro] + gy * p[s->sgo] + by * p[s->sbo] + (32 << 14) + (1 << 8)) >> 9);
    }
}
/* rgb24ToUV_c / rgb24ToUV_half_c and friends (input.c:1083-1172, :289-352): half = two neighbouring pixels summed */
static void rgb_to_uv(const OrcSws *s, int16_t *dU, int16_t *dV, const uint8_t *src, int w)
{
    const int ru = s->rgb2yuv[3], gu = s->rgb2yuv[4], bu = s->rgb2yuv[5], rv = s->rgb2yuv[6], gv = s->rgb2yuv[7], bv = s->rgb2yuv[8];
    for (int i = 0; i < w; i++) {
        if (s->chrSrcHSub) {
            const uint8_t *p = src + (size_t)2 * i * s->src_rgb, *q = p + s->src_rgb;
            const int r = p[s->sro] + q[s->sro], g = p[s->sgo] + q[s->sgo], b = p[s->sbo] + q[s->sbo];
            dU[i] = (int16_t)((ru * r + gu * g + bu * b + (256 << 15) + (1 << 9)) >> 10);
            dV[i] = (int16_t)((rv * r + gv * g + bv * b + (256 << 15) + (1 << 9)) >> 10);
        } else {
            const uint8_t *p = src + (size_t)i * s->src_rgb;
            const int r = p[s->sro], g = p[s->sgo], b = p[s->sbo];
            dU[i] = (int16_t)((ru * r + gu * g + bu * b + (256 << 14) + (1 << 8)) >> 9);
            dV[i] = (int16_t)((rv * r + gv * g + bv * b + (256 << 14) + (1 << 8)) >> 9);
        }
    }
}

/* ff_rgb24toyv12_c (rgb2rgb_template.c:580-641) on a bgr24 picture: truncating luma, chroma from the 2x2 box average */
static void bgr24_to_yv12(const OrcSws *s, const uint8_t *src, int ss, uint8_t *dy, int dys, uint8_t *du, int dus, uint8_t *dv, int dvs)
{
    const int *t = s->rgb2yuv;
    for (int y = 0; y < s->srcH; y += 2) {
        const uint8_t *s1 = src + (ptrdiff_t)y * ss, *s2 = y + 1 == s->srcH ? s1 : s1 + ss;
        uint8_t *y1 = dy + (ptrdiff_t)y * dys, *y2 = y + 1 == s->srcH ? y1 : y1 + dys;
        for (int i = 0; i < s->srcW >> 1; i++) {
            unsigned px[4][3];                                      /* b, g, r of the four pixels */
            for (int k = 0; k < 3; k++) { px[0][k] = s1[6 * i + k]; px[1][k] = s1[6 * i + 3 + k]; px[2][k] = s2[6 * i + k]; px[3][k] = s2[6 * i + 3 + k]; }
            uint8_t *yo[4] = { y1 + 2 * i, y1 + 2 * i + 1, y2 + 2 * i, y2 + 2 * i + 1 };
            for (int k = 0; k < 4; k++)
                *yo[k] = (uint8_t)((((unsigned)t[0] * px[k][2] + (unsigned)t[1] * px[k][1] + (unsigned)t[2] * px[k][0]) >> 15) + 16);
            const unsigned bx = (px[0][0] + px[1][0] + px[2][0] + px[3][0]) >> 2, gx = (px[0][1] + px[1][1] + px[2][1] + px[3][1]) >> 2,
                           rx = (px[0][2] + px[1][2] + px[2][2] + px[3][2]) >> 2;
            du[(ptrdiff_t)(y >> 1) * dus + i] = (uint8_t)((((unsigned)t[3] * rx + (unsigned)t[4] * gx + (unsigned)t[5] * bx) >> 15) + 128);
            dv[(ptrdiff_t)(y >> 1) * dvs + i] = (uint8_t)((((unsigned)t[6] * rx + (unsigned)t[7] * gx + (unsigned)t[8] * bx) >> 15) + 128);
        }
    }
}

/* rgbaToA_c / abgrToA_c (input.c:455-475): the alpha byte widened to 14 bits */
static void rgb_to_a(const OrcSws *s, int16_t *dst, const uint8_t *src, int w)
{
    const int sao = 6 - s->sro - s->sgo - s->sbo;
    for (int i = 0; i < w; i++) {
        const int a = src[(size_t)i * 4 + sao];
        dst[i] = (int16_t)(a << 6 | a >> 2);
    }
}

static void hpass(const OrcSws *s, int16_t *dst, int dstW, const uint8_t *src, int srcW, int chroma)
{
    if (s->fast_bilinear) hscale_fast(dst, dstW, src, srcW, chroma ? s->chrXInc : s->lumXInc, chroma);
    else if (chroma) orc_hscale8to15(dst, dstW, src, s->hChr, s->hChrPos, s->hChrSize);
    else orc_hscale8to15(dst, dstW, src, s->hLum, s->hLumPos, s->hLumSize);
}

/* the horizontal pass over a whole picture: L = srcH lines of dstW, CU / CV = chrSrcH lines of chrDstW.  For a packed RGB source
 * every line first goes through the input readers (lum_convert / chr_convert, hscale.c:103-160,229-290) */
static int hlines_a(const OrcSws *s, const uint8_t *y, int ys, const uint8_t *u, int us, const uint8_t *v, int vs,
                    int16_t *L, int16_t *CU, int16_t *CV, int16_t *A);
static int hlines(const OrcSws *s, const uint8_t *y, int ys, const uint8_t *u, int us, const uint8_t *v, int vs,
                  int16_t *L, int16_t *CU, int16_t *CV)
{
    return hlines_a(s, y, ys, u, us, v, vs, L, CU, CV, NULL);
}
/* A: the alpha lines (srcH lines of dstW) when the context carries alpha: read by alpToYV12, scaled with the luma filter (hscale.c:103-160) */
static int hlines_a(const OrcSws *s, const uint8_t *y, int ys, const uint8_t *u, int us, const uint8_t *v, int vs,
                    int16_t *L, int16_t *CU, int16_t *CV, int16_t *A)
{
    const int dstW = s->dstW, cW = s->chrDstW;
    if (s->src_rgb) {
        int16_t *t = calloc((size_t)3 * (s->srcW + 16), sizeof(int16_t));
        if (!t) return -1;
        int16_t *tY = t, *tU = t + s->srcW + 16, *tV = tU + s->srcW + 16;
        for (int r = 0; r < s->srcH; r++) {
            rgb_to_y(s, tY, y + (ptrdiff_t)r * ys, s->srcW);
            hscale16to15(L + (size_t)r * dstW, dstW, tY, s->hLum, s->hLumPos, s->hLumSize);
            if (A && s->need_alpha) {
                rgb_to_a(s, tY, y + (ptrdiff_t)r * ys, s->srcW);
                hscale16to15(A + (size_t)r * dstW, dstW, tY, s->hLum, s->hLumPos, s->hLumSize);
            }
        }
        for (int r = 0; r < s->chrSrcH; r++) {
            rgb_to_uv(s, tU, tV, y + (ptrdiff_t)r * ys, s->chrSrcW);
            hscale16to15(CU + (size_t)r * cW, cW, tU, s->hChr, s->hChrPos, s->hChrSize);
            hscale16to15(CV + (size_t)r * cW, cW, tV, s->hChr, s->hChrPos, s->hChrSize);
        }
        free(t);
        return 0;
    }
    for (int r = 0; r < s->srcH; r++)
        hpass(s, L + (size_t)r * dstW, dstW, y + (ptrdiff_t)r * ys, s->srcW, 0);
    for (int r = 0; r < s->chrSrcH; r++) {
        hpass(s, CU + (size_t)r * cW, cW, u + (ptrdiff_t)r * us, s->chrSrcW, 1);
        hpass(s, CV + (size_t)r * cW, cW, v + (ptrdiff_t)r * vs, s->chrSrcW, 1);
    }
    return 0;
}

static void range_line(int16_t *d, int w, int coeff, int offset, int clip);
/* the two stages in front of the vertical pass, exposed so that single kernels can be checked against them */
int orc_sws_hlines(const OrcSws *s, const uint8_t *y, int ys, const uint8_t *u, int us, const uint8_t *v, int vs,
                   int16_t *L, int16_t *CU, int16_t *CV)
{
    if (s->unscaled_lut) return -EINVAL;
    return hlines(s, y, ys, u, us, v, vs, L, CU, CV);
}
void orc_sws_range_lines(const OrcSws *s, int16_t *lines, int w, int n, int chroma)
{
    if (!s->range_conv) return;
    for (int r = 0; r < n; r++)
        range_line(lines + (size_t)r * w, w, chroma ? s->chr_coeff : s->lum_coeff, chroma ? s->chr_offset : s->lum_offset, s->range_conv == 1);
}

/* yuv2rgb_write, 24 and 32 bpp branches (output.c:1676-1713): one chroma pair -> LUT bases, two lumas -> two pixels */
static void put_pair(const OrcSws *s, uint8_t *d, int Y1, int Y2, int U, int V)
{
    const uint8_t *r = s->ytab + s->offR[V + HEADROOM];
    const uint8_t *g = s->ytab + s->offGU[U + HEADROOM] + s->offGV[V + HEADROOM];
    const uint8_t *b = s->ytab + s->offB[U + HEADROOM];
    uint8_t *e = d + s->bpp;
    d[s->ro] = r[Y1]; d[s->go] = g[Y1]; d[s->bo] = b[Y1];
    e[s->ro] = r[Y2]; e[s->go] = g[Y2]; e[s->bo] = b[Y2];
    if (s->ao >= 0) d[s->ao] = e[s->ao] = 255;
}

/* yuv2rgb_write_full (output.c:1998-2095) */
static void put_full(const OrcSws *s, uint8_t *d, int Y, int U, int V)
{
    unsigned y = (unsigned)(Y - s->y_offset) * (unsigned)s->y_coeff + (1U << 21);
    int R = (int)(y + (unsigned)V * (unsigned)s->v2r);
    int G = (int)(y + (unsigned)V * (unsigned)s->v2g + (unsigned)U * (unsigned)s->u2g);
    int B = (int)(y + (unsigned)U * (unsigned)s->u2b);
    if ((R | G | B) & 0xC0000000) {
#define CLIP30(a) (((a) & ~((1 << 30) - 1)) ? ((~(a)) >> 31 & ((1 << 30) - 1)) : (a))
        R = CLIP30(R); G = CLIP30(G); B = CLIP30(B);
#undef CLIP30
    }
    d[s->ro] = (uint8_t)(R >> 22); d[s->go] = (uint8_t)(G >> 22); d[s->bo] = (uint8_t)(B >> 22);
    if (s->ao >= 0) d[s->ao] = 255;
}

/* unscaled LUT converter yuv2rgb_c_24_rgb (yuv2rgb.c:137-236,530): chroma (x>>1, y>>1), no interpolation;
 * per line-pair it covers (dstW>>3)*8 + (dstW&4) + (dstW&2) pixels, an odd last column is never written. */
static void convert_unscaled(const OrcSws *s, const uint8_t *y, int ys, const uint8_t *u, int us,
                             const uint8_t *v, int vs, uint8_t *dst, int ds)
{
    int wpix = ((s->dstW >> 3) << 3) + (s->dstW & 4) + (s->dstW & 2);
    for (int row = 0; row < s->srcH; row++) {
        const uint8_t *py = y + (ptrdiff_t)row * ys;
        const uint8_t *pu = u + (ptrdiff_t)(row >> 1) * us;
        const uint8_t *pv = v + (ptrdiff_t)(row >> 1) * vs;
        uint8_t *d = dst + (ptrdiff_t)row * ds;
        for (int i = 0; i < wpix / 2; i++)
            put_pair(s, d + 2 * s->bpp * i, py[2 * i], py[2 * i + 1], pu[i], pv[i]);
    }
}

/* yuv2planeX_8_c / yuv2plane1_8_c (output.c:468-493) with the flat dither the 8-bit path uses (sws_pb_64, swscale.c:385-387):
 * one output line of one plane from `fs` int16 source lines */
static void vscale_plane_line(uint8_t *dst, int w, const int16_t *plane, int pw, int nlines, int first, const int16_t *f, int fs)
{
    for (int i = 0; i < w; i++) {
        int val;
        if (fs == 1) {
            int k = first < 0 ? 0 : first >= nlines ? nlines - 1 : first;
            val = (plane[(size_t)k * pw + i] + 64) >> 7;
        } else {
            val = 64 << 12;
            for (int j = 0; j < fs; j++) {
                int k = first + j;
                k = k < 0 ? 0 : k >= nlines ? nlines - 1 : k;
                val += plane[(size_t)k * pw + i] * f[j];
            }
            val >>= 19;
        }
        dst[i] = (uint8_t)clip_u8(val);
    }
}

/* lumRangeToJpeg_c / lumRangeFromJpeg_c / chrRange*_c (swscale.c:163-209), called per line right after the horizontal
 * pass (hscale.c:61-63, :195-197) */
static void range_line(int16_t *d, int w, int coeff, int offset, int clip)
{
    for (int i = 0; i < w; i++) {
        int v = (d[i] * coeff + offset) >> 14;
        if (clip && v > 32767) v = 32767;
        d[i] = (int16_t)v;
    }
}

/* yuv420p -> yuv420p: horizontal pass per plane, then lum_planar_vscale / chr_planar_vscale (vscale.c:34-107) */
/* nvXXtoUV_c (input.c:921-948): plane 1 of an nv12 / nv21 picture split into U and V planes of chrSrcW x chrSrcH */
static uint8_t *split_nv(const OrcSws *s, const uint8_t *uv, int uvs)
{
    uint8_t *t = malloc((size_t)2 * s->chrSrcW * s->chrSrcH);
    if (!t) return NULL;
    uint8_t *pu = t, *pv = t + (size_t)s->chrSrcW * s->chrSrcH;
    if (s->src_nv == 2) { uint8_t *x = pu; pu = pv; pv = x; }
    for (int r = 0; r < s->chrSrcH; r++)
        for (int i = 0; i < s->chrSrcW; i++) {
            pu[(size_t)r * s->chrSrcW + i] = uv[(ptrdiff_t)r * uvs + 2 * i];
            pv[(size_t)r * s->chrSrcW + i] = uv[(ptrdiff_t)r * uvs + 2 * i + 1];
        }
    return t;
}

int orc_sws_scale_planar(OrcSws *s, const uint8_t *y, int ys, const uint8_t *u, int us, const uint8_t *v, int vs,
                         uint8_t *dy, int dys, uint8_t *du, int dus, uint8_t *dv, int dvs)
{
    if (!s->planar) return -EINVAL;
    if (s->casc[0]) {                                              /* scale_cascaded (swscale.c:1001-1030): whole frames through both contexts */
        const int ts = (s->casc_w * 3 + 63) & ~63;
        /* zeroed: the unscaled converter of context 0 never writes an odd last column (yuv2rgb.c:137-236) and the reference then reads
         * whatever av_image_alloc returned; the product zeroes its intermediate picture too */
        uint8_t *t = calloc((size_t)ts * s->casc_h + 64, 1);
        if (!t) return -ENOMEM;
        int r = orc_sws_scale(s->casc[0], y, ys, u, us, v, vs, t, ts);
        if (r >= 0) r = orc_sws_scale_planar(s->casc[1], t, ts, t, ts, t, ts, dy, dys, du, dus, dv, dvs);
        free(t);
        return r;
    }
    if (s->dst_nv) {
        /* nv12 / nv21 destination: what the three-plane path writes, with the chroma planes interleaved — planarToNv12Wrapper
         * (swscale_unscaled.c:147-165) when unscaled; through the scaler yuv2nv12cX_c (output.c:495-528) forms the same sums
         * with the same flat dither as yuv2planeX_8_c / yuv2plane1_8_c and stores u, v (v, u for nv21) side by side */
        const int cw = s->chrDstW, ch = s->chrDstH, nv = s->dst_nv;
        uint8_t *t = malloc((size_t)2 * cw * ch);
        if (!t) return -ENOMEM;
        s->dst_nv = 0;
        int r = orc_sws_scale_planar(s, y, ys, u, us, v, vs, dy, dys, t, cw, t + (size_t)cw * ch, cw);
        s->dst_nv = nv;
        const uint8_t *a = nv == 1 ? t : t + (size_t)cw * ch, *b = nv == 1 ? t + (size_t)cw * ch : t;
        for (int j = 0; j < ch && r >= 0; j++)
            for (int i = 0; i < cw; i++) {
                du[(ptrdiff_t)j * dus + 2 * i] = a[(size_t)j * cw + i];
                du[(ptrdiff_t)j * dus + 2 * i + 1] = b[(size_t)j * cw + i];
            }
        free(t);
        return r;
    }
    if (s->src_nv) {
        uint8_t *t = split_nv(s, u, us);
        if (!t) return -ENOMEM;
        const int nv = s->src_nv;
        s->src_nv = 0;
        int r = orc_sws_scale_planar(s, y, ys, t, s->chrSrcW, t + (size_t)s->chrSrcW * s->chrSrcH, s->chrSrcW, dy, dys, du, dus, dv, dvs);
        s->src_nv = nv;
        free(t);
        return r;
    }
    if (s->unscaled_lut == 3) {
        bgr24_to_yv12(s, y, ys, dy, dys, du, dus, dv, dvs);
        return s->srcH;
    }
    if (s->unscaled_lut == 2) {
        for (int r = 0; r < s->srcH; r++) memcpy(dy + (ptrdiff_t)r * dys, y + (ptrdiff_t)r * ys, s->srcW);
        for (int r = 0; r < s->chrSrcH; r++) {
            memcpy(du + (ptrdiff_t)r * dus, u + (ptrdiff_t)r * us, s->chrSrcW);
            memcpy(dv + (ptrdiff_t)r * dvs, v + (ptrdiff_t)r * vs, s->chrSrcW);
        }
        return s->srcH;
    }
    const int dstW = s->dstW, cW = s->chrDstW;
    int16_t *L = malloc((size_t)s->srcH * dstW * sizeof(int16_t));
    int16_t *CU = malloc((size_t)s->chrSrcH * cW * sizeof(int16_t));
    int16_t *CV = malloc((size_t)s->chrSrcH * cW * sizeof(int16_t));
    if (!L || !CU || !CV) { free(L); free(CU); free(CV); return -ENOMEM; }
    if (hlines(s, y, ys, u, us, v, vs, L, CU, CV) < 0) { free(L); free(CU); free(CV); return -ENOMEM; }
    if (s->range_conv) {
        const int clip = s->range_conv == 1;
        for (int r = 0; r < s->srcH; r++) range_line(L + (size_t)r * dstW, dstW, s->lum_coeff, s->lum_offset, clip);
        for (int r = 0; r < s->chrSrcH; r++) {
            range_line(CU + (size_t)r * cW, cW, s->chr_coeff, s->chr_offset, clip);
            range_line(CV + (size_t)r * cW, cW, s->chr_coeff, s->chr_offset, clip);
        }
    }
    for (int d = 0; d < s->dstH; d++) {
        int first = s->vLumPos[d] > 1 - s->vLumSize ? s->vLumPos[d] : 1 - s->vLumSize;
        vscale_plane_line(dy + (ptrdiff_t)d * dys, dstW, L, dstW, s->srcH, first, s->vLum + d * s->vLumSize, s->vLumSize);
    }
    for (int d = 0; d < s->chrDstH; d++) {
        int first = s->vChrPos[d] > 1 - s->vChrSize ? s->vChrPos[d] : 1 - s->vChrSize;
        vscale_plane_line(du + (ptrdiff_t)d * dus, cW, CU, cW, s->chrSrcH, first, s->vChr + d * s->vChrSize, s->vChrSize);
        vscale_plane_line(dv + (ptrdiff_t)d * dvs, cW, CV, cW, s->chrSrcH, first, s->vChr + d * s->vChrSize, s->vChrSize);
    }
    free(L); free(CU); free(CV);
    return s->dstH;
}

int orc_sws_scale(OrcSws *s, const uint8_t *y, int ys, const uint8_t *u, int us,
                  const uint8_t *v, int vs, uint8_t *dst, int ds)
{
    if (s->planar) return -EINVAL;
    if (s->src_nv) {
        uint8_t *t = split_nv(s, u, us);
        if (!t) return -ENOMEM;
        const int nv = s->src_nv;
        s->src_nv = 0;
        int r = orc_sws_scale(s, y, ys, t, s->chrSrcW, t + (size_t)s->chrSrcW * s->chrSrcH, s->chrSrcW, dst, ds);
        s->src_nv = nv;
        free(t);
        return r;
    }
    if (s->unscaled_lut == 4) {                                    /* rgbToRgbWrapper / packedCopyWrapper: y is the packed source picture */
        const int sao = s->src_rgb == 4 ? 6 - s->sro - s->sgo - s->sbo : -1;
        for (int r = 0; r < s->srcH; r++)
            for (int x = 0; x < s->srcW; x++) {
                const uint8_t *p = y + (size_t)r * ys + (size_t)x * s->src_rgb;
                uint8_t *q = dst + (size_t)r * ds + (size_t)x * s->bpp;
                q[s->ro] = p[s->sro]; q[s->go] = p[s->sgo]; q[s->bo] = p[s->sbo];
                if (s->bpp == 4) q[s->ao] = sao >= 0 ? p[sao] : 255;
            }
        return s->srcH;
    }
    if (s->unscaled_lut) {
        convert_unscaled(s, y, ys, u, us, v, vs, dst, ds);
        return s->srcH;
    }
    const int dstW = s->dstW, cW = s->chrDstW;
    int16_t *L = malloc((size_t)s->srcH * dstW * sizeof(int16_t));
    int16_t *CU = malloc((size_t)s->chrSrcH * cW * sizeof(int16_t));
    int16_t *CV = malloc((size_t)s->chrSrcH * cW * sizeof(int16_t));
    int16_t *AL = s->need_alpha ? malloc((size_t)s->srcH * dstW * sizeof(int16_t)) : NULL;
    if (!L || !CU || !CV || (s->need_alpha && !AL)) { free(L); free(CU); free(CV); free(AL); return -ENOMEM; }
    if (hlines_a(s, y, ys, u, us, v, vs, L, CU, CV, AL) < 0) { free(L); free(CU); free(CV); free(AL); return -ENOMEM; }
    const int lfs = s->vLumSize, cfs = s->vChrSize, full = !s->chrDstHSub;
#define AROW(k) (AL + (size_t)((k) < 0 ? 0 : (k) >= s->srcH    ? s->srcH    - 1 : (k)) * dstW)
#define LROW(k) (L  + (size_t)((k) < 0 ? 0 : (k) >= s->srcH    ? s->srcH    - 1 : (k)) * dstW)
#define UROW(k) (CU + (size_t)((k) < 0 ? 0 : (k) >= s->chrSrcH ? s->chrSrcH - 1 : (k)) * cW)
#define VROW(k) (CV + (size_t)((k) < 0 ? 0 : (k) >= s->chrSrcH ? s->chrSrcH - 1 : (k)) * cW)
    for (int dy = 0; dy < s->dstH; dy++) {
        /* packed_vscale, vscale.c:109-171 */
        const int16_t *lf = s->vLum + dy * lfs, *cf = s->vChr + dy * cfs;
        int firstLum = s->vLumPos[dy] > 1 - lfs ? s->vLumPos[dy] : 1 - lfs;
        int firstChr = s->vChrPos[dy] > 1 - cfs ? s->vChrPos[dy] : 1 - cfs;
        uint8_t *d = dst + (ptrdiff_t)dy * ds;
        int mode;   /* 1: _1 writer, 2: _2 writer, 0: _X writer */
        int uvalpha = 0, yalpha = 0;
        if (lfs == 1 && cfs == 1) mode = 1;
        else if (lfs == 1 && cfs == 2 && (uint16_t)cf[0] + (uint16_t)cf[1] == 4096 && (uint16_t)cf[1] <= 4096U) {
            mode = 1; uvalpha = (uint16_t)cf[1];
        } else if (lfs == 2 && cfs == 2 &&
                   (uint16_t)lf[0] + (uint16_t)lf[1] == 4096 && (uint16_t)lf[1] <= 4096U &&
                   (uint16_t)cf[0] + (uint16_t)cf[1] == 4096 && (uint16_t)cf[1] <= 4096U) {
            mode = 2; yalpha = (uint16_t)lf[1]; uvalpha = (uint16_t)cf[1];
        } else mode = 0;

        if (!full) {
            for (int i = 0; i < (dstW + 1) >> 1; i++) {
                int Y1, Y2, U, V;
                if (mode == 1) {                                   /* output.c:1883-1939 */
                    const int16_t *b0 = LROW(firstLum);
                    Y1 = (b0[2 * i] + 64) >> 7; Y2 = (b0[2 * i + 1] + 64) >> 7;
                    if (!uvalpha) {
                        U = (UROW(firstChr)[i] + 64) >> 7; V = (VROW(firstChr)[i] + 64) >> 7;
                    } else {
                        int a1 = 4096 - uvalpha;
                        U = (UROW(firstChr)[i] * a1 + UROW(firstChr + 1)[i] * uvalpha + (128 << 11)) >> 19;
                        V = (VROW(firstChr)[i] * a1 + VROW(firstChr + 1)[i] * uvalpha + (128 << 11)) >> 19;
                    }
                } else if (mode == 2) {                            /* output.c:1843-1880 */
                    int ya1 = 4096 - yalpha, ua1 = 4096 - uvalpha;
                    const int16_t *b0 = LROW(firstLum), *b1 = LROW(firstLum + 1);
                    Y1 = (b0[2 * i] * ya1 + b1[2 * i] * yalpha) >> 19;
                    Y2 = (b0[2 * i + 1] * ya1 + b1[2 * i + 1] * yalpha) >> 19;
                    U = (UROW(firstChr)[i] * ua1 + UROW(firstChr + 1)[i] * uvalpha) >> 19;
                    V = (VROW(firstChr)[i] * ua1 + VROW(firstChr + 1)[i] * uvalpha) >> 19;
                } else {                                           /* output.c:1789-1840 */
                    unsigned a1 = 1 << 18, a2 = 1 << 18, au = 1 << 18, av = 1 << 18;
                    for (int j = 0; j < lfs; j++) {
                        a1 += (unsigned)(LROW(firstLum + j)[2 * i]     * (unsigned)(int)lf[j]);
                        a2 += (unsigned)(LROW(firstLum + j)[2 * i + 1] * (unsigned)(int)lf[j]);
                    }
                    for (int j = 0; j < cfs; j++) {
                        au += (unsigned)(UROW(firstChr + j)[i] * (unsigned)(int)cf[j]);
                        av += (unsigned)(VROW(firstChr + j)[i] * (unsigned)(int)cf[j]);
                    }
                    Y1 = (int)a1 >> 19; Y2 = (int)a2 >> 19; U = (int)au >> 19; V = (int)av >> 19;
                }
                put_pair(s, d + 2 * s->bpp * i, Y1, Y2, U, V);
                if (s->need_alpha) {                               /* the A1 / A2 of the three writers (output.c:1818-1830,1867-1872,1903-1908,1928-1933) */
                    int A1, A2;
                    if (mode == 1) {
                        const int16_t *a0 = AROW(firstLum);
                        if (uvalpha < 2048) { A1 = (a0[2 * i] * 255 + 16384) >> 15; A2 = (a0[2 * i + 1] * 255 + 16384) >> 15; }
                        else                { A1 = (a0[2 * i] + 64) >> 7;           A2 = (a0[2 * i + 1] + 64) >> 7; }
                        A1 = clip_u8(A1); A2 = clip_u8(A2);
                    } else if (mode == 2) {
                        const int16_t *a0 = AROW(firstLum), *a1 = AROW(firstLum + 1);
                        A1 = clip_u8((a0[2 * i] * (4096 - yalpha) + a1[2 * i] * yalpha) >> 19);
                        A2 = clip_u8((a0[2 * i + 1] * (4096 - yalpha) + a1[2 * i + 1] * yalpha) >> 19);
                    } else {
                        unsigned s1 = 1 << 18, s2 = 1 << 18;
                        for (int j = 0; j < lfs; j++) {
                            s1 += (unsigned)(AROW(firstLum + j)[2 * i]     * (unsigned)(int)lf[j]);
                            s2 += (unsigned)(AROW(firstLum + j)[2 * i + 1] * (unsigned)(int)lf[j]);
                        }
                        A1 = (int)s1 >> 19; A2 = (int)s2 >> 19;
                        if ((A1 | A2) & 0x100) { A1 = clip_u8(A1); A2 = clip_u8(A2); }
                    }
                    d[2 * s->bpp * i + s->ao] = (uint8_t)A1;
                    if (2 * i + 1 < dstW) d[2 * s->bpp * i + s->bpp + s->ao] = (uint8_t)A2;
                }
            }
        } else {
            for (int i = 0; i < dstW; i++) {
                int Y, U, V;
                if (mode == 1) {                                   /* output.c:2257-2310 */
                    Y = LROW(firstLum)[i] * 4;
                    if (!uvalpha) {
                        U = (UROW(firstChr)[i] - (128 << 7)) * 4; V = (VROW(firstChr)[i] - (128 << 7)) * 4;
                    } else {
                        int a1 = 4096 - uvalpha;
                        U = (UROW(firstChr)[i] * a1 + UROW(firstChr + 1)[i] * uvalpha - (128 << 19)) >> 10;
                        V = (VROW(firstChr)[i] * a1 + VROW(firstChr + 1)[i] * uvalpha - (128 << 19)) >> 10;
                    }
                } else if (mode == 2) {                            /* output.c:2211-2254 */
                    int ya1 = 4096 - yalpha, ua1 = 4096 - uvalpha;
                    Y = (LROW(firstLum)[i] * ya1 + LROW(firstLum + 1)[i] * yalpha) >> 10;
                    U = (UROW(firstChr)[i] * ua1 + UROW(firstChr + 1)[i] * uvalpha - (128 << 19)) >> 10;
                    V = (VROW(firstChr)[i] * ua1 + VROW(firstChr + 1)[i] * uvalpha - (128 << 19)) >> 10;
                } else {                                           /* output.c:2161-2208 */
                    unsigned ay = 1 << 9, au = (1 << 9) - (128 << 19), av = (1 << 9) - (128 << 19);
                    for (int j = 0; j < lfs; j++) ay += (unsigned)(LROW(firstLum + j)[i] * (unsigned)(int)lf[j]);
                    for (int j = 0; j < cfs; j++) {
                        au += (unsigned)(UROW(firstChr + j)[i] * (unsigned)(int)cf[j]);
                        av += (unsigned)(VROW(firstChr + j)[i] * (unsigned)(int)cf[j]);
                    }
                    Y = (int)ay >> 10; U = (int)au >> 10; V = (int)av >> 10;
                }
                put_full(s, d + s->bpp * i, Y, U, V);
                if (s->need_alpha) {                               /* output.c:2191-2199,2240-2244,2282-2286 */
                    int A;
                    if (mode == 1) A = (AROW(firstLum)[i] + 64) >> 7;
                    else if (mode == 2) A = (AROW(firstLum)[i] * (4096 - yalpha) + AROW(firstLum + 1)[i] * yalpha + (1 << 18)) >> 19;
                    else {
                        unsigned sa = 1 << 18;
                        for (int j = 0; j < lfs; j++) sa += (unsigned)(AROW(firstLum + j)[i] * (unsigned)(int)lf[j]);
                        A = (int)sa >> 19;
                    }
                    if (A & 0x100) A = clip_u8(A);
                    d[s->bpp * i + s->ao] = (uint8_t)A;
                }
            }
        }
    }
#undef LROW
#undef UROW
#undef VROW
#undef AROW
    free(L); free(CU); free(CV); free(AL);
    return s->dstH;
}
