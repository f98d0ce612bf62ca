/*
 * txi_oracle.c — TEST INFRASTRUCTURE ONLY.  CPU restatement of libavutil/tx's 32-bit fixed-point transforms (AV_TX_INT32_FFT,
 * AV_TX_INT32_MDCT, power-of-two lengths): the same split-radix structure as the float transforms (tx_oracle.c) with the
 * arithmetic of the TX_INT32 instantiation (libavutil/tx_int32.c, macros libavutil/tx_priv.h:113-155):
 *   sums wrap modulo 2^32 (BF on unsigned), products are 64-bit with rounding: (a*b - c*d + 2^30) >> 31 (CMUL),
 *   tables are RESCALE(x) = clip(llrintf((float)(x * 2^31))) (note the float rounding before the integer one),
 *   the forward MDCT folds its input with FOLD(a, b) = (a + b + 32) >> 6.
 * Follows (behaviour, not text) libavutil/tx_template.c: tables :65-77, butterflies / transform / combine :540-586, base cases
 * :631-704, recursion :615-629, FFT wrapper :763-778, MDCT :1223-1342, twiddles :2107-2134; permutation libavutil/tx.c:125-154.
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>

typedef struct { int32_t re, im; } cpxi;

struct OrcTxI {
    int type, inv, len, n;      /* type 4 FFT, 5 MDCT (AVTXType values) */
    int *map, *sub_map;
    cpxi *exp;
    int32_t *tab[18];
};

static int32_t rescale(double x)
{
    const float f = (float)(x * 2147483648.0);
    long long v = llrintf(f);
    if (v < INT32_MIN) v = INT32_MIN;
    if (v > INT32_MAX) v = INT32_MAX;
    return (int32_t)v;
}
static inline int32_t mulr(int64_t accu) { return (int32_t)((accu + 0x40000000) >> 31); }
/* CMUL(dre, dim, are, aim, bre, bim) */
static inline void cmul(uint32_t *dre, uint32_t *dim, int32_t are, int32_t aim, int32_t bre, int32_t bim)
{
    *dre = (uint32_t)mulr((int64_t)bre * are - (int64_t)bim * aim);
    *dim = (uint32_t)mulr((int64_t)bim * are + (int64_t)bre * aim);
}

static void make_tab(OrcTxI *t, int k)
{
    const int n = 1 << k;
    if (t->tab[k]) return;
    t->tab[k] = malloc(sizeof(int32_t) * (n / 4 + 1));
    const double freq = 2 * M_PI / n;
    for (int i = 0; i < n / 4; i++) t->tab[k][i] = rescale(cos(i * freq));
    t->tab[k][n / 4] = 0;
}

static inline void butterflies(cpxi *a0, cpxi *a1, cpxi *a2, cpxi *a3, uint32_t t1, uint32_t t2, uint32_t t5, uint32_t t6)
{
    const uint32_t r0 = a0->re, i0 = a0->im, r1 = a1->re, i1 = a1->im;
    const uint32_t t3 = t5 - t1; t5 = t5 + t1;
    a2->re = (int32_t)(r0 - t5); a0->re = (int32_t)(r0 + t5);
    a3->im = (int32_t)(i1 - t3); a1->im = (int32_t)(i1 + t3);
    const uint32_t t4 = t2 - t6; t6 = t2 + t6;
    a3->re = (int32_t)(r1 - t4); a1->re = (int32_t)(r1 + t4);
    a2->im = (int32_t)(i0 - t6); a0->im = (int32_t)(i0 + t6);
}
static inline void transform(cpxi *a0, cpxi *a1, cpxi *a2, cpxi *a3, int32_t wre, int32_t wim)
{
    uint32_t t1, t2, t5, t6;
    cmul(&t1, &t2, a2->re, a2->im, wre, -wim);
    cmul(&t5, &t6, a3->re, a3->im, wre, wim);
    butterflies(a0, a1, a2, a3, t1, t2, t5, t6);
}
static void fft4(cpxi *d, const cpxi *s)
{
    const uint32_t s0r = s[0].re, s1r = s[1].re, s2r = s[2].re, s3r = s[3].re, s0i = s[0].im, s1i = s[1].im, s2i = s[2].im, s3i = s[3].im;
    const uint32_t t3 = s0r - s1r, t1 = s0r + s1r, t8 = s3r - s2r, t6 = s3r + s2r;
    const uint32_t t4 = s0i - s1i, t2 = s0i + s1i, t7 = s2i - s3i, t5 = s2i + s3i;
    d[2].re = (int32_t)(t1 - t6); d[0].re = (int32_t)(t1 + t6);
    d[3].im = (int32_t)(t4 - t8); d[1].im = (int32_t)(t4 + t8);
    d[3].re = (int32_t)(t3 - t7); d[1].re = (int32_t)(t3 + t7);
    d[2].im = (int32_t)(t2 - t5); d[0].im = (int32_t)(t2 + t5);
}
static void fft_ns(OrcTxI *t, int k, cpxi *d, const cpxi *s)
{
    const int n = 1 << k;
    if (n == 2) {
        const uint32_t re = (uint32_t)s[0].re - (uint32_t)s[1].re, im = (uint32_t)s[0].im - (uint32_t)s[1].im;
        d[0].re = (int32_t)((uint32_t)s[0].re + (uint32_t)s[1].re); d[0].im = (int32_t)((uint32_t)s[0].im + (uint32_t)s[1].im);
        d[1].re = (int32_t)re; d[1].im = (int32_t)im;
    } else if (n == 4) {
        fft4(d, s);
    } else if (n == 8) {
        const int32_t c = t->tab[3][1];
        const cpxi s4 = s[4], s5 = s[5], s6 = s[6], s7 = s[7];
        fft4(d, s);
        const uint32_t t1 = (uint32_t)s4.re - (uint32_t)(-s5.re), t2 = (uint32_t)s4.im - (uint32_t)(-s5.im);
        const uint32_t t5 = (uint32_t)s6.re - (uint32_t)(-s7.re), t6 = (uint32_t)s6.im - (uint32_t)(-s7.im);
        d[5].re = (int32_t)((uint32_t)s4.re + (uint32_t)(-s5.re)); d[5].im = (int32_t)((uint32_t)s4.im + (uint32_t)(-s5.im));
        d[7].re = (int32_t)((uint32_t)s6.re + (uint32_t)(-s7.re)); d[7].im = (int32_t)((uint32_t)s6.im + (uint32_t)(-s7.im));
        butterflies(&d[0], &d[2], &d[4], &d[6], t1, t2, t5, t6);
        transform(&d[1], &d[3], &d[5], &d[7], c, c);
    } else if (n == 16) {
        const int32_t *c = t->tab[4];
        fft_ns(t, 3, d, s);
        fft4(d + 8, s + 8);
        fft4(d + 12, s + 12);
        butterflies(&d[0], &d[4], &d[8], &d[12], (uint32_t)d[8].re, (uint32_t)d[8].im, (uint32_t)d[12].re, (uint32_t)d[12].im);
        transform(&d[2], &d[6], &d[10], &d[14], c[2], c[2]);
        transform(&d[1], &d[5], &d[9], &d[13], c[1], c[3]);
        transform(&d[3], &d[7], &d[11], &d[15], c[3], c[1]);
    } else {
        const int n4 = n / 4;
        fft_ns(t, k - 1, d, s);
        fft_ns(t, k - 2, d + 2 * n4, s + 2 * n4);
        fft_ns(t, k - 2, d + 3 * n4, s + 3 * n4);
        const int len = n4 >> 1, o1 = 2 * len, o2 = 4 * len, o3 = 6 * len;
        const int32_t *cs = t->tab[k], *wim = cs + o1 - 7;
        cpxi *z = d;
        for (int i = 0; i < len; i += 4) {
            static const int order[8] = { 0, 2, 4, 6, 1, 3, 5, 7 };
            for (int q = 0; q < 8; q++) {
                const int j = order[q];
                transform(&z[j], &z[o1 + j], &z[o2 + j], &z[o3 + j], cs[j], wim[7 - j]);
            }
            z += 8; cs += 8; wim -= 8;
        }
    }
}

static int sr_perm(int i, int len, int inv)
{
    len >>= 1;
    if (len <= 1) return i & 1;
    if (!(i & len)) return sr_perm(i, len, inv) * 2;
    len >>= 1;
    return sr_perm(i, len, inv) * 4 + 1 - 2 * (!(i & len) ^ inv);
}
static int ilog2i(int n) { int k = 0; while ((1 << k) < n) k++; return k; }

OrcTxI *orc_txi_open(int type, int inv, int len, float scale, unsigned flags)
{
    if (flags || (type != 4 && type != 5) || len < 2 || (len & (len - 1))) return NULL;
    OrcTxI *t = calloc(1, sizeof(*t));
    t->type = type; t->inv = !!inv; t->len = len;
    t->n = type == 4 ? len : len >> 1;
    if (t->n < 1 || t->n > 131072) { free(t); return NULL; }
    const int k = ilog2i(t->n);
    for (int j = 3; j <= k; j++) make_tab(t, j);
    t->map = malloc(sizeof(int) * t->n);
    const int scatter = type == 5 && !inv;
    for (int i = 0; i < t->n; i++) {
        const int p = t->n == 1 ? 0 : (-sr_perm(i, t->n, t->inv)) & (t->n - 1);
        if (scatter) t->map[p] = i; else t->map[i] = p;
    }
    if (type == 5) {
        const int len4 = t->len >> 1;
        const double theta = (scale < 0 ? len4 : 0) + 1.0 / 8.0, sc = sqrt(fabs((double)scale));
        cpxi *full = malloc(sizeof(cpxi) * len4);
        for (int i = 0; i < len4; i++) {
            const double alpha = M_PI_2 * (i + theta) / len4;
            full[i].re = rescale(cos(alpha) * sc);
            full[i].im = rescale(sin(alpha) * sc);
        }
        if (inv) {
            t->exp = malloc(sizeof(cpxi) * 2 * len4);
            memcpy(t->exp + len4, full, sizeof(cpxi) * len4);
            for (int i = 0; i < len4; i++) t->exp[i] = full[t->map[i]];
            free(full);
        } else
            t->exp = full;
        t->sub_map = malloc(sizeof(int) * len4);
        for (int i = 0; i < len4; i++) t->sub_map[i] = inv ? t->map[i] << 1 : t->map[i];
    }
    return t;
}

void orc_txi_close(OrcTxI *t)
{
    if (!t) return;
    for (int i = 0; i < 18; i++) free(t->tab[i]);
    free(t->map); free(t->sub_map); free(t->exp); free(t);
}

static inline int32_t fold(int32_t x, int32_t y) { return (int32_t)((uint32_t)x + (uint32_t)y + 32) >> 6; }

static void run_mdct_inv(OrcTxI *t, int32_t *out, const int32_t *in, ptrdiff_t stride)
{
    cpxi *z = (cpxi *)out;
    const cpxi *e = t->exp;
    const int len2 = t->len >> 1, len4 = t->len >> 2;
    const int32_t *in1 = in, *in2 = in + (len2 * 2 - 1) * stride;
    for (int i = 0; i < len2; i++) {
        const int k = t->sub_map[i];
        uint32_t re, im;
        cmul(&re, &im, in2[-k * stride], in1[k * stride], e[i].re, e[i].im);
        z[i].re = (int32_t)re; z[i].im = (int32_t)im;
    }
    fft_ns(t, ilog2i(len2), z, z);
    e += len2;
    for (int i = 0; i < len4; i++) {
        const int i0 = len4 + i, i1 = len4 - i - 1;
        const cpxi s1 = { z[i1].im, z[i1].re }, s0 = { z[i0].im, z[i0].re };
        uint32_t a, b;
        cmul(&a, &b, s1.re, s1.im, e[i1].im, e[i1].re); z[i1].re = (int32_t)a; z[i0].im = (int32_t)b;
        cmul(&a, &b, s0.re, s0.im, e[i0].im, e[i0].re); z[i0].re = (int32_t)a; z[i1].im = (int32_t)b;
    }
}

static void run_mdct_fwd(OrcTxI *t, int32_t *dst, const int32_t *src, ptrdiff_t stride)
{
    cpxi *z = (cpxi *)dst;
    const cpxi *e = t->exp;
    const int len2 = t->len >> 1, len4 = t->len >> 2, len3 = len2 * 3;
    for (int i = 0; i < len2; i++) {
        const int k = 2 * i, idx = t->sub_map[i];
        int32_t re, im;
        if (k < len2) { re = fold(-src[len2 + k], src[1 * len2 - 1 - k]); im = fold(-src[len3 + k], -src[1 * len3 - 1 - k]); }
        else          { re = fold(-src[len2 + k], -src[5 * len2 - 1 - k]); im = fold(src[-len2 + k], -src[1 * len3 - 1 - k]); }
        uint32_t a, b;
        cmul(&a, &b, re, im, e[i].re, e[i].im);
        z[idx].im = (int32_t)a; z[idx].re = (int32_t)b;
    }
    fft_ns(t, ilog2i(len2), z, z);
    for (int i = 0; i < len4; i++) {
        const int i0 = len4 + i, i1 = len4 - i - 1;
        const cpxi s1 = z[i1], s0 = z[i0];
        uint32_t a, b;
        cmul(&a, &b, s0.re, s0.im, e[i0].im, e[i0].re); dst[2 * i1 * stride + stride] = (int32_t)a; dst[2 * i0 * stride] = (int32_t)b;
        cmul(&a, &b, s1.re, s1.im, e[i1].im, e[i1].re); dst[2 * i0 * stride + stride] = (int32_t)a; dst[2 * i1 * stride] = (int32_t)b;
    }
}

void orc_txi_run(OrcTxI *t, void *out, void *in, ptrdiff_t stride, int count, ptrdiff_t out_step, ptrdiff_t in_step)
{
    for (int c = 0; c < count; c++) {
        void *o = (uint8_t *)out + c * out_step, *i = (uint8_t *)in + c * in_step;
        if (t->type == 4) {
            cpxi *d = o; const cpxi *s = i;
            for (int j = 0; j < t->n; j++) d[j] = s[t->map[j]];
            fft_ns(t, ilog2i(t->n), d, d);
        } else if (t->inv) run_mdct_inv(t, o, i, stride / (ptrdiff_t)sizeof(int32_t));
        else run_mdct_fwd(t, o, i, stride / (ptrdiff_t)sizeof(int32_t));
    }
}
