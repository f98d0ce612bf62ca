/*
 * txd_oracle.c — TEST INFRASTRUCTURE ONLY.  CPU restatement of libavutil/tx's DOUBLE precision power-of-two FFT and MDCT (AV_TX_DOUBLE_FFT,
 * AV_TX_DOUBLE_MDCT: libavutil/tx_double.c instantiates tx_template.c with TXSample = double), i.e. the *_double_c codelets av_tx_init()
 * resolves to in a build without assembly.  Same structure and operation order as tx_oracle.c (float); the references there apply:
 *   tables tx_template.c:65-77, butterflies :540-560, split-radix combine :562-586, base cases :631-704, recursion :615-627,
 *   permutation tx.c:125-154, MDCT :1223-1342, twiddles :2107-2134 (scale is a const double * for the double types).
 * Compiled with -ffp-contract=off.  Pinned on the compiled reference (tests/test_oracle_more.py) and on fixtures generated from it.
 */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double re, im; } cpxd;

struct OrcTxD {
    int type, inv, len, n;      /* type 2 FFT, 3 MDCT (AVTXType values); n: FFT size actually run */
    int *map, *sub_map;
    cpxd *exp;
    double *tab[18];
};

static void make_tab(OrcTxD *t, int k)
{
    const int n = 1 << k;
    if (t->tab[k]) return;
    t->tab[k] = malloc(sizeof(double) * (n / 4 + 1));
    const double freq = 2 * M_PI / n;
    for (int i = 0; i < n / 4; i++) t->tab[k][i] = cos(i * freq);
    t->tab[k][n / 4] = 0;
}

/* inputs t1,t2 (from a2) and t5,t6 (from a3); tx_template.c:540-552 */
static inline void butterflies(cpxd *a0, cpxd *a1, cpxd *a2, cpxd *a3, double t1, double t2, double t5, double t6)
{
    const double r0 = a0->re, i0 = a0->im, r1 = a1->re, i1 = a1->im;
    const double t3 = t5 - t1; t5 = t5 + t1;
    a2->re = r0 - t5; a0->re = r0 + t5;
    a3->im = i1 - t3; a1->im = i1 + t3;
    const double t4 = t2 - t6; t6 = t2 + t6;
    a3->re = r1 - t4; a1->re = r1 + t4;
    a2->im = i0 - t6; a0->im = i0 + t6;
}

/* tx_template.c:554-560 */
static inline void transform(cpxd *a0, cpxd *a1, cpxd *a2, cpxd *a3, double wre, double wim)
{
    const double t1 = a2->re * wre - a2->im * (-wim);
    const double t2 = a2->re * (-wim) + a2->im * wre;
    const double t5 = a3->re * wre - a3->im * wim;
    const double t6 = a3->re * wim + a3->im * wre;
    butterflies(a0, a1, a2, a3, t1, t2, t5, t6);
}

static void fft4(cpxd *d, const cpxd *s)
{
    double t1, t2, t3, t4, t5, t6, t7, t8;
    t3 = s[0].re - s[1].re; t1 = s[0].re + s[1].re;
    t8 = s[3].re - s[2].re; t6 = s[3].re + s[2].re;
    const double d2re = t1 - t6, d0re = t1 + t6;
    t4 = s[0].im - s[1].im; t2 = s[0].im + s[1].im;
    t7 = s[2].im - s[3].im; t5 = s[2].im + s[3].im;
    d[2].re = d2re; d[0].re = d0re;
    d[3].im = t4 - t8; d[1].im = t4 + t8;
    d[3].re = t3 - t7; d[1].re = t3 + t7;
    d[2].im = t2 - t5; d[0].im = t2 + t5;
}

static void fft_ns(OrcTxD *t, int k, cpxd *d, const cpxd *s)
{
    const int n = 1 << k;
    if (n == 2) {
        const double re = s[0].re - s[1].re, im = s[0].im - s[1].im;
        d[0].re = s[0].re + s[1].re; d[0].im = s[0].im + s[1].im;
        d[1].re = re; d[1].im = im;
    } else if (n == 4) {
        fft4(d, s);
    } else if (n == 8) {
        const double c = t->tab[3][1];
        /* the reference runs fft4 first and then reads src[4..7]; src may alias dst, rows 4..7 are untouched by fft4 */
        fft4(d, s);
        const double t1 = s[4].re - (-s[5].re), d5re = s[4].re + (-s[5].re);
        const double t2 = s[4].im - (-s[5].im), d5im = s[4].im + (-s[5].im);
        const double t5 = s[6].re - (-s[7].re), d7re = s[6].re + (-s[7].re);
        const double t6 = s[6].im - (-s[7].im), d7im = s[6].im + (-s[7].im);
        d[5].re = d5re; d[5].im = d5im; d[7].re = d7re; d[7].im = d7im;
        butterflies(&d[0], &d[2], &d[4], &d[6], t1, t2, t5, t6);
        transform(&d[1], &d[3], &d[5], &d[7], c, c);
    } else if (n == 16) {
        const double *c = t->tab[4];
        fft_ns(t, 3, d, s);
        fft4(d + 8, s + 8);
        fft4(d + 12, s + 12);
        butterflies(&d[0], &d[4], &d[8], &d[12], d[8].re, d[8].im, d[12].re, d[12].im);
        transform(&d[2], &d[6], &d[10], &d[14], c[2], c[2]);
        transform(&d[1], &d[5], &d[9], &d[13], c[1], c[3]);
        transform(&d[3], &d[7], &d[11], &d[15], c[3], c[1]);
    } else {
        const int n4 = n / 4;
        fft_ns(t, k - 1, d, s);
        fft_ns(t, k - 2, d + 2 * n4, s + 2 * n4);
        fft_ns(t, k - 2, d + 3 * n4, s + 3 * n4);
        /* ff_tx_fft_sr_combine(d, tab_n, n/8) */
        const int len = n4 >> 1, o1 = 2 * len, o2 = 4 * len, o3 = 6 * len;
        const double *cs = t->tab[k], *wim = cs + o1 - 7;
        cpxd *z = d;
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


static void run_mdct_inv(OrcTxD *t, double *out, const double *in, ptrdiff_t stride)
{
    cpxd *z = (cpxd *)out;
    const cpxd *e = t->exp;
    const int len2 = t->len >> 1, len4 = t->len >> 2;
    const double *in1 = in, *in2 = in + (len2 * 2 - 1) * stride;
    for (int i = 0; i < len2; i++) {
        const int k = t->sub_map[i];
        const double are = in2[-k * stride], aim = in1[k * stride];
        z[i].re = are * e[i].re - aim * e[i].im;
        z[i].im = are * e[i].im + aim * e[i].re;
    }
    fft_ns(t, ilog2i(len2), z, z);
    e += len2;
    for (int i = 0; i < len4; i++) {
        const int i0 = len4 + i, i1 = len4 - i - 1;
        const cpxd s1 = { z[i1].im, z[i1].re }, s0 = { z[i0].im, z[i0].re };
        z[i1].re = s1.re * e[i1].im - s1.im * e[i1].re;
        z[i0].im = s1.re * e[i1].re + s1.im * e[i1].im;
        z[i0].re = s0.re * e[i0].im - s0.im * e[i0].re;
        z[i1].im = s0.re * e[i0].re + s0.im * e[i0].im;
    }
}

static void run_mdct_fwd(OrcTxD *t, double *dst, const double *src, ptrdiff_t stride)
{
    cpxd *z = (cpxd *)dst;
    const cpxd *e = t->exp;
    const int len2 = t->len >> 1, len4 = t->len >> 2, len3 = len2 * 3;
    for (int i = 0; i < len2; i++) {
        const int k = 2 * i, idx = t->sub_map[i];
        double re, im;
        if (k < len2) {
            re = -src[len2 + k] + src[1 * len2 - 1 - k];
            im = -src[len3 + k] + -src[1 * len3 - 1 - k];
        } else {
            re = -src[len2 + k] + -src[5 * len2 - 1 - k];
            im = src[-len2 + k] + -src[1 * len3 - 1 - k];
        }
        z[idx].im = re * e[i].re - im * e[i].im;
        z[idx].re = re * e[i].im + im * e[i].re;
    }
    fft_ns(t, ilog2i(len2), z, z);
    for (int i = 0; i < len4; i++) {
        const int i0 = len4 + i, i1 = len4 - i - 1;
        const cpxd s1 = z[i1], s0 = z[i0];
        dst[2 * i1 * stride + stride] = s0.re * e[i0].im - s0.im * e[i0].re;
        dst[2 * i0 * stride]          = s0.re * e[i0].re + s0.im * e[i0].im;
        dst[2 * i0 * stride + stride] = s1.re * e[i1].im - s1.im * e[i1].re;
        dst[2 * i1 * stride]          = s1.re * e[i1].re + s1.im * e[i1].im;
    }
}


OrcTxD *orc_txd_open(int type, int inv, int len, double scale, unsigned flags)
{
    if (flags || (type != 2 && type != 3) || len < 2 || (len & (len - 1))) return NULL;
    if (type == 3 && len < 4) return NULL;
    OrcTxD *t = calloc(1, sizeof(*t));
    t->type = type; t->inv = !!inv; t->len = len;
    t->n = type == 2 ? len : len >> 1;
    if (t->n < 1 || t->n > 131072) { free(t); return NULL; }
    const int k = ilog2i(t->n);
    for (int j = 3; j <= k; j++) make_tab(t, j);
    t->map = malloc(sizeof(int) * t->n);
    const int scatter = type == 3 && !inv;                     /* ff_tx_mdct_init: map_dir = !inv ? SCATTER : GATHER */
    for (int i = 0; i < t->n; i++) {
        const int p = t->n == 1 ? 0 : (-sr_perm(i, t->n, t->inv)) & (t->n - 1);
        if (scatter) t->map[p] = i; else t->map[i] = p;
    }
    if (type == 3) {
        const int len4 = t->len >> 1;
        const double theta = (scale < 0 ? len4 : 0) + 1.0 / 8.0, sc = sqrt(fabs(scale));
        cpxd *full = malloc(sizeof(cpxd) * len4);
        for (int i = 0; i < len4; i++) {
            const double alpha = M_PI_2 * (i + theta) / len4;
            full[i].re = cos(alpha) * sc;
            full[i].im = sin(alpha) * sc;
        }
        if (inv) {
            t->exp = malloc(sizeof(cpxd) * 2 * len4);
            memcpy(t->exp + len4, full, sizeof(cpxd) * len4);
            for (int i = 0; i < len4; i++) t->exp[i] = full[t->map[i]];
            free(full);
        } else
            t->exp = full;
        t->sub_map = malloc(sizeof(int) * len4);
        for (int i = 0; i < len4; i++) t->sub_map[i] = inv ? t->map[i] << 1 : t->map[i];
    }
    return t;
}

void orc_txd_close(OrcTxD *t)
{
    if (!t) return;
    for (int i = 0; i < 18; i++) free(t->tab[i]);
    free(t->map); free(t->sub_map); free(t->exp); free(t);
}

/* count transforms; in / out advance by in_step / out_step BYTES; stride is the av_tx_fn stride argument (bytes) */
void orc_txd_run(OrcTxD *t, void *out, void *in, ptrdiff_t stride, int count, ptrdiff_t out_step, ptrdiff_t in_step)
{
    for (int c = 0; c < count; c++) {
        void *o = (uint8_t *)out + c * out_step, *i = (uint8_t *)in + c * in_step;
        if (t->type == 2) {
            cpxd *d = o; const cpxd *s = i;
            for (int k = 0; k < t->n; k++) d[k] = s[t->map[k]];
            fft_ns(t, ilog2i(t->n), d, d);
        } else if (t->inv) run_mdct_inv(t, o, i, stride / (ptrdiff_t)sizeof(double));
        else run_mdct_fwd(t, o, i, stride / (ptrdiff_t)sizeof(double));
    }
}
