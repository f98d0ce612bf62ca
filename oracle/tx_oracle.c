/*
 * tx_oracle.c — TEST INFRASTRUCTURE ONLY.  CPU restatement of libavutil/tx's float32 power-of-two FFT and MDCT
 * (the *_float_c codelets, i.e. what av_tx_init() resolves to in a build without x86 assembly).
 *
 * Follows (behaviour and FLOAT OPERATION ORDER, not text):
 *   tables ............... libavutil/tx_template.c:65-77      tab_N[i] = (float)cos(2*pi*i/N), i < N/4, then 0
 *   butterfly / transform  libavutil/tx_template.c:540-560    BUTTERFLIES, TRANSFORM; CMUL/BF libavutil/tx_priv.h:89-110
 *   split-radix combine .. libavutil/tx_template.c:562-586    ff_tx_fft_sr_combine
 *   base cases ........... libavutil/tx_template.c:631-704    fft2/4/8/16
 *   recursion ............ libavutil/tx_template.c:615-627    fftN = fft(N/2) + 2 x fft(N/4) + combine(N/8)
 *   permutation .......... libavutil/tx.c:125-154             split_radix_permutation / ff_tx_gen_ptwo_revtab (gather / scatter)
 *   out-of-place wrapper . libavutil/tx_template.c:763-778    ff_tx_fft: dst[i] = src[map[i]], then the in-place codelet
 *   MDCT ................. libavutil/tx_template.c:1223-1342  ff_tx_mdct_init / _fwd / _inv (half iMDCT: len floats out)
 *   twiddles ............. libavutil/tx_template.c:2107-2134  ff_tx_mdct_gen_exp (double math, cast to float)
 * Compiled with -ffp-contract=off: every product and sum below is rounded separately, like the reference's C build.
 * Parity of this file with the compiled reference is bit-exact on all test vectors (tests/test_oracle.py); there is
 * no golden bit-level vector for float tx in the reference tree (its checkasm uses eps 5e-4), so the bit-level pin is
 * the compiled reference itself (oracle/_ref) and fixtures generated from it.
 */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float re, im; } cpx;

struct OrcTx {
    int type, inv, len;         /* type 0 FFT, 1 MDCT, 6 RDFT (AVTXType values); len: FFT points, or MDCT / RDFT len (len/2-point FFT inside) */
    int n;                      /* FFT size actually run */
    int *map;                   /* FFT permutation (gather for FFT / inverse MDCT, scatter for forward MDCT) */
    int *sub_map;               /* MDCT: map, doubled for the inverse */
    cpx *exp;                   /* MDCT twiddles */
    float *rexp;                /* RDFT: 8 factors, then cos and sin-like tables of len/4 entries each */
    float *tab[18];             /* tab[k] = cosine table of size 2^k */
    /* compound 15 x M MDCT (ff_tx_mdct_pfa_15xM_{inv,fwd}, tx_template.c:1471-1599): len/2 = 15 * m complex points */
    int full;                   /* AV_TX_FULL_IMDCT: ff_tx_mdct_inv_full around the inverse MDCT (tx_template.c:1372-1413) */
    int pfa_m;                  /* 0: power-of-two transform; else m (power of two >= 2) */
    int pfa_n;                  /* the odd factor: 15, 9, 7, 5 or 3 (ff_tx_mdct_pfa_{15,9,7,5,3}xM) */
    int *pfa_in, *pfa_out;      /* compound input map (15-point groups, 3x5 map embedded; doubled for the inverse) and CRT output map */
    int *pfa_sub;               /* scatter permutation of the m-point FFT (ff_tx_gen_ptwo_revtab, FF_TX_MAP_SCATTER) */
    cpx *pfa_tmp;
    float tab53[12 + 6 + 8];    /* ff_tx_tab_53 (tx_template.c:91-108), then ff_tx_tab_7 (:110-118) and ff_tx_tab_9 (:120-130) */
    /* AV_TX_FLOAT_DCT (type 9): DCT-II forward / DCT-III inverse around a real DFT (ff_tx_dctII / ff_tx_dctIII, tx_template.c:1832-1968) */
    OrcTx *dct_sub;             /* the r2c (forward) or c2r (inverse) transform of dct_n points */
    int dct_n;                  /* points: len forward, 2 * len inverse (ff_tx_dct_init doubles it) */
    float *dct_exp;             /* dct_n rotation factors, then dct_n / 2 pre- or post-scaling factors */
};

static void make_tab(OrcTx *t, int k)
{
    const int n = 1 << k;
    if (t->tab[k]) return;
    t->tab[k] = malloc(sizeof(float) * (n / 4 + 1));
    const double freq = 2 * M_PI / n;
    for (int i = 0; i < n / 4; i++) t->tab[k][i] = (float)cos(i * freq);
    t->tab[k][n / 4] = 0;
}

/* inputs t1,t2 (from a2) and t5,t6 (from a3); tx_template.c:540-552 */
static inline void butterflies(cpx *a0, cpx *a1, cpx *a2, cpx *a3, float t1, float t2, float t5, float t6)
{
    const float r0 = a0->re, i0 = a0->im, r1 = a1->re, i1 = a1->im;
    const float t3 = t5 - t1; t5 = t5 + t1;
    a2->re = r0 - t5; a0->re = r0 + t5;
    a3->im = i1 - t3; a1->im = i1 + t3;
    const float t4 = t2 - t6; t6 = t2 + t6;
    a3->re = r1 - t4; a1->re = r1 + t4;
    a2->im = i0 - t6; a0->im = i0 + t6;
}

/* tx_template.c:554-560 */
static inline void transform(cpx *a0, cpx *a1, cpx *a2, cpx *a3, float wre, float wim)
{
    const float t1 = a2->re * wre - a2->im * (-wim);
    const float t2 = a2->re * (-wim) + a2->im * wre;
    const float t5 = a3->re * wre - a3->im * wim;
    const float t6 = a3->re * wim + a3->im * wre;
    butterflies(a0, a1, a2, a3, t1, t2, t5, t6);
}

static void fft4(cpx *d, const cpx *s)
{
    float t1, t2, t3, t4, t5, t6, t7, t8;
    t3 = s[0].re - s[1].re; t1 = s[0].re + s[1].re;
    t8 = s[3].re - s[2].re; t6 = s[3].re + s[2].re;
    const float d2re = t1 - t6, d0re = t1 + t6;
    t4 = s[0].im - s[1].im; t2 = s[0].im + s[1].im;
    t7 = s[2].im - s[3].im; t5 = s[2].im + s[3].im;
    d[2].re = d2re; d[0].re = d0re;
    d[3].im = t4 - t8; d[1].im = t4 + t8;
    d[3].re = t3 - t7; d[1].re = t3 + t7;
    d[2].im = t2 - t5; d[0].im = t2 + t5;
}

static void fft_ns(OrcTx *t, int k, cpx *d, const cpx *s)
{
    const int n = 1 << k;
    if (n == 2) {
        const float re = s[0].re - s[1].re, im = s[0].im - s[1].im;
        d[0].re = s[0].re + s[1].re; d[0].im = s[0].im + s[1].im;
        d[1].re = re; d[1].im = im;
    } else if (n == 4) {
        fft4(d, s);
    } else if (n == 8) {
        const float c = t->tab[3][1];
        /* the reference runs fft4 first and then reads src[4..7]; src may alias dst, rows 4..7 are untouched by fft4 */
        fft4(d, s);
        const float t1 = s[4].re - (-s[5].re), d5re = s[4].re + (-s[5].re);
        const float t2 = s[4].im - (-s[5].im), d5im = s[4].im + (-s[5].im);
        const float t5 = s[6].re - (-s[7].re), d7re = s[6].re + (-s[7].re);
        const float t6 = s[6].im - (-s[7].im), d7im = s[6].im + (-s[7].im);
        d[5].re = d5re; d[5].im = d5im; d[7].re = d7re; d[7].im = d7im;
        butterflies(&d[0], &d[2], &d[4], &d[6], t1, t2, t5, t6);
        transform(&d[1], &d[3], &d[5], &d[7], c, c);
    } else if (n == 16) {
        const float *c = t->tab[4];
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
        const float *cs = t->tab[k], *wim = cs + o1 - 7;
        cpx *z = d;
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

/* ---- compound 15 x M MDCT: what av_tx_init(AV_TX_FLOAT_MDCT, len = 15 * 2^k) selects (mdct_pfa_15xM has the largest factor,
 * tx.c:391-395), e.g. the Opus CELT sizes 120 ... 960 (libavcodec/opus/dec_celt.c:569) -------------------------------------- */
static int mulinv(int n, int m)                                /* tx.c:34-42 */
{
    n = n % m;
    for (int x = 1; x < m; x++) if (((n * x) % m) == 1) return x;
    return 0;
}

/* ff_tx_tab_53 / _7 / _9 (tx_template.c:91-130) */
static void init_tab_odd(OrcTx *t)
{
    const double c5 = cos(2 * M_PI / 5), c10 = cos(2 * M_PI / 10), s5 = sin(2 * M_PI / 5), s10 = sin(2 * M_PI / 10);
    t->tab53[0] = t->tab53[1] = (float)c5; t->tab53[2] = t->tab53[3] = (float)c10;
    t->tab53[4] = t->tab53[5] = (float)s5; t->tab53[6] = t->tab53[7] = (float)s10;
    t->tab53[8] = t->tab53[9] = (float)cos(2 * M_PI / 12); t->tab53[10] = (float)cos(2 * M_PI / 6); t->tab53[11] = (float)cos(8 * M_PI / 6);
    float *t7 = t->tab53 + 12, *t9 = t->tab53 + 18;
    t7[0] = (float)cos(2 * M_PI / 7); t7[1] = (float)sin(2 * M_PI / 7); t7[2] = (float)sin(2 * M_PI / 28);
    t7[3] = (float)cos(2 * M_PI / 28); t7[4] = (float)cos(2 * M_PI / 14); t7[5] = (float)sin(2 * M_PI / 14);
    t9[0] = (float)cos(2 * M_PI / 3); t9[1] = (float)sin(2 * M_PI / 3); t9[2] = (float)cos(2 * M_PI / 9); t9[3] = (float)sin(2 * M_PI / 9);
    t9[4] = (float)cos(2 * M_PI / 36); t9[5] = (float)sin(2 * M_PI / 36); t9[6] = t9[2] + t9[5]; t9[7] = t9[3] - t9[4];
}

static OrcTx *open_mdct_pfa(int n, int inv, int len, float scale)
{
    OrcTx *t = calloc(1, sizeof(*t));
    const int l2 = len >> 1, m = l2 / n;
    t->type = 1; t->inv = !!inv; t->len = len; t->n = m; t->pfa_m = m; t->pfa_n = n;
    for (int j = 3; (1 << j) <= m; j++) make_tab(t, j);
    /* ff_tx_gen_compound_mapping(s, opts = NULL -> gather, inv, 15, m), tx.c:75-123 */
    int *in_map = t->pfa_in = malloc(sizeof(int) * l2), *out_map = t->pfa_out = malloc(sizeof(int) * l2);
    const int m_inv = mulinv(m, n), n_inv = mulinv(n, m);
    for (int j = 0; j < m; j++)
        for (int i = 0; i < n; i++) {
            in_map[j * n + i] = (i * m + j * n) % l2;
            out_map[(i * m * m_inv + j * n * n_inv) % l2] = i * m + j;
        }
    if (inv)
        for (int i = 0; i < m; i++) {
            int *in = &in_map[i * n + 1];                      /* skip the DC */
            for (int j = 0; j < ((n - 1) >> 1); j++) { int x = in[j]; in[j] = in[n - j - 2]; in[n - j - 2] = x; }
        }
    /* TX_EMBED_INPUT_PFA_MAP(map, len, 3, 5), tx_priv.h:275-284: the 15-point transform is itself a 3 x 5 compound */
    for (int k = 0; n == 15 && k < l2; k += 15) {
        int mt[15];
        memcpy(mt, &in_map[k], sizeof(mt));
        for (int b = 0; b < 5; b++) for (int a = 0; a < 3; a++) in_map[k + b * 3 + a] = mt[(b * 3 + a * 5) % 15];
    }
    /* ff_tx_mdct_gen_exp(s, inv ? s->map : NULL), tx_template.c:2107-2134 */
    const int len4 = l2;
    const double theta = (scale < 0 ? len4 : 0) + 1.0 / 8.0, sc = sqrt(fabs((double)scale));
    cpx *full = malloc(sizeof(cpx) * len4);
    for (int i = 0; i < len4; i++) {
        const double alpha = M_PI_2 * (i + theta) / len4;
        full[i].re = (float)(cos(alpha) * sc);
        full[i].im = (float)(sin(alpha) * sc);
    }
    if (inv) {
        t->exp = malloc(sizeof(cpx) * 2 * len4);
        memcpy(t->exp + len4, full, sizeof(cpx) * len4);
        for (int i = 0; i < len4; i++) t->exp[i] = full[in_map[i]];
        free(full);
    } else
        t->exp = full;
    for (int i = 0; i < l2; i++) in_map[i] <<= 1;              /* "saves multiplies in loops" (tx_template.c:1460-1462) */
    /* sub-transform: m-point FFT, in place, pre-shuffled, scatter map (ff_tx_mdct_pfa_init sub_opts; tx.c:136-154) */
    t->pfa_sub = malloc(sizeof(int) * m);
    for (int i = 0; i < m; i++) t->pfa_sub[(-sr_perm(i, m, t->inv)) & (m - 1)] = i;
    t->pfa_tmp = malloc(sizeof(cpx) * l2);
    init_tab_odd(t);
    return t;
}

/* fft3 (tx_template.c:172-209, float branch): out with stride */
static void pfa_fft3(const float *tab, cpx *out, const cpx *in, int stride)
{
    cpx t0 = in[0], t1, t2;
    t1.re = in[1].im - in[2].im; t2.im = in[1].im + in[2].im;
    t1.im = in[1].re - in[2].re; t2.re = in[1].re + in[2].re;
    out[0 * stride].re = t0.re + t2.re;
    out[0 * stride].im = t0.im + t2.im;
    t1.re = tab[8] * t1.re; t1.im = tab[9] * t1.im; t2.re = tab[10] * t2.re; t2.im = tab[10] * t2.im;
    out[1 * stride].re = t0.re - t2.re + t1.re;
    out[1 * stride].im = t0.im - t2.im - t1.im;
    out[2 * stride].re = t0.re - t2.re - t1.re;
    out[2 * stride].im = t0.im - t2.im + t1.im;
}

/* DECL_FFT5 (tx_template.c:211-250): d[] = the five output slots of the variant (fft5_m1 / _m2 / _m3) */
static void pfa_fft5(const float *tab, cpx *out, const cpx *in, int stride, const int d[5])
{
    cpx dc = in[0], z0[4], t[6];
    t[1].im = in[1].re - in[4].re; t[0].re = in[1].re + in[4].re;
    t[1].re = in[1].im - in[4].im; t[0].im = in[1].im + in[4].im;
    t[3].im = in[2].re - in[3].re; t[2].re = in[2].re + in[3].re;
    t[3].re = in[2].im - in[3].im; t[2].im = in[2].im + in[3].im;
    out[d[0] * stride].re = dc.re + t[0].re + t[2].re;
    out[d[0] * stride].im = dc.im + t[0].im + t[2].im;
    { const float a = tab[0] * t[2].re - tab[2] * t[0].re, b = tab[0] * t[0].re - tab[2] * t[2].re; t[4].re = a; t[0].re = b; }   /* SMUL */
    { const float a = tab[0] * t[2].im - tab[2] * t[0].im, b = tab[0] * t[0].im - tab[2] * t[2].im; t[4].im = a; t[0].im = b; }
    { const float a = tab[4] * t[3].re - tab[6] * t[1].re, b = tab[4] * t[1].re + tab[6] * t[3].re; t[5].re = a; t[1].re = b; }   /* CMUL */
    { const float a = tab[4] * t[3].im - tab[6] * t[1].im, b = tab[4] * t[1].im + tab[6] * t[3].im; t[5].im = a; t[1].im = b; }
    z0[0].re = t[0].re - t[1].re; z0[3].re = t[0].re + t[1].re;
    z0[0].im = t[0].im - t[1].im; z0[3].im = t[0].im + t[1].im;
    z0[2].re = t[4].re - t[5].re; z0[1].re = t[4].re + t[5].re;
    z0[2].im = t[4].im - t[5].im; z0[1].im = t[4].im + t[5].im;
    out[d[1] * stride].re = dc.re + z0[3].re; out[d[1] * stride].im = dc.im + z0[0].im;
    out[d[2] * stride].re = dc.re + z0[2].re; out[d[2] * stride].im = dc.im + z0[1].im;
    out[d[3] * stride].re = dc.re + z0[1].re; out[d[3] * stride].im = dc.im + z0[2].im;
    out[d[4] * stride].re = dc.re + z0[0].re; out[d[4] * stride].im = dc.im + z0[3].im;
}

/* fft15 (tx_template.c:465-476) */
static void pfa_fft15(const float *tab, cpx *out, const cpx *in, int stride)
{
    static const int m1[5] = { 0, 6, 12, 3, 9 }, m2[5] = { 10, 1, 7, 13, 4 }, m3[5] = { 5, 11, 2, 8, 14 };
    cpx tmp[15];
    for (int i = 0; i < 5; i++) pfa_fft3(tab, tmp + i, in + i * 3, 5);
    pfa_fft5(tab, out, tmp + 0, stride, m1);
    pfa_fft5(tab, out, tmp + 5, stride, m2);
    pfa_fft5(tab, out, tmp + 10, stride, m3);
}

/* fft7 (tx_template.c:252-339, float branch).  S[k] / D[k]: sum / difference of in[k + 1] and in[6 - k]; T = ff_tx_tab_7 as three
 * (re, im) pairs.  Every expression keeps the reference's left-to-right order of products and sums. */
static void pfa_fft7(const float *T, cpx *out, const cpx *in, int stride)
{
    const float c0 = T[0], s0 = T[1], c1 = T[2], s1 = T[3], c2 = T[4], s2 = T[5];
    const cpx dc = in[0];
    cpx S[3], D[3], z[3], u[3];
    for (int k = 0; k < 3; k++) {
        S[k].re = in[k + 1].re + in[6 - k].re; D[k].re = in[k + 1].re - in[6 - k].re;
        S[k].im = in[k + 1].im + in[6 - k].im; D[k].im = in[k + 1].im - in[6 - k].im;
    }
    out[0].re = dc.re + S[0].re + S[1].re + S[2].re;
    out[0].im = dc.im + S[0].im + S[1].im + S[2].im;
    z[0].re = c0 * S[0].re - c2 * S[2].re - c1 * S[1].re;
    z[1].re = c0 * S[2].re - c1 * S[0].re - c2 * S[1].re;
    z[2].re = c0 * S[1].re - c2 * S[0].re - c1 * S[2].re;
    z[0].im = c0 * S[0].im - c1 * S[1].im - c2 * S[2].im;
    z[1].im = c0 * S[2].im - c1 * S[0].im - c2 * S[1].im;
    z[2].im = c0 * S[1].im - c2 * S[0].im - c1 * S[2].im;
    u[0].re = s2 * D[0].im + s1 * D[2].im - s0 * D[1].im;
    u[1].re = s0 * D[2].im + s2 * D[1].im - s1 * D[0].im;
    u[2].re = s2 * D[2].im + s1 * D[1].im + s0 * D[0].im;
    u[0].im = s0 * D[0].re + s1 * D[1].re + s2 * D[2].re;
    u[1].im = s2 * D[1].re + s0 * D[2].re - s1 * D[0].re;
    u[2].im = s2 * D[0].re + s1 * D[2].re - s0 * D[1].re;
    out[1 * stride].re = dc.re + (z[0].re + u[2].re); out[1 * stride].im = dc.im + (z[0].im - u[0].im);
    out[2 * stride].re = dc.re + (z[1].re - u[1].re); out[2 * stride].im = dc.im + (z[1].im + u[1].im);
    out[3 * stride].re = dc.re + (z[2].re + u[0].re); out[3 * stride].im = dc.im + (z[2].im - u[2].im);
    out[4 * stride].re = dc.re + (z[2].re - u[0].re); out[4 * stride].im = dc.im + (z[2].im + u[2].im);
    out[5 * stride].re = dc.re + (z[1].re + u[1].re); out[5 * stride].im = dc.im + (z[1].im - u[1].im);
    out[6 * stride].re = dc.re + (z[0].re - u[2].re); out[6 * stride].im = dc.im + (z[0].im + u[0].im);
}

/* fft9 (tx_template.c:341-463, float branch).  S[k] / D[k]: sum / difference of in[k + 1] and in[8 - k]; T = ff_tx_tab_9 as four pairs. */
static void pfa_fft9(const float *T, cpx *out, const cpx *in, int stride)
{
    const cpx dc = in[0];
    cpx S[4], D[4], w[4], x[5], y[5], z0, z1;
    for (int k = 0; k < 4; k++) {
        S[k].re = in[k + 1].re + in[8 - k].re; D[k].re = in[k + 1].re - in[8 - k].re;
        S[k].im = in[k + 1].im + in[8 - k].im; D[k].im = in[k + 1].im - in[8 - k].im;
    }
    w[0].re = S[0].re - S[3].re; w[0].im = S[0].im - S[3].im;
    w[1].re = S[1].re - S[3].re; w[1].im = S[1].im - S[3].im;
    w[2].re = D[0].re - D[3].re; w[2].im = D[0].im - D[3].im;
    w[3].re = D[1].re + D[3].re; w[3].im = D[1].im + D[3].im;
    z0.re = dc.re + S[2].re; z0.im = dc.im + S[2].im;
    z1.re = S[0].re + S[1].re + S[3].re; z1.im = S[0].im + S[1].im + S[3].im;
    out[0].re = z0.re + z1.re; out[0].im = z0.im + z1.im;
    y[3].re = T[1] * (D[0].re - D[1].re + D[3].re);
    y[3].im = T[1] * (D[0].im - D[1].im + D[3].im);
    x[3].re = z0.re + T[0] * z1.re; x[3].im = z0.im + T[0] * z1.im;
    z0.re = dc.re + T[0] * S[2].re; z0.im = dc.im + T[0] * S[2].im;
    x[1].re = T[2] * w[0].re + T[5] * w[1].re; x[1].im = T[2] * w[0].im + T[5] * w[1].im;
    x[2].re = T[5] * w[0].re - T[6] * w[1].re; x[2].im = T[5] * w[0].im - T[6] * w[1].im;
    y[1].re = T[3] * w[2].re + T[4] * w[3].re; y[1].im = T[3] * w[2].im + T[4] * w[3].im;
    y[2].re = T[4] * w[2].re - T[7] * w[3].re; y[2].im = T[4] * w[2].im - T[7] * w[3].im;
    y[0].re = T[1] * D[2].re; y[0].im = T[1] * D[2].im;
    x[4].re = x[1].re + x[2].re; x[4].im = x[1].im + x[2].im;
    y[4].re = y[1].re - y[2].re; y[4].im = y[1].im - y[2].im;
    x[1].re = z0.re + x[1].re; x[1].im = z0.im + x[1].im;
    y[1].re = y[0].re + y[1].re; y[1].im = y[0].im + y[1].im;
    x[2].re = z0.re + x[2].re; x[2].im = z0.im + x[2].im;
    y[2].re = y[2].re - y[0].re; y[2].im = y[2].im - y[0].im;
    x[4].re = z0.re - x[4].re; x[4].im = z0.im - x[4].im;
    y[4].re = y[0].re - y[4].re; y[4].im = y[0].im - y[4].im;
    for (int k = 1; k <= 4; k++) {
        out[k * stride].re = x[k].re + y[k].im; out[k * stride].im = x[k].im - y[k].re;
        out[(9 - k) * stride].re = x[k].re - y[k].im; out[(9 - k) * stride].im = x[k].im + y[k].re;
    }
}

/* fft3 / fft5 / fft7 / fft9 / fft15 as the DECL_COMP_* macros instantiate them */
static void pfa_fftN(int n, const float *tab, cpx *out, const cpx *in, int stride)
{
    static const int d5[5] = { 0, 1, 2, 3, 4 };
    if (n == 3) pfa_fft3(tab, out, in, stride);
    else if (n == 5) pfa_fft5(tab, out, in, stride, d5);
    else if (n == 7) pfa_fft7(tab + 12, out, in, stride);
    else if (n == 9) pfa_fft9(tab + 18, out, in, stride);
    else pfa_fft15(tab, out, in, stride);
}

static void fft_ns(OrcTx *t, int k, cpx *d, const cpx *s);

/* ff_tx_mdct_pfa_15xM_inv (DECL_COMP_IMDCT, tx_template.c:1471-1511): len floats with a stride in, len/2 complex (= len floats) out */
static void run_mdct_pfa_inv(OrcTx *t, float *out, const float *in, ptrdiff_t stride)
{
    cpx *z = (cpx *)out, f15[15];
    const cpx *e = t->exp;
    const int len4 = t->len >> 2, len2 = t->len >> 1, m = t->pfa_m, N = t->pfa_n;
    const int *in_map = t->pfa_in, *sub_map = t->pfa_sub;
    const float *in1 = in, *in2 = in + ((N * m * 2) - 1) * stride;
    for (int i = 0; i < len2; i += N) {
        for (int j = 0; j < N; j++) {
            const int k = in_map[j];
            const float are = in2[-k * stride], aim = in1[k * stride];
            f15[j].re = are * e[j].re - aim * e[j].im;                  /* CMUL3 */
            f15[j].im = are * e[j].im + aim * e[j].re;
        }
        pfa_fftN(N, t->tab53, t->pfa_tmp + *(sub_map++), f15, m);
        e += N; in_map += N;
    }
    for (int i = 0; i < N; i++) fft_ns(t, ilog2i(m), t->pfa_tmp + m * i, t->pfa_tmp + m * i);
    for (int i = 0; i < len4; i++) {
        const int i0 = len4 + i, i1 = len4 - i - 1, s0 = t->pfa_out[i0], s1 = t->pfa_out[i1];
        const cpx src1 = { t->pfa_tmp[s1].im, t->pfa_tmp[s1].re }, src0 = { t->pfa_tmp[s0].im, t->pfa_tmp[s0].re };
        z[i1].re = src1.re * e[i1].im - src1.im * e[i1].re;
        z[i0].im = src1.re * e[i1].re + src1.im * e[i1].im;
        z[i0].re = src0.re * e[i0].im - src0.im * e[i0].re;
        z[i1].im = src0.re * e[i0].re + src0.im * e[i0].im;
    }
}

/* ff_tx_mdct_pfa_15xM_fwd (DECL_COMP_MDCT, tx_template.c:1533-1579): 2*len floats in, len floats with a stride out */
static void run_mdct_pfa_fwd(OrcTx *t, float *dst, const float *src, ptrdiff_t stride)
{
    cpx f15[15];
    const cpx *e = t->exp;
    const int m = t->pfa_m, N = t->pfa_n, len4 = N * m, len3 = len4 * 3, len8 = t->len >> 2;
    const int *in_map = t->pfa_in, *sub_map = t->pfa_sub;
    for (int i = 0; i < m; i++) {
        for (int j = 0; j < N; j++) {
            const int k = in_map[i * N + j];
            float re, im;
            if (k < len4) { re = -src[len4 + k] + src[1 * len4 - 1 - k]; im = -src[len3 + k] + -src[1 * len3 - 1 - k]; }
            else          { re = -src[len4 + k] + -src[5 * len4 - 1 - k]; im = src[-len4 + k] + -src[1 * len3 - 1 - k]; }
            f15[j].im = re * e[k >> 1].re - im * e[k >> 1].im;
            f15[j].re = re * e[k >> 1].im + im * e[k >> 1].re;
        }
        pfa_fftN(N, t->tab53, t->pfa_tmp + sub_map[i], f15, m);
    }
    for (int i = 0; i < N; i++) fft_ns(t, ilog2i(m), t->pfa_tmp + m * i, t->pfa_tmp + m * i);
    for (int i = 0; i < len8; i++) {
        const int i0 = len8 + i, i1 = len8 - i - 1, s0 = t->pfa_out[i0], s1 = t->pfa_out[i1];
        const cpx src1 = t->pfa_tmp[s1], src0 = t->pfa_tmp[s0];
        dst[2 * i1 * stride + stride] = src0.re * e[i0].im - src0.im * e[i0].re;
        dst[2 * i0 * stride]          = src0.re * e[i0].re + src0.im * e[i0].im;
        dst[2 * i0 * stride + stride] = src1.re * e[i1].im - src1.im * e[i1].re;
        dst[2 * i1 * stride]          = src1.re * e[i1].re + src1.im * e[i1].im;
    }
}

/* ff_tx_dct_init (tx_template.c:1832-1872): the inverse is defined on twice the length it is asked for (callers pass N / 2) */
static OrcTx *open_dct(int inv, int len, float scale)
{
    float rsc = scale;
    if (inv) { len *= 2; rsc *= 0.5f; }
    if (len < 4 || (len & (len - 1))) return NULL;
    OrcTx *sub = orc_tx_open(6, inv, len, rsc, 0);
    if (!sub) return NULL;
    OrcTx *t = calloc(1, sizeof(*t));
    t->type = 9; t->inv = !!inv; t->len = len; t->dct_n = len; t->dct_sub = sub;
    float *tab = t->dct_exp = malloc(sizeof(float) * (len / 2) * 3);
    const double freq = M_PI / (len * 2);
    for (int i = 0; i < len; i++) tab[i] = (float)(cos(i * freq) * (!inv + 1));
    for (int i = 0; i < len / 2; i++)
        tab[len + i] = inv ? (float)(0.5 / sin((2 * i + 1) * freq)) : (float)cos((len - 2 * i - 1) * freq);
    return t;
}

static void run_rdft(OrcTx *t, void *out, void *in);

/* ff_tx_dctII (tx_template.c:1874-1925): len floats in (overwritten), len + 2 floats of room in dst, len results */
static void run_dct2(OrcTx *t, float *dst, float *src)
{
    const int len = t->dct_n, len2 = len >> 1;
    const float *e = t->dct_exp;
    for (int i = 0; i < len2; i++) {
        const float in1 = src[i], in2 = src[len - i - 1], s = e[len + i];
        const float tmp1 = (in1 + in2) * 0.5f, tmp2 = (in1 - in2) * s;
        src[i] = tmp1 + tmp2;
        src[len - i - 1] = tmp1 - tmp2;
    }
    run_rdft(t->dct_sub, dst, src);
    float next = dst[len];
    for (int i = len - 2; i > 0; i -= 2) {
        const float are = e[len - i], aim = e[i], bre = dst[i + 0], bim = dst[i + 1];
        const float tmp = are * bre - aim * bim;                     /* CMUL(tmp, dst[i], exp[len - i], exp[i], dst[i], dst[i + 1]) */
        dst[i] = are * bim + aim * bre;
        dst[i + 1] = next;
        next += tmp;
    }
    dst[0] = e[0] * dst[0];
    dst[1] = next;
}

/* ff_tx_dctIII (tx_template.c:1927-1968): len + 2 floats in (len coefficients, two of padding; overwritten), len floats out */
static void run_dct3(OrcTx *t, float *dst, float *src)
{
    const int len = t->dct_n, len2 = len >> 1;
    const float *e = t->dct_exp;
    float tmp2 = 2 * src[len - 1];
    src[len] = tmp2;
    for (int i = len - 2; i >= 2; i -= 2) {
        const float val1 = src[i - 0], val2 = src[i - 1] - src[i + 1];
        const float are = e[len - i], aim = e[i];
        src[i + 1] = are * val1 - aim * val2;                        /* CMUL(src[i + 1], src[i], exp[len - i], exp[i], val1, val2) */
        src[i] = are * val2 + aim * val1;
    }
    run_rdft(t->dct_sub, dst, src);
    for (int i = 0; i < len2; i++) {
        const float in1 = dst[i], in2 = dst[len - i - 1], c = e[len + i];
        const float tmp1 = in1 + in2;
        tmp2 = in1 - in2;
        tmp2 *= c;
        dst[i] = tmp1 + tmp2;
        dst[len - i - 1] = tmp1 - tmp2;
    }
}

/* ---- compound N x M FFT: av_tx_init(AV_TX_FLOAT_FFT, len = N * 2^k), N = 15, 9, 7, 5 or 3, resolves to fft_pfa over fftN_ns and the
 * 2^k-point split-radix transform (codelet tree printed by the compiled reference: oracle/ref/ref_shim.c ffref_tx_describe; checkasm
 * lengths 120 / 960 / 1920, tests/checkasm/av_tx.c:38-40).  ff_tx_fft_pfa_init (tx_template.c:948-1057), ff_tx_fft_pfa (:1059-1080). */
static OrcTx *open_fft_pfa(int n, int inv, int len)
{
    OrcTx *t = calloc(1, sizeof(*t));
    const int m = len / n;
    t->type = 0; t->inv = !!inv; t->len = len; t->n = m; t->pfa_m = m; t->pfa_n = n;
    for (int j = 3; (1 << j) <= m; j++) make_tab(t, j);
    /* ff_tx_gen_compound_mapping(s, opts, 0, n, m): gather input map, CRT output map; the direction is NOT applied here */
    int *in_map = t->pfa_in = malloc(sizeof(int) * len), *out_map = t->pfa_out = malloc(sizeof(int) * len);
    const int m_inv = mulinv(m, n), n_inv = mulinv(n, m);
    for (int j = 0; j < m; j++)
        for (int i = 0; i < n; i++) {
            in_map[j * n + i] = (i * m + j * n) % len;
            out_map[(i * m * m_inv + j * n * n_inv) % len] = i * m + j;
        }
    /* the map of the first sub-transform (ff_tx_fft_factor_init, tx_template.c:478-494): 3 x 5 map for 15 points
     * (ff_tx_gen_pfa_input_map, tx.c:44-71), else ff_tx_gen_default_map (tx.c:525-542); the inverse reverses all but the DC */
    int sub[15];
    if (n == 15) {
        for (int b = 0; b < 5; b++)
            for (int a = 0; a < 3; a++) {
                if (inv) sub[(b * 3 + a * 5) % 15] = b * 3 + a;
                else     sub[b * 3 + a] = (b * 3 + a * 5) % 15;
            }
        if (inv) for (int w = 1; w <= 7; w++) { int x = sub[w]; sub[w] = sub[15 - w]; sub[15 - w] = x; }
    } else {
        sub[0] = 0;
        for (int i = 1; i < n; i++) sub[i] = inv ? n - i : i;
    }
    for (int k = 0; k < len; k += n) {                         /* "Flatten input map" (tx_template.c:1040-1046) */
        int tmp[15];
        memcpy(tmp, &in_map[k], sizeof(int) * n);
        for (int i = 0; i < n; i++) in_map[k + i] = tmp[sub[i]];
    }
    t->pfa_sub = malloc(sizeof(int) * m);
    for (int i = 0; i < m; i++) t->pfa_sub[(-sr_perm(i, m, t->inv)) & (m - 1)] = i;
    t->pfa_tmp = malloc(sizeof(cpx) * len);
    init_tab_odd(t);
    return t;
}

static void run_fft_pfa(OrcTx *t, cpx *out, const cpx *in, ptrdiff_t stride)
{
    const int n = t->pfa_n, m = t->pfa_m, l = t->len;
    cpx f[15];
    for (int i = 0; i < m; i++) {
        for (int j = 0; j < n; j++) f[j] = in[t->pfa_in[i * n + j]];
        pfa_fftN(n, t->tab53, t->pfa_tmp + t->pfa_sub[i], f, m);
    }
    for (int i = 0; i < n; i++) fft_ns(t, ilog2i(m), t->pfa_tmp + m * i, t->pfa_tmp + m * i);
    for (int i = 0; i < l; i++) out[i * stride] = t->pfa_tmp[t->pfa_out[i]];
}

OrcTx *orc_tx_open(int type, int inv, int len, float scale, unsigned flags)
{
    if (flags == 4) {                                          /* AV_TX_FULL_IMDCT: only the inverse MDCT has such a codelet (tx.c:762-771) */
        if (type != 1 || !inv) return NULL;
        OrcTx *sub = orc_tx_open(type, inv, len, scale, 0);
        if (sub) sub->full = 1;
        return sub;
    }
    if (!flags && type == 9) return open_dct(inv, len, scale);
    if (!flags && type == 1 && len >= 12 && !(len & 1)) {          /* compound MDCT: the largest odd factor wins (tx.c:391-395) */
        static const int factors[5] = { 15, 9, 7, 5, 3 };
        for (int f = 0; f < 5; f++) {
            const int n = factors[f], l2 = len >> 1, m = l2 / n;
            if (l2 % n == 0 && m >= 2 && !(m & (m - 1))) return open_mdct_pfa(n, inv, len, scale);
        }
    }
    if (!flags && type == 0 && len >= 6 && (len & (len - 1))) {   /* compound FFT: odd part 3, 5, 7, 9 or 15 times a power of two >= 2 */
        int m = 1;
        while (!(len & m)) m <<= 1;
        const int n = len / m;
        if (m >= 2 && (n == 3 || n == 5 || n == 7 || n == 9 || n == 15)) return open_fft_pfa(n, inv, len);
    }
    if (flags || (type != 0 && type != 1 && type != 6) || len < 2 || (len & (len - 1))) return NULL;
    if (type == 6 && len < 4) return NULL;                     /* ff_tx_rdft_*_def: min_len 4 */
    if (type == 1 && len < 4) return NULL;                     /* a 2-point MDCT has no 1-point FFT to sit on: the reference falls back to its naive MDCT */
    OrcTx *t = calloc(1, sizeof(*t));
    t->type = type; t->inv = !!inv; t->len = len;
    t->n = type == 0 ? len : len >> 1;
    if (t->n < 1 || t->n > 131072) { free(t); return NULL; }
    const int k = ilog2i(t->n);
    for (int j = 3; j <= k; j++) make_tab(t, j);
    t->map = malloc(sizeof(int) * t->n);
    const int scatter = type == 1 && !inv;                     /* ff_tx_mdct_init: map_dir = !inv ? SCATTER : GATHER */
    for (int i = 0; i < t->n; i++) {
        const int p = t->n == 1 ? 0 : (-sr_perm(i, t->n, t->inv)) & (t->n - 1);
        if (scatter) t->map[p] = i; else t->map[i] = p;
    }
    if (type == 1) {
        const int len4 = t->len >> 1;                          /* named len4 in ff_tx_mdct_gen_exp: it is s->len >> 1 */
        const double theta = (scale < 0 ? len4 : 0) + 1.0 / 8.0;
        const double sc = sqrt(fabs((double)scale));
        cpx *full = malloc(sizeof(cpx) * len4);
        for (int i = 0; i < len4; i++) {
            const double alpha = M_PI_2 * (i + theta) / len4;
            full[i].re = (float)(cos(alpha) * sc);
            full[i].im = (float)(sin(alpha) * sc);
        }
        if (inv) {                                             /* pre-shuffled copy first, natural order after it */
            t->exp = malloc(sizeof(cpx) * 2 * len4);
            memcpy(t->exp + len4, full, sizeof(cpx) * len4);
            for (int i = 0; i < len4; i++) t->exp[i] = full[t->map[i]];
            free(full);
        } else {
            t->exp = full;
        }
        t->sub_map = malloc(sizeof(int) * len4);
        for (int i = 0; i < len4; i++) t->sub_map[i] = inv ? t->map[i] << 1 : t->map[i];
    }
    if (type == 6) {                                           /* ff_tx_rdft_init, tx_template.c:1601-1653 */
        const int len4 = len >> 2;
        const double f = 2 * M_PI / len, m = inv ? 2 * (double)scale : (double)scale;
        float *tab = t->rexp = malloc(sizeof(float) * (8 + 2 * len4));
        *tab++ = (float)((inv ? 0.5 : 1.0) * m);
        *tab++ = (float)(inv ? 0.5 * m : 1.0 * m);
        *tab++ = (float)(m);
        *tab++ = (float)(-m);
        *tab++ = (float)((0.5 - 0.0) * m);
        *tab++ = (float)((0.0 - 0.5) * m);
        *tab++ = (float)((0.5 - inv) * m);
        *tab++ = (float)(-(0.5 - inv) * m);
        for (int i = 0; i < len4; i++) *tab++ = (float)cos(i * f);
        for (int i = 0; i < len4; i++) *tab++ = (float)cos(((len - i * 4) / 4.0) * f) * (inv ? 1 : -1);
    }
    return t;
}

void orc_tx_close(OrcTx *t)
{
    if (!t) return;
    for (int i = 0; i < 18; i++) free(t->tab[i]);
    free(t->map); free(t->sub_map); free(t->exp); free(t->rexp);
    free(t->pfa_in); free(t->pfa_out); free(t->pfa_sub); free(t->pfa_tmp);
    orc_tx_close(t->dct_sub); free(t->dct_exp); free(t);
}

static void run_fft(OrcTx *t, cpx *dst, const cpx *src)
{
    for (int i = 0; i < t->n; i++) dst[i] = src[t->map[i]];
    fft_ns(t, ilog2i(t->n), dst, dst);
}

static void run_mdct_inv(OrcTx *t, float *out, const float *in, ptrdiff_t stride)
{
    cpx *z = (cpx *)out;
    const cpx *e = t->exp;
    const int len2 = t->len >> 1, len4 = t->len >> 2;
    const float *in1 = in, *in2 = in + (len2 * 2 - 1) * stride;
    for (int i = 0; i < len2; i++) {
        const int k = t->sub_map[i];
        const float are = in2[-k * stride], aim = in1[k * stride];
        z[i].re = are * e[i].re - aim * e[i].im;
        z[i].im = are * e[i].im + aim * e[i].re;
    }
    fft_ns(t, ilog2i(len2), z, z);
    e += len2;
    for (int i = 0; i < len4; i++) {
        const int i0 = len4 + i, i1 = len4 - i - 1;
        const cpx s1 = { z[i1].im, z[i1].re }, s0 = { z[i0].im, z[i0].re };
        z[i1].re = s1.re * e[i1].im - s1.im * e[i1].re;
        z[i0].im = s1.re * e[i1].re + s1.im * e[i1].im;
        z[i0].re = s0.re * e[i0].im - s0.im * e[i0].re;
        z[i1].im = s0.re * e[i0].re + s0.im * e[i0].im;
    }
}

static void run_mdct_fwd(OrcTx *t, float *dst, const float *src, ptrdiff_t stride)
{
    cpx *z = (cpx *)dst;
    const cpx *e = t->exp;
    const int len2 = t->len >> 1, len4 = t->len >> 2, len3 = len2 * 3;
    for (int i = 0; i < len2; i++) {
        const int k = 2 * i, idx = t->sub_map[i];
        float re, im;
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
        const cpx s1 = z[i1], s0 = z[i0];
        dst[2 * i1 * stride + stride] = s0.re * e[i0].im - s0.im * e[i0].re;
        dst[2 * i0 * stride]          = s0.re * e[i0].re + s0.im * e[i0].im;
        dst[2 * i0 * stride + stride] = s1.re * e[i1].im - s1.im * e[i1].re;
        dst[2 * i1 * stride]          = s1.re * e[i1].re + s1.im * e[i1].im;
    }
}

/* ff_tx_rdft_r2c / ff_tx_rdft_c2r (DECL_RDFT, tx_template.c:1655-1724): a len/2-point complex FFT plus the even/odd
 * separation butterflies.  r2c: len floats in, len/2+1 complex out.  c2r: len/2+1 complex in (modified in place, like the
 * reference does), len floats out. */
static void rdft_butterflies(const OrcTx *t, cpx *data)
{
    const int len2 = t->len >> 1, len4 = t->len >> 2;
    const float *fact = t->rexp, *tcos = fact + 8, *tsin = tcos + len4;
    float t0re = data[0].re;
    data[0].re = t0re + data[0].im;
    data[0].im = t0re - data[0].im;
    data[0].re = fact[0] * data[0].re;
    data[0].im = fact[1] * data[0].im;
    data[len4].re = fact[2] * data[len4].re;
    data[len4].im = fact[3] * data[len4].im;
    for (int i = 1; i < len4; i++) {
        cpx a, b, c;
        a.re = fact[4] * (data[i].re + data[len2 - i].re);
        a.im = fact[5] * (data[i].im - data[len2 - i].im);
        b.re = fact[6] * (data[i].im + data[len2 - i].im);
        b.im = fact[7] * (data[i].re - data[len2 - i].re);
        c.re = b.re * tcos[i] - b.im * tsin[i];
        c.im = b.re * tsin[i] + b.im * tcos[i];
        data[i].re = a.re + c.re;
        data[i].im = c.im - a.im;
        data[len2 - i].re = a.re - c.re;
        data[len2 - i].im = c.im + a.im;
    }
}

static void run_rdft(OrcTx *t, void *out, void *in)
{
    const int len2 = t->len >> 1;
    if (!t->inv) {
        cpx *data = out;
        run_fft(t, data, in);
        rdft_butterflies(t, data);
        data[len2].re = data[0].im;
        data[0].im = data[len2].im = 0;
    } else {
        cpx *data = in;
        data[0].im = data[len2].re;
        rdft_butterflies(t, data);
        run_fft(t, out, data);
    }
}

void orc_tx_run(OrcTx *t, void *out, void *in, ptrdiff_t stride, int count, ptrdiff_t out_step, ptrdiff_t in_step)
{
    for (int c = 0; c < count; c++) {
        void *o = (uint8_t *)out + c * out_step, *i = (uint8_t *)in + c * in_step;
        if (t->full) {                                         /* ff_tx_mdct_inv_full: the half transform into the middle, then the two mirrors */
            float *d = o;
            const int n = t->len, n2 = n >> 1;
            const ptrdiff_t st = stride / (ptrdiff_t)sizeof(float);
            if (t->pfa_m) run_mdct_pfa_inv(t, d + n2, i, st); else run_mdct_inv(t, d + n2, i, st);
            for (int k = 0; k < n2; k++) {
                d[k * st] = -d[(n - k - 1) * st];
                d[(2 * n - k - 1) * st] = d[(n + k) * st];
            }
        }
        else if (t->type == 9) { if (t->inv) run_dct3(t, o, i); else run_dct2(t, o, i); }
        else if (t->type == 0 && t->pfa_m) run_fft_pfa(t, o, i, stride / (ptrdiff_t)sizeof(cpx));
        else if (t->type == 0) run_fft(t, o, i);
        else if (t->type == 6) run_rdft(t, o, i);
        else if (t->pfa_m && t->inv) run_mdct_pfa_inv(t, o, i, stride / (ptrdiff_t)sizeof(float));
        else if (t->pfa_m) run_mdct_pfa_fwd(t, o, i, stride / (ptrdiff_t)sizeof(float));
        else if (t->inv) run_mdct_inv(t, o, i, stride / (ptrdiff_t)sizeof(float));
        else run_mdct_fwd(t, o, i, stride / (ptrdiff_t)sizeof(float));
    }
}
