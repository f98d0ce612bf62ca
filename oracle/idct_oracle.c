/*
 * idct_oracle.c — TEST INFRASTRUCTURE ONLY.  CPU restatement of libavcodec's 8-bit "simple" IDCT.
 *
 * Follows (behaviour, not text):
 *   libavcodec/simple_idct_template.c:48-61   constants W1..W7 (W4 = 16383), ROW_SHIFT 11, COL_SHIFT 20, DC_SHIFT 3
 *   libavcodec/simple_idct_template.c:114-206 idctRowCondDC: rows whose AC terms are all zero become (row[0]*8)&0xffff,
 *                                             other rows use the rounded butterfly; results are stored back as int16
 *   libavcodec/simple_idct_template.c:209-327 IDCT_COLS + idctSparseCol{,Put,Add}
 *   libavcodec/simple_idct_template.c:329-368 ff_simple_idct_{put,add,}_int16_8bit
 *   libavcodec/idctdsp.c:73-165               put / put_signed / add _pixels_clamped
 * The zero-tests inside the reference's column pass only skip additions of zero, so they are dropped here.
 * All arithmetic is unsigned 32-bit (wrap-around) exactly like the reference's SUINT.
 */
#include "oracle.h"

enum { W1 = 22725, W2 = 21407, W3 = 19266, W4 = 16383, W5 = 12873, W6 = 8867, W7 = 4520 };

static uint8_t clip8(int a) { return (uint8_t)(a < 0 ? 0 : a > 255 ? 255 : a); }

static void row_pass(int16_t *r)
{
    if (!(r[1] | r[2] | r[3] | r[4] | r[5] | r[6] | r[7])) {
        int16_t dc = (int16_t)(uint16_t)(((unsigned)r[0] << 3) & 0xffff);
        for (int i = 0; i < 8; i++) r[i] = dc;
        return;
    }
    uint32_t a0 = (uint32_t)W4 * r[0] + (1u << 10), a1 = a0, a2 = a0, a3 = a0;
    a0 += (uint32_t)W2 * r[2]; a1 += (uint32_t)W6 * r[2];
    a2 -= (uint32_t)W6 * r[2]; a3 -= (uint32_t)W2 * r[2];
    uint32_t b0 = (uint32_t)(W1 * r[1]) + (uint32_t)(W3 * r[3]);
    uint32_t b1 = (uint32_t)(W3 * r[1]) - (uint32_t)(W7 * r[3]);
    uint32_t b2 = (uint32_t)(W5 * r[1]) - (uint32_t)(W1 * r[3]);
    uint32_t b3 = (uint32_t)(W7 * r[1]) - (uint32_t)(W5 * r[3]);
    a0 += (uint32_t)W4 * r[4] + (uint32_t)W6 * r[6];
    a1 += (uint32_t)-W4 * r[4] - (uint32_t)W2 * r[6];
    a2 += (uint32_t)-W4 * r[4] + (uint32_t)W2 * r[6];
    a3 += (uint32_t)W4 * r[4] - (uint32_t)W6 * r[6];
    b0 += (uint32_t)(W5 * r[5]) + (uint32_t)(W7 * r[7]);
    b1 += (uint32_t)(-W1 * r[5]) + (uint32_t)(-W5 * r[7]);
    b2 += (uint32_t)(W7 * r[5]) + (uint32_t)(W3 * r[7]);
    b3 += (uint32_t)(W3 * r[5]) + (uint32_t)(-W1 * r[7]);
    r[0] = (int16_t)((int32_t)(a0 + b0) >> 11); r[7] = (int16_t)((int32_t)(a0 - b0) >> 11);
    r[1] = (int16_t)((int32_t)(a1 + b1) >> 11); r[6] = (int16_t)((int32_t)(a1 - b1) >> 11);
    r[2] = (int16_t)((int32_t)(a2 + b2) >> 11); r[5] = (int16_t)((int32_t)(a2 - b2) >> 11);
    r[3] = (int16_t)((int32_t)(a3 + b3) >> 11); r[4] = (int16_t)((int32_t)(a3 - b3) >> 11);
}

static void col_pass(const int16_t *c, int out[8])
{
    uint32_t a0 = (uint32_t)W4 * (uint32_t)(c[0] + ((1 << 19) / W4)), a1 = a0, a2 = a0, a3 = a0;
    a0 += (uint32_t)W2 * c[16]; a1 += (uint32_t)W6 * c[16];
    a2 += (uint32_t)-W6 * c[16]; a3 += (uint32_t)-W2 * c[16];
    uint32_t b0 = (uint32_t)(W1 * c[8]), b1 = (uint32_t)(W3 * c[8]), b2 = (uint32_t)(W5 * c[8]), b3 = (uint32_t)(W7 * c[8]);
    b0 += (uint32_t)(W3 * c[24]); b1 += (uint32_t)(-W7 * c[24]); b2 += (uint32_t)(-W1 * c[24]); b3 += (uint32_t)(-W5 * c[24]);
    a0 += (uint32_t)W4 * c[32]; a1 += (uint32_t)-W4 * c[32]; a2 += (uint32_t)-W4 * c[32]; a3 += (uint32_t)W4 * c[32];
    b0 += (uint32_t)(W5 * c[40]); b1 += (uint32_t)(-W1 * c[40]); b2 += (uint32_t)(W7 * c[40]); b3 += (uint32_t)(W3 * c[40]);
    a0 += (uint32_t)W6 * c[48]; a1 += (uint32_t)-W2 * c[48]; a2 += (uint32_t)W2 * c[48]; a3 += (uint32_t)-W6 * c[48];
    b0 += (uint32_t)(W7 * c[56]); b1 += (uint32_t)(-W5 * c[56]); b2 += (uint32_t)(W3 * c[56]); b3 += (uint32_t)(-W1 * c[56]);
    out[0] = (int32_t)(a0 + b0) >> 20; out[1] = (int32_t)(a1 + b1) >> 20;
    out[2] = (int32_t)(a2 + b2) >> 20; out[3] = (int32_t)(a3 + b3) >> 20;
    out[4] = (int32_t)(a3 - b3) >> 20; out[5] = (int32_t)(a2 - b2) >> 20;
    out[6] = (int32_t)(a1 - b1) >> 20; out[7] = (int32_t)(a0 - b0) >> 20;
}

void orc_idct(int16_t *block)
{
    int o[8];
    for (int i = 0; i < 8; i++) row_pass(block + 8 * i);
    for (int i = 0; i < 8; i++) {
        col_pass(block + i, o);
        for (int k = 0; k < 8; k++) block[8 * k + i] = (int16_t)o[k];
    }
}

void orc_idct_put(uint8_t *dest, ptrdiff_t ls, int16_t *block)
{
    int o[8];
    for (int i = 0; i < 8; i++) row_pass(block + 8 * i);
    for (int i = 0; i < 8; i++) {
        col_pass(block + i, o);
        for (int k = 0; k < 8; k++) dest[k * ls + i] = clip8(o[k]);
    }
}

void orc_idct_add(uint8_t *dest, ptrdiff_t ls, int16_t *block)
{
    int o[8];
    for (int i = 0; i < 8; i++) row_pass(block + 8 * i);
    for (int i = 0; i < 8; i++) {
        col_pass(block + i, o);
        for (int k = 0; k < 8; k++) dest[k * ls + i] = clip8(dest[k * ls + i] + o[k]);
    }
}

void orc_idct_batch(int kind, int16_t *blocks, int nblocks, uint8_t *dest, ptrdiff_t line_size, const int64_t *dest_off)
{
    for (int i = 0; i < nblocks; i++) {
        int16_t *b = blocks + 64 * (size_t)i;
        if (kind == 0)      orc_idct(b);
        else if (kind == 1) orc_idct_put(dest + dest_off[i], line_size, b);
        else                orc_idct_add(dest + dest_off[i], line_size, b);
    }
}

/* ------------------------------------------------------------------ simple IDCT, 10 and 12 bit (int16 coefficients)
 * libavcodec/simple_idct_template.c:63-104 (constants: 10 bit W3 = 19265, W4 = 16384, ROW_SHIFT 12, COL_SHIFT 19, DC_SHIFT 2;
 * 12 bit W = 45451 ... 9041, ROW_SHIFT 16, COL_SHIFT 17, DC_SHIFT -1), :114-206 row pass with the DC-only shortcut
 * ((row[0] << 2) for 10 bit, (row[0] + 1) >> 1 for 12 bit, both & 0xffff), :209-257 columns, :281-368 put / add / in place on
 * uint16 pixels clipped to the bit depth.  Installed by ff_idctdsp_init for bits_per_raw_sample 9, 10 and 12
 * (idctdsp.c:248-266).  All sums mod 2^32 as in the reference (SUINT / unsigned MUL, MAC). */
typedef struct { uint32_t w[8]; int row_shift, col_shift, dc_shift, depth, extra; } HbdConst;
static const HbdConst HBD10 = { { 0, 22725, 21407, 19265, 16384, 12873, 8867, 4520 }, 12, 19, 2, 10, 0 };
static const HbdConst HBD12 = { { 0, 45451, 42813, 38531, 32767, 25746, 17734, 9041 }, 16, 17, -1, 12, 0 };
/* the EXTRA_SHIFT instantiation of the 10-bit constants that proresdsp.c makes (simple_idct_template.c:73-76: ROW_SHIFT 13,
 * COL_SHIFT 18, DC_SHIFT 1), run with extra_shift = 2 in the row pass (proresdsp.c:61-62) */
static const HbdConst PRORES10 = { { 0, 22725, 21407, 19265, 16384, 12873, 8867, 4520 }, 13, 18, 1, 10, 2 };

static void hbd_row(const HbdConst *k, int16_t *r)
{
    const uint32_t *W = k->w;
    if (!(r[1] | r[2] | r[3] | r[4] | r[5] | r[6] | r[7])) {
        const int dsh = k->dc_shift - k->extra;
        int t = dsh >= 0 ? r[0] * (1 << dsh) : (r[0] + (1 << (-dsh - 1))) >> -dsh;
        int16_t dc = (int16_t)(uint16_t)(t & 0xffff);
        for (int i = 0; i < 8; i++) r[i] = dc;
        return;
    }
    uint32_t a0 = W[4] * (uint32_t)r[0] + (1u << (k->row_shift + k->extra - 1)), a1 = a0, a2 = a0, a3 = a0;
    a0 += W[2] * (uint32_t)r[2]; a1 += W[6] * (uint32_t)r[2]; a2 -= W[6] * (uint32_t)r[2]; a3 -= W[2] * (uint32_t)r[2];
    uint32_t b0 = W[1] * (uint32_t)r[1] + W[3] * (uint32_t)r[3];
    uint32_t b1 = W[3] * (uint32_t)r[1] - W[7] * (uint32_t)r[3];
    uint32_t b2 = W[5] * (uint32_t)r[1] - W[1] * (uint32_t)r[3];
    uint32_t b3 = W[7] * (uint32_t)r[1] - W[5] * (uint32_t)r[3];
    a0 += W[4] * (uint32_t)r[4] + W[6] * (uint32_t)r[6];
    a1 += -W[4] * (uint32_t)r[4] - W[2] * (uint32_t)r[6];
    a2 += -W[4] * (uint32_t)r[4] + W[2] * (uint32_t)r[6];
    a3 += W[4] * (uint32_t)r[4] - W[6] * (uint32_t)r[6];
    b0 += W[5] * (uint32_t)r[5] + W[7] * (uint32_t)r[7];
    b1 += -W[1] * (uint32_t)r[5] - W[5] * (uint32_t)r[7];
    b2 += W[7] * (uint32_t)r[5] + W[3] * (uint32_t)r[7];
    b3 += W[3] * (uint32_t)r[5] - W[1] * (uint32_t)r[7];
    const int sh = k->row_shift + k->extra;
    r[0] = (int16_t)((int32_t)(a0 + b0) >> sh); r[7] = (int16_t)((int32_t)(a0 - b0) >> sh);
    r[1] = (int16_t)((int32_t)(a1 + b1) >> sh); r[6] = (int16_t)((int32_t)(a1 - b1) >> sh);
    r[2] = (int16_t)((int32_t)(a2 + b2) >> sh); r[5] = (int16_t)((int32_t)(a2 - b2) >> sh);
    r[3] = (int16_t)((int32_t)(a3 + b3) >> sh); r[4] = (int16_t)((int32_t)(a3 - b3) >> sh);
}

static void hbd_col(const HbdConst *k, const int16_t *c, int out[8])
{
    const uint32_t *W = k->w;
    uint32_t a0 = W[4] * (uint32_t)(c[0] + (int)((1u << (k->col_shift - 1)) / W[4])), a1 = a0, a2 = a0, a3 = a0;
    a0 += W[2] * (uint32_t)c[16]; a1 += W[6] * (uint32_t)c[16]; a2 -= W[6] * (uint32_t)c[16]; a3 -= W[2] * (uint32_t)c[16];
    uint32_t b0 = W[1] * (uint32_t)c[8], b1 = W[3] * (uint32_t)c[8], b2 = W[5] * (uint32_t)c[8], b3 = W[7] * (uint32_t)c[8];
    b0 += W[3] * (uint32_t)c[24]; b1 -= W[7] * (uint32_t)c[24]; b2 -= W[1] * (uint32_t)c[24]; b3 -= W[5] * (uint32_t)c[24];
    a0 += W[4] * (uint32_t)c[32]; a1 -= W[4] * (uint32_t)c[32]; a2 -= W[4] * (uint32_t)c[32]; a3 += W[4] * (uint32_t)c[32];
    b0 += W[5] * (uint32_t)c[40]; b1 -= W[1] * (uint32_t)c[40]; b2 += W[7] * (uint32_t)c[40]; b3 += W[3] * (uint32_t)c[40];
    a0 += W[6] * (uint32_t)c[48]; a1 -= W[2] * (uint32_t)c[48]; a2 += W[2] * (uint32_t)c[48]; a3 -= W[6] * (uint32_t)c[48];
    b0 += W[7] * (uint32_t)c[56]; b1 -= W[5] * (uint32_t)c[56]; b2 += W[3] * (uint32_t)c[56]; b3 -= W[1] * (uint32_t)c[56];
    const int sh = k->col_shift;
    out[0] = (int32_t)(a0 + b0) >> sh; out[1] = (int32_t)(a1 + b1) >> sh; out[2] = (int32_t)(a2 + b2) >> sh; out[3] = (int32_t)(a3 + b3) >> sh;
    out[4] = (int32_t)(a3 - b3) >> sh; out[5] = (int32_t)(a2 - b2) >> sh; out[6] = (int32_t)(a1 - b1) >> sh; out[7] = (int32_t)(a0 - b0) >> sh;
}

/* depth 10 (also what 9-bit content uses) or 12; kind 0 in place, 1 put, 2 add; dest = uint16 pixels, line_size in bytes */
int orc_idct_hbd(int depth, int kind, uint8_t *dest_, ptrdiff_t line_size, int16_t *block)
{
    const HbdConst *k = depth == 10 ? &HBD10 : depth == 12 ? &HBD12 : 0;
    if (!k || kind < 0 || kind > 2) return -1;
    uint16_t *dest = (uint16_t *)dest_;
    const ptrdiff_t ls = line_size / 2;
    const int maxv = (1 << depth) - 1;
    int o[8];
    for (int i = 0; i < 8; i++) hbd_row(k, block + 8 * i);
    for (int i = 0; i < 8; i++) {
        hbd_col(k, block + i, o);
        for (int j = 0; j < 8; j++) {
            if (kind == 0) block[8 * j + i] = (int16_t)o[j];
            else {
                int v = kind == 1 ? o[j] : dest[j * ls + i] + o[j];
                dest[j * ls + i] = (uint16_t)(v < 0 ? 0 : v > maxv ? maxv : v);
            }
        }
    }
    return 0;
}

/* ProresDSPContext.idct_put (libavcodec/proresdsp.c:56-82,102-167): block[i] *= qmat[i] (int16 wrap), the row pass (10 bit: the
 * extra-shift constants with two more bits of shift; 12 bit: the plain 12-bit rows), 8192 added to the first row, the in-place column
 * pass, then pixels clipped to [4, 2^bits - 5].  out = uint16 pixels, linesize in bytes. */
int orc_prores_idct_put(int bits, uint8_t *out_, ptrdiff_t linesize, int16_t *block, const int16_t *qmat)
{
    const HbdConst *k = bits == 10 ? &PRORES10 : bits == 12 ? &HBD12 : 0;
    if (!k) return -1;
    uint16_t *out = (uint16_t *)out_;
    const ptrdiff_t ls = linesize >> 1;
    const int lo = 4, hi = (1 << bits) - 4 - 1;
    int o[8];
    for (int i = 0; i < 64; i++) block[i] = (int16_t)(block[i] * qmat[i]);
    for (int i = 0; i < 8; i++) hbd_row(k, block + 8 * i);
    for (int i = 0; i < 8; i++) {
        block[i] = (int16_t)(block[i] + 8192);
        hbd_col(k, block + i, o);
        for (int j = 0; j < 8; j++) block[8 * j + i] = (int16_t)o[j];
    }
    for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++) {
            const int v = block[8 * y + x];
            out[y * ls + x] = (uint16_t)(v < lo ? lo : v > hi ? hi : v);
        }
    return 0;
}

/* idctdsp.c:73-165; kind 0 put, 1 put_signed, 2 add */
void orc_pixels_clamped(int kind, const int16_t *block, uint8_t *pixels, ptrdiff_t ls)
{
    for (int i = 0; i < 8; i++)
        for (int j = 0; j < 8; j++) {
            int b = block[8 * i + j];
            uint8_t *p = pixels + i * ls + j;
            if (kind == 0)      *p = clip8(b);
            else if (kind == 1) *p = b < -128 ? 0 : b > 127 ? 255 : (uint8_t)(b + 128);
            else                *p = clip8(*p + b);
        }
}

#include <string.h>
static int clip_u8_i(int a) { return a < 0 ? 0 : a > 255 ? 255 : a; }

/* ------------------------------------------------------------------ H.264 residual transforms (8 bit)
 * libavcodec/h264idct_template.c:33-70 (ff_h264_idct_add), :72-145 (ff_h264_idct8_add), :147-181 (the two DC-only adds).
 * Intermediate results go back into the int16 block (dctcoef), sums are formed mod 2^32 (SUINT), every function leaves
 * the coefficients it consumed zeroed.  kind 0: 4x4, 1: 8x8, 2: 4x4 DC only, 3: 8x8 DC only. */
static void h264_idct4_add(uint8_t *dst, int16_t *b, ptrdiff_t stride)
{
    b[0] += 1 << 5;
    for (int i = 0; i < 4; i++) {
        const unsigned z0 = b[i] + (unsigned)b[i + 8], z1 = b[i] - (unsigned)b[i + 8];
        const unsigned z2 = (b[i + 4] >> 1) - (unsigned)b[i + 12], z3 = b[i + 4] + (unsigned)(b[i + 12] >> 1);
        b[i] = (int16_t)(z0 + z3); b[i + 4] = (int16_t)(z1 + z2); b[i + 8] = (int16_t)(z1 - z2); b[i + 12] = (int16_t)(z0 - z3);
    }
    for (int i = 0; i < 4; i++) {
        const unsigned z0 = b[4 * i] + (unsigned)b[2 + 4 * i], z1 = b[4 * i] - (unsigned)b[2 + 4 * i];
        const unsigned z2 = (b[1 + 4 * i] >> 1) - (unsigned)b[3 + 4 * i], z3 = b[1 + 4 * i] + (unsigned)(b[3 + 4 * i] >> 1);
        dst[i + 0 * stride] = (uint8_t)clip_u8_i(dst[i + 0 * stride] + ((int)(z0 + z3) >> 6));
        dst[i + 1 * stride] = (uint8_t)clip_u8_i(dst[i + 1 * stride] + ((int)(z1 + z2) >> 6));
        dst[i + 2 * stride] = (uint8_t)clip_u8_i(dst[i + 2 * stride] + ((int)(z1 - z2) >> 6));
        dst[i + 3 * stride] = (uint8_t)clip_u8_i(dst[i + 3 * stride] + ((int)(z0 - z3) >> 6));
    }
    memset(b, 0, 16 * sizeof(int16_t));
}

/* one 8-point pass of ff_h264_idct8_add on in[0..7] (already sign-extended int16 values) -> out[0..7], sums mod 2^32 */
static void h264_idct8_1d(const int *in, unsigned *out)
{
    const unsigned a0 = in[0] + (unsigned)in[4], a2 = in[0] - (unsigned)in[4];
    const unsigned a4 = (in[2] >> 1) - (unsigned)in[6], a6 = (in[6] >> 1) + (unsigned)in[2];
    const unsigned b0 = a0 + a6, b2 = a2 + a4, b4 = a2 - a4, b6 = a0 - a6;
    const int a1 = (int)(-(unsigned)in[3] + (unsigned)in[5] - (unsigned)in[7] - (unsigned)(in[7] >> 1));
    const int a3 = (int)((unsigned)in[1] + (unsigned)in[7] - (unsigned)in[3] - (unsigned)(in[3] >> 1));
    const int a5 = (int)(-(unsigned)in[1] + (unsigned)in[7] + (unsigned)in[5] + (unsigned)(in[5] >> 1));
    const int a7 = (int)((unsigned)in[3] + (unsigned)in[5] + (unsigned)in[1] + (unsigned)(in[1] >> 1));
    const unsigned b1 = (unsigned)(a7 >> 2) + (unsigned)a1, b3 = (unsigned)a3 + (unsigned)(a5 >> 2);
    const unsigned b5 = (unsigned)(a3 >> 2) - (unsigned)a5, b7 = (unsigned)a7 - (unsigned)(a1 >> 2);
    out[0] = b0 + b7; out[7] = b0 - b7; out[1] = b2 + b5; out[6] = b2 - b5;
    out[2] = b4 + b3; out[5] = b4 - b3; out[3] = b6 + b1; out[4] = b6 - b1;
}

static void h264_idct8_add(uint8_t *dst, int16_t *b, ptrdiff_t stride)
{
    b[0] += 32;
    for (int i = 0; i < 8; i++) {
        int in[8]; unsigned out[8];
        for (int k = 0; k < 8; k++) in[k] = b[i + 8 * k];
        h264_idct8_1d(in, out);
        for (int k = 0; k < 8; k++) b[i + 8 * k] = (int16_t)out[k];
    }
    for (int i = 0; i < 8; i++) {
        int in[8]; unsigned out[8];
        for (int k = 0; k < 8; k++) in[k] = b[k + 8 * i];
        h264_idct8_1d(in, out);
        for (int k = 0; k < 8; k++) dst[i + k * stride] = (uint8_t)clip_u8_i(dst[i + k * stride] + ((int)out[k] >> 6));
    }
    memset(b, 0, 64 * sizeof(int16_t));
}

int orc_h264_idct(int kind, uint8_t *dst, int16_t *block, ptrdiff_t stride)
{
    if (kind == 0) { h264_idct4_add(dst, block, stride); return 0; }
    if (kind == 1) { h264_idct8_add(dst, block, stride); return 0; }
    if (kind == 2 || kind == 3) {
        const int n = kind == 2 ? 4 : 8, dc = (block[0] + 32) >> 6;
        block[0] = 0;
        for (int j = 0; j < n; j++)
            for (int i = 0; i < n; i++) dst[j * stride + i] = (uint8_t)clip_u8_i(dst[j * stride + i] + dc);
        return 0;
    }
    return -1;
}


/* ---- H.264 residual adds for 9 / 10 / 12 / 14 bit samples: h264idct_template.c:33-181 with dctcoef = int32_t, pixel = uint16_t
 * (h264dsp.c:66-158).  Same butterflies as the 8-bit functions on 32-bit values (the reference's SUINT sums: mod 2^32), clip to the depth. */
static int hb_px(int v, int maxv) { return v < 0 ? 0 : v > maxv ? maxv : v; }
static void hb_idct8_1d(const int32_t *s, int stride, int32_t *o)
{
    const uint32_t a0 = (uint32_t)s[0] + (uint32_t)s[4 * stride], a2 = (uint32_t)s[0] - (uint32_t)s[4 * stride];
    const uint32_t a4 = (uint32_t)(s[2 * stride] >> 1) - (uint32_t)s[6 * stride], a6 = (uint32_t)(s[6 * stride] >> 1) + (uint32_t)s[2 * stride];
    const uint32_t b0 = a0 + a6, b2 = a2 + a4, b4 = a2 - a4, b6 = a0 - a6;
    const int32_t a1 = (int32_t)(-(uint32_t)s[3 * stride] + (uint32_t)s[5 * stride] - (uint32_t)s[7 * stride] - (uint32_t)(s[7 * stride] >> 1));
    const int32_t a3 = (int32_t)((uint32_t)s[1 * stride] + (uint32_t)s[7 * stride] - (uint32_t)s[3 * stride] - (uint32_t)(s[3 * stride] >> 1));
    const int32_t a5 = (int32_t)(-(uint32_t)s[1 * stride] + (uint32_t)s[7 * stride] + (uint32_t)s[5 * stride] + (uint32_t)(s[5 * stride] >> 1));
    const int32_t a7 = (int32_t)((uint32_t)s[3 * stride] + (uint32_t)s[5 * stride] + (uint32_t)s[1 * stride] + (uint32_t)(s[1 * stride] >> 1));
    const uint32_t b1 = (uint32_t)(a7 >> 2) + (uint32_t)a1, b3 = (uint32_t)a3 + (uint32_t)(a5 >> 2);
    const uint32_t b5 = (uint32_t)(a3 >> 2) - (uint32_t)a5, b7 = (uint32_t)a7 - (uint32_t)(a1 >> 2);
    o[0] = (int32_t)(b0 + b7); o[7] = (int32_t)(b0 - b7); o[1] = (int32_t)(b2 + b5); o[6] = (int32_t)(b2 - b5);
    o[2] = (int32_t)(b4 + b3); o[5] = (int32_t)(b4 - b3); o[3] = (int32_t)(b6 + b1); o[4] = (int32_t)(b6 - b1);
}

int orc_h264_idct_hbd(int depth, int kind, uint8_t *dst8, int32_t *block, ptrdiff_t stride)
{
    if (kind < 0 || kind > 3 || (depth != 9 && depth != 10 && depth != 12 && depth != 14)) return -1;
    uint16_t *dst = (uint16_t *)dst8;
    const ptrdiff_t st = stride / 2;
    const int maxv = (1 << depth) - 1, n = (kind & 1) ? 8 : 4;
    if (kind >= 2) {
        const int dc = (block[0] + 32) >> 6;
        block[0] = 0;
        for (int j = 0; j < n; j++)
            for (int i = 0; i < n; i++) dst[j * st + i] = (uint16_t)hb_px(dst[j * st + i] + dc, maxv);
        return 0;
    }
    block[0] = (int32_t)((uint32_t)block[0] + 32u);
    if (kind == 0) {
        for (int i = 0; i < 4; i++) {
            const uint32_t z0 = (uint32_t)block[i] + (uint32_t)block[i + 8], z1 = (uint32_t)block[i] - (uint32_t)block[i + 8];
            const uint32_t z2 = (uint32_t)(block[i + 4] >> 1) - (uint32_t)block[i + 12], z3 = (uint32_t)block[i + 4] + (uint32_t)(block[i + 12] >> 1);
            block[i] = (int32_t)(z0 + z3); block[i + 4] = (int32_t)(z1 + z2); block[i + 8] = (int32_t)(z1 - z2); block[i + 12] = (int32_t)(z0 - z3);
        }
        for (int i = 0; i < 4; i++) {
            const uint32_t z0 = (uint32_t)block[4 * i] + (uint32_t)block[4 * i + 2], z1 = (uint32_t)block[4 * i] - (uint32_t)block[4 * i + 2];
            const uint32_t z2 = (uint32_t)(block[4 * i + 1] >> 1) - (uint32_t)block[4 * i + 3], z3 = (uint32_t)block[4 * i + 1] + (uint32_t)(block[4 * i + 3] >> 1);
            dst[i + 0 * st] = (uint16_t)hb_px(dst[i + 0 * st] + ((int32_t)(z0 + z3) >> 6), maxv);
            dst[i + 1 * st] = (uint16_t)hb_px(dst[i + 1 * st] + ((int32_t)(z1 + z2) >> 6), maxv);
            dst[i + 2 * st] = (uint16_t)hb_px(dst[i + 2 * st] + ((int32_t)(z1 - z2) >> 6), maxv);
            dst[i + 3 * st] = (uint16_t)hb_px(dst[i + 3 * st] + ((int32_t)(z0 - z3) >> 6), maxv);
        }
        memset(block, 0, 16 * sizeof(int32_t));
    } else {
        int32_t o[8];
        for (int i = 0; i < 8; i++) {
            hb_idct8_1d(block + i, 8, o);
            for (int r = 0; r < 8; r++) block[i + 8 * r] = o[r];
        }
        for (int i = 0; i < 8; i++) {
            hb_idct8_1d(block + 8 * i, 1, o);
            for (int r = 0; r < 8; r++) dst[i + r * st] = (uint16_t)hb_px(dst[i + r * st] + (o[r] >> 6), maxv);
        }
        memset(block, 0, 64 * sizeof(int32_t));
    }
    return 0;
}
