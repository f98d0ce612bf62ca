// fdct_dev.cuh — device pieces of the forward 8x8 DCTs of FDCTDSPContext (libavcodec/fdctdsp.c:27-45) shared by fdctdsp.cu (the transforms
// themselves) and mecmp_dct.cu (dct_sad / dct_max): a lane holds eight values of one row or column in registers.
#pragma once
#include "common.h"
// (everything here has internal linkage: each translation unit keeps its own copy)

static __device__ __forceinline__ int as_s16(int x) { return (int)(short)x; }
// DESCALE of jfdctint_template.c:70 on a product kept modulo 2^32, stored as int16
static __device__ __forceinline__ int round_s16(unsigned x, int n) { return as_s16(((int)x + (1 << (n - 1))) >> n); }

// lane r and lane k exchange so that afterwards v[k] of lane r is what v[r] of lane k was (lanes = the 8 of one block, row = lane & 7)
static __device__ __forceinline__ void transpose8(int (&v)[8], int row)
{
#pragma unroll
    for (int span = 1; span < 8; span <<= 1) {
        const bool upper = row & span;
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (!(k & span)) {
                const int give = upper ? v[k] : v[k + span];
                const int got = __shfl_xor_sync(0xffffffffu, give, span);
                if (upper) v[k] = got; else v[k + span] = got;
            }
    }
}

// Loeffler-Ligtenberg-Moschytz 1-D DCT with 13-bit constants (the "slow" integer DCT, jfdctint_template.c:173-340).  COL = second pass;
// P1 = PASS1_BITS (extra bits the row pass leaves: 4 for 8-bit samples, 1 for 10-bit), OUT = OUT_SHIFT (what the column pass removes: 4 / 2)
template <bool COL, int P1, int OUT>
static __device__ __forceinline__ void fdct_slow8(int (&v)[8])
{
    const int s0 = v[0] + v[7], s1 = v[1] + v[6], s2 = v[2] + v[5], s3 = v[3] + v[4];
    const int d0 = v[0] - v[7], d1 = v[1] - v[6], d2 = v[2] - v[5], d3 = v[3] - v[4];
    const int e0 = s0 + s3, e3 = s0 - s3, e1 = s1 + s2, e2 = s1 - s2;
    constexpr int SH = COL ? 13 + OUT : 13 - P1;
    v[0] = COL ? as_s16((e0 + e1 + (1 << (OUT - 1))) >> OUT) : as_s16((e0 + e1) * (1 << P1));
    v[4] = COL ? as_s16((e0 - e1 + (1 << (OUT - 1))) >> OUT) : as_s16((e0 - e1) * (1 << P1));
    const unsigned r = (unsigned)(e2 + e3) * 4433u;
    v[2] = round_s16(r + (unsigned)e3 * 6270u, SH);
    v[6] = round_s16(r - (unsigned)e2 * 15137u, SH);
    const unsigned q = (unsigned)(d3 + d1 + d2 + d0) * 9633u;
    const unsigned z1 = (unsigned)(d3 + d0) * 7373u, z2 = (unsigned)(d2 + d1) * 20995u;
    const unsigned z3 = q - (unsigned)(d3 + d1) * 16069u, z4 = q - (unsigned)(d2 + d0) * 3196u;
    v[7] = round_s16((unsigned)d3 * 2446u - z1 + z3, SH);
    v[5] = round_s16((unsigned)d2 * 16819u - z2 + z4, SH);
    v[3] = round_s16((unsigned)d1 * 25172u - z2 + z3, SH);
    v[1] = round_s16((unsigned)d0 * 12299u - z1 + z4, SH);
}

// Arai-Agui-Nakajima 1-D DCT with 8-bit constants, products shifted down without rounding and kept as int16 (the "fast" integer DCT)
static __device__ __forceinline__ int mul8(int x, int c) { return as_s16((x * c) >> 8); }
static __device__ __forceinline__ void fdct_fast8(int (&v)[8])
{
    const int s0 = v[0] + v[7], s1 = v[1] + v[6], s2 = v[2] + v[5], s3 = v[3] + v[4];
    const int d0 = v[0] - v[7], d1 = v[1] - v[6], d2 = v[2] - v[5], d3 = v[3] - v[4];
    const int e0 = s0 + s3, e3 = s0 - s3, e1 = s1 + s2, e2 = s1 - s2;
    v[0] = as_s16(e0 + e1); v[4] = as_s16(e0 - e1);
    const int r = mul8(e2 + e3, 181);
    v[2] = as_s16(e3 + r); v[6] = as_s16(e3 - r);
    const int a = d3 + d2, b = d2 + d1, c = d1 + d0;
    const int z5 = mul8(a - c, 98), z2 = mul8(a, 139) + z5, z4 = mul8(c, 334) + z5, z3 = mul8(b, 181);
    v[5] = as_s16(d0 - z3 + z2); v[3] = as_s16(d0 - z3 - z2); v[1] = as_s16(d0 + z3 + z4); v[7] = as_s16(d0 + z3 - z4);
}

// column pass of the 2-4-8 DCT (two interleaved fields: ff_fdct248_islow, jfdctint_template.c:347-412): the even part of the slow DCT on the
// sums of line pairs (-> rows 0 4 2 6) and again on their differences (-> rows 1 5 3 7)
template <int OUT>
static __device__ __forceinline__ void fdct248_slow_half(int a0, int a1, int a2, int a3, int &o0, int &o4, int &o2, int &o6)
{
    const int p = a0 + a3, q = a1 + a2, r = a1 - a2, s = a0 - a3;
    o0 = as_s16((p + q + (1 << (OUT - 1))) >> OUT);
    o4 = as_s16((p - q + (1 << (OUT - 1))) >> OUT);
    const unsigned z = (unsigned)(r + s) * 4433u;
    o2 = round_s16(z + (unsigned)s * 6270u, 13 + OUT);
    o6 = round_s16(z - (unsigned)r * 15137u, 13 + OUT);
}
template <int OUT>
static __device__ __forceinline__ void fdct248_slow8(int (&v)[8])
{
    const int s0 = v[0] + v[1], s1 = v[2] + v[3], s2 = v[4] + v[5], s3 = v[6] + v[7];
    const int d0 = v[0] - v[1], d1 = v[2] - v[3], d2 = v[4] - v[5], d3 = v[6] - v[7];
    fdct248_slow_half<OUT>(s0, s1, s2, s3, v[0], v[4], v[2], v[6]);
    fdct248_slow_half<OUT>(d0, d1, d2, d3, v[1], v[5], v[3], v[7]);
}
// the same for the fast DCT (ff_fdct_ifast248, jfdctfst.c:286-343)
static __device__ __forceinline__ void fdct248_fast_half(int a0, int a1, int a2, int a3, int &o0, int &o4, int &o2, int &o6)
{
    const int p = a0 + a3, q = a1 + a2, r = a1 - a2, s = a0 - a3;
    o0 = as_s16(p + q); o4 = as_s16(p - q);
    const int z = mul8(r + s, 181);
    o2 = as_s16(s + z); o6 = as_s16(s - z);
}
static __device__ __forceinline__ void fdct248_fast8(int (&v)[8])
{
    const int s0 = v[0] + v[1], s1 = v[2] + v[3], s2 = v[4] + v[5], s3 = v[6] + v[7];
    const int d0 = v[0] - v[1], d1 = v[2] - v[3], d2 = v[4] - v[5], d3 = v[6] - v[7];
    fdct248_fast_half(s0, s1, s2, s3, v[0], v[4], v[2], v[6]);
    fdct248_fast_half(d0, d1, d2, d3, v[1], v[5], v[3], v[7]);
}
