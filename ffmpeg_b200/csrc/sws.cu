// sws.cu — libswscale yuv420p -> rgb24 on sm_100a: kernels + the C ABI (include/b200dsp.h, "libswscale" section).
//
// Reference semantics reproduced bit-for-bit (see oracle/sws_oracle.c for the CPU restatement used as checker):
//   unscaled LUT converter  yuv2rgb_c_24_rgb          libswscale/yuv2rgb.c:137-236,530   -> sws_unscaled_kernel
//   horizontal FIR          hScale8To15_c             libswscale/swscale.c:128-142       -> sws_hscale_kernel
//   vertical FIR + writers  yuv2rgb24_{X,2,1}_c       libswscale/output.c:1789-1939      -> sws_vscale_rgb24_kernel
//   full-chroma writers     yuv2rgb24_full_{X,2,1}_c  libswscale/output.c:2161-2310      -> sws_vscale_rgb24_full_kernel
// The reference's pointer LUTs (table_rV/gU/gV/bU into a clipped luma ramp) are evaluated in closed form, which is
// exact because the ramp is clip_u8((yb0 + k*cy) >> 16) for every k (yuv2rgb.c:901-914).
//
// Data layout in HBM: planar u8 Y/U/V exactly as the caller's AVFrame planes (any stride), packed rgb24 rows out.
// The kernels are HBM-bound byte movers: each thread owns 16 horizontally adjacent pixels so that luma moves as one
// 128-bit load, chroma as 64-bit loads and rgb24 as three 128-bit stores (48 B = 16 px).
#include "common.h"
#include "sws_plan.h"
#include <vector>
#include <cstring>
#include <new>

// ------------------------------------------------------------------------------------------------ device helpers
struct SwsDevTables {            // device copies of the vertical banks and the per-line writer choice
    const int16_t *vLum; const int32_t *vLumPos; int vLumSize;
    const int16_t *vChr; const int32_t *vChrPos; int vChrSize;
    const int32_t *rowMode;
    const int16_t *hLum; const int32_t *hLumPos; int hLumSize;
    const int16_t *hChr; const int32_t *hChrPos; int hChrSize;
};

struct SwsFrameArgs {
    const uint8_t *y, *u, *v;         // u8 source planes (or int16 planes reinterpret_cast for the scaled path)
    long long ys, us, vs;             // strides in BYTES (may be negative for u8 planes)
    long long yfs, ufs, vfs;          // frame strides in bytes
    uint8_t *dst; long long ds, dfs;
    int srcH, chrSrcH, dstW, dstH, chrDstW;
};

__device__ __forceinline__ int clamp_u8(int v) { return min(max(v, 0), 255); }

// colour constants per chroma sample: A_c such that channel = clip_u8((A_c + Y*cy) >> 16)
struct ChromaBase { int r, g, b; };
__device__ __forceinline__ ChromaBase chroma_base(const SwsColorConst &c, int U, int V)
{
    const int u8 = clamp_u8(U), v8 = clamp_u8(V);
    ChromaBase o;
    o.r = c.yb0 + (c.baseR + ((v8 * c.crv) >> 16)) * c.cy;
    o.g = c.yb0 + (c.baseG + ((u8 * c.cgu) >> 16) + ((v8 * c.cgv) >> 16)) * c.cy;
    o.b = c.yb0 + (c.baseB + ((u8 * c.cbu) >> 16)) * c.cy;
    return o;
}

// writes 2 pixels (6 bytes) worth of channels into out[0..5]
__device__ __forceinline__ void pair_rgb(const SwsColorConst &c, const ChromaBase &cb, int Y1, int Y2, int *out)
{
    const int t1 = Y1 * c.cy, t2 = Y2 * c.cy;
    out[0] = clamp_u8((cb.r + t1) >> 16); out[1] = clamp_u8((cb.g + t1) >> 16); out[2] = clamp_u8((cb.b + t1) >> 16);
    out[3] = clamp_u8((cb.r + t2) >> 16); out[4] = clamp_u8((cb.g + t2) >> 16); out[5] = clamp_u8((cb.b + t2) >> 16);
}

__device__ __forceinline__ unsigned pack4(int a, int b, int c, int d)
{
    return (unsigned)a | ((unsigned)b << 8) | ((unsigned)c << 16) | ((unsigned)d << 24);
}

// store 16 px (48 channel values) starting at dst; vector path when `vec`, else byte stores of `npx` pixels
__device__ __forceinline__ void store_px16(uint8_t *dst, const int *ch, bool vec, int npx)
{
    if (vec) {
        uint4 *d4 = reinterpret_cast<uint4 *>(dst);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            uint4 v;
            v.x = pack4(ch[16 * k + 0], ch[16 * k + 1], ch[16 * k + 2], ch[16 * k + 3]);
            v.y = pack4(ch[16 * k + 4], ch[16 * k + 5], ch[16 * k + 6], ch[16 * k + 7]);
            v.z = pack4(ch[16 * k + 8], ch[16 * k + 9], ch[16 * k + 10], ch[16 * k + 11]);
            v.w = pack4(ch[16 * k + 12], ch[16 * k + 13], ch[16 * k + 14], ch[16 * k + 15]);
            d4[k] = v;
        }
    } else {
        for (int i = 0; i < npx * 3; i++) dst[i] = (uint8_t)ch[i];
    }
}

__device__ __forceinline__ void unpack16(const uint4 &q, int *o)
{
    const unsigned w[4] = { q.x, q.y, q.z, q.w };
#pragma unroll
    for (int k = 0; k < 4; k++) {
        o[4 * k + 0] = w[k] & 0xff; o[4 * k + 1] = (w[k] >> 8) & 0xff;
        o[4 * k + 2] = (w[k] >> 16) & 0xff; o[4 * k + 3] = w[k] >> 24;
    }
}
__device__ __forceinline__ void unpack8(const uint2 &q, int *o)
{
    o[0] = q.x & 0xff; o[1] = (q.x >> 8) & 0xff; o[2] = (q.x >> 16) & 0xff; o[3] = q.x >> 24;
    o[4] = q.y & 0xff; o[5] = (q.y >> 8) & 0xff; o[6] = (q.y >> 16) & 0xff; o[7] = q.y >> 24;
}

// 16 consecutive u8 samples (or fewer at the right edge: missing ones read as 0)
__device__ __forceinline__ void load_u8x16(const uint8_t *p, int avail, bool vec, int *o)
{
    if (vec && avail >= 16) {
        unpack16(__ldg(reinterpret_cast<const uint4 *>(p)), o);
    } else {
#pragma unroll
        for (int i = 0; i < 16; i++) o[i] = i < avail ? (int)__ldg(p + i) : 0;
    }
}
__device__ __forceinline__ void load_u8x8(const uint8_t *p, int avail, bool vec, int *o)
{
    if (vec && avail >= 8) {
        unpack8(__ldg(reinterpret_cast<const uint2 *>(p)), o);
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++) o[i] = i < avail ? (int)__ldg(p + i) : 0;
    }
}
__device__ __forceinline__ void load_s16(const int16_t *p, int n, int avail, int *o)
{
    for (int i = 0; i < n; i++) o[i] = i < avail ? (int)__ldg(p + i) : 0;
}

// ------------------------------------------------------------------------------------------------ kernel: unscaled LUT path
// One thread = 16 px x 2 lines (one chroma line).  Grid: x over 16-px groups, y over line pairs, z over frames.
__global__ void __launch_bounds__(256)
sws_unscaled_kernel(SwsFrameArgs a, SwsColorConst c, int wpix /* pixels the reference writes per line */, int vecOK)
{
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
    const int row = blockIdx.y * 2;
    if (x0 >= wpix) return;
    const long long f = blockIdx.z;
    const uint8_t *py = a.y + f * a.yfs + (long long)row * a.ys + x0;
    const uint8_t *pu = a.u + f * a.ufs + (long long)(row >> 1) * a.us + (x0 >> 1);
    const uint8_t *pv = a.v + f * a.vfs + (long long)(row >> 1) * a.vs + (x0 >> 1);
    uint8_t *d0 = a.dst + f * a.dfs + (long long)row * a.ds + (long long)x0 * 3;
    const int npx = min(16, wpix - x0);            // even by construction
    const bool vec = vecOK && npx == 16;
    int U[8], V[8], Y0[16], Y1[16];
    load_u8x8(pu, npx >> 1, vec, U);
    load_u8x8(pv, npx >> 1, vec, V);
    load_u8x16(py, npx, vec, Y0);
    load_u8x16(py + a.ys, npx, vec, Y1);
    int ch0[48], ch1[48];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const ChromaBase cb = chroma_base(c, U[i], V[i]);
        pair_rgb(c, cb, Y0[2 * i], Y0[2 * i + 1], ch0 + 6 * i);
        pair_rgb(c, cb, Y1[2 * i], Y1[2 * i + 1], ch1 + 6 * i);
    }
    store_px16(d0, ch0, vec, npx);
    store_px16(d0 + a.ds, ch1, vec, npx);
}

// ------------------------------------------------------------------------------------------------ kernel: horizontal FIR
// dst[i] = min((sum_j src[pos[i]+j] * coef[i*fs+j]) >> 7, 32767); one thread per output sample, grid y = lines, z = frames.
__global__ void __launch_bounds__(256)
sws_hscale_kernel(const uint8_t *src, long long sstride, long long sfs, int16_t *dst, int dstW, long long dfs,
                  const int16_t *coef, const int32_t *pos, int fs)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= dstW) return;
    const uint8_t *s = src + (long long)blockIdx.z * sfs + (long long)blockIdx.y * sstride + __ldg(pos + i);
    const int16_t *k = coef + (long long)i * fs;
    int acc = 0;
    for (int j = 0; j < fs; j++) acc += (int)__ldg(s + j) * (int)__ldg(k + j);
    acc >>= 7;
    dst[(long long)blockIdx.z * dfs + (long long)blockIdx.y * dstW + i] = (int16_t)min(acc, 32767);
}

// ------------------------------------------------------------------------------------------------ kernel: vertical FIR + rgb24 writer
// SRC8: the horizontal pass is the identity, so taps are read straight from the u8 planes and widened (<<7) here,
//       exactly what hScale8To15_c produces for coefficient 1<<14.  Otherwise taps come from int16 line planes.
// One thread = 16 px of one output line.  Grid: x over 16-px groups, y over lines, z over frames.
template <bool SRC8>
__global__ void __launch_bounds__(256)
sws_vscale_rgb24_kernel(SwsFrameArgs a, SwsDevTables t, SwsColorConst c, int vecOK)
{
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
    const int dy = blockIdx.y;
    if (x0 >= a.dstW) return;
    const long long f = blockIdx.z;
    const int npx = min(16, a.dstW - x0);          // dstW is even on this path
    const int nch = (npx + 1) >> 1;
    const bool vec = vecOK && npx == 16;
    const int lfs = t.vLumSize, cfs = t.vChrSize;
    const int16_t *lf = t.vLum + (long long)dy * lfs, *cf = t.vChr + (long long)dy * cfs;
    const int firstLum = max(1 - lfs, __ldg(t.vLumPos + dy));
    const int firstChr = max(1 - cfs, __ldg(t.vChrPos + dy));
    const int mode = __ldg(t.rowMode + 4 * dy), yalpha = __ldg(t.rowMode + 4 * dy + 1), uvalpha = __ldg(t.rowMode + 4 * dy + 2);
    const uint8_t *ybase = a.y + f * a.yfs, *ubase = a.u + f * a.ufs, *vbase = a.v + f * a.vfs;

    auto lum = [&](int line, int *o) {
        line = min(max(line, 0), a.srcH - 1);
        if (SRC8) {
            load_u8x16(ybase + line * a.ys + x0, npx, vec, o);
#pragma unroll
            for (int i = 0; i < 16; i++) o[i] <<= 7;
        } else {
            load_s16(reinterpret_cast<const int16_t *>(ybase + line * a.ys) + x0, 16, npx, o);
        }
    };
    auto chr = [&](const uint8_t *base, long long stride, int line, int *o) {
        line = min(max(line, 0), a.chrSrcH - 1);
        if (SRC8) {
            load_u8x8(base + line * stride + (x0 >> 1), nch, vec, o);
#pragma unroll
            for (int i = 0; i < 8; i++) o[i] <<= 7;
        } else {
            load_s16(reinterpret_cast<const int16_t *>(base + line * stride) + (x0 >> 1), 8, nch, o);
        }
    };

    int Y[16], U[8], V[8], s[16];
    if (mode == 1) {                                   // yuv2rgb_1_c_template, output.c:1883-1939
        lum(firstLum, s);
#pragma unroll
        for (int i = 0; i < 16; i++) Y[i] = (s[i] + 64) >> 7;
        if (!uvalpha) {
            chr(ubase, a.us, firstChr, s);
#pragma unroll
            for (int i = 0; i < 8; i++) U[i] = (s[i] + 64) >> 7;
            chr(vbase, a.vs, firstChr, s);
#pragma unroll
            for (int i = 0; i < 8; i++) V[i] = (s[i] + 64) >> 7;
        } else {
            const int a1 = 4096 - uvalpha;
            int s1[8];
            chr(ubase, a.us, firstChr, s); chr(ubase, a.us, firstChr + 1, s1);
#pragma unroll
            for (int i = 0; i < 8; i++) U[i] = (s[i] * a1 + s1[i] * uvalpha + (128 << 11)) >> 19;
            chr(vbase, a.vs, firstChr, s); chr(vbase, a.vs, firstChr + 1, s1);
#pragma unroll
            for (int i = 0; i < 8; i++) V[i] = (s[i] * a1 + s1[i] * uvalpha + (128 << 11)) >> 19;
        }
    } else if (mode == 2) {                            // yuv2rgb_2_c_template, output.c:1843-1880
        const int ya1 = 4096 - yalpha, ua1 = 4096 - uvalpha;
        int s1[16];
        lum(firstLum, s); lum(firstLum + 1, s1);
#pragma unroll
        for (int i = 0; i < 16; i++) Y[i] = (s[i] * ya1 + s1[i] * yalpha) >> 19;
        chr(ubase, a.us, firstChr, s); chr(ubase, a.us, firstChr + 1, s1);
#pragma unroll
        for (int i = 0; i < 8; i++) U[i] = (s[i] * ua1 + s1[i] * uvalpha) >> 19;
        chr(vbase, a.vs, firstChr, s); chr(vbase, a.vs, firstChr + 1, s1);
#pragma unroll
        for (int i = 0; i < 8; i++) V[i] = (s[i] * ua1 + s1[i] * uvalpha) >> 19;
    } else {                                           // yuv2rgb_X_c_template, output.c:1789-1840 (unsigned wrap-around sums)
        unsigned ay[16], au[8], av[8];
#pragma unroll
        for (int i = 0; i < 16; i++) ay[i] = 1u << 18;
#pragma unroll
        for (int i = 0; i < 8; i++) au[i] = av[i] = 1u << 18;
        for (int j = 0; j < lfs; j++) {
            const unsigned k = (unsigned)(int)__ldg(lf + j);
            lum(firstLum + j, s);
#pragma unroll
            for (int i = 0; i < 16; i++) ay[i] += (unsigned)s[i] * k;
        }
        for (int j = 0; j < cfs; j++) {
            const unsigned k = (unsigned)(int)__ldg(cf + j);
            chr(ubase, a.us, firstChr + j, s);
#pragma unroll
            for (int i = 0; i < 8; i++) au[i] += (unsigned)s[i] * k;
            chr(vbase, a.vs, firstChr + j, s);
#pragma unroll
            for (int i = 0; i < 8; i++) av[i] += (unsigned)s[i] * k;
        }
#pragma unroll
        for (int i = 0; i < 16; i++) Y[i] = (int)ay[i] >> 19;
#pragma unroll
        for (int i = 0; i < 8; i++) { U[i] = (int)au[i] >> 19; V[i] = (int)av[i] >> 19; }
    }

    int ch[48];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const ChromaBase cb = chroma_base(c, U[i], V[i]);
        pair_rgb(c, cb, Y[2 * i], Y[2 * i + 1], ch + 6 * i);
    }
    store_px16(a.dst + f * a.dfs + (long long)dy * a.ds + (long long)x0 * 3, ch, vec, npx);
}

// ------------------------------------------------------------------------------------------------ kernel: full-chroma writer
// chrDstW == dstW (SWS_FULL_CHR_H_INT, forced for odd widths).  One thread per pixel; taps always from int16 planes.
__global__ void __launch_bounds__(256)
sws_vscale_rgb24_full_kernel(SwsFrameArgs a, SwsDevTables t, SwsColorConst c)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int dy = blockIdx.y;
    if (x >= a.dstW) return;
    const long long f = blockIdx.z;
    const int lfs = t.vLumSize, cfs = t.vChrSize;
    const int16_t *lf = t.vLum + (long long)dy * lfs, *cf = t.vChr + (long long)dy * cfs;
    const int firstLum = max(1 - lfs, __ldg(t.vLumPos + dy));
    const int firstChr = max(1 - cfs, __ldg(t.vChrPos + dy));
    const int mode = __ldg(t.rowMode + 4 * dy), yalpha = __ldg(t.rowMode + 4 * dy + 1), uvalpha = __ldg(t.rowMode + 4 * dy + 2);
    auto L = [&](int line) { line = min(max(line, 0), a.srcH - 1);
        return (int)__ldg(reinterpret_cast<const int16_t *>(a.y + f * a.yfs + line * a.ys) + x); };
    auto Uu = [&](int line) { line = min(max(line, 0), a.chrSrcH - 1);
        return (int)__ldg(reinterpret_cast<const int16_t *>(a.u + f * a.ufs + line * a.us) + x); };
    auto Vv = [&](int line) { line = min(max(line, 0), a.chrSrcH - 1);
        return (int)__ldg(reinterpret_cast<const int16_t *>(a.v + f * a.vfs + line * a.vs) + x); };
    int Y, U, V;
    if (mode == 1) {                                   // output.c:2257-2310
        Y = L(firstLum) * 4;
        if (!uvalpha) { U = (Uu(firstChr) - (128 << 7)) * 4; V = (Vv(firstChr) - (128 << 7)) * 4; }
        else {
            const int a1 = 4096 - uvalpha;
            U = (Uu(firstChr) * a1 + Uu(firstChr + 1) * uvalpha - (128 << 19)) >> 10;
            V = (Vv(firstChr) * a1 + Vv(firstChr + 1) * uvalpha - (128 << 19)) >> 10;
        }
    } else if (mode == 2) {                            // output.c:2211-2254
        const int ya1 = 4096 - yalpha, ua1 = 4096 - uvalpha;
        Y = (L(firstLum) * ya1 + L(firstLum + 1) * yalpha) >> 10;
        U = (Uu(firstChr) * ua1 + Uu(firstChr + 1) * uvalpha - (128 << 19)) >> 10;
        V = (Vv(firstChr) * ua1 + Vv(firstChr + 1) * uvalpha - (128 << 19)) >> 10;
    } else {                                           // output.c:2161-2208
        unsigned ay = 1u << 9, au = (unsigned)((1 << 9) - (128 << 19)), av = au;
        for (int j = 0; j < lfs; j++) ay += (unsigned)L(firstLum + j) * (unsigned)(int)__ldg(lf + j);
        for (int j = 0; j < cfs; j++) {
            const unsigned k = (unsigned)(int)__ldg(cf + j);
            au += (unsigned)Uu(firstChr + j) * k;
            av += (unsigned)Vv(firstChr + j) * k;
        }
        Y = (int)ay >> 10; U = (int)au >> 10; V = (int)av >> 10;
    }
    // yuv2rgb_write_full, output.c:1998-2030
    const unsigned yy = (unsigned)(Y - c.y_offset) * (unsigned)c.y_coeff + (1u << 21);
    int R = (int)(yy + (unsigned)V * (unsigned)c.v2r);
    int G = (int)(yy + (unsigned)V * (unsigned)c.v2g + (unsigned)U * (unsigned)c.u2g);
    int B = (int)(yy + (unsigned)U * (unsigned)c.u2b);
    if ((R | G | B) & 0xC0000000) {
        R = (R & 0xC0000000) ? ((~R) >> 31 & 0x3FFFFFFF) : R;
        G = (G & 0xC0000000) ? ((~G) >> 31 & 0x3FFFFFFF) : G;
        B = (B & 0xC0000000) ? ((~B) >> 31 & 0x3FFFFFFF) : B;
    }
    uint8_t *d = a.dst + f * a.dfs + (long long)dy * a.ds + (long long)x * 3;
    d[0] = (uint8_t)(R >> 22); d[1] = (uint8_t)(G >> 22); d[2] = (uint8_t)(B >> 22);
}

// ------------------------------------------------------------------------------------------------ host side
struct B200SwsContext {
    B200Device *dev = nullptr;
    SwsPlan plan;
    void *tables = nullptr;          // one device allocation holding all banks
    SwsDevTables dt{};
    bool h_identity = false;         // both horizontal banks are the identity -> SRC8 kernels
    // intermediate int16 line planes for the scaled path (grown on demand, per batch)
    void *mid = nullptr; size_t mid_bytes = 0;
};

static int upload_tables(B200SwsContext *c)
{
    const SwsPlan &p = c->plan;
    if (c->tables) { cudaFree(c->tables); c->tables = nullptr; }
    if (p.unscaled_lut) return 0;
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    size_t off = 0;
    size_t o_vl = off;  off += al(p.vLum.coef.size() * 2);
    size_t o_vlp = off; off += al(p.vLum.pos.size() * 4);
    size_t o_vc = off;  off += al(p.vChr.coef.size() * 2);
    size_t o_vcp = off; off += al(p.vChr.pos.size() * 4);
    size_t o_rm = off;  off += al(p.rowMode.size() * 4);
    size_t o_hl = off;  off += al(p.hLum.coef.size() * 2);
    size_t o_hlp = off; off += al(p.hLum.pos.size() * 4);
    size_t o_hc = off;  off += al(p.hChr.coef.size() * 2);
    size_t o_hcp = off; off += al(p.hChr.pos.size() * 4);
    std::vector<uint8_t> host(off, 0);
    memcpy(&host[o_vl], p.vLum.coef.data(), p.vLum.coef.size() * 2);
    memcpy(&host[o_vlp], p.vLum.pos.data(), p.vLum.pos.size() * 4);
    memcpy(&host[o_vc], p.vChr.coef.data(), p.vChr.coef.size() * 2);
    memcpy(&host[o_vcp], p.vChr.pos.data(), p.vChr.pos.size() * 4);
    memcpy(&host[o_rm], p.rowMode.data(), p.rowMode.size() * 4);
    memcpy(&host[o_hl], p.hLum.coef.data(), p.hLum.coef.size() * 2);
    memcpy(&host[o_hlp], p.hLum.pos.data(), p.hLum.pos.size() * 4);
    memcpy(&host[o_hc], p.hChr.coef.data(), p.hChr.coef.size() * 2);
    memcpy(&host[o_hcp], p.hChr.pos.data(), p.hChr.pos.size() * 4);
    B200_CUDA_OK(cudaMalloc(&c->tables, off));
    B200_CUDA_OK(cudaMemcpy(c->tables, host.data(), off, cudaMemcpyHostToDevice));
    uint8_t *b = (uint8_t *)c->tables;
    c->dt.vLum = (const int16_t *)(b + o_vl); c->dt.vLumPos = (const int32_t *)(b + o_vlp); c->dt.vLumSize = p.vLum.size;
    c->dt.vChr = (const int16_t *)(b + o_vc); c->dt.vChrPos = (const int32_t *)(b + o_vcp); c->dt.vChrSize = p.vChr.size;
    c->dt.rowMode = (const int32_t *)(b + o_rm);
    c->dt.hLum = (const int16_t *)(b + o_hl); c->dt.hLumPos = (const int32_t *)(b + o_hlp); c->dt.hLumSize = p.hLum.size;
    c->dt.hChr = (const int16_t *)(b + o_hc); c->dt.hChrPos = (const int32_t *)(b + o_hcp); c->dt.hChrSize = p.hChr.size;
    c->h_identity = p.chrDstHSub == 1 && p.hLum.identity() && p.hChr.identity();
    return 0;
}

B200_API B200SwsContext *b200_sws_getContext(B200Device *dev, int srcW, int srcH, int srcFormat,
                                             int dstW, int dstH, int dstFormat, int flags)
{
    if (!dev) { b200_set_error("b200_sws_getContext: no device"); return nullptr; }
    if (srcFormat != B200_PIX_FMT_YUV420P || dstFormat != B200_PIX_FMT_RGB24) {
        b200_set_error("b200_sws_getContext: only yuv420p -> rgb24 is implemented");
        return nullptr;
    }
    B200SwsContext *c = new (std::nothrow) B200SwsContext();
    if (!c) return nullptr;
    c->dev = dev;
    int ret = sws_plan_build(c->plan, srcW, srcH, dstW, dstH, flags);
    if (ret < 0) { b200_set_error("b200_sws_getContext: unsupported configuration (%d)", ret); delete c; return nullptr; }
    cudaSetDevice(dev->ordinal);
    if (upload_tables(c) < 0) { delete c; return nullptr; }
    return c;
}

B200_API void b200_sws_freeContext(B200SwsContext *c)
{
    if (!c) return;
    cudaSetDevice(c->dev->ordinal);
    cudaStreamSynchronize(c->dev->stream);
    if (c->tables) cudaFree(c->tables);
    if (c->mid) cudaFree(c->mid);
    delete c;
}

B200_API int b200_sws_setColorspaceDetails(B200SwsContext *c, const int inv_table[4], int srcRange,
                                           const int table[4], int dstRange, int brightness, int contrast, int saturation)
{
    (void)table; (void)dstRange;          // dst is RGB: range_override_needed(), utils.c:844-880
    if (!c || !inv_table) return B200_EINVAL;
    return sws_plan_colorspace(c->plan, inv_table, srcRange, brightness, contrast, saturation);
}

B200_API int b200_sws_info(const B200SwsContext *c, int *o)
{
    if (!c || !o) return B200_EINVAL;
    const SwsPlan &p = c->plan;
    o[0] = p.hLum.size; o[1] = p.hChr.size; o[2] = p.vLum.size; o[3] = p.vChr.size;
    o[4] = p.chrSrcW; o[5] = p.chrSrcH; o[6] = p.chrDstW; o[7] = p.chrDstH;
    o[8] = p.unscaled_lut; o[9] = 1; o[10] = 1; o[11] = p.chrDstHSub; o[12] = 0;
    o[13] = p.dstW; o[14] = p.dstH; o[15] = 0;
    return 0;
}

B200_API int b200_sws_get_filter(const B200SwsContext *c, int which, int16_t *filter, int32_t *pos, int cap)
{
    if (!c) return B200_EINVAL;
    const SwsFilterBank &b = which == 0 ? c->plan.hLum : which == 1 ? c->plan.hChr : which == 2 ? c->plan.vLum : c->plan.vChr;
    int n = b.n < cap ? b.n : cap;
    if (filter) memcpy(filter, b.coef.data(), (size_t)n * b.size * 2);
    if (pos) memcpy(pos, b.pos.data(), (size_t)n * 4);
    return n;
}

B200_API int b200_sws_plan_probe(int srcW, int srcH, int dstW, int dstH, int flags, int which,
                                 int16_t *filter, int32_t *pos, int cap, int *o)
{
    SwsPlan p;
    int ret = sws_plan_build(p, srcW, srcH, dstW, dstH, flags);
    if (ret < 0) return ret;
    if (o) {
        o[0] = p.hLum.size; o[1] = p.hChr.size; o[2] = p.vLum.size; o[3] = p.vChr.size;
        o[4] = p.chrSrcW; o[5] = p.chrSrcH; o[6] = p.chrDstW; o[7] = p.chrDstH;
        o[8] = p.unscaled_lut; o[9] = 1; o[10] = 1; o[11] = p.chrDstHSub; o[12] = 0;
        o[13] = p.dstW; o[14] = p.dstH; o[15] = 0;
    }
    const SwsFilterBank &b = which == 0 ? p.hLum : which == 1 ? p.hChr : which == 2 ? p.vLum : p.vChr;
    int n = b.n < cap ? b.n : cap;
    if (filter && n) memcpy(filter, b.coef.data(), (size_t)n * b.size * 2);
    if (pos && n) memcpy(pos, b.pos.data(), (size_t)n * 4);
    return n;
}

static bool aligned16(const void *p, long long stride, long long fstride)
{
    return (((uintptr_t)p) & 15) == 0 && (stride & 15) == 0 && (fstride & 15) == 0;
}
static bool aligned8(const void *p, long long stride, long long fstride)
{
    return (((uintptr_t)p) & 7) == 0 && (stride & 7) == 0 && (fstride & 7) == 0;
}

// enqueue the conversion of nframes frames on `stream`
static int launch_batch(B200SwsContext *c, cudaStream_t stream, const uint8_t *const src[3], const long long sstr[3],
                        const long long sfs[3], uint8_t *dst, long long ds, long long dfs, int nframes)
{
    const SwsPlan &p = c->plan;
    if (nframes <= 0) return 0;
    SwsFrameArgs a{};
    a.y = src[0]; a.u = src[1]; a.v = src[2];
    a.ys = sstr[0]; a.us = sstr[1]; a.vs = sstr[2];
    a.yfs = sfs[0]; a.ufs = sfs[1]; a.vfs = sfs[2];
    a.dst = dst; a.ds = ds; a.dfs = dfs;
    a.srcH = p.srcH; a.chrSrcH = p.chrSrcH; a.dstW = p.dstW; a.dstH = p.dstH; a.chrDstW = p.chrDstW;
    const int vecSrc = aligned16(src[0], sstr[0], sfs[0]) && aligned8(src[1], sstr[1], sfs[1]) && aligned8(src[2], sstr[2], sfs[2]);
    const int vecOK = vecSrc && aligned16(dst, ds, dfs);
    for (int f0 = 0; f0 < nframes; f0 += 65535) {            // gridDim.z limit
        const int nf = nframes - f0 < 65535 ? nframes - f0 : 65535;
        SwsFrameArgs b = a;
        b.y += (long long)f0 * a.yfs; b.u += (long long)f0 * a.ufs; b.v += (long long)f0 * a.vfs; b.dst += (long long)f0 * a.dfs;
        if (p.unscaled_lut) {
            const int wpix = ((p.dstW >> 3) << 3) + (p.dstW & 4) + (p.dstW & 2);
            if (wpix == 0) continue;
            dim3 block(128), grid(b200_ceil_div(b200_ceil_div(wpix, 16), 128), p.dstH / 2, nf);
            sws_unscaled_kernel<<<grid, block, 0, stream>>>(b, p.color, wpix, vecOK);
            B200_LAUNCHED();
        } else if (c->h_identity) {
            dim3 block(128), grid(b200_ceil_div(b200_ceil_div(p.dstW, 16), 128), p.dstH, nf);
            sws_vscale_rgb24_kernel<true><<<grid, block, 0, stream>>>(b, c->dt, p.color, vecOK);
            B200_LAUNCHED();
        } else {
            // scaled path: horizontal pass into int16 line planes, then the vertical pass
            const size_t lumPlane = (size_t)p.srcH * p.dstW * 2, chrPlane = (size_t)p.chrSrcH * p.chrDstW * 2;
            const size_t perFrame = ((lumPlane + 2 * chrPlane) + 255) & ~(size_t)255;
            const size_t need = perFrame * nf;
            if (c->mid_bytes < need) {
                if (c->mid) { cudaStreamSynchronize(stream); cudaFree(c->mid); c->mid = nullptr; c->mid_bytes = 0; }
                B200_CUDA_OK(cudaMalloc(&c->mid, need));
                c->mid_bytes = need;
            }
            int16_t *mY = (int16_t *)c->mid;
            int16_t *mU = (int16_t *)((uint8_t *)c->mid + lumPlane);
            int16_t *mV = (int16_t *)((uint8_t *)c->mid + lumPlane + chrPlane);
            const long long mfs = (long long)(perFrame / 2);
            {
                dim3 block(256), grid(b200_ceil_div(p.dstW, 256), p.srcH, nf);
                sws_hscale_kernel<<<grid, block, 0, stream>>>(b.y, b.ys, b.yfs, mY, p.dstW, mfs, c->dt.hLum, c->dt.hLumPos, c->dt.hLumSize);
                B200_LAUNCHED();
                dim3 gridc(b200_ceil_div(p.chrDstW, 256), p.chrSrcH, nf);
                sws_hscale_kernel<<<gridc, block, 0, stream>>>(b.u, b.us, b.ufs, mU, p.chrDstW, mfs, c->dt.hChr, c->dt.hChrPos, c->dt.hChrSize);
                B200_LAUNCHED();
                sws_hscale_kernel<<<gridc, block, 0, stream>>>(b.v, b.vs, b.vfs, mV, p.chrDstW, mfs, c->dt.hChr, c->dt.hChrPos, c->dt.hChrSize);
                B200_LAUNCHED();
            }
            SwsFrameArgs m = b;
            m.y = (const uint8_t *)mY; m.u = (const uint8_t *)mU; m.v = (const uint8_t *)mV;
            m.ys = (long long)p.dstW * 2; m.us = m.vs = (long long)p.chrDstW * 2;
            m.yfs = m.ufs = m.vfs = (long long)perFrame;
            if (p.chrDstHSub) {
                dim3 block(128), grid(b200_ceil_div(b200_ceil_div(p.dstW, 16), 128), p.dstH, nf);
                sws_vscale_rgb24_kernel<false><<<grid, block, 0, stream>>>(m, c->dt, p.color, aligned16(dst, ds, dfs));
            } else {
                dim3 block(256), grid(b200_ceil_div(p.dstW, 256), p.dstH, nf);
                sws_vscale_rgb24_full_kernel<<<grid, block, 0, stream>>>(m, c->dt, p.color);
            }
            B200_LAUNCHED();
        }
    }
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

B200_API int b200_sws_scale_batch_device(B200SwsContext *c, const uint8_t *const src[3], const int srcStride[3],
                                         const int64_t srcFrameStride[3], uint8_t *dst, int dstStride,
                                         int64_t dstFrameStride, int nframes)
{
    if (!c || !src || !srcStride || !srcFrameStride || !dst) return B200_EINVAL;
    B200_CUDA_OK(cudaSetDevice(c->dev->ordinal));
    const long long ss[3] = { srcStride[0], srcStride[1], srcStride[2] };
    const long long fs[3] = { srcFrameStride[0], srcFrameStride[1], srcFrameStride[2] };
    return launch_batch(c, c->dev->stream, src, ss, fs, dst, dstStride, dstFrameStride, nframes);
}

// device-side packed layout used by the host-pointer entry points
struct PackedLayout {
    size_t yPitch, cPitch, dPitch, yOff, uOff, vOff, srcBytes, dstBytes;
};
static PackedLayout packed_layout(const SwsPlan &p)
{
    PackedLayout L;
    L.yPitch = ((size_t)p.srcW + 255) & ~(size_t)255;
    L.cPitch = ((size_t)p.chrSrcW + 255) & ~(size_t)255;
    L.dPitch = ((size_t)p.dstW * 3 + 255) & ~(size_t)255;
    L.yOff = 0; L.uOff = L.yPitch * p.srcH; L.vOff = L.uOff + L.cPitch * p.chrSrcH;
    L.srcBytes = L.vOff + L.cPitch * p.chrSrcH;
    L.dstBytes = L.dPitch * p.dstH;
    return L;
}

B200_API int b200_sws_scale_batch_host(B200SwsContext *c, const uint8_t *const src[3], const int srcStride[3],
                                       const int64_t srcFrameStride[3], uint8_t *dst, int dstStride,
                                       int64_t dstFrameStride, int nframes)
{
    if (!c || !src || !srcStride || !srcFrameStride || !dst || nframes < 0) return B200_EINVAL;
    B200Device *d = c->dev;
    B200_CUDA_OK(cudaSetDevice(d->ordinal));
    const SwsPlan &p = c->plan;
    const PackedLayout L = packed_layout(p);
    // chunk so that the 3 in-flight chunks stay within ~1.5 GB of scratch, and kernels stay large enough
    const size_t perFrame = L.srcBytes + L.dstBytes;
    int chunk = (int)((size_t)512 << 20) / (int)(perFrame ? perFrame : 1);
    if (chunk < 1) chunk = 1;
    if (chunk > 32) chunk = 32;
    if (chunk > nframes) chunk = nframes > 0 ? nframes : 1;
    const int K = B200Device::kPipe;
    uint8_t *scr = (uint8_t *)b200_scratch(d, perFrame * chunk * K);
    if (!scr) return B200_ENOMEM;
    B200_CUDA_OK(cudaStreamSynchronize(d->stream));
    const int wbytes[3] = { p.srcW, p.chrSrcW, p.chrSrcW }, rows[3] = { p.srcH, p.chrSrcH, p.chrSrcH };
    const size_t pitch[3] = { L.yPitch, L.cPitch, L.cPitch }, poff[3] = { L.yOff, L.uOff, L.vOff };
    int slot = 0;
    for (int f0 = 0; f0 < nframes; f0 += chunk, slot = (slot + 1) % K) {
        const int nf = nframes - f0 < chunk ? nframes - f0 : chunk;
        cudaStream_t st = d->pipe[slot];
        uint8_t *sbase = scr + (size_t)slot * perFrame * chunk;
        uint8_t *dbase = sbase + L.srcBytes * chunk;
        for (int f = 0; f < nf; f++)
            for (int pl = 0; pl < 3; pl++) {
                const uint8_t *hp = src[pl] + (int64_t)(f0 + f) * srcFrameStride[pl];
                long long hs = srcStride[pl];
                uint8_t *dp = sbase + (size_t)f * L.srcBytes + poff[pl];
                if (hs < 0) {                                           // bottom-up picture: copy from the lowest address, flip on the device side
                    return B200_ENOSYS;
                }
                B200_CUDA_OK(cudaMemcpy2DAsync(dp, pitch[pl], hp, (size_t)hs, wbytes[pl], rows[pl], cudaMemcpyHostToDevice, st));
            }
        const uint8_t *sp[3] = { sbase + L.yOff, sbase + L.uOff, sbase + L.vOff };
        const long long ss[3] = { (long long)L.yPitch, (long long)L.cPitch, (long long)L.cPitch };
        const long long fs[3] = { (long long)L.srcBytes, (long long)L.srcBytes, (long long)L.srcBytes };
        int ret = launch_batch(c, st, sp, ss, fs, dbase, (long long)L.dPitch, (long long)L.dstBytes, nf);
        if (ret < 0) return ret;
        for (int f = 0; f < nf; f++)
            B200_CUDA_OK(cudaMemcpy2DAsync(dst + (int64_t)(f0 + f) * dstFrameStride, (size_t)dstStride,
                                           dbase + (size_t)f * L.dstBytes, L.dPitch, (size_t)p.dstW * 3, p.dstH,
                                           cudaMemcpyDeviceToHost, st));
    }
    for (int i = 0; i < K; i++) B200_CUDA_OK(cudaStreamSynchronize(d->pipe[i]));
    return 0;
}

B200_API int b200_sws_scale(B200SwsContext *c, const uint8_t *const srcSlice[], const int srcStride[],
                            int srcSliceY, int srcSliceH, uint8_t *const dst[], const int dstStride[])
{
    if (!c || !srcSlice || !srcStride || !dst || !dstStride) return B200_EINVAL;
    const SwsPlan &p = c->plan;
    if (srcSliceY != 0 || srcSliceH != p.srcH) {
        b200_set_error("b200_sws_scale: only whole-frame calls are implemented (slice %d+%d of %d)", srcSliceY, srcSliceH, p.srcH);
        return B200_ENOSYS;
    }
    B200Device *d = c->dev;
    B200_CUDA_OK(cudaSetDevice(d->ordinal));
    const PackedLayout L = packed_layout(p);
    uint8_t *scr = (uint8_t *)b200_scratch(d, L.srcBytes + L.dstBytes);
    if (!scr) return B200_ENOMEM;
    cudaStream_t st = d->stream;
    const int wbytes[3] = { p.srcW, p.chrSrcW, p.chrSrcW }, rows[3] = { p.srcH, p.chrSrcH, p.chrSrcH };
    const size_t pitch[3] = { L.yPitch, L.cPitch, L.cPitch }, poff[3] = { L.yOff, L.uOff, L.vOff };
    const uint8_t *sp[3]; long long ss[3];
    for (int pl = 0; pl < 3; pl++) {
        const long long hs = srcStride[pl];
        const long long habs = hs < 0 ? -hs : hs;
        // negative stride (bottom-up, swscale.c:1141-1159): copy from the lowest address, walk upwards on the device
        const uint8_t *lo = hs < 0 ? srcSlice[pl] + (long long)(rows[pl] - 1) * hs : srcSlice[pl];
        B200_CUDA_OK(cudaMemcpy2DAsync(scr + poff[pl], pitch[pl], lo, (size_t)habs, wbytes[pl], rows[pl], cudaMemcpyHostToDevice, st));
        sp[pl] = hs < 0 ? scr + poff[pl] + (size_t)(rows[pl] - 1) * pitch[pl] : scr + poff[pl];
        ss[pl] = hs < 0 ? -(long long)pitch[pl] : (long long)pitch[pl];
    }
    const long long fs[3] = { 0, 0, 0 };
    uint8_t *dd = scr + L.srcBytes;
    int ret = launch_batch(c, st, sp, ss, fs, dd, (long long)L.dPitch, 0, 1);
    if (ret < 0) return ret;
    const long long dsl = dstStride[0];
    const long long dabs = dsl < 0 ? -dsl : dsl;
    if (dsl < 0) {
        // flip while copying back: row r of the device picture goes to dst[0] + r*dsl
        for (int r = 0; r < p.dstH; r++)
            B200_CUDA_OK(cudaMemcpyAsync(dst[0] + (long long)r * dsl, dd + (size_t)r * L.dPitch, (size_t)p.dstW * 3, cudaMemcpyDeviceToHost, st));
    } else {
        B200_CUDA_OK(cudaMemcpy2DAsync(dst[0], (size_t)dabs, dd, L.dPitch, (size_t)p.dstW * 3, p.dstH, cudaMemcpyDeviceToHost, st));
    }
    B200_CUDA_OK(cudaStreamSynchronize(st));
    return p.dstH;
}

B200_API int b200_sws_func(void *c, const uint8_t *const src[], const int srcStride[], int srcSliceY, int srcSliceH,
                           uint8_t *const dst[], const int dstStride[])
{
    return b200_sws_scale((B200SwsContext *)c, src, srcStride, srcSliceY, srcSliceH, dst, dstStride);
}
