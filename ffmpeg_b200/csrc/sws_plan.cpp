// sws_plan.cpp — host-side set-up of the swscale replacement: chroma geometry, scaler tap tables and the
// yuv->rgb constants.  This is cold code that must agree bit-for-bit with what the reference computes in
//   ff_sws_init_single_context  libswscale/utils.c:1137-1760
//   initFilter                  libswscale/utils.c:197-612  (filterAlign = 1: the generic-C layout; the reference's
//                               x86/NEON builds only append zero taps, utils.c:1675-1710, which cannot change results
//                               under SWS_BITEXACT)
//   ff_yuv2rgb_c_init_tables    libswscale/yuv2rgb.c:717-914
//   packed_vscale               libswscale/vscale.c:144-169 (writer choice per output line)
// The device kernels consume the tables produced here; nothing below touches the GPU.
#include "sws_plan.h"
#include "b200dsp.h"
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <algorithm>
#include <limits>

namespace {

using i64 = int64_t;

inline int floor_log2(unsigned v) { int n = 0; v |= 1; while (v >>= 1) ++n; return n; }
inline i64 iabs64(i64 v) { return v < 0 ? -v : v; }
inline int clip_byte(i64 v) { return v < 0 ? 0 : v > 255 ? 255 : (int)v; }
inline int chroma_shift_up(int v, int sh) { return -((-v) >> sh); }          // AV_CEIL_RSHIFT

// get_local_pos, utils.c:168-175
inline int sample_origin(int subsample, int pos)
{
    if (pos == -1 || pos <= -513) pos = (128 << subsample) - 128;
    return (pos + 128) >> subsample;
}

enum Kernel { K_POINT, K_LINEAR2, K_CUBIC, K_TRIANGLE, K_BOX };

// One row of raw (un-normalised, 64-bit) taps per output sample.
struct RawTaps {
    std::vector<i64> w;
    std::vector<int32_t> first;
    int taps = 0;
};

// The weight of a source sample at distance d (Q30 after the <<13) from the output sample centre.
constexpr double kParamDefault = 123456.0;                                       // SWS_PARAM_DEFAULT, swscale.h
i64 cubic_weight(i64 d, const double *param)
{
    // Mitchell-Netravali, B = param[0] (default 0), C = param[1] (default 0.6) in Q24, utils.c:312-332
    const i64 B = (param[0] != kParamDefault ? param[0] : 0) * (1 << 24), C = (param[1] != kParamDefault ? param[1] : 0.6) * (1 << 24);
    if (d >= (i64)1 << 31) return 0;
    const i64 dd = (d * d) >> 30, ddd = (dd * d) >> 30;
    if (d < (i64)1 << 30)
        return (12 * (1 << 24) - 9 * B - 6 * C) * ddd + (-18 * (1 << 24) + 12 * B + 6 * C) * dd +
               (6 * (1 << 24) - 2 * B) * ((i64)1 << 30);
    return (-B - 6 * C) * ddd + (6 * B + 30 * C) * dd + (-12 * B - 48 * C) * d + (8 * B + 24 * C) * ((i64)1 << 30);
}

// the kernels initFilter evaluates in double precision (utils.c:325-368); d = distance in 2.30 fixed point.
// The products and quotients keep the reference's order: these run on the host with the host's libm, like the reference's do.
i64 float_kernel_weight(int scaler, i64 d, i64 unit, const double *param)
{
    const double fd = d * (1.0 / (1 << 30));
    i64 w;
    if (scaler == B200_SWS_X) {
        double c = fd < 1.0 ? cos(fd * M_PI) : -1.0;
        const double A = param[0] != kParamDefault ? param[0] : 1.0;
        c = c < 0.0 ? -pow(-c, A) : pow(c, A);
        w = (c * 0.5 + 0.5) * unit;
    } else if (scaler == B200_SWS_GAUSS) {
        const double p = param[0] != kParamDefault ? param[0] : 3.0;
        w = exp2(-p * fd * fd) * unit;
    } else if (scaler == B200_SWS_SINC) {
        w = (d ? sin(fd * M_PI) / (fd * M_PI) : 1.0) * unit;
    } else if (scaler == B200_SWS_LANCZOS) {
        const double p = param[0] != kParamDefault ? param[0] : 3.0;
        w = (d ? sin(fd * M_PI) * sin(fd * M_PI / p) / (fd * fd * M_PI * M_PI / p) : 1.0) * unit;
        if (fd > p) w = 0;
    } else {                                                                     // spline: getSplineCoeff(1, 0, p, -p - 1, dist), utils.c:155-167
        double a = 1.0, b = 0.0, c = -2.196152422706632, e = 2.196152422706632 - 1.0, dist = fd;
        while (dist > 1.0) {
            const double nb = b + 2.0 * c + 3.0 * e, nc = c + 3.0 * e, ne = -b - 3.0 * c - 6.0 * e;
            a = 0.0; b = nb; c = nc; e = ne; dist -= 1.0;
        }
        w = (((e * dist + c) * dist + b) * dist + a) * unit;
    }
    return w;
}

int raw_taps(RawTaps &r, int scaler, int inc, int srcN, int dstN, int srcOrg, int dstOrg, i64 unit, const double *param)
{
    r.first.assign((size_t)dstN + 3, 0);
    if (std::abs(inc - 0x10000) < 10 && srcOrg == dstOrg) {                      // utils.c:221-231
        r.taps = 1;
        r.w.assign(dstN, unit);
        for (int i = 0; i < dstN; i++) r.first[i] = i;
        return 0;
    }
    if (scaler == B200_SWS_POINT) {                                              // utils.c:232-246
        r.taps = 1;
        r.w.assign(dstN, unit);
        i64 x = ((dstOrg * (i64)inc) >> 8) - ((srcOrg * 0x8000LL) >> 7);
        for (int i = 0; i < dstN; i++, x += inc) r.first[i] = (int32_t)((x + (1 << 15)) >> 16);
        return 0;
    }
    if ((inc <= (1 << 16) && scaler == B200_SWS_AREA) || scaler == B200_SWS_FAST_BILINEAR) {   // utils.c:247-271
        r.taps = 2;
        r.w.assign((size_t)dstN * 2, 0);
        i64 x = ((dstOrg * (i64)inc) >> 8) - ((srcOrg * 0x8000LL) >> 7);
        for (int i = 0; i < dstN; i++, x += inc) {
            int s = (int)((x - (1 << 15) + (1 << 15)) >> 16);
            r.first[i] = s;
            for (int j = 0; j < 2; j++, s++)
                r.w[(size_t)i * 2 + j] = std::max<i64>(0, unit - iabs64((i64)s * (1 << 16) - x) * (unit >> 16));
        }
        return 0;
    }
    int support;                                                                 // scale_algorithms[], utils.c:183-195
    switch (scaler) {
    case B200_SWS_BICUBIC:  support = 4; break;
    case B200_SWS_BILINEAR: support = 2; break;
    case B200_SWS_AREA:     support = 1; break;
    case B200_SWS_X: case B200_SWS_GAUSS:     support = 8; break;
    case B200_SWS_SINC: case B200_SWS_SPLINE: support = 20; break;
    case B200_SWS_LANCZOS:  support = param[0] != kParamDefault ? (int)ceil(2 * param[0]) : 6; break;   // utils.c:278-279
    default: return B200_ENOSYS;
    }
    if (support > 50 || support <= 0) return B200_EINVAL;                        // utils.c:280-285 (0 would trip the reference's assert)
    int taps = inc <= (1 << 16) ? 1 + support : 1 + (int)(((i64)support * srcN + dstN - 1) / dstN);
    taps = std::max(1, std::min(taps, srcN - 2));
    r.taps = taps;
    r.w.assign((size_t)dstN * taps, 0);
    i64 x = ((dstOrg * (i64)inc) >> 7) - ((srcOrg * 0x10000LL) >> 7);
    for (int i = 0; i < dstN; i++, x += 2LL * inc) {
        int s = (int)((x - (taps - 2) * ((i64)1 << 16)) / (1 << 17));
        r.first[i] = s;
        for (int j = 0; j < taps; j++, s++) {
            i64 d = iabs64((i64)s * (1 << 17) - x) << 13;
            if (inc > (1 << 16)) d = d * dstN / srcN;
            i64 w;
            if (scaler == B200_SWS_BICUBIC) {
                w = cubic_weight(d, param) / (((i64)1 << 54) / unit);
            } else if (scaler == B200_SWS_AREA) {
                const i64 d2 = d - (1 << 29);
                if (d2 * inc < -((i64)1 << 45))      w = (i64)1 << 46;
                else if (d2 * inc < ((i64)1 << 45))  w = -d2 * inc + ((i64)1 << 45);
                else                                 w = 0;
                w *= unit >> 46;
            } else if (scaler == B200_SWS_BILINEAR) {
                w = std::max<i64>(0, (1 << 30) - d) * (unit >> 30);
            } else {
                w = float_kernel_weight(scaler, d, unit, param);
            }
            r.w[(size_t)i * taps + j] = w;
        }
    }
    return 0;
}

int build_bank(SwsFilterBank &bank, int scaler, int inc, int srcN, int dstN, int one, int srcOrg, int dstOrg, const double *param,
               const std::vector<double> *srcVec = nullptr, int dstVecLen = 0)
{
    const i64 unit = (i64)1 << (54 - std::min(floor_log2((unsigned)(srcN / dstN)), 8));
    RawTaps r;
    int ret = raw_taps(r, scaler, inc, srcN, dstN, srcOrg, dstOrg, unit, param);
    if (ret < 0) return ret;
    // srcFilter / dstFilter (utils.c:384-413): the source vector is convolved into every row -- `filter2[...] += coeff * filter[...]` on an
    // int64 accumulator, i.e. through a double each time --, the destination vector only widens the row ("FIXME dstFilter" in the
    // reference); the window start moves by the difference of the two half widths
    const int sl = srcVec && !srcVec->empty() ? (int)srcVec->size() : 0;
    if (sl || dstVecLen > 0) {
        const int T0 = r.taps, T2 = T0 + (sl ? sl - 1 : 0) + (dstVecLen > 0 ? dstVecLen - 1 : 0);
        std::vector<i64> w2((size_t)dstN * T2, 0);
        for (int i = 0; i < dstN; i++) {
            if (sl) {
                for (int k = 0; k < sl; k++)
                    for (int j = 0; j < T0; j++) {
                        i64 &acc = w2[(size_t)i * T2 + k + j];
                        acc = (i64)((double)acc + (*srcVec)[k] * (double)r.w[(size_t)i * T0 + j]);
                    }
            } else
                for (int j = 0; j < T0; j++) w2[(size_t)i * T2 + j] = r.w[(size_t)i * T0 + j];
            r.first[i] += (T0 - 1) / 2 - (T2 - 1) / 2;
        }
        r.w.swap(w2);
        r.taps = T2;
    }
    const int T = r.taps;
    const double cutoff = 0.002 * (double)unit;                                  // SWS_MAX_REDUCE_CUTOFF, swscale.h:447

    // 1. trim negligible taps: shift them out on the left, count them on the right (utils.c:417-457)
    int keep = 0;
    for (int i = dstN - 1; i >= 0; i--) {
        i64 *w = &r.w[(size_t)i * T];
        i64 mass = 0;
        for (int j = 0; j < T; j++) {
            mass += iabs64(w[0]);
            if ((double)mass > cutoff) break;
            if (i < dstN - 1 && r.first[i] >= r.first[i + 1]) break;             // keep positions monotonic
            std::copy(w + 1, w + T, w);
            w[T - 1] = 0;
            r.first[i]++;
        }
        int need = T;
        mass = 0;
        for (int j = T - 1; j > 0; j--) {
            mass += iabs64(w[j]);
            if ((double)mass > cutoff) break;
            need--;
        }
        keep = std::max(keep, need);
    }
    if (keep >= 256) return B200_ENOSYS;                                         // reference cascades contexts here (utils.c:491-495)

    // 2. fold taps that fall outside [0, srcN) onto the border sample (utils.c:519-560)
    std::vector<i64> g((size_t)dstN * keep);
    for (int i = 0; i < dstN; i++) {
        i64 *o = &g[(size_t)i * keep];
        for (int j = 0; j < keep; j++) o[j] = j < T ? r.w[(size_t)i * T + j] : 0;
        int &p = r.first[i];
        if (p < 0) {
            for (int j = 1; j < keep; j++) {
                o[std::max(j + p, 0)] += o[j];
                o[j] = 0;
            }
            p = 0;
        }
        if (p + keep > srcN) {
            const int shift = p + std::min(keep - srcN, 0);
            i64 spill = 0;
            for (int j = keep - 1; j >= 0; j--)
                if (p + j >= srcN) { spill += o[j]; o[j] = 0; }
            for (int j = keep - 1; j >= 0; j--) o[j] = j < shift ? 0 : o[j - shift];
            p -= shift;
            o[srcN - 1 - p] += spill;
        }
    }

    // 3. normalise every row to `one` with error feedback (utils.c:568-588)
    bank.size = keep;
    bank.n = dstN;
    bank.coef.assign((size_t)dstN * keep, 0);
    bank.pos.assign(r.first.begin(), r.first.begin() + dstN);
    for (int i = 0; i < dstN; i++) {
        const i64 *o = &g[(size_t)i * keep];
        i64 sum = 0, carry = 0;
        for (int j = 0; j < keep; j++) sum += o[j];
        sum = (sum + one / 2) / one;
        if (!sum) sum = 1;
        for (int j = 0; j < keep; j++) {
            const i64 v = o[j] + carry;
            const int q = (int)(v >= 0 ? (v + (sum >> 1)) / sum : (v - (sum >> 1)) / sum);   // ROUNDED_DIV
            bank.coef[(size_t)i * keep + j] = (int16_t)q;
            carry = v - q * sum;
        }
    }
    return 0;
}

inline int16_t round_q16(i64 f)                                                  // roundToInt16, yuv2rgb.c:705-715
{
    int r = (int)((f + (1 << 15)) >> 16);
    if (r < -0x7FFF) return (int16_t)0x8000;
    if (r > 0x7FFF) return 0x7FFF;
    return (int16_t)r;
}

} // namespace

bool SwsFilterBank::identity() const
{
    if (size != 1) return false;
    for (int i = 0; i < n; i++)
        if (pos[i] != i || coef[i] != (1 << 14)) return false;
    return true;
}

int sws_plan_colorspace(SwsPlan &p, const int inv_table[4], int fullRange, int brightness, int contrast, int saturation)
{
    const int headroom = 512;                                                    // YUVRGB_TABLE_LUMA_HEADROOM
    const int yoffs = (fullRange ? 384 : 326) + headroom;
    i64 crv = inv_table[0], cbu = inv_table[1], cgu = -(i64)inv_table[2], cgv = -(i64)inv_table[3];
    i64 cy = 1 << 16, oy = 0;
    if (!fullRange) { cy = (cy * 255) / 219; oy = 16 << 16; }
    else { crv = (crv * 224) / 255; cbu = (cbu * 224) / 255; cgu = (cgu * 224) / 255; cgv = (cgv * 224) / 255; }
    cy  = (cy * contrast) >> 16;
    crv = (crv * contrast * saturation) >> 32;
    cbu = (cbu * contrast * saturation) >> 32;
    cgu = (cgu * contrast * saturation) >> 32;
    cgv = (cgv * contrast * saturation) >> 32;
    oy -= 256LL * brightness;

    SwsColorConst &c = p.color;
    c.y_coeff = round_q16(cy * (1 << 13));
    c.y_offset = round_q16(oy * (1 << 9));
    c.v2r = round_q16(crv * (1 << 13));
    c.v2g = round_q16(cgv * (1 << 13));
    c.u2g = round_q16(cgu * (1 << 13));
    c.u2b = round_q16(cbu * (1 << 13));

    const i64 den = std::max<i64>(cy, 1);
    crv = ((crv * (1 << 16)) + 0x8000) / den;
    cbu = ((cbu * (1 << 16)) + 0x8000) / den;
    cgu = ((cgu * (1 << 16)) + 0x8000) / den;
    cgv = ((cgv * (1 << 16)) + 0x8000) / den;

    const i64 yb0 = -((i64)384 << 16) - headroom * cy - oy + 0x8000;
    const i64 baseR = yoffs - (crv >> 9), baseB = yoffs - (cbu >> 9), baseG = yoffs - (cgu >> 9) - (cgv >> 9);
    // the kernels evaluate the tables in 32-bit arithmetic: refuse settings that would overflow there
    const i64 lim = std::numeric_limits<int32_t>::max();
    const i64 worst_idx = 4096;                                                  // |table index + luma| bound with margin
    if (cy <= 0 || iabs64(yb0) + worst_idx * cy >= lim ||
        255 * iabs64(crv) >= lim || 255 * iabs64(cbu) >= lim || 255 * iabs64(cgu) >= lim || 255 * iabs64(cgv) >= lim ||
        iabs64(baseR) > 2048 || iabs64(baseG) > 2048 || iabs64(baseB) > 2048)
        return B200_ENOSYS;
    c.cy = (int)cy; c.yb0 = (int)yb0;
    c.baseR = (int)baseR; c.baseG = (int)baseG; c.baseB = (int)baseB;
    c.crv = (int)crv; c.cgu = (int)cgu; c.cgv = (int)cgv; c.cbu = (int)cbu;
    return 0;
}

// solve_range_convert (swscale.c:577-589) for an 8-bit destination: src_bits 15, src_shift 7, mult_shift 14
// (init_range_convert_constants, :591-600); the line functions narrow the coefficient to uint16 and the offset to int32
static void solve_range(unsigned srcMin, unsigned srcMax, unsigned dstMin, unsigned dstMax, int &coeff, int &offset)
{
    const int srcShift = 7, multShift = 14, totalShift = srcShift + multShift;
    const uint64_t srcRange = srcMax - srcMin, dstRange = dstMax - dstMin;
    const uint64_t q = (dstRange << totalShift) / srcRange;
    const uint32_t c = (uint32_t)((q + (1u << srcShift) - 1) >> srcShift);                      // AV_CEIL_RSHIFT
    const i64 o = ((i64)dstMax << totalShift) - ((i64)srcMax << srcShift) * c + (1u << (multShift - 1));
    coeff = (uint16_t)c;
    offset = (int32_t)o;
}

// ff_sws_init_range_convert (swscale.c:626-660): only a yuv destination converts, and only when the ranges differ
static void plan_range_convert(SwsPlan &p)
{
    p.range_conv = 0;
    if (p.src_range == p.dst_range || !sws_out_is_yuv(p.out.kind)) return;
    if (p.src_range) {
        solve_range(0, 255, 16, 235, p.lumRangeCoeff, p.lumRangeOffset);
        solve_range(0, 255, 16, 240, p.chrRangeCoeff, p.chrRangeOffset);
        p.range_conv = 2;
    } else {
        solve_range(16, 235, 0, 255, p.lumRangeCoeff, p.lumRangeOffset);
        solve_range(16, 240, 0, 255, p.chrRangeCoeff, p.chrRangeOffset);
        p.range_conv = 1;
    }
}

static i64 rounded_div(i64 a, i64 b) { return a >= 0 ? (a + (b >> 1)) / b : (a - (b >> 1)) / b; }

// fill_rgb2yuv_table (utils.c:614-700): always the limited-range matrix (full range comes from the range conversion between
// the passes); the default BT.601 table gets the hand-rounded constants of utils.c:692-702
static void plan_rgb2yuv(SwsPlan &p, const int table[4])
{
    static const int bt601[4] = { 104597, 132201, 25675, 53279 };
    const i64 ONE = 65536;
    const i64 vr = table[0], ub = table[1], ug = -(i64)table[2], vg = -(i64)table[3], cy = ONE * 255 / 219;
    if (!vr || !ub) return;                                                       // (the reference would divide by zero)
    const i64 W = rounded_div(ONE * ONE * ug, ub), V = rounded_div(ONE * ONE * vg, vr), Z = ONE * ONE - W - V;
    const i64 Cy = rounded_div(cy * Z, ONE), Cu = rounded_div(ub * Z, ONE), Cv = rounded_div(vr * Z, ONE);
    if (!Cy || !Cu || !Cv) return;
    int *t = p.rgb2yuv;
    t[0] = (int)-rounded_div((1 << 15) * V, Cy);      t[1] = (int)rounded_div((1 << 15) * ONE * ONE, Cy);  t[2] = (int)-rounded_div((1 << 15) * W, Cy);
    t[3] = (int)rounded_div((1 << 15) * V, Cu);       t[4] = (int)-rounded_div((1 << 15) * ONE * ONE, Cu); t[5] = (int)rounded_div((1 << 15) * (Z + W), Cu);
    t[6] = (int)rounded_div((1 << 15) * (V + Z), Cv); t[7] = (int)-rounded_div((1 << 15) * ONE * ONE, Cv); t[8] = (int)rounded_div((1 << 15) * W, Cv);
    if (!memcmp(table, bt601, sizeof(bt601))) {
        t[2] =  (int)(0.114 * 219 / 255 * (1 << 15) + 0.5); t[8] = -(int)(0.081 * 224 / 255 * (1 << 15) + 0.5);
        t[5] =  (int)(0.500 * 224 / 255 * (1 << 15) + 0.5); t[1] =  (int)(0.587 * 219 / 255 * (1 << 15) + 0.5);
        t[7] = -(int)(0.419 * 224 / 255 * (1 << 15) + 0.5); t[4] = -(int)(0.331 * 224 / 255 * (1 << 15) + 0.5);
        t[0] =  (int)(0.299 * 219 / 255 * (1 << 15) + 0.5); t[6] =  (int)(0.500 * 224 / 255 * (1 << 15) + 0.5);
        t[3] = -(int)(0.169 * 224 / 255 * (1 << 15) + 0.5);
    }
}

int sws_plan_colorspace_details(SwsPlan &p, const int inv_table[4], int srcRange, const int table[4], int dstRange,
                                int brightness, int contrast, int saturation)
{
    const bool yuvDst = sws_out_is_yuv(p.out.kind);
    if (!yuvDst) dstRange = 0;                                                   // range_override_needed(dst), utils.c:877-878
    if (p.src_rgb) srcRange = 0;                                                 // range_override_needed(src), utils.c:879-880
    plan_rgb2yuv(p, table);                                                      // utils.c:1002
    memcpy(p.src_cs, inv_table, sizeof(p.src_cs));
    memcpy(p.dst_cs, table, sizeof(p.dst_cs));
    p.src_range = srcRange;
    p.dst_range = dstRange;
    plan_range_convert(p);
    if (yuvDst && !p.src_rgb)   // utils.c:910-989: with different matrices the reference cascades yuv -> bgr24 -> yuv: one plan cannot; b200_sws_setColorspaceDetails (sws.cu) builds the two contexts
        return memcmp(p.src_cs, p.dst_cs, sizeof(p.src_cs)) ? B200_ENOSYS : 0;
    if (yuvDst) return 0;       // RGB -> yuv: only the rgb2yuv table and the ranges matter
    return sws_plan_colorspace(p, inv_table, srcRange, brightness, contrast, saturation);
}

bool sws_out_format(int f, SwsOutFmt &o)
{
    switch (f) {
    case B200_PIX_FMT_RGB24: o = { SWS_OUT_RGB24, 3, 0, 1, 2, -1 }; return true;
    case B200_PIX_FMT_BGR24: o = { SWS_OUT_BGR24, 3, 2, 1, 0, -1 }; return true;
    case B200_PIX_FMT_RGBA:  o = { SWS_OUT_RGBA,  4, 0, 1, 2, 3 };  return true;
    case B200_PIX_FMT_BGRA:  o = { SWS_OUT_BGRA,  4, 2, 1, 0, 3 };  return true;
    case B200_PIX_FMT_ARGB:  o = { SWS_OUT_ARGB,  4, 1, 2, 3, 0 };  return true;
    case B200_PIX_FMT_ABGR:  o = { SWS_OUT_ABGR,  4, 3, 2, 1, 0 };  return true;
    case B200_PIX_FMT_YUV420P: o = { SWS_OUT_YUV420P, 1, 0, 0, 0, -1 }; return true;
    case B200_PIX_FMT_NV12:    o = { SWS_OUT_NV12, 1, 0, 0, 0, -1 }; return true;
    case B200_PIX_FMT_NV21:    o = { SWS_OUT_NV21, 1, 0, 0, 0, -1 }; return true;
    }
    return false;
}

int sws_plan_build(SwsPlan &p, int srcW, int srcH, int dstW, int dstH, int flags, int srcRange, int dstRange)
{
    if (srcW < 1 || srcH < 1 || dstW < 1 || dstH < 1) return B200_EINVAL;
    int algo = flags & 0x7FF;                                                    // scaler bits, utils.c:1196-1222
    if (!algo) { algo = B200_SWS_BICUBIC; flags |= algo; }
    else if (algo & (algo - 1)) return B200_EINVAL;
    if (algo == B200_SWS_FAST_BILINEAR && (srcW < 8 || dstW <= 8)) {             // utils.c:1224-1230
        algo = B200_SWS_BILINEAR;
        flags ^= B200_SWS_FAST_BILINEAR | algo;
    }
    // the fast horizontal functions exist for 8-bit input lines only (swscale.c:675-681); RGB sources are srcBpc 16
    // (utils.c:1407-1408) and use the 2-tap filter initFilter builds for the flag
    p.fast_bilinear = algo == B200_SWS_FAST_BILINEAR && !p.src_rgb;
    p.planar = sws_out_is_yuv(p.out.kind);
    p.dst_nv = p.out.kind == SWS_OUT_NV12 ? 1 : p.out.kind == SWS_OUT_NV21 ? 2 : 0;
    // utils.c:1256-1263,1608,1624: a source or destination filter vector longer than one tap rules out every unscaled special converter
    bool usesFilter = false;
    for (int k = 0; k < 4; k++) usesFilter = usesFilter || p.srcFilt[k].size() > 1 || p.dstFiltLen[k] > 1;
    p.rgb_shuffle = false;
    if (!usesFilter && p.src_rgb && !p.planar && srcW == dstW && srcH == dstH) {
        // same size, packed RGB both sides: a copy for equal formats (packedCopyWrapper, swscale_unscaled.c:2675-2690), else the byte
        // shuffles of rgbToRgbWrapper (:2001-2060) when findRgbConvFn (:1843-1998) has one: always between the four 32-bit orders,
        // between rgb24 and bgr24 and from 32 to 24 bits; from 24 to 32 bits unless SWS_BITEXACT asks for bgra / rgba (little endian:
        // RGB32 / BGR32, "maintain symmetry between endianness" :1992-1995), which then goes through the scaler like any other pair
        const bool to32 = p.src_rgb == 3 && p.out.bpp == 4;
        const bool a_last = p.out.bpp == 4 && p.out.ao == 3;                                 // bgra (AV_PIX_FMT_RGB32) / rgba (BGR32) on little endian
        if (!(to32 && a_last && (flags & B200_SWS_BITEXACT))) {
            const int sao = p.src_rgb == 4 ? 6 - p.sro - p.sgo - p.sbo : 4;                  // 4 selects the constant 255
            unsigned sel = 0;
            sel |= (unsigned)p.sro << (4 * p.out.ro);
            sel |= (unsigned)p.sgo << (4 * p.out.go);
            sel |= (unsigned)p.sbo << (4 * p.out.bo);
            if (p.out.bpp == 4) sel |= (unsigned)sao << (4 * p.out.ao);
            p.rgb_shuffle = true; p.shuffle_sel = sel;
            p.srcW = srcW; p.srcH = srcH; p.dstW = dstW; p.dstH = dstH; p.flags = flags;
            p.chrSrcW = srcW; p.chrSrcH = srcH; p.chrDstW = dstW; p.chrDstH = dstH; p.chrSrcHSub = p.chrSrcVSub = 0; p.chrDstHSub = 0;
            return 0;
        }
    }
    p.need_alpha = p.src_rgb == 4 && !p.planar && p.out.bpp == 4;
    if (!p.planar && (dstW & 1)) flags |= B200_SWS_FULL_CHR_H_INT;               // utils.c:1271-1276 (RGB destinations only)
    if (!p.planar && p.src_rgb && !(flags & B200_SWS_FAST_BILINEAR))
        flags |= B200_SWS_FULL_CHR_H_INT;                                        // utils.c:1277-1285: source chroma is not subsampled
    p.srcW = srcW; p.srcH = srcH; p.dstW = dstW; p.dstH = dstH; p.flags = flags;
    p.chrDstHSub = (!p.planar && (flags & B200_SWS_FULL_CHR_H_INT)) ? 0 : 1;     // utils.c:1359-1360
    const int chrDstVSub = p.planar ? 1 : 0;                                     // av_pix_fmt_get_chroma_sub_sample(dstFormat), utils.c:1266
    p.chrSrcHSub = 1; p.chrSrcVSub = 1;
    if (p.src_rgb) {                                                             // utils.c:1366-1393: every other pixel for chroma unless asked otherwise
        p.chrSrcVSub = 0;
        p.chrSrcHSub = (!(srcW & 1) && !(flags & B200_SWS_FULL_CHR_H_INP) &&
                        ((dstW >> p.chrDstHSub) <= (srcW >> 1) || (flags & B200_SWS_FAST_BILINEAR))) ? 1 : 0;
    }
    p.chrSrcW = chroma_shift_up(srcW, p.chrSrcHSub);
    p.chrSrcH = chroma_shift_up(srcH, p.chrSrcVSub);
    p.chrDstW = chroma_shift_up(dstW, p.chrDstHSub);
    p.chrDstH = chroma_shift_up(dstH, chrDstVSub);                               // packed RGB has no vertical chroma subsampling
    static const int bt601[4] = { 104597, 132201, 25675, 53279 };                // ff_yuv2rgb_coeffs[SWS_CS_DEFAULT], yuv2rgb.c:47-59
    // utils.c:1164-1167: the ranges the context holds at initialisation go through sws_setColorspaceDetails
    int ret = sws_plan_colorspace_details(p, bt601, srcRange, bt601, dstRange, 0, 1 << 16, 1 << 16);
    if (ret < 0) return ret;

    // unscaled special converter gate: swscale_unscaled.c:2426-2431 reached from utils.c:1623-1637; a yuv destination only
    // looks for one when no range conversion is due (utils.c:1624-1626)
    // bgr24ToYv12Wrapper (swscale_unscaled.c:2453-2457): bgr24 only, not with accurate_rnd, even width
    p.bgr24_yv12 = !usesFilter && p.planar && !p.dst_nv && p.src_rgb == 3 && p.sbo == 0 && srcW == dstW && srcH == dstH && p.src_range == p.dst_range &&
                   !(flags & B200_SWS_ACCURATE_RND) && !(dstW & 1);
    if (p.bgr24_yv12) return 0;
    // same-size yuv -> yuv: planarCopyWrapper, planarToNv12Wrapper or nv12ToPlanarWrapper (swscale_unscaled.c:2415-2419,2675-2693,147-188);
    // nv12 <-> nv21 has no such converter and goes through the scaler (it matters once sws_setColorspaceDetails changes a range)
    p.planar_copy = !usesFilter && p.planar && !p.src_rgb && srcW == dstW && srcH == dstH && p.src_range == p.dst_range &&
                    !(p.src_nv && p.dst_nv && p.src_nv != p.dst_nv);
    if (p.planar_copy) return 0;
    // (only planar yuv420p / yuv422p sources have the LUT converter; nv12 / nv21 go through the scaler)
    p.unscaled_lut = !usesFilter && !p.planar && !p.src_nv && !p.src_rgb && srcW == dstW && srcH == dstH && !(flags & B200_SWS_ACCURATE_RND) && !(dstH & 1);
    if (p.unscaled_lut) return 0;

    const int lumScaler = algo == B200_SWS_BICUBLIN ? B200_SWS_BICUBIC : algo;
    const int chrScaler = algo == B200_SWS_BICUBLIN ? B200_SWS_BILINEAR : algo;
    const i64 lumXInc = (((i64)srcW << 16) + (dstW >> 1)) / dstW;                // utils.c:1250-1251,1425-1426
    const i64 lumYInc = (((i64)srcH << 16) + (dstH >> 1)) / dstH;
    const i64 chrXInc = (((i64)p.chrSrcW << 16) + (p.chrDstW >> 1)) / p.chrDstW;
    const i64 chrYInc = (((i64)p.chrSrcH << 16) + (p.chrDstH >> 1)) / p.chrDstH;
    if (lumXInc < 10 || lumYInc < 10 || chrXInc < 10 || chrYInc < 10 ||
        lumXInc > INT32_MAX || lumYInc > INT32_MAX || chrXInc > INT32_MAX || chrYInc > INT32_MAX)
        return B200_ENOSYS;
    p.lumXInc = (int)lumXInc; p.chrXInc = (int)chrXInc;
    if ((ret = build_bank(p.hLum, lumScaler, (int)lumXInc, srcW, dstW, 1 << 14, sample_origin(0, 0), sample_origin(0, 0), p.param, &p.srcFilt[0], p.dstFiltLen[0])) < 0) return ret;
    if ((ret = build_bank(p.hChr, chrScaler, (int)chrXInc, p.chrSrcW, p.chrDstW, 1 << 14,
                          sample_origin(p.chrSrcHSub, -513), sample_origin(p.chrDstHSub, -513), p.param, &p.srcFilt[2], p.dstFiltLen[2])) < 0) return ret;
    if ((ret = build_bank(p.vLum, lumScaler, (int)lumYInc, srcH, dstH, 1 << 12, sample_origin(0, 0), sample_origin(0, 0), p.param, &p.srcFilt[1], p.dstFiltLen[1])) < 0) return ret;
    if ((ret = build_bank(p.vChr, chrScaler, (int)chrYInc, p.chrSrcH, p.chrDstH, 1 << 12,
                          sample_origin(p.chrSrcVSub, -513), sample_origin(chrDstVSub, -513), p.param, &p.srcFilt[3], p.dstFiltLen[3])) < 0) return ret;
    if (p.planar) return 0;

    // writer per output line, as packed_vscale decides it (vscale.c:144-169); coefficients are read as uint16 there
    p.rowMode.assign((size_t)dstH * 4, 0);
    const int L = p.vLum.size, C = p.vChr.size;
    for (int y = 0; y < dstH; y++) {
        const uint16_t *lf = (const uint16_t *)&p.vLum.coef[(size_t)y * L];
        const uint16_t *cf = (const uint16_t *)&p.vChr.coef[(size_t)y * C];
        int mode = 0, ya = 0, ua = 0;
        if (L == 1 && C == 1) mode = 1;
        else if (L == 1 && C == 2 && cf[0] + cf[1] == 4096 && cf[1] <= 4096) { mode = 1; ua = cf[1]; }
        else if (L == 2 && C == 2 && lf[0] + lf[1] == 4096 && lf[1] <= 4096 && cf[0] + cf[1] == 4096 && cf[1] <= 4096) {
            mode = 2; ya = lf[1]; ua = cf[1];
        }
        p.rowMode[(size_t)y * 4 + 0] = mode;
        p.rowMode[(size_t)y * 4 + 1] = ya;
        p.rowMode[(size_t)y * 4 + 2] = ua;
    }
    return 0;
}
