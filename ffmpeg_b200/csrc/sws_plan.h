// sws_plan.h — host-side plan of one yuv420p -> packed 8-bit RGB conversion (what SwsInternal holds after sws_init_context)
#pragma once
#include <vector>
#include <cstdint>

struct SwsFilterBank {            // one of hLum / hChr / vLum / vChr (SwsInternal fields, swscale_internal.h:437-448)
    std::vector<int16_t> coef;    // n * size, normalised to 1<<14 (horizontal) or 1<<12 (vertical)
    std::vector<int32_t> pos;     // n first-tap positions
    int size = 0;
    int n = 0;
    bool identity() const;        // size 1, pos[i] == i, coef == 1<<14  (unscaled horizontal pass)
};

// Closed form of the reference's yuv->rgb look-up tables (libswscale/yuv2rgb.c:717-914, 24 bpp case):
//   y_table[k] = clip_u8((yb0 + k*cy) >> 16)   with the +0x8000 rounding folded into yb0
//   r = y_table[baseR + ((clip_u8(V)*crv) >> 16) + Y]  etc.   (fill_table / fill_gv_table, yuv2rgb.c:680-703)
struct SwsColorConst {
    int cy, yb0;
    int baseR, baseG, baseB;      // baseG = yoffs - (cgu>>9) - (cgv>>9)
    int crv, cgu, cgv, cbu;
    // full-chroma writer (yuv2rgb_write_full, output.c:1998-2030)
    int y_offset, y_coeff, v2r, v2g, u2g, u2b;
};

// Packed 8-bit RGB outputs: bytes per pixel and the byte position of each channel (ao < 0: no alpha byte).
// 32-bit formats carry alpha = 255 (yuv2rgb.c:947-960 adds 255 << abase to the ramp; output.c:2066-2095 stores 255).
enum SwsOutKind { SWS_OUT_RGB24, SWS_OUT_BGR24, SWS_OUT_RGBA, SWS_OUT_BGRA, SWS_OUT_ARGB, SWS_OUT_ABGR,
                  SWS_OUT_YUV420P /* planar destination: three planes, bpp = 1 */,
                  SWS_OUT_NV12, SWS_OUT_NV21 /* luma plane + interleaved chroma plane (U,V / V,U): yuv420p with the chroma interleaved */ };
inline bool sws_out_is_yuv(int kind) { return kind == SWS_OUT_YUV420P || kind == SWS_OUT_NV12 || kind == SWS_OUT_NV21; }
struct SwsOutFmt { int kind, bpp, ro, go, bo, ao; };
bool sws_out_format(int av_pix_fmt, SwsOutFmt &o);

struct SwsPlan {
    int srcW = 0, srcH = 0, dstW = 0, dstH = 0, flags = 0;
    SwsOutFmt out{ SWS_OUT_RGB24, 3, 0, 1, 2, -1 };
    int chrSrcW = 0, chrSrcH = 0, chrDstW = 0, chrDstH = 0;
    int chrDstHSub = 1;
    bool unscaled_lut = false;    // reference installs yuv2rgb_c_24_rgb as convert_unscaled
    int src_nv = 0;               // 0: yuv420p source; 1: nv12, 2: nv21 (plane 1 holds U,V / V,U interleaved; nvXXtoUV_c, input.c:921-948)
    bool fast_bilinear = false;   // SWS_FAST_BILINEAR: horizontal pass = ff_hyscale_fast_c / ff_hcscale_fast_c (hscale_fast_bilinear.c:27-67)
    int lumXInc = 0, chrXInc = 0; // 16.16 horizontal steps (utils.c:1250,1425)
    bool planar = false;          // destination yuv420p / nv12 / nv21 (yuv2planeX / yuv2plane1 / yuv2nv12cX writers, vscale.c:34-107)
    int dst_nv = 0;               // 1: nv12, 2: nv21 destination (planarToNv12Wrapper swscale_unscaled.c:147-165; yuv2nv12cX_c output.c:495-528)
    bool planar_copy = false;     // same size yuv420p -> yuv420p: planarCopyWrapper (swscale_unscaled.c:2220,2675-2693)
    // yuv -> yuv range conversion of the 15-bit lines between the two passes (swscale.c:163-209; constants :577-624)
    int src_range = 0, dst_range = 0;            // SwsContext.src_range / .dst_range (0 limited, non-zero full)
    int range_conv = 0;           // 0 none, 1 limited -> full (lum/chrRangeToJpeg_c: clips at 2^15-1), 2 full -> limited (FromJpeg)
    int lumRangeCoeff = 0, lumRangeOffset = 0, chrRangeCoeff = 0, chrRangeOffset = 0;
    int src_cs[4] = { 104597, 132201, 25675, 53279 }, dst_cs[4] = { 104597, 132201, 25675, 53279 };   // colorspace tables as last set
    // packed RGB source: input readers (input.c:264-393,1068-1172) -> 16-bit lines -> hScale16To15_c (swscale.c:99-125)
    int src_rgb = 0;              // 0: yuv source; 3 / 4: bytes per source pixel (set, with the offsets, before sws_plan_build)
    int sro = 0, sgo = 0, sbo = 0;                // byte positions of R, G, B in a source pixel
    int chrSrcHSub = 1, chrSrcVSub = 1;           // chroma sampling of the source as the scaler sees it (utils.c:1366-1396)
    int rgb2yuv[9] = { 0 };       // input_rgb2yuv_table RY GY BY RU GU BU RV GV BV (swscale_internal.h:468-477; utils.c:614-700)
    bool need_alpha = false;      // c->needAlpha = isALPHA(src) && isALPHA(dst) (utils.c:1405): the alpha plane goes through the scaler
    bool rgb_shuffle = false;     // same size packed RGB -> packed RGB: rgbToRgbWrapper / packedCopyWrapper (swscale_unscaled.c:2001-2060,2138-2170)
    unsigned shuffle_sel = 0;     // __byte_perm selector: nibble j = source byte of destination byte j, 4 = the constant 255
    bool bgr24_yv12 = false;      // reference installs bgr24ToYv12Wrapper (ff_rgb24toyv12_c, rgb2rgb_template.c:580-641)
    // sws_getContext's srcFilter / dstFilter (SwsFilter, swscale.h:199-204), set before sws_plan_build: [0] lumH, [1] lumV, [2] chrH, [3] chrV.
    // initFilter convolves the source vector into every tap row; of the destination vector only the length counts (utils.c:384-413)
    std::vector<double> srcFilt[4];
    int dstFiltLen[4] = { 0, 0, 0, 0 };
    double param[2] = { 123456.0, 123456.0 };     // SwsContext.scaler_params; SWS_PARAM_DEFAULT = 123456 (set before sws_plan_build)
    SwsFilterBank hLum, hChr, vLum, vChr;
    SwsColorConst color{};
    // per output line: writer selected by packed_vscale (vscale.c:144-169): 0 = _X, 1 = _1, 2 = _2, plus alphas
    std::vector<int32_t> rowMode;  // dstH * 4: mode, yalpha, uvalpha, pad
};

// returns 0 or a negative B200_E* code
// srcRange / dstRange: the values SwsContext.src_range / .dst_range hold when sws_init_context runs
int sws_plan_build(SwsPlan &p, int srcW, int srcH, int dstW, int dstH, int flags, int srcRange = 0, int dstRange = 0);
// sws_setColorspaceDetails (utils.c:849-1004): RGB destination -> look-up constants; yuv destination -> ranges + range conversion
int sws_plan_colorspace_details(SwsPlan &p, const int inv_table[4], int srcRange, const int table[4], int dstRange,
                                int brightness, int contrast, int saturation);
int sws_plan_colorspace(SwsPlan &p, const int inv_table[4], int fullRange, int brightness, int contrast, int saturation);
