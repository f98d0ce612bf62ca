/*
 * h264lf_oracle.c — TEST INFRASTRUCTURE ONLY.  CPU restatement of libavcodec's H.264 in-loop deblocking filters, 8 bit: the twelve
 * loop-filter members of H264DSPContext (libavcodec/h264dsp.h:48-73) as ff_h264dsp_init(c, 8, chroma_format_idc) installs them
 * (libavcodec/h264dsp.c:109-132).
 *
 * Follows (behaviour, not text) libavcodec/h264dsp_template.c:
 *   :103-152  h264_loop_filter_luma        bS < 4: p0/q0 moved by a clipped delta, p1/q1 conditionally; a group of lines is skipped when tc0 < 0
 *   :166-222  h264_loop_filter_luma_intra  bS = 4: the strong 3-tap / weak 1-tap smoothing
 *   :236-271  h264_loop_filter_chroma      tc = ((tc0 - 1U) << (depth - 8)) + 1 = tc0 at 8 bit; a group with tc <= 0 is skipped
 *   :293-315  h264_loop_filter_chroma_intra
 * v = filter across a horizontal edge (neighbours one row apart), h = across a vertical edge; the variants differ only in the number of
 * lines per tc0 entry: luma 4, luma mbaff 2, chroma 2, chroma mbaff 1, chroma 4:2:2 h 4, chroma 4:2:2 h mbaff 2.
 */
#include "oracle.h"

static int iabs(int a) { return a < 0 ? -a : a; }
static int clip3(int v, int lo, int hi) { return v < lo ? lo : v > hi ? hi : v; }
static uint8_t clip8(int v) { return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); }

/* kind -> (intra, chroma, horizontal-edge-walk, lines per tc0 entry) */
static int lf_shape(int kind, int *intra, int *chroma, int *vert, int *iters)
{
    static const int tab[16][4] = {
        { 0, 0, 1, 4 }, { 0, 0, 0, 4 }, { 0, 0, 0, 2 },            /* 0 v_luma, 1 h_luma, 2 h_luma_mbaff */
        { 1, 0, 1, 4 }, { 1, 0, 0, 4 }, { 1, 0, 0, 2 },            /* 3 v_luma_intra, 4 h_luma_intra, 5 h_luma_mbaff_intra */
        { 0, 1, 1, 2 }, { 0, 1, 0, 2 }, { 0, 1, 0, 1 },            /* 6 v_chroma, 7 h_chroma, 8 h_chroma_mbaff */
        { 1, 1, 1, 2 }, { 1, 1, 0, 2 }, { 1, 1, 0, 1 },            /* 9 v_chroma_intra, 10 h_chroma_intra, 11 h_chroma_mbaff_intra */
        { 0, 1, 0, 4 }, { 0, 1, 0, 2 }, { 1, 1, 0, 4 }, { 1, 1, 0, 2 },   /* 12-15: the 4:2:2 h_chroma, h_chroma_mbaff, and their intra forms */
    };
    if (kind < 0 || kind > 15) return -1;
    *intra = tab[kind][0]; *chroma = tab[kind][1]; *vert = tab[kind][2]; *iters = tab[kind][3];
    return 0;
}

int orc_h264_loop_filter(int kind, uint8_t *pix, ptrdiff_t stride, int alpha, int beta, const int8_t *tc0)
{
    int intra, chroma, vert, iters;
    if (lf_shape(kind, &intra, &chroma, &vert, &iters) < 0) return -1;
    const ptrdiff_t xs = vert ? stride : 1, ys = vert ? 1 : stride;      /* xs: across the edge, ys: along it */
    for (int line = 0; line < 4 * iters; line++, pix += ys) {
        const int p0 = pix[-1 * xs], p1 = pix[-2 * xs], q0 = pix[0], q1 = pix[1 * xs];
        if (!intra) {
            const int t0 = tc0[line / iters];
            if (chroma ? t0 <= 0 : t0 < 0) continue;
            if (!(iabs(p0 - q0) < alpha && iabs(p1 - p0) < beta && iabs(q1 - q0) < beta)) continue;
            int tc = t0;                                            /* chroma: ((tc0 - 1U) << 0) + 1 = tc0 */
            if (!chroma) {
                const int p2 = pix[-3 * xs], q2 = pix[2 * xs];
                if (iabs(p2 - p0) < beta) {
                    if (t0) pix[-2 * xs] = (uint8_t)(p1 + clip3(((p2 + ((p0 + q0 + 1) >> 1)) >> 1) - p1, -t0, t0));
                    tc++;
                }
                if (iabs(q2 - q0) < beta) {
                    if (t0) pix[xs] = (uint8_t)(q1 + clip3(((q2 + ((p0 + q0 + 1) >> 1)) >> 1) - q1, -t0, t0));
                    tc++;
                }
            }
            const int delta = clip3((((q0 - p0) * 4) + (p1 - q1) + 4) >> 3, -tc, tc);
            pix[-xs] = clip8(p0 + delta);
            pix[0] = clip8(q0 - delta);
        } else {
            if (!(iabs(p0 - q0) < alpha && iabs(p1 - p0) < beta && iabs(q1 - q0) < beta)) continue;
            if (chroma) {
                pix[-xs] = (uint8_t)((2 * p1 + p0 + q1 + 2) >> 2);
                pix[0] = (uint8_t)((2 * q1 + q0 + p1 + 2) >> 2);
                continue;
            }
            const int p2 = pix[-3 * xs], q2 = pix[2 * xs];
            if (iabs(p0 - q0) < ((alpha >> 2) + 2)) {
                if (iabs(p2 - p0) < beta) {
                    const int p3 = pix[-4 * xs];
                    pix[-1 * xs] = (uint8_t)((p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3);
                    pix[-2 * xs] = (uint8_t)((p2 + p1 + p0 + q0 + 2) >> 2);
                    pix[-3 * xs] = (uint8_t)((2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3);
                } else
                    pix[-1 * xs] = (uint8_t)((2 * p1 + p0 + q1 + 2) >> 2);
                if (iabs(q2 - q0) < beta) {
                    const int q3 = pix[3 * xs];
                    pix[0 * xs] = (uint8_t)((p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3);
                    pix[1 * xs] = (uint8_t)((p0 + q0 + q1 + q2 + 2) >> 2);
                    pix[2 * xs] = (uint8_t)((2 * q3 + 3 * q2 + q1 + q0 + p0 + 4) >> 3);
                } else
                    pix[0 * xs] = (uint8_t)((2 * q1 + q0 + p1 + 2) >> 2);
            } else {
                pix[-1 * xs] = (uint8_t)((2 * p1 + p0 + q1 + 2) >> 2);
                pix[0 * xs] = (uint8_t)((2 * q1 + q0 + p1 + 2) >> 2);
            }
        }
    }
    return 0;
}


/* ---- the same members for 9 / 10 / 12 / 14 bit samples (h264dsp_template.c:103-340 with pixel = uint16_t): alpha and beta scaled by
 * << (depth - 8) (:109-110,:172-173,:241-242,:298-299), luma tc0 by * (1 << (depth - 8)) (:112), chroma tc = ((tc0 - 1U) << (depth - 8)) + 1
 * (:246), results clipped to the depth.  stride in BYTES. */
int orc_h264_loop_filter_hbd(int depth, int kind, uint8_t *pix8, ptrdiff_t stride, int alpha, int beta, const int8_t *tc0)
{
    int intra, chroma, vert, iters;
    if (lf_shape(kind, &intra, &chroma, &vert, &iters) < 0 || (depth != 9 && depth != 10 && depth != 12 && depth != 14)) return -1;
    uint16_t *pix = (uint16_t *)pix8;
    const ptrdiff_t st = stride / 2, xs = vert ? st : 1, ys = vert ? 1 : st;
    const int sh = depth - 8, maxv = (1 << depth) - 1;
    alpha <<= sh; beta <<= sh;
    for (int line = 0; line < 4 * iters; line++, pix += ys) {
        const int p0 = pix[-1 * xs], p1 = pix[-2 * xs], q0 = pix[0], q1 = pix[1 * xs];
        if (!intra) {
            const int raw = tc0[line / iters];
            const int t0 = chroma ? (int)(((unsigned)raw - 1U) << sh) + 1 : raw * (1 << sh);
            if (chroma ? t0 <= 0 : t0 < 0) continue;
            if (!(iabs(p0 - q0) < alpha && iabs(p1 - p0) < beta && iabs(q1 - q0) < beta)) continue;
            int tc = t0;
            if (!chroma) {
                const int p2 = pix[-3 * xs], q2 = pix[2 * xs];
                if (iabs(p2 - p0) < beta) {
                    if (t0) pix[-2 * xs] = (uint16_t)(p1 + clip3(((p2 + ((p0 + q0 + 1) >> 1)) >> 1) - p1, -t0, t0));
                    tc++;
                }
                if (iabs(q2 - q0) < beta) {
                    if (t0) pix[xs] = (uint16_t)(q1 + clip3(((q2 + ((p0 + q0 + 1) >> 1)) >> 1) - q1, -t0, t0));
                    tc++;
                }
            }
            const int delta = clip3((((q0 - p0) * 4) + (p1 - q1) + 4) >> 3, -tc, tc);
            pix[-xs] = (uint16_t)clip3(p0 + delta, 0, maxv);
            pix[0] = (uint16_t)clip3(q0 - delta, 0, maxv);
        } else {
            if (!(iabs(p0 - q0) < alpha && iabs(p1 - p0) < beta && iabs(q1 - q0) < beta)) continue;
            if (chroma) {
                pix[-xs] = (uint16_t)((2 * p1 + p0 + q1 + 2) >> 2);
                pix[0] = (uint16_t)((2 * q1 + q0 + p1 + 2) >> 2);
                continue;
            }
            const int p2 = pix[-3 * xs], q2 = pix[2 * xs];
            if (iabs(p0 - q0) < ((alpha >> 2) + 2)) {
                if (iabs(p2 - p0) < beta) {
                    const int p3 = pix[-4 * xs];
                    pix[-1 * xs] = (uint16_t)((p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3);
                    pix[-2 * xs] = (uint16_t)((p2 + p1 + p0 + q0 + 2) >> 2);
                    pix[-3 * xs] = (uint16_t)((2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3);
                } else
                    pix[-1 * xs] = (uint16_t)((2 * p1 + p0 + q1 + 2) >> 2);
                if (iabs(q2 - q0) < beta) {
                    const int q3 = pix[3 * xs];
                    pix[0 * xs] = (uint16_t)((p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3);
                    pix[1 * xs] = (uint16_t)((p0 + q0 + q1 + q2 + 2) >> 2);
                    pix[2 * xs] = (uint16_t)((2 * q3 + 3 * q2 + q1 + q0 + p0 + 4) >> 3);
                } else
                    pix[0 * xs] = (uint16_t)((2 * q1 + q0 + p1 + 2) >> 2);
            } else {
                pix[-1 * xs] = (uint16_t)((2 * p1 + p0 + q1 + 2) >> 2);
                pix[0 * xs] = (uint16_t)((2 * q1 + q0 + p1 + 2) >> 2);
            }
        }
    }
    return 0;
}
