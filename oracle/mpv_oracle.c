/*
 * mpv_oracle.c — TEST INFRASTRUCTURE ONLY.  CPU restatement of libavcodec's MPEG-1/2/4/H.263 inverse quantisers, the step in
 * front of the IDCT in mpv_reconstruct_mb (put_dct / add_dequant_dct, libavcodec/mpegvideo_dec.c).
 *
 * Follows (behaviour, not text) libavcodec/mpegvideo_unquantize.c:
 *   :50-79    dct_unquantize_mpeg1_intra_c   DC * dc_scale; AC (|l|*q*m)>>3, made odd by (v-1)|1
 *   :81-109   dct_unquantize_mpeg1_inter_c   ((2|l|+1)*q*m)>>4, made odd
 *   :111-140  dct_unquantize_mpeg2_intra_c   q = non-linear table or 2q; (|l|*q*m)>>4
 *   :142-176  dct_unquantize_mpeg2_intra_bitexact  the same plus mismatch control: block[63] ^= (sum-1)&1
 *   :178-211  dct_unquantize_mpeg2_inter_c   ((2|l|+1)*q*m)>>5 plus mismatch control
 *   :213-247  dct_unquantize_h263_intra_c    l*2q +- ((q-1)|1) in raster order (qadd 0 and no DC scaling with AIC)
 *   :249-276  dct_unquantize_h263_inter_c
 *   :36-48    ff_init_scantable (permutated scan and raster_end)
 * and libavcodec/mpegvideodata.c ff_mpeg2_non_linear_qscale (ISO 13818-2 table 7-6).
 * Results are stored back into int16 (wrap-around) as the reference's assignments do.
 */
#include "oracle.h"
#include <stdlib.h>

static const uint8_t nonlinear_qscale[32] = {                      /* ISO/IEC 13818-2 table 7-6, q_scale_type = 1 */
    0, 1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 14, 16, 18, 20, 22, 24, 28, 32, 36, 40, 44, 48, 52, 56, 64, 72, 80, 88, 96, 104, 112,
};

void orc_mpv_init_scantable(const uint8_t permutation[64], const uint8_t scan[64], uint8_t permutated[64], uint8_t raster_end[64])
{
    int end = -1;
    for (int i = 0; i < 64; i++) {
        int j = permutation[scan[i]];
        permutated[i] = (uint8_t)j;
        if (j > end) end = j;
        raster_end[i] = (uint8_t)end;
    }
}

static int dequant_level(int variant, int level, int q, int m)
{
    int a = level < 0 ? -level : level, v;
    switch (variant) {
    case ORC_UNQUANT_MPEG1_INTRA: v = (int)((unsigned)a * q * m) >> 3; v = (v - 1) | 1; break;
    case ORC_UNQUANT_MPEG1_INTER: v = (int)((unsigned)((a << 1) + 1) * q * m) >> 4; v = (v - 1) | 1; break;
    case ORC_UNQUANT_MPEG2_INTRA:
    case ORC_UNQUANT_MPEG2_INTRA_BITEXACT: v = (int)((unsigned)a * q * m) >> 4; break;
    default: v = (int)((unsigned)((a << 1) + 1) * q * m) >> 5; break;          /* MPEG2_INTER */
    }
    return level < 0 ? -v : v;
}

void orc_mpv_unquantize(int variant, const OrcMpvUnquant *p, int16_t *block, int n, int qscale, int last_index)
{
    const int intra = variant == ORC_UNQUANT_MPEG1_INTRA || variant == ORC_UNQUANT_MPEG2_INTRA ||
                      variant == ORC_UNQUANT_MPEG2_INTRA_BITEXACT || variant == ORC_UNQUANT_H263_INTRA;
    if (variant == ORC_UNQUANT_H263_INTRA || variant == ORC_UNQUANT_H263_INTER) {
        const int qmul = qscale << 1;
        int qadd = (qscale - 1) | 1, ncoef;
        if (intra) {
            if (!p->h263_aic) block[0] = (int16_t)(block[0] * (n < 4 ? p->y_dc_scale : p->c_dc_scale));
            else qadd = 0;
            ncoef = p->ac_pred ? 63 : p->raster_end[last_index];
        } else
            ncoef = p->raster_end[last_index];
        for (int i = intra; i <= ncoef; i++) {
            int level = block[i];
            if (level) block[i] = (int16_t)(level < 0 ? level * qmul - qadd : level * qmul + qadd);
        }
        return;
    }
    const int mpeg2 = variant >= ORC_UNQUANT_MPEG2_INTRA;
    int q = qscale, sum = -1;
    if (mpeg2) q = p->q_scale_type ? nonlinear_qscale[qscale & 31] : qscale << 1;
    const uint16_t *m = intra ? p->intra_matrix : p->inter_matrix;
    if (intra) {
        block[0] = (int16_t)(block[0] * (n < 4 ? p->y_dc_scale : p->c_dc_scale));
        sum += block[0];
    }
    for (int i = intra; i <= last_index; i++) {
        const int j = p->permutated[i];
        int level = block[j];
        if (level) {
            level = dequant_level(variant, level, q, m[j]);
            block[j] = (int16_t)level;
            sum += level;
        }
    }
    if (variant == ORC_UNQUANT_MPEG2_INTRA_BITEXACT || variant == ORC_UNQUANT_MPEG2_INTER)
        block[63] ^= sum & 1;
}

void orc_mpv_unquantize_batch(int variant, const OrcMpvUnquant *p, int16_t *blocks, int64_t nblocks, const uint8_t *blk_n,
                              const uint8_t *qscale, const int8_t *last_index)
{
    for (int64_t b = 0; b < nblocks; b++)
        orc_mpv_unquantize(variant, p, blocks + 64 * b, blk_n ? blk_n[b] : (int)(b % 6), qscale[b], last_index[b]);
}
