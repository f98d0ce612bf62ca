/*
 * fdsp_oracle.c — TEST INFRASTRUCTURE ONLY.  CPU restatement of libavutil's AVFloatDSPContext C functions, the element-wise
 * float work around the transforms (vector_fmul_window follows every iMDCT in AAC, libavcodec/aac/aacdec_dsp_template.c).
 *
 * Follows (behaviour, not text) libavutil/float_dsp.c:27-141 and libavutil/float_scalarproduct.c:25-33:
 *   vector_fmul / vector_dmul        dst[i] = src0[i] * src1[i]
 *   vector_fmac_scalar / dmac        dst[i] += src[i] * mul            (product rounded, then the sum rounded: no fused multiply-add)
 *   vector_fmul_scalar / dmul        dst[i] = src[i] * mul
 *   vector_fmul_window               dst[i] = s0*wj - s1*wi, dst[j] = s0*wi + s1*wj over the 2*len window, i from the middle out
 *   vector_fmul_add                  dst[i] = src0[i] * src1[i] + src2[i]
 *   vector_fmul_reverse              dst[i] = src0[i] * src1[len-1-i]
 *   butterflies_float                (v1, v2) <- (v1 + v2, v1 - v2)
 *   scalarproduct_float / double     sum accumulated left to right in the element type
 * Built with -ffp-contract=off (oracle/Makefile), like the generic-C build of the reference.
 */
#include "oracle.h"

int orc_float_dsp(int op, void *dst_, const void *src0_, const void *src1_, const void *src2_, double mul, int len)
{
    float *dst = dst_; const float *src0 = src0_, *src1 = src1_, *src2 = src2_;
    double *ddst = dst_; const double *dsrc0 = src0_, *dsrc1 = src1_;
    const float fmul = (float)mul;
    switch (op) {
    case ORC_FDSP_VECTOR_FMUL:        for (int i = 0; i < len; i++) dst[i] = src0[i] * src1[i]; return 0;
    case ORC_FDSP_VECTOR_FMAC_SCALAR: for (int i = 0; i < len; i++) dst[i] += src0[i] * fmul; return 0;
    case ORC_FDSP_VECTOR_DMAC_SCALAR: for (int i = 0; i < len; i++) ddst[i] += dsrc0[i] * mul; return 0;
    case ORC_FDSP_VECTOR_FMUL_SCALAR: for (int i = 0; i < len; i++) dst[i] = src0[i] * fmul; return 0;
    case ORC_FDSP_VECTOR_DMUL_SCALAR: for (int i = 0; i < len; i++) ddst[i] = dsrc0[i] * mul; return 0;
    case ORC_FDSP_VECTOR_FMUL_WINDOW:                                  /* src2 = window of 2*len floats, dst 2*len floats */
        for (int i = len - 1, j = len; i >= 0; i--, j++) {
            float s0 = src0[i], s1 = src1[j - len], wi = src2[i], wj = src2[j];
            dst[i] = s0 * wj - s1 * wi;
            dst[j] = s0 * wi + s1 * wj;
        }
        return 0;
    case ORC_FDSP_VECTOR_FMUL_ADD:     for (int i = 0; i < len; i++) dst[i] = src0[i] * src1[i] + src2[i]; return 0;
    case ORC_FDSP_VECTOR_FMUL_REVERSE: for (int i = 0; i < len; i++) dst[i] = src0[i] * src1[len - 1 - i]; return 0;
    case ORC_FDSP_BUTTERFLIES_FLOAT: {                                 /* dst = v1, src0 = v2 (both updated) */
        float *v2 = (float *)src0_;
        for (int i = 0; i < len; i++) { float t = dst[i] - v2[i]; dst[i] += v2[i]; v2[i] = t; }
        return 0;
    }
    case ORC_FDSP_SCALARPRODUCT_FLOAT: { float p = 0.0f; for (int i = 0; i < len; i++) p += src0[i] * src1[i]; dst[0] = p; return 0; }
    case ORC_FDSP_VECTOR_DMUL:         for (int i = 0; i < len; i++) ddst[i] = dsrc0[i] * dsrc1[i]; return 0;
    case ORC_FDSP_SCALARPRODUCT_DOUBLE: { double p = 0.0; for (int i = 0; i < len; i++) p += dsrc0[i] * dsrc1[i]; ddst[0] = p; return 0; }
    }
    return -1;
}
