/* TEST INFRASTRUCTURE — compile-time proof that the function tables declared in include/b200dsp.h can be copied member for member
 * into the reference's own context structs: this unit includes the reference's headers (where they lie under $(REF)) next to
 * b200dsp.h and asserts sizes, member offsets and function-pointer types.  It holds no code; it is compiled with libffref.so
 * (oracle/ref/Makefile), so a drift on either side breaks the oracle build, and tests/test_abi.py reads the marker below. */
#include <stddef.h>
#include "libavcodec/idctdsp.h"
#include "libavcodec/me_cmp.h"
#include "libavcodec/h264qpel.h"
#include "libavcodec/hpeldsp.h"
#include "libavcodec/h264chroma.h"
#include "libavcodec/videodsp.h"
#include "libavcodec/h264dsp.h"
#include "libavcodec/proresdsp.h"
#include "libavutil/float_dsp.h"
#include "libavutil/pixelutils.h"
#include "libavutil/tx.h"
#include "../../include/b200dsp.h"

#define SAME_SIZE(A, B)        _Static_assert(sizeof(A) == sizeof(B), "size of " #A " / " #B)
#define SAME_OFF(A, B, m)      _Static_assert(offsetof(A, m) == offsetof(B, m), "offset of " #m " in " #A " / " #B)
/* the two members have the same type (function pointer types must agree argument for argument) */
#define SAME_TYPE(A, B, m)     _Static_assert(__builtin_types_compatible_p(__typeof__(((A *)0)->m), __typeof__(((B *)0)->m)), "type of " #m " in " #A " / " #B)
#define SAME(A, B, m)          SAME_OFF(A, B, m); SAME_TYPE(A, B, m)
/* consecutive members of a larger reference struct against a table of ours: same distance from the first member */
#define SAME_REL(A, a0, B, b0, m) _Static_assert(offsetof(A, m) - offsetof(A, a0) == offsetof(B, m) - offsetof(B, b0), "relative offset of " #m); SAME_TYPE(A, B, m)

/* IDCTDSPContext, libavcodec/idctdsp.h:43-91 */
SAME_SIZE(IDCTDSPContext, B200IDCTDSPContext);
SAME(IDCTDSPContext, B200IDCTDSPContext, put_pixels_clamped);
SAME(IDCTDSPContext, B200IDCTDSPContext, put_signed_pixels_clamped);
SAME(IDCTDSPContext, B200IDCTDSPContext, add_pixels_clamped);
SAME(IDCTDSPContext, B200IDCTDSPContext, idct);
SAME(IDCTDSPContext, B200IDCTDSPContext, idct_put);
SAME(IDCTDSPContext, B200IDCTDSPContext, idct_add);
SAME(IDCTDSPContext, B200IDCTDSPContext, idct_permutation);
SAME_OFF(IDCTDSPContext, B200IDCTDSPContext, perm_type);
SAME_OFF(IDCTDSPContext, B200IDCTDSPContext, mpeg4_studio_profile);

/* FDCTDSPContext, libavcodec/fdctdsp.h */
#include "libavcodec/fdctdsp.h"
SAME_SIZE(FDCTDSPContext, B200FDCTDSPContext);
SAME(FDCTDSPContext, B200FDCTDSPContext, fdct);
SAME(FDCTDSPContext, B200FDCTDSPContext, fdct248);

/* ProresDSPContext, libavcodec/proresdsp.h */
SAME_SIZE(ProresDSPContext, B200ProresDSPContext);
SAME(ProresDSPContext, B200ProresDSPContext, idct_permutation_type);
SAME(ProresDSPContext, B200ProresDSPContext, idct_permutation);
SAME(ProresDSPContext, B200ProresDSPContext, idct_put);
SAME(ProresDSPContext, B200ProresDSPContext, idct_put_bayer);

/* MECmpContext, libavcodec/me_cmp.h:53-77 (the first argument of me_cmp_func is a struct pointer there, void * here) */
SAME_SIZE(MECmpContext, B200MECmpContext);
SAME_OFF(MECmpContext, B200MECmpContext, sad);
SAME_OFF(MECmpContext, B200MECmpContext, sse);
SAME_OFF(MECmpContext, B200MECmpContext, hadamard8_diff);
SAME_OFF(MECmpContext, B200MECmpContext, dct_sad);
SAME_OFF(MECmpContext, B200MECmpContext, quant_psnr);
SAME_OFF(MECmpContext, B200MECmpContext, bit);
SAME_OFF(MECmpContext, B200MECmpContext, rd);
SAME_OFF(MECmpContext, B200MECmpContext, vsad);
SAME_OFF(MECmpContext, B200MECmpContext, vsse);
SAME_OFF(MECmpContext, B200MECmpContext, nsse);
SAME_OFF(MECmpContext, B200MECmpContext, w53);
SAME_OFF(MECmpContext, B200MECmpContext, w97);
SAME_OFF(MECmpContext, B200MECmpContext, dct_max);
SAME_OFF(MECmpContext, B200MECmpContext, dct264_sad);
SAME_OFF(MECmpContext, B200MECmpContext, pix_abs);
SAME_OFF(MECmpContext, B200MECmpContext, median_sad);

/* H264QpelContext (h264qpel.h:27-30), HpelDSPContext (hpeldsp.h:39-97), H264ChromaContext (h264chroma.h), VideoDSPContext (videodsp.h) */
SAME_SIZE(H264QpelContext, B200H264QpelContext);
SAME(H264QpelContext, B200H264QpelContext, put_h264_qpel_pixels_tab);
SAME(H264QpelContext, B200H264QpelContext, avg_h264_qpel_pixels_tab);
SAME_SIZE(HpelDSPContext, B200HpelDSPContext);
SAME(HpelDSPContext, B200HpelDSPContext, put_pixels_tab);
SAME(HpelDSPContext, B200HpelDSPContext, avg_pixels_tab);
SAME(HpelDSPContext, B200HpelDSPContext, put_no_rnd_pixels_tab);
SAME(HpelDSPContext, B200HpelDSPContext, avg_no_rnd_pixels_tab);
SAME_SIZE(H264ChromaContext, B200H264ChromaContext);
SAME(H264ChromaContext, B200H264ChromaContext, put_h264_chroma_pixels_tab);
SAME(H264ChromaContext, B200H264ChromaContext, avg_h264_chroma_pixels_tab);
SAME_SIZE(VideoDSPContext, B200VideoDSPContext);
SAME(VideoDSPContext, B200VideoDSPContext, emulated_edge_mc);
SAME(VideoDSPContext, B200VideoDSPContext, prefetch);

/* H264DSPContext (h264dsp.h:42-118): three runs of consecutive members */
SAME_REL(H264DSPContext, weight_pixels_tab, B200H264WeightContext, weight_pixels_tab, weight_pixels_tab);
SAME_REL(H264DSPContext, weight_pixels_tab, B200H264WeightContext, weight_pixels_tab, biweight_pixels_tab);
_Static_assert(sizeof(B200H264WeightContext) == offsetof(H264DSPContext, v_loop_filter_luma) - offsetof(H264DSPContext, weight_pixels_tab), "the two weight tables and nothing else");
#define LF(m) _Static_assert(offsetof(H264DSPContext, m) - offsetof(H264DSPContext, v_loop_filter_luma) == offsetof(B200H264LoopFilterContext, m), \
                             "loop filter member " #m); \
              _Static_assert(__builtin_types_compatible_p(__typeof__(((H264DSPContext *)0)->m), __typeof__(((B200H264LoopFilterContext *)0)->m)), "type of " #m)
LF(v_loop_filter_luma); LF(h_loop_filter_luma); LF(h_loop_filter_luma_mbaff);
LF(v_loop_filter_luma_intra); LF(h_loop_filter_luma_intra); LF(h_loop_filter_luma_mbaff_intra);
LF(v_loop_filter_chroma); LF(h_loop_filter_chroma); LF(h_loop_filter_chroma_mbaff);
LF(v_loop_filter_chroma_intra); LF(h_loop_filter_chroma_intra); LF(h_loop_filter_chroma_mbaff_intra);
_Static_assert(sizeof(B200H264LoopFilterContext) == 12 * sizeof(void *), "twelve loop filter members");
#define ID(m) _Static_assert(__builtin_types_compatible_p(__typeof__(((H264DSPContext *)0)->m), b200_h264_idct_fn), "type of " #m)
ID(idct_add); ID(idct8_add); ID(idct_dc_add); ID(idct8_dc_add);
_Static_assert(offsetof(H264DSPContext, idct8_add) - offsetof(H264DSPContext, idct_add) == offsetof(B200H264IDCTContext, idct8_add) &&
               offsetof(H264DSPContext, idct_dc_add) - offsetof(H264DSPContext, idct_add) == offsetof(B200H264IDCTContext, idct_dc_add) &&
               offsetof(H264DSPContext, idct8_dc_add) - offsetof(H264DSPContext, idct_add) == offsetof(B200H264IDCTContext, idct8_dc_add),
               "the four residual-add members are consecutive");

/* AVFloatDSPContext, libavutil/float_dsp.h:24-210 */
SAME_SIZE(AVFloatDSPContext, B200FloatDSPContext);
SAME(AVFloatDSPContext, B200FloatDSPContext, vector_fmul);
SAME(AVFloatDSPContext, B200FloatDSPContext, vector_fmac_scalar);
SAME(AVFloatDSPContext, B200FloatDSPContext, vector_dmac_scalar);
SAME(AVFloatDSPContext, B200FloatDSPContext, vector_fmul_scalar);
SAME(AVFloatDSPContext, B200FloatDSPContext, vector_dmul_scalar);
SAME(AVFloatDSPContext, B200FloatDSPContext, vector_fmul_window);
SAME(AVFloatDSPContext, B200FloatDSPContext, vector_fmul_add);
SAME(AVFloatDSPContext, B200FloatDSPContext, vector_fmul_reverse);
SAME(AVFloatDSPContext, B200FloatDSPContext, butterflies_float);
SAME(AVFloatDSPContext, B200FloatDSPContext, scalarproduct_float);
SAME(AVFloatDSPContext, B200FloatDSPContext, vector_dmul);
SAME(AVFloatDSPContext, B200FloatDSPContext, scalarproduct_double);

/* public function types: av_pixelutils_sad_fn (pixelutils.h:31-34); av_tx_fn (tx.h:151) up to the context type; the tx flags */
_Static_assert(__builtin_types_compatible_p(av_pixelutils_sad_fn, b200_pixelutils_sad_fn), "av_pixelutils_sad_fn");
_Static_assert(AV_TX_INPLACE == B200_TX_INPLACE && AV_TX_UNALIGNED == B200_TX_UNALIGNED && AV_TX_FULL_IMDCT == B200_TX_FULL_IMDCT, "tx flags");
_Static_assert(AV_TX_FLOAT_FFT == B200_TX_FLOAT_FFT && AV_TX_FLOAT_MDCT == B200_TX_FLOAT_MDCT && AV_TX_FLOAT_RDFT == B200_TX_FLOAT_RDFT &&
               AV_TX_FLOAT_DCT == B200_TX_FLOAT_DCT && AV_TX_INT32_FFT == B200_TX_INT32_FFT && AV_TX_INT32_MDCT == B200_TX_INT32_MDCT, "tx types");

const char ffref_layout_check[] = "b200dsp.h tables match the reference's context structs";
