"""ctypes binding of libb200dsp.so (the C ABI declared in include/b200dsp.h).

The library is built in-tree by `__graft_entry__.build()` (ffmpeg_b200/csrc/Makefile, nvcc, sm_100a only).
There is no fallback: if the shared object is missing, or no CUDA device is usable, everything here raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libb200dsp.so")

u8p = C.POINTER(C.c_uint8)
i16p = C.POINTER(C.c_int16)
i32p = C.POINTER(C.c_int32)
i64p = C.POINTER(C.c_int64)
vp = C.c_void_p


class B200Error(RuntimeError):
    pass


class IDCTDSPContext(C.Structure):
    """Same member order as IDCTDSPContext, libavcodec/idctdsp.h:43-91."""
    _fields_ = [
        ("put_pixels_clamped", C.CFUNCTYPE(None, i16p, u8p, C.c_ssize_t)),
        ("put_signed_pixels_clamped", C.CFUNCTYPE(None, i16p, u8p, C.c_ssize_t)),
        ("add_pixels_clamped", C.CFUNCTYPE(None, i16p, u8p, C.c_ssize_t)),
        ("idct", C.CFUNCTYPE(None, i16p)),
        ("idct_put", C.CFUNCTYPE(None, u8p, C.c_ssize_t, i16p)),
        ("idct_add", C.CFUNCTYPE(None, u8p, C.c_ssize_t, i16p)),
        ("idct_permutation", C.c_uint8 * 64),
        ("perm_type", C.c_int),
        ("mpeg4_studio_profile", C.c_int),
    ]


_CMP = C.CFUNCTYPE(C.c_int, C.c_void_p, u8p, u8p, C.c_ssize_t, C.c_int)


class MECmpContext(C.Structure):
    """Same member order as MECmpContext, libavcodec/me_cmp.h:53-77."""
    _fields_ = [("sum_abs_dctelem", C.CFUNCTYPE(C.c_int, i16p))] + \
        [(n, _CMP * 6) for n in ("sad", "sse", "hadamard8_diff", "dct_sad", "quant_psnr", "bit", "rd", "vsad", "vsse",
                                 "nsse", "w53", "w97", "dct_max", "dct264_sad")] + \
        [("pix_abs", (_CMP * 4) * 2), ("median_sad", _CMP * 6)]


class FDCTDSPContext(C.Structure):
    """Same member order as FDCTDSPContext, libavcodec/fdctdsp.h:28-31."""
    _fields_ = [("fdct", C.CFUNCTYPE(None, i16p)), ("fdct248", C.CFUNCTYPE(None, i16p))]


_QPEL = C.CFUNCTYPE(None, u8p, u8p, C.c_ssize_t)
_HPEL = C.CFUNCTYPE(None, u8p, u8p, C.c_ssize_t, C.c_int)


class H264QpelContext(C.Structure):
    """libavcodec/h264qpel.h:27-30"""
    _fields_ = [("put_h264_qpel_pixels_tab", (_QPEL * 16) * 3), ("avg_h264_qpel_pixels_tab", (_QPEL * 16) * 3)]


class HpelDSPContext(C.Structure):
    """libavcodec/hpeldsp.h:39-97"""
    _fields_ = [("put_pixels_tab", (_HPEL * 4) * 4), ("avg_pixels_tab", (_HPEL * 4) * 4),
                ("put_no_rnd_pixels_tab", (_HPEL * 4) * 3), ("avg_no_rnd_pixels_tab", _HPEL * 4)]


_H264IDCT = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_ssize_t)


class SwsVector(C.Structure):
    """SwsVector (libswscale/swscale.h:187-192)"""
    _fields_ = [("coeff", C.POINTER(C.c_double)), ("length", C.c_int)]


class SwsFilter(C.Structure):
    """SwsFilter (libswscale/swscale.h:199-204)"""
    _fields_ = [("lumH", C.POINTER(SwsVector)), ("lumV", C.POINTER(SwsVector)), ("chrH", C.POINTER(SwsVector)), ("chrV", C.POINTER(SwsVector))]


class H264IDCTContext(C.Structure):
    """the IDCT members of H264DSPContext, libavcodec/h264dsp.h:81-88"""
    _fields_ = [("idct_add", _H264IDCT), ("idct8_add", _H264IDCT), ("idct_dc_add", _H264IDCT), ("idct8_dc_add", _H264IDCT)]


_H264W = C.CFUNCTYPE(None, C.c_void_p, C.c_ssize_t, C.c_int, C.c_int, C.c_int, C.c_int)
_H264BW = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_ssize_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int)


class H264WeightContext(C.Structure):
    """weight_pixels_tab / biweight_pixels_tab of H264DSPContext, libavcodec/h264dsp.h:33-45"""
    _fields_ = [("weight_pixels_tab", _H264W * 4), ("biweight_pixels_tab", _H264BW * 4)]


class FloatDSPContext(C.Structure):
    """AVFloatDSPContext, libavutil/float_dsp.h:24-210 (same member order)"""
    _F, _D = C.POINTER(C.c_float), C.POINTER(C.c_double)
    _fields_ = [("vector_fmul", C.CFUNCTYPE(None, _F, _F, _F, C.c_int)),
                ("vector_fmac_scalar", C.CFUNCTYPE(None, _F, _F, C.c_float, C.c_int)),
                ("vector_dmac_scalar", C.CFUNCTYPE(None, _D, _D, C.c_double, C.c_int)),
                ("vector_fmul_scalar", C.CFUNCTYPE(None, _F, _F, C.c_float, C.c_int)),
                ("vector_dmul_scalar", C.CFUNCTYPE(None, _D, _D, C.c_double, C.c_int)),
                ("vector_fmul_window", C.CFUNCTYPE(None, _F, _F, _F, _F, C.c_int)),
                ("vector_fmul_add", C.CFUNCTYPE(None, _F, _F, _F, _F, C.c_int)),
                ("vector_fmul_reverse", C.CFUNCTYPE(None, _F, _F, _F, C.c_int)),
                ("butterflies_float", C.CFUNCTYPE(None, _F, _F, C.c_int)),
                ("scalarproduct_float", C.CFUNCTYPE(C.c_float, _F, _F, C.c_int)),
                ("vector_dmul", C.CFUNCTYPE(None, _D, _D, _D, C.c_int)),
                ("scalarproduct_double", C.CFUNCTYPE(C.c_double, _D, _D, C.c_size_t))]


_LF_TC = C.CFUNCTYPE(None, C.c_void_p, C.c_ssize_t, C.c_int, C.c_int, C.c_void_p)
_LF_INTRA = C.CFUNCTYPE(None, C.c_void_p, C.c_ssize_t, C.c_int, C.c_int)


class H264LoopFilterContext(C.Structure):
    """the loop-filter members of H264DSPContext, libavcodec/h264dsp.h:48-73"""
    _fields_ = [("v_loop_filter_luma", _LF_TC), ("h_loop_filter_luma", _LF_TC), ("h_loop_filter_luma_mbaff", _LF_TC),
                ("v_loop_filter_luma_intra", _LF_INTRA), ("h_loop_filter_luma_intra", _LF_INTRA), ("h_loop_filter_luma_mbaff_intra", _LF_INTRA),
                ("v_loop_filter_chroma", _LF_TC), ("h_loop_filter_chroma", _LF_TC), ("h_loop_filter_chroma_mbaff", _LF_TC),
                ("v_loop_filter_chroma_intra", _LF_INTRA), ("h_loop_filter_chroma_intra", _LF_INTRA), ("h_loop_filter_chroma_mbaff_intra", _LF_INTRA)]


class ProresDSPContext(C.Structure):
    """libavcodec/proresdsp.h:28-35"""
    _fields_ = [("idct_permutation_type", C.c_int), ("idct_permutation", C.c_uint8 * 64),
                ("idct_put", C.CFUNCTYPE(None, C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_void_p)),
                ("idct_put_bayer", C.c_void_p)]


class MpvUnquant(C.Structure):
    """B200MpvUnquant: the MPVContext fields the inverse quantisers read (libavcodec/mpegvideo.h:70-77,201-203,258)"""
    _fields_ = [("intra_matrix", C.c_uint16 * 64), ("inter_matrix", C.c_uint16 * 64), ("permutated", C.c_uint8 * 64),
                ("raster_end", C.c_uint8 * 64), ("y_dc_scale", C.c_int32), ("c_dc_scale", C.c_int32), ("q_scale_type", C.c_int32),
                ("h263_aic", C.c_int32), ("ac_pred", C.c_int32)]


_CHROMA = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_ssize_t, C.c_int, C.c_int, C.c_int)


class H264ChromaContext(C.Structure):
    """libavcodec/h264chroma.h:26-31"""
    _fields_ = [("put_h264_chroma_pixels_tab", _CHROMA * 4), ("avg_h264_chroma_pixels_tab", _CHROMA * 4)]


_EDGE = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_ssize_t, C.c_ssize_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int)
_PREFETCH = C.CFUNCTYPE(None, C.c_void_p, C.c_ssize_t, C.c_int)


class VideoDSPContext(C.Structure):
    """libavcodec/videodsp.h:32-69"""
    _fields_ = [("emulated_edge_mc", _EDGE), ("prefetch", _PREFETCH)]


TX_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_void_p, C.c_ssize_t)

# name -> (restype, argtypes): every symbol include/b200dsp.h declares
PROTOTYPES = {
    "b200_abi_version": (C.c_int, []),
    "b200_last_error": (C.c_char_p, []),
    "b200_device_open": (C.c_int, [C.POINTER(vp), C.c_int, vp]),
    "b200_device_close": (None, [vp]),
    "b200_device_sync": (C.c_int, [vp]),
    "b200_device_ordinal": (C.c_int, [vp]),
    "b200_device_stream": (vp, [vp]),
    "b200_device_sm_count": (C.c_int, [vp]),
    "b200_set_default_device": (C.c_int, [vp]),
    "b200_malloc_device": (vp, [vp, C.c_size_t]),
    "b200_free_device": (None, [vp, vp]),
    "b200_malloc_host": (vp, [C.c_size_t]),
    "b200_free_host": (None, [vp]),
    "b200_memcpy_h2d": (C.c_int, [vp, vp, vp, C.c_size_t]),
    "b200_memcpy_d2h": (C.c_int, [vp, vp, vp, C.c_size_t]),
    "b200_launch_count": (C.c_uint64, []),
    "b200_sws_getContext": (vp, [vp] + [C.c_int] * 7),
    "b200_sws_getContext_range": (vp, [vp] + [C.c_int] * 9),
    "b200_sws_getContext_params": (vp, [vp] + [C.c_int] * 9 + [vp]),
    "b200_sws_getContext_filters": (vp, [vp] + [C.c_int] * 9 + [vp, vp, vp]),
    "b200_sws_freeContext": (None, [vp]),
    "b200_sws_setColorspaceDetails": (C.c_int, [vp, i32p, C.c_int, i32p, C.c_int, C.c_int, C.c_int, C.c_int]),
    "b200_sws_scale": (C.c_int, [vp, C.POINTER(vp), i32p, C.c_int, C.c_int, C.POINTER(vp), i32p]),
    "b200_sws_func": (C.c_int, [vp, C.POINTER(vp), i32p, C.c_int, C.c_int, C.POINTER(vp), i32p]),
    "b200_sws_scale_batch_device": (C.c_int, [vp, C.POINTER(vp), i32p, i64p, vp, C.c_int, C.c_int64, C.c_int]),
    "b200_sws_scale_batch_device_planar": (C.c_int, [vp, C.POINTER(vp), C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(vp), C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.c_int]),
    "b200_sws_scale_batch_host": (C.c_int, [vp, C.POINTER(vp), i32p, i64p, vp, C.c_int, C.c_int64, C.c_int]),
    "b200_sws_info": (C.c_int, [vp, i32p]),
    "b200_sws_get_filter": (C.c_int, [vp, C.c_int, i16p, i32p, C.c_int]),
    "b200_sws_last_path": (C.c_int, [vp]),
    "b200_sws_plan_probe": (C.c_int, [C.c_int] * 6 + [i16p, i32p, C.c_int, i32p]),
    "b200_sws_plan_probe2": (C.c_int, [i32p, i32p, C.c_int, i16p, i32p, C.c_int, i32p]),
    "b200_sws_mma_probe": (C.c_int, [i16p, i32p, C.c_int, C.c_int, i32p, C.c_int, C.POINTER(C.c_uint32), C.c_int, i32p]),
    "b200_idctdsp_init": (C.c_int, [C.POINTER(IDCTDSPContext), C.c_int, C.c_int, C.c_int]),
    "b200_idct_batch_device": (C.c_int, [vp, C.c_int, vp, C.c_int64, vp, vp, vp, C.c_int]),
    "b200_idct_mb420_device": (C.c_int, [vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.POINTER(vp), i32p, i64p]),
    "b200_idct_mb420_host": (C.c_int, [vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.POINTER(vp), i32p, i64p]),
    "b200_me_cmp_init": (C.c_int, [C.POINTER(MECmpContext), C.c_int]),
    "b200_mpv_unquant_idct_mb420_device": (C.c_int, [vp, C.c_int, vp, C.c_int, vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, i32p, i64p]),
    "b200_me_cmp_set_nsse_weight": (None, [C.c_int]),
    "b200_me_cmp_set_dct_algo": (C.c_int, [C.c_int]),
    "b200_sum_abs_dctelem_batch_device": (C.c_int, [vp, vp, C.c_int64, vp]),
    "b200_fdctdsp_init": (C.c_int, [C.POINTER(FDCTDSPContext), C.c_int, C.c_int]),
    "b200_fdct_batch_device": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, vp, C.c_int64]),
    "b200_me_cmp_batch_device": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, C.c_ssize_t, C.c_int, vp, vp, C.c_int64, vp]),
    "b200_me_esa_device": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int, vp, vp]),
    "b200_h264qpel_init": (C.c_int, [C.POINTER(H264QpelContext), C.c_int]),
    "b200_hpeldsp_init": (C.c_int, [C.POINTER(HpelDSPContext), C.c_int]),
    "b200_h264qpel_batch_device": (C.c_int, [vp, C.c_int64, vp, vp, vp, vp, vp, C.c_ssize_t]),
    "b200_h264qpel_hbd_batch_device": (C.c_int, [vp, C.c_int, C.c_int64, vp, vp, vp, vp, vp, C.c_ssize_t]),
    "b200_h264_loop_filter_hbd_batch_device": (C.c_int, [vp, C.c_int, C.c_int64, vp, vp, vp, C.c_ssize_t, vp, vp, vp]),
    "b200_h264_idct_hbd_batch_device": (C.c_int, [vp, C.c_int, C.c_int, C.c_int64, vp, vp, vp, vp, C.c_ssize_t]),
    "b200_h264_weight_hbd_batch_device": (C.c_int, [vp, C.c_int, C.c_int64, vp, vp, vp, vp, vp, C.c_ssize_t]),
    "b200_h264chroma_hbd_batch_device": (C.c_int, [vp, C.c_int64, vp, vp, vp, vp, vp, vp, vp, C.c_ssize_t]),
    "b200_emulated_edge_mc_hbd_batch_device": (C.c_int, [vp, C.c_int64, vp, vp, C.c_ssize_t, vp, vp, C.c_ssize_t, vp, C.c_int, C.c_int]),
    "b200_hpel_batch_device": (C.c_int, [vp, C.c_int64, vp, vp, vp, vp, vp, vp, C.c_ssize_t]),
    "b200_h264_idct_init": (C.c_int, [C.POINTER(H264IDCTContext), C.c_int, C.c_int]),
    "b200_h264_idct_batch_device": (C.c_int, [vp, C.c_int, C.c_int64, vp, vp, vp, vp, C.c_ssize_t]),
    "b200_h264_weight_init": (C.c_int, [C.POINTER(H264WeightContext), C.c_int]),
    "b200_h264_weight_batch_device": (C.c_int, [vp, C.c_int64, vp, vp, vp, vp, vp, C.c_ssize_t]),
    "b200_mpv_unquantize_batch_device": (C.c_int, [vp, C.c_int, vp, vp, C.c_int64, vp, vp, vp]),
    "b200_idctdsp_init_hbd": (C.c_int, [vp, C.c_int, C.c_int, C.c_int]),
    "b200_idct_hbd_batch_device": (C.c_int, [vp, C.c_int, C.c_int, vp, C.c_int64, vp, vp, vp, C.c_int]),
    "b200_tx_pfa_tables": (C.c_int, [C.c_int, C.c_int, C.c_float, vp, C.c_int, vp]),
    "b200_tx_dct_table": (C.c_int, [C.c_int, C.c_int, vp, C.c_int]),
    "b200_tx_i32_tables": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_float, vp, C.c_int, vp]),
    "b200_pixelutils_get_sad_fn": (vp, [C.c_int, C.c_int, C.c_int, vp]),
    "b200_pixelutils_sad_batch_device": (C.c_int, [vp, C.c_int, vp, C.c_ssize_t, vp, C.c_ssize_t, vp, vp, C.c_int64, vp]),
    "b200_h264_loop_filter_init": (C.c_int, [vp, C.c_int, C.c_int]),
    "b200_h264_loop_filter_batch_device": (C.c_int, [vp, C.c_int64, vp, vp, vp, C.c_ssize_t, vp, vp, vp]),
    "b200_proresdsp_init": (C.c_int, [vp, C.c_int]),
    "b200_prores_idct_put_batch_device": (C.c_int, [vp, C.c_int, vp, C.c_int64, vp, vp, vp, vp, C.c_int]),
    "b200_float_dsp_init": (C.c_int, [vp]),
    "b200_float_dsp_batch_device": (C.c_int, [vp, C.c_int, C.c_int64, C.c_int, vp, C.c_int64, vp, C.c_int64, vp, C.c_int64, vp, C.c_int64, C.c_double]),
    "b200_videodsp_init": (C.c_int, [C.POINTER(VideoDSPContext), C.c_int]),
    "b200_emulated_edge_mc_batch_device": (C.c_int, [vp, C.c_int64, vp, vp, C.c_ssize_t, vp, vp, C.c_ssize_t, vp, C.c_int, C.c_int]),
    "b200_h264chroma_init": (C.c_int, [C.POINTER(H264ChromaContext), C.c_int]),
    "b200_h264chroma_batch_device": (C.c_int, [vp, C.c_int64, vp, vp, vp, vp, vp, vp, vp, C.c_ssize_t]),
    "b200_tx_init": (C.c_int, [C.POINTER(vp), C.POINTER(TX_FN), C.c_int, C.c_int, C.c_int, vp, C.c_uint64]),
    "b200_tx_init_device": (C.c_int, [vp, C.POINTER(vp), C.POINTER(TX_FN), C.c_int, C.c_int, C.c_int, vp, C.c_uint64]),
    "b200_tx_uninit": (None, [C.POINTER(vp)]),
    "b200_tx_batch_device": (C.c_int, [vp, vp, vp, C.c_ssize_t, C.c_int64, C.c_ssize_t, C.c_ssize_t]),
    "b200_tx_r16_plan": (C.c_int, [C.c_int, C.c_int, vp, C.c_int]),
    "b200_tx_batch_host": (C.c_int, [vp, vp, vp, C.c_ssize_t, C.c_int64, C.c_ssize_t, C.c_ssize_t]),
    "b200_h264qpel_frames_host": (C.c_int, [vp, C.c_int, C.c_int64, vp, vp, vp, vp, vp, vp, C.c_ssize_t]),
    "b200_me_esa_host": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int, vp, vp]),
}

_lib = None


def lib():
    """Load libb200dsp.so (once).  Raises B200Error when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise B200Error(f"{SO_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(nvcc, sm_100a).  There is no CPU fallback.")
        L = C.CDLL(SO_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)            # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(ret, what):
    if ret is None or (isinstance(ret, int) and ret < 0):
        msg = lib().b200_last_error()
        raise B200Error(f"{what} failed ({ret}): {msg.decode() if msg else ''}")
    return ret
