"""Host-side mirror of libavcodec's mpegvideo inverse quantisers (MPVUnquantDSPContext, libavcodec/mpegvideo_unquantize.h:31-44;
functions libavcodec/mpegvideo_unquantize.c:50-276).  The reference's functions take the whole MPVContext; MpvUnquant carries
the fields they read.  Everything goes through the C ABI (b200_mpv_unquantize_batch_device)."""
import ctypes as C
from ._lib import lib, check, vp, MpvUnquant

(UNQUANT_MPEG1_INTRA, UNQUANT_MPEG1_INTER, UNQUANT_MPEG2_INTRA, UNQUANT_MPEG2_INTRA_BITEXACT, UNQUANT_MPEG2_INTER,
 UNQUANT_H263_INTRA, UNQUANT_H263_INTER) = range(7)


def _dptr(x):
    return int(x.data_ptr()) if hasattr(x, "data_ptr") else int(x)


def ff_init_scantable(permutation, scan):
    """ScanTable.permutated / .raster_end (mpegvideo_unquantize.c:36-48)"""
    permutated, raster_end, end = [], [], -1
    for i in range(64):
        j = int(permutation[int(scan[i])])
        permutated.append(j)
        end = max(end, j)
        raster_end.append(end)
    return permutated, raster_end


def unquant_params(intra_matrix, inter_matrix, permutated, raster_end, y_dc_scale, c_dc_scale, q_scale_type=0, h263_aic=0, ac_pred=0):
    p = MpvUnquant()
    for i in range(64):
        p.intra_matrix[i], p.inter_matrix[i] = int(intra_matrix[i]), int(inter_matrix[i])
        p.permutated[i], p.raster_end[i] = int(permutated[i]), int(raster_end[i])
    p.y_dc_scale, p.c_dc_scale, p.q_scale_type, p.h263_aic, p.ac_pred = y_dc_scale, c_dc_scale, q_scale_type, h263_aic, ac_pred
    return p


def unquantize_batch_device(device, variant, params, blocks, nblocks, blk_n, qscale, last_index):
    """in place on nblocks int16[64] device blocks; blk_n None = macroblock stream order (i % 6)"""
    return check(lib().b200_mpv_unquantize_batch_device(device.handle, variant, C.byref(params), vp(_dptr(blocks)), nblocks,
                                                        vp(_dptr(blk_n)) if blk_n is not None else None, vp(_dptr(qscale)),
                                                        vp(_dptr(last_index))), "mpv_unquantize_batch_device")


def unquant_idct_mb420_device(device, variant, params, kind, blocks, qscale, last_index, mb_w, mb_h, nframes, planes, linesize, frame_stride):
    """put_dct / add_dequant_dct (libavcodec/mpegvideo_dec.c:907-922) over a 4:2:0 macroblock stream: inverse quantiser fused into the
    simple IDCT put (kind 1) / add (kind 2); device tensors / pointers"""
    pl = (C.c_void_p * 3)(*[_dptr(x) for x in planes])
    ls = (C.c_int32 * 3)(*linesize)
    fs = (C.c_int64 * 3)(*frame_stride)
    return check(lib().b200_mpv_unquant_idct_mb420_device(device.handle, variant, C.byref(params), kind, vp(_dptr(blocks)), vp(_dptr(qscale)),
                                                          vp(_dptr(last_index)), mb_w, mb_h, nframes, pl, ls, fs), "unquant_idct_mb420_device")
