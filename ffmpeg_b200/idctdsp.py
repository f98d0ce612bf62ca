"""Host-side mirror of libavcodec's idctdsp interface: IDCTDSPContext as filled by ff_idctdsp_init
(libavcodec/idctdsp.c:228-314), plus the batched entry points.  Everything goes through the C ABI."""
import ctypes as C
import numpy as np
from ._lib import lib, check, vp, u8p, i16p, IDCTDSPContext, B200Error

FF_IDCT_AUTO, FF_IDCT_SIMPLE = 0, 2
IDCT, IDCT_PUT, IDCT_ADD = 0, 1, 2


def _dptr(x):
    if x is None:
        return None
    return int(x.data_ptr()) if hasattr(x, "data_ptr") else int(x)


def ff_idctdsp_init(idct_algo=FF_IDCT_SIMPLE, bits_per_raw_sample=8, lowres=0):
    """Returns the filled pointer table (ctypes struct): c.idct_put(dest, line_size, block) etc. on HOST pointers."""
    c = IDCTDSPContext()
    check(lib().b200_idctdsp_init(C.byref(c), idct_algo, bits_per_raw_sample, lowres), "ff_idctdsp_init")
    return c


def ff_idctdsp_init_hbd(idct_algo=FF_IDCT_SIMPLE, bits_per_raw_sample=10, lowres=0):
    """ff_idctdsp_init for bits_per_raw_sample 9 / 10 / 12 (idctdsp.c:248-266): uint16 pixels, line sizes in bytes."""
    c = IDCTDSPContext()
    check(lib().b200_idctdsp_init_hbd(C.byref(c), idct_algo, bits_per_raw_sample, lowres), "ff_idctdsp_init_hbd")
    return c


def idct_hbd_batch_device(device, depth, kind, blocks, nblocks, dest=None, dest_off=None, line_size=None, uniform_line_size=0):
    return check(lib().b200_idct_hbd_batch_device(device.handle, depth, kind, vp(_dptr(blocks)), nblocks,
                                                  vp(_dptr(dest)) if dest is not None else None,
                                                  vp(_dptr(dest_off)) if dest_off is not None else None,
                                                  vp(_dptr(line_size)) if line_size is not None else None, uniform_line_size),
                 "idct_hbd_batch_device")


def ff_proresdsp_init(bits_per_raw_sample):
    """ProresDSPContext (libavcodec/proresdsp.h): c.idct_put(out, linesize, block, qmat) on HOST pointers"""
    from ._lib import ProresDSPContext
    c = ProresDSPContext()
    check(lib().b200_proresdsp_init(C.byref(c), bits_per_raw_sample), "ff_proresdsp_init")
    return c


def prores_idct_put_batch_device(device, bits, blocks, nblocks, qmat, dest, dest_off, line_size=None, uniform_line_size=0):
    return check(lib().b200_prores_idct_put_batch_device(device.handle, bits, vp(_dptr(blocks)), nblocks, vp(_dptr(qmat)), vp(_dptr(dest)),
                                                         vp(_dptr(dest_off)), vp(_dptr(line_size)) if line_size is not None else None,
                                                         uniform_line_size), "prores_idct_put_batch_device")


def idct_batch_device(device, kind, blocks, nblocks, dest=None, dest_off=None, line_size=None, uniform_line_size=0):
    return check(lib().b200_idct_batch_device(device.handle, kind, vp(_dptr(blocks)), nblocks, vp(_dptr(dest)) if dest is not None else None,
                                              vp(_dptr(dest_off)) if dest_off is not None else None,
                                              vp(_dptr(line_size)) if line_size is not None else None, uniform_line_size),
                 "idct_batch_device")


def _mb420(fn, device, kind, blocks, mb_w, mb_h, nframes, planes, linesize, frame_stride, what):
    pp = (vp * 3)(*[_dptr(p) if not hasattr(p, "ctypes") else p.ctypes.data for p in planes])
    ls = (C.c_int32 * 3)(*linesize)
    fs = (C.c_int64 * 3)(*frame_stride)
    b = blocks.ctypes.data if hasattr(blocks, "ctypes") else _dptr(blocks)
    return check(fn(device.handle, kind, vp(b), mb_w, mb_h, nframes, pp, ls, fs), what)


def idct_mb420_device(device, kind, blocks, mb_w, mb_h, nframes, planes, linesize, frame_stride):
    return _mb420(lib().b200_idct_mb420_device, device, kind, blocks, mb_w, mb_h, nframes, planes, linesize, frame_stride,
                  "idct_mb420_device")


def idct_mb420_host(device, kind, blocks, mb_w, mb_h, nframes, planes, linesize, frame_stride):
    return _mb420(lib().b200_idct_mb420_host, device, kind, blocks, mb_w, mb_h, nframes, planes, linesize, frame_stride,
                  "idct_mb420_host")


# ---------------------------------------------------------------------------------------------- H.264 residual transforms
H264_IDCT4, H264_IDCT8, H264_IDCT4_DC, H264_IDCT8_DC = 0, 1, 2, 3


def ff_h264dsp_idct_init(bit_depth=8, chroma_format_idc=1):
    """The IDCT members of H264DSPContext as ff_h264dsp_init installs them (libavcodec/h264dsp.c:66-139)."""
    from ._lib import H264IDCTContext
    c = H264IDCTContext()
    check(lib().b200_h264_idct_init(C.byref(c), bit_depth, chroma_format_idc), "ff_h264dsp_init (idct)")
    return c


def h264_idct_batch_device(device, kind, n, blocks, blk_off, dst, dst_off, stride):
    return check(lib().b200_h264_idct_batch_device(device.handle, kind, n, vp(_dptr(blocks)), vp(_dptr(blk_off)), vp(_dptr(dst)),
                                                   vp(_dptr(dst_off)), stride), "h264_idct_batch_device")
