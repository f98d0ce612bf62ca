"""Host-side mirror of libavcodec's fdctdsp interface (FDCTDSPContext as filled by ff_fdctdsp_init, libavcodec/fdctdsp.c:27-45)."""
import ctypes as C
from ._lib import lib, check, vp, FDCTDSPContext

FF_DCT_AUTO, FF_DCT_FASTINT, FF_DCT_INT, FF_DCT_FAAN = 0, 1, 2, 6              # AVCodecContext.dct_algo (libavcodec/avcodec.h:1531-1537)


def ff_fdctdsp_init(dct_algo=FF_DCT_AUTO, bits_per_raw_sample=8):
    """the table of HOST-pointer functions (fdct, fdct248: 64 int16 in place)"""
    c = FDCTDSPContext()
    check(lib().b200_fdctdsp_init(C.byref(c), dct_algo, bits_per_raw_sample), "ff_fdctdsp_init")
    return c


def fdct_batch_device(device, blocks, n, dct_algo=FF_DCT_AUTO, bits_per_raw_sample=8, is248=False):
    """n blocks of 64 int16 on the DEVICE, transformed in place"""
    p = int(blocks.data_ptr()) if hasattr(blocks, "data_ptr") else int(blocks)
    return check(lib().b200_fdct_batch_device(device.handle, dct_algo, bits_per_raw_sample, int(bool(is248)), vp(p), n), "fdct_batch_device")
