"""Host-side mirror of libavutil's AVFloatDSPContext (libavutil/float_dsp.h:24-210, avpriv_float_dsp_alloc float_dsp.c:143-168).
Everything goes through the C ABI: the function table takes host pointers, float_dsp_batch_device device pointers."""
import ctypes as C
from ._lib import lib, check, vp, FloatDSPContext

(VECTOR_FMUL, VECTOR_FMAC_SCALAR, VECTOR_DMAC_SCALAR, VECTOR_FMUL_SCALAR, VECTOR_DMUL_SCALAR, VECTOR_FMUL_WINDOW, VECTOR_FMUL_ADD,
 VECTOR_FMUL_REVERSE, BUTTERFLIES_FLOAT, SCALARPRODUCT_FLOAT, VECTOR_DMUL, SCALARPRODUCT_DOUBLE) = range(12)


def _dptr(x):
    return None if x is None else vp(int(x.data_ptr()) if hasattr(x, "data_ptr") else int(x))


def avpriv_float_dsp_alloc(bit_exact=0):
    c = FloatDSPContext()
    check(lib().b200_float_dsp_init(C.byref(c)), "float_dsp_init")
    return c


def float_dsp_batch_device(device, op, nvec, length, dst, dst_stride, src0, src0_stride, src1=None, src1_stride=0, src2=None,
                           src2_stride=0, mul=0.0):
    """strides in elements between consecutive vectors of an operand (0 = shared)"""
    return check(lib().b200_float_dsp_batch_device(device.handle, op, nvec, length, _dptr(dst), dst_stride, _dptr(src0), src0_stride,
                                                   _dptr(src1), src1_stride, _dptr(src2), src2_stride, mul), "float_dsp_batch_device")
