"""Host-side mirror of libavutil/tx's public interface (av_tx_init / av_tx_fn / av_tx_uninit, libavutil/tx.h:151,202-208)
for the float, int32 and double transform types the library builds, plus the batched device entry points."""
import ctypes as C
from ._lib import lib, check, vp, TX_FN

AV_TX_FLOAT_FFT, AV_TX_FLOAT_MDCT, AV_TX_FLOAT_RDFT, AV_TX_FLOAT_DCT = 0, 1, 6, 9
AV_TX_INT32_FFT, AV_TX_INT32_MDCT = 4, 5
AV_TX_DOUBLE_FFT, AV_TX_DOUBLE_MDCT = 2, 3
AV_TX_INPLACE, AV_TX_UNALIGNED, AV_TX_FULL_IMDCT = 1, 2, 4          # flags (libavutil/tx.h:155-180)


def _dptr(x):
    return int(x.data_ptr()) if hasattr(x, "data_ptr") else int(x)


def _hptr(x):
    return int(x.ctypes.data) if hasattr(x, "ctypes") else _dptr(x)


class AVTXContext:
    def __init__(self, type, inv, len, scale=None, flags=0, device=None):
        self._h = vp()
        self._fn = TX_FN()
        # the scale of the double types is a const double * (libavutil/tx.h:44-58)
        sc = C.byref((C.c_double if type in (AV_TX_DOUBLE_FFT, AV_TX_DOUBLE_MDCT) else C.c_float)(scale)) if scale is not None else None
        if device is None:
            ret = lib().b200_tx_init(C.byref(self._h), C.byref(self._fn), type, inv, len, sc, flags)
        else:
            ret = lib().b200_tx_init_device(device.handle, C.byref(self._h), C.byref(self._fn), type, inv, len, sc, flags)
        check(ret, "av_tx_init")
        self.type, self.inv, self.len = type, inv, len

    def fn(self, out, inp, stride):
        """av_tx_fn on HOST numpy arrays."""
        self._fn(self._h, out.ctypes.data, inp.ctypes.data, stride)

    def batch_device(self, out, inp, stride, count, out_step, in_step):
        return check(lib().b200_tx_batch_device(self._h, vp(_dptr(out)), vp(_dptr(inp)), stride, count, out_step, in_step),
                     "tx_batch_device")

    def batch_host(self, out, inp, stride, count, out_step, in_step):
        """HOST buffers (numpy arrays / pinned torch tensors / raw addresses), pipelined H2D -> kernels -> D2H."""
        return check(lib().b200_tx_batch_host(self._h, vp(_hptr(out)), vp(_hptr(inp)), stride, count, out_step, in_step), "tx_batch_host")

    def uninit(self):
        if self._h:
            lib().b200_tx_uninit(C.byref(self._h))

    def __del__(self):
        try:
            self.uninit()
        except Exception:
            pass


def av_tx_init(type, inv, len, scale=None, flags=0, device=None):
    return AVTXContext(type, inv, len, scale, flags, device)
