"""Device handle: the role AVHWDeviceContext / AVCUDADeviceContext {CUcontext, CUstream} plays in the reference
(libavutil/hwcontext_cuda.h): one GPU ordinal + one stream that orders all work of the contexts created on it."""
import ctypes as C
from ._lib import lib, check, vp, B200Error


class Device:
    def __init__(self, ordinal=0, stream=None):
        """stream: raw CUstream/cudaStream_t handle (int) to share (e.g. torch.cuda.current_stream().cuda_stream)."""
        h = vp()
        check(lib().b200_device_open(C.byref(h), int(ordinal), vp(stream) if stream else None), "b200_device_open")
        self._h = h
        self.ordinal = int(ordinal)

    @property
    def handle(self):
        if not self._h:
            raise B200Error("device is closed")
        return self._h

    @property
    def stream(self):
        return lib().b200_device_stream(self.handle)

    @property
    def sm_count(self):
        return lib().b200_device_sm_count(self.handle)

    def sync(self):
        check(lib().b200_device_sync(self.handle), "b200_device_sync")

    def set_default(self):
        check(lib().b200_set_default_device(self.handle), "b200_set_default_device")

    def close(self):
        if self._h:
            lib().b200_device_close(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def launch_count():
    return int(lib().b200_launch_count())
