"""GPU tier: the fused scaler with the tensor-core horizontal pass (csrc/sws_mma.cuh) against the oracle, bit for bit, over ratios that
take one, two and more 32-column chunks per group, up- and down-scaling, every packed output and the three-plane output; each case
asserts that the tensor-core kernels actually ran (b200_sws_last_path)."""
import numpy as np
import pytest

import cpulibs as cl
from cases import FATE

pytestmark = pytest.mark.gpu

CASES = [(640, 368, 320, 184, FATE), (1920, 1088, 1280, 720, FATE), (704, 576, 1408, 1152, cl.SWS_BICUBIC), (1024, 608, 96, 64, FATE),
         (3840, 2160, 1920, 1080, FATE), (1920, 1080, 640, 352, cl.SWS_AREA)]


def run_batch(device, w, h, dw, dh, fl, fmt, frames):
    import torch
    from ffmpeg_b200 import swscale as sw
    n = len(frames)
    Y, U, V = (np.stack([f[k] for f in frames]) for k in range(3))
    cw, ch = (w + 1) // 2, (h + 1) // 2
    ctx = sw.sws_getContext(device, w, h, 0, dw, dh, fmt, fl)
    with torch.cuda.stream(torch.cuda.ExternalStream(device.stream)):
        dY, dU, dV = torch.from_numpy(Y).cuda(), torch.from_numpy(U).cuda(), torch.from_numpy(V).cuda()
        if fmt == sw.AV_PIX_FMT_YUV420P:
            cdw, cdh = (dw + 1) // 2, (dh + 1) // 2
            o = [torch.zeros((n, dh, dw), dtype=torch.uint8, device="cuda"), torch.zeros((n, cdh, cdw), dtype=torch.uint8, device="cuda"),
                 torch.zeros((n, cdh, cdw), dtype=torch.uint8, device="cuda")]
            ctx.scale_batch_device_planar([dY, dU, dV], [w, cw, cw], [w * h, cw * ch, cw * ch], o, [dw, cdw, cdw], [dw * dh, cdw * cdh, cdw * cdh], n)
            device.sync()
            got = [t.cpu().numpy() for t in o]
        else:
            o = torch.zeros((n, dh, dw * ctx.bpp), dtype=torch.uint8, device="cuda")
            ctx.scale_batch_device([dY, dU, dV], [w, cw, cw], [w * h, cw * ch, cw * ch], o, dw * ctx.bpp, dw * dh * ctx.bpp, n)
            device.sync()
            got = o.cpu().numpy()
    path = ctx.last_path()
    ctx.free()
    return got, path


@pytest.mark.parametrize("case", CASES + [(1280, 720, 1920, 1080, cl.SWS_BILINEAR | 0x40000 | 0x80000)])
def test_mma_planar_vs_oracle(device, case):
    w, h, dw, dh, fl = case
    frames = [cl.yuv_frame(w, h, 40 + i, kind) for i, kind in enumerate(("random", "limited"))]
    got, path = run_batch(device, w, h, dw, dh, fl, 0, frames)
    assert path == 4 or w > 4 * dw, path                         # 10:1: the tiles do not fit in shared memory (two passes)
    for i, fr in enumerate(frames):
        ref = cl.orc_sws_planar(w, h, dw, dh, fl, *fr)
        for k in range(3):
            assert np.array_equal(got[k][i], ref[k]), (case, i, "YUV"[k], int((got[k][i] != ref[k]).sum()))


@pytest.mark.parametrize("case", CASES)
def test_mma_rgb24_vs_oracle(device, case):
    w, h, dw, dh, fl = case
    frames = [cl.yuv_frame(w, h, 50 + i, kind) for i, kind in enumerate(("random", "smooth"))]
    got, path = run_batch(device, w, h, dw, dh, fl, cl.PIX_FMT_RGB24, frames)
    assert path == 4 or (w > 4 * dw and path == 1), path        # a 10:1 tile of three planes does not fit in shared memory: two passes
    for i, fr in enumerate(frames):
        ref = cl.orc_sws(w, h, dw, dh, fl, *fr)
        assert np.array_equal(got[i], ref), (case, i, int((got[i] != ref).sum()))


@pytest.mark.parametrize("name", ["bgr24", "rgba", "bgra", "argb", "abgr"])
def test_mma_other_packed_formats(device, name):
    w, h, dw, dh, fl = 1920, 1088, 1280, 720, FATE
    fmt = cl.PACKED_RGB_FORMATS[name]
    fr = cl.yuv_frame(w, h, 61, "random")
    got, path = run_batch(device, w, h, dw, dh, fl, fmt, [fr])
    assert path == 4, path
    ref = cl.orc_sws(w, h, dw, dh, fl, *fr, fmt=fmt)
    assert np.array_equal(got[0], ref), (name, int((got[0] != ref).sum()))
