"""GPU tier: the CUDA swscale path (through the C ABI) against the oracle, the golden fixtures and — at full 4K batch
size — against size-independent properties."""
import hashlib
import os

import numpy as np
import pytest

import cpulibs as cl
from cases import SWS_SMALL_CASES, FATE

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def gpu_sws(device, w, h, dw, dh, fl, y, u, v, dst_pad=0, colorspace=None, fmt=cl.PIX_FMT_RGB24, src_fmt=0):
    from ffmpeg_b200 import swscale as sw
    ctx = sw.sws_getContext(device, w, h, src_fmt, dw, dh, fmt, fl)
    try:
        if colorspace is not None:
            ctx.setColorspaceDetails(cl.COEFFS[colorspace[0]], colorspace[1], cl.COEFFS[colorspace[2]], colorspace[3],
                                     colorspace[4], colorspace[5], colorspace[6])
        return ctx.convert(y, u, v, dst_pad=dst_pad)
    finally:
        ctx.free()


def test_golden_small(device):
    g = np.load(os.path.join(G, "sws_small.npz"))
    for i, (w, h, dw, dh, fl, kind) in enumerate(SWS_SMALL_CASES):
        out = gpu_sws(device, w, h, dw, dh, fl, g[f"c{i}_y"], g[f"c{i}_u"], g[f"c{i}_v"])
        ref = g[f"c{i}_rgb"]
        assert np.array_equal(out, ref), (i, w, h, dw, dh, hex(fl), int((out != ref).sum()))


def test_golden_hashes(device):
    for line in open(os.path.join(G, "sws_hashes.txt")):
        i, w, h, dw, dh, fl, kind, hin, hout = line.split()
        i, w, h, dw, dh, fl = map(int, (i, w, h, dw, dh, fl))
        y, u, v = cl.yuv_frame(w, h, 200 + i, kind)
        out = gpu_sws(device, w, h, dw, dh, fl, y, u, v)
        assert sha(out) == hout, (i, w, h, dw, dh, hex(fl))


def test_colorspace_golden(device):
    g = np.load(os.path.join(G, "sws_colorspace.npz"))
    y, u, v = g["y"], g["u"], g["v"]
    css = [(1, 0, 1, 0, 0, 1 << 16, 1 << 16), (5, 1, 5, 1, 0, 1 << 16, 1 << 16),
           (9, 0, 9, 0, 3000, 70000, 80000), (7, 1, 7, 0, -2000, 60000, 50000)]
    for j, cs in enumerate(css):
        for k, fl in enumerate([FATE, cl.SWS_BICUBIC]):
            assert np.array_equal(gpu_sws(device, 64, 48, 64, 48, fl, y, u, v, colorspace=cs), g[f"cs{j}_{k}"]), (j, k)
            assert np.array_equal(gpu_sws(device, 64, 48, 96, 80, fl, y, u, v, colorspace=cs), g[f"cs{j}_{k}_s"]), (j, k)


@pytest.mark.parametrize("case", [
    (352, 288, 352, 288, FATE), (352, 288, 352, 288, cl.SWS_BICUBIC), (352, 288, 200, 100, FATE),
    (640, 360, 640, 360, FATE), (640, 360, 640, 360, cl.SWS_BICUBIC), (1280, 720, 1920, 1080, FATE),
    (100, 50, 37, 21, FATE), (350, 288, 350, 288, cl.SWS_BICUBIC), (346, 286, 346, 286, FATE),
    (352, 288, 640, 360, cl.SWS_BILINEAR), (352, 288, 300, 200, cl.SWS_BICUBLIN), (352, 288, 176, 144, cl.SWS_AREA),
    (3840, 2160, 3840, 2160, FATE), (3840, 2160, 3840, 2160, cl.SWS_BICUBIC), (3840, 2160, 1920, 1080, FATE),
])
def test_vs_oracle_padded_strides(device, case):
    """Seeded inputs with non-multiple-of-16 strides (scalar paths) and 16-aligned strides (vector paths)."""
    w, h, dw, dh, fl = case
    for seed, kind, pad, dpad in ((1, "random", 0, 0), (2, "limited", 7, 5), (3, "smooth", 16, 16)):
        if w >= 3840 and seed == 3:
            continue
        y, u, v = cl.yuv_frame(w, h, seed, kind, pad=pad)
        a = cl.orc_sws(w, h, dw, dh, fl, y, u, v, dst_pad=dpad)
        b = gpu_sws(device, w, h, dw, dh, fl, y, u, v, dst_pad=dpad)
        assert np.array_equal(a, b), (case, seed, int((a != b).sum()))


def test_other_formats_golden_hashes(device):
    """bgr24 / rgba / bgra / argb / abgr against the reference's outputs (sws_format_hashes.txt)."""
    n = 0
    for line in open(os.path.join(G, "sws_format_hashes.txt")):
        name, i, w, h, dw, dh, fl, kind, hout = line.split()
        i, w, h, dw, dh, fl = map(int, (i, w, h, dw, dh, fl))
        y, u, v = cl.yuv_frame(w, h, 500 + i, kind)
        out = gpu_sws(device, w, h, dw, dh, fl, y, u, v, fmt=cl.PACKED_RGB_FORMATS[name])
        assert sha(out) == hout, (name, i, w, h, dw, dh, hex(fl))
        n += 1
    assert n == 50


@pytest.mark.parametrize("name", ["bgr24", "rgba", "bgra", "argb", "abgr"])
def test_other_formats_vs_oracle(device, name):
    """Vector and scalar kernels of every writer: aligned and odd strides, 4K same-size (FATE flags and LUT), a rescale."""
    fmt = cl.PACKED_RGB_FORMATS[name]
    for (w, h, dw, dh, fl) in ((640, 360, 640, 360, FATE), (640, 360, 640, 360, cl.SWS_BICUBIC), (346, 286, 346, 286, FATE),
                               (352, 288, 300, 200, cl.SWS_BICUBLIN), (351, 288, 351, 288, FATE),
                               (3840, 2160, 3840, 2160, FATE), (3840, 2160, 3840, 2160, cl.SWS_BICUBIC)):
        for seed, kind, pad, dpad in ((1, "random", 0, 0), (2, "limited", 7, 5)):
            if w >= 3840 and (seed == 2 or name in ("argb", "abgr")):
                continue
            y, u, v = cl.yuv_frame(w, h, seed, kind, pad=pad)
            a = cl.orc_sws(w, h, dw, dh, fl, y, u, v, dst_pad=dpad, fmt=fmt)
            b = gpu_sws(device, w, h, dw, dh, fl, y, u, v, dst_pad=dpad, fmt=fmt)
            assert np.array_equal(a, b), (name, w, h, dw, dh, hex(fl), seed, int((a != b).sum()))


def test_other_formats_slices_and_batch(device):
    """bgra through the slice entry and the batched device entry: same bytes as the whole-frame call."""
    import torch
    from ffmpeg_b200 import swscale as sw
    w, h = 128, 96
    y, u, v = cl.yuv_frame(w, h, 77, "random")
    for fl in (FATE, cl.SWS_BICUBIC):
        ref = cl.orc_sws(w, h, w, h, fl, y, u, v, fmt=cl.PIX_FMT_BGRA)
        ctx = sw.sws_getContext(device, w, h, 0, w, h, sw.AV_PIX_FMT_BGRA, fl)
        out = np.full((h, w * 4), 0xA5, np.uint8)
        total = 0
        for sy in range(0, h, 32):
            total += ctx.scale([y[sy:], u[sy // 2:], v[sy // 2:]], [w, w // 2, w // 2], sy, 32, [out], [w * 4])
        assert total == h and np.array_equal(out, ref), hex(fl)
        with torch.cuda.stream(torch.cuda.ExternalStream(device.stream)):
            dy, du, dv = (torch.from_numpy(np.stack([a, a])).cuda() for a in (y, u, v))
            dout = torch.zeros((2, h, w * 4), dtype=torch.uint8, device="cuda")
            ctx.scale_batch_device([dy, du, dv], [w, w // 2, w // 2], [w * h, w * h // 4, w * h // 4], dout, w * 4, w * h * 4, 2)
            device.sync()
            got = dout.cpu().numpy()
        assert np.array_equal(got[0], ref) and np.array_equal(got[1], ref), hex(fl)
        ctx.free()
    with pytest.raises(Exception):
        sw.sws_getContext(device, 64, 48, 0, 64, 48, 4, FATE)           # AV_PIX_FMT_YUV422P as destination: not this path


def gpu_sws_planar(device, w, h, dw, dh, fl, y, u, v, dst_pad=0, src_fmt=0, ranges=(0, 0), details=None, dst_fmt=0):
    from ffmpeg_b200 import swscale as sw
    ctx = sw.sws_getContext(device, w, h, src_fmt, dw, dh, dst_fmt, fl, src_range=ranges[0], dst_range=ranges[1])
    try:
        if details is not None:
            assert ctx.setColorspaceDetails(cl.COEFFS[details[0]], details[1], cl.COEFFS[details[2]], details[3], *details[4:]) == 0
        return ctx.convert_planar(y, u, v, dst_pad=dst_pad)
    finally:
        ctx.free()


def test_planar_golden_hashes(device):
    """yuv420p -> yuv420p scaling against the reference's outputs (sws_planar_hashes.txt)."""
    n = 0
    for line in open(os.path.join(G, "sws_planar_hashes.txt")):
        i, w, h, dw, dh, fl, kind, hout = line.split()
        i, w, h, dw, dh, fl = map(int, (i, w, h, dw, dh, fl))
        y, u, v = cl.yuv_frame(w, h, 600 + i, kind)
        out = gpu_sws_planar(device, w, h, dw, dh, fl, y, u, v)
        assert sha(np.concatenate([p.ravel() for p in out])) == hout, (i, w, h, dw, dh, hex(fl))
        n += 1
    assert n == 14


@pytest.mark.parametrize("case", [(640, 360, 1280, 720, FATE), (1920, 1080, 1280, 720, cl.SWS_BICUBIC), (3840, 2160, 1920, 1080, FATE),
                                  (351, 287, 351, 287, cl.SWS_BICUBIC), (640, 360, 641, 361, cl.SWS_BILINEAR)])
def test_planar_vs_oracle(device, case):
    w, h, dw, dh, fl = case
    for seed, kind, pad, dpad in ((1, "random", 0, 0), (2, "limited", 5, 3)):
        y, u, v = cl.yuv_frame(w, h, seed, kind, pad=pad)
        a = cl.orc_sws_planar(w, h, dw, dh, fl, y, u, v, dst_pad=dpad)
        b = gpu_sws_planar(device, w, h, dw, dh, fl, y, u, v, dst_pad=dpad)
        for pa, pb, name in zip(a, b, "YUV"):
            assert np.array_equal(pa, pb), (case, seed, name, int((pa != pb).sum()))


def test_planar_batch_device_and_errors(device):
    import torch
    import ffmpeg_b200 as fb
    from ffmpeg_b200 import swscale as sw
    w, h, dw, dh, n = 320, 180, 480, 270, 3
    frames = [cl.yuv_frame(w, h, 90 + i, "random") for i in range(n)]
    Y, U, V = (np.stack([f[k] for f in frames]) for k in range(3))
    ctx = sw.sws_getContext(device, w, h, 0, dw, dh, sw.AV_PIX_FMT_YUV420P, FATE)
    cw, ch, cdw, cdh = w // 2, h // 2, dw // 2, dh // 2
    with torch.cuda.stream(torch.cuda.ExternalStream(device.stream)):
        dY, dU, dV = torch.from_numpy(Y).cuda(), torch.from_numpy(U).cuda(), torch.from_numpy(V).cuda()
        oY = torch.zeros((n, dh, dw), dtype=torch.uint8, device="cuda")
        oU = torch.zeros((n, cdh, cdw), dtype=torch.uint8, device="cuda")
        oV = torch.zeros((n, cdh, cdw), dtype=torch.uint8, device="cuda")
        ctx.scale_batch_device_planar([dY, dU, dV], [w, cw, cw], [w * h, cw * ch, cw * ch], [oY, oU, oV], [dw, cdw, cdw],
                                      [dw * dh, cdw * cdh, cdw * cdh], n)
        with pytest.raises(fb.B200Error):                      # single-plane entry point on a planar context
            ctx.scale_batch_device([dY, dU, dV], [w, cw, cw], [w * h, cw * ch, cw * ch], oY, dw, dw * dh, n)
        device.sync()
        got = [t.cpu().numpy() for t in (oY, oU, oV)]
    for i in range(n):
        ref = cl.orc_sws_planar(w, h, dw, dh, FATE, *frames[i])
        for k in range(3):
            assert np.array_equal(got[k][i], ref[k]), (i, k)
    planes = [np.zeros((dh, dw), np.uint8), np.zeros((cdh, cdw), np.uint8), np.zeros((cdh, cdw), np.uint8)]
    with pytest.raises(fb.B200Error):                          # invalid slice parameters (odd start line), as scale_internal rejects them
        ctx.scale([Y[0][17:], U[0][8:], V[0][8:]], [w, cw, cw], 17, 16, planes, [dw, cdw, cdw])
    ctx.free()
    rgb = sw.sws_getContext(device, w, h, 0, dw, dh, sw.AV_PIX_FMT_RGB24, FATE)
    with pytest.raises(fb.B200Error):
        rgb.scale_batch_device_planar([0, 0, 0], [w, cw, cw], [0, 0, 0], [0, 0, 0], [dw, cdw, cdw], [0, 0, 0], 1)
    rgb.free()


def test_fast_bilinear_golden_hashes_and_oracle(device):
    """SWS_FAST_BILINEAR: reference fixtures for both destinations, plus larger oracle comparisons (4K -> 1080p, upscales)."""
    n = 0
    for line in open(os.path.join(G, "sws_fastbil_hashes.txt")):
        i, w, h, dw, dh, fl, kind, hrgb, hyuv = line.split()
        i, w, h, dw, dh, fl = map(int, (i, w, h, dw, dh, fl))
        y, u, v = cl.yuv_frame(w, h, 800 + i, kind)
        assert sha(gpu_sws(device, w, h, dw, dh, fl, y, u, v)) == hrgb, (i, w, h, dw, dh, hex(fl))
        assert sha(np.concatenate([p.ravel() for p in gpu_sws_planar(device, w, h, dw, dh, fl, y, u, v)])) == hyuv, (i, w, h, dw, dh, hex(fl))
        n += 1
    assert n == 11
    for (w, h, dw, dh, fl) in ((3840, 2160, 1920, 1080, 1), (1280, 720, 1920, 1080, 1 | cl.SWS_ACCURATE_RND), (641, 361, 333, 201, 1)):
        y, u, v = cl.yuv_frame(w, h, 5, "random", pad=3)
        assert np.array_equal(cl.orc_sws(w, h, dw, dh, fl, y, u, v), gpu_sws(device, w, h, dw, dh, fl, y, u, v)), (w, h, dw, dh)
        a, b = cl.orc_sws_planar(w, h, dw, dh, fl, y, u, v), gpu_sws_planar(device, w, h, dw, dh, fl, y, u, v)
        assert all(np.array_equal(p, q) for p, q in zip(a, b)), (w, h, dw, dh)


def test_nv12_nv21_golden_hashes(device):
    n = 0
    for line in open(os.path.join(G, "sws_nv_hashes.txt")):
        name, i, w, h, dw, dh, fl, kind, hrgb, hbgra, hyuv = line.split()
        i, w, h, dw, dh, fl = map(int, (i, w, h, dw, dh, fl))
        sf = cl.PIX_FMT_NV12 if name == "nv12" else cl.PIX_FMT_NV21
        y, u, v = cl.yuv_frame(w, h, 1000 + i, kind)
        uv = cl.nv_interleave(u, v, sf)
        assert sha(gpu_sws(device, w, h, dw, dh, fl, y, uv, uv, src_fmt=sf)) == hrgb, (name, i)
        assert sha(gpu_sws(device, w, h, dw, dh, fl, y, uv, uv, src_fmt=sf, fmt=cl.PIX_FMT_BGRA)) == hbgra, (name, i)
        assert sha(np.concatenate([p.ravel() for p in gpu_sws_planar(device, w, h, dw, dh, fl, y, uv, uv, src_fmt=sf)])) == hyuv, (name, i)
        n += 1
    assert n == 18


def test_nv12_4k_slices_batches_and_bottom_up(device):
    """The decoder-output case: nv12 4K frames to rgb24 (vector kernels after the chroma split), through every entry point:
    whole frame, top-down slices, bottom-up strides, batched device pointers, batched host pointers."""
    import torch
    from ffmpeg_b200 import swscale as sw
    w, h, n = 3840, 2160, 3
    frames = [cl.yuv_frame(w, h, 60 + i, "random") for i in range(n)]
    uvs = [cl.nv_interleave(f[1], f[2], cl.PIX_FMT_NV12, pad=32) for f in frames]
    refs = [cl.orc_sws(w, h, w, h, FATE, frames[i][0], uvs[i], uvs[i], src_fmt=cl.PIX_FMT_NV12) for i in range(n)]
    # the split must give what the planar source gives
    assert np.array_equal(refs[0], cl.orc_sws(w, h, w, h, FATE, *frames[0]))
    ctx = sw.sws_getContext(device, w, h, sw.AV_PIX_FMT_NV12, w, h, sw.AV_PIX_FMT_RGB24, FATE)
    assert np.array_equal(ctx.convert(frames[0][0], uvs[0], uvs[0]), refs[0])
    out = np.zeros((h, w * 3), np.uint8)
    total = 0
    for sy in range(0, h, 540):
        total += ctx.scale([frames[1][0][sy:], uvs[1][sy // 2:]], [w, uvs[1].strides[0]], sy, 540, [out], [w * 3])
    assert total == h and np.array_equal(out, refs[1])
    # bottom-up source (negative strides): the picture is read upside down, like the reference does
    yf, uvf = np.ascontiguousarray(frames[2][0][::-1]), np.ascontiguousarray(uvs[2][::-1])
    ref_flipped = cl.orc_sws(w, h, w, h, FATE, yf, uvf, uvf, src_fmt=cl.PIX_FMT_NV12)
    out = np.zeros((h, w * 3), np.uint8)
    uvst = uvs[2].strides[0]
    assert ctx.scale([frames[2][0].ctypes.data + (h - 1) * w, uvs[2].ctypes.data + (h // 2 - 1) * uvst], [-w, -uvst], 0, h, [out], [w * 3]) == h
    assert np.array_equal(out, ref_flipped)
    Y = np.stack([f[0] for f in frames]); UV = np.stack([np.ascontiguousarray(x[:, :w]) for x in uvs])
    with torch.cuda.stream(torch.cuda.ExternalStream(device.stream)):
        dY, dUV = torch.from_numpy(Y).cuda(), torch.from_numpy(UV).cuda()
        do = torch.zeros((n, h, w * 3), dtype=torch.uint8, device="cuda")
        ctx.scale_batch_device([dY, dUV], [w, w], [w * h, w * h // 2], do, w * 3, w * h * 3, n)
        device.sync()
        got = do.cpu().numpy()
    ho = np.zeros((n, h, w * 3), np.uint8)
    ctx.scale_batch_host([Y.ctypes.data, UV.ctypes.data], [w, w], [w * h, w * h // 2], ho.ctypes.data, w * 3, w * h * 3, n)
    for i in range(n):
        assert np.array_equal(got[i], refs[i]) and np.array_equal(ho[i], refs[i]), i
    ctx.free()
    # nv12 -> yuv420p: scaled, and the same-size de-interleave (nv12ToPlanarWrapper)
    for (dw, dh) in ((1920, 1080), (w, h)):
        a = cl.orc_sws_planar(w, h, dw, dh, FATE, frames[0][0], uvs[0], uvs[0], src_fmt=cl.PIX_FMT_NV12)
        b = gpu_sws_planar(device, w, h, dw, dh, FATE, frames[0][0], uvs[0], uvs[0], src_fmt=cl.PIX_FMT_NV12)
        assert all(np.array_equal(p, q) for p, q in zip(a, b)), (dw, dh)


def test_bottom_up_strides(device):
    """Negative strides (bottom-up pictures) are legal for sws_scale (libswscale/swscale.c:1141-1159)."""
    from ffmpeg_b200 import swscale as sw
    w, h = 64, 48
    y, u, v = cl.yuv_frame(w, h, 9, "random")
    ref_out = cl.orc_sws(w, h, w, h, FATE, y[::-1].copy(), u[::-1].copy(), v[::-1].copy())
    ctx = sw.sws_getContext(device, w, h, 0, w, h, 2, FATE)
    out = np.zeros((h, w * 3), np.uint8)
    n = ctx.scale([y.ctypes.data + (h - 1) * w, u.ctypes.data + (h // 2 - 1) * (w // 2), v.ctypes.data + (h // 2 - 1) * (w // 2)],
                  [-w, -(w // 2), -(w // 2)], 0, h, [out], [w * 3])
    assert n == h and np.array_equal(out, ref_out)
    # and a bottom-up destination
    out2 = np.zeros((h, w * 3), np.uint8)
    ctx.scale([y, u, v], [w, w // 2, w // 2], 0, h, [out2.ctypes.data + (h - 1) * w * 3], [-w * 3])
    assert np.array_equal(out2[::-1], cl.orc_sws(w, h, w, h, FATE, y, u, v))
    ctx.free()


def test_slice_calls_match_reference(device):
    """sws_scale() band by band: per-call return values and the final picture equal the reference's (fixtures)."""
    from ffmpeg_b200 import swscale as sw
    from cases import SWS_SLICE_CASES
    g = np.load(os.path.join(G, "sws_slices.npz"))
    y, u, v = g["y"], g["u"], g["v"]
    for ci, (dw, dh, fl, bands) in enumerate(SWS_SLICE_CASES):
        ctx = sw.sws_getContext(device, 64, 48, 0, dw, dh, 2, fl)
        out = np.full((dh, dw * 3), 0xA5, np.uint8)
        rets = []
        for rep in range(2):                                        # the context is reusable for the next frame
            rets = [ctx.scale([y.ctypes.data + sy * 64, u.ctypes.data + (sy // 2) * 32, v.ctypes.data + (sy // 2) * 32],
                              [64, 32, 32], sy, sh, [out], [dw * 3]) for (sy, sh) in bands]
        assert rets == list(g[f"s{ci}_rets"]), (ci, rets, list(g[f"s{ci}_rets"]))
        assert np.array_equal(out, g[f"s{ci}_rgb"]), ci
        ctx.free()


def test_bad_slices_rejected(device):
    from ffmpeg_b200 import swscale as sw
    import ffmpeg_b200 as fb
    ctx = sw.sws_getContext(device, 64, 48, 0, 64, 48, 2, FATE)
    y, u, v = cl.yuv_frame(64, 48, 1)
    out = np.zeros((48, 192), np.uint8)
    with pytest.raises(fb.B200Error):
        ctx.scale([y, u, v], [64, 32, 32], 1, 16, [out], [192])     # odd start line: "Slice parameters invalid"
    with pytest.raises(fb.B200Error):
        ctx.scale([y, u, v], [64, 32, 32], 16, 16, [out], [192])    # "Slices start in the middle!"
    ctx.free()
    with pytest.raises(fb.B200Error):
        sw.sws_getContext(device, 64, 48, 0, 64, 48, 2, cl.SWS_BICUBIC | cl.SWS_BILINEAR)   # two scalers: EINVAL like the reference


@pytest.mark.parametrize("fl", [FATE, cl.SWS_BICUBIC])
def test_batch_device_and_host_4k(device, fl):
    """Batch entry points at the BASELINE size: every frame of the batch equals the oracle's output for that frame
    (frames are distinct), and device/host entry points agree byte for byte."""
    import torch
    from ffmpeg_b200 import swscale as sw
    w, h, n = 3840, 2160, 6
    frames = [cl.yuv_frame(w, h, 40 + i, "random" if i % 2 else "limited") for i in range(n)]
    Y = np.stack([f[0] for f in frames]); U = np.stack([f[1] for f in frames]); V = np.stack([f[2] for f in frames])
    ctx = sw.sws_getContext(device, w, h, 0, w, h, 2, fl)
    stream = torch.cuda.ExternalStream(device.stream)
    with torch.cuda.stream(stream):
        dY, dU, dV = torch.from_numpy(Y).cuda(), torch.from_numpy(U).cuda(), torch.from_numpy(V).cuda()
        out = torch.empty((n, h, w * 3), dtype=torch.uint8, device="cuda")
        ctx.scale_batch_device([dY, dU, dV], [w, w // 2, w // 2], [w * h, w * h // 4, w * h // 4], out, w * 3, w * h * 3, n)
        device.sync()
        got = out.cpu().numpy()
    host_out = np.zeros((n, h, w * 3), np.uint8)
    ctx.scale_batch_host([Y.ctypes.data, U.ctypes.data, V.ctypes.data], [w, w // 2, w // 2], [w * h, w * h // 4, w * h // 4],
                         host_out.ctypes.data, w * 3, w * h * 3, n)
    assert np.array_equal(got, host_out)
    for i in (0, n - 1):
        ref = cl.orc_sws(w, h, w, h, fl, *frames[i])
        assert np.array_equal(got[i], ref), (i, int((got[i] != ref).sum()))
    # checksum of checksums over the whole batch: each frame differs from its neighbours
    assert len({sha(got[i]) for i in range(n)}) == n
    ctx.free()


def test_full_batch_256_properties(device):
    """BASELINE config 2 at its full size (256 x 4K frames, 9.6 GB per call): the batch holds 8 distinct frames repeated 32
    times, so every output must equal the output of the same frame elsewhere in the batch (frames are independent, no
    state leaks across the batch), and two of the distinct frames are checked against the oracle bit for bit."""
    import torch
    from ffmpeg_b200 import swscale as sw
    w, h, n, nd = 3840, 2160, 256, 8
    frames = [cl.yuv_frame(w, h, 900 + i, "random" if i % 2 else "limited") for i in range(nd)]
    ctx = sw.sws_getContext(device, w, h, 0, w, h, sw.AV_PIX_FMT_RGB24, FATE)
    with torch.cuda.stream(torch.cuda.ExternalStream(device.stream)):
        dY, dU, dV = (torch.from_numpy(np.stack([f[k] for f in frames])).cuda().repeat(n // nd, 1, 1) for k in range(3))
        out = torch.empty((n, h, w * 3), dtype=torch.uint8, device="cuda")
        ctx.scale_batch_device([dY, dU, dV], [w, w // 2, w // 2], [w * h, w * h // 4, w * h // 4], out, w * 3, w * h * 3, n)
        device.sync()
        base = out[:nd]
        for r in range(1, n // nd):
            assert torch.equal(out[r * nd:(r + 1) * nd], base), r
        # checksum of checksums: the 8 distinct frames give 8 distinct pictures
        sums = [int(base[i].to(torch.int64).sum().item()) for i in range(nd)]
        assert len(set(sums)) == nd
        got = {i: base[i].cpu().numpy() for i in (0, nd - 1)}
    for i, g in got.items():
        assert np.array_equal(g, cl.orc_sws(w, h, w, h, FATE, *frames[i])), i
    ctx.free()


def test_vsynth1_frame0(device):
    """FATE's own test picture (vsynth1 frame 0) through the CUDA path, against the reference's output fixtures."""
    g = np.load(os.path.join(G, "vsynth1_f0.npz"))
    y, u, v = g["y"], g["u"], g["v"]
    assert np.array_equal(gpu_sws(device, 352, 288, 352, 288, FATE, y, u, v), g["rgb_same"])
    assert np.array_equal(gpu_sws(device, 352, 288, 200, 100, FATE, y, u, v), g["rgb_200x100"])
    assert np.array_equal(gpu_sws(device, 352, 288, 352, 288, cl.SWS_BICUBIC, y, u, v), g["rgb_lut"])


def test_fate_filter_pixfmts_md5(device):
    """The CUDA path's frames, wrapped by the reference's NUT muxer exactly as FATE does, hash to the md5 sums the
    reference tree commits for filter-pixfmts-{null,copy,vflip,hflip,crop,scale} and filter-pixdesc-* (5 frames) —
    tests/golden/fate_pixfmts.txt cites each line."""
    import functools
    import test_fate_golden as fg
    if not (cl.have_nut() and os.path.exists(cl.VIDEOGEN)):
        pytest.skip("oracle/_ref/libffnut.so / videogen not built")
    try:
        cl.nut_md5(np.zeros(16 * 16 * 3, np.uint8), 16, 16, cl.PIX_FMT_RGB24)
        cl.vsynth1_frames(1)
    except Exception as e:                                                # checker tools, not the product: skip, never fail
        pytest.skip(f"oracle/_ref tools not usable on this box: {e}")
    fg.check_all(functools.partial(gpu_sws, device), functools.partial(gpu_sws_planar, device), rgb_sources=False, nv_dest=False)


@pytest.mark.parametrize("fmt", [cl.PIX_FMT_RGB24, 0])
def test_bottom_up_slice_order(device, fmt):
    """A slice sequence that starts with the band touching the last line runs bottom-up: the reference flips the picture internally
    (swscale.c:1096-1159), i.e. the result is flip(convert(flip(source))); per-call return values add up to the picture height."""
    from ffmpeg_b200 import swscale as sw
    w, h, dw, dh = 128, 96, 96, 64
    y, u, v = cl.yuv_frame(w, h, 321, "random")
    fy, fu, fv = (np.ascontiguousarray(a[::-1]) for a in (y, u, v))
    ctx = sw.sws_getContext(device, w, h, 0, dw, dh, fmt, FATE)
    if fmt == 0:
        exp = [np.ascontiguousarray(a[::-1]) for a in cl.orc_sws_planar(w, h, dw, dh, FATE, fy, fu, fv)]
        out = [np.zeros((dh, dw), np.uint8), np.zeros((dh // 2, dw // 2), np.uint8), np.zeros((dh // 2, dw // 2), np.uint8)]
    else:
        exp = [np.ascontiguousarray(cl.orc_sws(w, h, dw, dh, FATE, fy, fu, fv)[::-1])]
        out = [np.zeros((dh, dw * 3), np.uint8)]
    total = 0
    for sy, sh in ((64, 32), (32, 32), (8, 24), (0, 8)):
        total += ctx.scale([y[sy:], u[sy // 2:], v[sy // 2:]], [w, w // 2, w // 2], sy, sh, out, [o.strides[0] for o in out])
    ctx.free()
    assert total == dh
    for a, b in zip(out, exp):
        assert np.array_equal(a, b), int((a != b).sum())
