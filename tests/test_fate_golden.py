"""CPU tier: the oracle's frames against the md5 sums FATE itself holds for this path (tests/ref/fate/filter-pixfmts-* and
filter-pixdesc-* of the reference tree, listed with file:line in tests/golden/fate_pixfmts.txt).  FATE hashes a NUT stream
of rawvideo frames; oracle/ref/ref_nut.c drives the reference's own NUT muxer to wrap our frames the same way.
Needs oracle/_ref/libffnut.so and oracle/_ref/videogen (built by oracle/ref/Makefile wherever the reference tree is
present; they travel with the repository snapshot)."""
import os

import numpy as np
import pytest

import cpulibs as cl
from cases import FATE

FMT_ID = dict(cl.PACKED_RGB_FORMATS, yuv420p=cl.PIX_FMT_YUV420P, nv12=cl.PIX_FMT_NV12, nv21=cl.PIX_FMT_NV21)
NV = ("nv12", "nv21")
needs_nut = pytest.mark.skipif(not (cl.have_nut() and os.path.exists(cl.VIDEOGEN)), reason="oracle/_ref/libffnut.so / videogen not built")


def filtered(planes, test, fmt):
    """the byte moves of FATE's vflip / hflip / crop=100:100:100:100 filters on a converted picture"""
    bpp = 1 if fmt == "yuv420p" or fmt in NV else cl.fmt_bpp(FMT_ID[fmt])
    out = []
    for k, p in enumerate(planes):
        px = p.reshape(p.shape[0], -1, 2 if (fmt in NV and k) else bpp)          # nv12 / nv21 chroma: one (u, v) pair per sample
        o = 100 >> (k > 0)                                             # chroma planes: half the offset and size
        if test == "vflip":
            px = px[::-1]
        elif test == "hflip":
            px = px[:, ::-1]
        elif test == "crop":
            px = px[o:2 * o, o:2 * o]
        out.append(np.ascontiguousarray(px).ravel())
    return np.concatenate(out)


def fate_stream(run_rgb, run_planar, frames, test, fmt, w, h):
    """`scale,format=fmt,<test filter>` on the given vsynth1 frames with FATE's scaler flags -> list of rawvideo packets"""
    pkts = []
    for y, u, v in frames:
        if fmt == "yuv420p":
            planes = run_planar(352, 288, w, h, FATE, y, u, v) if test == "scale" else (y, u, v)
        elif fmt in NV:
            planes = run_planar(352, 288, 352, 288, FATE, y, u, v, dst_fmt=FMT_ID[fmt])
            if test == "scale":                                     # the second scaler: semi-planar in, the same format out
                planes = run_planar(352, 288, w, h, FATE, planes[0], planes[1], planes[1], src_fmt=FMT_ID[fmt], dst_fmt=FMT_ID[fmt])
        else:
            pic = run_rgb(352, 288, 352, 288, FATE, y, u, v, fmt=FMT_ID[fmt])
            if test == "scale":                                     # the second scaler: packed RGB in, the same format out
                pic = run_rgb(352, 288, w, h, FATE, pic, pic, pic, fmt=FMT_ID[fmt], src_fmt=FMT_ID[fmt])
            planes = (pic,)
        pkts.append(filtered(planes, test, fmt))
    return pkts


def check_all(run_rgb, run_planar, rgb_sources=True, nv_dest=True):
    """rgb_sources=False: skip the rows whose second scaler reads packed RGB; nv_dest=False: skip the nv12 / nv21 rows (paths the
    caller does not have yet)"""
    rows = cl.fate_pixfmts_goldens()
    assert len(rows) == 59
    if not rgb_sources:
        rows = [r for r in rows if not (r[1] == "scale" and r[2] in cl.PACKED_RGB_FORMATS)]
    if not nv_dest:
        rows = [r for r in rows if r[2] not in NV]
    src = cl.vsynth1_frames(5)
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vsynth1_f0.npz"))
    assert all(np.array_equal(a, g[k]) for a, k in zip(src[0], "yuv"))              # the generator gives the committed frame 0
    for where, test, fmt, w, h, nframes, md5 in rows:
        pkts = fate_stream(run_rgb, run_planar, src[:nframes], test, fmt, w, h)
        assert cl.nut_md5(pkts, w, h, FMT_ID[fmt]) == md5, (where, test, fmt)


@needs_nut
def test_oracle_reproduces_fate_md5():
    check_all(cl.orc_sws, cl.orc_sws_planar)


@needs_nut
def test_reference_build_reproduces_fate_md5():
    """the compiled reference itself (oracle/_ref/libffref.so) through the same wrapper: validates the wrapper, not us"""
    if not cl.have_ref():
        pytest.skip("oracle/_ref/libffref.so not built")
    check_all(cl.ref_sws, cl.ref_sws_planar)


@needs_nut
def test_md5_wrapper_is_sensitive():
    """one flipped bit in any frame changes the stream md5 (the wrapper hashes the payload, not just headers)"""
    src = cl.vsynth1_frames(5)
    pkts = fate_stream(cl.orc_sws, cl.orc_sws_planar, src, "pixdesc", "rgb24", 352, 288)
    good = cl.nut_md5(pkts, 352, 288, FMT_ID["rgb24"])
    pkts[3] = pkts[3].copy()
    pkts[3][12345] ^= 1
    assert cl.nut_md5(pkts, 352, 288, FMT_ID["rgb24"]) != good
