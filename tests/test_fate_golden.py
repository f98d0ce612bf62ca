"""CPU tier: the oracle's frames against the md5 sums FATE itself holds for this path (tests/ref/fate/filter-pixfmts-null,
-copy and -scale of the reference tree, listed with file:line in tests/golden/fate_pixfmts.txt).  FATE hashes a NUT stream
with one rawvideo frame; oracle/ref/ref_nut.c drives the reference's own NUT muxer to wrap our frame the same way.
Needs oracle/_ref/libffnut.so (built by oracle/ref/Makefile wherever the reference tree is present; it travels with the
repository snapshot)."""
import os

import numpy as np
import pytest

import cpulibs as cl
from cases import FATE

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FMT_ID = dict(cl.PACKED_RGB_FORMATS, yuv420p=cl.PIX_FMT_YUV420P)
needs_nut = pytest.mark.skipif(not cl.have_nut(), reason="oracle/_ref/libffnut.so not built")


def fate_frame(run_rgb, run_planar, fmt, w, h):
    """vsynth1 frame 0 through `scale,format=fmt[,scale=WxH]` with FATE's scaler flags"""
    g = np.load(os.path.join(G, "vsynth1_f0.npz"))
    y, u, v = g["y"], g["u"], g["v"]
    if fmt == "yuv420p" and (w, h) == (352, 288):
        return np.concatenate([y.ravel(), u.ravel(), v.ravel()]), (y, u, v)          # the null filter on the source format
    if fmt == "yuv420p":
        return np.concatenate([p.ravel() for p in run_planar(352, 288, w, h, FATE, y, u, v)]), (y, u, v)
    return run_rgb(352, 288, w, h, FATE, y, u, v, fmt=FMT_ID[fmt]), (y, u, v)


@needs_nut
def test_oracle_reproduces_fate_filter_pixfmts_md5():
    rows = cl.fate_pixfmts_goldens()
    assert len(rows) == 8
    for where, test, fmt, w, h, md5 in rows:
        frame, _ = fate_frame(cl.orc_sws, cl.orc_sws_planar, fmt, w, h)
        assert cl.nut_md5(frame, w, h, FMT_ID[fmt]) == md5, (where, test, fmt)


@needs_nut
def test_reference_build_reproduces_fate_md5():
    """the compiled reference itself (oracle/_ref/libffref.so) through the same wrapper: validates the wrapper, not us"""
    if not cl.have_ref():
        pytest.skip("oracle/_ref/libffref.so not built")
    for where, test, fmt, w, h, md5 in cl.fate_pixfmts_goldens():
        frame, _ = fate_frame(cl.ref_sws, cl.ref_sws_planar, fmt, w, h)
        assert cl.nut_md5(frame, w, h, FMT_ID[fmt]) == md5, (where, test, fmt)


@needs_nut
def test_md5_wrapper_is_sensitive():
    """one flipped bit in the frame changes the stream md5 (the wrapper hashes the payload, not just headers)"""
    frame, _ = fate_frame(cl.orc_sws, cl.orc_sws_planar, "rgb24", 352, 288)
    bad = frame.copy()
    bad[100, 100] ^= 1
    assert cl.nut_md5(bad, 352, 288, FMT_ID["rgb24"]) != cl.nut_md5(frame, 352, 288, FMT_ID["rgb24"])
