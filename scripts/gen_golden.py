#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libffref.so, built by oracle/ref/Makefile
from /root/reference).  Run in the build container only; the fixtures travel to the GPU box.

    python scripts/gen_golden.py

Every fixture stores the inputs (or the seed + a sha256 of the inputs, for the larger cases) and the reference's
outputs, so the tests need neither /root/reference nor oracle/_ref at run time.
"""
import hashlib
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cpulibs as cl  # noqa: E402
from cases import SWS_SMALL_CASES, SWS_HASH_CASES, idct_blocks  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def gen_sws():
    d = {}
    for i, (w, h, dw, dh, fl, kind) in enumerate(SWS_SMALL_CASES):
        y, u, v = cl.yuv_frame(w, h, 100 + i, kind)
        out = cl.ref_sws(w, h, dw, dh, fl, y, u, v)
        assert out is not None, (w, h, dw, dh, fl)
        d[f"c{i}_y"], d[f"c{i}_u"], d[f"c{i}_v"], d[f"c{i}_rgb"] = y, u, v, out
    np.savez_compressed(os.path.join(OUT, "sws_small.npz"), **d)
    lines = []
    for i, (w, h, dw, dh, fl, kind) in enumerate(SWS_HASH_CASES):
        y, u, v = cl.yuv_frame(w, h, 200 + i, kind)
        out = cl.ref_sws(w, h, dw, dh, fl, y, u, v)
        lines.append(f"{i} {w} {h} {dw} {dh} {fl} {kind} {sha(np.concatenate([y.ravel(), u.ravel(), v.ravel()]))} {sha(out)}")
    open(os.path.join(OUT, "sws_hashes.txt"), "w").write("\n".join(lines) + "\n")
    # colourspace variants on one small frame
    d = {}
    y, u, v = cl.yuv_frame(64, 48, 300, "random")
    d["y"], d["u"], d["v"] = y, u, v
    FATE = cl.SWS_BICUBIC | cl.SWS_ACCURATE_RND | cl.SWS_BITEXACT
    for j, cs in enumerate([(1, 0, 1, 0, 0, 1 << 16, 1 << 16), (5, 1, 5, 1, 0, 1 << 16, 1 << 16),
                            (9, 0, 9, 0, 3000, 70000, 80000), (7, 1, 7, 0, -2000, 60000, 50000)]):
        for k, fl in enumerate([FATE, cl.SWS_BICUBIC]):
            d[f"cs{j}_{k}"] = cl.ref_sws(64, 48, 64, 48, fl, y, u, v, colorspace=cs)
            d[f"cs{j}_{k}_s"] = cl.ref_sws(64, 48, 96, 80, fl, y, u, v, colorspace=cs)
    np.savez_compressed(os.path.join(OUT, "sws_colorspace.npz"), **d)


def gen_idct():
    R = cl.ref()
    d = {}
    for kind in ("dense", "wide", "extreme", "sparse", "dc63", "dconly"):
        blk = idct_blocks(kind, 256, seed=7)
        rng = np.random.default_rng(11)
        dest0 = rng.integers(0, 256, (8, 256 * 8), dtype=np.uint8)
        off = (np.arange(256) * 8).astype(np.int64)
        d[f"{kind}_in"] = blk
        d[f"{kind}_dest"] = dest0
        for op in (0, 1, 2):
            b, de = blk.copy(), dest0.copy()
            R.ffref_idct_batch(op, cl.ptr(b, cl.i16p), 256, cl.ptr(de), 256 * 8, cl.ptr(off, cl.i64p))
            d[f"{kind}_op{op}"] = b if op == 0 else de
    np.savez_compressed(os.path.join(OUT, "idct.npz"), **d)


if __name__ == "__main__":
    assert cl.have_ref(), "build oracle/_ref first: make -C oracle/ref"
    os.makedirs(OUT, exist_ok=True)
    gen_sws()
    gen_idct()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
