"""GPU tier: the drop-in function tables behave like the pure C functions they replace when an un-modified caller uses them the way
frame- / slice-threaded decoders do — many threads at once (every host-pointer entry serialises on the device object that owns the
scratch buffer and the stream) — and with negative line sizes (bottom-up / flipped frames: row r at base + r * linesize)."""
import ctypes as C
import threading

import numpy as np
import pytest

import cpulibs as cl
from cases import idct_blocks

pytestmark = pytest.mark.gpu


def test_idct_table_concurrent_threads(device):
    """16 threads x 40 calls of idct_put / idct_add / me_cmp sad / qpel through the tables at the same time; every result equals the
    oracle's (ctypes releases the GIL, so the calls really overlap)."""
    from ffmpeg_b200 import idctdsp, me_cmp, pel
    from ffmpeg_b200._lib import u8p, i16p
    O = cl.oracle()
    c = idctdsp.ff_idctdsp_init(idctdsp.FF_IDCT_SIMPLE, 8, 0)
    m = me_cmp.ff_me_cmp_init(0)
    q = pel.ff_h264qpel_init(8)
    errs = []
    zero = np.zeros(1, np.int64)

    def worker(t):
        rng = np.random.default_rng(100 + t)
        try:
            for it in range(40):
                blk = idct_blocks("dense", 1, 1000 * t + it)[0]
                dest = rng.integers(0, 256, (8, 16 + 8 * (t % 3)), dtype=np.uint8)
                ls = dest.shape[1]
                b1, b2, d1, d2 = blk.copy(), blk.copy(), dest.copy(), dest.copy()
                op = 1 + (it & 1)
                (c.idct_put if op == 1 else c.idct_add)(d1.ctypes.data_as(u8p), ls, b1.ctypes.data_as(i16p))
                O.orc_idct_batch(op, cl.ptr(b2, cl.i16p), 1, cl.ptr(d2), ls, cl.ptr(zero, cl.i64p))
                if not np.array_equal(d1, d2):
                    errs.append(("idct", t, it))
                a = rng.integers(0, 256, (17, 48), dtype=np.uint8)
                b = rng.integers(0, 256, (17, 48), dtype=np.uint8)
                got = m.sad[0](None, a.ctypes.data_as(u8p), b.ctypes.data_as(u8p), 48, 16)
                if got != int(np.abs(a[:16, :16].astype(np.int32) - b[:16, :16].astype(np.int32)).sum()):
                    errs.append(("sad", t, it))
                src = rng.integers(0, 256, (32, 48), dtype=np.uint8)
                o1 = np.zeros((32, 48), np.uint8); o2 = o1.copy()
                pos = (t + it) % 16
                q.put_h264_qpel_pixels_tab[0][pos](C.cast(o1.ctypes.data + 8 * 48 + 8, u8p), C.cast(src.ctypes.data + 8 * 48 + 8, u8p), 48)
                O.orc_h264qpel(0, 0, pos, C.cast(o2.ctypes.data + 8 * 48 + 8, cl.u8p), C.cast(src.ctypes.data + 8 * 48 + 8, cl.u8p), 48)
                if not np.array_equal(o1, o2):
                    errs.append(("qpel", t, it))
        except Exception as ex:                       # pragma: no cover
            errs.append(("exception", t, repr(ex)))

    ts = [threading.Thread(target=worker, args=(t,)) for t in range(16)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs[:5]


def test_idct_table_negative_linesize(device):
    """idct_put / idct_add / the clamp helpers on a bottom-up picture (dest = last row, line_size < 0), as the C functions allow."""
    from ffmpeg_b200 import idctdsp
    from ffmpeg_b200._lib import u8p, i16p
    O = cl.oracle()
    c = idctdsp.ff_idctdsp_init(idctdsp.FF_IDCT_SIMPLE, 8, 0)
    rng = np.random.default_rng(5)
    zero = np.zeros(1, np.int64)
    for kind in ("dense", "sparse"):
        for op, fn in ((1, c.idct_put), (2, c.idct_add)):
            blk = idct_blocks(kind, 1, 9)[0]
            pic = rng.integers(0, 256, (12, 24), dtype=np.uint8)
            p1, p2 = pic.copy(), pic.copy()
            b1, b2 = blk.copy(), blk.copy()
            # flipped view: row 0 of the block is picture row 9, row 7 is picture row 2
            fn(C.cast(p1.ctypes.data + 9 * 24 + 8, u8p), -24, b1.ctypes.data_as(i16p))
            flip = np.ascontiguousarray(p2[::-1])
            O.orc_idct_batch(op, cl.ptr(b2, cl.i16p), 1, C.cast(flip.ctypes.data + 2 * 24 + 8, cl.u8p), 24, cl.ptr(zero, cl.i64p))
            assert np.array_equal(p1, flip[::-1]), (kind, op)
    for k, fn in enumerate((c.put_pixels_clamped, c.put_signed_pixels_clamped, c.add_pixels_clamped)):
        big = rng.integers(-600, 600, 64).astype(np.int16)
        pic = rng.integers(0, 256, (10, 16), dtype=np.uint8)
        p1 = pic.copy()
        flip = np.ascontiguousarray(pic[::-1])
        fn(big.ctypes.data_as(i16p), C.cast(p1.ctypes.data + 8 * 16 + 4, u8p), -16)
        O.orc_pixels_clamped(k, cl.ptr(big, cl.i16p), C.cast(flip.ctypes.data + 1 * 16 + 4, cl.u8p), 16)
        assert np.array_equal(p1, flip[::-1]), k


def test_pel_and_mecmp_tables_negative_stride(device):
    from ffmpeg_b200 import pel, me_cmp
    from ffmpeg_b200._lib import u8p
    O = cl.oracle()
    q = pel.ff_h264qpel_init(8)
    m = me_cmp.ff_me_cmp_init(0)
    rng = np.random.default_rng(6)
    src = rng.integers(0, 256, (40, 48), dtype=np.uint8)
    fsrc = np.ascontiguousarray(src[::-1])
    for pos in (0, 5, 10, 15):
        o1 = np.zeros((40, 48), np.uint8)
        o2 = np.zeros((40, 48), np.uint8)
        # rows run upwards in memory: block row 0 is array row 27
        q.put_h264_qpel_pixels_tab[0][pos](C.cast(o1.ctypes.data + 27 * 48 + 8, u8p), C.cast(src.ctypes.data + 27 * 48 + 8, u8p), -48)
        O.orc_h264qpel(0, 0, pos, C.cast(o2.ctypes.data + 12 * 48 + 8, cl.u8p), C.cast(fsrc.ctypes.data + 12 * 48 + 8, cl.u8p), 48)
        assert np.array_equal(o1, o2[::-1]), pos
    a = rng.integers(0, 256, (20, 32), dtype=np.uint8)
    b = rng.integers(0, 256, (20, 32), dtype=np.uint8)
    got = m.sad[0](None, C.cast(a.ctypes.data + 17 * 32, u8p), C.cast(b.ctypes.data + 17 * 32, u8p), -32, 16)
    assert got == int(np.abs(a[2:18, :16].astype(np.int32) - b[2:18, :16].astype(np.int32)).sum())
