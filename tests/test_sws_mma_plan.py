"""CPU replay of the tensor-core horizontal pass (csrc/sws_mma.cuh): the library's operand tables (b200_sws_mma_probe) are run through a
lane-by-lane model of mma.sync.m16n8k32 (PTX ISA fragment layouts: A row = lane/4 (+8), k = 4*(lane%4)+b (+16); B k likewise, n = lane/4;
D row = lane/4 (+8), columns 2*(lane%4), +1) with the kernel's addressing, and compared with hScale8To15_c (libswscale/swscale.c:128-142)
computed directly.  Filter banks: the library's own (every scaler / ratio of tests/cases.py) and random ones."""
import ctypes as C
import numpy as np
import pytest
import cpulibs as cl
from ffmpeg_b200._lib import lib


def mma_tables(coef, pos):
    n, size = coef.shape
    L = lib()
    ng = (n + 7) // 8
    ginfo = np.zeros((ng + 1, 2), np.int32)
    pitch = np.zeros(2, np.int32)
    words = L.b200_sws_mma_probe(cl.ptr(coef, cl.i16p), cl.ptr(pos, cl.i32p), n, size, cl.ptr(ginfo, cl.i32p), ginfo.size, None, 0, cl.ptr(pitch, cl.i32p))
    assert words > 0 and words % 128 == 0
    bfrag = np.zeros(words, np.uint32)
    assert L.b200_sws_mma_probe(cl.ptr(coef, cl.i16p), cl.ptr(pos, cl.i32p), n, size, None, 0,
                                bfrag.ctypes.data_as(C.POINTER(C.c_uint32)), words, None) == words
    return ginfo, bfrag.reshape(-1, 32, 4), pitch


def hscale_direct(src, coef, pos):
    n, size = coef.shape
    idx = pos[:, None] + np.arange(size)[None, :]
    acc = (src[:, idx].astype(np.int64) * coef[None].astype(np.int64)).sum(axis=2)
    return np.minimum(acc >> 7, 32767).astype(np.int16)


def replay_tiles(src, coef, pos, tile_groups):
    """The kernel's data flow: per tile the staged window [c0, c0 + 16 * nseg) (zeros beyond srcW), per group the chunked MMAs."""
    rows, srcW = src.shape
    n = coef.shape[0]
    ginfo, bfrag, pitch = mma_tables(coef, pos)
    ng = (n + 7) // 8
    SP = int(pitch[0] if tile_groups == 16 else pitch[1])
    assert SP % 32 == 16                                                      # 16 x odd: conflict-free fragment loads
    out = np.zeros((rows, ng * 8), np.int64)
    rpad = (rows + 15) & ~15
    for G0 in range(0, ng, tile_groups):
        gs = range(G0, min(G0 + tile_groups, ng))
        s = min(int(ginfo[G, 0]) for G in gs)
        e = max(int(ginfo[G, 0]) + 32 * int(ginfo[G + 1, 1] - ginfo[G, 1]) for G in gs)
        c0 = s & ~15
        nseg = (e - c0 + 15) >> 4
        assert nseg * 16 <= SP
        st = np.zeros((rpad, SP), np.uint8)
        w = max(0, min(srcW, c0 + 16 * nseg) - c0)
        st[:rows, :w] = src[:, c0:c0 + w]
        for G in gs:
            kstart, ch0 = int(ginfo[G, 0]), int(ginfo[G, 1])
            nch = int(ginfo[G + 1, 1]) - ch0
            assert kstart % 4 == 0
            krel = kstart - c0
            for rb in range(rpad // 16):
                hi = np.zeros((16, 8), np.int64); lo = np.zeros((16, 8), np.int64)
                for c in range(nch):
                    A = st[rb * 16:rb * 16 + 16, krel + 32 * c:krel + 32 * c + 32].astype(np.int64)      # 16 x 32 as the fragments address it
                    Bh = np.zeros((32, 8), np.int64); Bl = np.zeros((32, 8), np.int64)
                    for lane in range(32):
                        g, t = lane >> 2, lane & 3
                        wv = bfrag[ch0 + c, lane]
                        for half in range(2):
                            for b in range(4):
                                k = 16 * half + 4 * t + b
                                hb = (int(wv[half]) >> (8 * b)) & 0xff
                                Bh[k, g] = hb - 256 if hb >= 128 else hb                                  # .s8
                                Bl[k, g] = (int(wv[2 + half]) >> (8 * b)) & 0xff                          # .u8
                    hi += A @ Bh
                    lo += A @ Bl
                out[rb * 16:rb * 16 + 16, G * 8:G * 8 + 8][:max(0, min(16, rows - rb * 16))] = \
                    np.minimum((hi * 256 + lo) >> 7, 32767)[:max(0, min(16, rows - rb * 16))]
    return out[:, :n].astype(np.int16)


@pytest.mark.parametrize("srcW,dstW,flags", [(3840, 1920, 4 | 0x40000 | 0x80000), (1920, 1280, 4 | 0x40000 | 0x80000), (352, 200, 4 | 0x40000 | 0x80000),
                                             (640, 1280, 2), (1000, 96, 0x200 | 0x40000 | 0x80000), (720, 704, 0x20), (3840, 1280, 0x400 | 0x80000)])
def test_library_banks_replay(srcW, dstW, flags):
    rng = np.random.default_rng(srcW * 7 + dstW)
    L = lib()
    info = np.zeros(16, np.int32)
    assert L.b200_sws_plan_probe(srcW, 64, dstW, 64, flags, 0, None, None, 0, cl.ptr(info, cl.i32p)) >= 0
    for which, sw, n in ((0, srcW, dstW), (1, int(info[4]), int(info[6]))):
        size = int(info[which])
        f = np.zeros(n * size, np.int16); p = np.zeros(n, np.int32)
        assert L.b200_sws_plan_probe(srcW, 64, dstW, 64, flags, which, cl.ptr(f, cl.i16p), cl.ptr(p, cl.i32p), n, None) == n
        coef = f[:n * size].reshape(n, size).copy(); pos = p[:n].copy()
        assert int(pos.min()) >= 0 and int((pos + size).max()) <= sw
        src = rng.integers(0, 256, (37, sw), dtype=np.uint8)
        ref = hscale_direct(src, coef, pos)
        for tg in (16, 8):
            assert np.array_equal(replay_tiles(src, coef, pos, tg), ref), (which, tg)


def test_random_banks_replay():
    rng = np.random.default_rng(5)
    for _ in range(12):
        n = int(rng.integers(9, 300)); size = int(rng.integers(1, 40))
        step = float(rng.uniform(0.3, 6.0))
        pos = np.floor(np.arange(n) * step + rng.integers(0, 3, n)).astype(np.int32)
        srcW = int(pos.max()) + size + int(rng.integers(0, 20))
        coef = rng.integers(-16384, 16384, (n, size)).astype(np.int16)
        src = rng.integers(0, 256, (21, srcW), dtype=np.uint8)
        ref = hscale_direct(src, coef, pos)
        for tg in (16, 8):
            assert np.array_equal(replay_tiles(src, coef, pos, tg), ref)
