"""CPU tier: the plan of the register-resident FFT / inverse-MDCT kernel (ffmpeg_b200/csrc/tx_r16.cu) replayed in numpy.

The kernel itself needs a GPU (bulk async copies, mbarriers); what can go wrong in it without one is the schedule: which samples a
thread owns in every pass, the shared-memory indices, which chunk is one block and which is two, the butterfly factors, and the lane
pairing of the MDCT post-rotation.  b200_tx_r16_plan() returns the very words the kernel loads, and this file executes them the way
the kernel does — one numpy lane per CUDA thread, float32 operations rounded one by one — and compares with the checker bit for bit."""
import ctypes as C

import numpy as np
import pytest

import cpulibs as cl

F = np.float32


def plan_words(n, inv):
    import ffmpeg_b200 as fb
    L = fb.lib()
    cnt = L.b200_tx_r16_plan(n, inv, None, 0)
    assert cnt > 0, cnt
    w = np.zeros(cnt, np.uint32)
    assert L.b200_tx_r16_plan(n, inv, w.ctypes.data_as(C.c_void_p), cnt) == cnt
    return w.reshape(-1, n // 16)                                  # [word][thread]


def butterflies(a, i0, i1, i2, i3, t1, t2, t5, t6):
    r0, im0, r1, im1 = a[i0][0], a[i0][1], a[i1][0], a[i1][1]
    t3 = t5 - t1; t5 = t5 + t1
    a2x = r0 - t5; a0x = r0 + t5
    a3y = im1 - t3; a1y = im1 + t3
    t4 = t2 - t6; t6 = t2 + t6
    a3x = r1 - t4; a1x = r1 + t4
    a2y = im0 - t6; a0y = im0 + t6
    a[i0], a[i1], a[i2], a[i3] = (a0x, a0y), (a1x, a1y), (a2x, a2y), (a3x, a3y)


def nomul(a, i0, i1, i2, i3):
    butterflies(a, i0, i1, i2, i3, a[i2][0], a[i2][1], a[i3][0], a[i3][1])


def mul(a, i0, i1, i2, i3, wre, wim):
    t1 = a[i2][0] * wre - a[i2][1] * (-wim)
    t2 = a[i2][0] * (-wim) + a[i2][1] * wre
    t5 = a[i3][0] * wre - a[i3][1] * wim
    t6 = a[i3][0] * wim + a[i3][1] * wre
    butterflies(a, i0, i1, i2, i3, t1, t2, t5, t6)


def fft2r(a, i, j):
    s0, s1 = a[i], a[j]
    a[i] = (s0[0] + s1[0], s0[1] + s1[1])
    a[j] = (s0[0] - s1[0], s0[1] - s1[1])


def sel(mask, x, y):
    return (np.where(mask, x[0], y[0]), np.where(mask, x[1], y[1]))


def leaf16(v, full, c8, c1, c2, c3):
    fft2r(v, 0, 1); fft2r(v, 4, 5); fft2r(v, 6, 7)
    nomul(v, 0, 1, 2, 3); nomul(v, 0, 2, 4, 6); mul(v, 1, 3, 5, 7, c8, c8)
    fft2r(v, 8, 9); nomul(v, 8, 9, 10, 11); fft2r(v, 12, 13)
    a, b = list(v), list(v)
    nomul(a, 12, 13, 14, 15); nomul(a, 0, 4, 8, 12); mul(a, 2, 6, 10, 14, c2, c2); mul(a, 1, 5, 9, 13, c1, c3); mul(a, 3, 7, 11, 15, c3, c1)
    fft2r(b, 14, 15); nomul(b, 8, 10, 12, 14); mul(b, 9, 11, 13, 15, c8, c8)
    for i in range(16):
        v[i] = sel(full, a[i], b[i])


def combine(x, tw, full, NL):
    W = lambda k: (tw[k][0], tw[k][1])
    mul(x, 0, 1, 2, 3, *W(0))
    if NL == 3:
        mul(x, 0, 2, 4, 6, *W(1)); mul(x, 1, 3, 5, 7, *W(2)); mul(x, 8, 9, 10, 11, *W(0))
        a, b = list(x), list(x)
        mul(a, 12, 13, 14, 15, *W(0))
        for u in range(4):
            mul(a, u, u + 4, u + 8, u + 12, *W(3 + u))
        mul(b, 8, 10, 12, 14, *W(1)); mul(b, 9, 11, 13, 15, *W(2))
    else:
        a, b = list(x), list(x)
        mul(a, 0, 2, 4, 6, *W(1)); mul(a, 1, 3, 5, 7, *W(2))
        mul(b, 4, 5, 6, 7, *W(0))
    for i in range(len(x)):
        x[i] = sel(full, a[i], b[i])


def zidx(base, t, Q):
    return base + t * (Q + Q // 16) if Q >= 16 else base + 8 * t + (t >> 1)


NLS = {9: (3, 2), 10: (3, 3), 11: (3, 2, 2), 12: (3, 3, 2)}


def replay(n, inv, mode, src, exp_nat=None):
    """src: FFT: complex64[n]; inverse MDCT: float32[2n].  Returns complex64[n] (the MDCT output viewed as n pairs)."""
    P = plan_words(n, inv)
    G = n // 16
    logn = n.bit_length() - 1
    f32 = lambda w: P[w].view(np.float32)
    zs = n + n // 16
    buf = np.zeros((zs, 2), np.float32)
    v = []
    raw = src.view(np.float32)
    for i in range(16):
        o = (P[i >> 1] >> 16) if (i & 1) else (P[i >> 1] & 0xffff)
        o = o.astype(np.int64)
        if mode == 0:
            v.append((raw[o // 4], raw[o // 4 + 1]))
        else:
            aim = raw[o // 4]
            are = raw[(8 * n - 4 - o) // 4]
            wx, wy = exp_nat[o // 8, 0], exp_nat[o // 8, 1]
            v.append((are * wx - aim * wy, are * wy + aim * wx))
    leaf16(v, P[8] != 0, f32(9), f32(10), f32(11), f32(12))
    tg = np.arange(G)
    for i in range(16):
        buf[17 * tg + i, 0], buf[17 * tg + i, 1] = v[i]
    out = np.zeros((n, 2), np.float32)
    a, w0 = 4, 13
    nls = NLS[logn]
    for k, NL in enumerate(nls):
        last = k == len(nls) - 1
        Q, E = 1 << (a - 1), 2 << NL
        nit, ntw = (1, 7) if NL == 3 else (2, 3)
        for s in range(nit):
            wb = w0 + s * (3 + 2 * ntw)
            base, kind, gidx = P[wb].astype(np.int64), P[wb + 1] != 0, P[wb + 2].astype(np.int64)
            tw = [(f32(wb + 3 + 2 * j), f32(wb + 4 + 2 * j)) for j in range(ntw)]
            x = [(buf[zidx(base, t, Q), 0].copy(), buf[zidx(base, t, Q), 1].copy()) for t in range(E)]
            combine(x, tw, kind, NL)
            if not last:
                for t in range(E):
                    buf[zidx(base, t, Q), 0], buf[zidx(base, t, Q), 1] = x[t]
                continue
            if mode == 1:
                bs = []
                for t in range(E):
                    e = exp_nat[gidx + t * Q]
                    ax = x[t][1] * e[:, 1] - x[t][0] * e[:, 0]
                    bs.append(x[t][1] * e[:, 0] + x[t][0] * e[:, 1])
                    x[t] = (ax, x[t][1])
                lane = tg % 32
                partner = tg - lane + (31 - lane)                  # __shfl_sync(.., 31 - lane) inside the warp
                for t in range(E):
                    x[t] = (x[t][0], bs[E - 1 - t][partner])
            for t in range(E):
                out[gidx + t * Q, 0], out[gidx + t * Q, 1] = x[t]
        a += NL
        w0 += nit * (3 + 2 * ntw)
    assert a == logn
    return out


@pytest.mark.parametrize("n", [512, 1024, 2048, 4096])
@pytest.mark.parametrize("inv", [0, 1])
def test_fft_schedule_is_bit_identical_to_the_checker(n, inv):
    O = cl.oracle()
    rng = np.random.default_rng(n + inv)
    x = (rng.random((2, 2 * n), dtype=np.float32) - F(0.5)) * F(4)
    h = O.orc_tx_open(0, inv, n, 1.0, 0)
    exp = np.zeros_like(x)
    O.orc_tx_run(h, exp.ctypes.data, x.ctypes.data, 8, 2, 8 * n, 8 * n)
    O.orc_tx_close(h)
    for r in range(2):
        got = replay(n, inv, 0, x[r])
        assert np.array_equal(got.reshape(-1).view(np.uint32), exp[r].view(np.uint32)), (n, inv, r)


@pytest.mark.parametrize("length", [1024, 2048, 4096, 8192])
def test_imdct_schedule_is_bit_identical_to_the_checker(length):
    import math
    O = cl.oracle()
    n = length // 2
    scale = 1.0 / length
    rng = np.random.default_rng(length)
    x = (rng.random((2, length), dtype=np.float32) - F(0.5)) * F(4)
    h = O.orc_tx_open(1, 1, length, scale, 0)
    exp = np.zeros((2, length), np.float32)
    O.orc_tx_run(h, exp.ctypes.data, x.ctypes.data, 4, 2, 4 * length, 4 * length)
    O.orc_tx_close(h)
    # ff_tx_mdct_gen_exp (tx_template.c:2107-2134), natural order: what tx.cu hands the kernel
    i = np.arange(n, dtype=np.float64)
    theta = (n if scale < 0 else 0) + 1.0 / 8.0
    alpha = (math.pi / 2) * (i + theta) / n
    amp = math.sqrt(abs(np.float64(np.float32(scale))))
    en = np.stack([(np.cos(alpha) * amp).astype(np.float32), (np.sin(alpha) * amp).astype(np.float32)], axis=1)
    for r in range(2):
        got = replay(n, 1, 1, x[r], en)
        assert np.array_equal(got.reshape(-1).view(np.uint32), exp[r].view(np.uint32)), (length, r)
