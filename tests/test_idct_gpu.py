"""GPU tier: the CUDA IDCT path (through the C ABI) against the oracle and the golden fixtures."""
import ctypes as C
import os

import numpy as np
import pytest

import cpulibs as cl
from cases import idct_blocks

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KINDS = ("dense", "wide", "extreme", "sparse", "dc63", "dconly")


def gpu_batch(device, op, blk, dest, off, ls):
    import torch
    from ffmpeg_b200 import idctdsp
    st = torch.cuda.ExternalStream(device.stream)
    with torch.cuda.stream(st):
        db = torch.from_numpy(blk).cuda()
        dd = torch.from_numpy(dest).cuda()
        do = torch.from_numpy(off).cuda()
        idctdsp.idct_batch_device(device, op, db, blk.shape[0], dd, do, None, ls)
        device.sync()
        return db.cpu().numpy(), dd.cpu().numpy()


def test_golden(device):
    g = np.load(os.path.join(G, "idct.npz"))
    off = (np.arange(256) * 8).astype(np.int64)
    for kind in KINDS:
        for op in (0, 1, 2):
            b, d = gpu_batch(device, op, g[f"{kind}_in"].copy(), g[f"{kind}_dest"].copy(), off, 256 * 8)
            ref = g[f"{kind}_op{op}"]
            got = b if op == 0 else d
            assert np.array_equal(got, ref), (kind, op, int((got != ref).sum()))
            if op != 0:
                assert np.array_equal(b, g[f"{kind}_in"])      # coefficients untouched (documented difference)


@pytest.mark.parametrize("n", [1, 3, 4, 31, 33, 1000, 48960])
def test_vs_oracle_ragged_counts(device, n):
    O = cl.oracle()
    rng = np.random.default_rng(n)
    off = (rng.permutation(n) * 8).astype(np.int64)            # scattered, non-monotonic destinations
    for kind in KINDS:
        for op in (0, 1, 2):
            blk = idct_blocks(kind, n, seed=n + op)
            dest = rng.integers(0, 256, (8, n * 8), dtype=np.uint8)
            b2, d2 = blk.copy(), dest.copy()
            O.orc_idct_batch(op, cl.ptr(b2, cl.i16p), n, cl.ptr(d2), n * 8, cl.ptr(off, cl.i64p))
            b1, d1 = gpu_batch(device, op, blk.copy(), dest.copy(), off, n * 8)
            if op == 0:
                assert np.array_equal(b1, b2), (kind, n)
            else:
                assert np.array_equal(d1, d2), (kind, op, n)


def test_empty_batch(device):
    from ffmpeg_b200 import idctdsp
    import torch
    t = torch.zeros(64, dtype=torch.int16, device="cuda")
    assert idctdsp.idct_batch_device(device, 0, t, 0) == 0


def test_unaligned_dest(device):
    """dest rows that are not 8-byte aligned take the byte path."""
    O = cl.oracle()
    n = 64
    rng = np.random.default_rng(3)
    off = (np.arange(n) * 8 + 3).astype(np.int64)
    blk = idct_blocks("dense", n, 5)
    dest = rng.integers(0, 256, (8, n * 8 + 8), dtype=np.uint8)
    for op in (1, 2):
        b2, d2 = blk.copy(), dest.copy()
        O.orc_idct_batch(op, cl.ptr(b2, cl.i16p), n, cl.ptr(d2), n * 8 + 8, cl.ptr(off, cl.i64p))
        _, d1 = gpu_batch(device, op, blk.copy(), dest.copy(), off, n * 8 + 8)
        assert np.array_equal(d1, d2)


def test_pointer_table_dropin(device):
    """IDCTDSPContext filled by ff_idctdsp_init: per-block calls on HOST pointers, like checkasm drives the reference."""
    from ffmpeg_b200 import idctdsp
    from ffmpeg_b200._lib import u8p, i16p
    O = cl.oracle()
    c = idctdsp.ff_idctdsp_init(idctdsp.FF_IDCT_SIMPLE, 8, 0)
    assert c.perm_type == 0 and list(c.idct_permutation) == list(range(64))
    rng = np.random.default_rng(1)
    for kind in ("dense", "sparse", "dconly"):
        blk = idct_blocks(kind, 1, 3)[0]
        dest = rng.integers(0, 256, (8, 24), dtype=np.uint8)
        b1, b2, d1, d2 = blk.copy(), blk.copy(), dest.copy(), dest.copy()
        c.idct_put(d1[:, 8:].ctypes.data_as(u8p), 24, b1.ctypes.data_as(i16p))
        O.orc_idct_batch(1, cl.ptr(b2, cl.i16p), 1, d2[:, 8:].ctypes.data_as(cl.u8p), 24, cl.ptr(np.zeros(1, np.int64), cl.i64p))
        assert np.array_equal(d1, d2)
        b1, b2 = blk.copy(), blk.copy()
        c.idct_add(d1[:, 8:].ctypes.data_as(u8p), 24, b1.ctypes.data_as(i16p))
        O.orc_idct_batch(2, cl.ptr(b2, cl.i16p), 1, d2[:, 8:].ctypes.data_as(cl.u8p), 24, cl.ptr(np.zeros(1, np.int64), cl.i64p))
        assert np.array_equal(d1, d2)
        b1, b2 = blk.copy(), blk.copy()
        c.idct(b1.ctypes.data_as(i16p))
        O.orc_idct_batch(0, cl.ptr(b2, cl.i16p), 1, None, 0, None)
        assert np.array_equal(b1, b2)
        for k, fn in enumerate((c.put_pixels_clamped, c.put_signed_pixels_clamped, c.add_pixels_clamped)):
            big = rng.integers(-600, 600, 64).astype(np.int16)
            p1 = rng.integers(0, 256, (8, 16), dtype=np.uint8); p2 = p1.copy()
            fn(big.ctypes.data_as(i16p), p1.ctypes.data_as(u8p), 16)
            O.orc_pixels_clamped(k, cl.ptr(big, cl.i16p), cl.ptr(p2), 16)
            assert np.array_equal(p1, p2), k
    assert idctdsp.lib().b200_idctdsp_init(C.byref(c), 1, 8, 0) < 0      # FF_IDCT_INT: not implemented -> ENOSYS


@pytest.mark.parametrize("op", [1, 2])
@pytest.mark.parametrize("geom", [(120, 68, 64, 32), (45, 3, 3, 1), (7, 2, 0, 0), (8, 1, 16, 8)])
def test_mb420_geometries(device, op, geom):
    """Partial segments (mb_w % 8 != 0), odd line sizes (falls back to the warp-cooperative kernel) and tiny frames."""
    import torch
    from ffmpeg_b200 import idctdsp
    O = cl.oracle()
    mb_w, mb_h, padl, padc = geom
    nf = 2
    nblk = mb_w * mb_h * 6 * nf
    blk = idct_blocks("dense", nblk, 23)
    W, H = mb_w * 16, mb_h * 16
    ls = [W + padl, W // 2 + padc, W // 2 + padc]
    rng = np.random.default_rng(6)
    planes = [rng.integers(0, 256, (nf, H, ls[0]), dtype=np.uint8), rng.integers(0, 256, (nf, H // 2, ls[1]), dtype=np.uint8),
              rng.integers(0, 256, (nf, H // 2, ls[2]), dtype=np.uint8)]
    ref = [p.copy() for p in planes]
    b = np.arange(nblk)
    f, r = b // (mb_w * mb_h * 6), b % (mb_w * mb_h * 6)
    mb, k = r // 6, r % 6
    mby, mbx = mb // mb_w, mb % mb_w
    for pl in range(3):
        sel = (k < 4) if pl == 0 else (k == 3 + pl)
        if pl == 0:
            off = f * H * ls[0] + (mby * 16 + (k >> 1) * 8) * ls[0] + mbx * 16 + (k & 1) * 8
        else:
            off = f * (H // 2) * ls[pl] + (mby * 8) * ls[pl] + mbx * 8
        bs = np.ascontiguousarray(blk[sel])
        O.orc_idct_batch(op, cl.ptr(bs, cl.i16p), int(sel.sum()), cl.ptr(ref[pl]), ls[pl], cl.ptr(np.ascontiguousarray(off[sel]).astype(np.int64), cl.i64p))
    fs = [H * ls[0], (H // 2) * ls[1], (H // 2) * ls[2]]
    st = torch.cuda.ExternalStream(device.stream)
    with torch.cuda.stream(st):
        db = torch.from_numpy(blk).cuda()
        dp = [torch.from_numpy(p).cuda() for p in planes]
        idctdsp.idct_mb420_device(device, op, db, mb_w, mb_h, nf, dp, ls, fs)
        device.sync()
        for pl in range(3):
            got = dp[pl].cpu().numpy()
            assert np.array_equal(got, ref[pl]), (pl, int((got != ref[pl]).sum()))


@pytest.mark.parametrize("op", [1, 2])
def test_mb420_frames(device, op):
    """Macroblock stream of BASELINE config 3 (1080p = 120x68 MBs x 6 blocks): device and host entry points vs oracle."""
    import torch
    from ffmpeg_b200 import idctdsp
    O = cl.oracle()
    mb_w, mb_h, nf = 120, 68, 3
    nblk = mb_w * mb_h * 6 * nf
    blk = np.concatenate([idct_blocks(k, nblk // 4 + 1, 17) for k in ("dense", "sparse", "dc63", "wide")])[:nblk]
    blk = np.ascontiguousarray(blk[np.random.default_rng(2).permutation(nblk)])
    W, H = mb_w * 16, mb_h * 16
    ls = [W + 64, W // 2 + 32, W // 2 + 32]
    rng = np.random.default_rng(4)
    planes = [rng.integers(0, 256, (nf, H, ls[0]), dtype=np.uint8), rng.integers(0, 256, (nf, H // 2, ls[1]), dtype=np.uint8),
              rng.integers(0, 256, (nf, H // 2, ls[2]), dtype=np.uint8)]
    # oracle: per plane offset lists
    ref = [p.copy() for p in planes]
    b = np.arange(nblk)
    f, r = b // (mb_w * mb_h * 6), b % (mb_w * mb_h * 6)
    mb, k = r // 6, r % 6
    mby, mbx = mb // mb_w, mb % mb_w
    for pl in range(3):
        sel = (k < 4) if pl == 0 else (k == 3 + pl)
        if pl == 0:
            off = f * H * ls[0] + (mby * 16 + (k >> 1) * 8) * ls[0] + mbx * 16 + (k & 1) * 8
        else:
            off = f * (H // 2) * ls[pl] + (mby * 8) * ls[pl] + mbx * 8
        bs = np.ascontiguousarray(blk[sel])
        O.orc_idct_batch(op, cl.ptr(bs, cl.i16p), int(sel.sum()), cl.ptr(ref[pl]), ls[pl], cl.ptr(np.ascontiguousarray(off[sel]).astype(np.int64), cl.i64p))
    fs = [H * ls[0], (H // 2) * ls[1], (H // 2) * ls[2]]
    st = torch.cuda.ExternalStream(device.stream)
    with torch.cuda.stream(st):
        db = torch.from_numpy(blk).cuda()
        dp = [torch.from_numpy(p).cuda() for p in planes]
        idctdsp.idct_mb420_device(device, op, db, mb_w, mb_h, nf, dp, ls, fs)
        device.sync()
        for pl in range(3):
            got = dp[pl].cpu().numpy()
            assert np.array_equal(got, ref[pl]), (pl, int((got != ref[pl]).sum()))
    hp = [p.copy() for p in planes]
    idctdsp.idct_mb420_host(device, op, blk, mb_w, mb_h, nf, hp, ls, fs)
    for pl in range(3):
        # the host entry point only writes the picture area (W x H), padding columns stay untouched
        assert np.array_equal(hp[pl], ref[pl]), pl


@pytest.mark.parametrize("op", [1, 2])
def test_mb420_full_batch_properties(device, op):
    """BASELINE config 3 batch size (256 x 1080p frames = 12.5 M blocks per call): 4 distinct frames of coefficients repeated
    64 times into identical destination pictures must give 64 identical groups of pictures, and the first group equals the
    oracle bit for bit (put: the result does not depend on the old destination; add: it does, so destinations repeat too)."""
    import torch
    from ffmpeg_b200 import idctdsp
    O = cl.oracle()
    mb_w, mb_h, nd, reps = 120, 68, 4, 64
    per = mb_w * mb_h * 6
    nblk = per * nd
    blk = np.concatenate([idct_blocks(k, nblk // 4 + 1, 23) for k in ("dense", "sparse", "wide", "dc63")])[:nblk]
    blk = np.ascontiguousarray(blk[np.random.default_rng(6).permutation(nblk)])
    W, H = mb_w * 16, mb_h * 16
    ls = [W, W // 2, W // 2]
    rng = np.random.default_rng(7)
    planes = [rng.integers(0, 256, (nd, H, ls[0]), dtype=np.uint8), rng.integers(0, 256, (nd, H // 2, ls[1]), dtype=np.uint8),
              rng.integers(0, 256, (nd, H // 2, ls[2]), dtype=np.uint8)]
    ref = [p.copy() for p in planes]
    b = np.arange(nblk)
    f, r = b // per, b % per
    mb, k = r // 6, r % 6
    mby, mbx = mb // mb_w, mb % mb_w
    for pl in range(3):
        sel = (k < 4) if pl == 0 else (k == 3 + pl)
        if pl == 0:
            off = f * H * ls[0] + (mby * 16 + (k >> 1) * 8) * ls[0] + mbx * 16 + (k & 1) * 8
        else:
            off = f * (H // 2) * ls[pl] + (mby * 8) * ls[pl] + mbx * 8
        bs = np.ascontiguousarray(blk[sel])
        O.orc_idct_batch(op, cl.ptr(bs, cl.i16p), int(sel.sum()), cl.ptr(ref[pl]), ls[pl], cl.ptr(np.ascontiguousarray(off[sel]).astype(np.int64), cl.i64p))
    fs = [H * ls[0], (H // 2) * ls[1], (H // 2) * ls[2]]
    with torch.cuda.stream(torch.cuda.ExternalStream(device.stream)):
        db = torch.from_numpy(blk).cuda().repeat(reps, 1)
        dp = [torch.from_numpy(p).cuda().repeat(reps, 1, 1) for p in planes]
        idctdsp.idct_mb420_device(device, op, db, mb_w, mb_h, nd * reps, dp, ls, fs)
        device.sync()
        for pl in range(3):
            base = dp[pl][:nd]
            for rr in range(1, reps):
                assert torch.equal(dp[pl][rr * nd:(rr + 1) * nd], base), (pl, rr)
            assert np.array_equal(base.cpu().numpy(), ref[pl]), pl


# ---------------------------------------------------------------------------------------------- H.264 residual transforms
def test_h264_idct_pointer_table_golden(device):
    from ffmpeg_b200 import idctdsp
    g = np.load(os.path.join(G, "h264idct.npz"))
    t = idctdsp.ff_h264dsp_idct_init(8, 1)
    fns = [t.idct_add, t.idct8_add, t.idct_dc_add, t.idct8_dc_add]
    for kind in range(4):
        blk, out = g[f"k{kind}_in"].copy(), g[f"k{kind}_dst"].copy()
        n = blk.shape[0]
        for i in range(0, n, 3):
            fns[kind](out.ctypes.data + 8 * i, blk.ctypes.data + i * blk.strides[0], n * 8)
            N = 4 if kind in (0, 2) else 8
            assert np.array_equal(out[:, 8 * i:8 * i + N], g[f"k{kind}_out"][:, 8 * i:8 * i + N]), (kind, i)
            assert np.array_equal(blk[i], g[f"k{kind}_blk_after"][i]), (kind, i)
    assert idctdsp.ff_h264dsp_idct_init(10, 1).idct_add                       # the 16-bit functions (test_h264_idct_hbd)


@pytest.mark.parametrize("kind", [0, 1, 2, 3])
def test_h264_idct_batch_vs_oracle(device, kind):
    """A picture worth of residual blocks of one kind, in raster order over a 640x368 plane, checked against the oracle;
    the coefficient buffer must come back cleared exactly like the reference leaves it."""
    import torch
    from ffmpeg_b200 import idctdsp
    O = cl.oracle()
    N, nc = (4, 16) if kind in (0, 2) else (8, 64)
    W, H = 640, 368
    rng = np.random.default_rng(50 + kind)
    nb = (W // N) * (H // N)
    blk = np.zeros((nb, 64), np.int16)                          # blocks spaced 64 coefficients apart (16-byte aligned starts)
    vals = rng.integers(-700, 701, (nb, nc)).astype(np.int16)
    vals[::5] = rng.integers(-32768, 32768, (len(vals[::5]), nc))
    if kind >= 2:
        vals[:, 1:] = 0
    blk[:, :nc] = vals
    dst0 = rng.integers(0, 256, (H, W), dtype=np.uint8)
    by, bx = np.divmod(np.arange(nb), W // N)
    doff = (by * N * W + bx * N).astype(np.int64)
    boff = (np.arange(nb) * 64).astype(np.int64)
    exp, eb = dst0.copy(), blk.copy()
    for i in range(nb):
        O.orc_h264_idct(kind, C.cast(exp.ctypes.data + int(doff[i]), cl.u8p), C.cast(eb.ctypes.data + i * 128, cl.i16p), W)
    with torch.cuda.stream(torch.cuda.ExternalStream(device.stream)):
        db, dd = torch.from_numpy(blk).cuda(), torch.from_numpy(dst0).cuda()
        idctdsp.h264_idct_batch_device(device, kind, nb, db, torch.from_numpy(boff).cuda(), dd, torch.from_numpy(doff).cuda(), W)
        device.sync()
        assert np.array_equal(dd.cpu().numpy(), exp), kind
        assert np.array_equal(db.cpu().numpy(), eb), kind


@pytest.mark.parametrize("variant", range(7))
@pytest.mark.parametrize("op", [1, 2])
def test_unquant_idct_fused_mb420(device, variant, op):
    """put_dct / add_dequant_dct in one kernel (b200_mpv_unquant_idct_mb420_device) = the checker's inverse quantiser followed by the
    checker's IDCT put / add on every block; add leaves blocks with block_last_index < 0 alone (mpegvideo_dec.c:915-922)."""
    import torch
    from ffmpeg_b200 import mpegvideo
    from ffmpeg_b200._lib import MpvUnquant
    O = cl.oracle()
    mb_w, mb_h, nf = 21, 5, 2
    nblk = mb_w * mb_h * 6 * nf
    cfg, blocks, _, q, last = cl.unquant_case(4000 + variant * 3 + op, variant, nblocks=nblk)
    deq = cl.orc_unquant(variant, cfg, blocks, None, q, last)                   # block number inside the macroblock = i % 6
    W, H = mb_w * 16, mb_h * 16
    ls = [W + 32, W // 2 + 16, W // 2 + 16]
    rng = np.random.default_rng(16 + variant)
    planes = [rng.integers(0, 256, (nf, H, ls[0]), dtype=np.uint8), rng.integers(0, 256, (nf, H // 2, ls[1]), dtype=np.uint8),
              rng.integers(0, 256, (nf, H // 2, ls[2]), dtype=np.uint8)]
    ref = [p.copy() for p in planes]
    b = np.arange(nblk)
    f, r = b // (mb_w * mb_h * 6), b % (mb_w * mb_h * 6)
    mb, k = r // 6, r % 6
    mby, mbx = mb // mb_w, mb % mb_w
    live = (last >= 0) if op == 2 else np.ones(nblk, bool)
    for pl in range(3):
        sel = ((k < 4) if pl == 0 else (k == 3 + pl)) & live
        if pl == 0:
            off = f * H * ls[0] + (mby * 16 + (k >> 1) * 8) * ls[0] + mbx * 16 + (k & 1) * 8
        else:
            off = f * (H // 2) * ls[pl] + (mby * 8) * ls[pl] + mbx * 8
        bs = np.ascontiguousarray(deq[sel])
        O.orc_idct_batch(op, cl.ptr(bs, cl.i16p), int(sel.sum()), cl.ptr(ref[pl]), ls[pl], cl.ptr(np.ascontiguousarray(off[sel]).astype(np.int64), cl.i64p))
    fs = [H * ls[0], (H // 2) * ls[1], (H // 2) * ls[2]]
    params = cl.unquant_params(struct=MpvUnquant, **cfg)
    st = torch.cuda.ExternalStream(device.stream)
    with torch.cuda.stream(st):
        db = torch.from_numpy(blocks).cuda()
        dq, dl = torch.from_numpy(q).cuda(), torch.from_numpy(last).cuda()
        dp = [torch.from_numpy(p).cuda() for p in planes]
        mpegvideo.unquant_idct_mb420_device(device, variant, params, op, db, dq, dl, mb_w, mb_h, nf, dp, ls, fs)
        device.sync()
        assert np.array_equal(db.cpu().numpy(), blocks)                          # the coefficient stream is read-only
        for pl in range(3):
            got = dp[pl].cpu().numpy()
            assert np.array_equal(got, ref[pl]), (variant, op, pl, int((got != ref[pl]).sum()))
