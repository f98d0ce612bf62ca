"""smoke() checks for the paths added after the first slice: ESA motion search, H.264 qpel MC, float FFT / iMDCT —
one small call each, bit-compared with the oracle.  Lives under tests/ because it loads the checker; only
__graft_entry__.smoke() imports it."""
import ctypes as C
import numpy as np


def run(dev):
    import torch
    import cpulibs as cl
    from ffmpeg_b200 import me_cmp, pel, tx
    O = cl.oracle()
    rng = np.random.default_rng(3)
    st = torch.cuda.ExternalStream(dev.stream)
    with torch.cuda.stream(st):
        # exhaustive search on one 160x96 pair
        W, H, mb, sp = 160, 96, 16, 16
        cur = rng.integers(0, 256, (H, W), dtype=np.uint8)
        ref = np.roll(cur, (3, -4), (0, 1)).copy()
        bw, bh = W // mb, H // mb
        emv, ec = np.zeros((bh * bw, 2), np.int32), np.zeros(bh * bw, np.uint64)
        O.orc_esa_frame(cl.ptr(cur), cl.ptr(ref), W, W, H, mb, sp, 0, bh, cl.ptr(emv, cl.i32p), cl.ptr(ec, cl.u64p))
        mv = torch.zeros((bh * bw, 2), dtype=torch.int32, device="cuda")
        cost = torch.zeros(bh * bw, dtype=torch.int64, device="cuda")
        me_cmp.me_esa_device(dev, torch.from_numpy(cur).cuda(), torch.from_numpy(ref).cuda(), W, W, H, W * H, 1, mb, sp, mv, cost)
        dev.sync()
        assert np.array_equal(mv.cpu().numpy(), emv) and np.array_equal(cost.cpu().numpy().astype(np.uint64), ec), "esa mismatch"
        # qpel: all 16 positions, put and avg, on one 16x16 block each
        src = rng.integers(0, 256, (64, 64 * 32), dtype=np.uint8)
        dst = rng.integers(0, 256, (64, 64 * 32), dtype=np.uint8)
        ops = np.array([pel.qpel_op(a, 0, p) for a in (0, 1) for p in range(16)], np.uint8)
        offs = np.array([16 * src.shape[1] + 24 + 64 * i for i in range(32)], np.int64)
        exp = dst.copy()
        for o, off in zip(ops, offs):
            O.orc_h264qpel(int(o) & 1, 0, int(o) >> 3, C.cast(exp.ctypes.data + int(off), cl.u8p), C.cast(src.ctypes.data + int(off), cl.u8p), src.shape[1])
        d_dst = torch.from_numpy(dst).cuda()
        d_off = torch.from_numpy(offs).cuda()
        pel.h264qpel_batch_device(dev, 32, torch.from_numpy(ops).cuda(), d_dst, d_off, torch.from_numpy(src).cuda(), d_off, src.shape[1])
        dev.sync()
        assert np.array_equal(d_dst.cpu().numpy(), exp), "qpel mismatch"
        # h264chroma: all 64 phases, put and avg, 8x8
        ops = np.array([pel.chroma_op(a, 0) for a in (0, 1) for p in range(64)], np.uint8)
        xys = np.array([p for a in (0, 1) for p in range(64)], np.uint8)
        src, dst = src[:, :64 * 128 // 4], dst[:, :64 * 128 // 4]
        src, dst = np.ascontiguousarray(np.tile(src, (1, 4))), np.ascontiguousarray(np.tile(dst, (1, 4)))
        offs = np.array([16 * src.shape[1] + 24 + 64 * i for i in range(128)], np.int64)
        exp = dst.copy()
        for o, xy, off in zip(ops, xys, offs):
            O.orc_h264chroma(int(o) & 1, 0, C.cast(exp.ctypes.data + int(off), cl.u8p), C.cast(src.ctypes.data + int(off), cl.u8p), src.shape[1], 8, int(xy) & 7, int(xy) >> 3)
        d_dst = torch.from_numpy(dst).cuda()
        d_off = torch.from_numpy(offs).cuda()
        pel.h264chroma_batch_device(dev, 128, torch.from_numpy(ops).cuda(), torch.full((128,), 8, dtype=torch.uint8, device="cuda"),
                                    torch.from_numpy(xys).cuda(), d_dst, d_off, torch.from_numpy(src).cuda(), d_off, src.shape[1])
        dev.sync()
        assert np.array_equal(d_dst.cpu().numpy(), exp), "h264chroma mismatch"
        # H.264 residual add: 4x4 and 8x8 transforms, 64 blocks each
        from ffmpeg_b200 import idctdsp
        for kind, N in ((0, 4), (1, 8)):
            nb = 64
            coef = rng.integers(-600, 601, (nb, N * N)).astype(np.int16)
            pix = rng.integers(0, 256, (N, nb * 8), dtype=np.uint8)
            e_pix, e_coef = pix.copy(), coef.copy()
            for i in range(nb):
                O.orc_h264_idct(kind, C.cast(e_pix.ctypes.data + 8 * i, cl.u8p), C.cast(e_coef.ctypes.data + i * e_coef.strides[0], cl.i16p), nb * 8)
            d_coef, d_pix = torch.from_numpy(coef).cuda(), torch.from_numpy(pix).cuda()
            idctdsp.h264_idct_batch_device(dev, kind, nb, d_coef, torch.arange(nb, dtype=torch.int64, device="cuda") * (N * N), d_pix,
                                           torch.arange(nb, dtype=torch.int64, device="cuda") * 8, nb * 8)
            dev.sync()
            assert np.array_equal(d_pix.cpu().numpy(), e_pix) and np.array_equal(d_coef.cpu().numpy(), e_coef), "h264 idct mismatch"
        # tx: FFT-1024 and iMDCT-1024, 4 transforms each
        n = 1024
        x = rng.random((4, 2 * n), dtype=np.float32)
        for typ, inv, scale, inw, outw, stride in ((0, 0, 1.0, 2 * n, 2 * n, 8), (1, 1, 1.0 / n, n, n, 4), (6, 0, 1.0, n, n + 2, 4)):
            h = O.orc_tx_open(typ, inv, n, scale, 0)
            xi = np.ascontiguousarray(x[:, :inw])
            e = np.zeros((4, outw), np.float32)
            O.orc_tx_run(h, e.ctypes.data, xi.ctypes.data, stride, 4, e.strides[0], xi.strides[0])
            O.orc_tx_close(h)
            c = tx.av_tx_init(typ, inv, n, scale=scale if typ else None, device=dev)
            do = torch.zeros((4, outw), dtype=torch.float32, device="cuda")
            c.batch_device(do, torch.from_numpy(xi).cuda(), stride, 4, 4 * outw, 4 * inw)
            dev.sync()
            assert np.array_equal(do.cpu().numpy(), e), "tx mismatch"
            c.uninit()
