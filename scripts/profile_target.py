#!/usr/bin/env python
"""Small driver for ncu: launches each hot kernel a few times on device-resident synthetic data.
   python scripts/profile_target.py [sws|lut|idct|all] [nframes]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ffmpeg_b200 as fb
from ffmpeg_b200 import swscale as sw, idctdsp

what = sys.argv[1] if len(sys.argv) > 1 else "all"
nf = int(sys.argv[2]) if len(sys.argv) > 2 else 64
W, H = 3840, 2160
stream = torch.cuda.Stream()
dev = fb.Device(0, stream=stream.cuda_stream)
with torch.cuda.stream(stream):
    Y = torch.randint(0, 256, (nf, H, W), dtype=torch.uint8, device="cuda")
    U = torch.randint(0, 256, (nf, H // 2, W // 2), dtype=torch.uint8, device="cuda")
    V = torch.randint(0, 256, (nf, H // 2, W // 2), dtype=torch.uint8, device="cuda")
    O = torch.empty((nf, H, W * 3), dtype=torch.uint8, device="cuda")
    stream.synchronize()
    if what in ("sws", "all"):
        c = sw.sws_getContext(dev, W, H, 0, W, H, 2, 4 | 0x40000 | 0x80000)
        for _ in range(4):
            c.scale_batch_device([Y, U, V], [W, W // 2, W // 2], [W * H, W * H // 4, W * H // 4], O, W * 3, W * H * 3, nf)
        dev.sync(); c.free()
    if what in ("lut", "all"):
        c = sw.sws_getContext(dev, W, H, 0, W, H, 2, 4)
        for _ in range(4):
            c.scale_batch_device([Y, U, V], [W, W // 2, W // 2], [W * H, W * H // 4, W * H // 4], O, W * 3, W * H * 3, nf)
        dev.sync(); c.free()
    if what in ("scale", "all"):
        c = sw.sws_getContext(dev, W, H, 0, 1920, 1080, 2, 4 | 0x40000 | 0x80000)
        for _ in range(2):
            c.scale_batch_device([Y, U, V], [W, W // 2, W // 2], [W * H, W * H // 4, W * H // 4], O, 1920 * 3, 1920 * 1080 * 3, min(nf, 8))
        dev.sync(); c.free()
    if what in ("scalep", "all"):
        c = sw.sws_getContext(dev, W, H, 0, 1920, 1080, 0, 4 | 0x40000 | 0x80000)
        nb = min(nf, 8)
        oy = torch.empty((nb, 1080, 1920), dtype=torch.uint8, device="cuda")
        ou = torch.empty((nb, 540, 960), dtype=torch.uint8, device="cuda"); ov = torch.empty_like(ou)
        for _ in range(2):
            c.scale_batch_device_planar([Y, U, V], [W, W // 2, W // 2], [W * H, W * H // 4, W * H // 4], [oy, ou, ov], [1920, 960, 960],
                                        [1920 * 1080, 960 * 540, 960 * 540], nb)
        dev.sync(); c.free()
    if what in ("idct", "all"):
        mbw, mbh, fr = 120, 68, nf
        n = mbw * mbh * 6 * fr
        blk = torch.randint(-256, 257, (n, 64), dtype=torch.int16, device="cuda")
        pl = [torch.zeros((fr, mbh * 16, mbw * 16), dtype=torch.uint8, device="cuda"),
              torch.zeros((fr, mbh * 8, mbw * 8), dtype=torch.uint8, device="cuda"),
              torch.zeros((fr, mbh * 8, mbw * 8), dtype=torch.uint8, device="cuda")]
        for kind in (1, 2):
            for _ in range(4):
                idctdsp.idct_mb420_device(dev, kind, blk, mbw, mbh, fr, pl, [mbw * 16, mbw * 8, mbw * 8],
                                          [mbw * 16 * mbh * 16, mbw * 8 * mbh * 8, mbw * 8 * mbh * 8])
            dev.sync()
    if what in ("h264", "all"):
        Hh, Wh, hn = 1088, 1920, 16
        planes_h = torch.randint(0, 256, (hn, Hh, Wh), dtype=torch.uint8, device="cuda")
        for kind, N in ((0, 4), (1, 8)):
            nb = hn * (Hh // N) * (Wh // N)
            coef = torch.randint(-600, 601, (nb, N * N), dtype=torch.int16, device="cuda")
            bi = torch.arange(nb, device="cuda", dtype=torch.int64)
            per = (Hh // N) * (Wh // N)
            fr, r = bi // per, bi % per
            hdoff = (fr * (Hh * Wh) + (r // (Wh // N)) * (N * Wh) + (r % (Wh // N)) * N).contiguous()
            hboff = (bi * (N * N)).contiguous()
            for _ in range(2):
                idctdsp.h264_idct_batch_device(dev, kind, nb, coef, hboff, planes_h, hdoff, Wh)
            dev.sync()
    if what in ("tx", "all"):
        from ffmpeg_b200 import tx
        n, cnt = 1024, 1 << 16
        x = torch.rand((cnt, 2 * n), device="cuda"); y = torch.empty_like(x)
        c = tx.av_tx_init(0, 0, n, device=dev)
        for _ in range(3):
            c.batch_device(y, x, 8, cnt, 8 * n, 8 * n)
        dev.sync(); c.uninit()
        c = tx.av_tx_init(1, 1, n, scale=1.0 / n, device=dev)
        for _ in range(3):
            c.batch_device(y, x, 4, cnt, 4 * n, 4 * n)
        dev.sync(); c.uninit()
    if what in ("tx2048",):
        from ffmpeg_b200 import tx
        n, cnt = 2048, 1 << 15
        x = torch.rand((cnt, 2 * n), device="cuda"); y = torch.empty_like(x)
        c = tx.av_tx_init(0, 0, n, device=dev)
        for _ in range(3):
            c.batch_device(y, x, 8, cnt, 8 * n, 8 * n)
        dev.sync(); c.uninit()
        c = tx.av_tx_init(1, 1, n, scale=1.0 / n, device=dev)
        for _ in range(3):
            c.batch_device(y, x, 4, cnt, 4 * n, 4 * n)
        dev.sync(); c.uninit()
    if what in ("qpel", "all"):
        from ffmpeg_b200 import pel
        nfr, Wd, Hd, ap = max(16, nf), 1920, 1088, 32
        PW, PH = Wd + 2 * ap, Hd + 2 * ap
        refp = torch.randint(0, 256, (nfr, PH, PW), dtype=torch.uint8, device="cuda")
        dstp = torch.randint(0, 256, (nfr, PH, PW), dtype=torch.uint8, device="cuda")
        fi = torch.arange(nfr, device="cuda").view(-1, 1, 1)
        by = torch.arange(Hd // 16, device="cuda").view(1, -1, 1); bx = torch.arange(Wd // 16, device="cuda").view(1, 1, -1)
        base = fi * (PH * PW) + (by * 16 + ap) * PW + bx * 16 + ap
        dx = torch.randint(-16, 17, base.shape, device="cuda"); dy = torch.randint(-16, 17, base.shape, device="cuda")
        doff = base.reshape(-1).to(torch.int64).contiguous(); soff = (base + dy * PW + dx).reshape(-1).to(torch.int64).contiguous()
        ops = (torch.randint(0, 2, (doff.numel(),), device="cuda") | (torch.randint(0, 16, (doff.numel(),), device="cuda") << 3)).to(torch.uint8)
        for _ in range(3):
            pel.h264qpel_batch_device(dev, doff.numel(), ops, dstp, doff, refp, soff, PW)
        dev.sync()
    if what in ("chroma", "all"):
        from ffmpeg_b200 import pel
        nfr, Wd, Hd, ap = 32, 960, 544, 16
        PW, PH = Wd + 2 * ap, Hd + 2 * ap
        refp = torch.randint(0, 256, (nfr, PH, PW), dtype=torch.uint8, device="cuda")
        dstp = torch.randint(0, 256, (nfr, PH, PW), dtype=torch.uint8, device="cuda")
        fi = torch.arange(nfr, device="cuda").view(-1, 1, 1)
        by = torch.arange(Hd // 8, device="cuda").view(1, -1, 1); bx = torch.arange(Wd // 8, device="cuda").view(1, 1, -1)
        base = fi * (PH * PW) + (by * 8 + ap) * PW + bx * 8 + ap
        dx = torch.randint(-8, 9, base.shape, device="cuda"); dy = torch.randint(-8, 9, base.shape, device="cuda")
        doff = base.reshape(-1).to(torch.int64).contiguous(); soff = (base + dy * PW + dx).reshape(-1).to(torch.int64).contiguous()
        n = doff.numel()
        ops = torch.randint(0, 2, (n,), device="cuda").to(torch.uint8)
        hs = torch.full((n,), 8, dtype=torch.uint8, device="cuda")
        xys = torch.randint(0, 64, (n,), device="cuda").to(torch.uint8)
        for _ in range(3):
            pel.h264chroma_batch_device(dev, n, ops, hs, xys, dstp, doff, refp, soff, PW)
        dev.sync()
    if what in ("esa", "all"):
        from ffmpeg_b200 import me_cmp
        cur = torch.randint(0, 256, (1, H, W), dtype=torch.uint8, device="cuda")
        ref = torch.roll(cur, shifts=(7, -13), dims=(1, 2)).contiguous()
        nmb = (W // 16) * (H // 16)
        mv = torch.zeros((1, nmb, 2), dtype=torch.int32, device="cuda"); cost = torch.zeros((1, nmb), dtype=torch.int64, device="cuda")
        for _ in range(2):
            me_cmp.me_esa_device(dev, cur, ref, W, W, H, W * H, 1, 16, 32, mv, cost)
        dev.sync()
dev.close()
print("done", fb.launch_count())
