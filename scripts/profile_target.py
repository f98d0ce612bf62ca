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
dev.close()
print("done", fb.launch_count())
