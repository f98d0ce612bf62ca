#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "bottom_up or unquant or fused" > gpurun_out/t_new.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t_new.log
python - <<'PY'
import torch, sys, time
sys.path.insert(0, '.')
import ffmpeg_b200 as fb
from ffmpeg_b200 import mpegvideo, idctdsp
stream = torch.cuda.Stream(); dev = fb.Device(0, stream=stream.cuda_stream)
MB_W, MB_H, fr = 120, 68, 64
nbq = MB_W * MB_H * 6 * fr
zz = [0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
      35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63]
perm, rend = mpegvideo.ff_init_scantable(list(range(64)), zz)
prm = mpegvideo.unquant_params([8] + [16 + (i % 23) for i in range(1, 64)], [16] * 64, perm, rend, 8, 8)
with torch.cuda.stream(stream):
    qb = torch.randint(-40, 41, (nbq, 64), dtype=torch.int16, device="cuda")
    qsc = torch.randint(1, 32, (nbq,), dtype=torch.uint8, device="cuda")
    lastq = torch.full((nbq,), 63, dtype=torch.int8, device="cuda")
    pls = [torch.zeros((fr, MB_H * 16, MB_W * 16), dtype=torch.uint8, device="cuda"), torch.zeros((fr, MB_H * 8, MB_W * 8), dtype=torch.uint8, device="cuda"),
           torch.zeros((fr, MB_H * 8, MB_W * 8), dtype=torch.uint8, device="cuda")]
    ls = [MB_W * 16, MB_W * 8, MB_W * 8]; fs = [MB_W * 16 * MB_H * 16, MB_W * 8 * MB_H * 8, MB_W * 8 * MB_H * 8]
    def t(call, name):
        for _ in range(3): call()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(10): call()
        e1.record(stream); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"{name}: {nbq / ms / 1e6:.2f} G blocks/s ({ms:.3f} ms)")
    t(lambda: mpegvideo.unquantize_batch_device(dev, 2, prm, qb, nbq, None, qsc, lastq), "unquant alone (in place)")
    t(lambda: idctdsp.idct_mb420_device(dev, 1, qb, MB_W, MB_H, fr, pls, ls, fs), "idct put alone")
    t(lambda: mpegvideo.unquant_idct_mb420_device(dev, 2, prm, 1, qb, qsc, lastq, MB_W, MB_H, fr, pls, ls, fs), "fused unquant + idct put")
dev.close()
PY
