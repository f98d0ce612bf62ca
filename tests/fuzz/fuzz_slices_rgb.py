"""Differential fuzzing, CPU only (test infrastructure): sws_scale() band by band from packed RGB sources into packed RGB destinations (same-size byte
shuffles, alpha through the scaler, the scaler proper) with random band partitions on the emulated device vs the compiled reference (per-call
return values and the final picture).  Usage: python tests/fuzz/fuzz_slices_rgb.py SEED COUNT"""
import sys, random, ctypes as C
import os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, cpulibs as cl
import test_cuda_emu as te
fx = te.emusws; fn = getattr(fx, "__pytest_wrapped__", None); L = (fn.obj if fn else fx.__wrapped__)()
R = cl.ref()
seed = int(sys.argv[1]); N = int(sys.argv[2]); rnd = random.Random(seed)
FLAGS = [cl.SWS_BICUBIC, cl.SWS_BILINEAR, te.FATE, cl.SWS_BICUBIC | 0x40000, 1, 0x10, cl.SWS_BICUBIC | 0x80000]
names = list(cl.PACKED_RGB_FORMATS)
bad = ran = 0
for it in range(N):
    w, h = rnd.choice([16, 34, 64, 100]), rnd.choice([8, 16, 34, 48, 66])
    dw, dh = (w, h) if rnd.random() < 0.5 else (rnd.choice([8, 17, 32, 64, 100]), rnd.choice([4, 8, 17, 32, 64]))
    fl = rnd.choice(FLAGS); sn, dn = rnd.choice(names), rnd.choice(names)
    sf, df = cl.PACKED_RGB_FORMATS[sn], cl.PACKED_RGB_FORMATS[dn]
    sb, db = cl.fmt_bpp(sf), cl.fmt_bpp(df)
    src = cl.rgb_frame(w, h, seed * 100 + it, sb, "random", pad=rnd.choice([0, 0, 3]))
    cuts = sorted(set([0, h] + [rnd.randrange(1, h) for _ in range(rnd.randrange(0, 4))]))
    bands = [(a, b - a) for a, b in zip(cuts[:-1], cuts[1:])]
    rc = R.ffref_sws_open_io(sf, w, h, df, dw, dh, fl, 1)
    if not rc:
        continue
    rout = np.full((dh + 1, dw * db + 4), 0xA5, np.uint8); rrets = []
    for (sy, sh) in bands:
        rrets.append(R.ffref_sws_scale(rc, cl.ptr(src[sy:]), src.strides[0], cl.ptr(src[sy:]), src.strides[0], cl.ptr(src[sy:]), src.strides[0], sy, sh, cl.ptr(rout), rout.strides[0]))
    R.ffref_sws_close(rc)
    ctx = te._emu_ctx(L, w, h, sf, dw, dh, df, fl)
    if not ctx:
        print("product refused", (w, h, dw, dh, hex(fl), sn, dn)); bad += 1; continue
    out = np.full((dh + 1, dw * db + 4), 0xA5, np.uint8); rets = []
    for (sy, sh) in bands:
        sp = (C.c_void_p * 4)(src[sy:].ctypes.data, None, None, None); ss = (C.c_int32 * 4)(src.strides[0], 0, 0, 0)
        dp = (C.c_void_p * 4)(out.ctypes.data, None, None, None); ds = (C.c_int32 * 4)(out.strides[0], 0, 0, 0)
        rets.append(L.b200_sws_scale(ctx, sp, ss, sy, sh, dp, ds))
    L.b200_sws_freeContext(ctx)
    ran += 1
    # picture area only: the reference's 24 -> argb / abgr shuffle writes one byte past each line
    if rets != rrets or not np.array_equal(out[:dh, :dw * db], rout[:dh, :dw * db]):
        print("MISMATCH", (w, h, dw, dh, hex(fl), sn, dn), bands, rets, rrets, np.array_equal(out[:dh, :dw * db], rout[:dh, :dw * db])); bad += 1
print("seed", seed, "ran", ran, "bad", bad)
