"""Differential fuzzing, CPU only (test infrastructure): sws_scale() band by band into a yuv420p / nv12 destination on the emulated
device vs the compiled reference (per-call return values and the final planes).
Usage: python tests/fuzz/fuzz_slices_planar.py SEED COUNT
The reference itself asserts (swscale.c:474) on some partitions with fast_bilinear and a steep vertical reduction: skipped."""
import sys, random, ctypes as C
import os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, cpulibs as cl
import test_cuda_emu as te
fx = te.emusws; fn = getattr(fx, "__pytest_wrapped__", None); L = (fn.obj if fn else fx.__wrapped__)()
R = cl.ref()
R.ffref_sws_scale_planar.argtypes = [C.c_void_p] + [C.c_void_p, C.c_int] * 3 + [C.c_int, C.c_int] + [C.c_void_p, C.c_int] * 3
seed = int(sys.argv[1]); N = int(sys.argv[2]); rnd = random.Random(seed)
FLAGS = [cl.SWS_BICUBIC, cl.SWS_BILINEAR, te.FATE, 1, 0x10, 0x20, 0x40]     # the wide float kernels trip an assert in the reference (swscale.c:474) on some partitions
bad = 0
for it in range(N):
    w, h = rnd.choice([16, 34, 64, 100]), rnd.choice([8, 16, 34, 48, 66])
    dw, dh = (w, h) if rnd.random() < 0.3 else (rnd.choice([8, 18, 32, 64, 100, 200]), rnd.choice([4, 8, 18, 32, 64, 100]))
    fl = rnd.choice(FLAGS); df = rnd.choice([0, cl.PIX_FMT_NV12]); ranges = rnd.choice([(0, 0), (0, 1), (1, 0)])
    if fl == 1 and h >= 8 * dh:
        continue                                # fast_bilinear with a steep vertical reduction trips the reference's own assert (swscale.c:474) when sliced
    y, u, v = cl.yuv_frame(w, h, seed * 100 + it, "random")
    cuts = sorted(set([0, h] + [2 * rnd.randrange(1, h // 2) for _ in range(rnd.randrange(0, 4))]))
    bands = [(a, b - a) for a, b in zip(cuts[:-1], cuts[1:])]
    cw, ch = (dw + 1) // 2, (dh + 1) // 2
    def planes():
        p = [np.full((dh, dw), 0xA5, np.uint8), np.full((ch, 2 * cw if df else cw), 0xA5, np.uint8)]
        p.append(np.full((ch, cw), 0xA5, np.uint8))
        return p
    if os.environ.get("FUZZ_VERBOSE"): print("case", (w, h, dw, dh, hex(fl), df, ranges), bands, flush=True)
    rc = R.ffref_sws_open_range(0, w, h, ranges[0], df, dw, dh, ranges[1], fl, 1)
    if not rc:
        continue
    rp, rr = planes(), []
    for (sy, sh) in bands:
        rr.append(R.ffref_sws_scale_planar(rc, y[sy:].ctypes.data, y.strides[0], u[sy // 2:].ctypes.data, u.strides[0], v[sy // 2:].ctypes.data, v.strides[0],
                                           sy, sh, rp[0].ctypes.data, rp[0].strides[0], rp[1].ctypes.data, rp[1].strides[0], rp[2].ctypes.data, rp[2].strides[0]))
    R.ffref_sws_close(rc)
    ctx = te._emu_ctx(L, w, h, 0, dw, dh, df, fl, ranges)
    if not ctx:
        print("product refused", (w, h, dw, dh, hex(fl), df, ranges)); bad += 1; continue
    gp, gr = planes(), []
    for (sy, sh) in bands:
        sp = (C.c_void_p * 4)(y[sy:].ctypes.data, u[sy // 2:].ctypes.data, v[sy // 2:].ctypes.data, None); ss = (C.c_int32 * 4)(y.strides[0], u.strides[0], v.strides[0], 0)
        dp = (C.c_void_p * 4)(gp[0].ctypes.data, gp[1].ctypes.data, gp[2].ctypes.data, None); ds = (C.c_int32 * 4)(gp[0].strides[0], gp[1].strides[0], gp[2].strides[0], 0)
        gr.append(L.b200_sws_scale(ctx, sp, ss, sy, sh, dp, ds))
    L.b200_sws_freeContext(ctx)
    npl = 2 if df else 3
    same = all(np.array_equal(a, b) for a, b in zip(gp[:npl], rp[:npl]))
    if gr != rr or not same:
        print("MISMATCH", (w, h, dw, dh, hex(fl), df, ranges), bands, gr, rr, same); bad += 1
print("seed", seed, "bad", bad)
