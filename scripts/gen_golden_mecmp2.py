#!/usr/bin/env python
"""Fixture for the MECmpContext members added in round 2 (vsad / vsse (+ intra), nsse, median_sad, hadamard8_intra, sum_abs_dctelem):
values of the COMPILED REFERENCE (oracle/_ref/libffref.so, ff_me_cmp_init entries called with a NULL context like checkasm/motion.c).
Run in the build container: python scripts/gen_golden_mecmp2.py -> tests/golden/mecmp2.npz"""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cpulibs as cl

R = cl.ref()
R.ffref_sum_abs_dctelem.argtypes = [cl.i16p]
rng = np.random.default_rng(77)
img1 = rng.integers(0, 256, (64, 64), dtype=np.uint8)
img2 = (img1.astype(int) + rng.integers(-20, 21, img1.shape)).clip(0, 255).astype(np.uint8)
cases = []
for fn, idxs in ((3, (4, 5)), (4, (0, 1, 4, 5)), (5, (0, 1, 4, 5)), (6, (0, 1)), (7, (0, 1))):
    for idx in idxs:
        for h in (8, 16):
            for _ in range(6):
                x1, y1, x2, y2 = (int(v) for v in rng.integers(1, 40, 4))
                v = R.ffref_me_cmp(fn, idx, C.cast(img1.ctypes.data + y1 * 64 + x1, cl.u8p), C.cast(img2.ctypes.data + y2 * 64 + x2, cl.u8p), 64, h)
                cases.append((fn, idx, x1, y1, x2, y2, h, v))
blocks = rng.integers(-3000, 3001, (40, 64)).astype(np.int16)
sums = np.array([R.ffref_sum_abs_dctelem(cl.ptr(b, cl.i16p)) for b in blocks], np.int32)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "mecmp2.npz"), img1=img1, img2=img2, cases=np.array(cases, np.int64), blocks=blocks, sums=sums)
print(len(cases), "cases")
