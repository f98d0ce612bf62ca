"""Differential fuzzing, CPU only (test infrastructure): random swscale contexts (formats, sizes, scalers, ranges, colorspace details, padding) through the library on the emulated device vs the checker and the compiled reference.
Usage: python tests/fuzz/fuzz_sws.py SEED COUNT   — prints every disagreement and a summary line; the deterministic short form of the
swscale loop runs in the CPU tier (tests/test_cuda_emu.py::test_sws_differential_fuzz)."""
import sys, os, functools, random
import os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, cpulibs as cl
import test_cuda_emu as te
fx = te.emusws
fn = getattr(fx, "__pytest_wrapped__", None)
L = (fn.obj if fn else fx.__wrapped__)()
er, ep = functools.partial(te.emu_sws, L), functools.partial(te.emu_sws_planar, L)
seed = int(sys.argv[1]); N = int(sys.argv[2])
rnd = random.Random(seed)
FLAGS = [cl.SWS_BICUBIC, cl.SWS_BILINEAR, te.FATE, cl.SWS_BICUBIC | 0x40000, cl.SWS_BILINEAR | 0x80000, 1, cl.SWS_BICUBIC | 0x2000, te.FATE | 0x2000,
         cl.SWS_BICUBIC | 0x4000, 0x10, 0x200, 0x400]
srcs = [0, cl.PIX_FMT_NV12, cl.PIX_FMT_NV21] + list(cl.PACKED_RGB_FORMATS.values())
dsts = [0, cl.PIX_FMT_NV12, cl.PIX_FMT_NV21] + list(cl.PACKED_RGB_FORMATS.values())
bad = 0; ran = 0; refused = 0
for it in range(N):
    w, h = rnd.choice([2,4,6,8,10,16,18,34,66,100,130,258]), rnd.choice([2,4,6,8,10,16,18,34,50,98])
    if rnd.random() < 0.3: w += 1
    if rnd.random() < 0.3: h += 1
    same = rnd.random() < 0.3
    dw, dh = (w, h) if same else (rnd.choice([2,3,8,17,32,64,100,200,301]), rnd.choice([2,3,8,17,32,64,100,151]))
    fl = rnd.choice(FLAGS); sf = rnd.choice(srcs); df = rnd.choice(dsts)
    ranges = rnd.choice([(0,0),(0,0),(0,1),(1,0),(1,1)])
    kind = rnd.choice(["random","limited","smooth"]) if hasattr(cl,'yuv_frame') else "random"
    rgbsrc = sf in cl.PACKED_RGB_FORMATS.values(); rgbdst = df in cl.PACKED_RGB_FORMATS.values()
    cw, ch = (w+1)//2, (h+1)//2
    try:
        if rgbsrc:
            bpp = cl.fmt_bpp(sf); s = cl.rgb_frame(w, h, seed*1000+it, bpp); y=u=v=s
        else:
            y,u,v = cl.yuv_frame(w, h, seed*1000+it, kind)
            if sf: uv = cl.nv_interleave(u, v, sf); u = v = uv
    except Exception as e:
        continue
    dpad = rnd.choice([0,0,1,5,13])
    cs = None
    if rnd.random() < 0.4:
        cs = (rnd.choice([1,2,4,5,6,7,9]), rnd.choice([0,1]), rnd.choice([1,5,6,7,9]), rnd.choice([0,1]), rnd.choice([0, 1<<12, -(1<<13)]), rnd.choice([1<<16, 70000, 50000]), rnd.choice([1<<16, 80000, 40000]))
    filt = None
    if rnd.random() < 0.25:                                     # srcFilter / dstFilter of sws_getContext
        def vec():
            r = rnd.random()
            if r < 0.4: return None
            n = rnd.choice([1, 2, 3, 5, 7])
            c = [rnd.uniform(-0.3, 1.0) for _ in range(n)]
            t = sum(c)
            if abs(t) < 0.25: t = 1.0                         # a near-zero sum would blow the normalised taps up to +-50: the reference itself (and the
                                                              # checker, which restates it) crashes on such vectors (seed 777004, case 354) - out of domain
            return [x / t for x in c] if rnd.random() < 0.7 else c
        filt = (tuple(vec() for _ in range(4)), tuple(rnd.choice([0, 0, 1, 3, 5]) for _ in range(4)))
    fk = {} if filt is None else {"filters": filt}
    desc = (w,h,dw,dh,hex(fl),sf,df,ranges,dpad,cs,filt)
    if os.environ.get("FUZZ_VERBOSE"): print("case", it, desc, flush=True)
    try:
        if rgbdst:
            if ranges != (0,0): ranges=(0,0)
            exp = cl.orc_sws(w,h,dw,dh,fl,y,u,v,fmt=df,src_fmt=sf,dst_pad=dpad,colorspace=cs,**fk)
        else:
            exp = cl.orc_sws_planar(w,h,dw,dh,fl,y,u,v,src_fmt=sf,dst_fmt=df,ranges=ranges,dst_pad=dpad,details=cs,**fk)
    except Exception as e:
        exp = None
    try:
        got = er(w,h,dw,dh,fl,y,u,v,fmt=df,src_fmt=sf,dst_pad=dpad,colorspace=cs,**fk) if rgbdst else ep(w,h,dw,dh,fl,y,u,v,src_fmt=sf,dst_fmt=df,ranges=ranges,dst_pad=dpad,details=cs,**fk)
    except AssertionError as e:
        got = "ERR:"+str(e)[:80]
        import traceback
        tb = traceback.format_exc()
        if "_emu_ctx" in tb: got = None
    if got is None:
        refused += 1
        if exp is not None: print("REFUSED-BY-PRODUCT", desc)
        continue
    if exp is None:
        print("ORACLE REFUSED but product ran", desc); bad += 1; continue
    ran += 1
    if isinstance(got,str): print("PRODUCT ERROR", desc, got); bad += 1; continue
    try:
        rexp = cl.ref_sws(w,h,dw,dh,fl,y,u,v,fmt=df,src_fmt=sf,dst_pad=dpad,colorspace=cs,**fk) if rgbdst else cl.ref_sws_planar(w,h,dw,dh,fl,y,u,v,src_fmt=sf,dst_fmt=df,ranges=ranges,dst_pad=dpad,details=cs,**fk)
        if rgbdst:                                              # picture area only: rgbToRgbWrapper 24-bit -> argb / abgr writes one byte past each row (documented quirk)
            pw = dw * cl.fmt_bpp(df)
            rok = np.array_equal(np.asarray(rexp).reshape(dh, -1)[:, :pw], np.asarray(exp).reshape(dh, -1)[:, :pw])
        else:
            rok = all(np.array_equal(a,b) for a,b in zip(rexp,exp))
        if not rok: print("ORACLE != REFERENCE", desc); bad += 1
    except Exception as e:
        print("ref failed", desc, str(e)[:60])
    ok = np.array_equal(got, exp) if rgbdst else all(np.array_equal(a,b) for a,b in zip(got,exp))
    if not ok: print("MISMATCH", desc); bad += 1
print("seed",seed,"ran",ran,"refused",refused,"bad",bad)
