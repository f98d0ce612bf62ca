"""Differential fuzzing, CPU only (test infrastructure): sws_scale() band by band with random band partitions on the emulated device vs the compiled reference (per-call return values and the final picture).
Usage: python tests/fuzz/fuzz_slices.py SEED COUNT   — prints every disagreement and a summary line; the deterministic short form of the
swscale loop runs in the CPU tier (tests/test_cuda_emu.py::test_sws_differential_fuzz)."""
import sys, functools, random, ctypes as C
import os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, cpulibs as cl
import test_cuda_emu as te
fx = te.emusws; fn = getattr(fx, "__pytest_wrapped__", None); L = (fn.obj if fn else fx.__wrapped__)()
R = cl.ref()
seed=int(sys.argv[1]); N=int(sys.argv[2]); rnd=random.Random(seed)
FLAGS=[cl.SWS_BICUBIC, cl.SWS_BILINEAR, te.FATE, cl.SWS_BICUBIC|0x40000, 1, cl.SWS_BICUBIC|0x2000, 0x10, 0x200]
bad=0
for it in range(N):
    w,h = rnd.choice([16,34,64,100]), rnd.choice([8,16,34,48,66])
    dw,dh = (w,h) if rnd.random()<0.3 else (rnd.choice([8,17,32,64,100,200]), rnd.choice([4,8,17,32,64,100]))
    fl=rnd.choice(FLAGS); df=rnd.choice([2,3,cl.PACKED_RGB_FORMATS["bgra"]])
    y,u,v=cl.yuv_frame(w,h,seed*100+it,"random")
    # bands: even boundaries
    cuts=sorted(set([0,h]+[2*rnd.randrange(1,h//2) for _ in range(rnd.randrange(0,4))]))
    bands=[(a,b-a) for a,b in zip(cuts[:-1],cuts[1:])]
    bpp=cl.fmt_bpp(df)
    # reference
    rc=R.ffref_sws_open_io(0,w,h,df,dw,dh,fl,1)
    if not rc: continue
    rout=np.full((dh,dw*bpp),0xA5,np.uint8); rrets=[]
    for (sy,sh) in bands:
        rrets.append(R.ffref_sws_scale(rc, cl.ptr(y[sy:]), y.strides[0], cl.ptr(u[sy//2:]), u.strides[0], cl.ptr(v[sy//2:]), v.strides[0], sy, sh, cl.ptr(rout), rout.strides[0]))
    R.ffref_sws_close(rc)
    ctx=te._emu_ctx(L,w,h,0,dw,dh,df,fl)
    if not ctx: print("product refused",(w,h,dw,dh,hex(fl),df)); bad+=1; continue
    out=np.full((dh,dw*bpp),0xA5,np.uint8); rets=[]
    for (sy,sh) in bands:
        sp=(C.c_void_p*4)(y[sy:].ctypes.data,u[sy//2:].ctypes.data,v[sy//2:].ctypes.data,None); ss=(C.c_int32*4)(y.strides[0],u.strides[0],v.strides[0],0)
        dp=(C.c_void_p*4)(out.ctypes.data,None,None,None); ds=(C.c_int32*4)(out.strides[0],0,0,0)
        rets.append(L.b200_sws_scale(ctx,sp,ss,sy,sh,dp,ds))
    L.b200_sws_freeContext(ctx)
    if rets!=rrets or not np.array_equal(out,rout):
        print("MISMATCH",(w,h,dw,dh,hex(fl),df),bands,rets,rrets,np.array_equal(out,rout)); bad+=1
print("seed",seed,"bad",bad)
