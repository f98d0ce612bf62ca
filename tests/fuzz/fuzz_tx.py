"""Differential fuzzing, CPU only (test infrastructure): random av_tx types / lengths / scales / strides / padding through b200_tx_init_device + b200_tx_batch_device on the emulated device vs the checker.
Usage: python tests/fuzz/fuzz_tx.py SEED COUNT   — prints every disagreement and a summary line; the deterministic short form of the
swscale loop runs in the CPU tier (tests/test_cuda_emu.py::test_sws_differential_fuzz)."""
import sys, random, ctypes as C
import os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, cpulibs as cl
import test_cuda_emu as te
from test_oracle_more import _tx, _txi, _dct
from ffmpeg_b200._lib import TX_FN
fx = te.emutx; fn = getattr(fx, "__pytest_wrapped__", None); L = (fn.obj if fn else fx.__wrapped__)()
O = cl.oracle()
seed=int(sys.argv[1]); N=int(sys.argv[2]); rnd=random.Random(seed); rng=np.random.default_rng(seed)
bad=0; ran=0
POW2=[2,4,8,16,32,64,128,256]
PFA=[12,20,24,28,36,40,48,56,60,72,80,96,112,120,144,160,240]
for it in range(N):
    typ=rnd.choice([0,1,1,6,9,4,5]); inv=rnd.choice([0,1]); cnt=rnd.choice([1,2,5,9])
    flags=0
    if typ==0: n=rnd.choice(POW2 + [6,10,12,14,18,24,30,36,40,56,60,72,96,120,160,240]);     # power of two or compound N x 2^k
    elif typ==1: n=rnd.choice(POW2[1:]+PFA); flags = 4 if (inv and rnd.random()<0.3) else 0
    elif typ==6: n=rnd.choice(POW2[1:])
    elif typ==9: n=rnd.choice(POW2[1:6])
    elif typ==4: n=rnd.choice(POW2)
    else: n=rnd.choice(POW2[1:])
    sc=rnd.choice([1.0,1.0/n,-1.0,-1.0/32768,0.37])
    pad_in=rnd.choice([0,0,2,6]); pad_out=rnd.choice([0,0,2,10])
    stride_mul=1
    if typ in (1,5) and not flags and rnd.random()<0.3: stride_mul=rnd.choice([2,3])
    # element counts
    if typ in (0,4): ine=oute=2*n
    elif typ in (1,5): ine=(n if inv else 2*n); oute=(2*n if flags else n)
    elif typ==6: ine=(n+2 if inv else n); oute=(n if inv else n+2)
    else: ine=oute=(2*n if inv else n)   # dct: inverse works on 2*len
    isint = typ in (4,5)
    dt=np.int32 if isint else np.float32
    if isint: x=rng.integers(-(1<<20),1<<20,(cnt,ine)).astype(np.int32)
    else: x=(rng.random((cnt,ine),dtype=np.float32)*2-1).astype(np.float32)
    in_mul = stride_mul if (typ in (1,5) and inv) else 1
    out_mul = stride_mul if (typ in (1,5) and not inv) else 1
    xin=np.zeros((cnt,ine*in_mul+pad_in),dt); xin[:,:ine*in_mul:in_mul]=x
    out=np.zeros((cnt,oute*out_mul+pad_out),dt)
    ctx, f, scc = C.c_void_p(), TX_FN(), C.c_float(sc)
    ret=L.b200_tx_init_device(L.dev,C.byref(ctx),C.byref(f),typ,inv,n,C.byref(scc),flags)
    desc=(typ,inv,n,sc,flags,cnt,stride_mul,pad_in,pad_out)
    # oracle
    try:
        if typ==9:
            xp=np.zeros((cnt,ine+2),np.float32); xp[:,:ine]=x
            exp=_dct(O,"orc",inv,n,sc,xp,oute)
        elif isint:
            exp=_txi(O,"orc",typ,inv,n,sc,x,oute)
        else:
            exp=_tx(O,"orc",typ,inv,n,sc,x,oute,flags=flags) if flags else _tx(O,"orc",typ,inv,n,sc,x,oute)
    except AssertionError:
        exp=None
    if ret<0:
        if exp is not None: print("PRODUCT REFUSED", desc, ret); bad+=1
        continue
    if exp is None: print("ORACLE REFUSED", desc); bad+=1; L.b200_tx_uninit(C.byref(ctx)); continue
    st = (8 if typ in (0,4) else 4*stride_mul)
    r=L.b200_tx_batch_device(ctx,out.ctypes.data,xin.ctypes.data,st,cnt,out.strides[0],xin.strides[0])
    L.b200_tx_uninit(C.byref(ctx))
    if r!=0: print("BATCH ERROR",desc,r); bad+=1; continue
    got=out[:,:oute*out_mul:out_mul]
    ran+=1
    if not np.array_equal(got.view(np.uint32),exp.view(np.uint32)): print("MISMATCH",desc); bad+=1
    if pad_out and out[:,oute*out_mul:].any(): print("WROTE INTO PADDING",desc); bad+=1
print("seed",seed,"ran",ran,"bad",bad)
