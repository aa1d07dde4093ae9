#!/usr/bin/env python3
"""tools/fuzz_host_decoders.py -- memory-safety fuzzing of the single-warp CUDA decoders through their host emulation.

    g++ -O1 -g -fsanitize=address,undefined -fno-sanitize-recover=undefined -Wno-unknown-pragmas -shared -fPIC \\
        -o /tmp/asan/libqdec3_host.so tools/qdec3_host.cpp
    LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 python tools/fuzz_host_decoders.py

Feeds random garbage, bit-flipped, truncated and garbage-tailed streams to every decoder variant (q_decode3<0..2>, q_decode6 and q_decode8 with
both layouts, the fast and the adaptive decoder).  A decoder may return anything, but it must not touch memory outside its
shared-memory image, its cold-counter arrays and its output slice, and must terminate: on the GPU an out-of-bounds access is a
sticky fault that takes the whole process down.  Round 1: 2700 cases, no sanitizer report."""
import ctypes, sys, numpy as np
sys.path.insert(0,'/root/repo')
from oracle import pyoracle
lib=ctypes.CDLL('/tmp/asan/libqdec3_host.so')
vp,cu,ci=ctypes.c_void_p,ctypes.c_uint,ctypes.c_int
lib.qdec3_host_decode.restype=ci; lib.qdec3_host_decode.argtypes=[vp,cu,vp,cu,vp,ci]
lib.qdec6_host_decode.restype=ci; lib.qdec6_host_decode.argtypes=[vp,cu,vp,cu,vp,ci]
lib.qfast_host_decode.restype=ci; lib.qfast_host_decode.argtypes=[vp,cu,vp,cu,vp]
lib.qadapt_host_decode.restype=ci; lib.qadapt_host_decode.argtypes=[vp,cu,vp,cu,vp]
gen=pyoracle.Gen(); ref=pyoracle.best()
rng=np.random.default_rng(123)
base={}
a=ref.bwt_encode(gen.text(2,200000))[1]
for coder in (1,2,3):
    r,s=ref.encode_block(a,coder=coder); base[coder]=s
skew=gen.skew(3,100000)
r,s=ref.encode_block(skew,coder=1); base['skew1']=s
def run(fn,stream,cap,*extra):
    s=np.ascontiguousarray(stream,dtype=np.uint8)
    out=np.empty(cap+64,np.uint8)
    return fn(s.ctypes.data,s.size,out.ctypes.data,cap,None,*extra)
decs=[("d3m0",lib.qdec3_host_decode,(0,),1),("d3m1",lib.qdec3_host_decode,(1,),1),("d3m2",lib.qdec3_host_decode,(2,),1),
      ("d6full",lib.qdec6_host_decode,(0,),1),("d6diet",lib.qdec6_host_decode,(1,),1),("d8full",lib.qdec6_host_decode,(2,),1),("d8diet",lib.qdec6_host_decode,(3,),1),("fast",lib.qfast_host_decode,(),3),("adapt",lib.qadapt_host_decode,(),2)]
count=0
for it in range(300):
    for name,fn,extra,coder in decs:
        src=base[coder] if it%3 else base.get('skew1') if coder==1 else base[coder]
        s=src.copy()
        mode=it%4
        if mode==0:   # random garbage
            s=rng.integers(0,256,rng.integers(8,5000),dtype=np.uint8)
        elif mode==1: # flip bytes
            for _ in range(rng.integers(1,20)): s[rng.integers(0,s.size)]^=rng.integers(1,256)
        elif mode==2: # truncate
            s=s[:rng.integers(8,s.size)]
        else:         # garbage tail after valid prefix
            k=rng.integers(8,s.size); s[k:]=rng.integers(0,256,s.size-k,dtype=np.uint8)
        cap=200000
        r=run(fn,s,cap,*extra); count+=1
print("fuzz cases run without sanitizer reports:",count)
