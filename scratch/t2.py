import sys, random, ctypes
sys.path.insert(0,'/root/repo')
import numpy as np
from oracle import pasta as o
L=ctypes.CDLL('/root/repo/scratch/host_shim.so')
random.seed(2)
def b32(x): return (ctypes.c_uint8*32).from_buffer_copy(int(x).to_bytes(32,'little'))
def b64(p): return (ctypes.c_uint8*64).from_buffer_copy((int(p[0]).to_bytes(32,'little')+int(p[1]).to_bytes(32,'little')) if p else bytes(64))
def fo(f,op,a,b=0):
    out=(ctypes.c_uint8*32)(); L.hs_field(f,op,b32(a),b32(b),out); return int.from_bytes(bytes(out),'little')
def po(c,op,a,b):
    out=(ctypes.c_uint8*64)(); L.hs_point(c,op,b64(a),b if not isinstance(b,(tuple,type(None))) else b64(b),out); r=bytes(out)
    x,y=int.from_bytes(r[:32],'little'),int.from_bytes(r[32:],'little'); return None if x==0 and y==0 else (x,y)
for f,m in [(0,o.P),(1,o.Q)]:
    edge=[0,1,m-1,m-2,(1<<256)%m,2**32-1,2**32,(1<<255)%m, m>>1]
    vals=edge+[random.randrange(m) for _ in range(300)]
    for a in vals:
        for b in random.sample(vals,6)+edge[:4]:
            assert fo(f,0,a,b)==(a+b)%m; assert fo(f,1,a,b)==(a-b)%m; assert fo(f,2,a,b)==a*b%m
        assert fo(f,4,a)==(-a)%m; assert fo(f,5,a)==a*a%m
    for a in vals[:20]: assert fo(f,3,a)==pow(a,m-2,m)
print('field ok')
for c,cv,G in [(0,o.VESTA,o.VESTA_GEN),(1,o.PALLAS,o.PALLAS_GEN)]:
    A=cv.mul(12345,G); B=cv.mul(99999,G)
    assert po(c,0,A,B)==cv.add(A,B); assert po(c,0,A,A)==cv.add(A,A); assert po(c,0,A,cv.neg(A)) is None
    assert po(c,0,A,None)==A; assert po(c,0,None,B)==B
    assert po(c,2,A,None)==cv.add(A,A); assert po(c,4,A,None)==cv.add(A,A)
    assert po(c,3,A,B)==cv.add(cv.mul(2,A),cv.mul(3,B)); assert po(c,3,A,A)==cv.mul(5,A)
    for k in [0,1,2,15,16,cv.fs-1,random.randrange(cv.fs),random.randrange(cv.fs)]:
        assert po(c,1,A,b32(k))==cv.mul(k,A),k
print('curve ok')
