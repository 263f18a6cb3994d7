import sys, time, random
sys.path.insert(0,'/root/repo')
import numpy as np
from oracle import pasta as o, cpu as c
random.seed(1)
for f,m,root,z,d in [(c.FP,o.P,o.ROOT_P,o.ZETA_P,o.DELTA_P),(c.FQ,o.Q,o.ROOT_Q,o.ZETA_Q,o.DELTA_Q)]:
    k=c.field_consts(f,15)
    assert k['omega']==pow(root,1<<17,m) and k['delta']==d and k['zeta']==z and k['R']==(1<<256)%m and k['R2']==(1<<512)%m
    assert (k['inv64']*m)%(1<<64)==(1<<64)-1
    for _ in range(200):
        a,b=random.randrange(m),random.randrange(m)
        A,B=c.ints_to_bytes([a])[0],c.ints_to_bytes([b])[0]
        assert c.bytes_to_ints(c.field_op(f,0,A,B)[1])[0]==(a+b)%m
        assert c.bytes_to_ints(c.field_op(f,1,A,B)[1])[0]==(a-b)%m
        assert c.bytes_to_ints(c.field_op(f,2,A,B)[1])[0]==(a*b)%m
    a=random.randrange(m); A=c.ints_to_bytes([a])[0]
    assert c.bytes_to_ints(c.field_op(f,3,A)[1])[0]==pow(a,m-2,m)
    rc,s=c.field_op(f,4,c.ints_to_bytes([a*a%m])[0]); s=c.bytes_to_ints(s)[0]; assert rc==0 and s in (a,m-a)
    u=bytes(random.randrange(256) for _ in range(64)); assert c.bytes_to_ints(c.from_uniform(f,u))[0]==int.from_bytes(u,'little')%m
print('fields ok')
data=open('/root/reference/taiga_halo2/params/params_15','rb').read()
n=1<<15
t=time.time(); pts=c.decompress(c.VESTA, np.frombuffer(data[4:],np.uint8)); print('decompress',time.time()-t)
g=pts[:n]; gl=pts[n:2*n]; w=pts[2*n]; u=pts[2*n+1]
pr=o.read_params(data,limit=4)
for i in range(4):
    assert c.bytes_to_ints(g[i].reshape(2,32))==list(pr['g'][i]); assert c.bytes_to_ints(gl[i].reshape(2,32))==list(pr['g_lagrange'][i])
assert (c.compress(c.VESTA,pts).tobytes()==data[4:])
ones=c.ints_to_bytes([1]*n)
t=time.time(); r=c.msm(c.VESTA,ones,gl); print('msm',time.time()-t); assert r.tobytes()==g[0].tobytes()
wv=[1]; om=o.omega(15)
for i in range(n-1): wv.append(wv[-1]*om%o.P)
r=c.msm(c.VESTA,c.ints_to_bytes(wv),gl); assert r.tobytes()==g[1].tobytes()
r=c.msm(c.VESTA,c.ints_to_bytes([o.inv(n,o.P)]*n),g); assert r.tobytes()==gl[0].tobytes()
v=c.ints_to_bytes([random.randrange(o.P) for _ in range(n)])
t=time.time(); iv=c.ntt(c.FP,v,inverse=True); print('intt',time.time()-t)
t=time.time(); assert c.msm(c.VESTA,v,gl).tobytes()==c.msm(c.VESTA,iv,g).tobytes(); print('2 msm',time.time()-t)
a=[random.randrange(o.P) for _ in range(256)]
assert c.bytes_to_ints(c.ntt(c.FP,c.ints_to_bytes(a)))==o.ntt(a,o.omega(8))
assert c.bytes_to_ints(c.coeff_to_extended(8,5,c.ints_to_bytes(a)))==o.coeff_to_extended(a,8,10)
e=c.coeff_to_extended(8,5,c.ints_to_bytes(a)); assert c.bytes_to_ints(c.extended_to_coeff(8,5,e))[:256]==a
sc=[random.randrange(o.P) for _ in range(20)]; P=[pr['g'][i%4] for i in range(20)]
P=[o.VESTA.mul(i+1,p) for i,p in enumerate(P)]
ref=o.VESTA.msm(sc,P)
got=c.msm(c.VESTA,c.ints_to_bytes(sc),np.concatenate([c.ints_to_bytes(list(p)) for p in P]))
assert c.bytes_to_ints(got.reshape(2,32))==list(ref)
print('all ok')
