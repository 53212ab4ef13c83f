"""What wall contacts are in practice (CPU, oracle only): a rollout of the restated reference under random actions,
and at a few episode ages every (ant capsule, wall box) pair within the capsule radius classified by body (torso /
Aux / lower leg), closest segment point (t = 0 end, t = 1 end, interior), contact feature (face / edge / corner) and
normal axis. The design of the lower leg's inline wall path (ant_physics.cuh tip_wall) rests on these numbers.
    python tools/wall_stats.py ant_heavenhell 2048 450"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import envs as oenvs, cstep, threefry as tf, brax_v1 as bx
from tests import _parity as P
kind=sys.argv[1]; n=int(sys.argv[2]); T=int(sys.argv[3])
oenv=oenvs.create(kind)
S=oenv.env.sys
cstep.attach(S, threads=os.cpu_count())
s=oenv.reset(P.keys_for(n,0))
rng=np.random.default_rng(0)
def classify(qp,t):
    nbx=len(S.boxes)
    b=np.repeat(S.cap_body,nbx); ca=np.repeat(S.cap_a,nbx,axis=0); cb=np.repeat(S.cap_b,nbx,axis=0); rad=np.repeat(S.cap_rad,nbx)
    box=np.tile(S.boxes,(len(S.cap_body),1))
    pos,rot=qp.pos[:,b],qp.rot[:,b]
    apos=qp.pos[:,S.arena][:,None,:]
    a_w=pos+bx.rotate(np.broadcast_to(ca,pos.shape),rot); b_w=pos+bx.rotate(np.broadcast_to(cb,pos.shape),rot)
    lo,hi=apos+box[:,:3],apos+box[:,3:]
    sp,bp=S._closest_segment_box(a_w,b_w,lo,hi)
    d=sp-bp; dist=np.sqrt((d**2).sum(-1))
    touch=(dist<rad)&(dist>0)
    inside=dist==0
    da=np.abs(sp-a_w).sum(-1)<1e-7; db=np.abs(sp-b_w).sum(-1)<1e-7
    nnz=(np.abs(d)>1e-9).sum(-1)
    axis=np.argmax(np.abs(d),-1)
    K=len(S.cap_body)
    touch_r=touch.reshape(n,K,nbx)
    print(f't={t} contacts per env {touch.sum()/n:.3f}; inside-box {inside.sum()/n:.3f}')
    for name,sel in (('torso',S.cap_body==0),('aux',np.isin(S.cap_body,[1,3,5,7])),('lower',np.isin(S.cap_body,[2,4,6,8]))):
        m=np.repeat(sel,nbx)
        tt=touch[:,m]
        if tt.sum()==0: print('  ',name,'none'); continue
        print(f'   {name}: contacts/env {tt.sum()/n:.3f} | end a(t=0) {(tt&da[:,m]).sum()/tt.sum():.2f} end b(t=1) {(tt&db[:,m]&~da[:,m]).sum()/tt.sum():.2f} interior {(tt&~da[:,m]&~db[:,m]).sum()/tt.sum():.2f} | face {(tt&(nnz[:,m]==1)).sum()/tt.sum():.2f} edge {(tt&(nnz[:,m]==2)).sum()/tt.sum():.2f} corner {(tt&(nnz[:,m]==3)).sum()/tt.sum():.2f} | axis x {(tt&(axis[:,m]==0)&(nnz[:,m]==1)).sum()/tt.sum():.2f} y {(tt&(axis[:,m]==1)&(nnz[:,m]==1)).sum()/tt.sum():.2f} z {(tt&(axis[:,m]==2)&(nnz[:,m]==1)).sum()/tt.sum():.2f}')
        per_body=touch_r[:,sel].sum(-1)
        print(f'      bodies with 1 contact {(per_body==1).mean():.4f}, with >=2 {(per_body>=2).mean():.4f}')
    # lower leg: both ends in contact with the same box?
for t in range(T+1):
    if t in (5,50,150,300,450): classify(s.qp,t)
    a=rng.uniform(-1,1,(n,8)).astype(np.float32)
    s=oenv.step(s,a)
